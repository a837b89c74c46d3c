#!/usr/bin/env python3
"""bench.py -- reads/s through the Poisson-binomial filter (BASELINE.json metric).

Own arm:  python bench.py [--gpus N] [--steps K] [--warmup W]
  One rank per GPU (torchrun for N > 1).  Workload = BASELINE config C2: 10 M synthetic 253-bp V4
  contigs per GPU (MiSeq profile, weak scaling), --error_calc poisson_binomial, alpha 0.005,
  uncert 0.01, treat_as_errors.  A step = one pass of the filter over the GPU's 10 M reads.
    value   reads/s, whole job, inputs resident in HBM (moira_filter_device, decision mode)
    e2e     the same through moira_filter_batch with pinned HOST buffers: H2D of the slab, kernels,
            D2H of ee/Ns/flags/counters inside the timed region
    roofline       dominant kernel pb_tpr<K=4> against the FP64 (non-fused DMUL/DADD) issue peak
                   measured live with moira_fp64_peak, plus the HBM view
    cpu_baseline   the unmodified reference C core (oracle/_ref) on this box's host cores, N = 1 only
Reference arm:  python bench.py --impl reference ...   times oracle/_ref on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 253
STRIDE = 256
ALPHA = 0.005
UNCERT = 0.01
METRIC = "reads/s Poisson-binomial filter (253-bp)"
WORKLOAD = "C2: 10M synthetic 253-bp V4 contigs per GPU, MiSeq profile, poisson_binomial, alpha 0.005, uncert 0.01"
FALLBACK_HBM_GBS = 6650.0


def _cpu_sample(n_reads, seed):
    from moira_b200 import synth
    return synth.generate("v4", n_reads, seed)


def _time_reference(slab, off, ln, threads):
    """reads/s of the reference's own C core (bernoullimodule.c test()) over a packed sample."""
    from oracle import py_oracle as po
    t0 = time.perf_counter()
    if po.have_ref():
        po.ref_batch(slab, off, ln, ALPHA, n_threads=threads)
        kind = "reference"
    else:
        po.pb_batch(slab, off, ln, ALPHA, faithful=True)
        kind, threads = "port", 1
    dt = time.perf_counter() - t0
    return len(ln) / dt, dt, kind, threads


def _pool_chunk(chunk):
    """Pool worker: the reference module's per-read Python API over one chunk of (sequence, quality list) pairs."""
    from oracle import py_oracle as po
    ref = po.ref_module()
    seqs, quals = chunk
    return [ref.calculate_errors_PB(s_, q_, ALPHA) for s_, q_ in zip(seqs, quals)]


_POOL = None   # forked in main() before CUDA is initialised (the workers only ever run _pool_chunk)
_OUT = sys.stdout   # main() points this at the real stdout and file descriptor 1 at stderr


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    probe = _cpu_sample(20000, 20160105 + 1)
    rate, _, kind, threads = _time_reference(*probe, threads=cores)
    per_step = int(min(2_000_000, max(20000, rate * 4.0)))          # ~4 s of CPU work per step
    slab, off, ln = _cpu_sample(per_step, 20160105 + 1)
    for _ in range(args.warmup):
        _time_reference(slab[: 20000 * STRIDE], off[:20000], ln[:20000], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _time_reference(slab, off, ln, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d reads per step x %d steps of the C2 workload (same generator, seed 20160106)" % (per_step, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reads_per_step": per_step, "read_len": READ_LEN, "error_calc": "poisson_binomial"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, torch_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.ok = [], 0, False, False
        self.max_mhz = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # pragma: no cover
            self.err = repr(exc)

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                break
            time.sleep(0.002)

    def summary(self):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        reasons = [v for k, v in names.items() if self.reasons & k and v != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


def bind_to_gpu_numa(torch_index):
    """Multi-rank runs: pin this rank (and the pinned buffers it allocates, first touch) to the CPUs
    local to its GPU so that 8 ranks do not all stream host memory through one socket."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    import moira_b200
    from moira_b200 import FilterParams, synth
    from moira_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.reads
    ctx = moira_b200.Context(local_rank)

    # ---- synthetic input, resident in HBM (2.56 GB per GPU >> 126 MB L2: no flush needed) ----
    slab = synth.generate_v4_device(n, 20160105 + 1 + 1000 * rank, dev)
    ee = torch.empty(n, dtype=torch.float64, device=dev)
    ns = torch.empty(n, dtype=torch.int32, device=dev)
    fl = torch.empty(n, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    p_dec = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False)
    p_exact = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=True)

    def step(params):
        cnt.zero_()
        ctx.filter_device(slab.data_ptr(), None, None, STRIDE, READ_LEN, n, params, ee.data_ptr(), ns.data_ptr(),
                          fl.data_ptr(), cnt.data_ptr(), stream)
        if world > 1:   # the path's only collective: good/bad counts + error histogram
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(params, steps, timing=False):
        barrier()
        if timing:
            ctx.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        e0.record()
        for _ in range(steps):
            step(params)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        kms, kname = (ctx.last_kernel_ms() if timing else (0.0, ""))
        if timing:
            ctx.set_timing(False)
        return ms, launches, kms, kname

    for _ in range(max(3, args.warmup)):
        step(p_dec)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms, launches, kms, kname = timed(p_dec, args.steps, timing=True)
    sampler.stop_flag = True
    sampler.join()
    value = world * n * args.steps / (ms * 1e-3)
    counters = cnt.cpu().numpy().astype(np.int64)
    accepted_frac = float(counters[L.CNT_ACCEPTED]) / max(1, int(counters[L.CNT_READS]))

    # exact-ee mode (what --collapse needs), same inputs
    for _ in range(2):
        step(p_exact)
    ms_x, _, _, _ = timed(p_exact, max(3, args.steps // 4))
    value_exact = world * n * max(3, args.steps // 4) / (ms_x * 1e-3)

    # ---- parity on a bounded sample of THIS run's reads: GPU (exact mode, last step) vs the oracle ----
    parity = None
    if rank == 0:
        from oracle import py_oracle as po
        k = 20000   # the first and the last 10 000 reads (the last ones sit in the second ladder sub-batch)
        sel = torch.cat([torch.arange(0, k // 2, device=dev), torch.arange(n - k // 2, n, device=dev)])
        rows_s = slab[sel].cpu().numpy().reshape(-1)
        ee_o, ns_o = po.pb_batch(rows_s, np.arange(k, dtype=np.uint64) * STRIDE, np.full(k, READ_LEN, np.uint32), ALPHA)
        ee_g, ns_g, fl_g = ee[sel].cpu().numpy(), ns[sel].cpu().numpy(), fl[sel].cpu().numpy()
        ok_o = (ee_o + ns_o) <= READ_LEN * UNCERT
        near = (fl_g & L.FLAG_NEAR_CUTOFF) != 0
        parity = {"sample_reads": k, "ee_bit_mismatches": int((ee_g != ee_o).sum()), "ns_mismatches": int((ns_g != ns_o).sum()),
                  "decision_mismatches_outside_band": int((((fl_g & 1) != 0) != ok_o)[~near].sum()),
                  "near_cutoff_reads_whole_run": int(counters[L.CNT_NEAR_CUTOFF]), "checker": "oracle/pb_oracle.c"}

    # ---- roofline of the dominant kernel --------------------------------------------------------
    # The step's launches are one group, "pb_cascade<2,4>": pilot + verdict + two-entry sweep (+ the skipped full sweep)
    # + four-entry sweep over the escalated reads.  `achieved` follows SURVEY.md 8d: ALGORITHMIC flop per read (K = 4
    # entries for every base of every read, no credit for work avoided) x reads / measured time of the group.  The
    # cascade executes fewer FP64 operations than that, so this figure can exceed the pipe's peak; `executed` is the
    # operation count it really issues, and `single_sweep` is the K = 4 kernel on its own (cascade = 2), where
    # executed == algorithmic.
    k_dec = synth.decision_k(READ_LEN, UNCERT)
    kernel_ms = kms / max(1, min(args.steps, 256))   # the library keeps the first 256 timed launches
    w_fp64, w_hbm = synth.w_fp64(READ_LEN, k_dec), synth.w_hbm(READ_LEN)
    peak_ops, _ = ctx.fp64_peak(40000)
    achieved = n * w_fp64 / (kernel_ms * 1e-3)
    escalated_frac = float(counters[L.CNT_ESCALATED]) / max(1, int(counters[L.CNT_READS]))
    padded = (READ_LEN + 15) // 16 * 16
    # two-entry sweep: DSUB + 3 DMUL + DADD per swept position; escalated reads: 7 DMUL + 3 DADD per position again
    executed_per_read = 5 * padded + escalated_frac * 10 * padded if "cascade" in kname else 10 * padded
    p_single = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False, cascade=2)
    for _ in range(3):
        step(p_single)
    ms_s, _, kms_s, kname_s = timed(p_single, args.steps, timing=True)
    kernel_ms_s = kms_s / max(1, min(args.steps, 256))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = FALLBACK_HBM_GBS, "fallback"
    if os.path.exists(peaks_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    hbm_achieved = n * w_hbm / (kernel_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of the group's dominant kernel from the committed ncu --set full capture
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj[kname if kname in tj else "pb_tpr<K=4>"]["dram_bytes_per_read"] * n
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": kname, "achieved": achieved / 1e12, "peak": peak_ops / 1e12, "unit": "TFLOP/s",
        "frac": achieved / peak_ops, "traffic": traffic,
        "peak_source": "moira_fp64_peak: register-resident non-fused DMUL/DADD probe, measured live in this run",
        "flop_per_read": w_fp64, "kernel_ms": kernel_ms,
        "note": "achieved = algorithmic flop (SURVEY 8d: K=4 entries, every base, no credit for avoided work) / time of the step's "
                "launches; the cascade executes fewer operations (see executed), so frac may exceed 1",
        "executed": {"flop_per_read": executed_per_read, "escalated_fraction": escalated_frac,
                     "achieved": n * executed_per_read / (kernel_ms * 1e-3) / 1e12,
                     "frac": n * executed_per_read / (kernel_ms * 1e-3) / peak_ops},
        "single_sweep": {"kernel": kname_s, "value": world * n * args.steps / (ms_s * 1e-3), "kernel_ms": kernel_ms_s,
                         "achieved": n * w_fp64 / (kernel_ms_s * 1e-3) / 1e12, "frac": n * w_fp64 / (kernel_ms_s * 1e-3) / peak_ops,
                         "note": "cascade = 2: one K=4 sweep over every read; executed == algorithmic flop"},
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "bytes_per_read": w_hbm, "peak_source": hbm_src + " (MEASURED_PEAKS.json hbm_gbs)" if hbm_src == "measured" else "fallback 6650 GB/s"},
    }

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_slab = moira_b200.PinnedBuffer(n * STRIDE)
        torch.cuda.synchronize()
        h_t = torch.from_numpy(h_slab.u8)
        h_t.copy_(slab.view(-1))
        h_out = moira_b200.PinnedBuffer(n * 13 + 4096)
        out = moira_b200.FilterResult(h_out.view(np.float64, n), h_out.view(np.int32, n, n * 8),
                                      h_out.view(np.uint8, n, n * 12), np.zeros(L.N_COUNTERS, np.uint64))
        h_meta = moira_b200.PinnedBuffer(n * 12)
        off = h_meta.view(np.uint64, n)
        off[:] = np.arange(n, dtype=np.uint64) * STRIDE
        lens = h_meta.view(np.uint32, n, n * 8)
        lens[:] = READ_LEN
        e_steps = max(2, min(5, args.steps))

        def time_e2e(host_slab, params):
            ctx.filter_batch(host_slab, off, lens, params, out)         # warm-up (allocates device buffers)
            ctx.filter_batch(host_slab, off, lens, params, out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                ctx.filter_batch(host_slab, off, lens, params, out)
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            assert int(out.counters[L.CNT_READS]) == n, "e2e pass did not process every read"
            return dt

        dt8 = time_e2e(h_slab.u8, p_dec)
        acc8 = int(out.counters[L.CNT_ACCEPTED])
        e2e_q8 = {"value": world * n * e_steps / dt8, "unit": "reads/s", "h2d_bytes_per_step": n * STRIDE,
                  "d2h_bytes_per_step": n * 13 + L.N_COUNTERS * 8, "steps": e_steps, "ms_per_step": dt8 / e_steps * 1e3,
                  "api": "moira_filter_batch, one byte per base"}
        # the library's 6-bit transport image of the same slab (moira_pack_q6, packed once by the host packer
        # like the slab itself): 3/4 of the bytes cross PCIe, the device expands them before filtering
        h_img = moira_b200.PinnedBuffer(n * STRIDE // 16 * 12)
        moira_b200.pack_q6(h_slab.u8, out=h_img.u8)
        p_q6 = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False, slab_format="q6")
        dt6 = time_e2e(h_img.u8, p_q6)
        assert int(out.counters[L.CNT_ACCEPTED]) == acc8, "6-bit transport changed the result"
        e2e = {"value": world * n * e_steps / dt6, "unit": "reads/s", "h2d_bytes_per_step": h_img.nbytes,
               "d2h_bytes_per_step": n * 13 + L.N_COUNTERS * 8, "steps": e_steps, "ms_per_step": dt6 / e_steps * 1e3,
               "api": "moira_filter_batch (pinned host slab in the 6-bit transport format -> chunked H2D / device expand / "
                      "kernels / D2H on two streams; uniform rows: offsets/lengths stay on the host)",
               "q8": e2e_q8}

    # ---- other --error_calc modes on the same resident slab (kernel-only, decision mode) --------------
    modes = {}
    for calc in ("poisson", "expected_error"):
        pm = FilterParams(error_calc=calc, alpha=ALPHA, uncert=UNCERT, exact_ee=False)
        for _ in range(2):
            step(pm)
        ms_m, _, _, _ = timed(pm, 5)
        modes[calc] = {"value": world * n * 5 / (ms_m * 1e-3), "unit": "reads/s"}

    # ---- end to end INCLUDING host parsing: FASTQ text -> C parser (all host threads) -> filter ------
    e2e_parse = None
    if not args.no_e2e:
        m = min(n, 2_000_000)
        rows = slab[:m].cpu().numpy()
        rec = np.empty((m, 10 + 1 + READ_LEN + 3 + READ_LEN + 1), dtype=np.uint8)
        ids = np.char.zfill(np.arange(m).astype("U8"), 8)
        rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
        rec[:, 2:10] = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(m, 8)
        q = rows[:, :READ_LEN]
        isn = q == 0xFF
        rec[:, 10] = 10
        rec[:, 11:11 + READ_LEN] = np.where(isn, ord("N"), ord("A"))
        rec[:, 11 + READ_LEN:14 + READ_LEN] = np.frombuffer(b"\n+\n", dtype=np.uint8)
        rec[:, 14 + READ_LEN:14 + 2 * READ_LEN] = np.where(isn, 2, q) + 33
        rec[:, -1] = 10
        text = rec.tobytes()
        del rec, rows, q, isn, ids
        h_out2 = moira_b200.PinnedBuffer(m * 13 + 4096)
        out2 = moira_b200.FilterResult(h_out2.view(np.float64, m), h_out2.view(np.int32, m, m * 8),
                                       h_out2.view(np.uint8, m, m * 12), np.zeros(L.N_COUNTERS, np.uint64))

        h_text = moira_b200.PinnedBuffer(len(text))
        h_text.u8[:] = np.frombuffer(text, dtype=np.uint8)

        def parse_rate(src):
            r0 = ctx.filter_fastq(src, p_dec, 33, out2)[0]
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                r0 = ctx.filter_fastq(src, p_dec, 33, out2)[0]
            barrier()
            return (time.perf_counter() - t0) / 3, r0

        dt, r0 = parse_rate(h_text.u8)
        dt_pageable, _ = parse_rate(text)
        e2e_parse = {"value": world * m / dt, "unit": "reads/s", "reads": m, "fastq_bytes": len(text),
                     "text_gb_per_s": len(text) / dt / 1e9, "host_threads": os.cpu_count(),
                     "accepted": int(r0.counters[L.CNT_ACCEPTED]),
                     "pageable_text": {"value": world * m / dt_pageable, "unit": "reads/s",
                                       "note": "the same with the text in ordinary (pageable) host memory: staged through pinned buffers by the host threads"},
                     "api": "moira_filter_fastq: FASTQ text in pinned host memory -> H2D as it is (chunk cuts guessed from the local structure of the text, "
                            "verified by the device's line count) -> newline index, record table, slab conversion and the filter on the device -> "
                            "D2H of ee / Ns / flags / lengths; three 64 MB chunks in flight"}
        del text
        h_text.free()

    # ---- the step before the filter when reads come in pairs: contig construction, N = 1 only ----
    contigs = None
    if world == 1 and not args.no_e2e:
        from tools.bench_contigs import make_pairs
        from moira_b200 import ContigParams, PairResult, PinnedBuffer
        np_pairs, rl = 1 << 18, 251
        pf, pr_ = make_pairs(np_pairs, rl)
        keep = []

        def _pin(arr):
            pb = PinnedBuffer(arr.nbytes)
            keep.append(pb)
            v = pb.view(arr.dtype, arr.size)
            v[:] = arr
            return v

        pf = (_pin(pf[0]), _pin(pf[1]), pf[2], pf[3])
        pr_ = (_pin(pr_[0]), _pin(pr_[1]), pr_[2], pr_[3])
        pout = PairResult.allocate(np_pairs, (2 * rl + 15) // 16 * 16, True, pinned=True)
        ctx.filter_pairs(*pf, *pr_, ContigParams(), p_dec, out=pout)
        ctx.set_timing(True)
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.filter_pairs(*pf, *pr_, ContigParams(), p_dec, out=pout)
        dtp = (time.perf_counter() - t0) / 3
        cms, claunches = ctx.last_contig_ms()
        ctx.set_timing(False)
        contigs = {"kernel": {"value": np_pairs / (cms / 3 * 1e-3), "unit": "pairs/s", "cells_per_s": np_pairs * rl * rl / (cms / 3 * 1e-3)},
                   "e2e": {"value": np_pairs / dtp, "unit": "pairs/s"},
                   "workload": "%d synthetic 2 x %d bp MiSeq V4 pairs -> contigs -> filter (moira_filter_pairs); details: tools/bench_contigs.py, profiles/r01_contigs.json" % (np_pairs, rl),
                   "contig_kernel_launches_per_step": claunches / 3, "bad_pairs": int(np.count_nonzero(pout.status))}
        del pf, pr_, pout, keep

    # ---- CPU baseline: the reference's own C core on this box's host cores (rank 0, N = 1) ------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        probe_n = 20000
        rows = slab[:2_000_000].cpu().numpy().reshape(-1)
        offs = np.arange(2_000_000, dtype=np.uint64) * STRIDE
        lns = np.full(2_000_000, READ_LEN, dtype=np.uint32)
        rate, _, kind, threads = _time_reference(rows, offs[:probe_n], lns[:probe_n], cores)
        s_n = int(min(2_000_000, max(probe_n, rate * 12.0)))      # ~12 s of CPU work
        rate, dt, kind, threads = _time_reference(rows, offs[:s_n], lns[:s_n], cores)
        cpu = {"value": rate, "unit": "reads/s", "cores": threads, "kind": kind,
               "sample": "first %d reads of this run's GPU workload, %.1f s, all host threads, C loop over the reference test()" % (s_n, dt)}
        # the same core as moira itself calls it (moira.py:817): one Python call per read, list of ints, 1 core
        try:
            from oracle import py_oracle as po
            if po.have_ref():
                ref = po.ref_module()
                k = 20000
                seqs = ["".join("N" if v == 0xFF else "A" for v in r[:READ_LEN]) for r in rows.reshape(-1, STRIDE)[:k].tolist()]
                quals = [[2 if v >= 0xFE else (1 if v == 0 else v) for v in r[:READ_LEN]] for r in rows.reshape(-1, STRIDE)[:k].tolist()]
                t0 = time.perf_counter()
                for s_, q_ in zip(seqs, quals):
                    ref.calculate_errors_PB(s_, q_, ALPHA)
                cpu["python_api_1core"] = {"value": k / (time.perf_counter() - t0), "unit": "reads/s",
                                           "sample": "%d reads through bernoulli.calculate_errors_PB(str, list, float), 1 core" % k}
                # ... and as moira runs it on a multi-core host: multiprocessing.Pool, chunked map (moira.py:398-399, 431-438)
                if _POOL is not None:
                    per, rep = 2000, 8 * cores
                    chunks = [(seqs[i:i + per], quals[i:i + per]) for i in range(0, k, per)] * (rep * per // k + 1)
                    chunks = chunks[:rep]
                    _POOL.map(_pool_chunk, chunks[:cores])                       # warm-up: module import in every worker
                    t0 = time.perf_counter()
                    got = _POOL.map(_pool_chunk, chunks, chunksize=1)
                    dtp = time.perf_counter() - t0
                    cpu["python_api_pool"] = {"value": sum(len(g) for g in got) / dtp, "unit": "reads/s", "cores": cores,
                                              "sample": "%d reads in chunks of %d through multiprocessing.Pool(%d).map, reads pickled to the workers"
                                                        % (rep * per, per, cores)}
        except Exception as exc:  # pragma: no cover
            cpu["python_api_1core"] = {"error": repr(exc)}
        finally:
            if _POOL is not None:
                _POOL.terminate()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reads_per_gpu": n, "read_len": READ_LEN, "row_stride": STRIDE,
                       "error_calc": "poisson_binomial", "mode": "decision (exact ee for accepted reads, lower bound for certain rejects)",
                       "first_pass": "cascade = 0 (library default): pilot launch decides per batch between the two-entry sweep with Newton-bound rejects and the single K=4 sweep",
                       "l2": "inputs (2.56 GB/GPU) larger than L2; no flush", "accepted_fraction": accepted_frac,
                       "parallelism": "reads sharded by contiguous chunk, 1 rank per GPU, NCCL all-reduce of 80 counters per step",
                       "rank_cpu_affinity": numa_cpus},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "parity": parity, "modes": modes, "e2e_parse": e2e_parse, "contigs": contigs,
            "exact_ee": {"value": value_exact, "unit": "reads/s", "note": "exact statistic for every read (escalation ladder), device-resident"},
        }
        print(json.dumps(line), file=_OUT, flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    # NCCL_DEBUG is set on the box), so file descriptor 1 is pointed at stderr for the run and the line goes to the real one.
    global _POOL, _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl != "reference" and int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu:
        try:
            import multiprocessing as mp
            from oracle import py_oracle as po
            if po.have_ref():
                _POOL = mp.get_context("fork").Pool(os.cpu_count() or 1)   # before any CUDA call: forking is safe here
        except Exception:  # pragma: no cover
            _POOL = None
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
