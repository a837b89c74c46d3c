#!/usr/bin/env python3
"""bench.py -- reads/s through the Poisson-binomial filter (BASELINE.json metric).

Own arm:  python bench.py [--gpus N] [--steps K] [--warmup W]
  One rank per GPU (torchrun for N > 1).  Headline workload = BASELINE config C2: 10 M synthetic 253-bp V4 contigs per
  GPU (MiSeq profile, weak scaling), --error_calc poisson_binomial, alpha 0.005, uncert 0.01, treat_as_errors.
  A step = one pass of the filter over the GPU's reads; the counters accumulate over the steps and are summed over the
  GPUs by ONE all-reduce (the library's own NCCL call) behind the last step, inside the timed region.
    value          reads/s, whole job, inputs resident in HBM (moira_filter_device, decision mode)
    roofline       the step's first-pass launches against the FP64 (non-fused DMUL/DADD) issue peak measured live with
                   moira_fp64_peak; frac = FP64 operations EXECUTED (device counter) / time / peak, never above 1
    roofline_exact the same for exact-ee mode (what --collapse needs), with the algorithmic flop of SURVEY 8d beside it
    e2e            the same metric through moira_filter_batch with pinned HOST byte slabs: H2D, kernels, D2H inside the
                   timed region; frac_of_link against a copy probe run at the same N
    configs        the other BASELINE.json configs (C3 ragged ~450 bp, C4 1500 bp exact + collapse, C5 mixed lengths in
                   three modes) and a real-profile workload bootstrapped from the reference's own contig fixture, each with
                   reads/s, executed-flop roofline fraction and a bit-exact parity sample against oracle/
    cpu_baseline   the unmodified reference C core (oracle/_ref) on this box's host cores, N = 1 only
Reference arm:  python bench.py --impl reference ...   times oracle/_ref on the host cores (no CUDA library is loaded).
"""
import argparse
import importlib.util
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
SEED = 20160105


def shared_config(reads_per_gpu):
    """The `config` object of BOTH arms (the driver compares them key by key)."""
    return {"workload": WORKLOAD, "reads_per_gpu": reads_per_gpu, "read_len": READ_LEN, "row_stride": STRIDE,
            "error_calc": "poisson_binomial", "alpha": ALPHA, "uncert": UNCERT, "ambigs": "treat_as_errors"}


def _load_synth_standalone():
    """moira_b200/synth.py without importing the package (whose __init__ loads the CUDA library): the reference arm
    must not map the product's .so."""
    spec = importlib.util.spec_from_file_location("moira_synth_standalone", os.path.join(ROOT, "moira_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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
    synth = _load_synth_standalone()
    cores = os.cpu_count() or 1
    probe = synth.generate("v4", 20000, SEED + 1)
    rate, _, kind, threads = _time_reference(*probe, threads=cores)
    per_step = int(min(2_000_000, max(20000, rate * 4.0)))          # ~4 s of CPU work per step
    slab, off, ln = synth.generate("v4", per_step, SEED + 1)
    for _ in range(args.warmup):
        _time_reference(slab[: 20000 * STRIDE], off[:20000], ln[:20000], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _time_reference(slab, off, ln, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d reads per step x %d steps of the C2 workload (same generator, seed %d)" % (per_step, args.steps, SEED + 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.reads),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


def collapse_bench(ctx, w, px, world, barrier, max_over_ranks):
    """C4's second half: --collapse (moira.py:459-475, 491-504) next to the exact-ee filter.  The sequences are resident in
    HBM like the slab; moira_collapse_device labels them exactly (hash + byte compare), the host turns labels + ee into
    the reference's groups (moira_collapse_labels).  Parity: the unmodified-semantics host collapse (moira_collapse: hash +
    memcmp over the strings) on a bounded sample of THIS rank's reads."""
    import torch
    import moira_b200
    n, stride, L_ = w.n, w.stride, w.fixed
    dev = w.slab.device
    stream = torch.cuda.current_stream().cuda_stream
    labels = torch.zeros(n, dtype=torch.int32, device=dev)

    def dedup():
        ctx.collapse_device(w.seqs.data_ptr(), None, None, stride, L_, n, labels.data_ptr(), 0, stream)

    dedup()
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 2
    e0.record()
    for _ in range(reps):
        w.run(px)
    e1.record()
    for _ in range(reps):
        dedup()
    e2.record()
    barrier()
    ms_filter, ms_dedup = max_over_ranks(e0.elapsed_time(e1)) / reps, max_over_ranks(e1.elapsed_time(e2)) / reps
    # groups / representatives / names order / abundance order from labels + ee
    ee_fin = (w.ee + w.ns.to(torch.float64)).contiguous()             # the value process_data returns (treat_as_errors)
    torch.cuda.synchronize()
    col = ctx.collapse_groups(labels.data_ptr(), ee_fin.data_ptr(), n)     # (the first call allocates the scratch)
    barrier()
    t0 = time.perf_counter()
    col = ctx.collapse_groups(labels.data_ptr(), ee_fin.data_ptr(), n)     # device passes + D2H of the six uint32 arrays (pinned)
    t_host = max_over_ranks(time.perf_counter() - t0)
    n_groups, largest = int(col.rep.shape[0]), int(col.size.max()) if n else 0
    dev_arrays = {f: np.array(getattr(col, f), dtype=np.uint64) for f in ("group_of_read", "rep", "size", "member_start", "members", "order")}
    t0 = time.perf_counter()
    lab_h = labels.cpu().numpy().view(np.uint32)
    ee_h = ee_fin.cpu().numpy()
    col_host = moira_b200.collapse_labels(lab_h, ee_h)                # the host version, D2H of labels + ee included
    t_hostgroups = max_over_ranks(time.perf_counter() - t0)
    groups_differ = sum(int(not np.array_equal(dev_arrays[f], getattr(col_host, f))) for f in dev_arrays)
    del dev_arrays
    # parity on a bounded sample: device labels of the sample alone -> groups == host hash + memcmp collapse
    k = min(n, 200_000)
    sub = w.seqs[:k].contiguous()
    sub_lab = torch.zeros(k, dtype=torch.int32, device=dev)
    ctx.collapse_device(sub.data_ptr(), None, None, stride, L_, k, sub_lab.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    seq_h = sub.cpu().numpy().reshape(-1)
    off = np.arange(k, dtype=np.uint64) * np.uint64(stride)
    t0 = time.perf_counter()
    want = moira_b200.collapse(seq_h, off, np.full(k, L_, np.uint32), ee_h[:k])
    t_hostcollapse = time.perf_counter() - t0
    got = moira_b200.collapse_labels(sub_lab.cpu().numpy().view(np.uint32), ee_h[:k])
    mism = sum(int(not np.array_equal(getattr(got, f), getattr(want, f))) for f in ("group_of_read", "rep", "size", "member_start", "members", "order"))
    total_ms = ms_filter + ms_dedup + t_host * 1e3
    return {"reads_per_gpu": n, "groups": n_groups, "largest_group": largest,
            "filter_exact_ms": ms_filter, "dedup_kernels_ms": ms_dedup, "groups_device_ms": t_host * 1e3,
            "groups_host_ms": t_hostgroups * 1e3, "device_groups_arrays_differing_from_host_groups": groups_differ,
            "dedup_kernels_value": world * n / (ms_dedup * 1e-3), "dedup_seq_gb_per_s": n * L_ * 2 / (ms_dedup * 1e-3) / 1e9,
            "exact_plus_collapse_value": world * n / (total_ms * 1e-3), "filter_only_value": world * n / (ms_filter * 1e-3),
            "slowdown_vs_filter_only": total_ms / ms_filter, "unit": "reads/s",
            "host_collapse_sample": {"reads": k, "value": k / t_hostcollapse, "unit": "reads/s",
                                    "note": "moira_collapse (hash + memcmp on all host threads) over the sample's strings"},
            "parity": {"sample_reads_per_rank": k, "arrays_differing_from_host_collapse": mism},
            "note": "device: 128-bit hash + exact byte compare of the HBM-resident sequences -> labels (moira_collapse_device); "
                    "groups / representatives / names order / abundance order from labels + ee on the device too "
                    "(moira_collapse_groups on the device-resident labels and ee: device passes + D2H of the six uint32 result arrays into "
                    "pinned host memory inside groups_device_ms; groups_host_ms is the host version, moira_collapse_labels, with its "
                    "D2H of labels + ee, for comparison)"}


def make_cli_fastq(slab_rows, seed):
    """FASTQ text for the CLI run: the qualities of `slab_rows` (uint8 [m, 256], C2 workload), sequences drawn from 100 000
    distinct random strings with Zipf(1.2) abundance ('N' where the slab has the marker), so that --collapse has the work
    it has on an amplicon run."""
    m = slab_rows.shape[0]
    rng = np.random.Generator(np.random.PCG64(seed))
    pool = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (100_000, READ_LEN))]
    wts = np.arange(1, 100_001, dtype=np.float64) ** -1.2
    cdf = np.cumsum(wts / wts.sum())
    pid = np.minimum(np.searchsorted(cdf, rng.random(m)), 99_999)
    rec = np.empty((m, 10 + 1 + READ_LEN + 3 + READ_LEN + 1), dtype=np.uint8)
    ids = np.char.zfill(np.arange(m).astype("U8"), 8)
    rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
    rec[:, 2:10] = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(m, 8)
    q = slab_rows[:, :READ_LEN]
    isn = q == 0xFF
    rec[:, 10] = 10
    rec[:, 11:11 + READ_LEN] = np.where(isn, ord("N"), pool[pid])
    rec[:, 11 + READ_LEN:14 + READ_LEN] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 14 + READ_LEN:14 + 2 * READ_LEN] = np.where(isn, 2, q) + 33
    rec[:, -1] = 10
    return rec


def cli_bench(rec, accepted_per_read, world, rank, args):
    """The program a user runs (python -m moira_b200.cli, in-process), wall clock from the FASTQ FILE to the output FILES,
    on all N GPUs from one process (--devices): default flags (exact ee, --collapse, fasta + qual + names) and the
    streaming case (-c False -o fastq).  Rank 0 drives; the other ranks wait."""
    if rank != 0:
        return None
    import io as _io
    import shutil
    import tempfile
    from moira_b200 import cli
    m = rec.shape[0]
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 6 * rec.nbytes else None
    tmp = tempfile.mkdtemp(prefix="moira_cli_", dir=base)
    out = {"reads": m, "fastq_bytes": int(rec.nbytes), "devices": world, "tmpdir": "tmpfs" if base else "disk", "unit": "reads/s"}
    try:
        path = os.path.join(tmp, "in.fastq")
        with open(path, "wb") as fh:
            fh.write(rec)
        # the ranks' GPUs by number; when that is every GPU of the box, "all" -- the CLI then takes as many as the input can keep
        # busy (one per 2 GiB of text: contexts cost more than they save on a file this size)
        import torch
        devs = "all" if world > 1 and world == torch.cuda.device_count() else ",".join(str(i) for i in range(world))
        for tag, extra in (("collapse_default", []), ("no_collapse_fastq", ["-c", "False", "-o", "fastq"])):
            dt = None
            for _rep in range(2):      # the faster of two runs: page-cache and driver effects (mmap faults, cudaFree) vary by seconds
                log_r = _io.StringIO()
                t0 = time.perf_counter()
                rc_r = cli.main(cli.parse_arguments(["-ffq", path, "-op", os.path.join(tmp, tag), "--devices", devs] + extra), log_r)
                dt_r = time.perf_counter() - t0
                if dt is None or dt_r < dt:
                    dt, rc, log = dt_r, rc_r, log_r
            files = {f: os.path.getsize(os.path.join(tmp, f)) for f in sorted(os.listdir(tmp)) if f.startswith(tag + ".")}
            used = log.getvalue().split(" s to the decisions, ")[1].split(" GPU")[0] if " s to the decisions, " in log.getvalue() else None
            res = {"value": m / dt, "seconds": dt, "rc": rc, "runs": 2, "output_bytes": int(sum(files.values())),
                   "devices_asked": devs, "devices_used": int(used) if used else None,
                   "decisions_seconds": float(log.getvalue().split(" s to the decisions")[0].rsplit("(", 1)[1]) if " s to the decisions" in log.getvalue() else None}
            if tag == "no_collapse_fastq":
                with open(os.path.join(tmp, tag + ".qc.good.fastq"), "rb") as fh:
                    good = fh.read().count(b"\n") // 4
                res["good_records"] = good
                res["matches_device_decisions"] = bool(good == accepted_per_read)
            else:
                with open(os.path.join(tmp, tag + ".qc.good.names"), "rb") as fh:
                    gn = fh.read()
                with open(os.path.join(tmp, tag + ".qc.bad.names"), "rb") as fh:
                    bn = fh.read()
                res["groups"] = gn.count(b"\n") + bn.count(b"\n")
                res["names_cover_every_read"] = bool(gn.count(b"r") + bn.count(b"r") - res["groups"] == m)
            out[tag] = res
            for f in files:
                os.remove(os.path.join(tmp, f))
        if world == 1:
            # compressed files in and out (moira.py:1065-1068 sniffed gzip input, --output_compression gz): the whole file as
            # blocked gzip (BGZF, what bgzip / Illumina's writers produce: moira_gz_inflate on all host threads), a 2 M-read
            # prefix as one plain gzip member (one inflate thread: a deflate stream cannot be split), and gzip outputs
            from moira_b200 import gz_deflate
            import zlib
            gzp = os.path.join(tmp, "in.bgzf.fastq.gz")
            fd = os.open(gzp, os.O_CREAT | os.O_WRONLY, 0o644)
            t0 = time.perf_counter()
            gz_bytes = gz_deflate(rec.reshape(-1), fd, 0, 1, 0, eof=True)     # level 1: this only prepares the input
            t_def = time.perf_counter() - t0
            os.close(fd)
            m_plain = min(m, 2_000_000)
            plain_p = os.path.join(tmp, "in.plain.fastq.gz")
            zc = zlib.compressobj(1, zlib.DEFLATED, 31)
            with open(plain_p, "wb") as fh:
                fh.write(zc.compress(rec[:m_plain].tobytes()))
                fh.write(zc.flush())
            comp = {"bgzf_input_bytes": int(gz_bytes), "bgzf_written_in_seconds": t_def,
                    "bgzf_deflate_level1_gb_per_s": rec.nbytes / t_def / 1e9}
            for tag, pth, mm_, extra in (("bgzf_in", gzp, m, []), ("plain_gzip_in", plain_p, m_plain, []),
                                         ("bgzf_in_gz_out", gzp, m, ["-oc", "gz"]), ("bgzf_in_gz_out_zlib6", gzp, m, ["-oc", "gz"])):
                if tag.endswith("zlib6"):
                    os.environ["MOIRA_B200_GZ_LEVEL"] = "6"      # zlib at bgzip's default level instead of the library's own compressor
                else:
                    os.environ.pop("MOIRA_B200_GZ_LEVEL", None)
                t0 = time.perf_counter()
                rc = cli.main(cli.parse_arguments(["-ffq", pth, "-op", os.path.join(tmp, "z_" + tag), "--devices", devs] + extra), _io.StringIO())
                dt = time.perf_counter() - t0
                files = {f: os.path.getsize(os.path.join(tmp, f)) for f in sorted(os.listdir(tmp)) if f.startswith("z_" + tag + ".")}
                comp[tag] = {"value": mm_ / dt, "unit": "reads/s", "reads": mm_, "seconds": dt, "rc": rc, "output_bytes": int(sum(files.values()))}
                for f in files:
                    os.remove(os.path.join(tmp, f))
            os.environ.pop("MOIRA_B200_GZ_LEVEL", None)
            os.remove(gzp); os.remove(plain_p)
            comp["note"] = ("default flags (exact ee, collapse, fasta + qual + names); bgzf_in: every 64 KB member inflated on its own thread; "
                            "plain_gzip_in: one member, one inflate thread (the library's own DEFLATE decoder, CRC-checked, zlib behind it), %d reads; gz_out: outputs as BGZF members compressed on all host threads by the library's own compressor (the CLI's default, every piece inflated and compared before it is written); zlib6: zlib at level 6 instead (MOIRA_B200_GZ_LEVEL=6)" % m_plain)
            out["compressed"] = comp
            # the paired flow (moira's main use): two FASTQ files -> contigs -> filter -> collapse -> files
            from tools.bench_contigs import make_pairs
            n_pairs, rl = 400_000, 251
            paths = []
            for tag, (bases, quals, _off, _ln) in zip(("R1", "R2"), make_pairs(n_pairs, rl)):
                recp = np.empty((n_pairs, 10 + 1 + rl + 3 + rl + 1), dtype=np.uint8)
                ids = np.char.zfill(np.arange(n_pairs).astype("U8"), 8)
                recp[:, 0] = ord("@"); recp[:, 1] = ord("p")
                recp[:, 2:10] = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(n_pairs, 8)
                recp[:, 10] = 10
                recp[:, 11:11 + rl] = bases.reshape(n_pairs, rl)
                recp[:, 11 + rl:14 + rl] = np.frombuffer(b"\n+\n", dtype=np.uint8)
                recp[:, 14 + rl:14 + 2 * rl] = quals.reshape(n_pairs, rl) + 33
                recp[:, -1] = 10
                pth = os.path.join(tmp, "pairs_%s.fastq" % tag)
                with open(pth, "wb") as fh:
                    fh.write(recp)
                paths.append(pth)
                del recp
            best = None
            for _rep in range(2):
                log = _io.StringIO()
                t0 = time.perf_counter()
                rc = cli.main(cli.parse_arguments(["-ffq", paths[0], "-rfq", paths[1], "--paired", "-op", os.path.join(tmp, "paired"),
                                                   "--devices", devs]), log)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, rc)
            files = {f: os.path.getsize(os.path.join(tmp, f)) for f in sorted(os.listdir(tmp)) if f.startswith("paired.")}
            out["paired_default"] = {"value": n_pairs / best[0], "unit": "pairs/s", "pairs": n_pairs, "seconds": best[0], "rc": best[1],
                                     "output_bytes": int(sum(files.values())),
                                     "workload": "2 x %d bp synthetic MiSeq V4 pairs, two FASTQ files -> contigs -> filter -> collapse -> fasta + qual + names + contigs report (second of two runs)" % rl}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out["value"] = out["collapse_default"]["value"]
    out["api"] = "python -m moira_b200.cli -ffq FILE --devices 0..N-1: mmap -> record-aligned shards -> moira_filter_fastq_ex per GPU and host thread -> (device labels -> groups) -> moira_format_records -> files"
    return out


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
    from oracle import py_oracle as po     # the CHECKER of the parity samples (and the cpu_baseline leg), never the product

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else None
    ctx = moira_b200.Context(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # the path's only collective lives in the library: rank 0's communicator id travels over torch.distributed (plumbing)
        box = [moira_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    stream = torch.cuda.current_stream().cuda_stream
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = FALLBACK_HBM_GBS, "fallback 6650 GB/s"
    if os.path.exists(peaks_path):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.time()

    def log(tag, obj):
        if rank == 0:
            print("[bench %6.1fs] %s: %s" % (time.time() - t_start, tag, json.dumps(obj)), file=sys.stderr, flush=True)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    class Work:
        """One device-resident workload: slab (+ lengths, + row marks) and its output arrays."""

        def __init__(self, profile, n, seed, with_sequences=False):
            self.profile, self.n = profile, n
            self.stride, self.fixed = synth.DEVICE_LAYOUT[profile]
            self.slab, self.lengths, self.seqs = synth.generate_device(profile, n, seed, dev, with_sequences=with_sequences)
            self.max_len = self.fixed or int(self.lengths.max().item())
            self.min_len = self.fixed or int(self.lengths.min().item())
            self.ee = torch.empty(n, dtype=torch.float64, device=dev)
            self.ns = torch.empty(n, dtype=torch.int32, device=dev)
            self.fl = torch.empty(n, dtype=torch.uint8, device=dev)
            self.cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
            # row marks (Ns | has-N per row): what every producer of a slab inside the library leaves next to it (6-bit
            # expansion, FASTQ conversion); here the stand-alone producer, once, like the generator itself
            self.marks = torch.zeros(n, dtype=torch.int32, device=dev)
            ctx.count_marks_device(self.slab.data_ptr(), None, None if self.lengths is None else self.lengths.data_ptr(),
                                   self.stride, self.fixed or 0, n, self.marks.data_ptr(), 0, stream)
            torch.cuda.synchronize()

        def params(self, **kw):
            return FilterParams(alpha=ALPHA, uncert=UNCERT, max_length=self.max_len, min_length=self.min_len, **kw)

        def run(self, params, marks=True):
            ctx.filter_device(self.slab.data_ptr(), None, None if self.lengths is None else self.lengths.data_ptr(), self.stride,
                              self.fixed or 0, self.n, params, self.ee.data_ptr(), self.ns.data_ptr(), self.fl.data_ptr(),
                              self.cnt.data_ptr(), stream, self.marks.data_ptr() if marks else None)

        def timed(self, params, steps, warm=3, marks=True, kernel_timing=False):
            """(ms total, launches, first-pass kernel ms total, kernel name, counters summed over ranks and steps)."""
            for _ in range(warm):
                self.run(params, marks)
            barrier()
            if kernel_timing:
                ctx.set_timing(True)
            self.cnt.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launch_count
            e0.record()
            for _ in range(steps):
                self.run(params, marks)
            if world > 1:   # the path's only collective: good / bad counts + error histogram, once per job
                ctx.reduce_counters_device(self.cnt.data_ptr(), stream)
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            launches = ctx.launch_count - l0
            kms, kname = (ctx.last_kernel_ms() if kernel_timing else (0.0, ""))
            if kernel_timing:
                ctx.set_timing(False)
            return ms, launches, kms, kname, self.cnt.cpu().numpy().astype(np.int64)

        def lens_host(self, sel):
            return np.full(len(sel), self.fixed, np.uint32) if self.lengths is None else self.lengths[sel].cpu().numpy().astype(np.uint32)

        def parity(self, k, exact_params, decision_params, calc="poisson_binomial"):
            """GPU vs oracle/ on the first and the last k/2 reads of THIS rank's workload: ee / Ns bit-exact in exact mode,
            decisions (and ee of the reads that are not flagged lower-bound) in decision mode."""
            k = min(k, self.n)
            sel = torch.cat([torch.arange(0, k // 2, device=dev), torch.arange(self.n - (k - k // 2), self.n, device=dev)])
            rows = self.slab[sel].cpu().numpy().reshape(-1)
            lens = self.lens_host(sel)
            offs = np.arange(k, dtype=np.uint64) * np.uint64(self.stride)
            if calc == "poisson_binomial":
                ee_o, ns_o = po.pb_batch(rows, offs, lens, ALPHA)
            else:
                ee_o, ns_o, _ = po.poisson_batch(rows, offs, lens, ALPHA)
            ok_o = (ee_o + ns_o) <= lens.astype(np.float64) * UNCERT
            out = {"sample_reads_per_rank": int(k), "checker": "oracle/pb_oracle.c" if calc == "poisson_binomial" else "oracle/py_oracle.py"}
            self.run(exact_params)
            torch.cuda.synchronize()
            ee_g, ns_g, fl_g = self.ee[sel].cpu().numpy(), self.ns[sel].cpu().numpy(), self.fl[sel].cpu().numpy()
            near = (fl_g & L.FLAG_NEAR_CUTOFF) != 0
            if calc == "poisson_binomial":
                out["ee_bit_mismatches"] = int((ee_g != ee_o).sum())
            else:   # device exp / pow against glibc: the tolerance north_star states
                out["ee_rel_err_max"] = float(np.max(np.abs(ee_g - ee_o) / np.maximum(np.abs(ee_o), 1e-300)))
                out["ee_outside_1e-12"] = int((np.abs(ee_g - ee_o) > 1e-12 * np.maximum(np.abs(ee_o), 1e-300)).sum())
            out["ns_mismatches"] = int((ns_g != ns_o).sum())
            out["decision_mismatches_outside_band"] = int((((fl_g & 1) != 0) != ok_o)[~near].sum())
            if decision_params is not None:
                self.run(decision_params)
                torch.cuda.synchronize()
                ee_d, fl_d = self.ee[sel].cpu().numpy(), self.fl[sel].cpu().numpy()
                lb = (fl_d & L.FLAG_LOWER_BOUND) != 0
                near_d = (fl_d & L.FLAG_NEAR_CUTOFF) != 0
                out["decision_mode_decision_mismatches"] = int((((fl_d & 1) != 0) != ok_o)[~near_d].sum())
                if calc == "poisson_binomial":
                    out["decision_mode_ee_bit_mismatches"] = int((ee_d[~lb] != ee_o[~lb]).sum())
                out["decision_mode_lower_bound_violations"] = int((ee_d[lb] > ee_o[lb]).sum())
            for key in list(out):
                if key.endswith("mismatches") or key.endswith("violations") or key.endswith("band") or key.endswith("1e-12"):
                    out[key] = int(sum_over_ranks(out[key]))
            return out

        def algorithmic_flop_exact(self):
            """SURVEY 8d for exact-ee mode: sum_r (L'-1)(3 K_r - 2) + (K_r - 1) + 6 with K_r = j*_r + 1, from the ee the
            last exact pass left on the device (j* = ceil(ee), 1 where ee == 0)."""
            lens = torch.full((self.n,), float(self.fixed), device=dev, dtype=torch.float64) if self.lengths is None else self.lengths.double()
            lp = (lens - self.ns.double()).clamp_(min=1)
            k = torch.ceil(self.ee).clamp_(min=1) + 1
            return float(((lp - 1) * (3 * k - 2) + (k - 1) + 6).sum().item())

        def algorithmic_flop_decision(self):
            """SURVEY 8d for decision mode: K_r = floor(cutoff_r) + 2, cutoff_r = L * uncert - Ns."""
            lens = torch.full((self.n,), float(self.fixed), device=dev, dtype=torch.float64) if self.lengths is None else self.lengths.double()
            nsd = self.ns.double()
            lp = (lens - nsd).clamp_(min=1)
            k = torch.floor((lens * UNCERT - nsd).clamp_(min=0)) + 2
            return float(((lp - 1) * (3 * k - 2) + (k - 1) + 6).sum().item())

    peak_ops, _ = ctx.fp64_peak(40000)
    peak_ops, _ = ctx.fp64_peak(40000)

    def fp64_view(counters, steps, ms_total, n_reads):
        """executed-FP64 roofline of `steps` passes: device counter / time / live peak."""
        ops = float(counters[L.CNT_FP64_OPS]) / world / steps                 # per GPU and step
        sec = ms_total * 1e-3 / steps
        return {"executed_flop_per_read": ops / n_reads, "achieved": ops / sec / 1e12, "peak": peak_ops / 1e12, "unit": "TFLOP/s",
                "frac": ops / sec / peak_ops}

    # ==== C2, the headline ===================================================================================
    n = args.reads
    c2 = Work("v4", n, SEED + 1 + 1000 * rank)
    p_dec, p_exact = c2.params(exact_ee=False), c2.params(exact_ee=True)
    c2.timed(p_dec, max(3, args.warmup), warm=0)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms, launches, kms, kname, counters = c2.timed(p_dec, args.steps, warm=0, kernel_timing=True)
    sampler.stop_flag = True
    sampler.join()
    value = world * n * args.steps / (ms * 1e-3)
    reads_total = max(1, int(counters[L.CNT_READS]))
    accepted_frac = float(counters[L.CNT_ACCEPTED]) / reads_total
    escalated_frac = float(counters[L.CNT_ESCALATED]) / reads_total
    # all-reduced counters against the sum of every rank's own flags (N-rank GPU parity of the collective)
    local_acc = float((c2.fl & 1).sum().item()) * args.steps
    counters_check = {"accepted_allreduced": int(counters[L.CNT_ACCEPTED]), "accepted_sum_of_rank_flags": int(sum_over_ranks(local_acc)),
                      "reads_allreduced": int(counters[L.CNT_READS]), "reads_expected": world * n * args.steps}
    counters_check["ok"] = (counters_check["accepted_allreduced"] == counters_check["accepted_sum_of_rank_flags"]
                            and counters_check["reads_allreduced"] == counters_check["reads_expected"])
    kernel_ms = kms / max(1, min(args.steps, 256))            # the library keeps the first 256 timed launch groups
    k_dec = synth.decision_k(READ_LEN, UNCERT)
    w_fp64, w_hbm = synth.w_fp64(READ_LEN, k_dec), synth.w_hbm(READ_LEN)
    exec_ops = float(counters[L.CNT_FP64_OPS]) / world / args.steps
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj[kname if kname in tj else "pb_tpr<K=4>"]["dram_bytes_per_read"] * n
    except Exception:
        pass
    # the same pass without the row marks (the sweep counts N/n itself), and the single K=4 sweep (cascade = 2)
    ms_sc, _, kms_sc, _, cnt_sc = c2.timed(p_dec, args.steps, marks=False, kernel_timing=True)
    p_single = c2.params(exact_ee=False, cascade=2)
    ms_s, _, kms_s, kname_s, cnt_s = c2.timed(p_single, args.steps, kernel_timing=True)
    kernel_ms_s = kms_s / max(1, min(args.steps, 256))
    exec_ops_s = float(cnt_s[L.CNT_FP64_OPS]) / world / args.steps
    hbm_achieved = n * w_hbm / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "fp64", "kernel": kname, "achieved": exec_ops / (kernel_ms * 1e-3) / 1e12, "peak": peak_ops / 1e12, "unit": "TFLOP/s",
        "frac": exec_ops / (kernel_ms * 1e-3) / peak_ops, "traffic": traffic,
        "peak_source": "moira_fp64_peak: register-resident non-fused DMUL/DADD probe, measured live in this run",
        "kernel_ms": kernel_ms, "executed_flop_per_read": exec_ops / n, "escalated_fraction": escalated_frac,
        "note": "achieved = FP64 operations the launches EXECUTED (device counter MOIRA_CNT_FP64_OPS: swept positions x operations "
                "per position) / CUDA-event time of the step's first-pass launches",
        "algorithmic_flop_per_read": w_fp64, "algorithmic_tflops": n * w_fp64 / (kernel_ms * 1e-3) / 1e12,
        "algorithmic_note": "SURVEY 8d credit (K=4 entries for every base, no credit for work avoided): may exceed the peak, not a hardware fraction",
        "single_sweep_kernel": kname_s, "single_sweep_value": world * n * args.steps / (ms_s * 1e-3),
        "single_sweep_frac": exec_ops_s / (kernel_ms_s * 1e-3) / peak_ops, "single_sweep_kernel_ms": kernel_ms_s,
        "slab_only_value": world * n * args.steps / (ms_sc * 1e-3),
        "slab_only_frac": float(cnt_sc[L.CNT_FP64_OPS]) / world / args.steps / (kms_sc / max(1, min(args.steps, 256)) * 1e-3) / peak_ops,
        "hbm_frac": hbm_achieved / hbm_peak, "hbm_achieved_gbs": hbm_achieved, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "hbm_bytes_per_read": w_hbm,
    }

    # ---- exact-ee mode (what --collapse needs), same inputs ----
    x_steps = max(3, args.steps // 4)
    ms_x, _, _, _, cnt_x = c2.timed(p_exact, x_steps, warm=2)
    value_exact = world * n * x_steps / (ms_x * 1e-3)
    alg_x = c2.algorithmic_flop_exact()
    roofline_exact = dict(fp64_view(cnt_x, x_steps, ms_x, n), bound="fp64", kernel="pb_cascade<2,4> + escalation ladder (whole step)",
                          value=value_exact, ms_per_step=ms_x / x_steps,
                          algorithmic_flop_per_read=alg_x / n, algorithmic_frac=alg_x / (ms_x * 1e-3 / x_steps) / peak_ops,
                          note="frac = executed FP64 operations / whole-step time / peak; algorithmic = sum_r (L'-1)(3K_r-2)+(K_r-1)+6 with "
                               "K_r = j*_r + 1 (SURVEY 8d, exact mode) from the device's own ee")
    parity = c2.parity(20000, p_exact, p_dec)
    parity["near_cutoff_reads_whole_run"] = int(counters[L.CNT_NEAR_CUTOFF])
    parity["counters_vs_rank_flags"] = counters_check

    # ==== host link probe (every rank at once: the rate a rank gets while all of them stream) ===============
    link = None
    if not args.no_e2e:
        barrier()
        h2d_gbs, d2h_gbs = ctx.link_probe(1 << 30, 3)
        link = {"h2d_gb_per_s_per_gpu": h2d_gbs, "d2h_gb_per_s_per_gpu": d2h_gbs, "h2d_gb_per_s_all_gpus": sum_over_ranks(h2d_gbs),
                "note": "1 GiB cudaMemcpyAsync from / to pinned host memory, best of 3, every rank at the same time"}

    # ==== end to end through the host-buffer C-ABI call ======================================================
    e2e = None
    if not args.no_e2e:
        h_slab = moira_b200.PinnedBuffer(n * STRIDE)
        torch.cuda.synchronize()
        torch.from_numpy(h_slab.u8).copy_(c2.slab.view(-1))
        h_out = moira_b200.PinnedBuffer(n * 13 + 4096)
        out = moira_b200.FilterResult(h_out.view(np.float64, n), h_out.view(np.int32, n, n * 8),
                                      h_out.view(np.uint8, n, n * 12), np.zeros(L.N_COUNTERS, np.uint64))
        h_meta = moira_b200.PinnedBuffer(n * 12)
        off = h_meta.view(np.uint64, n)
        off[:] = np.arange(n, dtype=np.uint64) * STRIDE
        lens = h_meta.view(np.uint32, n, n * 8)
        lens[:] = READ_LEN
        e_steps = max(2, min(5, args.steps))
        p_e2e = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False)

        def time_e2e(host_slab, params, before=None):
            for _ in range(2):                                     # warm-up (allocates device buffers)
                ctx.filter_batch(host_slab, off, lens, params, out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                if before is not None:
                    before()
                ctx.filter_batch(host_slab, off, lens, params, out)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            assert int(out.counters[L.CNT_READS]) == n, "e2e pass did not process every read"
            return dt

        d2h = n * 13 + L.N_COUNTERS * 8
        dt8 = time_e2e(h_slab.u8, p_e2e)
        acc8 = int(out.counters[L.CNT_ACCEPTED])
        # the library's 6-bit transport image of the same slab: 3/4 of the bytes cross PCIe, the device expands them (and leaves
        # the row marks) before filtering.  q6_prepacked: the host packer wrote the image once (like the slab itself);
        # q6_pack_in_loop: moira_pack_q6 of the whole slab inside every timed step.
        h_img = moira_b200.PinnedBuffer(n * STRIDE // 16 * 12)
        moira_b200.pack_q6(h_slab.u8, out=h_img.u8)
        p_q6 = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False, slab_format="q6")
        dt6 = time_e2e(h_img.u8, p_q6)
        assert int(out.counters[L.CNT_ACCEPTED]) == acc8, "6-bit transport changed the result"
        dt6p = time_e2e(h_img.u8, p_q6, before=lambda: moira_b200.pack_q6(h_slab.u8, out=h_img.u8))
        e2e = {"value": world * n * e_steps / dt8, "unit": "reads/s", "h2d_bytes_per_step": n * STRIDE, "d2h_bytes_per_step": d2h,
               "steps": e_steps, "ms_per_step": dt8 / e_steps * 1e3,
               "api": "moira_filter_batch: pinned host slab, one byte per base -> chunked H2D / kernels / D2H on two streams",
               "h2d_gb_per_s_per_gpu": n * STRIDE * e_steps / dt8 / 1e9,
               "frac_of_link": (n * STRIDE * e_steps / dt8 / 1e9) / link["h2d_gb_per_s_per_gpu"],
               "link_probe": link,
               "q6_prepacked": {"value": world * n * e_steps / dt6, "unit": "reads/s", "h2d_bytes_per_step": h_img.nbytes,
                                "ms_per_step": dt6 / e_steps * 1e3, "frac_of_link": (h_img.nbytes * e_steps / dt6 / 1e9) / link["h2d_gb_per_s_per_gpu"]},
               "q6_pack_in_loop": {"value": world * n * e_steps / dt6p, "unit": "reads/s", "ms_per_step": dt6p / e_steps * 1e3,
                                   "note": "moira_pack_q6 of the 2.56 GB slab (all host threads of the rank) inside every timed step"}}
        h_img.free()

    # ==== other --error_calc modes on the same resident slab (kernel-only, decision mode) =====================
    modes = {}
    for calc in ("poisson", "expected_error"):
        pm = c2.params(error_calc=calc, exact_ee=False)
        ms_m, _, _, _, _ = c2.timed(pm, 5, warm=2)
        modes[calc] = {"value": world * n * 5 / (ms_m * 1e-3), "unit": "reads/s"}

    # ==== end to end INCLUDING parsing: FASTQ text -> device parser -> filter =================================
    e2e_parse = None
    text_for_cli, cli_accepted = None, 0
    if not args.no_e2e:
        m = min(n, 2_000_000)
        rows = c2.slab[:m].cpu().numpy()
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
            r0 = ctx.filter_fastq(src, p_e2e, 33, out2)[0]
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                r0 = ctx.filter_fastq(src, p_e2e, 33, out2)[0]
            barrier()
            return max_over_ranks(time.perf_counter() - t0) / 3, r0

        dt, r0 = parse_rate(h_text.u8)
        dt_pageable, _ = parse_rate(text)
        e2e_parse = {"value": world * m / dt, "unit": "reads/s", "reads": m, "fastq_bytes": len(text),
                     "text_gb_per_s_per_gpu": len(text) / dt / 1e9, "frac_of_link": (len(text) / dt / 1e9) / link["h2d_gb_per_s_per_gpu"],
                     "accepted": int(r0.counters[L.CNT_ACCEPTED]),
                     "pageable_text": {"value": world * m / dt_pageable, "unit": "reads/s"},
                     "api": "moira_filter_fastq: FASTQ text in pinned host memory -> H2D as it is -> newline index, record table, slab "
                            "conversion (+ row marks) and the filter on the device -> D2H of ee / Ns / flags / lengths; three 64 MB chunks in flight"}
        h_text.free()
        del text
        if rank == 0 and not args.no_cli:
            m_cli = min(n, args.cli_reads)
            c2.run(p_e2e)
            torch.cuda.synchronize()
            cli_accepted = int((c2.fl[:m_cli] & 1).sum().item())
            text_for_cli = make_cli_fastq(c2.slab[:m_cli].cpu().numpy(), SEED + 9)

    # ==== the other BASELINE.json configs + the real-profile workload, device-resident =========================
    configs = {}
    c2_cfg = {"workload": WORKLOAD, "reads_per_gpu": n, "decision": {"value": value, "unit": "reads/s", "frac": roofline["frac"]},
              "exact_ee": {"value": value_exact, "unit": "reads/s", "frac": roofline_exact["frac"]},
              "accepted_fraction": accepted_frac, "escalated_fraction": escalated_frac, "parity": parity}
    configs["C2"] = c2_cfg
    log("C2", {"value": value, "roofline": roofline, "roofline_exact": roofline_exact, "parity": parity, "e2e": e2e, "modes": modes,
               "e2e_parse": e2e_parse})
    del c2
    torch.cuda.empty_cache()

    def run_config(name, profile, n_c, seed, what, steps=3, parity_k=4000, collapse=False, modes_c=()):
        w = Work(profile, n_c, seed, with_sequences=collapse)
        pd, px = w.params(exact_ee=False), w.params(exact_ee=True)
        res = {"workload": what, "reads_per_gpu": n_c, "mean_length": float(w.fixed or w.lengths.double().mean().item()),
               "row_stride": w.stride, "first_pass_K": synth.decision_k(w.max_len, UNCERT)}
        ms_d, _, _, _, cd = w.timed(pd, steps, warm=2)
        alg_d = w.algorithmic_flop_decision()
        res["decision"] = dict(fp64_view(cd, steps, ms_d, n_c), value=world * n_c * steps / (ms_d * 1e-3), ms_per_step=ms_d / steps,
                               algorithmic_flop_per_read=alg_d / n_c)
        res["accepted_fraction"] = float(cd[L.CNT_ACCEPTED]) / max(1, int(cd[L.CNT_READS]))
        res["escalated_fraction"] = float(cd[L.CNT_ESCALATED]) / max(1, int(cd[L.CNT_READS]))
        # classify-first batches (decisions / statistics that need 9 or more entries): the fp32 classifier is the first pass and
        # the ladder sweeps every read once with its own K
        res["classified_first_fraction"] = float(cd[L.CNT_CLASSIFIED]) / max(1, int(cd[L.CNT_READS]))
        ms_x2, _, _, _, cx = w.timed(px, steps, warm=1)
        alg_x2 = w.algorithmic_flop_exact()
        res["exact_ee"] = dict(fp64_view(cx, steps, ms_x2, n_c), value=world * n_c * steps / (ms_x2 * 1e-3), ms_per_step=ms_x2 / steps,
                               algorithmic_flop_per_read=alg_x2 / n_c, algorithmic_frac=alg_x2 / (ms_x2 * 1e-3 / steps) / peak_ops)
        for calc in modes_c:
            pm = w.params(error_calc=calc, exact_ee=False)
            ms_m2, _, _, _, cm = w.timed(pm, steps, warm=1)
            res[calc] = {"value": world * n_c * steps / (ms_m2 * 1e-3), "unit": "reads/s", "ms_per_step": ms_m2 / steps,
                         "accepted_fraction": float(cm[L.CNT_ACCEPTED]) / max(1, int(cm[L.CNT_READS])),
                         "hbm_frac": (n_c * (w.stride + 16) / (ms_m2 * 1e-3 / steps) / 1e9) / hbm_peak}
            if calc == "poisson":
                res[calc]["parity"] = w.parity(min(parity_k, 1000), w.params(error_calc="poisson", exact_ee=True), None, calc="poisson")
        res["parity"] = w.parity(parity_k, px, pd)
        if collapse:
            res["collapse"] = run_collapse(w, px)
        configs[name] = res
        log(name, res)
        del w
        torch.cuda.empty_cache()

    def run_collapse(w, px):
        """C4: exact ee for every read + dereplication (moira.py:459-475, 491-504)."""
        return collapse_bench(ctx, w, px, world, barrier, max_over_ranks)

    if not args.no_configs:
        big = not args.small_configs
        run_config("real_profile", "real", n, SEED + 7 + 1000 * rank,
                   "253-bp rows bootstrapped from the reference's real contig fixture (test_results/paired.qc.*, +-1 jitter): SURVEY 8d's preferred generator",
                   steps=max(3, args.steps // 4), parity_k=20000)
        run_config("C3", "v3v4", 25_000_000 if big else 2_000_000, SEED + 2 + 1000 * rank,
                   "C3: synthetic ~450-bp V3-V4 contigs (ragged 420..480 bp, lengths known to the host), %s reads per GPU" % ("25M" if big else "2M"),
                   steps=3, parity_k=4000)
        run_config("C4", "ccs", (5_000_000 if big else 500_000) // world, SEED + 3 + 1000 * rank,
                   "C4: 5M synthetic 1500-bp CCS-like reads in total (sharded over the GPUs), exact ee for every read + collapse",
                   steps=2, parity_k=20000 if big else 2000, collapse=True)
        run_config("C5", "mixed", 10_000_000 if big else 1_000_000, SEED + 4 + 1000 * rank,
                   "C5: mixed lengths 100..600 bp, V3-V4 shape; poisson_binomial vs poisson vs expected_error",
                   steps=3, parity_k=4000, modes_c=("poisson", "expected_error"))
        pb_acc, po_acc, ee_acc = (configs["C5"]["accepted_fraction"], configs["C5"]["poisson"]["accepted_fraction"],
                                  configs["C5"]["expected_error"]["accepted_fraction"])
        configs["C5"]["cross_mode_accept_delta"] = {"poisson_minus_pb": po_acc - pb_acc, "expected_error_minus_pb": ee_acc - pb_acc}

    # ==== the CLI a user runs: FASTQ file -> output files, wall clock (N GPUs in ONE process, rank 0 drives) ======
    e2e_cli = None
    if not args.no_e2e and not args.no_cli:
        torch.cuda.empty_cache()
        barrier()
        if text_for_cli is not None:
            e2e_cli = cli_bench(text_for_cli, cli_accepted, world, rank, args)
            log("e2e_cli", e2e_cli)
    text_for_cli = None
    barrier()

    # ==== the step before the filter when reads come in pairs: contig construction, N = 1 only ====
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
        p_pairs = FilterParams(alpha=ALPHA, uncert=UNCERT, exact_ee=False)
        ctx.filter_pairs(*pf, *pr_, ContigParams(), p_pairs, out=pout)
        ctx.set_timing(True)
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.filter_pairs(*pf, *pr_, ContigParams(), p_pairs, out=pout)
        dtp = (time.perf_counter() - t0) / 3
        cms, claunches = ctx.last_contig_ms()
        ctx.set_timing(False)
        contigs = {"kernel": {"value": np_pairs / (cms / 3 * 1e-3), "unit": "pairs/s", "cells_per_s": np_pairs * rl * rl / (cms / 3 * 1e-3)},
                   "e2e": {"value": np_pairs / dtp, "unit": "pairs/s"},
                   "workload": "%d synthetic 2 x %d bp MiSeq V4 pairs -> contigs -> filter (moira_filter_pairs); details: tools/bench_contigs.py" % (np_pairs, rl),
                   "contig_kernel_launches_per_step": claunches / 3, "bad_pairs": int(np.count_nonzero(pout.status))}
        del pf, pr_, pout, keep

    # ==== CPU baseline: the reference's own C core on this box's host cores (rank 0, N = 1) ======================
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        probe_n = 20000
        rows, offs, lns = synth.generate("v4", 2_000_000, SEED + 1)
        rate, _, kind, threads = _time_reference(rows, offs[:probe_n], lns[:probe_n], cores)
        s_n = int(min(2_000_000, max(probe_n, rate * 12.0)))      # ~12 s of CPU work
        rate, dt, kind, threads = _time_reference(rows, offs[:s_n], lns[:s_n], cores)
        cpu = {"value": rate, "unit": "reads/s", "cores": threads, "kind": kind,
               "sample": "%d reads of the C2 workload (host generator, seed %d), %.1f s, all host threads, C loop over the reference test()" % (s_n, SEED + 1, dt)}
        # the same core as moira itself calls it (moira.py:817): one Python call per read, list of ints, 1 core
        try:
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
            "config": shared_config(n),
            "run": {"mode": "decision (exact ee for accepted reads, lower bound for certain rejects)",
                    "first_pass": "cascade = 0 (library default): a pilot launch decides per batch between the two-entry sweep with Newton-bound rejects and the single K=4 sweep",
                    "row_marks": "given: Ns / has-N per row as the library's slab producers leave them (moira_count_marks_device here); roofline.slab_only_value is the pass that counts N/n in the sweep",
                    "l2": "inputs (2.56 GB/GPU) larger than L2; no flush", "accepted_fraction": accepted_frac,
                    "parallelism": "reads sharded by contiguous chunk, 1 rank per GPU; one all-reduce of 80 counters (moira_reduce_counters_device, NCCL) behind the last step",
                    "rank_cpu_affinity": numa_cpus},
            "roofline": roofline, "roofline_exact": roofline_exact, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "parity": parity, "configs": configs, "modes": modes, "e2e_parse": e2e_parse, "e2e_cli": e2e_cli,
            "contigs": contigs, "exact_ee": {"value": value_exact, "unit": "reads/s"},
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
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (C2)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip C3 / C4 / C5 / real_profile")
    ap.add_argument("--small-configs", action="store_true", help="C3 / C4 / C5 at a tenth of their size (development)")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--cli-reads", type=int, default=10_000_000, help="reads of the FASTQ file the CLI run filters")
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
