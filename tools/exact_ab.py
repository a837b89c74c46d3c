"""Exact-mode (and decision-mode) step times per workload, for A/B runs of library variants / environment switches:
python tools/exact_ab.py [profile:n ...]   (default: the five bench workloads)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L

work = sys.argv[1:] or ["v4:10000000", "real:10000000", "v3v4:10000000", "mixed:10000000", "ccs:2000000"]
dev = torch.device("cuda", 0)
ctx = moira_b200.Context(0)
peak = ctx.fp64_peak()[0]
stream = torch.cuda.current_stream().cuda_stream
for w in work:
    profile, n = w.split(":")
    n = int(n)
    slab, lens, _ = synth.generate_device(profile, n, 20160106, dev)
    stride, fixed = synth.DEVICE_LAYOUT[profile]
    ee = torch.empty(n, dtype=torch.float64, device=dev); ns = torch.empty(n, dtype=torch.int32, device=dev)
    fl = torch.empty(n, dtype=torch.uint8, device=dev)
    marks = torch.zeros(n, dtype=torch.int32, device=dev)
    d_len = None if lens is None else lens.data_ptr()
    ctx.count_marks_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, marks.data_ptr(), 0, stream)
    max_len = fixed or int(lens.max().item())
    min_len = fixed or int(lens.min().item())
    out = [profile, "n=%d" % n]
    for exact in (False, True):
        p = FilterParams(exact_ee=exact, max_length=max_len, min_length=min_len, uncert=float(os.environ.get("AB_UNCERT", "0.01")))
        best = 1e9
        for rep in range(4):
            cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.filter_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                              cnt.data_ptr(), stream, marks.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            if rep: best = min(best, e0.elapsed_time(e1))
        ctx.set_timing(True)   # one more call with the library's own events around the first-pass launches
        ctx.filter_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                          cnt.data_ptr(), stream, marks.data_ptr())
        torch.cuda.synchronize()
        first_ms, first_name = ctx.last_kernel_ms()
        ctx.set_timing(False)
        c = cnt.cpu().numpy()
        c = c // 2   # two calls added into these counters
        ops = float(c[L.CNT_FP64_OPS])
        out.append("%s %.3f ms %.4g reads/s  exec %.0f flop/read  frac %.3f  esc %.3f" % (
            "exact" if exact else "decision", best, n / best * 1e3, ops / n, ops / (best * 1e-3) / peak, c[L.CNT_ESCALATED] / n)
            + "  first pass %.3f ms (%s)" % (first_ms, first_name))
    print(" | ".join(out), flush=True)
    del slab, lens, ee, ns, fl, marks
    torch.cuda.empty_cache()
ctx.close()
