#!/usr/bin/env python3
"""Kernel-only throughput of the other BASELINE.json configs (C3 ~450 bp, C4 1500 bp, C5 mixed 100-600 bp)
next to C2, in decision / exact-ee / poisson / expected-error modes.  Device-resident slabs, CUDA events.
Not the headline bench (bench.py is); writes a markdown table to stdout."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L

def run(ctx, name, profile, n, seed, uncert=0.01, maxerrors=None, reps=5, length_sort=0, cascade=0):
    t0 = time.time()
    slab, off, ln = synth.generate(profile, n, seed)
    gen_s = time.time() - t0
    dev = torch.device("cuda", 0)
    d_slab = torch.from_numpy(slab).to(dev)
    uniform = bool((ln == ln[0]).all())
    stride = int(off[1] - off[0]) if n > 1 else 0
    fixed_pitch = bool((np.diff(off.astype(np.int64)) == stride).all())      # rows at a fixed pitch (lengths may vary)
    d_off = None if fixed_pitch else torch.from_numpy(off.astype(np.int64)).to(dev)
    d_len = None if uniform else torch.from_numpy(ln.astype(np.int32)).to(dev)
    ee = torch.empty(n, dtype=torch.float64, device=dev)
    ns = torch.empty(n, dtype=torch.int32, device=dev)
    fl = torch.empty(n, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    for label, p in (("pb decision", FilterParams(uncert=uncert, maxerrors=maxerrors, exact_ee=False, length_sort=length_sort, cascade=cascade)),
                     ("pb exact-ee", FilterParams(uncert=uncert, maxerrors=maxerrors, exact_ee=True, length_sort=length_sort, cascade=cascade)),
                     ("poisson", FilterParams(error_calc="poisson", uncert=uncert, maxerrors=maxerrors, exact_ee=False, length_sort=length_sort)),
                     ("expected_error", FilterParams(error_calc="expected_error", uncert=uncert, maxerrors=maxerrors, exact_ee=False, length_sort=length_sort))):
        def step():
            cnt.zero_()
            ctx.filter_device(d_slab.data_ptr(), d_off.data_ptr() if d_off is not None else None,
                              d_len.data_ptr() if d_len is not None else None, stride if fixed_pitch else 0,
                              int(ln[0]) if uniform else 0, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        c = cnt.cpu().numpy()
        rows.append((label, n / ms * 1e3, ms, c[L.CNT_ACCEPTED] / n, c[L.CNT_ESCALATED] / n))
    mean_len = float(ln.mean())
    for label, rps, ms, acc, esc in rows:
        print("| %s | %s | %d | %.0f | %.3g | %.3f | %.3f | %.3f |" % (name, label, n, mean_len, rps, ms, acc, esc))
    sys.stdout.flush()

if __name__ == "__main__":
    ctx = moira_b200.Context(0)
    print("| config | mode | reads | mean length | reads/s | ms/pass | accepted | escalated by the first pass |\n|---|---|---|---|---|---|---|---|")
    if "--quick" not in sys.argv:
        run(ctx, "C2 v4 253 bp", "v4", 2_000_000, 20160106)
        run(ctx, "C2 v4 253 bp, cascade=2 (single sweep)", "v4", 2_000_000, 20160106, cascade=2)
        run(ctx, "C3 v3v4 ~450 bp", "v3v4", 2_000_000, 20160107)
        run(ctx, "C3 v3v4 ~450 bp, cascade=1 (forced)", "v3v4", 2_000_000, 20160107, cascade=1)
        run(ctx, "C3 v3v4 ~450 bp, cascade=2 (single sweep)", "v3v4", 2_000_000, 20160107, cascade=2)
        run(ctx, "C4 ccs 1500 bp", "ccs", 200_000, 20160108)
    run(ctx, "C5 mixed 100-600 bp", "mixed", 1_000_000, 20160109)
    run(ctx, "C5 mixed, length_sort=1", "mixed", 1_000_000, 20160109, length_sort=1)
    run(ctx, "C3 v3v4, length_sort=1", "v3v4", 1_000_000, 20160107, length_sort=1)
    ctx.close()
