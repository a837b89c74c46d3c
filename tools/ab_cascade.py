"""A/B of the cascaded first pass at BASELINE config C2 (10 M x 253 bp, device-resident): reads/s per cascade setting."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ctx = moira_b200.Context(0)
slab = synth.generate_v4_device(n, 20160106, dev)
ee = torch.empty(n, dtype=torch.float64, device=dev); ns = torch.empty(n, dtype=torch.int32, device=dev)
fl = torch.empty(n, dtype=torch.uint8, device=dev); cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for exact in (False, True):
    for cascade in (2, 1, 0):
        p = FilterParams(exact_ee=exact, cascade=cascade)
        def step():
            ctx.filter_device(slab.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream)
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cnt.zero_(); e0.record()
        for _ in range(10): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        c = cnt.cpu().numpy()
        print("exact %d cascade %d: %.3f ms  %.3e reads/s  accepted %d lower %d" % (exact, cascade, ms, n / ms * 1e3, c[L.CNT_ACCEPTED] // 10, c[L.CNT_LOWER_BOUND] // 10), flush=True)
ctx.close()
