"""Two steps over config C2 (for ncu captures): python tools/one_step.py [n_reads] [decision|single|exact|nomarks]"""
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
marks = torch.zeros(n, dtype=torch.int32, device=dev)
ctx.count_marks_device(slab.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n, marks.data_ptr(), 0, stream)
mode = sys.argv[2] if len(sys.argv) > 2 else "decision"   # decision | single | exact | nomarks
p = FilterParams(exact_ee=(mode == "exact"), cascade=2 if mode == "single" else 0)
for _ in range(2):
    ctx.filter_device(slab.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n, p, ee.data_ptr(),
                      ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream, None if mode == "nomarks" else marks.data_ptr())
torch.cuda.synchronize()
print(cnt.cpu().numpy()[:8])
ctx.close()
