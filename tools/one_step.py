"""Two steps over one workload (for ncu captures):
python tools/one_step.py [n_reads] [decision|single|exact|nomarks] [v4|real|v3v4|mixed|ccs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "decision"   # decision | single | exact | nomarks
profile = sys.argv[3] if len(sys.argv) > 3 else "v4"
dev = torch.device("cuda", 0)
ctx = moira_b200.Context(0)
slab, lens, _ = synth.generate_device(profile, n, 20160106, dev)
stride, fixed = synth.DEVICE_LAYOUT[profile]
ee = torch.empty(n, dtype=torch.float64, device=dev); ns = torch.empty(n, dtype=torch.int32, device=dev)
fl = torch.empty(n, dtype=torch.uint8, device=dev); cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
marks = torch.zeros(n, dtype=torch.int32, device=dev)
d_len = None if lens is None else lens.data_ptr()
ctx.count_marks_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, marks.data_ptr(), 0, stream)
max_len = fixed or int(lens.max().item())
min_len = fixed or int(lens.min().item())
p = FilterParams(exact_ee=(mode == "exact"), cascade=2 if mode == "single" else 0, max_length=max_len, min_length=min_len,
                 length_sort=1 if profile == "mixed" else 0)
for _ in range(2):
    ctx.filter_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, p, ee.data_ptr(),
                      ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream, None if mode == "nomarks" else marks.data_ptr())
torch.cuda.synchronize()
print(cnt.cpu().numpy()[:10])
ctx.close()
