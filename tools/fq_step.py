"""Two moira_filter_fastq calls over ~260 MB of synthetic FASTQ text in pinned memory (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L

m, RL = 500_000, 253
slab, off, ln = synth.generate("v4", m, 20160106)
rows = slab.reshape(m, -1)
rec = np.empty((m, 10 + 1 + RL + 3 + RL + 1), dtype=np.uint8)
ids = np.char.zfill(np.arange(m).astype("U8"), 8)
rec[:, 0] = ord("@"); rec[:, 1] = ord("r")
rec[:, 2:10] = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(m, 8)
q = rows[:, :RL]; isn = q == 0xFF
rec[:, 10] = 10
rec[:, 11:11 + RL] = np.where(isn, ord("N"), ord("A"))
rec[:, 11 + RL:14 + RL] = np.frombuffer(b"\n+\n", dtype=np.uint8)
rec[:, 14 + RL:14 + 2 * RL] = np.where(isn, 2, q) + 33
rec[:, -1] = 10
h = moira_b200.PinnedBuffer(rec.size)
h.u8[:] = rec.reshape(-1)
ctx = moira_b200.Context(0)
for _ in range(2):
    res, lengths = ctx.filter_fastq(h.u8, FilterParams(exact_ee=False))
print(len(lengths), int(res.counters[L.CNT_ACCEPTED]), rec.size)
ctx.close()
