#!/usr/bin/env python3
"""Host-side measurement of moira_collapse against the reference's dict loop (moira.py:459-475) on a
C4-shaped input: 1500-bp sequences drawn from a pool with Zipf(1.2) abundance."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moira_b200

n, pool_n, L = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 50_000, 1500
rng = np.random.default_rng(8)
pool = rng.integers(0, 4, size=(pool_n, L), dtype=np.uint8)
pool = np.frombuffer(b"ACGT", dtype=np.uint8)[pool]
idx = (rng.zipf(1.2, size=n) - 1) % pool_n
text = pool[idx].tobytes()
ee = rng.random(n) * 20
off = np.arange(n, dtype=np.uint64) * L
ln = np.full(n, L, dtype=np.uint32)
t0 = time.perf_counter()
col = moira_b200.collapse(text, off, ln, ee)
t_native = time.perf_counter() - t0
seqs = [text[i * L:(i + 1) * L] for i in range(n)]
t0 = time.perf_counter()
uniq = {}
for r in range(n):
    s = seqs[r]
    u = uniq.get(s)
    if u is None:
        uniq[s] = [r, ee[r], [r]]
    elif ee[r] < u[1]:
        u[0], u[1] = r, ee[r]
        u[2].insert(0, r)
    else:
        u[2].append(r)
order = sorted(uniq, key=lambda s: len(uniq[s][2]), reverse=True)
t_py = time.perf_counter() - t0
assert len(order) == len(col.order) and all(int(col.rep[g]) == uniq[s][0] for g, s in zip(col.order.tolist(), order))
print("collapse of %d x %d-bp reads into %d uniques: native %.3f s (%.3g reads/s, %d threads) | python dict loop %.3f s (%.3g reads/s) | identical representatives"
      % (n, L, len(order), t_native, n / t_native, os.cpu_count(), t_py, n / t_py))
