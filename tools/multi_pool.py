#!/usr/bin/env python
"""K independent device pools in ONE launch of the persistent kernel (tsb_nq_pool_run_multi):
python tools/multi_pool.py N M K [reps]   — the warm-up frontier dealt round-robin to K pools, all rounds, wall time"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200")]
import tsb200  # noqa: E402

N, M, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
m = 25
front, tree0, sol0 = tsb200.nqueens_warmup(N, K * m)
evs = [tsb200.NQueensEvaluator(N, M=M) for _ in range(K)]
for r in range(reps + 1):
    for i, ev in enumerate(evs):
        ev.pool_push(np.ascontiguousarray(front[i::K]))
    t0 = time.perf_counter()
    res = tsb200.nqueens_pool_run_multi(evs, m, M)
    dt = time.perf_counter() - t0
    rest = [ev.pool_drain() for ev in evs]  # fewer than m nodes each: finished on the host in the real driver
    tree = tree0 + sum(x[1] for x in res)
    print(f"N={N} M={M} K={K}: rounds {[x[0] for x in res]} parents {sum(x[1] for x in res)} solutions {sol0 + sum(x[3] for x in res)} "
          f"left {[len(x) for x in rest]}  {dt * 1e3:.1f} ms  {tree / dt / 1e9:.2f} Gnodes/s", flush=True)
for ev in evs:
    ev.close()
