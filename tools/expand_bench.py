#!/usr/bin/env python
"""Times the fused expand kernel (tsb_nq_expand_device) on one synthetic chunk, device resident:
usage: python tools/expand_bench.py [N] [M] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200")]
import bench  # noqa: E402
import tsb200  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 17
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda:0")
par = bench.synth_nq_parents(N, M, 7, tsb200.NQ_NODE_DTYPE)
nsets = 3
d_in = [torch.from_numpy(np.roll(par, 1000 * k, axis=0).view(np.uint8).reshape(-1).copy()).to(dev) for k in range(nsets)]
d_ch = [torch.empty(M * 21 * 3, dtype=torch.uint8, device=dev) for _ in range(nsets)]
d_lab = torch.empty(M * N, dtype=torch.uint8, device=dev)
with tsb200.NQueensEvaluator(N, 1, M) as ev:
    for k in range(3):
        nc, ns = ev.expand_device(d_in[k % nsets].data_ptr(), M, d_ch[k % nsets].data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        nc, ns = ev.expand_device(d_in[k % nsets].data_ptr(), M, d_ch[k % nsets].data_ptr())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"expand N={N} M={M}: {dt * 1e6:.1f} us/call (host wall incl. launch+sync), children={nc} sol={ns}, "
          f"{M / dt / 1e9:.2f} G parents/s, alg bytes {(M * 21 + nc * 21) / dt / 1e9:.0f} GB/s")
    s = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        for k in range(3):
            ev.evaluate_device(d_in[k % nsets].data_ptr(), M, d_lab.data_ptr(), s.cuda_stream)
        e0.record(s)
        for k in range(reps):
            ev.evaluate_device(d_in[k % nsets].data_ptr(), M, d_lab.data_ptr(), s.cuda_stream)
        e1.record(s)
    torch.cuda.synchronize()
    print(f"evaluate: {e0.elapsed_time(e1) / reps * 1e3:.1f} us/call")
