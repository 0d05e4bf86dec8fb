#!/usr/bin/env python
"""A few launches of one kernel family with device-resident inputs, for ncu captures (profiles/README.md):
python tools/prof_run.py nq | lb1 | lb2 | expand | rounds | pool"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200")]
import bench  # noqa: E402
import tsb200  # noqa: E402

what = sys.argv[1]
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
if what in ("nq", "expand"):
    N, M = 17, 1 << 22
    par = bench.synth_nq_parents(N, M, 7, tsb200.NQ_NODE_DTYPE)
    d_in = [torch.from_numpy(np.roll(par, 1000 * k, axis=0).view(np.uint8).reshape(-1).copy()).to(dev) for k in range(3)]
    with tsb200.NQueensEvaluator(N, 1, M) as ev:
        if what == "nq":
            d_lab = torch.empty(M * N, dtype=torch.uint8, device=dev)
            for k in range(5):
                ev.evaluate_device(d_in[k % 3].data_ptr(), M, d_lab.data_ptr(), stream.cuda_stream)
        else:
            d_ch = torch.empty(M * 21 * 3, dtype=torch.uint8, device=dev)
            for k in range(4):
                ev.expand_device(d_in[k % 3].data_ptr(), M, d_ch.data_ptr())
        torch.cuda.synchronize()
elif what in ("lb1", "lb2"):
    inst, M, hist, best = (14, 1 << 20, bench.PFSP_TA014_LB1_HIST, 1377) if what == "lb1" else \
        (20, 1 << 18, bench.PFSP_TA020_LB2_HIST, 1591)
    par = bench.synth_pfsp_parents(M, 1234, tsb200.PFSP_NODE_DTYPE, hist)
    d_in = [torch.from_numpy(np.roll(par, 7919 * k, axis=0).view(np.uint8).reshape(-1).copy()).to(dev) for k in range(3)]
    d_out = torch.empty(M * 80, dtype=torch.uint8, device=dev)
    with tsb200.PfspEvaluator(inst, M=M) as ev:
        for k in range(5):
            ev.evaluate_device(what, d_in[k % 3].data_ptr(), M, best, d_out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
elif what == "rounds":
    os.environ["TSB200_POOLS"] = "1"
    st = tsb200.nqueens_search_device(15, 1, 25, 50000, 1)  # 3 431 rounds in one launch of the persistent kernel
    assert (st.explored_tree, st.explored_sol) == (171129071, 2279184)
elif what == "rounds17":  # four pools per launch (the default of the search driver), N = 17: 2 048 rounds per pool and launch
    st = tsb200.nqueens_search_device(17, 1, 25, 50000, 1)
    assert (st.explored_tree, st.explored_sol) == (8017021931, 95815104)
elif what == "pool":
    st = tsb200.nqueens_search_device(17, 1, 25, 1 << 22, 1)  # 1 922 two-kernel rounds of up to 4 Mi parents
    assert (st.explored_tree, st.explored_sol) == (8017021931, 95815104)
print("done", what)
