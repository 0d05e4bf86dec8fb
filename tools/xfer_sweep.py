#!/usr/bin/env python
"""Sweep of the host-buffer entry points (tsb_nq_evaluate / tsb_pfsp_evaluate): transfer mode x chunk size.
Prints one JSON line per configuration; run on the GPU box.  Used to pick the AUTO thresholds."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200"))


def one(kind, M, reps):
    import numpy as np

    import bench
    import tsb200
    if kind == "nq":
        par = bench.synth_nq_parents(17, min(M, 1 << 18), 1, tsb200.NQ_NODE_DTYPE)
        par = np.ascontiguousarray(np.resize(par, M))
        ev = tsb200.NQueensEvaluator(17, 1, M)
        out = np.empty(M * 17, dtype=np.uint8)
        call = lambda: ev.evaluate_gpu(par, M * 17, out)  # noqa: E731
    else:
        par = bench.synth_pfsp_parents(M, 1, tsb200.PFSP_NODE_DTYPE)
        ev = tsb200.PfspEvaluator(14, M=M)
        out = np.empty(M * 20, dtype=np.int32)
        call = lambda: ev.evaluate_gpu(par, M * 20, 1377, "lb1", out)  # noqa: E731
    for _ in range(5):
        call()
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"kind": kind, "M": M, "xfer": os.environ.get("TSB200_XFER", "auto"),
                      "pipe_min": os.environ.get("TSB200_PIPE_MIN"), "pipe_chunk": os.environ.get("TSB200_PIPE_CHUNK"),
                      "us_per_call": dt * 1e6, "Mnodes_s": M / dt / 1e6}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    for kind in ("nq", "pfsp"):
        for M, reps in ((50000, 400), (200000, 200), (1 << 20, 50), (1 << 22, 15)):
            if kind == "pfsp" and M > (1 << 20):
                continue
            for env in ({"TSB200_XFER": "memcpy", "TSB200_PIPE_MIN": "1000000000"},
                        {"TSB200_XFER": "memcpy", "TSB200_PIPE_MIN": "1", "TSB200_PIPE_CHUNK": "16384"},
                        {"TSB200_XFER": "memcpy", "TSB200_PIPE_MIN": "1", "TSB200_PIPE_CHUNK": "65536"},
                        {"TSB200_XFER": "memcpy", "TSB200_PIPE_MIN": "1", "TSB200_PIPE_CHUNK": "262144"},
                        {"TSB200_XFER": "zerocopy"}):
                subprocess.run([sys.executable, __file__, kind, str(M), str(reps)], env={**os.environ, **env})
