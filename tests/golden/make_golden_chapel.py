#!/usr/bin/env python
"""Pin the ONE place where the Chapel program and the reference's C baseline differ — fill_min_heads_tails,
SURVEY.md Appendix A.1 — on REFERENCE-RUN output instead of on a hand-derived list.

oracle/Makefile builds the reference's PFSP sources a second time with the single line
    lb1_data->min_heads[k] = MIN(lb1_data->min_heads[k], tmp[k - 1]);      (baselines/pfsp/lib/c_bound_simple.c:299)
rewritten by a committed `sed` into what the Chapel source says at the same place
    data.min_heads[k] = min(max(int(32)), tmp0);                            (lib/pfsp/Bound_simple.chpl:271)
(oracle/_ref/libref_pfsp_chapel.so, oracle/_ref/pfsp_c_chapel.out).  This script runs that build:
  * min_heads of ta001..ta030 as its fill_min_heads_tails produces them;
  * lb1_children_bounds of the ROOT (limit1 = -1, the only node that reads min_heads) for the same instances;
  * the explored-tree / solution counts its sequential search prints for ta014 and ta020 with lb1_d
    (ta020: 836 490 312 nodes against 859 257 178 with the C line — about four minutes of CPU time).
Run in the build container only (needs /root/reference):  make -C oracle ref && python tests/golden/make_golden_chapel.py [--no-ta020]
"""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pfsp_chapel_heads.json")


def run_counts(inst):
    exe = os.path.join(ROOT, "oracle", "_ref", "pfsp_c_chapel.out")
    txt = subprocess.run([exe, "--inst", str(inst), "--lb", "0", "--ub", "1"], capture_output=True, text=True,
                         cwd="/tmp").stdout
    g = lambda pat: int(re.search(pat, txt).group(1))  # noqa: E731
    return {"tree": g(r"explored tree: (\d+)"), "sol": g(r"explored solutions: (\d+)"), "best": g(r"makespan: (\d+)")}


def main():
    L = po.ref_pfsp(chapel_heads=True)
    out = {"_source": "oracle/_ref/libref_pfsp_chapel.so + pfsp_c_chapel.out: the reference's C sources with "
                      "c_bound_simple.c:299 rewritten by sed to the Chapel statement of Bound_simple.chpl:271",
           "min_heads": {}, "root_lb1_children_bounds": {}, "counts": {}}
    for inst in range(1, 31):
        d1, _ = po.ref_pfsp_data(inst, chapel_heads=True)
        jobs, machines = d1.contents.nb_jobs, d1.contents.nb_machines
        out["min_heads"][f"ta{inst:03d}"] = [int(d1.contents.min_heads[k]) for k in range(machines)]
        prmu = np.arange(jobs, dtype=np.int32)
        lb_begin = np.zeros(jobs, dtype=np.int32)
        L.lb1_children_bounds(d1, prmu.ctypes.data_as(C.c_void_p), -1, jobs, lb_begin.ctypes.data_as(C.c_void_p))
        out["root_lb1_children_bounds"][f"ta{inst:03d}"] = [int(x) for x in lb_begin]
    out["counts"]["ta014_lb1d"] = run_counts(14)
    if "--no-ta020" not in sys.argv:
        out["counts"]["ta020_lb1d"] = run_counts(20)
    elif os.path.exists(OUT):
        out["counts"]["ta020_lb1d"] = json.load(open(OUT))["counts"].get("ta020_lb1d")
    json.dump(out, open(OUT, "w"), indent=1)
    print("wrote", OUT, out["counts"])


if __name__ == "__main__":
    main()
