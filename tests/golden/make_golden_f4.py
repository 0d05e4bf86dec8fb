#!/usr/bin/env python
"""Golden vectors for SURVEY §8(f4) — larger instances (MAX_JOBS = 50: ta031..ta060) and the lb2 variants
(LB2_NABESHIMA, LB2_LAGEWEG) — from the REFERENCE's own C sources, built by oracle/Makefile with the one constant
each of them hard-codes rewritten by sed (`#define MAX_JOBS 20`, baselines/pfsp/lib/PFSP_node.h:10;
`enum lb2_variant lb2_type = LB2_FULL;`, baselines/pfsp/lib/c_bound_johnson.c:15,55):

    make -C oracle ref && python tests/golden/make_golden_f4.py      ->  tests/golden/pfsp_f4.npz

Per case: seeded random nodes (every depth), the tables the reference's fill_* functions produce, and the bounds its
lb1_bound / lb1_children_bounds / lb2_bound return for every live child slot (computed slot by slot as decompose_*
does, pfsp_c.c:106-191)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from oracle import pyoracle50 as po50  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pfsp_f4.npz")
INT_MAX = 2**31 - 1


def random_nodes(rng, jobs, count, dtype):
    nodes = np.zeros(count, dtype=dtype)
    for i in range(count):
        depth = 1 + i % (jobs - 1)  # 1 .. jobs-1
        nodes["depth"][i] = depth
        nodes["limit1"][i] = depth - 1
        nodes["prmu"][i, :jobs] = rng.permutation(jobs).astype(np.int32)
    return nodes


def ref_bounds(L, d1, d2, lb_kind, nodes, jobs, best):
    out = np.full((nodes.shape[0], jobs), -0x32323233, dtype=np.int32)
    for i in range(nodes.shape[0]):
        prmu = np.ascontiguousarray(nodes["prmu"][i]).astype(np.int32)
        limit1, depth = int(nodes["limit1"][i]), int(nodes["depth"][i])
        if lb_kind == 0:
            lbb = np.zeros(jobs, dtype=np.int32)
            L.lb1_children_bounds(d1, prmu.ctypes.data_as(C.c_void_p), limit1, jobs, lbb.ctypes.data_as(C.c_void_p))
            for k in range(limit1 + 1, jobs):
                out[i, k] = lbb[prmu[k]]
        else:
            for k in range(limit1 + 1, jobs):
                child = prmu.copy()
                child[depth], child[k] = child[k], child[depth]
                p = child.ctypes.data_as(C.c_void_p)
                out[i, k] = (L.lb1_bound(d1, p, limit1 + 1, jobs) if lb_kind == 1
                             else L.lb2_bound(d1, d2, p, limit1 + 1, jobs, int(best)))
    return out.reshape(-1)


def tables_of(d1, d2, out, tag):
    jobs, machines, pairs = d1.contents.nb_jobs, d1.contents.nb_machines, d2.contents.nb_machine_pairs
    asarr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).astype(np.int32).copy()  # noqa: E731
    out[f"{tag}_dims"] = np.array([jobs, machines, pairs], dtype=np.int32)
    out[f"{tag}_p_times"] = asarr(d1.contents.p_times, jobs * machines)
    out[f"{tag}_min_tails"] = asarr(d1.contents.min_tails, machines)
    out[f"{tag}_lags"] = asarr(d2.contents.lags, pairs * jobs)
    out[f"{tag}_johnson_qsort"] = asarr(d2.contents.johnson_schedules, pairs * jobs)
    out[f"{tag}_mp0"] = asarr(d2.contents.machine_pairs_1, pairs)
    out[f"{tag}_mp1"] = asarr(d2.contents.machine_pairs_2, pairs)
    return jobs


def main():
    rng = np.random.default_rng(0xF4)
    out = {}
    # ---- lb2 variants on 20-job instances (ta014: 10 machines, ta021: 20 machines)
    for variant in ("nabeshima", "lageweg"):
        L = po.ref_pfsp_named(variant)
        for inst in (14, 21):
            d1, d2 = po.ref_pfsp_data(inst, L=L)
            tag = f"{variant}_ta{inst:03d}"
            jobs = tables_of(d1, d2, out, tag)
            nodes = random_nodes(rng, jobs, 57, po.PFSP_NODE_DTYPE)
            best = int(po.lib().or_taillard_best_ub(inst))
            out[f"{tag}_parents"] = nodes.view(np.uint8).reshape(-1)
            out[f"{tag}_lb2_best"] = ref_bounds(L, d1, d2, 2, nodes, jobs, best)
            out[f"{tag}_lb2_inf"] = ref_bounds(L, d1, d2, 2, nodes, jobs, INT_MAX)
    # ---- MAX_JOBS = 50: ta031 (50x5), ta041 (50x10), ta051 (50x20)
    L = po.ref_pfsp_named("50")
    for inst in (31, 41, 51):
        d1, d2 = po.ref_pfsp_data(inst, L=L)
        tag = f"jobs50_ta{inst:03d}"
        jobs = tables_of(d1, d2, out, tag)
        assert jobs == 50
        nodes = random_nodes(rng, jobs, 98 if inst != 51 else 49, po50.PFSP_NODE_DTYPE)
        best = int(po.lib().or_taillard_best_ub(inst))
        out[f"{tag}_parents"] = nodes.view(np.uint8).reshape(-1)
        out[f"{tag}_lb1"] = ref_bounds(L, d1, d2, 1, nodes, jobs, best)
        out[f"{tag}_lb1_d"] = ref_bounds(L, d1, d2, 0, nodes, jobs, best)
        out[f"{tag}_lb2_best"] = ref_bounds(L, d1, d2, 2, nodes, jobs, best)
        out[f"{tag}_lb2_inf"] = ref_bounds(L, d1, d2, 2, nodes, jobs, INT_MAX)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if k.endswith("_dims")})


if __name__ == "__main__":
    main()
