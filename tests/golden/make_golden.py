#!/usr/bin/env python
"""Generate tests/golden/*.npz / counts.json from the REFERENCE's own C sources.

Run in the build container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

It calls the reference functions compiled into oracle/_ref/libref_{nqueens,pfsp}.so
(isSafe, lb1_bound, lb1_children_bounds, lb2_bound, fill_*; baselines/nqueens/nqueens_c.c,
baselines/pfsp/lib/c_bound_*.c) on seeded random nodes, and runs the unmodified reference
binaries oracle/_ref/{nqueens_c,pfsp_c}.out for the explored-tree counts.  The fixtures are
committed; the GPU box never needs /root/reference.

Where the C baseline and the Chapel program differ (min_heads) the
fixture records the C value and says so (`heads_mode = 1`); the Chapel-semantics values
are pinned separately by the SURVEY Appendix B known answers in counts.json.
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

OUT = os.path.dirname(os.path.abspath(__file__))
INT_MAX = 2**31 - 1


def random_nq_nodes(rng, N, count):
    """every depth 0..N appears; board = random permutation of 0..N-1 (unconditioned: any
    input is a valid argument of the predicate), bytes N..19 zero like the reference root"""
    nodes = np.zeros(count, dtype=po.NQ_NODE_DTYPE)
    for i in range(count):
        nodes["depth"][i] = i % (N + 1)
        nodes["board"][i, :N] = rng.permutation(N).astype(np.uint8)
    return nodes


def ref_nq_labels(nodes, N, g=1):
    L = po.ref_nqueens()
    labels = np.full((nodes.shape[0], N), 0xCD, dtype=np.uint8)
    for i in range(nodes.shape[0]):
        board = np.ascontiguousarray(nodes["board"][i])
        depth = int(nodes["depth"][i])
        for k in range(depth, N):
            # decompose() in nqueens_c.c:97-104 calls isSafe(G, board, depth, board[j])
            labels[i, k] = L.isSafe(g, board.ctypes.data_as(C.c_void_p), depth, int(board[k]))
    return labels.reshape(-1)


def random_pfsp_nodes(rng, jobs, count, with_root):
    nodes = np.zeros(count, dtype=po.PFSP_NODE_DTYPE)
    for i in range(count):
        depth = i % jobs  # 0..jobs-1
        if depth == 0 and not with_root:
            depth = 1 + (i // jobs) % (jobs - 1)
        nodes["depth"][i] = depth
        nodes["limit1"][i] = depth - 1
        nodes["prmu"][i, :jobs] = rng.permutation(jobs).astype(np.int32)
    return nodes


def ref_pfsp_bounds(d1, d2, lb_kind, nodes, jobs, best):
    L = po.ref_pfsp()
    out = np.full((nodes.shape[0], jobs), -0x32323233, dtype=np.int32)
    for i in range(nodes.shape[0]):
        prmu = np.ascontiguousarray(nodes["prmu"][i]).astype(np.int32)
        limit1, depth = int(nodes["limit1"][i]), int(nodes["depth"][i])
        if lb_kind == 0:  # pfsp_c.c:134-162 decompose_lb1_d
            lbb = np.zeros(jobs, dtype=np.int32)
            L.lb1_children_bounds(d1, prmu.ctypes.data_as(C.c_void_p), limit1, jobs, lbb.ctypes.data_as(C.c_void_p))
            for k in range(limit1 + 1, jobs):
                out[i, k] = lbb[prmu[k]]
        else:  # pfsp_c.c:106-132 decompose_lb1 / :164-191 decompose_lb2
            for k in range(limit1 + 1, jobs):
                child = prmu.copy()
                child[depth], child[k] = child[k], child[depth]
                p = child.ctypes.data_as(C.c_void_p)
                out[i, k] = (L.lb1_bound(d1, p, limit1 + 1, jobs) if lb_kind == 1
                             else L.lb2_bound(d1, d2, p, limit1 + 1, jobs, int(best)))
    return out.reshape(-1)


def run_counts(cmd):
    txt = subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=600).stdout
    tree = int(re.search(r"Size of the explored tree: (\d+)", txt).group(1))
    sol = int(re.search(r"Number of explored solutions: (\d+)", txt).group(1))
    m = re.search(r"Optimal makespan: (\d+)", txt)
    return {"tree": tree, "sol": sol, **({"best": int(m.group(1))} if m else {})}


def main():
    assert po.ref_available(), "run `make -C oracle ref` first"
    rng = np.random.default_rng(0x5EED)

    # ---- N-Queens label vectors
    nq = {}
    for N in (5, 8, 14, 17, 19, 20):
        nodes = random_nq_nodes(rng, N, 12 * (N + 1))
        nq[f"parents_N{N}"] = nodes.view(np.uint8).reshape(-1)
        nq[f"labels_N{N}"] = ref_nq_labels(nodes, N)
    np.savez_compressed(os.path.join(OUT, "nqueens_labels.npz"), **nq)

    # ---- PFSP tables + bound vectors (ta001 20x5, ta014/ta020 20x10, ta021 20x20)
    pf = {}
    for inst in (1, 14, 20, 21):
        d1, d2 = po.ref_pfsp_data(inst)
        jobs, machines, pairs = d1.contents.nb_jobs, d1.contents.nb_machines, d2.contents.nb_machine_pairs
        tag = f"ta{inst:03d}"
        asarr = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).astype(np.int32).copy()  # noqa: E731
        pf[f"{tag}_dims"] = np.array([jobs, machines, pairs], dtype=np.int32)
        pf[f"{tag}_p_times"] = asarr(d1.contents.p_times, jobs * machines)
        pf[f"{tag}_min_heads_C"] = asarr(d1.contents.min_heads, machines)
        pf[f"{tag}_min_tails"] = asarr(d1.contents.min_tails, machines)
        pf[f"{tag}_lags"] = asarr(d2.contents.lags, pairs * jobs)
        pf[f"{tag}_johnson_qsort"] = asarr(d2.contents.johnson_schedules, pairs * jobs)
        pf[f"{tag}_mp0"] = asarr(d2.contents.machine_pairs_1, pairs)
        pf[f"{tag}_mp1"] = asarr(d2.contents.machine_pairs_2, pairs)
        ident = np.arange(jobs, dtype=np.int32)
        L = po.ref_pfsp()
        pf[f"{tag}_kat"] = np.array([
            L.eval_solution(d1, ident.ctypes.data_as(C.c_void_p)),
            L.lb1_bound(d1, ident.ctypes.data_as(C.c_void_p), 0, jobs),
            L.lb2_bound(d1, d2, ident.ctypes.data_as(C.c_void_p), 0, jobs, INT_MAX)], dtype=np.int32)
        best = int(po.lib().or_taillard_best_ub(inst))
        n = 60 if inst in (14, 20) else 40
        nodes = random_pfsp_nodes(rng, jobs, n, with_root=False)
        pf[f"{tag}_parents"] = nodes.view(np.uint8).reshape(-1)
        pf[f"{tag}_lb1"] = ref_pfsp_bounds(d1, d2, 1, nodes, jobs, best)
        pf[f"{tag}_lb1_d"] = ref_pfsp_bounds(d1, d2, 0, nodes, jobs, best)
        pf[f"{tag}_lb2_best"] = ref_pfsp_bounds(d1, d2, 2, nodes, jobs, best)
        pf[f"{tag}_lb2_inf"] = ref_pfsp_bounds(d1, d2, 2, nodes, jobs, INT_MAX)
        # the root (limit1 = -1) is only ever evaluated by lb1_d; C min_heads semantics here
        root = np.zeros(1, dtype=po.PFSP_NODE_DTYPE)
        root["limit1"][0] = -1
        root["prmu"][0, :jobs] = ident
        pf[f"{tag}_root_lb1_d_C"] = ref_pfsp_bounds(d1, d2, 0, root, jobs, best)
    np.savez_compressed(os.path.join(OUT, "pfsp_bounds.npz"), **pf)

    # ---- counts printed by the unmodified reference binaries
    counts = {"_source": "oracle/_ref/{nqueens_c,pfsp_c}.out = the reference's baselines compiled unmodified; "
                         "large entries from SURVEY.md Appendix B (same binaries, longer runs)",
              "nqueens": {}, "pfsp": {}}
    for N in range(4, 15):
        counts["nqueens"][str(N)] = run_counts([os.path.join(ROOT, "oracle/_ref/nqueens_c.out"), "-N", str(N)])
    counts["nqueens"]["15"] = {"tree": 171129071, "sol": 2279184}
    counts["nqueens"]["16"] = {"tree": 1141190302, "sol": 14772512}
    counts["nqueens"]["17"] = {"tree": 8017021931, "sol": 95815104}
    counts["nqueens_classical_solutions"] = {"4": 2, "5": 10, "6": 4, "7": 40, "8": 92, "9": 352, "10": 724,
                                             "11": 2680, "12": 14200, "13": 73712, "14": 365596, "15": 2279184,
                                             "16": 14772512, "17": 95815104, "18": 666090624, "19": 4968057848}
    # (only ta014 finishes in seconds with the sequential reference binary)
    for inst, lbs in ((14, (0, 1, 2)),):
        for lb in lbs:
            counts["pfsp"][f"ta{inst:03d}_lb{lb}_ub1"] = run_counts(
                [os.path.join(ROOT, "oracle/_ref/pfsp_c.out"), "--inst", str(inst), "--lb", str(lb), "--ub", "1"])
    counts["pfsp"]["ta020_lb2_ub1"] = {"tree": 4870386, "sol": 0, "best": 1591}
    counts["pfsp"]["ta020_lb1_ub1"] = {"tree": 859257178, "sol": 3764, "best": 1591}
    counts["pfsp"]["ta020_lb0_ub1_Csemantics"] = {"tree": 859257178, "sol": 3764, "best": 1591}
    counts["pfsp"]["ta020_lb0_ub1_Chapelsemantics"] = {"tree": 836490312, "sol": 3764, "best": 1591}
    counts["chapel_min_heads"] = {"ta014": [0, 32, 53, 79, 108, 159, 216, 290, 312, 358],
                                  "ta020": [0, 97, 153, 245, 310, 338, 364, 427, 457, 504]}
    with open(os.path.join(OUT, "counts.json"), "w") as f:
        json.dump(counts, f, indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
