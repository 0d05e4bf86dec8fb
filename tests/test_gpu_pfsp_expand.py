"""GPU tests of the PFSP fused evaluate + generate_children path and of the device-resident PFSP pool (SURVEY
§8f rows 1, 3): children arrays byte-identical to the oracle's evaluate + sequential generate_children, the
incumbent updated exactly as the reference does (including rounds in which a leaf improves it), pools
byte-identical after the same rounds, and whole searches with the reference's counts."""
import json
import os

import numpy as np
import pytest

import tsb200
from oracle import pyoracle as po
from test_gpu_parity import rand_pfsp

pytestmark = pytest.mark.gpu
LBS = ("lb1", "lb1_d", "lb2")


def _check_expand(ev, t, parents, lb, best):
    got, gsol, gbest = ev.expand(parents, lb, best)
    want, wsol, wbest = po.pfsp_expand(t, tsb200.LB_NAMES[lb], parents.view(po.PFSP_NODE_DTYPE), best)
    assert (gsol, gbest, got.shape[0]) == (wsol, wbest, want.shape[0])
    assert got.tobytes() == want.tobytes()
    return got.shape[0], gbest


@pytest.mark.parametrize("inst,opt", [(14, 1377), (20, 1591), (1, 1278), (21, 2297)])
@pytest.mark.parametrize("lb", LBS)
def test_expand_matches_oracle_children(inst, opt, lb):
    """random chunks of every size class; best = optimum (never improved), optimum + 60 (prunes less)"""
    rng = np.random.default_rng(900 + inst)
    t = po.tables(inst, heads_mode=0)
    with tsb200.PfspEvaluator(inst, M=6000) as ev:
        for count, lo in ((1, 1), (2, 1), (127, 1), (128, 1), (129, 3), (1000 + 17, 1), (6000, 8)):
            parents = rand_pfsp(rng, 20, count, depth_lo=lo)
            for best in (opt, opt + 60):
                _check_expand(ev, t, parents, lb, best)
        assert ev.slow_rounds == 0 or lb != "lb2"  # random leaves are far above the optimum


@pytest.mark.parametrize("lb", LBS)
def test_expand_when_a_leaf_improves_best(lb):
    """best = INT64 max / a loose value: leaf children (depth 19 parents) lower it in the middle of the chunk,
    and the rest of the chunk is pruned against the lowered value — the reference's sequential rule"""
    inst = 14
    rng = np.random.default_rng(77)
    t = po.tables(inst, heads_mode=0)
    parents = rand_pfsp(rng, 20, 3000, depth_lo=15)  # many depth-19 parents
    with tsb200.PfspEvaluator(inst, M=3000) as ev:
        for best in (2**63 - 1, 2**31 - 1, 1900, 1700):
            n, b = _check_expand(ev, t, parents, lb, best)
            assert b < best
        assert ev.slow_rounds >= 3


@pytest.mark.parametrize("lb,which", [("lb1", 10), ("lb1_d", 30), ("lb2", 5)])
def test_expand_on_captured_real_chunks(lb, which):
    parents, best = po.pfsp_capture_chunk(14, tsb200.LB_NAMES[lb], which)
    t = po.tables(14, heads_mode=0)
    with tsb200.PfspEvaluator(14, M=50000) as ev:
        _check_expand(ev, t, parents.view(tsb200.PFSP_NODE_DTYPE), lb, best)


def test_root_and_shallow_parents():
    """limit1 = -1 (root: lb1_d seeds the front with min_heads) and depth 0..2 parents: up to 20 children each,
    several passes of the staging image"""
    t = po.tables(14, heads_mode=0)
    rng = np.random.default_rng(5)
    parents = rand_pfsp(rng, 20, 700, depth_lo=0)
    parents["depth"][:300] = 0
    parents["limit1"][:300] = -1
    with tsb200.PfspEvaluator(14, M=700) as ev:
        for lb in ("lb1", "lb1_d"):
            n, _ = _check_expand(ev, t, parents, lb, 2**31 - 1)
            assert n > 300 * 19


@pytest.mark.parametrize("lb", LBS)
def test_device_pool_is_byte_identical_to_the_reference_pool(lb):
    inst, m, M, best = 14, 25, 300, 1377
    t = po.tables(inst, heads_mode=0)
    rng = np.random.default_rng(11)
    start = rand_pfsp(rng, 20, 40, depth_lo=2)
    start["depth"][:] = np.minimum(start["depth"], 6)
    start["limit1"][:] = start["depth"] - 1
    host_pool = start.copy()
    with tsb200.PfspEvaluator(inst, M=M) as ev:
        ev.pool_push(start)
        for _ in range(60):
            n_par, n_child, n_sol, best2 = ev.pool_step(lb, m, M, best)
            if host_pool.shape[0] < m:
                assert n_par == 0
                break
            n = min(host_pool.shape[0], M)
            chunk = np.ascontiguousarray(host_pool[host_pool.shape[0] - n:])
            kids, sol, wbest = po.pfsp_expand(t, tsb200.LB_NAMES[lb], chunk.view(po.PFSP_NODE_DTYPE), best)
            host_pool = np.concatenate([host_pool[: host_pool.shape[0] - n], kids.view(tsb200.PFSP_NODE_DTYPE)])
            assert (n_par, n_child, n_sol, best2) == (n, kids.shape[0], sol, wbest)
            best = best2
            assert ev.pool_size == host_pool.shape[0]
        rest = ev.pool_drain()
        assert rest.tobytes() == np.ascontiguousarray(host_pool).tobytes() and ev.pool_size == 0


@pytest.mark.parametrize("inst,lb,ub,m,M,D", [(14, "lb1", 1, 25, 50000, 1), (14, "lb1_d", 1, 25, 50000, 1),
                                              (14, "lb2", 1, 25, 50000, 1), (14, "lb1", 1, 25, 3000, 3),
                                              (14, "lb1", 1, 5, 1 << 20, 1), (14, "lb2", 1, 25, 700, 2),
                                              (14, "lb1_d", 0, 25, 50000, 1)])
def test_device_resident_search_counts(golden_dir, inst, lb, ub, m, M, D, monkeypatch):
    """whole searches: identical explored tree / solutions / optimum and the same chunk sequence as the reference
    driver (ub = 0: the incumbent is found on the way, several slow rounds; single task, so still deterministic)"""
    monkeypatch.setenv("TSB200_NO_STEAL", "1")  # the static split alone: the reference driver's chunk sequence
    st = tsb200.pfsp_search_device(inst, lb, ub, m, M, D)
    ref = po.pfsp_search_offload(inst, tsb200.LB_NAMES[lb], ub, m, M, D)
    assert (st.explored_tree, st.explored_sol, st.best) == (ref.tree, ref.sol, ref.best)
    assert (st.offloads, st.offloaded_parents) == (ref.offloads, ref.offloaded_parents)
    if ub == 1:
        counts = json.load(open(os.path.join(golden_dir, "counts.json")))["pfsp"]
        key = f"ta{inst:03d}_lb{tsb200.LB_NAMES[lb]}_ub1"
        assert (st.explored_tree, st.explored_sol, st.best) == (counts[key]["tree"], counts[key]["sol"], counts[key]["best"])


@pytest.mark.parametrize("inst,lb,m,M,D", [(14, "lb1", 25, 3000, 3), (14, "lb1_d", 25, 50000, 4), (14, "lb2", 25, 700, 2),
                                           (14, "lb1", 25, 500, 8)])
def test_device_resident_search_with_work_stealing(golden_dir, inst, lb, m, M, D):
    """D tasks with device pools that steal from each other (--ub 1: the counts do not depend on who explores what)"""
    st = tsb200.pfsp_search_device(inst, lb, 1, m, M, D)
    counts = json.load(open(os.path.join(golden_dir, "counts.json")))["pfsp"]
    key = f"ta{inst:03d}_lb{tsb200.LB_NAMES[lb]}_ub1"
    assert (st.explored_tree, st.explored_sol, st.best) == (counts[key]["tree"], counts[key]["sol"], counts[key]["best"])


def test_search_on_a_precreated_handle(golden_dir):
    counts = json.load(open(os.path.join(golden_dir, "counts.json")))["pfsp"]["ta014_lb1_ub1"]
    with tsb200.PfspEvaluator(14, M=50000) as ev:
        for _ in range(3):  # the handle (tables, arena, side arrays) is reused
            st = ev.search(14, "lb1", 1, 25, 50000)
            assert (st.explored_tree, st.explored_sol, st.best) == (counts["tree"], counts["sol"], counts["best"])
    nq = json.load(open(os.path.join(golden_dir, "counts.json")))["nqueens"]["13"]
    with tsb200.NQueensEvaluator(13, M=50000) as ev:
        for _ in range(3):
            st = ev.search(25, 50000)
            assert (st.explored_tree, st.explored_sol) == (nq["tree"], nq["sol"])


def test_ta020_lb1d_search_has_the_chapel_program_count(golden_dir):
    """ta020 with lb1_d is where the Chapel program (min_heads of Bound_simple.chpl) and the C baseline explore
    different trees (836 490 312 against 859 257 178 nodes): the device-pool search matches the count of the reference's C
    code rebuilt with the Chapel statement (tests/golden/make_golden_chapel.py)"""
    gold = json.load(open(os.path.join(golden_dir, "pfsp_chapel_heads.json")))["counts"].get("ta020_lb1d")
    if not gold:
        pytest.skip("golden count not generated (make_golden_chapel.py --no-ta020)")
    with tsb200.PfspEvaluator(20, M=1 << 20) as ev:
        st = ev.search(20, "lb1_d", 1, 25, 1 << 20)
    assert (st.explored_tree, st.explored_sol, st.best) == (gold["tree"], gold["sol"], gold["best"])


def test_pool_compaction_and_growth(monkeypatch):
    monkeypatch.setenv("TSB200_POOL_CAP", "4000")
    st = tsb200.pfsp_search_device(14, "lb1", 1, 25, 400, 1)
    ref = po.pfsp_search_offload(14, tsb200.LB_NAMES["lb1"], 1, 25, 400, 1)
    assert (st.explored_tree, st.explored_sol, st.best) == (ref.tree, ref.sol, ref.best)
    assert (st.offloads, st.offloaded_parents) == (ref.offloads, ref.offloaded_parents)
