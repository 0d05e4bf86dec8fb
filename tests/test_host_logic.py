"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol
include/tsb200.h declares, argument validation / error codes work without a GPU, and the host table
precompute (tsb_pfsp_tables_build) agrees with the oracle.  No compute entry point is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tsb200
from oracle import pyoracle as po
from tsb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tsb200.h")).read()
    declared = set(re.findall(r"\b(tsb_[a-z0-9_]+)\s*\(", header))
    declared -= {"tsb_nq", "tsb_pfsp"}
    assert declared, "no declarations parsed"
    L = tsb200.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libtsb200.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_struct_sizes_match_wire_format():
    assert tsb200.NQ_NODE_DTYPE.itemsize == 21 and tsb200.PFSP_NODE_DTYPE.itemsize == 88
    assert C.sizeof(_lib.PfspTables) == 4 * (3 + 400 + 20 + 20 + 3800 + 3800 + 3 * 190)


def test_error_codes_and_messages():
    L = tsb200.lib()
    for code in range(0, -7, -1):
        assert L.tsb_strerror(code)
    h = C.c_void_p()
    assert L.tsb_nq_create(C.byref(h), 0, 0, 1, 10) == _lib.EINVAL       # N out of range
    assert L.tsb_nq_create(C.byref(h), 0, 21, 1, 10) == _lib.EINVAL
    assert L.tsb_nq_create(C.byref(h), 0, 8, 0, 10) == _lib.EINVAL       # g < 1
    assert L.tsb_nq_create(None, 0, 8, 1, 10) == _lib.EINVAL
    assert L.tsb_nq_evaluate(None, None, 1, None) == _lib.EINVAL
    assert L.tsb_pfsp_evaluate(None, 1, None, 1, 0, None) == _lib.EINVAL
    st = _lib.SearchStats()
    assert L.tsb_nq_search(0, 1, 25, 50000, 1, C.byref(st)) == _lib.EINVAL
    assert L.tsb_pfsp_search(14, 3, 1, 25, 50000, 1, C.byref(st)) == _lib.EINVAL
    assert L.tsb_pfsp_search(31, 1, 1, 25, 50000, 1, C.byref(st)) == _lib.EUNSUPPORTED  # 50 jobs > MAX_JOBS
    t = _lib.PfspTables()
    assert L.tsb_pfsp_tables_build(C.byref(t), 0) == _lib.EINVAL
    assert L.tsb_pfsp_tables_build(C.byref(t), 31) == _lib.EUNSUPPORTED


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tsb200.TsbError):
        tsb200.NQueensEvaluator(8)
    with pytest.raises(tsb200.TsbError):
        tsb200.PfspEvaluator(14)
    with pytest.raises(tsb200.TsbError):
        tsb200.nqueens_search(8)


@pytest.mark.parametrize("inst", list(range(1, 31)))
def test_host_tables_equal_oracle(inst):
    t = tsb200.taillard_tables(inst)
    o = po.tables(inst, heads_mode=0)
    assert (t.jobs, t.machines, t.pairs) == (o.jobs, o.machines, o.pairs)
    assert tsb200.lib().tsb_taillard_best_ub(inst) == po.lib().or_taillard_best_ub(inst)
    for name in ("p_times", "min_heads", "min_tails", "lags", "mp0", "mp1", "mp_order", "johnson"):
        np.testing.assert_array_equal(np.ctypeslib.as_array(getattr(t, name)), o.arr(name), err_msg=name)


def test_taillard_shapes():
    L = tsb200.lib()
    for inst in range(1, 121):
        assert L.tsb_taillard_nb_jobs(inst) == po.lib().or_taillard_nb_jobs(inst)
        assert L.tsb_taillard_nb_machines(inst) == po.lib().or_taillard_nb_machines(inst)


def test_chapel_binding_declares_only_exported_symbols_with_matching_arity():
    """chapel/TSB200.chpl cannot be compiled here (no chpl): at least keep its extern procs in step with the header —
    every extern proc must exist in include/tsb200.h with the same number of parameters"""
    header = open(os.path.join(ROOT, "include", "tsb200.h")).read()
    chpl = open(os.path.join(ROOT, "gpu-accelerated-tree-search-chapel_b200", "chapel", "TSB200.chpl")).read()
    decls = {}
    for m in re.finditer(r"\b(tsb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    externs = re.findall(r"extern proc (tsb_[a-z0-9_]+)\s*\((.*?)\)\s*(?::|;)", chpl, flags=re.S)
    assert len(externs) >= 15
    for name, args in externs:
        assert name in decls, f"{name} is not declared in include/tsb200.h"
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == decls[name], f"{name}: {n} parameters in TSB200.chpl, {decls[name]} in the header"


@pytest.mark.parametrize("variant", ["full", "nabeshima", "lageweg", "learn"])
@pytest.mark.parametrize("inst", [3, 14, 21])
def test_host_tables_lb2_variants_equal_oracle(inst, variant):
    t = tsb200.taillard_tables(inst, variant)
    o = po.tables(inst, 0, po.LB2_VARIANTS[variant])
    assert (t.jobs, t.machines, t.pairs) == (o.jobs, o.machines, o.pairs)
    for name in ("lags", "mp0", "mp1", "mp_order", "johnson"):
        np.testing.assert_array_equal(np.ctypeslib.as_array(getattr(t, name)), o.arr(name), err_msg=name)


@pytest.mark.parametrize("inst", [31, 41, 51, 60])
def test_host_tables50_equal_oracle50(inst):
    """tables of the 50-job instances (a MAX_JOBS = 50 build) against the oracle built that way"""
    from oracle import pyoracle50 as po50
    t = tsb200.taillard_tables50(inst)
    o = po50.tables(inst, heads_mode=0)
    assert (t.jobs, t.machines, t.pairs) == (o.jobs, o.machines, o.pairs) and t.jobs == 50
    for name in ("p_times", "min_heads", "min_tails", "lags", "mp0", "mp1", "mp_order", "johnson"):
        np.testing.assert_array_equal(np.ctypeslib.as_array(getattr(t, name)), o.arr(name), err_msg=name)
    with pytest.raises(tsb200.TsbError):
        tsb200.taillard_tables(inst)  # does not fit a MAX_JOBS = 20 build
