"""PFSP side of the offload interface (pfsp_gpu_chpl.chpl / pfsp_multigpu_chpl.chpl)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import LB1, LB1_D, LB2, PfspTables, SearchStats, check, lib

# lib/pfsp/PFSP_node.chpl:9-12
PFSP_NODE_DTYPE = np.dtype([("depth", np.int32), ("limit1", np.int32), ("prmu", np.int32, (20,))])
assert PFSP_NODE_DTYPE.itemsize == 88
# the Chapel CLI spells the bounds as strings (pfsp_gpu_chpl.chpl:15), the C ABI as the C baseline's ints
LB_NAMES = {"lb1_d": LB1_D, "lb1": LB1, "lb2": LB2}


def taillard_tables(inst: int) -> PfspTables:
    """lbound1 / lbound2 as built at pfsp_gpu_chpl.chpl:325-332 (Chapel semantics, incl. its min_heads)"""
    t = PfspTables()
    check(lib().tsb_pfsp_tables_build(C.byref(t), inst), "tsb_pfsp_tables_build")
    return t


class PfspEvaluator:
    """Owns parents_d / bounds_d / lbound1_d / lbound2_d of pfsp_gpu_chpl.chpl:359-371."""

    def __init__(self, inst: int | None = None, tables: PfspTables | None = None, M: int = 50000, device: int = 0):
        self.tables = tables if tables is not None else taillard_tables(inst)
        self.jobs, self.machines, self.M = self.tables.jobs, self.tables.machines, M
        self._h = C.c_void_p()
        check(lib().tsb_pfsp_create_from_tables(C.byref(self._h), device, M, C.byref(self.tables)), "tsb_pfsp_create")

    def close(self):
        if self._h:
            lib().tsb_pfsp_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_xfer(self, mode: int):
        check(lib().tsb_pfsp_set_xfer(self._h, mode), "tsb_pfsp_set_xfer")

    @property
    def kernel_launches(self) -> int:
        return int(lib().tsb_pfsp_kernel_launches(self._h))

    def evaluate_gpu(self, parents: np.ndarray, size: int, best: int, lb, bounds: np.ndarray) -> None:
        """evaluate_gpu(parents_d, size, best, lbound1_d, lbound2_d, bounds_d) of pfsp_gpu_chpl.chpl:257-270 with
        the copies of :384/:386; `size` = jobs * poolSize; lb is "lb1" | "lb1_d" | "lb2" or the int code"""
        assert parents.dtype == PFSP_NODE_DTYPE and parents.flags.c_contiguous
        assert bounds.dtype == np.int32 and bounds.flags.c_contiguous
        kind = LB_NAMES[lb] if isinstance(lb, str) else int(lb)
        if size % self.jobs:
            raise ValueError("size must be jobs * poolSize")
        count = size // self.jobs
        assert parents.shape[0] >= count and bounds.size >= size
        check(lib().tsb_pfsp_evaluate(self._h, kind, parents.ctypes.data, count, int(best), bounds.ctypes.data),
              "tsb_pfsp_evaluate")

    def evaluate(self, parents: np.ndarray, lb, best: int) -> np.ndarray:
        bounds = np.empty(parents.shape[0] * self.jobs, dtype=np.int32)
        self.evaluate_gpu(parents, parents.shape[0] * self.jobs, best, lb, bounds)
        return bounds

    def evaluate_device(self, lb, parents_ptr: int, count: int, best: int, bounds_ptr: int, stream: int = 0) -> None:
        kind = LB_NAMES[lb] if isinstance(lb, str) else int(lb)
        check(lib().tsb_pfsp_evaluate_device(self._h, kind, parents_ptr, count, int(best), bounds_ptr, stream),
              "tsb_pfsp_evaluate_device")


def pfsp_search(inst: int = 14, lb="lb1", ub: int = 1, m: int = 25, M: int = 50000, D: int = 1) -> SearchStats:
    """pfsp_gpu_chpl.chpl:306-431 (D = 1) / pfsp_multigpu_chpl.chpl (static split), C++ emulation driver"""
    kind = LB_NAMES[lb] if isinstance(lb, str) else int(lb)
    st = SearchStats()
    check(lib().tsb_pfsp_search(inst, kind, ub, m, M, D, C.byref(st)), "tsb_pfsp_search")
    return st
