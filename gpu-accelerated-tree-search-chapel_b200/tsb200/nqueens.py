"""N-Queens side of the offload interface (nqueens_gpu_chpl.chpl / nqueens_multigpu_chpl.chpl)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import SearchStats, check, lib

# lib/nqueens/NQueens_node.chpl:9-11
NQ_NODE_DTYPE = np.dtype([("depth", np.uint8), ("board", np.uint8, (20,))])
assert NQ_NODE_DTYPE.itemsize == 21


class NQueensEvaluator:
    """Owns what `on device var parents_d, labels_d` owns in the reference (nqueens_gpu_chpl.chpl:194-195)."""

    def __init__(self, N: int, g: int = 1, M: int = 50000, device: int = 0):
        self.N, self.g, self.M, self.device = N, g, M, device
        self._h = C.c_void_p()
        check(lib().tsb_nq_create(C.byref(self._h), device, N, g, M), "tsb_nq_create")

    def close(self):
        if self._h:
            lib().tsb_nq_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_xfer(self, mode: int):
        check(lib().tsb_nq_set_xfer(self._h, mode), "tsb_nq_set_xfer")

    @property
    def kernel_launches(self) -> int:
        return int(lib().tsb_nq_kernel_launches(self._h))

    def evaluate_gpu(self, parents: np.ndarray, size: int, labels: np.ndarray) -> None:
        """evaluate_gpu(parents_d, size, labels_d) of nqueens_gpu_chpl.chpl:97-123 including the copies of
        :203/:205; `size` = N * poolSize as in the reference call (:201-204)."""
        assert parents.dtype == NQ_NODE_DTYPE and parents.flags.c_contiguous
        assert labels.dtype == np.uint8 and labels.flags.c_contiguous
        if size % self.N:
            raise ValueError("size must be N * poolSize")
        count = size // self.N
        assert parents.shape[0] >= count and labels.size >= size
        check(lib().tsb_nq_evaluate(self._h, parents.ctypes.data, count, labels.ctypes.data), "tsb_nq_evaluate")

    def evaluate(self, parents: np.ndarray) -> np.ndarray:
        labels = np.empty(parents.shape[0] * self.N, dtype=np.uint8)
        self.evaluate_gpu(parents, parents.shape[0] * self.N, labels)
        return labels

    def evaluate_device(self, parents_ptr: int, count: int, labels_ptr: int, stream: int = 0) -> None:
        """device-resident form; pointers are raw device addresses (e.g. torch.Tensor.data_ptr())"""
        check(lib().tsb_nq_evaluate_device(self._h, parents_ptr, count, labels_ptr, stream), "tsb_nq_evaluate_device")


def nqueens_search(N: int = 14, g: int = 1, m: int = 25, M: int = 50000, D: int = 1) -> SearchStats:
    """the 3-step search of nqueens_gpu_chpl.chpl:152-248 (D = 1) / nqueens_multigpu_chpl.chpl:158-352
    (static split, D GPUs), run by the C++ emulation driver inside libtsb200.so"""
    st = SearchStats()
    check(lib().tsb_nq_search(N, g, m, M, D, C.byref(st)), "tsb_nq_search")
    return st
