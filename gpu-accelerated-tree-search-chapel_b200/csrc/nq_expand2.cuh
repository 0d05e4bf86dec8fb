// nq_expand2.cuh — side words of the device pool (N-Queens): each node's attacked diagonals, kept next to the node.
// A/B EXPERIMENT, off by default (TSB200_AUX=1): measured no faster than the plain kernels, see below.
//
// nq_expand_count re-derives a parent's attacked values from its whole board (the O(depth) loop of isSafe,
// nqueens_gpu_chpl.chpl:79-94: ~200 of the count kernel's 310 warp instructions per parent) although the parent was
// BUILT one round earlier from a node whose diagonals were known.  The device pool therefore keeps one 8-byte side
// word per arena position — ld | rd << 20, the values attacked on the node's next row along the rising / falling
// diagonals (layout of nq_aux_pack, nq_rounds.cuh; the upper fields are not used here).  The AUX variants of the
// count / build kernels (nq_expand.cuh) read a parent's word instead of walking its board and write each child's
// word, derived in O(1) (nq_child_ldrd), next to the child.
//
// Measured alternative (N = 17, 4 Mi parents, B200): ALSO evaluating the child's mask when it is built — so that the
// count kernel only pop-counts side words (9.7 us) and the build kernel regenerates its items from them — moved 100
// instructions per child into the build kernel, which is latency- rather than issue-bound: 98 us against 51.5 us,
// 108 us per round against 96 us for the pair without side words (profiles/nq_build2_fullaux_r2_ncu.txt).  With the
// mask kept in the count kernel and only the O(1) diagonal update in the build kernel (this file + the AUX variants):
// count 36.8 us (22.7 M warp instructions instead of 36 M) + build 59.9 us = 96.7 us against 44.6 + 51.5 = 96.1 us
// (profiles/nq_{count,build}_sidewords_r2_ncu.txt): a round of 4 Mi parents moves 347 MB instead of ~280 MB, both
// kernels sit at 37-39 % of the DRAM peak with long-scoreboard as their first stall — the instructions saved are paid
// back in bytes.  The default therefore stays the pair WITHOUT side words.
//
// Nodes that did not come out of the build kernel (host pushes, stolen nodes, the persistent kernel's export) get
// their word from nq_aux_fill_kernel (the reference predicate, row by row).
#pragma once
#include "nq_expand.cuh"
#include "nq_rounds.cuh"

namespace tsb {

// side words of arena positions [lo, hi) from the boards
template <int N>
__global__ void __launch_bounds__(256) nq_aux_fill_kernel(const uint8_t* __restrict__ arena,
                                                         unsigned long long* __restrict__ aux, long long lo,
                                                         long long hi) {
  for (long long pos = lo + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pos < hi;
       pos += static_cast<long long>(gridDim.x) * blockDim.x)
    aux[pos] = nq_aux_of_node<N>(arena + pos * NQ_REC);
}

}  // namespace tsb
