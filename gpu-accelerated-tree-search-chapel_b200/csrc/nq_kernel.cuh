// nq_kernel.cuh — N-Queens batch conflict check for sm_100a.
//
// Replaces the reference's one-thread-per-(parent,k) foreach (nqueens_gpu_chpl.chpl:97-123;
// CUDA twin baselines/nqueens/nqueens_gpu_cuda.cu:137-164), which re-reads the 21-byte parent
// N times and runs an O(depth) loop per slot.  Here:
//   * the chunk is streamed through shared memory by the TMA engine (cp.async.bulk) in tiles of
//     512 parents (10 752 B in, 512*N B out), 4-stage mbarrier pipeline, persistent CTAs;
//   * one thread owns FOUR consecutive parents = 84 B = 21 aligned words in, N aligned words out,
//     so the 21-byte / N-byte records never need unaligned or byte-wide memory instructions and
//     the word strides (21, N odd for N = 17, 19) are bank-conflict free;
//   * per parent the placed queens are folded once into a 32-bit "attacked values" mask
//         U = OR_{i<depth} ( 1 << (board[i] + (depth-i)) | 1 << (board[i] - (depth-i)) )
//     (bits outside 0..N-1 fall off), so label[k] = !bit(U, board[k]) : O(depth + N) per parent
//     instead of O(depth * N).  Equivalent to the reference predicate
//         board[i] != board[k] - (depth-i)  &&  board[i] != board[k] + (depth-i)   for all i < depth
//     evaluated in int arithmetic (no uint8 wrap), nqueens_gpu_chpl.chpl:112-118.
// Slots k < depth are written 0 (the reference leaves them untouched).  `g` repeats an idempotent
// AND in the reference (:115-118); the result does not depend on it and the work is done once.
#pragma once
#include "tsb_ptx.cuh"

namespace tsb {

constexpr int NQ_THREADS = 128;
constexpr int NQ_QUAD = 4;                       // parents per thread
constexpr int NQ_TILE = NQ_THREADS * NQ_QUAD;    // 512 parents per tile
constexpr int NQ_REC = 21;                       // sizeof(tsb_nq_node)
constexpr int NQ_STAGES = 4;

template <int N>
using NqSmem = TileSmem<NQ_STAGES, NQ_TILE * NQ_REC, NQ_TILE * N>;

// byte `b` (compile-time) of a little-endian word array
template <int B>
__device__ __forceinline__ uint32_t byte_of(const uint32_t* w) {
  return (w[B >> 2] >> (8 * (B & 3))) & 0xFFu;
}

template <int N, int Q, int I>
struct NqRows {
  __device__ static __forceinline__ void run(const uint32_t* w, uint32_t depth, uint32_t& U) {
    if constexpr (I < N) {
      const uint32_t e = byte_of<21 * Q + 1 + I>(w);
      const uint32_t s = depth - I;  // > 0 for placed rows
      const uint32_t bits = shl_clamp(1u, e + s) | shl_clamp(1u, e - s);  // e - s < 0 wraps to >= 32 -> 0
      if (I < depth) U |= bits;
      NqRows<N, Q, I + 1>::run(w, depth, U);
    }
  }
};

template <int N, int Q, int K>
struct NqSlots {
  __device__ static __forceinline__ void run(const uint32_t* w, uint32_t depth, uint32_t safe, uint32_t* o) {
    if constexpr (K < N) {
      const uint32_t e = byte_of<21 * Q + 1 + K>(w);
      uint32_t bit = (safe >> e) & 1u;  // e <= 19
      if (K < depth) bit = 0;
      constexpr int OB = Q * N + K;  // output byte index inside this thread's 4N bytes
      o[OB >> 2] |= bit << (8 * (OB & 3));
      NqSlots<N, Q, K + 1>::run(w, depth, safe, o);
    }
  }
};

template <int N, int Q>
__device__ __forceinline__ void nq_one_parent(const uint32_t* w, uint32_t* o) {
  const uint32_t depth = byte_of<21 * Q>(w);
  uint32_t U = 0;
  NqRows<N, Q, 0>::run(w, depth, U);
  NqSlots<N, Q, 0>::run(w, depth, ~U, o);
}

template <int N>
__device__ __forceinline__ void nq_compute_tile(const uint8_t* in_tile, uint8_t* out_tile, int /*records*/) {
  const uint32_t* in_w = reinterpret_cast<const uint32_t*>(in_tile) + 21 * threadIdx.x;
  uint32_t* out_w = reinterpret_cast<uint32_t*>(out_tile) + N * threadIdx.x;
  uint32_t w[21];
#pragma unroll
  for (int i = 0; i < 21; i++) w[i] = in_w[i];
  uint32_t o[N];
#pragma unroll
  for (int i = 0; i < N; i++) o[i] = 0;
  nq_one_parent<N, 0>(w, o);
  nq_one_parent<N, 1>(w, o);
  nq_one_parent<N, 2>(w, o);
  nq_one_parent<N, 3>(w, o);
#pragma unroll
  for (int i = 0; i < N; i++) out_w[i] = o[i];
}

template <int N>
__global__ void __launch_bounds__(NQ_THREADS) nq_evaluate_kernel(const uint8_t* __restrict__ parents,
                                                                uint8_t* __restrict__ labels, long long count) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  NqSmem<N>& sm = *reinterpret_cast<NqSmem<N>*>(smem_raw);
  run_tile_pipeline<NQ_STAGES, NQ_TILE, NQ_REC, N>(
      sm, parents, labels, count,
      [](const uint8_t* in_tile, uint8_t* out_tile, int n) { nq_compute_tile<N>(in_tile, out_tile, n); });
}

}  // namespace tsb
