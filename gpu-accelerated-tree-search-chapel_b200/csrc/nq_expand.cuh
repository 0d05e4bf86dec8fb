// nq_expand.cuh — fused N-Queens evaluation + child generation on the device.
//
// Restates, on the GPU, evaluate_gpu (nqueens_gpu_chpl.chpl:97-123) followed by generate_children
// (:126-149): for every parent p of a chunk, in order, and every slot j = depth..N-1, in order, whose
// queen is safe, emit the child {depth+1, board with board[depth] <=> board[j]}; a parent with
// depth == N counts as one explored solution.  The children come out PACKED and IN THE REFERENCE'S
// ORDER, so a pool that appends them is byte-identical to the reference's pool after the same round.
//
// Three kernels per chunk:
//   K1 nq_expand_count : the evaluator of nq_kernel.cuh, but instead of N label bytes per parent it
//                        writes one 32-bit child mask (bit j <=> child j exists) and per-tile totals
//   K2 scan_tiles      : exclusive scan of the per-tile child counts (one CTA)
//   K3 nq_expand_write : per tile, exclusive scan of the per-parent counts, children built in shared
//                        memory as a contiguous byte image and written with one TMA bulk store plus
//                        < 16 head / tail bytes (21-byte records land at arbitrary alignment)
#pragma once
#include "nq_kernel.cuh"

namespace tsb {

struct ExpandCounters {
  unsigned long long children;   // total children of the chunk (written by K2)
  unsigned long long solutions;  // parents with depth == N (accumulated by K1)
};

// ------------------------------------------------------------------------------------------- K1
template <int N>
using NqCountSmem = TileSmem<NQ_STAGES, NQ_TILE * NQ_REC, NQ_TILE * 4>;

template <int N, int Q>
__device__ __forceinline__ uint32_t nq_child_mask(NqParent<N, Q, 0>& p) {
  const uint32_t S = ~p.U;
  uint32_t cm = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const uint32_t x = shf_r_wrap(S, 0u, p.amt[k]) & 1u;  // bit board[k] of the safe-value mask
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(cm) : "r"(x), "r"(1u << k));  // cm |= x << k on the FMA pipe
  }
  return cm & shl_clamp(0xFFFFFFFFu, p.depth);  // only slots k >= depth exist; depth >= 32 cannot occur
}

template <int N>
__device__ __forceinline__ void nq_count_tile(const uint8_t* in_tile, uint8_t* out_tile, int records, long long tile,
                                              int* __restrict__ tile_sums, ExpandCounters* __restrict__ ctr,
                                              int* red /* shared, 8 ints */) {
  const int t = threadIdx.x;
  const uint32_t* in_w = reinterpret_cast<const uint32_t*>(in_tile) + 21 * t;
  uint32_t w[21];
#pragma unroll
  for (int i = 0; i < 21; i++) w[i] = in_w[i];
  NqParent<N, 0, 0> p0;
  NqParent<N, 1, 0> p1;
  NqParent<N, 2, 0> p2;
  NqParent<N, 3, 0> p3;
  p0.init(w);
  p1.init(w);
  p2.init(w);
  p3.init(w);
  const uint32_t dmax = max(max(p0.depth, p1.depth), max(p2.depth, p3.depth));
#pragma unroll
  for (int j = 0; j < (N + 3) / 4; j++) {
    if (dmax > 4u * j) {
      switch (j) {
#define TSB_ROWS(J)                          \
  case J:                                    \
    p0.template rows<4 * J, 4 * J + 4>();    \
    p1.template rows<4 * J, 4 * J + 4>();    \
    p2.template rows<4 * J, 4 * J + 4>();    \
    p3.template rows<4 * J, 4 * J + 4>();    \
    break;
        TSB_ROWS(0) TSB_ROWS(1) TSB_ROWS(2) TSB_ROWS(3) TSB_ROWS(4)
#undef TSB_ROWS
      }
    }
  }
  uint32_t cm[4] = {nq_child_mask<N, 0>(p0), nq_child_mask<N, 1>(p1), nq_child_mask<N, 2>(p2),
                    nq_child_mask<N, 3>(p3)};
  const uint32_t dep[4] = {p0.depth, p1.depth, p2.depth, p3.depth};
  int cnt = 0, leaves = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const bool valid = 4 * t + q < records;
    if (!valid || dep[q] >= (uint32_t)N) cm[q] = 0;
    if (valid && dep[q] == (uint32_t)N) leaves++;
    cnt += __popc(cm[q]);
  }
  reinterpret_cast<uint4*>(out_tile)[t] = make_uint4(cm[0], cm[1], cm[2], cm[3]);
  // tile totals: warp shuffle reduction, then 4 warps through shared memory
  int packed = cnt | (leaves << 20);  // cnt <= 512*20 < 2^20
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xFFFFFFFFu, packed, o);
  __syncthreads();  // `red` may still be read by the previous tile's thread 0
  if ((t & 31) == 0) red[t >> 5] = packed;
  __syncthreads();
  if (t == 0) {
    const int tot = red[0] + red[1] + red[2] + red[3];
    tile_sums[tile] = tot & 0xFFFFF;
    if (tot >> 20) atomicAdd(&ctr->solutions, static_cast<unsigned long long>(tot >> 20));
  }
}

template <int N>
__global__ void __launch_bounds__(NQ_THREADS) nq_expand_count_kernel(const uint8_t* __restrict__ parents,
                                                                    uint8_t* __restrict__ cmask, long long count,
                                                                    int* __restrict__ tile_sums,
                                                                    ExpandCounters* __restrict__ ctr) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  NqCountSmem<N>& sm = *reinterpret_cast<NqCountSmem<N>*>(smem_raw);
  __shared__ int red[8];
  run_tile_pipeline<NQ_STAGES, NQ_TILE, NQ_REC, 4>(
      sm, parents, cmask, count, [&](const uint8_t* in_tile, uint8_t* out_tile, int n, long long tile) {
        nq_count_tile<N>(in_tile, out_tile, n, tile, tile_sums, ctr, red);
      });
}

// ------------------------------------------------------------------------------------------- K2
// exclusive scan of n ints in place (n <= a few 10^4), total into ctr->children; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int* __restrict__ v, int n, ExpandCounters* __restrict__ ctr) {
  __shared__ int warp_tot[32];
  __shared__ long long carry_s;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const int x = i < n ? v[i] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int wt = warp_tot[lane], wi = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= o) wi += y;
      }
      warp_tot[lane] = wi - wt;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const long long carry = carry_s;
    // child offsets of one chunk fit in 32 bits only up to 2^31 children; chunks are capped accordingly
    if (i < n) v[i] = static_cast<int>(carry + warp_tot[wid] + incl - x);
    __syncthreads();
    if (t == 1023) carry_s = carry + warp_tot[wid] + incl;
    __syncthreads();
  }
  if (t == 0) ctr->children = static_cast<unsigned long long>(carry_s);
}

// ------------------------------------------------------------------------------------------- K3
constexpr int EXP_CAP = 1536;  // children of one tile that fit the shared staging image (average is ~512)

struct NqWriteSmem {
  alignas(128) uint8_t in[NQ_TILE * NQ_REC];
  alignas(128) uint8_t stage[EXP_CAP * NQ_REC + 32];
  alignas(8) uint64_t full;
  int warp_tot[4];
};

template <int N>
__global__ void __launch_bounds__(NQ_THREADS) nq_expand_write_kernel(const uint8_t* __restrict__ parents,
                                                                    const uint32_t* __restrict__ cmask,
                                                                    const int* __restrict__ tile_off,
                                                                    long long count, uint8_t* __restrict__ children) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  NqWriteSmem& sm = *reinterpret_cast<NqWriteSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const long long tiles = (count + NQ_TILE - 1) / NQ_TILE;
  if (t == 0) {
    mbar_init(&sm.full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long first = tile * NQ_TILE;
    const int records = static_cast<int>(count - first < NQ_TILE ? count - first : NQ_TILE);
    // ---- stage the parents of this tile
    if (records == NQ_TILE) {
      if (t == 0) {
        mbar_arrive_expect_tx(&sm.full, NQ_TILE * NQ_REC);
        bulk_g2s(sm.in, parents + first * NQ_REC, NQ_TILE * NQ_REC, &sm.full);
      }
      mbar_wait(&sm.full, phase);
      phase ^= 1;
    } else {
      for (int i = t; i < records * NQ_REC; i += NQ_THREADS) sm.in[i] = parents[first * NQ_REC + i];
      __syncthreads();
    }
    // ---- per-parent counts and their exclusive scan over the tile (4 parents per thread, in order)
    uint4 cmv = make_uint4(0, 0, 0, 0);
    if (4 * t < records) cmv = reinterpret_cast<const uint4*>(cmask + first)[t];  // zero beyond `records` (K1)
    uint32_t cm[4] = {cmv.x, cmv.y, cmv.z, cmv.w};
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (4 * t + q >= records) cm[q] = 0;  // entries past the chunk's end were never written by K1
    const int mine = __popc(cm[0]) + __popc(cm[1]) + __popc(cm[2]) + __popc(cm[3]);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) sm.warp_tot[wid] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < wid) woff += sm.warp_tot[i];
      total += sm.warp_tot[i];
    }
    int pos = woff + incl - mine;  // index (within the tile) of this thread's first child
    const long long g_byte0 = (static_cast<long long>(tile_off[tile])) * NQ_REC;  // byte offset in `children`
    uint8_t* gdst = children + g_byte0;
    const bool staged = total <= EXP_CAP;
    // staging image starts at the same 16-byte phase as the global destination
    uint8_t* sdst = sm.stage + (reinterpret_cast<uintptr_t>(gdst) & 15);
    uint8_t* dst = staged ? sdst : gdst;
    // ---- build the children: copy the parent's 21 bytes, then patch depth and the two swapped queens
#pragma unroll
    for (int q = 0; q < 4; q++) {
      uint32_t m = cm[q];
      if (m) {
        const uint8_t* src = sm.in + (4 * t + q) * NQ_REC;
        uint8_t b[NQ_REC];
#pragma unroll
        for (int i = 0; i < NQ_REC; i++) b[i] = src[i];
        const int depth = b[0];
        const uint8_t qd = src[1 + depth];  // board[depth]
        while (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1;
          uint8_t* c = dst + static_cast<long long>(pos) * NQ_REC;
#pragma unroll
          for (int i = 0; i < NQ_REC; i++) c[i] = b[i];
          c[0] = static_cast<uint8_t>(depth + 1);
          c[1 + depth] = src[1 + k];  // child.board[depth] <=> child.board[k]
          c[1 + k] = qd;
          pos++;
        }
      }
    }
    if (staged) {
      fence_async_smem();
      __syncthreads();
      // head (< 16 B) and tail (< 16 B) by byte stores, the 16-byte aligned middle by one bulk store
      const int bytes = total * NQ_REC;
      const int head = min(bytes, static_cast<int>((16 - (reinterpret_cast<uintptr_t>(gdst) & 15)) & 15));
      const int mid = (bytes - head) & ~15;
      const int tail = bytes - head - mid;
      if (t < head) gdst[t] = sdst[t];
      if (t >= 32 && t - 32 < tail) gdst[head + mid + (t - 32)] = sdst[head + mid + (t - 32)];
      if (t == 0 && mid > 0) {
        bulk_s2g(gdst + head, sdst + head, static_cast<uint32_t>(mid));
        bulk_commit();
        bulk_wait_read<0>();  // the staging image is rewritten by the next tile
      }
    }
    __syncthreads();
  }
  if (t == 0) bulk_wait_all();
}

}  // namespace tsb
