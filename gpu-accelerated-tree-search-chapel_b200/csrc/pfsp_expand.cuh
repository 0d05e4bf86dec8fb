// pfsp_expand.cuh — PFSP bound evaluation + child generation on the device.
//
// Restates, on the GPU, evaluate_gpu (pfsp_gpu_chpl.chpl:192-270) followed by generate_children (:273-303):
// for every parent of a chunk, in order, and every slot j = limit1+1 .. jobs-1, in order: if the child is a
// leaf (depth + 1 == jobs) it counts as an explored solution and its bound competes for `best`; otherwise,
// if its bound is < best, the child {depth+1, limit1+1, prmu with prmu[depth] <=> prmu[j]} is emitted.  The
// children come out PACKED and IN THE REFERENCE'S ORDER.
//
// `best` is the value at launch for the whole chunk, exactly as in the reference's kernels (:385).  The
// reference's generate_children lowers `best` while it walks the chunk (a leaf with a smaller bound), which
// changes what the REST of the chunk pushes; the kernels report the minimum leaf bound of the chunk, and when
// it is below the launch value the host redoes that one round through the evaluate entry point and the
// sequential rule (tsb200_api.cu) — the device result is simply not committed (the chunk is read in place and
// the children land above the pool's top).  With --ub 1 (`best` = optimum) that never happens.
//
// Same two-kernel shape as nq_expand.cuh:
//   pfsp_expand_count_lb1 / _lb2 : the bound kernels of pfsp_kernels.cuh with a different epilogue — one
//                                  32-bit child mask per parent, one child count per tile, leaf statistics
//   pfsp_expand_build            : offsets of the CTA's own tiles (prologue), then per tile of 128 parents:
//                                  children copied word-wise into a shared-memory image (one thread per child,
//                                  11 x 8-byte loads/stores + 4 patched words) and written by one TMA bulk store
#pragma once
#include "expand_common.cuh"
#include "pfsp_kernels.cuh"

namespace tsb {

constexpr int PF_EXP_CAP = 256;  // children per pass of the staging image (a tile of 128 parents averages ~130)

// generic tile loop over the pieces of a round: f(in_tile, lin, abs_tile, lo, hi) is called by all threads for
// every tile of this CTA (lin = first, first+stride, ...); a __syncthreads follows each call
template <int STAGES, int TILE, int REC, typename F>
__device__ __forceinline__ void run_piece_tiles(uint8_t* in /* STAGES x TILE*REC, 128-B aligned */, uint64_t* full,
                                                const uint8_t* __restrict__ arena, const ExpandParams& prm, F&& f) {
  constexpr uint32_t IN_BYTES = TILE * REC;
  const int t = threadIdx.x;
  const int first = blockIdx.x, stride = gridDim.x;
  if (t == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int lin, int s) {  // thread 0
    long long at, lo, hi;
    piece_of(prm, lin, TILE, at, lo, hi);
    const uint32_t nb = tile_load_bytes(at, hi, TILE, REC);
    mbar_arrive_expect_tx(&full[s], nb);
    // (a half tile that lies wholly past the piece's end loads nothing: the phase completes on the arrival alone)
    if (nb) bulk_g2s(in + s * IN_BYTES, arena + at * IN_BYTES, nb, &full[s]);  // default L2 policy: the build kernel re-reads it
  };
  if (t == 0)
    for (int s = 0; s < STAGES; s++)
      if (first + s * stride < prm.n_tiles) issue(first + s * stride, s);
  unsigned it = 0;
  for (int lin = first; lin < prm.n_tiles; lin += stride, it++) {
    const int s = it % STAGES;
    long long at, lo, hi;
    piece_of(prm, lin, TILE, at, lo, hi);
    mbar_wait(&full[s], (it / STAGES) & 1u);
    f(in + s * IN_BYTES, lin, at, lo, hi);
    __syncthreads();
    if (t == 0 && lin + STAGES * stride < prm.n_tiles) issue(lin + STAGES * stride, s);
  }
}

// tile totals + leaf statistics of one tile, from per-thread values
__device__ __forceinline__ void pf_tile_totals(int* red /* 8 ints, shared */, int cnt, int leaves, int lin,
                                               int* __restrict__ tile_sums, unsigned& my_solutions) {
  const int t = threadIdx.x;
  int packed = cnt | (leaves << 16);  // children of a tile <= 128*20 < 2^16, leaves <= 128*20
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xFFFFFFFFu, packed, o);
  if ((t & 31) == 0) red[t >> 5] = packed;
  __syncthreads();
  if (t == 0) {
    const int tot = red[0] + red[1] + red[2] + red[3];
    tile_sums[lin] = tot & 0xFFFF;
    my_solutions += static_cast<unsigned>(tot >> 16);
  }
}

// ------------------------------------------------------------------------------------------- count: lb1 / lb1_d
struct Lb1CountSmem {
  Lb1Smem core;  // tiles.buf[0..1]: two input stages
  uint32_t cmask[PF_TILE];
  int red[8];
};

template <int KIND, int M, bool SIMD>
__global__ void __launch_bounds__(PF_THREADS) pfsp_expand_count_lb1_kernel(const uint8_t* __restrict__ arena,
                                                                          const __grid_constant__ ExpandParams prm,
                                                                          const PfspLb1Tables* __restrict__ tables,
                                                                          uint32_t* __restrict__ cmask,
                                                                          int* __restrict__ tile_sums,
                                                                          ExpandState* __restrict__ st) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Lb1CountSmem& sm = *reinterpret_cast<Lb1CountSmem*>(smem_raw);
  stage_blob(&sm.core.tab, tables, sizeof(PfspLb1Tables), &sm.core.tab_bar);
  const int jobs = sm.core.tab.jobs, best = prm.best;
  unsigned my_solutions = 0;
  run_piece_tiles<2, PF_TILE, PF_REC>(
      sm.core.tiles.buf[0], sm.core.tiles.full, arena, prm,
      [&](const uint8_t* in_tile, int lin, long long at, long long lo, long long hi) {
        const int rec_lo = static_cast<int>(lo - at * PF_TILE), rec_hi = static_cast<int>(hi - at * PF_TILE);
        uint32_t m = 0, live = 0;
        int leaf_lb = 0x7FFFFFFF;
        const int p = lb1_compute_tile<KIND, M, SIMD, false>(sm.core, in_tile, rec_lo, rec_hi,
                                                [&](int, int limit1, int g, const int(&v)[4]) {
#pragma unroll
                                                  for (int c = 0; c < 4; c++) {
                                                    const int k = 4 * g + c;
                                                    if (k > limit1) {
                                                      live |= 1u << k;
                                                      if (v[c] < best) m |= 1u << k;
                                                      leaf_lb = min(leaf_lb, v[c]);
                                                    }
                                                  }
                                                });
        int leaves = 0;
        if (live) {  // p is a valid parent with at least one slot
          const int depth = reinterpret_cast<const int32_t*>(in_tile)[22 * p];
          if (depth + 1 == jobs) {  // every child is a leaf (pfsp_gpu_chpl.chpl:283-288)
            leaves = __popc(live);
            m = 0;
            if (leaf_lb < best) atomicMin(&st->best, leaf_lb);
          }
        }
        sm.cmask[p] = m;  // p runs over all 128 records of the tile (the depth sort is a permutation)
        pf_tile_totals(sm.red, __popc(m), leaves, lin, tile_sums, my_solutions);  // (syncs: cmask complete)
        cmask[static_cast<long long>(lin) * PF_TILE + threadIdx.x] = sm.cmask[threadIdx.x];
      });
  if (threadIdx.x == 0 && my_solutions) atomicAdd(&st->solutions, static_cast<unsigned long long>(my_solutions));
}

// ------------------------------------------------------------------------------------------- count: lb2
template <int M>
struct Lb2CountSmem {
  Lb2Smem<M> core;  // tiles.in[0..1]: two input stages
  uint32_t cmask[LB2_TILE];
  uint32_t leafs[LB2_TILE];
  int red[8];
};

// (tiles of LB2_TILE = 64 parents: two linear tiles of this kernel make one 128-parent tile of the build
// kernel, so masks and counts are laid out for PF_TILE: mask index = lin * 64 + t, tile_sums[lin / 2] is
// accumulated with an atomicAdd — the host clears tile_sums before the launch)
template <int M, typename CT>
__global__ void __launch_bounds__(PF_THREADS) pfsp_expand_count_lb2_kernel(const uint8_t* __restrict__ arena,
                                                                          const __grid_constant__ ExpandParams prm,
                                                                          const PfspLb1Tables* __restrict__ tables1,
                                                                          const __grid_constant__ CT C,
                                                                          uint32_t* __restrict__ cmask,
                                                                          int* __restrict__ tile_sums,
                                                                          ExpandState* __restrict__ st) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Lb2CountSmem<M>& sm = *reinterpret_cast<Lb2CountSmem<M>*>(smem_raw);
  lb2_stage_tables(sm.core, tables1, C);
  const int jobs = sm.core.tab1.jobs, best = prm.best;
  unsigned my_solutions = 0;
  // this kernel walks the round in half tiles: linear half-tile h covers records [64h, 64h+64) of linear tile h/2
  ExpandParams half = prm;
  for (int i = 0; i < half.n_pieces; i++) {
    // a piece's first build tile starts at first_tile*128 = (2*first_tile)*64
    half.piece[i].first_tile *= PF_TILE / LB2_TILE;
    half.piece[i].tile_cum *= PF_TILE / LB2_TILE;
  }
  half.n_tiles *= PF_TILE / LB2_TILE;
  run_piece_tiles<2, LB2_TILE, PF_REC>(
      sm.core.tiles.in[0], sm.core.tiles.full, arena, half,
      [&](const uint8_t* in_tile, int lin, long long at, long long lo, long long hi) {
        const int rec_lo = static_cast<int>(lo - at * LB2_TILE), rec_hi = static_cast<int>(hi - at * LB2_TILE);
        const int t = threadIdx.x;
        const int32_t* nodes = reinterpret_cast<const int32_t*>(in_tile);
        if (t < LB2_TILE) {
          sm.cmask[t] = 0;
          sm.leafs[t] = 0;  // (lb2_compute_tile starts with a barrier)
        }
        lb2_compute_tile<M>(
            sm.core, C, in_tile, rec_lo, rec_hi, best,
            [&](int p, int k, int lb) {
              if (nodes[22 * p] + 1 == jobs) {  // leaf child (pfsp_gpu_chpl.chpl:283-288)
                atomicOr(&sm.leafs[p], 1u << k);
                if (lb < best) atomicMin(&st->best, lb);
              } else if (lb < best) {
                atomicOr(&sm.cmask[p], 1u << k);
              }
            },
            [](int, int) {});
        __syncthreads();
        const uint32_t m = t < LB2_TILE ? sm.cmask[t] : 0u;
        int packed = __popc(m) | ((t < LB2_TILE ? __popc(sm.leafs[t]) : 0) << 16);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xFFFFFFFFu, packed, o);
        if ((t & 31) == 0) sm.red[t >> 5] = packed;
        __syncthreads();
        if (t == 0) {
          const int tot = sm.red[0] + sm.red[1] + sm.red[2] + sm.red[3];
          if (tot & 0xFFFF) atomicAdd(&tile_sums[lin / (PF_TILE / LB2_TILE)], tot & 0xFFFF);
          my_solutions += static_cast<unsigned>(tot >> 16);
        }
        if (t < LB2_TILE) cmask[static_cast<long long>(lin) * LB2_TILE + t] = m;
      });
  if (threadIdx.x == 0 && my_solutions) atomicAdd(&st->solutions, static_cast<unsigned long long>(my_solutions));
}

// ------------------------------------------------------------------------------------------- build
struct PfBuildSmem {
  alignas(128) uint8_t in[2][PF_TILE * PF_REC];
  alignas(128) uint32_t mask[2][PF_TILE];
  alignas(128) uint8_t stage[PF_EXP_CAP * PF_REC + 32];
  alignas(8) uint64_t full[2];
  uint16_t item[PF_EXP_CAP];  // (record << 5) | slot, in child order
  int warp_tot[4];
  ScanSmem scan;
};

__global__ void __launch_bounds__(PF_THREADS) pfsp_expand_build_kernel(const uint8_t* __restrict__ arena,
                                                                      const __grid_constant__ ExpandParams prm,
                                                                      const uint32_t* __restrict__ cmask,
                                                                      const int* __restrict__ tile_sums,
                                                                      uint8_t* __restrict__ children,
                                                                      ExpandState* __restrict__ st,
                                                                      ExpandResult* __restrict__ res) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  PfBuildSmem& sm = *reinterpret_cast<PfBuildSmem*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  constexpr uint32_t IN_BYTES = PF_TILE * PF_REC;
  const int first = blockIdx.x, stride = gridDim.x;
  if (t == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint64_t pol = 0;
  if (t == 0) pol = policy_evict_first();
  auto issue = [&](int lin, int s) {  // thread 0: parents + masks of one tile
    long long at, lo, hi;
    piece_of(prm, lin, PF_TILE, at, lo, hi);
    const uint32_t nb = tile_load_bytes(at, hi, PF_TILE, PF_REC);
    mbar_arrive_expect_tx(&sm.full[s], nb + PF_TILE * 4);
    if (nb) bulk_g2s_stream(sm.in[s], arena + at * IN_BYTES, nb, &sm.full[s], pol);
    bulk_g2s_stream(sm.mask[s], cmask + static_cast<long long>(lin) * PF_TILE, PF_TILE * 4, &sm.full[s], pol);
  };
  if (t == 0) {
    if (first < prm.n_tiles) issue(first, 0);
    if (first + stride < prm.n_tiles) issue(first + stride, 1);
  }
  expand_own_offsets<PF_THREADS>(sm.scan, tile_sums, prm.n_tiles, first, stride);
  expand_publish(sm.scan, st, res, prm.epoch, 1);  // st->best restarts at INT_MAX every round
  unsigned it = 0;
  for (int lin = first; lin < prm.n_tiles; lin += stride, it++) {
    const int s = it & 1;
    mbar_wait(&sm.full[s], (it >> 1) & 1u);
    const uint32_t cm = sm.mask[s][t];
    const int mine = __popc(cm);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) sm.warp_tot[wid] = incl;
    if (t == 0) bulk_wait_read<0>();  // the previous tile's bulk store has drained the staging image
    __syncthreads();  // (A)
    int woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i < wid) woff += sm.warp_tot[i];
      total += sm.warp_tot[i];
    }
    const int pos0 = woff + incl - mine;  // index (within the tile) of this parent's first child
    uint8_t* const gtile = children + static_cast<long long>(sm.scan.own[it]) * PF_REC;
    for (int c0 = 0; c0 < total; c0 += PF_EXP_CAP) {  // windows of PF_EXP_CAP children
      const int cnt = min(PF_EXP_CAP, total - c0);
      if (c0 > 0 && t == 0) bulk_wait_read<0>();
      int pos = pos0 - c0;
      uint32_t m = cm;
      while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        if (pos >= 0 && pos < PF_EXP_CAP) sm.item[pos] = static_cast<uint16_t>((t << 5) | k);
        pos++;
      }
      __syncthreads();  // (B) items
      uint8_t* gdst = gtile + static_cast<long long>(c0) * PF_REC;
      const int phase = static_cast<int>(reinterpret_cast<uintptr_t>(gdst) & 15);  // 0 or 8 (children are 8-B aligned)
      uint8_t* sdst = sm.stage + phase;
      for (int c = t; c < cnt; c += PF_THREADS) {
        const int item = sm.item[c];
        const int r = item >> 5, k = item & 31;
        const int2* src = reinterpret_cast<const int2*>(sm.in[s] + r * PF_REC);
        int2* d = reinterpret_cast<int2*>(sdst + c * PF_REC);
        const int2 head = src[0];
#pragma unroll
        for (int i = 1; i < 11; i++) d[i] = src[i];
        d[0] = make_int2(head.x + 1, head.y + 1);  // depth + 1, limit1 + 1
        const int32_t* sp = reinterpret_cast<const int32_t*>(src) + 2;
        int32_t* dp = reinterpret_cast<int32_t*>(d) + 2;
        const int depth = head.x;
        const int a = sp[depth], b = sp[k];  // child.prmu[depth] <=> child.prmu[k]
        dp[depth] = b;
        dp[k] = a;
      }
      fence_async_smem();
      __syncthreads();  // (C) image complete
      const int bytes = cnt * PF_REC;
      const int head = min(bytes, (16 - phase) & 15);
      const int mid = (bytes - head) & ~15;
      const int tail = bytes - head - mid;
      if (t < head) gdst[t] = sdst[t];
      if (t >= 32 && t - 32 < tail) gdst[head + mid + (t - 32)] = sdst[head + mid + (t - 32)];
      if (t == 0 && mid > 0) {
        bulk_s2g(gdst + head, sdst + head, static_cast<uint32_t>(mid));
        bulk_commit();
      }
    }
    // a tile without children has no barrier after (A): without this one a fast warp could overwrite warp_tot
    // for tile it+1 while a slow warp still reads the totals of tile it (and thread 0 could re-arm full[s] twice
    // before a slow warp has tested the phase of tile it)
    if (total == 0) __syncthreads();
    if (t == 0 && lin + 2 * stride < prm.n_tiles) issue(lin + 2 * stride, s);
  }
  if (t == 0) bulk_wait_all();
}

}  // namespace tsb
