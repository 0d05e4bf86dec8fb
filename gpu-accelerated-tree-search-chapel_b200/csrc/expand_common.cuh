// expand_common.cuh — pieces shared by the "expand" kernels (evaluate + generate_children of one offload
// round on the device):
//   * the chunk is described by up to EXP_MAX_PIECES position ranges of a node ARENA (the device-resident
//     pool is a stack of extents inside one arena; a chunk = the newest nodes = the top extents, possibly
//     spanning holes), read IN PLACE.  Tiles are taken at ABSOLUTE arena positions (multiples of the tile
//     size), so that every tile load is one 16-byte aligned TMA bulk copy; records outside [lo, hi) are masked;
//   * two kernels per round, neither with a cross-tile dependency:
//       count  evaluates the parents, writes one child mask per parent and one child count per tile;
//       build  every CTA scans the tile counts (L2) for the offsets of its own tiles, re-reads its tiles (L2),
//              builds the children of a tile as a contiguous image in shared memory and stores it at its final
//              place, in the reference's order; CTA 0 publishes {children, solutions[, best]} to a host-mapped
//              record and re-arms the device counters — a round costs one stream synchronisation.
//     (A single-pass variant with a decoupled look-back — flat, 256-wide and two-level — was measured at
//     140-350 us per 4 Mi parents against ~70 us for this pipeline: persistent CTAs run in lock-step waves,
//     in-order commit with two tiles in flight per CTA left them waiting on each other half of the time.)
#pragma once
#include "tsb_ptx.cuh"

namespace tsb {

constexpr int EXP_MAX_PIECES = 8;

struct ExpandPiece {
  long long lo, hi;      // arena positions [lo, hi) of this piece (logical order = piece order)
  long long first_tile;  // absolute tile index of the tile that holds `lo`
  int tile_cum;          // linear index (within the round) of that tile
  int pad;
};
struct ExpandParams {
  ExpandPiece piece[EXP_MAX_PIECES];
  int n_pieces;
  int n_tiles;     // linear tiles of the round
  unsigned epoch;  // round number, echoed in the result record
  int best;        // PFSP: incumbent at launch (int32-clamped)
};

// device-side counters of a round, re-armed by the scan kernel
struct ExpandState {
  unsigned long long solutions;  // accumulated by the count kernel
  int best;                      // PFSP: running minimum over evaluated leaves (atomicMin)
  int pad;
};
// host-mapped (pinned) result of a round
struct ExpandResult {
  unsigned long long children;
  unsigned long long solutions;
  long long best;
  unsigned long long epoch;
};

// linear tile index of the round -> absolute tile + validity range
__device__ __forceinline__ void piece_of(const ExpandParams& prm, int lin, int tile_records, long long& abs_tile,
                                         long long& lo, long long& hi) {
  int k = 0;
#pragma unroll
  for (int i = 1; i < EXP_MAX_PIECES; i++)
    if (i < prm.n_pieces && lin >= prm.piece[i].tile_cum) k = i;
  abs_tile = prm.piece[k].first_tile + (lin - prm.piece[k].tile_cum);
  const long long t0 = abs_tile * tile_records;
  lo = prm.piece[k].lo > t0 ? prm.piece[k].lo : t0;
  hi = prm.piece[k].hi < t0 + tile_records ? prm.piece[k].hi : t0 + tile_records;
}

// Copy `bytes` bytes of a 16-byte aligned shared-memory image to an arbitrarily aligned global address:
// < 16 head / tail bytes one by one, the 16-byte aligned middle as STG.128 whose source words are realigned
// in registers by a funnel shift (the image does not have to share the destination's 16-byte phase, so it
// can be built before the destination is known).  Called by all `nthreads` threads of the CTA.
__device__ __forceinline__ void copy_image_to_global(uint8_t* gdst, const uint8_t* image, int bytes, int tid,
                                                     int nthreads) {
  const int head = min(bytes, static_cast<int>((16 - (reinterpret_cast<uintptr_t>(gdst) & 15)) & 15));
  const int nmid = (bytes - head) >> 4;
  const int tail = bytes - head - 16 * nmid;
  if (tid < head) gdst[tid] = image[tid];
  if (tid >= 32 && tid - 32 < tail) gdst[head + 16 * nmid + (tid - 32)] = image[head + 16 * nmid + (tid - 32)];
  const uint32_t* iw = reinterpret_cast<const uint32_t*>(image) + (head >> 2);
  const uint32_t sh = (head & 3) * 8;
  uint4* g4 = reinterpret_cast<uint4*>(gdst + head);
  for (int k = tid; k < nmid; k += nthreads) {
    const uint32_t* w = iw + 4 * k;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    g4[k] = make_uint4(shf_r_wrap(w0, w1, sh), shf_r_wrap(w1, w2, sh), shf_r_wrap(w2, w3, sh),
                       shf_r_wrap(w3, w4, sh));
  }
}

// bytes of a tile worth loading: up to the last valid record, rounded up to the 16-byte TMA granularity (a
// caller-owned chunk is never read more than 15 bytes past its end)
__device__ __forceinline__ uint32_t tile_load_bytes(long long abs_tile, long long hi, int tile_records, int rec) {
  const long long b = (hi - abs_tile * tile_records) * rec;
  const long long full = static_cast<long long>(tile_records) * rec;
  if (b <= 0) return 0u;  // a (half) tile that lies wholly past the piece's end: nothing to load
  return static_cast<uint32_t>(b >= full ? full : (b + 15) & ~15LL);
}

// Tile offsets without a scan kernel: every CTA of the build kernel scans the n tile counts itself (they sit in
// L2: 4 B per tile, 32 KB for a 4 Mi-parent round) while its first TMA loads are in flight, and keeps the
// offsets of its own tiles first, first+stride, ... in shared memory.  CTA 0 also publishes the totals of the
// round to the host-mapped record and re-arms the device counters.
constexpr int EXP_MAX_OWN = 256;  // tiles per CTA (host-checked)
struct ScanSmem {
  int own[EXP_MAX_OWN];  // offset of own tile i (= tile first + i*stride)
  int cnt[EXP_MAX_OWN];  // its child count
  int excl[256];         // children of all tiles before thread t's block of tile counts
  int part[32];
  long long total;
};
template <int THREADS>
__device__ __forceinline__ void expand_own_offsets(ScanSmem& sc, const int* __restrict__ tile_sums, int n, int first,
                                                   int stride) {
  static_assert(THREADS <= 256, "excl[] holds one entry per thread");
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (n + THREADS - 1) / THREADS;
  const int lo = min(n, t * per), hi = min(n, lo + per);
  // pass 1: sum of this thread's block (independent loads, the adds trail behind)
  int run = 0;
#pragma unroll 16
  for (int j = lo; j < hi; j++) run += __ldg(&tile_sums[j]);
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) sc.part[wid] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int i = 0; i < THREADS / 32; i++) {
    if (i < wid) woff += sc.part[i];
    total += sc.part[i];
  }
  sc.excl[t] = woff + incl - run;
  if (t == 0) sc.total = total;
  __syncthreads();
  // pass 2: one thread per own tile: block prefix + the counts before it inside its block
  for (int i = t; first + i * stride < n; i += THREADS) {
    const int j = first + i * stride;
    const int blk = j / per;
    int off = sc.excl[blk];
#pragma unroll 16
    for (int x = blk * per; x < j; x++) off += __ldg(&tile_sums[x]);
    sc.own[i] = off;
    sc.cnt[i] = __ldg(&tile_sums[j]);
  }
  __syncthreads();
}
__device__ __forceinline__ void expand_publish(const ScanSmem& sc, ExpandState* st, ExpandResult* res, unsigned epoch,
                                               int reset_best) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    res->children = static_cast<unsigned long long>(sc.total);
    res->solutions = st->solutions;
    res->best = st->best;
    __threadfence_system();  // the host may be polling `epoch` while this kernel still builds the children:
    *reinterpret_cast<volatile unsigned long long*>(&res->epoch) = epoch;  // payload first, then the flag
    st->solutions = 0ull;  // the next round's count kernel starts after this kernel
    if (reset_best) st->best = 0x7FFFFFFF;
  }
}

}  // namespace tsb
