// nq_rounds_ll.cuh — the persistent multi-round N-Queens kernel without fences on its critical path.
//
// nq_rounds.cuh (v2) orders round r+1 after round r with a release fence (MEMBAR.ALL.GPU, 1 500-3 500 cycles on
// B200 under load) and a "done" flag exchange among all CTAs, on top of the exchange that gathers the child counts:
// 7 us per round, of which ~1 us is work.  Here the pool lives, while the kernel runs, in a SELF-VALIDATING format
// (the "LL" idea of NCCL's low-latency protocol):
//
//   fat node = 8 x 8-byte words; word i = data32[i] | epoch << 32            (64 B per node, 16-byte aligned)
//   data32[0..5] = the 21 node bytes (depth, board[0..19]),  data32[6..7] = the node's aux word (nq_rounds.cuh:
//   diagonal masks, child mask, leaf flag)
//
// Every 8-byte word is written by one store (an element of a st.v2.u64) and is therefore seen whole or not at all;
// a reader that expects the children of round r polls the words of its slice until all eight carry r's epoch.  No
// fence, no "done" flags: the data is its own flag, and a round costs ONE flag exchange (the child counts) plus one
// store -> L2 -> poll hop for the nodes.  Epochs are 32 bits and never repeat, so stale words cannot alias.
//
// Nodes that are NOT children of the previous round (the chunk reaches below the newest layer when a round produced
// fewer than M children) carry the epoch of the round that stored them.  Every CTA keeps the LAYER STACK of the pool
// — (first position, epoch) of every round's children that are still in the pool, a deterministic function of the
// round totals like the rest of the pool state — and validates every piece against the epoch of the layer its
// position lies in.  Nodes that were in the pool when the kernel was launched form one trusted layer (a kernel
// boundary orders them).  (A first version fenced old rounds with a side warp and checked per-CTA fence flags before
// reading old nodes: 1 300 cycles on the critical path of every round that reaches below the newest layer.)
//
// Measured alternatives for the one remaining exchange (all ~130 CTAs polling the 2 G count words: 2 500-4 000 cycles):
// a dedicated aggregator CTA that scans the counts and writes every worker its own result line (two contention-free
// hops: 2 900-3 400 cycles, no gain), per-reader rows that only their owner polls (no gain at 128 CTAs), CTAs of 512
// threads (slower scan), fewer CTAs (cheaper exchange, more work per CTA: best at 128 of 148 SMs), a sentinel
// poll (one lane per warp watches one piece and the full sweeps start when it has arrived: the extra hop costs more
// than the poll traffic it saves, 3-12 % slower with one to four pools).
//
// The plain 21-byte arena is converted
// to and from the fat arena by nq_fat_import / nq_fat_export (whole pool, only when the host needs the plain form:
// drain, steal, pool_step, arena growth).
//
// One launch serves up to LL_MAX_POOLS INDEPENDENT pools (grid (G, pools), LlMultiParams): even so a pool's round stays
// a chain of L2 round trips with ~1 us of work in between, and nothing inside one pool can fill the waits — another
// pool's CTA on the same SM can.  Measured on the N = 17 search at M = 50000 (one B200): 0.742 s with one pool (128
// CTAs), 0.496 s with two (148 + 148), 0.411 s with three (3 x 98), 0.413 s with four (4 x 74 CTAs of 768 parents).
#pragma once
#include "nq_rounds.cuh"

namespace tsb {

constexpr int LL_T = 256;                    // threads per CTA
// parents per worker thread (PPT): 2 with one or two pools per launch (128 / 148 CTAs per pool), 3 with three or four
// (74 CTAs per pool: a round's count exchange among 74 CTAs costs half of one among 148, tools/flag_exchange.py, and
// every CTA brings 1.5x the work to hide it behind)
__host__ __device__ constexpr int ll_slice(int ppt) { return LL_T * ppt; }          // parents per CTA per round
// children per window of the staging buffer (a CTA's share of a round averages ~0.8 children per parent; a dense
// share takes several windows); 512 in the three-CTAs-per-SM build (MINB = 3: 64 KB of shared memory per CTA)
__host__ __device__ constexpr int ll_cap(int ppt, int minb) { return minb >= 3 ? 512 : ppt <= 2 ? 2048 : 1024; }
constexpr int LL_WORDS = 8;                  // 8-byte words per fat node
constexpr int LL_LAYERS = 1024;              // layers of the pool a CTA tracks (more: the kernel leaves and is relaunched)
constexpr unsigned LL_TRUSTED = 0u;          // layer epoch of the nodes that were in the pool at launch (epochs start at 1)

struct FatNode {
  unsigned long long w[LL_WORDS];
};
struct LlSync {
  unsigned long long slot[2][2 * RND_MAX_CTAS];  // by round parity, one per SUB-slice: epoch << 32 | leaves << 20 | children
  unsigned abort;
};
constexpr int LL_MAX_POOLS = 4;  // independent pools one launch can serve (blockIdx.y)
struct LlParams {
  FatNode* fat;
  long long cap;    // nodes the fat arena holds
  long long size0;  // nodes in the pool at launch: positions [0, size0), all stored before the launch
  unsigned epoch0;  // last epoch used so far (the import kernel's tag or the previous launch's last round)
  int m, M;
  long long max_rounds;
  int prof;
  LlSync* sync;
  RoundsState* state;
};
// Several INDEPENDENT pools in one cooperative launch: grid (G, pools), the CTAs of row y run the rounds of pool y and
// never look at another row.  A round of one pool is a chain of L2 round trips (count exchange, store -> poll) with
// ~1 us of work in between; two co-resident CTAs per SM working on different pools fill each other's waits.
struct LlMultiParams {
  LlParams pool[LL_MAX_POOLS];
};

// ---- CTA barriers on a named barrier (kept from the version that had a side warp outside them)
__device__ __forceinline__ void ll_bar(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }
__device__ __forceinline__ bool ll_bar_or(int threads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(static_cast<uint32_t>(pred)), "r"(threads)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void ld_fat2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// (no "memory" clobber: the compiler may move the shared-memory loads that feed later stores across this one)
__device__ __forceinline__ void st_fat2(unsigned long long* p, unsigned long long a, unsigned long long b) {
  asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b));
}

// ---- plain arena <-> fat arena (one thread per node; not performance critical: the whole pool, once per hand-over)
template <int N>
__global__ void nq_fat_import_kernel(const uint8_t* __restrict__ arena, FatNode* __restrict__ fat, long long size,
                                     unsigned epoch) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= size) return;
  const uint8_t* node = arena + p * NQ_REC;
  uint32_t d[LL_WORDS] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < NQ_REC; i++) d[i >> 2] |= static_cast<uint32_t>(node[i]) << (8 * (i & 3));
  const unsigned long long aux = nq_aux_of_node<N>(node);
  d[6] = static_cast<uint32_t>(aux);
  d[7] = static_cast<uint32_t>(aux >> 32);
  const unsigned long long e = static_cast<unsigned long long>(epoch) << 32;
  for (int i = 0; i < LL_WORDS; i += 2) st_fat2(&fat[p].w[i], d[i] | e, d[i + 1] | e);
}
__global__ void nq_fat_export_kernel(const FatNode* __restrict__ fat, uint8_t* __restrict__ arena, long long size) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= size) return;
  uint8_t* node = arena + p * NQ_REC;
  for (int i = 0; i < 6; i++) {
    const uint32_t d = static_cast<uint32_t>(fat[p].w[i]);
    for (int b = 0; b < 4 && 4 * i + b < NQ_REC; b++) node[4 * i + b] = static_cast<uint8_t>(d >> (8 * b));
  }
}

// warp 0: until all n slots carry `epoch`; sums of {leaves << 32 | children} over all slots and over the slots
// before k0 / before k1 (valid in every lane); false = abort
__device__ __forceinline__ bool warp_gather_slots2(const unsigned long long* slot, int n, int k0, int k1, unsigned epoch,
                                                   unsigned* abort_flag, unsigned long long& before0,
                                                   unsigned long long& before1, unsigned long long& all) {
  const int lane = threadIdx.x & 31;
  SpinGuard guard;
  for (;;) {
    bool ok = true;
    before0 = before1 = all = 0;
    for (int i = 2 * lane; i < n; i += 64) {
      unsigned long long v0, v1;
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(slot + i) : "memory");
      const bool has1 = i + 1 < n;
      ok &= static_cast<unsigned>(v0 >> 32) == epoch && (!has1 || static_cast<unsigned>(v1 >> 32) == epoch);
      const unsigned long long p0 = (v0 & 0xFFFFFull) | ((v0 >> 20) & 0xFFFull) << 32;
      const unsigned long long p1 = has1 ? (v1 & 0xFFFFFull) | ((v1 >> 20) & 0xFFFull) << 32 : 0ull;
      all += p0 + p1;
      if (i < k0) before0 += p0;
      if (i + 1 < k0) before0 += p1;
      if (i < k1) before1 += p0;
      if (i + 1 < k1) before1 += p1;
    }
    if (__all_sync(0xFFFFFFFFu, ok)) break;
    if (__any_sync(0xFFFFFFFFu, guard.expired(abort_flag))) return false;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before0 += __shfl_xor_sync(0xFFFFFFFFu, before0, o);
    before1 += __shfl_xor_sync(0xFFFFFFFFu, before1, o);
    all += __shfl_xor_sync(0xFFFFFFFFu, all, o);
  }
  return true;
}

template <int T, int PPT, int MINB>
struct LlSmem {
  alignas(16) uint32_t parent[T * PPT][8];           // the slice: data32[0..7] of every parent
  alignas(16) uint32_t stage[ll_cap(PPT, MINB)][8];  // the window's children: data32[0..7]
  alignas(16) uint16_t item[T * PPT * 20];     // (record << 5) | slot, in child order
  unsigned long long warp_tot64[T / 32];
  unsigned long long red[3];
  long long lay_start[LL_LAYERS];  // the pool's layers, bottom to top: first position ...
  unsigned lay_epoch[LL_LAYERS];   // ... and the epoch its nodes were stored with (LL_TRUSTED: before the launch)
};

// child `item` of the slice (pure data: no alignment games in the fat format) -> its eight data words
template <int N>
__device__ __forceinline__ void ll_build_child(const uint32_t (*parent)[8], int item, uint32_t (&c)[8]) {
  const int r = item >> 5, k = item & 31;
  const uint4 lo = *reinterpret_cast<const uint4*>(parent[r]), hi = *reinterpret_cast<const uint4*>(parent[r] + 4);
  uint32_t P[6] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y & 0xFFu};
  const uint32_t depth = P[0] & 0xFFu;
  const uint32_t p1 = 1u + depth, p2 = 1u + static_cast<uint32_t>(k);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(parent[r]);
  const uint32_t v = pb[p2];  // the queen placed on row `depth`
  const uint32_t D = static_cast<uint32_t>(pb[p1]) ^ v;
  const uint32_t x1 = D << ((p1 & 3u) * 8u), x2 = D << ((p2 & 3u) * 8u);
  const uint32_t w1 = p1 >> 2, w2 = p2 >> 2;
#pragma unroll
  for (uint32_t j = 0; j < 6; j++) P[j] ^= (j == w1 ? x1 : 0u) ^ (j == w2 ? x2 : 0u);
  P[0] += 1u;  // depth + 1
  const unsigned long long w = static_cast<unsigned long long>(hi.z) | static_cast<unsigned long long>(hi.w) << 32;
  const uint32_t ld = static_cast<uint32_t>(w) & 0xFFFFFu, rd = static_cast<uint32_t>(w >> 20) & 0xFFFFFu;
  const uint32_t bit = 1u << (v & 31u);
  const uint32_t ld2 = ((ld | bit) << 1) & ((1u << N) - 1u), rd2 = (rd | bit) >> 1;
  NqParent<N, 0, 0> cp;
  cp.init(P);  // depth + 1, shift amounts = the child's board
  cp.U = ld2 | rd2;
  const uint32_t cm = nq_child_mask<N, 0>(cp);  // slots >= depth + 1 whose value is safe (none for a leaf)
  const unsigned long long ca = nq_aux_pack(ld2, rd2, cm, depth + 1u == static_cast<uint32_t>(N));
#pragma unroll
  for (int j = 0; j < 6; j++) c[j] = P[j];
  c[6] = static_cast<uint32_t>(ca);
  c[7] = static_cast<uint32_t>(ca >> 32);
}

template <int N, int T, int MINB, int PPT>
__global__ void __launch_bounds__(T, MINB) nq_rounds_ll_kernel(const __grid_constant__ LlMultiParams mprm) {
  const LlParams& prm = mprm.pool[blockIdx.y];
  constexpr int LL_PPT = PPT, LL_CAP = ll_cap(PPT, MINB);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  LlSmem<T, PPT, MINB>& sm = *reinterpret_cast<LlSmem<T, PPT, MINB>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int k = blockIdx.x, G = gridDim.x;
  LlSync* const sy = prm.sync;
  FatNode* const fat = prm.fat;

  if (t == 0) {
    sm.lay_start[0] = 0;
    sm.lay_epoch[0] = LL_TRUSTED;
  }
  __syncthreads();
  int n_lay = prm.size0 > 0 ? 1 : 0;

  // ------------------------------------------------------------------------------------------ the workers
  long long size = prm.size0;
  unsigned epoch = prm.epoch0;
  unsigned long long rounds = 0, tot_parents = 0, tot_children = 0, tot_solutions = 0;
  int exit_code = RND_EXIT_PAUSE;
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = 0;
  const bool prof_on = prm.prof != 0 && k == 0 && t == 0;
#define TSB_PROF(i)                  \
  if (prof_on) {                     \
    const long long now = clock64(); \
    prof[i] += now - tp;             \
    tp = now;                        \
  }

  for (long long r = 0;; r++) {
    // ---- (0) the round's chunk: popBackBulk(m, M) (uniform decisions: every CTA holds the same state)
    if (size < prm.m) {
      exit_code = RND_EXIT_DONE;
      break;
    }
    if (r >= prm.max_rounds) {
      exit_code = RND_EXIT_PAUSE;
      break;
    }
    const long long n = size < prm.M ? size : prm.M;
    const long long s0 = size - n;  // position of the chunk's first parent = of the round's first child
    if (s0 + n * N > prm.cap) {     // worst case: every slot of every parent survives
      exit_code = RND_EXIT_SPACE;
      break;
    }
    if (n_lay >= LL_LAYERS) {  // (no room to record this round's children: start over with one trusted layer)
      exit_code = RND_EXIT_RELAUNCH;
      break;
    }
    ++epoch;
    if (prof_on) tp = clock64();
    // my share of the chunk: TWO sub-slices of n / 2G parents — number k from the bottom and number k from the top.
    // The bottom of a chunk holds the shallow nodes (many children), the top the deep ones (few): a single slice per
    // CTA left the bottom CTA with 3x the average children, and its build + 64-byte stores were the round's
    // critical path; pairing k with 2G-1-k evens the load without knowing it in advance.
    const int G2 = 2 * G;
    const unsigned n32 = static_cast<unsigned>(n), uG2 = static_cast<unsigned>(G2), uk = static_cast<unsigned>(k);
    // (n <= 768 G and k < G <= 256: the products fit 32 bits — four 64-bit divisions cost 700 cycles per round)
    const int a0 = static_cast<int>(n32 * uk / uG2), len0 = static_cast<int>(n32 * (uk + 1u) / uG2) - a0;
    const int a1 = static_cast<int>(n32 * (uG2 - 1u - uk) / uG2), len1 = static_cast<int>(n32 * (uG2 - uk) / uG2) - a1;
    const int len = len0 + len1;

    bool ok = true;
    TSB_PROF(0)

    // ---- (2) my slice -> shared memory, 16-byte piece by piece (4 pieces per node, consecutive lanes on consecutive
    // pieces: every warp load is 512 contiguous bytes); a piece is polled until both of its words carry the epoch of
    // the layer its node lies in
    {
      SpinGuard guard;
      const unsigned long long* src0 = fat[s0 + a0].w;
      const unsigned long long* src1 = fat[s0 + a1].w - 8 * len0;  // (indexed by the concatenated piece number)
      const int top = n_lay - 1;
      constexpr int PCS = 4 * LL_PPT;  // pieces per thread
      unsigned long long w0[PCS], w1[PCS];
      unsigned pending = 0;
#pragma unroll
      for (int j = 0; j < PCS; j++)
        if (t + j * T < 4 * len) pending |= 1u << j;
      // the epoch node i of my slice was stored with: that of the layer its position lies in (mostly the top one)
      const auto want_of = [&](int i) {
        const long long pos = s0 + (i < len0 ? a0 + i : a1 + (i - len0));
        int L = top;
        while (L > 0 && sm.lay_start[L] > pos) --L;
        return sm.lay_epoch[L];
      };
      while (pending) {
        // all loads of a sweep are issued back to back (a dependent re-poll per piece would serialise 8 L2 round trips)
#pragma unroll
        for (int j = 0; j < PCS; j++)
          if (pending & (1u << j)) {
            const int pc = t + j * T;
            ld_fat2((pc < 4 * len0 ? src0 : src1) + 2 * pc, w0[j], w1[j]);
          }
#pragma unroll
        for (int j = 0; j < PCS; j++)
          if (pending & (1u << j)) {
            const int pc = t + j * T, i = pc >> 2;
            const unsigned want = want_of(i);
            if (want == LL_TRUSTED || (static_cast<unsigned>(w0[j] >> 32) == want && static_cast<unsigned>(w1[j] >> 32) == want)) {
              *reinterpret_cast<uint2*>(&sm.parent[i][2 * (pc & 3)]) =
                  make_uint2(static_cast<uint32_t>(w0[j]), static_cast<uint32_t>(w1[j]));
              pending &= ~(1u << j);
            }
          }
        if (pending && guard.expired(&sy->abort)) {
          ok = false;
          break;
        }
      }
    }
    if (ll_bar_or(T, !ok)) {  // the slice is in shared memory
      exit_code = RND_EXIT_ABORT;
      break;
    }
    uint32_t cm[LL_PPT];
    int leaves = 0, mine = 0, mine0 = 0;
#pragma unroll
    for (int q = 0; q < LL_PPT; q++) {
      const int i = LL_PPT * t + q;
      cm[q] = 0;
      if (i < len) {
        const uint2 ax = *reinterpret_cast<const uint2*>(&sm.parent[i][6]);
        const unsigned long long aux = static_cast<unsigned long long>(ax.x) | static_cast<unsigned long long>(ax.y) << 32;
        cm[q] = static_cast<uint32_t>(aux >> 40) & 0xFFFFFu;
        leaves += static_cast<int>(aux >> 60) & 1;
        mine += __popc(cm[q]);
        if (i < len0) mine0 += __popc(cm[q]);
      }
    }
    TSB_PROF(1)
    // ---- (3) block scan: children | leaves << 20 | children of the bottom sub-slice << 32
    unsigned long long incl = static_cast<unsigned long long>(mine) | static_cast<unsigned long long>(leaves) << 20 |
                              static_cast<unsigned long long>(mine0) << 32;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) sm.warp_tot64[wid] = incl;
    ll_bar(T);
    unsigned long long woff = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < T / 32; i++) {
      if (i < wid) woff += sm.warp_tot64[i];
      tot += sm.warp_tot64[i];
    }
    const int my_children = static_cast<int>(tot & 0xFFFFF), my_leaves = static_cast<int>(tot >> 20) & 0xFFF;
    const int cnt0 = static_cast<int>(tot >> 32), cnt1 = my_children - cnt0;
    // ---- (4) publish {epoch, leaves, children} of my two sub-slices (slot s = sub-slice s, bottom to top)
    unsigned long long* const slots = sy->slot[epoch & 1u];
    if (t == 0) {
      st_relaxed_u64(&slots[k], static_cast<unsigned long long>(epoch) << 32 | static_cast<unsigned long long>(my_leaves) << 20 |
                                    static_cast<unsigned long long>(cnt0));
      st_relaxed_u64(&slots[G2 - 1 - k], static_cast<unsigned long long>(epoch) << 32 | static_cast<unsigned long long>(cnt1));
    }
    {
      uint16_t* it = sm.item + (static_cast<int>((woff + incl) & 0xFFFFF) - mine);
#pragma unroll
      for (int q = 0; q < LL_PPT; q++) {
        uint32_t m = cm[q];
        while (m) {
          const int s = __ffs(m) - 1;
          m &= m - 1;
          *it++ = static_cast<uint16_t>(((LL_PPT * t + q) << 5) | s);
        }
      }
    }
    ll_bar(T);  // items complete
    TSB_PROF(2)
    // ---- (5) my children (first window), built and evaluated while the other CTAs' counts are on their way
    auto build_window = [&](int c0, int cnt) {
      for (int c = t; c < cnt; c += T) {
        uint32_t ch[8];
        ll_build_child<N>(sm.parent, sm.item[c0 + c], ch);
        uint4* dst = reinterpret_cast<uint4*>(sm.stage[c]);
        dst[0] = make_uint4(ch[0], ch[1], ch[2], ch[3]);
        dst[1] = make_uint4(ch[4], ch[5], ch[6], ch[7]);
      }
    };
    build_window(0, min(LL_CAP, my_children));
    TSB_PROF(6)
    // ---- (6) all-to-all: everybody's {leaves, children}; my child offset and the round's totals
    unsigned long long before0 = 0, before1 = 0, all = 0;
    if (wid == 0) {
      ok = warp_gather_slots2(slots, G2, k, G2 - 1 - k, epoch, &sy->abort, before0, before1, all);
      if (lane == 0) {
        sm.red[0] = before0;
        sm.red[1] = before1;
        sm.red[2] = all;
      }
    }
    if (ll_bar_or(T, !ok)) {  // (also: the window is complete, red[] visible)
      exit_code = RND_EXIT_ABORT;
      break;
    }
    TSB_PROF(3)
    const long long off0 = static_cast<long long>(sm.red[0] & 0xFFFFFFFFull);
    const long long off1 = static_cast<long long>(sm.red[1] & 0xFFFFFFFFull);
    all = sm.red[2];
    const long long round_children = static_cast<long long>(all & 0xFFFFFFFFull);
    const long long round_leaves = static_cast<long long>(all >> 32);

    // ---- (7) my children, in place, tagged with this round's epoch (every slice of the chunk has been read: all
    // G slots carried this epoch)
    const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
    for (int c0 = 0; c0 < my_children; c0 += LL_CAP) {
      const int cnt = min(LL_CAP, my_children - c0);
      if (c0 > 0) {
        ll_bar(T);  // the previous window has been copied out
        build_window(c0, cnt);
        ll_bar(T);
      }
      // child c of my share goes to position s0 + off0 + c (bottom sub-slice) or s0 + off1 + (c - cnt0) (top one)
      unsigned long long* const dst0 = fat[s0 + off0 + c0].w;
      unsigned long long* const dst1 = fat[s0 + off1 + c0 - cnt0].w;
      const int npc = 4 * cnt;  // 16-byte pieces, consecutive lanes on consecutive pieces, four in flight per thread
      for (int pc = t; pc < npc; pc += 4 * T) {
        uint2 d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int x = min(pc + u * T, npc - 1);
          d[u] = *reinterpret_cast<const uint2*>(&sm.stage[x >> 2][2 * (x & 3)]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int x = pc + u * T;
          if (x < npc) st_fat2((c0 + (x >> 2) < cnt0 ? dst0 : dst1) + 2 * x, d[u].x | tag, d[u].y | tag);
        }
      }
    }
    TSB_PROF(4)
    // ---- (8) the pool's layers after the round: every layer that starts inside the chunk is consumed, the round's
    // children form the new top layer (same computation in every CTA)
    {
      int nl = n_lay;
      while (nl > 0 && sm.lay_start[nl - 1] >= s0) --nl;
      ll_bar(T);  // (everybody has read the old entries, the window buffers and sm.red)
      if (round_children > 0) {
        if (t == 0) {
          sm.lay_start[nl] = s0;
          sm.lay_epoch[nl] = epoch;
        }
        ++nl;
      }
      n_lay = nl;
      ll_bar(T);
    }
    TSB_PROF(5)
    // ---- (9) the pool after the round
    size = s0 + round_children;
    ++rounds;
    tot_parents += static_cast<unsigned long long>(n);
    tot_children += static_cast<unsigned long long>(round_children);
    tot_solutions += static_cast<unsigned long long>(round_leaves);
  }
  if (k == 0 && t == 0) {
    RoundsState* st = prm.state;
    st->size = size;
    st->epoch = epoch;
    st->rounds = rounds;
    st->parents = tot_parents;
    st->children = tot_children;
    st->solutions = tot_solutions;
    st->exit_code = exit_code;
    if (prm.prof)
      for (int i = 0; i < 8; i++) st->prof[i] = prof[i];
  }
#undef TSB_PROF
}

}  // namespace tsb
