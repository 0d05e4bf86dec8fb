// nq_rounds.cuh — many offload rounds of the N-Queens search per launch (the `--M 50000` regime).
//
// One round of the reference's step 2 (nqueens_gpu_chpl.chpl:197-215) is: popBackBulk(m, M) = the newest
// n = min(size, M) nodes of the pool (nothing if size < m, lib/commons/Pool.chpl:50-59), evaluate_gpu on them
// (:97-123), generate_children (:126-149) pushing the surviving children back, in order.  Round i+1 pops what
// round i pushed, so the rounds are inherently sequential; at the reference's default --M 50000 a round moves
// ~2 MB (0.3 us of HBM time) and the two-kernel pipeline of nq_expand.cuh is bound by launch + host latency
// (15 us per round, 160 348 rounds for N = 17).  This kernel keeps the whole loop on the device:
//
//   * cooperative launch, one CTA per SM, all co-resident; the pool is ONE contiguous stack [0, size) of an
//     arena in HBM.  Every CTA tracks the (tiny) pool state redundantly — it is a deterministic function of the
//     per-round totals, which every CTA learns anyway — so there is no shared state to broadcast;
//   * a round: CTA k takes slice k of the chunk (n/G parents, <= RND_SLICE), loads it into shared memory with
//     coalesced 16-byte loads, evaluates it (the attacked-values mask of nq_kernel.cuh), scans its child counts
//     and publishes {epoch, leaves, children} in ITS 64-bit slot; all CTAs poll all G slots (one all-to-all
//     flag exchange = one L2 round trip, no atomics) and derive their child offset and the round's totals; the
//     children are built in shared memory exactly as in nq_expand_build and stored IN PLACE — the chunk is the
//     top of the stack and every slice is known to be in shared memory once all G slots are visible, so the
//     children of the round overwrite the chunk: the pool stays one contiguous stack, byte-identical to the
//     reference's pool after every round, with no holes to compact;
//   * a second all-to-all flag exchange ("children of round r are in L2") orders round r+1's loads after
//     round r's stores.  Two flag exchanges per round, ~2 us per round instead of 15.
//
// A spin loop that waits longer than ~2 s raises a global abort flag and every CTA leaves (exit code ABORT):
// a logic error must never hang the GPU.
#pragma once
#include "expand_common.cuh"
#include "nq_expand.cuh"
#include "nq_kernel.cuh"

namespace tsb {

constexpr int RND_THREADS = 256;                   // (default; the kernel is templated on the CTA size)
constexpr int RND_PPT = 2;                         // parents per thread
constexpr int RND_SLICE = RND_THREADS * RND_PPT;   // parents per CTA per round at the default CTA size
constexpr int RND_MAX_CTAS = 256;

enum { RND_EXIT_DONE = 0, RND_EXIT_PAUSE = 1, RND_EXIT_SPACE = 2, RND_EXIT_ABORT = 3, RND_EXIT_RELAUNCH = 4 };

// device-resident synchronisation area (zeroed once; epochs increase monotonically across launches)
struct RoundsSync {
  unsigned long long slot[RND_MAX_CTAS];  // epoch << 32 | leaves << 20 | children of the CTA's slice
  unsigned done[RND_MAX_CTAS];            // epoch of the last round whose children this CTA has stored
  unsigned abort;
};
// in / out record of a launch (pinned + mapped host memory)
struct RoundsState {
  long long size;                 // nodes in the pool: arena positions [0, size)
  unsigned epoch;                 // last epoch used
  int exit_code;
  unsigned long long rounds, parents, children, solutions;  // of this launch
  long long prof[8];  // (prm.prof) cycles CTA 0 spent per phase: wait done, load, evaluate+scan, gather, build+store, release
};
struct RoundsParams {
  uint8_t* arena;
  long long cap;    // nodes the arena holds (plus slack for 16-byte over-reads)
  long long size0;  // nodes in the pool at launch: arena positions [0, size0)
  unsigned epoch0;  // last epoch used by the previous launch
  int m, M;
  long long max_rounds;
  int prof;  // accumulate per-phase cycle counts of CTA 0 into state->prof (env TSB200_ROUNDS_PROF)
  unsigned long long* aux;  // one side word per arena position (nq_aux_pack)
  long long aux_valid;      // positions [0, aux_valid) already hold the aux word of their node
  RoundsSync* sync;
  RoundsState* state;
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Per-node side word ("aux"), one per arena position, kept by this kernel next to the pool:
//   bits  0..19  ld: values attacked on the node's next row along the rising diagonals  {board[i] + (depth - i)}
//   bits 20..39  rd: ... along the falling diagonals                                     {board[i] - (depth - i)}
//   bits 40..59  the node's child mask: slot k set <=> k >= depth and board[k] is not attacked (evaluate_gpu's
//                label for slot k, nqueens_gpu_chpl.chpl:97-123)
//   bit  60      leaf (depth == N)
// A child's word follows from its parent's in O(1) — ld' = ((ld | 1 << v) << 1) mod 2^N, rd' = (rd | 1 << v) >> 1
// for the placed value v — so a node is evaluated ONCE, when it is created, without the O(depth) pass over its
// board; a round then only pop-counts the masks of its chunk.  Nodes that did not come out of this kernel
// (host pushes, other kernels) get their word in the kernel's prologue, from the board.
__device__ __forceinline__ unsigned long long nq_aux_pack(uint32_t ld, uint32_t rd, uint32_t cm, bool leaf) {
  return static_cast<unsigned long long>(ld) | static_cast<unsigned long long>(rd) << 20 |
         static_cast<unsigned long long>(cm) << 40 | static_cast<unsigned long long>(leaf ? 1u : 0u) << 60;
}
template <int N>
__device__ __forceinline__ unsigned long long nq_aux_of_node(const uint8_t* node) {  // the reference predicate, slot by slot
  const int d = node[0];
  uint32_t ld = 0, rd = 0;
  for (int i = 0; i < d && i < N; i++) {
    const int b = node[1 + i], s = d - i;
    if (b + s < N) ld |= 1u << (b + s);
    if (b - s >= 0) rd |= 1u << (b - s);
  }
  const uint32_t U = ld | rd;
  uint32_t cm = 0;
  for (int k = d; k < N; k++)
    if (!((U >> (node[1 + k] & 31)) & 1u)) cm |= 1u << k;
  return nq_aux_pack(ld, rd, cm, d == N);
}

constexpr int RND_CAP2 = 2048;  // children per window of the staging image
template <int T>
struct RoundsSmem {
  alignas(128) uint8_t raw[T * RND_PPT * NQ_REC + 48];   // the slice, at the 16-byte phase of its arena address
  alignas(128) uint8_t stage[RND_CAP2 * NQ_REC + 32];    // children of one window (any phase: realigned on the way out)
  alignas(16) unsigned long long aux_in[T * RND_PPT];     // aux words of the slice
  alignas(16) unsigned long long aux_out[RND_CAP2];       // aux words of the window's children
  alignas(16) uint16_t item[T * RND_PPT * 20];            // (record << 5) | slot, in child order
  int warp_tot[T / 32];
  unsigned long long red[2];
};

// Flag exchanges.  Every CTA needs every other CTA's flag, so the polling is done by ONE warp per CTA with
// coalesced 16-byte loads (10 cache lines per sweep of the slots, 5 of the done flags): 148 threads per CTA each
// spinning on its own flag put ~22 000 loads per sweep on a handful of L2 lines and made a round 9 us.
struct SpinGuard {  // watchdog of a spin loop: ~2 s, or another CTA's abort
  unsigned spins = 0;
  long long t0 = 0;
  __device__ __forceinline__ bool expired(unsigned* abort_flag) {
    if ((++spins & 0xFFu) != 0) return false;
    if (*reinterpret_cast<volatile unsigned*>(abort_flag)) return true;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    if (now - t0 > 4000000000LL) {
      *reinterpret_cast<volatile unsigned*>(abort_flag) = 1u;
      return true;
    }
    return false;
  }
};
// warp 0: until all G done flags equal `want`; false = abort
__device__ __forceinline__ bool warp_wait_done(const unsigned* done, int G, unsigned want, unsigned* abort_flag) {
  const int lane = threadIdx.x & 31;
  SpinGuard guard;
  for (;;) {
    bool ok = true;
    for (int i = 4 * lane; i < G; i += 128) {
      unsigned v0, v1, v2, v3;
      asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                   : "l"(done + i)
                   : "memory");
      ok &= v0 == want && (i + 1 >= G || v1 == want) && (i + 2 >= G || v2 == want) && (i + 3 >= G || v3 == want);
    }
    if (__all_sync(0xFFFFFFFFu, ok)) break;
    if (__any_sync(0xFFFFFFFFu, guard.expired(abort_flag))) return false;
  }
  // No acquire fence: a gpu-scope fence costs 1 000-2 000 cycles here (MEMBAR.SC + CCTL.IVALL, tools/flag_exchange.py),
  // and all it would add is an L1 invalidation — the slice is then read with ld.global.cg (L2 only), by loads
  // issued after this poll has returned (bar.sync in between), and the writers released at gpu scope.
  return true;
}
// warp 0: until all G slots carry `epoch`; sums of {leaves << 32 | children} over all slots and over the slots
// before k (valid in every lane); false = abort
__device__ __forceinline__ bool warp_gather_slots(const unsigned long long* slot, int G, int k, unsigned epoch,
                                                  unsigned* abort_flag, unsigned long long& before,
                                                  unsigned long long& all) {
  const int lane = threadIdx.x & 31;
  SpinGuard guard;
  for (;;) {
    bool ok = true;
    before = 0;
    all = 0;
    for (int i = 2 * lane; i < G; i += 64) {
      unsigned long long v0, v1;
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(slot + i) : "memory");
      const bool has1 = i + 1 < G;
      ok &= static_cast<unsigned>(v0 >> 32) == epoch && (!has1 || static_cast<unsigned>(v1 >> 32) == epoch);
      const unsigned long long p0 = (v0 & 0xFFFFFull) | ((v0 >> 20) & 0xFFFull) << 32;
      const unsigned long long p1 = has1 ? (v1 & 0xFFFFFull) | ((v1 >> 20) & 0xFFFull) << 32 : 0ull;
      all += p0 + p1;
      if (i < k) before += p0;
      if (i + 1 < k) before += p1;
    }
    if (__all_sync(0xFFFFFFFFu, ok)) break;
    if (__any_sync(0xFFFFFFFFu, guard.expired(abort_flag))) return false;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    before += __shfl_xor_sync(0xFFFFFFFFu, before, o);
    all += __shfl_xor_sync(0xFFFFFFFFu, all, o);
  }
  return true;
}

// one parent at an arbitrary byte address of shared memory -> parent-aligned words P[0..5]
__device__ __forceinline__ void nq_load_parent_words(const uint8_t* src, uint32_t (&P)[6]) {
  const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(src)) & 3u, a8 = mis * 8u;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src - mis);
  const uint32_t s0 = sw[0], s1 = sw[1], s2 = sw[2], s3 = sw[3], s4 = sw[4], s5 = sw[5];
  P[0] = shf_r_wrap(s0, s1, a8);
  P[1] = shf_r_wrap(s1, s2, a8);
  P[2] = shf_r_wrap(s2, s3, a8);
  P[3] = shf_r_wrap(s3, s4, a8);
  P[4] = shf_r_wrap(s4, s5, a8);
  P[5] = shf_r_wrap(s5, 0u, a8);
}

// child `item` of the slice -> bytes [B, B + 21) of the image (B = 21 c for child c of the window: lane & 3 ==
// B & 3, every word stored whole, see nq_build_child_warp) and its aux word, from the parent's
template <int N>
__device__ __forceinline__ unsigned long long nq_build_child_warp_aux(const uint8_t* recs, const unsigned long long* aux_in,
                                                                      int item, uint8_t* image, int B, bool active) {
  const int b = threadIdx.x & 3;
  const uint32_t b8 = b * 8;
  uint32_t W0 = 0, W1 = 0, W2 = 0, W3 = 0, W4 = 0, W5 = 0;
  unsigned long long child_aux = 0;
  if (active) {
    const int r = item >> 5, k = item & 31;
    const uint8_t* src = recs + r * NQ_REC;
    uint32_t P[6];
    nq_load_parent_words(src, P);
    P[5] &= 0xFFu;
    const uint32_t depth = P[0] & 0xFFu;
    const uint32_t p1 = 1u + depth, p2 = 1u + static_cast<uint32_t>(k);
    const uint32_t v = src[p2];  // the queen placed on row `depth`
    const uint32_t D = static_cast<uint32_t>(src[p1]) ^ v;
    const uint32_t x1 = D << ((p1 & 3u) * 8u), x2 = D << ((p2 & 3u) * 8u);
    const uint32_t w1 = p1 >> 2, w2 = p2 >> 2;
#pragma unroll
    for (uint32_t j = 0; j < 6; j++) P[j] ^= (j == w1 ? x1 : 0u) ^ (j == w2 ? x2 : 0u);
    P[0] += 1u;
    // the child's aux word from the parent's
    const unsigned long long w = aux_in[r];
    const uint32_t ld = static_cast<uint32_t>(w) & 0xFFFFFu, rd = static_cast<uint32_t>(w >> 20) & 0xFFFFFu;
    const uint32_t bit = 1u << (v & 31u);
    const uint32_t ld2 = ((ld | bit) << 1) & ((1u << N) - 1u), rd2 = (rd | bit) >> 1;
    NqParent<N, 0, 0> cp;
    cp.init(P);  // depth + 1, shift amounts = the child's board
    cp.U = ld2 | rd2;
    const uint32_t cm = nq_child_mask<N, 0>(cp);  // slots >= depth + 1 whose value is safe (none for a leaf)
    child_aux = nq_aux_pack(ld2, rd2, cm, depth + 1u == static_cast<uint32_t>(N));
    W0 = shf_l_wrap(0u, P[0], b8);
    W1 = shf_l_wrap(P[0], P[1], b8);
    W2 = shf_l_wrap(P[1], P[2], b8);
    W3 = shf_l_wrap(P[2], P[3], b8);
    W4 = shf_l_wrap(P[3], P[4], b8);
    W5 = shf_l_wrap(P[4], P[5], b8);
  }
  const uint32_t nb = __shfl_down_sync(0xFFFFFFFFu, W0, 1);  // the right-hand neighbour's first word (0 if none)
  if (active) {
    uint32_t* dw = reinterpret_cast<uint32_t*>(image + (B - b));
    if (b == 0) dw[0] = W0;
    dw[1] = W1;
    dw[2] = W2;
    dw[3] = W3;
    dw[4] = W4;
    dw[5] = b == 3 ? W5 : (W5 | nb);
  }
  return child_aux;
}

template <int N, int T>
__global__ void __launch_bounds__(T, 1) nq_rounds_kernel(const __grid_constant__ RoundsParams prm) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  RoundsSmem<T>& sm = *reinterpret_cast<RoundsSmem<T>*>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int k = blockIdx.x, G = gridDim.x;
  RoundsSync* const sy = prm.sync;
  uint8_t* const arena = prm.arena;
  unsigned long long* const aux = prm.aux;

  long long size = prm.size0;
  unsigned epoch = prm.epoch0;
  unsigned long long rounds = 0, tot_parents = 0, tot_children = 0, tot_solutions = 0;
  int exit_code = RND_EXIT_PAUSE;
  long long prof[6] = {0, 0, 0, 0, 0, 0}, tp = 0;
  const bool prof_on = prm.prof != 0 && k == 0 && t == 0;
#define TSB_PROF(i)                  \
  if (prof_on) {                     \
    const long long now = clock64(); \
    prof[i] += now - tp;             \
    tp = now;                        \
  }

  // ---- prologue: aux words of the nodes that did not come out of this kernel, then "round 0 is stored"
  for (long long pos = prm.aux_valid + static_cast<long long>(k) * T + t; pos < size; pos += static_cast<long long>(G) * T)
    __stcg(aux + pos, nq_aux_of_node<N>(arena + pos * NQ_REC));
  ++epoch;
  __syncthreads();
  if (t == 0) st_release_u32(&sy->done[k], epoch);

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
    const long long s0 = size - n;  // arena position of the chunk's first parent = of the round's first child
    if (s0 + n * N > prm.cap) {     // worst case: every slot of every parent survives
      exit_code = RND_EXIT_SPACE;
      break;
    }
    ++epoch;
    if (prof_on) tp = clock64();
    const int a = static_cast<int>(n * k / G), b = static_cast<int>(n * (k + 1) / G);  // my slice of the chunk
    const int len = b - a;

    // ---- (1) the children (and aux words) of the previous round are in L2: all CTAs have stored theirs
    bool ok = true;
    if (wid == 0) ok = warp_wait_done(sy->done, G, epoch - 1u, &sy->abort);
    if (__syncthreads_or(!ok)) {
      exit_code = RND_EXIT_ABORT;
      break;
    }
    TSB_PROF(0)

    // ---- (2) slice + its aux words -> shared memory (L2 loads: other SMs wrote these bytes)
    const uint8_t* gsrc = arena + (s0 + a) * NQ_REC;
    const uint32_t ph_in = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(gsrc)) & 15u;
    {
      const uint4* g4 = reinterpret_cast<const uint4*>(gsrc - ph_in);
      uint4* s4 = reinterpret_cast<uint4*>(sm.raw);
      const int n16 = (static_cast<int>(ph_in) + len * NQ_REC + 15) >> 4;
      for (int i = t; i < n16; i += T) s4[i] = __ldcg(g4 + i);
      const unsigned long long* ga = aux + (s0 + a);
      for (int i = t; i < len; i += T) sm.aux_in[i] = __ldcg(ga + i);
    }
    __syncthreads();
    TSB_PROF(1)
    const uint8_t* const recs = sm.raw + ph_in;  // record i of the slice at recs + 21 i

    // ---- (3) my parents' child masks (evaluated when the nodes were created), leaves
    uint32_t cm[RND_PPT];
    int leaves = 0, mine = 0;
#pragma unroll
    for (int q = 0; q < RND_PPT; q++) {
      const int i = RND_PPT * t + q;
      const unsigned long long w = i < len ? sm.aux_in[i] : 0ull;
      cm[q] = static_cast<uint32_t>(w >> 40) & 0xFFFFFu;
      leaves += static_cast<int>(w >> 60) & 1;
      mine += __popc(cm[q]);
    }
    // ---- (4) block scan of the child counts (leaves ride in the upper bits), items of the slice
    int incl = mine | (leaves << 20);  // children of a slice <= 1024 * 20 < 2^20, leaves <= 1024
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) sm.warp_tot[wid] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < T / 32; i++) {
      if (i < wid) woff += sm.warp_tot[i];
      tot += sm.warp_tot[i];
    }
    const int my_children = tot & 0xFFFFF, my_leaves = tot >> 20;
    // ---- (5) publish {epoch, leaves, children} of my slice
    if (t == 0)
      st_relaxed_u64(&sy->slot[k], static_cast<unsigned long long>(epoch) << 32 |
                                       static_cast<unsigned long long>(my_leaves) << 20 |
                                       static_cast<unsigned long long>(my_children));
    {
      uint16_t* it = sm.item + (((woff + incl) & 0xFFFFF) - mine);
#pragma unroll
      for (int q = 0; q < RND_PPT; q++) {
        uint32_t m = cm[q];
        while (m) {
          const int s = __ffs(m) - 1;
          m &= m - 1;
          *it++ = static_cast<uint16_t>(((RND_PPT * t + q) << 5) | s);
        }
      }
    }
    __syncthreads();  // items complete
    TSB_PROF(2)
    // ---- (6) first window of my children, built (and evaluated) while the other CTAs' counts are on their way
    auto build_window = [&](int c0, int cnt) {
      for (int cb = 0; cb < cnt; cb += T) {
        const int c = cb + t;
        const bool active = c < cnt;
        const unsigned long long ca = nq_build_child_warp_aux<N>(recs, sm.aux_in, active ? sm.item[c0 + c] : 0, sm.stage,
                                                                 c * NQ_REC, active);
        if (active) sm.aux_out[c] = ca;
      }
    };
    build_window(0, min(RND_CAP2, my_children));
    // ---- (7) all-to-all: everybody's {leaves, children}; my child offset and the round's totals
    unsigned long long before = 0, all = 0;  // packed leaves << 32 | children sums
    if (wid == 0) {
      ok = warp_gather_slots(sy->slot, G, k, epoch, &sy->abort, before, all);
      if (lane == 0) {
        sm.red[0] = before;
        sm.red[1] = all;
      }
    }
    if (__syncthreads_or(!ok)) {  // (also: the window's image and aux words are complete, red[] visible)
      exit_code = RND_EXIT_ABORT;
      break;
    }
    TSB_PROF(3)
    before = sm.red[0];
    all = sm.red[1];
    const long long child_off = static_cast<long long>(before & 0xFFFFFFFFull);
    const long long round_children = static_cast<long long>(all & 0xFFFFFFFFull);
    const long long round_leaves = static_cast<long long>(all >> 32);

    // ---- (8) my children, in place: arena positions s0 + child_off ...  (every slice of the chunk is in shared
    // memory by now: all G slots carried this epoch)
    for (int c0 = 0; c0 < my_children; c0 += RND_CAP2) {
      const int cnt = min(RND_CAP2, my_children - c0);
      if (c0 > 0) {
        __syncthreads();  // the previous window has been copied out
        build_window(c0, cnt);
        __syncthreads();
      }
      const long long pos = s0 + child_off + c0;
      copy_image_to_global(arena + pos * NQ_REC, sm.stage, cnt * NQ_REC, t, T);
      for (int c = t; c < cnt; c += T) __stcg(aux + pos + c, sm.aux_out[c]);
    }
    TSB_PROF(4)
    // ---- (9) my children are stored: release (the barrier orders every thread's stores before thread 0's
    // release, which is cumulative — the pattern of a cooperative-groups grid barrier)
    __syncthreads();
    if (t == 0) st_release_u32(&sy->done[k], epoch);
    TSB_PROF(5)
    // ---- (10) the pool after the round
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
      for (int i = 0; i < 6; i++) st->prof[i] = prof[i];
  }
#undef TSB_PROF
}

// ---- diagnostics: the bare flag-exchange skeleton of a round (no evaluation, no children), to measure the floor
// the exchanges put under a round.  variant bits: 1 = no release fence (plain store of the done flag); 2 = no
// acquire fence; 4 = every thread stores 16 bytes to global before the release (a round's children); 8 = polls
// are weak L2 loads (ld.global.cg) instead of ld.relaxed.gpu; 16 = only ONE exchange per round (the slots);
// 32 = one exchange through per-reader inboxes (every writer stores its flag into every reader's own row)
template <bool CG>
__device__ __forceinline__ void bench_ld4(const unsigned* p, unsigned& v0, unsigned& v1, unsigned& v2, unsigned& v3) {
  if constexpr (CG)
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p) : "memory");
  else
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p) : "memory");
}
// warp 0: all G 32-bit flags equal `want`
template <bool CG>
__device__ __forceinline__ bool bench_wait32(const unsigned* f, int G, unsigned want, unsigned* abort_flag) {
  const int lane = threadIdx.x & 31;
  SpinGuard guard;
  for (;;) {
    bool ok = true;
    for (int i = 4 * lane; i < G; i += 128) {
      unsigned v0, v1, v2, v3;
      bench_ld4<CG>(f + i, v0, v1, v2, v3);
      ok &= v0 == want && (i + 1 >= G || v1 == want) && (i + 2 >= G || v2 == want) && (i + 3 >= G || v3 == want);
    }
    if (__all_sync(0xFFFFFFFFu, ok)) return true;
    if (__any_sync(0xFFFFFFFFu, guard.expired(abort_flag))) return false;
  }
}
template <bool CG>
__device__ __forceinline__ void rounds_sync_bench_body(RoundsSync* sy, unsigned epoch0, int rounds, int variant,
                                                       uint4* scratch, long long* out_cycles) {
  const int t = threadIdx.x, wid = t >> 5, k = blockIdx.x, G = gridDim.x;
  unsigned* const slot32 = reinterpret_cast<unsigned*>(sy->slot);  // 32-bit slots: 592 B = 5 lines per sweep
  unsigned epoch = epoch0;
  const long long c0 = clock64();
  for (int r = 0; r < rounds; r++) {
    ++epoch;
    bool ok = true;
    if (!(variant & 16)) {
      if (r > 0 && wid == 0) {
        ok = bench_wait32<CG>(sy->done, G, epoch - 1u, &sy->abort);
        if (!(variant & 2)) __threadfence();
      }
      if (__syncthreads_or(!ok)) break;
    }
    if (variant & 32) {  // per-reader inboxes: writer k stores its flag into row j of every reader j; a reader polls
                         // only its own row (no line is polled by more than one CTA)
      unsigned* const inbox = reinterpret_cast<unsigned*>(scratch) + (r & 1) * 256 * 256;
      for (int j = t; j < G; j += RND_THREADS)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&inbox[j * 256 + k]), "r"(epoch) : "memory");
      if (wid == 0) ok = bench_wait32<CG>(inbox + k * 256, G, epoch, &sy->abort);
      if (__syncthreads_or(!ok)) break;
      continue;
    }
    unsigned* const sl = slot32 + 256 * (r & 1);  // (two slot arrays, by round parity: a single exchange per round
                                                  // lets a fast CTA publish round r+1 before a slow one has read r)
    if (t == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&sl[k]), "r"(epoch) : "memory");
    if (wid == 0) ok = bench_wait32<CG>(sl, G, epoch, &sy->abort);
    if (__syncthreads_or(!ok)) break;
    if (variant & 4) __stcg(scratch + (static_cast<long long>(k) * RND_THREADS + t), make_uint4(epoch, t, k, r));
    if (!(variant & 16)) {
      __syncthreads();
      if (t == 0) {
        if (variant & 1)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&sy->done[k]), "r"(epoch) : "memory");
        else
          st_release_u32(&sy->done[k], epoch);
      }
    }
  }
  if (k == 0 && t == 0) *out_cycles = clock64() - c0;
}
__global__ void __launch_bounds__(RND_THREADS, 1) rounds_sync_bench_kernel(RoundsSync* sy, unsigned epoch0, int rounds,
                                                                          int variant, uint4* scratch,
                                                                          long long* out_cycles) {
  if (variant & 8)
    rounds_sync_bench_body<true>(sy, epoch0, rounds, variant, scratch, out_cycles);
  else
    rounds_sync_bench_body<false>(sy, epoch0, rounds, variant, scratch, out_cycles);
}

}  // namespace tsb
