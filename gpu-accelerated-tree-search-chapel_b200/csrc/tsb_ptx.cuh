// tsb_ptx.cuh — sm_100a PTX helpers: mbarrier, 1-D TMA bulk copies (cp.async.bulk), proxy
// fences, and the persistent tile pipeline shared by all evaluator kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tsb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// global -> shared, 1-D bulk copy by the TMA engine (SASS: UBLKCP); size and both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// same with an L2 evict-first policy: node chunks are streamed exactly once
__device__ __forceinline__ void bulk_g2s_stream(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                                uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// shared -> global bulk store
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// PTX shifts with CLAMP semantics (amount >= 32 gives 0), which C++ << / >> do not guarantee
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t n) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t shr_clamp(uint32_t x, uint32_t n) {
  uint32_t r;
  asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
  return r;
}
// funnel shifts in WRAP mode: only the low 5 bits of the amount register are used
__device__ __forceinline__ uint32_t shf_l_wrap(uint32_t lo, uint32_t hi, uint32_t n) {
  uint32_t r;
  asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t shf_r_wrap(uint32_t lo, uint32_t hi, uint32_t n) {
  uint32_t r;
  asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(n));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Persistent tile pipeline.  A kernel processes `count` records in tiles of TILE records;
// a full tile is IN_BYTES in / OUT_BYTES out (both multiples of 16).  CTA b handles tiles
// b, b+grid, ...  STAGES input and output buffers live in shared memory; one elected thread
// drives the TMA engine:  g2s(tile i+STAGES) is issued as soon as tile i's input buffer has
// been consumed, s2g(tile i) as soon as its output buffer has been produced.  One
// __syncthreads per tile.  The last partial tile (count % TILE records) is done by the CTA
// that owns it with plain word/byte copies (bulk copies need 16-B granularity).
//
// Functor: compute(const uint8_t* in_tile, uint8_t* out_tile, int records_in_tile, long long tile_index)
// ---------------------------------------------------------------------------------------------
template <int STAGES, int IN_BYTES, int OUT_BYTES>
struct TileSmem {
  alignas(128) uint8_t in[STAGES][IN_BYTES];
  alignas(128) uint8_t out[STAGES][OUT_BYTES];
  alignas(8) uint64_t full[STAGES];
};

template <int STAGES, int TILE, int IN_REC, int OUT_REC, typename Smem, typename F>
__device__ __forceinline__ void run_tile_pipeline(Smem& sm, const uint8_t* __restrict__ in_g,
                                                  uint8_t* __restrict__ out_g, long long count, F&& compute) {
  constexpr uint32_t IN_BYTES = TILE * IN_REC;
  constexpr uint32_t OUT_BYTES = TILE * OUT_REC;
  static_assert(IN_BYTES % 16 == 0 && OUT_BYTES % 16 == 0, "tile sizes must be multiples of 16 B");
  const long long full_tiles = count / TILE;
  const int rem = static_cast<int>(count - full_tiles * TILE);
  const int tid = threadIdx.x;
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long my_tiles = first < full_tiles ? (full_tiles - first + stride - 1) / stride : 0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  uint64_t pol = 0;
  if (tid == 0) {
    pol = policy_evict_first();
    for (int s = 0; s < STAGES && s < my_tiles; s++) {
      const long long tile = first + s * stride;
      mbar_arrive_expect_tx(&sm.full[s], IN_BYTES);
      bulk_g2s_stream(sm.in[s], in_g + tile * IN_BYTES, IN_BYTES, &sm.full[s], pol);
    }
  }

  for (long long it = 0; it < my_tiles; it++) {
    const int s = static_cast<int>(it % STAGES);
    const uint32_t parity = static_cast<uint32_t>((it / STAGES) & 1);
    const long long tile = first + it * stride;
    mbar_wait(&sm.full[s], parity);
    compute(sm.in[s], sm.out[s], TILE, tile);
    fence_async_smem();
    // the store issued from out[(s+1)%STAGES] STAGES-1 iterations ago must have drained its
    // shared-memory reads before anyone writes that buffer in the next iteration
    if constexpr (STAGES >= 2) {
      if (tid == 0) bulk_wait_read<(STAGES >= 2 ? STAGES - 2 : 0)>();
    }
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(out_g + tile * OUT_BYTES, sm.out[s], OUT_BYTES);
      bulk_commit();
      // single-buffered: nobody can start the next tile before its load is issued below, so
      // draining the store's shared-memory reads here protects the one output buffer
      if constexpr (STAGES == 1) bulk_wait_read<0>();
      const long long nxt = it + STAGES;
      if (nxt < my_tiles) {
        mbar_arrive_expect_tx(&sm.full[s], IN_BYTES);
        bulk_g2s_stream(sm.in[s], in_g + (first + nxt * stride) * IN_BYTES, IN_BYTES, &sm.full[s], pol);
      }
    }
  }
  if (tid == 0) bulk_wait_all();

  // partial last tile: owned by the CTA that would own tile index `full_tiles`
  if (rem > 0 && (full_tiles % stride) == first) {
    __syncthreads();  // all bulk stores of this CTA have drained (thread 0 waited above)
    const uint8_t* src = in_g + full_tiles * IN_BYTES;
    uint8_t* dst = out_g + full_tiles * OUT_BYTES;
    uint8_t* sin = sm.in[0];
    uint8_t* sout = sm.out[0];
    const int in_b = rem * IN_REC, out_b = rem * OUT_REC;
    for (int i = tid; i < static_cast<int>(IN_BYTES); i += blockDim.x) sin[i] = i < in_b ? src[i] : 0;
    __syncthreads();
    compute(sin, sout, rem, full_tiles);
    __syncthreads();
    for (int i = tid; i < out_b; i += blockDim.x) dst[i] = sout[i];
  }
}

// ---------------------------------------------------------------------------------------------
// In-place variant for kernels whose output record is not larger than the input record: a ring of
// three tile buffers, each first the TMA destination of an input tile and then — after the
// functor has pulled its inputs into registers and synchronised — the TMA source of that tile's
// output.  At the end of iteration i the store of tile i is issued from buffer i%3 and, once the
// store of tile i-1 has drained its shared-memory reads (it had a whole tile time to do so), the
// load of tile i+2 goes into buffer (i-1)%3: loads stay two tiles ahead at the shared-memory
// cost of three input tiles and no output tiles.
// Functor: compute(uint8_t* tile /* in and out */, int records_in_tile, long long tile_index);
// it MUST __syncthreads() between its last read of the input and its first write of the output.
// ---------------------------------------------------------------------------------------------
template <int IN_BYTES>
struct RingSmem {
  alignas(128) uint8_t buf[3][IN_BYTES];
  alignas(8) uint64_t full[3];
};

template <int TILE, int IN_REC, int OUT_REC, typename Smem, typename F>
__device__ __forceinline__ void run_tile_ring_inplace(Smem& sm, const uint8_t* __restrict__ in_g,
                                                      uint8_t* __restrict__ out_g, long long count, F&& compute) {
  constexpr uint32_t IN_BYTES = TILE * IN_REC;
  constexpr uint32_t OUT_BYTES = TILE * OUT_REC;
  static_assert(IN_BYTES % 16 == 0 && OUT_BYTES % 16 == 0 && OUT_BYTES <= IN_BYTES, "tile sizes");
  const long long full_tiles = count / TILE;
  const int rem = static_cast<int>(count - full_tiles * TILE);
  const int tid = threadIdx.x;
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long my_tiles = first < full_tiles ? (full_tiles - first + stride - 1) / stride : 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 3; s++) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint64_t pol = 0;
  if (tid == 0) {
    pol = policy_evict_first();
    for (int s = 0; s < 2 && s < my_tiles; s++) {
      mbar_arrive_expect_tx(&sm.full[s], IN_BYTES);
      bulk_g2s_stream(sm.buf[s], in_g + (first + s * stride) * IN_BYTES, IN_BYTES, &sm.full[s], pol);
    }
  }
  for (long long it = 0; it < my_tiles; it++) {
    const int s = static_cast<int>(it % 3);
    const long long tile = first + it * stride;
    mbar_wait(&sm.full[s], static_cast<uint32_t>((it / 3) & 1));
    compute(sm.buf[s], TILE, tile);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(out_g + tile * OUT_BYTES, sm.buf[s], OUT_BYTES);
      bulk_commit();
      const long long nxt = it + 2;
      if (nxt < my_tiles) {
        bulk_wait_read<1>();  // the store of tile it-1 (buffer nxt % 3) has drained
        const int ns = static_cast<int>(nxt % 3);
        mbar_arrive_expect_tx(&sm.full[ns], IN_BYTES);
        bulk_g2s_stream(sm.buf[ns], in_g + (first + nxt * stride) * IN_BYTES, IN_BYTES, &sm.full[ns], pol);
      }
    }
  }
  if (tid == 0) bulk_wait_all();
  // partial last tile: owned by the CTA that would own tile index `full_tiles`
  if (rem > 0 && (full_tiles % stride) == first) {
    __syncthreads();
    const uint8_t* src = in_g + full_tiles * IN_BYTES;
    uint8_t* dst = out_g + full_tiles * OUT_BYTES;
    uint8_t* b = sm.buf[0];
    const int in_b = rem * IN_REC, out_b = rem * OUT_REC;
    for (int i = tid; i < static_cast<int>(IN_BYTES); i += blockDim.x) b[i] = i < in_b ? src[i] : 0;
    __syncthreads();
    compute(b, rem, full_tiles);
    __syncthreads();
    for (int i = tid; i < out_b; i += blockDim.x) dst[i] = b[i];
  }
}

}  // namespace tsb
