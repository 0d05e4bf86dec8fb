// tsb200_api.cu — C ABI of libtsb200.so (include/tsb200.h): handles, transfers, kernel launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "nq_expand.cuh"
#include "nq_kernel.cuh"
#include "pfsp_kernels.cuh"
#include "tsb200.h"

namespace {

thread_local std::string g_last_cuda_error;

#define TSB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      g_last_cuda_error = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
      (void)cudaGetLastError();                                                                \
      return e__ == cudaErrorMemoryAllocation ? TSB_ENOMEM : TSB_ECUDA;                        \
    }                                                                                          \
  } while (0)

int env_xfer() {
  const char* s = std::getenv("TSB200_XFER");
  if (!s) return TSB_XFER_AUTO;
  if (!std::strcmp(s, "memcpy")) return TSB_XFER_MEMCPY;
  if (!std::strcmp(s, "zerocopy")) return TSB_XFER_ZEROCOPY;
  return TSB_XFER_AUTO;
}
bool env_no_register() {
  const char* s = std::getenv("TSB200_NO_REGISTER");
  return s && *s && *s != '0';
}

// Host ranges kept page-locked + mapped between calls, so that cudaMemcpyAsync is truly
// asynchronous on the caller's own arrays and the zero-copy kernels can address them.
struct HostRange {
  uintptr_t base;
  size_t len;
};
struct HostRegistry {
  std::vector<HostRange> ranges;
  bool disabled = env_no_register();
  // returns true if [p, p+bytes) is (now) page-locked
  bool ensure(const void* p, size_t bytes) {
    if (disabled || !p || !bytes) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), b = a + bytes;
    for (auto& r : ranges)
      if (a >= r.base && b <= r.base + r.len) return true;
    // merge with any overlapping registered range (the caller's array seen with a larger count)
    uintptr_t na = a, nb = b;
    for (size_t i = 0; i < ranges.size();) {
      const uintptr_t ra = ranges[i].base, rb = ra + ranges[i].len;
      if (ra < nb && na < rb) {
        cudaHostUnregister(reinterpret_cast<void*>(ra));
        na = std::min(na, ra);
        nb = std::max(nb, rb);
        ranges.erase(ranges.begin() + i);
      } else {
        ++i;
      }
    }
    if (cudaHostRegister(reinterpret_cast<void*>(na), nb - na, cudaHostRegisterPortable | cudaHostRegisterMapped) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    ranges.push_back({na, nb - na});
    return true;
  }
  void release() {
    for (auto& r : ranges) cudaHostUnregister(reinterpret_cast<void*>(r.base));
    ranges.clear();
  }
};

struct DeviceInfo {
  int sms = 0;
  bool can_use_host_ptr = false;
};
int query_device(int device, DeviceInfo& di) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return TSB_ENODEV;
  }
  if (device < 0 || device >= n) return TSB_ENODEV;
  TSB_CUDA(cudaSetDevice(device));
  TSB_CUDA(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, device));
  int v = 0;
  TSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrCanUseHostPointerForRegisteredMem, device));
  di.can_use_host_ptr = v != 0;
  return TSB_OK;
}

// common part of both handle types
struct Base {
  int device = 0, M_max = 0, xfer = TSB_XFER_AUTO;
  DeviceInfo di;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  int pipe_min = 131072;  // records from which the memcpy path is split over two streams (env TSB200_PIPE_MIN)
  int pipe_chunk = 262144;
  uint8_t *d_in = nullptr, *d_out = nullptr;  // device chunk buffers (M_max records)
  uint8_t *h_in = nullptr, *h_out = nullptr;  // pinned+mapped staging, used when the caller's arrays cannot be locked
  size_t in_rec = 0, out_rec = 0;
  HostRegistry reg;
  uint64_t launches = 0;

  int init(int dev, int M, size_t irec, size_t orec) {
    device = dev;
    M_max = M;
    in_rec = irec;
    out_rec = orec;
    xfer = env_xfer();
    int rc = query_device(dev, di);
    if (rc != TSB_OK) return rc;
    TSB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    TSB_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    if (const char* v = std::getenv("TSB200_PIPE_MIN")) pipe_min = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("TSB200_PIPE_CHUNK")) pipe_chunk = std::max(1024, std::atoi(v)) & ~1023;
    TSB_CUDA(cudaMalloc(&d_in, in_rec * M + 256));
    TSB_CUDA(cudaMalloc(&d_out, out_rec * M + 256));
    return TSB_OK;
  }
  int ensure_staging() {
    if (!h_in) TSB_CUDA(cudaHostAlloc(&h_in, in_rec * M_max + 256, cudaHostAllocPortable | cudaHostAllocMapped));
    if (!h_out) TSB_CUDA(cudaHostAlloc(&h_out, out_rec * M_max + 256, cudaHostAllocPortable | cudaHostAllocMapped));
    return TSB_OK;
  }
  void fini() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    reg.release();
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
  }

  // Host-buffer evaluation shared by N-Queens and PFSP.  `launch(in_dev, out_dev, count, stream)`
  // enqueues the evaluator kernel.
  template <class Launch>
  int evaluate_host(const void* in, int count, void* out, Launch&& launch) {
    const size_t in_b = in_rec * count, out_b = out_rec * count;
    const bool in_locked = reg.ensure(in, in_b), out_locked = reg.ensure(out, out_b);
    // AUTO: zero-copy whenever the caller's arrays could be page-locked and are 16-byte aligned (measured
    // fastest at every chunk size, profiles/xfer_sweep_r1.txt); otherwise copies, pipelined when large
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const bool zc_ok = in_locked && out_locked && aligned && di.can_use_host_ptr;
    int mode = xfer == TSB_XFER_AUTO ? (zc_ok ? TSB_XFER_ZEROCOPY : TSB_XFER_MEMCPY) : xfer;
    if (mode == TSB_XFER_ZEROCOPY && !zc_ok) mode = TSB_XFER_MEMCPY;

    if (mode == TSB_XFER_ZEROCOPY) {
      // the kernel's TMA engine pulls the chunk over PCIe and pushes the results back: one launch,
      // reads and writes overlap on the full-duplex link
      int rc = launch(static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), count, stream);
      if (rc != TSB_OK) return rc;
      TSB_CUDA(cudaStreamSynchronize(stream));
      return TSB_OK;
    }
    const void* src = in;
    void* dst = out;
    if (!in_locked || !out_locked) {
      int rc = ensure_staging();
      if (rc != TSB_OK) return rc;
    }
    if (!in_locked) {
      std::memcpy(h_in, in, in_b);
      src = h_in;
    }
    if (!out_locked) dst = h_out;
    if (count >= pipe_min && count > pipe_chunk) {
      // large chunk: sub-chunks alternate between two streams so that the upload of one overlaps the
      // download of the previous one (PCIe is full duplex) and the kernel of the one in between
      const cudaStream_t st[2] = {stream, stream2};
      int i = 0;
      for (int off = 0; off < count; off += pipe_chunk, ++i) {
        const int n = std::min(pipe_chunk, count - off);
        cudaStream_t s = st[i & 1];
        TSB_CUDA(cudaMemcpyAsync(d_in + in_rec * off, static_cast<const uint8_t*>(src) + in_rec * off, in_rec * n,
                                 cudaMemcpyHostToDevice, s));
        int rc = launch(d_in + in_rec * off, d_out + out_rec * off, n, s);
        if (rc != TSB_OK) return rc;
        TSB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + out_rec * off, d_out + out_rec * off, out_rec * n,
                                 cudaMemcpyDeviceToHost, s));
      }
      TSB_CUDA(cudaStreamSynchronize(stream));
      TSB_CUDA(cudaStreamSynchronize(stream2));
    } else {
      TSB_CUDA(cudaMemcpyAsync(d_in, src, in_b, cudaMemcpyHostToDevice, stream));
      int rc = launch(d_in, d_out, count, stream);
      if (rc != TSB_OK) return rc;
      TSB_CUDA(cudaMemcpyAsync(dst, d_out, out_b, cudaMemcpyDeviceToHost, stream));
      TSB_CUDA(cudaStreamSynchronize(stream));
    }
    if (!out_locked) std::memcpy(out, h_out, out_b);
    return TSB_OK;
  }
};

// persistent grid: enough CTAs to fill the GPU, never more than there are full tiles
template <class K>
int grid_for(K kernel, int threads, size_t smem, long long count, int tile, int sms, int* grid, int* cache) {
  int per_sm = *cache;
  if (per_sm <= 0) {
    TSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    *cache = per_sm;
  }
  const long long tiles = std::max<long long>(1, count / tile);
  *grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(per_sm) * sms));
  return TSB_OK;
}

}  // namespace

// ============================================================================ N-Queens
struct tsb_nq : Base {
  int N = 0, g = 1;
  int variant = 0;  // env TSB200_NQ_VARIANT (kernel A/B experiments)
  int occ = 0;      // cached CTAs per SM
  bool attr_set = false;
  // fused expand (evaluate + generate_children on the device) and the device-resident pool
  uint8_t* d_cmask = nullptr;
  int* d_tile = nullptr;
  long long exp_cap = 0;  // parents the two arrays above are sized for
  tsb::ExpandCounters* d_ctr = nullptr;
  tsb::ExpandCounters* h_ctr = nullptr;  // pinned
  uint8_t* d_children = nullptr;         // host-buffer expand: device image of the children
  size_t d_children_bytes = 0;
  bool exp_attr_set = false;
  int occ_count = 0, occ_write = 0;
  uint8_t* pool = nullptr;  // device-resident pool: `pool_size` nodes of 21 B, capacity `pool_cap`
  long long pool_size = 0, pool_cap = 0;
};

namespace {

template <int N, int VAR>
int launch_nq_n(tsb_nq* h, const uint8_t* in, uint8_t* out, long long count, cudaStream_t s) {
  auto kernel = tsb::nq_evaluate_kernel<N, VAR>;
  const size_t smem = sizeof(tsb::NqSmem<N>) + 128;
  if (!h->attr_set) {
    TSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    h->attr_set = true;
  }
  int grid = 1;
  int rc = grid_for(kernel, tsb::NQ_THREADS, smem, count, tsb::NQ_TILE, h->di.sms, &grid, &h->occ);
  if (rc != TSB_OK) return rc;
  kernel<<<grid, tsb::NQ_THREADS, smem, s>>>(in, out, count);
  TSB_CUDA(cudaGetLastError());
  h->launches++;
  return TSB_OK;
}

int launch_nq(tsb_nq* h, const uint8_t* in, uint8_t* out, long long count, cudaStream_t s) {
  if (h->N == 17 && h->variant == 1) return launch_nq_n<17, 1>(h, in, out, count, s);  // A/B experiment: byte alignment as IMAD.HI
  switch (h->N) {
#define TSB_NQ_CASE(n) \
  case n:              \
    return launch_nq_n<n, 0>(h, in, out, count, s);
    TSB_NQ_CASE(1) TSB_NQ_CASE(2) TSB_NQ_CASE(3) TSB_NQ_CASE(4) TSB_NQ_CASE(5) TSB_NQ_CASE(6) TSB_NQ_CASE(7)
    TSB_NQ_CASE(8) TSB_NQ_CASE(9) TSB_NQ_CASE(10) TSB_NQ_CASE(11) TSB_NQ_CASE(12) TSB_NQ_CASE(13)
    TSB_NQ_CASE(14) TSB_NQ_CASE(15) TSB_NQ_CASE(16) TSB_NQ_CASE(17) TSB_NQ_CASE(18) TSB_NQ_CASE(19)
    TSB_NQ_CASE(20)
#undef TSB_NQ_CASE
  }
  return TSB_EINVAL;
}

}  // namespace

namespace {

int nq_expand_reserve(tsb_nq* h, long long count) {
  if (count > h->exp_cap) {
    if (h->d_cmask) cudaFree(h->d_cmask);
    if (h->d_tile) cudaFree(h->d_tile);
    h->d_cmask = nullptr;
    h->d_tile = nullptr;
    h->exp_cap = 0;
    const long long cap = std::max<long long>(count, h->M_max);
    TSB_CUDA(cudaMalloc(&h->d_cmask, static_cast<size_t>(cap + tsb::NQ_TILE) * 4));
    TSB_CUDA(cudaMalloc(&h->d_tile, static_cast<size_t>(cap / tsb::NQ_TILE + 4) * sizeof(int)));
    h->exp_cap = cap;
  }
  if (!h->d_ctr) TSB_CUDA(cudaMalloc(&h->d_ctr, sizeof(tsb::ExpandCounters)));
  if (!h->h_ctr) TSB_CUDA(cudaHostAlloc(&h->h_ctr, sizeof(tsb::ExpandCounters), cudaHostAllocPortable));
  return TSB_OK;
}

// K1 + K2 + K3 on `s`; children_d may have any alignment; synchronous (the counts come back)
template <int N>
int nq_expand_n(tsb_nq* h, const uint8_t* parents_d, long long count, uint8_t* children_d, cudaStream_t s,
                unsigned long long* n_children, unsigned long long* n_solutions) {
  int rc = nq_expand_reserve(h, count);
  if (rc != TSB_OK) return rc;
  auto k1 = tsb::nq_expand_count_kernel<N>;
  auto k3 = tsb::nq_expand_write_kernel<N>;
  const size_t smem1 = sizeof(tsb::NqCountSmem<N>) + 128, smem3 = sizeof(tsb::NqWriteSmem) + 128;
  if (!h->exp_attr_set) {
    TSB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem1)));
    TSB_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem3)));
    h->exp_attr_set = true;
  }
  const long long tiles = (count + tsb::NQ_TILE - 1) / tsb::NQ_TILE;
  int g1 = 1, g3 = 1;
  rc = grid_for(k1, tsb::NQ_THREADS, smem1, count, tsb::NQ_TILE, h->di.sms, &g1, &h->occ_count);
  if (rc != TSB_OK) return rc;
  rc = grid_for(k3, tsb::NQ_THREADS, smem3, tiles * tsb::NQ_TILE, tsb::NQ_TILE, h->di.sms, &g3, &h->occ_write);
  if (rc != TSB_OK) return rc;
  TSB_CUDA(cudaMemsetAsync(h->d_ctr, 0, sizeof(tsb::ExpandCounters), s));
  k1<<<g1, tsb::NQ_THREADS, smem1, s>>>(parents_d, h->d_cmask, count, h->d_tile, h->d_ctr);
  tsb::scan_tiles_kernel<<<1, 1024, 0, s>>>(h->d_tile, static_cast<int>(tiles), h->d_ctr);
  k3<<<g3, tsb::NQ_THREADS, smem3, s>>>(parents_d, reinterpret_cast<const uint32_t*>(h->d_cmask), h->d_tile, count,
                                        children_d);
  TSB_CUDA(cudaGetLastError());
  h->launches += 3;
  TSB_CUDA(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(tsb::ExpandCounters), cudaMemcpyDeviceToHost, s));
  TSB_CUDA(cudaStreamSynchronize(s));
  *n_children = h->h_ctr->children;
  *n_solutions = h->h_ctr->solutions;
  return TSB_OK;
}

int nq_expand_dispatch(tsb_nq* h, const uint8_t* parents_d, long long count, uint8_t* children_d, cudaStream_t s,
                       unsigned long long* nc, unsigned long long* ns) {
  switch (h->N) {
#define TSB_NQ_CASE(n) \
  case n:              \
    return nq_expand_n<n>(h, parents_d, count, children_d, s, nc, ns);
    TSB_NQ_CASE(1) TSB_NQ_CASE(2) TSB_NQ_CASE(3) TSB_NQ_CASE(4) TSB_NQ_CASE(5) TSB_NQ_CASE(6) TSB_NQ_CASE(7)
    TSB_NQ_CASE(8) TSB_NQ_CASE(9) TSB_NQ_CASE(10) TSB_NQ_CASE(11) TSB_NQ_CASE(12) TSB_NQ_CASE(13)
    TSB_NQ_CASE(14) TSB_NQ_CASE(15) TSB_NQ_CASE(16) TSB_NQ_CASE(17) TSB_NQ_CASE(18) TSB_NQ_CASE(19)
    TSB_NQ_CASE(20)
#undef TSB_NQ_CASE
  }
  return TSB_EINVAL;
}

}  // namespace

// ============================================================================ PFSP
struct tsb_pfsp : Base {
  int jobs = 0, machines = 0, pairs = 0, mt = 0;  // mt = template machine count (5, 10 or 20)
  tsb::PfspLb1Tables* d_tab1 = nullptr;
  tsb::PfspLb2Tables* d_tab2 = nullptr;
  bool attr_set[3] = {false, false, false};
  int occ[3] = {0, 0, 0};
};

namespace {

template <int KIND, int M>
int launch_lb1_km(tsb_pfsp* h, const uint8_t* in, uint8_t* out, long long count, cudaStream_t s) {
  auto kernel = tsb::pfsp_lb1_kernel<KIND, M>;
  const size_t smem = sizeof(tsb::Lb1Smem) + 128;
  if (!h->attr_set[KIND]) {
    TSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    h->attr_set[KIND] = true;
  }
  int grid = 1;
  int rc = grid_for(kernel, tsb::PF_THREADS, smem, count, tsb::PF_TILE, h->di.sms, &grid, &h->occ[KIND]);
  if (rc != TSB_OK) return rc;
  kernel<<<grid, tsb::PF_THREADS, smem, s>>>(in, out, count, h->d_tab1);
  TSB_CUDA(cudaGetLastError());
  h->launches++;
  return TSB_OK;
}

template <int M>
int launch_lb2_m(tsb_pfsp* h, const uint8_t* in, uint8_t* out, long long count, int best, cudaStream_t s) {
  auto kernel = tsb::pfsp_lb2_kernel<M>;
  const size_t smem = sizeof(tsb::Lb2Smem) + 128;
  if (!h->attr_set[2]) {
    TSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    h->attr_set[2] = true;
  }
  int grid = 1;
  int rc = grid_for(kernel, tsb::PF_THREADS, smem, count, tsb::PF_TILE, h->di.sms, &grid, &h->occ[2]);
  if (rc != TSB_OK) return rc;
  kernel<<<grid, tsb::PF_THREADS, smem, s>>>(in, out, count, h->d_tab1, h->d_tab2, best);
  TSB_CUDA(cudaGetLastError());
  h->launches++;
  return TSB_OK;
}

int launch_pfsp(tsb_pfsp* h, int lb_kind, const uint8_t* in, uint8_t* out, long long count, int64_t best64,
                cudaStream_t s) {
  // bounds are int32 and `lb > best` can never hold for best >= INT32_MAX (Chapel's max(int) under --ub 0)
  const int best = best64 > INT_MAX ? INT_MAX : best64 < INT_MIN ? INT_MIN : static_cast<int>(best64);
#define TSB_PF_DISPATCH(M)                                                          \
  if (lb_kind == TSB_LB1) return launch_lb1_km<1, M>(h, in, out, count, s);         \
  if (lb_kind == TSB_LB1_D) return launch_lb1_km<0, M>(h, in, out, count, s);       \
  return launch_lb2_m<M>(h, in, out, count, best, s);
  if (h->mt == 5) { TSB_PF_DISPATCH(5) }
  if (h->mt == 10) { TSB_PF_DISPATCH(10) }
  TSB_PF_DISPATCH(20)
#undef TSB_PF_DISPATCH
}

}  // namespace

// ============================================================================ exported C ABI
extern "C" {

const char* tsb_version(void) { return "tsb200 0.1 (sm_100a)"; }

const char* tsb_strerror(int code) {
  switch (code) {
    case TSB_OK: return "ok";
    case TSB_EINVAL: return "invalid argument";
    case TSB_ECUDA: return "CUDA runtime error (see tsb_last_cuda_error)";
    case TSB_ENOMEM: return "out of memory";
    case TSB_ENODEV: return "no such CUDA device";
    case TSB_EALIGN: return "device pointer not 16-byte aligned";
    case TSB_EUNSUPPORTED: return "unsupported instance shape (jobs must be 20, machines 1..20)";
  }
  return "unknown error";
}
const char* tsb_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int tsb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return TSB_ENODEV;
  }
  return n;
}

int tsb_init_devices(int n) {
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) {
    (void)cudaGetLastError();
    return TSB_ENODEV;
  }
  for (int d = 0; d < n && d < have; d++) {
    TSB_CUDA(cudaSetDevice(d));
    TSB_CUDA(cudaFree(nullptr));
  }
  return TSB_OK;
}

// ---------------------------------------------------------------- N-Queens
int tsb_nq_create(tsb_nq** out, int device, int N, int g, int M_max) {
  if (!out || N < 1 || N > TSB_MAX_QUEENS || g < 1 || M_max < 1) return TSB_EINVAL;
  tsb_nq* h = new (std::nothrow) tsb_nq();
  if (!h) return TSB_ENOMEM;
  h->N = N;
  h->g = g;
  if (const char* v = std::getenv("TSB200_NQ_VARIANT")) h->variant = std::atoi(v);
  int rc = h->init(device, M_max, sizeof(tsb_nq_node), static_cast<size_t>(N));
  if (rc != TSB_OK) {
    h->fini();
    delete h;
    return rc;
  }
  *out = h;
  return TSB_OK;
}

void tsb_nq_destroy(tsb_nq* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->d_cmask) cudaFree(h->d_cmask);
  if (h->d_tile) cudaFree(h->d_tile);
  if (h->d_ctr) cudaFree(h->d_ctr);
  if (h->h_ctr) cudaFreeHost(h->h_ctr);
  if (h->d_children) cudaFree(h->d_children);
  if (h->pool) cudaFree(h->pool);
  h->fini();
  delete h;
}

// ---- fused evaluate + generate_children, and the device-resident pool
int tsb_nq_expand_device(tsb_nq* h, const void* parents_d, int count, void* children_d, uint64_t* n_children,
                         uint64_t* n_solutions, void* stream) {
  if (!h || count < 0 || !n_children || !n_solutions) return TSB_EINVAL;
  *n_children = *n_solutions = 0;
  if (count == 0) return TSB_OK;
  if (!parents_d || !children_d) return TSB_EINVAL;
  if (reinterpret_cast<uintptr_t>(parents_d) & 15) return TSB_EALIGN;
  TSB_CUDA(cudaSetDevice(h->device));
  unsigned long long nc = 0, ns = 0;
  int rc = nq_expand_dispatch(h, static_cast<const uint8_t*>(parents_d), count, static_cast<uint8_t*>(children_d),
                              stream ? static_cast<cudaStream_t>(stream) : h->stream, &nc, &ns);
  *n_children = nc;
  *n_solutions = ns;
  return rc;
}

int tsb_nq_expand(tsb_nq* h, const void* parents, int count, void* children, uint64_t capacity, uint64_t* n_children,
                  uint64_t* n_solutions) {
  if (!h || count < 0 || count > h->M_max || !n_children || !n_solutions) return TSB_EINVAL;
  *n_children = *n_solutions = 0;
  if (count == 0) return TSB_OK;
  if (!parents || !children) return TSB_EINVAL;
  TSB_CUDA(cudaSetDevice(h->device));
  const size_t need = static_cast<size_t>(h->M_max) * h->N * sizeof(tsb_nq_node) + 64;
  if (h->d_children_bytes < need) {
    if (h->d_children) cudaFree(h->d_children);
    h->d_children = nullptr;
    h->d_children_bytes = 0;
    TSB_CUDA(cudaMalloc(&h->d_children, need));
    h->d_children_bytes = need;
  }
  const size_t in_b = sizeof(tsb_nq_node) * static_cast<size_t>(count);
  const bool in_locked = h->reg.ensure(parents, in_b);
  const void* src = parents;
  if (!in_locked) {
    int rc = h->ensure_staging();
    if (rc != TSB_OK) return rc;
    std::memcpy(h->h_in, parents, in_b);
    src = h->h_in;
  }
  TSB_CUDA(cudaMemcpyAsync(h->d_in, src, in_b, cudaMemcpyHostToDevice, h->stream));
  unsigned long long nc = 0, ns = 0;
  int rc = nq_expand_dispatch(h, h->d_in, count, h->d_children, h->stream, &nc, &ns);
  if (rc != TSB_OK) return rc;
  *n_children = nc;
  *n_solutions = ns;
  if (nc > capacity) return TSB_ENOMEM;  // the caller's children array is too small; counts are valid
  if (nc) TSB_CUDA(cudaMemcpy(children, h->d_children, nc * sizeof(tsb_nq_node), cudaMemcpyDeviceToHost));
  return TSB_OK;
}

int tsb_nq_pool_push(tsb_nq* h, const void* nodes, int64_t n) {
  if (!h || n < 0 || (n && !nodes)) return TSB_EINVAL;
  TSB_CUDA(cudaSetDevice(h->device));
  const long long need = h->pool_size + n;
  if (need > h->pool_cap) {
    // room for two worst-case rounds up front, so that growth (a device-wide realloc + copy) stays rare
    const long long cap = std::max<long long>({need, 2 * h->pool_cap, 1LL << 20, 2LL * h->M_max * h->N});
    uint8_t* np = nullptr;
    TSB_CUDA(cudaMalloc(&np, static_cast<size_t>(cap) * sizeof(tsb_nq_node) + 64));
    if (h->pool_size)
      TSB_CUDA(cudaMemcpy(np, h->pool, static_cast<size_t>(h->pool_size) * sizeof(tsb_nq_node), cudaMemcpyDeviceToDevice));
    if (h->pool) cudaFree(h->pool);
    h->pool = np;
    h->pool_cap = cap;
  }
  if (n)
    TSB_CUDA(cudaMemcpy(h->pool + h->pool_size * sizeof(tsb_nq_node), nodes, static_cast<size_t>(n) * sizeof(tsb_nq_node),
                        cudaMemcpyHostToDevice));
  h->pool_size = need;
  return TSB_OK;
}

int64_t tsb_nq_pool_size(const tsb_nq* h) { return h ? h->pool_size : -1; }

int tsb_nq_pool_step(tsb_nq* h, int m, int M, int64_t* n_parents, uint64_t* n_children, uint64_t* n_solutions) {
  if (!h || m < 1 || M < 1 || M > h->M_max || !n_parents || !n_children || !n_solutions) return TSB_EINVAL;
  *n_parents = 0;
  *n_children = *n_solutions = 0;
  if (h->pool_size < m) return TSB_OK;  // popBackBulk returns 0 below m (lib/commons/Pool.chpl:50-59)
  TSB_CUDA(cudaSetDevice(h->device));
  const long long n = std::min<long long>(h->pool_size, M);
  const long long base = h->pool_size - n;
  // room for the worst case (every slot of every parent survives) before anything is overwritten
  int rc = TSB_OK;
  const long long worst = base + n * h->N;
  if (worst > h->pool_cap) {
    const long long keep = h->pool_size;
    const long long cap = std::max<long long>(worst, 2 * h->pool_cap);
    uint8_t* np = nullptr;
    TSB_CUDA(cudaMalloc(&np, static_cast<size_t>(cap) * sizeof(tsb_nq_node) + 64));
    TSB_CUDA(cudaMemcpy(np, h->pool, static_cast<size_t>(keep) * sizeof(tsb_nq_node), cudaMemcpyDeviceToDevice));
    cudaFree(h->pool);
    h->pool = np;
    h->pool_cap = cap;
  }
  // the newest n nodes become the chunk (order preserved); their children are appended where they were
  TSB_CUDA(cudaMemcpyAsync(h->d_in, h->pool + base * sizeof(tsb_nq_node), static_cast<size_t>(n) * sizeof(tsb_nq_node),
                           cudaMemcpyDeviceToDevice, h->stream));
  unsigned long long nc = 0, ns = 0;
  rc = nq_expand_dispatch(h, h->d_in, n, h->pool + base * sizeof(tsb_nq_node), h->stream, &nc, &ns);
  if (rc != TSB_OK) return rc;
  h->pool_size = base + static_cast<long long>(nc);
  *n_parents = n;
  *n_children = nc;
  *n_solutions = ns;
  return TSB_OK;
}

int tsb_nq_pool_drain(tsb_nq* h, void* nodes, int64_t capacity, int64_t* n) {
  if (!h || !n || capacity < 0) return TSB_EINVAL;
  *n = h->pool_size;
  if (h->pool_size > capacity) return TSB_ENOMEM;
  TSB_CUDA(cudaSetDevice(h->device));
  if (h->pool_size)
    TSB_CUDA(cudaMemcpy(nodes, h->pool, static_cast<size_t>(h->pool_size) * sizeof(tsb_nq_node), cudaMemcpyDeviceToHost));
  h->pool_size = 0;
  return TSB_OK;
}

int tsb_nq_evaluate(tsb_nq* h, const void* parents, int count, uint8_t* labels) {
  if (!h || count < 0 || count > h->M_max) return TSB_EINVAL;
  if (count == 0) return TSB_OK;
  if (!parents || !labels) return TSB_EINVAL;
  TSB_CUDA(cudaSetDevice(h->device));
  return h->evaluate_host(parents, count, labels, [h](const uint8_t* in, uint8_t* out, int n, cudaStream_t s) {
    return launch_nq(h, in, out, n, s);
  });
}

int tsb_nq_evaluate_device(tsb_nq* h, const void* parents_d, int count, uint8_t* labels_d, void* stream) {
  if (!h || count < 0) return TSB_EINVAL;
  if (count == 0) return TSB_OK;
  if (!parents_d || !labels_d) return TSB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(parents_d) | reinterpret_cast<uintptr_t>(labels_d)) & 15) return TSB_EALIGN;
  TSB_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  return launch_nq(h, static_cast<const uint8_t*>(parents_d), labels_d, count, s);
}

int tsb_nq_set_xfer(tsb_nq* h, int mode) {
  if (!h || mode < 0 || mode > 2) return TSB_EINVAL;
  h->xfer = mode;
  return TSB_OK;
}
uint64_t tsb_nq_kernel_launches(const tsb_nq* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------- PFSP
int tsb_pfsp_create(tsb_pfsp** out, int device, int jobs, int machines, int M_max, const int32_t* p_times,
                    const int32_t* min_heads, const int32_t* min_tails, int nb_pairs, const int32_t* johnson,
                    const int32_t* lags, const int32_t* mp0, const int32_t* mp1, const int32_t* mp_order) {
  if (!out || !p_times || !min_heads || !min_tails || M_max < 1 || nb_pairs < 0) return TSB_EINVAL;
  if (nb_pairs > 0 && (!johnson || !lags || !mp0 || !mp1 || !mp_order)) return TSB_EINVAL;
  if (jobs != TSB_MAX_JOBS || machines < 1 || machines > TSB_MAX_MACHINES || nb_pairs > TSB_MAX_PAIRS)
    return TSB_EUNSUPPORTED;
  tsb_pfsp* h = new (std::nothrow) tsb_pfsp();
  if (!h) return TSB_ENOMEM;
  h->jobs = jobs;
  h->machines = machines;
  h->pairs = nb_pairs;
  h->mt = machines <= 5 ? 5 : machines <= 10 ? 10 : 20;
  int rc = h->init(device, M_max, sizeof(tsb_pfsp_node), static_cast<size_t>(jobs) * 4);
  // tables -> device blobs (zero padding up to the template machine count is value-neutral:
  // the reference itself evaluates 20-wide zero-padded tuples, lib/pfsp/Bound_simple.chpl:125-135)
  std::vector<tsb::PfspLb1Tables> t1v(1);
  std::vector<tsb::PfspLb2Tables> t2v(1);
  tsb::PfspLb1Tables& t1 = t1v[0];
  tsb::PfspLb2Tables& t2 = t2v[0];
  std::memset(&t1, 0, sizeof(t1));
  std::memset(&t2, 0, sizeof(t2));
  const int mp = tsb::row_stride(h->mt);
  t1.jobs = jobs;
  t1.machines = machines;
  t1.pairs = nb_pairs;
  t1.mp = mp;
  for (int k = 0; k < machines; k++) {
    t1.min_heads[k] = min_heads[k];
    t1.min_tails[k] = min_tails[k];
    for (int j = 0; j < jobs; j++) {
      t1.total[k] += p_times[k * jobs + j];
      t1.pj[j * mp + k] = p_times[k * jobs + j];
      t2.pm[k * jobs + j] = p_times[k * jobs + j];
    }
  }
  bool bad = false;
  for (int i = 0; i < nb_pairs; i++) {
    t2.mp0[i] = mp0[i];
    t2.mp1[i] = mp1[i];
    t2.order[i] = mp_order[i];
    bad |= mp0[i] < 0 || mp0[i] >= machines || mp1[i] < 0 || mp1[i] >= machines || mp_order[i] < 0 ||
           mp_order[i] >= nb_pairs;
    for (int j = 0; j < jobs; j++) {
      t2.johnson[i * jobs + j] = johnson[i * jobs + j];
      t2.lags[i * jobs + j] = lags[i * jobs + j];
      bad |= johnson[i * jobs + j] < 0 || johnson[i * jobs + j] >= jobs;
    }
  }
  if (rc == TSB_OK && bad) rc = TSB_EINVAL;
  auto upload = [&]() -> int {
    TSB_CUDA(cudaMalloc(&h->d_tab1, sizeof(t1)));
    TSB_CUDA(cudaMalloc(&h->d_tab2, sizeof(t2)));
    TSB_CUDA(cudaMemcpy(h->d_tab1, &t1, sizeof(t1), cudaMemcpyHostToDevice));
    TSB_CUDA(cudaMemcpy(h->d_tab2, &t2, sizeof(t2), cudaMemcpyHostToDevice));
    return TSB_OK;
  };
  if (rc == TSB_OK) rc = upload();
  if (rc != TSB_OK) {
    tsb_pfsp_destroy(h);
    return rc;
  }
  *out = h;
  return TSB_OK;
}

void tsb_pfsp_destroy(tsb_pfsp* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->d_tab1) cudaFree(h->d_tab1);
  if (h->d_tab2) cudaFree(h->d_tab2);
  h->fini();
  delete h;
}

int tsb_pfsp_evaluate(tsb_pfsp* h, int lb_kind, const void* parents, int count, int64_t best, int32_t* bounds) {
  if (!h || count < 0 || count > h->M_max || lb_kind < 0 || lb_kind > 2) return TSB_EINVAL;
  if (lb_kind == TSB_LB2 && h->pairs == 0) return TSB_EINVAL;
  if (count == 0) return TSB_OK;
  if (!parents || !bounds) return TSB_EINVAL;
  TSB_CUDA(cudaSetDevice(h->device));
  return h->evaluate_host(parents, count, bounds,
                          [h, lb_kind, best](const uint8_t* in, uint8_t* out, int n, cudaStream_t s) {
                            return launch_pfsp(h, lb_kind, in, out, n, best, s);
                          });
}

int tsb_pfsp_evaluate_device(tsb_pfsp* h, int lb_kind, const void* parents_d, int count, int64_t best,
                             int32_t* bounds_d, void* stream) {
  if (!h || count < 0 || lb_kind < 0 || lb_kind > 2) return TSB_EINVAL;
  if (lb_kind == TSB_LB2 && h->pairs == 0) return TSB_EINVAL;
  if (count == 0) return TSB_OK;
  if (!parents_d || !bounds_d) return TSB_EINVAL;
  if ((reinterpret_cast<uintptr_t>(parents_d) | reinterpret_cast<uintptr_t>(bounds_d)) & 15) return TSB_EALIGN;
  TSB_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  return launch_pfsp(h, lb_kind, static_cast<const uint8_t*>(parents_d), reinterpret_cast<uint8_t*>(bounds_d),
                     count, best, s);
}

int tsb_pfsp_set_xfer(tsb_pfsp* h, int mode) {
  if (!h || mode < 0 || mode > 2) return TSB_EINVAL;
  h->xfer = mode;
  return TSB_OK;
}
uint64_t tsb_pfsp_kernel_launches(const tsb_pfsp* h) { return h ? h->launches : 0; }

}  // extern "C"
