// C ABI of libngicp_b200 (include/ngicp_b200.h): handle / index bookkeeping that mirrors
// nano_gicp::NanoGICP's cloud, tree and covariance members (reference
// src/dlio/src/nano_gicp/nano_gicp.cc:97-203) and the host-side LM driver
// (src/dlio/src/nano_gicp/lsq_registration.cc:108-229). No CPU fallback anywhere: every compute
// entry point launches the sm_100a kernels or returns an error.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "internal.h"
#include "linearize.cuh"
#include "lm_host.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace ngicp {

// Targets up to this size get the fine index levels too (scan-to-scan targets have their covariances computed by K2); a
// submap's covariances come from its keyframes (odom.cc:1719-1729), so its table stays small for the correspondence search.
constexpr size_t kFineTargetMax = 262144;

static thread_local std::string g_thread_err;
void set_thread_error(const std::string& s) { g_thread_err = s; }
int fail(Handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_thread_err = msg;
  return code;
}

namespace {

__global__ void __launch_bounds__(256) keys_unsort_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ keys, int n,
                                                          unsigned long long* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[__float_as_int(__ldg(&pts[j].w))] = keys[j];
}
__global__ void __launch_bounds__(256) pack_f4_kernel(const float* __restrict__ in, int stride, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(in[(size_t)i * stride], in[(size_t)i * stride + 1], in[(size_t)i * stride + 2], 0.f);
}
__global__ void __launch_bounds__(256) unsort_points_kernel(const float4* __restrict__ pts, int n, float* __restrict__ out3) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float4 p = __ldg(pts + j);
  const int o = __float_as_int(p.w);
  out3[(size_t)o * 3] = p.x; out3[(size_t)o * 3 + 1] = p.y; out3[(size_t)o * 3 + 2] = p.z;
}

int use_device(Handle* h) {
  NGICP_CUDA(h, cudaSetDevice(h->device));
  return NGICP_OK;
}

int ensure_stage(Handle* h, size_t bytes) {
  if (h->stage_done) NGICP_CUDA(h, cudaEventSynchronize(h->stage_done));  // previous H2D out of the buffer has landed
  if (h->stage_cap >= bytes) return NGICP_OK;
  if (h->stage_host) NGICP_CUDA(h, cudaFreeHost(h->stage_host));
  h->stage_host = nullptr; h->stage_cap = 0;
  size_t cap = 1 << 20;
  while (cap < bytes) cap <<= 1;
  NGICP_CUDA(h, cudaHostAlloc(&h->stage_host, cap, cudaHostAllocDefault));
  h->stage_cap = cap;
  return NGICP_OK;
}

// host AoS (xyz at floats 0..2 of every `stride_bytes` record) -> device.
//  * Page-locked memory (cudaHostAlloc / cudaHostRegister) with a stride of at most 32 bytes is copied as it is,
//    asynchronously and without touching it on the CPU; the kernels read it with its stride (*stride_floats). The copy
//    is still in flight when this function returns: the public entry points call finish_input() before THEY return
//    (unless the caller opted into ngicp_set_async_input, include/ngicp_b200.h).
//  * Pageable memory (what PCL clouds are) is packed to 12 bytes per point into the handle's pinned staging buffer in
//    four slices, each slice's H2D copy leaving while the next one is packed; the caller's buffer is not referenced
//    after return.
void pack_xyz(float* __restrict__ st, const char* __restrict__ src, size_t n, size_t stride_bytes) {
#if defined(__SSE2__)
  // one unaligned 16-byte load and store per point (the 4th lane is overwritten by the next point's x)
  if (stride_bytes >= 16 && n > 1) {
    size_t i = 0;
    for (; i + 1 < n; i++) _mm_storeu_ps(st + 3 * i, _mm_loadu_ps(reinterpret_cast<const float*>(src + i * stride_bytes)));
    const float* p = reinterpret_cast<const float*>(src + i * stride_bytes);
    st[3 * i] = p[0]; st[3 * i + 1] = p[1]; st[3 * i + 2] = p[2];
    return;
  }
#endif
  for (size_t i = 0; i < n; i++) {
    const float* p = reinterpret_cast<const float*>(src + i * stride_bytes);
    st[3 * i] = p[0]; st[3 * i + 1] = p[1]; st[3 * i + 2] = p[2];
  }
}

int upload_xyz(Handle* h, const void* points, size_t n, size_t stride_bytes, float** d_xyz, int* stride_floats = nullptr) {
  if (stride_bytes < 12 || (stride_bytes % 4) != 0) return fail(h, NGICP_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
  if (stride_floats) {
    *stride_floats = 3;
    cudaPointerAttributes at;
    const cudaError_t q = cudaPointerGetAttributes(&at, points);
    if (q != cudaSuccess) cudaGetLastError();   // plain pageable memory on older drivers: not an error
    if (q == cudaSuccess && at.type == cudaMemoryTypeHost && stride_bytes <= 32) {
      NGICP_CUDA(h, dev_alloc(d_xyz, n * (stride_bytes / 4), h->stream));
      NGICP_CUDA(h, cudaMemcpyAsync(*d_xyz, points, n * stride_bytes, cudaMemcpyHostToDevice, h->stream));
      NGICP_CUDA(h, cudaEventRecord(h->input_copied, h->stream));
      h->input_pending = true;
      *stride_floats = (int)(stride_bytes / 4);
      return NGICP_OK;
    }
  }
  if (int rc = ensure_stage(h, n * 12)) return rc;
  float* st = static_cast<float*>(h->stage_host);
  const char* src = static_cast<const char*>(points);
  NGICP_CUDA(h, dev_alloc(d_xyz, n * 3, h->stream));
  if (stride_bytes == 12) {
    std::memcpy(st, src, n * 12);
    NGICP_CUDA(h, cudaMemcpyAsync(*d_xyz, st, n * 12, cudaMemcpyHostToDevice, h->stream));
  } else {
    const size_t slices = n >= 16384 ? 4 : 1, per = (n + slices - 1) / slices;
    for (size_t a = 0; a < n; a += per) {
      const size_t m = std::min(per, n - a);
      pack_xyz(st + 3 * a, src + a * stride_bytes, m, stride_bytes);
      NGICP_CUDA(h, cudaMemcpyAsync(*d_xyz + 3 * a, st + 3 * a, m * 12, cudaMemcpyHostToDevice, h->stream));
    }
  }
  NGICP_CUDA(h, cudaEventRecord(h->stage_done, h->stream));
  return NGICP_OK;
}

// called by the public entry points that took a host cloud, right before they return
int finish_input(Handle* h) {
  if (h->input_pending && !h->async_input) NGICP_CUDA(h, cudaEventSynchronize(h->input_copied));
  h->input_pending = false;
  return NGICP_OK;
}

void drop_covs(Handle* h, int which) {
  CovSet& c = h->covs[which];
  dev_free(c.cov6, h->stream);
  c.cov6 = nullptr; c.n = 0; c.valid = false; c.pending.clear(); c.pending_n = 0;
}

void release_index(Handle* h, Index* idx) {
  if (!idx) return;
  if (idx->refs.fetch_sub(1) == 1) free_index(idx, h ? h->stream : (cudaStream_t)0);
}

// materialise covariances handed over with set_covariances once an index of matching size is attached
int apply_pending(Handle* h, int which) {
  CovSet& c = h->covs[which];
  Index* idx = h->index[which];
  if (c.valid || c.pending_n == 0 || !idx || (size_t)idx->n != c.pending_n) return NGICP_OK;
  const size_t n = c.pending_n;
  double* d_in = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_in, n * 16, h->stream));
  NGICP_CUDA(h, cudaMemcpyAsync(d_in, c.pending.data(), n * 16 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  dev_free(c.cov6, h->stream);
  c.cov6 = nullptr;
  NGICP_CUDA(h, dev_alloc(&c.cov6, n * 6, h->stream));
  if (int rc = mat4_host_order_to_cov6(h, idx, d_in, c.cov6)) return rc;
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));  // the pageable source buffer must outlive the copy
  dev_free(d_in, h->stream);
  c.n = n; c.valid = true;
  c.pending.clear(); c.pending.shrink_to_fit(); c.pending_n = 0;
  return NGICP_OK;
}

int compute_covariances_impl(Handle* h, int which, float* density) {
  Index* idx = h->index[which];
  if (!idx) return fail(h, NGICP_ERR_INVALID, "calculate covariances: no cloud attached");
  const int k = h->params.k_correspondences;
  if (idx->n_seg == 1 && idx->n < k) return fail(h, NGICP_ERR_INVALID, "calculate covariances: fewer points than k_correspondences");
  for (size_t sgm = 0; sgm + 1 < idx->seg_offsets_host.size(); sgm++)     // batched clouds: neighbours never cross a segment
    if (idx->seg_offsets_host[sgm + 1] - idx->seg_offsets_host[sgm] < k)
      return fail(h, NGICP_ERR_INVALID, "calculate covariances: a segment has fewer points than k_correspondences");
  cudaStream_t s = h->stream;
  const size_t n = idx->n;
  CovSet& c = h->covs[which];
  dev_free(c.cov6, s);
  c.cov6 = nullptr; c.valid = false; c.pending.clear(); c.pending_n = 0;
  int* d_nbr = nullptr; double* d_dens = nullptr; double* d_sum = nullptr;
  NGICP_CUDA(h, dev_alloc(&c.cov6, n * 6, s));
  NGICP_CUDA(h, dev_alloc(&d_nbr, nbr_elems(n, k), s));
  NGICP_CUDA(h, dev_alloc(&d_dens, n, s));
  NGICP_CUDA(h, dev_alloc(&d_sum, (size_t)idx->n_seg, s));
  int rc = NGICP_OK;
  if (which == NGICP_SOURCE) rc = speculate_first_search(h);   // runs beside K2 + K3 on the second stream
  if (!rc) {
    StageTimer t(h, &h->t.knn_ms);
    rc = knn_self(h, idx, k, d_nbr, d_dens);
  }
  if (!rc) {
    // density = sum / N (nano_gicp.cc:389): K3 delivers the sum to host-mapped memory with the covariances. Callers that
    // do not need the value pass NULL and stay asynchronous.
    const bool want = density && idx->n_seg == 1;
    double sum = 0.0;
    {
      StageTimer t(h, &h->t.covariance_ms);
      rc = covariances_from_knn(h, idx, d_nbr, k, h->params.regularization, c.cov6, want ? d_dens : nullptr, want ? &sum : nullptr);
    }
    if (!rc && want) {
      c.density = (float)(sum / (double)n);
      *density = c.density;
    }
  }
  dev_free(d_nbr, s); dev_free(d_dens, s); dev_free(d_sum, s);
  if (rc) return rc;
  c.n = n; c.valid = true;
  return NGICP_OK;
}

}  // namespace
}  // namespace ngicp

using namespace ngicp;

static inline Handle* H(ngicp_handle* p) { return p; }
static inline Index* I(ngicp_index* p) { return p; }
static inline ngicp_index* wrap(Index* i) { return i; }

extern "C" {

const char* ngicp_version(void) { return "ngicp_b200 0.1 (sm_100a)"; }

void ngicp_default_params(ngicp_params* p) {
  if (!p) return;
  p->k_correspondences = 20;           // nano_gicp.cc:60
  p->max_corr_dist = (double)FLT_MAX;  // nano_gicp.cc:62
  p->regularization = NGICP_REG_PLANE; // nano_gicp.cc:64
  p->max_iterations = 64;              // lsq_registration.cc:55
  p->rotation_epsilon = 2e-3;          // :56
  p->transformation_epsilon = 5e-4;    // :57
  p->lm_init_lambda_factor = 1e-9;     // :63
  p->lm_max_iterations = 10;           // :62
  p->use_gauss_newton = 0;             // :59
}

const char* ngicp_last_error(const ngicp_handle* h) { return h ? h->err.c_str() : g_thread_err.c_str(); }

int ngicp_create(int device, ngicp_handle** out) {
  if (!out) return fail(nullptr, NGICP_ERR_INVALID, "ngicp_create: out is NULL");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, NGICP_ERR_NO_DEVICE, "no CUDA device: libngicp_b200 has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(nullptr, NGICP_ERR_INVALID, "ngicp_create: bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, NGICP_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(nullptr, NGICP_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library is built for sm_100a only");
  ngicp_handle* p = new (std::nothrow) ngicp_handle;
  if (!p) return fail(nullptr, NGICP_ERR_INVALID, "out of host memory");
  Handle* h = p;
  h->device = device;
  ngicp_default_params(&h->params);
  std::memset(&h->t, 0, sizeof h->t);
  if (const char* e = std::getenv("NGICP_K4_CMAX")) h->k4_cmax = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("NGICP_K2_CMAX_MULT")) h->k2_cmax_mult = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("NGICP_K2_LPQ")) h->k2_lpq = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K4_LPQ")) h->k4_lpq = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K2_LEAF")) h->k2_leaf = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K2_CAP2_MULT")) h->k2_cap2_mult = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("NGICP_FINE_OCC10")) h->fine_occ10 = std::max(0, std::atoi(e));
  if (const char* e = std::getenv("NGICP_K2_TMA")) h->k2_tma = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K2_CHUNK")) h->k2_chunk = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K4_BALL")) h->k4_ball = std::atoi(e);
  if (const char* e = std::getenv("NGICP_K4_SPEC")) h->k4_spec = std::atoi(e);
  for (int i = 0; i < 36; i++) h->final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;  // setIdentity, lsq_registration.cc:65
#define CREATE_CUDA(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      const int rc = fail(nullptr, NGICP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
      delete p;                                                                               \
      return rc;                                                                              \
    }                                                                                         \
  } while (0)
  CREATE_CUDA(cudaSetDevice(device));
  CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  CREATE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;  // keep freed blocks cached: per-scan allocations become free-list hits
  CREATE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  {
    // Reserve the workspace once per device and process: the first index of a large submap otherwise pays the growth of
    // the pool (tens of ms measured in the odom loop) in the middle of a sequence. NGICP_POOL_RESERVE_MB overrides (0 = off).
    static std::atomic<unsigned> reserved_mask{0};
    const unsigned bit = 1u << (device & 31);
    if (!(reserved_mask.fetch_or(bit) & bit)) {
      size_t mb = 1024;
      if (const char* e = std::getenv("NGICP_POOL_RESERVE_MB")) mb = (size_t)std::max(0, std::atoi(e));
      void* blk = nullptr;
      if (mb && cudaMallocAsync(&blk, mb << 20, h->stream) == cudaSuccess) cudaFreeAsync(blk, h->stream);
      else cudaGetLastError();   // not fatal: the pool simply grows on demand
    }
  }
  CREATE_CUDA(cudaMalloc(&h->partials, sizeof(double) * 32 * kMaxLinBlocks));
  CREATE_CUDA(cudaMalloc(&h->counter, sizeof(unsigned int) * kMaxBatch));
  CREATE_CUDA(cudaMemset(h->counter, 0, sizeof(unsigned int) * kMaxBatch));
  CREATE_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->slot_host), sizeof(ReduceSlot) * kMaxBatch, cudaHostAllocMapped));
  std::memset(h->slot_host, 0, sizeof(ReduceSlot) * kMaxBatch);
  CREATE_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->slot_dev), h->slot_host, 0));
  CREATE_CUDA(cudaEventCreate(&h->ev[0]));
  CREATE_CUDA(cudaEventCreate(&h->ev[1]));
  CREATE_CUDA(cudaEventCreate(&h->ev[2]));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->stage_done, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&h->input_copied, cudaEventDisableTiming));
  if (const char* e = std::getenv("NGICP_ASYNC_INPUT")) h->async_input = std::atoi(e) != 0;
#undef CREATE_CUDA
  *out = p;
  return NGICP_OK;
}

int ngicp_destroy(ngicp_handle* p) {
  if (!p) return NGICP_OK;
  Handle* h = H(p);
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (int w = 0; w < 2; w++) {
    drop_covs(h, w);
    release_index(h, h->index[w]);
    h->index[w] = nullptr;
  }
  if (h->stream2) cudaStreamSynchronize(h->stream2);
  cudaStreamSynchronize(h->stream);
  if (h->corr) cudaFree(h->corr);
  if (h->corr_alt) cudaFree(h->corr_alt);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_spec) cudaEventDestroy(h->ev_spec);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->scan_pts) cudaFree(h->scan_pts);
  if (h->scan_keys) cudaFree(h->scan_keys);
  if (h->heavy) cudaFree(h->heavy);
  if (h->heavy_count) cudaFree(h->heavy_count);
  if (h->k2_ctr) cudaFree(h->k2_ctr);
  if (h->partials) cudaFree(h->partials);
  if (h->batch_partials) cudaFree(h->batch_partials);
  if (h->counter) cudaFree(h->counter);
  if (h->slot_host) cudaFreeHost(h->slot_host);
  if (h->stage_host) cudaFreeHost(h->stage_host);
  if (h->ev[0]) cudaEventDestroy(h->ev[0]);
  if (h->ev[1]) cudaEventDestroy(h->ev[1]);
  if (h->ev[2]) cudaEventDestroy(h->ev[2]);
  if (h->stage_done) cudaEventDestroy(h->stage_done);
  if (h->input_copied) cudaEventDestroy(h->input_copied);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete p;
  return NGICP_OK;
}

int ngicp_set_params(ngicp_handle* p, const ngicp_params* prm) {
  if (!p || !prm) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_set_params: NULL argument");
  if (prm->k_correspondences < 1) return fail(H(p), NGICP_ERR_INVALID, "k_correspondences must be >= 1");
  if (prm->regularization < 0 || prm->regularization > NGICP_REG_FROBENIUS) return fail(H(p), NGICP_ERR_INVALID, "unknown regularization method");
  // a speculative search in flight (api.cu:compute_covariances_impl, ngicp_align) used the old gate: only a change of
  // the parameters the search depends on invalidates it — the drop-in header pushes the whole set before every align
  if (prm->max_corr_dist != H(p)->params.max_corr_dist || prm->k_correspondences != H(p)->params.k_correspondences) {
    if (int rc = use_device(H(p))) return rc;
    drop_speculation(H(p));
  }
  H(p)->params = *prm;
  return NGICP_OK;
}
int ngicp_get_params(const ngicp_handle* p, ngicp_params* prm) {
  if (!p || !prm) return NGICP_ERR_INVALID;
  *prm = p->params;
  return NGICP_OK;
}
int ngicp_set_async_input(ngicp_handle* p, int on) {
  if (!p) return NGICP_ERR_INVALID;
  H(p)->async_input = on != 0;
  return NGICP_OK;
}
void* ngicp_stream(ngicp_handle* p) { return p ? (void*)H(p)->stream : nullptr; }
int ngicp_synchronize(ngicp_handle* p) {
  if (!p) return NGICP_ERR_INVALID;
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));
  return NGICP_OK;
}

// ------------------------------------------------------------------------------------- index
int ngicp_index_build(ngicp_handle* p, const void* points, size_t n, size_t stride_bytes, ngicp_index** out) {
  if (!p || !out) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_index_build: NULL argument");
  Handle* h = H(p);
  *out = nullptr;
  if (!points || n == 0) return fail(h, NGICP_ERR_INVALID, "ngicp_index_build: empty cloud");
  if (int rc = use_device(h)) return rc;
  float* d_xyz = nullptr;
  if (int rc = upload_xyz(h, points, n, stride_bytes, &d_xyz)) return rc;
  Index* idx = nullptr;
  const int rc = build_index(h, d_xyz, 3, (int)n, nullptr, 1, &idx);
  dev_free(d_xyz, h->stream);
  if (rc) return rc;
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));  // a free-standing index may be adopted by another handle / stream
  *out = wrap(idx);
  return NGICP_OK;
}

int ngicp_index_build_device(ngicp_handle* p, const void* d_points_f4, size_t n, ngicp_index** out) {
  if (!p || !out) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_index_build_device: NULL argument");
  Handle* h = H(p);
  *out = nullptr;
  if (!d_points_f4 || n == 0) return fail(h, NGICP_ERR_INVALID, "ngicp_index_build_device: empty cloud");
  if (int rc = use_device(h)) return rc;
  Index* idx = nullptr;
  if (int rc = build_index(h, static_cast<const float*>(d_points_f4), 4, (int)n, nullptr, 1, &idx)) return rc;
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));
  *out = wrap(idx);
  return NGICP_OK;
}

int ngicp_index_retain(ngicp_index* idx) {
  if (!idx) return NGICP_ERR_INVALID;
  I(idx)->refs.fetch_add(1);
  return NGICP_OK;
}
int ngicp_index_release(ngicp_index* idx) {
  if (!idx) return NGICP_OK;
  cudaSetDevice(I(idx)->device);
  release_index(nullptr, I(idx));
  return NGICP_OK;
}
size_t ngicp_index_size(const ngicp_index* idx) { return idx ? (size_t)idx->n : 0; }

int ngicp_knn(ngicp_handle* p, const ngicp_index* idx, const void* queries, size_t nq, size_t stride_bytes, int k, int32_t* out_idx, float* out_sqd) {
  if (!p || !idx || !out_idx || !out_sqd) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_knn: NULL argument");
  Handle* h = H(p);
  if (nq == 0) return NGICP_OK;  // nothing to do
  if (!queries) return fail(h, NGICP_ERR_INVALID, "ngicp_knn: NULL queries");
  if (int rc = use_device(h)) return rc;
  cudaStream_t s = h->stream;
  NGICP_CUDA(h, order_after_build(idx, s));
  float* d_xyz = nullptr;
  if (int rc = upload_xyz(h, queries, nq, stride_bytes, &d_xyz)) return rc;
  float4* d_q = nullptr; int* d_i = nullptr; float* d_d = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_q, nq, s));
  NGICP_CUDA(h, dev_alloc(&d_i, nq * (size_t)k, s));
  NGICP_CUDA(h, dev_alloc(&d_d, nq * (size_t)k, s));
  pack_f4_kernel<<<((int)nq + 255) / 256, 256, 0, s>>>(d_xyz, 3, (int)nq, d_q);
  count_launch(h);
  int rc;
  {
    StageTimer t(h, &h->t.knn_ms);
    rc = knn_queries(h, idx, d_q, (int)nq, k, d_i, d_d);
  }
  if (!rc) {
    NGICP_CUDA(h, cudaMemcpyAsync(out_idx, d_i, sizeof(int) * nq * k, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaMemcpyAsync(out_sqd, d_d, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaStreamSynchronize(s));
  }
  dev_free(d_xyz, s); dev_free(d_q, s); dev_free(d_i, s); dev_free(d_d, s);
  return rc;
}

int ngicp_self_neighbours(ngicp_handle* p, int which, int k, int32_t* out_idx, double* out_density_terms) {
  if (!p || (which != 0 && which != 1) || !out_idx) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_self_neighbours: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  const Index* idx = h->index[which];
  if (!idx) return fail(h, NGICP_ERR_INVALID, "ngicp_self_neighbours: no cloud attached");
  if (idx->n_seg == 1 && idx->n < k) return fail(h, NGICP_ERR_INVALID, "ngicp_self_neighbours: fewer points than k");
  cudaStream_t s = h->stream;
  const size_t n = (size_t)idx->n;
  int *d_nbr = nullptr, *d_out = nullptr;
  double* d_dens = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_nbr, nbr_elems(n, k), s));
  NGICP_CUDA(h, dev_alloc(&d_out, n * (size_t)k, s));
  NGICP_CUDA(h, dev_alloc(&d_dens, n, s));
  int rc = knn_self(h, idx, k, d_nbr, d_dens);
  if (!rc) rc = export_self_rows(h, idx, d_nbr, k, d_out);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out_idx, d_out, sizeof(int) * n * k, cudaMemcpyDeviceToHost, s);
    // density terms are in sorted order on the device; hand them back per ORIGINAL point through the inverse permutation on the host
    std::vector<double> dens;
    std::vector<int> inv;
    if (e == cudaSuccess && out_density_terms) {
      dens.resize(n); inv.resize(n);
      e = cudaMemcpyAsync(dens.data(), d_dens, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(inv.data(), idx->inv, sizeof(int) * n, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(h, NGICP_ERR_CUDA, std::string("ngicp_self_neighbours: ") + cudaGetErrorString(e));
    else if (out_density_terms) for (size_t i = 0; i < n; i++) out_density_terms[i] = dens[inv[i]];
  }
  dev_free(d_nbr, s); dev_free(d_out, s); dev_free(d_dens, s);
  return rc;
}

int ngicp_index_keys(ngicp_handle* p, const ngicp_index* idx, uint64_t* out_keys, float origin_h0[4]) {
  if (!p || !idx) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_index_keys: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  cudaStream_t s = h->stream;
  const Index& ix = *idx;
  NGICP_CUDA(h, order_after_build(idx, s));
  if (out_keys) {
    unsigned long long* d_out = nullptr;
    NGICP_CUDA(h, dev_alloc(&d_out, (size_t)ix.n, s));
    keys_unsort_kernel<<<(ix.n + 255) / 256, 256, 0, s>>>(ix.pts, ix.keys, ix.n, d_out);
    count_launch(h);
    NGICP_CUDA(h, cudaMemcpyAsync(out_keys, d_out, sizeof(uint64_t) * ix.n, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaStreamSynchronize(s));
    dev_free(d_out, s);
  }
  if (origin_h0) {
    float4 o; GridMeta m;
    NGICP_CUDA(h, cudaMemcpyAsync(&o, ix.seg_origin, sizeof o, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaMemcpyAsync(&m, ix.meta, sizeof m, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaStreamSynchronize(s));
    origin_h0[0] = o.x; origin_h0[1] = o.y; origin_h0[2] = o.z; origin_h0[3] = m.h0;
  }
  return NGICP_OK;
}

// ------------------------------------------------------------------------- bookkeeping
int ngicp_attach_index(ngicp_handle* p, int which, ngicp_index* idx) {
  if (!p || (which != 0 && which != 1)) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_attach_index: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  if (idx && idx->device != h->device) return fail(h, NGICP_ERR_INVALID, "index lives on another device");
  Index* old = h->index[which];
  if (old == idx) return NGICP_OK;
  drop_speculation(h);
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));  // our pending work may still read the old index
  NGICP_CUDA(h, order_after_build(idx, h->stream)); // the index may still be under construction on its builder's stream
  CovSet& c = h->covs[which];
  if (c.valid) {
    // The reference keeps covariances across registerInput* / tree hand-over (nano_gicp.cc:119-132),
    // indexed by point number. Device covariances live in the old index's sorted order: re-sort
    // them for the new index when the sizes match, otherwise they cannot refer to this cloud.
    if (idx && old && c.n == (size_t)idx->n) {
      float *d_host_order = nullptr, *d_new = nullptr;
      NGICP_CUDA(h, dev_alloc(&d_host_order, c.n * 6, h->stream));
      NGICP_CUDA(h, dev_alloc(&d_new, c.n * 6, h->stream));
      if (int rc = cov6_to_host_order(h, old, c.cov6, d_host_order)) return rc;
      if (int rc = cov6_from_host_order(h, idx, d_host_order, d_new)) return rc;
      dev_free(d_host_order, h->stream);
      dev_free(c.cov6, h->stream);
      c.cov6 = d_new;
    } else {
      dev_free(c.cov6, h->stream);
      c.cov6 = nullptr; c.valid = false; c.n = 0;
    }
  }
  if (idx) idx->refs.fetch_add(1);
  release_index(h, old);
  h->index[which] = idx;
  h->lin_valid = false;
  return apply_pending(h, which);
}

ngicp_index* ngicp_get_index(ngicp_handle* p, int which) {
  if (!p || (which != 0 && which != 1)) return nullptr;
  return wrap(H(p)->index[which]);
}

extern "C++" {
namespace ngicp {
int swap_in_index(Handle* h, int which, Index* idx) {
  // same stream as everything else this handle does: no synchronisation needed to swap
  drop_speculation(h);
  Index* old = h->index[which];
  h->index[which] = idx;
  if (old) {
    if (old->refs.load() == 1) release_index(h, old);                 // sole owner: stream-ordered free
    else { cudaStreamSynchronize(h->stream); release_index(h, old); } // shared: our reads must be done first
  }
  drop_covs(h, which);  // nano_gicp.cc:146,160
  h->lin_valid = false;
  return NGICP_OK;
}
int select_device(Handle* h) { return use_device(h); }
int upload_points(Handle* h, const void* points, size_t n, size_t stride_bytes, float** d_xyz, int* stride_floats) {
  return upload_xyz(h, points, n, stride_bytes, d_xyz, stride_floats);
}
int finish_input_copy(Handle* h) { return finish_input(h); }
}  // namespace ngicp
}  // extern "C++"

int ngicp_set_input(ngicp_handle* p, int which, const void* points, size_t n, size_t stride_bytes) {
  if (!p || (which != 0 && which != 1)) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_set_input: bad argument");
  Handle* h = H(p);
  if (!points || n == 0) return fail(h, NGICP_ERR_INVALID, "ngicp_set_input: empty cloud");
  if (int rc = use_device(h)) return rc;
  float* d_xyz = nullptr;
  int stride = 3;
  if (int rc = upload_xyz(h, points, n, stride_bytes, &d_xyz, &stride)) return rc;
  Index* idx = nullptr;
  int rc = build_index(h, d_xyz, stride, (int)n, nullptr, 1, &idx, which == NGICP_SOURCE || n <= kFineTargetMax);
  dev_free(d_xyz, h->stream);
  if (!rc) rc = swap_in_index(h, which, idx);
  const int rc2 = finish_input(h);     // the caller's page-locked buffer is free again when this returns (see ngicp_set_async_input)
  return rc ? rc : rc2;
}

int ngicp_set_input_device(ngicp_handle* p, int which, const void* d_points_f4, size_t n) {
  if (!p || (which != 0 && which != 1)) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_set_input_device: bad argument");
  Handle* h = H(p);
  if (!d_points_f4 || n == 0) return fail(h, NGICP_ERR_INVALID, "ngicp_set_input_device: empty cloud");
  if (int rc = use_device(h)) return rc;
  Index* idx = nullptr;
  if (int rc = build_index(h, static_cast<const float*>(d_points_f4), 4, (int)n, nullptr, 1, &idx, which == NGICP_SOURCE || n <= kFineTargetMax)) return rc;
  return swap_in_index(h, which, idx);
}

int ngicp_swap_source_and_target(ngicp_handle* p) {
  if (!p) return NGICP_ERR_INVALID;
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  drop_speculation(h);
  std::swap(h->index[0], h->index[1]);
  std::swap(h->covs[0], h->covs[1]);
  h->lin_valid = false;  // correspondences_.clear(), nano_gicp.cc:102-103
  return NGICP_OK;
}

int ngicp_clear(ngicp_handle* p, int which) {
  if (!p || (which != 0 && which != 1)) return NGICP_ERR_INVALID;
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  drop_speculation(h);
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));
  release_index(h, h->index[which]);
  h->index[which] = nullptr;
  drop_covs(h, which);
  h->lin_valid = false;
  return NGICP_OK;
}

// ------------------------------------------------------------------------- covariances
int ngicp_compute_covariances(ngicp_handle* p, int which, float* density) {
  if (!p || (which != 0 && which != 1)) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_compute_covariances: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  return compute_covariances_impl(h, which, density);
}

int ngicp_has_covariances(const ngicp_handle* p, int which, size_t* n) {
  if (!p || (which != 0 && which != 1)) return 0;
  const CovSet& c = p->covs[which];
  if (n) *n = c.valid ? c.n : c.pending_n;
  return (c.valid || c.pending_n) ? 1 : 0;
}

int ngicp_get_covariances(ngicp_handle* p, int which, double* out, size_t n) {
  if (!p || (which != 0 && which != 1) || !out) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_get_covariances: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  CovSet& c = h->covs[which];
  if (!c.valid && c.pending_n == n && n) {  // handed over by the host and never needed on the device
    std::memcpy(out, c.pending.data(), n * 16 * sizeof(double));
    return NGICP_OK;
  }
  if (!c.valid || !h->index[which]) return fail(h, NGICP_ERR_INVALID, "no covariances");
  if (c.n != n) return fail(h, NGICP_ERR_INVALID, "ngicp_get_covariances: size mismatch");
  double* d_out = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_out, n * 16, h->stream));
  if (int rc = cov6_to_mat4_host_order(h, h->index[which], c.cov6, d_out)) return rc;
  NGICP_CUDA(h, cudaMemcpyAsync(out, d_out, n * 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));
  dev_free(d_out, h->stream);
  return NGICP_OK;
}

int ngicp_set_covariances(ngicp_handle* p, int which, const double* in, size_t n) {
  if (!p || (which != 0 && which != 1) || !in || n == 0) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_set_covariances: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  CovSet& c = h->covs[which];
  dev_free(c.cov6, h->stream);
  c.cov6 = nullptr; c.valid = false; c.n = 0;
  c.pending.assign(in, in + n * 16);
  c.pending_n = n;
  h->lin_valid = false;
  return apply_pending(h, which);
}

// ------------------------------------------------------------------------- registration
int ngicp_update_correspondences(ngicp_handle* p, const double T[16], int32_t* corr, float* sqd, double* mahal, int* ncorr) {
  if (!p || !T) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_update_correspondences: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  return export_correspondences(h, T, corr, sqd, mahal, ncorr);
}

int ngicp_linearize(ngicp_handle* p, const double T[16], double Hm[36], double b[6], double* error, int* ncorr) {
  if (!p || !T) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_linearize: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  return linearize_device(h, T, Hm && b, Hm, b, error, ncorr);
}

int ngicp_compute_error(ngicp_handle* p, const double T[16], double* error) {
  if (!p || !T) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_compute_error: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  return compute_error_device(h, T, error);
}

int ngicp_align(ngicp_handle* p, const float guess[16], float T_out[16], int* nr_iterations, int* converged, double H_final[36], double* final_error) {
  if (!p) return fail(nullptr, NGICP_ERR_INVALID, "ngicp_align: NULL handle");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  if (!h->index[0] || !h->index[1]) return fail(h, NGICP_ERR_INVALID, "align: source and target clouds are required");
  // NanoGICP::computeTransformation (nano_gicp.cc:194-203): lazily (re)compute missing covariances
  for (int w = 0; w < 2; w++) {
    if (int rc = apply_pending(h, w)) return rc;
    if (!h->covs[w].valid || h->covs[w].n != (size_t)h->index[w]->n)
      if (int rc = compute_covariances_impl(h, w, nullptr)) return rc;
  }
  const ngicp_params& prm = h->params;
  // LsqRegistration::computeTransformation (lsq_registration.cc:108-134)
  lm::Iso x0 = guess ? lm::from_colmajor_f(guess) : lm::identity();
  h->lm_lambda = -1.0;
  bool conv = false;
  int nr = 0;
  int status = NGICP_OK;
  for (int i = 0; i < prm.max_iterations && !conv; i++) {
    nr = i;
    lm::Iso delta = lm::identity();
    double Hm[36], b[6], y0 = 0.0, T[16];
    lm::to_colmajor(x0, T);
    if (int rc = linearize_device(h, T, true, Hm, b, &y0, nullptr)) return rc;
    bool step_ok = false;
    if (prm.use_gauss_newton) {  // step_gn, lsq_registration.cc:161-178
      double nb[6], d[6];
      for (int j = 0; j < 6; j++) nb[j] = -b[j];
      lm::solve6(Hm, nb, d);
      delta = lm::delta_from(d);
      x0 = lm::compose(delta, x0);
      std::memcpy(h->final_hessian, Hm, sizeof Hm);
      h->final_error = y0;
      step_ok = true;
    } else {  // step_lm, lsq_registration.cc:181-229
      if (h->lm_lambda < 0.0) {
        double mx = 0.0;
        for (int j = 0; j < 6; j++) mx = std::max(mx, std::fabs(Hm[7 * j]));
        h->lm_lambda = prm.lm_init_lambda_factor * mx;
      }
      double nu = 2.0;
      for (int it = 0; it < prm.lm_max_iterations; it++) {
        double A[36], nb[6], d[6];
        std::memcpy(A, Hm, sizeof A);
        for (int j = 0; j < 6; j++) { A[7 * j] += h->lm_lambda; nb[j] = -b[j]; }
        lm::solve6(A, nb, d);
        delta = lm::delta_from(d);
        const lm::Iso xi = lm::compose(delta, x0);
        double yi = 0.0, Ti[16];
        lm::to_colmajor(xi, Ti);
        // if this trial is accepted and the loop goes on, the next linearize needs the correspondences at xi: search
        // them on the second stream while K5 runs
        if (i + 1 < prm.max_iterations && !lm::is_converged(delta, prm.rotation_epsilon, prm.transformation_epsilon))
          if (int rc = speculate_search(h, Ti)) return rc;
        if (int rc = compute_error_device(h, Ti, &yi)) return rc;
        double den = 0.0;
        for (int j = 0; j < 6; j++) den += d[j] * (h->lm_lambda * d[j] - b[j]);
        const double rho = (y0 - yi) / den;
        if (rho < 0) {
          if (lm::is_converged(delta, prm.rotation_epsilon, prm.transformation_epsilon)) { step_ok = true; break; }
          h->lm_lambda = nu * h->lm_lambda;
          nu = 2 * nu;
          continue;
        }
        x0 = xi;
        h->lm_lambda = h->lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
        std::memcpy(h->final_hessian, Hm, sizeof Hm);
        h->final_error = yi;
        step_ok = true;
        break;
      }
    }
    if (!step_ok) { status = NGICP_ERR_LM_NOT_CONVERGED; fail(h, status, "lm not converged!!"); break; }
    conv = lm::is_converged(delta, prm.rotation_epsilon, prm.transformation_epsilon);
  }
  drop_speculation(h);
  if (T_out) {
    std::memset(T_out, 0, 16 * sizeof(float));
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) T_out[4 * c + r] = (float)x0.R[3 * r + c];
      T_out[12 + r] = (float)x0.t[r];
    }
    T_out[15] = 1.0f;
  }
  if (nr_iterations) *nr_iterations = nr;
  if (converged) *converged = conv ? 1 : 0;
  if (H_final) std::memcpy(H_final, h->final_hessian, sizeof h->final_hessian);
  if (final_error) *final_error = h->final_error;
  return status;
}

int ngicp_transform_source(ngicp_handle* p, const float T[16], void* out_points, size_t n, size_t stride_bytes) {
  if (!p || !T || !out_points) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_transform_source: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  Index* idx = h->index[0];
  if (!idx || (size_t)idx->n != n) return fail(h, NGICP_ERR_INVALID, "ngicp_transform_source: source cloud missing or size mismatch");
  if (stride_bytes < 12 || stride_bytes % 4) return fail(h, NGICP_ERR_INVALID, "bad stride");
  cudaStream_t s = h->stream;
  float *d_in = nullptr, *d_out = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_in, n * 3, s));
  NGICP_CUDA(h, dev_alloc(&d_out, n * 3, s));
  unsort_points_kernel<<<((int)n + 255) / 256, 256, 0, s>>>(idx->pts, (int)n, d_in);
  count_launch(h);
  if (int rc = transform_points_device(h, d_in, 3, (int)n, T, d_out)) return rc;
  if (int rc = ensure_stage(h, n * 12)) return rc;
  NGICP_CUDA(h, cudaMemcpyAsync(h->stage_host, d_out, n * 12, cudaMemcpyDeviceToHost, s));
  NGICP_CUDA(h, cudaStreamSynchronize(s));
  const float* st = static_cast<const float*>(h->stage_host);
  char* dst = static_cast<char*>(out_points);
  for (size_t i = 0; i < n; i++) {
    float* o = reinterpret_cast<float*>(dst + i * stride_bytes);
    o[0] = st[3 * i]; o[1] = st[3 * i + 1]; o[2] = st[3 * i + 2];
  }
  dev_free(d_in, s); dev_free(d_out, s);
  return NGICP_OK;
}

// ------------------------------------------------------------------------- batched units
int ngicp_batch_covariances(ngicp_handle* p, const void* points, size_t n, size_t stride_bytes, const int64_t* seg_offsets, int n_seg,
                            double* out_4x4, float* out_cov6, float* seg_density) {
  if (!p || !points || !seg_offsets || n == 0 || n_seg < 1) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_batch_covariances: bad argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  const int k = h->params.k_correspondences;
  for (int s = 0; s < n_seg; s++)
    if (seg_offsets[s + 1] - seg_offsets[s] < k) return fail(h, NGICP_ERR_INVALID, "ngicp_batch_covariances: a keyframe has fewer points than k_correspondences");
  cudaStream_t st = h->stream;
  float* d_xyz = nullptr;
  if (int rc = upload_xyz(h, points, n, stride_bytes, &d_xyz)) return rc;
  Index* idx = nullptr;
  int rc = build_index(h, d_xyz, 3, (int)n, seg_offsets, n_seg, &idx);
  dev_free(d_xyz, st);
  if (rc) return rc;
  int* d_nbr = nullptr; double* d_dens = nullptr; double* d_sum = nullptr; float* d_cov = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_nbr, nbr_elems(n, k), st));
  NGICP_CUDA(h, dev_alloc(&d_dens, n, st));
  NGICP_CUDA(h, dev_alloc(&d_sum, (size_t)n_seg, st));
  NGICP_CUDA(h, dev_alloc(&d_cov, n * 6, st));
  {
    StageTimer t(h, &h->t.knn_ms);
    rc = knn_self(h, idx, k, d_nbr, d_dens);
  }
  if (!rc) {
    StageTimer t(h, &h->t.covariance_ms);
    rc = covariances_from_knn(h, idx, d_nbr, k, h->params.regularization, d_cov);
  }
  if (!rc && out_4x4) {
    double* d_out = nullptr;
    NGICP_CUDA(h, dev_alloc(&d_out, n * 16, st));
    rc = cov6_to_mat4_host_order(h, idx, d_cov, d_out);
    if (!rc) NGICP_CUDA(h, cudaMemcpyAsync(out_4x4, d_out, n * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NGICP_CUDA(h, cudaStreamSynchronize(st));
    dev_free(d_out, st);
  }
  if (!rc && out_cov6) {
    float* d_out = nullptr;
    NGICP_CUDA(h, dev_alloc(&d_out, n * 6, st));
    rc = cov6_to_host_order(h, idx, d_cov, d_out);
    if (!rc) NGICP_CUDA(h, cudaMemcpyAsync(out_cov6, d_out, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    NGICP_CUDA(h, cudaStreamSynchronize(st));
    dev_free(d_out, st);
  }
  if (!rc && seg_density) {
    std::vector<double> sums(n_seg);
    rc = reduce_sum(h, d_dens, (int)n, idx->seg_start, n_seg, sums.data());
    for (int s = 0; !rc && s < n_seg; s++) seg_density[s] = (float)(sums[s] / (double)(seg_offsets[s + 1] - seg_offsets[s]));
  }
  NGICP_CUDA(h, cudaStreamSynchronize(st));
  dev_free(d_nbr, st); dev_free(d_dens, st); dev_free(d_sum, st); dev_free(d_cov, st);
  release_index(h, idx);
  return rc;
}

int ngicp_set_input_batch(ngicp_handle* p, int which, const void* points, size_t n, size_t stride_bytes, const int64_t* seg_offsets, int n_seg) {
  if (!p || (which != 0 && which != 1) || !seg_offsets) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_set_input_batch: bad argument");
  Handle* h = H(p);
  if (!points || n == 0 || n_seg < 1) return fail(h, NGICP_ERR_INVALID, "ngicp_set_input_batch: empty cloud");
  if (int rc = use_device(h)) return rc;
  float* d_xyz = nullptr;
  if (int rc = upload_xyz(h, points, n, stride_bytes, &d_xyz)) return rc;
  Index* idx = nullptr;
  const int rc = build_index(h, d_xyz, 3, (int)n, seg_offsets, n_seg, &idx);
  dev_free(d_xyz, h->stream);
  if (rc) return rc;
  return swap_in_index(h, which, idx);
}

int ngicp_batch_linearize(ngicp_handle* p, int n_scans, const double* T16s, double* H36s, double* b6s, double* errs, int* ncorrs) {
  if (!p || !T16s) return fail(p ? H(p) : nullptr, NGICP_ERR_INVALID, "ngicp_batch_linearize: NULL argument");
  Handle* h = H(p);
  if (int rc = use_device(h)) return rc;
  return batch_linearize_device(h, n_scans, T16s, H36s, b6s, errs, ncorrs);
}

// ------------------------------------------------------------------------- timing hooks
int ngicp_enable_timing(ngicp_handle* p, int on) {
  if (!p) return NGICP_ERR_INVALID;
  H(p)->timing = on != 0;
  return NGICP_OK;
}
int ngicp_get_timings(ngicp_handle* p, ngicp_timings* out, int reset) {
  if (!p) return NGICP_ERR_INVALID;
  if (out) *out = H(p)->t;
  if (reset) std::memset(&H(p)->t, 0, sizeof(ngicp_timings));
  return NGICP_OK;
}

}  // extern "C"
