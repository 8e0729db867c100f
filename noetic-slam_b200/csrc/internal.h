// Host-side objects behind the C ABI (include/ngicp_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/ngicp_b200.h"
#include "common.cuh"

namespace ngicp {
struct ReduceSlot;  // host-mapped result slot (linearize.cuh)

// Covariances of one side (source / target), device-resident in the index's sorted order.
struct CovSet {
  float* cov6 = nullptr;     // [n][6] xx,xy,xz,yy,yz,zz
  size_t n = 0;
  bool valid = false;        // sorted copy present
  std::vector<double> pending;  // [n][16] host-order Matrix4d list handed over before a matching index was attached
  size_t pending_n = 0;
  float density = 0.f;
};
}  // namespace ngicp

// Device cloud + voxel index. Replaces nanoflann::KdTreeFLANN<PointT>
// (reference src/dlio/include/nano_gicp/nanoflann_adaptor.h:57-152). Reference-counted because
// DLIO hands the submap tree from one NanoGICP instance to another (odom.cc:1737-1738 -> :995).
struct ngicp_index {
  std::atomic<int> refs{1};
  int device = 0;
  int n = 0;
  int n_seg = 1;
  char* arena = nullptr;                  // one allocation holding every array below
  float4* pts = nullptr;                  // [n] Morton-sorted, .w = original index
  int* inv = nullptr;                     // [n] original index -> sorted position
  unsigned long long* keys = nullptr;     // [n] sorted voxel keys
  ngicp::CellSlot* table = nullptr;
  uint32_t table_mask = 0;
  ngicp::GridMeta* meta = nullptr;
  float4* seg_origin = nullptr;           // [n_seg]
  int* seg_start = nullptr;               // [n_seg+1]
  std::vector<int64_t> seg_offsets_host;  // [n_seg+1]
  // recorded at the end of the build on the building handle's stream: any OTHER stream that touches the index
  // (a handle that adopts or queries it) waits on it first — DLIO builds the submap tree on one NanoGICP object and
  // hands it to another (odom.cc:1737-1738 -> :995)
  cudaEvent_t built = nullptr;
  cudaStream_t built_stream = nullptr;
  ngicp::GridView view() const {
    ngicp::GridView g;
    g.pts = pts; g.inv = inv; g.table = table; g.meta = meta; g.seg_origin = seg_origin; g.seg_start = seg_start;
    g.table_mask = table_mask; g.n = n; g.n_seg = n_seg;
    return g;
  }
};

struct ngicp_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  ngicp_params params;
  ngicp_index* index[2] = {nullptr, nullptr};
  ngicp::CovSet covs[2];
  // registration scratch
  int* corr = nullptr;          // [n_src] target sorted position or -1 (order: source sorted position)
  size_t corr_cap = 0;
  // speculative correspondence search (linearize.cu:speculate_search): second stream, alternate result buffer
  int* corr_alt = nullptr;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_main = nullptr, ev_spec = nullptr;
  bool spec_pending = false;
  bool spec_first = false;      // the pending search is the first one of an align (no hints, no previous linearize)
  double spec_T[16];
  int k4_spec = 1;              // NGICP_K4_SPEC=0 disables
  void* heavy = nullptr;        // [corr_cap] HeavyQuery (linearize.cu): the queries the fast search kernel deferred
  unsigned int* heavy_count = nullptr;   // [2] list lengths, alternating between searches (the heavy kernel clears the next one)
  unsigned int heavy_parity = 0;
  size_t corr_n = 0;            // number of source points the cached correspondences belong to
  // scan held between ngicp_scan_ingest and ngicp_scan_deskew (filters.cu): time-sorted, cropped, on the device
  float4* scan_pts = nullptr;               // [scan_n] xyz1 in ascending time-stamp order
  unsigned long long* scan_keys = nullptr;  // [scan_n] order-preserving integer image of the time stamps
  size_t scan_n = 0, scan_groups = 0;
  double lin_pose[12];          // pose of the last linearize (R row-major 9 + t 3), needed by compute_error
  bool lin_valid = false;
  double* partials = nullptr;   // [max_blocks][32] block partial sums
  unsigned int* counter = nullptr;
  ngicp::ReduceSlot* slot_host = nullptr;   // pinned + mapped
  ngicp::ReduceSlot* slot_dev = nullptr;    // device alias of slot_host
  unsigned long long seq = 0;
  // host staging (pinned), grows on demand
  void* stage_host = nullptr;
  size_t stage_cap = 0;
  cudaEvent_t stage_done = nullptr;  // last H2D out of stage_host
  // page-locked caller clouds are copied straight out of the caller's buffer (api.cu:upload_xyz): by default the entry
  // point returns only after that copy has landed; ngicp_set_async_input(h, 1) lets it return while the copy is in flight
  bool async_input = false;
  bool input_pending = false;
  cudaEvent_t input_copied = nullptr;
  // host work the odom loop (odom_loop.cu) wants done while the device is busy: run once by the next wait_count (filters.cu)
  void (*overlap_fn)(void*) = nullptr;
  void* overlap_arg = nullptr;
  int ingest_stamp_bits = -1;   // uint32 stamps of the next ngicp_scan_ingest are all < 2^bits (the caller looked); -1: unknown
  double* batch_partials = nullptr;  // batched reductions (allocated on first use)
  size_t batch_partials_cap = 0;
  // tuning knobs (env NGICP_K4_CMAX / NGICP_K2_CMAX_MULT override)
  int k4_cmax = 64;
  int k2_cmax_mult = 6;
  int fine_occ10 = 11;    // fine index levels are added while their mean occupancy stays >= fine_occ10 / 10 (NGICP_FINE_OCC10; 0: only the table limits)
  int k2_tma = 1;         // candidate staging of the leaf search: TMA bulk copies (1) or direct loads (0) (NGICP_K2_TMA)
  int k2_cap2_mult = 8;   // leaf rule of the leaf search: the parent of a leaf holds at most cap2 = k2_cap2_mult * cmax points (knn.cu)
  int k2_lpq = 0;   // lanes per query in K2 (0 = pick by cloud size)
  int k2_leaf = 1;  // leaf-scheduled K2 (lknn.cuh); NGICP_K2_LEAF=0 = the warp-cooperative search of round 1
  int k2_chunk = 128;              // staging chunk of the leaf search (128 or 256)
  unsigned int* k2_ctr = nullptr;  // [4] item count / next item / finished warps of the leaf search (zero between calls)
  int k4_lpq = 0;
  int k4_ball = 1;  // seed re-association with the previous correspondences (NGICP_K4_BALL=0 disables)
  // LM state (lsq_registration.h:151-168)
  double lm_lambda = -1.0;
  double final_hessian[36];
  double final_error = 0.0;
  int num_correspondences = 0;
  // timing
  bool timing = false;
  ngicp_timings t;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  std::string err;
};

namespace ngicp {
using Index = ::ngicp_index;
using Handle = ::ngicp_handle;

// ---- error plumbing -------------------------------------------------------------------------
void set_thread_error(const std::string& s);
int fail(Handle* h, int code, const std::string& msg);
#define NGICP_CUDA(h, expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::ngicp::fail((h), NGICP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ---- stage entry points (each in its own .cu) --------------------------------------------------
// K1. xyz: device array of n points, `stride_floats` (3 or 4) floats apart.
// fine: also insert the levels below the base level the table has room for (lknn.cuh needs them where a scan is dense).
// Clouds that are only ever searched INTO (the submap target: covariances come from its keyframes) are built without.
int build_index(Handle* h, const float* d_xyz, int stride_floats, int n, const int64_t* seg_offsets, int n_seg, Index** out, bool fine = true);
void free_index(Index* idx, cudaStream_t stream);
// make `stream` wait for the build of idx unless it IS the building stream (no-op then)
inline cudaError_t order_after_build(const Index* idx, cudaStream_t stream) {
  if (!idx || !idx->built || idx->built_stream == stream) return cudaSuccess;
  return cudaStreamWaitEvent(stream, idx->built, 0);
}
// K2. self k-NN of every indexed point, as sorted positions. Layout of the table (transient between K2 and K3):
// k = 16 / 20 (the compile-time paths of K3): TILED — points in tiles of 32, chunk c (int4 = neighbours 4c..4c+3) of
// the 32 points of a tile contiguous: int4 index ((j / 32) * (k / 4) + c) * 32 + j % 32, so that both K2's writes and
// K3's reads are contiguous across a warp; allocate nbr_elems(n, k) ints. Any other k: row-major nbr[n][k].
inline bool nbr_tiled(int k) { return k == 16 || k == 20; }
inline size_t nbr_elems(size_t n, int k) { return ((n + 31) / 32) * 32 * (size_t)k; }
// dens_term[n] (optional) = sum_{j>=1} d2_j / normalization, the per-point density term.
int knn_self(Handle* h, const Index* idx, int k, int* d_nbr, double* d_dens_term);
int export_self_rows(Handle* h, const Index* idx, const int* d_nbr, int k, int* d_out);   // knn.cu: K2's table as original-index rows
// public k-NN: queries on device (float4), results in ORIGINAL indices, canonical order
int knn_queries(Handle* h, const Index* idx, const float4* d_q, int nq, int k, int* d_out_idx, float* d_out_sqd);
// K3. covariance + regularisation from the k-NN table
// d_dens_term + density_sum (both optional, single-segment clouds): also delivers sum(d_dens_term) to the host (blocking)
int covariances_from_knn(Handle* h, const Index* idx, const int* d_nbr, int k, int reg, float* d_cov6, const double* d_dens_term = nullptr,
                         double* density_sum = nullptr);
int wait_slot(Handle* h, int n_slots, unsigned long long seq);   // linearize.cu: spin on the host-mapped result slots
// sum of n doubles -> *d_out (device), deterministic
// deterministic per-segment sums of a device array, delivered to host memory (linearize.cu; blocks until they arrive)
int reduce_sum(Handle* h, const double* d_in, int n, const int* seg_start_dev, int n_seg, double* host_out);
// covariance layout conversions (host order <-> sorted order)
int cov6_to_mat4_host_order(Handle* h, const Index* idx, const float* d_cov6, double* d_out16);
int mat4_host_order_to_cov6(Handle* h, const Index* idx, const double* d_in16, float* d_cov6);
int cov6_to_host_order(Handle* h, const Index* idx, const float* d_cov6, float* d_out6);
int cov6_from_host_order(Handle* h, const Index* idx, const float* d_in6, float* d_cov6);
// K4 / K5
int linearize_device(Handle* h, const double T[16], bool want_Hb, double H[36], double b[6], double* err, int* ncorr);
int compute_error_device(Handle* h, const double T[16], double* err);
int batch_linearize_device(Handle* h, int n_scans, const double* T16s, double* H36s, double* b6s, double* errs, int* ncorrs);
int export_correspondences(Handle* h, const double T[16], int32_t* corr, float* sqd, double* mahal, int* ncorr);
int speculate_search(Handle* h, const double T[16]);
int speculate_first_search(Handle* h);
void drop_speculation(Handle* h);
int transform_points_device(Handle* h, const float* d_xyz_in, int stride_floats, int n, const float T[16], float* d_xyz_out);

// bookkeeping shared with keyframe.cu
int swap_in_index(Handle* h, int which, Index* idx);   // make idx the source/target cloud, drop that side's covariances
int select_device(Handle* h);

// workspace helpers (stream-ordered pool)
template <typename T>
inline cudaError_t dev_alloc(T** p, size_t count, cudaStream_t s) {
  return cudaMallocAsync(reinterpret_cast<void**>(p), count ? count * sizeof(T) : sizeof(T), s);
}
template <typename T>
inline void dev_free(T* p, cudaStream_t s) {
  if (p) cudaFreeAsync(p, s);
}

inline void count_launch(Handle* h, int k = 1) { h->t.kernel_launches += k; }

struct StageTimer {
  Handle* h; float* acc;
  StageTimer(Handle* h_, float* acc_) : h(h_), acc(acc_) { if (h->timing) cudaEventRecord(h->ev[0], h->stream); }
  ~StageTimer() {
    if (!h->timing) return;
    cudaEventRecord(h->ev[1], h->stream);
    cudaEventSynchronize(h->ev[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    *acc += ms;
  }
};

}  // namespace ngicp
