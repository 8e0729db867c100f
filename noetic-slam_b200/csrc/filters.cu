// Scan pre-filters on the device (SURVEY.md §8f "next" row 2): the two PCL filters DLIO applies right before
// setInputSource — pcl::CropBox (reference src/dlio/src/dlio/odom.cc:114-116 configure, :500-502 apply) and
// pcl::VoxelGrid (odom.cc:118 configure, :575-584 apply). PCL (>= 1.10.0, apt libpcl-dev) is a third-party dependency
// that is not vendored in the reference; the algorithm restated here is PCL 1.10's
// filters/impl/crop_box.hpp (applyFilter) and filters/impl/voxel_grid.hpp (applyFilter, downsample_all_data_ = true,
// min_points_per_voxel_ = 0), xyz only:
//   CropBox   keep a finite point iff (inside box) != negative, inside = !(p < min || p > max) on any axis; order kept.
//   VoxelGrid min/max of the finite points -> min_b = floor(min * inv_leaf), div_b = max_b - min_b + 1 ->
//             idx = ijk0 + ijk1 * div_b0 + ijk2 * div_b0 * div_b1 with ijk = int(floor(p * inv_leaf) - float(min_b))
//             (all fp32) -> sort by idx -> one output per occupied voxel, in ascending idx order, = fp32 sum of its
//             points / count. If the voxel count overflows int32 PCL warns and returns the input unchanged; so do we.
// PCL sorts with std::sort, whose order inside a voxel is unspecified, so its fp32 sums are only defined up to
// rounding; this implementation (and the oracle) fix the order to ascending input index (stable sort).
// Same machinery as K1: fused key + digit-histogram kernel, the hand-written stable LSD radix sort, then one pass over
// the sorted keys. Everything stays in HBM; only the output count comes back (host-mapped slot).
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"
#include "linearize.cuh"
#include "radix_sort.cuh"
#include "cluster_sort.cuh"

namespace ngicp {

int upload_points(Handle* h, const void* points, size_t n, size_t stride_bytes, float** d_xyz, int* stride_floats);   // api.cu
int finish_input_copy(Handle* h);                                                                                     // api.cu

namespace {

constexpr int kFiltThreads = 1024;
constexpr unsigned long long kInvalidVoxel = 0xffffffffull;   // sorts behind every real voxel index (< 2^31)

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }
__device__ __forceinline__ bool finite3(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

struct Box {
  float mn[3], mx[3];
  int negative;
};
__device__ __forceinline__ bool crop_keeps(const Box& b, float x, float y, float z) {
  if (!finite3(x, y, z)) return false;                                          // crop_box.hpp: non-finite points are skipped
  const bool outside = (x < b.mn[0] || y < b.mn[1] || z < b.mn[2]) || (x > b.mx[0] || y > b.mx[1] || z > b.mx[2]);
  return outside ? (b.negative != 0) : (b.negative == 0);
}

// stable in-block rank of the threads whose flag is set; returns the block total through *total
__device__ __forceinline__ unsigned int block_rank(bool flag, unsigned int* total) {
  __shared__ unsigned int wsum[kFiltThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) wsum[warp] = __popc(m);
  __syncthreads();
  unsigned int before = 0, all = 0;
  for (int w = 0; w < kFiltThreads / 32; w++) {
    const unsigned int c = wsum[w];
    if (w < warp) before += c;
    all += c;
  }
  __syncthreads();
  *total = all;
  return before + __popc(m & ((1u << lane) - 1u));
}

// pass 1 of a stable compaction: kept points per block
__global__ void __launch_bounds__(kFiltThreads) crop_count_kernel(const float* __restrict__ xyz, int stride, int n, Box box, unsigned int* __restrict__ block_sums) {
  const int i = blockIdx.x * kFiltThreads + threadIdx.x;
  bool keep = false;
  if (i < n) keep = crop_keeps(box, xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]);
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned int)c;
}

// exclusive scan of the block sums by one block; the grand total goes to the host-mapped slot
__global__ void __launch_bounds__(1024) scan_sums_kernel(unsigned int* __restrict__ sums, int nblocks, unsigned int* __restrict__ total_dev,
                                                         ReduceSlot* __restrict__ slot, unsigned long long seq, const int* __restrict__ flag = nullptr) {
  __shared__ unsigned int s[1024];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned int v = i < nblocks ? sums[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const unsigned int t = threadIdx.x >= off ? s[threadIdx.x - off] : 0u;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nblocks) sums[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total_dev = carry;
    slot->v[0] = (double)carry;
    slot->v[1] = flag ? (double)*flag : 0.0;     // rides along with the count (the voxel grid's overflow flag)
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&slot->seq) = seq;
  }
}

__global__ void __launch_bounds__(kFiltThreads) crop_scatter_kernel(const float* __restrict__ xyz, int stride, int n, Box box,
                                                                    const unsigned int* __restrict__ block_offs, float4* __restrict__ out) {
  const int i = blockIdx.x * kFiltThreads + threadIdx.x;
  bool keep = false;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < n) {
    x = xyz[(size_t)i * stride]; y = xyz[(size_t)i * stride + 1]; z = xyz[(size_t)i * stride + 2];
    keep = crop_keeps(box, x, y, z);
  }
  unsigned int total;
  const unsigned int r = block_rank(keep, &total);
  if (keep) out[block_offs[blockIdx.x] + r] = make_float4(x, y, z, 1.0f);
}

// ---- VoxelGrid ----
struct VoxelMeta {
  unsigned int lo[3], hi[3];   // ordered-int min / max of the finite points
  int min_b[3], div_b[3];
  int overflow;                // voxel count does not fit int32: PCL returns the input unchanged
  int pad;
};

__global__ void __launch_bounds__(256) vg_init_kernel(VoxelMeta* m, uint32_t* digit_hist, int n_hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3) { m->lo[i] = 0xffffffffu; m->hi[i] = 0u; }
  if (i == 0) m->overflow = 0;
  if (i < n_hist) digit_hist[i] = 0u;
}

__global__ void __launch_bounds__(256) vg_minmax_kernel(const float* __restrict__ xyz, int stride, int n, VoxelMeta* __restrict__ m) {
  unsigned int l[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, u[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    if (!finite3(x, y, z)) continue;                                            // getMinMax3D on a non-dense cloud
    const unsigned int o[3] = {f2ord(x), f2ord(y), f2ord(z)};
#pragma unroll
    for (int a = 0; a < 3; a++) { l[a] = min(l[a], o[a]); u[a] = max(u[a], o[a]); }
  }
  __shared__ unsigned int sl[8][3], su[8][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      l[a] = min(l[a], __shfl_xor_sync(0xffffffffu, l[a], off));
      u[a] = max(u[a], __shfl_xor_sync(0xffffffffu, u[a], off));
    }
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5][a] = l[a]; su[threadIdx.x >> 5][a] = u[a]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned int ml = sl[0][threadIdx.x], mu = su[0][threadIdx.x];
    for (int w = 1; w < 8; w++) { ml = min(ml, sl[w][threadIdx.x]); mu = max(mu, su[w][threadIdx.x]); }
    atomicMin(&m->lo[threadIdx.x], ml);
    atomicMax(&m->hi[threadIdx.x], mu);
  }
}

// voxel index of every point (voxel_grid.hpp: "First pass: go over all points and insert them into the index_vector")
// + the radix sort's digit totals
__global__ void __launch_bounds__(256) vg_keys_kernel(const float* __restrict__ xyz, int stride, int n, float3 inv_leaf, VoxelMeta* __restrict__ m,
                                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ digit_hist) {
  __shared__ uint32_t hsm[4 * kSortRadix];
  for (int t = threadIdx.x; t < 4 * kSortRadix; t += blockDim.x) hsm[t] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float il[3] = {inv_leaf.x, inv_leaf.y, inv_leaf.z};
  int min_b[3], div_b[3];
  long long cells = 1;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float mn = ord2f(m->lo[a]), mx = ord2f(m->hi[a]);
    min_b[a] = (int)floorf(__fmul_rn(mn, il[a]));
    const int max_b = (int)floorf(__fmul_rn(mx, il[a]));
    div_b[a] = max_b - min_b[a] + 1;
    cells *= (long long)(__fmul_rn(__fsub_rn(mx, mn), il[a])) + 1;             // the int64 dx*dy*dz overflow test of PCL
  }
  if (i == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) { m->min_b[a] = min_b[a]; m->div_b[a] = div_b[a]; }
    m->overflow = cells > 2147483647ll ? 1 : 0;
  }
  if (i < n) {
    const float p[3] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
    unsigned long long key = kInvalidVoxel;
    if (finite3(p[0], p[1], p[2])) {
      int ijk[3];
#pragma unroll
      for (int a = 0; a < 3; a++) ijk[a] = (int)__fsub_rn(floorf(__fmul_rn(p[a], il[a])), (float)min_b[a]);
      key = (unsigned long long)(unsigned int)(ijk[0] + ijk[1] * div_b[0] + ijk[2] * div_b[0] * div_b[1]);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
#pragma unroll
    for (int ps = 0; ps < 4; ps++) atomicAdd(&hsm[ps * kSortRadix + sort_digit_of(key, 0, ps)], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * kSortRadix; t += blockDim.x)
    if (hsm[t]) atomicAdd(&digit_hist[t], hsm[t]);
}

__device__ __forceinline__ bool voxel_head(const unsigned long long* __restrict__ keys, int j, int n) {
  if (j >= n) return false;
  const unsigned long long k = keys[j];
  return k != kInvalidVoxel && (j == 0 || keys[j - 1] != k);
}

__global__ void __launch_bounds__(kFiltThreads) vg_count_kernel(const unsigned long long* __restrict__ keys, int n, unsigned int* __restrict__ block_sums) {
  const int j = blockIdx.x * kFiltThreads + threadIdx.x;
  const int c = __syncthreads_count(voxel_head(keys, j, n));
  if (threadIdx.x == 0) block_sums[blockIdx.x] = (unsigned int)c;
}

// one thread per occupied voxel (the first point of its run in the sorted order): fp32 sum of its points in ascending
// input order, divided by the count (pcl::CentroidPoint / AccumulatorXYZ); .w carries the count
__global__ void __launch_bounds__(kFiltThreads) vg_centroid_kernel(const float* __restrict__ xyz, int stride, int n, const unsigned long long* __restrict__ keys,
                                                                   const uint32_t* __restrict__ vals, const unsigned int* __restrict__ block_offs,
                                                                   float4* __restrict__ out, int* __restrict__ out_voxel) {
  const int j = blockIdx.x * kFiltThreads + threadIdx.x;
  const bool head = voxel_head(keys, j, n);
  unsigned int total;
  const unsigned int r = block_rank(head, &total);
  if (!head) return;
  const unsigned long long k = keys[j];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int cnt = 0;
  for (int t = j; t < n && keys[t] == k; t++) {
    const float* p = xyz + (size_t)vals[t] * stride;
    sx = __fadd_rn(sx, p[0]); sy = __fadd_rn(sy, p[1]); sz = __fadd_rn(sz, p[2]);
    cnt++;
  }
  const float fc = (float)cnt;
  const unsigned int pos = block_offs[blockIdx.x] + r;
  out[pos] = make_float4(__fdiv_rn(sx, fc), __fdiv_rn(sy, fc), __fdiv_rn(sz, fc), fc);
  if (out_voxel) out_voxel[pos] = (int)k;
}

__global__ void __launch_bounds__(256) f4_to_xyz_kernel(const float4* __restrict__ in, int n, float* __restrict__ out3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float4 p = in[i]; out3[3 * (size_t)i] = p.x; out3[3 * (size_t)i + 1] = p.y; out3[3 * (size_t)i + 2] = p.z; }
}
__global__ void __launch_bounds__(256) strided_to_f4_kernel(const float* __restrict__ in, int stride, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(in[(size_t)i * stride], in[(size_t)i * stride + 1], in[(size_t)i * stride + 2], 1.0f);
}

// ---- deskew (§8f row 3) ----
// order-preserving 64-bit integer image of a point's time stamp; cropped / non-finite points get the largest key
constexpr unsigned long long kDroppedStamp = ~0ull;
__device__ __forceinline__ unsigned long long stamp_key(const unsigned char* rec, int time_type) {
  if (time_type == 0) return (unsigned long long)*reinterpret_cast<const unsigned int*>(rec);          // Ouster `t`, uint32 ns
  if (time_type == 1) return (unsigned long long)f2ord(*reinterpret_cast<const float*>(rec));           // Velodyne `time`, float s
  const unsigned long long u = *reinterpret_cast<const unsigned long long*>(rec);                       // Hesai `timestamp`, double s
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double stamp_value(unsigned long long key, int time_type) {
  if (time_type == 0) return (double)(unsigned int)key;
  if (time_type == 1) {
    const unsigned int o = (unsigned int)key;
    const unsigned int u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return (double)f;
  }
  const unsigned long long u = (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
  double d;
  memcpy(&d, &u, 8);
  return d;
}

__global__ void __launch_bounds__(256) ingest_keys_kernel(const unsigned char* __restrict__ recs, size_t stride_bytes, size_t time_off, int time_type, int n,
                                                          Box box, int use_box, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                                          uint32_t* __restrict__ digit_hist, int passes) {
  __shared__ uint32_t hsm[8 * kSortRadix];
  for (int t = threadIdx.x; t < passes * kSortRadix; t += blockDim.x) hsm[t] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned char* rec = recs + (size_t)i * stride_bytes;
    const float* p = reinterpret_cast<const float*>(rec);
    const bool keep = use_box ? crop_keeps(box, p[0], p[1], p[2]) : finite3(p[0], p[1], p[2]);
    unsigned long long key = keep ? stamp_key(rec + time_off, time_type) : kDroppedStamp;
    if (keep && key == kDroppedStamp) key = kDroppedStamp - 1;   // a NaN stamp: keep the point, sort it last
    keys[i] = key;
    vals[i] = (uint32_t)i;
    for (int ps = 0; ps < passes; ps++) atomicAdd(&hsm[ps * kSortRadix + sort_digit_of(key, 0, ps)], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < passes * kSortRadix; t += blockDim.x)
    if (hsm[t]) atomicAdd(&digit_hist[t], hsm[t]);
}

__device__ __forceinline__ bool stamp_head(const unsigned long long* __restrict__ keys, int j, int n) {
  if (j >= n) return false;
  const unsigned long long k = keys[j];
  return k != kDroppedStamp && (j == 0 || keys[j - 1] != k);
}
// block_sums[b] = unique stamps that start in block b; block_sums[nb + 1 + b] = kept points in block b
__global__ void __launch_bounds__(kFiltThreads) ingest_count_kernel(const unsigned long long* __restrict__ keys, int n, unsigned int* __restrict__ head_sums,
                                                                    unsigned int* __restrict__ kept_total) {
  const int j = blockIdx.x * kFiltThreads + threadIdx.x;
  const int heads = __syncthreads_count(stamp_head(keys, j, n));
  const int kept = __syncthreads_count(j < n && keys[j] != kDroppedStamp);
  if (threadIdx.x == 0) { head_sums[blockIdx.x] = (unsigned int)heads; atomicAdd(kept_total, (unsigned int)kept); }
}
// gathers the kept points in time order and writes every unique stamp (ascending)
__global__ void __launch_bounds__(kFiltThreads) ingest_gather_kernel(const unsigned char* __restrict__ recs, size_t stride_bytes, int n,
                                                                     const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                     const unsigned int* __restrict__ head_offs, int time_type, float4* __restrict__ pts,
                                                                     double* __restrict__ unique_stamps) {
  const int j = blockIdx.x * kFiltThreads + threadIdx.x;
  const bool head = stamp_head(keys, j, n);
  unsigned int total;
  const unsigned int r = block_rank(head, &total);
  if (j >= n) return;
  const unsigned long long k = keys[j];
  if (k == kDroppedStamp) return;                   // dropped points sort last: kept points are exactly positions [0, kept)
  const float* p = reinterpret_cast<const float*>(recs + (size_t)vals[j] * stride_bytes);
  pts[j] = make_float4(p[0], p[1], p[2], 1.0f);
  if (head) unique_stamps[head_offs[blockIdx.x] + r] = stamp_value(k, time_type);
}
// pt = T_group * pt (odom.cc:690-701): ((m0 x + m1 y) + m2 z) + m3, fp32, no FMA; group = unique stamps up to this point - 1
__global__ void __launch_bounds__(kFiltThreads) deskew_apply_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ keys, int n,
                                                                    const unsigned int* __restrict__ head_offs, const float* __restrict__ frames, int n_frames,
                                                                    float4* __restrict__ out) {
  const int j = blockIdx.x * kFiltThreads + threadIdx.x;
  const bool head = stamp_head(keys, j, n);
  // inclusive count of heads up to this thread
  __shared__ unsigned int wsum[kFiltThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int m = __ballot_sync(0xffffffffu, head);
  if (lane == 0) wsum[warp] = __popc(m);
  __syncthreads();
  unsigned int before = 0;
  for (int w = 0; w < warp; w++) before += wsum[w];
  if (j >= n) return;
  const int g = (int)(head_offs[blockIdx.x] + before + __popc(m & ((2u << lane) - 1u))) - 1;
  const float* T = frames + (size_t)min(max(g, 0), n_frames - 1) * 16;   // column-major 4x4
  const float4 p = pts[j];
  float q[3];
#pragma unroll
  for (int r = 0; r < 3; r++)
    q[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[r], p.x), __fmul_rn(T[4 + r], p.y)), __fmul_rn(T[8 + r], p.z)), T[12 + r]);
  out[j] = make_float4(q[0], q[1], q[2], 1.0f);
}
__global__ void zero_u32_kernel(unsigned int* p) { *p = 0u; }

int wait_count(Handle* h, unsigned long long seq, size_t* out) {
  if (h->overlap_fn) {                 // the kernels are in flight: the caller's host work goes here
    auto fn = h->overlap_fn;
    h->overlap_fn = nullptr;
    fn(h->overlap_arg);
  }
  volatile unsigned long long* p = &h->slot_host[0].seq;
  unsigned long long spins = 0;
  while (*p != seq) {
    if ((++spins & 0x3fff) != 0) continue;
    const cudaError_t q = cudaStreamQuery(h->stream);
    if (q == cudaErrorNotReady) continue;
    if (q != cudaSuccess) return fail(h, NGICP_ERR_CUDA, std::string("filter kernel: ") + cudaGetErrorString(q));
    if (*p != seq) return fail(h, NGICP_ERR_CUDA, "filter kernel finished without writing its count");
  }
  __sync_synchronize();
  *out = (size_t)(h->slot_host[0].v[0] + 0.5);
  return NGICP_OK;
}

}  // namespace

// d_in: n points, `stride` floats apart, on the device. *d_out: new float4 array (caller frees), *n_out its length.
int crop_box_device(Handle* h, const float* d_in, int stride, int n, const float mn[3], const float mx[3], int negative, float4** d_out, size_t* n_out) {
  cudaStream_t s = h->stream;
  Box box;
  for (int a = 0; a < 3; a++) { box.mn[a] = mn[a]; box.mx[a] = mx[a]; }
  box.negative = negative;
  const int nb = (n + kFiltThreads - 1) / kFiltThreads;
  unsigned int* d_sums = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_sums, (size_t)nb + 1, s));
  NGICP_CUDA(h, dev_alloc(d_out, (size_t)n, s));
  const unsigned long long seq = ++h->seq;
  crop_count_kernel<<<nb, kFiltThreads, 0, s>>>(d_in, stride, n, box, d_sums);
  scan_sums_kernel<<<1, 1024, 0, s>>>(d_sums, nb, d_sums + nb, h->slot_dev, seq);
  crop_scatter_kernel<<<nb, kFiltThreads, 0, s>>>(d_in, stride, n, box, d_sums, *d_out);
  count_launch(h, 3);
  NGICP_CUDA(h, cudaGetLastError());
  const int rc = wait_count(h, seq, n_out);
  dev_free(d_sums, s);
  return rc;
}

// *d_out: new float4 array of the voxel centroids (w = points in the voxel), ascending voxel index.
int voxel_grid_device(Handle* h, const float* d_in, int stride, int n, const float leaf[3], float4** d_out, size_t* n_out, int* d_out_voxel) {
  cudaStream_t s = h->stream;
  for (int a = 0; a < 3; a++)
    if (!(leaf[a] > 0.f)) return fail(h, NGICP_ERR_INVALID, "voxel grid: leaf size must be positive");
  const float3 inv_leaf = make_float3(1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]);   // Array4f::Ones() / leaf_size_
  const int nbits = 32, passes = sort_num_passes(nbits);
  auto align_up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_keys_a = 0, o_keys_b = o_keys_a + align_up(sizeof(unsigned long long) * (size_t)n), o_vals_a = o_keys_b + align_up(sizeof(unsigned long long) * (size_t)n),
               o_vals_b = o_vals_a + align_up(sizeof(uint32_t) * (size_t)n), o_sort = o_vals_b + align_up(sizeof(uint32_t) * (size_t)n),
               o_meta = o_sort + align_up(sizeof(uint32_t) * sort_scratch_elems(n, nbits)), bytes = o_meta + align_up(sizeof(VoxelMeta));
  char* scratch = nullptr;
  NGICP_CUDA(h, dev_alloc(&scratch, bytes, s));
  unsigned long long* keys_a = reinterpret_cast<unsigned long long*>(scratch + o_keys_a);
  unsigned long long* keys_b = reinterpret_cast<unsigned long long*>(scratch + o_keys_b);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(scratch + o_vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(scratch + o_vals_b);
  uint32_t* sort_scratch = reinterpret_cast<uint32_t*>(scratch + o_sort);
  VoxelMeta* meta = reinterpret_cast<VoxelMeta*>(scratch + o_meta);
  const int nb256 = (n + 255) / 256;
  vg_init_kernel<<<(passes * kSortRadix + 255) / 256, 256, 0, s>>>(meta, sort_scratch, passes * kSortRadix);
  vg_minmax_kernel<<<std::min(nb256, 148), 256, 0, s>>>(d_in, stride, n, meta);
  vg_keys_kernel<<<nb256, 256, 0, s>>>(d_in, stride, n, inv_leaf, meta, keys_a, vals_a, sort_scratch);
  count_launch(h, 3);
  unsigned long long* keys_sorted = nullptr;
  uint32_t* vals_sorted = nullptr;
  cudaError_t cl_err = cudaSuccess;
  if (cluster_sort_identity(keys_a, n, nbits, keys_b, vals_b, s, &cl_err)) {      // scan-sized: one cluster kernel (cluster_sort.cuh)
    NGICP_CUDA(h, cl_err);
    count_launch(h);
    keys_sorted = keys_b; vals_sorted = vals_b;
  } else {
    count_launch(h, radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, sort_scratch, n, 0, nbits, s, &keys_sorted, &vals_sorted, true));
  }
  const int nb = (n + kFiltThreads - 1) / kFiltThreads;
  unsigned int* d_sums = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_sums, (size_t)nb + 1, s));
  NGICP_CUDA(h, dev_alloc(d_out, (size_t)n, s));
  const unsigned long long seq = ++h->seq;
  vg_count_kernel<<<nb, kFiltThreads, 0, s>>>(keys_sorted, n, d_sums);
  scan_sums_kernel<<<1, 1024, 0, s>>>(d_sums, nb, d_sums + nb, h->slot_dev, seq, &meta->overflow);
  vg_centroid_kernel<<<nb, kFiltThreads, 0, s>>>(d_in, stride, n, keys_sorted, vals_sorted, d_sums, *d_out, d_out_voxel);
  count_launch(h, 3);
  NGICP_CUDA(h, cudaGetLastError());
  int rc = wait_count(h, seq, n_out);
  const int overflow = !rc && h->slot_host[0].v[1] != 0.0;
  dev_free(d_sums, s);
  dev_free(scratch, s);
  if (!rc && overflow) {
    // "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input (voxel_grid.hpp)
    strided_to_f4_kernel<<<nb256, 256, 0, s>>>(d_in, stride, n, *d_out);
    count_launch(h);
    *n_out = (size_t)n;
  }
  return rc;
}

}  // namespace ngicp

using namespace ngicp;

extern "C" {

int ngicp_filter_scan(ngicp_handle* p, const void* points, size_t n, size_t stride_bytes, const float crop_min[3], const float crop_max[3],
                      int crop_negative, const float leaf[3], int set_as, float* out_xyz, size_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(p);
  if (!h || !points || n == 0 || !n_out) return fail(h, NGICP_ERR_INVALID, "ngicp_filter_scan: bad argument");
  if (set_as < -1 || set_as > 1) return fail(h, NGICP_ERR_INVALID, "ngicp_filter_scan: set_as must be -1, 0 (source) or 1 (target)");
  if ((crop_min == nullptr) != (crop_max == nullptr)) return fail(h, NGICP_ERR_INVALID, "ngicp_filter_scan: crop_min and crop_max go together");
  if (n >= (size_t)1 << 31) return fail(h, NGICP_ERR_UNSUPPORTED, "ngicp_filter_scan: more than 2^31 points");
  if (int rc = select_device(h)) return rc;
  cudaStream_t s = h->stream;
  float* d_in = nullptr;
  int stride = 3;
  if (int rc = upload_points(h, points, n, stride_bytes, &d_in, &stride)) return rc;
  const float* cur = d_in;
  size_t cur_n = n;
  float4* d_crop = nullptr;
  float4* d_vox = nullptr;
  int rc = NGICP_OK;
  if (crop_min) {
    rc = crop_box_device(h, cur, stride, (int)cur_n, crop_min, crop_max, crop_negative, &d_crop, &cur_n);
    if (!rc) { cur = reinterpret_cast<const float*>(d_crop); stride = 4; }
  }
  if (!rc && leaf && cur_n > 0) {
    rc = voxel_grid_device(h, cur, stride, (int)cur_n, leaf, &d_vox, &cur_n, nullptr);
    if (!rc) { cur = reinterpret_cast<const float*>(d_vox); stride = 4; }
  }
  if (!rc && out_xyz && cur_n > 0) {
    float* d_o = nullptr;
    NGICP_CUDA(h, dev_alloc(&d_o, cur_n * 3, s));
    if (stride == 4) {
      f4_to_xyz_kernel<<<((int)cur_n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4*>(cur), (int)cur_n, d_o);
      count_launch(h);
      NGICP_CUDA(h, cudaMemcpyAsync(out_xyz, d_o, cur_n * 12, cudaMemcpyDeviceToHost, s));
    } else {
      NGICP_CUDA(h, cudaMemcpy2DAsync(out_xyz, 12, cur, (size_t)stride * 4, 12, cur_n, cudaMemcpyDeviceToHost, s));
    }
    NGICP_CUDA(h, cudaStreamSynchronize(s));
    dev_free(d_o, s);
  }
  if (!rc && set_as >= 0) {
    if (cur_n == 0) rc = fail(h, NGICP_ERR_INVALID, "ngicp_filter_scan: the filters removed every point");
    else {
      Index* idx = nullptr;
      rc = build_index(h, cur, stride, (int)cur_n, nullptr, 1, &idx);
      if (!rc) rc = swap_in_index(h, set_as, idx);
    }
  }
  *n_out = cur_n;
  dev_free(d_in, s); dev_free(d_crop, s); dev_free(d_vox, s);
  const int rc2 = finish_input_copy(h);
  return rc ? rc : rc2;
}

int ngicp_scan_ingest(ngicp_handle* p, const void* points, size_t n, size_t stride_bytes, size_t time_offset_bytes, int time_type,
                      const float crop_min[3], const float crop_max[3], int crop_negative, double* unique_stamps, size_t* n_unique, size_t* n_kept) {
  Handle* h = reinterpret_cast<Handle*>(p);
  if (!h || !points || n == 0 || !n_unique) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_ingest: bad argument");
  if (time_type < 0 || time_type > 2) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_ingest: time_type must be 0 (uint32), 1 (float) or 2 (double)");
  const size_t tsize = time_type == 2 ? 8 : 4;
  if (stride_bytes < 12 || stride_bytes % 4 || time_offset_bytes % tsize || time_offset_bytes + tsize > stride_bytes)
    return fail(h, NGICP_ERR_INVALID, "ngicp_scan_ingest: bad stride or time-stamp offset");
  if ((crop_min == nullptr) != (crop_max == nullptr)) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_ingest: crop_min and crop_max go together");
  if (n >= (size_t)1 << 31) return fail(h, NGICP_ERR_UNSUPPORTED, "ngicp_scan_ingest: more than 2^31 points");
  if (int rc = select_device(h)) return rc;
  cudaStream_t s = h->stream;
  dev_free(h->scan_pts, s); dev_free(h->scan_keys, s);
  h->scan_pts = nullptr; h->scan_keys = nullptr; h->scan_n = 0; h->scan_groups = 0;
  Box box;
  for (int a = 0; a < 3; a++) { box.mn[a] = crop_min ? crop_min[a] : 0.f; box.mx[a] = crop_max ? crop_max[a] : 0.f; }
  box.negative = crop_negative;
  // bit 32 separates the dropped points (all-ones key) from real 32-bit stamps; when the caller knows that every uint32 stamp
  // is < 2^b, bit b does (fewer sort passes: an OS1 scan spans 1e8 ns = 27 bits, a MulRan scan has one stamp)
  int nbits = time_type == 2 ? 64 : 33;
  if (time_type == 0 && h->ingest_stamp_bits >= 0 && h->ingest_stamp_bits < 32) nbits = h->ingest_stamp_bits + 1;
  h->ingest_stamp_bits = -1;
  const int passes = sort_num_passes(nbits);
  unsigned char* d_recs = nullptr;
  unsigned long long *keys_a = nullptr, *keys_b = nullptr;
  uint32_t *vals_a = nullptr, *vals_b = nullptr, *sort_scratch = nullptr;
  unsigned int* d_sums = nullptr;
  double* d_unique = nullptr;
  const int nb = ((int)n + kFiltThreads - 1) / kFiltThreads;
  NGICP_CUDA(h, dev_alloc(&d_recs, n * stride_bytes, s));
  NGICP_CUDA(h, dev_alloc(&keys_a, n, s));
  NGICP_CUDA(h, dev_alloc(&keys_b, n, s));
  NGICP_CUDA(h, dev_alloc(&vals_a, n, s));
  NGICP_CUDA(h, dev_alloc(&vals_b, n, s));
  NGICP_CUDA(h, dev_alloc(&sort_scratch, sort_scratch_elems((int)n, nbits), s));
  NGICP_CUDA(h, dev_alloc(&d_sums, (size_t)nb + 2, s));
  NGICP_CUDA(h, dev_alloc(&d_unique, n, s));
  NGICP_CUDA(h, dev_alloc(&h->scan_pts, n, s));
  NGICP_CUDA(h, cudaMemcpyAsync(d_recs, points, n * stride_bytes, cudaMemcpyHostToDevice, s));
  NGICP_CUDA(h, cudaMemsetAsync(sort_scratch, 0, sizeof(uint32_t) * passes * kSortRadix, s));
  ingest_keys_kernel<<<((int)n + 255) / 256, 256, 0, s>>>(d_recs, stride_bytes, time_offset_bytes, time_type, (int)n, box, crop_min ? 1 : 0, keys_a, vals_a,
                                                         sort_scratch, passes);
  count_launch(h);
  unsigned long long* keys_sorted = nullptr;
  uint32_t* vals_sorted = nullptr;
  cudaError_t cl_err = cudaSuccess;
  if (cluster_sort_identity(keys_a, (int)n, nbits, keys_b, vals_b, s, &cl_err)) {   // scan-sized, <= 47 key bits: one cluster kernel
    NGICP_CUDA(h, cl_err);
    count_launch(h);
    keys_sorted = keys_b; vals_sorted = vals_b;
  } else {
    count_launch(h, radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, sort_scratch, (int)n, 0, nbits, s, &keys_sorted, &vals_sorted, true));
  }
  const unsigned long long seq = ++h->seq;
  zero_u32_kernel<<<1, 1, 0, s>>>(d_sums + nb + 1);
  ingest_count_kernel<<<nb, kFiltThreads, 0, s>>>(keys_sorted, (int)n, d_sums, d_sums + nb + 1);
  scan_sums_kernel<<<1, 1024, 0, s>>>(d_sums, nb, d_sums + nb, h->slot_dev, seq);
  ingest_gather_kernel<<<nb, kFiltThreads, 0, s>>>(d_recs, stride_bytes, (int)n, keys_sorted, vals_sorted, d_sums, time_type, h->scan_pts, d_unique);
  count_launch(h, 4);
  NGICP_CUDA(h, cudaGetLastError());
  size_t groups = 0;
  int rc = wait_count(h, seq, &groups);
  unsigned int kept = 0;
  if (!rc) {
    NGICP_CUDA(h, cudaMemcpyAsync(&kept, d_sums + nb + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    if (unique_stamps && groups) NGICP_CUDA(h, cudaMemcpyAsync(unique_stamps, d_unique, sizeof(double) * groups, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaStreamSynchronize(s));
    // keep the sorted keys (they define the groups) next to the sorted points
    h->scan_keys = keys_sorted == keys_a ? keys_a : keys_b;
    (keys_sorted == keys_a ? keys_a : keys_b) = nullptr;
    h->scan_n = kept;
    h->scan_groups = groups;
  }
  dev_free(d_recs, s); dev_free(keys_a, s); dev_free(keys_b, s); dev_free(vals_a, s); dev_free(vals_b, s);
  dev_free(sort_scratch, s); dev_free(d_sums, s); dev_free(d_unique, s);
  *n_unique = groups;
  if (n_kept) *n_kept = kept;
  return rc;
}

int ngicp_scan_deskew(ngicp_handle* p, const float* frames16, size_t n_frames, const float leaf[3], int set_as, float* out_xyz, size_t* n_out) {
  Handle* h = reinterpret_cast<Handle*>(p);
  if (!h || !frames16 || !n_out) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_deskew: bad argument");
  if (!h->scan_pts || h->scan_n == 0) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_deskew: no ingested scan (call ngicp_scan_ingest first)");
  if (n_frames != 1 && n_frames != h->scan_groups)
    return fail(h, NGICP_ERR_INVALID, "ngicp_scan_deskew: need one frame per unique time stamp (or exactly one frame for a rigid transform)");
  if (set_as < -1 || set_as > 1) return fail(h, NGICP_ERR_INVALID, "ngicp_scan_deskew: set_as must be -1, 0 (source) or 1 (target)");
  if (int rc = select_device(h)) return rc;
  cudaStream_t s = h->stream;
  const int n = (int)h->scan_n;
  const int nb = (n + kFiltThreads - 1) / kFiltThreads;
  float* d_frames = nullptr;
  unsigned int* d_sums = nullptr;
  float4* d_out = nullptr;
  float4* d_vox = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_frames, n_frames * 16, s));
  NGICP_CUDA(h, dev_alloc(&d_sums, (size_t)nb + 2, s));
  NGICP_CUDA(h, dev_alloc(&d_out, (size_t)n, s));
  NGICP_CUDA(h, cudaMemcpyAsync(d_frames, frames16, sizeof(float) * 16 * n_frames, cudaMemcpyHostToDevice, s));
  const unsigned long long seq = ++h->seq;     // the group count is known since the ingest: nobody waits for this one
  zero_u32_kernel<<<1, 1, 0, s>>>(d_sums + nb + 1);
  ingest_count_kernel<<<nb, kFiltThreads, 0, s>>>(h->scan_keys, n, d_sums, d_sums + nb + 1);
  scan_sums_kernel<<<1, 1024, 0, s>>>(d_sums, nb, d_sums + nb, h->slot_dev, seq);
  deskew_apply_kernel<<<nb, kFiltThreads, 0, s>>>(h->scan_pts, h->scan_keys, n, d_sums, d_frames, (int)n_frames, d_out);
  count_launch(h, 4);
  NGICP_CUDA(h, cudaGetLastError());
  int rc = NGICP_OK;
  const float* cur = reinterpret_cast<const float*>(d_out);
  size_t cur_n = (size_t)n;
  if (!rc && leaf) {
    rc = voxel_grid_device(h, cur, 4, n, leaf, &d_vox, &cur_n, nullptr);
    if (!rc) cur = reinterpret_cast<const float*>(d_vox);
  }
  if (!rc && out_xyz && cur_n > 0) {
    float* d_o = nullptr;
    NGICP_CUDA(h, dev_alloc(&d_o, cur_n * 3, s));
    f4_to_xyz_kernel<<<((int)cur_n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4*>(cur), (int)cur_n, d_o);
    count_launch(h);
    NGICP_CUDA(h, cudaMemcpyAsync(out_xyz, d_o, cur_n * 12, cudaMemcpyDeviceToHost, s));
    NGICP_CUDA(h, cudaStreamSynchronize(s));
    dev_free(d_o, s);
  }
  if (!rc && set_as >= 0) {
    Index* idx = nullptr;
    rc = build_index(h, cur, 4, (int)cur_n, nullptr, 1, &idx);
    if (!rc) rc = swap_in_index(h, set_as, idx);
  }
  NGICP_CUDA(h, cudaStreamSynchronize(s));   // frames16 is the caller's pageable buffer
  *n_out = cur_n;
  dev_free(d_frames, s); dev_free(d_sums, s); dev_free(d_out, s); dev_free(d_vox, s);
  return rc;
}

}  // extern "C"
