// K1 — spatial index build: voxel keys -> Morton -> radix sort -> reorder -> multi-level voxel hash.
// Replaces nanoflann::KdTreeFLANN::setInputCloud -> KDTreeSingleIndexAdaptor::buildIndex
// (reference src/dlio/include/nano_gicp/nanoflann_adaptor.h:132-138, nanoflann.h:1405-1417,
// divideTree/middleSplit_/planeSplit :1025-1185). Everything stays on the device: the grid
// parameters (origin, cell size, base level) are computed by kernels and consumed by kernels, so
// the build needs no host round trip.
#include <algorithm>
#include <cmath>

#include "internal.h"
#include "radix_sort.cuh"

namespace ngicp {

namespace {

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void set_pair_kernel(int* p, int a, int b) { p[0] = a; p[1] = b; }

__global__ void bbox_init_kernel(unsigned int* lo, unsigned int* hi, int n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * n_seg) { lo[i] = 0xffffffffu; hi[i] = 0u; }
}

// per-segment bounding boxes. Warp-aggregated when the whole warp sits in one segment.
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ xyz, int stride, int n, const int* __restrict__ seg_start, int n_seg,
                                                   unsigned int* __restrict__ lo, unsigned int* __restrict__ hi) {
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    const bool valid = i < n;
    const int ic = valid ? i : n - 1;
    const int seg = find_segment(seg_start, n_seg, ic);
    const float x = xyz[(size_t)ic * stride + 0], y = xyz[(size_t)ic * stride + 1], z = xyz[(size_t)ic * stride + 2];
    unsigned int lx = f2ord(x), ly = f2ord(y), lz = f2ord(z), hx = lx, hy = ly, hz = lz;
    const int seg0 = __shfl_sync(0xffffffffu, seg, 0);
    const bool uniform = __all_sync(0xffffffffu, seg == seg0);
    if (uniform) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        lx = min(lx, __shfl_xor_sync(0xffffffffu, lx, off)); hx = max(hx, __shfl_xor_sync(0xffffffffu, hx, off));
        ly = min(ly, __shfl_xor_sync(0xffffffffu, ly, off)); hy = max(hy, __shfl_xor_sync(0xffffffffu, hy, off));
        lz = min(lz, __shfl_xor_sync(0xffffffffu, lz, off)); hz = max(hz, __shfl_xor_sync(0xffffffffu, hz, off));
      }
      if ((threadIdx.x & 31) == 0) {
        atomicMin(&lo[3 * seg + 0], lx); atomicMax(&hi[3 * seg + 0], hx);
        atomicMin(&lo[3 * seg + 1], ly); atomicMax(&hi[3 * seg + 1], hy);
        atomicMin(&lo[3 * seg + 2], lz); atomicMax(&hi[3 * seg + 2], hz);
      }
    } else if (valid) {
      atomicMin(&lo[3 * seg + 0], lx); atomicMax(&hi[3 * seg + 0], hx);
      atomicMin(&lo[3 * seg + 1], ly); atomicMax(&hi[3 * seg + 1], hy);
      atomicMin(&lo[3 * seg + 2], lz); atomicMax(&hi[3 * seg + 2], hz);
    }
  }
}

// One block. Segment origins = lower bbox corners; one common power-of-two cell size h0 with
// 4096*h0 strictly larger than the largest extent of any segment.
__global__ void __launch_bounds__(256) grid_meta_kernel(const unsigned int* __restrict__ lo, const unsigned int* __restrict__ hi, int n_seg,
                                                        float4* __restrict__ seg_origin, GridMeta* __restrict__ meta) {
  __shared__ float smax[256];
  float ext = 0.f;
  for (int s = threadIdx.x; s < n_seg; s += blockDim.x) {
    const float ox = ord2f(lo[3 * s + 0]), oy = ord2f(lo[3 * s + 1]), oz = ord2f(lo[3 * s + 2]);
    seg_origin[s] = make_float4(ox, oy, oz, 0.f);
    ext = fmaxf(ext, fmaxf(ord2f(hi[3 * s + 0]) - ox, fmaxf(ord2f(hi[3 * s + 1]) - oy, ord2f(hi[3 * s + 2]) - oz)));
  }
  smax[threadIdx.x] = ext;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float e = smax[0];
    if (!(e > 9.765625e-4f)) e = 9.765625e-4f;  // degenerate clouds: keep a sane cell size
    if (!(e < 1e30f)) e = 1e30f;
    int ex;
    frexpf(e * 1.01f, &ex);                    // e*1.01 < 2^ex
    const float h0 = ldexpf(1.0f, ex - kBitsPerAxis);
    meta->h0 = h0;
    meta->inv_h0 = 1.0f / h0;                  // exact: power of two
    meta->margin = h0 * (1.0f / 512.0f);
    meta->base_level = 0;
    meta->cells_total = 0;
    for (int i = 0; i < 16; i++) meta->level_hist[i] = 0;
  }
}

__global__ void __launch_bounds__(256) keys_kernel(const float* __restrict__ xyz, int stride, int n, const int* __restrict__ seg_start, int n_seg,
                                                   const float4* __restrict__ seg_origin, const GridMeta* __restrict__ meta,
                                                   unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int seg = find_segment(seg_start, n_seg, i);
  const float4 o = __ldg(seg_origin + seg);
  const float inv_h0 = __ldg(&meta->inv_h0);
  const float x = xyz[(size_t)i * stride + 0], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
  const unsigned int cx = clampi(voxel_coord_unclamped(x, o.x, inv_h0), 0, kMaxCoord);
  const unsigned int cy = clampi(voxel_coord_unclamped(y, o.y, inv_h0), 0, kMaxCoord);
  const unsigned int cz = clampi(voxel_coord_unclamped(z, o.z, inv_h0), 0, kMaxCoord);
  keys[i] = ((unsigned long long)seg << kMortonBits) | morton3(cx, cy, cz);
  vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_points_kernel(const float* __restrict__ xyz, int stride, int n, const uint32_t* __restrict__ vals,
                                                            float4* __restrict__ pts, int* __restrict__ inv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t i = vals[j];
  const float* p = xyz + (size_t)i * stride;
  pts[j] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
  inv[i] = j;
}

// first level at which two keys fall into different cells, +1; kNumLevels (13) if they differ at the top
__device__ __forceinline__ int diff_levels(unsigned long long a, unsigned long long b) {
  const unsigned long long x = a ^ b;
  if (x == 0) return 0;
  const int msb = 63 - __clzll((long long)x);
  return min(msb / 3 + 1, kNumLevels);
}

__global__ void __launch_bounds__(256) level_hist_kernel(const unsigned long long* __restrict__ keys, int n, GridMeta* __restrict__ meta) {
  __shared__ unsigned int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int d = j == 0 ? kNumLevels : diff_levels(keys[j], keys[j - 1]);
    if (d) atomicAdd(&h[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 16 && h[threadIdx.x]) atomicAdd(&meta->level_hist[threadIdx.x], h[threadIdx.x]);
}

// cells(L) = #positions with diff_levels > L. Base level = finest level whose mean occupancy is
// >= occupancy and whose cumulative entry count (this level and all coarser ones) fits the table.
__global__ void choose_base_kernel(GridMeta* meta, int n, unsigned int max_entries, int occupancy) {
  if (threadIdx.x || blockIdx.x) return;
  unsigned int cells[kNumLevels];
  unsigned int acc = 0;
  for (int L = kTopLevel; L >= 0; L--) {
    acc += meta->level_hist[L + 1];
    cells[L] = acc;
  }
  int base = kTopLevel;
  unsigned int total = cells[kTopLevel];
  for (int L = kTopLevel - 1; L >= 0; L--) {
    if ((unsigned long long)cells[L] * (unsigned)occupancy > (unsigned long long)n) break;
    if (total + cells[L] > max_entries) break;
    total += cells[L];
    base = L;
  }
  meta->base_level = base;
  meta->cells_total = total;
}

__global__ void __launch_bounds__(256) table_insert_kernel(const unsigned long long* __restrict__ keys, int n, const GridMeta* __restrict__ meta,
                                                           CellSlot* __restrict__ table, uint32_t mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int base = meta->base_level;
  const unsigned long long k = keys[j];
  const int d = j == 0 ? kNumLevels : diff_levels(k, keys[j - 1]);
  for (int L = base; L < d; L++) {
    const unsigned long long ck = cell_key(k, L);
    uint32_t h = hash64(ck) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&table[h].key, kEmptyKey, ck);
      if (prev == kEmptyKey) { table[h].start = (uint32_t)j; break; }
      h = (h + 1) & mask;
    }
  }
}

__global__ void __launch_bounds__(256) table_close_kernel(const unsigned long long* __restrict__ keys, int n, const GridMeta* __restrict__ meta,
                                                          CellSlot* __restrict__ table, uint32_t mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int base = meta->base_level;
  const unsigned long long k = keys[j];
  const int d = j == n - 1 ? kNumLevels : diff_levels(k, keys[j + 1]);
  for (int L = base; L < d; L++) {
    const unsigned long long ck = cell_key(k, L);
    uint32_t h = hash64(ck) & mask;
    for (;;) {
      const unsigned long long cur = table[h].key;
      if (cur == ck) { table[h].end = (uint32_t)(j + 1); break; }
      if (cur == kEmptyKey) break;  // cannot happen: every closing cell was opened by table_insert_kernel
      h = (h + 1) & mask;
    }
  }
}

inline int ceil_log2(unsigned int v) { int b = 0; while ((1u << b) < v) b++; return b; }

}  // namespace

void free_index(Index* idx, cudaStream_t stream) {
  if (!idx) return;
  dev_free(idx->pts, stream);
  dev_free(idx->inv, stream);
  dev_free(idx->keys, stream);
  dev_free(idx->table, stream);
  dev_free(idx->meta, stream);
  dev_free(idx->seg_origin, stream);
  dev_free(idx->seg_start, stream);
  delete idx;
}

int build_index(Handle* h, const float* d_xyz, int stride, int n, const int64_t* seg_offsets, int n_seg, Index** out) {
  if (n <= 0) return fail(h, NGICP_ERR_INVALID, "index build: empty cloud");
  if (n_seg < 1 || n_seg > (1 << kMaxSegBits)) return fail(h, NGICP_ERR_INVALID, "index build: bad segment count");
  if ((long long)n >= (1ll << 31) - 4096) return fail(h, NGICP_ERR_UNSUPPORTED, "index build: more than 2^31 points");
  StageTimer timer(h, &h->t.index_ms);
  cudaStream_t s = h->stream;
  Index* idx = new Index;
  idx->device = h->device;
  idx->n = n;
  idx->n_seg = n_seg;
  idx->seg_offsets_host.resize(n_seg + 1);
  std::vector<int> seg32(n_seg + 1);
  for (int i = 0; i <= n_seg; i++) {
    const int64_t v = seg_offsets ? seg_offsets[i] : (i == 0 ? 0 : n);
    idx->seg_offsets_host[i] = v;
    seg32[i] = (int)v;
  }
  if (seg32[0] != 0 || seg32[n_seg] != n) { delete idx; return fail(h, NGICP_ERR_INVALID, "index build: segment offsets must span [0,n]"); }
  for (int i = 0; i < n_seg; i++)
    if (seg32[i + 1] <= seg32[i]) { delete idx; return fail(h, NGICP_ERR_INVALID, "index build: empty or unordered segment"); }

  const int nbits = kMortonBits + (n_seg > 1 ? ceil_log2((unsigned)n_seg) : 0);
  unsigned int cap = 1024;
  while (cap < (unsigned long long)n + n / 2) cap <<= 1;
  idx->table_mask = cap - 1;

  unsigned long long *keys_a = nullptr, *keys_b = nullptr, *keys_sorted = nullptr;
  uint32_t *vals_a = nullptr, *vals_b = nullptr, *vals_sorted = nullptr, *sort_scratch = nullptr;
  unsigned int* bbox = nullptr;
  auto cleanup = [&]() {
    dev_free(keys_b, s); dev_free(vals_a, s); dev_free(vals_b, s); dev_free(sort_scratch, s); dev_free(bbox, s);
  };
#define IDX_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      cleanup(); dev_free(keys_a, s); idx->keys = nullptr; free_index(idx, s);                           \
      return fail(h, NGICP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    }                                                                                                    \
  } while (0)

  IDX_CUDA(dev_alloc(&idx->pts, (size_t)n, s));
  IDX_CUDA(dev_alloc(&idx->inv, (size_t)n, s));
  IDX_CUDA(dev_alloc(&idx->table, (size_t)cap, s));
  IDX_CUDA(dev_alloc(&idx->meta, 1, s));
  IDX_CUDA(dev_alloc(&idx->seg_origin, (size_t)n_seg, s));
  IDX_CUDA(dev_alloc(&idx->seg_start, (size_t)n_seg + 1, s));
  IDX_CUDA(dev_alloc(&keys_a, (size_t)n, s));
  IDX_CUDA(dev_alloc(&keys_b, (size_t)n, s));
  IDX_CUDA(dev_alloc(&vals_a, (size_t)n, s));
  IDX_CUDA(dev_alloc(&vals_b, (size_t)n, s));
  IDX_CUDA(dev_alloc(&sort_scratch, sort_scratch_elems(n, nbits), s));
  IDX_CUDA(dev_alloc(&bbox, (size_t)6 * n_seg, s));
  if (n_seg == 1) {
    set_pair_kernel<<<1, 1, 0, s>>>(idx->seg_start, 0, n);  // no host buffer, no sync on the per-scan path
    count_launch(h);
  } else {
    IDX_CUDA(cudaMemcpyAsync(idx->seg_start, seg32.data(), sizeof(int) * (n_seg + 1), cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaStreamSynchronize(s));  // seg32 is a stack-lifetime buffer
  }
  IDX_CUDA(cudaMemsetAsync(idx->table, 0xff, sizeof(CellSlot) * (size_t)cap, s));

  const int tpb = 256;
  const int nb = (n + tpb - 1) / tpb;
  unsigned int *lo = bbox, *hi = bbox + 3 * n_seg;
  bbox_init_kernel<<<(3 * n_seg + 255) / 256, 256, 0, s>>>(lo, hi, n_seg);
  bbox_kernel<<<std::min(nb, 148 * 8), tpb, 0, s>>>(d_xyz, stride, n, idx->seg_start, n_seg, lo, hi);
  grid_meta_kernel<<<1, 256, 0, s>>>(lo, hi, n_seg, idx->seg_origin, idx->meta);
  keys_kernel<<<nb, tpb, 0, s>>>(d_xyz, stride, n, idx->seg_start, n_seg, idx->seg_origin, idx->meta, keys_a, vals_a);
  count_launch(h, 4);
  count_launch(h, radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, sort_scratch, n, nbits, s, &keys_sorted, &vals_sorted));
  gather_points_kernel<<<nb, tpb, 0, s>>>(d_xyz, stride, n, vals_sorted, idx->pts, idx->inv);
  level_hist_kernel<<<std::min(nb, 148 * 8), tpb, 0, s>>>(keys_sorted, n, idx->meta);
  choose_base_kernel<<<1, 32, 0, s>>>(idx->meta, n, cap / 2, 2);
  table_insert_kernel<<<nb, tpb, 0, s>>>(keys_sorted, n, idx->meta, idx->table, idx->table_mask);
  table_close_kernel<<<nb, tpb, 0, s>>>(keys_sorted, n, idx->meta, idx->table, idx->table_mask);
  count_launch(h, 5);
  IDX_CUDA(cudaGetLastError());
  // keep the sorted keys, drop the other ping-pong buffer
  idx->keys = keys_sorted;
  if (keys_sorted == keys_a) { /* keys_b freed by cleanup */ } else { dev_free(keys_a, s); keys_b = nullptr; }
  cleanup();
#undef IDX_CUDA
  *out = idx;
  return NGICP_OK;
}

}  // namespace ngicp
