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

// one launch that resets everything the build accumulates into: bounding boxes, the sort's digit totals, and
// (single-keyframe builds) the segment table, which then needs no host buffer and no sync on the per-scan path
__global__ void __launch_bounds__(256) index_prep_kernel(unsigned int* lo, unsigned int* hi, int n_seg, uint32_t* digit_hist, int n_hist,
                                                        int* seg_start, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * n_seg) { lo[i] = 0xffffffffu; hi[i] = 0u; }
  if (i < n_hist) digit_hist[i] = 0u;
  if (seg_start && i == 0) { seg_start[0] = 0; seg_start[1] = n; }
}

// per-segment bounding boxes. Single-keyframe clouds: registers -> warp shuffles -> shared memory -> six atomics per
// block. Multi-keyframe clouds: warp-aggregated when the whole warp sits in one segment.
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ xyz, int stride, int n, const int* __restrict__ seg_start, int n_seg,
                                                   unsigned int* __restrict__ lo, unsigned int* __restrict__ hi) {
  if (n_seg == 1) {
    unsigned int l[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, u[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const unsigned int o = f2ord(xyz[(size_t)i * stride + a]);
        l[a] = min(l[a], o); u[a] = max(u[a], o);
      }
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
      atomicMin(&lo[threadIdx.x], ml);
      atomicMax(&hi[threadIdx.x], mu);
    }
    return;
  }
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

// power-of-two level-0 cell size with 4096*h0 strictly larger than the largest bbox extent
__device__ __forceinline__ float cell_size_for(float e) {
  if (!(e > 9.765625e-4f)) e = 9.765625e-4f;  // degenerate clouds: keep a sane cell size
  if (!(e < 1e30f)) e = 1e30f;
  int ex;
  frexpf(e * 1.01f, &ex);                     // e*1.01 < 2^ex
  return ldexpf(1.0f, ex - kBitsPerAxis);
}
__device__ __forceinline__ void write_meta(GridMeta* meta, float h0) {
  meta->h0 = h0;
  meta->inv_h0 = 1.0f / h0;                   // exact: power of two
  meta->margin = h0 * (1.0f / 512.0f);
  meta->base_level = 0;
  meta->fine_level = 0;
  meta->cells_total = 0;
  for (int i = 0; i < 16; i++) meta->level_hist[i] = 0;
}

// One block. Segment origins = lower bbox corners; one common power-of-two cell size h0 with
// 4096*h0 strictly larger than the largest extent of any segment. (Multi-keyframe builds only; a
// single-keyframe build derives the same values inside keys_kernel.)
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
  if (threadIdx.x == 0) write_meta(meta, cell_size_for(smax[0]));
}

// voxel keys + (fused) the radix sort's digit totals for every pass. Single-keyframe builds also derive the grid
// parameters here, every thread from the same six bbox words, so no separate one-block launch sits on the path.
__global__ void __launch_bounds__(256) keys_kernel(const float* __restrict__ xyz, int stride, int n, const int* __restrict__ seg_start, int n_seg,
                                                   float4* __restrict__ seg_origin, GridMeta* __restrict__ meta,
                                                   const unsigned int* __restrict__ lo, const unsigned int* __restrict__ hi,
                                                   unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                                   uint32_t* __restrict__ digit_hist, int low_bit, int passes) {
  __shared__ uint32_t hsm[8 * kSortRadix];
  for (int t = threadIdx.x; t < passes * kSortRadix; t += blockDim.x) hsm[t] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int seg = 0;
    float4 o;
    float inv_h0;
    if (n_seg == 1) {
      o = make_float4(ord2f(lo[0]), ord2f(lo[1]), ord2f(lo[2]), 0.f);
      const float h0 = cell_size_for(fmaxf(ord2f(hi[0]) - o.x, fmaxf(ord2f(hi[1]) - o.y, ord2f(hi[2]) - o.z)));
      inv_h0 = 1.0f / h0;
      if (i == 0) { seg_origin[0] = o; write_meta(meta, h0); }
    } else {
      seg = find_segment(seg_start, n_seg, i);
      o = seg_origin[seg];
      inv_h0 = meta->inv_h0;
    }
    const float x = xyz[(size_t)i * stride + 0], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    const unsigned int cx = clampi(voxel_coord_unclamped(x, o.x, inv_h0), 0, kMaxCoord);
    const unsigned int cy = clampi(voxel_coord_unclamped(y, o.y, inv_h0), 0, kMaxCoord);
    const unsigned int cz = clampi(voxel_coord_unclamped(z, o.z, inv_h0), 0, kMaxCoord);
    const unsigned long long key = ((unsigned long long)seg << kMortonBits) | morton3(cx, cy, cz);
    keys[i] = key;
    vals[i] = (uint32_t)i;
    for (int p = 0; p < passes; p++) atomicAdd(&hsm[p * kSortRadix + sort_digit_of(key, low_bit, p)], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < passes * kSortRadix; t += blockDim.x)
    if (hsm[t]) atomicAdd(&digit_hist[t], hsm[t]);
}

// first level at which two keys fall into different cells, +1; kNumLevels (13) if they differ at the top
__device__ __forceinline__ int diff_levels(unsigned long long a, unsigned long long b) {
  const unsigned long long x = a ^ b;
  if (x == 0) return 0;
  const int msb = 63 - __clzll((long long)x);
  return min(msb / 3 + 1, kNumLevels);
}

// reorder the points into Morton order (+ inverse permutation) and, in the same pass over the sorted keys, count at
// which octree level every sorted position opens a new cell
__global__ void __launch_bounds__(256) gather_levels_kernel(const float* __restrict__ xyz, int stride, int n, const uint32_t* __restrict__ vals,
                                                            const unsigned long long* __restrict__ keys, float4* __restrict__ pts,
                                                            int* __restrict__ inv, GridMeta* __restrict__ meta) {
  __shared__ unsigned int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const uint32_t i = vals[j];
    const float* p = xyz + (size_t)i * stride;
    pts[j] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
    inv[i] = j;
    const int d = j == 0 ? kNumLevels : diff_levels(keys[j], keys[j - 1]);
    if (d) atomicAdd(&h[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 16 && h[threadIdx.x]) atomicAdd(&meta->level_hist[threadIdx.x], h[threadIdx.x]);
}

// cells(L) = #positions with diff_levels > L. Base level = finest level whose mean occupancy is >= occupancy and whose
// cumulative entry count (this level and all coarser ones) fits the table. Fine level = as many finer levels below the
// base level as still fit the table and still merge points (mean occupancy >= 1.11): a LiDAR scan is two orders of
// magnitude denser next to the sensor than its mean, and there the self k-NN wants cells finer than the base level.
__device__ __forceinline__ int choose_base(const GridMeta* meta, int n, unsigned int max_entries, int occupancy, bool want_fine, int fine_occ10, unsigned int* total_out,
                                           int* fine_out) {
  unsigned int cells[kNumLevels];
  unsigned int acc = 0;
  for (int L = kTopLevel; L >= 0; L--) {
    acc += meta->level_hist[L + 1];
    cells[L] = acc;
  }
  int base = kTopLevel;
  unsigned int total = cells[kTopLevel];
  for (int L = kTopLevel - 1; L >= kBaseFloor; L--) {
    if ((unsigned long long)cells[L] * (unsigned)occupancy > (unsigned long long)n) break;
    if (total + cells[L] > max_entries) break;
    total += cells[L];
    base = L;
  }
  int fine = base;
  for (int L = base - 1; want_fine && L >= kSortLevel; L--) {
    if ((unsigned long long)cells[L] * (unsigned long long)fine_occ10 > (unsigned long long)n * 10ull) break;
    if (total + cells[L] > max_entries) break;
    total += cells[L];
    fine = L;
  }
  if (total_out) *total_out = total;
  if (fine_out) *fine_out = fine;
  return base;
}

// Every sorted position opens the cells whose first point it is and closes the cells whose last point it is; both find
// the cell's slot by find-or-insert (atomicCAS on the key), so one pass over the sorted keys fills start AND end:
// whichever of the two threads arrives first claims the slot, the fields they write are disjoint.
__global__ void __launch_bounds__(256) table_build_kernel(const unsigned long long* __restrict__ keys, int n, GridMeta* __restrict__ meta,
                                                          CellSlot* __restrict__ table, uint32_t mask, unsigned int max_entries, int occupancy, bool want_fine, int fine_occ10) {
  __shared__ int s_fine;
  if (threadIdx.x == 0) {
    unsigned int total;
    int fine;
    const int b = choose_base(meta, n, max_entries, occupancy, want_fine, fine_occ10, &total, &fine);
    s_fine = fine;
    if (blockIdx.x == 0) { meta->base_level = b; meta->fine_level = fine; meta->cells_total = total; }   // published for every later kernel
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int base = s_fine;
  const unsigned long long k = keys[j];
  const int d_open = j == 0 ? kNumLevels : diff_levels(k, keys[j - 1]);
  const int d_close = j == n - 1 ? kNumLevels : diff_levels(k, keys[j + 1]);
  const int d = max(d_open, d_close);
  for (int L = base; L < d; L++) {
    const unsigned long long ck = cell_key(k, L);
    uint32_t h = hash64(ck) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&table[h].key, kEmptyKey, ck);
      if (prev == kEmptyKey || prev == ck) break;
      h = (h + 1) & mask;
    }
    if (L < d_open) table[h].start = (uint32_t)j;
    if (L < d_close) table[h].end = (uint32_t)(j + 1);
  }
}

// Scan-sized builds (levels already chosen by index_cluster_kernel). A position opens / closes d - fine cells (d = first
// level at which it differs from a neighbour), anything from 0 to 13, and the per-position loop of table_build_kernel is
// as slow as its longest chain of dependent find-or-inserts. Here row r < kTbRows - 1 of the grid handles the single level
// fine + r of every position (most cells live in the finest levels), and the last row loops over the few coarser ones.
// (Measured alternatives: one row per level for all 13 levels 22 us, warp-pooled (position, level) pairs dealt out
// round-robin 30-40 us, the per-position loop 31 us; the kernel is bound by the latency of the atomics.)
constexpr int kTbRows = 7;
__device__ __forceinline__ void table_insert(CellSlot* __restrict__ table, uint32_t mask, unsigned long long k, int L, int j, bool opens, bool closes) {
  const unsigned long long ck = cell_key(k, L);
  uint32_t h = hash64(ck) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&table[h].key, kEmptyKey, ck);
    if (prev == kEmptyKey || prev == ck) break;
    h = (h + 1) & mask;
  }
  if (opens) table[h].start = (uint32_t)j;
  if (closes) table[h].end = (uint32_t)(j + 1);
}
__global__ void __launch_bounds__(256) table_build_levels_kernel(const unsigned long long* __restrict__ keys, int n, const GridMeta* __restrict__ meta,
                                                                 CellSlot* __restrict__ table, uint32_t mask) {
  const int fine = meta->fine_level;
  const int row = (int)blockIdx.y;
  const int L0 = fine + row;
  if (L0 > kTopLevel) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const unsigned long long k = keys[j];
  const int d_open = j == 0 ? kNumLevels : diff_levels(k, keys[j - 1]);
  const int d_close = j == n - 1 ? kNumLevels : diff_levels(k, keys[j + 1]);
  if (row < kTbRows - 1) {
    if (L0 < d_open || L0 < d_close) table_insert(table, mask, k, L0, j, L0 < d_open, L0 < d_close);
  } else {
    const int d = max(d_open, d_close);
    for (int L = L0; L < d; L++) table_insert(table, mask, k, L, j, L < d_open, L < d_close);
  }
}

inline int ceil_log2(unsigned int v) { int b = 0; while ((1u << b) < v) b++; return b; }

}  // namespace
}  // namespace ngicp

#include "index_cluster.cuh"

namespace ngicp {

void free_index(Index* idx, cudaStream_t stream) {
  if (!idx) return;
  // the freeing stream may not be the one that built (or last adopted) the index: order the free after the build
  order_after_build(idx, stream);
  dev_free(idx->arena, stream);
  if (idx->built) cudaEventDestroy(idx->built);
  delete idx;
}

int build_index(Handle* h, const float* d_xyz, int stride, int n, const int64_t* seg_offsets, int n_seg, Index** out, bool fine) {
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

  const int low_bit = 3 * kSortLevel;
  const int nbits = kMortonBits - low_bit + (n_seg > 1 ? ceil_log2((unsigned)n_seg) : 0);
  const int passes = sort_num_passes(nbits);
  unsigned int cap = 1024;
  // 24-48 B of table per point; with the fine levels (choose_base) 64-128 B
  while (cap < (fine ? 4ull * (unsigned long long)n : (unsigned long long)n + n / 2)) cap <<= 1;
  idx->table_mask = cap - 1;

  // two stream-ordered allocations per build: the index's own arena and one scratch block
  auto align_up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_pts = 0, o_inv = o_pts + align_up(sizeof(float4) * (size_t)n), o_keys = o_inv + align_up(sizeof(int) * (size_t)n),
               o_table = o_keys + align_up(sizeof(unsigned long long) * (size_t)n), o_meta = o_table + align_up(sizeof(CellSlot) * (size_t)cap),
               o_sego = o_meta + align_up(sizeof(GridMeta)), o_segs = o_sego + align_up(sizeof(float4) * (size_t)n_seg),
               arena_bytes = o_segs + align_up(sizeof(int) * ((size_t)n_seg + 1));
  const size_t scratch_elems = sort_scratch_elems(n, nbits);
  const size_t s_keys = 0, s_vals_a = s_keys + align_up(sizeof(unsigned long long) * (size_t)n), s_vals_b = s_vals_a + align_up(sizeof(uint32_t) * (size_t)n),
               s_sort = s_vals_b + align_up(sizeof(uint32_t) * (size_t)n), s_bbox = s_sort + align_up(sizeof(uint32_t) * scratch_elems),
               scratch_bytes = s_bbox + align_up(sizeof(unsigned int) * 6 * (size_t)n_seg);
  char* scratch = nullptr;
  // scan-sized single clouds: the whole front end (bbox .. sorted points + level histogram) is one cluster kernel
  static const int k1_cluster = [] { const char* e = std::getenv("NGICP_K1_CLUSTER"); return e ? std::atoi(e) : 1; }();
  const ClusterShape cl_shape = (k1_cluster && n_seg == 1 && low_bit == 0) ? cluster_shape_for(n) : ClusterShape{0, 0};
  const bool use_cluster = cl_shape.ctas > 0;
#define IDX_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      dev_free(scratch, s); free_index(idx, s);                                                          \
      return fail(h, NGICP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    }                                                                                                    \
  } while (0)
  IDX_CUDA(dev_alloc(&idx->arena, arena_bytes, s));
  if (!use_cluster) IDX_CUDA(dev_alloc(&scratch, scratch_bytes, s));
  idx->pts = reinterpret_cast<float4*>(idx->arena + o_pts);
  idx->inv = reinterpret_cast<int*>(idx->arena + o_inv);
  idx->keys = reinterpret_cast<unsigned long long*>(idx->arena + o_keys);
  idx->table = reinterpret_cast<CellSlot*>(idx->arena + o_table);
  idx->meta = reinterpret_cast<GridMeta*>(idx->arena + o_meta);
  idx->seg_origin = reinterpret_cast<float4*>(idx->arena + o_sego);
  idx->seg_start = reinterpret_cast<int*>(idx->arena + o_segs);
  unsigned long long* keys_tmp = reinterpret_cast<unsigned long long*>(scratch + s_keys);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(scratch + s_vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(scratch + s_vals_b);
  uint32_t* sort_scratch = reinterpret_cast<uint32_t*>(scratch + s_sort);
  unsigned int* lo = reinterpret_cast<unsigned int*>(scratch + s_bbox);
  unsigned int* hi = lo + 3 * n_seg;
  // the sort ping-pongs; start in whichever buffer makes the LAST pass land in the arena
  unsigned long long* keys_a = (passes % 2 == 0) ? idx->keys : keys_tmp;
  unsigned long long* keys_b = (passes % 2 == 0) ? keys_tmp : idx->keys;

  if (n_seg > 1) {
    IDX_CUDA(cudaMemcpyAsync(idx->seg_start, seg32.data(), sizeof(int) * (n_seg + 1), cudaMemcpyHostToDevice, s));
    IDX_CUDA(cudaStreamSynchronize(s));  // seg32 is a stack-lifetime buffer
  }
  IDX_CUDA(cudaMemsetAsync(idx->table, 0xff, sizeof(CellSlot) * (size_t)cap, s));

  const int tpb = 256;
  const int nb = (n + tpb - 1) / tpb;
  const int n_hist = passes * kSortRadix;
  unsigned long long* keys_sorted = nullptr;
  uint32_t* vals_sorted = nullptr;
  if (use_cluster) {
    IDX_CUDA(launch_index_cluster(d_xyz, stride, n, passes, idx->seg_origin, idx->seg_start, idx->meta, idx->keys, idx->pts, idx->inv, cl_shape,
                                  fine ? cap / 2 + cap / 8 : cap / 2, 2, fine, h->fine_occ10, s));
    count_launch(h);
    keys_sorted = idx->keys;
  } else {
  index_prep_kernel<<<(std::max(3 * n_seg, n_hist) + 255) / 256, 256, 0, s>>>(lo, hi, n_seg, sort_scratch, n_hist, n_seg == 1 ? idx->seg_start : nullptr, n);
  bbox_kernel<<<std::min(nb, n_seg == 1 ? 148 : 148 * 8), tpb, 0, s>>>(d_xyz, stride, n, idx->seg_start, n_seg, lo, hi);
  count_launch(h, 2);
  if (n_seg > 1) {
    grid_meta_kernel<<<1, 256, 0, s>>>(lo, hi, n_seg, idx->seg_origin, idx->meta);
    count_launch(h);
  }
  keys_kernel<<<nb, tpb, 0, s>>>(d_xyz, stride, n, idx->seg_start, n_seg, idx->seg_origin, idx->meta, lo, hi, keys_a, vals_a, sort_scratch, low_bit, passes);
  count_launch(h);
  count_launch(h, radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, sort_scratch, n, low_bit, nbits, s, &keys_sorted, &vals_sorted, true));
  if (keys_sorted != idx->keys) {  // n == 1: nothing was sorted
    IDX_CUDA(cudaMemcpyAsync(idx->keys, keys_sorted, sizeof(unsigned long long) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    keys_sorted = idx->keys;
  }
  gather_levels_kernel<<<nb, tpb, 0, s>>>(d_xyz, stride, n, vals_sorted, keys_sorted, idx->pts, idx->inv, idx->meta);
  count_launch(h);
  }
  if (use_cluster) table_build_levels_kernel<<<dim3(nb, kTbRows), tpb, 0, s>>>(keys_sorted, n, idx->meta, idx->table, idx->table_mask);
  else table_build_kernel<<<nb, tpb, 0, s>>>(keys_sorted, n, idx->meta, idx->table, idx->table_mask, fine ? cap / 2 + cap / 8 : cap / 2, 2, fine, h->fine_occ10);
  count_launch(h);
  IDX_CUDA(cudaGetLastError());
  IDX_CUDA(cudaEventCreateWithFlags(&idx->built, cudaEventDisableTiming));
  IDX_CUDA(cudaEventRecord(idx->built, s));
  idx->built_stream = s;
  dev_free(scratch, s);
#undef IDX_CUDA
  *out = idx;
  return NGICP_OK;
}

}  // namespace ngicp
