// K1 front end for one scan-sized cloud in ONE kernel: bounding box -> grid parameters -> voxel keys -> stable LSD radix
// sort (cluster_sort.cuh) -> Morton-ordered points + inverse permutation + level histogram + choice of levels. Included by
// index.cu (uses its helpers). The reduction of the bounding box (before the passes) and of the level histogram (after the
// last pass) go through CTA 0's shared memory with DSMEM atomics. Replaces index_prep + bbox + keys + 4 x (count, scatter)
// + gather_levels = 12 launches.
#pragma once
#include "cluster_sort.cuh"

namespace ngicp {
namespace {

template <int kClItems>
__global__ void __launch_bounds__(kClThreads, 1) index_cluster_kernel(const float* __restrict__ xyz, int stride, int n, int passes,
                                                                      float4* __restrict__ seg_origin, int* __restrict__ seg_start,
                                                                      GridMeta* __restrict__ meta, unsigned long long* __restrict__ keys_sorted,
                                                                      float4* __restrict__ pts, int* __restrict__ inv, unsigned int max_entries, int occupancy,
                                                                      bool want_fine, int fine_occ10) {
  constexpr int kClTile = kClThreads * kClItems;
  using ClusterSmem = ClusterSmem<kClItems>;
  extern __shared__ __align__(16) unsigned char cl_raw[];
  ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>(cl_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int first = c * kClTile;                         // global position of this CTA's first record
  const int mine = max(0, min(kClTile, n - first));      // records this CTA owns (in every pass: positions are dense)

  // ---- bounding box: registers -> warp -> this CTA's shared memory (local atomics); after one cluster barrier every CTA
  //      folds the boxes of all CTAs itself (6 x ncta DSMEM loads), so nothing funnels through one CTA's shared memory ----
  if (threadIdx.x < 3) { sm.box_lo[threadIdx.x] = 0xffffffffu; sm.box_hi[threadIdx.x] = 0u; sm.gbox_lo[threadIdx.x] = 0xffffffffu; sm.gbox_hi[threadIdx.x] = 0u; }
  if (threadIdx.x < 16) sm.lvl[threadIdx.x] = 0u;
  __syncthreads();
  float px[kClItems], py[kClItems], pz[kClItems];
  unsigned int l[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, u[3] = {0u, 0u, 0u};
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const int i = first + warp * (32 * kClItems) + r * 32 + lane;
    if (i < n) {
      px[r] = xyz[(size_t)i * stride + 0]; py[r] = xyz[(size_t)i * stride + 1]; pz[r] = xyz[(size_t)i * stride + 2];
      const unsigned int ox = f2ord(px[r]), oy = f2ord(py[r]), oz = f2ord(pz[r]);
      l[0] = min(l[0], ox); u[0] = max(u[0], ox);
      l[1] = min(l[1], oy); u[1] = max(u[1], oy);
      l[2] = min(l[2], oz); u[2] = max(u[2], oz);
    } else {
      px[r] = py[r] = pz[r] = 0.f;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      l[a] = min(l[a], __shfl_xor_sync(0xffffffffu, l[a], off));
      u[a] = max(u[a], __shfl_xor_sync(0xffffffffu, u[a], off));
    }
  }
  if (lane < 3) {
    const unsigned int ml = lane == 0 ? l[0] : (lane == 1 ? l[1] : l[2]), mu = lane == 0 ? u[0] : (lane == 1 ? u[1] : u[2]);
    atomicMin(&sm.box_lo[lane], ml);
    atomicMax(&sm.box_hi[lane], mu);
  }
  cluster.sync();                                        // every CTA's box is final (and every CTA of the cluster is resident)
  if (threadIdx.x < 3 * ncta) {
    const ClusterSmem* rs = cluster.map_shared_rank(&sm, threadIdx.x / 3);
    atomicMin(&sm.gbox_lo[threadIdx.x % 3], rs->box_lo[threadIdx.x % 3]);
    atomicMax(&sm.gbox_hi[threadIdx.x % 3], rs->box_hi[threadIdx.x % 3]);
  }
  __syncthreads();
  // ---- grid parameters (every thread from the same six words) and voxel keys, exactly as keys_kernel ----
  const float4 o = make_float4(ord2f(sm.gbox_lo[0]), ord2f(sm.gbox_lo[1]), ord2f(sm.gbox_lo[2]), 0.f);
  const float h0 = cell_size_for(fmaxf(ord2f(sm.gbox_hi[0]) - o.x, fmaxf(ord2f(sm.gbox_hi[1]) - o.y, ord2f(sm.gbox_hi[2]) - o.z)));
  const float inv_h0 = 1.0f / h0;
  if (c == 0 && threadIdx.x == 0) { seg_origin[0] = o; write_meta(meta, h0); seg_start[0] = 0; seg_start[1] = n; }
  unsigned long long rec[kClItems];
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const int i = first + warp * (32 * kClItems) + r * 32 + lane;
    const unsigned int cx = clampi(voxel_coord_unclamped(px[r], o.x, inv_h0), 0, kMaxCoord);
    const unsigned int cy = clampi(voxel_coord_unclamped(py[r], o.y, inv_h0), 0, kMaxCoord);
    const unsigned int cz = clampi(voxel_coord_unclamped(pz[r], o.z, inv_h0), 0, kMaxCoord);
    rec[r] = i < n ? ((morton3(cx, cy, cz) << kClValBits) | (unsigned long long)i) : ~0ull;
  }

  cluster_lsd_passes<kClItems>(sm, cluster, rec, mine, passes);

  // ---- sorted: keys, Morton-ordered points, inverse permutation, level histogram ----
  unsigned int lv[kClItems];
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const int sl = warp * (32 * kClItems) + r * 32 + lane;
    lv[r] = 0;
    if (sl < mine) {
      const int j = first + sl;
      const unsigned long long key = rec[r] >> kClValBits;
      const uint32_t i = (uint32_t)(rec[r] & kClValMask);
      keys_sorted[j] = key;
      const float* q = xyz + (size_t)i * stride;
      pts[j] = make_float4(q[0], q[1], q[2], __int_as_float((int)i));
      inv[i] = j;
      int d = kNumLevels;
      if (j > 0) {
        const unsigned long long prev = sl > 0 ? sm.a[sl - 1] : cluster.map_shared_rank(&sm, c - 1)->a[kClTile - 1];
        d = diff_levels(key, prev >> kClValBits);
      }
      lv[r] = (unsigned int)d;
    }
  }
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const unsigned m = __match_any_sync(0xffffffffu, lv[r]);
    if (lv[r] && lane == __ffs(m) - 1) atomicAdd(&sm.lvl[lv[r]], (unsigned int)__popc(m));     // this CTA's own histogram
  }
  cluster.sync();                                        // every CTA's level histogram is final
  if (c == 0 && threadIdx.x < 16) {
    unsigned int t = 0;
    for (int cc = 0; cc < ncta; cc++) t += cluster.map_shared_rank(&sm, cc)->lvl[threadIdx.x];
    meta->level_hist[threadIdx.x] = t;
  }
  cluster.sync();                                        // nobody leaves while its shared memory may still be read
  if (c == 0) {
    __syncthreads();
    if (threadIdx.x == 0) {                              // the levels the table will hold (table_build_levels_kernel reads them)
      unsigned int total;
      int fine;
      const int b = choose_base(meta, n, max_entries, occupancy, want_fine, fine_occ10, &total, &fine);
      meta->base_level = b; meta->fine_level = fine; meta->cells_total = total;
    }
  }
}

inline ClusterShape cluster_shape_for(int n) {
  static int ok4_dev[kClMaxDevices][kClMaxCtas + 1] = {{0}}, ok8_dev[kClMaxDevices][kClMaxCtas + 1] = {{0}};   // function attributes are per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kClMaxDevices) return {0, 0};
  int *ok4 = ok4_dev[dev], *ok8 = ok8_dev[dev];
  if (n < 2 || n > kClMaxPoints) return {0, 0};
  const int c4 = (n + kClThreads * 4 - 1) / (kClThreads * 4), c8 = (n + kClThreads * 8 - 1) / (kClThreads * 8);
  if (c4 <= kClMaxCtas && cluster_launchable(index_cluster_kernel<4>, c4, sizeof(ClusterSmem<4>), ok4)) return {c4, 4};
  if (c8 <= kClMaxCtas && cluster_launchable(index_cluster_kernel<8>, c8, sizeof(ClusterSmem<8>), ok8)) return {c8, 8};
  return {0, 0};
}

inline cudaError_t launch_index_cluster(const float* xyz, int stride, int n, int passes, float4* seg_origin, int* seg_start, GridMeta* meta,
                                        unsigned long long* keys_sorted, float4* pts, int* inv, ClusterShape shape, unsigned int max_entries, int occupancy,
                                        bool want_fine, int fine_occ10, cudaStream_t s) {
  if (shape.items == 4)
    return cluster_launch(index_cluster_kernel<4>, shape, sizeof(ClusterSmem<4>), s, xyz, stride, n, passes, seg_origin, seg_start, meta, keys_sorted, pts, inv,
                          max_entries, occupancy, want_fine, fine_occ10);
  return cluster_launch(index_cluster_kernel<8>, shape, sizeof(ClusterSmem<8>), s, xyz, stride, n, passes, seg_origin, seg_start, meta, keys_sorted, pts, inv,
                        max_entries, occupancy, want_fine, fine_occ10);
}

}  // namespace
}  // namespace ngicp
