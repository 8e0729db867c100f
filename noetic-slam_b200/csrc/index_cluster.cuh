// K1 front end for one scan-sized cloud in ONE kernel: bounding box -> grid parameters -> voxel keys -> stable LSD radix
// sort -> Morton-ordered points + inverse permutation + level histogram. Included by index.cu (uses its helpers).
//
// A thread-block cluster of up to 16 CTAs holds the whole cloud in distributed shared memory: CTA c owns positions
// [c*8192, (c+1)*8192). A record is one 64-bit word, (key << 17) | original index (36 + 17 bits), so a pass moves 8 bytes
// per point and stability needs no tie-break. Per pass and CTA:
//   1. every thread holds 8 words in registers (warp w owns the contiguous run [w*256, (w+1)*256) so that
//      (cta, warp, round, lane) order is input order), warp-synchronous match_any counting per warp and digit
//   2. per-digit exclusive scan over the 32 warps -> the CTA's digit histogram, scanned over digits -> local starts
//   3. cluster barrier; every digit thread reads the histograms of all CTAs through DSMEM: global start of the digit
//      plus the keys of that digit in earlier CTAs
//   4. stable local rank -> the words are scattered into a CTA-local staging buffer (sorted by digit), then copied out in
//      staging order: consecutive threads write consecutive remote addresses (a digit run of this CTA is a contiguous run at the
//      destination), so the DSMEM stores coalesce instead of being 8-byte random writes
//   5. cluster barrier; reload registers from the incoming buffer
// The reduction of the bounding box (pass 0) and of the level histogram (after the last pass) go through CTA 0's shared
// memory with DSMEM atomics. Replaces index_prep + bbox + keys + 4 x (count, scatter) + gather_levels = 12 launches.
#pragma once
#include <cooperative_groups.h>

namespace ngicp {
namespace {

namespace cg = cooperative_groups;

constexpr int kClThreads = 1024;
constexpr int kClValBits = 17;                          // original index < 131,072 = 16 CTAs x 8192
constexpr int kClMaxCtas = 16;
constexpr int kClMaxPoints = kClMaxCtas * kClThreads * 8;
constexpr unsigned long long kClValMask = (1ull << kClValBits) - 1ull;

template <int ITEMS>
struct ClusterSmem {
  unsigned long long a[kClThreads * ITEMS];             // records owned by this CTA (filled by every CTA of the cluster)
  unsigned long long b[kClThreads * ITEMS];             // local staging, sorted by the digit of the pass
  unsigned short warp_cnt[kClThreads / 32][kSortRadix];
  uint32_t hist[kSortRadix];                            // this CTA's digit counts (read by the whole cluster)
  uint32_t lstart[kSortRadix];                          // exclusive scan of hist over digits
  uint32_t gbase[kSortRadix];                           // destination of this CTA's first record of every digit
  uint32_t half[kSortRadix];                            // keys of the digit in warps 16..31
  uint32_t lower[kSortRadix];                           // keys of the digit in warps 0..15
  uint32_t scan_tmp[32];
  unsigned int box_lo[3], box_hi[3];                    // CTA 0: cluster-wide bounding box (order-preserving integer images)
  unsigned int lvl[16];                                 // CTA 0: cluster-wide level histogram
};

// exclusive scan of one value per thread over the block (all 1024 threads call it)
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* tmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = tmp[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    tmp[lane] = w - tmp[lane];                           // exclusive warp offsets
  }
  __syncthreads();
  const uint32_t r = tmp[warp] + inc - v;
  __syncthreads();                                       // tmp is reused by the next call
  return r;
}

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
  ClusterSmem* sm0 = cluster.map_shared_rank(&sm, 0);
  const int first = c * kClTile;                         // global position of this CTA's first record
  const int mine = max(0, min(kClTile, n - first));      // records this CTA owns (in every pass: positions are dense)

  // ---- bounding box: registers -> warp -> CTA 0 (DSMEM atomics) ----
  if (c == 0 && threadIdx.x < 3) { sm.box_lo[threadIdx.x] = 0xffffffffu; sm.box_hi[threadIdx.x] = 0u; }
  if (c == 0 && threadIdx.x < 16) sm.lvl[threadIdx.x] = 0u;
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
  cluster.sync();                                        // CTA 0's box is initialised
  if (lane < 3) {
    const unsigned int ml = lane == 0 ? l[0] : (lane == 1 ? l[1] : l[2]), mu = lane == 0 ? u[0] : (lane == 1 ? u[1] : u[2]);
    atomicMin(&sm0->box_lo[lane], ml);
    atomicMax(&sm0->box_hi[lane], mu);
  }
  cluster.sync();
  // ---- grid parameters (every thread from the same six words) and voxel keys, exactly as keys_kernel ----
  const float4 o = make_float4(ord2f(sm0->box_lo[0]), ord2f(sm0->box_lo[1]), ord2f(sm0->box_lo[2]), 0.f);
  const float h0 = cell_size_for(fmaxf(ord2f(sm0->box_hi[0]) - o.x, fmaxf(ord2f(sm0->box_hi[1]) - o.y, ord2f(sm0->box_hi[2]) - o.z)));
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

  // ---- LSD passes ----
  for (int p = 0; p < passes; p++) {
    const int shift = kClValBits + p * kSortRadixBits;
    int dig[kClItems];
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      const bool valid = warp * (32 * kClItems) + r * 32 + lane < mine;
      dig[r] = valid ? (int)((rec[r] >> shift) & (kSortRadix - 1)) : (kSortRadix + lane);   // invalid lanes match nobody valid
    }
    {
      uint32_t* row = reinterpret_cast<uint32_t*>(sm.warp_cnt[warp]);
#pragma unroll
      for (int q = 0; q < kSortRadix / 64; q++) row[q * 32 + lane] = 0u;
    }
    __syncwarp();
    // rank inside the warp's run (kept in registers) while counting: the leader of every digit group reads the counter,
    // adds the group size, and hands the old value to the group
    uint32_t lrank[kClItems];
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
      const int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (dig[r] < kSortRadix && lane == leader) {
        old = sm.warp_cnt[warp][dig[r]];
        sm.warp_cnt[warp][dig[r]] = (unsigned short)(old + __popc(m));
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      lrank[r] = old + __popc(m & lt_mask);
    }
    __syncthreads();
    // per digit: exclusive scan over the 32 warps, two halves side by side (threads 512.. take warps 16..31)
    uint32_t total = 0;
    {
      const int d = threadIdx.x & (kSortRadix - 1), w0 = (threadIdx.x >> kSortRadixBits) * 16;
#pragma unroll
      for (int w = 0; w < 16; w++) {
        const unsigned short cnt = sm.warp_cnt[w0 + w][d];
        sm.warp_cnt[w0 + w][d] = (unsigned short)total;
        total += cnt;
      }
      if (w0) sm.half[d] = total;
    }
    __syncthreads();
    if (threadIdx.x < kSortRadix) {
      sm.lower[threadIdx.x] = total;                     // keys of this digit in warps 0..15
      total += sm.half[threadIdx.x];
      sm.hist[threadIdx.x] = total;
    } else {
      total = 0;
    }
    const uint32_t ls = block_excl_scan(total, sm.scan_tmp);
    if (threadIdx.x < kSortRadix) sm.lstart[threadIdx.x] = ls;
    cluster.sync();                                      // every histogram is up; everybody holds its records in registers
    uint32_t all = 0, below = 0;
    if (threadIdx.x < kSortRadix) {
#pragma unroll 1
      for (int c0 = 0; c0 < ncta; c0 += 8) {             // eight remote loads in flight
        uint32_t v[8];
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = c0 + q < ncta ? cluster.map_shared_rank(&sm, c0 + q)->hist[threadIdx.x] : 0u;
#pragma unroll
        for (int q = 0; q < 8; q++) { all += v[q]; below += c0 + q < c ? v[q] : 0u; }
      }
    }
    const uint32_t gs = block_excl_scan(all, sm.scan_tmp);
    if (threadIdx.x < kSortRadix) sm.gbase[threadIdx.x] = gs + below;
    // stable local rank -> staging buffer
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      if (dig[r] < kSortRadix) {
        const uint32_t lr = sm.lstart[dig[r]] + sm.warp_cnt[warp][dig[r]] + (warp >= 16 ? sm.lower[dig[r]] : 0u) + lrank[r];
        sm.b[lr] = rec[r];
      }
    }
    __syncthreads();
    // copy out in staging order
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      const int e = r * kClThreads + threadIdx.x;
      if (e < mine) {
        const unsigned long long w = sm.b[e];
        const int d = (int)((w >> shift) & (kSortRadix - 1));
        const uint32_t pos = sm.gbase[d] + ((uint32_t)e - sm.lstart[d]);
        cluster.map_shared_rank(&sm, pos / kClTile)->a[pos % kClTile] = w;
      }
    }
    cluster.sync();                                      // the incoming buffer is complete
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      const int sl = warp * (32 * kClItems) + r * 32 + lane;
      rec[r] = sl < mine ? sm.a[sl] : ~0ull;
    }
  }

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
    if (lv[r] && lane == __ffs(m) - 1) atomicAdd(&sm0->lvl[lv[r]], (unsigned int)__popc(m));
  }
  cluster.sync();                                        // nobody leaves while its shared memory may still be read
  if (c == 0 && threadIdx.x < 16) meta->level_hist[threadIdx.x] = sm.lvl[threadIdx.x];
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

// cluster shape for n points: 4096 records per CTA up to 65,536 points (16 CTAs), 8192 beyond; 0 CTAs = not available
struct ClusterShape { int ctas; int items; };

template <int ITEMS>
inline bool cluster_shape_ok(int ncta) {
  static int ok[kClMaxCtas + 1] = {0};       // 0 unknown, 1 yes, -1 no
  if (ok[ncta]) return ok[ncta] > 0;
  auto* fn = index_cluster_kernel<ITEMS>;
  bool good = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ClusterSmem<ITEMS>)) == cudaSuccess;
  if (good && ncta > 8) good = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (good) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(kClThreads); cfg.dynamicSmemBytes = sizeof(ClusterSmem<ITEMS>);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    good = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) == cudaSuccess && nclusters >= 1;
  }
  if (!good) cudaGetLastError();
  ok[ncta] = good ? 1 : -1;
  return good;
}

inline ClusterShape cluster_shape_for(int n) {
  if (n < 2 || n > kClMaxPoints) return {0, 0};
  const int c4 = (n + kClThreads * 4 - 1) / (kClThreads * 4), c8 = (n + kClThreads * 8 - 1) / (kClThreads * 8);
  if (c4 <= kClMaxCtas && cluster_shape_ok<4>(c4)) return {c4, 4};
  if (c8 <= kClMaxCtas && cluster_shape_ok<8>(c8)) return {c8, 8};
  return {0, 0};
}

inline cudaError_t launch_index_cluster(const float* xyz, int stride, int n, int passes, float4* seg_origin, int* seg_start, GridMeta* meta,
                                        unsigned long long* keys_sorted, float4* pts, int* inv, ClusterShape shape, unsigned int max_entries, int occupancy,
                                        bool want_fine, int fine_occ10, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(shape.ctas); cfg.blockDim = dim3(kClThreads); cfg.stream = s;
  cfg.dynamicSmemBytes = shape.items == 4 ? sizeof(ClusterSmem<4>) : sizeof(ClusterSmem<8>);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = shape.ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (shape.items == 4)
    return cudaLaunchKernelEx(&cfg, index_cluster_kernel<4>, xyz, stride, n, passes, seg_origin, seg_start, meta, keys_sorted, pts, inv, max_entries, occupancy,
                              want_fine, fine_occ10);
  return cudaLaunchKernelEx(&cfg, index_cluster_kernel<8>, xyz, stride, n, passes, seg_origin, seg_start, meta, keys_sorted, pts, inv, max_entries, occupancy,
                            want_fine, fine_occ10);
}

}  // namespace
}  // namespace ngicp
