// Stable LSD radix sort of one scan-sized array inside ONE thread-block cluster (used by the index build, the VoxelGrid
// filter and the time sort of the deskew; <= 131,072 records).
//
// A cluster of up to 16 CTAs holds the records in distributed shared memory: CTA c owns positions [c*TILE, (c+1)*TILE).
// A record is one 64-bit word, (key << 17) | original index, so a pass moves 8 bytes per record and stability needs no
// tie-break. Per pass and CTA:
//   1. every thread holds ITEMS records in registers (warp w owns the contiguous run [w*32*ITEMS, (w+1)*32*ITEMS) so that
//      (cta, warp, round, lane) order is input order); warp-synchronous match_any ranking inside the warp's run: the leader
//      of a digit group bumps a 16-bit per-warp counter and hands the old value to its group (ranks stay in registers)
//   2. per-digit exclusive scan over the 32 warps -> the CTA's digit histogram, scanned over digits -> local starts
//   3. cluster barrier; every digit thread reads the histograms of all CTAs through DSMEM (eight loads in flight): global
//      start of the digit plus the records of that digit in earlier CTAs
//   4. the records are scattered into a CTA-local staging buffer (sorted by digit), then copied out in staging order:
//      consecutive threads write consecutive remote addresses (a digit run of this CTA is a contiguous run at the
//      destination), so the DSMEM stores coalesce instead of being 8-byte random writes
//   5. cluster barrier; reload registers from the incoming buffer
#pragma once
#include <cooperative_groups.h>
#include <cstdlib>

#include "radix_sort.cuh"

namespace ngicp {
namespace {

namespace cg = cooperative_groups;

constexpr int kClThreads = 1024;
constexpr int kClValBits = 17;                          // original index < 131,072 = 16 CTAs x 8192
constexpr int kClMaxCtas = 16;
constexpr int kClMaxDevices = 64;
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
  unsigned int box_lo[3], box_hi[3];                    // this CTA's bounding box (order-preserving integer images)
  unsigned int gbox_lo[3], gbox_hi[3];                  // the cluster's bounding box, folded by every CTA for itself
  unsigned int lvl[16];                                 // this CTA's level histogram
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

// The LSD passes over the records the cluster holds in registers (rec[r] = slot warp*32*ITEMS + r*32 + lane of this CTA;
// `mine` slots are valid). Digits are taken from bit kClValBits upwards. On return rec[] and sm.a hold the sorted records.
template <int kClItems>
__device__ __forceinline__ void cluster_lsd_passes(ClusterSmem<kClItems>& sm, cg::cluster_group& cluster, unsigned long long (&rec)[kClItems], int mine, int passes) {
  constexpr int kClTile = kClThreads * kClItems;
  const int c = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
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
    unsigned mm[kClItems];
#pragma unroll
    for (int r = 0; r < kClItems; r++) mm[r] = __match_any_sync(0xffffffffu, dig[r]);     // independent: all in flight together
#pragma unroll
    for (int r = 0; r < kClItems; r++) {
      const unsigned m = mm[r];
      const int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (dig[r] < kSortRadix && lane == leader) {
        old = sm.warp_cnt[warp][dig[r]];
        sm.warp_cnt[warp][dig[r]] = (unsigned short)(old + __popc(m));
      }
      __syncwarp();                                      // the next round's leaders (other lanes) read what this round's wrote
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

}

// Generic form: sorts n (key, i) pairs by key bits [0, 9 * passes), stable; keys_out[j] = keys_in[order[j]], vals_out[j] = order[j].
// (The index build has its own kernel around the same passes, index_cluster.cuh.)
template <int kClItems>
__global__ void __launch_bounds__(kClThreads, 1) sort_cluster_kernel(const unsigned long long* __restrict__ keys_in, int n, int passes,
                                                                     unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int kClTile = kClThreads * kClItems;
  extern __shared__ __align__(16) unsigned char cl_raw[];
  ClusterSmem<kClItems>& sm = *reinterpret_cast<ClusterSmem<kClItems>*>(cl_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = c * kClTile;
  const int mine = max(0, min(kClTile, n - first));
  unsigned long long rec[kClItems];
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const int i = first + warp * (32 * kClItems) + r * 32 + lane;
    rec[r] = i < n ? ((keys_in[i] << kClValBits) | (unsigned long long)i) : ~0ull;
  }
  cluster.sync();                                        // every CTA of the cluster is resident before the first DSMEM access
  cluster_lsd_passes<kClItems>(sm, cluster, rec, mine, passes);
#pragma unroll
  for (int r = 0; r < kClItems; r++) {
    const int sl = warp * (32 * kClItems) + r * 32 + lane;
    if (sl < mine) {
      const uint32_t i = (uint32_t)(rec[r] & kClValMask);
      keys_out[first + sl] = keys_in[i];                 // the full 64-bit key (the record keeps only its low 47 bits)
      vals_out[first + sl] = i;
    }
  }
  cluster.sync();                                        // nobody leaves while its shared memory may still be written
}

// cluster shape for n records: 4096 per CTA up to 65,536 (16 CTAs), 8192 beyond; 0 CTAs = not available
struct ClusterShape { int ctas; int items; };

template <typename Kernel>
inline bool cluster_launchable(Kernel fn, int ncta, size_t smem, int* cache) {
  if (cache[ncta]) return cache[ncta] > 0;
  bool good = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
  if (good && ncta > 8) good = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
  if (good) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(kClThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    good = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) == cudaSuccess && nclusters >= 1;
  }
  if (!good) cudaGetLastError();
  cache[ncta] = good ? 1 : -1;
  return good;
}

template <typename... Args>
inline cudaError_t cluster_launch(void (*fn)(Args...), ClusterShape shape, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(shape.ctas); cfg.blockDim = dim3(kClThreads); cfg.stream = s; cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = shape.ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, fn, args...);
}

// Sorts (keys_in[i], i) by the low `nbits` key bits with one cluster kernel. Returns false (nothing launched) when the
// size or the device does not allow it: the caller then takes radix_sort_pairs.
inline bool cluster_sort_identity(const unsigned long long* keys_in, int n, int nbits, unsigned long long* keys_out, uint32_t* vals_out, cudaStream_t s,
                                  cudaError_t* err) {
  static const int enabled = [] { const char* e = std::getenv("NGICP_SORT_CLUSTER"); return e ? std::atoi(e) : 1; }();
  static int ok4_dev[kClMaxDevices][kClMaxCtas + 1] = {{0}}, ok8_dev[kClMaxDevices][kClMaxCtas + 1] = {{0}};   // function attributes are per device
  *err = cudaSuccess;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kClMaxDevices) return false;
  int *ok4 = ok4_dev[dev], *ok8 = ok8_dev[dev];
  if (!enabled || n < 2 || n > kClMaxPoints || nbits < 1 || nbits > 64 - kClValBits) return false;
  const int passes = (nbits + kSortRadixBits - 1) / kSortRadixBits;
  const int c4 = (n + kClThreads * 4 - 1) / (kClThreads * 4), c8 = (n + kClThreads * 8 - 1) / (kClThreads * 8);
  if (c4 <= kClMaxCtas && cluster_launchable(sort_cluster_kernel<4>, c4, sizeof(ClusterSmem<4>), ok4)) {
    *err = cluster_launch(sort_cluster_kernel<4>, ClusterShape{c4, 4}, sizeof(ClusterSmem<4>), s, keys_in, n, passes, keys_out, vals_out);
    return true;
  }
  if (c8 <= kClMaxCtas && cluster_launchable(sort_cluster_kernel<8>, c8, sizeof(ClusterSmem<8>), ok8)) {
    *err = cluster_launch(sort_cluster_kernel<8>, ClusterShape{c8, 8}, sizeof(ClusterSmem<8>), s, keys_in, n, passes, keys_out, vals_out);
    return true;
  }
  return false;
}

}  // namespace
}  // namespace ngicp
