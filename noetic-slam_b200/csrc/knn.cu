// K2 — exact k-nearest-neighbour search over the Morton-sorted voxel hash.
// Replaces nanoflann::KdTreeFLANN::nearestKSearch -> findNeighbors -> searchLevel
// (reference src/dlio/include/nano_gicp/nanoflann_adaptor.h:141-152, nanoflann.h:1436-1460,
// :1587-1666) and the k-NN call of calculate_covariances (src/nano_gicp/nano_gicp.cc:343).
// One thread per query; queries run in Morton order so that a warp walks the same few cells.
// Distances are the reference's fp32 metric bit for bit (common.cuh: sqdist_ref); result rows are
// ordered by (distance, original index).
#include <cstdio>
#include <cstdlib>

#include "internal.h"
#include "wknn.cuh"
#include "lknn.cuh"

namespace ngicp {

namespace {

// original index -> sorted position (what K3 gathers with); -1 stays -1
__device__ __forceinline__ int topos(const GridView& g, int orig) { return orig >= 0 ? __ldg(g.inv + orig) : -1; }

// where chunk c (neighbours 4c..4c+3) of point j goes: tiled for the k K3 reads with compile-time loops (internal.h)
template <int K>
__device__ __forceinline__ int4* nbr_chunk(int* __restrict__ nbr, int j, int c, bool tiled) {
  return tiled ? reinterpret_cast<int4*>(nbr) + ((size_t)(j >> 5) * (K / 4) + c) * 32 + (j & 31)
               : reinterpret_cast<int4*>(nbr + (size_t)j * K) + c;
}

template <class TK>
__device__ __forceinline__ void self_query(const GridView& g, int j, int k, int start_count, int normalization,
                                           int* __restrict__ nbr, double* __restrict__ dens_term, TK& best) {
  const float4 q = __ldg(g.pts + j);
  const int seg = find_segment(g.seg_start, g.n_seg, j);
  grid_knn(g, q.x, q.y, q.z, seg, k, start_count, __int_as_float(0x7f800000), best);
}

// Rows of the tiled table keep neighbour 0 (the query itself) first and the other k-1 in ASCENDING SORTED POSITION, not
// distance order: K3 sums over the set, and lane-adjacent (Morton-adjacent) queries then gather from nearby addresses with
// the same instruction — 9.0 instead of 11.7 distinct 128-byte lines per warp gather on an OS1-64 scan, which is what
// bounds K3 (L1 wavefronts). Bitonic network on registers, padded with INT_MAX.
template <int N>
__device__ __forceinline__ void sort_ascending(int (&a)[N]) {
  static_assert((N & (N - 1)) == 0, "power of two");
#pragma unroll
  for (int k = 2; k <= N; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int l = i ^ j;
        if (l > i) {
          const int lo = min(a[i], a[l]), hi = max(a[i], a[l]);
          if ((i & k) == 0) { a[i] = lo; a[l] = hi; } else { a[i] = hi; a[l] = lo; }
        }
      }
}
template <int K, class TK>
__device__ __forceinline__ void write_tiled_row(const GridView& g, const TK& best, int* __restrict__ nbr, int j) {
  constexpr int N = K <= 16 ? 16 : 32;
  int a[N];
#pragma unroll
  for (int i = 0; i < N; i++) a[i] = (i + 1 < K) ? topos(g, best.p[i + 1]) : 0x7fffffff;
  sort_ascending(a);
  const int self = topos(g, best.p[0]);
#pragma unroll
  for (int c = 0; c < K / 4; c++) {
    const int4 v = c == 0 ? make_int4(self, a[0], a[1], a[2]) : make_int4(a[4 * c - 1], a[4 * c], a[4 * c + 1], a[4 * c + 2]);
    *nbr_chunk<K>(nbr, j, c, true) = v;
  }
}

#ifdef NGICP_STATS
__device__ unsigned long long g_k2_items[4 * 16384];   // development: start / end (globaltimer, ns) of every K2 work item
#endif

// Production K2: a warp owns 32/LPQ consecutive (Morton-sorted) points, LPQ lanes per point, shared staged
// candidates (wknn.cuh).
constexpr int kSelfWarps = 1;
template <int K, int LPQ>
__global__ void __launch_bounds__(32 * kSelfWarps) knn_self_warp_kernel(GridView g, int k, int cmax, int normalization, bool tiled,
                                                                        int* __restrict__ nbr, double* __restrict__ dens_term) {
  __shared__ WarpScratch scratch[kSelfWarps];
  const int j = blockIdx.x * (blockDim.x / LPQ) + threadIdx.x / LPQ;
  const bool active = j < g.n;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  int seg = 0;
  if (active) { q = __ldg(g.pts + j); seg = find_segment(g.seg_start, g.n_seg, j); }
  TopK<K> best;
  uint32_t phase = wknn_init(scratch[threadIdx.x >> 5]);
#ifdef NGICP_STATS
  unsigned long long t_start; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
  warp_knn<LPQ>(g, active, q.x, q.y, q.z, seg, k, cmax, __int_as_float(0x7f800000), best, scratch[threadIdx.x >> 5], phase);
#ifdef NGICP_STATS
  if (threadIdx.x == 0 && blockIdx.x < 16384) {
    unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    g_k2_items[4 * blockIdx.x] = t_start; g_k2_items[4 * blockIdx.x + 1] = t_end;   // indexed by launch position
    g_k2_items[4 * blockIdx.x + 2] = ((unsigned long long)scratch[0].pad[0] << 32) | scratch[0].pad[2]; g_k2_items[4 * blockIdx.x + 3] = scratch[0].pad[1];
  }
#endif
  if (!active || (threadIdx.x & (LPQ - 1)) != 0) return;
  int* row = nbr + (size_t)j * k;
  if ((K == 16 || K == 20) && k == K && tiled) {
    write_tiled_row<K>(g, best, nbr, j);
  } else if (k == K && (K % 4) == 0) {
#pragma unroll
    for (int i = 0; i < K; i += 4)
      *nbr_chunk<K>(nbr, j, i / 4, tiled) = make_int4(topos(g, best.p[i]), topos(g, best.p[i + 1]), topos(g, best.p[i + 2]), topos(g, best.p[i + 3]));
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) if (i < k) row[i] = topos(g, best.p[i]);
  }
  if (dens_term) {
    double acc = 0.0;
#pragma unroll
    for (int i = 1; i < K; i++) if (i < k) acc += (double)best.d[i];
    dens_term[j] = acc / (double)normalization;
  }
}


// ---- leaf-scheduled K2 (lknn.cuh) ---------------------------------------------------------------------------------
// Work items of the search: one thread per sorted position; the heads of the finest grouping cells (level base+1) climb
// to their leaf = the largest ancestor with <= cmax points (the finest grouping cell itself if even that holds more) and
// the first point of a leaf appends ceil(members / 32) items. ctr[0] = number of items (see knn_leaf_kernel for the other counters).
__global__ void __launch_bounds__(256) leaf_items_kernel(GridView g, const unsigned long long* __restrict__ keys, int cmax, int cap2,
                                                         LeafItem* __restrict__ items, unsigned int* __restrict__ ctr) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int nit = 0, P = 0;
  uint32_t s = 0, e = 0;
  if (j < g.n) {
    const int base = __ldg(&g.meta->fine_level);     // the finest level in the table
    const unsigned long long key = __ldg(keys + j);
    const int sh = 3 * (base + 1);
    if (j == 0 || (__ldg(keys + j - 1) >> sh) != (key >> sh)) {          // the first point of its finest grouping cell
      // leaf = the largest ancestor P (from the finest grouping cell up) with <= cmax points whose parent holds <= cap2: the
      // block staged for a leaf is about as populated as the leaf's parent and its neighbours, so a sparse cell inside a
      // dense neighbourhood is taken at a finer level instead of dragging thousands of candidates through one warp. Both
      // counts grow with the level, so every point of a leaf finds the same P.
      P = min(base + 1, kTopLevel);
      if (cell_lookup(g.table, g.table_mask, cell_key(key, P), s, e)) {   // (always: every cell >= base is in the table)
        uint32_t s1 = 0, e1 = 0x7fffffffu;                                // the parent of the current candidate level
        bool have1 = P < kTopLevel && cell_lookup(g.table, g.table_mask, cell_key(key, P + 1), s1, e1);
        while (P < kTopLevel && have1 && (int)(e1 - s1) <= cmax) {
          uint32_t s2 = 0, e2 = 0;
          const bool have2 = P + 2 <= kTopLevel && cell_lookup(g.table, g.table_mask, cell_key(key, P + 2), s2, e2);
          if (have2 && (int)(e2 - s2) > cap2) break;
          P++; s = s1; e = e1;
          s1 = s2; e1 = e2; have1 = have2;
        }
        if ((int)s == j) nit = ((int)(e - s) + 31) / 32;                  // the leaf belongs to its first point
      }
    }
  }
  // one atomic per warp: exclusive prefix of the item counts of its lanes
  int inc = nit;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += a;
  }
  const int total = __shfl_sync(0xffffffffu, inc, 31);
  unsigned int at = 0;
  if (total > 0) {
    if (lane == 31) at = atomicAdd(&ctr[0], (unsigned int)total);
    at = __shfl_sync(0xffffffffu, at, 31) + (unsigned int)(inc - nit);
    const int mcount = (int)(e - s);
    for (int c = 0; c < nit; c++) {
      LeafItem it;
      it.start = (int)s + 32 * c;
      it.count_level = (min(32, mcount - 32 * c) << 8) | P;
      items[at + c] = it;
    }
  }
  // the last block to finish copies the item count to where the search keeps its queue state (knn_leaf_kernel)
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(&ctr[7], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    const unsigned int n1 = *reinterpret_cast<volatile unsigned int*>(&ctr[0]);
    ctr[1] = n1; ctr[4] = n1; ctr[6] = n1; ctr[7] = 0u;
  }
}

constexpr int kLeafWarps = 4;
#ifndef NGICP_K2_MINB
#define NGICP_K2_MINB 5      // CTAs per SM the leaf search is compiled for (register cap 96: at 80 the compiler rebuilds addresses from special registers inside the scan loop)
#endif
// Persistent warps over a work queue that the search itself may extend (leaf_knn_item re-queues the members of very
// heavy blocks as smaller items). ctr: [0] items published and [1] next index of the overflow items (those published during
// this launch, behind the n1 initial ones) — one 64-bit word, read with one load; [2] next ticket of the initial items,
// [3] warps that left, [4] slots reserved, [6] n1, [7] block counter of leaf_items_kernel. Overflow items come first (they
// are the pieces of the heaviest blocks: started late they would be the tail of the launch) and are drawn with a
// compare-and-swap, so that an index is only ever consumed when its item exists; initial items with a plain atomicAdd.
// Nobody waits: a warp leaves when both queues are empty at the moment it looks, and a warp that publishes items looks
// again afterwards, so whatever nobody else picked up it processes itself. The last warp to leave resets the counters, so
// the next k-NN call on this handle finds them zero without a memset.
template <int KP, int KC, int C>
__global__ void __launch_bounds__(32 * kLeafWarps, NGICP_K2_MINB) knn_leaf_kernel(GridView g, int k_rt, LeafItem* __restrict__ items, unsigned int capacity, unsigned int* __restrict__ ctr,
                                                                   int normalization, bool tiled, int* __restrict__ nbr, double* __restrict__ dens_term,
                                                                   unsigned long long* __restrict__ trace, int use_tma) {
  unsigned long long t_enter = 0, t_first = 0, t_last = 0;
  if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_enter));
  extern __shared__ __align__(16) unsigned char leaf_smem[];     // kLeafWarps x LeafScratch (dynamic: the big lists exceed 48 KB)
  LeafScratch<KP, C>& ws = reinterpret_cast<LeafScratch<KP, C>*>(leaf_smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int k = KC > 0 ? KC : k_rt;
  if (lane == 0) { mbar_init(&ws.mbar[0], 1); mbar_init(&ws.mbar[1], 1); ws.use_tma = use_tma; }
#pragma unroll
  for (int i = 0; i < kLeafPend; i++) ws.pend[i][lane] = kKeyEmpty;
  __syncwarp();
  uint32_t phase = 0u;
  volatile unsigned int* vctr = ctr;
  const unsigned int n1 = vctr[6];
  bool initial = true;
  for (;;) {
    unsigned int it = 0;
    bool got = false;
    if (lane == 0) {
      for (;;) {
        const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(ctr);
        const unsigned int published = (unsigned int)w, t2 = (unsigned int)(w >> 32);
        if (t2 < published) {                                       // an overflow item nobody has taken
          if (atomicCAS(&ctr[1], t2, t2 + 1u) == t2) { it = t2; got = true; break; }
          continue;
        }
        if (initial) { it = atomicAdd(&ctr[2], 1u); got = it < n1; initial = got; }
        if (got || !initial) break;
      }
    }
    got = __shfl_sync(0xffffffffu, got ? 1 : 0, 0) != 0;
    it = __shfl_sync(0xffffffffu, it, 0);
    if (!got) break;
    if (trace && t_first == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_first));
    LeafItem item;     // published by another SM during this launch, possibly: read through L2
    {
      const int2 raw = __ldcg(reinterpret_cast<const int2*>(items + it));
      item.start = raw.x; item.count_level = raw.y;
    }
#ifdef NGICP_STATS
    unsigned long long t_start; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    unsigned int* cur = g_leaf_cur[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & 8191];
    if (lane == 0) { for (int i = 0; i < 8; i++) cur[i] = 0; }
    __syncwarp();
#endif
    leaf_knn_item<KP, KC, C>(g, item, k_rt, ws, phase, [&](int j, int (&a)[KP < 16 ? 16 : KP], double dsum) {
      sort_ascending(a);
#pragma unroll
      for (int i = 0; i < KP; i++) if (i < k - 1 && a[i] == 0x7fffffff) a[i] = j;   // fewer than k points in reach: pad with the point itself
      if (KC > 0 && (KC % 4) == 0 && tiled) {
#pragma unroll
        for (int c = 0; c < KC / 4; c++) {
          const int4 v = c == 0 ? make_int4(j, a[0], a[1], a[2]) : make_int4(a[4 * c - 1], a[4 * c], a[4 * c + 1], a[4 * c + 2]);
          *nbr_chunk<KC>(nbr, j, c, true) = v;
        }
      } else {
        int* row = nbr + (size_t)j * k;
        row[0] = j;
#pragma unroll
        for (int i = 0; i + 1 < KP; i++) if (i + 1 < k) row[i + 1] = a[i];
      }
      if (dens_term) dens_term[j] = dsum / (double)normalization;
    }, [&](unsigned heads, int level) {
      // one new item per run of lanes; slots are reserved with one atomic, filled, and published in reservation order
      const unsigned FULL = 0xffffffffu;
      const int groups = __popc(heads);
      unsigned int at = 0;
      if (lane == 0) at = atomicAdd(&ctr[4], (unsigned int)groups);
      at = __shfl_sync(FULL, at, 0);
      if (at + (unsigned int)groups > capacity) {                 // no room: give the slots back in order and process the item as it is
        if (lane == 0) { while (vctr[0] != at) __nanosleep(100); atomicSub(&ctr[4], (unsigned int)groups); }
        __syncwarp();
        return false;
      }
      if ((heads >> lane) & 1u) {
        const unsigned above = heads & ~((2u << lane) - 1u);
        const int end = above ? __ffs(above) - 1 : (item.count_level >> 8);
        LeafItem ni;
        ni.start = item.start + lane;
        ni.count_level = ((end - lane) << 8) | level;
        items[at + __popc(heads & ((1u << lane) - 1u))] = ni;
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        while (vctr[0] != at) __nanosleep(100);                   // publish in reservation order
        atomicAdd(&ctr[0], (unsigned int)groups);
      }
      __syncwarp();
      return true;
    });
#ifdef NGICP_STATS
    if (lane == 0 && it < 8192) {
      unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      g_leaf_items[4 * it] = t_start; g_leaf_items[4 * it + 1] = t_end;
      g_leaf_items[4 * it + 2] = ((unsigned long long)(item.count_level & 0xff) << 56) | ((unsigned long long)(item.count_level >> 8) << 48) |
                                 ((unsigned long long)min(cur[0], 255u) << 40) | ((unsigned long long)min(cur[3], 255u) << 32) | (unsigned long long)cur[1];
      g_leaf_items[4 * it + 3] = ((unsigned long long)cur[2] << 32) | (unsigned long long)(unsigned)item.start;
      if (it < 4096) { g_leaf_phase[4 * it] = cur[4]; g_leaf_phase[4 * it + 1] = cur[5]; g_leaf_phase[4 * it + 2] = cur[6]; g_leaf_phase[4 * it + 3] = cur[7]; }
    }
#endif
    __syncwarp();
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_last));
  }
  if (trace && lane == 0) {     // development: [0] first warp entry, [1] last warp exit, [2] first item start, [3] last item end
    unsigned long long t_exit; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
    atomicMin(&trace[0], t_enter); atomicMax(&trace[1], t_exit);
    if (t_first) { atomicMin(&trace[2], t_first); atomicMax(&trace[3], t_last); }
  }
  if (lane == 0) {
    const unsigned int total = gridDim.x * kLeafWarps;
    if (atomicAdd(&ctr[3], 1u) == total - 1u) { ctr[0] = 0u; ctr[1] = 0u; ctr[2] = 0u; ctr[4] = 0u; ctr[6] = 0u; __threadfence(); ctr[3] = 0u; }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(64) knn_self_dyn_kernel(GridView g, int k, int start_count, int normalization,
                                                         int* __restrict__ nbr, double* __restrict__ dens_term) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  TopKDyn<KMAX> best;
  best.cap = k;
  self_query(g, j, k, start_count, normalization, nbr, dens_term, best);
  int* row = nbr + (size_t)j * k;
  for (int i = 0; i < k; i++) row[i] = topos(g, best.p[i]);
  if (dens_term) {
    double acc = 0.0;
    for (int i = 1; i < k; i++) acc += (double)best.d[i];
    dens_term[j] = acc / (double)normalization;
  }
}

template <class TK>
__device__ __forceinline__ void write_public(const GridView& g, const TK& best, int k, int cap, int* __restrict__ oi, float* __restrict__ od) {
  for (int i = 0; i < k; i++) {
    int p = -1; float d = __int_as_float(0x7f800000);
    if (i < cap) { p = best.p[i]; d = best.d[i]; }
    oi[i] = p;
    od[i] = p >= 0 ? d : __int_as_float(0x7f800000);
  }
}

template <int K>
__global__ void __launch_bounds__(128) knn_query_kernel(GridView g, const float4* __restrict__ q, int nq, int k, int start_count,
                                                        int* __restrict__ out_idx, float* __restrict__ out_sqd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float4 qq = __ldg(q + i);
  TopK<K> best;
  grid_knn(g, qq.x, qq.y, qq.z, 0, k, start_count, __int_as_float(0x7f800000), best);
  int* oi = out_idx + (size_t)i * k;
  float* od = out_sqd + (size_t)i * k;
#pragma unroll
  for (int t = 0; t < K; t++) {
    if (t < k) {
      const int p = best.p[t];
      oi[t] = p;
      od[t] = p >= 0 ? best.d[t] : __int_as_float(0x7f800000);
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(64) knn_query_dyn_kernel(GridView g, const float4* __restrict__ q, int nq, int k, int start_count,
                                                          int* __restrict__ out_idx, float* __restrict__ out_sqd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const float4 qq = __ldg(q + i);
  TopKDyn<KMAX> best;
  best.cap = k;
  grid_knn(g, qq.x, qq.y, qq.z, 0, k, start_count, __int_as_float(0x7f800000), best);
  write_public(g, best, k, k, out_idx + (size_t)i * k, out_sqd + (size_t)i * k);
}

constexpr int kMaxK = 128;
// inspection export (ngicp_self_neighbours): K2's table (sorted positions, tiled or row-major) -> rows of ORIGINAL indices in
// original point order; entry 0 of a row is the point itself
__global__ void __launch_bounds__(256) export_self_rows_kernel(GridView g, const int* __restrict__ nbr, int k, bool tiled, int* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  const int orig = __float_as_int(__ldg(&g.pts[j].w));
  for (int i = 0; i < k; i++) {
    const int pos = tiled ? nbr[(((size_t)(j >> 5) * (k / 4) + (i >> 2)) * 32 + (j & 31)) * 4 + (i & 3)] : nbr[(size_t)j * k + i];
    out[(size_t)orig * k + i] = (pos >= 0 && pos < g.n) ? __float_as_int(__ldg(&g.pts[pos].w)) : -1;
  }
}

inline int start_count_for(int k) { return k < 4 ? 1 : (k + 3) / 4; }
// group-cell capacity of the warp-cooperative search: large enough that the k-th neighbour of a member
// lies inside the staged block (block reach >= half the group cell), small enough to keep the scan short
inline int group_cap_for(int k, int mult) { return k * mult > 32 ? k * mult : 32; }

}  // namespace


// persistent grid of the leaf search: as many CTAs as the device holds at once (queried once per instantiation)
template <int KP, int KC, int C>
static int launch_leaf_search(Handle* h, const Index* idx, int k, int normalization, LeafItem* items, unsigned int capacity, int* d_nbr, double* d_dens_term,
                              unsigned long long* d_trace) {
  static int per_sm = 0;
  constexpr size_t smem = sizeof(LeafScratch<KP, C>) * kLeafWarps;
  if (per_sm == 0) {
    int v = 0;
    NGICP_CUDA(h, cudaFuncSetAttribute(knn_leaf_kernel<KP, KC, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NGICP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, knn_leaf_kernel<KP, KC, C>, 32 * kLeafWarps, smem));
    per_sm = v > 0 ? v : 1;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const int want = (idx->n / 16) / kLeafWarps + 1;     // an item is ~23 points on average
  const int grid = std::max(1, std::min(sms * per_sm, want));
  knn_leaf_kernel<KP, KC, C><<<grid, 32 * kLeafWarps, smem, h->stream>>>(idx->view(), k, items, capacity, h->k2_ctr, normalization, nbr_tiled(k), d_nbr, d_dens_term, d_trace, h->k2_tma);
  return NGICP_OK;
}

int knn_self(Handle* h, const Index* idx, int k, int* d_nbr, double* d_dens_term) {
  if (k < 1 || k > kMaxK) return fail(h, NGICP_ERR_UNSUPPORTED, "k must be in [1,128]");
  const GridView g = idx->view();
  const int n = idx->n;
  const int normalization = ((k - 1) * (2 + k)) / 2;  // integer arithmetic, nano_gicp.cc:345
  const int sc = start_count_for(k);
  cudaStream_t s = h->stream;
  if (k <= 32 && h->k2_leaf) {
    // leaf-scheduled search (lknn.cuh): work items first, then the persistent search over them
    if (!h->k2_ctr) {
      NGICP_CUDA(h, cudaMalloc(&h->k2_ctr, sizeof(unsigned int) * 8));
      NGICP_CUDA(h, cudaMemsetAsync(h->k2_ctr, 0, sizeof(unsigned int) * 8, s));
    }
    LeafItem* items = nullptr;
    const size_t capacity = (size_t)n + (size_t)n / 32 + (size_t)n / 4 + 64;     // leaves + their 32-point chunks + room for re-queued splits
    NGICP_CUDA(h, dev_alloc(&items, capacity, s));
    static const bool trace = std::getenv("NGICP_K2_TRACE") != nullptr;      // development: split of the two launches
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (trace) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], s); }
    leaf_items_kernel<<<(n + 255) / 256, 256, 0, s>>>(g, idx->keys, group_cap_for(k, h->k2_cmax_mult), group_cap_for(k, h->k2_cmax_mult) * h->k2_cap2_mult, items,
                                                    h->k2_ctr);
    unsigned long long* d_trace = nullptr;
    if (trace) {
      cudaEventRecord(ev[1], s);
      cudaMalloc(&d_trace, 32);
      const unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
      cudaMemcpyAsync(d_trace, init, 32, cudaMemcpyHostToDevice, s);
      cudaEventRecord(ev[1], s);
    }
    int rc;
#define LEAF(KP, KC) (h->k2_chunk == 128 ? launch_leaf_search<KP, KC, 128>(h, idx, k, normalization, items, (unsigned int)capacity, d_nbr, d_dens_term, d_trace) \
                                         : launch_leaf_search<KP, KC, 256>(h, idx, k, normalization, items, (unsigned int)capacity, d_nbr, d_dens_term, d_trace))
    if (k == 16) rc = LEAF(16, 16);
    else if (k == 20) rc = LEAF(32, 20);
    else if (k <= 8) rc = LEAF(8, 0);
    else if (k <= 16) rc = LEAF(16, 0);
    else rc = LEAF(32, 0);
#undef LEAF
    if (trace) {
      cudaEventRecord(ev[2], s);
      cudaEventSynchronize(ev[2]);
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
      unsigned long long t[4];
      cudaMemcpy(t, d_trace, sizeof t, cudaMemcpyDeviceToHost);
      cudaFree(d_trace);
      std::fprintf(stderr, "[k2 trace] n %d leaf_items %.3f ms search %.3f ms (events); in-kernel: enter..exit %.1f us, first item at +%.1f us, last item end at +%.1f us\n", n, a, b,
                   (t[1] - t[0]) * 1e-3, (t[2] - t[0]) * 1e-3, (t[3] - t[0]) * 1e-3);
      for (auto& e : ev) cudaEventDestroy(e);
    }
    dev_free(items, s);
    if (rc) return rc;
    count_launch(h, 2);
    NGICP_CUDA(h, cudaGetLastError());
    return NGICP_OK;
  }
  // lanes per query: small clouds need the extra warps, big ones prefer the sharing of one query per lane
  const int lpq = h->k2_lpq > 0 ? h->k2_lpq : (n < 1500000 ? 4 : 1);
#define LAUNCH_SELF_L(K, LPQ)                                                                                             \
  knn_self_warp_kernel<K, LPQ><<<(n + (32 * kSelfWarps / LPQ) - 1) / (32 * kSelfWarps / LPQ), 32 * kSelfWarps, 0, s>>>( \
      g, k, group_cap_for(k, h->k2_cmax_mult), normalization, nbr_tiled(k), d_nbr, d_dens_term)
#define LAUNCH_SELF(K) do { if (lpq >= 4) LAUNCH_SELF_L(K, 4); else LAUNCH_SELF_L(K, 1); } while (0)
  if (k == 1) LAUNCH_SELF(1);
  else if (k <= 8) LAUNCH_SELF(8);
  else if (k <= 16) LAUNCH_SELF(16);
  else if (k <= 20) LAUNCH_SELF(20);
  else if (k <= 32) LAUNCH_SELF(32);
  else knn_self_dyn_kernel<kMaxK><<<(n + 63) / 64, 64, 0, s>>>(g, k, sc, normalization, d_nbr, d_dens_term);
#undef LAUNCH_SELF
#undef LAUNCH_SELF_L
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

int export_self_rows(Handle* h, const Index* idx, const int* d_nbr, int k, int* d_out) {
  export_self_rows_kernel<<<(idx->n + 255) / 256, 256, 0, h->stream>>>(idx->view(), d_nbr, k, nbr_tiled(k), d_out);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

int knn_queries(Handle* h, const Index* idx, const float4* d_q, int nq, int k, int* d_out_idx, float* d_out_sqd) {
  if (k < 1 || k > kMaxK) return fail(h, NGICP_ERR_UNSUPPORTED, "k must be in [1,128]");
  if (idx->n_seg != 1) return fail(h, NGICP_ERR_UNSUPPORTED, "public k-NN queries need a single-segment index");
  if (nq <= 0) return NGICP_OK;
  const GridView g = idx->view();
  const int sc = start_count_for(k);
  cudaStream_t s = h->stream;
#define LAUNCH_Q(K) knn_query_kernel<K><<<(nq + 127) / 128, 128, 0, s>>>(g, d_q, nq, k, sc, d_out_idx, d_out_sqd)
  if (k == 1) LAUNCH_Q(1);
  else if (k <= 8) LAUNCH_Q(8);
  else if (k <= 16) LAUNCH_Q(16);
  else if (k <= 20) LAUNCH_Q(20);
  else if (k <= 32) LAUNCH_Q(32);
  else knn_query_dyn_kernel<kMaxK><<<(nq + 63) / 64, 64, 0, s>>>(g, d_q, nq, k, sc, d_out_idx, d_out_sqd);
#undef LAUNCH_Q
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

}  // namespace ngicp

#ifdef NGICP_STATS
extern "C" int ngicp_debug_stats_leaf(unsigned long long out[16], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_leaf_stats, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ngicp::g_leaf_stats, z, sizeof z); }
  return 0;
}
extern "C" int ngicp_debug_leaf_phase(unsigned int* out, int n_items) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_leaf_phase, sizeof(unsigned int) * 4 * n_items);
  return 0;
}
extern "C" int ngicp_debug_leaf_items(unsigned long long* out, int n_items) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_leaf_items, sizeof(unsigned long long) * 4 * n_items);
  return 0;
}
extern "C" int ngicp_debug_leaf_records(unsigned int* out /* 64 x 24 */, unsigned int* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_leaf_dbg, sizeof(unsigned int) * 64 * 24);
  cudaMemcpyFromSymbol(n, ngicp::g_leaf_dbg_n, sizeof(unsigned int));
  unsigned int z = 0; cudaMemcpyToSymbol(ngicp::g_leaf_dbg_n, &z, sizeof z);
  return 0;
}
extern "C" int ngicp_debug_items_knn(unsigned long long* out, int n_items) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_k2_items, sizeof(unsigned long long) * 4 * n_items);
  return 0;
}
extern "C" int ngicp_debug_stats_knn(unsigned long long out[8], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_wknn_stats, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ngicp::g_wknn_stats, z, sizeof z); }
  return 0;
}
#endif
