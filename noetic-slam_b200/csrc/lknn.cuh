// Leaf-scheduled exact self k-NN: the production search of K2 (covariance neighbourhoods).
// Replaces the per-query KD-tree descent of the reference
// (src/dlio/include/nano_gicp/nanoflann.h:1587-1666, called from nano_gicp.cc:343).
//
// Work item = up to 32 Morton-consecutive points of ONE leaf of the adaptive octree (a leaf = the largest cell
// with <= cmax points, found by knn.cu:leaf_items_kernel), one point per lane, so every lane of the warp shares one
// candidate set:
//   1. the 4x4x4 block of half-size cells around the leaf (its 8 children plus one ring) is located with 64 hash
//      probes (two per lane); the non-empty voxel buckets are sorted by their distance from the members' bounding box
//      (a 64-key bitonic sort across the warp) and staged into shared memory chunk by chunk with TMA bulk copies (every
//      bucket is a contiguous run of float4 of the Morton-sorted array), nearest buckets first;
//   2. SELECT: every lane scans the same staged candidates (broadcast shared-memory reads) and keeps its k smallest
//      KEYS in a sorted register list. A key is the fp32 squared distance — the reference's metric bit for bit — with
//      its low mantissa bits replaced by the candidate's slot in the staged list, so 32 bits name both. Candidates that
//      beat the lane's current k-th key go to a small per-lane pending buffer in shared memory; when any lane's buffer
//      fills up, all lanes sort their (<= 8) pending keys with a 19-comparator network and merge them into the list
//      with one bitonic merge (~110 integer min/max for up to 8 candidates per lane). Before another chunk is
//      requested the list of buckets is cut where the reach of the members' current k-th distances ends;
//   3. the block contains the 3x3x3 neighbourhood of every member's own half-size cell, so a member is done when
//      its k-th distance is closer than the nearest block face that still has grid behind it (same bound as
//      common.cuh:grid_knn); the others go one level up together, taking their k-th distance along as a bound;
//   4. a finished lane's list is CERTIFIED when the (k+1)-th smallest key differs from the k-th in its distance bits
//      (the smallest rejected / dropped key is tracked): then the slots of the list ARE the k nearest and are decoded
//      to sorted positions (binary search over the bucket prefix). The few lanes that are not certified (~0.3 %) select
//      once more with exact distance bits and COLLECT in a second scan: everything closer than the k-th distance tau,
//      plus as many candidates AT tau as the list holds — ties at tau beyond that are resolved by smallest original
//      index, the documented tie-break.
// The neighbour SET of a query is therefore exactly the k smallest (distance, original index) pairs.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ngicp {

#ifdef NGICP_STATS
// development counters: [0] items [1] passes [2] candidates staged (after pruning) [3] member lanes [4] lanes finished fast
// [5] lanes finished exact [6] passes with an exact re-selection [7] refused lanes [8] prune calls [9] candidates before
// pruning (pruned passes only) [10] candidates after [11] flushes [12] max M [13] passes that stream (M > 2C) [14] tie redos
static __device__ unsigned long long g_leaf_stats[16];
#define LEAF_STAT(i, v) do { const unsigned long long _sv = (unsigned long long)(v); if ((threadIdx.x & 31) == 0) atomicAdd(&g_leaf_stats[i], _sv); } while (0)
#define LEAF_STAT_MAX(i, v) do { if ((threadIdx.x & 31) == 0) atomicMax(&g_leaf_stats[i], (unsigned long long)(v)); } while (0)
// development: up to 64 records of lanes that finish with fewer than k valid list entries or a repeated key
static __device__ unsigned long long g_leaf_items[4 * 8192];   // per work item: start / end (globaltimer ns), (passes << 32 | members), candidates scanned
static __device__ unsigned int g_leaf_cur[8192][8];   // per resident warp: passes, candidates scanned, largest staged M, exact re-selections of the current item
static __device__ unsigned int g_leaf_phase[4 * 4096];   // per work item: cycles in probe, select (incl. flushes), flushes; number of flushes
static __device__ unsigned int g_leaf_dbg_n;
#define LEAF_CUR(i, v) do { if ((threadIdx.x & 31) == 0) { unsigned int* _c = g_leaf_cur[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & 8191]; if ((i) == 2) _c[2] = max(_c[2], (unsigned)(v)); else _c[i] += (unsigned)(v); } } while (0)
static __device__ unsigned int g_leaf_dbg[64][24];
#else
#define LEAF_STAT(i, v) do { } while (0)
#define LEAF_STAT_MAX(i, v) do { } while (0)
#define LEAF_CUR(i, v) do { } while (0)
#endif
#ifdef NGICP_STATS
#define LEAF_CLK(var) const long long var = clock64()
#define LEAF_CLK_ADD(i, t0) LEAF_CUR(i, (unsigned)(clock64() - (t0)))
#else
#define LEAF_CLK(var) do { } while (0)
#define LEAF_CLK_ADD(i, t0) do { } while (0)
#endif

#ifndef NGICP_LEAF_HEAVY
#define NGICP_LEAF_HEAVY 768
#endif
constexpr int kLeafHeavy = NGICP_LEAF_HEAVY;    // first passes that would stage more candidates than this are split by child cell and re-queued
constexpr int kLeafPend = 8;       // pending distances per lane between two merges

// one work item of the search (8 bytes)
struct __align__(8) LeafItem {
  int start;        // sorted position of the first member
  int count_level;  // (members << 8) | level of the leaf cell
};

template <int KP, int C>
struct __align__(16) LeafScratch {
  float4 pts[2][C];              // staged candidates, double buffered (written by the TMA bulk copies)
  unsigned pend[kLeafPend][32];  // pending keys, one column per lane
  int row[KP][32];               // collected neighbour positions, one column per lane
  uint32_t rstart[64];           // non-empty voxel buckets of the block, sorted by their distance from the members
  uint32_t rpre[68];             // exclusive prefix of their sizes; rpre[R] = M, 0xffffffff behind
  float rdist[64];               // lower bound of the squared distance between bucket and the members' bounding box (+inf behind R)
  unsigned long long mbar[2];    // one mbarrier per buffer
  int use_tma;                   // staging by TMA bulk copies (1) or by direct loads (0)
  int pad_[3];
};

// ---- sorted list of the K smallest keys of one lane (registers) -----------------------------------------------------
// A key is the bit pattern of a non-negative fp32 squared distance (unsigned integer order == float order), in the fast
// mode with its low `sb` mantissa bits replaced by the candidate's slot in the staged list. 0xffffffff = empty.
__device__ __forceinline__ void ce(unsigned& a, unsigned& b) { const unsigned lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

// 19-comparator sorting network for 8 values
__device__ __forceinline__ void sort8(unsigned (&b)[8]) {
  ce(b[0], b[1]); ce(b[2], b[3]); ce(b[4], b[5]); ce(b[6], b[7]);
  ce(b[0], b[2]); ce(b[1], b[3]); ce(b[4], b[6]); ce(b[5], b[7]);
  ce(b[1], b[2]); ce(b[5], b[6]); ce(b[0], b[4]); ce(b[3], b[7]);
  ce(b[1], b[5]); ce(b[2], b[6]);
  ce(b[1], b[4]); ce(b[3], b[6]);
  ce(b[2], b[4]); ce(b[3], b[5]);
  ce(b[3], b[4]);
}

// dl (ascending, KP entries) <- the KP smallest of dl and the 8 pending keys b (any order, 0xffffffff = empty).
// Returns the smallest key that dropped out (0xffffffff if none did).
template <int KP>
__device__ __forceinline__ unsigned merge_pending(unsigned (&dl)[KP], unsigned (&b)[8]) {
  sort8(b);
  // b ascending against the top end of dl descending: element-wise min leaves the KP smallest as a bitonic sequence
  unsigned out = 0xffffffffu;
  constexpr int NB = KP >= 8 ? 8 : KP;
#pragma unroll
  for (int i = 0; i < NB; i++) {
    out = min(out, max(dl[KP - 1 - i], b[i]));
    dl[KP - 1 - i] = min(dl[KP - 1 - i], b[i]);
  }
#pragma unroll
  for (int i = NB; i < 8; i++) out = min(out, b[i]);
#pragma unroll
  for (int s = KP / 2; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < KP; i++)
      if ((i & s) == 0) ce(dl[i], dl[i + s]);
  }
  return out;
}

// bitonic sorting network over N registers (N a power of two), ascending
template <int N>
__device__ __forceinline__ void sort_regs(unsigned (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned lo = min(a[i], a[l]), hi = max(a[i], a[l]);
          if ((i & k) == 0) { a[i] = lo; a[l] = hi; } else { a[i] = hi; a[l] = lo; }
        }
      }
}

template <int KP>
__device__ __forceinline__ unsigned kth_of(const unsigned (&dl)[KP], int k) {   // k in 1..KP (k > KP: empty), static register indices only
  unsigned v = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < KP; i++) if (i == k - 1) v = dl[i];
  return v;
}

// ---- staging ------------------------------------------------------------------------------------------------------
template <class WS, int C>
__device__ __forceinline__ void leaf_issue(const GridView& g, WS& ws, int buf, int lane, int R, uint32_t c0, int nch) {
  if (!ws.use_tma) {
    // direct staging: every lane fetches candidates c0 + lane + 32 t (coalesced inside a bucket) after locating their bucket
    // in the prefix array. Synchronous, but without the fixed cost of a TMA operation per (small) bucket.
#pragma unroll
    for (int t = 0; t < C / 32; t++) {
      const uint32_t e = (uint32_t)(lane + 32 * t);
      if ((int)e < nch) {
        const uint32_t c = c0 + e;
        int b = 0;
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1) b += (ws.rpre[b + step] <= c) ? step : 0;
        ws.pts[buf][e] = __ldg(g.pts + (ws.rstart[b] + (c - ws.rpre[b])));
      }
    }
    __syncwarp();
    return;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this buffer before async writes
  if (lane == 0) mbar_expect_tx(&ws.mbar[buf], (uint32_t)nch * 16u);
  const uint32_t c1 = c0 + (uint32_t)nch;
  for (int ri = lane; ri < R; ri += 32) {
    const uint32_t pre = ws.rpre[ri], nxt = ws.rpre[ri + 1];
    const uint32_t lo = max(pre, c0), hi = min(nxt, c1);
    if (lo < hi) tma_bulk_g2s(&ws.pts[buf][lo - c0], g.pts + (ws.rstart[ri] + (lo - pre)), (hi - lo) * 16u, &ws.mbar[buf]);
  }
}
template <class WS>
__device__ __forceinline__ void leaf_wait(WS& ws, int buf, uint32_t& phase) {
  if (!ws.use_tma) return;
  unsigned int spins = 0;
  while (!mbar_try_wait(&ws.mbar[buf], (phase >> buf) & 1u)) {
    if (++spins > (1u << 24)) __trap();   // never hang the GPU on a programming error
  }
  phase ^= 1u << buf;
  __syncwarp();
}

// Run body(P, c0, nch) over the staged candidates [0, M), chunk by chunk (chunk c+1 streams in while chunk c is scanned).
// `resident`: both buffers already hold chunks 0 and 1 of THIS bucket list and M <= 2C — nothing is copied again.
// `trim(limit)` is asked before every further chunk is requested and may shrink M (the buckets are sorted by their distance
// from the members, so everything behind the returned limit is out of every member's reach); never below `limit`, the end
// of what is already on its way.
template <int C, class WS, class Body, class Trim>
__device__ __forceinline__ void leaf_scan(const GridView& g, WS& ws, int lane, int R, uint32_t& M, uint32_t& phase, bool& resident, Body body, Trim trim) {
  if (M == 0) return;
  const bool copy = !(M <= 2u * C && resident);
  if (copy) leaf_issue<WS, C>(g, ws, 0, lane, R, 0u, (int)min((uint32_t)C, M));
  int buf = 0;
  for (uint32_t c0 = 0; c0 < M; c0 += C, buf ^= 1) {
    const int nch = (int)min((uint32_t)C, M - c0);
    if (c0 + C < M) {
      if (c0 > 0) M = max(min(M, trim()), c0 + (uint32_t)C);   // chunk c0 is already staged (or on its way): finish it in any case
      if (copy && c0 + C < M) leaf_issue<WS, C>(g, ws, buf ^ 1, lane, R, c0 + C, (int)min((uint32_t)C, M - c0 - C));
    }
    if (copy) { LEAF_CLK(t_wait); leaf_wait(ws, buf, phase); LEAF_CLK_ADD(7, t_wait); }
    body(ws.pts[buf], c0, nch);
    __syncwarp();
  }
  resident = M <= 2u * C;
}
template <int C, class WS, class Body>
__device__ __forceinline__ void leaf_scan(const GridView& g, WS& ws, int lane, int R, uint32_t& M, uint32_t& phase, bool& resident, Body body) {
  leaf_scan<C>(g, ws, lane, R, M, phase, resident, body, [&]() { return M; });
}

// cell offset (-1..2 per axis) of block cell id ci (0..63): ids 0..7 are the 8 children of the leaf, the ring follows
__device__ __forceinline__ void block_cell_offset(int ci, int& ox, int& oy, int& oz) {
  ox = (ci & 8) ? ((ci & 1) ? 2 : -1) : (ci & 1);
  oy = (ci & 16) ? ((ci & 2) ? 2 : -1) : ((ci >> 1) & 1);
  oz = (ci & 32) ? ((ci & 4) ? 2 : -1) : ((ci >> 2) & 1);
}

// 64 hash probes (two per lane) for the 4x4x4 block of level-Lg cells around parent cell (lpx,lpy,lpz). The non-empty
// buckets are SORTED by a lower bound of their squared distance from the members' bounding box (qlo..qhi, grid-relative
// coordinates; the cells the box touches come first, the children of the leaf in front) and listed with the exclusive
// prefix of their sizes. Returns R (buckets) and M (candidates).
template <class WS>
__device__ __forceinline__ void leaf_probe_block(const GridView& g, WS& ws, int lane, int sg, int Lg, int lpx, int lpy, int lpz, float h0, float margin,
                                                 const float qlo[3], const float qhi[3], bool want_sort, uint32_t sort_above, int& R, uint32_t& M) {
  const unsigned FULL = 0xffffffffu;
  const int maxc = kMaxCoord >> Lg;
  const float hL = h0 * (float)(1 << Lg), slack = 2.0f * margin;
  uint32_t cnt[2], st[2], key[2];
  {
    unsigned long long ck[2];
    uint32_t hh[2];
    uint4 raw[2];
    bool inside[2];
    float d2[2];
#pragma unroll
    for (int half = 0; half < 2; half++) {
      int ox, oy, oz;
      block_cell_offset(lane + 32 * half, ox, oy, oz);
      const int ax = 2 * lpx + ox, ay = 2 * lpy + oy, az = 2 * lpz + oz;
      inside[half] = ax >= 0 && ax <= maxc && ay >= 0 && ay <= maxc && az >= 0 && az <= maxc;
      ck[half] = pack_cell((unsigned)sg, Lg, (unsigned)ax, (unsigned)ay, (unsigned)az);
      hh[half] = hash64(ck[half]) & g.table_mask;
      const float bx = (float)ax * hL, by = (float)ay * hL, bz = (float)az * hL;
      const float ex = fmaxf(fmaxf(bx - slack - qhi[0], qlo[0] - (bx + hL + slack)), 0.0f);
      const float ey = fmaxf(fmaxf(by - slack - qhi[1], qlo[1] - (by + hL + slack)), 0.0f);
      const float ez = fmaxf(fmaxf(bz - slack - qhi[2], qlo[2] - (bz + hL + slack)), 0.0f);
      d2[half] = (ex * ex + ey * ey + ez * ez) * 0.999999f;
    }
#pragma unroll
    for (int half = 0; half < 2; half++) raw[half] = inside[half] ? __ldg(reinterpret_cast<const uint4*>(g.table + hh[half])) : make_uint4(~0u, ~0u, 0u, 0u);
#pragma unroll
    for (int half = 0; half < 2; half++) {
      uint32_t s = 0, e = 0;
      unsigned long long k = ((unsigned long long)raw[half].y << 32) | raw[half].x;
      uint32_t hcur = hh[half];
      while (k != kEmptyKey) {
        if (k == ck[half]) { s = raw[half].z; e = raw[half].w; break; }
        hcur = (hcur + 1) & g.table_mask;
        raw[half] = __ldg(reinterpret_cast<const uint4*>(g.table + hcur));
        k = ((unsigned long long)raw[half].y << 32) | raw[half].x;
      }
      st[half] = s;
      cnt[half] = e - s;
      // sort key: distance bits rounded down to a multiple of 64 (still a lower bound) with the cell id below; empty cells last
      key[half] = e > s ? ((__float_as_uint(d2[half]) & ~63u) | (uint32_t)(lane + 32 * half)) : 0xffffffffu;
    }
  }
  // unsorted total first: a pass whose candidates all fit the two staging buffers is never cut short, so the order of its
  // buckets does not matter and the sort is skipped (cell-id order, the children of the leaf in front)
  uint32_t inc0 = cnt[0], inc1 = cnt[1];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t a = __shfl_up_sync(FULL, inc0, off), b = __shfl_up_sync(FULL, inc1, off);
    if (lane >= off) { inc0 += a; inc1 += b; }
  }
  uint32_t tot0 = __shfl_sync(FULL, inc0, 31);
  M = tot0 + __shfl_sync(FULL, inc1, 31);
  const unsigned nz0 = __ballot_sync(FULL, cnt[0] != 0), nz1 = __ballot_sync(FULL, cnt[1] != 0);
  R = __popc(nz0) + __popc(nz1);
  __syncwarp();
  if (!want_sort && M <= sort_above) {
    const unsigned lt = (1u << lane) - 1u;
    if (cnt[0]) { const int i0 = __popc(nz0 & lt); ws.rstart[i0] = st[0]; ws.rpre[i0] = inc0 - cnt[0]; }
    if (cnt[1]) { const int i1 = __popc(nz0) + __popc(nz1 & lt); ws.rstart[i1] = st[1]; ws.rpre[i1] = tot0 + inc1 - cnt[1]; }
    for (int i = R + lane; i < 68; i += 32) ws.rpre[i] = i == R ? M : 0xffffffffu;
    ws.rdist[lane] = 0.0f; ws.rdist[lane + 32] = 0.0f;
    __syncwarp();
    return;
  }
  // bitonic sort of the 64 keys: element e = lane + 32 * half
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        const uint32_t lo = min(key[0], key[1]), hi = max(key[0], key[1]);
        key[0] = lo; key[1] = hi;
      } else {
#pragma unroll
        for (int half = 0; half < 2; half++) {
          const int e = lane + 32 * half;
          const uint32_t other = __shfl_xor_sync(FULL, key[half], j);
          const bool take_min = ((e & k) == 0) == ((e & j) == 0);
          key[half] = take_min ? min(key[half], other) : max(key[half], other);
        }
      }
    }
  }
  // payload of the cell that landed at sorted position p = lane + 32 * half
  uint32_t sst[2], scnt[2];
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const bool have = key[half] != 0xffffffffu;
    const int id = (int)(key[half] & 63u);
    const uint32_t s0 = __shfl_sync(FULL, st[0], id & 31), s1 = __shfl_sync(FULL, st[1], id & 31);
    const uint32_t c0 = __shfl_sync(FULL, cnt[0], id & 31), c1 = __shfl_sync(FULL, cnt[1], id & 31);
    sst[half] = have ? ((id >> 5) ? s1 : s0) : 0u;
    scnt[half] = have ? ((id >> 5) ? c1 : c0) : 0u;
  }
  inc0 = scnt[0]; inc1 = scnt[1];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t a = __shfl_up_sync(FULL, inc0, off), b = __shfl_up_sync(FULL, inc1, off);
    if (lane >= off) { inc0 += a; inc1 += b; }
  }
  tot0 = __shfl_sync(FULL, inc0, 31);
  // the non-empty buckets occupy positions 0..R-1; everything behind rpre[R] = M compares high (the slot -> bucket search of
  // the fast mode and the reach test read rpre / rdist without a range check)
  ws.rstart[lane] = sst[0]; ws.rstart[lane + 32] = sst[1];
  ws.rpre[lane] = scnt[0] ? inc0 - scnt[0] : (lane == R ? M : 0xffffffffu);
  ws.rpre[lane + 32] = scnt[1] ? tot0 + inc1 - scnt[1] : (lane + 32 == R ? M : 0xffffffffu);
  ws.rdist[lane] = scnt[0] ? __uint_as_float(key[0] & ~63u) : __int_as_float(0x7f800000);
  ws.rdist[lane + 32] = scnt[1] ? __uint_as_float(key[1] & ~63u) : __int_as_float(0x7f800000);
  if (lane < 4) ws.rpre[64 + lane] = (64 + lane == R) ? M : 0xffffffffu;
  __syncwarp();
}

// number of candidates in the buckets within squared distance `reach` of the members' box (a prefix of the sorted list)
template <class WS>
__device__ __forceinline__ uint32_t leaf_reach(const WS& ws, int lane, float reach) {
  const unsigned FULL = 0xffffffffu;
  const int cut = __popc(__ballot_sync(FULL, ws.rdist[lane] <= reach)) + __popc(__ballot_sync(FULL, ws.rdist[lane + 32] <= reach));
  return ws.rpre[cut];
}

// ---- one work item ------------------------------------------------------------------------------------------------
// KP = list size (8, 16 or 32), KC = compile-time k or 0 for a runtime k <= KP, C = staging chunk.
// Output per member j = item.start + lane: `emit(j, a[], dens_sum)`, a[] = the sorted positions of the other k-1
// neighbours in its first k-1 entries after an ascending sort (the rest is INT_MAX).
constexpr unsigned kKeyEmpty = 0xffffffffu;
constexpr unsigned kKeyInf = 0x7f800000u;     // keys from here on (inf, NaN) are never accepted

__device__ __forceinline__ float key_hi_float(unsigned key, unsigned smask) { return __uint_as_float(key | smask); }

// `split(heads, level)`: re-queue the members as separate work items, one per run of lanes that starts at a set bit of
// `heads`, at leaf level `level`; false if the queue has no room (the item is then processed as it is).
template <int KP, int KC, int C, class WS, class Emit, class Split>
__device__ __forceinline__ void leaf_knn_item(const GridView& g, const LeafItem item, int k_rt, WS& ws, uint32_t& phase, Emit emit, Split split) {
  const unsigned FULL = 0xffffffffu;
  const float INF = __int_as_float(0x7f800000);
  const int lane = threadIdx.x & 31;
  const int k = KC > 0 ? KC : k_rt;
  const int count = item.count_level >> 8;
  const bool active = lane < count;
  const int j = item.start + (active ? lane : 0);
  const float4 q = __ldg(g.pts + j);
  const GridMeta* __restrict__ m = g.meta;
  const float h0 = __ldg(&m->h0), inv_h0 = __ldg(&m->inv_h0), margin = __ldg(&m->margin);
  const int fine = __ldg(&m->fine_level);
  const int sg = find_segment(g.seg_start, g.n_seg, item.start);
  const float4 o = __ldg(g.seg_origin + sg);
  const float ux = __fsub_rn(q.x, o.x), uy = __fsub_rn(q.y, o.y), uz = __fsub_rn(q.z, o.z);
  const int c0x = voxel_coord_unclamped(q.x, o.x, inv_h0), c0y = voxel_coord_unclamped(q.y, o.y, inv_h0), c0z = voxel_coord_unclamped(q.z, o.z, inv_h0);

  unsigned dl[KP];
  bool done = !active;
  unsigned bound = kKeyInf;     // every candidate whose distance bits are below this may still belong to the lane's k nearest
  int Lg = (item.count_level & 0xff) - 1, Lnext = Lg + 1;
  for (bool first = true;; first = false, Lg = Lnext) {
    const bool member = !done;
    const unsigned mem_mask = __ballot_sync(FULL, member);
    if (!mem_mask) break;
    const int leader = __ffs(mem_mask) - 1;
    const int maxc = kMaxCoord >> Lg;
    const int px = clampi(c0x >> Lg, 0, maxc) >> 1, py = clampi(c0y >> Lg, 0, maxc) >> 1, pz = clampi(c0z >> Lg, 0, maxc) >> 1;
    const int lpx = __shfl_sync(FULL, px, leader), lpy = __shfl_sync(FULL, py, leader), lpz = __shfl_sync(FULL, pz, leader);
    // bounding box of the members (grid-relative): the candidate buckets are visited in the order of their distance from it
    float qlo[3] = {member ? ux : INF, member ? uy : INF, member ? uz : INF};
    float qhi[3] = {member ? ux : -INF, member ? uy : -INF, member ? uz : -INF};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        qlo[a] = fminf(qlo[a], __shfl_xor_sync(FULL, qlo[a], off));
        qhi[a] = fmaxf(qhi[a], __shfl_xor_sync(FULL, qhi[a], off));
      }
    }
    int R;
    uint32_t M;
    LEAF_CLK(t_probe);
    // (a pass with a bound carried over from the level below starts with a cut of the sorted list: it always sorts)
    leaf_probe_block(g, ws, lane, sg, Lg, lpx, lpy, lpz, h0, margin, qlo, qhi, !first, 2u * C, R, M);
    LEAF_CLK_ADD(4, t_probe);
    bool resident = false;
    LEAF_STAT(1, 1); LEAF_STAT(3, __popc(mem_mask)); if (first) LEAF_STAT(0, 1);
    LEAF_STAT(9, M);

    // ---- list state of one selection; `sel` = the lanes that select, the others accept nothing
    // pending keys go to column `lane` of ws.pend through a running shared-memory address (one register; the compiler
    // otherwise rebuilds the column address from special registers at every store)
    const uint32_t pend0 = smem_u32(&ws.pend[0][lane]);
    uint32_t pw = pend0;
    unsigned thr = 0u, next = kKeyEmpty, beff = 0u, cur_sm = 0u;
    bool sel = false;
    // sm = slot mask of the keys about to be offered: `bound` lives in exact distance bits, and a fast key may exceed the
    // distance bits it was made from by up to sm, so the bound is widened to the end of its truncation class
    auto reset = [&](bool lanes, unsigned sm) {
      sel = lanes;
      cur_sm = sm;
      pw = pend0;
      next = kKeyEmpty;
      beff = ((bound - 1u) | sm) + 1u;
      thr = sel ? beff : 0u;
#pragma unroll
      for (int i = 0; i < KP; i++) dl[i] = kKeyEmpty;
    };
    auto flush = [&]() {
      LEAF_CLK(t_flush);
      unsigned b[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { b[i] = ws.pend[i][lane]; ws.pend[i][lane] = kKeyEmpty; }
      next = min(next, merge_pending<KP>(dl, b));
      pw = pend0;
      LEAF_STAT(11, 1);
      LEAF_CLK_ADD(6, t_flush);
      if (sel) thr = min(beff, KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k));
    };
    auto offer = [&](float d, unsigned slot, unsigned dmask) {
      const unsigned key = (__float_as_uint(d) & dmask) | slot;
      const bool take = key < thr;
      if (take) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(pw), "r"(key) : "memory"); pw += 128u; }
      next = min(next, take ? kKeyEmpty : key);
    };
    // how far the selecting lanes still reach: the largest squared distance any of them could still accept (thr is an
    // exclusive key bound; a key below it was made from distance bits of at most thr | slot mask)
    auto reach = [&]() {
      float r = sel ? (thr >= kKeyInf ? INF : __uint_as_float(thr | cur_sm)) : 0.0f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, off));
      return r;
    };

    // a bound carried over from the level below: the buckets out of every member's reach are not even staged
    {
      float r = member ? (bound >= kKeyInf ? INF : __uint_as_float(bound)) : 0.0f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, off));
      if (r < INF) M = min(M, leaf_reach(ws, lane, r));
    }
    // A block with thousands of candidates around a leaf of a few dozen points (a sparse cell inside a dense neighbourhood):
    // hand the members back to the queue as one work item per child cell, one level down — smaller blocks around tighter
    // groups, spread over the warps that have run out of work. Members the smaller block does not satisfy climb back
    // here with a bound, and then only the buckets within their reach are staged.
    if (first && M > (uint32_t)kLeafHeavy && Lg > fine) {
      const int gx = clampi(c0x >> Lg, 0, maxc), gy = clampi(c0y >> Lg, 0, maxc), gz = clampi(c0z >> Lg, 0, maxc);
      const int qx_ = __shfl_up_sync(FULL, gx, 1), qy_ = __shfl_up_sync(FULL, gy, 1), qz_ = __shfl_up_sync(FULL, gz, 1);
      const bool head = active && (lane == 0 || gx != qx_ || gy != qy_ || gz != qz_);
      if (split(__ballot_sync(FULL, head), Lg)) { LEAF_STAT(8, 1); return; }
    }
    LEAF_STAT(2, M); LEAF_STAT_MAX(12, M); if (M > 2u * C) LEAF_STAT(13, 1);
    LEAF_CUR(0, 1); LEAF_CUR(2, M);

    // ---- SELECT, fast mode: the low `sb` mantissa bits of a key carry the candidate's slot in the staged list, so the
    //      k smallest keys name the neighbours themselves — valid as long as the (k+1)-th key (`next`: the smallest key that
    //      was rejected or dropped) does not have the same truncated distance as the k-th. sb = 0: exact keys, slots unknown.
    const int sb = M <= 16384u ? max(32 - __clz((int)max(M, 2u) - 1), 5) : 0;      // as few slot bits as the M slots need
    const unsigned smask = (1u << sb) - 1u;
    const int self_orig = __float_as_int(q.w);
    constexpr int N = KP < 16 ? 16 : KP;

    // sorted position of a slot of the staged list: largest b with rpre[b] <= slot (rpre is padded behind rpre[R])
    auto slot_pos = [&](unsigned slot) {
      int b = 0;
#pragma unroll
      for (int step = 32; step >= 1; step >>= 1) b += (ws.rpre[b + step] <= slot) ? step : 0;
      return (int)(ws.rstart[b] + (slot - ws.rpre[b]));
    };

    // ONE selection loop for both modes (fast first; exact keys for the lanes that end up needing them)
    unsigned sm = smask;
    bool lanes = member, finish = false, exact_lane = false;
    unsigned tau_key = kKeyEmpty;
    for (int round = 0;; round++) {
      LEAF_CLK(t_select);
      reset(lanes, sm);
      {
        const unsigned dmask = ~sm;
        leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
          LEAF_CUR(1, nch);
          int e = 0;
          for (; e + 4 <= nch; e += 4) {
            // four candidates at a time, loads first: independent chains for a warp that is alone on its scheduler
            const float4 p0 = P[e], p1 = P[e + 1], p2 = P[e + 2], p3 = P[e + 3];
            const float d0 = sqdist_ref(q.x, q.y, q.z, p0.x, p0.y, p0.z), d1 = sqdist_ref(q.x, q.y, q.z, p1.x, p1.y, p1.z);
            const float d2 = sqdist_ref(q.x, q.y, q.z, p2.x, p2.y, p2.z), d3 = sqdist_ref(q.x, q.y, q.z, p3.x, p3.y, p3.z);
            const unsigned s0 = c0 + (uint32_t)e;
            offer(d0, s0 & sm, dmask); offer(d1, (s0 + 1u) & sm, dmask); offer(d2, (s0 + 2u) & sm, dmask); offer(d3, (s0 + 3u) & sm, dmask);
            if (__any_sync(FULL, pw >= pend0 + 128u * (kLeafPend - 3))) flush();
          }
          for (; e < nch; e++) {
            const float4 p = P[e];
            offer(sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z), (c0 + (uint32_t)e) & sm, dmask);
          }
          if (__any_sync(FULL, pw >= pend0 + 128u * (kLeafPend - 3))) flush();
        }, [&]() {
          // before another chunk is requested: everything pending counts, then cut the list where the members' reach ends
          if (__any_sync(FULL, pw != pend0)) flush();
          const float r = reach();
          return r < INF ? leaf_reach(ws, lane, r) : 0xffffffffu;
        });
        if (__any_sync(FULL, pw != pend0)) flush();
      }
      LEAF_CLK_ADD(5, t_select);
      if (round == 1) break;

      // ---- termination test of the members (faces of the 4x4x4 block that still have grid behind them). In the fast mode
      //      the k-th distance is only known up to its truncation: test its upper end (refusing is always safe).
      tau_key = KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k);
      if (member) {
        const float hL = h0 * (float)(1 << Lg);
        float gap = INF;
        {
          const int lo_c = 2 * lpx - 1, hi_c = 2 * lpx + 2;
          if (lo_c > 0) gap = fminf(gap, fmaxf(ux - (float)lo_c * hL, 0.0f));
          if (hi_c < maxc) gap = fminf(gap, fmaxf((float)(hi_c + 1) * hL - ux, 0.0f));
        }
        {
          const int lo_c = 2 * lpy - 1, hi_c = 2 * lpy + 2;
          if (lo_c > 0) gap = fminf(gap, fmaxf(uy - (float)lo_c * hL, 0.0f));
          if (hi_c < maxc) gap = fminf(gap, fmaxf((float)(hi_c + 1) * hL - uy, 0.0f));
        }
        {
          const int lo_c = 2 * lpz - 1, hi_c = 2 * lpz + 2;
          if (lo_c > 0) gap = fminf(gap, fmaxf(uz - (float)lo_c * hL, 0.0f));
          if (hi_c < maxc) gap = fminf(gap, fmaxf((float)(hi_c + 1) * hL - uz, 0.0f));
        }
        const float covered = fmaxf(gap - margin, 0.0f);
        const float cov2 = covered * covered * 0.999999f;
        const bool have = tau_key < kKeyInf;
        finish = Lg >= kTopLevel || (have && key_hi_float(tau_key, smask) < cov2);
        if (!finish && have) bound = min(bound, (tau_key | smask) + 1u);
      }
      // next level: one up; further if a member's bound asks for it (the ring of a block is one cell wide, so a level whose
      // cells are wider than the bound's radius certainly settles the member)
      {
        int need = Lg + 1;
        if (member && !finish && bound < kKeyInf) {
          const float rb = sqrtf(__uint_as_float(bound)) * 1.000001f;
          while (need < kTopLevel && h0 * (float)(1 << need) * 0.999f - margin <= rb) need++;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) need = max(need, __shfl_xor_sync(FULL, need, off));
        Lnext = min(need, kTopLevel);
      }
      LEAF_STAT(7, __popc(__ballot_sync(FULL, member && !finish)));
      if (!__any_sync(FULL, finish)) break;
      exact_lane = finish;
      if (sb == 0) break;                              // the selection above already ran on exact keys

      // ---- fast mode: certify the lists
      const unsigned after = min(next, kth_of<KP>(dl, k + 1));        // the (k+1)-th smallest key overall
      const bool certain = tau_key >= kKeyInf || after == kKeyEmpty || ((after ^ tau_key) & ~smask) != 0u;
      // Not certified: the list and the rest share the truncated distance of the k-th key. Everything below that class is
      // settled; one more scan picks up the exact low bits of every candidate IN the class (a handful) and the list keeps the
      // closest of them. Only an exact tie at the cut (original indices decide) or an overfull class goes to the exact mode.
      bool resolved = false;
      const bool unc = finish && !certain;
      if (__any_sync(FULL, unc)) {
        const unsigned cls = tau_key & ~smask;
        int a_in = 0;
#pragma unroll
        for (int i = 0; i < KP; i++) a_in += (i < k && (dl[i] & ~smask) == cls) ? 1 : 0;
        int cc = 0;
        leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
          for (int e = 0; e < nch; e++) {
            const float4 p = P[e];
            const unsigned bits = __float_as_uint(sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z));
            const bool in = unc && (bits & ~smask) == cls;
            if (in && cc < KP) ws.row[cc][lane] = (int)(((bits & smask) << 16) | (c0 + (uint32_t)e));
            cc += in ? 1 : 0;
          }
        });
        bool ok = unc && cc <= KP && cc >= a_in;
        unsigned ev[KP];
#pragma unroll
        for (int i = 0; i < KP; i++) ev[i] = (ok && i < cc) ? (unsigned)ws.row[i][lane] : 0xffffffffu;
        sort_regs<KP>(ev);
        if (ok && a_in < cc && (kth_of<KP>(ev, a_in) >> 16) == (kth_of<KP>(ev, a_in + 1) >> 16)) ok = false;   // exact tie at the cut
        if (ok) {
#pragma unroll
          for (int i = 0; i < KP; i++) ws.row[i][lane] = (int)ev[i];
#pragma unroll
          for (int i = 0; i < KP; i++)
            if (i < k && i >= k - a_in) dl[i] = cls | ((unsigned)ws.row[i - (k - a_in)][lane] & 0xffffu);
        }
        resolved = ok;
        LEAF_STAT(6, 1); LEAF_CUR(3, 1);
      }
      const bool fast = finish && (certain || resolved);
      exact_lane = finish && !fast;

      // ---- decode the slots of the certified lists
      if (__any_sync(FULL, fast)) {
        int a[N];
        double dsum = 0.0;
        bool self_found = false;
#pragma unroll
        for (int i = 0; i < N; i++) a[i] = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < KP; i++) {
          if (i < k) {
            const unsigned key = dl[i];
            const bool valid = fast && key < kKeyInf;
            const unsigned slot = valid ? (key & smask) : 0u;
            const int pos = slot_pos(slot);
            float4 pt = q;
            if (valid) pt = resident ? ws.pts[slot / C][slot % C] : __ldg(g.pts + pos);
            dsum += valid ? (double)sqdist_ref(q.x, q.y, q.z, pt.x, pt.y, pt.z) : 0.0;
            const bool is_self = valid && pos == j;
            self_found |= is_self;
            a[i] = (valid && !is_self) ? pos : 0x7fffffff;
          }
        }
        // >= k exact duplicates of the query with smaller slots: drop one of them instead of the query itself
        if (!self_found) {
#pragma unroll
          for (int i = 0; i < KP; i++) if (i == k - 1) a[i] = 0x7fffffff;
        }
        if (fast) {
          emit(j, a, dsum);
          done = true;
        }
      }
      LEAF_STAT(4, __popc(__ballot_sync(FULL, fast)));
      if (!__any_sync(FULL, exact_lane)) break;
      LEAF_STAT(15, 1);
      // the others select once more, with exact keys
      lanes = exact_lane;
      sm = 0u;
    }
    if (!__any_sync(FULL, exact_lane)) continue;

    // ---- exact mode (rare). COLLECT the slots of the k-1 nearest other points of every remaining finished member:
    // dl[0] is the member itself (distance 0; an exact duplicate is the same point for every consumer), so the row holds
    // everything below the k-th distance tau plus need_eq candidates AT tau. If more candidates sit at tau than that, those
    // with the smallest ORIGINAL indices are kept (documented tie-break): a cutoff index is found by need_eq further scans.
    const unsigned tau = KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k);
    int n_lt = 0;
#pragma unroll
    for (int i = 1; i < KP; i++) n_lt += (i < k && dl[i] < tau) ? 1 : 0;
    const int need_eq = (k - 1) - n_lt;
    int cnt2 = 0, ties = 0, cutoff = 0x7fffffff, best = 0x7fffffff;
    bool redo = false;
    // passes of one scan loop: 0 = collect, 1 .. need_eq = find the next tied original index, last = collect again below the cutoff
    for (int pass = 0;; pass++) {
      const bool finding = pass >= 1 && redo && pass <= need_eq;
      const bool collecting = pass == 0 ? exact_lane : (redo && pass > need_eq);
      if (pass >= 1 && !__any_sync(FULL, redo && pass <= need_eq + 1)) break;
      if (collecting) cnt2 = 0;
      best = 0x7fffffff;
      leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
        for (int e = 0; e < nch; e++) {
          const float4 p = P[e];
          const unsigned d = __float_as_uint(sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z));
          const int oi = __float_as_int(p.w);
          const bool other = oi != self_orig;
          if (collecting && other) {
            const bool eq = d == tau;
            const bool take = d < tau || (eq && (pass == 0 ? ties < need_eq : oi <= cutoff));
            if (pass == 0) ties += eq ? 1 : 0;
            if (take && cnt2 < KP) { ws.row[cnt2][lane] = (int)(c0 + (uint32_t)e); cnt2++; }
          }
          if (finding && other && d == tau && (pass == 1 || oi > cutoff) && oi < best) best = oi;
        }
      });
      if (pass == 0) {
        redo = exact_lane && ties > need_eq;
        LEAF_STAT(5, __popc(__ballot_sync(FULL, exact_lane)));
        if (__any_sync(FULL, redo)) LEAF_STAT(14, 1);
        cutoff = -1;
      }
      if (finding) cutoff = best;
    }
    if (exact_lane) {
      int a[N];
#pragma unroll
      for (int i = 0; i < N; i++) a[i] = (i < KP && i < k - 1) ? (i < cnt2 ? slot_pos((unsigned)ws.row[i][lane]) : j) : 0x7fffffff;
      double dsum = 0.0;
#pragma unroll
      for (int i = 1; i < KP; i++) if (i < k && dl[i] < kKeyInf) dsum += (double)__uint_as_float(dl[i]);
      emit(j, a, dsum);
      done = true;
    }
  }
}

}  // namespace ngicp
