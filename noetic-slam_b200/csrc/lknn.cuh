// Leaf-scheduled exact self k-NN: the production search of K2 (covariance neighbourhoods).
// Replaces the per-query KD-tree descent of the reference
// (src/dlio/include/nano_gicp/nanoflann.h:1587-1666, called from nano_gicp.cc:343).
//
// Work item = up to 32 Morton-consecutive points of ONE leaf of the adaptive octree (a leaf = the largest cell
// with <= cmax points, found by knn.cu:leaf_items_kernel), one point per lane, so every lane of the warp shares one
// candidate set:
//   1. the 4x4x4 block of half-size cells around the leaf (its 8 children plus one ring) is located with 64 hash
//      probes (two per lane), the non-empty voxel buckets are compacted and staged into shared memory by TMA bulk
//      copies (every bucket is a contiguous run of float4 of the Morton-sorted array);
//   2. SELECT: every lane scans the same staged candidates (broadcast shared-memory reads) and keeps the k smallest
//      squared DISTANCES only — fp32, the reference's metric bit for bit. Candidates that beat the lane's current
//      k-th distance go to a small per-lane pending buffer in shared memory; when any lane's buffer fills up, all
//      lanes sort their (<= 8) pending values with a 19-comparator network and merge them into the sorted register
//      list with one bitonic merge. No index travels with the distances, so the list is K registers and a merge is
//      ~110 FMNMX for up to 8 candidates per lane (one insertion per candidate cost ~100 instructions before);
//   3. the block contains the 3x3x3 neighbourhood of every member's own half-size cell, so a member is done when
//      its k-th distance is closer than the nearest block face that still has grid behind it (same bound as
//      common.cuh:grid_knn); the others go one level up together;
//   4. COLLECT: the finished lanes scan the (still staged) candidates once more and write down the sorted positions
//      of everything closer than their k-th distance tau, plus as many candidates AT tau as the list holds — ties
//      at tau beyond that are resolved by smallest original index, the documented tie-break, on a slow path.
// The neighbour SET of a query is therefore exactly the k smallest (distance, original index) pairs.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ngicp {

constexpr int kLeafPend = 8;       // pending distances per lane between two merges
constexpr int kLeafPrune = 512;    // passes that stage more candidates than this first bound every member's k-th distance
                                   // (on the item's own 32-point run) and drop the voxel buckets out of everybody's reach

// one work item of the search (16 bytes)
struct __align__(16) LeafItem {
  int start;        // sorted position of the first member
  int count_level;  // (members << 8) | level of the leaf cell
  int pre;          // sorted position of a 32-point run of the same leaf that contains the members (bound pre-scan)
  int pre_count;    // its length (<= 32)
};

template <int KP, int C>
struct __align__(16) LeafScratch {
  float4 pts[2][C];              // staged candidates, double buffered (written by the TMA bulk copies)
  float pend[kLeafPend][32];     // pending distances, one column per lane
  int row[KP][32];               // collected neighbour positions, one column per lane
  uint32_t rstart[64];           // non-empty voxel buckets of the block, compacted, in scan order
  uint32_t rpre[68];             // exclusive prefix of their sizes; rpre[R] = M
  unsigned char rcode[64];       // cell id (0..63) of every compacted bucket, for box-distance pruning
  unsigned long long mbar[2];    // one mbarrier per buffer
};

// ---- sorted list of the K smallest distances of one lane (registers) ----------------------------------------------
__device__ __forceinline__ void ce(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

// 19-comparator sorting network for 8 values
__device__ __forceinline__ void sort8(float (&b)[8]) {
  ce(b[0], b[1]); ce(b[2], b[3]); ce(b[4], b[5]); ce(b[6], b[7]);
  ce(b[0], b[2]); ce(b[1], b[3]); ce(b[4], b[6]); ce(b[5], b[7]);
  ce(b[1], b[2]); ce(b[5], b[6]); ce(b[0], b[4]); ce(b[3], b[7]);
  ce(b[1], b[5]); ce(b[2], b[6]);
  ce(b[1], b[4]); ce(b[3], b[6]);
  ce(b[2], b[4]); ce(b[3], b[5]);
  ce(b[3], b[4]);
}

// dl (ascending, KP entries) <- the KP smallest of dl and the 8 pending values b (any order, +inf = empty)
template <int KP>
__device__ __forceinline__ void merge_pending(float (&dl)[KP], float (&b)[8]) {
  sort8(b);
  // b ascending against the top end of dl descending: element-wise min leaves the KP smallest as a bitonic sequence
  if (KP >= 8) {
#pragma unroll
    for (int i = 0; i < 8; i++) dl[KP - 1 - i] = fminf(dl[KP - 1 - i], b[i]);
  } else {
#pragma unroll
    for (int i = 0; i < KP; i++) dl[KP - 1 - i] = fminf(dl[KP - 1 - i], b[i]);
  }
#pragma unroll
  for (int s = KP / 2; s >= 1; s >>= 1) {
#pragma unroll
    for (int i = 0; i < KP; i++)
      if ((i & s) == 0) ce(dl[i], dl[i + s]);
  }
}

template <int KP>
__device__ __forceinline__ float kth_of(const float (&dl)[KP], int k) {   // k in 1..KP, static register indices only
  float v = dl[KP - 1];
#pragma unroll
  for (int i = 0; i < KP; i++) if (i == k - 1) v = dl[i];
  return v;
}

// ---- staging ------------------------------------------------------------------------------------------------------
template <class WS, int C>
__device__ __forceinline__ void leaf_issue(const GridView& g, WS& ws, int buf, int lane, int R, uint32_t c0, int nch) {
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
  unsigned int spins = 0;
  while (!mbar_try_wait(&ws.mbar[buf], (phase >> buf) & 1u)) {
    if (++spins > (1u << 24)) __trap();   // never hang the GPU on a programming error
  }
  phase ^= 1u << buf;
  __syncwarp();
}

// Run body(P, c0, nch) over all M staged candidates, chunk by chunk (chunk c+1 streams in while chunk c is scanned).
// `resident`: both buffers already hold chunks 0 and 1 of THIS bucket list and M <= 2C — nothing is copied again.
template <int C, class WS, class Body>
__device__ __forceinline__ void leaf_scan(const GridView& g, WS& ws, int lane, int R, uint32_t M, uint32_t& phase, bool& resident, Body body) {
  if (M == 0) return;
  const bool fits = M <= 2u * C;
  const bool copy = !(fits && resident);
  if (copy) leaf_issue<WS, C>(g, ws, 0, lane, R, 0u, (int)min((uint32_t)C, M));
  int buf = 0;
  for (uint32_t c0 = 0; c0 < M; c0 += C, buf ^= 1) {
    const int nch = (int)min((uint32_t)C, M - c0);
    if (copy) {
      if (c0 + C < M) leaf_issue<WS, C>(g, ws, buf ^ 1, lane, R, c0 + C, (int)min((uint32_t)C, M - c0 - C));
      leaf_wait(ws, buf, phase);
    }
    body(ws.pts[buf], c0, nch);
    __syncwarp();
  }
  resident = fits;
}

// cell offset (-1..2 per axis) of block cell id ci (0..63): ids 0..7 are the 8 children of the leaf, the ring follows
__device__ __forceinline__ void block_cell_offset(int ci, int& ox, int& oy, int& oz) {
  ox = (ci & 8) ? ((ci & 1) ? 2 : -1) : (ci & 1);
  oy = (ci & 16) ? ((ci & 2) ? 2 : -1) : ((ci >> 1) & 1);
  oz = (ci & 32) ? ((ci & 4) ? 2 : -1) : ((ci >> 2) & 1);
}

// 64 hash probes (two per lane) for the 4x4x4 block of level-Lg cells around parent cell (lpx,lpy,lpz); the non-empty
// buckets, compacted in scan order, with the exclusive prefix of their sizes. Returns R (buckets) and M (candidates).
template <class WS>
__device__ __forceinline__ void leaf_probe_block(const GridView& g, WS& ws, int lane, int sg, int Lg, int lpx, int lpy, int lpz, int& R, uint32_t& M) {
  const unsigned FULL = 0xffffffffu;
  const int maxc = kMaxCoord >> Lg;
  uint32_t cnt[2], st[2];
  {
    unsigned long long ck[2];
    uint32_t hh[2];
    uint4 raw[2];
    bool inside[2];
#pragma unroll
    for (int half = 0; half < 2; half++) {
      int ox, oy, oz;
      block_cell_offset(lane + 32 * half, ox, oy, oz);
      const int ax = 2 * lpx + ox, ay = 2 * lpy + oy, az = 2 * lpz + oz;
      inside[half] = ax >= 0 && ax <= maxc && ay >= 0 && ay <= maxc && az >= 0 && az <= maxc;
      ck[half] = pack_cell((unsigned)sg, Lg, (unsigned)ax, (unsigned)ay, (unsigned)az);
      hh[half] = hash64(ck[half]) & g.table_mask;
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
    }
  }
  uint32_t inc0 = cnt[0], inc1 = cnt[1];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t a = __shfl_up_sync(FULL, inc0, off), b = __shfl_up_sync(FULL, inc1, off);
    if (lane >= off) { inc0 += a; inc1 += b; }
  }
  const uint32_t tot0 = __shfl_sync(FULL, inc0, 31);
  M = tot0 + __shfl_sync(FULL, inc1, 31);
  const unsigned nz0 = __ballot_sync(FULL, cnt[0] != 0), nz1 = __ballot_sync(FULL, cnt[1] != 0);
  const unsigned lt = (1u << lane) - 1u;
  R = __popc(nz0) + __popc(nz1);
  __syncwarp();
  if (cnt[0]) { const int i0 = __popc(nz0 & lt); ws.rstart[i0] = st[0]; ws.rpre[i0] = inc0 - cnt[0]; ws.rcode[i0] = (unsigned char)lane; }
  if (cnt[1]) { const int i1 = __popc(nz0) + __popc(nz1 & lt); ws.rstart[i1] = st[1]; ws.rpre[i1] = tot0 + inc1 - cnt[1]; ws.rcode[i1] = (unsigned char)(lane + 32); }
  if (lane == 0) ws.rpre[R] = M;
  __syncwarp();
}

// Drop the buckets that lie farther from the members' bounding box (qlo..qhi, grid-relative coordinates) than the largest
// member bound; two buckets per lane, ballot compaction in place. A dropped bucket cannot hold any member's k nearest.
template <class WS>
__device__ __forceinline__ void leaf_prune_buckets(WS& ws, int lane, int Lg, int lpx, int lpy, int lpz, float h0, float margin,
                                                   const float qlo[3], const float qhi[3], float bound_max, int& R, uint32_t& M) {
  const unsigned FULL = 0xffffffffu;
  const float hL = h0 * (float)(1 << Lg), slack = 2.0f * margin;
  uint32_t bs[2], bc[2];
  int bcode[2];
  bool need[2];
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int bi = lane + 32 * half;
    const bool valid = bi < R;
    bs[half] = valid ? ws.rstart[bi] : 0u;
    bc[half] = valid ? ws.rpre[bi + 1] - ws.rpre[bi] : 0u;
    bcode[half] = valid ? (int)ws.rcode[bi] : 0;
    int ox, oy, oz;
    block_cell_offset(bcode[half], ox, oy, oz);
    const float ax = (float)(2 * lpx + ox) * hL, ay = (float)(2 * lpy + oy) * hL, az = (float)(2 * lpz + oz) * hL;
    const float ex = fmaxf(fmaxf(ax - slack - qhi[0], qlo[0] - (ax + hL + slack)), 0.0f);
    const float ey = fmaxf(fmaxf(ay - slack - qhi[1], qlo[1] - (ay + hL + slack)), 0.0f);
    const float ez = fmaxf(fmaxf(az - slack - qhi[2], qlo[2] - (az + hL + slack)), 0.0f);
    need[half] = valid && !((ex * ex + ey * ey + ez * ez) * 0.999999f > bound_max);
  }
  __syncwarp();
  const uint32_t k0 = need[0] ? bc[0] : 0u, k1 = need[1] ? bc[1] : 0u;
  uint32_t i0 = k0, i1 = k1;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t a = __shfl_up_sync(FULL, i0, off), b = __shfl_up_sync(FULL, i1, off);
    if (lane >= off) { i0 += a; i1 += b; }
  }
  const uint32_t t0 = __shfl_sync(FULL, i0, 31);
  const unsigned z0 = __ballot_sync(FULL, k0 != 0), z1 = __ballot_sync(FULL, k1 != 0);
  const unsigned ltm = (1u << lane) - 1u;
  if (k0) { const int w = __popc(z0 & ltm); ws.rstart[w] = bs[0]; ws.rpre[w] = i0 - k0; ws.rcode[w] = (unsigned char)bcode[0]; }
  __syncwarp();
  if (k1) { const int w = __popc(z0) + __popc(z1 & ltm); ws.rstart[w] = bs[1]; ws.rpre[w] = t0 + i1 - k1; ws.rcode[w] = (unsigned char)bcode[1]; }
  R = __popc(z0) + __popc(z1);
  M = t0 + __shfl_sync(FULL, i1, 31);
  if (lane == 0) ws.rpre[R] = M;
  __syncwarp();
}

// ---- one work item ------------------------------------------------------------------------------------------------
// KP = list size (8, 16 or 32), KC = compile-time k or 0 for a runtime k <= KP, C = staging chunk.
// Output per member j = item.start + lane: neighbour row (self first, the other k-1 in ascending sorted position)
// handed to `emit(j, a[], dens_sum)`, a[] = KP-or-16 ints, the first k-1 valid.
template <int KP, int KC, int C, class WS, class Emit>
__device__ __forceinline__ void leaf_knn_item(const GridView& g, const LeafItem item, int k_rt, WS& ws, uint32_t& phase, Emit emit) {
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
  const int sg = find_segment(g.seg_start, g.n_seg, item.start);
  const float4 o = __ldg(g.seg_origin + sg);
  const float ux = __fsub_rn(q.x, o.x), uy = __fsub_rn(q.y, o.y), uz = __fsub_rn(q.z, o.z);
  const int c0x = voxel_coord_unclamped(q.x, o.x, inv_h0), c0y = voxel_coord_unclamped(q.y, o.y, inv_h0), c0z = voxel_coord_unclamped(q.z, o.z, inv_h0);

  float dl[KP];
  bool done = !active;
  float bound = INF;            // a proven upper bound (exclusive) on the lane's k-th distance
  int Lg = (item.count_level & 0xff) - 1;
  for (bool first = true;; first = false, Lg++) {
    const bool member = !done;
    const unsigned mem_mask = __ballot_sync(FULL, member);
    if (!mem_mask) break;
    const int leader = __ffs(mem_mask) - 1;
    const int maxc = kMaxCoord >> Lg;
    const int px = clampi(c0x >> Lg, 0, maxc) >> 1, py = clampi(c0y >> Lg, 0, maxc) >> 1, pz = clampi(c0z >> Lg, 0, maxc) >> 1;
    const int lpx = __shfl_sync(FULL, px, leader), lpy = __shfl_sync(FULL, py, leader), lpz = __shfl_sync(FULL, pz, leader);
    int R;
    uint32_t M;
    leaf_probe_block(g, ws, lane, sg, Lg, lpx, lpy, lpz, R, M);
    bool resident = false;

    int cnt = 0;
    float thr = member ? bound : -1.0f;
#pragma unroll
    for (int i = 0; i < KP; i++) dl[i] = INF;
    auto flush = [&]() {
      float b[8];
#pragma unroll
      for (int i = 0; i < 8; i++) { b[i] = ws.pend[i][lane]; ws.pend[i][lane] = INF; }
      merge_pending<KP>(dl, b);
      cnt = 0;
      if (member) thr = fminf(bound, KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k));
    };

    if (M > (uint32_t)kLeafPrune) {
      if (first) {
        // bound every member's k-th distance on the item's own run of Morton-consecutive points
        const int pc = item.pre_count;
        const float4 pp = __ldg(g.pts + item.pre + min(lane, pc - 1));
        for (int t = 0; t < pc; t += 2) {
          const float ax = __shfl_sync(FULL, pp.x, t), ay = __shfl_sync(FULL, pp.y, t), az = __shfl_sync(FULL, pp.z, t);
          const int t1 = min(t + 1, pc - 1);
          const float bx = __shfl_sync(FULL, pp.x, t1), by = __shfl_sync(FULL, pp.y, t1), bz = __shfl_sync(FULL, pp.z, t1);
          const float d0 = sqdist_ref(q.x, q.y, q.z, ax, ay, az), d1 = sqdist_ref(q.x, q.y, q.z, bx, by, bz);
          if (d0 < thr) { ws.pend[cnt][lane] = d0; cnt++; }
          if (t + 1 < pc && d1 < thr) { ws.pend[cnt][lane] = d1; cnt++; }
          if (__any_sync(FULL, cnt >= kLeafPend - 1)) flush();
        }
        if (__any_sync(FULL, cnt > 0)) flush();
        const float kd = KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k);
        // candidates AT the bound must pass the strict test of the scan below
        if (member && kd < INF) bound = fminf(bound, __uint_as_float(__float_as_uint(kd) + 1u));
      }
      float qlo[3] = {member ? ux : INF, member ? uy : INF, member ? uz : INF};
      float qhi[3] = {member ? ux : -INF, member ? uy : -INF, member ? uz : -INF};
      float bmax = member ? bound : 0.0f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          qlo[a] = fminf(qlo[a], __shfl_xor_sync(FULL, qlo[a], off));
          qhi[a] = fmaxf(qhi[a], __shfl_xor_sync(FULL, qhi[a], off));
        }
        bmax = fmaxf(bmax, __shfl_xor_sync(FULL, bmax, off));
      }
      if (bmax < INF) leaf_prune_buckets(ws, lane, Lg, lpx, lpy, lpz, h0, margin, qlo, qhi, bmax, R, M);
      cnt = 0;
      thr = member ? bound : -1.0f;
#pragma unroll
      for (int i = 0; i < KP; i++) dl[i] = INF;
    }

    // ---- SELECT: the k smallest distances of every member
    leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
      for (int e = 0; e < nch; e += 2) {
        const float4 p0 = P[e], p1 = P[min(e + 1, nch - 1)];
        const float d0 = sqdist_ref(q.x, q.y, q.z, p0.x, p0.y, p0.z), d1 = sqdist_ref(q.x, q.y, q.z, p1.x, p1.y, p1.z);
        if (d0 < thr) { ws.pend[cnt][lane] = d0; cnt++; }
        if (e + 1 < nch && d1 < thr) { ws.pend[cnt][lane] = d1; cnt++; }
        if (__any_sync(FULL, cnt >= kLeafPend - 1)) flush();
      }
    });
    if (__any_sync(FULL, cnt > 0)) flush();

    // ---- termination test of the members (faces of the 4x4x4 block that still have grid behind them)
    const float tau = KC > 0 ? dl[KC - 1] : kth_of<KP>(dl, k);
    bool finish = false;
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
      finish = Lg >= kTopLevel || tau < cov2;
      if (!finish && tau < INF) bound = fminf(bound, __uint_as_float(__float_as_uint(tau) + 1u));
    }
    if (!__any_sync(FULL, finish)) continue;

    // ---- COLLECT: positions of the k-1 nearest other points of every finished member
    // dl[0] is the member itself (distance 0; an exact duplicate is the same point for every consumer): the row holds the
    // other k-1. need_eq = how many of them sit exactly AT tau.
    int n_lt = 0;
#pragma unroll
    for (int i = 1; i < KP; i++) n_lt += (i < k && dl[i] < tau) ? 1 : 0;
    const int need_eq = (k - 1) - n_lt;
    int cnt2 = 0, ties = 0;
    {
      int bcur = -1;
      uint32_t bend = 0;
      int base = 0;
      leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
        for (int e = 0; e < nch; e++) {
          const uint32_t c = c0 + (uint32_t)e;
          while (c >= bend) { bcur++; bend = ws.rpre[bcur + 1]; base = (int)ws.rstart[bcur] - (int)ws.rpre[bcur]; }
          const int pos = base + (int)c;
          const float4 p = P[e];
          const float d = sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z);
          if (finish && pos != j) {
            const bool eq = d == tau;
            const bool take = d < tau || (eq && ties < need_eq);
            ties += eq ? 1 : 0;
            if (take && cnt2 < KP) { ws.row[cnt2][lane] = pos; cnt2++; }
          }
        }
      });
    }
    // more candidates AT tau than the list holds: keep those with the smallest ORIGINAL indices (documented tie-break)
    bool redo = finish && ties > need_eq;
    if (__any_sync(FULL, redo)) {
      int cutoff = -1;   // largest original index taken among the ties
      for (int r = 0;; r++) {
        const bool want = redo && r < need_eq;
        if (!__any_sync(FULL, want)) break;
        int best = 0x7fffffff;
        int bcur = -1;
        uint32_t bend = 0;
        int base = 0;
        leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
          for (int e = 0; e < nch; e++) {
            const uint32_t c = c0 + (uint32_t)e;
            while (c >= bend) { bcur++; bend = ws.rpre[bcur + 1]; base = (int)ws.rstart[bcur] - (int)ws.rpre[bcur]; }
            const float4 p = P[e];
            const float d = sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z);
            const int oi = __float_as_int(p.w);
            if (want && base + (int)c != j && d == tau && oi > cutoff && oi < best) best = oi;
          }
        });
        if (want) cutoff = best;
      }
      if (redo) cnt2 = 0;
      int bcur = -1;
      uint32_t bend = 0;
      int base = 0;
      leaf_scan<C>(g, ws, lane, R, M, phase, resident, [&](const float4* __restrict__ P, uint32_t c0, int nch) {
        for (int e = 0; e < nch; e++) {
          const uint32_t c = c0 + (uint32_t)e;
          while (c >= bend) { bcur++; bend = ws.rpre[bcur + 1]; base = (int)ws.rstart[bcur] - (int)ws.rpre[bcur]; }
          const int pos = base + (int)c;
          const float4 p = P[e];
          const float d = sqdist_ref(q.x, q.y, q.z, p.x, p.y, p.z);
          if (redo && pos != j && (d < tau || (d == tau && __float_as_int(p.w) <= cutoff)) && cnt2 < KP) { ws.row[cnt2][lane] = pos; cnt2++; }
        }
      });
    }
    if (finish) {
      constexpr int N = KP < 16 ? 16 : KP;
      int a[N];
#pragma unroll
      for (int i = 0; i < N; i++) a[i] = (i < KP && i < k - 1) ? (i < cnt2 ? ws.row[i][lane] : j) : 0x7fffffff;
      double dsum = 0.0;
#pragma unroll
      for (int i = 1; i < KP; i++) if (i < k) dsum += (double)dl[i];
      emit(j, a, dsum);
      done = true;
    }
  }
}

}  // namespace ngicp
