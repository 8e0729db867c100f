// Warp-cooperative exact k-NN over the voxel hash: the production search of K2 (covariance k-NN) and
// K4 (correspondences). Replaces the per-query KD-tree descent of the reference
// (src/dlio/include/nano_gicp/nanoflann.h:1587-1666, called from nano_gicp.cc:224,343).
//
// A warp owns 32 consecutive queries (Morton order => spatially coherent). Repeatedly:
//   1. the first unresolved lane leads. Twelve lanes probe the leader's ancestor cells at all
//      levels at once; the group cell is the LARGEST ancestor holding <= cmax points (never finer
//      than the level a lane was already refused at, never coarser than the search radius needs);
//   2. every unresolved lane inside that group cell joins;
//   3. the 4x4x4 block of half-size cells around the group cell (its 8 children plus one ring) is
//      staged into shared memory: 64 hash probes, two per lane, then coalesced copies of the
//      (contiguous, Morton-sorted) voxel buckets, children first;
//   4. all members scan the SAME staged candidates from shared memory (broadcast reads, no per-lane
//      global gathers, no trip-count divergence), top-k in registers;
//   5. the block contains the 3x3x3 neighbourhood of every member's own half-size cell, so the
//      exactness proof of common.cuh:grid_knn carries over: a member is done when its k-th best is
//      closer than the nearest block face that still has grid behind it; otherwise it asks for a
//      coarser group next time.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace ngicp {

constexpr int kWarpChunk = 256;  // candidates staged per pass and warp
constexpr int kPruneStaged = 1024;  // passes that stage more candidates than this first drop the voxel buckets no member can need

#ifdef NGICP_STATS
// development counters: [0] passes, [1] staged candidates, [2] member lanes, [3] warp_knn calls, [4] refused members
static __device__ unsigned long long g_wknn_stats[8];
#define WKNN_STAT(i, v) do { const unsigned long long _sv = (unsigned long long)(v); if ((threadIdx.x & 31) == 0) atomicAdd(&g_wknn_stats[i], _sv); } while (0)
static __device__ unsigned long long g_bnn_stats[16];
#define BNN_STAT(i, v) atomicAdd(&g_bnn_stats[i], (unsigned long long)(v))
#else
#define WKNN_STAT(i, v) do { } while (0)
#define BNN_STAT(i, v) do { } while (0)
#endif

struct __align__(16) WarpScratch {
  float4 pts[2][kWarpChunk];    // staged candidates, double buffered (written by the TMA bulk copies)
  uint32_t rstart[128];         // non-empty voxel buckets of the block (<= 64) or ball (<= 125), compacted, in scan order
  uint32_t rpre[129];           // exclusive prefix of their sizes; rpre[R] = M
  uint32_t pad[3];
  unsigned char rcode[128];     // block mode: cell id (0..63, scan order) of every compacted bucket, for box-distance pruning
  unsigned long long mbar[2];   // one mbarrier per buffer: the bulk copies complete on it
  float4 qm[8];                 // member queries of the current pass (k = 1 path)
};


// Stage candidates [c0, c0+nch) of the concatenated bucket list into shared-memory buffer `buf`. Every voxel bucket
// is a contiguous run of float4 in the Morton-sorted array, so each overlapping bucket is ONE TMA bulk copy
// (cp.async.bulk, issued by the lane that owns the bucket) completing on the buffer's mbarrier. The original index of
// every candidate rides in the .w lane of the copied float4. Issue and wait are separate so that the next chunk
// streams in while the warp scans the current one.
__device__ __forceinline__ void wknn_issue(const GridView& g, WarpScratch& ws, int buf, int lane, int R, uint32_t c0, int nch) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this buffer before async writes
  if (lane == 0) mbar_expect_tx(&ws.mbar[buf], (uint32_t)nch * 16u);
  const uint32_t c1 = c0 + (uint32_t)nch;
  for (int ri = lane; ri < R; ri += 32) {
    const uint32_t pre = ws.rpre[ri], nxt = ws.rpre[ri + 1];
    const uint32_t lo = max(pre, c0), hi = min(nxt, c1);
    if (lo < hi) tma_bulk_g2s(&ws.pts[buf][lo - c0], g.pts + (ws.rstart[ri] + (lo - pre)), (hi - lo) * 16u, &ws.mbar[buf]);
  }
}
__device__ __forceinline__ void wknn_wait(WarpScratch& ws, int buf, uint32_t& phase) {
  unsigned int spins = 0;
  while (!mbar_try_wait(&ws.mbar[buf], (phase >> buf) & 1u)) {
    if (++spins > (1u << 24)) __trap();   // never hang the GPU on a programming error
  }
  phase ^= 1u << buf;
  __syncwarp();
}
// chunk loop used by every scan mode: prefetch chunk c+1, wait for chunk c, run BODY(P, nch) on it
#define WKNN_FOR_CHUNKS(...)                                                                               \
  {                                                                                                        \
    int buf_ = 0;                                                                                          \
    if (M > 0) wknn_issue(g, ws, 0, lane, R, 0u, (int)min((uint32_t)kWarpChunk, M));                       \
    for (uint32_t c0 = 0; c0 < M; c0 += kWarpChunk, buf_ ^= 1) {                                           \
      const int nch = (int)min((uint32_t)kWarpChunk, M - c0);                                              \
      if (c0 + kWarpChunk < M) wknn_issue(g, ws, buf_ ^ 1, lane, R, c0 + kWarpChunk, (int)min((uint32_t)kWarpChunk, M - c0 - kWarpChunk)); \
      wknn_wait(ws, buf_, phase);                                                                          \
      const float4* __restrict__ P = ws.pts[buf_];                                                         \
      __VA_ARGS__                                                                                          \
      __syncwarp();                                                                                        \
    }                                                                                                      \
  }

// Once per warp and kernel, before the first warp_knn: arm the warp's mbarrier. (Initialising an mbarrier twice is
// undefined, so persistent kernels that search many times keep one barrier and carry its phase parity along.)
__device__ __forceinline__ uint32_t wknn_init(WarpScratch& ws) {
  if ((threadIdx.x & 31) == 0) { mbar_init(&ws.mbar[0], 1); mbar_init(&ws.mbar[1], 1); }
#ifdef NGICP_STATS
  if ((threadIdx.x & 31) == 0) { ws.pad[0] = 0; ws.pad[1] = 0; ws.pad[2] = 0; }
#endif
  __syncwarp();
  return 0u;   // bit b = parity of the phase the next chunk staged into buffer b completes
}

// LPQ = lanes per query (1, 2, 4 or 8). The LPQ lanes of a query hold identical query state and split the
// staged candidates between them, which multiplies the number of warps a small cloud can keep in flight
// (a 65,536-point scan is only 2,048 warps at one query per lane) and shortens every warp's dependent chain.
template <int LPQ, class TK>
__device__ __forceinline__ void warp_knn(const GridView& g, bool active, float qx, float qy, float qz, int seg, int k, int cmax,
                                         float max_sqd, TK& best, WarpScratch& ws, uint32_t& phase) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const GridMeta* __restrict__ m = g.meta;
  const float h0 = __ldg(&m->h0), inv_h0 = __ldg(&m->inv_h0), margin = __ldg(&m->margin);
  const int base = __ldg(&m->base_level);
  if (!active) seg = 0;
  const float4 o = __ldg(g.seg_origin + seg);
  const float ux = __fsub_rn(qx, o.x), uy = __fsub_rn(qy, o.y), uz = __fsub_rn(qz, o.z);
  const int c0x = voxel_coord_unclamped(qx, o.x, inv_h0);
  const int c0y = voxel_coord_unclamped(qy, o.y, inv_h0);
  const int c0z = voxel_coord_unclamped(qz, o.z, inv_h0);

  // coarsest level the search radius can ever need: one block at L_reach covers max_sqd
  int L_reach = base;
  for (; L_reach < kTopLevel; L_reach++) {
    const float reach = h0 * (float)(1 << L_reach) - margin;
    if (reach * reach * 0.999999f >= max_sqd) break;
  }

  int Lmin = base;  // finest group level this lane still accepts
  bool done = !active;
  best.reset();
  WKNN_STAT(3, 1);

  for (;;) {
    const unsigned un = __ballot_sync(FULL, !done);
    if (!un) break;
    const int leader = __ffs(un) - 1;
    const int lcx = __shfl_sync(FULL, c0x, leader), lcy = __shfl_sync(FULL, c0y, leader), lcz = __shfl_sync(FULL, c0z, leader);
    const int sg = __shfl_sync(FULL, seg, leader);
    const int lLmin = __shfl_sync(FULL, Lmin, leader);

    // ---- 1. group level: largest ancestor of the leader (levels base+1 .. top) with <= cmax points
    int Lg;
    if (lLmin > base) {
      // the leader was refused one level below: go exactly one level up, no need to probe its ancestors again
      Lg = min(lLmin, kTopLevel);
    } else {
      const int P = base + 1 + lane;
      bool ok = false;
      if (P <= kTopLevel) {
        const int maxcP = kMaxCoord >> P;
        const unsigned int ax = clampi(lcx >> P, 0, maxcP), ay = clampi(lcy >> P, 0, maxcP), az = clampi(lcz >> P, 0, maxcP);
        const unsigned long long ck = pack_cell((unsigned)sg, P, ax, ay, az);
        uint32_t s = 0, e = 0;
        const bool hit = cell_lookup(g.table, g.table_mask, ck, s, e);
        ok = !hit || (int)(e - s) <= cmax;
      }
      const unsigned okm = __ballot_sync(FULL, ok);
      // counts grow with the level, so the ok lanes form a prefix; take its last lane
      const int nprefix = __ffs(~okm) - 1;          // number of leading ok lanes (0..12)
      Lg = base + (nprefix > 0 ? nprefix - 1 : 0);  // group cell level P = Lg + 1
      Lg = min(Lg, L_reach);
      Lg = min(Lg, kTopLevel);
    }

    const int maxc = kMaxCoord >> Lg;
    const int px = clampi(c0x >> Lg, 0, maxc) >> 1, py = clampi(c0y >> Lg, 0, maxc) >> 1, pz = clampi(c0z >> Lg, 0, maxc) >> 1;
    const int lpx = __shfl_sync(FULL, px, leader), lpy = __shfl_sync(FULL, py, leader), lpz = __shfl_sync(FULL, pz, leader);
    const bool member = !done && Lmin <= Lg && seg == sg && px == lpx && py == lpy && pz == lpz;

    // ---- 3. 64 hash probes, two per lane; ranges + exclusive prefix of their sizes into shared memory
    uint32_t cnt[2], st[2];
    {
      unsigned long long ck[2];
      uint32_t hh[2];
      uint4 raw[2];
      bool inside[2];
#pragma unroll
      for (int half = 0; half < 2; half++) {
        // scan order without a table (a lane-indexed __constant__ table serialised 32 ways): bits 5..3 of the cell id say
        // per axis whether the cell lies in the outer ring, bits 2..0 pick the side. Ids 0..7 are the 8 children of the
        // group cell (they hold the nearest candidates, so the top-k threshold tightens early), the ring follows.
        const int ci = lane + 32 * half;
        const int ox = (ci & 8) ? ((ci & 1) ? 2 : -1) : (ci & 1), oy = (ci & 16) ? ((ci & 2) ? 2 : -1) : ((ci >> 1) & 1),
                  oz = (ci & 32) ? ((ci & 4) ? 2 : -1) : ((ci >> 2) & 1);
        const int ax = 2 * lpx + ox, ay = 2 * lpy + oy, az = 2 * lpz + oz;
        inside[half] = ax >= 0 && ax <= maxc && ay >= 0 && ay <= maxc && az >= 0 && az <= maxc;
        ck[half] = pack_cell((unsigned)sg, Lg, (unsigned)ax, (unsigned)ay, (unsigned)az);
        hh[half] = hash64(ck[half]) & g.table_mask;
      }
      // both first probes are issued before either is consumed (one memory round trip instead of two)
#pragma unroll
      for (int half = 0; half < 2; half++) raw[half] = inside[half] ? __ldg(reinterpret_cast<const uint4*>(g.table + hh[half])) : make_uint4(~0u, ~0u, 0u, 0u);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t s = 0, e = 0;
        unsigned long long k = ((unsigned long long)raw[half].y << 32) | raw[half].x;
        uint32_t hcur = hh[half];
        while (k != kEmptyKey) {
          if (k == ck[half]) { s = raw[half].z; e = raw[half].w; break; }
          hcur = (hcur + 1) & g.table_mask;            // collision: continue the probe sequence
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
    const uint32_t M = tot0 + __shfl_sync(FULL, inc1, 31);
    // keep only the non-empty buckets, in scan order
    const unsigned nz0 = __ballot_sync(FULL, cnt[0] != 0), nz1 = __ballot_sync(FULL, cnt[1] != 0);
    const unsigned lt = (1u << lane) - 1u;
    const int R = __popc(nz0) + __popc(nz1);
    if (cnt[0]) { const int i0 = __popc(nz0 & lt); ws.rstart[i0] = st[0]; ws.rpre[i0] = inc0 - cnt[0]; ws.rcode[i0] = (unsigned char)lane; }
    if (cnt[1]) { const int i1 = __popc(nz0) + __popc(nz1 & lt); ws.rstart[i1] = st[1]; ws.rpre[i1] = tot0 + inc1 - cnt[1]; ws.rcode[i1] = (unsigned char)(lane + 32); }
    if (lane == 0) ws.rpre[R] = M;
    __syncwarp();
    if (member) best.reset();
    WKNN_STAT(0, 1); WKNN_STAT(1, M); WKNN_STAT(2, __popc(__ballot_sync(FULL, member)));
#ifdef NGICP_STATS
    if (lane == 0) { ws.pad[0] += 1; ws.pad[1] += M; ws.pad[2] = max(ws.pad[2], M); }   // per work item: passes, staged candidates, largest pass
#endif
#ifdef NGICP_STATS
    if (lane == 0) { atomicMax(&g_wknn_stats[5], (unsigned long long)M); if (M > 2048) { atomicAdd(&g_wknn_stats[6], 1ull); atomicAdd(&g_wknn_stats[7], (unsigned long long)M); } }
#endif

    // ---- 3b/4. stage buckets chunk by chunk and scan them. Three modes:
    //   shared      : every member scans every staged candidate (cost ~ M per lane, whatever the member count)
    //   split, k=1  : the 32 lanes split each chunk, one warp-argmin per member and chunk (cost ~ members*M/32)
    //   split, k>1  : per member, the lanes split all candidates into private top-k lists that are merged by
    //                 K rounds of warp-argmin; pays off for the heavy tail (few members, thousands of candidates)
    const unsigned mem_mask = __ballot_sync(FULL, member);
    const int nmem = __popc(mem_mask) / LPQ;
    const unsigned grp_mask = (LPQ == 32 ? 0xffffffffu : ((1u << LPQ) - 1u));
    if (TK::kK == 1 && LPQ >= 4) {
      // k = 1, at most 8 queries per warp: every lane keeps a running best for each member query over ITS share of the
      // candidates (8 per staged chunk) and the warp-argmin is taken once per member at the end of the pass, not per
      // chunk — the heavy tail (thousands of candidates) is no longer paced by shuffle reductions.
      constexpr int NS = 32 / LPQ;
      if (member && sub == 0) ws.qm[lane / LPQ] = make_float4(qx, qy, qz, 0.f);
      unsigned slots = 0;
#pragma unroll
      for (int sl = 0; sl < NS; sl++) slots |= ((mem_mask >> (sl * LPQ)) & 1u) << sl;
      __syncwarp();
      float bd[NS];
      int bp[NS];
#pragma unroll
      for (int sl = 0; sl < NS; sl++) { bd[sl] = __int_as_float(0x7f800000); bp[sl] = -1; }
      WKNN_FOR_CHUNKS({
        float4 cand[kWarpChunk / 32];
#pragma unroll
        for (int t = 0; t < kWarpChunk / 32; t++) {
          cand[t] = P[min(lane + 32 * t, nch - 1)];
          if (lane + 32 * t >= nch) cand[t].w = __int_as_float(-1);   // padding: loses every comparison below
        }
#pragma unroll
        for (int sl = 0; sl < NS; sl++) {
          if (!((slots >> sl) & 1u)) continue;
          const float4 mq = ws.qm[sl];
#pragma unroll
          for (int t = 0; t < kWarpChunk / 32; t++) {
            const float d = sqdist_ref(mq.x, mq.y, mq.z, cand[t].x, cand[t].y, cand[t].z);
            const int pi = __float_as_int(cand[t].w);
            if (pi >= 0 && TK::before(d, pi, bd[sl], bp[sl])) { bd[sl] = d; bp[sl] = pi; }
          }
        }
      })
#pragma unroll
      for (int sl = 0; sl < NS; sl++) {
        if (!((slots >> sl) & 1u)) continue;
        float rd = bd[sl];
        int rp = bp[sl];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float od = __shfl_xor_sync(FULL, rd, off);
          const int op = __shfl_xor_sync(FULL, rp, off);
          if (TK::before(od, op, rd, rp)) { rd = od; rp = op; }
        }
        if (lane / LPQ == sl && rp >= 0 && rd <= best.worst()) best.offer(rd, rp);
      }
    } else if (TK::kK == 1) {
      WKNN_FOR_CHUNKS({
        float4 cand[kWarpChunk / 32];
#pragma unroll
        for (int t = 0; t < kWarpChunk / 32; t++) cand[t] = P[min(lane + 32 * t, nch - 1)];
        unsigned rem = mem_mask;
        while (rem) {
          const int mi = __ffs(rem) - 1;                 // first lane of the next member query
          rem &= ~(grp_mask << (mi & ~(LPQ - 1)));
          const float mx = __shfl_sync(FULL, qx, mi), my = __shfl_sync(FULL, qy, mi), mz = __shfl_sync(FULL, qz, mi);
          float bd = __int_as_float(0x7f800000);
          int bp = -1;
#pragma unroll
          for (int t = 0; t < kWarpChunk / 32; t++) {
            const float d = sqdist_ref(mx, my, mz, cand[t].x, cand[t].y, cand[t].z);
            const int pi = __float_as_int(cand[t].w);
            if (lane + 32 * t < nch && TK::before(d, pi, bd, bp)) { bd = d; bp = pi; }
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(FULL, bd, off);
            const int op = __shfl_xor_sync(FULL, bp, off);
            if (TK::before(od, op, bd, bp)) { bd = od; bp = op; }
          }
          if ((lane & ~(LPQ - 1)) == (mi & ~(LPQ - 1)) && bp >= 0 && bd <= best.worst()) best.offer(bd, bp);
        }
      })
    } else {
      // k > 1. One scan loop and one merge for both ways of sharing the staged candidates:
      //   shared : every member query is scanned by its own LPQ lanes (width LPQ, one round)
      //   heavy  : few member queries, many candidates -> all 32 lanes split the candidates of one query at a time
      // Each lane keeps a private sorted list of packed (distance, index) keys over its share; the lists of the `width`
      // lanes are then merged by log2(width) bitonic steps (exchange with the xor partner, keep the K smallest of the
      // two sorted lists as a bitonic sequence, re-sort it with the half-cleaner network). Deliberately ONE code site
      // for the insertion and ONE for the merge: this kernel used to be 138 KB of SASS and stalled on instruction fetch.
      if (LPQ == 1) {
        WKNN_FOR_CHUNKS({
          if (member) {
#pragma unroll 4
            for (int e = 0; e < nch; e++) {
              const float4 p = P[e];
              const float d = sqdist_ref(qx, qy, qz, p.x, p.y, p.z);
              if (d <= best.worst()) best.offer(d, __float_as_int(p.w));
            }
          }
        })
      } else {
        // Big passes (a sparse group cell next to a dense one: thousands of staged candidates, nearly all of them far
        // from every member) run twice. Phase 0 scans only the first chunk — the group cell's own children come first
        // in scan order — which gives every member a valid bound on its k-th distance; the voxel buckets that no
        // member's bound reaches (box distance) are then dropped from the staging list and phase 1 is the normal scan,
        // from scratch, over what is left. Same result: a dropped bucket cannot hold any member's k nearest.
        uint32_t Mf = M;
        int Rf = R;
        for (int ph = (M > (uint32_t)kPruneStaged) ? 0 : 1; ph < 2; ph++) {
        const uint32_t M = ph == 0 ? (uint32_t)kWarpChunk : Mf;      // (shadow the pass totals: the chunk loop reads M and R)
        const int R = Rf;
        const bool heavy = ph == 0 ? false : (long long)nmem * (2ll * M + 1000) < (long long)(60 / LPQ) * M;
        const int width = heavy ? 32 : LPQ;
        const int s0 = lane & (width - 1);
        unsigned rem = heavy ? mem_mask : 1u;
        while (rem) {                                  // warp-uniform
          const int mi = __ffs(rem) - 1;               // heavy: first lane of the member query of this round
          rem = heavy ? (rem & ~(grp_mask << (mi & ~(LPQ - 1)))) : 0u;
          const float mx = heavy ? __shfl_sync(FULL, qx, mi) : qx;
          const float my = heavy ? __shfl_sync(FULL, qy, mi) : qy;
          const float mz = heavy ? __shfl_sync(FULL, qz, mi) : qz;
          const bool scan = heavy || member;
          KeyList<TK::kK> part;
          // The private lists together prove a bound on the final k-th distance long before any one of them is full:
          // when every lane of the query holds Q entries and width * Q >= K, at least K candidates are known at or
          // below the largest of the lanes' Q-th entries, so nothing farther can make the final list.
          constexpr int kQs = (TK::kK + LPQ - 1) / LPQ - 1;          // shared: LPQ lanes per query
          constexpr int kQh = (TK::kK + 31) / 32 - 1;                // heavy: 32 lanes per query
          float lim = __int_as_float(0x7f800000);
          WKNN_FOR_CHUNKS({
            int e_first = s0;
            if (c0 == 0) {
              // seed: the first KP candidates of this lane's share, sorted by a network instead of inserted one by one
#pragma unroll
              for (int i = 0; i < KeyList<TK::kK>::KP; i++) {
                const int e = s0 + i * width;
                const float4 p = P[min(e, nch - 1)];
                const float d = sqdist_ref(mx, my, mz, p.x, p.y, p.z);
                part.k[i] = (scan && e < nch) ? KeyList<TK::kK>::make_key(d, __float_as_int(p.w)) : KeyList<TK::kK>::kEmpty;
              }
              part.sort_all();
              e_first = s0 + KeyList<TK::kK>::KP * width;
            }
            {
              float b = heavy ? part.dist_at(kQh) : part.dist_at(kQs);
              for (int off = 1; off < width; off <<= 1) b = fmaxf(b, __shfl_xor_sync(FULL, b, off));
              lim = fminf(lim, b);
            }
            if (scan) {
#pragma unroll 2
              for (int e = e_first; e < nch; e += width) {
                const float4 p = P[e];
                const float d = sqdist_ref(mx, my, mz, p.x, p.y, p.z);
                if (d <= fminf(lim, part.worst())) part.offer(d, __float_as_int(p.w));
              }
            }
          })
          for (int off = 1; off < width; off <<= 1) part.merge_with_partner(off);
          const bool take = heavy ? ((lane & ~(LPQ - 1)) == (mi & ~(LPQ - 1))) : member;
          if (take) part.store(best);
        }
        if (ph == 0) {
          // ---- drop the buckets no member can need; two buckets per lane as at staging time
          uint32_t bs[2], bc[2];
          int bcode[2];
          bool need[2];
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const int bi = lane + 32 * half;
            const bool valid = bi < Rf;
            bs[half] = valid ? ws.rstart[bi] : 0u;
            bc[half] = valid ? ws.rpre[bi + 1] - ws.rpre[bi] : 0u;
            bcode[half] = valid ? (int)ws.rcode[bi] : 0;
            need[half] = false;
          }
          const float my_bound = best.worst();
          const float phL = h0 * (float)(1 << Lg), slack = 2.0f * margin;
          for (int sl = 0; sl < 32 / LPQ; sl++) {
            const int src = sl * LPQ;
            if (!((mem_mask >> src) & 1u)) continue;                                // uniform
            const float bux = __shfl_sync(FULL, ux, src), buy = __shfl_sync(FULL, uy, src), buz = __shfl_sync(FULL, uz, src);
            const float bb = __shfl_sync(FULL, my_bound, src);
#pragma unroll
            for (int half = 0; half < 2; half++) {
              const int ci = bcode[half];
              const int ox = (ci & 8) ? ((ci & 1) ? 2 : -1) : (ci & 1), oy = (ci & 16) ? ((ci & 2) ? 2 : -1) : ((ci >> 1) & 1),
                        oz = (ci & 32) ? ((ci & 4) ? 2 : -1) : ((ci >> 2) & 1);
              const float ax = (float)(2 * lpx + ox) * phL, ay = (float)(2 * lpy + oy) * phL, az = (float)(2 * lpz + oz) * phL;
              const float ex = fmaxf(fmaxf(ax - slack - bux, bux - (ax + phL + slack)), 0.0f);
              const float ey = fmaxf(fmaxf(ay - slack - buy, buy - (ay + phL + slack)), 0.0f);
              const float ez = fmaxf(fmaxf(az - slack - buz, buz - (az + phL + slack)), 0.0f);
              need[half] |= !((ex * ex + ey * ey + ez * ez) * 0.999999f > bb);
            }
          }
          __syncwarp();
          const uint32_t k0 = (need[0] && bc[0]) ? bc[0] : 0u, k1 = (need[1] && bc[1]) ? bc[1] : 0u;
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
          __syncwarp();   // the first half's compacted slots all lie below the second half's sources read above: order the writes anyway
          if (k1) { const int w = __popc(z0) + __popc(z1 & ltm); ws.rstart[w] = bs[1]; ws.rpre[w] = t0 + i1 - k1; ws.rcode[w] = (unsigned char)bcode[1]; }
          Rf = __popc(z0) + __popc(z1);
          Mf = t0 + __shfl_sync(FULL, i1, 31);
          if (lane == 0) ws.rpre[Rf] = Mf;
          __syncwarp();
        }
        }
      }
    }

    // ---- 5. termination test of the members (same bound as grid_knn, faces of the 4x4x4 block)
    if (member) {
      const float hL = h0 * (float)(1 << Lg);
      float gap = __int_as_float(0x7f800000);
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
      if (Lg >= kTopLevel || best.worst() < cov2 || cov2 >= max_sqd) done = true;
      else Lmin = Lg + 1;
    }
    WKNN_STAT(4, __popc(__ballot_sync(FULL, member && !done)));
  }
}

}  // namespace ngicp
