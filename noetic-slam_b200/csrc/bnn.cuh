// Bounded exact nearest neighbour: the correspondence search of K4a (reference nano_gicp.cc:219-227, one
// nanoflann knnSearch(k = 1) per source point and LM iteration).
//
// Every point that can beat or tie a known candidate at squared distance d lies in the closed ball of radius sqrt(d)
// around the query, and nothing farther than the correspondence gate matters at all. With a bound in hand the search
// is one pass over the voxels the ball touches, most of which are ruled out by their box distance before they are
// probed. The bound is the correspondence of the previous LM iteration re-measured at the new pose, or a point of
// the query's own base-level cell (or its parent), clipped to max_sqd.
//
// Two stages, two kernels (linearize.cu):
//   bnn_group  4 lanes per query, registers only. Resolves the LIGHT queries: ball within <= 32 base-level cells
//              (batches of four hash probes per lane, then the points of those cells as one flat list, four loads
//              in flight). A query never waits for the other queries of its warp beyond that fixed, short chain.
//              Everything else is HEAVY: its best known bound is left behind as a hint and its index is appended to
//              a list.
//   bnn_warp   one warp per heavy query: up to 5x5x5 cells at the finest level that allows it, four probes per lane
//              in flight, the non-empty voxel buckets staged with TMA bulk copies (wknn.cuh machinery) and scanned by
//              all 32 lanes. Declines only when the ball is larger than that; the caller then runs warp_knn.
// Exactness: same (distance, original index) order and the same fp32 metric as everywhere else; the cell range and
// the box-distance test carry the key-rounding margin of GridMeta (doubled in the box test).
#pragma once
#include "wknn.cuh"

namespace ngicp {

constexpr int kBnnSeedPerLane = 8;            // seed points examined per lane (own cell, 4 lanes)
constexpr int kBnnLightCells = 32;            // a light query touches at most this many base-level cells (two batches of 4 probes per lane)
constexpr int kBnnSpan = 5;                   // heavy: cells per axis the ball may span at the chosen level
constexpr int kBnnMaxLevelsAboveBase = 2;

struct NNBest {
  float d;   // squared distance (fp32 reference metric)
  int i;     // original index (tie-break)
  int pos;   // sorted position (what the correspondence array stores), -1 if unknown / none
};
__device__ __forceinline__ void nn_offer(NNBest& b, float d, int i, int pos) {
  if (TopK<1>::before(d, i, b.d, b.i)) { b.d = d; b.i = i; b.pos = pos; }
}
template <int W>
__device__ __forceinline__ void nn_reduce(NNBest& b, unsigned mask) {
#pragma unroll
  for (int off = W / 2; off > 0; off >>= 1) {
    const float od = __shfl_xor_sync(mask, b.d, off);
    const int oi = __shfl_xor_sync(mask, b.i, off);
    const int op = __shfl_xor_sync(mask, b.pos, off);
    nn_offer(b, od, oi, op);
  }
}

// hash probe whose first load was issued by the caller (raw); walks the collision chain if needed
__device__ __forceinline__ void bnn_resolve(const GridView& g, unsigned long long ck, uint32_t h, uint4 raw, uint32_t& s, uint32_t& e) {
  s = 0; e = 0;
  unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
  while (k != kEmptyKey) {
    if (k == ck) { s = raw.z; e = raw.w; return; }
    h = (h + 1) & g.table_mask;
    raw = __ldg(reinterpret_cast<const uint4*>(g.table + h));
    k = ((unsigned long long)raw.y << 32) | raw.x;
  }
}

// lower bound (squared, deflated) of the distance from the query (grid coordinates u) to anything stored in a cell
__device__ __forceinline__ float bnn_cell_lb(int ix, int iy, int iz, float hL, float slack, float ux, float uy, float uz) {
  const float ex = fmaxf(fmaxf((float)ix * hL - slack - ux, ux - ((float)(ix + 1) * hL + slack)), 0.0f);
  const float ey = fmaxf(fmaxf((float)iy * hL - slack - uy, uy - ((float)(iy + 1) * hL + slack)), 0.0f);
  const float ez = fmaxf(fmaxf((float)iz * hL - slack - uz, uz - ((float)(iz + 1) * hL + slack)), 0.0f);
  return (ex * ex + ey * ey + ez * ez) * 0.999999f;
}

// ceil(65536 / n) for the 1..5 cells a ball spans per axis: c / n == (c * r) >> 16 for c < 125, without an integer division
__device__ __forceinline__ unsigned int bnn_recip16(int n) {
  return n <= 1 ? 65536u : (n == 2 ? 32768u : (n == 3 ? 21846u : (n == 4 ? 16384u : 13108u)));
}

struct BallCells {
  int lox, loy, loz, nx, ny, nz, total;
};
__device__ __forceinline__ BallCells bnn_cells(float ux, float uy, float uz, float r, float inv_hL, int maxc) {
  BallCells c;
  c.lox = max((int)floorf((ux - r) * inv_hL), 0);
  c.loy = max((int)floorf((uy - r) * inv_hL), 0);
  c.loz = max((int)floorf((uz - r) * inv_hL), 0);
  c.nx = min((int)floorf((ux + r) * inv_hL), maxc) - c.lox + 1;
  c.ny = min((int)floorf((uy + r) * inv_hL), maxc) - c.loy + 1;
  c.nz = min((int)floorf((uz + r) * inv_hL), maxc) - c.loz + 1;
  c.total = (c.nx > 0 && c.ny > 0 && c.nz > 0) ? c.nx * c.ny * c.nz : 0;
  return c;
}

// ---------------------------------------------------------------------------------------------- stage 1
// Returns true when the query is resolved (best = exact nearest neighbour within max_sqd, or none), false when it is
// heavy (best = tightest bound found so far, to be stored as a hint). All 4 lanes of a group hold identical inputs
// and return identical results. have_hint: best already holds a bound (d, i, pos) from the caller.
__device__ __forceinline__ bool bnn_group(const GridView& g, float qx, float qy, float qz, int seg, bool have_hint, float max_sqd, NNBest& best) {
  constexpr int LPQ = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPQ - 1);
  const unsigned gmask = 0xfu << (lane & ~(LPQ - 1));
  const GridMeta* __restrict__ m = g.meta;
  const float h0 = __ldg(&m->h0), inv_h0 = __ldg(&m->inv_h0), margin = __ldg(&m->margin);
  const int base = __ldg(&m->base_level);
  const float4 o = __ldg(g.seg_origin + seg);
  const float ux = __fsub_rn(qx, o.x), uy = __fsub_rn(qy, o.y), uz = __fsub_rn(qz, o.z);
  int seed_level = -1, sx = 0, sy = 0, sz = 0;   // cell the seed step scanned completely
  const float hB = h0 * (float)(1 << base);
  if (!have_hint) { best.d = __int_as_float(0x7f800000); best.i = -1; best.pos = -1; }
  // a hint farther than half a base cell (the pose moved a lot since it was found) is worth a look at the own cell too
  if (!have_hint || !(best.d <= 0.25f * hB * hB)) {
    const int c0x = clampi(voxel_coord_unclamped(qx, o.x, inv_h0), 0, kMaxCoord);
    const int c0y = clampi(voxel_coord_unclamped(qy, o.y, inv_h0), 0, kMaxCoord);
    const int c0z = clampi(voxel_coord_unclamped(qz, o.z, inv_h0), 0, kMaxCoord);
    // own cell at the base level and its parent, both probes in flight together
    unsigned long long ck[2];
    uint32_t hh[2];
    uint4 raw[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
      ck[t] = pack_cell((unsigned)seg, base + t, (unsigned)(c0x >> (base + t)), (unsigned)(c0y >> (base + t)), (unsigned)(c0z >> (base + t)));
      hh[t] = hash64(ck[t]) & g.table_mask;
    }
#pragma unroll
    for (int t = 0; t < 2; t++) raw[t] = __ldg(reinterpret_cast<const uint4*>(g.table + hh[t]));
    uint32_t s, e;
    int lvl = base;
    bnn_resolve(g, ck[0], hh[0], raw[0], s, e);
    if (e == s) { bnn_resolve(g, ck[1], hh[1], raw[1], s, e); lvl = base + 1; }
    const uint32_t cnt = e - s;
    if (cnt > 0) {
      if (cnt <= (uint32_t)(kBnnSeedPerLane * LPQ)) { seed_level = lvl; sx = c0x >> lvl; sy = c0y >> lvl; sz = c0z >> lvl; }
#pragma unroll
      for (int half = 0; half < kBnnSeedPerLane / 4; half++) {
        if ((uint32_t)(LPQ * 4 * half) >= cnt) break;   // uniform in the group
        float4 p[4];
        uint32_t jj[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          jj[t] = (uint32_t)(sub + LPQ * (4 * half + t));
          p[t] = __ldg(g.pts + s + min(jj[t], cnt - 1));
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (jj[t] < cnt) nn_offer(best, sqdist_ref(qx, qy, qz, p[t].x, p[t].y, p[t].z), __float_as_int(p[t].w), (int)(s + jj[t]));
      }
      nn_reduce<LPQ>(best, gmask);
    }
  }
  if (!(best.d <= max_sqd)) { best.d = max_sqd; best.i = -1; best.pos = -1; }   // also "no bound at all" (+inf)
  const float r = __fsqrt_ru(best.d) * 1.000001f + margin;   // covers the fp32 rounding of the metric and of the keys
  const float hL = hB;
  if (sub == 0) { BNN_STAT(0, 1); if (best.i < 0) BNN_STAT(5, 1); if (have_hint) BNN_STAT(4, 1); }
  if (!(r <= 2.0f * hL)) { if (sub == 0) BNN_STAT(1, 1); return false; }                       // also r = +inf / NaN
  const float inv_hL = inv_h0 / (float)(1 << base);          // powers of two: exact
  const BallCells bc = bnn_cells(ux, uy, uz, r, inv_hL, kMaxCoord >> base);
  if (bc.total > kBnnLightCells) { if (sub == 0) BNN_STAT(2, 1); return false; }
  // ---- batches of four probes per lane
  const unsigned int rx = bnn_recip16(bc.nx), ry = bnn_recip16(bc.ny);   // c / nx == (c * rx) >> 16 for c < 125
  const float slack = 2.0f * margin;
  constexpr int NB = 4;
  for (int cb = 0; cb < bc.total; cb += NB * LPQ) {          // uniform in the group
  unsigned long long ck[NB];
  uint32_t hh[NB];
  uint4 raw[NB];
  bool want[NB];
#pragma unroll
  for (int t = 0; t < NB; t++) {
    if (t >= 2 && bc.total <= cb + 2 * LPQ) {                  // the usual 2x2x2 ball: two cells per lane, skip the other two slots
      want[t] = false; ck[t] = 0ull; hh[t] = 0u;
      continue;
    }
    const int c = cb + sub + LPQ * t;
    const int cyz = (int)(((unsigned)c * rx) >> 16);
    const int ox = c - cyz * bc.nx;
    const int oz = (int)(((unsigned)cyz * ry) >> 16);
    const int oy = cyz - oz * bc.ny;
    const int ix = bc.lox + ox, iy = bc.loy + oy, iz = bc.loz + oz;
    want[t] = c < bc.total && !(bnn_cell_lb(ix, iy, iz, hL, slack, ux, uy, uz) > best.d) &&
              !(base == seed_level && ix == sx && iy == sy && iz == sz);
    ck[t] = pack_cell((unsigned)seg, base, (unsigned)ix, (unsigned)iy, (unsigned)iz);
    hh[t] = hash64(ck[t]) & g.table_mask;
  }
#pragma unroll
  for (int t = 0; t < NB; t++) raw[t] = want[t] ? __ldg(reinterpret_cast<const uint4*>(g.table + hh[t])) : make_uint4(~0u, ~0u, 0u, 0u);
  uint32_t st[NB], pre[NB + 1];
  pre[0] = 0;
#pragma unroll
  for (int t = 0; t < NB; t++) {
    uint32_t e;
    bnn_resolve(g, ck[t], hh[t], raw[t], st[t], e);
    pre[t + 1] = pre[t] + (e - st[t]);
  }
  // ---- the points of this lane's cells as one flat list, four independent loads in flight per step
  const uint32_t tot = pre[NB];
  for (uint32_t i0 = 0; i0 < tot; i0 += 4) {
    float4 p[4];
    uint32_t pos[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t i = min(i0 + u, tot - 1);
      uint32_t a = st[0] + i;
#pragma unroll
      for (int t = 1; t < NB; t++) a = i >= pre[t] ? st[t] + (i - pre[t]) : a;
      pos[u] = a;
      p[u] = __ldg(g.pts + a);
    }
#pragma unroll
    for (int u = 0; u < 4; u++)   // re-offering a duplicate of the last point is harmless
      nn_offer(best, sqdist_ref(qx, qy, qz, p[u].x, p[u].y, p[u].z), __float_as_int(p[u].w), (int)pos[u]);
  }
  }
  nn_reduce<LPQ>(best, gmask);
  return true;
}

// ---------------------------------------------------------------------------------------------- stage 2
// One warp, one query (all lanes hold identical inputs). best: in = bound already clipped to max_sqd (i = pos = -1 if
// it is only the clip), out = exact nearest neighbour within the bound. Returns false when the ball is too large.
__device__ __forceinline__ bool bnn_warp(const GridView& g, float qx, float qy, float qz, int seg, NNBest& best, WarpScratch& ws, uint32_t& phase) {
  const float bd = best.d;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const GridMeta* __restrict__ m = g.meta;
  const float h0 = __ldg(&m->h0), inv_h0 = __ldg(&m->inv_h0), margin = __ldg(&m->margin);
  const int base = __ldg(&m->base_level);
  const float4 o = __ldg(g.seg_origin + seg);
  const float ux = __fsub_rn(qx, o.x), uy = __fsub_rn(qy, o.y), uz = __fsub_rn(qz, o.z);
  const float r = __fsqrt_ru(bd) * 1.000001f + margin;
  // finest level at which the ball spans at most kBnnSpan cells per axis
  int L = base;
  float hL = h0 * (float)(1 << base);
  while (hL * (0.5f * (kBnnSpan - 1)) < r && L < base + kBnnMaxLevelsAboveBase && L < kTopLevel) { hL *= 2.0f; L++; }
  if (!(hL * (0.5f * (kBnnSpan - 1)) >= r)) { if (lane == 0) BNN_STAT(10, 1); return false; }
  if (lane == 0) { BNN_STAT(6, 1); BNN_STAT(7 + (L - base), 1); }
  const float inv_hL = inv_h0 / (float)(1 << L);
  const BallCells bc = bnn_cells(ux, uy, uz, r, inv_hL, kMaxCoord >> L);
  if (bc.nx > kBnnSpan || bc.ny > kBnnSpan || bc.nz > kBnnSpan) return false;   // cannot happen; stay exact if it ever does
  const unsigned int rx = bnn_recip16(bc.nx), ry = bnn_recip16(bc.ny);
  const float slack = 2.0f * margin;
  constexpr int NB = (kBnnSpan * kBnnSpan * kBnnSpan + 31) / 32;   // 4 probes per lane, all in flight
  unsigned long long ck[NB];
  uint32_t hh[NB];
  uint4 raw[NB];
  bool want[NB];
#pragma unroll
  for (int t = 0; t < NB; t++) {
    const int c = lane + 32 * t;
    const int cyz = (int)(((unsigned)c * rx) >> 16);
    const int ox = c - cyz * bc.nx;
    const int oz = (int)(((unsigned)cyz * ry) >> 16);
    const int oy = cyz - oz * bc.ny;
    const int ix = bc.lox + ox, iy = bc.loy + oy, iz = bc.loz + oz;
    want[t] = c < bc.total && !(bnn_cell_lb(ix, iy, iz, hL, slack, ux, uy, uz) > bd);
    ck[t] = pack_cell((unsigned)seg, L, (unsigned)ix, (unsigned)iy, (unsigned)iz);
    hh[t] = hash64(ck[t]) & g.table_mask;
  }
#pragma unroll
  for (int t = 0; t < NB; t++) raw[t] = want[t] ? __ldg(reinterpret_cast<const uint4*>(g.table + hh[t])) : make_uint4(~0u, ~0u, 0u, 0u);
  uint32_t st[NB], cnt[NB];
#pragma unroll
  for (int t = 0; t < NB; t++) {
    uint32_t e;
    bnn_resolve(g, ck[t], hh[t], raw[t], st[t], e);
    cnt[t] = e - st[t];
  }
  // compact the non-empty buckets into the warp's range list (wknn.cuh: rstart / rpre), order (t, lane)
  const unsigned lt = (1u << lane) - 1u;
  uint32_t M = 0;
  int R = 0;
#pragma unroll
  for (int t = 0; t < NB; t++) {
    uint32_t inc = cnt[t];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t a = __shfl_up_sync(FULL, inc, off);
      if (lane >= off) inc += a;
    }
    const unsigned nz = __ballot_sync(FULL, cnt[t] != 0);
    if (cnt[t]) { const int i = R + __popc(nz & lt); ws.rstart[i] = st[t]; ws.rpre[i] = M + inc - cnt[t]; }
    M += __shfl_sync(FULL, inc, 31);
    R += __popc(nz);
  }
  if (lane == 0) ws.rpre[R] = M;
  if (lane == 0) { BNN_STAT(11, M); BNN_STAT(12, R); BNN_STAT(13, bc.total); }
  __syncwarp();
  NNBest mine;   // pos = number of the candidate in the concatenated bucket list
  mine.d = __int_as_float(0x7f800000); mine.i = -1; mine.pos = -1;
  WKNN_FOR_CHUNKS({
#pragma unroll
    for (int t = 0; t < kWarpChunk / 32; t++) {
      const int e = lane + 32 * t;
      const float4 c = P[min(e, nch - 1)];
      const float d = sqdist_ref(qx, qy, qz, c.x, c.y, c.z);
      if (e < nch) nn_offer(mine, d, __float_as_int(c.w), (int)c0 + e);
    }
  })
  nn_reduce<32>(mine, FULL);
  if (mine.i >= 0 && TopK<1>::before(mine.d, mine.i, best.d, best.i)) {
    // candidate number -> sorted position: the bucket that holds it (shared-memory binary search, uniform)
    int lo = 0, hi = R;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ws.rpre[mid] <= (uint32_t)mine.pos) lo = mid; else hi = mid;
    }
    best.d = mine.d; best.i = mine.i;
    best.pos = (int)(ws.rstart[lo] + ((uint32_t)mine.pos - ws.rpre[lo]));
  }
  __syncwarp();
  return true;
}

}  // namespace ngicp
