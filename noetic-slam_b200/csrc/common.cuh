// Shared device/host definitions for libngicp_b200 (sm_100a only).
//
// Data layout in HBM:
//   points      float4 per point, Morton-sorted; .w carries the ORIGINAL index (int bits)
//   covariances 6 x fp32 per point (xx,xy,xz,yy,yz,zz), same (sorted) order as the points
//   voxel index open-addressing hash of CellSlot{key,start,end} over ALL levels >= base_level of an
//               implicit octree: a level-L cell (side h0 * 2^L) is a contiguous range of the
//               Morton-sorted array, so one sorted array serves every level.
// Replaces the nanoflann KD-tree of the reference (src/dlio/include/nano_gicp/nanoflann.h:1025-1185,
// 1405-1417 build; :1436-1460,1587-1666 search).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngicp {

constexpr int kBitsPerAxis = 12;                  // integer voxel coordinate bits per axis
constexpr int kMortonBits = 3 * kBitsPerAxis;     // 36
constexpr int kTopLevel = kBitsPerAxis;           // one cell spans the whole grid at this level
constexpr int kNumLevels = kTopLevel + 1;         // levels 0..12
constexpr int kMaxSegBits = 20;                   // keyframes per batched index
constexpr int kSortLevel = 0;                     // keys are sorted on all 36 bits (four 9-bit passes): where a scan is dense (the
                                                  // returns next to the sensor) the self k-NN needs cells finer than the base level
constexpr int kBaseFloor = 3;                     // the base level (correspondence / public searches) is never finer than this:
                                                  // 512 cells per axis; the bounded 1-NN search is tuned to cells of that size
constexpr int kMaxCoord = (1 << kBitsPerAxis) - 1;
constexpr unsigned long long kEmptyKey = ~0ull;

struct __align__(16) CellSlot {
  unsigned long long key;  // (((seg << 36 | morton) >> 3L) << 4) | L ; kEmptyKey if free
  uint32_t start;          // first sorted position of the cell
  uint32_t end;            // one past the last
};

// Device-resident description of one index, produced by the build kernels without a host round trip.
struct GridMeta {
  float h0;          // level-0 cell side, a power of two
  float inv_h0;
  float margin;      // absolute slack subtracted from every "covered radius" (fp32 rounding of keys)
  int base_level;    // search level of the correspondence / public k-NN searches: finest level with mean occupancy >= 2
  unsigned int level_hist[16];  // level_hist[d] = #sorted positions whose key first differs from the
                                // predecessor at level d-1 (d = 13: differs at the top / first point)
  unsigned int cells_total;     // entries inserted into the hash
  int fine_level;               // finest level present in the hash (<= base_level): as fine as the table has room for; only
                                // the leaf-scheduled self k-NN (lknn.cuh) goes below base_level, where the cloud is dense
  unsigned int pad[2];
};

// Read-only view handed to the kernels by value.
struct GridView {
  const float4* pts;        // [n] sorted, .w = original index
  const int* inv;           // [n] original index -> sorted position
  const CellSlot* table;    // [table_mask+1]
  const GridMeta* meta;
  const float4* seg_origin; // [n_seg] lower bbox corner of each segment (keyframe)
  const int* seg_start;     // [n_seg+1] sorted-position boundaries (== original boundaries)
  uint32_t table_mask;
  int n;
  int n_seg;
};

// ------------------------------------------------------------------------------------ morton
__host__ __device__ __forceinline__ unsigned long long expand3(unsigned int v) {
  unsigned long long x = v & 0x1fffffull;
  x = (x | (x << 32)) & 0x1f00000000ffffull;
  x = (x | (x << 16)) & 0x1f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}
__host__ __device__ __forceinline__ unsigned long long morton3(unsigned int x, unsigned int y, unsigned int z) {
  return expand3(x) | (expand3(y) << 1) | (expand3(z) << 2);
}

// Voxel key spec (bit-exact with oracle: tests/test_keys*.py):
//   u_a  = fl32(p_a - origin_a)                  (round-to-nearest, no FMA)
//   c_a  = clamp((int)floor(fl32(u_a * inv_h0)), 0, 4095)     inv_h0 is a power of two => exact
//   key  = (seg << 36) | morton3(c_x, c_y, c_z)
__device__ __forceinline__ int voxel_coord_unclamped(float p, float o, float inv_h0) {
  float t = floorf(__fmul_rn(__fsub_rn(p, o), inv_h0));
  t = fminf(fmaxf(t, -1073741824.0f), 1073741824.0f);
  return (int)t;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// inverse of expand3: every third bit of m, compacted
__host__ __device__ __forceinline__ unsigned int compact3(unsigned long long m) {
  unsigned long long x = m & 0x1249249249249249ull;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
  x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
  x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
  x = (x ^ (x >> 32)) & 0x1fffffull;
  return (unsigned int)x;
}
// Hash key of a cell: plain packed level-L coordinates (no bit interleaving on the query side — the searches
// probe dozens of cells per pass and Morton expansion was 40% of their instructions). The sort key stays Morton.
__device__ __forceinline__ unsigned long long pack_cell(unsigned int seg, int level, unsigned int cx, unsigned int cy, unsigned int cz) {
  return ((((unsigned long long)seg << 36) | ((unsigned long long)cz << 24) | ((unsigned long long)cy << 12) | (unsigned long long)cx) << 4) |
         (unsigned long long)level;
}
// the same key from a sorted (segment, Morton) key — used by the build, once per cell
__device__ __forceinline__ unsigned long long cell_key(unsigned long long seg_morton, int level) {
  const unsigned long long m = seg_morton & ((1ull << kMortonBits) - 1ull);
  return pack_cell((unsigned int)(seg_morton >> kMortonBits), level, compact3(m) >> level, compact3(m >> 1) >> level, compact3(m >> 2) >> level);
}
__device__ __forceinline__ uint32_t hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

__device__ __forceinline__ bool cell_lookup(const CellSlot* __restrict__ table, uint32_t mask, unsigned long long ckey,
                                            uint32_t& start, uint32_t& end) {
  uint32_t h = hash64(ckey) & mask;
  for (;;) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(table + h));
    const unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
    if (k == ckey) { start = raw.z; end = raw.w; return true; }
    if (k == kEmptyKey) return false;
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ int find_segment(const int* __restrict__ seg_start, int n_seg, int i) {
  if (n_seg <= 1) return 0;
  int lo = 0, hi = n_seg;  // largest s with seg_start[s] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seg_start + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// exact reference metric (nanoflann.h:509-520): ((dx*dx)+(dy*dy))+(dz*dz), fp32, no FMA, query minus data
__device__ __forceinline__ float sqdist_ref(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ------------------------------------------------------------------------------------ top-K
// Register-resident sorted list of the K best (distance, ORIGINAL index) pairs, ascending by
// (distance, index) — the documented tie-break. The original index rides in the .w lane of every
// staged point, so ties are broken in registers; kernels that need sorted positions translate
// through the index's inverse permutation when they write their results. Empty slots hold
// (+inf, -1); indices compare as unsigned so -1 sorts last.
template <int K>
struct TopK {
  static constexpr int kK = K;
  float d[K];
  int p[K];
  // shift register: drop d[0], append at the back (static indices only)
  __device__ __forceinline__ void append_shift(float dn, int pn) {
#pragma unroll
    for (int i = 0; i + 1 < K; i++) { d[i] = d[i + 1]; p[i] = p[i + 1]; }
    d[K - 1] = dn; p[K - 1] = pn;
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int i = 0; i + 1 < K; i++) { d[i] = d[i + 1]; p[i] = p[i + 1]; }
    d[K - 1] = __int_as_float(0x7f800000); p[K - 1] = -1;
  }
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < K; i++) { d[i] = __int_as_float(0x7f800000); p[i] = -1; }
  }
  __device__ __forceinline__ static bool before(float da, int pa, float db, int pb) {
    return da < db || (da == db && (unsigned)pa < (unsigned)pb);
  }
  // slow path: full (distance, original index) ordering, only taken when a distance is exactly tied
  __device__ __forceinline__ void offer_tied(float dn, int pn) {
    if (!before(dn, pn, d[K - 1], p[K - 1])) return;
    d[K - 1] = dn; p[K - 1] = pn;
#pragma unroll
    for (int i = K - 1; i > 0; i--) {
      if (before(d[i], p[i], d[i - 1], p[i - 1])) {
        const float td = d[i]; d[i] = d[i - 1]; d[i - 1] = td;
        const int tp = p[i]; p[i] = p[i - 1]; p[i - 1] = tp;
      }
    }
  }
  // fast path: rank-and-shift with independent selects (no dependent compare-swap chain)
  __device__ __forceinline__ void offer(float dn, int pn) {
    if (dn > d[K - 1]) return;
    bool tie = false;
#pragma unroll
    for (int i = 0; i < K; i++) tie |= (d[i] == dn);
    if (tie) { offer_tied(dn, pn); return; }
#pragma unroll
    for (int i = K - 1; i > 0; i--) {
      const bool lti = d[i] < dn, ltm = d[i - 1] < dn;
      d[i] = lti ? d[i] : (ltm ? dn : d[i - 1]);
      p[i] = lti ? p[i] : (ltm ? pn : p[i - 1]);
    }
    const bool lt0 = d[0] < dn;
    d[0] = lt0 ? d[0] : dn;
    p[0] = lt0 ? p[0] : pn;
  }
  __device__ __forceinline__ float worst() const { return d[K - 1]; }
  __device__ __forceinline__ float kth(int k) const {  // k in 1..K
    float v = d[K - 1];
#pragma unroll
    for (int i = 0; i < K; i++) if (i == k - 1) v = d[i];
    return v;
  }
};

// The same list as packed 64-bit keys (distance bits << 32 | original index): squared distances are non-negative, so
// the unsigned integer order of the keys IS the (distance, index) order — no separate tie path, and two sorted lists
// merge with a plain compare-exchange network. Used for the per-lane partial lists of the warp search.
template <int K>
struct KeyList {
  static constexpr int KP = K <= 1 ? 1 : (K <= 2 ? 2 : (K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 16 ? 16 : (K <= 32 ? 32 : 64)))));   // network size
  unsigned long long k[KP];
  static constexpr unsigned long long kEmpty = (0x7f800000ull << 32) | 0xffffffffull;   // (+inf, -1)
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int i = 0; i < KP; i++) k[i] = kEmpty;
  }
  __device__ __forceinline__ float worst() const { return __uint_as_float((unsigned int)(k[K - 1] >> 32)); }
  __device__ __forceinline__ void offer(float dn, int pn) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(dn) << 32) | (unsigned int)pn;
    if (key >= k[K - 1]) return;
#pragma unroll
    for (int i = K - 1; i > 0; i--) {
      const bool lti = k[i] < key, ltm = k[i - 1] < key;
      k[i] = lti ? k[i] : (ltm ? key : k[i - 1]);
    }
    k[0] = k[0] < key ? k[0] : key;
  }
  static __device__ __forceinline__ unsigned long long make_key(float d, int p) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)p;
  }
  // full bitonic sorting network over the KP slots (static indices only): used once per scan to seed the list from the
  // first KP candidates, which is several times cheaper than KP insertions
  __device__ __forceinline__ void sort_all() {
#pragma unroll
    for (int size = 2; size <= KP; size <<= 1) {
#pragma unroll
      for (int stride = size / 2; stride > 0; stride >>= 1) {
#pragma unroll
        for (int i = 0; i < KP; i++) {
          const int j = i ^ stride;
          if (j > i) {
            const bool up = (i & size) == 0;
            const unsigned long long a = k[i], b = k[j];
            const bool sw = up ? (b < a) : (a < b);
            k[i] = sw ? b : a;
            k[j] = sw ? a : b;
          }
        }
      }
    }
  }
  __device__ __forceinline__ float dist_at(int i) const { return __uint_as_float((unsigned int)(k[i] >> 32)); }
  // Both lanes of an xor pair end up with the K smallest keys of their two sorted lists, sorted.
  __device__ __forceinline__ void merge_with_partner(int off) {
    unsigned long long o[KP];
#pragma unroll
    for (int i = 0; i < KP; i++) o[i] = __shfl_xor_sync(0xffffffffu, k[i], off);
#pragma unroll
    for (int i = 0; i < KP; i++) k[i] = k[i] < o[KP - 1 - i] ? k[i] : o[KP - 1 - i];   // KP smallest, bitonic
#pragma unroll
    for (int s = KP / 2; s >= 1; s >>= 1) {
#pragma unroll
      for (int i = 0; i < KP; i++) {
        if ((i & s) == 0) {
          const unsigned long long a = k[i], b = k[i + s];
          const bool lt = a < b;
          k[i] = lt ? a : b;
          k[i + s] = lt ? b : a;
        }
      }
    }
  }
  __device__ __forceinline__ void store(TopK<K>& t) const {
#pragma unroll
    for (int i = 0; i < K; i++) { t.d[i] = __uint_as_float((unsigned int)(k[i] >> 32)); t.p[i] = (int)(unsigned int)k[i]; }
  }
};

// Slow-path list for k > 32 (the reference accepts any k): same ordering, runtime capacity, lives
// in local memory.
template <int KMAX>
struct TopKDyn {
  float d[KMAX];
  int p[KMAX];
  int cap;
  __device__ __forceinline__ void reset() {
    for (int i = 0; i < cap; i++) { d[i] = __int_as_float(0x7f800000); p[i] = -1; }
  }
  __device__ __forceinline__ void offer(float dn, int pn) {
    if (dn > d[cap - 1]) return;
    if (!TopK<1>::before(dn, pn, d[cap - 1], p[cap - 1])) return;
    int i = cap - 1;
    while (i > 0 && TopK<1>::before(dn, pn, d[i - 1], p[i - 1])) { d[i] = d[i - 1]; p[i] = p[i - 1]; i--; }
    d[i] = dn; p[i] = pn;
  }
  __device__ __forceinline__ float worst() const { return d[cap - 1]; }
  __device__ __forceinline__ float kth(int k) const { return d[k - 1]; }
};

// ------------------------------------------------------------------------------------ search
// Exact k-NN of one query inside one segment of a grid. Strategy:
//   1. start at the finest level whose own cell holds >= start_count points,
//   2. visit the 3x3x3 block of cells around the query at that level,
//   3. the block guarantees every unvisited point is farther than `covered` (distance to the
//      nearest block face that still has grid behind it); if the k-th best is closer, done,
//   4. otherwise restart one level up (cells twice as large). The top level is one cell = exhaustive.
// max_sqd bounds the search radius for correspondence queries: once `covered^2 >= max_sqd` nothing
// unvisited can be a valid correspondence.
template <class TK>
__device__ __forceinline__ void grid_knn(const GridView& g, float qx, float qy, float qz, int seg, int k, int start_count,
                                         float max_sqd, TK& best) {
  const GridMeta* __restrict__ m = g.meta;
  const float h0 = __ldg(&m->h0), inv_h0 = __ldg(&m->inv_h0), margin = __ldg(&m->margin);
  const int base = __ldg(&m->base_level);
  const float4 o = __ldg(g.seg_origin + seg);
  const float ux = __fsub_rn(qx, o.x), uy = __fsub_rn(qy, o.y), uz = __fsub_rn(qz, o.z);
  const int c0x = voxel_coord_unclamped(qx, o.x, inv_h0);
  const int c0y = voxel_coord_unclamped(qy, o.y, inv_h0);
  const int c0z = voxel_coord_unclamped(qz, o.z, inv_h0);
  int L = base;
  // 1. start level from own-cell occupancy
  for (; L < kTopLevel; L++) {
    // a block at this level already covers the whole correspondence radius: never start coarser
    const float reach = h0 * (float)(1 << L) - margin;
    if (reach * reach * 0.999999f >= max_sqd) break;
    const int maxc = kMaxCoord >> L;
    const unsigned int cx = clampi(c0x >> L, 0, maxc), cy = clampi(c0y >> L, 0, maxc), cz = clampi(c0z >> L, 0, maxc);
    uint32_t s, e;
    const unsigned long long ck = pack_cell((unsigned)seg, L, cx, cy, cz);
    if (cell_lookup(g.table, g.table_mask, ck, s, e) && (int)(e - s) >= start_count) break;
  }

  for (;; L++) {
    best.reset();
    const int maxc = kMaxCoord >> L;
    const int cx = clampi(c0x >> L, 0, maxc), cy = clampi(c0y >> L, 0, maxc), cz = clampi(c0z >> L, 0, maxc);
#pragma unroll 1
    for (int az = cz - 1; az <= cz + 1; az++) {
      if (az < 0 || az > maxc) continue;
#pragma unroll 1
      for (int ay = cy - 1; ay <= cy + 1; ay++) {
        if (ay < 0 || ay > maxc) continue;
#pragma unroll 1
        for (int ax = cx - 1; ax <= cx + 1; ax++) {
          if (ax < 0 || ax > maxc) continue;
          const unsigned long long ck = pack_cell((unsigned)seg, L, (unsigned)ax, (unsigned)ay, (unsigned)az);
          uint32_t s, e;
          if (!cell_lookup(g.table, g.table_mask, ck, s, e)) continue;
          for (uint32_t j = s; j < e; j++) {
            const float4 p = __ldg(g.pts + j);
            best.offer(sqdist_ref(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
          }
        }
      }
    }
    if (L >= kTopLevel) break;
    // 3. covered radius of the block
    const float hL = h0 * (float)(1 << L);
    float gap = __int_as_float(0x7f800000);
    {
      const float lo = (float)(cx - 1) * hL, hi = (float)(cx + 2) * hL;
      if (cx - 1 > 0) gap = fminf(gap, fmaxf(ux - lo, 0.0f));
      if (cx + 1 < maxc) gap = fminf(gap, fmaxf(hi - ux, 0.0f));
    }
    {
      const float lo = (float)(cy - 1) * hL, hi = (float)(cy + 2) * hL;
      if (cy - 1 > 0) gap = fminf(gap, fmaxf(uy - lo, 0.0f));
      if (cy + 1 < maxc) gap = fminf(gap, fmaxf(hi - uy, 0.0f));
    }
    {
      const float lo = (float)(cz - 1) * hL, hi = (float)(cz + 2) * hL;
      if (cz - 1 > 0) gap = fminf(gap, fmaxf(uz - lo, 0.0f));
      if (cz + 1 < maxc) gap = fminf(gap, fmaxf(hi - uz, 0.0f));
    }
    const float covered = fmaxf(gap - margin, 0.0f);
    const float cov2 = covered * covered * 0.999999f;
    if (best.worst() < cov2) break;   // every unvisited point is strictly farther than the k-th best
    if (cov2 >= max_sqd) break;      // nothing unvisited can be within the correspondence radius
  }
}

}  // namespace ngicp
