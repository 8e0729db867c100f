// K4 / K5 — fused correspondence search + GICP linearisation, and the cached-correspondence error.
// Replaces (reference src/dlio/src/nano_gicp/nano_gicp.cc):
//   update_correspondences :206-245  fp32 transform, 1-NN in the target, distance gate,
//                                     RCR = C_B + R C_A R^T, Mahalanobis = RCR^-1
//   linearize              :248-302  e = p_B - T p_A, J = [skew(T p_A) | -I], H += J^T M J, b += J^T M e
//   compute_error          :305-326  sum e^T M e at a trial pose with the cached correspondences
// One thread per source point (Morton order, so a warp probes neighbouring target cells). The 21
// upper-triangular terms of H, the 6 of b, the error and the correspondence count are reduced with
// warp shuffles, then per block, then by the last block to finish (fixed order => run-to-run
// deterministic, unlike the reference's per-thread OpenMP partials), with Neumaier-compensated
// accumulation across blocks. The result lands in host-mapped pinned memory; the host's 6x6 LM
// solve waits on a sequence number, not on a stream synchronise.
// compute_error does not cache the Mahalanobis matrices: it rebuilds them from the pose of the last
// linearize with the very same device function, so both kernels see bit-identical M.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "linearize.cuh"
#include "bnn.cuh"

namespace ngicp {

namespace {

#ifdef NGICP_STATS
__device__ unsigned long long g_lin_hist[32];
__device__ unsigned int g_item_cycles[1 << 16];   // per work item: cycles of the search   // [0..15] per-item cycles (log2 buckets from 2^10), [16..31] per-warp totals
#endif

constexpr int kTerms = 29;  // 21 H + 6 b + error + count(c>0)

// symmetric RCR^-1 with explicit fma so K4 and K5 agree bit for bit
__device__ __forceinline__ void mahalanobis(const PoseArg& P, const float* __restrict__ ca, const float* __restrict__ cb, double M[6]) {
  const double a00 = ca[0], a01 = ca[1], a02 = ca[2], a11 = ca[3], a12 = ca[4], a22 = ca[5];
  const double* R = P.R;
  // T = R * C_A
  double T[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
    T[3 * r + 0] = __fma_rn(r2, a02, __fma_rn(r1, a01, __dmul_rn(r0, a00)));
    T[3 * r + 1] = __fma_rn(r2, a12, __fma_rn(r1, a11, __dmul_rn(r0, a01)));
    T[3 * r + 2] = __fma_rn(r2, a22, __fma_rn(r1, a12, __dmul_rn(r0, a02)));
  }
  // RCR = C_B + T R^T (upper triangle)
  double s[6];
  const int ri[6] = {0, 0, 0, 1, 1, 2}, ci[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) {
    const int r = ri[e], c = ci[e];
    const double v = __fma_rn(T[3 * r + 2], R[3 * c + 2], __fma_rn(T[3 * r + 1], R[3 * c + 1], __dmul_rn(T[3 * r], R[3 * c])));
    s[e] = __dadd_rn((double)cb[e], v);
  }
  const double xx = s[0], xy = s[1], xz = s[2], yy = s[3], yz = s[4], zz = s[5];
  const double c00 = __fma_rn(yy, zz, -__dmul_rn(yz, yz));
  const double c01 = __fma_rn(xz, yz, -__dmul_rn(xy, zz));
  const double c02 = __fma_rn(xy, yz, -__dmul_rn(xz, yy));
  const double c11 = __fma_rn(xx, zz, -__dmul_rn(xz, xz));
  const double c12 = __fma_rn(xy, xz, -__dmul_rn(xx, yz));
  const double c22 = __fma_rn(xx, yy, -__dmul_rn(xy, xy));
  const double det = __fma_rn(xz, c02, __fma_rn(xy, c01, __dmul_rn(xx, c00)));
  const double inv = __ddiv_rn(1.0, det);
  M[0] = __dmul_rn(c00, inv); M[1] = __dmul_rn(c01, inv); M[2] = __dmul_rn(c02, inv);
  M[3] = __dmul_rn(c11, inv); M[4] = __dmul_rn(c12, inv); M[5] = __dmul_rn(c22, inv);
}

// residual e = p_B - (R p_A + t) and q = R p_A + t, fp64 (nano_gicp.cc:266-271)
__device__ __forceinline__ void residual(const PoseArg& P, const float4& pa, const float4& pb, double q[3], double e[3]) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
    q[r] = __dadd_rn(__fma_rn(P.R[3 * r + 2], (double)pa.z, __fma_rn(P.R[3 * r + 1], (double)pa.y, __dmul_rn(P.R[3 * r], (double)pa.x))), P.t[r]);
  }
  e[0] = __dsub_rn((double)pb.x, q[0]); e[1] = __dsub_rn((double)pb.y, q[1]); e[2] = __dsub_rn((double)pb.z, q[2]);
}

__device__ __forceinline__ double quad_form(const double M[6], const double e[3], double Me[3]) {
  Me[0] = __fma_rn(M[2], e[2], __fma_rn(M[1], e[1], __dmul_rn(M[0], e[0])));
  Me[1] = __fma_rn(M[4], e[2], __fma_rn(M[3], e[1], __dmul_rn(M[1], e[0])));
  Me[2] = __fma_rn(M[5], e[2], __fma_rn(M[4], e[1], __dmul_rn(M[2], e[0])));
  return __fma_rn(e[2], Me[2], __fma_rn(e[1], Me[1], __dmul_rn(e[0], Me[0])));
}

__device__ __forceinline__ void cross3(const double a[3], double bx, double by, double bz, double o[3]) {
  o[0] = a[1] * bz - a[2] * by;
  o[1] = a[2] * bx - a[0] * bz;
  o[2] = a[0] * by - a[1] * bx;
}

// Deterministic per-segment sum (the density term of nano_gicp.cc:389): grid = (blocks per segment, segments), same
// fixed-order reduction and host-mapped result slot as K4b / K5.
__global__ void __launch_bounds__(kLinThreads) segment_sum_kernel(const double* __restrict__ in, const int* __restrict__ seg_start, double* __restrict__ partials,
                                                                   unsigned int* __restrict__ counters, ReduceSlot* __restrict__ slots, unsigned long long seq) {
  __shared__ double wsum[kLinThreads / 32][1];
  const int begin = __ldg(seg_start + blockIdx.y), end = __ldg(seg_start + blockIdx.y + 1);
  double acc = 0.0;
  for (int i = begin + blockIdx.x * kLinThreads + threadIdx.x; i < end; i += gridDim.x * kLinThreads) acc += __ldg(in + i);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][0] = acc;
  block_publish<1>(wsum, partials, counters, slots, seq);
}

__device__ __forceinline__ PoseArg load_pose(const PoseArg& p0, const PoseArg* __restrict__ poses) {
  if (!poses) return p0;
  return poses[blockIdx.y];
}

// K4a. Correspondence search (nano_gicp.cc:219-227): fp32 transform of every source point, exact 1-NN in the target
// index, strict distance gate. Two kernels (bnn.cuh): the fast one resolves every query whose bounded ball is small
// and lists the rest; the heavy one gives each listed query a whole warp. grid.y = scans; source scan b = source
// segment b, its target segment is target_seg[b] (or 0).
// corr[j] = target sorted position, or -2-pos for a neighbour that failed the gate / a bound that is not yet the
// answer (both are hints for the next search of this point), or -1.
__device__ __forceinline__ void transform_query(const PoseArg& P, const float4& pa, float qf[3]) {
  // ((r0*x + r1*y) + r2*z) + t*w with w = 1, fp32 (nano_gicp.cc:222)
#pragma unroll
  for (int r = 0; r < 3; r++)
    qf[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.Rf[3 * r], pa.x), __fmul_rn(P.Rf[3 * r + 1], pa.y)), __fmul_rn(P.Rf[3 * r + 2], pa.z)), P.tf[r]);
}
__device__ __forceinline__ int corr_code(int pos, float d, double thr2) {
  const bool valid = pos >= 0 && (double)d < thr2;                     // strict, float promoted to double (nano_gicp.cc:227)
  return valid ? pos : (pos >= 0 ? -2 - pos : -1);
}
__device__ __forceinline__ int corr_hint(int c) { return c >= 0 ? c : (c <= -2 ? -2 - c : -1); }

// a query the fast kernel deferred: transformed point, bound (squared distance, original index, sorted position; the
// clip to max_sqd has i = pos = -1), its place in the correspondence array and its target segment
struct __align__(16) HeavyQuery {
  float4 q;   // x, y, z, bound d
  int4 m;     // bound i, bound pos, j, target segment
};

template <bool USE_PREV>
__global__ void __launch_bounds__(kLinThreads) correspond_fast_kernel(GridView src, GridView tgt, PoseArg pose0, const PoseArg* __restrict__ poses,
                                                                       const int* __restrict__ target_seg, double thr2, float max_sqd,
                                                                       const int* corr_in, int* corr, HeavyQuery* __restrict__ heavy,
                                                                       unsigned int* __restrict__ heavy_count) {
  const PoseArg P = load_pose(pose0, poses);
  const int b = blockIdx.y;
  const int begin = src.n_seg > 1 ? __ldg(src.seg_start + b) : 0;
  const int end = src.n_seg > 1 ? __ldg(src.seg_start + b + 1) : src.n;
  const int tseg = target_seg ? __ldg(target_seg + b) : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LPQ = 4, kPer = 32 / LPQ;
  const int gwarp = blockIdx.x * (kLinThreads / 32) + warp, nwarps = gridDim.x * (kLinThreads / 32);
  for (int j0 = begin + gwarp * kPer; j0 < end; j0 += nwarps * kPer) {
    const int j = j0 + lane / LPQ;
    const bool active = j < end;
    bool is_heavy = false;
    float qf[3] = {0.f, 0.f, 0.f};
    NNBest best;
    best.d = 0.f; best.i = -1; best.pos = -1;
    if (active) {
      const float4 pa = __ldg(src.pts + j);
      transform_query(P, pa, qf);
      bool have = false;
      if (USE_PREV) {
        const int prev = corr_hint(corr_in[j]);   // hints may live in another buffer than the results (speculative search)
        if (prev >= 0) {
          const float4 pb0 = __ldg(tgt.pts + prev);
          have = true;
          best.d = sqdist_ref(qf[0], qf[1], qf[2], pb0.x, pb0.y, pb0.z);
          best.i = __float_as_int(pb0.w);
          best.pos = prev;
        }
      }
      const bool resolved = bnn_group(tgt, qf[0], qf[1], qf[2], tseg, have, max_sqd, best);
      is_heavy = !resolved && (lane & (LPQ - 1)) == 0;
      if (resolved && (lane & (LPQ - 1)) == 0) corr[j] = corr_code(best.pos, best.d, thr2);
    }
    // one atomic per warp reserves list slots for its heavy queries
    const unsigned hm = __ballot_sync(0xffffffffu, is_heavy);
    if (hm) {
      unsigned int slot = 0;
      if (lane == __ffs(hm) - 1) slot = atomicAdd(heavy_count, (unsigned int)__popc(hm));
      slot = __shfl_sync(0xffffffffu, slot, __ffs(hm) - 1);
      if (is_heavy) {
        HeavyQuery* q = heavy + slot + __popc(hm & ((1u << lane) - 1u));
        q->q = make_float4(qf[0], qf[1], qf[2], best.d);
        q->m = make_int4(best.i, best.pos, j, tseg);
      }
    }
  }
}

__global__ void __launch_bounds__(kLinThreads, 5) correspond_heavy_kernel(GridView tgt, double thr2, float max_sqd, int cmax, int* __restrict__ corr,
                                                                           const HeavyQuery* __restrict__ heavy, const unsigned int* __restrict__ heavy_count,
                                                                           unsigned int* __restrict__ next_count) {
  __shared__ WarpScratch scratch[kLinThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int count = __ldg(heavy_count);
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0;      // the list of the next search starts empty
  const int gwarp = blockIdx.x * (kLinThreads / 32) + warp, nwarps = gridDim.x * (kLinThreads / 32);
  if ((unsigned int)gwarp >= count) return;
  uint32_t phase = wknn_init(scratch[warp]);
  for (unsigned int e = gwarp; e < count; e += nwarps) {
    const float4 q = __ldg(&heavy[e].q);
    const int4 m = __ldg(&heavy[e].m);
    NNBest best;
    best.d = q.w; best.i = m.x; best.pos = m.y;
    const int j = m.z, tseg = m.w;
    if (!bnn_warp(tgt, q.x, q.y, q.z, tseg, best, scratch[warp], phase)) {
      // ball larger than the bounded search handles: general search, lanes 0..3 carry the query
      TopK<1> top;
      warp_knn<4>(tgt, lane < 4, q.x, q.y, q.z, tseg, 1, cmax, max_sqd, top, scratch[warp], phase);
      best.d = __shfl_sync(0xffffffffu, top.d[0], 0);
      best.i = __shfl_sync(0xffffffffu, top.p[0], 0);
      best.pos = best.i >= 0 ? __ldg(tgt.inv + best.i) : -1;
    }
    if (lane == 0) corr[j] = corr_code(best.pos, best.d, thr2);
  }
}

// K4b. Fused linearisation (nano_gicp.cc:237-241,259-299): per matched source point the Mahalanobis matrix
// (C_B + R C_A R^T)^-1, the residual, the 6-DoF Jacobian and the 21 + 6 + 1 (+ count) contributions, accumulated in
// fp64 registers, reduced by warp shuffles, per block, and by the last block to finish (fixed order, compensated).
// A streaming kernel: 16 B p_A + 24 B C_A + 4 B corr + gathers of 16 B p_B + 24 B C_B per point, nothing written per point.
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// PF: also pull the NEXT trip's gather targets (p_B, C_B; correspondence fetched two trips ahead) into L1
template <bool WANT_HB, bool PF = false>
__global__ void __launch_bounds__(kLinThreads) linearize_kernel(GridView src, GridView tgt, const float* __restrict__ cov_src, const float* __restrict__ cov_tgt,
                                                                 PoseArg pose0, const PoseArg* __restrict__ poses, const int* __restrict__ corr,
                                                                 double* __restrict__ partials, unsigned int* __restrict__ counters,
                                                                 ReduceSlot* __restrict__ slots, unsigned long long seq) {
  const PoseArg P = load_pose(pose0, poses);
  const int b = blockIdx.y;
  const int begin = src.n_seg > 1 ? __ldg(src.seg_start + b) : 0;
  const int end = src.n_seg > 1 ? __ldg(src.seg_start + b + 1) : src.n;
  __shared__ double wsum[kLinThreads / 32][kTerms];
  double acc[kTerms];
#pragma unroll
  for (int t = 0; t < kTerms; t++) acc[t] = 0.0;
  // the correspondence of the NEXT trip is fetched one trip ahead, so the dependent gathers (p_B, C_B) leave together
  // with the streaming loads (p_A, C_A) instead of one memory latency behind them
  const int stride = gridDim.x * kLinThreads;
  int j = begin + blockIdx.x * kLinThreads + threadIdx.x;
  int pos_next = j < end ? __ldg(corr + j) : -1;
  int pos_next2 = (PF && j + stride < end) ? __ldg(corr + j + stride) : -1;
  for (; j < end; j += stride) {
    const int pos = pos_next;
    if (PF) {
      pos_next = pos_next2;
      pos_next2 = j + 2 * stride < end ? __ldg(corr + j + 2 * stride) : -1;
      if (pos_next >= 0) {
        prefetch_l1(tgt.pts + pos_next);
        prefetch_l1(cov_tgt + (size_t)pos_next * 6);
        prefetch_l1(cov_tgt + (size_t)pos_next * 6 + 5);
      }
    } else {
      pos_next = j + stride < end ? __ldg(corr + j + stride) : -1;
    }
    if (pos < 0) continue;
    const float4 pa = __ldg(src.pts + j);
    const float4 pb = __ldg(tgt.pts + pos);
    float ca[6], cb[6];
    {
      const float2* a2 = reinterpret_cast<const float2*>(cov_src + (size_t)j * 6);
      const float2* b2 = reinterpret_cast<const float2*>(cov_tgt + (size_t)pos * 6);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const float2 va = __ldg(a2 + i), vb = __ldg(b2 + i);
        ca[2 * i] = va.x; ca[2 * i + 1] = va.y; cb[2 * i] = vb.x; cb[2 * i + 1] = vb.y;
      }
    }
    double M[6], q[3], e[3], Me[3];
    mahalanobis(P, ca, cb, M);
    residual(P, pa, pb, q, e);
    acc[27] += quad_form(M, e, Me);
    acc[28] += (__float_as_int(pb.w) > 0) ? 1.0 : 0.0;  // num_correspondences counts indices > 0 (nano_gicp.cc:244)
    if (WANT_HB) {
      // G = S M (S = skew(q)): column k = q x M[:,k];  H_rr row i = q x G[i,:];  H_rt = G;  H_tt = M
      double g0[3], g1[3], g2[3];
      cross3(q, M[0], M[1], M[2], g0);
      cross3(q, M[1], M[3], M[4], g1);
      cross3(q, M[2], M[4], M[5], g2);
      double h0[3], h1[3], h2[3];   // rows of G: G[i][k] = gk[i]
      cross3(q, g0[0], g1[0], g2[0], h0);
      cross3(q, g0[1], g1[1], g2[1], h1);
      cross3(q, g0[2], g1[2], g2[2], h2);
      // upper triangle, row-major: (0,0..5) (1,1..5) (2,2..5) (3,3..5) (4,4..5) (5,5)
      acc[0] += h0[0]; acc[1] += h0[1]; acc[2] += h0[2]; acc[3] += g0[0]; acc[4] += g1[0]; acc[5] += g2[0];
      acc[6] += h1[1]; acc[7] += h1[2]; acc[8] += g0[1]; acc[9] += g1[1]; acc[10] += g2[1];
      acc[11] += h2[2]; acc[12] += g0[2]; acc[13] += g1[2]; acc[14] += g2[2];
      acc[15] += M[0]; acc[16] += M[1]; acc[17] += M[2];
      acc[18] += M[3]; acc[19] += M[4];
      acc[20] += M[5];
      double qm[3];   // b = J^T M e = [ -q x Me ; -Me ]
      cross3(q, Me[0], Me[1], Me[2], qm);
      acc[21] -= qm[0]; acc[22] -= qm[1]; acc[23] -= qm[2];
      acc[24] -= Me[0]; acc[25] -= Me[1]; acc[26] -= Me[2];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < kTerms; t++) {
    double v = acc[t];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) wsum[warp][t] = v;
  }
  block_publish<kTerms>(wsum, partials, counters, slots, seq);
}

// K5. Trial pose P, Mahalanobis rebuilt at the linearisation pose P0.
__global__ void __launch_bounds__(kLinThreads) error_kernel(GridView src, GridView tgt, const float* __restrict__ cov_src, const float* __restrict__ cov_tgt,
                                                            PoseArg pose0, const PoseArg* __restrict__ poses, PoseArg lin0, const PoseArg* __restrict__ lins,
                                                            const int* __restrict__ corr, double* __restrict__ partials, unsigned int* __restrict__ counters,
                                                            ReduceSlot* __restrict__ slots, unsigned long long seq) {
  const PoseArg P = load_pose(pose0, poses);
  const PoseArg P0 = load_pose(lin0, lins);
  const int b = blockIdx.y;
  const int begin = src.n_seg > 1 ? __ldg(src.seg_start + b) : 0;
  const int end = src.n_seg > 1 ? __ldg(src.seg_start + b + 1) : src.n;
  __shared__ double wsum[kLinThreads / 32][1];
  double acc[1] = {0.0};
  const int stride = gridDim.x * kLinThreads;
  int j = begin + blockIdx.x * kLinThreads + threadIdx.x;
  int pos_next = j < end ? __ldg(corr + j) : -1;   // one trip ahead, as in K4b
  for (; j < end; j += stride) {
    const int pos = pos_next;
    pos_next = j + stride < end ? __ldg(corr + j + stride) : -1;
    if (pos < 0) continue;
    const float4 pa = __ldg(src.pts + j);
    const float4 pb = __ldg(tgt.pts + pos);
    float ca[6], cb[6];
    const float2* a2 = reinterpret_cast<const float2*>(cov_src + (size_t)j * 6);
    const float2* b2 = reinterpret_cast<const float2*>(cov_tgt + (size_t)pos * 6);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float2 va = __ldg(a2 + i), vb = __ldg(b2 + i);
      ca[2 * i] = va.x; ca[2 * i + 1] = va.y; cb[2 * i] = vb.x; cb[2 * i + 1] = vb.y;
    }
    double M[6], q[3], e[3], Me[3];
    mahalanobis(P0, ca, cb, M);
    residual(P, pa, pb, q, e);
    acc[0] += quad_form(M, e, Me);
  }
  {
    double v = acc[0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][0] = v;
  }
  block_publish<1>(wsum, partials, counters, slots, seq);
}

// API-parity export of update_correspondences in ORIGINAL source order (slow path, tests only)
__global__ void __launch_bounds__(128) export_corr_kernel(GridView src, GridView tgt, const float* __restrict__ cov_src, const float* __restrict__ cov_tgt,
                                                          PoseArg P, const int* __restrict__ corr_sorted, int* __restrict__ corr_out,
                                                          float* __restrict__ sqd_out, double* __restrict__ mahal_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= src.n) return;
  const float4 pa = __ldg(src.pts + j);
  const int orig = __float_as_int(pa.w);
  const int pos = corr_sorted[j];
  if (corr_out) corr_out[orig] = pos >= 0 ? __float_as_int(__ldg(&tgt.pts[pos].w)) : -1;
  float d = __int_as_float(0x7f800000);
  double M4[16];
#pragma unroll
  for (int i = 0; i < 16; i++) M4[i] = 0.0;
  if (pos >= 0) {
    const float4 pb = __ldg(tgt.pts + pos);
    float qf[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      qf[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.Rf[3 * r], pa.x), __fmul_rn(P.Rf[3 * r + 1], pa.y)), __fmul_rn(P.Rf[3 * r + 2], pa.z)), P.tf[r]);
    d = sqdist_ref(qf[0], qf[1], qf[2], pb.x, pb.y, pb.z);
    float ca[6], cb[6];
    for (int i = 0; i < 6; i++) { ca[i] = cov_src[(size_t)j * 6 + i]; cb[i] = cov_tgt[(size_t)pos * 6 + i]; }
    double M[6];
    mahalanobis(P, ca, cb, M);
    M4[0] = M[0]; M4[1] = M[1]; M4[2] = M[2];
    M4[4] = M[1]; M4[5] = M[3]; M4[6] = M[4];
    M4[8] = M[2]; M4[9] = M[4]; M4[10] = M[5];
  }
  if (sqd_out) sqd_out[orig] = d;
  if (mahal_out)
    for (int i = 0; i < 16; i++) mahal_out[(size_t)orig * 16 + i] = M4[i];
}

__global__ void __launch_bounds__(256) transform_points_kernel(const float* __restrict__ in, int stride, int n, PoseArg P, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[(size_t)i * stride], y = in[(size_t)i * stride + 1], z = in[(size_t)i * stride + 2];
#pragma unroll
  for (int r = 0; r < 3; r++)
    out[(size_t)i * 3 + r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.Rf[3 * r], x), __fmul_rn(P.Rf[3 * r + 1], y)), __fmul_rn(P.Rf[3 * r + 2], z)), P.tf[r]);
}

}  // namespace

PoseArg make_pose(const double T[16]) {
  PoseArg P;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) { P.R[3 * r + c] = T[4 * c + r]; P.Rf[3 * r + c] = (float)T[4 * c + r]; }
    P.t[r] = T[12 + r];
    P.tf[r] = (float)T[12 + r];
  }
  return P;
}

static float max_sqd_for(double thr) {
  // smallest float that is certainly >= thr^2 (so that "covered^2 >= max_sqd" implies nothing unvisited can pass the gate)
  const double t2 = thr * thr;
  if (!(t2 < 3.0e38)) return __builtin_inff();
  float f = (float)t2;
  f = __builtin_nextafterf(f, __builtin_inff());
  return __builtin_nextafterf(f, __builtin_inff());
}

static int g_lin_block_cap = -1;
int lin_blocks_for(int n) {
  if (g_lin_block_cap < 0) { const char* e = std::getenv("NGICP_LIN_BLOCKS"); g_lin_block_cap = e ? std::atoi(e) : 0; }
  if (g_lin_block_cap > 0) return std::max(1, std::min((n + kLinThreads - 1) / kLinThreads, g_lin_block_cap));
  // two blocks per SM: enough threads in flight for the gathers, few enough partials for a short reduction tail
  const int want = (n + kLinThreads - 1) / kLinThreads;
  return std::max(1, std::min(want, 148 * 2));
}

int wait_slot(Handle* h, int n_slots, unsigned long long seq) {
  // spin on the host-mapped sequence numbers; poll the stream every so often so that a failed
  // launch cannot hang the caller
  for (int b = 0; b < n_slots; b++) {
    volatile unsigned long long* p = &h->slot_host[b].seq;
    unsigned long long spins = 0;
    while (*p != seq) {
      if ((++spins & 0x3fff) != 0) continue;
      const cudaError_t q = cudaStreamQuery(h->stream);
      if (q == cudaErrorNotReady) continue;
      if (q != cudaSuccess) return fail(h, NGICP_ERR_CUDA, std::string("reduction kernel: ") + cudaGetErrorString(q));
      if (*p != seq) return fail(h, NGICP_ERR_CUDA, "reduction kernel finished without writing its result slot");
    }
  }
  __sync_synchronize();
  return NGICP_OK;
}

int reduce_sum(Handle* h, const double* d_in, int n, const int* seg_start_dev, int n_seg, double* host_out) {
  for (int s0 = 0; s0 < n_seg; s0 += kMaxBatch) {
    const int ns = std::min(kMaxBatch, n_seg - s0);
    const int per_seg = std::max(1, n / std::max(n_seg, 1));
    const int blocks = std::max(1, std::min(std::min(64, kMaxLinBlocks / ns), (per_seg + kLinThreads * 8 - 1) / (kLinThreads * 8)));
    const unsigned long long seq = ++h->seq;
    segment_sum_kernel<<<dim3(blocks, ns), kLinThreads, 0, h->stream>>>(d_in, seg_start_dev + s0, h->partials, h->counter, h->slot_dev, seq);
    count_launch(h);
    NGICP_CUDA(h, cudaGetLastError());
    if (int rc = wait_slot(h, ns, seq)) return rc;
    for (int s = 0; s < ns; s++) host_out[s0 + s] = h->slot_host[s].v[0];
  }
  return NGICP_OK;
}

static int check_ready(Handle* h) {
  for (int w = 0; w < 2; w++) {
    if (!h->index[w]) return fail(h, NGICP_ERR_INVALID, w == 0 ? "no source cloud" : "no target cloud");
    if (!h->covs[w].valid || h->covs[w].n != (size_t)h->index[w]->n)
      return fail(h, NGICP_ERR_INVALID, w == 0 ? "source covariances missing" : "target covariances missing");
  }
  return NGICP_OK;
}

static int ensure_corr(Handle* h, size_t n) {
  if (h->corr_cap >= n) return NGICP_OK;
  // cudaFree / cudaMalloc synchronise the device (and were seen to take tens to hundreds of ms in a process whose memory
  // pool is warm): grow rarely — the voxel-filtered scans of a sequence differ by a few percent from one to the next
  n = ((n + n / 4 + 4095) / 4096) * 4096;
  drop_speculation(h);
  if (h->corr) NGICP_CUDA(h, cudaFree(h->corr));
  if (h->corr_alt) NGICP_CUDA(h, cudaFree(h->corr_alt));
  if (h->heavy) NGICP_CUDA(h, cudaFree(h->heavy));
  h->corr = nullptr; h->corr_alt = nullptr; h->heavy = nullptr; h->corr_cap = 0;
  NGICP_CUDA(h, cudaMalloc(&h->corr, sizeof(int) * n));
  NGICP_CUDA(h, cudaMalloc(&h->corr_alt, sizeof(int) * n));
  NGICP_CUDA(h, cudaMalloc(&h->heavy, sizeof(HeavyQuery) * n));
  if (!h->heavy_count) {
    NGICP_CUDA(h, cudaMalloc(&h->heavy_count, sizeof(unsigned int) * 2));
    NGICP_CUDA(h, cudaMemsetAsync(h->heavy_count, 0, sizeof(unsigned int) * 2, h->stream));
  }
  h->corr_cap = n;
  return NGICP_OK;
}

// search grid: persistent warps, enough of them to fill the machine several times over
static int search_blocks_for(int n, int lpq) {
  const int items = (n * lpq + 31) / 32;                         // one warp-sized work item per 32/lpq points
  const int want = (items + (kLinThreads / 32) - 1) / (kLinThreads / 32);
  // exactly the blocks that are resident at once (the warps stride over the queries): a partial second wave of blocks
  // would run its whole share on a nearly empty machine
  static const int per_sm = [] {
    if (const char* e = std::getenv("NGICP_K4_BLOCKS_PER_SM")) return std::max(1, std::atoi(e));
    int a = 0, b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, correspond_fast_kernel<true>, kLinThreads, 0) != cudaSuccess) a = 5;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, correspond_fast_kernel<false>, kLinThreads, 0) != cudaSuccess) b = 5;
    return std::max(1, std::min(a, b));
  }();
  return std::max(1, std::min(want, 148 * per_sm));
}

// K4 = correspondence search + fused linearisation over n_scans source segments (n_scans = 1: the reference's
// linearize). Poses, target segments and per-scan results live in device / host-mapped arrays when batched.
// the two search kernels of K4a on `stream`: hints from corr_in (if use_prev), results to corr_out
static void launch_search(Handle* h, cudaStream_t stream, int n_scans, const PoseArg& P0, const PoseArg* d_poses, const int* d_target_seg, bool use_prev,
                          const int* corr_in, int* corr_out) {
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  const double thr = h->params.max_corr_dist;
  const double thr2 = thr * thr;
  const int per_scan = si->n / n_scans + 1;
  const dim3 sgrid(search_blocks_for(per_scan, 4), n_scans);
  const int hgrid = 148 * 5;   // heavy queries: one warp each, persistent over the list
  unsigned int* cnt = h->heavy_count + (h->heavy_parity & 1u);
  unsigned int* cnt_next = h->heavy_count + ((h->heavy_parity + 1u) & 1u);
  h->heavy_parity++;
  if (use_prev)
    correspond_fast_kernel<true><<<sgrid, kLinThreads, 0, stream>>>(si->view(), ti->view(), P0, d_poses, d_target_seg, thr2, max_sqd_for(thr), corr_in, corr_out,
                                                                    static_cast<HeavyQuery*>(h->heavy), cnt);
  else
    correspond_fast_kernel<false><<<sgrid, kLinThreads, 0, stream>>>(si->view(), ti->view(), P0, d_poses, d_target_seg, thr2, max_sqd_for(thr), corr_in, corr_out,
                                                                     static_cast<HeavyQuery*>(h->heavy), cnt);
  correspond_heavy_kernel<<<hgrid, kLinThreads, 0, stream>>>(ti->view(), thr2, max_sqd_for(thr), h->k4_cmax, corr_out, static_cast<const HeavyQuery*>(h->heavy), cnt,
                                                             cnt_next);
  count_launch(h, 2);
}

// K4 = correspondence search + fused linearisation over n_scans source segments (n_scans = 1: the reference's
// linearize). Poses, target segments and per-scan results live in device / host-mapped arrays when batched.
// search_done: h->corr already holds the correspondences for this pose (speculative search, see speculate_search).
static int launch_linearize(Handle* h, int n_scans, const PoseArg& P0, const PoseArg* d_poses, const int* d_target_seg, bool want_Hb,
                            double* partials, unsigned long long seq, bool search_done = false) {
  const Index* si = h->index[0];
  const int per_scan = si->n / n_scans + 1;
  // batched: few fat blocks per scan (many points per thread amortise the 29-term block reduction); single scan: wide
  static const int lin_mult = getenv("NGICP_K4B_MULT") ? atoi(getenv("NGICP_K4B_MULT")) : 8;   // development switches
  static const bool lin_pf = getenv("NGICP_K4B_PF") && atoi(getenv("NGICP_K4B_PF"));
  const dim3 lgrid(std::max(1, std::min(lin_blocks_for(per_scan), (148 * lin_mult) / n_scans)), n_scans);
  // the correspondences of the previous linearize (same clouds) seed this one
  const bool use_prev = h->lin_valid && h->k4_ball && h->corr_n == (size_t)si->n;
  if (h->timing) cudaEventRecord(h->ev[0], h->stream);
  if (!search_done) launch_search(h, h->stream, n_scans, P0, d_poses, d_target_seg, use_prev, h->corr, h->corr);
  if (h->timing) cudaEventRecord(h->ev[2], h->stream);
  const Index* ti = h->index[1];
  if (want_Hb && lin_pf)
    linearize_kernel<true, true><<<lgrid, kLinThreads, 0, h->stream>>>(si->view(), ti->view(), h->covs[0].cov6, h->covs[1].cov6, P0, d_poses, h->corr, partials,
                                                                      h->counter, h->slot_dev, seq);
  else if (want_Hb)
    linearize_kernel<true><<<lgrid, kLinThreads, 0, h->stream>>>(si->view(), ti->view(), h->covs[0].cov6, h->covs[1].cov6, P0, d_poses, h->corr, partials,
                                                                h->counter, h->slot_dev, seq);
  else
    linearize_kernel<false><<<lgrid, kLinThreads, 0, h->stream>>>(si->view(), ti->view(), h->covs[0].cov6, h->covs[1].cov6, P0, d_poses, h->corr, partials,
                                                                 h->counter, h->slot_dev, seq);
  h->corr_n = (size_t)si->n;
  count_launch(h, 1);
  NGICP_CUDA(h, cudaGetLastError());
  if (h->timing) cudaEventRecord(h->ev[1], h->stream);
  if (int rc = wait_slot(h, n_scans, seq)) return rc;
  if (h->timing) {
    cudaEventSynchronize(h->ev[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[2]);
    h->t.correspond_ms += ms;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[1]);
    h->t.linearize_ms += ms;
  }
  h->t.linearize_calls++;
  return NGICP_OK;
}

// the second stream outranks the main one: a speculative search shares the GPU with K2 (8,192 blocks) or K5 and must get
// its blocks resident as soon as slots free up, or it would simply queue behind them
static cudaError_t create_search_stream(Handle* h) {
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  return cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, greatest);
}

// While K5 evaluates a trial pose on the main stream, search the correspondences AT that pose on a second stream: if
// the LM step is accepted (the usual case) the next linearize starts with its search already done. Hints come from the
// current correspondences (read-only here and in K5), results go to the alternate buffer; nothing is consumed unless
// the next linearize asks for exactly this pose.
int speculate_search(Handle* h, const double T[16]) {
  const Index* si = h->index[0];
  if (!h->k4_spec || h->timing || !h->lin_valid || !h->corr_alt || h->corr_n != (size_t)si->n || si->n_seg != 1) return NGICP_OK;
  if (!h->stream2) {
    NGICP_CUDA(h, create_search_stream(h));
    NGICP_CUDA(h, cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    NGICP_CUDA(h, cudaEventCreateWithFlags(&h->ev_spec, cudaEventDisableTiming));
  }
  NGICP_CUDA(h, cudaEventRecord(h->ev_main, h->stream));          // everything the search reads is ordered on the main stream
  NGICP_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_main, 0));
  const PoseArg P = make_pose(T);
  launch_search(h, h->stream2, 1, P, nullptr, nullptr, h->k4_ball != 0, h->corr, h->corr_alt);
  NGICP_CUDA(h, cudaGetLastError());
  NGICP_CUDA(h, cudaEventRecord(h->ev_spec, h->stream2));
  std::memcpy(h->spec_T, T, sizeof h->spec_T);
  h->spec_pending = true;
  return NGICP_OK;
}

// The FIRST search of the coming align does not need the source covariances: started from calculate*Covariances for the
// source (api.cu:compute_covariances_impl) on the second stream, it runs beside K2 + K3. The pose is the identity — DLIO
// hands GICP scans that are already in the world frame and aligns without a guess (reference src/dlio/src/dlio/odom.cc:1005);
// any other first pose, a changed target / source / parameter set in between simply discards the result.
int speculate_first_search(Handle* h) {
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  if (!h->k4_spec || h->timing || !si || !ti || si->n_seg != 1 || ti->n_seg != 1) return NGICP_OK;
  if (int rc = ensure_corr(h, si->n)) return rc;
  drop_speculation(h);
  if (!h->stream2) {
    NGICP_CUDA(h, create_search_stream(h));
    NGICP_CUDA(h, cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming));
    NGICP_CUDA(h, cudaEventCreateWithFlags(&h->ev_spec, cudaEventDisableTiming));
  }
  NGICP_CUDA(h, cudaEventRecord(h->ev_main, h->stream));          // the source index build is ordered on the main stream
  NGICP_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_main, 0));
  double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const PoseArg P = make_pose(T);
  launch_search(h, h->stream2, 1, P, nullptr, nullptr, false, h->corr, h->corr_alt);
  NGICP_CUDA(h, cudaGetLastError());
  NGICP_CUDA(h, cudaEventRecord(h->ev_spec, h->stream2));
  std::memcpy(h->spec_T, T, sizeof h->spec_T);
  h->spec_pending = true;
  h->spec_first = true;
  return NGICP_OK;
}

// orders the main stream after any speculative search still in flight (before the index, the list or the buffers it
// uses are touched again) and forgets it
void drop_speculation(Handle* h) {
  if (h->spec_pending) cudaStreamWaitEvent(h->stream, h->ev_spec, 0);
  h->spec_pending = false;
  h->spec_first = false;
}

static void unpack_result(const double* v, bool want_Hb, double H[36], double b[6], double* err, int* ncorr) {
  if (H && b && want_Hb) {
    int t = 0;
    for (int r = 0; r < 6; r++)
      for (int c = r; c < 6; c++) { H[6 * r + c] = v[t]; H[6 * c + r] = v[t]; t++; }
    for (int i = 0; i < 6; i++) b[i] = v[21 + i];
  }
  if (err) *err = v[27];
  if (ncorr) *ncorr = (int)(v[28] + 0.5);
}

int linearize_device(Handle* h, const double T[16], bool want_Hb, double H[36], double b[6], double* err, int* ncorr) {
  if (int rc = check_ready(h)) return rc;
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  if (si->n_seg != 1 || ti->n_seg != 1) return fail(h, NGICP_ERR_UNSUPPORTED, "linearize: use ngicp_batch_linearize for multi-segment clouds");
  if (int rc = ensure_corr(h, si->n)) return rc;
  const PoseArg P = make_pose(T);
  const unsigned long long seq = ++h->seq;
  bool search_done = false;
  if (h->spec_pending) {
    search_done = (h->spec_first || (h->lin_valid && h->corr_n == (size_t)si->n)) && std::memcmp(h->spec_T, T, sizeof h->spec_T) == 0;
    drop_speculation(h);                                   // the main stream now follows the speculative search
    if (search_done) std::swap(h->corr, h->corr_alt);
  }
  if (int rc = launch_linearize(h, 1, P, nullptr, nullptr, want_Hb, h->partials, seq, search_done)) return rc;
  int nc = 0;
  unpack_result(h->slot_host[0].v, want_Hb, H, b, err, &nc);
  h->num_correspondences = nc;
  if (ncorr) *ncorr = nc;
  for (int i = 0; i < 9; i++) h->lin_pose[i] = P.R[i];
  for (int i = 0; i < 3; i++) h->lin_pose[9 + i] = P.t[i];
  h->lin_valid = true;
  return NGICP_OK;
}

// Batched linearize: source = n_scans segments (ngicp_set_input_batch), every scan with its own pose against the
// (single-segment) target. One search launch and one linearisation launch for all scans; per-scan results.
int batch_linearize_device(Handle* h, int n_scans, const double* T16s, double* H36s, double* b6s, double* errs, int* ncorrs) {
  if (int rc = check_ready(h)) return rc;
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  if (si->n_seg != n_scans) return fail(h, NGICP_ERR_INVALID, "batch linearize: the source must hold one segment per scan");
  if (ti->n_seg != 1) return fail(h, NGICP_ERR_UNSUPPORTED, "batch linearize: the target must be a single cloud");
  if (n_scans < 1 || n_scans > kMaxBatch) return fail(h, NGICP_ERR_INVALID, "batch linearize: 1..256 scans");
  if (int rc = ensure_corr(h, si->n)) return rc;
  const int per_scan = si->n / n_scans + 1;
  const size_t need = (size_t)n_scans * lin_blocks_for(per_scan) * 32;   // upper bound on blocks per scan
  if (h->batch_partials_cap < need) {
    if (h->batch_partials) NGICP_CUDA(h, cudaFree(h->batch_partials));
    h->batch_partials = nullptr; h->batch_partials_cap = 0;
    NGICP_CUDA(h, cudaMalloc(&h->batch_partials, sizeof(double) * need));
    h->batch_partials_cap = need;
  }
  std::vector<PoseArg> poses(n_scans);
  for (int s = 0; s < n_scans; s++) poses[s] = make_pose(T16s + 16 * (size_t)s);
  PoseArg* d_poses = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_poses, (size_t)n_scans, h->stream));
  NGICP_CUDA(h, cudaMemcpyAsync(d_poses, poses.data(), sizeof(PoseArg) * n_scans, cudaMemcpyHostToDevice, h->stream));
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));   // poses is a stack-lifetime pageable buffer
  const unsigned long long seq = ++h->seq;
  const int rc = launch_linearize(h, n_scans, poses[0], d_poses, nullptr, true, h->batch_partials, seq);
  dev_free(d_poses, h->stream);
  if (rc) return rc;
  for (int s = 0; s < n_scans; s++)
    unpack_result(h->slot_host[s].v, true, H36s ? H36s + 36 * (size_t)s : nullptr, b6s ? b6s + 6 * (size_t)s : nullptr, errs ? errs + s : nullptr,
                  ncorrs ? ncorrs + s : nullptr);
  h->lin_valid = false;   // the single-scan LM state does not apply to a batch
  return NGICP_OK;
}

int compute_error_device(Handle* h, const double T[16], double* err) {
  if (int rc = check_ready(h)) return rc;
  if (!h->lin_valid) return fail(h, NGICP_ERR_INVALID, "compute_error before linearize (no cached correspondences)");
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  const PoseArg P = make_pose(T);
  PoseArg P0;
  for (int i = 0; i < 9; i++) { P0.R[i] = h->lin_pose[i]; P0.Rf[i] = (float)h->lin_pose[i]; }
  for (int i = 0; i < 3; i++) { P0.t[i] = h->lin_pose[9 + i]; P0.tf[i] = (float)h->lin_pose[9 + i]; }
  const dim3 grid(lin_blocks_for(si->n), 1);
  const unsigned long long seq = ++h->seq;
  if (h->timing) cudaEventRecord(h->ev[0], h->stream);
  error_kernel<<<grid, kLinThreads, 0, h->stream>>>(si->view(), ti->view(), h->covs[0].cov6, h->covs[1].cov6, P, nullptr, P0, nullptr, h->corr, h->partials,
                                                   h->counter, h->slot_dev, seq);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  if (h->timing) cudaEventRecord(h->ev[1], h->stream);
  if (int rc = wait_slot(h, 1, seq)) return rc;
  if (h->timing) {
    cudaEventSynchronize(h->ev[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->t.error_ms += ms;
  }
  h->t.error_calls++;
  if (err) *err = h->slot_host[0].v[0];
  return NGICP_OK;
}

int export_correspondences(Handle* h, const double T[16], int32_t* corr, float* sqd, double* mahal, int* ncorr) {
  double e;
  if (int rc = linearize_device(h, T, false, nullptr, nullptr, &e, ncorr)) return rc;
  const Index* si = h->index[0];
  const Index* ti = h->index[1];
  const int n = si->n;
  cudaStream_t s = h->stream;
  int* d_corr = nullptr; float* d_sqd = nullptr; double* d_m = nullptr;
  if (corr) NGICP_CUDA(h, dev_alloc(&d_corr, (size_t)n, s));
  if (sqd) NGICP_CUDA(h, dev_alloc(&d_sqd, (size_t)n, s));
  if (mahal) NGICP_CUDA(h, dev_alloc(&d_m, (size_t)n * 16, s));
  export_corr_kernel<<<(n + 127) / 128, 128, 0, s>>>(si->view(), ti->view(), h->covs[0].cov6, h->covs[1].cov6, make_pose(T), h->corr, d_corr, d_sqd, d_m);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  if (corr) NGICP_CUDA(h, cudaMemcpyAsync(corr, d_corr, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  if (sqd) NGICP_CUDA(h, cudaMemcpyAsync(sqd, d_sqd, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  if (mahal) NGICP_CUDA(h, cudaMemcpyAsync(mahal, d_m, sizeof(double) * 16 * n, cudaMemcpyDeviceToHost, s));
  NGICP_CUDA(h, cudaStreamSynchronize(s));
  dev_free(d_corr, s); dev_free(d_sqd, s); dev_free(d_m, s);
  return NGICP_OK;
}

int transform_points_device(Handle* h, const float* d_xyz_in, int stride_floats, int n, const float T[16], float* d_xyz_out) {
  double Td[16];
  for (int i = 0; i < 16; i++) Td[i] = T[i];
  PoseArg P = make_pose(Td);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) P.Rf[3 * r + c] = T[4 * c + r]; P.tf[r] = T[12 + r]; }
  transform_points_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(d_xyz_in, stride_floats, n, P, d_xyz_out);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

}  // namespace ngicp

#ifdef NGICP_STATS
extern "C" int ngicp_debug_stats_lin(unsigned long long out[8], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_wknn_stats, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ngicp::g_wknn_stats, z, sizeof z); }
  return 0;
}
#endif

#ifdef NGICP_STATS
extern "C" int ngicp_debug_bnn(unsigned long long out[16], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_bnn_stats, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ngicp::g_bnn_stats, z, sizeof z); }
  return 0;
}
extern "C" int ngicp_debug_lin_hist(unsigned long long out[32], int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_lin_hist, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(ngicp::g_lin_hist, z, sizeof z); }
  return 0;
}
#endif

#ifdef NGICP_STATS
extern "C" int ngicp_debug_item_cycles(unsigned int* out, int n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ngicp::g_item_cycles, sizeof(unsigned int) * n);
  return 0;
}
#endif
