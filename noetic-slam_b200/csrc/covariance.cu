// K3 — per-point k-neighbour 3x3 covariance + regularisation, closed-form symmetric eigen-solve in
// registers. Replaces NanoGICP::calculate_covariances, the part after the k-NN call
// (reference src/dlio/src/nano_gicp/nano_gicp.cc:345-389):
//   neighbours -> fp64, subtract the mean, cov = N N^T / k (divide by k, self included)   :348-354
//   PLANE (default): JacobiSVD, C = U diag(1,1,1e-3) V^T                                   :365-385
//   NONE / FROBENIUS / MIN_EIG / NORMALIZED_MIN_EIG                                        :356-363,375-381
// For a symmetric PSD matrix U == V (distinct, positive singular values), so PLANE is
// I - (1-1e-3) n n^T with n the eigenvector of the smallest eigenvalue: n n^T = adj(A - lam I) / tr(adj), lam from
// three fp64 Newton steps on the characteristic polynomial (eig3.cuh).
// Ill-conditioned neighbourhoods (smallest two eigenvalues ~equal) fall back to cyclic Jacobi.
// All arithmetic is fp64 from fp32 inputs, as in the reference; output is 6 x fp32 per point.
#include "internal.h"
#include "eig3.cuh"
#include "linearize.cuh"

namespace ngicp {

namespace {

// cyclic Jacobi eigen-decomposition of a symmetric 3x3: A = V diag(w) V^T (columns of V)
__device__ __noinline__ void jacobi_eig3(const Sym3& A, double w[3], double V[3][3]) {
  double a[3][3] = {{A.xx, A.xy, A.xz}, {A.xy, A.yy, A.yz}, {A.xz, A.yz, A.zz}};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-18 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < 3; r++) {  // A <- A J
          const double arp = a[r][p], arq = a[r][q];
          a[r][p] = c * arp - s * arq; a[r][q] = s * arp + c * arq;
        }
        for (int r = 0; r < 3; r++) {  // A <- J^T A
          const double apr = a[p][r], aqr = a[q][r];
          a[p][r] = c * apr - s * aqr; a[q][r] = s * apr + c * aqr;
        }
        for (int r = 0; r < 3; r++) {
          const double vrp = V[r][p], vrq = V[r][q];
          V[r][p] = c * vrp - s * vrq; V[r][q] = s * vrp + c * vrq;
        }
      }
  }
  w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
}

// exact path of PLANE for the neighbourhoods plane_regularize_fast (eig3.cuh) refuses: unit eigenvector of the
// smallest eigenvalue by cyclic Jacobi
__device__ __noinline__ void plane_regularize_exact(const Sym3& A, Sym3& o) {
  double w[3], V[3][3], n[3];
  jacobi_eig3(A, w, V);
  int m = 0;
  if (w[1] < w[m]) m = 1;
  if (w[2] < w[m]) m = 2;
  // exact ties (e.g. the all-zero matrix): the reference's sorted SVD keeps the LAST column small
  if (w[0] == w[1] && w[1] == w[2]) m = 2;
  n[0] = V[0][m]; n[1] = V[1][m]; n[2] = V[2][m];
  if (!(w[0] == w[0])) { n[0] = 0; n[1] = 0; n[2] = 1; }
  const double f = 1.0 - 1e-3;
  o.xx = 1.0 - f * n[0] * n[0]; o.xy = -f * n[0] * n[1]; o.xz = -f * n[0] * n[2];
  o.yy = 1.0 - f * n[1] * n[1]; o.yz = -f * n[1] * n[2]; o.zz = 1.0 - f * n[2] * n[2];
}

__device__ __forceinline__ Sym3 inverse_sym3(const Sym3& a) {
  Sym3 c;
  c.xx = a.yy * a.zz - a.yz * a.yz;
  c.xy = a.xz * a.yz - a.xy * a.zz;
  c.xz = a.xy * a.yz - a.xz * a.yy;
  c.yy = a.xx * a.zz - a.xz * a.xz;
  c.yz = a.xy * a.xz - a.xx * a.yz;
  c.zz = a.xx * a.yy - a.xy * a.xy;
  const double det = a.xx * c.xx + a.xy * c.xy + a.xz * c.xz;
  const double inv = 1.0 / det;
  c.xx *= inv; c.xy *= inv; c.xz *= inv; c.yy *= inv; c.yz *= inv; c.zz *= inv;
  return c;
}

template <int REG>
__device__ __forceinline__ Sym3 regularize(const Sym3& cov) {
  if (REG == NGICP_REG_NONE) return cov;
  if (REG == NGICP_REG_PLANE) {
    Sym3 o;
    if (!plane_regularize_fast(cov, o)) plane_regularize_exact(cov, o);
    return o;
  }
  if (REG == NGICP_REG_FROBENIUS) {  // nano_gicp.cc:358-363
    Sym3 C = cov;
    C.xx += 1e-3; C.yy += 1e-3; C.zz += 1e-3;
    Sym3 Ci = inverse_sym3(C);
    const double nrm = sqrt(Ci.xx * Ci.xx + Ci.yy * Ci.yy + Ci.zz * Ci.zz + 2.0 * (Ci.xy * Ci.xy + Ci.xz * Ci.xz + Ci.yz * Ci.yz));
    const double inv = 1.0 / nrm;
    Ci.xx *= inv; Ci.xy *= inv; Ci.xz *= inv; Ci.yy *= inv; Ci.yz *= inv; Ci.zz *= inv;
    return inverse_sym3(Ci);
  }
  // MIN_EIG / NORMALIZED_MIN_EIG (nano_gicp.cc:375-381): clamp the spectrum from below
  double w[3], V[3][3];
  jacobi_eig3(cov, w, V);
  double vals[3];
  const double mx = fmax(fabs(w[0]), fmax(fabs(w[1]), fabs(w[2])));
  for (int i = 0; i < 3; i++) {
    const double sv = fabs(w[i]);  // singular value of a symmetric matrix
    const double v = REG == NGICP_REG_NORMALIZED_MIN_EIG ? sv / mx : sv;
    vals[i] = (w[i] < 0 ? -1.0 : 1.0) * fmax(v, 1e-3);  // U column = sign * V column
  }
  Sym3 o;
  o.xx = vals[0] * V[0][0] * V[0][0] + vals[1] * V[0][1] * V[0][1] + vals[2] * V[0][2] * V[0][2];
  o.xy = vals[0] * V[0][0] * V[1][0] + vals[1] * V[0][1] * V[1][1] + vals[2] * V[0][2] * V[1][2];
  o.xz = vals[0] * V[0][0] * V[2][0] + vals[1] * V[0][1] * V[2][1] + vals[2] * V[0][2] * V[2][2];
  o.yy = vals[0] * V[1][0] * V[1][0] + vals[1] * V[1][1] * V[1][1] + vals[2] * V[1][2] * V[1][2];
  o.yz = vals[0] * V[1][0] * V[2][0] + vals[1] * V[1][1] * V[2][1] + vals[2] * V[1][2] * V[2][2];
  o.zz = vals[0] * V[2][0] * V[2][0] + vals[1] * V[2][1] * V[2][1] + vals[2] * V[2][2] * V[2][2];
  return o;
}

// scatter matrix of the K gathered neighbours about the query point, covariance, regularisation, 24-byte store.
// Neighbour 0 is the query itself (or an exact duplicate of it, distance 0): its difference is exactly zero, so it is
// neither gathered nor accumulated — it only counts in the divisor k.
template <int K, int REG, typename Row4>
__device__ __forceinline__ void covariance_from_ids(const float4* __restrict__ pts, const float4 pj, Row4 row4, float* __restrict__ cov6, int j) {
  const double ox = (double)pj.x, oy = (double)pj.y, oz = (double)pj.z;
  double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
  // gathers in batches of 8: enough loads in flight per thread, few enough registers for 28 warps per SM
  int id[K];
#pragma unroll
  for (int i = 0; i < K; i += 4) {
    const int4 v = row4(i / 4);
    id[i] = v.x; id[i + 1] = v.y; id[i + 2] = v.z; id[i + 3] = v.w;
  }
#pragma unroll
  for (int b0 = 1; b0 < K; b0 += 8) {
    constexpr int kB = 8;
    float4 nb[kB];
#pragma unroll
    for (int i = 0; i < kB; i++) if (b0 + i < K) nb[i] = __ldg(pts + id[b0 + i]);
#pragma unroll
    for (int i = 0; i < kB; i++) if (b0 + i < K) {
      const double dx = (double)nb[i].x - ox, dy = (double)nb[i].y - oy, dz = (double)nb[i].z - oz;
      sx += dx; sy += dy; sz += dz;
      sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
  }
  // cov = (S - s s^T / k) / k : covariance about the mean, divided by k (nano_gicp.cc:353-354)
  const double ik = 1.0 / (double)K;
  Sym3 c;
  c.xx = (sxx - sx * sx * ik) * ik; c.xy = (sxy - sx * sy * ik) * ik; c.xz = (sxz - sx * sz * ik) * ik;
  c.yy = (syy - sy * sy * ik) * ik; c.yz = (syz - sy * sz * ik) * ik; c.zz = (szz - sz * sz * ik) * ik;
  const Sym3 o = regularize<REG>(c);
  float2* out = reinterpret_cast<float2*>(cov6 + (size_t)j * 6);
  __stcs(out, make_float2((float)o.xx, (float)o.xy));
  __stcs(out + 1, make_float2((float)o.xz, (float)o.yy));
  __stcs(out + 2, make_float2((float)o.yz, (float)o.zz));
}

// One thread per point. K = compile-time k (vector index loads, fully unrolled gathers) or 0 for a runtime k.
template <int K, int REG>
__device__ __forceinline__ void covariance_point(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k_rt,
                                                 float* __restrict__ cov6, int j) {
  const float4 pj = __ldg(pts + j);
  if constexpr (K > 0) {
    // tiled k-NN table (internal.h:nbr_tiled): chunk c of the 32 points of a tile is contiguous, so a warp reads 512
    // contiguous bytes per load (4 L1 wavefronts instead of the 16 of a 64-byte-strided row read)
    const int4* tile = reinterpret_cast<const int4*>(nbr) + (size_t)(j >> 5) * (K / 4) * 32 + (j & 31);
    // the index rows are read once: streaming loads (evict-first) leave L1 / L2 to the gathered neighbour points
    covariance_from_ids<K, REG>(pts, pj, [&](int c) { return __ldcs(tile + c * 32); }, cov6, j);
  } else {
    const int k = k_rt;
    // the table is tiled for the k that K2 writes tiled (internal.h:nbr_tiled), row-major otherwise
    const bool tiled = k == 16 || k == 20;
    const int* row = tiled ? nbr + ((size_t)(j >> 5) * (k / 4) * 32 + (j & 31)) * 4 : nbr + (size_t)j * k;
    const double ox = (double)pj.x, oy = (double)pj.y, oz = (double)pj.z;
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    for (int i = 1; i < k; i++) {     // neighbour 0 is the query itself: zero difference
      const float4 p = __ldg(pts + __ldg(tiled ? row + (i >> 2) * 128 + (i & 3) : row + i));
      const double dx = (double)p.x - ox, dy = (double)p.y - oy, dz = (double)p.z - oz;
      sx += dx; sy += dy; sz += dz;
      sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
    const double ik = 1.0 / (double)k;
    Sym3 c;
    c.xx = (sxx - sx * sx * ik) * ik; c.xy = (sxy - sx * sy * ik) * ik; c.xz = (sxz - sx * sz * ik) * ik;
    c.yy = (syy - sy * sy * ik) * ik; c.yz = (syz - sy * sz * ik) * ik; c.zz = (szz - sz * sz * ik) * ik;
    const Sym3 o = regularize<REG>(c);
    float2* out = reinterpret_cast<float2*>(cov6 + (size_t)j * 6);
    out[0] = make_float2((float)o.xx, (float)o.xy);
    out[1] = make_float2((float)o.xz, (float)o.yy);
    out[2] = make_float2((float)o.yz, (float)o.zz);
  }
}

template <int K, int REG>
__global__ void __launch_bounds__(128) covariance_kernel(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k_rt,
                                                         float* __restrict__ cov6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) covariance_point<K, REG>(pts, nbr, n, k_rt, cov6, j);
}

// Single clouds: the density sum of nano_gicp.cc:389 (per-point terms from K2) rides on the same launch — block sums,
// then the fixed-order fold and host-mapped result slot of K4b / K5 (linearize.cuh:block_publish) — instead of a reduction
// kernel of its own behind K3.
template <int K, int REG>
__global__ void __launch_bounds__(kLinThreads) covariance_density_kernel(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k_rt,
                                                                          float* __restrict__ cov6, const double* __restrict__ dens_term,
                                                                          double* __restrict__ partials, unsigned int* __restrict__ counters,
                                                                          ReduceSlot* __restrict__ slots, unsigned long long seq) {
  __shared__ double wsum[kLinThreads / 32][1];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (j < n) {
    v = __ldg(dens_term + j);
    covariance_point<K, REG>(pts, nbr, n, k_rt, cov6, j);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5][0] = v;
  block_publish<1>(wsum, partials, counters, slots, seq);
}

// ---- layout conversions between the host's CovarianceList order and the device's sorted order ----
__global__ void __launch_bounds__(256) cov6_to_mat4_kernel(const float4* __restrict__ pts, const float* __restrict__ cov6, int n, double* __restrict__ out16) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int orig = __float_as_int(__ldg(&pts[j].w));
  const float* c = cov6 + (size_t)j * 6;
  double* o = out16 + (size_t)orig * 16;
  const double xx = c[0], xy = c[1], xz = c[2], yy = c[3], yz = c[4], zz = c[5];
  o[0] = xx; o[1] = xy; o[2] = xz; o[3] = 0;
  o[4] = xy; o[5] = yy; o[6] = yz; o[7] = 0;
  o[8] = xz; o[9] = yz; o[10] = zz; o[11] = 0;
  o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 0;
}
__global__ void __launch_bounds__(256) mat4_to_cov6_kernel(const float4* __restrict__ pts, const double* __restrict__ in16, int n, float* __restrict__ cov6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int orig = __float_as_int(__ldg(&pts[j].w));
  const double* m = in16 + (size_t)orig * 16;  // column-major; upper triangle (r<=c) = m[4c+r]
  float* c = cov6 + (size_t)j * 6;
  c[0] = (float)m[0]; c[1] = (float)m[4]; c[2] = (float)m[8]; c[3] = (float)m[5]; c[4] = (float)m[9]; c[5] = (float)m[10];
}
__global__ void __launch_bounds__(256) cov6_unsort_kernel(const float4* __restrict__ pts, const float* __restrict__ cov6, int n, float* __restrict__ out6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int orig = __float_as_int(__ldg(&pts[j].w));
  for (int i = 0; i < 6; i++) out6[(size_t)orig * 6 + i] = cov6[(size_t)j * 6 + i];
}

__global__ void __launch_bounds__(256) cov6_sort_kernel(const float4* __restrict__ pts, const float* __restrict__ in6, int n, float* __restrict__ cov6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int orig = __float_as_int(__ldg(&pts[j].w));
  for (int i = 0; i < 6; i++) cov6[(size_t)j * 6 + i] = in6[(size_t)orig * 6 + i];
}

}  // namespace

int cov6_from_host_order(Handle* h, const Index* idx, const float* d_in6, float* d_cov6) {
  cov6_sort_kernel<<<(idx->n + 255) / 256, 256, 0, h->stream>>>(idx->pts, d_in6, idx->n, d_cov6);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

int covariances_from_knn(Handle* h, const Index* idx, const int* d_nbr, int k, int reg, float* d_cov6, const double* d_dens_term, double* density_sum) {
  const int n = idx->n;
  const int nb = (n + 127) / 128;
  cudaStream_t s = h->stream;
  // density fused into the launch when the block partials fit the handle's reduction buffers (single scans)
  const bool fuse = d_dens_term && density_sum && idx->n_seg == 1 && nb <= kMaxLinBlocks;
  const unsigned long long seq = fuse ? ++h->seq : 0ull;
#define LAUNCH_COV(K, REG)                                                                                                              \
  do {                                                                                                                                  \
    if (fuse) covariance_density_kernel<K, REG><<<nb, kLinThreads, 0, s>>>(idx->pts, d_nbr, n, k, d_cov6, d_dens_term, h->partials,     \
                                                                           h->counter, h->slot_dev, seq);                              \
    else covariance_kernel<K, REG><<<nb, 128, 0, s>>>(idx->pts, d_nbr, n, k, d_cov6);                                                   \
  } while (0)
#define LAUNCH_COV_K(REG)            \
  do {                               \
    if (k == 16) LAUNCH_COV(16, REG); \
    else if (k == 20) LAUNCH_COV(20, REG); \
    else LAUNCH_COV(0, REG);         \
  } while (0)
  switch (reg) {
    case NGICP_REG_NONE: LAUNCH_COV_K(NGICP_REG_NONE); break;
    case NGICP_REG_MIN_EIG: LAUNCH_COV(0, NGICP_REG_MIN_EIG); break;
    case NGICP_REG_NORMALIZED_MIN_EIG: LAUNCH_COV(0, NGICP_REG_NORMALIZED_MIN_EIG); break;
    case NGICP_REG_PLANE: LAUNCH_COV_K(NGICP_REG_PLANE); break;
    case NGICP_REG_FROBENIUS: LAUNCH_COV(0, NGICP_REG_FROBENIUS); break;
    default: return fail(h, NGICP_ERR_INVALID, "unknown regularization method");  // reference aborts (nano_gicp.cc:369-371)
  }
#undef LAUNCH_COV_K
#undef LAUNCH_COV
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  if (fuse) {
    if (int rc = wait_slot(h, 1, seq)) return rc;
    *density_sum = h->slot_host[0].v[0];
  } else if (d_dens_term && density_sum && idx->n_seg == 1) {
    if (int rc = reduce_sum(h, d_dens_term, n, idx->seg_start, 1, density_sum)) return rc;
  }
  return NGICP_OK;
}

int cov6_to_mat4_host_order(Handle* h, const Index* idx, const float* d_cov6, double* d_out16) {
  cov6_to_mat4_kernel<<<(idx->n + 255) / 256, 256, 0, h->stream>>>(idx->pts, d_cov6, idx->n, d_out16);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}
int mat4_host_order_to_cov6(Handle* h, const Index* idx, const double* d_in16, float* d_cov6) {
  mat4_to_cov6_kernel<<<(idx->n + 255) / 256, 256, 0, h->stream>>>(idx->pts, d_in16, idx->n, d_cov6);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}
int cov6_to_host_order(Handle* h, const Index* idx, const float* d_cov6, float* d_out6) {
  cov6_unsort_kernel<<<(idx->n + 255) / 256, 256, 0, h->stream>>>(idx->pts, d_cov6, idx->n, d_out6);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

}  // namespace ngicp
