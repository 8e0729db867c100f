// K3 — per-point k-neighbour 3x3 covariance + regularisation, closed-form symmetric eigen-solve in
// registers. Replaces NanoGICP::calculate_covariances, the part after the k-NN call
// (reference src/dlio/src/nano_gicp/nano_gicp.cc:345-389):
//   neighbours -> fp64, subtract the mean, cov = N N^T / k (divide by k, self included)   :348-354
//   PLANE (default): JacobiSVD, C = U diag(1,1,1e-3) V^T                                   :365-385
//   NONE / FROBENIUS / MIN_EIG / NORMALIZED_MIN_EIG                                        :356-363,375-381
// For a symmetric PSD matrix U == V (distinct, positive singular values), so PLANE is
// I - (1-1e-3) n n^T with n the eigenvector of the smallest eigenvalue; it is found in fp64 from the
// trigonometric eigenvalue formula and the largest cross product of two rows of (A - lambda I).
// Ill-conditioned neighbourhoods (smallest two eigenvalues ~equal) fall back to cyclic Jacobi.
// All arithmetic is fp64 from fp32 inputs, as in the reference; output is 6 x fp32 per point.
#include "internal.h"

namespace ngicp {

namespace {

struct Sym3 { double xx, xy, xz, yy, yz, zz; };

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

// unit eigenvector of the smallest eigenvalue of a symmetric PSD 3x3
__device__ __forceinline__ void smallest_eigvec(const Sym3& A, double n[3]) {
  const double scale = fmax(fmax(fabs(A.xx), fabs(A.yy)), fmax(fabs(A.zz), fmax(fabs(A.xy), fmax(fabs(A.xz), fabs(A.yz)))));
  bool ok = scale > 0.0;
  if (ok) {
    const double inv = 1.0 / scale;
    const double a00 = A.xx * inv, a01 = A.xy * inv, a02 = A.xz * inv, a11 = A.yy * inv, a12 = A.yz * inv, a22 = A.zz * inv;
    const double q = (a00 + a11 + a22) * (1.0 / 3.0);
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt(p2 * (1.0 / 6.0));
    double lam = q;
    if (p > 1e-300) {
      const double ip = 1.0 / p;
      const double c00 = b00 * ip, c01 = a01 * ip, c02 = a02 * ip, c11 = b11 * ip, c12 = a12 * ip, c22 = b22 * ip;
      double r = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
      r = fmin(1.0, fmax(-1.0, r));
      const double phi = acos(r) * (1.0 / 3.0);
      lam = q + 2.0 * p * cos(phi + 2.0943951023931954923);  // smallest root
    }
    // rows of (A - lam I); the eigenvector is orthogonal to all of them
    const double r0x = a00 - lam, r0y = a01, r0z = a02;
    const double r1x = a01, r1y = a11 - lam, r1z = a12;
    const double r2x = a02, r2y = a12, r2z = a22 - lam;
    const double c0x = r0y * r1z - r0z * r1y, c0y = r0z * r1x - r0x * r1z, c0z = r0x * r1y - r0y * r1x;
    const double c1x = r0y * r2z - r0z * r2y, c1y = r0z * r2x - r0x * r2z, c1z = r0x * r2y - r0y * r2x;
    const double c2x = r1y * r2z - r1z * r2y, c2y = r1z * r2x - r1x * r2z, c2z = r1x * r2y - r1y * r2x;
    const double n0 = c0x * c0x + c0y * c0y + c0z * c0z;
    const double n1 = c1x * c1x + c1y * c1y + c1z * c1z;
    const double n2 = c2x * c2x + c2y * c2y + c2z * c2z;
    double bx = c0x, by = c0y, bz = c0z, bn = n0;
    if (n1 > bn) { bx = c1x; by = c1y; bz = c1z; bn = n1; }
    if (n2 > bn) { bx = c2x; by = c2y; bz = c2z; bn = n2; }
    // |cross| ~ (lam_mid - lam_min)(lam_max - lam_min) on the unit-scaled matrix; tiny => the two
    // smallest eigenvalues (nearly) coincide and the direction is ill-defined: use Jacobi.
    if (bn > 1e-16) {
      const double rn = rsqrt(bn);
      n[0] = bx * rn; n[1] = by * rn; n[2] = bz * rn;
      return;
    }
  }
  double w[3], V[3][3];
  jacobi_eig3(A, w, V);
  int m = 0;
  if (w[1] < w[m]) m = 1;
  if (w[2] < w[m]) m = 2;
  if (!(w[0] == w[0])) { n[0] = 0; n[1] = 0; n[2] = 1; return; }
  // exact ties (e.g. the all-zero matrix): the reference's sorted SVD keeps the LAST column small
  if (w[0] == w[1] && w[1] == w[2]) m = 2;
  n[0] = V[0][m]; n[1] = V[1][m]; n[2] = V[2][m];
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
    double n[3];
    smallest_eigvec(cov, n);
    const double f = 1.0 - 1e-3;
    Sym3 o;
    o.xx = 1.0 - f * n[0] * n[0]; o.xy = -f * n[0] * n[1]; o.xz = -f * n[0] * n[2];
    o.yy = 1.0 - f * n[1] * n[1]; o.yz = -f * n[1] * n[2]; o.zz = 1.0 - f * n[2] * n[2];
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

// K = compile-time k (vector index loads, fully unrolled gathers) or 0 for a runtime k.
template <int K, int REG>
__global__ void __launch_bounds__(128) covariance_kernel(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k_rt,
                                                         float* __restrict__ cov6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int k = K > 0 ? K : k_rt;
  const float4 pj = __ldg(pts + j);
  const double ox = (double)pj.x, oy = (double)pj.y, oz = (double)pj.z;
  double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
  const int* row = nbr + (size_t)j * k;
  if (K > 0) {
    int id[K > 0 ? K : 1];
#pragma unroll
    for (int i = 0; i < K; i += 4) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(row) + i / 4);
      id[i] = v.x; id[i + 1] = v.y; id[i + 2] = v.z; id[i + 3] = v.w;
    }
    // gathers in batches of kBatch: enough loads in flight per thread, few enough registers for ~40 warps per SM
    // (the kernel is bound by the fp64 pipe and needs the occupancy to keep it fed)
    constexpr int kBatch = (K > 0 && K % 8 == 0) ? 8 : 4;
#pragma unroll
    for (int b0 = 0; b0 < K; b0 += kBatch) {
      float4 nb[kBatch];
#pragma unroll
      for (int i = 0; i < kBatch; i++) nb[i] = __ldg(pts + id[b0 + i]);
#pragma unroll
      for (int i = 0; i < kBatch; i++) {
        const double dx = (double)nb[i].x - ox, dy = (double)nb[i].y - oy, dz = (double)nb[i].z - oz;
        sx += dx; sy += dy; sz += dz;
        sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
      }
    }
  } else {
    for (int i = 0; i < k; i++) {
      const float4 p = __ldg(pts + __ldg(row + i));
      const double dx = (double)p.x - ox, dy = (double)p.y - oy, dz = (double)p.z - oz;
      sx += dx; sy += dy; sz += dz;
      sxx += dx * dx; sxy += dx * dy; sxz += dx * dz; syy += dy * dy; syz += dy * dz; szz += dz * dz;
    }
  }
  // cov = (S - s s^T / k) / k : covariance about the mean, divided by k (nano_gicp.cc:353-354)
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

int covariances_from_knn(Handle* h, const Index* idx, const int* d_nbr, int k, int reg, float* d_cov6) {
  const int n = idx->n;
  const int nb = (n + 127) / 128;
  cudaStream_t s = h->stream;
#define LAUNCH_COV(K, REG) covariance_kernel<K, REG><<<nb, 128, 0, s>>>(idx->pts, d_nbr, n, k, d_cov6)
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
