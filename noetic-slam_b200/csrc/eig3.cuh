// Plane regularisation of a symmetric PSD 3x3 covariance without an eigen-decomposition:
//   C_reg = U diag(1,1,eps) V^T = I - (1-eps) n n^T                       (nano_gicp.cc:365-385, PLANE)
// with n the unit eigenvector of the smallest eigenvalue lam3. For B = A - lam3 I (rank 2, symmetric)
//   adj(B) = (lam1-lam3)(lam2-lam3) n n^T     =>     n n^T = adj(B) / tr(adj(B)),
// so neither n nor a normalisation is ever formed. lam3 is the smallest root of the monic characteristic
// polynomial g(l) = l^3 - c2 l^2 + c1 l - c0: an fp32 trigonometric seed pushed just below the root, then
// three fp64 Newton steps (monotone from the left for a real-rooted cubic), each dividing through an
// approximate reciprocal (MUFU.RCP64H, ~20 bits: the iteration stays self-correcting). tr(adj(B)) = g'(lam).
// The fast path is accepted only when the last step proves convergence and the root is the smallest one;
// everything else (tiny spectral gap, degenerate neighbourhoods, non-finite input) takes the Jacobi path
// of the caller. fp64 work: 17 (coefficients) + 18 (Newton) + 28 (adjugate, scale, output) instructions,
// no fp64 division, square root or trigonometric call.
#pragma once
#include <math.h>

namespace ngicp {

#if defined(__CUDACC__)
#define NGICP_HD __host__ __device__ __forceinline__
#else
#define NGICP_HD inline
#endif

struct Sym3 { double xx, xy, xz, yy, yz, zz; };

// ~20-bit reciprocal of a normal, non-zero double
NGICP_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return (double)(1.0f / (float)x);
#endif
}

// returns false when the caller must take the exact (Jacobi) path
NGICP_HD bool plane_regularize_fast(const Sym3& A, Sym3& o) {
  // characteristic polynomial
  const double myz = A.yy * A.zz - A.yz * A.yz;  // principal minors
  const double mxz = A.xx * A.zz - A.xz * A.xz;
  const double mxy = A.xx * A.yy - A.xy * A.xy;
  const double c2 = A.xx + A.yy + A.zz;
  const double c1 = myz + mxz + mxy;
  const double k12 = A.xy * A.zz - A.yz * A.xz;
  const double k13 = A.xy * A.yz - A.yy * A.xz;
  const double c0 = A.xx * myz - A.xy * k12 + A.xz * k13;
  // fp32 seed: lam = q + 2 p cos(acos(r)/3 + 2pi/3), q = c2/3, p = sqrt(c2^2 - 3 c1)/3, r = -g(q) / (2 p^3)
  const double q = c2 * (1.0 / 3.0);
  const double gq = ((q - c2) * q + c1) * q - c0;
  const double p2 = c2 * c2 - 3.0 * c1;  // = 9 p^2 >= 0
  const float p2f = (float)p2, gqf = (float)gq;
#if defined(__CUDA_ARCH__)
  const float ip = rsqrtf(p2f);
#else
  const float ip = 1.0f / sqrtf(p2f);
#endif
  float r = -13.5f * gqf * ip * ip * ip;
  r = fminf(1.0f, fmaxf(-1.0f, r));
  const float phi = acosf(r) * (1.0f / 3.0f) + 2.0943951f;
#if defined(__CUDA_ARCH__)
  const float cs = __cosf(phi);
#else
  const float cs = cosf(phi);
#endif
  const float tf = (2.0f / 3.0f) * p2f * ip * cs;  // 2 p cos(.)
  // push the seed below the root: 4e-6 of the trace dominates the fp32 error of the seed away from a
  // vanishing gap (where the acceptance test below sends the point to the exact path anyway)
  double lam = (q + (double)tf) - 4e-6 * c2;
  const double m2c2 = -2.0 * c2;
  const double c2sq = c2 * c2;
  const double tol = 3e-17 * c2sq;  // |step| <= 5e-9 c2: the NEXT iterate is converged to the rounding floor of g
  double g, gp, d = 0.0;
#pragma unroll
  for (int it = 0; it < 3; it++) {
    g = ((lam - c2) * lam + c1) * lam - c0;
    gp = (3.0 * lam + m2c2) * lam + c1;
    d = g * rcp_seed(gp);
    lam -= d;
  }
  // small spectral gap (seed error comparable to the gap): Newton is still monotone but only linear at first
  for (int it = 0; it < 6 && d * d > tol; it++) {
    g = ((lam - c2) * lam + c1) * lam - c0;
    gp = (3.0 * lam + m2c2) * lam + c1;
    d = g * rcp_seed(gp);
    lam -= d;
  }
  // B = A - lam I, adj(B) (symmetric), tr(adj(B)) = g'(lam)
  const double bxx = A.xx - lam, byy = A.yy - lam, bzz = A.zz - lam;
  const double axx = byy * bzz - A.yz * A.yz;
  const double ayy = bxx * bzz - A.xz * A.xz;
  const double azz = bxx * byy - A.xy * A.xy;
  const double axy = A.xz * A.yz - A.xy * bzz;
  const double axz = A.xy * A.yz - A.xz * byy;
  const double ayz = A.xy * A.xz - A.yz * bxx;
  const double tr = axx + ayy + azz;
  // accept: the last Newton step d bounds the remaining error of lam by ~d^2 / gap; g'(lam) > 0 and lam <= q single
  // out the smallest root, and (gap/c2)(spread/c2) > 1e-9 keeps n defined (same bound as the exact path's switch)
  const bool ok = (d * d <= tol) && (tr > 1e-9 * c2sq) && (3.0 * lam <= c2);
  if (!ok) return false;
  double w = rcp_seed(tr);
  w = fma(w, fma(-tr, w, 1.0), w);
  w = fma(w, fma(-tr, w, 1.0), w);
  w *= (1.0 - 1e-3);
  o.xx = fma(-w, axx, 1.0); o.xy = -w * axy; o.xz = -w * axz;
  o.yy = fma(-w, ayy, 1.0); o.yz = -w * ayz; o.zz = fma(-w, azz, 1.0);
  return true;
}

}  // namespace ngicp
