// TEST INFRASTRUCTURE — part of the CPU oracle. Not shipped, not on the product path.
//
// Tiny fixed-size fp64 linear algebra used by oracle.cc to restate the Eigen calls the
// reference makes on the nano_gicp path (Eigen is absent from this image):
//   * JacobiSVD<Matrix3d>(ComputeFullU|ComputeFullV)   nano_gicp.cc:365
//   * Matrix3d::inverse(), Matrix3d::norm()            nano_gicp.cc:360-363
//   * Matrix4d::inverse() of blockdiag(A,1)            nano_gicp.cc:240
//   * LDLT<Matrix<double,6,6>>::solve                  lsq_registration.cc:166-167,192-193
//   * Quaterniond::toRotationMatrix                    lsq_registration.cc:196
// Rounding differs from Eigen at O(1e-15) relative, far inside the 1e-4 parity tolerances.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

namespace orc {

// 3x3, row-major: m[3*r+c]
struct M3 {
  double m[9];
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
  static M3 zero() { M3 a; std::memset(a.m, 0, sizeof a.m); return a; }
  static M3 identity() { M3 a = zero(); a(0, 0) = a(1, 1) = a(2, 2) = 1.0; return a; }
};

inline M3 mul(const M3& a, const M3& b) {
  M3 c = M3::zero();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline M3 transpose(const M3& a) {
  M3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t(i, j) = a(j, i);
  return t;
}
inline M3 add(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 9; i++) c.m[i] = a.m[i] + b.m[i];
  return c;
}
inline double det(const M3& a) {
  return a(0, 0) * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) -
         a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
         a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
}
// cofactor inverse (what Eigen does for fixed 3x3)
inline M3 inverse(const M3& a) {
  M3 c;
  c(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  c(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2);
  c(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
  c(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  c(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0);
  c(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
  c(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  c(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1);
  c(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
  const double d = a(0, 0) * c(0, 0) + a(0, 1) * c(1, 0) + a(0, 2) * c(2, 0);
  const double inv = 1.0 / d;
  for (int i = 0; i < 9; i++) c.m[i] *= inv;
  return c;
}
inline double frobenius(const M3& a) {
  double s = 0.0;
  for (int i = 0; i < 9; i++) s += a.m[i] * a.m[i];
  return std::sqrt(s);
}

// One-sided (Hestenes) Jacobi SVD of a general 3x3: A = U diag(s) V^T, s descending, U and V
// full orthogonal. Restates what the reference asks of Eigen::JacobiSVD (nano_gicp.cc:365).
inline void svd3(const M3& A, M3& U, double s[3], M3& V) {
  M3 W = A;
  V = M3::identity();
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; r++) {
          alpha += W(r, p) * W(r, p);
          beta += W(r, q) * W(r, q);
          gamma += W(r, p) * W(r, q);
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-300) continue;
        if (std::fabs(gamma) <= 2.3e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int r = 0; r < 3; r++) {
          const double wp = W(r, p), wq = W(r, q);
          W(r, p) = c * wp - sn * wq;
          W(r, q) = sn * wp + c * wq;
          const double vp = V(r, p), vq = V(r, q);
          V(r, p) = c * vp - sn * vq;
          V(r, q) = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  for (int j = 0; j < 3; j++) nrm[j] = std::sqrt(W(0, j) * W(0, j) + W(1, j) * W(1, j) + W(2, j) * W(2, j));
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return nrm[a] > nrm[b]; });
  M3 Vs, Us = M3::zero();
  for (int j = 0; j < 3; j++) {
    s[j] = nrm[ord[j]];
    for (int r = 0; r < 3; r++) Vs(r, j) = V(r, ord[j]);
    if (s[j] > 0.0)
      for (int r = 0; r < 3; r++) Us(r, j) = W(r, ord[j]) / s[j];
  }
  // complete U for (numerically) zero singular values so that it stays orthogonal
  const double tiny = s[0] * 1e-300;
  auto col = [&](const M3& m, int j, double* o) { o[0] = m(0, j); o[1] = m(1, j); o[2] = m(2, j); };
  auto setcol = [&](M3& m, int j, const double* o) { m(0, j) = o[0]; m(1, j) = o[1]; m(2, j) = o[2]; };
  auto cross = [](const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
  };
  if (!(s[0] > 0.0)) {
    Us = M3::identity();
  } else {
    if (!(s[1] > tiny)) {
      double u0[3]; col(Us, 0, u0);
      double e[3] = {0, 0, 0};
      int k = 0;
      if (std::fabs(u0[1]) < std::fabs(u0[k])) k = 1;
      if (std::fabs(u0[2]) < std::fabs(u0[k])) k = 2;
      e[k] = 1.0;
      double u1[3]; cross(u0, e, u1);
      const double n = std::sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
      for (double& v : u1) v /= n;
      setcol(Us, 1, u1);
    }
    if (!(s[2] > tiny)) {
      double u0[3], u1[3], u2[3];
      col(Us, 0, u0); col(Us, 1, u1); cross(u0, u1, u2);
      const double n = std::sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
      for (double& v : u2) v /= n;
      setcol(Us, 2, u2);
    }
  }
  U = Us;
  V = Vs;
}

// 6x6 symmetric solve through a pivoted LDL^T (Eigen::LDLT pivots on the largest diagonal).
// A row-major 6x6, returns x with A x = rhs.
inline void ldlt6_solve(const double Ain[36], const double rhs[6], double x[6]) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Ain, sizeof A);
  int perm[6] = {0, 1, 2, 3, 4, 5};
  double L[36] = {0}, D[6];
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i * n + i]) > std::fabs(A[piv * n + piv])) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[piv * n + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * n + k], A[i * n + piv]);
      for (int j = 0; j < k; j++) std::swap(L[k * n + j], L[piv * n + j]);
      std::swap(perm[k], perm[piv]);
    }
    D[k] = A[k * n + k];
    L[k * n + k] = 1.0;
    for (int i = k + 1; i < n; i++) L[i * n + k] = (D[k] != 0.0) ? A[i * n + k] / D[k] : 0.0;
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j < n; j++) A[i * n + j] -= L[i * n + k] * D[k] * L[j * n + k];
  }
  double y[6], z[6];
  for (int i = 0; i < n; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i * n + j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) y[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= L[j * n + i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < n; i++) x[perm[i]] = z[i];
}

// Rigid transform, fp64: x -> R x + t.  (Eigen::Isometry3d in the reference.)
struct Iso3 {
  M3 R;
  double t[3];
  static Iso3 identity() { Iso3 a; a.R = M3::identity(); a.t[0] = a.t[1] = a.t[2] = 0; return a; }
};
inline Iso3 compose(const Iso3& a, const Iso3& b) {  // a * b
  Iso3 c;
  c.R = mul(a.R, b.R);
  for (int i = 0; i < 3; i++) c.t[i] = a.R(i, 0) * b.t[0] + a.R(i, 1) * b.t[1] + a.R(i, 2) * b.t[2] + a.t[i];
  return c;
}

// Quaternion (w,x,y,z) -> rotation matrix, the formula Eigen::Quaternion::toRotationMatrix uses.
inline M3 quat_to_rot(double w, double x, double y, double z) {
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3 r;
  r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz;       r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;       r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;       r(2, 1) = tyz + twx;       r(2, 2) = 1 - (txx + tyy);
  return r;
}

}  // namespace orc
