// Host side of the optimiser: the 6x6 Gauss-Newton / Levenberg-Marquardt step on SE(3).
// Restates LsqRegistration (reference src/dlio/src/nano_gicp/lsq_registration.cc:108-229,
// include/nano_gicp/lsq_registration.h:70-101) with fixed-size arrays; the north star keeps exactly
// this part on the host. linearize()/compute_error() are the device calls (K4/K5).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ngicp {
namespace lm {

struct Iso {  // x -> R x + t, fp64 (Eigen::Isometry3d)
  double R[9];  // row-major
  double t[3];
};
inline Iso identity() { Iso a{}; a.R[0] = a.R[4] = a.R[8] = 1.0; return a; }
inline Iso compose(const Iso& a, const Iso& b) {  // a * b
  Iso c{};
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a.R[3 * i + k] * b.R[3 * k + j];
      c.R[3 * i + j] = s;
    }
    c.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
  }
  return c;
}
inline void to_colmajor(const Iso& x, double T[16]) {
  std::memset(T, 0, 16 * sizeof(double));
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T[4 * c + r] = x.R[3 * r + c];
    T[12 + r] = x.t[r];
  }
  T[15] = 1.0;
}
inline Iso from_colmajor_f(const float T[16]) {
  Iso x{};
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) x.R[3 * r + c] = (double)T[4 * c + r];
    x.t[r] = (double)T[12 + r];
  }
  return x;
}

// so3_exp (lsq_registration.h:82-101) followed by Quaterniond::toRotationMatrix (un-normalised)
inline void so3_exp(const double w[3], double R[9]) {
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
inline Iso delta_from(const double d[6]) {
  Iso x{};
  so3_exp(d, x.R);
  x.t[0] = d[3]; x.t[1] = d[4]; x.t[2] = d[5];
  return x;
}

// is_converged (lsq_registration.cc:137-146)
inline bool is_converged(const Iso& delta, double rot_eps, double trans_eps) {
  double rmax = 0, tmax = 0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) rmax = std::max(rmax, 1.0 / rot_eps * std::fabs(delta.R[3 * r + c] - (r == c ? 1.0 : 0.0)));
    tmax = std::max(tmax, 1.0 / trans_eps * std::fabs(delta.t[r]));
  }
  return std::max(rmax, tmax) < 1;
}

// (H) d = rhs for symmetric 6x6 H, LDL^T with diagonal pivoting (Eigen::LDLT, lsq_registration.cc:166-167,192-193)
inline void solve6(const double Hin[36], const double rhs[6], double x[6]) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Hin, sizeof A);
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
    for (int i = k + 1; i < n; i++) L[i * n + k] = D[k] != 0.0 ? A[i * n + k] / D[k] : 0.0;
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j < n; j++) A[i * n + j] -= L[i * n + k] * D[k] * L[j * n + k];
  }
  double y[6], z[6];
  for (int i = 0; i < n; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i * n + j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) y[i] = D[i] != 0.0 ? y[i] / D[i] : 0.0;
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= L[j * n + i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < n; i++) x[perm[i]] = z[i];
}

}  // namespace lm
}  // namespace ngicp
