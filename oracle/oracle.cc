// TEST INFRASTRUCTURE — CPU ORACLE. Not shipped, never on the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / the reported CPU baseline.
//
// What it is: a CPU restatement of the reference's nano_gicp scan-to-map path
// (reference = /root/reference/src/dlio, citations below are relative to that directory):
//   src/nano_gicp/nano_gicp.cc:206-245   update_correspondences
//   src/nano_gicp/nano_gicp.cc:248-302   linearize
//   src/nano_gicp/nano_gicp.cc:305-326   compute_error
//   src/nano_gicp/nano_gicp.cc:330-392   calculate_covariances
//   src/nano_gicp/lsq_registration.cc:108-229, include/nano_gicp/lsq_registration.h:70-101  LM on SE(3)
//   include/nano_gicp/nanoflann_adaptor.h:141-152  nearestKSearch wrapper semantics
// Eigen, PCL and Boost are absent from this image, so those .cc files cannot be compiled; the
// arithmetic is restated here with oracle/linalg_small.h standing in for the Eigen calls.
//
// Two builds of this one file (see oracle/Makefile):
//   * liboracle_port.so  — self-contained. The neighbour search is an independent exact k-NN
//     (median-split k-d tree, fp32 distance ((dx*dx)+(dy*dy))+(dz*dz), result order = ascending
//     (distance, index)). This is the build that travels and that tests use as the checker.
//   * _ref/liboracle_ref.so — compiled with -DORC_USE_REF_NANOFLANN and
//     -I/root/reference/src/dlio/include, so the neighbour search IS the reference's own
//     nanoflann.h (KDTreeSingleIndexAdaptor<SO3_Adaptor<float,..>,..,3,int>, leaf 25, exactly the
//     instantiation of nanoflann_adaptor.h:100-102,114). Used to pin the port and as CPU baseline.
//
// Parity pin status: the reference holds NO tests / golden vectors for this path;
// k-NN is pinned against the reference's own nanoflann compiled here; covariance / linearise /
// LM are pinned only against numpy/scipy cross-checks => "parity unpinned" for those.
//
// Build flags follow the reference's Release build: no FMA contraction, no -march=native
// (src/dlio/CMakeLists.txt:16-17): -O3 -fopenmp -ffp-contract=off.

#include <cstdint>
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <utility>
#include <memory>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "linalg_small.h"

#ifdef ORC_USE_REF_NANOFLANN
#include <nano_gicp/nanoflann.h>
#endif

namespace orc {

// ------------------------------------------------------------------------------------------
// Point cloud view: the reference's 32-byte AoS dlio::Point (include/dlio/dlio.h:85-108) has
// xyz at floats 0..2 and w=1 at float 3; any float stride >= 3 is accepted here.
// ------------------------------------------------------------------------------------------
struct Cloud {
  std::vector<float> xyz;  // packed N x 3 copy
  size_t n = 0;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, int dim) const { return xyz[3 * idx + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
  const float* pt(size_t i) const { return &xyz[3 * i]; }
};

static inline float sqdist_f32(const float* a, const float* b) {
  // nanoflann.h:509-520 (L2_Simple_Adaptor::evalMetric): result += diff*diff over dims 0,1,2, fp32
  float r = 0.0f;
  for (int i = 0; i < 3; i++) {
    const float d = a[i] - b[i];
    r += d * d;
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Port k-NN: independent exact search. Canonical order (distance asc, then index asc).
// ------------------------------------------------------------------------------------------
class PortTree {
 public:
  explicit PortTree(std::shared_ptr<const Cloud> c) : cloud_(std::move(c)) {
    const size_t n = cloud_->n;
    perm_.resize(n);
    std::iota(perm_.begin(), perm_.end(), 0);
    if (n) build(0, n);
  }
  // returns number found (<= k)
  int knn(const float* q, int k, int* idx, float* sqd) const {
    if (cloud_->n == 0 || k <= 0) return 0;
    Result r{idx, sqd, k, 0};
    search(0, q, r);
    return r.count;
  }
  const std::shared_ptr<const Cloud>& cloud() const { return cloud_; }

 private:
  struct Node {
    double lo[3], hi[3];
    int left = -1, right = -1;  // children node ids; leaf if left < 0
    size_t begin = 0, end = 0;
  };
  struct Result {
    int* idx; float* sqd; int cap; int count;
    bool better(float d, int i) const {
      if (count < cap) return true;
      return d < sqd[cap - 1] || (d == sqd[cap - 1] && i < idx[cap - 1]);
    }
    void add(float d, int i) {
      if (!better(d, i)) return;
      int pos = count < cap ? count : cap - 1;
      while (pos > 0 && (sqd[pos - 1] > d || (sqd[pos - 1] == d && idx[pos - 1] > i))) {
        sqd[pos] = sqd[pos - 1]; idx[pos] = idx[pos - 1]; pos--;
      }
      sqd[pos] = d; idx[pos] = i;
      if (count < cap) count++;
    }
  };
  int build(size_t b, size_t e) {
    const int id = (int)nodes_.size();
    nodes_.emplace_back();
    {
      Node& nd = nodes_[id];
      nd.begin = b; nd.end = e;
      for (int d = 0; d < 3; d++) { nd.lo[d] = DBL_MAX; nd.hi[d] = -DBL_MAX; }
      for (size_t i = b; i < e; i++)
        for (int d = 0; d < 3; d++) {
          const double v = cloud_->pt(perm_[i])[d];
          nd.lo[d] = std::min(nd.lo[d], v); nd.hi[d] = std::max(nd.hi[d], v);
        }
    }
    if (e - b <= 16) return id;
    int dim = 0;
    {
      const Node& nd = nodes_[id];
      double best = -1;
      for (int d = 0; d < 3; d++) if (nd.hi[d] - nd.lo[d] > best) { best = nd.hi[d] - nd.lo[d]; dim = d; }
      if (best <= 0.0) return id;  // all points identical: keep as a (large) leaf
    }
    const size_t mid = b + (e - b) / 2;
    std::nth_element(perm_.begin() + b, perm_.begin() + mid, perm_.begin() + e, [&](int a, int c) {
      const float va = cloud_->pt(a)[dim], vc = cloud_->pt(c)[dim];
      return va < vc || (va == vc && a < c);
    });
    const int l = build(b, mid);
    const int r = build(mid, e);
    nodes_[id].left = l; nodes_[id].right = r;
    return id;
  }
  static double box_lb(const Node& nd, const float* q) {
    double s = 0;
    for (int d = 0; d < 3; d++) {
      const double v = q[d];
      const double g = v < nd.lo[d] ? nd.lo[d] - v : (v > nd.hi[d] ? v - nd.hi[d] : 0.0);
      s += g * g;
    }
    return s * (1.0 - 1e-6);  // conservative against the fp32 rounding of the metric
  }
  void search(int id, const float* q, Result& r) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (size_t i = nd.begin; i < nd.end; i++) {
        const int pi = perm_[i];
        r.add(sqdist_f32(q, cloud_->pt(pi)), pi);
      }
      return;
    }
    const double ll = box_lb(nodes_[nd.left], q), lr = box_lb(nodes_[nd.right], q);
    const int first = ll <= lr ? nd.left : nd.right, second = ll <= lr ? nd.right : nd.left;
    const double lsecond = ll <= lr ? lr : ll;
    search(first, q, r);
    if (r.count < r.cap || lsecond <= (double)r.sqd[r.cap - 1]) search(second, q, r);
  }
  std::shared_ptr<const Cloud> cloud_;
  std::vector<int> perm_;
  std::vector<Node> nodes_;
};

#ifdef ORC_USE_REF_NANOFLANN
// The reference tree, instantiated as nanoflann_adaptor.h:100-102 does, leaf size 25 (:114).
class RefTree {
 public:
  using KD = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::SO3_Adaptor<float, Cloud>, Cloud, 3, int>;
  explicit RefTree(std::shared_ptr<const Cloud> c)
      : cloud_(std::move(c)), kd_(3, *cloud_, nanoflann::KDTreeSingleIndexAdaptorParams(25)) {
    // KDTreeSingleIndexAdaptor's ctor runs buildIndex() (nanoflann.h:1396-1398). The reference
    // constructs on an empty adaptor and calls buildIndex() once the cloud is attached
    // (nanoflann_adaptor.h:132-138); here the cloud is attached first, so the ctor's build is
    // that one build.
    params_.sorted = false;  // nanoflann_adaptor.h:113-117
  }
  int knn(const float* q, int k, int* idx, float* sqd) const {
    // nanoflann_adaptor.h:141-152
    nanoflann::KNNResultSet<float, int> rs(k);
    rs.init(idx, sqd);
    kd_.findNeighbors(rs, q, params_);
    return (int)rs.size();
  }
  const std::shared_ptr<const Cloud>& cloud() const { return cloud_; }

 private:
  std::shared_ptr<const Cloud> cloud_;
  KD kd_;
  nanoflann::SearchParams params_;
};
using Tree = RefTree;
static const char* kTreeKind = "reference-nanoflann";
#else
using Tree = PortTree;
static const char* kTreeKind = "port-kdtree";
#endif

// Matrix4d as the reference stores it: 16 doubles, column-major (Eigen default). Only the upper
// 3x3 block is ever non-zero on this path.
struct Mat4 {
  double m[16];
  double& operator()(int r, int c) { return m[4 * c + r]; }
  double operator()(int r, int c) const { return m[4 * c + r]; }
  static Mat4 zero() { Mat4 a; std::memset(a.m, 0, sizeof a.m); return a; }
};
static M3 block3(const Mat4& a) {
  M3 b;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) b(r, c) = a(r, c);
  return b;
}
static void set_block3(Mat4& a, const M3& b) {
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) a(r, c) = b(r, c);
}

enum RegMethod { REG_NONE = 0, REG_MIN_EIG = 1, REG_NORMALIZED_MIN_EIG = 2, REG_PLANE = 3, REG_FROBENIUS = 4 };  // nano_gicp.h:61

// lsq_registration.h:82-101
static M3 so3_exp(const double omega[3]) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  // Eigen::Quaterniond(w,x,y,z) is NOT normalised by the ctor; toRotationMatrix uses it as is.
  return quat_to_rot(real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]);
}

struct Gicp {
  // configuration (defaults: nano_gicp.cc:53-66, lsq_registration.cc:53-67)
  int num_threads = 1;
  int k_correspondences = 20;
  double corr_dist_threshold = (double)FLT_MAX;
  int reg_method = REG_PLANE;
  int max_iterations = 64;
  double rotation_epsilon = 2e-3;
  double transformation_epsilon = 5e-4;
  int lm_max_iterations = 10;
  double lm_init_lambda_factor = 1e-9;
  bool use_gauss_newton = false;

  std::shared_ptr<const Cloud> input, target;
  std::shared_ptr<const Tree> source_tree, target_tree;
  std::shared_ptr<std::vector<Mat4>> source_covs, target_covs;
  float source_density = 0, target_density = 0;
  int num_correspondences = 0;

  std::vector<Mat4> mahalanobis;
  std::vector<int> correspondences;
  std::vector<float> sq_distances;

  // LM state
  double lm_lambda = -1.0;
  double final_hessian[36];
  double final_error = 0.0;
  bool converged = false;
  int nr_iterations = 0;
  float final_transformation[16];  // column-major 4x4

  Gicp() {
#ifdef _OPENMP
    num_threads = omp_get_max_threads();
#endif
    for (int i = 0; i < 36; i++) final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;
  }

  // nano_gicp.cc:330-392
  bool calculate_covariances(const Cloud& cloud, const Tree& tree, std::vector<Mat4>& covs, float& density) {
    const int k = k_correspondences;
    covs.resize(cloud.n);
    float sum_k_sq_distances = 0.0f;
    const int n = (int)cloud.n;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8) reduction(+ : sum_k_sq_distances)
    for (int i = 0; i < n; i++) {
      std::vector<int> k_indices(k);
      std::vector<float> k_sq_distances(k);
      tree.knn(cloud.pt(i), k, k_indices.data(), k_sq_distances.data());

      const int normalization = ((k - 1) * (2 + k)) / 2;
      double acc = 0.0;  // std::accumulate(..., 0.0): double accumulator
      for (int j = 1; j < k; j++) acc += k_sq_distances[j];
      sum_k_sq_distances += acc / normalization;  // float += double  (the reduction variable is float)

      std::vector<double> nb(4 * (size_t)k);
      for (int j = 0; j < k; j++) {
        const float* p = cloud.pt(k_indices[j]);
        nb[4 * j + 0] = p[0]; nb[4 * j + 1] = p[1]; nb[4 * j + 2] = p[2]; nb[4 * j + 3] = 1.0;
      }
      double mean[4] = {0, 0, 0, 0};
      for (int r = 0; r < 4; r++) {
        double s = 0;
        for (int j = 0; j < k; j++) s += nb[4 * j + r];
        mean[r] = s / k;
      }
      for (int j = 0; j < k; j++) for (int r = 0; r < 4; r++) nb[4 * j + r] -= mean[r];
      Mat4 cov = Mat4::zero();
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double s = 0;
          for (int j = 0; j < k; j++) s += nb[4 * j + r] * nb[4 * j + c];
          cov(r, c) = s / k;
        }

      if (reg_method == REG_NONE) {
        covs[i] = cov;
      } else if (reg_method == REG_FROBENIUS) {
        const double lambda = 1e-3;
        M3 C = block3(cov);
        C(0, 0) += lambda; C(1, 1) += lambda; C(2, 2) += lambda;
        M3 C_inv = inverse(C);
        const double nrm = frobenius(C_inv);
        for (double& v : C_inv.m) v /= nrm;
        covs[i] = Mat4::zero();
        set_block3(covs[i], inverse(C_inv));
      } else {
        M3 U, V; double sv[3];
        svd3(block3(cov), U, sv, V);
        double values[3];
        switch (reg_method) {
          default:
          case REG_PLANE: values[0] = 1; values[1] = 1; values[2] = 1e-3; break;
          case REG_MIN_EIG:
            for (int a = 0; a < 3; a++) values[a] = std::max(sv[a], 1e-3);
            break;
          case REG_NORMALIZED_MIN_EIG: {
            const double mx = std::max(sv[0], std::max(sv[1], sv[2]));
            for (int a = 0; a < 3; a++) values[a] = std::max(sv[a] / mx, 1e-3);
          } break;
        }
        M3 UD = U;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) UD(r, c) = U(r, c) * values[c];
        covs[i] = Mat4::zero();
        set_block3(covs[i], mul(UD, transpose(V)));
      }
    }
    density = sum_k_sq_distances / cloud.n;
    return true;
  }

  // nano_gicp.cc:206-245
  void update_correspondences(const Iso3& trans) {
    float Rf[9], tf[3];
    for (int i = 0; i < 9; i++) Rf[i] = (float)trans.R.m[i];
    for (int i = 0; i < 3; i++) tf[i] = (float)trans.t[i];
    const int n = (int)input->n;
    correspondences.resize(n);
    sq_distances.resize(n);
    mahalanobis.resize(n);
    const double thr2 = corr_dist_threshold * corr_dist_threshold;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      const float* p = input->pt(i);
      float pt[3];
      // Isometry3f * Vector4f (w = 1), fp32, column-accumulated: ((r0*x + r1*y) + r2*z) + t*w
      for (int r = 0; r < 3; r++) pt[r] = ((Rf[3 * r + 0] * p[0] + Rf[3 * r + 1] * p[1]) + Rf[3 * r + 2] * p[2]) + tf[r] * 1.0f;
      int k_index = 0; float k_sq = 0.0f;
      target_tree->knn(pt, 1, &k_index, &k_sq);
      sq_distances[i] = k_sq;
      correspondences[i] = ((double)k_sq < thr2) ? k_index : -1;
      if (correspondences[i] < 0) continue;
      const M3 cov_A = block3((*source_covs)[i]);
      const M3 cov_B = block3((*target_covs)[correspondences[i]]);
      const M3 RCR = add(cov_B, mul(mul(trans.R, cov_A), transpose(trans.R)));
      Mat4 M = Mat4::zero();          // inverse of blockdiag(RCR,1) with (3,3) zeroed afterwards
      set_block3(M, inverse(RCR));
      mahalanobis[i] = M;
    }
    num_correspondences = (int)std::count_if(correspondences.begin(), correspondences.end(), [](int c) { return c > 0; });
  }

  // shared per-point residual: returns false if no correspondence
  inline bool residual(const Iso3& trans, int i, double q[3], double e[3]) const {
    const int ti = correspondences[i];
    if (ti < 0) return false;
    const float* a = input->pt(i);
    const float* b = target->pt(ti);
    for (int r = 0; r < 3; r++) {
      q[r] = trans.R(r, 0) * (double)a[0] + trans.R(r, 1) * (double)a[1] + trans.R(r, 2) * (double)a[2] + trans.t[r];
      e[r] = (double)b[r] - q[r];
    }
    return true;
  }

  // nano_gicp.cc:248-302. H row-major 6x6 (symmetric), b 6.
  double linearize(const Iso3& trans, double* H, double* b) {
    update_correspondences(trans);
    double sum_errors = 0.0;
    const int nt = std::max(1, num_threads);
    std::vector<std::array<double, 36>> Hs(nt);
    std::vector<std::array<double, 6>> bs(nt);
    for (int t = 0; t < nt; t++) { Hs[t].fill(0.0); bs[t].fill(0.0); }
    const int n = (int)input->n;
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      double q[3], e[3];
      if (!residual(trans, i, q, e)) continue;
      const M3 M = block3(mahalanobis[i]);
      double Me[3];
      for (int r = 0; r < 3; r++) Me[r] = M(r, 0) * e[0] + M(r, 1) * e[1] + M(r, 2) * e[2];
      sum_errors += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
      if (H == nullptr || b == nullptr) continue;
      // J = [ skew(q) | -I ]   (3x6; row 3 of the reference's 4x6 is zero)
      double J[3][6] = {{0, -q[2], q[1], -1, 0, 0}, {q[2], 0, -q[0], 0, -1, 0}, {-q[1], q[0], 0, 0, 0, -1}};
      double MJ[3][6];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 6; c++) MJ[r][c] = M(r, 0) * J[0][c] + M(r, 1) * J[1][c] + M(r, 2) * J[2][c];
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) Hs[tid][6 * r + c] += J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c] + J[2][r] * MJ[2][c];
        bs[tid][r] += J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2];
      }
    }
    if (H && b) {
      std::fill(H, H + 36, 0.0); std::fill(b, b + 6, 0.0);
      for (int t = 0; t < nt; t++) {
        for (int i = 0; i < 36; i++) H[i] += Hs[t][i];
        for (int i = 0; i < 6; i++) b[i] += bs[t][i];
      }
    }
    return sum_errors;
  }

  // nano_gicp.cc:305-326 (cached correspondences / mahalanobis; no re-association)
  double compute_error(const Iso3& trans) {
    double sum_errors = 0.0;
    const int n = (int)input->n;
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      double q[3], e[3];
      if (!residual(trans, i, q, e)) continue;
      const M3 M = block3(mahalanobis[i]);
      double Me[3];
      for (int r = 0; r < 3; r++) Me[r] = M(r, 0) * e[0] + M(r, 1) * e[1] + M(r, 2) * e[2];
      sum_errors += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
    }
    return sum_errors;
  }

  // lsq_registration.cc:137-146
  bool is_converged(const Iso3& delta) const {
    double rmax = 0, tmax = 0;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) rmax = std::max(rmax, 1.0 / rotation_epsilon * std::fabs(delta.R(r, c) - (r == c ? 1.0 : 0.0)));
      tmax = std::max(tmax, 1.0 / transformation_epsilon * std::fabs(delta.t[r]));
    }
    return std::max(rmax, tmax) < 1;
  }

  static Iso3 delta_from(const double d[6]) {
    Iso3 delta;
    delta.R = so3_exp(d);
    delta.t[0] = d[3]; delta.t[1] = d[4]; delta.t[2] = d[5];
    return delta;
  }

  // lsq_registration.cc:161-178
  bool step_gn(Iso3& x0, Iso3& delta) {
    double H[36], b[6], nb[6], d[6];
    const double y0 = linearize(x0, H, b);
    for (int i = 0; i < 6; i++) nb[i] = -b[i];
    ldlt6_solve(H, nb, d);
    delta = delta_from(d);
    x0 = compose(delta, x0);
    std::memcpy(final_hessian, H, sizeof H);
    final_error = y0;
    return true;
  }

  // lsq_registration.cc:181-229
  bool step_lm(Iso3& x0, Iso3& delta) {
    double H[36], b[6];
    const double y0 = linearize(x0, H, b);
    if (lm_lambda < 0.0) {
      double mx = 0;
      for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[7 * i]));
      lm_lambda = lm_init_lambda_factor * mx;
    }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations; i++) {
      double A[36], nb[6], d[6];
      std::memcpy(A, H, sizeof A);
      for (int j = 0; j < 6; j++) { A[7 * j] += lm_lambda; nb[j] = -b[j]; }
      ldlt6_solve(A, nb, d);
      delta = delta_from(d);
      const Iso3 xi = compose(delta, x0);
      const double yi = compute_error(xi);
      double den = 0;
      for (int j = 0; j < 6; j++) den += d[j] * (lm_lambda * d[j] - b[j]);
      const double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_converged(delta)) return true;
        lm_lambda = nu * lm_lambda;
        nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(final_hessian, H, sizeof H);
      final_error = yi;
      return true;
    }
    return false;
  }

  // nano_gicp.cc:194-203 + lsq_registration.cc:108-134 (+ PCL align(): guess handed through)
  // guess / result: column-major float 4x4. Returns 0, or 1 if the LM inner loop was exhausted.
  int align(const float guess[16]) {
    if (!source_covs || source_covs->size() != input->n) calc_covs(0);
    if (!target_covs || target_covs->size() != target->n) calc_covs(1);
    Iso3 x0;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) x0.R(r, c) = (double)guess[4 * c + r];
      x0.t[r] = (double)guess[12 + r];
    }
    lm_lambda = -1.0;
    converged = false;
    nr_iterations = 0;
    int lm_failed = 0;
    for (int i = 0; i < max_iterations && !converged; i++) {
      nr_iterations = i;
      Iso3 delta;
      const bool ok = use_gauss_newton ? step_gn(x0, delta) : step_lm(x0, delta);
      if (!ok) { lm_failed = 1; break; }  // reference prints "lm not converged!!" and breaks
      converged = is_converged(delta);
    }
    std::memset(final_transformation, 0, sizeof final_transformation);
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) final_transformation[4 * c + r] = (float)x0.R(r, c);
      final_transformation[12 + r] = (float)x0.t[r];
    }
    final_transformation[15] = 1.0f;
    return lm_failed;
  }

  bool calc_covs(int which) {
    auto covs = std::make_shared<std::vector<Mat4>>();
    float density = 0;
    bool ret;
    if (which == 0) {
      ret = calculate_covariances(*input, *source_tree, *covs, density);
      source_covs = covs; source_density = density;
    } else {
      ret = calculate_covariances(*target, *target_tree, *covs, density);
      target_covs = covs; target_density = density;
    }
    return ret;
  }
};

static std::shared_ptr<const Cloud> make_cloud(const float* pts, size_t n, size_t stride_floats) {
  auto c = std::make_shared<Cloud>();
  c->n = n;
  c->xyz.resize(3 * n);
  for (size_t i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) c->xyz[3 * i + d] = pts[i * stride_floats + d];
  return c;
}

static Iso3 iso_from_colmajor(const double T[16]) {
  Iso3 x;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) x.R(r, c) = T[4 * c + r];
    x.t[r] = T[12 + r];
  }
  return x;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------
// C ABI (ctypes-friendly). All 4x4 matrices are column-major (Eigen's default storage).
// ------------------------------------------------------------------------------------------
using namespace orc;

struct orc_tree { std::shared_ptr<const Tree> t; };

extern "C" {

const char* orc_tree_kind() { return kTreeKind; }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

orc_tree* orc_tree_build(const float* pts, size_t n, size_t stride_floats) {
  auto* t = new orc_tree;
  t->t = std::make_shared<const Tree>(make_cloud(pts, n, stride_floats));
  return t;
}
void orc_tree_free(orc_tree* t) { delete t; }

// k-NN of nq queries (float stride qstride). canonical != 0 re-orders every result row to
// ascending (distance, index) — the documented tie-break of the new build.
// Rows with fewer than k hits are padded with idx -1 / dist +inf. Returns 0.
int orc_knn(const orc_tree* t, const float* q, size_t nq, size_t qstride, int k, int* idx, float* sqd, int canonical, int num_threads) {
  if (num_threads <= 0) num_threads = orc_max_threads();
  const long long n = (long long)nq;
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (long long i = 0; i < n; i++) {
    int* ri = idx + (size_t)i * k;
    float* rd = sqd + (size_t)i * k;
    const int cnt = t->t->knn(q + (size_t)i * qstride, k, ri, rd);
    for (int j = cnt; j < k; j++) { ri[j] = -1; rd[j] = INFINITY; }
    if (canonical) {
      std::vector<std::pair<float, int>> row(cnt);
      for (int j = 0; j < cnt; j++) row[j] = {rd[j], ri[j]};
      std::sort(row.begin(), row.end());
      for (int j = 0; j < cnt; j++) { rd[j] = row[j].first; ri[j] = row[j].second; }
    }
  }
  return 0;
}

Gicp* orc_gicp_create() { return new Gicp; }
void orc_gicp_destroy(Gicp* g) { delete g; }

void orc_gicp_set_params(Gicp* g, int num_threads, int k, double max_corr_dist, int reg_method, int max_iterations,
                         double rot_eps, double trans_eps, double lm_init_lambda_factor, int lm_max_iterations, int gauss_newton) {
  g->num_threads = num_threads > 0 ? num_threads : orc_max_threads();
  g->k_correspondences = k;
  g->corr_dist_threshold = max_corr_dist;
  g->reg_method = reg_method;
  g->max_iterations = max_iterations;
  g->rotation_epsilon = rot_eps;
  g->transformation_epsilon = trans_eps;
  g->lm_init_lambda_factor = lm_init_lambda_factor;
  g->lm_max_iterations = lm_max_iterations;
  g->use_gauss_newton = gauss_newton != 0;
}

// setInputSource / setInputTarget (nano_gicp.cc:135-161): store cloud, build tree, drop covariances.
void orc_gicp_set_cloud(Gicp* g, int which, const float* pts, size_t n, size_t stride_floats) {
  auto c = make_cloud(pts, n, stride_floats);
  auto t = std::make_shared<const Tree>(c);
  if (which == 0) { g->input = c; g->source_tree = t; g->source_covs.reset(); }
  else { g->target = c; g->target_tree = t; g->target_covs.reset(); }
}
int orc_gicp_calc_covs(Gicp* g, int which, float* density) {
  const bool r = g->calc_covs(which);
  if (density) *density = which == 0 ? g->source_density : g->target_density;
  return r ? 0 : 1;
}
size_t orc_gicp_get_covs(const Gicp* g, int which, double* out) {
  const auto& c = which == 0 ? g->source_covs : g->target_covs;
  if (!c) return 0;
  if (out) std::memcpy(out, c->data(), c->size() * sizeof(Mat4));
  return c->size();
}
void orc_gicp_set_covs(Gicp* g, int which, const double* in, size_t n) {
  auto c = std::make_shared<std::vector<Mat4>>(n);
  std::memcpy(c->data(), in, n * sizeof(Mat4));
  (which == 0 ? g->source_covs : g->target_covs) = c;
}
void orc_gicp_update_correspondences(Gicp* g, const double T[16], int* corr, float* sqd, double* mahal) {
  g->update_correspondences(iso_from_colmajor(T));
  const size_t n = g->input->n;
  if (corr) std::memcpy(corr, g->correspondences.data(), n * sizeof(int));
  if (sqd) std::memcpy(sqd, g->sq_distances.data(), n * sizeof(float));
  if (mahal)
    for (size_t i = 0; i < n; i++) {
      if (g->correspondences[i] < 0) std::memset(mahal + 16 * i, 0, 16 * sizeof(double));
      else std::memcpy(mahal + 16 * i, g->mahalanobis[i].m, 16 * sizeof(double));
    }
}
int orc_gicp_num_correspondences(const Gicp* g) { return g->num_correspondences; }
double orc_gicp_linearize(Gicp* g, const double T[16], double* H, double* b) { return g->linearize(iso_from_colmajor(T), H, b); }
double orc_gicp_compute_error(Gicp* g, const double T[16]) { return g->compute_error(iso_from_colmajor(T)); }
int orc_gicp_align(Gicp* g, const float guess[16], float T_out[16], int* nr_iterations, int* converged, double* H_final, double* final_err) {
  const int rc = g->align(guess);
  if (T_out) std::memcpy(T_out, g->final_transformation, sizeof g->final_transformation);
  if (nr_iterations) *nr_iterations = g->nr_iterations;
  if (converged) *converged = g->converged ? 1 : 0;
  if (H_final) std::memcpy(H_final, g->final_hessian, sizeof g->final_hessian);
  if (final_err) *final_err = g->final_error;
  return rc;
}

// ---- scan pre-filters ------------------------------------------------------------------------
// PCL (>= 1.10.0, apt libpcl-dev; src/dlio/README.md:31) is a third-party dependency that is NOT part of the reference
// tree, so these two functions restate PCL 1.10's published algorithms as DLIO calls them (xyz only):
//   pcl::CropBox<PointT>::applyFilter   (filters/impl/crop_box.hpp)    reference call sites odom.cc:114-116, :500-502
//   pcl::VoxelGrid<PointT>::applyFilter (filters/impl/voxel_grid.hpp)  reference call sites odom.cc:118, :575-584
// Parity with a real PCL is unpinned (no PCL in this image, no fixtures in the reference). One deliberate pin: PCL sorts
// the (voxel index, point index) pairs with std::sort, whose order inside a voxel is unspecified, so PCL's own fp32
// centroid sums are defined only up to rounding; here the order is ascending point index (std::stable_sort).

// keeps a finite point iff (inside the box) != negative; returns the number kept; out = n x 3 floats capacity
size_t orc_crop_box(const float* pts, size_t n, size_t stride_floats, const float mn[3], const float mx[3], int negative, float* out) {
  size_t m = 0;
  for (size_t i = 0; i < n; i++) {
    const float* p = pts + i * stride_floats;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;   // crop_box.hpp: "Check if the point is invalid"
    const bool outside = (p[0] < mn[0] || p[1] < mn[1] || p[2] < mn[2]) || (p[0] > mx[0] || p[1] > mx[1] || p[2] > mx[2]);
    if (outside ? (negative != 0) : (negative == 0)) {
      out[3 * m] = p[0]; out[3 * m + 1] = p[1]; out[3 * m + 2] = p[2];
      m++;
    }
  }
  return m;
}

// out = n x 3 floats capacity, out_voxel / out_count (optional) = voxel index and population of every output point
size_t orc_voxel_grid(const float* pts, size_t n, size_t stride_floats, const float leaf[3], float* out, int* out_voxel, int* out_count) {
  float inv[3], mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int a = 0; a < 3; a++) inv[a] = 1.0f / leaf[a];                                     // Array4f::Ones() / leaf_size_
  for (size_t i = 0; i < n; i++) {                                                          // getMinMax3D, non-dense cloud
    const float* p = pts + i * stride_floats;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
  }
  long long cells = 1;
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++) {
    cells *= static_cast<long long>((mx[a] - mn[a]) * inv[a]) + 1;
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv[a]));
    div_b[a] = static_cast<int>(std::floor(mx[a] * inv[a])) - min_b[a] + 1;
  }
  if (cells > static_cast<long long>(INT32_MAX)) {                                          // "Leaf size is too small": output = input
    for (size_t i = 0; i < n; i++) for (int a = 0; a < 3; a++) out[3 * i + a] = pts[i * stride_floats + a];
    return n;
  }
  std::vector<std::pair<unsigned int, unsigned int>> iv;                                    // (voxel idx, point index)
  iv.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const float* p = pts + i * stride_floats;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    int ijk[3];
    for (int a = 0; a < 3; a++) ijk[a] = static_cast<int>(std::floor(p[a] * inv[a]) - static_cast<float>(min_b[a]));
    iv.emplace_back(static_cast<unsigned int>(ijk[0] + ijk[1] * div_b[0] + ijk[2] * div_b[0] * div_b[1]), static_cast<unsigned int>(i));
  }
  std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned int, unsigned int>& a, const std::pair<unsigned int, unsigned int>& b) { return a.first < b.first; });
  size_t m = 0;
  for (size_t j = 0; j < iv.size();) {
    float s[3] = {0.f, 0.f, 0.f};                                                           // AccumulatorXYZ: Eigen::Vector3f xyz
    size_t e = j;
    for (; e < iv.size() && iv[e].first == iv[j].first; e++)
      for (int a = 0; a < 3; a++) s[a] += pts[iv[e].second * stride_floats + a];
    const float cnt = static_cast<float>(e - j);
    for (int a = 0; a < 3; a++) out[3 * m + a] = s[a] / cnt;
    if (out_voxel) out_voxel[m] = static_cast<int>(iv[j].first);
    if (out_count) out_count[m] = static_cast<int>(e - j);
    m++;
    j = e;
  }
  return m;
}

// ---- deskew: dlio::OdomNode::deskewPointcloud, odom.cc:588-706 ------------------------------------
// ingest: drop non-finite points (:496-498), optional CropBox (:500-502), sort by the per-point time stamp (:634-635;
// stable here, std::partial_sort_copy leaves ties unspecified), unique stamps (:638-650). Outputs: xyz in time order,
// group (index of its unique stamp) per point, unique stamps as raw field values. Returns the number of points kept.
size_t orc_scan_ingest(const unsigned char* recs, size_t n, size_t stride_bytes, size_t time_off, int time_type, const float* mn, const float* mx,
                       int negative, float* out_xyz, int* out_group, double* unique_stamps, size_t* n_unique) {
  struct Item { double stamp; size_t i; };
  std::vector<Item> items;
  items.reserve(n);
  for (size_t i = 0; i < n; i++) {
    const unsigned char* rec = recs + i * stride_bytes;
    const float* p = reinterpret_cast<const float*>(rec);
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    if (mn && mx) {
      const bool outside = (p[0] < mn[0] || p[1] < mn[1] || p[2] < mn[2]) || (p[0] > mx[0] || p[1] > mx[1] || p[2] > mx[2]);
      if (!(outside ? (negative != 0) : (negative == 0))) continue;
    }
    double st;
    if (time_type == 0) { std::uint32_t t; std::memcpy(&t, rec + time_off, 4); st = static_cast<double>(t); }
    else if (time_type == 1) { float t; std::memcpy(&t, rec + time_off, 4); st = static_cast<double>(t); }
    else { std::memcpy(&st, rec + time_off, 8); }
    items.push_back({st, i});
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.stamp < b.stamp; });
  size_t groups = 0;
  for (size_t j = 0; j < items.size(); j++) {
    if (j == 0 || items[j].stamp != items[j - 1].stamp) unique_stamps[groups++] = items[j].stamp;
    const float* p = reinterpret_cast<const float*>(recs + items[j].i * stride_bytes);
    out_xyz[3 * j] = p[0]; out_xyz[3 * j + 1] = p[1]; out_xyz[3 * j + 2] = p[2];
    out_group[j] = static_cast<int>(groups - 1);
  }
  *n_unique = groups;
  return items.size();
}
// pt = T_group * pt with w = 1 (odom.cc:690-701), fp32: ((m0 x + m1 y) + m2 z) + m3; frames16 column-major; n_frames == 1
// applies the one matrix to every point (pcl::transformPointCloud, :659 / :681)
void orc_scan_deskew(const float* xyz, const int* group, size_t n, const float* frames16, size_t n_frames, float* out) {
  for (size_t j = 0; j < n; j++) {
    const float* T = frames16 + (n_frames == 1 ? 0 : static_cast<size_t>(group[j])) * 16;
    const float x = xyz[3 * j], y = xyz[3 * j + 1], z = xyz[3 * j + 2];
    for (int r = 0; r < 3; r++) {
      volatile float a = T[r] * x, b = T[4 + r] * y, c = T[8 + r] * z;   // volatile: no FMA contraction
      volatile float s1 = a + b;
      volatile float s2 = s1 + c;
      out[3 * j + r] = s2 + T[12 + r];
    }
  }
}

}  // extern "C"
