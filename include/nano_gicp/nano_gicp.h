// Drop-in replacement of the reference's include/nano_gicp/nano_gicp.h (:63-150) and
// src/nano_gicp/nano_gicp.cc: the nano_gicp::NanoGICP<PointSource,PointTarget> surface DLIO's odom
// node links against (call sites: src/dlio/odom.cc:89-107,715,721-722,829,992-1008,1428,1592,1619,
// 1737-1738) is unchanged; every member forwards to the C ABI in ../ngicp_b200.h, which runs the
// hand-written sm_100a kernels. Header-only (no explicit instantiation needed).
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/registration/registration.h>

#include "../ngicp_b200.h"
#include "lsq_registration.h"
#include "nanoflann_adaptor.h"

namespace nano_gicp {

typedef std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> CovarianceList;

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };

template <typename PointSource, typename PointTarget>
class NanoGICP : public LsqRegistration<PointSource, PointTarget> {
 public:
  using Scalar = float;
  using Matrix4 = typename pcl::Registration<PointSource, PointTarget, Scalar>::Matrix4;
  using PointCloudSource = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

 protected:
  using pcl::Registration<PointSource, PointTarget, Scalar>::reg_name_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::input_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::target_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::final_transformation_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::converged_;
  using pcl::Registration<PointSource, PointTarget, Scalar>::nr_iterations_;
  using LsqRegistration<PointSource, PointTarget>::params_;
  using LsqRegistration<PointSource, PointTarget>::final_hessian_;
  using LsqRegistration<PointSource, PointTarget>::final_error_;

 public:
  explicit NanoGICP(int device = 0) {
    reg_name_ = "NanoGICP";
    if (ngicp_create(device, &h_) != NGICP_OK) throw std::runtime_error(ngicp_last_error(nullptr));  // no CPU fallback
    source_density_ = target_density_ = 0.f;
    num_correspondences = 0;
  }
  virtual ~NanoGICP() override { if (h_) ngicp_destroy(h_); }
  NanoGICP(const NanoGICP&) = delete;
  NanoGICP& operator=(const NanoGICP&) = delete;

  void setNumThreads(int) {}  // OpenMP team size of the reference (nano_gicp.cc:71-79): meaningless on the GPU
  void setCorrespondenceRandomness(int k) { params_.k_correspondences = k; }
  void setMaxCorrespondenceDistance(double corr) { params_.max_corr_dist = corr; }
  void setRegularizationMethod(RegularizationMethod m) { params_.regularization = static_cast<int>(m); }

  virtual void swapSourceAndTarget() override {  // nano_gicp.cc:97-104
    input_.swap(target_);
    source_kdtree_.swap(target_kdtree_);
    source_covs_.swap(target_covs_);
    check(ngicp_swap_source_and_target(h_));
    std::swap(attached_[0], attached_[1]);
    uploaded_covs_[0].swap(uploaded_covs_[1]);
    std::swap(device_covs_[0], device_covs_[1]);
  }
  virtual void clearSource() override { input_.reset(); source_covs_.reset(); device_covs_[0] = false; check(ngicp_clear(h_, NGICP_SOURCE)); attached_[0] = nullptr; }  // nano_gicp.cc:107-110
  virtual void clearTarget() override { target_.reset(); target_covs_.reset(); device_covs_[1] = false; check(ngicp_clear(h_, NGICP_TARGET)); attached_[1] = nullptr; }  // nano_gicp.cc:113-116

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) override {   // nano_gicp.cc:135-147
    if (input_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputSource(cloud);
    source_kdtree_ = build(NGICP_SOURCE, cloud);
    source_covs_.reset();
    device_covs_[NGICP_SOURCE] = false;
  }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) override {   // nano_gicp.cc:150-161
    if (target_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputTarget(cloud);
    target_kdtree_ = build(NGICP_TARGET, cloud);
    target_covs_.reset();
    device_covs_[NGICP_TARGET] = false;
  }
  // Additive: pcl::CropBox -> pcl::VoxelGrid -> setInputSource in one call that keeps the scan in
  // HBM (replaces odom.cc:501-502, :579-580 and :721). crop_size <= 0: no crop (DLIO's crop is negative: it removes the
  // box around the sensor); leaf <= 0: no voxel grid. The filtered cloud (xyz filled in, other fields default) becomes
  // input_ and is returned, so the caller can still publish it.
  PointCloudSourceConstPtr setInputSourceFiltered(const PointCloudSourceConstPtr& raw, float crop_size, float leaf) {
    const float mn[3] = {-crop_size, -crop_size, -crop_size}, mx[3] = {crop_size, crop_size, crop_size}, lf[3] = {leaf, leaf, leaf};
    std::vector<float> xyz(raw->points.size() * 3);
    size_t n_out = 0;
    check(ngicp_filter_scan(h_, raw->points.data(), raw->points.size(), sizeof(PointSource), crop_size > 0 ? mn : nullptr,
                            crop_size > 0 ? mx : nullptr, 1, leaf > 0 ? lf : nullptr, NGICP_SOURCE, xyz.data(), &n_out));
    PointCloudSourcePtr cloud(new PointCloudSource);   // PointCloud::Ptr is boost::shared_ptr up to PCL 1.10, std::shared_ptr from 1.11
    cloud->points.resize(n_out);
    for (size_t i = 0; i < n_out; i++) { cloud->points[i].x = xyz[3 * i]; cloud->points[i].y = xyz[3 * i + 1]; cloud->points[i].z = xyz[3 * i + 2]; }
    cloud->width = static_cast<std::uint32_t>(n_out); cloud->height = 1; cloud->is_dense = true;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputSource(cloud);
    auto tree = std::make_shared<nanoflann::KdTreeFLANN<PointSource>>();
    tree->adopt(cloud, ngicp_get_index(h_, NGICP_SOURCE));
    attached_[NGICP_SOURCE] = tree->index();
    uploaded_covs_[NGICP_SOURCE].reset();
    source_kdtree_ = tree;
    source_covs_.reset();
    device_covs_[NGICP_SOURCE] = false;
    return cloud;
  }
  ngicp_handle* handle() const { return h_; }
  // Additive: align(output) fills `output` with the transformed source as PCL does; a caller that never reads it
  // (DLIO, odom.cc:1004-1005) can skip that host pass over the cloud.
  void setComputeOutputCloud(bool on) { compute_output_ = on; }
  virtual void setSourceCovariances(const std::shared_ptr<const CovarianceList>& covs) { source_covs_ = covs; device_covs_[0] = false; }  // :164-166
  virtual void setTargetCovariances(const std::shared_ptr<const CovarianceList>& covs) { target_covs_ = covs; device_covs_[1] = false; }  // :169-171
  virtual void registerInputSource(const PointCloudSourceConstPtr& cloud) {   // nano_gicp.cc:119-124
    if (input_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputSource(cloud);
  }
  virtual void registerInputTarget(const PointCloudTargetConstPtr& cloud) {   // nano_gicp.cc:127-132
    if (target_ == cloud) return;
    pcl::Registration<PointSource, PointTarget, Scalar>::setInputTarget(cloud);
  }

  virtual bool calculateSourceCovariances() { return calculate(NGICP_SOURCE); }   // nano_gicp.cc:174-181
  virtual bool calculateTargetCovariances() { return calculate(NGICP_TARGET); }   // nano_gicp.cc:184-191

  // Covariances computed on the device are copied to a host CovarianceList only when somebody asks
  // (DLIO does for keyframes: odom.cc:715,1592), not on every scan.
  std::shared_ptr<const CovarianceList> getSourceCovariances() const { materialize(NGICP_SOURCE); return source_covs_; }
  std::shared_ptr<const CovarianceList> getTargetCovariances() const { materialize(NGICP_TARGET); return target_covs_; }

  virtual void update_correspondences(const Eigen::Isometry3d& trans) {          // nano_gicp.cc:206-245
    sync();
    check(ngicp_update_correspondences(h_, trans.matrix().data(), nullptr, nullptr, nullptr, &num_correspondences));
  }

 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {   // nano_gicp.cc:194-203
    sync();
    int nr = 0, conv = 0;
    const int rc = ngicp_align(h_, guess.data(), final_transformation_.data(), &nr, &conv, final_hessian_.data(), &final_error_);
    if (rc != NGICP_OK && rc != NGICP_ERR_LM_NOT_CONVERGED) throw std::runtime_error(ngicp_last_error(h_));
    nr_iterations_ = nr;
    converged_ = conv != 0;
    // covariances computed lazily inside align (:195-200) live on the device until asked for
    for (int which = 0; which < 2; which++) {
      const auto& covs = which == NGICP_SOURCE ? source_covs_ : target_covs_;
      if (!covs && !device_covs_[which]) device_covs_[which] = ngicp_has_covariances(h_, which, nullptr) != 0;
    }
    // pcl::transformPointCloud(*input_, output, final_transformation_)  (lsq_registration.cc:133). The cloud is the
    // caller's host memory and every other field of a point is copied as it is, so this stays a host loop exactly like the
    // reference's (fp32, ((m0 x + m1 y) + m2 z) + m3 per row); DLIO discards the result (odom.cc:1004-1005) and may switch
    // it off with setComputeOutputCloud(false).
    if (compute_output_) {
      // pcl::Registration::align has already copied *input_ into `output` (every field); only xyz is rewritten, from input_
      if (output.points.size() != input_->points.size()) output = *input_;
      const Matrix4& M = final_transformation_;
      const float m00 = M(0, 0), m01 = M(0, 1), m02 = M(0, 2), m03 = M(0, 3), m10 = M(1, 0), m11 = M(1, 1), m12 = M(1, 2), m13 = M(1, 3);
      const float m20 = M(2, 0), m21 = M(2, 1), m22 = M(2, 2), m23 = M(2, 3);
      const size_t n = input_->points.size();
      const PointSource* in = input_->points.data();
      PointSource* out = output.points.data();
      for (size_t i = 0; i < n; i++) {
        const float x = in[i].x, y = in[i].y, z = in[i].z;
        out[i].x = ((m00 * x + m01 * y) + m02 * z) + m03;
        out[i].y = ((m10 * x + m11 * y) + m12 * z) + m13;
        out[i].z = ((m20 * x + m21 * y) + m22 * z) + m23;
      }
    }
  }

 public:
  std::shared_ptr<const nanoflann::KdTreeFLANN<PointSource>> source_kdtree_;
  std::shared_ptr<const nanoflann::KdTreeFLANN<PointTarget>> target_kdtree_;
  mutable std::shared_ptr<const CovarianceList> source_covs_;
  mutable std::shared_ptr<const CovarianceList> target_covs_;
  float source_density_;
  float target_density_;
  int num_correspondences;

 private:
  void check(int rc) const { if (rc != NGICP_OK) throw std::runtime_error(ngicp_last_error(h_)); }

  template <typename CloudPtr>
  std::shared_ptr<const nanoflann::KdTreeFLANN<PointSource>> build(int which, const CloudPtr& cloud) {
    check(ngicp_set_input(h_, which, cloud->points.data(), cloud->points.size(), sizeof(PointSource)));
    auto tree = std::make_shared<nanoflann::KdTreeFLANN<PointSource>>();
    tree->adopt(cloud, ngicp_get_index(h_, which));
    attached_[which] = tree->index();
    uploaded_covs_[which].reset();
    return tree;
  }
  bool calculate(int which) {
    sync_tree(which);
    check(ngicp_set_params(h_, &params_));
    float density = 0.f;
    check(ngicp_compute_covariances(h_, which, &density));
    (which == NGICP_SOURCE ? source_density_ : target_density_) = density;
    (which == NGICP_SOURCE ? source_covs_ : target_covs_).reset();
    device_covs_[which] = true;
    return true;  // the reference always returns true (nano_gicp.cc:391)
  }
  void materialize(int which) const {
    if (!device_covs_[which]) return;
    size_t n = 0;
    if (!ngicp_has_covariances(h_, which, &n) || n == 0) { device_covs_[which] = false; return; }
    auto covs = std::make_shared<CovarianceList>(n);
    check(ngicp_get_covariances(h_, which, (*covs)[0].data(), n));
    (which == NGICP_SOURCE ? source_covs_ : target_covs_) = covs;
    uploaded_covs_[which] = covs;
    device_covs_[which] = false;
  }
  // DLIO assigns the public members directly (target_kdtree_ = submap_kdtree, odom.cc:995;
  // setTargetCovariances(submap_normals), :998): push whatever changed to the device before use.
  void sync_tree(int which) {
    const auto& tree = which == NGICP_SOURCE ? source_kdtree_ : target_kdtree_;
    ngicp_index* want = tree ? tree->index() : nullptr;
    if (want != attached_[which]) { check(ngicp_attach_index(h_, which, want)); attached_[which] = want; uploaded_covs_[which].reset(); }
  }
  void sync() {
    check(ngicp_set_params(h_, &params_));
    for (int which = 0; which < 2; which++) {
      sync_tree(which);
      const auto& covs = which == NGICP_SOURCE ? source_covs_ : target_covs_;
      // uploaded_covs_ keeps the list it uploaded alive, so an address can never be reused by a different list (no ABA).
      // The device stores covariances as 6 x fp32 (the reference keeps fp64 Matrix4d): user-supplied lists are rounded.
      if (covs && covs != uploaded_covs_[which]) {
        check(ngicp_set_covariances(h_, which, (*covs)[0].data(), covs->size()));
        uploaded_covs_[which] = covs;
      }
    }
  }

  ngicp_handle* h_ = nullptr;
  ngicp_index* attached_[2] = {nullptr, nullptr};
  mutable std::shared_ptr<const CovarianceList> uploaded_covs_[2];
  bool compute_output_ = true;
  mutable bool device_covs_[2] = {false, false};   // covariances valid on the device, host list not materialised yet
};

}  // namespace nano_gicp
