// Drop-in replacement of the reference's include/nano_gicp/lsq_registration.h (:103-168).
// Same class, same setters and getters; the Gauss-Newton / Levenberg-Marquardt loop itself
// (src/nano_gicp/lsq_registration.cc:108-229) runs inside ngicp_align — 6x6 solve on the host,
// linearize / compute_error as CUDA kernels — so this class only carries the configuration.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>

#include <pcl/point_cloud.h>
#include <pcl/registration/registration.h>

#include "../ngicp_b200.h"

namespace nano_gicp {

enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };

template <typename PointSource, typename PointTarget>
class LsqRegistration : public pcl::Registration<PointSource, PointTarget, float> {
 public:
  using Scalar = float;
  using Matrix4 = typename pcl::Registration<PointSource, PointTarget, Scalar>::Matrix4;
  using PointCloudSource = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudSource;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename pcl::Registration<PointSource, PointTarget, Scalar>::PointCloudTarget;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  LsqRegistration() {
    this->reg_name_ = "LsqRegistration";
    ngicp_default_params(&params_);   // lsq_registration.cc:53-67 defaults
    final_hessian_.setIdentity();
    final_error_ = 0.;
  }
  virtual ~LsqRegistration() {}

  void setRotationEpsilon(double eps) { params_.rotation_epsilon = eps; }
  void setTransformationEpsilon(double eps) { params_.transformation_epsilon = eps; }
  void setMaximumIterations(int iter) { params_.max_iterations = iter; }
  void setInitialLambdaFactor(double f) { params_.lm_init_lambda_factor = f; }
  void setDebugPrint(bool) {}   // the reference's LM table (lsq_registration.cc:203-209) is not reproduced

  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }
  double getFinalError() const { return final_error_; }

  virtual void swapSourceAndTarget() {}
  virtual void clearSource() {}
  virtual void clearTarget() {}

 protected:
  ngicp_params params_;
  Eigen::Matrix<double, 6, 6> final_hessian_;
  double final_error_;
};

}  // namespace nano_gicp
