// TEST-ONLY shim of pcl::Registration<Source,Target,Scalar> with the members and the align() sequence
// SURVEY.md Appendix C recalls from PCL 1.10. Not shipped.
#pragma once
#include <string>
#include <Eigen/Core>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    output = *input_;
    computeTransformation(output, guess);
  }
 protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_;
  bool converged_ = false;
  int nr_iterations_ = 0;
};
}  // namespace pcl
