// TEST-ONLY shim of pcl::PointCloud<PointT> (PCL is absent from this image). Not shipped.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  const PointT& at(size_t i) const { return points.at(i); }
};
}  // namespace pcl
