// TEST-ONLY shim of pcl::PointCloud<PointT> (PCL is absent from this image). Not shipped.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <pcl/pcl_config.h>
#if PCL_VERSION_COMPARE(<, 1, 11, 0)
#include <boost/shared_ptr.hpp>
#endif
namespace pcl {
#if PCL_VERSION_COMPARE(<, 1, 11, 0)
template <typename T> using shared_ptr = boost::shared_ptr<T>;   // PCL <= 1.10
#else
template <typename T> using shared_ptr = std::shared_ptr<T>;     // PCL >= 1.11
#endif
template <typename PointT>
struct PointCloud {
  using Ptr = pcl::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = pcl::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  const PointT& at(size_t i) const { return points.at(i); }
};
}  // namespace pcl
