// Drop-in replacement of the reference's include/nano_gicp/nanoflann_adaptor.h (:57-193):
// nanoflann::KdTreeFLANN<PointT> keeps its name and its setInputCloud / nearestKSearch / getInputCloud
// members (DLIO declares std::shared_ptr<const nanoflann::KdTreeFLANN<PointType>> submap_kdtree,
// include/dlio/odom.h:164), but the "tree" is the B200 voxel-hash index behind the C ABI.
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include <pcl/point_cloud.h>

#include "../ngicp_b200.h"

namespace nanoflann {

template <typename PointT>
class KdTreeFLANN {
 public:
  typedef typename pcl::PointCloud<PointT> PointCloud;
  typedef typename pcl::PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  typedef std::shared_ptr<std::vector<int>> IndicesPtr;
  typedef std::shared_ptr<const std::vector<int>> IndicesConstPtr;

  explicit KdTreeFLANN(bool /*sorted*/ = false, int device = 0) : device_(device) {}
  ~KdTreeFLANN() {
    if (index_) ngicp_index_release(index_);
    if (handle_) ngicp_destroy(handle_);
  }
  KdTreeFLANN(const KdTreeFLANN&) = delete;
  KdTreeFLANN& operator=(const KdTreeFLANN&) = delete;

  void setEpsilon(float) {}          // the search is exact (reference default eps = 0)
  void setSortedResults(bool) {}     // rows are always sorted by (distance, index)

  // nanoflann_adaptor.h:132-138 -> ngicp_index_build
  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) {
    if (indices) throw std::runtime_error("KdTreeFLANN(b200): index subsets are not supported");
    cloud_ = cloud;
    if (index_) { ngicp_index_release(index_); index_ = nullptr; }
    if (!cloud || cloud->points.empty()) return;
    ensure_handle();
    if (ngicp_index_build(handle_, cloud->points.data(), cloud->points.size(), sizeof(PointT), &index_) != NGICP_OK)
      throw std::runtime_error(ngicp_last_error(handle_));
  }
  // adopt an index built elsewhere (NanoGICP::setInputSource/Target build through their own handle)
  void adopt(const PointCloudConstPtr& cloud, ngicp_index* idx) {
    cloud_ = cloud;
    if (index_) ngicp_index_release(index_);
    index_ = idx;
    if (index_) ngicp_index_retain(index_);
  }
  inline PointCloudConstPtr getInputCloud() const { return cloud_; }

  // nanoflann_adaptor.h:141-152
  int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    k_indices.resize(k);
    k_sqr_distances.resize(k);
    if (!cloud_ || cloud_->points.empty()) return 0;   // nanoflann.h:1441
    if (!index_) throw std::runtime_error("[nanoflann] findNeighbors() called before building the index.");  // nanoflann.h:1442-1445
    const_cast<KdTreeFLANN*>(this)->ensure_handle();
    if (ngicp_knn(handle_, index_, &point, 1, sizeof(PointT), k, k_indices.data(), k_sqr_distances.data()) != NGICP_OK)
      throw std::runtime_error(ngicp_last_error(handle_));
    int found = 0;
    while (found < k && k_indices[found] >= 0) found++;
    return found;
  }
  ngicp_index* index() const { return index_; }

 private:
  void ensure_handle() {
    if (!handle_ && ngicp_create(device_, &handle_) != NGICP_OK) throw std::runtime_error(ngicp_last_error(nullptr));
  }
  int device_;
  ngicp_handle* handle_ = nullptr;
  ngicp_index* index_ = nullptr;
  PointCloudConstPtr cloud_;
};

}  // namespace nanoflann
