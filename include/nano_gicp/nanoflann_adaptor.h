// Drop-in replacement of the reference's include/nano_gicp/nanoflann_adaptor.h (:57-193):
// nanoflann::KdTreeFLANN<PointT> keeps its name and its setInputCloud / nearestKSearch / getInputCloud
// members (DLIO declares std::shared_ptr<const nanoflann::KdTreeFLANN<PointType>> submap_kdtree,
// include/dlio/odom.h:164), but the "tree" is the B200 voxel-hash index behind the C ABI.
#pragma once
#include <memory>
#include <stdexcept>
#include <vector>

#include <pcl/pcl_config.h>
#include <pcl/point_cloud.h>
#if PCL_VERSION_COMPARE(<, 1, 11, 0)
#include <boost/shared_ptr.hpp>   // PCL <= 1.10 (DLIO's Ubuntu 20.04 image) hands clouds and index lists around as boost::shared_ptr
#endif

#include "../ngicp_b200.h"

namespace nanoflann {

template <typename PointT>
class KdTreeFLANN {
 public:
  typedef typename pcl::PointCloud<PointT> PointCloud;
  typedef typename pcl::PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename pcl::PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  // reference nanoflann_adaptor.h:66-69
#if PCL_VERSION_COMPARE(<, 1, 11, 0)
  typedef boost::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef boost::shared_ptr<const KdTreeFLANN<PointT>> ConstPtr;
  typedef boost::shared_ptr<std::vector<int>> IndicesPtr;
  typedef boost::shared_ptr<const std::vector<int>> IndicesConstPtr;
#else
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef std::shared_ptr<const KdTreeFLANN<PointT>> ConstPtr;
  typedef std::shared_ptr<std::vector<int>> IndicesPtr;
  typedef std::shared_ptr<const std::vector<int>> IndicesConstPtr;
#endif

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
  // nanoflann_adaptor.h:155-174 (unused by GICP). As in the reference, `radius` bounds the SQUARED distance (it is handed to
  // nanoflann's RadiusResultSet as is) and the test is strict; results come back sorted by (distance, index).
  int radiusSearch(const PointT& point, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    k_indices.clear();
    k_sqr_distances.clear();
    if (!cloud_ || cloud_->points.empty()) return 0;
    if (!index_) throw std::runtime_error("[nanoflann] findNeighbors() called before building the index.");
    const_cast<KdTreeFLANN*>(this)->ensure_handle();
    const size_t n = ngicp_index_size(index_);
    for (size_t k = 32;; k *= 4) {      // exact k-NN with a growing k until the k-th neighbour falls outside the radius
      if (k > n) k = n;
      if (k > 128) k = 128;             // ngicp_knn serves k <= 128
      std::vector<int> idx(k);
      std::vector<float> sqd(k);
      if (ngicp_knn(handle_, index_, &point, 1, sizeof(PointT), (int)k, idx.data(), sqd.data()) != NGICP_OK)
        throw std::runtime_error(ngicp_last_error(handle_));
      size_t found = 0;
      while (found < k && idx[found] >= 0 && (double)sqd[found] < radius) found++;
      if (found < k || k == n || k == 128) {
        if (found == k && k == 128 && n > 128) throw std::runtime_error("KdTreeFLANN(b200)::radiusSearch: more than 128 points inside the radius");
        k_indices.assign(idx.begin(), idx.begin() + found);
        k_sqr_distances.assign(sqd.begin(), sqd.begin() + found);
        return (int)found;
      }
    }
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
