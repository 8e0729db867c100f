// TEST-ONLY: exercises include/nano_gicp/nano_gicp.h the way dlio::OdomNode does (src/dlio/odom.cc:89-107,
// 721-722, 992-1008, 1737-1738) against the shims in this directory. Reads two clouds of 32-byte dlio::Point
// records (float32 xyz + padding) from raw files, prints pose / iterations / density as one line.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <nano_gicp/nano_gicp.h>

struct Point {  // layout of dlio::Point (src/dlio/include/dlio/dlio.h:85-108): 32 bytes, xyz first, w = 1
  float x, y, z, w = 1.f;
  float intensity = 0.f;
  float pad_[3] = {0, 0, 0};
};
static_assert(sizeof(Point) == 32, "dlio::Point is 32 bytes");

static pcl::PointCloud<Point>::Ptr load(const char* path) {
  pcl::PointCloud<Point>::Ptr c(new pcl::PointCloud<Point>);   // boost::shared_ptr with PCL <= 1.10, std::shared_ptr from 1.11
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  float v[3];
  while (std::fread(v, sizeof(float), 3, f) == 3) { Point p; p.x = v[0]; p.y = v[1]; p.z = v[2]; c->points.push_back(p); }
  std::fclose(f);
  return c;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  auto src = load(argv[1]), tgt = load(argv[2]);
  nano_gicp::NanoGICP<Point, Point> gicp, gicp_temp;
  for (auto* g : {&gicp, &gicp_temp}) {   // odom.cc:89-101
    g->setCorrespondenceRandomness(16);
    g->setMaxCorrespondenceDistance(0.5);
    g->setMaximumIterations(32);
    g->setTransformationEpsilon(0.01);
    g->setRotationEpsilon(0.01);
    g->setInitialLambdaFactor(1e-9);
  }
  // submap thread: build the target tree on gicp_temp, hand it over (odom.cc:1737-1738 -> :992-998)
  gicp_temp.setInputTarget(tgt);
  gicp_temp.calculateTargetCovariances();
  auto submap_normals = gicp_temp.getTargetCovariances();
  gicp.registerInputTarget(tgt);
  gicp.target_kdtree_ = gicp_temp.target_kdtree_;
  gicp.setTargetCovariances(submap_normals);
  // lidar thread (odom.cc:721-722, :1005-1008)
  gicp.setInputSource(src);
  gicp.calculateSourceCovariances();
  pcl::PointCloud<Point> aligned;
  gicp.align(aligned);
  const auto T = gicp.getFinalTransformation();
  std::printf("T");
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf(" %.9g", T(r, c));
  std::printf(" iters %d converged %d density %.9g ncov %zu aligned0 %.9g %.9g %.9g\n", 0, gicp.hasConverged() ? 1 : 0, gicp.source_density_,
              gicp.getSourceCovariances()->size(), aligned.points[0].x, aligned.points[0].y, aligned.points[0].z);
  std::vector<int> ki; std::vector<float> kd;
  const int found = gicp.target_kdtree_->nearestKSearch(src->points[0], 3, ki, kd);
  std::printf("knn %d %d %d %d %.9g %.9g %.9g\n", found, ki[0], ki[1], ki[2], kd[0], kd[1], kd[2]);
  // crop + voxel grid + setInputSource in one device call (replaces odom.cc:501-502, :579-580, :721), then align again
  auto filtered = gicp.setInputSourceFiltered(src, 1.0f, 0.25f);
  gicp.calculateSourceCovariances();
  gicp.align(aligned);
  const auto T2 = gicp.getFinalTransformation();
  std::printf("filtered %zu first %.9g %.9g %.9g T", filtered->points.size(), filtered->points[0].x, filtered->points[0].y, filtered->points[0].z);
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf(" %.9g", T2(r, c));
  std::printf("\n");
  // the hand-over DLIO actually makes (odom.cc:1737-1738 -> :992-998): the submap tree is built on gicp_temp and adopted by
  // gicp WITHOUT anything on gicp_temp that waits for the build (no calculateTargetCovariances there: the submap's
  // covariances are the keyframes'); the adopter must order itself after the build
  {
    nano_gicp::NanoGICP<Point, Point> a, b;
    for (auto* g : {&a, &b}) {
      g->setCorrespondenceRandomness(16);
      g->setMaxCorrespondenceDistance(0.5);
      g->setMaximumIterations(32);
      g->setTransformationEpsilon(0.01);
      g->setRotationEpsilon(0.01);
    }
    b.setInputTarget(tgt);                       // index build in flight on b's stream
    a.registerInputTarget(tgt);
    a.target_kdtree_ = b.target_kdtree_;         // adopted at once
    a.setTargetCovariances(submap_normals);
    a.setInputSource(src);
    a.calculateSourceCovariances();
    a.align(aligned);
    const auto T3 = a.getFinalTransformation();
    std::printf("handover T");
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf(" %.9g", T3(r, c));
    std::printf("\n");
  }
  return 0;
}
