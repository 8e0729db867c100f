// BENCH/TEST helper: the unmodified call sequence of dlio::OdomNode (src/dlio/src/dlio/odom.cc:89-107, 721-722, 992-1008,
// 1719-1738) over include/nano_gicp/nano_gicp.h, compiled against the PCL/Eigen shims of this directory (PCL is absent from
// this image). Scans are fresh pageable pcl::PointCloud objects, as in DLIO; the timed step is setInputSource +
// calculateSourceCovariances + align(output).
//   dropin_bench <target.bin> <bounds.bin> <steps> <warmup> <compute_output 0|1> <scan0.bin> [scan1.bin ...]
// target/scans: float32 xyz triples; bounds: int64 keyframe boundaries. Prints one JSON line.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <nano_gicp/nano_gicp.h>

struct Point {  // layout of dlio::Point (src/dlio/include/dlio/dlio.h:85-108): 32 bytes, xyz first, w = 1
  float x, y, z, w = 1.f;
  float intensity = 0.f;
  float pad_[3] = {0, 0, 0};
};
static_assert(sizeof(Point) == 32, "dlio::Point is 32 bytes");
using Cloud = pcl::PointCloud<Point>;

static std::vector<float> read_f32(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<float> v(n / sizeof(float));
  if (std::fread(v.data(), sizeof(float), v.size(), f) != v.size()) std::exit(2);
  std::fclose(f);
  return v;
}
static Cloud::Ptr cloud_from(const float* xyz, size_t n) {
  Cloud::Ptr c(new Cloud);
  c->points.resize(n);
  for (size_t i = 0; i < n; i++) { c->points[i].x = xyz[3 * i]; c->points[i].y = xyz[3 * i + 1]; c->points[i].z = xyz[3 * i + 2]; }
  c->width = (std::uint32_t)n; c->height = 1;
  return c;
}

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const std::vector<float> tgt = read_f32(argv[1]);
  std::vector<int64_t> bounds;
  {
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 2;
    int64_t v;
    while (std::fread(&v, sizeof v, 1, f) == 1) bounds.push_back(v);
    std::fclose(f);
  }
  const int steps = std::atoi(argv[3]), warmup = std::atoi(argv[4]);
  const bool compute_output = std::atoi(argv[5]) != 0;
  std::vector<std::vector<float>> scans;
  for (int i = 6; i < argc; i++) scans.push_back(read_f32(argv[i]));

  nano_gicp::NanoGICP<Point, Point> gicp, gicp_temp;
  for (auto* g : {&gicp, &gicp_temp}) {   // odom.cc:89-101, cfg/params.yaml:57-63
    g->setCorrespondenceRandomness(16);
    g->setMaxCorrespondenceDistance(0.5);
    g->setMaximumIterations(32);
    g->setTransformationEpsilon(0.01);
    g->setRotationEpsilon(0.01);
    g->setInitialLambdaFactor(1e-9);
  }
  gicp.setComputeOutputCloud(compute_output);
  // keyframes: covariances computed when each was a scan (odom.cc:721-722), kept per keyframe (:1592), concatenated per
  // submap (:1719-1729)
  auto submap_normals = std::make_shared<nano_gicp::CovarianceList>();
  for (size_t s = 0; s + 1 < bounds.size(); s++) {
    auto kf = cloud_from(tgt.data() + 3 * bounds[s], (size_t)(bounds[s + 1] - bounds[s]));
    gicp_temp.setInputSource(kf);
    gicp_temp.calculateSourceCovariances();
    auto c = gicp_temp.getSourceCovariances();
    submap_normals->insert(submap_normals->end(), c->begin(), c->end());
  }
  auto submap_cloud = cloud_from(tgt.data(), tgt.size() / 3);
  gicp_temp.setInputTarget(submap_cloud);                  // odom.cc:1737
  gicp.registerInputTarget(submap_cloud);                  // odom.cc:992-998
  gicp.target_kdtree_ = gicp_temp.target_kdtree_;
  gicp.setTargetCovariances(submap_normals);

  double total = 0.0;
  int iters = 0;
  for (int i = 0; i < warmup + steps; i++) {
    const auto& sc = scans[i % scans.size()];
    auto cloud = cloud_from(sc.data(), sc.size() / 3);     // a fresh pageable cloud per scan, outside the timed region
    Cloud::Ptr aligned(new Cloud);
    const auto t0 = std::chrono::steady_clock::now();
    gicp.setInputSource(cloud);
    gicp.calculateSourceCovariances();
    gicp.align(*aligned);
    const auto t1 = std::chrono::steady_clock::now();
    if (i >= warmup) { total += std::chrono::duration<double>(t1 - t0).count(); iters += 1; }
  }
  const auto T = gicp.getFinalTransformation();
  std::printf("{\"steps\": %d, \"ms_per_step\": %.6f, \"scans_per_s\": %.3f, \"converged\": %d, \"compute_output\": %d, \"t_last\": [%.6f, %.6f, %.6f]}\n", iters,
              1e3 * total / iters, iters / total, gicp.hasConverged() ? 1 : 0, compute_output ? 1 : 0, T(0, 3), T(1, 3), T(2, 3));
  return 0;
}
