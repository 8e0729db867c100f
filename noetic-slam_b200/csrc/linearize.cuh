// Shared between linearize.cu and api.cu.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace ngicp {

constexpr int kMaxLinBlocks = 148 * 16;
constexpr int kMaxBatch = 256;

// Pose handed to the kernels by value: fp64 for the residual / Mahalanobis, fp32 for the
// correspondence query (the reference casts the Isometry3d to float once, nano_gicp.cc:210).
struct PoseArg {
  double R[9];  // row-major
  double t[3];
  float Rf[9];
  float tf[3];
};

// Host-mapped result slot of one reduction: 29 doubles + a sequence number written last.
struct ReduceSlot {
  double v[32];
  unsigned long long seq;
  unsigned long long pad[7];
};

PoseArg make_pose(const double T_colmajor[16]);
int lin_blocks_for(int n);

}  // namespace ngicp
