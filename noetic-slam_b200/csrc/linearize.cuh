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

constexpr int kLinThreads = 128;   // block size of every kernel that ends in block_publish

#ifdef __CUDACC__
// Reduce `NT` per-thread doubles over the block, publish the block partial, and let the last
// block of this batch entry (blockIdx.y) fold all partials (Neumaier) into the result slot.
template <int NT>
__device__ __forceinline__ void block_publish(double (*sm)[NT], double* __restrict__ partials, unsigned int* __restrict__ counters,
                                              ReduceSlot* __restrict__ slots, unsigned long long seq) {
  // sm[warp][term] holds every warp's running sums (written by that warp's lane 0)
  constexpr int kWarps = kLinThreads / 32;
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  const int b = blockIdx.y;
  double* my = partials + ((size_t)b * gridDim.x + blockIdx.x) * 32;
  if (threadIdx.x < NT) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) v += sm[w][threadIdx.x];
    my[threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(&counters[b], 1u);
    is_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // Fixed-order fold of the block partials by all 128 threads: thread (term, subset) = (tid % TP, tid / TP) takes
  // partials subset, subset + SUB, ... (Neumaier, eight independent loads in flight: this tail is pure L2 latency),
  // then a shuffle tree over the subsets of a warp and a fold over the four warps. One row write to the host-mapped
  // slot, one system fence, then the sequence number.
  constexpr int TP = NT <= 1 ? 1 : (NT <= 2 ? 2 : (NT <= 4 ? 4 : (NT <= 8 ? 8 : (NT <= 16 ? 16 : 32))));
  constexpr int SUB = kLinThreads / TP;
  __shared__ double fin[kWarps][32];
  const double* all = partials + (size_t)b * gridDim.x * 32;
  const unsigned int nb = gridDim.x;
  const int term = threadIdx.x % TP, subset = threadIdx.x / TP;
  double s = 0.0, c = 0.0;
  if (term < NT) {
    for (unsigned int i0 = subset; i0 < nb; i0 += SUB * 8) {
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const unsigned int i = i0 + u * SUB;
        x[u] = i < nb ? __ldcg(all + (size_t)i * 32 + term) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double y = s + x[u];
        c += (fabs(s) >= fabs(x[u])) ? ((s - y) + x[u]) : ((x[u] - y) + s);
        s = y;
      }
    }
  }
  double v = s + c;
#pragma unroll
  for (int off = TP; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane < TP) fin[warp][lane] = v;
  __syncthreads();
  if (warp != 0) return;
  if (lane < NT) {
    double t = fin[0][lane];
#pragma unroll
    for (int w = 1; w < kWarps; w++) t += fin[w][lane];
    slots[b].v[lane] = t;
  }
  __threadfence_system();
  __syncwarp();
  if (lane == 0) {
    counters[b] = 0;  // ready for the next launch
    *reinterpret_cast<volatile unsigned long long*>(&slots[b].seq) = seq;
  }
}

#endif

PoseArg make_pose(const double T_colmajor[16]);
int lin_blocks_for(int n);

}  // namespace ngicp
