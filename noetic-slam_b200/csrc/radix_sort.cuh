// Stable LSD radix sort of (64-bit voxel key, 32-bit point index) pairs — the "radix-sorted integer
// keys" step of the Morton/voxel-hash index that replaces the reference's recursive KD-tree build
// (reference src/dlio/include/nano_gicp/nanoflann.h:1025-1185).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngicp {

// 9-bit digits: the 36 key bits of a single cloud's index sort in four passes (32-bit VoxelGrid indices too). One thread
// per digit in the count / scan steps, so a tile is 512 threads.
constexpr int kSortRadixBits = 9;
constexpr int kSortRadix = 1 << kSortRadixBits;
constexpr int kSortThreads = kSortRadix;
// keys per thread: 1 for small inputs (a 65,536-point scan becomes 128 tiles instead of 32 and reaches most SMs), 4 otherwise
inline int sort_items_for(int n) { return n <= 131072 ? 1 : 4; }
inline int sort_num_blocks(int n) { const int tile = kSortThreads * sort_items_for(n); return (n + tile - 1) / tile; }
inline int sort_num_passes(int nbits) { return (nbits + kSortRadixBits - 1) / kSortRadixBits; }
// scratch (uint32 elements): per-pass digit totals [passes*radix] + per-tile digit counts [radix*tiles]
inline size_t sort_scratch_elems(int n, int nbits) {
  return (size_t)sort_num_passes(nbits) * kSortRadix + (size_t)kSortRadix * sort_num_blocks(n) + 64;
}

// Sorts n pairs by key bits [low_bit, low_bit + nbits). (keys_a, vals_a) hold the input; (keys_b, vals_b) are
// ping-pong buffers of the same size. On return *out_keys / *out_vals point at whichever pair
// holds the sorted result (pair A after an even number of passes, pair B after an odd number).
// hist_ready: the caller already accumulated the raw per-pass digit totals into scratch[0 .. passes*256)
// (sort_digit_of gives the digit), which saves the histogram launch. Returns the number of kernels launched.
int radix_sort_pairs(unsigned long long* keys_a, uint32_t* vals_a, unsigned long long* keys_b, uint32_t* vals_b,
                     uint32_t* scratch, int n, int low_bit, int nbits, cudaStream_t stream,
                     unsigned long long** out_keys, uint32_t** out_vals, bool hist_ready);

__host__ __device__ __forceinline__ int sort_digit_of(unsigned long long key, int low_bit, int pass) {
  return (int)((key >> (low_bit + pass * kSortRadixBits)) & (kSortRadix - 1));
}

}  // namespace ngicp
