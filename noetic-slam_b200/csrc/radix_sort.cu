// Stable LSD radix sort, 9-bit digits (radix_sort.cuh), hand-written for sm_100a (no CUB/Thrust).
//
// Per sort:   1 histogram kernel over all passes (digit totals do not depend on order)
//             1 tiny scan kernel -> global start of every digit of every pass
// Per pass:   count   : per-tile digit counts            counts[digit][tile]
//             scanrow : exclusive scan of every digit row over tiles
//             scatter : warp-synchronous stable ranking (match_any) + scatter
// A tile is 512 threads x ITEMS keys (4, or 1 for small inputs); warp w of a tile owns the contiguous chunk of 32*ITEMS keys
// [w*32*ITEMS, (w+1)*32*ITEMS) so that (tile, warp, round, lane) order == input order, which makes the
// scatter stable. All global reads are coalesced 256-byte warp rows.
#include "radix_sort.cuh"

namespace ngicp {

namespace {

__global__ void __launch_bounds__(256) sort_histogram_kernel(const unsigned long long* __restrict__ keys, int n, int low_bit, int passes,
                                                             uint32_t* __restrict__ digit_hist) {
  __shared__ uint32_t h[8 * kSortRadix];
  for (int i = threadIdx.x; i < passes * kSortRadix; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    for (int p = 0; p < passes; p++) atomicAdd(&h[p * kSortRadix + sort_digit_of(k, low_bit, p)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kSortRadix; i += blockDim.x)
    if (h[i]) atomicAdd(&digit_hist[i], h[i]);
}

// counts layout: [tile][digit] when the scatter kernel sums the earlier tiles itself (few tiles), [digit][tile] when
// sort_scan_rows_kernel turns every digit row into an exclusive prefix first (many tiles)
template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) sort_count_kernel(const unsigned long long* __restrict__ keys, int n, int shift,
                                                                 uint32_t* __restrict__ counts, int nblocks, int rows_scanned) {
  __shared__ uint32_t h[kSortRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * (kSortThreads * ITEMS);
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & (kSortRadix - 1))], 1u);
  }
  __syncthreads();
  counts[rows_scanned ? threadIdx.x * nblocks + blockIdx.x : blockIdx.x * kSortRadix + threadIdx.x] = h[threadIdx.x];
}

// grid = 256 blocks (one digit row each), 256 threads; exclusive scan of counts[d][0..nblocks)
__global__ void __launch_bounds__(256) sort_scan_rows_kernel(uint32_t* __restrict__ counts, int nblocks) {
  __shared__ uint32_t s[256];
  uint32_t* row = counts + (size_t)blockIdx.x * nblocks;
  const int chunk = (nblocks + 255) / 256;
  const int b0 = threadIdx.x * chunk;
  const int b1 = min(b0 + chunk, nblocks);
  uint32_t sum = 0;
  for (int b = b0; b < b1; b++) sum += row[b];
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {
    const uint32_t t = threadIdx.x >= off ? s[threadIdx.x - off] : 0u;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = s[threadIdx.x] - sum;
  for (int b = b0; b < b1; b++) {
    const uint32_t c = row[b];
    row[b] = run;
    run += c;
  }
}

template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                   unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                   const uint32_t* __restrict__ counts, const uint32_t* __restrict__ digit_total,
                                                                   int n, int shift, int nblocks, int rows_scanned) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t warp_cnt[kWarps][kSortRadix];
  __shared__ uint32_t digit_base[kSortRadix];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int w = 0; w < kWarps; w++) warp_cnt[w][threadIdx.x] = 0;
  {
    // global start of this thread's digit = exclusive scan of the 256 raw digit totals of the pass
    const uint32_t v = digit_total[threadIdx.x];
    digit_base[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kSortRadix; off <<= 1) {
      const uint32_t t = threadIdx.x >= off ? digit_base[threadIdx.x - off] : 0u;
      __syncthreads();
      digit_base[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t row = 0;  // keys with this digit in earlier tiles
    if (rows_scanned) row = counts[threadIdx.x * nblocks + blockIdx.x];
    else {
      // coalesced rows, sixteen loads in flight: the last tiles sum > 100 rows and this chain is the kernel's long pole
      int b = 0;
      for (; b + 16 <= (int)blockIdx.x; b += 16) {
        uint32_t v16[16];
#pragma unroll
        for (int u = 0; u < 16; u++) v16[u] = __ldg(counts + (b + u) * kSortRadix + threadIdx.x);
#pragma unroll
        for (int u = 0; u < 16; u++) row += v16[u];
      }
      for (; b + 4 <= (int)blockIdx.x; b += 4)
        row += (__ldg(counts + b * kSortRadix + threadIdx.x) + __ldg(counts + (b + 1) * kSortRadix + threadIdx.x)) +
               (__ldg(counts + (b + 2) * kSortRadix + threadIdx.x) + __ldg(counts + (b + 3) * kSortRadix + threadIdx.x));
      for (; b < (int)blockIdx.x; b++) row += __ldg(counts + b * kSortRadix + threadIdx.x);
    }
    const uint32_t start = digit_base[threadIdx.x] - v;
    __syncthreads();
    digit_base[threadIdx.x] = start + row;
  }
  __syncthreads();

  const int chunk0 = blockIdx.x * (kSortThreads * ITEMS) + warp * (32 * ITEMS);
  unsigned long long key[ITEMS];
  int dig[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const int i = chunk0 + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? keys_in[i] : 0ull;
    dig[r] = valid ? (int)((key[r] >> shift) & (kSortRadix - 1)) : (kSortRadix + lane);  // invalid lanes match nobody valid
  }
  // phase 1: per-warp digit counts
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
    if (dig[r] < kSortRadix && lane == __ffs(m) - 1) warp_cnt[warp][dig[r]] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over the warps of the tile, per digit (thread == digit)
  {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const uint32_t c = warp_cnt[w][threadIdx.x];
      warp_cnt[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase 2: stable rank + scatter
#pragma unroll
  for (int r = 0; r < ITEMS; r++) {
    const unsigned m = __match_any_sync(0xffffffffu, dig[r]);
    const bool valid = dig[r] < kSortRadix;
    uint32_t pos = 0;
    if (valid) pos = digit_base[dig[r]] + warp_cnt[warp][dig[r]] + __popc(m & lt_mask);
    __syncwarp();
    if (valid && lane == __ffs(m) - 1) warp_cnt[warp][dig[r]] += __popc(m);
    __syncwarp();
    if (valid) {
      const int i = chunk0 + r * 32 + lane;
      keys_out[pos] = key[r];
      vals_out[pos] = vals_in[i];
    }
  }
}

}  // namespace

int radix_sort_pairs(unsigned long long* keys_a, uint32_t* vals_a, unsigned long long* keys_b, uint32_t* vals_b,
                     uint32_t* scratch, int n, int low_bit, int nbits, cudaStream_t stream,
                     unsigned long long** out_keys, uint32_t** out_vals, bool hist_ready) {
  *out_keys = keys_a;
  *out_vals = vals_a;
  if (n <= 1 || nbits <= 0) return 0;
  const int passes = sort_num_passes(nbits);
  const int nblocks = sort_num_blocks(n);
  uint32_t* digit_hist = scratch;                       // [passes][256] raw digit totals
  uint32_t* counts = scratch + passes * kSortRadix;     // [256][nblocks]
  int launches = 0;
  if (!hist_ready) {
    cudaMemsetAsync(digit_hist, 0, sizeof(uint32_t) * passes * kSortRadix, stream);
    const int hist_blocks = max(1, min(nblocks, 148 * 4));
    sort_histogram_kernel<<<hist_blocks, 256, 0, stream>>>(keys_a, n, low_bit, passes, digit_hist);
    launches += 1;
  }
  const bool scan_rows = nblocks > 256;   // few tiles: the scatter kernel sums the earlier tiles itself
  const int items = sort_items_for(n);
  unsigned long long* kin = keys_a; uint32_t* vin = vals_a;
  unsigned long long* kout = keys_b; uint32_t* vout = vals_b;
  for (int p = 0; p < passes; p++) {
    const int shift = low_bit + p * kSortRadixBits;
    if (items == 1) {
      sort_count_kernel<1><<<nblocks, kSortThreads, 0, stream>>>(kin, n, shift, counts, nblocks, scan_rows ? 1 : 0);
      if (scan_rows) sort_scan_rows_kernel<<<kSortRadix, 256, 0, stream>>>(counts, nblocks);
      sort_scatter_kernel<1><<<nblocks, kSortThreads, 0, stream>>>(kin, vin, kout, vout, counts, digit_hist + p * kSortRadix, n, shift, nblocks, scan_rows ? 1 : 0);
    } else {
      sort_count_kernel<4><<<nblocks, kSortThreads, 0, stream>>>(kin, n, shift, counts, nblocks, scan_rows ? 1 : 0);
      if (scan_rows) sort_scan_rows_kernel<<<kSortRadix, 256, 0, stream>>>(counts, nblocks);
      sort_scatter_kernel<4><<<nblocks, kSortThreads, 0, stream>>>(kin, vin, kout, vout, counts, digit_hist + p * kSortRadix, n, shift, nblocks, scan_rows ? 1 : 0);
    }
    launches += scan_rows ? 3 : 2;
    unsigned long long* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *out_keys = kin;
  *out_vals = vin;
  return launches;
}

}  // namespace ngicp
