// Device-side LSD radix sort (u64 key, u32 payload), stable, and the ordered-walk reductions the
// correctors need.  Replaces the global std::priority_queue<KmerSign> drain of
// sorted_aggregator::run (reference include/kmdiff/aggregator.hpp:282-310, 325-341) and the
// per-record corrector->apply loop of aggregator::worker (aggregator.hpp:137-163).
//
// Sizes here are the prefilter survivors (<= ~1e-3 of the rows), so the sort favours simplicity:
// 8 passes of 8 bits; each warp owns a tile of consecutive keys, ranks equal digits with
// __match_any_sync in lane order (=> stable) and keeps running per-digit offsets in shared memory;
// the eight warps of a CTA share one histogram column, so the table that is prefix-summed between the
// two kernels of a pass has 256 x (n / 4096) entries.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace kmd {

constexpr int kSortTile = 512;   // keys per warp
constexpr int kSortWarps = 8;    // warps per CTA; a CTA owns kSortTile * kSortWarps consecutive keys
constexpr int kSortCtaKeys = kSortTile * kSortWarps;

// digit counts of each warp's sub-tile of the CTA's keys -> counts[warp][digit] (shared memory)
__device__ __forceinline__ void sort_count_tile(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                uint32_t (*counts)[256])
{
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t d = lane; d < 256; d += 32) counts[warp][d] = 0;
  __syncwarp();
  const uint32_t begin = blockIdx.x * kSortCtaKeys + warp * kSortTile;
  const uint32_t end = min(n, begin + kSortTile);
  for (uint32_t i = begin + lane; i < end; i += 32)
    atomicAdd(&counts[warp][(keys[i] >> shift) & 0xFF], 1u);
  __syncthreads();
}

// one histogram per CTA (not per warp): the table the single-CTA prefix sum walks stays small
__global__ void __launch_bounds__(kSortWarps * 32)
sort_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t n_ctas,
                 uint32_t* __restrict__ hist /* [256][n_ctas] */)
{
  __shared__ uint32_t counts[kSortWarps][256];
  sort_count_tile(keys, n, shift, counts);
  const uint32_t d = threadIdx.x;   // kSortWarps * 32 == 256 digits
  uint32_t s = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++) s += counts[w][d];
  hist[d * n_ctas + blockIdx.x] = s;
}

// Exclusive prefix sum of `n` u32 values in place, one CTA (n is at most a few hundred thousand).
// Optionally writes the grand total to *total.
__global__ void __launch_bounds__(1024) scan_exclusive_kernel(uint32_t* data, uint32_t n, uint32_t* total)
{
  __shared__ uint32_t partial[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; i++) s += data[i];
  partial[t] = s;
  __syncthreads();
  // Hillis-Steele over the 1024 partials
  for (uint32_t o = 1; o < 1024; o <<= 1)
  {
    uint32_t v = t >= o ? partial[t - o] : 0;
    __syncthreads();
    partial[t] += v;
    __syncthreads();
  }
  uint32_t run = t ? partial[t - 1] : 0;
  for (uint32_t i = lo; i < hi; i++)
  {
    const uint32_t v = data[i];
    data[i] = run;
    run += v;
  }
  if (total && t == 1023) *total = partial[1023];
}

// vals_in == nullptr: keys only
__global__ void __launch_bounds__(kSortWarps * 32)
sort_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                    int shift, uint32_t n_ctas, const uint32_t* __restrict__ offsets /* scanned hist */)
{
  __shared__ uint32_t base[kSortWarps][256];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sort_count_tile(keys_in, n, shift, base);
  {
    // counts -> first output slot of (warp, digit): the CTA's offset of the digit + the counts of the warps before
    const uint32_t d = threadIdx.x;
    uint32_t run = offsets[d * n_ctas + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; w++)
    {
      const uint32_t c = base[w][d];
      base[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  const uint32_t begin = blockIdx.x * kSortCtaKeys + warp * kSortTile;
  const uint32_t end = min(n, begin + kSortTile);
  for (uint32_t i0 = begin; i0 < end; i0 += 32)
  {
    const uint32_t i = i0 + lane;
    const bool valid = i < end;
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, valid);
    if (valid)
    {
      const uint64_t key = keys_in[i];
      const uint32_t d = (key >> shift) & 0xFF;
      const uint32_t peers = __match_any_sync(active, d);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1));
      const uint32_t dst = base[warp][d] + rank;
      __syncwarp(active);
      if (rank == 0) base[warp][d] += __popc(peers);
      __syncwarp(active);
      keys_out[dst] = key;
      if (vals_in) vals_out[dst] = vals_in[i];
    }
  }
}

// ---- ordered-walk helpers --------------------------------------------------------------------

// BENJAMINI / HOLM over p-values already in ascending order: index of the first rejected element
// (corrector.cpp:27-35: p < (rank/m)*q with rank = i+1;  corrector.cpp:68-71: p < t/(m-i)).
// `first` must be preset to n.
__global__ void first_reject_kernel(const double* __restrict__ p_sorted, uint32_t n, int holm,
                                    double level, uint64_t total, uint32_t* first)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool pass;
  if (holm)
    pass = p_sorted[i] < __ddiv_rn(level, static_cast<double>(total - i));
  else
    pass = p_sorted[i] < __dmul_rn(__ddiv_rn(static_cast<double>(static_cast<uint64_t>(i) + 1),
                                             static_cast<double>(total)), level);
  if (!pass) atomicMin(first, i);
}

// the same walk with m read from device memory (multi-GPU exchange: m = sum of the ranks' row counts)
__global__ void first_reject_dev_kernel(const double* __restrict__ p_sorted, uint32_t n, int holm, double level,
                                        const uint64_t* __restrict__ total_ptr, uint32_t* first)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t total = *total_ptr;
  bool pass;
  if (holm)
    pass = p_sorted[i] < __ddiv_rn(level, static_cast<double>(total - i));
  else
    pass = p_sorted[i] < __dmul_rn(__ddiv_rn(static_cast<double>(static_cast<uint64_t>(i) + 1),
                                             static_cast<double>(total)), level);
  if (!pass) atomicMin(first, i);
}

}  // namespace kmd
