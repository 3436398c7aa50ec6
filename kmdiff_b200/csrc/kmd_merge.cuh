// Device-side N-way merge of per-sample sorted (k-mer, count) streams into dense count-matrix rows.
//
// Replaces km::KmerMerger::next (reference thirdparty/kmtricks/include/kmtricks/merge.hpp:183-260) as
// kmdiff drives it (include/kmdiff/merge.hpp:266-267: recurrence_min = 1, save_if = 0) when the run
// directory holds counts/partition_{p}/<sample>.kmer[.lz4] files instead of matrices
// (cmd/diff.hpp:88-95, kmdir.hpp:48-73; record layout kmer_file.hpp:127-137).
//
// The reference scans all N stream heads for every output row (O(N) compares per row, one thread per
// partition).  Here the N sorted streams of a chunk are the leaves of a pairwise merge tree:
// ceil(log2 N) merge-path passes over (key, payload) arrays in HBM, each pass one coalesced read and
// one coalesced write of every record; then equal keys are adjacent, a head flag + prefix sum numbers
// the rows, and one warp per row writes the dense row (zeros + the samples that carry the k-mer) in
// kmtricks matrix layout, ready for lrt_scan_kernel without leaving the device.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace kmd {
namespace mrg {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;   // outputs per CTA of a merge pass
constexpr int kScanTile = 4096;            // flags per CTA of the row-numbering scan

// payload: bits 0..15 stream (sample) index, bit 16 "solid" (count >= ab_min of that sample), bits 32..63 count
__host__ __device__ inline uint64_t make_payload(uint32_t stream, uint32_t count, bool solid)
{
  return static_cast<uint64_t>(stream) | (solid ? (1ull << 16) : 0ull) | (static_cast<uint64_t>(count) << 32);
}
__host__ __device__ inline uint32_t pay_stream(uint64_t p) { return static_cast<uint32_t>(p & 0xFFFFu); }
__host__ __device__ inline bool pay_solid(uint64_t p) { return (p >> 16) & 1u; }
__host__ __device__ inline uint32_t pay_count(uint64_t p) { return static_cast<uint32_t>(p >> 32); }

template <int W>
struct Key;
template <>
struct Key<1>
{
  uint64_t lo;
  __device__ bool operator<=(const Key& o) const { return lo <= o.lo; }
  __device__ bool operator==(const Key& o) const { return lo == o.lo; }
};
template <>
struct Key<2>
{
  uint64_t lo, hi;
  __device__ bool operator<=(const Key& o) const { return hi < o.hi || (hi == o.hi && lo <= o.lo); }
  __device__ bool operator==(const Key& o) const { return hi == o.hi && lo == o.lo; }
};

template <int W>
struct KeyArrays
{
  uint64_t* lo;
  uint64_t* hi;   // unused for W == 1
  __device__ Key<W> get(uint32_t i) const
  {
    Key<W> k;
    k.lo = lo[i];
    if constexpr (W == 2) k.hi = hi[i];
    return k;
  }
  __device__ void put(uint32_t i, const Key<W>& k) const
  {
    lo[i] = k.lo;
    if constexpr (W == 2) hi[i] = k.hi;
  }
};

// ---- records -> (key, payload) -----------------------------------------------------------------

struct UnpackArgs
{
  const uint8_t* raw;          // the streams' records back to back
  const uint64_t* stream_off;  // [n_streams + 1] record offsets
  const uint32_t* ab_min;      // [n_streams]
  uint32_t n_streams;
  uint32_t n;
  uint32_t kmer_words, count_bytes;
  uint64_t* klo;
  uint64_t* khi;
  uint64_t* pay;
};

__device__ inline uint64_t load_le(const uint8_t* p, int bytes)
{
  uint64_t v = 0;
  for (int b = 0; b < bytes; b++) v |= static_cast<uint64_t>(p[b]) << (8 * b);
  return v;
}

__global__ void __launch_bounds__(256) unpack_kernel(const UnpackArgs A)
{
  const uint32_t rec = 8 * A.kmer_words + A.count_bytes;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A.n; j += gridDim.x * blockDim.x)
  {
    // stream of record j: last s with stream_off[s] <= j
    uint32_t lo = 0, hi = A.n_streams;
    while (hi - lo > 1)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (A.stream_off[mid] <= j) lo = mid; else hi = mid;
    }
    const uint8_t* r = A.raw + static_cast<size_t>(j) * rec;
    A.klo[j] = load_le(r, 8);
    if (A.kmer_words == 2) A.khi[j] = load_le(r + 8, 8);
    const uint32_t c = static_cast<uint32_t>(load_le(r + 8 * A.kmer_words, A.count_bytes));
    // merge.hpp:204-223: a count below the sample's abundance_min is reported as 0 and does not
    // count towards the recurrence of the row
    A.pay[j] = make_payload(lo, c, c >= A.ab_min[lo]);
  }
}

// ---- one pass of the merge tree -------------------------------------------------------------------

// Runs r = 0..n_runs-1 cover [run_off[r], run_off[r+1]).  Pair q merges runs 2q and 2q+1 (a lone last
// run is a pair with an empty right side) into the same index range of the output arrays.
// pair_tile[q] = first CTA of pair q (host-built prefix sum of ceil(len/kTile)).
template <int W>
struct MergeArgs
{
  KeyArrays<W> in, out;
  const uint64_t* pay_in;
  uint64_t* pay_out;
  const uint32_t* run_off;
  const uint32_t* pair_tile;
  uint32_t n_runs, n_pairs;
};

template <int W, class GetA, class GetB>
__device__ inline uint32_t merge_path(uint32_t diag, uint32_t len_a, uint32_t len_b, GetA a, GetB b)
{
  uint32_t lo = diag > len_b ? diag - len_b : 0, hi = diag < len_a ? diag : len_a;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (a(mid) <= b(diag - 1 - mid)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// pair and tile geometry of CTA `cta` of a pass
template <int W>
struct TileGeom { uint32_t a0, b0, len_a, len_b, d0, d1; };

template <int W>
__device__ inline TileGeom<W> tile_geom(const MergeArgs<W>& A, uint32_t cta)
{
  uint32_t ql = 0, qh = A.n_pairs;
  while (qh - ql > 1)
  {
    const uint32_t mid = (ql + qh) >> 1;
    if (A.pair_tile[mid] <= cta) ql = mid; else qh = mid;
  }
  const uint32_t q = ql;
  TileGeom<W> g;
  g.a0 = A.run_off[2 * q];
  const uint32_t a1 = A.run_off[2 * q + 1];
  g.b0 = a1;
  const uint32_t b1 = (2 * q + 2 <= A.n_runs) ? A.run_off[2 * q + 2] : a1;
  g.len_a = a1 - g.a0;
  g.len_b = b1 - g.b0;
  g.d0 = (cta - A.pair_tile[q]) * kTile;
  g.d1 = min(g.d0 + kTile, g.len_a + g.len_b);
  return g;
}

// split[cta] = how many records of run A lie before the first output of tile `cta` (merge-path search in global
// memory, one thread per tile): the merge kernel then starts without a search or a barrier of its own
template <int W>
__global__ void __launch_bounds__(128) merge_partition_kernel(const MergeArgs<W> A, uint32_t tiles, uint32_t* __restrict__ split)
{
  const uint32_t cta = blockIdx.x * blockDim.x + threadIdx.x;
  if (cta >= tiles) return;
  const TileGeom<W> g = tile_geom<W>(A, cta);
  auto ga = [&](uint32_t i) { return A.in.get(g.a0 + i); };
  auto gb = [&](uint32_t i) { return A.in.get(g.b0 + i); };
  split[cta] = merge_path<W>(g.d0, g.len_a, g.len_b, ga, gb);
}

// shared memory of one merge_pass_kernel CTA: the staged input slices and a per-warp output staging area
constexpr int kOutTile = kTile + kTile / kItems;   // output staging, skewed by one slot per kItems (bank conflicts)
__device__ __forceinline__ uint32_t out_slot(uint32_t i) { return i + i / kItems; }
// STAGE_OUT: the warp's outputs go through its own slice of a second shared-memory buffer and leave coalesced;
// otherwise every thread stores its kItems consecutive outputs (64 contiguous bytes per array) straight from registers —
// half the shared memory per CTA, i.e. twice the resident warps, for partially filled store sectors.
template <int W, bool STAGE_OUT>
constexpr size_t merge_smem_bytes() { return static_cast<size_t>(kTile + (STAGE_OUT ? kOutTile : 0)) * 8 * (W + 1); }

// One tile of kTile outputs: stage both input slices in shared memory (coalesced), ONE block barrier, every thread
// finds its merge-path start in shared memory and merges kItems consecutive outputs in registers, each warp stages its
// 32 * kItems outputs in its own slice of a second buffer (warp barrier only) and stores them coalesced.
template <int W, bool STAGE_OUT>
__global__ void __launch_bounds__(kThreads) merge_pass_kernel(const MergeArgs<W> A, const uint32_t* __restrict__ split)
{
  extern __shared__ __align__(16) uint64_t smem[];
  uint64_t* s_lo = smem;
  uint64_t* s_hi = s_lo + (W == 2 ? kTile : 0);      // W == 1: aliases s_lo, never touched
  uint64_t* s_pay = s_hi + kTile;
  uint64_t* o_lo = s_pay + kTile;
  uint64_t* o_hi = o_lo + (W == 2 ? kOutTile : 0);
  uint64_t* o_pay = o_hi + kOutTile;

  const TileGeom<W> g = tile_geom<W>(A, blockIdx.x);
  const uint32_t ai0 = split[blockIdx.x];
  const uint32_t ai1 = g.d1 == g.len_a + g.len_b ? g.len_a : split[blockIdx.x + 1];
  const uint32_t bi0 = g.d0 - ai0, bi1 = g.d1 - ai1;
  const uint32_t na = ai1 - ai0, nb = bi1 - bi0, nt = na + nb;   // nt == d1 - d0 <= kTile

  // stage the two input slices: A's slice first, then B's
  const uint32_t src_a = g.a0 + ai0, src_b = g.b0 + bi0;
#pragma unroll
  for (int e = 0; e < kItems; e++)
  {
    const uint32_t i = threadIdx.x + e * kThreads;
    if (i < nt)
    {
      const uint32_t src = i < na ? src_a + i : src_b + (i - na);
      s_lo[i] = A.in.lo[src];
      if constexpr (W == 2) s_hi[i] = A.in.hi[src];
      s_pay[i] = A.pay_in[src];
    }
  }
  __syncthreads();

  auto sk = [&](uint32_t i) {
    Key<W> k;
    k.lo = s_lo[i];
    if constexpr (W == 2) k.hi = s_hi[i];
    return k;
  };
  auto sa = [&](uint32_t i) { return sk(i); };
  auto sb = [&](uint32_t i) { return sk(na + i); };

  const uint32_t t0 = min(threadIdx.x * kItems, nt);
  const uint32_t t1 = min(t0 + kItems, nt);
  uint32_t ia = merge_path<W>(t0, na, nb, sa, sb);
  uint32_t ib = t0 - ia;
  // the two heads live in registers; every step reloads only the side that advanced
  Key<W> ka{}, kb{};
  if (ia < na) ka = sa(ia);
  if (ib < nb) kb = sb(ib);
  const uint32_t o0 = g.a0 + g.d0;
#pragma unroll
  for (int e = 0; e < kItems; e++)
  {
    if (t0 + e < t1)
    {
      const bool take_a = ib >= nb || (ia < na && ka <= kb);
      const uint32_t src = take_a ? ia : na + ib;
      if constexpr (STAGE_OUT)
      {
        // thread t owns outputs [t * kItems, (t + 1) * kItems): slot stride kItems + 1 between lanes, conflict-free
        const uint32_t o = t0 + t0 / kItems + e;
        o_lo[o] = take_a ? ka.lo : kb.lo;
        if constexpr (W == 2) o_hi[o] = take_a ? ka.hi : kb.hi;
        o_pay[o] = s_pay[src];
      }
      else
      {
        A.out.lo[o0 + t0 + e] = take_a ? ka.lo : kb.lo;
        if constexpr (W == 2) A.out.hi[o0 + t0 + e] = take_a ? ka.hi : kb.hi;
        A.pay_out[o0 + t0 + e] = s_pay[src];
      }
      if (take_a) { ia++; if (ia < na) ka = sa(ia); }
      else { ib++; if (ib < nb) kb = sb(ib); }
    }
  }
  if constexpr (STAGE_OUT)
  {
    __syncwarp();
    // the warp's 32 * kItems outputs are contiguous in the output staging area: coalesced stores
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w0 = (threadIdx.x & ~31u) * kItems;
#pragma unroll
    for (int e = 0; e < kItems; e++)
    {
      const uint32_t i = w0 + e * 32 + lane;
      if (i < nt)
      {
        const uint32_t o = out_slot(i);
        A.out.lo[o0 + i] = o_lo[o];
        if constexpr (W == 2) A.out.hi[o0 + i] = o_hi[o];
        A.pay_out[o0 + i] = o_pay[o];
      }
    }
  }
}

// ---- row numbering ------------------------------------------------------------------------------

// flag[i] = 1 when record i opens a run of equal k-mers of which at least one sample is solid
// (recurrence >= 1, merge.hpp:239-240); tile_sum[cta] = flags of the CTA's kScanTile records.
template <int W>
__global__ void __launch_bounds__(256) heads_kernel(KeyArrays<W> keys, const uint64_t* __restrict__ pay, uint32_t n,
                                                    uint32_t* __restrict__ flag, uint32_t* __restrict__ tile_sum)
{
  __shared__ uint32_t s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kScanTile;
  uint32_t mine = 0;
  for (uint32_t i = base + threadIdx.x; i < min(n, base + kScanTile); i += blockDim.x)
  {
    const Key<W> k = keys.get(i);
    uint32_t f = 0;
    if (i == 0 || !(keys.get(i - 1) == k))
    {
      bool any = false;
      for (uint32_t j = i; j < n && !any; j++)
      {
        if (j > i && !(keys.get(j) == k)) break;
        any = pay_solid(pay[j]);
      }
      f = any;
    }
    flag[i] = f;
    mine += f;
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_sum, mine);
  __syncthreads();
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = s_sum;
}

// head_pos[row] = index of the record that opens row `row` (tile_off = exclusive scan of tile_sum)
__global__ void __launch_bounds__(256) head_positions_kernel(const uint32_t* __restrict__ flag, uint32_t n,
                                                             const uint32_t* __restrict__ tile_off,
                                                             uint32_t* __restrict__ head_pos)
{
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_run;
  if (threadIdx.x == 0) s_run = tile_off[blockIdx.x];
  __syncthreads();
  const uint32_t base = blockIdx.x * kScanTile;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t c = 0; c < kScanTile; c += 256)
  {
    const uint32_t i = base + c + threadIdx.x;
    const uint32_t f = i < n ? flag[i] : 0;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_run;
    for (uint32_t w = 0; w < warp; w++) before += s_warp[w];
    if (f) head_pos[before + __popc(bal & ((1u << lane) - 1))] = i;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      uint32_t t = 0;
      for (int w = 0; w < 8; w++) t += s_warp[w];
      s_run += t;
    }
    __syncthreads();
  }
}

// ---- dense rows -----------------------------------------------------------------------------------

// One warp per row: k-mer words, then N counts of B bytes (0 unless the sample carries the k-mer with a
// solid count), i.e. the row km::MatrixWriter::write would emit (matrix_file.hpp:121-127).
template <int W, int B>
__global__ void __launch_bounds__(256) rows_kernel(KeyArrays<W> keys, const uint64_t* __restrict__ pay, uint32_t n,
                                                   const uint32_t* __restrict__ head_pos, uint32_t n_rows,
                                                   uint32_t n_samples, uint8_t* __restrict__ rows)
{
  using unit_t = typename std::conditional<B == 4, uint32_t, typename std::conditional<B == 2, uint16_t, uint8_t>::type>::type;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t kmer_units = 8 * W / B;
  const size_t row_units = kmer_units + n_samples;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps)
  {
    const uint32_t i = head_pos[r];
    const Key<W> k = keys.get(i);
    unit_t* out = reinterpret_cast<unit_t*>(rows) + static_cast<size_t>(r) * row_units;
    for (uint32_t u = lane; u < row_units; u += 32)
    {
      unit_t v = 0;
      if (u < kmer_units)
      {
        const uint32_t byte = u * B;
        uint64_t word = k.lo;
        if constexpr (W == 2) word = byte < 8 ? k.lo : k.hi;
        v = static_cast<unit_t>(word >> (8 * (byte & 7)));
      }
      out[u] = v;
    }
    __syncwarp();
    for (uint32_t base = i;; base += 32)
    {
      const uint32_t j = base + lane;
      const bool same = j < n && keys.get(j) == k;
      if (same)
      {
        const uint64_t p = pay[j];
        if (pay_solid(p)) out[kmer_units + pay_stream(p)] = static_cast<unit_t>(pay_count(p));
      }
      if (!__all_sync(0xFFFFFFFFu, same)) break;
    }
  }
}

}  // namespace mrg
}  // namespace kmd
