// Device-side N-way merge of per-sample sorted (k-mer, count) streams, second form: the k-mers are merged, the
// counts never move.
//
// Replaces km::KmerMerger::next (reference thirdparty/kmtricks/include/kmtricks/merge.hpp:183-260) like
// kmd_merge.cuh does, for the common case in which every record is solid (count >= abundance_min of its sample:
// what `kmtricks count --hard-min` writes).  The pairwise tree of kmd_merge.cuh carries (key, payload) pairs through
// ceil(log2 N) full passes; but a dense row only needs to know WHERE each k-mer goes.  So:
//
//   1. the leaves: the records themselves when the counts have 4 bytes (RawKeys: the first level reads the k-mer words
//      where they lie), else keys_kernel unpacks the k-mer words first, SoA;
//   2. level_partition_kernel + union_level_kernel x ceil(log2 N): pairwise merge of sorted DISTINCT key runs that drops
//      the k-mer both sides hold.  Cohort samples share most k-mers, so a run is hardly longer than its inputs and the
//      levels shrink (N = 20 streams of ~n keys: 20n keys in, then 10n, 5n, 3n, 2n — 2 full passes' worth instead of 5,
//      keys only); the root is U, the sorted k-mers of the rows.  Output sizes are data dependent: a run starts where the
//      host knows (the sum of the leaf lengths before it), is as long as the level below says, and a tile learns its
//      output offset inside its run by a decoupled look-back over the tiles before it (single pass, no second read);
//      tiles are handed out by a ticket, round robin over the pairs (tile_order_kernel);
//   3. scatter_partition_kernel (one lower bound per row tile and stream), then assemble_rows_kernel (dense cohorts: a
//      CTA assembles its rows in shared memory — their k-mers, zeros, and every stream's counts that fall into them,
//      each record's row found by a binary search narrowed by the record's rank in its stream — and writes them whole)
//      or scatter_rows_kernel (sparse cohorts: 4-byte stores into rows zeroed by a memset) — kmtricks matrix layout
//      (matrix_file.hpp:121-127), ready for lrt_scan_kernel.
//
// A record below its sample's abundance_min counts as absent for the row's existence (merge.hpp:204-223, 239-240);
// the leaf level (or keys_kernel) reports whether there is one and the caller then takes the pairwise tree of
// kmd_merge.cuh instead.
#pragma once

#include "kmd_merge.cuh"

namespace kmd {
namespace uni {

using mrg::Key;
using mrg::KeyArrays;
using mrg::merge_path;

constexpr int kItems = 8;                       // merged positions per thread of a level: a tile is threads * kItems
constexpr int kThreads = 256;                   // threads per CTA of a level (128 and 512 measured: within 10 %, profiles/r02u)
constexpr int kRowTile = 2048;                  // most rows per CTA of the scatter / the row assembly
constexpr int kMaxFlat = 2048;                  // most streams assemble_rows_kernel keeps a prefix sum for

// ---- the k-mer words of record i, read where the record lies (4-byte counts: records are whole 32-bit words) ----

template <int W>
struct RawKeys
{
  static constexpr bool kLeaf = true;
  const uint32_t* w;   // the streams' records back to back, 2W + 1 words each
  __device__ Key<W> get(uint32_t i) const
  {
    const uint32_t* r = w + static_cast<size_t>(i) * (2 * W + 1);
    Key<W> k;
    k.lo = r[0] | (static_cast<uint64_t>(r[1]) << 32);
    if constexpr (W == 2) k.hi = r[2] | (static_cast<uint64_t>(r[3]) << 32);
    return k;
  }
  __device__ uint32_t count(uint32_t i) const { return w[static_cast<size_t>(i) * (2 * W + 1) + 2 * W]; }
};

template <int W>
struct SoaKeys : KeyArrays<W>
{
  static constexpr bool kLeaf = false;
  __device__ uint32_t count(uint32_t) const { return 0; }
};

// ---- records -> k-mer words ---------------------------------------------------------------------

struct KeysArgs
{
  const uint8_t* raw;          // the streams' records back to back
  const uint64_t* stream_off;  // [n_streams + 1] record offsets
  const uint32_t* ab_min;      // [n_streams]
  uint32_t n_streams;
  uint32_t n;
  uint32_t kmer_words, count_bytes;
  uint32_t ab_min_max;         // max over ab_min: a count at or above it is solid whatever its sample
  uint64_t* klo;
  uint64_t* khi;
  uint32_t* nonsolid;          // set to 1 when a record is below its sample's abundance_min
};

__global__ void __launch_bounds__(256) keys_kernel(const KeysArgs A)
{
  const uint32_t rec = 8 * A.kmer_words + A.count_bytes;
  const bool words = (rec & 3u) == 0;   // 4-byte counts: the record is a whole number of aligned 32-bit words
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < A.n; j += gridDim.x * blockDim.x)
  {
    uint64_t lo, hi = 0;
    uint32_t c;
    if (words)
    {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(A.raw) + static_cast<size_t>(j) * (rec >> 2);
      lo = w[0] | (static_cast<uint64_t>(w[1]) << 32);
      if (A.kmer_words == 2) hi = w[2] | (static_cast<uint64_t>(w[3]) << 32);
      c = w[2 * A.kmer_words];
    }
    else
    {
      const uint8_t* r = A.raw + static_cast<size_t>(j) * rec;
      lo = mrg::load_le(r, 8);
      if (A.kmer_words == 2) hi = mrg::load_le(r + 8, 8);
      c = static_cast<uint32_t>(mrg::load_le(r + 8 * A.kmer_words, A.count_bytes));
    }
    A.klo[j] = lo;
    if (A.kmer_words == 2) A.khi[j] = hi;
    if (c < A.ab_min_max)
    {
      // rare: look the sample up (last s with stream_off[s] <= j) for the exact test
      uint32_t l = 0, h = A.n_streams;
      while (h - l > 1)
      {
        const uint32_t mid = (l + h) >> 1;
        if (A.stream_off[mid] <= j) l = mid; else h = mid;
      }
      if (c < A.ab_min[l]) *A.nonsolid = 1u;
    }
  }
}

// ---- one level of the union tree ------------------------------------------------------------------

// What level_partition_kernel knows about a tile: its pair's runs and the merge-path split at the tile's first
// position.  The split at its end is the next tile's.
struct __align__(16) TileDesc
{
  uint32_t a0, b0;    // where the pair's runs start
  uint32_t la, lb;    // their lengths
  uint32_t ai0;       // keys of run A before the tile's first merged position
  uint32_t pair;
  uint32_t lt;        // index of the tile inside its pair: first position = lt * tile size
  uint32_t valid;     // 0: the pair ends before this tile (the host's tile count is an upper bound)
};

// Input run r of the level covers [run_base[r], run_base[r] + run_len[r]): the base is what the host knows (the sum of
// the leaf lengths before it), the length is what the level below produced.  Pair q = runs 2q and 2q + 1 (a lone last
// run is a pair with an empty right side) -> output run q at run_base[2q], length out_len[q].
// pair_tile[q] = first tile of pair q, from the host's upper bound of the pair's length.
// In = RawKeys (the first level reads the k-mers inside the records and checks that every count is solid: the records
// of run r belong to sample r) or SoaKeys (the output of the level below).
template <int W, class In>
struct LevelArgs
{
  In in;
  KeyArrays<W> out;
  const uint32_t* ab_min;        // RawKeys only: [n_runs] abundance_min of the samples, their maximum, and the flag
  uint32_t ab_min_max;           //   that is set when a count lies below its sample's
  uint32_t* nonsolid;
  const uint32_t* run_base;
  const uint32_t* run_len;
  const uint32_t* pair_tile;
  uint32_t* out_len;             // zero before the launch
  uint32_t n_runs, n_pairs, tiles;
  uint32_t tile;                 // merged positions per tile = threads per CTA * kItems
  TileDesc* desc;                // [tiles]
  unsigned long long* status;    // [tiles], zero before the launch: flag << 32 | keys (flag 1 = the tile's own count,
                                 // 2 = the count of the tile and of every tile of the pair before it)
  uint32_t* ticket;              // zero before the launch
  const uint32_t* order;         // [tiles] ticket -> tile (tile_order_kernel): round robin over the pairs, so that the tiles in
                                 // flight belong to many runs and a tile's predecessors (same run) are long done when it
                                 // looks back; null = tiles in index order
};

// order[ticket] = tile: the pairs' first tiles, then their second ones, ... (round robin over the pairs that still have
// tiles).  One thread per tile; its ticket = the tiles with a smaller index inside their pair, plus the pairs before its
// own that reach its index.
__global__ void __launch_bounds__(256) tile_order_kernel(const uint32_t* __restrict__ pair_tile, uint32_t n_pairs, uint32_t tiles,
                                                         uint32_t* __restrict__ order)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= tiles) return;
  uint32_t ql = 0, qh = n_pairs;
  while (qh - ql > 1)
  {
    const uint32_t mid = (ql + qh) >> 1;
    if (pair_tile[mid] <= g) ql = mid; else qh = mid;
  }
  const uint32_t lt = g - pair_tile[ql];
  uint32_t ticket = 0;
  for (uint32_t q = 0; q < n_pairs; q++)
  {
    const uint32_t cnt = pair_tile[q + 1] - pair_tile[q];
    ticket += min(cnt, lt) + (q < ql && cnt > lt ? 1u : 0u);
  }
  order[ticket] = g;
}

// merge_path by a whole warp: 32 probes per step instead of one (the probes are dependent loads from global memory,
// ~0.7 us each: 5 steps instead of 21 for runs of a million keys)
template <int W, class GetA, class GetB>
__device__ inline uint32_t merge_path_warp(uint32_t diag, uint32_t len_a, uint32_t len_b, GetA a, GetB b)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t lo = diag > len_b ? diag - len_b : 0, hi = diag < len_a ? diag : len_a;
  // a(mid) <= b(diag - 1 - mid) holds for every mid below the answer and for none from it on
  while (lo < hi)
  {
    const uint32_t width = hi - lo;
    const uint32_t mid = width <= 32 ? lo + lane : lo + static_cast<uint32_t>(static_cast<uint64_t>(width) * (lane + 1) / 33);
    const bool valid = mid < hi;
    const bool below = valid && (a(mid) <= b(diag - 1 - mid));
    const uint32_t k = __popc(__ballot_sync(0xFFFFFFFFu, below));       // probes 0 .. k-1 lie below the answer
    const uint32_t n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    const uint32_t p_last = __shfl_sync(0xFFFFFFFFu, mid, k ? k - 1 : 0);
    const uint32_t p_next = __shfl_sync(0xFFFFFFFFu, mid, k < 32 ? k : 31);
    if (k < n_valid) hi = p_next;
    if (k) lo = p_last + 1;
  }
  return lo;
}

// one warp per tile: the pair it belongs to and the merge-path split at its first position
template <int W, class In>
__global__ void __launch_bounds__(128) level_partition_kernel(const LevelArgs<W, In> A)
{
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= A.tiles) return;
  uint32_t ql = 0, qh = A.n_pairs;
  while (qh - ql > 1)
  {
    const uint32_t mid = (ql + qh) >> 1;
    if (A.pair_tile[mid] <= t) ql = mid; else qh = mid;
  }
  const uint32_t q = ql;
  const bool has_b = 2 * q + 1 < A.n_runs;
  TileDesc d{};
  d.pair = q;
  d.lt = t - A.pair_tile[q];
  d.a0 = A.run_base[2 * q];
  d.la = A.run_len[2 * q];
  d.b0 = has_b ? A.run_base[2 * q + 1] : 0;
  d.lb = has_b ? A.run_len[2 * q + 1] : 0;
  const uint32_t d0 = d.lt * A.tile;
  if (d0 < d.la + d.lb)
  {
    d.valid = 1;
    auto ga = [&](uint32_t i) { return A.in.get(d.a0 + i); };
    auto gb = [&](uint32_t i) { return A.in.get(d.b0 + i); };
    d.ai0 = merge_path_warp<W>(d0, d.la, d.lb, ga, gb);
  }
  if ((threadIdx.x & 31) == 0) A.desc[t] = d;
}

__device__ __forceinline__ uint32_t stage_slot(uint32_t i) { return i + i / kItems; }

// One tile: both slices into shared memory, every thread merges kItems consecutive positions of the merged order in
// registers and keeps a B key only when run A does not hold it (ties put A first, so the twin is the A key right
// before it), block-wide prefix of the kept counts, look-back for the tile's offset inside the output run, coalesced
// stores.  Tiles are handed out by a ticket, earlier tiles of a run before later ones, so a tile that is waited for
// has already started.
template <int W, int kThreads, class In>
__global__ void __launch_bounds__(kThreads) union_level_kernel(const LevelArgs<W, In> A)
{
  constexpr int kTile = kThreads * kItems;        // merged inputs (A slice + B slice) of the CTA
  constexpr int kStage = kTile + kTile / kItems;  // shared key slots: inputs (+1 halo), then the skewed outputs
  __shared__ __align__(16) uint64_t s_lo[kStage];
  __shared__ __align__(16) uint64_t s_hi[W == 2 ? kStage : 1];
  __shared__ uint32_t s_warp[kThreads / 32];
  __shared__ uint32_t s_tile, s_prefix;

  if (threadIdx.x == 0)
  {
    const uint32_t ticket = atomicAdd(A.ticket, 1u);
    s_tile = A.order ? A.order[ticket] : ticket;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const TileDesc d = A.desc[tile];
  if (!d.valid) return;
  const uint32_t run_total = d.la + d.lb, d0 = d.lt * kTile, d1 = min(d0 + kTile, run_total);
  const bool last = d1 == run_total;
  const uint32_t ai1 = last ? d.la : A.desc[tile + 1].ai0;   // not the pair's last tile: the next one is valid and of this pair
  const uint32_t a_src = d.a0 + d.ai0, b_src = d.b0 + (d0 - d.ai0);
  const uint32_t na = ai1 - d.ai0, nb = (d1 - ai1) - (d0 - d.ai0), nt = na + nb;
  const bool halo = d.ai0 > 0;

  // slot 0 = the A key before the slice, then A's slice, then B's
#pragma unroll
  for (int e = 0; e < kItems; e++)
  {
    const uint32_t i = threadIdx.x + e * kThreads;
    if (i < nt)
    {
      const uint32_t src = i < na ? a_src + i : b_src + (i - na);
      const Key<W> k = A.in.get(src);
      s_lo[1 + i] = k.lo;
      if constexpr (W == 2) s_hi[1 + i] = k.hi;
      if constexpr (In::kLeaf)
      {
        // merge.hpp:204-223: a count below the sample's abundance_min makes the record absent — not this tree's case
        const uint32_t c = A.in.count(src);
        if (c < A.ab_min_max && c < A.ab_min[2 * d.pair + (i < na ? 0 : 1)]) *A.nonsolid = 1u;
      }
    }
  }
  if (threadIdx.x == 0 && halo)
  {
    const Key<W> k = A.in.get(a_src - 1);
    s_lo[0] = k.lo;
    if constexpr (W == 2) s_hi[0] = k.hi;
  }
  __syncthreads();

  auto sk = [&](uint32_t slot) {
    Key<W> k;
    k.lo = s_lo[slot];
    if constexpr (W == 2) k.hi = s_hi[slot];
    return k;
  };
  auto sa = [&](uint32_t i) { return sk(1 + i); };
  auto sb = [&](uint32_t i) { return sk(1 + na + i); };

  const uint32_t t0 = min(threadIdx.x * kItems, nt);
  const uint32_t t1 = min(t0 + kItems, nt);
  uint32_t ia = merge_path<W>(t0, na, nb, sa, sb);
  uint32_t ib = t0 - ia;
  Key<W> ka{}, kb{}, pa{};
  bool pa_ok = ia > 0 || halo;   // pa = the last A key before the current position
  if (pa_ok) pa = sk(ia);
  if (ia < na) ka = sa(ia);
  if (ib < nb) kb = sb(ib);
  Key<W> outk[kItems];
  uint32_t keep = 0;
#pragma unroll
  for (int e = 0; e < kItems; e++)
  {
    outk[e] = Key<W>{};
    if (t0 + e < t1)
    {
      const bool take_a = ib >= nb || (ia < na && ka <= kb);
      if (take_a)
      {
        outk[e] = ka;
        keep |= 1u << e;
        pa = ka;
        pa_ok = true;
        ia++;
        if (ia < na) ka = sa(ia);
      }
      else
      {
        if (!(pa_ok && pa == kb)) { outk[e] = kb; keep |= 1u << e; }
        ib++;
        if (ib < nb) kb = sb(ib);
      }
    }
  }

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t cnt = __popc(keep);
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= static_cast<uint32_t>(o)) inc += v;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();   // every thread is done with the staged inputs
  uint32_t before = 0, total = 0;
#pragma unroll
  for (uint32_t w = 0; w < kThreads / 32; w++)
  {
    const uint32_t v = s_warp[w];
    if (w < warp) before += v;
    total += v;
  }
  // the kept keys, compacted, into the same shared memory (skewed: a thread's keys are kItems apart from its neighbour's)
  uint32_t o = before + inc - cnt;
#pragma unroll
  for (int e = 0; e < kItems; e++)
  {
    if ((keep >> e) & 1u)
    {
      const uint32_t slot = stage_slot(o);
      s_lo[slot] = outk[e].lo;
      if constexpr (W == 2) s_hi[slot] = outk[e].hi;
      o++;
    }
  }
  if (warp == 0)
  {
    // keys of the pair's earlier tiles: their own counts (flag 1) back to the nearest tile that already knows its
    // prefix (flag 2; the pair's first tile always does), 32 tiles per step
    volatile unsigned long long* st = A.status;
    uint32_t prefix = 0;
    if (d.lt != 0)
    {
      if (lane == 0) st[tile] = (1ull << 32) | total;
      for (uint32_t back = 1;; back += 32)
      {
        const uint32_t dist = back + lane;
        const bool in_pair = dist <= d.lt;
        unsigned long long v = 0;
        if (in_pair)
          while (((v = st[tile - dist]) >> 32) == 0) __nanosleep(64);
        const uint32_t closed = __ballot_sync(0xFFFFFFFFu, in_pair && (v >> 32) == 2);
        const uint32_t upto = closed ? __ffs(closed) - 1 : 31;
        uint32_t part = (in_pair && lane <= upto) ? static_cast<uint32_t>(v) : 0u;
#pragma unroll
        for (int sh = 16; sh; sh >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, sh);
        prefix += part;
        if (closed) break;
      }
    }
    if (lane == 0)
    {
      st[tile] = (2ull << 32) | (prefix + total);
      if (last) A.out_len[d.pair] = prefix + total;
      s_prefix = prefix;
    }
  }
  __syncthreads();
  const uint32_t out0 = d.a0 + s_prefix;
  for (uint32_t i = threadIdx.x; i < total; i += kThreads)
  {
    const uint32_t slot = stage_slot(i);
    A.out.lo[out0 + i] = s_lo[slot];
    if constexpr (W == 2) A.out.hi[out0 + i] = s_hi[slot];
  }
}

// ---- counts -> dense rows -----------------------------------------------------------------------------

template <int W>
struct ScatterArgs
{
  KeyArrays<W> leaf;           // the streams' keys, back to back (keys_kernel; not with 4-byte counts: they are read in the records)
  KeyArrays<W> rows_key;       // U: the k-mer of every row, ascending
  const uint8_t* raw;
  const uint64_t* stream_off;  // [n_streams + 1]
  const uint32_t* ab_min;      // [n_streams]
  uint32_t* split;             // [(tiles + 1) * n_streams]: records of stream s before the first row of tile b
  uint32_t n_streams, n_rows, tiles;
  uint32_t row_tile;           // rows per CTA (<= kRowTile; a multiple of 16)
  uint32_t rec, row_bytes;
  uint8_t* rows;               // n_rows * row_bytes, zero before the launch
  uint32_t* err;               // |= 4 when a record's k-mer is not in its tile (never: it would be a bug in the tree)
};

// split[b][s] = lower bound of the first k-mer of row tile b in stream s (b == tiles: the stream's length)
template <int W, class Leaf>
__global__ void __launch_bounds__(256) scatter_partition_kernel(const ScatterArgs<W> A, const Leaf leaf)
{
  const uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<uint64_t>(A.tiles + 1) * A.n_streams) return;
  const uint32_t b = static_cast<uint32_t>(idx / A.n_streams), s = static_cast<uint32_t>(idx % A.n_streams);
  const uint32_t off = static_cast<uint32_t>(A.stream_off[s]);
  const uint32_t len = static_cast<uint32_t>(A.stream_off[s + 1]) - off;
  uint32_t lo = 0, hi = len;
  if (b < A.tiles)
  {
    const Key<W> key = A.rows_key.get(b * A.row_tile);
    while (lo < hi)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (key <= leaf.get(off + mid)) hi = mid; else lo = mid + 1;
    }
  }
  else lo = len;
  A.split[idx] = lo;
}

template <int W, int B>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const ScatterArgs<W> A)
{
  using unit_t = typename std::conditional<B == 4, uint32_t, typename std::conditional<B == 2, uint16_t, uint8_t>::type>::type;
  __shared__ __align__(16) uint64_t s_lo[kRowTile];
  __shared__ __align__(16) uint64_t s_hi[W == 2 ? kRowTile : 1];
  __shared__ uint32_t s_total;

  const uint32_t N = A.n_streams;
  const uint32_t r0 = blockIdx.x * A.row_tile, rt = min(A.row_tile, A.n_rows - r0);
  const uint32_t* sp0 = A.split + static_cast<size_t>(blockIdx.x) * N;
  const uint32_t* sp1 = sp0 + N;
  constexpr uint32_t kmer_units = 8 * W / B;
  const size_t row_units = A.row_bytes / B;
  unit_t* rows = reinterpret_cast<unit_t*>(A.rows);

  if (threadIdx.x == 0) s_total = 0;
  __syncthreads();
  // the tile's k-mers: into shared memory and into the rows
  for (uint32_t i = threadIdx.x; i < rt; i += blockDim.x)
  {
    const Key<W> k = A.rows_key.get(r0 + i);
    s_lo[i] = k.lo;
    if constexpr (W == 2) s_hi[i] = k.hi;
    unit_t* out = rows + static_cast<size_t>(r0 + i) * row_units;
#pragma unroll
    for (uint32_t u = 0; u < kmer_units; u++)
    {
      const uint32_t byte = u * B;
      uint64_t word = k.lo;
      if constexpr (W == 2) word = byte < 8 ? k.lo : k.hi;
      out[u] = static_cast<unit_t>(word >> (8 * (byte & 7)));
    }
  }
  uint32_t mine = 0;
  for (uint32_t s = threadIdx.x; s < N; s += blockDim.x) mine += sp1[s] - sp0[s];
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_total, mine);
  __syncthreads();

  auto row_key = [&](uint32_t i) {
    Key<W> k;
    k.lo = s_lo[i];
    if constexpr (W == 2) k.hi = s_hi[i];
    return k;
  };
  // record j of stream s (its a-th .. e-th records fall into this tile)
  auto place = [&](uint32_t s, uint32_t j, uint32_t a, uint32_t e, uint32_t off, uint32_t ab) {
    Key<W> key;
    uint32_t c;
    if constexpr (B == 4)
    {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(A.raw) + static_cast<size_t>(off + j) * (2 * W + 1);
      key.lo = w[0] | (static_cast<uint64_t>(w[1]) << 32);
      if constexpr (W == 2) key.hi = w[2] | (static_cast<uint64_t>(w[3]) << 32);
      c = w[2 * W];
    }
    else
    {
      key = A.leaf.get(off + j);
      c = *reinterpret_cast<const unit_t*>(A.raw + static_cast<size_t>(off + j) * A.rec + 8 * W);
    }
    // the stream's keys are distinct and all among the tile's: j - a of them lie before this one, e - 1 - j after
    uint32_t lo = j - a, hi = rt - (e - j);
    if (hi >= rt || lo > hi) { atomicOr(A.err, 4u); return; }
    while (lo < hi)
    {
      const uint32_t mid = (lo + hi) >> 1;
      if (key <= row_key(mid)) hi = mid; else lo = mid + 1;
    }
    if (!(row_key(lo) == key)) { atomicOr(A.err, 4u); return; }
    if (c >= ab) rows[static_cast<size_t>(r0 + lo) * row_units + kmer_units + s] = static_cast<unit_t>(c);
  };

  if (s_total >= 64u * N)
  {
    // dense tile: every stream's records are spread over the whole CTA
    for (uint32_t s = 0; s < N; s++)
    {
      const uint32_t a = sp0[s], e = sp1[s], off = static_cast<uint32_t>(A.stream_off[s]), ab = A.ab_min[s];
      for (uint32_t j = a + threadIdx.x; j < e; j += blockDim.x) place(s, j, a, e, off, ab);
    }
  }
  else
  {
    const uint32_t lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    for (uint32_t s = threadIdx.x >> 5; s < N; s += warps)
    {
      const uint32_t a = sp0[s], e = sp1[s];
      if (a == e) continue;
      const uint32_t off = static_cast<uint32_t>(A.stream_off[s]), ab = A.ab_min[s];
      for (uint32_t j = a + lane; j < e; j += 32) place(s, j, a, e, off, ab);
    }
  }
}


// The same rows, assembled in shared memory and written once, whole (dense cohorts: scattered 4-byte stores into rows
// that do not all stay in L2 until they are complete cost a read-modify-write of every sector).  A CTA owns row_tile
// rows (as many as fit the shared-memory budget), keeps the prefix sum of the streams' record counts inside the tile
// and walks the tile's records as one flat list — balanced whatever the streams' densities.
// Dynamic shared memory: rows [row_tile * row_bytes, 16-aligned] | key words [row_tile * W] | prefix [N + 1] | first record [N].
template <int W>
__host__ __device__ inline size_t assemble_smem_bytes(uint32_t row_tile, uint32_t row_bytes, uint32_t n_streams)
{
  const size_t rows = (static_cast<size_t>(row_tile) * row_bytes + 15) / 16 * 16;
  return rows + static_cast<size_t>(row_tile) * 8 * W + (2ull * n_streams + 2) * 4;
}

template <int W, int B>
__global__ void __launch_bounds__(256) assemble_rows_kernel(const ScatterArgs<W> A)
{
  using unit_t = typename std::conditional<B == 4, uint32_t, typename std::conditional<B == 2, uint16_t, uint8_t>::type>::type;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ uint32_t s_w[8];

  const uint32_t N = A.n_streams, R = A.row_bytes;
  const uint32_t r0 = blockIdx.x * A.row_tile, rt = min(A.row_tile, A.n_rows - r0);
  const size_t rows_bytes = (static_cast<size_t>(A.row_tile) * R + 15) / 16 * 16;
  uint8_t* s_rows = smem_raw;
  uint64_t* s_lo = reinterpret_cast<uint64_t*>(smem_raw + rows_bytes);
  uint64_t* s_hi = s_lo + (W == 2 ? A.row_tile : 0);
  uint32_t* s_pre = reinterpret_cast<uint32_t*>(s_hi + A.row_tile);   // [N + 1] records of the streams before s, inside the tile
  uint32_t* s_rec = s_pre + N + 1;                                    // [N] index of the stream's first record of the tile - s_pre[s]
  const uint32_t* sp0 = A.split + static_cast<size_t>(blockIdx.x) * N;
  const uint32_t* sp1 = sp0 + N;
  constexpr uint32_t kmer_units = 8 * W / B;
  const uint32_t row_units = R / B;

  {
    uint4* z = reinterpret_cast<uint4*>(s_rows);
    for (uint32_t i = threadIdx.x; i < rows_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  for (uint32_t s = threadIdx.x; s < N; s += blockDim.x) s_pre[s] = sp1[s] - sp0[s];
  for (uint32_t i = threadIdx.x; i < rt; i += blockDim.x)
  {
    const Key<W> k = A.rows_key.get(r0 + i);
    s_lo[i] = k.lo;
    if constexpr (W == 2) s_hi[i] = k.hi;
  }
  __syncthreads();
  // exclusive prefix sum of the counts, in place: every thread owns `per` consecutive streams
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = (N + blockDim.x - 1) / blockDim.x;
  const uint32_t sb = min(threadIdx.x * per, N), se = min(sb + per, N);
  uint32_t local = 0;
  for (uint32_t s = sb; s < se; s++) local += s_pre[s];
  uint32_t inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= static_cast<uint32_t>(o)) inc += v;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t run = inc - local;
  for (uint32_t w = 0; w < warp; w++) run += s_w[w];
  for (uint32_t s = sb; s < se; s++)
  {
    const uint32_t c = s_pre[s];
    s_pre[s] = run;
    s_rec[s] = static_cast<uint32_t>(A.stream_off[s]) + sp0[s] - run;
    run += c;
  }
  if (threadIdx.x == blockDim.x - 1) s_pre[N] = run;
  // the k-mer words of the tile's rows
  for (uint32_t i = threadIdx.x; i < rt; i += blockDim.x)
  {
    unit_t* out = reinterpret_cast<unit_t*>(s_rows) + static_cast<size_t>(i) * row_units;
#pragma unroll
    for (uint32_t u = 0; u < kmer_units; u++)
    {
      const uint32_t byte = u * B;
      uint64_t word = s_lo[i];
      if constexpr (W == 2) word = byte < 8 ? s_lo[i] : s_hi[i];
      out[u] = static_cast<unit_t>(word >> (8 * (byte & 7)));
    }
  }
  __syncthreads();

  auto row_key = [&](uint32_t i) {
    Key<W> k;
    k.lo = s_lo[i];
    if constexpr (W == 2) k.hi = s_hi[i];
    return k;
  };
  const uint32_t M = s_pre[N];
  // kBatch records per thread at a time: their loads are in flight together, then each is placed
  constexpr int kBatch = 4;
  uint32_t walk = 0;   // stream of the thread's previous record
  for (uint32_t m0 = threadIdx.x; m0 < M; m0 += blockDim.x * kBatch)
  {
    Key<W> key[kBatch];
    uint32_t cnt[kBatch], str[kBatch], lo[kBatch], hi[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++)
    {
      const uint32_t m = m0 + u * blockDim.x;
      str[u] = 0xFFFFFFFFu;
      if (m < M)
      {
        // the stream of flat record m: last s with s_pre[s] <= m (an empty stream shares its prefix with the next
        // one).  A thread's records ascend, so it walks on from the stream of its previous record; a long walk
        // (many short streams) turns into a binary search.
        uint32_t sl = walk;
        if (s_pre[sl + 1] <= m)
        {
          uint32_t steps = 0;
          do { sl++; steps++; } while (steps < 4 && s_pre[sl + 1] <= m);
          if (s_pre[sl + 1] <= m)
          {
            uint32_t sh = N;
            while (sh - sl > 1)
            {
              const uint32_t mid = (sl + sh) >> 1;
              if (s_pre[mid] <= m) sl = mid; else sh = mid;
            }
          }
        }
        walk = sl;
        str[u] = sl;
        const uint32_t j = s_rec[sl] + m;   // record index in the raw buffer
        if constexpr (B == 4)
        {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(A.raw) + static_cast<size_t>(j) * (2 * W + 1);
          key[u].lo = w[0] | (static_cast<uint64_t>(w[1]) << 32);
          if constexpr (W == 2) key[u].hi = w[2] | (static_cast<uint64_t>(w[3]) << 32);
          cnt[u] = w[2 * W];
        }
        else
        {
          key[u] = A.leaf.get(j);
          cnt[u] = *reinterpret_cast<const unit_t*>(A.raw + static_cast<size_t>(j) * A.rec + 8 * W);
        }
        // the stream's keys are distinct and all among the tile's: m - s_pre[s] of them lie before this one, the rest after
        lo[u] = m - s_pre[sl];
        hi[u] = rt - (s_pre[sl + 1] - m);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++)
    {
      if (str[u] == 0xFFFFFFFFu) continue;
      uint32_t l = lo[u], h = hi[u];
      if (h >= rt || l > h) { atomicOr(A.err, 4u); continue; }
      while (l < h)
      {
        const uint32_t mid = (l + h) >> 1;
        if (key[u] <= row_key(mid)) h = mid; else l = mid + 1;
      }
      if (!(row_key(l) == key[u])) { atomicOr(A.err, 4u); continue; }
      if (cnt[u] >= A.ab_min[str[u]])
        reinterpret_cast<unit_t*>(s_rows)[static_cast<size_t>(l) * row_units + kmer_units + str[u]] = static_cast<unit_t>(cnt[u]);
    }
  }
  __syncthreads();
  // the tile's rows are contiguous in the matrix: whole 16-byte stores (the tile starts on a multiple of 16 rows)
  const size_t bytes = static_cast<size_t>(rt) * R;
  uint8_t* dst = A.rows + static_cast<size_t>(r0) * R;
  const uint4* src16 = reinterpret_cast<const uint4*>(s_rows);
  uint4* dst16 = reinterpret_cast<uint4*>(dst);
  for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst16[i] = src16[i];
  for (size_t i = bytes / 16 * 16 + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = s_rows[i];
}

}  // namespace uni
}  // namespace kmd
