// Seeded, counter-based generator of kmtricks-format count-matrix rows.
//
// Integer-only after the inverse-CDF tables are built (once, on the host), so host and device
// produce bit-identical rows for any (seed, row index): the CPU baseline, the oracle and the CUDA
// path can all be fed the same prefix of the same synthetic matrix without moving it around.
//
// Row r:  k-mer  = strictly increasing in r (rows of a kmtricks partition are sorted ascending,
//                  kmtricks/io/matrix_file.hpp + merge.hpp), 2 bits/base, top bits zero;
//         counts = Poisson(lambda_class(r)) per sample; for `enriched_ppm` rows per million one
//                  cohort (mostly the cases) is drawn from Poisson(40) instead;
//         never all-zero (kmtricks never emits such rows, kmtricks/merge.hpp:249-250).
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define KMD_SHD __host__ __device__ __forceinline__
#else
#define KMD_SHD inline
#endif

namespace kmd {
namespace synth {

constexpr int kClasses = 17;      // 16 background abundance classes + 1 enriched class
constexpr int kEnrichedClass = 16;
constexpr int kCdfLen = 96;       // counts are clipped at 95 (P[Poisson(40) > 95] < 1e-12)

// background lambdas: a coarse Gamma(2,3)-like spread of per-sample abundances
constexpr double kLambda[kClasses] = {0.4, 0.8, 1.2, 1.7, 2.3, 3.0, 3.8, 4.7, 5.7, 6.8,
                                      8.0, 9.5, 11.5, 14.0, 17.5, 23.0, 40.0};

struct Tables
{
  uint32_t cdf[kClasses][kCdfLen];  // cdf[c][i] = floor(P[X <= i] * 2^32), saturated
};

struct Params
{
  uint64_t seed;
  uint64_t kmer_stride;  // spacing of consecutive k-mers (top word when kmer_words == 2)
  uint32_t nb_controls;
  uint32_t nb_cases;
  uint32_t kmer_words;
  uint32_t count_bytes;
  uint32_t enriched_ppm;
  uint32_t row_bytes;
};

KMD_SHD uint64_t mix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// smallest i with u < cdf[i]  (binary search over the saturated table)
KMD_SHD uint32_t sample(const uint32_t* cdf, uint32_t u)
{
  uint32_t lo = 0, hi = kCdfLen - 1;
  while (lo < hi)
  {
    uint32_t mid = (lo + hi) >> 1;
    if (u < cdf[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Byte sink that packs into 8-byte little-endian words when the row start is 8-byte aligned and
// falls back to single-byte stores otherwise.
struct RowSink
{
  uint8_t* dst;
  uint64_t acc = 0;
  uint32_t fill = 0;
  bool wide;
  KMD_SHD RowSink(uint8_t* d) : dst(d), wide((reinterpret_cast<uintptr_t>(d) & 7u) == 0) {}
  KMD_SHD void push(uint32_t v, uint32_t nbytes)
  {
    if (wide)
    {
      acc |= static_cast<uint64_t>(v) << (8 * fill);
      fill += nbytes;
      if (fill == 8)
      {
        *reinterpret_cast<uint64_t*>(dst) = acc;
        dst += 8; acc = 0; fill = 0;
      }
    }
    else
      for (uint32_t b = 0; b < nbytes; b++) *dst++ = static_cast<uint8_t>(v >> (8 * b));
  }
  KMD_SHD void flush()
  {
    for (uint32_t b = 0; b < fill; b++) dst[b] = static_cast<uint8_t>(acc >> (8 * b));
  }
};

// Writes one row (p.row_bytes bytes) at `dst`.
KMD_SHD void make_row(const Params& p, const Tables& t, uint64_t r, uint8_t* dst)
{
  const uint64_t h0 = mix64(p.seed ^ (r * 0xD6E8FEB86659FD93ull));
  const uint64_t h1 = mix64(h0 ^ 0xA5A5A5A5A5A5A5A5ull);
  const uint32_t n = p.nb_controls + p.nb_cases;
  RowSink sink(dst);

  // k-mer
  const uint64_t top = r * p.kmer_stride + (p.kmer_stride > 1 ? h1 % p.kmer_stride : 0);
  if (p.kmer_words == 1)
  {
    sink.push(static_cast<uint32_t>(top), 4);
    sink.push(static_cast<uint32_t>(top >> 32), 4);
  }
  else
  {
    const uint64_t low = mix64(h1);
    sink.push(static_cast<uint32_t>(low), 4);
    sink.push(static_cast<uint32_t>(low >> 32), 4);
    sink.push(static_cast<uint32_t>(top), 4);
    sink.push(static_cast<uint32_t>(top >> 32), 4);
  }

  // abundance class, enrichment
  const uint32_t base_class = static_cast<uint32_t>(h0 & 15u);
  const bool enriched = ((h0 >> 8) % 1000000ull) < p.enriched_ppm;
  const bool enrich_controls = ((h0 >> 40) & 7u) == 0;  // 1 in 8 enriched rows favours controls
  const uint32_t cmax = p.count_bytes == 1 ? 0xFFu : (p.count_bytes == 2 ? 0xFFFFu : 0xFFFFFFFFu);
  // if every draw comes out zero this sample gets a 1 (kmtricks never emits all-zero rows)
  const uint32_t forced = static_cast<uint32_t>(h1 % n);

  bool any = false;
  for (uint32_t j = 0; j < n; j++)
  {
    const bool is_ctrl = j < p.nb_controls;
    const uint32_t cls = (enriched && (is_ctrl == enrich_controls)) ? kEnrichedClass : base_class;
    const uint32_t u = static_cast<uint32_t>(mix64(h0 + 0x632BE59BD9B4E019ull * (j + 1)) >> 32);
    uint32_t v = sample(t.cdf[cls], u);
    if (v > cmax) v = cmax;
    any |= v != 0;
    if (j == n - 1 && !any && forced == j) v = 1;
    sink.push(v, p.count_bytes);
  }
  sink.flush();
  if (!any && forced != n - 1) dst[8 * p.kmer_words + forced * p.count_bytes] = 1;
}

}  // namespace synth
}  // namespace kmd
