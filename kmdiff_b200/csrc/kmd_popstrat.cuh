// Population-stratification kernels for sm_100a (`kmdiff diff --pop-correction`).
//
// Replaces (reference file:line):
//   Sampler::sample                     include/kmdiff/popstrat.hpp:107-146  (std::default_random_engine ==
//                                       minstd_rand0 + uniform_real_distribution<double>, one draw per row
//                                       in row-visit order)
//   EigGenoFile::push                   include/kmdiff/popstrat.hpp:56-80    (presence = count > 0)
//   glm_irls / lu / inverse / sigmoid   src/linear_model.cpp:297-410, 94-189, 191-195
//   pop_strat_corrector::apply(ks)      include/kmdiff/popstrat.hpp:249-333
//
// The sampler needs no memory traffic at all: the state of the LCG before the draw of global row r is
// s0 * a^(2r) mod (2^31-1) (jump-ahead), so every thread owns a run of consecutive rows and emits the
// indices of the sampled ones.  The GLM is one warp per surviving k-mer: every floating-point sum is
// accumulated in the order the reference accumulates it (lanes own whole output elements, never parts
// of a sum), without FMA contraction, so the fit follows the reference iteration by iteration.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kmd_math.cuh"

namespace kmd {
namespace pop {

// ---- minstd_rand0 -------------------------------------------------------------------------------

constexpr uint32_t kLcgA = 16807u;
constexpr uint32_t kLcgM = 2147483647u;  // 2^31 - 1

// x * a mod (2^31 - 1) for x, a in [1, M)
KMD_HD uint32_t lcg_mul(uint32_t x, uint32_t a)
{
  const uint64_t p = static_cast<uint64_t>(x) * a;
  uint64_t y = (p & kLcgM) + (p >> 31);
  if (y >= kLcgM) y -= kLcgM;
  return static_cast<uint32_t>(y);
}

// a^e mod (2^31 - 1)
KMD_HD uint32_t lcg_pow(uint32_t a, uint64_t e)
{
  e %= (kLcgM - 1u);  // Fermat: a^(M-1) == 1
  uint32_t r = 1, b = a;
  while (e)
  {
    if (e & 1) r = lcg_mul(r, b);
    b = lcg_mul(b, b);
    e >>= 1;
  }
  return r;
}

// linear_congruential_engine::seed (libstdc++): state = seed % m, 0 -> 1; the Sampler seeds only when
// seed != 0 (popstrat.hpp:113-117), a default-constructed engine starts at 1.
KMD_HD uint32_t lcg_seed(uint64_t seed)
{
  if (seed == 0) return 1u;
  const uint32_t s = static_cast<uint32_t>(seed % kLcgM);
  return s == 0 ? 1u : s;
}

// generate_canonical<double, 53> over minstd_rand0 (libstdc++ random.tcc): two engine calls,
// ((x1-1) + (x2-1)*R) / (R*R) with R = 2147483646, every operation rounded to double.
KMD_HD double canonical_draw(uint32_t x1, uint32_t x2)
{
  const double R = 2147483646.0;
  double sum = static_cast<double>(x1 - 1u);
  sum = add(sum, mul(static_cast<double>(x2 - 1u), R));
  const double tmp = mul(R, R);
  double ret = div(sum, tmp);
  if (ret >= 1.0) ret = 0.99999999999999988898;  // nextafter(1, 0)
  return ret;
}

struct SamplerArgs
{
  uint32_t state0;        // engine state before the draw of global row 0
  double v;               // --kmer-pca
  double v_scaled;        // v * R
  uint64_t first_row;     // global index (survivor/geno row tag) of row 0 of this batch
  uint64_t first_draw;    // index of the draw of row 0 of this batch (== first_row unless overridden)
  uint64_t n_rows;
  uint32_t rows_per_thread;
  unsigned long long* total;  // number of sampled rows so far (all batches)
  uint64_t* out_rows;     // global row index of each sampled row, in slot order
  uint64_t cap;
};

// Sampler::sample() for the rows of one batch.  A sampled row takes the next free slot; slots are
// put into row order later (the .geno export), the Gram matrix does not depend on the order.
__global__ void __launch_bounds__(256) sampler_kernel(const SamplerArgs A)
{
  const uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t r0 = t * A.rows_per_thread;
  if (r0 >= A.n_rows) return;
  const uint64_t r1 = min(r0 + A.rows_per_thread, A.n_rows);
  // state before draw first_draw + r0: two engine calls per earlier row
  uint32_t x = lcg_mul(A.state0, lcg_pow(kLcgA, 2ull * ((A.first_draw + r0) % (kLcgM - 1u))));
  for (uint64_t r = r0; r < r1; r++)
  {
    const uint32_t x1 = lcg_mul(x, kLcgA);
    const uint32_t x2 = lcg_mul(x1, kLcgA);
    x = x2;
    // draw lies in [(x2-1)/R, x2/R): decide by integers unless x2 is within 2 of v*R
    bool take;
    const double x2d = static_cast<double>(x2);
    if (x2d < A.v_scaled - 2.0) take = true;
    else if (x2d - 1.0 > A.v_scaled + 2.0) take = false;
    else take = canonical_draw(x1, x2) < A.v;
    if (take)
    {
      const unsigned long long slot = atomicAdd(A.total, 1ull);
      if (slot < A.cap) A.out_rows[slot] = A.first_row + r;
    }
  }
}

struct PresenceArgs
{
  const uint8_t* rows;    // the batch
  uint64_t row_base;      // global index of its first row
  uint32_t row_bytes;
  uint32_t kmer_bytes;
  uint32_t n;             // samples
  uint32_t count_bytes;
  const unsigned long long* total;
  unsigned long long* done;  // slots [0, done) already have their presence vectors
  const uint64_t* slot_rows;
  uint8_t* geno;          // [cap][n] 0/1
  uint32_t* carriers;     // [cap] number of samples with count > 0
  uint64_t cap;
};

// EigGenoFile::push for the rows sampled in this batch: one warp per sampled row.
__global__ void __launch_bounds__(256) presence_kernel(const PresenceArgs A)
{
  const uint64_t lo = *A.done;
  const uint64_t hi = min(static_cast<uint64_t>(*A.total), A.cap);
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) >> 5;
  const uint64_t n_warps = (gridDim.x * static_cast<uint64_t>(blockDim.x)) >> 5;
  for (uint64_t s = lo + warp; s < hi; s += n_warps)
  {
    const uint8_t* row = A.rows + (A.slot_rows[s] - A.row_base) * A.row_bytes + A.kmer_bytes;
    uint32_t c = 0;
    for (uint32_t i = lane; i < A.n; i += 32)
    {
      uint32_t v;
      if (A.count_bytes == 4)
        v = __ldg(row + 4 * i) | __ldg(row + 4 * i + 1) | __ldg(row + 4 * i + 2) | __ldg(row + 4 * i + 3);
      else if (A.count_bytes == 2)
        v = __ldg(row + 2 * i) | __ldg(row + 2 * i + 1);
      else
        v = __ldg(row + i);
      const uint8_t bit = v != 0;
      A.geno[s * A.n + i] = bit;
      c += bit;
    }
    for (uint32_t o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if (lane == 0) A.carriers[s] = c;
  }
}

__global__ void presence_done_kernel(const unsigned long long* total, unsigned long long* done, uint64_t cap,
                                     uint32_t* overflow)
{
  if (*total > cap) *overflow = 1;
  *done = min(static_cast<uint64_t>(*total), cap);
}

// ---- logistic GLM by IRLS, one warp per k-mer ---------------------------------------------------

constexpr int kMaxP = 16;       // features per model (1 + n_pc + total + ratio <= 13 in the reference)
constexpr int kIrlsQ = (kMaxP * kMaxP + kMaxP + 31) / 32;   // [hessian | X^T S z] elements per lane, at most
constexpr int kIrlsWarps = 10;    // at most; fewer when the per-warp state of a large cohort fills shared memory

struct IrlsArgs
{
  uint32_t n;             // samples
  uint32_t p;             // features of this model
  uint32_t ldf;           // leading dimension of `features`
  uint32_t own_last;      // 1: column p-1 is count_i / total_i of the k-mer (alt model); 0: all from `features`
  const double* features; // [n][ldf] global features after standardize()
  const double* y;        // [n]
  const double* totals;   // [n] per-sample totals as double
  uint32_t max_iter;
  double llr_epsilon;     // s_epsilon
  double lgam_half;
  // alt model: the survivors
  uint32_t n_items;
  const uint32_t* n_items_dev;  // if not null: number of items is min(*n_items_dev, n_items)
  const uint8_t* counts;  // [item][n] raw counts of count_bytes each
  uint32_t count_bytes;
  double null_likelihood;
  // outputs (each may be null)
  double* p_value;        // [item]
  double* model;          // [item][p]
  double* llr;            // [item]
  double* likelihood;     // [item]
  int32_t* iters;         // [item]
  int32_t* flags;         // [item] bit0 singular, bit1 nan
  uint32_t feat_in_smem;  // features staged in shared memory
};

// linear_model.cpp:191-195
__device__ __forceinline__ double sigmoid_ref(double x)
{
  return div(1.0, add(1.0, pow(2.718281828459045235360, -x)));
}

struct IrlsWarp
{
  const double* F;   // features (shared or global), leading dimension ldf
  const double* xl;  // own last column (shared) or null
  uint32_t ldf, p, n;
  __device__ __forceinline__ double x(uint32_t i, uint32_t j) const
  {
    return (xl != nullptr && j == p - 1) ? xl[i] : F[i * ldf + j];
  }
};

// Shared memory layout (doubles): [features n*ldf (optional)] [y n] then per warp:
//   xl[n] mu[n] eta[n] d2[n] H[p*p] LU[p*p] INV[p*p] v[p] w[p] weight[p]
__host__ __device__ inline size_t irls_warp_doubles(uint32_t n, uint32_t p) { return 4ull * n + 3ull * p * p + 3ull * p; }

// P > 0: p == P features, staged in shared memory; every lane owns a 2 x 4 block of [hessian | X^T S z] and the row
// walk is straight-line code on shared-memory byte offsets.  P == 0: generic form
// (any p; features in global memory when the cohort is too large for shared memory).
template <int P>
__global__ void __launch_bounds__(kIrlsWarps * 32) glm_irls_kernel(const IrlsArgs A)
{
  // P > 0: the feature count is a compile-time constant (the host dispatches on it when the features fit shared memory):
  // the lanes' blocks of the hessian and every loop over features have fixed shapes
  extern __shared__ double sm[];
  const uint32_t n = A.n, p = P > 0 ? static_cast<uint32_t>(P) : A.p;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cursor = sm;
  const double* F = A.features;
  if (A.feat_in_smem)
  {
    for (uint32_t i = threadIdx.x; i < n * A.ldf; i += blockDim.x) cursor[i] = A.features[i];
    F = cursor;
    cursor += static_cast<size_t>(n) * A.ldf;
  }
  double* ys = cursor;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) ys[i] = A.y[i];
  cursor += n;
  double* mine = cursor + warp * irls_warp_doubles(n, p);
  double* xl = mine;
  double* mu = xl + n;
  double* eta = mu + n;
  double* d2 = eta + n;      // (y - mu)^2 of the current iteration
  double* H = d2 + n;        // hessian, later the forward-substitution results y[row][invc]
  double* LU = H + p * p;    // L strictly below the diagonal (unit diagonal implied), U on and above
  double* INV = LU + p * p;
  double* v = INV + p * p;   // X^T S z
  double* w = v + p;
  double* weight = w + p;
  __syncthreads();

  const uint32_t n_items = A.n_items_dev ? min(*A.n_items_dev, A.n_items) : A.n_items;
  IrlsWarp X{F, A.own_last ? xl : nullptr, A.ldf, p, n};
  const uint32_t pp = p * p;

  // elements of [hessian | X^T S z] owned by this lane: e = lane + 32 q; where column p-1 is the k-mer's own
  // (alt model) it is read from xl with stride 1, every other column from the feature matrix with stride ldf
  const uint32_t nq = (pp + p + 31) / 32;
  const double* el_a[kIrlsQ];
  const double* el_b[kIrlsQ];
  uint32_t el_sa[kIrlsQ], el_sb[kIrlsQ];
  bool el_v[kIrlsQ];
#pragma unroll
  for (int q = 0; q < kIrlsQ; q++)
  {
    const uint32_t e = lane + 32 * q;
    const bool is_v = e >= pp;
    uint32_t a = is_v ? e - pp : e / p, b = is_v ? 0 : e - (e / p) * p;
    if (a >= p) a = 0;   // unused slot: any valid address
    const bool own_a = A.own_last && a == p - 1, own_b = A.own_last && b == p - 1;
    el_a[q] = own_a ? xl : F + a;
    el_sa[q] = own_a ? 1 : A.ldf;
    el_b[q] = own_b ? xl : F + b;
    el_sb[q] = own_b ? 1 : A.ldf;
    el_v[q] = is_v;
  }

  // P > 0: a lane owns a 2 x 4 block of the p x (p + 1) matrix [hessian | X^T S z] — rows (2r, 2r+1), columns 4c .. 4c+3,
  // lane = r * ceil((p + 1) / 4) + c — so one row of the walk costs it 2 + 4 shared-memory loads for 8 slots, and the
  // four products S_k x_b are shared by its two rows.  Column p is X^T S z: it multiplies by z[k] (kept in eta[]
  // during the walk); S_k * z_k is the very product the generic form reads back as gz, so every slot is
  // acc += x_a[k] * (S_k * x_b[k]) and the walk has no per-slot branch.  Each slot still sums over k in row order with
  // the reference's operands, whatever lane holds it.  Offsets are shared-memory BYTE offsets (relative to sm) of row 0,
  // strides bytes per row; slots outside the matrix read a valid address and are dropped.
  constexpr int kCQ = P > 0 ? (P + 1 + 3) / 4 : 1;
  static_assert(P == 0 || ((P + 1) / 2) * kCQ <= 32, "the 2 x 4 blocks of [H | v] must fit one warp");
  uint32_t of_a[2], st_a[2], of_b[4], st_b[4];
  const uint32_t blk_r = lane / kCQ, blk_c = lane - blk_r * kCQ;
  if constexpr (P > 0)
  {
#pragma unroll
    for (int i = 0; i < 2; i++)
    {
      uint32_t a = 2 * blk_r + i;
      if (a >= p) a = 0;
      const bool own = A.own_last && a == p - 1;
      of_a[i] = static_cast<uint32_t>((own ? xl : F + a) - sm) * 8u;
      st_a[i] = (own ? 1u : A.ldf) * 8u;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
      uint32_t c = 4 * blk_c + j;
      if (c > p) c = 0;
      const bool is_v = c == p, own = A.own_last && c == p - 1;
      of_b[j] = static_cast<uint32_t>((is_v ? eta : (own ? xl : F + c)) - sm) * 8u;
      st_b[j] = (is_v || own) ? 8u : A.ldf * 8u;
    }
  }
  const char* const smb = reinterpret_cast<const char*>(sm);

  const uint32_t warps_per_block = blockDim.x >> 5;
  for (uint32_t item = blockIdx.x * warps_per_block + warp; item < n_items; item += gridDim.x * warps_per_block)
  {
    // ---- local_features[i][p-1] = counts[i] / totals[i]  (popstrat.hpp:254-257) ----
    if (A.own_last)
    {
      const uint8_t* c = A.counts + static_cast<size_t>(item) * n * A.count_bytes;
      for (uint32_t i = lane; i < n; i += 32)
      {
        uint32_t cv;
        if (A.count_bytes == 4) cv = reinterpret_cast<const uint32_t*>(c)[i];
        else if (A.count_bytes == 2) cv = reinterpret_cast<const uint16_t*>(c)[i];
        else cv = c[i];
        xl[i] = div(static_cast<double>(cv), A.totals[i]);
      }
    }
    // ---- glm_irls (linear_model.cpp:297-410) ----
    for (uint32_t i = lane; i < n; i += 32)
    {
      const double m0 = div(add(ys[i], 0.5), 2.0);
      mu[i] = m0;
      eta[i] = log(div(m0, sub(1.0, m0)));
    }
    for (uint32_t j = lane; j < p; j += 32) { weight[j] = 1.0; w[j] = 1.0; }
    __syncwarp();

    double prev_error = 1e18;
    int iter = 0, ise = 0, ine = 0;
    while (true)
    {
      // Per-row terms, rows spread over the lanes (each is a function of its own row only):
      //   d2 = (y - mu)^2, g = S[k] = mu (1 - mu), gz = S[k] * z[k] with z = eta + (y - mu) / (g + 1e-305).
      // mu[] and eta[] are dead until the new weights rebuild them below, so they take g and gz.
      for (uint32_t i = lane; i < n; i += 32)
      {
        const double m = mu[i];
        const double d = sub(ys[i], m);
        const double g = mul(m, sub(1.0, m));
        d2[i] = mul(d, d);
        const double z = add(eta[i], div(d, add(g, 1e-305)));
        mu[i] = g;
        eta[i] = P > 0 ? z : mul(g, z);
      }
      __syncwarp();

      // One walk over the rows, in row order, carries every sum of the iteration in the reference's order:
      //   error         = sum_k d2[k]                               (every lane, same value)
      //   hessian[a][b] = sum_{good k} X[k][a] * (S[k] * X[k][b])   (element e = a*p + b owned by lane e % 32)
      //   (X^T S z)[j]  = sum_{good k} X[k][j] * (S[k] * z[k])      (element p*p + j)
      double error = 0.0;
      uint32_t n_good = 0;
      if constexpr (P > 0)
      {
        double acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
#pragma unroll 2
        for (uint32_t k = 0; k < n; k++)
        {
          error = add(error, d2[k]);
          // a row with S_k <= 1e-305 is skipped by the reference; here it adds x_a * (0 * x_b) = +-0 to every slot, which
          // leaves a sum unchanged (the features are finite) — no branch in the walk
          const double gk = mu[k];
          const bool good = gk > 1e-305;
          n_good += good;
          const double g = good ? gk : 0.0;
          const double xa0 = *reinterpret_cast<const double*>(smb + (k * st_a[0] + of_a[0]));
          const double xa1 = *reinterpret_cast<const double*>(smb + (k * st_a[1] + of_a[1]));
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const double gx = mul(g, *reinterpret_cast<const double*>(smb + (k * st_b[j] + of_b[j])));
            acc[0][j] = add(acc[0][j], mul(xa0, gx));
            acc[1][j] = add(acc[1][j], mul(xa1, gx));
          }
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const uint32_t a = 2 * blk_r + i, c = 4 * blk_c + j;
            if (a < p && c < p) H[a * p + c] = acc[i][j];
            else if (a < p && c == p) v[a] = acc[i][j];
          }
      }
      else
      {
        double acc[kIrlsQ];
#pragma unroll
        for (int q = 0; q < kIrlsQ; q++) acc[q] = 0.0;
        for (uint32_t k = 0; k < n; k++)
        {
          error = add(error, d2[k]);
          const double g = mu[k];
          if (!(g > 1e-305)) continue;   // warp-uniform
          n_good++;
          const double gz = eta[k];
#pragma unroll
          for (int q = 0; q < kIrlsQ; q++)
          {
            if (q < nq)                  // warp-uniform
            {
              const double xa = el_a[q][k * el_sa[q]];
              const double xb = el_b[q][k * el_sb[q]];
              const double other = el_v[q] ? gz : mul(g, xb);
              acc[q] = add(acc[q], mul(xa, other));
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kIrlsQ; q++)
        {
          const uint32_t e = lane + 32 * q;
          if (e < pp) H[e] = acc[q];
          else if (e < pp + p) v[e - pp] = acc[q];
        }
      }
      if (n_good == 0) break;
      error = div(error, static_cast<double>(n));
      if (fabs(sub(error, prev_error)) < 1e-6) break;
      prev_error = error;
      __syncwarp();

      // ---- inverse(hessian): Doolittle LU without pivoting (linear_model.cpp:94-137) ----
      for (uint32_t i = 0; i < p; i++)
      {
        if (lane >= i && lane < p)
        {
          double s = 0.0;
          for (uint32_t j = 0; j < i; j++) s = add(s, mul(LU[i * p + j], LU[j * p + lane]));
          LU[i * p + lane] = sub(H[i * p + lane], s);  // upper[i][k]
        }
        __syncwarp();
        if (lane > i && lane < p)
        {
          double s = 0.0;
          for (uint32_t j = 0; j < i; j++) s = add(s, mul(LU[lane * p + j], LU[j * p + i]));
          LU[lane * p + i] = div(sub(H[lane * p + i], s), LU[i * p + i]);  // lower[k][i]
        }
        __syncwarp();
      }
      // forward and back substitution, one lane per column of the inverse (linear_model.cpp:139-189)
      if (lane < p)
      {
        const uint32_t c = lane;
        double* Y = H;  // y[row] of column c at Y[row * p + c]
        for (uint32_t row = 0; row < p; row++) Y[row * p + c] = 0.0;
        Y[c] = (c == 0) ? 1.0 : 0.0;
        for (uint32_t row = 1; row < p; row++)
        {
          double s = 0.0;
          for (uint32_t col = 0; col < p; col++)
          {
            const double l = col < row ? LU[row * p + col] : (col == row ? 1.0 : 0.0);
            s = add(s, mul(l, Y[col * p + c]));
          }
          Y[row * p + c] = sub(row == c ? 1.0 : 0.0, s);
        }
        for (uint32_t row = 0; row < p; row++) INV[row * p + c] = 0.0;
        INV[(p - 1) * p + c] = div(Y[(p - 1) * p + c], LU[(p - 1) * p + (p - 1)]);
        for (int row = static_cast<int>(p) - 2; row > -1; row--)
        {
          double s = 0.0;
          for (uint32_t col = row + 1; col < p; col++) s = add(s, mul(LU[row * p + col], INV[col * p + c]));
          INV[row * p + c] = div(sub(Y[row * p + c], s), LU[row * p + row]);
        }
      }
      // det: the reference multiplies the diagonal of U (last to first) once per inverse column
      double det = 1.0;
      for (uint32_t c = 0; c < p; c++)
        for (int row = static_cast<int>(p) - 1; row > -1; row--) det = mul(det, LU[row * p + row]);
      __syncwarp();
      const bool sing = det == 0.0;
      const bool isn = !sing && isnan(det);
      if (sing || isn)
      {
        ise = sing;
        ine = isn;
        break;
      }
      // w = hinv * (X^T S z)
      if (lane < p)
      {
        double r = 0.0;
        for (uint32_t k = 0; k < p; k++) r = add(r, mul(INV[lane * p + k], v[k]));
        w[lane] = r;
      }
      __syncwarp();
      iter += 1;
      if (iter >= static_cast<int>(A.max_iter)) break;
      if (lane < p) weight[lane] = w[lane];
      for (uint32_t i = lane; i < n; i += 32)
      {
        double e = 0.0;
        for (uint32_t j = 0; j < p; j++) e = add(e, mul(X.x(i, j), w[j]));
        eta[i] = e;
        mu[i] = sigmoid_ref(e);
      }
      __syncwarp();
    }
    __syncwarp();

    // ---- likelihood of the fitted model (popstrat.hpp:268-289): product in row order ----
    for (uint32_t i = lane; i < n; i += 32)
    {
      double s = 0.0;
      for (uint32_t j = 0; j < p; j++) s = add(s, mul(weight[j], X.x(i, j)));
      const double pr = sigmoid_ref(s);
      mu[i] = ys[i] == 1.0 ? pr : sub(1.0, pr);
    }
    __syncwarp();
    double lik = 1.0;
    for (uint32_t i = 0; i < n; i++) lik = mul(lik, mu[i]);

    if (lane == 0)
    {
      double alt = lik, null = A.null_likelihood;
      if (null == 0.0 && alt == 0.0)
      {
        null = 0.001;
        alt = 1.0;
      }
      const double ratio = div(null, alt);
      double llr = mul(-2.0, log(ratio));
      if (fabs(llr) < A.llr_epsilon || llr < 0.0 || isnan(lik)) llr = 0.0;
      if (A.p_value) A.p_value[item] = igamc_half(div(llr, 2.0), A.lgam_half);
      if (A.llr) A.llr[item] = llr;
      if (A.likelihood) A.likelihood[item] = lik;
      if (A.iters) A.iters[item] = iter;
      if (A.flags) A.flags[item] = ise | (ine << 1);
    }
    if (A.model && lane < p) A.model[static_cast<size_t>(item) * p + lane] = weight[lane];
    __syncwarp();
  }
}

}  // namespace pop
}  // namespace kmd
