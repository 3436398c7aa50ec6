// Exact fp64 arithmetic of the `kmdiff diff` model, usable from host and device.
//
// Every operation below is a single correctly-rounded fp64 op issued in the order the reference
// source states it (no FMA contraction: explicit __dadd_rn/__dmul_rn/__ddiv_rn on the device, and
// the host build targets baseline x86-64 which has no FMA).  log()/exp() come from CUDA's
// libdevice on the device (<= 1 ulp) and from glibc on the host.
//
// Reference being restated (file:line relative to the reference root):
//   PoissonLikelihood::poisson_prob / process   include/kmdiff/model.hpp:133-176
//   LogFactorialTable::operator[]               include/kmdiff/log_factorial_table.hpp:14-18
//   incompletegammac / incompletegamma          thirdparty/alglib/src/specialfunctions.cpp:4578-4739
//   chisquarecdistribution                      thirdparty/alglib/src/specialfunctions.cpp:9559-9567
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define KMD_HD __host__ __device__ __forceinline__
#else
#define KMD_HD inline
#endif

namespace kmd {

#if defined(__CUDA_ARCH__)
KMD_HD double add(double a, double b) { return __dadd_rn(a, b); }
KMD_HD double sub(double a, double b) { return __dsub_rn(a, b); }
KMD_HD double mul(double a, double b) { return __dmul_rn(a, b); }
KMD_HD double div(double a, double b) { return __ddiv_rn(a, b); }
#else
KMD_HD double add(double a, double b) { volatile double r = a + b; return r; }
KMD_HD double sub(double a, double b) { volatile double r = a - b; return r; }
KMD_HD double mul(double a, double b) { volatile double r = a * b; return r; }
KMD_HD double div(double a, double b) { volatile double r = a / b; return r; }
#endif

// lngamma(0.5) as ALGLIB's lngamma evaluates it (x < 13 branch: z = 1/0.5/1.5, rational tail);
// computed once on the host by lngamma_half() below and handed to the kernels as a constant.
constexpr double kIgamEps = 0.000000000000001;
constexpr double kMinLog = -709.78271289338399;
constexpr double kBig = 4503599627370496.0;
constexpr double kBigInv = 2.22044604925031308085 * 0.0000000000000001;

// specialfunctions.cpp:3744-3860, only the 0 < x < 13 branch that a = v/2 = 0.5 ever reaches.
inline double lngamma_small(double x)
{
  double z = 1, p = 0, u = x;
  while (u >= 3) { p = p - 1; u = x + p; z = z * u; }
  while (u < 2) { z = z / u; p = p + 1; u = x + p; }
  if (z < 0) z = -z;
  if (u == 2) return std::log(z);
  p = p - 2;
  x = x + p;
  double b = -1378.25152569120859100;
  b = -38801.6315134637840924 + x * b;
  b = -331612.992738871184744 + x * b;
  b = -1162370.97492762307383 + x * b;
  b = -1721737.00820839662146 + x * b;
  b = -853555.664245765465627 + x * b;
  double c = 1;
  c = -351.815701436523470549 + x * c;
  c = -17064.2106651881159223 + x * c;
  c = -220528.590553854454839 + x * c;
  c = -1139334.44367982507207 + x * c;
  c = -2532523.07177582951285 + x * c;
  c = -2018891.41433532773231 + x * c;
  p = x * b / c;
  return std::log(z) + p;
}

// specialfunctions.cpp:4598-4617 with a = 0.5: power series of the lower incomplete gamma
// (the part of incompletegamma after its argument checks).
KMD_HD double igam_series_half(double x, double lgam_a)
{
  const double a = 0.5;
  double ax = sub(sub(mul(a, log(x)), x), lgam_a);
  if (ax < kMinLog) return 0;
  ax = exp(ax);
  double r = a, c = 1, ans = 1;
  do
  {
    r = add(r, 1);
    c = div(mul(c, x), r);
    ans = add(ans, c);
  } while (div(c, ans) > kIgamEps);
  return div(mul(ans, ax), a);
}

// specialfunctions.cpp:4689-4737 with a = 0.5: continued fraction of the upper incomplete gamma
// (the part of incompletegammac after its argument checks).
KMD_HD double igamc_cfrac_half(double x, double lgam_a)
{
  const double a = 0.5;
  double ax = sub(sub(mul(a, log(x)), x), lgam_a);
  if (ax < kMinLog) return 0;
  ax = exp(ax);
  double y = sub(1, a);
  double z = add(add(x, y), 1);
  double c = 0;
  double pkm2 = 1;
  double qkm2 = x;
  double pkm1 = add(x, 1);
  double qkm1 = mul(z, x);
  double ans = div(pkm1, qkm1);
  double t;
  do
  {
    c = add(c, 1);
    y = add(y, 1);
    z = add(z, 2);
    double yc = mul(y, c);
    double pk = sub(mul(pkm1, z), mul(pkm2, yc));
    double qk = sub(mul(qkm1, z), mul(qkm2, yc));
    if (qk != 0)
    {
      double r = div(pk, qk);
      t = fabs(div(sub(ans, r), r));
      ans = r;
    }
    else
      t = 1;
    pkm2 = pkm1;
    pkm1 = pk;
    qkm2 = qkm1;
    qkm1 = qk;
    if (fabs(pk) > kBig)
    {
      pkm2 = mul(pkm2, kBigInv);
      pkm1 = mul(pkm1, kBigInv);
      qkm2 = mul(qkm2, kBigInv);
      qkm1 = mul(qkm1, kBigInv);
    }
  } while (t > kIgamEps);
  return mul(ans, ax);
}

// incompletegammac(0.5, x), specialfunctions.cpp:4655-4739.  The reference's two functions call
// each other once at most (igamc -> 1 - igam for x < 1; igam -> 1 - igamc for x > 1), so the
// dispatch is flattened here.
KMD_HD double igamc_half(double x, double lgam_a)
{
  if (x <= 0) return 1;
  if (x < 1 || x < 0.5) return sub(1, igam_series_half(x, lgam_a));
  return igamc_cfrac_half(x, lgam_a);
}

// Constants of one PoissonLikelihood instance (model.hpp:105-118, 182-191).
struct ModelConst
{
  double sum_controls;  // (double)Sc
  double sum_cases;     // (double)Sk
  double sum_all;       // (double)(Sc + Sk), integer addition first (model.hpp:148)
  double lgam_half;     // lngamma(0.5) by ALGLIB's formula
  const double* lf;     // log-factorial table (device or host pointer, matching the caller)
  uint64_t lf_size;
};

struct RowStat
{
  double p;
  int32_t sign;
  double mean_control;
  double mean_case;
};

// log_factorial_table.hpp:14-18.  Beyond the table the reference re-sums log(j) for j = k..2, an
// O(k) loop per call; here ln(k!) comes from lgamma(k+1).  The two differ by ~1e-16 relative, and
// the term cancels between the alternative and the null hypothesis, so the p-value moves by less
// than the fp64 noise the reference itself carries at such abundances (parity notes).
template <class TableLoad>
KMD_HD double log_factorial(const ModelConst& m, uint64_t k, TableLoad load)
{
  if (k < m.lf_size) return load(k);
  return lgamma(static_cast<double>(k) + 1.0);
}

// model.hpp:133-138.  The reference narrows the (double) sum to `int k`; sums >= 2^31 are
// undefined behaviour there.  Here k keeps its exact integer value.
template <class TableLoad>
KMD_HD double poisson_prob(const ModelConst& m, uint64_t k, double lambda, TableLoad load)
{
  if (lambda <= 0) return 0;
  double kl = mul(static_cast<double>(k), log(lambda));
  return add(-lambda, sub(kl, log_factorial(m, k, load)));
}

// model.hpp:142-176 given the two row sums (exact integers, < 2^53, so the reference's
// double-accumulated sums are the same numbers).
template <class TableLoad>
KMD_HD RowStat lrt_exact(const ModelConst& m, uint64_t sum_ctrl_i, uint64_t sum_case_i, TableLoad load)
{
  const double kc = static_cast<double>(sum_ctrl_i);
  const double kk = static_cast<double>(sum_case_i);
  const double mean = div(add(kc, kk), m.sum_all);

  double alt = 0, null = 0;
  alt = add(alt, poisson_prob(m, sum_ctrl_i, kc, load));
  alt = add(alt, poisson_prob(m, sum_case_i, kk, load));
  null = add(null, poisson_prob(m, sum_ctrl_i, mul(mean, m.sum_controls), load));
  null = add(null, poisson_prob(m, sum_case_i, mul(mean, m.sum_cases), load));

  double lr = sub(alt, null);
  if (lr < 0) lr = 0;
  // chisquarecdistribution(1, 2*lr) = incompletegammac(1/2.0, (2*lr)/2.0); (2*lr)/2 == lr exactly
  // unless 2*lr overflows, which needs lr > 8.9e307.
  RowStat out;
  out.p = igamc_half(div(mul(2, lr), 2.0), m.lgam_half);

  const double scaled = div(mul(kc, m.sum_cases), m.sum_controls);
  out.sign = scaled < kk ? 1 : (scaled > kk ? 0 : 2);
  out.mean_control = scaled;
  out.mean_case = kk;
  return out;
}

}  // namespace kmd
