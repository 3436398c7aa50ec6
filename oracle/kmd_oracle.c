/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See kmd_oracle.h for scope, parity status and rules.
 *
 * Plain C restatement of the reference's `kmdiff diff` arithmetic.  Compiled with
 * -ffp-contract=off so that every fp64 operation rounds exactly once, in the order the reference
 * source states it (the reference build in oracle/_ref is likewise compiled without FMA
 * contraction being available: no -march=native, x86-64 baseline has no FMA).
 */
#include "kmd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Log-factorial table
 * ------------------------------------------------------------------------------------------- */

/* src/log_factorial_table.cpp:13-22.  The order matters for bit-exactness: the reference starts
 * from log(k) and walks DOWN to log(2) (tests/factorial_test.cpp pins t[10], t[50], t[51], t[100]). */
double kmo_log_factorial(uint64_t k)
{
  double acc = 0;
  for (; k > 1; k--) acc += log((double)k);
  return acc;
}

/* src/log_factorial_table.cpp:5-11 */
void kmo_lf_table_fill(double* table, uint64_t size)
{
  for (uint64_t i = 0; i < size; i++) table[i] = kmo_log_factorial(i);
}

/* include/kmdiff/log_factorial_table.hpp:14-18 */
double kmo_lf_lookup(const double* table, uint64_t size, uint64_t i)
{
  return i < size ? table[i] : kmo_log_factorial(i);
}

/* ---------------------------------------------------------------------------------------------
 * Cephes gamma family as vendored in ALGLIB 3.17.0
 * ------------------------------------------------------------------------------------------- */

/* alglib/specialfunctions.cpp:3744-3860 */
double kmo_lngamma(double x, double* sgngam)
{
  const double log_pi = 1.14472988584940017414;
  const double log_sqrt_2pi = 0.91893853320467274178;
  double dummy;
  *sgngam = 1;

  if (x < -34.0)
  {
    /* reflection */
    double q = -x;
    double w = kmo_lngamma(q, &dummy);
    double p = floor(q);
    long i = lround(p);
    *sgngam = (i % 2 == 0) ? -1 : 1;
    double z = q - p;
    if (z > 0.5)
    {
      p = p + 1;
      z = p - q;
    }
    z = q * sin(M_PI * z);
    return log_pi - log(z) - w;
  }

  if (x < 13)
  {
    /* shift the argument into [2,3) and use the rational approximation */
    double z = 1, p = 0, u = x;
    while (u >= 3)
    {
      p = p - 1;
      u = x + p;
      z = z * u;
    }
    while (u < 2)
    {
      z = z / u;
      p = p + 1;
      u = x + p;
    }
    if (z < 0)
    {
      *sgngam = -1;
      z = -z;
    }
    else
      *sgngam = 1;
    if (u == 2) return log(z);
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
    return log(z) + p;
  }

  /* Stirling */
  double q = (x - 0.5) * log(x) - x + log_sqrt_2pi;
  if (x > 100000000) return q;
  double p = 1 / (x * x);
  if (x >= 1000.0)
  {
    q = q + ((7.9365079365079365079365 * 0.0001 * p - 2.7777777777777777777778 * 0.001) * p +
             0.0833333333333333333333) / x;
  }
  else
  {
    double a = 8.11614167470508450300 * 0.0001;
    a = -5.95061904284301438324 * 0.0001 + p * a;
    a = 7.93650340457716943945 * 0.0001 + p * a;
    a = -2.77777777730099687205 * 0.001 + p * a;
    a = 8.33333333333331927722 * 0.01 + p * a;
    q = q + a / x;
  }
  return q;
}

static const double KMO_IGAM_EPS = 0.000000000000001;
static const double KMO_MIN_LOG = -709.78271289338399;

/* alglib/specialfunctions.cpp:4578-4619 — power series for the lower incomplete gamma */
double kmo_igam(double a, double x)
{
  double sg;
  if (x <= 0 || a <= 0) return 0;
  if (x > 1 && x > a) return 1 - kmo_igamc(a, x);
  double ax = a * log(x) - x - kmo_lngamma(a, &sg);
  if (ax < KMO_MIN_LOG) return 0;
  ax = exp(ax);
  double r = a, c = 1, ans = 1;
  do
  {
    r = r + 1;
    c = c * x / r;
    ans = ans + c;
  } while (c / ans > KMO_IGAM_EPS);
  return ans * ax / a;
}

/* alglib/specialfunctions.cpp:4655-4739 — continued fraction for the upper incomplete gamma */
double kmo_igamc(double a, double x)
{
  const double big = 4503599627370496.0;
  const double biginv = 2.22044604925031308085 * 0.0000000000000001;
  double sg;
  if (x <= 0 || a <= 0) return 1;
  if (x < 1 || x < a) return 1 - kmo_igam(a, x);
  double ax = a * log(x) - x - kmo_lngamma(a, &sg);
  if (ax < KMO_MIN_LOG) return 0;
  ax = exp(ax);
  double y = 1 - a;
  double z = x + y + 1;
  double c = 0;
  double pkm2 = 1;
  double qkm2 = x;
  double pkm1 = x + 1;
  double qkm1 = z * x;
  double ans = pkm1 / qkm1;
  double t;
  do
  {
    c = c + 1;
    y = y + 1;
    z = z + 2;
    double yc = y * c;
    double pk = pkm1 * z - pkm2 * yc;
    double qk = qkm1 * z - qkm2 * yc;
    if (qk != 0)
    {
      double r = pk / qk;
      t = fabs((ans - r) / r);
      ans = r;
    }
    else
      t = 1;
    pkm2 = pkm1;
    pkm1 = pk;
    qkm2 = qkm1;
    qkm1 = qk;
    if (fabs(pk) > big)
    {
      pkm2 = pkm2 * biginv;
      pkm1 = pkm1 * biginv;
      qkm2 = qkm2 * biginv;
      qkm1 = qkm1 * biginv;
    }
  } while (t > KMO_IGAM_EPS);
  return ans * ax;
}

/* alglib/specialfunctions.cpp:9559-9567 (the domain assert is the caller's problem here) */
double kmo_chisqc(double v, double x) { return kmo_igamc(v / 2.0, x / 2.0); }

/* ---------------------------------------------------------------------------------------------
 * Poisson likelihood-ratio model
 * ------------------------------------------------------------------------------------------- */

/* include/kmdiff/model.hpp:133-138.  `k` arrives as a double (a sum of counts) and is narrowed
 * to int by the call, like the reference's `poisson_prob(int k, double lambda)`.  Sums >= 2^31
 * are undefined behaviour in the reference; the oracle refuses to guess and saturates. */
static double kmo_poisson_prob(const kmo_model* m, double k_as_double, double lambda)
{
  int k = k_as_double >= 2147483647.0 ? 2147483647 : (int)k_as_double;
  if (lambda <= 0) return 0;
  if (k < 0) k = 0;
  return (-lambda + (k * log(lambda) - kmo_lf_lookup(m->lf_table, m->lf_size, (uint64_t)k)));
}

/* include/kmdiff/model.hpp:70-81 — the sum runs in double, in sample order */
static double kmo_sum(const void* counts, int count_bytes, uint32_t start, uint32_t n)
{
  double s = 0;
  for (uint32_t i = start; i < start + n; i++)
  {
    switch (count_bytes)
    {
      case 1: s += ((const uint8_t*)counts)[i]; break;
      case 2: s += ((const uint16_t*)counts)[i]; break;
      default: s += ((const uint32_t*)counts)[i]; break;
    }
  }
  return s;
}

/* include/kmdiff/model.hpp:142-176 */
void kmo_poisson_process(const kmo_model* m, const void* counts, int count_bytes, double* p,
                         int32_t* sign, double* mean_control, double* mean_case)
{
  double sum_ctrl = kmo_sum(counts, count_bytes, 0, m->nb_controls);
  double sum_case = kmo_sum(counts, count_bytes, m->nb_controls, m->nb_cases);

  /* m_sum_controls + m_sum_cases is a size_t addition, then one conversion (model.hpp:148) */
  double mean = (sum_ctrl + sum_case) / (double)(m->sum_controls + m->sum_cases);

  double alt = 0, null = 0;
  alt += kmo_poisson_prob(m, sum_ctrl, sum_ctrl);
  alt += kmo_poisson_prob(m, sum_case, sum_case);
  null += kmo_poisson_prob(m, sum_ctrl, mean * (double)m->sum_controls);
  null += kmo_poisson_prob(m, sum_case, mean * (double)m->sum_cases);

  double lr = alt - null;
  if (lr < 0) lr = 0;
  *p = kmo_chisqc(1, 2 * lr);

  double scaled_ctrl = sum_ctrl * (double)m->sum_cases / (double)m->sum_controls;
  if (scaled_ctrl < sum_case)
    *sign = KMO_CASE;
  else if (scaled_ctrl > sum_case)
    *sign = KMO_CONTROL;
  else
    *sign = KMO_NO;
  *mean_control = scaled_ctrl;
  *mean_case = sum_case;
}

void kmo_model_rows(const kmo_model* m, const uint8_t* rows, uint64_t n_rows, int kmer_words,
                    int count_bytes, double* p, int32_t* sign, double* mc, double* kk)
{
  const uint64_t kb = (uint64_t)kmer_words * 8;
  const uint64_t R = kb + (uint64_t)(m->nb_controls + m->nb_cases) * count_bytes;
  /* counts may be unaligned inside the AoS payload when R is odd-sized: copy out first */
  void* tmp = malloc((size_t)(R - kb) + 8);
  for (uint64_t r = 0; r < n_rows; r++)
  {
    memcpy(tmp, rows + r * R + kb, (size_t)(R - kb));
    kmo_poisson_process(m, tmp, count_bytes, p + r, sign + r, mc + r, kk + r);
  }
  free(tmp);
}

/* include/kmdiff/merge.hpp:194-203 + 68-103 */
uint64_t kmo_diff_rows(const kmo_model* m, const uint8_t* rows, uint64_t n_rows, int kmer_words,
                       int count_bytes, double thr, uint64_t row_base, kmo_survivor* out,
                       uint64_t cap, uint64_t* n_control, uint64_t* n_case)
{
  const uint64_t kb = (uint64_t)kmer_words * 8;
  const uint64_t R = kb + (uint64_t)(m->nb_controls + m->nb_cases) * count_bytes;
  void* tmp = malloc((size_t)(R - kb) + 8);
  uint64_t found = 0, nctl = 0, ncas = 0;
  for (uint64_t r = 0; r < n_rows; r++)
  {
    double p, mc, kk;
    int32_t sign;
    memcpy(tmp, rows + r * R + kb, (size_t)(R - kb));
    kmo_poisson_process(m, tmp, count_bytes, &p, &sign, &mc, &kk);
    if (p <= thr) /* merge.hpp:78 — inclusive */
    {
      if (found < cap)
      {
        kmo_survivor* s = out + found;
        memset(s, 0, sizeof(*s));
        memcpy(s->kmer, rows + r * R, kb > 16 ? 16 : (size_t)kb);
        s->p = p;
        s->sign = sign;
        s->mean_control = mc;
        s->mean_case = kk;
        s->row = row_base + r;
      }
      if (sign == KMO_CONTROL) /* merge.hpp:95-98: NO lands in the case bucket */
        nctl++;
      else
        ncas++;
      found++;
    }
  }
  free(tmp);
  if (n_control) *n_control = nctl;
  if (n_case) *n_case = ncas;
  return found;
}

/* ---------------------------------------------------------------------------------------------
 * Correctors and aggregation
 * ------------------------------------------------------------------------------------------- */

void kmo_corrector_init(kmo_corrector* c, int type, double threshold, uint64_t total)
{
  c->type = type;
  c->threshold = threshold;
  c->total = total;
  c->rank = 1;
}

int kmo_corrector_apply(kmo_corrector* c, double p)
{
  switch (c->type)
  {
    case KMO_BONFERRONI: /* src/corrector.cpp:9-12 */
      return p < (c->threshold / (double)c->total);
    case KMO_BENJAMINI: /* src/corrector.cpp:27-35 */
      if (p < (((double)c->rank / (double)c->total) * c->threshold))
      {
        c->rank++;
        return 1;
      }
      return 0;
    case KMO_SIDAK: /* src/corrector.cpp:50-53 */
      return p < (1 - pow(1 - c->threshold, 1.0 / (double)c->total));
    case KMO_HOLM: /* src/corrector.cpp:68-71 — post-decrement of an unsigned total */
    {
      uint64_t m = c->total--;
      return p < (c->threshold / (double)m);
    }
    default: /* src/corrector.cpp:86-89 */
      return p < c->threshold;
  }
}

typedef struct
{
  double p;
  uint64_t idx;
} kmo_pi;

static int kmo_pi_cmp(const void* a, const void* b)
{
  const kmo_pi* x = (const kmo_pi*)a;
  const kmo_pi* y = (const kmo_pi*)b;
  if (x->p < y->p) return -1;
  if (x->p > y->p) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* include/kmdiff/aggregator.hpp:137-163 (stateless) and 282-310 (sorted, break at first reject).
 * The reference's heap order among equal p-values is unspecified; ties cannot straddle the
 * break (thresholds are non-decreasing in rank for BH, and for Holm t/(m-i) grows with i), so the
 * kept SET does not depend on it.  We fix ties by input index. */
uint64_t kmo_aggregate(const double* p, uint64_t n, int type, double threshold, uint64_t total,
                       uint64_t* keep_idx)
{
  kmo_corrector c;
  kmo_corrector_init(&c, type, threshold, total);
  uint64_t kept = 0;
  if (type == KMO_BENJAMINI || type == KMO_HOLM)
  {
    kmo_pi* v = (kmo_pi*)malloc(sizeof(kmo_pi) * (size_t)(n ? n : 1));
    for (uint64_t i = 0; i < n; i++)
    {
      v[i].p = p[i];
      v[i].idx = i;
    }
    qsort(v, (size_t)n, sizeof(kmo_pi), kmo_pi_cmp);
    for (uint64_t i = 0; i < n; i++)
    {
      if (!kmo_corrector_apply(&c, v[i].p)) break;
      keep_idx[kept++] = v[i].idx;
    }
    free(v);
  }
  else
  {
    for (uint64_t i = 0; i < n; i++)
      if (kmo_corrector_apply(&c, p[i])) keep_idx[kept++] = i;
  }
  return kept;
}

/* kmtricks kmer.hpp:38 (bToN = {A, C, T, G}) and :797-810 */
void kmo_kmer_to_string(const uint64_t* words, int kmer_words, int k, char* out)
{
  static const char alphabet[4] = {'A', 'C', 'T', 'G'};
  for (int i = 0; i < k; i++)
  {
    int bit = 2 * (k - 1 - i); /* first base sits in the most significant used bits */
    int w = bit / 64;
    uint64_t v = w < kmer_words ? (words[w] >> (bit % 64)) & 3u : 0;
    out[i] = alphabet[v];
  }
  out[k] = 0;
}
