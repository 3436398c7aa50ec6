/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See kmd_popstrat_oracle.h for scope, citations and
 * parity status (sampler / IRLS / pop_strat_corrector / smartpca restatement all pinned).
 * Compile with -ffp-contract=off: the reference is built for baseline x86-64 (no FMA). */
#define _USE_MATH_DEFINES
#include "kmd_popstrat_oracle.h"
#include "kmd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================================================
 * Sampler — popstrat.hpp:107-146 over libstdc++'s minstd_rand0 + generate_canonical<double,53>
 * ================================================================================================ */

#define KMO_LCG_A 16807ull
#define KMO_LCG_M 2147483647ull

void kmo_sampler_init(kmo_sampler* s, uint64_t seed)
{
  /* default-constructed engine: state 1; Sampler seeds only when seed != 0 (popstrat.hpp:113-117);
   * linear_congruential_engine::seed: c == 0 and s % m == 0 -> 1 */
  if (seed == 0) { s->state = 1; return; }
  s->state = seed % KMO_LCG_M;
  if (s->state == 0) s->state = 1;
}

static uint64_t lcg_next(kmo_sampler* s)
{
  s->state = (s->state * KMO_LCG_A) % KMO_LCG_M;
  return s->state;
}

double kmo_sampler_draw(kmo_sampler* s)
{
  /* generate_canonical: range R = max-min+1 = 2147483646, log2 R = 30 -> k = 2 calls;
   * sum += (urng()-min) * tmp; tmp *= R; ret = sum / tmp (all in double) */
  const double R = 2147483646.0;
  double sum = 0.0, tmp = 1.0;
  sum += (double)(lcg_next(s) - 1) * tmp;
  tmp *= R;
  sum += (double)(lcg_next(s) - 1) * tmp;
  tmp *= R;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  /* uniform_real_distribution(0,1): ret * (b - a) + a */
  return ret * (1.0 - 0.0) + 0.0;
}

uint64_t kmo_sampler_rows(uint64_t seed, double kmer_pca, uint64_t first_row, uint64_t n, uint8_t* out)
{
  kmo_sampler s;
  kmo_sampler_init(&s, seed);
  for (uint64_t r = 0; r < first_row; r++) { lcg_next(&s); lcg_next(&s); }
  uint64_t taken = 0;
  for (uint64_t i = 0; i < n; i++)
  {
    out[i] = kmo_sampler_draw(&s) < kmer_pca;
    taken += out[i];
  }
  return taken;
}

/* ================================================================================================
 * linear_model.cpp
 * ================================================================================================ */

double kmo_sigmoid(double x)
{
  double e = M_E;
  return 1.0 / (1.0 + pow(e, -x));
}

void kmo_lu(const double* a, int n, double* lower, double* upper)
{
  memset(lower, 0, sizeof(double) * n * n);
  memset(upper, 0, sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
  {
    for (int k = i; k < n; k++)
    {
      double sum = 0.0;
      for (int j = 0; j < i; j++) sum += lower[i * n + j] * upper[j * n + k];
      upper[i * n + k] = a[i * n + k] - sum;
    }
    for (int k = i; k < n; k++)
    {
      if (i == k) lower[i * n + i] = 1;
      else
      {
        double sum = 0;
        for (int j = 0; j < i; j++) sum += lower[k * n + j] * upper[j * n + i];
        lower[k * n + i] = (a[k * n + i] - sum) / upper[i * n + i];
      }
    }
  }
}

void kmo_inverse(const double* a, int n, double* inv, int* singular, int* is_nan)
{
  double* lower = (double*)malloc(sizeof(double) * n * n);
  double* upper = (double*)malloc(sizeof(double) * n * n);
  double* b = (double*)malloc(sizeof(double) * n);
  double* y = (double*)malloc(sizeof(double) * n);
  double* x = (double*)malloc(sizeof(double) * n);
  kmo_lu(a, n, lower, upper);
  memset(inv, 0, sizeof(double) * n * n);
  double det = 1;
  for (int invc = 0; invc < n; invc++)
  {
    for (int j = 0; j < n; j++) b[j] = (invc == j) ? 1 : 0;
    for (int j = 0; j < n; j++) y[j] = 0;
    det *= lower[0];
    y[0] = b[0];
    for (int row = 1; row < n; row++)
    {
      double sum = 0;
      for (int col = 0; col < n; col++) sum += lower[row * n + col] * y[col];
      y[row] = b[row] - sum;
      det *= lower[row * n + row];
    }
    for (int j = 0; j < n; j++) x[j] = 0;
    x[n - 1] = y[n - 1] / upper[(n - 1) * n + n - 1];
    det *= upper[(n - 1) * n + n - 1];
    for (int row = n - 2; row > -1; row--)
    {
      double sum = 0;
      for (int col = row + 1; col < n; col++) sum += upper[row * n + col] * x[col];
      x[row] = (y[row] - sum) / upper[row * n + row];
      det *= upper[row * n + row];
    }
    for (int j = 0; j < n; j++) inv[j * n + invc] = x[j];
  }
  *singular = 0;
  *is_nan = 0;
  if (det == 0) *singular = 1;
  else if (isnan(det)) *is_nan = 1;
  free(lower); free(upper); free(b); free(y); free(x);
}

void kmo_glm_irls(const double* x, const double* y, int n, int p, int max_iters, double* weight,
                  int* singular, int* is_nan, double* ret_error_out, int* iters)
{
  const double epsilon = 1e-6;
  int iter = 0, ein = 0, ise = 0, ine = 0;
  double ret_error = 0.0;
  double* w = (double*)malloc(sizeof(double) * p);
  double* eta = (double*)malloc(sizeof(double) * n);
  double* mu = (double*)malloc(sizeof(double) * n);
  int* good = (int*)malloc(sizeof(int) * n);      /* indices of the rows kept in X */
  double* S = (double*)malloc(sizeof(double) * n);
  double* z = (double*)malloc(sizeof(double) * n);
  double* S_X = (double*)malloc(sizeof(double) * n * p);
  double* S_z = (double*)malloc(sizeof(double) * n);
  double* hessian = (double*)malloc(sizeof(double) * p * p);
  double* hinv = (double*)malloc(sizeof(double) * p * p);
  double* XTSz = (double*)malloc(sizeof(double) * p);
  for (int i = 0; i < p; i++) { weight[i] = 1; w[i] = 1; }
  for (int i = 0; i < n; i++)
  {
    mu[i] = (y[i] + 0.5) / 2;
    eta[i] = log(mu[i] / (1 - mu[i]));
  }
  double prev_error = 1e18;

  while (1)
  {
    double error = 0.0;
    int num_of_good = 0;
    for (int i = 0; i < n; i++)
    {
      double g_i = mu[i] * (1.0 - mu[i]);
      if (g_i > 1e-305)
      {
        good[num_of_good] = i;
        z[num_of_good] = eta[i] + (y[i] - mu[i]) / (g_i + 1e-305);
        S[num_of_good] = g_i;
        num_of_good++;
      }
      error += (y[i] - mu[i]) * (y[i] - mu[i]);
    }
    if (num_of_good == 0) break;
    error /= n;
    ret_error = error;
    if (fabs(error - prev_error) < epsilon) break;
    prev_error = error;

    for (int i = 0; i < num_of_good; i++)
      for (int j = 0; j < p; j++) S_X[i * p + j] = S[i] * x[good[i] * p + j];
    /* hessian = X^T (S X): res[a][b] = res[a][b] + X_T[a][k] * S_X[k][b], k ascending, from 0.0 */
    for (int a = 0; a < p; a++)
      for (int b = 0; b < p; b++)
      {
        double r = 0.0;
        for (int k = 0; k < num_of_good; k++) r = r + x[good[k] * p + a] * S_X[k * p + b];
        hessian[a * p + b] = r;
      }
    int sing, nan_;
    kmo_inverse(hessian, p, hinv, &sing, &nan_);
    if (sing || nan_)
    {
      ise = sing;
      ine = nan_;
      ret_error = prev_error;
      break;
    }
    for (int i = 0; i < num_of_good; i++) S_z[i] = S[i] * z[i];
    for (int a = 0; a < p; a++)
    {
      double r = 0.0;
      for (int k = 0; k < num_of_good; k++) r = r + x[good[k] * p + a] * S_z[k];
      XTSz[a] = r;
    }
    for (int a = 0; a < p; a++)
    {
      double r = 0.0;
      for (int k = 0; k < p; k++) r = r + hinv[a * p + k] * XTSz[k];
      w[a] = r;
    }
    iter += 1;
    ein = iter;
    if (iter >= max_iters) break;
    prev_error = error;
    ret_error = prev_error;
    for (int i = 0; i < p; i++) weight[i] = w[i];
    for (int i = 0; i < n; i++)
    {
      eta[i] = 0;
      for (int j = 0; j < p; j++) eta[i] += x[i * p + j] * w[j];
      mu[i] = kmo_sigmoid(eta[i]);
    }
  }
  if (singular) *singular = ise;
  if (is_nan) *is_nan = ine;
  if (ret_error_out) *ret_error_out = ret_error;
  if (iters) *iters = ein;
  free(w); free(eta); free(mu); free(good); free(S); free(z); free(S_X); free(S_z);
  free(hessian); free(hinv); free(XTSz);
}

/* ================================================================================================
 * pop_strat_corrector
 * ================================================================================================ */

/* popstrat.cpp:327-370, bugs included: column means divided by the column COUNT, deviations
 * indexed by column but applied by ROW, so only rows 1..ncols-1 are rescaled. */
static void standardize(kmo_popstrat* ps)
{
  const int n = ps->n, pc = ps->null_p, pa = ps->alt_p;
  double* means = (double*)calloc(pc, sizeof(double));
  double* stddev = (double*)calloc(n, sizeof(double));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < pc; j++) means[j] += ps->null_features[i * pc + j];
  for (int i = 1; i < pc; i++) means[i] /= pc;
  for (int i = 0; i < n; i++)
    for (int j = 1; j < pc; j++)
    {
      const double d = ps->null_features[i * pc + j] - means[j];
      stddev[j] += d * d; /* std::pow(d, 2) */
    }
  for (int i = 1; i < pc; i++)
  {
    stddev[i] /= n;
    stddev[i] = sqrt(stddev[i]);
  }
  for (int i = 0; i < n; i++)
    for (int j = 1; j < pc; j++)
      if (fabs(stddev[i]) > 1e-305)
      {
        ps->null_features[i * pc + j] = (ps->null_features[i * pc + j] - means[j]) / stddev[i];
        ps->alt_features[i * pa + j] = (ps->alt_features[i * pa + j] - means[j]) / stddev[i];
      }
  free(means);
  free(stddev);
}

int kmo_popstrat_init(kmo_popstrat* ps, int nb_controls, int nb_cases, const uint64_t* totals,
                      const double* Z, int npc)
{
  memset(ps, 0, sizeof(*ps));
  const int n = nb_controls + nb_cases;
  if (n < npc + 2) return 1; /* standardize() indexes stddev[ncols-1] in a vector of n entries */
  ps->n = n;
  ps->nb_controls = nb_controls;
  ps->nb_cases = nb_cases;
  ps->npc = npc;
  ps->null_p = 1 + npc + 0 + 1;
  ps->alt_p = 1 + ps->null_p;
  ps->max_iter = 100;
  ps->epsilon = 1e-30;
  const int pc = ps->null_p, pa = ps->alt_p;
  ps->null_features = (double*)calloc((size_t)n * pc, sizeof(double));
  ps->alt_features = (double*)calloc((size_t)n * pa, sizeof(double));
  ps->Y = (double*)malloc(sizeof(double) * n);
  ps->totals = (double*)malloc(sizeof(double) * n);
  ps->null_model = (double*)malloc(sizeof(double) * pc);
  for (int i = 0; i < n; i++)
  {
    ps->Y[i] = i < nb_controls ? 1.0 : 0.0; /* load_Y: "Case" ? 0.0 : 1.0 */
    ps->totals[i] = (double)totals[i];
    ps->null_features[i * pc + 0] = 1;
    ps->alt_features[i * pa + 0] = 1;
    for (int zc = 0; zc < npc; zc++)
    {
      ps->null_features[i * pc + zc + 1] = Z[i * 10 + zc];
      ps->alt_features[i * pa + zc + 1] = Z[i * 10 + zc];
    }
    /* m_unkg != 0 (load_ginfo never recognises a gender, popstrat.cpp:236-268) */
    ps->null_features[i * pc + 1 + npc] = ps->totals[i];
    ps->alt_features[i * pa + 1 + npc] = ps->totals[i];
  }
  standardize(ps); /* s_stand is always true (set_params only ever sets it) */
  kmo_glm_irls(ps->null_features, ps->Y, n, pc, ps->max_iter, ps->null_model, NULL, NULL, NULL, NULL);
  return 0;
}

void kmo_popstrat_free(kmo_popstrat* ps)
{
  free(ps->null_features); free(ps->alt_features); free(ps->Y); free(ps->totals); free(ps->null_model);
  memset(ps, 0, sizeof(*ps));
}

static double predict(const double* model, const double* data, int p)
{
  double s = 0.0;
  for (int i = 0; i < p; i++) s += model[i] * data[i];
  return kmo_sigmoid(s);
}

double kmo_popstrat_apply(const kmo_popstrat* ps, const double* counts, double* model_out, int* iters,
                          double* llr_out)
{
  const int n = ps->n, pc = ps->null_p, pa = ps->alt_p;
  double* local = (double*)malloc(sizeof(double) * n * pa);
  double* model = (double*)malloc(sizeof(double) * pa);
  memcpy(local, ps->alt_features, sizeof(double) * n * pa);
  for (int i = 0; i < n; i++) local[i * pa + pa - 1] = counts[i] / ps->totals[i];
  kmo_glm_irls(local, ps->Y, n, pa, ps->max_iter, model, NULL, NULL, NULL, iters);

  double alt_likelihood = 1.0;
  for (int f = 0; f < n; f++)
  {
    const double p = predict(model, local + f * pa, pa);
    if (ps->Y[f] == 1) alt_likelihood = alt_likelihood * p;
    else alt_likelihood *= 1.0 - p;
  }
  double null_likelihood = 1.0;
  for (int f = 0; f < n; f++)
  {
    const double p = predict(ps->null_model, ps->null_features + f * pc, pc);
    if (ps->Y[f] == 1) null_likelihood *= p;
    else null_likelihood *= 1.0 - p;
  }
  if (null_likelihood == 0.0 && alt_likelihood == 0.0)
  {
    null_likelihood = 0.001;
    alt_likelihood = 1.0;
  }
  const double likelihood_ratio = null_likelihood / alt_likelihood;
  double llr = -2.0 * log(likelihood_ratio);
  if (fabs(llr) < ps->epsilon || llr < 0.0 || isnan(alt_likelihood)) llr = 0.0;
  if (model_out) memcpy(model_out, model, sizeof(double) * pa);
  if (llr_out) *llr_out = llr;
  free(local);
  free(model);
  return kmo_chisqc(1, llr);
}

/* ================================================================================================
 * smartpca (HAWK's EIG6.0.1) — pinned against the real program, tests/test_smartpca_reference.py
 * ================================================================================================ */

/* getcolxz + fvadjust for one SNP.  Returns 0 if the SNP is ignored (no carrier). */
static int snp_column(const uint8_t* g, int n, int is_diploid, double* cc, double* pmean, double* pfancy)
{
  int c = 0;
  for (int i = 0; i < n; i++) c += g[i] > 0;
  /* fvadjust: binarise, mean over non-missing */
  double ynum = 0.0, ysum = 0.0;
  for (int i = 0; i < n; i++) { cc[i] = g[i] > 0 ? 1 : 0; ++ynum; ysum += cc[i]; }
  const double ymean = ysum / ynum;
  double yfancy = 1.0;
  ysum = 0;
  for (int i = 0; i < n; i++) ysum += (cc[i] - ymean) * (cc[i] - ymean);
  if (ysum > 0) yfancy = 1.0 / sqrt(ysum / (n - 1));
  for (int i = 0; i < n; i++) cc[i] -= ymean;
  /* fancynorm (usenorm: YES) */
  const double p = is_diploid ? 1 - sqrt(1 - ymean) : ymean;
  const double y = p * (1.0 - p);
  if (y > 0.0) yfancy = 1.0 / sqrt(y);
  for (int i = 0; i < n; i++) cc[i] = cc[i] * yfancy; /* vst */
  *pmean = ymean * yfancy;                            /* xmean[col] */
  *pfancy = yfancy;                                   /* xfancy[col] */
  /* t = MIN(n0, n1) = c; ignored when t < minallelecnt (1) or t == 0 */
  return c > 0;
}

uint64_t kmo_smartpca_xtx(const uint8_t* geno, uint64_t m, int n, int is_diploid, double* xtx)
{
  double* cc = (double*)malloc(sizeof(double) * n);
  memset(xtx, 0, sizeof(double) * n * n);
  uint64_t used = 0;
  for (uint64_t s = 0; s < m; s++)
  {
    double pm, pf;
    if (!snp_column(geno + s * n, n, is_diploid, cc, &pm, &pf)) continue;
    used++;
    for (int i = 0; i < n; i++)
    {
      const double v = cc[i];
      for (int j = 0; j <= i; j++) xtx[i * n + j] += v * cc[j];
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) xtx[j * n + i] = xtx[i * n + j];
  double tr = 0.0;
  for (int i = 0; i < n; i++) tr += xtx[i * n + i];
  const double y = tr / (double)(n - 1);
  const double inv = 1.0 / y;
  for (int i = 0; i < n * n; i++) xtx[i] = xtx[i] * inv;
  free(cc);
  return used;
}

void kmo_sym_eig(const double* a, int n, double* evals, double* evecs)
{
  double* A = (double*)malloc(sizeof(double) * n * n);
  double* V = (double*)malloc(sizeof(double) * n * n);
  memcpy(A, a, sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = i == j;
  for (int sweep = 0; sweep < 60; sweep++)
  {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; i++)
    {
      diag += A[i * n + i] * A[i * n + i];
      for (int j = 0; j < i; j++) off += A[i * n + j] * A[i * n + j];
    }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++)
      {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++)
        {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++)
        {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++)
        {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  int* order = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 1; i < n; i++)
  { /* insertion sort, descending eigenvalue */
    int o = order[i], j = i;
    while (j > 0 && A[order[j - 1] * n + order[j - 1]] < A[o * n + o]) { order[j] = order[j - 1]; j--; }
    order[j] = o;
  }
  for (int j = 0; j < n; j++)
  {
    const int col = order[j];
    evals[j] = A[col * n + col];
    int big = 0;
    for (int i = 1; i < n; i++)
      if (fabs(V[i * n + col]) > fabs(V[big * n + col])) big = i;
    const double sgn = V[big * n + col] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < n; i++) evecs[j * n + i] = sgn * V[i * n + col];
  }
  free(A); free(V); free(order);
}

void kmo_smartpca_pcs(const uint8_t* geno, uint64_t m, int n, int is_diploid, const double* evecs,
                      int numeigs, double* pcs)
{
  double* cc = (double*)malloc(sizeof(double) * n);
  double* xmean = (double*)malloc(sizeof(double) * m);
  double* xfancy = (double*)malloc(sizeof(double) * m);
  double* ffvecs = (double*)calloc((size_t)numeigs * m, sizeof(double));
  double* fxscal = (double*)calloc(numeigs, sizeof(double));
  double* xrow = (double*)malloc(sizeof(double) * m);
  /* fvecs = 10 * evecs (qpsubs.c:428-447, plotmode NO); ffvecs[j][snp] = sum_k fvecs[j][k]*cc[k]
   * with cc from getcolxf (zero for ignored SNPs), smartpca.c:1186-1195 */
  for (uint64_t s = 0; s < m; s++)
  {
    const int used = snp_column(geno + s * n, n, is_diploid, cc, &xmean[s], &xfancy[s]);
    if (!used) continue;
    for (int j = 0; j < numeigs; j++)
      for (int k = 0; k < n; k++) ffvecs[j * m + s] += (evecs[j * n + k] * 10.0) * cc[k];
  }
  /* smartpca.c:1204-1228 */
  for (int i = 0; i < n; i++)
  {
    for (uint64_t s = 0; s < m; s++)
    { /* loadxdataind + fixxrow (qpsubs.c:282-303) */
      double v = (double)geno[s * n + i] * xfancy[s];
      if (v < -0.1) v = 0.0; else v -= xmean[s];
      xrow[s] = v;
    }
    for (int j = 0; j < numeigs; j++)
    {
      double y = 0.0;
      for (uint64_t s = 0; s < m; s++) y += xrow[s] * ffvecs[j * m + s];
      pcs[i * numeigs + j] = y; /* fxvecs, rescaled below (bcoeffs = fxscal * the same dot) */
      fxscal[j] += y * y;
    }
  }
  for (int j = 0; j < numeigs; j++) fxscal[j] = 1.0 / sqrt(fxscal[j]);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < numeigs; j++)
    {
      char buf[64];
      snprintf(buf, sizeof buf, "%10.4f", fxscal[j] * pcs[i * numeigs + j]);
      pcs[i * numeigs + j] = strtod(buf, NULL);
    }
  free(cc); free(xmean); free(xfancy); free(ffvecs); free(fxscal); free(xrow);
}
