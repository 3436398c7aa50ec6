/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64, -ffp-contract=off) of the population-stratification half of
 * `kmdiff diff --pop-correction`, the parity oracle for the CUDA
 * sampler / GRM / IRLS kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.
 *
 * Parity status
 *   sampler, glm_irls, pop_strat_corrector (features, standardize, apply): PINNED — bit-exact
 *     against the unmodified reference compiled behind oracle/ref_shim.cpp (libstdc++'s
 *     default_random_engine / uniform_real_distribution, src/linear_model.cpp, src/popstrat.cpp),
 *     tests/test_popstrat_oracle.py, plus golden vectors made from that library
 *     (tests/golden/popstrat_vectors.json, generator committed).
 *   smartpca (HAWK's modified EIGENSOFT 6.0.1, thirdparty/hawk/EIG6.0.1-Hawk): PINNED — `make -C oracle
 *     smartpca` builds the program from the reference's vendored sources (LAPACK from scipy's bundled
 *     OpenBLAS, the GSL-only "fastmode" stubbed).  tests/test_smartpca_reference.py holds kmo_smartpca_*
 *     to its pcs.evec / eigenvalue output (tests/golden/smartpca_vectors.json, generator committed): every
 *     printed digit equal, eigenvector signs aside (LAPACK's dspev leaves them unspecified), and — with
 *     oracle/_ref/kmdiff next to it — to the reference binary's whole `--pop-correction` run (.geno,
 *     pcs.evec, LRT and GLM p-values of the spill files, bit for bit).
 *
 * Citations: "popstrat.hpp" = include/kmdiff/popstrat.hpp, "popstrat.cpp" = src/popstrat.cpp,
 * "linear_model.cpp" = src/linear_model.cpp, "smartpca.c" =
 * thirdparty/hawk/EIG6.0.1-Hawk/src/eigensrc/smartpca.c, "qpsubs.c" = .../src/qpsubs.c.
 */
#ifndef KMD_POPSTRAT_ORACLE_H
#define KMD_POPSTRAT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Sampler (popstrat.hpp:107-146) ---------------------------------------------------------
 * std::default_random_engine == minstd_rand0 (x <- 16807 x mod 2^31-1, libstdc++ random.h) seeded
 * only when seed != 0 (popstrat.hpp:113-117; default state 1, seed % m == 0 -> 1), driven by
 * uniform_real_distribution<double>(0,1) == generate_canonical<double,53>: two engine calls per
 * draw, (x1-1) + (x2-1)*R over R*R with R = 2147483646 (libstdc++ random.tcc). */
typedef struct { uint64_t state; } kmo_sampler;
void kmo_sampler_init(kmo_sampler* s, uint64_t seed);
double kmo_sampler_draw(kmo_sampler* s);
/* sample() for rows first_row.. first_row+n-1 of the visiting order (draw index == row index):
 * out[i] = draw < kmer_pca.  Returns how many were taken. */
uint64_t kmo_sampler_rows(uint64_t seed, double kmer_pca, uint64_t first_row, uint64_t n, uint8_t* out);

/* ---- linear_model.cpp ---------------------------------------------------------------------- */
/* :191-195  1/(1+pow(M_E,-x)) */
double kmo_sigmoid(double x);
/* :94-137 Doolittle LU without pivoting; lower/upper are n*n row-major */
void kmo_lu(const double* a, int n, double* lower, double* upper);
/* :139-189 inverse through LU; flags as the reference sets them (det == 0 / isnan(det)) */
void kmo_inverse(const double* a, int n, double* inv, int* singular, int* is_nan);
/* :297-410 glm_irls(x, y, max_iters): x is n*p row-major.  weight[p] = returned model. */
void kmo_glm_irls(const double* x, const double* y, int n, int p, int max_iters, double* weight,
                  int* singular, int* is_nan, double* ret_error, int* iters);

/* ---- pop_strat_corrector (popstrat.hpp:148-372, popstrat.cpp:136-370) ---------------------- */
typedef struct
{
  int n, nb_controls, nb_cases, npc;
  int null_p, alt_p;     /* 1+npc+1, 1+npc+2 (no covariates, gender unknown: popstrat.cpp:270-313) */
  int max_iter;          /* s_max_iter = 100 */
  double epsilon;        /* s_epsilon = 1e-30 */
  double* null_features; /* n * null_p, after standardize() */
  double* alt_features;  /* n * alt_p, last column filled per k-mer */
  double* Y;             /* Control -> 1, Case -> 0 (popstrat.cpp:164-171) */
  double* totals;        /* per-sample totals as double */
  double* null_model;    /* null_p */
} kmo_popstrat;

/* Z: n rows x 10 principal components as load_Z reads them (popstrat.cpp:153-162);
 * totals: per-sample totals, controls first.  Runs init_global_features incl. standardize and the
 * null-model IRLS.  Returns 0 on success. */
int kmo_popstrat_init(kmo_popstrat* ps, int nb_controls, int nb_cases, const uint64_t* totals,
                      const double* Z, int npc);
void kmo_popstrat_free(kmo_popstrat* ps);
/* pop_strat_corrector::apply(KmerSign&) (popstrat.hpp:249-333): counts[n] as doubles (the record's
 * m_counts_ratio); returns the corrected p-value.  Optional outputs: model (alt_p weights),
 * iterations, log-likelihood ratio. */
double kmo_popstrat_apply(const kmo_popstrat* ps, const double* counts, double* model, int* iters,
                          double* llr);

/* ---- smartpca as HAWK ships it (pinned against the real program, see header) ----------------- */
/* geno: m rows (sampled k-mers) x n individuals, values 0/1 (what EigGenoFile writes,
 * popstrat.hpp:56-80).  Follows smartpca.c:907-1025: per SNP getcolxz (:2598-2700) + fvadjust
 * (:1684-1800, usenorm YES), SNPs with no carrier dropped (:921), XTX += col col^T in input order
 * (block_increment_normal :2985-3004), symmetrise, divide by trace/(n-1) (:1019-1023).
 * xtx: n*n row-major.  Returns the number of SNPs used. */
uint64_t kmo_smartpca_xtx(const uint8_t* geno, uint64_t m, int n, int is_diploid, double* xtx);
/* symmetric eigen-decomposition, eigenvalues descending (eigsubs.c:35-51 eigvecs: LAPACK dspev on
 * -A), cyclic Jacobi here; evecs[j*n+i] = component i of eigenvector j, sign: largest-magnitude
 * component positive (the reference's sign is whatever LAPACK returns — unspecified). */
void kmo_sym_eig(const double* a, int n, double* evals, double* evecs);
/* the numeigs (=10) output coordinates of every individual (smartpca.c:1140-1329: fvecs -> ffvecs
 * -> fxscal -> bcoeffs) printed with %10.4f and parsed back (evec2pca.perl + load_Z).
 * pcs: n x numeigs row-major. */
void kmo_smartpca_pcs(const uint8_t* geno, uint64_t m, int n, int is_diploid, const double* evecs,
                      int numeigs, double* pcs);

#ifdef __cplusplus
}
#endif
#endif
