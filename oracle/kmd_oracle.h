/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64, -ffp-contract=off) of the `kmdiff diff` hot path of
 * tlemane/kmdiff, used as the parity oracle for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own golden values (tests/factorial_test.cpp, model_test.cpp,
 *       corrector_test.cpp — see tests/golden/reference_tests.json),
 *   (2) oracle/_ref/libkmdiff_ref.so = the unmodified reference compiled from /root/reference
 *       (bit-exact on p-values over seeded rows), and
 *   (3) tests/golden JSON vectors generated from that library (script committed).
 *
 * Each function cites the reference file:line it restates (paths relative to the reference root;
 * "alglib/" = thirdparty/alglib/src/).
 */
#ifndef KMD_ORACLE_H
#define KMD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kmdiff::Significance, include/kmdiff/kmer.hpp:33-38 */
enum { KMO_CONTROL = 0, KMO_CASE = 1, KMO_NO = 2 };

/* kmdiff::CorrectionType, include/kmdiff/correction.hpp */
enum { KMO_NOTHING = 0, KMO_BONFERRONI = 1, KMO_BENJAMINI = 2, KMO_SIDAK = 3, KMO_HOLM = 4 };

/* src/log_factorial_table.cpp:13-22 — sum of log(j), j = k down to 2 */
double kmo_log_factorial(uint64_t k);
/* src/log_factorial_table.cpp:5-11 */
void kmo_lf_table_fill(double* table, uint64_t size);
/* include/kmdiff/log_factorial_table.hpp:14-18 — table hit, else recompute */
double kmo_lf_lookup(const double* table, uint64_t size, uint64_t i);

/* alglib/specialfunctions.cpp:3744-3860 (Cephes lgam) */
double kmo_lngamma(double x, double* sgngam);
/* alglib/specialfunctions.cpp:4578-4619 (Cephes igam) */
double kmo_igam(double a, double x);
/* alglib/specialfunctions.cpp:4655-4739 (Cephes igamc) */
double kmo_igamc(double a, double x);
/* alglib/specialfunctions.cpp:9559-9567 */
double kmo_chisqc(double v, double x);

/* PoissonLikelihood state: include/kmdiff/model.hpp:105-118,182-191 */
typedef struct
{
  uint32_t nb_controls;
  uint32_t nb_cases;
  uint64_t sum_controls; /* Σ per-sample totals of controls (model.hpp:185-186) */
  uint64_t sum_cases;    /* Σ per-sample totals of cases    (model.hpp:188)     */
  const double* lf_table;
  uint64_t lf_size;
} kmo_model;

/* include/kmdiff/model.hpp:142-176 on one row of counts (count_bytes = 1, 2 or 4) */
void kmo_poisson_process(const kmo_model* m, const void* counts, int count_bytes,
                         double* p, int32_t* sign, double* mean_control, double* mean_case);

/* Same over an AoS matrix payload (kmtricks io/matrix_file.hpp:144-154): per row kmer_words*8
 * bytes of k-mer then (nc+nk) counts. */
void kmo_model_rows(const kmo_model* m, const uint8_t* rows, uint64_t n_rows, int kmer_words,
                    int count_bytes, double* p, int32_t* sign, double* mc, double* kk);

/* Survivor record == fields of KmerSign (include/kmdiff/kmer.hpp:88-127) + row index */
typedef struct
{
  uint64_t kmer[2];
  double p;
  int32_t sign;
  int32_t pad_;
  double mean_control;
  double mean_case;
  uint64_t row;
} kmo_survivor;

/* The per-row loop of matrix_proxy::merge + diff_observer::process
 * (include/kmdiff/merge.hpp:194-203, 68-103): keep rows with p <= thr (note <=), count signs
 * (NO counted as case).  Returns the number of survivors found (may exceed cap; only the first
 * cap are stored). */
uint64_t kmo_diff_rows(const kmo_model* m, const uint8_t* rows, uint64_t n_rows, int kmer_words,
                       int count_bytes, double thr, uint64_t row_base, kmo_survivor* out,
                       uint64_t cap, uint64_t* n_control, uint64_t* n_case);

/* Correctors: src/corrector.cpp:9-12, 27-35, 50-53, 68-71, 86-89 (stateful like the reference) */
typedef struct
{
  int type;
  double threshold;
  uint64_t total; /* m_total (holm decrements it) */
  uint64_t rank;  /* benjamini m_rank, starts at 1 */
} kmo_corrector;

void kmo_corrector_init(kmo_corrector* c, int type, double threshold, uint64_t total);
int kmo_corrector_apply(kmo_corrector* c, double p);

/* aggregator::run / sorted_aggregator::run (include/kmdiff/aggregator.hpp:137-220, 240-322):
 * for NOTHING/BONFERRONI/SIDAK every survivor is tested independently and kept survivors stay in
 * input order; for BENJAMINI/HOLM survivors are visited in ascending p (stable on input order
 * for ties) and the walk stops at the first rejection.  Writes the indices of the kept survivors,
 * in output order, to keep_idx; returns how many. */
uint64_t kmo_aggregate(const double* p, uint64_t n, int type, double threshold, uint64_t total,
                       uint64_t* keep_idx);

/* km::Kmer::to_string (kmtricks kmer.hpp:38, 797-810): 2 bits/base A0 C1 T2 G3, first base most
 * significant; words[0] = least-significant 64 bits.  out must hold k+1 bytes. */
void kmo_kmer_to_string(const uint64_t* words, int kmer_words, int k, char* out);

#ifdef __cplusplus
}
#endif
#endif
