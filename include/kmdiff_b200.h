/*
 * kmdiff_b200 — C-ABI of the B200-native `kmdiff diff` hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types, integer status
 * codes and a last-error string, no exceptions across it.  The reference (tlemane/kmdiff, pure
 * C++17/CPU) has no FFI of its own for this path; each entry point below names the reference
 * interface it replaces (file:line relative to the reference root; "kmtricks/" =
 * thirdparty/kmtricks/include/kmtricks/).  INTEGRATION.md shows the binding a reference
 * maintainer would add inside include/kmdiff/cmd/diff.hpp.
 *
 * Threading: a context is thread-compatible (one thread at a time per context); distinct
 * contexts are independent.  All device work of a context is ordered on one CUDA stream
 * (its own, or the caller's via kmd_set_stream).
 *
 * Row layout everywhere = the payload of a kmtricks count matrix (kmtricks/io/matrix_file.hpp:
 * 144-154): per row kmer_words*8 bytes of k-mer (little-endian u64 words, word 0 least
 * significant) followed by nb_controls+nb_cases counts of count_bytes each, controls first.
 */
#ifndef KMDIFF_B200_H
#define KMDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMD_ABI_VERSION 1

/* status codes */
enum
{
  KMD_OK = 0,
  KMD_ERR_ARG = 1,      /* invalid argument / configuration                     */
  KMD_ERR_CUDA = 2,     /* a CUDA runtime call failed (see kmd_last_error)       */
  KMD_ERR_NOMEM = 3,    /* host or device allocation failed                      */
  KMD_ERR_OVERFLOW = 4, /* more candidates/survivors than max_survivors          */
  KMD_ERR_STATE = 5,    /* call order violated (e.g. fetch before correct)       */
  KMD_ERR_NO_DEVICE = 6 /* no usable sm_100 device; there is NO CPU fallback     */
};

/* kmdiff::Significance (include/kmdiff/kmer.hpp:33-38) */
enum { KMD_SIGN_CONTROL = 0, KMD_SIGN_CASE = 1, KMD_SIGN_NO = 2 };

/* kmdiff::CorrectionType (include/kmdiff/correction.hpp) */
enum
{
  KMD_CORR_NOTHING = 0,
  KMD_CORR_BONFERRONI = 1,
  KMD_CORR_BENJAMINI = 2,
  KMD_CORR_SIDAK = 3,
  KMD_CORR_HOLM = 4
};

/* kmd_config.flags */
enum
{
  /* keep the N counts of every prefilter survivor (KmerSign::m_counts_ratio, kmer.hpp:118-123: what a default
   * WITH_POPSTRAT build of the reference appends to every spill record); read back with kmd_survivor_counts.
   * Implied by kmd_popstrat_enable.  Costs max_survivors * N * count_bytes of device memory. */
  KMD_FLAG_KEEP_COUNTS = 1
};

typedef struct kmd_context kmd_context;

/* Replaces the constructor arguments of PoissonLikelihood (include/kmdiff/model.hpp:105-118),
 * diff_observer (include/kmdiff/merge.hpp:48-66) and global_merge (merge.hpp:211-236). */
typedef struct kmd_config
{
  uint32_t abi_version;   /* KMD_ABI_VERSION                                               */
  int32_t device;         /* CUDA device ordinal                                           */
  uint32_t nb_controls;   /* -1/--nb-controls                                              */
  uint32_t nb_cases;      /* -2/--nb-cases                                                 */
  uint32_t kmer_size;     /* k                                                             */
  uint32_t kmer_words;    /* on-disk k-mer words = ceil(k/32) (matrix_file.hpp:109); 1 or 2 */
  uint32_t count_bytes;   /* 1, 2 or 4 (compile-time MAX_C in the reference)               */
  uint32_t flags;         /* KMD_FLAG_*                                                    */
  uint64_t sum_controls;  /* Σ per-sample totals of the controls (model.hpp:185-186)       */
  uint64_t sum_cases;     /* Σ per-sample totals of the cases (model.hpp:188)              */
  double threshold;       /* prefilter: significance / cutoff (cmd/diff.hpp:146)           */
  const double* lf_table; /* host log-factorial table in reference order, or NULL: the     */
  uint64_t lf_size;       /*   library then builds lf_size entries itself (--log-factorial) */
  uint64_t max_survivors; /* capacity of the device survivor buffers; 0 = 1<<22            */
} kmd_config;

/* KmerSign (include/kmdiff/kmer.hpp:88-127) plus the global row index */
typedef struct kmd_survivor
{
  uint64_t kmer[2];    /* kmer[0] = least-significant word                        */
  double p_value;      /* PoissonLikelihood::process p-value                      */
  int32_t sign;        /* KMD_SIGN_*                                              */
  uint32_t partition;  /* partition tag given to kmd_scan_*                       */
  double mean_control; /* sum(ctrl) * sum_cases / sum_controls (model.hpp:164)    */
  double mean_case;    /* sum(case)                                               */
  uint64_t row;        /* row_base + index inside the scanned batch               */
} kmd_survivor;

/* merge.hpp:105-113, 319-337 (total / nb_sign / signs) */
typedef struct kmd_counters
{
  uint64_t rows;       /* rows tested so far = m for the correctors (merge.hpp:316) */
  uint64_t candidates; /* rows that passed the fp32 screen (diagnostic)             */
  uint64_t survivors;  /* rows with p <= threshold                                  */
  uint64_t n_control;  /* survivors with sign CONTROL                               */
  uint64_t n_case;     /* survivors with sign CASE or NO (merge.hpp:95-98)          */
} kmd_counters;

/* ---- lifecycle ------------------------------------------------------------------------------ */

int kmd_abi_version(void);
/* number of CUDA devices visible to the library; 0 => every compute entry point fails loudly */
int kmd_device_count(void);

int kmd_create(const kmd_config* cfg, kmd_context** out);
void kmd_destroy(kmd_context* ctx);
/* message of the last failing call on ctx (ctx == NULL: last failing kmd_create of this thread) */
const char* kmd_last_error(const kmd_context* ctx);
/* run all device work of ctx on the caller's cudaStream_t (NULL restores the context's own) */
int kmd_set_stream(kmd_context* ctx, void* cuda_stream);
int kmd_sync(kmd_context* ctx);
/* forget survivors and counters (start of a new `diff` run, cmd/diff.hpp:66) */
int kmd_reset(kmd_context* ctx);

/* ---- the partition loop: global_merge::merge(paths) -> matrix_proxy::merge ->
 *      diff_observer::process (merge.hpp:341-397, 194-203, 68-103) --------------------------- */

/* Test every row of a device-resident batch (Poisson LRT + chi-square p-value + prefilter) and
 * append the survivors to the context.  d_rows must be 16-byte aligned.  Asynchronous. */
int kmd_scan_device(kmd_context* ctx, const void* d_rows, uint64_t n_rows, uint64_t row_base,
                    uint32_t partition);
/* Same for a HOST buffer: chunked host->device copies (async when h_rows is pinned, see
 * kmd_host_alloc) overlapped with the scan of the previous chunk.  Returns after the last chunk
 * has been enqueued and the host buffer is no longer needed. */
int kmd_scan_host(kmd_context* ctx, const void* h_rows, uint64_t n_rows, uint64_t row_base,
                  uint32_t partition);
/* The same without waiting for the copy: returns as soon as the copies and scans are enqueued.  h_rows (pinned) must
 * stay untouched until kmd_host_wait(ctx, *ticket) returns — the only entry point that may be called from a thread
 * other than the one driving ctx (reader threads recycling their staging buffers while the feeder keeps enqueuing;
 * the host-side replacement of the one-row-at-a-time istream reads of matrix_proxy::merge, merge.hpp:194-203).
 * At most 255 tickets may be outstanding. */
int kmd_scan_host_async(kmd_context* ctx, const void* h_rows, uint64_t n_rows, uint64_t row_base,
                        uint32_t partition, uint64_t* ticket);
int kmd_host_wait(kmd_context* ctx, uint64_t ticket);
int kmd_counters_get(kmd_context* ctx, kmd_counters* out);
/* prefilter survivors ordered by row (what the FileAccumulators hold, accumulator.hpp:156-285) */
int kmd_survivors(kmd_context* ctx, kmd_survivor* out, uint64_t cap, uint64_t* n);
/* Append n host-side survivor records (FileAccumulator::push, accumulator.hpp:176-188) and count
 * rows_tested more rows as tested: the entry for survivors that were not produced by kmd_scan_* —
 * rows judged on the host by a user IModel plugin (--cmodel), or spill files of an earlier run being
 * re-corrected.  Sign counters are updated like merge.hpp:95-98.  The correction, sorting and
 * output stages then treat them exactly like scanned survivors. */
int kmd_push_survivors(kmd_context* ctx, const kmd_survivor* records, uint64_t n, uint64_t rows_tested);
/* the same with the records' counts (n x N values of count_bytes, parallel to records; may be NULL): spill files with
 * the WITH_POPSTRAT tail re-opened for a new correction or a new stratification fit (cmd/diff.hpp:342-348) */
int kmd_push_survivors2(kmd_context* ctx, const kmd_survivor* records, const void* counts, uint64_t n,
                        uint64_t rows_tested);

/* ---- per-sample count files instead of matrices: global_merge::merge() -> km::KmerMerger
 *      (merge.hpp:237-309; kmtricks/merge.hpp:102-361; record layout kmtricks/io/kmer_file.hpp:127-137) ---- */

/* N-way merge of the samples' sorted (k-mer, count) streams into dense matrix rows on the device,
 * then the same test as kmd_scan_device on those rows (they never leave HBM).  streams[s] points to
 * n_records[s] HOST records of sample s (controls first, nb_controls + nb_cases streams): kmer_words*8
 * bytes of k-mer then count_bytes of count, ascending and unique within a stream.  A count below
 * ab_min[s] is reported as 0 and a row is kept when at least one sample is solid (KmerMerger with
 * recurrence_min = 1, save_if = 0, as merge.hpp:266-267 constructs it).  The streams of one call must
 * cover one k-mer range completely (a whole partition, or the records <= a common bound of every
 * stream) and hold fewer than 2^31 records in total.  n_rows (may be NULL) = rows produced; they are
 * numbered row_base, row_base+1, ... in ascending k-mer order.  Synchronous for the host buffers. */
int kmd_merge_scan_host(kmd_context* ctx, const void* const* streams, const uint64_t* n_records,
                        const uint32_t* ab_min, uint64_t row_base, uint32_t partition, uint64_t* n_rows);
/* The same for the samples' files of a WHOLE partition as they lie on disk: frames[s] = the payload of
 * counts/partition_p/<sample>.kmer.lz4 after its 41-byte header (HOST; pinned for full PCIe rate), frame_bytes[s] its
 * size; compressed[s] (NULL = all) says whether it is an LZ4 frame (km::lz4_stream: 64 KiB linked blocks,
 * kmtricks/io/lz4_stream.hpp:42,128-140) or raw records.  Compressed payloads cross PCIe compressed and are decoded on
 * the device, one warp per stream — the host decoder manages ~1 GB/s per thread, a GPU ingests ~55 GB/s.
 * max_records != 0: when the partition decodes to more records than that, nothing is merged and KMD_ERR_OVERFLOW is
 * returned (*n_records says how many): the caller then streams the partition in k-mer ranges through
 * kmd_merge_scan_host.  n_rows / n_records may be NULL. */
int kmd_merge_scan_lz4_host(kmd_context* ctx, const void* const* frames, const uint64_t* frame_bytes,
                            const uint8_t* compressed, const uint32_t* ab_min, uint64_t row_base, uint32_t partition,
                            uint64_t max_records, uint64_t* n_rows, uint64_t* n_records);
/* kmd_merge_scan_host for records that are already in device memory: d_records = the streams' records back to back
 * (sample 0's n_records[0] records, then sample 1's, ...), 4-byte aligned; it is only read and must stay untouched
 * until the call returns (the call synchronises once, for the row count). */
int kmd_merge_scan_device(kmd_context* ctx, const void* d_records, const uint64_t* n_records, const uint32_t* ab_min,
                          uint64_t row_base, uint32_t partition, uint64_t* n_rows);
/* Start the host->device copies of the NEXT kmd_merge_scan_host (frames == 0: blobs[s] = the records of sample s,
 * blob_bytes[s] = n_records[s] * record size) or kmd_merge_scan_lz4_host (frames != 0: all payloads compressed) call on
 * the copy stream and return at once: they overlap the merge the context is still busy with.  The next merge call
 * that is given exactly these pointers and sizes uses the prefetched copy; any other call ignores it.  The host
 * buffers must stay untouched until that call has returned. */
int kmd_merge_prefetch(kmd_context* ctx, const void* const* blobs, const uint64_t* blob_bytes, int frames);
/* the device LZ4 decoder alone: nb_controls + nb_cases frames in, their decoded records back to back out (HOST);
 * device_ms (may be NULL) = device time of the two decode passes (tests, tools) */
int kmd_lz4_decode_host(kmd_context* ctx, const void* const* frames, const uint64_t* frame_bytes, void* records_out,
                        uint64_t cap_records, uint64_t* n_records, float* device_ms);
/* The merged rows themselves (what KmerMerger::write_as_bin would put in a matrix file,
 * kmtricks/merge.hpp:262-272), copied back to the host: tests, tools.  No test is run. */
int kmd_merge_rows_host(kmd_context* ctx, const void* const* streams, const uint64_t* n_records,
                        const uint32_t* ab_min, void* rows_out, uint64_t cap_rows, uint64_t* n_rows);

/* device times (CUDA events on the context's stream) of the last kmd_merge_* call: record unpacking (nothing when
 * the first level of the union tree reads the k-mers inside the records), the levels / passes of the tree, row
 * numbering + dense-row writing (the latter includes one host round trip for the row count); the number of levels
 * and records of that call.  Any output may be NULL. */
int kmd_merge_ms(kmd_context* ctx, float* unpack_ms, float* passes_ms, float* rows_ms, uint32_t* n_passes,
                 uint64_t* n_records);

/* Which merge the last kmd_merge_* call ran and what its tree moved.  algo: 0 = union tree (kmd_union.cuh: the k-mers
 * are merged pairwise with the duplicates dropped, the counts are scattered into the rows; taken when every record is
 * solid), 1 = pairwise tree of (k-mer, payload) pairs (kmd_merge.cuh; also with KMD_MERGE_PAIRWISE=1 in the
 * environment).  rows_mode (union tree): 1 = the rows were assembled in shared memory and written whole (dense cohorts),
 * 2 = the counts were scattered into zeroed rows (sparse ones); KMD_UNION_ROWS=1|2 forces one.  keys_in / keys_out: keys
 * read / written by all levels of the union tree (0 for algo 1, where every pass reads and writes every record).  kmd_merge_ms' three times are then: key extraction | levels | zeroing the rows
 * + scatter (incl. the host round trip for the row count).  Any output may be NULL. */
int kmd_merge_info(kmd_context* ctx, uint32_t* algo, uint32_t* rows_mode, uint64_t* keys_in, uint64_t* keys_out);

/* diagnostics for bench.py: kernels launched on behalf of ctx so far, and the device time of the
 * most recent lrt_scan kernel launch (CUDA events on the context's stream) */
uint64_t kmd_launch_count(const kmd_context* ctx);
int kmd_scan_kernel_ms(kmd_context* ctx, float* ms);

/* ---- IModel::process, batched (include/kmdiff/imodel.hpp:37, model.hpp:142-176) ------------ */

/* p-value, sign and the two "means" of EVERY row (no screen, no prefilter); rows on the device
 * (out arrays on the device too) or on the host (out arrays on the host). */
int kmd_model_rows_device(kmd_context* ctx, const void* d_rows, uint64_t n_rows, double* d_p,
                          int32_t* d_sign, double* d_mean_control, double* d_mean_case);
int kmd_model_rows_host(kmd_context* ctx, const void* h_rows, uint64_t n_rows, double* p,
                        int32_t* sign, double* mean_control, double* mean_case);

/* ---- do_correction: make_corrector + (sorted_)aggregator::run
 *      (cmd/diff.hpp:227-260, src/corrector.cpp, aggregator.hpp:137-322) --------------------- */

/* Apply `correction` with significance level `significance` and m = total_rows to the prefilter
 * survivors held by ctx.  NOTHING/BONFERRONI/SIDAK: independent strict `<` tests, output in row
 * order.  BENJAMINI/HOLM: device radix sort by p, walk ascending, stop at the first rejection
 * (the reference's step-down-with-break).  Returns counts of kept control / case k-mers. */
int kmd_correct(kmd_context* ctx, int correction, double significance, uint64_t total_rows,
                uint64_t* n_control, uint64_t* n_case);
/* kept survivors in output order (row order, or ascending p for BENJAMINI/HOLM) */
int kmd_significant(kmd_context* ctx, kmd_survivor* out, uint64_t cap, uint64_t* n);

/* Multi-GPU BENJAMINI/HOLM (partitions sharded over ranks): every rank exports its survivors'
 * p-values ascending, the host side exchanges them (NCCL all-gather), computes the first
 * rejected p-value with kmd_first_reject, and every rank keeps its survivors with p < p_cut. */
int kmd_sorted_pvalues(kmd_context* ctx, double* out, uint64_t cap, uint64_t* n);
int kmd_keep_below(kmd_context* ctx, double p_cut, uint64_t* n_control, uint64_t* n_case);
/* The same exchange without leaving the devices: the survivors' p-values (accumulator order, unsorted)
 * copied into a caller-owned DEVICE buffer on the context's stream; and the ascending walk over a
 * DEVICE array of merged, sorted p-values (+inf padding allowed at the end), one thread per rank
 * (both rank tests are monotone, so the first failing position is a min-reduction).  n_accepted =
 * index of the first rejection, p_cut = its p-value (+inf if none).  Synchronises the stream. */
int kmd_survivor_pvalues_device(kmd_context* ctx, double* d_out, uint64_t cap, uint64_t* n);
int kmd_first_reject_device(kmd_context* ctx, const double* d_p_sorted, uint64_t n, int correction,
                            double significance, uint64_t total_rows, uint64_t* n_accepted, double* p_cut);
/* The whole multi-GPU correction stream-ordered on the devices, one host synchronisation per run
 * (replaces global_merge's total, merge.hpp:316 + 390-397, and sorted_aggregator::run's global heap,
 * aggregator.hpp:282-310, when partitions are sharded over ranks; any corrector):
 *   kmd_exchange_export   this rank's slot = KMD_EXCHANGE_HEADER u64 {rows, controls, cases, survivors, overflow, 0,0,0}
 *                         followed by `cap` p-values (accumulator order, +inf padded); d_slot is DEVICE memory of
 *                         KMD_EXCHANGE_HEADER + cap words.  Asynchronous on the context's stream (kmd_get_stream).
 *   (caller)              all-gathers the world slots into d_all (NCCL / peer copies), ordered after the export and
 *                         before the next call with events on that stream.
 *   kmd_exchange_correct  m = sum of the slots' rows; BENJAMINI / HOLM: device sort of the union, first rejected rank
 *                         (same walk as kmd_first_reject), p_cut = its p-value; other correctors: p_cut from m;
 *                         this rank's survivors with p < p_cut are kept.  Asynchronous.
 *   kmd_exchange_result   synchronises and reports.  overflow != 0: some rank had more than cap survivors, nothing was
 *                         decided — repeat the three calls with cap >= max_rank_survivors.  Otherwise kmd_significant
 *                         returns this rank's kept records (ascending p, ties by row, for BENJAMINI / HOLM). */
#define KMD_EXCHANGE_HEADER 8
typedef struct kmd_exchange_decision
{
  uint64_t total_rows;         /* m: rows tested by all ranks                          */
  uint64_t total_survivors;    /* prefilter survivors of all ranks                     */
  uint64_t max_rank_survivors; /* largest per-rank survivor count                      */
  uint64_t n_accepted;         /* BENJAMINI / HOLM: global number accepted, else 0      */
  double p_cut;                /* survivors with p < p_cut are significant             */
  uint64_t kept_control;       /* this rank's significant control / case k-mers        */
  uint64_t kept_case;
  uint32_t overflow;
  uint32_t reserved;
} kmd_exchange_decision;
void* kmd_get_stream(kmd_context* ctx); /* the cudaStream_t all work of ctx is ordered on */
int kmd_exchange_export(kmd_context* ctx, uint64_t* d_slot, uint64_t cap);
int kmd_exchange_correct(kmd_context* ctx, const uint64_t* d_all, uint32_t world, uint64_t cap, int correction,
                         double significance);
int kmd_exchange_result(kmd_context* ctx, kmd_exchange_decision* out);
/* The same exchange inside ONE process that drives several devices (n contexts, one per device — the host program):
 * every context exports its slot, pulls the other contexts' slots with peer-to-peer copies (NVLink where the topology
 * has it) ordered by events, and corrects.  Synchronises; decisions[i] belongs to ctxs[i]. */
int kmd_exchange_local(kmd_context* const* ctxs, uint32_t n, int correction, double significance,
                       kmd_exchange_decision* decisions);
/* host helper: walk p_sorted (ascending, all ranks merged) exactly like sorted_aggregator::run;
 * returns the number accepted and writes the first rejected p-value (+inf if none) */
uint64_t kmd_first_reject(const double* p_sorted, uint64_t n, int correction, double significance,
                          uint64_t total_rows, double* p_cut);

/* ---- --pop-correction: Sampler -> smartpca Gram matrix -> PCs -> per-k-mer logistic GLM
 *      (cmd/diff.hpp:168-224, 307-362; include/kmdiff/popstrat.hpp; src/popstrat.cpp;
 *       src/linear_model.cpp; thirdparty/hawk/EIG6.0.1-Hawk/src/eigensrc/smartpca.c) ------------ */

/* Sampler<MAX_C> constructor arguments (popstrat.hpp:107-118) + smartpca's `-p` and capacity */
typedef struct kmd_popstrat_config
{
  double kmer_pca;      /* --kmer-pca: probability that a row is sampled                   */
  uint64_t seed;        /* --random-seed; 0 = default-constructed engine (popstrat.hpp:116) */
  uint32_t is_diploid;  /* --ploidy 2 (run_eigenstrat_smartpca's is_diploid)               */
  uint32_t n_pc;        /* --n-pc: principal components used as covariates (<= 10)          */
  uint64_t max_sampled; /* capacity of the presence matrix; 0 = 1<<21 sampled k-mers        */
} kmd_popstrat_config;

/* Turn the context into diff_observer_strat (merge.hpp:129-177): from now on every kmd_scan_* also
 * runs the Sampler on its rows (draw index = row_base + row, i.e. the reference's `-t 1` visiting
 * order when row_base is the number of rows of the partitions visited before) and keeps the N counts
 * of every prefilter survivor.  Must be called before the first scan. */
int kmd_popstrat_enable(kmd_context* ctx, const kmd_popstrat_config* cfg);
/* Decouple the Sampler's draw index from row_base: the next scanned row takes draw `next_draw`, every
 * further row the following one, whatever row_base the scans are given.  The host program numbers
 * survivor rows (partition << 40 | row) but draws in the reference's `-t 1` visiting order, i.e. with
 * next_draw = number of rows in the partitions visited before (merge.hpp:341-397). */
int kmd_popstrat_set_draw_index(kmd_context* ctx, uint64_t next_draw);
int kmd_popstrat_sampled(kmd_context* ctx, uint64_t* n);
/* gwas_eigenstratX.geno (popstrat.hpp:56-80) as bytes: geno[s*N + i] = sample i carries sampled
 * k-mer s, rows[s] = its global row index; ordered by row like the file the reference writes. */
int kmd_popstrat_geno(kmd_context* ctx, uint8_t* geno, uint64_t* rows, uint64_t cap, uint64_t* n);
/* smartpca's XTX (smartpca.c:907-1025) of the sampled k-mers: int8 tensor-core Gram matrices per
 * carrier-count bucket, combined in fp64.  xtx: N*N row-major.  normalize != 0 divides by
 * trace/(N-1) (smartpca.c:1019-1023); multi-GPU callers sum the un-normalised matrices of all ranks
 * first and call kmd_smartpca_normalize.  n_used = SNPs with at least one carrier. */
int kmd_popstrat_gram(kmd_context* ctx, double* xtx, int normalize, uint64_t* n_used);
/* Multi-GPU form (partitions, hence sampled k-mers, sharded over ranks): the INTEGER parts only — G_c, the co-carrier
 * counts of the SNPs with c carriers, written to d_G + (c - 1) * N * N (DEVICE memory of N^3 int32, caller-owned), and the
 * carrier histogram d_hist[0..N] (u64, DEVICE).  The ranks sum both (NCCL all-reduce: exact and order-free), then every
 * rank calls kmd_popstrat_gram_combine on the sums and obtains the XTX a single context would have computed from all the
 * sampled k-mers, bit for bit.  (smartpca.c:907-1025 accumulates the same sums SNP by SNP.) */
int kmd_popstrat_gram_partial(kmd_context* ctx, int32_t* d_G, uint64_t* d_hist);
int kmd_popstrat_gram_combine(kmd_context* ctx, const int32_t* d_G, const uint64_t* d_hist, double* xtx, int normalize,
                              uint64_t* n_used);
/* The same inside ONE process that drives several devices (the host program): partials on every context, summed on
 * ctxs[0]'s device over peer copies, combined there.  xtx: N*N row-major (host). */
int kmd_popstrat_gram_local(kmd_context* const* ctxs, uint32_t n_ctx, double* xtx, int normalize, uint64_t* n_used);
/* device time (ms) of the tensor-core kernel of the last kmd_popstrat_gram */
int kmd_popstrat_gram_ms(kmd_context* ctx, float* ms);
/* pop_strat_corrector: load_Z/load_Y + init_global_features (src/popstrat.cpp:153-171, 270-370),
 * including standardize() and the null-model IRLS.  Z: N x 10 row-major (what load_Z parses from
 * pcs.evec); totals: per-sample k-mer totals, controls first.  null_model: null_p = n_pc + 2
 * weights, may be NULL. */
int kmd_popstrat_init(kmd_context* ctx, const double* Z, const uint64_t* totals, double* null_model);
/* pop_strat_corrector::apply over all accumulators (popstrat.hpp:186-246): every prefilter
 * survivor's p-value is replaced by the GLM likelihood-ratio p-value (one warp per k-mer). */
int kmd_popstrat_apply(kmd_context* ctx);
/* pop_strat_corrector::apply(KmerSign&) (popstrat.hpp:249-333) on caller-supplied count vectors:
 * counts = n_kmers x N host counts of count_bytes each.  Outputs may be NULL: p (corrected
 * p-value), model (n_kmers x alt_p weights, alt_p = n_pc + 3), llr, iters. */
int kmd_popstrat_fit(kmd_context* ctx, const void* counts, uint64_t n_kmers, double* p, double* model,
                     double* llr, int32_t* iters);
/* counts of the prefilter survivors, row order like kmd_survivors: n x N values of count_bytes */
int kmd_survivor_counts(kmd_context* ctx, void* out, uint64_t cap, uint64_t* n);

/* host-side (no device needed) */
void kmd_smartpca_normalize(double* xtx, uint32_t n);
/* all eigenvalues (descending) and unit eigenvectors (evecs[j*n+i]; largest component positive) of
 * a symmetric matrix: smartpca's eigvecs (eigsubs.c:35-51).  Returns 0 on success. */
int kmd_sym_eig(const double* a, uint32_t n, double* evals, double* evecs);
/* the numeigs leading principal components as the reference reads them back: eigenvectors of the
 * normalised XTX, projected and rescaled to unit length (smartpca.c:1140-1329), printed "%10.4f"
 * (pcs.evec) and parsed (load_Z).  evals: n (may be NULL); pcs: n x numeigs row-major. */
int kmd_smartpca_pcs(const double* xtx, uint32_t n, uint32_t numeigs, double* evals, double* pcs);
/* Sampler decisions of rows [first_row, first_row+n) on the host (same arithmetic as the kernel) */
uint64_t kmd_sampler_rows(uint64_t seed, double kmer_pca, uint64_t first_row, uint64_t n, uint8_t* out);

/* ---- host-side pieces of the path that stay on the CPU -------------------------------------- */

/* LogFactorialTable (src/log_factorial_table.cpp:5-22): table[i] = log(i)+log(i-1)+...+log(2) */
void kmd_lf_table_fill(double* table, uint64_t size);
/* alglib::chisquarecdistribution(v, x) (thirdparty/alglib/src/specialfunctions.cpp:9559) */
double kmd_chisqc(double v, double x);
/* one corrector predicate (src/corrector.cpp): rank = 1-based rank for BENJAMINI, number of
 * earlier applications for HOLM, ignored otherwise */
int kmd_corrector_apply(int correction, double significance, uint64_t total_rows, uint64_t rank,
                        double p);
/* km::Kmer::to_string (kmtricks/kmer.hpp:38, 797-810); out holds k+1 bytes */
void kmd_kmer_to_string(const uint64_t* words, uint32_t kmer_words, uint32_t k, char* out);
/* one FASTA record exactly as aggregator.hpp:55-62 + kseq++ (60-column wrap) print it; returns
 * the number of bytes written (excluding the NUL), or the size needed if cap is too small */
size_t kmd_format_fasta(uint64_t index, const kmd_survivor* s, uint32_t kmer_words, uint32_t k,
                        char* out, size_t cap);

/* Bind the calling host thread to `device`: makes it the thread's current CUDA device and — where sysfs exposes the
 * GPU's NUMA node — restricts the thread to that node's CPUs and prefers its memory, so that staging buffers pinned
 * by this thread (kmd_host_alloc) and the copies into them stay next to the GPU's PCIe root.  A no-op on single-node
 * hosts.  Call before kmd_host_alloc; one process per GPU calls it once at start-up. */
int kmd_bind_host_thread(int device);
/* create the device's primary CUDA context now (hosts overlap this ~0.5 s with their own start-up work) */
int kmd_warmup(int device);
/* pinned host memory for kmd_scan_host / kmd_model_rows_host */
void* kmd_host_alloc(size_t bytes);
void kmd_host_free(void* p);

/* ---- seeded synthetic kmtricks-format rows (bench / tests) ------------------ */

typedef struct kmd_synth
{
  uint64_t seed;
  uint64_t total_rows;  /* rows of the whole synthetic matrix (fixes the k-mer spacing)   */
  uint32_t nb_controls;
  uint32_t nb_cases;
  uint32_t kmer_size;
  uint32_t kmer_words;
  uint32_t count_bytes;
  uint32_t enriched_ppm; /* rows per million whose cases (or controls) are over-abundant  */
} kmd_synth;

/* rows [row_base, row_base+n_rows) of the synthetic matrix, bit-identical on host and device */
int kmd_synth_fill_host(const kmd_synth* s, void* h_rows, uint64_t n_rows, uint64_t row_base,
                        int threads);
int kmd_synth_fill_device(const kmd_synth* s, int device, void* d_rows, uint64_t n_rows,
                          uint64_t row_base, void* cuda_stream);
/* per-sample totals the synthetic cohort declares (what the .hist files would hold) */
void kmd_synth_totals(const kmd_synth* s, uint64_t* totals);

#ifdef __cplusplus
}
#endif
#endif /* KMDIFF_B200_H */
