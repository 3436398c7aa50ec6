/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of km::KmerMerger (reference thirdparty/kmtricks/include/kmtricks/merge.hpp:102-361)
 * as kmdiff drives it for run directories without matrices (include/kmdiff/merge.hpp:237-309,
 * cmd/diff.hpp:88-95), and of the per-sample record files it reads
 * (thirdparty/kmtricks/include/kmtricks/io/kmer_file.hpp:26-67, 127-137).
 *
 * Parity status: PINNED.  tests/test_merge_oracle.py checks it against (1) the reference's own
 * fixture tests/data_test/km_out_dir and the expectations of tests/merge_test.cpp (320 merged rows,
 * per-sample totals 160/160, no significant k-mer; fixture files committed as hex in tests/golden/km_out_dir.json
 * by tests/golden/make_merge_golden.py), (2) the real km::KmerMerger behind oracle/_ref
 * (ref_kmer_merge in ref_shim.cpp) on seeded streams, and (3) vectors generated from it
 * (tests/golden/merge_vectors.json).
 */
#ifndef KMD_MERGE_ORACLE_H
#define KMD_MERGE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* KmerMerger::next + keep() (merge.hpp:183-260) over in-memory streams until every stream is exhausted.
 * streams[s]: n_records[s] records of kmer_words*8 + count_bytes bytes, ascending k-mers.
 * ab_min[s]: abundance_min of sample s; recurrence_min, save_if as in the constructor (merge.hpp:115-126).
 * Kept rows (kmer words, then n_streams counts of count_bytes) are appended to rows_out (capacity
 * cap_rows rows; NULL = count only).  Returns the number of kept rows. */
uint64_t kmo_kmer_merge(const uint8_t* const* streams, const uint64_t* n_records, uint32_t n_streams,
                        int kmer_words, int count_bytes, const uint32_t* ab_min, uint32_t recurrence_min,
                        uint32_t save_if, uint8_t* rows_out, uint64_t cap_rows);

#ifdef __cplusplus
}
#endif
#endif
