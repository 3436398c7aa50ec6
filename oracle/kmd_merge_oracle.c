/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See kmd_merge_oracle.h. */
#include "kmd_merge_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct
{
  uint64_t w[2]; /* w[0] least significant (kmtricks kmer.hpp: get_data64) */
} kmer_t;

static int kmer_lt(const kmer_t* a, const kmer_t* b) { return a->w[1] != b->w[1] ? a->w[1] < b->w[1] : a->w[0] < b->w[0]; }
static int kmer_eq(const kmer_t* a, const kmer_t* b) { return a->w[0] == b->w[0] && a->w[1] == b->w[1]; }

/* KmerMerger::element (merge.hpp:107-112) */
typedef struct
{
  kmer_t value;
  uint32_t count;
  int is_set;
  uint64_t pos; /* next record of the stream */
} element;

/* KmerReader::read (kmer_file.hpp:127-137): kmer_slots*8 bytes then count_slots bytes; false at EOF */
static int read_next(element* e, const uint8_t* stream, uint64_t n, int kw, int cb)
{
  if (e->pos >= n) return 0;
  const uint8_t* r = stream + e->pos * (size_t)(8 * kw + cb);
  e->value.w[0] = e->value.w[1] = 0;
  memcpy(e->value.w, r, (size_t)8 * kw);
  e->count = 0;
  memcpy(&e->count, r + 8 * kw, (size_t)cb);
  e->pos++;
  return 1;
}

uint64_t kmo_kmer_merge(const uint8_t* const* streams, const uint64_t* n_records, uint32_t n_streams,
                        int kmer_words, int count_bytes, const uint32_t* ab_min, uint32_t recurrence_min,
                        uint32_t save_if, uint8_t* rows_out, uint64_t cap_rows)
{
  const size_t R = (size_t)8 * kmer_words + (size_t)n_streams * count_bytes;
  element* el = (element*)calloc(n_streams ? n_streams : 1, sizeof(element));
  uint32_t* counts = (uint32_t*)calloc(n_streams ? n_streams : 1, sizeof(uint32_t));
  uint32_t* need_check = (uint32_t*)calloc(n_streams ? n_streams : 1, sizeof(uint32_t));
  kmer_t current = {{0, 0}}, next = {{0, 0}};
  int current_set = 0;
  uint64_t kept = 0;

  /* init_state (merge.hpp:148-166) */
  for (uint32_t i = 0; i < n_streams; i++)
  {
    if (read_next(&el[i], streams[i], n_records[i], kmer_words, count_bytes)) el[i].is_set = 1;
    if ((!current_set || kmer_lt(&el[i].value, &current)) && el[i].is_set)
    {
      current = next = el[i].value;
      current_set = 1;
    }
  }

  for (;;)
  {
    /* next() (merge.hpp:183-260) */
    int keep = 0, finish = 1, next_set = 0;
    uint32_t recurrence = 0, solid_in = 0, n_check = 0;
    current = next;
    for (uint32_t i = 0; i < n_streams; i++)
    {
      if (el[i].is_set && kmer_eq(&el[i].value, &current))
      {
        finish = 0;
        counts[i] = el[i].count;
        if (counts[i] >= ab_min[i]) { recurrence++; solid_in++; }
        else if (save_if) need_check[n_check++] = i;
        else counts[i] = 0;
        if (!read_next(&el[i], streams[i], n_records[i], kmer_words, count_bytes)) el[i].is_set = 0;
      }
      else counts[i] = 0;
      if (el[i].is_set && (!next_set || kmer_lt(&el[i].value, &next)))
      {
        next = el[i].value;
        next_set = 1;
      }
    }
    for (uint32_t f = 0; f < n_check; f++)
      if (!(solid_in >= save_if)) counts[need_check[f]] = 0;
    if (recurrence >= recurrence_min) keep = 1;
    if (finish) break;
    /* merge(obs) (merge.hpp:318-323): the observer sees kept rows only */
    if (keep)
    {
      if (rows_out && kept < cap_rows)
      {
        uint8_t* o = rows_out + kept * R;
        memcpy(o, current.w, (size_t)8 * kmer_words);
        for (uint32_t i = 0; i < n_streams; i++) memcpy(o + 8 * kmer_words + (size_t)i * count_bytes, &counts[i], (size_t)count_bytes);
      }
      kept++;
    }
  }
  free(el); free(counts); free(need_check);
  return kept;
}
