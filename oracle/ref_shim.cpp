// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-linkage shim around the UNMODIFIED reference implementation (tlemane/kmdiff).
// It is compiled by oracle/Makefile against the reference sources where they lie
// under $(REF) (= /root/reference); no reference source is copied into this repo.
// The resulting oracle/_ref/libkmdiff_ref.so is used ONLY by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs,
// to (a) pin oracle/kmd_oracle.c against the real thing and (b) time the
// reference's own CPU model loop.
//
// What it binds (reference file:line):
//   kmdiff::LogFactorialTable                 include/kmdiff/log_factorial_table.hpp:9-26
//   kmdiff::PoissonLikelihood<MAX_C>::process include/kmdiff/model.hpp:142-176
//   alglib::chisquarecdistribution            thirdparty/alglib/src/specialfunctions.cpp:2750
//   kmdiff::make_corrector / ICorrector       src/corrector.cpp:101-117
//   the per-row loop of diff_observer::process include/kmdiff/merge.hpp:68-103
//   kmdiff::Sampler<MAX_C>::sample()          include/kmdiff/popstrat.hpp:107-146
//   kmdiff::glm_irls / inverse / sigmoid      src/linear_model.cpp:297-410, 139-189, 191-195
//   kmdiff::pop_strat_corrector               include/kmdiff/popstrat.hpp:148-372, src/popstrat.cpp
//   km::KmerMerger / km::KmerReader           thirdparty/kmtricks/include/kmtricks/merge.hpp:102-361, io/kmer_file.hpp
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <fstream>
#include <map>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <tuple>

#include <kmdiff/model.hpp>
#include <kmdiff/corrector.hpp>
#include <kmdiff/log_factorial_table.hpp>
#include <kmdiff/linear_model.hpp>
#include <specialfunctions.h>
#include <kmtricks/merge.hpp>

// pop_strat_corrector::apply(KmerSign&) and its feature matrices are private; the shim reads them
// to pin the oracle, so the class is opened up for this translation unit only (every standard
// header it uses is already included above).
#define private public
#include <kmdiff/popstrat.hpp>
#undef private

namespace {

template <std::size_t MAX_C>
struct RefModel
{
  using count_t = typename km::selectC<MAX_C>::type;
  std::vector<std::size_t> tc, tk;
  std::size_t nc, nk;
  std::unique_ptr<kmdiff::PoissonLikelihood<MAX_C>> model;

  RefModel(std::size_t nc_, std::size_t nk_, const std::uint64_t* totals, std::size_t preload)
    : tc(totals, totals + nc_), tk(totals + nc_, totals + nc_ + nk_), nc(nc_), nk(nk_)
  {
    model = std::make_unique<kmdiff::PoissonLikelihood<MAX_C>>(nc, nk, tc, tk, preload);
  }

  // One call of IModel::process on one row, exactly as diff_observer does it
  // (merge.hpp:70-73): two Range views over one count vector.
  void row(std::vector<count_t>& v, double* p, std::int32_t* sign, double* mc, double* kk)
  {
    kmdiff::Range<count_t> rc(v, 0, nc);
    kmdiff::Range<count_t> rk(v, nc, nk);
    auto [pv, s, a, b] = model->process(rc, rk);
    *p = pv; *sign = static_cast<std::int32_t>(s); *mc = a; *kk = b;
  }

  // rows: AoS matrix payload (kmtricks matrix_file.hpp:144-154): kmer_bytes of k-mer then
  // (nc+nk) counts of sizeof(count_t) per row.
  void rows(const std::uint8_t* buf, std::size_t n_rows, std::size_t kmer_bytes,
            double* p, std::int32_t* sign, double* mc, double* kk, int threads)
  {
    const std::size_t n = nc + nk;
    const std::size_t R = kmer_bytes + n * sizeof(count_t);
    auto work = [&](std::size_t lo, std::size_t hi) {
      std::vector<count_t> v(n);
      for (std::size_t r = lo; r < hi; r++)
      {
        std::memcpy(v.data(), buf + r * R + kmer_bytes, n * sizeof(count_t));
        row(v, p + r, sign + r, mc + r, kk + r);
      }
    };
    if (threads <= 1) { work(0, n_rows); return; }
    std::vector<std::thread> pool;
    std::size_t chunk = (n_rows + threads - 1) / threads;
    for (int t = 0; t < threads; t++)
    {
      std::size_t lo = std::min(n_rows, t * chunk), hi = std::min(n_rows, lo + chunk);
      pool.emplace_back(work, lo, hi);
    }
    for (auto& t : pool) t.join();
  }
};

} // namespace

// km::KmerMerger over per-sample .kmer[.lz4] files exactly as global_merge::merge() builds it
// (include/kmdiff/merge.hpp:266-267: recurrence_min 1, save_if 0) with a row-collecting observer in place
// of diff_observer.  Rows: kmer_slots*8 bytes + n_paths counts of count_bytes.  Returns kept rows
// (rows beyond cap_rows are counted, not stored).
namespace {
template <std::size_t K, std::size_t MAX_C>
struct RowCollector : km::IMergeObserver<K, MAX_C>
{
  using count_t = typename km::selectC<MAX_C>::type;
  std::uint8_t* out; std::size_t cap, n = 0, kw;
  RowCollector(std::uint8_t* o, std::size_t c, std::size_t w) : out(o), cap(c), kw(w) {}
  void process(km::Kmer<K>& kmer, std::vector<count_t>& counts) override
  {
    const std::size_t R = 8 * kw + counts.size() * sizeof(count_t);
    if (out && n < cap)
    {
      std::memcpy(out + n * R, kmer.get_data64(), 8 * kw);
      std::memcpy(out + n * R + 8 * kw, counts.data(), counts.size() * sizeof(count_t));
    }
    n++;
  }
};

template <std::size_t K, std::size_t MAX_C>
std::size_t kmer_merge(const char* const* paths, std::size_t n_paths, std::uint32_t kmer_size, const std::uint32_t* ab_min,
                       std::uint8_t* rows_out, std::size_t cap_rows)
{
  std::vector<std::string> p(paths, paths + n_paths);
  std::vector<std::uint32_t> a(ab_min, ab_min + n_paths);
  km::KmerMerger<K, MAX_C> merger(p, a, kmer_size, 1, 0);
  auto obs = std::make_shared<RowCollector<K, MAX_C>>(rows_out, cap_rows, (kmer_size + 31) / 32);
  merger.merge(obs);
  return obs->n;
}
}


extern "C" {

void ref_lf_table(std::size_t size, std::size_t n_query, const std::uint64_t* idx, double* out)
{
  kmdiff::LogFactorialTable t(size);
  for (std::size_t i = 0; i < n_query; i++) out[i] = t[idx[i]];
}

double ref_chisqc(double v, double x) { return alglib::chisquarecdistribution(v, x); }

void* ref_model_create(int count_bytes, std::size_t nc, std::size_t nk,
                       const std::uint64_t* totals, std::size_t preload)
{
  switch (count_bytes)
  {
    case 1: return new RefModel<255>(nc, nk, totals, preload);
    case 2: return new RefModel<65535>(nc, nk, totals, preload);
    case 4: return new RefModel<4294967295>(nc, nk, totals, preload);
  }
  return nullptr;
}

void ref_model_destroy(int count_bytes, void* h)
{
  switch (count_bytes)
  {
    case 1: delete static_cast<RefModel<255>*>(h); break;
    case 2: delete static_cast<RefModel<65535>*>(h); break;
    case 4: delete static_cast<RefModel<4294967295>*>(h); break;
  }
}

void ref_model_rows(int count_bytes, void* h, const std::uint8_t* buf, std::size_t n_rows,
                    std::size_t kmer_bytes, double* p, std::int32_t* sign, double* mc, double* kk,
                    int threads)
{
  switch (count_bytes)
  {
    case 1: static_cast<RefModel<255>*>(h)->rows(buf, n_rows, kmer_bytes, p, sign, mc, kk, threads); break;
    case 2: static_cast<RefModel<65535>*>(h)->rows(buf, n_rows, kmer_bytes, p, sign, mc, kk, threads); break;
    case 4: static_cast<RefModel<4294967295>*>(h)->rows(buf, n_rows, kmer_bytes, p, sign, mc, kk, threads); break;
  }
}

// type: kmdiff::CorrectionType (0 NOTHING, 1 BONFERRONI, 2 BENJAMINI, 3 SIDAK, 4 HOLM)
void* ref_corrector_create(int type, double threshold, std::size_t total)
{
  auto* sp = new std::shared_ptr<kmdiff::ICorrector>(
    kmdiff::make_corrector(static_cast<kmdiff::CorrectionType>(type), threshold, total));
  return sp;
}

int ref_corrector_apply(void* h, double p)
{
  return (*static_cast<std::shared_ptr<kmdiff::ICorrector>*>(h))->apply(p) ? 1 : 0;
}

void ref_corrector_destroy(void* h)
{
  delete static_cast<std::shared_ptr<kmdiff::ICorrector>*>(h);
}


// ---- population stratification ------------------------------------------------------------------

// Sampler::sample() n times: out[i] = (draw i < v).  geno/snp sinks are not touched by sample().
std::size_t ref_sampler_rows(std::size_t seed, double v, std::size_t n, std::uint8_t* out)
{
  kmdiff::Sampler<4294967295> s(nullptr, nullptr, v, seed);
  std::size_t taken = 0;
  for (std::size_t i = 0; i < n; i++) { out[i] = s.sample() ? 1 : 0; taken += out[i]; }
  return taken;
}

double ref_sigmoid(double x) { return kmdiff::sigmoid(x); }

void ref_inverse(const double* a, int n, double* inv, int* singular, int* is_nan)
{
  kmdiff::matrix_t m(n, kmdiff::vector_t(n));
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) m[i][j] = a[i * n + j];
  auto [r, s, q] = kmdiff::inverse(m, n);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = r[i][j];
  *singular = s; *is_nan = q;
}

void ref_glm_irls(const double* x, const double* y, int n, int p, int max_iters, double* weight,
                  int* singular, int* is_nan, double* err, int* iters)
{
  kmdiff::matrix_t X(n, kmdiff::vector_t(p));
  kmdiff::vector_t Y(y, y + n);
  for (int i = 0; i < n; i++) for (int j = 0; j < p; j++) X[i][j] = x[i * p + j];
  auto [w, s, q, e, it] = kmdiff::glm_irls(X, Y, max_iters);
  for (int j = 0; j < p; j++) weight[j] = w[j];
  *singular = s; *is_nan = q; *err = e; *iters = it;
}

// pop_strat_corrector built the way do_pop builds it (cmd/diff.hpp:203-211): load_Z / load_Y read
// the files the caller wrote (pcs.evec: 10 numbers per line; .ind: "id\tU\tControl|Case").
void* ref_popstrat_create(std::size_t nc, std::size_t nk, const std::uint64_t* totals, std::size_t npc,
                          const char* pcs_evec, const char* ind_file)
{
  std::vector<std::size_t> tc(totals, totals + nc), tk(totals + nc, totals + nc + nk);
  auto* c = new kmdiff::pop_strat_corrector(nc, nk, tc, tk, npc);
  c->load_Z(pcs_evec);
  c->load_Y(ind_file);
  c->load_C("");
  c->load_ginfo("");
  c->init_global_features();
  return c;
}

void ref_popstrat_destroy(void* h) { delete static_cast<kmdiff::pop_strat_corrector*>(h); }

// null / alt global feature matrices after standardize() and the null model
void ref_popstrat_state(void* h, double* null_features, double* alt_features, double* null_model)
{
  auto* c = static_cast<kmdiff::pop_strat_corrector*>(h);
  const std::size_t n = c->m_size, pc = c->m_null_feature_count, pa = c->m_alt_feature_count;
  for (std::size_t i = 0; i < n; i++)
  {
    for (std::size_t j = 0; j < pc; j++) null_features[i * pc + j] = c->m_null_global_features[i][j];
    for (std::size_t j = 0; j < pa; j++) alt_features[i * pa + j] = c->m_alt_global_features[i][j];
  }
  for (std::size_t j = 0; j < pc; j++) null_model[j] = c->m_null_model[j];
}

// pop_strat_corrector::apply(KmerSign&) on one survivor's counts -> corrected p-value
double ref_popstrat_apply(void* h, const double* counts, std::size_t n)
{
  auto* c = static_cast<kmdiff::pop_strat_corrector*>(h);
  std::vector<double> v(counts, counts + n);
  km::Kmer<32> k;
  kmdiff::KmerSign<32> ks(std::move(k), 0.5, kmdiff::Significance::CASE, v, 0.0, 0.0);
  c->apply(ks);
  return ks.m_pvalue;
}

std::size_t ref_kmer_merge(const char* const* paths, std::size_t n_paths, std::uint32_t kmer_size, int count_bytes,
                           const std::uint32_t* ab_min, std::uint8_t* rows_out, std::size_t cap_rows)
{
  // main.cc:74-75 dispatches on k < 32 / < 64; count width is the compile-time MAX_C of the build
  if (count_bytes == 4)
    return kmer_size < 32 ? kmer_merge<32, 4294967295>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows)
                          : kmer_merge<64, 4294967295>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows);
  if (count_bytes == 2)
    return kmer_size < 32 ? kmer_merge<32, 65535>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows)
                          : kmer_merge<64, 65535>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows);
  return kmer_size < 32 ? kmer_merge<32, 255>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows)
                        : kmer_merge<64, 255>(paths, n_paths, kmer_size, ab_min, rows_out, cap_rows);
}

} // extern "C"
