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
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <kmdiff/model.hpp>
#include <kmdiff/corrector.hpp>
#include <kmdiff/log_factorial_table.hpp>
#include <specialfunctions.h>

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

} // extern "C"
