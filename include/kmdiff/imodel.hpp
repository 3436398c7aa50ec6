// kmdiff_b200 — plugin API, part 2: the model interface of `kmdiff diff --cmodel lib.so`.
//
// Source- and ABI-compatible with the reference's include/kmdiff/imodel.hpp:13-72 (+ Significance,
// include/kmdiff/kmer.hpp:33-38, and km::selectC, kmtricks/utils.hpp:312-327): same namespace, names,
// template parameter, virtual-function order (destructor, configure, process) and return tuple.  A plugin
// written for the reference (plugins/ex_model.cpp there) builds against this header unchanged:
//
//   extern "C" std::string plugin_name();
//   extern "C" kmdiff::IModel<kmdiff::maxc8>*  create8();     // 1-byte counts
//   extern "C" kmdiff::IModel<kmdiff::maxc16>* create16();    // 2-byte counts
//   extern "C" kmdiff::IModel<kmdiff::maxc32>* create32();    // 4-byte counts
//
// The built-in Poisson likelihood-ratio model runs on the GPU (csrc/); a user model necessarily runs
// per row on the host, behind the same matrix reader and the same correction / output stages.  One
// instance is shared by all reader threads: process() must be re-entrant (as in the reference,
// merge.hpp:281-283).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <numeric>
#include <string>
#include <tuple>

#include <kmdiff/range.hpp>

#ifndef KMDIFF_B200_NO_KM_SELECTC
namespace km {
// count type for a maximum count value: <= 255 -> u8, <= 65535 -> u16, else u32
template <std::size_t C> struct requiredC { enum { value = C <= 0xFF ? 8 : C <= 0xFFFF ? 16 : 32 }; };
template <int bits> struct select_;
template <> struct select_<8> { typedef std::uint8_t type; };
template <> struct select_<16> { typedef std::uint16_t type; };
template <> struct select_<32> { typedef std::uint32_t type; };
template <std::size_t C> struct selectC : select_<requiredC<C>::value> {};
}  // namespace km
#endif

namespace kmdiff {

enum class Significance { CONTROL, CASE, NO };

using pvalue_t = double;
// (p-value, enrichment sign, "mean" of the controls, "mean" of the cases) — what ends up in the FASTA header
using model_ret_t = std::tuple<pvalue_t, Significance, double, double>;

constexpr std::size_t maxc8 = std::numeric_limits<std::uint8_t>::max();
constexpr std::size_t maxc16 = std::numeric_limits<std::uint16_t>::max();
constexpr std::size_t maxc32 = std::numeric_limits<std::uint32_t>::max();

template <std::size_t MAX_C>
class IModel
{
 public:
  using count_type = typename km::selectC<MAX_C>::type;
  using range_type = Range<count_type>;

  IModel() = default;
  virtual ~IModel() {}

  // --config string of the command line
  virtual void configure(const std::string& config) = 0;
  // one k-mer row: the counts of the controls and of the cases
  virtual model_ret_t process(const range_type& controls, const range_type& cases) = 0;

  // helpers the reference offers to plugin authors
  template <typename Iterable>
  static double mean(const Iterable& v)
  {
    return std::accumulate(v.begin(), v.end(), 0.0) / v.size();
  }

  template <typename Iterable>
  static std::tuple<double, std::size_t> sum_count(const Iterable& v)
  {
    double total = 0;
    std::size_t positive = 0;
    for (auto& e : v) { total += e; positive += e > 0 ? 1 : 0; }
    return std::make_tuple(total, positive);
  }

  template <typename Iterable>
  static std::tuple<double, std::size_t> mean_count(const Iterable& v)
  {
    auto [total, positive] = sum_count(v);
    return std::make_tuple(total / static_cast<double>(v.size()), positive);
  }

  template <typename Iterable>
  static double sd(const Iterable& v)
  {
    const double m = mean(v);
    const double sq = std::inner_product(v.begin(), v.end(), v.begin(), 0.0);
    return std::sqrt(sq / v.size() - m * m);
  }
};

}  // namespace kmdiff
