// Example model plugin for `kmdiff-b200 diff --cmodel libex_model.so` — the counterpart of the
// reference's plugins/ex_model.cpp (same exported symbols, same IModel interface); build with
//   g++ -std=c++17 -O2 -fPIC -shared -I include plugins/ex_model.cpp -o libex_model.so
// This one flags a k-mer when its mean abundance differs by more than `ratio` (from --config) between cohorts.
#include <cstdlib>

#include <kmdiff/imodel.hpp>

template <std::size_t MAX_C>
class FoldChangeModel : public kmdiff::IModel<MAX_C>
{
  using base = kmdiff::IModel<MAX_C>;
  using count_type = typename base::count_type;
  double m_ratio = 4.0;

 public:
  void configure(const std::string& config) override
  {
    if (!config.empty()) m_ratio = std::atof(config.c_str());
  }

  kmdiff::model_ret_t process(const kmdiff::Range<count_type>& controls, const kmdiff::Range<count_type>& cases) override
  {
    const double mc = base::mean(controls), mk = base::mean(cases);
    const bool up = mk > m_ratio * (mc + 1), down = mc > m_ratio * (mk + 1);
    const double p = (up || down) ? 1e-12 : 1.0;
    return std::make_tuple(p, up ? kmdiff::Significance::CASE : down ? kmdiff::Significance::CONTROL : kmdiff::Significance::NO, mc, mk);
  }
};

extern "C" std::string plugin_name() { return "FoldChangeModel"; }
extern "C" FoldChangeModel<kmdiff::maxc8>* create8() { return new FoldChangeModel<kmdiff::maxc8>(); }
extern "C" FoldChangeModel<kmdiff::maxc16>* create16() { return new FoldChangeModel<kmdiff::maxc16>(); }
extern "C" FoldChangeModel<kmdiff::maxc32>* create32() { return new FoldChangeModel<kmdiff::maxc32>(); }
