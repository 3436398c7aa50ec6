// kmdiff_b200 — plugin API, part 1: the count-row view handed to IModel::process.
//
// ABI-compatible with the reference's include/kmdiff/range.hpp:7-35 (a reference to the row's
// std::vector, a start offset and a length, in that order), so a model plugin compiled against
// either header can be loaded by either host.
#pragma once

#include <cstddef>
#include <vector>

namespace kmdiff {

template <typename T>
class Range
{
  std::vector<T>& m_data;   // the whole row: controls first, then cases
  std::size_t m_start;
  std::size_t m_size;

 public:
  Range(std::vector<T>& row, std::size_t first, std::size_t count) : m_data(row), m_start(first), m_size(count) {}

  auto begin() const { return m_data.begin() + static_cast<std::ptrdiff_t>(m_start); }
  auto end() const { return m_data.begin() + static_cast<std::ptrdiff_t>(m_start + m_size); }
  std::size_t size() const { return m_size; }
  const T& operator[](std::size_t i) const { return m_data[m_start + i]; }
};

}  // namespace kmdiff
