// The kmtricks on-disk formats `kmdiff diff` touches, and the run-directory helpers around them.
// Replaces (reference file:line; "km/" = thirdparty/kmtricks/include/kmtricks/):
//   get_kmtricks_config / get_total_kmer      src/kmtricks_utils.cpp:29-76, 78-139
//   km::Fof                                   km/io/fof.hpp:39-40 ("ID : path ; path ! min_abundance")
//   km::KmHeader + MatrixFileHeader           km/io/io_common.hpp:125-158, km/io/matrix_file.hpp:26-91
//   km::MatrixReader / MatrixWriter           km/io/matrix_file.hpp:93-154
//   km::HistFileHeader / HistReader           km/io/hist_file.hpp:25-144
//   km::KmerFileHeader / KmerReader           km/io/kmer_file.hpp:26-67, 112-137
//   km::KmDir::get_files_to_merge             km/kmdir.hpp:48-73
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lz4_frame.hpp"

namespace kmd_host {

namespace fs = std::filesystem;

struct KmdError : std::runtime_error
{
  std::string name;
  KmdError(const std::string& n, const std::string& msg) : std::runtime_error(msg), name(n) {}
};

constexpr uint64_t kKmMagic = 0x736b636972746d6bull;      // io_common.hpp:50-62
constexpr uint64_t kMatrixMagic = 0x6b5f78697274616dull;
constexpr uint64_t kHistMagic = 0x747369686bull;
constexpr size_t kMatrixHeaderBytes = 45;

#pragma pack(push, 1)
struct MatrixHeader
{
  uint64_t km_magic;
  uint32_t km_version;
  uint8_t compressed;
  uint64_t matrix_magic;
  uint32_t kmer_size;
  uint32_t kmer_slots;
  uint32_t count_slots;   // ignored by the reference readers (matrix_file.hpp:149-150)
  uint32_t nb_counts;
  uint32_t id;
  uint32_t partition;
};
#pragma pack(pop)
static_assert(sizeof(MatrixHeader) == kMatrixHeaderBytes, "matrix header is 45 packed bytes");

struct RunConfig
{
  uint32_t kmer_size = 0;
  uint32_t abundance_min = 0;
  uint32_t nb_partitions = 0;
};

inline std::string trim(const std::string& s)
{
  const size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  return s.substr(a, s.find_last_not_of(" \t\r\n") - a + 1);
}

// kmdiff-count.opt (kmer_size=, abundance_min=) when present, else kmtricks' options.txt (kmer_size=,
// c_ab_min=): one comma-separated line of key=value pairs.  nb_partitions = entries of <run>/counts.
inline RunConfig get_kmtricks_config(const std::string& run_dir)
{
  std::string path = run_dir + "/kmdiff-count.opt", ab_key = "abundance_min";
  if (!fs::exists(path) || fs::file_size(path) == 0) { path = run_dir + "/options.txt"; ab_key = "c_ab_min"; }
  std::ifstream in(path);
  RunConfig c;
  std::string line;
  while (std::getline(in, line))
  {
    if (line.find("kmer_size") == std::string::npos) continue;
    std::stringstream ss(line);
    std::string item;
    while (std::getline(ss, item, ','))
    {
      const size_t eq = item.find('=');
      if (eq == std::string::npos) continue;
      const std::string key = trim(item.substr(0, eq)), val = trim(item.substr(eq + 1));
      if (key.find("kmer_size") != std::string::npos) c.kmer_size = static_cast<uint32_t>(std::stoul(val));
      else if (key.find(ab_key) != std::string::npos) c.abundance_min = static_cast<uint32_t>(std::stoul(val));
    }
  }
  const std::string counts = run_dir + "/counts";
  if (fs::is_directory(counts))
    for (auto& e : fs::directory_iterator(counts)) { (void)e; c.nb_partitions++; }
  if (c.kmer_size == 0 || c.nb_partitions == 0)
    throw KmdError("ConfigError", "Unable to load config from " + path + ".");
  return c;
}

struct FofEntry { std::string id; uint32_t min_abundance; };

inline std::vector<FofEntry> parse_fof(const std::string& path)
{
  std::ifstream in(path);
  if (!in) throw KmdError("IOError", "Unable to open " + path);
  std::vector<FofEntry> out;
  std::string line;
  while (std::getline(in, line))
  {
    line = trim(line);
    if (line.empty()) continue;
    const size_t colon = line.find(':');
    if (colon == std::string::npos) throw KmdError("IOError", "Invalid fof line: " + line);
    FofEntry e{trim(line.substr(0, colon)), 0};
    const size_t bang = line.find('!', colon);
    if (bang != std::string::npos)
    {
      const std::string v = trim(line.substr(bang + 1));
      if (!v.empty()) e.min_abundance = static_cast<uint32_t>(std::stoul(v));
    }
    out.push_back(e);
  }
  return out;
}

// Payload bytes of a kmtricks file after its header: raw, or one LZ4 frame.
class PayloadReader
{
 public:
  PayloadReader(FILE* f, bool compressed) : m_f(f) { if (compressed) m_lz4 = std::make_unique<Lz4FrameReader>(f); }
  size_t read(void* dst, size_t n)
  {
    if (m_lz4) return m_lz4->read(static_cast<uint8_t*>(dst), n);
    return fread(dst, 1, n, m_f);
  }
 private:
  FILE* m_f;
  std::unique_ptr<Lz4FrameReader> m_lz4;
};

// per-sample total = hist.total - sum_{j=1}^{ab_min-1} j * hist_u[j-1]   (kmtricks_utils.cpp:93-107)
inline uint64_t hist_total(const std::string& path, uint32_t ab_min)
{
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw KmdError("IOError", "Unable to open " + path);
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
#pragma pack(push, 1)
  struct { uint64_t km_magic; uint32_t version; uint8_t compressed; } base;
  struct { uint64_t magic; uint32_t k, id; uint64_t lower, upper, uniq, total, oob_ln, oob_lu, oob_un, oob_uu; } h;
#pragma pack(pop)
  if (fread(&base, sizeof base, 1, f) != 1 || fread(&h, sizeof h, 1, f) != 1 || base.km_magic != kKmMagic || h.magic != kHistMagic)
    throw KmdError("IOError", "Invalid file format: " + path);
  uint64_t total = h.total;
  if (ab_min > 1)
  {
    const uint64_t n = h.upper - h.lower + 1;
    std::vector<uint64_t> hu(n);
    PayloadReader pr(f, base.compressed);
    if (pr.read(hu.data(), n * 8) != n * 8) throw KmdError("IOError", "Truncated histogram: " + path);
    for (uint64_t j = 1; j < ab_min && j - 1 < n; j++) total -= j * hu[j - 1];
  }
  return total;
}

inline void get_total_kmer(const std::string& run_dir, size_t nb_controls, size_t nb_cases, uint32_t ab_min_default,
                           std::vector<uint64_t>& controls, std::vector<uint64_t>& cases)
{
  const auto fof = parse_fof(run_dir + "/kmtricks.fof");
  if (fof.size() < nb_controls + nb_cases)
    throw KmdError("ConfigError", "kmtricks.fof lists fewer samples than --nb-controls + --nb-cases.");
  controls.clear(); cases.clear();
  for (size_t i = 0; i < nb_controls + nb_cases; i++)
  {
    const uint32_t ab = fof[i].min_abundance ? fof[i].min_abundance : ab_min_default;
    const uint64_t t = hist_total(run_dir + "/histograms/" + fof[i].id + ".hist", ab);
    (i < nb_controls ? controls : cases).push_back(t);
  }
}

// <run>/matrices/* in directory order sorted by name (the reference iterates fs::directory_iterator order,
// cmd/diff.hpp:80-86; partition order does not change the result set)
inline std::vector<std::string> list_matrices(const std::string& run_dir)
{
  std::vector<std::string> out;
  const std::string d = run_dir + "/matrices";
  if (fs::is_directory(d))
    for (auto& e : fs::directory_iterator(d)) out.push_back(e.path().string());
  std::sort(out.begin(), out.end());
  return out;
}

struct MatrixFile
{
  FILE* f = nullptr;
  MatrixHeader hdr{};
  std::unique_ptr<PayloadReader> payload;
  uint64_t rows_if_raw = 0;   // (file size - 45) / R for an uncompressed matrix, else 0

  MatrixFile(const std::string& path, size_t row_bytes)
  {
    f = fopen(path.c_str(), "rb");
    if (!f) throw KmdError("IOError", "Unable to open " + path);
    if (fread(&hdr, sizeof hdr, 1, f) != 1 || hdr.km_magic != kKmMagic || hdr.matrix_magic != kMatrixMagic)
    {
      fclose(f); f = nullptr;
      throw KmdError("IOError", "Invalid file format: " + path);
    }
    payload = std::make_unique<PayloadReader>(f, hdr.compressed != 0);
    if (!hdr.compressed) rows_if_raw = (fs::file_size(path) - kMatrixHeaderBytes) / row_bytes;
  }
  ~MatrixFile() { if (f) fclose(f); }
  MatrixFile(const MatrixFile&) = delete;
  MatrixFile& operator=(const MatrixFile&) = delete;
};

// counts/partition_{p}/<sample>.kmer[.lz4]: 41-byte header, then records of kmer_slots*8 + count_slots bytes
constexpr uint64_t kKmerMagic = 0x72656d6bull;
constexpr size_t kKmerHeaderBytes = 41;

#pragma pack(push, 1)
struct KmerHeader
{
  uint64_t km_magic;
  uint32_t km_version;
  uint8_t compressed;
  uint64_t kmer_magic;
  uint32_t kmer_size;
  uint32_t kmer_slots;
  uint32_t count_slots;   // width of the count on disk (KmerReader::read, kmer_file.hpp:132)
  uint32_t id;
  uint32_t partition;
};
#pragma pack(pop)
static_assert(sizeof(KmerHeader) == kKmerHeaderBytes, "kmer file header is 41 packed bytes");

struct KmerFile
{
  FILE* f = nullptr;
  KmerHeader hdr{};
  std::unique_ptr<PayloadReader> payload;

  explicit KmerFile(const std::string& path)
  {
    f = fopen(path.c_str(), "rb");
    if (!f) throw KmdError("IOError", "Unable to open " + path);
    if (fread(&hdr, sizeof hdr, 1, f) != 1 || hdr.km_magic != kKmMagic || hdr.kmer_magic != kKmerMagic)
    {
      fclose(f); f = nullptr;
      throw KmdError("IOError", "Invalid file format: " + path);
    }
    payload = std::make_unique<PayloadReader>(f, hdr.compressed != 0);
  }
  ~KmerFile() { if (f) fclose(f); }
  KmerFile(const KmerFile&) = delete;
  KmerFile& operator=(const KmerFile&) = delete;
};

// KmDir::get_files_to_merge(p, true, KMER): one file per fof line, in fof order, named `.kmer.lz4` whatever
// the header's compressed flag says; like the reference, only the first path is checked for existence.
inline std::vector<std::string> get_files_to_merge(const std::string& run_dir, uint32_t partition)
{
  std::vector<std::string> paths;
  for (const FofEntry& e : parse_fof(run_dir + "/kmtricks.fof"))
  {
    paths.push_back(run_dir + "/counts/partition_" + std::to_string(partition) + "/" + e.id + ".kmer.lz4");
    if (!fs::exists(paths.front())) throw KmdError("FileNotFoundError", paths.front() + " is missing.");
  }
  return paths;
}

// --save-sk: km::MatrixWriter<65536>(path, k, count_size = 4, nb_counts, id = 0, partition, lz4 = true)
// (include/kmdiff/merge.hpp:270-278, km/io/matrix_file.hpp:93-128): the 45-byte header, then one LZ4 frame of rows
class MatrixWriter
{
 public:
  MatrixWriter(const std::string& path, uint32_t k, uint32_t kmer_slots, uint32_t count_slots, uint32_t nb_counts, uint32_t partition)
  {
    m_f = fopen(path.c_str(), "wb");
    if (!m_f) throw KmdError("IOError", "Unable to open " + path);
    MatrixHeader h{kKmMagic, 0, 1, kMatrixMagic, k, kmer_slots, count_slots, nb_counts, 0, partition};
    fwrite(&h, sizeof h, 1, m_f);
    m_lz4 = std::make_unique<Lz4FrameWriter>(m_f);
  }
  void write(const void* row, size_t n) { m_lz4->write(row, n); }
  ~MatrixWriter() { if (m_lz4) m_lz4->close(); if (m_f) fclose(m_f); }
 private:
  FILE* m_f;
  std::unique_ptr<Lz4FrameWriter> m_lz4;
};

}  // namespace kmd_host
