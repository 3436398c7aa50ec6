// kmdiff-b200 — the `kmdiff diff` command line over the B200 hot path.
//
// Host-side counterpart of (reference file:line):
//   main / CLI                      src/main.cc:52-103, src/cli.cpp:148-362
//   main_diff / do_diff / do_correction   include/kmdiff/cmd/diff.hpp:66-164, 227-260, 262-377
//   global_merge::merge(paths) + matrix_proxy   include/kmdiff/merge.hpp:194-203, 341-397
//   plugin_manager                  include/kmdiff/model_manager.hpp:33-94
//   FileAccumulator / KmerSign::dump   include/kmdiff/accumulator.hpp:156-285, kmer.hpp:113-127
//   aggregator::writer              include/kmdiff/aggregator.hpp:27-71, 197-199
//
// What runs where: this program only frames bytes (run-dir parsing, matrix files -> pinned buffers,
// spill files, FASTA text).  Every row is tested on the GPU through the C-ABI (include/kmdiff_b200.h);
// there is no CPU implementation of the built-in model here.  Only a user `--cmodel` plugin — by
// contract a host C++ object called once per row — is evaluated on the host, and its survivors are
// pushed into the same device-side accumulators for correction and output.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <unistd.h>

#include <kmdiff/imodel.hpp>
#include <kmdiff_b200.h>

#include "km_formats.hpp"

using namespace kmd_host;

// ---- logging: spdlog pattern "[%Y-%m-%d %H:%M:%S.%e] [%^%l%$] %v" on stderr (main.cc:62-64) ----------

namespace {

enum Level { L_DEBUG = 0, L_INFO = 1, L_WARN = 2, L_ERROR = 3 };
Level g_level = L_INFO;
std::mutex g_log_mutex;

void logf(Level lv, const char* fmt, ...)
{
  if (lv < g_level) return;
  static const char* names[] = {"debug", "info", "warning", "error"};
  char msg[2048];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
  timeval tv;
  gettimeofday(&tv, nullptr);
  tm t;
  localtime_r(&tv.tv_sec, &t);
  char ts[32];
  strftime(ts, sizeof ts, "%Y-%m-%d %H:%M:%S", &t);
  std::lock_guard<std::mutex> lock(g_log_mutex);
  fprintf(stderr, "[%s.%03d] [%s] %s\n", ts, static_cast<int>(tv.tv_usec / 1000), names[lv], msg);
}

struct Timer  // include/kmdiff/time.hpp, src/time.cpp:29-47
{
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  std::string formatted() const
  {
    long s = static_cast<long>(seconds());
    const long d = s / 86400; s %= 86400;
    const long h = s / 3600; s %= 3600;
    const long m = s / 60; s %= 60;
    char buf[64]; int n = 0;
    if (d) n += snprintf(buf + n, sizeof buf - n, "%02ldd", d);
    if (h) n += snprintf(buf + n, sizeof buf - n, "%02ldh", h);
    if (m) n += snprintf(buf + n, sizeof buf - n, "%02ldm", m);
    snprintf(buf + n, sizeof buf - n, "%02lds", s);
    return buf;
  }
};

// ---- options (cmd/diff_opt.hpp, cli.cpp:148-362) ------------------------------------------------------

struct DiffOptions
{
  std::string kmtricks_dir, output_directory = "./kmdiff_output", verbosity = "info";
  size_t nb_controls = 0, nb_cases = 0, nb_threads = std::max(1u, std::thread::hardware_concurrency());
  double threshold = 0.05;
  size_t cutoff = 100000;
  int correction = KMD_CORR_BONFERRONI;
  bool kff = false, in_memory = false, cpr = false, keep_tmp = false, save_sk = false;
  std::string model_lib_path, model_config;
  bool pop_correction = false;
  std::string gender, covariates;
  double kmer_pca = 0.001;
  size_t ploidy = 2, npc = 2, seed = 0, log_size = 10000;
  // extensions of this implementation
  int gpus = 0;               // 0 = chosen from the input size: one device per 64 GiB
  size_t count_bytes = 4;     // the reference fixes this at compile time (MAX_C); here a runtime option
  size_t max_survivors = 0;
  size_t max_sampled = 0;     // capacity of the per-GPU presence matrix (--pop-correction); 0 = 2097152
  bool device_lz4 = false;    // count files: decode the LZ4 frames on the device whatever the cohort size
  size_t merge_records = 32u << 20;   // count-file input: records merged on the device per call (all samples together)
  size_t total_kmers = 0;
};

[[noreturn]] void usage(int code)
{
  fputs(
    "kmdiff-b200 diff — differential k-mers analysis (B200-native `kmdiff diff`)\n\n"
    "  -d/--km-run DIR         kmtricks run directory.\n"
    "  -o/--output-dir DIR     output directory. {./kmdiff_output}\n"
    "  -1/--nb-controls INT    number of controls.\n"
    "  -2/--nb-cases INT       number of cases.\n"
    "  -s/--significance FLOAT significance threshold in [0, 0.5]. {0.05}\n"
    "  -u/--cutoff INT         divide the significance threshold by N for the prefilter. {100000}\n"
    "  -c/--correction STR     bonferroni|benjamini|sidak|holm|disabled {bonferroni}\n"
    "  -f/--kff-output         output significant k-mers in kff format (no p-values, no counts).\n"
    "  -m/--in-memory  -r/--cpr   accepted and ignored, as in the reference\n"
    "  --keep-tmp              keep partitions/p{i}_uncorrected spill files.\n"
    "  --save-sk               build the matrix of significant k-mers.\n"
    " [population stratification]\n"
    "  --pop-correction        apply correction of population stratification.\n"
    "  --kmer-pca FLOAT        proportion of k-mers used for PCA (in [0.0, 0.05]). {0.001}\n"
    "  --ploidy INT            ploidy level. {2}\n"
    "  --n-pc INT              number of principal components (in [2, 10]). {2}\n"
    "  --random-seed INT       seed of the k-mer sampler. {0}\n"
    "  --gender/--covariates   accepted; both fail in the reference too (see DESIGN.md).\n"
    " [custom model]\n"
    "  --cmodel STR            path to model shared library (IModel plugin).\n"
    "  --config STR            model config.\n"
    " [dev]\n"
    "  --log-factorial INT     size of precomputed table. {10000}\n"
    " [common]\n"
    "  -t/--threads INT        host threads (matrix readers / plugin model). {hw}\n"
    "  -v/--verbose STR        debug|info|warning|error {info}\n"
    "  -h/--help  --version\n"
    " [this implementation]\n"
    "  --gpus INT              GPUs to shard the partitions over. {one per 64 GiB of input}\n"
    "  --count-bytes INT       count width of the matrices: 1, 2 or 4 (the reference's compile-time MAX_C). {4}\n"
    "  --max-survivors INT     capacity of the per-GPU survivor buffers. {4194304}\n"
    "  --max-sampled INT       capacity of the per-GPU sampled k-mer buffer (--pop-correction). {2097152}\n"
    "  --device-lz4            count files: decode the LZ4 frames on the GPU (default from 2048 samples on).\n",
    code ? stderr : stdout);
  exit(code);
}

[[noreturn]] void cli_error(const std::string& msg)
{
  fprintf(stderr, "[Error] %s\n", msg.c_str());
  exit(1);
}

bool is_number(const std::string& s) { return !s.empty() && s.find_first_not_of("0123456789") == std::string::npos; }

DiffOptions parse_cli(int argc, char** argv)
{
  if (argc < 2) usage(1);
  const std::string cmd = argv[1];
  if (cmd == "-h" || cmd == "--help") usage(0);
  if (cmd == "--version") { puts("kmdiff-b200 0.1.0 (kmdiff 1.1.0 `diff` interface)"); exit(0); }
  if (cmd == "infos") { puts("kmdiff-b200 0.1.0: sm_100a CUDA hot path, C-ABI v1; commands: diff"); exit(0); }
  if (cmd != "diff") cli_error("unknown command '" + cmd + "' (this build implements `diff`; `count` is kmtricks' job)");
  DiffOptions o;
  bool have_d = false, have_1 = false, have_2 = false;
  for (int i = 2; i < argc; i++)
  {
    const std::string a = argv[i];
    auto value = [&]() -> std::string {
      if (i + 1 >= argc) cli_error(a + ": missing value");
      return argv[++i];
    };
    auto number = [&]() -> size_t {
      const std::string v = value();
      if (!is_number(v)) cli_error(a + " " + v + " : Not a number!");
      return std::stoull(v);
    };
    if (a == "-d" || a == "--km-run")
    {
      o.kmtricks_dir = value(); have_d = true;
      if (!fs::is_directory(o.kmtricks_dir)) cli_error(a + " " + o.kmtricks_dir + " : Directory doesn't exist.");
      if (!fs::exists(o.kmtricks_dir + "/kmtricks.fof")) cli_error(a + " " + o.kmtricks_dir + " : Not a kmtricks runtime directory.");
    }
    else if (a == "-o" || a == "--output-dir") o.output_directory = value();
    else if (a == "-1" || a == "--nb-controls") { o.nb_controls = number(); have_1 = true; }
    else if (a == "-2" || a == "--nb-cases") { o.nb_cases = number(); have_2 = true; }
    else if (a == "-s" || a == "--significance")
    {
      const std::string v = value();
      char* end = nullptr;
      o.threshold = strtod(v.c_str(), &end);
      if (end == v.c_str()) cli_error(a + " " + v + ": Not a number!");
      if (o.threshold < 0.0 || o.threshold > 0.5) cli_error("Not in range [0.0, 0.5]");
    }
    else if (a == "-u" || a == "--cutoff") o.cutoff = number();
    else if (a == "-c" || a == "--correction")
    {
      const std::string v = value();
      if (v == "bonferroni") o.correction = KMD_CORR_BONFERRONI;
      else if (v == "benjamini") o.correction = KMD_CORR_BENJAMINI;
      else if (v == "sidak") o.correction = KMD_CORR_SIDAK;
      else if (v == "holm") o.correction = KMD_CORR_HOLM;
      else if (v == "disabled") o.correction = KMD_CORR_NOTHING;
      else cli_error(a + " " + v + " : Not in bonferroni|benjamini|sidak|holm|disabled");
    }
    else if (a == "-f" || a == "--kff-output") o.kff = true;
    else if (a == "-m" || a == "--in-memory") o.in_memory = true;
    else if (a == "-r" || a == "--cpr") o.cpr = true;
    else if (a == "--keep-tmp") o.keep_tmp = true;
    else if (a == "--save-sk") o.save_sk = true;
    else if (a == "--cmodel")
    {
      o.model_lib_path = value();
      const std::string ext = fs::path(o.model_lib_path).extension().string();
      if (ext != ".so" && ext != ".dylib") cli_error(a + " " + o.model_lib_path + " : extension must be so|dylib");
    }
    else if (a == "--config") o.model_config = value();
    else if (a == "--pop-correction") o.pop_correction = true;
    else if (a == "--gender") o.gender = value();
    else if (a == "--covariates") o.covariates = value();
    else if (a == "--kmer-pca") o.kmer_pca = atof(value().c_str());
    else if (a == "--ploidy") o.ploidy = number();
    else if (a == "--n-pc") o.npc = number();
    else if (a == "--learning-rate" || a == "--max-iteration" || a == "--epsilon") value();  // inert in the reference too
    else if (a == "--stand" || a == "--irls") {}
    else if (a == "--random-seed") o.seed = number();
    else if (a == "--log-factorial") o.log_size = number();
    else if (a == "-t" || a == "--threads") o.nb_threads = std::max<size_t>(1, number());
    else if (a == "-v" || a == "--verbose")
    {
      o.verbosity = value();
      if (o.verbosity != "debug" && o.verbosity != "info" && o.verbosity != "warning" && o.verbosity != "error")
        cli_error(a + " " + o.verbosity + " : Not in debug|info|warning|error");
    }
    else if (a == "-h" || a == "--help") usage(0);
    else if (a == "--version") { puts("kmdiff-b200 0.1.0"); exit(0); }
    else if (a == "--gpus") o.gpus = static_cast<int>(number());
    else if (a == "--count-bytes") o.count_bytes = number();
    else if (a == "--max-survivors") o.max_survivors = number();
    else if (a == "--max-sampled") o.max_sampled = number();
    else if (a == "--merge-records") o.merge_records = number();
    else if (a == "--device-lz4") o.device_lz4 = true;
    else cli_error("unknown parameter " + a);
  }
  if (!have_d) cli_error("-d/--km-run is required");
  if (!have_1 || !have_2) cli_error("-1/--nb-controls and -2/--nb-cases are required");
  if (o.count_bytes != 1 && o.count_bytes != 2 && o.count_bytes != 4) cli_error("--count-bytes must be 1, 2 or 4");
  return o;
}

// options.bin: 37 bytes (cmd/diff_opt.hpp:78-88) + total_kmers appended — the reference forgets it and a
// resumed run then corrects with m = 0; the extra 8 bytes are ignored by the reference's loader.
void dump_opt(const DiffOptions& o, const std::string& path)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return;
  const double thr = o.threshold, cutoff = static_cast<double>(o.cutoff), pca = o.kmer_pca;
  const int32_t corr = o.correction;
  const uint8_t pop = o.pop_correction;
  const uint64_t npc = o.npc, total = o.total_kmers;
  fwrite(&thr, 8, 1, f); fwrite(&cutoff, 8, 1, f); fwrite(&corr, 4, 1, f); fwrite(&pop, 1, 1, f);
  fwrite(&pca, 8, 1, f); fwrite(&npc, 8, 1, f); fwrite(&total, 8, 1, f);
  fclose(f);
}

// load_opt (cmd/diff_opt.hpp:90-104) + our total_kmers extension (0 when the file was written by the reference)
struct PrevOptions { bool valid = false; double threshold = 0, cutoff = 0, kmer_pca = 0; int32_t correction = 0; uint8_t pop = 0; uint64_t npc = 0, total_kmers = 0; };

PrevOptions load_opt(const std::string& path)
{
  PrevOptions o;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return o;
  uint8_t b[45] = {0};
  const size_t got = fread(b, 1, sizeof b, f);
  fclose(f);
  if (got < 37) return o;
  memcpy(&o.threshold, b, 8); memcpy(&o.cutoff, b + 8, 8); memcpy(&o.correction, b + 16, 4); o.pop = b[20];
  memcpy(&o.kmer_pca, b + 21, 8); memcpy(&o.npc, b + 29, 8);
  if (got >= 45) memcpy(&o.total_kmers, b + 37, 8);
  o.valid = true;
  return o;
}

// compare_opt (cmd/diff_opt.hpp:106-133): bit 0 redo the scan, bit 1 redo the stratification, bit 2 redo the correction
unsigned compare_opt(const DiffOptions& o, const PrevOptions& prev)
{
  unsigned r = 0;
  if (o.threshold != prev.threshold || static_cast<double>(o.cutoff) != prev.cutoff) r |= 0b1;
  if (prev.pop && o.pop_correction)
  {
    if (o.kmer_pca != prev.kmer_pca) r |= 0b11;
    if (o.npc != prev.npc) r |= 0b10;
  }
  if (!prev.pop && o.pop_correction) r |= 0b11;
  if (o.correction != prev.correction) r |= 0b100;
  if (prev.pop && !o.pop_correction) r |= 0b100;
  return r;
}

// ---- device contexts ------------------------------------------------------------------------------------

struct CtxDeleter { void operator()(kmd_context* c) const { kmd_destroy(c); } };
using ctx_ptr = std::unique_ptr<kmd_context, CtxDeleter>;

void ck(kmd_context* ctx, int rc, const char* what)
{
  if (rc != KMD_OK) throw KmdError("GpuError", std::string(what) + ": " + kmd_last_error(ctx));
}

// Partition tag and row index of a survivor: rows are numbered inside their partition; the partition
// index sits above bit 40 so that device-side row order == (partition, row) order.
constexpr int kPartShift = 40;

struct Partition
{
  std::string path;                        // matrix_{p}.count[.lz4]; empty in count-file mode
  uint32_t index;
  uint64_t bytes;
  uint64_t draw_base = 0;
  bool raw = false;                        // uncompressed matrix: read in slices by several threads
  std::vector<std::string> sample_paths;   // count-file mode: counts/partition_{p}/<sample>.kmer.lz4, fof order
};

// One GPU's share.  Matrix input: reader threads bring slices of the partition files into pinned staging buffers
// (pread of the page cache for raw matrices — many slices per file, so every -t thread has work whatever the
// partition count; one sequential LZ4 stream per compressed file), the feeder thread enqueues the asynchronous
// H2D copy + scan of every buffer (kmd_scan_host_async) and hands the buffer straight back; a reader waits for its
// buffer's copy (kmd_host_wait) only when it is about to refill it.  Replaces the per-row istream reads of
// matrix_proxy::merge (merge.hpp:194-203) under global_merge's partition pool (merge.hpp:341-388).
class GpuWorker
{
 public:
  // buf_bytes: size of one pinned staging buffer; max_bufs: pool limit (buffers are pinned on demand by the reader
  // threads — pinning costs ~0.3 s per GB and overlaps the first reads)
  GpuWorker(int device, const DiffOptions& o, const RunConfig& cfg, uint64_t sum_controls, uint64_t sum_cases, size_t n_readers = 1,
            size_t buf_bytes = 16u << 20, size_t max_bufs = 0)
    : m_device(device), m_opt(o), m_n_readers(std::max<size_t>(1, n_readers)),
      m_bufs(max_bufs ? max_bufs : std::max<size_t>(1, n_readers) + 8)
  {
    kmd_bind_host_thread(device);   // CPU affinity + memory policy of the GPU's NUMA node (no-op when unknown)
    kmd_config c{};
    c.abi_version = KMD_ABI_VERSION;
    c.device = device;
    c.nb_controls = static_cast<uint32_t>(o.nb_controls);
    c.nb_cases = static_cast<uint32_t>(o.nb_cases);
    c.kmer_size = cfg.kmer_size;
    c.kmer_words = (cfg.kmer_size + 31) / 32;
    c.count_bytes = static_cast<uint32_t>(o.count_bytes);
    c.sum_controls = sum_controls;
    c.sum_cases = sum_cases;
    c.threshold = o.threshold / static_cast<double>(o.cutoff);   // cmd/diff.hpp:146
    c.lf_table = nullptr;
    c.lf_size = o.log_size;
    c.max_survivors = o.max_survivors;
    // spill files carry every survivor's counts, like a default (WITH_POPSTRAT) build of the reference (kmer.hpp:118-123)
    c.flags = (o.keep_tmp || o.save_sk) ? KMD_FLAG_KEEP_COUNTS : 0;
    kmd_context* raw = nullptr;
    const int rc = kmd_create(&c, &raw);
    if (rc != KMD_OK) throw KmdError("GpuError", std::string("kmd_create: ") + kmd_last_error(nullptr));
    m_ctx.reset(raw);
    if (o.pop_correction)
    {
      // Sampler(geno, snp, opt->kmer_pca, opt->seed) (cmd/diff.hpp:325-329)
      kmd_popstrat_config pc{};
      pc.kmer_pca = o.kmer_pca;
      pc.seed = o.seed;
      pc.is_diploid = o.ploidy == 2;
      pc.n_pc = static_cast<uint32_t>(o.npc);
      pc.max_sampled = o.max_sampled;
      ck(raw, kmd_popstrat_enable(raw, &pc), "kmd_popstrat_enable");
    }
    m_row_bytes = 8ull * c.kmer_words + (o.nb_controls + o.nb_cases) * o.count_bytes;
    m_buf_rows = std::max<uint64_t>(1, buf_bytes / m_row_bytes);
  }
  ~GpuWorker()
  {
    for (auto& b : m_bufs) kmd_host_free(b.data);
    kmd_host_free(m_arena);
    for (int fd : m_fds) if (fd >= 0) ::close(fd);
  }

  kmd_context* ctx() { return m_ctx.get(); }
  int device() const { return m_device; }
  std::vector<Partition> partitions;
  std::exception_ptr error;

  void run()
  {
    try
    {
      if (!partitions.empty() && !partitions.front().sample_paths.empty())
      {
        merge_loop();
        ck(m_ctx.get(), kmd_sync(m_ctx.get()), "kmd_sync");
        return;
      }
      plan_items();
      m_next_item = 0;
      m_readers_left = static_cast<int>(m_n_readers);
      std::vector<std::thread> readers;
      for (size_t r = 0; r < m_n_readers; r++) readers.emplace_back([this] { read_loop(); });
      std::exception_ptr feed_error;
      try { feed_loop(false); }
      catch (...)
      {
        // keep recycling the readers' buffers (without scanning) until the last one arrives, so that no reader
        // stays blocked in acquire_free() and every thread can be joined before the error is reported
        feed_error = std::current_exception();
        m_next_item = m_items.size();
        try { feed_loop(true); } catch (...) {}
      }
      for (auto& r : readers) r.join();
      if (feed_error) std::rethrow_exception(feed_error);
      if (m_reader_error) std::rethrow_exception(m_reader_error);
      ck(m_ctx.get(), kmd_sync(m_ctx.get()), "kmd_sync");
    }
    catch (...) { error = std::current_exception(); }
  }

 private:
  struct Buf { uint8_t* data = nullptr; uint64_t rows = 0; uint64_t row_base = 0; uint64_t draw_base = 0; uint32_t partition = 0; bool last = false;
               uint64_t ticket = 0; /* kmd_scan_host_async ticket of the copy that last read this buffer */ };
  // one unit of reader work: rows [first_row, first_row + rows) of a raw matrix, or a whole compressed matrix (rows == 0)
  struct Item { size_t part; uint64_t first_row; uint64_t rows; };

  void plan_items()
  {
    m_items.clear();
    m_fds.assign(partitions.size(), -1);
    for (size_t pi = 0; pi < partitions.size(); pi++)
    {
      const Partition& p = partitions[pi];
      MatrixFile mf(p.path, m_row_bytes);
      const uint32_t slots = (mf.hdr.kmer_size + 31) / 32;
      if (mf.hdr.nb_counts != m_opt.nb_controls + m_opt.nb_cases)
        throw KmdError("InputError", p.path + ": " + std::to_string(mf.hdr.nb_counts) + " counts per row, expected " +
                                       std::to_string(m_opt.nb_controls + m_opt.nb_cases));
      if (mf.hdr.kmer_slots != slots) throw KmdError("InputError", p.path + ": unexpected kmer_slots");
      if (mf.hdr.compressed) { m_items.push_back({pi, 0, 0}); continue; }
      if ((p.bytes - kMatrixHeaderBytes) % m_row_bytes) throw KmdError("InputError", p.path + ": truncated row");
      m_fds[pi] = ::open(p.path.c_str(), O_RDONLY);
      if (m_fds[pi] < 0) throw KmdError("IOError", "Unable to open " + p.path);
      for (uint64_t r = 0; r < mf.rows_if_raw; r += m_buf_rows) m_items.push_back({pi, r, std::min(m_buf_rows, mf.rows_if_raw - r)});
      logf(L_DEBUG, "partition %u: %llu rows (gpu %d)", p.index, static_cast<unsigned long long>(mf.rows_if_raw), m_device);
    }
  }

  // Run directories without matrices (cmd/diff.hpp:88-95 -> global_merge::merge(), merge.hpp:237-309): every
  // partition is N per-sample files of sorted (k-mer, count) records.  The reference walks them with
  // km::KmerMerger, one row at a time; here each stream is buffered, the records up to the smallest "last
  // buffered k-mer" of the streams still open form a complete k-mer range, and that range is merged and tested
  // on the device (kmd_merge_scan_host).  ab_min is 1 for every sample, as cmd/diff.hpp:109 passes it.
  void merge_loop()
  {
    const uint32_t N = static_cast<uint32_t>(m_opt.nb_controls + m_opt.nb_cases);
    const uint32_t kw = static_cast<uint32_t>((m_row_bytes - static_cast<uint64_t>(N) * m_opt.count_bytes) / 8);
    const uint32_t B = static_cast<uint32_t>(m_opt.count_bytes);
    const size_t rec = 8ull * kw + B;
    const uint64_t budget = std::max<uint64_t>(1ull << 12, m_opt.merge_records);
    // per-stream buffer: the call budget shared by the samples, but no more than the largest stream can decode to
    // (LZ4 shrinks sorted records to no less than ~1/6): pinning costs ~0.4 s per GB
    uint64_t max_file = 0;
    for (const Partition& p : partitions)
      for (const std::string& sp : p.sample_paths) { std::error_code ec; const auto sz = fs::file_size(sp, ec); if (!ec) max_file = std::max<uint64_t>(max_file, sz); }
    const size_t per = static_cast<size_t>(std::max<uint64_t>(1024, std::min<uint64_t>(budget / N, max_file * 6 / rec + 1024)));
    const std::vector<uint32_t> ab_min(N, 1);
    // record buffers in pinned memory: the H2D copies of kmd_merge_scan_host then run at PCIe rate
    struct PinnedBuf
    {
      uint8_t* p = nullptr;
      explicit PinnedBuf(size_t bytes) : p(static_cast<uint8_t*>(kmd_host_alloc(bytes))) { if (!p) throw KmdError("GpuError", "kmd_host_alloc failed"); }
      ~PinnedBuf() { kmd_host_free(p); }
      PinnedBuf(const PinnedBuf&) = delete;
      PinnedBuf& operator=(const PinnedBuf&) = delete;
      uint8_t* data() const { return p; }
    };
    struct Stream { std::unique_ptr<KmerFile> file; std::unique_ptr<PinnedBuf> pin; std::vector<uint8_t> disk; size_t n = 0; bool eof = false; };
    auto key_less = [kw](const uint8_t* a, const uint8_t* b) {
      for (int w = static_cast<int>(kw) - 1; w >= 0; w--)
      {
        uint64_t x, y;
        memcpy(&x, a + 8 * w, 8); memcpy(&y, b + 8 * w, 8);
        if (x != y) return x < y;
      }
      return false;
    };
    // Lanes take partitions in turn: while one lane's range is merged and tested on the device (the context is used by
    // one thread at a time: m_gpu_mu), another lane can read and decode its next range on the host threads.  Measured
    // (profiles/r02m: 1.6e7 rows as 32 x 20 LZ4 count files): one lane 0.85 s, two 1.1-1.3 s, three 2.2 s — the lanes'
    // pinned allocations and copies contend inside the driver for more than the overlap wins, so ONE lane is the
    // default and KMD_MERGE_LANES the knob.  The Sampler's draw index follows the rows merged so far, so
    // --pop-correction always keeps one lane (one ordered stream).
    size_t n_lanes = 1;
    if (const char* e = getenv("KMD_MERGE_LANES")) n_lanes = std::max<size_t>(1, static_cast<size_t>(atol(e)));
    if (m_opt.pop_correction || partitions.size() < 2) n_lanes = 1;
    const size_t fill_threads = std::max<size_t>(1, std::min<size_t>(m_opt.nb_threads / n_lanes, N));
    std::atomic<size_t> next_part{0};
    std::exception_ptr lane_error;
    std::mutex lane_mu;
    auto lane = [&] {
      kmd_bind_host_thread(m_device);
      try
      {
        std::vector<std::unique_ptr<PinnedBuf>> pins(N);
        for (auto& pb : pins) pb = std::make_unique<PinnedBuf>(per * rec);
        for (size_t pi; (pi = next_part.fetch_add(1)) < partitions.size();)
        {
          { std::lock_guard<std::mutex> lk(lane_mu); if (lane_error) break; }
          const Partition& p = partitions[pi];
          {
            if (p.sample_paths.size() != N) throw KmdError("ConfigError", "kmtricks.fof does not list --nb-controls + --nb-cases samples.");
            {
              std::lock_guard<std::mutex> gpu(m_gpu_mu);
              if (merge_whole_partition(p, N, kw, B, budget)) continue;
            }
            std::vector<Stream> st(N);
            struct Returner { std::vector<Stream>& st; std::vector<std::unique_ptr<PinnedBuf>>& pins;
                              ~Returner() { for (size_t s = 0; s < st.size(); s++) if (st[s].pin) pins[s] = std::move(st[s].pin); } } returner{st, pins};
            for (uint32_t s = 0; s < N; s++)
            {
              st[s].pin = std::move(pins[s]);
              st[s].file = std::make_unique<KmerFile>(p.sample_paths[s]);
              const KmerHeader& h = st[s].file->hdr;
              if (h.kmer_slots != kw) throw KmdError("InputError", p.sample_paths[s] + ": unexpected kmer_slots");
              if (h.count_slots == 0 || h.count_slots > 4) throw KmdError("InputError", p.sample_paths[s] + ": unexpected count_slots");
            }
            uint64_t local = 0;
            for (;;)
            {
              // refill every open stream up to `per` buffered records, the samples spread over the reader threads
              std::atomic<uint32_t> next{0};
              std::exception_ptr err;
              std::mutex err_mu;
              auto fill = [&] {
                for (uint32_t s; (s = next.fetch_add(1)) < N;)
                {
                  Stream& t = st[s];
                  if (t.eof || t.n >= per) continue;
                  try
                  {
                    const size_t want = per - t.n;
                    const uint32_t cs = t.file->hdr.count_slots;
                    size_t got_rec;
                    if (cs == B)
                    {
                      const size_t got = t.file->payload->read(t.pin->data() + t.n * rec, want * rec);
                      if (got % rec) throw KmdError("InputError", p.sample_paths[s] + ": truncated record");
                      got_rec = got / rec;
                    }
                    else
                    {
                      // count stored narrower or wider than this run's count type: KmerReader::read copies count_slots
                      // bytes into a count_type (kmer_file.hpp:132); widen with zeros / keep the low bytes
                      const size_t drec = 8ull * kw + cs;
                      t.disk.resize(want * drec);
                      const size_t got = t.file->payload->read(t.disk.data(), want * drec);
                      if (got % drec) throw KmdError("InputError", p.sample_paths[s] + ": truncated record");
                      got_rec = got / drec;
                      for (size_t i = 0; i < got_rec; i++)
                      {
                        uint8_t* dst = t.pin->data() + (t.n + i) * rec;
                        memset(dst, 0, rec);
                        memcpy(dst, t.disk.data() + i * drec, 8ull * kw + std::min<uint32_t>(cs, B));
                      }
                    }
                    t.n += got_rec;
                    if (got_rec < want) t.eof = true;
                  }
                  catch (...) { std::lock_guard<std::mutex> lk(err_mu); err = std::current_exception(); }
                }
              };
              const size_t n_thr = fill_threads;
              std::vector<std::thread> pool;
              for (size_t t = 1; t < n_thr; t++) pool.emplace_back(fill);
              fill();
              for (auto& t : pool) t.join();
              if (err) std::rethrow_exception(err);

              // every stream still open may later deliver k-mers above its last buffered one only
              const uint8_t* bound = nullptr;
              for (const Stream& t : st)
                if (!t.eof && t.n)
                {
                  const uint8_t* last = t.pin->data() + (t.n - 1) * rec;
                  if (!bound || key_less(last, bound)) bound = last;
                }
              std::vector<uint8_t> bound_key;
              if (bound) bound_key.assign(bound, bound + 8 * kw);
              std::vector<const void*> ptrs(N);
              std::vector<uint64_t> cnt(N);
              uint64_t total = 0;
              for (uint32_t s = 0; s < N; s++)
              {
                size_t cut = st[s].n;
                if (bound)
                {
                  // first record whose k-mer is > bound
                  size_t lo = 0, hi = st[s].n;
                  while (lo < hi)
                  {
                    const size_t mid = (lo + hi) / 2;
                    if (key_less(bound_key.data(), st[s].pin->data() + mid * rec)) hi = mid; else lo = mid + 1;
                  }
                  cut = lo;
                }
                ptrs[s] = st[s].pin->data();
                cnt[s] = cut;
                total += cut;
              }
              if (total == 0) break;   // every stream is exhausted
              uint64_t n_rows = 0;
              {
                std::lock_guard<std::mutex> gpu(m_gpu_mu);
                if (m_opt.pop_correction)
                  ck(m_ctx.get(), kmd_popstrat_set_draw_index(m_ctx.get(), m_draw_next), "kmd_popstrat_set_draw_index");
                ck(m_ctx.get(), kmd_merge_scan_host(m_ctx.get(), ptrs.data(), cnt.data(), ab_min.data(),
                                                    (static_cast<uint64_t>(p.index) << kPartShift) + local, p.index, &n_rows),
                   "kmd_merge_scan_host");
                m_draw_next += n_rows;
              }
              local += n_rows;
              for (uint32_t s = 0; s < N; s++)
              {
                Stream& t = st[s];
                if (cnt[s] && cnt[s] < t.n) memmove(t.pin->data(), t.pin->data() + cnt[s] * rec, (t.n - cnt[s]) * rec);
                t.n -= cnt[s];
              }
            }
            logf(L_DEBUG, "partition %u: %llu rows merged from %u count files (gpu %d)", p.index, static_cast<unsigned long long>(local), N, m_device);
          }
        }
      }
      catch (...) { std::lock_guard<std::mutex> lk(lane_mu); if (!lane_error) lane_error = std::current_exception(); }
    };
    std::vector<std::thread> lanes;
    for (size_t l = 1; l < n_lanes; l++) lanes.emplace_back(lane);
    lane();
    for (auto& t : lanes) t.join();
    if (lane_error) std::rethrow_exception(lane_error);
  }

  // A partition whose files are small enough goes to the device as it lies on disk: the N files are read whole into
  // one pinned arena (the -t threads share them), cross PCIe still compressed, and are decoded there — one warp per
  // sample stream (kmd_merge_scan_lz4_host) — then merged and tested in one call.  Returns false when the partition has
  // to be streamed in k-mer ranges by the loop above instead (too many records for one call, or a count width on
  // disk that differs from this run's).
  bool merge_whole_partition(const Partition& p, uint32_t N, uint32_t kw, uint32_t B, uint64_t budget)
  {
    // One warp decodes one stream at ~27 MB/s over its two passes (a serial chain of short sequences;
    // profiles/r02x/dev_lz4_rate.json), a host thread at ~3 GB/s (profiles/r03k/host_lz4.jsonl): the device draws level
    // with 16 host threads at ~1 700 streams in flight, so it is taken from 2048 samples per partition on (or on
    // request: --device-lz4 / KMD_DEVICE_LZ4=1, e.g. where few host cores feed many GPUs); smaller cohorts are decoded by
    // the -t host threads.
    const char* force = getenv("KMD_DEVICE_LZ4");
    const bool on = force ? atoi(force) != 0 : (m_opt.device_lz4 || N >= 2048);
    if (!on) return false;
    const size_t rec = 8ull * kw + B;
    std::vector<uint64_t> fsize(N), foff(N + 1, 0);
    for (uint32_t s = 0; s < N; s++)
    {
      std::error_code ec;
      fsize[s] = static_cast<uint64_t>(fs::file_size(p.sample_paths[s], ec));
      if (ec || fsize[s] < kKmerHeaderBytes) throw KmdError("IOError", "Unable to open " + p.sample_paths[s]);
      foff[s + 1] = foff[s] + (fsize[s] + 63) / 64 * 64;
    }
    // compressed k-mer records shrink to no less than ~1/4: do not even read a partition that cannot fit
    if (foff[N] / rec > 4 * budget) return false;
    if (foff[N] > m_arena_bytes)
    {
      kmd_host_free(m_arena);
      m_arena = nullptr; m_arena_bytes = 0;
      const size_t want = static_cast<size_t>(foff[N] + foff[N] / 4 + (1u << 20));
      m_arena = static_cast<uint8_t*>(kmd_host_alloc(want));
      if (!m_arena) throw KmdError("GpuError", "kmd_host_alloc failed");
      m_arena_bytes = want;
    }
    std::atomic<uint32_t> next{0};
    std::exception_ptr err;
    std::mutex err_mu;
    auto slurp = [&] {
      for (uint32_t s; (s = next.fetch_add(1)) < N;)
      {
        try
        {
          const int fd = ::open(p.sample_paths[s].c_str(), O_RDONLY);
          if (fd < 0) throw KmdError("IOError", "Unable to open " + p.sample_paths[s]);
          size_t got = 0;
          while (got < fsize[s])
          {
            const ssize_t n = ::pread(fd, m_arena + foff[s] + got, fsize[s] - got, static_cast<off_t>(got));
            if (n <= 0) { ::close(fd); throw KmdError("IOError", p.sample_paths[s] + ": short read"); }
            got += static_cast<size_t>(n);
          }
          ::close(fd);
        }
        catch (...) { std::lock_guard<std::mutex> lk(err_mu); err = std::current_exception(); }
      }
    };
    {
      const size_t n_thr = std::max<size_t>(1, std::min<size_t>(m_opt.nb_threads, N));
      std::vector<std::thread> pool;
      for (size_t t = 1; t < n_thr; t++) pool.emplace_back(slurp);
      slurp();
      for (auto& t : pool) t.join();
      if (err) std::rethrow_exception(err);
    }
    std::vector<const void*> frames(N);
    std::vector<uint64_t> bytes(N);
    std::vector<uint8_t> compressed(N);
    for (uint32_t s = 0; s < N; s++)
    {
      KmerHeader h;
      memcpy(&h, m_arena + foff[s], sizeof h);
      if (h.km_magic != kKmMagic || h.kmer_magic != kKmerMagic) throw KmdError("IOError", "Invalid file format: " + p.sample_paths[s]);
      if (h.kmer_slots != kw) throw KmdError("InputError", p.sample_paths[s] + ": unexpected kmer_slots");
      if (h.count_slots != B) return false;   // widened / narrowed record by record on the host
      frames[s] = m_arena + foff[s] + kKmerHeaderBytes;
      bytes[s] = fsize[s] - kKmerHeaderBytes;
      compressed[s] = h.compressed != 0;
    }
    const std::vector<uint32_t> ab_min(N, 1);   // as cmd/diff.hpp:109 passes it
    if (m_opt.pop_correction)
      ck(m_ctx.get(), kmd_popstrat_set_draw_index(m_ctx.get(), m_draw_next), "kmd_popstrat_set_draw_index");
    uint64_t n_rows = 0, n_rec = 0;
    const int rc = kmd_merge_scan_lz4_host(m_ctx.get(), frames.data(), bytes.data(), compressed.data(), ab_min.data(),
                                           static_cast<uint64_t>(p.index) << kPartShift, p.index, budget, &n_rows, &n_rec);
    if (rc == KMD_ERR_OVERFLOW) return false;
    ck(m_ctx.get(), rc, "kmd_merge_scan_lz4_host");
    m_draw_next += n_rows;
    logf(L_DEBUG, "partition %u: %llu rows merged from %u count files, %llu records decoded on the device (gpu %d)", p.index,
         static_cast<unsigned long long>(n_rows), N, static_cast<unsigned long long>(n_rec), m_device);
    return true;
  }

  void read_loop()
  {
    kmd_bind_host_thread(m_device);
    try
    {
      for (size_t ii; (ii = m_next_item.fetch_add(1)) < m_items.size();)
      {
        const Item& it = m_items[ii];
        const Partition& p = partitions[it.part];
        if (it.rows)
        {
          // a slice of a raw matrix: page cache -> pinned buffer
          Buf* b = acquire_free();
          const size_t want = it.rows * m_row_bytes;
          size_t got = 0;
          while (got < want)
          {
            const ssize_t n = ::pread(m_fds[it.part], b->data + got, want - got,
                                      static_cast<off_t>(kMatrixHeaderBytes + it.first_row * m_row_bytes + got));
            if (n <= 0) { publish_empty(b); throw KmdError("IOError", p.path + ": short read"); }
            got += static_cast<size_t>(n);
          }
          b->rows = it.rows;
          b->row_base = (static_cast<uint64_t>(p.index) << kPartShift) + it.first_row;
          b->draw_base = p.draw_base + it.first_row;
          b->partition = p.index;
          publish(b);
          continue;
        }
        // a compressed matrix: one sequential LZ4 frame
        MatrixFile mf(p.path, m_row_bytes);
        uint64_t local = 0;
        for (;;)
        {
          Buf* b = acquire_free();
          size_t got;
          try { got = mf.payload->read(b->data, m_buf_rows * m_row_bytes); }
          catch (...) { publish_empty(b); throw; }
          if (got % m_row_bytes) { publish_empty(b); throw KmdError("InputError", p.path + ": truncated row"); }
          b->rows = got / m_row_bytes;
          b->row_base = (static_cast<uint64_t>(p.index) << kPartShift) + local;
          b->draw_base = p.draw_base + local;
          b->partition = p.index;
          local += b->rows;
          const bool eof = b->rows < m_buf_rows;
          publish(b);
          if (eof) break;
        }
        logf(L_DEBUG, "partition %u: %llu rows (gpu %d)", p.index, static_cast<unsigned long long>(local), m_device);
      }
    }
    catch (...)
    {
      std::lock_guard<std::mutex> lk(m_mu);
      if (!m_reader_error) m_reader_error = std::current_exception();
      m_next_item = m_items.size();   // the other readers stop after their current item
    }
    if (m_readers_left.fetch_sub(1) == 1)
    {
      // the last reader to finish closes the stream of buffers
      Buf* b = acquire_free();
      b->rows = 0; b->last = true;
      publish(b);
    }
  }

  // drain == true: an error is being unwound — buffers are recycled without being scanned
  void feed_loop(bool drain)
  {
    for (;;)
    {
      Buf* b;
      {
        std::unique_lock<std::mutex> lk(m_mu);
        m_cv_ready.wait(lk, [this] { return !m_ready.empty(); });
        b = m_ready.front(); m_ready.pop();
      }
      const bool last = b->last;
      struct Recycle
      {
        GpuWorker* w; Buf* b;
        ~Recycle()
        {
          { std::lock_guard<std::mutex> lk(w->m_mu); b->last = false; b->rows = 0; w->m_free.push(b); }
          w->m_cv_free.notify_one();
        }
      } recycle{this, b};
      if (b->rows && !drain)
      {
        // Sampler draws follow the reference's `-t 1` visiting order, not the (partition, row) tag
        if (m_opt.pop_correction) ck(m_ctx.get(), kmd_popstrat_set_draw_index(m_ctx.get(), b->draw_base), "kmd_popstrat_set_draw_index");
        ck(m_ctx.get(), kmd_scan_host_async(m_ctx.get(), b->data, b->rows, b->row_base, b->partition, &b->ticket), "kmd_scan_host_async");
      }
      if (last) break;
    }
  }

  // a free staging buffer whose last copy has completed; pinned on first use, by the reader that needs it
  Buf* acquire_free()
  {
    Buf* b = nullptr;
    {
      std::unique_lock<std::mutex> lk(m_mu);
      if (m_free.empty() && m_allocated < m_bufs.size()) b = &m_bufs[m_allocated++];
      else
      {
        m_cv_free.wait(lk, [this] { return !m_free.empty(); });
        b = m_free.front(); m_free.pop();
      }
    }
    if (!b->data)
    {
      b->data = static_cast<uint8_t*>(kmd_host_alloc(m_buf_rows * m_row_bytes));
      if (!b->data) { publish_empty(b); throw KmdError("GpuError", "kmd_host_alloc failed"); }
    }
    if (b->ticket)
    {
      if (kmd_host_wait(m_ctx.get(), b->ticket) != KMD_OK) { publish_empty(b); throw KmdError("GpuError", "kmd_host_wait failed"); }
      b->ticket = 0;
    }
    return b;
  }
  void publish(Buf* b)
  {
    { std::lock_guard<std::mutex> lk(m_mu); m_ready.push(b); }
    m_cv_ready.notify_one();
  }
  // hand a buffer back through the feeder without work (error paths: the pool must stay complete)
  void publish_empty(Buf* b) { b->rows = 0; b->last = false; publish(b); }

  int m_device;
  const DiffOptions& m_opt;
  ctx_ptr m_ctx;
  uint64_t m_row_bytes = 0, m_buf_rows = 0;
  std::mutex m_gpu_mu;          // count-file mode: the lanes take turns on the context
  uint8_t* m_arena = nullptr;   // count-file mode: pinned arena holding one partition's files
  size_t m_arena_bytes = 0;
  uint64_t m_draw_next = 0;   // count-file mode: Sampler draw index = rows merged so far (one worker, partition order)
  size_t m_n_readers = 1;
  std::vector<Buf> m_bufs;
  std::vector<Item> m_items;
  std::vector<int> m_fds;             // raw matrices, parallel to `partitions`
  std::atomic<size_t> m_next_item{0};
  std::atomic<int> m_readers_left{0};
  std::queue<Buf*> m_free, m_ready;
  size_t m_allocated = 0;             // buffers of m_bufs handed out so far
  std::mutex m_mu;
  std::condition_variable m_cv_free, m_cv_ready;
  std::exception_ptr m_reader_error;
};

// ---- user model plugins (model_manager.hpp:33-94) --------------------------------------------------------

template <size_t MAX_C>
struct PluginModel
{
  using model_t = kmdiff::IModel<MAX_C>;
  using count_t = typename model_t::count_type;
  void* handle = nullptr;
  std::string name;
  std::shared_ptr<model_t> model;

  void init(const std::string& path, const std::string& config)
  {
    handle = dlopen(path.c_str(), RTLD_LAZY);
    if (!handle) throw KmdError("PluginError", std::string("dlerror: ") + dlerror());
    dlerror();
    auto name_fn = reinterpret_cast<std::string (*)()>(dlsym(handle, "plugin_name"));
    if (const char* e = dlerror()) throw KmdError("PluginError", std::string("dlerror: ") + e);
    name = name_fn();
    const std::string sym = "create" + std::to_string(sizeof(count_t) * 8);
    auto create = reinterpret_cast<model_t* (*)()>(dlsym(handle, sym.c_str()));
    if (const char* e = dlerror()) throw KmdError("PluginError", std::string("dlerror: ") + e);
    model_t* m = create();
    m->configure(config);
    model.reset(m);   // deleted through the virtual destructor; the handle stays open, as in the reference
    logf(L_INFO, "Plugin '%s' loaded.", name.c_str());
  }
};

// The reference's diff_observer::process (merge.hpp:68-103) with a user model: per row on the host, one
// shared model instance, `threads` reader threads over the partitions.  Survivors go to the device-side
// accumulators of `ctx` through kmd_push_survivors.
template <size_t MAX_C>
void run_plugin_model(PluginModel<MAX_C>& pm, const std::vector<Partition>& parts, const DiffOptions& o, const RunConfig& cfg,
                      kmd_context* ctx)
{
  using count_t = typename PluginModel<MAX_C>::count_t;
  const uint32_t kw = (cfg.kmer_size + 31) / 32;
  const size_t n = o.nb_controls + o.nb_cases;
  const size_t R = 8 * kw + n * sizeof(count_t);
  const double thr = o.threshold / static_cast<double>(o.cutoff);
  std::atomic<size_t> next{0};
  std::mutex push_mu;
  std::exception_ptr err;
  auto work = [&] {
    try
    {
      std::vector<uint8_t> buf(R * 4096);
      std::vector<count_t> row(n);
      std::vector<kmd_survivor> out;
      for (size_t pi = next++; pi < parts.size(); pi = next++)
      {
        MatrixFile mf(parts[pi].path, R);
        uint64_t local = 0;
        out.clear();
        for (;;)
        {
          const size_t got = mf.payload->read(buf.data(), buf.size());
          const size_t rows = got / R;
          for (size_t r = 0; r < rows; r++)
          {
            const uint8_t* src = buf.data() + r * R;
            memcpy(row.data(), src + 8 * kw, n * sizeof(count_t));
            kmdiff::Range<count_t> rc(row, 0, o.nb_controls), rk(row, o.nb_controls, o.nb_cases);
            auto [p, sign, mc, mk] = pm.model->process(rc, rk);
            if (p <= thr)
            {
              kmd_survivor s{};
              memcpy(s.kmer, src, 8 * kw);
              s.p_value = p; s.sign = static_cast<int32_t>(sign); s.partition = parts[pi].index;
              s.mean_control = mc; s.mean_case = mk;
              s.row = (static_cast<uint64_t>(parts[pi].index) << kPartShift) + local + r;
              out.push_back(s);
            }
          }
          local += rows;
          if (got < buf.size()) break;
        }
        std::lock_guard<std::mutex> lk(push_mu);
        ck(ctx, kmd_push_survivors(ctx, out.data(), out.size(), local), "kmd_push_survivors");
      }
    }
    catch (...) { std::lock_guard<std::mutex> lk(push_mu); err = std::current_exception(); }
  };
  std::vector<std::thread> pool;
  for (size_t t = 0; t < std::min(o.nb_threads, std::max<size_t>(1, parts.size())); t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  if (err) std::rethrow_exception(err);
}

// ---- spill files: partitions/p{i}_uncorrected (accumulator.hpp:156-285, kmer.hpp:113-127) --------------

// KmerSign::dump of a default (WITH_POPSTRAT=ON, CMakeLists.txt:8) build of the reference (kmer.hpp:113-127):
//   k-mer | f64 p | i32 sign | f64 mean_control | f64 mean_case | u16 N | N x f64 counts
// in one LZ4 frame per partition.  For k <= 32 the files are interchangeable with the reference's.  For k > 32 the
// reference writes only 8 of the 16 k-mer bytes (SURVEY.md §8 hazard 1); all words are written here, and both record
// sizes are accepted on input.
void write_spill_files(const std::vector<kmd_survivor>& surv, uint32_t nb_partitions, uint32_t kmer_words, const std::string& dir,
                       const char* suffix, const std::vector<double>& counts, uint16_t n_counts)
{
  std::vector<std::unique_ptr<Lz4FrameWriter>> w(nb_partitions);
  std::vector<FILE*> files(nb_partitions, nullptr);
  for (uint32_t p = 0; p < nb_partitions; p++)
  {
    files[p] = fopen((dir + "/p" + std::to_string(p) + suffix).c_str(), "wb");
    if (!files[p]) throw KmdError("IOError", "Unable to create spill file in " + dir);
    w[p] = std::make_unique<Lz4FrameWriter>(files[p]);
  }
  for (size_t i = 0; i < surv.size(); i++)
  {
    const kmd_survivor& s = surv[i];
    if (s.partition >= nb_partitions) continue;
    Lz4FrameWriter& o = *w[s.partition];
    o.write(s.kmer, 8 * kmer_words);
    o.write(&s.p_value, 8);
    o.write(&s.sign, 4);
    o.write(&s.mean_control, 8);
    o.write(&s.mean_case, 8);
    o.write(&n_counts, 2);
    o.write(counts.data() + i * n_counts, 8ull * n_counts);
  }
  for (uint32_t p = 0; p < nb_partitions; p++) { w[p]->close(); fclose(files[p]); }
}

// FileAccumulator::get over a spill file of an earlier run — of this program or of the reference
// (accumulator.hpp:190-226, KmerSign::load kmer.hpp:101-111).  Rows are renumbered in file order.  counts_out
// receives the records' counts as count_bytes-wide integers (what kmd_push_survivors2 takes).
std::vector<kmd_survivor> read_spill_file(const std::string& path, uint32_t partition, uint32_t kmer_words, uint32_t n_samples,
                                          uint32_t count_bytes, std::vector<uint8_t>* counts_out)
{
  std::vector<kmd_survivor> out;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw KmdError("IOError", "Unable to open " + path);
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
  Lz4FrameReader in(f);
  std::vector<uint8_t> all;
  {
    std::vector<uint8_t> buf(1 << 20);
    for (size_t got; (got = in.read(buf.data(), buf.size())) > 0;) all.insert(all.end(), buf.begin(), buf.begin() + got);
  }
  if (all.empty()) return out;
  // record layout: the k-mer field is 8 * kmer_words bytes in our files, 8 bytes in the reference's whatever k is
  size_t kbytes = 8ull * kmer_words;
  auto tail_ok = [&](size_t kb) {
    uint16_t n = 0;
    if (all.size() < kb + 30) return false;
    memcpy(&n, all.data() + kb + 28, 2);
    return n == n_samples && all.size() % (kb + 30 + 8ull * n_samples) == 0;
  };
  if (!tail_ok(kbytes))
  {
    if (kmer_words > 1 && tail_ok(8)) kbytes = 8;   // a reference file for k > 32: the high word was never written
    else throw KmdError("IOError", path + ": not a spill file of this cohort (record size / sample count mismatch)");
  }
  const size_t rec = kbytes + 30 + 8ull * n_samples;
  uint64_t local = 0;
  for (size_t at = 0; at < all.size(); at += rec)
  {
    const uint8_t* r = all.data() + at;
    kmd_survivor s{};
    memcpy(s.kmer, r, kbytes);
    memcpy(&s.p_value, r + kbytes, 8);
    memcpy(&s.sign, r + kbytes + 8, 4);
    memcpy(&s.mean_control, r + kbytes + 12, 8);
    memcpy(&s.mean_case, r + kbytes + 20, 8);
    s.partition = partition;
    s.row = (static_cast<uint64_t>(partition) << kPartShift) + local++;
    out.push_back(s);
    if (counts_out)
      for (uint32_t j = 0; j < n_samples; j++)
      {
        double v;
        memcpy(&v, r + kbytes + 30 + 8ull * j, 8);
        const uint32_t c = v <= 0 ? 0u : v >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(v);
        const size_t o = counts_out->size();
        counts_out->resize(o + count_bytes);
        memcpy(counts_out->data() + o, &c, count_bytes);
      }
  }
  return out;
}

// ---- FASTA output (aggregator.hpp:27-71, 197-199) -----------------------------------------------------------

// ---- KFF output (-f/--kff-output): KffWriter (include/kmdiff/kff_utils.hpp:32-107) over kff-cpp-api 1.0 ----
// (thirdparty/kff-cpp-api/kff_io.cpp: Kff_file::open :67-114, complete_header :241-257, Section_GV :709-800,
// Section_Raw :891-953, 1018-1032, Section_Index :849-871, write_footer :549-566, close :167-206).
// The file the reference leaves: 8-byte header ("KFF", version 1.0, encoding byte A0 C1 G3 T2 = 0x1e, not unique,
// not canonical), empty metadata, one 'v' section {data_size = 0, k, max = 1} (names in std::map order), one 'r'
// section with one block per k-mer (max = 1 -> no k-mer-count field; ceil(k/4) bytes, first byte holds the k % 4
// leading bases; no data), an 'i' index of the two sections (offsets relative to the end of the index), a footer
// 'v' section {first_index, footer_size = 49} and the "KFF" signature.  All integers big-endian.
class KffWriter
{
 public:
  KffWriter(const std::string& path, uint32_t k) : m_k(k)
  {
    m_f = fopen(path.c_str(), "wb");
    if (!m_f) throw KmdError("IOError", "Unable to create " + path);
    const uint8_t head[8] = {'K', 'F', 'F', 1, 0, 0x1e, 0, 0};
    put(head, 8);
    be(0, 4);                       // metadata size
    m_pos_v = m_pos;
    put("v", 1); be(3, 8);
    var("data_size", 0); var("k", k); var("max", 1);
    m_pos_r = m_pos;
    put("r", 1); be(0, 8);          // block count, patched by close()
  }
  // 2 bits per base, first base most significant.  kmtricks words hold A0 C1 T2 G3 (kmer.hpp:38) and the
  // reference packs (char >> 1) & 3 (kff_utils.hpp:72-82), the same code — hence the header's encoding byte 0x1e
  void write(const uint64_t* words, uint32_t kmer_words)
  {
    uint8_t enc[64];
    const uint32_t nbytes = (m_k + 3) / 4, rem = m_k % 4;
    memset(enc, 0, nbytes);
    for (uint32_t i = 0; i < m_k; i++)
    {
      const uint32_t bit = 2 * (m_k - 1 - i);                 // base i (0 = first) sits at bits [bit, bit+1]
      const uint32_t v = bit / 64 < kmer_words ? static_cast<uint32_t>(words[bit / 64] >> (bit % 64)) & 3u : 0u;
      // byte 0 holds the `rem` leading bases (if any), then 4 bases per byte
      const uint32_t slot = rem ? (i < rem ? 0 : 1 + (i - rem) / 4) : i / 4;
      const uint32_t in_byte = rem ? (i < rem ? rem - 1 - i : 3 - (i - rem) % 4) : 3 - i % 4;
      enc[slot] |= static_cast<uint8_t>(v << (2 * in_byte));
    }
    put(enc, nbytes);
    m_blocks++;
  }
  void close()
  {
    if (!m_f) return;
    // footer: index section, then the variables that locate it
    const uint64_t index_pos = m_pos;
    const int64_t end_of_index = static_cast<int64_t>(index_pos) + 17 + 9 * 2;
    put("i", 1); be(2, 8);
    put("v", 1); be(static_cast<uint64_t>(static_cast<int64_t>(m_pos_v) - end_of_index), 8);
    put("r", 1); be(static_cast<uint64_t>(static_cast<int64_t>(m_pos_r) - end_of_index), 8);
    be(0, 8);                       // next index
    put("v", 1); be(2, 8);
    var("first_index", index_pos); var("footer_size", 9 + 2 * (12 + 8));
    put("KFF", 3);
    fseek(m_f, static_cast<long>(m_pos_r + 1), SEEK_SET);
    m_pos = m_pos_r + 1;
    be(m_blocks, 8);
    fclose(m_f);
    m_f = nullptr;
  }
  ~KffWriter() { close(); }

 private:
  void put(const void* p, size_t n) { fwrite(p, 1, n, m_f); m_pos += n; }
  void be(uint64_t v, int bytes)
  {
    uint8_t b[8];
    for (int i = 0; i < bytes; i++) b[i] = static_cast<uint8_t>(v >> (8 * (bytes - 1 - i)));
    put(b, bytes);
  }
  void var(const char* name, uint64_t v) { put(name, strlen(name) + 1); be(v, 8); }
  FILE* m_f = nullptr;
  uint32_t m_k;
  uint64_t m_pos = 0, m_pos_v = 0, m_pos_r = 0, m_blocks = 0;
};

struct FastaPair
{
  FILE* control = nullptr;
  FILE* cases = nullptr;
  std::unique_ptr<KffWriter> kff_control, kff_cases;    // -f/--kff-output: {control,case}_kmers.kff instead (aggregator.hpp:197-199)
  uint64_t n_control = 0, n_case = 0;
  FastaPair(const std::string& dir, bool kff, uint32_t k)
  {
    if (kff)
    {
      kff_control = std::make_unique<KffWriter>(dir + "/control_kmers.kff", k);
      kff_cases = std::make_unique<KffWriter>(dir + "/case_kmers.kff", k);
      return;
    }
    control = fopen((dir + "/control_kmers.fasta").c_str(), "w");
    cases = fopen((dir + "/case_kmers.fasta").c_str(), "w");
    if (!control || !cases) throw KmdError("IOError", "Unable to create output FASTA files in " + dir);
  }
  void write(const kmd_survivor& s, uint32_t kmer_words, uint32_t k)
  {
    const bool is_control = s.sign == KMD_SIGN_CONTROL;   // NO goes to the case file (aggregator.hpp:155-162)
    if (kff_control)
      (is_control ? kff_control : kff_cases)->write(s.kmer, kmer_words);
    else
    {
      char buf[1024];
      const size_t n = kmd_format_fasta(is_control ? n_control : n_case, &s, kmer_words, k, buf, sizeof buf);
      fwrite(buf, 1, n, is_control ? control : cases);
    }
    (is_control ? n_control : n_case)++;
  }
  // FASTA only: the records are formatted by `threads` threads (the running index of a record = its rank among the
  // records of its own file, aggregator.hpp:55-62) and written in order
  void write_all(const std::vector<kmd_survivor>& recs, uint32_t kmer_words, uint32_t k, size_t threads)
  {
    const size_t n = recs.size();
    threads = std::max<size_t>(1, std::min<size_t>(threads, 32));
    const size_t per = (n + threads - 1) / threads;
    // per chunk: how many control / case records precede it
    std::vector<uint64_t> ctl_before(threads + 1, 0), case_before(threads + 1, 0);
    for (size_t t = 0; t < threads; t++)
    {
      uint64_t c = 0, k2 = 0;
      for (size_t i = std::min(n, t * per); i < std::min(n, (t + 1) * per); i++) (recs[i].sign == KMD_SIGN_CONTROL ? c : k2)++;
      ctl_before[t + 1] = ctl_before[t] + c;
      case_before[t + 1] = case_before[t] + k2;
    }
    std::vector<std::string> ctl_text(threads), case_text(threads);
    auto work = [&](size_t t) {
      uint64_t ic = n_control + ctl_before[t], ik = n_case + case_before[t];
      char buf[1024];
      for (size_t i = std::min(n, t * per); i < std::min(n, (t + 1) * per); i++)
      {
        const bool is_control = recs[i].sign == KMD_SIGN_CONTROL;
        const size_t len = kmd_format_fasta(is_control ? ic++ : ik++, &recs[i], kmer_words, k, buf, sizeof buf);
        (is_control ? ctl_text[t] : case_text[t]).append(buf, len);
      }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (size_t t = 0; t < threads; t++)
    {
      fwrite(ctl_text[t].data(), 1, ctl_text[t].size(), control);
      fwrite(case_text[t].data(), 1, case_text[t].size(), cases);
    }
    n_control += ctl_before[threads];
    n_case += case_before[threads];
  }
  ~FastaPair() { if (control) fclose(control); if (cases) fclose(cases); }
};

size_t peak_rss_mb()
{
  rusage ru;
  getrusage(RUSAGE_SELF, &ru);
  return static_cast<size_t>(ru.ru_maxrss * 0.0009765625);
}

std::vector<kmd_survivor> fetch(kmd_context* ctx, bool significant)
{
  uint64_t n = 0;
  ck(ctx, significant ? kmd_significant(ctx, nullptr, 0, &n) : kmd_survivors(ctx, nullptr, 0, &n), "count records");
  std::vector<kmd_survivor> out(n);
  if (n) ck(ctx, significant ? kmd_significant(ctx, out.data(), n, &n) : kmd_survivors(ctx, out.data(), n, &n), "fetch records");
  return out;
}


// ---- do_pop (cmd/diff.hpp:168-224): eigenstrat inputs, PCA, per-k-mer GLM ------------------------------------

std::vector<double> fetch_counts_as_double(kmd_context* ctx, size_t n_samples, size_t count_bytes, uint64_t expect)
{
  uint64_t n = 0;
  ck(ctx, kmd_survivor_counts(ctx, nullptr, 0, &n), "kmd_survivor_counts");
  if (n != expect) throw KmdError("GpuError", "survivor counts out of step with the survivor records");
  std::vector<uint8_t> raw(n * n_samples * count_bytes);
  if (n) ck(ctx, kmd_survivor_counts(ctx, raw.data(), n, &n), "kmd_survivor_counts");
  std::vector<double> out(n * n_samples);
  for (size_t i = 0; i < out.size(); i++)
  {
    uint32_t v = 0;
    memcpy(&v, raw.data() + i * count_bytes, count_bytes);   // little-endian counts
    out[i] = static_cast<double>(v);
  }
  return out;
}

void do_pop(std::vector<std::unique_ptr<GpuWorker>>& workers, const DiffOptions& opt, const std::string& pop_dir,
            const std::vector<uint64_t>& total_controls, const std::vector<uint64_t>& total_cases)
{
  Timer pca_time;
  const uint32_t n = static_cast<uint32_t>(opt.nb_controls + opt.nb_cases);

  // write_gwas_info x2 + control.ind / case.ind, write_gwas_eigenstrat_total, write_parfile (src/popstrat.cpp:9-96)
  const std::vector<FofEntry> fof = parse_fof(opt.kmtricks_dir + "/kmtricks.fof");
  {
    std::ofstream par(pop_dir + "/parfile.txt");
    // std::map order of parfile_map (popstrat.hpp:28-37)
    par << "evaloutname: gwas_eigenstrat.eval\n" << "evecoutname: gwas_eigenstrat.evec\n" << "genotypename: gwas_eigenstratX.geno\n"
        << "indivname: gwas_eigenstratX.ind\n" << "numoutevec: 10\n" << "numoutlieriter: 0\n" << "snpname: gwas_eigenstratX.snp\n"
        << "usenorm: YES\n";
    for (const char* name : {"/gwas_infos.txt", "/gwas_eigenstratX.ind"})
    {
      std::ofstream out(pop_dir + name), coind(pop_dir + "/control.ind"), caind(pop_dir + "/case.ind");
      for (size_t i = 0; i < fof.size(); i++)
      {
        const bool control = i < opt.nb_controls;
        out << fof[i].id << "\tU\t" << (control ? "Control" : "Case") << "\n";
        (control ? coind : caind) << fof[i].id << "\tU\t" << (control ? "Control" : "Case") << "\n";
      }
    }
    std::ofstream tot(pop_dir + "/gwas_eigenstratX.total");
    for (uint64_t t : total_controls) tot << t << "\n";
    for (uint64_t t : total_cases) tot << t << "\n";
  }

  // gwas_eigenstratX.geno / .snp: the sampled k-mers of all GPUs in (partition, row) order
  // (EigGenoFile::push / EigSnpFile::push, popstrat.hpp:65-72, 92-96)
  struct Sampled { uint64_t row; size_t worker; uint64_t slot; };
  std::vector<Sampled> sampled;
  std::vector<std::vector<uint8_t>> geno(workers.size());
  for (size_t w = 0; w < workers.size(); w++)
  {
    kmd_context* ctx = workers[w]->ctx();
    uint64_t m = 0;
    ck(ctx, kmd_popstrat_sampled(ctx, &m), "kmd_popstrat_sampled");
    geno[w].resize(m * n);
    std::vector<uint64_t> rows(m);
    if (m) ck(ctx, kmd_popstrat_geno(ctx, geno[w].data(), rows.data(), m, &m), "kmd_popstrat_geno");
    for (uint64_t s = 0; s < m; s++) sampled.push_back({rows[s], w, s});
  }
  std::sort(sampled.begin(), sampled.end(), [](const Sampled& a, const Sampled& b) { return a.row < b.row; });
  {
    FILE* g = fopen((pop_dir + "/gwas_eigenstratX.geno").c_str(), "w");
    FILE* snp = fopen((pop_dir + "/gwas_eigenstratX.snp").c_str(), "w");
    if (!g || !snp) throw KmdError("IOError", "Unable to create eigenstrat inputs in " + pop_dir);
    std::string line(2ull * n + 1, '\t');
    line[2ull * n] = '\n';
    for (size_t i = 0; i < sampled.size(); i++)
    {
      const uint8_t* v = geno[sampled[i].worker].data() + sampled[i].slot * n;
      for (uint32_t j = 0; j < n; j++) line[2ull * j] = v[j] ? '1' : '0';
      fwrite(line.data(), 1, line.size(), g);
      fprintf(snp, "%zu\t1\t0.0\t0\n", i);
    }
    fclose(g); fclose(snp);
  }
  logf(L_DEBUG, "%zu k-mers sampled for the PCA", sampled.size());

  // smartpca (run_eigenstrat_smartpca, src/popstrat.cpp:98-137): the GPUs' integer co-carrier matrices summed exactly
  // on one device, one fp64 combination, one host eigensolve — the XTX of a one-GPU run, bit for bit
  std::vector<double> xtx(static_cast<size_t>(n) * n, 0.0);
  uint64_t used = 0;
  {
    std::vector<kmd_context*> ctxs;
    for (auto& w : workers) ctxs.push_back(w->ctx());
    ck(ctxs[0], kmd_popstrat_gram_local(ctxs.data(), static_cast<uint32_t>(ctxs.size()), xtx.data(), 0, &used), "kmd_popstrat_gram_local");
  }
  if (used == 0 || n < 2) throw KmdError("EigenStratError", "eigenstrat/smartpca failed.");
  kmd_smartpca_normalize(xtx.data(), n);
  const uint32_t numeigs = std::min<uint32_t>(n, 10);
  std::vector<double> evals(n), pcs(static_cast<size_t>(n) * numeigs);
  if (kmd_smartpca_pcs(xtx.data(), n, numeigs, evals.data(), pcs.data()) != KMD_OK)
    throw KmdError("EigenStratError", "eigenstrat/smartpca failed.");
  {
    // gwas_eigenstrat.eval ("%12.6f" per eigenvalue) and pcs.evec as evec2pca.perl prints it
    // (" " + " " for positive values + "%.04f" per component, one line per individual)
    FILE* ev = fopen((pop_dir + "/gwas_eigenstrat.eval").c_str(), "w");
    FILE* pc = fopen((pop_dir + "/pcs.evec").c_str(), "w");
    if (!ev || !pc) throw KmdError("IOError", "Unable to create PCA outputs in " + pop_dir);
    for (uint32_t i = 0; i < n; i++) fprintf(ev, "%12.6f\n", evals[i]);
    for (uint32_t i = 0; i < n; i++)
    {
      for (uint32_t j = 0; j < numeigs; j++)
      {
        const double v = pcs[static_cast<size_t>(i) * numeigs + j];
        fprintf(pc, v > 0 ? "  %.04f" : " %.04f", v);
      }
      fputc('\n', pc);
    }
    fclose(ev); fclose(pc);
  }
  logf(L_INFO, "PCA done. (%s).", pca_time.formatted().c_str());

  Timer pop_time;
  logf(L_INFO, "Apply population stratification correction...");
  // load_Z: 10 numbers per individual (src/popstrat.cpp:153-162); fewer than 10 samples leave the rest 0
  std::vector<double> Z(static_cast<size_t>(n) * 10, 0.0);
  for (uint32_t i = 0; i < n; i++)
    for (uint32_t j = 0; j < numeigs; j++) Z[static_cast<size_t>(i) * 10 + j] = pcs[static_cast<size_t>(i) * numeigs + j];
  std::vector<uint64_t> totals(total_controls);
  totals.insert(totals.end(), total_cases.begin(), total_cases.end());
  for (auto& w : workers)
  {
    ck(w->ctx(), kmd_popstrat_init(w->ctx(), Z.data(), totals.data(), nullptr), "kmd_popstrat_init");
    ck(w->ctx(), kmd_popstrat_apply(w->ctx()), "kmd_popstrat_apply");
    ck(w->ctx(), kmd_sync(w->ctx()), "kmd_sync");
  }
  logf(L_INFO, "Population correction done. (%s).", pop_time.formatted().c_str());
}

// Hide all but the first n devices from the CUDA driver (must run before the first CUDA call of the process).
void restrict_visible_devices(size_t n)
{
  if (n == 0) return;
  std::vector<std::string> ids;
  if (const char* cur = getenv("CUDA_VISIBLE_DEVICES"))
  {
    std::stringstream ss(cur);
    for (std::string item; std::getline(ss, item, ',');) if (!item.empty()) ids.push_back(item);
  }
  else
  {
    std::error_code ec;
    size_t count = 0;
    if (fs::is_directory("/proc/driver/nvidia/gpus", ec))
      for (auto& e : fs::directory_iterator("/proc/driver/nvidia/gpus", ec)) { (void)e; count++; }
    for (size_t i = 0; i < count; i++) ids.push_back(std::to_string(i));
  }
  if (ids.size() <= n) return;
  std::string v;
  for (size_t i = 0; i < n; i++) v += (i ? "," : "") + ids[i];
  setenv("CUDA_VISIBLE_DEVICES", v.c_str(), 1);
  logf(L_DEBUG, "CUDA_VISIBLE_DEVICES=%s (%zu of %zu devices)", v.c_str(), n, ids.size());
}

// ---- main_diff (cmd/diff.hpp:262-377) ---------------------------------------------------------------------

void main_diff(DiffOptions& opt)
{
  Timer whole;
  if (opt.pop_correction && !opt.covariates.empty())
    throw KmdError("OptionError", "--covariates is not supported: the reference's load_C never terminates on a non-empty file (src/popstrat.cpp:205-213).");
  if (opt.pop_correction && !opt.gender.empty())
    throw KmdError("IOError", "Unknown gender: the reference rejects every --gender file (src/popstrat.cpp:26-41).");
  if (opt.pop_correction && opt.npc > 10) throw KmdError("OptionError", "--n-pc must be <= 10 (pcs.evec holds 10 components).");

  // CUDA start-up (driver initialisation + one primary context per device, ~0.5 s) runs beside the host-side
  // start-up work below (run-dir config, N histogram files, P matrix headers, pinning the first staging buffers)
  struct Warmup
  {
    std::thread th;
    std::atomic<int> n_dev{0};
    void start(int want)
    {
      th = std::thread([this, want] {
        const int n = std::min(kmd_device_count(), want);
        std::vector<std::thread> pool;
        for (int d = 1; d < n; d++) pool.emplace_back([d] { kmd_warmup(d); });
        if (n > 0) kmd_warmup(0);
        for (auto& t : pool) t.join();
        n_dev = n;
      });
    }
    int wait() { if (th.joinable()) th.join(); return n_dev; }
    ~Warmup() { if (th.joinable()) th.join(); }
  } warmup;
  {
    // How many devices: --gpus, else one per 64 GiB of input — the host side (page cache -> pinned memory at ~40 GB/s on
    // a 16-32 core box) feeds one or two GPUs at most, and every further device costs start-up time.  The choice is made
    // BEFORE the CUDA driver is initialised and the other devices are hidden from it (CUDA_VISIBLE_DEVICES): the
    // driver's initialisation time grows with every visible device (measured: 0.6 s with 1, 5-7 s with 8 B200s visible).
    const auto mats = list_matrices(opt.kmtricks_dir);
    size_t n_units = mats.size();
    uint64_t bytes = 0;
    std::error_code ec;
    for (const std::string& mpath : mats) { const auto sz = fs::file_size(mpath, ec); if (!ec) bytes += sz; }
    if (n_units == 0)
    {
      n_units = 1u << 16;     // count-file mode: the partition count is read below
      if (fs::is_directory(opt.kmtricks_dir + "/counts"))
        for (auto& e : fs::recursive_directory_iterator(opt.kmtricks_dir + "/counts", ec))
          if (e.is_regular_file(ec)) bytes += static_cast<uint64_t>(e.file_size(ec));
    }
    const bool single = !opt.model_lib_path.empty();
    size_t want = opt.gpus > 0 ? static_cast<size_t>(opt.gpus) : std::max<uint64_t>(1, bytes >> 36);
    want = single ? 1 : std::min(want, n_units);
    restrict_visible_devices(want);
    if (opt.gpus == 0) opt.gpus = static_cast<int>(want);
    warmup.start(static_cast<int>(want));
  }

  const RunConfig cfg = get_kmtricks_config(opt.kmtricks_dir);
  if (cfg.kmer_size > 64) throw KmdError("OptionError", "k > 64 is not supported (the reference cannot write such k-mers either).");
  const uint32_t kw = (cfg.kmer_size + 31) / 32;
  const std::string part_dir = opt.output_directory + "/partitions";
  fs::create_directories(part_dir);

  // ---- an earlier run in the same output directory (cmd/diff.hpp:277-303) ----
  // With its spill files kept (--keep-tmp) and the same prefilter, the scan is skipped and only the correction
  // is redone (cmd/diff.hpp:340-349); when nothing changed and the FASTA files exist, nothing is redone.
  // The reference forgets total_kmers in options.bin and would correct with m = 0; ours stores it.  Spill files
  // of a --pop-correction run carry per-sample counts and are not resumed from: such a run is redone in full.
  const PrevOptions prev = load_opt(opt.output_directory + "/options.bin");
  bool resumed = false;          // the scan is skipped: survivors come from kept spill files
  bool resumed_pop = false;      // ... and so is the stratification: the GLM p-values are in p{i}_popstrat_uncorrected
  uint64_t resumed_total = 0;
  unsigned action = 0;
  if (prev.valid)
  {
    action = compare_opt(opt, prev);
    bool prev_1 = true, prev_2 = true;
    for (uint32_t p = 0; p < cfg.nb_partitions; p++)
    {
      prev_1 = prev_1 && fs::exists(part_dir + "/p" + std::to_string(p) + "_uncorrected");
      prev_2 = prev_2 && fs::exists(part_dir + "/p" + std::to_string(p) + "_popstrat_uncorrected");
    }
    const bool prev_f = fs::exists(opt.output_directory + "/control_kmers.fasta") && fs::exists(opt.output_directory + "/case_kmers.fasta");
    logf(L_DEBUG, "prev1 -> %d, prev2 -> %d, prevf -> %d, action -> %u", prev_1, prev_2, prev_f, action);
    // m of the correctors: options.bin of this program stores it; an output directory written by the reference does
    // not (it would correct with m = 0), but raw matrices give their row count away: (file size - 45) / R
    resumed_total = prev.total_kmers;
    if (resumed_total == 0)
    {
      const uint64_t R = 8ull * kw + (opt.nb_controls + opt.nb_cases) * opt.count_bytes;
      uint64_t rows = 0;
      bool known = true;
      const auto mats = list_matrices(opt.kmtricks_dir);
      for (const std::string& mpath : mats)
      {
        MatrixFile mf(mpath, R);
        if (mf.hdr.compressed) { known = false; break; }
        rows += mf.rows_if_raw;
      }
      if (known && !mats.empty()) resumed_total = rows;
    }
    const bool pop_kept = opt.pop_correction && prev.pop && prev_2 && !(action & 0b10);
    resumed = prev_1 && !(action & 0b1) && resumed_total > 0 && opt.model_lib_path.empty() && !opt.save_sk &&
              (!opt.pop_correction || pop_kept);
    resumed_pop = resumed && pop_kept;
    if (resumed && prev_f && action == 0)
    {
      logf(L_INFO, "Done in %s, Peak RSS -> %zu MB.", whole.formatted().c_str(), peak_rss_mb());
      return;
    }
  }

  // ---- do_diff ----
  Timer merge_time;
  if (!resumed) logf(L_INFO, "Process partitions");
  const std::vector<std::string> matrix_paths = resumed ? std::vector<std::string>{"<spill files>"} : list_matrices(opt.kmtricks_dir);
  // no matrices: the per-sample count files of every partition are merged on the fly (cmd/diff.hpp:88-95)
  const bool from_matrix = !matrix_paths.empty();
  if (from_matrix && !resumed) { std::ofstream touch(opt.kmtricks_dir + "/kmdiff-count.opt", std::ios::app); }   // the reference leaves this marker (cmd/diff.hpp:100)

  std::vector<uint64_t> total_controls, total_cases;
  get_total_kmer(opt.kmtricks_dir, opt.nb_controls, opt.nb_cases, cfg.abundance_min, total_controls, total_cases);
  uint64_t sum_c = 0, sum_k = 0;
  for (uint64_t t : total_controls) sum_c += t;
  for (uint64_t t : total_cases) sum_k += t;

  std::vector<Partition> parts;
  for (const std::string& p : resumed ? std::vector<std::string>{} : matrix_paths)
  {
    // partition index from the header (matrix_{p}.count[.lz4])
    MatrixFile mf(p, 8ull * kw + (opt.nb_controls + opt.nb_cases) * opt.count_bytes);
    Partition part{};
    part.path = p; part.index = mf.hdr.partition; part.bytes = static_cast<uint64_t>(fs::file_size(p)); part.raw = !mf.hdr.compressed;
    parts.push_back(part);
  }
  if (!from_matrix)
  {
    if (!opt.model_lib_path.empty())
      throw KmdError("OptionError", "--cmodel needs count matrices (`kmtricks pipeline --until merge --mode kmer:count:bin`): "
                                    "per-sample count files are merged on the GPU, user models run on the host.");
    for (uint32_t p = 0; p < cfg.nb_partitions; p++)
    {
      Partition part{};
      part.index = p;
      part.sample_paths = get_files_to_merge(opt.kmtricks_dir, p);
      if (part.sample_paths.size() != opt.nb_controls + opt.nb_cases)
        throw KmdError("ConfigError", "kmtricks.fof does not list --nb-controls + --nb-cases samples.");
      std::error_code ec;
      for (const std::string& sp : part.sample_paths) { const auto sz = fs::file_size(sp, ec); if (!ec) part.bytes += sz; }
      parts.push_back(part);
    }
  }
  uint32_t nb_partitions = cfg.nb_partitions;
  for (const Partition& p : parts) nb_partitions = std::max(nb_partitions, p.index + 1);
  std::sort(parts.begin(), parts.end(), [](const Partition& a, const Partition& b) { return a.index < b.index; });

  int n_dev = kmd_device_count();
  if (n_dev <= 0) throw KmdError("GpuError", "no CUDA device visible: kmdiff-b200 has no CPU fallback.");
  if (opt.gpus > 0) n_dev = std::min(n_dev, opt.gpus);
  warmup.wait();
  logf(L_DEBUG, "CUDA warm-up joined (%.3f s since start)", whole.seconds());
  const bool plugin = !opt.model_lib_path.empty();
  if (plugin)
  {
    if (opt.pop_correction) logf(L_WARN, "population stratification correction disabled with custom models.");
    opt.pop_correction = false;
    n_dev = 1;
  }
  if (!from_matrix && opt.pop_correction) n_dev = 1;   // Sampler draws follow the rows merged so far: one ordered stream
  n_dev = std::max(1, std::min<int>(n_dev, static_cast<int>(parts.size())));
  if (resumed) n_dev = 1;

  std::string pop_dir;
  if (opt.pop_correction)
  {
    pop_dir = opt.output_directory + "/popstrat";
    fs::create_directories(pop_dir);
  }
  if (opt.pop_correction && from_matrix)
  {
    // The Sampler draws once per row in visiting order (popstrat.hpp:121-135).  With `-t 1` the reference visits the
    // partitions one after the other, so the draw index of a row = rows of the partitions before it + its own index;
    // partitions are taken in index order here (the reference follows directory order, which is unspecified).
    const size_t R = 8ull * kw + (opt.nb_controls + opt.nb_cases) * opt.count_bytes;
    std::vector<uint64_t> rows(parts.size(), 0);
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex err_mu;
    auto count = [&] {
      for (size_t i; (i = next.fetch_add(1)) < parts.size();)
      {
        try
        {
          MatrixFile mf(parts[i].path, R);
          if (!mf.hdr.compressed) { rows[i] = mf.rows_if_raw; continue; }
          std::vector<uint8_t> buf(R * 16384);
          uint64_t bytes = 0;
          for (size_t got; (got = mf.payload->read(buf.data(), buf.size())) > 0;) bytes += got;
          rows[i] = bytes / R;
        }
        catch (...) { std::lock_guard<std::mutex> lk(err_mu); err = std::current_exception(); }
      }
    };
    std::vector<std::thread> pool;
    for (size_t t = 0; t < std::min(opt.nb_threads, parts.size()); t++) pool.emplace_back(count);
    for (auto& t : pool) t.join();
    if (err) std::rethrow_exception(err);
    uint64_t base = 0;
    for (size_t i = 0; i < parts.size(); i++) { parts[i].draw_base = base; base += rows[i]; }
  }

  // The workers own pinned staging buffers and device contexts whose orderly release (unpinning, cudaFree, context
  // teardown) costs more than a short run's scan; the process is about to end, so they are deliberately left to the OS
  // (main() leaves through _exit once every output file is closed).
  auto& workers = *new std::vector<std::unique_ptr<GpuWorker>>;
  // host threads (-t) are split over the devices: each GPU gets its own readers (page cache / LZ4 decode -> pinned).
  // Raw matrices are read in slices, so the reader count is not limited by the partition count.
  uint64_t input_bytes = 0;
  bool any_raw = false;
  for (const Partition& p : parts) { input_bytes += p.bytes; any_raw |= p.raw; }
  const size_t parts_per_gpu = (parts.size() + static_cast<size_t>(n_dev) - 1) / static_cast<size_t>(n_dev);
  const size_t reader_cap = any_raw ? size_t{32} : std::max<size_t>(1, parts_per_gpu);
  size_t readers_per_gpu = std::max<size_t>(1, std::min({size_t{32}, opt.nb_threads / static_cast<size_t>(n_dev), reader_cap}));
  if (const char* e = getenv("KMD_READERS")) readers_per_gpu = std::max<size_t>(1, static_cast<size_t>(atol(e)));   // tuning experiments
  // Pinning host memory costs ~0.3 s per GB when it is allocated and again when the process exits, so the staging
  // pool is kept small: 4 MiB buffers (a 76 us copy at PCIe rate, far above the launch overhead; measured best of
  // 4 / 16 / 32 MiB, profiles/r02d), readers + 8 of them; small inputs get smaller ones.
  size_t buf_bytes = static_cast<size_t>(std::min<uint64_t>(4ull << 20, std::max<uint64_t>(1ull << 20,
                         input_bytes / static_cast<uint64_t>(n_dev) / (2 * (readers_per_gpu + 2)))));
  if (const char* e = getenv("KMD_BUF_MB")) buf_bytes = std::max<size_t>(1, static_cast<size_t>(atol(e))) << 20;
  // one thread per device: CUDA context creation and the device allocations of kmd_create run side by side
  workers.resize(static_cast<size_t>(n_dev));
  {
    std::vector<std::thread> pool;
    std::exception_ptr err;
    std::mutex err_mu;
    for (int d = 0; d < n_dev; d++)
      pool.emplace_back([&, d] {
        try { workers[static_cast<size_t>(d)] = std::make_unique<GpuWorker>(d, opt, cfg, sum_c, sum_k, readers_per_gpu, buf_bytes); }
        catch (...) { std::lock_guard<std::mutex> lk(err_mu); err = std::current_exception(); }
      });
    for (auto& t : pool) t.join();
    if (err) std::rethrow_exception(err);
  }
  logf(L_DEBUG, "%d device context(s) ready (%.3f s), %zu reader thread(s) per device, %zu KiB staging buffers", n_dev,
       merge_time.seconds(), readers_per_gpu, buf_bytes >> 10);

  if (resumed)
  {
    // accumulators re-opened on the kept spill files (cmd/diff.hpp:342-348); after a stratified run the p-values to
    // correct are the GLM ones of p{i}_popstrat_uncorrected
    uint64_t tested = resumed_total;
    const uint32_t n_samples = static_cast<uint32_t>(opt.nb_controls + opt.nb_cases);
    const char* suffix = resumed_pop ? "_popstrat_uncorrected" : "_uncorrected";
    for (uint32_t p = 0; p < cfg.nb_partitions; p++)
    {
      std::vector<uint8_t> counts;
      const std::vector<kmd_survivor> recs = read_spill_file(part_dir + "/p" + std::to_string(p) + suffix, p, kw, n_samples,
                                                             static_cast<uint32_t>(opt.count_bytes), opt.keep_tmp ? &counts : nullptr);
      ck(workers[0]->ctx(), kmd_push_survivors2(workers[0]->ctx(), recs.data(), opt.keep_tmp ? counts.data() : nullptr, recs.size(), tested),
         "kmd_push_survivors2");
      tested = 0;
    }
    if (tested) ck(workers[0]->ctx(), kmd_push_survivors(workers[0]->ctx(), nullptr, 0, tested), "kmd_push_survivors");
  }
  else if (plugin)
  {
    switch (opt.count_bytes)
    {
      case 1: { PluginModel<kmdiff::maxc8> pm; pm.init(opt.model_lib_path, opt.model_config); run_plugin_model(pm, parts, opt, cfg, workers[0]->ctx()); break; }
      case 2: { PluginModel<kmdiff::maxc16> pm; pm.init(opt.model_lib_path, opt.model_config); run_plugin_model(pm, parts, opt, cfg, workers[0]->ctx()); break; }
      default: { PluginModel<kmdiff::maxc32> pm; pm.init(opt.model_lib_path, opt.model_config); run_plugin_model(pm, parts, opt, cfg, workers[0]->ctx()); break; }
    }
  }
  else
  {
    // size-balanced partition -> GPU assignment (largest first onto the least loaded device)
    std::vector<size_t> order(parts.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return parts[a].bytes != parts[b].bytes ? parts[a].bytes > parts[b].bytes : a < b; });
    std::vector<uint64_t> load(n_dev, 0);
    for (size_t i : order)
    {
      const int d = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
      workers[d]->partitions.push_back(parts[i]);
      load[d] += parts[i].bytes;
    }
    if (!from_matrix)   // index order per device: the running Sampler draw index follows it
      for (auto& w : workers)
        std::sort(w->partitions.begin(), w->partitions.end(), [](const Partition& a, const Partition& b) { return a.index < b.index; });
    std::vector<std::thread> pool;
    for (auto& w : workers) pool.emplace_back([&w] { w->run(); });
    for (auto& t : pool) t.join();
    for (auto& w : workers) if (w->error) std::rethrow_exception(w->error);
  }

  logf(L_DEBUG, "scan finished (%.3f s since start)", whole.seconds());
  uint64_t total_kmers = 0, nb_sign = 0, sign_controls = 0, sign_cases = 0;
  for (auto& w : workers)
  {
    kmd_counters c;
    ck(w->ctx(), kmd_counters_get(w->ctx(), &c), "kmd_counters_get");
    total_kmers += c.rows; nb_sign += c.survivors; sign_controls += c.n_control; sign_cases += c.n_case;
  }
  opt.total_kmers = total_kmers;
  if (!resumed)
  {
    logf(L_INFO, "Partitions processed (%s)", merge_time.formatted().c_str());
    logf(L_INFO, "%llu/%llu significant k-mers.", static_cast<unsigned long long>(nb_sign), static_cast<unsigned long long>(total_kmers));
    logf(L_INFO, "Before correction: %llu (control), %llu (case).", static_cast<unsigned long long>(sign_controls),
         static_cast<unsigned long long>(sign_cases));
  }

  dump_opt(opt, opt.output_directory + "/options.bin");
  const uint16_t n_samples = static_cast<uint16_t>(opt.nb_controls + opt.nb_cases);
  // spill files of the prefilter survivors with their counts (the reference's default record layout)
  auto spill = [&](const char* suffix) {
    std::vector<kmd_survivor> all;
    std::vector<double> counts;
    for (auto& w : workers)
    {
      auto s = fetch(w->ctx(), false);
      auto c = fetch_counts_as_double(w->ctx(), n_samples, opt.count_bytes, s.size());
      counts.insert(counts.end(), c.begin(), c.end());
      all.insert(all.end(), s.begin(), s.end());
    }
    write_spill_files(all, nb_partitions, kw, part_dir, suffix, counts, n_samples);
  };
  if (opt.keep_tmp && !resumed) spill("_uncorrected");

  if (opt.save_sk)
  {
    // --save-sk (cmd/diff.hpp:137-147, merge.hpp:270-278, 83-86): positive_kmer_matrix/ gets the run directory's
    // skeleton (copy_kdir, cmd/diff.hpp:52-64) and one matrix_{p}.count.lz4 per partition holding the rows that passed
    // the PREFILTER (written inside diff_observer::process, before any correction), count_slots = 4 in the header,
    // rows in partition order.  With --pop-correction the reference passes no writer (merge.hpp:286-289): skeleton only.
    const std::string root = opt.output_directory + "/positive_kmer_matrix";
    const std::string dir = root + "/matrices";
    fs::create_directories(dir);
    fs::create_directories(root + "/repartition_gatb");
    fs::create_directories(root + "/config_gatb");
    std::error_code ec;
    const auto copts = fs::copy_options::recursive | fs::copy_options::overwrite_existing;
    for (const char* sub : {"/config_gatb", "/repartition_gatb"})
      if (fs::exists(opt.kmtricks_dir + sub)) fs::copy(opt.kmtricks_dir + sub, root + sub, copts, ec);
    for (const char* file : {"/options.txt", "/kmtricks.fof"})
      if (fs::exists(opt.kmtricks_dir + file)) fs::copy(opt.kmtricks_dir + file, root, copts, ec);
    if (!opt.pop_correction)
    {
      const size_t cb = (opt.nb_controls + opt.nb_cases) * opt.count_bytes;
      std::vector<std::unique_ptr<MatrixWriter>> w(nb_partitions);
      for (const Partition& p : parts)
        w[p.index] = std::make_unique<MatrixWriter>(dir + "/matrix_" + std::to_string(p.index) + ".count.lz4", cfg.kmer_size, kw, 4,
                                                    static_cast<uint32_t>(opt.nb_controls + opt.nb_cases), p.index);
      for (auto& wk : workers)
      {
        const auto s = fetch(wk->ctx(), false);
        uint64_t n = 0;
        ck(wk->ctx(), kmd_survivor_counts(wk->ctx(), nullptr, 0, &n), "kmd_survivor_counts");
        if (n != s.size()) throw KmdError("GpuError", "survivor counts out of step with the survivor records");
        std::vector<uint8_t> raw(n * cb);
        if (n) ck(wk->ctx(), kmd_survivor_counts(wk->ctx(), raw.data(), n, &n), "kmd_survivor_counts");
        for (size_t i = 0; i < s.size(); i++)
        {
          if (s[i].partition >= nb_partitions || !w[s[i].partition]) continue;
          w[s[i].partition]->write(s[i].kmer, 8ull * kw);
          w[s[i].partition]->write(raw.data() + i * cb, cb);
        }
      }
    }
  }

  if (opt.pop_correction && !resumed_pop)
  {
    do_pop(workers, opt, pop_dir, total_controls, total_cases);
    if (opt.keep_tmp) spill("_popstrat_uncorrected");
  }

  // ---- do_correction ----
  Timer agg_time;
  logf(L_INFO, opt.correction == KMD_CORR_NOTHING ? "Aggregate partitions..." : "Aggregate partitions and apply significance correction...");
  const bool sorted = opt.correction == KMD_CORR_BENJAMINI || opt.correction == KMD_CORR_HOLM;
  std::vector<std::vector<kmd_survivor>> kept(workers.size());
  if (workers.size() > 1)
  {
    // m and (BENJAMINI / HOLM) one global ascending walk over all GPUs' survivors (aggregator.hpp:282-310), on the
    // devices: every context exports its counters and p-values, pulls the others' over the peer links, sorts the union
    // and keeps its own survivors below the first rejected p-value (kmd_exchange_local)
    std::vector<kmd_context*> ctxs;
    for (auto& w : workers) ctxs.push_back(w->ctx());
    std::vector<kmd_exchange_decision> dec(ctxs.size());
    ck(ctxs[0], kmd_exchange_local(ctxs.data(), static_cast<uint32_t>(ctxs.size()), opt.correction, opt.threshold, dec.data()),
       "kmd_exchange_local");
    if (dec[0].total_rows != total_kmers) throw KmdError("GpuError", "exchange: row totals out of step");
    for (size_t i = 0; i < workers.size(); i++) kept[i] = fetch(workers[i]->ctx(), true);
  }
  else
  {
    for (size_t i = 0; i < workers.size(); i++)
    {
      ck(workers[i]->ctx(), kmd_correct(workers[i]->ctx(), opt.correction, opt.threshold, total_kmers, nullptr, nullptr), "kmd_correct");
      kept[i] = fetch(workers[i]->ctx(), true);
    }
  }
  logf(L_DEBUG, "correction done, records fetched (%.3f s since start)", whole.seconds());
  // k-way merge of the per-GPU outputs: ascending p for BH/Holm, (partition, row) otherwise
  std::vector<kmd_survivor> final_records;
  for (auto& k : kept) final_records.insert(final_records.end(), k.begin(), k.end());
  if (workers.size() > 1)
  {
    if (sorted)
      std::stable_sort(final_records.begin(), final_records.end(), [](const kmd_survivor& a, const kmd_survivor& b) {
        return a.p_value != b.p_value ? a.p_value < b.p_value : a.row < b.row; });
    else
      std::stable_sort(final_records.begin(), final_records.end(), [](const kmd_survivor& a, const kmd_survivor& b) { return a.row < b.row; });
  }
  {
    FastaPair out(opt.output_directory, opt.kff, cfg.kmer_size);
    if (opt.kff || final_records.size() < 4096)
      for (const kmd_survivor& s : final_records) out.write(s, kw, cfg.kmer_size);
    else
      out.write_all(final_records, kw, cfg.kmer_size, opt.nb_threads);
    logf(L_INFO, "Partitions aggregated (%s)", agg_time.formatted().c_str());
    logf(L_INFO, "Significant k-mers: %llu (control), %llu (case).", static_cast<unsigned long long>(out.n_control),
         static_cast<unsigned long long>(out.n_case));
  }
  if (!opt.keep_tmp)
  {
    std::error_code ec;
    for (uint32_t p = 0; p < nb_partitions; p++)
    {
      fs::remove(part_dir + "/p" + std::to_string(p) + "_uncorrected", ec);
      fs::remove(part_dir + "/p" + std::to_string(p) + "_popstrat_uncorrected", ec);
    }
  }
  logf(L_INFO, "Done in %s, Peak RSS -> %zu MB.", whole.formatted().c_str(), peak_rss_mb());
}

// Hidden developer commands used by tests/test_host_cli.py to exercise the host-side format code without a GPU:
//   dev-cat FILE SKIP [lz4]  decoded payload of a kmtricks file (SKIP header bytes; LZ4 frame iff its compressed flag is set or "lz4" is given)
//   dev-pack IN OUT          IN wrapped into one LZ4 frame of stored blocks
//   dev-config DIR NC NK     run-dir config and per-sample totals
//   dev-synth OUT ...        a synthetic kmtricks run directory (bench inputs)
int dev_command(int argc, char** argv)
{
  const std::string cmd = argv[1];
  if (cmd == "dev-cat" && (argc == 4 || argc == 5))
  {
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    std::vector<uint8_t> head(static_cast<size_t>(atol(argv[3])));
    if (!head.empty() && fread(head.data(), 1, head.size(), f) != head.size()) return 2;
    const bool compressed = argc == 5 || (head.size() >= 13 && head[12] != 0);
    PayloadReader pr(f, compressed);
    std::vector<uint8_t> buf(1 << 20);
    for (size_t n; (n = pr.read(buf.data(), buf.size())) > 0;) fwrite(buf.data(), 1, n, stdout);
    fclose(f);
    return 0;
  }
  if (cmd == "dev-pack" && argc == 4)
  {
    FILE* in = fopen(argv[2], "rb");
    FILE* out = fopen(argv[3], "wb");
    if (!in || !out) return 2;
    Lz4FrameWriter w(out);
    std::vector<uint8_t> buf(100000);
    for (size_t n; (n = fread(buf.data(), 1, buf.size(), in)) > 0;) w.write(buf.data(), n);
    w.close();
    fclose(in); fclose(out);
    return 0;
  }
  if (cmd == "dev-kff" && argc >= 4)
  {
    // dev-kff OUT K KMER...: the given k-mers (ACGT strings) as a KFF file
    const uint32_t k = static_cast<uint32_t>(atol(argv[3]));
    KffWriter w(argv[2], k);
    for (int i = 4; i < argc; i++)
    {
      uint64_t words[2] = {0, 0};
      for (uint32_t b = 0; b < k; b++)
      {
        const char c = argv[i][b];
        const uint64_t v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : 3;      // kmtricks kmer.hpp:38
        const uint32_t bit = 2 * (k - 1 - b);
        words[bit / 64] |= v << (bit % 64);
      }
      w.write(words, (k + 31) / 32);
    }
    w.close();
    return 0;
  }
  if (cmd == "dev-synth" && argc >= 12)
  {
    // dev-synth OUT NC NK K COUNT_BYTES TOTAL_ROWS SEED PPM N_ROWS PARTITIONS matrix|counts [THREADS]
    // A kmtricks run directory holding rows [0, N_ROWS) of the seeded synthetic cohort (kmd_synth_*), as raw count
    // matrices or as per-sample sorted count files: the input of bench.py's reference arm and of the CLI timings,
    // written by a process of its own so that the timed reference never shares a process with this library.
    const std::string out = argv[2];
    kmd_synth sy{};
    sy.nb_controls = static_cast<uint32_t>(atol(argv[3])); sy.nb_cases = static_cast<uint32_t>(atol(argv[4]));
    sy.kmer_size = static_cast<uint32_t>(atol(argv[5])); sy.kmer_words = (sy.kmer_size + 31) / 32;
    sy.count_bytes = static_cast<uint32_t>(atol(argv[6])); sy.total_rows = strtoull(argv[7], nullptr, 10);
    sy.seed = strtoull(argv[8], nullptr, 10); sy.enriched_ppm = static_cast<uint32_t>(atol(argv[9]));
    const uint64_t n_rows = strtoull(argv[10], nullptr, 10);
    const uint32_t parts = static_cast<uint32_t>(std::max(1l, atol(argv[11])));
    const bool count_files = argc > 12 && !strcmp(argv[12], "counts");
    const size_t threads = argc > 13 ? static_cast<size_t>(std::max(1l, atol(argv[13]))) : std::max(1u, std::thread::hardware_concurrency());
    const uint32_t n = sy.nb_controls + sy.nb_cases, kw = sy.kmer_words, B = sy.count_bytes;
    const size_t R = 8ull * kw + static_cast<size_t>(n) * B;
    fs::create_directories(out + "/matrices");
    fs::create_directories(out + "/histograms");
    std::vector<std::string> ids;
    for (uint32_t i = 0; i < sy.nb_controls; i++) ids.push_back("C" + std::to_string(i));
    for (uint32_t i = 0; i < sy.nb_cases; i++) ids.push_back("K" + std::to_string(i));
    {
      std::ofstream fof(out + "/kmtricks.fof"), opt(out + "/options.txt");
      for (const std::string& id : ids) fof << id << " : /dev/null/" << id << ".fa\n";
      opt << "kmer_size=" << sy.kmer_size << ", c_ab_min=1,\n";
    }
    fs::remove(out + "/kmdiff-count.opt");
    std::vector<uint64_t> totals(n);
    kmd_synth_totals(&sy, totals.data());
    for (uint32_t j = 0; j < n; j++)
    {
      // km::HistFileHeader + two empty histograms over [1, 255] (hist_file.hpp:31-45, 129-144); only `total` is read
      FILE* f = fopen((out + "/histograms/" + ids[j] + ".hist").c_str(), "wb");
      if (!f) return 2;
#pragma pack(push, 1)
      struct { uint64_t km_magic; uint32_t version; uint8_t compressed; uint64_t magic; uint32_t k, id;
               uint64_t lower, upper, uniq, total, oob_ln, oob_lu, oob_un, oob_uu; } h{kKmMagic, 0, 0, kHistMagic, sy.kmer_size, j, 1, 255, 0, totals[j], 0, 0, 0, 0};
#pragma pack(pop)
      fwrite(&h, sizeof h, 1, f);
      const std::vector<uint64_t> zeros(2 * 255, 0);
      fwrite(zeros.data(), 8, zeros.size(), f);
      fclose(f);
    }
    const uint64_t per = (n_rows + parts - 1) / parts;
    std::atomic<uint32_t> next{0};
    std::atomic<bool> failed{false};
    auto work = [&] {
      const uint64_t chunk = std::max<uint64_t>(1, (8u << 20) / R);
      std::vector<uint8_t> rows(chunk * R);
      for (uint32_t p; (p = next.fetch_add(1)) < parts;)
      {
        fs::create_directories(out + "/counts/partition_" + std::to_string(p));
        const uint64_t lo = std::min(n_rows, p * per), hi = std::min(n_rows, (p + 1) * per);
        std::vector<FILE*> files;
        if (count_files)
        {
          for (uint32_t j = 0; j < n; j++)
          {
            FILE* f = fopen((out + "/counts/partition_" + std::to_string(p) + "/" + ids[j] + ".kmer.lz4").c_str(), "wb");
            if (!f) { failed = true; return; }
            const KmerHeader h{kKmMagic, 0, 0, kKmerMagic, sy.kmer_size, kw, B, j, p};
            fwrite(&h, sizeof h, 1, f);
            files.push_back(f);
          }
        }
        else
        {
          FILE* f = fopen((out + "/matrices/matrix_" + std::to_string(p) + ".count").c_str(), "wb");
          if (!f) { failed = true; return; }
          const MatrixHeader h{kKmMagic, 0, 0, kMatrixMagic, sy.kmer_size, kw, B, n, 0, p};
          fwrite(&h, sizeof h, 1, f);
          files.push_back(f);
        }
        for (uint64_t r = lo; r < hi; r += chunk)
        {
          const uint64_t m = std::min(chunk, hi - r);
          if (kmd_synth_fill_host(&sy, rows.data(), m, r, 1) != KMD_OK) { failed = true; break; }
          if (!count_files) { fwrite(rows.data(), R, m, files[0]); continue; }
          // KmerWriter::write (kmer_file.hpp:94-101): one (k-mer, count) record per sample that carries the k-mer
          for (uint64_t i = 0; i < m; i++)
          {
            const uint8_t* row = rows.data() + i * R;
            for (uint32_t j = 0; j < n; j++)
            {
              const uint8_t* c = row + 8ull * kw + static_cast<size_t>(j) * B;
              bool nz = false;
              for (uint32_t b = 0; b < B; b++) nz |= c[b] != 0;
              if (!nz) continue;
              fwrite(row, 1, 8ull * kw, files[j]);
              fwrite(c, 1, B, files[j]);
            }
          }
        }
        for (FILE* f : files) fclose(f);
      }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < std::min<size_t>(threads, parts); t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return failed ? 2 : 0;
  }
  if (cmd == "dev-lz4bench" && argc >= 4)
  {
    // dev-lz4bench THREADS FILE...: decode the payloads of the given kmtricks files (45-byte matrix or 41-byte k-mer
    // header, then one LZ4 frame) with THREADS host threads, one file per thread at a time, and print the rates
    const size_t threads = static_cast<size_t>(std::max(1l, atol(argv[2])));
    std::vector<std::string> files(argv + 3, argv + argc);
    std::atomic<size_t> next{0};
    std::atomic<uint64_t> out_bytes{0}, in_bytes{0};
    const Timer t;
    auto work = [&] {
      std::vector<uint8_t> buf(4u << 20);
      for (size_t i; (i = next.fetch_add(1)) < files.size();)
      {
        FILE* f = fopen(files[i].c_str(), "rb");
        if (!f) continue;
        uint8_t head[45];
        if (fread(head, 1, 21, f) != 21) { fclose(f); continue; }
        uint64_t magic; memcpy(&magic, head + 13, 8);
        const size_t rest = magic == kMatrixMagic ? 45 - 21 : 41 - 21;
        if (fread(head + 21, 1, rest, f) != rest) { fclose(f); continue; }
        PayloadReader pr(f, head[12] != 0);
        uint64_t n = 0;
        for (size_t got; (got = pr.read(buf.data(), buf.size())) > 0;) n += got;
        out_bytes += n;
        in_bytes += static_cast<uint64_t>(fs::file_size(files[i]));
        fclose(f);
      }
    };
    std::vector<std::thread> pool;
    for (size_t i = 1; i < threads; i++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    const double s = t.seconds();
    printf("{\"threads\": %zu, \"files\": %zu, \"compressed_bytes\": %llu, \"decoded_bytes\": %llu, \"seconds\": %.4f, "
           "\"decoded_gbs\": %.3f, \"compressed_gbs\": %.3f}\n", threads, files.size(), static_cast<unsigned long long>(in_bytes.load()),
           static_cast<unsigned long long>(out_bytes.load()), s, out_bytes.load() / s / 1e9, in_bytes.load() / s / 1e9);
    return 0;
  }
  if (cmd == "dev-config" && argc == 5)
  {
    const RunConfig c = get_kmtricks_config(argv[2]);
    std::vector<uint64_t> tc, tk;
    get_total_kmer(argv[2], static_cast<size_t>(atol(argv[3])), static_cast<size_t>(atol(argv[4])), c.abundance_min, tc, tk);
    printf("kmer_size=%u abundance_min=%u nb_partitions=%u\n", c.kmer_size, c.abundance_min, c.nb_partitions);
    for (uint64_t t : tc) printf("control %llu\n", static_cast<unsigned long long>(t));
    for (uint64_t t : tk) printf("case %llu\n", static_cast<unsigned long long>(t));
    for (const std::string& m : list_matrices(argv[2])) printf("matrix %s\n", m.c_str());
    return 0;
  }
  return 2;
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc >= 2 && !strncmp(argv[1], "dev-", 4))
  {
    try { return dev_command(argc, argv); }
    catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 3; }
  }
  const Timer process_time;
  DiffOptions opt = parse_cli(argc, argv);
  g_level = opt.verbosity == "debug" ? L_DEBUG : opt.verbosity == "info" ? L_INFO : opt.verbosity == "warning" ? L_WARN : L_ERROR;
  if (opt.correction == KMD_CORR_BENJAMINI || opt.correction == KMD_CORR_HOLM)
    logf(L_WARN, "-c/--correction %s: all significants k-mers will live in memory.", opt.correction == KMD_CORR_BENJAMINI ? "benjamini" : "holm");
  if (opt.in_memory) logf(L_WARN, "-m/--in-memory: all significants k-mers will live in memory.");
  try
  {
    main_diff(opt);
  }
  catch (const KmdError& e)
  {
    logf(L_ERROR, "%s - %s", e.name.c_str(), e.what());
  }
  catch (const std::exception& e)
  {
    logf(L_ERROR, "%s", e.what());
  }
  // like the reference: errors are logged, the exit code stays 0 (main.cc:93-102).  Every output file has been closed
  // by its owner; _exit skips the CUDA runtime's exit-time teardown (~0.1-0.2 s for nothing).
  logf(L_DEBUG, "leaving main (%.3f s since process start)", process_time.seconds());
  fflush(nullptr);
  _exit(0);
}
