// C-ABI of the B200-native `kmdiff diff` hot path (include/kmdiff_b200.h) and the kernels behind it
// that are not the streaming scan (kmd_scan.cuh) or the sort (kmd_sort.cuh).
//
// There is no CPU fallback in this file: every compute entry point needs a context, and a context
// needs an sm_100 device.
#include "../../include/kmdiff_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "kmd_eig.hpp"
#include "kmd_grm.cuh"
#include "kmd_lz4.cuh"
#include "kmd_math.cuh"
#include "kmd_merge.cuh"
#include "kmd_union.cuh"
#include "kmd_popstrat.cuh"
#include "kmd_scan.cuh"
#include "kmd_sort.cuh"
#include "kmd_synth.cuh"

namespace {

using namespace kmd;

thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------

struct SurvivorArrays
{
  uint64_t* kmer0 = nullptr;
  uint64_t* kmer1 = nullptr;
  double* p = nullptr;
  int32_t* sign = nullptr;
  uint32_t* part = nullptr;
  double* mc = nullptr;
  double* kk = nullptr;
  uint64_t* row = nullptr;
};

struct DeviceCounters
{
  uint32_t cand_count;      // candidates of the batch in flight
  uint32_t surv_count;      // survivors appended so far
  uint32_t n_control;
  uint32_t n_case;
  unsigned long long cand_total;
  uint32_t overflow;        // 1: candidate buffer overflowed, 2: survivor buffer overflowed
  uint32_t first_reject;    // scratch for first_reject_kernel
  uint32_t keep_control;
  uint32_t keep_case;
  uint32_t keep_total;
  uint32_t pad_;
};

// --pop-correction state: the Sampler's output (presence vectors of the sampled k-mers), the counts of
// the prefilter survivors (KmerSign::m_counts_ratio under WITH_POPSTRAT) and the GLM features.
struct PopState
{
  bool enabled = false;
  kmd_popstrat_config cfg{};
  uint32_t state0 = 1;
  bool draw_override = false;              // kmd_popstrat_set_draw_index: draws follow their own counter
  uint64_t draw_next = 0;
  uint64_t cap = 0;                        // max sampled k-mers
  unsigned long long* d_total = nullptr;   // sampled so far
  unsigned long long* d_done = nullptr;    // slots with a presence vector
  uint32_t* d_overflow = nullptr;
  uint64_t* d_slot_rows = nullptr;
  uint8_t* d_geno = nullptr;               // [cap][N] 0/1
  uint32_t* d_carriers = nullptr;          // [cap]
  uint8_t* d_surv_counts = nullptr;        // [max_survivors][N * count_bytes]
  // GLM
  bool glm_ready = false;
  uint32_t null_p = 0, alt_p = 0;
  double* d_null_feat = nullptr;           // [N][null_p]
  double* d_alt_feat = nullptr;            // [N][alt_p]
  double* d_y = nullptr;
  double* d_totals = nullptr;
  double null_likelihood = 0;
  std::vector<double> null_model;
  float gram_ms = 0;                       // device time of the last grm_umma_kernel launch
  // grow-only scratch of kmd_popstrat_gram (bucketing, transposed presence matrix, integer Gram matrices):
  // the presence matrix alone is ~0.5 GB at 1e6 sampled k-mers, too much to allocate and free per call
  void* gram_scratch[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t gram_scratch_cap[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// per-sample count-file input (kmd_merge_*): grow-only device buffers of the merge tree
struct MergeState
{
  uint8_t* raw = nullptr; size_t raw_cap = 0;
  uint64_t* klo[2] = {nullptr, nullptr}; size_t klo_cap[2] = {0, 0};
  uint64_t* khi[2] = {nullptr, nullptr}; size_t khi_cap[2] = {0, 0};
  uint64_t* pay[2] = {nullptr, nullptr}; size_t pay_cap[2] = {0, 0};
  uint32_t* flag = nullptr; size_t flag_cap = 0;
  uint32_t* tile = nullptr; size_t tile_cap = 0;
  uint32_t* head_pos = nullptr; size_t head_cap = 0;
  uint8_t* rows = nullptr; size_t rows_cap = 0;
  uint64_t* stream_off = nullptr; size_t off_cap = 0;
  uint32_t* ab_min = nullptr; size_t ab_cap = 0;
  uint32_t* plan = nullptr; size_t plan_cap = 0;
  uint32_t* split = nullptr; size_t split_cap = 0;   // merge-path split of every tile of the pass in flight
  // device LZ4 decode of compressed count files (kmd_merge_scan_lz4_host): the frames and their per-stream tables
  uint8_t* lz_in = nullptr; size_t lz_in_cap = 0;
  uint64_t* lz_tab = nullptr; size_t lz_tab_cap = 0;   // [5][n]: in_off, in_size, out_off, out_cap, out_size
  uint32_t* lz_status = nullptr; size_t lz_status_cap = 0;
  cudaEvent_t lz_ev[2] = {nullptr, nullptr};
  float last_lz4_ms = 0;
  // kmd_merge_prefetch: the NEXT call's host blobs, already on their way to the device on the copy stream
  // (two slots: the chunk about to be merged and the one after it)
  struct Prefetch
  {
    uint8_t* buf = nullptr; size_t cap = 0;
    std::vector<const void*> ptrs;
    std::vector<uint64_t> bytes;
    int kind = -1;                  // -1 free, 0 records back to back, 1 frames at 16-byte aligned offsets
    cudaEvent_t done = nullptr;
  } pf[2];
  uint32_t* total = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // unpack | merge passes | row numbering + dense rows
  bool timed = false;
  uint32_t last_passes = 0;
  uint64_t last_records = 0;
  // the union tree (kmd_union.cuh): tile descriptors, look-back words (+ the ticket), the scatter's splits, flags
  uni::TileDesc* udesc = nullptr; size_t udesc_cap = 0;
  unsigned long long* ustatus = nullptr; size_t ustatus_cap = 0;
  uint32_t* usplit = nullptr; size_t usplit_cap = 0;
  uint32_t* uorder = nullptr; size_t uorder_cap = 0;
  uint32_t* uflags = nullptr;
  int algo = -1;            // 0 union tree (pairwise tree when a record is not solid), 1 pairwise tree; -1 = ask KMD_MERGE_PAIRWISE
  int last_algo = 1;        // what the last call ran
  bool asm_attr_set[2] = {false, false};
  bool knob_raw_leaf = true;     // KMD_UNION_RAW_LEAF=0: unpack the keys first even with 4-byte counts
  uint32_t knob_rows = 0, knob_smem = 0, knob_scatter_tile = 1024;   // KMD_UNION_ROWS / _SMEM_KB / _SCATTER_TILE
  int last_rows_mode = 0;   // union tree: 1 rows assembled in shared memory, 2 scattered into zeroed rows
  uint64_t last_keys_in = 0, last_keys_out = 0;   // keys read / written by the levels of the last union tree
};

// multi-GPU correction (kmd_exchange_*): the union of the ranks' p-values, sorted here, and the decision
struct ExchangeDecision
{
  unsigned long long total_rows, total_survivors, max_rank_survivors, n_accepted;
  double p_cut;
  uint32_t first;        // scratch of first_reject_dev_kernel
  uint32_t overflow;     // some rank had more survivors than `cap`
};

struct ExchangeState
{
  uint64_t* keys[2] = {nullptr, nullptr};
  size_t keys_cap = 0;
  ExchangeDecision* d_dec = nullptr;
  ExchangeDecision* h_dec = nullptr;   // pinned
  bool pending = false;                // kmd_exchange_correct enqueued, result not read yet
  bool by_p = false;
  // kmd_exchange_local (one process, several devices): this context's gather buffer and "my slot is written" event
  uint64_t* gather = nullptr;
  size_t gather_words = 0;
  cudaEvent_t ev_export = nullptr;
};

}  // namespace

struct kmd_context
{
  kmd_config cfg{};
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  std::string err;

  // model
  kmd::ModelConst model{};      // lf points to device memory
  double* d_lf = nullptr;
  double lr_cut = 0;            // LR*(threshold) on the host
  // scan plan
  kmd::ScanParams plan{};
  uint32_t consumer_warps = 16;
  size_t scan_smem = 0;
  bool direct = false;          // stream rows with plain global loads instead of the TMA ring

  // device state
  uint64_t cap = 0;
  DeviceCounters* d_cnt = nullptr;
  DeviceCounters* h_cnt = nullptr;  // pinned
  uint32_t* d_cand = nullptr;
  SurvivorArrays surv;
  uint64_t rows_tested = 0;
  uint64_t launches = 0;          // kernels launched on behalf of this context
  cudaEvent_t ev_k0 = nullptr;    // around the last lrt_scan kernel launch
  cudaEvent_t ev_k1 = nullptr;
  bool scan_timed = false;

  // sort / correction state
  uint64_t* d_keys[2] = {nullptr, nullptr};
  uint32_t* d_vals[2] = {nullptr, nullptr};
  uint32_t* d_hist = nullptr;
  size_t hist_entries = 0;
  double* d_p_sorted = nullptr;
  kmd_survivor* d_out = nullptr;
  uint64_t out_cap = 0;
  uint32_t* d_order = nullptr;   // final order (survivor slots), valid after correct/keep_below
  uint32_t n_final = 0;
  int order_state = 0;           // 0 none, 1 row order over all survivors, 2 p order over all survivors
  bool final_valid = false;
  bool final_sorted = false;     // d_order is in output order (row order, or ascending p); else sorted lazily at fetch
  bool final_by_p = false;       // output order of an unsorted d_order: ascending p (ties by row) instead of row order
  uint64_t max_row_key = 0;      // upper bound of every survivor's row index: bounds the radix passes of a row sort

  // host staging for kmd_scan_host / kmd_model_rows_host
  uint8_t* d_stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_scan[2] = {nullptr, nullptr};
  int stage_next = 0;             // staging buffer the next chunk goes to (persists across calls: copy i+1 overlaps scan i)
  // kmd_scan_host_async tickets: event t % kTicketRing is recorded on the copy stream after the last H2D copy of ticket t
  static constexpr uint64_t kTicketRing = 256;
  cudaEvent_t ev_ticket[kTicketRing] = {};
  uint64_t next_ticket = 1;

  PopState pop;
  MergeState mrg;
  ExchangeState xch;
};

namespace {

void pop_free(kmd_context* ctx);   // kmd_popstrat_api.inl
void merge_free(kmd_context* ctx); // kmd_merge_api.inl
void exchange_free(kmd_context* ctx); // kmd_exchange_api.inl
int pop_reset(kmd_context* ctx);

int fail(kmd_context* ctx, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

#define KMD_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? KMD_ERR_NOMEM : KMD_ERR_CUDA,           \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
  } while (0)

struct DeviceBuf
{
  void* p = nullptr;
  ~DeviceBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

uint32_t env_u32(const char* name, uint32_t dflt)
{
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return static_cast<uint32_t>(strtoul(v, nullptr, 10));
}

// ------------------------------------------------------------------------------------------------
// host-side model pieces
// ------------------------------------------------------------------------------------------------

// LogFactorialTable (reference src/log_factorial_table.cpp:5-22): entry k is log(k) + log(k-1) +
// ... + log(2), accumulated in exactly that order.  log(j) is evaluated once per j instead of once
// per (k, j) pair; the additions and their order are the reference's, so the table is bit-identical
// to the reference's on the same libm.
void lf_table_fill(double* table, uint64_t size)
{
  if (size == 0) return;
  std::vector<double> logs(size);
  for (uint64_t j = 2; j < size; j++) logs[j] = std::log(static_cast<double>(j));
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned nt = size < 4096 ? 1 : std::min(hw, 16u);
  auto work = [&](unsigned t) {
    // interleave so every thread gets short and long chains
    for (uint64_t k = t; k < size; k += nt)
    {
      double acc = 0;
      for (uint64_t j = k; j > 1; j--) acc += logs[j];
      table[k] = acc;
    }
  };
  if (nt == 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; t++) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
}

double host_igamc_half(double x)
{
  static const double lg = kmd::lngamma_small(0.5);
  return kmd::igamc_half(x, lg);
}

// Largest LR known to give p > thr (bisection on the monotone igamc); +inf if nothing can pass,
// -1 if everything passes.
double lr_cut_for(double thr)
{
  if (!(thr >= 0)) return std::numeric_limits<double>::infinity();  // negative or NaN: no survivor
  if (thr >= 1) return -1;
  double lo = 0, hi = 1;
  while (host_igamc_half(hi) > thr)
  {
    lo = hi;
    hi *= 2;
    if (hi > 4096) return std::numeric_limits<double>::infinity();
  }
  for (int i = 0; i < 200; i++)
  {
    double mid = 0.5 * (lo + hi);
    if (host_igamc_half(mid) > thr) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------
// scan plan: tile / stage / lanes-per-row geometry
// ------------------------------------------------------------------------------------------------

uint32_t conflict_degree(uint32_t R, uint32_t kmer_bytes, uint32_t g, bool fast, uint32_t B)
{
  uint32_t worst = 1;
  if (fast)
  {
    for (uint32_t half = 0; half < 2; half++)
    {
      uint32_t cnt[16] = {0};
      uint64_t seen[16][16];
      for (uint32_t i = 0; i < 16; i++)
      {
        uint32_t lane = half * 16 + i;
        uint64_t addr = static_cast<uint64_t>(lane / g) * R + kmer_bytes + 8 * (lane % g);
        uint32_t bank = (addr / 8) % 16;
        bool dup = false;
        for (uint32_t k = 0; k < cnt[bank]; k++) dup |= seen[bank][k] == addr;
        if (!dup) seen[bank][cnt[bank]++] = addr;
      }
      for (uint32_t b = 0; b < 16; b++) worst = std::max(worst, cnt[b]);
    }
  }
  else
  {
    uint32_t cnt[32] = {0};
    uint64_t seen[32][32];
    for (uint32_t lane = 0; lane < 32; lane++)
    {
      uint64_t addr = (static_cast<uint64_t>(lane / g) * R + kmer_bytes + B * (lane % g)) / 4;
      uint32_t bank = addr % 32;
      bool dup = false;
      for (uint32_t k = 0; k < cnt[bank]; k++) dup |= seen[bank][k] == addr;
      if (!dup) seen[bank][cnt[bank]++] = addr;
    }
    for (uint32_t b = 0; b < 32; b++) worst = std::max(worst, cnt[b]);
  }
  return worst;
}

int make_plan(kmd_context* ctx)
{
  const kmd_config& c = ctx->cfg;
  ScanParams& P = ctx->plan;
  const uint32_t N = c.nb_controls + c.nb_cases;
  const uint32_t B = c.count_bytes;
  P.row_bytes = 8 * c.kmer_words + N * B;
  P.kmer_bytes = 8 * c.kmer_words;
  P.nb_controls = c.nb_controls;
  P.nb_cases = c.nb_cases;
  P.fast = (static_cast<uint64_t>(N) * B) % 8 == 0;
  P.words_total = P.fast ? N * B / 8 : 0;
  P.words_ctrl = P.fast ? c.nb_controls * B / 8 : 0;
  P.straddle_elems = P.fast ? (c.nb_controls * B % 8) / B : 0;

  // consumer geometry: `groups` groups of `gw` warps; a group sums one tile at a time.  Presets by row size, from the
  // sweep of profiles/r02j (scan GB/s on 10+10 / 50+50 / 100+100 / 200+200 cohorts: more consumer warps pay for narrow
  // and for ~1 KB rows, the 3 x 8 geometry for the sizes between, 16 lanes per row for very wide rows); every knob can
  // still be overridden from the environment.
  struct Preset { uint32_t max_row_bytes, groups, warps, tile_bytes, lanes; };
  static const Preset presets[] = {{128, 3, 10, 32768, 0}, {512, 3, 8, 24576, 0}, {1024, 5, 6, 32768, 0}, {0xFFFFFFFFu, 3, 8, 24576, 16}};
  const Preset* preset = presets;
  while (P.row_bytes > preset->max_row_bytes) preset++;
  uint32_t groups = std::max<uint32_t>(1, env_u32("KMD_SCAN_GROUPS", preset->groups));
  uint32_t gw = std::max<uint32_t>(1, env_u32("KMD_SCAN_WARPS", preset->warps));
  if (groups * gw > kMaxConsumerWarps) return fail(ctx, KMD_ERR_ARG, "KMD_SCAN_GROUPS x KMD_SCAN_WARPS exceeds 31 consumer warps");
  const uint32_t warps = groups * gw;
  ctx->consumer_warps = warps;
  P.consumer_warps = warps;
  P.groups = groups;
  P.group_warps = gw;

  const uint32_t ring_budget = 200 * 1024;
  const uint32_t target_tile = env_u32("KMD_SCAN_TILE_BYTES", preset->tile_bytes);
  const uint32_t units = P.fast ? P.words_total : N;

  // lanes per row: cheapest estimated issue cost among the choices whose full pass fits the ring
  uint32_t best_g = 32;
  double best_cost = 1e30;
  for (uint32_t g = 1; g <= 32; g *= 2)
  {
    const uint64_t pass_rows = static_cast<uint64_t>(gw) * (32 / g);
    if (pass_rows * P.row_bytes * std::max<uint32_t>(2, groups) > ring_budget && g < 32) continue;
    const uint32_t conf = conflict_degree(P.row_bytes, P.kmer_bytes, g, P.fast, B);
    uint32_t lg = 0;
    while ((1u << lg) < g) lg++;
    const double per_pass = std::ceil(static_cast<double>(units) / g) * (2.0 + 2.0 * conf) + 8.0 * lg + 40.0;
    const double cost = per_pass * g / 32.0;
    if (cost < best_cost) { best_cost = cost; best_g = g; }
  }
  P.lanes_per_row = env_u32("KMD_SCAN_G", preset->lanes ? preset->lanes : best_g);
  if (P.lanes_per_row == 0 || (P.lanes_per_row & (P.lanes_per_row - 1)) || P.lanes_per_row > 32)
    return fail(ctx, KMD_ERR_ARG, "KMD_SCAN_G must be a power of two in [1,32]");

  // rows per tile: whole passes of one consumer group, rounded to 16 rows so that tile bytes are a
  // multiple of 16 for any R
  uint32_t pass_rows = gw * (32 / P.lanes_per_row);
  pass_rows = (pass_rows + 15) / 16 * 16;
  uint32_t passes = std::max<uint32_t>(1, target_tile / (pass_rows * P.row_bytes));
  P.tile_rows = env_u32("KMD_SCAN_TILE_ROWS", pass_rows * passes);
  if (P.tile_rows % 16) return fail(ctx, KMD_ERR_ARG, "tile rows must be a multiple of 16");
  const uint64_t tile_bytes = static_cast<uint64_t>(P.tile_rows) * P.row_bytes;
  P.stage_bytes = static_cast<uint32_t>((tile_bytes + 127) / 128 * 128);
  uint32_t stages = tile_bytes ? static_cast<uint32_t>(ring_budget / P.stage_bytes) : 0;
  stages = std::min<uint32_t>(stages, kMaxStages);
  stages = std::min<uint32_t>(stages, std::max<uint32_t>(1, env_u32("KMD_SCAN_STAGES", kMaxStages)));
  // A consumer group must be the only reader of a ring slot in consecutive phases (a group that ran a
  // whole ring ahead of another would alias the mbarrier parity), so the ring holds a multiple of
  // `groups` stages; with fewer stages than groups the rows are streamed with plain loads instead.
  if (stages < groups)
  {
    P.groups = 1;  // one group of all consumer warps
    P.group_warps = warps;
  }
  else
    stages -= stages % groups;
  P.stages = stages;
  const char* mode = getenv("KMD_SCAN_MODE");
  ctx->direct = stages < 2 || tile_bytes >= (1u << 20) || (mode && !strcmp(mode, "direct"));
  ctx->scan_smem = ctx->direct ? 0 : kBarrierBytes + static_cast<size_t>(P.stages) * P.stage_bytes;

  // screen
  const double Sc = static_cast<double>(c.sum_controls), Sk = static_cast<double>(c.sum_cases);
  const double S = static_cast<double>(c.sum_controls + c.sum_cases);
  P.screen_off = (c.sum_controls == 0 || c.sum_cases == 0 || ctx->lr_cut < 0) ? 1 : 0;
  if (env_u32("KMD_SCREEN_OFF", 0)) P.screen_off = 1;
  P.frac_ctrl = static_cast<float>(Sc / S);
  P.frac_case = static_cast<float>(Sk / S);
  if (std::isinf(ctx->lr_cut))
    P.lr2_cut = std::numeric_limits<float>::infinity();
  else
  {
    const double cut2 = ctx->lr_cut * (1.0 - 1e-9) / 0.6931471805599453;
    P.lr2_cut = std::nextafterf(static_cast<float>(cut2), -std::numeric_limits<float>::infinity());
  }
  return KMD_OK;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------

// Same work as lrt_scan_kernel without the shared-memory ring: rows are read with plain global
// loads.  Used for rows too large for the ring, and as the A/B baseline (KMD_SCAN_MODE=direct).
template <int B>
__global__ void __launch_bounds__(kMaxScanThreads) lrt_scan_direct_kernel(const ScanParams P, uint32_t warps)
{
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const uint64_t n_chunks = (P.n_rows + P.tile_rows - 1) / P.tile_rows;
  for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
  {
    const uint64_t first = chunk * P.tile_rows;
    const uint32_t n = static_cast<uint32_t>(min(static_cast<uint64_t>(P.tile_rows), P.n_rows - first));
    GmemLoader ld{P.rows + first * P.row_bytes};
    scan_rows<B>(P, ld, n, static_cast<uint32_t>(first), warp, warps, lane);
  }
}

struct ExactArgs
{
  const uint8_t* rows;
  uint64_t row_base;
  uint32_t partition;
  double threshold;
  ModelConst model;
  ScanParams geom;        // lanes_per_row forced to 1
  const uint32_t* cand;
  DeviceCounters* cnt;
  SurvivorArrays out;
  uint32_t cap;
  uint32_t kmer_words;
  uint8_t* counts_out;    // --pop-correction: the survivor's N counts (merge.hpp:161-163), else null
  uint32_t counts_bytes;  // N * count_bytes
};

// diff_observer::process on the candidate rows: exact fp64 LRT in reference order, the `<=`
// prefilter, sign counting (NO counted as case) and the KmerSign fields (merge.hpp:68-103).
template <int B>
__global__ void __launch_bounds__(256) lrt_exact_kernel(const ExactArgs A)
{
  const uint32_t n_cand = min(A.cnt->cand_count, A.geom.cand_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicAdd(&A.cnt->cand_total, static_cast<unsigned long long>(A.cnt->cand_count));
    if (A.cnt->cand_count > A.geom.cand_cap) atomicOr(&A.cnt->overflow, 1u);
  }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x)
  {
    const uint32_t r = A.cand[i];
    const uint8_t* row = A.rows + static_cast<uint64_t>(r) * A.geom.row_bytes;
    GmemLoader ld{row};
    uint64_t sc, sk;
    row_partial<B>(A.geom, ld, 0, 0, 1, sc, sk);
    const double* lf = A.model.lf;
    const RowStat st = lrt_exact(A.model, sc, sk, [lf](uint64_t k) { return __ldg(lf + k); });
    if (st.p <= A.threshold)
    {
      const uint32_t pos = atomicAdd(&A.cnt->surv_count, 1u);
      if (st.sign == KMD_SIGN_CONTROL) atomicAdd(&A.cnt->n_control, 1u); else atomicAdd(&A.cnt->n_case, 1u);
      if (pos < A.cap)
      {
        uint64_t k0 = 0, k1 = 0;
        if ((A.geom.row_bytes & 7u) == 0)
        {
          k0 = ld.ld64(0);
          if (A.kmer_words > 1) k1 = ld.ld64(8);
        }
        else
        {
          for (int b = 0; b < 8; b++) k0 |= static_cast<uint64_t>(ld.ld8(b)) << (8 * b);
          if (A.kmer_words > 1)
            for (int b = 0; b < 8; b++) k1 |= static_cast<uint64_t>(ld.ld8(8 + b)) << (8 * b);
        }
        A.out.kmer0[pos] = k0;
        A.out.kmer1[pos] = k1;
        A.out.p[pos] = st.p;
        A.out.sign[pos] = st.sign;
        A.out.part[pos] = A.partition;
        A.out.mc[pos] = st.mean_control;
        A.out.kk[pos] = st.mean_case;
        A.out.row[pos] = A.row_base + r;
        if (A.counts_out)
        {
          uint8_t* dst = A.counts_out + static_cast<size_t>(pos) * A.counts_bytes;
          for (uint32_t b = 0; b < A.counts_bytes; b++) dst[b] = ld.ld8(A.geom.kmer_bytes + b);
        }
      }
      else
        atomicOr(&A.cnt->overflow, 2u);
    }
  }
}

// IModel::process for every row (imodel.hpp:37): one thread per row, the log-factorial table
// staged in shared memory (each row does four look-ups).
template <int B>
__global__ void __launch_bounds__(256) lrt_model_rows_kernel(const uint8_t* rows, uint64_t n_rows, ScanParams geom,
                                                            ModelConst model, uint32_t lf_smem, double* p,
                                                            int32_t* sign, double* mc, double* kk)
{
  extern __shared__ double lf_s[];
  for (uint32_t i = threadIdx.x; i < lf_smem; i += blockDim.x) lf_s[i] = model.lf[i];
  __syncthreads();
  const double* lf = model.lf;
  for (uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; r < n_rows;
       r += static_cast<uint64_t>(gridDim.x) * blockDim.x)
  {
    GmemLoader ld{rows + r * geom.row_bytes};
    uint64_t sc, sk;
    row_partial<B>(geom, ld, 0, 0, 1, sc, sk);
    const RowStat st = lrt_exact(model, sc, sk, [lf, lf_smem](uint64_t k) {
      return k < lf_smem ? lf_s[k] : __ldg(lf + k);
    });
    p[r] = st.p;
    sign[r] = st.sign;
    mc[r] = st.mean_control;
    kk[r] = st.mean_case;
  }
}

__global__ void iota_keys_kernel(uint32_t n, const uint64_t* row, const double* p, int by_p, const uint32_t* order_in,
                                 uint64_t* keys, uint32_t* vals)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = order_in ? order_in[i] : i;
  keys[i] = by_p ? static_cast<uint64_t>(__double_as_longlong(p[slot])) : row[slot];
  vals[i] = slot;
}

__global__ void set_u32_kernel(uint32_t* dst, uint32_t v) { *dst = v; }

__global__ void gather_p_kernel(uint32_t n, const uint32_t* order, const double* p, double* out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = p[order[i]];
}

__global__ void count_prefix_kernel(uint32_t n, const uint32_t* order, const int32_t* sign, DeviceCounters* cnt)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (sign[order[i]] == KMD_SIGN_CONTROL) atomicAdd(&cnt->keep_control, 1u); else atomicAdd(&cnt->keep_case, 1u);
}

__global__ void gather_records_kernel(uint32_t n, const uint32_t* order, SurvivorArrays s, kmd_survivor* out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = order[i];
  kmd_survivor r;
  r.kmer[0] = s.kmer0[slot];
  r.kmer[1] = s.kmer1[slot];
  r.p_value = s.p[slot];
  r.sign = s.sign[slot];
  r.partition = s.part[slot];
  r.mean_control = s.mc[slot];
  r.mean_case = s.kk[slot];
  r.row = s.row[slot];
  out[i] = r;
}

// Stateless correctors (aggregator::worker, aggregator.hpp:137-163): every survivor is tested on its own, so the
// kept slots are appended in arrival order (warp-aggregated atomics) and only ordered when they are fetched.
// cut_ptr != null: the cut is read from device memory (multi-GPU exchange, no host round trip).
__global__ void __launch_bounds__(256) keep_unordered_kernel(DeviceCounters* cnt, uint32_t cap, const double* p, double cut,
                                                             const double* cut_ptr, const int32_t* sign, uint32_t* out_order)
{
  if (cut_ptr) cut = *cut_ptr;
  const uint32_t n = min(cnt->surv_count, cap);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (n + stride - 1) / stride;
  uint32_t n_ctl = 0, n_case = 0;
  for (uint32_t r = 0, i = blockIdx.x * blockDim.x + threadIdx.x; r < rounds; r++, i += stride)
  {
    const bool keep = i < n && p[i] < cut;
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
    if (!ballot) continue;
    uint32_t pos = 0;
    if (lane == 0) pos = atomicAdd(&cnt->keep_total, __popc(ballot));
    pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(ballot & ((1u << lane) - 1));
    if (keep)
    {
      out_order[pos] = i;
      if (sign[i] == KMD_SIGN_CONTROL) n_ctl++; else n_case++;
    }
  }
  for (uint32_t o = 16; o > 0; o >>= 1)
  {
    n_ctl += __shfl_xor_sync(0xFFFFFFFFu, n_ctl, o);
    n_case += __shfl_xor_sync(0xFFFFFFFFu, n_case, o);
  }
  if (lane == 0)
  {
    if (n_ctl) atomicAdd(&cnt->keep_control, n_ctl);
    if (n_case) atomicAdd(&cnt->keep_case, n_case);
  }
}

// FileAccumulator::push for records that arrive from the host (plugin models, resumed spill files)
// counts_in / counts_out (may be null): the records' per-sample counts, counts_bytes each, stored at the record's slot
__global__ void push_records_kernel(uint32_t n, const kmd_survivor* in, SurvivorArrays s, DeviceCounters* cnt, uint32_t cap,
                                    const uint8_t* counts_in, uint8_t* counts_out, uint32_t counts_bytes)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const kmd_survivor r = in[i];
  const uint32_t pos = atomicAdd(&cnt->surv_count, 1u);
  if (pos < cap && counts_out)
  {
    uint8_t* dst = counts_out + static_cast<size_t>(pos) * counts_bytes;
    if (counts_in)
      for (uint32_t b = 0; b < counts_bytes; b++) dst[b] = counts_in[static_cast<size_t>(i) * counts_bytes + b];
    else
      for (uint32_t b = 0; b < counts_bytes; b++) dst[b] = 0;
  }
  if (r.sign == KMD_SIGN_CONTROL) atomicAdd(&cnt->n_control, 1u); else atomicAdd(&cnt->n_case, 1u);
  if (pos >= cap) { atomicOr(&cnt->overflow, 2u); return; }
  s.kmer0[pos] = r.kmer[0];
  s.kmer1[pos] = r.kmer[1];
  s.p[pos] = r.p_value;
  s.sign[pos] = r.sign;
  s.part[pos] = r.partition;
  s.mc[pos] = r.mean_control;
  s.kk[pos] = r.mean_case;
  s.row[pos] = r.row;
}

__global__ void synth_kernel(synth::Params p, const synth::Tables* t, uint8_t* dst, uint64_t n_rows, uint64_t row_base)
{
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n_rows;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    synth::make_row(p, *t, row_base + i, dst + i * p.row_bytes);
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------

template <int B>
int launch_scan(kmd_context* ctx, const uint8_t* d_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition)
{
  ScanParams P = ctx->plan;
  P.rows = d_rows;
  P.n_rows = n_rows;
  P.cand_count = &ctx->d_cnt->cand_count;
  P.cand = ctx->d_cand;
  P.cand_cap = static_cast<uint32_t>(ctx->cap);
  KMD_CUDA(ctx, cudaMemsetAsync(&ctx->d_cnt->cand_count, 0, sizeof(uint32_t), ctx->stream));
  if (ctx->direct)
  {
    P.n_tiles = 0;
    const uint64_t chunks = (n_rows + P.tile_rows - 1) / P.tile_rows;
    const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>(chunks, static_cast<uint64_t>(ctx->sm_count) * 2));
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    lrt_scan_direct_kernel<B><<<grid, ctx->consumer_warps * 32, 0, ctx->stream>>>(P, ctx->consumer_warps);
    cudaEventRecord(ctx->ev_k1, ctx->stream);
  }
  else
  {
    P.n_tiles = static_cast<uint32_t>(n_rows / P.tile_rows);
    const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(P.n_tiles, ctx->sm_count));
    cudaEventRecord(ctx->ev_k0, ctx->stream);
    lrt_scan_kernel<B><<<grid, (ctx->consumer_warps + 1) * 32, ctx->scan_smem, ctx->stream>>>(P);
    cudaEventRecord(ctx->ev_k1, ctx->stream);
  }
  ctx->scan_timed = true;
  ctx->launches += 2;
  KMD_CUDA(ctx, cudaGetLastError());

  ExactArgs A;
  A.rows = d_rows;
  A.row_base = row_base;
  A.partition = partition;
  A.threshold = ctx->cfg.threshold;
  A.model = ctx->model;
  A.geom = P;
  A.geom.lanes_per_row = 1;
  A.cand = ctx->d_cand;
  A.cnt = ctx->d_cnt;
  A.out = ctx->surv;
  A.cap = static_cast<uint32_t>(ctx->cap);
  A.kmer_words = ctx->cfg.kmer_words;
  A.counts_out = ctx->pop.d_surv_counts;   // --pop-correction or KMD_FLAG_KEEP_COUNTS, else null
  A.counts_bytes = (ctx->cfg.nb_controls + ctx->cfg.nb_cases) * ctx->cfg.count_bytes;
  lrt_exact_kernel<B><<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(A);
  KMD_CUDA(ctx, cudaGetLastError());
  if (ctx->pop.enabled)
  {
    // diff_observer_strat::process (merge.hpp:143-174): the Sampler sees every row, draw index = global row index
    PopState& ps = ctx->pop;
    pop::SamplerArgs S;
    S.state0 = ps.state0;
    S.v = ps.cfg.kmer_pca;
    S.v_scaled = ps.cfg.kmer_pca * 2147483646.0;
    S.first_row = row_base;
    S.first_draw = ps.draw_override ? ps.draw_next : row_base;
    ps.draw_next += n_rows;
    S.n_rows = n_rows;
    S.rows_per_thread = 64;
    S.total = ps.d_total;
    S.out_rows = ps.d_slot_rows;
    S.cap = ps.cap;
    const uint64_t threads = (n_rows + S.rows_per_thread - 1) / S.rows_per_thread;
    pop::sampler_kernel<<<static_cast<uint32_t>((threads + 255) / 256), 256, 0, ctx->stream>>>(S);
    pop::PresenceArgs Q;
    Q.rows = d_rows;
    Q.row_base = row_base;
    Q.row_bytes = P.row_bytes;
    Q.kmer_bytes = P.kmer_bytes;
    Q.n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
    Q.count_bytes = ctx->cfg.count_bytes;
    Q.total = ps.d_total;
    Q.done = ps.d_done;
    Q.slot_rows = ps.d_slot_rows;
    Q.geno = ps.d_geno;
    Q.carriers = ps.d_carriers;
    Q.cap = ps.cap;
    pop::presence_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(Q);
    pop::presence_done_kernel<<<1, 1, 0, ctx->stream>>>(ps.d_total, ps.d_done, ps.cap, ps.d_overflow);
    ctx->launches += 3;
    KMD_CUDA(ctx, cudaGetLastError());
  }
  return KMD_OK;
}

int scan_device_impl(kmd_context* ctx, const uint8_t* d_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition)
{
  // the kernels index rows of one launch with 32 bits
  const uint64_t max_batch = (1ull << 31);
  uint64_t done = 0;
  while (done < n_rows)
  {
    uint64_t n = std::min(max_batch, n_rows - done);
    if (n < n_rows - done) n -= n % 16;  // keep the next sub-batch 16-byte aligned for any row size
    const uint8_t* base = d_rows + done * ctx->plan.row_bytes;
    int rc;
    switch (ctx->cfg.count_bytes)
    {
      case 1: rc = launch_scan<1>(ctx, base, n, row_base + done, partition); break;
      case 2: rc = launch_scan<2>(ctx, base, n, row_base + done, partition); break;
      default: rc = launch_scan<4>(ctx, base, n, row_base + done, partition); break;
    }
    if (rc) return rc;
    done += n;
  }
  ctx->rows_tested += n_rows;
  ctx->max_row_key = std::max(ctx->max_row_key, row_base + n_rows);
  ctx->order_state = 0;
  ctx->final_valid = false;
  return KMD_OK;
}

int ensure_stage(kmd_context* ctx)
{
  if (ctx->d_stage[0]) return KMD_OK;
  size_t want = static_cast<size_t>(env_u32("KMD_STAGE_MB", 64)) << 20;
  const size_t unit = static_cast<size_t>(ctx->plan.row_bytes) * 16;
  want = std::max(unit, want / unit * unit);
  for (int b = 0; b < 2; b++)
  {
    KMD_CUDA(ctx, cudaMalloc(&ctx->d_stage[b], want));
    KMD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[b], cudaEventDisableTiming));
    KMD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_scan[b], cudaEventDisableTiming));
  }
  ctx->stage_bytes = want;
  return KMD_OK;
}

int fetch_counters(kmd_context* ctx)
{
  KMD_CUDA(ctx, cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->h_cnt->overflow & 4u)
    return fail(ctx, KMD_ERR_STATE, "internal error: a record's k-mer was not among the merged rows (kmd_union.cuh)");
  if (ctx->h_cnt->overflow)
    return fail(ctx, KMD_ERR_OVERFLOW, "%s buffer overflow: raise kmd_config.max_survivors (now %llu)",
                (ctx->h_cnt->overflow & 1) ? "candidate" : "survivor", static_cast<unsigned long long>(ctx->cap));
  return KMD_OK;
}

// stable LSD radix sort of (key[, slot]) pairs over the low 8*n_passes key bits (higher bits must be equal in all
// keys), ping-ponging between the two buffers; *result = index of the buffer that holds the sorted output.
// vals may be null (keys only).
int radix_sort_buffers(kmd_context* ctx, uint64_t* const keys[2], uint32_t* const vals[2], uint32_t n, int n_passes, int* result)
{
  *result = 0;
  if (n < 2 || n_passes <= 0) return KMD_OK;
  const uint32_t n_ctas = (n + kSortCtaKeys - 1) / kSortCtaKeys;
  const size_t need = static_cast<size_t>(n_ctas) * 256;
  if (need > ctx->hist_entries)
  {
    if (ctx->d_hist) cudaFree(ctx->d_hist);
    ctx->d_hist = nullptr;
    ctx->hist_entries = 0;
    KMD_CUDA(ctx, cudaMalloc(&ctx->d_hist, need * sizeof(uint32_t)));
    ctx->hist_entries = need;
  }
  int cur = 0;
  for (int pass = 0; pass < n_passes; pass++)
  {
    const int shift = 8 * pass;
    sort_hist_kernel<<<n_ctas, kSortWarps * 32, 0, ctx->stream>>>(keys[cur], n, shift, n_ctas, ctx->d_hist);
    scan_exclusive_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_hist, static_cast<uint32_t>(need), nullptr);
    sort_scatter_kernel<<<n_ctas, kSortWarps * 32, 0, ctx->stream>>>(keys[cur], vals ? vals[cur] : nullptr, keys[cur ^ 1],
                                                                     vals ? vals[cur ^ 1] : nullptr, n, shift, n_ctas, ctx->d_hist);
    ctx->launches += 3;
    cur ^= 1;
  }
  KMD_CUDA(ctx, cudaGetLastError());
  *result = cur;
  return KMD_OK;
}

// the context's survivor (key, slot) pairs in d_keys[0]/d_vals[0]; the sorted slots end up in d_vals[0]
int radix_sort(kmd_context* ctx, uint32_t n, int n_passes = 8)
{
  int cur = 0;
  int rc = radix_sort_buffers(ctx, ctx->d_keys, ctx->d_vals, n, n_passes, &cur);
  if (rc) return rc;
  if (cur == 1)
    KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_vals[0], ctx->d_vals[1], n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return KMD_OK;
}

int row_key_passes(const kmd_context* ctx)
{
  int bits = 0;
  for (uint64_t v = ctx->max_row_key; v; v >>= 1) bits++;
  return std::max(1, (bits + 7) / 8);
}

inline uint32_t blocks_for(uint32_t n) { return std::max<uint32_t>(1, (n + 255) / 256); }

// order survivors by row (state 1) or by (p, row) (state 2); the order ends up in d_vals[0]
int ensure_order(kmd_context* ctx, int want)
{
  int rc = fetch_counters(ctx);
  if (rc) return rc;
  const uint32_t n = ctx->h_cnt->surv_count;
  if (ctx->order_state == want || n == 0) { ctx->order_state = want; return KMD_OK; }
  if (ctx->order_state == 0)
  {
    iota_keys_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(n, ctx->surv.row, ctx->surv.p, 0, nullptr, ctx->d_keys[0], ctx->d_vals[0]);
    ctx->launches++;
    if ((rc = radix_sort(ctx, n, row_key_passes(ctx)))) return rc;
    ctx->order_state = 1;
  }
  if (want == 2 && ctx->order_state == 1)
  {
    // stable sort by p of the row-ordered slots: ties in p keep row order
    KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_vals[1], ctx->d_vals[0], n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    iota_keys_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(n, ctx->surv.row, ctx->surv.p, 1, ctx->d_vals[1], ctx->d_keys[0], ctx->d_vals[0]);
    ctx->launches++;
    if ((rc = radix_sort(ctx, n))) return rc;
    ctx->order_state = 2;
  }
  else if (want == 1 && ctx->order_state == 2)
  {
    iota_keys_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(n, ctx->surv.row, ctx->surv.p, 0, nullptr, ctx->d_keys[0], ctx->d_vals[0]);
    ctx->launches++;
    if ((rc = radix_sort(ctx, n, row_key_passes(ctx)))) return rc;
    ctx->order_state = 1;
  }
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

int ensure_out(kmd_context* ctx, uint64_t n)
{
  if (n <= ctx->out_cap) return KMD_OK;
  if (ctx->d_out) cudaFree(ctx->d_out);
  ctx->d_out = nullptr;
  ctx->out_cap = 0;
  KMD_CUDA(ctx, cudaMalloc(&ctx->d_out, n * sizeof(kmd_survivor)));
  ctx->out_cap = n;
  return KMD_OK;
}

int download_records(kmd_context* ctx, const uint32_t* d_order, uint32_t n, kmd_survivor* out, uint64_t cap, uint64_t* n_out)
{
  if (n_out) *n_out = n;
  if (n == 0 || !out) return KMD_OK;
  if (cap < n) return fail(ctx, KMD_ERR_ARG, "output capacity %llu < %u records", static_cast<unsigned long long>(cap), n);
  int rc = ensure_out(ctx, n);
  if (rc) return rc;
  gather_records_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(n, d_order, ctx->surv, ctx->d_out);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out, n * sizeof(kmd_survivor), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KMD_OK;
}

// keep survivors (visited in the current order) with p < cut; stable
// survivors with p < cut, kept in arrival order; kmd_significant puts them in ascending p (ties by row) when asked
int keep_below_impl(kmd_context* ctx, double cut, uint64_t* n_control, uint64_t* n_case)
{
  KMD_CUDA(ctx, cudaMemsetAsync(&ctx->d_cnt->keep_control, 0, 3 * sizeof(uint32_t), ctx->stream));
  keep_unordered_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_cnt, static_cast<uint32_t>(ctx->cap), ctx->surv.p, cut,
                                                                     nullptr, ctx->surv.sign, ctx->d_order);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  int rc = fetch_counters(ctx);
  if (rc) return rc;
  ctx->n_final = ctx->h_cnt->keep_total;
  ctx->final_valid = true;
  ctx->final_sorted = false;
  ctx->final_by_p = true;
  if (n_control) *n_control = ctx->h_cnt->keep_control;
  if (n_case) *n_case = ctx->h_cnt->keep_case;
  return KMD_OK;
}

synth::Tables* build_synth_tables()
{
  static synth::Tables t;
  static bool done = false;
  if (!done)
  {
    for (int c = 0; c < synth::kClasses; c++)
    {
      const double lam = synth::kLambda[c];
      double pmf = std::exp(-lam), cdf = 0;
      for (int i = 0; i < synth::kCdfLen; i++)
      {
        cdf += pmf;
        double v = std::floor(cdf * 4294967296.0);
        t.cdf[c][i] = v >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(v);
        pmf = pmf * lam / (i + 1);
      }
      t.cdf[c][synth::kCdfLen - 1] = 0xFFFFFFFFu;
    }
    done = true;
  }
  return &t;
}

int synth_params(const kmd_synth* s, synth::Params* p)
{
  if (!s || !s->nb_controls || !s->nb_cases || (s->kmer_words != 1 && s->kmer_words != 2) ||
      (s->count_bytes != 1 && s->count_bytes != 2 && s->count_bytes != 4) || !s->total_rows)
    return KMD_ERR_ARG;
  p->seed = s->seed;
  p->nb_controls = s->nb_controls;
  p->nb_cases = s->nb_cases;
  p->kmer_words = s->kmer_words;
  p->count_bytes = s->count_bytes;
  p->enriched_ppm = s->enriched_ppm;
  p->row_bytes = 8 * s->kmer_words + (s->nb_controls + s->nb_cases) * s->count_bytes;
  int bits = 2 * static_cast<int>(s->kmer_size) - (s->kmer_words == 2 ? 64 : 0);
  bits = std::max(1, std::min(bits, 62));
  p->kmer_stride = std::max<uint64_t>(1, (1ull << bits) / s->total_rows);
  return KMD_OK;
}

}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================

extern "C" {

int kmd_abi_version(void) { return KMD_ABI_VERSION; }

int kmd_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* kmd_last_error(const kmd_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void kmd_destroy(kmd_context* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  cudaFree(ctx->d_lf);
  cudaFree(ctx->d_cnt);
  if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
  cudaFree(ctx->d_cand);
  cudaFree(ctx->surv.kmer0); cudaFree(ctx->surv.kmer1); cudaFree(ctx->surv.p); cudaFree(ctx->surv.sign);
  cudaFree(ctx->surv.part); cudaFree(ctx->surv.mc); cudaFree(ctx->surv.kk); cudaFree(ctx->surv.row);
  for (int b = 0; b < 2; b++)
  {
    cudaFree(ctx->d_keys[b]); cudaFree(ctx->d_vals[b]); cudaFree(ctx->d_stage[b]);
    if (ctx->ev_copy[b]) cudaEventDestroy(ctx->ev_copy[b]);
    if (ctx->ev_scan[b]) cudaEventDestroy(ctx->ev_scan[b]);
  }
  cudaFree(ctx->d_hist); cudaFree(ctx->d_p_sorted); cudaFree(ctx->d_out); cudaFree(ctx->d_order);
  for (cudaEvent_t e : ctx->ev_ticket) if (e) cudaEventDestroy(e);
  pop_free(ctx);
  merge_free(ctx);
  exchange_free(ctx);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int kmd_create(const kmd_config* cfg, kmd_context** out)
{
  if (!cfg || !out) return fail(nullptr, KMD_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != KMD_ABI_VERSION) return fail(nullptr, KMD_ERR_ARG, "ABI version %u != %u", cfg->abi_version, KMD_ABI_VERSION);
  if (!cfg->nb_controls || !cfg->nb_cases) return fail(nullptr, KMD_ERR_ARG, "nb_controls and nb_cases must be > 0");
  if (cfg->kmer_words != 1 && cfg->kmer_words != 2) return fail(nullptr, KMD_ERR_ARG, "kmer_words must be 1 or 2 (k <= 64)");
  if (cfg->count_bytes != 1 && cfg->count_bytes != 2 && cfg->count_bytes != 4) return fail(nullptr, KMD_ERR_ARG, "count_bytes must be 1, 2 or 4");
  if (static_cast<uint64_t>(cfg->nb_controls) + cfg->nb_cases > 65536) return fail(nullptr, KMD_ERR_ARG, "more than 65536 samples");

  // KMD_TIMING=1: where the start-up time of a context goes (stderr)
  const bool timing = env_u32("KMD_TIMING", 0) != 0;
  const auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (timing)
      fprintf(stderr, "[kmd_create] %-28s +%.3f s\n", what,
              std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
  };
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
  {
    cudaGetLastError();
    return fail(nullptr, KMD_ERR_NO_DEVICE, "no CUDA device visible: kmdiff_b200 has no CPU fallback");
  }
  lap("cudaGetDeviceCount");
  if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, KMD_ERR_ARG, "device %d out of range (0..%d)", cfg->device, n_dev - 1);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, KMD_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major < 10)
    return fail(nullptr, KMD_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

  kmd_context* ctx = new (std::nothrow) kmd_context;
  if (!ctx) return fail(nullptr, KMD_ERR_NOMEM, "out of host memory");
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cap = cfg->max_survivors ? cfg->max_survivors : (1ull << 22);
  if (ctx->cap >= (1ull << 31)) { delete ctx; return fail(nullptr, KMD_ERR_ARG, "max_survivors must be < 2^31"); }

  auto bail = [&](int rc) { g_create_error = ctx->err; kmd_destroy(ctx); return rc; };
#define KMD_TRY(expr) do { int rc_ = (expr); if (rc_) return bail(rc_); } while (0)
#define KMD_CUDA_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fail(ctx, e_ == cudaErrorMemoryAllocation ? KMD_ERR_NOMEM : KMD_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    return bail(e_ == cudaErrorMemoryAllocation ? KMD_ERR_NOMEM : KMD_ERR_CUDA); } } while (0)

  KMD_CUDA_C(cudaSetDevice(ctx->device));
  KMD_CUDA_C(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  KMD_CUDA_C(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  KMD_CUDA_C(cudaEventCreate(&ctx->ev_k0));
  KMD_CUDA_C(cudaEventCreate(&ctx->ev_k1));
  lap("context + streams");

  // model constants + log-factorial table
  std::vector<double> table;
  const double* lf_host = cfg->lf_table;
  if (!lf_host && cfg->lf_size)
  {
    table.resize(cfg->lf_size);
    lf_table_fill(table.data(), cfg->lf_size);
    lf_host = table.data();
  }
  if (cfg->lf_size)
  {
    KMD_CUDA_C(cudaMalloc(&ctx->d_lf, cfg->lf_size * sizeof(double)));
    KMD_CUDA_C(cudaMemcpy(ctx->d_lf, lf_host, cfg->lf_size * sizeof(double), cudaMemcpyHostToDevice));
  }
  ctx->model.sum_controls = static_cast<double>(cfg->sum_controls);
  ctx->model.sum_cases = static_cast<double>(cfg->sum_cases);
  ctx->model.sum_all = static_cast<double>(cfg->sum_controls + cfg->sum_cases);
  ctx->model.lgam_half = kmd::lngamma_small(0.5);
  ctx->model.lf = ctx->d_lf;
  ctx->model.lf_size = cfg->lf_size;
  ctx->lr_cut = lr_cut_for(cfg->threshold);
  lap("lf table + model constants");

  KMD_TRY(make_plan(ctx));
  // The attribute belongs to the function and the device, not to a context: always raise it to the largest value any
  // context can need (the 200 KiB ring budget + barriers; 12288 table entries), so that contexts with different
  // row shapes or table sizes on one device do not lower each other's limit.
  {
    const int scan_smem_max = static_cast<int>(kBarrierBytes + 200 * 1024 + 128);
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem_max));
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem_max));
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_scan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, scan_smem_max));
    const int lf_smem_bytes = 12288 * static_cast<int>(sizeof(double));
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_model_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem_bytes));
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_model_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem_bytes));
    KMD_CUDA_C(cudaFuncSetAttribute(lrt_model_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem_bytes));
  }

  lap("kernel attributes");
  // buffers
  const size_t cap = ctx->cap;
  KMD_CUDA_C(cudaMalloc(&ctx->d_cnt, sizeof(DeviceCounters)));
  KMD_CUDA_C(cudaMemset(ctx->d_cnt, 0, sizeof(DeviceCounters)));
  lap("first cudaMalloc + memset");
  KMD_CUDA_C(cudaMallocHost(&ctx->h_cnt, sizeof(DeviceCounters)));
  memset(ctx->h_cnt, 0, sizeof(DeviceCounters));
  lap("first cudaMallocHost");
  KMD_CUDA_C(cudaMalloc(&ctx->d_cand, cap * sizeof(uint32_t)));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.kmer0, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.kmer1, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.p, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.sign, cap * 4));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.part, cap * 4));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.mc, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.kk, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->surv.row, cap * 8));
  for (int b = 0; b < 2; b++)
  {
    KMD_CUDA_C(cudaMalloc(&ctx->d_keys[b], cap * 8));
    KMD_CUDA_C(cudaMalloc(&ctx->d_vals[b], cap * 4));
  }
  KMD_CUDA_C(cudaMalloc(&ctx->d_p_sorted, cap * 8));
  KMD_CUDA_C(cudaMalloc(&ctx->d_order, cap * 4));
  if (cfg->flags & KMD_FLAG_KEEP_COUNTS)   // KmerSign::m_counts_ratio of a default (WITH_POPSTRAT) reference build, kmer.hpp:118-123
    KMD_CUDA_C(cudaMalloc(&ctx->pop.d_surv_counts, cap * (cfg->nb_controls + cfg->nb_cases) * cfg->count_bytes));
  lap("device buffers");
#undef KMD_TRY
#undef KMD_CUDA_C
  *out = ctx;
  return KMD_OK;
}

int kmd_set_stream(kmd_context* ctx, void* cuda_stream)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return KMD_OK;
}

int kmd_sync(kmd_context* ctx)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KMD_OK;
}

int kmd_reset(kmd_context* ctx)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaMemsetAsync(ctx->d_cnt, 0, sizeof(DeviceCounters), ctx->stream));
  ctx->rows_tested = 0;
  ctx->max_row_key = 0;
  ctx->order_state = 0;
  ctx->final_valid = false;
  ctx->final_sorted = false;
  ctx->n_final = 0;
  return pop_reset(ctx);
}

int kmd_scan_device(kmd_context* ctx, const void* d_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition)
{
  if (!ctx) return KMD_ERR_ARG;
  if (n_rows == 0) return KMD_OK;
  if (!d_rows) return fail(ctx, KMD_ERR_ARG, "null row pointer");
  if (reinterpret_cast<uintptr_t>(d_rows) & 15u) return fail(ctx, KMD_ERR_ARG, "d_rows must be 16-byte aligned");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return scan_device_impl(ctx, static_cast<const uint8_t*>(d_rows), n_rows, row_base, partition);
}

// chunked H2D copies on the copy stream, each followed by the scan of the chunk on the context's stream; the two
// staging buffers alternate across calls, so the copy of chunk i+1 overlaps the scan of chunk i whatever the
// size of the caller's buffers
static int scan_host_enqueue(kmd_context* ctx, const void* h_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition)
{
  int rc = ensure_stage(ctx);
  if (rc) return rc;
  const uint64_t R = ctx->plan.row_bytes;
  const uint64_t chunk_rows = ctx->stage_bytes / R;
  const uint8_t* src = static_cast<const uint8_t*>(h_rows);
  uint64_t done = 0;
  while (done < n_rows)
  {
    const int b = ctx->stage_next;
    ctx->stage_next ^= 1;
    const uint64_t n = std::min(chunk_rows, n_rows - done);
    // the scan that last read this staging buffer (possibly enqueued by an earlier call) must be done;
    // waiting on a never-recorded event is a no-op
    KMD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_scan[b], 0));
    KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage[b], src + done * R, n * R, cudaMemcpyHostToDevice, ctx->copy_stream));
    KMD_CUDA(ctx, cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
    KMD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
    if ((rc = scan_device_impl(ctx, ctx->d_stage[b], n, row_base + done, partition))) return rc;
    KMD_CUDA(ctx, cudaEventRecord(ctx->ev_scan[b], ctx->stream));
    done += n;
  }
  return KMD_OK;
}

int kmd_scan_host(kmd_context* ctx, const void* h_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition)
{
  if (!ctx) return KMD_ERR_ARG;
  if (n_rows == 0) return KMD_OK;
  if (!h_rows) return fail(ctx, KMD_ERR_ARG, "null row pointer");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = scan_host_enqueue(ctx, h_rows, n_rows, row_base, partition);
  if (rc) return rc;
  // the caller may reuse h_rows as soon as we return
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return KMD_OK;
}

int kmd_scan_host_async(kmd_context* ctx, const void* h_rows, uint64_t n_rows, uint64_t row_base, uint32_t partition,
                        uint64_t* ticket)
{
  if (!ctx || !ticket) return KMD_ERR_ARG;
  *ticket = 0;
  if (n_rows == 0) return KMD_OK;
  if (!h_rows) return fail(ctx, KMD_ERR_ARG, "null row pointer");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = scan_host_enqueue(ctx, h_rows, n_rows, row_base, partition);
  if (rc) return rc;
  const uint64_t t = ctx->next_ticket;
  cudaEvent_t& ev = ctx->ev_ticket[t % kmd_context::kTicketRing];
  if (!ev) KMD_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  KMD_CUDA(ctx, cudaEventRecord(ev, ctx->copy_stream));
  // published last: kmd_host_wait may run on another thread and must find the event recorded
  __atomic_store_n(&ctx->next_ticket, t + 1, __ATOMIC_RELEASE);
  *ticket = t;
  return KMD_OK;
}

int kmd_host_wait(kmd_context* ctx, uint64_t ticket)
{
  if (!ctx) return KMD_ERR_ARG;
  if (ticket == 0) return KMD_OK;
  const uint64_t next = __atomic_load_n(&ctx->next_ticket, __ATOMIC_ACQUIRE);
  if (ticket >= next) return KMD_ERR_ARG;   // (no message: ctx->err belongs to the scanning thread)
  // the ring holds the last kTicketRing tickets; an older one was overwritten only after kTicketRing later copies
  // were enqueued behind it on the same stream, and callers keep fewer buffers than that in flight
  if (ticket + kmd_context::kTicketRing <= next) return KMD_OK;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return KMD_ERR_CUDA;
  return cudaEventSynchronize(ctx->ev_ticket[ticket % kmd_context::kTicketRing]) == cudaSuccess ? KMD_OK : KMD_ERR_CUDA;
}

int kmd_counters_get(kmd_context* ctx, kmd_counters* out)
{
  if (!ctx || !out) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = fetch_counters(ctx);
  out->rows = ctx->rows_tested;
  out->candidates = ctx->h_cnt->cand_total;
  out->survivors = ctx->h_cnt->surv_count;
  out->n_control = ctx->h_cnt->n_control;
  out->n_case = ctx->h_cnt->n_case;
  return rc;
}

uint64_t kmd_launch_count(const kmd_context* ctx) { return ctx ? ctx->launches : 0; }

int kmd_scan_kernel_ms(kmd_context* ctx, float* ms)
{
  if (!ctx || !ms) return KMD_ERR_ARG;
  if (!ctx->scan_timed) return fail(ctx, KMD_ERR_STATE, "no scan has been launched yet");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaEventSynchronize(ctx->ev_k1));
  KMD_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return KMD_OK;
}

int kmd_survivors(kmd_context* ctx, kmd_survivor* out, uint64_t cap, uint64_t* n)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_order(ctx, 1);
  if (rc) return rc;
  return download_records(ctx, ctx->d_vals[0], ctx->h_cnt->surv_count, out, cap, n);
}

int kmd_push_survivors2(kmd_context* ctx, const kmd_survivor* records, const void* counts, uint64_t n, uint64_t rows_tested)
{
  if (!ctx) return KMD_ERR_ARG;
  if (n && !records) return fail(ctx, KMD_ERR_ARG, "null records");
  if (n >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "too many records in one call");
  if (counts && !ctx->pop.d_surv_counts)
    return fail(ctx, KMD_ERR_STATE, "counts pushed into a context that keeps none (KMD_FLAG_KEEP_COUNTS / kmd_popstrat_enable)");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n)
  {
    int rc = ensure_out(ctx, n);
    if (rc) return rc;
    const uint32_t counts_bytes = (ctx->cfg.nb_controls + ctx->cfg.nb_cases) * ctx->cfg.count_bytes;
    DeviceBuf d_counts;
    if (counts)
    {
      KMD_CUDA(ctx, cudaMalloc(&d_counts.p, n * counts_bytes));
      KMD_CUDA(ctx, cudaMemcpyAsync(d_counts.p, counts, n * counts_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_out, records, n * sizeof(kmd_survivor), cudaMemcpyHostToDevice, ctx->stream));
    push_records_kernel<<<blocks_for(static_cast<uint32_t>(n)), 256, 0, ctx->stream>>>(static_cast<uint32_t>(n), ctx->d_out, ctx->surv,
                                                                                       ctx->d_cnt, static_cast<uint32_t>(ctx->cap),
                                                                                       d_counts.as<uint8_t>(), ctx->pop.d_surv_counts, counts_bytes);
    ctx->launches++;
    KMD_CUDA(ctx, cudaGetLastError());
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `records` / `counts` may be pageable and reused by the caller
  }
  for (uint64_t i = 0; i < n; i++) ctx->max_row_key = std::max(ctx->max_row_key, records[i].row);
  ctx->rows_tested += rows_tested;
  ctx->order_state = 0;
  ctx->final_valid = false;
  return KMD_OK;
}

int kmd_push_survivors(kmd_context* ctx, const kmd_survivor* records, uint64_t n, uint64_t rows_tested)
{
  return kmd_push_survivors2(ctx, records, nullptr, n, rows_tested);
}

static int model_rows_launch(kmd_context* ctx, const uint8_t* d_rows, uint64_t n_rows, double* p, int32_t* sign, double* mc, double* kk)
{
  ScanParams G = ctx->plan;
  G.lanes_per_row = 1;
  const uint32_t lf_smem = static_cast<uint32_t>(std::min<uint64_t>(ctx->cfg.lf_size, 12288));
  const size_t smem = lf_smem * sizeof(double);
  const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>((n_rows + 255) / 256, static_cast<uint64_t>(ctx->sm_count) * 8));
  switch (ctx->cfg.count_bytes)
  {
    case 1: lrt_model_rows_kernel<1><<<grid, 256, smem, ctx->stream>>>(d_rows, n_rows, G, ctx->model, lf_smem, p, sign, mc, kk); break;
    case 2: lrt_model_rows_kernel<2><<<grid, 256, smem, ctx->stream>>>(d_rows, n_rows, G, ctx->model, lf_smem, p, sign, mc, kk); break;
    default: lrt_model_rows_kernel<4><<<grid, 256, smem, ctx->stream>>>(d_rows, n_rows, G, ctx->model, lf_smem, p, sign, mc, kk); break;
  }
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

int kmd_model_rows_device(kmd_context* ctx, const void* d_rows, uint64_t n_rows, double* d_p, int32_t* d_sign,
                          double* d_mean_control, double* d_mean_case)
{
  if (!ctx) return KMD_ERR_ARG;
  if (n_rows == 0) return KMD_OK;
  if (!d_rows || !d_p || !d_sign || !d_mean_control || !d_mean_case) return fail(ctx, KMD_ERR_ARG, "null pointer");
  if (reinterpret_cast<uintptr_t>(d_rows) & 15u) return fail(ctx, KMD_ERR_ARG, "d_rows must be 16-byte aligned");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return model_rows_launch(ctx, static_cast<const uint8_t*>(d_rows), n_rows, d_p, d_sign, d_mean_control, d_mean_case);
}

int kmd_model_rows_host(kmd_context* ctx, const void* h_rows, uint64_t n_rows, double* p, int32_t* sign,
                        double* mean_control, double* mean_case)
{
  if (!ctx) return KMD_ERR_ARG;
  if (n_rows == 0) return KMD_OK;
  if (!h_rows || !p || !sign || !mean_control || !mean_case) return fail(ctx, KMD_ERR_ARG, "null pointer");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_stage(ctx);
  if (rc) return rc;
  const uint64_t R = ctx->plan.row_bytes;
  // chunk so that rows + 28 bytes of results per row fit one staging buffer
  const uint64_t chunk_rows = std::max<uint64_t>(1, ctx->stage_bytes / (R + 32));
  // one allocation for the four result arrays: nothing to leak when it fails
  uint8_t* d_res = nullptr;
  KMD_CUDA(ctx, cudaMalloc(&d_res, chunk_rows * 28));
  double* d_p = reinterpret_cast<double*>(d_res);
  double* d_mc = d_p + chunk_rows;
  double* d_kk = d_mc + chunk_rows;
  int32_t* d_sign = reinterpret_cast<int32_t*>(d_kk + chunk_rows);
  const uint8_t* src = static_cast<const uint8_t*>(h_rows);
  for (uint64_t done = 0; done < n_rows && rc == KMD_OK; done += chunk_rows)
  {
    const uint64_t n = std::min(chunk_rows, n_rows - done);
    cudaError_t e = cudaMemcpyAsync(ctx->d_stage[0], src + done * R, n * R, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = model_rows_launch(ctx, ctx->d_stage[0], n, d_p, d_sign, d_mc, d_kk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p + done, d_p, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mean_control + done, d_mc, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mean_case + done, d_kk, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sign + done, d_sign, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, KMD_ERR_CUDA, "model_rows_host: %s", cudaGetErrorString(e));
  }
  cudaFree(d_res);
  return rc;
}

int kmd_correct(kmd_context* ctx, int correction, double significance, uint64_t total_rows, uint64_t* n_control, uint64_t* n_case)
{
  if (!ctx) return KMD_ERR_ARG;
  if (correction < KMD_CORR_NOTHING || correction > KMD_CORR_HOLM) return fail(ctx, KMD_ERR_ARG, "unknown correction %d", correction);
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool sorted = correction == KMD_CORR_BENJAMINI || correction == KMD_CORR_HOLM;
  if (!sorted)
  {
    double cut;
    switch (correction)
    {
      case KMD_CORR_BONFERRONI: cut = significance / static_cast<double>(total_rows); break;                    // corrector.cpp:11
      case KMD_CORR_SIDAK: cut = 1 - std::pow(1 - significance, 1.0 / static_cast<double>(total_rows)); break;   // corrector.cpp:52
      default: cut = significance; break;                                                                        // corrector.cpp:88
    }
    KMD_CUDA(ctx, cudaMemsetAsync(&ctx->d_cnt->keep_control, 0, 3 * sizeof(uint32_t), ctx->stream));
    keep_unordered_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_cnt, static_cast<uint32_t>(ctx->cap), ctx->surv.p, cut,
                                                                       nullptr, ctx->surv.sign, ctx->d_order);
    ctx->launches++;
    KMD_CUDA(ctx, cudaGetLastError());
    int rc0 = fetch_counters(ctx);
    if (rc0) return rc0;
    ctx->n_final = ctx->h_cnt->keep_total;
    ctx->final_valid = true;
    ctx->final_sorted = false;
    ctx->final_by_p = false;
    if (n_control) *n_control = ctx->h_cnt->keep_control;
    if (n_case) *n_case = ctx->h_cnt->keep_case;
    return KMD_OK;
  }
  int rc = ensure_order(ctx, 2);
  if (rc) return rc;
  const uint32_t n = ctx->h_cnt->surv_count;
  // BENJAMINI / HOLM: ascending walk, stop at the first rejection (aggregator.hpp:282-310)
  KMD_CUDA(ctx, cudaMemsetAsync(&ctx->d_cnt->keep_control, 0, 3 * sizeof(uint32_t), ctx->stream));
  uint32_t n_keep = 0;
  if (n)
  {
    gather_p_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(n, ctx->d_vals[0], ctx->surv.p, ctx->d_p_sorted);
    set_u32_kernel<<<1, 1, 0, ctx->stream>>>(&ctx->d_cnt->first_reject, n);
    first_reject_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(ctx->d_p_sorted, n, correction == KMD_CORR_HOLM, significance,
                                                                total_rows, &ctx->d_cnt->first_reject);
    KMD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;
    if ((rc = fetch_counters(ctx))) return rc;
    n_keep = ctx->h_cnt->first_reject;
    if (n_keep)
    {
      KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_order, ctx->d_vals[0], n_keep * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      count_prefix_kernel<<<blocks_for(n_keep), 256, 0, ctx->stream>>>(n_keep, ctx->d_order, ctx->surv.sign, ctx->d_cnt);
      ctx->launches++;
      KMD_CUDA(ctx, cudaGetLastError());
    }
  }
  if ((rc = fetch_counters(ctx))) return rc;
  ctx->n_final = n_keep;
  ctx->final_valid = true;
  ctx->final_sorted = true;
  if (n_control) *n_control = ctx->h_cnt->keep_control;
  if (n_case) *n_case = ctx->h_cnt->keep_case;
  return KMD_OK;
}

int kmd_significant(kmd_context* ctx, kmd_survivor* out, uint64_t cap, uint64_t* n)
{
  if (!ctx) return KMD_ERR_ARG;
  if (!ctx->final_valid) return fail(ctx, KMD_ERR_STATE, "kmd_significant called before kmd_correct / kmd_keep_below");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->final_sorted && ctx->n_final > 1)
  {
    // stateless correctors and kmd_keep_below leave the kept slots in arrival order: put them in row order now,
    // then (kmd_keep_below: BENJAMINI / HOLM output) stably in ascending p
    iota_keys_kernel<<<blocks_for(ctx->n_final), 256, 0, ctx->stream>>>(ctx->n_final, ctx->surv.row, ctx->surv.p, 0, ctx->d_order,
                                                                         ctx->d_keys[0], ctx->d_vals[0]);
    ctx->launches++;
    int rc = radix_sort(ctx, ctx->n_final, row_key_passes(ctx));
    if (rc) return rc;
    if (ctx->final_by_p)
    {
      KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_vals[1], ctx->d_vals[0], ctx->n_final * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      iota_keys_kernel<<<blocks_for(ctx->n_final), 256, 0, ctx->stream>>>(ctx->n_final, ctx->surv.row, ctx->surv.p, 1, ctx->d_vals[1],
                                                                           ctx->d_keys[0], ctx->d_vals[0]);
      ctx->launches++;
      if ((rc = radix_sort(ctx, ctx->n_final))) return rc;
    }
    KMD_CUDA(ctx, cudaMemcpyAsync(ctx->d_order, ctx->d_vals[0], ctx->n_final * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->order_state = 0;   // the full-survivor ordering in d_vals[0] was overwritten
  }
  ctx->final_sorted = true;
  return download_records(ctx, ctx->d_order, ctx->n_final, out, cap, n);
}

int kmd_sorted_pvalues(kmd_context* ctx, double* out, uint64_t cap, uint64_t* n)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_order(ctx, 2);
  if (rc) return rc;
  const uint32_t cnt = ctx->h_cnt->surv_count;
  if (n) *n = cnt;
  if (!cnt || !out) return KMD_OK;
  if (cap < cnt) return fail(ctx, KMD_ERR_ARG, "output capacity %llu < %u", static_cast<unsigned long long>(cap), cnt);
  gather_p_kernel<<<blocks_for(cnt), 256, 0, ctx->stream>>>(cnt, ctx->d_vals[0], ctx->surv.p, ctx->d_p_sorted);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_p_sorted, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KMD_OK;
}

int kmd_survivor_pvalues_device(kmd_context* ctx, double* d_out, uint64_t cap, uint64_t* n)
{
  if (!ctx || !n) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = fetch_counters(ctx);
  if (rc) return rc;
  *n = ctx->h_cnt->surv_count;
  if (!d_out || *n == 0) return KMD_OK;
  if (cap < *n) return fail(ctx, KMD_ERR_ARG, "output capacity %llu < %llu survivors", static_cast<unsigned long long>(cap),
                            static_cast<unsigned long long>(*n));
  KMD_CUDA(ctx, cudaMemcpyAsync(d_out, ctx->surv.p, *n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return KMD_OK;
}

int kmd_first_reject_device(kmd_context* ctx, const double* d_p_sorted, uint64_t n, int correction, double significance,
                            uint64_t total_rows, uint64_t* n_accepted, double* p_cut)
{
  if (!ctx) return KMD_ERR_ARG;
  if (correction != KMD_CORR_BENJAMINI && correction != KMD_CORR_HOLM) return fail(ctx, KMD_ERR_ARG, "BENJAMINI or HOLM expected");
  if (n >= (1ull << 32)) return fail(ctx, KMD_ERR_ARG, "too many p-values");
  if (n && !d_p_sorted) return fail(ctx, KMD_ERR_ARG, "null p-values");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t first = static_cast<uint32_t>(n);
  double cut = std::numeric_limits<double>::infinity();
  if (n)
  {
    set_u32_kernel<<<1, 1, 0, ctx->stream>>>(&ctx->d_cnt->first_reject, first);
    first_reject_kernel<<<blocks_for(first), 256, 0, ctx->stream>>>(d_p_sorted, first, correction == KMD_CORR_HOLM, significance,
                                                                    total_rows, &ctx->d_cnt->first_reject);
    ctx->launches += 2;
    KMD_CUDA(ctx, cudaGetLastError());
    KMD_CUDA(ctx, cudaMemcpyAsync(&first, &ctx->d_cnt->first_reject, sizeof first, cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (first < n)
    {
      KMD_CUDA(ctx, cudaMemcpyAsync(&cut, d_p_sorted + first, sizeof cut, cudaMemcpyDeviceToHost, ctx->stream));
      KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  if (n_accepted) *n_accepted = first;
  if (p_cut) *p_cut = cut;
  return KMD_OK;
}

int kmd_keep_below(kmd_context* ctx, double p_cut, uint64_t* n_control, uint64_t* n_case)
{
  if (!ctx) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return keep_below_impl(ctx, p_cut, n_control, n_case);
}

int kmd_corrector_apply(int correction, double significance, uint64_t total_rows, uint64_t rank, double p)
{
  switch (correction)
  {
    case KMD_CORR_BONFERRONI: return p < (significance / static_cast<double>(total_rows));
    case KMD_CORR_BENJAMINI: return p < ((static_cast<double>(rank) / static_cast<double>(total_rows)) * significance);
    case KMD_CORR_SIDAK: return p < (1 - std::pow(1 - significance, 1.0 / static_cast<double>(total_rows)));
    case KMD_CORR_HOLM: return p < (significance / static_cast<double>(total_rows - rank));
    default: return p < significance;
  }
}

uint64_t kmd_first_reject(const double* p_sorted, uint64_t n, int correction, double significance, uint64_t total_rows, double* p_cut)
{
  uint64_t i = 0;
  for (; i < n; i++)
  {
    const uint64_t rank = correction == KMD_CORR_BENJAMINI ? i + 1 : i;
    if (!kmd_corrector_apply(correction, significance, total_rows, rank, p_sorted[i])) break;
  }
  if (p_cut) *p_cut = i < n ? p_sorted[i] : std::numeric_limits<double>::infinity();
  return i;
}

void kmd_lf_table_fill(double* table, uint64_t size) { lf_table_fill(table, size); }

double kmd_chisqc(double v, double x)
{
  // only v == 1 (a = 0.5) is on the path; other v are outside this library's scope
  if (v != 1.0) return std::numeric_limits<double>::quiet_NaN();
  return host_igamc_half(x / 2.0);
}

void kmd_kmer_to_string(const uint64_t* words, uint32_t kmer_words, uint32_t k, char* out)
{
  static const char alphabet[4] = {'A', 'C', 'T', 'G'};
  for (uint32_t i = 0; i < k; i++)
  {
    const uint32_t bit = 2 * (k - 1 - i);
    const uint32_t w = bit / 64;
    out[i] = alphabet[w < kmer_words ? (words[w] >> (bit % 64)) & 3u : 0];
  }
  out[k] = 0;
}

// fmt's default ("{}") double formatting: shortest digits that round-trip; fixed notation when the
// decimal exponent is in [-4, 16), exponent notation otherwise.
static int format_shortest(double x, char* out, size_t cap)
{
  if (std::isnan(x)) return snprintf(out, cap, "nan");
  if (std::isinf(x)) return snprintf(out, cap, x < 0 ? "-inf" : "inf");
  if (x == 0) return snprintf(out, cap, std::signbit(x) ? "-0" : "0");
  char buf[40];
  int prec = 0;
  for (; prec < 17; prec++)
  {
    snprintf(buf, sizeof buf, "%.*e", prec, x);
    if (strtod(buf, nullptr) == x) break;
  }
  // buf = [-]d[.ddd]e[+-]XX
  std::string digits;
  const char* q = buf;
  bool neg = *q == '-';
  if (neg) q++;
  for (; *q && *q != 'e'; q++)
    if (*q != '.') digits.push_back(*q);
  const int exp10 = atoi(q + 1);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string s = neg ? "-" : "";
  if (exp10 >= -4 && exp10 < 16)
  {
    if (exp10 < 0)
    {
      s += "0.";
      s.append(static_cast<size_t>(-exp10 - 1), '0');
      s += digits;
    }
    else
    {
      const size_t int_len = static_cast<size_t>(exp10) + 1;
      if (digits.size() <= int_len)
      {
        s += digits;
        s.append(int_len - digits.size(), '0');
      }
      else
      {
        s += digits.substr(0, int_len);
        s += ".";
        s += digits.substr(int_len);
      }
    }
  }
  else
  {
    s += digits.substr(0, 1);
    if (digits.size() > 1) { s += "."; s += digits.substr(1); }
    char e[16];
    snprintf(e, sizeof e, "e%c%02d", exp10 < 0 ? '-' : '+', std::abs(exp10));
    s += e;
  }
  return snprintf(out, cap, "%s", s.c_str());
}

size_t kmd_format_fasta(uint64_t index, const kmd_survivor* s, uint32_t kmer_words, uint32_t k, char* out, size_t cap)
{
  char mean_case[48];
  format_shortest(s->mean_case, mean_case, sizeof mean_case);
  const double mc = s->mean_control;
  const unsigned long long mc_int = !(mc >= 0) ? 0ull : (mc >= 18446744073709551615.0 ? ~0ull : static_cast<unsigned long long>(mc));
  char head[192];
  const int hl = snprintf(head, sizeof head, ">%llu_pval=%g_control=%llu_case=%s\n", static_cast<unsigned long long>(index),
                          s->p_value, mc_int, mean_case);
  std::vector<char> seq(k + 1);
  kmd_kmer_to_string(s->kmer, kmer_words, k, seq.data());
  const size_t lines = (k + 59) / 60;
  const size_t need = static_cast<size_t>(hl) + k + lines;
  if (need + 1 > cap || !out) return need;
  char* w = out;
  memcpy(w, head, hl); w += hl;
  for (uint32_t i = 0; i < k; i += 60)
  {
    const uint32_t n = std::min<uint32_t>(60, k - i);
    memcpy(w, seq.data() + i, n); w += n;
    *w++ = '\n';
  }
  *w = 0;
  return need;
}

// CPU set "0-3,8,10-11" (sysfs cpulist) -> cpu_set_t restricted to the CPUs the thread may already use
static bool parse_cpulist(const char* text, const cpu_set_t& allowed, cpu_set_t* out)
{
  CPU_ZERO(out);
  int n = 0;
  const char* p = text;
  while (*p)
  {
    char* end = nullptr;
    const long a = strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    p = end;
    if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++)
      if (c >= 0 && CPU_ISSET(c, &allowed)) { CPU_SET(c, out); n++; }
    if (*p == ',') p++;
  }
  return n > 0;
}

int kmd_bind_host_thread(int device)
{
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return KMD_ERR_ARG; }
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return KMD_OK; }
  for (char* c = bus; *c; c++) *c = static_cast<char>(tolower(*c));
  auto slurp = [&](const char* leaf, char* buf, size_t cap) {
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/%s", bus, leaf);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const size_t n = fread(buf, 1, cap - 1, f);
    fclose(f);
    buf[n] = 0;
    return n > 0;
  };
  char text[4096];
  int node = -1;
  if (slurp("numa_node", text, sizeof text)) node = atoi(text);
  if (node < 0) return KMD_OK;   // single-node host or a VM that hides the topology: nothing to bind
  cpu_set_t allowed, want;
  if (sched_getaffinity(0, sizeof allowed, &allowed) == 0 && slurp("local_cpulist", text, sizeof text) &&
      parse_cpulist(text, allowed, &want))
    sched_setaffinity(0, sizeof want, &want);
  // pages this thread touches from now on (pinned staging buffers above all) come from the GPU's node
  unsigned long mask[16] = {0};
  if (node < static_cast<int>(sizeof mask * 8))
  {
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
  }
  return KMD_OK;
}

int kmd_warmup(int device)
{
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return KMD_ERR_ARG; }
  return cudaFree(nullptr) == cudaSuccess ? KMD_OK : KMD_ERR_CUDA;
}

void* kmd_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void kmd_host_free(void* p) { if (p) cudaFreeHost(p); }

void kmd_synth_totals(const kmd_synth* s, uint64_t* totals)
{
  const uint32_t n = s->nb_controls + s->nb_cases;
  // every sample is drawn from the same abundance classes, so the declared per-sample totals are
  // equal up to +-1 % (a sequencing-depth jitter), seeded like the rows
  for (uint32_t j = 0; j < n; j++)
    totals[j] = 990000000ull + synth::mix64(s->seed * 0x9E3779B97F4A7C15ull + j) % 20000001ull;
}

int kmd_synth_fill_host(const kmd_synth* s, void* h_rows, uint64_t n_rows, uint64_t row_base, int threads)
{
  synth::Params p;
  if (synth_params(s, &p) || !h_rows) return KMD_ERR_ARG;
  const synth::Tables* t = build_synth_tables();
  uint8_t* dst = static_cast<uint8_t*>(h_rows);
  auto work = [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) synth::make_row(p, *t, row_base + i, dst + i * p.row_bytes);
  };
  if (threads <= 1 || n_rows < 4096) { work(0, n_rows); return KMD_OK; }
  std::vector<std::thread> pool;
  const uint64_t chunk = (n_rows + threads - 1) / threads;
  for (int i = 0; i < threads; i++)
  {
    const uint64_t lo = std::min(n_rows, i * chunk), hi = std::min(n_rows, lo + chunk);
    pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return KMD_OK;
}

int kmd_synth_fill_device(const kmd_synth* s, int device, void* d_rows, uint64_t n_rows, uint64_t row_base, void* cuda_stream)
{
  synth::Params p;
  if (synth_params(s, &p) || !d_rows) return KMD_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return KMD_ERR_CUDA;
  static synth::Tables* d_tables[64] = {nullptr};
  if (device < 0 || device >= 64) return KMD_ERR_ARG;
  if (!d_tables[device])
  {
    if (cudaMalloc(&d_tables[device], sizeof(synth::Tables)) != cudaSuccess) return KMD_ERR_NOMEM;
    if (cudaMemcpy(d_tables[device], build_synth_tables(), sizeof(synth::Tables), cudaMemcpyHostToDevice) != cudaSuccess) return KMD_ERR_CUDA;
  }
  if (n_rows == 0) return KMD_OK;
  const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>((n_rows + 127) / 128, 148 * 16));
  synth_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(p, d_tables[device], static_cast<uint8_t*>(d_rows), n_rows, row_base);
  return cudaGetLastError() == cudaSuccess ? KMD_OK : KMD_ERR_CUDA;
}

}  // extern "C"

#include "kmd_popstrat_api.inl"
#include "kmd_merge_api.inl"
#include "kmd_exchange_api.inl"
