// LRT scan kernel for sm_100a — the streaming half of diff_observer::process.
//
// Replaces, per row (reference file:line):
//   matrix_proxy::merge / MatrixReader::read   include/kmdiff/merge.hpp:194-203,
//                                              kmtricks/io/matrix_file.hpp:144-154
//   Model::compute_sum_e (controls, cases)     include/kmdiff/model.hpp:70-81
//   the decision `p_value <= m_threshold`      include/kmdiff/merge.hpp:78
//
// Design: the kernel is bound by HBM bandwidth — R = 8W + N*B bytes per row and a
// few dozen integer/fp32 instructions.  A persistent CTA per SM streams tiles of whole rows from
// global memory into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx; SASS: UBLKCP), issued by one elected producer thread, so tens of KB per SM are in
// flight without holding registers.  Consumer warps read rows back from shared memory (a row is
// handled by `g` lanes, g chosen on the host so the LDS.64 pattern is bank-conflict free), sum
// controls and cases exactly in integers, and evaluate a *conservative fp32 screen* of the
// likelihood ratio:
//
//     LR = kc*ln(kc/ec) + kk*ln(kk/ek),  ec = (kc+kk)*Sc/(Sc+Sk),  ek = (kc+kk)*Sk/(Sc+Sk)
//
// (the closed form of alt-null of model.hpp:153-160: the factorial and the linear terms cancel).
// p <= thr  <=>  LR >= LR*(thr) because igamc(0.5, .) is decreasing; rows whose screened LR plus a
// proven error margin reaches LR* are emitted as *candidates* (warp-ballot compaction); everything
// that decides a p-value is then done in fp64, reference operation order, by lrt_exact_kernel
// (kmd_api.cu) on the candidates only (~1e-3 of the rows).  The screen can only err towards
// keeping a row, never towards dropping one.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace kmd {

constexpr int kMaxConsumerWarps = 31;                    // + 1 producer warp = 1024 threads
constexpr int kMaxScanThreads = (kMaxConsumerWarps + 1) * 32;
constexpr int kMaxStages = 16;
constexpr int kBarrierBytes = 16 * kMaxStages;           // full[] + empty[] mbarriers

struct ScanParams
{
  const uint8_t* rows;     // device, 16-byte aligned
  uint64_t n_rows;         // < 2^32
  uint32_t row_bytes;      // R
  uint32_t kmer_bytes;     // 8W
  uint32_t nb_controls;
  uint32_t nb_cases;
  uint32_t tile_rows;      // rows per ring tile; tile_rows*R is a multiple of 16
  uint32_t n_tiles;        // full tiles streamed through the ring
  uint32_t stages;
  uint32_t stage_bytes;    // >= tile_rows*R, multiple of 128
  uint32_t lanes_per_row;  // g: power of two, 1..32
  uint32_t consumer_warps; // warps 0..consumer_warps-1 consume, warp consumer_warps produces
  uint32_t groups;         // consumer groups; group j consumes the CTA's tiles t = j, j+groups, ...
  uint32_t group_warps;    // consumer_warps / groups; stages % groups == 0
  uint32_t fast;           // 1: (N*B) % 8 == 0 -> 8-byte word path
  uint32_t words_ctrl;     // words made of control counts only
  uint32_t straddle_elems; // control counts in the word that straddles the cohorts (0: none)
  uint32_t words_total;    // N*B/8
  uint32_t screen_off;     // 1: every row is a candidate (thr >= 1, or a degenerate cohort total)
  float frac_ctrl;         // Sc/(Sc+Sk)
  float frac_case;         // Sk/(Sc+Sk)
  float lr2_cut;           // (LR*(thr) - slack) / ln 2
  uint32_t* cand_count;
  uint32_t* cand;          // row index inside the batch
  uint32_t cand_cap;
};

// ---- PTX wrappers ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- loaders: the same row arithmetic runs on the shared-memory ring and on global memory ----

struct SmemLoader
{
  uint32_t base;
  __device__ __forceinline__ uint64_t ld64(uint32_t off) const
  {
    uint64_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(base + off));
    return v;
  }
  __device__ __forceinline__ uint32_t ld32(uint32_t off) const
  {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + off));
    return v;
  }
  __device__ __forceinline__ uint32_t ld16(uint32_t off) const
  {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + off));
    return v;
  }
  __device__ __forceinline__ uint32_t ld8(uint32_t off) const
  {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + off));
    return v;
  }
};

struct GmemLoader
{
  const uint8_t* base;
  __device__ __forceinline__ uint64_t ld64(uint32_t off) const
  {
    return __ldg(reinterpret_cast<const unsigned long long*>(base + off));
  }
  __device__ __forceinline__ uint32_t ld32(uint32_t off) const
  {
    return __ldg(reinterpret_cast<const uint32_t*>(base + off));
  }
  __device__ __forceinline__ uint32_t ld16(uint32_t off) const
  {
    return __ldg(reinterpret_cast<const uint16_t*>(base + off));
  }
  __device__ __forceinline__ uint32_t ld8(uint32_t off) const { return __ldg(base + off); }
};

// ---- exact integer row sums (model.hpp:70-81; integer sums == the reference's double sums) ----

template <int B>
struct Acc;

template <>
struct Acc<4>
{
  uint64_t v = 0;
  // one IMAD.WIDE.U32 per count: 64-bit accumulate of a 32-bit value on the FMA pipe
  __device__ __forceinline__ void elem(uint32_t e) { asm("mad.wide.u32 %0, %1, 1, %0;" : "+l"(v) : "r"(e)); }
  __device__ __forceinline__ void word(uint64_t x)
  {
    elem(static_cast<uint32_t>(x));
    elem(static_cast<uint32_t>(x >> 32));
  }
  __device__ __forceinline__ uint64_t total() const { return v; }
};
template <>
struct Acc<2>
{
  uint32_t v = 0;  // N*65535 < 2^32 for any N <= 65536 samples
  // dp2a: v += lo16(x)*1 + hi16(x)*1 — one instruction per two counts
  __device__ __forceinline__ void word(uint64_t x)
  {
    v = __dp2a_lo(static_cast<uint32_t>(x), 0x00000101u, v);
    v = __dp2a_lo(static_cast<uint32_t>(x >> 32), 0x00000101u, v);
  }
  __device__ __forceinline__ void elem(uint32_t e) { v += e; }
  __device__ __forceinline__ uint64_t total() const { return v; }
};
template <>
struct Acc<1>
{
  uint32_t v = 0;
  __device__ __forceinline__ void word(uint64_t x)
  {
    v = __dp4a(static_cast<uint32_t>(x), 0x01010101u, v);
    v = __dp4a(static_cast<uint32_t>(x >> 32), 0x01010101u, v);
  }
  __device__ __forceinline__ void elem(uint32_t e) { v += e; }
  __device__ __forceinline__ uint64_t total() const { return v; }
};

template <int B, class Loader>
__device__ __forceinline__ uint32_t load_elem(const Loader& ld, uint32_t off)
{
  if (B == 4) return ld.ld32(off);
  if (B == 2) return ld.ld16(off);
  return ld.ld8(off);
}

// Partial sums of one row for lane `l` of the `g` lanes that share it.
template <int B, class Loader>
__device__ __forceinline__ void row_partial(const ScanParams& P, const Loader& ld, uint32_t row_off,
                                            uint32_t l, uint32_t g, uint64_t& sum_ctrl, uint64_t& sum_case)
{
  Acc<B> ac, ak;
  const uint32_t cbase = row_off + P.kmer_bytes;
  if (P.fast)
  {
    for (uint32_t w = l; w < P.words_ctrl; w += g) ac.word(ld.ld64(cbase + 8 * w));
    uint32_t first_case = P.words_ctrl;
    if (P.straddle_elems)
    {
      if (l == (P.words_ctrl & (g - 1)))
      {
        const uint64_t x = ld.ld64(cbase + 8 * P.words_ctrl);
#pragma unroll
        for (uint32_t e = 0; e < 8 / B; e++)
        {
          const uint32_t v = static_cast<uint32_t>(x >> (8 * B * e)) & (B == 4 ? 0xFFFFFFFFu : (B == 2 ? 0xFFFFu : 0xFFu));
          if (e < P.straddle_elems) ac.elem(v); else ak.elem(v);
        }
      }
      first_case++;
    }
    for (uint32_t w = first_case + ((l - first_case) & (g - 1)); w < P.words_total; w += g)
      ak.word(ld.ld64(cbase + 8 * w));
  }
  else
  {
    const uint32_t n = P.nb_controls + P.nb_cases;
    for (uint32_t j = l; j < n; j += g)
    {
      const uint32_t v = load_elem<B>(ld, cbase + j * B);
      if (j < P.nb_controls) ac.elem(v); else ak.elem(v);
    }
  }
  sum_ctrl = ac.total();
  sum_case = ak.total();
}

// Conservative fp32 screen.  Returns true when the row MAY satisfy p <= thr.
// Error budget (log2 units, per unit of count): conversion u64->f32 6e-8, ec/ek 1.8e-7,
// approximate division 2.4e-7 => relative error of the ratio <= 4.2e-7 => 6.1e-7 in log2;
// lg2.approx 2.4e-7 absolute + 2.4e-7 relative; products/sum 1.2e-7 relative.  Total
// <= 8.5e-7*k + 3.6e-7*|k*log2 x| per cohort; the margin charges 2.5e-6 on both, plus 0.08
// absolute (host slack on LR* included).  tests/test_gpu_parity.py attacks this bound (adversarial rows against
// the oracle; 1e8 (kc, kk) pairs up to 2^32 - 1 per cohort against the exact kernel) and tests/test_screen_emulation.py
// replays the same arithmetic in numpy float32 on the CPU.
__device__ __forceinline__ bool screen_pass(const ScanParams& P, uint64_t sum_ctrl, uint64_t sum_case)
{
  const float kc = static_cast<float>(sum_ctrl);
  const float kk = static_cast<float>(sum_case);
  const float t = kc + kk;
  const float ec = t * P.frac_ctrl;
  const float ek = t * P.frac_case;
  const float lc = kc > 0.f ? __log2f(__fdividef(kc, ec)) : 0.f;
  const float lk = kk > 0.f ? __log2f(__fdividef(kk, ek)) : 0.f;
  const float tc = kc * lc;
  const float tk = kk * lk;
  const float lr2 = tc + tk;
  const float margin = 0.08f + 2.5e-6f * (fabsf(tc) + fabsf(tk) + t);
  return !(lr2 + margin < P.lr2_cut);  // NaN => keep
}

// Sum, screen and compact the rows [first_row, first_row + n) of one tile / tail.
template <int B, class Loader>
__device__ __forceinline__ void scan_rows(const ScanParams& P, const Loader& ld, uint32_t n,
                                          uint32_t first_row, uint32_t warp, uint32_t n_warps,
                                          uint32_t lane)
{
  const uint32_t g = P.lanes_per_row;
  const uint32_t rows_per_warp = 32 / g;
  const uint32_t sub = lane / g;
  const uint32_t l = lane & (g - 1);
  for (uint32_t r0 = warp * rows_per_warp; r0 < n; r0 += n_warps * rows_per_warp)
  {
    const uint32_t r = r0 + sub;
    const bool valid = r < n;
    uint64_t sc = 0, sk = 0;
    if (valid) row_partial<B>(P, ld, r * P.row_bytes, l, g, sc, sk);
    for (uint32_t o = g >> 1; o > 0; o >>= 1)
    {
      sc += __shfl_xor_sync(0xFFFFFFFFu, sc, o);
      sk += __shfl_xor_sync(0xFFFFFFFFu, sk, o);
    }
    const bool keep = valid && l == 0 && (P.screen_off || screen_pass(P, sc, sk));
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
    if (ballot)
    {
      uint32_t pos = 0;
      if (lane == 0) pos = atomicAdd(P.cand_count, __popc(ballot));
      pos = __shfl_sync(0xFFFFFFFFu, pos, 0) + __popc(ballot & ((1u << lane) - 1));
      if (keep && pos < P.cand_cap) P.cand[pos] = first_row + r;
    }
  }
}

template <int B>
__global__ void __launch_bounds__(kMaxScanThreads, 1) lrt_scan_kernel(const ScanParams P)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t bars = smem_u32(smem);              // full[s] at +8s, empty[s] at +8*kMaxStages+8s
  const uint32_t ring = bars + kBarrierBytes;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x == 0)
  {
    for (uint32_t s = 0; s < P.stages; s++)
    {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * kMaxStages + 8 * s, P.group_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t tile_bytes = P.tile_rows * P.row_bytes;

  if (warp == P.consumer_warps)
  {
    // ---- producer: one thread keeps `stages` tiles in flight ----
    if (lane == 0)
    {
      uint32_t i = 0;
      for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, i++)
      {
        const uint32_t s = i % P.stages;
        const uint32_t phase = (i / P.stages) & 1;
        mbar_wait(bars + 8 * kMaxStages + 8 * s, phase ^ 1);  // consumers released this slot
        mbar_expect_tx(bars + 8 * s, tile_bytes);
        bulk_g2s(ring + s * P.stage_bytes, P.rows + static_cast<uint64_t>(tile) * tile_bytes,
                 tile_bytes, bars + 8 * s);
      }
    }
  }
  else
  {
    // ---- consumers: group j of group_warps warps owns the CTA's tiles t = j, j+groups, ... so that
    //      several tiles are being summed at once while each stays small enough for a deep ring ----
    const uint32_t group = warp / P.group_warps;
    const uint32_t gwarp = warp - group * P.group_warps;
    for (uint32_t t = group;; t += P.groups)
    {
      const uint64_t tile = blockIdx.x + static_cast<uint64_t>(t) * gridDim.x;
      if (tile >= P.n_tiles) break;
      const uint32_t s = t % P.stages;
      const uint32_t phase = (t / P.stages) & 1;
      mbar_wait(bars + 8 * s, phase);
      SmemLoader ld{ring + s * P.stage_bytes};
      scan_rows<B>(P, ld, P.tile_rows, static_cast<uint32_t>(tile) * P.tile_rows, gwarp, P.group_warps, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * kMaxStages + 8 * s);
    }
    // ---- tail: rows after the last full tile, straight from global memory, by the last CTA ----
    const uint64_t done = static_cast<uint64_t>(P.n_tiles) * P.tile_rows;
    if (blockIdx.x == gridDim.x - 1 && done < P.n_rows)
    {
      GmemLoader ld{P.rows + done * P.row_bytes};
      scan_rows<B>(P, ld, static_cast<uint32_t>(P.n_rows - done), static_cast<uint32_t>(done), warp,
                   P.consumer_warps, lane);
    }
  }
}

}  // namespace kmd
