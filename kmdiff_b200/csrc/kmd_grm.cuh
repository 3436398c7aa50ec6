// Eigenstrat Gram matrix of the sampled k-mer presence vectors as an exact int8 tensor-core GEMM
// (SURVEY.md §8 a18).
//
// Replaces HAWK's smartpca accumulation (thirdparty/hawk/EIG6.0.1-Hawk/src/eigensrc/smartpca.c:
// 907-1025 main loop, getcolxz :2598-2700, fvadjust :1684-1800, block_increment_normal :2985-3004):
// per SNP the column is centred by mu = c/n and scaled by f(c), both functions of the carrier count c
// only, and XTX += col col^T in fp64.  Here SNPs are bucketed by c, so
//
//     XTX = sum_c f_c^2 [ G_c - mu_c (r_c 1^T + 1 r_c^T) + N_c mu_c^2 1 1^T ],   G_c = sum_{s in c} b_s b_s^T
//
// with b_s the 0/1 presence vector: G_c is an integer matrix (r_c is its diagonal) computed exactly by
// tcgen05.mma kind::i8 with int32 accumulators in TMEM; the fp64 combination runs once per bucket.
//
// Layout: Pt[i][k] (u8, K-major: sample i, bucket-sorted SNP position k), every bucket padded with
// zero columns to a multiple of 128 so that a 128-byte K block belongs to one bucket.  A work item is
// (bucket segment of <= kMaxKBlocks K blocks, 128-row M tile, 512-column N half); its 128 x <=512
// int32 accumulator occupies the CTA's TMEM and is added into G_c with integer atomics (exact and
// order-independent, so split-K is deterministic).
//
// Kernel roles (one CTA per SM, 192 threads): warp 0 = TMA producer (2-D tiled loads, 128B swizzle),
// warp 1 = TMEM allocation + single-thread tcgen05.mma issue, warps 2-5 = epilogue (tcgen05.ld ->
// red.global.add.s32).
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "kmd_scan.cuh"  // mbarrier wrappers

namespace kmd {
namespace grm {

constexpr uint32_t kBlockK = 128;        // bytes (= int8 elements) per K block: one 128B swizzle span
constexpr uint32_t kUmmaK = 32;          // K of one tcgen05.mma kind::i8
constexpr uint32_t kTileM = 128;
constexpr uint32_t kMaxN = 512;          // TMEM columns
constexpr uint32_t kStages = 2;
constexpr uint32_t kChunkBytes = kTileM * kBlockK;             // one TMA box: 128 rows x 128 B = 16 KB
constexpr uint32_t kStageBytes = kChunkBytes * (1 + kMaxN / 128);  // A + up to 4 B chunks = 80 KB
constexpr uint32_t kMaxKBlocks = 64;     // K blocks per work item (8192 SNPs)
constexpr uint32_t kThreads = 192;
constexpr size_t kSmemBytes = 1024 /*alignment slack*/ + kStages * kStageBytes + 256 /*barriers*/;

struct WorkItem
{
  uint32_t bucket;   // compact bucket index -> G + bucket * n * n
  uint32_t kb0;      // first K block (global, in units of kBlockK)
  uint32_t kb_cnt;
  uint16_t mtile;
  uint16_t nhalf;
};

// ---- bucket sort of the sampled SNPs by carrier count ----------------------------------------

__global__ void carriers_hist_kernel(const uint32_t* carriers, uint64_t m, uint32_t* hist /* n+1 */)
{
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < m;
       s += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&hist[carriers[s]], 1u);
}

// perm[koff[c] + j] = j-th SNP (in arrival order of the atomics) with carrier count c; padding keeps ~0u
__global__ void carriers_scatter_kernel(const uint32_t* carriers, uint64_t m, const uint64_t* koff, uint32_t* cursor,
                                        uint32_t* perm)
{
  for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < m;
       s += static_cast<uint64_t>(gridDim.x) * blockDim.x)
  {
    const uint32_t c = carriers[s];
    if (c == 0) continue;  // no carrier: ignored by smartpca (smartpca.c:917-930)
    perm[koff[c] + atomicAdd(&cursor[c], 1u)] = static_cast<uint32_t>(s);
  }
}

// Pt[i][k] = geno[perm[k]][i].  A lane owns 4 consecutive SNPs k and walks the samples 4 at a time: four 32-bit loads
// (one per SNP; the rows are read front to back, so the 32-byte sectors they touch are reused from L1), a 4 x 4 byte
// transpose in registers (PRMT) and four 32-bit stores — the warp's stores of one sample are 128 contiguous bytes.
// Rows n .. rows_alloc-1 and columns without a SNP are written as zeros.
__global__ void __launch_bounds__(256) transpose_presence_kernel(const uint8_t* __restrict__ geno, uint32_t n,
                                                                const uint32_t* __restrict__ perm, uint64_t k_pad,
                                                                uint32_t rows_alloc, uint8_t* __restrict__ pt)
{
  const uint64_t k0 = (static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (k0 >= k_pad) return;   // k_pad is a multiple of 4
  const uint8_t* src[4];
#pragma unroll
  for (int q = 0; q < 4; q++)
  {
    const uint32_t s = perm[k0 + q];
    src[q] = s != 0xFFFFFFFFu ? geno + static_cast<uint64_t>(s) * n : nullptr;
  }
  const bool words = (n & 3u) == 0;
  for (uint32_t i = 0; i < rows_alloc; i += 4)
  {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
      w[q] = 0;
      if (src[q] && i < n)
      {
        if (words) w[q] = *reinterpret_cast<const uint32_t*>(src[q] + i);
        else
          for (uint32_t b = 0; b < 4 && i + b < n; b++) w[q] |= static_cast<uint32_t>(src[q][i + b]) << (8 * b);
      }
    }
    // o[j] = (w0.byte j, w1.byte j, w2.byte j, w3.byte j)
    const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140);   // w0.b0 w1.b0 w0.b1 w1.b1
    const uint32_t t1 = __byte_perm(w[2], w[3], 0x5140);
    const uint32_t t2 = __byte_perm(w[0], w[1], 0x7362);   // w0.b2 w1.b2 w0.b3 w1.b3
    const uint32_t t3 = __byte_perm(w[2], w[3], 0x7362);
    const uint32_t o0 = __byte_perm(t0, t1, 0x5410), o1 = __byte_perm(t0, t1, 0x7632);
    const uint32_t o2 = __byte_perm(t2, t3, 0x5410), o3 = __byte_perm(t2, t3, 0x7632);
    uint8_t* dst = pt + static_cast<uint64_t>(i) * k_pad + k0;
    *reinterpret_cast<uint32_t*>(dst) = o0;
    *reinterpret_cast<uint32_t*>(dst + k_pad) = o1;
    *reinterpret_cast<uint32_t*>(dst + 2 * k_pad) = o2;
    *reinterpret_cast<uint32_t*>(dst + 3 * k_pad) = o3;
  }
}

// ---- tcgen05 / TMA PTX ----------------------------------------------------------------------------

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity)
{
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done)
  {
    asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 8000000000ll) __trap();  // ~4 s: a lost arrival must not hang the GPU
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t x, int32_t y)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
    "{\n.reg .pred p;\n"
    "setp.ne.b32 p, %4, 0;\n"
    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
    ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO); the
// leading-dimension offset is unused for a swizzled K-major tile no wider than the swizzle span.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address
  d |= static_cast<uint64_t>(1) << 16;                        // leading byte offset (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                        // SWIZZLE_128B
  return d;
}

// instruction descriptor, kind::i8: D = S32, A = B = unsigned 8-bit, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_i8(uint32_t n_cols)
{
  return (2u << 4) | ((n_cols >> 3) << 17) | ((kTileM >> 4) << 24);
}

struct GemmArgs
{
  const WorkItem* items;
  uint32_t n_items;
  uint32_t n;          // samples
  uint32_t n_pad16;    // n rounded up to 16 (UMMA N granularity)
  uint32_t rows_alloc; // rows of Pt (n rounded up to 128)
  int32_t* G;          // [bucket][n][n]
};

__global__ void __launch_bounds__(kThreads, 1) grm_umma_kernel(const __grid_constant__ CUtensorMap tmap, const GemmArgs A)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 tiles need 1024-byte alignment
  const uint32_t bars = smem0 + kStages * kStageBytes;
  // full[s] at bars + 8s, empty[s] at bars + 64 + 8s, tmem_full at +128, tmem_empty at +136, tmem base at +144
  const uint32_t bar_tfull = bars + 128, bar_tempty = bars + 136, tmem_slot = bars + 144;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0)
  {
    for (uint32_t s = 0; s < kStages; s++)
    {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 64 + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, kMaxN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0)
  {
    // ===== TMA producer =====
    if (lane == 0)
    {
      uint32_t it = 0;  // K blocks issued by this CTA
      for (uint32_t w = blockIdx.x; w < A.n_items; w += gridDim.x)
      {
        const WorkItem item = A.items[w];
        const uint32_t col0 = item.nhalf * kMaxN;
        const uint32_t ncols = min(kMaxN, A.n_pad16 - col0);
        const uint32_t b_chunks = (ncols + 127) / 128;
        for (uint32_t kb = 0; kb < item.kb_cnt; kb++, it++)
        {
          const uint32_t s = it % kStages, phase = (it / kStages) & 1;
          mbar_wait_bounded(bars + 64 + 8 * s, phase ^ 1);
          const uint32_t stage = smem0 + s * kStageBytes;
          mbar_expect_tx(bars + 8 * s, kChunkBytes * (1 + b_chunks));
          const int32_t x = static_cast<int32_t>((item.kb0 + kb) * kBlockK);
          tma_load_2d(stage, &tmap, bars + 8 * s, x, static_cast<int32_t>(item.mtile * kTileM));
          for (uint32_t c = 0; c < b_chunks; c++)
            tma_load_2d(stage + kChunkBytes * (1 + c), &tmap, bars + 8 * s, x, static_cast<int32_t>(col0 + 128 * c));
        }
      }
    }
  }
  else if (warp == 1)
  {
    // ===== MMA issuer: one thread =====
    if (lane == 0)
    {
      uint32_t it = 0, n_done = 0;
      for (uint32_t w = blockIdx.x; w < A.n_items; w += gridDim.x, n_done++)
      {
        const WorkItem item = A.items[w];
        const uint32_t col0 = item.nhalf * kMaxN;
        const uint32_t ncols = min(kMaxN, A.n_pad16 - col0);
        const uint32_t n1 = min(ncols, 256u), n2 = ncols - n1;
        const uint32_t idesc1 = umma_idesc_i8(n1), idesc2 = umma_idesc_i8(n2);
        mbar_wait_bounded(bar_tempty, (n_done & 1) ^ 1);  // epilogue drained the accumulator
        tc_fence_after();
        for (uint32_t kb = 0; kb < item.kb_cnt; kb++, it++)
        {
          const uint32_t s = it % kStages, phase = (it / kStages) & 1;
          mbar_wait_bounded(bars + 8 * s, phase);
          tc_fence_after();
          const uint32_t stage = smem0 + s * kStageBytes;
          const uint64_t a_desc = umma_desc_sw128(stage);
          const uint64_t b1_desc = umma_desc_sw128(stage + kChunkBytes);
          const uint64_t b2_desc = umma_desc_sw128(stage + kChunkBytes * 3);  // rows 256.. of the B tile
#pragma unroll
          for (uint32_t k = 0; k < kBlockK / kUmmaK; k++)
          {
            const uint64_t adv = (k * kUmmaK) >> 4;  // start-address field advances in 16-byte units
            const uint32_t acc = (kb | k) != 0;
            umma_i8(tmem_base, a_desc + adv, b1_desc + adv, idesc1, acc);
            if (n2) umma_i8(tmem_base + 256, a_desc + adv, b2_desc + adv, idesc2, acc);
          }
          umma_commit(bars + 64 + 8 * s);  // frees the smem stage when these MMAs have read it
        }
        umma_commit(bar_tfull);  // accumulator complete
      }
    }
  }
  else
  {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4).. =====
    const uint32_t q = warp & 3;
    uint32_t n_done = 0;
    for (uint32_t w = blockIdx.x; w < A.n_items; w += gridDim.x, n_done++)
    {
      const WorkItem item = A.items[w];
      const uint32_t col0 = item.nhalf * kMaxN;
      const uint32_t ncols = min(kMaxN, A.n_pad16 - col0);
      mbar_wait_bounded(bar_tfull, n_done & 1);
      tc_fence_after();
      const uint32_t row = item.mtile * kTileM + 32 * q + lane;
      int32_t* out = A.G + (static_cast<size_t>(item.bucket) * A.n + row) * A.n + col0;
      for (uint32_t c = 0; c < ncols; c += 16)
      {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((32 * q) << 16) + c;
        asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < A.n)
        {
#pragma unroll
          for (uint32_t j = 0; j < 16; j++)
            if (col0 + c + j < A.n && r[j] != 0) atomicAdd(out + c + j, static_cast<int32_t>(r[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kMaxN);
}

// ---- fp64 combination over the buckets, ascending carrier count -----------------------------------

struct BucketInfo
{
  double weight;   // f_c^2
  double mu;       // c / n
  double count;    // N_c (SNPs in the bucket)
};

// slot (may be null): G slot of bucket b (dense-by-carrier-count layouts); default b
__global__ void grm_combine_kernel(const int32_t* G, const BucketInfo* info, const uint32_t* slot, uint32_t n_buckets, uint32_t n,
                                   double* xtx)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = blockIdx.y;
  if (j >= n) return;
  double acc = 0.0;
  for (uint32_t b = 0; b < n_buckets; b++)
  {
    const int32_t* g = G + static_cast<size_t>(slot ? slot[b] : b) * n * n;
    const double gij = static_cast<double>(g[static_cast<size_t>(i) * n + j]);
    const double ri = static_cast<double>(g[static_cast<size_t>(i) * n + i]);
    const double rj = static_cast<double>(g[static_cast<size_t>(j) * n + j]);
    const BucketInfo bi = info[b];
    const double inner = gij - bi.mu * (ri + rj) + bi.count * bi.mu * bi.mu;
    acc += bi.weight * inner;
  }
  xtx[static_cast<size_t>(i) * n + j] = acc;
}

__global__ void add_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, size_t n)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] += src[i];
}

__global__ void add_u64_kernel(uint64_t* __restrict__ dst, const uint64_t* __restrict__ src, uint32_t n)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

__global__ void hist_to_u64_kernel(const uint32_t* hist, uint32_t n, uint64_t* out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hist[i];
}

}  // namespace grm
}  // namespace kmd
