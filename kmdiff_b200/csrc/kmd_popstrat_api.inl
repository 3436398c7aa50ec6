// C-ABI entry points of the --pop-correction path (included at the end of kmd_api.cu, after the helpers
// of its anonymous namespace).  Kernels: kmd_popstrat.cuh (sampler, presence, IRLS),
// kmd_grm.cuh (bucketed int8 tcgen05 Gram); host eigensolve: kmd_eig.hpp.

namespace {

void pop_free(kmd_context* ctx)
{
  PopState& ps = ctx->pop;
  cudaFree(ps.d_total); cudaFree(ps.d_done); cudaFree(ps.d_overflow); cudaFree(ps.d_slot_rows);
  cudaFree(ps.d_geno); cudaFree(ps.d_carriers); cudaFree(ps.d_surv_counts);
  cudaFree(ps.d_null_feat); cudaFree(ps.d_alt_feat); cudaFree(ps.d_y); cudaFree(ps.d_totals);
  for (void* q : ps.gram_scratch) cudaFree(q);
  ps = PopState{};
}

int pop_reset(kmd_context* ctx)
{
  PopState& ps = ctx->pop;
  if (!ps.enabled) return KMD_OK;
  KMD_CUDA(ctx, cudaMemsetAsync(ps.d_total, 0, sizeof(unsigned long long), ctx->stream));
  KMD_CUDA(ctx, cudaMemsetAsync(ps.d_done, 0, sizeof(unsigned long long), ctx->stream));
  KMD_CUDA(ctx, cudaMemsetAsync(ps.d_overflow, 0, sizeof(uint32_t), ctx->stream));
  return KMD_OK;
}

// number of sampled k-mers held (synchronises); KMD_ERR_OVERFLOW if the Sampler took more than max_sampled
int pop_sampled(kmd_context* ctx, uint64_t* m)
{
  PopState& ps = ctx->pop;
  unsigned long long done = 0;
  uint32_t ovf = 0;
  KMD_CUDA(ctx, cudaMemcpyAsync(&done, ps.d_done, sizeof done, cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(&ovf, ps.d_overflow, sizeof ovf, cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ovf)
    return fail(ctx, KMD_ERR_OVERFLOW, "sampler overflow: more than max_sampled = %llu k-mers were sampled",
                static_cast<unsigned long long>(ps.cap));
  *m = done;
  return KMD_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_presence_map(kmd_context* ctx, CUtensorMap* map, void* pt, uint64_t k_pad, uint64_t rows_alloc)
{
  static EncodeTiledFn fn = nullptr;
  if (!fn)
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    KMD_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) return fail(ctx, KMD_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t gdim[2] = {k_pad, rows_alloc};
  const cuuint64_t gstride[1] = {k_pad};  // bytes between rows
  const cuuint32_t box[2] = {grm::kBlockK, grm::kTileM};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, pt, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, KMD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return KMD_OK;
}

// getcolxz + fvadjust for a SNP with c carriers out of n (smartpca.c:2598-2700, 1684-1800; usenorm YES):
// mean c/n, scale 1/sqrt(p(1-p)) with p = 1 - sqrt(1 - mean) for diploids, else mean.
grm::BucketInfo bucket_info(uint32_t c, uint32_t n, bool diploid, uint32_t count)
{
  const double ymean = static_cast<double>(c) / static_cast<double>(n);
  double yfancy = 1.0;
  const double p = diploid ? 1 - std::sqrt(1 - ymean) : ymean;
  const double y = p * (1.0 - p);
  if (y > 0.0) yfancy = 1.0 / std::sqrt(y);
  // 0 < c < n always has y > 0; c == n gives a column of zeros whatever the scale
  grm::BucketInfo b;
  b.weight = yfancy * yfancy;
  b.mu = ymean;
  b.count = static_cast<double>(count);
  return b;
}

// non-owning view of one slot of PopState::gram_scratch
struct ScratchBuf
{
  void* p = nullptr;
  template <class T> T* as() { return static_cast<T*>(p); }
};

int gram_scratch(kmd_context* ctx, int slot, size_t bytes, ScratchBuf* out)
{
  PopState& ps = ctx->pop;
  if (bytes > ps.gram_scratch_cap[slot])
  {
    if (ps.gram_scratch[slot]) { KMD_CUDA(ctx, cudaFree(ps.gram_scratch[slot])); ps.gram_scratch[slot] = nullptr; ps.gram_scratch_cap[slot] = 0; }
    const size_t want = bytes + bytes / 8;
    KMD_CUDA(ctx, cudaMalloc(&ps.gram_scratch[slot], want));
    ps.gram_scratch_cap[slot] = want;
  }
  out->p = ps.gram_scratch[slot];
  return KMD_OK;
}

// dense_G == null: this context's XTX (integer Gram matrices of the carrier-count buckets + fp64 combination).
// dense_G != null (multi-GPU): only the integers — G_c of bucket c goes to dense_G + (c - 1) * n * n (caller-owned
// DEVICE memory of n^3 int32, zeroed here) and the carrier histogram to dense_hist[0..n] (u64) — for the ranks to sum.
int gram_impl(kmd_context* ctx, double* xtx, int normalize, uint64_t* n_used, int32_t* dense_G = nullptr, uint64_t* dense_hist = nullptr)
{
  PopState& ps = ctx->pop;
  const uint32_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  uint64_t m = 0;
  int rc = pop_sampled(ctx, &m);
  if (rc) return rc;
  if (xtx) std::fill(xtx, xtx + static_cast<size_t>(n) * n, 0.0);
  if (n_used) *n_used = 0;
  ps.gram_ms = 0;
  if (dense_G)
  {
    KMD_CUDA(ctx, cudaMemsetAsync(dense_G, 0, static_cast<size_t>(n) * n * n * sizeof(int32_t), ctx->stream));
    KMD_CUDA(ctx, cudaMemsetAsync(dense_hist, 0, (n + 1) * sizeof(uint64_t), ctx->stream));
  }
  if (m == 0) return KMD_OK;
  if (m >= (1ull << 32)) return fail(ctx, KMD_ERR_ARG, "more than 2^32 sampled k-mers");

  // ---- bucket the SNPs by carrier count ----
  ScratchBuf d_hist, d_koff, d_cursor, d_perm, d_pt, d_G, d_info, d_items, d_xtx;
  if ((rc = gram_scratch(ctx, 0, (n + 1) * sizeof(uint32_t), &d_hist))) return rc;
  KMD_CUDA(ctx, cudaMemsetAsync(d_hist.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
  grm::carriers_hist_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ps.d_carriers, m, d_hist.as<uint32_t>());
  std::vector<uint32_t> hist(n + 1);
  KMD_CUDA(ctx, cudaMemcpyAsync(hist.data(), d_hist.p, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

  const uint32_t rows_alloc = (n + grm::kTileM - 1) / grm::kTileM * grm::kTileM;
  const uint32_t n_pad16 = (n + 15) / 16 * 16;
  const uint32_t mtiles = rows_alloc / grm::kTileM;
  const uint32_t nhalves = (n_pad16 + grm::kMaxN - 1) / grm::kMaxN;
  std::vector<uint64_t> koff(n + 1, 0);
  std::vector<grm::BucketInfo> info;
  std::vector<grm::WorkItem> items;
  uint64_t k_pad = 0, used = 0;
  for (uint32_t c = 1; c <= n; c++)
  {
    koff[c] = k_pad;
    if (!hist[c]) continue;
    used += hist[c];
    const uint32_t bucket = dense_G ? c - 1 : static_cast<uint32_t>(info.size());
    info.push_back(bucket_info(c, n, ps.cfg.is_diploid != 0, hist[c]));
    const uint32_t kblocks = (hist[c] + grm::kBlockK - 1) / grm::kBlockK;
    for (uint32_t kb = 0; kb < kblocks; kb += grm::kMaxKBlocks)
      for (uint32_t mt = 0; mt < mtiles; mt++)
        for (uint32_t nh = 0; nh < nhalves; nh++)
        {
          grm::WorkItem w;
          w.bucket = bucket;
          w.kb0 = static_cast<uint32_t>(k_pad / grm::kBlockK) + kb;
          w.kb_cnt = std::min(grm::kMaxKBlocks, kblocks - kb);
          w.mtile = static_cast<uint16_t>(mt);
          w.nhalf = static_cast<uint16_t>(nh);
          items.push_back(w);
        }
    k_pad += static_cast<uint64_t>(kblocks) * grm::kBlockK;
  }
  if (n_used) *n_used = used;
  if (info.empty()) return KMD_OK;
  if (k_pad >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "presence matrix too wide for one tensor map (%llu columns)", static_cast<unsigned long long>(k_pad));
  // longest items first: the persistent CTAs take items round-robin
  std::stable_sort(items.begin(), items.end(), [](const grm::WorkItem& a, const grm::WorkItem& b) { return a.kb_cnt > b.kb_cnt; });

  const size_t g_bytes = dense_G ? 0 : info.size() * static_cast<size_t>(n) * n * sizeof(int32_t);
  if ((rc = gram_scratch(ctx, 1, (n + 1) * sizeof(uint64_t), &d_koff))) return rc;
  if ((rc = gram_scratch(ctx, 2, (n + 1) * sizeof(uint32_t), &d_cursor))) return rc;
  if ((rc = gram_scratch(ctx, 3, k_pad * sizeof(uint32_t), &d_perm))) return rc;
  if ((rc = gram_scratch(ctx, 4, static_cast<size_t>(rows_alloc) * k_pad, &d_pt))) return rc;
  if (!dense_G && (rc = gram_scratch(ctx, 5, g_bytes, &d_G))) return rc;
  if ((rc = gram_scratch(ctx, 6, info.size() * sizeof(grm::BucketInfo), &d_info))) return rc;
  if ((rc = gram_scratch(ctx, 7, items.size() * sizeof(grm::WorkItem), &d_items))) return rc;
  if ((rc = gram_scratch(ctx, 8, static_cast<size_t>(n) * n * sizeof(double), &d_xtx))) return rc;
  KMD_CUDA(ctx, cudaMemcpyAsync(d_koff.p, koff.data(), (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(d_info.p, info.data(), info.size() * sizeof(grm::BucketInfo), cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(grm::WorkItem), cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemsetAsync(d_cursor.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
  KMD_CUDA(ctx, cudaMemsetAsync(d_perm.p, 0xFF, k_pad * sizeof(uint32_t), ctx->stream));
  if (!dense_G) KMD_CUDA(ctx, cudaMemsetAsync(d_G.p, 0, g_bytes, ctx->stream));

  grm::carriers_scatter_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ps.d_carriers, m, d_koff.as<uint64_t>(),
                                                                         d_cursor.as<uint32_t>(), d_perm.as<uint32_t>());
  const uint32_t tgrid = static_cast<uint32_t>((k_pad / 4 + 255) / 256);
  grm::transpose_presence_kernel<<<tgrid, 256, 0, ctx->stream>>>(ps.d_geno, n, d_perm.as<uint32_t>(), k_pad, rows_alloc, d_pt.as<uint8_t>());
  KMD_CUDA(ctx, cudaGetLastError());

  // ---- exact integer Gram matrices on the tensor cores ----
  CUtensorMap tmap;
  if ((rc = encode_presence_map(ctx, &tmap, d_pt.p, k_pad, rows_alloc))) return rc;
  grm::GemmArgs G;
  G.items = d_items.as<grm::WorkItem>();
  G.n_items = static_cast<uint32_t>(items.size());
  G.n = n;
  G.n_pad16 = n_pad16;
  G.rows_alloc = rows_alloc;
  G.G = dense_G ? dense_G : d_G.as<int32_t>();
  KMD_CUDA(ctx, cudaFuncSetAttribute(grm::grm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(grm::kSmemBytes)));
  const uint32_t grid = std::min<uint32_t>(G.n_items, ctx->sm_count);
  cudaEventRecord(ctx->ev_k0, ctx->stream);
  grm::grm_umma_kernel<<<grid, grm::kThreads, grm::kSmemBytes, ctx->stream>>>(tmap, G);
  cudaEventRecord(ctx->ev_k1, ctx->stream);
  KMD_CUDA(ctx, cudaGetLastError());
  ctx->scan_timed = false;
  if (dense_G)
  {
    grm::hist_to_u64_kernel<<<(n + 256) / 256, 256, 0, ctx->stream>>>(d_hist.as<uint32_t>(), n + 1, dense_hist);
    ctx->launches += 5;
    KMD_CUDA(ctx, cudaGetLastError());
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KMD_CUDA(ctx, cudaEventElapsedTime(&ps.gram_ms, ctx->ev_k0, ctx->ev_k1));
    return KMD_OK;
  }

  // ---- fp64 combination, buckets in ascending carrier count ----
  const dim3 cgrid((n + 127) / 128, n);
  grm::grm_combine_kernel<<<cgrid, 128, 0, ctx->stream>>>(d_G.as<int32_t>(), d_info.as<grm::BucketInfo>(), nullptr,
                                                          static_cast<uint32_t>(info.size()), n, d_xtx.as<double>());
  ctx->launches += 5;
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaMemcpyAsync(xtx, d_xtx.p, static_cast<size_t>(n) * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  KMD_CUDA(ctx, cudaEventElapsedTime(&ps.gram_ms, ctx->ev_k0, ctx->ev_k1));
  if (normalize) kmd_smartpca_normalize(xtx, n);
  return KMD_OK;
}

// shared memory of one glm_irls_kernel CTA, and how many warps it can hold
size_t irls_smem(uint32_t n, uint32_t p, uint32_t ldf, bool feat_in_smem, uint32_t warps)
{
  return sizeof(double) * ((feat_in_smem ? static_cast<size_t>(n) * ldf : 0) + n + warps * pop::irls_warp_doubles(n, p));
}

int launch_irls(kmd_context* ctx, pop::IrlsArgs A, uint32_t n_items_host)
{
  const size_t budget = 226 * 1024;   // of the 227 KB a CTA may take
  uint32_t warps = pop::kIrlsWarps;
  bool in_smem = true;
  while (warps > 1 && irls_smem(A.n, A.p, A.ldf, in_smem, warps) > budget) warps--;
  if (irls_smem(A.n, A.p, A.ldf, in_smem, warps) > budget)
  {
    in_smem = false;
    warps = pop::kIrlsWarps;
    while (warps > 1 && irls_smem(A.n, A.p, A.ldf, in_smem, warps) > budget) warps--;
  }
  const size_t smem = irls_smem(A.n, A.p, A.ldf, in_smem, warps);
  if (smem > budget) return fail(ctx, KMD_ERR_ARG, "cohort of %u samples does not fit the GLM kernel's shared memory", A.n);
  A.feat_in_smem = in_smem;
  const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>((n_items_host + warps - 1) / warps, ctx->sm_count));
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kernel<<<blocks, warps * 32, smem, ctx->stream>>>(A);
    return cudaGetLastError();
  };
  // the feature count as a template parameter (null model: n_pc + 2, alternative model: n_pc + 3; --n-pc is 2..10)
  cudaError_t le;
  switch (in_smem ? A.p : 0u)
  {
    case 2: le = launch(pop::glm_irls_kernel<2>); break;
    case 3: le = launch(pop::glm_irls_kernel<3>); break;
    case 4: le = launch(pop::glm_irls_kernel<4>); break;
    case 5: le = launch(pop::glm_irls_kernel<5>); break;
    case 6: le = launch(pop::glm_irls_kernel<6>); break;
    case 7: le = launch(pop::glm_irls_kernel<7>); break;
    case 8: le = launch(pop::glm_irls_kernel<8>); break;
    case 9: le = launch(pop::glm_irls_kernel<9>); break;
    case 10: le = launch(pop::glm_irls_kernel<10>); break;
    case 11: le = launch(pop::glm_irls_kernel<11>); break;
    case 12: le = launch(pop::glm_irls_kernel<12>); break;
    case 13: le = launch(pop::glm_irls_kernel<13>); break;
    default: le = launch(pop::glm_irls_kernel<0>); break;
  }
  KMD_CUDA(ctx, le);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

pop::IrlsArgs alt_args(kmd_context* ctx)
{
  PopState& ps = ctx->pop;
  pop::IrlsArgs A{};
  A.n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  A.p = ps.alt_p;
  A.ldf = ps.alt_p;
  A.own_last = 1;
  A.features = ps.d_alt_feat;
  A.y = ps.d_y;
  A.totals = ps.d_totals;
  A.max_iter = 100;        // pop_strat_corrector::s_max_iter
  A.llr_epsilon = 1e-30;   // pop_strat_corrector::s_epsilon
  A.lgam_half = ctx->model.lgam_half;
  A.count_bytes = ctx->cfg.count_bytes;
  A.null_likelihood = ps.null_likelihood;
  return A;
}

}  // namespace

extern "C" {

int kmd_popstrat_enable(kmd_context* ctx, const kmd_popstrat_config* cfg)
{
  if (!ctx || !cfg) return KMD_ERR_ARG;
  if (ctx->rows_tested) return fail(ctx, KMD_ERR_STATE, "kmd_popstrat_enable must precede the first scan");
  if (cfg->n_pc > 10) return fail(ctx, KMD_ERR_ARG, "n_pc must be <= 10 (pcs.evec holds 10 components)");
  if (!(cfg->kmer_pca >= 0.0)) return fail(ctx, KMD_ERR_ARG, "kmer_pca must be >= 0");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  pop_free(ctx);
  PopState& ps = ctx->pop;
  ps.cfg = *cfg;
  ps.cap = cfg->max_sampled ? cfg->max_sampled : (1ull << 21);
  ps.state0 = pop::lcg_seed(cfg->seed);
  const uint64_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  KMD_CUDA(ctx, cudaMalloc(&ps.d_total, sizeof(unsigned long long)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_done, sizeof(unsigned long long)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_overflow, sizeof(uint32_t)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_slot_rows, ps.cap * sizeof(uint64_t)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_geno, ps.cap * n));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_carriers, ps.cap * sizeof(uint32_t)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_surv_counts, ctx->cap * n * ctx->cfg.count_bytes));
  ps.enabled = true;
  return pop_reset(ctx);
}

int kmd_popstrat_set_draw_index(kmd_context* ctx, uint64_t next_draw)
{
  if (!ctx) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  ctx->pop.draw_override = true;
  ctx->pop.draw_next = next_draw;
  return KMD_OK;
}

int kmd_popstrat_sampled(kmd_context* ctx, uint64_t* n)
{
  if (!ctx || !n) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return pop_sampled(ctx, n);
}

int kmd_popstrat_geno(kmd_context* ctx, uint8_t* geno, uint64_t* rows, uint64_t cap, uint64_t* n_out)
{
  if (!ctx || !n_out) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  PopState& ps = ctx->pop;
  uint64_t m = 0;
  int rc = pop_sampled(ctx, &m);
  if (rc) return rc;
  *n_out = m;
  if (m == 0 || (!geno && !rows)) return KMD_OK;
  if (cap < m) return fail(ctx, KMD_ERR_ARG, "output capacity %llu < %llu sampled k-mers", static_cast<unsigned long long>(cap), static_cast<unsigned long long>(m));
  const uint64_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  std::vector<uint64_t> slot_rows(m);
  KMD_CUDA(ctx, cudaMemcpy(slot_rows.data(), ps.d_slot_rows, m * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> order(m);
  for (uint64_t i = 0; i < m; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return slot_rows[a] < slot_rows[b]; });
  if (rows) for (uint64_t i = 0; i < m; i++) rows[i] = slot_rows[order[i]];
  if (geno)
  {
    std::vector<uint8_t> raw(m * n);
    KMD_CUDA(ctx, cudaMemcpy(raw.data(), ps.d_geno, m * n, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < m; i++) memcpy(geno + i * n, raw.data() + order[i] * n, n);
  }
  return KMD_OK;
}

int kmd_popstrat_gram(kmd_context* ctx, double* xtx, int normalize, uint64_t* n_used)
{
  if (!ctx || !xtx) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return gram_impl(ctx, xtx, normalize, n_used);
}

int kmd_popstrat_gram_partial(kmd_context* ctx, int32_t* d_G, uint64_t* d_hist)
{
  if (!ctx || !d_G || !d_hist) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return gram_impl(ctx, nullptr, 0, nullptr, d_G, d_hist);
}

int kmd_popstrat_gram_combine(kmd_context* ctx, const int32_t* d_G, const uint64_t* d_hist, double* xtx, int normalize, uint64_t* n_used)
{
  if (!ctx || !d_G || !d_hist || !xtx) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  std::vector<uint64_t> hist(n + 1);
  KMD_CUDA(ctx, cudaMemcpyAsync(hist.data(), d_hist, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<grm::BucketInfo> info;
  std::vector<uint32_t> slot;
  uint64_t used = 0;
  for (uint32_t c = 1; c <= n; c++)
  {
    if (!hist[c]) continue;
    if (hist[c] >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "carrier-count bucket %u overflows int32 (%llu SNPs)", c, static_cast<unsigned long long>(hist[c]));
    used += hist[c];
    info.push_back(bucket_info(c, n, ctx->pop.cfg.is_diploid != 0, static_cast<uint32_t>(hist[c])));
    slot.push_back(c - 1);
  }
  if (n_used) *n_used = used;
  std::fill(xtx, xtx + static_cast<size_t>(n) * n, 0.0);
  if (info.empty()) return KMD_OK;
  ScratchBuf d_info, d_slot, d_xtx;
  int rc;
  if ((rc = gram_scratch(ctx, 6, info.size() * sizeof(grm::BucketInfo), &d_info))) return rc;
  if ((rc = gram_scratch(ctx, 2, std::max<size_t>((n + 1), slot.size()) * sizeof(uint32_t), &d_slot))) return rc;
  if ((rc = gram_scratch(ctx, 8, static_cast<size_t>(n) * n * sizeof(double), &d_xtx))) return rc;
  KMD_CUDA(ctx, cudaMemcpyAsync(d_info.p, info.data(), info.size() * sizeof(grm::BucketInfo), cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(d_slot.p, slot.data(), slot.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  const dim3 cgrid((n + 127) / 128, n);
  grm::grm_combine_kernel<<<cgrid, 128, 0, ctx->stream>>>(d_G, d_info.as<grm::BucketInfo>(), d_slot.as<uint32_t>(),
                                                          static_cast<uint32_t>(info.size()), n, d_xtx.as<double>());
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaMemcpyAsync(xtx, d_xtx.p, static_cast<size_t>(n) * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (normalize) kmd_smartpca_normalize(xtx, n);
  return KMD_OK;
}

int kmd_popstrat_gram_local(kmd_context* const* ctxs, uint32_t n_ctx, double* xtx, int normalize, uint64_t* n_used)
{
  if (!ctxs || !n_ctx || !xtx) return KMD_ERR_ARG;
  for (uint32_t i = 0; i < n_ctx; i++)
    if (!ctxs[i] || !ctxs[i]->pop.enabled) return KMD_ERR_ARG;
  kmd_context* c0 = ctxs[0];
  if (n_ctx == 1) return kmd_popstrat_gram(c0, xtx, normalize, n_used);
  const uint32_t n = c0->cfg.nb_controls + c0->cfg.nb_cases;
  const size_t g_words = static_cast<size_t>(n) * n * n;
  if (g_words * sizeof(int32_t) > (2ull << 30))
    return fail(c0, KMD_ERR_ARG, "%u samples: the dense integer Gram exchange would need %zu MiB per device", n, g_words * 4 >> 20);
  // every context: its integer co-carrier matrices by carrier count + its carrier histogram, on its own device
  std::vector<DeviceBuf> G(n_ctx), H(n_ctx);
  for (uint32_t i = 0; i < n_ctx; i++)
  {
    kmd_context* c = ctxs[i];
    KMD_CUDA(c, cudaSetDevice(c->device));
    KMD_CUDA(c, cudaMalloc(&G[i].p, g_words * sizeof(int32_t)));
    KMD_CUDA(c, cudaMalloc(&H[i].p, (n + 1) * sizeof(uint64_t)));
    int rc = gram_impl(c, nullptr, 0, nullptr, G[i].as<int32_t>(), H[i].as<uint64_t>());   // synchronises c's stream
    if (rc) return rc;
  }
  // exact sum on device 0 (peer copies; the integers make the order irrelevant), then one fp64 combination
  KMD_CUDA(c0, cudaSetDevice(c0->device));
  DeviceBuf tg, th;
  KMD_CUDA(c0, cudaMalloc(&tg.p, g_words * sizeof(int32_t)));
  KMD_CUDA(c0, cudaMalloc(&th.p, (n + 1) * sizeof(uint64_t)));
  for (uint32_t i = 1; i < n_ctx; i++)
  {
    KMD_CUDA(c0, cudaMemcpyPeerAsync(tg.p, c0->device, G[i].p, ctxs[i]->device, g_words * sizeof(int32_t), c0->stream));
    KMD_CUDA(c0, cudaMemcpyPeerAsync(th.p, c0->device, H[i].p, ctxs[i]->device, (n + 1) * sizeof(uint64_t), c0->stream));
    grm::add_i32_kernel<<<c0->sm_count * 8, 256, 0, c0->stream>>>(G[0].as<int32_t>(), tg.as<int32_t>(), g_words);
    grm::add_u64_kernel<<<(n + 256) / 256, 256, 0, c0->stream>>>(H[0].as<uint64_t>(), th.as<uint64_t>(), n + 1);
    c0->launches += 2;
    KMD_CUDA(c0, cudaGetLastError());
  }
  int rc = kmd_popstrat_gram_combine(c0, G[0].as<int32_t>(), H[0].as<uint64_t>(), xtx, normalize, n_used);
  // the buffers of the other devices are freed on their own devices
  for (uint32_t i = 1; i < n_ctx; i++)
  {
    cudaSetDevice(ctxs[i]->device);
    cudaFree(G[i].p); G[i].p = nullptr;
    cudaFree(H[i].p); H[i].p = nullptr;
  }
  cudaSetDevice(c0->device);
  return rc;
}

int kmd_popstrat_gram_ms(kmd_context* ctx, float* ms)
{
  if (!ctx || !ms) return KMD_ERR_ARG;
  *ms = ctx->pop.gram_ms;
  return KMD_OK;
}

int kmd_popstrat_init(kmd_context* ctx, const double* Z, const uint64_t* totals, double* null_model)
{
  if (!ctx || !Z || !totals) return KMD_ERR_ARG;
  if (!ctx->pop.enabled) return fail(ctx, KMD_ERR_STATE, "population stratification is not enabled on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  PopState& ps = ctx->pop;
  const uint32_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  const uint32_t npc = ps.cfg.n_pc;
  // init_global_features (popstrat.cpp:270-313) without covariates and with every gender unknown
  // (load_ginfo never recognises one, popstrat.cpp:236-268): [1, PC_1..PC_npc, total_i (, ratio)]
  const uint32_t pc = 1 + npc + 1, pa = pc + 1;
  if (n < pc) return fail(ctx, KMD_ERR_ARG, "standardize() needs at least n_pc + 2 samples");  // it indexes stddev[ncols-1] of n entries
  std::vector<double> nf(static_cast<size_t>(n) * pc, 0.0), af(static_cast<size_t>(n) * pa, 0.0), y(n), tot(n);
  for (uint32_t i = 0; i < n; i++)
  {
    y[i] = i < ctx->cfg.nb_controls ? 1.0 : 0.0;  // load_Y: "Case" -> 0, else 1 (popstrat.cpp:164-171)
    tot[i] = static_cast<double>(totals[i]);
    nf[static_cast<size_t>(i) * pc] = 1;
    af[static_cast<size_t>(i) * pa] = 1;
    for (uint32_t z = 0; z < npc; z++)
    {
      nf[static_cast<size_t>(i) * pc + z + 1] = Z[static_cast<size_t>(i) * 10 + z];
      af[static_cast<size_t>(i) * pa + z + 1] = Z[static_cast<size_t>(i) * 10 + z];
    }
    nf[static_cast<size_t>(i) * pc + 1 + npc] = tot[i];
    af[static_cast<size_t>(i) * pa + 1 + npc] = tot[i];
  }
  // standardize() (popstrat.cpp:327-370) as written: column means are divided by the number of
  // COLUMNS, the deviations are indexed by column but applied by ROW, so only rows 1..ncols-1 change.
  {
    std::vector<double> means(pc, 0.0), stddev(n, 0.0);
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = 0; j < pc; j++) means[j] += nf[static_cast<size_t>(i) * pc + j];
    for (uint32_t j = 1; j < pc; j++) means[j] /= pc;
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = 1; j < pc; j++)
      {
        const double d = nf[static_cast<size_t>(i) * pc + j] - means[j];
        stddev[j] += d * d;
      }
    for (uint32_t j = 1; j < pc; j++) stddev[j] = std::sqrt(stddev[j] / n);
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = 1; j < pc; j++)
        if (std::fabs(stddev[i]) > 1e-305)
        {
          nf[static_cast<size_t>(i) * pc + j] = (nf[static_cast<size_t>(i) * pc + j] - means[j]) / stddev[i];
          af[static_cast<size_t>(i) * pa + j] = (af[static_cast<size_t>(i) * pa + j] - means[j]) / stddev[i];
        }
  }
  cudaFree(ps.d_null_feat); cudaFree(ps.d_alt_feat); cudaFree(ps.d_y); cudaFree(ps.d_totals);
  ps.d_null_feat = ps.d_alt_feat = ps.d_y = ps.d_totals = nullptr;
  ps.glm_ready = false;
  KMD_CUDA(ctx, cudaMalloc(&ps.d_null_feat, nf.size() * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_alt_feat, af.size() * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_y, n * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&ps.d_totals, n * sizeof(double)));
  KMD_CUDA(ctx, cudaMemcpy(ps.d_null_feat, nf.data(), nf.size() * sizeof(double), cudaMemcpyHostToDevice));
  KMD_CUDA(ctx, cudaMemcpy(ps.d_alt_feat, af.data(), af.size() * sizeof(double), cudaMemcpyHostToDevice));
  KMD_CUDA(ctx, cudaMemcpy(ps.d_y, y.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  KMD_CUDA(ctx, cudaMemcpy(ps.d_totals, tot.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  ps.null_p = pc;
  ps.alt_p = pa;

  // m_null_model = glm_irls(null features, Y) (popstrat.cpp:315-323) and the likelihood of that model,
  // which apply() recomputes for every k-mer from the same inputs (popstrat.hpp:291-311)
  DeviceBuf d_model, d_lik;
  KMD_CUDA(ctx, cudaMalloc(&d_model.p, pc * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&d_lik.p, sizeof(double)));
  pop::IrlsArgs A = alt_args(ctx);
  A.p = pc;
  A.ldf = pc;
  A.own_last = 0;
  A.features = ps.d_null_feat;
  A.n_items = 1;
  A.null_likelihood = 1.0;
  A.model = d_model.as<double>();
  A.likelihood = d_lik.as<double>();
  int rc = launch_irls(ctx, A, 1);
  if (rc) return rc;
  ps.null_model.resize(pc);
  KMD_CUDA(ctx, cudaMemcpyAsync(ps.null_model.data(), d_model.p, pc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(&ps.null_likelihood, d_lik.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (null_model) memcpy(null_model, ps.null_model.data(), pc * sizeof(double));
  ps.glm_ready = true;
  return KMD_OK;
}

int kmd_popstrat_apply(kmd_context* ctx)
{
  if (!ctx) return KMD_ERR_ARG;
  if (!ctx->pop.glm_ready) return fail(ctx, KMD_ERR_STATE, "kmd_popstrat_init has not been called");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = fetch_counters(ctx);
  if (rc) return rc;
  const uint32_t n_surv = ctx->h_cnt->surv_count;
  if (n_surv == 0) return KMD_OK;
  pop::IrlsArgs A = alt_args(ctx);
  A.n_items = n_surv;
  A.counts = ctx->pop.d_surv_counts;
  A.p_value = ctx->surv.p;  // ks.set_pval(corrected) (popstrat.hpp:330)
  if ((rc = launch_irls(ctx, A, n_surv))) return rc;
  // d_vals[0] holds the p order in state 2 (ensure_order overwrote the row order with it): both orders are rebuilt
  if (ctx->order_state != 1) ctx->order_state = 0;
  ctx->final_valid = false;
  return KMD_OK;
}

int kmd_popstrat_fit(kmd_context* ctx, const void* counts, uint64_t n_kmers, double* p, double* model, double* llr, int32_t* iters)
{
  if (!ctx || (!counts && n_kmers)) return KMD_ERR_ARG;
  if (!ctx->pop.glm_ready) return fail(ctx, KMD_ERR_STATE, "kmd_popstrat_init has not been called");
  if (n_kmers == 0) return KMD_OK;
  if (n_kmers >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "too many k-mers in one call");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  PopState& ps = ctx->pop;
  const uint32_t n = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  const size_t cbytes = n_kmers * n * ctx->cfg.count_bytes;
  DeviceBuf d_counts, d_p, d_model, d_llr, d_iters;
  KMD_CUDA(ctx, cudaMalloc(&d_counts.p, cbytes));
  KMD_CUDA(ctx, cudaMalloc(&d_p.p, n_kmers * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&d_model.p, n_kmers * ps.alt_p * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&d_llr.p, n_kmers * sizeof(double)));
  KMD_CUDA(ctx, cudaMalloc(&d_iters.p, n_kmers * sizeof(int32_t)));
  KMD_CUDA(ctx, cudaMemcpyAsync(d_counts.p, counts, cbytes, cudaMemcpyHostToDevice, ctx->stream));
  pop::IrlsArgs A = alt_args(ctx);
  A.n_items = static_cast<uint32_t>(n_kmers);
  A.counts = d_counts.as<uint8_t>();
  A.p_value = d_p.as<double>();
  A.model = d_model.as<double>();
  A.llr = d_llr.as<double>();
  A.iters = d_iters.as<int32_t>();
  int rc = launch_irls(ctx, A, A.n_items);
  if (rc) return rc;
  if (p) KMD_CUDA(ctx, cudaMemcpyAsync(p, d_p.p, n_kmers * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (model) KMD_CUDA(ctx, cudaMemcpyAsync(model, d_model.p, n_kmers * ps.alt_p * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (llr) KMD_CUDA(ctx, cudaMemcpyAsync(llr, d_llr.p, n_kmers * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (iters) KMD_CUDA(ctx, cudaMemcpyAsync(iters, d_iters.p, n_kmers * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KMD_OK;
}

int kmd_survivor_counts(kmd_context* ctx, void* out, uint64_t cap, uint64_t* n_out)
{
  if (!ctx || !n_out) return KMD_ERR_ARG;
  if (!ctx->pop.d_surv_counts)
    return fail(ctx, KMD_ERR_STATE, "survivor counts are kept only with KMD_FLAG_KEEP_COUNTS or population stratification enabled");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_order(ctx, 1);
  if (rc) return rc;
  const uint32_t n = ctx->h_cnt->surv_count;
  *n_out = n;
  if (n == 0 || !out) return KMD_OK;
  if (cap < n) return fail(ctx, KMD_ERR_ARG, "output capacity %llu < %u records", static_cast<unsigned long long>(cap), n);
  const size_t rec = static_cast<size_t>(ctx->cfg.nb_controls + ctx->cfg.nb_cases) * ctx->cfg.count_bytes;
  std::vector<uint32_t> order(n);
  std::vector<uint8_t> raw(n * rec);
  KMD_CUDA(ctx, cudaMemcpyAsync(order.data(), ctx->d_vals[0], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(raw.data(), ctx->pop.d_surv_counts, n * rec, cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (uint32_t i = 0; i < n; i++) memcpy(static_cast<uint8_t*>(out) + i * rec, raw.data() + order[i] * rec, rec);
  return KMD_OK;
}

// ---- host-only helpers ---------------------------------------------------------------------------

void kmd_smartpca_normalize(double* xtx, uint32_t n)
{
  // smartpca.c:1019-1023: y = trace / (n - 1); XTX *= 1 / y
  double tr = 0.0;
  for (uint32_t i = 0; i < n; i++) tr += xtx[static_cast<size_t>(i) * n + i];
  const double y = tr / static_cast<double>(n - 1);
  const double inv = 1.0 / y;
  for (size_t i = 0; i < static_cast<size_t>(n) * n; i++) xtx[i] = xtx[i] * inv;
}

int kmd_sym_eig(const double* a, uint32_t n, double* evals, double* evecs)
{
  if (!a || !evals || !evecs) return KMD_ERR_ARG;
  return kmd::sym_eig(a, n, evals, evecs) ? KMD_ERR_STATE : KMD_OK;
}

int kmd_smartpca_pcs(const double* xtx, uint32_t n, uint32_t numeigs, double* evals_out, double* pcs)
{
  if (!xtx || !pcs || numeigs > n) return KMD_ERR_ARG;
  // all eigenvalues, but only the numeigs leading eigenvectors (inverse iteration: O(n^2) each instead of accumulating
  // all n through the QL sweeps)
  std::vector<double> evals(n), evecs(static_cast<size_t>(std::max<uint32_t>(numeigs, 1)) * n);
  if (kmd::sym_eig_top(xtx, n, numeigs, evals.data(), evecs.data())) return KMD_ERR_STATE;
  if (evals_out) memcpy(evals_out, evals.data(), n * sizeof(double));
  // fxvecs[j] = X X^T fvecs[j] (smartpca.c:1186-1228: ffvecs = X^T fvecs, then one dot product per
  // individual), rescaled to unit length (fxscal, :1230-1240) and written "%10.4f" (:1296-1329)
  std::vector<double> yv(n);
  for (uint32_t j = 0; j < numeigs; j++)
  {
    const double* e = &evecs[static_cast<size_t>(j) * n];
    double norm2 = 0.0;
    for (uint32_t i = 0; i < n; i++)
    {
      double s = 0.0;
      for (uint32_t k = 0; k < n; k++) s += xtx[static_cast<size_t>(i) * n + k] * e[k];
      yv[i] = s;
      norm2 += s * s;
    }
    const double scal = norm2 > 0 ? 1.0 / std::sqrt(norm2) : 0.0;
    for (uint32_t i = 0; i < n; i++)
    {
      char buf[64];
      snprintf(buf, sizeof buf, "%10.4f", scal * yv[i]);
      pcs[static_cast<size_t>(i) * numeigs + j] = strtod(buf, nullptr);
    }
  }
  return KMD_OK;
}

uint64_t kmd_sampler_rows(uint64_t seed, double kmer_pca, uint64_t first_row, uint64_t n, uint8_t* out)
{
  uint32_t x = pop::lcg_mul(pop::lcg_seed(seed), pop::lcg_pow(pop::kLcgA, 2ull * (first_row % (pop::kLcgM - 1u))));
  uint64_t taken = 0;
  for (uint64_t i = 0; i < n; i++)
  {
    const uint32_t x1 = pop::lcg_mul(x, pop::kLcgA);
    const uint32_t x2 = pop::lcg_mul(x1, pop::kLcgA);
    x = x2;
    const bool take = pop::canonical_draw(x1, x2) < kmer_pca;
    if (out) out[i] = take;
    taken += take;
  }
  return taken;
}

}  // extern "C"
