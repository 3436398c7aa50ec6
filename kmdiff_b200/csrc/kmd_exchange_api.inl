// C-ABI entry points of the multi-GPU correction (included at the end of kmd_api.cu).
//
// kmtricks partitions are sharded over ranks (one process per GPU).  The reference's only cross-partition
// steps are the sum of the per-partition row counts = m of the correctors (include/kmdiff/merge.hpp:316,
// 390-397) and, for BENJAMINI / HOLM, ONE ascending walk over the survivors of all partitions
// (include/kmdiff/aggregator.hpp:282-310, src/corrector.cpp:27-35, 68-71).  Here every rank
//   1. kmd_exchange_export:  writes its slot {header: rows, controls, cases, survivors | cap p-values, +inf padded}
//      on its stream, without knowing the survivor count on the host,
//   2. (caller) all-gathers the slots — NCCL over NVLink, one collective, ordered by events against this stream,
//   3. kmd_exchange_correct: sums the headers (m), sorts the union of the p-values, finds the first rejected rank
//      of the union and keeps its own survivors below that p-value — all enqueued, nothing read back,
//   4. kmd_exchange_result:  the one host synchronisation of the step.
// Every rank computes the same decision from the same gathered bytes.

namespace {

void exchange_free(kmd_context* ctx)
{
  ExchangeState& x = ctx->xch;
  cudaFree(x.keys[0]);
  cudaFree(x.keys[1]);
  cudaFree(x.d_dec);
  cudaFree(x.gather);
  if (x.ev_export) cudaEventDestroy(x.ev_export);
  if (x.h_dec) cudaFreeHost(x.h_dec);
  x = ExchangeState{};
}

int exchange_ensure(kmd_context* ctx, size_t n_keys)
{
  ExchangeState& x = ctx->xch;
  if (!x.d_dec)
  {
    KMD_CUDA(ctx, cudaMalloc(&x.d_dec, sizeof(ExchangeDecision)));
    KMD_CUDA(ctx, cudaMallocHost(&x.h_dec, sizeof(ExchangeDecision)));
    memset(x.h_dec, 0, sizeof(ExchangeDecision));
  }
  if (n_keys > x.keys_cap)
  {
    for (int b = 0; b < 2; b++) { if (x.keys[b]) cudaFree(x.keys[b]); x.keys[b] = nullptr; }
    x.keys_cap = 0;
    const size_t want = n_keys + n_keys / 4;
    KMD_CUDA(ctx, cudaMalloc(&x.keys[0], want * sizeof(uint64_t)));
    KMD_CUDA(ctx, cudaMalloc(&x.keys[1], want * sizeof(uint64_t)));
    x.keys_cap = want;
  }
  return KMD_OK;
}

__global__ void __launch_bounds__(256) exchange_export_kernel(const DeviceCounters* cnt, uint64_t rows_tested, const double* p,
                                                              uint64_t* slot, uint32_t cap)
{
  const uint32_t n = cnt->surv_count;
  if (blockIdx.x == 0 && threadIdx.x < KMD_EXCHANGE_HEADER)
  {
    uint64_t v = 0;
    switch (threadIdx.x)
    {
      case 0: v = rows_tested; break;
      case 1: v = cnt->n_control; break;
      case 2: v = cnt->n_case; break;
      case 3: v = n; break;
      case 4: v = (n > cap || cnt->overflow) ? 1 : 0; break;
      default: break;
    }
    slot[threadIdx.x] = v;
  }
  double* out = reinterpret_cast<double*>(slot + KMD_EXCHANGE_HEADER);
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
    out[i] = i < n ? p[i] : inf;
}

// headers -> m, survivor totals, overflow; for the order-free correctors also the cut itself
__global__ void exchange_header_kernel(const uint64_t* all, uint32_t world, uint64_t slot_words, int correction, double significance,
                                       double host_cut, ExchangeDecision* dec)
{
  if (threadIdx.x || blockIdx.x) return;
  unsigned long long rows = 0, surv = 0, mx = 0;
  uint32_t overflow = 0;
  for (uint32_t r = 0; r < world; r++)
  {
    const uint64_t* h = all + r * slot_words;
    rows += h[0];
    surv += h[3];
    mx = max(mx, static_cast<unsigned long long>(h[3]));
    overflow |= h[4] ? 1u : 0u;
  }
  dec->total_rows = rows;
  dec->total_survivors = surv;
  dec->max_rank_survivors = mx;
  dec->overflow = overflow;
  dec->n_accepted = 0;
  switch (correction)
  {
    case KMD_CORR_BONFERRONI: dec->p_cut = __ddiv_rn(significance, static_cast<double>(rows)); break;   // corrector.cpp:11
    case KMD_CORR_SIDAK: dec->p_cut = host_cut; break;   // pow() is evaluated on the host (corrector.cpp:52)
    case KMD_CORR_NOTHING: dec->p_cut = significance; break;                                            // corrector.cpp:88
    default: dec->p_cut = 0; break;                       // BENJAMINI / HOLM: set by exchange_cut_kernel
  }
}

__global__ void __launch_bounds__(256) exchange_keys_kernel(const uint64_t* all, uint32_t world, uint32_t cap, uint64_t slot_words,
                                                            uint64_t* keys, uint32_t* first)
{
  const uint32_t n = world * cap;
  if (blockIdx.x == 0 && threadIdx.x == 0) *first = n;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    keys[i] = all[(i / cap) * slot_words + KMD_EXCHANGE_HEADER + i % cap];
}

__global__ void exchange_cut_kernel(const double* p_sorted, uint32_t n, ExchangeDecision* dec)
{
  if (threadIdx.x || blockIdx.x) return;
  const uint32_t first = dec->first;
  // +inf padding never passes a rank test, so first <= total_survivors whenever no slot overflowed
  dec->n_accepted = min(static_cast<unsigned long long>(first), dec->total_survivors);
  dec->p_cut = first < n ? p_sorted[first] : __longlong_as_double(0x7FF0000000000000ll);
}

}  // namespace

extern "C" {

void* kmd_get_stream(kmd_context* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int kmd_exchange_export(kmd_context* ctx, uint64_t* d_slot, uint64_t cap)
{
  if (!ctx || !d_slot) return KMD_ERR_ARG;
  if (cap == 0 || cap >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "exchange capacity must be in [1, 2^31)");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t grid = static_cast<uint32_t>(std::min<uint64_t>((cap + 255) / 256, static_cast<uint64_t>(ctx->sm_count) * 8));
  exchange_export_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_cnt, ctx->rows_tested, ctx->surv.p, d_slot, static_cast<uint32_t>(cap));
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

int kmd_exchange_result(kmd_context* ctx, kmd_exchange_decision* out);

int kmd_exchange_correct(kmd_context* ctx, const uint64_t* d_all, uint32_t world, uint64_t cap, int correction, double significance)
{
  if (!ctx || !d_all || !world) return KMD_ERR_ARG;
  if (correction < KMD_CORR_NOTHING || correction > KMD_CORR_HOLM) return fail(ctx, KMD_ERR_ARG, "unknown correction %d", correction);
  const uint64_t n64 = static_cast<uint64_t>(world) * cap;
  if (cap == 0 || n64 >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "world x capacity must be in [1, 2^31)");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool sorted = correction == KMD_CORR_BENJAMINI || correction == KMD_CORR_HOLM;
  int rc = exchange_ensure(ctx, sorted ? n64 : 0);
  if (rc) return rc;
  ExchangeState& x = ctx->xch;
  const uint64_t slot_words = KMD_EXCHANGE_HEADER + cap;
  const uint32_t n = static_cast<uint32_t>(n64);

  double host_cut = 0;
  if (correction == KMD_CORR_SIDAK)
  {
    // 1 - pow(1 - t, 1/m) must be the host libm's pow (parity with corrector.cpp:52): m is needed here, one read-back
    std::vector<uint64_t> head(static_cast<size_t>(world) * KMD_EXCHANGE_HEADER);
    KMD_CUDA(ctx, cudaMemcpy2DAsync(head.data(), KMD_EXCHANGE_HEADER * 8, d_all, slot_words * 8, KMD_EXCHANGE_HEADER * 8, world,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t m = 0;
    for (uint32_t r = 0; r < world; r++) m += head[static_cast<size_t>(r) * KMD_EXCHANGE_HEADER];
    host_cut = 1 - std::pow(1 - significance, 1.0 / static_cast<double>(m));
  }
  exchange_header_kernel<<<1, 32, 0, ctx->stream>>>(d_all, world, slot_words, correction, significance, host_cut, x.d_dec);
  ctx->launches++;
  if (sorted)
  {
    const uint32_t grid = std::min<uint32_t>((n + 255) / 256, ctx->sm_count * 8);
    exchange_keys_kernel<<<grid, 256, 0, ctx->stream>>>(d_all, world, static_cast<uint32_t>(cap), slot_words, x.keys[0], &x.d_dec->first);
    ctx->launches++;
    int cur = 0;
    // p >= 0, so the order of the bit patterns is the numeric order; +inf padding sorts last
    if ((rc = radix_sort_buffers(ctx, x.keys, nullptr, n, 8, &cur))) return rc;
    const double* p_sorted = reinterpret_cast<const double*>(x.keys[cur]);
    first_reject_dev_kernel<<<blocks_for(n), 256, 0, ctx->stream>>>(p_sorted, n, correction == KMD_CORR_HOLM, significance,
                                                                    reinterpret_cast<const uint64_t*>(&x.d_dec->total_rows), &x.d_dec->first);
    exchange_cut_kernel<<<1, 32, 0, ctx->stream>>>(p_sorted, n, x.d_dec);
    ctx->launches += 2;
  }
  KMD_CUDA(ctx, cudaGetLastError());
  // every rank keeps its own survivors below the cut (strict <, like every corrector)
  KMD_CUDA(ctx, cudaMemsetAsync(&ctx->d_cnt->keep_control, 0, 3 * sizeof(uint32_t), ctx->stream));
  keep_unordered_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->d_cnt, static_cast<uint32_t>(ctx->cap), ctx->surv.p, 0.0,
                                                                     &x.d_dec->p_cut, ctx->surv.sign, ctx->d_order);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaMemcpyAsync(x.h_dec, x.d_dec, sizeof(ExchangeDecision), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, ctx->stream));
  x.pending = true;
  x.by_p = sorted;
  ctx->final_valid = false;
  return KMD_OK;
}

int kmd_exchange_local(kmd_context* const* ctxs, uint32_t n, int correction, double significance, kmd_exchange_decision* decisions)
{
  if (!ctxs || !n || !decisions) return KMD_ERR_ARG;
  for (uint32_t i = 0; i < n; i++)
    if (!ctxs[i]) return KMD_ERR_ARG;
  // slot capacity = the largest survivor count (this entry point synchronises anyway)
  uint64_t cap = 1;
  for (uint32_t i = 0; i < n; i++)
  {
    kmd_context* c = ctxs[i];
    KMD_CUDA(c, cudaSetDevice(c->device));
    int rc = fetch_counters(c);
    if (rc) return rc;
    cap = std::max<uint64_t>(cap, c->h_cnt->surv_count);
  }
  const uint64_t slot_words = KMD_EXCHANGE_HEADER + cap;
  // every context writes its slot into its OWN gather buffer ...
  for (uint32_t i = 0; i < n; i++)
  {
    kmd_context* c = ctxs[i];
    ExchangeState& x = c->xch;
    KMD_CUDA(c, cudaSetDevice(c->device));
    if (x.gather_words < slot_words * n)
    {
      if (x.gather) KMD_CUDA(c, cudaFree(x.gather));
      x.gather = nullptr;
      x.gather_words = 0;
      const size_t want = slot_words * n + slot_words * n / 4;
      KMD_CUDA(c, cudaMalloc(&x.gather, want * sizeof(uint64_t)));
      x.gather_words = want;
    }
    if (!x.ev_export) KMD_CUDA(c, cudaEventCreateWithFlags(&x.ev_export, cudaEventDisableTiming));
    for (uint32_t j = 0; j < n; j++)
      if (j != i && ctxs[j]->device != c->device)
      {
        const cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);   // NVLink P2P when the topology has it
        if (e != cudaSuccess) cudaGetLastError();                               // already enabled / unsupported: staged copies
      }
    int rc = kmd_exchange_export(c, x.gather + i * slot_words, cap);
    if (rc) return rc;
    KMD_CUDA(c, cudaEventRecord(x.ev_export, c->stream));
  }
  // ... and pulls every other context's slot over the peer link once that slot is written
  for (uint32_t j = 0; j < n; j++)
  {
    kmd_context* c = ctxs[j];
    KMD_CUDA(c, cudaSetDevice(c->device));
    for (uint32_t i = 0; i < n; i++)
    {
      if (i == j) continue;
      KMD_CUDA(c, cudaStreamWaitEvent(c->stream, ctxs[i]->xch.ev_export, 0));
      KMD_CUDA(c, cudaMemcpyPeerAsync(c->xch.gather + i * slot_words, c->device, ctxs[i]->xch.gather + i * slot_words,
                                      ctxs[i]->device, slot_words * sizeof(uint64_t), c->stream));
    }
    int rc = kmd_exchange_correct(c, c->xch.gather, n, cap, correction, significance);
    if (rc) return rc;
  }
  for (uint32_t j = 0; j < n; j++)
  {
    int rc = kmd_exchange_result(ctxs[j], &decisions[j]);
    if (rc) return rc;
    if (decisions[j].overflow) return fail(ctxs[j], KMD_ERR_STATE, "exchange slot overflow with an exact capacity");
  }
  return KMD_OK;
}

int kmd_exchange_result(kmd_context* ctx, kmd_exchange_decision* out)
{
  if (!ctx || !out) return KMD_ERR_ARG;
  ExchangeState& x = ctx->xch;
  if (!x.pending) return fail(ctx, KMD_ERR_STATE, "kmd_exchange_result without kmd_exchange_correct");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  x.pending = false;
  const ExchangeDecision& d = *x.h_dec;
  out->total_rows = d.total_rows;
  out->total_survivors = d.total_survivors;
  out->max_rank_survivors = d.max_rank_survivors;
  out->n_accepted = d.n_accepted;
  out->p_cut = d.p_cut;
  out->kept_control = ctx->h_cnt->keep_control;
  out->kept_case = ctx->h_cnt->keep_case;
  out->overflow = d.overflow;
  out->reserved = 0;
  if (ctx->h_cnt->overflow)
    return fail(ctx, KMD_ERR_OVERFLOW, "%s buffer overflow: raise kmd_config.max_survivors (now %llu)",
                (ctx->h_cnt->overflow & 1) ? "candidate" : "survivor", static_cast<unsigned long long>(ctx->cap));
  if (d.overflow) return KMD_OK;   // decision void: the caller repeats the exchange with cap >= max_rank_survivors
  ctx->n_final = ctx->h_cnt->keep_total;
  ctx->final_valid = true;
  ctx->final_sorted = false;
  ctx->final_by_p = x.by_p;
  return KMD_OK;
}

}  // extern "C"
