// C-ABI entry points of the per-sample count-file input path (included at the end of kmd_api.cu).
// Kernels: kmd_merge.cuh.  Reference: km::KmerMerger (kmtricks/merge.hpp:102-361) driven by
// global_merge::merge() (include/kmdiff/merge.hpp:237-309).

namespace {

template <class T>
int grow(kmd_context* ctx, T*& p, size_t& cap, size_t want)
{
  if (want <= cap) return KMD_OK;
  if (p) { KMD_CUDA(ctx, cudaFree(p)); p = nullptr; cap = 0; }
  want += want / 8;   // head room: the chunks of one partition have similar sizes
  void* q = nullptr;
  KMD_CUDA(ctx, cudaMalloc(&q, want * sizeof(T)));
  p = static_cast<T*>(q);
  cap = want;
  return KMD_OK;
}

void merge_free(kmd_context* ctx)
{
  MergeState& m = ctx->mrg;
  cudaFree(m.raw);
  for (int b = 0; b < 2; b++) { cudaFree(m.klo[b]); cudaFree(m.khi[b]); cudaFree(m.pay[b]); }
  cudaFree(m.flag); cudaFree(m.tile); cudaFree(m.head_pos); cudaFree(m.rows); cudaFree(m.stream_off);
  cudaFree(m.ab_min); cudaFree(m.plan); cudaFree(m.total); cudaFree(m.split);
  cudaFree(m.lz_in); cudaFree(m.lz_tab); cudaFree(m.lz_status);
  cudaFree(m.udesc); cudaFree(m.ustatus); cudaFree(m.usplit); cudaFree(m.uflags); cudaFree(m.uorder);
  for (auto& f : m.pf) { cudaFree(f.buf); if (f.done) cudaEventDestroy(f.done); }
  for (cudaEvent_t e : m.ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : m.lz_ev) if (e) cudaEventDestroy(e);
  m = MergeState{};
}

struct MergePass { size_t run_off, pair_tile; uint32_t n_runs, n_pairs, tiles; };

template <int W>
int merge_passes(kmd_context* ctx, const std::vector<MergePass>& passes, int* final_buf)
{
  MergeState& m = ctx->mrg;
  uint32_t max_tiles = 0;
  for (const MergePass& ps : passes) max_tiles = std::max(max_tiles, ps.tiles);
  int rc = grow(ctx, m.split, m.split_cap, static_cast<size_t>(max_tiles) + 1);
  if (rc) return rc;
  static const bool stage_out = env_u32("KMD_MERGE_STAGE_OUT", 1) != 0;
  const size_t smem = stage_out ? mrg::merge_smem_bytes<W, true>() : mrg::merge_smem_bytes<W, false>();
  if (stage_out) KMD_CUDA(ctx, cudaFuncSetAttribute(mrg::merge_pass_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  else KMD_CUDA(ctx, cudaFuncSetAttribute(mrg::merge_pass_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int cur = 0;
  for (const MergePass& ps : passes)
  {
    mrg::MergeArgs<W> A{};
    A.in = {m.klo[cur], m.khi[cur]};
    A.out = {m.klo[cur ^ 1], m.khi[cur ^ 1]};
    A.pay_in = m.pay[cur];
    A.pay_out = m.pay[cur ^ 1];
    A.run_off = m.plan + ps.run_off;
    A.pair_tile = m.plan + ps.pair_tile;
    A.n_runs = ps.n_runs;
    A.n_pairs = ps.n_pairs;
    if (ps.tiles)
    {
      mrg::merge_partition_kernel<W><<<(ps.tiles + 127) / 128, 128, 0, ctx->stream>>>(A, ps.tiles, m.split);
      if (stage_out) mrg::merge_pass_kernel<W, true><<<ps.tiles, mrg::kThreads, smem, ctx->stream>>>(A, m.split);
      else mrg::merge_pass_kernel<W, false><<<ps.tiles, mrg::kThreads, smem, ctx->stream>>>(A, m.split);
      ctx->launches += 2;
    }
    cur ^= 1;
  }
  KMD_CUDA(ctx, cudaGetLastError());
  *final_buf = cur;
  return KMD_OK;
}

template <int W>
int merge_rows(kmd_context* ctx, int buf, uint32_t n, uint32_t* n_rows_out)
{
  MergeState& m = ctx->mrg;
  const uint32_t n_samples = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  const uint32_t tiles = (n + mrg::kScanTile - 1) / mrg::kScanTile;
  int rc;
  if ((rc = grow(ctx, m.flag, m.flag_cap, n))) return rc;
  if ((rc = grow(ctx, m.tile, m.tile_cap, tiles))) return rc;
  mrg::KeyArrays<W> keys{m.klo[buf], m.khi[buf]};
  mrg::heads_kernel<W><<<tiles, 256, 0, ctx->stream>>>(keys, m.pay[buf], n, m.flag, m.tile);
  scan_exclusive_kernel<<<1, 1024, 0, ctx->stream>>>(m.tile, tiles, m.total);
  ctx->launches += 2;
  uint32_t n_rows = 0;
  KMD_CUDA(ctx, cudaMemcpyAsync(&n_rows, m.total, sizeof n_rows, cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // also: every H2D copy of the caller's streams is done
  *n_rows_out = n_rows;
  if (n_rows == 0) return KMD_OK;
  if ((rc = grow(ctx, m.head_pos, m.head_cap, n_rows))) return rc;
  if ((rc = grow(ctx, m.rows, m.rows_cap, static_cast<size_t>(n_rows) * ctx->plan.row_bytes + 16))) return rc;
  mrg::head_positions_kernel<<<tiles, 256, 0, ctx->stream>>>(m.flag, n, m.tile, m.head_pos);
  const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>((n_rows + 7) / 8, ctx->sm_count * 8));
  switch (ctx->cfg.count_bytes)
  {
    case 1: mrg::rows_kernel<W, 1><<<blocks, 256, 0, ctx->stream>>>(keys, m.pay[buf], n, m.head_pos, n_rows, n_samples, m.rows); break;
    case 2: mrg::rows_kernel<W, 2><<<blocks, 256, 0, ctx->stream>>>(keys, m.pay[buf], n, m.head_pos, n_rows, n_samples, m.rows); break;
    default: mrg::rows_kernel<W, 4><<<blocks, 256, 0, ctx->stream>>>(keys, m.pay[buf], n, m.head_pos, n_rows, n_samples, m.rows); break;
  }
  ctx->launches += 2;
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

struct UnionLevel { size_t run_base, run_len, pair_tile, out_len; uint32_t n_runs, n_pairs, tiles; };

// The rows of a chunk by the union tree of kmd_union.cuh: keys -> levels -> (one synchronisation: row count) -> scatter.
// *used = false (and no rows) when a record lies below its sample's abundance_min: the caller then runs the pairwise
// tree, which carries the solid flags along.
template <int W>
int merge_union(kmd_context* ctx, const uint8_t* raw, const std::vector<uint64_t>& off, const uint32_t* ab_min, uint32_t n,
                uint32_t* n_rows_out, bool* used)
{
  MergeState& m = ctx->mrg;
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  const uint32_t B = ctx->cfg.count_bytes;
  *used = false;

  // the tree: a run's start is the sum of the leaf lengths before it (host), its length is the level's output (device)
  const uint32_t tile = uni::kThreads * uni::kItems;
  std::vector<uint32_t> plan;
  std::vector<UnionLevel> levels;
  {
    std::vector<uint32_t> base(N), ub(N);
    for (uint32_t s = 0; s < N; s++) { base[s] = static_cast<uint32_t>(off[s]); ub[s] = static_cast<uint32_t>(off[s + 1] - off[s]); }
    bool leaves = true;
    while (base.size() > 1)
    {
      const uint32_t n_runs = static_cast<uint32_t>(base.size()), n_pairs = (n_runs + 1) / 2;
      UnionLevel lv{};
      lv.n_runs = n_runs; lv.n_pairs = n_pairs;
      lv.run_base = plan.size();
      plan.insert(plan.end(), base.begin(), base.end());
      lv.run_len = plan.size();
      if (leaves) plan.insert(plan.end(), ub.begin(), ub.end()); else plan.insert(plan.end(), n_runs, 0u);
      leaves = false;
      lv.pair_tile = plan.size();
      std::vector<uint32_t> nbase(n_pairs), nub(n_pairs);
      uint32_t tiles = 0;
      for (uint32_t q = 0; q < n_pairs; q++)
      {
        nbase[q] = base[2 * q];
        nub[q] = ub[2 * q] + (2 * q + 1 < n_runs ? ub[2 * q + 1] : 0u);
        plan.push_back(tiles);
        tiles += (nub[q] + tile - 1) / tile;
      }
      plan.push_back(tiles);
      lv.tiles = tiles;
      if (!levels.empty()) levels.back().out_len = lv.run_len;
      levels.push_back(lv);
      base.swap(nbase);
      ub.swap(nub);
    }
    if (!levels.empty()) { levels.back().out_len = plan.size(); plan.push_back(0u); }
  }
  uint32_t max_tiles = 0;
  for (const UnionLevel& lv : levels) max_tiles = std::max(max_tiles, lv.tiles);
  int rc;
  if ((rc = grow(ctx, m.plan, m.plan_cap, std::max<size_t>(1, plan.size())))) return rc;
  if ((rc = grow(ctx, m.udesc, m.udesc_cap, std::max<size_t>(1, max_tiles)))) return rc;
  if ((rc = grow(ctx, m.uorder, m.uorder_cap, std::max<size_t>(1, max_tiles)))) return rc;
  if ((rc = grow(ctx, m.ustatus, m.ustatus_cap, static_cast<size_t>(max_tiles) + 1))) return rc;
  if (!m.uflags) KMD_CUDA(ctx, cudaMalloc(&m.uflags, 4 * sizeof(uint32_t)));
  KMD_CUDA(ctx, cudaMemsetAsync(m.uflags, 0, 4 * sizeof(uint32_t), ctx->stream));
  if (!plan.empty())
    KMD_CUDA(ctx, cudaMemcpyAsync(m.plan, plan.data(), plan.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));

  // leaves = klo[0] / khi[0] (or the records themselves: 4-byte counts, see RawKeys); the levels alternate between
  // klo[1] / khi[1] and pay[0] / pay[1]
  mrg::KeyArrays<W> leaf{m.klo[0], m.khi[0]};
  mrg::KeyArrays<W> ping{m.klo[1], m.khi[1]}, pong{m.pay[0], m.pay[1]};
  const uint32_t ab_min_max = *std::max_element(ab_min, ab_min + N);
  const bool raw_leaf = B == 4 && !levels.empty() && m.knob_raw_leaf;
  const uni::RawKeys<W> raw_keys{reinterpret_cast<const uint32_t*>(raw)};
  KMD_CUDA(ctx, cudaEventRecord(m.ev[0], ctx->stream));
  if (!raw_leaf)
  {
    uni::KeysArgs K{};
    K.raw = raw; K.stream_off = m.stream_off; K.ab_min = m.ab_min; K.n_streams = N; K.n = n;
    K.kmer_words = W; K.count_bytes = B;
    K.ab_min_max = ab_min_max;
    K.klo = leaf.lo; K.khi = leaf.hi; K.nonsolid = m.uflags;
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>((n + 255) / 256, ctx->sm_count * 8));
    uni::keys_kernel<<<blocks, 256, 0, ctx->stream>>>(K);
    ctx->launches++;
  }
  KMD_CUDA(ctx, cudaEventRecord(m.ev[1], ctx->stream));
  mrg::KeyArrays<W> cur = leaf;
  int flip = 0;
  bool first = true;
  for (const UnionLevel& lv : levels)
  {
    const bool from_raw = first && raw_leaf;
    first = false;
    mrg::KeyArrays<W> out = flip ? pong : ping;
    if (lv.tiles)   // else: nothing but empty runs
    {
      KMD_CUDA(ctx, cudaMemsetAsync(m.ustatus, 0, (static_cast<size_t>(lv.tiles) + 1) * sizeof(unsigned long long), ctx->stream));
      // tickets go round robin over the pairs (not worth computing for one pair, or for thousands of short ones)
      const bool round_robin = lv.n_pairs > 1 && lv.n_pairs <= 1024;
      if (round_robin)
      {
        uni::tile_order_kernel<<<(lv.tiles + 255) / 256, 256, 0, ctx->stream>>>(m.plan + lv.pair_tile, lv.n_pairs, lv.tiles, m.uorder);
        ctx->launches++;
      }
      auto fill = [&](auto& A) {
        A.out = out;
        A.ab_min = m.ab_min; A.ab_min_max = ab_min_max; A.nonsolid = m.uflags;
        A.run_base = m.plan + lv.run_base;
        A.run_len = m.plan + lv.run_len;
        A.pair_tile = m.plan + lv.pair_tile;
        A.out_len = m.plan + lv.out_len;
        A.n_runs = lv.n_runs; A.n_pairs = lv.n_pairs; A.tiles = lv.tiles; A.tile = tile;
        A.desc = m.udesc;
        A.status = m.ustatus;
        A.ticket = reinterpret_cast<uint32_t*>(m.ustatus + lv.tiles);
        A.order = round_robin ? m.uorder : nullptr;
      };
      if (from_raw)
      {
        uni::LevelArgs<W, uni::RawKeys<W>> A{};
        A.in = raw_keys;
        fill(A);
        uni::level_partition_kernel<W><<<(lv.tiles + 3) / 4, 128, 0, ctx->stream>>>(A);   // one warp per tile
        uni::union_level_kernel<W, uni::kThreads><<<lv.tiles, uni::kThreads, 0, ctx->stream>>>(A);
      }
      else
      {
        uni::LevelArgs<W, uni::SoaKeys<W>> A{};
        A.in.lo = cur.lo; A.in.hi = cur.hi;
        fill(A);
        uni::level_partition_kernel<W><<<(lv.tiles + 3) / 4, 128, 0, ctx->stream>>>(A);
        uni::union_level_kernel<W, uni::kThreads><<<lv.tiles, uni::kThreads, 0, ctx->stream>>>(A);
      }
      ctx->launches += 2;
    }
    cur = out;
    flip ^= 1;
  }
  KMD_CUDA(ctx, cudaGetLastError());
  KMD_CUDA(ctx, cudaEventRecord(m.ev[2], ctx->stream));

  uint32_t flags[4] = {0, 0, 0, 0};
  KMD_CUDA(ctx, cudaMemcpyAsync(flags, m.uflags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
  if (!plan.empty())
    KMD_CUDA(ctx, cudaMemcpyAsync(plan.data(), m.plan, plan.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // also: every H2D copy of the caller's streams is done
  if (flags[0]) return KMD_OK;
  *used = true;
  uint32_t n_rows = n;   // one stream: its keys are the rows
  m.last_keys_in = m.last_keys_out = 0;
  if (!levels.empty())
  {
    n_rows = plan[levels.back().out_len];
    for (const UnionLevel& lv : levels)
    {
      for (uint32_t r = 0; r < lv.n_runs; r++) m.last_keys_in += plan[lv.run_len + r];
      for (uint32_t q = 0; q < lv.n_pairs; q++) m.last_keys_out += plan[lv.out_len + q];
    }
  }
  m.last_passes = static_cast<uint32_t>(levels.size());
  *n_rows_out = n_rows;
  if (n_rows == 0) return KMD_OK;

  // Rows assembled in shared memory (whole rows written once) when the cohort is dense enough to fill them, scattered
  // into zeroed rows otherwise (a sparse row is mostly the zeros).  KMD_UNION_ROWS: 0 choose, 1 shared memory, 2 scatter.
  const size_t row_bytes = ctx->plan.row_bytes;
  const uint32_t rows_mode = m.knob_rows, smem_budget = m.knob_smem, scatter_tile = m.knob_scatter_tile;
  // Rows per assembly tile.  Narrow rows: ~28 KB of them, so that 6-7 CTAs share an SM (c2m: 1.23 ms per step against
  // 1.42 with 56 KB tiles); wide rows: no fewer than 80 (every tile costs one search per stream: c3m 2.20 ms with 88-row
  // tiles against 2.39 with 64) as long as the tile stays under 100 KB (profiles/r03d).  KMD_UNION_SMEM_KB fixes the budget.
  uint32_t asm_tile = 0;
  if (N <= static_cast<uint32_t>(uni::kMaxFlat))
  {
    const size_t per_row = row_bytes + 8 * W, fixed = (2ull * N + 2) * 4 + 16;
    size_t rows_fit;
    if (smem_budget) rows_fit = smem_budget > fixed ? (smem_budget - fixed) / per_row : 0;
    else
    {
      rows_fit = (28u << 10) / per_row;
      if (rows_fit < 80) rows_fit = std::min<size_t>(80, ((100u << 10) - fixed) / per_row);
    }
    asm_tile = static_cast<uint32_t>(std::min<size_t>(uni::kRowTile, rows_fit) / 16 * 16);
  }
  const bool dense = static_cast<uint64_t>(n) * 4 >= static_cast<uint64_t>(n_rows) * N;   // a quarter of the cells or more
  const bool assemble = asm_tile >= 16 && (rows_mode == 1 || (rows_mode == 0 && dense));
  const uint32_t row_tile = assemble ? asm_tile : scatter_tile;
  const uint32_t row_tiles = (n_rows + row_tile - 1) / row_tile;
  if ((rc = grow(ctx, m.rows, m.rows_cap, static_cast<size_t>(n_rows) * row_bytes + 16))) return rc;
  if ((rc = grow(ctx, m.usplit, m.usplit_cap, (static_cast<size_t>(row_tiles) + 1) * N))) return rc;
  uni::ScatterArgs<W> S{};
  S.leaf = leaf;
  S.rows_key = cur;
  S.raw = raw; S.stream_off = m.stream_off; S.ab_min = m.ab_min; S.split = m.usplit;
  S.n_streams = N; S.n_rows = n_rows; S.tiles = row_tiles; S.row_tile = row_tile;
  S.rec = 8 * W + B; S.row_bytes = static_cast<uint32_t>(row_bytes);
  S.rows = m.rows;
  S.err = &ctx->d_cnt->overflow;
  const uint64_t splits = (static_cast<uint64_t>(row_tiles) + 1) * N;
  if (raw_leaf) uni::scatter_partition_kernel<W><<<static_cast<uint32_t>((splits + 255) / 256), 256, 0, ctx->stream>>>(S, raw_keys);
  else uni::scatter_partition_kernel<W><<<static_cast<uint32_t>((splits + 255) / 256), 256, 0, ctx->stream>>>(S, leaf);
  if (assemble)
  {
    const size_t smem = uni::assemble_smem_bytes<W>(row_tile, S.row_bytes, N);
    bool& attr_set = m.asm_attr_set[W - 1];   // the attribute belongs to (function, device): once per context is enough
    if (!attr_set)
    {
      const int most = 200 << 10;
      KMD_CUDA(ctx, cudaFuncSetAttribute(uni::assemble_rows_kernel<W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
      KMD_CUDA(ctx, cudaFuncSetAttribute(uni::assemble_rows_kernel<W, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
      KMD_CUDA(ctx, cudaFuncSetAttribute(uni::assemble_rows_kernel<W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
      attr_set = true;
    }
    switch (B)
    {
      case 1: uni::assemble_rows_kernel<W, 1><<<row_tiles, 256, smem, ctx->stream>>>(S); break;
      case 2: uni::assemble_rows_kernel<W, 2><<<row_tiles, 256, smem, ctx->stream>>>(S); break;
      default: uni::assemble_rows_kernel<W, 4><<<row_tiles, 256, smem, ctx->stream>>>(S); break;
    }
  }
  else
  {
    KMD_CUDA(ctx, cudaMemsetAsync(m.rows, 0, static_cast<size_t>(n_rows) * row_bytes, ctx->stream));
    switch (B)
    {
      case 1: uni::scatter_rows_kernel<W, 1><<<row_tiles, 256, 0, ctx->stream>>>(S); break;
      case 2: uni::scatter_rows_kernel<W, 2><<<row_tiles, 256, 0, ctx->stream>>>(S); break;
      default: uni::scatter_rows_kernel<W, 4><<<row_tiles, 256, 0, ctx->stream>>>(S); break;
    }
  }
  m.last_rows_mode = assemble ? 1 : 2;
  ctx->launches += 2;
  KMD_CUDA(ctx, cudaGetLastError());
  return KMD_OK;
}

// off[s] .. off[s + 1]: records of sample s inside the chunk (ascending and unique per stream); fill_raw() enqueues
// whatever brings those records to m.raw + off[s] * rec (H2D copies of host records, or the device LZ4 decoder).
// Leaves the dense rows in ctx->mrg.rows.
// ext_raw (device, may be null): the records are already there, back to back, and fill_raw has nothing to bring.
template <class FillRaw>
int merge_from_offsets(kmd_context* ctx, const std::vector<uint64_t>& off, const uint32_t* ab_min, uint32_t* n_rows, FillRaw&& fill_raw,
                       const uint8_t* ext_raw = nullptr)
{
  MergeState& m = ctx->mrg;
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  const uint32_t W = ctx->cfg.kmer_words, B = ctx->cfg.count_bytes;
  const size_t rec = 8ull * W + B;
  const uint64_t n64 = off[N];
  *n_rows = 0;
  if (n64 == 0) return KMD_OK;
  if (n64 >= (1ull << 31)) return fail(ctx, KMD_ERR_ARG, "more than 2^31 records in one merge call: split the partition into k-mer ranges");
  const uint32_t n = static_cast<uint32_t>(n64);

  int rc;
  if (!ext_raw && (rc = grow(ctx, m.raw, m.raw_cap, n * rec + 16))) return rc;
  const uint8_t* raw = ext_raw ? ext_raw : m.raw;
  for (int b = 0; b < 2; b++)
  {
    if ((rc = grow(ctx, m.klo[b], m.klo_cap[b], n))) return rc;
    if (W == 2 && (rc = grow(ctx, m.khi[b], m.khi_cap[b], n))) return rc;
    if ((rc = grow(ctx, m.pay[b], m.pay_cap[b], n))) return rc;
  }
  if ((rc = grow(ctx, m.stream_off, m.off_cap, N + 1))) return rc;
  if ((rc = grow(ctx, m.ab_min, m.ab_cap, N))) return rc;
  if (!m.total) KMD_CUDA(ctx, cudaMalloc(&m.total, sizeof(uint32_t)));
  for (cudaEvent_t& e : m.ev) if (!e) KMD_CUDA(ctx, cudaEventCreate(&e));
  m.timed = false;

  KMD_CUDA(ctx, cudaMemcpyAsync(m.stream_off, off.data(), (N + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(m.ab_min, ab_min, N * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = fill_raw())) return rc;

  // the union tree (kmd_union.cuh) unless a record is below its sample's abundance_min or the caller asks for the pairwise tree
  if (m.algo < 0)
  {
    // the knobs are read once per context
    m.algo = env_u32("KMD_MERGE_PAIRWISE", 0) ? 1 : 0;
    m.knob_rows = env_u32("KMD_UNION_ROWS", 0);
    m.knob_raw_leaf = env_u32("KMD_UNION_RAW_LEAF", 1) != 0;
    m.knob_smem = std::min<uint32_t>(200, env_u32("KMD_UNION_SMEM_KB", 0)) << 10;   // 0: chosen from the row size
    m.knob_scatter_tile = std::max<uint32_t>(16, std::min<uint32_t>(uni::kRowTile, env_u32("KMD_UNION_SCATTER_TILE", 1024) / 16 * 16));
  }
  m.last_algo = 1;
  if (m.algo == 0)
  {
    bool used = false;
    rc = W == 2 ? merge_union<2>(ctx, raw, off, ab_min, n, n_rows, &used) : merge_union<1>(ctx, raw, off, ab_min, n, n_rows, &used);
    if (rc) return rc;
    if (used)
    {
      KMD_CUDA(ctx, cudaEventRecord(m.ev[3], ctx->stream));
      m.timed = true;
      m.last_algo = 0;
      m.last_records = n;
      return KMD_OK;
    }
  }

  // the merge tree: leaves = streams, every pass halves the number of runs
  std::vector<uint32_t> plan;
  std::vector<MergePass> passes;
  {
    std::vector<uint32_t> runs(off.begin(), off.end());   // n_runs + 1 offsets
    while (runs.size() - 1 > 1)
    {
      const uint32_t n_runs = static_cast<uint32_t>(runs.size() - 1), n_pairs = (n_runs + 1) / 2;
      MergePass ps{};
      ps.n_runs = n_runs; ps.n_pairs = n_pairs;
      ps.run_off = plan.size();
      plan.insert(plan.end(), runs.begin(), runs.end());
      ps.pair_tile = plan.size();
      std::vector<uint32_t> next(1, 0);
      uint32_t tiles = 0;
      for (uint32_t q = 0; q < n_pairs; q++)
      {
        const uint32_t end = runs[std::min<uint32_t>(2 * q + 2, n_runs)];
        plan.push_back(tiles);
        tiles += (end - runs[2 * q] + mrg::kTile - 1) / mrg::kTile;
        next.push_back(end);
      }
      plan.push_back(tiles);
      ps.tiles = tiles;
      passes.push_back(ps);
      runs.swap(next);
    }
  }
  if ((rc = grow(ctx, m.plan, m.plan_cap, std::max<size_t>(1, plan.size())))) return rc;

  if (!plan.empty())
    KMD_CUDA(ctx, cudaMemcpyAsync(m.plan, plan.data(), plan.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));

  KMD_CUDA(ctx, cudaEventRecord(m.ev[0], ctx->stream));
  mrg::UnpackArgs U{};
  U.raw = raw; U.stream_off = m.stream_off; U.ab_min = m.ab_min; U.n_streams = N; U.n = n;
  U.kmer_words = W; U.count_bytes = B;
  U.klo = m.klo[0]; U.khi = m.khi[0]; U.pay = m.pay[0];
  const uint32_t ublocks = std::max<uint32_t>(1, std::min<uint32_t>((n + 255) / 256, ctx->sm_count * 8));
  mrg::unpack_kernel<<<ublocks, 256, 0, ctx->stream>>>(U);
  ctx->launches++;
  KMD_CUDA(ctx, cudaGetLastError());

  int buf = 0;
  KMD_CUDA(ctx, cudaEventRecord(m.ev[1], ctx->stream));
  rc = W == 2 ? merge_passes<2>(ctx, passes, &buf) : merge_passes<1>(ctx, passes, &buf);
  if (rc) return rc;
  KMD_CUDA(ctx, cudaEventRecord(m.ev[2], ctx->stream));
  // plan/off/ab_min are host temporaries: merge_rows synchronises the stream before we return
  rc = W == 2 ? merge_rows<2>(ctx, buf, n, n_rows) : merge_rows<1>(ctx, buf, n, n_rows);
  if (rc) return rc;
  KMD_CUDA(ctx, cudaEventRecord(m.ev[3], ctx->stream));
  m.timed = true;
  m.last_passes = static_cast<uint32_t>(passes.size());
  m.last_records = n;
  return KMD_OK;
}

// If (ptrs, bytes) is exactly what kmd_merge_prefetch put on the copy stream, make `dst` that buffer (the old one becomes
// the next prefetch target) and order the context's stream behind the copies.  Returns true when taken.
bool take_prefetch(kmd_context* ctx, int kind, const void* const* ptrs, const uint64_t* bytes, uint32_t n, uint8_t*& dst, size_t& dst_cap)
{
  for (auto& f : ctx->mrg.pf)
  {
    if (f.kind != kind || f.ptrs.size() != n) continue;
    bool same = true;
    for (uint32_t s = 0; s < n && same; s++) same = f.ptrs[s] == ptrs[s] && f.bytes[s] == bytes[s];
    if (!same) continue;
    std::swap(dst, f.buf);
    std::swap(dst_cap, f.cap);
    f.kind = -1;
    cudaStreamWaitEvent(ctx->stream, f.done, 0);
    return true;
  }
  return false;
}

// Host blobs -> device, blob s to dst + at[s].  One copy per blob: neighbours that happen to be adjacent in the address
// space may still belong to two pinned allocations, and a copy must not span two.
int copy_blobs(kmd_context* ctx, uint8_t* dst, const std::vector<uint64_t>& at, const void* const* blobs, const uint64_t* bytes,
               uint32_t n, cudaStream_t stream)
{
  for (uint32_t s = 0; s < n; s++)
    if (bytes[s])
      KMD_CUDA(ctx, cudaMemcpyAsync(dst + at[s], blobs[s], bytes[s], cudaMemcpyHostToDevice, stream));
  return KMD_OK;
}

// streams[s] = n_records[s] HOST records of sample s
int merge_impl(kmd_context* ctx, const void* const* streams, const uint64_t* n_records, const uint32_t* ab_min, uint32_t* n_rows)
{
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  if (N > 0xFFFFu) return fail(ctx, KMD_ERR_ARG, "more than 65535 samples");
  const size_t rec = 8ull * ctx->cfg.kmer_words + ctx->cfg.count_bytes;
  std::vector<uint64_t> off(N + 1, 0);
  for (uint32_t s = 0; s < N; s++)
  {
    if (n_records[s] && !streams[s]) return fail(ctx, KMD_ERR_ARG, "null stream pointer (sample %u)", s);
    off[s + 1] = off[s] + n_records[s];
  }
  std::vector<uint64_t> bytes(N);
  for (uint32_t s = 0; s < N; s++) bytes[s] = n_records[s] * rec;
  // a prefetched chunk is already (on its way) in the other record buffer: swap it in before the buffers are sized
  const bool prefetched = take_prefetch(ctx, 0, streams, bytes.data(), N, ctx->mrg.raw, ctx->mrg.raw_cap);
  return merge_from_offsets(ctx, off, ab_min, n_rows, [&]() -> int {
    if (prefetched) return KMD_OK;
    std::vector<uint64_t> at(N);
    for (uint32_t s = 0; s < N; s++) at[s] = off[s] * rec;
    return copy_blobs(ctx, ctx->mrg.raw, at, streams, bytes.data(), N, ctx->stream);
  });
}

// The samples' files of one partition as they lie on disk: frames[s] = the payload after the 41-byte header (HOST),
// an LZ4 frame when compressed[s] != 0, raw records otherwise.  Compressed payloads are uploaded as they are and
// decoded on the device, one warp per stream (kmd_lz4.cuh); the decoded records land where the merge tree reads them.
// records_out (host, may be null): when given, the decoded records of every stream are copied there back to back
// and no merge is run (tests / tools).
int merge_lz4_impl(kmd_context* ctx, const void* const* frames, const uint64_t* frame_bytes, const uint8_t* compressed,
                   const uint32_t* ab_min, uint64_t max_records, uint32_t* n_rows, uint64_t* n_records_out, void* records_out,
                   uint64_t records_cap)
{
  MergeState& m = ctx->mrg;
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  if (N > 0xFFFFu) return fail(ctx, KMD_ERR_ARG, "more than 65535 samples");
  const size_t rec = 8ull * ctx->cfg.kmer_words + ctx->cfg.count_bytes;
  *n_rows = 0;
  // ---- upload the compressed frames, 16-byte aligned ----
  std::vector<uint64_t> tab(5ull * N, 0);
  uint64_t* in_off = tab.data();
  uint64_t* in_size = in_off + N;
  uint64_t* out_off = in_size + N;
  uint64_t* out_cap = out_off + N;
  uint64_t* out_size = out_cap + N;
  uint64_t in_total = 0;
  uint32_t n_lz = 0;
  for (uint32_t s = 0; s < N; s++)
  {
    if (frame_bytes[s] && !frames[s]) return fail(ctx, KMD_ERR_ARG, "null stream pointer (sample %u)", s);
    const bool lz = !compressed || compressed[s];
    in_off[s] = in_total;
    in_size[s] = lz ? frame_bytes[s] : 0;
    if (lz) { in_total += (frame_bytes[s] + 15) / 16 * 16; n_lz++; }
    else
    {
      if (frame_bytes[s] % rec) return fail(ctx, KMD_ERR_ARG, "sample %u: payload is not a whole number of records", s);
      out_size[s] = frame_bytes[s];
    }
  }
  int rc;
  bool prefetched = false;
  if (n_lz == N) prefetched = take_prefetch(ctx, 1, frames, frame_bytes, N, m.lz_in, m.lz_in_cap);
  if (!prefetched && (rc = grow(ctx, m.lz_in, m.lz_in_cap, in_total + 16))) return rc;
  if ((rc = grow(ctx, m.lz_tab, m.lz_tab_cap, 5ull * N))) return rc;
  if ((rc = grow(ctx, m.lz_status, m.lz_status_cap, N))) return rc;
  for (cudaEvent_t& e : m.lz_ev) if (!e) KMD_CUDA(ctx, cudaEventCreate(&e));
  for (uint32_t s = 0; s < N && !prefetched; s++)
    if (in_size[s])
      KMD_CUDA(ctx, cudaMemcpyAsync(m.lz_in + in_off[s], frames[s], in_size[s], cudaMemcpyHostToDevice, ctx->stream));
  KMD_CUDA(ctx, cudaMemcpyAsync(m.lz_tab, tab.data(), 5ull * N * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  lz4::Args A{};
  A.in = m.lz_in;
  A.in_off = m.lz_tab;
  A.in_size = m.lz_tab + N;
  A.out_off = m.lz_tab + 2ull * N;
  A.out_cap = m.lz_tab + 3ull * N;
  A.out_size = m.lz_tab + 4ull * N;
  A.status = m.lz_status;
  A.n = N;
  std::vector<uint32_t> status(N, 0);
  const uint32_t grid = (N + lz4::kWarps - 1) / lz4::kWarps;
  cudaEvent_t lz0 = m.lz_ev[0], lz1 = m.lz_ev[1];
  KMD_CUDA(ctx, cudaEventRecord(lz0, ctx->stream));
  if (n_lz)
  {
    // ---- pass 1: decoded sizes (the frames carry no content size) ----
    A.out = nullptr;
    lz4::frames_kernel<<<grid, lz4::kWarps * 32, 0, ctx->stream>>>(A);
    ctx->launches++;
    KMD_CUDA(ctx, cudaGetLastError());
    std::vector<uint64_t> sizes(N);
    KMD_CUDA(ctx, cudaMemcpyAsync(sizes.data(), A.out_size, N * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaMemcpyAsync(status.data(), m.lz_status, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < N; s++)
    {
      if (!in_size[s]) continue;
      if (status[s]) return fail(ctx, KMD_ERR_ARG, "sample %u: malformed LZ4 frame (decoder status %u)", s, status[s]);
      if (sizes[s] % rec) return fail(ctx, KMD_ERR_ARG, "sample %u: decoded payload is not a whole number of records", s);
      out_size[s] = sizes[s];
    }
  }
  std::vector<uint64_t> off(N + 1, 0);
  for (uint32_t s = 0; s < N; s++) off[s + 1] = off[s] + out_size[s] / rec;
  if (n_records_out) *n_records_out = off[N];
  if (max_records && off[N] > max_records)
    return fail(ctx, KMD_ERR_OVERFLOW, "partition holds %llu records, more than the %llu one merge call may take: stream it in k-mer ranges",
                static_cast<unsigned long long>(off[N]), static_cast<unsigned long long>(max_records));
  for (uint32_t s = 0; s < N; s++) { out_off[s] = off[s] * rec; out_cap[s] = out_size[s]; }

  auto fill = [&]() -> int {
    // ---- pass 2: decode into the merge tree's record buffer; raw streams are plain copies ----
    KMD_CUDA(ctx, cudaMemcpyAsync(m.lz_tab, tab.data(), 5ull * N * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (n_lz)
    {
      lz4::Args D = A;
      D.out = m.raw;
      lz4::frames_kernel<<<grid, lz4::kWarps * 32, 0, ctx->stream>>>(D);
      ctx->launches++;
      KMD_CUDA(ctx, cudaGetLastError());
    }
    for (uint32_t s = 0; s < N; s++)
      if (!in_size[s] && out_size[s])
        KMD_CUDA(ctx, cudaMemcpyAsync(m.raw + out_off[s], frames[s], out_size[s], cudaMemcpyHostToDevice, ctx->stream));
    KMD_CUDA(ctx, cudaEventRecord(lz1, ctx->stream));
    return KMD_OK;
  };
  if (records_out)
  {
    if (off[N] > records_cap) return fail(ctx, KMD_ERR_ARG, "output holds %llu records, the streams decode to %llu",
                                          static_cast<unsigned long long>(records_cap), static_cast<unsigned long long>(off[N]));
    if ((rc = grow(ctx, m.raw, m.raw_cap, off[N] * rec + 16))) return rc;
    if ((rc = fill())) return rc;
    KMD_CUDA(ctx, cudaMemcpyAsync(records_out, m.raw, off[N] * rec, cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaMemcpyAsync(status.data(), m.lz_status, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < N; s++)
      if (in_size[s] && status[s]) return fail(ctx, KMD_ERR_ARG, "sample %u: LZ4 decode failed (status %u)", s, status[s]);
    KMD_CUDA(ctx, cudaEventElapsedTime(&m.last_lz4_ms, lz0, lz1));
    return KMD_OK;
  }
  rc = merge_from_offsets(ctx, off, ab_min, n_rows, fill);
  if (rc) return rc;
  // merge_from_offsets synchronised the stream (row count read-back): the decode status is final
  KMD_CUDA(ctx, cudaMemcpy(status.data(), m.lz_status, N * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  for (uint32_t s = 0; s < N; s++)
    if (in_size[s] && status[s]) return fail(ctx, KMD_ERR_ARG, "sample %u: LZ4 decode failed (status %u)", s, status[s]);
  return KMD_OK;
}

int merge_check(kmd_context* ctx, const void* const* streams, const uint64_t* n_records, const uint32_t* ab_min)
{
  if (!ctx) return KMD_ERR_ARG;
  if (!streams || !n_records || !ab_min) return fail(ctx, KMD_ERR_ARG, "null argument");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  return KMD_OK;
}

}  // namespace

extern "C" {

int kmd_merge_scan_host(kmd_context* ctx, const void* const* streams, const uint64_t* n_records, const uint32_t* ab_min,
                        uint64_t row_base, uint32_t partition, uint64_t* n_rows)
{
  int rc = merge_check(ctx, streams, n_records, ab_min);
  if (rc) return rc;
  uint32_t rows = 0;
  if ((rc = merge_impl(ctx, streams, n_records, ab_min, &rows))) return rc;
  if (n_rows) *n_rows = rows;
  if (rows == 0) return KMD_OK;
  return scan_device_impl(ctx, ctx->mrg.rows, rows, row_base, partition);
}

int kmd_merge_scan_device(kmd_context* ctx, const void* d_records, const uint64_t* n_records, const uint32_t* ab_min,
                          uint64_t row_base, uint32_t partition, uint64_t* n_rows)
{
  if (!ctx) return KMD_ERR_ARG;
  if (!n_records || !ab_min) return fail(ctx, KMD_ERR_ARG, "null argument");
  if (reinterpret_cast<uintptr_t>(d_records) & 3u) return fail(ctx, KMD_ERR_ARG, "d_records must be 4-byte aligned");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  if (N > 0xFFFFu) return fail(ctx, KMD_ERR_ARG, "more than 65535 samples");
  std::vector<uint64_t> off(N + 1, 0);
  for (uint32_t s = 0; s < N; s++) off[s + 1] = off[s] + n_records[s];
  if (off[N] && !d_records) return fail(ctx, KMD_ERR_ARG, "null record pointer");
  uint32_t rows = 0;
  int rc = merge_from_offsets(ctx, off, ab_min, &rows, []() -> int { return KMD_OK; }, static_cast<const uint8_t*>(d_records));
  if (rc) return rc;
  if (n_rows) *n_rows = rows;
  if (rows == 0) return KMD_OK;
  return scan_device_impl(ctx, ctx->mrg.rows, rows, row_base, partition);
}

int kmd_merge_info(kmd_context* ctx, uint32_t* algo, uint32_t* rows_mode, uint64_t* keys_in, uint64_t* keys_out)
{
  if (!ctx) return KMD_ERR_ARG;
  const MergeState& m = ctx->mrg;
  if (!m.timed) return fail(ctx, KMD_ERR_STATE, "no merge call has completed on this context");
  if (algo) *algo = static_cast<uint32_t>(m.last_algo);
  if (rows_mode) *rows_mode = m.last_algo == 0 ? static_cast<uint32_t>(m.last_rows_mode) : 0;
  if (keys_in) *keys_in = m.last_algo == 0 ? m.last_keys_in : 0;
  if (keys_out) *keys_out = m.last_algo == 0 ? m.last_keys_out : 0;
  return KMD_OK;
}

int kmd_merge_prefetch(kmd_context* ctx, const void* const* blobs, const uint64_t* blob_bytes, int frames)
{
  if (!ctx || !blobs || !blob_bytes) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  MergeState& m = ctx->mrg;
  const uint32_t N = ctx->cfg.nb_controls + ctx->cfg.nb_cases;
  uint64_t total = 0;
  std::vector<uint64_t> at(N);
  for (uint32_t s = 0; s < N; s++)
  {
    if (blob_bytes[s] && !blobs[s]) return fail(ctx, KMD_ERR_ARG, "null blob pointer (sample %u)", s);
    at[s] = total;
    total += frames ? (blob_bytes[s] + 15) / 16 * 16 : blob_bytes[s];
  }
  // a free slot: its buffer was the record / frame buffer of an earlier call whose readers (unpack / decode) completed
  // before that call returned, so the copy stream may overwrite it at once
  MergeState::Prefetch* f = nullptr;
  for (auto& cand : m.pf) if (cand.kind == -1) { f = &cand; break; }
  if (!f) return fail(ctx, KMD_ERR_STATE, "two prefetched chunks are already waiting for their merge calls");
  int rc;
  if ((rc = grow(ctx, f->buf, f->cap, total + 16))) return rc;
  if (!f->done) KMD_CUDA(ctx, cudaEventCreateWithFlags(&f->done, cudaEventDisableTiming));
  if ((rc = copy_blobs(ctx, f->buf, at, blobs, blob_bytes, N, ctx->copy_stream))) return rc;
  KMD_CUDA(ctx, cudaEventRecord(f->done, ctx->copy_stream));
  f->ptrs.assign(blobs, blobs + N);
  f->bytes.assign(blob_bytes, blob_bytes + N);
  f->kind = frames ? 1 : 0;
  return KMD_OK;
}

int kmd_merge_scan_lz4_host(kmd_context* ctx, const void* const* frames, const uint64_t* frame_bytes, const uint8_t* compressed,
                            const uint32_t* ab_min, uint64_t row_base, uint32_t partition, uint64_t max_records, uint64_t* n_rows,
                            uint64_t* n_records)
{
  int rc = merge_check(ctx, frames, frame_bytes, ab_min);
  if (rc) return rc;
  uint32_t rows = 0;
  if ((rc = merge_lz4_impl(ctx, frames, frame_bytes, compressed, ab_min, max_records, &rows, n_records, nullptr, 0))) return rc;
  if (n_rows) *n_rows = rows;
  if (rows == 0) return KMD_OK;
  return scan_device_impl(ctx, ctx->mrg.rows, rows, row_base, partition);
}

int kmd_lz4_decode_host(kmd_context* ctx, const void* const* frames, const uint64_t* frame_bytes, void* records_out,
                        uint64_t cap_records, uint64_t* n_records, float* device_ms)
{
  if (!ctx || !frames || !frame_bytes || !records_out || !n_records) return KMD_ERR_ARG;
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t rows = 0;
  const std::vector<uint32_t> ab(ctx->cfg.nb_controls + ctx->cfg.nb_cases, 1);
  int rc = merge_lz4_impl(ctx, frames, frame_bytes, nullptr, ab.data(), 0, &rows, n_records, records_out, cap_records);
  if (rc) return rc;
  if (device_ms) *device_ms = ctx->mrg.last_lz4_ms;
  return KMD_OK;
}

int kmd_merge_ms(kmd_context* ctx, float* unpack_ms, float* passes_ms, float* rows_ms, uint32_t* n_passes, uint64_t* n_records)
{
  if (!ctx) return KMD_ERR_ARG;
  MergeState& m = ctx->mrg;
  if (!m.timed) return fail(ctx, KMD_ERR_STATE, "no merge call has completed on this context");
  KMD_CUDA(ctx, cudaSetDevice(ctx->device));
  KMD_CUDA(ctx, cudaEventSynchronize(m.ev[3]));
  float a = 0, b = 0, c = 0;
  KMD_CUDA(ctx, cudaEventElapsedTime(&a, m.ev[0], m.ev[1]));
  KMD_CUDA(ctx, cudaEventElapsedTime(&b, m.ev[1], m.ev[2]));
  KMD_CUDA(ctx, cudaEventElapsedTime(&c, m.ev[2], m.ev[3]));
  if (unpack_ms) *unpack_ms = a;
  if (passes_ms) *passes_ms = b;
  if (rows_ms) *rows_ms = c;
  if (n_passes) *n_passes = m.last_passes;
  if (n_records) *n_records = m.last_records;
  return KMD_OK;
}

int kmd_merge_rows_host(kmd_context* ctx, const void* const* streams, const uint64_t* n_records, const uint32_t* ab_min,
                        void* rows_out, uint64_t cap_rows, uint64_t* n_rows)
{
  int rc = merge_check(ctx, streams, n_records, ab_min);
  if (rc) return rc;
  uint32_t rows = 0;
  if ((rc = merge_impl(ctx, streams, n_records, ab_min, &rows))) return rc;
  if (n_rows) *n_rows = rows;
  if (rows == 0) return KMD_OK;
  if (!rows_out || rows > cap_rows) return fail(ctx, KMD_ERR_ARG, "output holds %llu rows, the merge produced %u",
                                                static_cast<unsigned long long>(cap_rows), rows);
  KMD_CUDA(ctx, cudaMemcpyAsync(rows_out, ctx->mrg.rows, static_cast<size_t>(rows) * ctx->plan.row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  KMD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KMD_OK;
}

}  // extern "C"
