// A whole tree under the CPU emulator: the launch sequence of UpdateTreeImpl (xgboost_b200/csrc/updater.cu) with the
// real control, evaluation, partition and finalise kernels.  The histogram build is either a plain CPU accumulation
// over the kernels' own work lists (hist_mode 0) or the REAL histogram kernels chosen like LaunchBuildHist does
// (hist_mode 1, and 2 = with the pair-kernel hint); their inline-PTX shared atomics have emulator twins in histogram.cu
// (B200_EMU), the async-copy ring variants are not emulated.  Checks what no single-kernel test can: the order in which nodes are evaluated (colsample_bynode draws,
// the interaction-constraint buffer), categorical routing through partition and FinalisePosition, loss-guided growth.
#define B200_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "emu_io.h"

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../../xgboost_b200/csrc/colsample.cu"
#include "../../xgboost_b200/csrc/comm.cu"
#include "../../xgboost_b200/csrc/evaluate.cu"
#include "../../xgboost_b200/csrc/histogram.cu"
#include "../../xgboost_b200/csrc/partition.cu"
#include "../../xgboost_b200/csrc/updater.cu"

// The fixed-point sums wrap as the device's 64-bit atomics do (a zero hessian total quantises to INT64_MIN): the host
// restatement adds them as unsigned so that the wrap is defined behaviour here too.
static inline long long WrapAdd(long long a, long long b) { return (long long)((unsigned long long)a + (unsigned long long)b); }
static inline long long WrapSub(long long a, long long b) { return (long long)((unsigned long long)a - (unsigned long long)b); }

using namespace b200;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  // meta: n_rows, n_features, row_bytes, bin_bytes, null_symbol, has_mono, has_cat, onehot, threshold, cat_words,
  //       bin_words, max_cat_bins, n_ic_sets, n_node_sets, n_per_node, has_fs, n_fs, n_levels,
  //       [18] n_col_seeds, [19] col_n_level, [20] col_n_node, [21] bylevel_is_one, [22] bynode_is_one,
  //       [23] row_stride of a SPARSE page (0 = dense layout): bins = uint32[n][row_stride] global bins, null = total_bins
  //       [24] hist_mode
  auto meta = ReadAll<int64_t>(dir + "/meta.bin");
  auto prm = ReadAll<float>(dir + "/param_f.bin");
  auto prm_i = ReadAll<int32_t>(dir + "/param_i.bin");
  auto quant = ReadAll<double>(dir + "/quant.bin");
  auto root = ReadAll<long long>(dir + "/root.bin");  // root_g, root_h, h_zero_any
  auto bins = ReadAll<uint8_t>(dir + "/bins.bin");
  auto cut_ptrs = ReadAll<uint32_t>(dir + "/cut_ptrs.bin");
  auto cut_values = ReadAll<float>(dir + "/cut_values.bin");
  auto gq = ReadAll<longlong2>(dir + "/gq.bin");
  auto mono_c = ReadAll<int32_t>(dir + "/mono.bin", true);
  auto is_cat = ReadAll<uint8_t>(dir + "/is_cat.bin", true);
  auto ic_set_ptr = ReadAll<int32_t>(dir + "/ic_set_ptr.bin", true);
  auto ic_set_feat = ReadAll<int32_t>(dir + "/ic_set_feat.bin", true);
  auto node_sets = ReadAll<int32_t>(dir + "/node_sets.bin", true);
  auto fs = ReadAll<int32_t>(dir + "/feature_set.bin", true);
  auto col_seeds = ReadAll<uint32_t>(dir + "/col_seeds.bin", true);
  auto col_tree_set = ReadAll<int32_t>(dir + "/col_tree_set.bin", true);
  uint64_t n = (uint64_t)meta[0];
  const uint32_t nf = (uint32_t)meta[1];
  EllView view{};
  view.bins = bins.data();
  view.cut_ptrs = cut_ptrs.data();
  view.cut_values = cut_values.data();
  view.n_rows = n;
  view.n_features = nf;
  view.total_bins = (uint32_t)cut_values.size();
  view.row_bytes = (uint32_t)meta[2];
  view.bin_bytes = (uint32_t)meta[3];
  view.null_symbol = (uint32_t)meta[4];
  const bool sparse = meta.size() > 23 && meta[23] > 0;
  if (sparse) { view.is_sparse = 1; view.row_stride = (uint32_t)meta[23]; }
  const ParamDev p{prm_i[0], prm_i[1], prm_i[2], prm[0], prm[1], prm[2], prm[3], prm[4], prm[5]};
  QuantDev q{};
  q.to_fixed_g = quant[0]; q.to_fixed_h = quant[1]; q.to_float_g = quant[2]; q.to_float_h = quant[3];
  q.root_g = root[0]; q.root_h = root[1]; q.h_zero_any = (unsigned)root[2];
  q.h_lo_or = root.size() > 3 ? (unsigned)root[3] : 1u;

  // ---- what b200hist_updater_create derives on the host ----
  long long by_depth = p.max_depth > 0 ? ((1ll << (p.max_depth + 1)) - 1) : (1ll << 40);
  long long by_leaves = p.max_leaves > 0 ? (2ll * p.max_leaves - 1) : (1ll << 40);
  const int32_t max_nodes = (int32_t)std::min(by_depth, by_leaves);
  const bool dw = p.grow_policy == 0;
  const int n_tree_levels = dw ? (p.max_depth > 0 ? p.max_depth : p.max_leaves - 1) : (max_nodes - 1) / 2;
  // [26] level batch (0 = the library's default): valid entries one step expands at most -- node batching
  int level_batch = p.max_leaves > 0 ? 1024 : kMaxLevelNodes;
  if (meta.size() > 26 && meta[26] > 0) level_batch = (int)meta[26];
  std::vector<uint32_t> step_cap, step_slot_base;
  std::vector<int> step_level, step_level_start, step_children, step_reuses;
  uint32_t base = 1;
  int n_steps = 0;
  unsigned long long total_slots = 1, max_window = 0;
  for (int st = 0; st < n_tree_levels; ++st) {
    long long width = dw ? (1ll << std::min(st, 40)) : 1ll;
    if (p.max_leaves > 0) width = std::min<long long>(width, p.max_leaves);
    const bool children = p.max_depth == 0 || !dw || st + 1 < p.max_depth;
    const long long n_sub = dw ? (width + level_batch - 1) / level_batch : 1;
    for (long long b = 0; b < n_sub; ++b) {
      const uint32_t cap = (uint32_t)std::min<long long>(level_batch, width - b * level_batch);
      step_cap.push_back(cap);
      step_level.push_back(st);
      step_level_start.push_back(b == 0 ? 1 : 0);
      step_children.push_back(children ? 1 : 0);
      if (children) { total_slots += 2ull * cap; max_window = std::max<unsigned long long>(max_window, 2ull * cap); }
      ++n_steps;
    }
  }
  // [27] histogram cache capacity in slots (0 = a slot per possible node): a smaller cache is a ring whose windows wrap;
  // splits whose parent histogram has been overwritten build both children
  unsigned long long ring = total_slots - 1;
  if (meta.size() > 27 && meta[27] > 0 && (unsigned long long)meta[27] < total_slots)
    ring = std::max<unsigned long long>(max_window, (unsigned long long)meta[27] - 1);
  {
    unsigned long long pos = 0;
    bool wrapped = false;
    for (int st = 0; st < n_steps; ++st) {
      const unsigned long long w = step_children[st] ? 2ull * step_cap[st] : 0ull;
      if (pos + w > ring) { pos = 0; wrapped = true; }
      step_slot_base.push_back((uint32_t)(1 + pos));
      step_reuses.push_back(wrapped ? 1 : 0);
      pos += w;
    }
  }
  base = (uint32_t)(ring + 1);
  const uint32_t n_slots = base;

  // ---- [28] world > 1: row split over `world` PROCESSES (one per emulated GPU), like torchrun's ranks.  The histogram
  // caches, flag arrays + mailboxes and evaluation scratch tables live in one shared mapping created before the fork, so
  // that "peer memory" is what it is on the device: memory another rank's kernels load from and store to while this
  // rank's kernels run.  Rank r keeps rows [n r / world, n (r+1) / world) (or shard_ptr.bin's) and the kernels'
  // own flag barriers (comm.cu) order the ranks.  [29] = 1: the fused reduce-scatter + all-gather instead of the
  // evaluation by feature slice.
  const int world = meta.size() > 28 && meta[28] > 1 ? (int)meta[28] : 1;
  const int hist_mode_early = meta.size() > 24 ? (int)meta[24] : 0;
  const bool full_exchange = meta.size() > 29 && meta[29] == 1;
  // [29] = 2: feature-block shard (north_star's decomposition): ALL rows on every rank, rank r builds the histograms of
  // its contiguous feature block only and pushes that bin range into every peer's cache (HistExchangeKernel); evaluation,
  // partition and finalise are replicated.  hist_mode 0 only (the CPU accumulation, restricted to the block).
  const bool feature_block = meta.size() > 29 && meta[29] == 2;
  int rank = 0;
  const size_t hist_elems = (size_t)n_slots * view.total_bins;
  const size_t scratch_elems = (size_t)2 * 64 * nf;  // the emulated trees expand at most 64 nodes per step
  const size_t flag_bytes = kMailboxOffset + 2 * kMaxPeers * sizeof(MailRecord);
  auto round_up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t per_rank = round_up(hist_elems * sizeof(longlong2)) + round_up(flag_bytes) + round_up(scratch_elems * sizeof(FeatCand));
  std::vector<longlong2> hist_local;
  std::vector<FeatCand> scratch_local;
  longlong2* peer_hist[kMaxPeers] = {};
  uint32_t* peer_flags[kMaxPeers] = {};
  FeatCand* peer_scratch[kMaxPeers] = {};
  if (world > 1) {
    if (world > kMaxPeers) return 3;
    auto* shm = static_cast<uint8_t*>(mmap(nullptr, per_rank * world, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (shm == MAP_FAILED) return 4;
    std::memset(shm, 0, per_rank * world);
    for (int r = 0; r < world; ++r) {
      uint8_t* b = shm + per_rank * r;
      peer_hist[r] = reinterpret_cast<longlong2*>(b);
      peer_flags[r] = reinterpret_cast<uint32_t*>(b + round_up(hist_elems * sizeof(longlong2)));
      peer_scratch[r] = reinterpret_cast<FeatCand*>(b + round_up(hist_elems * sizeof(longlong2)) + round_up(flag_bytes));
    }
    auto shard_ptr = ReadAll<int64_t>(dir + "/shard_ptr.bin", true);
    std::vector<pid_t> kids;
    bool child = false;
    for (int r = 0; r < world; ++r) {
      const pid_t pid = fork();  // before the first launch: the emulator's thread pool does not exist yet
      if (pid < 0) return 5;
      if (pid == 0) { rank = r; child = true; break; }
      kids.push_back(pid);
    }
    if (!child) {
      int bad = 0;
      for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 6;
      }
      return bad;
    }
    if (!feature_block) {
    const uint64_t r0 = shard_ptr.empty() ? n * (uint64_t)rank / world : (uint64_t)shard_ptr[rank];
    const uint64_t r1 = shard_ptr.empty() ? n * (uint64_t)(rank + 1) / world : (uint64_t)shard_ptr[rank + 1];
    bins = std::vector<uint8_t>(bins.begin() + r0 * view.row_bytes, bins.begin() + r1 * view.row_bytes);
    gq = std::vector<longlong2>(gq.begin() + r0, gq.begin() + r1);
    view.bins = bins.data();
    n = r1 - r0;
    view.n_rows = n;
    // what QuantConvertKernel leaves on this rank: the sums and hessian flags of ITS rows (ScalarExchangeKernel folds them)
    unsigned long long sg = 0, sh = 0;
    unsigned lo_or = 0, zero_any = 0;
    for (auto const& v : gq) {
      sg += (unsigned long long)v.x;
      sh += (unsigned long long)v.y;
      lo_or |= (unsigned)((unsigned long long)v.y & 0xFFFFFFFFull);
      zero_any |= v.y == 0 ? 1u : 0u;
    }
    q.root_g = (long long)sg; q.root_h = (long long)sh; q.h_lo_or = lo_or; q.h_zero_any = zero_any;
    }
  } else {
    hist_local.assign(hist_elems, make_longlong2(0, 0));
    scratch_local.resize(scratch_elems);
  }
  longlong2* const hist_p = world > 1 ? peer_hist[rank] : hist_local.data();
  FeatCand* const scratch_p = world > 1 ? peer_scratch[rank] : scratch_local.data();
  const bool sliced = world > 1 && !full_exchange && !feature_block && meta[6] == 0 && nf >= (uint32_t)world;
  // the block of rank r: ShardGroups' contiguous split (comm.cu), over features here (the product splits 32-feature groups)
  uint32_t fb0 = 0, fb1 = nf;
  if (feature_block && world > 1) {
    const uint32_t bs = nf / (uint32_t)world, rem = nf % (uint32_t)world, r = (uint32_t)rank;
    fb0 = r * bs + std::min(r, rem);
    fb1 = fb0 + bs + (r < rem ? 1u : 0u);
    if (hist_mode_early != 0) return 7;
  }
  const EvalShard shard{sliced ? rank : 0, sliced ? world : 1, nf};
  uint32_t slice_lo = 0, slice_hi = view.total_bins;
  if (sliced) {  // SliceBins (updater.cu): the bins of the features this rank owns
    uint32_t f0 = nf, f1 = 0;
    for (uint32_t f = 0; f < nf; ++f)
      if (shard.Owns((int)f)) { f0 = std::min(f0, f); f1 = std::max(f1, f + 1); }
    slice_lo = f0 < f1 ? cut_ptrs[f0] : 0u;
    slice_hi = f0 < f1 ? cut_ptrs[f1] : 0u;
  }
  uint32_t epoch = 0;
  unsigned int done_counter = 0;
  const std::string out_suffix = world > 1 ? "_r" + std::to_string(rank) : "";
  const size_t part_tiles = n / kPartTile + (size_t)max_nodes + 2, hist_tiles = n / kHistTile + 2 * (size_t)max_nodes + 2;
  std::vector<uint32_t> ridx0(n), ridx1(n);
  for (uint64_t i = 0; i < n; ++i) ridx0[i] = (uint32_t)i;
  std::vector<int32_t> position(n, -12345);
  TreeState ts{};
  std::vector<NodeDev> nodes((size_t)max_nodes);
  std::vector<b200hist_candidate> cand((size_t)max_nodes);
  std::vector<PartEntry> part(kMaxLevelNodes);
  std::vector<BuildEntry> build(2 * kMaxLevelNodes), fin((size_t)max_nodes);
  std::vector<int32_t> slot_owner((size_t)n_slots, -1);
  std::vector<SubEntry> sub(kMaxLevelNodes);
  std::vector<int32_t> eval_nodes(2 * kMaxLevelNodes);
  std::vector<uint32_t> part_tile_seg(part_tiles), build_tile_seg(hist_tiles), tile_left(part_tiles), tile_excl(part_tiles + 1),
      tile_left_off(part_tiles), flags(part_tiles * (kPartTile / 32)), fin_tile_seg(part_tiles), fin_tile_begin((size_t)max_nodes + 1);
  std::vector<float> lower((size_t)max_nodes), upper((size_t)max_nodes);
  const MonoView mono{meta[5] ? mono_c.data() : nullptr, lower.data(), upper.data()};
  const bool has_cat = meta[6] != 0;
  std::vector<uint32_t> cat_bits((size_t)max_nodes * std::max<int64_t>(1, meta[9])), bin_bits((size_t)max_nodes * std::max<int64_t>(1, meta[10]));
  std::vector<uint8_t> node_is_cat((size_t)max_nodes, 0);
  const CatView catv = has_cat ? CatView{is_cat.data(), (int32_t)meta[7], (int32_t)meta[8], cat_bits.data(), bin_bits.data(),
                                         node_is_cat.data(), (int32_t)meta[9], (int32_t)meta[10], (int32_t)std::max<int64_t>(1, meta[11])}
                               : CatView{nullptr, 0, 0, nullptr, nullptr, nullptr, 0, 0, 0};
  const CatRoute cr{has_cat ? bin_bits.data() : nullptr, (int32_t)meta[10]};
  const bool has_ic = meta[12] > 0;
  const int ic_words = (int)((nf + 63) / 64);
  std::vector<unsigned long long> ic_bits((size_t)max_nodes * ic_words, 0ull);
  std::vector<int32_t> ic_list(nf), ic_count(2 * kMaxLevelNodes);
  const bool has_node_sets = meta[13] > 0;
  const int32_t n_per_node = (int32_t)meta[14];
  std::vector<int32_t> node_list((size_t)2 * 64 * std::max(1, n_per_node)), node_count(2 * 64);
  int32_t node_seq = 0;
  const int32_t* d_fs = meta[15] ? fs.data() : nullptr;
  const int32_t n_fs = (int32_t)meta[16], n_levels = (int32_t)meta[17];
  const size_t cat_smem = (size_t)kEvalWarps * (size_t)catv.max_cat_bins * (sizeof(float) + sizeof(uint16_t));
  // seed-driven column sampler (b200hist_updater_set_column_sampler)
  const bool has_col = meta[18] > 0;
  const int32_t col_n_level = (int32_t)meta[19], col_n_node = (int32_t)meta[20];
  const int32_t col_max_levels = p.max_depth > 0 ? p.max_depth + 1 : std::max(2, p.max_leaves);
  std::vector<int32_t> col_cursor(2, 0), col_level_sets((size_t)col_max_levels * std::max(1, col_n_level), -9),
      col_level_drawn((size_t)col_max_levels, 0), col_node_list((size_t)2 * 64 * std::max(1, col_n_node), -9),
      col_node_count(2 * 64, -9);
  const ColDrawDev cold{col_seeds.data(), (int32_t)meta[18], col_cursor.data(), col_tree_set.data(),
                        (int32_t)col_tree_set.size(), col_level_sets.data(), col_n_level, col_level_drawn.data(),
                        col_max_levels, (int32_t)meta[21]};

  // ---- the real histogram kernels, selected like LaunchBuildHist (default implementation) ----
  const int hist_mode = meta.size() > 24 ? (int)meta[24] : 0;
  // [25] page_rows > 0: external memory -- every phase that reads the matrix runs once per page on a view of a COPY of
  // that page alone (exact-size heap buffer: a read outside the page is an ASan report), like UpdateTreeImpl's
  // for_each_page over the staging buffers
  const uint64_t page_rows = meta.size() > 25 ? (uint64_t)meta[25] : 0;
  auto for_each_page = [&](auto&& fn) {
    if (page_rows == 0) { fn(view); return; }
    for (uint64_t lo = 0; lo < n; lo += page_rows) {
      const uint64_t hi = std::min<uint64_t>(n, lo + page_rows);
      std::vector<uint8_t> stage(bins.begin() + lo * view.row_bytes, bins.begin() + hi * view.row_bytes);
      EllView pv = view;
      pv.bins = stage.data() - lo * (uint64_t)view.row_bytes;
      pv.bins_t = nullptr;
      pv.page_lo = (uint32_t)lo;
      pv.page_hi = (uint32_t)hi;
      fn(pv);
    }
  };
  b200hist_ellpack_s ell_host;
  ell_host.n_rows = n; ell_host.n_features = nf; ell_host.total_bins = view.total_bins; ell_host.row_bytes = view.row_bytes;
  ell_host.bin_bytes = view.bin_bytes; ell_host.null_symbol = view.null_symbol;
  ell_host.cut_ptrs = cut_ptrs; ell_host.cut_values = cut_values;
  ell_host.is_sparse = sparse; ell_host.row_stride = view.row_stride;
  ell_host.max_feature_bins = 0;
  for (uint32_t f = 0; f < nf; ++f) ell_host.max_feature_bins = std::max(ell_host.max_feature_bins, cut_ptrs[f + 1] - cut_ptrs[f]);
  const std::vector<GroupDesc> groups = hist_mode ? MakeFeatureGroups(ell_host, 0) : std::vector<GroupDesc>();
  auto build_hist_kernels = [&](const EllView& view) {
    const bool paged = view.page_hi != 0xFFFFFFFFu;
    const int n_groups = (int)groups.size();
    if (sparse) {
      uint32_t b0 = 0xFFFFFFFFu, b1 = 0;
      for (auto const& g : groups) { b0 = std::min(b0, g.bin_begin); b1 = std::max(b1, g.bin_end); }
      EMU_LAUNCH((3), (256), 0, HistKernelSparse(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, hist_p, b0, b1));
      return;
    }
    const bool banked = view.bin_bytes == 1 || (view.bin_bytes == 2 && !groups.empty() && groups[0].n_blocks > 0);
    if (!banked) {
      uint32_t max_bins = 0;
      for (auto const& g : groups) max_bins = std::max(max_bins, g.bin_end - g.bin_begin);
      EMU_LAUNCH((2, (unsigned)n_groups), (kHistThreads), GenericSmemBytes(max_bins),
                 HistKernelGeneric<uint16_t>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), hist_p));
      return;
    }
    const uint32_t sym_bytes = view.bin_bytes;
    size_t smem = 0;
    for (auto const& g : groups) smem = std::max(smem, BankedSmemBytes(g));
    bool pair = hist_mode == 2 && n_groups >= 2 && view.row_bytes % (8u * sym_bytes) == 0u;
    uint32_t pair_sym = 0;
    for (int i = 0; pair && i < n_groups; ++i) {
      const GroupDesc& g = groups[i];
      pair = g.n_blocks == 1 && g.byte_begin % (8u * sym_bytes) == 0u &&
             (i == 0 || (g.f_begin == groups[i - 1].f_end && g.byte_begin == groups[i - 1].byte_begin + 32u * sym_bytes));
      pair_sym = std::max(pair_sym, g.max_sym);
    }
    if (pair) {
      const size_t pair_smem = (size_t)pair_sym * 192 * 4;
      const int n_pairs = (n_groups + 1) / 2;
      if (paged && sym_bytes == 2) {
        EMU_LAUNCH((2, (unsigned)n_pairs), (kHistThreads), std::max(pair_smem, smem),
                   (HistKernelPair<2, 2, true>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), (uint32_t)n_groups, hist_p, &q)));
      } else if (paged) {
        EMU_LAUNCH((2, (unsigned)n_pairs), (kHistThreads), std::max(pair_smem, smem),
                   (HistKernelPair<3, 1, true>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), (uint32_t)n_groups, hist_p, &q)));
      } else if (sym_bytes == 2) {
        EMU_LAUNCH((2, (unsigned)n_pairs), (kHistThreads), std::max(pair_smem, smem),
                   (HistKernelPair<2, 2>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), (uint32_t)n_groups, hist_p, &q)));
      } else {
        EMU_LAUNCH((2, (unsigned)n_pairs), (kHistThreads), std::max(pair_smem, smem),
                   (HistKernelPair<3, 1>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), (uint32_t)n_groups, hist_p, &q)));
      }
      return;
    }
    if (paged && sym_bytes == 2) {
      EMU_LAUNCH((2, (unsigned)n_groups), (kHistThreads), smem,
                 (HistKernelBankedFast<2, true>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), hist_p, &q)));
    } else if (paged) {
      EMU_LAUNCH((2, (unsigned)n_groups), (kHistThreads), smem,
                 (HistKernelBankedFast<1, true>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), hist_p, &q)));
    } else if (sym_bytes == 2) {
      EMU_LAUNCH((2, (unsigned)n_groups), (kHistThreads), smem,
                 HistKernelBankedFast<2>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), hist_p, &q));
    } else {
      EMU_LAUNCH((2, (unsigned)n_groups), (kHistThreads), smem,
                 HistKernelBankedFast<1>(view, gq.data(), ridx0.data(), ridx1.data(), build.data(), build_tile_seg.data(), &ts, groups.data(), hist_p, &q));
    }
  };
  auto build_local = [&]() {  // the CPU stand-in for LaunchBuildHist over the kernels' own work list
    if (hist_mode) { for_each_page(build_hist_kernels); return; }
    for (int k = 0; k < ts.n_build; ++k) {
      const BuildEntry e = build[k];
      const uint32_t* rows = (e.buf ? ridx1.data() : ridx0.data()) + e.begin;
      longlong2* h = hist_p + (size_t)e.slot * view.total_bins;
      for (uint32_t j = 0; j < e.count; ++j) {
        const uint32_t r = rows[j];
        if (sparse) {
          const uint32_t* row = reinterpret_cast<const uint32_t*>(bins.data()) + (size_t)r * view.row_stride;
          for (uint32_t k = 0; k < view.row_stride; ++k) {
            if (row[k] >= view.total_bins) continue;
            h[row[k]].x = WrapAdd(h[row[k]].x, gq[r].x);
            h[row[k]].y = WrapAdd(h[row[k]].y, gq[r].y);
          }
          continue;
        }
        for (uint32_t f = fb0; f < fb1; ++f) {
          const uint32_t sym = view.bin_bytes == 1 ? (uint32_t)bins[(size_t)r * view.row_bytes + f]
                                                   : (uint32_t)reinterpret_cast<const uint16_t*>(bins.data() + (size_t)r * view.row_bytes)[f];
          if (sym == view.null_symbol) continue;
          h[cut_ptrs[f] + sym].x = WrapAdd(h[cut_ptrs[f] + sym].x, gq[r].x);
          h[cut_ptrs[f] + sym].y = WrapAdd(h[cut_ptrs[f] + sym].y, gq[r].y);
        }
      }
    }
  };
  // AllReduceHist of UpdateTreeImpl's build_hist: one collective per level over the level's contiguous built slots
  auto build_hist = [&](uint32_t slot_base, uint32_t slot_cap) {
    build_local();
    if (world == 1) return;
    if (feature_block) {
      ++epoch;
      EMU_LAUNCH((3), (256), 0, HistExchangeKernel(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, 1, view.total_bins, cut_ptrs[fb0], cut_ptrs[fb1], epoch, &done_counter));
      return;
    }
    if (sliced) {
      ++epoch;
      if (world == 2) { EMU_LAUNCH((3), (256), 0, HistSliceReduceKernel<2>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, slice_lo, slice_hi, epoch)); }
      else if (world == 4) { EMU_LAUNCH((3), (256), 0, HistSliceReduceKernel<4>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, slice_lo, slice_hi, epoch)); }
      else if (world == 8) { EMU_LAUNCH((3), (256), 0, HistSliceReduceKernel<8>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, slice_lo, slice_hi, epoch)); }
      else { EMU_LAUNCH((3), (256), 0, HistSliceReduceKernel<0>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, slice_lo, slice_hi, epoch)); }
      return;
    }
    const uint32_t ea = ++epoch, eb = ++epoch;
    if (world == 2) { EMU_LAUNCH((3), (256), 0, HistAllReduceKernel<2>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, ea, eb, &done_counter)); }
    else if (world == 4) { EMU_LAUNCH((3), (256), 0, HistAllReduceKernel<4>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, ea, eb, &done_counter)); }
    else if (world == 8) { EMU_LAUNCH((3), (256), 0, HistAllReduceKernel<8>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, ea, eb, &done_counter)); }
    else { EMU_LAUNCH((3), (256), 0, HistAllReduceKernel<0>(peer_hist, peer_flags, rank, world, &ts, slot_base, slot_cap, view.total_bins, ea, eb, &done_counter)); }
  };
  auto subtract = [&]() {
    for (int k = 0; k < ts.n_sub; ++k) {
      const SubEntry e = sub[k];
      for (uint32_t b = slice_lo; b < slice_hi; ++b) {  // LaunchSubtract's bin range: this rank's slice, or everything
        hist_p[(size_t)e.sub * view.total_bins + b].x = WrapSub(hist_p[(size_t)e.parent * view.total_bins + b].x, hist_p[(size_t)e.built * view.total_bins + b].x);
        hist_p[(size_t)e.sub * view.total_bins + b].y = WrapSub(hist_p[(size_t)e.parent * view.total_bins + b].y, hist_p[(size_t)e.built * view.total_bins + b].y);
      }
    }
  };
  auto evaluate = [&](int32_t max_eval, IcView ic, const int32_t* feature_set, int32_t n_feature_set, int32_t levels) {
    const int32_t stride = feature_set ? n_feature_set : (int32_t)nf;
    const long long items = (long long)max_eval * stride;
    if (items == 0) return;
    const unsigned b1 = (unsigned)std::min<long long>((items + kEvalWarps - 1) / kEvalWarps, 4);
    const unsigned b2 = (unsigned)std::min<long long>(((long long)max_eval + kEvalWarps - 1) / kEvalWarps, 2);
    const longlong2* h2 = hist_p;
    if (ic.count != nullptr) {
      EMU_LAUNCH((b1), (kEvalWarps * 32), 0, FeatureScanKernel<true>(view, h2, eval_nodes.data(), &ts, nodes.data(), feature_set, stride, p, &q, scratch_p, mono, levels, ic, shard));
    } else {
      EMU_LAUNCH((b1), (kEvalWarps * 32), 0, FeatureScanKernel<false>(view, h2, eval_nodes.data(), &ts, nodes.data(), feature_set, stride, p, &q, scratch_p, mono, levels, ic, shard));
    }
    if (has_cat)
      EMU_LAUNCH((b1), (kEvalWarps * 32), cat_smem, CatFeatureScanKernel(view, h2, eval_nodes.data(), &ts, nodes.data(), feature_set, stride, p, &q, scratch_p, mono, levels, ic, catv));
    if (sliced) {  // LaunchEvaluate's `between`: the per-feature winners this rank owns travel to every peer
      const unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((items + 255) / 256, 3));
      ++epoch;
      EMU_LAUNCH((bx), (256), 0, ScratchExchangeKernel(peer_scratch, peer_flags, shard, &ts, eval_nodes.data(), nodes.data(), feature_set, stride, levels, ic, epoch, &done_counter));
    }
    if (ic.count != nullptr) {
      EMU_LAUNCH((b2), (kEvalWarps * 32), 0, NodeReduceKernel<true>(view, eval_nodes.data(), &ts, nodes.data(), feature_set, stride, p, &q, scratch_p, cand.data(), mono, levels, &ts.max_eval_depth, ic));
    } else {
      EMU_LAUNCH((b2), (kEvalWarps * 32), 0, NodeReduceKernel<false>(view, eval_nodes.data(), &ts, nodes.data(), feature_set, stride, p, &q, scratch_p, cand.data(), mono, levels, &ts.max_eval_depth, ic));
    }
    if (has_cat)
      EMU_LAUNCH((b2), (kEvalWarps * 32), cat_smem, CatFinishKernel(view, h2, eval_nodes.data(), &ts, nodes.data(), p, &q, cand.data(), catv));
  };
  auto node_sets_view = [&]() {
    EMU_LAUNCH((1), (kCtlThreads), 0, NodeSetLevelKernel(&ts, node_sets.data(), (int32_t)meta[13], n_per_node, &node_seq, node_list.data(), node_count.data()));
    return IcView{node_list.data(), node_count.data(), n_per_node};
  };

  auto evaluate_with_column_sampler = [&](int32_t max_eval, bool is_root) {
    EMU_LAUNCH((1), (kCsThreads), sizeof(int32_t) * 2 * (size_t)col_n_level, LevelDrawKernel(&ts, nodes.data(), eval_nodes.data(), cold));
    if (has_ic && !is_root) {
      const IcDev icd{(int32_t)nf, ic_words, (int32_t)meta[12], ic_set_ptr.data(), ic_set_feat.data(), ic_bits.data(), ic_list.data(), ic_count.data()};
      if (!meta[22]) {
        EMU_LAUNCH(((unsigned)max_eval), (kCsThreads), sizeof(int32_t) * 2 * (size_t)col_n_node,
                   NodeDrawKernel(&ts, nodes.data(), eval_nodes.data(), cold, col_n_node, col_node_list.data(), col_node_count.data()));
        EMU_LAUNCH((1), (kCtlThreads), sizeof(unsigned long long) * (size_t)ic_words,
                   IcLevelKernel(&ts, nodes.data(), part.data(), eval_nodes.data(), icd, nullptr, 0, 0, col_node_list.data(), col_n_node));
      } else {
        EMU_LAUNCH((1), (kCtlThreads), sizeof(unsigned long long) * (size_t)ic_words,
                   IcLevelKernel(&ts, nodes.data(), part.data(), eval_nodes.data(), icd, col_level_sets.data(), col_n_level, col_max_levels, nullptr, 0));
      }
      evaluate(max_eval, IcView{ic_list.data(), ic_count.data(), 0}, nullptr, 0, 0);
      return;
    }
    if (!meta[22]) {
      EMU_LAUNCH(((unsigned)max_eval), (kCsThreads), sizeof(int32_t) * 2 * (size_t)col_n_node,
                 NodeDrawKernel(&ts, nodes.data(), eval_nodes.data(), cold, col_n_node, col_node_list.data(), col_node_count.data()));
      evaluate(max_eval, IcView{col_node_list.data(), col_node_count.data(), col_n_node}, nullptr, 0, 0);
    } else {
      evaluate(max_eval, IcView{nullptr, nullptr, 0}, col_level_sets.data(), col_n_level, col_max_levels);
    }
  };

  // ---- UpdateTreeImpl ----
  if (world > 1) {
    // nobody may push into a cache that its owner has not cleared yet (a pure flag barrier), then the root sum and the
    // hessian flags of all rows (the scale was derived from all rows by the caller: ScalarExchangeKernel's which = 0
    // is covered where the quantiser kernels are, tests/test_multi_gpu.py on hardware)
    ++epoch;
    EMU_LAUNCH((3), (256), 0, HistExchangeKernel(peer_hist, peer_flags, rank, world, &ts, 0, 0, 0, view.total_bins, 0, 0, epoch, &done_counter));
    if (!feature_block) ++epoch;
    if (!feature_block) EMU_LAUNCH((1), (32), 0, ScalarExchangeKernel(peer_flags, rank, world, epoch, 1, &q, (unsigned long long)n));
  }
  EMU_LAUNCH((1), (kCtlThreads), 0, InitRootKernel(p, (uint32_t)n, &q, &ts, nodes.data(), build.data(), build_tile_seg.data(), eval_nodes.data(), mono, slot_owner.data()));
  build_hist(0, 1);
  if (has_col) evaluate_with_column_sampler(1, true);
  else if (has_node_sets) evaluate(1, node_sets_view(), nullptr, 0, 0);
  else evaluate(1, IcView{nullptr, nullptr, 0}, d_fs, n_fs, n_levels);
  for (int step = 0; step < n_steps; ++step) {
    const long long cap = step_cap[step];
    if (has_cat) {
      EMU_LAUNCH((1), (kCtlThreads), 0, ApplyKernel<true>(view, p, &q, &ts, nodes.data(), cand.data(), part.data(), part_tile_seg.data(), max_nodes, mono, node_is_cat.data(), step_level_start[step], level_batch));
    } else {
      EMU_LAUNCH((1), (kCtlThreads), 0, ApplyKernel<false>(view, p, &q, &ts, nodes.data(), cand.data(), part.data(), part_tile_seg.data(), max_nodes, mono, nullptr, step_level_start[step], level_batch));
    }
    const bool children_possible = p.max_depth == 0 || !dw || step_level[step] + 1 < p.max_depth;
    if (!children_possible) continue;
    if (page_rows > 0) {  // the updater clears the accumulators before the first page
      std::fill(flags.begin(), flags.end(), 0u);
      std::fill(tile_left.begin(), tile_left.end(), 0u);
    }
    for_each_page([&](const EllView& view) {
    if (sparse) {
      if (has_cat) { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint32_t, true>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
      else { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint32_t, false>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
    } else if (view.bin_bytes == 1) {
      if (has_cat) { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint8_t, true>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
      else { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint8_t, false>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
    } else {
      if (has_cat) { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint16_t, true>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
      else { EMU_LAUNCH((3), (kPartThreads), 0, (PartitionCountKernel<uint16_t, false>(view, ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left.data(), cr))); }
    }
    });
    EMU_LAUNCH((1), (kCtlThreads), 0, ScanKernel(&ts, nodes.data(), part.data(), part_tile_seg.data(), tile_left.data(), tile_excl.data(), tile_left_off.data(), build.data(), sub.data(), eval_nodes.data(), build_tile_seg.data(), step_slot_base[step], step_cap[step], world > 1 && !feature_block ? 1 : 0, slot_owner.data()));
    EMU_LAUNCH((3), (kPartThreads), 0, PartitionScatterKernel(ridx0.data(), ridx1.data(), ridx0.data(), ridx1.data(), part.data(), part_tile_seg.data(), &ts, flags.data(), tile_left_off.data()));
    if (step_reuses[step])  // a reused window holds other nodes' sums
      std::fill(hist_p + (size_t)step_slot_base[step] * view.total_bins,
                hist_p + (size_t)(step_slot_base[step] + 2 * step_cap[step]) * view.total_bins, make_longlong2(0, 0));
    build_hist(step_slot_base[step], ring < total_slots - 1 ? 2 * step_cap[step] : step_cap[step]);
    subtract();
    IcView icv{nullptr, nullptr, 0};
    if (has_col) {
      evaluate_with_column_sampler((int32_t)(2 * cap), false);
      continue;
    }
    if (has_node_sets) {
      icv = node_sets_view();
      evaluate((int32_t)(2 * cap), icv, nullptr, 0, 0);
      continue;
    }
    if (has_ic) {
      const IcDev icd{(int32_t)nf, ic_words, (int32_t)meta[12], ic_set_ptr.data(), ic_set_feat.data(), ic_bits.data(), ic_list.data(), ic_count.data()};
      EMU_LAUNCH((1), (kCtlThreads), sizeof(unsigned long long) * (size_t)ic_words, IcLevelKernel(&ts, nodes.data(), part.data(), eval_nodes.data(), icd, d_fs, n_fs, n_levels, nullptr, 0));
      icv = IcView{ic_list.data(), ic_count.data(), 0};
    }
    evaluate((int32_t)(2 * cap), icv, d_fs, n_fs, n_levels);
  }
  EMU_LAUNCH((1), (kCtlThreads), 0, FinalListKernel(&ts, nodes.data(), fin.data(), fin_tile_seg.data(), fin_tile_begin.data()));
  for_each_page([&](const EllView& view) {
  const uint32_t r_lo = page_rows ? view.page_lo : 0u, r_hi = page_rows ? view.page_hi : (uint32_t)n;
  if (sparse) {
    if (has_cat) { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint32_t, true>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
    else { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint32_t, false>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
  } else if (view.bin_bytes == 1) {
    if (has_cat) { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint8_t, true>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
    else { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint8_t, false>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
  } else {
    if (has_cat) { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint16_t, true>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
    else { EMU_LAUNCH((2), (256), 0, (FinaliseWindowKernel<uint16_t, false>(view, ridx0.data(), ridx1.data(), gq.data(), nodes.data(), fin.data(), &ts, position.data(), &q, r_lo, r_hi, cr))); }
  }
  });
  // ---- results ----
  std::vector<b200hist_tree_node> out((size_t)ts.n_nodes);
  for (int32_t i = 0; i < ts.n_nodes; ++i) {
    const NodeDev& s = nodes[i];
    b200hist_tree_node d{};
    d.parent = s.parent; d.left = s.left; d.right = s.right; d.findex = s.findex;
    d.fvalue = s.is_leaf ? s.leaf_value : s.fvalue;
    d.default_left = s.default_left; d.is_leaf = s.is_leaf; d.leaf_value = s.leaf_value;
    d.loss_chg = s.loss_chg; d.sum_hess = s.sum_hess; d.base_weight = s.base_weight; d.depth = s.depth;
    out[i] = d;
  }
  WriteAll(dir + "/nodes.bin" + out_suffix, out);
  WriteAll(dir + "/position.bin" + out_suffix, position);
  std::vector<int32_t> flags_out((size_t)ts.n_nodes);
  for (int32_t i = 0; i < ts.n_nodes; ++i) flags_out[i] = (!nodes[i].is_leaf && node_is_cat[i]) ? 1 : 0;
  WriteAll(dir + "/is_cat.bin" + out_suffix, flags_out);
  WriteAll(dir + "/cat_bits.bin" + out_suffix, cat_bits);
  WriteAll(dir + "/eval_count.bin" + out_suffix, std::vector<int32_t>{node_seq, ts.max_eval_depth, col_cursor[0] + col_cursor[1]});
  return 0;
}
