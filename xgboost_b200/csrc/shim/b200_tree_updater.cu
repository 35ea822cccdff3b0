// Reference-side binding: xgboost::TreeUpdater served by libb200hist.so.
//
// This translation unit is the ONLY file compiled against the reference's headers.  It subclasses
// xgboost::TreeUpdater (/root/reference/include/xgboost/tree_updater.h:37-108) and registers itself in the
// reference's dmlc registry as "grow_gpu_hist" -- the name GBTree::Configure resolves for
// tree_method="hist" on CUDA (src/gbm/gbtree.cc:45-66,129-139) -- and as "grow_b200_hist" so both
// implementations can be A/B-ed inside one process (param updater="grow_b200_hist").
// dmlc::Registry::__REGISTER__ returns the existing entry for a duplicate name
// (dmlc-core/include/dmlc/registry.h:77-86) and set_body() overwrites its factory, so loading this shared
// object AFTER libxgboost.so replaces the factory without touching the reference.
//
// Everything it does is marshalling: raw device pointers and scalars out of the reference's objects into
// the C ABI (include/b200hist.h), and the returned node array back into RegTree::ExpandNode calls.
//
// Build (tools/build_shim.sh): nvcc -shared against /root/reference/{include,src,dmlc-core/include} and
// linked with oracle/_ref_cuda/libxgboost.so + xgboost_b200/lib/libb200hist.so.
#include <dmlc/registry.h>
#include <xgboost/base.h>
#include <xgboost/context.h>
#include <xgboost/data.h>
#include <xgboost/gradient.h>
#include <xgboost/host_device_vector.h>
#include <xgboost/json.h>
#include <xgboost/linalg.h>
#include <xgboost/logging.h>
#include <xgboost/tree_model.h>
#include <xgboost/tree_updater.h>

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <type_traits>
#include <utility>
#include <string>
#include <vector>

#include "../../../include/b200hist.h"
#include "src/collective/allgather.h"       // collective::Allgather (IPC handle exchange)
#include "src/collective/allreduce.h"       // collective::Allreduce (barrier before unmapping)
#include "src/collective/communicator-inl.h"  // IsDistributed, GetRank, GetWorldSize, GetProcessorName
#include "src/common/bitfield.h"            // CatBitField
#include "src/common/categorical.h"         // CatBitField alias
#include "src/common/cuda_context.cuh"      // CUDAContext::Stream
#include "src/common/random.h"              // ColumnSampler (the reference's own colsample_bytree draw)
#include "src/data/ellpack_page.cuh"        // EllpackPageImpl
#include "src/tree/hist/hist_param.h"       // HistMakerTrainParam
#include "src/tree/param.h"                 // TrainParam
#include "src/tree/updater_gpu_common.cuh"  // HistBatch

namespace xgboost::tree {

#define B200_SHIM_CHECK(call)                                          \
  do {                                                                 \
    if ((call) != 0) LOG(FATAL) << "[b200hist] " << b200hist_last_error(); \
  } while (0)

class B200HistMaker : public TreeUpdater {
 public:
  explicit B200HistMaker(Context const* ctx, ObjInfo const* task) : TreeUpdater(ctx), task_{task} {}
  ~B200HistMaker() override { this->Release(); }

  void Configure(Args const& args) override {
    hist_maker_param_.UpdateAllowUnknown(args);  // unknown keys pass through like the reference
    // max_cached_hist_node (hist_param.h:18-45): honoured when the user sets it; otherwise the library sizes its
    // histogram cache from the free device memory (the reference's CUDA default of 4096 nodes would make deep trees
    // rebuild histograms that fit comfortably)
    for (auto const& kv : args) {
      if (kv.first == "max_cached_hist_node") max_cached_hist_node_ = std::max<long long>(0, std::atoll(kv.second.c_str()));
    }
    this->Release();                              // caches are invalidated by Configure (updater_gpu_hist.cu:764-771)
  }
  void LoadConfig(Json const& in) override {
    auto const& config = get<Object const>(in);
    FromJson(config.at("hist_train_param"), &this->hist_maker_param_);
    this->Release();
  }
  void SaveConfig(Json* p_out) const override { (*p_out)["hist_train_param"] = ToJson(hist_maker_param_); }

  [[nodiscard]] char const* Name() const override { return "grow_gpu_hist"; }
  [[nodiscard]] bool HasNodePosition() const override { return true; }

  void Update(TrainParam const* param, GradientContainer* in_gpair, DMatrix* p_fmat,
              common::Span<HostDeviceVector<bst_node_t>> out_position,
              std::vector<RegTree*> const& trees) override {
    CHECK_EQ(in_gpair->gpair.Shape(1), 1) << "[b200hist] single-target trees only";
    in_gpair->gpair.SetDevice(ctx_->Device());
    this->InitDataOnce(param, p_fmat);
    auto stream = static_cast<void*>(ctx_->CUDACtx()->Stream());
    std::size_t t_idx{0};
    for (RegTree* p_tree : trees) {
      auto& pos = out_position[t_idx];
      pos.SetDevice(ctx_->Device());
      pos.Resize(p_fmat->Info().num_row_);
      // GradientPair = {float grad, float hess}: exactly the n_rows x 2 float32 layout of the ABI
      auto const* gpair = reinterpret_cast<float const*>(in_gpair->gpair.Data()->ConstDevicePointer());
      this->DrawSampling(param, p_fmat);
      std::int32_t n_nodes{0};
      B200_SHIM_CHECK(b200hist_updater_update_tree(updater_, gpair, stream, nodes_.data(), &n_nodes,
                                                   pos.DevicePointer()));
      this->FinishSampling(param, p_fmat);
      this->ReplayInto(p_tree, n_nodes);
      hist_maker_param_.CheckTreesSynchronized(ctx_, p_tree);
      ++t_idx;
    }
    p_last_fmat_ = p_fmat;
    p_last_tree_ = trees.empty() ? nullptr : trees.back();
  }

  bool UpdatePredictionCache(DMatrix const* p_fmat, common::Span<HostDeviceVector<bst_node_t>> out_position,
                             linalg::MatrixView<float> p_out_preds) override {
    if (updater_ == nullptr || p_last_fmat_ == nullptr || p_last_fmat_ != p_fmat) return false;
    if (out_position.size() > 1) return false;
    CHECK(p_last_tree_ != nullptr);
    CHECK(p_out_preds.Device().IsCUDA());
    CHECK_EQ(p_out_preds.Device().ordinal, ctx_->Ordinal());
    CHECK_EQ(out_position.size(), 1);
    CHECK_EQ(p_out_preds.Shape(1), 1);
    CHECK_EQ(p_out_preds.Size(), out_position.front().Size());
    // updater_gpu_hist.cu:582-608.  Two things the reference's call depends on and a plain `preds[i] += leaf` misses:
    //  * the view is out.Slice(All, Range(gid, gid + 1)) of the (n_rows, n_groups) cache (gbtree.cc DoBoost): element
    //    (i, 0) sits at Values().data() + i * Stride(0);
    //  * the leaf values come from the RegTree, because ObjFunction::UpdateTreeLeaf (adaptive objectives such as
    //    reg:absoluteerror) rewrites them between Update and this call.
    auto const& tree_nodes = p_last_tree_->GetNodes(DeviceOrd::CPU());
    leaf_values_.assign(tree_nodes.size(), 0.0f);
    for (std::size_t i = 0; i < tree_nodes.size(); ++i) {
      if (tree_nodes[i].IsLeaf()) leaf_values_[i] = tree_nodes[i].LeafValue();
    }
    B200_SHIM_CHECK(b200hist_updater_update_prediction_cache_ex(
        updater_, p_out_preds.Values().data(), static_cast<std::uint64_t>(p_out_preds.Stride(0)),
        out_position.front().ConstDevicePointer(), leaf_values_.data(), static_cast<std::int32_t>(leaf_values_.size()),
        static_cast<void*>(ctx_->CUDACtx()->Stream())));
    return true;
  }

 private:
  // InitDataOnce (updater_gpu_hist.cu:803-827): borrow the reference's EllpackPage and cuts.
  void InitDataOnce(TrainParam const* param, DMatrix* p_fmat) {
    if (updater_ != nullptr && p_last_fmat_ == p_fmat) return;
    this->Release();
    // In-core DMatrix: exactly one page.  ExtMemQuantileDMatrix: the pages come one at a time (host or disk cache);
    // each is decoded into its row range of ONE resident page of ours -- an external-memory matrix that fits HBM
    // trains like an in-core one (histograms are integer sums: the result does not depend on the paging).
    std::uint64_t const n_total = p_fmat->Info().num_row_;
    for (auto const& page : p_fmat->GetBatches<EllpackPage>(ctx_, cuda_impl::HistBatch(*param))) {
      auto const* impl = page.Impl();
      if (ellpack_ == nullptr) {
        auto cuts = impl->CutsShared();
        auto const& ptrs = cuts->cut_ptrs_.ConstHostVector();
        auto const& vals = cuts->cut_values_.ConstHostVector();
        ellpack_cut_bins_.clear();
        for (std::size_t f = 0; f + 1 < ptrs.size(); ++f) ellpack_cut_bins_.push_back(ptrs[f + 1] - ptrs[f]);
        if (impl->IsDenseCompressed()) {
          int const has_missing = impl->IsDense() ? 0 : 1;
          // External memory (ExtMemQuantileDMatrix: several pages) larger than the resident budget: our page set keeps
          // the matrix in pinned host memory and the updater streams it (b200hist_ellpack_create_paged); within the
          // budget the pages are decoded into ONE resident page as before.  Budget = B200HIST_RESIDENT_BUDGET_BYTES,
          // default 60 % of the free device memory.
          auto const n_feat = static_cast<std::uint32_t>(ptrs.size() - 1);
          std::uint64_t const row_bytes =
              static_cast<std::uint64_t>((n_feat + 31u) & ~31u) * (impl->NumSymbols() <= 256 ? 1u : 2u);
          std::uint64_t budget = 0;
          if (char const* e = std::getenv("B200HIST_RESIDENT_BUDGET_BYTES")) budget = std::strtoull(e, nullptr, 10);
          if (budget == 0) {
            std::size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            budget = static_cast<std::uint64_t>(free_b) * 6 / 10;
          }
          if (!p_fmat->SingleColBlock() && n_total * row_bytes > budget) {
            std::uint64_t const page_rows = std::max<std::uint64_t>(4096, budget / (2 * row_bytes));
            B200_SHIM_CHECK(b200hist_ellpack_create_paged(ctx_->Ordinal(), n_total, n_feat, ptrs.data(), vals.data(),
                                                          has_missing, page_rows, &ellpack_));
          } else {
            B200_SHIM_CHECK(b200hist_ellpack_create(ctx_->Ordinal(), n_total, n_feat, ptrs.data(), vals.data(),
                                                    has_missing, &ellpack_));
          }
        } else {  // sparse ELLPACK: global-bin symbols, row_stride < n_features (ellpack_page.cu:140-183)
          B200_SHIM_CHECK(b200hist_ellpack_create_sparse(ctx_->Ordinal(), n_total,
                                                         static_cast<std::uint32_t>(ptrs.size() - 1), ptrs.data(),
                                                         vals.data(), static_cast<std::uint32_t>(impl->info.row_stride),
                                                         &ellpack_));
        }
      }
      // ExtMem "direct access" read path (ellpack_page_source.cu:349-360): the page is split between gidx_buffer (host
      // part) and d_gidx_buffer (device part) and read through DoubleCompressedIter; decoding n_rows rows from
      // gidx_buffer alone would run past its end.
      CHECK(impl->d_gidx_buffer.empty())
          << "[b200hist] Ellpack pages split between a host and a device buffer (ExtMem direct-access cache) are not "
             "imported; build the ExtMemQuantileDMatrix with cache_host_ratio=1.0 or 0.0";
      B200_SHIM_CHECK(b200hist_ellpack_import_reference_rows(
          ellpack_, static_cast<std::uint64_t>(impl->base_rowid), static_cast<std::uint64_t>(impl->n_rows),
          impl->gidx_buffer.data(), static_cast<std::uint32_t>(impl->info.row_stride),
          static_cast<std::uint32_t>(impl->NumSymbols()), impl->IsDense() ? 1 : 0,
          static_cast<void*>(ctx_->CUDACtx()->Stream())));
      // the reference may recycle the page's buffer as soon as the iterator advances
      ctx_->CUDACtx()->Stream().Sync();
    }
    CHECK(ellpack_ != nullptr) << "[b200hist] the DMatrix yielded no Ellpack page";
    b200hist_train_param p{};
    p.max_depth = param->max_depth;
    p.max_leaves = param->max_leaves;
    p.grow_policy = param->grow_policy == TrainParam::kDepthWise ? 0 : 1;
    p.learning_rate = param->learning_rate;
    p.reg_lambda = param->reg_lambda;
    p.reg_alpha = param->reg_alpha;
    p.min_child_weight = param->min_child_weight;
    p.max_delta_step = param->max_delta_step;
    p.min_split_loss = param->min_split_loss;
    p.reserved[0] = static_cast<std::int32_t>(std::min<long long>(max_cached_hist_node_, 0x7fffffff));
    CHECK(param->colsample_bynode == 1.0f ||
          (param->colsample_bylevel == 1.0f && param->interaction_constraints.empty()) ||
          (p_fmat->Info().feature_weights.Empty() && p_fmat->Info().num_col_ <= 8192))
        << "[b200hist] colsample_bynode with colsample_bylevel or interaction constraints needs the device sampler "
           "(no feature weights, <= 8192 features)";
    B200_SHIM_CHECK(b200hist_updater_create(ellpack_, &p, &updater_));
    // Sampler ctor (sampler.cu:343-363): the strategy follows sampling_method whenever subsample < 1
    B200_SHIM_CHECK(b200hist_updater_set_sampling_method(
        updater_, param->subsample < 1.0f && param->sampling_method == TrainParam::kGradientBased ? 1 : 0));
    this->ConnectWorkers(p_fmat);
    if (!param->monotone_constraints.empty()) {  // TreeEvaluator ctor (split_evaluator.h:37-57)
      std::vector<std::int32_t> mono(param->monotone_constraints.cbegin(), param->monotone_constraints.cend());
      CHECK_LE(mono.size(), p_fmat->Info().num_col_)
          << "The size of monotone constraint should be less or equal to the number of features.";
      B200_SHIM_CHECK(b200hist_updater_set_monotone_constraints(updater_, mono.data(),
                                                                static_cast<std::uint32_t>(mono.size())));
    }
    if (!param->interaction_constraints.empty()) {  // FeatureInteractionConstraintDevice::Configure (constraints.cu:25-44)
      std::vector<std::vector<bst_feature_t>> sets;
      ParseInteractionConstraint(param->interaction_constraints, &sets);
      std::vector<std::int32_t> set_ptr{0}, set_feat;
      for (auto const& set : sets) {
        for (auto f : set) set_feat.push_back(static_cast<std::int32_t>(f));
        set_ptr.push_back(static_cast<std::int32_t>(set_feat.size()));
      }
      if (set_feat.empty()) set_feat.push_back(0);
      B200_SHIM_CHECK(b200hist_updater_set_interaction_constraints(updater_, set_ptr.data(), set_feat.data(),
                                                                   static_cast<std::uint32_t>(sets.size())));
    }
    has_categorical_ = p_fmat->Info().HasCategorical();
    if (has_categorical_) {
      auto const& ft = p_fmat->Info().feature_types.ConstHostVector();
      std::vector<std::uint8_t> flags(ft.size());
      for (std::size_t i = 0; i < ft.size(); ++i) flags[i] = ft[i] == FeatureType::kCategorical ? 1 : 0;
      B200_SHIM_CHECK(b200hist_updater_set_categorical(updater_, flags.data(), static_cast<std::uint32_t>(flags.size()),
                                                       static_cast<std::int32_t>(param->max_cat_to_onehot),
                                                       static_cast<std::int32_t>(param->max_cat_threshold)));
      B200_SHIM_CHECK(b200hist_updater_cat_words(updater_, &cat_words_));
    }
    std::int32_t max_nodes{0};
    B200_SHIM_CHECK(b200hist_updater_max_nodes(updater_, &max_nodes));
    nodes_.resize(static_cast<std::size_t>(max_nodes));
    if (has_categorical_) {
      node_is_cat_.resize(static_cast<std::size_t>(max_nodes));
      cat_bits_.resize(static_cast<std::size_t>(max_nodes) * cat_words_);
    }
  }

  // Distributed training (collective::IsDistributed(), data split by rows -- the reference's default): the reference
  // sums every level's partial histograms with collective::Allreduce -> ncclAllReduce (histogram.cu:598-609), takes the
  // quantiser's scale and the root sum from collective::GlobalSum (quantiser.cu:88-123, updater_gpu_hist.cu:661-668).
  // Ours does the same three exchanges inside its own kernels over peer memory: every worker's histogram cache is
  // IPC-mapped into the others (b200hist_updater_ipc_handle / set_peers; the 64-byte handles travel through the
  // reference's OWN communicator, collective::Allgather on the CPU backend), then b200hist_updater_set_row_split.
  // Requirements checked here, loudly: row split, one GPU per worker, all workers on one host (NVLink peer access).
  void ConnectWorkers(DMatrix* p_fmat) {
    if (!collective::IsDistributed()) return;
    CHECK(p_fmat->Info().IsRowSplit())
        << "[b200hist] distributed training is implemented for row-split data only (column split: use the "
           "reference's updater, updater=grow_gpu_hist_reference)";
    CHECK(!collective::IsFederated()) << "[b200hist] federated learning is not supported";
    auto const world = collective::GetWorldSize();
    auto const rank = collective::GetRank();
    // same host?  compare a hash of the processor name of every worker
    auto const name = collective::GetProcessorName();
    std::vector<std::uint64_t> hosts(static_cast<std::size_t>(world), 0);
    hosts[rank] = std::hash<std::string>{}(name);
    auto rc = collective::Allgather(ctx_, linalg::MakeVec(hosts.data(), hosts.size()));
    collective::SafeColl(rc);
    for (auto h : hosts) {
      CHECK_EQ(h, hosts[rank]) << "[b200hist] workers on different hosts: the peer-memory histogram exchange needs all "
                                  "GPUs in one NVLink domain of one node";
    }
    std::vector<std::uint8_t> handles(static_cast<std::size_t>(world) * 64, 0);
    B200_SHIM_CHECK(b200hist_updater_ipc_handle(updater_, handles.data() + static_cast<std::size_t>(rank) * 64));
    rc = collective::Allgather(ctx_, linalg::MakeVec(handles.data(), handles.size()));
    collective::SafeColl(rc);
    B200_SHIM_CHECK(b200hist_updater_set_peers(updater_, rank, world, handles.data()));
    B200_SHIM_CHECK(b200hist_updater_set_row_split(updater_, 1));
    connected_ = true;
  }

  // GPUHistMakerDevice::Reset (updater_gpu_hist.cu:172-211) draws from ctx->Rng() in this order for every tree:
  // the row sampler's seed (sampler.cu:165, only when subsample < 1), then ColumnSampler::Init's shuffle seed
  // (random.cu:74, only when colsample_bytree < 1).  The column draw is the reference's own ColumnSampler.
  void DrawSampling(TrainParam const* param, DMatrix* p_fmat) {
    bool const rows = param->subsample < 1.0f;
    bool const cols = param->colsample_bytree < 1.0f;
    bool const levels = param->colsample_bylevel < 1.0f;
    bool const by_node = param->colsample_bynode < 1.0f;
    if (!rows && !cols && !levels && !by_node && !sampling_set_) return;
    std::uint32_t seed = rows ? static_cast<std::uint32_t>(ctx_->Rng()()) : 0u;
    auto const& info = p_fmat->Info();
    std::vector<std::int32_t> fs;
    level_draws_pending_ = false;
    node_draws_pending_ = false;
    if (by_node) {
      // ColumnSampler::GetFeatureSet draws a fresh set per call (random.h:156-175; one ctx->Rng()() each, random.cu:67):
      // once for the root, then for the left and right child of every valid candidate.  The stream does not depend on
      // the data, only its length does: draw the longest possible stream from a COPY of the engine, let the updater
      // take draw k for the k-th evaluated node, and replay the real engine by the count it reports.
      rng_before_columns_ = ctx_->Rng();
      std::int32_t max_nodes{0};
      B200_SHIM_CHECK(b200hist_updater_max_nodes(updater_, &max_nodes));
      std::size_t per_node = 0;
      node_draws_by_seed_ = info.feature_weights.Empty() && info.num_col_ <= 8192;
      if (node_draws_by_seed_) {
        // The whole column sampler of the tree on the device: the tree-level set comes from the reference's sampler
        // (Init consumes the colsample_bytree draw exactly as it would); every later GetFeatureSet call -- a lazy level
        // draw the first time a depth is evaluated, a node draw per evaluated node -- takes the engine's next output as
        // its shuffle seed (random.cu:67).  The updater gets the stream of outputs and keeps the cursor itself.
        column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, 1.0f, 1.0f, param->colsample_bytree);
        auto p_tree = column_sampler_.GetFeatureSet(ctx_, 0);
        auto const& h_tree = p_tree->ConstHostVector();
        std::vector<std::int32_t> tree_set(h_tree.cbegin(), h_tree.cend());
        rng_after_init_ = ctx_->Rng();
        auto ahead = ctx_->Rng();  // a copy: the real engine moves in FinishSampling by the outputs really used
        std::int32_t const max_levels =
            param->max_depth > 0 ? param->max_depth + 1 : std::max<std::int32_t>(2, param->max_leaves);
        std::vector<std::uint32_t> seeds(static_cast<std::size_t>(max_nodes + max_levels));
        for (auto& sd : seeds) sd = static_cast<std::uint32_t>(ahead());
        B200_SHIM_CHECK(b200hist_updater_set_level_feature_sets(updater_, nullptr, 0, 0));
        B200_SHIM_CHECK(b200hist_updater_set_tree_sampling(updater_, nullptr, 0, rows ? param->subsample : 1.0f, seed));
        B200_SHIM_CHECK(b200hist_updater_set_column_sampler(updater_, tree_set.data(),
                                                            static_cast<std::uint32_t>(tree_set.size()),
                                                            param->colsample_bylevel, param->colsample_bynode,
                                                            seeds.data(), static_cast<std::uint32_t>(seeds.size())));
        node_draws_pending_ = true;
        sampling_set_ = true;
        return;
      } else {  // feature weights: weighted sampling without replacement, through the reference's own sampler
        column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, param->colsample_bynode, 1.0f,
                             param->colsample_bytree);
        for (std::int32_t k = 0; k < max_nodes; ++k) {
          auto p_set = column_sampler_.GetFeatureSet(ctx_, 0);
          auto const& h_set = p_set->ConstHostVector();
          if (k == 0) per_node = h_set.size();
          CHECK_EQ(h_set.size(), per_node);
          fs.insert(fs.end(), h_set.cbegin(), h_set.cend());
        }
      }
      B200_SHIM_CHECK(b200hist_updater_set_level_feature_sets(updater_, nullptr, 0, 0));
      B200_SHIM_CHECK(b200hist_updater_set_tree_sampling(updater_, nullptr, 0, rows ? param->subsample : 1.0f, seed));
      B200_SHIM_CHECK(b200hist_updater_set_node_feature_sets(updater_, fs.data(), static_cast<std::uint32_t>(max_nodes),
                                                             static_cast<std::uint32_t>(per_node)));
      node_draws_pending_ = true;
    } else if (levels) {
      // ColumnSampler draws a level's set lazily, when the first node of that depth is evaluated
      // (random.h:156-175).  Our tree is built without coming back to the host, so the sets of ALL possible depths
      // are drawn up front from a COPY of the engine state; FinishSampling() then replays exactly the draws the
      // reference would have made (the depths the tree really evaluated) on the real engine.
      rng_before_columns_ = ctx_->Rng();
      std::int32_t const n_levels =
          param->max_depth > 0 ? param->max_depth : std::max<std::int32_t>(1, param->max_leaves - 1);
      column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, 1.0f, param->colsample_bylevel,
                           param->colsample_bytree);
      std::size_t per_level = 0;
      for (std::int32_t d = 0; d < n_levels; ++d) {
        auto p_set = column_sampler_.GetFeatureSet(ctx_, d);
        auto const& h_set = p_set->ConstHostVector();
        if (d == 0) per_level = h_set.size();
        CHECK_EQ(h_set.size(), per_level);
        fs.insert(fs.end(), h_set.cbegin(), h_set.cend());
      }
      B200_SHIM_CHECK(b200hist_updater_set_tree_sampling(updater_, nullptr, 0, rows ? param->subsample : 1.0f, seed));
      B200_SHIM_CHECK(b200hist_updater_set_level_feature_sets(updater_, fs.data(), static_cast<std::uint32_t>(n_levels),
                                                              static_cast<std::uint32_t>(per_level)));
      level_draws_pending_ = true;
    } else {
      if (cols) {
        column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, 1.0f, 1.0f, param->colsample_bytree);
        auto p_set = column_sampler_.GetFeatureSet(ctx_, 0);
        auto const& h_set = p_set->ConstHostVector();
        fs.assign(h_set.cbegin(), h_set.cend());
      }
      B200_SHIM_CHECK(b200hist_updater_set_level_feature_sets(updater_, nullptr, 0, 0));
      B200_SHIM_CHECK(b200hist_updater_set_tree_sampling(updater_, cols ? fs.data() : nullptr,
                                                         static_cast<std::uint32_t>(fs.size()),
                                                         rows ? param->subsample : 1.0f, seed));
    }
    sampling_set_ = rows || cols || levels || by_node;
  }

  // colsample_bylevel: leave ctx->Rng() where the reference would have left it after this tree.
  void FinishSampling(TrainParam const* param, DMatrix* p_fmat) {
    if (node_draws_pending_) {  // colsample_bynode: as many draws as nodes were evaluated
      std::int32_t n_draws = 0;
      if (node_draws_by_seed_) {  // one engine output per ColSample call, level draws included
        B200_SHIM_CHECK(b200hist_updater_last_draw_count(updater_, &n_draws));
        ctx_->Rng() = rng_after_init_;
        ctx_->Rng().discard(static_cast<unsigned long long>(n_draws));
        node_draws_pending_ = false;
        return;
      }
      B200_SHIM_CHECK(b200hist_updater_last_eval_count(updater_, &n_draws));
      ctx_->Rng() = rng_before_columns_;
      auto const& info = p_fmat->Info();
      column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, param->colsample_bynode, 1.0f,
                           param->colsample_bytree);
      for (std::int32_t k = 0; k < n_draws; ++k) (void)column_sampler_.GetFeatureSet(ctx_, 0);
      node_draws_pending_ = false;
      return;
    }
    if (!level_draws_pending_) return;
    std::int32_t deepest = 0;
    B200_SHIM_CHECK(b200hist_updater_last_max_eval_depth(updater_, &deepest));
    ctx_->Rng() = rng_before_columns_;
    auto const& info = p_fmat->Info();
    column_sampler_.Init(ctx_, info.num_col_, info.feature_weights, 1.0f, param->colsample_bylevel,
                         param->colsample_bytree);
    for (std::int32_t d = 0; d <= deepest; ++d) (void)column_sampler_.GetFeatureSet(ctx_, d);
    level_draws_pending_ = false;
  }

  // ApplySplit (updater_gpu_hist.cu:610-655): our nodes are numbered in the reference's expansion order,
  // so replaying the internal nodes by ascending left-child id reproduces RegTree's AllocNode sequence.
  void ReplayInto(RegTree* p_tree, std::int32_t n_nodes) {
    auto& tree = *p_tree;
    auto const& root = nodes_[0];
    if (has_categorical_)
      B200_SHIM_CHECK(b200hist_updater_last_tree_categories(updater_, node_is_cat_.data(), cat_bits_.data()));
    tree.Stat(RegTree::kRoot).sum_hess = root.sum_hess;
    tree.Stat(RegTree::kRoot).base_weight = root.base_weight;
    tree[RegTree::kRoot].SetLeaf(root.leaf_value);
    std::vector<std::int32_t> internal;
    for (std::int32_t i = 0; i < n_nodes; ++i) {
      if (!nodes_[i].is_leaf) internal.push_back(i);
    }
    std::sort(internal.begin(), internal.end(),
              [&](std::int32_t a, std::int32_t b) { return nodes_[a].left < nodes_[b].left; });
    for (auto nidx : internal) {
      auto const& n = nodes_[nidx];
      auto const& l = nodes_[n.left];
      auto const& r = nodes_[n.right];
      // children are created as leaves holding learning_rate * weight; expanded later if they split
      float const lw = l.is_leaf ? l.leaf_value : LeafOf(n.left);
      float const rw = r.is_leaf ? r.leaf_value : LeafOf(n.right);
      if (has_categorical_ && node_is_cat_[nidx] != 0) {
        // updater_gpu_hist.cu:628-642: the set is stored in as many words as the feature's bins need
        auto const n_bins_feature = ellpack_cut_bins_.at(static_cast<std::size_t>(n.findex));
        std::size_t const words = std::min<std::size_t>(common::CatBitField::ComputeStorageSize(n_bins_feature), cat_words_);
        common::Span<std::uint32_t const> cats{cat_bits_.data() + static_cast<std::size_t>(nidx) * cat_words_, words};
        tree.ExpandCategorical(nidx, static_cast<bst_feature_t>(n.findex), cats, n.default_left != 0, n.base_weight, lw,
                               rw, n.loss_chg, n.sum_hess, l.sum_hess, r.sum_hess);
      } else {
        tree.ExpandNode(nidx, static_cast<bst_feature_t>(n.findex), n.fvalue, n.default_left != 0, n.base_weight, lw,
                        rw, n.loss_chg, n.sum_hess, l.sum_hess, r.sum_hess);
      }
      CHECK_EQ(tree[nidx].LeftChild(), n.left);
      CHECK_EQ(tree[nidx].RightChild(), n.right);
    }
  }
  // leaf value a node had before it was expanded = learning_rate * its weight (kept by the library)
  [[nodiscard]] float LeafOf(std::int32_t nidx) const { return nodes_[nidx].leaf_value; }

  void Release() {
    if (updater_ != nullptr && connected_) {
      // unmap the peers' caches; nobody frees a cache that a peer still maps: meet the other workers first (when the
      // communicator is still up -- after XGCommunicatorFinalize every worker has finished its last exchange, whose
      // closing barrier is the guarantee that no peer reads this cache any more)
      ctx_->CUDACtx()->Stream().Sync();
      b200hist_updater_clear_peers(updater_);
      if (collective::IsDistributed()) {
        std::int32_t one{1};
        auto rc = collective::Allreduce(ctx_, linalg::MakeVec(&one, 1), collective::Op::kSum);
        (void)rc;
      }
    }
    connected_ = false;
    if (updater_ != nullptr) b200hist_updater_destroy(updater_);
    if (ellpack_ != nullptr) b200hist_ellpack_destroy(ellpack_);
    updater_ = nullptr;
    ellpack_ = nullptr;
    p_last_fmat_ = nullptr;
    p_last_tree_ = nullptr;
    sampling_set_ = false;
  }

  ObjInfo const* task_{nullptr};
  HistMakerTrainParam hist_maker_param_;
  b200hist_ellpack_t ellpack_{nullptr};
  b200hist_updater_t updater_{nullptr};
  std::vector<b200hist_tree_node> nodes_;
  DMatrix* p_last_fmat_{nullptr};
  RegTree const* p_last_tree_{nullptr};
  std::vector<float> leaf_values_;
  common::ColumnSampler column_sampler_;
  bool sampling_set_{false};
  bool connected_{false};
  long long max_cached_hist_node_{0};  // 0 = not set by the user
  bool level_draws_pending_{false};
  bool node_draws_pending_{false};
  bool node_draws_by_seed_{false};
  bool has_categorical_{false};
  std::uint32_t cat_words_{0};
  std::vector<std::int32_t> node_is_cat_;
  std::vector<std::uint32_t> cat_bits_;
  std::vector<std::uint32_t> ellpack_cut_bins_;
  std::remove_reference_t<decltype(std::declval<Context const&>().Rng())> rng_before_columns_;
  std::remove_reference_t<decltype(std::declval<Context const&>().Rng())> rng_after_init_;
};

// Replaces the factory of the reference's own "grow_gpu_hist" entry (keeping the original reachable as
// "grow_gpu_hist_reference" for A/B runs in one process) and adds the distinct name "grow_b200_hist".
struct B200Registrar {
  B200Registrar() {
    auto* reg = ::dmlc::Registry<::xgboost::TreeUpdaterReg>::Get();
    auto make = [](Context const* ctx, ObjInfo const* task) -> TreeUpdater* { return new B200HistMaker(ctx, task); };
    if (auto const* orig = reg->Find("grow_gpu_hist")) {
      auto body = orig->body;  // copy before it is overwritten
      reg->__REGISTER__("grow_gpu_hist_reference").describe("The reference's own GPU hist updater.").set_body(body);
    }
    reg->__REGISTER__("grow_gpu_hist")
        .describe("Grow tree with the Blackwell-native hist updater (libb200hist.so).")
        .set_body(make);
    reg->__REGISTER__("grow_b200_hist")
        .describe("Grow tree with the Blackwell-native hist updater (libb200hist.so).")
        .set_body(make);
  }
};
static B200Registrar b200_registrar__;

}  // namespace xgboost::tree
