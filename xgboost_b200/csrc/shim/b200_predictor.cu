// Reference-side binding 2: xgboost::Predictor served by libb200hist.so for quantised (Ellpack-only) matrices.
//
// Compiled against the reference's headers, next to b200_tree_updater.cu, into the same shim library.  It subclasses
// xgboost::Predictor (/root/reference/include/xgboost/predictor.h:63-178), keeps the reference's own "gpu_predictor"
// alive inside itself and forwards EVERYTHING to it except the one call the hot path owns:
//
//   PredictBatch and PredictLeaf on a DMatrix that has no SparsePage, i.e. a QuantileDMatrix
//   (src/predictor/gpu_predictor.cu:565-589 routes exactly those through EllpackLoader), scalar leaves (one output
//   per row, or one per class with tree_info naming each tree's output), no tree weights, no category re-coding, no
//   column split.
//
// The page is read in place (b200hist_forest_predict_reference_page decodes EllpackPageImpl::gidx_buffer itself):
// nothing is copied, and a call costs what the reference's own walk costs in memory traffic.
//
// That is the call GBTree::PredictBatchImpl makes for a validation set every boosting iteration
// (src/gbm/gbtree.cc:514-560: tree_begin = the cache's version, so only the NEW trees are walked).  Anything else --
// and any tree b200hist_forest_add_tree refuses -- goes to the reference's predictor before an output is touched.
//
// STATUS: on since round 2.  libb200hist's forest kernels equal the oracle under the CPU emulator
// (tests/test_kernel_emulation.py), the oracle equals the reference's quantised prediction bit for bit
// (tests/test_reference_golden.py), and on a B200 this binding returns the reference predictor's margins and leaf ids
// bit for bit (tests/test_gpu_next.py::test_predictor_binding_serves_quantised_matrices_like_the_reference: dense,
// missing values, uint16 symbols, multi-class; timing: tools/bench_predictor.py, profiles/).  B200HIST_SHIM_PREDICTOR=0
// when the shim is loaded leaves the reference's own factory in place.
#include <dmlc/registry.h>
#include <xgboost/base.h>
#include <xgboost/context.h>
#include <xgboost/data.h>
#include <xgboost/host_device_vector.h>
#include <xgboost/learner.h>
#include <xgboost/logging.h>
#include <xgboost/predictor.h>
#include <xgboost/tree_model.h>

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "../../../include/b200hist.h"
#include "src/common/cuda_context.cuh"  // CUDAContext::Stream
#include "src/data/batch_utils.h"       // cuda_impl::StaticBatch
#include "src/data/cat_container.h"     // CatContainer::NeedRecode
#include "src/data/ellpack_page.cuh"    // EllpackPageImpl
#include "src/gbm/gbtree_model.h"       // GBTreeModel

namespace xgboost::predictor {

class B200Predictor : public Predictor {
 public:
  B200Predictor(Context const* ctx, std::unique_ptr<Predictor> reference)
      : Predictor{ctx}, ref_{std::move(reference)} {}
  ~B200Predictor() override { this->ReleaseScratch(); }

  void Configure(Args const& args) override { ref_->Configure(args); }
  void InitOutPredictions(MetaInfo const& info, HostDeviceVector<float>* out_predt,
                          gbm::GBTreeModel const& model) const override {
    ref_->InitOutPredictions(info, out_predt, model);
  }

  void PredictBatch(DMatrix* dmat, PredictionCacheEntry* out_preds, gbm::GBTreeModel const& model,
                    bst_tree_t tree_begin, bst_tree_t tree_end = 0,
                    std::vector<float> const* tree_weights = nullptr) const override {
    bst_tree_t const end = tree_end == 0 ? static_cast<bst_tree_t>(model.trees.size()) : tree_end;
    if (this->Eligible(dmat, model, tree_begin, end, tree_weights) &&
        out_preds->predictions.Size() == dmat->Info().num_row_ * model.learner_model_param->OutputLength() &&
        this->PredictQuantised(dmat, &out_preds->predictions, model, tree_begin, end, /*leaf=*/false)) {
      return;
    }
    ref_->PredictBatch(dmat, out_preds, model, tree_begin, tree_end, tree_weights);
  }

  bool InplacePredict(std::shared_ptr<DMatrix> p_fmat, gbm::GBTreeModel const& model, float missing,
                      PredictionCacheEntry* out_preds, bst_tree_t tree_begin = 0, bst_tree_t tree_end = 0,
                      std::vector<float> const* tree_weights = nullptr) const override {
    return ref_->InplacePredict(std::move(p_fmat), model, missing, out_preds, tree_begin, tree_end, tree_weights);
  }
  void PredictLeaf(DMatrix* dmat, HostDeviceVector<float>* out_preds, gbm::GBTreeModel const& model,
                   bst_tree_t tree_end = 0) const override {
    // GPUPredictor::PredictLeaf (gpu_predictor.cu:793-830): trees [0, GetTreeLimit), ids as floats, row-major
    bst_tree_t const end = tree_end == 0 || tree_end > static_cast<bst_tree_t>(model.trees.size())
                               ? static_cast<bst_tree_t>(model.trees.size())
                               : tree_end;
    if (this->Eligible(dmat, model, 0, end, nullptr)) {
      out_preds->SetDevice(ctx_->Device());
      out_preds->Resize(dmat->Info().num_row_ * static_cast<std::size_t>(end));
      if (this->PredictQuantised(dmat, out_preds, model, 0, end, /*leaf=*/true)) return;
    }
    ref_->PredictLeaf(dmat, out_preds, model, tree_end);
  }
  void PredictContribution(DMatrix* dmat, HostDeviceVector<float>* out_contribs, gbm::GBTreeModel const& model,
                           bst_tree_t tree_end = 0, std::vector<float> const* tree_weights = nullptr,
                           bool approximate = false, int condition = 0,
                           unsigned condition_feature = 0) const override {
    ref_->PredictContribution(dmat, out_contribs, model, tree_end, tree_weights, approximate, condition,
                              condition_feature);
  }
  void PredictInteractionContributions(DMatrix* dmat, HostDeviceVector<float>* out_contribs,
                                       gbm::GBTreeModel const& model, bst_tree_t tree_end = 0,
                                       std::vector<float> const* tree_weights = nullptr,
                                       bool approximate = false) const override {
    ref_->PredictInteractionContributions(dmat, out_contribs, model, tree_end, tree_weights, approximate);
  }

 private:
  // The conditions under which GPUPredictor::PredictDMatrix (gpu_predictor.cu:627-660) reaches EllpackLoader with
  // the no-op category accessor and unit tree weights -- the only configuration the forest kernels restate.
  bool Eligible(DMatrix* dmat, gbm::GBTreeModel const& model, bst_tree_t begin, bst_tree_t end,
                std::vector<float> const* tree_weights) const {
    if (!ctx_->IsCUDA() || tree_weights != nullptr || end <= begin) return false;
    if (end > static_cast<bst_tree_t>(model.trees.size())) return false;
    auto const& info = dmat->Info();
    if (info.IsColumnSplit() || model.learner_model_param->IsVectorLeaf()) return false;
    if (model.learner_model_param->OutputLength() > 1 &&
        model.tree_info.Size() < static_cast<std::size_t>(end)) return false;
    if (dmat->PageExists<SparsePage>() || !dmat->PageExists<EllpackPage>() || !dmat->SingleColBlock()) return false;
    if (dmat->Cats() != nullptr && dmat->Cats()->NeedRecode()) return false;
    if (info.num_col_ > model.learner_model_param->num_feature) return false;  // the reference raises its own error
    for (bst_tree_t t = begin; t < end; ++t) {
      if (model.trees[t]->IsMultiTarget()) return false;
    }
    return true;
  }

  // The forest resolves split conditions against a page object of ours that carries the cuts; a one-row page is all
  // it needs (the rows are read straight from the reference's buffer).  Rebuilt when the cuts change.
  bool FitScratch(EllpackPageImpl const* impl) const {
    auto cuts = impl->CutsShared();
    auto const& ptrs = cuts->cut_ptrs_.ConstHostVector();
    auto const& vals = cuts->cut_values_.ConstHostVector();
    if (scratch_ != nullptr && scratch_ptrs_ == ptrs && scratch_vals_ == vals) return true;
    this->ReleaseScratch();
    if (b200hist_ellpack_create(ctx_->Ordinal(), 1, static_cast<std::uint32_t>(ptrs.size() - 1), ptrs.data(),
                                vals.data(), /*has_missing=*/1, &scratch_) != 0) {
      scratch_ = nullptr;
      return false;
    }
    scratch_ptrs_ = ptrs;
    scratch_vals_ = vals;
    return true;
  }

  void ReleaseScratch() const {
    if (scratch_ != nullptr) b200hist_ellpack_destroy(scratch_);
    scratch_ = nullptr;
  }

  // Trees [begin, end) of the model as the C ABI's node arrays, resolved against the scratch page's cuts.
  // RegTree::GetNodes / Node accessors: include/xgboost/tree_model.h:96-190,256-261; categorical storage: :491-520 (one LBitField32
  // segment per node, the layout b200hist_forest_add_tree takes).
  bool BuildForest(gbm::GBTreeModel const& model, bst_tree_t begin, bst_tree_t end, b200hist_forest_t* out) const {
    b200hist_forest_t forest{nullptr};
    if (b200hist_forest_create(scratch_, &forest) != 0) return false;
    std::vector<b200hist_tree_node> nodes;
    std::vector<std::int32_t> is_cat;
    std::vector<std::uint32_t> bits;
    for (bst_tree_t t = begin; t < end; ++t) {
      RegTree const& tree = *model.trees[t];
      auto const n = static_cast<std::size_t>(tree.NumNodes());
      auto types = tree.GetSplitTypes(DeviceOrd::CPU());
      auto cats = tree.GetSplitCategories(DeviceOrd::CPU());
      auto tree_nodes = tree.GetNodes(DeviceOrd::CPU());
      auto const& segments = tree.GetSplitCategoriesPtr();
      std::size_t words{0};
      for (std::size_t nid = 0; nid < n; ++nid) {
        auto const& node = tree_nodes[nid];
        if (!node.IsDeleted() && !node.IsLeaf() && nid < types.size() && types[nid] == FeatureType::kCategorical &&
            nid < segments.size()) {
          words = std::max(words, segments[nid].size);
        }
      }
      nodes.assign(n, b200hist_tree_node{});
      is_cat.assign(n, 0);
      bits.assign(n * std::max<std::size_t>(words, 1), 0u);
      for (std::size_t nid = 0; nid < n; ++nid) {
        auto const& node = tree_nodes[nid];
        auto& d = nodes[nid];
        d.parent = node.IsDeleted() || node.IsRoot() ? -1 : static_cast<std::int32_t>(node.Parent());
        if (node.IsDeleted() || node.IsLeaf()) {  // a deleted node is unreachable
          d.left = d.right = -1;
          d.findex = -1;
          d.is_leaf = 1;
          d.leaf_value = node.IsDeleted() ? 0.0f : node.LeafValue();
          continue;
        }
        d.left = node.LeftChild();
        d.right = node.RightChild();
        d.findex = static_cast<std::int32_t>(node.SplitIndex());
        d.fvalue = node.SplitCond();
        d.default_left = node.DefaultLeft() ? 1 : 0;
        if (words > 0 && nid < types.size() && types[nid] == FeatureType::kCategorical && nid < segments.size()) {
          is_cat[nid] = 1;
          auto const& seg = segments[nid];
          for (std::size_t w = 0; w < seg.size && seg.beg + w < cats.size(); ++w) {
            bits[nid * words + w] = cats[seg.beg + w];
          }
        }
      }
      bool const any_cat = words > 0;
      if (b200hist_forest_add_tree(forest, nodes.data(), static_cast<std::int32_t>(n),
                                   any_cat ? is_cat.data() : nullptr, any_cat ? bits.data() : nullptr,
                                   static_cast<std::uint32_t>(words)) != 0) {
        b200hist_forest_destroy(forest);
        return false;  // e.g. a pruned tree whose node ids are no longer in expansion order
      }
    }
    auto const n_groups = model.learner_model_param->OutputLength();
    if (n_groups > 1) {  // gbm::GBTreeModel::tree_info: the output every tree feeds
      auto const& info = model.tree_info.ConstHostVector();
      std::vector<std::uint32_t> groups(info.cbegin() + begin, info.cbegin() + end);
      if (b200hist_forest_set_tree_groups(forest, groups.data(), static_cast<std::uint32_t>(groups.size())) != 0) {
        b200hist_forest_destroy(forest);
        return false;
      }
    }
    *out = forest;
    return true;
  }

  // leaf == false: margins are ADDED to out_preds (n_rows x n_groups); leaf == true: out_preds (n_rows x n_trees) is
  // overwritten with leaf ids.
  bool PredictQuantised(DMatrix* dmat, HostDeviceVector<float>* out_preds, gbm::GBTreeModel const& model,
                        bst_tree_t begin, bst_tree_t end, bool leaf) const {
    std::lock_guard<std::mutex> guard{mu_};
    out_preds->SetDevice(ctx_->Device());
    auto* stream = static_cast<void*>(ctx_->CUDACtx()->Stream());
    b200hist_forest_t forest{nullptr};
    bool wrote{false}, ok{true};
    bst_idx_t offset{0};
    for (auto const& page : dmat->GetBatches<EllpackPage>(ctx_, cuda_impl::StaticBatch(true))) {
      auto const* impl = page.Impl();
      ok = this->FitScratch(impl) && this->BuildForest(model, begin, end, &forest);
      if (!ok) break;
      auto const n_trees = static_cast<std::uint32_t>(end - begin);
      auto const n_rows = static_cast<std::uint64_t>(impl->n_rows);
      auto const n_groups = static_cast<std::uint32_t>(model.learner_model_param->OutputLength());
      // EllpackPageImpl: gidx_buffer, info.row_stride, NumSymbols, IsDenseCompressed, IsDense
      // (src/data/ellpack_page.cuh:181-391) -- read in place, nothing is copied
      int const mode = leaf ? 2 : (n_groups == 1 ? 0 : 1);
      std::uint64_t const width = leaf ? n_trees : n_groups;
      ok = b200hist_forest_predict_reference_page(
               forest, impl->gidx_buffer.data(), n_rows, static_cast<std::uint32_t>(impl->info.row_stride),
               static_cast<std::uint32_t>(impl->NumSymbols()), impl->IsDenseCompressed() ? 1 : 0,
               impl->IsDense() ? 1 : 0, mode, 0, n_trees, n_groups, out_preds->DevicePointer() + offset * width,
               stream) == 0;
      ctx_->CUDACtx()->Stream().Sync();  // the forest is freed below; the reference may recycle the page
      b200hist_forest_destroy(forest);
      forest = nullptr;
      if (!ok) break;
      wrote = true;
      offset += impl->n_rows;
    }
    if (!ok && wrote) LOG(FATAL) << "[b200hist] prediction failed half way: " << b200hist_last_error();
    return ok && wrote;
  }

  std::unique_ptr<Predictor> ref_;
  mutable std::mutex mu_;
  mutable b200hist_ellpack_t scratch_{nullptr};
  mutable std::vector<std::uint32_t> scratch_ptrs_;
  mutable std::vector<float> scratch_vals_;
};

// Replaces the factory of "gpu_predictor" (the name Learner/GBTree resolve on CUDA, src/gbm/gbtree.cc:566-590) --
// unless B200HIST_SHIM_PREDICTOR=0, see STATUS above -- and keeps the original reachable as "gpu_predictor_reference".
struct B200PredictorRegistrar {
  B200PredictorRegistrar() {
    char const* e = std::getenv("B200HIST_SHIM_PREDICTOR");
    if (e != nullptr && e[0] == '0') return;
    auto* reg = ::dmlc::Registry<::xgboost::PredictorReg>::Get();
    auto const* orig = reg->Find("gpu_predictor");
    if (orig == nullptr) return;
    auto body = orig->body;  // copy before it is overwritten
    reg->__REGISTER__("gpu_predictor_reference").describe("The reference's own GPU predictor.").set_body(body);
    reg->__REGISTER__("gpu_predictor")
        .describe("GPU predictor; quantised matrices are served by libb200hist.so.")
        .set_body([body](Context const* ctx) -> Predictor* {
          return new B200Predictor(ctx, std::unique_ptr<Predictor>(body(ctx)));
        });
  }
};
static B200PredictorRegistrar b200_predictor_registrar__;

}  // namespace xgboost::predictor
