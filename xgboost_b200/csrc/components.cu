// Component entry points of the C ABI: each runs ONE stage of the hot path on caller-provided device
// buffers so the parity tests can compare it with the oracle in isolation (SURVEY.md 8(a) rows
// H2, H4, E1-E4, P2).  They launch exactly the kernels the updater launches; only the small work
// tables that the updater derives on the device are prepared on the host here.
#include <algorithm>
#include <cmath>

#include "common.cuh"

using namespace b200;

namespace {
template <typename T>
struct DevBuf {
  T* p{nullptr};
  explicit DevBuf(size_t n) : p(DevAlloc<T>(n)) {}
  ~DevBuf() { cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};
}  // namespace

extern "C" {

int b200hist_build_hist(b200hist_ellpack_t ell, const b200hist_gpair64* gq_dev, const uint32_t* ridx_dev,
                        const uint64_t* seg_begin_host, const uint64_t* seg_count_host, int32_t n_nodes,
                        b200hist_gpair64* hist_dev, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(ell && gq_dev && ridx_dev && seg_begin_host && seg_count_host && hist_dev, "null argument");
  B200_CHECK(n_nodes >= 0 && n_nodes <= kMaxLevelNodes, "n_nodes out of range");
  DeviceGuard guard(ell->device);
  auto s = static_cast<cudaStream_t>(stream);
  B200_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(b200hist_gpair64) * (size_t)n_nodes * ell->total_bins, s));
  if (n_nodes == 0) return 0;
  PrepareHistKernels(ell->device);
  auto groups = MakeFeatureGroups(*ell, ell->device);
  std::vector<BuildEntry> build(n_nodes);
  std::vector<uint32_t> tile_seg;
  for (int32_t k = 0; k < n_nodes; ++k) {
    B200_CHECK(seg_begin_host[k] + seg_count_host[k] <= (1ull << 32), "segment out of range");
    uint32_t tiles = (uint32_t)((seg_count_host[k] + kHistTile - 1) / kHistTile);
    build[k] = BuildEntry{k, 0, (uint32_t)seg_begin_host[k], (uint32_t)seg_count_host[k], (uint32_t)tile_seg.size(), (uint32_t)k};
    tile_seg.insert(tile_seg.end(), tiles, (uint32_t)k);
  }
  TreeState ts{};
  ts.n_build = n_nodes;
  ts.build_tiles = (uint32_t)tile_seg.size();
  DevBuf<BuildEntry> d_build(n_nodes);
  DevBuf<uint32_t> d_tile_seg(tile_seg.size());
  DevBuf<TreeState> d_ts(1);
  DevBuf<GroupDesc> d_groups(groups.size());
  B200_CUDA(cudaMemcpyAsync(d_build.p, build.data(), sizeof(BuildEntry) * n_nodes, cudaMemcpyHostToDevice, s));
  if (!tile_seg.empty())
    B200_CUDA(cudaMemcpyAsync(d_tile_seg.p, tile_seg.data(), sizeof(uint32_t) * tile_seg.size(), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_ts.p, &ts, sizeof(ts), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_groups.p, groups.data(), sizeof(GroupDesc) * groups.size(), cudaMemcpyHostToDevice, s));
  B200_CHECK(!ell->paged, "component entry points need an in-core page");
  LaunchBuildHist(ell->View(), gq_dev, ridx_dev, ridx_dev, d_build.p, d_tile_seg.p, d_ts.p, d_groups.p, groups,
                  hist_dev, nullptr, NumSMs(ell->device), s);
  B200_CUDA(cudaStreamSynchronize(s));
  B200_API_END();
}

int b200hist_subtract(int device, const b200hist_gpair64* parent_dev, const b200hist_gpair64* built_dev,
                      b200hist_gpair64* out_dev, uint64_t n_bins, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(parent_dev && built_dev && out_dev, "null argument");
  DeviceGuard guard(device);
  LaunchSubtractFlat(parent_dev, built_dev, out_dev, n_bins, static_cast<cudaStream_t>(stream));
  B200_API_END();
}

int b200hist_evaluate_splits(b200hist_ellpack_t ell, const b200hist_gpair64* hist_dev,
                             const b200hist_gpair64* parent_sums_host, int32_t n_nodes,
                             const int32_t* feature_set_host, int32_t n_feature_set,
                             const b200hist_train_param* param, const b200hist_quantiser* quantiser,
                             b200hist_candidate* out_host, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(ell && hist_dev && parent_sums_host && param && quantiser && out_host, "null argument");
  B200_CHECK(n_nodes >= 0 && n_nodes <= 2 * kMaxLevelNodes, "n_nodes out of range");
  if (n_nodes == 0) return 0;
  if (feature_set_host)
    for (int32_t i = 0; i < n_feature_set; ++i)
      B200_CHECK(feature_set_host[i] >= 0 && (uint32_t)feature_set_host[i] < ell->n_features, "feature id out of range");
  DeviceGuard guard(ell->device);
  auto s = static_cast<cudaStream_t>(stream);
  std::vector<NodeDev> nodes(n_nodes);
  std::vector<int32_t> eval_nodes(n_nodes);
  for (int32_t k = 0; k < n_nodes; ++k) {
    NodeDev nd{};
    nd.parent = -1; nd.left = -1; nd.right = -1; nd.is_leaf = 1; nd.findex = -1; nd.seg_buf = -1;
    nd.sum_g = parent_sums_host[k].grad;
    nd.sum_h = parent_sums_host[k].hess;
    nd.hist_slot = k;
    nodes[k] = nd;
    eval_nodes[k] = k;
  }
  TreeState ts{};
  ts.n_eval = n_nodes;
  QuantDev q{};
  q.to_fixed_g = quantiser->to_fixed_grad; q.to_fixed_h = quantiser->to_fixed_hess;
  q.to_float_g = quantiser->to_float_grad; q.to_float_h = quantiser->to_float_hess;
  const int32_t n_fs = feature_set_host ? n_feature_set : (int32_t)ell->n_features;
  DevBuf<NodeDev> d_nodes(n_nodes);
  DevBuf<int32_t> d_eval(n_nodes);
  DevBuf<TreeState> d_ts(1);
  DevBuf<QuantDev> d_q(1);
  DevBuf<int32_t> d_fs(std::max(1, n_fs));
  DevBuf<FeatCand> d_scratch((size_t)n_nodes * std::max(1, n_fs));
  DevBuf<b200hist_candidate> d_cand(n_nodes);
  B200_CUDA(cudaMemcpyAsync(d_nodes.p, nodes.data(), sizeof(NodeDev) * n_nodes, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_eval.p, eval_nodes.data(), sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_ts.p, &ts, sizeof(ts), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_q.p, &q, sizeof(q), cudaMemcpyHostToDevice, s));
  if (feature_set_host && n_fs > 0)
    B200_CUDA(cudaMemcpyAsync(d_fs.p, feature_set_host, sizeof(int32_t) * n_fs, cudaMemcpyHostToDevice, s));
  // a node with an empty feature set keeps the default candidate
  std::vector<b200hist_candidate> init(n_nodes);
  for (auto& c : init) {
    c = b200hist_candidate{};
    c.loss_chg = -3.402823466e+38f;
    c.findex = -1;
    c.split_gbin = -1;
    c.base_weight = c.left_weight = c.right_weight = NAN;
  }
  B200_CUDA(cudaMemcpyAsync(d_cand.p, init.data(), sizeof(b200hist_candidate) * n_nodes, cudaMemcpyHostToDevice, s));
  if (n_fs > 0) {
    LaunchEvaluate(ell->View(), hist_dev, d_eval.p, d_ts.p, n_nodes, d_nodes.p, feature_set_host ? d_fs.p : nullptr,
                   n_fs, ToParamDev(*param), d_q.p, d_scratch.p, d_cand.p, d_nodes.p, NumSMs(ell->device), s);
  }
  B200_CUDA(cudaMemcpyAsync(out_host, d_cand.p, sizeof(b200hist_candidate) * n_nodes, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  B200_API_END();
}

int b200hist_evaluate_splits_cat(b200hist_ellpack_t ell, const b200hist_gpair64* hist_dev,
                                 const b200hist_gpair64* parent_sums_host, int32_t n_nodes,
                                 const int32_t* feature_set_host, int32_t n_feature_set,
                                 const b200hist_train_param* param, const b200hist_quantiser* quantiser,
                                 const uint8_t* feature_is_cat_host, int32_t max_cat_to_onehot,
                                 int32_t max_cat_threshold, b200hist_candidate* out_host, int32_t* out_is_cat_host,
                                 uint32_t* out_cat_bits_host, uint32_t n_cat_words, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(ell && hist_dev && parent_sums_host && param && quantiser && out_host && feature_is_cat_host &&
                 out_is_cat_host && out_cat_bits_host,
             "null argument");
  B200_CHECK(n_nodes >= 0 && n_nodes <= 2 * kMaxLevelNodes, "n_nodes out of range");
  B200_CHECK(n_cat_words >= 1, "n_cat_words must be at least 1");
  if (n_nodes == 0) return 0;
  if (feature_set_host)
    for (int32_t i = 0; i < n_feature_set; ++i)
      B200_CHECK(feature_set_host[i] >= 0 && (uint32_t)feature_set_host[i] < ell->n_features, "feature id out of range");
  int32_t max_cat_bins = 1;
  for (uint32_t f = 0; f < ell->n_features; ++f)
    if (feature_is_cat_host[f])
      max_cat_bins = std::max<int32_t>(max_cat_bins, (int32_t)(ell->cut_ptrs[f + 1] - ell->cut_ptrs[f]));
  B200_CHECK(max_cat_bins <= kMaxCatBins, "a categorical feature has more bins than the sort scratch holds");
  DeviceGuard guard(ell->device);
  auto s = static_cast<cudaStream_t>(stream);
  std::vector<NodeDev> nodes(n_nodes);
  std::vector<int32_t> eval_nodes(n_nodes);
  for (int32_t k = 0; k < n_nodes; ++k) {
    NodeDev nd{};
    nd.parent = -1; nd.left = -1; nd.right = -1; nd.is_leaf = 1; nd.findex = -1; nd.seg_buf = -1;
    nd.sum_g = parent_sums_host[k].grad;
    nd.sum_h = parent_sums_host[k].hess;
    nd.hist_slot = k;
    nodes[k] = nd;
    eval_nodes[k] = k;
  }
  TreeState ts{};
  ts.n_eval = n_nodes;
  QuantDev q{};
  q.to_fixed_g = quantiser->to_fixed_grad; q.to_fixed_h = quantiser->to_fixed_hess;
  q.to_float_g = quantiser->to_float_grad; q.to_float_h = quantiser->to_float_hess;
  const int32_t n_fs = feature_set_host ? n_feature_set : (int32_t)ell->n_features;
  const int32_t bin_words = (max_cat_bins + 31) / 32;
  DevBuf<NodeDev> d_nodes(n_nodes);
  DevBuf<int32_t> d_eval(n_nodes);
  DevBuf<TreeState> d_ts(1);
  DevBuf<QuantDev> d_q(1);
  DevBuf<int32_t> d_fs(std::max(1, n_fs));
  DevBuf<FeatCand> d_scratch((size_t)n_nodes * std::max(1, n_fs));
  DevBuf<b200hist_candidate> d_cand(n_nodes);
  DevBuf<uint8_t> d_is_cat(ell->n_features);
  DevBuf<uint32_t> d_cat_bits((size_t)n_nodes * n_cat_words);
  DevBuf<uint32_t> d_bin_bits((size_t)n_nodes * bin_words);
  DevBuf<uint8_t> d_node_is_cat(n_nodes);
  B200_CUDA(cudaMemcpyAsync(d_nodes.p, nodes.data(), sizeof(NodeDev) * n_nodes, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_eval.p, eval_nodes.data(), sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_ts.p, &ts, sizeof(ts), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_q.p, &q, sizeof(q), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_is_cat.p, feature_is_cat_host, ell->n_features, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemsetAsync(d_cat_bits.p, 0, sizeof(uint32_t) * (size_t)n_nodes * n_cat_words, s));
  B200_CUDA(cudaMemsetAsync(d_node_is_cat.p, 0, (size_t)n_nodes, s));
  if (feature_set_host && n_fs > 0)
    B200_CUDA(cudaMemcpyAsync(d_fs.p, feature_set_host, sizeof(int32_t) * n_fs, cudaMemcpyHostToDevice, s));
  std::vector<b200hist_candidate> init(n_nodes);
  for (auto& c : init) {
    c = b200hist_candidate{};
    c.loss_chg = -3.402823466e+38f;
    c.findex = -1;
    c.split_gbin = -1;
    c.base_weight = c.left_weight = c.right_weight = NAN;
  }
  B200_CUDA(cudaMemcpyAsync(d_cand.p, init.data(), sizeof(b200hist_candidate) * n_nodes, cudaMemcpyHostToDevice, s));
  if (n_fs > 0) {
    const CatView cat{d_is_cat.p, max_cat_to_onehot, max_cat_threshold, d_cat_bits.p, d_bin_bits.p, d_node_is_cat.p,
                      (int32_t)n_cat_words, bin_words, max_cat_bins};
    LaunchEvaluate(ell->View(), hist_dev, d_eval.p, d_ts.p, n_nodes, d_nodes.p, feature_set_host ? d_fs.p : nullptr,
                   n_fs, ToParamDev(*param), d_q.p, d_scratch.p, d_cand.p, d_nodes.p, NumSMs(ell->device), s,
                   MonoView{nullptr, nullptr, nullptr}, 0, nullptr, IcView{nullptr, nullptr, 0}, cat);
  }
  std::vector<uint8_t> node_is_cat(n_nodes);
  B200_CUDA(cudaMemcpyAsync(out_host, d_cand.p, sizeof(b200hist_candidate) * n_nodes, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(node_is_cat.data(), d_node_is_cat.p, (size_t)n_nodes, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(out_cat_bits_host, d_cat_bits.p, sizeof(uint32_t) * (size_t)n_nodes * n_cat_words,
                            cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  for (int32_t k = 0; k < n_nodes; ++k) out_is_cat_host[k] = node_is_cat[k];
  B200_API_END();
}

int b200hist_partition(b200hist_ellpack_t ell, const uint32_t* ridx_in_dev, uint32_t* ridx_out_dev,
                       const uint64_t* seg_begin_host, const uint64_t* seg_count_host, const int32_t* findex_host,
                       const float* split_cond_host, const int32_t* default_left_host, int32_t n_nodes,
                       uint64_t* left_count_host, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(ell && ridx_in_dev && ridx_out_dev && seg_begin_host && seg_count_host && findex_host &&
                 split_cond_host && default_left_host,
             "null argument");
  B200_CHECK(ridx_in_dev != ridx_out_dev, "partition is out of place");
  B200_CHECK(n_nodes >= 0 && n_nodes <= kMaxLevelNodes, "n_nodes out of range");
  if (n_nodes == 0) return 0;
  DeviceGuard guard(ell->device);
  auto s = static_cast<cudaStream_t>(stream);
  std::vector<PartEntry> part(n_nodes);
  std::vector<uint32_t> tile_seg;
  for (int32_t k = 0; k < n_nodes; ++k) {
    B200_CHECK(findex_host[k] >= 0 && (uint32_t)findex_host[k] < ell->n_features, "feature id out of range");
    const uint32_t beg = ell->cut_ptrs[findex_host[k]], end = ell->cut_ptrs[findex_host[k] + 1];
    // cut_values[bin] <= split_cond  <=>  local bin < number of cuts <= split_cond
    uint32_t lbin = (uint32_t)(std::upper_bound(ell->cut_values.begin() + beg, ell->cut_values.begin() + end,
                                                split_cond_host[k]) -
                               (ell->cut_values.begin() + beg));
    PartEntry e{};
    e.nidx = k; e.left = -1; e.right = -1; e.findex = findex_host[k]; e.split_lbin = (int32_t)lbin;
    e.default_left = default_left_host[k] != 0; e.begin = (uint32_t)seg_begin_host[k];
    e.count = (uint32_t)seg_count_host[k]; e.tile_begin = (uint32_t)tile_seg.size(); e.src_buf = 0;
    uint32_t tiles = (e.count + kPartTile - 1) / kPartTile;
    tile_seg.insert(tile_seg.end(), tiles, (uint32_t)k);
    part[k] = e;
  }
  const uint32_t T = (uint32_t)tile_seg.size();
  TreeState ts{};
  ts.n_part = n_nodes;
  ts.part_tiles = T;
  DevBuf<PartEntry> d_part(n_nodes);
  DevBuf<uint32_t> d_tile_seg(T), d_tile_left(T), d_tile_off(T), d_flags((size_t)T * (kPartTile / 32));
  DevBuf<TreeState> d_ts(1);
  B200_CUDA(cudaMemcpyAsync(d_part.p, part.data(), sizeof(PartEntry) * n_nodes, cudaMemcpyHostToDevice, s));
  if (T) B200_CUDA(cudaMemcpyAsync(d_tile_seg.p, tile_seg.data(), sizeof(uint32_t) * T, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_ts.p, &ts, sizeof(ts), cudaMemcpyHostToDevice, s));
  const int sms = NumSMs(ell->device);
  B200_CHECK(!ell->paged, "component entry points need an in-core page");
  LaunchPartitionCount(ell->View(), ridx_in_dev, ridx_in_dev, d_part.p, d_tile_seg.p, d_ts.p, d_flags.p, d_tile_left.p, sms, s);
  std::vector<uint32_t> tile_left(T), tile_off(T);
  if (T) B200_CUDA(cudaMemcpyAsync(tile_left.data(), d_tile_left.p, sizeof(uint32_t) * T, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  for (int32_t k = 0; k < n_nodes; ++k) {
    uint32_t tb = part[k].tile_begin, te = (k + 1 < n_nodes) ? part[k + 1].tile_begin : T;
    uint32_t acc = 0;
    for (uint32_t t = tb; t < te; ++t) {
      tile_off[t] = acc;
      acc += tile_left[t];
    }
    part[k].n_left = acc;
    if (left_count_host) left_count_host[k] = acc;
  }
  B200_CUDA(cudaMemcpyAsync(d_part.p, part.data(), sizeof(PartEntry) * n_nodes, cudaMemcpyHostToDevice, s));
  if (T) B200_CUDA(cudaMemcpyAsync(d_tile_off.p, tile_off.data(), sizeof(uint32_t) * T, cudaMemcpyHostToDevice, s));
  // src_buf = 0 for every entry: reads ridx_in (buffer 0), writes the "other" buffer = ridx_out
  LaunchPartitionScatter(ridx_in_dev, ridx_in_dev, ridx_out_dev, ridx_out_dev, d_part.p, d_tile_seg.p, d_ts.p,
                         d_flags.p, d_tile_off.p, sms, s);
  B200_CUDA(cudaStreamSynchronize(s));
  B200_API_END();
}

}  // extern "C"
