// Split evaluation under the CPU emulator: the kernel sources of xgboost_b200/csrc/evaluate.cu in the order
// LaunchEvaluate issues them -- FeatureScanKernel<kIc>, CatFeatureScanKernel, NodeReduceKernel<kIc>, CatFinishKernel --
// on histograms written by tests/test_kernel_emulation.py.  Node id = evaluation slot = histogram slot.
#define B200_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "emu_io.h"

#include "../../xgboost_b200/csrc/evaluate.cu"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  // meta: n_features, total_bins, n_eval, n_fs, has_fs, n_levels, ic_mode (0 none, 1 shared list, 2 per-slot lists),
  //       ic_stride, has_mono, has_cat, max_cat_to_onehot, max_cat_threshold, cat_words, bin_words, max_cat_bins
  auto meta = ReadAll<int32_t>(dir + "/meta.bin");
  auto prm = ReadAll<float>(dir + "/param_f.bin");   // learning_rate, lambda, alpha, min_child_weight, max_delta_step, gamma
  auto prm_i = ReadAll<int32_t>(dir + "/param_i.bin");  // max_depth, max_leaves, grow_policy
  auto quant = ReadAll<double>(dir + "/quant.bin");
  auto cut_ptrs = ReadAll<uint32_t>(dir + "/cut_ptrs.bin");
  auto cut_values = ReadAll<float>(dir + "/cut_values.bin");
  auto hist = ReadAll<longlong2>(dir + "/hist.bin");
  auto parents = ReadAll<long long>(dir + "/parents.bin");
  auto depth = ReadAll<int32_t>(dir + "/depth.bin");
  auto fs = ReadAll<int32_t>(dir + "/feature_set.bin", true);
  auto ic_list = ReadAll<int32_t>(dir + "/ic_list.bin", true);
  auto ic_count = ReadAll<int32_t>(dir + "/ic_count.bin", true);
  auto mono_c = ReadAll<int32_t>(dir + "/mono.bin", true);
  auto lower = ReadAll<float>(dir + "/lower.bin", true);
  auto upper = ReadAll<float>(dir + "/upper.bin", true);
  auto is_cat = ReadAll<uint8_t>(dir + "/is_cat.bin", true);
  const int n_features = meta[0], total_bins = meta[1], n_eval = meta[2], n_fs = meta[3];
  b200::EllView ell{};
  ell.cut_ptrs = cut_ptrs.data();
  ell.cut_values = cut_values.data();
  ell.n_features = (uint32_t)n_features;
  ell.total_bins = (uint32_t)total_bins;
  b200::ParamDev p{prm_i[0], prm_i[1], prm_i[2], prm[0], prm[1], prm[2], prm[3], prm[4], prm[5]};
  b200::QuantDev q{};
  q.to_fixed_g = quant[0]; q.to_fixed_h = quant[1]; q.to_float_g = quant[2]; q.to_float_h = quant[3];
  std::vector<b200::NodeDev> nodes((size_t)n_eval);
  std::vector<int32_t> eval_nodes((size_t)n_eval);
  for (int k = 0; k < n_eval; ++k) {
    b200::NodeDev nd{};
    nd.parent = -1; nd.left = -1; nd.right = -1; nd.is_leaf = 1; nd.findex = -1; nd.seg_buf = -1;
    nd.depth = depth[k];
    nd.sum_g = parents[2 * k]; nd.sum_h = parents[2 * k + 1];
    nd.hist_slot = k;
    nodes[k] = nd;
    eval_nodes[k] = k;
  }
  b200::TreeState ts{};
  ts.n_eval = n_eval;
  std::vector<b200::FeatCand> scratch((size_t)n_eval * n_fs);
  std::vector<b200hist_candidate> cand((size_t)n_eval);
  for (auto& c : cand) { c = b200hist_candidate{}; c.loss_chg = -3.402823466e+38f; c.findex = -1; c.split_gbin = -1; }
  const int32_t* d_fs = meta[4] ? fs.data() : nullptr;
  const b200::MonoView mono{meta[8] ? mono_c.data() : nullptr, lower.data(), upper.data()};
  const b200::IcView ic{meta[6] ? ic_list.data() : nullptr, meta[6] ? ic_count.data() : nullptr, meta[7]};
  std::vector<uint32_t> cat_bits((size_t)n_eval * std::max(1, meta[12])), bin_bits((size_t)n_eval * std::max(1, meta[13]));
  std::vector<uint8_t> node_is_cat((size_t)n_eval, 0);
  const b200::CatView cat{meta[9] ? is_cat.data() : nullptr, meta[10], meta[11], cat_bits.data(), bin_bits.data(),
                          node_is_cat.data(), meta[12], meta[13], std::max(1, meta[14])};
  const long long items = (long long)n_eval * n_fs;
  const unsigned blocks = (unsigned)std::min<long long>((items + b200::kEvalWarps - 1) / b200::kEvalWarps, 6);
  const unsigned blocks2 = (unsigned)std::min<long long>(((long long)n_eval + b200::kEvalWarps - 1) / b200::kEvalWarps, 3);
  const size_t cat_smem = (size_t)b200::kEvalWarps * (size_t)cat.max_cat_bins * (sizeof(float) + sizeof(uint16_t));
  int32_t max_depth_seen = 0;
  if (meta[6]) {
    EMU_LAUNCH((blocks), (b200::kEvalWarps * 32), 0,
               b200::FeatureScanKernel<true>(ell, hist.data(), eval_nodes.data(), &ts, nodes.data(), d_fs, n_fs, p, &q,
                                             scratch.data(), mono, meta[5], ic, b200::EvalShard{0, 1, (uint32_t)ell.n_features}));
  } else {
    EMU_LAUNCH((blocks), (b200::kEvalWarps * 32), 0,
               b200::FeatureScanKernel<false>(ell, hist.data(), eval_nodes.data(), &ts, nodes.data(), d_fs, n_fs, p, &q,
                                              scratch.data(), mono, meta[5], ic, b200::EvalShard{0, 1, (uint32_t)ell.n_features}));
  }
  if (meta[9])
    EMU_LAUNCH((blocks), (b200::kEvalWarps * 32), cat_smem,
               b200::CatFeatureScanKernel(ell, hist.data(), eval_nodes.data(), &ts, nodes.data(), d_fs, n_fs, p, &q,
                                          scratch.data(), mono, meta[5], ic, cat));
  if (meta[6]) {
    EMU_LAUNCH((blocks2), (b200::kEvalWarps * 32), 0,
               b200::NodeReduceKernel<true>(ell, eval_nodes.data(), &ts, nodes.data(), d_fs, n_fs, p, &q, scratch.data(),
                                            cand.data(), mono, meta[5], &max_depth_seen, ic));
  } else {
    EMU_LAUNCH((blocks2), (b200::kEvalWarps * 32), 0,
               b200::NodeReduceKernel<false>(ell, eval_nodes.data(), &ts, nodes.data(), d_fs, n_fs, p, &q, scratch.data(),
                                             cand.data(), mono, meta[5], &max_depth_seen, ic));
  }
  if (meta[9])
    EMU_LAUNCH((blocks2), (b200::kEvalWarps * 32), cat_smem,
               b200::CatFinishKernel(ell, hist.data(), eval_nodes.data(), &ts, nodes.data(), p, &q, cand.data(), cat));
  WriteAll(dir + "/cand.bin", cand);
  WriteAll(dir + "/node_is_cat.bin", node_is_cat);
  WriteAll(dir + "/cat_bits.bin", cat_bits);
  std::vector<int32_t> open((size_t)n_eval);
  for (int k = 0; k < n_eval; ++k) open[k] = nodes[k].cand_open;
  WriteAll(dir + "/cand_open.bin", open);
  return 0;
}
