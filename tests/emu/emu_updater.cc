// Control kernels of xgboost_b200/csrc/updater.cu and the categorical re-binning kernel of ellpack.cu under the CPU
// emulator: NodeSetLevelKernel, IcLevelKernel, RebinCategoricalKernel.
#define B200_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "emu_io.h"

#include "../../xgboost_b200/csrc/ellpack.cu"
#include "../../xgboost_b200/csrc/updater.cu"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1], dir = argv[2];
  if (mode == "nodeset") {
    auto meta = ReadAll<int32_t>(dir + "/meta.bin");  // n_sets, n_per, seq0, n_eval (several steps: n_eval list)
    auto sets = ReadAll<int32_t>(dir + "/sets.bin");
    auto steps = ReadAll<int32_t>(dir + "/steps.bin");
    const int n_sets = meta[0], n_per = meta[1];
    int32_t seq = meta[2];
    std::vector<int32_t> all_lists, all_counts;
    for (int32_t n_eval : steps) {
      b200::TreeState ts{};
      ts.n_eval = n_eval;
      std::vector<int32_t> list((size_t)std::max(1, n_eval) * n_per, -1), count((size_t)std::max(1, n_eval), -1);
      EMU_LAUNCH((1), (b200::kCtlThreads), 0,
                 b200::NodeSetLevelKernel(&ts, sets.data(), n_sets, n_per, &seq, list.data(), count.data()));
      all_lists.insert(all_lists.end(), list.begin(), list.begin() + (size_t)n_eval * n_per);
      all_counts.insert(all_counts.end(), count.begin(), count.begin() + n_eval);
    }
    WriteAll(dir + "/lists.bin", all_lists);
    WriteAll(dir + "/counts.bin", all_counts);
    WriteAll(dir + "/seq.bin", std::vector<int32_t>{seq});
    return 0;
  }
  if (mode == "iclevel") {
    // meta: n_features, n_sets, n_nodes, n_part, n_fs, n_levels, has_fs
    auto meta = ReadAll<int32_t>(dir + "/meta.bin");
    auto set_ptr = ReadAll<int32_t>(dir + "/set_ptr.bin");
    auto set_feat = ReadAll<int32_t>(dir + "/set_feat.bin");
    auto bits = ReadAll<unsigned long long>(dir + "/bits.bin");
    auto part_raw = ReadAll<int32_t>(dir + "/part.bin");  // n_part x (nidx, findex, left, right)
    auto depth = ReadAll<int32_t>(dir + "/depth.bin");
    auto fs = ReadAll<int32_t>(dir + "/feature_set.bin", true);
    const int nf = meta[0], n_nodes = meta[2], n_part = meta[3];
    const int words = (nf + 63) / 64;
    std::vector<b200::NodeDev> nodes((size_t)n_nodes);
    for (int i = 0; i < n_nodes; ++i) { nodes[i] = b200::NodeDev{}; nodes[i].depth = depth[i]; }
    std::vector<b200::PartEntry> part((size_t)std::max(1, n_part));
    std::vector<int32_t> eval_nodes((size_t)std::max(1, 2 * n_part));
    for (int k = 0; k < n_part; ++k) {
      b200::PartEntry e{};
      e.nidx = part_raw[4 * k]; e.findex = part_raw[4 * k + 1]; e.left = part_raw[4 * k + 2]; e.right = part_raw[4 * k + 3];
      part[k] = e;
      eval_nodes[2 * k] = e.left;
      eval_nodes[2 * k + 1] = e.right;
    }
    b200::TreeState ts{};
    ts.n_part = n_part;
    ts.n_eval = 2 * n_part;
    std::vector<int32_t> list((size_t)nf, -7), count((size_t)std::max(1, 2 * n_part), -7);
    const b200::IcDev icd{nf, words, meta[1], set_ptr.data(), set_feat.data(), bits.data(), list.data(), count.data()};
    EMU_LAUNCH((1), (b200::kCtlThreads), sizeof(unsigned long long) * (size_t)words,
               b200::IcLevelKernel(&ts, nodes.data(), part.data(), eval_nodes.data(), icd, meta[6] ? fs.data() : nullptr,
                                   meta[4], meta[5], nullptr, 0));
    WriteAll(dir + "/bits_out.bin", bits);
    WriteAll(dir + "/list.bin", list);
    WriteAll(dir + "/count.bin", count);
    return 0;
  }
  if (mode == "rebin") {
    // meta: n_rows, ld, n_features, row_bytes, bin_bytes, null_symbol
    auto meta = ReadAll<uint32_t>(dir + "/meta.bin");
    auto x = ReadAll<float>(dir + "/x.bin");
    auto cat_features = ReadAll<uint32_t>(dir + "/cat_features.bin");
    auto cut_ptrs = ReadAll<uint32_t>(dir + "/cut_ptrs.bin");
    auto cut_values = ReadAll<float>(dir + "/cut_values.bin");
    auto bins = ReadAll<uint8_t>(dir + "/bins.bin");
    if (meta[4] == 1) {
      EMU_LAUNCH((2), (256), 0,
                 b200::RebinCategoricalKernel<uint8_t>(x.data(), meta[0], meta[1], cat_features.data(),
                                                       (uint32_t)cat_features.size(), cut_ptrs.data(), cut_values.data(),
                                                       bins.data(), 0, meta[3], meta[5]));
    } else {
      EMU_LAUNCH((2), (256), 0,
                 b200::RebinCategoricalKernel<uint16_t>(x.data(), meta[0], meta[1], cat_features.data(),
                                                        (uint32_t)cat_features.size(), cut_ptrs.data(), cut_values.data(),
                                                        bins.data(), 0, meta[3], meta[5]));
    }
    WriteAll(dir + "/bins_out.bin", bins);
    return 0;
  }
  return 2;
}
