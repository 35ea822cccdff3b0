// Runs ColSampleStreamKernel and PredictMarginKernel / PredictLeafKernel -- the kernel SOURCES of
// xgboost_b200/csrc, compiled for the CPU emulator -- on inputs read from stdin-named files and writes their outputs.
// Driven by tests/test_kernel_emulation.py, which compares with the oracle.
#define B200_KERNELS_ONLY 1
#include "cuda_emu.h"

#include <cstdio>
#include <string>

#include "../../xgboost_b200/csrc/colsample.cu"
#include <fstream>

#include "../../xgboost_b200/csrc/predict.cu"

template <typename T>
static std::vector<T> ReadAll(const std::string& path) {
  std::vector<T> out;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize((size_t)n / sizeof(T));
  if (n && std::fread(out.data(), 1, (size_t)n, f) != (size_t)n) std::exit(2);
  std::fclose(f);
  return out;
}
template <typename T>
static void WriteAll(const std::string& path, const std::vector<T>& v) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) std::exit(2);
  std::fwrite(v.data(), sizeof(T), v.size(), f);
  std::fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1], dir = argv[2];
  if (mode == "colsample") {
    auto features = ReadAll<int32_t>(dir + "/features.bin");
    auto seeds = ReadAll<uint32_t>(dir + "/seeds.bin");
    auto meta = ReadAll<int32_t>(dir + "/meta.bin");  // n_sel
    const int32_t n = (int32_t)features.size(), n_sel = meta[0];
    std::vector<int32_t> out((size_t)seeds.size() * n_sel, -1);
    EMU_LAUNCH(((unsigned)seeds.size()), (b200::kCsThreads), sizeof(int32_t) * 2 * (size_t)n_sel,
               b200::ColSampleStreamKernel(features.data(), n, n_sel, seeds.data(), out.data()));
    WriteAll(dir + "/out.bin", out);
    return 0;
  }
  if (mode == "predict") {
    // meta: n_rows, n_features, row_bytes, bin_bytes, null_symbol, n_trees, bin_words
    auto meta = ReadAll<uint32_t>(dir + "/meta.bin");
    auto bins = ReadAll<uint8_t>(dir + "/bins.bin");
    auto nodes = ReadAll<b200::PredNode>(dir + "/nodes.bin");
    auto roots = ReadAll<uint32_t>(dir + "/roots.bin");
    auto bin_bits = ReadAll<uint32_t>(dir + "/bin_bits.bin");
    auto cut_ptrs = ReadAll<uint32_t>(dir + "/cut_ptrs.bin");
    auto cut_values = ReadAll<float>(dir + "/cut_values.bin");
    auto base = ReadAll<float>(dir + "/base.bin");
    b200::EllView ell{};
    ell.bins = bins.data();
    ell.cut_ptrs = cut_ptrs.data();
    ell.cut_values = cut_values.data();
    ell.n_rows = meta[0];
    ell.n_features = meta[1];
    ell.total_bins = (uint32_t)cut_values.size();
    ell.row_bytes = meta[2];
    ell.bin_bytes = meta[3];
    ell.null_symbol = meta[4];
    ell.bins_t = nullptr;
    ell.is_sparse = 0;
    ell.row_sig = nullptr;
    if (meta.size() > 7 && meta[7] > 0) {  // sparse ELLPACK: uint32 global bins, row_stride entries per row
      ell.is_sparse = 1;
      ell.row_stride = meta[7];
    }
    const b200::ForestView fv{nodes.data(), roots.data(), bin_bits.data(), (int32_t)meta[6]};
    std::vector<float> out = base;
    std::vector<int32_t> leaves((size_t)meta[0] * meta[5]);
    if (ell.is_sparse) {
      EMU_LAUNCH((3), (256), 0, b200::PredictMarginKernel<uint32_t>(ell, fv, 0, meta[5], 0, meta[0], out.data()));
      EMU_LAUNCH((3), (256), 0, b200::PredictLeafKernel<uint32_t>(ell, fv, 0, meta[5], 0, meta[0], leaves.data()));
    } else if (meta[3] == 1) {
      EMU_LAUNCH((3), (256), 0, b200::PredictMarginKernel<uint8_t>(ell, fv, 0, meta[5], 0, meta[0], out.data()));
      EMU_LAUNCH((3), (256), 0, b200::PredictLeafKernel<uint8_t>(ell, fv, 0, meta[5], 0, meta[0], leaves.data()));
    } else {
      EMU_LAUNCH((3), (256), 0, b200::PredictMarginKernel<uint16_t>(ell, fv, 0, meta[5], 0, meta[0], out.data()));
      EMU_LAUNCH((3), (256), 0, b200::PredictLeafKernel<uint16_t>(ell, fv, 0, meta[5], 0, meta[0], leaves.data()));
    }
    WriteAll(dir + "/margin.bin", out);
    WriteAll(dir + "/leaves.bin", leaves);
    // optional: several output groups (tree_groups.bin, base_groups.bin [n_rows * n_groups]) and float leaf ids
    if (std::ifstream(dir + "/tree_groups.bin").good()) {
      auto groups = ReadAll<uint32_t>(dir + "/tree_groups.bin");
      auto gout = ReadAll<float>(dir + "/base_groups.bin");
      const uint32_t n_groups = (uint32_t)(gout.size() / meta[0]);
      b200::ForestView fg = fv;
      fg.tree_group = groups.data();
      std::vector<float> fleaves((size_t)meta[0] * meta[5]);
      if (ell.is_sparse) {
        EMU_LAUNCH((3), (256), 0, b200::PredictMarginGroupsKernel<uint32_t>(ell, fg, 0, meta[5], n_groups, 0, meta[0], gout.data()));
        EMU_LAUNCH((3), (256), 0, (b200::PredictLeafKernel<uint32_t, float>(ell, fv, 0, meta[5], 0, meta[0], fleaves.data())));
      } else if (meta[3] == 1) {
        EMU_LAUNCH((3), (256), 0, b200::PredictMarginGroupsKernel<uint8_t>(ell, fg, 0, meta[5], n_groups, 0, meta[0], gout.data()));
        EMU_LAUNCH((3), (256), 0, (b200::PredictLeafKernel<uint8_t, float>(ell, fv, 0, meta[5], 0, meta[0], fleaves.data())));
      } else {
        EMU_LAUNCH((3), (256), 0, b200::PredictMarginGroupsKernel<uint16_t>(ell, fg, 0, meta[5], n_groups, 0, meta[0], gout.data()));
        EMU_LAUNCH((3), (256), 0, (b200::PredictLeafKernel<uint16_t, float>(ell, fv, 0, meta[5], 0, meta[0], fleaves.data())));
      }
      WriteAll(dir + "/margin_groups.bin", gout);
      WriteAll(dir + "/leaves_f32.bin", fleaves);
      // the same three results straight from the reference's compressed page (ref_page.bin + ref_meta.bin:
      // row_stride, n_symbols, dense_layout, is_dense)
      if (std::ifstream(dir + "/ref_page.bin").good()) {
        auto ref = ReadAll<uint8_t>(dir + "/ref_page.bin");
        auto rm = ReadAll<uint32_t>(dir + "/ref_meta.bin");
        b200::RefPageView pg{};
        pg.buf = ref.data();
        pg.cut_ptrs = cut_ptrs.data();
        pg.symbol_bits = 1;
        while ((1ull << pg.symbol_bits) < rm[1]) ++pg.symbol_bits;
        pg.row_stride = rm[0];
        pg.ref_null = (rm[2] && !rm[3]) ? rm[1] - 1 : 0xFFFFFFFFu;
        pg.total_bins = (uint32_t)cut_values.size();
        pg.dense_layout = (int32_t)rm[2];
        std::vector<float> r0 = base, r1 = ReadAll<float>(dir + "/base_groups.bin"), r2((size_t)meta[0] * meta[5]);
        EMU_LAUNCH((3), (256), 0, b200::PredictRefKernel<0>(pg, fg, 0, meta[5], n_groups, meta[0], r0.data()));
        EMU_LAUNCH((3), (256), 0, b200::PredictRefKernel<1>(pg, fg, 0, meta[5], n_groups, meta[0], r1.data()));
        EMU_LAUNCH((3), (256), 0, b200::PredictRefKernel<2>(pg, fg, 0, meta[5], n_groups, meta[0], r2.data()));
        WriteAll(dir + "/ref_margin.bin", r0);
        WriteAll(dir + "/ref_margin_groups.bin", r1);
        WriteAll(dir + "/ref_leaves.bin", r2);
      }
    }
    return 0;
  }
  return 2;
}
