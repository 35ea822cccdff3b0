// A seconds-long hardware check through the C ABI only (no Python start-up): reads binary inputs written by
// tools/quick_check/make_inputs.py, grows a few trees with libb200hist.so and writes the node arrays to
// gpurun_out/quick_check/, which tests/quick_check_compare.py then checks against the oracle on the CPU.
// Cases: plain tree, interaction constraints (overlapping sets), categorical features.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/b200hist.h"

#define CK(call)                                                                   \
  do {                                                                             \
    if ((call) != 0) {                                                             \
      std::fprintf(stderr, "FAIL %s: %s\n", #call, b200hist_last_error());         \
      return 1;                                                                    \
    }                                                                              \
  } while (0)
#define CU(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "CUDA %s: %s\n", #call, cudaGetErrorString(e_));        \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

template <typename T>
static std::vector<T> ReadAll(const std::string& path) {
  std::vector<T> out;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize((size_t)n / sizeof(T));
  if (std::fread(out.data(), 1, (size_t)n, f) != (size_t)n) std::exit(2);
  std::fclose(f);
  return out;
}
template <typename T>
static void WriteAll(const std::string& path, const T* p, size_t n) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); std::exit(2); }
  std::fwrite(p, sizeof(T), n, f);
  std::fclose(f);
}

struct Data {
  uint64_t n;
  uint32_t f;
  std::vector<float> x, gpair, cut_values;
  std::vector<uint32_t> cut_ptrs;
  float* d_x{nullptr};
  float* d_gpair{nullptr};
  b200hist_ellpack_t ell{nullptr};
};

static int Load(const std::string& dir, const std::string& name, Data* d) {
  auto meta = ReadAll<uint64_t>(dir + "/" + name + "_meta.bin");
  d->n = meta[0];
  d->f = (uint32_t)meta[1];
  d->x = ReadAll<float>(dir + "/" + name + "_x.bin");
  d->gpair = ReadAll<float>(dir + "/" + name + "_gpair.bin");
  d->cut_ptrs = ReadAll<uint32_t>(dir + "/" + name + "_cut_ptrs.bin");
  d->cut_values = ReadAll<float>(dir + "/" + name + "_cut_values.bin");
  CU(cudaMalloc(&d->d_x, sizeof(float) * d->x.size()));
  CU(cudaMalloc(&d->d_gpair, sizeof(float) * d->gpair.size()));
  CU(cudaMemcpy(d->d_x, d->x.data(), sizeof(float) * d->x.size(), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d->d_gpair, d->gpair.data(), sizeof(float) * d->gpair.size(), cudaMemcpyHostToDevice));
  CK(b200hist_ellpack_create(0, d->n, d->f, d->cut_ptrs.data(), d->cut_values.data(), (int)meta[2], &d->ell));
  CK(b200hist_ellpack_push_dense(d->ell, 0, d->d_x, d->n, d->f, nullptr));
  return 0;
}

static int Grow(const Data& d, b200hist_updater_t u, const std::string& out, const std::string& name, bool cats) {
  int32_t max_nodes = 0, n_nodes = 0;
  CK(b200hist_updater_max_nodes(u, &max_nodes));
  std::vector<b200hist_tree_node> nodes((size_t)max_nodes);
  int32_t* d_pos = nullptr;
  CU(cudaMalloc(&d_pos, sizeof(int32_t) * d.n));
  CK(b200hist_updater_update_tree(u, d.d_gpair, nullptr, nodes.data(), &n_nodes, d_pos));
  std::vector<int32_t> pos(d.n);
  CU(cudaMemcpy(pos.data(), d_pos, sizeof(int32_t) * d.n, cudaMemcpyDeviceToHost));
  WriteAll(out + "/" + name + "_nodes.bin", nodes.data(), (size_t)n_nodes);
  WriteAll(out + "/" + name + "_pos.bin", pos.data(), pos.size());
  if (cats) {
    uint32_t words = 0;
    CK(b200hist_updater_cat_words(u, &words));
    std::vector<int32_t> is_cat((size_t)n_nodes);
    std::vector<uint32_t> bits((size_t)n_nodes * words);
    CK(b200hist_updater_last_tree_categories(u, is_cat.data(), bits.data()));
    WriteAll(out + "/" + name + "_is_cat.bin", is_cat.data(), is_cat.size());
    WriteAll(out + "/" + name + "_cat_bits.bin", bits.data(), bits.size());
  }
  std::printf("%s: %d nodes\n", name.c_str(), n_nodes);
  cudaFree(d_pos);
  return 0;
}

int main(int argc, char** argv) {
  const std::string in = argc > 1 ? argv[1] : "tools/quick_check/inputs";
  const std::string out = argc > 2 ? argv[2] : "gpurun_out/quick_check";
  b200hist_train_param p{};
  p.max_depth = 5; p.learning_rate = 0.5f; p.reg_lambda = 1.0f; p.min_child_weight = 1.0f;
  {  // numerical data: plain tree, then overlapping interaction constraints
    Data d;
    if (Load(in, "num", &d)) return 1;
    b200hist_updater_t u = nullptr;
    CK(b200hist_updater_create(d.ell, &p, &u));
    if (Grow(d, u, out, "plain", false)) return 1;
    const int32_t set_ptr[] = {0, 2, 4, 6, 8, 10};
    const int32_t set_feat[] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 5};
    CK(b200hist_updater_set_interaction_constraints(u, set_ptr, set_feat, 5));
    if (Grow(d, u, out, "interaction", false)) return 1;
    CK(b200hist_updater_set_interaction_constraints(u, nullptr, nullptr, 0));
    if (Grow(d, u, out, "plain_again", false)) return 1;
    CK(b200hist_updater_destroy(u));
  }
  {  // categorical data (category codes shifted by -0.5 so that the numerical binning puts each in its own bin)
    Data d;
    if (Load(in, "cat", &d)) return 1;
    auto flags = ReadAll<uint8_t>(in + "/cat_is_cat.bin");
    b200hist_updater_t u = nullptr;
    CK(b200hist_updater_create(d.ell, &p, &u));
    CK(b200hist_updater_set_categorical(u, flags.data(), (uint32_t)flags.size(), 4, 64));
    if (Grow(d, u, out, "categorical", true)) return 1;
    CK(b200hist_updater_set_categorical(u, flags.data(), (uint32_t)flags.size(), 100, 64));
    if (Grow(d, u, out, "categorical_onehot", true)) return 1;
    CK(b200hist_updater_destroy(u));
  }
  std::printf("quick check done\n");
  return 0;
}
