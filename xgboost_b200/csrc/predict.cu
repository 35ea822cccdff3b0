// Prediction on the quantised matrix: the step that follows the updater in every boosting
// iteration whenever the prediction cache misses (validation sets, a fresh DMatrix bound to the training cuts).
//
// Replaces (reference paths relative to /root/reference):
//   PredictKernel with EllpackLoader, PredictBatch over Ellpack pages    src/predictor/gpu_predictor.cu:180-230,316-351
//   EllpackDeviceAccessor::GetFvalue (bin -> representative value)       src/data/ellpack_page.cuh:147-170
//   common::Decision (categorical routing)                               src/common/categorical.h:49-60
//
// The reference turns a row's bin back into a feature value (the bin's lower cut, gpu_data_accessor.cuh:95-107) and
// compares it with the split condition.  That comparison is decided by the bin alone: the forest is resolved
// against the page's cuts once, when a tree is added (split_lbin = 1 + number of the feature's cuts < the
// condition; for a tree grown on these cuts that is the updater's ApplySplit threshold), and the kernel compares
// small integers.  Categorical splits carry the
// set of the feature's local bins that go right.  A row sums the leaves of trees [begin, end) in tree order into a
// float that starts at 0 and is then added to the output (gpu_predictor.cu:221-228) -- the order is part of the
// result.
//
// One thread per row, grid-strided; a row's symbols are one contiguous run of the row-major page (<= 512 B at the
// headline width), so after the first tree they are L1/L2 hits; the forest (24 B per node) is read through the
// read-only path.  HBM-bound on the first touch of each row: row_bytes per row + 4 B output.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace b200 {

struct PredNode {
  int32_t findex;       // -1: leaf
  int32_t split_lbin;   // go left iff local bin < split_lbin; -1: categorical (bin set)
  int32_t left, right;  // absolute indices into the forest's node array
  float leaf_value;
  int32_t default_left;
};

struct ForestView {
  const PredNode* nodes;
  const uint32_t* tree_root;  // [n_trees] index of every tree's root
  const uint32_t* bin_bits;   // [n_nodes][bin_words] (categorical nodes only meaningful), may be null
  int32_t bin_words;
  const uint32_t* tree_group{nullptr};  // [n_trees] output group of every tree (multi-class), null: all 0
};

template <typename BinT>
__device__ __forceinline__ float WalkTree(const EllView& ell, const ForestView& fv, uint32_t root, uint64_t r) {
  uint32_t i = root;
  PredNode nd = fv.nodes[i];
  while (nd.findex >= 0) {
    int32_t lbin;
    if (ell.is_sparse) {
      const uint32_t fb = ell.cut_ptrs[nd.findex], fe = ell.cut_ptrs[nd.findex + 1];
      const uint32_t gb = SparseRowBin(ell, (uint32_t)r, fb, fe, (uint32_t)nd.findex);
      lbin = gb == 0xFFFFFFFFu ? -1 : (int32_t)(gb - fb);
    } else {
      const uint32_t sym = (uint32_t)__ldg(reinterpret_cast<const BinT*>(ell.bins + r * ell.row_bytes) + nd.findex);
      lbin = sym == ell.null_symbol ? -1 : (int32_t)sym;
    }
    bool left;
    if (lbin < 0) {
      left = nd.default_left != 0;
    } else if (nd.split_lbin < 0) {  // categorical: a bin in the set goes right
      left = ((__ldg(fv.bin_bits + (size_t)i * fv.bin_words + (lbin >> 5)) >> (lbin & 31)) & 1u) == 0u;
    } else {
      left = lbin < nd.split_lbin;
    }
    i = (uint32_t)(left ? nd.left : nd.right);
    nd = fv.nodes[i];
  }
  return nd.leaf_value;
}

template <typename BinT>
__global__ void __launch_bounds__(256) PredictMarginKernel(EllView ell, ForestView fv, uint32_t tree_begin,
                                                            uint32_t tree_end, uint64_t row_begin, uint64_t n_rows,
                                                            float* __restrict__ out) {
  for (uint64_t k = blockIdx.x * 256ull + threadIdx.x; k < n_rows; k += (uint64_t)gridDim.x * 256ull) {
    const uint64_t r = row_begin + k;
    float sum = 0.0f;
    for (uint32_t t = tree_begin; t < tree_end; ++t) sum += WalkTree<BinT>(ell, fv, fv.tree_root[t], r);
    out[k] += sum;
  }
}

// Several output groups (multi-class: tree t feeds output tree_group[t]).  The reference's kernel then adds every
// leaf STRAIGHT onto the output, tree after tree (gpu_predictor.cu:229-239) -- not into a sum that starts at 0 as in
// the single-group branch -- so per output the chain is out, out + l1, (out + l1) + l2, ...  One thread per
// (row, group): the output is read once, the group's trees of [begin, end) are walked in order, and it is written once.
template <typename BinT>
__global__ void __launch_bounds__(256) PredictMarginGroupsKernel(EllView ell, ForestView fv, uint32_t tree_begin,
                                                                  uint32_t tree_end, uint32_t n_groups,
                                                                  uint64_t row_begin, uint64_t n_rows,
                                                                  float* __restrict__ out) {
  for (uint64_t k = blockIdx.x * 256ull + threadIdx.x; k < n_rows * n_groups; k += (uint64_t)gridDim.x * 256ull) {
    const uint64_t row = k / n_groups;
    const uint32_t g = (uint32_t)(k - row * n_groups);
    float acc = out[k];
    for (uint32_t t = tree_begin; t < tree_end; ++t) {
      if (__ldg(fv.tree_group + t) == g) acc += WalkTree<BinT>(ell, fv, fv.tree_root[t], row_begin + row);
    }
    out[k] = acc;
  }
}

// PredictLeaf (gpu_predictor.cu:196-201): out[row * n_trees + t] = id of the leaf inside tree t
// (OutT = float is the layout Predictor::PredictLeaf hands back: the id as a float.)
template <typename BinT, typename OutT = int32_t>
__global__ void __launch_bounds__(256) PredictLeafKernel(EllView ell, ForestView fv, uint32_t tree_begin,
                                                          uint32_t tree_end, uint64_t row_begin, uint64_t n_rows,
                                                          OutT* __restrict__ out) {
  const uint32_t nt = tree_end - tree_begin;
  for (uint64_t k = blockIdx.x * 256ull + threadIdx.x; k < n_rows * nt; k += (uint64_t)gridDim.x * 256ull) {
    const uint64_t row = k / nt;
    const uint32_t t = tree_begin + (uint32_t)(k - row * nt);
    const uint64_t r = row_begin + row;
    uint32_t i = fv.tree_root[t];
    PredNode nd = fv.nodes[i];
    while (nd.findex >= 0) {
      int32_t lbin;
      if (ell.is_sparse) {
        const uint32_t fb = ell.cut_ptrs[nd.findex], fe = ell.cut_ptrs[nd.findex + 1];
        const uint32_t gb = SparseRowBin(ell, (uint32_t)r, fb, fe, (uint32_t)nd.findex);
        lbin = gb == 0xFFFFFFFFu ? -1 : (int32_t)(gb - fb);
      } else {
        const uint32_t sym = (uint32_t)__ldg(reinterpret_cast<const BinT*>(ell.bins + r * ell.row_bytes) + nd.findex);
        lbin = sym == ell.null_symbol ? -1 : (int32_t)sym;
      }
      bool left;
      if (lbin < 0) left = nd.default_left != 0;
      else if (nd.split_lbin < 0)
        left = ((__ldg(fv.bin_bits + (size_t)i * fv.bin_words + (lbin >> 5)) >> (lbin & 31)) & 1u) == 0u;
      else left = lbin < nd.split_lbin;
      i = (uint32_t)(left ? nd.left : nd.right);
      nd = fv.nodes[i];
    }
    out[k] = (OutT)(int32_t)(i - fv.tree_root[t]);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// The same walks straight on a page of the REFERENCE (EllpackPageImpl::gidx_buffer: 8 bytes of padding, then an
// MSB-first bit stream of ceil(log2(n_symbols))-bit symbols, src/common/compressed_iterator.h:37-40,140-152,228-268),
// so that the Predictor binding does not have to copy the page first.  Dense layout: row_stride == n_features, symbols
// are feature-local bins, null = n_symbols - 1 when the page has missing values.  Sparse layout: global bins in
// ascending order, null = total_bins at the end of a row; a feature is found by binary search over the row
// (EllpackAccessorImpl::GetBinIndex, src/data/ellpack_page.cuh:101-122).
// ---------------------------------------------------------------------------------------------------------------------
struct RefPageView {
  const uint8_t* buf;         // gidx_buffer, including its 8 bytes of padding
  const uint32_t* cut_ptrs;   // the forest's cuts (device)
  uint32_t symbol_bits;
  uint32_t row_stride;
  uint32_t ref_null;          // dense layout: the null symbol, 0xFFFFFFFF without missing values
  uint32_t total_bins;
  int32_t dense_layout;
};

__device__ __forceinline__ uint32_t RefSymbol(const uint8_t* __restrict__ buf, uint64_t i, uint32_t bits) {
  const uint64_t bit0 = i * bits;
  const uint64_t byte0 = bit0 >> 3;
  const uint32_t shift = (uint32_t)(bit0 & 7);
  const uint32_t nbytes = (shift + bits + 7) >> 3;
  uint64_t v = 0;
  for (uint32_t k = 0; k < nbytes; ++k) v = (v << 8) | __ldg(buf + 8 + byte0 + k);
  return (uint32_t)((v >> (8 * nbytes - shift - bits)) & ((1ull << bits) - 1ull));
}

// local bin of (row, feature), -1 when missing
__device__ __forceinline__ int32_t RefLocalBin(const RefPageView& pg, uint64_t r, uint32_t f) {
  if (pg.dense_layout) {
    const uint32_t sym = RefSymbol(pg.buf, r * pg.row_stride + f, pg.symbol_bits);
    return sym == pg.ref_null ? -1 : (int32_t)sym;
  }
  const uint32_t fb = pg.cut_ptrs[f], fe = pg.cut_ptrs[f + 1];
  uint32_t lo = 0, hi = pg.row_stride;  // first slot whose symbol >= fb
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (RefSymbol(pg.buf, r * pg.row_stride + mid, pg.symbol_bits) < fb) lo = mid + 1; else hi = mid;
  }
  if (lo == pg.row_stride) return -1;
  const uint32_t sym = RefSymbol(pg.buf, r * pg.row_stride + lo, pg.symbol_bits);
  return sym < fe ? (int32_t)(sym - fb) : -1;
}

__device__ __forceinline__ uint32_t WalkTreeRef(const RefPageView& pg, const ForestView& fv, uint32_t root, uint64_t r) {
  uint32_t i = root;
  PredNode nd = fv.nodes[i];
  while (nd.findex >= 0) {
    const int32_t lbin = RefLocalBin(pg, r, (uint32_t)nd.findex);
    bool left;
    if (lbin < 0) {
      left = nd.default_left != 0;
    } else if (nd.split_lbin < 0) {
      left = ((__ldg(fv.bin_bits + (size_t)i * fv.bin_words + (lbin >> 5)) >> (lbin & 31)) & 1u) == 0u;
    } else {
      left = lbin < nd.split_lbin;
    }
    i = (uint32_t)(left ? nd.left : nd.right);
    nd = fv.nodes[i];
  }
  return i;
}

// kMode 0: one output per row (sum from 0, then added); 1: n_groups outputs per row (leaves straight onto the output);
// 2: leaf ids as floats, [row][tree].
template <int kMode>
__global__ void __launch_bounds__(256) PredictRefKernel(RefPageView pg, ForestView fv, uint32_t tree_begin,
                                                         uint32_t tree_end, uint32_t n_groups, uint64_t n_rows,
                                                         float* __restrict__ out) {
  const uint32_t nt = tree_end - tree_begin;
  const uint64_t per_row = kMode == 0 ? 1 : (kMode == 1 ? n_groups : nt);
  for (uint64_t k = blockIdx.x * 256ull + threadIdx.x; k < n_rows * per_row; k += (uint64_t)gridDim.x * 256ull) {
    const uint64_t row = k / per_row;
    const uint32_t sub = (uint32_t)(k - row * per_row);
    if (kMode == 0) {
      float sum = 0.0f;
      for (uint32_t t = tree_begin; t < tree_end; ++t) sum += fv.nodes[WalkTreeRef(pg, fv, fv.tree_root[t], row)].leaf_value;
      out[k] += sum;
    } else if (kMode == 1) {
      float acc = out[k];
      for (uint32_t t = tree_begin; t < tree_end; ++t) {
        if (__ldg(fv.tree_group + t) == sub) acc += fv.nodes[WalkTreeRef(pg, fv, fv.tree_root[t], row)].leaf_value;
      }
      out[k] = acc;
    } else {
      const uint32_t t = tree_begin + sub;
      out[k] = (float)(int32_t)(WalkTreeRef(pg, fv, fv.tree_root[t], row) - fv.tree_root[t]);
    }
  }
}

}  // namespace b200

#ifndef B200_KERNELS_ONLY
using namespace b200;

struct b200hist_forest_s {
  b200hist_ellpack_t ell{nullptr};  // the cuts the split conditions are resolved against (borrowed)
  int device{0};
  std::vector<PredNode> nodes;
  std::vector<uint32_t> roots;
  std::vector<uint32_t> bin_bits;  // [nodes][bin_words]
  std::vector<uint32_t> groups;    // [trees] output group of every tree (0 unless set_tree_groups said otherwise)
  int32_t bin_words{0};
  // device copy, refreshed lazily
  PredNode* d_nodes{nullptr};
  uint32_t* d_roots{nullptr};
  uint32_t* d_bin_bits{nullptr};
  uint32_t* d_groups{nullptr};
  uint32_t* d_cut_ptrs{nullptr};  // device copy of the cuts' feature segments (reference-page path)
  bool dirty{true};
};

template <typename OutT>
static int PredictLeafImpl(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin, uint32_t tree_end,
                           uint64_t row_begin, uint64_t n_rows, OutT* out_dev, void* stream);

extern "C" {

int b200hist_forest_create(b200hist_ellpack_t ell, b200hist_forest_t* out) {
  B200_API_BEGIN();
  B200_CHECK(ell != nullptr && out != nullptr, "null argument");
  auto* f = new b200hist_forest_s();
  f->ell = ell;
  f->device = ell->device;
  f->bin_words = (int32_t)((ell->max_feature_bins + 31) / 32);
  if (f->bin_words < 1) f->bin_words = 1;
  *out = f;
  B200_API_END();
}

int b200hist_forest_destroy(b200hist_forest_t f) {
  B200_API_BEGIN();
  if (f) {
    DeviceGuard guard(f->device);
    cudaFree(f->d_nodes);
    cudaFree(f->d_roots);
    cudaFree(f->d_bin_bits);
    cudaFree(f->d_groups);
    cudaFree(f->d_cut_ptrs);
    delete f;
  }
  B200_API_END();
}

int b200hist_forest_num_trees(b200hist_forest_t f, uint32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && out != nullptr, "null argument");
  *out = (uint32_t)f->roots.size();
  B200_API_END();
}

int b200hist_forest_add_tree(b200hist_forest_t f, const b200hist_tree_node* nodes_host, int32_t n_nodes,
                             const int32_t* node_is_cat_host, const uint32_t* cat_bits_host, uint32_t cat_words) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && nodes_host != nullptr && n_nodes >= 1, "null argument");
  const auto* ell = f->ell;
  const uint32_t base = (uint32_t)f->nodes.size();
  std::vector<PredNode> add((size_t)n_nodes);
  std::vector<uint32_t> bits((size_t)n_nodes * f->bin_words, 0u);
  for (int32_t i = 0; i < n_nodes; ++i) {
    const b200hist_tree_node& s = nodes_host[i];
    PredNode d{};
    d.leaf_value = s.leaf_value;
    d.default_left = s.default_left;
    if (s.is_leaf) {
      d.findex = -1; d.split_lbin = 0; d.left = d.right = (int32_t)(base + i);
    } else {
      B200_CHECK(s.findex >= 0 && (uint32_t)s.findex < ell->n_features, "split feature out of range");
      B200_CHECK(s.left > i && s.left < n_nodes && s.right > i && s.right < n_nodes, "child index out of range");
      d.findex = s.findex;
      d.left = (int32_t)(base + s.left);
      d.right = (int32_t)(base + s.right);
      const uint32_t beg = ell->cut_ptrs[s.findex], end = ell->cut_ptrs[s.findex + 1];
      if (node_is_cat_host != nullptr && node_is_cat_host[i]) {
        B200_CHECK(cat_bits_host != nullptr && cat_words >= 1, "categorical node without a category set");
        d.split_lbin = -1;
        for (uint32_t b = beg; b < end; ++b) {  // the feature's bins whose category is in the set
          const float cv = ell->cut_values[b];
          if (!(cv >= 0.0f)) continue;
          const uint32_t c = (uint32_t)cv;
          if (c / 32 < cat_words && ((cat_bits_host[(size_t)i * cat_words + c / 32] >> (31 - c % 32)) & 1u))
            bits[(size_t)i * f->bin_words + ((b - beg) >> 5)] |= 1u << ((b - beg) & 31);
        }
      } else {
        B200_CHECK(!std::isnan(s.fvalue), "numerical split with a NaN condition");
        // The reference's loader hands out the bin's lower cut (-inf for the first bin) and goes left iff that value
        // < fvalue (gpu_data_accessor.cuh:95-107): local bin b goes left iff b == 0 or cut[b-1] < fvalue, i.e. iff
        // b < 1 + (number of the feature's cuts < fvalue).  For a condition that is one of the cuts this is the
        // updater's cut[b] <= fvalue; it also holds for trees grown on other cuts.
        const float* lo = ell->cut_values.data() + beg;
        const float* hi = ell->cut_values.data() + end;
        d.split_lbin = s.fvalue > -INFINITY ? 1 + (int32_t)(std::lower_bound(lo, hi, s.fvalue) - lo) : 0;
      }
    }
    add[i] = d;
  }
  f->roots.push_back(base);
  f->groups.push_back(0u);
  f->nodes.insert(f->nodes.end(), add.begin(), add.end());
  f->bin_bits.insert(f->bin_bits.end(), bits.begin(), bits.end());
  f->dirty = true;
  B200_API_END();
}

static void SyncForest(b200hist_forest_s* f, cudaStream_t s) {
  if (!f->dirty) return;
  cudaFree(f->d_nodes);
  cudaFree(f->d_roots);
  cudaFree(f->d_bin_bits);
  cudaFree(f->d_groups);
  f->d_groups = DevAlloc<uint32_t>(std::max<size_t>(f->groups.size(), 1));
  B200_CUDA(cudaMemcpyAsync(f->d_groups, f->groups.data(), sizeof(uint32_t) * f->groups.size(), cudaMemcpyHostToDevice, s));
  f->d_nodes = DevAlloc<PredNode>(std::max<size_t>(f->nodes.size(), 1));
  f->d_roots = DevAlloc<uint32_t>(std::max<size_t>(f->roots.size(), 1));
  f->d_bin_bits = DevAlloc<uint32_t>(std::max<size_t>(f->bin_bits.size(), 1));
  B200_CUDA(cudaMemcpyAsync(f->d_nodes, f->nodes.data(), sizeof(PredNode) * f->nodes.size(), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(f->d_roots, f->roots.data(), sizeof(uint32_t) * f->roots.size(), cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(f->d_bin_bits, f->bin_bits.data(), sizeof(uint32_t) * f->bin_bits.size(),
                            cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaStreamSynchronize(s));  // the host vectors may grow again right after this call
  f->dirty = false;
}

static void CheckPage(const b200hist_forest_s* f, const b200hist_ellpack_s* page) {
  B200_CHECK(page->device == f->device, "the page lives on another device than the forest");
  B200_CHECK(page->n_features == f->ell->n_features && page->cut_ptrs == f->ell->cut_ptrs &&
                 page->cut_values == f->ell->cut_values,
             "the page must be quantised with the cuts the forest was resolved against");
}

int b200hist_forest_predict_margin(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin,
                                   uint32_t tree_end, uint64_t row_begin, uint64_t n_rows, float* out_dev,
                                   void* stream) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && page != nullptr && (out_dev != nullptr || n_rows == 0), "null argument");
  B200_CHECK(tree_begin <= tree_end && tree_end <= f->roots.size(), "tree range out of bounds");
  B200_CHECK(row_begin + n_rows <= page->n_rows, "row range exceeds the page");
  CheckPage(f, page);
  if (n_rows == 0 || tree_begin == tree_end) return 0;
  DeviceGuard guard(f->device);
  auto s = static_cast<cudaStream_t>(stream);
  SyncForest(f, s);
  const ForestView fv{f->d_nodes, f->d_roots, f->d_bin_bits, f->bin_words};
  B200_CHECK(!page->paged, "prediction on external-memory page sets is not implemented: predict page by page");
  const EllView view = page->View();
  const int blocks = (int)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)NumSMs(f->device) * 16);
  if (page->is_sparse) {
    PredictMarginKernel<uint32_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  } else if (page->bin_bytes == 1) {
    PredictMarginKernel<uint8_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  } else {
    PredictMarginKernel<uint16_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  }
  B200_LAUNCHED();
  B200_API_END();
}

int b200hist_forest_set_tree_groups(b200hist_forest_t f, const uint32_t* groups_host, uint32_t n_trees) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && (groups_host != nullptr || n_trees == 0), "null argument");
  B200_CHECK(n_trees == f->roots.size(), "one group per tree of the forest");
  f->groups.assign(groups_host, groups_host + n_trees);
  f->dirty = true;
  B200_API_END();
}

int b200hist_forest_predict_margin_groups(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin,
                                          uint32_t tree_end, uint32_t n_groups, uint64_t row_begin, uint64_t n_rows,
                                          float* out_dev, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && page != nullptr && (out_dev != nullptr || n_rows == 0), "null argument");
  B200_CHECK(n_groups >= 1, "at least one output group");
  B200_CHECK(tree_begin <= tree_end && tree_end <= f->roots.size(), "tree range out of bounds");
  B200_CHECK(row_begin + n_rows <= page->n_rows, "row range exceeds the page");
  for (uint32_t t = tree_begin; t < tree_end; ++t) B200_CHECK(f->groups[t] < n_groups, "tree group out of range");
  CheckPage(f, page);
  if (n_rows == 0 || tree_begin == tree_end) return 0;
  DeviceGuard guard(f->device);
  auto s = static_cast<cudaStream_t>(stream);
  SyncForest(f, s);
  ForestView fv{f->d_nodes, f->d_roots, f->d_bin_bits, f->bin_words};
  fv.tree_group = f->d_groups;
  B200_CHECK(!page->paged, "prediction on external-memory page sets is not implemented: predict page by page");
  const EllView view = page->View();
  const uint64_t total = n_rows * (uint64_t)n_groups;
  const int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)NumSMs(f->device) * 16);
  if (page->is_sparse) {
    PredictMarginGroupsKernel<uint32_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, n_groups, row_begin,
                                                               n_rows, out_dev);
  } else if (page->bin_bytes == 1) {
    PredictMarginGroupsKernel<uint8_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, n_groups, row_begin,
                                                              n_rows, out_dev);
  } else {
    PredictMarginGroupsKernel<uint16_t><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, n_groups, row_begin,
                                                               n_rows, out_dev);
  }
  B200_LAUNCHED();
  B200_API_END();
}

int b200hist_forest_predict_leaf(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin, uint32_t tree_end,
                                 uint64_t row_begin, uint64_t n_rows, int32_t* out_dev, void* stream) {
  return PredictLeafImpl<int32_t>(f, page, tree_begin, tree_end, row_begin, n_rows, out_dev, stream);
}

int b200hist_forest_predict_leaf_f32(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin,
                                     uint32_t tree_end, uint64_t row_begin, uint64_t n_rows, float* out_dev,
                                     void* stream) {
  return PredictLeafImpl<float>(f, page, tree_begin, tree_end, row_begin, n_rows, out_dev, stream);
}

int b200hist_forest_predict_reference_page(b200hist_forest_t f, const uint8_t* gidx_buffer_dev, uint64_t n_rows,
                                           uint32_t row_stride, uint32_t n_symbols, int dense_layout_in,
                                           int is_dense, int mode,
                                           uint32_t tree_begin, uint32_t tree_end, uint32_t n_groups, float* out_dev,
                                           void* stream) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && (n_rows == 0 || (gidx_buffer_dev != nullptr && out_dev != nullptr)), "null argument");
  B200_CHECK(mode >= 0 && mode <= 2, "mode: 0 margin, 1 margin per group, 2 leaf ids");
  B200_CHECK(tree_begin <= tree_end && tree_end <= f->roots.size(), "tree range out of bounds");
  B200_CHECK(n_symbols >= 1 && row_stride >= 1, "empty page geometry");
  const auto* ell = f->ell;
  const bool dense_layout = dense_layout_in != 0;
  if (!dense_layout) {
    B200_CHECK(n_symbols == ell->total_bins + 1, "sparse reference page: n_symbols must be total_bins + 1");
  } else {
    B200_CHECK(row_stride == ell->n_features, "dense-layout reference pages need row_stride == n_features");
    B200_CHECK(n_symbols >= ell->max_feature_bins + (is_dense ? 0u : 1u),
               "dense reference page: n_symbols below the widest feature (+1 with missing values)");
  }
  if (mode == 1) {
    B200_CHECK(n_groups >= 1, "at least one output group");
    for (uint32_t t = tree_begin; t < tree_end; ++t) B200_CHECK(f->groups[t] < n_groups, "tree group out of range");
  }
  if (n_rows == 0 || tree_begin == tree_end) return 0;
  DeviceGuard guard(f->device);
  auto s = static_cast<cudaStream_t>(stream);
  SyncForest(f, s);
  if (f->d_cut_ptrs == nullptr) {
    f->d_cut_ptrs = DevAlloc<uint32_t>(ell->cut_ptrs.size());
    B200_CUDA(cudaMemcpyAsync(f->d_cut_ptrs, ell->cut_ptrs.data(), sizeof(uint32_t) * ell->cut_ptrs.size(),
                              cudaMemcpyHostToDevice, s));
    B200_CUDA(cudaStreamSynchronize(s));
  }
  ForestView fv{f->d_nodes, f->d_roots, f->d_bin_bits, f->bin_words};
  fv.tree_group = f->d_groups;
  RefPageView pg{};
  pg.buf = gidx_buffer_dev;
  pg.cut_ptrs = f->d_cut_ptrs;
  pg.symbol_bits = 1;
  while ((1ull << pg.symbol_bits) < n_symbols) ++pg.symbol_bits;  // SymbolBits (compressed_iterator.h:37-40)
  pg.row_stride = row_stride;
  pg.ref_null = (dense_layout && !is_dense) ? n_symbols - 1 : 0xFFFFFFFFu;
  pg.total_bins = ell->total_bins;
  pg.dense_layout = dense_layout ? 1 : 0;
  const uint64_t per_row = mode == 0 ? 1 : (mode == 1 ? n_groups : tree_end - tree_begin);
  const int blocks = (int)std::min<uint64_t>((n_rows * per_row + 255) / 256, (uint64_t)NumSMs(f->device) * 16);
  if (mode == 0) {
    PredictRefKernel<0><<<blocks, 256, 0, s>>>(pg, fv, tree_begin, tree_end, n_groups, n_rows, out_dev);
  } else if (mode == 1) {
    PredictRefKernel<1><<<blocks, 256, 0, s>>>(pg, fv, tree_begin, tree_end, n_groups, n_rows, out_dev);
  } else {
    PredictRefKernel<2><<<blocks, 256, 0, s>>>(pg, fv, tree_begin, tree_end, n_groups, n_rows, out_dev);
  }
  B200_LAUNCHED();
  B200_API_END();
}

}  // extern "C"

template <typename OutT>
static int PredictLeafImpl(b200hist_forest_t f, b200hist_ellpack_t page, uint32_t tree_begin, uint32_t tree_end,
                           uint64_t row_begin, uint64_t n_rows, OutT* out_dev, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(f != nullptr && page != nullptr && (out_dev != nullptr || n_rows == 0), "null argument");
  B200_CHECK(tree_begin <= tree_end && tree_end <= f->roots.size(), "tree range out of bounds");
  B200_CHECK(row_begin + n_rows <= page->n_rows, "row range exceeds the page");
  CheckPage(f, page);
  if (n_rows == 0 || tree_begin == tree_end) return 0;
  DeviceGuard guard(f->device);
  auto s = static_cast<cudaStream_t>(stream);
  SyncForest(f, s);
  const ForestView fv{f->d_nodes, f->d_roots, f->d_bin_bits, f->bin_words};
  B200_CHECK(!page->paged, "prediction on external-memory page sets is not implemented: predict page by page");
  const EllView view = page->View();
  const uint64_t total = n_rows * (uint64_t)(tree_end - tree_begin);
  const int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)NumSMs(f->device) * 16);
  if (page->is_sparse) {
    PredictLeafKernel<uint32_t, OutT><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  } else if (page->bin_bytes == 1) {
    PredictLeafKernel<uint8_t, OutT><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  } else {
    PredictLeafKernel<uint16_t, OutT><<<blocks, 256, 0, s>>>(view, fv, tree_begin, tree_end, row_begin, n_rows, out_dev);
  }
  B200_LAUNCHED();
  B200_API_END();
}
#endif  // B200_KERNELS_ONLY
