// Column sampling draws on the device: many ColSample results at once, one CTA per draw.
//
// Replaces (reference paths relative to /root/reference):
//   ColumnSampler::ColSample on CUDA, no feature weights      src/common/random.cc:27-46
//   cuda_impl::SampleFeature: thrust::shuffle + sort           src/common/random.cu:62-96
//   thrust::shuffle = feistel_bijection + flagged gather       <toolkit>/thrust/system/detail/generic/shuffle.inl
//
// The reference draws a node's feature set with one host call each (one engine draw, a device shuffle, a device sort,
// a few allocations): up to 2 * n_nodes calls per tree.  A tree's draws depend only on their seeds, so this kernel takes
// the whole stream of seeds and produces all sets in one launch: CTA d seeds thrust's minstd_rand with seeds[d], takes
// the 24 round keys, walks the cipher domain [0, 2^bits) in index order keeping the images below n (a block scan
// gives each kept element its output position -- the same compaction thrust's inclusive scan of flags performs),
// stops after n_sel elements and sorts them ascending (rank sort in shared memory).
// Pinned on the CPU: the same arithmetic in oracle/hist_oracle.c (orc_thrust_shuffle) reproduces permutations of the
// toolkit's own thrust::shuffle (tests/golden/thrust_shuffle_known_answers.json).  STATUS: green on a B200 since round 2 (tests/test_gpu_next.py).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b200 {

constexpr int kCsThreads = 256;
constexpr int kCsMaxFeatures = 8192;  // 2 x 32 KiB of shared memory (selected ids + sort keys): opt-in above 48 KiB

__device__ __forceinline__ uint32_t BlockExclusiveScanCs(uint32_t v, uint32_t* s_warp, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kCsThreads / 32 ? s_warp[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kCsThreads / 32) s_warp[lane] = winc - w;
    if (lane == kCsThreads / 32 - 1) s_warp[32] = winc;
  }
  __syncthreads();
  const uint32_t ex = s_warp[warp] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return ex;
}

// One draw by the whole CTA: ColSample(features[0..n), n_sel, seed) -> out[0..n_sel), ascending.  s_sel: 2 * n_sel ints.
__device__ __forceinline__ void ColSampleBlock(const int32_t* __restrict__ features, int32_t n, int32_t n_sel,
                                               uint32_t seed, int32_t* s_sel, uint32_t* s_key, uint32_t* s_warp,
                                               int32_t* __restrict__ out) {
  int32_t* s_sorted = s_sel + n_sel;
  __syncthreads();  // the caller's earlier use of the scratch is over
  if (threadIdx.x == 0) {  // minstd_rand seeded like thrust: seed mod m, 0 -> 1; a key is the next state
    const unsigned long long mod = 2147483647ull;
    unsigned long long x = seed % mod;
    if (x == 0) x = 1;
    for (int i = 0; i < 24; ++i) {
      x = (x * 48271ull) % mod;
      s_key[i] = (uint32_t)x;
    }
  }
  __syncthreads();
  unsigned long long bits = 4;
  if (n > 16) {
    unsigned long long t = (unsigned long long)n - 1;
    bits = 0;
    while (t != 0) { ++bits; t >>= 1; }
  }
  const unsigned long long left_bits = bits / 2, right_bits = bits - left_bits;
  const uint32_t left_mask = (uint32_t)((1ull << left_bits) - 1), right_mask = (uint32_t)((1ull << right_bits) - 1);
  const unsigned long long domain = 1ull << bits;
  uint32_t running = 0;
  for (unsigned long long base = 0; base < domain && running < (uint32_t)n_sel; base += kCsThreads) {
    const unsigned long long idx = base + threadIdx.x;
    uint32_t flag = 0;
    unsigned long long image = 0;
    if (idx < domain) {
      uint32_t s0 = (uint32_t)(idx >> right_bits), s1 = (uint32_t)(idx & right_mask);
#pragma unroll 4
      for (int i = 0; i < 24; ++i) {
        const unsigned long long product = 0xD2B74407B1CE6E93ull * (unsigned long long)s0;
        const uint32_t hi = (uint32_t)(product >> 32);
        uint32_t lo = (uint32_t)product;
        lo = (lo << (right_bits - left_bits)) | (s1 >> left_bits);
        s0 = ((hi ^ s_key[i]) ^ s1) & left_mask;
        s1 = lo & right_mask;
      }
      image = ((unsigned long long)s0 << right_bits) | (unsigned long long)s1;
      flag = image < (unsigned long long)n ? 1u : 0u;
    }
    uint32_t total;
    const uint32_t pos = running + BlockExclusiveScanCs(flag, s_warp, &total);
    if (flag && pos < (uint32_t)n_sel) s_sel[pos] = features[image];
    running += total;
  }
  __syncthreads();
  // ascending order (thrust::sort of distinct ids): rank = number of smaller ids
  for (int i = threadIdx.x; i < n_sel; i += kCsThreads) {
    const int32_t v = s_sel[i];
    int rank = 0;
    for (int j = 0; j < n_sel; ++j) rank += s_sel[j] < v ? 1 : 0;
    s_sorted[rank] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_sel; i += kCsThreads) out[i] = s_sorted[i];
}

__global__ void __launch_bounds__(kCsThreads) ColSampleStreamKernel(const int32_t* __restrict__ features, int32_t n,
                                                                     int32_t n_sel, const uint32_t* __restrict__ seeds,
                                                                     int32_t* __restrict__ out) {
  B200_DYNAMIC_SHARED(int32_t, s_sel);  // n_sel selected ids, then n_sel sorted ids
  __shared__ uint32_t s_key[24];
  __shared__ uint32_t s_warp[33];
  ColSampleBlock(features, n, n_sel, seeds[blockIdx.x], s_sel, s_key, s_warp, out + (size_t)blockIdx.x * n_sel);
}

// ---------------------------------------------------------------------------------------------
// The column sampler of one tree, driven by the engine's outputs (ColumnSampler::GetFeatureSet, random.h:156-175).
// A GetFeatureSet call for a node of depth d first draws the depth's level set if it has none yet
// (colsample_bylevel < 1: one engine output), then the node's set from it (colsample_bynode < 1: one output); the calls
// come in evaluation order: the root, then left and right child of every valid candidate in Pop order.  With both
// fractions below 1 the position of a level draw in the stream depends on how many nodes were evaluated before the
// level's first node -- known on the device only.  So the device holds the stream of outputs (`seeds`) and a cursor.
// All evaluation slots of one step share a depth (a level, or the two children of one loss-guided split).
//   LevelDrawKernel  1 CTA : cursor += draws of the previous step; draw the step's level set if it is missing
//   NodeDrawKernel   1 CTA per evaluation slot: slot j takes output cursor + j
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCsThreads) LevelDrawKernel(const TreeState* ts, const NodeDev* __restrict__ nodes,
                                                               const int32_t* __restrict__ eval_nodes, ColDrawDev cd) {
  B200_DYNAMIC_SHARED(int32_t, s_sel);
  __shared__ uint32_t s_key[24];
  __shared__ uint32_t s_warp[33];
  __shared__ int s_draw, s_pos;
  if (threadIdx.x == 0) {
    cd.cursor[0] += cd.cursor[1];
    cd.cursor[1] = 0;
    s_draw = 0;
    if (ts->n_eval > 0) {
      const int d = min(nodes[eval_nodes[0]].depth, cd.max_levels - 1);
      if (!cd.level_drawn[d]) {
        s_draw = 1 + d;
        s_pos = cd.cursor[0];
        cd.level_drawn[d] = 1;
        if (!cd.bylevel_is_one) cd.cursor[0] += 1;
      }
    }
  }
  __syncthreads();
  if (s_draw == 0) return;
  int32_t* out = cd.level_sets + (size_t)(s_draw - 1) * cd.n_level;
  if (cd.bylevel_is_one) {
    for (int i = threadIdx.x; i < cd.n_tree; i += kCsThreads) out[i] = cd.tree_set[i];
    return;
  }
  ColSampleBlock(cd.tree_set, cd.n_tree, cd.n_level, cd.seeds[min(s_pos, cd.n_seeds - 1)], s_sel, s_key, s_warp, out);
}

__global__ void __launch_bounds__(kCsThreads) NodeDrawKernel(const TreeState* ts, const NodeDev* __restrict__ nodes,
                                                              const int32_t* __restrict__ eval_nodes, ColDrawDev cd,
                                                              int32_t n_node, int32_t* __restrict__ list,
                                                              int32_t* __restrict__ count) {
  B200_DYNAMIC_SHARED(int32_t, s_sel);
  __shared__ uint32_t s_key[24];
  __shared__ uint32_t s_warp[33];
  const int n_eval = ts->n_eval;
  const int j = blockIdx.x;
  if (j >= n_eval) return;
  const int d = min(nodes[eval_nodes[0]].depth, cd.max_levels - 1);
  const int32_t* level = cd.level_sets + (size_t)d * cd.n_level;
  const int pos = cd.cursor[0] + j;
  ColSampleBlock(level, cd.n_level, n_node, cd.seeds[min(pos, cd.n_seeds - 1)], s_sel, s_key, s_warp,
                 list + (size_t)j * n_node);
  if (threadIdx.x == 0) {
    count[j] = n_node;
    if (j == 0) cd.cursor[1] = n_eval;  // nobody reads [1] in this launch; LevelDrawKernel folds it into the cursor
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchLevelDraw(const TreeState* ts, const NodeDev* nodes, const int32_t* eval_nodes, const ColDrawDev& cd,
                     cudaStream_t s) {
  AllowDynamicShared(LevelDrawKernel, sizeof(int32_t) * 2 * (size_t)cd.n_level);
  LevelDrawKernel<<<1, kCsThreads, sizeof(int32_t) * 2 * (size_t)cd.n_level, s>>>(ts, nodes, eval_nodes, cd);
  B200_LAUNCHED();
}

void LaunchNodeDraw(const TreeState* ts, const NodeDev* nodes, const int32_t* eval_nodes, const ColDrawDev& cd,
                    int32_t n_node, int32_t max_eval, int32_t* list, int32_t* count, cudaStream_t s) {
  if (max_eval == 0) return;
  AllowDynamicShared(NodeDrawKernel, sizeof(int32_t) * 2 * (size_t)n_node);
  NodeDrawKernel<<<max_eval, kCsThreads, sizeof(int32_t) * 2 * (size_t)n_node, s>>>(ts, nodes, eval_nodes, cd, n_node,
                                                                                     list, count);
  B200_LAUNCHED();
}

void LaunchColSampleStream(const int32_t* features_dev, int32_t n, int32_t n_sel, const uint32_t* seeds_dev,
                           int32_t n_draws, int32_t* out_dev, cudaStream_t s) {
  if (n_draws == 0) return;
  AllowDynamicShared(ColSampleStreamKernel, sizeof(int32_t) * 2 * (size_t)n_sel);
  ColSampleStreamKernel<<<n_draws, kCsThreads, sizeof(int32_t) * 2 * (size_t)n_sel, s>>>(features_dev, n, n_sel,
                                                                                         seeds_dev, out_dev);
  B200_LAUNCHED();
}

#endif  // B200_KERNELS_ONLY

}  // namespace b200

#ifndef B200_KERNELS_ONLY
using namespace b200;

// Stand-alone entry: out_host[n_draws][n_sel], n_sel = max(1, int(colsample * n)).
extern "C" int b200hist_col_sample_stream(int device, const int32_t* features_host, uint32_t n, float colsample,
                                          const uint32_t* seeds_host, uint32_t n_draws, int32_t* out_host,
                                          uint32_t* n_sel_out, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(features_host != nullptr && seeds_host != nullptr && out_host != nullptr, "null argument");
  B200_CHECK(n >= 1 && n <= (uint32_t)kCsMaxFeatures, "feature count out of range for the device draw");
  B200_CHECK(colsample > 0.0f && colsample <= 1.0f, "colsample must be in (0, 1]");
  const int32_t n_sel = std::max<int32_t>(1, (int32_t)(colsample * (float)n));
  if (n_sel_out) *n_sel_out = (uint32_t)n_sel;
  if (n_draws == 0) return 0;
  DeviceGuard guard(device);
  auto s = static_cast<cudaStream_t>(stream);
  if (colsample == 1.0f) {  // ColSample returns the input untouched, no engine draw
    for (uint32_t d = 0; d < n_draws; ++d) std::copy(features_host, features_host + n, out_host + (size_t)d * n);
    return 0;
  }
  int32_t* d_feat = DevAlloc<int32_t>(n);
  uint32_t* d_seeds = DevAlloc<uint32_t>(n_draws);
  int32_t* d_out = DevAlloc<int32_t>((size_t)n_draws * n_sel);
  B200_CUDA(cudaMemcpyAsync(d_feat, features_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
  B200_CUDA(cudaMemcpyAsync(d_seeds, seeds_host, sizeof(uint32_t) * n_draws, cudaMemcpyHostToDevice, s));
  LaunchColSampleStream(d_feat, (int32_t)n, n_sel, d_seeds, (int32_t)n_draws, d_out, s);
  B200_CUDA(cudaMemcpyAsync(out_host, d_out, sizeof(int32_t) * (size_t)n_draws * n_sel, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_feat);
  cudaFree(d_seeds);
  cudaFree(d_out);
  B200_API_END();
}
#endif  // B200_KERNELS_ONLY
