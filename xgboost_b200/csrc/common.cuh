// Internal shared declarations for libb200hist.so (sm_100a only).
// Nothing here crosses the C ABI (include/b200hist.h).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <functional>
#include <vector>

#include "b200hist.h"

// Dynamic shared memory of a kernel.  The CPU emulator of the tests (tests/emu/cuda_emu.h) supplies its own definition
// so that the kernel sources run unmodified there; B200_KERNELS_ONLY (also set by those tests) compiles the host
// launchers out, which the emulator cannot parse (<<< >>>).
// Body of a flag spin loop (comm.cu): nothing on the device; the emulator yields the core to the process it waits for.
#ifndef B200_SPIN_PAUSE
#define B200_SPIN_PAUSE() ((void)0)
#endif
#ifndef B200_DYNAMIC_SHARED
#define B200_DYNAMIC_SHARED(type, name) extern __shared__ type name[]
#define B200_DYNAMIC_SHARED_ALIGNED(type, name, bytes) extern __shared__ __align__(bytes) type name[]
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Errors: C++ exceptions inside, converted to status + thread-local message at the ABI.
// Mirrors API_BEGIN/API_END of the reference (src/c_api/c_api_error.h:27-42).
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
void SetLastError(const std::string& msg);

#define B200_CHECK(cond, msg)                                                              \
  do {                                                                                     \
    if (!(cond)) throw ::b200::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + \
                                     ": check failed: " #cond ": " + (msg));               \
  } while (0)

#define B200_CUDA(expr)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      throw ::b200::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": CUDA " + \
                          cudaGetErrorName(e__) + ": " + cudaGetErrorString(e__) + " in " #expr); \
  } while (0)

#define B200_API_BEGIN() try {
#define B200_API_END()                       \
  }                                          \
  catch (std::exception const& e__) {        \
    ::b200::SetLastError(e__.what());        \
    return -1;                               \
  }                                          \
  return 0;

extern std::atomic<uint64_t> g_launch_count;
// Count + check every kernel launch.
#define B200_LAUNCHED()                          \
  do {                                           \
    ::b200::g_launch_count.fetch_add(1);         \
    B200_CUDA(cudaGetLastError());               \
  } while (0)

// A launch whose static + dynamic shared memory passes 48 KiB needs the per-kernel opt-in.  Called right before such a
// launch with exactly the dynamic bytes it asks for (the static part of every kernel here is below 18 KiB, so dynamic
// sizes up to 16 KiB can never reach the limit and skip the call).
#ifndef B200_KERNELS_ONLY
template <typename Kernel>
inline void AllowDynamicShared(Kernel kernel, size_t dynamic_bytes) {
  if (dynamic_bytes > 16 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynamic_bytes));
  }
}
#endif

struct DeviceGuard {
  int prev{-1};
  explicit DeviceGuard(int dev) {
    B200_CUDA(cudaGetDevice(&prev));
    if (prev != dev) B200_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

template <typename T>
T* DevAlloc(size_t n) {
  void* p = nullptr;
  B200_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
  return static_cast<T*>(p);
}

int NumSMs(int device);

// ---------------------------------------------------------------------------------------------
// Ellpack page (ours).  bins[n_rows][row_bytes], feature-local symbols, see include/b200hist.h.
// ---------------------------------------------------------------------------------------------
struct EllView {
  const uint8_t* bins;
  const uint32_t* cut_ptrs;  // device, n_features+1
  const float* cut_values;   // device, total_bins
  uint64_t n_rows;
  uint32_t n_features;
  uint32_t total_bins;
  uint32_t row_bytes;    // multiple of 16
  uint32_t bin_bytes;    // 1 or 2
  uint32_t null_symbol;  // symbol meaning "missing"; 0xFFFFFFFF when the page has no missing
  // column-major copy bins_t[f][col_stride] (uint8 pages only, may be null): one byte per (feature, row),
  // used by the single-feature gathers of the row partitioner (32 consecutive rows share a sector)
  const uint8_t* bins_t;
  uint64_t col_stride;
  // sparse ELLPACK layout (ellpack_page.cu:140-183 "Ellpack is sparse"): bins = uint32[n_rows][row_stride]
  // holding GLOBAL bin indices in feature order, padded with the null symbol total_bins
  uint32_t is_sparse;
  uint32_t row_stride;
  // sparse pages: 64-bit presence signature per row, bit (f mod 64) set for every feature f the row holds.  A row
  // whose bit for the split feature is clear certainly lacks the feature, so the partitioner reads 8 bytes of it
  // instead of its whole 4*row_stride-byte entry list (may be null)
  const unsigned long long* row_sig;
  // External-memory pages (b200hist_ellpack_create_paged): `bins` addresses ONE resident page, biased so that
  // `bins + r * row_bytes` is row r for the page's GLOBAL row ids r in [page_lo, page_hi); rows outside that range
  // are not resident and must not be read.  In-core pages: page_lo = 0, page_hi = 0xFFFFFFFF (every row).
  uint32_t page_lo = 0u;
  uint32_t page_hi = 0xFFFFFFFFu;
};

#ifdef __CUDACC__
// A node's row segment lists its rows in ascending order (the partition is stable), so the rows of one page are a
// contiguous slice of it: [lo, hi) -> the slice whose row ids fall in [page_lo, page_hi).
__device__ __forceinline__ void ClipToPage(const uint32_t* __restrict__ ridx, uint32_t& lo, uint32_t& hi,
                                           uint32_t page_lo, uint32_t page_hi) {
  uint32_t a = lo, b = hi;
  while (a < b) {  // first index with ridx >= page_lo
    const uint32_t mid = (a + b) >> 1;
    if (__ldg(ridx + mid) < page_lo) a = mid + 1; else b = mid;
  }
  const uint32_t first = a;
  b = hi;
  while (a < b) {  // first index with ridx >= page_hi
    const uint32_t mid = (a + b) >> 1;
    if (__ldg(ridx + mid) < page_hi) a = mid + 1; else b = mid;
  }
  lo = first;
  hi = a;
}
#endif

}  // namespace b200

struct b200hist_comm_s {
  int device{0};
  int world{1};
  int rank{0};
  void* comm{nullptr};  // ncclComm_t
};

struct b200hist_ellpack_s {
  int device{0};
  uint64_t n_rows{0};
  uint32_t n_features{0};
  uint32_t total_bins{0};
  uint32_t row_bytes{0};
  uint32_t bin_bytes{1};
  uint32_t null_symbol{0xFFFFFFFFu};
  uint32_t max_feature_bins{0};
  bool has_missing{false};
  std::vector<uint32_t> cut_ptrs;
  std::vector<float> cut_values;
  uint32_t* d_cut_ptrs{nullptr};
  float* d_cut_values{nullptr};
  uint8_t* d_bins{nullptr};
  uint8_t* d_bins_t{nullptr};  // column-major copy (uint8 pages)
  uint64_t col_stride{0};
  bool is_sparse{false};
  uint32_t row_stride{0};  // entries per row (sparse layout)
  unsigned long long* d_row_sig{nullptr};  // presence signatures (sparse layout)
  // categorical features (dense layout): ids of the categorical columns, re-binned by SearchCatBin after the
  // numerical binning kernel has run (nullptr = none)
  uint32_t* d_cat_features{nullptr};
  uint32_t n_cat_features{0};
  // External memory (b200hist_ellpack_create_paged): the quantised matrix lives in PINNED HOST memory (one
  // allocation, row-major, our layout); page p = rows [p * page_rows, min(n_rows, (p + 1) * page_rows)).  The updater
  // streams the pages through two device staging buffers (double-buffered H2D on `copy_stream`); incoming batches
  // are binned / decoded into `d_import` on the device and copied out.  d_bins stays null.
  bool paged{false};
  uint64_t page_rows{0};
  uint8_t* h_bins{nullptr};
  uint8_t* d_stage[2]{nullptr, nullptr};
  uint8_t* d_import{nullptr};
  size_t import_bytes{0};
  void* copy_stream{nullptr};              // cudaStream_t
  void* ev_ready[2]{nullptr, nullptr};     // cudaEvent_t: the page has landed in stage b (recorded on copy_stream)
  void* ev_done[2]{nullptr, nullptr};      // cudaEvent_t: the kernels reading stage b have finished (compute stream)
  uint64_t stage_uses{0};                  // staging buffer of the next page = stage_uses & 1
  uint64_t h2d_bytes{0};                   // bytes streamed from the host so far (bench / tests)
  uint32_t NumPages() const { return paged ? (uint32_t)((n_rows + page_rows - 1) / page_rows) : 1u; }
  b200::EllView View() const {
    return b200::EllView{d_bins, d_cut_ptrs, d_cut_values, n_rows, n_features, total_bins,
                         row_bytes, bin_bytes, null_symbol, d_bins_t, col_stride,
                         is_sparse ? 1u : 0u, row_stride, d_row_sig, 0u, 0xFFFFFFFFu};
  }
  // the view of page p once it is resident in staging buffer b
  b200::EllView PageView(uint32_t p, int b) const {
    b200::EllView v = View();
    const uint64_t lo = (uint64_t)p * page_rows;
    const uint64_t hi = lo + page_rows < n_rows ? lo + page_rows : n_rows;
    v.bins = d_stage[b] - lo * (uint64_t)row_bytes;  // biased: never dereferenced outside [lo, hi)
    v.bins_t = nullptr;
    v.page_lo = (uint32_t)lo;
    v.page_hi = (uint32_t)hi;
    return v;
  }
};

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Tree-building state shared by the kernels.  All of it lives in device memory so a whole tree
// is built without the host reading anything back (no per-level sync).
// ---------------------------------------------------------------------------------------------
constexpr int kPartTile = 2048;       // rows per partition tile (64 flag words)
constexpr int kHistTile = 512;        // rows per histogram tile
constexpr int kMaxLevelNodes = 4096;  // nodes expanded in one step
constexpr int kMaxCatBins = 1024;     // bins of one categorical feature (8 warps x 6 B x 1024 = 48 KiB of sort scratch)

struct NodeDev {
  int32_t parent, left, right, depth;
  int32_t findex;
  float fvalue;
  int32_t default_left;
  int32_t is_leaf;
  float leaf_value, loss_chg, sum_hess, base_weight;
  int64_t sum_g, sum_h;  // quantised gradient sum of the node's rows
  uint32_t seg_begin, seg_count;
  int32_t seg_buf;       // which ridx buffer holds the node's row segment
  int32_t partitioned;   // rows have been handed to the children
  int32_t split_lbin;    // go left iff feature-local bin < split_lbin
  int32_t cand_open;     // has a queued expand entry (driver.h Push filter passed)
  int32_t hist_slot;     // slot of the node's histogram in the histogram cache (-1 = none)
  int32_t pad;
};

struct PartEntry {
  int32_t nidx, left, right;
  int32_t findex;
  int32_t split_lbin;
  int32_t default_left;
  uint32_t begin, count;
  uint32_t tile_begin;
  int32_t src_buf;
  uint32_t n_left;  // filled by the scan kernel
  uint32_t pad;
};

struct BuildEntry {
  int32_t nidx;
  int32_t buf;
  uint32_t begin, count;
  uint32_t tile_begin;
  uint32_t slot;  // histogram slot the rows accumulate into
};

struct SubEntry {
  int32_t parent, built, sub;  // histogram SLOTS
};

struct TreeState {
  int32_t n_nodes;
  int32_t num_leaves;
  int32_t level_begin, level_end;  // node-id range of the open level (depth-wise)
  // Node batching (Driver::Pop returns at most max_node_batch_size entries of one level, driver.h:71-100;
  // updater_gpu_hist.cuh:34 kMaxNodeBatchSize): a level wider than a step's capacity is expanded by several steps.
  // level_cursor = first node of the level no step has consumed yet; next_begin = first node of the NEXT level
  // (the children created so far by this level's steps).
  int32_t level_cursor, next_begin;
  int32_t n_part;
  int32_t n_build;
  int32_t n_sub;
  int32_t n_eval;
  uint32_t part_tiles;
  uint32_t build_tiles;
  uint32_t final_tiles;
  int32_t n_final;
  unsigned long long row_visits;
  int32_t stopped;  // lossguide: Pop() returned an invalid top -> growth ended
  int32_t max_eval_depth;  // deepest node that went through split evaluation
};

struct QuantDev {
  double to_fixed_g, to_fixed_h, to_float_g, to_float_h;
  double sums[4];
  long long root_g, root_h;
  unsigned int h_zero_any;  // some quantised hessian is 0 (position must be bit-complemented for those rows)
  unsigned int pad0;
  unsigned int h_lo_or;  // OR of the low 32 bits of every quantised hessian (0 => histogram skips h.lo)
  unsigned int pad;
};

// Monotone constraints (TreeEvaluator, src/tree/split_evaluator.h:25-205): sign per feature (device, n_features
// entries), weight bounds per node (device, max_nodes entries).  constraints == nullptr: no constraints.
struct MonoView {
  const int32_t* constraints;
  float* lower;
  float* upper;
};

// Feature interaction constraints (FeatureInteractionConstraintDevice, src/tree/constraints.cu) as the evaluation
// kernels see them.  Every Query of the reference answers into the front of ONE device buffer and the spans of a
// whole level are collected before the evaluation launch (constraints.cu:125,199-240; updater_gpu_hist.cu:251-268),
// so the kernel reads the buffer as the level's queries left it, cut to each node's own count: `list` is that buffer
// image, count[slot] what evaluation slot `slot` reads of it.  count == nullptr: no constraints (and always for the
// root, which Query never filters).
// The same view carries colsample_bynode: every evaluation slot then owns a list (stride > 0: slot s reads
// list + s * stride), the next draws of the tree's stream of node feature sets (NodeSetLevelKernel, updater.cu).
struct IcView {
  const int32_t* list;
  const int32_t* count;
  int32_t stride;  // 0: one list shared by all slots
};

// Categorical features (src/common/categorical.h, evaluate_splits.cu:148-252,313-348, evaluator.cu:69-128).
// A categorical feature's bins ARE its categories: cut_values[b] is the category code of bin b.  is_cat == nullptr:
// every feature is numerical.  Per node the winner's category set is kept twice: cat_bits = the reference's
// LBitField32 over category codes (category c = bit 31 - c % 32 of word c / 32; what RegTree stores) and bin_bits =
// the same set over the feature's LOCAL bins (bin b = bit b % 32 of word b / 32; what the partitioner tests).
// A row whose category is in the set goes RIGHT (common::Decision, categorical.h:49-60).
struct CatView {
  const uint8_t* is_cat;       // [n_features]
  int32_t max_cat_to_onehot;   // fewer bins than this: one-hot split (UseOneHot, categorical.h:86-89)
  int32_t max_cat_threshold;   // sorted-partition scan looks at this many categories from either end
  uint32_t* cat_bits;          // [max_nodes][cat_words]
  uint32_t* bin_bits;          // [max_nodes][bin_words]
  uint8_t* node_is_cat;        // [max_nodes] the node's candidate is a categorical split
  int32_t cat_words, bin_words;
  int32_t max_cat_bins;        // most bins of any categorical feature (sizes the sort scratch)
};

// Row routing of a categorical split: NodeDev/PartEntry::split_lbin == -1 marks it, the node's set over the split
// feature's LOCAL bins says which rows go RIGHT (CatView::bin_bits).  kCat == false compiles to the numerical test only.
struct CatRoute {
  const uint32_t* bin_bits;  // [max_nodes][bin_words]
  int32_t bin_words;
};
template <bool kCat>
__device__ __forceinline__ bool GoLeftBin(int32_t lbin, int32_t split_lbin, int nidx, const CatRoute& cr) {
  if constexpr (kCat) {
    if (split_lbin < 0) return ((__ldg(cr.bin_bits + (size_t)nidx * cr.bin_words + (lbin >> 5)) >> (lbin & 31)) & 1u) == 0u;
  }
  return lbin < split_lbin;
}

struct ParamDev {
  int32_t max_depth, max_leaves, grow_policy;
  float learning_rate, reg_lambda, reg_alpha, min_child_weight, max_delta_step, min_split_loss;
};

inline ParamDev ToParamDev(const b200hist_train_param& p) {
  return ParamDev{p.max_depth, p.max_leaves, p.grow_policy, p.learning_rate, p.reg_lambda,
                  p.reg_alpha, p.min_child_weight, p.max_delta_step, p.min_split_loss};
}

#ifdef __CUDACC__
// Sparse page: global bin of feature [fb, fe) in row r, or 0xFFFFFFFF.  Entries are in feature order, so the
// row (<= row_stride symbols, null-padded with values >= total_bins) is binary-searched
// (EllpackAccessor::GetBinIndex / BinarySearchBin, src/data/ellpack_page.cuh:101-122).
__device__ __forceinline__ uint32_t SparseRowBin(const EllView& ell, uint32_t r, uint32_t fb, uint32_t fe,
                                                 uint32_t f) {
  if (ell.row_sig != nullptr && ((__ldg(ell.row_sig + r) >> (f & 63u)) & 1ull) == 0ull) return 0xFFFFFFFFu;
  const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(ell.bins) + (size_t)r * ell.row_stride;
  uint32_t lo = 0, hi = ell.row_stride;
  while (lo < hi) {  // first entry with symbol >= fb
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(row + mid) < fb) lo = mid + 1; else hi = mid;
  }
  if (lo < ell.row_stride) {
    const uint32_t s = __ldg(row + lo);
    if (s < fe) return s;
  }
  return 0xFFFFFFFFu;
}
#endif

// Feature group: a contiguous feature range whose bins fit the shared-memory histogram.
struct GroupDesc {
  uint32_t f_begin, f_end;      // features
  uint32_t bin_begin, bin_end;  // global bins
  uint32_t byte_begin;          // byte offset of f_begin inside a row (multiple of 4)
  uint32_t n_words;             // 32-bit words per row covering the group
  uint32_t n_blocks;            // bank-mapped layout: 32-feature blocks in the group (0 = generic layout)
  uint32_t max_sym;             // bank-mapped layout: symbols per feature slot (max bins of the group)
};

// Row split over several GPUs, evaluation by feature slice: rank r of `world` owns the features f with
// f * world / n_features == r -- after the reduce step it holds the complete histograms of exactly those features'
// bins, scans only them and pushes their per-feature winners to the peers (the reference's column-split evaluation,
// evaluate_splits.cu:412-435, applied to the slices a reduce-scatter leaves behind).  world == 1: everything.
struct EvalShard {
  int32_t rank;
  int32_t world;
  uint32_t n_features;
  __host__ __device__ bool Owns(int f) const {
    return world <= 1 || (int32_t)(((unsigned long long)(uint32_t)f * (uint32_t)world) / n_features) == rank;
  }
};

// Per-(node, feature) best split produced by the feature scan.
struct FeatCand {
  float loss_chg;
  int32_t dir;
  int32_t gbin;  // -1 = none
  int32_t pad;
  long long left_g, left_h;
};

// ---------------------------------------------------------------------------------------------
// Kernel launchers (one per SURVEY.md 8(a) row)
// ---------------------------------------------------------------------------------------------
// quantise.cu
void LaunchQuantise(const float* gpair, uint64_t n, QuantDev* q, b200hist_gpair64* gq, uint32_t* ridx_init,
                    double* partials /* >= 4*kQuantBlocks */, unsigned int* counter, int num_sms, cudaStream_t s,
                    float subsample = 1.0f, uint32_t sample_seed = 0);
void LaunchQuantReduce(const float* gpair, uint64_t n, QuantDev* q, double* partials, unsigned int* counter,
                       int num_sms, cudaStream_t s);
void LaunchQuantConvert(const float* gpair, uint64_t n, QuantDev* q, b200hist_gpair64* gq, uint32_t* ridx_init,
                        int num_sms, cudaStream_t s, float subsample = 1.0f, uint32_t sample_seed = 0);
constexpr int kQuantBlocksPerSM = 4;

// histogram.cu
void PrepareHistKernels(int device);
void LaunchBuildHist(const EllView& ell, const b200hist_gpair64* gq, const uint32_t* ridx0, const uint32_t* ridx1,
                     const BuildEntry* build, const uint32_t* tile_seg, const TreeState* ts,
                     const GroupDesc* groups_dev, const std::vector<GroupDesc>& groups_host,
                     b200hist_gpair64* hist, const QuantDev* quant /* nullable */, int num_sms, cudaStream_t s,
                     bool pair_hint = false /* expect hessians without low words: 64-feature pair kernel */);
void LaunchSubtract(const SubEntry* sub, const TreeState* ts, b200hist_gpair64* hist, uint32_t total_bins,
                    int num_sms, cudaStream_t s, uint32_t bin_begin = 0, uint32_t bin_end = 0xFFFFFFFFu);
void LaunchSubtractFlat(const b200hist_gpair64* parent, const b200hist_gpair64* built, b200hist_gpair64* out,
                        uint64_t n_bins, cudaStream_t s);
std::vector<GroupDesc> MakeFeatureGroups(const b200hist_ellpack_s& ell, int device);

// comm.cu
void CommAllReduceInt64(b200hist_comm_s* c, void* buf, size_t count, cudaStream_t s);
void ShardGroups(uint32_t n_groups, int world, int rank, uint32_t* g0, uint32_t* g1);
void LaunchHistExchange(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                        const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, bool use_n_build,
                        uint32_t total_bins, uint32_t bin_begin, uint32_t bin_end, uint32_t epoch,
                        unsigned int* done_counter, int num_sms, cudaStream_t s);
constexpr int kMaxPeers = 16;
void LaunchHistAllReduce(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                         const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, uint32_t total_bins,
                         uint32_t epoch_a, uint32_t epoch_b, unsigned int* done_counter, int num_sms, cudaStream_t s);
void LaunchScalarExchange(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch, int which /* 0 sums, 1 root */,
                          QuantDev* q, uint64_t n_local, cudaStream_t s);
// row split, evaluation by feature slice: sum the peers' partial histograms over this rank's bin range into its OWN
// cache (reduce-scatter, no all-gather) ...
void LaunchHistSliceReduce(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                           const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, uint32_t total_bins,
                           uint32_t bin_begin, uint32_t bin_end, uint32_t epoch, int num_sms, cudaStream_t s);
// ... and after the feature scan push the owned per-feature winners into every peer's scratch table, then the barrier
void LaunchScratchExchange(FeatCand* const* peer_scratch, uint32_t* const* peer_flags, const EvalShard& shard,
                           const TreeState* ts, const int32_t* eval_nodes, const NodeDev* nodes,
                           const int32_t* feature_set, int32_t n_fs, int32_t n_levels, IcView ic, int32_t max_eval,
                           uint32_t epoch, unsigned int* done_counter, int num_sms, cudaStream_t s);

// evaluate.cu
void LaunchEvaluate(const EllView& ell, const b200hist_gpair64* hist /* slot = node id */,
                    const int32_t* eval_nodes, const TreeState* ts /* n_eval */, int32_t max_eval,
                    const NodeDev* nodes, const int32_t* feature_set, int32_t n_feature_set,
                    const ParamDev& p, const QuantDev* q, FeatCand* scratch, b200hist_candidate* cand,
                    NodeDev* nodes_mut, int num_sms, cudaStream_t s, MonoView mono = MonoView{nullptr, nullptr, nullptr},
                    int32_t n_levels = 0 /* > 0: feature_set is [n_levels][n_feature_set], row = node depth */,
                    int32_t* max_eval_depth = nullptr /* device: atomicMax of the depths evaluated */,
                    IcView ic = IcView{nullptr, nullptr, 0},
                    CatView cat = CatView{nullptr, 0, 0, nullptr, nullptr, nullptr, 0, 0, 0},
                    EvalShard shard = EvalShard{0, 1, 0} /* scan only the features this rank owns */,
                    const std::function<void()>* between = nullptr /* runs after the feature scans, before the fold */);

// colsample.cu: the seed-driven column sampler of one tree (see the kernels there)
struct ColDrawDev {
  const uint32_t* seeds;  // engine outputs after ColumnSampler::Init
  int32_t n_seeds;
  int32_t* cursor;        // [0] outputs consumed so far, [1] outputs the last NodeDrawKernel consumed
  const int32_t* tree_set;
  int32_t n_tree;
  int32_t* level_sets;    // [max_levels][n_level], row = depth (the last row serves deeper nodes)
  int32_t n_level;
  int32_t* level_drawn;   // [max_levels]
  int32_t max_levels;
  int32_t bylevel_is_one;  // colsample_bylevel == 1: a level set is the tree set, no output consumed
};
void LaunchLevelDraw(const TreeState* ts, const NodeDev* nodes, const int32_t* eval_nodes, const ColDrawDev& cd,
                     cudaStream_t s);
void LaunchNodeDraw(const TreeState* ts, const NodeDev* nodes, const int32_t* eval_nodes, const ColDrawDev& cd,
                    int32_t n_node, int32_t max_eval, int32_t* list, int32_t* count, cudaStream_t s);

// sampling.cu: gradient-based row sampling on the quantised pairs (after LaunchQuantise without uniform sampling);
// recomputes the record's root sum and hessian flags from the sampled pairs
size_t GradientBasedScratchBytes(uint64_t n, int num_sms);
void LaunchGradientBasedSample(b200hist_gpair64* gq, uint64_t n, QuantDev* q, float subsample, uint32_t seed,
                               void* scratch, size_t scratch_bytes, int num_sms, cudaStream_t s);

// partition.cu
void LaunchPartitionCount(const EllView& ell, const uint32_t* ridx0, const uint32_t* ridx1, const PartEntry* part,
                          const uint32_t* tile_seg, const TreeState* ts, uint32_t* flags, uint32_t* tile_left,
                          int num_sms, cudaStream_t s, CatRoute cr = CatRoute{nullptr, 0});
void LaunchPartitionScatter(const uint32_t* ridx0, const uint32_t* ridx1, uint32_t* out0, uint32_t* out1,
                            const PartEntry* part, const uint32_t* tile_seg, const TreeState* ts,
                            const uint32_t* flags, const uint32_t* tile_left_off, int num_sms, cudaStream_t s);

}  // namespace b200
