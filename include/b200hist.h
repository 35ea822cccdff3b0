/*
 * b200hist.h -- C ABI of libb200hist.so, the Blackwell-native (sm_100a) `gpu_hist` tree updater.
 *
 * This is the drop-in boundary of the hot path.  Plain pointers and sizes only: no C++ types, no
 * torch types, no exceptions cross it.  Every function returns 0 on success and non-zero on error;
 * the message is then available from b200hist_last_error() (thread-local), mirroring the
 * reference's `XGB_DLL int fn(...)` + `XGBGetLastError()` contract
 * (/root/reference/include/xgboost/c_api.h:90, src/c_api/c_api_error.h:27-42).
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the `xgboost::TreeUpdater` shim (registered as "grow_gpu_hist") that a
 * maintainer links against these symbols.
 *
 * Pointers suffixed `_dev` are CUDA device pointers on the updater's device; `_host` are host
 * pointers.  `stream` is a `cudaStream_t` passed as `void*` (0 = default stream).  All work is
 * enqueued on that stream; functions that return host results synchronise it before returning.
 */
#ifndef B200HIST_H_
#define B200HIST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200HIST_API __attribute__((visibility("default")))
#define B200HIST_VERSION 100

/* ---------------------------------------------------------------------------------------------
 * Plain-old-data types
 * ------------------------------------------------------------------------------------------- */

/* Subset of TrainParam the updater consumes (src/tree/param.h:25-185) plus GPUTrainingParam
 * (src/tree/updater_gpu_common.cuh:20-40).  Field meanings and defaults are the reference's. */
typedef struct b200hist_train_param {
  int32_t max_depth;        /* default 6; 0 = unlimited (requires max_leaves > 0) */
  int32_t max_leaves;       /* default 0 = unlimited */
  int32_t grow_policy;      /* 0 = depthwise (default), 1 = lossguide */
  float learning_rate;      /* eta, default 0.3 */
  float reg_lambda;         /* default 1 */
  float reg_alpha;          /* default 0 */
  float min_child_weight;   /* default 1 */
  float max_delta_step;     /* default 0 */
  float min_split_loss;     /* gamma, default 0 */
  int32_t reserved[7];      /* [0]: histogram cache capacity in nodes, HistMakerTrainParam::max_cached_hist_node
                             * (src/tree/hist/hist_param.h:18-45); 0 = sized from the free device memory.  A cache
                             * smaller than the tree is a ring: a split whose parent histogram has been overwritten
                             * builds both children (histogram.cuh:72-168, updater_gpu_hist.cu:309-325).  [1..6]: 0 */
} b200hist_train_param;

/* GradientPairInt64 (include/xgboost/base.h:253-304): 64-bit fixed-point gradient statistics. */
typedef struct b200hist_gpair64 {
  int64_t grad;
  int64_t hess;
} b200hist_gpair64;

/* GradientQuantiser state (src/tree/gpu_hist/quantiser.cuh:52-81). */
typedef struct b200hist_quantiser {
  double to_fixed_grad, to_fixed_hess;
  double to_float_grad, to_float_hess;
} b200hist_quantiser;

/* One split decision: DeviceSplitCandidate + the weights of GPUExpandEntry
 * (src/tree/updater_gpu_common.cuh:52-83, src/tree/gpu_hist/expand_entry.cuh:15-34). */
typedef struct b200hist_candidate {
  float loss_chg;
  int32_t dir;              /* 0 = kLeftDir (missing goes left), 1 = kRightDir */
  int32_t findex;           /* -1 = no valid split */
  float fvalue;             /* split condition (a cut value, or -inf) */
  b200hist_gpair64 left_sum;
  b200hist_gpair64 right_sum;
  float base_weight, left_weight, right_weight;
  int32_t split_gbin;       /* global bin whose exclusive scan is the left sum; -1 if none */
} b200hist_candidate;

/* One tree node, carrying exactly what the reference records through RegTree::ExpandNode and
 * RegTree::Stat (src/tree/tree_model.cc:805-830, src/tree/updater_gpu_hist.cu:610-655,680-684). */
typedef struct b200hist_tree_node {
  int32_t parent, left, right;  /* -1 when absent */
  int32_t findex;               /* split feature; -1 for a leaf */
  float fvalue;                 /* split condition; for a leaf: the leaf value */
  int32_t default_left;
  int32_t is_leaf;
  float leaf_value;             /* learning_rate * weight (valid for leaves) */
  float loss_chg, sum_hess, base_weight; /* RTreeNodeStat */
  int32_t depth;
} b200hist_tree_node;

typedef struct b200hist_ellpack_s* b200hist_ellpack_t;  /* quantised feature matrix (EllpackPage) */
typedef struct b200hist_updater_s* b200hist_updater_t;  /* tree updater (GPUHistMaker) */
typedef struct b200hist_comm_s* b200hist_comm_t;        /* NCCL communicator of the histogram allreduce */
typedef struct b200hist_forest_s* b200hist_forest_t;    /* trees resolved against a page's cuts, for prediction */

/* ---------------------------------------------------------------------------------------------
 * Errors / info
 * ------------------------------------------------------------------------------------------- */
/* replaces XGBGetLastError (include/xgboost/c_api.h:90) for this library */
B200HIST_API const char* b200hist_last_error(void);
B200HIST_API int b200hist_version(void);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches` evidence). */
B200HIST_API uint64_t b200hist_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Ellpack page: HistogramCuts + quantised bin matrix
 *   replaces EllpackPageImpl (src/data/ellpack_page.cuh:181-391) and HistogramCuts
 *   (src/common/hist_util.h:38-190) as seen by the updater through
 *   DMatrix::GetBatches<EllpackPage> (src/tree/updater_gpu_hist.cu:735-756).
 * HBM layout (ours): row-major bins[n_rows][row_bytes]; one symbol per feature holding the
 * FEATURE-LOCAL bin (uint8 when every feature has <= 256 symbols incl. the null symbol, else
 * uint16); rows are zero-padded to whole 32-feature blocks (row_bytes = ceil32(n_features) *
 * symbol bytes: every row 32-byte aligned); the null symbol of a page with missing values is the
 * first symbol no feature uses (max bins per feature).  A column-major copy in the same symbol
 * width serves the partitioner's single-feature gathers.
 * ------------------------------------------------------------------------------------------- */
/* cut_ptrs_host: n_features+1 uint32 (HistogramCuts::cut_ptrs_), cut_values_host: total_bins f32.
 * has_missing: reserve a null symbol (dense-with-missing layout, ellpack_page.cu:140-183). */
B200HIST_API int b200hist_ellpack_create(int device, uint64_t n_rows, uint32_t n_features,
                                         const uint32_t* cut_ptrs_host, const float* cut_values_host,
                                         int has_missing, b200hist_ellpack_t* out);
/* External memory: the same dense page, but the quantised matrix lives in PINNED HOST memory and is streamed through
 * two device staging buffers of `page_rows` rows each (double-buffered H2D on a copy stream owned by the page set) by
 * every phase of the updater that reads it -- histogram build, partition, FinalisePosition -- once per tree level.
 *   replaces the Ellpack page sources and caches of ExtMemQuantileDMatrix (src/data/ellpack_page_source.h:37-337,
 *   src/data/extmem_quantile_dmatrix.cu) and the per-page loop of the updater (src/tree/updater_gpu_hist.cu:475-499,
 *   529-580, src/tree/gpu_hist/row_partitioner.cuh:449-501).
 * Device memory: 2 * page_rows * row_bytes for the pages + ~28 bytes per row of updater state; push_dense /
 * import_reference_rows work as for an in-core page (any row range, in any order).  Trees are bit-identical to the
 * in-core page's: histograms are integer sums, their value does not depend on how rows are grouped into pages. */
B200HIST_API int b200hist_ellpack_create_paged(int device, uint64_t n_rows, uint32_t n_features,
                                               const uint32_t* cut_ptrs_host, const float* cut_values_host,
                                               int has_missing, uint64_t page_rows, b200hist_ellpack_t* out);
/* n_pages = 0 for an in-core page; h2d_bytes = bytes streamed host -> device by the updater so far. */
B200HIST_API int b200hist_ellpack_paging_info(b200hist_ellpack_t ell, uint32_t* n_pages, uint64_t* page_rows,
                                              uint64_t* h2d_bytes);
/* Q2: quantile sketch -> HistogramCuts of a dense f32 batch resident on the device (NaN = missing), for the
 * in-core, unweighted, single-window case.  replaces DeviceSketch / SketchContainer::{Push, ScanInput, MakeCuts}
 * (src/common/hist_util.cu:234-397, src/common/quantile.cu:108-153,340-418,658-709, quantile.h:280-340).
 * cut_ptrs_out_host: n_features + 1; cut_values_out_host: capacity >= n_features * max_bin.  n < 2^24 rows. */
B200HIST_API int b200hist_sketch_dense(int device, const float* x_dev, uint64_t n, uint64_t ld, uint32_t n_features,
                                       int32_t max_bin, uint32_t* cut_ptrs_out_host, float* cut_values_out_host,
                                       uint64_t cut_values_capacity, void* stream);

/* Q2, batch by batch: the resident sketch behind QuantileDMatrix(DataIter).  replaces common::SketchContainer
 * (src/common/quantile.cuh:38-200) as driven by cuda_impl::MakeSketches (src/data/quantile_dmatrix.cu:29-126):
 *   create     = SketchContainer ctor (quantile.cuh:95-112), numerical features only;
 *   push_dense = one DataIter batch: AdapterDeviceSketch (src/common/hist_util.cuh:399-443; the batch is ONE sketch
 *                window, so n < 2^24 rows and n * n_features < 2^31) -> SketchContainer::Push = PruneImpl + ScanInput
 *                + Merge (MergeImpl, SegmentedUnique, FixError) + Prune(IntermediateNumCuts)
 *                (src/common/quantile.cu:340-378,491-566), then MakeSketches' Prune to
 *                RequiredSampleCutsPerColumn(max_bin, rows pushed so far) (quantile_dmatrix.cu:73-90);
 *   make_cuts  = SketchContainer::MakeCuts on one worker (quantile.cu:658-709).  fix_error_changes_out (optional)
 *                receives how many entries FixError had to repair over the sketch's life: the reference repairs in
 *                place from concurrent threads, so cut values are reproducible only while this stays 0. */
typedef struct b200hist_sketch_s* b200hist_sketch_t;
B200HIST_API int b200hist_sketch_create(int device, uint32_t n_features, int32_t max_bin, b200hist_sketch_t* out);
B200HIST_API int b200hist_sketch_push_dense(b200hist_sketch_t sketch, const float* x_dev, uint64_t n, uint64_t ld,
                                            void* stream);
/* Feature types of the sketch (SketchContainer's feature_types): a categorical column (flag != 0) gets the cut values
 * 0, 1, .., its largest category code instead of quantiles (quantile.cu:684-696, AddCategories); a negative code or
 * one beyond 2^24 is an error as in the reference (InvalidCat).  Host-side only: call any time before make_cuts. */
B200HIST_API int b200hist_sketch_set_feature_types(b200hist_sketch_t sketch, const uint8_t* feature_is_cat_host,
                                                   uint32_t n);
B200HIST_API int b200hist_sketch_make_cuts(b200hist_sketch_t sketch, uint32_t* cut_ptrs_out_host,
                                           float* cut_values_out_host, uint64_t cut_values_capacity,
                                           uint32_t* fix_error_changes_out, void* stream);
B200HIST_API int b200hist_sketch_destroy(b200hist_sketch_t sketch);

/* Sparse ELLPACK page (the reference's layout when row_stride < n_features, ellpack_page.cu:140-183):
 * uint32 symbols[n_rows][row_stride] holding GLOBAL bin indices in feature order, null = total_bins.
 * row_stride = the longest row (entries); the caller computes it like GetRowStride (ellpack_page.cuh:393+). */
B200HIST_API int b200hist_ellpack_create_sparse(int device, uint64_t n_rows, uint32_t n_features,
                                                const uint32_t* cut_ptrs_host, const float* cut_values_host,
                                                uint32_t row_stride, b200hist_ellpack_t* out);
/* Bin a CSR batch resident on the device into rows [row_begin, row_begin + n): indptr int64[n+1] (batch
 * relative), indices int32[nnz] ascending per row, values float32[nnz]; NaN values are skipped.
 * replaces CompressBinEllpackKernel<.., kDenseCompressed=false> (src/data/ellpack_page.cu:56-134). */
B200HIST_API int b200hist_ellpack_push_csr(b200hist_ellpack_t ell, uint64_t row_begin, const int64_t* indptr_dev,
                                           const int32_t* indices_dev, const float* values_dev, uint64_t n,
                                           void* stream);
B200HIST_API int b200hist_ellpack_destroy(b200hist_ellpack_t ell);
/* Bin a dense row-major f32 batch resident on the device into rows [row_begin, row_begin+n).
 * NaN = missing.  replaces CompressBinEllpackKernel / WriteCompressedEllpackFunctor
 * (src/data/ellpack_page.cu:56-134,240-298) + HistogramCuts::SearchBin (hist_util.h:112-121). */
B200HIST_API int b200hist_ellpack_push_dense(b200hist_ellpack_t ell, uint64_t row_begin,
                                             const float* x_dev, uint64_t n, uint64_t ld, void* stream);
/* Feature types of a dense-layout page: the categorical columns (flag != 0) of every later push_dense are binned by
 * HistogramCuts::SearchCatBin (src/common/hist_util.h:132-144: lower bound of the truncated value among the feature's
 * category codes, clamped to the last bin) instead of SearchBin.  Call before push_dense.
 * STATUS: green on a B200 since round 2 (tests/test_gpu_next.py, categorical cases; on by default). */
B200HIST_API int b200hist_ellpack_set_feature_types(b200hist_ellpack_t ell, const uint8_t* feature_is_cat_host,
                                                    uint32_t n);
/* Import a reference EllpackPage: decode its big-endian bit-packed gidx_buffer
 * (src/common/compressed_iterator.h:110-152,228-268) into our layout.  Dense layouts only
 * (row_stride == n_features); null_value as EllpackAccessor::NullValue (ellpack_page.cuh:85-87). */
B200HIST_API int b200hist_ellpack_import_reference(b200hist_ellpack_t ell, const uint8_t* gidx_buffer_dev,
                                                   uint32_t row_stride, uint32_t n_symbols, int is_dense,
                                                   void* stream);
/* The same for ONE page of a multi-page DMatrix (ExtMemQuantileDMatrix: GetBatches<EllpackPage> yields the pages one
 * at a time, src/data/extmem_quantile_dmatrix.cu): its n_rows rows land at [row_begin, row_begin + n_rows) of this
 * page, so that an external-memory matrix that fits HBM is trained from one resident page. */
B200HIST_API int b200hist_ellpack_import_reference_rows(b200hist_ellpack_t e, uint64_t row_begin, uint64_t n_rows,
                                                        const uint8_t* gidx_buffer_dev, uint32_t row_stride,
                                                        uint32_t n_symbols, int is_dense, void* stream);
/* Accessors (EllpackPageImpl::{Size, info.row_stride, NumSymbols}, ellpack_page.cuh:250-300). */
B200HIST_API int b200hist_ellpack_info(b200hist_ellpack_t ell, uint64_t* n_rows, uint32_t* n_features,
                                       uint32_t* total_bins, uint32_t* row_bytes, uint32_t* bin_bytes,
                                       uint32_t* null_symbol, const void** bins_dev);

/* Copy the bin matrix (n_rows * row_bytes bytes) to host memory; tests / debugging only. */
B200HIST_API int b200hist_ellpack_copy_bins_to_host(b200hist_ellpack_t ell, void* out_host);

/* ---------------------------------------------------------------------------------------------
 * Updater: replaces GPUHistMaker / GPUHistMakerDevice (src/tree/updater_gpu_hist.cu:104-891)
 * ------------------------------------------------------------------------------------------- */
/* TreeUpdater::Configure + InitDataOnce (updater_gpu_hist.cu:764-771,803-827).  The ellpack is
 * borrowed and must outlive the updater. */
B200HIST_API int b200hist_updater_create(b200hist_ellpack_t ell, const b200hist_train_param* param,
                                         b200hist_updater_t* out);
B200HIST_API int b200hist_updater_destroy(b200hist_updater_t u);
/* Upper bound on nodes a tree can have under the configured param (sizes out_nodes). */
B200HIST_API int b200hist_updater_max_nodes(b200hist_updater_t u, int32_t* out);

/* TreeUpdater::Update for one tree (updater_gpu_hist.cu:785-801 -> UpdateTree :693-732).
 * gpair_dev: n_rows x {grad, hess} float32 (GradientPair).  out_position_dev (nullable): n_rows
 * int32, leaf id per row, bitwise-NOT when the row's quantised hessian is 0
 * (updater_gpu_hist.cuh:48-56).  out_nodes_host: capacity >= max_nodes. */
B200HIST_API int b200hist_updater_update_tree(b200hist_updater_t u, const float* gpair_dev, void* stream,
                                              b200hist_tree_node* out_nodes_host, int32_t* out_n_nodes,
                                              int32_t* out_position_dev);
/* Same, with the gradient in HOST memory (XGBoosterTrainOneIter with host arrays,
 * include/xgboost/c_api.h:1080): H2D of the gradient and D2H of the tree happen inside. */
B200HIST_API int b200hist_updater_update_tree_host(b200hist_updater_t u, const float* gpair_host,
                                                   void* stream, b200hist_tree_node* out_nodes_host,
                                                   int32_t* out_n_nodes, int32_t* out_position_dev);
/* The same tree with a DEFERRED read-back: every kernel and the device -> host copy of the tree are enqueued on
 * `stream` and the call returns without waiting (the reference synchronises twice per level).  The prediction cache
 * (below) reads the device-side tree, so objective gradient -> update_tree_async -> update_prediction_cache of the
 * next iteration can be enqueued while this tree is still being built; b200hist_updater_fetch_tree returns the trees
 * in order (it waits for the oldest one in flight; at most two may be in flight). */
B200HIST_API int b200hist_updater_update_tree_async(b200hist_updater_t u, const float* gpair_dev, void* stream,
                                                    int32_t* out_position_dev);
B200HIST_API int b200hist_updater_fetch_tree(b200hist_updater_t u, b200hist_tree_node* out_nodes_host,
                                             int32_t* out_n_nodes);
/* TreeUpdater::UpdatePredictionCache (updater_gpu_hist.cu:582-608,848-869):
 * preds_dev[i] += leaf_value[decode(position[i])] for the last tree built. */
B200HIST_API int b200hist_updater_update_prediction_cache(b200hist_updater_t u, float* preds_dev,
                                                          void* stream);
/* The same with everything the reference's call can vary (updater_gpu_hist.cu:582-608):
 *   stride            elements between consecutive rows of the output column -- GBTree::DoBoost hands class `gid` of a
 *                     multi-class model as out.Slice(All, gid): Values() starts at column gid, Stride(0) = n_groups
 *                     (src/gbm/gbtree.cc:224-253);
 *   position_dev      the positions Update wrote (NULL = the buffer of the last Update call, which the caller must
 *                     then still own);
 *   leaf_values_host  leaf value per node id of the last tree (NULL = the values Update computed): adaptive
 *                     objectives rewrite RegTree's leaves between Update and UpdatePredictionCache
 *                     (ObjFunction::UpdateTreeLeaf, src/objective/adaptive.cu), the reference therefore reads
 *                     p_tree->GetNodes() here. */
B200HIST_API int b200hist_updater_update_prediction_cache_ex(b200hist_updater_t u, float* preds_dev, uint64_t stride,
                                                             const int32_t* position_dev,
                                                             const float* leaf_values_host, int32_t n_leaf_values,
                                                             void* stream);
/* ---------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU): feature-block sharded histogram + ONE int64 allreduce per tree level.
 *   replaces DeviceHistogramBuilder::AllReduceHist (src/tree/gpu_hist/histogram.cu:598-609) ->
 *   collective::Allreduce/NCCLColl (src/collective/aggregator.h:138-160, src/collective/coll.cu:251-277)
 *   and the communicator bootstrap (src/collective/comm.cu:45-101).
 * Every rank holds the full page and the full row state; rank r builds only the histogram slices of its
 * feature groups, the allreduce (a sum over disjoint slices, hence exact) completes them, and all ranks then
 * evaluate and partition identically.  Results are bit-identical to a single GPU.
 * ------------------------------------------------------------------------------------------- */
/* 128-byte NCCL unique id, created on one rank and distributed by the caller's control plane. */
B200HIST_API int b200hist_comm_unique_id(uint8_t* out_id128);
B200HIST_API int b200hist_comm_create(int device, int world, int rank, const uint8_t* id128, b200hist_comm_t* out);
B200HIST_API int b200hist_comm_destroy(b200hist_comm_t comm);
/* Rank `rank` of `world` owns feature groups [g_begin, g_end) out of n_groups (host logic only). */
B200HIST_API int b200hist_shard_groups(uint32_t n_groups, int world, int rank, uint32_t* g_begin, uint32_t* g_end);
/* Attach a communicator: the updater builds only its shard and allreduces once per level.  The
 * communicator is borrowed.  Passing NULL returns to single-GPU mode. */
B200HIST_API int b200hist_updater_set_comm(b200hist_updater_t u, b200hist_comm_t comm);
B200HIST_API int b200hist_updater_num_feature_groups(b200hist_updater_t u, uint32_t* out);
/* Peer-memory exchange (default multi-GPU collective; NVLink P2P instead of NCCL): every rank exports the IPC
 * handle of its histogram cache (64 bytes, cudaIpcMemHandle_t), the caller all-gathers the handles, and
 * b200hist_updater_set_peers maps the peers' caches.  Afterwards each level's built histogram slices are pushed
 * into every peer by one kernel that ends in a cross-GPU flag barrier; no NCCL call is made.
 * handles_host: world x 64 bytes in rank order (the own entry is ignored). */
B200HIST_API int b200hist_updater_ipc_handle(b200hist_updater_t u, uint8_t* out_handle64);
B200HIST_API int b200hist_updater_set_peers(b200hist_updater_t u, int rank, int world, const uint8_t* handles_host);
/* Unmap the peers' caches (cudaIpcCloseMemHandle) and return to single-GPU operation.  Freeing memory that another
 * process still maps is undefined, so a job tears down in two steps: every rank calls this, the ranks meet at a
 * barrier of the caller's control plane, then the updaters are destroyed. */
B200HIST_API int b200hist_updater_clear_peers(b200hist_updater_t u);

/* Per-tree sampling, applied to every following update_tree call until changed:
 *   feature_set_host (nullable = all features): the tree's colsample_bytree draw, strictly ascending feature ids --
 *     what ColumnSampler::GetFeatureSet(depth) returns when only colsample_bytree < 1 (src/common/random.h:101-132,
 *     updater_gpu_hist.cu:199-217,257-261); split evaluation visits exactly these features in this order.
 *   subsample < 1: uniform row sampling = cuda_impl::UniformSampling::Sample (src/tree/gpu_hist/sampler.cu:160-173):
 *     the quantised gradient pair of row i is zeroed when RandomWeight(subsample_seed)(i) > subsample
 *     (minstd_rand + discard(i) + uniform float, sampler.cu:36-60); subsample_seed is the tree's ctx->Rng()() draw;
 *     i is the row's index in THIS page (under row split: the rank-local index, as in the reference's workers).
 * colsample_bylevel / bynode and gradient_based sampling are not built. */
B200HIST_API int b200hist_updater_set_tree_sampling(b200hist_updater_t u, const int32_t* feature_set_host,
                                                    uint32_t n_feature_set, float subsample, uint32_t subsample_seed);

/* colsample_bylevel: one feature list per depth, sets_host[n_levels][n_per_level], each strictly ascending -- what
 * ColumnSampler::GetFeatureSet(depth) returns for depth = 0, 1, ... (src/common/random.h:156-175; every level draws
 * the same number of features from the tree's set); a node at depth d is evaluated on row min(d, n_levels - 1).
 * Replaces the tree-level set of b200hist_updater_set_tree_sampling (row sampling is kept); NULL / 0 removes it.
 * The reference draws a level's set lazily, when the first node of that depth is evaluated:
 * b200hist_updater_last_max_eval_depth tells the caller how many of the pre-drawn levels the last tree really used,
 * so that it can leave its random engine exactly where the reference would (the shim does). */
B200HIST_API int b200hist_updater_set_level_feature_sets(b200hist_updater_t u, const int32_t* sets_host,
                                                         uint32_t n_levels, uint32_t n_per_level);
B200HIST_API int b200hist_updater_last_max_eval_depth(b200hist_updater_t u, int32_t* out);

/* TrainParam::sampling_method: 0 = uniform (default), 1 = gradient_based = GradientBasedSampling::Sample
 * (src/tree/gpu_hist/sampler.cu:194-312): with subsample < 1 (b200hist_updater_set_tree_sampling; its seed is the one
 * ctx->Rng()() draw the reference makes per tree, sampler.cu:308) rows are kept with probability min(1, rag_i / mu),
 * rag_i = sqrt(g_i^2 + 0.1 h_i^2) on the quantised pair, mu chosen so that the expected sample size is
 * n * subsample; kept rows below mu are scaled by 1 / p.  Single GPU.
 * STATUS: green on a B200 since round 2: the reference's own gpu_hist trees under sampling_method=gradient_based are
 * committed as fixtures (tests/golden/ref_gpu_next_trees.json, tools/make_golden_gpu_next.py) and reproduced bit for bit
 * by this path and by the oracle; drop-in test through the shim in tests/test_gpu_next.py. */
B200HIST_API int b200hist_updater_set_sampling_method(b200hist_updater_t u, int gradient_based);

/* colsample_bynode = ColumnSampler::GetFeatureSet's per-call draw (src/common/random.h:156-175, random.cu:62-96).
 * gpu_hist asks for a set once for the root and then for the left and the right child of every valid candidate in
 * Pop order (updater_gpu_hist.cu:217,257-264): the draws of a tree form one stream, independent of the data.
 * sets_host holds the first n_sets >= max_nodes draws of that stream, n_per_set ascending feature ids each; the k-th
 * node evaluated in the tree is evaluated on draw k (position in the list = tie rank).  b200hist_updater_last_eval_count
 * tells the caller how many draws the last tree used, so that it can leave its random engine where the reference
 * would.  Replaces tree- and level-wise sets while set; NULL / 0 removes it.  Not combinable with interaction
 * constraints yet, nor with colsample_bylevel (the reference interleaves the lazy level draws with the node draws in
 * one engine, which makes the stream data dependent).
 * STATUS: green on a B200 since round 2 (tests/test_gpu_next.py). */
B200HIST_API int b200hist_updater_set_node_feature_sets(b200hist_updater_t u, const int32_t* sets_host, uint32_t n_sets,
                                                        uint32_t n_per_set);
B200HIST_API int b200hist_updater_last_eval_count(b200hist_updater_t u, int32_t* out);

/* The whole column sampler of a tree on the device (ColumnSampler::GetFeatureSet, src/common/random.h:156-175): tree_set =
 * what ColumnSampler::Init produced (colsample_bytree), seeds_host = the values ctx->Rng()() returns after Init, in
 * order.  A node of depth d is evaluated on: the depth's level set, drawn from tree_set with the next output the first
 * time a node of that depth is evaluated (colsample_bylevel < 1), and -- colsample_bynode < 1 -- a node set drawn from
 * the level set with the next output.  Calls happen in evaluation order, so the position of a level draw in the stream
 * depends on the data; the cursor therefore lives on the device.  n_seeds >= max_nodes + max depth + 1.
 * b200hist_updater_last_draw_count = outputs the last tree consumed: the caller advances its engine by that many.
 * Covers colsample_bynode with or without colsample_bylevel (and colsample_bylevel alone); replaces
 * b200hist_updater_set_node_feature_sets / set_level_feature_sets while set; NULL / 0 removes it.  With interaction
 * constraints a node's Query runs on its OWN sampled list (updater_gpu_hist.cu:257-264).
 * STATUS: green on a B200 since round 2: colsample_bylevel / bynode through the shim equal the reference's own trees
 * (tests/test_gpu_next.py, tests/golden/ref_gpu_next_trees.json); whole trees also under the CPU emulator. */
B200HIST_API int b200hist_updater_set_column_sampler(b200hist_updater_t u, const int32_t* tree_set_host,
                                                     uint32_t n_tree_set, float colsample_bylevel,
                                                     float colsample_bynode, const uint32_t* seeds_host,
                                                     uint32_t n_seeds);
B200HIST_API int b200hist_updater_last_draw_count(b200hist_updater_t u, int32_t* out);

/* Column-sampling draws, many at once: ColumnSampler::ColSample on CUDA without feature weights (src/common/random.cc:
 * 27-46; cuda_impl::SampleFeature, random.cu:62-96 = thrust::shuffle with default_random_engine(seed), keep the first
 * n_sel = max(1, int(colsample * n)), sort ascending).  Draw d uses seeds_host[d] -- the value the reference obtains
 * from ctx->Rng()() for that call.  out_host: n_draws x n_sel.  One launch instead of one host call per draw; n <= 8192.
 * STATUS: green on a B200 since round 2 (tests/test_gpu_next.py); the arithmetic is also pinned on the CPU against the toolkit's
 * own thrust::shuffle (tests/golden/thrust_shuffle_known_answers.json). */
B200HIST_API int b200hist_col_sample_stream(int device, const int32_t* features_host, uint32_t n, float colsample,
                                            const uint32_t* seeds_host, uint32_t n_draws, int32_t* out_host,
                                            uint32_t* n_sel_out, void* stream);

/* Monotone constraints = TreeEvaluator (src/tree/split_evaluator.h:25-205): constraints_host[f] in {-1, 0, +1} for the
 * first n features (the rest 0); n = 0 or NULL removes them.  A non-empty list switches every split gain to the
 * weight-based form -(2 g w + (h + lambda) w^2) with child weights clamped into the bounds of the node being split
 * (CalcSplitGain / CalcWeight :73-112), rejects splits whose child weights violate the sign, and propagates
 * mid = (w_left + w_right) / 2 as a bound to the children (AddSplit :163-203), exactly like the reference. */
B200HIST_API int b200hist_updater_set_monotone_constraints(b200hist_updater_t u, const int32_t* constraints_host,
                                                           uint32_t n);

/* Feature interaction constraints = FeatureInteractionConstraintDevice (src/tree/constraints.cu:25-329, used at
 * src/tree/updater_gpu_hist.cu:217-220,251-268,653-654).  The parsed `interaction_constraints` parameter as CSR:
 * set s holds features set_feat_host[set_ptr_host[s] .. set_ptr_host[s + 1]); n_sets = 0 removes them.
 * Every node carries the set of features it may split on: a split on f adds every set that holds f (and f itself)
 * to the node and hands the result to both children; the root is unrestricted.  One property of the reference is
 * reproduced on purpose because its trees depend on it (tests/golden/ref_gpu_interaction_trees.json): all Query
 * calls of a level answer into ONE buffer before the evaluation kernel runs, so a node is evaluated on the first
 * `its own count` entries of what the level's later answers left there.  With disjoint sets every node of a tree has
 * the same answer and the two readings coincide.
 * STATUS: green on a B200 since round 2: the reference's own trees bit for bit (overlapping sets, lossguide, 130 features,
 * with colsample_bylevel / monotone constraints) and the drop-in through the shim (tests/test_gpu_next.py); on by default. */
B200HIST_API int b200hist_updater_set_interaction_constraints(b200hist_updater_t u, const int32_t* set_ptr_host,
                                                              const int32_t* set_feat_host, uint32_t n_sets);

/* Categorical features in the tree updater.  feature_is_cat_host[f] != 0 for the first n
 * features marks a categorical feature (MetaInfo::feature_types == kCategorical); its bins are its categories and
 * its cut values the category codes, as the reference's sketch emits them (AddCategories) -- checked here.
 * max_cat_to_onehot / max_cat_threshold are TrainParam's (src/tree/param.h; defaults 4 and 64).  All zero / NULL
 * removes them.  Evaluation: see b200hist_evaluate_splits_cat.  An applied categorical split stores NaN as its
 * split condition (updater_gpu_hist.cu:628-642); rows whose category is in the node's set go RIGHT, missing values
 * follow default_left (common::Decision, src/common/categorical.h:49-60) in the partitioner and in FinalisePosition.
 * b200hist_updater_last_tree_categories: per node of the last tree, 1 if it is a categorical split, and its
 * category set as cat_words uint32 words in the reference's LBitField32 layout (category c = bit 31 - c % 32 of
 * word c / 32) -- what RegTree::ExpandCategorical takes.  Buffers: n_nodes and n_nodes * cat_words entries.
 * STATUS: green on a B200 since round 2: one-hot and sorted-partition splits, the reference's own trees bit for bit and the
 * drop-in through the shim (tests/test_gpu_next.py); on by default. */
B200HIST_API int b200hist_updater_set_categorical(b200hist_updater_t u, const uint8_t* feature_is_cat_host, uint32_t n,
                                                  int32_t max_cat_to_onehot, int32_t max_cat_threshold);
B200HIST_API int b200hist_updater_cat_words(b200hist_updater_t u, uint32_t* out);
B200HIST_API int b200hist_updater_last_tree_categories(b200hist_updater_t u, int32_t* node_is_cat_host,
                                                       uint32_t* cat_bits_host);

/* Row split (the reference's default distributed mode): every rank holds a page of ITS rows with the
 * same cuts and builds full-width partial histograms; one peer-memory kernel per level sums them (reduce-scatter +
 * all-gather over NVLink, integer, bit-identical to one GPU) = AllReduceHist (src/tree/gpu_hist/histogram.cu:598-609);
 * the quantiser scale and the root sum come from rank-ordered scalar exchanges = the GlobalSum calls of
 * GradientQuantiserGroup (src/tree/gpu_hist/quantiser.cu:88-123) and InitRoot (updater_gpu_hist.cu:661-668); the child
 * to build is chosen by hessian sums like the reference (updater_gpu_hist.cu:461-467) so that all ranks agree.
 * Call after b200hist_updater_set_peers; on = 0 returns to feature-block sharding. */
B200HIST_API int b200hist_updater_set_row_split(b200hist_updater_t u, int on);

/* Introspection of the last update (parity tests): quantiser, per-node histogram, per-node
 * candidate, per-node row counts. */
B200HIST_API int b200hist_updater_last_quantiser(b200hist_updater_t u, b200hist_quantiser* out_host);
B200HIST_API int b200hist_updater_node_histogram(b200hist_updater_t u, int32_t nidx,
                                                 b200hist_gpair64* out_host /* total_bins */);
B200HIST_API int b200hist_updater_node_candidate(b200hist_updater_t u, int32_t nidx,
                                                 b200hist_candidate* out_host);
/* (d) measurement: the shared-memory atomic ceiling the histogram kernel (H2) is judged against, MEASURED on this
 * device: the kernel's own accumulation sequence (conflict-free 32-bit shared atomics with carry) on
 * register-generated symbols, no global loads, same CTA shape; returns (row, feature) updates per clock per SM.
 * h_lo_zero != 0 measures the variant that skips the hessian's low word.  No reference counterpart. */
B200HIST_API int b200hist_smem_atomic_peak(int device, int h_lo_zero, double* updates_per_clk_per_sm, void* stream);

/* Per-phase device time (ms) of the last update_tree when profiling is on: quantise, hist,
 * subtract, evaluate, apply, partition, finalise (7 floats) and row-visits V of the hist phase. */
B200HIST_API int b200hist_updater_set_profiling(b200hist_updater_t u, int on);
B200HIST_API int b200hist_updater_last_profile(b200hist_updater_t u, float* out_ms7, uint64_t* out_row_visits);
/* Device time of the histogram phase of every growth step (level) of the last profiled tree, in launch order
 * (root first); *out_n = number of steps, at most `capacity` values are written. */
B200HIST_API int b200hist_updater_last_hist_steps(b200hist_updater_t u, float* out_ms, int32_t capacity,
                                                  int32_t* out_n);
/* with profiling on: device time of the last tree's cross-GPU histogram exchanges (already contained in the
 * "hist" phase of b200hist_updater_last_profile); 0 on one GPU. */
B200HIST_API int b200hist_updater_last_comm_ms(b200hist_updater_t u, float* out_ms);

/* ---------------------------------------------------------------------------------------------
 * Component entry points (each a row of SURVEY.md 8(a)); used by the parity tests.
 * ------------------------------------------------------------------------------------------- */
/* G1+G2+G3: GradientQuantiserGroup + CalcQuantizedGpairs + root sum
 * (src/tree/gpu_hist/quantiser.cu:55-149, updater_gpu_hist.cu:661-664). */
B200HIST_API int b200hist_quantise(int device, const float* gpair_dev, uint64_t n,
                                   b200hist_quantiser* out_host, b200hist_gpair64* gq_dev,
                                   b200hist_gpair64* root_sum_host, void* stream);
/* H2: DeviceHistogramBuilder::BuildHistogram (src/tree/gpu_hist/histogram.cu:570-582) for a batch
 * of nodes in one launch.  Node k owns rows ridx_dev[seg_begin[k] .. +seg_count[k]); its histogram
 * is hist_dev + k*total_bins (cleared by the call). */
B200HIST_API int b200hist_build_hist(b200hist_ellpack_t ell, const b200hist_gpair64* gq_dev,
                                     const uint32_t* ridx_dev, const uint64_t* seg_begin_host,
                                     const uint64_t* seg_count_host, int32_t n_nodes,
                                     b200hist_gpair64* hist_dev, void* stream);
/* H4: SubtractionTrick (src/tree/gpu_hist/histogram.cuh:208-221): out = parent - built. */
B200HIST_API int b200hist_subtract(int device, const b200hist_gpair64* parent_dev,
                                   const b200hist_gpair64* built_dev, b200hist_gpair64* out_dev,
                                   uint64_t n_bins, void* stream);
/* E1-E4: GPUHistEvaluator::EvaluateSplits (src/tree/gpu_hist/evaluate_splits.cu:400-470).
 * hist_dev: n_nodes x total_bins; parent_sums_host: n_nodes; feature_set_host (nullable): the
 * sampled feature ids, shared by all nodes. */
B200HIST_API int b200hist_evaluate_splits(b200hist_ellpack_t ell, const b200hist_gpair64* hist_dev,
                                          const b200hist_gpair64* parent_sums_host, int32_t n_nodes,
                                          const int32_t* feature_set_host, int32_t n_feature_set,
                                          const b200hist_train_param* param,
                                          const b200hist_quantiser* quantiser,
                                          b200hist_candidate* out_host, void* stream);
/* E1-E4 with categorical features: EvaluateSplitAgent::{OneHot, Partition}
 * (src/tree/gpu_hist/evaluate_splits.cu:148-252), GPUHistEvaluator::SortHistogram (evaluator.cu:69-128),
 * DeviceSplitCandidate::UpdateCat (updater_gpu_common.cuh:88-104), SetCategoricalSplit (evaluate_splits.cu:313-348).
 * feature_is_cat_host[f] != 0 marks a categorical feature, whose bins are its categories (cut value = category
 * code).  A feature with fewer than max_cat_to_onehot bins is split one category against the rest; otherwise the
 * categories are scanned in the order of their leaf weights from both ends, max_cat_threshold deep.  Per node:
 * out_host (fvalue = NaN, split_gbin = -1 for a categorical winner), out_is_cat_host, and the categories that go
 * RIGHT as the reference's LBitField32 over category codes in out_cat_bits_host[node * n_cat_words ...]
 * (category c = bit 31 - c % 32 of word c / 32).
 * STATUS: green on a B200 since round 2: the reference's twelve literal vectors and random histograms against the oracle
 * (tests/test_gpu_next.py). */
B200HIST_API int b200hist_evaluate_splits_cat(b200hist_ellpack_t ell, const b200hist_gpair64* hist_dev,
                                              const b200hist_gpair64* parent_sums_host, int32_t n_nodes,
                                              const int32_t* feature_set_host, int32_t n_feature_set,
                                              const b200hist_train_param* param,
                                              const b200hist_quantiser* quantiser,
                                              const uint8_t* feature_is_cat_host, int32_t max_cat_to_onehot,
                                              int32_t max_cat_threshold, b200hist_candidate* out_host,
                                              int32_t* out_is_cat_host, uint32_t* out_cat_bits_host,
                                              uint32_t n_cat_words, void* stream);
/* P2: RowPartitioner::UpdatePositionBatch (src/tree/gpu_hist/row_partitioner.cuh:343-418) with the
 * numerical GoLeftOp (updater_gpu_hist.cu:398-419).  Segment k of ridx_in_dev is split stably into
 * ridx_out_dev (left rows first); left_count_host[k] receives the number of left rows. */
B200HIST_API int b200hist_partition(b200hist_ellpack_t ell, const uint32_t* ridx_in_dev,
                                    uint32_t* ridx_out_dev, const uint64_t* seg_begin_host,
                                    const uint64_t* seg_count_host, const int32_t* findex_host,
                                    const float* split_cond_host, const int32_t* default_left_host,
                                    int32_t n_nodes, uint64_t* left_count_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Prediction on the quantised matrix
 *   replaces GPUPredictor's PredictKernel / PredictBatch over Ellpack pages
 *   (src/predictor/gpu_predictor.cu:180-230,316-351; EllpackLoader::GetElement,
 *   src/predictor/gpu_data_accessor.cuh:95-107: a bin stands for its lower cut, -inf for a feature's first bin).
 *   A row is routed by its bin alone: add_tree resolves every condition against the page's cuts (threshold = 1 +
 *   number of the feature's cuts < the condition).  For trees grown on these cuts that is the updater's own
 *   threshold; for any other forest (a loaded model, a page with its own cuts) it is still the reference's result
 *   on that page.
 * forest_create borrows `ell` for its cuts (keep it alive); add_tree takes a tree as update_tree returned it, with the
 * categorical flags / LBitField32 sets of b200hist_updater_last_tree_categories (NULL, NULL, 0 without categorical
 * splits).  predict_margin: for rows [row_begin, row_begin + n_rows) of `page` (any page quantised with the same
 * cuts, e.g. a validation set), out_dev[k] += sum over trees [tree_begin, tree_end) of the leaf the row reaches, the
 * sum taken in tree order from 0.0f like the reference (gpu_predictor.cu:221-228).  predict_leaf: out_dev[k *
 * n_trees + t] = the leaf's node id inside tree t (:196-201).
 * STATUS: green on a B200 since round 2 (tests/test_gpu_next.py: margins and leaf ids against the oracle, which reproduces
 * the reference CPU predictor's margins bit for bit on the training cuts and on a QuantileDMatrix with foreign cuts,
 * tests/test_reference_golden.py); the kernels also equal the oracle under the CPU emulator. */
B200HIST_API int b200hist_forest_create(b200hist_ellpack_t ell, b200hist_forest_t* out);
B200HIST_API int b200hist_forest_destroy(b200hist_forest_t forest);
B200HIST_API int b200hist_forest_num_trees(b200hist_forest_t forest, uint32_t* out);
B200HIST_API int b200hist_forest_add_tree(b200hist_forest_t forest, const b200hist_tree_node* nodes_host,
                                          int32_t n_nodes, const int32_t* node_is_cat_host,
                                          const uint32_t* cat_bits_host, uint32_t cat_words);
B200HIST_API int b200hist_forest_predict_margin(b200hist_forest_t forest, b200hist_ellpack_t page,
                                                uint32_t tree_begin, uint32_t tree_end, uint64_t row_begin,
                                                uint64_t n_rows, float* out_dev, void* stream);
B200HIST_API int b200hist_forest_predict_leaf(b200hist_forest_t forest, b200hist_ellpack_t page, uint32_t tree_begin,
                                              uint32_t tree_end, uint64_t row_begin, uint64_t n_rows,
                                              int32_t* out_dev, void* stream);
/* Several output groups (multi-class: tree t feeds output groups_host[t]; gbm::GBTreeModel::tree_info).  With more
 * than one group the reference adds every leaf straight onto the output, tree after tree
 * (gpu_predictor.cu:229-239), instead of summing from 0.0f first: out_dev[(k * n_groups) + g] is read once, the
 * leaves of the group's trees in [tree_begin, tree_end) are added in tree order, and it is written once.
 * set_tree_groups replaces the group of EVERY tree added so far (n_trees must equal forest_num_trees; trees added
 * later start in group 0).  predict_leaf_f32 is predict_leaf with the id stored as a float, the layout
 * Predictor::PredictLeaf returns (gpu_predictor.cu:793-830).
 * STATUS: green on a B200 since round 2 (multi-class margins and leaf ids through the Predictor binding,
 * tests/test_gpu_next.py::test_predictor_binding_serves_quantised_matrices_like_the_reference). */
B200HIST_API int b200hist_forest_set_tree_groups(b200hist_forest_t forest, const uint32_t* groups_host,
                                                 uint32_t n_trees);
B200HIST_API int b200hist_forest_predict_margin_groups(b200hist_forest_t forest, b200hist_ellpack_t page,
                                                       uint32_t tree_begin, uint32_t tree_end, uint32_t n_groups,
                                                       uint64_t row_begin, uint64_t n_rows, float* out_dev,
                                                       void* stream);
/* The same three results straight from a page of the REFERENCE, without importing it: gidx_buffer_dev is
 * EllpackPageImpl::gidx_buffer (8 bytes of padding, then the MSB-first bit stream of ceil(log2(n_symbols))-bit
 * symbols, src/common/compressed_iterator.h:37-40,228-268); dense_layout = EllpackPageImpl::IsDenseCompressed
 * (feature-local symbols, row_stride == n_features), is_dense = no missing values (null = n_symbols - 1 otherwise);
 * a sparse page holds ascending global bins and is searched per row like EllpackAccessorImpl::GetBinIndex
 * (src/data/ellpack_page.cuh:101-122).  The page must be quantised with the cuts the forest was created on.
 * mode 0: out_dev[n_rows] += margin (sum from 0.0f first); 1: out_dev[n_rows * n_groups], leaves straight onto the
 * output; 2: out_dev[n_rows * n_trees] = leaf ids as floats.  This is what the Predictor binding calls.
 * STATUS: green on a B200 since round 2: what the shim's Predictor serves by default -- bit-identical to the reference's
 * GPUPredictor, margins 1.26x its speed on 10M x 64 x 50 trees (profiles/r02_predictor_bench.json). */
B200HIST_API int b200hist_forest_predict_reference_page(b200hist_forest_t forest, const uint8_t* gidx_buffer_dev,
                                                        uint64_t n_rows, uint32_t row_stride, uint32_t n_symbols,
                                                        int dense_layout, int is_dense, int mode,
                                                        uint32_t tree_begin, uint32_t tree_end, uint32_t n_groups,
                                                        float* out_dev, void* stream);
B200HIST_API int b200hist_forest_predict_leaf_f32(b200hist_forest_t forest, b200hist_ellpack_t page,
                                                  uint32_t tree_begin, uint32_t tree_end, uint64_t row_begin,
                                                  uint64_t n_rows, float* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200HIST_H_ */
