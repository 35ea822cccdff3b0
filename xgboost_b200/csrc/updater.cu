// Tree-level driver of the B200 hist updater and the updater part of the C ABI.
//
// Replaces (reference paths relative to /root/reference):
//   GPUHistMakerDevice::{Reset, InitRoot, UpdateTree, ApplySplit, PartitionAndBuildHist,
//     EvaluateSplits, FinalisePosition, UpdatePredictionCache}   src/tree/updater_gpu_hist.cu:172-732
//   Driver<GPUExpandEntry> (Push/Pop/IsChildValid)                src/tree/driver.h:29-107
//   GPUExpandEntry::IsValid                                       src/tree/gpu_hist/expand_entry.cuh:33-50
//   AssignNodes                                                   src/tree/updater_gpu_hist.cuh:59-81
//   RegTree::ExpandNode bookkeeping                               src/tree/tree_model.cc:805-830
//
// B200-first organisation: the reference keeps the tree, the expand queue and the row segments on
// the HOST and synchronises the stream twice per level.  Here the whole growth state (nodes,
// candidates, row segments, work lists, tile tables) lives in device memory and is advanced by small
// single-CTA "control" kernels (ApplyKernel, ScanKernel), so one tree is a fixed, data-independent
// sequence of launches with no host read-back until the finished tree is copied out once.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace b200 {

constexpr int kCtlThreads = 1024;

// ---------------------------------------------------------------------------------------------
// Block-wide exclusive scan for the control kernels (kCtlThreads threads).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t BlockExclusiveScan(uint32_t v, uint32_t* s_warp /* 33 words */, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect s_warp reuse
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0u;  // blocks smaller than 32 warps
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + inc - v;
}

// Segment lookup: largest k with tile_begin[k] <= t (tile_begin ascending, n >= 1)
__device__ __forceinline__ uint32_t FindSegment(const uint32_t* tile_begin, uint32_t n, uint32_t t) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (tile_begin[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ double CalcWeightDev(const ParamDev& p, double g, double h) {
  if (h < p.min_child_weight || h <= 0.0) return 0.0;
  double t = g > +p.reg_alpha ? g - p.reg_alpha : (g < -p.reg_alpha ? g + p.reg_alpha : 0.0);
  double dw = -t / (h + p.reg_lambda);
  if (p.max_delta_step != 0.0f && fabs(dw) > p.max_delta_step) dw = copysign((double)p.max_delta_step, dw);
  return dw;
}

__device__ __forceinline__ void InitLeaf(NodeDev* n, int parent, int depth, float leaf, float sum_hess, long long g,
                                         long long h) {
  n->parent = parent; n->left = -1; n->right = -1; n->depth = depth;
  n->findex = -1; n->fvalue = leaf; n->default_left = 0; n->is_leaf = 1;
  n->leaf_value = leaf; n->loss_chg = 0.0f; n->sum_hess = sum_hess; n->base_weight = leaf;
  n->sum_g = g; n->sum_h = h;
  n->seg_begin = 0; n->seg_count = 0; n->seg_buf = -1; n->partitioned = 0; n->split_lbin = 0; n->cand_open = 0;
  n->hist_slot = -1; n->pad = 0;
}

// ---------------------------------------------------------------------------------------------
// InitRoot (updater_gpu_hist.cu:657-691) minus the histogram build / evaluation launches.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtlThreads) InitRootKernel(ParamDev p, uint32_t n_rows, const QuantDev* q,
                                                               TreeState* ts, NodeDev* nodes, BuildEntry* build,
                                                               uint32_t* build_tile_seg, int32_t* eval_nodes,
                                                               MonoView mono, int32_t* slot_owner) {
  const uint32_t tiles = (n_rows + kHistTile - 1) / kHistTile;
  for (uint32_t t = threadIdx.x; t < tiles; t += kCtlThreads) build_tile_seg[t] = 0;
  if (threadIdx.x == 0) {
    double rg = (double)q->root_g * q->to_float_g, rh = (double)q->root_h * q->to_float_h;
    float w = (float)CalcWeightDev(p, rg, rh);
    NodeDev* r = &nodes[0];
    InitLeaf(r, -1, 0, p.learning_rate * w, (float)rh, q->root_g, q->root_h);
    r->base_weight = w;
    r->seg_begin = 0; r->seg_count = n_rows; r->seg_buf = 0;
    r->hist_slot = 0;
    slot_owner[0] = 0;  // slot 0 is the root's for the whole tree (outside the ring)
    build[0] = BuildEntry{0, 0, 0u, n_rows, 0u, 0u};
    eval_nodes[0] = 0;
    if (mono.constraints) { mono.lower[0] = -3.402823466e+38f; mono.upper[0] = 3.402823466e+38f; }
    ts->n_nodes = 1; ts->num_leaves = 1; ts->level_begin = 0; ts->level_end = 1;
    ts->level_cursor = 0; ts->next_begin = 0;  // the first step opens level [next_begin, n_nodes) = [0, 1)
    ts->n_part = 0; ts->n_build = 1; ts->n_sub = 0; ts->n_eval = 1;
    ts->part_tiles = 0; ts->build_tiles = tiles; ts->final_tiles = 0; ts->n_final = 0;
    ts->row_visits = n_rows; ts->stopped = 0; ts->max_eval_depth = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Driver::Pop + IsValid + ApplySplit + IsChildValid + partition work list for one growth step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool EntryValidNoLeafLimit(const ParamDev& p, const b200hist_candidate& c, int depth) {
  if (c.loss_chg <= 1e-6f) return false;                          // kRtEps
  if (c.left_sum.hess == 0 || c.right_sum.hess == 0) return false;
  if (c.loss_chg < p.min_split_loss) return false;
  if (p.max_depth > 0 && depth == p.max_depth) return false;
  return true;
}

template <bool kCat>
__device__ void ApplyOne(const EllView& ell, const ParamDev& p, const QuantDev* q, NodeDev* nodes,
                         const b200hist_candidate& c, int nidx, int left, PartEntry* part_slot, const MonoView& mono,
                         const uint8_t* node_is_cat) {
  NodeDev* n = &nodes[nidx];
  const int right = left + 1;
  const double tfh = q->to_float_h;
  n->left = left; n->right = right;
  n->findex = c.findex; n->fvalue = c.fvalue; n->default_left = (c.dir == 0); n->is_leaf = 0;
  n->loss_chg = c.loss_chg;
  n->sum_hess = (float)((double)(c.left_sum.hess + c.right_sum.hess) * tfh);
  n->base_weight = c.base_weight;
  const float lw = c.left_weight * p.learning_rate, rw = c.right_weight * p.learning_rate;
  InitLeaf(&nodes[left], nidx, n->depth + 1, lw, (float)((double)c.left_sum.hess * tfh), c.left_sum.grad, c.left_sum.hess);
  InitLeaf(&nodes[right], nidx, n->depth + 1, rw, (float)((double)c.right_sum.hess * tfh), c.right_sum.grad, c.right_sum.hess);
  if (mono.constraints) {  // TreeEvaluator::AddSplit (split_evaluator.h:163-203): weights before the learning rate
    mono.lower[left] = mono.lower[right] = mono.lower[nidx];
    mono.upper[left] = mono.upper[right] = mono.upper[nidx];
    const int32_t cst = mono.constraints[c.findex];
    const float mid = (c.left_weight + c.right_weight) / 2;
    if (cst < 0) { mono.lower[left] = mid; mono.upper[right] = mid; }
    else if (cst > 0) { mono.upper[left] = mid; mono.lower[right] = mid; }
  }
  // go left iff cut_values[bin] <= fvalue  <=>  feature-local bin < (number of cuts <= fvalue)
  const uint32_t beg = ell.cut_ptrs[c.findex], end = ell.cut_ptrs[c.findex + 1];
  uint32_t lo = beg, hi = end;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (ell.cut_values[mid] <= c.fvalue) lo = mid + 1; else hi = mid;
  }
  n->split_lbin = (int32_t)(lo - beg);
  if constexpr (kCat) {
    // categorical split (tree.ExpandCategorical, updater_gpu_hist.cu:628-642): the rows are routed by the node's
    // category set (CatView::bin_bits, written by CatFinishKernel), -1 marks it (GoLeftBin)
    if (node_is_cat[nidx]) n->split_lbin = -1;
  }
  PartEntry e;
  e.nidx = nidx; e.left = left; e.right = right; e.findex = c.findex; e.split_lbin = n->split_lbin;
  e.default_left = n->default_left; e.begin = n->seg_begin; e.count = n->seg_count; e.tile_begin = 0;
  e.src_buf = n->seg_buf; e.n_left = 0; e.pad = 0;
  *part_slot = e;
}

template <bool kCat>
__global__ void __launch_bounds__(kCtlThreads) ApplyKernel(EllView ell, ParamDev p, const QuantDev* q, TreeState* ts,
                                                            NodeDev* nodes, const b200hist_candidate* cand,
                                                            PartEntry* part, uint32_t* part_tile_seg,
                                                            int32_t max_nodes, MonoView mono,
                                                            const uint8_t* node_is_cat, int level_start,
                                                            int batch_cap) {
  // level_start (depth-wise): this step opens the next level; otherwise it continues the open level where the
  // previous step stopped.  batch_cap: valid entries one step pops at most (Driver's max_node_batch_size).
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_tile_begin[kMaxLevelNodes + 1];
  __shared__ int s_n_expand, s_child_valid, s_cutoff;
  __shared__ float s_best_loss[kCtlThreads / 32];
  __shared__ int s_best_idx[kCtlThreads / 32];
  const int n_nodes0 = ts->n_nodes, leaves0 = ts->num_leaves;
  const int allow = p.max_leaves > 0 ? max(0, p.max_leaves - leaves0) : 0x7fffffff;
  int n_expand = 0;
  int level_depth = 0;

  if (ts->stopped) {
    if (threadIdx.x == 0) { ts->n_part = 0; ts->n_build = 0; ts->n_sub = 0; ts->n_eval = 0; ts->part_tiles = 0; ts->build_tiles = 0; }
    return;
  }

  if (p.grow_policy == 1) {
    // loss guide: pop the single best open entry (ties -> smaller nidx); an invalid top ends growth
    float bl = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_nodes0; i += kCtlThreads)
      if (nodes[i].cand_open) {
        float l = cand[i].loss_chg;
        if (l > bl || (l == bl && i < bi)) { bl = l; bi = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ol = __shfl_xor_sync(0xffffffffu, bl, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ol > bl || (ol == bl && oi < bi))) { bl = ol; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_best_loss[threadIdx.x >> 5] = bl; s_best_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kCtlThreads / 32; ++w) {
        float ol = s_best_loss[w];
        int oi = s_best_idx[w];
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ol > bl || (ol == bl && oi < bi))) { bl = ol; bi = oi; }
      }
      int ne = 0;
      if (bi == 0x7fffffff) {
        ts->stopped = 1;
      } else {
        nodes[bi].cand_open = 0;
        b200hist_candidate c = cand[bi];
        bool ok = EntryValidNoLeafLimit(p, c, nodes[bi].depth) && !(p.max_leaves > 0 && leaves0 == p.max_leaves) &&
                  n_nodes0 + 2 <= max_nodes;
        if (ok) {
          ApplyOne<kCat>(ell, p, q, nodes, c, bi, n_nodes0, &part[0], mono, node_is_cat);
          ne = 1;
        } else {
          ts->stopped = 1;
        }
      }
      s_n_expand = ne;
      level_depth = ne ? nodes[bi].depth : 0;
      int leaves_new = leaves0 + ne;
      s_child_valid = ne && (p.max_depth == 0 || level_depth + 1 < p.max_depth) &&
                      (p.max_leaves == 0 || leaves_new < p.max_leaves);
    }
    __syncthreads();
    n_expand = s_n_expand;
  } else {
    // depth wise: the open entries of the level in ascending nidx, at most batch_cap VALID ones per step
    // (Driver::Pop, driver.h:84-99: invalid entries are popped and dropped without counting)
    if (threadIdx.x == 0) {
      if (level_start) {
        ts->level_begin = ts->next_begin;
        ts->level_end = n_nodes0;
        ts->level_cursor = ts->next_begin;
        ts->next_begin = n_nodes0;
      }
      s_cutoff = ts->level_end;
    }
    __syncthreads();
    const int lb = ts->level_cursor, le = ts->level_end;
    const int cap = min(batch_cap, kMaxLevelNodes);
    level_depth = lb < le ? nodes[lb].depth : 0;
    uint32_t running = 0;
    int base = lb;
    for (; base < le && (int)running < cap; base += kCtlThreads) {
      const int i = base + threadIdx.x;
      bool valid = false;
      b200hist_candidate c;
      if (i < le && nodes[i].cand_open) {
        c = cand[i];
        valid = EntryValidNoLeafLimit(p, c, nodes[i].depth);
      }
      uint32_t total;
      const uint32_t rank = running + BlockExclusiveScan(valid ? 1u : 0u, s_warp, &total);
      // consumed by this step: everything in front of the (cap + 1)-th valid entry
      const bool consumed = i < le && (int)rank < cap;
      if (i < le && !consumed) atomicMin(&s_cutoff, i);
      if (consumed) nodes[i].cand_open = 0;
      bool accepted = consumed && valid && (int)rank < allow && n_nodes0 + 2 * ((int)rank + 1) <= max_nodes;
      if (accepted) ApplyOne<kCat>(ell, p, q, nodes, c, i, n_nodes0 + 2 * (int)rank, &part[rank], mono, node_is_cat);
      running += total;
    }
    // the step stopped in front of `base` (batch full at a chunk boundary) or inside the last chunk scanned
    if (threadIdx.x == 0) atomicMin(&s_cutoff, min(le, base));
    running = min(running, (uint32_t)cap);
    n_expand = min((int)running, allow);
    n_expand = min(n_expand, (max_nodes - n_nodes0) / 2);
    n_expand = min(n_expand, kMaxLevelNodes);
    if (threadIdx.x == 0) {
      int leaves_new = leaves0 + n_expand;
      s_child_valid = n_expand > 0 && (p.max_depth == 0 || level_depth + 1 < p.max_depth) &&
                      (p.max_leaves == 0 || leaves_new < p.max_leaves);
    }
    __syncthreads();
  }

  const int n_part = s_child_valid ? n_expand : 0;
  // partition tile table
  uint32_t running_tiles = 0;
  for (int base = 0; base < n_part; base += kCtlThreads) {
    const int k = base + threadIdx.x;
    uint32_t tiles = 0;
    if (k < n_part) {
      tiles = (part[k].count + kPartTile - 1) / kPartTile;
      nodes[part[k].nidx].partitioned = 1;
    }
    uint32_t total;
    uint32_t ex = running_tiles + BlockExclusiveScan(tiles, s_warp, &total);
    if (k < n_part) {
      part[k].tile_begin = ex;
      s_tile_begin[k] = ex;
    }
    running_tiles += total;
  }
  __syncthreads();
  if (n_part > 0)
    for (uint32_t t = threadIdx.x; t < running_tiles; t += kCtlThreads)
      part_tile_seg[t] = FindSegment(s_tile_begin, (uint32_t)n_part, t);
  if (threadIdx.x == 0) {
    ts->n_nodes = n_nodes0 + 2 * n_expand;
    ts->num_leaves = leaves0 + n_expand;
    if (p.grow_policy == 0) ts->level_cursor = s_cutoff;  // what is left of the level for the next step
    ts->n_part = n_part;
    ts->part_tiles = running_tiles;
    ts->n_build = 0; ts->n_sub = 0; ts->n_eval = 0; ts->build_tiles = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// After the count pass: per-tile offsets, children's row segments, AssignNodes (build the child
// with fewer rows, subtract the other), evaluation list and the histogram tile table.
//
// Histogram slots (DeviceHistogramStorage, histogram.cuh:72-168).  The step owns the slot window
// [slot_base, slot_base + 2 * slot_cap); the cache behind it may be a RING that later steps reuse (the host plans the
// windows), so every slot has an owner: a node's histogram is usable iff slot_owner[hist_slot] still names the node.
// The step first takes its window (whoever owned those slots loses them), then for every split:
//   parent's histogram still cached  -> build the smaller child, sibling = parent - built (SubtractHist);
//   parent's histogram gone          -> build BOTH children (the reference's slow path, updater_gpu_hist.cu:309-325).
// Built histograms take the first n_build slots of the window -- one contiguous range for the level's exchange --,
// subtracted ones follow.  With a cache that holds every node (the default) nothing is ever lost.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtlThreads) ScanKernel(TreeState* ts, NodeDev* nodes, PartEntry* part,
                                                           const uint32_t* part_tile_seg, const uint32_t* tile_left,
                                                           uint32_t* tile_excl /* scratch, part_tiles+1 */,
                                                           uint32_t* tile_left_off, BuildEntry* build, SubEntry* sub,
                                                           int32_t* eval_nodes, uint32_t* build_tile_seg,
                                                           uint32_t slot_base, uint32_t slot_cap, int build_by_hess,
                                                           int32_t* slot_owner) {
  __shared__ uint32_t s_warp[33];
  __shared__ uint32_t s_tile_begin[2 * kMaxLevelNodes + 1];
  const int n_part = ts->n_part;
  const uint32_t T = ts->part_tiles;
  if (n_part == 0) return;
  // 0. take the window
  for (uint32_t i = threadIdx.x; i < 2u * slot_cap; i += kCtlThreads) slot_owner[slot_base + i] = -1;
  // 1. exclusive scan of the tile left-counts over the whole level
  uint32_t running = 0;
  for (uint32_t base = 0; base < T; base += kCtlThreads) {
    const uint32_t t = base + threadIdx.x;
    uint32_t v = t < T ? tile_left[t] : 0u;
    uint32_t total;
    uint32_t ex = running + BlockExclusiveScan(v, s_warp, &total);
    if (t < T) tile_excl[t] = ex;
    running += total;
  }
  if (threadIdx.x == 0) tile_excl[T] = running;
  __syncthreads();
  // 2. per segment: left count, children segments, build list (1 or 2 entries per split)
  unsigned long long visits = 0;
  uint32_t n_build = 0;
  for (int base = 0; base < n_part; base += kCtlThreads) {
    const int k = base + threadIdx.x;
    PartEntry e{};
    bool parent_cached = true;
    bool build_right = false;
    if (k < n_part) {
      e = part[k];
      const uint32_t tb = e.tile_begin;
      const uint32_t te = (k + 1 < n_part) ? part[k + 1].tile_begin : T;
      const uint32_t n_left = tile_excl[te] - tile_excl[tb];
      part[k].n_left = n_left;
      NodeDev* l = &nodes[e.left];
      NodeDev* r = &nodes[e.right];
      const int dst_buf = 1 - e.src_buf;
      l->seg_begin = e.begin; l->seg_count = n_left; l->seg_buf = dst_buf;
      r->seg_begin = e.begin + n_left; r->seg_count = e.count - n_left; r->seg_buf = dst_buf;
      // Which child gets built (the other is parent - built): the one with fewer LOCAL rows is the cheaper one;
      // under row split every rank must pick the same child, so the hessian sums decide, as in the reference
      // (updater_gpu_hist.cu:461-467: right iff hess_R < hess_L).  The results do not depend on the choice.
      build_right = build_by_hess ? (r->sum_h < l->sum_h) : (r->seg_count < l->seg_count);
      const int32_t ps = nodes[e.nidx].hist_slot;
      parent_cached = ps >= 0 && slot_owner[ps] == e.nidx;
      eval_nodes[2 * k] = e.left;
      eval_nodes[2 * k + 1] = e.right;
    }
    uint32_t total;
    const uint32_t first = n_build + BlockExclusiveScan(k < n_part ? (parent_cached ? 1u : 2u) : 0u, s_warp, &total);
    if (k < n_part) {
      const int dst_buf = 1 - e.src_buf;
      const int b_idx = build_right ? e.right : e.left;
      const int s_idx = build_right ? e.left : e.right;
      const NodeDev* b = &nodes[b_idx];
      build[first] = BuildEntry{b_idx, dst_buf, b->seg_begin, b->seg_count, 0u, slot_base + first};
      nodes[b_idx].hist_slot = (int32_t)(slot_base + first);
      slot_owner[slot_base + first] = b_idx;
      visits += b->seg_count;
      if (!parent_cached) {
        const NodeDev* o = &nodes[s_idx];
        build[first + 1] = BuildEntry{s_idx, dst_buf, o->seg_begin, o->seg_count, 0u, slot_base + first + 1};
        nodes[s_idx].hist_slot = (int32_t)(slot_base + first + 1);
        slot_owner[slot_base + first + 1] = s_idx;
        visits += o->seg_count;
      }
      part[k].pad = first | (parent_cached ? 0x80000000u : 0u);  // for the second pass
    }
    n_build += total;
  }
  __syncthreads();
  // 2b. subtracted siblings: slots behind the built ones, in split order
  const uint32_t n_sub = 2u * (uint32_t)n_part - n_build;
  for (int k = threadIdx.x; k < n_part; k += kCtlThreads) {
    const uint32_t tag = part[k].pad;
    if ((tag & 0x80000000u) == 0u) continue;
    const uint32_t first = tag & 0x7FFFFFFFu;
    const uint32_t rank = 2u * (uint32_t)k - first;  // cached parents in front of split k
    const PartEntry e = part[k];
    const int b_idx = build[first].nidx;
    const int s_idx = b_idx == e.left ? e.right : e.left;
    const uint32_t s_slot = slot_base + n_build + rank;
    nodes[s_idx].hist_slot = (int32_t)s_slot;
    slot_owner[s_slot] = s_idx;
    sub[rank] = SubEntry{nodes[e.nidx].hist_slot, (int32_t)(slot_base + first), (int32_t)s_slot};
  }
  // 3. segment-relative tile offsets for the scatter pass
  for (uint32_t t = threadIdx.x; t < T; t += kCtlThreads)
    tile_left_off[t] = tile_excl[t] - tile_excl[part[part_tile_seg[t]].tile_begin];
  // 4. histogram tile table
  uint32_t running_tiles = 0;
  for (uint32_t base = 0; base < n_build; base += kCtlThreads) {
    const uint32_t j = base + threadIdx.x;
    uint32_t tiles = j < n_build ? (build[j].count + kHistTile - 1) / kHistTile : 0u;
    uint32_t total;
    uint32_t ex = running_tiles + BlockExclusiveScan(tiles, s_warp, &total);
    if (j < n_build) {
      build[j].tile_begin = ex;
      s_tile_begin[j] = ex;
    }
    running_tiles += total;
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < running_tiles; t += kCtlThreads)
    build_tile_seg[t] = FindSegment(s_tile_begin, n_build, t);
  if (visits) atomicAdd(&ts->row_visits, visits);
  if (threadIdx.x == 0) {
    ts->n_build = (int32_t)n_build;
    ts->n_sub = (int32_t)n_sub;
    ts->n_eval = 2 * n_part;
    ts->build_tiles = running_tiles;
  }
}

// ---------------------------------------------------------------------------------------------
// Feature interaction constraints (FeatureInteractionConstraintDevice, src/tree/constraints.cu).
// One allowed-feature bit field per node (bit f of word f / 64), cleared per tree (Reset, constraints.cu:134-138).
// After a level's ScanKernel this control kernel
//   1. applies Split(nidx, findex, left, right) of every split whose children get evaluated (constraints.cu:269-329:
//      node |= the sets that hold findex; left |= node; right |= node; then bit findex in all three),
//   2. answers Query(sampled features of the depth, child) for every evaluation slot in the order of the reference's
//      loop (left0, right0, left1, ...; updater_gpu_hist.cu:251-268): the allowed features in ASCENDING order,
//   3. and lays those answers over each other the way the reference's single result buffer ends up
//      (constraints.cu:125,199-240): entry i of the buffer is entry i of the LAST answer longer than i.  A node then
//      reads the first count[slot] entries (IcView, evaluate.cu).  Pinned by tests/golden/ref_gpu_interaction_trees.json.
// Splits whose children are not evaluated need no bit fields: nothing reads them again.
// ---------------------------------------------------------------------------------------------
struct IcDev {
  int32_t n_features, words, n_sets;
  const int32_t* set_ptr;   // CSR of the constraint sets
  const int32_t* set_feat;
  unsigned long long* bits;  // [max_nodes][words]
  int32_t* list;             // [n_features] the level's buffer image
  int32_t* count;            // [2 * kMaxLevelNodes]
};

__global__ void __launch_bounds__(kCtlThreads) IcLevelKernel(const TreeState* ts, const NodeDev* nodes,
                                                              const PartEntry* part, const int32_t* eval_nodes,
                                                              IcDev ic, const int32_t* feature_set, int32_t n_fs,
                                                              int32_t n_levels, const int32_t* slot_lists,
                                                              int32_t slot_stride) {
  B200_DYNAMIC_SHARED(unsigned long long, s_mask);  // [words] the depth's sampled features
  const int n_part = ts->n_part;
  if (n_part == 0) return;
  const int W = ic.words;
  // 1. Split: one thread per applied split; parents of a level are distinct nodes, children are fresh
  for (int k = threadIdx.x; k < n_part; k += kCtlThreads) {
    const PartEntry e = part[k];
    unsigned long long* nb = ic.bits + (size_t)e.nidx * W;
    unsigned long long* lb = ic.bits + (size_t)e.left * W;
    unsigned long long* rb = ic.bits + (size_t)e.right * W;
    for (int s = 0; s < ic.n_sets; ++s) {
      bool holds = false;
      for (int j = ic.set_ptr[s]; j < ic.set_ptr[s + 1]; ++j) holds = holds || ic.set_feat[j] == e.findex;
      if (!holds) continue;
      for (int j = ic.set_ptr[s]; j < ic.set_ptr[s + 1]; ++j) nb[ic.set_feat[j] >> 6] |= 1ull << (ic.set_feat[j] & 63);
    }
    nb[e.findex >> 6] |= 1ull << (e.findex & 63);
    for (int w = 0; w < W; ++w) { lb[w] |= nb[w]; rb[w] |= nb[w]; }
  }
  // the depth's sampled features as a bit mask (all features without column sampling)
  const int depth = nodes[eval_nodes[0]].depth;
  for (int w = threadIdx.x; w < W; w += kCtlThreads) {
    unsigned long long m = 0;
    if (feature_set == nullptr) {
      const int lo = w * 64, hi = min(ic.n_features, lo + 64);
      m = hi - lo >= 64 ? ~0ull : ((1ull << (hi - lo)) - 1ull);
    }
    s_mask[w] = m;
  }
  __syncthreads();
  if (feature_set != nullptr) {
    const int32_t* fs = n_levels > 0 ? feature_set + (size_t)min(depth, n_levels - 1) * n_fs : feature_set;
    for (int j = threadIdx.x; j < n_fs; j += kCtlThreads)
      atomicOr(&s_mask[fs[j] >> 6], 1ull << (fs[j] & 63));
  }
  __syncthreads();
  // 2. counts of the answers.  slot_lists (colsample_bynode): Query(the node's OWN sampled list, nidx) -- the list is
  //    ascending, so the answer is the list filtered by the node's bit field, in order
  const int n_eval = 2 * n_part;
  for (int j = threadIdx.x; j < n_eval; j += kCtlThreads) {
    const unsigned long long* b = ic.bits + (size_t)eval_nodes[j] * W;
    int c = 0;
    if (slot_lists != nullptr) {
      const int32_t* mine = slot_lists + (size_t)j * slot_stride;
      for (int k = 0; k < slot_stride; ++k) c += (int)((b[mine[k] >> 6] >> (mine[k] & 63)) & 1ull);
    } else {
      for (int w = 0; w < W; ++w) c += __popcll(b[w] & s_mask[w]);
    }
    ic.count[j] = c;
  }
  __syncthreads();
  // 3. buffer image: answer j survives at positions [max count of the later answers, count_j)
  for (int j = threadIdx.x; j < n_eval; j += kCtlThreads) {
    const int cj = ic.count[j];
    int later = 0;
    for (int j2 = j + 1; j2 < n_eval; ++j2) later = max(later, ic.count[j2]);
    if (cj <= later) continue;
    const unsigned long long* b = ic.bits + (size_t)eval_nodes[j] * W;
    int pos = 0;
    if (slot_lists != nullptr) {
      const int32_t* mine = slot_lists + (size_t)j * slot_stride;
      for (int k = 0; k < slot_stride && pos < cj; ++k) {
        if ((b[mine[k] >> 6] >> (mine[k] & 63)) & 1ull) {
          if (pos >= later) ic.list[pos] = mine[k];
          ++pos;
        }
      }
      continue;
    }
    for (int w = 0; w < W && pos < cj; ++w) {
      unsigned long long v = b[w] & s_mask[w];
      while (v) {
        const int bit = __ffsll((long long)v) - 1;
        v &= v - 1;
        if (pos >= later) ic.list[pos] = w * 64 + bit;
        ++pos;
      }
    }
  }
}

// colsample_bynode (ColumnSampler::GetFeatureSet, src/common/random.h:156-175): every call draws a fresh set from
// the level's set, and gpu_hist calls it once for the root and then for the left and the right child of every valid
// candidate in Pop order (updater_gpu_hist.cu:217,257-264).  The draws form ONE stream per tree, so the caller
// hands over the stream's first max_nodes sets and evaluation slot j of the current step takes draw seq + j.
__global__ void __launch_bounds__(kCtlThreads) NodeSetLevelKernel(const TreeState* ts, const int32_t* __restrict__ sets,
                                                                   int32_t n_sets, int32_t n_per,
                                                                   int32_t* __restrict__ seq,
                                                                   int32_t* __restrict__ list,
                                                                   int32_t* __restrict__ count) {
  const int n_eval = ts->n_eval;
  const int base = *seq;
  __syncthreads();  // everybody has read the sequence number before it moves
  for (int i = threadIdx.x; i < n_eval * n_per; i += kCtlThreads) {
    const int j = i / n_per, k = i - j * n_per;
    const int draw = min(base + j, n_sets - 1);  // the caller provides max_nodes draws: never short
    list[i] = sets[(size_t)draw * n_per + k];
  }
  for (int j = threadIdx.x; j < n_eval; j += kCtlThreads) count[j] = n_per;
  if (threadIdx.x == 0) *seq = base + n_eval;
}

// ---------------------------------------------------------------------------------------------
// FinalisePosition (updater_gpu_hist.cu:529-580): list the nodes that still own a row segment,
// then walk each row from that node to its leaf (at most the splits applied after the last
// partition) and write position[row] = leaf, bitwise-NOT when the row's hessian is 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCtlThreads) FinalListKernel(TreeState* ts, const NodeDev* nodes, BuildEntry* fin,
                                                                uint32_t* fin_tile_seg, uint32_t* fin_tile_begin) {
  __shared__ uint32_t s_warp[33];
  const int n_nodes = ts->n_nodes;
  uint32_t running_n = 0, running_tiles = 0;
  for (int base = 0; base < n_nodes; base += kCtlThreads) {
    const int i = base + threadIdx.x;
    bool term = i < n_nodes && nodes[i].seg_buf >= 0 && !nodes[i].partitioned;
    uint32_t tiles = term ? (nodes[i].seg_count + kPartTile - 1) / kPartTile : 0u;
    uint32_t total_n, total_t;
    uint32_t slot = running_n + BlockExclusiveScan(term ? 1u : 0u, s_warp, &total_n);
    uint32_t tb = running_tiles + BlockExclusiveScan(tiles, s_warp, &total_t);
    if (term) {
      fin[slot] = BuildEntry{i, nodes[i].seg_buf, nodes[i].seg_begin, nodes[i].seg_count, tb, 0u};
      fin_tile_begin[slot] = tb;
    }
    running_n += total_n;
    running_tiles += total_t;
  }
  __syncthreads();
  __threadfence_block();
  for (uint32_t t = threadIdx.x; t < running_tiles; t += kCtlThreads)
    fin_tile_seg[t] = FindSegment(fin_tile_begin, running_n, t);
  if (threadIdx.x == 0) {
    ts->n_final = (int32_t)running_n;
    ts->final_tiles = running_tiles;
  }
}

template <typename BinT, bool kCat>
__global__ void __launch_bounds__(256) FinaliseKernel(EllView ell, const uint32_t* __restrict__ ridx0,
                                                       const uint32_t* __restrict__ ridx1,
                                                       const longlong2* __restrict__ gq,
                                                       const NodeDev* __restrict__ nodes,
                                                       const BuildEntry* __restrict__ fin,
                                                       const uint32_t* __restrict__ fin_tile_seg,
                                                       const TreeState* __restrict__ ts, int32_t* __restrict__ position,
                                                       const QuantDev* __restrict__ qd, CatRoute cr) {
  const uint32_t n_tiles = ts->final_tiles;
  const bool any_h_zero = qd->h_zero_any != 0u;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const BuildEntry e = fin[fin_tile_seg[t]];
    const uint32_t off = (t - e.tile_begin) * kPartTile;
    const uint32_t n = min((uint32_t)kPartTile, e.count - off);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin + off;
    for (uint32_t j = threadIdx.x; j < n; j += 256) {
      const uint32_t r = ridx[j];
      int nidx = e.nidx;
      while (!nodes[nidx].is_leaf) {
        const NodeDev& nd = nodes[nidx];
        if (ell.is_sparse) {
          const uint32_t fb = ell.cut_ptrs[nd.findex], fe = ell.cut_ptrs[nd.findex + 1];
          const uint32_t gb = SparseRowBin(ell, r, fb, fe, (uint32_t)nd.findex);
          const bool go_left = (gb == 0xFFFFFFFFu) ? (nd.default_left != 0)
                                                   : GoLeftBin<kCat>((int32_t)(gb - fb), nd.split_lbin, nidx, cr);
          nidx = go_left ? nd.left : nd.right;
          continue;
        }
        const uint32_t sym = (sizeof(BinT) <= 2 && ell.bins_t != nullptr)
                                 ? (uint32_t)reinterpret_cast<const BinT*>(ell.bins_t)[(uint64_t)nd.findex * ell.col_stride + r]
                                 : (uint32_t)reinterpret_cast<const BinT*>(ell.bins + (size_t)r * ell.row_bytes)[nd.findex];
        const bool left = (sym == ell.null_symbol) ? (nd.default_left != 0)
                                                   : GoLeftBin<kCat>((int32_t)sym, nd.split_lbin, nidx, cr);
        nidx = left ? nd.left : nd.right;
      }
      // EncodeOp (updater_gpu_hist.cuh:48-56): rows whose hessian is 0 get the bit-complemented id; the
      // quantiser tells us when no such row exists so the gradient gather can be skipped
      position[r] = (any_h_zero && gq[r].y == 0) ? ~nidx : nidx;
    }
  }
}

// Windowed FinalisePosition: the tile version above scatters 4-byte position writes all over the row range (a
// terminal segment holds every ~2^depth-th row), so each write dirties its own 32-byte sector.  Every segment
// lists its rows in ascending order (the partition is stable), so a CTA can instead take a WINDOW of
// kFinWindow consecutive row ids, find by binary search the slice of every terminal segment that falls inside
// it, resolve those rows' leaves into a shared-memory image of the window, and write the window's positions
// (and read the hessians for the bit-complement rule) fully coalesced.
constexpr uint32_t kFinWindow = 8192;

__device__ __forceinline__ uint32_t LowerBoundRow(const uint32_t* __restrict__ r, uint32_t lo, uint32_t hi, uint32_t v) {
  while (lo < hi) {  // first index in [lo, hi) with r[idx] >= v
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(r + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename BinT, bool kCat>
__device__ __forceinline__ int WalkToLeaf(const EllView& ell, const NodeDev* __restrict__ nodes, int nidx, uint32_t r,
                                          const CatRoute& cr) {
  while (!nodes[nidx].is_leaf) {
    const NodeDev& nd = nodes[nidx];
    if (ell.is_sparse) {
      const uint32_t fb = ell.cut_ptrs[nd.findex], fe = ell.cut_ptrs[nd.findex + 1];
      const uint32_t gb = SparseRowBin(ell, r, fb, fe, (uint32_t)nd.findex);
      const bool go_left = (gb == 0xFFFFFFFFu) ? (nd.default_left != 0)
                                               : GoLeftBin<kCat>((int32_t)(gb - fb), nd.split_lbin, nidx, cr);
      nidx = go_left ? nd.left : nd.right;
      continue;
    }
    const uint32_t sym = (sizeof(BinT) <= 2 && ell.bins_t != nullptr)
                             ? (uint32_t)reinterpret_cast<const BinT*>(ell.bins_t)[(uint64_t)nd.findex * ell.col_stride + r]
                             : (uint32_t)reinterpret_cast<const BinT*>(ell.bins + (size_t)r * ell.row_bytes)[nd.findex];
    const bool left = (sym == ell.null_symbol) ? (nd.default_left != 0)
                                               : GoLeftBin<kCat>((int32_t)sym, nd.split_lbin, nidx, cr);
    nidx = left ? nd.left : nd.right;
  }
  return nidx;
}

template <typename BinT, bool kCat>
__global__ void __launch_bounds__(256) FinaliseWindowKernel(EllView ell, const uint32_t* __restrict__ ridx0,
                                                             const uint32_t* __restrict__ ridx1,
                                                             const longlong2* __restrict__ gq,
                                                             const NodeDev* __restrict__ nodes,
                                                             const BuildEntry* __restrict__ fin,
                                                             const TreeState* __restrict__ ts,
                                                             int32_t* __restrict__ position,
                                                             const QuantDev* __restrict__ qd, uint32_t row_lo,
                                                             uint32_t n_rows, CatRoute cr) {
  // rows [row_lo, n_rows): the whole matrix, or the rows of the resident page (external memory: one launch per page)
  __shared__ int32_t s_pos[kFinWindow];
  __shared__ uint32_t s_lo[256], s_hi[256], s_off[257], s_warp[33];
  const uint32_t n_final = (uint32_t)ts->n_final;
  const bool any_h_zero = qd->h_zero_any != 0u;
  const uint32_t n_windows = (n_rows - row_lo + kFinWindow - 1) / kFinWindow;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // a CTA takes CONSECUTIVE windows: a slice starts where the previous window's slice ended, so only the first
  // window of a CTA pays the full binary search per segment (when all segments fit one 256-thread chunk)
  const uint32_t w_first = (uint32_t)(((unsigned long long)blockIdx.x * n_windows) / gridDim.x);
  const uint32_t w_last = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_windows) / gridDim.x);
  const bool carry = n_final <= 256u;
  for (uint32_t wdx = w_first; wdx < w_last; ++wdx) {
    const uint32_t r_begin = row_lo + wdx * kFinWindow, r_end = min(n_rows, r_begin + kFinWindow);
    for (uint32_t base = 0; base < n_final; base += 256) {
      if (base + threadIdx.x < n_final) {  // one thread per terminal segment: its slice inside the window
        const BuildEntry e = fin[base + threadIdx.x];
        const uint32_t* __restrict__ r = (e.buf ? ridx1 : ridx0) + e.begin;
        const uint32_t lo = (carry && wdx > w_first) ? s_hi[threadIdx.x] : LowerBoundRow(r, 0u, e.count, r_begin);
        s_lo[threadIdx.x] = lo;
        s_hi[threadIdx.x] = LowerBoundRow(r, lo, min(e.count, lo + kFinWindow), r_end);
      }
      __syncthreads();
      const uint32_t m = min(256u, n_final - base);
      // rows of all slices of this chunk, concatenated: entry e belongs to the slice whose offset range holds it.
      // Work is dealt out by ROWS, 32 consecutive entries per warp step, so one huge leaf (an unbalanced tree keeps
      // most rows in a few leaves) is shared by all warps instead of serialising on one.
      uint32_t total_rows;
      const uint32_t cnt = threadIdx.x < m ? s_hi[threadIdx.x] - s_lo[threadIdx.x] : 0u;
      const uint32_t off = BlockExclusiveScan(cnt, s_warp, &total_rows);
      s_off[threadIdx.x] = off;
      if (threadIdx.x == 0) s_off[256] = total_rows;
      __syncthreads();
      // four entries per lane and step: their row-id loads are issued together (the chain row id -> node -> symbol
      // is latency bound), then the walks
      for (uint32_t e0 = warp * 128u; e0 < total_rows; e0 += 8u * 128u) {
        uint32_t row[4];
        int nid[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t eidx = e0 + u * 32u + lane;
          ok[u] = eidx < total_rows;
          row[u] = 0u;
          nid[u] = 0;
          if (ok[u]) {
            uint32_t lo = 0, hi = m;  // last slice with s_off[k] <= eidx
            while (hi - lo > 1) {
              const uint32_t mid = (lo + hi) >> 1;
              if (s_off[mid] <= eidx) lo = mid; else hi = mid;
            }
            const BuildEntry en = fin[base + lo];
            nid[u] = en.nidx;
            row[u] = __ldg((en.buf ? ridx1 : ridx0) + en.begin + s_lo[lo] + (eidx - s_off[lo]));
          }
        }
        if (ell.is_sparse || ell.bins_t == nullptr) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (ok[u]) s_pos[row[u] - r_begin] = WalkToLeaf<BinT, kCat>(ell, nodes, nid[u], row[u], cr);
        } else {  // the four walks level by level, so that their symbol gathers overlap
          bool live[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) live[u] = ok[u] && !nodes[nid[u]].is_leaf;
          while (live[0] || live[1] || live[2] || live[3]) {
            uint32_t sym[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              sym[u] = live[u] ? (uint32_t)__ldg(reinterpret_cast<const BinT*>(ell.bins_t) +
                                                 (uint64_t)nodes[nid[u]].findex * ell.col_stride + row[u])
                               : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (live[u]) {
                const NodeDev& nd = nodes[nid[u]];
                const bool left = (sym[u] == ell.null_symbol) ? (nd.default_left != 0)
                                                              : GoLeftBin<kCat>((int32_t)sym[u], nd.split_lbin, nid[u], cr);
                nid[u] = left ? nd.left : nd.right;
                live[u] = !nodes[nid[u]].is_leaf;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (ok[u]) s_pos[row[u] - r_begin] = nid[u];
        }
      }
      __syncthreads();
    }
    // EncodeOp (updater_gpu_hist.cuh:48-56): rows whose hessian is 0 get the bit-complemented id
    for (uint32_t i = threadIdx.x; i < r_end - r_begin; i += 256) {
      int32_t p = s_pos[i];
      if (any_h_zero && gq[r_begin + i].y == 0) p = ~p;
      position[r_begin + i] = p;
    }
    __syncthreads();
  }
}

// UpdatePredictionCache (updater_gpu_hist.cu:582-608)
__global__ void __launch_bounds__(256) PredCacheKernel(const int32_t* __restrict__ position,
                                                        const NodeDev* __restrict__ nodes, uint64_t n,
                                                        float* __restrict__ preds) {
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
    int32_t p = position[i];
    p = p < 0 ? ~p : p;  // SamplePosition::Decode
    preds[i] += nodes[p].leaf_value;
  }
}

// The same update for a caller whose output is a strided column (multi-class: out.Slice(All, gid), stride = n_groups,
// gbtree.cc DoBoost) and/or whose leaf values were rewritten after Update (adaptive objectives: obj->UpdateTreeLeaf
// runs between Update and UpdatePredictionCache, which is why the reference reads p_tree->GetNodes() there,
// updater_gpu_hist.cu:594-598).  `leaves` is indexed by node id.
__global__ void __launch_bounds__(256) PredCacheStridedKernel(const int32_t* __restrict__ position,
                                                               const float* __restrict__ leaves, uint64_t n,
                                                               uint64_t stride, float* __restrict__ preds) {
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
    int32_t p = position[i];
    p = p < 0 ? ~p : p;  // SamplePosition::Decode
    preds[i * stride] += leaves[p];
  }
}

__global__ void __launch_bounds__(256) GatherLeavesKernel(const NodeDev* __restrict__ nodes, int32_t n,
                                                           float* __restrict__ leaves) {
  int32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) leaves[i] = nodes[i].leaf_value;
}

}  // namespace b200

#ifndef B200_KERNELS_ONLY
using namespace b200;

// =============================================================================================
// Updater object
// =============================================================================================
struct b200hist_updater_s {
  b200hist_ellpack_t ell{nullptr};
  int device{0};
  int num_sms{0};
  b200hist_train_param param{};
  ParamDev pdev{};
  int32_t max_nodes{0};
  int32_t n_steps{0};  // growth steps (levels for depth-wise)
  std::vector<GroupDesc> groups;
  // device state
  GroupDesc* d_groups{nullptr};
  QuantDev* d_quant{nullptr};
  double* d_partials{nullptr};
  unsigned int* d_counter{nullptr};
  b200hist_gpair64* d_gq{nullptr};
  uint32_t* d_ridx[2]{nullptr, nullptr};
  int32_t* d_position{nullptr};
  float* d_gpair_staging{nullptr};
  TreeState* d_ts{nullptr};
  NodeDev* d_nodes{nullptr};
  b200hist_candidate* d_cand{nullptr};
  PartEntry* d_part{nullptr};
  BuildEntry* d_build{nullptr};
  SubEntry* d_sub{nullptr};
  int32_t* d_eval_nodes{nullptr};
  uint32_t* d_part_tile_seg{nullptr};
  uint32_t* d_build_tile_seg{nullptr};
  uint32_t* d_tile_left{nullptr};
  uint32_t* d_tile_excl{nullptr};
  uint32_t* d_tile_left_off{nullptr};
  uint32_t* d_flags{nullptr};
  BuildEntry* d_fin{nullptr};
  uint32_t* d_fin_tile_seg{nullptr};
  uint32_t* d_fin_tile_begin{nullptr};
  FeatCand* d_scratch{nullptr};
  b200hist_gpair64* d_hist{nullptr};
  size_t hist_bytes{0};
  std::vector<uint32_t> step_cap;        // max nodes expanded in step s (host-known)
  std::vector<uint32_t> step_slot_base;  // first histogram slot of step s
  std::vector<int> step_level;           // tree level (depth-wise) / split number (loss-guide) the step belongs to
  std::vector<int> step_level_start;     // 1: the step opens its level, 0: it continues it (node batching)
  int level_batch{kMaxLevelNodes};       // valid entries one step expands at most
  std::vector<int> step_reuses;          // 1: the step's slot window has held other histograms in this tree (ring wrapped)
  int32_t* d_slot_owner{nullptr};        // [n_slots] node that owns the slot's histogram, -1 = none
  bool ring_reuse{false};                // the cache is smaller than the tree: windows are reused, parents may get lost
  uint32_t n_slots{1};
  // multi-GPU: feature-group shard of this rank and the allreduce communicator (borrowed)
  b200hist_comm_s* comm{nullptr};
  uint32_t group_begin{0}, group_end{0};
  // peer-memory exchange (P2P over NVLink): IPC-mapped histogram caches of all ranks
  bool p2p{false};
  int p2p_rank{0}, p2p_world{1};
  // per-tree sampling state (set through b200hist_updater_set_tree_sampling, persists until changed)
  float subsample{1.0f};
  uint32_t sample_seed{0};
  bool gradient_based{false};       // TrainParam::sampling_method == kGradientBased (applies when subsample < 1)
  void* d_gbs_scratch{nullptr};
  size_t gbs_scratch_bytes{0};
  int32_t* d_monotone{nullptr};     // monotone constraint sign per feature (nullptr = none)
  float* d_lower{nullptr};          // TreeEvaluator weight bounds per node (allocated with d_monotone)
  float* d_upper{nullptr};
  int32_t* d_feature_set{nullptr};  // sorted feature ids of the tree's colsample_bytree draw (nullptr = all)
  int32_t n_feature_set{0};
  int32_t n_feature_levels{0};      // > 0: d_feature_set holds one list per depth (colsample_bylevel)
  int32_t last_max_eval_depth{0};
  // feature interaction constraints (nullptr = none): constraint sets as CSR, per-node allowed bit fields, the
  // level's Query buffer image and per-slot counts (IcLevelKernel)
  int32_t* d_ic_set_ptr{nullptr};
  int32_t* d_ic_set_feat{nullptr};
  unsigned long long* d_ic_bits{nullptr};
  int32_t* d_ic_list{nullptr};
  int32_t* d_ic_count{nullptr};
  int32_t ic_n_sets{0}, ic_words{0};
  // colsample_bynode (nullptr = none): the tree's stream of node feature sets, the per-slot lists of the current
  // step, the stream position (device) and the position after the last tree (host)
  int32_t* d_node_sets{nullptr};
  int32_t* d_node_list{nullptr};
  int32_t* d_node_count{nullptr};
  int32_t* d_node_seq{nullptr};
  int32_t n_node_sets{0}, n_per_node{0}, last_eval_count{0};
  // seed-driven column sampler (nullptr = none): engine outputs, cursor, tree set, lazily drawn level sets
  uint32_t* d_col_seeds{nullptr};
  int32_t* d_col_cursor{nullptr};
  int32_t* d_col_tree_set{nullptr};
  int32_t* d_col_level_sets{nullptr};
  int32_t* d_col_level_drawn{nullptr};
  int32_t* d_col_node_list{nullptr};
  int32_t* d_col_node_count{nullptr};
  int32_t col_n_seeds{0}, col_n_tree{0}, col_n_level{0}, col_n_node{0}, col_max_levels{0}, last_draw_count{0};
  bool col_bylevel_is_one{true}, col_bynode_is_one{true};
  // categorical features (nullptr = none): flags per feature, per-node category sets (CatView)
  uint8_t* d_is_cat{nullptr};
  uint32_t* d_cat_bits{nullptr};
  uint32_t* d_bin_bits{nullptr};
  uint8_t* d_node_is_cat{nullptr};
  int32_t cat_words{0}, cat_bin_words{0}, max_cat_bins{0};
  int32_t max_cat_to_onehot{4}, max_cat_threshold{64};
  bool row_split{false};                   // ranks own row slices (full-width partial histograms, summed) instead of feature blocks
  size_t flags_offset{0};                  // byte offset of the flag array inside the histogram allocation
  std::vector<void*> peer_base;            // opened IPC mappings (own entry = nullptr)
  longlong2** d_peer_hist{nullptr};
  uint32_t** d_peer_flags{nullptr};
  FeatCand** d_peer_scratch{nullptr};      // every rank's evaluation scratch table (inside the shared allocation)
  size_t scratch_offset{0};                // byte offset of the scratch table inside the histogram allocation
  bool last_sliced{false};                 // the last tree was evaluated by feature slice (node_histogram assembles)
  unsigned int* d_exchange_counter{nullptr};
  uint32_t exchange_epoch{0};
  // host staging
  NodeDev* h_nodes{nullptr};  // pinned
  TreeState* h_ts{nullptr};   // pinned
  QuantDev* h_quant{nullptr}; // pinned copy of the last tree's quantiser record
  // deferred read-back (b200hist_updater_update_tree_async / fetch_tree): up to two trees in flight, each with its own
  // pinned staging, "tree landed" event and phase-event pool
  struct PendingTree {
    NodeDev* h_nodes{nullptr};
    TreeState* h_ts{nullptr};
    QuantDev* h_quant{nullptr};
    cudaEvent_t landed{nullptr};
    std::vector<cudaEvent_t> events;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
    bool busy{false};
  } pending[2];
  uint64_t async_issued{0}, async_fetched{0};
  bool pair_hint{false};      // last tree's hessians had zero low words -> launch the 64-feature pair kernel
  int32_t* last_position{nullptr};
  int32_t last_n_nodes{0};
  float* d_leaf_override{nullptr};  // leaf values handed to update_prediction_cache_ex (or gathered for a strided output)
  // profiling
  bool profiling{false};
  std::vector<cudaEvent_t> events;
  float last_ms[8]{0, 0, 0, 0, 0, 0, 0, 0};  // [7] = the cross-GPU exchange, also contained in [1] (hist)
  uint64_t last_row_visits{0};
  std::vector<float> last_hist_ms;  // the histogram span of every growth step of the last profiled tree
};

namespace {

enum Phase { kQuant = 0, kHist = 1, kSub = 2, kEval = 3, kApply = 4, kPart = 5, kFinal = 6, kComm = 7 };

struct PhaseTimer {
  b200hist_updater_s* u;
  cudaStream_t s;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
  size_t next{0};
  std::vector<cudaEvent_t>* pool{nullptr};  // nullptr: the updater's own pool (synchronous trees)
  cudaEvent_t Get() {
    auto& ev = pool ? *pool : u->events;
    if (next == ev.size()) {
      cudaEvent_t e;
      B200_CUDA(cudaEventCreate(&e));
      ev.push_back(e);
    }
    return ev[next++];
  }
  cudaEvent_t Begin() {
    if (!u->profiling) return nullptr;
    cudaEvent_t e = Get();
    B200_CUDA(cudaEventRecord(e, s));
    return e;
  }
  void End(int phase, cudaEvent_t b) {
    if (!u->profiling) return;
    cudaEvent_t e = Get();
    B200_CUDA(cudaEventRecord(e, s));
    spans.push_back({phase, {b, e}});
  }
  void Collect() {
    if (!u->profiling) return;
    for (auto& m : u->last_ms) m = 0;
    u->last_hist_ms.clear();
    for (auto& sp : spans) {
      float ms = 0;
      B200_CUDA(cudaEventElapsedTime(&ms, sp.second.first, sp.second.second));
      u->last_ms[sp.first] += ms;
      if (sp.first == kHist) u->last_hist_ms.push_back(ms);  // one span per growth step (level), exchange included
    }
  }
};

int32_t MaxNodesFor(const b200hist_train_param& p) {
  long long by_depth = p.max_depth > 0 ? ((1ll << (p.max_depth + 1)) - 1) : (1ll << 40);
  long long by_leaves = p.max_leaves > 0 ? (2ll * p.max_leaves - 1) : (1ll << 40);
  long long n = std::min(by_depth, by_leaves);
  return (int32_t)n;
}

// bins of the features rank owns under EvalShard::Owns: f * world / n_features == rank
void SliceBins(const b200hist_ellpack_s& ell, int rank, int world, uint32_t* lo, uint32_t* hi) {
  const EvalShard sh{rank, world, ell.n_features};
  uint32_t f0 = ell.n_features, f1 = 0;
  for (uint32_t f = 0; f < ell.n_features; ++f)
    if (sh.Owns((int)f)) { f0 = std::min(f0, f); f1 = std::max(f1, f + 1); }
  if (f0 >= f1) { *lo = 0; *hi = 0; return; }
  *lo = ell.cut_ptrs[f0];
  *hi = ell.cut_ptrs[f1];
}

void ToAbiNodes(const NodeDev* src, int32_t n, b200hist_tree_node* dst) {
  for (int32_t i = 0; i < n; ++i) {
    const NodeDev& s = src[i];
    b200hist_tree_node d;
    d.parent = s.parent; d.left = s.left; d.right = s.right; d.findex = s.findex;
    d.fvalue = s.is_leaf ? s.leaf_value : s.fvalue;
    d.default_left = s.default_left; d.is_leaf = s.is_leaf; d.leaf_value = s.leaf_value;
    d.loss_chg = s.loss_chg; d.sum_hess = s.sum_hess; d.base_weight = s.base_weight; d.depth = s.depth;
    dst[i] = d;
  }
}

void UpdateTreeImpl(b200hist_updater_s* u, const float* gpair_dev, cudaStream_t s, b200hist_tree_node* out_nodes,
                    int32_t* out_n_nodes, int32_t* out_position, b200hist_updater_s::PendingTree* defer = nullptr) {
  auto* ell = u->ell;
  const EllView view = ell->View();
  const uint64_t n = ell->n_rows;
  PhaseTimer timer{u, s};
  if (defer) timer.pool = &defer->events;
  // External memory (b200hist_ellpack_create_paged): every phase that reads the quantised matrix -- histogram build,
  // partition count, FinalisePosition -- runs once per page while the NEXT page travels host -> device on the page
  // set's copy stream (two staging buffers).  Replaces the per-page loop of GPUHistMakerDevice
  // (updater_gpu_hist.cu:475-499, 529-580) over EllpackPageSource (ellpack_page_source.h:128-337).  The row
  // state (gradient pairs, row ids, positions) stays resident for ALL rows; a node's segment lists its rows in ascending
  // order, so the rows of one page are a contiguous slice of it (ClipToPage): no per-page partitioner is needed, and
  // integer histograms make the result independent of the paging -- the trees are those of the in-core matrix.
  const uint32_t n_pages = ell->NumPages();
  auto for_each_page = [&](auto&& fn) {
    if (!ell->paged) {
      fn(view);
      return;
    }
    auto cs = static_cast<cudaStream_t>(ell->copy_stream);
    auto prefetch = [&](uint32_t p) {
      const int b = (int)((ell->stage_uses + p) & 1u);
      const uint64_t lo = (uint64_t)p * ell->page_rows;
      const uint64_t rows = std::min<uint64_t>(ell->page_rows, n - lo);
      B200_CUDA(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(ell->ev_done[b]), 0));  // last readers of stage b
      B200_CUDA(cudaMemcpyAsync(ell->d_stage[b], ell->h_bins + lo * (uint64_t)ell->row_bytes,
                                rows * (uint64_t)ell->row_bytes, cudaMemcpyHostToDevice, cs));
      B200_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ell->ev_ready[b]), cs));
      ell->h2d_bytes += rows * (uint64_t)ell->row_bytes;
    };
    prefetch(0);
    for (uint32_t p = 0; p < n_pages; ++p) {
      const int b = (int)((ell->stage_uses + p) & 1u);
      if (p + 1 < n_pages) prefetch(p + 1);
      B200_CUDA(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(ell->ev_ready[b]), 0));
      fn(ell->PageView(p, b));
      B200_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(ell->ev_done[b]), s));
    }
    ell->stage_uses += n_pages;
  };
  if (ell->paged) {
    B200_CHECK(!u->p2p && u->comm == nullptr, "external-memory pages and multi-GPU training are not combined yet");
    B200_CHECK(!ell->is_sparse, "external-memory pages use the dense layout");
  }
  int32_t* position = out_position ? out_position : u->d_position;
  const MonoView mono{u->d_monotone, u->d_lower, u->d_upper};

  // Reset (updater_gpu_hist.cu:172-211): quantiser, quantised gpairs, row indices, histograms
  auto t0 = timer.Begin();
  B200_CUDA(cudaMemsetAsync(u->d_hist, 0, u->hist_bytes, s));
  if (u->p2p)  // nobody may push into a cache that its owner has not cleared yet: pure flag barrier
    LaunchHistExchange(u->d_peer_hist, u->d_peer_flags, u->p2p_rank, u->p2p_world, u->d_ts, 0, 0, false,
                       ell->total_bins, 0, 0, ++u->exchange_epoch, u->d_exchange_counter, u->num_sms, s);
  const bool gbs = u->gradient_based && u->subsample < 1.0f;
  if (gbs) {
    B200_CHECK(!u->p2p && u->comm == nullptr, "gradient-based sampling is single-GPU only so far (the threshold needs a global sort)");
    const size_t need = GradientBasedScratchBytes(n, u->num_sms);
    if (need > u->gbs_scratch_bytes) {
      if (u->d_gbs_scratch) cudaFree(u->d_gbs_scratch);
      u->d_gbs_scratch = nullptr;
      B200_CUDA(cudaMalloc(&u->d_gbs_scratch, need));
      u->gbs_scratch_bytes = need;
    }
  }
  if (u->p2p && u->row_split) {  // global quantiser scale and root sum (quantiser.cu:88-123 GlobalSum)
    LaunchQuantReduce(gpair_dev, n, u->d_quant, u->d_partials, u->d_counter, u->num_sms, s);
    LaunchScalarExchange(u->d_peer_flags, u->p2p_rank, u->p2p_world, ++u->exchange_epoch, 0, u->d_quant, n, s);
    LaunchQuantConvert(gpair_dev, n, u->d_quant, u->d_gq, u->d_ridx[0], u->num_sms, s, u->subsample, u->sample_seed);
    LaunchScalarExchange(u->d_peer_flags, u->p2p_rank, u->p2p_world, ++u->exchange_epoch, 1, u->d_quant, n, s);
  } else {
    // gradient-based sampling works on the quantised pairs of ALL rows: no Bernoulli trial in the conversion
    LaunchQuantise(gpair_dev, n, u->d_quant, u->d_gq, u->d_ridx[0], u->d_partials, u->d_counter, u->num_sms, s,
                   gbs ? 1.0f : u->subsample, u->sample_seed);
    if (gbs)
      LaunchGradientBasedSample(u->d_gq, n, u->d_quant, u->subsample, u->sample_seed, u->d_gbs_scratch,
                                u->gbs_scratch_bytes, u->num_sms, s);
  }
  timer.End(kQuant, t0);
  const bool has_ic = u->d_ic_bits != nullptr;
  const bool has_cat = u->d_is_cat != nullptr;
  const bool has_node_sets = u->d_node_sets != nullptr;
  const bool has_col = u->d_col_seeds != nullptr;
  const ColDrawDev cold{u->d_col_seeds, u->col_n_seeds, u->d_col_cursor, u->d_col_tree_set, u->col_n_tree,
                        u->d_col_level_sets, u->col_n_level, u->d_col_level_drawn, u->col_max_levels,
                        u->col_bylevel_is_one ? 1 : 0};
  if (has_col) {
    B200_CUDA(cudaMemsetAsync(u->d_col_cursor, 0, 2 * sizeof(int32_t), s));
    B200_CUDA(cudaMemsetAsync(u->d_col_level_drawn, 0, sizeof(int32_t) * (size_t)u->col_max_levels, s));
  }
  // Row split: evaluation by feature slice (comm.cu, HistSliceReduceKernel / ScratchExchangeKernel).  Categorical
  // features keep the fused reduce-scatter + all-gather (CatFinishKernel re-reads the winner's histogram, which may be
  // another rank's slice); B200HIST_ROWSPLIT_FULL=1 forces it everywhere (A/B runs).
  static const bool force_full_exchange = std::getenv("B200HIST_ROWSPLIT_FULL") != nullptr;
  const bool sliced = u->p2p && u->row_split && u->d_is_cat == nullptr && !force_full_exchange &&
                      ell->n_features >= (uint32_t)u->p2p_world;
  u->last_sliced = sliced;
  const EvalShard shard{sliced ? u->p2p_rank : 0, sliced ? u->p2p_world : 1, ell->n_features};
  uint32_t slice_bin_lo = 0, slice_bin_hi = ell->total_bins;
  if (sliced) SliceBins(*ell, u->p2p_rank, u->p2p_world, &slice_bin_lo, &slice_bin_hi);
  auto evaluate = [&](int32_t max_eval, const int32_t* fset, int32_t n_fset, int32_t n_lv, IcView icv, const CatView& cv) {
    const int32_t n_fs = fset ? n_fset : (int32_t)ell->n_features;
    std::function<void()> exchange = [&]() {  // its time is part of the evaluation phase
      LaunchScratchExchange(u->d_peer_scratch, u->d_peer_flags, shard, u->d_ts, u->d_eval_nodes, u->d_nodes, fset, n_fs,
                            n_lv, icv, max_eval, ++u->exchange_epoch, u->d_exchange_counter, u->num_sms, s);
    };
    LaunchEvaluate(view, u->d_hist, u->d_eval_nodes, u->d_ts, max_eval, u->d_nodes, fset, n_fset, u->pdev, u->d_quant,
                   u->d_scratch, u->d_cand, u->d_nodes, u->num_sms, s, mono, n_lv, &u->d_ts->max_eval_depth, icv, cv,
                   shard, sliced ? &exchange : nullptr);
  };
  // evaluation through the seed-driven column sampler: level set of the step's depth (drawn on first use), then a
  // node set per evaluation slot (colsample_bynode < 1) or the level set itself
  auto evaluate_with_column_sampler = [&](int32_t max_eval, const CatView& cv, bool root) {
    LaunchLevelDraw(u->d_ts, u->d_nodes, u->d_eval_nodes, cold, s);
    if (has_ic && !root) {  // Query(the node's sampled list, nidx): the root is never filtered (constraints.cu:201-203)
      const IcDev icd{(int32_t)ell->n_features, u->ic_words, u->ic_n_sets, u->d_ic_set_ptr, u->d_ic_set_feat,
                      u->d_ic_bits, u->d_ic_list, u->d_ic_count};
      if (!u->col_bynode_is_one) {
        LaunchNodeDraw(u->d_ts, u->d_nodes, u->d_eval_nodes, cold, u->col_n_node, max_eval, u->d_col_node_list,
                       u->d_col_node_count, s);
        IcLevelKernel<<<1, kCtlThreads, sizeof(unsigned long long) * (size_t)u->ic_words, s>>>(
            u->d_ts, u->d_nodes, u->d_part, u->d_eval_nodes, icd, nullptr, 0, 0, u->d_col_node_list, u->col_n_node);
      } else {
        IcLevelKernel<<<1, kCtlThreads, sizeof(unsigned long long) * (size_t)u->ic_words, s>>>(
            u->d_ts, u->d_nodes, u->d_part, u->d_eval_nodes, icd, u->d_col_level_sets, u->col_n_level, u->col_max_levels,
            nullptr, 0);
      }
      B200_LAUNCHED();
      evaluate(max_eval, nullptr, 0, 0, IcView{u->d_ic_list, u->d_ic_count, 0}, cv);
      return;
    }
    if (!u->col_bynode_is_one) {
      LaunchNodeDraw(u->d_ts, u->d_nodes, u->d_eval_nodes, cold, u->col_n_node, max_eval, u->d_col_node_list,
                     u->d_col_node_count, s);
      evaluate(max_eval, nullptr, 0, 0, IcView{u->d_col_node_list, u->d_col_node_count, u->col_n_node}, cv);
    } else {
      evaluate(max_eval, u->d_col_level_sets, u->col_n_level, u->col_max_levels, IcView{nullptr, nullptr, 0}, cv);
    }
  };
  if (has_node_sets) B200_CUDA(cudaMemsetAsync(u->d_node_seq, 0, sizeof(int32_t), s));
  auto node_sets_view = [&]() {  // the next n_eval draws of the stream, one list per evaluation slot
    NodeSetLevelKernel<<<1, kCtlThreads, 0, s>>>(u->d_ts, u->d_node_sets, u->n_node_sets, u->n_per_node, u->d_node_seq,
                                                 u->d_node_list, u->d_node_count);
    B200_LAUNCHED();
    return IcView{u->d_node_list, u->d_node_count, u->n_per_node};
  };
  const CatView catv = has_cat ? CatView{u->d_is_cat, u->max_cat_to_onehot, u->max_cat_threshold, u->d_cat_bits,
                                         u->d_bin_bits, u->d_node_is_cat, u->cat_words, u->cat_bin_words, u->max_cat_bins}
                               : CatView{nullptr, 0, 0, nullptr, nullptr, nullptr, 0, 0, 0};
  if (has_ic)  // FeatureInteractionConstraintDevice::Reset
    B200_CUDA(cudaMemsetAsync(u->d_ic_bits, 0, sizeof(unsigned long long) * (size_t)u->max_nodes * u->ic_words, s));

  // InitRoot
  auto t1 = timer.Begin();
  B200_CUDA(cudaMemsetAsync(u->d_slot_owner, 0xFF, sizeof(int32_t) * (size_t)u->n_slots, s));  // -1: nobody
  InitRootKernel<<<1, kCtlThreads, 0, s>>>(u->pdev, (uint32_t)n, u->d_quant, u->d_ts, u->d_nodes, u->d_build,
                                           u->d_build_tile_seg, u->d_eval_nodes, mono, u->d_slot_owner);
  B200_LAUNCHED();
  timer.End(kApply, t1);
  auto t2 = timer.Begin();
  const std::vector<GroupDesc> my_groups(u->groups.begin() + u->group_begin, u->groups.begin() + u->group_end);
  auto build_hist = [&](uint32_t slot_base, uint32_t slot_cap) {
    if (!my_groups.empty())
      for_each_page([&](const EllView& pv) {
        LaunchBuildHist(pv, u->d_gq, u->d_ridx[0], u->d_ridx[1], u->d_build, u->d_build_tile_seg, u->d_ts,
                        u->d_groups + u->group_begin, my_groups, u->d_hist, u->d_quant, u->num_sms, s, u->pair_hint);
      });
    // AllReduceHist: one collective per level over the level's contiguous built slots
    auto tc = timer.Begin();
    if (sliced) {  // reduce-scatter only: this rank's bin range of every built slot, summed into its own cache
      LaunchHistSliceReduce(u->d_peer_hist, u->d_peer_flags, u->p2p_rank, u->p2p_world, u->d_ts, slot_base, slot_cap,
                            ell->total_bins, slice_bin_lo, slice_bin_hi, ++u->exchange_epoch, u->num_sms, s);
    } else if (u->p2p && u->row_split) {
      const uint32_t ea = ++u->exchange_epoch, eb = ++u->exchange_epoch;
      LaunchHistAllReduce(u->d_peer_hist, u->d_peer_flags, u->p2p_rank, u->p2p_world, u->d_ts, slot_base, slot_cap,
                          ell->total_bins, ea, eb, u->d_exchange_counter, u->num_sms, s);
    } else if (u->p2p) {
      const uint32_t b0 = my_groups.empty() ? 0u : my_groups.front().bin_begin;
      const uint32_t b1 = my_groups.empty() ? 0u : my_groups.back().bin_end;
      LaunchHistExchange(u->d_peer_hist, u->d_peer_flags, u->p2p_rank, u->p2p_world, u->d_ts, slot_base, slot_cap,
                         /*use_n_build=*/true, ell->total_bins, b0, b1, ++u->exchange_epoch, u->d_exchange_counter,
                         u->num_sms, s);
    } else if (u->comm != nullptr) {
      CommAllReduceInt64(u->comm, u->d_hist + (size_t)slot_base * ell->total_bins,
                         (size_t)slot_cap * ell->total_bins * 2, s);
    }
    timer.End(kComm, tc);
  };
  build_hist(0, 1);
  timer.End(kHist, t2);
  auto t3 = timer.Begin();
  if (has_col) {
    evaluate_with_column_sampler(1, catv, true);
  } else if (has_node_sets) {
    const IcView nv = node_sets_view();
    evaluate(1, nullptr, 0, 0, nv, catv);
  } else {
    evaluate(1, u->d_feature_set, u->n_feature_set, u->n_feature_levels, IcView{nullptr, nullptr, 0}, catv);
  }
  timer.End(kEval, t3);

  const bool depthwise = u->param.grow_policy == 0;
  for (int step = 0; step < u->n_steps; ++step) {
    // upper bound on nodes expanded in this step (host-known, data independent)
    const long long cap = u->step_cap[step];
    auto ta = timer.Begin();
    if (has_cat) {
      ApplyKernel<true><<<1, kCtlThreads, 0, s>>>(view, u->pdev, u->d_quant, u->d_ts, u->d_nodes, u->d_cand, u->d_part,
                                                  u->d_part_tile_seg, u->max_nodes, mono, u->d_node_is_cat,
                                                  u->step_level_start[step], u->level_batch);
    } else {
      ApplyKernel<false><<<1, kCtlThreads, 0, s>>>(view, u->pdev, u->d_quant, u->d_ts, u->d_nodes, u->d_cand, u->d_part,
                                                   u->d_part_tile_seg, u->max_nodes, mono, nullptr,
                                                   u->step_level_start[step], u->level_batch);
    }
    B200_LAUNCHED();
    timer.End(kApply, ta);
    const bool children_possible =
        u->param.max_depth == 0 || !depthwise || u->step_level[step] + 1 < u->param.max_depth;
    if (!children_possible) continue;
    auto tp = timer.Begin();
    if (ell->paged) {  // decision bits and tile counts accumulate over the page launches
      const size_t pt = n / kPartTile + kMaxLevelNodes + 2;
      B200_CUDA(cudaMemsetAsync(u->d_flags, 0, sizeof(uint32_t) * pt * (kPartTile / 32), s));
      B200_CUDA(cudaMemsetAsync(u->d_tile_left, 0, sizeof(uint32_t) * pt, s));
    }
    for_each_page([&](const EllView& pv) {
      LaunchPartitionCount(pv, u->d_ridx[0], u->d_ridx[1], u->d_part, u->d_part_tile_seg, u->d_ts, u->d_flags,
                           u->d_tile_left, u->num_sms, s, CatRoute{u->d_bin_bits, u->cat_bin_words});
    });
    ScanKernel<<<1, kCtlThreads, 0, s>>>(u->d_ts, u->d_nodes, u->d_part, u->d_part_tile_seg, u->d_tile_left,
                                         u->d_tile_excl, u->d_tile_left_off, u->d_build, u->d_sub, u->d_eval_nodes,
                                         u->d_build_tile_seg, u->step_slot_base[step], u->step_cap[step],
                                         u->row_split ? 1 : 0, u->d_slot_owner);
    B200_LAUNCHED();
    LaunchPartitionScatter(u->d_ridx[0], u->d_ridx[1], u->d_ridx[0], u->d_ridx[1], u->d_part, u->d_part_tile_seg,
                           u->d_ts, u->d_flags, u->d_tile_left_off, u->num_sms, s);
    timer.End(kPart, tp);
    auto th = timer.Begin();
    if (u->step_reuses[step])  // a reused window holds other nodes' sums: the build accumulates into zeros
      B200_CUDA(cudaMemsetAsync(u->d_hist + (size_t)u->step_slot_base[step] * ell->total_bins, 0,
                                sizeof(b200hist_gpair64) * (size_t)2 * u->step_cap[step] * ell->total_bins, s));
    // the exchange covers the built slots: at most both children of every split when parents were lost
    build_hist(u->step_slot_base[step], u->ring_reuse ? 2 * u->step_cap[step] : u->step_cap[step]);
    timer.End(kHist, th);
    auto tsb = timer.Begin();
    LaunchSubtract(u->d_sub, u->d_ts, u->d_hist, ell->total_bins, u->num_sms, s, slice_bin_lo, slice_bin_hi);
    timer.End(kSub, tsb);
    auto te = timer.Begin();
    IcView icv{nullptr, nullptr, 0};
    if (has_col) {
      evaluate_with_column_sampler((int32_t)(2 * cap), catv, false);
      timer.End(kEval, te);
      continue;
    }
    if (has_node_sets) {
      icv = node_sets_view();
      evaluate((int32_t)(2 * cap), nullptr, 0, 0, icv, catv);
      timer.End(kEval, te);
      continue;
    }
    if (has_ic) {
      const IcDev icd{(int32_t)ell->n_features, u->ic_words, u->ic_n_sets, u->d_ic_set_ptr, u->d_ic_set_feat,
                      u->d_ic_bits, u->d_ic_list, u->d_ic_count};
      IcLevelKernel<<<1, kCtlThreads, sizeof(unsigned long long) * (size_t)u->ic_words, s>>>(
          u->d_ts, u->d_nodes, u->d_part, u->d_eval_nodes, icd, u->d_feature_set, u->n_feature_set,
          u->n_feature_levels, nullptr, 0);
      B200_LAUNCHED();
      icv = IcView{u->d_ic_list, u->d_ic_count, 0};
    }
    evaluate((int32_t)(2 * cap), u->d_feature_set, u->n_feature_set, u->n_feature_levels, icv, catv);
    timer.End(kEval, te);
  }

  auto tf = timer.Begin();
  FinalListKernel<<<1, kCtlThreads, 0, s>>>(u->d_ts, u->d_nodes, u->d_fin, u->d_fin_tile_seg, u->d_fin_tile_begin);
  B200_LAUNCHED();
  static const bool fin_tile = std::getenv("B200HIST_FINALISE_TILE") != nullptr;  // the older per-tile scatter
  const auto* gq64 = reinterpret_cast<const longlong2*>(u->d_gq);
  const int fin_grid = u->num_sms * 6;
  const CatRoute cr{u->d_bin_bits, u->cat_bin_words};
  auto finalise = [&](auto bin_tag, auto cat_tag) {
    using BinT = decltype(bin_tag);
    constexpr bool kCat = decltype(cat_tag)::value;
    if (fin_tile && !ell->paged) {
      FinaliseKernel<BinT, kCat><<<u->num_sms * 8, 256, 0, s>>>(view, u->d_ridx[0], u->d_ridx[1], gq64, u->d_nodes,
                                                                u->d_fin, u->d_fin_tile_seg, u->d_ts, position,
                                                                u->d_quant, cr);
      B200_LAUNCHED();
    } else {  // the windows of row ids cover the whole matrix, or page after page the rows that are resident
      for_each_page([&](const EllView& pv) {
        const uint32_t r_lo = ell->paged ? pv.page_lo : 0u, r_hi = ell->paged ? pv.page_hi : (uint32_t)n;
        FinaliseWindowKernel<BinT, kCat><<<fin_grid, 256, 0, s>>>(pv, u->d_ridx[0], u->d_ridx[1], gq64, u->d_nodes,
                                                                  u->d_fin, u->d_ts, position, u->d_quant, r_lo, r_hi, cr);
        B200_LAUNCHED();
      });
    }
  };
  auto finalise_bins = [&](auto cat_tag) {
    if (ell->is_sparse) finalise(uint32_t{}, cat_tag);
    else if (ell->bin_bytes == 1) finalise(uint8_t{}, cat_tag);
    else finalise(uint16_t{}, cat_tag);
  };
  if (has_cat) finalise_bins(std::true_type{});
  else finalise_bins(std::false_type{});
  timer.End(kFinal, tf);

  if (defer) {  // deferred read-back: the copies are enqueued, nobody waits; b200hist_updater_fetch_tree finishes the job
    B200_CUDA(cudaMemcpyAsync(defer->h_ts, u->d_ts, sizeof(TreeState), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(defer->h_nodes, u->d_nodes, sizeof(NodeDev) * u->max_nodes, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(defer->h_quant, u->d_quant, sizeof(QuantDev), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaEventRecord(defer->landed, s));
    defer->spans = std::move(timer.spans);
    defer->busy = true;
    u->last_position = position;
    return;
  }
  // the one read-back of the tree
  B200_CUDA(cudaMemcpyAsync(u->h_ts, u->d_ts, sizeof(TreeState), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(u->h_nodes, u->d_nodes, sizeof(NodeDev) * u->max_nodes, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaMemcpyAsync(u->h_quant, u->d_quant, sizeof(QuantDev), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  u->pair_hint = u->h_quant->h_lo_or == 0u;  // the objective decides this; it rarely changes between trees
  timer.Collect();
  u->last_row_visits = u->h_ts->row_visits;
  u->last_n_nodes = u->h_ts->n_nodes;
  u->last_max_eval_depth = u->h_ts->max_eval_depth;
  if (has_node_sets)
    B200_CUDA(cudaMemcpy(&u->last_eval_count, u->d_node_seq, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (has_col) {
    int32_t cur[2] = {0, 0};
    B200_CUDA(cudaMemcpy(cur, u->d_col_cursor, sizeof(cur), cudaMemcpyDeviceToHost));
    u->last_draw_count = cur[0] + cur[1];
  }
  u->last_position = position;
  if (out_nodes) ToAbiNodes(u->h_nodes, u->last_n_nodes, out_nodes);
  if (out_n_nodes) *out_n_nodes = u->last_n_nodes;
}

}  // namespace

extern "C" {

int b200hist_updater_create(b200hist_ellpack_t ell, const b200hist_train_param* param, b200hist_updater_t* out) {
  B200_API_BEGIN();
  B200_CHECK(ell != nullptr && param != nullptr && out != nullptr, "null argument");
  B200_CHECK(param->max_depth >= 0 && param->max_leaves >= 0, "negative max_depth / max_leaves");
  B200_CHECK(param->max_depth > 0 || param->max_leaves > 0, "max_depth = 0 requires max_leaves > 0");
  B200_CHECK(param->grow_policy == 0 || param->grow_policy == 1, "grow_policy must be 0 (depthwise) or 1 (lossguide)");
  B200_CHECK(param->max_depth <= 24, "max_depth > 24 is not supported");
  DeviceGuard guard(ell->device);
  auto* u = new b200hist_updater_s();
  u->ell = ell;
  u->device = ell->device;
  u->num_sms = NumSMs(ell->device);
  u->param = *param;
  u->pdev = ToParamDev(*param);
  long long mn = MaxNodesFor(*param);
  B200_CHECK(mn < (1ll << 26), "tree too large: bound max_depth or max_leaves");
  u->max_nodes = (int32_t)mn;
  if (param->grow_policy == 0) {
    // depth-wise: one step per level; with only max_leaves the deepest chain has max_leaves-1 levels
    u->n_steps = param->max_depth > 0 ? param->max_depth : param->max_leaves - 1;
  } else {
    u->n_steps = (u->max_nodes - 1) / 2;  // one split per step
  }
  PrepareHistKernels(u->device);
  u->groups = MakeFeatureGroups(*ell, u->device);
  u->group_begin = 0;
  u->group_end = (uint32_t)u->groups.size();
  {  // histogram slots: slot 0 = root; step s owns [base, base+cap) built + [base+cap, base+2cap) subtracted
    // Depth-wise: a level that may hold more nodes than one step expands (level_batch) takes several steps -- node
    // batching, updater_gpu_hist.cuh:34 kMaxNodeBatchSize + Driver::Pop (driver.h:84-99).  With a leaf budget the
    // batch size is part of the semantics (IsChildValid sees the leaf count after each batch), so it is the
    // reference's 1024 there; without one it is only a capacity and we take 4096.
    const bool dw = param->grow_policy == 0;
    u->level_batch = param->max_leaves > 0 ? 1024 : kMaxLevelNodes;
    if (const char* e = std::getenv("B200HIST_LEVEL_BATCH")) u->level_batch = std::max(1, std::min(std::atoi(e), kMaxLevelNodes));
    const int n_levels = u->n_steps;
    u->n_steps = 0;
    std::vector<int> step_children;
    unsigned long long total_slots = 1, max_window = 0;
    for (int st = 0; st < n_levels; ++st) {
      long long width = dw ? (1ll << std::min(st, 40)) : 1ll;
      if (param->max_leaves > 0) width = std::min<long long>(width, param->max_leaves);
      const bool children = param->max_depth == 0 || !dw || st + 1 < param->max_depth;
      const long long n_sub = dw ? (width + u->level_batch - 1) / u->level_batch : 1;
      B200_CHECK(u->n_steps + n_sub < (1ll << 22), "tree too large: bound max_depth or max_leaves");
      for (long long b = 0; b < n_sub; ++b) {
        const uint32_t cap = (uint32_t)std::min<long long>(u->level_batch, width - b * u->level_batch);
        u->step_cap.push_back(cap);
        u->step_level.push_back(st);
        u->step_level_start.push_back(b == 0 ? 1 : 0);
        step_children.push_back(children ? 1 : 0);
        if (children) { total_slots += 2ull * cap; max_window = std::max<unsigned long long>(max_window, 2ull * cap); }
        ++u->n_steps;
      }
    }
    // The cache holds a slot per possible node when that fits (nothing is ever rebuilt); otherwise it is a ring of
    // `ring` slots behind the root's: the steps' windows wrap around it and a split whose parent histogram has been
    // overwritten builds both children (DeviceHistogramStorage's overflow, histogram.cuh:72-168).  Capacity:
    // train_param.reserved[0] (= HistMakerTrainParam::max_cached_hist_node) or B200HIST_CACHE_SLOTS, else half of the
    // free device memory.
    size_t free_b = 0, total_b = 0;
    B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const unsigned long long slot_bytes = (unsigned long long)ell->total_bins * sizeof(b200hist_gpair64);
    unsigned long long ring = total_slots - 1;
    unsigned long long budget_slots = std::max<unsigned long long>(1, (unsigned long long)(free_b / 2) / slot_bytes);
    if (param->reserved[0] > 0) budget_slots = (unsigned long long)param->reserved[0];
    if (const char* e = std::getenv("B200HIST_CACHE_SLOTS")) budget_slots = std::strtoull(e, nullptr, 10);
    if (ring + 1 > budget_slots) ring = std::max<unsigned long long>(max_window, budget_slots > 1 ? budget_slots - 1 : 1);
    B200_CHECK((ring + 1) * slot_bytes < (unsigned long long)free_b,
               "histogram cache: even one step's window (" + std::to_string(((ring + 1) * slot_bytes) >> 20) +
                   " MiB) does not fit in free device memory; lower max_bin / max_depth");
    u->ring_reuse = ring + 1 < total_slots;
    unsigned long long pos = 0;
    bool wrapped = false;
    for (int st = 0; st < u->n_steps; ++st) {
      const unsigned long long w = step_children[(size_t)st] ? 2ull * u->step_cap[(size_t)st] : 0ull;
      if (pos + w > ring) { pos = 0; wrapped = true; }
      u->step_slot_base.push_back((uint32_t)(1 + pos));
      u->step_reuses.push_back(wrapped ? 1 : 0);
      pos += w;
    }
    u->n_slots = (uint32_t)(ring + 1);
    if (std::getenv("B200HIST_VERBOSE") != nullptr)
      std::fprintf(stderr, "[b200hist] updater: %llu rows, %d steps, histogram cache %u of %llu slots x %llu KiB%s, free %zu MiB\n",
                   (unsigned long long)ell->n_rows, u->n_steps, u->n_slots, total_slots, slot_bytes >> 10,
                   u->ring_reuse ? " (ring: parents may be rebuilt)" : "", free_b >> 20);
  }
  const uint64_t n = ell->n_rows;
  const size_t part_tiles = n / kPartTile + kMaxLevelNodes + 2;
  const size_t hist_tiles = n / kHistTile + 2 * kMaxLevelNodes + 2;  // a split may build both children
  const size_t fin_tiles = n / kPartTile + (size_t)u->max_nodes + 2;
  u->d_groups = DevAlloc<GroupDesc>(u->groups.size());
  B200_CUDA(cudaMemcpy(u->d_groups, u->groups.data(), sizeof(GroupDesc) * u->groups.size(), cudaMemcpyHostToDevice));
  u->d_quant = DevAlloc<QuantDev>(1);
  u->d_partials = DevAlloc<double>((size_t)4 * u->num_sms * kQuantBlocksPerSM);
  u->d_counter = DevAlloc<unsigned int>(1);
  B200_CUDA(cudaMemset(u->d_counter, 0, sizeof(unsigned int)));
  u->d_gq = DevAlloc<b200hist_gpair64>(n);
  u->d_ridx[0] = DevAlloc<uint32_t>(n);
  u->d_ridx[1] = DevAlloc<uint32_t>(n);
  u->d_position = DevAlloc<int32_t>(n);
  u->d_ts = DevAlloc<TreeState>(1);
  u->d_nodes = DevAlloc<NodeDev>(u->max_nodes);
  u->d_cand = DevAlloc<b200hist_candidate>(u->max_nodes);
  u->d_part = DevAlloc<PartEntry>(kMaxLevelNodes);
  u->d_build = DevAlloc<BuildEntry>(2 * kMaxLevelNodes);
  u->d_sub = DevAlloc<SubEntry>(kMaxLevelNodes);
  u->d_eval_nodes = DevAlloc<int32_t>(2 * kMaxLevelNodes);
  u->d_part_tile_seg = DevAlloc<uint32_t>(part_tiles);
  u->d_build_tile_seg = DevAlloc<uint32_t>(hist_tiles);
  u->d_tile_left = DevAlloc<uint32_t>(part_tiles);
  u->d_tile_excl = DevAlloc<uint32_t>(part_tiles + 1);
  u->d_tile_left_off = DevAlloc<uint32_t>(part_tiles);
  u->d_flags = DevAlloc<uint32_t>(part_tiles * (kPartTile / 32));
  u->d_fin = DevAlloc<BuildEntry>(u->max_nodes);
  u->d_fin_tile_seg = DevAlloc<uint32_t>(fin_tiles);
  u->d_fin_tile_begin = DevAlloc<uint32_t>((size_t)u->max_nodes + 1);
  u->hist_bytes = (size_t)u->n_slots * ell->total_bins * sizeof(b200hist_gpair64);
  u->d_slot_owner = DevAlloc<int32_t>(u->n_slots);
  // one allocation = histogram cache + flag array of the peer-memory exchange (one IPC handle covers both);
  // at least 4 MiB so that it is a stand-alone driver allocation whose IPC mapping starts at the same address
  u->flags_offset = (u->hist_bytes + 255) & ~size_t(255);
  {
    // ... + the evaluation scratch table: under row split the ranks push per-feature winners into each other's
    u->scratch_offset = u->flags_offset + 4096;
    const size_t scratch_bytes = sizeof(FeatCand) * (size_t)2 * kMaxLevelNodes * ell->n_features;
    const size_t alloc = std::max<size_t>(u->scratch_offset + scratch_bytes, size_t(4) << 20);
    void* p = nullptr;
    B200_CUDA(cudaMalloc(&p, alloc));
    B200_CUDA(cudaMemset(p, 0, alloc));
    u->d_hist = static_cast<b200hist_gpair64*>(p);
    u->d_scratch = reinterpret_cast<FeatCand*>(static_cast<uint8_t*>(p) + u->scratch_offset);
  }
  u->d_exchange_counter = DevAlloc<unsigned int>(1);
  B200_CUDA(cudaMemset(u->d_exchange_counter, 0, sizeof(unsigned int)));
  B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&u->h_nodes), sizeof(NodeDev) * u->max_nodes));
  B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&u->h_ts), sizeof(TreeState)));
  B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&u->h_quant), sizeof(QuantDev)));
  *out = u;
  B200_API_END();
}

int b200hist_updater_destroy(b200hist_updater_t u) {
  B200_API_BEGIN();
  if (u) {
    DeviceGuard guard(u->device);
    void* ptrs[] = {u->d_groups, u->d_quant, u->d_partials, u->d_counter, u->d_gq, u->d_ridx[0], u->d_ridx[1],
                    u->d_position, u->d_gpair_staging, u->d_ts, u->d_nodes, u->d_cand, u->d_part, u->d_build,
                    u->d_sub, u->d_eval_nodes, u->d_part_tile_seg, u->d_build_tile_seg, u->d_tile_left,
                    u->d_tile_excl, u->d_tile_left_off, u->d_flags, u->d_fin, u->d_fin_tile_seg,
                    u->d_fin_tile_begin, u->d_hist, u->d_slot_owner};
    for (void* p : ptrs) cudaFree(p);
    for (void* p : u->peer_base)
      if (p != nullptr) cudaIpcCloseMemHandle(p);
    cudaFree(u->d_peer_hist);
    cudaFree(u->d_peer_flags);
    cudaFree(u->d_peer_scratch);
    cudaFree(u->d_exchange_counter);
    if (u->d_feature_set) cudaFree(u->d_feature_set);
    for (void* p : {(void*)u->d_monotone, (void*)u->d_lower, (void*)u->d_upper})
      if (p) cudaFree(p);
    for (void* p : {(void*)u->d_ic_set_ptr, (void*)u->d_ic_set_feat, (void*)u->d_ic_bits, (void*)u->d_ic_list,
                    (void*)u->d_ic_count})
      if (p) cudaFree(p);
    for (void* p : {(void*)u->d_is_cat, (void*)u->d_cat_bits, (void*)u->d_bin_bits, (void*)u->d_node_is_cat})
      if (p) cudaFree(p);
    for (void* p : {(void*)u->d_node_sets, (void*)u->d_node_list, (void*)u->d_node_count, (void*)u->d_node_seq})
      if (p) cudaFree(p);
    if (u->d_leaf_override) cudaFree(u->d_leaf_override);
    if (u->d_gbs_scratch) cudaFree(u->d_gbs_scratch);
    for (void* p : {(void*)u->d_col_seeds, (void*)u->d_col_cursor, (void*)u->d_col_tree_set, (void*)u->d_col_level_sets,
                    (void*)u->d_col_level_drawn, (void*)u->d_col_node_list, (void*)u->d_col_node_count})
      if (p) cudaFree(p);
    cudaFreeHost(u->h_nodes);
    cudaFreeHost(u->h_ts);
    cudaFreeHost(u->h_quant);
    for (auto& slot : u->pending) {
      if (slot.landed) {
        cudaEventSynchronize(slot.landed);
        cudaEventDestroy(slot.landed);
      }
      if (slot.h_nodes) cudaFreeHost(slot.h_nodes);
      if (slot.h_ts) cudaFreeHost(slot.h_ts);
      if (slot.h_quant) cudaFreeHost(slot.h_quant);
      for (auto e : slot.events) cudaEventDestroy(e);
    }
    for (auto e : u->events) cudaEventDestroy(e);
    delete u;
  }
  B200_API_END();
}

int b200hist_updater_max_nodes(b200hist_updater_t u, int32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = u->max_nodes;
  B200_API_END();
}

int b200hist_updater_update_tree(b200hist_updater_t u, const float* gpair_dev, void* stream,
                                 b200hist_tree_node* out_nodes_host, int32_t* out_n_nodes, int32_t* out_position_dev) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && gpair_dev != nullptr, "null argument");
  DeviceGuard guard(u->device);
  static const bool verbose = std::getenv("B200HIST_VERBOSE") != nullptr;
  if (verbose) {  // wall time of the whole call and its device phases, one line per tree (diagnostics)
    const bool was = u->profiling;
    u->profiling = true;
    B200_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    const auto t0 = std::chrono::steady_clock::now();
    UpdateTreeImpl(u, gpair_dev, static_cast<cudaStream_t>(stream), out_nodes_host, out_n_nodes, out_position_dev);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "[b200hist] update_tree %.2f ms wall: quant %.2f hist %.2f (exchange %.2f) sub %.2f eval %.2f apply %.2f "
                         "part %.2f final %.2f, %llu row visits, rank %d/%d\n",
                 ms, u->last_ms[0], u->last_ms[1], u->last_ms[7], u->last_ms[2], u->last_ms[3], u->last_ms[4], u->last_ms[5],
                 u->last_ms[6], (unsigned long long)u->last_row_visits, u->p2p_rank, u->p2p_world);
    u->profiling = was;
    return 0;
  }
  UpdateTreeImpl(u, gpair_dev, static_cast<cudaStream_t>(stream), out_nodes_host, out_n_nodes, out_position_dev);
  B200_API_END();
}

int b200hist_updater_update_tree_host(b200hist_updater_t u, const float* gpair_host, void* stream,
                                      b200hist_tree_node* out_nodes_host, int32_t* out_n_nodes,
                                      int32_t* out_position_dev) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && gpair_host != nullptr, "null argument");
  DeviceGuard guard(u->device);
  auto s = static_cast<cudaStream_t>(stream);
  const size_t bytes = sizeof(float) * 2 * u->ell->n_rows;
  if (u->d_gpair_staging == nullptr) u->d_gpair_staging = DevAlloc<float>(2 * u->ell->n_rows);
  B200_CUDA(cudaMemcpyAsync(u->d_gpair_staging, gpair_host, bytes, cudaMemcpyHostToDevice, s));
  UpdateTreeImpl(u, u->d_gpair_staging, s, out_nodes_host, out_n_nodes, out_position_dev);
  B200_API_END();
}

int b200hist_updater_update_tree_async(b200hist_updater_t u, const float* gpair_dev, void* stream,
                                       int32_t* out_position_dev) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && gpair_dev != nullptr, "null argument");
  B200_CHECK(u->async_issued - u->async_fetched < 2, "two trees are in flight already: call b200hist_updater_fetch_tree");
  B200_CHECK(u->d_node_sets == nullptr && u->d_col_seeds == nullptr,
             "deferred read-back is not available with colsample_bynode (the draw count is read back per tree)");
  DeviceGuard guard(u->device);
  auto& slot = u->pending[u->async_issued & 1u];
  if (slot.landed == nullptr) {
    B200_CUDA(cudaEventCreateWithFlags(&slot.landed, cudaEventDisableTiming));
    B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&slot.h_nodes), sizeof(NodeDev) * u->max_nodes));
    B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&slot.h_ts), sizeof(TreeState)));
    B200_CUDA(cudaMallocHost(reinterpret_cast<void**>(&slot.h_quant), sizeof(QuantDev)));
  }
  UpdateTreeImpl(u, gpair_dev, static_cast<cudaStream_t>(stream), nullptr, nullptr, out_position_dev, &slot);
  ++u->async_issued;
  B200_API_END();
}

int b200hist_updater_fetch_tree(b200hist_updater_t u, b200hist_tree_node* out_nodes_host, int32_t* out_n_nodes) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  B200_CHECK(u->async_fetched < u->async_issued, "no tree in flight");
  DeviceGuard guard(u->device);
  auto& slot = u->pending[u->async_fetched & 1u];
  B200_CUDA(cudaEventSynchronize(slot.landed));
  u->pair_hint = slot.h_quant->h_lo_or == 0u;
  if (u->profiling) {
    for (auto& m : u->last_ms) m = 0;
    u->last_hist_ms.clear();
    for (auto& sp : slot.spans) {
      float ms = 0;
      B200_CUDA(cudaEventElapsedTime(&ms, sp.second.first, sp.second.second));
      u->last_ms[sp.first] += ms;
      if (sp.first == kHist) u->last_hist_ms.push_back(ms);
    }
  }
  u->last_row_visits = slot.h_ts->row_visits;
  u->last_n_nodes = slot.h_ts->n_nodes;
  u->last_max_eval_depth = slot.h_ts->max_eval_depth;
  std::memcpy(u->h_nodes, slot.h_nodes, sizeof(NodeDev) * (size_t)u->last_n_nodes);  // node_histogram & co read h_nodes
  if (out_nodes_host) ToAbiNodes(slot.h_nodes, u->last_n_nodes, out_nodes_host);
  if (out_n_nodes) *out_n_nodes = u->last_n_nodes;
  slot.busy = false;
  ++u->async_fetched;
  B200_API_END();
}

int b200hist_updater_update_prediction_cache(b200hist_updater_t u, float* preds_dev, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && preds_dev != nullptr, "null argument");
  B200_CHECK(u->last_position != nullptr, "UpdatePredictionCache called before Update (updater_gpu_hist.cu:851)");
  DeviceGuard guard(u->device);
  auto s = static_cast<cudaStream_t>(stream);
  PredCacheKernel<<<u->num_sms * 8, 256, 0, s>>>(u->last_position, u->d_nodes, u->ell->n_rows, preds_dev);
  B200_LAUNCHED();
  B200_API_END();
}

int b200hist_updater_update_prediction_cache_ex(b200hist_updater_t u, float* preds_dev, uint64_t stride,
                                                const int32_t* position_dev, const float* leaf_values_host,
                                                int32_t n_leaf_values, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && preds_dev != nullptr, "null argument");
  B200_CHECK(stride >= 1, "stride must be >= 1");
  const int32_t* position = position_dev != nullptr ? position_dev : u->last_position;
  B200_CHECK(position != nullptr, "UpdatePredictionCache called before Update (updater_gpu_hist.cu:851)");
  B200_CHECK(leaf_values_host == nullptr || (n_leaf_values >= u->last_n_nodes && n_leaf_values <= u->max_nodes),
             "leaf_values_host must hold one value per node of the last tree");
  DeviceGuard guard(u->device);
  auto s = static_cast<cudaStream_t>(stream);
  if (stride == 1 && leaf_values_host == nullptr) {
    PredCacheKernel<<<u->num_sms * 8, 256, 0, s>>>(position, u->d_nodes, u->ell->n_rows, preds_dev);
    B200_LAUNCHED();
  } else {
    if (u->d_leaf_override == nullptr) u->d_leaf_override = DevAlloc<float>((size_t)u->max_nodes);
    if (leaf_values_host != nullptr) {
      // pageable host memory: the copy is staged by the driver before the call returns, the caller may reuse its buffer
      B200_CUDA(cudaMemcpyAsync(u->d_leaf_override, leaf_values_host, sizeof(float) * (size_t)n_leaf_values,
                                cudaMemcpyHostToDevice, s));
    } else {
      GatherLeavesKernel<<<(u->max_nodes + 255) / 256, 256, 0, s>>>(u->d_nodes, u->max_nodes, u->d_leaf_override);
      B200_LAUNCHED();
    }
    PredCacheStridedKernel<<<u->num_sms * 8, 256, 0, s>>>(position, u->d_leaf_override, u->ell->n_rows, stride,
                                                          preds_dev);
    B200_LAUNCHED();
  }
  B200_API_END();
}

int b200hist_updater_last_quantiser(b200hist_updater_t u, b200hist_quantiser* out_host) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_host != nullptr, "null argument");
  DeviceGuard guard(u->device);
  QuantDev h;
  B200_CUDA(cudaMemcpy(&h, u->d_quant, sizeof(h), cudaMemcpyDeviceToHost));
  *out_host = b200hist_quantiser{h.to_fixed_g, h.to_fixed_h, h.to_float_g, h.to_float_h};
  B200_API_END();
}

int b200hist_updater_node_histogram(b200hist_updater_t u, int32_t nidx, b200hist_gpair64* out_host) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_host != nullptr, "null argument");
  B200_CHECK(nidx >= 0 && nidx < u->last_n_nodes, "node index out of range");
  DeviceGuard guard(u->device);
  const int32_t slot = u->h_nodes[nidx].hist_slot;
  B200_CHECK(slot >= 0, "node " + std::to_string(nidx) + " has no histogram");
  if (u->p2p && u->last_sliced) {
    // evaluation by feature slice: rank p holds the complete sums of ITS features' bins only -- assemble the node's
    // histogram from every rank's slice through the peer mappings
    for (int p = 0; p < u->p2p_world; ++p) {
      uint32_t lo = 0, hi = 0;
      SliceBins(*u->ell, p, u->p2p_world, &lo, &hi);
      if (hi <= lo) continue;
      const auto* base = p == u->p2p_rank ? u->d_hist : static_cast<const b200hist_gpair64*>(u->peer_base[(size_t)p]);
      B200_CUDA(cudaMemcpy(out_host + lo, base + (size_t)slot * u->ell->total_bins + lo,
                           sizeof(b200hist_gpair64) * (hi - lo), cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  B200_CUDA(cudaMemcpy(out_host, u->d_hist + (size_t)slot * u->ell->total_bins,
                       sizeof(b200hist_gpair64) * u->ell->total_bins, cudaMemcpyDeviceToHost));
  B200_API_END();
}

int b200hist_updater_node_candidate(b200hist_updater_t u, int32_t nidx, b200hist_candidate* out_host) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_host != nullptr, "null argument");
  B200_CHECK(nidx >= 0 && nidx < u->last_n_nodes, "node index out of range");
  DeviceGuard guard(u->device);
  B200_CUDA(cudaMemcpy(out_host, u->d_cand + nidx, sizeof(b200hist_candidate), cudaMemcpyDeviceToHost));
  B200_API_END();
}

int b200hist_updater_set_comm(b200hist_updater_t u, b200hist_comm_t comm) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  if (comm == nullptr) {
    u->comm = nullptr;
    u->group_begin = 0;
    u->group_end = (uint32_t)u->groups.size();
  } else {
    B200_CHECK(comm->device == u->device, "communicator lives on a different device");
    u->comm = comm;
    ShardGroups((uint32_t)u->groups.size(), comm->world, comm->rank, &u->group_begin, &u->group_end);
  }
  B200_API_END();
}

int b200hist_updater_ipc_handle(b200hist_updater_t u, uint8_t* out_handle64) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_handle64 != nullptr, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard guard(u->device);
  cudaIpcMemHandle_t h;
  B200_CUDA(cudaIpcGetMemHandle(&h, u->d_hist));
  std::memcpy(out_handle64, &h, 64);
  B200_API_END();
}

int b200hist_updater_set_peers(b200hist_updater_t u, int rank, int world, const uint8_t* handles_host) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && handles_host != nullptr, "null argument");
  B200_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank");
  DeviceGuard guard(u->device);
  B200_CHECK(!u->p2p, "peers are already set");
  u->peer_base.assign(world, nullptr);
  std::vector<longlong2*> hist(world);
  std::vector<uint32_t*> flags(world);
  for (int p = 0; p < world; ++p) {
    void* base = u->d_hist;
    if (p != rank) {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles_host + 64 * p, 64);
      B200_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
      u->peer_base[p] = base;
    }
    hist[p] = static_cast<longlong2*>(base);
    flags[p] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(base) + u->flags_offset);
  }
  std::vector<FeatCand*> scratch(world);
  for (int p = 0; p < world; ++p)
    scratch[p] = reinterpret_cast<FeatCand*>(reinterpret_cast<uint8_t*>(hist[p]) + u->scratch_offset);
  u->d_peer_hist = DevAlloc<longlong2*>(world);
  u->d_peer_flags = DevAlloc<uint32_t*>(world);
  u->d_peer_scratch = DevAlloc<FeatCand*>(world);
  B200_CUDA(cudaMemcpy(u->d_peer_hist, hist.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(u->d_peer_flags, flags.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(u->d_peer_scratch, scratch.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
  ShardGroups((uint32_t)u->groups.size(), world, rank, &u->group_begin, &u->group_end);
  u->p2p_rank = rank;
  u->p2p_world = world;
  u->p2p = true;
  u->exchange_epoch = 0;
  B200_API_END();
}

int b200hist_updater_set_tree_sampling(b200hist_updater_t u, const int32_t* feature_set_host, uint32_t n_feature_set,
                                       float subsample, uint32_t subsample_seed) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  B200_CHECK(subsample > 0.0f && subsample <= 1.0f, "subsample must be in (0, 1]");
  DeviceGuard guard(u->device);
  if (u->d_feature_set) {
    cudaFree(u->d_feature_set);
    u->d_feature_set = nullptr;
    u->n_feature_set = 0;
  }
  u->n_feature_levels = 0;
  if (feature_set_host != nullptr) {
    B200_CHECK(n_feature_set >= 1 && n_feature_set <= u->ell->n_features, "feature set size out of range");
    for (uint32_t i = 0; i < n_feature_set; ++i) {
      B200_CHECK(feature_set_host[i] >= 0 && (uint32_t)feature_set_host[i] < u->ell->n_features,
                 "feature id out of range in feature set");
      B200_CHECK(i == 0 || feature_set_host[i] > feature_set_host[i - 1], "feature set must be strictly ascending");
    }
    u->d_feature_set = DevAlloc<int32_t>(n_feature_set);
    B200_CUDA(cudaMemcpy(u->d_feature_set, feature_set_host, sizeof(int32_t) * n_feature_set, cudaMemcpyHostToDevice));
    u->n_feature_set = (int32_t)n_feature_set;
  }
  u->subsample = subsample;
  u->sample_seed = subsample_seed;
  B200_API_END();
}

int b200hist_updater_set_level_feature_sets(b200hist_updater_t u, const int32_t* sets_host, uint32_t n_levels,
                                            uint32_t n_per_level) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  DeviceGuard guard(u->device);
  if ((sets_host == nullptr || n_levels == 0) && u->n_feature_levels == 0) return 0;  // nothing to remove: a tree-level
                                                                                       // set stays in force
  if (u->d_feature_set) {
    cudaFree(u->d_feature_set);
    u->d_feature_set = nullptr;
  }
  u->n_feature_set = 0;
  u->n_feature_levels = 0;
  if (sets_host != nullptr && n_levels > 0) {
    B200_CHECK(n_per_level >= 1 && n_per_level <= u->ell->n_features, "feature set size out of range");
    for (uint32_t l = 0; l < n_levels; ++l)
      for (uint32_t i = 0; i < n_per_level; ++i) {
        const int32_t f = sets_host[(size_t)l * n_per_level + i];
        B200_CHECK(f >= 0 && (uint32_t)f < u->ell->n_features, "feature id out of range in level feature set");
        B200_CHECK(i == 0 || f > sets_host[(size_t)l * n_per_level + i - 1], "level feature sets must be strictly ascending");
      }
    const size_t total = (size_t)n_levels * n_per_level;
    u->d_feature_set = DevAlloc<int32_t>(total);
    B200_CUDA(cudaMemcpy(u->d_feature_set, sets_host, sizeof(int32_t) * total, cudaMemcpyHostToDevice));
    u->n_feature_set = (int32_t)n_per_level;
    u->n_feature_levels = (int32_t)n_levels;
  }
  B200_API_END();
}

int b200hist_updater_last_max_eval_depth(b200hist_updater_t u, int32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = u->last_max_eval_depth;
  B200_API_END();
}

int b200hist_updater_set_monotone_constraints(b200hist_updater_t u, const int32_t* constraints_host, uint32_t n) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  B200_CHECK(n <= u->ell->n_features, "more monotone constraints than features");
  DeviceGuard guard(u->device);
  for (void** p : {(void**)&u->d_monotone, (void**)&u->d_lower, (void**)&u->d_upper}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  bool any = false;
  for (uint32_t i = 0; constraints_host != nullptr && i < n; ++i) any = any || constraints_host[i] != 0;
  // TreeEvaluator sets has_constraint_ whenever the parameter is non-empty, even if every sign is 0
  // (split_evaluator.h:45-57): that alone switches the gain formula, so "non-empty" is what is kept here
  if (constraints_host != nullptr && n > 0) {
    (void)any;
    std::vector<int32_t> full(u->ell->n_features, 0);
    std::copy(constraints_host, constraints_host + n, full.begin());
    u->d_monotone = DevAlloc<int32_t>(full.size());
    B200_CUDA(cudaMemcpy(u->d_monotone, full.data(), sizeof(int32_t) * full.size(), cudaMemcpyHostToDevice));
    u->d_lower = DevAlloc<float>((size_t)u->max_nodes);
    u->d_upper = DevAlloc<float>((size_t)u->max_nodes);
  }
  B200_API_END();
}

int b200hist_updater_set_interaction_constraints(b200hist_updater_t u, const int32_t* set_ptr_host,
                                                 const int32_t* set_feat_host, uint32_t n_sets) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  DeviceGuard guard(u->device);
  for (void** p : {(void**)&u->d_ic_set_ptr, (void**)&u->d_ic_set_feat, (void**)&u->d_ic_bits, (void**)&u->d_ic_list,
                   (void**)&u->d_ic_count}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  u->ic_n_sets = 0;
  u->ic_words = 0;
  if (n_sets > 0) {
    B200_CHECK(set_ptr_host != nullptr && set_feat_host != nullptr, "null constraint sets");
    B200_CHECK(u->d_node_sets == nullptr, "interaction constraints and node feature sets cannot be combined yet");
    B200_CHECK(set_ptr_host[0] == 0, "set_ptr must start at 0");
    for (uint32_t i = 0; i < n_sets; ++i) B200_CHECK(set_ptr_host[i] <= set_ptr_host[i + 1], "set_ptr must not decrease");
    const int32_t n_feat = set_ptr_host[n_sets];
    for (int32_t i = 0; i < n_feat; ++i)
      B200_CHECK(set_feat_host[i] >= 0 && (uint32_t)set_feat_host[i] < u->ell->n_features,
                 "interaction constraint names a feature outside the matrix");
    u->ic_n_sets = (int32_t)n_sets;
    u->ic_words = (int32_t)((u->ell->n_features + 63) / 64);
    B200_CHECK((size_t)u->ic_words * sizeof(unsigned long long) <= 40 * 1024, "too many features for the level mask");
    u->d_ic_set_ptr = DevAlloc<int32_t>(n_sets + 1);
    u->d_ic_set_feat = DevAlloc<int32_t>(std::max<int32_t>(n_feat, 1));
    B200_CUDA(cudaMemcpy(u->d_ic_set_ptr, set_ptr_host, sizeof(int32_t) * (n_sets + 1), cudaMemcpyHostToDevice));
    if (n_feat > 0)
      B200_CUDA(cudaMemcpy(u->d_ic_set_feat, set_feat_host, sizeof(int32_t) * n_feat, cudaMemcpyHostToDevice));
    u->d_ic_bits = DevAlloc<unsigned long long>((size_t)u->max_nodes * u->ic_words);
    u->d_ic_list = DevAlloc<int32_t>(u->ell->n_features);
    u->d_ic_count = DevAlloc<int32_t>(2 * (size_t)kMaxLevelNodes);
  }
  B200_API_END();
}

int b200hist_updater_set_column_sampler(b200hist_updater_t u, const int32_t* tree_set_host, uint32_t n_tree_set,
                                        float colsample_bylevel, float colsample_bynode, const uint32_t* seeds_host,
                                        uint32_t n_seeds) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  DeviceGuard guard(u->device);
  // the usual caller sets a fresh stream of outputs before every tree: same shapes -> only the two copies
  if (seeds_host != nullptr && u->d_col_seeds != nullptr && (int32_t)n_seeds == u->col_n_seeds &&
      (int32_t)n_tree_set == u->col_n_tree && tree_set_host != nullptr &&
      u->col_bylevel_is_one == (colsample_bylevel == 1.0f) && u->col_bynode_is_one == (colsample_bynode == 1.0f) &&
      u->col_n_level == (colsample_bylevel == 1.0f ? (int32_t)n_tree_set
                                                    : std::max(1, (int32_t)(colsample_bylevel * (float)n_tree_set))) &&
      u->col_n_node == (colsample_bynode == 1.0f ? u->col_n_level
                                                  : std::max(1, (int32_t)(colsample_bynode * (float)u->col_n_level)))) {
    for (uint32_t i = 0; i < n_tree_set; ++i)
      B200_CHECK(tree_set_host[i] >= 0 && (uint32_t)tree_set_host[i] < u->ell->n_features, "feature id out of range");
    B200_CUDA(cudaMemcpy(u->d_col_seeds, seeds_host, sizeof(uint32_t) * n_seeds, cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(u->d_col_tree_set, tree_set_host, sizeof(int32_t) * n_tree_set, cudaMemcpyHostToDevice));
    return 0;
  }
  for (void** p : {(void**)&u->d_col_seeds, (void**)&u->d_col_cursor, (void**)&u->d_col_tree_set,
                   (void**)&u->d_col_level_sets, (void**)&u->d_col_level_drawn, (void**)&u->d_col_node_list,
                   (void**)&u->d_col_node_count}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  u->col_n_seeds = 0;
  if (seeds_host != nullptr && n_seeds > 0) {
    B200_CHECK(tree_set_host != nullptr && n_tree_set >= 1 && n_tree_set <= 8192, "tree set out of range");
    B200_CHECK(colsample_bylevel > 0.0f && colsample_bylevel <= 1.0f && colsample_bynode > 0.0f && colsample_bynode <= 1.0f,
               "colsample fractions must be in (0, 1]");
    B200_CHECK(u->d_node_sets == nullptr, "the column sampler cannot be combined with explicit node sets");
    for (uint32_t i = 0; i < n_tree_set; ++i)
      B200_CHECK(tree_set_host[i] >= 0 && (uint32_t)tree_set_host[i] < u->ell->n_features, "feature id out of range");
    // a tree consumes at most one output per possible depth and one per node
    const int32_t max_levels = u->param.max_depth > 0 ? u->param.max_depth + 1 : std::max(2, u->param.max_leaves);
    B200_CHECK((int64_t)n_seeds >= (int64_t)u->max_nodes + max_levels, "need max_nodes + max depth + 1 engine outputs");
    u->col_n_seeds = (int32_t)n_seeds;
    u->col_n_tree = (int32_t)n_tree_set;
    u->col_bylevel_is_one = colsample_bylevel == 1.0f;
    u->col_bynode_is_one = colsample_bynode == 1.0f;
    u->col_n_level = u->col_bylevel_is_one ? u->col_n_tree : std::max(1, (int32_t)(colsample_bylevel * (float)n_tree_set));
    u->col_n_node = u->col_bynode_is_one ? u->col_n_level : std::max(1, (int32_t)(colsample_bynode * (float)u->col_n_level));
    u->col_max_levels = max_levels;
    u->d_col_seeds = DevAlloc<uint32_t>(n_seeds);
    B200_CUDA(cudaMemcpy(u->d_col_seeds, seeds_host, sizeof(uint32_t) * n_seeds, cudaMemcpyHostToDevice));
    u->d_col_cursor = DevAlloc<int32_t>(2);
    u->d_col_tree_set = DevAlloc<int32_t>(n_tree_set);
    B200_CUDA(cudaMemcpy(u->d_col_tree_set, tree_set_host, sizeof(int32_t) * n_tree_set, cudaMemcpyHostToDevice));
    u->d_col_level_sets = DevAlloc<int32_t>((size_t)max_levels * u->col_n_level);
    u->d_col_level_drawn = DevAlloc<int32_t>((size_t)max_levels);
    size_t max_cap = 1;
    for (uint32_t c : u->step_cap) max_cap = std::max<size_t>(max_cap, c);
    u->d_col_node_list = DevAlloc<int32_t>(2 * max_cap * (size_t)u->col_n_node);
    u->d_col_node_count = DevAlloc<int32_t>(2 * max_cap);
  }
  B200_API_END();
}

int b200hist_updater_last_draw_count(b200hist_updater_t u, int32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = u->last_draw_count;
  B200_API_END();
}

int b200hist_updater_set_sampling_method(b200hist_updater_t u, int gradient_based) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  u->gradient_based = gradient_based != 0;
  B200_API_END();
}

int b200hist_updater_set_node_feature_sets(b200hist_updater_t u, const int32_t* sets_host, uint32_t n_sets,
                                           uint32_t n_per_set) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  DeviceGuard guard(u->device);
  for (void** p : {(void**)&u->d_node_sets, (void**)&u->d_node_list, (void**)&u->d_node_count, (void**)&u->d_node_seq}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  u->n_node_sets = u->n_per_node = 0;
  if (sets_host != nullptr && n_sets > 0) {
    B200_CHECK(u->d_ic_bits == nullptr, "node feature sets and interaction constraints cannot be combined yet");
    B200_CHECK(n_per_set >= 1 && n_per_set <= u->ell->n_features, "n_per_set out of range");
    B200_CHECK((int64_t)n_sets >= (int64_t)u->max_nodes, "one draw per possible node is needed (max_nodes sets)");
    for (size_t i = 0; i < (size_t)n_sets * n_per_set; ++i)
      B200_CHECK(sets_host[i] >= 0 && (uint32_t)sets_host[i] < u->ell->n_features, "feature id out of range");
    u->n_node_sets = (int32_t)n_sets;
    u->n_per_node = (int32_t)n_per_set;
    u->d_node_sets = DevAlloc<int32_t>((size_t)n_sets * n_per_set);
    B200_CUDA(cudaMemcpy(u->d_node_sets, sets_host, sizeof(int32_t) * (size_t)n_sets * n_per_set, cudaMemcpyHostToDevice));
    // evaluation slots of one step: twice the most nodes any step can expand (host-known, data independent)
    size_t max_cap = 1;
    for (uint32_t c : u->step_cap) max_cap = std::max<size_t>(max_cap, c);
    u->d_node_list = DevAlloc<int32_t>(2 * max_cap * n_per_set);
    u->d_node_count = DevAlloc<int32_t>(2 * (size_t)kMaxLevelNodes);
    u->d_node_seq = DevAlloc<int32_t>(1);
    B200_CUDA(cudaMemset(u->d_node_seq, 0, sizeof(int32_t)));
  }
  B200_API_END();
}

int b200hist_updater_last_eval_count(b200hist_updater_t u, int32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = u->last_eval_count;
  B200_API_END();
}

int b200hist_updater_set_categorical(b200hist_updater_t u, const uint8_t* feature_is_cat_host, uint32_t n,
                                     int32_t max_cat_to_onehot, int32_t max_cat_threshold) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  B200_CHECK(n <= u->ell->n_features, "more feature types than features");
  DeviceGuard guard(u->device);
  for (void** p : {(void**)&u->d_is_cat, (void**)&u->d_cat_bits, (void**)&u->d_bin_bits, (void**)&u->d_node_is_cat}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  u->cat_words = u->cat_bin_words = u->max_cat_bins = 0;
  bool any = false;
  for (uint32_t f = 0; feature_is_cat_host != nullptr && f < n; ++f) any = any || feature_is_cat_host[f] != 0;
  if (any) {
    B200_CHECK(max_cat_to_onehot >= 1 && max_cat_threshold >= 1, "max_cat_to_onehot and max_cat_threshold must be positive");
    std::vector<uint8_t> flags(u->ell->n_features, 0);
    int32_t max_bins = 1;
    float max_code = 0.0f;
    for (uint32_t f = 0; f < n; ++f) {
      if (!feature_is_cat_host[f]) continue;
      flags[f] = 1;
      const uint32_t b = u->ell->cut_ptrs[f], e = u->ell->cut_ptrs[f + 1];
      max_bins = std::max<int32_t>(max_bins, (int32_t)(e - b));
      for (uint32_t i = b; i < e; ++i) {
        const float c = u->ell->cut_values[i];
        // a categorical feature's cut values are its category codes (AddCategories, src/common/quantile.cc)
        B200_CHECK(c >= 0.0f && c < 16777216.0f && c == std::floor(c), "categorical cut value is not a category code");
        max_code = std::max(max_code, c);
      }
    }
    B200_CHECK(max_bins <= kMaxCatBins, "a categorical feature has more bins than the sort scratch holds");
    u->max_cat_bins = max_bins;
    u->cat_bin_words = (max_bins + 31) / 32;
    u->cat_words = (int32_t)max_code / 32 + 1;
    u->max_cat_to_onehot = max_cat_to_onehot;
    u->max_cat_threshold = max_cat_threshold;
    u->d_is_cat = DevAlloc<uint8_t>(flags.size());
    B200_CUDA(cudaMemcpy(u->d_is_cat, flags.data(), flags.size(), cudaMemcpyHostToDevice));
    u->d_cat_bits = DevAlloc<uint32_t>((size_t)u->max_nodes * u->cat_words);
    u->d_bin_bits = DevAlloc<uint32_t>((size_t)u->max_nodes * u->cat_bin_words);
    u->d_node_is_cat = DevAlloc<uint8_t>((size_t)u->max_nodes);
    B200_CUDA(cudaMemset(u->d_cat_bits, 0, sizeof(uint32_t) * (size_t)u->max_nodes * u->cat_words));
    B200_CUDA(cudaMemset(u->d_bin_bits, 0, sizeof(uint32_t) * (size_t)u->max_nodes * u->cat_bin_words));
    B200_CUDA(cudaMemset(u->d_node_is_cat, 0, (size_t)u->max_nodes));
  }
  B200_API_END();
}

int b200hist_updater_cat_words(b200hist_updater_t u, uint32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = (uint32_t)u->cat_words;
  B200_API_END();
}

int b200hist_updater_last_tree_categories(b200hist_updater_t u, int32_t* node_is_cat_host, uint32_t* cat_bits_host) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && node_is_cat_host != nullptr && cat_bits_host != nullptr, "null argument");
  B200_CHECK(u->d_is_cat != nullptr, "no categorical features configured");
  DeviceGuard guard(u->device);
  const int32_t nn = u->last_n_nodes;
  std::vector<uint8_t> flags((size_t)std::max(nn, 1));
  B200_CUDA(cudaMemcpy(flags.data(), u->d_node_is_cat, (size_t)nn, cudaMemcpyDeviceToHost));
  B200_CUDA(cudaMemcpy(cat_bits_host, u->d_cat_bits, sizeof(uint32_t) * (size_t)nn * u->cat_words, cudaMemcpyDeviceToHost));
  for (int32_t i = 0; i < nn; ++i) {
    // the flag belongs to the node's CANDIDATE; only an applied one is a split of the tree
    const bool split = !u->h_nodes[i].is_leaf && flags[i] != 0;
    node_is_cat_host[i] = split ? 1 : 0;
    if (!split)
      for (int32_t w = 0; w < u->cat_words; ++w) cat_bits_host[(size_t)i * u->cat_words + w] = 0u;
  }
  B200_API_END();
}

int b200hist_updater_set_row_split(b200hist_updater_t u, int on) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  B200_CHECK(u->p2p, "row split needs the peer-memory exchange: call b200hist_updater_set_peers first");
  u->row_split = on != 0;
  if (u->row_split) {  // every rank builds every feature group over its own rows
    u->group_begin = 0;
    u->group_end = (uint32_t)u->groups.size();
  } else {
    ShardGroups((uint32_t)u->groups.size(), u->p2p_world, u->p2p_rank, &u->group_begin, &u->group_end);
  }
  B200_API_END();
}

int b200hist_updater_clear_peers(b200hist_updater_t u) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  DeviceGuard guard(u->device);
  for (void*& p : u->peer_base) {
    if (p != nullptr) cudaIpcCloseMemHandle(p);
    p = nullptr;
  }
  if (u->d_peer_hist) cudaFree(u->d_peer_hist);
  if (u->d_peer_flags) cudaFree(u->d_peer_flags);
  if (u->d_peer_scratch) cudaFree(u->d_peer_scratch);
  u->d_peer_hist = nullptr;
  u->d_peer_flags = nullptr;
  u->d_peer_scratch = nullptr;
  u->p2p = false;
  u->row_split = false;
  u->p2p_world = 1;
  u->p2p_rank = 0;
  u->group_begin = 0;
  u->group_end = (uint32_t)u->groups.size();
  B200_API_END();
}

int b200hist_updater_num_feature_groups(b200hist_updater_t u, uint32_t* out) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out != nullptr, "null argument");
  *out = (uint32_t)u->groups.size();
  B200_API_END();
}

int b200hist_updater_set_profiling(b200hist_updater_t u, int on) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  u->profiling = on != 0;
  B200_API_END();
}

int b200hist_updater_last_comm_ms(b200hist_updater_t u, float* out_ms) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_ms != nullptr, "null argument");
  *out_ms = u->last_ms[7];
  B200_API_END();
}

int b200hist_updater_last_hist_steps(b200hist_updater_t u, float* out_ms, int32_t capacity, int32_t* out_n) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr && out_n != nullptr, "null argument");
  const int32_t n = (int32_t)u->last_hist_ms.size();
  *out_n = n;
  if (out_ms)
    for (int32_t i = 0; i < n && i < capacity; ++i) out_ms[i] = u->last_hist_ms[(size_t)i];
  B200_API_END();
}

int b200hist_updater_last_profile(b200hist_updater_t u, float* out_ms7, uint64_t* out_row_visits) {
  B200_API_BEGIN();
  B200_CHECK(u != nullptr, "null argument");
  if (out_ms7) std::memcpy(out_ms7, u->last_ms, sizeof(float) * 7);
  if (out_row_visits) *out_row_visits = u->last_row_visits;
  B200_API_END();
}

}  // extern "C"
#endif  // B200_KERNELS_ONLY
