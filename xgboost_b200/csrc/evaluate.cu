// Split evaluation: best (feature, bin, missing direction) for every child node of a tree level.
//
// Replaces (reference paths relative to /root/reference):
//   EvaluateSplitAgent::{ReduceFeature, Numerical}, LossChangeMissing   src/tree/gpu_hist/evaluate_splits.cu:17-32,90-146
//   EvaluateSplitsKernel + cub segmented reduce + entry lambda          evaluate_splits.cu:255-311,350-384,441-467
//   DeviceSplitCandidate::Update                                        src/tree/updater_gpu_common.cuh:68-83
//   SplitEvaluator::{CalcSplitGain, CalcGainGivenWeight, Divide}        src/tree/split_evaluator.h:73-148
//   CalcWeight / CalcGain / ThresholdL1                                 src/tree/param.h:215-267
//
// The arithmetic follows the reference operation for operation (fp64 ToFloatingPoint, fp32 gain via
// __fdividef == div.approx.ftz.f32, fp32 adds, tie-breaks) because split decisions must be
// bit-identical.  What differs is the organisation:
//   * FeatureScanKernel: one WARP per (node, feature) packed 8 warps to a CTA and grid-strided, instead
//     of one 32-thread CTA per pair; prefix sums are warp shuffles over int64 (exact), the arg-max is a
//     shuffle reduction with the same "max gain, lowest lane" rule as cub::ArgMax.
//   * NodeReduceKernel: one warp per node folds the per-feature winners ("later feature wins only if
//     strictly greater") and applies the parent-gain / weight post-pass -- two launches per level in
//     total instead of kernel + CUB segmented reduce + transform.
// Histograms are read straight from L2 (they were just written by the build/subtract kernels).
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ double ThresholdL1(double w, float alpha) {
  if (w > +alpha) return w - alpha;
  if (w < -alpha) return w + alpha;
  return 0.0;
}

__device__ __forceinline__ double CalcWeightD(const ParamDev& p, double g, double h) {
  if (h < p.min_child_weight || h <= 0.0) return 0.0;
  double dw = -ThresholdL1(g, p.reg_alpha) / (h + p.reg_lambda);
  if (p.max_delta_step != 0.0f && fabs(dw) > p.max_delta_step) dw = copysign((double)p.max_delta_step, dw);
  return dw;
}

__device__ __forceinline__ double CalcGainD(const ParamDev& p, double g, double h) {
  if (h < p.min_child_weight || h <= 0.0) return 0.0;
  if (p.max_delta_step == 0.0f) {
    if (p.reg_alpha == 0.0f) return (g * g) / (h + p.reg_lambda);
    double t = ThresholdL1(g, p.reg_alpha);
    return (t * t) / (h + p.reg_lambda);
  }
  double w = CalcWeightD(p, g, h);
  double ret = -(2.0 * g * w + (h + p.reg_lambda) * (w * w));
  if (p.reg_alpha == 0.0f) return ret;
  return ret + p.reg_alpha * fabs(w);
}

// SplitEvaluator::CalcGainGivenWeight without monotone constraints (split_evaluator.h:130-143)
__device__ __forceinline__ float GainOfStats(const ParamDev& p, double g, double h) {
  if (h <= 0) return 0.0f;
  if (p.max_delta_step == 0.0f) {
    double t = ThresholdL1(g, p.reg_alpha);
    return __fdividef((float)(t * t), (float)(h + p.reg_lambda));
  }
  float w = (float)CalcWeightD(p, g, h);
  float fg = (float)g, fh = (float)h;
  return -(2.0f * fg * w + (fh + p.reg_lambda) * (w * w));
}

__device__ __forceinline__ float SplitGain(const ParamDev& p, double tfg, double tfh, long long lg, long long lh,
                                           long long rg, long long rh) {
  float gl = GainOfStats(p, (double)lg * tfg, (double)lh * tfh);
  float gr = GainOfStats(p, (double)rg * tfg, (double)rh * tfh);
  return gl + gr;
}

// Monotone constraints (TreeEvaluator, split_evaluator.h:25-205): sign per feature, weight bounds per node.
// When any constraint is set EVERY gain becomes the weight-based -(2 g w + (h + lambda) w^2) with the child
// weights clamped into the bounds of the node being split, and a split whose weights violate the sign is -inf.
__device__ __forceinline__ float BoundedWeight(const ParamDev& p, double g, double h, float lower, float upper) {
  float w = (float)CalcWeightD(p, g, h);
  if (w < lower) return lower;
  if (w > upper) return upper;
  return w;
}
__device__ __forceinline__ float SplitGainMono(const ParamDev& p, double tfg, double tfh, long long lg, long long lh,
                                               long long rg, long long rh, int constraint, float lower, float upper) {
  const double dlg = (double)lg * tfg, dlh = (double)lh * tfh, drg = (double)rg * tfg, drh = (double)rh * tfh;
  const float wl = BoundedWeight(p, dlg, dlh, lower, upper), wr = BoundedWeight(p, drg, drh, lower, upper);
  const float gl = dlh <= 0 ? 0.0f : -(2.0f * (float)dlg * wl + ((float)dlh + p.reg_lambda) * (wl * wl));
  const float gr = drh <= 0 ? 0.0f : -(2.0f * (float)drg * wr + ((float)drh + p.reg_lambda) * (wr * wr));
  const float gain = gl + gr;
  if (constraint == 0) return gain;
  if (constraint > 0) return wl <= wr ? gain : -INFINITY;
  return wl >= wr ? gain : -INFINITY;
}

constexpr int kEvalWarps = 8;

// One warp per (node slot, feature position).  kIc: interaction constraints -- the feature list is the level's
// shared Query buffer image and a slot reads only its own count of it (IcView, common.cuh); n_fs is then the
// stride of the scratch table, not the length of a list.
template <bool kIc>
__global__ void __launch_bounds__(kEvalWarps * 32)
FeatureScanKernel(EllView ell, const longlong2* __restrict__ hist, const int32_t* __restrict__ eval_nodes,
                  const TreeState* __restrict__ ts, const NodeDev* __restrict__ nodes,
                  const int32_t* __restrict__ feature_set, int32_t n_fs, ParamDev p,
                  const QuantDev* __restrict__ q, FeatCand* __restrict__ scratch, MonoView mono, int32_t n_levels,
                  IcView ic, EvalShard shard) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kEvalWarps + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * kEvalWarps;
  const int n_eval = ts->n_eval;
  const long long total = (long long)n_eval * n_fs;
  const bool has_mono = mono.constraints != nullptr;
  const double tfg = q->to_float_g, tfh = q->to_float_h;
  for (long long item = warp_global; item < total; item += n_warps) {
    const int slot = (int)(item / n_fs);
    const int fpos = (int)(item - (long long)slot * n_fs);
    const int nidx = eval_nodes[slot];
    // colsample_bylevel: feature_set holds one list of n_fs ids per depth (the last row serves deeper nodes)
    const int32_t* __restrict__ fs =
        (feature_set && n_levels > 0) ? feature_set + (size_t)min(nodes[nidx].depth, n_levels - 1) * n_fs : feature_set;
    if constexpr (kIc) {
      if (fpos >= ic.count[slot]) continue;  // warp-uniform: the node reads only its own count of the buffer
      fs = ic.list + (size_t)slot * ic.stride;
    }
    const int f = fs ? fs[fpos] : fpos;
    if (!shard.Owns(f)) continue;  // warp-uniform: another rank holds this feature's complete histogram
    const longlong2* h = hist + (size_t)nodes[nidx].hist_slot * ell.total_bins;
    const long long pg = nodes[nidx].sum_g, ph = nodes[nidx].sum_h;
    const uint32_t beg = ell.cut_ptrs[f], end = ell.cut_ptrs[f + 1];
    const int constraint = has_mono ? mono.constraints[f] : 0;
    const float w_lo = has_mono ? mono.lower[nidx] : 0.0f, w_hi = has_mono ? mono.upper[nidx] : 0.0f;

    // ReduceFeature: missing = parent - sum(bins)
    long long sg = 0, sh = 0;
    for (uint32_t b = beg + lane; b < end; b += 32) {
      longlong2 v = h[b];
      sg += v.x;
      sh += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sh += __shfl_xor_sync(0xffffffffu, sh, o);
    }
    const long long mg = pg - sg, mh = ph - sh;

    float best_loss = -3.402823466e+38f;  // -FLT_MAX, DeviceSplitCandidate default
    int best_dir = 0, best_bin = -1;
    long long best_lg = 0, best_lh = 0;
    long long run_g = 0, run_h = 0;  // exclusive prefix carried across chunks
    for (uint32_t cb = beg; cb < end; cb += 32) {
      const uint32_t b = cb + lane;
      const bool active = b < end;
      longlong2 v = active ? h[b] : make_longlong2(0, 0);
      long long ig = v.x, ih = v.y;  // inclusive warp scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long tg = __shfl_up_sync(0xffffffffu, ig, o);
        long long th = __shfl_up_sync(0xffffffffu, ih, o);
        if (lane >= o) { ig += tg; ih += th; }
      }
      const long long scan_g = run_g + ig - v.x, scan_h = run_h + ih - v.y;
      run_g += __shfl_sync(0xffffffffu, ig, 31);
      run_h += __shfl_sync(0xffffffffu, ih, 31);

      float gain = -INFINITY;
      bool missing_left = true;
      if (active) {  // LossChangeMissing
        const long long lg = scan_g + mg, lh = scan_h + mh;
        float g_ml, g_mr;
        if (has_mono) {
          g_ml = SplitGainMono(p, tfg, tfh, lg, lh, pg - lg, ph - lh, constraint, w_lo, w_hi);
          g_mr = SplitGainMono(p, tfg, tfh, scan_g, scan_h, pg - scan_g, ph - scan_h, constraint, w_lo, w_hi);
        } else {
          g_ml = SplitGain(p, tfg, tfh, lg, lh, pg - lg, ph - lh);
          g_mr = SplitGain(p, tfg, tfh, scan_g, scan_h, pg - scan_g, ph - scan_h);
        }
        missing_left = g_ml > g_mr;
        gain = missing_left ? g_ml : g_mr;
      }
      // cub::ArgMax: larger value wins, equal values -> lower lane
      float bg = gain;
      int bl = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float og = __shfl_xor_sync(0xffffffffu, bg, o);
        int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (og > bg || (og == bg && ol < bl)) { bg = og; bl = ol; }
      }
      // the winning lane proposes; DeviceSplitCandidate::Update accepts or rejects
      const long long lg = missing_left ? scan_g + mg : scan_g;
      const long long lh = missing_left ? scan_h + mh : scan_h;
      const long long rh = ph - lh;
      bool accept = (lane == bl) && active && (gain > best_loss) &&
                    ((double)lh * tfh >= p.min_child_weight) && ((double)rh * tfh >= p.min_child_weight);
      unsigned m = __ballot_sync(0xffffffffu, accept);
      if (m) {
        int src = __ffs(m) - 1;
        best_loss = __shfl_sync(0xffffffffu, gain, src);
        best_dir = __shfl_sync(0xffffffffu, missing_left ? 0 : 1, src);
        best_bin = (int)cb + src;
        best_lg = __shfl_sync(0xffffffffu, lg, src);
        best_lh = __shfl_sync(0xffffffffu, lh, src);
      }
    }
    if (lane == 0) {
      FeatCand c;
      c.loss_chg = best_loss;
      c.dir = best_dir;
      c.gbin = best_bin;
      c.pad = 0;
      c.left_g = best_lg;
      c.left_h = best_lh;
      scratch[item] = c;
    }
  }
}

// One warp per node slot: fold per-feature winners in feature order, then the post-pass.
template <bool kIc>
__global__ void __launch_bounds__(kEvalWarps * 32)
NodeReduceKernel(EllView ell, const int32_t* __restrict__ eval_nodes, const TreeState* __restrict__ ts,
                 NodeDev* __restrict__ nodes, const int32_t* __restrict__ feature_set, int32_t n_fs, ParamDev p,
                 const QuantDev* __restrict__ q, const FeatCand* __restrict__ scratch,
                 b200hist_candidate* __restrict__ cand, MonoView mono, int32_t n_levels, int32_t* max_eval_depth,
                 IcView ic) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * kEvalWarps + (threadIdx.x >> 5);
  const int n_warps = gridDim.x * kEvalWarps;
  const int n_eval = ts->n_eval;
  const double tfg = q->to_float_g, tfh = q->to_float_h;
  for (int slot = warp_global; slot < n_eval; slot += n_warps) {
    const int nidx = eval_nodes[slot];
    const FeatCand* sc = scratch + (size_t)slot * n_fs;
    // Cross-feature fold.  The reference reduces with cub::DeviceSegmentedReduce::Sum and the NON-commutative
    // operator "b.loss_chg > a.loss_chg ? b : a" (evaluate_splits.cu:308-311,371-383).  CUB (CCCL 2.8, Policy600
    // scaled for the 56-byte candidate: 256 threads x 1 item) loads a segment STRIPED -- thread t folds items
    // t, t+256, t+512, ... in order, then threads are folded in thread order -- so among equal gains the winner
    // is the one with the smallest (position mod 256), then the smallest position.  For <= 256 features that is
    // "lowest feature wins" (test_evaluate_splits.cu:411-434); beyond 256 it is not, and bit-parity needs it.
    auto earlier = [](int p2, int p1) {
      return (p2 & 255) < (p1 & 255) || ((p2 & 255) == (p1 & 255) && p2 < p1);
    };
    float bl = -3.402823466e+38f;
    int bpos = 0x7fffffff;
    int n_here = n_fs;
    if constexpr (kIc) n_here = ic.count[slot];
    for (int fpos = lane; fpos < n_here; fpos += 32) {
      float l = sc[fpos].loss_chg;
      if (bpos == 0x7fffffff || l > bl || (l == bl && earlier(fpos, bpos))) { bl = l; bpos = fpos; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ol = __shfl_xor_sync(0xffffffffu, bl, o);
      int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      if (op != 0x7fffffff && (bpos == 0x7fffffff || ol > bl || (ol == bl && earlier(op, bpos)))) { bl = ol; bpos = op; }
    }
    if (lane == 0) {
      const long long pg = nodes[nidx].sum_g, ph = nodes[nidx].sum_h;
      b200hist_candidate c;
      c.loss_chg = -3.402823466e+38f;
      c.dir = 0;
      c.findex = -1;
      c.fvalue = 0.0f;
      c.left_sum = b200hist_gpair64{0, 0};
      c.right_sum = b200hist_gpair64{0, 0};
      c.split_gbin = -1;
      if (bpos != 0x7fffffff && sc[bpos].gbin >= 0) {
        FeatCand w = sc[bpos];
        const int32_t* fs = (feature_set && n_levels > 0)
                                ? feature_set + (size_t)min(nodes[nidx].depth, n_levels - 1) * n_fs
                                : feature_set;
        if constexpr (kIc) fs = ic.list + (size_t)slot * ic.stride;
        int f = fs ? fs[bpos] : bpos;
        c.loss_chg = w.loss_chg;
        c.dir = w.dir;
        c.findex = f;
        c.split_gbin = w.gbin;
        int split_gidx = w.gbin - 1;
        c.fvalue = split_gidx < (int)ell.cut_ptrs[f] ? -INFINITY : ell.cut_values[split_gidx];
        c.left_sum = b200hist_gpair64{w.left_g, w.left_h};
        c.right_sum = b200hist_gpair64{pg - w.left_g, ph - w.left_h};
      }
      // post-pass (evaluate_splits.cu:441-467)
      float parent_gain = (float)CalcGainD(p, (double)pg * tfg, (double)ph * tfh);
      c.loss_chg -= parent_gain;
      long long tg = c.left_sum.grad + c.right_sum.grad, th = c.left_sum.hess + c.right_sum.hess;
      // evaluator.CalcWeight(input.nidx, ...): clamped into the bounds of the node being split when constrained
      const float w_lo = mono.constraints ? mono.lower[nidx] : -3.402823466e+38f;
      const float w_hi = mono.constraints ? mono.upper[nidx] : 3.402823466e+38f;
      c.base_weight = BoundedWeight(p, (double)tg * tfg, (double)th * tfh, w_lo, w_hi);
      c.left_weight = BoundedWeight(p, (double)c.left_sum.grad * tfg, (double)c.left_sum.hess * tfh, w_lo, w_hi);
      c.right_weight = BoundedWeight(p, (double)c.right_sum.grad * tfg, (double)c.right_sum.hess * tfh, w_lo, w_hi);
      cand[nidx] = c;
      if (max_eval_depth) atomicMax(max_eval_depth, nodes[nidx].depth);
      // Driver::Push filter (driver.h:42-50): kRtEps = 1e-6f
      nodes[nidx].cand_open = c.loss_chg > 1e-6f ? 1 : 0;
    }
  }
}

// =============================================================================================
// Categorical features.  Replaces
//   EvaluateSplitAgent::{OneHot, Partition, PartitionUpdate}            evaluate_splits.cu:148-252
//   GPUHistEvaluator::SortHistogram (stable sort of a feature's bins by CalcWeightCat)   evaluator.cu:69-128
//   DeviceSplitCandidate::UpdateCat                                     updater_gpu_common.cuh:88-104
//   SetCategoricalSplit                                                 evaluate_splits.cu:313-348
// Organisation: FeatureScanKernel has already filled the scratch entry of every (node, feature) as if numerical;
// CatFeatureScanKernel (one warp per pair, same packing) OVERWRITES the entries of categorical features, the fold
// and post-pass of NodeReduceKernel are shared, and CatFinishKernel turns a categorical winner into its category
// set.  The reference sorts the bins of every categorical feature of every node in one device-wide
// stable_sort_by_key and keeps the order; here a warp ranks its own feature's bins in shared memory (position of
// bin i = #{w_j < w_i} + #{j < i : w_j == w_i}: exactly the stable ascending order) and the winner's order is
// recomputed by CatFinishKernel instead of being stored for all pairs.
// FeatCand of a categorical feature: gbin = global bin of the chosen category (one-hot) or its position in the
// sorted order (partition), -1 none.
// =============================================================================================
__device__ __forceinline__ float CatGain(bool has_mono, const ParamDev& p, double tfg, double tfh, long long lg,
                                         long long lh, long long rg, long long rh, int constraint, float lo, float hi) {
  return has_mono ? SplitGainMono(p, tfg, tfh, lg, lh, rg, rh, constraint, lo, hi)
                  : SplitGain(p, tfg, tfh, lg, lh, rg, rh);
}

// cub::ArgMax over the warp: larger value wins, equal values -> lower lane.  Returns the winning lane.
__device__ __forceinline__ int WarpArgMaxLane(float gain, int lane) {
  float bg = gain;
  int bl = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float og = __shfl_xor_sync(0xffffffffu, bg, o);
    int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (og > bg || (og == bg && ol < bl)) { bg = og; bl = ol; }
  }
  return bl;
}

// Stable ascending order of a feature's bins by (float)CalcWeight(bin sums): order[k] = local bin at position k.
__device__ __forceinline__ void WarpSortFeatureBins(const longlong2* __restrict__ h, uint32_t beg, int nb,
                                                    const ParamDev& p, double tfg, double tfh, float* w_s,
                                                    uint16_t* order_s, int lane) {
  for (int i = lane; i < nb; i += 32) {
    const longlong2 v = h[beg + i];
    w_s[i] = (float)CalcWeightD(p, (double)v.x * tfg, (double)v.y * tfh);
  }
  __syncwarp();
  for (int i = lane; i < nb; i += 32) {
    const float wi = w_s[i];
    int rank = 0;
    for (int j = 0; j < nb; ++j) {
      const float wj = w_s[j];
      rank += (wj < wi || (wj == wi && j < i)) ? 1 : 0;
    }
    order_s[rank] = (uint16_t)i;
  }
  __syncwarp();
}

// UpdateCat from the lane that won the chunk; all lanes receive the result.
struct CatBest {
  float loss;
  int dir, enc;
  long long lg, lh;
};
__device__ __forceinline__ void CatPropose(CatBest& best, bool proposes, float gain, int dir, int enc, long long lg,
                                           long long lh, long long rh, double tfh, const ParamDev& p) {
  const bool accept = proposes && gain > best.loss && ((double)lh * tfh >= p.min_child_weight) &&
                      ((double)rh * tfh >= p.min_child_weight);
  const unsigned m = __ballot_sync(0xffffffffu, accept);
  if (m) {
    const int src = __ffs(m) - 1;
    best.loss = __shfl_sync(0xffffffffu, gain, src);
    best.dir = dir;
    best.enc = __shfl_sync(0xffffffffu, enc, src);
    best.lg = __shfl_sync(0xffffffffu, lg, src);
    best.lh = __shfl_sync(0xffffffffu, lh, src);
  }
}

__global__ void __launch_bounds__(kEvalWarps * 32)
CatFeatureScanKernel(EllView ell, const longlong2* __restrict__ hist, const int32_t* __restrict__ eval_nodes,
                     const TreeState* __restrict__ ts, const NodeDev* __restrict__ nodes,
                     const int32_t* __restrict__ feature_set, int32_t n_fs, ParamDev p,
                     const QuantDev* __restrict__ q, FeatCand* __restrict__ scratch, MonoView mono, int32_t n_levels,
                     IcView ic, CatView cat) {
  B200_DYNAMIC_SHARED(unsigned char, s_cat_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* w_s = reinterpret_cast<float*>(s_cat_raw) + (size_t)warp * cat.max_cat_bins;
  uint16_t* order_s = reinterpret_cast<uint16_t*>(reinterpret_cast<float*>(s_cat_raw) + (size_t)kEvalWarps * cat.max_cat_bins) +
                      (size_t)warp * cat.max_cat_bins;
  const int warp_global = blockIdx.x * kEvalWarps + warp;
  const int n_warps = gridDim.x * kEvalWarps;
  const int n_eval = ts->n_eval;
  const long long total = (long long)n_eval * n_fs;
  const bool has_mono = mono.constraints != nullptr;
  const double tfg = q->to_float_g, tfh = q->to_float_h;
  for (long long item = warp_global; item < total; item += n_warps) {
    const int slot = (int)(item / n_fs);
    const int fpos = (int)(item - (long long)slot * n_fs);
    const int nidx = eval_nodes[slot];
    const int32_t* __restrict__ fs =
        (feature_set && n_levels > 0) ? feature_set + (size_t)min(nodes[nidx].depth, n_levels - 1) * n_fs : feature_set;
    if (ic.count != nullptr) {
      if (fpos >= ic.count[slot]) continue;
      fs = ic.list + (size_t)slot * ic.stride;
    }
    const int f = fs ? fs[fpos] : fpos;
    if (!cat.is_cat[f]) continue;  // warp-uniform
    const longlong2* h = hist + (size_t)nodes[nidx].hist_slot * ell.total_bins;
    const long long pg = nodes[nidx].sum_g, ph = nodes[nidx].sum_h;
    const uint32_t beg = ell.cut_ptrs[f], end = ell.cut_ptrs[f + 1];
    const int nb = (int)(end - beg);
    const int constraint = has_mono ? mono.constraints[f] : 0;
    const float w_lo = has_mono ? mono.lower[nidx] : 0.0f, w_hi = has_mono ? mono.upper[nidx] : 0.0f;
    CatBest best{-3.402823466e+38f, 0, -1, 0, 0};

    if (nb < cat.max_cat_to_onehot) {
      // OneHot: one category (with or without the missing values) against everything else
      long long sg = 0, sh = 0;
      for (uint32_t b = beg + lane; b < end; b += 32) {
        const longlong2 v = h[b];
        sg += v.x;
        sh += v.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sh += __shfl_xor_sync(0xffffffffu, sh, o);
      }
      const long long mg = pg - sg, mh = ph - sh;
      for (uint32_t cb = beg; cb < end; cb += 32) {
        const uint32_t b = cb + lane;
        const bool active = b < end;
        float gain = -INFINITY;
        bool missing_left = true;
        long long bg_ = 0, bh_ = 0;
        if (active) {
          const longlong2 v = h[b];
          bg_ = pg - v.x - mg;  // everything but this category and the missing values
          bh_ = ph - v.y - mh;
          const long long lg = bg_ + mg, lh = bh_ + mh;
          const float g_ml = CatGain(has_mono, p, tfg, tfh, lg, lh, pg - lg, ph - lh, constraint, w_lo, w_hi);
          const float g_mr = CatGain(has_mono, p, tfg, tfh, bg_, bh_, pg - bg_, ph - bh_, constraint, w_lo, w_hi);
          missing_left = g_ml > g_mr;
          gain = missing_left ? g_ml : g_mr;
        }
        const int win = WarpArgMaxLane(gain, lane);
        const long long lg = missing_left ? bg_ + mg : bg_;
        const long long lh = missing_left ? bh_ + mh : bh_;
        const int dir = __shfl_sync(0xffffffffu, missing_left ? 0 : 1, win);
        CatPropose(best, lane == win && active, gain, dir, (int)b, lg, lh, ph - lh, tfh, p);
      }
    } else {
      // Partition: scan the categories in the order of their leaf weights, from both ends
      WarpSortFeatureBins(h, beg, nb, p, tfg, tfh, w_s, order_s, lane);
      const int n_scan = min(nb, cat.max_cat_threshold);
      {  // forward: the first it+1 sorted categories go right, the rest (with the missing values) left
        const int it_end = n_scan - 1;
        long long run_g = 0, run_h = 0;
        for (int cb = 0; cb < it_end; cb += 32) {
          const int it = cb + lane;
          const bool active = it < it_end;
          const longlong2 v = active ? h[beg + order_s[it]] : make_longlong2(0, 0);
          long long ig = v.x, ih = v.y;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const long long tg = __shfl_up_sync(0xffffffffu, ig, o);
            const long long th = __shfl_up_sync(0xffffffffu, ih, o);
            if (lane >= o) { ig += tg; ih += th; }
          }
          const long long rg = run_g + ig, rh = run_h + ih;  // inclusive
          run_g += __shfl_sync(0xffffffffu, ig, 31);
          run_h += __shfl_sync(0xffffffffu, ih, 31);
          const long long lg = pg - rg, lh = ph - rh;
          const float gain = active ? CatGain(has_mono, p, tfg, tfh, lg, lh, rg, rh, constraint, w_lo, w_hi) : -INFINITY;
          const int win = WarpArgMaxLane(gain, lane);
          CatPropose(best, lane == win && active, gain, 0 /* kLeftDir */, it, lg, lh, rh, tfh, p);
        }
      }
      {  // backward: the last sorted categories go left, the rest (with the missing values) right
        const int it_begin = nb - 1, it_end = it_begin - n_scan + 1;
        long long run_g = 0, run_h = 0;
        for (int cb = it_begin; cb > it_end; cb -= 32) {
          const int it = cb - lane;
          const bool active = it > it_end;
          const longlong2 v = active ? h[beg + order_s[it]] : make_longlong2(0, 0);
          long long ig = v.x, ih = v.y;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const long long tg = __shfl_up_sync(0xffffffffu, ig, o);
            const long long th = __shfl_up_sync(0xffffffffu, ih, o);
            if (lane >= o) { ig += tg; ih += th; }
          }
          const long long lg = run_g + ig, lh = run_h + ih;  // inclusive
          run_g += __shfl_sync(0xffffffffu, ig, 31);
          run_h += __shfl_sync(0xffffffffu, ih, 31);
          const long long rh = ph - lh;
          const float gain = active ? CatGain(has_mono, p, tfg, tfh, lg, lh, pg - lg, rh, constraint, w_lo, w_hi) : -INFINITY;
          const int win = WarpArgMaxLane(gain, lane);
          CatPropose(best, lane == win && active, gain, 1 /* kRightDir */, it, lg, lh, rh, tfh, p);
        }
      }
      __syncwarp();  // the next item of this warp reuses the sort scratch
    }
    if (lane == 0) {
      FeatCand c;
      c.loss_chg = best.loss;
      c.dir = best.dir;
      c.gbin = best.enc;
      c.pad = 0;
      c.left_g = best.lg;
      c.left_h = best.lh;
      scratch[item] = c;
    }
  }
}

// One warp per evaluation slot, after NodeReduceKernel: a categorical winner gets its category set, a NaN split
// condition (evaluate_splits.cu:316-348, updater_gpu_hist.cu:631) and the flag the apply step looks at.
__global__ void __launch_bounds__(kEvalWarps * 32)
CatFinishKernel(EllView ell, const longlong2* __restrict__ hist, const int32_t* __restrict__ eval_nodes,
                const TreeState* __restrict__ ts, const NodeDev* __restrict__ nodes, ParamDev p,
                const QuantDev* __restrict__ q, b200hist_candidate* __restrict__ cand, CatView cat) {
  B200_DYNAMIC_SHARED(unsigned char, s_cat_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* w_s = reinterpret_cast<float*>(s_cat_raw) + (size_t)warp * cat.max_cat_bins;
  uint16_t* order_s = reinterpret_cast<uint16_t*>(reinterpret_cast<float*>(s_cat_raw) + (size_t)kEvalWarps * cat.max_cat_bins) +
                      (size_t)warp * cat.max_cat_bins;
  const int warp_global = blockIdx.x * kEvalWarps + warp;
  const int n_warps = gridDim.x * kEvalWarps;
  const int n_eval = ts->n_eval;
  const double tfg = q->to_float_g, tfh = q->to_float_h;
  for (int slot = warp_global; slot < n_eval; slot += n_warps) {
    const int nidx = eval_nodes[slot];
    const int f = cand[nidx].findex;
    const bool is_cat = f >= 0 && cat.is_cat[f] != 0;
    if (lane == 0) cat.node_is_cat[nidx] = is_cat ? 1 : 0;
    if (!is_cat) continue;
    uint32_t* cbits = cat.cat_bits + (size_t)nidx * cat.cat_words;
    uint32_t* bbits = cat.bin_bits + (size_t)nidx * cat.bin_words;
    for (int w = lane; w < cat.cat_words; w += 32) cbits[w] = 0u;
    for (int w = lane; w < cat.bin_words; w += 32) bbits[w] = 0u;
    __syncwarp();
    const uint32_t beg = ell.cut_ptrs[f], end = ell.cut_ptrs[f + 1];
    const int nb = (int)(end - beg);
    const int enc = cand[nidx].split_gbin;
    const int dir = cand[nidx].dir;
    if (nb < cat.max_cat_to_onehot) {
      if (lane == 0 && enc >= 0) {
        const int b = enc - (int)beg;
        const uint32_t c = (uint32_t)ell.cut_values[enc];
        if ((int)(c >> 5) < cat.cat_words) cbits[c >> 5] |= 1u << (31 - (c & 31));
        bbits[b >> 5] |= 1u << (b & 31);
      }
    } else if (enc >= 0) {
      const longlong2* h = hist + (size_t)nodes[nidx].hist_slot * ell.total_bins;
      WarpSortFeatureBins(h, beg, nb, p, tfg, tfh, w_s, order_s, lane);
      const int partition = dir == 0 ? enc + 1 : enc;
      for (int k = lane; k < partition; k += 32) {
        const int b = order_s[k];
        const uint32_t c = (uint32_t)ell.cut_values[beg + b];
        if ((int)(c >> 5) < cat.cat_words) atomicOr(&cbits[c >> 5], 1u << (31 - (c & 31)));
        atomicOr(&bbits[b >> 5], 1u << (b & 31));
      }
      __syncwarp();
    }
    __syncwarp();  // every lane has read the candidate before it changes
    if (lane == 0) {
      cand[nidx].fvalue = NAN;
      cand[nidx].split_gbin = -1;
    }
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchEvaluate(const EllView& ell, const b200hist_gpair64* hist, const int32_t* eval_nodes, const TreeState* ts,
                    int32_t max_eval, const NodeDev* nodes, const int32_t* feature_set, int32_t n_feature_set,
                    const ParamDev& p, const QuantDev* q, FeatCand* scratch, b200hist_candidate* cand,
                    NodeDev* nodes_mut, int num_sms, cudaStream_t s, MonoView mono, int32_t n_levels,
                    int32_t* max_eval_depth, IcView ic, CatView cat, EvalShard shard,
                    const std::function<void()>* between) {
  const int32_t n_fs = feature_set ? n_feature_set : (int32_t)ell.n_features;
  if (shard.world > 1) shard.n_features = ell.n_features;
  B200_CHECK(shard.world <= 1 || cat.is_cat == nullptr, "internal: evaluation by feature slice has no categorical path");
  long long max_items = (long long)max_eval * n_fs;
  if (max_items == 0) return;
  int blocks = (int)std::min<long long>((max_items + kEvalWarps - 1) / kEvalWarps, (long long)num_sms * 8);
  int blocks2 = (int)std::min<long long>(((long long)max_eval + kEvalWarps - 1) / kEvalWarps, (long long)num_sms * 4);
  const auto* h2 = reinterpret_cast<const longlong2*>(hist);
  // sort scratch of the categorical kernels: per warp max_cat_bins weights (float) + positions (uint16)
  const size_t cat_smem = (size_t)kEvalWarps * (size_t)cat.max_cat_bins * (sizeof(float) + sizeof(uint16_t));
  auto scan_cat = [&]() {
    if (cat.is_cat == nullptr) return;
    AllowDynamicShared(CatFeatureScanKernel, cat_smem);
    CatFeatureScanKernel<<<blocks, kEvalWarps * 32, cat_smem, s>>>(ell, h2, eval_nodes, ts, nodes, feature_set, n_fs, p,
                                                                   q, scratch, mono, n_levels, ic, cat);
    B200_LAUNCHED();
  };
  auto finish_cat = [&]() {
    if (cat.is_cat == nullptr) return;
    AllowDynamicShared(CatFinishKernel, cat_smem);
    CatFinishKernel<<<blocks2, kEvalWarps * 32, cat_smem, s>>>(ell, h2, eval_nodes, ts, nodes, p, q, cand, cat);
    B200_LAUNCHED();
  };
  if (ic.count != nullptr) {  // an answer is a subset of the sampled list, so n_fs bounds every count
    FeatureScanKernel<true><<<blocks, kEvalWarps * 32, 0, s>>>(ell, h2, eval_nodes, ts, nodes, feature_set, n_fs, p, q,
                                                               scratch, mono, n_levels, ic, shard);
    B200_LAUNCHED();
    scan_cat();
    if (between) (*between)();
    NodeReduceKernel<true><<<blocks2, kEvalWarps * 32, 0, s>>>(ell, eval_nodes, ts, nodes_mut, feature_set, n_fs, p, q,
                                                               scratch, cand, mono, n_levels, max_eval_depth, ic);
    B200_LAUNCHED();
    finish_cat();
    return;
  }
  FeatureScanKernel<false><<<blocks, kEvalWarps * 32, 0, s>>>(ell, h2, eval_nodes, ts, nodes, feature_set, n_fs, p, q,
                                                              scratch, mono, n_levels, ic, shard);
  B200_LAUNCHED();
  scan_cat();
  if (between) (*between)();
  NodeReduceKernel<false><<<blocks2, kEvalWarps * 32, 0, s>>>(ell, eval_nodes, ts, nodes_mut, feature_set, n_fs, p, q,
                                                              scratch, cand, mono, n_levels, max_eval_depth, ic);
  B200_LAUNCHED();
  finish_cat();
}

#endif  // B200_KERNELS_ONLY

}  // namespace b200
