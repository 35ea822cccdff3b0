/*
 * hist_oracle.c -- CPU restatement of the reference `gpu_hist` tree-updater algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (xgboost_b200/, include/) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED against (a) the reference's own known-answer tests restated as data in
 * tests/golden/eval_known_answers.json (tests/cpp/tree/gpu_hist/test_evaluate_splits.cu:292-472,
 * tests/cpp/data/test_ellpack_page.cu, test_histogram.cu:286-355) and (b) tree dumps produced by
 * the unmodified reference built from /root/reference (tools/make_golden.py -> tests/golden/).
 * The only step a CPU cannot reproduce bit-for-bit is `__fdividef` (device approximate
 * division, src/tree/split_evaluator.h:122-128); this file uses IEEE division and the tests
 * treat |delta gain| <= 4 ulp as a tie needing confirmation.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * The structure is deliberately plain sequential C: it states WHAT the reference computes,
 * not how its CUDA kernels are organised.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math -o liboracle.so hist_oracle.c -lm
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Parameters (src/tree/param.h:25-185 TrainParam; src/tree/updater_gpu_common.cuh:20-40)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t max_depth;       /* 0 = no limit */
  int32_t max_leaves;      /* 0 = no limit */
  int32_t grow_policy;     /* 0 depthwise, 1 lossguide */
  float learning_rate;
  float reg_lambda;
  float reg_alpha;
  float min_child_weight;
  float max_delta_step;
  float min_split_loss;
} orc_param;

/* int64 fixed-point gradient pair: include/xgboost/base.h:253-304 */
typedef struct {
  int64_t g, h;
} orc_gpair64;

/* split candidate: src/tree/updater_gpu_common.cuh:52-83 + expand entry expand_entry.cuh:15-34 */
typedef struct {
  float loss_chg;
  int32_t dir; /* 0 = left (missing goes left), 1 = right */
  int32_t findex;
  float fvalue;
  orc_gpair64 left_sum, right_sum;
  float base_weight, left_weight, right_weight;
  int32_t split_gbin; /* global bin index b whose scan produced the split (left = bins < b) */
} orc_candidate;

/* tree node in the layout the updater writes through RegTree::ExpandNode
 * (src/tree/tree_model.cc:805-830, updater_gpu_hist.cu:610-655,680-684) */
typedef struct {
  int32_t parent, left, right; /* -1 when absent */
  int32_t findex;
  float fvalue; /* split condition; leaf: leaf value */
  int32_t default_left;
  int32_t is_leaf;
  float leaf_value;
  float loss_chg, sum_hess, base_weight;
  int32_t depth;
} orc_node;

/* ------------------------------------------------------------------------------------------
 * Q1  HistogramCuts::SearchBin  (src/common/hist_util.h:112-121; device twin
 *     src/data/ellpack_page.cu:103-120): upper_bound over the feature's cuts, clamped to last.
 * ---------------------------------------------------------------------------------------- */
ORC_API int32_t orc_search_bin(float v, const float* cut_values, uint32_t beg, uint32_t end) {
  uint32_t lo = beg, hi = end;
  while (lo < hi) { /* first index with cut > v */
    uint32_t mid = lo + (hi - lo) / 2;
    if (!(v < cut_values[mid])) lo = mid + 1; else hi = mid;
  }
  if (lo == end) lo = end - 1;
  return (int32_t)lo;
}

/* Dense binning: X is row-major n x F float32, NaN = missing.  out[r*F+f] = GLOBAL bin index or
 * -1 when missing.  (ellpack_page.cu:56-134: bin = upper_bound; clamp; dense layouts then store
 * the feature-local bin `gbin - cut_ptrs[f]`, sparse layouts the global one.) */
ORC_API void orc_bin_dense(const float* X, uint64_t n, uint32_t F, const uint32_t* cut_ptrs,
                           const float* cut_values, int32_t* out) {
  for (uint64_t r = 0; r < n; ++r)
    for (uint32_t f = 0; f < F; ++f) {
      float v = X[r * F + f];
      out[r * F + f] = isnan(v) ? -1 : orc_search_bin(v, cut_values, cut_ptrs[f], cut_ptrs[f + 1]);
    }
}

/* ------------------------------------------------------------------------------------------
 * Q3/Q4  Reference Ellpack bit packing (src/common/compressed_iterator.h:37-40,110-152,228-268).
 * symbol_bits = max(1, ceil(log2(n_symbols))); symbols are written big-endian bit-packed
 * starting at byte kPadding=8; buffer size rounded up to 4 bytes (min kPadding+1).
 * ---------------------------------------------------------------------------------------- */
ORC_API uint32_t orc_symbol_bits(uint64_t n_symbols) {
  uint32_t bits = (uint32_t)ceil(log2((double)n_symbols));
  return bits < 1 ? 1 : bits;
}
ORC_API uint64_t orc_ellpack_buffer_size(uint64_t n_elements, uint64_t n_symbols) {
  uint64_t compressed = (uint64_t)ceil((double)(orc_symbol_bits(n_symbols) * n_elements) / 8.0);
  uint64_t ret = (uint64_t)(ceil((double)(compressed + 8) / 4.0) * 4.0);
  return ret < 9 ? 9 : ret;
}
/* AtomicWriteSymbol semantics: OR the symbol's bits MSB-first into the stream at bit offset
 * i*symbol_bits after the padding.  buffer must be zero-initialised. */
ORC_API void orc_ellpack_pack(const uint32_t* symbols, uint64_t n_elements, uint64_t n_symbols,
                              uint8_t* buffer) {
  uint32_t bits = orc_symbol_bits(n_symbols);
  for (uint64_t i = 0; i < n_elements; ++i) {
    uint64_t ibit_start = i * bits, ibit_end = (i + 1) * bits - 1;
    int64_t ibyte_start = (int64_t)(ibit_start / 8), ibyte_end = (int64_t)(ibit_end / 8);
    uint64_t sym = (uint64_t)symbols[i] << (7 - ibit_end % 8);
    for (int64_t ib = ibyte_end; ib >= ibyte_start; --ib) {
      buffer[8 + ib] |= (uint8_t)(sym & 0xff);
      sym >>= 8;
    }
  }
}
/* CompressedIterator::operator[] (compressed_iterator.h:228-268): read symbol i. */
ORC_API uint32_t orc_ellpack_read(const uint8_t* buffer, uint64_t n_symbols, uint64_t i) {
  uint32_t bits = orc_symbol_bits(n_symbols);
  uint64_t ibit_start = i * bits;
  uint64_t v = 0;
  for (uint32_t k = 0; k < bits; ++k) {
    uint64_t b = ibit_start + k;
    uint32_t bit = (buffer[8 + b / 8] >> (7 - b % 8)) & 1u;
    v = (v << 1) | bit;
  }
  return (uint32_t)v;
}

/* ------------------------------------------------------------------------------------------
 * G1  Quantiser scale (src/tree/gpu_hist/quantiser.cu:30-85, src/common/deterministic.cuh:24-38)
 * scale[0..3] = to_fixed_g, to_fixed_h, to_float_g, to_float_h.  Sums are sequential doubles
 * here; the reference uses a device tree reduction -- results differ only if the sum lands
 * within an ulp of a power-of-two boundary (SURVEY 7 "Quantiser scale parity").
 * ---------------------------------------------------------------------------------------- */
static double rounding_factor(double max_abs, uint64_t n) {
  double delta = max_abs / (1.0 - 2.0 * (double)n * DBL_EPSILON);
  int e;
  (void)frexp(delta, &e);
  return ldexp(1.0, e);
}
ORC_API void orc_quantiser_from_sums(const double sums[4] /* +g,+h,-g,-h */, uint64_t n,
                                     double scale[4]) {
  double mg = sums[0] > sums[2] ? sums[0] : sums[2];
  double mh = sums[1] > sums[3] ? sums[1] : sums[3];
  double rg = rounding_factor(mg, n), rh = rounding_factor(mh, n);
  double div = (double)((int64_t)1 << 62);
  scale[2] = rg / div;
  scale[3] = rh / div;
  scale[0] = 1.0 / scale[2];
  scale[1] = 1.0 / scale[3];
}
ORC_API void orc_quantiser(const float* gpair, uint64_t n, double scale[4], double sums_out[4]) {
  double s[4] = {0, 0, 0, 0};
  for (uint64_t i = 0; i < n; ++i) { /* Clip (quantiser.cu:30-44): float clip, double sum */
    float g = gpair[2 * i], h = gpair[2 * i + 1];
    s[0] += (double)(g > 0 ? g : 0.0f);
    s[1] += (double)(h > 0 ? h : 0.0f);
    s[2] += (double)(g < 0 ? fabsf(g) : 0.0f);
    s[3] += (double)(h < 0 ? fabsf(h) : 0.0f);
  }
  if (sums_out) memcpy(sums_out, s, sizeof(s));
  orc_quantiser_from_sums(s, n, scale);
}
/* G2  ToFixedPoint (quantiser.cuh:65-69): (int64)((double)v * to_fixed), truncation toward 0. */
ORC_API void orc_quantise(const float* gpair, uint64_t n, const double scale[4], orc_gpair64* out) {
  for (uint64_t i = 0; i < n; ++i) {
    out[i].g = (int64_t)((double)gpair[2 * i] * scale[0]);
    out[i].h = (int64_t)((double)gpair[2 * i + 1] * scale[1]);
  }
}

/* ------------------------------------------------------------------------------------------
 * H2  Histogram semantics (src/tree/gpu_hist/histogram.cu:160-183): for every row of the node and
 * every stored entry that is not null, hist[global_bin] += gq[row].  `gbins` is n_total x stride
 * GLOBAL bin indices, -1 = null.  hist is NOT cleared (accumulates).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_build_hist(const int32_t* gbins, uint32_t stride, const orc_gpair64* gq,
                            const uint32_t* ridx, uint64_t n_rows, orc_gpair64* hist) {
  for (uint64_t i = 0; i < n_rows; ++i) {
    uint32_t r = ridx ? ridx[i] : (uint32_t)i;
    const int32_t* row = gbins + (uint64_t)r * stride;
    for (uint32_t k = 0; k < stride; ++k) {
      int32_t b = row[k];
      if (b >= 0) {
        hist[b].g += gq[r].g;
        hist[b].h += gq[r].h;
      }
    }
  }
}
/* H4  SubtractionTrick (src/tree/gpu_hist/histogram.cuh:208-221) */
ORC_API void orc_subtract(const orc_gpair64* parent, const orc_gpair64* built, uint32_t n_bins,
                          orc_gpair64* out) {
  for (uint32_t b = 0; b < n_bins; ++b) {
    out[b].g = parent[b].g - built[b].g;
    out[b].h = parent[b].h - built[b].h;
  }
}
/* AtomicAdd64As32 (src/tree/gpu_hist/histogram.cuh:46-59): 64-bit add decomposed into two
 * 32-bit adds with carry -- restated for the known-answer test (test_histogram.cu:286-340). */
ORC_API int64_t orc_add64_as32(int64_t acc, int64_t val) {
  uint32_t lo = (uint32_t)((uint64_t)acc & 0xffffffffu), hi = (uint32_t)((uint64_t)acc >> 32);
  uint32_t vlo = (uint32_t)((uint64_t)val & 0xffffffffu), vhi = (uint32_t)((uint64_t)val >> 32);
  uint32_t old = lo;
  lo += vlo;
  uint32_t carry = (lo < old) ? 1u : 0u; /* old + vlo overflowed */
  hi += vhi + carry;
  return (int64_t)(((uint64_t)hi << 32) | lo);
}

/* ------------------------------------------------------------------------------------------
 * E3  Gain / weight arithmetic (src/tree/split_evaluator.h:73-148, src/tree/param.h:215-267)
 * ---------------------------------------------------------------------------------------- */
static double threshold_l1(double w, float alpha) { /* param.h:215-224, T1=double T2=float */
  if (w > +alpha) return w - alpha;
  if (w < -alpha) return w + alpha;
  return 0.0;
}
static double calc_weight_d(const orc_param* p, double g, double h) { /* param.h:233-244 */
  if (h < p->min_child_weight || h <= 0.0) return 0.0;
  double dw = -threshold_l1(g, p->reg_alpha) / (h + p->reg_lambda);
  if (p->max_delta_step != 0.0f && fabs(dw) > p->max_delta_step) dw = copysign(p->max_delta_step, dw);
  return dw;
}
static double calc_gain_d(const orc_param* p, double g, double h) { /* param.h:247-267 */
  if (h < p->min_child_weight || h <= 0.0) return 0.0;
  if (p->max_delta_step == 0.0f) {
    if (p->reg_alpha == 0.0f) return (g * g) / (h + p->reg_lambda);
    double t = threshold_l1(g, p->reg_alpha);
    return (t * t) / (h + p->reg_lambda);
  } else {
    double w = calc_weight_d(p, g, h);
    double ret = -(2.0 * g * w + (h + p->reg_lambda) * (w * w));
    if (p->reg_alpha == 0.0f) return ret;
    return ret + p->reg_alpha * fabs(w);
  }
}
/* SplitEvaluator::CalcGainGivenWeight (split_evaluator.h:130-143), no monotone constraints.
 * Device uses __fdividef; IEEE division here (see header). */
static float gain_given_stats(const orc_param* p, double g, double h) {
  if (h <= 0) return 0.0f;
  if (p->max_delta_step == 0.0f) {
    double t = threshold_l1(g, p->reg_alpha);
    float num = (float)(t * t);
    float den = (float)(h + p->reg_lambda);
    return num / den;
  }
  /* tree::CalcGainGivenWeight<ParamT,float>(p, (float)g, (float)h, w), param.h:227-230 */
  float w = (float)calc_weight_d(p, g, h);
  float fg = (float)g, fh = (float)h;
  return -(2.0f * fg * w + (fh + p->reg_lambda) * (w * w));
}
/* Monotone constraints (TreeEvaluator, split_evaluator.h:25-205): per-feature sign, per-node weight bounds. */
typedef struct {
  int has;            /* has_constraint_: switches EVERY gain to the weight-based formula */
  int constraint;     /* monotone[fidx] of the feature under evaluation */
  float lower, upper; /* bounds of the node being split */
} orc_mono;
static const orc_mono kNoMono = {0, 0, -FLT_MAX, FLT_MAX};

/* SplitEvaluator::CalcWeight (split_evaluator.h:94-112): float weight clamped into the node's bounds */
static float calc_weight_bounded(const orc_param* p, const orc_mono* m, double g, double h) {
  float w = (float)calc_weight_d(p, g, h);
  if (!m->has) return w;
  if (w < m->lower) return m->lower;
  if (w > m->upper) return m->upper;
  return w;
}
static float split_gain_m(const orc_param* p, const double scale[4], orc_gpair64 l, orc_gpair64 r, const orc_mono* m) {
  double lg = (double)l.g * scale[2], lh = (double)l.h * scale[3];
  double rg = (double)r.g * scale[2], rh = (double)r.h * scale[3];
  if (!m->has) { /* CalcSplitGain (split_evaluator.h:73-92): gain(left) + gain(right) in fp32 */
    return gain_given_stats(p, lg, lh) + gain_given_stats(p, rg, rh);
  }
  /* with constraints: CalcGainGivenWeight<ParamT,float>(p, g, h, w) = -(2 g w + (h + lambda) w^2), param.h:227-230 */
  float wl = calc_weight_bounded(p, m, lg, lh), wr = calc_weight_bounded(p, m, rg, rh);
  float gl = lh <= 0 ? 0.0f : -(2.0f * (float)lg * wl + ((float)lh + p->reg_lambda) * (wl * wl));
  float gr = rh <= 0 ? 0.0f : -(2.0f * (float)rg * wr + ((float)rh + p->reg_lambda) * (wr * wr));
  float gain = gl + gr;
  if (m->constraint == 0) return gain;
  if (m->constraint > 0) return wl <= wr ? gain : -INFINITY;
  return wl >= wr ? gain : -INFINITY;
}
static float split_gain(const orc_param* p, const double scale[4], orc_gpair64 l, orc_gpair64 r) {
  return split_gain_m(p, scale, l, r, &kNoMono);
}
static orc_gpair64 gp_add(orc_gpair64 a, orc_gpair64 b) { orc_gpair64 c = {a.g + b.g, a.h + b.h}; return c; }
static orc_gpair64 gp_sub(orc_gpair64 a, orc_gpair64 b) { orc_gpair64 c = {a.g - b.g, a.h - b.h}; return c; }

static void cand_init(orc_candidate* c) {
  memset(c, 0, sizeof(*c));
  c->loss_chg = -FLT_MAX; /* DeviceSplitCandidate default (updater_gpu_common.cuh:53-57) */
  c->dir = 0;
  c->findex = -1;
  c->fvalue = 0;
  c->split_gbin = -1;
  c->base_weight = c->left_weight = c->right_weight = NAN;
}

/* ------------------------------------------------------------------------------------------
 * E1/E2/E4  Best split of one node (src/tree/gpu_hist/evaluate_splits.cu:17-32,90-146,255-311,
 * 441-467; DeviceSplitCandidate::Update updater_gpu_common.cuh:68-83).
 * feature_set: list of feature ids to evaluate in order (NULL = all 0..F-1).
 * `gain_out` (optional, size total_bins*2): gain of every (bin, missing-left?0:1) for tie checks.
 * ---------------------------------------------------------------------------------------- */
/* One numerical feature of one node: ReduceFeature + the 32-bin chunks of EvaluateSplitAgent::Numerical
 * (evaluate_splits.cu:90-146) with DeviceSplitCandidate::Update (updater_gpu_common.cuh:68-83). */
static void eval_feature_numerical(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                                   const float* cut_values, int32_t f, const orc_param* p, const double scale[4],
                                   const orc_mono* mono, float* gain_out, orc_candidate* fbest_out) {
  uint32_t beg = cut_ptrs[f], end = cut_ptrs[f + 1];
  cand_init(fbest_out);
  /* ReduceFeature: missing = parent - sum(bins of f)   (evaluate_splits.cu:90-101) */
  orc_gpair64 fsum = {0, 0};
  for (uint32_t b = beg; b < end; ++b) fsum = gp_add(fsum, hist[b]);
  orc_gpair64 missing = gp_sub(parent_sum, fsum);
  orc_gpair64 running = {0, 0}; /* exclusive prefix carried across 32-bin chunks */
  for (uint32_t cb = beg; cb < end; cb += 32) {
    float best_gain = -INFINITY;
    int best_t = 0; /* cub::ArgMax over 32 lanes: max value, ties -> lowest lane; all -inf -> 0 */
    int best_ml = 1;
    orc_gpair64 best_scan = {0, 0};
    int any = 0;
    for (int t = 0; t < 32; ++t) {
      uint32_t b = cb + t;
      int active = b < end;
      float gain = -INFINITY;
      int ml = 1;
      orc_gpair64 scan = running;
      if (active) {
        /* LossChangeMissing (evaluate_splits.cu:17-32) */
        orc_gpair64 left = gp_add(scan, missing);
        float g_ml = split_gain_m(p, scale, left, gp_sub(parent_sum, left), mono);
        float g_mr = split_gain_m(p, scale, scan, gp_sub(parent_sum, scan), mono);
        ml = g_ml > g_mr;
        gain = ml ? g_ml : g_mr;
        if (gain_out) {
          gain_out[2 * b] = g_ml;
          gain_out[2 * b + 1] = g_mr;
        }
        running = gp_add(running, hist[b]);
      }
      if (!any || gain > best_gain) {
        if (!any || gain > best_gain) {
          best_gain = gain; best_t = t; best_ml = ml; best_scan = scan;
        }
        any = 1;
      }
    }
    /* Only the chunk winner calls Update (evaluate_splits.cu:126-142): a winner rejected by the
     * min_child_weight test means the whole chunk contributes nothing. */
    uint32_t b = cb + best_t;
    if (b < end) {
      int32_t split_gidx = (int32_t)b - 1;
      float fvalue = split_gidx < (int32_t)beg ? -INFINITY : cut_values[split_gidx];
      orc_gpair64 left = best_ml ? gp_add(best_scan, missing) : best_scan;
      orc_gpair64 right = gp_sub(parent_sum, left);
      double lh = (double)left.h * scale[3], rh = (double)right.h * scale[3];
      if (best_gain > fbest_out->loss_chg && lh >= p->min_child_weight && rh >= p->min_child_weight) {
        fbest_out->loss_chg = best_gain;
        fbest_out->dir = best_ml ? 0 : 1;
        fbest_out->fvalue = fvalue;
        fbest_out->findex = f;
        fbest_out->left_sum = left;
        fbest_out->right_sum = right;
        fbest_out->split_gbin = (int32_t)b;
      }
    }
  }
}

ORC_API void orc_eval_node_mono(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                                const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                                int32_t n_feature_set, const orc_param* p, const double scale[4],
                                orc_candidate* out, float* gain_out, const int32_t* monotone, float lower,
                                float upper);

ORC_API void orc_eval_node(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                           const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                           int32_t n_feature_set, const orc_param* p, const double scale[4],
                           orc_candidate* out, float* gain_out) {
  orc_eval_node_mono(hist, parent_sum, cut_ptrs, cut_values, n_features, feature_set, n_feature_set, p, scale, out,
                     gain_out, NULL, -FLT_MAX, FLT_MAX);
}

/* monotone (NULL = no constraints): sign per feature; lower/upper: weight bounds of THIS node */
ORC_API void orc_eval_node_mono(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                                const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                                int32_t n_feature_set, const orc_param* p, const double scale[4],
                                orc_candidate* out, float* gain_out, const int32_t* monotone, float lower,
                                float upper) {
  orc_mono mono = {monotone != NULL, 0, lower, upper};
  orc_candidate best;
  cand_init(&best);
  int32_t best_pos = -1;
  int32_t nf = feature_set ? n_feature_set : (int32_t)n_features;
  for (int32_t fi = 0; fi < nf; ++fi) {
    int32_t f = feature_set ? feature_set[fi] : fi;
    mono.constraint = monotone ? monotone[f] : 0;
    orc_candidate fbest;
    eval_feature_numerical(hist, parent_sum, cut_ptrs, cut_values, f, p, scale, &mono, gain_out, &fbest);
    /* Cross-feature fold: operator+ "b.loss_chg > a.loss_chg ? b : a" (evaluate_splits.cu:308-311) applied by
     * cub::DeviceSegmentedReduce::Sum (:371-383).  The operator is not commutative and CUB (CCCL 2.8 Policy600
     * scaled to the 56-byte candidate: 256 threads x 1 item, agent_reduce.cuh ConsumeRange) folds a segment
     * STRIPED: thread t takes positions t, t+256, ... in order, threads are then folded in thread order.  So
     * among equal gains the smallest (position mod 256), then the smallest position, wins.  Up to 256 features
     * this is "lowest feature wins" (test_evaluate_splits.cu:411-434). */
    {
      int take = 0;
      if (best_pos < 0) take = 1; /* first item of the fold (the init value is the default candidate) */
      else if (fbest.loss_chg > best.loss_chg) take = 1;
      else if (fbest.loss_chg == best.loss_chg) {
        int a = fi & 255, b = best_pos & 255;
        take = (a < b) || (a == b && fi < best_pos);
      }
      if (take) { best = fbest; best_pos = fi; }
    }
  }
  /* post-pass (evaluate_splits.cu:441-467) */
  double pg = (double)parent_sum.g * scale[2], ph = (double)parent_sum.h * scale[3];
  float parent_gain = (float)calc_gain_d(p, pg, ph);
  best.loss_chg -= parent_gain;
  orc_gpair64 tot = gp_add(best.left_sum, best.right_sum);
  /* evaluator.CalcWeight(input.nidx, ...): clamped into the bounds of the node being split (:457-463) */
  best.base_weight = calc_weight_bounded(p, &mono, (double)tot.g * scale[2], (double)tot.h * scale[3]);
  best.left_weight = calc_weight_bounded(p, &mono, (double)best.left_sum.g * scale[2], (double)best.left_sum.h * scale[3]);
  best.right_weight = calc_weight_bounded(p, &mono, (double)best.right_sum.g * scale[2], (double)best.right_sum.h * scale[3]);
  *out = best;
}

/* ------------------------------------------------------------------------------------------
 * Categorical features: one-hot and partition-based split evaluation
 *   EvaluateSplitAgent::{OneHot,Partition,PartitionUpdate}     src/tree/gpu_hist/evaluate_splits.cu:148-252
 *   kernel dispatch by feature type / UseOneHot                evaluate_splits.cu:286-299, src/common/categorical.h:86-89
 *   GPUHistEvaluator::SortHistogram (stable sort of a feature's bins by CalcWeightCat)
 *                                                              src/tree/gpu_hist/evaluator.cu:69-128
 *   DeviceSplitCandidate::UpdateCat                            src/tree/updater_gpu_common.cuh:88-104
 *   SetCategoricalSplit (category bit set of the winner)       evaluate_splits.cu:313-348
 * Pinned by the reference's literal test vectors (tests/cpp/tree/gpu_hist/test_evaluate_splits.cu:39-335,
 * tests/cpp/tree/test_evaluate_splits.h:56-92) in tests/golden/eval_categorical_known_answers.json.
 * A categorical feature's bins ARE its categories (cut_values[b] = category code of bin b).
 * ---------------------------------------------------------------------------------------- */
static void update_cat(orc_candidate* c, int32_t* c_is_cat, int32_t* c_thresh, float gain, int dir, int32_t thresh,
                       int32_t f, orc_gpair64 left, orc_gpair64 right, const orc_param* p, const double scale[4]) {
  double lh = (double)left.h * scale[3], rh = (double)right.h * scale[3];
  if (gain > c->loss_chg && lh >= p->min_child_weight && rh >= p->min_child_weight) {
    c->loss_chg = gain;
    c->dir = dir;
    c->fvalue = NAN;
    c->findex = f;
    c->left_sum = left;
    c->right_sum = right;
    c->split_gbin = -1;
    *c_is_cat = 1;
    *c_thresh = thresh;
  }
}

/* stable ascending order of the feature's bins by (float)CalcWeight(stats), evaluator.cu:83-126 */
static void sort_feature_bins(const orc_gpair64* hist, uint32_t beg, uint32_t end, const orc_param* p,
                              const double scale[4], uint32_t* order /* end - beg global bin ids */) {
  uint32_t n = end - beg;
  float* w = (float*)malloc(sizeof(float) * (n ? n : 1));
  for (uint32_t i = 0; i < n; ++i) {
    order[i] = beg + i;
    w[i] = (float)calc_weight_d(p, (double)hist[beg + i].g * scale[2], (double)hist[beg + i].h * scale[3]);
  }
  for (uint32_t i = 1; i < n; ++i) { /* insertion sort: stable */
    uint32_t oi = order[i];
    float wi = w[i];
    uint32_t j = i;
    while (j > 0 && wi < w[j - 1]) { order[j] = order[j - 1]; w[j] = w[j - 1]; --j; }
    order[j] = oi;
    w[j] = wi;
  }
  free(w);
}

static void eval_feature_categorical(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                                     const float* cut_values, int32_t f, const orc_param* p, const double scale[4],
                                     const orc_mono* mono, int32_t max_cat_to_onehot, int32_t max_cat_threshold,
                                     orc_candidate* fbest, int32_t* is_cat, int32_t* thresh, uint32_t* sorted_out) {
  uint32_t beg = cut_ptrs[f], end = cut_ptrs[f + 1];
  uint32_t n_bins_feature = end - beg;
  cand_init(fbest);
  *is_cat = 0;
  *thresh = -1;
  if (n_bins_feature < (uint32_t)max_cat_to_onehot) { /* UseOneHot */
    orc_gpair64 fsum = {0, 0};
    for (uint32_t b = beg; b < end; ++b) fsum = gp_add(fsum, hist[b]);
    orc_gpair64 missing = gp_sub(parent_sum, fsum);
    for (uint32_t cb = beg; cb < end; cb += 32) {
      float best_gain = -INFINITY;
      int best_t = 0, best_ml = 1, any = 0;
      orc_gpair64 best_bin = {0, 0};
      for (int t = 0; t < 32; ++t) {
        uint32_t b = cb + t;
        float gain = -INFINITY;
        int ml = 1;
        orc_gpair64 bin = {0, 0};
        if (b < end) { /* everything but this category (and the missing values) on one side */
          bin = gp_sub(gp_sub(parent_sum, hist[b]), missing);
          orc_gpair64 left = gp_add(bin, missing);
          float g_ml = split_gain_m(p, scale, left, gp_sub(parent_sum, left), mono);
          float g_mr = split_gain_m(p, scale, bin, gp_sub(parent_sum, bin), mono);
          ml = g_ml > g_mr;
          gain = ml ? g_ml : g_mr;
        }
        if (!any || gain > best_gain) { best_gain = gain; best_t = t; best_ml = ml; best_bin = bin; any = 1; }
      }
      uint32_t b = cb + best_t;
      if (b < end) {
        orc_gpair64 left = best_ml ? gp_add(best_bin, missing) : best_bin;
        update_cat(fbest, is_cat, thresh, best_gain, best_ml ? 0 : 1, (int32_t)cut_values[b], f, left,
                   gp_sub(parent_sum, left), p, scale);
      }
    }
    return;
  }
  /* partition-based: scan the categories in the order of their leaf weights, from both ends */
  sort_feature_bins(hist, beg, end, p, scale, sorted_out);
  int32_t n_bins = (int32_t)n_bins_feature < max_cat_threshold ? (int32_t)n_bins_feature : max_cat_threshold;
  { /* forward: the first it+1 sorted categories go right, everything else (with the missing values) left */
    int32_t it_end = n_bins - 1;
    orc_gpair64 prefix = {0, 0};
    for (int32_t cb = 0; cb < it_end; cb += 32) {
      float best_gain = -INFINITY;
      int best_t = 0, any = 0;
      orc_gpair64 best_l = {0, 0}, best_r = {0, 0};
      for (int t = 0; t < 32; ++t) {
        int32_t it = cb + t;
        float gain = -INFINITY;
        orc_gpair64 l = {0, 0}, r = {0, 0};
        if (it < it_end) {
          prefix = gp_add(prefix, hist[sorted_out[it]]);
          r = prefix;
          l = gp_sub(parent_sum, r);
          gain = split_gain_m(p, scale, l, r, mono);
        }
        if (!any || gain > best_gain) { best_gain = gain; best_t = t; best_l = l; best_r = r; any = 1; }
      }
      if (cb + best_t < it_end)
        update_cat(fbest, is_cat, thresh, best_gain, 0 /* kLeftDir */, cb + best_t, f, best_l, best_r, p, scale);
    }
  }
  { /* backward: the last sorted categories go left, everything else (with the missing values) right */
    int32_t it_begin = (int32_t)n_bins_feature - 1, it_end = it_begin - n_bins + 1;
    orc_gpair64 prefix = {0, 0};
    for (int32_t cb = it_begin; cb > it_end; cb -= 32) {
      float best_gain = -INFINITY;
      int best_t = 0, any = 0;
      orc_gpair64 best_l = {0, 0}, best_r = {0, 0};
      for (int t = 0; t < 32; ++t) {
        int32_t it = cb - t;
        float gain = -INFINITY;
        orc_gpair64 l = {0, 0}, r = {0, 0};
        if (it > it_end) {
          prefix = gp_add(prefix, hist[sorted_out[it]]);
          l = prefix;
          r = gp_sub(parent_sum, l);
          gain = split_gain_m(p, scale, l, r, mono);
        }
        if (!any || gain > best_gain) { best_gain = gain; best_t = t; best_l = l; best_r = r; any = 1; }
      }
      if (cb - best_t > it_end)
        update_cat(fbest, is_cat, thresh, best_gain, 1 /* kRightDir */, cb - best_t, f, best_l, best_r, p, scale);
    }
  }
}

/* Best split of one node over numerical and categorical features.  feature_is_cat: n_features flags.
 * Outputs next to the candidate: is_cat, thresh (one-hot: the category; partition: position in the sorted order) and
 * the category bit set of SetCategoricalSplit in LBitField32 layout (category c = bit 31 - c % 32 of word c / 32). */
static void eval_node_cat_m(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                            const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                            int32_t n_feature_set, const orc_param* p, const double scale[4],
                            const uint8_t* feature_is_cat, int32_t max_cat_to_onehot, int32_t max_cat_threshold,
                            orc_candidate* out, int32_t* out_is_cat, int32_t* out_thresh, uint32_t* out_cat_bits,
                            uint32_t n_cat_words, const int32_t* monotone, float lower, float upper);

ORC_API void orc_eval_node_cat(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                               const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                               int32_t n_feature_set, const orc_param* p, const double scale[4],
                               const uint8_t* feature_is_cat, int32_t max_cat_to_onehot, int32_t max_cat_threshold,
                               orc_candidate* out, int32_t* out_is_cat, int32_t* out_thresh, uint32_t* out_cat_bits,
                               uint32_t n_cat_words) {
  eval_node_cat_m(hist, parent_sum, cut_ptrs, cut_values, n_features, feature_set, n_feature_set, p, scale,
                  feature_is_cat, max_cat_to_onehot, max_cat_threshold, out, out_is_cat, out_thresh, out_cat_bits,
                  n_cat_words, NULL, -FLT_MAX, FLT_MAX);
}

/* monotone / lower / upper as in orc_eval_node_mono: the categorical gains go through the same evaluator
 * (evaluate_splits.cu:160-166,203-206) */
static void eval_node_cat_m(const orc_gpair64* hist, orc_gpair64 parent_sum, const uint32_t* cut_ptrs,
                            const float* cut_values, uint32_t n_features, const int32_t* feature_set,
                            int32_t n_feature_set, const orc_param* p, const double scale[4],
                            const uint8_t* feature_is_cat, int32_t max_cat_to_onehot, int32_t max_cat_threshold,
                            orc_candidate* out, int32_t* out_is_cat, int32_t* out_thresh, uint32_t* out_cat_bits,
                            uint32_t n_cat_words, const int32_t* monotone, float lower, float upper) {
  orc_mono mono = {monotone != NULL, 0, lower, upper};
  orc_candidate best;
  cand_init(&best);
  int32_t best_pos = -1, best_is_cat = 0, best_thresh = -1;
  uint32_t max_bins = 1;
  for (uint32_t f = 0; f < n_features; ++f)
    if (cut_ptrs[f + 1] - cut_ptrs[f] > max_bins) max_bins = cut_ptrs[f + 1] - cut_ptrs[f];
  uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * max_bins);
  uint32_t* best_sorted = (uint32_t*)malloc(sizeof(uint32_t) * max_bins);
  int32_t nf = feature_set ? n_feature_set : (int32_t)n_features;
  for (int32_t fi = 0; fi < nf; ++fi) {
    int32_t f = feature_set ? feature_set[fi] : fi;
    orc_candidate fbest;
    int32_t f_is_cat = 0, f_thresh = -1;
    mono.constraint = monotone ? monotone[f] : 0;
    if (feature_is_cat && feature_is_cat[f]) {
      eval_feature_categorical(hist, parent_sum, cut_ptrs, cut_values, f, p, scale, &mono, max_cat_to_onehot,
                               max_cat_threshold, &fbest, &f_is_cat, &f_thresh, sorted);
    } else {
      eval_feature_numerical(hist, parent_sum, cut_ptrs, cut_values, f, p, scale, &mono, NULL, &fbest);
    }
    int take = 0; /* striped cross-feature fold, see orc_eval_node_mono */
    if (best_pos < 0) take = 1;
    else if (fbest.loss_chg > best.loss_chg) take = 1;
    else if (fbest.loss_chg == best.loss_chg) {
      int a = fi & 255, b = best_pos & 255;
      take = (a < b) || (a == b && fi < best_pos);
    }
    if (take) {
      best = fbest; best_pos = fi; best_is_cat = f_is_cat; best_thresh = f_thresh;
      if (f_is_cat) memcpy(best_sorted, sorted, sizeof(uint32_t) * (cut_ptrs[f + 1] - cut_ptrs[f]));
    }
  }
  /* post-pass (evaluate_splits.cu:441-467) */
  double pg = (double)parent_sum.g * scale[2], ph = (double)parent_sum.h * scale[3];
  best.loss_chg -= (float)calc_gain_d(p, pg, ph);
  for (uint32_t w = 0; w < n_cat_words; ++w) out_cat_bits[w] = 0;
  if (best_is_cat) { /* SetCategoricalSplit */
    uint32_t nb = cut_ptrs[best.findex + 1] - cut_ptrs[best.findex];
    if (nb < (uint32_t)max_cat_to_onehot) {
      uint32_t c = (uint32_t)best_thresh;
      if (c / 32 < n_cat_words) out_cat_bits[c / 32] |= 1u << (31 - c % 32);
    } else if (best_thresh != -1) {
      int32_t partition = best.dir == 0 ? best_thresh + 1 : best_thresh;
      for (int32_t k = 0; k < partition; ++k) {
        uint32_t c = (uint32_t)cut_values[best_sorted[k]];
        if (c / 32 < n_cat_words) out_cat_bits[c / 32] |= 1u << (31 - c % 32);
      }
    }
  }
  orc_gpair64 tot = gp_add(best.left_sum, best.right_sum);
  best.base_weight = calc_weight_bounded(p, &mono, (double)tot.g * scale[2], (double)tot.h * scale[3]);
  best.left_weight = calc_weight_bounded(p, &mono, (double)best.left_sum.g * scale[2], (double)best.left_sum.h * scale[3]);
  best.right_weight = calc_weight_bounded(p, &mono, (double)best.right_sum.g * scale[2], (double)best.right_sum.h * scale[3]);
  *out = best;
  *out_is_cat = best_is_cat;
  *out_thresh = best_thresh;
  free(sorted);
  free(best_sorted);
}

/* ------------------------------------------------------------------------------------------
 * P2  go-left predicate (src/tree/updater_gpu_hist.cu:398-419, ellpack GetFvalue
 * src/data/ellpack_page.cuh:147-153): fvalue = cut_values[bin]; missing -> default direction;
 * else fvalue <= split_cond.
 * ---------------------------------------------------------------------------------------- */
static int32_t row_feature_gbin(const int32_t* gbins, uint32_t stride, int is_dense_layout,
                                const uint32_t* cut_ptrs, uint64_t r, int32_t f) {
  const int32_t* row = gbins + r * stride;
  if (is_dense_layout) return row[f];
  for (uint32_t k = 0; k < stride; ++k) { /* sparse ELL: BinarySearchBin over the row */
    int32_t b = row[k];
    if (b >= (int32_t)cut_ptrs[f] && b < (int32_t)cut_ptrs[f + 1]) return b;
  }
  return -1;
}
static int go_left(const int32_t* gbins, uint32_t stride, int is_dense_layout, const uint32_t* cut_ptrs,
                   const float* cut_values, uint64_t r, int32_t findex, float split_cond,
                   int default_left) {
  int32_t b = row_feature_gbin(gbins, stride, is_dense_layout, cut_ptrs, r, findex);
  if (b < 0) return default_left;
  return cut_values[b] <= split_cond;
}

/* EllpackLoader::GetElement + the numerical test of GetNextNode (src/predictor/gpu_data_accessor.cuh:95-107,
 * src/predictor/predict_fn.h): value = -inf for the feature's first bin, else the cut BELOW the bin; left iff
 * value < split_cond. */
static int go_left_predict(const int32_t* gbins, uint32_t stride, int is_dense_layout, const uint32_t* cut_ptrs,
                           const float* cut_values, uint64_t r, int32_t findex, float split_cond, int default_left) {
  int32_t b = row_feature_gbin(gbins, stride, is_dense_layout, cut_ptrs, r, findex);
  if (b < 0) return default_left;
  float v = ((uint32_t)b == cut_ptrs[findex]) ? -INFINITY : cut_values[b - 1];
  return v < split_cond;
}

/* common::Decision (src/common/categorical.h:49-60) on a training row: the category of the row's bin is looked up in
 * the split's set (LBitField32: category c = bit 31 - c % 32 of word c / 32); in the set -> RIGHT; a category beyond
 * the set's storage or a missing value -> default direction / left as the reference does. */
static int go_left_cat(const int32_t* gbins, uint32_t stride, int is_dense_layout, const uint32_t* cut_ptrs,
                       const float* cut_values, uint64_t r, int32_t findex, const uint32_t* cat_bits,
                       uint32_t n_cat_words, int default_left) {
  int32_t b = row_feature_gbin(gbins, stride, is_dense_layout, cut_ptrs, r, findex);
  if (b < 0) return default_left;
  uint32_t c = (uint32_t)cut_values[b];
  if (c / 32 >= n_cat_words) return 1;
  return ((cat_bits[c / 32] >> (31 - c % 32)) & 1u) == 0u;
}

/* Stable partition of one node's row segment: left rows first (row order inside a node does not
 * affect any result; the reference writes right rows reversed, row_partitioner.cuh:111-140). */
ORC_API uint64_t orc_partition(const int32_t* gbins, uint32_t stride, int is_dense_layout,
                               const uint32_t* cut_ptrs, const float* cut_values, const uint32_t* ridx_in,
                               uint64_t n, int32_t findex, float split_cond, int default_left,
                               uint32_t* ridx_out) {
  uint64_t nl = 0;
  for (uint64_t i = 0; i < n; ++i)
    if (go_left(gbins, stride, is_dense_layout, cut_ptrs, cut_values, ridx_in[i], findex, split_cond, default_left)) nl++;
  uint64_t l = 0, r = nl;
  for (uint64_t i = 0; i < n; ++i) {
    if (go_left(gbins, stride, is_dense_layout, cut_ptrs, cut_values, ridx_in[i], findex, split_cond, default_left))
      ridx_out[l++] = ridx_in[i];
    else
      ridx_out[r++] = ridx_in[i];
  }
  return nl;
}

/* ----------------------------------------------------------------------------------------
 * Feature interaction constraints: FeatureInteractionConstraintDevice (src/tree/constraints.cu).
 * Every node owns a bit field of allowed features (all clear at Reset, constraints.cu:134-138).
 *   Split(node, f, left, right)  (constraints.cu:269-329): node |= union of the constraint sets that
 *     hold f; left |= node; right |= node; then bit f is set in all three.  The PARENT's field changes too
 *     (it has been evaluated by then).  A feature in no set adds only itself.
 *   Query(list, nidx)  (constraints.cu:199-240): nidx == 0 returns the list untouched; otherwise the
 *     features of `list` that the node allows, in ASCENDING feature order (copy_if over a counting
 *     iterator), whatever the order of `list` was -- the position in this list is the tie rank of evaluation.
 * Sets come as CSR (set_ptr[n_sets + 1], set_feat[...]), the parsed form of the JSON parameter
 * (constraints.cu:33-44).  Bit fields here: LBitField64, bit f of word f / 64 (value 1 << (f % 64)); only
 * membership is observable, the word layout never leaves the device in the reference.
 * ---------------------------------------------------------------------------------------- */
/* Driver's max_node_batch_size (driver.h:36-40; the GPU updater passes kMaxNodeBatchSize = 1024,
 * updater_gpu_hist.cuh:34, updater_gpu_hist.cu:695).  Settable so that tests can batch small trees. */
static int32_t g_node_batch = 1024;
ORC_API void orc_set_node_batch(int32_t b) { g_node_batch = b > 0 && b <= 4096 ? b : 1024; }

typedef struct {
  int32_t n_features, n_sets;
  const int32_t* set_ptr;
  const int32_t* set_feat;
  /* 1 = model the single result buffer all Query calls of one EvaluateSplits batch share
   * (constraints.cu:125, updater_gpu_hist.cu:251-268): see orc_update_tree_ic */
  int32_t shared_result_buffer;
} orc_ic;

static int ic_words(int32_t n_features) { return (n_features + 63) / 64; }
static int ic_check(const uint64_t* bits, int32_t f) { return (int)((bits[f >> 6] >> (f & 63)) & 1u); }
static void ic_set(uint64_t* bits, int32_t f) { bits[f >> 6] |= (uint64_t)1 << (f & 63); }

ORC_API void orc_ic_split(const orc_ic* ic, uint64_t* node_bits, int32_t node, int32_t fid, int32_t left,
                          int32_t right) {
  int w = ic_words(ic->n_features);
  uint64_t* nb = node_bits + (size_t)node * w;
  uint64_t* lb = node_bits + (size_t)left * w;
  uint64_t* rb = node_bits + (size_t)right * w;
  for (int32_t s = 0; s < ic->n_sets; ++s) { /* RestoreFeatureListFromSetsKernel */
    int holds = 0;
    for (int32_t k = ic->set_ptr[s]; k < ic->set_ptr[s + 1]; ++k) holds |= (ic->set_feat[k] == fid);
    if (!holds) continue;
    for (int32_t k = ic->set_ptr[s]; k < ic->set_ptr[s + 1]; ++k) ic_set(nb, ic->set_feat[k]);
  }
  for (int i = 0; i < w; ++i) { lb[i] |= nb[i]; rb[i] |= nb[i]; } /* InteractionConstraintSplitKernel */
  ic_set(nb, fid); ic_set(lb, fid); ic_set(rb, fid);
}

/* QueryNode (constraints.cu:162-183): the node's allowed features, ascending */
ORC_API int32_t orc_ic_query_node(const orc_ic* ic, const uint64_t* node_bits, int32_t node, int32_t* out) {
  const uint64_t* nb = node_bits + (size_t)node * ic_words(ic->n_features);
  int32_t n = 0;
  for (int32_t f = 0; f < ic->n_features; ++f) if (ic_check(nb, f)) out[n++] = f;
  return n;
}

/* Query: list == NULL means all features 0..n_features-1 */
ORC_API int32_t orc_ic_query(const orc_ic* ic, const uint64_t* node_bits, const int32_t* list, int32_t n_list,
                             int32_t node, int32_t* out) {
  if (node == 0) {
    if (!list) { for (int32_t f = 0; f < ic->n_features; ++f) out[f] = f; return ic->n_features; }
    memcpy(out, list, sizeof(int32_t) * (size_t)n_list);
    return n_list;
  }
  int w = ic_words(ic->n_features);
  const uint64_t* nb = node_bits + (size_t)node * w;
  uint64_t* in = (uint64_t*)calloc((size_t)w, sizeof(uint64_t));
  if (!list) for (int32_t f = 0; f < ic->n_features; ++f) ic_set(in, f);
  else for (int32_t k = 0; k < n_list; ++k) ic_set(in, list[k]);
  int32_t n = 0;
  for (int32_t f = 0; f < ic->n_features; ++f) if (ic_check(nb, f) && ic_check(in, f)) out[n++] = f;
  free(in);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * T1  Whole tree: GPUHistMakerDevice::UpdateTree (src/tree/updater_gpu_hist.cu:693-732) with
 * InitRoot :657-691, ApplySplit :610-655, PartitionAndBuildHist :440-506, Driver driver.h:29-107,
 * GPUExpandEntry::IsValid expand_entry.cuh:33-50, FinalisePosition :529-580.
 *
 * Outputs: nodes[max_nodes] (returns number of nodes), position[n_rows] (leaf id, bitwise-NOT
 * when the row's quantised hessian is 0: updater_gpu_hist.cuh:48-56), optional per-node
 * histograms dump (hist_dump: max_nodes x total_bins, zero where not computed) and per-node
 * candidates (cand_dump[max_nodes]).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t nidx, depth;
  orc_candidate c;
  int open;
} entry_t;

static int entry_valid(const entry_t* e, const orc_param* p, int32_t num_leaves) {
  /* GPUExpandEntry::IsValid (expand_entry.cuh:33-50); kRtEps = 1e-6f (base.h) */
  if (e->c.loss_chg <= 1e-6f) return 0;
  if (e->c.left_sum.h == 0 || e->c.right_sum.h == 0) return 0;
  if (e->c.loss_chg < p->min_split_loss) return 0;
  if (p->max_depth > 0 && e->depth == p->max_depth) return 0;
  if (p->max_leaves > 0 && num_leaves == p->max_leaves) return 0;
  return 1;
}

/* Uniform row sampling: cuda_impl::UniformSampling::Sample (src/tree/gpu_hist/sampler.cu:160-173) zeroes the
 * QUANTISED gradient pair of row i when BernoulliTrial(seed, subsample)(i) holds, i.e. when
 * RandomWeight(seed)(i) > subsample (sampler.cu:36-60).  RandomWeight is thrust::default_random_engine
 * (= minstd_rand: x <- 48271 x mod 2^31-1, seeded with seed mod m, 0 -> 1), discard(i), then
 * thrust::uniform_real_distribution<float>: float(x - 1) / (1.0f + float(max - min)) with max - min = 2^31 - 3,
 * which rounds to 2^31 in float (thrust/random/detail/uniform_real_distribution.inl, CUDA 12.9 toolkit). */
ORC_API float orc_random_weight(uint32_t seed, uint64_t i) {
  const uint64_t m = 2147483647ull;
  uint64_t x = seed % m;
  if (x == 0) x = 1;
  /* discard(i) then one draw = multiply by a^(i+1) */
  uint64_t z = i + 1, mult = 48271ull, acc = 1;
  while (z > 0) {
    if (z & 1) acc = (acc * mult) % m;
    z >>= 1;
    mult = (mult * mult) % m;
  }
  x = (acc * x) % m;
  float r = (float)(uint32_t)(x - 1);
  r /= (1.0f + (float)(uint32_t)(m - 1 - 1));
  return r * (1.0f - 0.0f) + 0.0f;
}

/* thrust::shuffle as the reference's column sampler calls it (src/common/random.cu:86-91: thrust::shuffle on the device
 * with thrust::default_random_engine seeded by one ctx->Rng()() draw): the toolkit's generic algorithm
 * (thrust/system/detail/generic/shuffle.inl, CUDA 12.9): a 24-round Feistel bijection over [0, 2^bits), bits =
 * max(4, ceil(log2 m)), whose round keys are the engine's next 24 outputs; out = the inputs at the images below m, in
 * index order.  minstd_rand: x <- 48271 x mod (2^31 - 1), seeded with seed mod m (0 -> 1); a key is the uint32 state.
 * Pinned by tests/golden/thrust_shuffle_known_answers.json (the toolkit's own host backend). */
ORC_API void orc_thrust_shuffle(uint32_t seed, uint64_t m, uint32_t* perm /* m entries: out[i] = index of the input */) {
  const uint64_t mod = 2147483647ull;
  uint64_t x = seed % mod;
  if (x == 0) x = 1;
  uint32_t key[24];
  for (int i = 0; i < 24; ++i) { x = (x * 48271ull) % mod; key[i] = (uint32_t)x; }
  uint64_t bits = 4;
  if (m > 16) { uint64_t t = m - 1; bits = 0; while (t != 0) { bits++; t >>= 1; } }
  const uint64_t left_bits = bits / 2, right_bits = bits - left_bits;
  const uint64_t left_mask = (1ull << left_bits) - 1, right_mask = (1ull << right_bits) - 1;
  uint64_t count = 0;
  for (uint64_t idx = 0; idx < (1ull << bits); ++idx) {
    uint32_t s0 = (uint32_t)(idx >> right_bits), s1 = (uint32_t)(idx & right_mask);
    for (int i = 0; i < 24; ++i) {
      uint64_t product = 0xD2B74407B1CE6E93ull * (uint64_t)s0;
      uint32_t hi = (uint32_t)(product >> 32), lo = (uint32_t)product;
      lo = (lo << (right_bits - left_bits)) | (s1 >> left_bits);
      s0 = ((hi ^ key[i]) ^ s1) & (uint32_t)left_mask;
      s1 = lo & (uint32_t)right_mask;
    }
    uint64_t image = ((uint64_t)s0 << right_bits) | (uint64_t)s1;
    if (image < m) perm[count++] = (uint32_t)image;
  }
}

/* ColumnSampler::ColSample on CUDA without feature weights (src/common/random.cc:27-46, random.cu:62-96): n_sel =
 * max(1, int(colsample * size)); shuffle a copy of the features with the seed, keep the first n_sel, sort ascending.
 * Returns n_sel. */
ORC_API int32_t orc_col_sample(const int32_t* features, int32_t n, float colsample, uint32_t seed, int32_t* out) {
  if (colsample == 1.0f) { memcpy(out, features, sizeof(int32_t) * (size_t)n); return n; }
  int32_t n_sel = (int32_t)(colsample * (float)n);
  if (n_sel < 1) n_sel = 1;
  uint32_t* perm = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
  orc_thrust_shuffle(seed, (uint64_t)n, perm);
  for (int32_t i = 0; i < n_sel; ++i) out[i] = features[perm[i]];
  for (int32_t i = 1; i < n_sel; ++i) { /* ascending */
    int32_t v = out[i], j = i;
    while (j > 0 && out[j - 1] > v) { out[j] = out[j - 1]; --j; }
    out[j] = v;
  }
  free(perm);
  return n_sel;
}

/* Gradient-based row sampling: cuda_impl::GradientBasedSampling::Sample (src/tree/gpu_hist/sampler.cu:98-115,
 * 257-312; helpers src/tree/hist/sampler.h:48-74, src/tree/gpu_hist/quantiser.cuh:22-47) on the QUANTISED pairs of one
 * target, in place:
 *   rag_i  = sqrtf(float(g_i^2 + 0.1 * h_i^2))   with (g, h) = the pair converted back to double (MvsGradOp)
 *   thr    = sort(rag) ++ [FLT_MAX]
 *   csum_i = float(ToFloatingPoint(sum_{k<=i} int64(double(thr_k) * to_fixed)))   (FloatQuantiser of sum(thr), n)
 *   i*     = first i with denom = float(sample_rows) - float(n - i - 1) > 0 and thr_i < u = csum_i / denom <= thr_{i+1};
 *            thr_{i*+1} := u;  mu = thr[i* + 1]   (no such i: min_element picks index 0, mu = thr[1])
 *   row i: pair (0,0) stays; p = rag_i / mu (|mu| < 1e-6 -> 1e-6); p >= 1 keeps the pair; otherwise the pair becomes
 *          ToFixedPoint(pair_as_double * (1.0f / p)) when RandomWeight(seed)(i) <= p and (0,0) when not.
 * sample_rows = (uint64)(n * subsample) with the product taken in float (bst_idx_t = n_samples * subsample_).
 * PINNED (round 2): no literal vector of the reference covers this (its tests are statistical,
 * tests/cpp/tree/gpu_hist/test_sampler.cu:79-179), so the unmodified reference's gpu_hist was run on a B200 with
 * sampling_method=gradient_based and its trees committed (tests/golden/ref_gpu_next_trees.json, generator
 * tools/make_golden_gpu_next.py); tests/test_reference_golden.py::test_oracle_reproduces_reference_gpu_gradient_based_sampling
 * rebuilds them bit for bit from this function.  The double sum behind the quantiser is taken sequentially here (device
 * reduction order there: matters only at a power of two).
 * Returns mu. */
ORC_API float orc_gradient_based_sample(orc_gpair64* gq, uint64_t n, const double scale[4], float subsample,
                                        uint32_t seed, uint64_t* threshold_index_out, float* rag_out) {
  float* rag = (float*)malloc(sizeof(float) * (n ? n : 1));
  float* thr = (float*)malloc(sizeof(float) * (n + 1));
  for (uint64_t i = 0; i < n; ++i) {
    double g = (double)gq[i].g * scale[2], h = (double)gq[i].h * scale[3];
    float m = (float)(g * g + (double)0.1f * (h * h)); /* Sqr(g) + lambda * Sqr(h): float lambda promoted to double */
    rag[i] = sqrtf(m);
    thr[i] = rag[i];
  }
  thr[n] = FLT_MAX;
  /* ascending sort (thrust::sort of floats: any correct sort gives this array) */
  for (uint64_t gap = n / 2; gap > 0; gap /= 2) /* shell sort keeps the oracle dependency-free */
    for (uint64_t i = gap; i < n; ++i) {
      float v = thr[i];
      uint64_t j = i;
      while (j >= gap && thr[j - gap] > v) { thr[j] = thr[j - gap]; j -= gap; }
      thr[j] = v;
    }
  double total = 0.0;
  for (uint64_t i = 0; i < n; ++i) total += (double)thr[i];
  double to_float = rounding_factor(total, n) / (double)((int64_t)1 << 62);
  double to_fixed = 1.0 / to_float;
  uint64_t sample_rows = (uint64_t)((float)n * subsample);
  int64_t run = 0;
  uint64_t best = 0;
  int found = 0;
  for (uint64_t i = 0; i < n; ++i) {
    run += (int64_t)((double)thr[i] * to_fixed);
    float csum = (float)((double)run * to_float);
    if (found) continue;
    float denom = (float)sample_rows - (float)(n - i - 1);
    if (denom <= 0) continue;
    float u = csum / denom;
    if (u > thr[i] && u <= thr[i + 1]) { thr[i + 1] = u; best = i; found = 1; }
  }
  uint64_t tidx = (found ? best : 0) + 1;
  float mu = n ? thr[tidx] : FLT_MAX;
  if (threshold_index_out) *threshold_index_out = tidx;
  for (uint64_t i = 0; i < n; ++i) {
    if (gq[i].g == 0 && gq[i].h == 0) continue;
    float u = mu;
    if (fabsf(u) < 1e-6f) u = copysignf(1e-6f, u);
    float pr = rag[i] / u;
    if (pr >= 1.0f) continue;
    if (orc_random_weight(seed, i) <= pr) {
      float inv = 1.0f / pr;
      double g = ((double)gq[i].g * scale[2]) * (double)inv, h = ((double)gq[i].h * scale[3]) * (double)inv;
      gq[i].g = (int64_t)(g * scale[0]);
      gq[i].h = (int64_t)(h * scale[1]);
    } else {
      gq[i].g = 0;
      gq[i].h = 0;
    }
  }
  if (rag_out) memcpy(rag_out, rag, sizeof(float) * n);
  free(rag);
  free(thr);
  return mu;
}

ORC_API int32_t orc_update_tree_ex(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                   uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                   const float* gpair, const orc_param* p, const double* scale_override,
                                   orc_node* nodes, int32_t max_nodes, int32_t* position,
                                   orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                   const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                   uint32_t sample_seed, const int32_t* monotone, int32_t n_levels);
ORC_API int32_t orc_update_tree_ic(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                   uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                   const float* gpair, const orc_param* p, const double* scale_override,
                                   orc_node* nodes, int32_t max_nodes, int32_t* position,
                                   orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                   const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                   uint32_t sample_seed, const int32_t* monotone, int32_t n_levels,
                                   const orc_ic* ic);
typedef struct {
  const uint8_t* feature_is_cat;
  int32_t max_cat_to_onehot, max_cat_threshold;
  uint32_t n_cat_words;
  int32_t* node_is_cat;     /* out [max_nodes] */
  uint32_t* node_cat_bits;  /* out [max_nodes][n_cat_words] */
  /* colsample_bynode (ColumnSampler::GetFeatureSet, src/common/random.h:156-175): node_sets (NULL = none) is the
   * tree's stream of per-call draws, [n_node_sets][n_per_node] ascending feature ids; gpu_hist asks once for the root
   * and then for the left and right child of every valid candidate in Pop order (updater_gpu_hist.cu:217,257-264),
   * so the k-th node evaluated takes draw k.  feature_is_cat may be NULL when only this part is used. */
  const int32_t* node_sets;
  int32_t n_node_sets, n_per_node;
  int32_t* eval_count; /* out: draws used */
  int32_t gradient_based; /* subsample < 1: TrainParam::sampling_method == kGradientBased (orc_gradient_based_sample) */
  /* The column sampler driven by the engine's outputs (ColumnSampler::GetFeatureSet, src/common/random.h:156-175, with
   * ColSample on CUDA = orc_col_sample): seeds (NULL = off) are the values ctx->Rng()() returns AFTER ColumnSampler::Init,
   * tree_set the tree-level set Init produced.  A call for a node of depth d first draws the level's set if the depth
   * has none yet (colsample_bylevel < 1: one seed), then the node's set from it (colsample_bynode < 1: one seed).
   * Calls happen in evaluation order.  draws_used (out) = seeds consumed. */
  const uint32_t* seeds;
  int32_t n_seeds;
  const int32_t* tree_set;
  int32_t n_tree_set;
  float colsample_bylevel, colsample_bynode;
  int32_t* draws_used;
} orc_cat;
ORC_API int32_t orc_update_tree_full(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                     uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                     const float* gpair, const orc_param* p, const double* scale_override,
                                     orc_node* nodes, int32_t max_nodes, int32_t* position,
                                     orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                     const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                     uint32_t sample_seed, const int32_t* monotone, int32_t n_levels,
                                     const orc_ic* ic, const orc_cat* cat);

ORC_API int32_t orc_update_tree(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                const float* gpair, const orc_param* p, const double* scale_override,
                                orc_node* nodes, int32_t max_nodes, int32_t* position,
                                orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4]) {
  return orc_update_tree_ex(gbins, n_rows, stride, is_dense_layout, n_features, cut_ptrs, cut_values, gpair, p,
                            scale_override, nodes, max_nodes, position, hist_dump, cand_dump, scale_out, NULL, 0, 1.0f,
                            0u, NULL, 0);
}

/* feature_set (sorted ascending, NULL = all): the tree's colsample_bytree draw, evaluated in this order
 * (column_sampler_->GetFeatureSet, updater_gpu_hist.cu:217,257-261); subsample < 1: uniform row sampling with the
 * tree's seed (one ctx->Rng()() draw per tree, sampler.cu:165). */
ORC_API int32_t orc_update_tree_ex(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                   uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                   const float* gpair, const orc_param* p, const double* scale_override,
                                   orc_node* nodes, int32_t max_nodes, int32_t* position,
                                   orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                   const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                   uint32_t sample_seed, const int32_t* monotone, int32_t n_levels) {
  return orc_update_tree_ic(gbins, n_rows, stride, is_dense_layout, n_features, cut_ptrs, cut_values, gpair, p,
                            scale_override, nodes, max_nodes, position, hist_dump, cand_dump, scale_out, feature_set,
                            n_feature_set, subsample, sample_seed, monotone, n_levels, NULL);
}

/* ic (NULL = none): interaction constraints.  A node other than the root is evaluated on
 * Query(GetFeatureSet(depth), nidx) (updater_gpu_hist.cu:217-220,257-264) and every applied split calls
 * Split(nidx, findex, left, right) (updater_gpu_hist.cu:653-654).
 * ic->shared_result_buffer: every Query of the reference writes its answer to the front of ONE device buffer and
 * returns a span over it (constraints.cu:125,199-240); EvaluateSplits collects the spans of the whole batch
 * (left0, right0, left1, ...) before the evaluation kernel runs, so a node reads the buffer as the LAST query left
 * it, cut to its own count.  0 restates what Query returns at the moment of the call (what the unit tests of the
 * reference pin); 1 restates what the evaluation kernel then reads -- and that is what the reference's trees follow:
 * tests/golden/ref_gpu_interaction_trees.json (oracle/_ref_cuda on a B200, overlapping sets) is rebuilt bit for bit
 * with 1 and not with 0 (tests/test_reference_golden.py). */
ORC_API int32_t orc_update_tree_ic(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                   uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                   const float* gpair, const orc_param* p, const double* scale_override,
                                   orc_node* nodes, int32_t max_nodes, int32_t* position,
                                   orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                   const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                   uint32_t sample_seed, const int32_t* monotone, int32_t n_levels,
                                   const orc_ic* ic) {
  return orc_update_tree_full(gbins, n_rows, stride, is_dense_layout, n_features, cut_ptrs, cut_values, gpair, p,
                              scale_override, nodes, max_nodes, position, hist_dump, cand_dump, scale_out, feature_set,
                              n_feature_set, subsample, sample_seed, monotone, n_levels, ic, NULL);
}

/* cat (NULL = none): categorical features.  Evaluation through eval_node_cat_m; an applied categorical split stores
 * NaN as its condition (updater_gpu_hist.cu:628-642) and routes rows by its category set (go_left_cat) in the
 * partitioner and in FinalisePosition.  Outputs: cat->node_is_cat[max_nodes], cat->node_cat_bits[max_nodes][n_cat_words]
 * (set only for applied categorical splits). */
ORC_API int32_t orc_update_tree_full(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                                     uint32_t n_features, const uint32_t* cut_ptrs, const float* cut_values,
                                     const float* gpair, const orc_param* p, const double* scale_override,
                                     orc_node* nodes, int32_t max_nodes, int32_t* position,
                                     orc_gpair64* hist_dump, orc_candidate* cand_dump, double scale_out[4],
                                     const int32_t* feature_set, int32_t n_feature_set, float subsample,
                                     uint32_t sample_seed, const int32_t* monotone, int32_t n_levels,
                                     const orc_ic* ic, const orc_cat* cat) {
  /* candidates' category sets, kept until the candidate is applied */
  const int has_cat = cat && cat->feature_is_cat;
  uint32_t* cand_bits = has_cat ? (uint32_t*)calloc((size_t)max_nodes * cat->n_cat_words, sizeof(uint32_t)) : NULL;
  int32_t* cand_is_cat = has_cat ? (int32_t*)calloc((size_t)max_nodes, sizeof(int32_t)) : NULL;
  if (cat && cat->feature_is_cat) {
    memset(cat->node_is_cat, 0, sizeof(int32_t) * (size_t)max_nodes);
    memset(cat->node_cat_bits, 0, sizeof(uint32_t) * (size_t)max_nodes * cat->n_cat_words);
  }
  int32_t eval_seq = 0;
  /* lazily drawn level sets of the seed-driven column sampler */
  enum { kOrcMaxLevels = 64 };
  int32_t* lvl_sets = NULL;
  int32_t seed_pos = 0;
  int32_t lvl_drawn[kOrcMaxLevels], lvl_n[kOrcMaxLevels];
  int32_t* node_draw = NULL;
  memset(lvl_drawn, 0, sizeof(lvl_drawn));
  if (cat && cat->seeds) {
    lvl_sets = (int32_t*)malloc(sizeof(int32_t) * (size_t)kOrcMaxLevels * (size_t)(cat->n_tree_set ? cat->n_tree_set : 1));
    node_draw = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cat->n_tree_set ? cat->n_tree_set : 1));
  }
/* the column sampler's answer for node nidx_ (explicit node sets, or the seed-driven lazy sampler): overrides fs_v / nfs_v */
#define ORC_DRAW(nidx_, fs_v, nfs_v)                                                                                  \
  do {                                                                                                                \
    if (cat && cat->node_sets) {                                                                                      \
      int32_t k_ = eval_seq < cat->n_node_sets ? eval_seq : cat->n_node_sets - 1;                                     \
      fs_v = cat->node_sets + (size_t)k_ * cat->n_per_node;                                                           \
      nfs_v = cat->n_per_node;                                                                                        \
      eval_seq++;                                                                                                     \
    }                                                                                                                 \
    if (cat && cat->seeds) {                                                                                          \
      int32_t d_ = nodes[nidx_].depth < kOrcMaxLevels ? nodes[nidx_].depth : kOrcMaxLevels - 1;                       \
      int32_t* lv_ = lvl_sets + (size_t)d_ * cat->n_tree_set;                                                         \
      if (!lvl_drawn[d_]) {                                                                                           \
        uint32_t sd_ = 0;                                                                                             \
        if (cat->colsample_bylevel != 1.0f) sd_ = cat->seeds[seed_pos < cat->n_seeds ? seed_pos : cat->n_seeds - 1], seed_pos++; \
        lvl_n[d_] = orc_col_sample(cat->tree_set, cat->n_tree_set, cat->colsample_bylevel, sd_, lv_);                  \
        lvl_drawn[d_] = 1;                                                                                            \
      }                                                                                                               \
      {                                                                                                               \
        uint32_t sd_ = 0;                                                                                             \
        if (cat->colsample_bynode != 1.0f) sd_ = cat->seeds[seed_pos < cat->n_seeds ? seed_pos : cat->n_seeds - 1], seed_pos++; \
        nfs_v = orc_col_sample(lv_, lvl_n[d_], cat->colsample_bynode, sd_, node_draw);                                 \
        fs_v = node_draw;                                                                                             \
      }                                                                                                               \
    }                                                                                                                 \
  } while (0)
/* evaluation of node nidx_ on the list fs_ / nfs_ */
#define ORC_EVAL_ON(nidx_, fs_, nfs_, out_)                                                                           \
  do {                                                                                                                \
    if (cat && cat->feature_is_cat) {                                                                                 \
      int32_t th_;                                                                                                    \
      eval_node_cat_m(hist[nidx_], node_sum[nidx_], cut_ptrs, cut_values, n_features, fs_, nfs_, p, scale,            \
                      cat->feature_is_cat, cat->max_cat_to_onehot, cat->max_cat_threshold, out_, &cand_is_cat[nidx_], \
                      &th_, cand_bits + (size_t)(nidx_) * cat->n_cat_words, cat->n_cat_words, monotone, lower[nidx_], \
                      upper[nidx_]);                                                                                  \
    } else {                                                                                                          \
      orc_eval_node_mono(hist[nidx_], node_sum[nidx_], cut_ptrs, cut_values, n_features, fs_, nfs_, p, scale, out_,   \
                         NULL, monotone, lower[nidx_], upper[nidx_]);                                                 \
    }                                                                                                                 \
  } while (0)
#define ORC_EVAL(nidx_, fs_in_, nfs_in_, out_)                                                                        \
  do {                                                                                                                \
    const int32_t* fsd_ = fs_in_;                                                                                     \
    int32_t nfsd_ = nfs_in_;                                                                                          \
    ORC_DRAW(nidx_, fsd_, nfsd_);                                                                                     \
    ORC_EVAL_ON(nidx_, fsd_, nfsd_, out_);                                                                            \
  } while (0)
  uint64_t* ic_bits = ic ? (uint64_t*)calloc((size_t)max_nodes * (size_t)ic_words(ic->n_features), sizeof(uint64_t)) : NULL;
  int32_t* ic_buffer = ic ? (int32_t*)calloc((size_t)n_features + 1, sizeof(int32_t)) : NULL;
  /* n_levels > 0 (colsample_bylevel): feature_set is [n_levels][n_feature_set], a node at depth d is evaluated on
   * row min(d, n_levels - 1) = ColumnSampler::GetFeatureSet(depth), src/common/random.h:156-175 */
#define ORC_LEVEL_SET(d) \
  ((feature_set && n_levels > 0) ? feature_set + (size_t)((d) < n_levels ? (d) : n_levels - 1) * n_feature_set : feature_set)
  /* monotone (NULL or n_features signs): TreeEvaluator bounds per node, -/+FLT_MAX at the root */
  float* lower = (float*)malloc(sizeof(float) * (size_t)max_nodes);
  float* upper = (float*)malloc(sizeof(float) * (size_t)max_nodes);
  for (int32_t i = 0; i < max_nodes; ++i) { lower[i] = -FLT_MAX; upper[i] = FLT_MAX; }
  uint32_t total_bins = cut_ptrs[n_features];
  double scale[4];
  if (scale_override) memcpy(scale, scale_override, sizeof(scale));
  else orc_quantiser(gpair, n_rows, scale, NULL);
  if (scale_out) memcpy(scale_out, scale, sizeof(scale));
  orc_gpair64* gq = (orc_gpair64*)malloc(sizeof(orc_gpair64) * (n_rows ? n_rows : 1));
  orc_quantise(gpair, n_rows, scale, gq);
  if (subsample < 1.0f && cat && cat->gradient_based) {
    (void)orc_gradient_based_sample(gq, n_rows, scale, subsample, sample_seed, NULL, NULL);
  } else if (subsample < 1.0f) {
    for (uint64_t i = 0; i < n_rows; ++i)
      if (orc_random_weight(sample_seed, i) > subsample) { gq[i].g = 0; gq[i].h = 0; }
  }

  /* per-node state */
  orc_gpair64** hist = (orc_gpair64**)calloc((size_t)max_nodes, sizeof(orc_gpair64*));
  uint64_t* seg_beg = (uint64_t*)calloc((size_t)max_nodes, sizeof(uint64_t));
  uint64_t* seg_cnt = (uint64_t*)calloc((size_t)max_nodes, sizeof(uint64_t));
  orc_gpair64* node_sum = (orc_gpair64*)calloc((size_t)max_nodes, sizeof(orc_gpair64));
  uint32_t* ridx = (uint32_t*)malloc(sizeof(uint32_t) * (n_rows ? n_rows : 1));
  uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (n_rows ? n_rows : 1));
  entry_t* queue = (entry_t*)calloc((size_t)max_nodes, sizeof(entry_t)); /* open entries by nidx */
  for (uint64_t i = 0; i < n_rows; ++i) ridx[i] = (uint32_t)i;

  int32_t n_nodes = 1, num_leaves = 1;
  memset(&nodes[0], 0, sizeof(orc_node));
  nodes[0].parent = nodes[0].left = nodes[0].right = -1;
  nodes[0].is_leaf = 1;
  nodes[0].findex = -1;

  /* InitRoot */
  orc_gpair64 root = {0, 0};
  for (uint64_t i = 0; i < n_rows; ++i) root = gp_add(root, gq[i]);
  node_sum[0] = root;
  seg_beg[0] = 0;
  seg_cnt[0] = n_rows;
  hist[0] = (orc_gpair64*)calloc(total_bins, sizeof(orc_gpair64));
  orc_build_hist(gbins, stride, gq, ridx, n_rows, hist[0]);
  {
    double rg = (double)root.g * scale[2], rh = (double)root.h * scale[3];
    float w = (float)calc_weight_d(p, rg, rh);
    nodes[0].sum_hess = (float)rh;
    nodes[0].base_weight = w;
    nodes[0].leaf_value = p->learning_rate * w;
    nodes[0].fvalue = nodes[0].leaf_value;
  }
  entry_t root_e;
  root_e.nidx = 0; root_e.depth = 0; root_e.open = 1;
  ORC_EVAL(0, ORC_LEVEL_SET(0), n_feature_set, &root_e.c);
  if (cand_dump) cand_dump[0] = root_e.c;
  /* driver.Push(vector) filters loss_chg > kRtEps (driver.h:42-50) */
  int n_open = 0;
  if (root_e.c.loss_chg > 1e-6f) { queue[0] = root_e; n_open = 1; }

  while (n_open > 0) {
    /* ---- Driver::Pop (driver.h:72-100) ---- */
    int32_t expand[4096];
    int n_expand = 0;
    if (p->grow_policy == 1) { /* loss guide: single best entry; ties -> smaller nidx */
      int32_t bi = -1;
      for (int32_t i = 0; i < n_nodes; ++i)
        if (queue[i].open && (bi < 0 || queue[i].c.loss_chg > queue[bi].c.loss_chg)) bi = i;
      queue[bi].open = 0; n_open--;
      if (entry_valid(&queue[bi], p, num_leaves)) { num_leaves++; expand[n_expand++] = bi; }
      else break; /* Pop returned {} -> UpdateTree loop ends */
    } else { /* depth wise: all entries of the smallest-nidx entry's depth, ascending nidx */
      int32_t level = -1;
      for (int32_t i = 0; i < n_nodes && n_expand < g_node_batch; ++i) {
        if (!queue[i].open) continue;
        if (level < 0) level = queue[i].depth;
        if (queue[i].depth != level) break;
        queue[i].open = 0; n_open--;
        if (entry_valid(&queue[i], p, num_leaves)) { num_leaves++; expand[n_expand++] = i; }
      }
      if (n_expand == 0) break;
    }
    /* ---- ApplySplit (updater_gpu_hist.cu:610-655) ---- */
    for (int k = 0; k < n_expand; ++k) {
      int32_t nidx = expand[k];
      orc_candidate* c = &queue[nidx].c;
      if (n_nodes + 2 > max_nodes) { n_open = 0; break; }
      int32_t l = n_nodes, r = n_nodes + 1;
      n_nodes += 2;
      orc_gpair64 tot = gp_add(c->left_sum, c->right_sum);
      nodes[nidx].left = l; nodes[nidx].right = r;
      nodes[nidx].findex = c->findex;
      nodes[nidx].fvalue = c->fvalue;
      nodes[nidx].default_left = (c->dir == 0);
      nodes[nidx].is_leaf = 0;
      nodes[nidx].loss_chg = c->loss_chg;
      nodes[nidx].sum_hess = (float)((double)tot.h * scale[3]);
      nodes[nidx].base_weight = c->base_weight;
      float lw = c->left_weight * p->learning_rate, rw = c->right_weight * p->learning_rate;
      orc_node ln, rn;
      memset(&ln, 0, sizeof(ln));
      ln.parent = nidx; ln.left = ln.right = -1; ln.findex = -1; ln.is_leaf = 1;
      ln.depth = queue[nidx].depth + 1;
      rn = ln;
      ln.leaf_value = ln.fvalue = lw; ln.base_weight = lw;
      ln.sum_hess = (float)((double)c->left_sum.h * scale[3]);
      rn.leaf_value = rn.fvalue = rw; rn.base_weight = rw;
      rn.sum_hess = (float)((double)c->right_sum.h * scale[3]);
      nodes[l] = ln; nodes[r] = rn;
      node_sum[l] = c->left_sum; node_sum[r] = c->right_sum;
      if (ic) orc_ic_split(ic, ic_bits, nidx, c->findex, l, r);
      if (has_cat && cand_is_cat[nidx]) {
        cat->node_is_cat[nidx] = 1;
        memcpy(cat->node_cat_bits + (size_t)nidx * cat->n_cat_words, cand_bits + (size_t)nidx * cat->n_cat_words,
               sizeof(uint32_t) * cat->n_cat_words);
      }
      if (monotone) { /* TreeEvaluator::AddSplit (split_evaluator.h:163-203), weights before the learning rate */
        lower[l] = lower[r] = lower[nidx];
        upper[l] = upper[r] = upper[nidx];
        int32_t cst = monotone[c->findex];
        float mid = (c->left_weight + c->right_weight) / 2;
        if (cst < 0) { lower[l] = mid; upper[r] = mid; }
        else if (cst > 0) { upper[l] = mid; lower[r] = mid; }
      }
    }
    if (n_open == 0 && n_nodes + 2 > max_nodes) break;
    /* ---- PartitionAndBuildHist for the valid candidates ---- */
    int32_t eval_list[2 * 4096];
    int n_eval = 0;
    for (int k = 0; k < n_expand; ++k) {
      int32_t nidx = expand[k];
      int32_t depth = queue[nidx].depth;
      /* Driver::IsChildValid (driver.h:62-66) */
      if (p->max_depth > 0 && depth + 1 >= p->max_depth) continue;
      if (p->max_leaves > 0 && num_leaves >= p->max_leaves) continue;
      int32_t l = nodes[nidx].left, r = nodes[nidx].right;
      uint64_t nl;
      if (has_cat && cat->node_is_cat[nidx]) {
        const uint32_t* bits = cat->node_cat_bits + (size_t)nidx * cat->n_cat_words;
        const uint32_t* in = ridx + seg_beg[nidx];
        uint32_t* outp = tmp + seg_beg[nidx];
        uint64_t cnt = seg_cnt[nidx], li = 0, ri;
        nl = 0;
        for (uint64_t i = 0; i < cnt; ++i)
          nl += go_left_cat(gbins, stride, is_dense_layout, cut_ptrs, cut_values, in[i], nodes[nidx].findex, bits,
                            cat->n_cat_words, nodes[nidx].default_left) ? 1 : 0;
        ri = nl;
        for (uint64_t i = 0; i < cnt; ++i) {
          if (go_left_cat(gbins, stride, is_dense_layout, cut_ptrs, cut_values, in[i], nodes[nidx].findex, bits,
                          cat->n_cat_words, nodes[nidx].default_left)) outp[li++] = in[i];
          else outp[ri++] = in[i];
        }
      } else {
        nl = orc_partition(gbins, stride, is_dense_layout, cut_ptrs, cut_values, ridx + seg_beg[nidx],
                           seg_cnt[nidx], nodes[nidx].findex, nodes[nidx].fvalue,
                           nodes[nidx].default_left, tmp + seg_beg[nidx]);
      }
      memcpy(ridx + seg_beg[nidx], tmp + seg_beg[nidx], sizeof(uint32_t) * seg_cnt[nidx]);
      seg_beg[l] = seg_beg[nidx]; seg_cnt[l] = nl;
      seg_beg[r] = seg_beg[nidx] + nl; seg_cnt[r] = seg_cnt[nidx] - nl;
      /* build the child with fewer rows, subtract the other (any choice gives identical integers;
       * the reference compares hessian sums, updater_gpu_hist.cu:461-467) */
      int32_t build = seg_cnt[r] < seg_cnt[l] ? r : l, sub = build == l ? r : l;
      hist[build] = (orc_gpair64*)calloc(total_bins, sizeof(orc_gpair64));
      hist[sub] = (orc_gpair64*)calloc(total_bins, sizeof(orc_gpair64));
      orc_build_hist(gbins, stride, gq, ridx + seg_beg[build], seg_cnt[build], hist[build]);
      orc_subtract(hist[nidx], hist[build], total_bins, hist[sub]);
      eval_list[n_eval++] = l;
      eval_list[n_eval++] = r;
    }
    /* ---- EvaluateSplits over the batch: all Query calls first, then one evaluation launch ---- */
    int32_t* ic_lists = NULL; /* [n_eval][n_features] answers as returned */
    int32_t* ic_count = NULL;
    if (ic && n_eval > 0) {
      ic_lists = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_eval * (n_features ? n_features : 1));
      ic_count = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_eval);
      for (int k = 0; k < n_eval; ++k) {
        int32_t d = nodes[eval_list[k]].depth;
        int32_t* out = ic_lists + (size_t)k * n_features;
        const int32_t* sampled = feature_set ? ORC_LEVEL_SET(d) : NULL;
        int32_t n_sampled = sampled ? n_feature_set : 0;
        ORC_DRAW(eval_list[k], sampled, n_sampled); /* colsample_bynode: Query(the node's own draw, nidx) */
        ic_count[k] = orc_ic_query(ic, ic_bits, sampled, n_sampled, eval_list[k], out);
        memcpy(ic_buffer, out, sizeof(int32_t) * (size_t)ic_count[k]);
      }
    }
    for (int k = 0; k < n_eval; ++k) {
      int32_t c = eval_list[k];
      entry_t e;
      e.nidx = c; e.depth = nodes[c].depth; e.open = 1;
      const int32_t* fs = ORC_LEVEL_SET(e.depth);
      int32_t nfs = n_feature_set;
      if (ic) {
        fs = ic->shared_result_buffer ? ic_buffer : ic_lists + (size_t)k * n_features;
        nfs = ic_count[k];
      }
      /* an empty answer is a non-NULL list of length 0: nothing is scanned, the candidate stays invalid */
      if (ic) ORC_EVAL_ON(c, fs, nfs, &e.c); /* the draws were made with the queries */
      else ORC_EVAL(c, fs, nfs, &e.c);
      if (cand_dump) cand_dump[c] = e.c;
      if (e.c.loss_chg > 1e-6f) { queue[c] = e; n_open++; }
    }
    free(ic_lists); free(ic_count);
  }

  /* ---- FinalisePosition: walk every row from the root to its leaf ---- */
  if (position) {
    for (uint64_t r = 0; r < n_rows; ++r) {
      int32_t nidx = 0;
      while (!nodes[nidx].is_leaf) {
        int gl = (has_cat && cat->node_is_cat[nidx])
                     ? go_left_cat(gbins, stride, is_dense_layout, cut_ptrs, cut_values, r, nodes[nidx].findex,
                                   cat->node_cat_bits + (size_t)nidx * cat->n_cat_words, cat->n_cat_words,
                                   nodes[nidx].default_left)
                     : go_left(gbins, stride, is_dense_layout, cut_ptrs, cut_values, r, nodes[nidx].findex,
                               nodes[nidx].fvalue, nodes[nidx].default_left);
        nidx = gl ? nodes[nidx].left : nodes[nidx].right;
      }
      position[r] = gq[r].h != 0 ? nidx : ~nidx;
    }
  }
  if (hist_dump) {
    for (int32_t i = 0; i < n_nodes; ++i)
      if (hist[i]) memcpy(hist_dump + (uint64_t)i * total_bins, hist[i], sizeof(orc_gpair64) * total_bins);
  }
  for (int32_t i = 0; i < max_nodes; ++i) free(hist[i]);
  free(hist); free(seg_beg); free(seg_cnt); free(node_sum); free(ridx); free(tmp); free(queue); free(gq);
  free(lower); free(upper); free(ic_bits); free(ic_buffer); free(cand_bits); free(cand_is_cat);
  if (cat && cat->eval_count) *cat->eval_count = eval_seq;
  if (cat && cat->draws_used) *cat->draws_used = seed_pos;
  free(lvl_sets); free(node_draw);
#undef ORC_LEVEL_SET
#undef ORC_EVAL
#undef ORC_EVAL_ON
#undef ORC_DRAW
  return n_nodes;
}

/* R4  Prediction on the quantised matrix: GPUPredictor's kernel over an EllpackPage
 * (src/predictor/gpu_predictor.cu:180-230 PredictKernel with EllpackLoader, :316-351 PredictBatch): every row walks
 * every tree from the root -- missing -> default child, categorical split -> common::Decision on the bin's category,
 * else fvalue < split_cond, where the loader (src/predictor/gpu_data_accessor.cuh:95-107) hands out the bin's LOWER
 * cut as its value (-inf for a feature's first bin); for a condition that is one of the cuts this agrees with the
 * updater's cut_values[bin] <= split_cond, for any other condition only this rule is the reference's -- and the
 * leaf values of trees [0, n) are summed IN TREE ORDER into a float that starts at 0 and is then added to the output
 * (:221-228).  This function does one tree: out_leaf[r] = leaf the row reaches (cat arguments may be NULL). */
ORC_API void orc_predict_leaf(const int32_t* gbins, uint64_t n_rows, uint32_t stride, int is_dense_layout,
                              const uint32_t* cut_ptrs, const float* cut_values, const orc_node* nodes,
                              const int32_t* node_is_cat, const uint32_t* node_cat_bits, uint32_t n_cat_words,
                              int32_t* out_leaf) {
  for (uint64_t r = 0; r < n_rows; ++r) {
    int32_t nidx = 0;
    while (!nodes[nidx].is_leaf) {
      int gl = (node_is_cat && node_is_cat[nidx])
                   ? go_left_cat(gbins, stride, is_dense_layout, cut_ptrs, cut_values, r, nodes[nidx].findex,
                                 node_cat_bits + (size_t)nidx * n_cat_words, n_cat_words, nodes[nidx].default_left)
                   : go_left_predict(gbins, stride, is_dense_layout, cut_ptrs, cut_values, r, nodes[nidx].findex,
                                     nodes[nidx].fvalue, nodes[nidx].default_left);
      nidx = gl ? nodes[nidx].left : nodes[nidx].right;
    }
    out_leaf[r] = nidx;
  }
}

/* A3  UpdatePredictionCache (updater_gpu_hist.cu:582-608): pred[i] += leaf_value[decode(pos[i])] */
ORC_API void orc_update_prediction_cache(const orc_node* nodes, const int32_t* position, uint64_t n, float* preds) {
  for (uint64_t i = 0; i < n; ++i) {
    int32_t nidx = position[i] < 0 ? ~position[i] : position[i];
    preds[i] += nodes[nidx].leaf_value;
  }
}
