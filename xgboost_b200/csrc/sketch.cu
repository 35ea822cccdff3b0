// Quantile sketch -> HistogramCuts for dense batches resident on the device (unweighted, numerical features):
// the cut finding that precedes Ellpack binning, one sketch window per pushed batch, merged into a resident
// per-column summary exactly like the reference's QuantileDMatrix construction.
//
// Replaces (reference paths relative to /root/reference):
//   DeviceSketchWithHessian / ProcessWeightedBatch (thrust::sort of Entry{fidx, fvalue})   src/common/hist_util.cu:234-397
//   RequiredSampleCutsPerColumn, SketchEpsilon, WQuantileSketch::LimitSizeLevel            hist_util.cu:31-35, quantile.h:579-601,734-738
//   SketchContainer::Push -> PruneImpl + BinarySearchQueryIndex                            src/common/quantile.cu:36-61,108-153,340-378
//   SketchContainer::ScanInput (duplicate folding)                                         quantile.cu:380-418
//   SketchContainer::{Merge,FixError,Prune}, MergeImpl, MergePartition, SelectPruneIndices  quantile.cu:63-96,180-335,420-566
//   SketchContainer::IntermediateNumCuts                                                   quantile.cuh:68-72
//   AdapterDeviceSketch (window per batch), cuda_impl::MakeSketches (QuantileDMatrix)      hist_util.cuh:399-443, quantile_dmatrix.cu:29-126
//   SketchContainer::MakeCuts -> WQSummary::QueryCutValues (host, as in the reference)      quantile.cu:658-709, quantile.h:280-340
//
// A window's summary is fully determined by per-column ORDER STATISTICS: PruneImpl keeps
// `to = min(column_size, limit)` samples whose ranks follow from float32 arithmetic on (idx, to, n) alone.
// So the device work is (1) a transpose that turns the row-major batch into per-feature key columns
// (order-preserving uint32 image of the float, NaN dropped to the end), (2) a hand-written 4-pass LSD radix
// sort of every column at once (grid.y = feature; per-chunk digit histograms in shared memory, one scan CTA per
// column, stable scatter by warp match + per-warp digit counts), (3) a gather of the selected ranks into
// summary entries {rmin, rmax, wmin, value}.  (4) One CTA per column then folds duplicates (ScanInput), merges
// the window into the resident summary (each output entry finds its merge partition by binary search, as in
// the reference), drops equal values, repairs rank monotonicity (FixError) and (5) prunes the column to the
// budget of the rows seen so far.  All rank arithmetic is float32 in the reference's operation order.  The
// max_bin rank queries run on the host exactly like the reference's MakeCuts.  Cut values are bit-identical
// to the reference's on the golden inputs (tests/golden/ref_gpu_trees.json, ref_quantile_dmatrix_cuts.json).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace b200 {

constexpr int kSortChunk = 4096;   // keys per CTA per pass
constexpr int kSortThreads = 256;

__device__ __forceinline__ uint32_t FloatToKey(float v) {
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t KeyToFloatBits(uint32_t k) {
  return (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
}

// Row-major float batch -> column-major keys[f][n]; NaN becomes 0xFFFFFFFF (sorts last) and is not counted.
__global__ void __launch_bounds__(256) KeysTransposeKernel(const float* __restrict__ x, uint64_t n, uint64_t ld,
                                                            uint32_t n_features, uint32_t* __restrict__ keys,
                                                            uint64_t col_stride, unsigned long long* __restrict__ valid) {
  __shared__ uint32_t tile[32][33];
  __shared__ uint32_t s_valid[32];
  const uint32_t f0 = blockIdx.y * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (threadIdx.x < 32) s_valid[threadIdx.x] = 0u;
  // a CTA walks kTilesPerCta row tiles so that the per-feature counts cost one global atomic per CTA, not per warp
  constexpr uint32_t kTilesPerCta = 16;
  for (uint32_t it = 0; it < kTilesPerCta; ++it) {
    const uint64_t r0 = ((uint64_t)blockIdx.x * kTilesPerCta + it) * 32;
    if (r0 >= n) break;
    __syncthreads();
    for (uint32_t k = ty; k < 32; k += 8) {
      const uint64_t r = r0 + k;
      const uint32_t f = f0 + tx;
      uint32_t key = 0xFFFFFFFFu;
      if (r < n && f < n_features) {
        const float v = x[r * ld + f];
        if (!isnan(v)) key = FloatToKey(v);
      }
      tile[k][tx] = key;
    }
    __syncthreads();
    for (uint32_t k = ty; k < 32; k += 8) {
      const uint32_t f = f0 + k;
      const uint64_t r = r0 + tx;
      const uint32_t key = tile[tx][k];
      const bool ok = f < n_features && r < n;
      if (ok) keys[(uint64_t)f * col_stride + r] = key;
      const uint32_t m = __ballot_sync(0xffffffffu, ok && key != 0xFFFFFFFFu);
      if (tx == 0 && m) atomicAdd(&s_valid[k], (uint32_t)__popc(m));
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && f0 + threadIdx.x < n_features && s_valid[threadIdx.x])
    atomicAdd(&valid[f0 + threadIdx.x], (unsigned long long)s_valid[threadIdx.x]);
}

// pass 1/3: digit histogram of every (column, chunk)
__global__ void __launch_bounds__(kSortThreads) RadixCountKernel(const uint32_t* __restrict__ keys, uint64_t n,
                                                                  uint64_t col_stride, uint32_t shift,
                                                                  uint32_t* __restrict__ counts, uint32_t n_chunks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t f = blockIdx.y, c = blockIdx.x;
  const uint64_t b = (uint64_t)c * kSortChunk, e = min(n, b + kSortChunk);
  const uint32_t* __restrict__ col = keys + (uint64_t)f * col_stride;
  for (uint64_t i = b + threadIdx.x; i < e; i += kSortThreads) atomicAdd(&h[(col[i] >> shift) & 255u], 1u);
  __syncthreads();
  counts[((uint64_t)f * n_chunks + c) * 256 + threadIdx.x] = h[threadIdx.x];
}

// pass 2/3: one CTA per column: base[c][d] = sum_{d'<d} total[d'] + sum_{c'<c} counts[c'][d]
__global__ void __launch_bounds__(256) RadixScanKernel(uint32_t* __restrict__ counts, uint32_t n_chunks) {
  __shared__ uint32_t total[256];
  const uint32_t f = blockIdx.x, d = threadIdx.x;
  uint32_t* __restrict__ cf = counts + (uint64_t)f * n_chunks * 256;
  uint32_t run = 0;
  for (uint32_t c = 0; c < n_chunks; ++c) {  // exclusive over chunks, per digit
    const uint32_t v = cf[(uint64_t)c * 256 + d];
    cf[(uint64_t)c * 256 + d] = run;
    run += v;
  }
  total[d] = run;
  __syncthreads();
  if (d == 0) {  // exclusive over digits
    uint32_t acc = 0;
    for (int k = 0; k < 256; ++k) {
      const uint32_t v = total[k];
      total[k] = acc;
      acc += v;
    }
  }
  __syncthreads();
  const uint32_t base = total[d];
  for (uint32_t c = 0; c < n_chunks; ++c) cf[(uint64_t)c * 256 + d] += base;
}

// pass 3/3: stable scatter of one chunk.  Rounds of 256 keys; inside a round the rank of a key among equal
// digits = (keys of earlier warps) + (earlier lanes of its warp with the same digit, via match_any).
__global__ void __launch_bounds__(kSortThreads) RadixScatterKernel(const uint32_t* __restrict__ in,
                                                                    uint32_t* __restrict__ out, uint64_t n,
                                                                    uint64_t col_stride, uint32_t shift,
                                                                    const uint32_t* __restrict__ bases,
                                                                    uint32_t n_chunks) {
  __shared__ uint32_t running[256];         // next free slot of each digit for this chunk
  __shared__ uint32_t warp_cnt[8][256];     // per round: keys of each warp per digit
  const uint32_t f = blockIdx.y, c = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  running[threadIdx.x] = bases[((uint64_t)f * n_chunks + c) * 256 + threadIdx.x];
  const uint64_t b = (uint64_t)c * kSortChunk, e = min(n, b + kSortChunk);
  const uint32_t* __restrict__ src = in + (uint64_t)f * col_stride;
  uint32_t* __restrict__ dst = out + (uint64_t)f * col_stride;
  for (uint64_t r0 = b; r0 < e; r0 += kSortThreads) {
    for (uint32_t k = threadIdx.x; k < 8 * 256; k += kSortThreads) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const uint64_t i = r0 + threadIdx.x;
    const bool ok = i < e;
    const uint32_t key = ok ? src[i] : 0u;
    const uint32_t d = (key >> shift) & 255u;
    const uint32_t active = __ballot_sync(0xffffffffu, ok);
    uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : 0xFFFFFFFFu) & active;
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank_in_warp == 0) warp_cnt[warp][d] = __popc(peers);
    __syncthreads();
    uint32_t before = 0;  // same digit in earlier warps of this round
    if (ok)
      for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[w][d];
    if (ok) dst[running[d] + before + rank_in_warp] = key;
    __syncthreads();
    {  // advance the running offsets by this round's totals
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += warp_cnt[w][threadIdx.x];
      running[threadIdx.x] += t;
    }
    __syncthreads();
  }
}

// PruneImpl (quantile.cu:108-153) for unweighted entries e_i = {rmin = i, rmax = i + 1, wmin = 1}:
// output slot idx of a column keeps sample BinarySearchQueryIndex(q), all in float32 like the reference.
__device__ __forceinline__ uint32_t QueryIndex(uint32_t n, float rank) {
  const float rank2 = rank * 2.0f;
  if (rank2 < 0.0f + 1.0f) return 0u;
  if (rank2 >= (float)(n - 1) + (float)n) return n - 1;
  uint32_t lo = 1, hi = n - 1;  // upper_bound of rank2 over (2i+1), i in [1, n-1)
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (!(rank2 < (float)mid + (float)(mid + 1))) lo = mid + 1; else hi = mid;
  }
  const uint32_t i = lo - 1;
  return (rank2 < ((float)i + 1.0f) + ((float)(i + 2) - 1.0f)) ? i : i + 1;
}

// Summary entry = float4 {x: rmin, y: rmax, z: wmin, w: value}  (SketchEntry, quantile.h)
__global__ void __launch_bounds__(256) PruneSelectKernel(const uint32_t* __restrict__ sorted_keys, uint64_t col_stride,
                                                          const unsigned long long* __restrict__ valid,
                                                          uint32_t per_col, float4* __restrict__ out) {
  const uint32_t f = blockIdx.y;
  const uint32_t n = (uint32_t)valid[f];
  const uint32_t to = min(n, per_col);
  const uint32_t* __restrict__ col = sorted_keys + (uint64_t)f * col_stride;
  for (uint32_t idx = blockIdx.x * 256 + threadIdx.x; idx < to; idx += gridDim.x * 256) {
    uint32_t sel;
    if (n <= to) {
      sel = idx;
    } else if (idx == 0) {
      sel = 0;
    } else if (idx == to - 1) {
      sel = n - 1;
    } else {
      const float front_rmax = 1.0f;
      const float w = (float)(n - 1) - front_rmax;
      const float q = ((float)idx * w) / ((float)to - 1.0f) + front_rmax;
      sel = QueryIndex(n, q);
    }
    out[(uint64_t)f * per_col + idx] =
        make_float4((float)sel, (float)(sel + 1), 1.0f, __uint_as_float(KeyToFloatBits(col[sel])));
  }
}

constexpr int kColThreads = 1024;

// Exclusive prefix sum of one value per thread over the CTA; *total = sum.  `ws` = 33 words of shared memory.
__device__ __forceinline__ uint32_t BlockExclusiveScan(uint32_t v, uint32_t* ws, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (uint32_t)d) inc += t;
  }
  __syncthreads();  // ws may still be read from a previous call
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t nw = blockDim.x >> 5;
    uint32_t w = lane < nw ? ws[lane] : 0u, winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= (uint32_t)d) winc += t;
    }
    if (lane < nw) ws[lane] = winc - w;
    if (lane == 31) ws[32] = winc;
  }
  __syncthreads();
  *total = ws[32];
  return ws[warp] + inc - v;
}

// Stable compaction inside one CTA: every thread owns a contiguous slice of [0, n); emit(i, slot) is called for
// kept items in order.  Returns the number kept.
template <typename Keep, typename Emit>
__device__ __forceinline__ uint32_t BlockCompact(uint32_t n, Keep keep, Emit emit, uint32_t* ws) {
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t b = min(n, threadIdx.x * per), e = min(n, b + per);
  uint32_t c = 0;
  for (uint32_t i = b; i < e; ++i) c += keep(i) ? 1u : 0u;
  uint32_t total;
  uint32_t slot = BlockExclusiveScan(c, ws, &total);
  for (uint32_t i = b; i < e; ++i)
    if (keep(i)) emit(i, slot++);
  __syncthreads();
  return total;
}

__device__ __forceinline__ float RMinNext(const float4& e) { return e.x + e.z; }
__device__ __forceinline__ float RMaxPrev(const float4& e) { return e.y - e.z; }

// MergeImpl::merge_entry_at (quantile.cu:232-322) for output position k of one column; x = resident (m > 0
// entries), y = incoming (n > 0 entries).
__device__ __forceinline__ float4 MergeEntryAt(const float4* x, uint32_t m, const float4* y, uint32_t n, uint32_t k) {
  // MergePartition (quantile.cu:180-198): first i in [low, high] for which NOT (j > 0 && i < m && y[j-1] >= x[i])
  uint32_t low = k > n ? k - n : 0u;
  const uint32_t high = min(k, m);
  uint32_t len = high - low + 1;
  while (len > 0) {
    const uint32_t half = len >> 1, i = low + half, j = k - i;
    const bool need_more_x = j > 0 && i < m && y[j - 1].w >= x[i].w;
    if (need_more_x) {
      low = i + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  const uint32_t a = low, b = k - low;
  auto other_rmin = [](const float4* c, uint32_t ind) { return ind == 0 ? 0.0f : RMinNext(c[ind - 1]); };
  auto other_rmax = [](const float4* c, uint32_t size, uint32_t ind) {
    return ind == size ? c[size - 1].y : RMaxPrev(c[ind]);
  };
  if (a == m) {
    const float4 ye = y[b];
    return make_float4(other_rmin(x, a) + ye.x, other_rmax(x, m, a) + ye.y, ye.z, ye.w);
  }
  const float4 xe = x[a];
  if (b == n) return make_float4(xe.x + other_rmin(y, b), xe.y + other_rmax(y, n, b), xe.z, xe.w);
  const float4 ye = y[b];
  if (xe.w == ye.w) return make_float4(xe.x + ye.x, xe.y + ye.y, xe.z + ye.z, xe.w);
  if (xe.w < ye.w) return make_float4(xe.x + other_rmin(y, b), xe.y + other_rmax(y, n, b), xe.z, xe.w);
  return make_float4(other_rmin(x, a) + ye.x, other_rmax(x, m, a) + ye.y, ye.z, ye.w);
}

// SketchContainer::Push after PruneImpl (quantile.cu:370-377): ScanInput, Merge (+ SegmentedUnique + FixError).
// One CTA per column.  `inc` holds the window's selected entries (stride inc_stride, min(valid, inc_stride) used),
// `res` the resident column (len[f] entries), `m1`/`m2` are scratch columns of the resident stride; the new
// resident column is left in `m1` (the host swaps res <-> m1).
__global__ void __launch_bounds__(kColThreads) ColumnPushKernel(const float4* __restrict__ inc, uint32_t inc_stride,
                                                                 const unsigned long long* __restrict__ valid,
                                                                 uint32_t n_features, float4* inc_unique,
                                                                 const float4* __restrict__ res, float4* m1, float4* m2,
                                                                 uint64_t stride,
                                                                 uint32_t* __restrict__ len,
                                                                 uint32_t* __restrict__ fix_changes) {
  __shared__ uint32_t ws[33];
  __shared__ uint32_t any_s;
  const uint32_t f = blockIdx.x;
  // `if (n_uniques == 0) return;` -- a window without a single valid entry leaves the sketch untouched
  if (threadIdx.x == 0) any_s = 0;
  __syncthreads();
  {
    uint32_t any = 0;
    for (uint32_t k = threadIdx.x; k < n_features; k += blockDim.x) any |= valid[k] != 0 ? 1u : 0u;
    if (any) atomicOr(&any_s, 1u);
  }
  __syncthreads();
  const uint32_t m = len[f];
  const float4* __restrict__ x = res + (uint64_t)f * stride;
  float4* a1 = m1 + (uint64_t)f * stride;  // written and re-read inside this kernel: no __restrict__/read-only path
  float4* a2 = m2 + (uint64_t)f * stride;
  if (any_s == 0) {  // keep the resident column where the host expects it after the swap
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) a1[i] = x[i];
    return;
  }
  const uint32_t to = (uint32_t)min((unsigned long long)inc_stride, valid[f]);
  const float4* __restrict__ src = inc + (uint64_t)f * inc_stride;
  float4* y = inc_unique + (uint64_t)f * inc_stride;
  // ScanInput (quantile.cu:380-418): the reverse scan adds the weight of the entry to the right when it holds the
  // same value at a different rmin; SegmentedUnique then keeps the first entry of every run of equal values.
  const uint32_t n = BlockCompact(
      to, [&](uint32_t i) { return i == 0 || (src[i].w - src[i - 1].w) != 0.0f; },
      [&](uint32_t i, uint32_t slot) {
        float4 e = src[i];
        uint32_t last = i;  // end of the chain of same-value entries whose rmin keeps changing
        while (last + 1 < to && src[last + 1].w - e.w == 0.0f && src[last].x - src[last + 1].x != 0.0f) ++last;
        if (last > i) {
          float w = src[last].z;
          for (uint32_t k = last; k > i; --k) w = src[k - 1].z + w;  // l.wmin + r.wmin, right to left
          e.z = w;
          e.y = e.x + w;
        }
        y[slot] = e;
      },
      ws);
  // Merge (quantile.cu:491-546)
  const uint32_t total = m + n;
  if (m == 0) {
    for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) a2[k] = y[k];
  } else if (n == 0) {
    for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) a2[k] = x[k];
  } else {
    for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) a2[k] = MergeEntryAt(x, m, y, n, k);
  }
  __syncthreads();
  // normalize_merged: SegmentedUnique(SketchUnique) ...
  const uint32_t nu = BlockCompact(
      total, [&](uint32_t i) { return i == 0 || (a2[i].w - a2[i - 1].w) != 0.0f; },
      [&](uint32_t i, uint32_t slot) { a1[slot] = a2[i]; }, ws);
  // ... + FixError (quantile.cu:548-566).  The reference patches entries in place from concurrent threads; here
  // every entry reads its left neighbour as it was before the pass (identical whenever no repair cascades).
  uint32_t changed = 0;
  for (uint32_t i = threadIdx.x; i < nu; i += blockDim.x) {
    float4 e = a1[i];
    const float4 before = e;
    const float prev_rmin = i == 0 ? 0.0f : a1[i - 1].x;
    const float prev_rmax = i == 0 ? 0.0f : a1[i - 1].y;
    if (e.x < prev_rmin) e.x = prev_rmin;
    if (e.y < prev_rmax) e.y = prev_rmax;
    const float rmin_next = e.x + e.z;
    if (e.y < rmin_next) e.y = rmin_next;
    a2[i] = e;
    changed += (e.x != before.x || e.y != before.y) ? 1u : 0u;
  }
  if (changed) atomicAdd(fix_changes, changed);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nu; i += blockDim.x) a1[i] = a2[i];
  if (threadIdx.x == 0) len[f] = nu;
}

// BinarySearchQueryIndex (quantile.cu:36-61) over general entries.
__device__ __forceinline__ uint32_t QueryIndexGeneral(const float4* c, uint32_t n, float rank) {
  rank *= 2.0f;
  if (rank < c[0].x + c[0].y) return 0u;
  if (rank >= c[n - 1].x + c[n - 1].y) return n - 1;
  uint32_t first = 1, cnt = n - 2;  // upper_bound over [1, n-1)
  while (cnt > 0) {
    const uint32_t half = cnt >> 1, mid = first + half;
    if (!(rank < c[mid].x + c[mid].y)) {
      first = mid + 1;
      cnt -= half + 1;
    } else {
      cnt = half;
    }
  }
  const uint32_t i = first - 1;
  return (rank < RMinNext(c[i]) + RMaxPrev(c[i + 1])) ? i : i + 1;
}

// SketchContainer::Prune (quantile.cu:420-489) for one column per CTA: SelectPruneIndices, SegmentedUnique of
// the selected indices, gather, SegmentedUnique of the values.  Columns no longer than `to` are left alone.
__global__ void __launch_bounds__(kColThreads) ColumnPruneKernel(float4* res, float4* m1, uint32_t* sel_scratch,
                                                                  uint64_t stride,
                                                                  uint32_t* __restrict__ len, uint32_t to) {
  __shared__ uint32_t ws[33];
  const uint32_t f = blockIdx.x;
  const uint32_t n = len[f];
  if (n <= to) return;
  float4* c = res + (uint64_t)f * stride;
  float4* t = m1 + (uint64_t)f * stride;
  uint32_t* sel = sel_scratch + (uint64_t)f * stride;
  const float4 front = c[0], back = c[n - 1];
  for (uint32_t idx = threadIdx.x; idx < to; idx += blockDim.x) {
    uint32_t s;
    if (idx == 0) {
      s = 0;
    } else if (idx == to - 1) {
      s = n - 1;
    } else {
      const float w = back.x - front.y;
      const float q = ((float)idx * w) / ((float)to - 1.0f) + front.y;
      s = QueryIndexGeneral(c, n, q);
    }
    sel[idx] = s;
  }
  __syncthreads();
  const uint32_t ns = BlockCompact(
      to, [&](uint32_t i) { return i == 0 || sel[i] != sel[i - 1]; },
      [&](uint32_t i, uint32_t slot) { t[slot] = c[sel[i]]; }, ws);
  const uint32_t nu = BlockCompact(
      ns, [&](uint32_t i) { return i == 0 || (t[i].w - t[i - 1].w) != 0.0f; },
      [&](uint32_t i, uint32_t slot) { c[slot] = t[i]; }, ws);
  if (threadIdx.x == 0) len[f] = nu;
}

// ---- host side ----------------------------------------------------------------------------------------
static size_t LimitSizeLevel(size_t maxn, double eps) {  // quantile.h:579-601, kFactor = 8
  if (maxn == 0) return 1;
  const double internal_eps = eps / 8.0;
  size_t nlevel = 1, limit = 1;
  while (true) {
    limit = static_cast<size_t>(std::ceil(nlevel / internal_eps)) + 1;
    limit = std::min(maxn, limit);
    if ((1ull << nlevel) * limit >= maxn) break;
    ++nlevel;
  }
  return limit;
}
static double SketchEpsilon(int32_t max_bin, size_t n) {  // quantile.h:734-738
  return 1.0 / static_cast<double>(std::min<size_t>((size_t)max_bin, std::max<size_t>(1, n)));
}
static size_t RequiredSampleCutsPerColumn(int32_t max_bin, size_t n) {  // hist_util.cu:31-35
  return std::min(LimitSizeLevel(n, SketchEpsilon(max_bin, n)), n);
}

struct SummaryEntry {
  float rmin, rmax, wmin, value;
};

// WQSummary::QueryCutValues (quantile.h:280-340)
static void QueryCutValues(const SummaryEntry* s, size_t n, size_t max_bin, std::vector<float>* out) {
  if (n == 0) {
    out->push_back(1e-5f);
    return;
  }
  auto advance = [&](size_t cursor, float v) {
    while (cursor < n && s[cursor].value <= v) ++cursor;
    return cursor;
  };
  float last_cut = s[0].value;
  size_t next = advance(1, last_cut);
  if (n <= max_bin) {
    while (next < n) {
      float cpt = s[next].value;
      out->push_back(cpt);
      last_cut = cpt;
      next = advance(next + 1, last_cut);
    }
  } else {
    const double total = static_cast<double>(s[n - 1].rmax);
    size_t qc = 0;
    for (size_t i = 1; i < max_bin; ++i) {
      const double rank2 = 2.0 * (static_cast<double>(i) * total / static_cast<double>(max_bin));
      while (qc < n - 2 && rank2 >= static_cast<double>(s[qc + 1].rmin + s[qc + 1].rmax)) ++qc;
      const float rmin_next = s[qc].rmin + s[qc].wmin;
      const float rmax_prev = s[qc + 1].rmax - s[qc + 1].wmin;
      const SummaryEntry& q = rank2 < static_cast<double>(rmin_next + rmax_prev) ? s[qc] : s[qc + 1];
      float cpt = q.value;
      if (cpt <= last_cut) {
        next = advance(next, last_cut);
        if (next == n) break;
        cpt = s[next].value;
      } else if (next < n && s[next].value <= cpt) {
        next = advance(next + 1, cpt);
      }
      out->push_back(cpt);
      last_cut = cpt;
    }
  }
  const float cpt = s[n - 1].value;
  out->push_back(cpt + (std::fabs(cpt) + 1e-5f));
}

}  // namespace b200

using namespace b200;

// Resident sketch (SketchContainer, quantile.cuh:38-200): columns at a fixed stride, lengths on the device.
struct b200hist_sketch_s {
  int device = 0;
  uint32_t n_features = 0;
  int32_t max_bin = 0;
  std::vector<uint8_t> is_cat;  // feature types (empty = all numerical): a categorical column's cuts are its codes
  uint64_t rows_seen = 0;       // SketchContainer::rows_seen_
  uint64_t accumulated = 0;     // ExternalDataInfo::accumulated_rows
  uint64_t stride = 0;          // entries per column in res/m1/m2/sel
  uint64_t len_bound = 0;       // host upper bound of any column's length
  float4* res = nullptr;
  float4* m1 = nullptr;
  float4* m2 = nullptr;
  uint32_t* sel = nullptr;
  uint32_t* len = nullptr;
  uint32_t* fix_changes = nullptr;
  // window scratch, grow-only: sort double buffer, digit counts, valid counts, selected / folded entries
  uint32_t* keys_a = nullptr;
  uint32_t* keys_b = nullptr;
  uint32_t* counts = nullptr;
  unsigned long long* valid = nullptr;
  float4* inc = nullptr;
  float4* inc_unique = nullptr;
  uint64_t cap_keys = 0, cap_counts = 0, cap_inc = 0;
  ~b200hist_sketch_s() {
    for (void* p : {(void*)res, (void*)m1, (void*)m2, (void*)sel, (void*)len, (void*)fix_changes, (void*)keys_a,
                    (void*)keys_b, (void*)counts, (void*)valid, (void*)inc, (void*)inc_unique})
      if (p) cudaFree(p);
  }
};

namespace {

void EnsureStride(b200hist_sketch_s* sk, uint64_t need, cudaStream_t s) {
  if (need <= sk->stride) return;
  const uint64_t ns = std::max<uint64_t>(need, sk->stride + sk->stride / 2);
  const uint64_t F = sk->n_features;
  float4* nres = DevAlloc<float4>(F * ns);
  if (sk->res && sk->len_bound > 0)
    B200_CUDA(cudaMemcpy2DAsync(nres, ns * sizeof(float4), sk->res, sk->stride * sizeof(float4),
                                sk->len_bound * sizeof(float4), F, cudaMemcpyDeviceToDevice, s));
  B200_CUDA(cudaStreamSynchronize(s));
  for (void* p : {(void*)sk->res, (void*)sk->m1, (void*)sk->m2, (void*)sk->sel})
    if (p) cudaFree(p);
  sk->res = nres;
  sk->m1 = DevAlloc<float4>(F * ns);
  sk->m2 = DevAlloc<float4>(F * ns);
  sk->sel = DevAlloc<uint32_t>(F * ns);
  sk->stride = ns;
}

void PruneTo(b200hist_sketch_s* sk, size_t to, cudaStream_t s) {
  if (sk->len_bound <= to) return;  // no column can be longer than `to`
  ColumnPruneKernel<<<sk->n_features, kColThreads, 0, s>>>(sk->res, sk->m1, sk->sel, sk->stride, sk->len,
                                                           (uint32_t)to);
  B200_LAUNCHED();
  sk->len_bound = to;
}

// One sketch window = one batch (AdapterDeviceSketch with the whole batch in one window, hist_util.cuh:399-443),
// followed by the per-batch prune of QuantileDMatrix (quantile_dmatrix.cu:73-90).
void PushDense(b200hist_sketch_s* sk, const float* x_dev, uint64_t n, uint64_t ld, cudaStream_t s) {
  const uint32_t F = sk->n_features;
  if (n == 0) return;      // an empty batch opens no window (hist_util.cuh:415) and changes no budget
  sk->accumulated += n;
  sk->rows_seen += n;      // approx_n_samples = ceil(n * F / F)
  const uint32_t per_col = (uint32_t)RequiredSampleCutsPerColumn(sk->max_bin, n);
  const uint64_t col_stride = (n + 3) & ~uint64_t(3);
  const uint32_t n_chunks = (uint32_t)((n + kSortChunk - 1) / kSortChunk);
  auto grow = [&](auto*& ptr, uint64_t& cap, uint64_t need) {
    if (need <= cap) return;
    B200_CUDA(cudaStreamSynchronize(s));
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    ptr = DevAlloc<std::remove_pointer_t<std::remove_reference_t<decltype(ptr)>>>(need);
    cap = need;
  };
  {
    uint64_t c = sk->cap_keys;
    grow(sk->keys_a, c, (uint64_t)F * col_stride);
    grow(sk->keys_b, sk->cap_keys, (uint64_t)F * col_stride);
    grow(sk->counts, sk->cap_counts, (uint64_t)F * n_chunks * 256);
    c = sk->cap_inc;
    grow(sk->inc, c, (uint64_t)F * per_col);
    grow(sk->inc_unique, sk->cap_inc, (uint64_t)F * per_col);
    if (!sk->valid) sk->valid = DevAlloc<unsigned long long>(F);
  }
  uint32_t* keys_a = sk->keys_a;
  uint32_t* keys_b = sk->keys_b;
  uint32_t* counts = sk->counts;
  unsigned long long* valid = sk->valid;
  float4* inc = sk->inc;
  float4* inc_unique = sk->inc_unique;
  B200_CUDA(cudaMemsetAsync(valid, 0, sizeof(unsigned long long) * F, s));
  dim3 tgrid((unsigned)((n + 32 * 16 - 1) / (32 * 16)), (F + 31) / 32);
  KeysTransposeKernel<<<tgrid, 256, 0, s>>>(x_dev, n, ld, F, keys_a, col_stride, valid);
  B200_LAUNCHED();
  uint32_t* src = keys_a;
  uint32_t* dst = keys_b;
  dim3 sgrid(n_chunks, F);
  for (uint32_t shift = 0; shift < 32; shift += 8) {
    RadixCountKernel<<<sgrid, kSortThreads, 0, s>>>(src, n, col_stride, shift, counts, n_chunks);
    B200_LAUNCHED();
    RadixScanKernel<<<F, 256, 0, s>>>(counts, n_chunks);
    B200_LAUNCHED();
    RadixScatterKernel<<<sgrid, kSortThreads, 0, s>>>(src, dst, n, col_stride, shift, counts, n_chunks);
    B200_LAUNCHED();
    std::swap(src, dst);
  }
  dim3 pgrid((per_col + 255) / 256, F);
  PruneSelectKernel<<<pgrid, 256, 0, s>>>(src, col_stride, valid, per_col, inc);
  B200_LAUNCHED();
  EnsureStride(sk, sk->len_bound + per_col, s);
  ColumnPushKernel<<<F, kColThreads, 0, s>>>(inc, per_col, valid, F, inc_unique, sk->res, sk->m1, sk->m2, sk->stride,
                                             sk->len, sk->fix_changes);
  B200_LAUNCHED();
  std::swap(sk->res, sk->m1);
  sk->len_bound += per_col;
  // Push ends with Prune(IntermediateNumCuts()) (quantile.cu:375-377, quantile.cuh:68-72) ...
  const size_t rows = std::max<size_t>(1, sk->rows_seen);
  PruneTo(sk, LimitSizeLevel(rows, SketchEpsilon(sk->max_bin, rows)) * (size_t)F, s);
  // ... and MakeSketches prunes to the budget of the rows accumulated so far (quantile_dmatrix.cu:73-90)
  PruneTo(sk, RequiredSampleCutsPerColumn(sk->max_bin, sk->accumulated), s);
  // the caller may free or overwrite the batch as soon as the call returns (like a DataIter does)
  B200_CUDA(cudaStreamSynchronize(s));
}

}  // namespace

extern "C" int b200hist_sketch_create(int device, uint32_t n_features, int32_t max_bin, b200hist_sketch_t* out) {
  B200_API_BEGIN();
  B200_CHECK(out != nullptr, "null argument");
  B200_CHECK(n_features > 0 && max_bin >= 2, "need n_features > 0 and max_bin >= 2");
  DeviceGuard guard(device);
  auto* sk = new b200hist_sketch_s();
  sk->device = device;
  sk->n_features = n_features;
  sk->max_bin = max_bin;
  sk->len = DevAlloc<uint32_t>(n_features);
  sk->fix_changes = DevAlloc<uint32_t>(1);
  B200_CUDA(cudaMemset(sk->len, 0, sizeof(uint32_t) * n_features));
  B200_CUDA(cudaMemset(sk->fix_changes, 0, sizeof(uint32_t)));
  *out = sk;
  B200_API_END();
}

extern "C" int b200hist_sketch_destroy(b200hist_sketch_t sk) {
  B200_API_BEGIN();
  if (sk) {
    DeviceGuard guard(sk->device);
    delete sk;
  }
  B200_API_END();
}

extern "C" int b200hist_sketch_push_dense(b200hist_sketch_t sk, const float* x_dev, uint64_t n, uint64_t ld,
                                          void* stream) {
  B200_API_BEGIN();
  B200_CHECK(sk != nullptr && (x_dev != nullptr || n == 0), "null argument");
  B200_CHECK(ld >= sk->n_features, "leading dimension smaller than n_features");
  B200_CHECK(n < (1ull << 24), "a sketch window holds fewer than 2^24 rows (float32 ranks); push smaller batches");
  B200_CHECK(n * (uint64_t)sk->n_features <= 0x7fffffffull,
             "a sketch window holds at most 2^31-1 elements (hist_util.cu:89,127); push smaller batches");
  DeviceGuard guard(sk->device);
  PushDense(sk, x_dev, n, ld, static_cast<cudaStream_t>(stream));
  B200_API_END();
}

extern "C" int b200hist_sketch_set_feature_types(b200hist_sketch_t sk, const uint8_t* feature_is_cat_host, uint32_t n) {
  B200_API_BEGIN();
  B200_CHECK(sk != nullptr, "null argument");
  B200_CHECK(n <= sk->n_features, "more feature types than features");
  sk->is_cat.clear();
  bool any = false;
  for (uint32_t f = 0; feature_is_cat_host != nullptr && f < n; ++f) any = any || feature_is_cat_host[f] != 0;
  if (any) {
    sk->is_cat.assign(sk->n_features, 0);
    for (uint32_t f = 0; f < n; ++f) sk->is_cat[f] = feature_is_cat_host[f] != 0 ? 1 : 0;
  }
  B200_API_END();
}

extern "C" int b200hist_sketch_make_cuts(b200hist_sketch_t sk, uint32_t* cut_ptrs_out_host, float* cut_values_out_host,
                                         uint64_t cut_values_capacity, uint32_t* fix_error_changes_out, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(sk != nullptr && cut_ptrs_out_host != nullptr && cut_values_out_host != nullptr, "null argument");
  DeviceGuard guard(sk->device);
  auto s = static_cast<cudaStream_t>(stream);
  const uint32_t F = sk->n_features;
  std::vector<uint32_t> h_len(F, 0);
  B200_CUDA(cudaMemcpyAsync(h_len.data(), sk->len, sizeof(uint32_t) * F, cudaMemcpyDeviceToHost, s));
  uint32_t h_fix = 0;
  B200_CUDA(cudaMemcpyAsync(&h_fix, sk->fix_changes, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  const uint64_t width = sk->len_bound;
  std::vector<SummaryEntry> h_entries((size_t)F * width);
  if (width > 0)
    B200_CUDA(cudaMemcpy2DAsync(h_entries.data(), width * sizeof(SummaryEntry), sk->res, sk->stride * sizeof(float4),
                                width * sizeof(float4), F, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  if (fix_error_changes_out) *fix_error_changes_out = h_fix;
  uint64_t total = 0;
  cut_ptrs_out_host[0] = 0;
  std::vector<float> cuts;
  for (uint32_t f = 0; f < F; ++f) {
    B200_CHECK(h_len[f] <= width, "internal: column longer than its bound");
    cuts.clear();
    if (!sk->is_cat.empty() && sk->is_cat[f]) {
      // categorical column (quantile.cu:684-696): cut values = every code 0 .. max category.  The column's largest
      // value is the last entry of its summary whatever was pruned (a summary always keeps its extremes); the
      // reference keeps categorical columns unpruned, which changes nothing here.
      const SummaryEntry* col = h_entries.data() + (size_t)f * width;
      for (uint32_t i = 0; i < h_len[f]; ++i)
        B200_CHECK(col[i].value >= 0.0f && col[i].value < 16777216.0f,  // InvalidCat (categorical.h:33-36)
                   "invalid category: negative or not representable as an integer in float32");
      const float feature_max = h_len[f] == 0 ? 0.0f : col[h_len[f] - 1].value;
      for (size_t c = 0; c <= static_cast<size_t>(feature_max); ++c) cuts.push_back(static_cast<float>(c));
    } else {
      QueryCutValues(h_entries.data() + (size_t)f * width, h_len[f], (size_t)sk->max_bin, &cuts);
    }
    B200_CHECK(total + cuts.size() <= cut_values_capacity, "cut_values buffer too small");
    std::memcpy(cut_values_out_host + total, cuts.data(), sizeof(float) * cuts.size());
    total += cuts.size();
    cut_ptrs_out_host[f + 1] = (uint32_t)total;
  }
  B200_API_END();
}

// In-core single-window sketch (DeviceSketchWithHessian on a SimpleDMatrix, hist_util.cu:362-397) = one push.
extern "C" int b200hist_sketch_dense(int device, const float* x_dev, uint64_t n, uint64_t ld, uint32_t n_features,
                                     int32_t max_bin, uint32_t* cut_ptrs_out_host, float* cut_values_out_host,
                                     uint64_t cut_values_capacity, void* stream) {
  b200hist_sketch_t sk = nullptr;
  int rc = b200hist_sketch_create(device, n_features, max_bin, &sk);
  if (rc == 0) rc = b200hist_sketch_push_dense(sk, x_dev, n, ld, stream);
  if (rc == 0) rc = b200hist_sketch_make_cuts(sk, cut_ptrs_out_host, cut_values_out_host, cut_values_capacity, nullptr, stream);
  if (sk) delete sk;  // not through the ABI: keep the first error message
  return rc;
}
