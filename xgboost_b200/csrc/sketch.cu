// Quantile sketch -> HistogramCuts for a dense batch resident on the device (in-core, unweighted, single
// sketch window): the cut finding that precedes Ellpack binning.
//
// Replaces (reference paths relative to /root/reference):
//   DeviceSketchWithHessian / ProcessWeightedBatch (thrust::sort of Entry{fidx, fvalue})   src/common/hist_util.cu:234-397
//   RequiredSampleCutsPerColumn, SketchEpsilon, WQuantileSketch::LimitSizeLevel            hist_util.cu:31-35, quantile.h:579-601,734-738
//   SketchContainer::Push -> PruneImpl + BinarySearchQueryIndex                            src/common/quantile.cu:36-61,108-153,340-378
//   SketchContainer::ScanInput (duplicate folding)                                         quantile.cu:380-418
//   SketchContainer::MakeCuts -> WQSummary::QueryCutValues (host, as in the reference)      quantile.cu:658-709, quantile.h:280-340
//
// For this case the reference's summary is fully determined by per-column ORDER STATISTICS: PruneImpl keeps
// `to = min(column_size, limit)` samples whose ranks follow from float32 arithmetic on (idx, to, n) alone.
// So the device work is (1) a transpose that turns the row-major batch into per-feature key columns
// (order-preserving uint32 image of the float, NaN dropped to the end), (2) a hand-written 4-pass LSD radix
// sort of every column at once (grid.y = feature; per-chunk digit histograms in shared memory, one scan CTA per
// column, stable scatter by warp match + per-warp digit counts), (3) a gather of the selected ranks.  The
// O(limit * F) tail -- duplicate folding and the max_bin rank queries -- runs on the host exactly like the
// reference's MakeCuts.  Cut values are bit-identical to the reference's on the golden inputs.
#include <algorithm>
#include <cmath>
#include <cstring>

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
  const uint64_t r0 = (uint64_t)blockIdx.x * 32;
  const uint32_t f0 = blockIdx.y * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
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
    if (tx == 0 && m && f < n_features) atomicAdd(&valid[f], (unsigned long long)__popc(m));
  }
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

__global__ void __launch_bounds__(256) PruneSelectKernel(const uint32_t* __restrict__ sorted_keys, uint64_t col_stride,
                                                          const unsigned long long* __restrict__ valid,
                                                          uint32_t per_col, uint32_t* __restrict__ out_value_bits,
                                                          uint32_t* __restrict__ out_rank) {
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
    out_value_bits[(uint64_t)f * per_col + idx] = KeyToFloatBits(col[sel]);
    out_rank[(uint64_t)f * per_col + idx] = sel;
  }
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

struct SummaryEntry {
  float rmin, rmax, wmin, value;
};

// WQSummary::QueryCutValues (quantile.h:280-340)
static void QueryCutValues(const std::vector<SummaryEntry>& s, size_t max_bin, std::vector<float>* out) {
  const size_t n = s.size();
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

extern "C" int b200hist_sketch_dense(int device, const float* x_dev, uint64_t n, uint64_t ld, uint32_t n_features,
                                     int32_t max_bin, uint32_t* cut_ptrs_out_host, float* cut_values_out_host,
                                     uint64_t cut_values_capacity, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(x_dev != nullptr && cut_ptrs_out_host != nullptr && cut_values_out_host != nullptr, "null argument");
  B200_CHECK(n_features > 0 && max_bin >= 2, "need n_features > 0 and max_bin >= 2");
  B200_CHECK(ld >= n_features, "leading dimension smaller than n_features");
  B200_CHECK(n < (1ull << 24), "float32 ranks are exact only below 2^24 rows per sketch window (hist_util.cu:87-128 "
                               "splits larger inputs into windows; merge across windows is not implemented)");
  DeviceGuard guard(device);
  auto s = static_cast<cudaStream_t>(stream);
  const double eps = 1.0 / static_cast<double>(std::min<size_t>((size_t)max_bin, std::max<size_t>(1, n)));
  const uint32_t per_col = (uint32_t)std::min<size_t>(LimitSizeLevel(n, eps), n);  // RequiredSampleCutsPerColumn
  std::vector<uint32_t> h_bits, h_rank;
  std::vector<unsigned long long> h_valid(n_features, 0);
  if (n > 0) {
    const uint64_t col_stride = (n + 3) & ~uint64_t(3);
    const uint32_t n_chunks = (uint32_t)((n + kSortChunk - 1) / kSortChunk);
    uint32_t* keys_a = DevAlloc<uint32_t>((uint64_t)n_features * col_stride);
    uint32_t* keys_b = DevAlloc<uint32_t>((uint64_t)n_features * col_stride);
    uint32_t* counts = DevAlloc<uint32_t>((uint64_t)n_features * n_chunks * 256);
    unsigned long long* valid = DevAlloc<unsigned long long>(n_features);
    uint32_t* d_bits = DevAlloc<uint32_t>((uint64_t)n_features * per_col);
    uint32_t* d_rank = DevAlloc<uint32_t>((uint64_t)n_features * per_col);
    B200_CUDA(cudaMemsetAsync(valid, 0, sizeof(unsigned long long) * n_features, s));
    dim3 tgrid((unsigned)((n + 31) / 32), (n_features + 31) / 32);
    KeysTransposeKernel<<<tgrid, 256, 0, s>>>(x_dev, n, ld, n_features, keys_a, col_stride, valid);
    B200_LAUNCHED();
    uint32_t* src = keys_a;
    uint32_t* dst = keys_b;
    dim3 sgrid(n_chunks, n_features);
    for (uint32_t shift = 0; shift < 32; shift += 8) {
      RadixCountKernel<<<sgrid, kSortThreads, 0, s>>>(src, n, col_stride, shift, counts, n_chunks);
      B200_LAUNCHED();
      RadixScanKernel<<<n_features, 256, 0, s>>>(counts, n_chunks);
      B200_LAUNCHED();
      RadixScatterKernel<<<sgrid, kSortThreads, 0, s>>>(src, dst, n, col_stride, shift, counts, n_chunks);
      B200_LAUNCHED();
      std::swap(src, dst);
    }
    dim3 pgrid((per_col + 255) / 256, n_features);
    PruneSelectKernel<<<pgrid, 256, 0, s>>>(src, col_stride, valid, per_col, d_bits, d_rank);
    B200_LAUNCHED();
    h_bits.resize((size_t)n_features * per_col);
    h_rank.resize((size_t)n_features * per_col);
    B200_CUDA(cudaMemcpyAsync(h_bits.data(), d_bits, sizeof(uint32_t) * h_bits.size(), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(h_rank.data(), d_rank, sizeof(uint32_t) * h_rank.size(), cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaMemcpyAsync(h_valid.data(), valid, sizeof(unsigned long long) * n_features, cudaMemcpyDeviceToHost, s));
    B200_CUDA(cudaStreamSynchronize(s));
    for (void* p : {(void*)keys_a, (void*)keys_b, (void*)counts, (void*)valid, (void*)d_bits, (void*)d_rank}) cudaFree(p);
  }
  // host tail (MakeCuts): fold duplicates (ScanInput) and answer the rank queries per column
  uint64_t total = 0;
  cut_ptrs_out_host[0] = 0;
  std::vector<SummaryEntry> summary;
  std::vector<float> cuts;
  for (uint32_t f = 0; f < n_features; ++f) {
    const uint32_t to = (uint32_t)std::min<unsigned long long>(h_valid[f], per_col);
    std::vector<SummaryEntry> e(to);
    for (uint32_t i = 0; i < to; ++i) {
      const uint32_t r = h_rank[(size_t)f * per_col + i];
      float v;
      std::memcpy(&v, &h_bits[(size_t)f * per_col + i], 4);
      e[i] = SummaryEntry{(float)r, (float)(r + 1), 1.0f, v};
    }
    for (int64_t i = (int64_t)to - 2; i >= 0; --i) {  // reverse scan: fold "first type" duplicates leftwards
      if (e[i].value - e[i + 1].value == 0 && e[i].rmin - e[i + 1].rmin != 0) {
        const float w = e[i].wmin + e[i + 1].wmin;
        e[i].wmin = w;
        e[i].rmax = e[i].rmin + w;
      }
    }
    summary.clear();
    for (uint32_t i = 0; i < to; ++i)
      if (i == 0 || e[i].value - e[i - 1].value != 0) summary.push_back(e[i]);  // SegmentedUnique keeps the first
    cuts.clear();
    QueryCutValues(summary, (size_t)max_bin, &cuts);
    B200_CHECK(total + cuts.size() <= cut_values_capacity, "cut_values buffer too small");
    std::memcpy(cut_values_out_host + total, cuts.data(), sizeof(float) * cuts.size());
    total += cuts.size();
    cut_ptrs_out_host[f + 1] = (uint32_t)total;
  }
  B200_API_END();
}
