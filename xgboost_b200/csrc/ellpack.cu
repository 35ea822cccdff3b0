// Ellpack page: quantile binning of raw feature values into the HBM-resident bin matrix.
//
// Replaces (reference paths relative to /root/reference):
//   HistogramCuts::SearchBin                     src/common/hist_util.h:112-121
//   CompressBinEllpackKernel / WriteCompressed.. src/data/ellpack_page.cu:56-134,240-298
//   CalcNumSymbols                               src/data/ellpack_page.cu:140-183
//   CompressedIterator (import only)             src/common/compressed_iterator.h:228-268
//
// Layout: bins[n_rows][row_bytes], one feature-local symbol per feature, uint8 when every feature
// needs <= 256 symbols (incl. the null symbol when the page may hold missing values) else uint16.
// row_bytes is padded to 16 so that every row is 16-byte aligned: a row (or a 16-byte aligned
// slice of it) is a legal source for cp.async.bulk (TMA 1-D bulk copy) and for 128-bit loads.
// The reference's own 8-bit layout starts at byte 8 (kPadding) and is therefore only 8-byte
// aligned -- which is why we keep our own copy instead of reading the reference buffer in place.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace b200 {

std::atomic<uint64_t> g_launch_count{0};
static thread_local std::string g_last_error;
void SetLastError(const std::string& msg) { g_last_error = msg; }

int NumSMs(int device) {
  int n = 0;
  B200_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  return n;
}

// One thread per (row, 4-feature word): four upper_bound searches, one 32-bit store (uint8 bins).
// Reads of X are coalesced along the feature dimension; cut values are L1/L2 resident (<= 256 KiB
// for the headline 256x256 configuration).
template <typename BinT>
__global__ void __launch_bounds__(256) BinDenseKernel(const float* __restrict__ x, uint64_t n, uint64_t ld,
                                                       uint32_t n_features, const uint32_t* __restrict__ cut_ptrs,
                                                       const float* __restrict__ cut_values, uint8_t* __restrict__ bins,
                                                       uint64_t row_begin, uint32_t row_bytes, uint32_t null_symbol) {
  constexpr int kPerWord = 4 / sizeof(BinT);
  const uint32_t words_per_row = row_bytes / 4;
  const uint64_t total = n * words_per_row;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = idx / words_per_row;
    uint32_t w = (uint32_t)(idx - r * words_per_row);
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < kPerWord; ++k) {
      uint32_t f = w * kPerWord + k;
      uint32_t sym = 0;
      if (f < n_features) {
        float v = x[r * ld + f];
        if (isnan(v)) {
          sym = null_symbol;
        } else {
          uint32_t beg = cut_ptrs[f], end = cut_ptrs[f + 1];
          uint32_t lo = beg, hi = end;  // first cut strictly greater than v
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (!(v < cut_values[mid])) lo = mid + 1; else hi = mid;
          }
          if (lo == end) lo = end - 1;
          sym = lo - beg;
        }
      }
      packed |= (sym & ((1u << (8 * sizeof(BinT))) - 1u)) << (8 * sizeof(BinT) * k);
    }
    *reinterpret_cast<uint32_t*>(bins + (row_begin + r) * row_bytes + 4ull * w) = packed;
  }
}

// Categorical columns (HistogramCuts::SearchCatBin, src/common/hist_util.h:132-144): the bin of a category is its
// position among the feature's cut values (= category codes), found by lower bound on the truncated value and
// clamped to the last bin.  Runs after the numerical binning kernel and rewrites only the categorical columns'
// symbols: one thread per (row, categorical column).
template <typename BinT>
__global__ void __launch_bounds__(256) RebinCategoricalKernel(const float* __restrict__ x, uint64_t n, uint64_t ld,
                                                               const uint32_t* __restrict__ cat_features,
                                                               uint32_t n_cat, const uint32_t* __restrict__ cut_ptrs,
                                                               const float* __restrict__ cut_values,
                                                               uint8_t* __restrict__ bins, uint64_t row_begin,
                                                               uint32_t row_bytes, uint32_t null_symbol) {
  const uint64_t total = n * n_cat;
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx / n_cat;
    const uint32_t f = cat_features[idx - r * n_cat];
    const float raw = x[r * ld + f];
    uint32_t sym = null_symbol;
    if (!isnan(raw)) {
      const float v = (float)(int)raw;  // AsCat: truncation (categorical.h:25-31)
      const uint32_t beg = cut_ptrs[f], end = cut_ptrs[f + 1];
      uint32_t lo = beg, hi = end;  // first cut not less than v
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cut_values[mid] < v) lo = mid + 1; else hi = mid;
      }
      if (lo == end) lo = end - 1;
      sym = lo - beg;
    }
    reinterpret_cast<BinT*>(bins + (row_begin + r) * row_bytes)[f] = (BinT)sym;
  }
}

// Faster binning for uint8 pages: a CTA owns a block of 32 features, stages their cut values in shared memory
// (<= 32 KB at 256 bins) and walks the rows; the 8-step binary searches then run out of shared memory instead
// of L2.  Thread = (row, 4-feature word); a warp reads 4 rows x 128 contiguous bytes of X.
__global__ void __launch_bounds__(256) BinDenseSmemKernel(const float* __restrict__ x, uint64_t n, uint64_t ld,
                                                           uint32_t n_features, const uint32_t* __restrict__ cut_ptrs,
                                                           const float* __restrict__ cut_values,
                                                           uint8_t* __restrict__ bins, uint64_t row_begin,
                                                           uint32_t row_bytes, uint32_t null_symbol) {
  B200_DYNAMIC_SHARED(float, s_cuts);
  __shared__ uint32_t s_ptr[33];
  const uint32_t f0 = blockIdx.y * 32;
  const uint32_t nf = min(32u, n_features - f0);
  if (threadIdx.x <= nf) s_ptr[threadIdx.x] = cut_ptrs[f0 + threadIdx.x] - cut_ptrs[f0];
  __syncthreads();
  const uint32_t n_cuts = s_ptr[nf];
  for (uint32_t i = threadIdx.x; i < n_cuts; i += 256) s_cuts[i] = cut_values[cut_ptrs[f0] + i];
  __syncthreads();
  const uint32_t w = threadIdx.x & 7u;  // word inside the feature block
  for (uint64_t r = (uint64_t)blockIdx.x * 32 + (threadIdx.x >> 3); r < n; r += (uint64_t)gridDim.x * 32) {
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t fl = 4 * w + k;
      uint32_t sym = 0;
      if (fl < nf) {
        const float v = x[r * ld + f0 + fl];
        if (isnan(v)) {
          sym = null_symbol;
        } else {
          const uint32_t beg = s_ptr[fl], end = s_ptr[fl + 1];
          uint32_t lo = beg, hi = end;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (!(v < s_cuts[mid])) lo = mid + 1; else hi = mid;
          }
          if (lo == end) lo = end - 1;
          sym = lo - beg;
        }
      }
      packed |= (sym & 0xffu) << (8 * k);
    }
    if (f0 + 4 * w < row_bytes)
      *reinterpret_cast<uint32_t*>(bins + (row_begin + r) * row_bytes + f0 + 4ull * w) = packed;
  }
}

// Column-major copy of rows [row_begin, row_begin + n): bins_t[f][r].  128 rows x 64 features per CTA through
// shared memory; reads are 32-bit row-major words, writes are 128 contiguous bytes per feature.
__global__ void __launch_bounds__(256) TransposeKernel(const uint8_t* __restrict__ bins, uint32_t row_bytes,
                                                        uint32_t n_features, uint64_t row_begin, uint64_t n,
                                                        uint8_t* __restrict__ bins_t, uint64_t col_stride) {
  __shared__ uint8_t tile[64][128 + 4];
  const uint64_t r0 = row_begin + (uint64_t)blockIdx.x * 128;
  const uint32_t f0 = blockIdx.y * 64;
  const uint64_t r_end = row_begin + n;
  // load: 128 rows x 16 words
  for (uint32_t i = threadIdx.x; i < 128 * 16; i += 256) {
    const uint32_t r = i >> 4, w = i & 15u;
    uint32_t word = 0;
    if (r0 + r < r_end && f0 + 4 * w < row_bytes)
      word = *reinterpret_cast<const uint32_t*>(bins + (r0 + r) * row_bytes + f0 + 4 * w);
#pragma unroll
    for (int k = 0; k < 4; ++k) tile[4 * w + k][r] = (uint8_t)(word >> (8 * k));
  }
  __syncthreads();
  // store: 64 features x 32 words of 4 rows
  for (uint32_t i = threadIdx.x; i < 64 * 32; i += 256) {
    const uint32_t f = i >> 5, w = i & 31u;
    if (f0 + f < n_features && r0 + 4 * w < r_end) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&tile[f][4 * w]);
      uint8_t* dst = bins_t + (uint64_t)(f0 + f) * col_stride + r0 + 4 * w;
      if (r0 + 4 * w + 4 <= r_end) {
        *reinterpret_cast<uint32_t*>(dst) = v;
      } else {
        for (uint32_t k = 0; r0 + 4 * w + k < r_end; ++k) dst[k] = (uint8_t)(v >> (8 * k));
      }
    }
  }
}

// The same for uint16 pages: bins_t is uint16[f][col_stride]; 128 rows x 32 features per CTA.
__global__ void __launch_bounds__(256) TransposeKernel16(const uint8_t* __restrict__ bins, uint32_t row_bytes,
                                                          uint32_t n_features, uint64_t row_begin, uint64_t n,
                                                          uint16_t* __restrict__ bins_t, uint64_t col_stride) {
  __shared__ uint16_t tile[32][128 + 2];
  const uint64_t r0 = row_begin + (uint64_t)blockIdx.x * 128;
  const uint32_t f0 = blockIdx.y * 32;
  const uint64_t r_end = row_begin + n;
  for (uint32_t i = threadIdx.x; i < 128 * 16; i += 256) {  // 128 rows x 16 words of 2 symbols
    const uint32_t r = i >> 4, w = i & 15u;
    uint32_t word = 0;
    if (r0 + r < r_end && (f0 + 2 * w) * 2 < row_bytes)
      word = *reinterpret_cast<const uint32_t*>(bins + (r0 + r) * row_bytes + (f0 + 2 * w) * 2);
    tile[2 * w][r] = (uint16_t)(word & 0xFFFFu);
    tile[2 * w + 1][r] = (uint16_t)(word >> 16);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 32 * 64; i += 256) {  // 32 features x 64 words of 2 rows
    const uint32_t f = i >> 6, w = i & 63u;
    if (f0 + f < n_features && r0 + 2 * w < r_end) {
      uint16_t* dst = bins_t + (uint64_t)(f0 + f) * col_stride + r0 + 2 * w;
      if (r0 + 2 * w + 2 <= r_end) {
        *reinterpret_cast<uint32_t*>(dst) = (uint32_t)tile[f][2 * w] | ((uint32_t)tile[f][2 * w + 1] << 16);
      } else {
        dst[0] = tile[f][2 * w];
      }
    }
  }
}

#ifndef B200_KERNELS_ONLY
// External memory: scratch for one incoming batch in OUR row layout, and its copy-out into the pinned host image.
static uint8_t* ImportScratch(b200hist_ellpack_s* e, uint64_t n) {
  const size_t need = (size_t)n * e->row_bytes + 16;
  if (need > e->import_bytes) {
    if (e->d_import) cudaFree(e->d_import);
    e->d_import = nullptr;
    e->d_import = DevAlloc<uint8_t>(need);
    e->import_bytes = need;
  }
  return e->d_import;
}
static void ExportImported(b200hist_ellpack_s* e, uint64_t row_begin, uint64_t n, cudaStream_t s) {
  B200_CUDA(cudaMemcpyAsync(e->h_bins + row_begin * (uint64_t)e->row_bytes, e->d_import, n * (uint64_t)e->row_bytes,
                            cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));  // the scratch is reused by the next batch
}

static void LaunchTranspose(b200hist_ellpack_s* e, uint64_t row_begin, uint64_t n, cudaStream_t s) {
  if (e->d_bins_t == nullptr || n == 0) return;
  // tiles are aligned to 128 rows of the page so that 32-bit stores stay aligned
  const uint64_t a0 = row_begin & ~uint64_t(127);
  const uint64_t span = row_begin + n - a0;
  if (e->bin_bytes == 2) {
    dim3 grid((unsigned)((span + 127) / 128), (e->n_features + 31) / 32);
    TransposeKernel16<<<grid, 256, 0, s>>>(e->d_bins, e->row_bytes, e->n_features, a0, span,
                                           reinterpret_cast<uint16_t*>(e->d_bins_t), e->col_stride);
  } else {
    dim3 grid((unsigned)((span + 127) / 128), (e->n_features + 63) / 64);
    TransposeKernel<<<grid, 256, 0, s>>>(e->d_bins, e->row_bytes, e->n_features, a0, span, e->d_bins_t, e->col_stride);
  }
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// Sparse layout: one warp per CSR row; entry k -> global bin of (feature, value), compacted over NaNs, padded
// with the null symbol.  Entries stay in feature order, so a row can be binary-searched by feature.
__global__ void __launch_bounds__(256) BinCsrKernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                     const float* __restrict__ values, uint64_t n,
                                                     const uint32_t* __restrict__ cut_ptrs,
                                                     const float* __restrict__ cut_values, uint32_t* __restrict__ ell,
                                                     uint64_t row_begin, uint32_t row_stride, uint32_t null_symbol,
                                                     uint32_t n_features, int* __restrict__ error_flag) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n; r += n_warps) {
    const int64_t beg = indptr[r], end = indptr[r + 1];
    uint32_t* out = ell + (row_begin + r) * row_stride;
    uint32_t written = 0;
    for (int64_t base = beg; base < end; base += 32) {
      const int64_t i = base + lane;
      uint32_t sym = null_symbol;
      bool keep = false;
      if (i < end) {
        const float v = values[i];
        const uint32_t f = (uint32_t)indices[i];
        if (f >= n_features) {
          *error_flag = 1;
        } else if (!isnan(v)) {
          const uint32_t cb = cut_ptrs[f], ce = cut_ptrs[f + 1];
          uint32_t lo = cb, hi = ce;
          while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (!(v < cut_values[mid])) lo = mid + 1; else hi = mid;
          }
          if (lo == ce) lo = ce - 1;
          sym = lo;
          keep = true;
        }
      }
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      const uint32_t pos = written + __popc(m & ((1u << lane) - 1u));
      if (keep) {
        if (pos < row_stride) out[pos] = sym; else *error_flag = 2;
      }
      written += __popc(m);
    }
    for (uint32_t k = written + lane; k < row_stride; k += 32) out[k] = null_symbol;
  }
}

// Presence signature of sparse rows [row_begin, row_begin + n): bit (f mod 64) for every feature the row holds.
__global__ void __launch_bounds__(256) RowSignatureKernel(const uint32_t* __restrict__ ell, uint64_t row_begin, uint64_t n,
                                                           uint32_t row_stride, const uint32_t* __restrict__ cut_ptrs,
                                                           uint32_t n_features, uint32_t total_bins,
                                                           unsigned long long* __restrict__ sig) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t* __restrict__ row = ell + (row_begin + i) * row_stride;
    unsigned long long m = 0ull;
    for (uint32_t k = 0; k < row_stride; ++k) {
      const uint32_t sym = row[k];
      if (sym >= total_bins) continue;  // null padding
      uint32_t lo = 0, hi = n_features;  // feature whose [cut_ptrs[f], cut_ptrs[f+1]) holds sym
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cut_ptrs[mid] <= sym) lo = mid; else hi = mid;
      }
      m |= 1ull << (lo & 63u);
    }
    sig[row_begin + i] = m;
  }
}

// Reference sparse page (global-bin symbols, null = total_bins) -> uint32 ELL.
__global__ void __launch_bounds__(256) ImportReferenceSparseKernel(const uint8_t* __restrict__ buf, uint64_t total,
                                                                    uint32_t symbol_bits, uint32_t ref_stride,
                                                                    uint32_t our_stride, uint32_t* __restrict__ ell) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t bit0 = i * symbol_bits;
    uint64_t v = 0;
    uint64_t byte0 = bit0 >> 3;
    uint32_t shift = (uint32_t)(bit0 & 7);
    uint32_t nbytes = (shift + symbol_bits + 7) >> 3;
    for (uint32_t k = 0; k < nbytes; ++k) v = (v << 8) | buf[8 + byte0 + k];
    const uint64_t row = i / ref_stride;
    ell[row * our_stride + (i - row * ref_stride)] =
        (uint32_t)((v >> (8 * nbytes - shift - symbol_bits)) & ((1ull << symbol_bits) - 1ull));
  }
}

// Decode the reference's big-endian bit-packed symbol stream into our layout.
template <typename BinT>
__global__ void __launch_bounds__(256) ImportReferenceKernel(const uint8_t* __restrict__ buf, uint64_t n_rows,
                                                              uint32_t n_features, uint32_t symbol_bits,
                                                              uint32_t ref_null, uint8_t* __restrict__ bins,
                                                              uint32_t row_bytes, uint32_t null_symbol) {
  const uint64_t total = n_rows * n_features;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t bit0 = i * symbol_bits;
    uint64_t v = 0;
    uint64_t byte0 = bit0 >> 3;
    uint32_t shift = (uint32_t)(bit0 & 7);
    uint32_t nbytes = (shift + symbol_bits + 7) >> 3;                 // bytes that hold the symbol
    for (uint32_t k = 0; k < nbytes; ++k) v = (v << 8) | buf[8 + byte0 + k];  // kPadding = 8
    uint32_t sym = (uint32_t)((v >> (8 * nbytes - shift - symbol_bits)) & ((1ull << symbol_bits) - 1ull));
    uint64_t r = i / n_features;
    uint32_t f = (uint32_t)(i - r * n_features);
    if (sym == ref_null) sym = null_symbol;
    reinterpret_cast<BinT*>(bins + r * row_bytes)[f] = (BinT)sym;
  }
}

}  // namespace b200

#ifndef B200_KERNELS_ONLY
using namespace b200;

extern "C" {

const char* b200hist_last_error(void) { return g_last_error.c_str(); }
int b200hist_version(void) { return B200HIST_VERSION; }
uint64_t b200hist_launch_count(void) { return g_launch_count.load(); }

static int EllpackCreateDense(int device, uint64_t n_rows, uint32_t n_features, const uint32_t* cut_ptrs_host,
                              const float* cut_values_host, int has_missing, uint64_t page_rows,
                              b200hist_ellpack_t* out) {
  B200_API_BEGIN();
  B200_CHECK(out != nullptr && cut_ptrs_host != nullptr && cut_values_host != nullptr, "null argument");
  B200_CHECK(n_features > 0, "n_features must be positive");
  B200_CHECK(n_rows < (1ull << 32), "RowIndexT is uint32 (row_partitioner.cuh:22): n_rows must be < 2^32");
  DeviceGuard guard(device);
  auto* e = new b200hist_ellpack_s();
  e->device = device;
  e->n_rows = n_rows;
  e->n_features = n_features;
  e->cut_ptrs.assign(cut_ptrs_host, cut_ptrs_host + n_features + 1);
  e->total_bins = e->cut_ptrs[n_features];
  B200_CHECK(e->cut_ptrs[0] == 0, "cut_ptrs[0] must be 0");
  e->cut_values.assign(cut_values_host, cut_values_host + e->total_bins);
  uint32_t mx = 0;
  for (uint32_t f = 0; f < n_features; ++f) {
    B200_CHECK(e->cut_ptrs[f + 1] > e->cut_ptrs[f], "every feature needs at least one bin");
    mx = std::max(mx, e->cut_ptrs[f + 1] - e->cut_ptrs[f]);
  }
  e->max_feature_bins = mx;
  e->has_missing = has_missing != 0;
  uint32_t n_symbols = mx + (e->has_missing ? 1 : 0);  // CalcNumSymbols (ellpack_page.cu:140-183)
  B200_CHECK(n_symbols <= 65536, "more than 65536 symbols per feature is not supported");
  e->bin_bytes = n_symbols <= 256 ? 1 : 2;
  // null = the first symbol no feature uses (the reference's NullValue for dense pages with missing values is
  // n_symbols - 1 as well, ellpack_page.cu:140-183): the histogram kernels give it the accumulator slot mx
  e->null_symbol = e->has_missing ? mx : 0xFFFFFFFFu;
  // rows padded to whole 32-feature blocks (zero symbols), so the histogram kernel's unpredicated 32-feature loop
  // also serves feature counts that are not multiples of 32 (the padding is dropped at flush time)
  e->row_bytes = ((n_features + 31u) & ~31u) * e->bin_bytes;
  e->d_cut_ptrs = DevAlloc<uint32_t>(n_features + 1);
  e->d_cut_values = DevAlloc<float>(e->total_bins);
  B200_CUDA(cudaMemcpy(e->d_cut_ptrs, e->cut_ptrs.data(), sizeof(uint32_t) * (n_features + 1), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(e->d_cut_values, e->cut_values.data(), sizeof(float) * e->total_bins, cudaMemcpyHostToDevice));
  if (page_rows > 0) {  // external memory: pinned host storage + two device staging buffers of one page each
    e->paged = true;
    e->page_rows = std::min<uint64_t>(page_rows, std::max<uint64_t>(n_rows, 1));
    const size_t host_bytes = (size_t)n_rows * e->row_bytes + 16;
    B200_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&e->h_bins), host_bytes, cudaHostAllocPortable));
    std::memset(e->h_bins, 0, host_bytes);
    const size_t stage_bytes = (size_t)e->page_rows * e->row_bytes + 16;
    for (int b = 0; b < 2; ++b) {
      e->d_stage[b] = DevAlloc<uint8_t>(stage_bytes);
      cudaEvent_t ev;
      B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->ev_ready[b] = ev;
      B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->ev_done[b] = ev;
    }
    cudaStream_t cs;
    B200_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    e->copy_stream = cs;
    *out = e;
    return 0;
  }
  e->d_bins = DevAlloc<uint8_t>(n_rows * (uint64_t)e->row_bytes + 16);
  B200_CUDA(cudaMemset(e->d_bins, 0, n_rows * (uint64_t)e->row_bytes + 16));
  // Column-major copy in the page's symbol width (col_stride in symbols): the partitioner's single-feature gathers
  // read it (32 consecutive rows share a sector: 4.9 instead of 16 ms of partition per tree at 100M x 256).  It
  // doubles the page, so it is an accelerator, not a requirement: kept when it fits next to the updater's state
  // (gradient pairs, row ids, positions: ~40 B per row, plus the histogram cache) with a fifth of the device to
  // spare, dropped otherwise -- the partition kernels then gather from the row-major page.
  // B200HIST_NO_COLMAJOR=1 drops it always, =0 keeps it always.
  const char* no_t = std::getenv("B200HIST_NO_COLMAJOR");
  const uint64_t col_stride = (n_rows + 127) & ~uint64_t(127);
  const uint64_t copy_bytes = (uint64_t)n_features * col_stride * e->bin_bytes + 16;
  bool want_copy = true;
  if (no_t && no_t[0] == '1') {
    want_copy = false;
  } else if (!(no_t && no_t[0] == '0')) {
    size_t free_b = 0, total_b = 0;
    B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t state_bytes = n_rows * 40ull + (1ull << 30);
    want_copy = (uint64_t)free_b >= copy_bytes + state_bytes + (uint64_t)total_b / 5;
  }
  if (want_copy) {
    e->col_stride = col_stride;
    e->d_bins_t = DevAlloc<uint8_t>(copy_bytes);
  }
  *out = e;
  B200_API_END();
}

int b200hist_ellpack_create(int device, uint64_t n_rows, uint32_t n_features, const uint32_t* cut_ptrs_host,
                            const float* cut_values_host, int has_missing, b200hist_ellpack_t* out) {
  return EllpackCreateDense(device, n_rows, n_features, cut_ptrs_host, cut_values_host, has_missing, 0, out);
}

int b200hist_ellpack_create_paged(int device, uint64_t n_rows, uint32_t n_features, const uint32_t* cut_ptrs_host,
                                  const float* cut_values_host, int has_missing, uint64_t page_rows,
                                  b200hist_ellpack_t* out) {
  if (page_rows == 0) {
    ::b200::SetLastError("b200hist_ellpack_create_paged: page_rows must be positive");
    return -1;
  }
  return EllpackCreateDense(device, n_rows, n_features, cut_ptrs_host, cut_values_host, has_missing, page_rows, out);
}

int b200hist_ellpack_paging_info(b200hist_ellpack_t e, uint32_t* n_pages, uint64_t* page_rows, uint64_t* h2d_bytes) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr, "null ellpack");
  if (n_pages) *n_pages = e->paged ? e->NumPages() : 0u;
  if (page_rows) *page_rows = e->paged ? e->page_rows : 0ull;
  if (h2d_bytes) *h2d_bytes = e->h2d_bytes;
  B200_API_END();
}

int b200hist_ellpack_create_sparse(int device, uint64_t n_rows, uint32_t n_features, const uint32_t* cut_ptrs_host,
                                   const float* cut_values_host, uint32_t row_stride, b200hist_ellpack_t* out) {
  B200_API_BEGIN();
  B200_CHECK(out != nullptr && cut_ptrs_host != nullptr && cut_values_host != nullptr, "null argument");
  B200_CHECK(n_features > 0, "n_features must be positive");
  B200_CHECK(n_rows < (1ull << 32), "RowIndexT is uint32 (row_partitioner.cuh:22): n_rows must be < 2^32");
  B200_CHECK(row_stride > 0 && row_stride <= n_features, "row_stride must be in [1, n_features]");
  DeviceGuard guard(device);
  auto* e = new b200hist_ellpack_s();
  e->device = device;
  e->n_rows = n_rows;
  e->n_features = n_features;
  e->cut_ptrs.assign(cut_ptrs_host, cut_ptrs_host + n_features + 1);
  e->total_bins = e->cut_ptrs[n_features];
  B200_CHECK(e->cut_ptrs[0] == 0, "cut_ptrs[0] must be 0");
  e->cut_values.assign(cut_values_host, cut_values_host + e->total_bins);
  for (uint32_t f = 0; f < n_features; ++f) {
    B200_CHECK(e->cut_ptrs[f + 1] > e->cut_ptrs[f], "every feature needs at least one bin");
    e->max_feature_bins = std::max(e->max_feature_bins, e->cut_ptrs[f + 1] - e->cut_ptrs[f]);
  }
  e->has_missing = true;
  e->is_sparse = true;
  e->bin_bytes = 4;
  e->null_symbol = e->total_bins;  // NullValue of a sparse page (ellpack_page.cu:147)
  e->row_stride = (row_stride + 3u) & ~3u;  // rows stay 16-byte aligned
  e->row_bytes = e->row_stride * 4u;
  e->d_cut_ptrs = DevAlloc<uint32_t>(n_features + 1);
  e->d_cut_values = DevAlloc<float>(e->total_bins);
  B200_CUDA(cudaMemcpy(e->d_cut_ptrs, e->cut_ptrs.data(), sizeof(uint32_t) * (n_features + 1), cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(e->d_cut_values, e->cut_values.data(), sizeof(float) * e->total_bins, cudaMemcpyHostToDevice));
  e->d_bins = DevAlloc<uint8_t>(n_rows * (uint64_t)e->row_bytes + 16);
  // every slot starts as 0xFFFFFFFF; kernels treat any symbol >= total_bins as null
  B200_CUDA(cudaMemset(e->d_bins, 0xFF, n_rows * (uint64_t)e->row_bytes + 16));
  static const bool no_sig = std::getenv("B200HIST_NO_ROWSIG") != nullptr;
  if (!no_sig) {  // all-ones until the rows are written: "may hold any feature" is always safe
    e->d_row_sig = DevAlloc<unsigned long long>(n_rows);
    B200_CUDA(cudaMemset(e->d_row_sig, 0xFF, sizeof(unsigned long long) * n_rows));
  }
  *out = e;
  B200_API_END();
}

int b200hist_ellpack_push_csr(b200hist_ellpack_t e, uint64_t row_begin, const int64_t* indptr_dev,
                              const int32_t* indices_dev, const float* values_dev, uint64_t n, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr && indptr_dev != nullptr, "null argument");
  B200_CHECK(e->is_sparse, "push_csr needs a page created with b200hist_ellpack_create_sparse");
  B200_CHECK(row_begin + n <= e->n_rows, "batch exceeds the page");
  if (n == 0) return 0;
  DeviceGuard guard(e->device);
  auto s = static_cast<cudaStream_t>(stream);
  int* d_err = DevAlloc<int>(1);
  B200_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), s));
  int blocks = (int)std::min<uint64_t>((n * 32 + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
  BinCsrKernel<<<blocks, 256, 0, s>>>(indptr_dev, indices_dev, values_dev, n, e->d_cut_ptrs, e->d_cut_values,
                                      reinterpret_cast<uint32_t*>(e->d_bins), row_begin, e->row_stride, e->null_symbol,
                                      e->n_features, d_err);
  B200_LAUNCHED();
  if (e->d_row_sig) {
    int sb = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
    RowSignatureKernel<<<sb, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(e->d_bins), row_begin, n, e->row_stride,
                                          e->d_cut_ptrs, e->n_features, e->total_bins, e->d_row_sig);
    B200_LAUNCHED();
  }
  int h_err = 0;
  B200_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_err);
  B200_CHECK(h_err != 1, "CSR column index out of range");
  B200_CHECK(h_err != 2, "a CSR row holds more entries than row_stride");
  B200_API_END();
}

int b200hist_ellpack_destroy(b200hist_ellpack_t e) {
  B200_API_BEGIN();
  if (e) {
    DeviceGuard guard(e->device);
    cudaFree(e->d_cut_ptrs);
    cudaFree(e->d_cut_values);
    cudaFree(e->d_bins);
    cudaFree(e->d_bins_t);
    if (e->paged) {
      if (e->copy_stream) {
        cudaStreamSynchronize(static_cast<cudaStream_t>(e->copy_stream));
        cudaStreamDestroy(static_cast<cudaStream_t>(e->copy_stream));
      }
      for (int b = 0; b < 2; ++b) {
        cudaFree(e->d_stage[b]);
        if (e->ev_ready[b]) cudaEventDestroy(static_cast<cudaEvent_t>(e->ev_ready[b]));
        if (e->ev_done[b]) cudaEventDestroy(static_cast<cudaEvent_t>(e->ev_done[b]));
      }
      cudaFree(e->d_import);
      cudaFreeHost(e->h_bins);
    }
    if (e->d_row_sig) cudaFree(e->d_row_sig);
    if (e->d_cat_features) cudaFree(e->d_cat_features);
    delete e;
  }
  B200_API_END();
}

int b200hist_ellpack_push_dense(b200hist_ellpack_t e, uint64_t row_begin, const float* x_dev, uint64_t n,
                                uint64_t ld, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr && x_dev != nullptr, "null argument");
  B200_CHECK(!e->is_sparse, "push_dense needs a dense-layout page");
  B200_CHECK(row_begin + n <= e->n_rows, "batch exceeds the page");
  B200_CHECK(ld >= e->n_features, "leading dimension smaller than n_features");
  if (n == 0) return 0;
  DeviceGuard guard(e->device);
  auto s = static_cast<cudaStream_t>(stream);
  // external memory: the batch is binned into the import scratch (biased so that the kernels' row_begin indexing
  // lands at its start) and copied out to the pinned host image afterwards
  uint8_t* const dst_bins = e->paged ? ImportScratch(e, n) - row_begin * (uint64_t)e->row_bytes : e->d_bins;
  if (e->paged) B200_CUDA(cudaMemsetAsync(e->d_import, 0, n * (uint64_t)e->row_bytes, s));
  uint64_t total = n * (e->row_bytes / 4);
  int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
  uint32_t max_block_cuts = 0;  // cut values of the widest 32-feature block
  for (uint32_t f0 = 0; f0 < e->n_features; f0 += 32)
    max_block_cuts = std::max(max_block_cuts, e->cut_ptrs[std::min(e->n_features, f0 + 32)] - e->cut_ptrs[f0]);
  if (e->bin_bytes == 1 && max_block_cuts * sizeof(float) <= 40 * 1024) {
    dim3 grid((unsigned)std::min<uint64_t>((n + 31) / 32, (uint64_t)NumSMs(e->device) * 8), (e->n_features + 31) / 32);
    BinDenseSmemKernel<<<grid, 256, max_block_cuts * sizeof(float), s>>>(x_dev, n, ld, e->n_features, e->d_cut_ptrs,
                                                                        e->d_cut_values, dst_bins, row_begin,
                                                                        e->row_bytes, e->null_symbol);
  } else if (e->bin_bytes == 1) {
    BinDenseKernel<uint8_t><<<blocks, 256, 0, s>>>(x_dev, n, ld, e->n_features, e->d_cut_ptrs, e->d_cut_values,
                                                  dst_bins, row_begin, e->row_bytes, e->null_symbol);
  } else {
    BinDenseKernel<uint16_t><<<blocks, 256, 0, s>>>(x_dev, n, ld, e->n_features, e->d_cut_ptrs, e->d_cut_values,
                                                   dst_bins, row_begin, e->row_bytes, e->null_symbol);
  }
  B200_LAUNCHED();
  if (e->d_cat_features != nullptr) {
    const uint64_t cat_total = n * (uint64_t)e->n_cat_features;
    const int cb = (int)std::min<uint64_t>((cat_total + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
    if (e->bin_bytes == 1) {
      RebinCategoricalKernel<uint8_t><<<cb, 256, 0, s>>>(x_dev, n, ld, e->d_cat_features, e->n_cat_features,
                                                         e->d_cut_ptrs, e->d_cut_values, dst_bins, row_begin,
                                                         e->row_bytes, e->null_symbol);
    } else {
      RebinCategoricalKernel<uint16_t><<<cb, 256, 0, s>>>(x_dev, n, ld, e->d_cat_features, e->n_cat_features,
                                                          e->d_cut_ptrs, e->d_cut_values, dst_bins, row_begin,
                                                          e->row_bytes, e->null_symbol);
    }
    B200_LAUNCHED();
  }
  if (e->paged) ExportImported(e, row_begin, n, s);
  LaunchTranspose(e, row_begin, n, s);
  B200_API_END();
}

int b200hist_ellpack_set_feature_types(b200hist_ellpack_t e, const uint8_t* feature_is_cat_host, uint32_t n) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr, "null argument");
  B200_CHECK(!e->is_sparse, "feature types are supported on dense-layout pages only");
  B200_CHECK(n <= e->n_features, "more feature types than features");
  DeviceGuard guard(e->device);
  if (e->d_cat_features) cudaFree(e->d_cat_features);
  e->d_cat_features = nullptr;
  e->n_cat_features = 0;
  std::vector<uint32_t> ids;
  for (uint32_t f = 0; feature_is_cat_host != nullptr && f < n; ++f)
    if (feature_is_cat_host[f]) ids.push_back(f);
  if (!ids.empty()) {
    e->d_cat_features = DevAlloc<uint32_t>(ids.size());
    B200_CUDA(cudaMemcpy(e->d_cat_features, ids.data(), sizeof(uint32_t) * ids.size(), cudaMemcpyHostToDevice));
    e->n_cat_features = (uint32_t)ids.size();
  }
  B200_API_END();
}

int b200hist_ellpack_import_reference_rows(b200hist_ellpack_t e, uint64_t row_begin, uint64_t n_rows,
                                           const uint8_t* gidx_buffer_dev, uint32_t row_stride, uint32_t n_symbols,
                                           int is_dense, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr && gidx_buffer_dev != nullptr, "null argument");
  B200_CHECK(row_begin + n_rows <= e->n_rows, "reference page exceeds this page");
  if (n_rows == 0) return 0;
  if (e->is_sparse) {  // global-bin symbols, n_symbols = total_bins + 1, null = total_bins
    B200_CHECK(n_symbols == e->total_bins + 1, "sparse reference page: n_symbols must be total_bins + 1");
    B200_CHECK(row_stride <= e->row_stride, "the reference page's row_stride exceeds this page's");
    uint32_t sbits = 1;
    while ((1ull << sbits) < n_symbols) ++sbits;
    DeviceGuard guard_s(e->device);
    auto ss = static_cast<cudaStream_t>(stream);
    uint64_t tot = n_rows * (uint64_t)row_stride;
    int nb = (int)std::min<uint64_t>((tot + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
    // slots beyond the reference's row_stride keep the fill value (>= total_bins == null)
    ImportReferenceSparseKernel<<<nb, 256, 0, ss>>>(gidx_buffer_dev, tot, sbits, row_stride, e->row_stride,
                                                    reinterpret_cast<uint32_t*>(e->d_bins) + row_begin * e->row_stride);
    B200_LAUNCHED();
    if (e->d_row_sig) {
      int sb = (int)std::min<uint64_t>((n_rows + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
      RowSignatureKernel<<<sb, 256, 0, ss>>>(reinterpret_cast<const uint32_t*>(e->d_bins), row_begin, n_rows,
                                             e->row_stride, e->d_cut_ptrs, e->n_features, e->total_bins, e->d_row_sig);
      B200_LAUNCHED();
    }
    return 0;
  }
  B200_CHECK(row_stride == e->n_features, "dense-layout reference pages need row_stride == n_features");
  uint32_t bits = 1;
  while ((1ull << bits) < n_symbols) ++bits;  // SymbolBits (compressed_iterator.h:37-40)
  // dense & no missing: n_symbols = max bins/feature; relatively dense: +1 with null = n_symbols-1
  uint32_t ref_null = is_dense ? 0xFFFFFFFFu : n_symbols - 1;
  B200_CHECK(is_dense || e->has_missing, "page with missing values needs an ellpack created with has_missing=1");
  DeviceGuard guard(e->device);
  auto s = static_cast<cudaStream_t>(stream);
  uint64_t total = n_rows * (uint64_t)e->n_features;
  int blocks = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)NumSMs(e->device) * 16);
  uint8_t* dst = e->paged ? ImportScratch(e, n_rows) : e->d_bins + row_begin * (uint64_t)e->row_bytes;
  if (e->paged) B200_CUDA(cudaMemsetAsync(e->d_import, 0, n_rows * (uint64_t)e->row_bytes, s));
  if (e->bin_bytes == 1) {
    ImportReferenceKernel<uint8_t><<<blocks, 256, 0, s>>>(gidx_buffer_dev, n_rows, e->n_features, bits, ref_null, dst,
                                                         e->row_bytes, e->null_symbol);
  } else {
    ImportReferenceKernel<uint16_t><<<blocks, 256, 0, s>>>(gidx_buffer_dev, n_rows, e->n_features, bits, ref_null, dst,
                                                          e->row_bytes, e->null_symbol);
  }
  B200_LAUNCHED();
  if (e->paged) ExportImported(e, row_begin, n_rows, s);
  LaunchTranspose(e, row_begin, n_rows, s);
  B200_API_END();
}

int b200hist_ellpack_import_reference(b200hist_ellpack_t e, const uint8_t* gidx_buffer_dev, uint32_t row_stride,
                                      uint32_t n_symbols, int is_dense, void* stream) {
  if (e == nullptr) return b200hist_ellpack_import_reference_rows(e, 0, 0, gidx_buffer_dev, row_stride, n_symbols, is_dense, stream);
  return b200hist_ellpack_import_reference_rows(e, 0, e->n_rows, gidx_buffer_dev, row_stride, n_symbols, is_dense,
                                                stream);
}

int b200hist_ellpack_info(b200hist_ellpack_t e, uint64_t* n_rows, uint32_t* n_features, uint32_t* total_bins,
                          uint32_t* row_bytes, uint32_t* bin_bytes, uint32_t* null_symbol, const void** bins_dev) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr, "null ellpack");
  if (n_rows) *n_rows = e->n_rows;
  if (n_features) *n_features = e->n_features;
  if (total_bins) *total_bins = e->total_bins;
  if (row_bytes) *row_bytes = e->row_bytes;
  if (bin_bytes) *bin_bytes = e->bin_bytes;
  if (null_symbol) *null_symbol = e->null_symbol;
  if (bins_dev) *bins_dev = e->d_bins;
  B200_API_END();
}

int b200hist_ellpack_copy_bins_to_host(b200hist_ellpack_t e, void* out_host) {
  B200_API_BEGIN();
  B200_CHECK(e != nullptr && out_host != nullptr, "null argument");
  DeviceGuard guard(e->device);
  if (e->paged) {
    std::memcpy(out_host, e->h_bins, e->n_rows * (uint64_t)e->row_bytes);
    return 0;
  }
  B200_CUDA(cudaMemcpy(out_host, e->d_bins, e->n_rows * (uint64_t)e->row_bytes, cudaMemcpyDeviceToHost));
  B200_API_END();
}

}  // extern "C"
#endif  // B200_KERNELS_ONLY
