// Gradient-based row sampling (TrainParam::sampling_method = gradient_based), on the quantised pairs of one tree.
//
// Replaces (reference paths relative to /root/reference):
//   GradientBasedSampling::Sample, ReduceGrad, CalcThresholdIndex, CalculateThresholdIndex, SampleRateDelta,
//   PoissonSampling                                      src/tree/gpu_hist/sampler.cu:63-157,194-312
//   MvsGradOp, SamplingProbability, RescaleGrad          src/tree/hist/sampler.h:48-74
//   FloatQuantiser, ToFixedPointOp, ToFloatingPointOp    src/tree/gpu_hist/quantiser.cuh:22-47
//
// Steps (all on the caller's stream, nothing read back):
//   1. rag_i = sqrtf(float(g_i^2 + 0.1 h_i^2)) with (g, h) = the quantised pair as doubles
//   2. thr = sort(rag) ++ [FLT_MAX]                      -- cub::DeviceRadixSort (the reference: thrust::sort)
//   3. FloatQuantiser(sum(thr), n); csum = inclusive sum of int64(double(thr_k) * to_fixed)   -- cub::DeviceScan on
//      integers: exact, so the threshold does not depend on a scan order (the reference quantises for this reason)
//   4. first i with denom = float(sample_rows) - float(n - i - 1) > 0 and thr_i < u = float(csum_i * to_float) / denom
//      <= thr_{i+1}: mu = u (no such i: mu = thr[1], what min_element over all-FLT_MAX picks)
//   5. row i: pair (0,0) stays; p = rag_i / mu; p >= 1 keeps the pair; else RandomWeight(seed)(i) <= p keeps it scaled
//      by 1.0f / p in the quantised domain, otherwise the pair becomes (0,0)
//   6. the root sum and the two hessian flags of the quantiser record are recomputed from the sampled pairs.
// The sort and the scan are library primitives (CUB, part of the toolkit): any correct sort / integer scan gives the
// same arrays, and this is not the hot path of a tree (one pass over n floats; the histogram visits n * F symbols).
// STATUS: green on a B200 since round 2; pinned on the reference's own gpu_hist trees under sampling_method=gradient_based
// (tests/golden/ref_gpu_next_trees.json, generated on a B200 by tools/make_golden_gpu_next.py), which this path and the
// oracle (orc_gradient_based_sample) reproduce bit for bit.
#ifndef B200_KERNELS_ONLY
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

#include <cfloat>

#include "common.cuh"

namespace b200 {

constexpr int kSampThreads = 256;

struct GbsState {     // device-resident scalars of one sampling pass
  double total;       // sum of the regularised absolute gradients
  double to_fixed, to_float;
  unsigned long long first_hit;  // smallest i that satisfies the rate condition (ULLONG_MAX: none)
  float mu;
  unsigned int pad;
};

__device__ __forceinline__ float RandomWeightS(uint32_t seed, unsigned long long i) {
  // RandomWeight of quantise.cu (sampler.cu:36-47): minstd_rand, discard(i), float(x - 1) / 2^31
  const unsigned long long m = 2147483647ull;
  unsigned long long x = seed % m;
  if (x == 0) x = 1;
  unsigned long long z = i + 1, mult = 48271ull, acc = 1;
  while (z > 0) {
    if (z & 1) acc = (acc * mult) % m;
    z >>= 1;
    mult = (mult * mult) % m;
  }
  x = (acc * x) % m;
  float r = (float)(uint32_t)(x - 1);
  r /= (1.0f + (float)2147483645u);
  return r;
}

__global__ void __launch_bounds__(kSampThreads) RagKernel(const longlong2* __restrict__ gq, uint64_t n,
                                                           const QuantDev* __restrict__ q, float* __restrict__ rag) {
  const double tg = q->to_float_g, th = q->to_float_h;
  for (uint64_t i = blockIdx.x * (uint64_t)kSampThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kSampThreads) {
    const longlong2 v = gq[i];
    const double g = (double)v.x * tg, h = (double)v.y * th;
    const float lambda = 0.1f;                               // kDefaultMvsLambda
    rag[i] = sqrtf((float)(g * g + lambda * (h * h)));       // MvsGradOp, then cuda::std::sqrt(float)
  }
}

// sum of the sorted values in double with a fixed two-level tree (only its power-of-two ceiling is used)
__global__ void __launch_bounds__(kSampThreads) TotalKernel(const float* __restrict__ thr, uint64_t n,
                                                             double* __restrict__ partials) {
  __shared__ double red[kSampThreads / 32];
  double s = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)kSampThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kSampThreads)
    s += (double)thr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSampThreads / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void PrepareKernel(const double* __restrict__ partials, int n_partials, uint64_t n, GbsState* st,
                              float* __restrict__ thr) {
  double t = 0.0;
  for (int b = 0; b < n_partials; ++b) t += partials[b];
  // FloatQuantiser (quantiser.cuh:22-32) = CreateRoundingFactor<double>(total, n) / 2^62
  double delta = t / (1.0 - 2.0 * (double)n * 2.220446049250313e-16);
  int e;
  (void)frexp(delta, &e);
  const double rounding = ldexp(1.0, e);
  st->total = t;
  st->to_float = rounding / 4611686018427387904.0;
  st->to_fixed = 1.0 / st->to_float;
  st->first_hit = 0xFFFFFFFFFFFFFFFFull;
  st->mu = FLT_MAX;
  thr[n] = FLT_MAX;  // sentinel upper bound (sampler.cu:103-104)
}

// ToFixedPointOp (quantiser.cuh:34-40) on the sorted values; the scan then runs in place on the integers
__global__ void __launch_bounds__(kSampThreads) ToFixedKernel(const float* __restrict__ thr, uint64_t n,
                                                               const GbsState* __restrict__ st,
                                                               long long* __restrict__ out) {
  const double f = st->to_fixed;
  for (uint64_t i = blockIdx.x * (uint64_t)kSampThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kSampThreads)
    out[i] = (long long)((double)thr[i] * f);
}

// SampleRateDelta + min_element: the smallest index whose rate matches
__global__ void __launch_bounds__(kSampThreads) ThresholdKernel(const float* __restrict__ thr,
                                                                 const long long* __restrict__ csum, uint64_t n,
                                                                 uint64_t sample_rows, GbsState* st) {
  const double tf = st->to_float;
  unsigned long long best = 0xFFFFFFFFFFFFFFFFull;
  for (uint64_t i = blockIdx.x * (uint64_t)kSampThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kSampThreads) {
    const float lower = thr[i], upper = thr[i + 1];
    const float denom = (float)sample_rows - (float)(n - i - 1);
    if (denom <= 0.0f) continue;
    const float u = (float)((double)csum[i] * tf) / denom;
    if (u > lower && u <= upper && i < best) best = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best != 0xFFFFFFFFFFFFFFFFull) atomicMin(&st->first_hit, best);
}

__global__ void MuKernel(const float* __restrict__ thr, const long long* __restrict__ csum, uint64_t n,
                         uint64_t sample_rows, GbsState* st) {
  if (n == 0) return;
  const unsigned long long i = st->first_hit;
  if (i == 0xFFFFFFFFFFFFFFFFull) {
    st->mu = thr[n >= 1 ? 1 : 0];  // min_element over all-FLT_MAX returns index 0: threshold index 1
  } else {
    const float denom = (float)sample_rows - (float)(n - i - 1);
    st->mu = (float)((double)csum[i] * st->to_float) / denom;  // the value the reference stores into thr[i + 1]
  }
}

__global__ void __launch_bounds__(kSampThreads) PoissonKernel(longlong2* __restrict__ gq, uint64_t n,
                                                               const float* __restrict__ rag,
                                                               const GbsState* __restrict__ st, QuantDev* __restrict__ q,
                                                               uint32_t seed) {
  const double tg = q->to_float_g, th = q->to_float_h, fg = q->to_fixed_g, fh = q->to_fixed_h;
  float u = st->mu;
  if (fabsf(u) < 1e-6f) u = copysignf(1e-6f, u);  // SamplingProbability (hist/sampler.h:60-65), kRtEps
  long long sg = 0, sh = 0;
  unsigned int lo_or = 0u;
  bool h_zero = false;
  for (uint64_t i = blockIdx.x * (uint64_t)kSampThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kSampThreads) {
    longlong2 v = gq[i];
    if (!(v.x == 0 && v.y == 0)) {
      const float p = rag[i] / u;
      if (p < 1.0f) {
        if (RandomWeightS(seed, i) <= p) {
          const float inv = 1.0f / p;  // RescaleGrad: GradientPairPrecise * float
          v.x = (long long)((((double)v.x * tg) * inv) * fg);
          v.y = (long long)((((double)v.y * th) * inv) * fh);
        } else {
          v = make_longlong2(0, 0);
        }
        gq[i] = v;
      }
    }
    sg += v.x;
    sh += v.y;
    lo_or |= (unsigned int)(unsigned long long)v.y;
    h_zero |= (v.y == 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_down_sync(0xffffffffu, sg, o);
    sh += __shfl_down_sync(0xffffffffu, sh, o);
  }
  lo_or = __reduce_or_sync(0xffffffffu, lo_or);
  const bool any_zero = __any_sync(0xffffffffu, h_zero);
  if ((threadIdx.x & 31) == 0) {  // integer sums: any order gives the same root sum
    atomicAdd(reinterpret_cast<unsigned long long*>(&q->root_g), (unsigned long long)sg);
    atomicAdd(reinterpret_cast<unsigned long long*>(&q->root_h), (unsigned long long)sh);
    if (lo_or) atomicOr(&q->h_lo_or, lo_or);
    if (any_zero) atomicOr(&q->h_zero_any, 1u);
  }
}

__global__ void ResetRootKernel(QuantDev* q) {
  q->root_g = 0;
  q->root_h = 0;
  q->h_lo_or = 0u;
  q->h_zero_any = 0u;
}

#ifndef B200_KERNELS_ONLY
size_t GradientBasedScratchBytes(uint64_t n, int num_sms) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const float*)nullptr, (float*)nullptr, (int64_t)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const long long*)nullptr, (long long*)nullptr, (int64_t)n);
  const size_t tmp = std::max(sort_bytes, scan_bytes);
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  // rag[n], thr[n + 1], csum[n] (int64), partials, state, cub temp
  return align(sizeof(float) * n) + align(sizeof(float) * (n + 1)) + align(sizeof(long long) * n) +
         align(sizeof(double) * (size_t)num_sms * 8) + align(sizeof(GbsState)) + align(tmp) + 256;
}

void LaunchGradientBasedSample(b200hist_gpair64* gq, uint64_t n, QuantDev* q, float subsample, uint32_t seed,
                               void* scratch, size_t scratch_bytes, int num_sms, cudaStream_t s) {
  if (n == 0) return;
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  uint8_t* p = static_cast<uint8_t*>(scratch);
  float* rag = reinterpret_cast<float*>(p); p += align(sizeof(float) * n);
  float* thr = reinterpret_cast<float*>(p); p += align(sizeof(float) * (n + 1));
  long long* csum = reinterpret_cast<long long*>(p); p += align(sizeof(long long) * n);
  double* partials = reinterpret_cast<double*>(p); p += align(sizeof(double) * (size_t)num_sms * 8);
  GbsState* st = reinterpret_cast<GbsState*>(p); p += align(sizeof(GbsState));
  void* tmp = p;
  size_t tmp_bytes = scratch_bytes - (size_t)(p - static_cast<uint8_t*>(scratch));
  auto* g2 = reinterpret_cast<longlong2*>(gq);
  const int blocks = num_sms * 8;
  RagKernel<<<blocks, kSampThreads, 0, s>>>(g2, n, q, rag);
  B200_LAUNCHED();
  size_t need = tmp_bytes;
  B200_CUDA(cub::DeviceRadixSort::SortKeys(tmp, need, (const float*)rag, thr, (int64_t)n, 0, 32, s));
  TotalKernel<<<blocks, kSampThreads, 0, s>>>(thr, n, partials);
  B200_LAUNCHED();
  PrepareKernel<<<1, 1, 0, s>>>(partials, blocks, n, st, thr);
  B200_LAUNCHED();
  ToFixedKernel<<<blocks, kSampThreads, 0, s>>>(thr, n, st, csum);
  B200_LAUNCHED();
  need = tmp_bytes;
  B200_CUDA(cub::DeviceScan::InclusiveSum(tmp, need, (const long long*)csum, csum, (int64_t)n, s));
  const uint64_t sample_rows = (uint64_t)((float)n * subsample);  // bst_idx_t sample_rows = n_samples * subsample_
  ThresholdKernel<<<blocks, kSampThreads, 0, s>>>(thr, csum, n, sample_rows, st);
  B200_LAUNCHED();
  MuKernel<<<1, 1, 0, s>>>(thr, csum, n, sample_rows, st);
  B200_LAUNCHED();
  ResetRootKernel<<<1, 1, 0, s>>>(q);
  B200_LAUNCHED();
  PoissonKernel<<<blocks, kSampThreads, 0, s>>>(g2, n, rag, st, q, seed);
  B200_LAUNCHED();
}

#endif  // B200_KERNELS_ONLY

}  // namespace b200
