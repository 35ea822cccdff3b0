// Gradient quantiser: float gradient pairs -> int64 fixed point with a power-of-two scale, so that
// every later sum is exact integer arithmetic and therefore order-independent.
//
// Replaces (reference paths relative to /root/reference):
//   MakeQuantiserForTarget / BuildQuantiserFromPair   src/tree/gpu_hist/quantiser.cu:55-85
//   CreateRoundingFactor                              src/common/deterministic.cuh:24-38
//   CalcQuantizedGpairs / ToFixedPoint                src/tree/gpu_hist/quantiser.cu:131-149, quantiser.cuh:65-69
//   root sum dh::Reduce                               src/tree/updater_gpu_hist.cu:661-664
//   RowPartitioner::Reset (thrust::sequence)          src/tree/gpu_hist/row_partitioner.cu:13-28
//
// Two streaming kernels, both HBM-bound:
//   QuantReduceKernel   reads 8 B/row  : clipped +/- sums in fp64 with a FIXED reduction tree
//                       (per-thread strided sum -> warp shuffle -> block -> last block folds the
//                       per-block partials in block order), then derives the scale on the device.
//   QuantConvertKernel  reads 8 B/row, writes 16 B + 4 B/row: fixed-point conversion, fused with the
//                       exact int64 root sum and the row-index initialisation.
// Algorithmic bytes: N*(8 + 8 + 16 + 4).
#include "common.cuh"

namespace b200 {

constexpr int kQThreads = 256;

__device__ __forceinline__ double WarpSumD(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long WarpSumLL(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double RoundingFactor(double max_abs, unsigned long long n) {
  // deterministic.cuh:24-38.  DBL_EPSILON = 2^-52.
  double delta = max_abs / (1.0 - 2.0 * (double)n * 2.220446049250313e-16);
  int e;
  (void)frexp(delta, &e);
  return ldexp(1.0, e);
}

__global__ void __launch_bounds__(kQThreads) QuantReduceKernel(const float2* __restrict__ gpair, uint64_t n,
                                                                double* __restrict__ partials,
                                                                unsigned int* __restrict__ counter,
                                                                QuantDev* __restrict__ q) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  // four loads in flight per thread (one 8-byte load per iteration left the kernel latency bound at ~1.1 TB/s);
  // the per-thread summation order stays fixed, so the result is still deterministic
  const uint64_t stride = (uint64_t)gridDim.x * kQThreads;
  uint64_t i = blockIdx.x * (uint64_t)kQThreads + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(gpair + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s0 += (double)(v[k].x > 0 ? v[k].x : 0.0f);
      s1 += (double)(v[k].y > 0 ? v[k].y : 0.0f);
      s2 += (double)(v[k].x < 0 ? fabsf(v[k].x) : 0.0f);
      s3 += (double)(v[k].y < 0 ? fabsf(v[k].y) : 0.0f);
    }
  }
  for (; i < n; i += stride) {
    float2 v = __ldg(gpair + i);
    // Clip (quantiser.cu:30-44): clip in float, accumulate in double
    s0 += (double)(v.x > 0 ? v.x : 0.0f);
    s1 += (double)(v.y > 0 ? v.y : 0.0f);
    s2 += (double)(v.x < 0 ? fabsf(v.x) : 0.0f);
    s3 += (double)(v.y < 0 ? fabsf(v.y) : 0.0f);
  }
  __shared__ double sh[4][kQThreads / 32];
  __shared__ bool is_last;
  s0 = WarpSumD(s0); s1 = WarpSumD(s1); s2 = WarpSumD(s2); s3 = WarpSumD(s3);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; sh[2][warp] = s2; sh[3][warp] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int w = 0; w < kQThreads / 32; ++w) t += sh[threadIdx.x][w];
    partials[4 * blockIdx.x + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double t[4] = {0, 0, 0, 0};
    for (unsigned int b = 0; b < gridDim.x; ++b)
      for (int k = 0; k < 4; ++k) t[k] += ((volatile double*)partials)[4 * b + k];
    double mg = t[0] > t[2] ? t[0] : t[2];
    double mh = t[1] > t[3] ? t[1] : t[3];
    double rg = RoundingFactor(mg, n), rh = RoundingFactor(mh, n);
    const double div = 4611686018427387904.0;  // 2^62: keep one bit for the sign (quantiser.cu:77-80)
    q->to_float_g = rg / div;
    q->to_float_h = rh / div;
    q->to_fixed_g = 1.0 / q->to_float_g;
    q->to_fixed_h = 1.0 / q->to_float_h;
    for (int k = 0; k < 4; ++k) q->sums[k] = t[k];
    q->root_g = 0;
    q->root_h = 0;
    q->h_lo_or = 0u;
    q->h_zero_any = 0u;
    *counter = 0;  // ready for the next tree
  }
}

// RandomWeight(seed)(i) of the reference's row sampler (src/tree/gpu_hist/sampler.cu:36-47): minstd_rand seeded
// with `seed`, discard(i), one uniform_real_distribution<float> draw = float(x - 1) / 2^31 (the toolkit's
// thrust/random/detail/uniform_real_distribution.inl: 1.0f + float(max - min) rounds to 2^31).
__device__ __forceinline__ float RandomWeight(uint32_t seed, unsigned long long i) {
  const unsigned long long m = 2147483647ull;
  unsigned long long x = seed % m;
  if (x == 0) x = 1;
  unsigned long long z = i + 1, mult = 48271ull, acc = 1;  // discard(i) + one step = multiply by a^(i+1)
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

__global__ void __launch_bounds__(kQThreads) QuantConvertKernel(const float2* __restrict__ gpair, uint64_t n,
                                                                 QuantDev* __restrict__ q,
                                                                 longlong2* __restrict__ gq,
                                                                 uint32_t* __restrict__ ridx, float subsample,
                                                                 uint32_t sample_seed) {
  const double fg = q->to_fixed_g, fh = q->to_fixed_h;
  long long sg = 0, sh_ = 0;
  unsigned int lo_or = 0u;
  bool h_zero = false;
  for (uint64_t i = blockIdx.x * (uint64_t)kQThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kQThreads) {
    float2 v = __ldg(gpair + i);
    // ToFixedPoint (quantiser.cuh:65-69): C cast of double -> int64 truncates toward zero
    long long g = (long long)((double)v.x * fg);
    long long h = (long long)((double)v.y * fh);
    // UniformSampling::Sample (sampler.cu:160-173): the quantised pair of a row that loses its Bernoulli trial
    // becomes zero (the quantiser scale above was derived from ALL rows, as in the reference)
    if (subsample < 1.0f && RandomWeight(sample_seed, i) > subsample) { g = 0; h = 0; }
    gq[i] = make_longlong2(g, h);
    if (ridx != nullptr) ridx[i] = (uint32_t)i;
    sg += g;
    sh_ += h;
    lo_or |= (unsigned int)(unsigned long long)h;
    h_zero |= (h == 0);
  }
  __shared__ long long red[2][kQThreads / 32];
  sg = WarpSumLL(sg);
  sh_ = WarpSumLL(sh_);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  lo_or = __reduce_or_sync(0xffffffffu, lo_or);
  if (lane == 0) {
    red[0][warp] = sg;
    red[1][warp] = sh_;
    if (lo_or) atomicOr(&q->h_lo_or, lo_or);
  }
  if (__any_sync(0xffffffffu, h_zero) && lane == 0) {
    atomicOr(&q->h_zero_any, 1u);
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    long long t = 0;
    for (int w = 0; w < kQThreads / 32; ++w) t += red[threadIdx.x][w];
    // integer addition is associative: any order gives the same root sum
    atomicAdd(reinterpret_cast<unsigned long long*>(threadIdx.x == 0 ? &q->root_g : &q->root_h),
              (unsigned long long)t);
  }
}

void LaunchQuantReduce(const float* gpair, uint64_t n, QuantDev* q, double* partials, unsigned int* counter,
                       int num_sms, cudaStream_t s) {
  QuantReduceKernel<<<num_sms * kQuantBlocksPerSM, kQThreads, 0, s>>>(reinterpret_cast<const float2*>(gpair), n,
                                                                      partials, counter, q);
  B200_LAUNCHED();
}

void LaunchQuantConvert(const float* gpair, uint64_t n, QuantDev* q, b200hist_gpair64* gq, uint32_t* ridx_init,
                        int num_sms, cudaStream_t s, float subsample, uint32_t sample_seed) {
  QuantConvertKernel<<<num_sms * kQuantBlocksPerSM, kQThreads, 0, s>>>(reinterpret_cast<const float2*>(gpair), n, q,
                                                                       reinterpret_cast<longlong2*>(gq), ridx_init,
                                                                       subsample, sample_seed);
  B200_LAUNCHED();
}

void LaunchQuantise(const float* gpair, uint64_t n, QuantDev* q, b200hist_gpair64* gq, uint32_t* ridx_init,
                    double* partials, unsigned int* counter, int num_sms, cudaStream_t s, float subsample,
                    uint32_t sample_seed) {
  LaunchQuantReduce(gpair, n, q, partials, counter, num_sms, s);
  LaunchQuantConvert(gpair, n, q, gq, ridx_init, num_sms, s, subsample, sample_seed);
}

}  // namespace b200

using namespace b200;

extern "C" int b200hist_quantise(int device, const float* gpair_dev, uint64_t n, b200hist_quantiser* out_host,
                                 b200hist_gpair64* gq_dev, b200hist_gpair64* root_sum_host, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(gpair_dev != nullptr && gq_dev != nullptr, "null argument");
  DeviceGuard guard(device);
  auto s = static_cast<cudaStream_t>(stream);
  int sms = NumSMs(device);
  QuantDev* q = DevAlloc<QuantDev>(1);
  double* partials = DevAlloc<double>(4 * sms * kQuantBlocksPerSM);
  unsigned int* counter = DevAlloc<unsigned int>(1);
  B200_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
  LaunchQuantise(gpair_dev, n, q, gq_dev, nullptr, partials, counter, sms, s);
  QuantDev h;
  B200_CUDA(cudaMemcpyAsync(&h, q, sizeof(h), cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  cudaFree(q);
  cudaFree(partials);
  cudaFree(counter);
  if (out_host) *out_host = b200hist_quantiser{h.to_fixed_g, h.to_fixed_h, h.to_float_g, h.to_float_h};
  if (root_sum_host) *root_sum_host = b200hist_gpair64{h.root_g, h.root_h};
  B200_API_END();
}
