// Shared-memory accumulator layout probe for the histogram kernel (DESIGN.md "The histogram kernel").
//
// ncu on HistKernelPair (profiles/r01c_pair_kernels_ncu_summary.json) counts 1.27 data-pipe wavefronts per shared
// atomic instruction (61.8 M bank conflicts on 230.7 M ATOMS) although every lane of an instruction addresses a
// different bank by construction; the 32-feature kernel's layout (512-byte symbol stride) runs at 1.00-1.02.  This
// program times the kernel's own accumulation sequence (atom.shared.add + carry + 2 x red.shared.add per update,
// random symbols from registers, no global loads) for several placements of the same 12-byte accumulators:
//
//   0  32 features, 16-byte slots:   sym*512 + c*128 + bank*4                  (HistKernelBankedFast, h.lo skipped)
//   1  64 features, pair layout:     sym*768 + half*384 + c*128 + bank*4       (HistKernelPair today)
//   2  64 features, plane-major:     c*65536 + sym*256 + half*128 + bank*4
//   3  64 features, half-major:      half*98304 + sym*384 + c*128 + bank*4
//   4  64 features, component rows:  sym*768 + c*256 + half*128 + bank*4
//   5  layout 1 with red instead of atom for g.lo (no carry: NOT exact, only to price the returned value)
//
// Output: updates per clock per SM for each (clock64 on every SM, 148 CTAs x 1024 threads, 192 KiB dynamic smem).
// Build + run on the GPU box:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/smem_layout_probe.cu && /tmp/probe
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e__ = (x);                                                            \
    if (e__ != cudaSuccess) {                                                         \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

constexpr int kThreads = 1024;

__device__ __forceinline__ uint32_t AtomAdd(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void RedAdd(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t CarryHi(uint32_t old, uint32_t vlo, uint32_t vhi) {
  uint32_t hi;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(hi) : "r"(old), "r"(vlo), "r"(vhi));
  return hi;
}

template <int kLayout>
__device__ __forceinline__ uint32_t SlotAddr(uint32_t base, uint32_t sym, uint32_t half, uint32_t bank) {
  if (kLayout == 0) return base + sym * 512u + bank * 4u;
  if (kLayout == 1 || kLayout == 5) return base + sym * 768u + half * 384u + bank * 4u;
  if (kLayout == 2) return base + sym * 256u + half * 128u + bank * 4u;
  if (kLayout == 3) return base + half * 98304u + sym * 384u + bank * 4u;
  return base + sym * 768u + half * 128u + bank * 4u;
}
template <int kLayout>
__device__ __forceinline__ uint32_t CompOff(int c) {  // byte offset of component c (0 g.lo, 1 g.hi, 2 h.hi)
  if (kLayout == 0) return c == 0 ? 0u : (c == 1 ? 128u : 384u);
  if (kLayout == 2) return (uint32_t)c * 65536u;
  if (kLayout == 4) return (uint32_t)c * 256u;
  return (uint32_t)c * 128u;
}

template <int kLayout>
__global__ void __launch_bounds__(kThreads, 1) Probe(uint32_t iters, uint32_t words, unsigned long long* clocks,
                                                      uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t sh[];
  __shared__ unsigned long long clk_s;
  for (uint32_t i = threadIdx.x; i < words; i += kThreads) sh[i] = 0u;
  if (threadIdx.x == 0) clk_s = 0ull;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, q = lane >> 3, w = lane & 7u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t bank[4], shift[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    bank[k] = 4u * w + jj;
    shift[k] = 8u * jj;
  }
  constexpr int kHalves = kLayout == 0 ? 1 : 2;
  uint32_t x = 0x9E3779B9u * (threadIdx.x + 1u) + blockIdx.x;
  const uint32_t glo = 0xFFFFFFF0u, ghi = 3u, hhi = 1u;
  const long long t0 = clock64();
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {      // two rows per pass
#pragma unroll
      for (int half = 0; half < kHalves; ++half) {
        x = x * 1664525u + 1013904223u;
        const uint32_t word = x ^ (x >> 13);
        uint32_t addr[4], og[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) addr[k] = SlotAddr<kLayout>(base, (word >> shift[k]) & 0xFFu, (uint32_t)half, bank[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kLayout == 5) { RedAdd(addr[k] + CompOff<kLayout>(0), glo); og[k] = 0u; }
          else og[k] = AtomAdd(addr[k] + CompOff<kLayout>(0), glo);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          RedAdd(addr[k] + CompOff<kLayout>(1), CarryHi(og[k], glo, ghi));
          RedAdd(addr[k] + CompOff<kLayout>(2), hhi);
        }
      }
    }
  }
  const long long t1 = clock64();
  atomicMax(&clk_s, (unsigned long long)(t1 - t0));
  __syncthreads();
  if (threadIdx.x == 0) {
    clocks[blockIdx.x] = clk_s;
    sink[blockIdx.x] = sh[x % words];
  }
}

template <int kLayout>
int Run(int n_sms, const char* what) {
  const uint32_t words = kLayout == 0 ? 256u * 32u * 4u : 256u * 192u;   // 128 KiB / 192 KiB
  const size_t smem = (size_t)words * 4;
  CK(cudaFuncSetAttribute(Probe<kLayout>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* clocks;
  uint32_t* sink;
  CK(cudaMalloc(&clocks, sizeof(unsigned long long) * n_sms));
  CK(cudaMalloc(&sink, sizeof(uint32_t) * n_sms));
  const uint32_t iters = 10000;
  for (int rep = 0; rep < 2; ++rep) Probe<kLayout><<<n_sms, kThreads, smem>>>(iters, words, clocks, sink);
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> h(n_sms);
  CK(cudaMemcpy(h.data(), clocks, sizeof(unsigned long long) * n_sms, cudaMemcpyDeviceToHost));
  double sum = 0;
  for (auto c : h) sum += (double)c;
  const double updates = (double)kThreads * iters * 2.0 * (kLayout == 0 ? 1.0 : 2.0) * 4.0;
  std::printf("layout %d  %-52s %6.2f updates/clk/SM  (%.2f shared atomic instr/clk/SM)\n", kLayout, what,
              updates / (sum / n_sms), 3.0 * updates / 32.0 / (sum / n_sms));
  cudaFree(clocks);
  cudaFree(sink);
  return 0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  std::printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int n = p.multiProcessorCount;
  if (Run<0>(n, "32 features, sym*512 (BankedFast, h.lo skipped)")) return 1;
  if (Run<1>(n, "64 features, sym*768 + half*384 + c*128 (Pair today)")) return 1;
  if (Run<2>(n, "64 features, c*65536 + sym*256 + half*128 (planes)")) return 1;
  if (Run<3>(n, "64 features, half*98304 + sym*384 + c*128")) return 1;
  if (Run<4>(n, "64 features, sym*768 + c*256 + half*128")) return 1;
  if (Run<5>(n, "layout 1, g.lo by red (no returned value; inexact)")) return 1;
  return 0;
}
