// What the global loads cost the histogram kernel's shared-atomic pipe (DESIGN.md "The histogram kernel").
//
// tools/smem_layout_probe.cu: the accumulation sequence alone issues ONE shared atomic instruction per clock per SM
// (10.65 updates/clk/SM) whatever the accumulator layout and whether or not g.lo's old value is returned.  The real
// kernel reaches 6.0.  This program re-creates HistKernelPair's inner loop -- 8 rows per warp pass, lane (q, w) reads
// 8 symbol bytes of rows q and q+4, the rows' quantised gradient pairs and the row ids, two passes prefetched in
// registers -- over a matrix of `rows` x 256 bytes and times variants of the LOADS, everything else unchanged:
//
//   A  no loads (symbols from a register LCG)                      -> the pipe's ceiling with this loop structure
//   B  row ids + symbols + gradient pairs as today (LDG.32 / LDG.64 / LDG.128 by every lane)
//   C  like B, gradient pair as LDG.64 (g only: a constant hessian needs no load)
//   D  like B, gradient pair loaded by ONE lane per row and shuffled to the other seven
//   E  like B without the gradient-pair load at all (constant g: prices the pair load)
//   F  like B without the symbol load (symbols from the LCG, pair + row id loaded)
// each for sequential row ids (the root) and for a random permutation's prefix (deep levels).
// Output: updates per clock per SM.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/p tools/hist_pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e__ = (x);                                                                      \
    if (e__ != cudaSuccess) {                                                                   \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);     \
      std::exit(1);                                                                             \
    }                                                                                           \
  } while (0)

constexpr int kThreads = 1024;
constexpr uint32_t kRowBytes = 256;

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

struct Pass {
  uint32_t w[2][2];  // [row][half]
  uint32_t glo[2], ghi[2], hhi[2];
};

enum Mode { kNoLoads = 0, kAsToday = 1, kPair64 = 2, kPairShfl = 3, kNoPair = 4, kNoSym = 5 };

template <int kMode>
__device__ __forceinline__ void Load(Pass& d, const uint32_t (&r)[2], const longlong2* __restrict__ gq,
                                     const uint8_t* __restrict__ col, uint32_t& x, uint32_t w) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (kMode == kNoLoads || kMode == kNoSym) {
      x = x * 1664525u + 1013904223u;
      d.w[u][0] = x ^ (x >> 13);
      d.w[u][1] = x ^ (x >> 7);
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(col + (size_t)r[u] * kRowBytes));
      d.w[u][0] = v.x;
      d.w[u][1] = v.y;
    }
    if (kMode == kNoLoads || kMode == kNoPair) {
      d.glo[u] = 0xFFFFFFF0u; d.ghi[u] = 3u; d.hhi[u] = 1u;
    } else if (kMode == kPair64) {
      const long long g = __ldg(reinterpret_cast<const long long*>(gq + r[u]));
      d.glo[u] = (uint32_t)(unsigned long long)g; d.ghi[u] = (uint32_t)((unsigned long long)g >> 32); d.hhi[u] = 1u;
    } else if (kMode == kPairShfl) {
      longlong2 gp = make_longlong2(0, 0);
      if (w == 0) gp = __ldg(gq + r[u]);
      const uint32_t src = (threadIdx.x & 31u) & ~7u;  // lane (q, 0)
      d.glo[u] = __shfl_sync(0xffffffffu, (uint32_t)(unsigned long long)gp.x, src);
      d.ghi[u] = __shfl_sync(0xffffffffu, (uint32_t)((unsigned long long)gp.x >> 32), src);
      d.hhi[u] = __shfl_sync(0xffffffffu, (uint32_t)((unsigned long long)gp.y >> 32), src);
    } else {
      const longlong2 gp = __ldg(gq + r[u]);
      d.glo[u] = (uint32_t)(unsigned long long)gp.x; d.ghi[u] = (uint32_t)((unsigned long long)gp.x >> 32);
      d.hhi[u] = (uint32_t)((unsigned long long)gp.y >> 32);
    }
  }
}

__device__ __forceinline__ void Process(const Pass& d, uint32_t base, const uint32_t (&shift)[4], const uint32_t (&off)[4]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t addr[4], og[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) addr[k] = base + ((d.w[u][half] >> shift[k]) & 0xFFu) * 768u + half * 384u + off[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) og[k] = AtomAdd(addr[k], d.glo[u]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        RedAdd(addr[k] + 128u, CarryHi(og[k], d.glo[u], d.ghi[u]));
        RedAdd(addr[k] + 256u, d.hhi[u]);
      }
    }
  }
}

template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
Probe(const uint8_t* __restrict__ bins, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx,
      uint32_t n_list, uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t sh[];
  for (uint32_t i = threadIdx.x; i < 256u * 192u; i += kThreads) sh[i] = 0u;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, q = lane >> 3, w = lane & 7u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t shift[4], off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    shift[k] = 8u * jj;
    off[k] = (4u * w + jj) * 4u;
  }
  // this CTA's contiguous share of the row list; a warp pass = 8 rows, passes of the 32 warps interleave
  const uint32_t per = n_list / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = lo + per;
  const uint8_t* __restrict__ col = bins + 8u * w;
  constexpr uint32_t kStride = 32u * 8u;
  const uint32_t n_pass = per / kStride;
  const uint32_t* __restrict__ rp = ridx + lo + warp * 8u + q;
  uint32_t x = 0x9E3779B9u * (threadIdx.x + 1u) + blockIdx.x;
  auto rid = [&](uint32_t i, uint32_t (&r)[2]) {
    const uint32_t* p = rp + (size_t)(i < n_pass ? i : n_pass - 1) * kStride;
    if (kMode == kNoLoads) { r[0] = i; r[1] = i + 4; }
    else { r[0] = __ldg(p); r[1] = __ldg(p + 4); }
  };
  uint32_t r0[2], r1[2], r2[2];
  Pass d0, d1, d2;
  rid(0, r0); rid(1, r1); rid(2, r2);
  Load<kMode>(d0, r0, gq, col, x, w); rid(3, r0);
  Load<kMode>(d1, r1, gq, col, x, w); rid(4, r1);
  uint32_t i = 0;
  for (; i + 2 < n_pass; i += 3) {
    Load<kMode>(d2, r2, gq, col, x, w); rid(i + 5, r2);
    Process(d0, base, shift, off);
    Load<kMode>(d0, r0, gq, col, x, w); rid(i + 6, r0);
    Process(d1, base, shift, off);
    Load<kMode>(d1, r1, gq, col, x, w); rid(i + 7, r1);
    Process(d2, base, shift, off);
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sh[(x + hi) % (256u * 192u)];
}

// ---------------------------------------------------------------------------------------------------------------
// Variant G: FOUR lanes per row.  Lane (q', w') = (lane >> 2, lane & 3) reads 16 symbol bytes of row q' with one
// LDG.128, so one warp instruction covers 8 rows x 64 features (today: 4 rows), and the pair / row-id loads serve 8
// rows each as well: 3 LDG per 512 updates instead of 6.  Bank freedom: the lane's four words are rotated by q' & 3
// after the load (two conditional swaps), its bytes by (q' >> 2) + 2 * (w' >> 1) through the PRMT selector, so that in
// step (j, k) the 32 lanes address 32 different (feature mod 32) = banks.
// kRowsPerLane = 1: 8 rows per warp pass; 2: 16 rows (two rows in flight per lane and buffer).
// ---------------------------------------------------------------------------------------------------------------
struct PassG {
  uint32_t w[4];
  uint32_t glo, ghi, hhi;
};
__device__ __forceinline__ void LoadG(PassG& d, uint32_t r, const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                      uint32_t rw) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(col + (size_t)r * kRowBytes));
  const longlong2 gp = __ldg(gq + r);
  // rotate the words by rw: W'[j] = W[(j + rw) & 3]
  uint32_t a0 = v.x, a1 = v.y, a2 = v.z, a3 = v.w;
  if (rw & 1u) { const uint32_t t = a0; a0 = a1; a1 = a2; a2 = a3; a3 = t; }
  if (rw & 2u) { uint32_t t = a0; a0 = a2; a2 = t; t = a1; a1 = a3; a3 = t; }
  d.w[0] = a0; d.w[1] = a1; d.w[2] = a2; d.w[3] = a3;
  d.glo = (uint32_t)(unsigned long long)gp.x; d.ghi = (uint32_t)((unsigned long long)gp.x >> 32);
  d.hhi = (uint32_t)((unsigned long long)gp.y >> 32);
}
__device__ __forceinline__ void ProcessG(const PassG& d, const uint32_t (&shift)[4], const uint32_t (&off_j)[4],
                                         const uint32_t (&off_k)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t addr[4], og[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) addr[k] = ((d.w[j] >> shift[k]) & 0xFFu) * 768u + off_j[j] + off_k[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) og[k] = AtomAdd(addr[k], d.glo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      RedAdd(addr[k] + 128u, CarryHi(og[k], d.glo, d.ghi));
      RedAdd(addr[k] + 256u, d.hhi);
    }
  }
}

template <int kBufs>
__global__ void __launch_bounds__(kThreads, 1)
ProbeG(const uint8_t* __restrict__ bins, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx,
       uint32_t n_list, uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t sh[];
  for (uint32_t i = threadIdx.x; i < 256u * 192u; i += kThreads) sh[i] = 0u;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, qp = lane >> 2, wp = lane & 3u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
  const uint32_t rw = qp & 3u, rb = (qp >> 2) + 2u * (wp >> 1);
  uint32_t shift[4], off_j[4], off_k[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    shift[k] = 8u * ((k + rb) & 3u);
    off_k[k] = 4u * ((k + rb) & 3u);
    off_j[k] = base + (wp >> 1) * 384u + 64u * (wp & 1u) + 16u * ((k + rw) & 3u);
  }
  const uint32_t per = n_list / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = lo + per;
  const uint8_t* __restrict__ col = bins + 16u * wp;
  constexpr uint32_t kStride = 32u * 8u;
  const uint32_t n_pass = per / kStride;
  const uint32_t* __restrict__ rp = ridx + lo + warp * 8u + qp;
  auto rid = [&](uint32_t i) { return __ldg(rp + (size_t)(i < n_pass ? i : n_pass - 1) * kStride); };
  if (kBufs == 4) {
    uint32_t r0 = rid(0), r1 = rid(1), r2 = rid(2), r3 = rid(3);
    PassG d0, d1, d2, d3;
    LoadG(d0, r0, gq, col, rw); r0 = rid(4);
    LoadG(d1, r1, gq, col, rw); r1 = rid(5);
    LoadG(d2, r2, gq, col, rw); r2 = rid(6);
    for (uint32_t i = 0; i + 3 < n_pass; i += 4) {
      LoadG(d3, r3, gq, col, rw); r3 = rid(i + 7);
      ProcessG(d0, shift, off_j, off_k);
      LoadG(d0, r0, gq, col, rw); r0 = rid(i + 8);
      ProcessG(d1, shift, off_j, off_k);
      LoadG(d1, r1, gq, col, rw); r1 = rid(i + 9);
      ProcessG(d2, shift, off_j, off_k);
      LoadG(d2, r2, gq, col, rw); r2 = rid(i + 10);
      ProcessG(d3, shift, off_j, off_k);
    }
  } else {
    uint32_t r0 = rid(0), r1 = rid(1), r2 = rid(2);
    PassG d0, d1, d2;
    LoadG(d0, r0, gq, col, rw); r0 = rid(3);
    LoadG(d1, r1, gq, col, rw); r1 = rid(4);
    for (uint32_t i = 0; i + 2 < n_pass; i += 3) {
      LoadG(d2, r2, gq, col, rw); r2 = rid(i + 5);
      ProcessG(d0, shift, off_j, off_k);
      LoadG(d0, r0, gq, col, rw); r0 = rid(i + 6);
      ProcessG(d1, shift, off_j, off_k);
      LoadG(d1, r1, gq, col, rw); r1 = rid(i + 7);
      ProcessG(d2, shift, off_j, off_k);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sh[hi % (256u * 192u)];
}

template <int kBufs>
double RunG(const uint8_t* bins, const longlong2* gq, const uint32_t* ridx, uint32_t n_list, int n_sms, uint32_t* sink) {
  const size_t smem = 256u * 192u * 4u;
  CK(cudaFuncSetAttribute(ProbeG<kBufs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  ProbeG<kBufs><<<n_sms, kThreads, smem>>>(bins, gq, ridx, n_list, sink);
  CK(cudaEventRecord(e0));
  for (int rep = 0; rep < 3; ++rep) ProbeG<kBufs><<<n_sms, kThreads, smem>>>(bins, gq, ridx, n_list, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 3;
  const uint32_t per = n_list / n_sms;
  const double rows = (double)(per / 256u / kBufs * kBufs) * 256.0 * n_sms;
  return rows * 64.0 / (ms * 1e-3) / 1.965e9 / n_sms;
}

// ---------------------------------------------------------------------------------------------------------------
// Variant H: the gradient pair (int4 texel) and the row id (uint texel) come through the TEXTURE unit
// (tex1Dfetch on linear memory), the symbols through LDG.64 as today: does the TEX data path spare the LSU data stage?
// ---------------------------------------------------------------------------------------------------------------
template <bool kRidTex>
__global__ void __launch_bounds__(kThreads, 1)
ProbeH(const uint8_t* __restrict__ bins, cudaTextureObject_t gq_tex, cudaTextureObject_t rid_tex,
       const uint32_t* __restrict__ ridx, uint32_t n_list, uint32_t* sink) {
  extern __shared__ __align__(16) uint32_t sh[];
  for (uint32_t i = threadIdx.x; i < 256u * 192u; i += kThreads) sh[i] = 0u;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, q = lane >> 3, w = lane & 7u;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t shift[4], off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    shift[k] = 8u * jj;
    off[k] = (4u * w + jj) * 4u;
  }
  const uint32_t per = n_list / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = lo + per;
  const uint8_t* __restrict__ col = bins + 8u * w;
  constexpr uint32_t kStride = 32u * 8u;
  const uint32_t n_pass = per / kStride;
  const uint32_t rbase = lo + warp * 8u + q;
  auto rid = [&](uint32_t i, uint32_t (&r)[2]) {
    const uint32_t at = rbase + (i < n_pass ? i : n_pass - 1) * kStride;
    if (kRidTex) { r[0] = tex1Dfetch<uint32_t>(rid_tex, (int)at); r[1] = tex1Dfetch<uint32_t>(rid_tex, (int)at + 4); }
    else { r[0] = __ldg(ridx + at); r[1] = __ldg(ridx + at + 4); }
  };
  auto load = [&](Pass& d, const uint32_t (&r)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(col + (size_t)r[u] * kRowBytes));
      d.w[u][0] = v.x; d.w[u][1] = v.y;
      const int4 gp = tex1Dfetch<int4>(gq_tex, (int)r[u]);
      d.glo[u] = (uint32_t)gp.x; d.ghi[u] = (uint32_t)gp.y; d.hhi[u] = (uint32_t)gp.w;
    }
  };
  uint32_t r0[2], r1[2], r2[2];
  Pass d0, d1, d2;
  rid(0, r0); rid(1, r1); rid(2, r2);
  load(d0, r0); rid(3, r0);
  load(d1, r1); rid(4, r1);
  for (uint32_t i = 0; i + 2 < n_pass; i += 3) {
    load(d2, r2); rid(i + 5, r2);
    Process(d0, base, shift, off);
    load(d0, r0); rid(i + 6, r0);
    Process(d1, base, shift, off);
    load(d1, r1); rid(i + 7, r1);
    Process(d2, base, shift, off);
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sh[hi % (256u * 192u)];
}

template <bool kRidTex>
double RunH(const uint8_t* bins, const longlong2* gq, const uint32_t* ridx, uint32_t n_rows, uint32_t n_list, int n_sms,
            uint32_t* sink) {
  const size_t smem = 256u * 192u * 4u;
  CK(cudaFuncSetAttribute(ProbeH<kRidTex>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto make_tex = [&](const void* ptr, size_t bytes, cudaChannelFormatDesc desc) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = const_cast<void*>(ptr);
    rd.res.linear.desc = desc;
    rd.res.linear.sizeInBytes = bytes;
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    return t;
  };
  cudaTextureObject_t gq_tex = make_tex(gq, (size_t)n_rows * 16, cudaCreateChannelDesc<int4>());
  cudaTextureObject_t rid_tex = make_tex(ridx, (size_t)n_list * 4, cudaCreateChannelDesc<uint32_t>());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  ProbeH<kRidTex><<<n_sms, kThreads, smem>>>(bins, gq_tex, rid_tex, ridx, n_list, sink);
  CK(cudaEventRecord(e0));
  for (int rep = 0; rep < 3; ++rep) ProbeH<kRidTex><<<n_sms, kThreads, smem>>>(bins, gq_tex, rid_tex, ridx, n_list, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 3;
  CK(cudaDestroyTextureObject(gq_tex));
  CK(cudaDestroyTextureObject(rid_tex));
  const uint32_t per = n_list / n_sms;
  const double rows = (double)(per / 256u / 3u * 3u) * 256.0 * n_sms;
  return rows * 64.0 / (ms * 1e-3) / 1.965e9 / n_sms;
}

// ---------------------------------------------------------------------------------------------------------------
// Variant I: the SYMBOLS arrive by TMA (cp.async.bulk.tensor ... tile::gather4, SASS UTMALDG.2D.GATHER4: four arbitrary
// rows x 64 bytes per instruction, issued by lane 0 of each warp into a per-warp ring of kStagesI stages of 4 rows,
// completion on an mbarrier per stage); the consumers read their 8 bytes with LDS.64.  Gradient pairs and row ids stay
// on LDG as today.  Shared memory: 192 KiB histogram + 32 warps x 4 stages x 256 B + barriers = 225 KiB.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStagesI = 4;
__device__ __forceinline__ bool BarDone(uint32_t bar_a, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar_a), "r"(parity)
      : "memory");
  return done != 0;
}

__global__ void __launch_bounds__(kThreads, 1)
ProbeI(const __grid_constant__ CUtensorMap tm, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx,
       uint32_t n_list, uint32_t* sink, int* stuck) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint32_t* sh = reinterpret_cast<uint32_t*>(smem_raw);
  constexpr uint32_t kHistBytes = 256u * 768u;
  constexpr uint32_t kSlot = 256u;  // 4 rows x 64 bytes
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, q = lane >> 3, w = lane & 7u;
  uint8_t* ring = smem_raw + kHistBytes + warp * (kStagesI * kSlot);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + kHistBytes + 32u * kStagesI * kSlot) + warp * kStagesI;
  for (uint32_t i = threadIdx.x; i < 256u * 192u; i += kThreads) sh[i] = 0u;
  if (lane == 0) {
    for (int s = 0; s < kStagesI; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t shift[4], off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    shift[k] = 8u * jj;
    off[k] = (4u * w + jj) * 4u;
  }
  const uint32_t per = (n_list / gridDim.x) & ~255u;   // 16-byte aligned row-id groups
  const uint32_t lo = blockIdx.x * per;
  constexpr uint32_t kStride = 32u * 8u;
  const uint32_t n_pass = per / kStride;
  // group g of this warp = 4 consecutive entries of its row list: pass i holds groups 2i (rows q) and 2i + 1 (rows q + 4)
  const uint32_t* __restrict__ wl = ridx + lo + warp * 8u;
  auto issue = [&](uint32_t g) {  // lane 0: group g -> stage g % kStagesI
    const uint32_t st = g % kStagesI;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bars[st]);
    const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(ring + st * kSlot);
    const uint4 r = *reinterpret_cast<const uint4*>(wl + (size_t)(g >> 1) * kStride + (g & 1u) * 4u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(256));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst_a), "l"(&tm), "r"(0), "r"((int)r.x), "r"((int)r.y), "r"((int)r.z), "r"((int)r.w), "r"(bar_a)
        : "memory");
  };
  const uint32_t n_groups = 2u * n_pass;
  if (lane == 0)
    for (uint32_t g = 0; g < (uint32_t)kStagesI && g < n_groups; ++g) issue(g);
  // gradient pairs + row ids: two passes ahead in registers, as today
  auto pair_of = [&](uint32_t i, uint32_t u, uint32_t& glo, uint32_t& ghi, uint32_t& hhi) {
    const uint32_t r = __ldg(wl + (size_t)(i < n_pass ? i : n_pass - 1) * kStride + u * 4u + q);
    const longlong2 gp = __ldg(gq + r);
    glo = (uint32_t)(unsigned long long)gp.x; ghi = (uint32_t)((unsigned long long)gp.x >> 32);
    hhi = (uint32_t)((unsigned long long)gp.y >> 32);
  };
  uint32_t parity = 0;
  auto consume = [&](uint32_t i, const uint32_t (&pgv)[2][3]) {
#pragma unroll
    for (uint32_t u = 0; u < 2; ++u) {
      const uint32_t g = 2u * i + u, st = g % kStagesI;
      const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bars[st]);
      int spin = 0;
      while (!BarDone(bar_a, parity)) {
        if (++spin > (1 << 22)) { if (lane == 0) atomicAdd(stuck, 1); return false; }
      }
      const uint2 v = *reinterpret_cast<const uint2*>(ring + st * kSlot + q * 64u + w * 8u);
      __syncwarp();
      if (lane == 0 && g + kStagesI < n_groups) issue(g + kStagesI);
      const uint32_t words[2] = {v.x, v.y};
      const uint32_t glo = pgv[u][0], ghi = pgv[u][1], hhi = pgv[u][2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t addr[4], og[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) addr[k] = base + ((words[half] >> shift[k]) & 0xFFu) * 768u + half * 384u + off[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) og[k] = AtomAdd(addr[k], glo);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          RedAdd(addr[k] + 128u, CarryHi(og[k], glo, ghi));
          RedAdd(addr[k] + 256u, hhi);
        }
      }
      if (st == kStagesI - 1) parity ^= 1u;
    }
    return true;
  };
  uint32_t pa[2][3], pb[2][3];   // gradient pairs of two passes in registers
  pair_of(0, 0, pa[0][0], pa[0][1], pa[0][2]);
  pair_of(0, 1, pa[1][0], pa[1][1], pa[1][2]);
  for (uint32_t i = 0; i + 1 < n_pass; i += 2) {
    pair_of(i + 1, 0, pb[0][0], pb[0][1], pb[0][2]);
    pair_of(i + 1, 1, pb[1][0], pb[1][1], pb[1][2]);
    if (!consume(i, pa)) return;
    pair_of(i + 2, 0, pa[0][0], pa[0][1], pa[0][2]);
    pair_of(i + 2, 1, pa[1][0], pa[1][1], pa[1][2]);
    if (!consume(i + 1, pb)) return;
  }
  __syncthreads();
  if (threadIdx.x == 0) sink[blockIdx.x] = sh[(lo + n_pass) % (256u * 192u)];
}

double RunI(const CUtensorMap& tm, const longlong2* gq, const uint32_t* ridx, uint32_t n_list, int n_sms, uint32_t* sink) {
  const size_t smem = 256u * 768u + 32u * kStagesI * 256u + 32u * kStagesI * 8u;
  CK(cudaFuncSetAttribute(ProbeI, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int* stuck;
  CK(cudaMalloc(&stuck, 4));
  CK(cudaMemset(stuck, 0, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  ProbeI<<<n_sms, kThreads, smem>>>(tm, gq, ridx, n_list, sink, stuck);
  CK(cudaEventRecord(e0));
  for (int rep = 0; rep < 3; ++rep) ProbeI<<<n_sms, kThreads, smem>>>(tm, gq, ridx, n_list, sink, stuck);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 3;
  int h_stuck = 0;
  CK(cudaMemcpy(&h_stuck, stuck, 4, cudaMemcpyDeviceToHost));
  if (h_stuck) std::printf("  [variant I: %d warps gave up waiting]\n", h_stuck);
  cudaFree(stuck);
  const uint32_t per = (n_list / n_sms) & ~255u;
  const double rows = (double)(per / 256u / 2u * 2u) * 256.0 * n_sms;
  return rows * 64.0 / (ms * 1e-3) / 1.965e9 / n_sms;
}

template <int kMode>
double Run(const uint8_t* bins, const longlong2* gq, const uint32_t* ridx, uint32_t n_list, int n_sms, uint32_t* sink) {
  const size_t smem = 256u * 192u * 4u;
  CK(cudaFuncSetAttribute(Probe<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  Probe<kMode><<<n_sms, kThreads, smem>>>(bins, gq, ridx, n_list, sink);
  CK(cudaEventRecord(e0));
  for (int rep = 0; rep < 3; ++rep) Probe<kMode><<<n_sms, kThreads, smem>>>(bins, gq, ridx, n_list, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 3;
  const uint32_t per = n_list / n_sms;
  const double rows = (double)(per / 256u / 3u * 3u) * 256.0 * n_sms;  // rows really processed (whole triples of passes)
  return rows * 64.0 / (ms * 1e-3) / 1.965e9 / n_sms;
}

int main(int argc, char** argv) {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int n_sms = p.multiProcessorCount;
  const uint32_t rows = argc > 1 ? (uint32_t)std::atoll(argv[1]) : 40000000u;
  std::printf("%s, %d SMs, matrix %u x %u B, 1.965 GHz assumed\n", p.name, n_sms, rows, kRowBytes);
  uint8_t* bins;
  longlong2* gq;
  uint32_t *seq, *perm, *sink;
  CK(cudaMalloc(&bins, (size_t)rows * kRowBytes));
  CK(cudaMalloc(&gq, (size_t)rows * sizeof(longlong2)));
  CK(cudaMalloc(&seq, (size_t)rows * 4));
  CK(cudaMalloc(&perm, (size_t)rows * 4));
  CK(cudaMalloc(&sink, 4096));
  {
    std::vector<uint8_t> h((size_t)1 << 26);
    uint32_t x = 12345u;
    for (auto& b : h) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
    for (size_t o = 0; o < (size_t)rows * kRowBytes; o += h.size())
      CK(cudaMemcpy(bins + o, h.data(), std::min(h.size(), (size_t)rows * kRowBytes - o), cudaMemcpyHostToDevice));
    CK(cudaMemset(gq, 0x11, (size_t)rows * sizeof(longlong2)));
    std::vector<uint32_t> s(rows), pm;
    for (uint32_t i = 0; i < rows; ++i) s[i] = i;
    CK(cudaMemcpy(seq, s.data(), (size_t)rows * 4, cudaMemcpyHostToDevice));
    // a deep level: every 8th row on average, ascending (what a node's segment looks like)
    pm.reserve(rows / 8 + 16);
    x = 777u;
    for (uint32_t i = 0; i < rows; ++i) { x = x * 1664525u + 1013904223u; if ((x >> 29) == 0u) pm.push_back(i); }
    while (pm.size() < rows / 8) pm.push_back(rows - 1);
    CK(cudaMemcpy(perm, pm.data(), (size_t)(rows / 8) * 4, cudaMemcpyHostToDevice));
  }
  const uint32_t n_seq = rows, n_sparse = rows / 8;
  std::printf("%-66s %8s %8s\n", "variant (updates/clk/SM)", "root", "1-in-8");
#define ROW(MODE, WHAT)                                                                                       \
  std::printf("%-66s %8.2f %8.2f\n", WHAT, Run<MODE>(bins, gq, seq, n_seq, n_sms, sink),                      \
              Run<MODE>(bins, gq, perm, n_sparse, n_sms, sink));
  ROW(kNoLoads, "A  no loads (register symbols)")
  ROW(kAsToday, "B  as today: LDG.32 row id, LDG.64 symbols, LDG.128 pair")
  ROW(kPair64, "C  pair as LDG.64 (g only, constant hessian)")
  ROW(kPairShfl, "D  pair by one lane per row + 3 SHFL")
  ROW(kNoPair, "E  no pair load (constant g)")
  ROW(kNoSym, "F  no symbol load (register symbols)")
  std::printf("%-66s %8.2f %8.2f\n", "G3 four lanes per row: LDG.128 symbols, 8 rows per LDG, 3 buffers",
              RunG<3>(bins, gq, seq, n_seq, n_sms, sink), RunG<3>(bins, gq, perm, n_sparse, n_sms, sink));
  std::printf("%-66s %8.2f %8.2f\n", "H1 pair through the texture unit (tex1Dfetch<int4>), rest as today",
              RunH<false>(bins, gq, seq, rows, n_seq, n_sms, sink), RunH<false>(bins, gq, perm, rows, n_sparse, n_sms, sink));
  std::printf("%-66s %8.2f %8.2f\n", "H2 pair AND row ids through the texture unit",
              RunH<true>(bins, gq, seq, rows, n_seq, n_sms, sink), RunH<true>(bins, gq, perm, rows, n_sparse, n_sms, sink));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                 CUtensorMapFloatOOBfill)>(fn);
    CUtensorMap tm;
    cuuint64_t dims[2] = {kRowBytes, rows};
    cuuint64_t strides[1] = {kRowBytes};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, bins, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { std::printf("cuTensorMapEncodeTiled -> %d\n", (int)rc); return 1; }
    std::printf("%-66s %8.2f %8.2f\n", "I  symbols by TMA gather4 (4-stage ring per warp) + LDS, pair/ids LDG",
                RunI(tm, gq, seq, n_seq, n_sms, sink), RunI(tm, gq, perm, n_sparse, n_sms, sink));
  }
  std::printf("%-66s %8.2f %8.2f\n", "G4 the same with 4 buffers",
              RunG<4>(bins, gq, seq, n_seq, n_sms, sink), RunG<4>(bins, gq, perm, n_sparse, n_sms, sink));
  return 0;
}
