// Microbenchmark for DESIGN.md (f) item 1: can TMA row gathers (UTMALDG.2D.GATHER4, see tools/gather4_probe.cu) feed
// the histogram kernel faster than per-lane global loads do?  The histogram kernel spends 23 of its 95 busy points of
// the L1TEX data stage on gather wavefronts (one per gathered row for the symbols, one for the gradient pair); a TMA
// copy lands in shared memory without them, but the earlier one-row-per-copy ring lost to the TMA issue rate.
// gather4 moves four arbitrary rows per instruction.  This tool measures, on a row-major uint8 matrix
// [n_rows][row_bytes] and a random row list:
//   (a) baseline: every lane loads `box` bytes of its row with LDG (what the kernel does today),
//   (b) gather4 : one elected lane per warp keeps a ring of kStages gather4 copies in flight; the warp consumes a slot
//                 (reads it from shared memory) and refills it,
// for box = 32 / 64 / 128 bytes and 4 / 8 / 16 warps per CTA, one CTA per SM x 2, and prints rows per clock per SM
// plus a checksum that must agree between (a) and (b).  The histogram kernel needs ~0.18 rows/clk/SM per 32-feature
// block at today's speed (5.8 updates/clk/SM); the gradient pairs are the same pattern with a 16-byte box.
// Never run yet.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather4_throughput tools/gather4_throughput.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                  \
  do {                                                                                                         \
    cudaError_t e_ = (x);                                                                                      \
    if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } \
  } while (0)

// (a) per-lane loads: lane l of a warp handles row ridx[base + l]; reads `box` bytes as uint4 words
template <int kBox>
__global__ void __launch_bounds__(512) LdgGather(const uint8_t* __restrict__ bins, uint32_t row_bytes,
                                                 const uint32_t* __restrict__ ridx, uint64_t n_list,
                                                 unsigned long long* __restrict__ sink) {
  const uint64_t n_threads = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_list; i += n_threads) {
    const uint4* p = reinterpret_cast<const uint4*>(bins + (size_t)ridx[i] * row_bytes);
#pragma unroll
    for (int k = 0; k < kBox / 16; ++k) {
      const uint4 v = __ldg(p + k);
      acc += (unsigned long long)v.x + v.y + v.z + v.w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sink, acc);
}

__device__ __forceinline__ bool BarDone(uint32_t bar_a, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar_a), "r"(parity)
      : "memory");
  return done != 0;
}

// (b) gather4 ring per warp.  Slot layout per warp: kStages x (4 rows x kBox bytes), 128-byte aligned slots.
template <int kBox, int kStages>
__global__ void __launch_bounds__(512) Gather4Ring(const __grid_constant__ CUtensorMap tm,
                                                   const uint32_t* __restrict__ ridx, uint64_t n_list,
                                                   unsigned long long* __restrict__ sink, int* __restrict__ stuck) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kSlot = ((4 * kBox + 127) / 128) * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  uint8_t* ring = smem + (size_t)warp * kStages * kSlot;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)n_warps * kStages * kSlot) + warp * kStages;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  // this warp's share of the list, in groups of 4 rows
  const uint64_t n_groups = n_list / 4;
  const uint64_t warps_total = (uint64_t)gridDim.x * n_warps;
  const uint64_t wid = (uint64_t)blockIdx.x * n_warps + warp;
  const uint64_t g_begin = n_groups * wid / warps_total, g_end = n_groups * (wid + 1) / warps_total;
  auto issue = [&](uint64_t g, int s) {  // lane 0 only
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bars[s]);
    const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(ring + (size_t)s * kSlot);
    const uint4 r = *reinterpret_cast<const uint4*>(ridx + 4 * g);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(4 * kBox));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst_a), "l"(&tm), "r"(0), "r"((int)r.x), "r"((int)r.y), "r"((int)r.z), "r"((int)r.w), "r"(bar_a)
        : "memory");
  };
  if (lane == 0)
    for (int s = 0; s < kStages && g_begin + s < g_end; ++s) issue(g_begin + s, s);
  unsigned long long acc = 0;
  uint32_t parity = 0;
  int s = 0;
  for (uint64_t g = g_begin; g < g_end; ++g) {
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bars[s]);
    int spin = 0;
    while (!BarDone(bar_a, parity)) {
      if (++spin > (1 << 22)) { if (lane == 0) atomicAdd(stuck, 1); return; }  // never hang the GPU
    }
    // consume: 4 x kBox bytes as 32-bit words spread over the lanes
    const uint32_t* slot = reinterpret_cast<const uint32_t*>(ring + (size_t)s * kSlot);
    for (int w = lane; w < kBox; w += 32) acc += slot[w];  // kBox words = 4 * kBox bytes
    __syncwarp();
    if (lane == 0 && g + kStages < g_end) issue(g + kStages, s);
    if (++s == kStages) { s = 0; parity ^= 1u; }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) atomicAdd(sink, acc);
}

template <int kBox, int kStages>
static void RunCase(const CUtensorMap& tm, const uint8_t* d_bins, uint32_t row_bytes, const uint32_t* d_ridx,
                    uint64_t n_list, int sms, double clock_ghz, int warps) {
  unsigned long long *d_sink = nullptr, h_a = 0, h_b = 0;
  int *d_stuck = nullptr, h_stuck = 0;
  CK(cudaMalloc(&d_sink, 8));
  CK(cudaMalloc(&d_stuck, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = sms * 2, threads = warps * 32;
  float ms_a = 0, ms_b = 0;
  for (int rep = 0; rep < 3; ++rep) {  // last repetition is the one reported
    CK(cudaMemset(d_sink, 0, 8));
    CK(cudaEventRecord(e0));
    LdgGather<kBox><<<grid, threads>>>(d_bins, row_bytes, d_ridx, n_list, d_sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_a, e0, e1));
  }
  CK(cudaMemcpy(&h_a, d_sink, 8, cudaMemcpyDeviceToHost));
  constexpr int kSlot = ((4 * kBox + 127) / 128) * 128;
  const size_t smem = (size_t)warps * kStages * kSlot + (size_t)warps * kStages * 8;
  CK(cudaFuncSetAttribute(Gather4Ring<kBox, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(d_sink, 0, 8));
    CK(cudaMemset(d_stuck, 0, 4));
    CK(cudaEventRecord(e0));
    Gather4Ring<kBox, kStages><<<grid, threads, smem>>>(tm, d_ridx, n_list, d_sink, d_stuck);
    CK(cudaEventRecord(e1));
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { std::printf("  gather4 box %d stages %d warps %d: %s\n", kBox, kStages, warps, cudaGetErrorString(e)); std::exit(1); }
    CK(cudaEventElapsedTime(&ms_b, e0, e1));
  }
  CK(cudaMemcpy(&h_b, d_sink, 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_stuck, d_stuck, 4, cudaMemcpyDeviceToHost));
  const uint64_t rows4 = (n_list / 4) * 4;  // the ring handles whole groups
  const double per_clk_a = (double)n_list / (ms_a * 1e-3) / (clock_ghz * 1e9) / sms;
  const double per_clk_b = (double)rows4 / (ms_b * 1e-3) / (clock_ghz * 1e9) / sms;
  std::printf("box %3d B  stages %d  warps %2d : LDG %.3f ms (%.3f rows/clk/SM)   gather4 %.3f ms (%.3f rows/clk/SM)   "
              "checksum %s%s\n",
              kBox, kStages, warps, ms_a, per_clk_a, ms_b, per_clk_b,
              (h_a == h_b || n_list % 4 != 0) ? "equal" : "DIFFERS", h_stuck ? "  [a warp gave up waiting]" : "");
  cudaFree(d_sink);
  cudaFree(d_stuck);
}

int main(int argc, char** argv) {
  const uint64_t n_rows = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 20000000ull;
  const uint32_t row_bytes = 256;
  const uint64_t n_list = n_rows / 2 / 4 * 4;  // a node holding half of the rows, ascending row ids as in a real segment
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double clock_ghz = khz * 1e-6;
  std::printf("%s, %d SMs, %.2f GHz nominal; matrix %llu x %u B, row list %llu\n", prop.name, sms, clock_ghz,
              (unsigned long long)n_rows, row_bytes, (unsigned long long)n_list);
  uint8_t* d_bins = nullptr;
  CK(cudaMalloc(&d_bins, n_rows * row_bytes));
  CK(cudaMemset(d_bins, 1, n_rows * row_bytes));
  std::vector<uint32_t> h_ridx(n_list);
  uint64_t state = 88172645463325252ull;
  uint32_t r = 0;
  for (uint64_t i = 0; i < n_list; ++i) {  // every row with probability 1/2, ascending
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    r += 1 + (uint32_t)(state & 1) * 2;
    h_ridx[i] = r % (uint32_t)n_rows;
  }
  uint32_t* d_ridx = nullptr;
  CK(cudaMalloc(&d_ridx, n_list * 4));
  CK(cudaMemcpy(d_ridx, h_ridx.data(), n_list * 4, cudaMemcpyHostToDevice));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                               CUtensorMapFloatOOBfill)>(fn);
  auto make_map = [&](uint32_t box_cols, CUtensorMap* tm) {
    cuuint64_t dims[2] = {row_bytes, n_rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_cols, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_bins, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { std::printf("cuTensorMapEncodeTiled(box %u) -> %d\n", box_cols, (int)rc); std::exit(1); }
  };
  CUtensorMap tm32, tm64, tm128;
  make_map(32, &tm32);
  make_map(64, &tm64);
  make_map(128, &tm128);
  for (int warps : {4, 8, 16}) {
    RunCase<32, 4>(tm32, d_bins, row_bytes, d_ridx, n_list, sms, clock_ghz, warps);
    RunCase<64, 4>(tm64, d_bins, row_bytes, d_ridx, n_list, sms, clock_ghz, warps);
    RunCase<64, 8>(tm64, d_bins, row_bytes, d_ridx, n_list, sms, clock_ghz, warps);
    RunCase<128, 4>(tm128, d_bins, row_bytes, d_ridx, n_list, sms, clock_ghz, warps);
  }
  cudaFree(d_ridx);
  cudaFree(d_bins);
  return 0;
}
