// Probe for the TMA row-gather path the histogram kernel could use next:
//   cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4  ->  SASS UTMALDG.2D.GATHER4 (verified with nvcc 12.9,
//   -gencode arch=compute_100a,code=sm_100a; no public documentation in this environment, hence this probe).
// It builds a uint8 matrix [n_rows][row_bytes] with row r filled with (r + column) mod 251, asks the TMA for 4
// arbitrary rows x `box_cols` bytes starting at column `col0`, and prints what landed in shared memory, for a few
// candidate tensor-map box shapes.  RESULT on a B200 (driver 580, CUDA 12.9): with boxDim = {cols, 1} the four rows
// {r0, r1, r2, r3} land back to back in shared memory (slot k = row r_k, cols bytes each) and the mbarrier completes
// after 4 * cols bytes; boxDim = {cols, 4} raises "illegal instruction".  Build and run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe tools/gather4_probe.cu -lcuda && ./gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } \
  } while (0)

__global__ void GatherKernel(const __grid_constant__ CUtensorMap tm, int col0, int r0, int r1, int r2, int r3,
                             uint32_t bytes, uint8_t* out, uint32_t dst_offset) {
  extern __shared__ __align__(128) uint8_t raw[];
  uint8_t* buf = raw + dst_offset;  // open question for the ring layout: how aligned must the destination be?
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t buf_a = (uint32_t)__cvta_generic_to_shared(buf);
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) buf[i] = 0xEE;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(buf_a), "l"(&tm), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a)
        : "memory");
  }
  // bounded, non-blocking wait: a wrong byte count must not hang the GPU
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 20) && !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
    if (!done) __nanosleep(200);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
  if (threadIdx.x == 0) out[bytes] = (uint8_t)done;
}

int main() {
  const uint64_t n_rows = 4096, row_bytes = 256;
  std::vector<uint8_t> h(n_rows * row_bytes);
  for (uint64_t r = 0; r < n_rows; ++r)
    for (uint64_t c = 0; c < row_bytes; ++c) h[r * row_bytes + c] = (uint8_t)((r + c) % 251);
  uint8_t* d = nullptr;
  CK(cudaMalloc(&d, h.size()));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                               CUtensorMapFloatOOBfill)>(fn);
  const uint32_t box_cols = 64;
  const uint32_t candidates[][2] = {{box_cols, 1}, {box_cols, 4}};  // {box dim 0 (bytes), box dim 1 (rows)}
  for (auto const& box_shape : candidates) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {row_bytes, n_rows};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_shape[0], box_shape[1]};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    std::printf("box {%u, %u}: cuTensorMapEncodeTiled -> %d\n", box[0], box[1], (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    const uint32_t bytes = 4 * box_cols;
    uint8_t* out = nullptr;
    CK(cudaMalloc(&out, bytes + 16));
    const int rows[4] = {7, 1000, 33, 4095};
    GatherKernel<<<1, 128, bytes + 256>>>(tm, 32, rows[0], rows[1], rows[2], rows[3], bytes, out, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> got(bytes + 16);
    CK(cudaMemcpy(got.data(), out, bytes + 1, cudaMemcpyDeviceToHost));
    std::printf("  barrier completed: %d\n", (int)got[bytes]);
    for (int k = 0; k < 4; ++k) {
      int ok = 1;
      for (uint32_t c = 0; c < box_cols; ++c) ok &= got[k * box_cols + c] == (uint8_t)((rows[k] + 32 + c) % 251);
      std::printf("  slot %d: first bytes %u %u %u ... expected row %d col 32 -> %s\n", k, got[k * box_cols],
                  got[k * box_cols + 1], got[k * box_cols + 2], rows[k], ok ? "MATCH" : "differs");
    }
    // destination alignment: the histogram kernel's per-warp ring wants 64-byte (gradient pairs: 4 x 16 B) and
    // 256-byte (symbols: 4 x 64 B) slots packed back to back
    for (uint32_t off : {128u, 64u, 16u}) {
      GatherKernel<<<1, 128, bytes + 256>>>(tm, 32, rows[0], rows[1], rows[2], rows[3], bytes, out, off);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        std::printf("  dst offset %u: %s (context is gone, stopping)\n", off, cudaGetErrorString(e));
        return 1;
      }
      CK(cudaMemcpy(got.data(), out, bytes + 1, cudaMemcpyDeviceToHost));
      int ok = got[bytes] != 0;
      for (int k = 0; k < 4 && ok; ++k)
        for (uint32_t c = 0; c < box_cols; ++c) ok &= got[k * box_cols + c] == (uint8_t)((rows[k] + 32 + c) % 251);
      std::printf("  dst offset %u: %s\n", off, ok ? "MATCH" : "differs / barrier not completed");
    }
    cudaFree(out);
    break;  // the {cols, 4} box is known to raise an illegal instruction: not retried
  }
  cudaFree(d);
  return 0;
}
