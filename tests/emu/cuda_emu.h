// A small CPU emulator for CUDA kernels -- TEST INFRASTRUCTURE (tests/emu/): it lets the kernel sources of
// xgboost_b200/csrc execute, unmodified, against the oracle where there is no GPU (this container; the `not gpu` suite).
//
// Model: every CUDA thread is an OS thread; the CTAs of a launch run one after another.  __syncthreads and the warp
// collectives (__shfl_*_sync, __ballot_sync, __any_sync, __reduce_or_sync, __syncwarp) are barriers that tolerate
// threads which have already returned.  `__shared__` variables become function statics (one CTA runs at a time);
// dynamic shared memory comes from B200_DYNAMIC_SHARED (common.cuh).  Atomics are GCC __atomic builtins.
// What this checks: indexing, control flow, arithmetic and the data flow through shared memory / shuffles of the
// kernel text.  What it cannot check: anything that depends on the hardware's memory model, scheduling or on
// device-only arithmetic (__fdividef is an IEEE division here, as in the oracle).
#pragma once
#define B200_EMU 1
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>
// a rank of a multi-process emulation that waits for a peer's flag gives its core to that peer
#define B200_SPIN_PAUSE() sched_yield()

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n)
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
// aligned like the device types, so that -fsanitize=undefined reports a vector access the hardware would fault on
struct alignas(16) longlong2 { long long x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
inline longlong2 make_longlong2(long long a, long long b) { return longlong2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

// ---- the bits of the runtime API that common.cuh names (never executed under B200_KERNELS_ONLY) ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
inline const char* cudaGetErrorName(cudaError_t) { return "emu"; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n); return 0; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(std::malloc(n)); return 0; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = attr == 97 ? 232448 : 1;  // 97 = cudaDevAttrMaxSharedMemoryPerBlockOptin: B200's 227 KiB; anything else: 1
  return 0;
}

namespace emu {

class Barrier {  // reusable; a thread that returns from the kernel drops out
 public:
  explicit Barrier(int n) : expected_(n) {}
  void ArriveAndWait() {
    std::unique_lock<std::mutex> l(m_);
    const uint64_t g = gen_;
    if (++waiting_ >= expected_) {
      waiting_ = 0;
      Advance();
      l.unlock();
      cv_.notify_all();
      return;
    }
    if (expected_ <= 32) {  // warp-sized groups meet within microseconds: spin a little before sleeping
      l.unlock();
      for (int spin = 0; spin < 2000; ++spin) {
        if (__atomic_load_n(&gen_, __ATOMIC_ACQUIRE) != g) return;
        if ((spin & 63) == 63) std::this_thread::yield();
      }
      l.lock();
    }
    cv_.wait(l, [&] { return gen_ != g; });
  }
  void Drop() {
    std::unique_lock<std::mutex> l(m_);
    --expected_;
    if (expected_ > 0 && waiting_ >= expected_) { waiting_ = 0; Advance(); l.unlock(); cv_.notify_all(); }
  }
 private:
  // a release store: the spinning waiters leave on an acquire load without taking the mutex, and what the other
  // threads wrote before the barrier must be ordered before what they read after it (for -fsanitize=thread as well)
  void Advance() { __atomic_store_n(&gen_, gen_ + 1, __ATOMIC_RELEASE); }
  std::mutex m_;
  std::condition_variable cv_;
  int expected_, waiting_{0};
  uint64_t gen_{0};
};

struct Warp {
  explicit Warp(int n) : bar(n) {}
  Barrier bar;
  uint64_t slot[32];
  bool alive[32] = {};
};
struct Block {
  explicit Block(int n) : bar(n) {}
  Barrier bar;
  std::vector<unsigned char> dyn;
};
struct Ctx { Block* block; Warp* warp; int lane; };
inline thread_local Ctx g_ctx{nullptr, nullptr, 0};
inline void* DynShared() { return g_ctx.block->dyn.data(); }

template <typename T> inline uint64_t ToSlot(T v) { uint64_t s = 0; static_assert(sizeof(T) <= 8, "shuffle width"); std::memcpy(&s, &v, sizeof(T)); return s; }
template <typename T> inline T FromSlot(uint64_t s) { T v; std::memcpy(&v, &s, sizeof(T)); return v; }
template <typename T> inline T Exchange(T v, int src) {  // every live lane publishes, then reads lane `src`
  Warp* w = g_ctx.warp;
  w->slot[g_ctx.lane] = ToSlot(v);
  w->bar.ArriveAndWait();
  const T r = (src >= 0 && src < 32 && w->alive[src]) ? FromSlot<T>(w->slot[src]) : v;
  w->bar.ArriveAndWait();
  return r;
}

// grid x block launch of `body` (a callable that invokes the kernel with its arguments)
void Launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);

}  // namespace emu

inline thread_local uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0};
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { emu::g_ctx.block->bar.ArriveAndWait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_ctx.warp->bar.ArriveAndWait(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu::Exchange(v, src & 31); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::Exchange(v, emu::g_ctx.lane ^ m); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu::Exchange(v, emu::g_ctx.lane - d); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu::Exchange(v, emu::g_ctx.lane + d); }
inline unsigned __ballot_sync(unsigned, bool p) {
  emu::Warp* w = emu::g_ctx.warp;
  w->slot[emu::g_ctx.lane] = p ? 1u : 0u;
  w->bar.ArriveAndWait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) if (w->alive[i] && w->slot[i]) m |= 1u << i;
  w->bar.ArriveAndWait();
  return m;
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  emu::Warp* w = emu::g_ctx.warp;
  w->slot[emu::g_ctx.lane] = v;
  w->bar.ArriveAndWait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (w->alive[i]) r |= (unsigned)w->slot[i];
  w->bar.ArriveAndWait();
  return r;
}

using std::isnan;
using std::isinf;
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline float __fdividef(float a, float b) { return a / b; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }

template <typename T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T> inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

// CUDA's overloaded min / max (mixed signedness follows the usual arithmetic conversions there too)
template <typename A, typename B> inline typename std::common_type<A, B>::type min(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <typename A, typename B> inline typename std::common_type<A, B>::type max(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a > (C)b ? (C)a : (C)b;
}

// shared-window addresses (what __cvta_generic_to_shared yields on the device) are byte offsets into the CTA's dynamic
// shared memory here; SharedWord turns one back into a pointer
namespace emu {
inline uint32_t* SharedWord(uint32_t addr) { return reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(DynShared()) + addr); }
}  // namespace emu
inline size_t __cvta_generic_to_shared(const void* p) {
  return (size_t)(static_cast<const unsigned char*>(p) - static_cast<const unsigned char*>(emu::DynShared()));
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {  // PRMT, default mode
  const unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xF;
    unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
    if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;  // sign replication
    r |= byte << (8 * i);
  }
  return r;
}
inline long long clock64() { return 0; }
enum { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };

#define B200_DYNAMIC_SHARED(type, name) type* name = reinterpret_cast<type*>(::emu::DynShared())
#define B200_DYNAMIC_SHARED_ALIGNED(type, name, bytes) type* name = reinterpret_cast<type*>(::emu::DynShared())
#define EMU_LAUNCH(grid, block, smem, call) ::emu::Launch(dim3 grid, dim3 block, (smem), [&]() { call; })
