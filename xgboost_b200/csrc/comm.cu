// Histogram allreduce over NVLink/NVSwitch: the one collective of the hot path.
//
// Replaces (reference paths relative to /root/reference):
//   DeviceHistogramBuilder::AllReduceHist      src/tree/gpu_hist/histogram.cu:598-609
//   collective::Allreduce -> NCCLColl          src/collective/aggregator.h:138-160, src/collective/coll.cu:251-277
//   NCCLComm bootstrap                         src/collective/comm.cu:45-101
//   dlopen of NCCL                             src/collective/nccl_stub.cc
//
// One process per GPU.  The communicator is created from a 128-byte NCCL unique id that the caller
// broadcasts with whatever control plane it has (bench.py: torch.distributed); the allreduce itself is
// issued by the C++ tree driver on the SAME stream as the histogram kernels, once per tree level, over
// the level's contiguous histogram slots (int64 sum).  With feature-block sharding the ranks' slices are
// disjoint, so the sum is exact and order-free: results are bit-identical to one GPU.  No helper thread,
// no host synchronisation (the reference blocks the host on every collective, coll.cu:274-276).
// NCCL is resolved with dlopen so that single-GPU use does not need it.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace b200 {

struct NcclUniqueId {
  char internal[128];
};
using NcclComm = void*;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* handle = nullptr;
};

static NcclApi& Nccl() {
  static NcclApi api = [] {
    NcclApi a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (!a.handle) throw Error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* s) {
      void* p = dlsym(a.handle, s);
      if (!p) throw Error(std::string("NCCL symbol missing: ") + s);
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    return a;
  }();
  return api;
}

#define B200_NCCL(expr)                                                                       \
  do {                                                                                        \
    int r__ = (expr);                                                                         \
    if (r__ != 0) throw ::b200::Error(std::string("NCCL error in " #expr ": ") + Nccl().GetErrorString(r__)); \
  } while (0)

}  // namespace b200

namespace b200 {

// int64 sum over all ranks, in place, on stream s
void CommAllReduceInt64(b200hist_comm_s* c, void* buf, size_t count, cudaStream_t s) {
  constexpr int kNcclInt64 = 4, kNcclSum = 0;
  B200_NCCL(Nccl().AllReduce(buf, buf, count, kNcclInt64, kNcclSum, c->comm, s));
}

// Contiguous split of n_groups feature groups over `world` ranks (rank r owns [g0, g1)).
void ShardGroups(uint32_t n_groups, int world, int rank, uint32_t* g0, uint32_t* g1) {
  const uint32_t base = n_groups / (uint32_t)world, rem = n_groups % (uint32_t)world;
  const uint32_t r = (uint32_t)rank;
  *g0 = r * base + std::min(r, rem);
  *g1 = *g0 + base + (r < rem ? 1u : 0u);
}

}  // namespace b200

namespace b200 {

// ---------------------------------------------------------------------------------------------
// Peer-memory histogram exchange (NVLink / NVSwitch P2P), the default multi-GPU collective.
//
// With feature-block sharding the level's "allreduce" is really an all-gather of disjoint slices, so instead
// of calling NCCL every rank PUSHES the bin ranges it owns of the level's built histograms straight into the
// same addresses of every peer's histogram cache (16-byte stores through IPC-mapped peer pointers), then the
// ranks meet at a flag barrier: after a system-scope fence each rank writes the exchange's epoch into its slot
// of every peer's flag array and spins until all of its own slots show the epoch.  One launch per level on the
// kernels' stream; no host involvement, no extra buffer, half the wire bytes of a ring allreduce.
// The reference's counterpart is collective::Allreduce -> ncclAllReduce with a host-side block
// (src/tree/gpu_hist/histogram.cu:598-609, src/collective/coll.cu:251-277).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
HistExchangeKernel(longlong2* const* __restrict__ peer_hist, uint32_t* const* __restrict__ peer_flags, int rank,
                   int world, const TreeState* __restrict__ ts, uint32_t slot_base, uint32_t slot_cap, int use_n_build,
                   uint32_t total_bins, uint32_t bin_begin, uint32_t bin_end, uint32_t epoch,
                   unsigned int* __restrict__ done_counter) {
  const uint32_t n_slots = use_n_build ? min(slot_cap, (uint32_t)ts->n_build) : slot_cap;
  const uint32_t width = bin_end - bin_begin;
  const unsigned long long total = (unsigned long long)n_slots * width;
  const longlong2* __restrict__ mine = peer_hist[rank];
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256ull) {
    const uint32_t s = (uint32_t)(i / width);
    const uint32_t b = bin_begin + (uint32_t)(i - (unsigned long long)s * width);
    const size_t off = (size_t)(slot_base + s) * total_bins + b;
    const longlong2 v = mine[off];
    for (int p = 0; p < world; ++p)
      if (p != rank) peer_hist[p][off] = v;
  }
  // barrier: the last CTA to finish publishes the epoch to every peer and waits for all peers
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < (unsigned)world && (int)threadIdx.x != rank) {
      volatile uint32_t* f = peer_flags[threadIdx.x] + rank;  // my slot in peer's flag array
      *f = epoch;
    }
    __threadfence_system();
    if (threadIdx.x < (unsigned)world && (int)threadIdx.x != rank) {
      volatile uint32_t* f = peer_flags[rank] + threadIdx.x;  // peer's slot in my flag array
      while ((int32_t)(*f - epoch) < 0) { /* spin: the peer's stores are ordered before its flag */ }
    }
    __syncthreads();
    if (threadIdx.x == 0) *done_counter = 0;
    __threadfence_system();
  }
}

void LaunchHistExchange(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                        const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, bool use_n_build,
                        uint32_t total_bins, uint32_t bin_begin, uint32_t bin_end, uint32_t epoch,
                        unsigned int* done_counter, int num_sms, cudaStream_t s) {
  HistExchangeKernel<<<num_sms * 2, 256, 0, s>>>(peer_hist, peer_flags, rank, world, ts, slot_base, slot_cap,
                                                 use_n_build ? 1 : 0, total_bins, bin_begin, bin_end, epoch,
                                                 done_counter);
  B200_LAUNCHED();
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200hist_comm_unique_id(uint8_t* out_id128) {
  B200_API_BEGIN();
  B200_CHECK(out_id128 != nullptr, "null argument");
  NcclUniqueId id;
  B200_NCCL(Nccl().GetUniqueId(&id));
  std::memcpy(out_id128, id.internal, 128);
  B200_API_END();
}

int b200hist_comm_create(int device, int world, int rank, const uint8_t* id128, b200hist_comm_t* out) {
  B200_API_BEGIN();
  B200_CHECK(id128 != nullptr && out != nullptr, "null argument");
  B200_CHECK(world >= 1 && rank >= 0 && rank < world, "bad world / rank");
  DeviceGuard guard(device);
  auto* c = new b200hist_comm_s();
  c->device = device;
  c->world = world;
  c->rank = rank;
  NcclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  B200_NCCL(Nccl().CommInitRank(&c->comm, world, id, rank));
  *out = c;
  B200_API_END();
}

int b200hist_comm_destroy(b200hist_comm_t c) {
  B200_API_BEGIN();
  if (c) {
    if (c->comm) Nccl().CommDestroy(c->comm);
    delete c;
  }
  B200_API_END();
}

int b200hist_shard_groups(uint32_t n_groups, int world, int rank, uint32_t* g_begin, uint32_t* g_end) {
  B200_API_BEGIN();
  B200_CHECK(g_begin != nullptr && g_end != nullptr, "null argument");
  B200_CHECK(world >= 1 && rank >= 0 && rank < world, "bad world / rank");
  ShardGroups(n_groups, world, rank, g_begin, g_end);
  B200_API_END();
}

}  // extern "C"
