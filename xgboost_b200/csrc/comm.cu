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
#ifndef B200_KERNELS_ONLY
#include <dlfcn.h>
#endif

#include <cstring>

#include "common.cuh"

#ifndef B200_KERNELS_ONLY
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

#endif  // B200_KERNELS_ONLY

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
      while ((int32_t)(*f - epoch) < 0) { B200_SPIN_PAUSE(); /* the peer's stores are ordered before its flag */ }
    }
    __syncthreads();
    if (threadIdx.x == 0) *done_counter = 0;
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------------
// Row split (every rank owns a slice of the ROWS and builds full-width partial histograms): the level's
// histograms are SUMMED across ranks by one kernel over peer memory -- a reduce-scatter and an all-gather fused:
//   barrier A  every rank has finished building (a flag per peer, written through NVLink),
//   reduce     rank r owns the r-th slice of the level's flat bin range: it loads that slice from every peer
//              (P2P loads), adds in integer arithmetic (order-free, hence bit-identical to one GPU) and
//   broadcast  stores the sum into every rank's histogram, its own included; nobody else touches that slice,
//   barrier B  all stores have landed: the level's histograms are complete everywhere.
// Replaces collective::Allreduce over the built nodes (src/tree/gpu_hist/histogram.cu:598-609).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void SignalPeers(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch) {
  if (threadIdx.x < (unsigned)world && (int)threadIdx.x != rank) {
    volatile uint32_t* f = peer_flags[threadIdx.x] + rank;  // my slot in the peer's flag array
    *f = epoch;
  }
}
__device__ __forceinline__ void WaitPeers(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch) {
  if (threadIdx.x < (unsigned)world && (int)threadIdx.x != rank) {
    volatile uint32_t* f = peer_flags[rank] + threadIdx.x;  // the peer's slot in my flag array
    while ((int32_t)(*f - epoch) < 0) { B200_SPIN_PAUSE(); /* the peer's stores are ordered before its flag */ }
  }
}

template <int kWorld>  // 0 = run-time world size
__global__ void __launch_bounds__(256)
HistAllReduceKernel(longlong2* const* __restrict__ peer_hist, uint32_t* const* __restrict__ peer_flags, int rank,
                    int world, const TreeState* __restrict__ ts, uint32_t slot_base, uint32_t slot_cap,
                    uint32_t total_bins, uint32_t epoch_a, uint32_t epoch_b, unsigned int* __restrict__ done_counter) {
  const uint32_t n_slots = min(slot_cap, (uint32_t)ts->n_build);
  // barrier A: my build kernel precedes this kernel in the stream, so my partial sums are visible
  if (blockIdx.x == 0) {
    __threadfence_system();
    SignalPeers(peer_flags, rank, world, epoch_a);
  }
  WaitPeers(peer_flags, rank, world, epoch_a);
  __syncthreads();
  const unsigned long long total = (unsigned long long)n_slots * total_bins;
  const unsigned long long c0 = total * (unsigned)rank / (unsigned)world, c1 = total * (unsigned)(rank + 1) / (unsigned)world;
  const size_t base = (size_t)slot_base * total_bins;
  // all peers' values of an element are requested before any is used: a remote load costs microseconds, and a
  // load-add-load-add chain serialises them per thread.  The common world sizes are compiled with a fixed trip
  // count so that the loads of one element (and of the next) are issued back to back.
  if (kWorld > 0) {
    const longlong2* src[kWorld > 0 ? kWorld : 1];
#pragma unroll
    for (int p = 0; p < kWorld; ++p) src[p] = peer_hist[p] + base;
    for (unsigned long long i = c0 + blockIdx.x * 256ull + threadIdx.x; i < c1;
         i += (unsigned long long)gridDim.x * 256ull) {
      longlong2 v[kWorld > 0 ? kWorld : 1];
#pragma unroll
      for (int p = 0; p < kWorld; ++p) v[p] = src[p][i];
      longlong2 acc = make_longlong2(0, 0);
#pragma unroll
      for (int p = 0; p < kWorld; ++p) {
        acc.x += v[p].x;
        acc.y += v[p].y;
      }
#pragma unroll
      for (int p = 0; p < kWorld; ++p) peer_hist[p][base + i] = acc;
    }
  } else {
    for (unsigned long long i = c0 + blockIdx.x * 256ull + threadIdx.x; i < c1;
         i += (unsigned long long)gridDim.x * 256ull) {
      longlong2 acc = make_longlong2(0, 0);
      for (int p = 0; p < world; ++p) {
        const longlong2 v = peer_hist[p][base + i];
        acc.x += v.x;
        acc.y += v.y;
      }
      for (int p = 0; p < world; ++p) peer_hist[p][base + i] = acc;
    }
  }
  // barrier B: the last CTA to finish publishes the epoch and waits for every peer's
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    SignalPeers(peer_flags, rank, world, epoch_b);
    __threadfence_system();
    WaitPeers(peer_flags, rank, world, epoch_b);
    __syncthreads();
    if (threadIdx.x == 0) *done_counter = 0;
    __threadfence_system();
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchHistAllReduce(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                         const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, uint32_t total_bins,
                         uint32_t epoch_a, uint32_t epoch_b, unsigned int* done_counter, int num_sms, cudaStream_t s) {
  auto kernel = world == 2 ? HistAllReduceKernel<2>
                : world == 4 ? HistAllReduceKernel<4>
                : world == 8 ? HistAllReduceKernel<8>
                             : HistAllReduceKernel<0>;
  kernel<<<num_sms * 2, 256, 0, s>>>(peer_hist, peer_flags, rank, world, ts, slot_base, slot_cap, total_bins, epoch_a,
                                     epoch_b, done_counter);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// Row split, the two scalar exchanges of a tree: every rank posts one 64-byte record into slot `rank` of each
// peer's mailbox (it lives behind the flag array of the shared allocation), then every rank folds the records
// in rank order -- the same order everywhere, so all ranks derive bit-identical values.
//   which = 0: clipped gradient sums + row count -> the GLOBAL quantiser scale   (quantiser.cu:88-123 GlobalSum)
//   which = 1: the exact int64 root sum and the hessian flags                     (updater_gpu_hist.cu:661-668)
struct MailRecord {
  double d[4];
  unsigned long long u[4];
};
static_assert(sizeof(MailRecord) == 64, "mailbox record is 64 bytes");
constexpr uint32_t kMailboxOffset = 256;  // bytes after the flag array

__device__ __forceinline__ double RoundingFactorD(double max_abs, unsigned long long n) {
  double delta = max_abs / (1.0 - 2.0 * (double)n * 2.220446049250313e-16);  // deterministic.cuh:24-38
  int e;
  (void)frexp(delta, &e);
  return ldexp(1.0, e);
}

__global__ void __launch_bounds__(32)
ScalarExchangeKernel(uint32_t* const* __restrict__ peer_flags, int rank, int world, uint32_t epoch, int which,
                     QuantDev* __restrict__ q, unsigned long long n_local) {
  MailRecord rec;
  if (which == 0) {
    for (int k = 0; k < 4; ++k) { rec.d[k] = q->sums[k]; rec.u[k] = 0ull; }
    rec.u[0] = n_local;
  } else {
    for (int k = 0; k < 4; ++k) rec.d[k] = 0.0;
    rec.u[0] = (unsigned long long)q->root_g;
    rec.u[1] = (unsigned long long)q->root_h;
    rec.u[2] = q->h_lo_or;
    rec.u[3] = q->h_zero_any;
  }
  if (threadIdx.x < (unsigned)world) {
    auto* box = reinterpret_cast<volatile MailRecord*>(reinterpret_cast<uint8_t*>(peer_flags[threadIdx.x]) +
                                                       kMailboxOffset + which * kMaxPeers * sizeof(MailRecord));
    for (int k = 0; k < 4; ++k) { box[rank].d[k] = rec.d[k]; box[rank].u[k] = rec.u[k]; }
  }
  __threadfence_system();
  __syncwarp();
  SignalPeers(peer_flags, rank, world, epoch);
  WaitPeers(peer_flags, rank, world, epoch);
  __syncwarp();
  if (threadIdx.x == 0) {
    auto* box = reinterpret_cast<volatile MailRecord*>(reinterpret_cast<uint8_t*>(peer_flags[rank]) + kMailboxOffset +
                                                       which * kMaxPeers * sizeof(MailRecord));
    if (which == 0) {
      double t[4] = {0, 0, 0, 0};
      unsigned long long n = 0;
      for (int p = 0; p < world; ++p) {
        for (int k = 0; k < 4; ++k) t[k] += box[p].d[k];
        n += box[p].u[0];
      }
      const double mg = t[0] > t[2] ? t[0] : t[2];
      const double mh = t[1] > t[3] ? t[1] : t[3];
      const double rg = RoundingFactorD(mg, n), rh = RoundingFactorD(mh, n);
      const double div = 4611686018427387904.0;  // 2^62
      q->to_float_g = rg / div;
      q->to_float_h = rh / div;
      q->to_fixed_g = 1.0 / q->to_float_g;
      q->to_fixed_h = 1.0 / q->to_float_h;
      for (int k = 0; k < 4; ++k) q->sums[k] = t[k];
    } else {
      unsigned long long g = 0, h = 0, lo = 0, hz = 0;
      for (int p = 0; p < world; ++p) {
        g += box[p].u[0];
        h += box[p].u[1];
        lo |= box[p].u[2];
        hz |= box[p].u[3];
      }
      q->root_g = (long long)g;
      q->root_h = (long long)h;
      q->h_lo_or = (unsigned int)lo;
      q->h_zero_any = (unsigned int)hz;
    }
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchScalarExchange(uint32_t* const* peer_flags, int rank, int world, uint32_t epoch, int which, QuantDev* q,
                          uint64_t n_local, cudaStream_t s) {
  ScalarExchangeKernel<<<1, 32, 0, s>>>(peer_flags, rank, world, epoch, which, q, (unsigned long long)n_local);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// ---------------------------------------------------------------------------------------------
// Row split, evaluation by feature slice (the default row-split path): a reduce-scatter WITHOUT the all-gather.
//   HistSliceReduceKernel   barrier A (everyone has built), then rank r sums the peers' partial histograms over ITS
//                           bin range [bin_begin, bin_end) of every built slot (P2P loads) into its own cache.  Nobody
//                           writes into a peer: each rank ends up with the complete histograms of its features only.
//   (local)                 SubtractKernel and FeatureScanKernel over that bin / feature range.
//   ScratchExchangeKernel   every rank pushes the per-feature winners it owns (32 bytes per (node, feature)) into all
//                           peers' scratch tables and meets them at barrier B; the fold over features (NodeReduceKernel)
//                           then runs replicated on identical tables -- the reference's column-split evaluation
//                           (evaluate_splits.cu:412-435: allgather of candidates).
// Per level a rank reads (P-1)/P of the built histograms' bytes from its peers and writes 32 B x nodes x features / P
// to each: half the NVLink traffic of the fused reduce-scatter + all-gather, and evaluation / subtraction shrink by P.
// ---------------------------------------------------------------------------------------------
template <int kWorld>
__global__ void __launch_bounds__(256)
HistSliceReduceKernel(longlong2* const* __restrict__ peer_hist, uint32_t* const* __restrict__ peer_flags, int rank,
                      int world, const TreeState* __restrict__ ts, uint32_t slot_base, uint32_t slot_cap,
                      uint32_t total_bins, uint32_t bin_begin, uint32_t bin_end, uint32_t epoch) {
  const uint32_t n_slots = min(slot_cap, (uint32_t)ts->n_build);
  if (blockIdx.x == 0) {  // barrier A: my build kernel precedes this kernel in the stream
    __threadfence_system();
    SignalPeers(peer_flags, rank, world, epoch);
  }
  WaitPeers(peer_flags, rank, world, epoch);
  __syncthreads();
  const uint32_t width = bin_end - bin_begin;
  const unsigned long long total = (unsigned long long)n_slots * width;
  longlong2* __restrict__ mine = peer_hist[rank];
  if (kWorld > 0) {
    const longlong2* src[kWorld > 0 ? kWorld : 1];
#pragma unroll
    for (int p = 0; p < kWorld; ++p) src[p] = peer_hist[p];
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256ull) {
      const uint32_t sl = (uint32_t)(i / width);
      const size_t off = (size_t)(slot_base + sl) * total_bins + bin_begin + (uint32_t)(i - (unsigned long long)sl * width);
      longlong2 v[kWorld > 0 ? kWorld : 1];
#pragma unroll
      for (int p = 0; p < kWorld; ++p) v[p] = src[p][off];
      longlong2 acc = make_longlong2(0, 0);
#pragma unroll
      for (int p = 0; p < kWorld; ++p) { acc.x += v[p].x; acc.y += v[p].y; }
      mine[off] = acc;
    }
  } else {
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256ull) {
      const uint32_t sl = (uint32_t)(i / width);
      const size_t off = (size_t)(slot_base + sl) * total_bins + bin_begin + (uint32_t)(i - (unsigned long long)sl * width);
      longlong2 acc = make_longlong2(0, 0);
      for (int p = 0; p < world; ++p) {
        const longlong2 v = peer_hist[p][off];
        acc.x += v.x;
        acc.y += v.y;
      }
      mine[off] = acc;
    }
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchHistSliceReduce(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                           const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, uint32_t total_bins,
                           uint32_t bin_begin, uint32_t bin_end, uint32_t epoch, int num_sms, cudaStream_t s) {
  auto kernel = world == 2 ? HistSliceReduceKernel<2>
                : world == 4 ? HistSliceReduceKernel<4>
                : world == 8 ? HistSliceReduceKernel<8>
                             : HistSliceReduceKernel<0>;
  kernel<<<num_sms * 2, 256, 0, s>>>(peer_hist, peer_flags, rank, world, ts, slot_base, slot_cap, total_bins, bin_begin,
                                     bin_end, epoch);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// One thread per (evaluation slot, feature position): the entry travels iff this rank scanned it.  The feature of an
// item is looked up exactly as FeatureScanKernel does (evaluate.cu).
__global__ void __launch_bounds__(256)
ScratchExchangeKernel(FeatCand* const* __restrict__ peer_scratch, uint32_t* const* __restrict__ peer_flags,
                      EvalShard shard, const TreeState* __restrict__ ts, const int32_t* __restrict__ eval_nodes,
                      const NodeDev* __restrict__ nodes, const int32_t* __restrict__ feature_set, int32_t n_fs,
                      int32_t n_levels, IcView ic, uint32_t epoch, unsigned int* __restrict__ done_counter) {
  const int rank = shard.rank, world = shard.world;
  const long long total = (long long)ts->n_eval * n_fs;
  const FeatCand* __restrict__ mine = peer_scratch[rank];
  for (long long item = blockIdx.x * 256ll + threadIdx.x; item < total; item += (long long)gridDim.x * 256ll) {
    const int slot = (int)(item / n_fs);
    const int fpos = (int)(item - (long long)slot * n_fs);
    const int nidx = eval_nodes[slot];
    const int32_t* __restrict__ fs =
        (feature_set && n_levels > 0) ? feature_set + (size_t)min(nodes[nidx].depth, n_levels - 1) * n_fs : feature_set;
    if (ic.count != nullptr) {
      if (fpos >= ic.count[slot]) continue;
      fs = ic.list + (size_t)slot * ic.stride;
    }
    const int f = fs ? fs[fpos] : fpos;
    if (!shard.Owns(f)) continue;
    const FeatCand c = mine[item];
    for (int p = 0; p < world; ++p)
      if (p != rank) peer_scratch[p][item] = c;
  }
  // barrier B: the last CTA to finish publishes the epoch to every peer and waits for all peers
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    SignalPeers(peer_flags, rank, world, epoch);
    __threadfence_system();
    WaitPeers(peer_flags, rank, world, epoch);
    __syncthreads();
    if (threadIdx.x == 0) *done_counter = 0;
    __threadfence_system();
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchScratchExchange(FeatCand* const* peer_scratch, uint32_t* const* peer_flags, const EvalShard& shard,
                           const TreeState* ts, const int32_t* eval_nodes, const NodeDev* nodes,
                           const int32_t* feature_set, int32_t n_fs, int32_t n_levels, IcView ic, int32_t max_eval,
                           uint32_t epoch, unsigned int* done_counter, int num_sms, cudaStream_t s) {
  const long long items = (long long)max_eval * n_fs;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, (long long)num_sms * 2));
  ScratchExchangeKernel<<<blocks, 256, 0, s>>>(peer_scratch, peer_flags, shard, ts, eval_nodes, nodes, feature_set,
                                               n_fs, n_levels, ic, epoch, done_counter);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

#ifndef B200_KERNELS_ONLY
void LaunchHistExchange(longlong2* const* peer_hist, uint32_t* const* peer_flags, int rank, int world,
                        const TreeState* ts, uint32_t slot_base, uint32_t slot_cap, bool use_n_build,
                        uint32_t total_bins, uint32_t bin_begin, uint32_t bin_end, uint32_t epoch,
                        unsigned int* done_counter, int num_sms, cudaStream_t s) {
  HistExchangeKernel<<<num_sms * 2, 256, 0, s>>>(peer_hist, peer_flags, rank, world, ts, slot_base, slot_cap,
                                                 use_n_build ? 1 : 0, total_bins, bin_begin, bin_end, epoch,
                                                 done_counter);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

}  // namespace b200

using namespace b200;

#ifndef B200_KERNELS_ONLY
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
#endif  // B200_KERNELS_ONLY
