// Gradient histogram build (BuildHist) and histogram subtraction.
//
// Replaces (reference paths relative to /root/reference):
//   FeatureGroups                       src/tree/gpu_hist/feature_groups.cu:14-46
//   DeviceHistogramBuilder::BuildHistogram / StHistKernel / HistKernelOneNodeTarget
//                                       src/tree/gpu_hist/histogram.cu:134-244,382-420,570-582
//   SubtractionTrick                    src/tree/gpu_hist/histogram.cuh:208-221
//
// Semantics (SURVEY 8(a')): H[n][b] = sum over rows r of node n and features f with bin(r,f)=b of
// (gq[r], hq[r]) in exact int64.  Integer addition is associative, so any decomposition of the sum
// gives bit-identical histograms.
//
// Design (differs from the reference on purpose):
//   * ONE launch builds the histograms of every node of a tree level (the reference launches once
//     per node).  Work = row tiles of all build nodes (tile table written on the device by the
//     partition scan, no host sync); each persistent CTA owns a CONTIGUOUS range of tiles for one
//     feature group, so its shared-memory histogram is flushed only when the node changes.
//   * BANK-MAPPED shared histogram (uint8 pages, HistKernelBanked): the privatised histogram is stored
//     as four planes {g.lo, g.hi, h.lo, h.hi} of [block][symbol][32 lanes] 32-bit words, i.e. feature
//     (f mod 32) lives entirely in shared-memory bank (f mod 32).  A warp processes 4 rows x 32
//     features per pass: lane (q, w) loads one 32-bit word (4 symbols) of row q and in step k uses
//     byte (k + q) mod 4, so in every step the 32 lanes touch 32 DIFFERENT features = 32 different
//     banks.  Every shared atomic of the inner loop is therefore bank-conflict free by construction
//     (the reference's layout, and our first version, alias random bins onto few banks; the first
//     version measured 0.71 updates/clk/SM on B200 because of it).
//   * int64 accumulation = 32-bit shared atomic on the low word whose returned old value yields the
//     carry into the high word (two ATOMS per 64-bit add, exact).
//   * flush to global memory uses 64-bit reductions (RED.64) on non-zero bins only.
//   * Default kernels: HistKernelBankedFast<sym bytes> (32 features x 16-byte accumulators per CTA pass) and
//     HistKernelPair<bufs, sym bytes> (64 features x 12-byte accumulators, hessians without low words).  No
//     predicated variant: rows are zero-padded to whole 32-feature blocks and the null symbol of a page with
//     missing values owns an accumulator slot; both are dropped at flush time.  uint16 pages (257..~440 symbols per
//     feature, e.g. 256 bins + missing values) run the same kernels with two registers of symbols per lane; only
//     wider pages use the generic kernel (HistKernelGeneric).
//   * The binding resource is the L1TEX data stage (shared atomics + global-load wavefronts): ncu shows it at
//     94-95 % of peak for the pair kernel (profiles/r01c_pair_kernels_ncu_summary.json).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace b200 {

constexpr int kHistThreads = 1024;
constexpr uint32_t kMaxGroupFeatures = 512;

static inline size_t Round16(size_t x) { return (x + 15) & ~size_t(15); }

static size_t GenericSmemBytes(uint32_t group_bins) {
  return Round16((size_t)group_bins * 16) + (size_t)kHistTile * 16 + (size_t)kHistTile * 4 + kMaxGroupFeatures * 4;
}
static size_t BankedSmemBytes(const GroupDesc& g) { return (size_t)4 * g.n_blocks * g.max_sym * 32 * 4; }

static int SmemOptin(int device) {
  int optin = 0;
  B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  return optin;
}

// Histogram kernel selection (B200HIST_HIST_IMPL):
//   0 = direct global loads, software pipelined in registers (HistKernelBanked)  [default: fastest, see DESIGN.md]
//   1 = per-warp TMA bulk-copy ring (HistKernelStaged<.., true>)
//   2 = per-warp cp.async ring      (HistKernelStaged<.., false>)
static int HistImpl() {
  static int impl = [] {
    const char* e = std::getenv("B200HIST_HIST_IMPL");
    return e ? std::atoi(e) : 0;
  }();
  return impl;
}
constexpr uint32_t kRingStages = 6;
static size_t RingSmemBytes(uint32_t blocks) {
  return (HistImpl() == 0 || HistImpl() == 3) ? 0 : (size_t)(kHistThreads / 32) * kRingStages * (8 * (32 * blocks + 16) + 8);
}

// Feature groups.
//  uint8 pages : groups of 32-feature blocks for the bank-mapped kernel; a group holds as many blocks as
//                fit (4 planes x blocks x max_sym x 32 words).
//  uint16 pages: greedy contiguous ranges whose bins fit a linear shared histogram (generic kernel).
std::vector<GroupDesc> MakeFeatureGroups(const b200hist_ellpack_s& ell, int device) {
  std::vector<GroupDesc> groups;
  const size_t optin = (size_t)SmemOptin(device);
  if (ell.is_sparse) {
    // global-memory histogram: "groups" only describe how features are sharded across ranks (64 per shard unit)
    for (uint32_t f = 0; f < ell.n_features; f += 64) {
      GroupDesc g{};
      g.f_begin = f;
      g.f_end = std::min(ell.n_features, f + 64);
      g.bin_begin = ell.cut_ptrs[g.f_begin];
      g.bin_end = ell.cut_ptrs[g.f_end];
      groups.push_back(g);
    }
    return groups;
  }
  // uint16 pages take the bank-mapped kernels too while one 32-feature block of 16-byte accumulators fits
  // (<= ~440 symbols per feature: 256 bins + the null symbol is the common case); wider ones the generic kernel
  const bool banked16 = ell.bin_bytes == 2 && HistImpl() == 0 &&
                        (size_t)4 * (ell.max_feature_bins + 1u) * 32 * 4 <= optin - 1024 &&
                        ell.row_bytes >= ((ell.n_features + 31u) & ~31u) * 2u;
  if (ell.bin_bytes == 1 || banked16) {
    const size_t budget = optin - 1024;
    uint32_t f = 0;
    while (f < ell.n_features) {
      GroupDesc g{};
      g.f_begin = f;
      uint32_t blocks = 0, max_sym = 0, f_end = f;
      while (f_end < ell.n_features) {
        uint32_t b_end = std::min(ell.n_features, f_end + 32);
        uint32_t ms = max_sym;
        for (uint32_t k = f_end; k < b_end; ++k) ms = std::max(ms, ell.cut_ptrs[k + 1] - ell.cut_ptrs[k]);
        // pages with missing values: the null symbol (255) gets an accumulator slot of its own, so the kernels need
        // no predicate for it; the slot is dropped at flush time (missing = parent - sum(bins) downstream anyway)
        if (ell.null_symbol != 0xFFFFFFFFu) ms = std::max(ms, ell.null_symbol + 1u);
        if ((size_t)4 * (blocks + 1) * ms * 32 * 4 + RingSmemBytes(blocks + 1) > budget) break;
        // default kernels: one block per group, whatever the symbol count -- equal work per grid row (a group of 6
        // blocks next to one of 2 leaves CTAs idle) and every pair of blocks is eligible for the 64-feature kernel
        if (HistImpl() == 0 && blocks == 1) break;
        max_sym = ms;
        ++blocks;
        f_end = b_end;
      }
      B200_CHECK(blocks > 0, "a 32-feature block does not fit in shared memory");
      g.f_end = f_end;
      g.bin_begin = ell.cut_ptrs[f];
      g.bin_end = ell.cut_ptrs[f_end];
      g.byte_begin = f * ell.bin_bytes;
      g.n_words = (f_end - f + 3) / 4;
      g.n_blocks = blocks;
      g.max_sym = max_sym;
      groups.push_back(g);
      f = f_end;
    }
    return groups;
  }
  const size_t fixed = (size_t)kHistTile * 20 + kMaxGroupFeatures * 4 + 256;
  B200_CHECK(optin > fixed + 4096, "device has too little shared memory");
  const uint32_t max_group_bins = (uint32_t)((optin - fixed) / 16);
  const uint32_t per_word = 4 / ell.bin_bytes;
  uint32_t f = 0;
  while (f < ell.n_features) {
    uint32_t f_end = f, bins = 0;
    while (f_end < ell.n_features && f_end - f < kMaxGroupFeatures) {
      uint32_t fb = ell.cut_ptrs[f_end + 1] - ell.cut_ptrs[f_end];
      if (bins + fb > max_group_bins) break;
      bins += fb;
      ++f_end;
    }
    B200_CHECK(f_end > f, "a single feature has more bins than fit in shared memory");
    if (f_end < ell.n_features) {  // keep the next group word-aligned
      uint32_t aligned = f + ((f_end - f) / per_word) * per_word;
      B200_CHECK(aligned > f, "cannot form a word-aligned feature group");
      f_end = aligned;
    }
    GroupDesc g{};
    g.f_begin = f;
    g.f_end = f_end;
    g.bin_begin = ell.cut_ptrs[f];
    g.bin_end = ell.cut_ptrs[f_end];
    g.byte_begin = f * ell.bin_bytes;
    g.n_words = (f_end - f + per_word - 1) / per_word;
    g.n_blocks = 0;
    g.max_sym = 0;
    groups.push_back(g);
    f = f_end;
  }
  return groups;
}

// 64-bit add into a shared accumulator whose low / high 32-bit words are `stride` words apart.
__device__ __forceinline__ void SmemAdd64(uint32_t* lo, uint32_t* hi, long long v) {
  const uint32_t vlo = (uint32_t)(unsigned long long)v;
  const uint32_t vhi = (uint32_t)((unsigned long long)v >> 32);
  const uint32_t old = atomicAdd(lo, vlo);
  const uint32_t hi_add = vhi + ((uint32_t)(old + vlo) < old ? 1u : 0u);
  if (hi_add != 0u) atomicAdd(hi, hi_add);
}

// ---------------------------------------------------------------------------------------------
// Bank-mapped kernel (uint8 symbols)
//
// Shared layout: word index = ((blk * S + sym) * 4 + c) * 32 + lane, c in {g.lo, g.hi, h.lo, h.hi}.
// Bank = lane = (feature mod 32); the four words of one accumulator sit at +0/+128/+256/+384 bytes, so
// they are immediate offsets of one address register.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kRowsPerPass = 8;  // rows a warp handles per pass: 2 quads of 4 rows

struct PassData {       // what one lane holds for one pass (2 rows)
  uint32_t word[2];     // 4 symbols of the lane's 32-bit column word, per row
  uint32_t glo[2], ghi[2], hlo[2], hhi[2];
};

#ifdef B200_EMU
// CPU emulator (tests/emu): the same operations on the emulated shared window; `addr` is an offset into it
__device__ __forceinline__ uint32_t AtomAddShared(uint32_t addr, uint32_t v) { return atomicAdd(emu::SharedWord(addr), v); }
__device__ __forceinline__ uint32_t AtomAddShared256(uint32_t addr, uint32_t v) { return atomicAdd(emu::SharedWord(addr + 256), v); }
__device__ __forceinline__ uint32_t EmuCarryHi(uint32_t old, uint32_t vlo, uint32_t vhi) {
  return vhi + (uint32_t)(((unsigned long long)old + vlo) >> 32);
}
__device__ __forceinline__ void RedAddHi128(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  atomicAdd(emu::SharedWord(addr + 128), EmuCarryHi(old, vlo, vhi));
}
__device__ __forceinline__ void RedAdd384(uint32_t addr, uint32_t v) { atomicAdd(emu::SharedWord(addr + 384), v); }
__device__ __forceinline__ void RedAddHi384(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  atomicAdd(emu::SharedWord(addr + 384), EmuCarryHi(old, vlo, vhi));
}
#else
__device__ __forceinline__ uint32_t AtomAddShared(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t AtomAddShared256(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1+256], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
// hi += vhi + carry(old + vlo), no return value needed
__device__ __forceinline__ void RedAddHi128(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  uint32_t hi;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(hi) : "r"(old), "r"(vlo), "r"(vhi));
  asm volatile("red.shared.add.u32 [%0+128], %1;" ::"r"(addr), "r"(hi) : "memory");
}
__device__ __forceinline__ void RedAdd384(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+384], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void RedAddHi384(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  uint32_t hi;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(hi) : "r"(old), "r"(vlo), "r"(vhi));
  asm volatile("red.shared.add.u32 [%0+384], %1;" ::"r"(addr), "r"(hi) : "memory");
}
#endif  // B200_EMU

// kChecked = false: every block of the group is a full 32 features and the page has no null symbol, so
// the inner loop carries no predicate at all.
template <bool kChecked>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernelBanked(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
                 const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
                 const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
                 const GroupDesc* __restrict__ groups, longlong2* __restrict__ hist,
                 const QuantDev* __restrict__ qd) {
  B200_DYNAMIC_SHARED_ALIGNED(uint32_t, sh, 16);
  const GroupDesc g = groups[blockIdx.y];
  const uint32_t S = g.max_sym, B = g.n_blocks;
  const uint32_t slots = B * S * 32;  // (block, symbol, lane) accumulators
  const uint32_t nf = g.f_end - g.f_begin;

  const uint32_t n_tiles = ts->build_tiles;
  if (n_tiles == 0) return;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  if (t0 >= t1) return;
  const uint32_t n_build = (uint32_t)ts->n_build;
  const bool h_lo_zero = qd != nullptr && qd->h_lo_or == 0u;  // uniform over the grid

  for (uint32_t i = threadIdx.x; i < 4 * slots; i += kHistThreads) sh[i] = 0u;
  __syncthreads();

  auto flush = [&](int32_t nidx) {
    __syncthreads();
    longlong2* dst = hist + (size_t)nidx * ell.total_bins;
    for (uint32_t p = threadIdx.x; p < slots; p += kHistThreads) {
      const uint32_t ln = p & 31u, rest = p >> 5;  // rest = blk * S + sym
      uint32_t* a = sh + rest * 128 + ln;
      const uint32_t glo = a[0], ghi = a[32], hlo = a[64], hhi = a[96];
      if (glo | ghi | hlo | hhi) {
        const uint32_t f = g.f_begin + (rest / S) * 32 + ln;  // non-zero implies a valid (feature, bin)
        longlong2* d = dst + ell.cut_ptrs[f] + (rest % S);
        if (glo | ghi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->x), ((unsigned long long)ghi << 32) | glo);
        if (hlo | hhi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->y), ((unsigned long long)hhi << 32) | hlo);
        a[0] = 0u; a[32] = 0u; a[64] = 0u; a[96] = 0u;
      }
    }
    __syncthreads();
  };

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t q = lane >> 3, w = lane & 7u;  // row of the quad, 32-bit word of the block
  constexpr uint32_t kWarps = kHistThreads / 32;
  constexpr uint32_t kStride = kWarps * kRowsPerPass;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh);
  // step k of a lane uses byte (k + q) mod 4 of its word: 32 lanes -> 32 distinct features -> 32 banks
  uint32_t sel[4], lane_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    sel[k] = 0x4440u | jj;               // PRMT selector: byte jj -> byte 0, zeros elsewhere
    lane_off[k] = (4u * w + jj) * 4u;    // byte offset of the lane slot
  }

  uint32_t t = t0;
  while (t < t1) {
    const uint32_t seg = tile_seg[t];
    const BuildEntry e = build[seg];
    const uint32_t seg_tile_end = (seg + 1 < n_build) ? build[seg + 1].tile_begin : n_tiles;
    const uint32_t te = min(t1, seg_tile_end);
    uint32_t lo = (t - e.tile_begin) * kHistTile;
    uint32_t hi = min(e.count, (te - e.tile_begin) * kHistTile);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin;

    for (uint32_t blk = 0; blk < B; ++blk) {
      const uint32_t blk_base = sh_base + blk * S * 512u;
      const uint8_t* __restrict__ col = ell.bins + g.byte_begin + blk * 32u + 4u * w;
      const bool in_row = g.byte_begin + blk * 32u + 4u * w < ell.row_bytes;  // partial last block
      bool fok[4];
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) fok[k] = blk * 32u + 4u * w + ((k + q) & 3u) < nf;

      auto load_r = [&](uint32_t jb, uint32_t (&r)[2], bool (&v)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t j = jb + u * 4 + q;
          v[u] = j < hi;
          r[u] = __ldg(ridx + (v[u] ? j : lo));  // rows past the end re-read a valid row and add zero
        }
      };
      auto load_d = [&](const uint32_t (&r)[2], const bool (&v)[2], PassData& d) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const longlong2 gp = __ldg(gq + r[u]);
          const uint32_t wd = (!kChecked || in_row)
                                  ? __ldg(reinterpret_cast<const uint32_t*>(col + (size_t)r[u] * ell.row_bytes))
                                  : 0u;
          d.word[u] = wd;
          d.glo[u] = v[u] ? (uint32_t)(unsigned long long)gp.x : 0u;
          d.ghi[u] = v[u] ? (uint32_t)((unsigned long long)gp.x >> 32) : 0u;
          d.hlo[u] = v[u] ? (uint32_t)(unsigned long long)gp.y : 0u;
          d.hhi[u] = v[u] ? (uint32_t)((unsigned long long)gp.y >> 32) : 0u;
        }
      };
      auto process = [&](const PassData& d, auto hz) {
        constexpr bool kHLoZero = decltype(hz)::value;  // every hq has a zero low word: no h.lo atomic
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t addr[4], og[4], oh[4], vgl[4], vgh[4], vhl[4], vhh[4];
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t sym = __byte_perm(d.word[u], 0u, sel[k]);
            vgl[k] = d.glo[u]; vgh[k] = d.ghi[u]; vhl[k] = d.hlo[u]; vhh[k] = d.hhi[u];
            if (kChecked) {
              const bool ok = fok[k] && sym < S;  // partial block / null symbol: add zero to slot 0
              addr[k] = blk_base + (ok ? sym : 0u) * 512u + lane_off[k];
              if (!ok) { vgl[k] = 0u; vgh[k] = 0u; vhl[k] = 0u; vhh[k] = 0u; }
            } else {
              addr[k] = blk_base + sym * 512u + lane_off[k];
            }
          }
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            og[k] = AtomAddShared(addr[k], vgl[k]);
            if (!kHLoZero) oh[k] = AtomAddShared256(addr[k], vhl[k]);
          }
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            RedAddHi128(addr[k], og[k], vgl[k], vgh[k]);
            if (kHLoZero) RedAdd384(addr[k], vhh[k]);
            else RedAddHi384(addr[k], oh[k], vhl[k], vhh[k]);
          }
        }
      };

      auto run = [&](auto hz) {
        // software pipeline: data of pass i+1 and row ids of pass i+2 are in flight while pass i accumulates
        uint32_t jb = lo + warp * kRowsPerPass;
        if (jb < hi) {
          uint32_t r1[2], r2[2] = {0u, 0u};
          bool v1[2], v2[2] = {false, false};
          PassData d1, d2;
          load_r(jb, r1, v1);
          uint32_t jb2 = jb + kStride;
          bool have2 = jb2 < hi;
          if (have2) load_r(jb2, r2, v2);
          load_d(r1, v1, d1);
          while (true) {
            const uint32_t jb3 = jb2 + kStride;
            const bool have3 = have2 && jb3 < hi;
            uint32_t r3[2] = {0u, 0u};
            bool v3[2] = {false, false};
            if (have2) load_d(r2, v2, d2);
            if (have3) load_r(jb3, r3, v3);
            process(d1, hz);
            if (!have2) break;
            d1 = d2;
            r2[0] = r3[0]; r2[1] = r3[1]; v2[0] = v3[0]; v2[1] = v3[1];
            jb2 = jb3;
            have2 = have3;
          }
        }
      };
      if (h_lo_zero) run(std::true_type{}); else run(std::false_type{});
    }
    flush((int32_t)e.slot);
    t = te;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast bank-mapped kernel for the common case (every block is a full 32 features, no null symbol):
// the same accumulation as HistKernelBanked<false>, hand-scheduled to keep the instruction stream lean --
// full passes only in the main loop (no validity selects), ping-pong register buffers (no copies),
// pointer-increment addressing, block base folded into the lane offsets.  A warp's last partial pass goes
// through a small checked tail.
// ---------------------------------------------------------------------------------------------
// The 4 feature symbols a lane holds of one row: one register for uint8 pages, two for uint16 pages (pages with
// 257..~440 symbols per feature, e.g. 256 bins + the null symbol of a matrix with missing values).
template <int kSymBytes>
struct Sym4;
template <>
struct Sym4<1> {
  using T = uint32_t;
  static __device__ __forceinline__ T Load(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ uint32_t Get(T w, uint32_t sel) { return __byte_perm(w, 0u, sel); }
  static __device__ __forceinline__ uint32_t Sel(uint32_t jj) { return 0x4440u | jj; }
};
template <>
struct Sym4<2> {
  using T = uint2;
  static __device__ __forceinline__ T Load(const uint8_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
  static __device__ __forceinline__ uint32_t Get(T w, uint32_t sel) { return __byte_perm(w.x, w.y, sel) & 0xFFFFu; }
  static __device__ __forceinline__ uint32_t Sel(uint32_t jj) { return (2u * jj) | ((2u * jj + 1u) << 4); }
};

template <int kSymBytes>
struct FastData {
  typename Sym4<kSymBytes>::T word[2];
  uint32_t glo[2], ghi[2], hlo[2], hhi[2];
};

template <bool kHLoZero, int kSymBytes>
__device__ __forceinline__ void FastProcess(const FastData<kSymBytes>& d, const uint32_t (&sel)[4],
                                            const uint32_t (&off)[4]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t addr[4], og[4], oh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) addr[k] = Sym4<kSymBytes>::Get(d.word[u], sel[k]) * 512u + off[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      og[k] = AtomAddShared(addr[k], d.glo[u]);
      if (!kHLoZero) oh[k] = AtomAddShared256(addr[k], d.hlo[u]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      RedAddHi128(addr[k], og[k], d.glo[u], d.ghi[u]);
      if (kHLoZero) RedAdd384(addr[k], d.hhi[u]);
      else RedAddHi384(addr[k], oh[k], d.hlo[u], d.hhi[u]);
    }
  }
}

template <int kSymBytes>
__device__ __forceinline__ void FastLoad(FastData<kSymBytes>& d, const uint32_t (&r)[2],
                                         const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                         uint32_t row_bytes) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const longlong2 gp = __ldg(gq + r[u]);
    d.word[u] = Sym4<kSymBytes>::Load(col + (size_t)r[u] * row_bytes);
    d.glo[u] = (uint32_t)(unsigned long long)gp.x;
    d.ghi[u] = (uint32_t)((unsigned long long)gp.x >> 32);
    d.hlo[u] = (uint32_t)(unsigned long long)gp.y;
    d.hhi[u] = (uint32_t)((unsigned long long)gp.y >> 32);
  }
}

template <bool kHLoZero, int kSymBytes>
__device__ __forceinline__ void FastRows(const uint32_t* __restrict__ ridx, uint32_t lo, uint32_t hi,
                                         const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                         uint32_t row_bytes, uint32_t warp, uint32_t q,
                                         const uint32_t (&sel)[4], const uint32_t (&off)[4]) {
  constexpr uint32_t kWarps = kHistThreads / 32;
  constexpr uint32_t kStride = kWarps * kRowsPerPass;
  const uint32_t first = lo + warp * kRowsPerPass;
  if (first >= hi) return;
  // passes of this warp whose 8 rows all exist
  const uint32_t n_full = (hi >= first + kRowsPerPass) ? (hi - first - kRowsPerPass) / kStride + 1u : 0u;
  const uint32_t* __restrict__ rp = ridx + first + q;  // row id of (u = 0); u = 1 is +4
  if (n_full > 0) {
    // clamp so that prefetches past the end re-read the last full pass (discarded) instead of branching
    const uint32_t last = n_full - 1;
    auto rid = [&](uint32_t i, uint32_t (&r)[2]) {
      const uint32_t* p = rp + (size_t)min(i, last) * kStride;
      r[0] = __ldg(p);
      r[1] = __ldg(p + 4);
    };
    uint32_t ra[2], rb[2];
    FastData<kSymBytes> da, db;
    rid(0, ra);
    rid(1, rb);
    FastLoad(da, ra, gq, col, row_bytes);
    rid(2, ra);
    uint32_t i = 0;
    for (; i + 1 < n_full; i += 2) {
      FastLoad(db, rb, gq, col, row_bytes);  // pass i+1
      rid(i + 3, rb);
      FastProcess<kHLoZero>(da, sel, off);   // pass i
      FastLoad(da, ra, gq, col, row_bytes);  // pass i+2
      rid(i + 4, ra);
      FastProcess<kHLoZero>(db, sel, off);   // pass i+1
    }
    if (i < n_full) FastProcess<kHLoZero>(da, sel, off);  // odd count: pass n_full-1 is in da
  }
  // partial last pass (fewer than 8 rows): zero the gradients of the missing rows
  const uint32_t jb = first + n_full * kStride;
  if (jb < hi) {
    uint32_t r[2];
    FastData<kSymBytes> d;
#pragma unroll
    for (int u = 0; u < 2; ++u) r[u] = __ldg(ridx + min(jb + u * 4 + q, hi - 1));
    FastLoad(d, r, gq, col, row_bytes);
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (jb + u * 4 + q >= hi) { d.glo[u] = 0u; d.ghi[u] = 0u; d.hlo[u] = 0u; d.hhi[u] = 0u; }
    FastProcess<false>(d, sel, off);
  }
}

// h.lo-skip variant with THREE register buffers: the data of passes i+1 and i+2 are in flight while pass i
// accumulates (global-load latency, not issue or atomics, limits the two-buffer version: ncu shows the
// long-scoreboard stall dominating at 46 % issue utilisation).  Only 8 registers per pass are needed here.
template <int kSymBytes>
struct FastDataZ {
  typename Sym4<kSymBytes>::T word[2];
  uint32_t glo[2], ghi[2], hhi[2];
};
template <int kSymBytes>
__device__ __forceinline__ void FastLoadZ(FastDataZ<kSymBytes>& d, const uint32_t (&r)[2],
                                          const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                          uint32_t row_bytes) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const longlong2 gp = __ldg(gq + r[u]);
    d.word[u] = Sym4<kSymBytes>::Load(col + (size_t)r[u] * row_bytes);
    d.glo[u] = (uint32_t)(unsigned long long)gp.x;
    d.ghi[u] = (uint32_t)((unsigned long long)gp.x >> 32);
    d.hhi[u] = (uint32_t)((unsigned long long)gp.y >> 32);
  }
}
template <int kSymBytes>
__device__ __forceinline__ void FastProcessZ(const FastDataZ<kSymBytes>& d, const uint32_t (&sel)[4],
                                             const uint32_t (&off)[4]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t addr[4], og[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) addr[k] = Sym4<kSymBytes>::Get(d.word[u], sel[k]) * 512u + off[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) og[k] = AtomAddShared(addr[k], d.glo[u]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      RedAddHi128(addr[k], og[k], d.glo[u], d.ghi[u]);
      RedAdd384(addr[k], d.hhi[u]);
    }
  }
}

template <int kSymBytes>
__device__ __forceinline__ void FastRowsZ3(const uint32_t* __restrict__ ridx, uint32_t lo, uint32_t hi,
                                           const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                           uint32_t row_bytes, uint32_t warp, uint32_t q,
                                           const uint32_t (&sel)[4], const uint32_t (&off)[4]) {
  constexpr uint32_t kWarps = kHistThreads / 32;
  constexpr uint32_t kStride = kWarps * kRowsPerPass;
  const uint32_t first = lo + warp * kRowsPerPass;
  if (first >= hi) return;
  const uint32_t n_full = (hi >= first + kRowsPerPass) ? (hi - first - kRowsPerPass) / kStride + 1u : 0u;
  const uint32_t* __restrict__ rp = ridx + first + q;
  if (n_full > 0) {
    const uint32_t last = n_full - 1;
    auto rid = [&](uint32_t i, uint32_t (&r)[2]) {
      const uint32_t* p = rp + (size_t)min(i, last) * kStride;
      r[0] = __ldg(p);
      r[1] = __ldg(p + 4);
    };
    uint32_t r0[2], r1[2], r2[2];
    FastDataZ<kSymBytes> d0, d1, d2;
    rid(0, r0); rid(1, r1); rid(2, r2);
    FastLoadZ(d0, r0, gq, col, row_bytes); rid(3, r0);
    FastLoadZ(d1, r1, gq, col, row_bytes); rid(4, r1);
    uint32_t i = 0;
    for (; i + 2 < n_full; i += 3) {
      FastLoadZ(d2, r2, gq, col, row_bytes); rid(i + 5, r2);  // pass i+2
      FastProcessZ(d0, sel, off);                             // pass i
      FastLoadZ(d0, r0, gq, col, row_bytes); rid(i + 6, r0);  // pass i+3
      FastProcessZ(d1, sel, off);                             // pass i+1
      FastLoadZ(d1, r1, gq, col, row_bytes); rid(i + 7, r1);  // pass i+4
      FastProcessZ(d2, sel, off);                             // pass i+2
    }
    if (i < n_full) FastProcessZ(d0, sel, off);
    if (i + 1 < n_full) FastProcessZ(d1, sel, off);
  }
  const uint32_t jb = first + n_full * kStride;
  if (jb < hi) {  // partial last pass
    uint32_t r[2];
    FastData<kSymBytes> d;
#pragma unroll
    for (int u = 0; u < 2; ++u) r[u] = __ldg(ridx + min(jb + u * 4 + q, hi - 1));
    FastLoad(d, r, gq, col, row_bytes);
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (jb + u * 4 + q >= hi) { d.glo[u] = 0u; d.ghi[u] = 0u; d.hlo[u] = 0u; d.hhi[u] = 0u; }
    FastProcess<false>(d, sel, off);
  }
}

// One group (B blocks of 32 features, 16-byte accumulators) over the CTA's tile range [t0, t1).
template <int kSymBytes, bool kPaged = false>
__device__ __forceinline__ void BankedFastGroup(const EllView& ell, const longlong2* __restrict__ gq,
                                                const uint32_t* __restrict__ ridx0, const uint32_t* __restrict__ ridx1,
                                                const BuildEntry* __restrict__ build,
                                                const uint32_t* __restrict__ tile_seg, uint32_t n_tiles,
                                                uint32_t n_build, const GroupDesc& g, longlong2* __restrict__ hist,
                                                bool h_lo_zero, uint32_t* sh, uint32_t t0, uint32_t t1) {
  const uint32_t S = g.max_sym, B = g.n_blocks;
  const uint32_t slots = B * S * 32;

  for (uint32_t i = threadIdx.x; i < 4 * slots; i += kHistThreads) sh[i] = 0u;
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t q = lane >> 3, w = lane & 7u;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t sel[4], lane_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    sel[k] = Sym4<kSymBytes>::Sel(jj);
    lane_off[k] = (4u * w + jj) * 4u;
  }

  uint32_t t = t0;
  while (t < t1) {
    const uint32_t seg = tile_seg[t];
    const BuildEntry e = build[seg];
    const uint32_t seg_tile_end = (seg + 1 < n_build) ? build[seg + 1].tile_begin : n_tiles;
    const uint32_t te = min(t1, seg_tile_end);
    uint32_t lo = (t - e.tile_begin) * kHistTile;
    uint32_t hi = min(e.count, (te - e.tile_begin) * kHistTile);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin;
    if constexpr (kPaged) ClipToPage(ridx, lo, hi, ell.page_lo, ell.page_hi);  // external memory: this page's rows
    for (uint32_t blk = 0; blk < B; ++blk) {
      uint32_t off[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) off[k] = sh_base + blk * S * 512u + lane_off[k];
      const uint8_t* __restrict__ col = ell.bins + g.byte_begin + (blk * 32u + 4u * w) * kSymBytes;
      if (h_lo_zero) FastRowsZ3<kSymBytes>(ridx, lo, hi, gq, col, ell.row_bytes, warp, q, sel, off);
      else FastRows<false, kSymBytes>(ridx, lo, hi, gq, col, ell.row_bytes, warp, q, sel, off);
    }
    // flush this CTA's share of the node (RED.64 on non-zero accumulators) and clear
    __syncthreads();
    longlong2* dst = hist + (size_t)e.slot * ell.total_bins;
    for (uint32_t p = threadIdx.x; p < slots; p += kHistThreads) {
      const uint32_t ln = p & 31u, rest = p >> 5;
      uint32_t* a = sh + rest * 128 + ln;
      const uint32_t glo = a[0], ghi = a[32], hlo = a[64], hhi = a[96];
      if (glo | ghi | hlo | hhi) {
        const uint32_t f = g.f_begin + (rest / S) * 32 + ln;
        // dropped: features past the end (row padding, symbol 0) and symbols past the feature's bins (the null slot)
        if (f < ell.n_features && ell.cut_ptrs[f] + (rest % S) < ell.cut_ptrs[f + 1]) {
          longlong2* d = dst + ell.cut_ptrs[f] + (rest % S);
          if (glo | ghi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->x), ((unsigned long long)ghi << 32) | glo);
          if (hlo | hhi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->y), ((unsigned long long)hhi << 32) | hlo);
        }
        a[0] = 0u; a[32] = 0u; a[64] = 0u; a[96] = 0u;
      }
    }
    __syncthreads();
    t = te;
  }
}

template <int kSymBytes, bool kPaged = false>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernelBankedFast(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
                     const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
                     const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
                     const GroupDesc* __restrict__ groups, longlong2* __restrict__ hist,
                     const QuantDev* __restrict__ qd) {
  B200_DYNAMIC_SHARED_ALIGNED(uint32_t, sh, 16);
  const uint32_t n_tiles = ts->build_tiles;
  if (n_tiles == 0) return;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  if (t0 >= t1) return;
  const GroupDesc g = groups[blockIdx.y];
  BankedFastGroup<kSymBytes, kPaged>(ell, gq, ridx0, ridx1, build, tile_seg, n_tiles, (uint32_t)ts->n_build, g, hist,
                                     qd != nullptr && qd->h_lo_or == 0u, sh, t0, t1);
}

// ---------------------------------------------------------------------------------------------
// Pair kernel: TWO adjacent 32-feature blocks (64 features) per CTA pass, 12-byte accumulators.
//
// ncu on HistKernelBankedFast shows the L1TEX data pipe (l1tex__data_pipe_lsu_wavefronts) at 83-86 % of peak:
// 57 % of its wavefronts are the shared atomics, 43 % are global-load wavefronts -- one per gathered row for the
// symbols, one per row for the gradient pair, per 32-feature block.  When every quantised hessian has a zero
// low word (h.lo is never accumulated, e.g. reg:squarederror) an accumulator is 12 bytes, so 64 features x 256
// symbols fit one CTA (192 KiB): a lane loads 8 symbols of a row with ONE 64-bit load and the row's gradient
// pair serves 64 features, which halves the global-load wavefronts per update.
//
// Shared layout: byte address = sym * 768 + half * 384 + c * 128 + bank * 4, c in {g.lo, g.hi, h.hi};
// lane (q, w) holds the row's bytes 8w .. 8w+7: word `half` carries features 8w + 4*half + b, accumulated in
// bank 4w + b -- for a fixed step the 32 lanes hit 32 different banks, as in the 32-feature kernel.
// If the hessians turn out NOT to have zero low words (the launch is chosen from the previous tree's flag), the
// CTA falls back to the 16-byte layout and runs its two blocks one after the other.
// ---------------------------------------------------------------------------------------------
#ifdef B200_EMU
template <int kOff>
__device__ __forceinline__ uint32_t AtomAddSharedOff(uint32_t addr, uint32_t v) { return atomicAdd(emu::SharedWord(addr + kOff), v); }
template <int kOff>
__device__ __forceinline__ void RedAddOff(uint32_t addr, uint32_t v) { atomicAdd(emu::SharedWord(addr + kOff), v); }
template <int kOff>
__device__ __forceinline__ void RedAddHiOff(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  atomicAdd(emu::SharedWord(addr + kOff), EmuCarryHi(old, vlo, vhi));
}
#else
template <int kOff>
__device__ __forceinline__ uint32_t AtomAddSharedOff(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1+%3], %2;" : "=r"(old) : "r"(addr), "r"(v), "n"(kOff) : "memory");
  return old;
}
template <int kOff>
__device__ __forceinline__ void RedAddOff(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(kOff) : "memory");
}
template <int kOff>
__device__ __forceinline__ void RedAddHiOff(uint32_t addr, uint32_t old, uint32_t vlo, uint32_t vhi) {
  uint32_t hi;
  asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(hi) : "r"(old), "r"(vlo), "r"(vhi));
  asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(hi), "n"(kOff) : "memory");
}
#endif  // B200_EMU

template <int kSymBytes>
struct FastDataP {
  typename Sym4<kSymBytes>::T word[2][2];  // [row][half]: 2 x 4 symbols of a row
  uint32_t glo[2], ghi[2], hhi[2];
};
__device__ __forceinline__ void LoadSym8(const uint8_t* p, uint32_t (&w)[2]) {
  const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
  w[0] = v.x;
  w[1] = v.y;
}
__device__ __forceinline__ void LoadSym8(const uint8_t* p, uint2 (&w)[2]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  w[0] = make_uint2(v.x, v.y);
  w[1] = make_uint2(v.z, v.w);
}
template <int kSymBytes>
__device__ __forceinline__ void FastLoadP(FastDataP<kSymBytes>& d, const uint32_t (&r)[2],
                                          const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                          uint32_t row_bytes) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const longlong2 gp = __ldg(gq + r[u]);
    LoadSym8(col + (size_t)r[u] * row_bytes, d.word[u]);
    d.glo[u] = (uint32_t)(unsigned long long)gp.x;
    d.ghi[u] = (uint32_t)((unsigned long long)gp.x >> 32);
    d.hhi[u] = (uint32_t)((unsigned long long)gp.y >> 32);
  }
}
template <int kHalf, int kSymBytes>
__device__ __forceinline__ void FastProcessPHalf(typename Sym4<kSymBytes>::T word, uint32_t glo, uint32_t ghi,
                                                 uint32_t hhi, const uint32_t (&sel)[4], const uint32_t (&off)[4]) {
  uint32_t addr[4], og[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) addr[k] = Sym4<kSymBytes>::Get(word, sel[k]) * 768u + off[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) og[k] = AtomAddSharedOff<kHalf * 384>(addr[k], glo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    RedAddHiOff<kHalf * 384 + 128>(addr[k], og[k], glo, ghi);
    RedAddOff<kHalf * 384 + 256>(addr[k], hhi);
  }
}
template <int kSymBytes>
__device__ __forceinline__ void FastProcessP(const FastDataP<kSymBytes>& d, const uint32_t (&sel)[4],
                                             const uint32_t (&off)[4]) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    FastProcessPHalf<0, kSymBytes>(d.word[u][0], d.glo[u], d.ghi[u], d.hhi[u], sel, off);
    FastProcessPHalf<1, kSymBytes>(d.word[u][1], d.glo[u], d.ghi[u], d.hhi[u], sel, off);
  }
}

template <int kBufs, int kSymBytes>
__device__ __forceinline__ void FastRowsP(const uint32_t* __restrict__ ridx, uint32_t lo, uint32_t hi,
                                          const longlong2* __restrict__ gq, const uint8_t* __restrict__ col,
                                          uint32_t row_bytes, uint32_t warp, uint32_t q,
                                          const uint32_t (&sel)[4], const uint32_t (&off)[4]) {
  constexpr uint32_t kWarps = kHistThreads / 32;
  constexpr uint32_t kStride = kWarps * kRowsPerPass;
  const uint32_t first = lo + warp * kRowsPerPass;
  if (first >= hi) return;
  const uint32_t n_full = (hi >= first + kRowsPerPass) ? (hi - first - kRowsPerPass) / kStride + 1u : 0u;
  const uint32_t* __restrict__ rp = ridx + first + q;
  if (n_full > 0) {
    const uint32_t last = n_full - 1;
    auto rid = [&](uint32_t i, uint32_t (&r)[2]) {
      const uint32_t* p = rp + (size_t)min(i, last) * kStride;
      r[0] = __ldg(p);
      r[1] = __ldg(p + 4);
    };
    if (kBufs == 3) {
      uint32_t r0[2], r1[2], r2[2];
      FastDataP<kSymBytes> d0, d1, d2;
      rid(0, r0); rid(1, r1); rid(2, r2);
      FastLoadP(d0, r0, gq, col, row_bytes); rid(3, r0);
      FastLoadP(d1, r1, gq, col, row_bytes); rid(4, r1);
      uint32_t i = 0;
      for (; i + 2 < n_full; i += 3) {
        FastLoadP(d2, r2, gq, col, row_bytes); rid(i + 5, r2);  // pass i+2
        FastProcessP(d0, sel, off);                             // pass i
        FastLoadP(d0, r0, gq, col, row_bytes); rid(i + 6, r0);  // pass i+3
        FastProcessP(d1, sel, off);                             // pass i+1
        FastLoadP(d1, r1, gq, col, row_bytes); rid(i + 7, r1);  // pass i+4
        FastProcessP(d2, sel, off);                             // pass i+2
      }
      if (i < n_full) FastProcessP(d0, sel, off);
      if (i + 1 < n_full) FastProcessP(d1, sel, off);
    } else {
      uint32_t ra[2], rb[2];
      FastDataP<kSymBytes> da, db;
      rid(0, ra);
      rid(1, rb);
      FastLoadP(da, ra, gq, col, row_bytes);
      rid(2, ra);
      uint32_t i = 0;
      for (; i + 1 < n_full; i += 2) {
        FastLoadP(db, rb, gq, col, row_bytes); rid(i + 3, rb);  // pass i+1
        FastProcessP(da, sel, off);                             // pass i
        FastLoadP(da, ra, gq, col, row_bytes); rid(i + 4, ra);  // pass i+2
        FastProcessP(db, sel, off);                             // pass i+1
      }
      if (i < n_full) FastProcessP(da, sel, off);
    }
  }
  const uint32_t jb = first + n_full * kStride;
  if (jb < hi) {  // partial last pass: rows past the end add zero
    uint32_t r[2];
    FastDataP<kSymBytes> d;
#pragma unroll
    for (int u = 0; u < 2; ++u) r[u] = __ldg(ridx + min(jb + u * 4 + q, hi - 1));
    FastLoadP(d, r, gq, col, row_bytes);
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (jb + u * 4 + q >= hi) { d.glo[u] = 0u; d.ghi[u] = 0u; d.hhi[u] = 0u; }
    FastProcessP(d, sel, off);
  }
}

template <int kBufs, int kSymBytes, bool kPaged = false>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernelPair(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
               const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
               const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
               const GroupDesc* __restrict__ groups, uint32_t n_groups, longlong2* __restrict__ hist,
               const QuantDev* __restrict__ qd) {
  B200_DYNAMIC_SHARED_ALIGNED(uint32_t, sh, 16);
  const uint32_t n_tiles = ts->build_tiles;
  if (n_tiles == 0) return;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  if (t0 >= t1) return;
  const uint32_t n_build = (uint32_t)ts->n_build;
  const bool h_lo_zero = qd != nullptr && qd->h_lo_or == 0u;
  const uint32_t gi = blockIdx.y * 2u;
  const bool two = gi + 1u < n_groups;
  const GroupDesc g0 = groups[gi];
  if (!two || !h_lo_zero) {  // odd block out, or hessians with low words: 16-byte layout, block after block
    BankedFastGroup<kSymBytes, kPaged>(ell, gq, ridx0, ridx1, build, tile_seg, n_tiles, n_build, g0, hist, h_lo_zero, sh, t0, t1);
    if (two) {
      const GroupDesc g1 = groups[gi + 1u];
      BankedFastGroup<kSymBytes, kPaged>(ell, gq, ridx0, ridx1, build, tile_seg, n_tiles, n_build, g1, hist, h_lo_zero, sh,
                                         t0, t1);
    }
    return;
  }
  const uint32_t S = max(g0.max_sym, groups[gi + 1u].max_sym);
  const uint32_t words = S * 192u;  // 2 halves x 3 components x 32 banks per symbol
  for (uint32_t i = threadIdx.x; i < words; i += kHistThreads) sh[i] = 0u;
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t q = lane >> 3, w = lane & 7u;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t sel[4], off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    sel[k] = Sym4<kSymBytes>::Sel(jj);
    off[k] = sh_base + (4u * w + jj) * 4u;
  }
  const uint8_t* __restrict__ col = ell.bins + g0.byte_begin + 8u * w * kSymBytes;

  uint32_t t = t0;
  while (t < t1) {
    const uint32_t seg = tile_seg[t];
    const BuildEntry e = build[seg];
    const uint32_t seg_tile_end = (seg + 1 < n_build) ? build[seg + 1].tile_begin : n_tiles;
    const uint32_t te = min(t1, seg_tile_end);
    uint32_t lo = (t - e.tile_begin) * kHistTile;
    uint32_t hi = min(e.count, (te - e.tile_begin) * kHistTile);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin;
    if constexpr (kPaged) ClipToPage(ridx, lo, hi, ell.page_lo, ell.page_hi);  // external memory: this page's rows
    FastRowsP<kBufs, kSymBytes>(ridx, lo, hi, gq, col, ell.row_bytes, warp, q, sel, off);
    __syncthreads();
    // flush this CTA's share of the node (RED.64 on non-zero accumulators) and clear
    longlong2* dst = hist + (size_t)e.slot * ell.total_bins;
    for (uint32_t p = threadIdx.x; p < S * 64u; p += kHistThreads) {
      const uint32_t bank = p & 31u, half = (p >> 5) & 1u, sym = p >> 6;
      uint32_t* a = sh + sym * 192u + half * 96u + bank;
      const uint32_t glo = a[0], ghi = a[32], hhi = a[64];
      if (glo | ghi | hhi) {
        const uint32_t f = g0.f_begin + 8u * (bank >> 2) + 4u * half + (bank & 3u);
        // dropped: features past the end (row padding, symbol 0) and symbols past the feature's bins (the null slot)
        if (f < ell.n_features && ell.cut_ptrs[f] + sym < ell.cut_ptrs[f + 1]) {
          longlong2* d = dst + ell.cut_ptrs[f] + sym;
          if (glo | ghi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->x), ((unsigned long long)ghi << 32) | glo);
          if (hhi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->y), (unsigned long long)hhi << 32);
        }
        a[0] = 0u; a[32] = 0u; a[64] = 0u;
      }
    }
    __syncthreads();
    t = te;
  }
}

#ifndef B200_EMU  // async-copy rings (mbarrier / cp.async / TMA): not emulated
// ---------------------------------------------------------------------------------------------
// Staged bank-mapped kernel: the same accumulation, but every warp feeds itself through a private ring of
// kStages shared-memory stages filled by asynchronous copies, so the accumulating instruction stream
// contains no global loads or 64-bit address arithmetic -- only LDS, PRMT/IMAD and the atomics.
//
//   kBulk = true : one TMA 1-D bulk copy (cp.async.bulk.shared::cluster.global, SASS UBLKCP) per gathered
//                  row slice (32*B bytes of symbols, 16-byte aligned by construction of the page layout) and
//                  one per quantised gradient pair, completion counted on a per-(warp, stage) mbarrier.
//   kBulk = false: 16-byte cp.async (LDGSTS) per chunk, completion by commit/wait groups.
// A stage holds the 8 rows of one pass: [8][32*B] symbol bytes then [8] x 16-byte gradient pairs.
// Row ids for the copies are fetched two passes ahead, copies are issued kStages-1 passes ahead.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kStages = kRingStages;

__device__ __forceinline__ void MbarInit(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void BulkCopyG2S(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void CpAsync16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void CpAsyncWait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t LdsU32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 LdsV4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <bool kChecked, bool kBulk>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernelStaged(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
                 const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
                 const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
                 const GroupDesc* __restrict__ groups, longlong2* __restrict__ hist,
                 const QuantDev* __restrict__ qd) {
  B200_DYNAMIC_SHARED_ALIGNED(uint32_t, sh, 16);
  const GroupDesc g = groups[blockIdx.y];
  const uint32_t S = g.max_sym, B = g.n_blocks;
  const uint32_t slots = B * S * 32;
  const uint32_t nf = g.f_end - g.f_begin;
  constexpr uint32_t kWarps = kHistThreads / 32;
  const uint32_t row_b = 32u * B;                   // symbol bytes of one staged row
  const uint32_t stage_bytes = 8u * (row_b + 16u);  // 8 rows + 8 gradient pairs
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh);
  const uint32_t ring_base = sh_base + slots * 16u + warp * kStages * stage_bytes;
  const uint32_t bar_base = sh_base + slots * 16u + kWarps * kStages * stage_bytes + warp * kStages * 8u;

  const uint32_t n_tiles = ts->build_tiles;
  if (n_tiles == 0) return;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  if (t0 >= t1) return;
  const uint32_t n_build = (uint32_t)ts->n_build;
  const bool h_lo_zero = qd != nullptr && qd->h_lo_or == 0u;

  for (uint32_t i = threadIdx.x; i < 4 * slots; i += kHistThreads) sh[i] = 0u;
  if (kBulk && lane == 0) {
#pragma unroll
    for (uint32_t s = 0; s < kStages; ++s) MbarInit(bar_base + 8u * s, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto flush = [&](int32_t nidx) {
    __syncthreads();
    longlong2* dst = hist + (size_t)nidx * ell.total_bins;
    for (uint32_t p = threadIdx.x; p < slots; p += kHistThreads) {
      const uint32_t ln = p & 31u, rest = p >> 5;
      uint32_t* a = sh + rest * 128 + ln;
      const uint32_t glo = a[0], ghi = a[32], hlo = a[64], hhi = a[96];
      if (glo | ghi | hlo | hhi) {
        const uint32_t f = g.f_begin + (rest / S) * 32 + ln;
        longlong2* d = dst + ell.cut_ptrs[f] + (rest % S);
        if (glo | ghi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->x), ((unsigned long long)ghi << 32) | glo);
        if (hlo | hhi) atomicAdd(reinterpret_cast<unsigned long long*>(&d->y), ((unsigned long long)hhi << 32) | hlo);
        a[0] = 0u; a[32] = 0u; a[64] = 0u; a[96] = 0u;
      }
    }
    __syncthreads();
  };

  const uint32_t q = lane >> 3, w = lane & 7u;
  uint32_t sel[4], lane_off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    sel[k] = 0x4440u | jj;
    lane_off[k] = (4u * w + jj) * 4u;
  }
  // copy roles of this lane inside a pass (chunk = 16 bytes): chunks [0, 16B) are symbol chunks
  // (row = c / 2B, part = c % 2B), chunks [16B, 16B+8) the gradient pairs of rows 0..7
  const uint32_t n_chunks = 16u * B + 8u;
  uint32_t mbar_phase = 0;  // bit s = parity to wait for on stage s

  uint32_t t = t0;
  while (t < t1) {
    const uint32_t seg = tile_seg[t];
    const BuildEntry e = build[seg];
    const uint32_t seg_tile_end = (seg + 1 < n_build) ? build[seg + 1].tile_begin : n_tiles;
    const uint32_t te = min(t1, seg_tile_end);
    uint32_t lo = (t - e.tile_begin) * kHistTile;
    uint32_t hi = min(e.count, (te - e.tile_begin) * kHistTile);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin;

    // passes of this warp: rows [jb, jb+8), jb = lo + (warp + i * kWarps) * 8
    const uint32_t first = lo + warp * kRowsPerPass;
    const uint32_t n_pass = first < hi ? (hi - first + kWarps * kRowsPerPass - 1) / (kWarps * kRowsPerPass) : 0u;

    // row this lane copies in every pass (its first chunk); its row id is prefetched two passes early
    auto copy_row_of = [&](uint32_t c) -> uint32_t { return c < 16u * B ? c / (2u * B) : c - 16u * B; };
    const uint32_t my_row = kBulk ? (lane & 7u) : copy_row_of(lane < n_chunks ? lane : 0u);
    auto fetch_rid_row = [&](uint32_t i, uint32_t row) -> uint32_t {
      const uint32_t j = first + i * kWarps * kRowsPerPass + row;
      return __ldg(ridx + (j < hi ? j : lo));  // rows past the end re-read a valid row (zeroed by the consumer)
    };
    auto issue_copies = [&](uint32_t i, uint32_t rid) {
      const uint32_t s = i % kStages;
      const uint32_t stage = ring_base + s * stage_bytes;
      if (kBulk) {
        const uint32_t bar = bar_base + 8u * s;
        if (lane == 0) MbarExpectTx(bar, stage_bytes);
        // lanes 0..7: symbol slice of row `lane`; lanes 8..15: gradient pair of row lane-8
        if (lane < 8u)
          BulkCopyG2S(stage + my_row * row_b, ell.bins + (size_t)rid * ell.row_bytes + g.byte_begin, row_b, bar);
        else if (lane < 16u)
          BulkCopyG2S(stage + 8u * row_b + my_row * 16u, gq + rid, 16u, bar);
      } else {
        for (uint32_t c = lane; c < n_chunks; c += 32u) {
          if (c != lane) rid = fetch_rid_row(i, copy_row_of(c));
          if (c < 16u * B) {
            const uint32_t row = c / (2u * B), part = c % (2u * B);
            CpAsync16(stage + row * row_b + part * 16u,
                      ell.bins + (size_t)rid * ell.row_bytes + g.byte_begin + part * 16u);
          } else {
            const uint32_t row = c - 16u * B;
            CpAsync16(stage + 8u * row_b + row * 16u, gq + rid);
          }
        }
        CpAsyncCommit();
      }
    };

    auto consume = [&](uint32_t i) {
      const uint32_t s = i % kStages;
      const uint32_t stage = ring_base + s * stage_bytes;
      const uint32_t jb = first + i * kWarps * kRowsPerPass;
      const bool tail = jb + kRowsPerPass > hi;  // warp-uniform
      uint32_t glo[2], ghi[2], hlo[2], hhi[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint4 gp = LdsV4(stage + 8u * row_b + (u * 4 + q) * 16u);
        glo[u] = gp.x; ghi[u] = gp.y; hlo[u] = gp.z; hhi[u] = gp.w;
        if (tail && jb + u * 4 + q >= hi) { glo[u] = 0u; ghi[u] = 0u; hlo[u] = 0u; hhi[u] = 0u; }
      }
      for (uint32_t blk = 0; blk < B; ++blk) {
        const uint32_t blk_base = sh_base + blk * S * 512u;
        if (!kChecked) {  // lean path: same hand-scheduled accumulation as HistKernelBankedFast
          FastData<1> d;
          uint32_t off[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) off[k] = blk_base + lane_off[k];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            d.word[u] = LdsU32(stage + (u * 4 + q) * row_b + blk * 32u + 4u * w);
            d.glo[u] = glo[u]; d.ghi[u] = ghi[u]; d.hlo[u] = hlo[u]; d.hhi[u] = hhi[u];
          }
          if (h_lo_zero && !tail) FastProcess<true>(d, sel, off);
          else FastProcess<false>(d, sel, off);
          continue;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t word = LdsU32(stage + (u * 4 + q) * row_b + blk * 32u + 4u * w);
          uint32_t addr[4], og[4], oh[4], vgl[4], vgh[4], vhl[4], vhh[4];
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t sym = __byte_perm(word, 0u, sel[k]);
            vgl[k] = glo[u]; vgh[k] = ghi[u]; vhl[k] = hlo[u]; vhh[k] = hhi[u];
            const bool ok = (blk * 32u + 4u * w + ((k + q) & 3u) < nf) && sym < S;
            addr[k] = blk_base + (ok ? sym : 0u) * 512u + lane_off[k];
            if (!ok) { vgl[k] = 0u; vgh[k] = 0u; vhl[k] = 0u; vhh[k] = 0u; }
          }
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            og[k] = AtomAddShared(addr[k], vgl[k]);
            oh[k] = AtomAddShared256(addr[k], vhl[k]);
          }
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) {
            RedAddHi128(addr[k], og[k], vgl[k], vgh[k]);
            RedAddHi384(addr[k], oh[k], vhl[k], vhh[k]);
          }
        }
      }
    };

    if (n_pass > 0) {
      // prologue: copies for passes 0 .. kStages-2; row ids for the next two passes are kept in flight
      for (uint32_t i = 0; i + 1 < kStages; ++i) {
        if (i < n_pass) issue_copies(i, fetch_rid_row(i, my_row));
        else if (!kBulk) CpAsyncCommit();
      }
      uint32_t rid_a = (kStages - 1 < n_pass) ? fetch_rid_row(kStages - 1, my_row) : 0u;
      uint32_t rid_b = (kStages < n_pass) ? fetch_rid_row(kStages, my_row) : 0u;
      for (uint32_t i = 0; i < n_pass; ++i) {
        const uint32_t s = i % kStages;
        if (kBulk) {
          MbarWait(bar_base + 8u * s, (mbar_phase >> s) & 1u);
          mbar_phase ^= 1u << s;
        } else {
          CpAsyncWait<kStages - 2>();
        }
        __syncwarp();  // pass i is visible to every lane; every lane is done reading stage (i-1) % kStages
        const uint32_t ahead = i + kStages - 1;
        if (ahead < n_pass) {
          issue_copies(ahead, rid_a);
          rid_a = rid_b;
          if (ahead + 2 < n_pass) rid_b = fetch_rid_row(ahead + 2, my_row);
        } else if (!kBulk) {
          CpAsyncCommit();  // keep the group count uniform
        }
        consume(i);
      }
      if (!kBulk) CpAsyncWait<0>();
    }
    flush((int32_t)e.slot);
    t = te;
  }
}

#endif  // B200_EMU
// ---------------------------------------------------------------------------------------------
// Generic kernel (uint16 symbols): linear {g.lo, g.hi, h.lo, h.hi} per bin, tile staged through smem
// ---------------------------------------------------------------------------------------------
template <typename BinT>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernelGeneric(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
                  const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
                  const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
                  const GroupDesc* __restrict__ groups, longlong2* __restrict__ hist) {
  B200_DYNAMIC_SHARED_ALIGNED(uint8_t, smem, 16);
  constexpr int kPerWord = 4 / sizeof(BinT);
  constexpr uint32_t kSymMask = (1u << (8 * sizeof(BinT))) - 1u;

  const GroupDesc g = groups[blockIdx.y];
  const uint32_t gbins = g.bin_end - g.bin_begin;
  const uint32_t nf = g.f_end - g.f_begin;
  uint32_t* sh_hist = reinterpret_cast<uint32_t*>(smem);
  longlong2* s_gq = reinterpret_cast<longlong2*>(smem + (((size_t)gbins * 16 + 15) & ~size_t(15)));
  uint32_t* s_ridx = reinterpret_cast<uint32_t*>(s_gq + kHistTile);
  uint32_t* s_binoff = s_ridx + kHistTile;

  const uint32_t n_tiles = ts->build_tiles;
  if (n_tiles == 0) return;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  if (t0 >= t1) return;

  for (uint32_t i = threadIdx.x; i < nf; i += kHistThreads)
    s_binoff[i] = ell.cut_ptrs[g.f_begin + i] - g.bin_begin;
  for (uint32_t i = threadIdx.x; i < gbins * 4; i += kHistThreads) sh_hist[i] = 0u;
  __syncthreads();

  auto flush = [&](int32_t nidx) {
    __syncthreads();
    longlong2* dst = hist + (size_t)nidx * ell.total_bins + g.bin_begin;
    for (uint32_t b = threadIdx.x; b < gbins; b += kHistThreads) {
      uint4 v = reinterpret_cast<uint4*>(sh_hist)[b];
      if (v.x | v.y) atomicAdd(reinterpret_cast<unsigned long long*>(&dst[b].x), ((unsigned long long)v.y << 32) | v.x);
      if (v.z | v.w) atomicAdd(reinterpret_cast<unsigned long long*>(&dst[b].y), ((unsigned long long)v.w << 32) | v.z);
      reinterpret_cast<uint4*>(sh_hist)[b] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
  };

  const uint32_t W = g.n_words;
  int32_t cur_seg = -1;
  int32_t cur_nidx = -1;
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t seg = tile_seg[t];
    const BuildEntry e = build[seg];
    if ((int32_t)seg != cur_seg) {
      if (cur_seg >= 0) flush(cur_nidx);
      cur_seg = (int32_t)seg;
      cur_nidx = (int32_t)e.slot;
    }
    const uint32_t off = (t - e.tile_begin) * kHistTile;
    const uint32_t n = min((uint32_t)kHistTile, e.count - off);
    const uint32_t* __restrict__ ridx = e.buf ? ridx1 : ridx0;
    for (uint32_t j = threadIdx.x; j < n; j += kHistThreads) {
      uint32_t r = ridx[e.begin + off + j];
      s_ridx[j] = r;
      s_gq[j] = __ldg(gq + r);
    }
    __syncthreads();
    const uint32_t items = n * W;
    for (uint32_t idx = threadIdx.x; idx < items; idx += kHistThreads) {
      const uint32_t j = idx / W;
      const uint32_t w = idx - j * W;
      const uint32_t r = s_ridx[j];
      if (r < ell.page_lo || r >= ell.page_hi) continue;  // external memory: only the resident page's rows
      const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(ell.bins + (size_t)r * ell.row_bytes + g.byte_begin) + w);
      const longlong2 gp = s_gq[j];
#pragma unroll
      for (int k = 0; k < kPerWord; ++k) {
        const uint32_t fl = w * kPerWord + k;
        if (fl < nf) {
          const uint32_t sym = (word >> (8 * sizeof(BinT) * k)) & kSymMask;
          if (sym != ell.null_symbol) {
            uint32_t* acc = sh_hist + 4u * (s_binoff[fl] + sym);
            SmemAdd64(acc, acc + 1, gp.x);
            SmemAdd64(acc + 2, acc + 3, gp.y);
          }
        }
      }
    }
    __syncthreads();
  }
  flush(cur_nidx);
}

// ---------------------------------------------------------------------------------------------
// Sparse ELLPACK pages (global-bin symbols, ~1 % density): the histogram of a node spans every bin of every
// feature (4 MiB at 1024 x 256 bins), far beyond shared memory, and a row contributes only its ~10 present
// entries, so the accumulation goes straight to L2 with 64-bit reductions (RED.64) -- the same choice the
// reference makes for sparse pages (feature_groups.cu: single group => global-memory histogram).
// One thread per (row, entry); `bin_begin/bin_end` restrict a rank to its feature shard.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
HistKernelSparse(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
                 const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
                 const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
                 longlong2* __restrict__ hist, uint32_t bin_begin, uint32_t bin_end) {
  const uint32_t n_tiles = ts->build_tiles;
  const uint32_t stride = ell.row_stride;
  const uint32_t* __restrict__ sym_base = reinterpret_cast<const uint32_t*>(ell.bins);
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const BuildEntry e = build[tile_seg[t]];
    const uint32_t off = (t - e.tile_begin) * kHistTile;
    const uint32_t n = min((uint32_t)kHistTile, e.count - off);
    const uint32_t* __restrict__ ridx = (e.buf ? ridx1 : ridx0) + e.begin + off;
    longlong2* __restrict__ dst = hist + (size_t)e.slot * ell.total_bins;
    const uint32_t items = n * stride;
    for (uint32_t idx = threadIdx.x; idx < items; idx += 256) {
      const uint32_t j = idx / stride, k = idx - j * stride;
      const uint32_t r = __ldg(ridx + j);
      const uint32_t sym = __ldg(sym_base + (size_t)r * stride + k);
      if (sym >= bin_begin && sym < bin_end) {  // also drops the null symbol (>= total_bins)
        const longlong2 gp = __ldg(gq + r);
        atomicAdd(reinterpret_cast<unsigned long long*>(&dst[sym].x), (unsigned long long)gp.x);
        atomicAdd(reinterpret_cast<unsigned long long*>(&dst[sym].y), (unsigned long long)gp.y);
      }
    }
  }
}

#ifndef B200_KERNELS_ONLY
void PrepareHistKernels(int device) {
  int optin = SmemOptin(device);
  B200_CUDA(cudaFuncSetAttribute(HistKernelBanked<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelBankedFast<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelBankedFast<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelPair<2, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelPair<3, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelPair<2, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelPair<3, 1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelPair<2, 2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelBankedFast<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute((HistKernelBankedFast<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelBanked<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelStaged<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelStaged<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelStaged<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelStaged<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernelGeneric<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
}
#endif  // B200_KERNELS_ONLY

#ifndef B200_KERNELS_ONLY
void LaunchBuildHist(const EllView& ell, const b200hist_gpair64* gq, const uint32_t* ridx0, const uint32_t* ridx1,
                     const BuildEntry* build, const uint32_t* tile_seg, const TreeState* ts,
                     const GroupDesc* groups_dev, const std::vector<GroupDesc>& groups_host,
                     b200hist_gpair64* hist, const QuantDev* quant, int num_sms, cudaStream_t s, bool pair_hint) {
  if (ell.is_sparse) {
    uint32_t b0 = 0xFFFFFFFFu, b1 = 0;
    for (auto const& g : groups_host) { b0 = std::min(b0, g.bin_begin); b1 = std::max(b1, g.bin_end); }
    HistKernelSparse<<<num_sms * 8, 256, 0, s>>>(ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1, build,
                                                 tile_seg, ts, reinterpret_cast<longlong2*>(hist), b0, b1);
    B200_LAUNCHED();
    return;
  }
  const int n_groups = (int)groups_host.size();
  const int workers = std::max(1, num_sms / n_groups);
  dim3 grid(workers, n_groups);
  const bool banked = ell.bin_bytes == 1 || (ell.bin_bytes == 2 && !groups_host.empty() && groups_host[0].n_blocks > 0);
  if (banked) {
    const uint32_t sym_bytes = ell.bin_bytes;
    size_t smem = 0;
    for (auto const& g : groups_host) smem = std::max(smem, BankedSmemBytes(g));
    const bool paged = ell.page_hi != 0xFFFFFFFFu;
    const int impl = paged ? 0 : HistImpl();
    // the default kernels give the null symbol its own accumulator slot (MakeFeatureGroups); the others predicate
    bool checked = ell.null_symbol != 0xFFFFFFFFu && impl != 0;
    uint32_t max_blocks = 1;
    // a trailing partial block is fine for the default kernels when the rows are padded to whole blocks
    const bool padded_rows = impl == 0 && ell.row_bytes >= ((ell.n_features + 31u) & ~31u) * sym_bytes;
    for (auto const& g : groups_host) {
      const bool partial = (g.f_end - g.f_begin) % 32u != 0u;
      checked = checked || (partial && !(padded_rows && g.f_end == ell.n_features));
      max_blocks = std::max(max_blocks, g.n_blocks);
    }
    // Pair kernel (64 features per CTA pass, 12-byte accumulators): chosen when the caller expects hessians
    // without low words and the groups are single full blocks laid out back to back; it checks the flag itself.
    static const bool pair_off = std::getenv("B200HIST_NO_PAIR") != nullptr;
    bool pair = pair_hint && !pair_off && impl == 0 && !checked && n_groups >= 2 &&
                ell.row_bytes % (8u * sym_bytes) == 0u;
    uint32_t pair_sym = 0;
    for (int i = 0; pair && i < n_groups; ++i) {
      const GroupDesc& g = groups_host[i];
      pair = g.n_blocks == 1 && g.byte_begin % (8u * sym_bytes) == 0u &&
             (i == 0 || (g.f_begin == groups_host[i - 1].f_end &&
                         g.byte_begin == groups_host[i - 1].byte_begin + 32u * sym_bytes));
      pair_sym = std::max(pair_sym, g.max_sym);
    }
    const size_t pair_smem = (size_t)pair_sym * 192 * 4;
    int cur_dev = 0;
    if (pair) B200_CUDA(cudaGetDevice(&cur_dev));
    if (pair && std::max(pair_smem, smem) <= (size_t)SmemOptin(cur_dev)) {
      const int n_pairs = (n_groups + 1) / 2;
      dim3 pgrid(std::max(1, num_sms / n_pairs), n_pairs);
      static const int pair_bufs = [] {  // register buffers of prefetch in the pair kernel (2 or 3)
        const char* e = std::getenv("B200HIST_PAIR_BUFS");
        return e ? std::atoi(e) : 3;
      }();
      // uint16 symbols cost two more registers per row buffer: two buffers there (measured equal on uint8)
      auto pk = sym_bytes == 2 ? HistKernelPair<2, 2> : (pair_bufs == 2 ? HistKernelPair<2, 1> : HistKernelPair<3, 1>);
      if (paged) pk = sym_bytes == 2 ? HistKernelPair<2, 2, true> : HistKernelPair<3, 1, true>;
      pk<<<pgrid, kHistThreads, std::max(pair_smem, smem), s>>>(
          ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1, build, tile_seg, ts, groups_dev, (uint32_t)n_groups,
          reinterpret_cast<longlong2*>(hist), quant);
      B200_LAUNCHED();
      return;
    }
    using KernelT = void (*)(EllView, const longlong2*, const uint32_t*, const uint32_t*, const BuildEntry*,
                             const uint32_t*, const TreeState*, const GroupDesc*, longlong2*, const QuantDev*);
    KernelT kernel;
    if (paged) {  // external memory: the default kernels with the page clip compiled in
      B200_CHECK(!checked, "internal: external-memory pages are padded to whole feature blocks");
      kernel = sym_bytes == 2 ? HistKernelBankedFast<2, true> : HistKernelBankedFast<1, true>;
    } else if (impl == 0 || impl == 3) {  // 3 = the unscheduled direct-load kernel (for comparison)
      B200_CHECK(sym_bytes == 1 || (impl == 0 && !checked), "internal: uint16 pages use the default banked kernels only");
      kernel = checked ? HistKernelBanked<true>
                       : (impl == 0 ? (sym_bytes == 2 ? HistKernelBankedFast<2> : HistKernelBankedFast<1>)
                                    : HistKernelBanked<false>);
    } else {
      smem += RingSmemBytes(max_blocks);
      if (impl == 1) kernel = checked ? HistKernelStaged<true, true> : HistKernelStaged<false, true>;
      else kernel = checked ? HistKernelStaged<true, false> : HistKernelStaged<false, false>;
    }
    kernel<<<grid, kHistThreads, smem, s>>>(ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1, build, tile_seg,
                                           ts, groups_dev, reinterpret_cast<longlong2*>(hist), quant);
  } else {
    uint32_t max_bins = 0;
    for (auto const& g : groups_host) max_bins = std::max(max_bins, g.bin_end - g.bin_begin);
    HistKernelGeneric<uint16_t><<<grid, kHistThreads, GenericSmemBytes(max_bins), s>>>(
        ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1, build, tile_seg, ts, groups_dev,
        reinterpret_cast<longlong2*>(hist));
  }
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// sibling = parent - built for every (parent, built, sibling) triple of the level, one launch.
// [bin_begin, bin_end): the bins to subtract (a rank's feature slice under row split; everything otherwise)
__global__ void __launch_bounds__(256) SubtractKernel(const SubEntry* __restrict__ sub, const TreeState* __restrict__ ts,
                                                       longlong2* __restrict__ hist, uint32_t total_bins,
                                                       uint32_t bin_begin, uint32_t bin_end) {
  const uint32_t width = bin_end - bin_begin;
  const unsigned long long total = (unsigned long long)ts->n_sub * width;
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256ull) {
    uint32_t k = (uint32_t)(i / width);
    uint32_t b = bin_begin + (uint32_t)(i - (unsigned long long)k * width);
    SubEntry e = sub[k];
    longlong2 p = hist[(size_t)e.parent * total_bins + b];
    longlong2 c = hist[(size_t)e.built * total_bins + b];
    hist[(size_t)e.sub * total_bins + b] = make_longlong2(p.x - c.x, p.y - c.y);
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchSubtract(const SubEntry* sub, const TreeState* ts, b200hist_gpair64* hist, uint32_t total_bins,
                    int num_sms, cudaStream_t s, uint32_t bin_begin, uint32_t bin_end) {
  SubtractKernel<<<num_sms * 8, 256, 0, s>>>(sub, ts, reinterpret_cast<longlong2*>(hist), total_bins, bin_begin,
                                             std::min(bin_end, total_bins));
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

__global__ void __launch_bounds__(256) SubtractFlatKernel(const longlong2* __restrict__ p, const longlong2* __restrict__ c,
                                                           longlong2* __restrict__ o, uint64_t n) {
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
    longlong2 a = p[i], b = c[i];
    o[i] = make_longlong2(a.x - b.x, a.y - b.y);
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchSubtractFlat(const b200hist_gpair64* parent, const b200hist_gpair64* built, b200hist_gpair64* out,
                        uint64_t n_bins, cudaStream_t s) {
  int blocks = (int)std::min<uint64_t>((n_bins + 255) / 256, 148 * 8);
  if (blocks == 0) return;
  SubtractFlatKernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const longlong2*>(parent),
                                            reinterpret_cast<const longlong2*>(built),
                                            reinterpret_cast<longlong2*>(out), n_bins);
  B200_LAUNCHED();
}
#endif  // B200_KERNELS_ONLY

// ---------------------------------------------------------------------------------------------
// Measurement helper: the shared-memory atomic ceiling of the histogram kernel, measured rather than modelled.
// Runs the exact accumulation sequence of HistKernelBankedFast (FastProcess: conflict-free ATOMS with return,
// carry, two reductions per int64 component) on register-generated symbols -- no global loads, same 1024-thread
// CTA, same shared-memory footprint, one CTA per SM -- and reports updates per clock per SM from clock64().
// ---------------------------------------------------------------------------------------------
#ifndef B200_EMU
template <bool kHLoZero>
__global__ void __launch_bounds__(kHistThreads, 1) SmemAtomicPeakKernel(uint32_t iters, uint32_t words,
                                                                         unsigned long long* __restrict__ clocks,
                                                                         uint32_t* __restrict__ sink) {
  B200_DYNAMIC_SHARED_ALIGNED(uint32_t, sh, 16);
  __shared__ unsigned long long clk_s;
  for (uint32_t i = threadIdx.x; i < words; i += kHistThreads) sh[i] = 0u;
  if (threadIdx.x == 0) clk_s = 0ull;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t q = lane >> 3, w = lane & 7u;
  const uint32_t sh_base = (uint32_t)__cvta_generic_to_shared(sh);
  uint32_t sel[4], off[4];
#pragma unroll
  for (uint32_t k = 0; k < 4; ++k) {
    const uint32_t jj = (k + q) & 3u;
    sel[k] = 0x4440u | jj;
    off[k] = sh_base + (4u * w + jj) * 4u;
  }
  FastData<1> d;
  uint32_t x = 0x9E3779B9u * (threadIdx.x + 1u) + blockIdx.x;
#pragma unroll
  for (int u = 0; u < 2; ++u) { d.glo[u] = 0xFFFFFFF0u + u; d.ghi[u] = 3u; d.hlo[u] = kHLoZero ? 0u : 0xFFFFFF00u; d.hhi[u] = 1u; }
  const long long t0 = clock64();
  for (uint32_t it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    d.word[0] = x;
    d.word[1] = x ^ (x >> 11) ^ 0x5bd1e995u;
    FastProcess<kHLoZero>(d, sel, off);
  }
  const long long t1 = clock64();
  atomicMax(&clk_s, (unsigned long long)(t1 - t0));
  __syncthreads();
  if (threadIdx.x == 0) {
    clocks[blockIdx.x] = clk_s;
    sink[blockIdx.x] = sh[x % words];
  }
}
#endif  // B200_EMU

#ifndef B200_KERNELS_ONLY
double MeasureSmemAtomicPeak(int device, bool h_lo_zero, cudaStream_t s) {
  const uint32_t words = 256u * 32u * 4u;  // one block of 32 features x 256 symbols x int64 pair = 128 KiB
  const size_t smem = words * sizeof(uint32_t);
  const int n_sms = NumSMs(device);
  const uint32_t iters = 20000;
  auto kernel = h_lo_zero ? SmemAtomicPeakKernel<true> : SmemAtomicPeakKernel<false>;
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* clocks = DevAlloc<unsigned long long>(n_sms);
  uint32_t* sink = DevAlloc<uint32_t>(n_sms);
  for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
    kernel<<<n_sms, kHistThreads, smem, s>>>(iters, words, clocks, sink);
    B200_LAUNCHED();
  }
  std::vector<unsigned long long> h(n_sms);
  B200_CUDA(cudaMemcpyAsync(h.data(), clocks, sizeof(unsigned long long) * n_sms, cudaMemcpyDeviceToHost, s));
  B200_CUDA(cudaStreamSynchronize(s));
  cudaFree(clocks);
  cudaFree(sink);
  double sum = 0;
  for (auto c : h) sum += (double)c;
  const double avg_clocks = sum / n_sms;
  return (double)kHistThreads * 8.0 * iters / avg_clocks;
}
#endif  // B200_KERNELS_ONLY

}  // namespace b200

#ifndef B200_KERNELS_ONLY
extern "C" int b200hist_smem_atomic_peak(int device, int h_lo_zero, double* updates_per_clk_per_sm, void* stream) {
  B200_API_BEGIN();
  B200_CHECK(updates_per_clk_per_sm != nullptr, "null argument");
  b200::DeviceGuard guard(device);
  *updates_per_clk_per_sm = b200::MeasureSmemAtomicPeak(device, h_lo_zero != 0, static_cast<cudaStream_t>(stream));
  B200_API_END();
}
#endif  // B200_KERNELS_ONLY
