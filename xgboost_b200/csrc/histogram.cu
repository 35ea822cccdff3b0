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
//   * feature groups are sized for the 227 KB of shared memory of an sm_100 SM: 5 groups for
//     256 features x 256 bins instead of the reference's 11 (96 KiB cap), so gpair/ridx are
//     re-read 5x instead of 11x.
//   * the shared histogram is int64 stored as two 32-bit words; accumulation is a 32-bit shared
//     atomic on the low word whose returned old value yields the carry into the high word.
//   * flush to global memory uses 64-bit reductions (RED.64) on non-zero bins only.
#include "common.cuh"

namespace b200 {

constexpr int kHistThreads = 1024;
constexpr uint32_t kMaxGroupFeatures = 512;

static inline size_t Round16(size_t x) { return (x + 15) & ~size_t(15); }

size_t HistSmemBytesFor(uint32_t group_bins) {
  return Round16((size_t)group_bins * 16) + (size_t)kHistTile * 16 + (size_t)kHistTile * 4 + kMaxGroupFeatures * 4;
}

uint32_t HistMaxGroupBins(int device) {
  int optin = 0;
  B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  size_t fixed = (size_t)kHistTile * 20 + kMaxGroupFeatures * 4 + 256;
  B200_CHECK((size_t)optin > fixed + 4096, "device has too little shared memory");
  return (uint32_t)(((size_t)optin - fixed) / 16);
}

// Greedy split of the feature range into groups whose bins fit the shared histogram and whose first
// feature sits on a 32-bit word boundary of the row (4 uint8 / 2 uint16 symbols per word).
std::vector<GroupDesc> MakeFeatureGroups(const b200hist_ellpack_s& ell, uint32_t max_group_bins) {
  std::vector<GroupDesc> groups;
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
    GroupDesc g;
    g.f_begin = f;
    g.f_end = f_end;
    g.bin_begin = ell.cut_ptrs[f];
    g.bin_end = ell.cut_ptrs[f_end];
    g.byte_begin = f * ell.bin_bytes;
    g.n_words = (f_end - f + per_word - 1) / per_word;
    groups.push_back(g);
    f = f_end;
  }
  return groups;
}

// 64-bit add into a shared-memory accumulator held as {lo, hi} 32-bit words.
__device__ __forceinline__ void SmemAdd64(uint32_t* acc, long long v) {
  const uint32_t vlo = (uint32_t)(unsigned long long)v;
  const uint32_t vhi = (uint32_t)((unsigned long long)v >> 32);
  const uint32_t old = atomicAdd(acc, vlo);
  const uint32_t hi_add = vhi + ((uint32_t)(old + vlo) < old ? 1u : 0u);
  if (hi_add != 0u) atomicAdd(acc + 1, hi_add);
}

template <typename BinT>
__global__ void __launch_bounds__(kHistThreads, 1)
HistKernel(EllView ell, const longlong2* __restrict__ gq, const uint32_t* __restrict__ ridx0,
           const uint32_t* __restrict__ ridx1, const BuildEntry* __restrict__ build,
           const uint32_t* __restrict__ tile_seg, const TreeState* __restrict__ ts,
           const GroupDesc* __restrict__ groups, longlong2* __restrict__ hist) {
  extern __shared__ __align__(16) uint8_t smem[];
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
      cur_nidx = e.nidx;
    }
    const uint32_t off = (t - e.tile_begin) * kHistTile;
    const uint32_t n = min((uint32_t)kHistTile, e.count - off);
    const uint32_t* __restrict__ ridx = e.buf ? ridx1 : ridx0;
    // stage the tile's row ids and quantised gradient pairs once; every feature word reuses them
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
      const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(ell.bins + (size_t)r * ell.row_bytes + g.byte_begin) + w);
      const longlong2 gp = s_gq[j];
#pragma unroll
      for (int k = 0; k < kPerWord; ++k) {
        const uint32_t fl = w * kPerWord + k;  // feature index inside the group
        if (fl < nf) {
          const uint32_t sym = (word >> (8 * sizeof(BinT) * k)) & kSymMask;
          if (sym != ell.null_symbol) {
            uint32_t* acc = sh_hist + 4u * (s_binoff[fl] + sym);
            SmemAdd64(acc, gp.x);
            SmemAdd64(acc + 2, gp.y);
          }
        }
      }
    }
    __syncthreads();
  }
  flush(cur_nidx);
}

void PrepareHistKernels(int device) {
  int optin = 0;
  B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  B200_CUDA(cudaFuncSetAttribute(HistKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  B200_CUDA(cudaFuncSetAttribute(HistKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
}

void LaunchBuildHist(const EllView& ell, const b200hist_gpair64* gq, const uint32_t* ridx0, const uint32_t* ridx1,
                     const BuildEntry* build, const uint32_t* tile_seg, const TreeState* ts,
                     const GroupDesc* groups_dev, const std::vector<GroupDesc>& groups_host,
                     b200hist_gpair64* hist, int num_sms, cudaStream_t s) {
  uint32_t max_bins = 0;
  for (auto const& g : groups_host) max_bins = std::max(max_bins, g.bin_end - g.bin_begin);
  size_t smem = HistSmemBytesFor(max_bins);
  int n_groups = (int)groups_host.size();
  int workers = std::max(1, num_sms / n_groups);
  dim3 grid(workers, n_groups);
  if (ell.bin_bytes == 1) {
    HistKernel<uint8_t><<<grid, kHistThreads, smem, s>>>(ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1,
                                                        build, tile_seg, ts, groups_dev,
                                                        reinterpret_cast<longlong2*>(hist));
  } else {
    HistKernel<uint16_t><<<grid, kHistThreads, smem, s>>>(ell, reinterpret_cast<const longlong2*>(gq), ridx0, ridx1,
                                                         build, tile_seg, ts, groups_dev,
                                                         reinterpret_cast<longlong2*>(hist));
  }
  B200_LAUNCHED();
}

// sibling = parent - built for every (parent, built, sibling) triple of the level, one launch.
__global__ void __launch_bounds__(256) SubtractKernel(const SubEntry* __restrict__ sub, const TreeState* __restrict__ ts,
                                                       longlong2* __restrict__ hist, uint32_t total_bins) {
  const unsigned long long total = (unsigned long long)ts->n_sub * total_bins;
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * 256ull) {
    uint32_t k = (uint32_t)(i / total_bins);
    uint32_t b = (uint32_t)(i - (unsigned long long)k * total_bins);
    SubEntry e = sub[k];
    longlong2 p = hist[(size_t)e.parent * total_bins + b];
    longlong2 c = hist[(size_t)e.built * total_bins + b];
    hist[(size_t)e.sub * total_bins + b] = make_longlong2(p.x - c.x, p.y - c.y);
  }
}

void LaunchSubtract(const SubEntry* sub, const TreeState* ts, b200hist_gpair64* hist, uint32_t total_bins,
                    int num_sms, cudaStream_t s) {
  SubtractKernel<<<num_sms * 8, 256, 0, s>>>(sub, ts, reinterpret_cast<longlong2*>(hist), total_bins);
  B200_LAUNCHED();
}

__global__ void __launch_bounds__(256) SubtractFlatKernel(const longlong2* __restrict__ p, const longlong2* __restrict__ c,
                                                           longlong2* __restrict__ o, uint64_t n) {
  for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
    longlong2 a = p[i], b = c[i];
    o[i] = make_longlong2(a.x - b.x, a.y - b.y);
  }
}

void LaunchSubtractFlat(const b200hist_gpair64* parent, const b200hist_gpair64* built, b200hist_gpair64* out,
                        uint64_t n_bins, cudaStream_t s) {
  int blocks = (int)std::min<uint64_t>((n_bins + 255) / 256, 148 * 8);
  if (blocks == 0) return;
  SubtractFlatKernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const longlong2*>(parent),
                                            reinterpret_cast<const longlong2*>(built),
                                            reinterpret_cast<longlong2*>(out), n_bins);
  B200_LAUNCHED();
}

}  // namespace b200
