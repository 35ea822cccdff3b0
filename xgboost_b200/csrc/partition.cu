// Row partitioning (ApplySplit on rows): every node segment of the level is split stably into
// [left rows | right rows] in ONE count launch + ONE scatter launch for the whole level.
//
// Replaces (reference paths relative to /root/reference):
//   RowPartitioner::UpdatePositionBatch / SortPositionBatch (cub::DispatchScan + output iterator)
//                                         src/tree/gpu_hist/row_partitioner.cuh:98-205,343-418
//   SortPositionCopyKernel (copy-back)    row_partitioner.cuh:76-87
//   GoLeftOp                              src/tree/updater_gpu_hist.cu:398-419
//
// Radix-style two-pass design:
//   pass 1 (count)   per 2048-row tile: read ridx, gather the split feature's symbol, decide, publish
//                    the decision bits (64 words per tile) and the tile's left count.
//   scan             (updater.cu, single CTA) exclusive scan of tile counts -> per-tile offsets,
//                    children's segments, next level's histogram work list.  Stays on the device:
//                    the reference synchronises the stream here to read counts on the host
//                    (row_partitioner.cuh:401-403).
//   pass 2 (scatter) re-reads ridx + decision bits (not the bin matrix) and writes each row id to its
//                    final slot in the OTHER ridx buffer -- no copy-back pass.
// Algorithmic bytes per partitioned row: 4 (ridx) + 1 symbol [a 32 B sector in practice] + 1/8 +
// 4 (ridx) + 1/8 + 4 (write).
#include "common.cuh"

namespace b200 {

constexpr int kPartThreads = 256;
constexpr int kRowsPerThread = kPartTile / kPartThreads;  // 8
constexpr int kWordsPerTile = kPartTile / 32;             // 64

template <typename BinT, bool kCat>
__global__ void __launch_bounds__(kPartThreads)
PartitionCountKernel(EllView ell, const uint32_t* __restrict__ ridx0, const uint32_t* __restrict__ ridx1,
                     const PartEntry* __restrict__ part, const uint32_t* __restrict__ tile_seg,
                     const TreeState* __restrict__ ts, uint32_t* __restrict__ flags,
                     uint32_t* __restrict__ tile_left, CatRoute cr) {
  __shared__ uint32_t s_cnt[kPartThreads / 32];
  const uint32_t n_tiles = ts->part_tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const PartEntry e = part[tile_seg[t]];
    const uint32_t off = (t - e.tile_begin) * kPartTile;
    const uint32_t n = min((uint32_t)kPartTile, e.count - off);
    const uint32_t* __restrict__ ridx = (e.src_buf ? ridx1 : ridx0) + e.begin + off;
    const uint8_t* __restrict__ col = ell.bins + (size_t)e.findex * sizeof(BinT);
    const BinT* __restrict__ colt =
        ell.bins_t ? reinterpret_cast<const BinT*>(ell.bins_t) + (uint64_t)e.findex * ell.col_stride : nullptr;
    uint32_t cnt = 0;
    // External memory (one launch per resident page): only the rows of [page_lo, page_hi) are decided here; the
    // decision bits and the tile's left count ACCUMULATE over the page launches (the updater clears them before the
    // first page).  A tile lists ascending row ids, so one that lies entirely outside the page is skipped after two loads.
    const bool paged = ell.page_hi != 0xFFFFFFFFu;
    if (paged && (__ldg(ridx + n - 1) < ell.page_lo || __ldg(ridx) >= ell.page_hi)) continue;
    if (ell.is_sparse) {
#pragma unroll
      for (int k = 0; k < kRowsPerThread; ++k) {
        const uint32_t j = k * kPartThreads + threadIdx.x;
        bool left = false;
        if (j < n) {
          const uint32_t fb = ell.cut_ptrs[e.findex], fe = ell.cut_ptrs[e.findex + 1];
          const uint32_t gb = SparseRowBin(ell, ridx[j], fb, fe, (uint32_t)e.findex);
          left = (gb == 0xFFFFFFFFu) ? (e.default_left != 0) : GoLeftBin<kCat>((int32_t)(gb - fb), e.split_lbin, e.nidx, cr);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, left);
        if (lane == 0) {
          flags[(size_t)t * kWordsPerTile + (j >> 5)] = m;
          cnt += __popc(m);
        }
      }
    } else {
      // all 8 row ids, then all 8 symbol gathers, then the decisions: 8 independent gathers in flight per thread
      // (the kernel is latency bound: ncu shows long-scoreboard stalls at 20-26 per issue with the loads serialised)
      uint32_t r[kRowsPerThread], sym[kRowsPerThread];
#pragma unroll
      for (int k = 0; k < kRowsPerThread; ++k) {
        const uint32_t j = k * kPartThreads + threadIdx.x;
        r[k] = __ldg(ridx + min(j, n - 1));
      }
      bool in_page[kRowsPerThread];
#pragma unroll
      for (int k = 0; k < kRowsPerThread; ++k) {
        in_page[k] = r[k] >= ell.page_lo && r[k] < ell.page_hi;  // always true for an in-core page
        sym[k] = !in_page[k] ? 0u
                 : (sizeof(BinT) <= 2 && colt != nullptr)
                     ? (uint32_t)__ldg(colt + r[k])
                     : (uint32_t)__ldg(reinterpret_cast<const BinT*>(col + (size_t)r[k] * ell.row_bytes));
      }
#pragma unroll
      for (int k = 0; k < kRowsPerThread; ++k) {
        const uint32_t j = k * kPartThreads + threadIdx.x;
        const bool left = j < n && in_page[k] &&
                          ((sym[k] == ell.null_symbol) ? (e.default_left != 0)
                                                       : GoLeftBin<kCat>((int32_t)sym[k], e.split_lbin, e.nidx, cr));
        const uint32_t m = __ballot_sync(0xffffffffu, left);
        if (lane == 0) {
          uint32_t* fw = flags + (size_t)t * kWordsPerTile + (j >> 5);
          *fw = paged ? (*fw | m) : m;
          cnt += __popc(m);
        }
      }
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t c = 0;
#pragma unroll
      for (int w = 0; w < kPartThreads / 32; ++w) c += s_cnt[w];
      tile_left[t] = paged ? tile_left[t] + c : c;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPartThreads)
PartitionScatterKernel(const uint32_t* __restrict__ ridx0, const uint32_t* __restrict__ ridx1,
                       uint32_t* __restrict__ out0, uint32_t* __restrict__ out1,
                       const PartEntry* __restrict__ part, const uint32_t* __restrict__ tile_seg,
                       const TreeState* __restrict__ ts, const uint32_t* __restrict__ flags,
                       const uint32_t* __restrict__ tile_left_off) {
  __shared__ uint32_t s_word[kWordsPerTile];
  __shared__ uint32_t s_pref[kWordsPerTile];  // exclusive prefix of left counts over the tile's words
  __shared__ uint32_t s_first_total;          // left count of words 0..31
  const uint32_t n_tiles = ts->part_tiles;
  const int lane = threadIdx.x & 31;
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const PartEntry e = part[tile_seg[t]];
    const uint32_t off = (t - e.tile_begin) * kPartTile;
    const uint32_t n = min((uint32_t)kPartTile, e.count - off);
    const uint32_t* __restrict__ ridx = (e.src_buf ? ridx1 : ridx0) + e.begin + off;
    uint32_t* __restrict__ out = (e.src_buf ? out0 : out1) + e.begin;  // the other buffer
    if (threadIdx.x < kWordsPerTile) {
      // two warps: each scans 32 words, then warp 1 adds warp 0's total
      uint32_t w = flags[(size_t)t * kWordsPerTile + threadIdx.x];
      uint32_t c = __popc(w), inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      s_word[threadIdx.x] = w;
      s_pref[threadIdx.x] = inc - c;
      if (threadIdx.x == 31) s_first_total = inc;
    }
    __syncthreads();
    const uint32_t first_total = s_first_total;
    const uint32_t left_base = tile_left_off[t];           // lefts of this segment before this tile
    const uint32_t right_base = e.n_left + (off - left_base);  // rights before this tile, after all lefts
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
      const uint32_t j = k * kPartThreads + threadIdx.x;
      if (j < n) {
        const uint32_t wi = j >> 5;
        const uint32_t w = s_word[wi];
        const uint32_t lefts_before = s_pref[wi] + (wi >= 32 ? first_total : 0u) + __popc(w & ((1u << lane) - 1u));
        const bool left = (w >> lane) & 1u;
        const uint32_t dst = left ? (left_base + lefts_before) : (right_base + (j - lefts_before));
        out[dst] = ridx[j];
      }
    }
    __syncthreads();
  }
}

#ifndef B200_KERNELS_ONLY
void LaunchPartitionCount(const EllView& ell, const uint32_t* ridx0, const uint32_t* ridx1, const PartEntry* part,
                          const uint32_t* tile_seg, const TreeState* ts, uint32_t* flags, uint32_t* tile_left,
                          int num_sms, cudaStream_t s, CatRoute cr) {
  int blocks = num_sms * 8;
  if (cr.bin_bits != nullptr) {  // a level may hold categorical splits
    if (ell.bin_bytes == 1) {
      PartitionCountKernel<uint8_t, true><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
    } else if (ell.is_sparse) {
      PartitionCountKernel<uint32_t, true><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
    } else {
      PartitionCountKernel<uint16_t, true><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
    }
    B200_LAUNCHED();
    return;
  }
  if (ell.bin_bytes == 1) {
    PartitionCountKernel<uint8_t, false><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
  } else if (ell.is_sparse) {
    PartitionCountKernel<uint32_t, false><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
  } else {
    PartitionCountKernel<uint16_t, false><<<blocks, kPartThreads, 0, s>>>(ell, ridx0, ridx1, part, tile_seg, ts, flags, tile_left, cr);
  }
  B200_LAUNCHED();
}

void LaunchPartitionScatter(const uint32_t* ridx0, const uint32_t* ridx1, uint32_t* out0, uint32_t* out1,
                            const PartEntry* part, const uint32_t* tile_seg, const TreeState* ts,
                            const uint32_t* flags, const uint32_t* tile_left_off, int num_sms, cudaStream_t s) {
  PartitionScatterKernel<<<num_sms * 8, kPartThreads, 0, s>>>(ridx0, ridx1, out0, out1, part, tile_seg, ts, flags,
                                                              tile_left_off);
  B200_LAUNCHED();
}

#endif  // B200_KERNELS_ONLY

}  // namespace b200
