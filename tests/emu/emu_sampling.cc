// Gradient-based row sampling (xgboost_b200/csrc/sampling.cu) under the CPU emulator: the kernels in the order
// LaunchGradientBasedSample issues them, with std::sort / a sequential integer prefix sum where the device calls CUB
// (any correct sort / integer scan produces the same arrays).
#define B200_KERNELS_ONLY 1
#include "cuda_emu.h"
#include "emu_io.h"

#include <algorithm>

#include "../../xgboost_b200/csrc/sampling.cu"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  auto gq = ReadAll<longlong2>(dir + "/gq.bin");
  auto quant = ReadAll<double>(dir + "/quant.bin");
  auto prm = ReadAll<float>(dir + "/param.bin");      // subsample
  auto seed = ReadAll<uint32_t>(dir + "/seed.bin");
  const uint64_t n = gq.size();
  b200::QuantDev q{};
  q.to_fixed_g = quant[0]; q.to_fixed_h = quant[1]; q.to_float_g = quant[2]; q.to_float_h = quant[3];
  q.root_g = 123; q.root_h = 456; q.h_lo_or = 77u; q.h_zero_any = 1u;   // stale values the pass must replace
  std::vector<float> rag(n), thr(n + 1);
  std::vector<long long> csum(n);
  const unsigned blocks = 5;
  std::vector<double> partials(blocks);
  b200::GbsState st{};
  EMU_LAUNCH((blocks), (b200::kSampThreads), 0, b200::RagKernel(gq.data(), n, &q, rag.data()));
  std::copy(rag.begin(), rag.end(), thr.begin());
  std::sort(thr.begin(), thr.begin() + n);
  EMU_LAUNCH((blocks), (b200::kSampThreads), 0, b200::TotalKernel(thr.data(), n, partials.data()));
  EMU_LAUNCH((1), (1), 0, b200::PrepareKernel(partials.data(), (int)blocks, n, &st, thr.data()));
  EMU_LAUNCH((blocks), (b200::kSampThreads), 0, b200::ToFixedKernel(thr.data(), n, &st, csum.data()));
  for (uint64_t i = 1; i < n; ++i) csum[i] += csum[i - 1];
  const uint64_t sample_rows = (uint64_t)((float)n * prm[0]);
  EMU_LAUNCH((blocks), (b200::kSampThreads), 0, b200::ThresholdKernel(thr.data(), csum.data(), n, sample_rows, &st));
  EMU_LAUNCH((1), (1), 0, b200::MuKernel(thr.data(), csum.data(), n, sample_rows, &st));
  EMU_LAUNCH((1), (1), 0, b200::ResetRootKernel(&q));
  EMU_LAUNCH((blocks), (b200::kSampThreads), 0, b200::PoissonKernel(gq.data(), n, rag.data(), &st, &q, seed[0]));
  WriteAll(dir + "/gq_out.bin", gq);
  WriteAll(dir + "/rag.bin", rag);
  WriteAll(dir + "/mu.bin", std::vector<float>{st.mu});
  WriteAll(dir + "/hit.bin", std::vector<unsigned long long>{st.first_hit});
  WriteAll(dir + "/root.bin", std::vector<long long>{q.root_g, q.root_h, (long long)q.h_lo_or, (long long)q.h_zero_any});
  return 0;
}
