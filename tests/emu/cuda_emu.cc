#include "cuda_emu.h"

#include <memory>

namespace emu {

void Launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
  const int n_threads = (int)(block.x * block.y * block.z);
  const int n_warps = (n_threads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk(n_threads);
        blk.dyn.assign(dyn_smem + 256, 0);
        std::vector<std::unique_ptr<Warp>> warps;
        for (int w = 0; w < n_warps; ++w) {
          const int lanes = std::min(32, n_threads - w * 32);
          warps.emplace_back(new Warp(lanes));
          for (int l = 0; l < lanes; ++l) warps.back()->alive[l] = true;
        }
        std::vector<std::thread> threads;
        threads.reserve(n_threads);
        for (int t = 0; t < n_threads; ++t) {
          threads.emplace_back([&, t]() {
            threadIdx = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            blockIdx = uint3{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            Warp* w = warps[t / 32].get();
            g_ctx = Ctx{&blk, w, t % 32};
            body();
            // a returned thread no longer takes part in collectives or barriers
            w->alive[t % 32] = false;
            w->bar.Drop();
            blk.bar.Drop();
          });
        }
        for (auto& th : threads) th.join();
      }
}

}  // namespace emu
