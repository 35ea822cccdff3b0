#include "cuda_emu.h"

#include <atomic>
#include <memory>
#ifdef __SANITIZE_ADDRESS__
#include <sanitizer/asan_interface.h>
#endif

namespace emu {

namespace {

// A pool of worker threads reused by every launch: a CTA needs all of its threads alive at once (they meet at
// barriers), so the pool grows to the largest CTA seen and each worker takes exactly one CUDA thread of the current CTA.
class Pool {
 public:
  ~Pool() {
    {
      std::unique_lock<std::mutex> l(m_);
      stop_ = true;
      ++generation_;
    }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void Run(int n_threads, const std::function<void(int)>& job) {
    while ((int)workers_.size() < n_threads) {
      const int id = (int)workers_.size();
      workers_.emplace_back([this, id]() { Worker(id); });
    }
    {
      std::unique_lock<std::mutex> l(m_);
      job_ = &job;
      active_ = n_threads;
      remaining_ = n_threads;
      ++generation_;
    }
    cv_work_.notify_all();
    std::unique_lock<std::mutex> l(m_);
    cv_done_.wait(l, [&] { return remaining_ == 0; });
    job_ = nullptr;
  }

 private:
  void Worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_work_.wait(l, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
        if (id >= active_) continue;
        job = job_;
      }
      (*job)(id);
      {
        std::unique_lock<std::mutex> l(m_);
        if (--remaining_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* job_{nullptr};
  int active_{0}, remaining_{0};
  uint64_t generation_{0};
  bool stop_{false};
};

Pool& ThePool() {
  static Pool pool;
  return pool;
}

}  // namespace

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
        const std::function<void(int)> job = [&](int t) {
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
        };
#ifdef __SANITIZE_ADDRESS__
        // the slack behind the CTA's dynamic shared memory is off limits: an index past what the launch asked for
        // is a report, as it would be a fault (or a neighbour's data) on the device
        const size_t used = (dyn_smem + 7) / 8 * 8;
        ASAN_POISON_MEMORY_REGION(blk.dyn.data() + used, blk.dyn.size() - used);
#endif
        ThePool().Run(n_threads, job);
#ifdef __SANITIZE_ADDRESS__
        ASAN_UNPOISON_MEMORY_REGION(blk.dyn.data() + used, blk.dyn.size() - used);
#endif
      }
}

}  // namespace emu
