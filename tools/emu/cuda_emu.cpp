// TEST-ONLY CPU shim runtime (see cuda_emu.h).  One CTA at a time, one OS thread per CUDA thread.
#include "cuda_emu.h"

#include <atomic>

namespace vx_emu {

thread_local uint3e t_threadIdx, t_blockIdx;
thread_local int t_lane_slot = 0;
dim3 g_blockDim, g_gridDim;

static std::vector<unsigned char> g_smem(256 * 1024 + 64);
unsigned char* dyn_smem() {
  uintptr_t p = reinterpret_cast<uintptr_t>(g_smem.data());
  return reinterpret_cast<unsigned char*>((p + 63) & ~uintptr_t(63));
}

// Simple generation barrier usable with a thread count fixed per CTA run.
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, waiting = 0;
  uint64_t gen = 0;
  void reset(int n) { count = n; waiting = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t g = gen;
    if (++waiting == count) { waiting = 0; ++gen; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != g; });
  }
};

static Barrier g_cta_bar;
static Barrier g_warp_bar[32];
static uint32_t g_warp_slot[32][32];

void cta_barrier() { g_cta_bar.wait(); }

uint32_t shfl(uint32_t bits, int src_lane) {
  const int w = t_lane_slot >> 5, lane = t_lane_slot & 31;
  g_warp_slot[w][lane] = bits;
  g_warp_bar[w].wait();
  const uint32_t r = g_warp_slot[w][src_lane & 31];
  g_warp_bar[w].wait();
  return r;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads > 1024 || smem > 256 * 1024) { fprintf(stderr, "vx_emu: bad launch\n"); abort(); }
  g_blockDim = block;
  g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_cta_bar.reset(nthreads);
        for (int w = 0; w < 32; ++w) {
          int n = nthreads - w * 32;
          g_warp_bar[w].reset(n > 32 ? 32 : (n > 0 ? n : 1));
        }
        std::vector<std::thread> ts;
        ts.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
          ts.emplace_back([&, t]() {
            t_lane_slot = t;
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
            body();
          });
        }
        for (auto& th : ts) th.join();
      }
}

}  // namespace vx_emu
