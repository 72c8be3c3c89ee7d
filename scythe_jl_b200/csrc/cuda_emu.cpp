// cuda_emu.cpp -- runtime of the TEST-ONLY CPU emulation shim (see cuda_emu.h).
#ifdef SB_EMU
#include "cuda_emu.h"

#include <chrono>
#include <mutex>

namespace sbemu {
thread_local uint3 t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local int t_lin;
std::barrier<>* g_block_barrier = nullptr;
std::vector<std::unique_ptr<std::barrier<>>> g_warp_barriers;
double g_warp_buf[64][32];
double g_warp_buf2[64][32];
unsigned char* g_dyn_smem = nullptr;

static std::mutex g_named_mu;
static std::unique_ptr<std::barrier<>> g_named[16];
static int g_named_cnt[16];
void named_barrier(int id, int nthreads) {
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> lk(g_named_mu);
    if (!g_named[id] || g_named_cnt[id] != nthreads) {
      g_named[id].reset(new std::barrier<>(nthreads));
      g_named_cnt[id] = nthreads;
    }
    b = g_named[id].get();
  }
  b->arrive_and_wait();
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  for (int i = 0; i < 16; ++i) g_named[i].reset();
  std::barrier<> bar(nthreads);
  g_block_barrier = &bar;
  g_warp_barriers.clear();
  for (int w = 0; w * 32 < nthreads; ++w) {
    int cnt = std::min(32, nthreads - w * 32);
    g_warp_barriers.emplace_back(new std::barrier<>(cnt));
  }
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
  auto worker = [&](int lin) {
    t_lin = lin;
    t_blockDim = block;
    t_gridDim = grid;
    t_threadIdx.x = lin % block.x;
    t_threadIdx.y = (lin / block.x) % block.y;
    t_threadIdx.z = lin / (block.x * block.y);
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx.x = bx;
          t_blockIdx.y = by;
          t_blockIdx.z = bz;
          body();
          bar.arrive_and_wait();  // block boundary: statics (__shared__) are reused by the next block
        }
  };
  std::vector<std::thread> ths;
  ths.reserve(nthreads);
  for (int i = 1; i < nthreads; ++i) ths.emplace_back(worker, i);
  worker(0);
  for (auto& t : ths) t.join();
  g_block_barrier = nullptr;
  g_dyn_smem = nullptr;
}
}  // namespace sbemu

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new sbemu_event{0.0}; return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }
#endif
