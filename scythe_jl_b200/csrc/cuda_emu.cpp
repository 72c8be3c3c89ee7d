// cuda_emu.cpp -- runtime of the TEST-ONLY CPU emulation shim (see cuda_emu.h).
#ifdef SB_EMU
#include "cuda_emu.h"

#include <chrono>
#include <stdexcept>
#include <ucontext.h>

namespace sbemu {
uint3 t_threadIdx, t_blockIdx;
dim3 t_blockDim, t_gridDim;
int t_lin;
FiberBarrier* g_block_barrier = nullptr;
std::vector<std::unique_ptr<FiberBarrier>> g_warp_barriers;
double g_warp_buf[64][32];
double g_warp_buf2[64][32];
unsigned char* g_dyn_smem = nullptr;
uint32_t g_tmem[128][512];

static std::unique_ptr<FiberBarrier> g_named[16];
void named_barrier(int id, int nthreads) {
  if (!g_named[id] || g_named[id]->n != nthreads) g_named[id].reset(new FiberBarrier(nthreads));
  g_named[id]->arrive_and_wait();
}

// ---- fibers: one ucontext per CUDA thread of the block, all on the calling OS thread
namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
  ucontext_t ctx;
  bool done = false;
};
std::vector<Fiber> g_fibers;
std::vector<unsigned char*> g_stacks;      // kept across launches
ucontext_t g_main;
int g_cur = -1, g_n = 0, g_alive = 0;
dim3 g_block;
const std::function<void()>* g_body = nullptr;

void enter(int lin) {                      // make fiber `lin` the running CUDA thread
  g_cur = lin;
  t_lin = lin;
  t_threadIdx.x = lin % g_block.x;
  t_threadIdx.y = (lin / g_block.x) % g_block.y;
  t_threadIdx.z = lin / (g_block.x * g_block.y);
}
int next_alive(int from) {
  for (int k = 1; k <= g_n; ++k) {
    const int j = (from + k) % g_n;
    if (!g_fibers[j].done) return j;
  }
  return -1;
}
void trampoline() {
  (*g_body)();
  Fiber& me = g_fibers[g_cur];
  me.done = true;
  --g_alive;
  const int nxt = next_alive(g_cur);
  if (nxt < 0) {
    setcontext(&g_main);                   // the block is finished
  } else {
    enter(nxt);
    setcontext(&g_fibers[nxt].ctx);
  }
}
}  // namespace

void fiber_yield() {
  const int me = g_cur, nxt = next_alive(me);
  if (nxt < 0 || nxt == me) throw std::runtime_error("emulated kernel deadlock: a fiber waits on a barrier nobody else can reach");
  enter(nxt);
  swapcontext(&g_fibers[me].ctx, &g_fibers[nxt].ctx);
  // resumed: enter(me) was done by whoever switched back
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  for (int i = 0; i < 16; ++i) g_named[i].reset();
  FiberBarrier bar(nthreads);
  g_block_barrier = &bar;
  g_warp_barriers.clear();
  for (int w = 0; w * 32 < nthreads; ++w) {
    int cnt = std::min(32, nthreads - w * 32);
    g_warp_barriers.emplace_back(new FiberBarrier(cnt));
  }
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
  while ((int)g_stacks.size() < nthreads) g_stacks.push_back(static_cast<unsigned char*>(std::malloc(kStack)));
  g_fibers.assign(nthreads, Fiber{});
  g_n = nthreads;
  g_block = block;
  g_body = &body;
  t_blockDim = block;
  t_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        t_blockIdx.x = bx;
        t_blockIdx.y = by;
        t_blockIdx.z = bz;
        // barriers start clean for every block (a block that ended mid-phase would be a kernel bug)
        bar.count = 0;
        for (auto& wb : g_warp_barriers) wb->count = 0;
        for (int i = 0; i < 16; ++i) g_named[i].reset();
        for (int i = 0; i < nthreads; ++i) {
          Fiber& f = g_fibers[i];
          f.done = false;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = g_stacks[i];
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, trampoline, 0);
        }
        g_alive = nthreads;
        enter(0);
        swapcontext(&g_main, &g_fibers[0].ctx);   // returns when the last fiber of the block has finished
      }
  g_block_barrier = nullptr;
  g_dyn_smem = nullptr;
  g_body = nullptr;
  g_cur = -1;
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
