// cuda_emu.h -- TEST-ONLY host emulation of the small CUDA subset the kernels use.
//
// There is no GPU in the build container, so the CPU test-suite compiles the *same*
// kernel sources with g++ (-DSB_EMU) against this shim to check indexing and maths
// against the oracle before spending GPU time.  The emulated library is built as
// tests/_emu/libscythe_b200_emu.so and is loaded ONLY by tests; the product loader
// (scythe_jl_b200/_lib.py) never looks for it and fails loudly without the sm_100a build.
//
// Model: the CUDA threads of a block are cooperative fibers (ucontext) on the calling OS thread, blocks executed
// one after another; a fiber runs until it reaches a barrier (__syncthreads, warp exchange, named bar.sync) and
// then hands over to the next one, so a barrier costs user-level context switches instead of futex sleeps of
// hundreds of oversubscribed OS threads.  __shared__ = function-level static (safe because only one block is live
// at a time), warp shuffles through a per-warp exchange buffer.  Kernels must not spin on another thread's flag
// without a barrier (none does: there is no inter-thread communication outside barriers in this code base).
#pragma once
#ifdef SB_EMU
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }

namespace sbemu {
extern uint3 t_threadIdx, t_blockIdx;   // of the fiber that is running (set at every switch)
extern dim3 t_blockDim, t_gridDim;
extern int t_lin;                       // linear thread id in block
void fiber_yield();                     // run the other fibers of the block once, then come back
struct FiberBarrier {                   // all `n` fibers arrive before any leaves; reusable
  int n, count = 0;
  unsigned gen = 0;
  explicit FiberBarrier(int n_) : n(n_) {}
  void arrive_and_wait() {
    const unsigned g = gen;
    if (++count == n) { count = 0; ++gen; return; }
    while (gen == g) fiber_yield();
  }
};
extern FiberBarrier* g_block_barrier;
extern std::vector<std::unique_ptr<FiberBarrier>> g_warp_barriers;
extern double g_warp_buf[64][32];
extern double g_warp_buf2[64][32];
extern unsigned char* g_dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace sbemu

#define threadIdx (sbemu::t_threadIdx)
#define blockIdx (sbemu::t_blockIdx)
#define blockDim (sbemu::t_blockDim)
#define gridDim (sbemu::t_gridDim)

static inline void __syncthreads() { sbemu::g_block_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { sbemu::g_warp_barriers[sbemu::t_lin / 32]->arrive_and_wait(); }
static inline double __shfl_sync(unsigned, double v, int src, int = 32) {
  int w = sbemu::t_lin / 32, l = sbemu::t_lin % 32;
  sbemu::g_warp_buf[w][l] = v;
  sbemu::g_warp_barriers[w]->arrive_and_wait();
  double r = sbemu::g_warp_buf[w][src & 31];
  sbemu::g_warp_barriers[w]->arrive_and_wait();
  return r;
}
static inline double __shfl_xor_sync(unsigned m, double v, int lanemask, int = 32) {
  return __shfl_sync(m, v, (sbemu::t_lin % 32) ^ lanemask);
}
static inline double __shfl_down_sync(unsigned m, double v, unsigned d, int = 32) {
  int l = sbemu::t_lin % 32;
  return __shfl_sync(m, v, (l + (int)d < 32) ? l + (int)d : l);
}
static inline double __shfl_up_sync(unsigned m, double v, unsigned d, int = 32) {
  int l = sbemu::t_lin % 32;
  return __shfl_sync(m, v, (l - (int)d >= 0) ? l - (int)d : l);
}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void sincospi(double x, double* s, double* c) {
  *s = std::sin(M_PI * x);
  *c = std::cos(M_PI * x);
}
static inline double atomicAdd(double* p, double v) {  // blocks run serially, threads concurrently
  static std::atomic_flag lock = ATOMIC_FLAG_INIT;
  while (lock.test_and_set(std::memory_order_acquire)) {}
  double o = *p;
  *p = o + v;
  lock.clear(std::memory_order_release);
  return o;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline void __threadfence_block() {}
static inline void __threadfence() {}
using std::fma;
using std::fabs;
using std::sqrt;

// ---- tiny runtime shim ------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct sbemu_event* cudaEvent_t;
struct sbemu_event { double t; };
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaStreamNonBlocking = 1 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }   // everything runs at once
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return 0; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)1 << 40; return 0; }
cudaError_t cudaEventCreate(cudaEvent_t* e);
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);

// FP64 tensor-core MMA m8n8k4 (row.col): A[i][k] from lane i*4+k, B[k][j] from lane j*4+k,
// D[i][2q..2q+1] in lane i*4+q
static inline void sb_dmma(double& d0, double& d1, double a, double b) {
  const int w = sbemu::t_lin / 32, lane = sbemu::t_lin % 32, i = lane / 4, q = lane % 4;
  sbemu::g_warp_buf[w][lane] = a;
  sbemu::g_warp_buf2[w][lane] = b;
  sbemu::g_warp_barriers[w]->arrive_and_wait();
  for (int k = 0; k < 4; ++k) {
    const double aik = sbemu::g_warp_buf[w][i * 4 + k];
    d0 += aik * sbemu::g_warp_buf2[w][(2 * q) * 4 + k];
    d1 += aik * sbemu::g_warp_buf2[w][(2 * q + 1) * 4 + k];
  }
  sbemu::g_warp_barriers[w]->arrive_and_wait();
}
// asynchronous global->shared copies (LDGSTS): the emulation copies at once
static inline void sb_cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
static inline void sb_cp_async8(void* dst, const void* src) { std::memcpy(dst, src, 8); }
static inline void sb_prefetch_l2(const void*) {}
static inline void sb_cp_commit() {}
template <int N> static inline void sb_cp_wait() {}
// ---- Blackwell data movement used by sb_ringfft4.cu / sb_chebmma.cu, emulated ------------------------------------
// Tensor Memory (tcgen05.alloc / st / ld, 32x32b shape): 128 lanes x 512 columns of 32 bits per CTA; warp w reaches lanes
// 32 (w % 4) .. +31, thread l of the warp its lane 32 (w % 4) + l.  One block is live at a time, so one static array.
namespace sbemu { extern uint32_t g_tmem[128][512]; }
static inline uint32_t sb_tmem_alloc(uint32_t* smem_slot, int ncols) { (void)ncols; *smem_slot = 0u; return 0u; }   // warp 0 only
static inline void sb_tmem_dealloc(uint32_t, int) {}
static inline uint32_t sb_tmem_warp_base(uint32_t base) { return base + ((uint32_t)(32 * ((sbemu::t_lin / 32) & 3)) << 16); }
static inline void sb_tmem_st4(uint32_t addr, const uint32_t (&r)[4]) {
  const int lane = (int)(addr >> 16) + sbemu::t_lin % 32, col = (int)(addr & 0xffffu);
  for (int q = 0; q < 4; ++q) sbemu::g_tmem[lane][col + q] = r[q];
}
static inline void sb_tmem_ld4(uint32_t addr, uint32_t (&r)[4]) {
  const int lane = (int)(addr >> 16) + sbemu::t_lin % 32, col = (int)(addr & 0xffffu);
  for (int q = 0; q < 4; ++q) r[q] = sbemu::g_tmem[lane][col + q];
}
static inline void sb_tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  const int lane = (int)(addr >> 16) + sbemu::t_lin % 32, col = (int)(addr & 0xffffu);
  for (int q = 0; q < 16; ++q) r[q] = sbemu::g_tmem[lane][col + q];
}
static inline void sb_tmem_wait_st() {}
static inline void sb_tmem_wait_ld4(uint32_t (&)[4]) {}
static inline void sb_tmem_wait_ld16(uint32_t (&)[16]) {}
static inline void sb_tmem_fence_before_sync() {}
static inline void sb_tmem_fence_after_sync() {}
// mbarrier + bulk asynchronous copy (cp.async.bulk, SASS UBLKCP): the emulation copies at issue, so a wait never blocks;
// a copy issued before every reader of the destination has passed a barrier corrupts the emulated run exactly as it
// could on the device (fibers run to the next barrier one after another), which is what the tests are for.
// (the emulated barrier object is 16 bytes; kernels place barriers 16 bytes apart)
struct sb_mbar_t { uint32_t phases, count, pending, tx; };
static inline void sbemu_mbar_check(sb_mbar_t* b) {
  if (b->pending == 0 && b->tx == 0) { ++b->phases; b->pending = b->count; }      // phase complete: re-arm
}
static inline void sb_mbar_init(sb_mbar_t* b, int count) { b->phases = 0; b->count = b->pending = (uint32_t)count; b->tx = 0; }
static inline void sb_mbar_arrive(sb_mbar_t* b) { --b->pending; sbemu_mbar_check(b); }
static inline void sb_mbar_expect_tx(sb_mbar_t* b, unsigned bytes) { b->tx += bytes; --b->pending; sbemu_mbar_check(b); }   // arrive + expect
static inline void sb_mbar_wait(sb_mbar_t* b, unsigned parity) {      // returns once the phase of that parity has completed
  while ((b->phases & 1u) == (parity & 1u)) sbemu::fiber_yield();
}
static inline void sb_fence_mbar_init() {}
static inline void sb_fence_proxy_async() {}
static inline void sb_bulk_g2s(void* dst, const void* src, unsigned bytes, sb_mbar_t* b) {
  std::memcpy(dst, src, bytes);
  b->tx -= bytes;
  sbemu_mbar_check(b);
}
static inline void sb_bulk_s2g(void* dst, const void* src, unsigned bytes) { std::memcpy(dst, src, bytes); }
static inline void sb_bulk_commit() {}
template <int N> static inline void sb_bulk_wait_read() {}
template <int N> static inline void sb_bulk_wait_all() {}
// named barrier for a subset of the block (bar.sync id, nthreads)
namespace sbemu { void named_barrier(int id, int nthreads); }
static inline void sb_bar_sync(int id, int nthreads) { sbemu::named_barrier(id, nthreads); }
#define SB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(sbemu::g_dyn_smem)
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  sbemu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })

#else  // ---------------------------------------------------------------- real CUDA
#include <cuda_runtime.h>
#ifdef __CUDACC__
// FP64 tensor-core MMA (SASS: DMMA).  Measured on B200: 37.0 TFLOP/s, the same pipe as DFMA
// (36.7 TFLOP/s; both together 36.9) -- profiles/fp64_peak_b200.json.
__device__ __forceinline__ void sb_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void sb_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sb_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sb_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void sb_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sb_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sb_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// ---- Blackwell data movement: Tensor Memory as per-thread table storage, mbarrier + bulk asynchronous copies ----------
// tcgen05.alloc / st / ld (SASS UTCALLOC? / STTM / LDTM), 32x32b shape: thread l of warp w reads / writes lane
// 32 (w % 4) + l, consecutive columns.  All of them are warp-collective (.sync.aligned): call from converged warps only.
__device__ __forceinline__ uint32_t sb_tmem_alloc(uint32_t* smem_slot, int ncols) {   // one warp; result lands in *smem_slot
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  return 0u;
}
__device__ __forceinline__ void sb_tmem_dealloc(uint32_t addr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ uint32_t sb_tmem_warp_base(uint32_t base) { return base + ((uint32_t)(32 * ((threadIdx.x >> 5) & 3)) << 16); }
__device__ __forceinline__ void sb_tmem_st4(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void sb_tmem_ld4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sb_tmem_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr) : "memory");
}
__device__ __forceinline__ void sb_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the loaded registers are in/out operands of the wait, so nothing that reads them can be scheduled above it
__device__ __forceinline__ void sb_tmem_wait_ld4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
}
__device__ __forceinline__ void sb_tmem_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void sb_tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void sb_tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// mbarrier (SASS SYNCS) + cp.async.bulk (SASS UBLKCP): one thread arms the barrier with the byte count and issues the
// copy; the copy engine moves the bytes with no register or LSU instruction per element and completes the barrier
struct sb_mbar_t { unsigned long long state; };
__device__ __forceinline__ void sb_mbar_init(sb_mbar_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void sb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sb_mbar_arrive(sb_mbar_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void sb_mbar_expect_tx(sb_mbar_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(sb_mbar_t* b, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "SB_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SB_MBAR_DONE;\n\t"
      "bra SB_MBAR_WAIT;\n\t"
      "SB_MBAR_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sb_bulk_g2s(void* dst, const void* src, unsigned bytes, sb_mbar_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void sb_bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void sb_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif
#define SB_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char _sb_dyn_smem[];   \
  type* name = reinterpret_cast<type*>(_sb_dyn_smem)
#define SB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
