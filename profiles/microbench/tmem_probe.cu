// tmem_probe.cu -- is Tensor Memory usable as per-thread table storage for a non-MMA kernel on B200?
//   1. mapping check: warps 0-3 write f(lane, column) with tcgen05.st.32x32b, all 16 warps of the CTA read it back
//      (warp w sees lanes 32 (w % 4) .. +31) -> error count must be 0;
//   2. throughput of tcgen05.ld.32x32b.x4 / .x16 per SM (bytes per cycle), 16 warps;
//   3. does it run beside the shared-memory pipe?  LDS.128 loop alone, TMEM loop alone, both interleaved;
//   4. the same beside an FP64 FMA stream.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_alloc(uint32_t* smem_dst, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t addr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
}
__device__ __forceinline__ void tm_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint32_t fval(int lane, int col) { return 0x9E3779B9u * (uint32_t)(lane * 512 + col + 1); }

// mode: 0 = mapping check; 1 = ld.x4 loop; 2 = ld.x16 loop; 3 = LDS.128 loop; 4 = LDS.128 + ld.x4 interleaved;
//       5 = DFMA loop; 6 = DFMA + ld.x4 interleaved; 7 = DFMA + LDS.128 interleaved
__global__ void __launch_bounds__(512, 1) k_probe(int mode, int iters, unsigned long long* out, long long* cycles) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = reinterpret_cast<double2*>(smraw);
  __shared__ uint32_t s_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tm_alloc(&s_base, 512);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t base = s_base;
  const int lane = tid & 127;                                   // TMEM lane this thread can reach
  const uint32_t tb = base + ((uint32_t)(32 * (warp & 3)) << 16);   // warp's lane quarter
  for (int i = tid; i < 8192; i += 512) sm[i] = make_double2(1.0 + i, 2.0 - i);
  if (warp < 4) {
    for (int c = 0; c < 512; c += 4) tm_st4(tb + c, fval(lane, c), fval(lane, c + 1), fval(lane, c + 2), fval(lane, c + 3));
    tm_wait_st();
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  unsigned long long acc = 0;
  double facc = 0.0;
  double f0 = 1.0 + tid, f1 = 0.5, f2 = 0.25, f3 = 2.0, f4 = 3.0, f5 = 4.0, f6 = 5.0, f7 = 6.0;
  const double ca = 1.0000001, cb = 1e-9;
  long long t0 = clock64();
  if (mode == 0) {
    for (int c = 0; c < 512; c += 4) {
      uint32_t r[4];
      tm_ld4(tb + c, r);
      tm_wait_ld4(r);
      for (int q = 0; q < 4; ++q) acc += (r[q] != fval(lane, c + q));
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t r[4];
        tm_ld4(tb + ((it * 32 + u * 4) & 511), r);
        tm_wait_ld4(r);
        acc += r[0] ^ r[3];
      }
    }
  } else if (mode == 2) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t r[16];
        tm_ld16(tb + ((it * 32 + u * 16) & 511), r);
        tm_wait_ld16(r);
        acc += r[0] ^ r[15] ^ r[7];
      }
    }
  } else if (mode == 3 || mode == 4 || mode == 7) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 x = sm[(tid + 512 * ((it + u) & 15)) & 8191];
        facc += x.x + x.y;
        if (mode == 4) {
          uint32_t r[4];
          tm_ld4(tb + ((it * 32 + u * 4) & 511), r);
          tm_wait_ld4(r);
          acc += r[0] ^ r[3];
        }
        if (mode == 7) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            f0 = fma(f0, ca, cb); f1 = fma(f1, ca, cb); f2 = fma(f2, ca, cb); f3 = fma(f3, ca, cb);
            f4 = fma(f4, ca, cb); f5 = fma(f5, ca, cb); f6 = fma(f6, ca, cb); f7 = fma(f7, ca, cb);
          }
        }
      }
    }
  } else if (mode == 5 || mode == 6) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          f0 = fma(f0, ca, cb); f1 = fma(f1, ca, cb); f2 = fma(f2, ca, cb); f3 = fma(f3, ca, cb);
          f4 = fma(f4, ca, cb); f5 = fma(f5, ca, cb); f6 = fma(f6, ca, cb); f7 = fma(f7, ca, cb);
        }
        if (mode == 6) {
          uint32_t r[4];
          tm_ld4(tb + ((it * 32 + u * 4) & 511), r);
          tm_wait_ld4(r);
          acc += r[0] ^ r[3];
        }
      }
    }
  }
  long long t1 = clock64();
  facc += f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
  out[blockIdx.x * 512 + tid] = acc + (unsigned long long)(facc * 1e-300);
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) tm_dealloc(base, 512);
}

int main() {
  unsigned long long* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8);
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  const char* names[8] = {"mapping check (errors must be 0)", "TMEM ld.x4", "TMEM ld.x16", "LDS.128", "LDS.128 + TMEM ld.x4",
                          "DFMA x16", "DFMA x16 + TMEM ld.x4", "DFMA x16 + LDS.128"};
  const int iters = 2000;
  for (int mode = 0; mode < 8; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      k_probe<<<148, 512, 8192 * 16>>>(mode, iters, out, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    static unsigned long long h[148 * 512];
    long long hc[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += hc[i];
    mean /= 148;
    if (mode == 0) {
      unsigned long long errs = 0;
      for (int i = 0; i < 148 * 512; ++i) errs += h[i];
      printf("mode 0 %-36s errors = %llu\n", names[0], errs);
    } else {
      const double n8 = (double)iters * 8;      // per-thread unit operations (one LDS.128 / one ld.x4 / 16 DFMA)
      double tm_bytes = (mode == 2 ? (double)iters * 2 * 64 : n8 * 16) * 512;
      printf("mode %d %-36s %10.0f cycles  = %.2f cycles per unit-op per thread", mode, names[mode], mean, mean / n8);
      if (mode == 1 || mode == 2) printf("   TMEM read %.1f B/cycle/SM", tm_bytes / mean);
      if (mode == 3) printf("   smem read %.1f B/cycle/SM", n8 * 16 * 512 / mean);
      if (mode == 5) printf("   %.1f DFMA/cycle/SM", n8 * 16 * 512 / mean);
      printf("\n");
    }
  }
  return 0;
}
