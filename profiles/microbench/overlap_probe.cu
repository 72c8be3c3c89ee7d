// overlap_probe.cu -- can an FP64-bound, whole-SM kernel (the ring FFT: 512 threads, 204 KB shared memory) and an
// HBM-bound streaming kernel share the chip by SM partition?  Kernel A (FFT-like: FP64 + shared-memory exchange, one
// persistent CTA per SM, DYNAMIC work distribution) runs with grid 148 - k on a high-priority stream; kernel B (copy,
// 256 threads, dynamic work distribution, small footprint) on a second stream.  Times: A alone, B alone, both.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k_fp64(int nchunks, int* counter, double* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = reinterpret_cast<double2*>(smraw);
  __shared__ int s_chunk;
  const int tid = threadIdx.x;
  double2 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = make_double2(1.0 + 1e-3 * (tid + k), 0.5);
  for (;;) {
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    const int c = s_chunk;
    __syncthreads();
    if (c >= nchunks) break;
    for (int it = 0; it < 40; ++it) {
#pragma unroll
      for (int r = 0; r < 24; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k].x = fma(v[k].x, 0.9999999, v[k].y * 1e-9); v[k].y = fma(v[k].y, 0.9999999, 1e-9); }
#pragma unroll
      for (int k = 0; k < 8; ++k) sm[tid + 512 * k] = v[k];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) { double2 x = sm[((tid + 64) & 511) + 512 * k]; v[k].x += x.x * 1e-30; }
      __syncthreads();
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k].x + v[k].y;
  out[blockIdx.x * 512 + tid] = s;
}

// streaming copy, dynamic tiles of 64 KB
__global__ void __launch_bounds__(256, 2) k_copy(const double2* __restrict__ src, double2* __restrict__ dst, int ntiles, int* counter) {
  __shared__ int s_tile; extern __shared__ __align__(16) unsigned char pad_[]; if (threadIdx.x == 999) pad_[0] = 1;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_tile;
    __syncthreads();
    if (t >= ntiles) break;
    const double2* s = src + (size_t)t * 4096;
    double2* d = dst + (size_t)t * 4096;
    double2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = s[threadIdx.x + 256 * i];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[threadIdx.x + 256 * i] = x[i];
  }
}

int main() {
  const size_t bytes = (size_t)4 << 30;                 // 4 GiB read + 4 GiB written
  double2 *src, *dst;
  cudaMalloc(&src, bytes); cudaMalloc(&dst, bytes);
  cudaMemset(src, 1, bytes);
  int* cnt; cudaMalloc(&cnt, 1024);
  double* out; cudaMalloc(&out, 148 * 512 * 8);
  const int ntiles = (int)(bytes / 65536);
  cudaFuncSetAttribute(k_fp64, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  int lo, hi;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t sa, sb;
  cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
  cudaEvent_t a0, a1, b0, b1;
  cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
  const int nchunks = 148 * 24;
  auto runA = [&](int grid) { cudaMemsetAsync(cnt, 0, 4, sa); cudaEventRecord(a0, sa); k_fp64<<<grid, 512, 200 * 1024, sa>>>(nchunks, cnt, out); cudaEventRecord(a1, sa); };
  auto runB = [&](int grid) { cudaMemsetAsync(cnt + 64, 0, 4, sb); cudaEventRecord(b0, sb); k_copy<<<grid, 256, 40 * 1024, sb>>>(src, dst, ntiles, cnt + 64); cudaEventRecord(b1, sb); };
  float ta, tb;
  for (int rep = 0; rep < 2; ++rep) { runA(148); cudaDeviceSynchronize(); }
  cudaEventElapsedTime(&ta, a0, a1);
  printf("A alone, grid 148: %.3f ms\n", ta);
  for (int rep = 0; rep < 2; ++rep) { runB(296); cudaDeviceSynchronize(); }
  cudaEventElapsedTime(&tb, b0, b1);
  printf("B alone, grid 296: %.3f ms  (%.0f GB/s read+write)\n", tb, 2.0 * bytes / 1e6 / tb);
  for (int g : {8, 16, 24, 32, 48}) {
    runB(g * 2); cudaDeviceSynchronize();
    cudaEventElapsedTime(&tb, b0, b1);
    printf("B alone on %2d SMs (grid %3d): %.3f ms  (%.0f GB/s, %.1f GB/s per SM)\n", g, 2 * g, tb, 2.0 * bytes / 1e6 / tb, 2.0 * bytes / 1e6 / tb / g);
  }
  for (int k : {0, 8, 16, 24, 32}) {
    for (int order = 0; order < 2; ++order) {
      // order 0: A first then B; order 1: B first then A
      cudaDeviceSynchronize();
      cudaEvent_t w0, w1; cudaEventCreate(&w0); cudaEventCreate(&w1);
      cudaEventRecord(w0, 0);
      cudaStreamWaitEvent(sa, w0, 0); cudaStreamWaitEvent(sb, w0, 0);
      if (order == 0) { runA(148 - k); runB(296); } else { runB(296); runA(148 - k); }
      cudaDeviceSynchronize();
      cudaEventElapsedTime(&ta, a0, a1);
      cudaEventElapsedTime(&tb, b0, b1);
      float sa0, sb1;
      cudaEventElapsedTime(&sa0, w0, a1);
      cudaEventElapsedTime(&sb1, w0, b1);
      printf("both, A grid %3d (k = %2d), %s: A %.3f ms, B %.3f ms, makespan %.3f ms\n", 148 - k, k, order == 0 ? "A launched first" : "B launched first",
             ta, tb, sa0 > sb1 ? sa0 : sb1);
    }
  }
  return 0;
}
