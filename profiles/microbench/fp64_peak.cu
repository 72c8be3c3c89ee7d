// fp64_peak.cu -- measures sustained FP64 FMA (vector pipe) and DMMA m8n8k4 (tensor pipe) throughput
// and both together, on the launching device.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_both(double* out, int iters) {   // even warps DFMA, odd warps DMMA
  if ((threadIdx.x >> 5) & 1) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F> float timeit(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, blocks = sms * 4, threads = 512, iters = 4096;
  double* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  float t1 = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); }, 5);
  double f1 = 2.0 * 16 * iters * (double)blocks * threads / (t1 * 1e-3) / 1e12;
  float t2 = timeit([&] { k_dmma<<<blocks, threads>>>(out, iters); }, 5);
  double f2 = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32) / (t2 * 1e-3) / 1e12;
  float t3 = timeit([&] { k_both<<<blocks, threads>>>(out, iters); }, 5);
  double f3 = (2.0 * 16 * iters * (double)blocks * (threads / 2) + 2.0 * 256 * 8 * iters * (double)blocks * (threads / 64)) / (t3 * 1e-3) / 1e12;
  size_t n = (size_t)1 << 27;  // 2 GiB of double2 per buffer
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
  float t4 = timeit([&] { k_copy<<<sms * 16, 512>>>(a, b, n); }, 5);
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_plus_dmma_tflops\": %.2f, \"copy_gbs\": %.1f}\n",
         p.name, sms, f1, f2, f3, 2.0 * n * 16 / (t4 * 1e-3) / 1e9);
  return 0;
}
