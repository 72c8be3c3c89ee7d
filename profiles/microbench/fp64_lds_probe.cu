// fp64_lds_probe.cu -- how do FP64 arithmetic and shared-memory accesses share an SM sub-partition on B200?
// k_inv_l2 (ring FFT) runs at FP64-pipe time + shared-memory time, as if the two never overlapped.  This probe times
// controlled mixes: 512 threads per SM (4 warps per sub-partition, like the FFT kernel), per iteration NF independent
// DFMA (8 chains) and NL shared-memory accesses, interleaved or in separate phases, results consumed by integer or FP64.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_lds_probe fp64_lds_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds128(const double2* p) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return r;
}
__device__ __forceinline__ double lds64(const double* p) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
  return r;
}
__device__ __forceinline__ void sts128(double2* p, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr) : "memory");
}
__device__ __forceinline__ void tm_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void tm_st4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ double2 u2c(const uint32_t (&r)[16], int q) {
  return make_double2(__hiloint2double((int)r[4 * q + 1], (int)r[4 * q]), __hiloint2double((int)r[4 * q + 3], (int)r[4 * q + 2]));
}

// MODE 0: NF DFMA only                 1: NL LDS.128 only (integer consume)
//      2: interleaved, integer consume 3: phases (all LDS, then all DFMA), integer consume
//      4: interleaved, LDS feeds DFMA  5: interleaved STS.128            6: interleaved LDS.64 x2
//      7: phases, LDS feeds DFMA (loads first, then arithmetic on them)
template <int MODE, int NF, int NL>
__global__ void __launch_bounds__(512, 1) k_mix(int iters, double* out, long long* cycles, int nthreads_active) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = reinterpret_cast<double2*>(smraw);
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = make_double2(1.0 + 1e-9 * i, 1.0 - 1e-9 * i);
  __syncthreads();
  if (tid >= nthreads_active) return;
  double f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = 1.0 + 0.001 * (tid + c);
  const double ca = 1.0000001, cb = 1e-9;
  unsigned long long acc = 0;
  const double2* my = sm + tid;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double2* p = my + 512 * (it & 7);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i & 7] = fma(f[i & 7], ca, cb);
    } else if (MODE == 1) {
#pragma unroll
      for (int l = 0; l < NL; ++l) { double2 x = lds128(p + 512 * (l & 7)); acc ^= (unsigned long long)__double_as_longlong(x.x) + (unsigned long long)__double_as_longlong(x.y); }
    } else if (MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        if (MODE == 2) { double2 x = lds128(p + 512 * (l & 7)); acc ^= (unsigned long long)__double_as_longlong(x.x) + (unsigned long long)__double_as_longlong(x.y); }
        if (MODE == 4) { double2 x = lds128(p + 512 * (l & 7)); f[l & 7] = fma(f[l & 7], x.x, x.y * 1e-12); }
        if (MODE == 5) { sts128(const_cast<double2*>(p) + 512 * (l & 7), make_double2(f[l & 7], f[(l + 1) & 7])); }
        if (MODE == 6) { const double* q = reinterpret_cast<const double*>(p + 512 * (l & 7)); double a = lds64(q), b = lds64(q + 1); acc ^= (unsigned long long)__double_as_longlong(a) + (unsigned long long)__double_as_longlong(b); }
#pragma unroll
        for (int i = 0; i < NF / NL; ++i) f[i & 7] = fma(f[i & 7], ca, cb);
      }
    } else if (MODE == 3) {
#pragma unroll
      for (int l = 0; l < NL; ++l) { double2 x = lds128(p + 512 * (l & 7)); acc ^= (unsigned long long)__double_as_longlong(x.x) + (unsigned long long)__double_as_longlong(x.y); }
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i & 7] = fma(f[i & 7], ca, cb);
    } else if (MODE == 7) {
      double2 x[NL];
#pragma unroll
      for (int l = 0; l < NL; ++l) x[l] = lds128(p + 512 * (l & 7));
#pragma unroll
      for (int l = 0; l < NL; ++l) f[l & 7] = fma(f[l & 7], x[l].x, x[l].y * 1e-12);
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i & 7] = fma(f[i & 7], ca, cb);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += f[c];
  out[blockIdx.x * 512 + tid] = s + (double)acc * 1e-300;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// FFT-like pass: 16 LDS.128 -> NB dependent-ish FP64 ops on 16 complex values -> 15 LDS.128 (twiddles) + 60 FP64 -> 16 STS.128
// -> named barrier of 128 threads.  VAR 0: as described; 1: no twiddle loads (TMEM-like: tables elsewhere); 2: no barrier;
// 3: no twiddle loads and no barrier; 4: exchange only (no FP64); 5: FP64 only (no shared memory)
//   6: twiddles from Tensor Memory (4 x ld16 + wait each); 7: the same, next batch requested before the current one is used
template <int VAR>
__global__ void __launch_bounds__(512, 1) k_fftlike(int iters, double* out, long long* cycles) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = reinterpret_cast<double2*>(smraw);
  const int tid = threadIdx.x, team = tid >> 7, tl = tid & 127;
  for (int i = tid; i < 12288; i += 512) sm[i] = make_double2(1.0 + 1e-9 * i, 1e-9 * i);
  __syncthreads();
  double2* buf = sm + team * 2176;
  const double2* tw = sm + 8704 + tl;      // 15 x 128 twiddles
  __shared__ uint32_t s_tm;
  uint32_t tb = 0;
  if (VAR == 6 || VAR == 7) {
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tm)), "r"(64) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tb = s_tm + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16);
    if (tid < 128) {
      for (int k = 0; k < 16; ++k) {
        const double wr = 0.999 + 1e-4 * k, wi = 1e-3 * k;
        tm_st4(tb + 4 * k, (uint32_t)__double2loint(wr), (uint32_t)__double2hiint(wr), (uint32_t)__double2loint(wi), (uint32_t)__double2hiint(wi));
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  double2 v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = make_double2(1.0 + 0.01 * k + 1e-3 * tid, 0.5 - 0.01 * k);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (VAR != 5) {
#pragma unroll
      for (int k = 0; k < 16; ++k) { double2 x = lds128(buf + tl + 136 * k); v[k].x += x.x * 1e-30; v[k].y += x.y * 1e-30; }
    }
    if (VAR != 4) {
      // radix-16-like butterfly network: 4 stages of 16 complex add/sub (128 FP64) + 10 constant complex multiplies (40)
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const int h = 1 << st;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (!(k & h)) {
            double2 a = v[k], b = v[k | h];
            v[k] = make_double2(a.x + b.x, a.y + b.y);
            v[k | h] = make_double2(a.x - b.x, a.y - b.y);
          }
        }
        if (st == 1 || st == 2) {
#pragma unroll
          for (int k = 1; k < 16; k += 3) {
            const double wr = 0.92387953251128673848, wi = -0.38268343236508978178;
            v[k] = make_double2(v[k].x * wr - v[k].y * wi, v[k].x * wi + v[k].y * wr);
          }
        }
      }
      // 15 twiddle multiplies
      if (VAR == 6) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t r[16];
          tm_ld16(tb + 16 * b, r);
          tm_wait16(r);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 4 * b + q;
            if (k) { const double2 w = u2c(r, q); v[k] = make_double2(v[k].x * w.x - v[k].y * w.y, v[k].x * w.y + v[k].y * w.x); }
          }
        }
      } else if (VAR == 7) {
        uint32_t r0[16], r1[16];
        tm_ld16(tb, r0);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint32_t (&cur)[16] = (b & 1) ? r1 : r0;
          uint32_t (&nxt)[16] = (b & 1) ? r0 : r1;
          tm_wait16(cur);
          if (b < 3) tm_ld16(tb + 16 * (b + 1), nxt);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k = 4 * b + q;
            if (k) { const double2 w = u2c(cur, q); v[k] = make_double2(v[k].x * w.x - v[k].y * w.y, v[k].x * w.y + v[k].y * w.x); }
          }
        }
      } else
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        double2 w;
        if (VAR == 0 || VAR == 2) w = lds128(tw + 128 * (k - 1));
        else w = make_double2(0.999 + 1e-4 * k, 1e-3 * k);
        v[k] = make_double2(v[k].x * w.x - v[k].y * w.y, v[k].x * w.y + v[k].y * w.x);
      }
      // keep magnitudes bounded
#pragma unroll
      for (int k = 0; k < 16; ++k) { v[k].x *= 0.0625; v[k].y *= 0.0625; }
    }
    if (VAR != 5) {
#pragma unroll
      for (int k = 0; k < 16; ++k) sts128(buf + tl * 17 + k, v[k]);
    }
    if (VAR == 0 || VAR == 1 || VAR == 4 || VAR == 6 || VAR == 7) asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(128) : "memory");
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += v[k].x + v[k].y;
  out[blockIdx.x * 512 + tid] = s;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (VAR == 6 || VAR == 7) {
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tm), "r"(64) : "memory");
  }
}

template <int VAR>
void run_fft(const char* name, double* out, long long* cyc) {
  const int iters = 500;
  cudaFuncSetAttribute(k_fftlike<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 16);
  for (int rep = 0; rep < 2; ++rep) {
    k_fftlike<VAR><<<148, 512, 12288 * 16>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long hc[148];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += hc[i];
  printf("fft-like pass %-64s %8.1f cycles/pass\n", name, mean / 148 / iters);
}

// dependent chain latency, one warp per SM
__global__ void k_lat(int iters, double* out, long long* cycles, int kind) {
  double f = 1.0 + threadIdx.x * 1e-3, g = 0.5;
  const double ca = 1.0000001, cb = 1e-9;
  long long t0 = clock64();
  if (kind == 0) for (int i = 0; i < iters; ++i) f = fma(f, ca, cb);
  else if (kind == 1) for (int i = 0; i < iters; ++i) f = f + cb;
  else for (int i = 0; i < iters; ++i) { f = fma(f, ca, cb); g = fma(g, ca, cb); }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = f + g;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int NF, int NL>
void run(const char* name, int threads, double* out, long long* cyc) {
  const int iters = 1000;
  cudaFuncSetAttribute(k_mix<MODE, NF, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  for (int rep = 0; rep < 2; ++rep) {
    k_mix<MODE, NF, NL><<<148, 512, 8192 * 16>>>(iters, out, cyc, threads);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long hc[148];
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += hc[i];
  mean /= 148 * (double)iters;
  const int nw = threads / 128;   // warps per sub-partition
  printf("%-58s warps/SMSP %d  NF %3d NL %2d : %8.1f cycles/iter", name, nw, NF, NL, mean);
  const double fp = NF * 2.0 * nw, ld = NL * 4.0 * (threads / 32);
  printf("   [FP64 pipe floor %6.1f, smem floor %6.1f, sum %6.1f]\n", (MODE == 1 ? 0 : fp), (MODE == 0 ? 0 : ld), (MODE == 1 ? 0 : fp) + (MODE == 0 ? 0 : ld));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8);
  cudaMalloc(&cyc, 148 * 8);
  for (int kind = 0; kind < 3; ++kind) {
    k_lat<<<148, 32>>>(4000, out, cyc, kind);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    printf("latency kind %d (%s): %.2f cycles per dependent op\n", kind, kind == 0 ? "DFMA chain" : kind == 1 ? "DADD chain" : "2 DFMA chains", hc[0] / 4000.0);
  }
  for (int threads : {128, 256, 512}) {
    if (threads == 128) { run<0, 128, 8>("DFMA only", 128, out, cyc); run<1, 128, 8>("LDS.128 only", 128, out, cyc); }
    if (threads == 256) { run<0, 128, 8>("DFMA only", 256, out, cyc); run<1, 128, 8>("LDS.128 only", 256, out, cyc); }
    if (threads == 512) { run<0, 128, 8>("DFMA only", 512, out, cyc); run<1, 128, 8>("LDS.128 only", 512, out, cyc); }
  }
  run_fft<0>("(32 exchange + 15 twiddle LDS/STS.128, ~290 FP64, team barrier)", out, cyc);
  run_fft<1>("(no twiddle loads)", out, cyc);
  run_fft<2>("(no barrier)", out, cyc);
  run_fft<3>("(no twiddle loads, no barrier)", out, cyc);
  run_fft<4>("(exchange + barrier only, no FP64)", out, cyc);
  run_fft<5>("(FP64 only)", out, cyc);
  run_fft<6>("(twiddles from TMEM, 4 x ld16 + wait; barrier)", out, cyc);
  run_fft<7>("(twiddles from TMEM, double-buffered ld16; barrier)", out, cyc);
  run<2, 128, 8>("interleaved 16 DFMA : 1 LDS.128, int consume", 512, out, cyc);
  run<3, 128, 8>("phases: 8 LDS.128 then 128 DFMA, int consume", 512, out, cyc);
  run<4, 128, 8>("interleaved, LDS feeds DFMA", 512, out, cyc);
  run<7, 128, 8>("phases: 8 LDS.128, then DFMA on them, then 128 DFMA", 512, out, cyc);
  run<5, 128, 8>("interleaved 16 DFMA : 1 STS.128", 512, out, cyc);
  run<6, 128, 8>("interleaved 16 DFMA : 2 LDS.64", 512, out, cyc);
  run<2, 128, 16>("interleaved 8 DFMA : 1 LDS.128, int consume", 512, out, cyc);
  run<3, 128, 16>("phases: 16 LDS.128 then 128 DFMA, int consume", 512, out, cyc);
  run<7, 128, 16>("phases: 16 LDS.128, DFMA on them, 128 DFMA", 512, out, cyc);
  run<2, 128, 32>("interleaved 4 DFMA : 1 LDS.128, int consume", 512, out, cyc);
  run<3, 128, 32>("phases: 32 LDS.128 then 128 DFMA, int consume", 512, out, cyc);
  run<2, 128, 8>("interleaved 16:1, 2 warps/SMSP", 256, out, cyc);
  run<2, 128, 8>("interleaved 16:1, 1 warp/SMSP", 128, out, cyc);
  return 0;
}
