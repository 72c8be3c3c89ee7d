// sb_ringfft.cu -- register-resident Bluestein ring FFT for convolution lengths L = 256 .. 8192.
//
// One "team" of T = L/16 threads owns one complex sequence of length L; every thread keeps 16 complex
// values in registers.  A length-L FFT is nfull radix-16 decimation-in-frequency passes (strided
// butterflies, twiddle, exchange through padded shared memory) followed by one register-local pass
// of radix rf = L / 16^nfull on contiguous blocks.  The output order is digit-reversed; the
// pre-transformed chirp FH is stored in exactly that order (host_fft_dif16), so the pointwise
// product and the first inverse pass happen in registers and no permutation is ever done.
// Shared-memory index padding i + (i >> 4) makes every 128-bit access of every pass conflict-free.
#include "sb_internal.hpp"
#include "sb_fftcore.hpp"

#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace sb {


struct FastCls {
  int log2L, nfull, rf, T, nteams, iters, nrows;
  const double2* twp;   // per strided pass p: [15][Ms_p] twiddles W_{Lb_p}^{j k1}, k1 = 1..15
  int twoff[4];
};

// circular convolution of the team's sequence (held in v as element n1*M + tl) with the chirp.
// FHt is the pre-transformed chirp in DIF16 order, laid out [e][tl] (coalesced).
// On return v holds the natural-order result, element n1*M + tl.
__device__ __forceinline__ void team_conv(double2 (&v)[16], double2* buf, const FastCls& fc, const double2* __restrict__ FHt,
                                          int tl, bool active) {
  const int L = 1 << fc.log2L, T = fc.T;
  // ---- forward strided passes
  int Lb = L;
  for (int p = 0; p < fc.nfull; ++p) {
    const int Ms = Lb >> 4;
    const int b = tl / Ms, j = tl - b * Ms, base = b * Lb + j;
    if (p > 0) {
      __syncthreads();
      if (active) {
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = buf[padi(base + n * Ms)];
      }
    }
    if (active) {
      fft16<false>(v);
      const double2* tw = fc.twp + fc.twoff[p] + j;
#pragma unroll
      for (int k = 1; k < 16; ++k) v[k] = cm(v[k], tw[(k - 1) * Ms]);
#pragma unroll
      for (int k = 0; k < 16; ++k) buf[padi(base + k * Ms)] = v[k];
    }
    Lb = Ms;
  }
  __syncthreads();
  // ---- final forward pass, pointwise product, first inverse pass: all in registers
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = buf[padi(tl * 16 + e)];
    fft_final<false>(v, fc.rf);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = cm(v[e], FHt[e * T + tl]);
    fft_final<true>(v, fc.rf);
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[padi(tl * 16 + e)] = v[e];
  }
  // ---- inverse strided passes (reverse order)
  for (int p = fc.nfull - 1; p >= 0; --p) {
    const int Lbp = L >> (4 * p), Ms = Lbp >> 4;
    const int b = tl / Ms, j = tl - b * Ms, base = b * Lbp + j;
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = buf[padi(base + k * Ms)];
      const double2* tw = fc.twp + fc.twoff[p] + j;
#pragma unroll
      for (int k = 1; k < 16; ++k) v[k] = cmc(v[k], tw[(k - 1) * Ms]);
      fft16<true>(v);
      if (p > 0) {
#pragma unroll
        for (int n = 0; n < 16; ++n) buf[padi(base + n * Ms)] = v[n];
      }
    }
  }
}

// =====================================================================================
// forward: real ring rows -> retained coefficients (see k_fwd_l in sb_transforms.cu for the maths)
// =====================================================================================
__global__ void __launch_bounds__(512) k_fwd_l_fast(DevGrid g, const LWork* __restrict__ work, FastCls fc,
                                                    const RingPlan* __restrict__ plans, const double* __restrict__ blob,
                                                    const double* __restrict__ in, long long in_vs,
                                                    double* __restrict__ mirror, long long mirror_vs,
                                                    double* __restrict__ out, long long out_vs) {
  SB_DYN_SMEM(double2, sm);
  const LWork wk = work[blockIdx.x];
  const int v_ = blockIdx.y;
  const int tid = threadIdx.x;
  const RingPlan pl = plans[wk.r];
  const int n = pl.n, m = pl.m, L = 1 << fc.log2L, T = fc.T, M = L >> 4;
  const int Lp = L + (L >> 4);
  const double2* chirp = reinterpret_cast<const double2*>(blob + pl.off);
  const double2* wkk = chirp + m;
  const double2* ph = wkk + m;
  const double2* FHt = ph + m;
  const int team = tid / T, tl = tid - team * T;
  double2* buf = sm + (size_t)team * Lp;
  double2* stage = sm + (size_t)fc.nteams * Lp;     // [2*nrows - nteams][m], used when iters > 1
  const long long hoff = g.ring_hoff[wk.r];
  const double* src = in + (long long)v_ * in_vs + (long long)g.bz * hoff;
  double* mir = mirror ? mirror + (long long)v_ * mirror_vs + (long long)g.bz * hoff : nullptr;
  const int nseq = 2 * wk.nrows;
  for (int it = 0; it < fc.iters; ++it) {
    const int s = it * fc.nteams + team;             // sequence: row = s/2, half = s&1
    const bool active = s < nseq;
    double2 v[16];
    if (active) {
      const int row = s >> 1, half = s & 1;
      const double* rp = src + (long long)(wk.row0 + row) * n + 2 * half;
      double* mp = mir ? mir + (long long)(wk.row0 + row) * n + 2 * half : nullptr;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int a = n1 * M + tl;
        double2 y = make_double2(0.0, 0.0);
        if (a < m) {
          const double2 x = *reinterpret_cast<const double2*>(rp + 4 * a);
          if (mp) *reinterpret_cast<double2*>(mp + 4 * a) = x;
          y = cm(x, chirp[a]);
        }
        v[n1] = y;
      }
    }
    if (it > 0) __syncthreads();
    team_conv(v, buf, fc, FHt, tl, active);
    // un-chirp and park Y(k), k < m: in the team's own buffer (last iteration) or the stage area
    __syncthreads();
    if (active) {
      const bool last = (it == fc.iters - 1);
      double2* Y = last ? buf : stage + (size_t)s * m;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int a = n1 * M + tl;
        if (a < m) Y[last ? padi(a) : a] = cm(v[n1], chirp[a]);
      }
    }
  }
  __syncthreads();
  // separate the four real sub-sequences, radix-4 combine for k < m, phase + scale
  const double invn = 1.0 / n;
  const int nstage = nseq - fc.nteams;               // sequences parked in `stage` (<= 0: none)
  double* dst = out + (long long)v_ * out_vs + g.ring_woff[wk.r];
  for (int i = tid; i < wk.nrows * m; i += blockDim.x) {
    const int row = i / m, k = i - row * m;
    const int km = k ? m - k : 0;
    const int s0 = 2 * row, s1 = s0 + 1;
    double2 Y0, Y0m, Y1, Y1m;
    if (s0 < nstage) { const double2* y = stage + (size_t)s0 * m; Y0 = y[k]; Y0m = y[km]; }
    else { const double2* y = sm + (size_t)(s0 - (nstage > 0 ? nstage : 0)) * Lp; Y0 = y[padi(k)]; Y0m = y[padi(km)]; }
    if (s1 < nstage) { const double2* y = stage + (size_t)s1 * m; Y1 = y[k]; Y1m = y[km]; }
    else { const double2* y = sm + (size_t)(s1 - (nstage > 0 ? nstage : 0)) * Lp; Y1 = y[padi(k)]; Y1m = y[padi(km)]; }
    double2 S0 = make_double2(0.5 * (Y0.x + Y0m.x), 0.5 * (Y0.y - Y0m.y));
    double2 S1 = make_double2(0.5 * (Y0.y + Y0m.y), -0.5 * (Y0.x - Y0m.x));
    double2 S2 = make_double2(0.5 * (Y1.x + Y1m.x), 0.5 * (Y1.y - Y1m.y));
    double2 S3 = make_double2(0.5 * (Y1.y + Y1m.y), -0.5 * (Y1.x - Y1m.x));
    double2 w1 = wkk[k], w2 = cm(w1, w1), w3 = cm(w2, w1);
    double2 X = (S0 + cm(S1, w1)) + (cm(S2, w2) + cm(S3, w3));
    X = cm(X, ph[k]);
    double* o = dst + (long long)(wk.row0 + row) * g.W;
    if (k == 0) {
      o[0] = X.x * invn;
    } else {
      o[2 * k - 1] = X.x * invn;
      o[2 * k] = X.y * invn;
    }
  }
}

// =====================================================================================
// inverse: spectra (value, d/dr, d2/dr2) -> 5 real rows; rows of a ring are rho = zb*5 + f
// =====================================================================================
__global__ void __launch_bounds__(512) k_inv_l_fast(DevGrid g, const LWork* __restrict__ work, FastCls fc,
                                                    const RingPlan* __restrict__ plans, const double* __restrict__ blob,
                                                    const double* __restrict__ in, long long in_fs, long long in_vs,
                                                    double* __restrict__ out, long long out_fs, long long out_vs,
                                                    int out_is_phys, int var0, unsigned lmask) {
  SB_DYN_SMEM(double2, sm);
  const LWork wk = work[blockIdx.x];
  const int v_ = blockIdx.y;
  const int tid = threadIdx.x;
  const RingPlan pl = plans[wk.r];
  const int n = pl.n, m = pl.m, L = 1 << fc.log2L, T = fc.T, M = L >> 4;
  const int Lp = L + (L >> 4);
  const double2* chirp = reinterpret_cast<const double2*>(blob + pl.off);
  const double2* wkk = chirp + m;
  const double2* ph = wkk + m;
  const double2* FHt = ph + m;
  const int team = tid / T, tl = tid - team * T;
  double2* buf = sm + (size_t)team * Lp;
  const long long woff = g.ring_woff[wk.r], hoff = (out_is_phys == 2 ? g.ring_hoffp : g.ring_hoff)[wk.r];
  const int nseq = 2 * wk.nrows;
  for (int it = 0; it < fc.iters; ++it) {
    const int s = it * fc.nteams + team;
    const int row = s >> 1, half = s & 1;
    const int rho = wk.row0 + row;
    const int zb = rho / 5, f = rho - zb * 5;
    const bool active = s < nseq && ((lmask >> f) & 1);   // rows the equation set does not read are skipped
    double2 v[16];
    if (it > 0) __syncthreads();   // the previous sequence's last pass has finished reading buf
    if (active) {
      const int fin = (f < 3) ? f : 0;
      const double* sp = in + (long long)fin * in_fs + (long long)v_ * in_vs + (long long)zb * g.W + woff;
      // the spectrum-side prologue is built in a ROLLED loop through the team's own smem slots (each
      // thread re-reads exactly what it wrote): keeps the 16-element register tile out of the way of
      // the prologue's many temporaries (fully unrolled it spilled 1 KB under the 128-register cap)
#pragma unroll 2
      for (int n1 = 0; n1 < 16; ++n1) {
        const int k = n1 * M + tl;
        double2 u = make_double2(0.0, 0.0);
        if (k < m) {
          const int km = k ? m - k : 0;
          double2 dk, dm;
          if (k == 0) {
            dk = make_double2((f < 3) ? sp[0] : 0.0, 0.0);
            dm = dk;
          } else {
            dk = cmc(make_double2(2.0 * sp[2 * k - 1], 2.0 * sp[2 * k]), ph[k]);
            dm = cmc(make_double2(2.0 * sp[2 * km - 1], 2.0 * sp[2 * km]), ph[km]);
            if (f == 3) {
              dk = make_double2(-(double)k * dk.y, (double)k * dk.x);
              dm = make_double2(-(double)km * dm.y, (double)km * dm.x);
            } else if (f == 4) {
              const double sk = -(double)k * (double)k, sm_ = -(double)km * (double)km;
              dk = make_double2(sk * dk.x, sk * dk.y);
              dm = make_double2(sm_ * dm.x, sm_ * dm.y);
            }
          }
          // g_b[q] = d(q) conj(w_q)^b,  b = 2*half, 2*half+1 ;  G_b[k] = (g_b[k] + conj g_b[m-k]) / 2
          const double2 wk1 = wkk[k], wm1 = wkk[km];
          double2 ga = dk, gma = dm;
          if (half) {
            const double2 wk2 = cm(wk1, wk1), wm2 = cm(wm1, wm1);
            ga = cmc(dk, wk2);
            gma = cmc(dm, wm2);
          }
          const double2 gb = cmc(ga, wk1), gmb = cmc(gma, wm1);
          const double2 Ga = make_double2(0.5 * (ga.x + gma.x), 0.5 * (ga.y - gma.y));
          const double2 Gb = make_double2(0.5 * (gb.x + gmb.x), 0.5 * (gb.y - gmb.y));
          // h = Ga + i Gb ; feed conj(h) * chirp
          u = cm(make_double2(Ga.x - Gb.y, -(Ga.y + Gb.x)), chirp[k]);
        }
        buf[padi(k)] = u;
      }
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = buf[padi(n1 * M + tl)];
    }
    team_conv(v, buf, fc, FHt, tl, active);
    if (active) {
      const RowDst orow = row_dst(g, out, out_fs, out_vs, out_is_phys, f, v_, var0, hoff, n, zb);
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int a = n1 * M + tl;
        if (a < m) {
          const double2 Y = cm(v[n1], chirp[a]);
          *reinterpret_cast<double2*>(orow.at(4 * a + 2 * half)) = make_double2(Y.x, -Y.y);
        }
      }
    }
  }
}

// =====================================================================================
// host side: class configuration, twiddle tables, FH in DIF16 order
// =====================================================================================
void fast_class_config(int L, int* log2L, int* nfull, int* rf, int* T, int* nteams, int* iters, int* nrows) {
  int p = 0;
  while ((1 << p) < L) ++p;
  *log2L = p;
  *nfull = (p - 1) / 4;
  *rf = 1 << (p - 4 * *nfull);
  *T = L / 16;
  if (L <= 2048) { *nteams = 256 / *T; *iters = 1; }
  else if (L == 4096) { *nteams = 2; *iters = 1; }
  else { *nteams = 1; *iters = 2; }
  *nrows = (*nteams * *iters) / 2;
}

bool fast_class_supported(int L) {
  static const bool generic_only = std::getenv("SB_FFT_GENERIC") != nullptr;   // A/B switch for tests/profiling
  // L = 32 .. 128 (rings of 36 .. 256 points): the v1 register kernel with sub-warp teams of 2 .. 8 threads (one strided
  // radix-16 pass + a register-local radix 2 / 4 / 8 pass); below that the generic shared-memory kernel.  SB_FFT_FAST_MINL: A/B switch
  static const int minL = std::getenv("SB_FFT_FAST_MINL") ? std::atoi(std::getenv("SB_FFT_FAST_MINL")) : 32;
  return !generic_only && L >= minL && L >= 32 && L <= 8192;
}

// host reference of the device pass structure (builds FH in the order the kernel expects)
void host_fft_dif16(double* x, int L) {
  int log2L, nfull, rf, T, nteams, iters, nrows;
  fast_class_config(L, &log2L, &nfull, &rf, &T, &nteams, &iters, &nrows);
  const double PI = 3.14159265358979323846264338327950288;
  std::vector<double> c16(16), s16(16);
  for (int k = 0; k < 16; ++k) { c16[k] = std::cos(-2.0 * PI * k / 16.0); s16[k] = std::sin(-2.0 * PI * k / 16.0); }
  c16[4] = 0.0; c16[12] = 0.0; s16[0] = 0.0; s16[8] = 0.0; c16[0] = 1.0; c16[8] = -1.0; s16[4] = -1.0; s16[12] = 1.0;
  int Lb = L;
  for (int p = 0; p < nfull; ++p) {
    const int Ms = Lb / 16;
    for (int b = 0; b < L; b += Lb)
      for (int j = 0; j < Ms; ++j) {
        double vr[16], vi[16], yr[16], yi[16];
        for (int n1 = 0; n1 < 16; ++n1) { vr[n1] = x[2 * (b + j + n1 * Ms)]; vi[n1] = x[2 * (b + j + n1 * Ms) + 1]; }
        for (int k = 0; k < 16; ++k) {
          long double sr = 0, si = 0;
          for (int n1 = 0; n1 < 16; ++n1) {
            const int q = (n1 * k) & 15;
            sr += (long double)vr[n1] * c16[q] - (long double)vi[n1] * s16[q];
            si += (long double)vr[n1] * s16[q] + (long double)vi[n1] * c16[q];
          }
          const long long jk = ((long long)j * k) % Lb;
          const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)jk / Lb;
          const long double wr = cosl(ang), wi = sinl(ang);
          yr[k] = (double)(sr * wr - si * wi);
          yi[k] = (double)(sr * wi + si * wr);
        }
        for (int k = 0; k < 16; ++k) { x[2 * (b + j + k * Ms)] = yr[k]; x[2 * (b + j + k * Ms) + 1] = yi[k]; }
      }
    Lb = Ms;
  }
  for (int b = 0; b < L; b += rf) {
    double yr[16], yi[16];
    for (int k = 0; k < rf; ++k) {
      long double sr = 0, si = 0;
      for (int n1 = 0; n1 < rf; ++n1) {
        const int q = ((n1 * k) % rf) * (16 / rf);
        sr += (long double)x[2 * (b + n1)] * c16[q] - (long double)x[2 * (b + n1) + 1] * s16[q];
        si += (long double)x[2 * (b + n1)] * s16[q] + (long double)x[2 * (b + n1) + 1] * c16[q];
      }
      yr[k] = (double)sr; yi[k] = (double)si;
    }
    for (int k = 0; k < rf; ++k) { x[2 * (b + k)] = yr[k]; x[2 * (b + k) + 1] = yi[k]; }
  }
}

// per-pass twiddle tables [15][Ms] for class L; offsets (in complex elements) into `tab`
void fast_class_twiddles(int L, std::vector<double>& tab, int off[4]) {
  int log2L, nfull, rf, T, nteams, iters, nrows;
  fast_class_config(L, &log2L, &nfull, &rf, &T, &nteams, &iters, &nrows);
  tab.clear();
  int Lb = L;
  for (int p = 0; p < 4; ++p) off[p] = 0;
  for (int p = 0; p < nfull; ++p) {
    const int Ms = Lb / 16;
    off[p] = (int)(tab.size() / 2);
    tab.resize(tab.size() + (size_t)2 * 15 * Ms);
    double* t = tab.data() + 2 * (size_t)off[p];
    for (int k = 1; k < 16; ++k)
      for (int j = 0; j < Ms; ++j) {
        const long long jk = ((long long)j * k) % Lb;
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)jk / Lb;
        t[2 * ((size_t)(k - 1) * Ms + j)] = (double)cosl(ang);
        t[2 * ((size_t)(k - 1) * Ms + j) + 1] = (double)sinl(ang);
      }
    Lb = Ms;
  }
}

static inline void fcount(const LaunchCtx& c) { if (c.launches) ++*c.launches; }

static FastCls make_cls(int L, const double* twp, const int off[4]) {
  FastCls fc;
  fast_class_config(L, &fc.log2L, &fc.nfull, &fc.rf, &fc.T, &fc.nteams, &fc.iters, &fc.nrows);
  fc.twp = reinterpret_cast<const double2*>(twp);
  for (int i = 0; i < 4; ++i) fc.twoff[i] = off[i];
  return fc;
}

static size_t fast_smem(const FastCls& fc, int max_m) {
  const size_t Lp = ((size_t)1 << fc.log2L) + ((size_t)1 << (fc.log2L - 4));
  size_t stage = (fc.iters > 1) ? (size_t)(2 * fc.nrows - fc.nteams) * max_m : 0;
  return (fc.nteams * Lp + stage) * sizeof(double2);
}

void launch_fwd_l_fast(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                       const int twoff[4], const RingPlan* plans, const double* blob, int nvars, const double* in,
                       long long in_vs, double* mirror, long long mirror_vs, double* out, long long out_vs) {
  FastCls fc = make_cls(L, twp, twoff);
  size_t smem = fast_smem(fc, L / 2);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_fwd_l_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  }
  SB_LAUNCH(k_fwd_l_fast, dim3(nwork, nvars), dim3(fc.nteams * fc.T), smem, c.stream, g, work, fc, plans, blob, in, in_vs,
            mirror, mirror_vs, out, out_vs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_l_fast launch: ") + cudaGetErrorString(e));
  fcount(c);
}

void launch_inv_l_fast(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                       const int twoff[4], const RingPlan* plans, const double* blob, int nvars, const double* in,
                       long long in_fs, long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys,
                       int var0) {
  FastCls fc = make_cls(L, twp, twoff);
  size_t smem = (size_t)fc.nteams * (((size_t)1 << fc.log2L) + ((size_t)1 << (fc.log2L - 4))) * sizeof(double2);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_inv_l_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  }
  SB_LAUNCH(k_inv_l_fast, dim3(nwork, nvars), dim3(fc.nteams * fc.T), smem, c.stream, g, work, fc, plans, blob, in, in_fs,
            in_vs, out, out_fs, out_vs, out_is_phys, var0, c.need.lmask);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_l_fast launch: ") + cudaGetErrorString(e));
  fcount(c);
}

}  // namespace sb
