// sb_transforms.cu -- hand-written sm_100a kernels for the semi-spectral transforms.
//
//   forward  (K1, spectralTransform!):  fwd_z (Chebyshev) -> fwd_l (ring FFT) -> fwd_r (spline inner product)
//   solve    (K2, splineTransform!):    spline_solve (banded Cholesky with BC fold, or dense for PERIODIC)
//   inverse  (K3, tileTransform!):      inv_r (spline evaluate) -> inv_l (ring inverse FFT) -> inv_z (Chebyshev)
//
// Layouts (all Float64):
//   physical  P[(d*V + v)*N + i],  i = (hoff[r] + j)*zDim + z            (API layout, z fastest)
//   SZ        per (field,var): bz*hoff[r] + zb*n_r + j                    (z-mode rows of a ring contiguous in lambda)
//   SL        per (field,var): zb*W + woff[r] + p,  p = 0 | 2k-1 (Re) | 2k (Im), k <= ri
//   spectral  B/A[v*S + (zb*ncolp + p)*b_rDim + m]                        (API layout)
// Ring FFTs of length n = 4m (m = ri+1, arbitrary) use a 4-way decimation into two packed complex
// DFT_m, each evaluated with Bluestein's chirp-z algorithm on a power-of-two FFT held in shared memory.
#include "sb_internal.hpp"

#include <algorithm>
#include <cstdio>
#include <stdexcept>

namespace sb {

int sb_sm_count() {
#ifdef SB_EMU
  return 4;                    // the emulation runs blocks one after another: a small persistent grid keeps the tests fast
#else
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
#endif
}

#define SB_CHECK_LAUNCH()                                                             \
  do {                                                                                \
    cudaError_t e_ = cudaGetLastError();                                              \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

static inline void count(const LaunchCtx& c) { if (c.launches) ++*c.launches; }

template <class K>
static void opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  }
}

// =====================================================================================
// Chebyshev forward: u[col][z] -> out[zb][col]
// =====================================================================================
#define ZT_COLS 32

__global__ void __launch_bounds__(512) k_fwd_z(DevGrid g, const ZTile* __restrict__ tiles, int ntiles,
                                               const double* __restrict__ in, long long in_vs,
                                               double* __restrict__ mirror, long long mirror_vs,
                                               double* __restrict__ out, long long out_vs,
                                               const double* __restrict__ fwdT) {
  SB_DYN_SMEM(double, sm);
  const int zDim = g.zDim, bz = g.bz, bzp = g.bzp, zs = zDim | 1;
  double* Ct = sm;                       // [zDim][bzp]
  double* u = sm + (size_t)zDim * bzp;   // [32][zs]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int v = blockIdx.y;
  for (int i = tid; i < zDim * bzp; i += nthr) Ct[i] = fwdT[i];
  const double* src_v = in + (long long)v * in_vs;
  double* mir_v = mirror ? mirror + (long long)v * mirror_vs : nullptr;
  double* out_v = out + (long long)v * out_vs;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const ZTile zt = tiles[t];
    __syncthreads();
    const double* src = src_v + (long long)zt.hcol0 * zDim;
    const int cnt = zt.ncols * zDim;
    for (int i = tid; i < cnt; i += nthr) {
      double val = src[i];
      int c = i / zDim, z = i - c * zDim;
      u[c * zs + z] = val;
      if (mir_v) mir_v[(long long)zt.hcol0 * zDim + i] = val;
    }
    __syncthreads();
    for (int gq = warp; gq * 4 < bzp; gq += nwarps) {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      if (lane < zt.ncols) {
        const double* uc = u + lane * zs;
        const double* cc = Ct + gq * 4;
        for (int z = 0; z < zDim; ++z) {
          double uu = uc[z];
          const double* c4 = cc + z * bzp;
          a0 = fma(c4[0], uu, a0); a1 = fma(c4[1], uu, a1); a2 = fma(c4[2], uu, a2); a3 = fma(c4[3], uu, a3);
        }
        double* o = out_v + zt.out_base + lane;
        int zb = gq * 4;
        if (zb < bz) o[(long long)zb * zt.out_stride] = a0;
        if (zb + 1 < bz) o[(long long)(zb + 1) * zt.out_stride] = a1;
        if (zb + 2 < bz) o[(long long)(zb + 2) * zt.out_stride] = a2;
        if (zb + 3 < bz) o[(long long)(zb + 3) * zt.out_stride] = a3;
      }
    }
  }
}

void launch_fwd_z(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars,
                  const double* in, long long in_vstride, double* mirror, long long mirror_vstride,
                  double* out, long long out_vstride, const double* fwdT) {
  ProfScope prof_scope_(c, "fwd_z");
  int ngroups = g.bzp / 4;
  int nwarps = ngroups < 16 ? ngroups : 16;
  if (nwarps < 1) nwarps = 1;
  size_t smem = ((size_t)g.zDim * g.bzp + (size_t)ZT_COLS * (g.zDim | 1)) * sizeof(double);
  opt_in_smem(k_fwd_z, smem);
  const int nsm = sb_sm_count();
  int gx = ntiles < nsm * 4 ? ntiles : nsm * 4;
  SB_LAUNCH(k_fwd_z, dim3(gx, nvars), dim3(32 * nwarps), smem, c.stream, g, tiles, ntiles, in, in_vstride,
            mirror, mirror_vstride, out, out_vstride, fwdT);
  SB_CHECK_LAUNCH();
  count(c);
}

// =====================================================================================
// Chebyshev inverse: nfields x a[zb][col]  ->  physical slots [col][z] (+ d/dz, d2/dz2 of field 0)
// =====================================================================================
__global__ void __launch_bounds__(512) k_inv_z(DevGrid g, const ZTile* __restrict__ tiles, int ntiles, int var0,
                                               int nfields, const double* __restrict__ in, long long in_fs,
                                               long long in_vs, double* __restrict__ phys,
                                               const double* __restrict__ invM) {
  SB_DYN_SMEM(double, sm);
  const int zDim = g.zDim, bz = g.bz, zp = (zDim + 3) & ~3;
  double* Mt = sm;                                  // [3][bz][zp]
  double* a = sm + (size_t)3 * bz * zp;             // [nfields][bz][32]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  const int v = blockIdx.y;
  const double* Mg = invM + (size_t)(var0 + v) * 3 * bz * zp;
  for (int i = tid; i < 3 * bz * zp; i += nthr) Mt[i] = Mg[i];
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const ZTile zt = tiles[t];
    __syncthreads();
    for (int i = tid; i < nfields * bz * 32; i += nthr) {
      int c = i & 31, rest = i >> 5;
      int zb = rest % bz, f = rest / bz;
      double val = 0.0;
      if (c < zt.ncols)
        val = in[(long long)f * in_fs + (long long)v * in_vs + zt.out_base + (long long)zb * zt.out_stride + c];
      a[i] = val;
    }
    __syncthreads();
    if (lane < zt.ncols) {
      const long long colbase = ((long long)zt.hcol0 + lane) * zDim;
      for (int zg = warp; zg * 4 < zDim; zg += nwarps) {
        const int z0 = zg * 4;
        for (int f = 0; f < nfields; ++f) {
          const double* af = a + (size_t)f * bz * 32 + lane;
          double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
          double d0 = 0, d1 = 0, d2 = 0, d3 = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
          if (f == 0) {
            for (int zb = 0; zb < bz; ++zb) {
              double x = af[zb * 32];
              const double* m0 = Mt + (size_t)zb * zp + z0;
              const double* m1 = m0 + (size_t)bz * zp;
              const double* m2 = m1 + (size_t)bz * zp;
              s0 = fma(m0[0], x, s0); s1 = fma(m0[1], x, s1); s2 = fma(m0[2], x, s2); s3 = fma(m0[3], x, s3);
              d0 = fma(m1[0], x, d0); d1 = fma(m1[1], x, d1); d2 = fma(m1[2], x, d2); d3 = fma(m1[3], x, d3);
              e0 = fma(m2[0], x, e0); e1 = fma(m2[1], x, e1); e2 = fma(m2[2], x, e2); e3 = fma(m2[3], x, e3);
            }
          } else {
            for (int zb = 0; zb < bz; ++zb) {
              double x = af[zb * 32];
              const double* m0 = Mt + (size_t)zb * zp + z0;
              s0 = fma(m0[0], x, s0); s1 = fma(m0[1], x, s1); s2 = fma(m0[2], x, s2); s3 = fma(m0[3], x, s3);
            }
          }
          double* o = phys + ((long long)f * g.V + var0 + v) * g.N + colbase + z0;
          if (z0 < zDim) o[0] = s0;
          if (z0 + 1 < zDim) o[1] = s1;
          if (z0 + 2 < zDim) o[2] = s2;
          if (z0 + 3 < zDim) o[3] = s3;
          if (f == 0) {
            double* oz = phys + ((long long)nfields * g.V + var0 + v) * g.N + colbase + z0;
            double* ozz = phys + ((long long)(nfields + 1) * g.V + var0 + v) * g.N + colbase + z0;
            if (z0 < zDim) { oz[0] = d0; ozz[0] = e0; }
            if (z0 + 1 < zDim) { oz[1] = d1; ozz[1] = e1; }
            if (z0 + 2 < zDim) { oz[2] = d2; ozz[2] = e2; }
            if (z0 + 3 < zDim) { oz[3] = d3; ozz[3] = e3; }
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// parity fast path (no vertical BCs, even zDim <= 64, bz <= KMAX): T_k(-xi) = (-1)^k T_k(xi), so
// level z and its mirror zDim-1-z share the even-mode and odd-mode partial sums E, O:
//   out[z] = E + O,  out[zDim-1-z] = sigma (E - O)   (sigma = -1 for d/dz).
// lane = level (matrix row in REGISTERS), coefficients broadcast from smem as 128-bit pairs,
// warp = columns; every store is a coalesced 256-byte row segment.
// parM: [3][KMAX][32] (matrix, mode, lane), zero padded.
// -------------------------------------------------------------------------------------
#define ZPAR_KMAX 44
#define ZPAR_ZS (ZPAR_KMAX + 2)

__global__ void __launch_bounds__(256, 1) k_inv_z_par(DevGrid g, const ZTile* __restrict__ tiles, int ntiles, int var0,
                                                      int nfields, const double* __restrict__ in, long long in_fs,
                                                      long long in_vs, double* __restrict__ phys,
                                                      const double* __restrict__ parM) {
  SB_DYN_SMEM(double, a);   // [nfields][32][ZPAR_ZS]
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int v = blockIdx.y;
  for (int i = tid; i < nfields * 32 * ZPAR_ZS; i += 256) a[i] = 0.0;   // the zero padding k >= bz stays zero
  const long long slotN = (long long)g.V * g.N;
  double* const pv = phys + (long long)(var0 + v) * g.N;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const ZTile zt = tiles[t];
    __syncthreads();
    for (int row = warp; row < nfields * bz; row += 8) {
      const int f = row / bz, zb = row - f * bz;
      double val = 0.0;
      if (lane < zt.ncols)
        val = in[(long long)f * in_fs + (long long)v * in_vs + zt.out_base + (long long)zb * zt.out_stride + lane];
      a[(f * 32 + lane) * ZPAR_ZS + zb] = val;
    }
    __syncthreads();
    const long long colbase = (long long)zt.hcol0 * zDim;
#pragma unroll 1
    for (int mat = 0; mat < 3; ++mat) {
      double Mr[ZPAR_KMAX];
#pragma unroll
      for (int k = 0; k < ZPAR_KMAX; ++k) Mr[k] = parM[(mat * ZPAR_KMAX + k) * 32 + lane];
      if (mat == 0) {
        // value matrix: all fields of one column at once (2*nfields independent FMA chains)
        for (int c = warp; c < zt.ncols; c += 8) {
          double E[5] = {0, 0, 0, 0, 0}, O[5] = {0, 0, 0, 0, 0};
          const double2* ap = reinterpret_cast<const double2*>(a + c * ZPAR_ZS);
#pragma unroll
          for (int k2 = 0; k2 < ZPAR_KMAX / 2; ++k2) {
#pragma unroll
            for (int f = 0; f < 5; ++f) {
              if (f < nfields) {
                const double2 x = ap[f * (32 * ZPAR_ZS / 2) + k2];
                E[f] = fma(Mr[2 * k2], x.x, E[f]);
                O[f] = fma(Mr[2 * k2 + 1], x.y, O[f]);
              }
            }
          }
          if (lane < zh) {
            double* o = pv + colbase + (long long)c * zDim;
#pragma unroll
            for (int f = 0; f < 5; ++f) {
              if (f < nfields) {
                o[f * slotN + lane] = E[f] + O[f];
                o[f * slotN + zDim - 1 - lane] = E[f] - O[f];
              }
            }
          }
        }
      } else {
        // d/dz (mat 1, sigma = -1) and d2/dz2 (mat 2) of field 0: four columns at once
        const double sigma = (mat == 1) ? -1.0 : 1.0;
        double* const oslot = pv + (long long)(nfields + mat - 1) * slotN + colbase;
        for (int c0 = warp * 4; c0 < zt.ncols; c0 += 32) {   // 8 warps x 4 columns = one 32-column tile
          double E[4] = {0, 0, 0, 0}, O[4] = {0, 0, 0, 0};
          const double2* ap = reinterpret_cast<const double2*>(a + c0 * ZPAR_ZS);
#pragma unroll
          for (int k2 = 0; k2 < ZPAR_KMAX / 2; ++k2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double2 x = ap[j * (ZPAR_ZS / 2) + k2];
              E[j] = fma(Mr[2 * k2], x.x, E[j]);
              O[j] = fma(Mr[2 * k2 + 1], x.y, O[j]);
            }
          }
          if (lane < zh) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + j < zt.ncols) {
                double* o = oslot + (long long)(c0 + j) * zDim;
                o[lane] = E[j] + O[j];
                o[zDim - 1 - lane] = sigma * (E[j] - O[j]);
              }
            }
          }
        }
      }
    }
  }
}

bool inv_z_par_ok(const DevGrid& g, int nfields) {
  return g.zDim >= 2 && g.zDim <= 64 && (g.zDim & 1) == 0 && g.bz <= ZPAR_KMAX && nfields <= 5;
}

// host: parity matrices [3][KMAX][32] from the full synthesis matrices T_k[z][q] (no BC fold)
void build_inv_z_par_tables(int zDim, int bz, const double* T0, const double* T1, const double* T2, std::vector<double>& out) {
  out.assign((size_t)3 * ZPAR_KMAX * 32, 0.0);
  const double* T[3] = {T0, T1, T2};
  for (int m = 0; m < 3; ++m)
    for (int k = 0; k < bz && k < ZPAR_KMAX; ++k)
      for (int z = 0; z < zDim / 2 && z < 32; ++z) out[((size_t)m * ZPAR_KMAX + k) * 32 + z] = T[m][(size_t)z * zDim + k];
}

void launch_inv_z_par(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                      int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                      const double* parM) {
  ProfScope prof_scope_(c, "inv_z");
  size_t smem = (size_t)nfields * 32 * ZPAR_ZS * sizeof(double);
  opt_in_smem(k_inv_z_par, smem);
  const int nsm = sb_sm_count();
  int gx = ntiles < nsm * 2 * 4 ? ntiles : nsm * 2 * 4;
  SB_LAUNCH(k_inv_z_par, dim3(gx, nvars), dim3(256), smem, c.stream, g, tiles, ntiles, var0, nfields, in, in_fstride,
            in_vstride, phys, parM);
  SB_CHECK_LAUNCH();
  count(c);
}

void launch_inv_z(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                  int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                  const double* invM) {
  ProfScope prof_scope_(c, "inv_z");
  int zp = (g.zDim + 3) & ~3;
  int ngroups = zp / 4;
  int nwarps = ngroups < 16 ? ngroups : 16;
  size_t smem = ((size_t)3 * g.bz * zp + (size_t)nfields * g.bz * 32) * sizeof(double);
  opt_in_smem(k_inv_z, smem);
  const int nsm = sb_sm_count();
  int gx = ntiles < nsm * 2 ? ntiles : nsm * 2;
  SB_LAUNCH(k_inv_z, dim3(gx, nvars), dim3(32 * nwarps), smem, c.stream, g, tiles, ntiles, var0, nfields, in,
            in_fstride, in_vstride, phys, invM);
  SB_CHECK_LAUNCH();
  count(c);
}

// =====================================================================================
// power-of-two complex FFT in shared memory (radix-4 DIF forward / DIT inverse, no bit reversal)
// =====================================================================================
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

__device__ void fft_dif(double2* buf, int log2L, int ndft, const double2* __restrict__ tw, int tid, int nthr) {
  const int L = 1 << log2L;
  int lNs = log2L;
  for (; lNs >= 2; lNs -= 2) {
    const int lNq = lNs - 2, Nq = 1 << lNq, lstep = log2L - lNs;
    const int nb = ndft << (log2L - 2);
    for (int b = tid; b < nb; b += nthr) {
      int d = b >> (log2L - 2), w = b & ((L >> 2) - 1);
      int j = w & (Nq - 1), blk = w >> lNq;
      double2* p = buf + ((size_t)d << log2L) + ((size_t)blk << lNs) + j;
      double2 a0 = p[0], a1 = p[Nq], a2 = p[2 * Nq], a3 = p[3 * Nq];
      double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
      double2 t3 = make_double2(a1.y - a3.y, -(a1.x - a3.x));
      double2 y0 = cadd(t0, t2), y1 = cadd(t1, t3), y2 = csub(t0, t2), y3 = csub(t1, t3);
      if (j) {
        int ti = j << lstep;
        y1 = cmul(y1, tw[ti]);
        y2 = cmul(y2, tw[2 * ti]);
        y3 = cmul(y3, tw[3 * ti]);
      }
      p[0] = y0; p[Nq] = y1; p[2 * Nq] = y2; p[3 * Nq] = y3;
    }
    __syncthreads();
  }
  if (lNs == 1) {
    const int nb = ndft << (log2L - 1);
    for (int b = tid; b < nb; b += nthr) {
      double2* p = buf + 2 * (size_t)b;
      double2 a0 = p[0], a1 = p[1];
      p[0] = cadd(a0, a1);
      p[1] = csub(a0, a1);
    }
    __syncthreads();
  }
}

__device__ void fft_dit(double2* buf, int log2L, int ndft, const double2* __restrict__ tw, int tid, int nthr) {
  const int L = 1 << log2L;
  int lNs = 2;
  if (log2L & 1) {
    const int nb = ndft << (log2L - 1);
    for (int b = tid; b < nb; b += nthr) {
      double2* p = buf + 2 * (size_t)b;
      double2 a0 = p[0], a1 = p[1];
      p[0] = cadd(a0, a1);
      p[1] = csub(a0, a1);
    }
    __syncthreads();
    lNs = 3;
  }
  for (; lNs <= log2L; lNs += 2) {
    const int lNq = lNs - 2, Nq = 1 << lNq, lstep = log2L - lNs;
    const int nb = ndft << (log2L - 2);
    for (int b = tid; b < nb; b += nthr) {
      int d = b >> (log2L - 2), w = b & ((L >> 2) - 1);
      int j = w & (Nq - 1), blk = w >> lNq;
      double2* p = buf + ((size_t)d << log2L) + ((size_t)blk << lNs) + j;
      double2 a0 = p[0], a1 = p[Nq], a2 = p[2 * Nq], a3 = p[3 * Nq];
      if (j) {
        int ti = j << lstep;
        a1 = cmulc(a1, tw[ti]);
        a2 = cmulc(a2, tw[2 * ti]);
        a3 = cmulc(a3, tw[3 * ti]);
      }
      double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
      double2 t3 = make_double2(-(a1.y - a3.y), a1.x - a3.x);
      p[0] = cadd(t0, t2); p[Nq] = cadd(t1, t3); p[2 * Nq] = csub(t0, t2); p[3 * Nq] = csub(t1, t3);
    }
    __syncthreads();
  }
}

// circular convolution with the pre-transformed chirp: buf <- IFFT( FFT(buf) * FH )
__device__ void bluestein_conv(double2* buf, int log2L, int ndft, const double2* __restrict__ tw,
                               const double2* __restrict__ FH, int tid, int nthr) {
  const int L = 1 << log2L;
  fft_dif(buf, log2L, ndft, tw, tid, nthr);
  for (int i = tid; i < ndft * L; i += nthr) buf[i] = cmul(buf[i], FH[i & (L - 1)]);
  __syncthreads();
  fft_dit(buf, log2L, ndft, tw, tid, nthr);
}

// =====================================================================================
// ring forward FFT: rows of n real points -> retained coefficients k = 0..ri (phase-corrected, /n)
// =====================================================================================
__global__ void __launch_bounds__(512) k_fwd_l(DevGrid g, const LWork* __restrict__ work, int log2L_arg,
                                               const double2* __restrict__ tw_arg, const RingPlan* __restrict__ plans,
                                               const double* __restrict__ blob, const double* __restrict__ in,
                                               long long in_vs, double* __restrict__ mirror, long long mirror_vs,
                                               double* __restrict__ out, long long out_vs, const SmallCls* __restrict__ cls) {
  SB_DYN_SMEM(double2, buf);
  const LWork wk = work[blockIdx.x];
  const int v = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const RingPlan pl = plans[wk.r];
  // merged launch over all small classes (cls != null): the item's ring names its class
  const int log2L = cls ? cls[pl.cls].log2L : log2L_arg;
  const double2* __restrict__ tw = cls ? cls[pl.cls].tw : tw_arg;
  const int n = pl.n, m = pl.m, L = 1 << log2L;
  const double2* chirp = reinterpret_cast<const double2*>(blob + pl.off);
  const double2* wkk = chirp + m;
  const double2* ph = wkk + m;
  const double2* FH = ph + m;
  const long long hoff = g.ring_hoff[wk.r];
  const double* src = in + (long long)v * in_vs + (long long)g.bz * hoff;
  double* mir = mirror ? mirror + (long long)v * mirror_vs + (long long)g.bz * hoff : nullptr;
  const int ndft = 2 * wk.nrows;
  // 1. load, decimate by 4, pack pairs, chirp-multiply, zero pad
  for (int i = tid; i < wk.nrows * L; i += nthr) {
    int row = i >> log2L, a = i & (L - 1);
    double2 y0 = make_double2(0.0, 0.0), y1 = y0;
    if (a < m) {
      const long long o = (long long)(wk.row0 + row) * n + 4 * a;
      const double2 x01 = *reinterpret_cast<const double2*>(src + o);
      const double2 x23 = *reinterpret_cast<const double2*>(src + o + 2);
      if (mir) {
        *reinterpret_cast<double2*>(mir + o) = x01;
        *reinterpret_cast<double2*>(mir + o + 2) = x23;
      }
      const double2 c = chirp[a];
      y0 = cmul(x01, c);
      y1 = cmul(x23, c);
    }
    buf[((size_t)(2 * row) << log2L) + a] = y0;
    buf[((size_t)(2 * row + 1) << log2L) + a] = y1;
  }
  __syncthreads();
  // 2-4. Bluestein convolution
  bluestein_conv(buf, log2L, ndft, tw, FH, tid, nthr);
  // 5. un-chirp, separate the four real sub-sequences, radix-4 combine for k < m, phase + scale
  const double invn = 1.0 / n;
  double* dst = out + (long long)v * out_vs + g.ring_woff[wk.r];
  for (int i = tid; i < wk.nrows * m; i += nthr) {
    int row = i / m, k = i - row * m;
    int km = k ? m - k : 0;
    const double2* b0 = buf + ((size_t)(2 * row) << log2L);
    const double2* b1 = b0 + L;
    double2 ck = chirp[k], ckm = chirp[km];
    double2 Y0 = cmul(b0[k], ck), Y0m = cmul(b0[km], ckm);
    double2 Y1 = cmul(b1[k], ck), Y1m = cmul(b1[km], ckm);
    // S_even = (Y + conj Ym)/2 ; S_odd = (Y - conj Ym)/(2i)
    double2 S0 = make_double2(0.5 * (Y0.x + Y0m.x), 0.5 * (Y0.y - Y0m.y));
    double2 S1 = make_double2(0.5 * (Y0.y + Y0m.y), -0.5 * (Y0.x - Y0m.x));
    double2 S2 = make_double2(0.5 * (Y1.x + Y1m.x), 0.5 * (Y1.y - Y1m.y));
    double2 S3 = make_double2(0.5 * (Y1.y + Y1m.y), -0.5 * (Y1.x - Y1m.x));
    double2 w1 = wkk[k], w2 = cmul(w1, w1), w3 = cmul(w2, w1);
    double2 X = cadd(cadd(S0, cmul(S1, w1)), cadd(cmul(S2, w2), cmul(S3, w3)));
    X = cmul(X, ph[k]);
    double* o = dst + (long long)(wk.row0 + row) * g.W;
    if (k == 0) {
      o[0] = X.x * invn;
    } else {
      o[2 * k - 1] = X.x * invn;
      o[2 * k] = X.y * invn;
    }
  }
}

int sb_rows_per_cta(int L, bool fast) {
  if (fast) {
    int log2L, nfull, rf, T, nteams, iters, nrows;
    fast_class_config(L, &log2L, &nfull, &rf, &T, &nteams, &iters, &nrows);
    return nrows;
  }
  int r = (192 * 1024) / (2 * L * 16);
  if (r > 16) r = 16;
  if (r < 1) r = 1;
  return r;
}

// items of a ring work list (sorted by ring, descending) whose ring lies in [c.r_lo, c.r_hi): a contiguous block
static void ring_block(const LaunchCtx& c, const std::vector<LWork>& w, int& i0, int& i1) {
  const int n = (int)w.size();
  if (c.r_hi < 0) { i0 = 0; i1 = n; return; }
  auto first_below = [&](int r) {      // first index whose ring is < r
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (w[mid].r < r) hi = mid; else lo = mid + 1; }
    return lo;
  };
  i0 = first_below(c.r_hi);
  i1 = first_below(c.r_lo);
}

void launch_fwd_l(const LaunchCtx& c, const DevGrid& g, const std::vector<std::vector<LWork>>& hostwork,
                  const LWork* const* work, const std::vector<FftClass>& classes, const double* const* tw,
                  const double* const* twp, const RingPlan* plans, const double* blob, int nvars, const double* in,
                  long long in_vstride, int /*in_is_z*/, double* mirror, long long mirror_vstride, double* out,
                  long long out_vstride, const std::vector<std::vector<LWork>>* hostwork2, const LWork* const* work2,
                  double* fft3_scratch) {
  ProfScope prof_scope_(c, "fwd_l");
  const bool merged = c.small && c.small->nfwork > 0 && c.r_hi < 0;   // every small (generic-kernel) class in one launch
  for (size_t ci = classes.size(); ci-- > 0;) {   // largest convolution length first
    if (merged && !classes[ci].fast && classes[ci].R != 3) continue;
    int i0, i1, j0 = 0, j1 = 0;
    ring_block(c, hostwork[ci], i0, i1);
    const int nwork = i1 - i0;
    if (hostwork2) ring_block(c, (*hostwork2)[ci], j0, j1);
    const int nwork2 = j1 - j0;
    if (!nwork) continue;
    int L = classes[ci].L;
    if (classes[ci].R == 3) {
      if (!hostwork2 || !nwork2 || !fft3_scratch) throw std::runtime_error("composite FFT class without a v2 work list");
      if (fft5_supported(L) && (uintptr_t)in % 16 == 0 && in_vstride % 2 == 0)
        launch_fwd_l5(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_vstride, mirror,
                      mirror_vstride, out, out_vstride, fft3_scratch);
      else
        launch_fwd_l3(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_vstride, mirror,
                      mirror_vstride, out, out_vstride, fft3_scratch);
      continue;
    }
    if (classes[ci].fast && nwork2 && fft4_supported(L, true) && (uintptr_t)in % 16 == 0 &&
        in_vstride % 2 == 0) {   // rows are bulk-copied: 16-byte aligned
      launch_fwd_l4(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_vstride, mirror,
                    mirror_vstride, out, out_vstride);
      continue;
    }
    if (classes[ci].fast && nwork2) {
      launch_fwd_l2(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_vstride, mirror,
                    mirror_vstride, out, out_vstride);
      continue;
    }
    if (classes[ci].fast) {
      launch_fwd_l_fast(c, g, work[ci] + i0, nwork, L, twp[ci], classes[ci].twoff, plans, blob, nvars, in, in_vstride, mirror,
                        mirror_vstride, out, out_vstride);
      continue;
    }
    int nr = sb_rows_per_cta(L, false);
    size_t smem = (size_t)2 * nr * L * 16;
    opt_in_smem(k_fwd_l, smem);
    int threads = (nr * L / 2 >= 512) ? 512 : ((nr * L / 2 >= 256) ? 256 : 128);
    SB_LAUNCH(k_fwd_l, dim3(nwork, nvars), dim3(threads), smem, c.stream, g, work[ci] + i0, classes[ci].log2L,
              reinterpret_cast<const double2*>(tw[ci]), plans, blob, in, in_vstride, mirror, mirror_vstride, out,
              out_vstride, nullptr);
    SB_CHECK_LAUNCH();
    count(c);
  }
  if (merged) {
    opt_in_smem(k_fwd_l, c.small->smem);
    SB_LAUNCH(k_fwd_l, dim3(c.small->nfwork, nvars), dim3(256), c.small->smem, c.stream, g, c.small->fwork, 0, nullptr, plans, blob,
              in, in_vstride, mirror, mirror_vstride, out, out_vstride, c.small->cls);
    SB_CHECK_LAUNCH();
    count(c);
  }
}

// =====================================================================================
// ring inverse FFT: spectra (value, d/dr, d2/dr2) -> 5 real rows (value, r, rr, lambda, lambda-lambda)
// rows of one ring are indexed rho = zb*5 + f
// =====================================================================================
__global__ void __launch_bounds__(512) k_inv_l(DevGrid g, const LWork* __restrict__ work, int log2L_arg,
                                               const double2* __restrict__ tw_arg, const RingPlan* __restrict__ plans,
                                               const double* __restrict__ blob, const double* __restrict__ in,
                                               long long in_fs, long long in_vs, double* __restrict__ out,
                                               long long out_fs, long long out_vs, int out_is_phys, int var0,
                                               unsigned lmask, const SmallCls* __restrict__ cls) {
  SB_DYN_SMEM(double2, buf);
  const LWork wk = work[blockIdx.x];
  const int v = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const RingPlan pl = plans[wk.r];
  const int log2L = cls ? cls[pl.cls].log2L : log2L_arg;
  const double2* __restrict__ tw = cls ? cls[pl.cls].tw : tw_arg;
  const int n = pl.n, m = pl.m, L = 1 << log2L;
  const double2* chirp = reinterpret_cast<const double2*>(blob + pl.off);
  const double2* wkk = chirp + m;
  const double2* ph = wkk + m;
  const double2* FH = ph + m;
  const int ndft = 2 * wk.nrows;
  const long long woff = g.ring_woff[wk.r];
  // 1. build the two Hermitian-symmetrised, packed spectra per row; feed conj(h)*chirp
  for (int i = tid; i < wk.nrows * L; i += nthr) {
    int row = i >> log2L, k = i & (L - 1);
    double2 u0 = make_double2(0.0, 0.0), u1 = u0;
    if (k < m) {
      const int rho = wk.row0 + row;
      const int zb = rho / 5, f = rho - zb * 5;
      const int fin = (f < 3) ? f : 0;
      const double* sp = in + (long long)fin * in_fs + (long long)v * in_vs + (long long)zb * g.W + woff;
      // d(q) = spectrum coefficient of exp(+i q lambda_local), q = 0..m-1, incl. derivative factor
      auto dval = [&](int q) -> double2 {
        if (q == 0) return make_double2((f < 3) ? sp[0] : 0.0, 0.0);
        double2 cq = make_double2(2.0 * sp[2 * q - 1], 2.0 * sp[2 * q]);
        cq = cmulc(cq, ph[q]);
        if (f == 3) cq = make_double2(-(double)q * cq.y, (double)q * cq.x);
        else if (f == 4) { double s = -(double)q * (double)q; cq = make_double2(s * cq.x, s * cq.y); }
        return cq;
      };
      const int km = k ? m - k : 0;
      double2 dk = dval(k), dm = dval(km);
      double2 wk1 = wkk[k], wm1 = wkk[km];
      // g_b[q] = d(q) * conj(w_q)^b ; G_b[k] = (g_b[k] + conj(g_b[m-k]))/2
      double2 gk0 = dk, gm0 = dm;
      double2 gk1 = cmulc(gk0, wk1), gm1 = cmulc(gm0, wm1);
      double2 gk2 = cmulc(gk1, wk1), gm2 = cmulc(gm1, wm1);
      double2 gk3 = cmulc(gk2, wk1), gm3 = cmulc(gm2, wm1);
      double2 G0 = make_double2(0.5 * (gk0.x + gm0.x), 0.5 * (gk0.y - gm0.y));
      double2 G1 = make_double2(0.5 * (gk1.x + gm1.x), 0.5 * (gk1.y - gm1.y));
      double2 G2 = make_double2(0.5 * (gk2.x + gm2.x), 0.5 * (gk2.y - gm2.y));
      double2 G3 = make_double2(0.5 * (gk3.x + gm3.x), 0.5 * (gk3.y - gm3.y));
      // h01 = G0 + i G1 ; input to the forward machinery is conj(h) * chirp
      double2 h01 = make_double2(G0.x - G1.y, G0.y + G1.x);
      double2 h23 = make_double2(G2.x - G3.y, G2.y + G3.x);
      const double2 c = chirp[k];
      u0 = cmul(make_double2(h01.x, -h01.y), c);
      u1 = cmul(make_double2(h23.x, -h23.y), c);
    }
    buf[((size_t)(2 * row) << log2L) + k] = u0;
    buf[((size_t)(2 * row + 1) << log2L) + k] = u1;
  }
  __syncthreads();
  bluestein_conv(buf, log2L, ndft, tw, FH, tid, nthr);
  // 3. un-chirp; conj(Y) = x_{4a} + i x_{4a+1} (first DFT), x_{4a+2} + i x_{4a+3} (second)
  const long long hoff = (out_is_phys == 2 ? g.ring_hoffp : g.ring_hoff)[wk.r];
  for (int i = tid; i < wk.nrows * m; i += nthr) {
    int row = i / m, a = i - row * m;
    const int rho = wk.row0 + row;
    const int zb = rho / 5, f = rho - zb * 5;
    if (!((lmask >> f) & 1)) continue;     // row nobody reads: not stored (its destination may not exist)
    const double2* b0 = buf + ((size_t)(2 * row) << log2L);
    const double2* b1 = b0 + L;
    const double2 c = chirp[a];
    double2 Y0 = cmul(b0[a], c), Y1 = cmul(b1[a], c);
    const RowDst o = row_dst(g, out, out_fs, out_vs, out_is_phys, f, v, var0, hoff, n, zb);
    *reinterpret_cast<double2*>(o.at(4 * a)) = make_double2(Y0.x, -Y0.y);
    *reinterpret_cast<double2*>(o.at(4 * a + 2)) = make_double2(Y1.x, -Y1.y);
  }
}

void launch_inv_l(const LaunchCtx& c, const DevGrid& g, const std::vector<std::vector<LWork>>& hostwork,
                  const LWork* const* work, const std::vector<FftClass>& classes, const double* const* tw,
                  const double* const* twp, const RingPlan* plans, const double* blob, int nvars, const double* in,
                  long long in_fstride, long long in_vstride, double* out, long long out_fstride, long long out_vstride,
                  int out_is_phys, int var0, const std::vector<std::vector<LWork>>* hostwork2, const LWork* const* work2) {
  ProfScope prof_scope_(c, "inv_l");
  const bool merged = c.small && c.small->niwork > 0 && c.r_hi < 0;
  for (size_t ci = classes.size(); ci-- > 0;) {   // largest convolution length first
    if (merged && !classes[ci].fast && classes[ci].R != 3) continue;
    int i0, i1, j0 = 0, j1 = 0;
    ring_block(c, hostwork[ci], i0, i1);
    const int nwork = i1 - i0;
    if (hostwork2) ring_block(c, (*hostwork2)[ci], j0, j1);
    const int nwork2 = j1 - j0;
    if (!nwork) continue;
    int L = classes[ci].L;
    if (classes[ci].R == 3) {
      if (!hostwork2 || !nwork2) throw std::runtime_error("composite FFT class without a v2 work list");
      if (fft5_supported(L))
        launch_inv_l5(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_fstride, in_vstride,
                      out, out_fstride, out_vstride, out_is_phys, var0);
      else
        launch_inv_l3(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_fstride, in_vstride,
                      out, out_fstride, out_vstride, out_is_phys, var0);
      continue;
    }
    if (classes[ci].fast && nwork2 && fft4_supported(L, false)) {
      launch_inv_l4(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_fstride, in_vstride,
                    out, out_fstride, out_vstride, out_is_phys, var0);
      continue;
    }
    if (classes[ci].fast && nwork2) {
      launch_inv_l2(c, g, work2[ci] + j0, nwork2, L, twp[ci], plans, blob, nvars, in, in_fstride, in_vstride,
                    out, out_fstride, out_vstride, out_is_phys, var0);
      continue;
    }
    if (classes[ci].fast) {
      launch_inv_l_fast(c, g, work[ci] + i0, nwork, L, twp[ci], classes[ci].twoff, plans, blob, nvars, in, in_fstride,
                        in_vstride, out, out_fstride, out_vstride, out_is_phys, var0);
      continue;
    }
    int nr = sb_rows_per_cta(L, false);
    size_t smem = (size_t)2 * nr * L * 16;
    opt_in_smem(k_inv_l, smem);
    int threads = (nr * L / 2 >= 512) ? 512 : ((nr * L / 2 >= 256) ? 256 : 128);
    SB_LAUNCH(k_inv_l, dim3(nwork, nvars), dim3(threads), smem, c.stream, g, work[ci] + i0, classes[ci].log2L,
              reinterpret_cast<const double2*>(tw[ci]), plans, blob, in, in_fstride, in_vstride, out, out_fstride,
              out_vstride, out_is_phys, var0, c.need.lmask, nullptr);
    SB_CHECK_LAUNCH();
    count(c);
  }
  if (merged) {
    opt_in_smem(k_inv_l, c.small->smem);
    SB_LAUNCH(k_inv_l, dim3(c.small->niwork, nvars), dim3(256), c.small->smem, c.stream, g, c.small->iwork, 0, nullptr, plans, blob,
              in, in_fstride, in_vstride, out, out_fstride, out_vstride, out_is_phys, var0, c.need.lmask, c.small->cls);
    SB_CHECK_LAUNCH();
    count(c);
  }
}

// =====================================================================================
// radial forward: b_m = sum_r w_r phi_m(r) f(r) per spline column (sliding 4-wide window)
// =====================================================================================
#define RQ 128

__global__ void __launch_bounds__(RQ) k_fwd_r(DevGrid g, const double* __restrict__ in, long long in_vs,
                                              double* __restrict__ B, long long B_vs) {
  __shared__ double tile[RQ][33];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * RQ, q = q0 + tid;
  const int zb = blockIdx.y, v = blockIdx.z;
  const bool valid = q < g.ncolp;
  const double* plane = in + (long long)v * in_vs + (long long)zb * g.W;
  double* Bv = B + (long long)v * B_vs;
  double w[3][4];
#pragma unroll
  for (int mu = 0; mu < 3; ++mu)
#pragma unroll
    for (int j = 0; j < 4; ++j) w[mu][j] = g.wq[mu] * g.phi[0][mu][j];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int M = g.b_rDim, nc = g.num_cells;
  for (int m = 0; m < M; ++m) {
    if (m < nc) {
#pragma unroll
      for (int mu = 0; mu < 3; ++mu) {
        const int r = 3 * m + mu;
        const long long wo = g.ring_woff[r];
        const int ncol_r = (int)(g.ring_woff[r + 1] - wo);
        if (q0 < ncol_r) {
          double f = (valid && q < ncol_r) ? plane[wo + q] : 0.0;
          a0 = fma(w[mu][0], f, a0); a1 = fma(w[mu][1], f, a1); a2 = fma(w[mu][2], f, a2); a3 = fma(w[mu][3], f, a3);
        }
      }
    }
    tile[tid][m & 31] = a0;
    a0 = a1; a1 = a2; a2 = a3; a3 = 0.0;
    if ((m & 31) == 31 || m == M - 1) {
      __syncthreads();
      const int m0 = m & ~31, cnt = m - m0 + 1;
      const int lane = tid & 31, warp = tid >> 5;
      for (int qq = warp; qq < RQ; qq += RQ / 32) {
        if (q0 + qq < g.ncolp && lane < cnt)
          Bv[((long long)zb * g.ncolp + q0 + qq) * M + m0 + lane] = tile[qq][lane];
      }
      __syncthreads();
    }
  }
}

// chunked version: thread = (spline column, 32 consecutive coefficients).  b_m only needs cells m-3..m, so a
// chunk re-reads 3 cells of overlap and the radial loop of 1000 dependent iterations becomes 11x more
// CTAs with 105 independent coalesced loads each; all-zero chunks (column beyond every ring's wavenumber
// range) are written without touching the input.
#define RM 32
__global__ void __launch_bounds__(RQ) k_fwd_r2(DevGrid g, int nvars, const double* __restrict__ in, long long in_vs,
                                               double* __restrict__ B, long long B_vs, PeerScatter ps, int var0) {
  __shared__ double tile[RQ][RM + 1];
  __shared__ long long s_wo[3 * (RM + 3) + 1];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * RQ, q = q0 + tid;
  const int m0 = blockIdx.y * RM;
  const int zb = blockIdx.z / nvars, v = blockIdx.z - zb * nvars;
  const int M = g.b_rDim, nc = g.num_cells;
  const int mcnt = (M - m0 < RM) ? M - m0 : RM;
  const int c_lo = (m0 - 3 > 0) ? m0 - 3 : 0;
  const int c_hi = (m0 + RM - 1 < nc - 1) ? m0 + RM - 1 : nc - 1;       // inclusive
  for (int i = tid; i <= 3 * (c_hi - c_lo + 1); i += RQ) s_wo[i] = g.ring_woff[3 * c_lo + i];
  __syncthreads();
  const double* plane = in + (long long)v * in_vs + (long long)zb * g.W;
  double* Bv = B + (long long)v * B_vs;
  const int nrings = 3 * (c_hi - c_lo + 1);
  const int ncol_max = nrings > 0 ? (int)(s_wo[nrings] - s_wo[nrings - 1]) : 0;   // widest ring of the chunk
  if (q0 < ncol_max) {
    double w[3][4];
#pragma unroll
    for (int mu = 0; mu < 3; ++mu)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[mu][j] = g.wq[mu] * g.phi[0][mu][j];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    // four cells = twelve ring values requested before the first is used (ncu: 69 % of the stall samples of the
    // one-cell-at-a-time loop were long-scoreboard waits on its three loads); same accumulation order
    constexpr int UB = 4;
    for (int cc0 = m0 - 3; cc0 < m0 + mcnt; cc0 += UB) {
      double f[UB][3];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int cc = cc0 + u;
        const bool in = cc >= c_lo && cc <= c_hi && cc < m0 + mcnt;
#pragma unroll
        for (int mu = 0; mu < 3; ++mu) {
          const int ir = in ? 3 * (cc - c_lo) + mu : 0;
          const long long wo = s_wo[ir];
          const int ncol_r = (int)(s_wo[ir + 1] - wo);
          f[u][mu] = (in && q < ncol_r) ? plane[wo + q] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int cc = cc0 + u;
        if (cc < m0 + mcnt) {
          if (cc >= c_lo && cc <= c_hi) {
#pragma unroll
            for (int mu = 0; mu < 3; ++mu) {
              a0 = fma(w[mu][0], f[u][mu], a0); a1 = fma(w[mu][1], f[u][mu], a1);
              a2 = fma(w[mu][2], f[u][mu], a2); a3 = fma(w[mu][3], f[u][mu], a3);
            }
          }
          if (cc >= m0) tile[tid][cc - m0] = a0;
          a0 = a1; a1 = a2; a2 = a3; a3 = 0.0;
        }
      }
    }
  } else {
    for (int j = 0; j < mcnt; ++j) tile[tid][j] = 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  // plane owner of this block's z-mode (peer scatter): its slot of this tile, in that GPU's memory
  double* peer = nullptr;
  if (ps.nranks > 0) {
    int k = 0;
    while (k + 1 < ps.nranks && zb >= ps.z0[k + 1]) ++k;
    peer = ps.base[k] + (long long)(var0 + v) * ps.vstride[k] + (long long)(zb - ps.z0[k]) * g.ncolp * M;
  }
  for (int qq = warp; qq < RQ; qq += RQ / 32) {
    if (q0 + qq < g.ncolp && lane < mcnt) {
      const double val = tile[qq][lane];
      Bv[((long long)zb * g.ncolp + q0 + qq) * M + m0 + lane] = val;
      if (peer) peer[(long long)(q0 + qq) * M + m0 + lane] = val;
    }
  }
}

void launch_fwd_r(const LaunchCtx& c, const DevGrid& g, int nvars, const double* in, long long in_vstride, double* B,
                  long long B_vstride, const PeerScatter* scatter, int var0) {
  ProfScope prof_scope_(c, "fwd_r");
  static const bool v1 = std::getenv("SB_RADIAL_V1") != nullptr;   // A/B switch
  if (v1 && !scatter) {
    dim3 grid((g.ncolp + RQ - 1) / RQ, g.bz, nvars);
    SB_LAUNCH(k_fwd_r, grid, dim3(RQ), 0, c.stream, g, in, in_vstride, B, B_vstride);
  } else {
    PeerScatter ps{};
    if (scatter) ps = *scatter;
    dim3 grid((g.ncolp + RQ - 1) / RQ, (g.b_rDim + RM - 1) / RM, g.bz * nvars);
    SB_LAUNCH(k_fwd_r2, grid, dim3(RQ), 0, c.stream, g, nvars, in, in_vstride, B, B_vstride, ps, var0);
  }
  SB_CHECK_LAUNCH();
  count(c);
}

// =====================================================================================
// radial inverse: evaluate value, d/dr, d2/dr2 of every spline column at the tile's mish radii
// =====================================================================================
__global__ void __launch_bounds__(RQ) k_inv_r(DevGrid t, DevGrid p, const double* __restrict__ A, long long A_vs,
                                              double* __restrict__ out, long long out_fs, long long out_vs,
                                              int out_is_phys, int var0, unsigned smask) {
  __shared__ double tile[RQ][33];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * RQ, q = q0 + tid;
  const int zb = blockIdx.y, v = blockIdx.z;
  const bool valid = q < t.ncolp;
  const double* Av = A + (long long)v * A_vs;
  const int Mt = t.b_rDim, Mp = p.b_rDim, cofs = t.coefOffset - p.coefOffset;
  const int lane = tid & 31, warp = tid >> 5;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int m = 0; m < Mt; ++m) {
    if ((m & 31) == 0) {
      __syncthreads();
      const int cnt = (Mt - m < 32) ? Mt - m : 32;
      for (int qq = warp; qq < RQ; qq += RQ / 32) {
        double val = 0.0;
        if (q0 + qq < t.ncolp && lane < cnt)
          val = Av[((long long)zb * p.ncolp + q0 + qq) * Mp + cofs + m + lane];
        tile[qq][lane] = val;
      }
      __syncthreads();
    }
    a0 = a1; a1 = a2; a2 = a3; a3 = tile[tid][m & 31];
    if (m >= 3) {
      const int c = m - 3;
#pragma unroll
      for (int mu = 0; mu < 3; ++mu) {
        const int r = 3 * c + mu;
        const long long wo = t.ring_woff[r];
        const int ncol_r = (int)(t.ring_woff[r + 1] - wo);
        if (valid && q < ncol_r) {
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (!((smask >> d) & 1)) continue;
            double s = t.phi[d][mu][0] * a0;
            s = fma(t.phi[d][mu][1], a1, s);
            s = fma(t.phi[d][mu][2], a2, s);
            s = fma(t.phi[d][mu][3], a3, s);
            if (out_is_phys)
              out[((long long)d * t.V + var0 + v) * t.N + r] = s;
            else
              out[(long long)d * out_fs + (long long)v * out_vs + (long long)zb * t.W + wo + q] = s;
          }
        }
      }
    }
  }
}

// chunked version: thread = (spline column, 32 consecutive cells); the 35 coefficients it needs arrive
// through a transposed shared-memory tile, the 9 outputs per cell leave coalesced along the ring.
__global__ void __launch_bounds__(RQ) k_inv_r2(DevGrid t, DevGrid p, int nvars, const double* __restrict__ A, long long A_vs,
                                               double* __restrict__ out, long long out_fs, long long out_vs,
                                               int out_is_phys, int var0, unsigned smask) {
  __shared__ double tile[RQ][RM + 5];          // odd row stride: the per-thread column walk is conflict-free
  __shared__ long long s_wo[3 * RM + 1];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * RQ, q = q0 + tid;
  const int c0 = blockIdx.y * RM;              // first cell of the chunk
  const int zb = blockIdx.z / nvars, v = blockIdx.z - zb * nvars;
  const int nc = t.num_cells, Mp = p.b_rDim, cofs = t.coefOffset - p.coefOffset;
  const int ccnt = (nc - c0 < RM) ? nc - c0 : RM;
  for (int i = tid; i <= 3 * ccnt; i += RQ) s_wo[i] = t.ring_woff[3 * c0 + i];
  __syncthreads();
  const int ncol_max = (int)(s_wo[3 * ccnt] - s_wo[3 * ccnt - 1]);
  if (q0 >= ncol_max) return;                  // no ring of this chunk carries these columns
  const double* Av = A + (long long)v * A_vs;
  const int lane = tid & 31, warp = tid >> 5;
  const int need = ccnt + 3;
  // eight columns per round: sixteen loads in flight per warp (ncu: 57 % of the stall samples of the one-column-at-a-time
  // loop were long-scoreboard waits on these two loads)
  for (int qq0 = warp; qq0 < RQ; qq0 += 8 * (RQ / 32)) {
    double x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int qq = qq0 + u * (RQ / 32);
      const bool ok = q0 + qq < t.ncolp;
      const double* src = Av + ((long long)zb * p.ncolp + q0 + qq) * Mp + cofs + c0;
      x[u] = (ok && lane < need) ? src[lane] : 0.0;
      y[u] = (lane < 4 && ok && 32 + lane < need) ? src[32 + lane] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int qq = qq0 + u * (RQ / 32);
      tile[qq][lane] = x[u];
      if (lane < 4) tile[qq][32 + lane] = y[u];
    }
  }
  __syncthreads();
  if (q >= t.ncolp) return;
  double a0 = tile[tid][0], a1 = tile[tid][1], a2 = tile[tid][2], a3;
  for (int c = 0; c < ccnt; ++c) {
    a3 = tile[tid][c + 3];
#pragma unroll
    for (int mu = 0; mu < 3; ++mu) {
      const long long wo = s_wo[3 * c + mu];
      const int ncol_r = (int)(s_wo[3 * c + mu + 1] - wo);
      if (q < ncol_r) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (!((smask >> d) & 1)) continue;
          double s = t.phi[d][mu][0] * a0;
          s = fma(t.phi[d][mu][1], a1, s);
          s = fma(t.phi[d][mu][2], a2, s);
          s = fma(t.phi[d][mu][3], a3, s);
          if (out_is_phys)
            out[((long long)d * t.V + var0 + v) * t.N + 3 * (c0 + c) + mu] = s;
          else
            out[(long long)d * out_fs + (long long)v * out_vs + (long long)zb * t.W + wo + q] = s;
        }
      }
    }
    a0 = a1; a1 = a2; a2 = a3;
  }
}

void launch_inv_r(const LaunchCtx& c, const DevGrid& tile, const DevGrid& patch, int nvars, const double* A,
                  long long A_vstride, double* out, long long out_fstride, long long out_vstride, int out_is_phys,
                  int var0) {
  ProfScope prof_scope_(c, "inv_r");
  static const bool v1 = std::getenv("SB_RADIAL_V1") != nullptr;   // A/B switch
  if (v1) {
    dim3 grid((tile.ncolp + RQ - 1) / RQ, tile.bz, nvars);
    SB_LAUNCH(k_inv_r, grid, dim3(RQ), 0, c.stream, tile, patch, A, A_vstride, out, out_fstride, out_vstride,
              out_is_phys, var0, c.need.smask);
  } else {
    dim3 grid((tile.ncolp + RQ - 1) / RQ, (tile.num_cells + RM - 1) / RM, tile.bz * nvars);
    SB_LAUNCH(k_inv_r2, grid, dim3(RQ), 0, c.stream, tile, patch, nvars, A, A_vstride, out, out_fstride, out_vstride,
              out_is_phys, var0, c.need.smask);
  }
  SB_CHECK_LAUNCH();
  count(c);
}

// =====================================================================================
// K2: a = Gamma^T (Gamma (P+Q) Gamma^T)^-1 Gamma b  for every spline column
// =====================================================================================
__global__ void k_spline_solve(DevSplineFactor f, int ncols, int qc, const double* __restrict__ B,
                               double* __restrict__ A) {
  SB_DYN_SMEM(double, sm);
  const int M = f.M, Ms = M | 1, n = f.nfree;
  double* chol = sm;                    // [n][4]
  double* x = sm + (size_t)4 * n;       // [qc][Ms]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const long long c0 = (long long)blockIdx.x * qc;
  const int nq = (int)((ncols - c0 < qc) ? ncols - c0 : qc);
  for (int i = tid; i < 4 * n; i += nthr) chol[i] = f.chol[i];
  for (int i = tid; i < nq * M; i += nthr) {
    int cq = i / M, mm = i - cq * M;
    x[cq * Ms + mm] = B[c0 * M + i];
  }
  __syncthreads();
  if (tid < nq) {
    double* b = x + tid * Ms;
    const int rL = f.rL, rR = f.rR;
    // fold: b~ = Gamma b (in place on the free range b[rL .. rL+n-1])
    if (rL == 1) { b[1] += f.foldL[0] * b[0]; b[2] += f.foldL[1] * b[0]; }
    else if (rL == 2) { b[2] += f.foldL[0] * b[0] + f.foldL[1] * b[1]; }
    if (rR == 1) { b[M - 2] += f.foldR[0] * b[M - 1]; b[M - 3] += f.foldR[1] * b[M - 1]; }
    else if (rR == 2) { b[M - 3] += f.foldR[0] * b[M - 1] + f.foldR[1] * b[M - 2]; }
    double* y = b + rL;
    // forward substitution L y = b~
    double y1 = 0, y2 = 0, y3 = 0;
    for (int i = 0; i < n; ++i) {
      const double* l = chol + 4 * i;
      double s = y[i];
      s = fma(-l[1], y1, s); s = fma(-l[2], y2, s); s = fma(-l[3], y3, s);
      s *= l[0];
      y[i] = s;
      y3 = y2; y2 = y1; y1 = s;
    }
    // back substitution L^T x = y
    double x1 = 0, x2 = 0, x3 = 0;
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      if (i + 1 < n) s = fma(-chol[4 * (i + 1) + 1], x1, s);
      if (i + 2 < n) s = fma(-chol[4 * (i + 2) + 2], x2, s);
      if (i + 3 < n) s = fma(-chol[4 * (i + 3) + 3], x3, s);
      s *= chol[4 * i];
      y[i] = s;
      x3 = x2; x2 = x1; x1 = s;
    }
    // unfold: a = Gamma^T a~
    if (rL == 1) b[0] = f.foldL[0] * b[1] + f.foldL[1] * b[2];
    else if (rL == 2) { b[0] = f.foldL[0] * b[2]; b[1] = f.foldL[1] * b[2]; }
    else if (rL == 3) { b[0] = b[1] = b[2] = 0.0; }
    if (rR == 1) b[M - 1] = f.foldR[0] * b[M - 2] + f.foldR[1] * b[M - 3];
    else if (rR == 2) { b[M - 1] = f.foldR[0] * b[M - 3]; b[M - 2] = f.foldR[1] * b[M - 3]; }
    else if (rR == 3) { b[M - 1] = b[M - 2] = b[M - 3] = 0.0; }
  }
  __syncthreads();
  for (int i = tid; i < nq * M; i += nthr) {
    int cq = i / M, mm = i - cq * M;
    A[c0 * M + i] = x[cq * Ms + mm];
  }
}

// streaming version: thread = column, but only a [128 columns][32 coefficients] tile lives in shared memory.
// Forward substitution walks the tiles upward writing y to A, back substitution walks them downward
// (A is read back: 4S instead of 2S of traffic, but 10x more columns in flight per SM than the
// whole-column kernel above, whose 172 KB of shared memory allowed 64 columns per SM).
#define SSQ 128
__global__ void __launch_bounds__(SSQ, 4) k_spline_solve2(const DevSplineFactor* __restrict__ fs, int ncols, int chol_in_smem,
                                                       const double* __restrict__ B, double* __restrict__ A, long long vstride) {
  __shared__ double tile[SSQ][33];
  SB_DYN_SMEM(double, s_chol);
  const DevSplineFactor f = fs[blockIdx.y];        // one launch covers every variable (blockIdx.y), each with its own BCs
  if (f.periodic || f.nfree < 3) return;           // handled by k_spline_dense / k_spline_solve
  B += (long long)blockIdx.y * vstride;
  A += (long long)blockIdx.y * vstride;
  const int M = f.M, n = f.nfree, rL = f.rL, rR = f.rR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c0 = (long long)blockIdx.x * SSQ;
  const int nq = (int)((ncols - c0 < SSQ) ? ncols - c0 : SSQ);
  const bool valid = tid < nq;
  const double* chol = f.chol;
  if (chol_in_smem) {
    for (int i = tid; i < 4 * n; i += SSQ) s_chol[i] = f.chol[i];
    chol = s_chol;
  }
  const int ntile = (M + 31) / 32;
  const int last = M - 1 - rR;                 // last free index
  double bM1 = 0.0, bM2 = 0.0, bl0 = 0.0, bl1 = 0.0;
  if (valid) {
    bM1 = B[(c0 + tid) * M + M - 1];
    if (M >= 2) bM2 = B[(c0 + tid) * M + M - 2];
  }
  // ---------------- fold + forward substitution  L y = Gamma b
  // Software pipeline: the warp's 32 columns of the NEXT tile are requested into registers (32 loads in flight per thread)
  // before the substitution of the current tile starts, so HBM works while the dependent FMA chains run (ncu: 2.8 TB/s with
  // load, substitution and store phases taking turns).
  constexpr int NPW = SSQ / (SSQ / 32);          // columns per warp (qq = warp + u * SSQ/32)
  double xr[NPW];
  auto fetch = [&](const double* __restrict__ src, int t) {
    const int m0 = t * 32, cnt = (M - m0 < 32) ? M - m0 : 32;
#pragma unroll
    for (int u = 0; u < NPW; ++u) {
      const int qq = warp + u * (SSQ / 32);
      xr[u] = (qq < nq && lane < cnt) ? src[(c0 + qq) * M + m0 + lane] : 0.0;
    }
  };
  auto park = [&](int t) {
    const int m0 = t * 32, cnt = (M - m0 < 32) ? M - m0 : 32;
#pragma unroll
    for (int u = 0; u < NPW; ++u) {
      const int qq = warp + u * (SSQ / 32);
      if (qq < nq && lane < cnt) tile[qq][lane] = xr[u];
    }
  };
  double y1 = 0, y2 = 0, y3 = 0;
  fetch(B, 0);
  for (int t = 0; t < ntile; ++t) {
    const int m0 = t * 32, cnt = (M - m0 < 32) ? M - m0 : 32;
    __syncthreads();
    park(t);
    __syncthreads();
    if (t + 1 < ntile) fetch(B, t + 1);
    if (valid) {
      for (int j = 0; j < cnt; ++j) {
        const int i = m0 + j;
        double s = tile[tid][j];
        if (i < rL) { if (i == 0) bl0 = s; else if (i == 1) bl1 = s; continue; }
        if (i > last) continue;
        const int fi = i - rL;
        if (rL == 1) { if (fi == 0) s += f.foldL[0] * bl0; else if (fi == 1) s += f.foldL[1] * bl0; }
        else if (rL == 2) { if (fi == 0) s += f.foldL[0] * bl0 + f.foldL[1] * bl1; }
        if (rR == 1) { if (i == M - 2) s += f.foldR[0] * bM1; else if (i == M - 3) s += f.foldR[1] * bM1; }
        else if (rR == 2) { if (i == M - 3) s += f.foldR[0] * bM1 + f.foldR[1] * bM2; }
        const double* l = chol + 4 * fi;
        s = fma(-l[1], y1, s); s = fma(-l[2], y2, s); s = fma(-l[3], y3, s);
        s *= l[0];
        tile[tid][j] = s;
        y3 = y2; y2 = y1; y1 = s;
      }
    }
    __syncthreads();
    for (int qq = warp; qq < SSQ; qq += SSQ / 32)
      if (qq < nq && lane < cnt) A[(c0 + qq) * M + m0 + lane] = tile[qq][lane];
  }
  // ---------------- back substitution  L^T x = y, unfold a = Gamma^T x
  double x1 = 0, x2 = 0, x3 = 0, xa = 0, xb = 0;      // xa = x[n-1], xb = x[n-2]
  // (the forward sweep's last tile is still in shared memory: the first fetch of the way down re-reads what this block
  //  itself has just stored -- program order per thread, same addresses)
  __syncthreads();
  fetch(A, ntile - 1);
  for (int t = ntile - 1; t >= 0; --t) {
    const int m0 = t * 32, cnt = (M - m0 < 32) ? M - m0 : 32;
    __syncthreads();
    park(t);
    __syncthreads();
    if (t > 0) fetch(A, t - 1);
    if (valid) {
      for (int j = cnt - 1; j >= 0; --j) {
        const int i = m0 + j;
        if (i > last || i < rL) continue;
        const int fi = i - rL;
        double s = tile[tid][j];
        if (fi + 1 < n) s = fma(-chol[4 * (fi + 1) + 1], x1, s);
        if (fi + 2 < n) s = fma(-chol[4 * (fi + 2) + 2], x2, s);
        if (fi + 3 < n) s = fma(-chol[4 * (fi + 3) + 3], x3, s);
        s *= chol[4 * fi];
        tile[tid][j] = s;
        if (fi == n - 1) xa = s;
        if (fi == n - 2) xb = s;
        x3 = x2; x2 = x1; x1 = s;
      }
      if (t == 0) {   // x1 = x[0], x2 = x[1] here
        if (rL == 1) tile[tid][0] = f.foldL[0] * x1 + f.foldL[1] * x2;
        else if (rL == 2) { tile[tid][0] = f.foldL[0] * x1; tile[tid][1] = f.foldL[1] * x1; }
        else if (rL == 3) { tile[tid][0] = 0.0; tile[tid][1] = 0.0; tile[tid][2] = 0.0; }
      }
    }
    __syncthreads();
    for (int qq = warp; qq < SSQ; qq += SSQ / 32)
      if (qq < nq && lane < cnt) A[(c0 + qq) * M + m0 + lane] = tile[qq][lane];
  }
  __syncthreads();
  if (valid) {        // right-edge unfold: the tiles above stored stale values there
    double* a = A + (c0 + tid) * M;
    if (rR == 1) a[M - 1] = f.foldR[0] * xa + f.foldR[1] * xb;
    else if (rR == 2) { a[M - 1] = f.foldR[0] * xa; a[M - 2] = f.foldR[1] * xa; }
    else if (rR == 3) { a[M - 1] = 0.0; a[M - 2] = 0.0; a[M - 3] = 0.0; }
  }
}

__global__ void k_spline_dense(DevSplineFactor f, int ncols, const double* __restrict__ B, double* __restrict__ A) {
  const int M = f.M;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)ncols * M) return;
  const int col = (int)(idx / M), mm = (int)(idx - (long long)col * M);
  const double* b = B + (long long)col * M;
  const double* trow = f.dense + (size_t)mm * M;
  double s = 0.0;
  for (int j = 0; j < M; ++j) s = fma(trow[j], b[j], s);
  A[idx] = s;
}

void launch_spline_solve(const LaunchCtx& c, const DevGrid& g, const DevSplineFactor* dfactors,
                         const std::vector<DevSplineFactor>& hf, const double* B, double* A) {
  ProfScope prof_scope_(c, "spline_solve");
  const int ncols = g.bz * g.ncolp;
  static const bool v1 = std::getenv("SB_RADIAL_V1") != nullptr;
  // streaming kernel: all variables with banded (non-periodic) factors in ONE launch
  bool merged = false;
  if (!v1 && dfactors) {
    int nmax = 0;
    for (int v = 0; v < g.V; ++v)
      if (!hf[v].periodic && hf[v].nfree >= 3) nmax = std::max(nmax, hf[v].nfree);
    if (nmax > 0) {
      const int in_smem = (size_t)4 * nmax * 8 <= 96 * 1024;
      size_t smem = in_smem ? (size_t)4 * nmax * 8 : 0;
      if (smem > 8 * 1024) {   // static tile (33 KB) + table may exceed the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(k_spline_solve2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
      }
      SB_LAUNCH(k_spline_solve2, dim3((ncols + SSQ - 1) / SSQ, g.V), dim3(SSQ), smem, c.stream, dfactors, ncols, in_smem, B, A,
                g.S);
      SB_CHECK_LAUNCH();
      count(c);
      merged = true;
    }
  }
  for (int v = 0; v < g.V; ++v) {
    const DevSplineFactor& f = hf[v];
    const double* Bv = B + (long long)v * g.S;
    double* Av = A + (long long)v * g.S;
    if (f.periodic) {
      long long tot = (long long)ncols * f.M;
      SB_LAUNCH(k_spline_dense, dim3((unsigned)((tot + 127) / 128)), dim3(128), 0, c.stream, f, ncols, Bv, Av);
    } else if (merged && f.nfree >= 3) {
      continue;
    } else {
      int Ms = f.M | 1;
      int qc = 64;
      while (qc > 1 && ((size_t)qc * Ms + 4 * (size_t)f.nfree) * 8 > 200 * 1024) qc >>= 1;
      if (qc > ncols) { qc = 1; while (qc * 2 <= ncols) qc *= 2; }
      size_t smem = ((size_t)qc * Ms + 4 * (size_t)f.nfree) * 8;
      opt_in_smem(k_spline_solve, smem);
      int threads = qc < 128 ? 128 : qc;
      SB_LAUNCH(k_spline_solve, dim3((ncols + qc - 1) / qc), dim3(threads), smem, c.stream, f, ncols, qc, Bv, Av);
    }
    SB_CHECK_LAUNCH();
    count(c);
  }
}

// =====================================================================================
// shared-spectral assembly (own block assigned, halo added), copy, NaN scan
// =====================================================================================
__global__ void k_assemble(DevGrid p, DevGrid t, const double* __restrict__ tileB, int has_prev, DevGrid pv,
                           const double* __restrict__ prevB, double* __restrict__ shared) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_v = t.S;
  if (idx >= per_v * t.V) return;
  const int v = (int)(idx / per_v);
  long long rem = idx - (long long)v * per_v;
  const int m = (int)(rem % t.b_rDim);
  rem /= t.b_rDim;
  const int pcol = (int)(rem % t.ncolp), zb = (int)(rem / t.ncolp);
  double val = tileB[idx];
  if (has_prev && m < 3 && pcol < pv.ncolp)
    val += prevB[(long long)v * pv.S + ((long long)zb * pv.ncolp + pcol) * pv.b_rDim + pv.b_rDim - 3 + m];
  shared[(long long)v * p.S + ((long long)zb * p.ncolp + pcol) * p.b_rDim + (t.coefOffset - p.coefOffset) + m] = val;
}

void launch_assemble(const LaunchCtx& c, const DevGrid& patch, const DevGrid& tile, const double* tileB,
                     const DevGrid* prev, const double* prevB, int /*last*/, double* shared) {
  ProfScope prof_scope_(c, "assemble");
  long long tot = tile.S * tile.V;
  SB_LAUNCH(k_assemble, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, c.stream, patch, tile, tileB,
            prev ? 1 : 0, prev ? *prev : tile, prevB, shared);
  SB_CHECK_LAUNCH();
  count(c);
}

// inverse of k_assemble without the halo: the coefficients tile t evaluates, cut out of the solved patch planes
__global__ void k_extract(DevGrid p, DevGrid t, const double* __restrict__ A, double* __restrict__ tileA, long long dst_vs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_v = t.S;
  if (idx >= per_v * t.V) return;
  const int v = (int)(idx / per_v);
  long long rem = idx - (long long)v * per_v;
  const int m = (int)(rem % t.b_rDim);
  rem /= t.b_rDim;
  const int pcol = (int)(rem % t.ncolp), zb = (int)(rem / t.ncolp);
  // dst_vs != 0: straight into the tile's own A (possibly another GPU's memory), whose variables are S_tile apart
  const long long o = dst_vs ? (long long)v * dst_vs + (idx - (long long)v * per_v) : idx;
  tileA[o] = A[(long long)v * p.S + ((long long)zb * p.ncolp + pcol) * p.b_rDim + (t.coefOffset - p.coefOffset) + m];
}

void launch_extract(const LaunchCtx& c, const DevGrid& patch, const DevGrid& tile, const double* A, double* tileA,
                    long long dst_vstride) {
  ProfScope prof_scope_(c, "extract");
  long long tot = tile.S * tile.V;
  SB_LAUNCH(k_extract, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, c.stream, patch, tile, A, tileA, dst_vstride);
  SB_CHECK_LAUNCH();
  count(c);
}

// batched column operator (Chebyshev column API): out[r][c] = sum_k M[r][k] in[k][c] + C0, column-major batches
__global__ void k_column_op(const double* __restrict__ M, int rows, int cols, const double* __restrict__ in,
                            double* __restrict__ out, long long ncols, double C0) {
  SB_DYN_SMEM(double, x);                // [columns of this block][cols]
  const int r = threadIdx.x, cl = threadIdx.y;
  const long long c = (long long)blockIdx.x * blockDim.y + cl;
  for (int k = r; k < cols; k += blockDim.x)
    x[cl * cols + k] = c < ncols ? in[c * cols + k] : 0.0;
  __syncthreads();
  if (c >= ncols) return;
  for (int rr = r; rr < rows; rr += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < cols; ++k) s = fma(M[(size_t)rr * cols + k], x[cl * cols + k], s);
    out[c * rows + rr] = s + C0;
  }
}

void launch_column_op(const LaunchCtx& c, const double* M, int rows, int cols, const double* in, double* out,
                      long long ncols, double C0) {
  ProfScope prof_scope_(c, "column_op");
  const int tx = 32, ty = 8;
  SB_LAUNCH(k_column_op, dim3((unsigned)((ncols + ty - 1) / ty)), dim3(tx, ty), (size_t)ty * cols * sizeof(double), c.stream,
            M, rows, cols, in, out, ncols, C0);
  SB_CHECK_LAUNCH();
  count(c);
}

__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

void launch_copy(const LaunchCtx& c, double* dst, const double* src, long long n) {
  ProfScope prof_scope_(c, "copy");
  long long blocks = (n + 255) / 256;
  if (blocks > sb_sm_count() * 16) blocks = sb_sm_count() * 16;
  if (blocks < 1) blocks = 1;
  SB_LAUNCH(k_copy, dim3((unsigned)blocks), dim3(256), 0, c.stream, dst, src, n);
  SB_CHECK_LAUNCH();
  count(c);
}

// result[0] = smallest flat index (v*N + i) holding a NaN, or LLONG_MAX
__global__ void k_nan_scan(const double* __restrict__ phys, long long total, long long* result) {
  long long best = 0x7fffffffffffffffLL;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    double x = phys[i];
    if (x != x && i < best) best = i;
  }
  if (best != 0x7fffffffffffffffLL) {
#ifdef SB_EMU
    static std::atomic_flag lk = ATOMIC_FLAG_INIT;
    while (lk.test_and_set()) {}
    if (best < *result) *result = best;
    lk.clear();
#else
    atomicMin(reinterpret_cast<long long*>(result), best);
#endif
  }
}

void launch_nan_scan(const LaunchCtx& c, const double* phys, long long N, int V, long long* result) {
  ProfScope prof_scope_(c, "nan_scan");
  long long total = N * V;
  long long blocks = (total + 255) / 256;
  if (blocks > sb_sm_count() * 8) blocks = sb_sm_count() * 8;
  if (blocks < 1) blocks = 1;
  SB_LAUNCH(k_nan_scan, dim3((unsigned)blocks), dim3(256), 0, c.stream, phys, total, result);
  SB_CHECK_LAUNCH();
  count(c);
}

}  // namespace sb
