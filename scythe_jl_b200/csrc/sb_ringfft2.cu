// sb_ringfft2.cu -- persistent, table-resident Bluestein ring FFT (convolution lengths L = 256 .. 8192).
//
// Same mathematics as sb_ringfft.cu (ring of n = 4m points -> two packed complex DFT_m per real row,
// each a chirp-z convolution of power-of-two length L >= 2m-1 done as radix-16 passes with 16 complex
// values per thread), re-organised around what ncu showed on B200: the v1 kernel ran the FP64 pipe at
// 33 % because every pass waited on an L2 round trip for its twiddles / chirp / FH tables (one CTA per
// row, tables re-fetched each time) and both sequences of a CTA met at every __syncthreads.  Here
//   * one persistent 512-thread CTA per SM walks a list of (ring, row-range) items; the class twiddles,
//     the ring's pre-transformed chirp FH and the chirp itself live in shared memory for the whole item;
//   * teams of T = L/16 threads are decoupled: they meet only on their own named barrier (bar.sync id, T);
//   * L is a template parameter, so every pass stride / padded index is an immediate;
//   * the zero-padded upper half of the Bluestein input and the unused upper half of its output are pruned
//     from the first forward and the last inverse radix-16 pass;
//   * the spectrum-side prologue (Hermitian symmetrisation, derivative factor, ring phase, 4-way
//     decimation weights, chirp) is two complex multiplies against tables pre-combined on the host
//     (sb_tables.cpp: P_h, Q_h), and so is the forward epilogue (A0..A3);  both work straight from
//     registers, without the staging pass through shared memory.
#include "sb_internal.hpp"
#include "sb_fftcore.hpp"

#include <cstdlib>
#include <stdexcept>

namespace sb {

template <int LOG2L>
struct RCfg {
  static constexpr int L = 1 << LOG2L;
  static constexpr int T = L / 16;                       // threads per team (one complex sequence)
  static constexpr int NFULL = (LOG2L - 1) / 4;          // strided radix-16 passes
  static constexpr int RF = 1 << (LOG2L - 4 * NFULL);    // register-local final radix
  static constexpr int LP = L + L / 16;                  // padded team buffer (complex)
  static constexpr int NT = 512;
  static constexpr int NTEAMS = NT / T;
  static constexpr int TW0 = 15 * T;                     // pass-0 twiddles (complex)
  static constexpr int TWR = (NFULL >= 2 ? 15 * (L >> 8) : 0) + (NFULL >= 3 ? 15 * (L >> 12) : 0);
  static constexpr int TWRP = (TWR + 1) & ~1;
  static constexpr bool FH_SMEM = LOG2L <= 12;
  static constexpr bool CH_SMEM = LOG2L <= 11;
#ifndef SB_TW0_SMEM_MAX
#define SB_TW0_SMEM_MAX 11
#endif
  static constexpr bool TW0_SMEM = LOG2L <= SB_TW0_SMEM_MAX;
  static constexpr size_t SMEM = sizeof(double2) * ((size_t)NTEAMS * LP + TWRP + (TW0_SMEM ? TW0 : 0) + (FH_SMEM ? L : 0) +
                                                   (CH_SMEM ? L / 2 : 0));
};

// One prefetch.global.L2 per thread for the NEXT sequence's input while the current convolution runs: the input rows
// come from DRAM (the SZ / SL scratch does not fit L2) and nothing else requests them early -- every register is taken
// (128 x 512) and so is the shared memory, so the row cannot be staged; 25 % of k_fwd_l2's stall samples were
// long-scoreboard waits on these loads (profiles/r1m_ncu_full_k_fwd_l2.txt).  A hint only: results cannot change.
#ifndef SB_FFT_L2PF
#define SB_FFT_L2PF 1
#endif

#ifndef SB_PRO_BATCH
#define SB_PRO_BATCH 2
#endif
constexpr int PB = SB_PRO_BATCH;   // spectrum elements whose loads are in flight together in the inverse prologue

template <int T>
__device__ __forceinline__ void team_sync(int team) {
  if (T >= 64) sb_bar_sync(1 + team, T);
  else __syncwarp();
}
template <int T>
__device__ __forceinline__ void pair_sync(int pair) {
  if (2 * T >= 64) sb_bar_sync((T >= 64 ? 9 : 1) + pair, 2 * T);
  else __syncwarp();
}

// w[k] = w1^k, k = 1..15, by a depth-4 product tree: trades the 15 shared-memory twiddle loads of the second
// strided pass for 14 complex multiplies.  Measured on B200 (C4): inverse kernel -2.5 %, forward kernel +1 %
// (shared-memory wavefronts, not FP64 issue, are the tighter resource of the inverse kernel), so only the inverse
// kernel enables it (TW1C); doing the same for the first pass loses (19.4 vs 18.8 ms).
__device__ __forceinline__ void twiddle_powers(double2 w1, double2 (&w)[16]) {
  w[1] = w1;
  w[2] = cm(w1, w1); w[3] = cm(w[2], w1);
  w[4] = cm(w[2], w[2]); w[5] = cm(w[4], w1); w[6] = cm(w[4], w[2]); w[7] = cm(w[4], w[3]);
  w[8] = cm(w[4], w[4]); w[9] = cm(w[8], w1); w[10] = cm(w[8], w[2]); w[11] = cm(w[8], w[3]);
  w[12] = cm(w[8], w[4]); w[13] = cm(w[8], w[5]); w[14] = cm(w[8], w[6]); w[15] = cm(w[8], w[7]);
}

// circular convolution with the pre-transformed chirp.  In: v[n1] = element n1*T + tl, n1 < 8 (upper half zero).
// Out: v[n1] = element n1*T + tl of the result for n1 < 8 (the upper half is not produced).
template <int LOG2L, bool PRUNE = true, bool TW1C = false>
__device__ __forceinline__ void conv2(double2 (&v)[16], double2* buf, const double2* __restrict__ tw0,
                                      const double2* __restrict__ twr, const double2* __restrict__ FHt, int tl, int team,
                                      bool active) {
  typedef RCfg<LOG2L> C;
  constexpr int L = C::L, T = C::T;
  // ---- forward strided passes
#pragma unroll
  for (int p = 0; p < C::NFULL; ++p) {
    const int Ms = L >> (4 * (p + 1)), Lb = Ms << 4;
    const int b = tl / Ms, j = tl - b * Ms, base = padi(b * Lb + j);
    if (p > 0) {
      team_sync<T>(team);
      if (active) {
        // butterfly order (0,4,8,12, 1,5,9,13, ...): the first radix-4 can start after four loads
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = (i >> 2) + 4 * (i & 3);
          v[n] = buf[base + n * Ms + ((n * Ms) >> 4)];
        }
      }
    }
    if (active) {
      if (p == 0 && PRUNE) fft16_fwd_lo8(v); else fft16<false>(v);
      const double2* tw = (p == 0 ? tw0 : twr + (p == 1 ? 0 : 15 * (L >> 8))) + j;
      if (TW1C && p == 1) {
        double2 w[16];
        twiddle_powers(tw[0], w);
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cm(v[k], w[k]);
      } else {
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = cm(v[k], tw[(k - 1) * Ms]);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) buf[base + k * Ms + ((k * Ms) >> 4)] = v[k];
    }
  }
  team_sync<T>(team);
  // ---- final forward pass, pointwise product, first inverse pass: all in registers
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = buf[tl * 17 + e];
    fft_final<false>(v, C::RF);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = cm(v[e], FHt[e * T + tl]);
    fft_final<true>(v, C::RF);
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[tl * 17 + e] = v[e];
  }
  // ---- inverse strided passes (reverse order)
#pragma unroll
  for (int p = C::NFULL - 1; p >= 0; --p) {
    const int Ms = L >> (4 * (p + 1)), Lb = Ms << 4;
    const int b = tl / Ms, j = tl - b * Ms, base = padi(b * Lb + j);
    team_sync<T>(team);
    if (active) {
      const double2* tw = (p == 0 ? tw0 : twr + (p == 1 ? 0 : 15 * (L >> 8))) + j;
      if (TW1C && p == 1) {
        double2 w[16];
        twiddle_powers(tw[0], w);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const double2 x = buf[base + k * Ms + ((k * Ms) >> 4)];
          v[k] = k ? cmc(x, w[k]) : x;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {       // butterfly order, data and twiddle fetched together
          const int k = (i >> 2) + 4 * (i & 3);
          const double2 x = buf[base + k * Ms + ((k * Ms) >> 4)];
          v[k] = k ? cmc(x, tw[(k - 1) * Ms]) : x;
        }
      }
      fft16<true>(v);
      if (p > 0) {
#pragma unroll
        for (int n = 0; n < 16; ++n) buf[base + n * Ms + ((n * Ms) >> 4)] = v[n];
      }
    }
  }
}

// class tables -> shared memory (once per CTA); returns the pointers the passes use
template <int LOG2L>
__device__ __forceinline__ void load_class_tables(double2* sm, const double2* __restrict__ twp, const double2*& tw0,
                                                  const double2*& twr, double2*& s_FH, double2*& s_ch) {
  typedef RCfg<LOG2L> C;
  double2* s_twr = sm + (size_t)C::NTEAMS * C::LP;
  double2* s_tw0 = s_twr + C::TWRP;
  s_FH = s_tw0 + (C::TW0_SMEM ? C::TW0 : 0);
  s_ch = s_FH + (C::FH_SMEM ? C::L : 0);
  for (int i = threadIdx.x; i < C::TWR; i += C::NT) s_twr[i] = twp[C::TW0 + i];
  if (C::TW0_SMEM)
    for (int i = threadIdx.x; i < C::TW0; i += C::NT) s_tw0[i] = twp[i];
  tw0 = C::TW0_SMEM ? s_tw0 : twp;
  twr = s_twr;
}

// =====================================================================================
// inverse: spectra (value, d/dr, d2/dr2) -> 5 real rows; rows of a ring are rho = zb*5 + f
// =====================================================================================
template <int LOG2L>
__global__ void __launch_bounds__(512, 1) k_inv_l2(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                   const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                   const double* __restrict__ blob, const double* __restrict__ in,
                                                   long long in_fs, long long in_vs, double* __restrict__ out,
                                                   long long out_fs, long long out_vs, int out_is_phys, int var0,
                                                   unsigned lmask) {
  typedef RCfg<LOG2L> C;
  constexpr int L = C::L, T = C::T;
  SB_DYN_SMEM(double2, sm);
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  double2* const buf = sm + (size_t)team * C::LP;
  const double2 *tw0, *twr;
  double2 *s_FH, *s_ch;
  load_class_tables<LOG2L>(sm, twp, tw0, twr, s_FH, s_ch);
  int cur_ring = -1;
  const int total = nwork * nvars;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* PQ = reinterpret_cast<const double2*>(blob + pl.off2);
    if (wk.r != cur_ring) {          // ring tables -> shared memory
      __syncthreads();
      if (C::FH_SMEM)
        for (int i = tid; i < L; i += C::NT) s_FH[i] = FH_g[i];
      if (C::CH_SMEM)
        for (int i = tid; i < m; i += C::NT) s_ch[i] = chirp_g[i];
      __syncthreads();
      cur_ring = wk.r;
    }
    const double2* FHt = C::FH_SMEM ? s_FH : FH_g;
    const double2* chirp = C::CH_SMEM ? s_ch : chirp_g;
    const long long woff = g.ring_woff[wk.r], hoff = (out_is_phys == 2 ? g.ring_hoffp : g.ring_hoff)[wk.r];
    const int nseq = 2 * wk.nrows;
    for (int s = team; s - team < nseq; s += C::NTEAMS) {
      const int row = s >> 1, half = s & 1;
      const int rho = wk.row0 + row;
      const int zb = rho / 5, f = rho - zb * 5;
      // a row the equation set does not read is skipped.  Teams of >= 16 threads: the two teams of a warp work on the two
      // halves of ONE row, so the decision is warp-uniform and the trip is left before any barrier.  Smaller teams (L <= 128)
      // share a warp with teams on OTHER rows: they stay in step (the barriers below are __syncwarp) and merely go inactive.
      const bool wanted = ((lmask >> f) & 1) != 0;
      if (T >= 16 && !wanted) continue;
      const bool active = s < nseq && wanted;
      double2 v[16];
      team_sync<T>(team);            // the team's previous sequence has finished reading buf
      if (SB_FFT_L2PF && s + C::NTEAMS < nseq) {   // the team's next spectrum row: (2m-1) doubles <= T lines of 128 B
        const int rho2 = wk.row0 + ((s + C::NTEAMS) >> 1), zb2 = rho2 / 5, f2 = rho2 - zb2 * 5;
        const double* sp2 = in + (long long)(f2 < 3 ? f2 : 0) * in_fs + (long long)v_ * in_vs + (long long)zb2 * g.W + woff;
        if (tl * 16 < 2 * m) sb_prefetch_l2(sp2 + tl * 16);
      }
      if (active) {
        const int fin = (f < 3) ? f : 0;
        const double* sp = in + (long long)fin * in_fs + (long long)v_ * in_vs + (long long)zb * g.W + woff;
        const double2* Ph = PQ + (size_t)(2 * half) * m;
        const double2* Qh = Ph + m;
        // loads are unconditional (clamped indices) and issued four elements at a time, so that one L2
        // round trip covers a whole batch instead of one per `k < m` branch
#pragma unroll
        for (int h4 = 0; h4 < 8 / PB; ++h4) {
          double cx[PB], cy[PB], qx[PB], qy[PB];
          double2 Pk[PB], Qk[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int k0 = (h4 * PB + j) * T + tl;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            cx[j] = sp[k ? 2 * k - 1 : 0]; cy[j] = sp[2 * k];
            qx[j] = sp[km ? 2 * km - 1 : 0]; qy[j] = sp[2 * km];
            Pk[j] = Ph[k]; Qk[j] = Qh[k];
          }
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int k0 = (h4 * PB + j) * T + tl;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            const double2 ck = make_double2(cx[j], k ? cy[j] : 0.0);
            const double2 cq = make_double2(qx[j], km ? qy[j] : 0.0);
            // derivative factor D(q): 1 | i q | -q^2 ;  X = conj(c_k D_k), Y = c_km D_km
            double2 X, Y;
            if (f < 3) {
              X = make_double2(ck.x, -ck.y);
              Y = cq;
            } else if (f == 3) {
              const double dk = (double)k, dq = (double)km;
              X = make_double2(-dk * ck.y, -dk * ck.x);
              Y = make_double2(-dq * cq.y, dq * cq.x);
            } else {
              const double sk = -(double)k * (double)k, sq = -(double)km * (double)km;
              X = make_double2(sk * ck.x, -sk * ck.y);
              Y = make_double2(sq * cq.x, sq * cq.y);
            }
            const double2 u = cm(X, Pk[j]) + cm(Y, Qk[j]);
            v[h4 * PB + j] = k0 < m ? u : make_double2(0.0, 0.0);
          }
        }
      }
      conv2<LOG2L, true, true>(v, buf, tw0, twr, FHt, tl, team, active);
      if (active) {
        const RowDst orow = row_dst(g, out, out_fs, out_vs, out_is_phys, f, v_, var0, hoff, n, zb);
        double* const o0 = orow.at(4 * tl + 2 * half);
        const long long ostep = orow.step(T);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a = n1 * T + tl;
          if (a < m) {
            const double2 Y = cm(v[n1], chirp[a]);
            // (teams of two threads, L = 32: consecutive outputs are 8 points apart -- no constant stride in the blocked layout)
            double* const dst = (T % 4 == 0) ? o0 + n1 * ostep : orow.at(4 * a + 2 * half);
            *reinterpret_cast<double2*>(dst) = make_double2(Y.x, -Y.y);
          }
        }
      }
    }
  }
}

// =====================================================================================
// forward: real ring rows -> retained coefficients k = 0..ri.  Teams work in pairs (the two packed
// sequences of one row); the raw convolution outputs are parked in the teams' own buffers and the
// pair combines them with the pre-combined tables A0..A3.
// =====================================================================================
template <int LOG2L>
__global__ void __launch_bounds__(512, 1) k_fwd_l2(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                   const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                   const double* __restrict__ blob, const double* __restrict__ in,
                                                   long long in_vs, double* __restrict__ mirror, long long mirror_vs,
                                                   double* __restrict__ out, long long out_vs) {
  typedef RCfg<LOG2L> C;
  constexpr int L = C::L, T = C::T, NPAIRS = C::NTEAMS / 2;
  SB_DYN_SMEM(double2, sm);
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int pair = team >> 1, half = team & 1;
  double2* const buf = sm + (size_t)team * C::LP;
  const double2* const buf0 = sm + (size_t)(2 * pair) * C::LP;
  const double2* const buf1 = buf0 + C::LP;
  const double2 *tw0, *twr;
  double2 *s_FH, *s_ch;
  load_class_tables<LOG2L>(sm, twp, tw0, twr, s_FH, s_ch);
  int cur_ring = -1;
  const int total = nwork * nvars;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* AF = reinterpret_cast<const double2*>(blob + pl.off2) + (size_t)4 * m;
    if (wk.r != cur_ring) {
      __syncthreads();
      if (C::FH_SMEM)
        for (int i = tid; i < L; i += C::NT) s_FH[i] = FH_g[i];
      if (C::CH_SMEM)
        for (int i = tid; i < m; i += C::NT) s_ch[i] = chirp_g[i];
      __syncthreads();
      cur_ring = wk.r;
    }
    const double2* FHt = C::FH_SMEM ? s_FH : FH_g;
    const double2* chirp = C::CH_SMEM ? s_ch : chirp_g;
    const long long hoff = g.ring_hoff[wk.r];
    const double* src = in + (long long)v_ * in_vs + (long long)g.bz * hoff;
    double* mir = mirror ? mirror + (long long)v_ * mirror_vs + (long long)g.bz * hoff : nullptr;
    double* dst = out + (long long)v_ * out_vs + g.ring_woff[wk.r];
    for (int row = pair; row - pair < wk.nrows; row += NPAIRS) {
      const bool active = row < wk.nrows;
      double2 v[16];
      pair_sync<T>(pair);            // the pair has finished combining the previous row out of buf0 / buf1
      if (SB_FFT_L2PF && row + NPAIRS < wk.nrows) {   // the pair's next row: n doubles <= 2T lines of 128 B
        const int ln = half * T + tl;
        if (ln * 16 < n) sb_prefetch_l2(src + (long long)(wk.row0 + row + NPAIRS) * n + ln * 16);
      }
      if (active) {
        const double* rp = src + (long long)(wk.row0 + row) * n + 2 * half;
        double* mp = mir ? mir + (long long)(wk.row0 + row) * n + 2 * half : nullptr;
        double2 x[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {   // unconditional (clamped) loads: one round trip for the whole row slice
          const int a0 = n1 * T + tl, a = a0 < m ? a0 : m - 1;
          x[n1] = *reinterpret_cast<const double2*>(rp + 4 * a);
        }
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a0 = n1 * T + tl, a = a0 < m ? a0 : m - 1;
          if (mp && a0 < m) *reinterpret_cast<double2*>(mp + 4 * a) = x[n1];
          const double2 y = cm(x[n1], chirp[a]);
          v[n1] = a0 < m ? y : make_double2(0.0, 0.0);
        }
      }
      conv2<LOG2L>(v, buf, tw0, twr, FHt, tl, team, active);
      team_sync<T>(team);            // every thread of the team has read its last-pass inputs
      if (active) {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a = n1 * T + tl;
          if (a < m) buf[a] = v[n1];
        }
      }
      pair_sync<T>(pair);
      if (active) {
        const double invn = 1.0;      // 1/n is folded into A0..A3
        (void)invn;
        double* o = dst + (long long)(wk.row0 + row) * g.W;
        for (int k = half * T + tl; k < m; k += 2 * T) {
          const int km = k ? m - k : 0;
          const double2 b0 = buf0[k], b0m = buf0[km], b1 = buf1[k], b1m = buf1[km];
          double2 X = cm(b0, AF[k]) + cm(make_double2(b0m.x, -b0m.y), AF[m + k]);
          X = X + cm(b1, AF[2 * m + k]) + cm(make_double2(b1m.x, -b1m.y), AF[3 * m + k]);
          if (k == 0) {
            o[0] = X.x;
          } else {
            o[2 * k - 1] = X.x;
            o[2 * k] = X.y;
          }
        }
      }
    }
  }
}


// =====================================================================================
// composite convolution lengths L = 3 * L2 (L2 = 512, 1024, 2048) for rings with L2 < m <= 1.5 L2, which a
// power of two would pad to 4 * L2.  Radix-3 decimation in frequency of the zero-padded input (its last
// third is empty) gives three sequences  s_r[j] = (x[j] + w3^r x[j+L2]) W_L^{jr},  r = 0,1,2,  each
// convolved by one team of L2/16 threads with the r-th third of the pre-transformed chirp; the three results
// recombine as  y[j + s L2] = sum_r w3^{-rs} W_L^{-jr} c_r[j].  A group of three teams shares the prologue
// (one third of x each), meets on its own named barrier, and splits the outputs the same way.
// =====================================================================================
#define W3R (-0.5)
#define W3I (-0.86602540378443864676)   // w3 = exp(-2 pi i / 3)

template <int LOG2L2>
struct R3Cfg {
  typedef RCfg<LOG2L2> B;
  static constexpr int L2 = B::L, T = B::T;
  static constexpr int NGROUPS = (512 / T) / 3;
  static constexpr int NTEAMS = 3 * NGROUPS;
  static constexpr int NT = NTEAMS * T;
  static constexpr int XT = 2 * T + 32;                    // TWB1[T] TWB2[T] C1[16] C2[16]
  static constexpr bool FH_SMEM = LOG2L2 <= 10;
  static constexpr int CH = 3 * L2 / 2;                    // chirp entries (m <= 1.5 L2)
  static constexpr size_t SMEM = sizeof(double2) * ((size_t)NTEAMS * B::LP + B::TWRP + B::TW0 + XT + (FH_SMEM ? 3 * L2 : 0) + CH);
};

template <int T>
__device__ __forceinline__ void group_sync(int grp) {
  sb_bar_sync((T >= 64 ? 9 : 1) + grp, 3 * T);
}

// state shared by the two composite kernels: smem carve-up + class tables
template <int LOG2L2>
struct R3Ctx {
  typedef R3Cfg<LOG2L2> C;
  double2 *bufs, *s_twr, *s_tw0, *s_x, *s_FH, *s_ch;
  __device__ __forceinline__ void init(double2* sm, const double2* __restrict__ twp) {
    bufs = sm;
    s_twr = sm + (size_t)C::NTEAMS * C::B::LP;
    s_tw0 = s_twr + C::B::TWRP;
    s_x = s_tw0 + C::B::TW0;
    s_FH = s_x + C::XT;
    s_ch = s_FH + (C::FH_SMEM ? 3 * C::L2 : 0);
    for (int i = threadIdx.x; i < C::B::TWR; i += C::NT) s_twr[i] = twp[C::B::TW0 + i];
    for (int i = threadIdx.x; i < C::B::TW0; i += C::NT) s_tw0[i] = twp[i];
    for (int i = threadIdx.x; i < C::XT; i += C::NT) s_x[i] = twp[C::B::TW0 + C::B::TWR + i];
  }
  // W_L^{j r} for j = n1*T + tl (r = 1, 2)
  __device__ __forceinline__ double2 twid(int r, int n1, int tl) const {
    return cm(s_x[(r - 1) * C::T + tl], s_x[2 * C::T + (r - 1) * 16 + n1]);
  }
};

// steps shared by forward and inverse: x thirds are in the group's buffers -> s_r -> convolution -> c_r' back in the buffers
template <int LOG2L2, bool TW1C = false>
__device__ __forceinline__ void conv3(const R3Ctx<LOG2L2>& cx, double2* const (&gb)[3], const double2* __restrict__ FHt,
                                      int r, int grp, int team, int tl, bool active) {
  typedef R3Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  double2 v[16];
  if (active) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = n1 * T + tl;
      double2 x0 = (n1 < 8) ? gb[0][j] : gb[1][j - L2 / 2];
      if (n1 < 8) {
        const double2 x1 = gb[2][j];
        if (r == 0) x0 = x0 + x1;
        else if (r == 1) x0 = x0 + make_double2(W3R * x1.x - W3I * x1.y, W3R * x1.y + W3I * x1.x);
        else x0 = x0 + make_double2(W3R * x1.x + W3I * x1.y, W3R * x1.y - W3I * x1.x);       // w3^2 = conj(w3)
      }
      v[n1] = r ? cm(x0, cx.twid(r, n1, tl)) : x0;
    }
  }
  group_sync<T>(grp);              // every team has gathered its s_r: the buffers may be overwritten
  conv2<LOG2L2, false, TW1C>(v, gb[r], cx.s_tw0, cx.s_twr, FHt + (size_t)r * L2, tl, team, active);
  team_sync<T>(team);              // the team's last-pass loads are done
  if (active) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const double2 c = r ? cmc(v[n1], cx.twid(r, n1, tl)) : v[n1];
      gb[r][n1 * T + tl] = c;
    }
  }
  group_sync<T>(grp);              // c_0, c_1', c_2' are in the three buffers
}

// this team's eight outputs: r = 0: a in [0, L2/2), r = 1: [L2/2, L2), r = 2: [L2, 3 L2/2)
template <int LOG2L2>
__device__ __forceinline__ double2 combine3(double2* const (&gb)[3], int r, int i) {
  const double2 b0 = gb[0][i], b1 = gb[1][i], b2 = gb[2][i];
  if (r < 2) return b0 + b1 + b2;
  // s = 1: w3^-1 b1 + w3^-2 b2 = conj(w3) b1 + w3 b2
  return b0 + make_double2(W3R * b1.x + W3I * b1.y, W3R * b1.y - W3I * b1.x) +
         make_double2(W3R * b2.x - W3I * b2.y, W3R * b2.y + W3I * b2.x);
}

template <int LOG2L2>
__global__ void __launch_bounds__(R3Cfg<LOG2L2>::NT, 1) k_inv_l3(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                                const double2* __restrict__ twp,
                                                                const RingPlan* __restrict__ plans,
                                                                const double* __restrict__ blob,
                                                                const double* __restrict__ in, long long in_fs,
                                                                long long in_vs, double* __restrict__ out, long long out_fs,
                                                                long long out_vs, int out_is_phys, int var0, unsigned lmask) {
  typedef R3Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  SB_DYN_SMEM(double2, sm);
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int grp = team / 3, r = team - 3 * grp;
  R3Ctx<LOG2L2> cx;
  cx.init(sm, twp);
  double2* const gb[3] = {cx.bufs + (size_t)(3 * grp) * C::B::LP, cx.bufs + (size_t)(3 * grp + 1) * C::B::LP,
                          cx.bufs + (size_t)(3 * grp + 2) * C::B::LP};
  int cur_ring = -1;
  const int total = nwork * nvars;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* PQ = reinterpret_cast<const double2*>(blob + pl.off2);
    if (wk.r != cur_ring) {
      __syncthreads();
      if (C::FH_SMEM)
        for (int i = tid; i < 3 * L2; i += C::NT) cx.s_FH[i] = FH_g[i];
      for (int i = tid; i < m; i += C::NT) cx.s_ch[i] = chirp_g[i];
      __syncthreads();
      cur_ring = wk.r;
    }
    const double2* FHt = C::FH_SMEM ? cx.s_FH : FH_g;
    const long long woff = g.ring_woff[wk.r], hoff = (out_is_phys == 2 ? g.ring_hoffp : g.ring_hoff)[wk.r];
    const int nseq = 2 * wk.nrows;
    for (int s = grp; s - grp < nseq; s += C::NGROUPS) {
      const bool active = s < nseq;
      const int row = s >> 1, half = s & 1;
      const int rho = wk.row0 + row;
      const int zb = rho / 5, f = rho - zb * 5;
      if (!((lmask >> f) & 1)) continue;   // group-uniform
      group_sync<T>(grp);            // the group's previous outputs have been read out of the buffers
      if (active) {                  // prologue: this team's third of x (see k_inv_l2 for the formula)
        const int fin = (f < 3) ? f : 0;
        const double* sp = in + (long long)fin * in_fs + (long long)v_ * in_vs + (long long)zb * g.W + woff;
        const double2* Ph = PQ + (size_t)(2 * half) * m;
        const double2* Qh = Ph + m;
#pragma unroll
        for (int h4 = 0; h4 < 8 / PB; ++h4) {
          double cx_[PB], cy_[PB], qx_[PB], qy_[PB];
          double2 Pk[PB], Qk[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int k0 = r * (L2 / 2) + (h4 * PB + j) * T + tl;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            cx_[j] = sp[k ? 2 * k - 1 : 0]; cy_[j] = sp[2 * k];
            qx_[j] = sp[km ? 2 * km - 1 : 0]; qy_[j] = sp[2 * km];
            Pk[j] = Ph[k]; Qk[j] = Qh[k];
          }
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int i = (h4 * PB + j) * T + tl;
            const int k0 = r * (L2 / 2) + i;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            const double2 ck = make_double2(cx_[j], k ? cy_[j] : 0.0);
            const double2 cq = make_double2(qx_[j], km ? qy_[j] : 0.0);
            double2 X, Y;
            if (f < 3) {
              X = make_double2(ck.x, -ck.y);
              Y = cq;
            } else if (f == 3) {
              const double dk = (double)k, dq = (double)km;
              X = make_double2(-dk * ck.y, -dk * ck.x);
              Y = make_double2(-dq * cq.y, dq * cq.x);
            } else {
              const double sk = -(double)k * (double)k, sq = -(double)km * (double)km;
              X = make_double2(sk * ck.x, -sk * ck.y);
              Y = make_double2(sq * cq.x, sq * cq.y);
            }
            const double2 u = cm(X, Pk[j]) + cm(Y, Qk[j]);
            gb[r][i] = k0 < m ? u : make_double2(0.0, 0.0);
          }
        }
      }
      group_sync<T>(grp);            // x is complete
      conv3<LOG2L2, true>(cx, gb, FHt, r, grp, team, tl, active);
      if (active) {
        const RowDst orow = row_dst(g, out, out_fs, out_vs, out_is_phys, f, v_, var0, hoff, n, zb);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int i = (r == 1 ? L2 / 2 : 0) + n1 * T + tl;
          const int a = (r == 2 ? L2 : 0) + i;
          if (a < m) {
            const double2 Y = cm(combine3<LOG2L2>(gb, r, i), cx.s_ch[a]);
            *reinterpret_cast<double2*>(orow.at(4 * a + 2 * half)) = make_double2(Y.x, -Y.y);
          }
        }
      }
    }
  }
}

// forward: both packed sequences of a row go through the group one after the other; their raw convolution
// outputs are parked in a per-group global scratch line (L2-resident) and combined with A0..A3 at the end.
template <int LOG2L2>
__global__ void __launch_bounds__(R3Cfg<LOG2L2>::NT, 1) k_fwd_l3(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                                const double2* __restrict__ twp,
                                                                const RingPlan* __restrict__ plans,
                                                                const double* __restrict__ blob,
                                                                const double* __restrict__ in, long long in_vs,
                                                                double* __restrict__ mirror, long long mirror_vs,
                                                                double* __restrict__ out, long long out_vs, double2* scratch) {
  typedef R3Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  SB_DYN_SMEM(double2, sm);
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int grp = team / 3, r = team - 3 * grp;
  R3Ctx<LOG2L2> cx;
  cx.init(sm, twp);
  double2* const gb[3] = {cx.bufs + (size_t)(3 * grp) * C::B::LP, cx.bufs + (size_t)(3 * grp + 1) * C::B::LP,
                          cx.bufs + (size_t)(3 * grp + 2) * C::B::LP};
  double2* const park = scratch + ((size_t)blockIdx.x * C::NGROUPS + grp) * (2 * C::CH);   // [2 halves][CH]
  int cur_ring = -1;
  const int total = nwork * nvars;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* AF = reinterpret_cast<const double2*>(blob + pl.off2) + (size_t)4 * m;
    if (wk.r != cur_ring) {
      __syncthreads();
      if (C::FH_SMEM)
        for (int i = tid; i < 3 * L2; i += C::NT) cx.s_FH[i] = FH_g[i];
      for (int i = tid; i < m; i += C::NT) cx.s_ch[i] = chirp_g[i];
      __syncthreads();
      cur_ring = wk.r;
    }
    const double2* FHt = C::FH_SMEM ? cx.s_FH : FH_g;
    const long long hoff = g.ring_hoff[wk.r];
    const double* src = in + (long long)v_ * in_vs + (long long)g.bz * hoff;
    double* mir = mirror ? mirror + (long long)v_ * mirror_vs + (long long)g.bz * hoff : nullptr;
    double* dst = out + (long long)v_ * out_vs + g.ring_woff[wk.r];
    for (int row = grp; row - grp < wk.nrows; row += C::NGROUPS) {
      const bool active = row < wk.nrows;
      for (int half = 0; half < 2; ++half) {
        group_sync<T>(grp);
        if (active) {
          const double* rp = src + (long long)(wk.row0 + row) * n + 2 * half;
          double* mp = mir ? mir + (long long)(wk.row0 + row) * n + 2 * half : nullptr;
          double2 x[8];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) {
            const int a0 = r * (L2 / 2) + n1 * T + tl, a = a0 < m ? a0 : m - 1;
            x[n1] = *reinterpret_cast<const double2*>(rp + 4 * a);
          }
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) {
            const int i = n1 * T + tl;
            const int a0 = r * (L2 / 2) + i, a = a0 < m ? a0 : m - 1;
            if (mp && a0 < m) *reinterpret_cast<double2*>(mp + 4 * a) = x[n1];
            const double2 y = cm(x[n1], cx.s_ch[a]);
            gb[r][i] = a0 < m ? y : make_double2(0.0, 0.0);
          }
        }
        group_sync<T>(grp);
        conv3<LOG2L2>(cx, gb, FHt, r, grp, team, tl, active);
        if (active) {
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) {
            const int i = (r == 1 ? L2 / 2 : 0) + n1 * T + tl;
            const int a = (r == 2 ? L2 : 0) + i;
            if (a < m) park[(size_t)half * C::CH + a] = combine3<LOG2L2>(gb, r, i);
          }
        }
      }
      group_sync<T>(grp);            // both halves are parked (global writes of the group are ordered by the barrier)
      if (active) {
        double* o = dst + (long long)(wk.row0 + row) * g.W;
        const double2* p0 = park;
        const double2* p1 = park + C::CH;
        for (int k = r * T + tl; k < m; k += 3 * T) {
          const int km = k ? m - k : 0;
          const double2 b0 = p0[k], b0m = p0[km], b1 = p1[k], b1m = p1[km];
          double2 X = cm(b0, AF[k]) + cm(make_double2(b0m.x, -b0m.y), AF[m + k]);
          X = X + cm(b1, AF[2 * m + k]) + cm(make_double2(b1m.x, -b1m.y), AF[3 * m + k]);
          if (k == 0) {
            o[0] = X.x;
          } else {
            o[2 * k - 1] = X.x;
            o[2 * k] = X.y;
          }
        }
      }
    }
  }
}

// =====================================================================================
// launchers
// =====================================================================================
bool fft2_supported(int L, bool forward) {
  static const char* env = std::getenv("SB_FFT");
  if (env && std::string(env) == "v1") return false;   // A/B switch
  if (L % 3 == 0) return L == 1536 || L == 3072 || L == 6144;
  // L = 32 .. 128 stay in the v1 register kernel (k_*_l_fast): these kernels with sub-warp teams of 2 .. 8 threads (RCfg<5..7>)
  // were instantiated, passed every test and measured 1 % slower on those classes (inv_l 6.61 vs 6.55 ms at C4)
  return L >= 256 && L <= (forward ? 4096 : 8192);
}

int fft2_rows_per_item(int L, bool forward) {
  // experiment switch: SB_FFT_ITEM_ROWS_INV / _FWD = rows per work item for every class
  static const int env_i = std::getenv("SB_FFT_ITEM_ROWS_INV") ? std::atoi(std::getenv("SB_FFT_ITEM_ROWS_INV")) : 0;
  static const int env_f = std::getenv("SB_FFT_ITEM_ROWS_FWD") ? std::atoi(std::getenv("SB_FFT_ITEM_ROWS_FWD")) : 0;
  if (!forward && env_i > 0) return env_i;
  if (forward && env_f > 0) return env_f;
  if (L % 3 == 0) {               // composite: a group of three teams per sequence
    const int ngroups = (512 / (L / 48)) / 3;
    return forward ? 4 * ngroups : 8 * (ngroups > 1 ? ngroups / 2 + (ngroups & 1) : 1);
  }
  const int T = L / 16, nteams = 512 / T;
  const int wave = forward ? (nteams / 2) : (nteams > 1 ? nteams / 2 : 1);   // rows in flight per CTA
  // rows per item: 11 / 27 per wave of rows in flight, at most 22 / 54 -- at L <= 2048 two items of 22 rows per ring forward and four of 54
  // inverse (43 z-modes x 1 | 5 rows); measured at C4 against 16 / 32 (three / seven uneven items): fwd_l 2.66 -> 2.62 ms,
  // inv_l 6.56 -> 6.48 ms; 11 / 108 rows are slower again.  The caller evens the items of a ring out (sb_api.cpp).
  const int w = wave < 1 ? 1 : (wave > 2 ? 2 : wave);      // (the sweep ran 22 / 54 rows for EVERY class: the shorter classes like them too)
  return (forward ? 11 : 27) * w;
}

template <int LOG2L>
static void launch_inv2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs,
                        long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  const size_t smem = RCfg<LOG2L>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(k_inv_l2<LOG2L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, gx = total < sb_sm_count() ? total : sb_sm_count();
  SB_LAUNCH(k_inv_l2<LOG2L>, dim3(gx), dim3(512), smem, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0,
            c.need.lmask);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_l2 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_inv_l2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
#define SB_INV2(LG)                                                                                                    \
  case (1 << LG):                                                                                                      \
    launch_inv2<LG>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); \
    break;
  switch (L) {
    SB_INV2(8) SB_INV2(9) SB_INV2(10) SB_INV2(11) SB_INV2(12) SB_INV2(13)
    default: throw std::runtime_error("launch_inv_l2: unsupported convolution length");
  }
#undef SB_INV2
}

template <int LOG2L>
static void launch_fwd2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs,
                        double* mirror, long long mirror_vs, double* out, long long out_vs) {
  const size_t smem = RCfg<LOG2L>::SMEM;
  cudaError_t e = cudaFuncSetAttribute(k_fwd_l2<LOG2L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, gx = total < sb_sm_count() ? total : sb_sm_count();
  SB_LAUNCH(k_fwd_l2<LOG2L>, dim3(gx), dim3(512), smem, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_vs, mirror, mirror_vs, out, out_vs);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_l2 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_fwd_l2(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs) {
#define SB_FWD2(LG)                                                                                             \
  case (1 << LG):                                                                                               \
    launch_fwd2<LG>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs);     \
    break;
  switch (L) {
    SB_FWD2(8) SB_FWD2(9) SB_FWD2(10) SB_FWD2(11) SB_FWD2(12)
    default: throw std::runtime_error("launch_fwd_l2: unsupported convolution length");
  }
#undef SB_FWD2
}


// ---- composite lengths
bool fft3_enabled() {
  static const char* e3 = std::getenv("SB_FFT3");
  static const char* e = std::getenv("SB_FFT");
  if (e && std::string(e) == "v1") return false;
  return !(e3 && std::string(e3) == "0");
}

// TWB1[T] = W_L^tl, TWB2[T] = W_L^(2 tl), C1[16] = W_L^(T n1), C2[16] = W_L^(2 T n1)   (L = 3 L2, T = L2/16)
void fft3_class_tables(int L2, std::vector<double>& tab) {
  const int T = L2 / 16, L = 3 * L2;
  const long double PI = 3.14159265358979323846264338327950288L;
  tab.assign((size_t)2 * (2 * T + 32), 0.0);
  auto put = [&](size_t i, long long e) {
    const long double ang = -2.0L * PI * (long double)(e % L) / (long double)L;
    tab[2 * i] = (double)cosl(ang); tab[2 * i + 1] = (double)sinl(ang);
  };
  for (int tl = 0; tl < T; ++tl) { put(tl, tl); put(T + tl, 2LL * tl); }
  for (int n1 = 0; n1 < 16; ++n1) { put(2 * T + n1, (long long)T * n1); put(2 * T + 16 + n1, 2LL * T * n1); }
}

size_t fft3_scratch_doubles(int L) {
  const int L2 = L / 3, T = L2 / 16, ngroups = (512 / T) / 3;
  return (size_t)2 * sb_sm_count() * ngroups * 2 * (3 * L2 / 2);
}

template <int LOG2L2>
static void launch_inv3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs,
                        long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  typedef R3Cfg<LOG2L2> C;
  cudaError_t e = cudaFuncSetAttribute(k_inv_l3<LOG2L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, gx = total < sb_sm_count() ? total : sb_sm_count();
  SB_LAUNCH(k_inv_l3<LOG2L2>, dim3(gx), dim3(C::NT), C::SMEM, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0,
            c.need.lmask);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_l3 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_inv_l3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  switch (L / 3) {
    case 512: launch_inv3<9>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 1024: launch_inv3<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 2048: launch_inv3<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    default: throw std::runtime_error("launch_inv_l3: unsupported convolution length");
  }
}

template <int LOG2L2>
static void launch_fwd3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs,
                        double* mirror, long long mirror_vs, double* out, long long out_vs, double* scratch) {
  typedef R3Cfg<LOG2L2> C;
  cudaError_t e = cudaFuncSetAttribute(k_fwd_l3<LOG2L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, gx = total < sb_sm_count() ? total : sb_sm_count();
  SB_LAUNCH(k_fwd_l3<LOG2L2>, dim3(gx), dim3(C::NT), C::SMEM, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_vs, mirror, mirror_vs, out, out_vs,
            reinterpret_cast<double2*>(scratch));
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_l3 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_fwd_l3(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs, double* scratch) {
  switch (L / 3) {
    case 512: launch_fwd3<9>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs, scratch); break;
    case 1024: launch_fwd3<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs, scratch); break;
    case 2048: launch_fwd3<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs, scratch); break;
    default: throw std::runtime_error("launch_fwd_l3: unsupported convolution length");
  }
}

}  // namespace sb
