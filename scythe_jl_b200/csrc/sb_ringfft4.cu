// sb_ringfft4.cu -- ring FFT v4 (inverse): the v2 Bluestein convolution (sb_ringfft2.cu) with Blackwell data movement.
//
// What ncu and the probes under profiles/microbench/ showed about v2 (k_inv_l2, 42 % of the C4 step):
//   * the kernel is co-limited by the FP64 pipe and the shared-memory pipe (both ~50 % busy, the round takes about the
//     SUM of the two), with 21 % of the stall samples on the global loads of the prologue (spectrum row + the ring's
//     P/Q tables) -- and neither a register nor a byte of shared memory was left to stage anything ahead;
//   * 54 of the 180 16-byte shared-memory accesses per thread and sequence are TABLE reads whose index depends on the
//     thread only (pass twiddles W_L^{k tl}, the pre-transformed chirp FH[e T + tl], the chirp, P/Q): thread-private data;
//   * Tensor Memory (256 KB per SM, unused by a non-MMA kernel) is exactly per-thread-private storage: with the 32x32b
//     shape of tcgen05.ld / tcgen05.st, thread l of warp w owns lane 32 (w % 4) + l and 512 32-bit columns of it.  Probe
//     (profiles/microbench/tmem_probe.cu, fp64_lds_probe.cu): >= 570 B/cycle/SM of TMEM reads, and a radix-16 pass whose
//     twiddles come from TMEM costs 2860 cycles against 3100 with the twiddles in shared memory (2815 with none at all).
// So here
//   * every per-thread table lives in Tensor Memory (class twiddles for the whole kernel; FH, chirp and P/Q per ring),
//     read with tcgen05.ld (SASS LDTM) in batches of four complex values;
//   * the 78 KB of shared memory this frees hold one staging row per team, filled by the bulk-copy engine
//     (cp.async.bulk + mbarrier, SASS UBLKCP / SYNCS): the NEXT sequence's spectrum row is requested by one thread right
//     after the first exchange barrier of the current convolution and lands while the FP64 pipe works; no register, no
//     LSU instruction, no long-scoreboard stall in the prologue;
//   * the second-pass twiddles come from their table again (v2 rebuilt them with 14 complex multiplies per pass to spare
//     shared-memory wavefronts).
// Same arithmetic as v2 wherever a value is produced (butterflies, table values, operation order), so the result is
// bit-identical to v2 except for the second-pass twiddles (table instead of product tree: <= 1 ulp apart).
//
// Lengths: L = 512 .. 4096 (two strided radix-16 passes + a register-local final pass).  L = 4096 has teams of 256
// threads = two threads per TMEM lane: two table sets of 224 columns, and P/Q stay in global memory for that class.
#include "sb_internal.hpp"
#include "sb_fftcore.hpp"

#include <cstdlib>
#include <stdexcept>

namespace sb {

template <int LOG2L>
struct R4Cfg {
  static constexpr int L = 1 << LOG2L;
  static constexpr int T = L / 16;                       // threads per team (one complex sequence)
  static constexpr int NFULL = (LOG2L - 1) / 4;          // strided radix-16 passes (2 for every supported length)
  static constexpr int RF = 1 << (LOG2L - 4 * NFULL);    // register-local final radix
  static constexpr int LP = L + L / 16;                  // padded team buffer (complex)
  static constexpr int NT = 512;
  static constexpr int NTEAMS = NT / T;
  static constexpr int MS1 = L >> 8;                     // stride of the second pass
  static constexpr int SETS = T > 128 ? T / 128 : 1;     // threads per TMEM lane
  // Tensor Memory columns of one table set (one complex double = 4 columns)
  static constexpr int C_TW0 = 0;                        // [16] pass-0 twiddles W_L^{k tl} (entry 0 unused)
  static constexpr int C_TW1 = 64;                       // [16] pass-1 twiddles
  static constexpr int C_FH = 128;                       // [16] FH[e T + tl]
  static constexpr int C_CH = 192;                       // [8]  chirp[n1 T + tl]
  static constexpr int C_PQ = 224;                       // [2 halves][8][P, Q]
  static constexpr bool PQ_TMEM = SETS == 1;
  static constexpr int SETCOLS = PQ_TMEM ? 352 : 224;
  static constexpr int NFILL = SETS == 1 ? 128 : T;      // threads that write the tables (whole warps, every lane / set)
  static constexpr int STG = L + 16;                     // staging doubles per team: a spectrum row (<= L - 1 doubles) + slack
  static constexpr size_t SMEM = sizeof(double2) * (size_t)NTEAMS * LP + sizeof(double) * (size_t)NTEAMS * STG + 16 * NTEAMS;
  static_assert(NFULL == 2, "v4 covers the lengths with two strided passes");
  static_assert(MS1 <= 16 && 32 % MS1 == 0, "second-pass blocks must stay inside a warp");
  static_assert(SETS * SETCOLS <= 512, "Tensor Memory has 512 columns");
};

__device__ __forceinline__ double sb_u2d(uint32_t lo, uint32_t hi) {
#ifdef SB_EMU
  const unsigned long long b = ((unsigned long long)hi << 32) | lo;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
#else
  return __hiloint2double((int)hi, (int)lo);
#endif
}
__device__ __forceinline__ void sb_d2u(double d, uint32_t& lo, uint32_t& hi) {
#ifdef SB_EMU
  unsigned long long b;
  std::memcpy(&b, &d, 8);
  lo = (uint32_t)b; hi = (uint32_t)(b >> 32);
#else
  lo = (uint32_t)__double2loint(d); hi = (uint32_t)__double2hiint(d);
#endif
}
__device__ __forceinline__ double2 tm_c(const uint32_t (&r)[16], int q) {
  return make_double2(sb_u2d(r[4 * q], r[4 * q + 1]), sb_u2d(r[4 * q + 2], r[4 * q + 3]));
}
__device__ __forceinline__ void tm_put(uint32_t addr, double2 z) {     // one complex value -> 4 columns of this thread's lane
  uint32_t r[4];
  sb_d2u(z.x, r[0], r[1]);
  sb_d2u(z.y, r[2], r[3]);
  sb_tmem_st4(addr, r);
}

template <int T>
__device__ __forceinline__ void team_sync4(int team) {
  if (T >= 64) sb_bar_sync(1 + team, T);
  else __syncwarp();
}

// v[k] *= W[k] (CONJ: conj(W[k])), k = 1..15, W = 16 complex values at TMEM columns taddr..taddr+63.  Four batches of
// four; the next batch is requested before the current one is used.
template <bool CONJ>
__device__ __forceinline__ void tm_twiddle(double2 (&v)[16], uint32_t taddr) {
  uint32_t r0[16], r1[16];
  sb_tmem_ld16(taddr, r0);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    uint32_t(&cur)[16] = (b & 1) ? r1 : r0;
    uint32_t(&nxt)[16] = (b & 1) ? r0 : r1;
    sb_tmem_wait_ld16(cur);
    if (b < 3) sb_tmem_ld16(taddr + 16 * (b + 1), nxt);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 4 * b + q;
      if (k) v[k] = CONJ ? cmc(v[k], tm_c(cur, q)) : cm(v[k], tm_c(cur, q));
    }
  }
}

// circular convolution with the pre-transformed chirp; tables from Tensor Memory (tb = this thread's table set).
// In: v[n1] = element n1*T + tl, n1 < 8 (upper half zero).  Out: v[n1] = element n1*T + tl of the result, n1 < 8.
// `after_first_exchange` runs once, after the first team barrier: by then every thread of the team has finished the
// prologue that precedes the call (the staging row may be overwritten).
template <int LOG2L, class Hook, bool PRUNE = true>
__device__ __forceinline__ void conv4(double2 (&v)[16], double2* buf, uint32_t tb, int tl, int team, bool active, Hook&& after_first_exchange,
                                      int fh_col = R4Cfg<LOG2L>::C_FH) {
  typedef R4Cfg<LOG2L> C;
  constexpr int T = C::T, M1 = C::MS1, LB1 = M1 << 4;
  const int b1 = tl / M1, j1 = tl - b1 * M1, base1 = padi(b1 * LB1 + j1), base0 = padi(tl);
  // ---- forward pass 0 (stride T): pruned radix-16, twiddle, store
  if (active) {
    if (PRUNE) fft16_fwd_lo8(v); else fft16<false>(v);
    tm_twiddle<false>(v, tb + C::C_TW0);
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[base0 + k * T + ((k * T) >> 4)] = v[k];
  }
  team_sync4<T>(team);
  after_first_exchange();
  __syncwarp();                 // the hook is one thread's work: the warp is whole again before the warp-collective TMEM loads
  // ---- forward pass 1 (stride M1)
  if (active) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {          // butterfly order: the first radix-4 can start after four loads
      const int n = (i >> 2) + 4 * (i & 3);
      v[n] = buf[base1 + n * M1 + ((n * M1) >> 4)];
    }
    fft16<false>(v);
    tm_twiddle<false>(v, tb + C::C_TW1);
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[base1 + k * M1 + ((k * M1) >> 4)] = v[k];
  }
  // the second pass works on blocks of 16 M1 elements owned by M1 <= 16 consecutive threads, and so do the final pass
  // and the first inverse pass: these two exchanges never leave a warp
  __syncwarp();
  // ---- final forward pass, pointwise product with FH, first inverse pass: all in registers
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = buf[tl * 17 + e];
    fft_final<false>(v, C::RF);
    {
      uint32_t r0[16], r1[16];
      sb_tmem_ld16(tb + fh_col, r0);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint32_t(&cur)[16] = (b & 1) ? r1 : r0;
        uint32_t(&nxt)[16] = (b & 1) ? r0 : r1;
        sb_tmem_wait_ld16(cur);
        if (b < 3) sb_tmem_ld16(tb + fh_col + 16 * (b + 1), nxt);
#pragma unroll
        for (int q = 0; q < 4; ++q) v[4 * b + q] = cm(v[4 * b + q], tm_c(cur, q));
      }
    }
    fft_final<true>(v, C::RF);
#pragma unroll
    for (int e = 0; e < 16; ++e) buf[tl * 17 + e] = v[e];
  }
  __syncwarp();
  // ---- inverse pass 1
  if (active) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = (i >> 2) + 4 * (i & 3);
      v[k] = buf[base1 + k * M1 + ((k * M1) >> 4)];
    }
    tm_twiddle<true>(v, tb + C::C_TW1);
    fft16<true>(v);
#pragma unroll
    for (int n = 0; n < 16; ++n) buf[base1 + n * M1 + ((n * M1) >> 4)] = v[n];
  }
  team_sync4<T>(team);
  // ---- inverse pass 0: result stays in registers
  if (active) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = (i >> 2) + 4 * (i & 3);
      v[k] = buf[base0 + k * T + ((k * T) >> 4)];
    }
    tm_twiddle<true>(v, tb + C::C_TW0);
    fft16<true>(v);
  }
}

// =====================================================================================
// inverse: spectra (value, d/dr, d2/dr2) -> 5 real rows; rows of a ring are rho = zb*5 + f
// =====================================================================================
// BLK: rows go to the blocked SZ layout (sb_internal.hpp RowDst, out_is_phys == 2) -- a separate instantiation, so that the
// ring-row / physical destinations keep their compile-time store stride
template <int LOG2L, bool BLK>
__global__ void __launch_bounds__(512, 1) k_inv_l4(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                   const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                   const double* __restrict__ blob, const double* __restrict__ in,
                                                   long long in_fs, long long in_vs, double* __restrict__ out,
                                                   long long out_fs, long long out_vs, int out_is_phys, int var0,
                                                   unsigned lmask, int* counter, int chunk) {
  typedef R4Cfg<LOG2L> C;
  constexpr int L = C::L, T = C::T;
  SB_DYN_SMEM(double2, sm);
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  double2* const buf = sm + (size_t)team * C::LP;
  double* const stg_all = reinterpret_cast<double*>(sm + (size_t)C::NTEAMS * C::LP);
  double* const stg = stg_all + (size_t)team * C::STG;                                   // 16-byte aligned (STG even)
  sb_mbar_t* const mbar = reinterpret_cast<sb_mbar_t*>(reinterpret_cast<char*>(stg_all + (size_t)C::NTEAMS * C::STG) + 16 * team);
  if (tid < 32) sb_tmem_alloc(&s_tmem, 512);
  if (tl == 0) sb_mbar_init(mbar, 1);
  sb_fence_mbar_init();
  sb_tmem_fence_before_sync();
  __syncthreads();
  sb_tmem_fence_after_sync();
  const int set = (C::SETS > 1) ? tl / 128 : 0;
  const uint32_t tb = sb_tmem_warp_base(s_tmem) + (uint32_t)(set * C::SETCOLS);
  // ---- class twiddles -> Tensor Memory (once per CTA): thread tid < NFILL writes lane tid % 128, set tid / 128
  if (tid < C::NFILL) {
    const int tf = tid % T;                       // the tl whose tables this (lane, set) holds
    const uint32_t tf_b = sb_tmem_warp_base(s_tmem) + (uint32_t)((C::SETS > 1 ? tid / 128 : 0) * C::SETCOLS);
    const double2* tw1 = twp + 15 * T;
    tm_put(tf_b + C::C_TW0, make_double2(1.0, 0.0));
    tm_put(tf_b + C::C_TW1, make_double2(1.0, 0.0));
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      tm_put(tf_b + C::C_TW0 + 4 * k, twp[(k - 1) * T + tf]);
      tm_put(tf_b + C::C_TW1 + 4 * k, tw1[(k - 1) * C::MS1 + (tf % C::MS1)]);
    }
    sb_tmem_wait_st();
  }
  unsigned phase = 0;                             // parity of the team's staging barrier
  int cur_ring = -1;
  // contiguous share of the work list (items are ordered by ring): a CTA changes ring -- and reloads the ring's tables --
  // once every few items instead of at every item
  const int total = nwork * nvars;
  const int per_cta = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  __shared__ int s_chunk;
  int w_begin = (int)blockIdx.x * per_cta, w_end = w_begin + per_cta < total ? w_begin + per_cta : total;
  for (;;) {
  if (counter) {                    // dynamic shares (launches that share the chip with other kernels): `chunk` consecutive items
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    w_begin = s_chunk * chunk;
    w_end = w_begin + chunk < total ? w_begin + chunk : total;
    __syncthreads();
    if (w_begin >= total) break;
  }
  for (int w = w_begin; w < w_end; ++w) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* PQ = reinterpret_cast<const double2*>(blob + pl.off2);
    if (wk.r != cur_ring) {          // ring tables -> Tensor Memory
      sb_tmem_fence_before_sync();
      __syncthreads();               // nobody is still reading the previous ring's tables
      sb_tmem_fence_after_sync();
      if (tid < C::NFILL) {
        const int tf = tid % T;
        const uint32_t tf_b = sb_tmem_warp_base(s_tmem) + (uint32_t)((C::SETS > 1 ? tid / 128 : 0) * C::SETCOLS);
#pragma unroll
        for (int e0 = 0; e0 < 16; e0 += 8) {
          double2 x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = FH_g[(e0 + e) * T + tf];
#pragma unroll
          for (int e = 0; e < 8; ++e) tm_put(tf_b + C::C_FH + 4 * (e0 + e), x[e]);
        }
        {
          double2 x[8];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) { const int a0 = n1 * T + tf; x[n1] = chirp_g[a0 < m ? a0 : m - 1]; }
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) tm_put(tf_b + C::C_CH + 4 * n1, x[n1]);
        }
        if (C::PQ_TMEM) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double2 p[8], q[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
              const int k0 = n1 * T + tf, k = k0 < m ? k0 : m - 1;
              p[n1] = PQ[(size_t)(2 * h) * m + k];
              q[n1] = PQ[(size_t)(2 * h + 1) * m + k];
            }
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
              tm_put(tf_b + C::C_PQ + 4 * ((h * 8 + n1) * 2), p[n1]);
              tm_put(tf_b + C::C_PQ + 4 * ((h * 8 + n1) * 2 + 1), q[n1]);
            }
          }
        }
        sb_tmem_wait_st();
      }
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      cur_ring = wk.r;
    }
    const long long woff = g.ring_woff[wk.r], hoff = BLK ? g.ring_hoffp[wk.r] : g.ring_hoff[wk.r];
    const int nseq = 2 * wk.nrows;
    // spectrum row of sequence s (rows f = 0, 3, 4 read the value spectrum)
    auto row_of = [&](int s) -> const double* {
      const int rho = wk.row0 + (s >> 1), zb = rho / 5, f = rho - zb * 5;
      return in + (long long)(f < 3 ? f : 0) * in_fs + (long long)v_ * in_vs + (long long)zb * g.W + woff;
    };
    auto wanted = [&](int s) { const int rho = wk.row0 + (s >> 1); return ((lmask >> (rho % 5)) & 1u) != 0u; };
    // one thread asks the copy engine for a row: from the 16-byte boundary at or below its first double, a multiple of
    // 16 bytes (the row has 2m-1 doubles and starts at an arbitrary double; at most one foreign double on either side is
    // copied along -- both lie inside the SL scratch array)
    auto request = [&](int s) {
      const double* sp = row_of(s);
      const int sh = (int)(((uintptr_t)sp >> 3) & 1);
      const unsigned bytes = (unsigned)(((sh + 2 * m - 1) * 8 + 15) & ~15);
      sb_fence_proxy_async();
      sb_mbar_expect_tx(mbar, bytes);
      sb_bulk_g2s(stg, sp - sh, bytes, mbar);
    };
    // first wanted sequence of this team in the item
    int s = team;
    while (s < nseq && !wanted(s)) s += C::NTEAMS;
    team_sync4<T>(team);                 // the team's previous sequence (previous item) has left the staging row and buf
    if (s < nseq && tl == 0) request(s);
    // every team runs the same number of loop trips, so that whole warps stay together at the warp-collective TMEM loads
    for (int trip = team; trip - team < nseq; trip += C::NTEAMS) {
      const bool active = s < nseq;
      int s_next = nseq;
      if (active) {
        s_next = s + C::NTEAMS;
        while (s_next < nseq && !wanted(s_next)) s_next += C::NTEAMS;
      }
      const int row = s >> 1, half = s & 1;
      const int rho = wk.row0 + row;
      const int zb = rho / 5, f = rho - zb * 5;
      double2 v[16];
      if (active) {
        sb_mbar_wait(mbar, phase);       // the row has landed in the staging buffer
        phase ^= 1u;
        const double* st = stg + (int)(((uintptr_t)row_of(s) >> 3) & 1);
        const double2* Ph = PQ + (size_t)(2 * half) * m;
        const double2* Qh = Ph + m;
#pragma unroll
        for (int h4 = 0; h4 < 4; ++h4) {           // two spectrum elements per batch: P, Q of both in one TMEM load
          uint32_t r[16];
          double2 Pk[2], Qk[2];
          if (C::PQ_TMEM) {
            sb_tmem_ld16(tb + C::C_PQ + 4 * ((half * 8 + 2 * h4) * 2), r);
          }
          double cx[2], cy[2], qx[2], qy[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k0 = (h4 * 2 + j) * T + tl;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            cx[j] = st[k ? 2 * k - 1 : 0]; cy[j] = st[2 * k];
            qx[j] = st[km ? 2 * km - 1 : 0]; qy[j] = st[2 * km];
            if (!C::PQ_TMEM) { Pk[j] = Ph[k]; Qk[j] = Qh[k]; }
          }
          if (C::PQ_TMEM) {
            sb_tmem_wait_ld16(r);
            Pk[0] = tm_c(r, 0); Qk[0] = tm_c(r, 1); Pk[1] = tm_c(r, 2); Qk[1] = tm_c(r, 3);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k0 = (h4 * 2 + j) * T + tl;
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
            v[h4 * 2 + j] = k0 < m ? u : make_double2(0.0, 0.0);
          }
        }
      }
      conv4<LOG2L>(v, buf, tb, tl, team, active, [&]() {
        if (tl == 0 && s_next < nseq) request(s_next);     // lands while this convolution runs
      });
      if (active) {
        double* orow;                 // destination of this thread's first output a = tl; the others follow at a constant stride
        long long ostep = 4 * T;
        if (BLK) {
          const int k = zb >> 1, j0 = 4 * tl + 2 * half;
          orow = out + (long long)f * out_fs + (long long)v_ * out_vs + (long long)g.bz * hoff + 16 * ((zb & 1) * ((g.bz + 1) >> 1) + k) +
                 (long long)(j0 >> 4) * (16 * g.bz) + ((j0 & 15) ^ ((k & 3) << 2));
          ostep = (long long)(4 * T) * g.bz;
        } else if (out_is_phys) {
          orow = out + ((long long)f * g.V + var0 + v_) * g.N + hoff + 2 * half + 4 * tl;
        } else {
          orow = out + (long long)f * out_fs + (long long)v_ * out_vs + (long long)g.bz * hoff + (long long)zb * n + 2 * half + 4 * tl;
        }
        uint32_t r0[16], r1[16];
        sb_tmem_ld16(tb + C::C_CH, r0);
        sb_tmem_ld16(tb + C::C_CH + 16, r1);
        sb_tmem_wait_ld16(r0);
        sb_tmem_wait_ld16(r1);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a = n1 * T + tl;
          const double2 ch = n1 < 4 ? tm_c(r0, n1 & 3) : tm_c(r1, n1 & 3);
          if (a < m) {
            const double2 Y = cm(v[n1], ch);
            *reinterpret_cast<double2*>(orow + (BLK ? n1 * ostep : (long long)(4 * T * n1))) = make_double2(Y.x, -Y.y);
          }
        }
      }
      team_sync4<T>(team);            // the team has finished reading buf (inverse pass 0) before the next forward pass 0 writes it
      s = s_next;
    }
  }
  if (!counter) break;
  }
  sb_tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) sb_tmem_dealloc(s_tmem, 512);
}


// =====================================================================================
// forward: real ring rows -> retained coefficients k = 0..ri.  Teams work in pairs (the two packed sequences of one
// row).  The row (n = 4m doubles, contiguous and 32-byte aligned) is staged once per pair by the bulk-copy engine; the
// pair's "full" barrier tells both teams it has landed, its "consumed" barrier (one arrival per team, after the team's
// first exchange barrier) tells the requesting thread that the staging buffer may be refilled with the next row.  The
// raw convolution outputs are parked in the teams' own buffers and the pair combines them with the pre-combined tables
// A0..A3, which live in Tensor Memory like every other per-thread table (both halves' entries in every lane).
// =====================================================================================
template <int LOG2L>
struct R4FCfg : R4Cfg<LOG2L> {
  typedef R4Cfg<LOG2L> B;
  static constexpr int NPAIRS = B::NTEAMS / 2;
  static constexpr int STGF = 2 * B::L;                  // staging doubles per pair: one real row (n = 4m <= 2L)
  static constexpr size_t SMEMF = sizeof(double2) * (size_t)B::NTEAMS * B::LP + sizeof(double) * (size_t)NPAIRS * STGF + 16 * B::NTEAMS;
};

template <int T>
__device__ __forceinline__ void pair_sync4(int pair) {
  if (2 * T >= 64) sb_bar_sync((T >= 64 ? 9 : 1) + pair, 2 * T);
  else __syncwarp();
}

template <int LOG2L>
__global__ void __launch_bounds__(512, 1) k_fwd_l4(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                   const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                   const double* __restrict__ blob, const double* __restrict__ in,
                                                   long long in_vs, double* __restrict__ mirror, long long mirror_vs,
                                                   double* __restrict__ out, long long out_vs, int* counter, int chunk) {
  typedef R4FCfg<LOG2L> C;
  constexpr int T = C::T, NPAIRS = C::NPAIRS;
  SB_DYN_SMEM(double2, sm);
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int pair = team >> 1, half = team & 1;
  double2* const buf = sm + (size_t)team * C::LP;
  const double2* const buf0 = sm + (size_t)(2 * pair) * C::LP;
  const double2* const buf1 = buf0 + C::LP;
  double* const stg_all = reinterpret_cast<double*>(sm + (size_t)C::NTEAMS * C::LP);
  const double* const stg = stg_all + (size_t)pair * C::STGF;                                 // 16-byte aligned
  char* const mb = reinterpret_cast<char*>(stg_all + (size_t)NPAIRS * C::STGF) + 32 * pair;
  sb_mbar_t* const full = reinterpret_cast<sb_mbar_t*>(mb);
  sb_mbar_t* const consumed = reinterpret_cast<sb_mbar_t*>(mb + 16);
  if (tid < 32) sb_tmem_alloc(&s_tmem, 512);
  if (half == 0 && tl == 0) { sb_mbar_init(full, 1); sb_mbar_init(consumed, 2); }
  sb_fence_mbar_init();
  sb_tmem_fence_before_sync();
  __syncthreads();
  sb_tmem_fence_after_sync();
  const int set = (C::SETS > 1) ? tl / 128 : 0;
  const uint32_t tb = sb_tmem_warp_base(s_tmem) + (uint32_t)(set * C::SETCOLS);
  if (tid < C::NFILL) {          // class twiddles -> Tensor Memory (as k_inv_l4)
    const int tf = tid % T;
    const uint32_t tf_b = sb_tmem_warp_base(s_tmem) + (uint32_t)((C::SETS > 1 ? tid / 128 : 0) * C::SETCOLS);
    const double2* tw1 = twp + 15 * T;
    tm_put(tf_b + C::C_TW0, make_double2(1.0, 0.0));
    tm_put(tf_b + C::C_TW1, make_double2(1.0, 0.0));
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      tm_put(tf_b + C::C_TW0 + 4 * k, twp[(k - 1) * T + tf]);
      tm_put(tf_b + C::C_TW1 + 4 * k, tw1[(k - 1) * C::MS1 + (tf % C::MS1)]);
    }
    sb_tmem_wait_st();
  }
  unsigned ph_full = 0, ph_cons = 0;
  int cur_ring = -1;
  const int total = nwork * nvars;
  const int per_cta = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  __shared__ int s_chunk;
  int w_begin = (int)blockIdx.x * per_cta, w_end = w_begin + per_cta < total ? w_begin + per_cta : total;
  for (;;) {
  if (counter) {                    // dynamic shares (launches that share the chip with other kernels): `chunk` consecutive items
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    w_begin = s_chunk * chunk;
    w_end = w_begin + chunk < total ? w_begin + chunk : total;
    __syncthreads();
    if (w_begin >= total) break;
  }
  for (int w = w_begin; w < w_end; ++w) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* AF = reinterpret_cast<const double2*>(blob + pl.off2) + (size_t)4 * m;
    if (wk.r != cur_ring) {          // ring tables -> Tensor Memory
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      if (tid < C::NFILL) {
        const int tf = tid % T;
        const uint32_t tf_b = sb_tmem_warp_base(s_tmem) + (uint32_t)((C::SETS > 1 ? tid / 128 : 0) * C::SETCOLS);
#pragma unroll
        for (int e0 = 0; e0 < 16; e0 += 8) {
          double2 x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = FH_g[(e0 + e) * T + tf];
#pragma unroll
          for (int e = 0; e < 8; ++e) tm_put(tf_b + C::C_FH + 4 * (e0 + e), x[e]);
        }
        {
          double2 x[8];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) { const int a0 = n1 * T + tf; x[n1] = chirp_g[a0 < m ? a0 : m - 1]; }
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) tm_put(tf_b + C::C_CH + 4 * n1, x[n1]);
        }
        if (C::PQ_TMEM) {            // A0..A3 at k = h T + tf + 2 T i, h = 0, 1, i = 0..3
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k0 = h * T + tf + 2 * T * i, k = k0 < m ? k0 : m - 1;
              double2 a[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) a[q] = AF[(size_t)q * m + k];
#pragma unroll
              for (int q = 0; q < 4; ++q) tm_put(tf_b + C::C_PQ + 4 * ((h * 4 + i) * 4 + q), a[q]);
            }
          }
        }
        sb_tmem_wait_st();
      }
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      cur_ring = wk.r;
    }
    const long long hoff = g.ring_hoff[wk.r];
    const double* src = in + (long long)v_ * in_vs + (long long)g.bz * hoff;
    double* mir = mirror ? mirror + (long long)v_ * mirror_vs + (long long)g.bz * hoff : nullptr;
    double* dst = out + (long long)v_ * out_vs + g.ring_woff[wk.r];
    auto request = [&](int row) {          // the whole real row: n doubles, 32-byte aligned
      sb_fence_proxy_async();
      sb_mbar_expect_tx(full, (unsigned)n * 8u);
      sb_bulk_g2s(const_cast<double*>(stg), src + (long long)(wk.row0 + row) * n, (unsigned)n * 8u, full);
    };
    pair_sync4<T>(pair);             // the pair's previous row (previous item) has left the staging row and the buffers
    if (half == 1 && tl == 0 && pair < wk.nrows) request(pair);
    for (int row = pair; row - pair < wk.nrows; row += NPAIRS) {
      const bool active = row < wk.nrows;
      double2 v[16];
      if (active) {
        sb_mbar_wait(full, ph_full);
        ph_full ^= 1u;
        double* mp = mir ? mir + (long long)(wk.row0 + row) * n + 2 * half : nullptr;
        const double* rp = stg + 2 * half;
        uint32_t r0[16], r1[16];
        sb_tmem_ld16(tb + C::C_CH, r0);
        sb_tmem_ld16(tb + C::C_CH + 16, r1);
        double2 x[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a0 = n1 * T + tl, a = a0 < m ? a0 : m - 1;
          x[n1] = *reinterpret_cast<const double2*>(rp + 4 * a);
        }
        sb_tmem_wait_ld16(r0);
        sb_tmem_wait_ld16(r1);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a0 = n1 * T + tl, a = a0 < m ? a0 : m - 1;
          if (mp && a0 < m) *reinterpret_cast<double2*>(mp + 4 * a) = x[n1];
          const double2 y = cm(x[n1], n1 < 4 ? tm_c(r0, n1 & 3) : tm_c(r1, n1 & 3));
          v[n1] = a0 < m ? y : make_double2(0.0, 0.0);
        }
      }
      conv4<LOG2L>(v, buf, tb, tl, team, active, [&]() {
        // this team has finished reading the staged row; when both teams have, the next row may overwrite it
        if (active && tl == 0) {
          sb_mbar_arrive(consumed);
          if (half == 1) {
            sb_mbar_wait(consumed, ph_cons);
            ph_cons ^= 1u;
            if (row + NPAIRS < wk.nrows) request(row + NPAIRS);
          }
        }
      });
      team_sync4<T>(team);            // every thread of the team has read its last-pass inputs
      if (active) {
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int a = n1 * T + tl;
          if (a < m) buf[a] = v[n1];
        }
      }
      pair_sync4<T>(pair);
      if (active) {
        double* o = dst + (long long)(wk.row0 + row) * g.W;
        double2 Ag[C::PQ_TMEM ? 1 : 4][4];          // L = 4096: A0..A3 come from global memory, all sixteen loads up front
        if (!C::PQ_TMEM) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k0 = half * T + tl + 2 * T * i, k = k0 < m ? k0 : m - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) Ag[C::PQ_TMEM ? 0 : i][q] = AF[(size_t)q * m + k];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = half * T + tl + 2 * T * i;
          uint32_t r[16];
          double2 A0, A1, A2, A3;
          if (C::PQ_TMEM) {
            sb_tmem_ld16(tb + C::C_PQ + 4 * ((half * 4 + i) * 4), r);
            sb_tmem_wait_ld16(r);
            A0 = tm_c(r, 0); A1 = tm_c(r, 1); A2 = tm_c(r, 2); A3 = tm_c(r, 3);
          }
          if (k < m) {
            if (!C::PQ_TMEM) { A0 = Ag[C::PQ_TMEM ? 0 : i][0]; A1 = Ag[C::PQ_TMEM ? 0 : i][1]; A2 = Ag[C::PQ_TMEM ? 0 : i][2]; A3 = Ag[C::PQ_TMEM ? 0 : i][3]; }
            const int km = k ? m - k : 0;
            const double2 b0 = buf0[k], b0m = buf0[km], b1 = buf1[k], b1m = buf1[km];
            double2 X = cm(b0, A0) + cm(make_double2(b0m.x, -b0m.y), A1);
            X = X + cm(b1, A2) + cm(make_double2(b1m.x, -b1m.y), A3);
            if (k == 0) {
              o[0] = X.x;
            } else {
              o[2 * k - 1] = X.x;
              o[2 * k] = X.y;
            }
          }
        }
      }
      pair_sync4<T>(pair);            // the pair has finished combining out of buf0 / buf1 before the next row's first pass writes them
    }
  }
  if (!counter) break;
  }
  sb_tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) sb_tmem_dealloc(s_tmem, 512);
}


// =====================================================================================
// composite convolution lengths L = 3 * L2 (L2 = 1024, 2048) with the v4 data movement: rings with L2 < m <= 1.5 L2 of the
// outer tiles of a multi-GPU patch (m up to 3072), which a power of two would pad to 4 * L2.  Mathematics as k_inv_l3 /
// k_fwd_l3 (sb_ringfft2.cu): radix-3 decimation in frequency of the zero-padded input over a group of three teams, each
// convolving one residue class with its third of the pre-transformed chirp, recombination of the three results.  Per-thread
// tables (pass twiddles, the three FH thirds, the chirp at this thread's outputs, W_L^{tl}, W_L^{2 tl}) live in Tensor
// Memory; the spectrum row (inverse) / real row (forward) of the group's next sequence is staged by the bulk-copy engine.
// =====================================================================================
template <int LOG2L2>
struct R5Cfg {
  typedef R4Cfg<LOG2L2> B;
  static constexpr int L2 = B::L, T = B::T;
  static constexpr int NGROUPS = (512 / T) / 3;
  static constexpr int NTEAMS = 3 * NGROUPS;
  static constexpr int NT = NTEAMS * T;
  static constexpr int C_FH3 = 128;                      // [3][16] FH thirds
  static constexpr int C_CH3 = 320;                      // [3][8] chirp at the outputs of team r
  static constexpr int C_WB = 416;                       // [2] W_L^{tl}, W_L^{2 tl}  (L = 3 L2)
  static constexpr int STGI = 3 * L2 + 16;               // inverse staging doubles per group: spectrum row (2m-1 <= 3 L2 - 1)
  static constexpr int STGF = 6 * L2;                    // forward staging doubles per group: real row (n = 4m <= 6 L2)
  static constexpr int CH = 3 * L2 / 2;                  // parked outputs per half (forward)
  static constexpr size_t SMEMI = sizeof(double2) * ((size_t)NTEAMS * B::LP + 32) + sizeof(double) * (size_t)NGROUPS * STGI + 16 * NGROUPS;
  static constexpr size_t SMEMF = sizeof(double2) * ((size_t)NTEAMS * B::LP + 32) + sizeof(double) * (size_t)NGROUPS * STGF + 16 * NGROUPS;
  static_assert(T <= 128 && 128 % T == 0, "one table set per TMEM lane");
};

template <int T>
__device__ __forceinline__ void group_sync5(int grp) {
  sb_bar_sync((T >= 64 ? 9 : 1) + grp, 3 * T);
}

#define W3R5 (-0.5)
#define W3I5 (-0.86602540378443864676)   // w3 = exp(-2 pi i / 3)

// class + ring tables of a composite class -> Tensor Memory (threads 0..127, lane tid, tl = tid % T)
template <int LOG2L2>
__device__ __forceinline__ void fill_class5(uint32_t s_tmem, const double2* __restrict__ twp, int tid) {
  typedef R5Cfg<LOG2L2> C;
  constexpr int T = C::T;
  const int tf = tid % T;
  const uint32_t tf_b = sb_tmem_warp_base(s_tmem);
  const double2* tw1 = twp + 15 * T;
  const double2* xt = tw1 + 15 * C::B::MS1;               // TWB1[T] TWB2[T] C1[16] C2[16]
  tm_put(tf_b + C::B::C_TW0, make_double2(1.0, 0.0));
  tm_put(tf_b + C::B::C_TW1, make_double2(1.0, 0.0));
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    tm_put(tf_b + C::B::C_TW0 + 4 * k, twp[(k - 1) * T + tf]);
    tm_put(tf_b + C::B::C_TW1 + 4 * k, tw1[(k - 1) * C::B::MS1 + (tf % C::B::MS1)]);
  }
  tm_put(tf_b + C::C_WB, xt[tf]);
  tm_put(tf_b + C::C_WB + 4, xt[T + tf]);
  sb_tmem_wait_st();
}
template <int LOG2L2>
__device__ __forceinline__ void fill_ring5(uint32_t s_tmem, const double2* __restrict__ FH_g, const double2* __restrict__ chirp_g, int m, int tid) {
  typedef R5Cfg<LOG2L2> C;
  constexpr int T = C::T, L2 = C::L2;
  const int tf = tid % T;
  const uint32_t tf_b = sb_tmem_warp_base(s_tmem);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int e0 = 0; e0 < 16; e0 += 8) {
      double2 x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = FH_g[(size_t)r * L2 + (e0 + e) * T + tf];
#pragma unroll
      for (int e = 0; e < 8; ++e) tm_put(tf_b + C::C_FH3 + 64 * r + 4 * (e0 + e), x[e]);
    }
    double2 c[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      const int a0 = (r == 2 ? L2 : 0) + (r == 1 ? L2 / 2 : 0) + n1 * T + tf;
      c[n1] = chirp_g[a0 < m ? a0 : m - 1];
    }
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) tm_put(tf_b + C::C_CH3 + 32 * r + 4 * n1, c[n1]);
  }
  sb_tmem_wait_st();
}

// W_L^{j r} for j = n1*T + tl (r = 1, 2): W_L^{r tl} from Tensor Memory times W_L^{r T n1} from the 32-entry shared table
__device__ __forceinline__ double2 twid5(const double2 wb, const double2* __restrict__ ctab, int r, int n1) {
  return cm(wb, ctab[(r - 1) * 16 + n1]);
}

// x thirds are in the group's buffers -> s_r -> convolution -> c_r' back in the buffers
struct GB3 {                     // the three exchange buffers of a group and the calling team's own one
  double2 *b0, *b1, *b2, *mine;
};
template <int LOG2L2>
__device__ __forceinline__ void conv5(const GB3& gb, uint32_t tb, const double2* __restrict__ ctab, int r, int grp, int team, int tl) {
  typedef R5Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  double2 v[16];
  double2 wb = make_double2(1.0, 0.0);
  if (r) {
    uint32_t q[4];
    sb_tmem_ld4(tb + C::C_WB + 4 * (r - 1), q);
    sb_tmem_wait_ld4(q);
    wb = make_double2(sb_u2d(q[0], q[1]), sb_u2d(q[2], q[3]));
  }
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int j = n1 * T + tl;
    double2 x0 = (n1 < 8) ? gb.b0[j] : gb.b1[j - L2 / 2];
    if (n1 < 8) {
      const double2 x1 = gb.b2[j];
      if (r == 0) x0 = x0 + x1;
      else if (r == 1) x0 = x0 + make_double2(W3R5 * x1.x - W3I5 * x1.y, W3R5 * x1.y + W3I5 * x1.x);
      else x0 = x0 + make_double2(W3R5 * x1.x + W3I5 * x1.y, W3R5 * x1.y - W3I5 * x1.x);       // w3^2 = conj(w3)
    }
    v[n1] = r ? cm(x0, twid5(wb, ctab, r, n1)) : x0;
  }
  group_sync5<T>(grp);              // every team has gathered its s_r: the buffers may be overwritten
  conv4<LOG2L2, void (*)(), false>(v, gb.mine, tb, tl, team, true, +[]() {}, C::C_FH3 + 64 * r);
  team_sync4<T>(team);             // the team's last-pass loads are done
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const double2 c = r ? cmc(v[n1], twid5(wb, ctab, r, n1)) : v[n1];
    gb.mine[n1 * T + tl] = c;
  }
  group_sync5<T>(grp);              // c_0, c_1', c_2' are in the three buffers
}

// this team's eight outputs: r = 0: a in [0, L2/2), r = 1: [L2/2, L2), r = 2: [L2, 3 L2/2)
__device__ __forceinline__ double2 combine5(const GB3& gb, int r, int i) {
  const double2 b0 = gb.b0[i], b1 = gb.b1[i], b2 = gb.b2[i];
  if (r < 2) return b0 + b1 + b2;
  return b0 + make_double2(W3R5 * b1.x + W3I5 * b1.y, W3R5 * b1.y - W3I5 * b1.x) +
         make_double2(W3R5 * b2.x - W3I5 * b2.y, W3R5 * b2.y + W3I5 * b2.x);
}

template <int LOG2L2>
__global__ void __launch_bounds__(R5Cfg<LOG2L2>::NT, 1) k_inv_l5(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                                const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                                const double* __restrict__ blob, const double* __restrict__ in,
                                                                long long in_fs, long long in_vs, double* __restrict__ out,
                                                                long long out_fs, long long out_vs, int out_is_phys, int var0,
                                                                unsigned lmask, int* counter, int chunk) {
  typedef R5Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  SB_DYN_SMEM(double2, sm);
  __shared__ uint32_t s_tmem;
  __shared__ int s_chunk;
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int grp = team / 3, r = team - 3 * grp;
  GB3 gb;
  gb.b0 = sm + (size_t)(3 * grp) * C::B::LP; gb.b1 = gb.b0 + C::B::LP; gb.b2 = gb.b1 + C::B::LP;
  gb.mine = gb.b0 + (size_t)r * C::B::LP;
  double2* const ctab = sm + (size_t)C::NTEAMS * C::B::LP;                 // C1[16] C2[16]
  double* const stg_all = reinterpret_cast<double*>(ctab + 32);
  double* const stg = stg_all + (size_t)grp * C::STGI;
  sb_mbar_t* const mbar = reinterpret_cast<sb_mbar_t*>(reinterpret_cast<char*>(stg_all + (size_t)C::NGROUPS * C::STGI) + 16 * grp);
  if (tid < 32) sb_tmem_alloc(&s_tmem, 512);
  if (r == 0 && tl == 0) sb_mbar_init(mbar, 1);
  if (tid < 32) ctab[tid] = twp[15 * T + 15 * C::B::MS1 + 2 * T + tid];
  sb_fence_mbar_init();
  sb_tmem_fence_before_sync();
  __syncthreads();
  sb_tmem_fence_after_sync();
  const uint32_t tb = sb_tmem_warp_base(s_tmem);
  if (tid < 128) fill_class5<LOG2L2>(s_tmem, twp, tid);
  unsigned phase = 0;
  int cur_ring = -1;
  const int total = nwork * nvars;
  const int per_cta = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  int w_begin = (int)blockIdx.x * per_cta, w_end = w_begin + per_cta < total ? w_begin + per_cta : total;
  for (;;) {
  if (counter) {
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    w_begin = s_chunk * chunk;
    w_end = w_begin + chunk < total ? w_begin + chunk : total;
    __syncthreads();
    if (w_begin >= total) break;
  }
  for (int w = w_begin; w < w_end; ++w) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* PQ = reinterpret_cast<const double2*>(blob + pl.off2);
    if (wk.r != cur_ring) {
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      if (tid < 128) fill_ring5<LOG2L2>(s_tmem, FH_g, chirp_g, m, tid);
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      cur_ring = wk.r;
    }
    const long long woff = g.ring_woff[wk.r], hoff = (out_is_phys == 2 ? g.ring_hoffp : g.ring_hoff)[wk.r];
    const int nseq = 2 * wk.nrows;
    auto row_of = [&](int s) -> const double* {
      const int rho = wk.row0 + (s >> 1), zb = rho / 5, f = rho - zb * 5;
      return in + (long long)(f < 3 ? f : 0) * in_fs + (long long)v_ * in_vs + (long long)zb * g.W + woff;
    };
    auto wanted = [&](int s) { const int rho = wk.row0 + (s >> 1); return ((lmask >> (rho % 5)) & 1u) != 0u; };
    auto request = [&](int s) {
      const double* sp = row_of(s);
      const int sh = (int)(((uintptr_t)sp >> 3) & 1);
      const unsigned bytes = (unsigned)(((sh + 2 * m - 1) * 8 + 15) & ~15);
      sb_fence_proxy_async();
      sb_mbar_expect_tx(mbar, bytes);
      sb_bulk_g2s(stg, sp - sh, bytes, mbar);
    };
    int s = grp;
    while (s < nseq && !wanted(s)) s += C::NGROUPS;
    group_sync5<T>(grp);               // the group's previous sequence (previous item) has left the staging row and the buffers
    if (s < nseq && r == 0 && tl == 0) request(s);
    while (s < nseq) {
      int s_next = s + C::NGROUPS;
      while (s_next < nseq && !wanted(s_next)) s_next += C::NGROUPS;
      const int row = s >> 1, half = s & 1;
      const int rho = wk.row0 + row;
      const int zb = rho / 5, f = rho - zb * 5;
      sb_mbar_wait(mbar, phase);
      phase ^= 1u;
      {                              // prologue: this team's third of x (formula: k_inv_l4)
        const double* st = stg + (int)(((uintptr_t)row_of(s) >> 3) & 1);
        const double2* Ph = PQ + (size_t)(2 * half) * m;
        const double2* Qh = Ph + m;
#pragma unroll
        for (int h4 = 0; h4 < 4; ++h4) {
          double cx[2], cy[2], qx[2], qy[2];
          double2 Pk[2], Qk[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k0 = r * (L2 / 2) + (h4 * 2 + j) * T + tl;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            cx[j] = st[k ? 2 * k - 1 : 0]; cy[j] = st[2 * k];
            qx[j] = st[km ? 2 * km - 1 : 0]; qy[j] = st[2 * km];
            Pk[j] = Ph[k]; Qk[j] = Qh[k];
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int i = (h4 * 2 + j) * T + tl;
            const int k0 = r * (L2 / 2) + i;
            const int k = k0 < m ? k0 : m - 1;
            const int km = k ? m - k : 0;
            const double2 ck = make_double2(cx[j], k ? cy[j] : 0.0);
            const double2 cq = make_double2(qx[j], km ? qy[j] : 0.0);
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
            gb.mine[i] = k0 < m ? u : make_double2(0.0, 0.0);
          }
        }
      }
      group_sync5<T>(grp);            // x is complete and the staging row has been read by all three teams
      if (r == 0 && tl == 0 && s_next < nseq) request(s_next);
      __syncwarp();
      conv5<LOG2L2>(gb, tb, ctab, r, grp, team, tl);
      {
        const RowDst orow = row_dst(g, out, out_fs, out_vs, out_is_phys, f, v_, var0, hoff, n, zb);
        uint32_t r0[16], r1[16];
        sb_tmem_ld16(tb + C::C_CH3 + 32 * r, r0);
        sb_tmem_ld16(tb + C::C_CH3 + 32 * r + 16, r1);
        sb_tmem_wait_ld16(r0);
        sb_tmem_wait_ld16(r1);
#pragma unroll
        double* const o0 = orow.at(4 * ((r == 2 ? L2 : 0) + (r == 1 ? L2 / 2 : 0) + tl) + 2 * half);
        const long long ostep = orow.step(T);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int i = (r == 1 ? L2 / 2 : 0) + n1 * T + tl;
          const int a = (r == 2 ? L2 : 0) + i;
          const double2 ch = n1 < 4 ? tm_c(r0, n1 & 3) : tm_c(r1, n1 & 3);
          if (a < m) {
            const double2 Y = cm(combine5(gb, r, i), ch);
            *reinterpret_cast<double2*>(o0 + n1 * ostep) = make_double2(Y.x, -Y.y);
          }
        }
      }
      group_sync5<T>(grp);            // the group's outputs have been read out of the buffers
      s = s_next;
    }
  }
  if (!counter) break;
  }
  sb_tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) sb_tmem_dealloc(s_tmem, 512);
}

// forward: both packed sequences of a row go through the group one after the other (the real row is staged once); their
// raw convolution outputs are parked in a per-group global scratch line (L2-resident) and combined with A0..A3 at the end.
template <int LOG2L2>
__global__ void __launch_bounds__(R5Cfg<LOG2L2>::NT, 1) k_fwd_l5(DevGrid g, const LWork* __restrict__ work, int nwork, int nvars,
                                                                const double2* __restrict__ twp, const RingPlan* __restrict__ plans,
                                                                const double* __restrict__ blob, const double* __restrict__ in,
                                                                long long in_vs, double* __restrict__ mirror, long long mirror_vs,
                                                                double* __restrict__ out, long long out_vs, double2* scratch,
                                                                int* counter, int chunk) {
  typedef R5Cfg<LOG2L2> C;
  constexpr int L2 = C::L2, T = C::T;
  SB_DYN_SMEM(double2, sm);
  __shared__ uint32_t s_tmem;
  __shared__ int s_chunk;
  const int tid = threadIdx.x, team = tid / T, tl = tid - team * T;
  const int grp = team / 3, r = team - 3 * grp;
  GB3 gb;
  gb.b0 = sm + (size_t)(3 * grp) * C::B::LP; gb.b1 = gb.b0 + C::B::LP; gb.b2 = gb.b1 + C::B::LP;
  gb.mine = gb.b0 + (size_t)r * C::B::LP;
  double2* const ctab = sm + (size_t)C::NTEAMS * C::B::LP;
  double* const stg_all = reinterpret_cast<double*>(ctab + 32);
  const double* const stg = stg_all + (size_t)grp * C::STGF;
  sb_mbar_t* const mbar = reinterpret_cast<sb_mbar_t*>(reinterpret_cast<char*>(stg_all + (size_t)C::NGROUPS * C::STGF) + 16 * grp);
  double2* const park = scratch + ((size_t)blockIdx.x * C::NGROUPS + grp) * (2 * C::CH);   // [2 halves][CH]
  if (tid < 32) sb_tmem_alloc(&s_tmem, 512);
  if (r == 0 && tl == 0) sb_mbar_init(mbar, 1);
  if (tid < 32) ctab[tid] = twp[15 * T + 15 * C::B::MS1 + 2 * T + tid];
  sb_fence_mbar_init();
  sb_tmem_fence_before_sync();
  __syncthreads();
  sb_tmem_fence_after_sync();
  const uint32_t tb = sb_tmem_warp_base(s_tmem);
  if (tid < 128) fill_class5<LOG2L2>(s_tmem, twp, tid);
  unsigned phase = 0;
  int cur_ring = -1;
  const int total = nwork * nvars;
  const int per_cta = (total + (int)gridDim.x - 1) / (int)gridDim.x;
  int w_begin = (int)blockIdx.x * per_cta, w_end = w_begin + per_cta < total ? w_begin + per_cta : total;
  for (;;) {
  if (counter) {
    if (tid == 0) s_chunk = atomicAdd(counter, 1);
    __syncthreads();
    w_begin = s_chunk * chunk;
    w_end = w_begin + chunk < total ? w_begin + chunk : total;
    __syncthreads();
    if (w_begin >= total) break;
  }
  for (int w = w_begin; w < w_end; ++w) {
    const int item = w / nvars, v_ = w - item * nvars;
    const LWork wk = work[item];
    const RingPlan pl = plans[wk.r];
    const int n = pl.n, m = pl.m;
    const double2* chirp_g = reinterpret_cast<const double2*>(blob + pl.off);
    const double2* FH_g = chirp_g + 3 * m;
    const double2* AF = reinterpret_cast<const double2*>(blob + pl.off2) + (size_t)4 * m;
    if (wk.r != cur_ring) {
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      if (tid < 128) fill_ring5<LOG2L2>(s_tmem, FH_g, chirp_g, m, tid);
      sb_tmem_fence_before_sync();
      __syncthreads();
      sb_tmem_fence_after_sync();
      cur_ring = wk.r;
    }
    const long long hoff = g.ring_hoff[wk.r];
    const double* src = in + (long long)v_ * in_vs + (long long)g.bz * hoff;
    double* mir = mirror ? mirror + (long long)v_ * mirror_vs + (long long)g.bz * hoff : nullptr;
    double* dst = out + (long long)v_ * out_vs + g.ring_woff[wk.r];
    auto request = [&](int row) {
      sb_fence_proxy_async();
      sb_mbar_expect_tx(mbar, (unsigned)n * 8u);
      sb_bulk_g2s(const_cast<double*>(stg), src + (long long)(wk.row0 + row) * n, (unsigned)n * 8u, mbar);
    };
    group_sync5<T>(grp);
    if (r == 0 && tl == 0 && grp < wk.nrows) request(grp);
    for (int row = grp; row < wk.nrows; row += C::NGROUPS) {
      sb_mbar_wait(mbar, phase);
      phase ^= 1u;
      for (int half = 0; half < 2; ++half) {
        group_sync5<T>(grp);          // the buffers are free (previous half's outputs parked)
        {
          const double* rp = stg + 2 * half;
          double* mp = mir ? mir + (long long)(wk.row0 + row) * n + 2 * half : nullptr;
          uint32_t r0[16], r1[16];
          sb_tmem_ld16(tb + C::C_CH3 + 32 * r, r0);
          sb_tmem_ld16(tb + C::C_CH3 + 32 * r + 16, r1);
          double2 x[8];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) {
            const int a0 = (r == 2 ? L2 : 0) + (r == 1 ? L2 / 2 : 0) + n1 * T + tl, a = a0 < m ? a0 : m - 1;
            x[n1] = *reinterpret_cast<const double2*>(rp + 4 * a);
          }
          sb_tmem_wait_ld16(r0);
          sb_tmem_wait_ld16(r1);
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) {
            const int i = n1 * T + tl;
            const int a0 = (r == 2 ? L2 : 0) + (r == 1 ? L2 / 2 : 0) + i, a = a0 < m ? a0 : m - 1;
            if (mp && a0 < m) *reinterpret_cast<double2*>(mp + 4 * a) = x[n1];
            const double2 y = cm(x[n1], n1 < 4 ? tm_c(r0, n1 & 3) : tm_c(r1, n1 & 3));
            // team r holds x[r L2/2 + i] at gb[r][i]: thirds of L2/2 entries each... (layout of k_fwd_l3: third r at gb[r][0 .. L2/2))
            gb.mine[i] = a0 < m ? y : make_double2(0.0, 0.0);
          }
        }
        group_sync5<T>(grp);          // x is complete
        if (half == 1 && r == 0 && tl == 0 && row + C::NGROUPS < wk.nrows) request(row + C::NGROUPS);   // staging row read by all
        __syncwarp();
        conv5<LOG2L2>(gb, tb, ctab, r, grp, team, tl);
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const int i = (r == 1 ? L2 / 2 : 0) + n1 * T + tl;
          const int a = (r == 2 ? L2 : 0) + i;
          if (a < m) park[(size_t)half * C::CH + a] = combine5(gb, r, i);
        }
      }
      group_sync5<T>(grp);            // both halves are parked (global writes of the group are ordered by the barrier)
      {
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
  if (!counter) break;
  }
  sb_tmem_fence_before_sync();
  __syncthreads();
  if (tid < 32) sb_tmem_dealloc(s_tmem, 512);
}

// =====================================================================================
// launchers
// =====================================================================================
// Overlapped step (sb_api.cpp, tile_step_overlapped): the FP64-bound ring FFTs run on a high-priority stream with a grid
// that leaves `sm_reserve` SMs to the HBM-bound kernels of the other stream, and take their work from an atomic counter
// so that a CTA that becomes resident late (or never) does not hold a fixed share.
static int fft_grid_cap(const LaunchCtx& c) {
  const int n = sb_sm_count() - (c.counters ? c.sm_reserve : 0);
  return n < 1 ? 1 : n;
}
static int* take_counter(const LaunchCtx& c) {
  if (!c.counters || !c.counter_next) return nullptr;
  int* p = c.counters + (*c.counter_next)++ % c.ncounters;
  cudaMemsetAsync(p, 0, sizeof(int), c.stream);
  return p;
}

bool fft4_supported(int L, bool forward) {
  static const char* env = std::getenv("SB_FFT4");
  if (env && std::atoi(env) == 0) return false;   // A/B switch: v2 kernels
  (void)forward;
  return L == 512 || L == 1024 || L == 2048 || L == 4096;
}

template <int LOG2L>
static void launch_inv4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs,
                        long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  const size_t smem = R4Cfg<LOG2L>::SMEM;
  auto kern = out_is_phys == 2 ? k_inv_l4<LOG2L, true> : k_inv_l4<LOG2L, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, cap = fft_grid_cap(c), gx = total < cap ? total : cap;
  int* counter = take_counter(c);     // non-null: the launch shares the chip with another stream's kernels
  SB_LAUNCH(kern, dim3(gx), dim3(512), smem, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0,
            c.need.lmask, counter, c.fft_chunk);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_l4 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_inv_l4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  switch (L) {
    case 512: launch_inv4<9>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 1024: launch_inv4<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 2048: launch_inv4<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 4096: launch_inv4<12>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    default: throw std::runtime_error("launch_inv_l4: unsupported convolution length");
  }
}


template <int LOG2L>
static void launch_fwd4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs,
                        double* mirror, long long mirror_vs, double* out, long long out_vs) {
  const size_t smem = R4FCfg<LOG2L>::SMEMF;
  cudaError_t e = cudaFuncSetAttribute(k_fwd_l4<LOG2L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, cap = fft_grid_cap(c), gx = total < cap ? total : cap;
  int* counter = take_counter(c);
  SB_LAUNCH(k_fwd_l4<LOG2L>, dim3(gx), dim3(512), smem, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_vs, mirror, mirror_vs, out, out_vs, counter, c.fft_chunk);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_l4 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_fwd_l4(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs) {
  switch (L) {
    case 512: launch_fwd4<9>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs); break;
    case 1024: launch_fwd4<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs); break;
    case 2048: launch_fwd4<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs); break;
    case 4096: launch_fwd4<12>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs); break;
    default: throw std::runtime_error("launch_fwd_l4: unsupported convolution length");
  }
}

// ---- composite lengths with the v4 data movement
bool fft5_supported(int L) {
  static const char* env = std::getenv("SB_FFT4");
  if (env && std::atoi(env) == 0) return false;
  return L == 3072 || L == 6144;
}

template <int LOG2L2>
static void launch_inv5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs,
                        long long in_vs, double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  typedef R5Cfg<LOG2L2> C;
  cudaError_t e = cudaFuncSetAttribute(k_inv_l5<LOG2L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEMI);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, cap = fft_grid_cap(c), gx = total < cap ? total : cap;
  int* counter = take_counter(c);
  SB_LAUNCH(k_inv_l5<LOG2L2>, dim3(gx), dim3(C::NT), C::SMEMI, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0,
            c.need.lmask, counter, c.fft_chunk);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_l5 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_inv_l5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_fs, long long in_vs,
                   double* out, long long out_fs, long long out_vs, int out_is_phys, int var0) {
  switch (L / 3) {
    case 1024: launch_inv5<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    case 2048: launch_inv5<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_fs, in_vs, out, out_fs, out_vs, out_is_phys, var0); break;
    default: throw std::runtime_error("launch_inv_l5: unsupported convolution length");
  }
}

template <int LOG2L2>
static void launch_fwd5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, const double* twp,
                        const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs,
                        double* mirror, long long mirror_vs, double* out, long long out_vs, double* scratch) {
  typedef R5Cfg<LOG2L2> C;
  cudaError_t e = cudaFuncSetAttribute(k_fwd_l5<LOG2L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEMF);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int total = nwork * nvars, cap = fft_grid_cap(c), gx = total < cap ? total : cap;
  int* counter = take_counter(c);
  SB_LAUNCH(k_fwd_l5<LOG2L2>, dim3(gx), dim3(C::NT), C::SMEMF, c.stream, g, work, nwork, nvars,
            reinterpret_cast<const double2*>(twp), plans, blob, in, in_vs, mirror, mirror_vs, out, out_vs,
            reinterpret_cast<double2*>(scratch), counter, c.fft_chunk);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_l5 launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_fwd_l5(const LaunchCtx& c, const DevGrid& g, const LWork* work, int nwork, int L, const double* twp,
                   const RingPlan* plans, const double* blob, int nvars, const double* in, long long in_vs, double* mirror,
                   long long mirror_vs, double* out, long long out_vs, double* scratch) {
  switch (L / 3) {
    case 1024: launch_fwd5<10>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs, scratch); break;
    case 2048: launch_fwd5<11>(c, g, work, nwork, twp, plans, blob, nvars, in, in_vs, mirror, mirror_vs, out, out_vs, scratch); break;
    default: throw std::runtime_error("launch_fwd_l5: unsupported convolution length");
  }
}

}  // namespace sb
