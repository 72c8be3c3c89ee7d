// sb_chebmma.cu -- vertical Chebyshev synthesis on the FP64 tensor cores (DMMA m8n8k4).
//
// The inverse Chebyshev stage of K3 is a true small-matrix contraction: for every column,
//   out_f[z] = sum_k M[z][k] a_f[k]      (zDim x b_zDim, 7 outputs per column at RLZ),
// i.e. Out^T[cols x z] = A[cols x k] . M^T[k x z].  ncu showed the FMA formulation issue-bound
// (FP64 pipe 8 % busy, 255 registers for the matrix rows), so the contraction is issued as DMMA:
// 256 FMAs per instruction and 1 register per operand fragment.  DMMA shares the FP64 pipe with
// DFMA on B200 (measured 37.0 vs 36.7 TFLOP/s, profiles/fp64_peak_b200.json) -- the gain is issue
// slots and registers, not peak.
//
// Parity: without vertical BCs M[zDim-1-z][k] = sigma (-1)^k M[z][k] (sigma = -1 for d/dz), so only the
// lower half of the levels is computed, split into even-mode (E) and odd-mode (O) products:
//   out[z] = E + O,   out[zDim-1-z] = sigma (E - O).
// warp = (z-tile of 8 levels, column group); B fragments (the matrix) stay in registers for the
// whole kernel; A fragments come from a [mode][column] smem tile whose stride (40) makes both the
// transposing stores and the fragment loads bank-conflict-free.
#include "sb_internal.hpp"
#include "sb_eqcore.hpp"

#include <cstdint>
#include <cstdlib>
#include <stdexcept>

namespace sb {

#define ZM_KT 6                 // k-tiles (of 4 modes) per parity  -> b_zDim <= 48
#define ZM_KK (4 * ZM_KT)       // modes per parity held in smem
// column stride of the smem tile = COLS + 4 (== 4 mod 16: the half-warp fragment load q*CS + i, q,i = 0..3
// touches 16 distinct banks)

// Persistent CTAs, warps = (4 level tiles) x (COLS/8 column groups); the [mode][column] tile of the NEXT work
// item streams into the second smem buffer with cp.async (LDGSTS) while the tensor cores work on the
// current one.  CB = bytes per async copy (16 when every row is 16-byte aligned, else 8).  COLS = 32: one
// 512-thread CTA per SM; COLS = 16: two 256-thread CTAs per SM that drift out of phase, so one CTA's copy
// issue / barrier overlaps the other's DMMA burst (each 32-column ZTile is split in two).
template <int CB, int COLS>
__global__ void __launch_bounds__(COLS * 16, 32 / COLS) k_inv_z_mma(DevGrid g, const ZTile* __restrict__ tiles, int ntiles,
                                                             int nvars, int var0, int nfields,
                                                             const double* __restrict__ in, long long in_fs,
                                                             long long in_vs, double* __restrict__ phys,
                                                             const double* __restrict__ parB, unsigned zmask,
                                                             unsigned zsel) {
  SB_DYN_SMEM(double, a);       // [2 buffers][nfields][2 parities][ZM_KK][ZM_CS]
  constexpr int ZM_CS = COLS + 4, ZM_THREADS = COLS * 16, SPLIT = 32 / COLS;
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1, nzt = zh >> 3, ncg = (ZM_THREADS / 32) / nzt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, i = lane >> 2;
  const int zt = warp % nzt, cg = warp / nzt;
  const int bufsz = nfields * 2 * ZM_KK * ZM_CS;
  double B[3][2][ZM_KT];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int kt = 0; kt < ZM_KT; ++kt) B[m][p][kt] = parB[((((m * 2 + p) * ZM_KT + kt) * 4) + zt) * 32 + lane];
  for (int j = tid; j < 2 * bufsz; j += ZM_THREADS) a[j] = 0.0;   // zero padding (modes >= bz) stays zero
  __syncthreads();
  const long long slotN = (long long)g.V * g.N;
  const int z0 = zt * 8 + 2 * q;                 // this lane's pair of levels (z0, z0+1) and mirror (zDim-2-z0, +1)
  const int nsub = ntiles * SPLIT;               // sub-tiles of COLS columns
  const int nwork = nsub * nvars;
  constexpr int CPR = COLS * 8 / CB, CE = CB / 8;     // copies per row, doubles per copy
  // per input row (field f, mode zb): {f, zb, offset of the row in the smem tile}; built once, so the copy
  // loop is a table look-up + two multiply-adds instead of two integer divisions per 16 bytes
  int4* rowtab = reinterpret_cast<int4*>(a + 2 * bufsz);
  const int nrow = nfields * bz;
  for (int r = tid; r < nrow; r += ZM_THREADS) {
    const int f = r / bz, zb = r - f * bz;
    rowtab[r] = make_int4(f, zb, ((f * 2 + (zb & 1)) * ZM_KK + (zb >> 1)) * ZM_CS, 0);
  }
  __syncthreads();
  auto desc = [&](int w) {       // COLS-column sub-tile of a 32-column ZTile
    const int s_ = w % nsub;
    ZTile t = tiles[s_ / SPLIT];
    const int off = (s_ % SPLIT) * COLS;
    t.hcol0 += off; t.out_base += off;
    t.ncols = t.ncols - off < COLS ? t.ncols - off : COLS;   // may be <= 0: empty sub-tile
    return t;
  };
  auto issue = [&](int w, const ZTile& ztile, double* dst) {
    const int v = w / nsub;
    const double* src = in + (long long)v * in_vs + ztile.out_base;
    const int total = nrow * CPR;
    for (int c = tid; c < total; c += ZM_THREADS) {
      const int row = c / CPR, col = (c - row * CPR) * CE;
      const int4 rt = rowtab[row];
      if (col < ztile.ncols && ((zmask >> rt.x) & 1)) {   // fields nobody reads are not fetched
        const double* sp = src + (long long)rt.x * in_fs + (long long)rt.y * ztile.out_stride + col;
        double* dp = dst + rt.z + col;
        if (CB == 16) sb_cp_async16(dp, sp); else sb_cp_async8(dp, sp);
      }
    }
    sb_cp_commit();
  };
  // the next tile's descriptor is fetched before the barrier, so its L2 round trip overlaps the wait
  const int G = gridDim.x;
  int w = blockIdx.x, cur = 0;
  ZTile zt0 = desc(w < nwork ? w : 0), zt1;
  if (w < nwork) issue(w, zt0, a);
  for (; w < nwork; w += G) {
    zt1 = desc(w + G < nwork ? w + G : w);
    sb_cp_wait<0>();
    __syncthreads();            // tile w has landed; everybody is done with the other buffer
    if (w + G < nwork) issue(w + G, zt1, a + (cur ^ 1) * bufsz);
    const int v = w / nsub;
    const ZTile ztile = zt0;
    zt0 = zt1;
    const double* ab = a + cur * bufsz;
    double* const pv = phys + (long long)(var0 + v) * g.N;
    for (int ct = cg; ct < COLS / 8; ct += ncg) {
      const int c = ct * 8 + i;
      const bool live = c < ztile.ncols;
      double* const o = pv + ((long long)ztile.hcol0 + c) * zDim;
      const double* ap = ab + q * ZM_CS + c;        // + (parity*ZM_KK + kt*4) * ZM_CS per fragment
      // ---- field 0: value, d/dz, d2/dz2 share the A fragments (zsel: which of the three anybody reads)
      if (zsel == 7u) {
        double E0[2] = {0, 0}, O0[2] = {0, 0}, E1[2] = {0, 0}, O1[2] = {0, 0}, E2[2] = {0, 0}, O2[2] = {0, 0};
#pragma unroll
        for (int kt = 0; kt < ZM_KT; ++kt) {
          const double aE = ap[(kt * 4) * ZM_CS];
          const double aO = ap[(ZM_KK + kt * 4) * ZM_CS];
          sb_dmma(E0[0], E0[1], aE, B[0][0][kt]);
          sb_dmma(O0[0], O0[1], aO, B[0][1][kt]);
          sb_dmma(E1[0], E1[1], aE, B[1][0][kt]);
          sb_dmma(O1[0], O1[1], aO, B[1][1][kt]);
          sb_dmma(E2[0], E2[1], aE, B[2][0][kt]);
          sb_dmma(O2[0], O2[1], aO, B[2][1][kt]);
        }
        if (live) {
          *reinterpret_cast<double2*>(o + z0) = make_double2(E0[0] + O0[0], E0[1] + O0[1]);
          *reinterpret_cast<double2*>(o + zDim - 2 - z0) = make_double2(E0[1] - O0[1], E0[0] - O0[0]);
          double* oz = o + (long long)nfields * slotN;
          *reinterpret_cast<double2*>(oz + z0) = make_double2(E1[0] + O1[0], E1[1] + O1[1]);
          *reinterpret_cast<double2*>(oz + zDim - 2 - z0) = make_double2(-(E1[1] - O1[1]), -(E1[0] - O1[0]));
          double* ozz = oz + slotN;
          *reinterpret_cast<double2*>(ozz + z0) = make_double2(E2[0] + O2[0], E2[1] + O2[1]);
          *reinterpret_cast<double2*>(ozz + zDim - 2 - z0) = make_double2(E2[1] - O2[1], E2[0] - O2[0]);
        }
      } else {
        // one matrix at a time (same DMMA order per output as above, so the values are bit-identical)
#pragma unroll
        for (int mtx = 0; mtx < 3; ++mtx) {
          if (!((zsel >> mtx) & 1)) continue;
          double E[2] = {0, 0}, O[2] = {0, 0};
#pragma unroll
          for (int kt = 0; kt < ZM_KT; ++kt) {
            sb_dmma(E[0], E[1], ap[(kt * 4) * ZM_CS], B[mtx][0][kt]);
            sb_dmma(O[0], O[1], ap[(ZM_KK + kt * 4) * ZM_CS], B[mtx][1][kt]);
          }
          if (live) {
            const double sg = (mtx == 1) ? -1.0 : 1.0;   // d/dz is odd under z -> zDim-1-z
            double* om = o + (mtx ? (long long)(nfields + mtx - 1) * slotN : 0);
            *reinterpret_cast<double2*>(om + z0) = make_double2(E[0] + O[0], E[1] + O[1]);
            *reinterpret_cast<double2*>(om + zDim - 2 - z0) = make_double2(sg * (E[1] - O[1]), sg * (E[0] - O[0]));
          }
        }
      }
      // ---- remaining fields: value matrix only
      for (int f = 1; f < nfields; ++f) {
        if (!((zmask >> f) & 1)) continue;
        const double* af = ap + (size_t)f * 2 * ZM_KK * ZM_CS;
        double E[2] = {0, 0}, O[2] = {0, 0};
#pragma unroll
        for (int kt = 0; kt < ZM_KT; ++kt) {
          sb_dmma(E[0], E[1], af[(kt * 4) * ZM_CS], B[0][0][kt]);
          sb_dmma(O[0], O[1], af[(ZM_KK + kt * 4) * ZM_CS], B[0][1][kt]);
        }
        if (live) {
          double* of = o + (long long)f * slotN;
          *reinterpret_cast<double2*>(of + z0) = make_double2(E[0] + O[0], E[1] + O[1]);
          *reinterpret_cast<double2*>(of + zDim - 2 - z0) = make_double2(E[1] - O[1], E[0] - O[0]);
        }
      }
    }
    cur ^= 1;
  }
}

bool inv_z_mma_ok(const DevGrid& g, int nfields) {
  return (g.zDim == 16 || g.zDim == 32 || g.zDim == 64) && g.bz <= 2 * ZM_KK && nfields >= 1 && nfields <= 5;
}

// parB[3][2][ZM_KT][4][32]: B fragment of DMMA = M_mat[z = zt*8 + lane/4][mode = 2*(kt*4 + lane%4) + parity]
void build_inv_z_mma_tables(int zDim, int bz, const double* T0, const double* T1, const double* T2, std::vector<double>& out) {
  out.assign((size_t)3 * 2 * ZM_KT * 4 * 32, 0.0);
  const double* T[3] = {T0, T1, T2};
  const int nzt = (zDim / 2) / 8;
  for (int m = 0; m < 3; ++m)
    for (int p = 0; p < 2; ++p)
      for (int kt = 0; kt < ZM_KT; ++kt)
        for (int zt = 0; zt < nzt && zt < 4; ++zt)
          for (int lane = 0; lane < 32; ++lane) {
            const int z = zt * 8 + lane / 4, mode = 2 * (kt * 4 + lane % 4) + p;
            if (mode < bz && z < zDim / 2)
              out[((((size_t)(m * 2 + p) * ZM_KT + kt) * 4) + zt) * 32 + lane] = T[m][(size_t)z * zDim + mode];
          }
}

template <int CB, int COLS>
static void launch_inv_z_mma_t(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                               int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                               const double* parB) {
  const size_t smem = (size_t)2 * nfields * 2 * ZM_KK * (COLS + 4) * sizeof(double) + (size_t)nfields * g.bz * 16;
  cudaError_t e = cudaFuncSetAttribute(k_inv_z_mma<CB, COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  const int nwork = ntiles * (32 / COLS) * nvars;
  const int cap = sb_sm_count() * (32 / COLS);
  const int gx = nwork < cap ? nwork : cap;
  SB_LAUNCH((k_inv_z_mma<CB, COLS>), dim3(gx), dim3(COLS * 16), smem, c.stream, g, tiles, ntiles, nvars, var0, nfields, in,
            in_fstride, in_vstride, phys, parB, c.need.zmask, c.need.zsel);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_z_mma launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

void launch_inv_z_mma(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                      int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                      const double* parB) {
  ProfScope prof_scope_(c, "inv_z");
  // 16-byte async copies need every [mode] row of every tile 16-byte aligned
  const bool al16 = ((uintptr_t)in % 16 == 0) && in_fstride % 2 == 0 && in_vstride % 2 == 0 && (g.has_l || g.rDim % 2 == 0);
  static const int cols = std::getenv("SB_INVZ_COLS") ? std::atoi(std::getenv("SB_INVZ_COLS")) : 16;   // A/B switch
  if (cols == 8 && g.zDim == 64) {   // 4 level tiles x 1 column group = 4 warps, four CTAs per SM
    if (al16) launch_inv_z_mma_t<16, 8>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
    else launch_inv_z_mma_t<8, 8>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
  } else if (cols == 32) {
    if (al16) launch_inv_z_mma_t<16, 32>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
    else launch_inv_z_mma_t<8, 32>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
  } else {
    if (al16) launch_inv_z_mma_t<16, 16>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
    else launch_inv_z_mma_t<8, 16>(c, g, tiles, ntiles, nvars, var0, nfields, in, in_fstride, in_vstride, phys, parB);
  }
}

// =====================================================================================
// K3's last stage + K4 in one kernel for LinearAdvectionRLZ (src/testModels.jl:75-98): the Chebyshev
// synthesis of the seven rows the equation reads -- h: value, r, rr, lambda, lambda-lambda; u, v: value -- stays in
// registers, the tendency and the Euler/AB2/AB3 step (src/semiimplicit.jl:672-698) run on it, and only var_np1 and
// expdot_n are written: the derivative slots never exist in HBM.  Same DMMA order and the same tendency / AB3
// expressions (sb_eqcore.hpp) as k_inv_z_mma + k_pointwise, so the state is bit-identical to the two-kernel path.
// in: [ZF_NF field rows][bz][ring rows] (SZ layout), field row s = 0..4: h fields, 5: u, 6: v.
// 16-column tiles, 256 threads (4 level tiles x 2 column groups at 64 levels), two CTAs per SM, cp.async double buffer.
// =====================================================================================
#define ZF_NF 7
template <int CB>
__global__ void __launch_bounds__(256, 2) k_inv_z_advection(DevGrid g, const ZTile* __restrict__ tiles, int ntiles,
                                                            const double* __restrict__ in, long long in_fs,
                                                            const double* __restrict__ parB, EqParams p, ModelArrays arr, int t) {
  SB_DYN_SMEM(double, a);       // [2 buffers][ZF_NF][2 parities][ZM_KK][ZM_CS]
  constexpr int COLS = 16, ZM_CS = COLS + 4, ZM_THREADS = COLS * 16, SPLIT = 32 / COLS;
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1, nzt = zh >> 3, ncg = (ZM_THREADS / 32) / nzt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, i = lane >> 2;
  const int zt = warp % nzt, cg = warp / nzt;
  constexpr int bufsz = ZF_NF * 2 * ZM_KK * ZM_CS;
  double B[2][ZM_KT];           // value matrix only
#pragma unroll
  for (int par = 0; par < 2; ++par)
#pragma unroll
    for (int kt = 0; kt < ZM_KT; ++kt) B[par][kt] = parB[(((par * ZM_KT + kt) * 4) + zt) * 32 + lane];
  for (int j = tid; j < 2 * bufsz; j += ZM_THREADS) a[j] = 0.0;   // zero padding (modes >= bz) stays zero
  __syncthreads();
  const int z0 = zt * 8 + 2 * q;                 // this lane's pair of levels (z0, z0+1) and mirror (zDim-2-z0, +1)
  const int nwork = ntiles * SPLIT;
  constexpr int CPR = COLS * 8 / CB, CE = CB / 8;
  int4* rowtab = reinterpret_cast<int4*>(a + 2 * bufsz);
  const int nrow = ZF_NF * bz;
  for (int r = tid; r < nrow; r += ZM_THREADS) {
    const int f = r / bz, zb = r - f * bz;
    rowtab[r] = make_int4(f, zb, ((f * 2 + (zb & 1)) * ZM_KK + (zb >> 1)) * ZM_CS, 0);
  }
  __syncthreads();
  auto desc = [&](int w) {       // COLS-column sub-tile of a 32-column ZTile
    ZTile tl = tiles[w / SPLIT];
    const int off = (w % SPLIT) * COLS;
    tl.hcol0 += off; tl.out_base += off;
    tl.ncols = tl.ncols - off < COLS ? tl.ncols - off : COLS;   // may be <= 0: empty sub-tile
    return tl;
  };
  auto issue = [&](const ZTile& ztile, double* dst) {
    const double* src = in + ztile.out_base;
    const int total = nrow * CPR;
    for (int c = tid; c < total; c += ZM_THREADS) {
      const int row = c / CPR, col = (c - row * CPR) * CE;
      const int4 rt = rowtab[row];
      if (col < ztile.ncols) {
        const double* sp = src + (long long)rt.x * in_fs + (long long)rt.y * ztile.out_stride + col;
        double* dp = dst + rt.z + col;
        if (CB == 16) sb_cp_async16(dp, sp); else sb_cp_async8(dp, sp);
      }
    }
    sb_cp_commit();
  };
  const long long N = g.N;
  const double ts = p.ts, K = p.K;
  const int G = gridDim.x;
  int w = blockIdx.x, cur = 0;
  ZTile zt0 = desc(w < nwork ? w : 0), zt1;
  // the tile's radius (a ZTile lies inside one ring) is fetched with the descriptor, one tile ahead: h2r -> rad is a
  // dependent pair of global loads that used to sit right behind the barrier (12 % of the stall samples)
  auto radius = [&](const ZTile& z) { return z.ncols > 0 ? z.rad : 1.0; };     // (rides in the descriptor: no dependent lookups)
  double r0 = radius(zt0), r1;
  if (w < nwork) issue(zt0, a);
  for (; w < nwork; w += G) {
    zt1 = desc(w + G < nwork ? w + G : w);
    r1 = radius(zt1);
    sb_cp_wait<0>();
    __syncthreads();            // tile w has landed; everybody is done with the other buffer
    if (w + G < nwork) issue(zt1, a + (cur ^ 1) * bufsz);
    const ZTile ztile = zt0;
    zt0 = zt1;
    const double* ab = a + cur * bufsz;
    const double ri = 1.0 / r0, ri2 = ri * ri;
    r0 = r1;
    for (int ct = cg; ct < COLS / 8; ct += ncg) {
      const int c = ct * 8 + i;
      const bool live = c < ztile.ncols;
      const double* ap = ab + q * ZM_CS + c;
      // f[s][0..3]: field row s at levels z0, z0+1, zDim-2-z0, zDim-1-z0
      double f[ZF_NF][4];
#pragma unroll
      for (int s = 0; s < ZF_NF; ++s) {
        const double* af = ap + (size_t)s * 2 * ZM_KK * ZM_CS;
        double E[2] = {0, 0}, O[2] = {0, 0};
#pragma unroll
        for (int kt = 0; kt < ZM_KT; ++kt) {
          sb_dmma(E[0], E[1], af[(kt * 4) * ZM_CS], B[0][kt]);
          sb_dmma(O[0], O[1], af[(ZM_KK + kt * 4) * ZM_CS], B[1][kt]);
        }
        f[s][0] = E[0] + O[0]; f[s][1] = E[1] + O[1]; f[s][2] = E[1] - O[1]; f[s][3] = E[0] - O[0];
      }
      if (live) {
        const long long col0 = ((long long)ztile.hcol0 + c) * zDim;
#pragma unroll
        for (int hm = 0; hm < 2; ++hm) {          // the level pair and its mirror pair
          const long long o0 = col0 + (hm ? zDim - 2 - z0 : z0);
          // the six history loads of this pair first (the stores below may alias them as far as the compiler knows, so
          // it would otherwise serialise load -> use -> store per variable and pay the memory latency each time)
          double2 h1[3], h2[3];
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            h1[v] = t >= 2 ? *reinterpret_cast<const double2*>(arr.exp_nm1 + (long long)v * N + o0) : make_double2(0.0, 0.0);
            h2[v] = t >= 3 ? *reinterpret_cast<const double2*>(arr.exp_nm2 + (long long)v * N + o0) : make_double2(0.0, 0.0);
          }
          double e[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int x = 2 * hm + k;
            e[k] = advection_rl_tendency(f[5][x], f[6][x], f[1][x], f[3][x], f[2][x], f[4][x], ri, ri2, K);
          }
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const long long o = (long long)v * N + o0;
            const int s = v == 0 ? 0 : 4 + v;
            const double2 f1 = h1[v], f2 = h2[v];
            const double fn0 = v == 0 ? e[0] : 0.0, fn1 = v == 0 ? e[1] : 0.0;
            *reinterpret_cast<double2*>(arr.exp_n + o) = make_double2(fn0, fn1);
            *reinterpret_cast<double2*>(arr.var_np1 + o) =
                make_double2(ab_step(t, ts, f[s][2 * hm], fn0, f1.x, f2.x), ab_step(t, ts, f[s][2 * hm + 1], fn1, f1.y, f2.y));
          }
        }
      }
    }
    cur ^= 1;
  }
}


// -------------------------------------------------------------------------------------------------------------------
// The same kernel with Blackwell data movement (the default; SB_INVZ_BULK=0 selects the cp.async version above).
// ncu on the version above (profiles/r1m_ncu_full_k_inv_z_advection_t1.txt): 55 % of the measured HBM peak,
// the warps wait on the history loads of the epilogue (plain LDG behind the DMMA burst), and with 128 registers there
// is no room to request them early.  Here nothing the kernel reads goes through a register or an LSU instruction:
//   * the [mode][column] tile (7 fields x 43 rows of <= 128 bytes) and the six history blocks (exp_nm1, exp_nm2 of the
//     three variables: one 512-byte run per column) are fetched by the bulk-copy engine (cp.async.bulk, SASS UBLKCP)
//     into shared memory and announced by two mbarriers (SASS SYNCS);
//   * per CTA one tile buffer + one history buffer (109 KB, two CTAs per SM): the NEXT tile's rows are requested as soon
//     as the DMMA phase has drained the tile buffer and land during the epilogue, its history as soon as the epilogue
//     has drained the history buffer and lands during the next DMMA phase;
//   * history columns sit 72 doubles apart in shared memory: the epilogue's double2 reads (8 columns x 4 level pairs per
//     warp) are bank-conflict free.
// Arithmetic, DMMA order and store pattern are those of the kernel above: bit-identical results.
// -------------------------------------------------------------------------------------------------------------------
// BLK: the input is in the blocked SZ layout (sb_internal.hpp RowDst): a tile's 16 points x bz modes of a field are one
// contiguous run, even modes first, so the tile arrives with TWO bulk copies per field (14 per tile, 2.7 KB each) instead
// of one 128-byte copy per (field, mode) row (301 per tile).  The copy engine serves a request in some tens of cycles
// whatever its size (B300_MICROARCH "TMA service/SM"): at 301 + 96 requests per tile the kernel was bound by the REQUEST
// rate -- 13.5 GB moved in 3.8 ms, and dropping 6.2 GB of history traffic bought only 6 % -- not by HBM or the DMMAs.
// Rows are 16 doubles, the point index XOR-swizzled by the mode (conflict-free fragment loads without padding); the
// history blocks arrive as one 8 KB run per array (columns 64 doubles apart; the epilogue's four double2 reads per thread
// are then 2-way bank conflicts, which is noise).
#define ZB_HS 72                // column stride (doubles) of a history block in shared memory
template <bool BLK>
__global__ void __launch_bounds__(256, 2) k_inv_z_advection_bulk(DevGrid g, const ZTile* __restrict__ tiles, int ntiles,
                                                                 const double* __restrict__ in, long long in_fs,
                                                                 const double* __restrict__ parB, EqParams p, ModelArrays arr, int t) {
  SB_DYN_SMEM(double, a);       // [ZF_NF][2 parities][ZM_KK][ZM_CS] | history [2][3][16][ZB_HS] | row table | 2 mbarriers
  constexpr int COLS = 16, ZM_CS = BLK ? COLS : COLS + 4, ZM_THREADS = COLS * 16, SPLIT = 32 / COLS;
  constexpr int HS = BLK ? 64 : ZB_HS;               // column stride of a history block (BLK: as in HBM, zDim == 64)
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1, nzt = zh >> 3, ncg = (ZM_THREADS / 32) / nzt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, i = lane >> 2;
  const int zt = warp % nzt, cg = warp / nzt;
  constexpr int bufsz = ZF_NF * 2 * ZM_KK * ZM_CS;
  constexpr int histsz = 6 * COLS * HS;
  double* const hist = a + bufsz;
  int4* const rowtab = reinterpret_cast<int4*>(hist + histsz);
  const int nrow = ZF_NF * bz;
  char* const mb = reinterpret_cast<char*>(rowtab + ((nrow + 1) & ~1));
  sb_mbar_t* const full_in = reinterpret_cast<sb_mbar_t*>(mb);
  sb_mbar_t* const full_hist = reinterpret_cast<sb_mbar_t*>(mb + 16);
  double B[2][ZM_KT];           // value matrix only
#pragma unroll
  for (int par = 0; par < 2; ++par)
#pragma unroll
    for (int kt = 0; kt < ZM_KT; ++kt) B[par][kt] = parB[(((par * ZM_KT + kt) * 4) + zt) * 32 + lane];
  for (int j = tid; j < bufsz; j += ZM_THREADS) a[j] = 0.0;   // zero padding (modes >= bz) stays zero
  for (int r = tid; r < nrow; r += ZM_THREADS) {
    const int f = r / bz, zb = r - f * bz;
    rowtab[r] = make_int4(f, zb, ((f * 2 + (zb & 1)) * ZM_KK + (zb >> 1)) * ZM_CS, 0);
  }
  if (tid == 0) { sb_mbar_init(full_in, 1); sb_mbar_init(full_hist, 1); }
  sb_fence_mbar_init();
  __syncthreads();
  const int z0 = zt * 8 + 2 * q;                 // this lane's pair of levels (z0, z0+1) and mirror (zDim-2-z0, +1)
  const int nwork = ntiles * SPLIT;
  auto desc = [&](int w) {       // COLS-column sub-tile of a 32-column ZTile
    ZTile tl = tiles[w / SPLIT];
    const int off = (w % SPLIT) * COLS;
    tl.hcol0 += off; tl.out_base += off; tl.blk += (long long)off * bz;
    tl.ncols = tl.ncols - off < COLS ? tl.ncols - off : COLS;   // may be <= 0: empty sub-tile
    return tl;
  };
  const long long N = g.N;
  const int nh = (t >= 3 ? 2 : (t >= 2 ? 1 : 0));     // history arrays the AB step reads: exp_nm1 (t >= 2), exp_nm2 (t >= 3)
  // rows of the tile: every thread asks for its rows (<= 2), thread 0 arms the barrier with the byte count of the tile
  // offset of a tile's block in a field row of the blocked layout (fetched one tile ahead, like the radius)
  auto block_of = [&](const ZTile& z) -> long long { return z.blk; };
  auto issue_in = [&](const ZTile& ztile, long long blk) {
    const int nc = ztile.ncols > 0 ? ztile.ncols : 0;
    sb_fence_proxy_async();
    if (BLK) {       // two runs per field: the even modes, then the odd modes, of the tile's 16-point block
      if (tid == 0) sb_mbar_expect_tx(full_in, nc > 0 ? (unsigned)(ZF_NF * bz * COLS * 8) : 0u);
      if (nc > 0 && tid < 2 * ZF_NF) {
        const int f = tid >> 1, par = tid & 1, k0 = (bz + 1) >> 1;
        sb_bulk_g2s(a + (f * 2 + par) * ZM_KK * ZM_CS, in + (long long)f * in_fs + blk + (par ? k0 * COLS : 0),
                    (unsigned)((par ? bz - k0 : k0) * COLS * 8), full_in);
      }
      return;
    }
    if (tid == 0) sb_mbar_expect_tx(full_in, (unsigned)(nrow * nc * 8));
    if (nc > 0) {
      const double* src = in + ztile.out_base;
      for (int r = tid; r < nrow; r += ZM_THREADS) {
        const int4 rt = rowtab[r];
        sb_bulk_g2s(a + rt.z, src + (long long)rt.x * in_fs + (long long)rt.y * ztile.out_stride, (unsigned)(nc * 8), full_in);
      }
    }
  };
  // history: one 8 zDim-byte run per (array, variable, column)
  // u and v carry no tendency (src/testModels.jl:93 writes expdot[:, 1] only): while their history arrays still hold the
  // zeros they were allocated with (arr.passive, cleared by the host when somebody stores into them) they are neither
  // fetched nor is exp_n = 0 written back -- 6 of the kernel's 16 array passes.  The AB step below runs on zeros either way.
  const bool uv_live = (arr.passive & 6u) != 6u;
  const int nvh = uv_live ? 3 : 1;                     // variables whose history is fetched: h | h, u, v
  auto issue_hist = [&](const ZTile& ztile) {
    const int nc = ztile.ncols > 0 ? ztile.ncols : 0;
    sb_fence_proxy_async();
    if (tid == 0) sb_mbar_expect_tx(full_hist, (unsigned)(nh * nvh * nc * zDim * 8));
    if (BLK) {       // the nc columns of a tile are one contiguous run of every [V][N] array
      if (nc > 0 && tid < nh * nvh) {
        const int v = tid % nvh, hk = tid / nvh;
        const double* src = (hk ? arr.exp_nm2 : arr.exp_nm1) + (long long)v * N + (long long)ztile.hcol0 * zDim;
        sb_bulk_g2s(hist + (hk * 3 + v) * COLS * HS, src, (unsigned)(nc * zDim * 8), full_hist);
      }
      return;
    }
    const int total = nh * nvh * nc;
    for (int j = tid; j < total; j += ZM_THREADS) {
      const int c = j % nc, hv = j / nc, v = hv % nvh, hk = hv / nvh;    // hk = 0: exp_nm1, 1: exp_nm2
      const double* src = (hk ? arr.exp_nm2 : arr.exp_nm1) + (long long)v * N + ((long long)ztile.hcol0 + c) * zDim;
      sb_bulk_g2s(hist + ((hk * 3 + v) * COLS + c) * HS, src, (unsigned)(zDim * 8), full_hist);
    }
  };
  const double ts = p.ts, K = p.K;
  const int G = gridDim.x;
  int w = blockIdx.x;
  unsigned ph_in = 0, ph_h = 0;
  ZTile zt0 = desc(w < nwork ? w : 0), zt1;
  auto radius = [&](const ZTile& z) { return z.ncols > 0 ? z.rad : 1.0; };     // (rides in the descriptor: no dependent lookups)
  double r0 = radius(zt0), r1;
  long long b1 = 0;
  if (w < nwork) { issue_in(zt0, block_of(zt0)); issue_hist(zt0); }
  for (; w < nwork; w += G) {
    const bool more = w + G < nwork;
    zt1 = desc(more ? w + G : w);
    r1 = radius(zt1);
    b1 = block_of(zt1);
    const ZTile ztile = zt0;
    zt0 = zt1;
    const double ri = 1.0 / r0, ri2 = ri * ri;
    r0 = r1;
    sb_mbar_wait(full_in, ph_in);          // the tile's rows have landed
    ph_in ^= 1u;
    // COLS / 8 column groups; with nzt = 4 level tiles the 8 warps are 4 x 2: one column group per warp
    const int c = cg * 8 + i;
    const bool live = c < ztile.ncols && cg < COLS / 8;
    double f[ZF_NF][4];         // f[s][0..3]: field row s at levels z0, z0+1, zDim-2-z0, zDim-1-z0
    {
      const double* ap = a + q * ZM_CS + (BLK ? ((cg < COLS / 8 ? c : i) ^ (q << 2)) : (cg < COLS / 8 ? c : i));
      // k-tile outermost: the 14 accumulator pairs (7 fields x 2 parities) are independent chains, so consecutive DMMAs
      // never wait for each other (field-outermost left two chains in flight: 40 % of the DMMA phase's stall samples were
      // fixed-latency waits, profiles/r2za_ncu_full_k_inv_z_advection_bulk_blocked.txt).  Each chain still accumulates its
      // k-tiles in ascending order: the values are bit-identical.
      double E[ZF_NF][2], O[ZF_NF][2];
#pragma unroll
      for (int s = 0; s < ZF_NF; ++s) { E[s][0] = 0.0; E[s][1] = 0.0; O[s][0] = 0.0; O[s][1] = 0.0; }
#pragma unroll
      for (int kt = 0; kt < ZM_KT; ++kt) {
#pragma unroll
        for (int s = 0; s < ZF_NF; ++s) {
          const double* af = ap + (size_t)s * 2 * ZM_KK * ZM_CS;
          sb_dmma(E[s][0], E[s][1], af[(kt * 4) * ZM_CS], B[0][kt]);
          sb_dmma(O[s][0], O[s][1], af[(ZM_KK + kt * 4) * ZM_CS], B[1][kt]);
        }
      }
#pragma unroll
      for (int s = 0; s < ZF_NF; ++s) {
        f[s][0] = E[s][0] + O[s][0]; f[s][1] = E[s][1] + O[s][1]; f[s][2] = E[s][1] - O[s][1]; f[s][3] = E[s][0] - O[s][0];
      }
    }
    __syncthreads();            // every warp has drained the tile buffer
    if (more) issue_in(zt1, b1);    // lands during the epilogue
    sb_mbar_wait(full_hist, ph_h);
    ph_h ^= 1u;
    if (live) {
      const long long col0 = ((long long)ztile.hcol0 + c) * zDim;
#pragma unroll
      for (int hm = 0; hm < 2; ++hm) {          // the level pair and its mirror pair
        const int zl = hm ? zDim - 2 - z0 : z0;
        const long long o0 = col0 + zl;
        double2 h1[3], h2[3];
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const bool hv = v == 0 || uv_live;
          h1[v] = (t >= 2 && hv) ? *reinterpret_cast<const double2*>(hist + (v * COLS + c) * HS + zl) : make_double2(0.0, 0.0);
          h2[v] = (t >= 3 && hv) ? *reinterpret_cast<const double2*>(hist + ((3 + v) * COLS + c) * HS + zl) : make_double2(0.0, 0.0);
        }
        double e[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int x = 2 * hm + k;
          e[k] = advection_rl_tendency(f[5][x], f[6][x], f[1][x], f[3][x], f[2][x], f[4][x], ri, ri2, K);
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const long long o = (long long)v * N + o0;
          const int s = v == 0 ? 0 : 4 + v;
          const double2 f1 = h1[v], f2 = h2[v];
          const double fn0 = v == 0 ? e[0] : 0.0, fn1 = v == 0 ? e[1] : 0.0;
          if (v == 0 || uv_live) *reinterpret_cast<double2*>(arr.exp_n + o) = make_double2(fn0, fn1);
          *reinterpret_cast<double2*>(arr.var_np1 + o) =
              make_double2(ab_step(t, ts, f[s][2 * hm], fn0, f1.x, f2.x), ab_step(t, ts, f[s][2 * hm + 1], fn1, f1.y, f2.y));
        }
      }
    }
    __syncthreads();            // every warp has drained the history buffer
    if (more) issue_hist(zt1);  // lands during the next DMMA phase
  }
}

static bool inv_z_bulk_enabled() {
  static const char* e = std::getenv("SB_INVZ_BULK");
  return !(e && std::atoi(e) == 0);
}

bool inv_z_advection_ok(const DevGrid& g) {
  return g.has_l && g.has_z && g.V == 3 && (g.zDim == 16 || g.zDim == 32 || g.zDim == 64) && g.bz <= 2 * ZM_KK && g.N % 2 == 0;
}

// blocked SZ layout: the bulk-copy kernel at 64 levels (history blocks as in HBM); SB_SZ_BLOCKED=0: A/B switch
bool inv_z_advection_blocked(const DevGrid& g) {
  static const char* e = std::getenv("SB_SZ_BLOCKED");
  return inv_z_advection_ok(g) && inv_z_bulk_enabled() && g.zDim == 64 && !(e && std::atoi(e) == 0);
}

void launch_inv_z_advection(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, const double* in,
                            long long in_fstride, const double* parB, const EqParams& p, const ModelArrays& arr, int t,
                            bool blocked) {
  ProfScope prof_scope_(c, "inv_z_k4");
  const bool al16 = ((uintptr_t)in % 16 == 0) && in_fstride % 2 == 0;
  const size_t smem = (size_t)2 * ZF_NF * 2 * ZM_KK * (16 + 4) * sizeof(double) + (size_t)ZF_NF * g.bz * 16;
  const int nwork = ntiles * 2;
  const int sms = c.mem_grid_sms > 0 ? c.mem_grid_sms : sb_sm_count();     // overlapped step: the SMs the ring FFTs leave free
  const int gx = nwork < sms * 2 ? nwork : sms * 2;
  cudaError_t e;
  // bulk copies need 16-byte aligned rows (in, in_fs even, ring rows are multiples of 4 doubles) and whole history columns
  if (blocked) {
    if (!al16 || !inv_z_advection_blocked(g)) throw std::runtime_error("blocked SZ layout handed to a grid / buffer that cannot use it");
    const size_t smem_b = (size_t)(ZF_NF * 2 * ZM_KK * 16 + 6 * 16 * 64) * sizeof(double) + (size_t)((ZF_NF * g.bz + 1) & ~1) * 16 + 32;
    e = cudaFuncSetAttribute(k_inv_z_advection_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
    SB_LAUNCH(k_inv_z_advection_bulk<true>, dim3(gx), dim3(256), smem_b, c.stream, g, tiles, ntiles, in, in_fstride, parB, p, arr, t);
  } else if (al16 && inv_z_bulk_enabled() && g.zDim <= ZB_HS - 8) {
    const size_t smem_b = (size_t)(ZF_NF * 2 * ZM_KK * (16 + 4) + 6 * 16 * ZB_HS) * sizeof(double) +
                          (size_t)((ZF_NF * g.bz + 1) & ~1) * 16 + 32;
    e = cudaFuncSetAttribute(k_inv_z_advection_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
    SB_LAUNCH(k_inv_z_advection_bulk<false>, dim3(gx), dim3(256), smem_b, c.stream, g, tiles, ntiles, in, in_fstride, parB, p, arr, t);
  } else if (al16) {
    e = cudaFuncSetAttribute(k_inv_z_advection<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
    SB_LAUNCH((k_inv_z_advection<16>), dim3(gx), dim3(256), smem, c.stream, g, tiles, ntiles, in, in_fstride, parB, p, arr, t);
  } else {
    e = cudaFuncSetAttribute(k_inv_z_advection<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
    SB_LAUNCH((k_inv_z_advection<8>), dim3(gx), dim3(256), smem, c.stream, g, tiles, ntiles, in, in_fstride, parB, p, arr, t);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_z_advection launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

// =====================================================================================
// Chebyshev analysis (forward): b[zb] = sum_z fwd[zb][z] u[z].  fwd[zb][zDim-1-z] = (-1)^zb fwd[zb][z], so
// even modes see s = u[z] + u[zDim-1-z] and odd modes d = u[z] - u[zDim-1-z] over the lower half only.
// D[8 columns x 8 modes] += A[8 columns x 4 levels] . B[4 levels x 8 modes]; warp = (column group, parity).
// The [column][level] input tile is contiguous in HBM and arrives by cp.async (double buffered); the
// same tile is mirrored into physical slot 0 (calcTendency's `physical .= var_np1`).
// =====================================================================================
#define FZ_US 68                // level stride of the u tile (== 4 mod 16, 16-byte aligned rows)
#define FZ_THREADS 256

__global__ void __launch_bounds__(FZ_THREADS, 3) k_fwd_z_mma(DevGrid g, const ZTile* __restrict__ tiles, int ntiles, int nvars,
                                                          const double* __restrict__ in, long long in_vs,
                                                          double* __restrict__ mirror, long long mirror_vs,
                                                          double* __restrict__ out, long long out_vs,
                                                          const double* __restrict__ fwdB) {
  SB_DYN_SMEM(double, u);       // [3 stages][32 columns][FZ_US]
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1, nkt = zh >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, i = lane >> 2;
  const int cg = warp & 3, par = warp >> 2;
  double Bf[3][8];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt)
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) Bf[nt][kt] = fwdB[(((par * 3 + nt) * 8 + kt) * 32) + lane];
  const int nwork = ntiles * nvars;
  const int lcpc = (zDim == 64) ? 5 : (zDim == 32 ? 4 : 3), cpc = 1 << lcpc;     // 16-byte copies per column (2^lcpc)
  auto issue = [&](int w, const ZTile& zt, double* dst) {
    const int v = w / ntiles;
    const double* src = in + (long long)v * in_vs + (long long)zt.hcol0 * zDim;
    const int total = zt.ncols << lcpc;
    for (int c = tid; c < total; c += FZ_THREADS) {
      const int col = c >> lcpc, z = (c & (cpc - 1)) * 2;
      sb_cp_async16(dst + col * FZ_US + z, src + (long long)col * zDim + z);
    }
    sb_cp_commit();
  };
  // three-stage ring of tiles: two tiles are always in flight behind the one being transformed
  // (tried: fetching the descriptor of the tile to be issued one iteration earlier, because 13 % of the kernel's stall samples
  //  sit on that dependent load -- 1.21 -> 1.28 ms: the two extra 40-byte descriptor loads per tile cost more than the wait)
  const int G = gridDim.x;
  auto desc = [&](int w) { return w < nwork ? tiles[w % ntiles] : ZTile{}; };
  auto issue_w = [&](int w, int stage) {          // always commits a group (possibly empty): uniform group counting
    if (w < nwork) issue(w, desc(w), u + stage * 32 * FZ_US);
    else sb_cp_commit();
  };
  int w = blockIdx.x, cur = 0;
  issue_w(w, 0);
  issue_w(w + G, 1);
  for (; w < nwork; w += G) {
    const ZTile zt = desc(w);     // (cached: it was fetched when the tile was issued) -- overlaps the wait below
    sb_cp_wait<1>();              // everything but the newest group has landed: tile w is in stage `cur`
    __syncthreads();              // ... for every thread, and everybody has left the stage tile w-G used
    issue_w(w + 2 * G, cur == 0 ? 2 : cur - 1);
    const int v = w / ntiles;
    const double* ub = u + cur * 32 * FZ_US;
    if (mirror) {
      double* mv = mirror + (long long)v * mirror_vs + (long long)zt.hcol0 * zDim;
      const int total = zt.ncols << lcpc;
      for (int c = tid; c < total; c += FZ_THREADS) {
        const int col = c >> lcpc, z = (c & (cpc - 1)) * 2;
        *reinterpret_cast<double2*>(mv + (long long)col * zDim + z) = *reinterpret_cast<const double2*>(ub + col * FZ_US + z);
      }
    }
    const int col = cg * 8 + i;
    const double* uc = ub + col * FZ_US;
    double D[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      if (kt < nkt) {
        const int z = kt * 4 + q;
        const double x = uc[z], y = uc[zDim - 1 - z];
        const double av = par ? x - y : x + y;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) sb_dmma(D[nt][0], D[nt][1], av, Bf[nt][kt]);
      }
    }
    if (col < zt.ncols) {
      double* o = out + (long long)v * out_vs + zt.out_base + col;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const int zb0 = 2 * (nt * 8 + 2 * q) + par, zb1 = zb0 + 2;
        if (zb0 < bz) o[(long long)zb0 * zt.out_stride] = D[nt][0];
        if (zb1 < bz) o[(long long)zb1 * zt.out_stride] = D[nt][1];
      }
    }
    cur = (cur == 2) ? 0 : cur + 1;
  }
}

bool fwd_z_mma_ok(const DevGrid& g) { return (g.zDim == 16 || g.zDim == 32 || g.zDim == 64) && g.bz <= 48; }

// fwdB[2 parities][3 n-tiles][8 k-tiles][32 lanes]: B fragment = fwd[mode = 2*(nt*8 + lane/4) + par][level = kt*4 + lane%4]
void build_fwd_z_mma_tables(int zDim, int bz, const double* fwd /*[bz][zDim]*/, std::vector<double>& out) {
  out.assign((size_t)2 * 3 * 8 * 32, 0.0);
  for (int par = 0; par < 2; ++par)
    for (int nt = 0; nt < 3; ++nt)
      for (int kt = 0; kt < 8; ++kt)
        for (int lane = 0; lane < 32; ++lane) {
          const int mode = 2 * (nt * 8 + lane / 4) + par, z = kt * 4 + lane % 4;
          if (mode < bz && z < zDim / 2) out[(((size_t)(par * 3 + nt) * 8 + kt) * 32) + lane] = fwd[(size_t)mode * zDim + z];
        }
}

void launch_fwd_z_mma(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, const double* in,
                      long long in_vstride, double* mirror, long long mirror_vstride, double* out, long long out_vstride,
                      const double* fwdB) {
  ProfScope prof_scope_(c, "fwd_z");
  const size_t smem = (size_t)3 * 32 * FZ_US * sizeof(double);
  const int nwork = ntiles * nvars;
  if (smem > 48 * 1024) {
    cudaError_t e0 = cudaFuncSetAttribute(k_fwd_z_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e0 != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e0));
  }
  // persistent grid = exactly the CTAs that are resident at once (80 registers x 256 threads -> 3 per SM, although four
  // would fit in shared memory): with 148 x 4 CTAs the last 148 only started when the first wave had finished its share
  // and ran one per SM (ncu: 1.33 waves)
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd_z_mma, FZ_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 3;
  const int cap = (c.mem_grid_sms > 0 ? c.mem_grid_sms : sb_sm_count()) * per_sm;
  const int gx = nwork < cap ? nwork : cap;
  SB_LAUNCH(k_fwd_z_mma, dim3(gx), dim3(FZ_THREADS), smem, c.stream, g, tiles, ntiles, nvars, in, in_vstride, mirror,
            mirror_vstride, out, out_vstride, fwdB);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_fwd_z_mma launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

}  // namespace sb
