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

#include <stdexcept>

namespace sb {

#define ZM_KT 6                 // k-tiles (of 4 modes) per parity  -> b_zDim <= 48
#define ZM_KK (4 * ZM_KT)       // modes per parity held in smem
#define ZM_CS 40                // column stride of the smem tile: 32 columns + 8 pad (== 8 mod 16)

__global__ void __launch_bounds__(256, 2) k_inv_z_mma(DevGrid g, const ZTile* __restrict__ tiles, int ntiles, int var0,
                                                      int nfields, const double* __restrict__ in, long long in_fs,
                                                      long long in_vs, double* __restrict__ phys,
                                                      const double* __restrict__ parB) {
  SB_DYN_SMEM(double, a);       // [nfields][2 parities][ZM_KK][ZM_CS]
  const int zDim = g.zDim, bz = g.bz, zh = zDim >> 1, nzt = zh >> 3, ncg = 8 / nzt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, i = lane >> 2;
  const int zt = warp % nzt, cg = warp / nzt;
  const int v = blockIdx.y;
  double B[3][2][ZM_KT];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int kt = 0; kt < ZM_KT; ++kt) B[m][p][kt] = parB[((((m * 2 + p) * ZM_KT + kt) * 4) + zt) * 32 + lane];
  for (int j = tid; j < nfields * 2 * ZM_KK * ZM_CS; j += 256) a[j] = 0.0;   // zero padding (modes >= bz) stays zero
  const long long slotN = (long long)g.V * g.N;
  double* const pv = phys + (long long)(var0 + v) * g.N;
  const int z0 = zt * 8 + 2 * q;                 // this lane's pair of levels (z0, z0+1) and mirror (zDim-2-z0, +1)
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const ZTile ztile = tiles[t];
    __syncthreads();
    for (int row = warp; row < nfields * bz; row += 8) {
      const int f = row / bz, zb = row - f * bz;
      double val = 0.0;
      if (lane < ztile.ncols)
        val = in[(long long)f * in_fs + (long long)v * in_vs + ztile.out_base + (long long)zb * ztile.out_stride + lane];
      a[((f * 2 + (zb & 1)) * ZM_KK + (zb >> 1)) * ZM_CS + lane] = val;
    }
    __syncthreads();
    for (int ct = cg; ct < 4; ct += ncg) {
      const int c = ct * 8 + i;
      const bool live = c < ztile.ncols;
      double* const o = pv + ((long long)ztile.hcol0 + c) * zDim;
      const double* ap = a + q * ZM_CS + c;        // + (parity*ZM_KK + kt*4) * ZM_CS per fragment
      // ---- field 0: value, d/dz, d2/dz2 share the A fragments
      {
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
      }
      // ---- remaining fields: value matrix only
      for (int f = 1; f < nfields; ++f) {
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

void launch_inv_z_mma(const LaunchCtx& c, const DevGrid& g, const ZTile* tiles, int ntiles, int nvars, int var0,
                      int nfields, const double* in, long long in_fstride, long long in_vstride, double* phys,
                      const double* parB) {
  ProfScope prof_scope_(c, "inv_z");
  size_t smem = (size_t)nfields * 2 * ZM_KK * ZM_CS * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_inv_z_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) throw std::runtime_error(std::string("smem opt-in: ") + cudaGetErrorString(e));
  }
  int gx = ntiles < 148 * 2 * 8 ? ntiles : 148 * 2 * 8;
  SB_LAUNCH(k_inv_z_mma, dim3(gx, nvars), dim3(256), smem, c.stream, g, tiles, ntiles, var0, nfields, in, in_fstride,
            in_vstride, phys, parB);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("k_inv_z_mma launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

}  // namespace sb
