// sb_eqcore.hpp -- arithmetic shared by the equation-set kernels (sb_model.cu) and the kernels that fuse an
// equation set into the last stage of the inverse transform (sb_chebmma.cu): one definition, so both paths
// contract to the same FMAs and leave bit-identical states.
#pragma once
#include "cuda_emu.h"

namespace sb {

// Every product below is either rounded on its own or fused by an explicit fma(), so the compiler has no contraction
// choice left and every kernel that inlines these produces the same bits.

// explicit_timestep for one value (src/semiimplicit.jl:682-696)
__device__ __forceinline__ double ab_step(int t, double ts, double u, double fn, double fnm1, double fnm2) {
  if (t == 1) return fma(ts, fn, u);                                   // u + ts f_n
  if (t == 2) return fma(0.5 * ts, fma(3.0, fn, -fnm1), u);            // u + ts/2 (3 f_n - f_nm1)
  return fma(ts / 12.0, fma(5.0, fnm2, fma(23.0, fn, -(16.0 * fnm1))), u);   // u + ts/12 (23 f_n - 16 f_nm1 + 5 f_nm2)
}

// LinearAdvectionRL (K > 0) / LinearAdvectionRLZ: dh/dt = -u h_r - v h_l / r + K (h_r / r + h_rr + h_ll / r^2)
// (src/testModels.jl:62-68, :93)
// The three divisions by r and r^2 are multiplications by ri = 1/r and ri2 = ri*ri, which the caller computes ONCE per
// radius (a ring shares r; the fused kernel has one r per tile): 12 FP64 divisions per thread of k_inv_z_advection were
// 9.5 % of its stall samples (profiles/r1m_ncu_full_k_inv_z_advection_t1.txt).  Differs from `/ r` by <= 1 ulp per term.
__device__ __forceinline__ double advection_rl_tendency(double u, double v, double hr, double hl, double hrr, double hll,
                                                        double ri, double ri2, double K) {
  const double q = hl * ri;
  const double lap = ((hr * ri) + hrr) + (hll * ri2);
  return fma(K, lap, fma(-v, q, -(u * hr)));
}

}  // namespace sb
