// sb_model.cu -- fused physical-space tendency + explicit AB3 time step (+ semi-implicit
// adjustment) kernels: the reference's equation-set plugins as built-in CUDA kernels.
//
// Reference: /root/reference/src/testModels.jl:1-215, src/shallowWaterModels.jl:1-298,346-511,
// explicit_timestep src/semiimplicit.jl:672-698, semiimplicit_adjustment :521-597.
// One pass over the point reads the derivative slots it needs and the two history arrays and
// writes var_np1 and expdot_n; the history rotation (:689-695) is a pointer rotation on the host.
#include "sb_internal.hpp"
#include "sb_eqcore.hpp"
#include "sb_thermo.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace sb {

static const char* kEqNames[EQ_COUNT] = {
    "LinearAdvection1D", "LinearAdvectionRZ", "LinearAdvectionRL", "LinearAdvectionRLZ",
    "LinearShallowWater1D", "LinearShallowWaterRL", "Oneway_ShallowWater_Slab", "Twoway_ShallowWater_Slab",
    "Oneway_ShallowWater_HeightResolvedBL", "Euler_test", "BF02_test", "rainfall_test"};

int equation_set_from_name(const char* name) {
  for (int i = 0; i < EQ_COUNT; ++i)
    if (std::strcmp(name, kEqNames[i]) == 0) return i;
  return -1;
}

// Physical slots each equation-set kernel below READS, per variable (bit d = slot d; slot order R {f,r,rr},
// RL {f,r,rr,l,ll}, RZ {f,r,rr,z,zz}, RLZ {f,r,rr,l,ll,z,zz}).  tiles_physics hands this to K3 so that rows nobody reads
// are neither transformed nor written.  0 = the variable is a diagnostic the kernel itself writes (never read).
// Keep in step with the P(v, d) reads of the kernels: tests/test_gpu_parity.py::test_needed_slots_* runs every
// equation set with the unread slots poisoned (SB_K3_POISON) against the all-slots path.
void equation_set_needs(int eq, const EqParams& p, const DevGrid& g, unsigned* need) {
  const unsigned all = (1u << g.D) - 1u;
  for (int v = 0; v < g.V; ++v) need[v] = all;
  auto set = [&](int v, unsigned m) { if (v < g.V) need[v] = m & all; };
  const unsigned F = 1u, R = 2u, RR = 4u, S3 = 8u, S4 = 16u, S5 = 32u;   // S3.. = slots 3, 4, 5 of the geometry
  switch (eq) {
    case EQ_LinearAdvection1D: set(0, F | R | RR); break;
    case EQ_LinearShallowWater1D: set(0, F | R); set(1, F | R | RR); break;
    case EQ_LinearAdvectionRZ:                         // h: r, rr, z, zz ; every other variable: value only
      for (int v = 1; v < g.V; ++v) set(v, F);
      break;
    case EQ_LinearAdvectionRL:
    case EQ_LinearAdvectionRLZ:                        // h: r, l (+ rr, ll for the diffusion term) ; others: value only
      set(0, (eq == EQ_LinearAdvectionRL && !(p.K > 0.0)) ? (F | R | S3) : (F | R | RR | S3 | S4));
      for (int v = 1; v < g.V; ++v) set(v, F);
      break;
    case EQ_LinearShallowWaterRL: set(0, F | R | S3); set(1, F | R | RR | S4); set(2, F | R | RR | S3 | S4); break;
    case EQ_Oneway_ShallowWater_Slab:
    case EQ_Twoway_ShallowWater_Slab:                  // h, ug, vg: r, l ; ub, vb: all five ; w: written, never read
      set(0, F | R | S3); set(1, F | R | S3); set(2, F | R | S3);
      set(3, F | R | RR | S3 | S4); set(4, F | R | RR | S3 | S4); set(5, 0u);
      break;
    case EQ_Oneway_ShallowWater_HeightResolvedBL:      // as the slab + d/dz of ub, vb ; wb: written, never read
      set(0, F | R | S3); set(1, F | R | S3); set(2, F | R | S3);
      set(3, F | R | RR | S3 | S4 | S5); set(4, F | R | RR | S3 | S4 | S5); set(5, 0u);
      break;
    case EQ_Euler_test:                                // RZ: s, mu, u, w all five ; xi: r, z
      set(1, F | R | S3);
      break;
    case EQ_BF02_test:                                 // s, mu, u, w, mu_l all five ; xi, qss: r, z ; var 8 (mu_r): value
      set(1, F | R | S3); set(6, F | R | S3); set(7, F);
      break;
    case EQ_rainfall_test:                             // s, mu, u, w, mu_c, mu_r all five ; xi, qss: r, z
      set(1, F | R | S3); set(7, F | R | S3);
      break;
    default: break;
  }
}


// Variables whose tendency column the equation set never writes: expdot_n[:, v] keeps the zeros it was allocated with
// (LinearAdvectionRZ/RL/RLZ set expdot[:, 1] only, src/testModels.jl:40,68,93), so explicit_timestep adds
// ts/12 (23*0 - 16*0 + 5*0) to them.  The fused K3+K4 kernel uses this to leave those history arrays alone.
unsigned equation_set_passive(int eq, int V) {
  switch (eq) {
    case EQ_LinearAdvectionRZ:
    case EQ_LinearAdvectionRL:
    case EQ_LinearAdvectionRLZ: return V >= 32 ? ~1u : (((1u << V) - 1u) & ~1u);
    // the slab / boundary-layer sets diagnose w (column 6) and step it with a zero tendency (src/shallowWaterModels.jl:66-69,108)
    case EQ_Oneway_ShallowWater_Slab:
    case EQ_Twoway_ShallowWater_Slab:
    case EQ_Oneway_ShallowWater_HeightResolvedBL: return V >= 6 ? (1u << 5) : 0u;
    default: return 0u;
  }
}

struct PointCtx {
  const DevGrid& g;
  const ModelArrays& a;
  long long i;
  __device__ __forceinline__ double P(int v, int d) const { return a.phys[((long long)d * g.V + v) * g.N + i]; }
  __device__ __forceinline__ void setP(int v, int d, double x) const { a.phys[((long long)d * g.V + v) * g.N + i] = x; }
  __device__ __forceinline__ void advance(int v, int t, double ts, double u, double fn) const {
    const long long o = (long long)v * g.N + i;
    if ((a.passive >> v) & 1u) {          // no tendency and an all-zero history (ModelArrays::passive): same arithmetic on zeros,
      a.var_np1[o] = ab_step(t, ts, u, 0.0, 0.0, 0.0);   // the three history passes over this variable are left out
      return;
    }
    double f1 = (t >= 2) ? a.exp_nm1[o] : 0.0;
    double f2 = (t >= 3) ? a.exp_nm2[o] : 0.0;
    a.exp_n[o] = fn;
    a.var_np1[o] = ab_step(t, ts, u, fn, f1, f2);
  }
};

__device__ __forceinline__ double point_radius(const DevGrid& g, long long i) {
  long long h = g.has_z ? i / g.zDim : i;
  return g.rad[g.h2r[h]];
}

// shallow-water / slab boundary layer tendencies shared by Oneway/Twoway (src/shallowWaterModels.jl:60-108,176-228)
__device__ __forceinline__ void slab_tendencies(const PointCtx& c, const EqParams& p, double r, bool twoway, int t) {
  const double g = p.g, K = p.K, Cd = p.Cd, Hfree = p.Hfree, Hb = p.Hb, f = p.f;
  double h = c.P(0, 0), hr = c.P(0, 1), hl = c.P(0, 3);
  double ug = c.P(1, 0), ugr = c.P(1, 1), ugl = c.P(1, 3);
  double vg = c.P(2, 0), vgr = c.P(2, 1), vgl = c.P(2, 3);
  double ub = c.P(3, 0), ubr = c.P(3, 1), ubrr = c.P(3, 2), ubl = c.P(3, 3), ubll = c.P(3, 4);
  double vb = c.P(4, 0), vbr = c.P(4, 1), vbrr = c.P(4, 2), vbl = c.P(4, 3), vbll = c.P(4, 4);
  double U = 0.78 * sqrt((ub * ub) + (vb * vb));
  double w = -Hb * ((ub / r) + ubr + (vbl / r));
  c.setP(5, 0, w);
  double w_ = 0.5 * fabs(w) - w;
  double e0 = ((-vg * hl / r) + (-ug * hr)) + (-(Hfree + h) * ((ug / r) + ugr + (vgl / r)));
  if (twoway) e0 += -(Hfree + h) * w * p.S1;
  double e1 = ((-vg * ugl / r) + (-ug * ugr)) + (-g * hr) + (vg * (f + (vg / r)));
  double e2 = ((-vg * vgl / r) + (-ug * vgr)) + (-g * (hl / r)) + (-ug * (f + (vg / r)));
  double e3 = ((-vb * ubl / r) + (-ub * ubr)) + (-g * hr) + (vb * (f + (vb / r))) + (-(Cd * U * ub / Hb)) +
              (w_ * (ug - ub) / Hb) +
              (K * ((ubr / r) + ubrr - (ub / (r * r)) + (ubll / (r * r)) - (2.0 * vbl / (r * r))));
  double e4 = ((-vb * vbl / r) + (-ub * vbr)) + (-g * (hl / r)) + (-ub * (f + (vb / r))) + (-(Cd * U * vb / Hb)) +
              (w_ * (vg - vb) / Hb) +
              (K * ((vbr / r) + vbrr - (vb / (r * r)) + (vbll / (r * r)) + (2.0 * ubl / (r * r))));
  c.advance(0, t, p.ts, h, e0);
  c.advance(1, t, p.ts, ug, e1);
  c.advance(2, t, p.ts, vg, e2);
  c.advance(3, t, p.ts, ub, e3);
  c.advance(4, t, p.ts, vb, e4);
  c.advance(5, t, p.ts, w, 0.0);
}

template <int EQ>
__global__ void __launch_bounds__(256) k_pointwise(DevGrid g, EqParams p, ModelArrays a, int t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.N) return;
  PointCtx c{g, a, i};
  const double ts = p.ts;
  if (EQ == EQ_LinearAdvection1D) {  // src/testModels.jl:15
    double e = -(p.c_0 * c.P(0, 1)) + (p.K * c.P(0, 2));
    c.advance(0, t, ts, c.P(0, 0), e);
  } else if (EQ == EQ_LinearShallowWater1D) {  // src/shallowWaterModels.jl:253-254
    double e0 = -p.H * c.P(1, 1);
    double e1 = (-p.g * c.P(0, 1)) + (p.K * c.P(1, 2));
    c.advance(0, t, ts, c.P(0, 0), e0);
    c.advance(1, t, ts, c.P(1, 0), e1);
  } else if (EQ == EQ_LinearAdvectionRZ) {  // src/testModels.jl:40 (vars h=1,u=2,w=4)
    double r = point_radius(g, i);
    double hr = c.P(0, 1), hrr = c.P(0, 2), hz = c.P(0, 3), hzz = c.P(0, 4);
    double u = c.P(1, 0), w = c.P(3, 0);
    double e = (-u * hr) + (-w * hz) + (p.K * ((hr / r) + hrr + hzz));
    c.advance(0, t, ts, c.P(0, 0), e);
    for (int v = 1; v < g.V; ++v) c.advance(v, t, ts, c.P(v, 0), 0.0);
  } else if (EQ == EQ_LinearAdvectionRL || EQ == EQ_LinearAdvectionRLZ) {  // src/testModels.jl:62-68, :93
    double r = point_radius(g, i);
    double hr = c.P(0, 1), hl = c.P(0, 3);
    double u = c.P(1, 0), v = c.P(2, 0);
    double e;
    if (EQ == EQ_LinearAdvectionRL && !(p.K > 0.0)) {
      e = (-u * hr) - (v * (hl / r));
    } else {
      double hrr = c.P(0, 2), hll = c.P(0, 4);
      const double ri = 1.0 / r;
      e = advection_rl_tendency(u, v, hr, hl, hrr, hll, ri, ri * ri, p.K);
    }
    c.advance(0, t, ts, c.P(0, 0), e);
    c.advance(1, t, ts, u, 0.0);
    c.advance(2, t, ts, v, 0.0);
    for (int vv = 3; vv < g.V; ++vv) c.advance(vv, t, ts, c.P(vv, 0), 0.0);
  } else if (EQ == EQ_LinearShallowWaterRL) {  // src/shallowWaterModels.jl:290-292
    double r = point_radius(g, i);
    double hr = c.P(0, 1), hl = c.P(0, 3);
    double u = c.P(1, 0), ur = c.P(1, 1), urr = c.P(1, 2), ull = c.P(1, 4);
    double v = c.P(2, 0), vr = c.P(2, 1), vrr = c.P(2, 2), vl = c.P(2, 3), vll = c.P(2, 4);
    double e0 = -p.H * ((u / r) + ur + (vl / r));
    double e1 = (-p.g * hr) + (p.K * ((ur / r) + urr + (ull / (r * r))));
    double e2 = (-p.g * (hl / r)) + (p.K * ((vr / r) + vrr + (vll / (r * r))));
    c.advance(0, t, ts, c.P(0, 0), e0);
    c.advance(1, t, ts, u, e1);
    c.advance(2, t, ts, v, e2);
  } else if (EQ == EQ_Oneway_ShallowWater_Slab) {
    slab_tendencies(c, p, point_radius(g, i), false, t);
  } else if (EQ == EQ_Twoway_ShallowWater_Slab) {
    slab_tendencies(c, p, point_radius(g, i), true, t);
  }
}

// ------------------------------------------------------------------------------------
// column kernels: one thread per level, blockDim = (zDim, columns per block)
// colops (transposed, [k][z'][z]): 0 = CB->CA->CI, 1 = CB->CA->CIx, 2 = CB->CA->CIInt  (of variable "h")
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double col_matvec(const double* __restrict__ Mt, const double* x, int nz, int z) {
  double s = 0.0;
  for (int k = 0; k < nz; ++k) s = fma(Mt[(size_t)k * nz + z], x[k], s);
  return s;
}

// Oneway_ShallowWater_HeightResolvedBL, src/shallowWaterModels.jl:346-511
__global__ void k_heightresolved_bl(DevGrid g, EqParams p, ModelArrays a, int t) {
  SB_DYN_SMEM(double, sm);
  const int nz = g.zDim, z = threadIdx.x, cl = threadIdx.y;
  const long long col = (long long)blockIdx.x * blockDim.y + cl;
  const bool live = col < g.hpoints;
  double* s0 = sm + (size_t)cl * 4 * nz;  // per-column scratch: 4 vectors
  double* s1 = s0 + nz;
  double* s2 = s1 + nz;
  double* s3 = s2 + nz;
  const long long i = live ? col * nz + z : 0;
  PointCtx c{g, a, i};
  const int ring = live ? g.h2r[col] : 0;
  const double r = g.rad[ring];
  const double zz = g.zlev[z];
  const double* Mint = a.colops + (size_t)2 * nz * nz;
  const double* Mdz = a.colops + (size_t)1 * nz * nz;
  double h = c.P(0, 0), hr = c.P(0, 1), hl = c.P(0, 3);
  double ug = c.P(1, 0), ugr = c.P(1, 1), ugl = c.P(1, 3);
  double vg = c.P(2, 0), vgr = c.P(2, 1), vgl = c.P(2, 3);
  double ub = c.P(3, 0), ubr = c.P(3, 1), ubrr = c.P(3, 2), ubl = c.P(3, 3), ubll = c.P(3, 4), ubz = c.P(3, 5);
  double vb = c.P(4, 0), vbr = c.P(4, 1), vbrr = c.P(4, 2), vbl = c.P(4, 3), vbll = c.P(4, 4), vbz = c.P(4, 5);
  double S = sqrt((ubz * ubz) + (vbz * vbz));
  double l = 1.0 / ((1.0 / (0.4 * zz)) + (1.0 / 80.0));
  double Kv = (l * l) * S;
  s0[z] = -((ub / r) + ubr + (vbl / r));
  s1[z] = ub;
  s2[z] = vb;
  __syncthreads();
  double wb = col_matvec(Mint, s0, nz, z);
  // storm-motion surface wind rotated by the column's azimuth; 10 m wind = level 2
  const int jring = (int)(col - g.ring_hoff[ring]);
  const int n = g.ring_n[ring], ri = g.ring_ri[ring];
  const double dl = 6.283185307179586476925286766559 / n;
  const double lam = 0.5 * dl * (ri - 1) + dl * jring;
  const double sfcu = (p.Um * cos(lam)) + (p.Vm * sin(lam));
  const double sfcv = (p.Vm * cos(lam)) - (p.Um * sin(lam));
  const double u10 = s1[1] + sfcu, v10 = s2[1] + sfcv;
  const double U10 = sqrt(u10 * u10 + v10 * v10);
  double Cd = p.Cd;
  if (U10 < 5.2) Cd = 1.0e-3;
  else if (U10 < 33.6) Cd = 4.4e-4 * sqrt(U10);
  __syncthreads();
  s0[z] = (z == 0) ? Cd * U10 * u10 : Kv * ubz;
  s3[z] = (z == 0) ? Cd * U10 * v10 : Kv * vbz;
  __syncthreads();
  double vdu = col_matvec(Mdz, s0, nz, z);
  double vdv = col_matvec(Mdz, s3, nz, z);
  if (!live) return;
  const double gg = p.g, Kh = p.Kh, Hfree = p.Hfree, f = p.f;
  c.setP(5, 0, wb);
  double e0 = ((-vg * hl / r) + (-ug * hr)) + (-(Hfree + h) * ((ug / r) + ugr + (vgl / r)));
  double e1 = ((-vg * ugl / r) + (-ug * ugr)) + (-gg * hr) + (vg * (f + (vg / r)));
  double e2 = ((-vg * vgl / r) + (-ug * vgr)) + (-gg * (hl / r)) + (-ug * (f + (vg / r)));
  double hdu = Kh * ((ubr / r) + ubrr - (ub / (r * r)) + (ubll / (r * r)) - (2.0 * vbl / (r * r)));
  double hdv = Kh * ((vbr / r) + vbrr - (vb / (r * r)) + (vbll / (r * r)) + (2.0 * ubl / (r * r)));
  double e3 = ((-vb * ubl / r) + (-ub * ubr) + (-wb * ubz)) + (-gg * hr) + (vb * (f + (vb / r))) + vdu + hdu;
  double e4 = ((-vb * vbl / r) + (-ub * vbr) + (-wb * vbz)) + (-gg * (hl / r)) + (-ub * (f + (vb / r))) + vdv + hdv;
  c.advance(0, t, p.ts, h, e0);
  c.advance(1, t, p.ts, ug, e1);
  c.advance(2, t, p.ts, vg, e2);
  c.advance(3, t, p.ts, ub, e3);
  c.advance(4, t, p.ts, vb, e4);
  c.advance(5, t, p.ts, wb, 0.0);
}


// ------------------------------------------------------------------------------------
// tensor-core version: the three 64x64 column operators (integral of the divergence, d/dz of the two
// vertical fluxes) of 8 columns at a time are [8 x 64].[64 x 64] products on DMMA (mma.sync.m8n8k4.f64):
// 384 DMMA per 8 columns instead of 3 x 64 x (matrix load + broadcast load + FMA) per point.  block =
// (zDim levels, 8 columns); warp w: level tile nt = w % (zDim/8); products 1 and 2 share the d/dz operator.
// colfrag: [2][zDim/8][zDim/4][32] B fragments (host: build_colop_fragments), read through L1.
#define HB_COLS 8
__global__ void __launch_bounds__(512, 2) k_heightresolved_bl2(DevGrid g, EqParams p, ModelArrays a, int t, long long ngroups) {
  SB_DYN_SMEM(double, sm);
  const int nz = g.zDim, z = threadIdx.x, cl = threadIdx.y;
  const int xs = nz + 4;                       // row stride of X / O (== 4 mod 16: conflict-free fragment loads)
  double* X = sm;                              // [3][HB_COLS][xs]
  double* O = sm + (size_t)3 * HB_COLS * xs;   // [3][HB_COLS][xs]
  double* sfc = O + (size_t)3 * HB_COLS * xs;  // [HB_COLS][2] storm-motion surface wind of the column
  const int tid = cl * nz + z, lane = tid & 31, warp = tid >> 5;
  const int q = lane & 3, li = lane >> 2;
  const int nnt = nz >> 3, nkt = nz >> 2;
  // per-thread constants (the level is fixed for the thread's whole life)
  const double zz = g.zlev[z];
  const double lmix = 1.0 / ((1.0 / (0.4 * zz)) + (1.0 / 80.0));
  const double lmix2 = lmix * lmix;
  const double gg = p.g, Kh = p.Kh, Hfree = p.Hfree, f = p.f;
  auto product = [&](int m, int mat) {   // O[m] = X[m] . colop[mat]   (this warp's level tile)
    const int nt = warp % nnt;
    const double* bf = a.colfrag + ((size_t)(mat * nnt + nt) * nkt) * 32 + lane;
    const double* xa = X + (m * HB_COLS + li) * xs + q;
    double d0 = 0.0, d1 = 0.0;
    for (int kt = 0; kt < nkt; ++kt) sb_dmma(d0, d1, xa[kt * 4], bf[kt * 32]);
    *reinterpret_cast<double2*>(O + (m * HB_COLS + li) * xs + nt * 8 + 2 * q) = make_double2(d0, d1);
  };
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long col = grp * HB_COLS + cl;
    const bool live = col < g.hpoints;
    const long long i = live ? col * nz + z : 0;
    PointCtx c{g, a, i};
    const int ring = live ? g.h2r[col] : 0;
    const double r = g.rad[ring];
    // one reciprocal per point instead of ~25 FP64 divisions (each a ~25-instruction Newton sequence)
    const double ri_ = 1.0 / r, ri2 = ri_ * ri_;
    double h = c.P(0, 0), hr = c.P(0, 1), hl = c.P(0, 3);
    double ug = c.P(1, 0), ugr = c.P(1, 1), ugl = c.P(1, 3);
    double vg = c.P(2, 0), vgr = c.P(2, 1), vgl = c.P(2, 3);
    double ub = c.P(3, 0), ubr = c.P(3, 1), ubrr = c.P(3, 2), ubl = c.P(3, 3), ubll = c.P(3, 4), ubz = c.P(3, 5);
    double vb = c.P(4, 0), vbr = c.P(4, 1), vbrr = c.P(4, 2), vbl = c.P(4, 3), vbll = c.P(4, 4), vbz = c.P(4, 5);
    const double Kv = lmix2 * sqrt((ubz * ubz) + (vbz * vbz));
    X[(0 * HB_COLS + cl) * xs + z] = -((ub * ri_) + ubr + (vbl * ri_));
    X[(1 * HB_COLS + cl) * xs + z] = ub;
    X[(2 * HB_COLS + cl) * xs + z] = vb;
    if (z == 0) {   // storm-motion surface wind rotated by the column's azimuth: once per column
      const int jring = (int)(col - g.ring_hoff[ring]);
      const int n = g.ring_n[ring], rix = g.ring_ri[ring];
      const double dl = 6.283185307179586476925286766559 / n;
      const double lam = 0.5 * dl * (rix - 1) + dl * jring;
      double sl, cl_;
      sincos(lam, &sl, &cl_);
      sfc[2 * cl] = (p.Um * cl_) + (p.Vm * sl);
      sfc[2 * cl + 1] = (p.Vm * cl_) - (p.Um * sl);
    }
    __syncthreads();
    // 10 m wind = level 2 of the column
    const double u10 = X[(1 * HB_COLS + cl) * xs + 1] + sfc[2 * cl], v10 = X[(2 * HB_COLS + cl) * xs + 1] + sfc[2 * cl + 1];
    if (warp < nnt) product(0, 0);               // wb = CIInt(divergence)
    const double U10 = sqrt(u10 * u10 + v10 * v10);
    double Cd = p.Cd;
    if (U10 < 5.2) Cd = 1.0e-3;
    else if (U10 < 33.6) Cd = 4.4e-4 * sqrt(U10);
    __syncthreads();
    X[(1 * HB_COLS + cl) * xs + z] = (z == 0) ? Cd * U10 * u10 : Kv * ubz;
    X[(2 * HB_COLS + cl) * xs + z] = (z == 0) ? Cd * U10 * v10 : Kv * vbz;
    __syncthreads();
    product(1 + warp / nnt, 1);                  // d/dz of the two vertical fluxes (2 * nnt warps == all warps)
    __syncthreads();
    if (live) {
      const double wb = O[(0 * HB_COLS + cl) * xs + z];
      const double vdu = O[(1 * HB_COLS + cl) * xs + z], vdv = O[(2 * HB_COLS + cl) * xs + z];
      c.setP(5, 0, wb);
      double e0 = ((-vg * hl * ri_) + (-ug * hr)) + (-(Hfree + h) * ((ug * ri_) + ugr + (vgl * ri_)));
      double e1 = ((-vg * ugl * ri_) + (-ug * ugr)) + (-gg * hr) + (vg * (f + (vg * ri_)));
      double e2 = ((-vg * vgl * ri_) + (-ug * vgr)) + (-gg * (hl * ri_)) + (-ug * (f + (vg * ri_)));
      double hdu = Kh * ((ubr * ri_) + ubrr - (ub * ri2) + (ubll * ri2) - (2.0 * vbl * ri2));
      double hdv = Kh * ((vbr * ri_) + vbrr - (vb * ri2) + (vbll * ri2) + (2.0 * ubl * ri2));
      double e3 = ((-vb * ubl * ri_) + (-ub * ubr) + (-wb * ubz)) + (-gg * hr) + (vb * (f + (vb * ri_))) + vdu + hdu;
      double e4 = ((-vb * vbl * ri_) + (-ub * vbr) + (-wb * vbz)) + (-gg * (hl * ri_)) + (-ub * (f + (vb * ri_))) + vdv + hdv;
      c.advance(0, t, p.ts, h, e0);
      c.advance(1, t, p.ts, ug, e1);
      c.advance(2, t, p.ts, vg, e2);
      c.advance(3, t, p.ts, ub, e3);
      c.advance(4, t, p.ts, vb, e4);
      c.advance(5, t, p.ts, wb, 0.0);
    }
    // (the next iteration's first writes to X / sfc are ordered behind this iteration's reads by the barriers above:
    //  X[0..2] were last read before the final barrier; O is rewritten only after the next two barriers)
  }
}

// B fragments of a transposed column operator Mt[k][z] (out[z] = sum_k Mt[k][z] x[k]): [zDim/8][zDim/4][32],
// lane = (level n = lane/4 of the tile, k = lane%4 of the k-tile)
void build_colop_fragments(int nz, const double* Mt, std::vector<double>& out) {
  const int nnt = nz / 8, nkt = nz / 4;
  out.assign((size_t)nnt * nkt * 32, 0.0);
  for (int nt = 0; nt < nnt; ++nt)
    for (int kt = 0; kt < nkt; ++kt)
      for (int lane = 0; lane < 32; ++lane)
        out[((size_t)nt * nkt + kt) * 32 + lane] = Mt[(size_t)(kt * 4 + lane % 4) * nz + nt * 8 + lane / 4];
}

// semiimplicit_adjustment (src/semiimplicit.jl:521-597) for one level of a column; every thread of the column calls it
// (two barriers inside).  xi uses impdot[:, xi] = "wdot", w uses impdot[:, w] = "xidot"; o1 / o4 = offsets of this point
// in the xi / w columns of the [V][N] state arrays; s0, s1 = the column's two shared vectors.
// sicols (transposed [k][z'][z]): 0 = F (CB->CA->CI of xi), 1 = Dz (CB->CA->CIx of xi),
//   2,3 = W,X for tau = 0.5 ts ; 4,5 = W,X for tau = 1.25 ts   (W = dct H^-1 Shift, X = dct1 H^-1 Shift)
__device__ __forceinline__ void semi_adjust(const ModelArrays& a, const EqParams& p, int t, int nz, int z, long long o1,
                                            long long o4, double imp1, double imp4, double* s0, double* s1, double& xi_np1,
                                            double& w_np1) {
  const double ts = p.ts;
  double wdot_n = imp1, xidot_n = imp4;
  double wdot_nm1 = (t >= 2) ? a.imp_nm1[o1] : 0.0, wdot_nm2 = (t >= 3) ? a.imp_nm2[o1] : 0.0;
  double xidot_nm1 = (t >= 2) ? a.imp_nm1[o4] : 0.0, xidot_nm2 = (t >= 3) ? a.imp_nm2[o4] : 0.0;
  double tau, w_ns, xi_ns;
  if (t == 1) {
    tau = 0.5 * ts;
    w_ns = w_np1 - (ts * xidot_n) + (ts * 0.5 * xidot_n);
    xi_ns = xi_np1 - (ts * wdot_n) + (ts * 0.5 * wdot_n);
  } else if (t == 2) {
    tau = 1.25 * ts;
    w_ns = w_np1 - (0.5 * ts) * ((3.0 * xidot_n) - xidot_nm1) - (ts * xidot_n) + (ts * 0.75 * xidot_nm1);
    xi_ns = xi_np1 - (0.5 * ts) * ((3.0 * wdot_n) - wdot_nm1) - (ts * wdot_n) + (ts * 0.75 * wdot_nm1);
  } else {
    tau = 1.25 * ts;
    w_ns = w_np1 - ((ts / 12.0) * ((23.0 * xidot_n) - (16.0 * xidot_nm1) + (5.0 * xidot_nm2))) - (ts * xidot_n) +
           (ts * 0.75 * xidot_nm1);
    xi_ns = xi_np1 - ((ts / 12.0) * ((23.0 * wdot_n) - (16.0 * wdot_nm1) + (5.0 * wdot_nm2))) - (ts * wdot_n) +
            (ts * 0.75 * wdot_nm1);
  }
  const size_t nn = (size_t)nz * nz;
  s0[z] = xi_ns;
  __syncthreads();
  double xi_f = col_matvec(a.sicols, s0, nz, z);
  double xi_fz = col_matvec(a.sicols + nn, s0, nz, z);
  s1[z] = tau * p.Pxi_bar * xi_fz - w_ns;   // g before the BC-row shift (folded into W, X)
  __syncthreads();
  const double* Wm = a.sicols + ((t == 1) ? 2 : 4) * nn;
  const double* Xm = Wm + nn;
  w_np1 = col_matvec(Wm, s1, nz, z);
  xi_np1 = xi_f - tau * col_matvec(Xm, s1, nz, z);
}

// Euler_test + semiimplicit_adjustment, src/testModels.jl:100-215, src/semiimplicit.jl:521-597
__global__ void k_euler_test(DevGrid g, EqParams p, ModelArrays a, Thermo th, int semi, int t) {
  SB_DYN_SMEM(double, sm);
  const int nz = g.zDim, z = threadIdx.x, cl = threadIdx.y;
  const long long col = (long long)blockIdx.x * blockDim.y + cl;
  const bool live = col < g.hpoints;
  double* s0 = sm + (size_t)cl * 2 * nz;
  double* s1 = s0 + nz;
  const long long i = live ? col * nz + z : 0;
  PointCtx c{g, a, i};
  const double ts = p.ts, K = p.K;
  const double* rs = a.refstate;  // [profile][deriv][z]
  const double sbar = rs[z], sbar_z = rs[nz + z];
  const double xibar = rs[3 * nz + z], xibar_z = rs[4 * nz + z];
  const double mubar = rs[6 * nz + z], mubar_z = rs[7 * nz + z];
  double s = c.P(0, 0), s_x = c.P(0, 1), s_xx = c.P(0, 2), s_z = c.P(0, 3), s_zz = c.P(0, 4);
  double xi = c.P(1, 0), xi_x = c.P(1, 1), xi_z = c.P(1, 3);
  double mu = c.P(2, 0), mu_x = c.P(2, 1), mu_xx = c.P(2, 2), mu_z = c.P(2, 3), mu_zz = c.P(2, 4);
  double u = c.P(3, 0), u_x = c.P(3, 1), u_xx = c.P(3, 2), u_z = c.P(3, 3), u_zz = c.P(3, 4);
  double w = c.P(4, 0), w_x = c.P(4, 1), w_xx = c.P(4, 2), w_z = c.P(4, 3), w_zz = c.P(4, 4);
  // thermodynamic_tuple
  double q_v = th_ahyp(mu + mubar);
  double rho_d = th.rho_d0 * exp(xi + xibar);
  double Cfac = TH_Cvd + (q_v * TH_Cvv);
  double qfac = (q_v != 0.0) ? pow(rho_d * q_v / th.rho_v0, (q_v * TH_Rv) / Cfac) : 1.0;
  double Tk = TH_T0 * exp(((s + sbar) - (q_v * th.Lv_T0 / TH_T0)) / Cfac) * pow(rho_d / th.rho_d0, TH_Rd / Cfac) * qfac;
  double rho_t = rho_d * (1.0 + q_v);
  double dm = th_dmudq(mu + mubar, q_v);
  double qvp_x = mu_x / dm, qvp_z = mu_z / dm;
  double rhobar = (th.rho_d0 * exp(xibar)) * (1.0 + th_ahyp(mubar));
  double rho_p = rho_t - rhobar;
  double e0 = ((-u * s_x) + (-w * (s_z + sbar_z))) + (K * (s_xx + s_zz));
  double e1 = ((-u * xi_x) + (-w * (xi_z + xibar_z))) - u_x - w_z;
  double e2 = ((-u * mu_x) + (-w * (mu_z + mubar_z))) + (K * (mu_xx + mu_zz));
  double e3 = ((-u * u_x) + (-w * u_z)) + (-(th_pgrad(th, Tk, rho_d, q_v, s_x, xi_x, qvp_x) / rho_t)) + (K * (u_xx + u_zz));
  double e4 = ((-u * w_x) + (-w * w_z)) +
              (-(TH_g * rho_p / rho_t) - (th_pgrad(th, Tk, rho_d, q_v, s_z, xi_z, qvp_z) / rho_t)) + (K * (w_xx + w_zz));
  const long long o1 = (long long)1 * g.N + i, o4 = (long long)4 * g.N + i;
  const double imp1 = -w_z, imp4 = -(p.Pxi_bar * xi_z);
  double f1, f2;
  f1 = (t >= 2) ? a.exp_nm1[o1] : 0.0; f2 = (t >= 3) ? a.exp_nm2[o1] : 0.0;
  double xi_np1 = ab_step(t, ts, xi, e1, f1, f2);
  f1 = (t >= 2) ? a.exp_nm1[o4] : 0.0; f2 = (t >= 3) ? a.exp_nm2[o4] : 0.0;
  double w_np1 = ab_step(t, ts, w, e4, f1, f2);
  if (live) {
    c.advance(0, t, ts, s, e0);
    c.advance(2, t, ts, mu, e2);
    c.advance(3, t, ts, u, e3);
    a.exp_n[o1] = e1;
    a.exp_n[o4] = e4;
    if (a.imp_n) { a.imp_n[o1] = imp1; a.imp_n[o4] = imp4; }
  }
  if (!semi) {
    if (live) { a.var_np1[o1] = xi_np1; a.var_np1[o4] = w_np1; }
    return;
  }
  semi_adjust(a, p, t, nz, z, o1, o4, imp1, imp4, s0, s1, xi_np1, w_np1);
  if (live) {
    a.var_np1[o4] = w_np1;
    a.var_np1[o1] = xi_np1;
  }
}

// lexicographic isless(A, B) of two column vectors in shared memory (Julia: isless(::AbstractVector, ::AbstractVector) =
// cmp(A, B) < 0, first !isequal pair decides): what the un-dotted min / max of condensation_adjustment reduce to
__device__ __forceinline__ bool column_lex_less(const double* A, const double* B, int nz) {
  for (int k = 0; k < nz; ++k)
    if (!th_isequal(A[k], B[k])) return th_isless(A[k], B[k]);
  return false;
}

// BF02_test (RAIN = false, src/testModels.jl:217-385) and rainfall_test (RAIN = true, :387-586): moist compressible
// RZ test sets with bulk condensation (+ warm-rain microphysics), then explicit_timestep (src/semiimplicit.jl:672-698),
// semiimplicit_adjustment (:521-597) and condensation_adjustment (src/microphysics.jl:141-195) on var_np1.
// Variables by column: s, xi, mu, u, w, then BF02: mu_l (named mu_c), qss, mu_r (no tendency) | rainfall: mu_c, mu_r, qss.
// Reference quirks kept as written: the un-dotted vector min / max of condensation_adjustment (lexicographic, whole
// column), sedimentation's clamp that leaves Vt = 0, dmudq(mu_l, q_l) with the perturbation mu_l and the total q_l.
// impdot[:, mu] = q_v and impdot[:, qss] = qss (:354,:364 / :551,:573) are placeholders nothing reads: not stored.
template <bool RAIN>
__global__ void k_moist_test(DevGrid g, EqParams p, ModelArrays a, Thermo th, int semi, int t) {
  SB_DYN_SMEM(double, sm);
  const int nz = g.zDim, z = threadIdx.x, cl = threadIdx.y;
  const long long col = (long long)blockIdx.x * blockDim.y + cl;
  const bool live = col < g.hpoints;
  double* s0 = sm + (size_t)cl * 2 * nz;
  double* s1 = s0 + nz;
  const long long i = live ? col * nz + z : 0;
  PointCtx c{g, a, i};
  const double ts = p.ts, K = p.K;
  const double* rs = a.refstate;  // [profile][deriv][z]
  const double sbar = rs[z], sbar_z = rs[nz + z];
  const double xibar = rs[3 * nz + z], xibar_z = rs[4 * nz + z];
  const double mubar = rs[6 * nz + z], mubar_z = rs[7 * nz + z];
  const int IQ = RAIN ? 7 : 6;                       // column of qss
  double s = c.P(0, 0), s_x = c.P(0, 1), s_xx = c.P(0, 2), s_z = c.P(0, 3), s_zz = c.P(0, 4);
  double xi = c.P(1, 0), xi_x = c.P(1, 1), xi_z = c.P(1, 3);
  double mu = c.P(2, 0), mu_x = c.P(2, 1), mu_xx = c.P(2, 2), mu_z = c.P(2, 3), mu_zz = c.P(2, 4);
  double u = c.P(3, 0), u_x = c.P(3, 1), u_xx = c.P(3, 2), u_z = c.P(3, 3), u_zz = c.P(3, 4);
  double w = c.P(4, 0), w_x = c.P(4, 1), w_xx = c.P(4, 2), w_z = c.P(4, 3), w_zz = c.P(4, 4);
  double m5 = c.P(5, 0), m5_x = c.P(5, 1), m5_xx = c.P(5, 2), m5_z = c.P(5, 3), m5_zz = c.P(5, 4);   // mu_l | mu_c
  double qss = c.P(IQ, 0), qss_x = c.P(IQ, 1), qss_z = c.P(IQ, 3);
  const double mu_total = mu + mubar;
  const ThermoPoint tp = th_tuple(th, s + sbar, xi + xibar, mu_total);
  const double q_v = tp.q_v, rho_d = tp.rho_d, Tk = tp.Tk, pr = tp.p;
  double q_c = 0.0, q_r = 0.0, q_l, rho_t;
  double m6 = 0.0, m6_x = 0.0, m6_xx = 0.0, m6_z = 0.0, m6_zz = 0.0;     // rainfall: mu_r
  if (RAIN) {
    m6 = c.P(6, 0); m6_x = c.P(6, 1); m6_xx = c.P(6, 2); m6_z = c.P(6, 3); m6_zz = c.P(6, 4);
    q_c = th_ahyp(m5);
    q_r = th_ahyp(m6);
    q_l = q_c + q_r;
    rho_t = rho_d * (1.0 + (q_v + q_l));
  } else {
    q_l = th_ahyp(m5 + rs[9 * nz + z]);
    rho_t = rho_d * (1.0 + q_v + q_l);
  }
  const double mu_factor = th_dmudq(mu_total, q_v);
  const double qvp_x = mu_x / mu_factor, qvp_z = mu_z / mu_factor;
  const double rhobar = (th.rho_d0 * exp(xibar)) * (1.0 + th_ahyp(mubar));
  const double rho_p = rho_t - rhobar;
  const double dpdx = th_pgrad(th, Tk, rho_d, q_v, s_x, xi_x, qvp_x);
  const double dpdz = th_pgrad(th, Tk, rho_d, q_v, s_z, xi_z, qvp_z);
  const double Cm = (q_l * TH_Cl) / (TH_Cvd + (q_v * TH_Cvv) + (q_l * TH_Cl));
  const double s_div = Cm * (TH_Rd + q_v * TH_Rv) * (u_x + w_z);
  const double N_c = RAIN ? 100.0 : 500.0, r_c = 10.0;
  const SatPoint sp = th_sat(Tk, pr);
  const double cloudtau = th_invtau(Tk, pr, N_c, r_c);
  double q_cond = qss / (1.0 + th_Q_s(sp, Tk, q_v, q_l));        // q_condensation, src/microphysics.jl:84-93
  q_cond = fmin(q_v, q_cond);
  q_cond = fmax(-q_l, q_cond);
  q_cond = q_cond * cloudtau;
  const double s_cond = th_s_condensation(sp, q_cond, Tk, q_v, q_l, pr);
  const double lift = (u * dpdx) + (w * (dpdz - rhobar * TH_g));
  const double dq = th_dqsdp(sp, pr, rho_d, q_v, q_l);
  double qss_cond, q_evap = 0.0, q_auto = 0.0, q_coll = 0.0, Vt_flux = 0.0;
  if (RAIN) {
    const double rho_r = q_r * rho_d, fice = th_f_ice(Tk);
    double f_vent = 1.6 + 30.39 * pow(rho_r, 0.2046) * pow(fice, 1.5);
    if (f_vent < 0.0) f_vent = 0.0;
    const double rho_vs = sp.e_s / (TH_Rv * Tk);
    double raintau = (f_vent * pow(rho_r, 0.525)) / (1.0e4 * ((2.03 * rho_vs) + (3.337 / Tk)));
    if (raintau < 0.0) raintau = 0.0;
    q_evap = -qss * raintau;
    qss_cond = dq * lift - qss * (cloudtau + raintau);
    q_auto = 0.001 * (q_c - 0.001);
    if (q_auto < 0.0) q_auto = 0.0;
    q_coll = 2.20 * q_c * pow(q_r, 0.875) * fice;
    if (q_coll < 0.0) q_coll = 0.0;
    double Vt = -14.164 * pow(rho_r, 0.1364) * pow(th.rho_d0 / rho_d, 0.5) * fice;
    if (Vt < 0.0) Vt = 0.0;
    s0[z] = q_r * Vt;                                             // col.uMish (:526), CB -> CA -> CIx of the mu_r column
    __syncthreads();
    Vt_flux = col_matvec(a.sicols + (size_t)6 * nz * nz, s0, nz, z) / rho_d;
    __syncthreads();
  } else {
    qss_cond = dq * lift - qss * cloudtau;
  }
  const double e0 = ((-u * s_x) + (-w * (s_z + sbar_z))) + (s_cond + s_div) + (K * (s_xx + s_zz));
  const double e1 = ((-u * xi_x) + (-w * (xi_z + xibar_z))) + (-u_x - w_z);
  const double e2 = ((-u * mu_x) + (-w * (mu_z + mubar_z))) + (RAIN ? (mu_factor * (q_evap - q_cond)) : (-q_cond * mu_factor)) +
                    (K * (mu_xx + mu_zz));
  const double e3 = ((-u * u_x) + (-w * u_z)) + (-dpdx / rho_t) + (K * (u_xx + u_zz));
  const double e4 = ((-u * w_x) + (-w * w_z)) + (((-TH_g * rho_p) - dpdz) / rho_t) + (K * (w_xx + w_zz));
  double e5, e6 = 0.0;
  if (RAIN) {
    e5 = ((-u * m5_x) + (-w * m5_z)) + (th_dmudq(m5, q_c) * (q_cond - q_auto - q_coll)) + (K * (m5_xx + m5_zz));
    e6 = ((-u * m6_x) + (-w * m6_z)) + (th_dmudq(m6, q_r) * (q_auto + q_coll - q_evap - Vt_flux)) + (K * (m6_xx + m6_zz));
  } else {
    e5 = ((-u * m5_x) + (-w * (m5_z + rs[10 * nz + z]))) + (q_cond * th_dmudq(m5, q_l)) + (K * (m5_xx + m5_zz));
  }
  const double eq = ((-u * qss_x) + (-w * qss_z)) + qss_cond;
  auto step = [&](int v, double u0, double fn) {                  // explicit_timestep of one variable; var_np1 stays in a register
    const long long o = (long long)v * g.N + i;
    const double f1 = (t >= 2) ? a.exp_nm1[o] : 0.0;
    const double f2 = (t >= 3) ? a.exp_nm2[o] : 0.0;
    if (live) a.exp_n[o] = fn;
    return ab_step(t, ts, u0, fn, f1, f2);
  };
  double s_np1 = step(0, s, e0), xi_np1 = step(1, xi, e1), mu_np1 = step(2, mu, e2), u_np1 = step(3, u, e3);
  double w_np1 = step(4, w, e4), m5_np1 = step(5, m5, e5), qss_np1 = step(IQ, qss, eq);
  double mur_np1 = RAIN ? step(6, m6, e6) : step(7, c.P(7, 0), 0.0);
  const long long o1 = (long long)1 * g.N + i, o4 = (long long)4 * g.N + i;
  const double imp1 = -w_z, imp4 = -(p.Pxi_bar * xi_z);
  if (live && a.imp_n) { a.imp_n[o1] = imp1; a.imp_n[o4] = imp4; }
  if (semi) semi_adjust(a, p, t, nz, z, o1, o4, imp1, imp4, s0, s1, xi_np1, w_np1);
  // ---- condensation_adjustment (src/microphysics.jl:141-195) on the advanced state; mu_c = column 5, mu_r = column 6 | 7
  {
    const double mu_total2 = mu_np1 + mubar;
    const ThermoPoint t2 = th_tuple(th, s_np1 + sbar, xi_np1 + xibar, mu_total2);
    const double qc2 = th_ahyp(m5_np1), qr2 = th_ahyp(mur_np1), ql2 = qc2 + qr2;
    const SatPoint sp2 = th_sat(t2.Tk, t2.p);
    const double Qs2 = th_Q_s(sp2, t2.Tk, t2.q_v, ql2);
    double qcond2 = (t2.q_v - sp2.q_sat - qss_np1) / (1.0 + Qs2);
    __syncthreads();
    s0[z] = qcond2; s1[z] = t2.q_v;
    __syncthreads();
    if (!column_lex_less(s0, s1, nz)) qcond2 = t2.q_v;            // min(q_v, q_cond) on the column vectors (:185)
    __syncthreads();
    s0[z] = qcond2; s1[z] = -qc2;
    __syncthreads();
    if (column_lex_less(s0, s1, nz)) qcond2 = -qc2;               // max(-q_c, q_cond) (:187)
    const double tau_r = 0.25;
    mu_np1 = mu_np1 - tau_r * th_dmudq(mu_total2, t2.q_v) * qcond2;
    m5_np1 = m5_np1 + tau_r * th_dmudq(m5_np1, qc2) * qcond2;
    s_np1 = s_np1 + tau_r * th_s_condensation(sp2, qcond2, t2.Tk, t2.q_v, ql2, t2.p);
  }
  if (live) {
    const long long N = g.N;
    a.var_np1[i] = s_np1; a.var_np1[N + i] = xi_np1; a.var_np1[2 * N + i] = mu_np1; a.var_np1[3 * N + i] = u_np1;
    a.var_np1[4 * N + i] = w_np1; a.var_np1[5 * N + i] = m5_np1; a.var_np1[(long long)IQ * N + i] = qss_np1;
    a.var_np1[(long long)(RAIN ? 6 : 7) * N + i] = mur_np1;
  }
}

template <int EQ>
static void launch_pw(const LaunchCtx& c, const DevGrid& g, const EqParams& p, const ModelArrays& a, int t) {
  long long blocks = (g.N + 255) / 256;
  SB_LAUNCH(k_pointwise<EQ>, dim3((unsigned)blocks), dim3(256), 0, c.stream, g, p, a, t);
}

void launch_equation_set(const LaunchCtx& c, int eq, const DevGrid& g, const EqParams& p, const ModelArrays& a,
                         int tstep) {
  ProfScope prof_scope_(c, "equation_set");
  switch (eq) {
    case EQ_LinearAdvection1D: launch_pw<EQ_LinearAdvection1D>(c, g, p, a, tstep); break;
    case EQ_LinearAdvectionRZ: launch_pw<EQ_LinearAdvectionRZ>(c, g, p, a, tstep); break;
    case EQ_LinearAdvectionRL: launch_pw<EQ_LinearAdvectionRL>(c, g, p, a, tstep); break;
    case EQ_LinearAdvectionRLZ: launch_pw<EQ_LinearAdvectionRLZ>(c, g, p, a, tstep); break;
    case EQ_LinearShallowWater1D: launch_pw<EQ_LinearShallowWater1D>(c, g, p, a, tstep); break;
    case EQ_LinearShallowWaterRL: launch_pw<EQ_LinearShallowWaterRL>(c, g, p, a, tstep); break;
    case EQ_Oneway_ShallowWater_Slab: launch_pw<EQ_Oneway_ShallowWater_Slab>(c, g, p, a, tstep); break;
    case EQ_Twoway_ShallowWater_Slab: launch_pw<EQ_Twoway_ShallowWater_Slab>(c, g, p, a, tstep); break;
    case EQ_Oneway_ShallowWater_HeightResolvedBL: {
      static const bool v1 = std::getenv("SB_TCBL_V1") != nullptr;   // A/B switch
      if (a.colfrag && !v1 && g.zDim % 8 == 0 && g.zDim <= 64) {
        long long ngroups = (g.hpoints + HB_COLS - 1) / HB_COLS;
        long long blocks = ngroups < sb_sm_count() * 8 ? ngroups : sb_sm_count() * 8;
        size_t smem = ((size_t)6 * HB_COLS * (g.zDim + 4) + 2 * HB_COLS) * sizeof(double);
        SB_LAUNCH(k_heightresolved_bl2, dim3((unsigned)blocks), dim3(g.zDim, HB_COLS), smem, c.stream, g, p, a, tstep, ngroups);
        break;
      }
      int cpb = 128 / g.zDim; if (cpb < 1) cpb = 1;
      long long blocks = (g.hpoints + cpb - 1) / cpb;
      size_t smem = (size_t)cpb * 4 * g.zDim * sizeof(double);
      SB_LAUNCH(k_heightresolved_bl, dim3((unsigned)blocks), dim3(g.zDim, cpb), smem, c.stream, g, p, a, tstep);
      break;
    }
    case EQ_Euler_test: {
      int cpb = 128 / g.zDim; if (cpb < 1) cpb = 1;
      long long blocks = (g.hpoints + cpb - 1) / cpb;
      size_t smem = (size_t)cpb * 2 * g.zDim * sizeof(double);
      SB_LAUNCH(k_euler_test, dim3((unsigned)blocks), dim3(g.zDim, cpb), smem, c.stream, g, p, a, make_thermo(),
                a.imp_n ? 1 : 0, tstep);
      break;
    }
    case EQ_BF02_test:
    case EQ_rainfall_test: {
      int cpb = 128 / g.zDim; if (cpb < 1) cpb = 1;
      long long blocks = (g.hpoints + cpb - 1) / cpb;
      size_t smem = (size_t)cpb * 2 * g.zDim * sizeof(double);
      if (eq == EQ_rainfall_test)
        SB_LAUNCH(k_moist_test<true>, dim3((unsigned)blocks), dim3(g.zDim, cpb), smem, c.stream, g, p, a, make_thermo(),
                  a.imp_n ? 1 : 0, tstep);
      else
        SB_LAUNCH(k_moist_test<false>, dim3((unsigned)blocks), dim3(g.zDim, cpb), smem, c.stream, g, p, a, make_thermo(),
                  a.imp_n ? 1 : 0, tstep);
      break;
    }
    default: throw std::runtime_error("equation set has no CUDA kernel");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("equation-set launch: ") + cudaGetErrorString(e));
  if (c.launches) ++*c.launches;
}

}  // namespace sb
