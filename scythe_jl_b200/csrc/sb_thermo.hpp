// sb_thermo.hpp -- moist thermodynamic closure and bulk microphysics shared by the RZ test equation sets
// (Euler_test in sb_model.cu; BF02_test / rainfall_test in sb_moist.cu).
// Reference: /root/reference/src/thermodynamics.jl:2-17,31-32,67-80,96-178,184-269; /root/reference/src/microphysics.jl:81-264.
#pragma once
#include "cuda_emu.h"

#include <cmath>

namespace sb {

#define TH_Rd 287.04
#define TH_Rv 461.50
#define TH_Cvd 716.96
#define TH_Cvv 1410.0
#define TH_Cl 4186.0
#define TH_g 9.81
#define TH_Lv0 2.501e6
#define TH_T0 273.16
#define TH_p0 1000.0
#define TH_q0 1.0e-7

struct Thermo { double rho_d0, rho_v0, Lv_T0; };
static inline Thermo make_thermo() {
  Thermo th;
  th.rho_d0 = 100.0 * TH_p0 / (TH_T0 * TH_Rd);
  double Tc = TH_T0 - 273.15;
  double es = 6.112 * std::exp(17.67 * Tc / (Tc + 243.5));
  th.rho_v0 = 100.0 * es / (TH_T0 * TH_Rv);
  th.Lv_T0 = TH_Lv0 + (((TH_Cvv + TH_Rv) - TH_Cl) * (TH_T0 - TH_T0));
  return th;
}
__device__ __forceinline__ double th_ahyp(double mu) {
  return (mu < 0.0) ? 0.0 : sqrt(mu * mu + TH_q0 * TH_q0) + mu - TH_q0;
}
__device__ __forceinline__ double th_dmudq(double mu, double q_v) { return ((q_v + TH_q0) - mu) / (q_v + TH_q0); }
__device__ __forceinline__ double th_P_s(double Tk, double rho_d, double q_v) {
  return Tk * ((rho_d * TH_Rd) + (q_v * rho_d * TH_Rv)) / (TH_Cvd + (q_v * TH_Cvv));
}
__device__ __forceinline__ double th_pgrad(const Thermo& th, double Tk, double rho_d, double q_v, double s_x,
                                           double xi_x, double qv_x) {
  double Ps = th_P_s(Tk, rho_d, q_v);
  double Pxi = (TH_Rd + (q_v * rho_d * TH_Rv)) * ((rho_d * Tk) + Ps);
  double Pqv = 0.0;
  if (q_v != 0.0) {
    double rho_v = q_v * rho_d;
    double qf = TH_Rv * (1 + log(rho_v / th.rho_v0)) - (TH_Cvv * log(Tk / TH_T0)) - th.Lv_T0 / TH_T0;
    Pqv = (rho_d * TH_Rv * Tk) + qf * Ps;
  }
  return (Ps * s_x) + (Pxi * xi_x) + (Pqv * qv_x);
}


// ---- moist additions (BF02_test / rainfall_test)
#define TH_Eps (TH_Rd / TH_Rv)
#define TH_Cpd (TH_Cvd + TH_Rd)
#define TH_Cpv (TH_Cvv + TH_Rv)

struct ThermoPoint { double q_v, rho_d, Tk, p; };
// thermodynamic_tuple (src/thermodynamics.jl:248-257)
__device__ __forceinline__ ThermoPoint th_tuple(const Thermo& th, double s, double xi, double mu) {
  ThermoPoint t;
  t.q_v = th_ahyp(mu);
  t.rho_d = th.rho_d0 * exp(xi);
  const double Cfac = TH_Cvd + (t.q_v * TH_Cvv);
  const double qfac = (t.q_v != 0.0) ? pow(t.rho_d * t.q_v / th.rho_v0, (t.q_v * TH_Rv) / Cfac) : 1.0;
  t.Tk = TH_T0 * exp((s - (t.q_v * th.Lv_T0 / TH_T0)) / Cfac) * pow(t.rho_d / th.rho_d0, TH_Rd / Cfac) * qfac;
  const double pd = 0.01 * TH_Rd * t.Tk * t.rho_d;
  const double e = 0.01 * TH_Rv * t.Tk * t.rho_d * t.q_v;
  t.p = pd + e;
  return t;
}
__device__ __forceinline__ double th_L_v(double Tk) { return TH_Lv0 + ((TH_Cpv - TH_Cl) * (Tk - TH_T0)); }
__device__ __forceinline__ double th_vapor_pressure(double p, double q_v) { return (p * q_v) / (TH_Eps + q_v); }
// Buck (1981) saturation vapour pressure over liquid and its temperature derivative (src/thermodynamics.jl:108-150)
__device__ __forceinline__ void th_buck(double Tk, double p, double& es, double& des_dT) {
  const double Tc = Tk - 273.15;
  const double fw4 = 1.0 + 7.2e-4 + (p * (3.20e-6 + (5.9e-10 * (Tc * Tc))));
  const double d_fw4 = 2.0 * p * 5.9e-10 * Tc;
  const double b = 18.729, c = 257.87, d = 227.3;
  const double ew4 = 6.1121 * exp((b - (Tc / d)) * Tc / (Tc + c));
  const double T1 = (d * b - (2.0 * Tc)) * (d * (Tc + c)) - d * ((d * b * Tc) - (Tc * Tc));
  const double T2 = (d * (Tc + c)) * (d * (Tc + c));
  const double d_ew4 = ew4 * T1 / T2;
  es = fw4 * ew4;
  des_dT = ew4 * d_fw4 + fw4 * d_ew4;
}
struct SatPoint { double e_s, q_sat, dqsdT; };
__device__ __forceinline__ SatPoint th_sat(double Tk, double p) {
  SatPoint s;
  double des;
  th_buck(Tk, p, s.e_s, des);
  s.q_sat = TH_Eps * s.e_s / (p - s.e_s);                       // q_sat_liquid :171-178
  s.dqsdT = des * TH_Eps * p / ((p - s.e_s) * (p - s.e_s));
  return s;
}
__device__ __forceinline__ double th_Q_s(const SatPoint& s, double Tk, double q_v, double q_l) {   // microphysics.jl:108-114
  return th_L_v(Tk) * s.dqsdT / (TH_Cpd + (q_v * TH_Cpv) + (q_l * TH_Cl));
}
__device__ __forceinline__ double th_dqsdp(const SatPoint& s, double p, double rho_d, double q_v, double q_l) {   // :116-123
  return s.q_sat / (100.0 * (p - s.e_s)) - (s.dqsdT / (rho_d * (TH_Cpd + (q_v * TH_Cpv) + (q_l * TH_Cl))));
}
__device__ __forceinline__ double th_invtau(double Tk, double p, double N_c, double r_c) {   // :125-139
  const double Dv = 0.211 * pow(Tk / 273.15, 1.94) * (1013.25 / p);
  return 4 * 3.141592653589793 * Dv * N_c * (r_c * 1.0e-4);
}
__device__ __forceinline__ double th_s_condensation(const SatPoint& s, double q_cond, double Tk, double q_v, double q_l, double p) {   // :96-105
  const double Cm = (q_l * TH_Cl) / (TH_Cvd + (q_v * TH_Cvv) + (q_l * TH_Cl));
  const double e = th_vapor_pressure(p, q_v);
  return q_cond * (((-th_L_v(Tk) * Cm) / Tk) - (TH_Cl * log(Tk / TH_T0)) + (TH_Rv * log(e / s.e_s)));
}
__device__ __forceinline__ double th_f_ice(double Tk) { return Tk < 273.15 ? 0.2 + 0.8 / cosh((273.15 - Tk) / 5.0) : 1.0; }   // :217-225
// Julia isless / isequal for Float64 (NaN above everything, -0.0 < 0.0): the vector min/max of condensation_adjustment
__device__ __forceinline__ bool th_neg(double a) { return copysign(1.0, a) < 0.0; }
__device__ __forceinline__ bool th_isless(double a, double b) {
  if (a != a) return false;
  if (b != b) return true;
  return a < b || (a == b && th_neg(a) && !th_neg(b));
}
__device__ __forceinline__ bool th_isequal(double a, double b) {
  return (a == b && th_neg(a) == th_neg(b)) || (a != a && b != b);
}

}  // namespace sb
