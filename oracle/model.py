"""Oracle: Scythe.jl driver, time stepping and the in-scope equation sets.

TEST INFRASTRUCTURE ONLY.  These parts DO live in the reference, restated here with
citations (paths relative to /root/reference):

* ``ModelParameters``                         -- src/Scythe.jl:8-21
* ``ModelTile`` / ``createModelTile``           -- src/semiimplicit.jl:18-124
* ``initialize_model`` / ``run_model`` / ``model_loop`` / ``advanceTimestep`` -- src/semiimplicit.jl:126-332
* ``explicit_timestep`` (Euler -> AB2 -> AB3) -- src/semiimplicit.jl:672-698
* ``semiimplicit_adjustment`` + Helmholtz     -- src/semiimplicit.jl:521-597, 768-781
* ``calcTendency`` / ``checkCFL``               -- src/semiimplicit.jl:728-751
* equation sets                               -- src/testModels.jl:1-215, src/shallowWaterModels.jl:1-298,346-511
* thermodynamic closure for Euler_test        -- src/thermodynamics.jl:2-17,31-32,67-80,184-269
* reference state                             -- src/reference_state.jl:4-10,138-199
* moist test sets BF02_test / rainfall_test   -- src/testModels.jl:217-385, 387-586
* bulk microphysics + condensation_adjustment -- src/microphysics.jl:81-264, src/thermodynamics.jl:96-167

The Distributed/SharedArray/RemoteChannel plumbing is replaced by an in-process loop over
tiles that performs the same assignments in the same order (own block assigned, halo added).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import chebyshev as cheb
from . import grids as G


@dataclass
class ModelParameters:  # src/Scythe.jl:8-21
    ts: float = 0.0
    integration_time: float = 1.0
    output_interval: float = 1.0
    equation_set: str = "LinearAdvection1D"
    initial_conditions: str = "ic.csv"
    output_dir: str = "./output/"
    ref_state_file: str = ""
    grid_params: G.GridParameters = None
    physical_params: dict = field(default_factory=dict)
    options: dict = field(default_factory=lambda: {"semiimplicit": False, "exact_reference_state": False})


# ------------------------------------------------------------------ thermodynamics subset
Rd = 287.04
Rv = 461.50
Cvd = 716.96
Cvv = 1410.0
Cpv = Cvv + Rv
Cl = 4186.0
gravity = 9.81
L_v0 = 2.501e6
T_0 = 273.16
p_0 = 1000.0
q0 = 1.0e-7


def sat_pressure_liquid(Tk):
    Tc = Tk - 273.15
    return 6.112 * np.exp(17.67 * Tc / (Tc + 243.5))


rho_d0 = 100.0 * p_0 / (T_0 * Rd)
rho_v0 = 100.0 * float(sat_pressure_liquid(T_0)) / (T_0 * Rv)


def L_v(Tk):
    return L_v0 + ((Cpv - Cl) * (Tk - T_0))


def ahyp(mu):
    mu = np.asarray(mu, dtype=np.float64)
    return np.where(mu < 0.0, 0.0, np.sqrt(mu * mu + q0 * q0) + mu - q0)


def dmudq(mu, q_v):
    return ((q_v + q0) - mu) / (q_v + q0)


def dry_density(xi):
    return rho_d0 * np.exp(xi)


def temperature(s, rho_d, q_v):
    Cfactor = Cvd + (q_v * Cvv)
    safe = np.where(q_v != 0.0, rho_d * q_v / rho_v0, 1.0)
    qfactor = np.where(q_v != 0.0, safe ** ((q_v * Rv) / Cfactor), 1.0)
    rhofactor = (rho_d / rho_d0) ** (Rd / Cfactor)
    Tfactor = np.exp((s - (q_v * L_v(T_0) / T_0)) / Cfactor)
    return T_0 * Tfactor * rhofactor * qfactor


def thermodynamic_tuple(s, xi, mu):
    q_v = ahyp(mu)
    rho_d = dry_density(xi)
    Tk = temperature(s, rho_d, q_v)
    pd = 0.01 * Rd * Tk * rho_d
    e = 0.01 * Rv * Tk * rho_d * q_v
    return q_v, rho_d, Tk, pd + e


def P_s(Tk, rho_d, q_v):
    Cfactor = Cvd + (q_v * Cvv)
    return Tk * ((rho_d * Rd) + (q_v * rho_d * Rv)) / Cfactor


def P_xi(Tk, rho_d, q_v):
    return (Rd + (q_v * rho_d * Rv)) * ((rho_d * Tk) + P_s(Tk, rho_d, q_v))


def P_qv(Tk, rho_d, q_v):
    rho_v = np.where(q_v != 0.0, q_v * rho_d, rho_v0)
    qfactor = Rv * (1 + np.log(rho_v / rho_v0)) - (Cvv * np.log(Tk / T_0)) - L_v(T_0) / T_0
    qfactor = qfactor * P_s(Tk, rho_d, q_v)
    return np.where(q_v != 0.0, (rho_d * Rv * Tk) + qfactor, 0.0)


def pressure_gradient(Tk, rho_d, q_v, s_x, xi_x, qv_x):
    return (P_s(Tk, rho_d, q_v) * s_x) + (P_xi(Tk, rho_d, q_v) * xi_x) + (P_qv(Tk, rho_d, q_v) * qv_x)


# ---- moist thermodynamics / bulk microphysics (src/thermodynamics.jl:96-167, src/microphysics.jl:81-264)
Eps = Rd / Rv
Cpd = Cvd + Rd


def entropy(Tk, rho_d, q_v):  # src/thermodynamics.jl:44-54 (q_v > 0)
    qfactor = q_v * (Rv * np.log(q_v * rho_d / rho_v0) - (L_v(T_0) / T_0))
    return ((Cvd + (q_v * Cvv)) * np.log(Tk / T_0)) - (Rd * np.log(rho_d / rho_d0)) - qfactor


def bhyp(q_v):  # src/thermodynamics.jl:196-200
    return 0.5 * ((q_v + q0) - (q0 * q0 / (q_v + q0)))


def vapor_pressure(p, q_v):  # src/thermodynamics.jl:96-101
    return (p * q_v) / (Eps + q_v)


def _buck(Tk, phPa):  # shared pieces of src/thermodynamics.jl:108-150
    Tc = Tk - 273.15
    fw4 = 1.0 + 7.2e-4 + (phPa * (3.20e-6 + (5.9e-10 * (Tc * Tc))))
    ew4 = 6.1121 * np.exp((18.729 - (Tc / 227.3)) * Tc / (Tc + 257.87))
    return Tc, fw4, ew4


def sat_pressure_liquid_buck(Tk, phPa):  # src/thermodynamics.jl:108-125
    _, fw4, ew4 = _buck(Tk, phPa)
    return fw4 * ew4


def sat_pressure_liquid_buck_dT(Tk, phPa):  # src/thermodynamics.jl:127-150
    Tc, fw4, ew4 = _buck(Tk, phPa)
    d_fw4 = 2.0 * phPa * 5.9e-10 * Tc
    b, c, d = 18.729, 257.87, 227.3
    T1 = (d * b - (2.0 * Tc)) * (d * (Tc + c)) - d * ((d * b * Tc) - (Tc * Tc))
    T2 = (d * (Tc + c)) * (d * (Tc + c))
    d_ew4 = ew4 * T1 / T2
    return ew4 * d_fw4 + fw4 * d_ew4


def q_sat_liquid(Tk, phPa):  # src/thermodynamics.jl:171-178
    ew = sat_pressure_liquid_buck(Tk, phPa)
    return Eps * ew / (phPa - ew)


def Q_s_factor(Tk, p, q_v, q_l):  # src/microphysics.jl:108-114
    e_s = sat_pressure_liquid_buck(Tk, p)
    dqsdT = sat_pressure_liquid_buck_dT(Tk, p) * Eps * p / ((p - e_s) * (p - e_s))
    return L_v(Tk) * dqsdT / (Cpd + (q_v * Cpv) + (q_l * Cl))


def dqsdp(Tk, p, rho_d, q_v, q_l):  # src/microphysics.jl:116-123
    q_sat = q_sat_liquid(Tk, p)
    e_s = sat_pressure_liquid_buck(Tk, p)
    dqsdT = sat_pressure_liquid_buck_dT(Tk, p) * Eps * p / ((p - e_s) * (p - e_s))
    return q_sat / (100.0 * (p - e_s)) - (dqsdT / (rho_d * (Cpd + (q_v * Cpv) + (q_l * Cl))))


def vapor_diffusity(Tk, p):  # src/microphysics.jl:133-139
    return 0.211 * (Tk / 273.15) ** 1.94 * (1013.25 / p)


def invtau_condensation(Tk, p, N_c, r_c):  # src/microphysics.jl:125-131
    return 4 * math.pi * vapor_diffusity(Tk, p) * N_c * (r_c * 1.0e-4)


def q_condensation(qss, Tk, p, q_v, q_l, N_c, r_c):  # src/microphysics.jl:84-93 (scalar min / max)
    q_cond = qss / (1.0 + Q_s_factor(Tk, p, q_v, q_l))
    q_cond = np.minimum(q_v, q_cond)
    q_cond = np.maximum(-q_l, q_cond)
    return q_cond * invtau_condensation(Tk, p, N_c, r_c)


def s_condensation(q_cond, Tk, rho_d, q_v, q_l, p):  # src/microphysics.jl:96-105
    Cm = (q_l * Cl) / (Cvd + (q_v * Cvv) + (q_l * Cl))
    e = vapor_pressure(p, q_v)
    sat_e = sat_pressure_liquid_buck(Tk, p)
    return q_cond * (((-L_v(Tk) * Cm) / Tk) - (Cl * np.log(Tk / T_0)) + (Rv * np.log(e / sat_e)))


def autoconversion(q_c, rho_d):  # src/microphysics.jl:197-205
    return np.maximum(0.001 * (q_c - 0.001), 0.0)


def f_ice(Tk):  # src/microphysics.jl:217-225
    return np.where(Tk < 273.15, 0.2 + 0.8 / np.cosh((273.15 - Tk) / 5.0), 1.0)


def collection(q_c, q_r, rho_d, Tk):  # src/microphysics.jl:207-215
    return np.maximum(2.20 * q_c * q_r ** 0.875 * f_ice(Tk), 0.0)


def f_ventilation(q_r, rho_d, Tk):  # src/microphysics.jl:241-250
    rho_r = q_r * rho_d
    return np.maximum(1.6 + 30.39 * rho_r ** 0.2046 * f_ice(Tk) ** 1.5, 0.0)


def rain_evaporation(q_r, rho_d, Tk, p):  # src/microphysics.jl:227-239
    e_s = sat_pressure_liquid_buck(Tk, p)
    rho_vs = e_s / (Rv * Tk)
    rho_r = q_r * rho_d
    q_evap = (f_ventilation(q_r, rho_d, Tk) * rho_r ** 0.525) / (1.0e4 * ((2.03 * rho_vs) + (3.337 / Tk)))
    return np.maximum(q_evap, 0.0)


def sedimentation(q_r, rho_d, Tk):  # src/microphysics.jl:252-264 -- the clamp `Vt < 0 -> 0` leaves Vt = 0 (or -0.0) always
    rho_r = q_r * rho_d
    Vt = -14.164 * rho_r ** 0.1364 * (rho_d0 / rho_d) ** 0.5 * f_ice(Tk)
    return np.where(Vt < 0.0, 0.0, Vt)


def _isless(a, b):
    """Julia isless for Float64 (NaN is larger than everything, -0.0 < 0.0)."""
    return np.where(np.isnan(a), False, np.where(np.isnan(b), True,
                    (a < b) | ((a == b) & np.signbit(a) & ~np.signbit(b))))


def _lex_less(A, B):
    """Julia `isless(A::Vector, B::Vector)` per column: A, B are [ncols, nz]; lexicographic, first unequal element decides."""
    neq = ~(((A == B) & (np.signbit(A) == np.signbit(B))) | (np.isnan(A) & np.isnan(B)))       # !isequal
    first = np.argmax(neq, axis=1)
    rows = np.arange(A.shape[0])
    return neq.any(axis=1) & _isless(A[rows, first], B[rows, first])


def condensation_adjustment(mtile, t):
    """src/microphysics.jl:141-195.  `min(q_v, q_cond)` / `max(-q_c, q_cond)` (:185-187) are applied there to the column
    VECTORS without a dot: Julia's generic min/max fall back to `isless`, which is lexicographic for vectors, so each
    call returns one of its two arguments WHOLE, per column.  Restated as written."""
    v = mtile.model.grid_params.vars
    try:
        si, xii, mui, mci, mri, qi = (v[k] - 1 for k in ("s", "xi", "mu", "mu_c", "mu_r", "qss"))
    except KeyError as e:
        raise KeyError(f"key {e.args[0]!r} not found") from None
    nz = mtile.model.grid_params.zDim
    ncols = mtile.tile.N // nz
    ref = mtile.ref_state
    rep = lambda a: np.tile(a, ncols)  # noqa: E731
    np1 = mtile.var_np1
    s, xi, mu, mu_c, mu_r, qss = (np1[:, i] for i in (si, xii, mui, mci, mri, qi))
    mu_total = mu + rep(ref.mubar[:, 0])
    q_v, rho_d, Tk, p = thermodynamic_tuple(s + rep(ref.sbar[:, 0]), xi + rep(ref.xibar[:, 0]), mu_total)
    q_c = ahyp(mu_c)
    q_r = ahyp(mu_r)
    q_l = q_c + q_r
    q_sat = q_sat_liquid(Tk, p)
    Q_s = Q_s_factor(Tk, p, q_v, q_l)
    tau_r = 0.25
    q_cond = (q_v - q_sat - qss) / (1.0 + Q_s)
    col = lambda a: a.reshape(ncols, nz)  # noqa: E731
    qc2, qv2, nqc2 = col(q_cond), col(q_v), col(-q_c)
    qc2 = np.where(_lex_less(qc2, qv2)[:, None], qc2, qv2)          # min(q_v, q_cond) = isless(q_cond, q_v) ? q_cond : q_v
    qc2 = np.where(_lex_less(qc2, nqc2)[:, None], nqc2, qc2)        # max(-q_c, q_cond) = isless(q_cond, -q_c) ? -q_c : q_cond
    q_cond = qc2.reshape(-1)
    np1[:, mui] = mu - tau_r * dmudq(mu_total, q_v) * q_cond
    np1[:, mci] = mu_c + tau_r * dmudq(mu_c, q_c) * q_cond
    np1[:, si] = s + tau_r * s_condensation(q_cond, Tk, rho_d, q_v, q_l, p)


@dataclass
class ReferenceState:  # src/reference_state.jl:4-10
    sbar: np.ndarray = None
    xibar: np.ndarray = None
    mubar: np.ndarray = None
    mu_lbar: np.ndarray = None
    Pxi_bar: float = 0.0


def transform_reference_state(gp: G.GridParameters, prof: np.ndarray) -> np.ndarray:
    """src/reference_state.jl:138-157 -- filtered value, d/dz, d2/dz2 without BCs."""
    col = cheb.Chebyshev1D(cheb.ChebyshevParameters(gp.zmin, gp.zmax, gp.zDim, gp.b_zDim))
    a = col.CAtransform(col.CBtransform(prof))
    return np.stack([col.CItransform(a), col.CIxtransform(a), col.CIxxtransform(a)], axis=1)


def exact_reference_state_from_profiles(gp, sbar, xibar, mubar, mu_lbar) -> ReferenceState:
    """src/reference_state.jl:159-199 with the file already parsed into per-level profiles."""
    s3 = transform_reference_state(gp, np.asarray(sbar, float))
    x3 = transform_reference_state(gp, np.asarray(xibar, float))
    m3 = transform_reference_state(gp, np.asarray(mubar, float))
    l3 = transform_reference_state(gp, np.asarray(mu_lbar, float))
    q_v, rho_d, Tk, _ = thermodynamic_tuple(s3[:, 0], x3[:, 0], m3[:, 0])
    Pxi = P_xi(Tk, rho_d, q_v)
    Pxi_bar = float(np.mean(Pxi / (dry_density(x3[:, 0]) * (1.0 + ahyp(m3[:, 0])))))
    return ReferenceState(s3, x3, m3, l3, Pxi_bar)


def interpolate_reference_file(gp, text: str, z: np.ndarray) -> ReferenceState:
    """src/reference_state.jl:17-136 with the sounding file given as text (first line: surface pressure [hPa], theta [K],
    q_v [g/kg]; then altitude [m], theta, q_v): interpolation to the model levels, hydrostatic integration, re-integration
    through the Chebyshev column, entropy variables and their vertical derivatives."""
    lines = text.split("\n")
    first = lines[0].split()
    sfc_pressure = float(first[0])
    alt, theta_in, q_in = [0.0], [float(first[1])], [float(first[2])]
    for ln in lines[1:]:
        if not ln.strip():
            break
        a, th, q = ln.split()[:3]
        alt.append(float(a)); theta_in.append(float(th)); q_in.append(float(q))
    n = len(z)
    theta, q_v = np.zeros(n), np.zeros(n)
    theta[0], q_v[0] = theta_in[0], q_in[0]
    for i in range(1, n):
        found = False
        for j in range(1, len(alt)):
            if alt[j - 1] < z[i] and alt[j] > z[i]:
                theta[i] = theta_in[j - 1] + (z[i] - alt[j - 1]) * (theta_in[j] - theta_in[j - 1]) / (alt[j] - alt[j - 1])
                q_v[i] = q_in[j - 1] + (z[i] - alt[j - 1]) * (q_in[j] - q_in[j - 1]) / (alt[j] - alt[j - 1])
                found = True
            elif alt[j] == z[i]:
                theta[i], q_v[i] = theta_in[j], q_in[j]
                found = True
        if not found:
            raise ValueError(f"DomainError({i + 1}): Can't find an interpolating level for reference state")
    q_v = q_v * 1.0e-3
    Tk, p, rho_d, rho_t = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    p[0] = sfc_pressure
    Tk[0] = theta[0] / (p_0 / p[0]) ** (Rd / Cpd)
    rho_d[0] = 100.0 * (p[0] - vapor_pressure(p[0], q_v[0])) / (Tk[0] * Rd)
    rho_t[0] = rho_d[0] * (1.0 + q_v[0])
    dlnpdz = -gravity * rho_t[0] / (p[0] * 100.0)
    for i in range(1, n):
        p[i] = math.exp(math.log(p[i - 1]) + (dlnpdz * (z[i] - z[i - 1])))
        Tk[i] = theta[i] / (p_0 / p[i]) ** (Rd / Cpd)
        rho_d[i] = 100.0 * (p[i] - vapor_pressure(p[i], q_v[i])) / (Tk[i] * Rd)
        rho_t[i] = rho_d[i] * (1.0 + q_v[i])
        dlnpdz = -gravity * rho_t[i] / (p[i] * 100.0)
    col = cheb.Chebyshev1D(cheb.ChebyshevParameters(gp.zmin, gp.zmax, gp.zDim, gp.b_zDim))
    p_new = col.CIInttransform(col.CAtransform(col.CBtransform(-gravity * rho_t)), sfc_pressure * 100.0) / 100.0
    Tk = theta / (p_0 / p_new) ** (Rd / Cpd)
    rho_d = 100.0 * (p_new - vapor_pressure(p_new, q_v)) / (Tk * Rd)
    s3 = transform_reference_state(gp, entropy_any(Tk, rho_d, q_v))
    x3 = transform_reference_state(gp, np.log(rho_d / rho_d0))
    m3 = transform_reference_state(gp, bhyp(q_v))
    qq, rr, TT, _ = thermodynamic_tuple(s3[:, 0], x3[:, 0], m3[:, 0])
    Pxi_bar = float(np.mean(P_xi(TT, rr, qq) / (dry_density(x3[:, 0]) * (1.0 + ahyp(m3[:, 0])))))
    return ReferenceState(s3, x3, m3, np.zeros((n, 3)), Pxi_bar)


def entropy_any(Tk, rho_d, q_v):  # src/thermodynamics.jl:44-54 including the q_v == 0 branch
    q_v = np.asarray(q_v, dtype=np.float64)
    safe = np.where(q_v != 0.0, q_v * rho_d / rho_v0, 1.0)
    qfactor = np.where(q_v != 0.0, q_v * (Rv * np.log(safe) - (L_v(T_0) / T_0)), 0.0)
    return ((Cvd + (q_v * Cvv)) * np.log(Tk / T_0)) - (Rd * np.log(rho_d / rho_d0)) - qfactor


# ------------------------------------------------------------------ model tile
class ModelTile:  # src/semiimplicit.jl:18-124
    def __init__(self, patch: G.Grid, tile: G.Grid, model: ModelParameters,
                 haloReceiveRows: np.ndarray, ref_state: ReferenceState | None = None):
        N, V = tile.N, tile.V
        self.model = model
        self.tile = tile
        self.var_np1 = np.zeros((N, V))
        self.expdot_n = np.zeros((N, V))
        self.expdot_nm1 = np.zeros((N, V))
        self.expdot_nm2 = np.zeros((N, V))
        self.impdot_n = np.zeros((N, V))
        self.impdot_nm1 = np.zeros((N, V))
        self.impdot_nm2 = np.zeros((N, V))
        self.tilepoints = tile.getGridpoints()
        self.ref_state = ref_state if ref_state is not None else ReferenceState()
        self.patchSplines = patch.splines
        self.patchParams = patch.params
        self.patchSpectral = patch.spectral.copy()
        self.patchRows, self.tileRows = G.calcPatchMap(patch, tile)
        self.haloSendRows, self.haloTileRows = G.calcHaloMap(patch, tile)
        self.haloReceiveRows = haloReceiveRows
        self.h_matrix = None
        if model.options.get("semiimplicit", False):
            self.h_matrix = calc_Helmholtz_semiimplicit_matrix(model, self.ref_state.Pxi_bar, 1.25 * model.ts)


def explicit_timestep(mtile: ModelTile, t: int):  # src/semiimplicit.jl:672-698
    ts = mtile.model.ts
    phys = mtile.tile.physical[:, :, 0]
    if t == 1:
        mtile.var_np1[:] = phys + (ts * mtile.expdot_n)
        mtile.expdot_nm1[:] = mtile.expdot_n
    elif t == 2:
        mtile.var_np1[:] = phys + (0.5 * ts) * ((3.0 * mtile.expdot_n) - mtile.expdot_nm1)
        mtile.expdot_nm2[:] = mtile.expdot_nm1
        mtile.expdot_nm1[:] = mtile.expdot_n
    else:
        mtile.var_np1[:] = phys + ((ts / 12.0) * ((23.0 * mtile.expdot_n) - (16.0 * mtile.expdot_nm1)
                                                  + (5.0 * mtile.expdot_nm2)))
        mtile.expdot_nm2[:] = mtile.expdot_nm1
        mtile.expdot_nm1[:] = mtile.expdot_n


def calc_Helmholtz_semiimplicit_matrix(model: ModelParameters, Pxi_bar: float, ts_term: float) -> np.ndarray:
    """src/semiimplicit.jl:768-781 (returned unfactorised; solve with numpy)."""
    gp = model.grid_params
    nz = gp.zDim
    L = gp.zmax - gp.zmin
    dct = cheb.dct_matrix(nz)
    dct2 = cheb.dct_2nd_derivative(nz, L)
    h = (ts_term * ts_term * Pxi_bar) * dct2 - dct
    bc1 = (ts_term * ts_term * Pxi_bar) * dct[0, :]
    bc2 = (ts_term * ts_term * Pxi_bar) * dct[nz - 1, :]
    return np.vstack([bc1[None, :], bc2[None, :], h[1:nz - 1, :]])


def semiimplicit_adjustment(mtile: ModelTile, t: int):  # src/semiimplicit.jl:521-597, all columns at once
    gp = mtile.model.grid_params
    w_i = gp.vars["w"] - 1
    xi_i = gp.vars["xi"] - 1
    ts = mtile.model.ts
    nz = gp.zDim
    col = lambda a: a.reshape(-1, nz).T  # noqa: E731  [nz, ncols] view
    xi_nstar = col(mtile.var_np1[:, xi_i]).copy()
    wdot_n, wdot_nm1, wdot_nm2 = (col(a[:, xi_i]) for a in (mtile.impdot_n, mtile.impdot_nm1, mtile.impdot_nm2))
    w_nstar = col(mtile.var_np1[:, w_i]).copy()
    xidot_n, xidot_nm1, xidot_nm2 = (col(a[:, w_i]) for a in (mtile.impdot_n, mtile.impdot_nm1, mtile.impdot_nm2))
    Pxi_bar = mtile.ref_state.Pxi_bar
    if t == 1:
        ts_term = 0.5 * ts
        w_nstar = w_nstar - (ts * xidot_n) + (ts * 0.5 * xidot_n)
        xi_nstar = xi_nstar - (ts * wdot_n) + (ts * 0.5 * wdot_n)
    elif t == 2:
        ts_term = 1.25 * ts
        w_nstar = w_nstar - (0.5 * ts) * ((3.0 * xidot_n) - xidot_nm1) - (ts * xidot_n) + (ts * 0.75 * xidot_nm1)
        xi_nstar = xi_nstar - (0.5 * ts) * ((3.0 * wdot_n) - wdot_nm1) - (ts * wdot_n) + (ts * 0.75 * wdot_nm1)
    else:
        ts_term = 1.25 * ts
        w_nstar = w_nstar - ((ts / 12.0) * ((23.0 * xidot_n) - (16.0 * xidot_nm1) + (5.0 * xidot_nm2))) \
            - (ts * xidot_n) + (ts * 0.75 * xidot_nm1)
        xi_nstar = xi_nstar - ((ts / 12.0) * ((23.0 * wdot_n) - (16.0 * wdot_nm1) + (5.0 * wdot_nm2))) \
            - (ts * wdot_n) + (ts * 0.75 * wdot_nm1)
    # rotate implicit history for the two variables (views write through)
    xidot_nm2[:] = xidot_nm1
    xidot_nm1[:] = xidot_n
    wdot_nm2[:] = wdot_nm1
    wdot_nm1[:] = wdot_n

    xi_col = mtile.tile.columns[xi_i]
    a = xi_col.CAtransform(xi_col.CBtransform(xi_nstar))
    xi_nstar = xi_col.CItransform(a)
    xi_nstar_z = ts_term * Pxi_bar * xi_col.CIxtransform(a)
    g = xi_nstar_z - w_nstar
    g = np.vstack([np.zeros((2, g.shape[1])), g[1:nz - 1]])
    w_col = mtile.tile.columns[w_i]
    if t == 1:
        h_a = calc_Helmholtz_semiimplicit_matrix(mtile.model, Pxi_bar, ts_term)
    else:
        h_a = mtile.h_matrix
    wa = np.linalg.solve(h_a, g)
    mtile.var_np1[:, w_i] = w_col.CItransform(wa).T.reshape(-1)
    mtile.var_np1[:, xi_i] = (xi_nstar - (ts_term * w_col.CIxtransform(wa))).T.reshape(-1)


# ------------------------------------------------------------------ equation sets
def _slots(grid, v):
    return [grid.physical[:, v, d] for d in range(grid.D)]


def LinearAdvection1D(mtile, t):  # src/testModels.jl:1-20
    p = mtile.model.physical_params
    g = mtile.tile
    mtile.expdot_n[:, 0] = -(p["c_0"] * g.physical[:, 0, 1]) + (p["K"] * g.physical[:, 0, 2])
    explicit_timestep(mtile, t)


def LinearAdvectionRZ(mtile, t):  # src/testModels.jl:22-45
    K = mtile.model.physical_params["K"]
    g = mtile.tile
    r = mtile.tilepoints[:, 0]
    hr, hrr, hz, hzz = (g.physical[:, 0, d] for d in (1, 2, 3, 4))
    u = g.physical[:, 1, 0]
    w = g.physical[:, 3, 0]
    mtile.expdot_n[:, 0] = (-u * hr) + (-w * hz) + (K * ((hr / r) + hrr + hzz))
    explicit_timestep(mtile, t)


def LinearAdvectionRL(mtile, t):  # src/testModels.jl:47-73
    K = mtile.model.physical_params["K"]
    g = mtile.tile
    r = mtile.tilepoints[:, 0]
    hr, hrr, hl, hll = (g.physical[:, 0, d] for d in (1, 2, 3, 4))
    u = g.physical[:, 1, 0]
    v = g.physical[:, 2, 0]
    if K > 0.0:
        mtile.expdot_n[:, 0] = (-u * hr) - (v * (hl / r)) + (K * ((hr / r) + hrr + (hll / (r * r))))
    else:
        mtile.expdot_n[:, 0] = (-u * hr) - (v * (hl / r))
    explicit_timestep(mtile, t)


def LinearAdvectionRLZ(mtile, t):  # src/testModels.jl:75-98
    K = mtile.model.physical_params["K"]
    g = mtile.tile
    r = mtile.tilepoints[:, 0]
    hr, hrr, hl, hll = (g.physical[:, 0, d] for d in (1, 2, 3, 4))
    u = g.physical[:, 1, 0]
    v = g.physical[:, 2, 0]
    mtile.expdot_n[:, 0] = (-u * hr) - (v * (hl / r)) + (K * ((hr / r) + hrr + (hll / (r * r))))
    explicit_timestep(mtile, t)


def LinearShallowWater1D(mtile, t):  # src/shallowWaterModels.jl:235-259
    p = mtile.model.physical_params
    g = mtile.tile
    mtile.expdot_n[:, 0] = -p["H"] * g.physical[:, 1, 1]
    mtile.expdot_n[:, 1] = (-p["g"] * g.physical[:, 0, 1]) + (p["K"] * g.physical[:, 1, 2])
    explicit_timestep(mtile, t)


def LinearShallowWaterRL(mtile, t):  # src/shallowWaterModels.jl:261-298
    p = mtile.model.physical_params
    gg, K, H = p["g"], p["K"], p["H"]
    g = mtile.tile
    r = mtile.tilepoints[:, 0]
    h, hr, hrr, hl, hll = _slots(g, 0)
    u, ur, urr, ul, ull = _slots(g, 1)
    v, vr, vrr, vl, vll = _slots(g, 2)
    mtile.expdot_n[:, 0] = -H * ((u / r) + ur + (vl / r))
    mtile.expdot_n[:, 1] = (-gg * hr) + (K * ((ur / r) + urr + (ull / (r * r))))
    mtile.expdot_n[:, 2] = (-gg * (hl / r)) + (K * ((vr / r) + vrr + (vll / (r * r))))
    explicit_timestep(mtile, t)


def _shallow_water_slab(mtile, t, twoway: bool):  # src/shallowWaterModels.jl:1-113, 115-233
    p = mtile.model.physical_params
    g, K, Cd, Hfree, Hb, f = p["g"], p["K"], p["Cd"], p["Hfree"], p["Hb"], p["f"]
    grid = mtile.tile
    e = mtile.expdot_n
    r = mtile.tilepoints[:, 0]
    h, hr, hrr, hl, hll = _slots(grid, 0)
    ug, ugr, ugrr, ugl, ugll = _slots(grid, 1)
    vg, vgr, vgrr, vgl, vgll = _slots(grid, 2)
    ub, ubr, ubrr, ubl, ubll = _slots(grid, 3)
    vb, vbr, vbrr, vbl, vbll = _slots(grid, 4)
    U = 0.78 * np.sqrt((ub * ub) + (vb * vb))
    w = grid.physical[:, 5, 0]
    w[:] = -Hb * ((ub / r) + ubr + (vbl / r))
    w_ = 0.5 * np.abs(w) - w
    e[:, 5] = 0.0
    ADV = (-vg * hl / r) + (-ug * hr)
    PGF = (-(Hfree + h) * ((ug / r) + ugr + (vgl / r)))
    if twoway:
        COR = -(Hfree + h) * w * p["S1"]
        e[:, 0] = ADV + PGF + COR
    else:
        e[:, 0] = ADV + PGF
    e[:, 1] = ((-vg * ugl / r) + (-ug * ugr)) + (-g * hr) + (vg * (f + (vg / r)))
    e[:, 2] = ((-vg * vgl / r) + (-ug * vgr)) + (-g * (hl / r)) + (-ug * (f + (vg / r)))
    ADV = (-vb * ubl / r) + (-ub * ubr)
    PGF = (-g * hr)
    COR = (vb * (f + (vb / r)))
    DRAG = -(Cd * U * ub / Hb)
    W_ = w_ * (ug - ub) / Hb
    KDIFF = K * ((ubr / r) + ubrr - (ub / (r * r)) + (ubll / (r * r)) - (2.0 * vbl / (r * r)))
    e[:, 3] = ADV + PGF + COR + DRAG + W_ + KDIFF
    ADV = (-vb * vbl / r) + (-ub * vbr)
    PGF = (-g * (hl / r))
    COR = (-ub * (f + (vb / r)))
    DRAG = -(Cd * U * vb / Hb)
    W_ = w_ * (vg - vb) / Hb
    KDIFF = K * ((vbr / r) + vbrr - (vb / (r * r)) + (vbll / (r * r)) + (2.0 * ubl / (r * r)))
    e[:, 4] = ADV + PGF + COR + DRAG + W_ + KDIFF
    explicit_timestep(mtile, t)


def Oneway_ShallowWater_Slab(mtile, t):
    _shallow_water_slab(mtile, t, False)


def Twoway_ShallowWater_Slab(mtile, t):
    _shallow_water_slab(mtile, t, True)


def Oneway_ShallowWater_HeightResolvedBL(mtile, t):  # src/shallowWaterModels.jl:346-511, all columns at once
    p = mtile.model.physical_params
    g, Kh, Cd_user, Hfree, f, Um, Vm = p["g"], p["Kh"], p["Cd"], p["Hfree"], p["f"], p["Um"], p["Vm"]
    grid = mtile.tile
    gp = mtile.model.grid_params
    nz = gp.zDim
    e = mtile.expdot_n
    r, lam, z = mtile.tilepoints[:, 0], mtile.tilepoints[:, 1], mtile.tilepoints[:, 2]
    h, hr, hrr, hl, hll = (grid.physical[:, 0, d] for d in range(5))
    ug, ugr, ugrr, ugl, ugll = (grid.physical[:, 1, d] for d in range(5))
    vg, vgr, vgrr, vgl, vgll = (grid.physical[:, 2, d] for d in range(5))
    ub, ubr, ubrr, ubl, ubll, ubz, ubzz = (grid.physical[:, 3, d] for d in range(7))
    vb, vbr, vbrr, vbl, vbll, vbz, vbzz = (grid.physical[:, 4, d] for d in range(7))
    S = np.sqrt((ubz * ubz) + (vbz * vbz))
    with np.errstate(divide="ignore"):
        l = 1.0 / ((1.0 / (0.4 * z)) + (1.0 / 80.0))
    Kv = (l ** 2) * S
    hcol = grid.columns[gp.vars["h"] - 1]
    col = lambda a: a.reshape(-1, nz).T  # noqa: E731
    flat = lambda a: a.T.reshape(-1)  # noqa: E731
    div = -((ub / r) + ubr + (vbl / r))
    wb = flat(hcol.CIInttransform(hcol.CAtransform(hcol.CBtransform(col(div)))))
    grid.physical[:, 5, 0] = wb
    e[:, 5] = 0.0
    e[:, 0] = ((-vg * hl / r) + (-ug * hr)) + (-(Hfree + h) * ((ug / r) + ugr + (vgl / r)))
    e[:, 1] = ((-vg * ugl / r) + (-ug * ugr)) + (-g * hr) + (vg * (f + (vg / r)))
    e[:, 2] = ((-vg * vgl / r) + (-ug * vgr)) + (-g * (hl / r)) + (-ug * (f + (vg / r)))
    # surface wind from storm motion, 10 m wind = level 2
    lam0 = col(lam)[0]
    sfcu = (Um * np.cos(lam0)) + (Vm * np.sin(lam0))
    sfcv = (Vm * np.cos(lam0)) - (Um * np.sin(lam0))
    u10 = col(ub)[1] + sfcu
    v10 = col(vb)[1] + sfcv
    U10 = np.sqrt(u10 ** 2 + v10 ** 2)
    Cd = np.where(U10 < 5.2, 1.0e-3, np.where(U10 < 33.6, 4.4e-4 * U10 ** 0.5, Cd_user))

    def vdiff(fz, sfc):
        m = col(Kv * fz).copy()
        m[0] = Cd * U10 * sfc
        return flat(hcol.CIxtransform(hcol.CAtransform(hcol.CBtransform(m))))

    ADV = (-vb * ubl / r) + (-ub * ubr) + (-wb * ubz)
    PGF = (-g * hr)
    COR = (vb * (f + (vb / r)))
    HDIFF = Kh * ((ubr / r) + ubrr - (ub / (r * r)) + (ubll / (r * r)) - (2.0 * vbl / (r * r)))
    e[:, 3] = ADV + PGF + COR + vdiff(ubz, u10) + HDIFF
    ADV = (-vb * vbl / r) + (-ub * vbr) + (-wb * vbz)
    PGF = (-g * (hl / r))
    COR = (-ub * (f + (vb / r)))
    HDIFF = Kh * ((vbr / r) + vbrr - (vb / (r * r)) + (vbll / (r * r)) + (2.0 * ubl / (r * r)))
    e[:, 4] = ADV + PGF + COR + vdiff(vbz, v10) + HDIFF
    explicit_timestep(mtile, t)


def Euler_test(mtile, t):  # src/testModels.jl:100-215
    K = mtile.model.physical_params["K"]
    grid = mtile.tile
    nz = mtile.model.grid_params.zDim
    e, imp = mtile.expdot_n, mtile.impdot_n
    ref = mtile.ref_state
    ncols = grid.N // nz
    rep = lambda a: np.tile(a, ncols)  # noqa: E731
    s, s_x, s_xx, s_z, s_zz = _slots(grid, 0)
    xi, xi_x, xi_xx, xi_z, xi_zz = _slots(grid, 1)
    mu, mu_x, mu_xx, mu_z, mu_zz = _slots(grid, 2)
    u, u_x, u_xx, u_z, u_zz = _slots(grid, 3)
    w, w_x, w_xx, w_z, w_zz = _slots(grid, 4)
    sbar, sbar_z = rep(ref.sbar[:, 0]), rep(ref.sbar[:, 1])
    xibar, xibar_z = rep(ref.xibar[:, 0]), rep(ref.xibar[:, 1])
    mubar, mubar_z = rep(ref.mubar[:, 0]), rep(ref.mubar[:, 1])
    q_v, rho_d, Tk, p = thermodynamic_tuple(s + sbar, xi + xibar, mu + mubar)
    rho_t = rho_d * (1.0 + q_v)
    qvp_x = mu_x / dmudq(mu + mubar, q_v)
    qvp_z = mu_z / dmudq(mu + mubar, q_v)
    rhobar = dry_density(xibar) * (1.0 + ahyp(mubar))
    rho_p = rho_t - rhobar
    Pxi_bar = ref.Pxi_bar
    e[:, 0] = ((-u * s_x) + (-w * (s_z + sbar_z))) + (K * (s_xx + s_zz))
    e[:, 1] = ((-u * xi_x) + (-w * (xi_z + xibar_z))) - u_x - w_z
    imp[:, 1] = -w_z
    e[:, 2] = ((-u * mu_x) + (-w * (mu_z + mubar_z))) + (K * (mu_xx + mu_zz))
    e[:, 3] = ((-u * u_x) + (-w * u_z)) + (-(pressure_gradient(Tk, rho_d, q_v, s_x, xi_x, qvp_x) / rho_t)) \
        + (K * (u_xx + u_zz))
    PGF = -(gravity * rho_p / rho_t) - (pressure_gradient(Tk, rho_d, q_v, s_z, xi_z, qvp_z) / rho_t)
    e[:, 4] = ((-u * w_x) + (-w * w_z)) + PGF + (K * (w_xx + w_zz))
    imp[:, 4] = -(Pxi_bar * xi_z)
    explicit_timestep(mtile, t)
    if mtile.model.options.get("semiimplicit", False):
        semiimplicit_adjustment(mtile, t)


def _moist_test(mtile, t, rain: bool):
    """BF02_test (src/testModels.jl:217-385) and rainfall_test (:387-586): shared structure, restated term by term."""
    K = mtile.model.physical_params["K"]
    grid = mtile.tile
    nz = mtile.model.grid_params.zDim
    e, imp = mtile.expdot_n, mtile.impdot_n
    ref = mtile.ref_state
    ncols = grid.N // nz
    rep = lambda a: np.tile(a, ncols)  # noqa: E731
    s, s_x, s_xx, s_z, s_zz = _slots(grid, 0)
    xi, xi_x, xi_xx, xi_z, xi_zz = _slots(grid, 1)
    mu, mu_x, mu_xx, mu_z, mu_zz = _slots(grid, 2)
    u, u_x, u_xx, u_z, u_zz = _slots(grid, 3)
    w, w_x, w_xx, w_z, w_zz = _slots(grid, 4)
    sbar, sbar_z = rep(ref.sbar[:, 0]), rep(ref.sbar[:, 1])
    xibar, xibar_z = rep(ref.xibar[:, 0]), rep(ref.xibar[:, 1])
    mubar, mubar_z = rep(ref.mubar[:, 0]), rep(ref.mubar[:, 1])
    mu_total = mu + mubar
    q_v, rho_d, Tk, p = thermodynamic_tuple(s + sbar, xi + xibar, mu_total)
    if rain:
        mu_c, mu_c_x, mu_c_xx, mu_c_z, mu_c_zz = _slots(grid, 5)
        mu_r, mu_r_x, mu_r_xx, mu_r_z, mu_r_zz = _slots(grid, 6)
        qss, qss_x, _, qss_z, _ = _slots(grid, 7)
        q_c = ahyp(mu_c)
        q_r = ahyp(mu_r)
        q_l = q_c + q_r
        q_t = q_v + q_l
        rho_t = rho_d * (1.0 + q_t)
        N_c = 100.0
    else:
        mu_l, mu_l_x, mu_l_xx, mu_l_z, mu_l_zz = _slots(grid, 5)
        qss, qss_x, _, qss_z, _ = _slots(grid, 6)
        mu_lbar, mu_lbar_z = rep(ref.mu_lbar[:, 0]), rep(ref.mu_lbar[:, 1])
        q_l = ahyp(mu_l + mu_lbar)
        rho_t = rho_d * (1.0 + q_v + q_l)
        N_c = 500.0
    r_c = 10.0
    mu_factor = dmudq(mu_total, q_v)
    qvp_x = mu_x / mu_factor
    qvp_z = mu_z / mu_factor
    rhobar = dry_density(xibar) * (1.0 + ahyp(mubar))
    rho_p = rho_t - rhobar
    Pxi_bar = ref.Pxi_bar
    dpdx = pressure_gradient(Tk, rho_d, q_v, s_x, xi_x, qvp_x)
    dpdz = pressure_gradient(Tk, rho_d, q_v, s_z, xi_z, qvp_z)
    Cm = (q_l * Cl) / (Cvd + (q_v * Cvv) + (q_l * Cl))
    s_div = Cm * (Rd + q_v * Rv) * (u_x + w_z)
    q_cond = q_condensation(qss, Tk, p, q_v, q_l, N_c, r_c)
    s_cond = s_condensation(q_cond, Tk, rho_d, q_v, q_l, p)
    cloudtau = invtau_condensation(Tk, p, N_c, r_c)
    lift = (u * dpdx) + (w * (dpdz - rhobar * gravity))
    if rain:
        raintau = rain_evaporation(q_r, rho_d, Tk, p)
        q_evap = -qss * raintau
        qss_cond = dqsdp(Tk, p, rho_d, q_v, q_l) * lift - qss * (cloudtau + raintau)
        q_auto = autoconversion(q_c, rho_d)
        q_coll = collection(q_c, q_r, rho_d, Tk)
        Vt = sedimentation(q_r, rho_d, Tk)
        col = grid.columns[mtile.model.grid_params.vars["mu_r"] - 1]     # :525-529
        flux = (q_r * Vt).reshape(ncols, nz).T
        Vt_flux = col.CIxtransform(col.CAtransform(col.CBtransform(flux))).T.reshape(-1) / rho_d
    else:
        qss_cond = dqsdp(Tk, p, rho_d, q_v, q_l) * lift - qss * cloudtau
    e[:, 0] = ((-u * s_x) + (-w * (s_z + sbar_z))) + (s_cond + s_div) + (K * (s_xx + s_zz))
    e[:, 1] = ((-u * xi_x) + (-w * (xi_z + xibar_z))) + (-u_x - w_z)
    imp[:, 1] = -w_z
    if rain:
        e[:, 2] = ((-u * mu_x) + (-w * (mu_z + mubar_z))) + (mu_factor * (q_evap - q_cond)) + (K * (mu_xx + mu_zz))
    else:
        e[:, 2] = ((-u * mu_x) + (-w * (mu_z + mubar_z))) + (-q_cond * mu_factor) + (K * (mu_xx + mu_zz))
    imp[:, 2] = q_v
    e[:, 3] = ((-u * u_x) + (-w * u_z)) + (-dpdx / rho_t) + (K * (u_xx + u_zz))
    e[:, 4] = ((-u * w_x) + (-w * w_z)) + (((-gravity * rho_p) - dpdz) / rho_t) + (K * (w_xx + w_zz))
    imp[:, 4] = -(Pxi_bar * xi_z)
    if rain:
        e[:, 5] = ((-u * mu_c_x) + (-w * mu_c_z)) + (dmudq(mu_c, q_c) * (q_cond - q_auto - q_coll)) + (K * (mu_c_xx + mu_c_zz))
        e[:, 6] = ((-u * mu_r_x) + (-w * mu_r_z)) + (dmudq(mu_r, q_r) * (q_auto + q_coll - q_evap - Vt_flux)) \
            + (K * (mu_r_xx + mu_r_zz))
        e[:, 7] = ((-u * qss_x) + (-w * qss_z)) + qss_cond
        imp[:, 7] = qss
    else:
        e[:, 5] = ((-u * mu_l_x) + (-w * (mu_l_z + mu_lbar_z))) + (q_cond * dmudq(mu_l, q_l)) + (K * (mu_l_xx + mu_l_zz))
        e[:, 6] = ((-u * qss_x) + (-w * qss_z)) + qss_cond
        imp[:, 6] = qss
    explicit_timestep(mtile, t)
    if mtile.model.options.get("semiimplicit", False):
        semiimplicit_adjustment(mtile, t)
    condensation_adjustment(mtile, t)


def BF02_test(mtile, t):  # src/testModels.jl:217-385
    _moist_test(mtile, t, rain=False)


def rainfall_test(mtile, t):  # src/testModels.jl:387-586
    _moist_test(mtile, t, rain=True)


EQUATION_SETS = {f.__name__: f for f in (
    LinearAdvection1D, LinearAdvectionRZ, LinearAdvectionRL, LinearAdvectionRLZ,
    LinearShallowWater1D, LinearShallowWaterRL, Oneway_ShallowWater_Slab, Twoway_ShallowWater_Slab,
    Oneway_ShallowWater_HeightResolvedBL, Euler_test, BF02_test, rainfall_test)}


def physical_model(mtile, t):  # src/semiimplicit.jl:357-363
    try:
        fn = EQUATION_SETS[mtile.model.equation_set]
    except KeyError:
        raise KeyError(f"equation set {mtile.model.equation_set!r} is not defined") from None
    fn(mtile, t)


def checkCFL(grid: G.Grid):  # src/semiimplicit.jl:737-751
    for name, v in grid.params.vars.items():
        bad = np.flatnonzero(np.isnan(grid.physical[:, v - 1, 0]))
        if bad.size:
            raise RuntimeError(f"NaN found in variable {name} at index{bad[0] + 1} ! CFL condition likely violated")


# ------------------------------------------------------------------ driver
class ModelRun:
    """initialize_model + run_model with `num_tiles` in-process tiles (src/semiimplicit.jl:126-299)."""

    def __init__(self, model: ModelParameters, num_tiles: int, ic: np.ndarray,
                 ref_state: ReferenceState | None = None, workers: int = 1):
        self.model = model
        self.patch = G.createGrid(model.grid_params)
        self.patch.workers = workers
        self.patch.physical[:, :, 0] = ic
        self.patch.spectralTransform()
        self.patch.gridTransform()
        self.tile_params = G.calcTileSizes(self.patch, num_tiles)
        self.mtiles = []
        halo = np.zeros(0, dtype=np.int64)
        for t in range(num_tiles):
            tile = G.createGrid(G.tile_params(self.patch, self.tile_params, t))
            tile.workers = workers
            mt = ModelTile(self.patch, tile, model, halo, ref_state)
            halo = mt.haloSendRows
            self.mtiles.append(mt)
        self.lastHaloRows = halo
        self.sharedSpectral = self.patch.spectral.copy()
        self._spline_transform_all()
        self.t = 0

    def _spline_transform_all(self):
        # every worker solves the whole patch (replicated); do it once and share the result
        A = np.empty_like(self.sharedSpectral)
        G.splineTransform(self.patch.splines, A, self.patch.params, self.sharedSpectral)
        for mt in self.mtiles:
            mt.patchSpectral = A

    def advanceTimestep(self, mt: ModelTile, t: int, haloIn: np.ndarray | None):
        G.tileTransform(mt.patchSplines, mt.patchSpectral, mt.patchParams, mt.tile)
        physical_model(mt, t)
        mt.tile.physical[:, :, :] = mt.var_np1[:, :, None]      # calcTendency (:731)
        mt.tile.spectralTransform()
        haloOut = mt.tile.spectral[mt.haloTileRows].copy()       # put!(haloSend, haloSendView)
        self.sharedSpectral[mt.patchRows] = mt.tile.spectral[mt.tileRows]
        if haloIn is not None and mt.haloReceiveRows.size:
            self.sharedSpectral[mt.haloReceiveRows] += haloIn
        return haloOut

    def step(self):
        self.t += 1
        self.sharedSpectral[:] = 0.0
        halo = None
        for mt in self.mtiles:
            halo = self.advanceTimestep(mt, self.t, halo)
        self.sharedSpectral[self.lastHaloRows] += halo
        self._spline_transform_all()

    def run(self, nsteps: int):
        for _ in range(nsteps):
            self.step()

    def output_patch(self) -> np.ndarray:
        """patch.spectral <- A; tileTransform!(patch) (src/semiimplicit.jl:289-291)."""
        A = self.mtiles[0].patchSpectral
        G.tileTransform(self.patch.splines, A, self.patch.params, self.patch, patch=self.patch)
        checkCFL(self.patch)
        return self.patch.physical


def integrate_model(model: ModelParameters, ic: np.ndarray, num_tiles: int = 1, ref_state=None) -> np.ndarray:
    run = ModelRun(model, num_tiles, ic, ref_state)
    run.run(int(round(model.integration_time / model.ts)))
    return run.output_patch()
