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


EQUATION_SETS = {f.__name__: f for f in (
    LinearAdvection1D, LinearAdvectionRZ, LinearAdvectionRL, LinearAdvectionRLZ,
    LinearShallowWater1D, LinearShallowWaterRL, Oneway_ShallowWater_Slab, Twoway_ShallowWater_Slab,
    Oneway_ShallowWater_HeightResolvedBL, Euler_test)}


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
