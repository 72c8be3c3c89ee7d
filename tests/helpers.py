"""Shared builders for the parity tests: identical seeded inputs for the oracle and the CUDA path."""
from __future__ import annotations

import numpy as np

import scythe_jl_b200 as S
from oracle import chebyshev as ch
from oracle import grids as G
from oracle import model as M
from oracle import splines as spl

TRANSFORM_TOL = 1e-12   # BASELINE.json north_star: transform round-trip <= 1e-12 (relative)
STATE_TOL = 1e-9        # N-step state <= 1e-9 (relative)


def to_pkg(gp: G.GridParameters) -> S.GridParameters:
    names = gp.var_names()
    sb = lambda d: {n: getattr(S.CubicBSpline, spl.bc_name(gp.var_bc(d, n))) for n in names}  # noqa: E731
    cb = lambda d: {n: getattr(S.Chebyshev, ch.bc_name(gp.var_bc(d, n))) for n in names}  # noqa: E731
    return S.GridParameters(geometry=gp.geometry, xmin=gp.xmin, xmax=gp.xmax, num_cells=gp.num_cells, l_q=gp.l_q,
                            BCL=sb("BCL"), BCR=sb("BCR"), zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, b_zDim=gp.b_zDim,
                            BCB=cb("BCB"), BCT=cb("BCT"), vars=dict(gp.vars), spectralIndexL=gp.spectralIndexL)


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max |b| over the whole array (slots that are identically ~0 must be compared
    against the scale of the full field, so callers pass whole [N,V] slots of non-trivial data)."""
    s = float(np.abs(b).max())
    return float(np.abs(a - b).max()) / (s if s > 0 else 1.0)


def slot_errs(a: np.ndarray, b: np.ndarray) -> list[float]:
    return [rel_err(a[:, :, d], b[:, :, d]) for d in range(b.shape[2])]


# ------------------------------------------------------------------ transform cases
def transform_cases():
    B = spl
    return {
        "R_periodic": G.GridParameters(geometry="R", xmin=-50, xmax=50, num_cells=40, BCL={"u": B.PERIODIC},
                                       BCR={"u": B.PERIODIC}, vars={"u": 1}),
        "R_mixed_bcs": G.GridParameters(geometry="R", xmin=0, xmax=10, num_cells=17,
                                        BCL={"a": B.R0, "b": B.R1T0, "c": B.R1T2, "d": B.R2T10, "e": B.R3},
                                        BCR={"a": B.R1T1, "b": B.R2T20, "c": B.R0, "d": B.R1T0, "e": B.R3},
                                        vars={"a": 1, "b": 2, "c": 3, "d": 4, "e": 5}),
        # b_rDim = 103 > 3 tiles of 32: streaming banded solve / chunked radial kernels across tile seams
        "R_100cells_bcs": G.GridParameters(geometry="R", xmin=0, xmax=10, num_cells=100,
                                           BCL={"a": B.R0, "b": B.R1T0, "c": B.R1T2, "d": B.R2T10, "e": B.R3, "g": B.R1T1},
                                           BCR={"a": B.R1T1, "b": B.R2T20, "c": B.R0, "d": B.R1T0, "e": B.R3, "g": B.R2T10},
                                           vars={"a": 1, "b": 2, "c": 3, "d": 4, "e": 5, "g": 6}),
        "R_61cells_bcs": G.GridParameters(geometry="R", xmin=0, xmax=10, num_cells=61,
                                          BCL={"a": B.R3, "b": B.R2T20}, BCR={"a": B.R2T20, "b": B.R1T1}, vars={"a": 1, "b": 2}),
        # b_rDim = 66 / 65: the right-edge BC fold straddles the last two 32-coefficient tiles
        "R_63cells_bcs": G.GridParameters(geometry="R", xmin=0, xmax=10, num_cells=63,
                                          BCL={"a": B.R1T0, "b": B.R0, "c": B.R1T1}, BCR={"a": B.R2T10, "b": B.R1T2, "c": B.R3},
                                          vars={"a": 1, "b": 2, "c": 3}),
        "R_62cells_bcs": G.GridParameters(geometry="R", xmin=0, xmax=10, num_cells=62,
                                          BCL={"a": B.R2T20, "b": B.R1T1}, BCR={"a": B.R2T20, "b": B.R1T0}, vars={"a": 1, "b": 2}),
        "RL": G.GridParameters(geometry="RL", xmin=0, xmax=10, num_cells=6, BCL={"h": B.R1T1, "u": B.R1T0},
                               BCR={"h": B.R0, "u": B.R1T1}, vars={"h": 1, "u": 2}),
        "RL_tile": G.GridParameters(geometry="RL", xmin=3, xmax=7, num_cells=4, vars={"h": 1}, spectralIndexL=4),
        "RZ": G.GridParameters(geometry="RZ", xmin=0, xmax=10, num_cells=7, zmin=0, zmax=5, zDim=12,
                               BCL={"s": B.R1T1, "w": B.R2T10}, BCR={"s": B.R3, "w": B.R1T2},
                               BCB={"s": ch.R0, "w": ch.R1T0}, BCT={"s": ch.R0, "w": ch.R1T0}, vars={"s": 1, "w": 2}),
        "RLZ": G.GridParameters(geometry="RLZ", xmin=0, xmax=10, num_cells=4, zmin=0, zmax=5, zDim=10,
                                BCL={"h": B.R1T1, "u": B.R1T0}, BCR={"h": B.R0, "u": B.R2T20},
                                BCB={"h": ch.R0, "u": ch.R1T1}, BCT={"h": ch.R0, "u": ch.R1T2},
                                vars={"h": 1, "u": 2}),
        # no vertical BCs + zDim in {16, 32, 64}: the DMMA (tensor-core) Chebyshev synthesis path
        "RLZ_z16_nobc": G.GridParameters(geometry="RLZ", xmin=0, xmax=10, num_cells=3, zmin=0, zmax=5, zDim=16,
                                         BCL={"h": B.R1T1, "u": B.R1T0}, vars={"h": 1, "u": 2}),
        # rings with m = ri+1 >= 65 use the persistent table-resident Bluestein kernels (L = 256)
        "RL_24cells_fft2": G.GridParameters(geometry="RL", xmin=0, xmax=24, num_cells=24, BCL={"h": B.R1T1}, vars={"h": 1}),
        "RLZ_22cells_fft2": G.GridParameters(geometry="RLZ", xmin=0, xmax=22, num_cells=22, zmin=0, zmax=5, zDim=8,
                                             BCL={"h": B.R1T1}, vars={"h": 1}),
        # outer-tile rings with 512 < m <= 768: composite Bluestein length L = 1536 = 3 x 512 (three-team radix-3 split)
        "RL_tile_fft3": G.GridParameters(geometry="RL", xmin=171, xmax=174, num_cells=3, vars={"h": 1, "u": 2}, spectralIndexL=172),
        "RLZ_tile_fft3": G.GridParameters(geometry="RLZ", xmin=171, xmax=173, num_cells=2, zmin=0, zmax=5, zDim=8,
                                          vars={"h": 1}, spectralIndexL=172),
        # outer-tile rings in every power-of-two convolution class of the v4 kernels (Tensor-Memory tables, bulk-copy staging):
        # m = ri + 1 = 181..183 (L = 512), 301..303 (1024), 871..876 (2048; odd and even row offsets), 1801..1803 (4096: two
        # table sets per TMEM lane)
        "RL_tile_fft4_L512": G.GridParameters(geometry="RL", xmin=60, xmax=61, num_cells=1, vars={"h": 1, "u": 2}, spectralIndexL=61),
        "RL_tile_fft4_L1024": G.GridParameters(geometry="RL", xmin=100, xmax=101, num_cells=1, vars={"h": 1}, spectralIndexL=101),
        "RL_tile_fft4_L2048": G.GridParameters(geometry="RL", xmin=290, xmax=292, num_cells=2, vars={"h": 1, "u": 2}, spectralIndexL=291),
        "RLZ_tile_fft4_L2048": G.GridParameters(geometry="RLZ", xmin=290, xmax=291, num_cells=1, zmin=0, zmax=5, zDim=8,
                                                vars={"h": 1}, spectralIndexL=291),
        "RL_tile_fft4_L4096": G.GridParameters(geometry="RL", xmin=600, xmax=601, num_cells=1, vars={"h": 1}, spectralIndexL=601),
        # composite lengths with the v4 data movement (k_inv_l5 / k_fwd_l5): m = 1030..1032 (L = 3072 = 3 x 1024, two groups of three
        # teams) and m = 2071..2073 (L = 6144 = 3 x 2048, one group)
        "RL_tile_fft5_L3072": G.GridParameters(geometry="RL", xmin=343, xmax=344, num_cells=1, vars={"h": 1, "u": 2}, spectralIndexL=344),
        "RLZ_tile_fft5_L3072": G.GridParameters(geometry="RLZ", xmin=343, xmax=344, num_cells=1, zmin=0, zmax=5, zDim=8,
                                                vars={"h": 1}, spectralIndexL=344),
        "RL_tile_fft5_L6144": G.GridParameters(geometry="RL", xmin=690, xmax=691, num_cells=1, vars={"h": 1}, spectralIndexL=691),
        "RZ_z32_nobc": G.GridParameters(geometry="RZ", xmin=0, xmax=10, num_cells=13, zmin=0, zmax=5, zDim=32,
                                        vars={"s": 1, "w": 2}),
    }


def check_transforms(gp: G.GridParameters, lib, seed=0):
    """spectralTransform! and gridTransform! of seeded random data: CUDA path vs oracle."""
    og = G.createGrid(gp)
    g = S.createGrid(to_pkg(gp), lib=lib)
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((og.N, og.V))
    og.physical[:, :, 0] = u
    g.physical[:, :, 0] = u
    pts = S.getGridpoints(g)
    assert np.abs(pts.reshape(og.N, -1) - og.getGridpoints().reshape(og.N, -1)).max() < 1e-12 * max(1.0, abs(gp.xmax))
    og.spectralTransform()
    S.spectralTransform(g)
    eB = rel_err(g.spectral, og.spectral)
    og.gridTransform()
    S.gridTransform(g)
    eP = slot_errs(g.physical, og.physical)
    g.close()
    return eB, eP


# ------------------------------------------------------------------ model cases
def _slab_case(nc=6):
    names = ["h", "u", "v", "ub", "vb", "wb"]
    BCL = {"h": spl.R1T1, "u": spl.R1T0, "v": spl.R1T0, "ub": spl.R1T0, "vb": spl.R1T0, "wb": spl.R1T1}
    BCR = {"h": spl.R0, "u": spl.R1T1, "v": spl.R0, "ub": spl.R1T1, "vb": spl.R0, "wb": spl.R0}
    gp = G.GridParameters(geometry="RL", xmin=0, xmax=3e5, num_cells=nc, BCL=BCL, BCR=BCR,
                          vars={n: i + 1 for i, n in enumerate(names)})
    r, l = G.createGrid(gp).getGridpoints().T
    Rmax, V0 = 5e4, 50.0 / 5e4
    vbar = np.where(r < Rmax, V0 * r, Rmax * Rmax * V0 / r)
    ic = np.zeros((r.size, 6))
    ic[:, 2] = vbar * (1 + 0.05 * np.cos(2 * l))
    ic[:, 4] = vbar
    ic[:, 3] = -0.1 * vbar
    ic[:, 0] = 100 * np.exp(-(r / 1e5) ** 2)
    prm = dict(g=9.81, K=5000.0, Cd=2.4e-3, Hfree=2000.0, Hb=1000.0, f=5e-5, S1=1e-5)
    return gp, ic, prm


def fused_advection_case(num_cells, tiles=(1, 2), zDim=16):
    """no vertical BCs: the fused synthesis + tendency + AB3 kernel (k_inv_z_advection); 64 levels: its blocked-SZ form"""
    gpf = G.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=num_cells, zmin=0, zmax=1e3, zDim=zDim,
                           vars={"h": 1, "u": 2, "v": 3})
    r, l, z = G.createGrid(gpf).getGridpoints().T
    icf = np.zeros((r.size, 3))
    icf[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8) * np.cos(z / 400.0)
    icf[:, 1] = 5 * np.cos(l) * (1 + 0.2 * np.sin(z / 300.0))
    icf[:, 2] = -5 * np.sin(l) * np.exp(-z / 900.0)
    return dict(gp=gpf, eq="LinearAdvectionRLZ", prm={"K": 100.0}, ts=50.0, n=4, ic=icf, tiles=tiles)


def model_cases(small=True):
    cases = {}
    gp = G.GridParameters(geometry="R", xmin=-50, xmax=50, num_cells=30, BCL={"u": spl.PERIODIC},
                          BCR={"u": spl.PERIODIC}, vars={"u": 1})
    x = G.createGrid(gp).getGridpoints()
    cases["LinearAdvection1D"] = dict(gp=gp, eq="LinearAdvection1D", prm={"c_0": 1.0, "K": 0.01}, ts=0.05, n=8,
                                      ic=np.exp(-(x / 20) ** 2)[:, None], tiles=(1, 3))
    gp = G.GridParameters(geometry="R", xmin=0, xmax=100, num_cells=24, BCL={"h": spl.R1T1, "u": spl.R1T0},
                          BCR={"h": spl.R1T1, "u": spl.R1T0}, vars={"h": 1, "u": 2})
    x = G.createGrid(gp).getGridpoints()
    cases["LinearShallowWater1D"] = dict(gp=gp, eq="LinearShallowWater1D", prm={"g": 9.81, "K": 0.5, "H": 10.0}, ts=0.02,
                                         n=5, ic=np.stack([np.exp(-((x - 50) / 10) ** 2), 0 * x], 1), tiles=(2,))
    gp, ic, prm = _slab_case(6)
    cases["Oneway_ShallowWater_Slab"] = dict(gp=gp, eq="Oneway_ShallowWater_Slab", prm=prm, ts=3.0, n=4, ic=ic, tiles=(1, 2))
    cases["Twoway_ShallowWater_Slab"] = dict(gp=gp, eq="Twoway_ShallowWater_Slab", prm=prm, ts=3.0, n=3, ic=ic, tiles=(2,))
    gp3 = G.GridParameters(geometry="RL", xmin=0, xmax=1e5, num_cells=6, vars={"h": 1, "u": 2, "v": 3})
    r, l = G.createGrid(gp3).getGridpoints().T
    ic3 = np.zeros((r.size, 3))
    ic3[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8)
    ic3[:, 1] = 5 * np.cos(l)
    ic3[:, 2] = -5 * np.sin(l)
    cases["LinearAdvectionRL_K0"] = dict(gp=gp3, eq="LinearAdvectionRL", prm={"K": 0.0}, ts=50.0, n=3, ic=ic3, tiles=(2,))
    cases["LinearAdvectionRL"] = dict(gp=gp3, eq="LinearAdvectionRL", prm={"K": 100.0}, ts=50.0, n=3, ic=ic3, tiles=(1,))
    cases["LinearShallowWaterRL"] = dict(gp=gp3, eq="LinearShallowWaterRL", prm={"K": 100.0, "g": 9.81, "H": 1000.0},
                                         ts=5.0, n=3, ic=ic3, tiles=(2,))
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=6, zmin=0, zmax=1e3, zDim=10,
                          vars={"h": 1, "u": 2, "v": 3})
    r, l, z = G.createGrid(gp).getGridpoints().T
    ic = np.zeros((r.size, 3))
    ic[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8) * np.cos(z / 400.0)
    ic[:, 1] = 5 * np.cos(l) * (1 + 0.2 * np.sin(z / 300.0))
    ic[:, 2] = -5 * np.sin(l) * np.exp(-z / 900.0)
    cases["LinearAdvectionRLZ"] = dict(gp=gp, eq="LinearAdvectionRLZ", prm={"K": 100.0}, ts=50.0, n=3, ic=ic, tiles=(1, 2))
    cases["LinearAdvectionRLZ_z16_fused"] = fused_advection_case(6)
    # 64 levels: k_inv_z_advection_bulk<true> on the blocked SZ layout (rings of 8 ... 64 points: whole, partial and empty 16-point blocks)
    cases["LinearAdvectionRLZ_z64_blocked"] = fused_advection_case(6, zDim=64)
    gp = G.GridParameters(geometry="RZ", xmin=0, xmax=1e5, num_cells=12, zmin=0, zmax=1e4, zDim=12,
                          vars={"h": 1, "u": 2, "x": 3, "w": 4})
    r, z = G.createGrid(gp).getGridpoints().T
    ic = np.zeros((r.size, 4))
    ic[:, 0] = np.exp(-((r - 5e4) ** 2 / 4e8 + (z - 5e3) ** 2 / 4e6))
    ic[:, 1] = 5.0 * np.cos(z / 4e3)
    ic[:, 2] = np.sin(r / 2e4) * np.cos(z / 3e3)
    ic[:, 3] = 0.5 * np.sin(r / 3e4 + z / 5e3)
    cases["LinearAdvectionRZ"] = dict(gp=gp, eq="LinearAdvectionRZ", prm={"K": 10.0}, ts=20.0, n=3, ic=ic, tiles=(1, 3))
    names = ["h", "u", "v", "ub", "vb", "wb"]
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=2e5, num_cells=6, zmin=0, zmax=2e3, zDim=10,
                          vars={n: i + 1 for i, n in enumerate(names)},
                          BCL={"h": spl.R1T1, "u": spl.R1T0, "v": spl.R1T0, "ub": spl.R1T0, "vb": spl.R1T0, "wb": spl.R1T1})
    r, l, z = G.createGrid(gp).getGridpoints().T
    Rmax, V0 = 5e4, 30.0 / 5e4
    vbar = np.where(r < Rmax, V0 * r, Rmax * Rmax * V0 / r)
    ic = np.zeros((r.size, 6))
    ic[:, 0] = 50 * np.exp(-(r / 1e5) ** 2) * (1 + 0.1 * np.cos(l))
    ic[:, 1] = 0.05 * vbar * np.sin(l)
    ic[:, 2] = vbar
    ic[:, 3] = -0.2 * vbar * np.exp(-z / 500)
    ic[:, 4] = vbar * (1 - np.exp(-(z + 50) / 300))
    gp16 = G.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=3, zmin=0, zmax=2e3, zDim=16,
                            vars={n: i + 1 for i, n in enumerate(names)},
                            BCL={"h": spl.R1T1, "u": spl.R1T0, "v": spl.R1T0, "ub": spl.R1T0, "vb": spl.R1T0, "wb": spl.R1T1})
    r16, l16, z16 = G.createGrid(gp16).getGridpoints().T
    vbar16 = np.where(r16 < Rmax, V0 * r16, Rmax * Rmax * V0 / r16)
    ic16 = np.zeros((r16.size, 6))
    ic16[:, 0] = 50 * np.exp(-(r16 / 1e5) ** 2) * (1 + 0.1 * np.cos(l16))
    ic16[:, 1] = 0.05 * vbar16 * np.sin(l16)
    ic16[:, 2] = vbar16
    ic16[:, 3] = -0.2 * vbar16 * np.exp(-z16 / 500)
    ic16[:, 4] = vbar16 * (1 - np.exp(-(z16 + 50) / 300))
    # 16 levels: the tensor-core (DMMA) column operators of k_heightresolved_bl2
    cases["Oneway_ShallowWater_HeightResolvedBL_z16"] = dict(
        gp=gp16, eq="Oneway_ShallowWater_HeightResolvedBL",
        prm=dict(g=9.81, Kh=1500.0, Cd=2.4e-3, Hfree=2000.0, f=5e-5, Um=3.0, Vm=-2.0), ts=2.0, n=2, ic=ic16, tiles=(1,))
    cases["Oneway_ShallowWater_HeightResolvedBL"] = dict(
        gp=gp, eq="Oneway_ShallowWater_HeightResolvedBL",
        prm=dict(g=9.81, Kh=1500.0, Cd=2.4e-3, Hfree=2000.0, f=5e-5, Um=3.0, Vm=-2.0), ts=2.0, n=3, ic=ic, tiles=(1, 2))
    gp = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=8, zmin=0, zmax=1e4, zDim=16,
                          vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5})
    zc = ch.mish_points(ch.ChebyshevParameters(0, 1e4, 16, 11))
    T = 280.0
    rho = (1000e2 / (T * M.Rd)) * np.exp(-M.gravity * zc / (M.Rd * T))
    xibar = np.log(rho / M.rho_d0)
    sbar = M.Cvd * np.log(T / M.T_0) - M.Rd * np.log(rho / M.rho_d0)
    mubar = 2e-3 * np.exp(-zc / 3e3) - 5e-4   # moist below, dry (mu<0) aloft: both ahyp branches
    ref = M.exact_reference_state_from_profiles(gp, sbar, xibar, mubar, mubar)
    x, z = G.createGrid(gp).getGridpoints().T
    ic = np.zeros((x.size, 5))
    ic[:, 0] = 2.0 * np.exp(-((x) ** 2 + (z - 3e3) ** 2) / 2e3 ** 2)
    ic[:, 2] = 1e-4 * np.exp(-((x - 2e3) ** 2 + (z - 2e3) ** 2) / 2e3 ** 2)
    ic[:, 3] = 1.0 * np.sin(z / 2e3)
    # moist test sets (src/testModels.jl:217-586): a hydrostatic moist sounding as the reference state, a warm moist bubble,
    # cloud and rain blobs, and a supersaturation field whose sign changes along the bottom level.  Of the two whole-column
    # (lexicographic) selections of condensation_adjustment, max(-q_c, q_cond) goes both ways across the columns;
    # min(q_v, q_cond) always picks q_cond here (its other outcome sets q_cond = q_v in the whole column, which the
    # reference's own formulas amplify to overflow within two steps; tests/test_oracle_golden.py covers the selection rule)
    gpm = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=8, zmin=0, zmax=1e4, zDim=16,
                           vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_c": 6, "mu_r": 7, "qss": 8})
    refm, icm = moist_case(gpm, rain=True)
    cases["rainfall_test_explicit"] = dict(gp=gpm, eq="rainfall_test", prm={"K": 50.0}, ts=0.05, n=3, ic=icm, tiles=(1,), ref=refm)
    cases["rainfall_test_semiimplicit"] = dict(gp=gpm, eq="rainfall_test", prm={"K": 50.0}, ts=0.5, n=3, ic=icm, tiles=(1, 2),
                                               ref=refm, opts={"semiimplicit": True, "exact_reference_state": True})
    gpb = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=8, zmin=0, zmax=1e4, zDim=16,
                           vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_c": 6, "qss": 7, "mu_r": 8})
    refb, icb = moist_case(gpb, rain=False)
    cases["BF02_test_semiimplicit"] = dict(gp=gpb, eq="BF02_test", prm={"K": 50.0}, ts=0.5, n=3, ic=icb, tiles=(1, 2),
                                           ref=refb, opts={"semiimplicit": True, "exact_reference_state": True})
    cases["Euler_test_explicit"] = dict(gp=gp, eq="Euler_test", prm={"K": 50.0}, ts=0.05, n=3, ic=ic, tiles=(1,), ref=ref)
    cases["Euler_test_semiimplicit"] = dict(gp=gp, eq="Euler_test", prm={"K": 50.0}, ts=1.0, n=4, ic=ic, tiles=(1, 2),
                                            ref=ref, opts={"semiimplicit": True, "exact_reference_state": True})
    return cases


def moist_case(gp, rain: bool):
    """Reference state + initial condition for BF02_test / rainfall_test on `gp` (8 variables)."""
    zc = ch.mish_points(ch.ChebyshevParameters(gp.zmin, gp.zmax, gp.zDim, gp.b_zDim))
    Tk = 300.0 - 6.5e-3 * zc
    p = 1000.0 * (Tk / 300.0) ** (M.gravity / (M.Rd * 6.5e-3))
    q_v = 0.95 * float(M.q_sat_liquid(np.float64(300.0), np.float64(1000.0))) * np.exp(-zc / 2500.0)   # 95 % RH at the surface
    e = M.vapor_pressure(p, q_v)
    rho_d = 100.0 * (p - e) / (M.Rd * Tk)
    sbar = M.entropy(Tk, rho_d, q_v)
    xibar = np.log(rho_d / M.rho_d0)
    mubar = M.bhyp(q_v)
    mu_lbar = M.bhyp(1e-4 * np.exp(-((zc - 3e3) / 1e3) ** 2))
    ref = M.exact_reference_state_from_profiles(gp, sbar, xibar, mubar, mu_lbar)
    x, z = G.createGrid(gp).getGridpoints().T
    v = {k: i - 1 for k, i in gp.vars.items()}
    ic = np.zeros((x.size, 8))
    ic[:, v["s"]] = 2.0 * np.exp(-((x) ** 2 + (z - 3e3) ** 2) / 2e3 ** 2)
    # a moist and a dry blob near the surface: max(-q_c, q_cond) of condensation_adjustment picks -q_c in the dry columns
    ic[:, v["mu"]] = 2e-3 * np.exp(-((x - 4e3) / 3e3) ** 2 - (z / 3e3) ** 2) - 3.5e-3 * np.exp(-((x + 5e3) / 2.5e3) ** 2 - (z / 2e3) ** 2)
    ic[:, v["u"]] = 2.0 * np.sin(z / 2e3)
    ic[:, v["w"]] = 0.5 * np.exp(-((x + 1e3) ** 2 + (z - 4e3) ** 2) / 2.5e3 ** 2)
    # haze everywhere (mu_c, mu_r > 0: where they go negative the reference's dmudq = 1 + |mu| / q0 makes its own tendencies
    # explosive), q_c above and below the 1 g/kg autoconversion threshold
    ic[:, v["mu_c"]] = 8e-4 + 1.5e-3 * np.exp(-((x - 1e3) ** 2 + (z - 3.5e3) ** 2) / 3e3 ** 2)
    ic[:, v["mu_r"]] = 2e-4 + 1.0e-3 * np.exp(-((x + 3e3) ** 2 + (z - 2.5e3) ** 2) / 3e3 ** 2)
    ic[:, v["qss"]] = 2e-4 * np.sin(x / 2.5e3) * np.exp(-z / 4e3)
    return ref, ic


def benchmarked_shape_cases():
    """N-step cases at the shapes bench.py and profiles/microbench time: 64 levels select other
    kernel instantiations than the 10/16-level cases above (k_heightresolved_bl2 runs (64, 8)-thread blocks with
    [zDim/8][zDim/4][32] fragments, k_inv_z_advection<16> its 64-level parity split, k_euler_test its 64 x 64 composite
    column operators), and 100 radial cells is the reference's own production grid (C2).  GPU tests only: the oracle
    needs 2-8 s for each."""
    cases = {}
    # C2, /root/reference/models/cha_bell2024/Oneway_ShallowWater_Slab.jl:1-40 (100 cells, 181,800 points, 6 variables)
    gp, ic, prm = _slab_case(100)
    cases["Oneway_ShallowWater_Slab_C2_100cells"] = dict(gp=gp, eq="Oneway_ShallowWater_Slab", prm=prm, ts=3.0, n=4, ic=ic,
                                                         tiles=(1, 2))
    # LinearAdvectionRLZ through the fused Chebyshev synthesis + tendency + AB3 kernel at the C4 level count
    gpf = G.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=24, zmin=0, zmax=1e3, zDim=64,
                           vars={"h": 1, "u": 2, "v": 3})
    r, l, z = G.createGrid(gpf).getGridpoints().T
    icf = np.zeros((r.size, 3))
    icf[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8) * np.cos(z / 400.0)
    icf[:, 1] = 5 * np.cos(l) * (1 + 0.2 * np.sin(z / 300.0))
    icf[:, 2] = -5 * np.sin(l) * np.exp(-z / 900.0)
    cases["LinearAdvectionRLZ_z64_24cells_fused"] = dict(gp=gpf, eq="LinearAdvectionRLZ", prm={"K": 100.0}, ts=50.0, n=4,
                                                         ic=icf, tiles=(1, 2))
    # the TC boundary-layer set (bench.py `tcbl`) at 64 levels
    names = ["h", "u", "v", "ub", "vb", "wb"]
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=2e5, num_cells=8, zmin=0, zmax=2e3, zDim=64,
                          vars={n: i + 1 for i, n in enumerate(names)},
                          BCL={"h": spl.R1T1, "u": spl.R1T0, "v": spl.R1T0, "ub": spl.R1T0, "vb": spl.R1T0, "wb": spl.R1T1})
    r, l, z = G.createGrid(gp).getGridpoints().T
    Rmax, V0 = 5e4, 30.0 / 5e4
    vbar = np.where(r < Rmax, V0 * r, Rmax * Rmax * V0 / r)
    ic = np.zeros((r.size, 6))
    ic[:, 0] = 50 * np.exp(-(r / 1e5) ** 2) * (1 + 0.1 * np.cos(l))
    ic[:, 1] = 0.05 * vbar * np.sin(l)
    ic[:, 2] = vbar
    ic[:, 3] = -0.2 * vbar * np.exp(-z / 500)
    ic[:, 4] = vbar * (1 - np.exp(-(z + 50) / 300))
    cases["Oneway_ShallowWater_HeightResolvedBL_z64"] = dict(
        gp=gp, eq="Oneway_ShallowWater_HeightResolvedBL",
        prm=dict(g=9.81, Kh=1500.0, Cd=2.4e-3, Hfree=2000.0, f=5e-5, Um=3.0, Vm=-2.0), ts=2.0, n=3, ic=ic, tiles=(1, 2))
    # C3: RZ 334 cells x 64 levels, Euler_test with the semi-implicit adjustment (/root/reference/src/testModels.jl:100-215,
    # src/semiimplicit.jl:521-597)
    gp = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=334, zmin=0, zmax=1e4, zDim=64,
                          vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5})
    zc = ch.mish_points(ch.ChebyshevParameters(0, 1e4, 64, 43))
    T = 280.0
    rho = (1000e2 / (T * M.Rd)) * np.exp(-M.gravity * zc / (M.Rd * T))
    xibar = np.log(rho / M.rho_d0)
    sbar = M.Cvd * np.log(T / M.T_0) - M.Rd * np.log(rho / M.rho_d0)
    mubar = 2e-3 * np.exp(-zc / 3e3) - 5e-4
    ref = M.exact_reference_state_from_profiles(gp, sbar, xibar, mubar, mubar)
    x, z = G.createGrid(gp).getGridpoints().T
    ic = np.zeros((x.size, 5))
    ic[:, 0] = 2.0 * np.exp(-((x) ** 2 + (z - 3e3) ** 2) / 2e3 ** 2)
    ic[:, 2] = 1e-4 * np.exp(-((x - 2e3) ** 2 + (z - 2e3) ** 2) / 2e3 ** 2)
    ic[:, 3] = 1.0 * np.sin(z / 2e3)
    # the moist sets at the C3 level count (64 levels: four warps per column in k_moist_test, 64 x 64 column operators)
    gpm = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=100, zmin=0, zmax=1e4, zDim=64,
                           vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_c": 6, "mu_r": 7, "qss": 8})
    refm, icm = moist_case(gpm, rain=True)
    cases["rainfall_test_semiimplicit_100cells_z64"] = dict(gp=gpm, eq="rainfall_test", prm={"K": 50.0}, ts=0.5, n=3, ic=icm,
                                                            tiles=(1, 2), ref=refm,
                                                            opts={"semiimplicit": True, "exact_reference_state": True})
    gpb = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=100, zmin=0, zmax=1e4, zDim=64,
                           vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_c": 6, "qss": 7, "mu_r": 8})
    refb, icb = moist_case(gpb, rain=False)
    cases["BF02_test_semiimplicit_100cells_z64"] = dict(gp=gpb, eq="BF02_test", prm={"K": 50.0}, ts=0.5, n=3, ic=icb, tiles=(1, 2),
                                                        ref=refb, opts={"semiimplicit": True, "exact_reference_state": True})
    cases["Euler_test_semiimplicit_C3_334cells_z64"] = dict(gp=gp, eq="Euler_test", prm={"K": 50.0}, ts=1.0, n=4, ic=ic,
                                                            tiles=(1, 2), ref=ref,
                                                            opts={"semiimplicit": True, "exact_reference_state": True})
    return cases


def run_oracle(case, ntiles=1):
    opts = case.get("opts", {"semiimplicit": False})
    mp = M.ModelParameters(ts=case["ts"], integration_time=case["ts"] * case["n"], equation_set=case["eq"],
                           grid_params=case["gp"], physical_params=case["prm"], options=opts)
    run = M.ModelRun(mp, ntiles, case["ic"], case.get("ref"))
    run.run(case["n"])
    return run


def pkg_model(case, ntiles, lib, **kw):
    opts = case.get("opts", {"semiimplicit": False})
    mp = S.ModelParameters(ts=case["ts"], integration_time=case["ts"] * case["n"], equation_set=case["eq"],
                           grid_params=to_pkg(case["gp"]), physical_params=case["prm"], options=opts)
    ref = case.get("ref")
    sref = S.ReferenceState(ref.sbar, ref.xibar, ref.mubar, ref.mu_lbar, ref.Pxi_bar) if ref is not None else None
    return S.Model(mp, num_tiles=ntiles, ref_state=sref, lib=lib, **kw)


def check_model(case, lib, **kw):
    """N steps on `tiles` tiles: final tile state (var_np1, expdot history) vs the oracle."""
    errs = []
    for nt in case["tiles"]:
        orun = run_oracle(case, nt)
        m = pkg_model(case, nt, lib, **kw)
        m.initialize(case["ic"])
        m.run(case["n"])
        for i, mt in enumerate(orun.mtiles):
            errs.append(rel_err(m.state(i, "var_np1"), mt.var_np1))
            errs.append(rel_err(m.state(i, "expdot_nm1"), mt.expdot_nm1))
        out = m.output()
        oout = orun.output_patch()
        errs.append(rel_err(out[:, :, 0], oout[:, :, 0]))   # state
        errs.append(rel_err(out[:, :, 1], oout[:, :, 1]))   # radial derivative
        m.close()
    return max(errs)


def check_host_pipeline(case, lib, ntiles=1, nsteps=5, pinned=False, **kw):
    """Host-driven stepping: the pipelined cycle_host (sb_model_stage_in/out: copy streams, double staging buffers)
    must leave, for every step, exactly the bytes that set_state -> cycle -> get_state leaves.  Every step gets a
    different input and its own output buffer, so a staging buffer reused too early or a copy ordered wrongly shows."""
    def alloc(shape):
        if not pinned:
            return np.zeros(shape, order="F")
        import torch
        t = torch.zeros((shape[1], shape[0]), dtype=torch.float64, pin_memory=True)
        keep.append(t)
        return t.numpy().T
    keep = []
    a = pkg_model(case, ntiles, lib, **kw)
    a.initialize(case["ic"])
    a.run(1)
    nmodel, ntiles = ntiles, len(a.tiles)   # tiles of the patch / tiles this process owns (distributed: its share)
    base = [a.state(i, "var_np1") for i in range(ntiles)]
    ins = [[alloc(b.shape) for b in base] for _ in range(nsteps)]
    for s_, row in enumerate(ins):
        for x, b in zip(row, base):
            x[...] = b * (1.0 + 0.05 * s_)
    ref = []
    for s_ in range(nsteps):
        for i in range(ntiles):
            a.set_state(i, ins[s_][i])
        a.cycle()
        ref.append([a.state(i, "var_np1") for i in range(ntiles)])
    hist_a = [a.state(i, "expdot_nm1") for i in range(ntiles)]
    a.close()
    b = pkg_model(case, nmodel, lib, **kw)
    b.initialize(case["ic"])
    b.run(1)
    outs = [[alloc(x.shape) for x in base] for _ in range(nsteps)]
    for s_ in range(nsteps):
        b.cycle_host(ins[s_], outs[s_])
    b.drain()
    for s_ in range(nsteps):
        for i in range(ntiles):
            assert np.array_equal(outs[s_][i], ref[s_][i]), f"step {s_} tile {i}: pipelined host cycle differs"
    for i in range(ntiles):
        assert np.array_equal(b.state(i, "expdot_nm1"), hist_a[i])
        assert np.array_equal(b.state(i, "var_np1"), ref[-1][i])
    b.sync()
    b.close()


def check_needed_slots(case, lib, ntiles=None, **kw):
    """In-step tileTransform! producing only the slots the equation-set kernel reads (every other slot NaN) must leave
    exactly the state that producing all D slots leaves (src/semiimplicit.jl:305-314): bit-identical, no NaN."""
    nt = ntiles or case["tiles"][-1]
    states = {}
    for mode in ("all", "needed-poisoned", "fused"):
        m = pkg_model(case, nt, lib, **kw)
        m.set_k3_slots(mode)
        m.initialize(case["ic"])
        m.run(case["n"])
        states[mode] = [(m.state(i, "var_np1").copy(), m.state(i, "expdot_nm1").copy()) for i in range(nt)]
        m.close()
    for (a0, a1), (b0, b1), (c0, c1) in zip(states["all"], states["needed-poisoned"], states["fused"]):
        assert np.isfinite(b0).all() and np.isfinite(b1).all(), "a slot outside the declared mask was read"
        assert np.array_equal(a0, b0) and np.array_equal(a1, b1), "needed-slots state differs from the all-slots state"
        assert np.array_equal(a0, c0) and np.array_equal(a1, c1), (
            f"fused K3+K4 state differs from the all-slots state: max |d var_np1| {np.abs(a0 - c0).max():.3e} of "
            f"{np.abs(a0).max():.3e}, max |d expdot| {np.abs(a1 - c1).max():.3e} of {np.abs(a1).max():.3e}")


def check_overlapped_step(case, lib, ntiles, monkeypatch, batches=(1, 3, 5), **kw):
    """The overlapped step (two streams, ring batches, dynamic FFT work shares, SB_OVERLAP=1 forces it on small grids) must
    leave exactly the bits of the sequential step, for any number of ring batches."""
    monkeypatch.setenv("SB_OVERLAP", "0")
    m = pkg_model(case, ntiles, lib, **kw)
    m.initialize(case["ic"])
    m.run(case["n"])
    ref = [(m.state(i, "var_np1").copy(), m.state(i, "expdot_nm1").copy(), m.state(i, "expdot_nm2").copy()) for i in range(ntiles)]
    m.close()
    for nb in batches:
        monkeypatch.setenv("SB_OVERLAP", "1")
        monkeypatch.setenv("SB_OVERLAP_BATCHES", str(nb))
        m = pkg_model(case, ntiles, lib, **kw)
        m.initialize(case["ic"])
        m.run(case["n"])
        for i in range(ntiles):
            got = (m.state(i, "var_np1"), m.state(i, "expdot_nm1"), m.state(i, "expdot_nm2"))
            for a, b, name in zip(got, ref[i], ("var_np1", "expdot_nm1", "expdot_nm2")):
                assert np.array_equal(a, b), f"{nb} batches, tile {i}, {name}: overlapped step differs from the sequential step"
        m.close()
