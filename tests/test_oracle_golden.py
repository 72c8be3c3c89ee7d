"""Pins for the oracle itself (CPU): the reference's only known answer (a sanity BAND, see
oracle/__init__.py -- parity unpinned), analytic exactness, and structural identities."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import chebyshev as ch
from oracle import fourier
from oracle import grids as G
from oracle import model as M
from oracle import splines as spl

GOLD = json.loads((Path(__file__).parent / "golden" / "linear_advection_notebook.json").read_text())


def _c1(num_tiles):
    gp = G.GridParameters(geometry="R", xmin=-50.0, xmax=50.0, num_cells=100, BCL={"u": spl.PERIODIC},
                          BCR={"u": spl.PERIODIC}, vars={"u": 1})
    mp = M.ModelParameters(ts=0.05, integration_time=100.0, output_interval=50.0, equation_set="LinearAdvection1D",
                           grid_params=gp, physical_params={"c_0": 1.0, "K": 0.0})
    x = G.createGrid(gp).getGridpoints()
    run = M.ModelRun(mp, num_tiles, np.exp(-(x / 20.0) ** 2)[:, None])
    u0 = run.output_patch()[:, 0, 0].copy()
    run.run(2000)
    return x, u0, run.output_patch()[:, 0, 0].copy()


def test_linear_advection_notebook_band():
    """notebooks/LinearAdvection_example.ipynb:91-116 (grid), :229-254 (u at t=100), :307 (L2)."""
    x, u0, uf = _c1(2)
    assert np.allclose(x[:13], GOLD["gridpoints_first13"], rtol=0, atol=1e-13)
    assert np.allclose(x[-12:], GOLD["gridpoints_last12"], rtol=0, atol=1e-13)
    gold = np.array(GOLD["final_u_first13"] + GOLD["final_u_last12"])
    got = np.concatenate([uf[:13], uf[-12:]])
    assert np.abs(got / gold - 1).max() < 5e-3          # 0.5 % band
    l2 = np.sqrt(((u0 - uf) ** 2).sum())
    assert abs(l2 / GOLD["l2_norm"] - 1) < 0.05        # 5 % band


def test_grid_parameter_derivations():
    """printed GridParameters dump, notebooks/LinearAdvection_example.ipynb:43."""
    gp = G.GridParameters(geometry="R", xmin=-50.0, xmax=50.0, num_cells=100)
    assert (gp.rDim, gp.b_rDim, gp.l_q, gp.spectralIndexL, gp.spectralIndexR, gp.patchOffsetL, gp.patchOffsetR) == \
        (300, 103, 2.0, 1, 103, 0, 300)
    assert G.GridParameters(geometry="RLZ", num_cells=334, xmax=1.0, zDim=64).b_zDim == 43


def test_north_star_dimensions():
    """N = 128,897,280, S = 29,054,455, 86,215 spline columns."""
    nc, zDim, bz = 334, 64, 43
    ri = np.arange(1, 3 * nc + 1)
    h = int((4 + 4 * ri).sum())
    assert h == 2_014_020 and h * zDim == 128_897_280
    assert bz * (nc + 3) * (1 + 2 * 3 * nc) == 29_054_455 and bz * (1 + 2 * 3 * nc) == 86_215


def test_spline_reproduces_cubics_without_filter():
    sp = spl.Spline1D(spl.SplineParameters(xmin=0.0, xmax=10.0, num_cells=12, l_q=1e-3))
    x = sp.mishPoints
    f = 0.3 * x ** 3 - 2 * x ** 2 + x - 4
    a = sp.SAtransform(sp.SBtransform(f))
    assert np.abs(sp.SItransform(a) - f).max() < 1e-10 * np.abs(f).max()
    assert np.abs(sp.SItransform(a, 1) - (0.9 * x ** 2 - 4 * x + 1)).max() < 1e-9 * np.abs(f).max()
    assert np.abs(sp.SItransform(a, 2) - (1.8 * x - 4)).max() < 1e-8 * np.abs(f).max()


@pytest.mark.parametrize("bcl,bcr", [("R1T0", "R1T1"), ("R1T2", "R2T10"), ("R2T20", "R3"), ("R3", "R1T0")])
def test_spline_boundary_conditions_hold(bcl, bcr):
    p = spl.SplineParameters(xmin=0.0, xmax=5.0, num_cells=9, BCL=spl.BC_BY_NAME[bcl], BCR=spl.BC_BY_NAME[bcr])
    sp = spl.Spline1D(p)
    rng = np.random.default_rng(1)
    a = sp.SAtransform(sp.SBtransform(rng.standard_normal(p.mishDim)))
    ends = np.array([p.xmin, p.xmax])
    val = [spl.basis_matrix(p, ends, d) @ a for d in range(3)]
    want = {"R1T0": [0], "R1T1": [1], "R1T2": [2], "R2T10": [0, 1], "R2T20": [0, 2], "R3": [0, 1, 2]}
    for side, bc in ((0, bcl), (1, bcr)):
        for d in want[bc]:
            assert abs(val[d][side]) < 1e-10 * max(1.0, np.abs(a).max() * 10 ** d)


def test_ring_fft_exact_for_retained_wavenumbers():
    for ri in (1, 2, 7, 30, 101):
        lam = fourier.ring_lambdas(ri)
        k = ri
        u = 1.5 + np.cos(k * lam + 0.4) - 0.5 * np.sin(max(k - 1, 1) * lam)
        c = fourier.ring_forward(u, ri)
        assert np.abs(fourier.ring_inverse(c, ri) - u).max() < 1e-12
        du = -k * np.sin(k * lam + 0.4) - 0.5 * max(k - 1, 1) * np.cos(max(k - 1, 1) * lam)
        assert np.abs(fourier.ring_inverse(c, ri, 1) - du).max() < 1e-11 * k


def test_chebyshev_exact_for_polynomials_and_integral():
    cp = ch.ChebyshevParameters(zmin=0.0, zmax=2.0, zDim=16, bDim=11)
    col = ch.Chebyshev1D(cp)
    z = col.mishPoints
    f = z ** 5 - 3 * z ** 2 + 1
    a = col.CAtransform(col.CBtransform(f))
    assert np.abs(col.CItransform(a) - f).max() < 1e-12 * np.abs(f).max()
    assert np.abs(col.CIxtransform(a) - (5 * z ** 4 - 6 * z)).max() < 1e-11 * np.abs(f).max() * 10
    assert np.abs(col.CIxxtransform(a) - (20 * z ** 3 - 6)).max() < 1e-10 * np.abs(f).max() * 100
    assert np.abs(col.CIInttransform(a) - (z ** 6 / 6 - z ** 3 + z)).max() < 1e-12 * np.abs(f).max()
    assert z[0] == 0.0 and abs(z[-1] - 2.0) < 1e-15    # level 1 is the bottom


def test_helmholtz_matrix_consistency():
    """dct_matrix is the synthesis matrix of CItransform (rows = levels bottom->top)."""
    cp = ch.ChebyshevParameters(zmin=0.0, zmax=1e4, zDim=12, bDim=12)
    col = ch.Chebyshev1D(cp)
    rng = np.random.default_rng(2)
    a = rng.standard_normal(12)
    import scipy.fft as sfft
    assert np.abs(ch.dct_matrix(12) @ a - sfft.dct(a, type=1)).max() < 1e-12     # FFTW REDFT00
    assert np.abs(col.CItransform(a) - ch.dct_matrix(12) @ a).max() == 0.0


def test_tile_sum_identity_and_halo_maps():
    """sum over tiles of B_tile (own block + 3-coefficient halo) == B_patch."""
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=10, num_cells=12, zmin=0, zmax=1, zDim=8, vars={"a": 1, "b": 2})
    patch = G.createGrid(gp)
    rng = np.random.default_rng(3)
    patch.physical[:, :, 0] = rng.standard_normal((patch.N, 2))
    patch.spectralTransform()
    tp = G.calcTileSizes(patch, 3)
    shared = np.zeros_like(patch.spectral)
    start = 0
    for t in range(3):
        tile = G.createGrid(G.tile_params(patch, tp, t))
        n = tile.N
        tile.physical[:, :, 0] = patch.physical[start:start + n, :, 0]
        start += n
        tile.spectralTransform()
        prow, trow = G.calcPatchMap(patch, tile)
        hrow, htrow = G.calcHaloMap(patch, tile)
        assert len(hrow) == 3 * tile.ncolp * tile.b_zDim
        shared[prow] += tile.spectral[trow]
        shared[hrow] += tile.spectral[htrow]
    assert start == patch.N
    assert np.abs(shared - patch.spectral).max() < 1e-13 * np.abs(patch.spectral).max()


def test_moist_closure_identities_and_whole_column_selection():
    """The moist closure of BF02_test / rainfall_test (src/thermodynamics.jl, src/microphysics.jl) has no golden vector in
    the reference tree; what can be pinned is its internal consistency and the selection rule of the un-dotted
    min / max in condensation_adjustment (src/microphysics.jl:185-187: Julia's generic min(x, y) = ifelse(isless(y, x), y, x),
    isless lexicographic for vectors, isequal / isless for the elements: NaN above everything, -0.0 < 0.0)."""
    Tk, p = np.float64(288.0), np.float64(900.0)
    h = 1e-4
    d = (M.sat_pressure_liquid_buck(Tk + h, p) - M.sat_pressure_liquid_buck(Tk - h, p)) / (2 * h)
    assert abs(M.sat_pressure_liquid_buck_dT(Tk, p) / d - 1) < 1e-8                  # analytic dT == finite difference
    es = M.sat_pressure_liquid_buck(Tk, p)
    assert abs(M.vapor_pressure(p, M.q_sat_liquid(Tk, p)) / es - 1) < 1e-14         # q_sat <-> vapour pressure inverse pair
    q = np.array([0.0, 1e-9, 1e-4, 2e-2])
    assert np.allclose(M.ahyp(M.bhyp(q)), q, rtol=1e-12, atol=1e-20)                # hyperbolic water variable round trip
    rho_d, q_v = np.float64(1.05), np.float64(0.012)
    s = M.entropy(Tk, rho_d, q_v)
    assert abs(M.temperature(s, rho_d, q_v) / Tk - 1) < 1e-13                        # entropy <-> temperature
    assert M.sedimentation(np.float64(1e-3), rho_d, Tk) == 0.0                       # the clamp leaves no fall speed (quirk)
    A = np.array([[1.0, 5.0, 0.0], [1.0, 2.0, 9.0], [0.0, 1.0, 1.0], [-0.0, 7.0, 7.0], [np.nan, 0.0, 0.0], [2.0, 2.0, 2.0]])
    B = np.array([[1.0, 6.0, -9.0], [1.0, 2.0, 3.0], [-0.0, 9.0, 9.0], [0.0, 0.0, 0.0], [5.0, 9.0, 9.0], [2.0, 2.0, 2.0]])
    #            first unequal pair decides | later pair | 0.0 vs -0.0      | -0.0 < 0.0      | NaN is largest  | equal
    assert list(M._lex_less(A, B)) == [True, False, False, True, False, False]
    assert list(M._lex_less(B, A)) == [False, True, True, False, True, False]


def test_moist_sets_fail_like_the_reference_without_the_named_variables():
    """condensation_adjustment looks up "mu_c", "mu_r", "qss" by name: BF02_test with a variable list that calls column 6
    "mu_l" dies with a KeyError on its first step in the reference (src/microphysics.jl:158-165)."""
    gp = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=4, zmin=0, zmax=1e4, zDim=8,
                          vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_l": 6, "qss": 7})
    from helpers import moist_case
    gp8 = G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=4, zmin=0, zmax=1e4, zDim=8,
                           vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5, "mu_c": 6, "qss": 7, "mu_r": 8})
    ref, ic = moist_case(gp8, rain=False)
    mp = M.ModelParameters(ts=0.1, equation_set="BF02_test", grid_params=gp, physical_params={"K": 1.0})
    run = M.ModelRun(mp, 1, ic[:, :7], ref)
    with pytest.raises(KeyError, match="mu_c"):
        run.step()
