"""GPU parity tests (-m gpu): the product library through the C ABI vs the oracle, on the same
seeded inputs; plus size-independent properties at sizes the oracle cannot cover quickly."""
import numpy as np
import pytest

import scythe_jl_b200 as S
from helpers import (STATE_TOL, TRANSFORM_TOL, benchmarked_shape_cases, check_model, check_needed_slots, check_transforms, model_cases, pkg_model,
                     rel_err, slot_errs, to_pkg, transform_cases)
from oracle import grids as G
from oracle import splines as spl

pytestmark = pytest.mark.gpu
T_CASES = transform_cases()
M_CASES = model_cases()
B_CASES = benchmarked_shape_cases()


def test_native_library_is_the_product_build(gpu_lib):
    assert b"sm_100a" in gpu_lib.sb_version()
    assert gpu_lib.path.name == "libscythe_b200.so"


@pytest.mark.parametrize("name", sorted(T_CASES))
def test_transforms_match_oracle(name, gpu_lib):
    eB, eP = check_transforms(T_CASES[name], gpu_lib)
    assert eB <= TRANSFORM_TOL, f"spectralTransform rel err {eB}"
    assert max(eP) <= TRANSFORM_TOL, f"gridTransform rel err per slot {eP}"


@pytest.mark.parametrize("name", sorted(M_CASES))
def test_timestep_matches_oracle(name, gpu_lib):
    assert check_model(M_CASES[name], gpu_lib) <= STATE_TOL


@pytest.mark.parametrize("name", sorted(B_CASES))
def test_timestep_matches_oracle_at_benchmarked_shapes(name, gpu_lib):
    """N-step state vs the oracle at the shapes that are timed: C2 (100 cells), C3 (334 cells x 64 levels, semi-implicit
    Euler_test), the TC boundary-layer set and the fused LinearAdvectionRLZ kernel at 64 levels."""
    assert check_model(B_CASES[name], gpu_lib) <= STATE_TOL


@pytest.mark.parametrize("name", sorted(M_CASES))
def test_needed_slots_state_is_bit_identical(name, gpu_lib):
    """Every equation set: K3 restricted to the (variable, slot) pairs its kernel reads, all other slots NaN,
    vs K3 producing all D slots (the reference's dataflow)."""
    for nt in M_CASES[name]["tiles"]:
        check_needed_slots(M_CASES[name], gpu_lib, ntiles=nt)


@pytest.mark.parametrize("geometry,num_cells,zDim", [("RL", 180, 0), ("RLZ", 40, 16), ("RLZ", 24, 64)])
def test_needed_slots_large_rings(geometry, num_cells, zDim, gpu_lib):
    """The same identity on rings long enough for the persistent Bluestein kernels (power-of-two and 3 x 2^a
    convolution lengths) and the DMMA Chebyshev synthesis, where the row / field masks are applied."""
    kw = dict(zmin=0, zmax=1e3, zDim=zDim) if zDim else {}
    gp = G.GridParameters(geometry=geometry, xmin=0, xmax=1e3 * num_cells, num_cells=num_cells, vars={"h": 1, "u": 2, "v": 3}, **kw)
    pts = G.createGrid(gp).getGridpoints()
    r, l = pts[:, 0], pts[:, 1]
    zf = np.cos(pts[:, 2] / 400.0) if zDim else 1.0
    ic = np.zeros((r.size, 3))
    ic[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8) * zf
    ic[:, 1] = 5 * np.cos(l)
    ic[:, 2] = -5 * np.sin(l) * zf
    case = dict(gp=gp, eq="LinearAdvection" + geometry, prm={"K": 100.0}, ts=20.0, n=2, ic=ic, tiles=(1,))
    check_needed_slots(case, gpu_lib, ntiles=1)
    check_needed_slots(case, gpu_lib, ntiles=2)


def test_cha_bell_rl_config_transforms(gpu_lib):
    """C2: RL, 100 cells, 6 variables with the Cha & Bell (2024) boundary conditions
    (/root/reference/models/cha_bell2024/Oneway_ShallowWater_Slab.jl:9-33), rings 8..1204."""
    names = ["h", "u", "v", "ub", "vb", "wb"]
    BCL = {"h": spl.R1T1, "u": spl.R1T0, "v": spl.R1T0, "ub": spl.R1T0, "vb": spl.R1T0, "wb": spl.R1T1}
    BCR = {"h": spl.R0, "u": spl.R1T1, "v": spl.R0, "ub": spl.R1T1, "vb": spl.R0, "wb": spl.R0}
    gp = G.GridParameters(geometry="RL", xmin=0, xmax=3e5, num_cells=100, BCL=BCL, BCR=BCR,
                          vars={n: i + 1 for i, n in enumerate(names)})
    eB, eP = check_transforms(gp, gpu_lib, seed=3)
    assert eB <= TRANSFORM_TOL and max(eP) <= TRANSFORM_TOL, (eB, eP)


def test_rlz_64_levels_transforms(gpu_lib):
    """RLZ with the north-star vertical resolution (64 levels -> 43 modes), 20 cells (rings 8..244)."""
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=6e4, num_cells=20, zmin=0, zmax=1.5e4, zDim=64,
                          BCL={"h": spl.R1T1, "u": spl.R1T0}, vars={"h": 1, "u": 2})
    eB, eP = check_transforms(gp, gpu_lib, seed=5)
    assert eB <= TRANSFORM_TOL and max(eP) <= TRANSFORM_TOL, (eB, eP)


def test_rz_north_star_config(gpu_lib):
    """C3: RZ 334 cells x 64 levels, 5 variables."""
    gp = G.GridParameters(geometry="RZ", xmin=0, xmax=1e6, num_cells=334, zmin=0, zmax=2e4, zDim=64,
                          vars={"s": 1, "xi": 2, "mu": 3, "u": 4, "w": 5})
    eB, eP = check_transforms(gp, gpu_lib, seed=7)
    assert eB <= TRANSFORM_TOL and max(eP) <= TRANSFORM_TOL, (eB, eP)


@pytest.mark.parametrize("sil,expect_L", [(172, 1536), (343, 3072), (690, 6144), (1000, 8192)])
def test_outer_tile_rings_composite_bluestein_lengths(sil, expect_L, gpu_lib):
    """Outer radial tiles of a multi-GPU patch (rings of up to 12,000 points): every convolution-length class,
    including the composite 3 x 2^a ones, against the oracle (expect_L documents the class the rings fall into)."""
    gp = G.GridParameters(geometry="RL", xmin=float(sil - 1), xmax=float(sil + 1), num_cells=2, vars={"h": 1, "u": 2},
                          spectralIndexL=sil)
    eB, eP = check_transforms(gp, gpu_lib, seed=sil)
    assert eB <= TRANSFORM_TOL, eB
    assert max(eP[:4]) <= TRANSFORM_TOL and eP[4] <= 1e-9, eP     # d2/dlambda2 of white noise: round-off x kDim^2
    gz = G.GridParameters(geometry="RLZ", xmin=float(sil - 1), xmax=float(sil), num_cells=1, zmin=0, zmax=1e4, zDim=16,
                          vars={"h": 1}, spectralIndexL=sil)
    eB, eP = check_transforms(gz, gpu_lib, seed=sil + 1)
    assert eB <= TRANSFORM_TOL and max(eP[:4]) <= TRANSFORM_TOL and max(eP[5:]) <= TRANSFORM_TOL and eP[4] <= 1e-9, (eB, eP)


def test_long_radial_columns(gpu_lib):
    """b_rDim = 1203 (the C5 weak-scaling patch is 948): streaming banded solve with a 38 KB Cholesky table."""
    gp = G.GridParameters(geometry="R", xmin=0, xmax=1e6, num_cells=1200,
                          BCL={"a": spl.R1T1, "b": spl.R1T0, "c": spl.R0}, BCR={"a": spl.R0, "b": spl.R2T10, "c": spl.R3},
                          vars={"a": 1, "b": 2, "c": 3})
    eB, eP = check_transforms(gp, gpu_lib, seed=17)
    assert eB <= TRANSFORM_TOL and max(eP) <= TRANSFORM_TOL, (eB, eP)


def test_plane_distributed_solve_matches_shared_array_scheme(gpu_lib):
    """exchange="columns" (one process owning every z-mode plane) == the reference's shared-array scheme."""
    case = dict(M_CASES["LinearAdvectionRLZ"])
    case["tiles"] = (2,)
    assert check_model(case, gpu_lib, exchange="columns") <= STATE_TOL
    case = dict(M_CASES["Oneway_ShallowWater_HeightResolvedBL"])
    case["tiles"] = (2,)
    assert check_model(case, gpu_lib, exchange="columns") <= STATE_TOL


def test_tiles_equal_single_tile_and_oracle_on_larger_rl(gpu_lib):
    """N-tile == 1-tile (overlap-add of the 3 seam coefficients) at a size with uneven tiles."""
    case = dict(M_CASES["Oneway_ShallowWater_Slab"])
    from helpers import _slab_case
    gp, ic, prm = _slab_case(24)
    case.update(gp=gp, ic=ic, prm=prm, n=5)
    outs = {}
    for nt in (1, 4, 8):
        m = pkg_model(case, nt, gpu_lib)
        m.initialize(ic)
        m.run(case["n"])
        outs[nt] = m.output()
        m.close()
    for nt in (4, 8):
        assert max(slot_errs(outs[nt], outs[1])) <= 1e-11
    from helpers import run_oracle
    oout = run_oracle(case, 1).output_patch()
    assert max(slot_errs(outs[1], oout)) <= STATE_TOL


def test_full_size_ring_ffts_pure_mode_property(gpu_lib):
    """Every ring size up to the north-star 4012-point ring (all 1002 Bluestein plans):
    for u = f(r) cos(k lambda + a) the transform pair must give  u_ll = -k^2 u  and  u_l = d/dlambda u
    at every point, and spectral energy only in wavenumber k, independent of the radial filter."""
    gp = S.GridParameters(geometry="RL", xmin=0, xmax=1e6, num_cells=334, vars={"u": 1})
    g = S.createGrid(gp, lib=gpu_lib)
    r, l = S.getGridpoints(g).T
    k = 3
    f = np.exp(-((r - 4e5) / 2e5) ** 2)
    g.physical[:, 0, 0] = f * np.cos(k * l + 0.3)
    S.spectralTransform(g)
    spec = g.spectral[:, 0].reshape(-1, g.b_rDim)      # [1+2kDim, b_rDim]
    keep = np.abs(spec).max()
    other = np.delete(spec, [2 * k - 1, 2 * k], axis=0)
    assert np.abs(other).max() <= 1e-13 * keep
    S.gridTransform(g)
    u, ul, ull = g.physical[:, 0, 0], g.physical[:, 0, 3], g.physical[:, 0, 4]
    ring_id = np.repeat(np.arange(g.rDim), 4 + 4 * (np.arange(g.rDim) + 1))
    sel = (ring_id + 1) >= k                           # rings that carry wavenumber k
    scale = np.abs(u).max()
    # round-off at wavenumber q is amplified by q^2 in the second derivative: floor ~ eps * kDim^2
    assert np.abs(ull + k * k * u)[sel].max() <= 20 * np.finfo(float).eps * g.kDim ** 2 * scale
    # u = A(r) cos(k l + 0.3)  ->  u_l^2 + k^2 u^2 = k^2 A^2 is constant on each ring
    amp2 = (ul ** 2 + (k * u) ** 2)
    mx = np.maximum.reduceat(amp2, np.r_[0, np.cumsum(4 + 4 * (np.arange(g.rDim) + 1))[:-1]])
    mn = np.minimum.reduceat(amp2, np.r_[0, np.cumsum(4 + 4 * (np.arange(g.rDim) + 1))[:-1]])
    rsel = np.arange(g.rDim) + 1 >= k
    assert ((mx - mn)[rsel]).max() <= 20 * np.finfo(float).eps * g.kDim * (k * scale) ** 2
    g.close()


def test_north_star_horizontal_grid_matches_oracle(gpu_lib):
    """Maximum ring sizes: the C4 horizontal grid (334 cells, rings 8..4012 points, every Bluestein
    plan) as an RL grid, directly against the oracle."""
    gp = G.GridParameters(geometry="RL", xmin=0, xmax=1e6, num_cells=334, BCL={"u": spl.R1T1}, vars={"u": 1})
    eB, eP = check_transforms(gp, gpu_lib, seed=13)
    assert eB <= TRANSFORM_TOL, eB
    # value, d/dr, d2/dr2, d/dlambda within 1e-12; d2/dlambda2 of white noise amplifies round-off by kDim^2
    assert max(eP[:4]) <= TRANSFORM_TOL and eP[4] <= 1e-10, eP


def test_c4_full_size_vertical_exactness(gpu_lib):
    """BASELINE.json's full C4 grid (334 cells, rings 8..4012 points, 64 levels, N = 128,897,280 points, one
    variable): u = p(z), a degree-5 polynomial, must come back from spectralTransform!/gridTransform! with
    u_z = p'(z), u_zz = p''(z) and zero radial / azimuthal derivatives at EVERY point -- constants pass the
    spline filter and wavenumber 0 passes every ring's Bluestein plan exactly (all convolution-length classes,
    the DMMA Chebyshev analysis/synthesis and the streaming spline solve at the size bench.py times)."""
    zmax = 2.0e4
    gp = S.GridParameters(geometry="RLZ", xmin=0, xmax=1e6, num_cells=334, zmin=0, zmax=zmax, zDim=64, vars={"u": 1})
    g = S.createGrid(gp, lib=gpu_lib)
    assert g.N == 128_897_280 and g.S == 29_054_455
    zl = 0.5 * zmax * (1.0 - np.cos(np.pi * np.arange(64) / 63))
    x = 2.0 * zl / zmax - 1.0                                   # [-1, 1]
    c = np.array([0.3, -1.1, 0.7, 0.45, -0.6, 0.25])            # p(x) = sum c_k x^k
    p0 = np.polyval(c[::-1], x)
    p1 = np.polyval(np.polyder(c[::-1]), x) * (2.0 / zmax)
    p2 = np.polyval(np.polyder(c[::-1], 2), x) * (2.0 / zmax) ** 2
    ncol = g.N // 64
    g.physical[:, 0, 0] = np.tile(p0, ncol)
    S.spectralTransform(g)
    S.gridTransform(g)
    ph = g.physical[:, 0, :].reshape(ncol, 64, 7)
    s0, s1, s2 = np.abs(p0).max(), np.abs(p1).max(), np.abs(p2).max()
    assert np.abs(ph[:, :, 0] - p0).max() <= 1e-11 * s0
    assert np.abs(ph[:, :, 5] - p1).max() <= 1e-10 * s1
    assert np.abs(ph[:, :, 6] - p2).max() <= 1e-9 * s2          # second Chebyshev derivative amplifies round-off by ~zDim^4
    dx = 1e6 / 334
    assert np.abs(ph[:, :, 1]).max() <= 1e-9 * s0 / dx and np.abs(ph[:, :, 2]).max() <= 1e-8 * s0 / dx ** 2
    assert np.abs(ph[:, :, 3]).max() <= 1e-10 * s0 and np.abs(ph[:, :, 4]).max() <= 1e-7 * s0
    g.close()


def test_c4_full_size_two_tiles_equal_one_tile(gpu_lib):
    """bench.py's workload (C4, LinearAdvectionRLZ, synthetic vortex): two steps on one tile and on two radial
    tiles with the plane-distributed solve must give the same state (seam overlap-add of the 3 shared spline
    coefficients, tile-local evaluation of A) at the full 128.9 M-point size."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench as B
    gp = S.GridParameters(geometry="RLZ", xmin=0.0, xmax=B.XMAX, num_cells=B.C4_CELLS, zmin=0.0, zmax=B.ZMAX, zDim=B.ZDIM,
                          vars={"h": 1, "u": 2, "v": 3})
    mp = S.ModelParameters(ts=B.TS, integration_time=B.TS * 10, equation_set="LinearAdvectionRLZ", grid_params=gp,
                           physical_params={"K": B.KDIFF})
    DX = B.XMAX / B.C4_CELLS
    finals = {}
    for nt, ex in ((1, "torch"), (2, "columns")):
        m = S.Model(mp, num_tiles=nt, lib=gpu_lib, exchange=ex)
        tp = m.tile_params
        ics = [B.synthetic_state(tp[0, t], DX, int(tp[2, t]), (int(tp[3, t]) - 1) * 3, B.ZDIM, B.ZMAX) for t in range(nt)]
        m.initialize_tiles(ics)
        m.run(2)
        finals[nt] = np.concatenate([m.state(t, "var_np1") for t in range(nt)], axis=0)
        m.close()
    assert finals[1].shape == finals[2].shape == (128_897_280, 3)
    for v in range(3):
        assert rel_err(finals[2][:, v], finals[1][:, v]) <= 1e-11
    assert np.isfinite(finals[1]).all() and np.abs(finals[1][:, 0]).max() > 1.0


def test_notebook_run_through_integrate_model_and_csv(gpu_lib, tmp_path):
    """The reference's only end-to-end known answer (notebooks/LinearAdvection_example.ipynb: R grid, 100 cells,
    PERIODIC, 2 workers, 2000 steps of 0.05 s, output every 50 s) through the driver a Scythe user calls:
    initial conditions from a CSV (read_physical_grid, src/semiimplicit.jl:134), integrate_model
    (src/Scythe.jl:37-62), physical_out_<t>.csv files named as src/io.jl:5 does.  Checked against the notebook's
    printed values (sanity band) and against the oracle's run of the same model (<= 1e-9)."""
    import json
    from pathlib import Path
    from oracle import model as OM
    gold = json.loads((Path(__file__).parent / "golden" / "linear_advection_notebook.json").read_text())
    gp = S.GridParameters(geometry="R", xmin=-50.0, xmax=50.0, num_cells=100, BCL={"u": S.CubicBSpline.PERIODIC},
                          BCR={"u": S.CubicBSpline.PERIODIC}, vars={"u": 1})
    g = S.createGrid(gp, lib=gpu_lib)
    x = S.getGridpoints(g)
    g.close()
    ic_csv = tmp_path / "gaussian_ic.csv"
    np.savetxt(ic_csv, np.stack([x, np.exp(-(x / 20.0) ** 2)], 1), delimiter=",", header="r,u", comments="", fmt="%.17g")
    mp = S.ModelParameters(ts=0.05, integration_time=100.0, output_interval=50.0, equation_set="LinearAdvection1D",
                           initial_conditions=str(ic_csv), output_dir=str(tmp_path / "out"), grid_params=gp,
                           physical_params={"c_0": 1.0, "K": 0.0})
    final = S.integrate_model(mp, num_tiles=2, write=True, lib=gpu_lib)
    files = sorted(p.name for p in (tmp_path / "out").iterdir())
    assert files == ["physical_out_0.0.csv", "physical_out_100.0.csv", "physical_out_50.0.csv"]
    csv = np.loadtxt(tmp_path / "out" / "physical_out_100.0.csv", delimiter=",", skiprows=1)
    assert np.array_equal(csv[:, 0], x) and np.array_equal(csv[:, 1], final[:, 0, 0])
    uf = final[:, 0, 0]
    got = np.concatenate([uf[:13], uf[-12:]])
    band = np.array(gold["final_u_first13"] + gold["final_u_last12"])
    assert np.abs(got / band - 1).max() < 5e-3
    u0 = np.loadtxt(tmp_path / "out" / "physical_out_0.0.csv", delimiter=",", skiprows=1)[:, 1]
    assert abs(np.sqrt(((u0 - uf) ** 2).sum()) / gold["l2_norm"] - 1) < 0.05
    ogp = G.GridParameters(geometry="R", xmin=-50.0, xmax=50.0, num_cells=100, BCL={"u": spl.PERIODIC},
                           BCR={"u": spl.PERIODIC}, vars={"u": 1})
    omp = OM.ModelParameters(ts=0.05, integration_time=100.0, output_interval=50.0, equation_set="LinearAdvection1D",
                             grid_params=ogp, physical_params={"c_0": 1.0, "K": 0.0})
    run = OM.ModelRun(omp, 2, np.exp(-(x / 20.0) ** 2)[:, None])
    run.run(2000)
    assert rel_err(uf, run.output_patch()[:, 0, 0]) <= STATE_TOL


@pytest.mark.parametrize("bcb,bct", [("R0", "R0"), ("R1T0", "R1T1"), ("R1T2", "R1T0")])
def test_chebyshev_column_api(bcb, bct, gpu_lib):
    from oracle import chebyshev as och
    from test_kernels_emulated import check_chebyshev_column_api
    check_chebyshev_column_api(S, och, gpu_lib, bcb, bct, nz=64, ncol=1000)


@pytest.mark.parametrize("ntiles", [1, 2])
def test_pipelined_host_cycle_is_bit_identical(ntiles, gpu_lib):
    """Host-driven stepping through the copy streams (sb_model_stage_in/out, page-locked host buffers, 4.2 M points x 3
    variables = 101 MB each way per step so that copies and kernels really overlap): every step's result is
    bit-identical to set_state -> cycle -> get_state."""
    from helpers import check_host_pipeline
    from oracle import grids as G
    gp = G.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=60, zmin=0, zmax=1e3, zDim=64, vars={"h": 1, "u": 2, "v": 3})
    r, l, z = G.createGrid(gp).getGridpoints().T
    ic = np.zeros((r.size, 3))
    ic[:, 0] = np.exp(-((r * np.cos(l) - 3e4) ** 2 + (r * np.sin(l)) ** 2) / 4e8) * np.cos(z / 400.0)
    ic[:, 1] = 5 * np.cos(l) * (1 + 0.2 * np.sin(z / 300.0))
    ic[:, 2] = -5 * np.sin(l) * np.exp(-z / 900.0)
    case = dict(gp=gp, eq="LinearAdvectionRLZ", prm={"K": 100.0}, ts=50.0, n=8, ic=ic, tiles=(1,))
    check_host_pipeline(case, gpu_lib, ntiles=ntiles, nsteps=6, pinned=True)
    check_host_pipeline(M_CASES["LinearAdvectionRLZ"], gpu_lib, ntiles=ntiles, nsteps=5, pinned=False)   # pageable buffers: still correct


def test_linearity_rlz(gpu_lib):
    gp = S.GridParameters(geometry="RLZ", xmin=0, xmax=1e5, num_cells=30, zmin=0, zmax=1e4, zDim=64, vars={"a": 1, "b": 2, "c": 3})
    g = S.createGrid(gp, lib=gpu_lib)
    rng = np.random.default_rng(11)
    x, y = rng.standard_normal((2, g.N))
    g.physical[:, 0, 0], g.physical[:, 1, 0], g.physical[:, 2, 0] = x, y, 2.5 * x - 0.75 * y
    S.spectralTransform(g)
    sp = g.spectral
    assert rel_err(sp[:, 2], 2.5 * sp[:, 0] - 0.75 * sp[:, 1]) <= 1e-13
    S.gridTransform(g)
    ph = g.physical
    for d in range(g.D):
        assert rel_err(ph[:, 2, d], 2.5 * ph[:, 0, d] - 0.75 * ph[:, 1, d]) <= 1e-12
    g.close()


def test_error_behaviour(gpu_lib):
    with pytest.raises(S.DomainError):
        S.createGrid(S.GridParameters(geometry="XYZ", xmin=0, xmax=1, num_cells=4), lib=gpu_lib)
    with pytest.raises(S.DomainError):
        S.calcTileSizes(S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=8), 3, lib=gpu_lib)
    mp = S.ModelParameters(ts=1.0, equation_set="Kepert2017_TCBL",
                           grid_params=S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=8))
    with pytest.raises(S.UnsupportedError):
        S.Model(mp, lib=gpu_lib)
    # checkCFL: NaN -> error naming the variable and the 1-based index (src/semiimplicit.jl:745)
    g = S.createGrid(S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=8, vars={"u": 1, "q": 2}), lib=gpu_lib)
    g.physical[:] = 0.0
    g.physical[5, 1, 0] = np.nan
    g.upload_physical(0, 1)
    with pytest.raises(S.ScytheError, match="NaN found in variable q at index6"):
        S.checkCFL(g)
    g.close()


def test_checkpoint_restart_is_exact_on_device(gpu_lib, tmp_path):
    """Model.checkpoint / restore on the real device (fused K3+K4 default path, 16 levels; and the semi-implicit set with
    its impdot history): 4 steps straight == 2 steps, checkpoint, restore into a fresh model, 2 more steps, bit for bit."""
    for name, ntiles in (("LinearAdvectionRLZ_z16_fused", 2), ("Euler_test_semiimplicit", 1)):
        case = M_CASES[name]
        a = pkg_model(case, ntiles, gpu_lib)
        a.initialize(case["ic"])
        a.run(2)
        a.checkpoint(tmp_path / f"{name}.npz")
        a.run(2)
        b = pkg_model(case, ntiles, gpu_lib)
        b.restore(tmp_path / f"{name}.npz")
        assert b.t == 2
        b.run(2)
        for i in range(ntiles):
            for k in ("var_np1", "expdot_nm1", "expdot_nm2"):
                assert np.array_equal(a.state(i, k), b.state(i, k)), (name, i, k)
        a.close(); b.close()
    other = pkg_model(M_CASES["LinearAdvectionRLZ"], 2, gpu_lib)
    with pytest.raises(ValueError):          # a checkpoint of another grid is refused before anything is written
        other.restore(tmp_path / "LinearAdvectionRLZ_z16_fused.npz")
    other.close()


def test_multi_gpu_exchange_modes(gpu_lib):
    """Two ranks on two GPUs under torchrun/NCCL (tests/dist_gpu_check.py): every exchange transport incl. the default
    CUDA-IPC peer stores gives identical tile states, those match the oracle, and an injected one-sided IPC failure
    makes every rank fall back.  Needs >= 2 GPUs (the single-GPU box of the round-end run skips it; bench.py's
    multi-rank preflight covers the default path there)."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    if gpu_lib.sb_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    port = str(29500 + os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", port, str(root / "tests" / "dist_gpu_check.py")],
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port), capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "vs oracle rel_err" in r.stdout


@pytest.mark.parametrize("name", ["Oneway_ShallowWater_Slab", "LinearAdvection1D", "Euler_test_semiimplicit", "LinearAdvectionRLZ_z16_fused"])
def test_graph_replayed_steps_are_bit_identical(name, gpu_lib):
    """sb_model_run replays three AB3 steps per CUDA-graph launch on launch-bound grids (one graph per phase of the
    history pointer rotation); Model.step() launches every kernel eagerly.  Same bits after 14 steps, with run() entered
    at different rotation phases."""
    case = M_CASES[name]
    nt = case["tiles"][-1]
    a = pkg_model(case, nt, gpu_lib)
    a.initialize(case["ic"])
    for _ in range(14):
        a.step()
    b = pkg_model(case, nt, gpu_lib)
    b.initialize(case["ic"])
    b.run(7)          # steps 1-3 eager, 4-6 one graph, 7 eager
    b.run(4)          # phase shifted by one: a second graph
    b.run(3)          # third phase
    assert b.launch_count() == a.launch_count()
    for i in range(nt):
        for k in ("var_np1", "expdot_nm1", "expdot_nm2"):
            assert np.array_equal(a.state(i, k), b.state(i, k)), (i, k)
    a.close(); b.close()


@pytest.mark.parametrize("ntiles", [1, 2])
def test_overlapped_step_is_bit_identical(ntiles, gpu_lib, monkeypatch):
    """The default C4 step runs the FP64-bound ring FFTs and the HBM-bound Chebyshev stages side by side on two streams,
    by ring batches, with dynamic FFT work shares on 148 - k SMs.  Forced here on a 24-cell x 64-level grid: same bits as
    the one-stream step for 1, 4 and 7 batches, and the state still matches the oracle."""
    from helpers import check_overlapped_step
    case = B_CASES["LinearAdvectionRLZ_z64_24cells_fused"]
    check_overlapped_step(case, gpu_lib, ntiles, monkeypatch, batches=(1, 4, 7))
    monkeypatch.setenv("SB_OVERLAP", "1")
    monkeypatch.setenv("SB_OVERLAP_BATCHES", "5")
    c1 = dict(case, tiles=(ntiles,))
    assert check_model(c1, gpu_lib) <= STATE_TOL


def test_launcher_on_device(gpu_lib, tmp_path, monkeypatch):
    """python -m scythe_jl_b200.run (the run_Scythe.jl replacement) on the device: reference-style model file in, CSV + NetCDF
    out, checkpoint + restart ending on the same bytes (tests/test_driver_io.py holds the checks)."""
    from test_driver_io import check_launcher, check_netcdf_round_trip
    check_netcdf_round_trip(gpu_lib, tmp_path)
    check_launcher(gpu_lib, tmp_path, monkeypatch, nsteps=40)


def test_moist_error_behaviour(gpu_lib):
    from test_kernels_emulated import check_moist_error_behaviour
    check_moist_error_behaviour(S, gpu_lib)


def test_reference_state_files(gpu_lib, tmp_path):
    from test_kernels_emulated import check_reference_state_files
    check_reference_state_files(S, gpu_lib, tmp_path)


@pytest.mark.parametrize("ntiles", [1, 2])
def test_passive_history_of_tendency_free_variables(ntiles, gpu_lib):
    from test_kernels_emulated import check_passive_history
    check_passive_history(S, gpu_lib, B_CASES["LinearAdvectionRLZ_z64_24cells_fused"], ntiles, n0=9, n2=6)


def test_launcher_two_gpus(gpu_lib, tmp_path):
    """python -m scythe_jl_b200.run --gpus 2 model.jl (one radial tile per GPU under torchrun / NCCL, rank 0 writes): same
    final CSV, to round-off, as the 2-tile run on one GPU; checkpoint files are per rank.  Needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    if gpu_lib.sb_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    gp = S.GridParameters(geometry="RL", xmin=0.0, xmax=24.0, num_cells=24, BCL={"h": S.CubicBSpline.R1T1, "u": S.CubicBSpline.R1T0,
                          "v": S.CubicBSpline.R1T0}, BCR={"h": S.CubicBSpline.R0, "u": S.CubicBSpline.R0, "v": S.CubicBSpline.R0},
                          vars={"h": 1, "u": 2, "v": 3})
    g = S.createGrid(gp, lib=gpu_lib)
    pts = S.getGridpoints(g).reshape(g.N, -1)
    g.close()
    r, l = pts[:, 0], pts[:, 1]
    ic = np.stack([r, l, np.exp(-((r * np.cos(l) - 8.0) ** 2 + (r * np.sin(l)) ** 2) / 16.0), 0.5 * np.cos(l), -0.5 * np.sin(l)], 1)
    np.savetxt(tmp_path / "ic.csv", ic, delimiter=",", header="r,l,h,u,v", comments="", fmt="%.17g")

    def model_text(out):
        return f'''model = ModelParameters(
            ts = 0.05, integration_time = 1.0, output_interval = 0.5, equation_set = "LinearAdvectionRL",
            initial_conditions = "{tmp_path / "ic.csv"}", output_dir = "{out}/",
            grid_params = GridParameters(geometry = "RL", xmin = 0.0, xmax = 24.0, num_cells = 24,
                BCL = Dict("h" => CubicBSpline.R1T1, "u" => CubicBSpline.R1T0, "v" => CubicBSpline.R1T0),
                BCR = Dict("h" => CubicBSpline.R0, "u" => CubicBSpline.R0, "v" => CubicBSpline.R0),
                vars = Dict("h" => 1, "u" => 2, "v" => 3)),
            physical_params = Dict(:K => 0.01))'''
    env = dict(os.environ, PYTHONPATH=str(root) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    (tmp_path / "one.jl").write_text(model_text(tmp_path / "one"))
    (tmp_path / "two.jl").write_text(model_text(tmp_path / "two"))
    a = subprocess.run([sys.executable, "-m", "scythe_jl_b200.run", "-w", "2", str(tmp_path / "one.jl")], env=env, cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert a.returncode == 0, a.stdout[-2000:] + a.stderr[-2000:]
    b = subprocess.run([sys.executable, "-m", "scythe_jl_b200.run", "--gpus", "2", "--master-port", str(29500 + os.getpid() % 2000),
                        "--checkpoint", str(tmp_path / "ck.npz"), str(tmp_path / "two.jl")], env=env, cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert b.returncode == 0, b.stdout[-3000:] + b.stderr[-3000:]
    assert "Model complete!" in b.stdout
    names = sorted(p.name for p in (tmp_path / "two").iterdir())
    assert names == ["physical_out_0.0.csv", "physical_out_0.5.csv", "physical_out_1.0.csv"]
    one = np.loadtxt(tmp_path / "one" / "physical_out_1.0.csv", delimiter=",", skiprows=1)
    two = np.loadtxt(tmp_path / "two" / "physical_out_1.0.csv", delimiter=",", skiprows=1)
    assert np.array_equal(one[:, :2], two[:, :2])
    assert rel_err(two[:, 2:], one[:, 2:]) <= 1e-12
    assert sorted(p.name for p in tmp_path.glob("ck.tiles*.npz")) == ["ck.tiles0-0.npz", "ck.tiles1-1.npz"]
