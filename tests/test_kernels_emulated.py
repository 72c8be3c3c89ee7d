"""CPU (-m "not gpu") check of the *kernel sources* themselves: the .cu files are compiled for the
host against the test-only thread emulation shim (csrc/cuda_emu.h) and compared with the oracle.
This catches indexing / maths errors before GPU time is spent; the GPU parity tests proper are in
test_gpu_parity.py.  The emulated library is never reachable from the product package."""
import pytest

from helpers import (STATE_TOL, TRANSFORM_TOL, check_model, check_needed_slots, check_transforms, model_cases,
                     transform_cases)

T_CASES = transform_cases()
M_CASES = model_cases()


@pytest.mark.parametrize("name", sorted(T_CASES))
def test_transforms_match_oracle(name, emu_lib):
    eB, eP = check_transforms(T_CASES[name], emu_lib)
    assert eB <= TRANSFORM_TOL, f"spectralTransform rel err {eB}"
    assert max(eP) <= TRANSFORM_TOL, f"gridTransform rel err per slot {eP}"


@pytest.mark.parametrize("name", sorted(M_CASES))
def test_timestep_matches_oracle(name, emu_lib):
    case = dict(M_CASES[name])
    case["n"] = min(case["n"], 4)
    assert check_model(case, emu_lib) <= STATE_TOL


@pytest.mark.parametrize("name", sorted(M_CASES))
def test_needed_slots_state_is_bit_identical(name, emu_lib):
    case = dict(M_CASES[name])
    case["n"] = min(case["n"], 3)     # Euler, AB2, AB3
    check_needed_slots(case, emu_lib)


@pytest.mark.parametrize("name,ntiles", [("LinearAdvection1D", 3), ("LinearAdvectionRLZ", 2), ("Euler_test_semiimplicit", 2)])
def test_plane_distributed_solve_single_process(name, ntiles, emu_lib):
    """exchange="columns": B stays tile-local, the z-mode-plane owner overlap-adds, solves and hands every tile
    its own slice of A (here one process owns every plane) -- same answer as the shared-array scheme."""
    case = dict(M_CASES[name])
    case["tiles"] = (ntiles,)
    case["n"] = min(case["n"], 3)
    assert check_model(case, emu_lib, exchange="columns") <= STATE_TOL
    if name == "LinearAdvectionRLZ":
        # peer-memory form (here every "peer" buffer is this process's own): fwd_r scatters into the owner's slots,
        # the cut-out kernel writes the tiles' A in place
        assert check_model(case, emu_lib, exchange="columns-p2p") <= STATE_TOL


@pytest.mark.parametrize("name,ntiles", [("LinearAdvectionRLZ", 2), ("LinearAdvection1D", 1)])
def test_pipelined_host_cycle_is_bit_identical(name, ntiles, emu_lib):
    """sb_model_stage_in / cycle / sb_model_stage_out (asynchronous, double-buffered) == set_state / cycle / get_state."""
    from helpers import check_host_pipeline
    check_host_pipeline(M_CASES[name], emu_lib, ntiles=ntiles, nsteps=5)
    if name == "LinearAdvection1D":      # argument checking of the asynchronous calls (no hidden copies, no bad tiles)
        import ctypes as C
        import numpy as np
        from helpers import pkg_model
        m = pkg_model(M_CASES[name], 1, emu_lib)
        m.initialize(M_CASES[name]["ic"])
        g = m.tiles[0]
        good = np.zeros((g.N, g.V), order="F")
        with pytest.raises(ValueError):
            m.stage_in(0, np.zeros((g.N, g.V + 1), order="F")[:, 1:][::1].astype(np.float32))
        with pytest.raises(ValueError):
            m.stage_out(0, np.zeros((g.V, 2 * g.N)).T[::2])          # strided view
        buf = good.ctypes.data_as(C.POINTER(C.c_double))
        assert emu_lib.sb_model_stage_in(m.handle, 7, buf) != 0 and b"bad argument" in emu_lib.sb_last_error()
        assert emu_lib.sb_model_stage_out(m.handle, -1, buf) != 0
        assert emu_lib.sb_model_stage_drain(None, 1) != 0
        m.stage_in(0, good); m.drain()
        m.close()


@pytest.mark.parametrize("name,ntiles,exchange", [("LinearAdvectionRLZ", 2, "torch"), ("Euler_test_semiimplicit", 2, "columns")])
def test_checkpoint_restart_is_exact(name, ntiles, exchange, emu_lib, tmp_path):
    """4 steps straight == 2 steps, checkpoint, restore into a fresh model, 2 more steps (bit for bit): the AB3 /
    semi-implicit history travels with the checkpoint, unlike the reference's restart from an output file."""
    import numpy as np
    from helpers import pkg_model
    case = M_CASES[name]
    a = pkg_model(case, ntiles, emu_lib, exchange=exchange)
    a.initialize(case["ic"])
    a.run(2)
    a.checkpoint(tmp_path / "ck.npz")
    a.run(2)
    b = pkg_model(case, ntiles, emu_lib, exchange=exchange)
    b.restore(tmp_path / "ck.npz")
    assert b.t == 2
    b.run(2)
    for i in range(ntiles):
        for k in ("var_np1", "expdot_nm1", "expdot_nm2"):
            assert np.array_equal(a.state(i, k), b.state(i, k)), (i, k)
    a.close(); b.close()


def test_integrate_model_driver_and_csv_files(emu_lib, tmp_path):
    """integrate_model (src/Scythe.jl:37-62): CSV initial conditions in, physical_out_<t>.csv out at the output
    cadence of model_loop (src/semiimplicit.jl:288-293), file names as src/io.jl:5; final state == oracle."""
    import numpy as np
    import scythe_jl_b200 as S
    from helpers import run_oracle, rel_err
    case = dict(M_CASES["LinearAdvection1D"])
    from helpers import to_pkg
    gp = to_pkg(case["gp"])
    g = S.createGrid(gp, lib=emu_lib)
    x = S.getGridpoints(g)
    g.close()
    ic = tmp_path / "ic.csv"
    np.savetxt(ic, np.stack([x, case["ic"][:, 0]], 1), delimiter=",", header="r,u", comments="", fmt="%.17g")
    mp = S.ModelParameters(ts=case["ts"], integration_time=case["ts"] * 6, output_interval=case["ts"] * 3,
                           equation_set=case["eq"], initial_conditions=str(ic), output_dir=str(tmp_path / "out"),
                           grid_params=gp, physical_params=case["prm"])
    final = S.integrate_model(mp, num_tiles=3, write=True, lib=emu_lib)
    names = sorted(p.name for p in (tmp_path / "out").iterdir())
    assert names == ["physical_out_0.0.csv", "physical_out_0.15.csv", "physical_out_0.3.csv"]
    case["n"] = 6
    assert rel_err(final[:, 0, 0], run_oracle(case, 3).output_patch()[:, 0, 0]) <= STATE_TOL
    with pytest.raises(S.ScytheError):
        S.integrate_model(mp, num_tiles=0, lib=emu_lib)      # "Need to add at least 1 worker process" (src/Scythe.jl:39-41)


@pytest.mark.parametrize("bcb,bct", [("R0", "R0"), ("R1T0", "R1T1"), ("R1T2", "R1T0"), ("R1T1", "R0")])
def test_chebyshev_column_api_matches_oracle(bcb, bct, emu_lib):
    """Chebyshev1D column transforms (CB, CA, CI, CIx, CIxx, CIInt) and the dct_* matrices the semi-implicit
    Helmholtz operator is built from (src/semiimplicit.jl:569-574, 768-781), batched over 37 columns."""
    import numpy as np
    import scythe_jl_b200 as S
    from oracle import chebyshev as och
    check_chebyshev_column_api(S, och, emu_lib, bcb, bct)


def check_chebyshev_column_api(S, och, lib, bcb, bct, nz=24, ncol=37):
    import numpy as np
    ocp = och.ChebyshevParameters(0.5, 7.5, nz, 0, getattr(och, bcb), getattr(och, bct))
    oc = och.Chebyshev1D(ocp)
    c = S.Chebyshev1D(S.ChebyshevParameters(zmin=0.5, zmax=7.5, zDim=nz, bDim=0, BCB=getattr(S.Chebyshev, bcb),
                                            BCT=getattr(S.Chebyshev, bct)), lib=lib)
    assert np.abs(c.mishPoints - oc.mishPoints).max() <= 1e-14 * 7.5
    rng = np.random.default_rng(4)
    u = rng.standard_normal((nz, ncol))
    b, ob = c.CBtransform(u), oc.CBtransform(u)
    assert b.shape == ob.shape and np.abs(b - ob).max() <= 1e-14
    a, oa = c.CAtransform(ob), oc.CAtransform(ob)
    assert a.shape == oa.shape == (nz, ncol) and np.abs(a - oa).max() <= 1e-13
    for name, tol in (("CItransform", 1e-13), ("CIxtransform", 1e-11), ("CIxxtransform", 1e-9)):
        got, want = getattr(c, name)(oa), getattr(oc, name)(oa)
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), name
    got, want = c.CIInttransform(oa, 2.5), oc.CIInttransform(oa, 2.5)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    assert np.abs(c.CBtransform(u[:, 0]) - ob[:, 0]).max() <= 1e-14          # single column
    L = 7.0
    assert np.abs(S.dct_matrix(nz, lib=lib) - och.dct_matrix(nz)).max() <= 1e-13
    d1, o1 = S.dct_1st_derivative(nz, L, lib=lib), och.dct_1st_derivative(nz, L)
    d2, o2 = S.dct_2nd_derivative(nz, L, lib=lib), och.dct_2nd_derivative(nz, L)
    assert np.abs(d1 - o1).max() <= 1e-12 * np.abs(o1).max() and np.abs(d2 - o2).max() <= 1e-12 * np.abs(o2).max()


@pytest.mark.parametrize("name", ["RL", "RLZ", "RZ"])
def test_patch_and_halo_maps_match_oracle_and_overlap_add(name, emu_lib):
    """calcPatchMap / calcHaloMap (src/semiimplicit.jl:79-86) as host index maps: the reference's host-side overlap-add  shared[patchMap] = tileView; shared[haloMap(prev)] += haloView  (:323-329)
    of the tiles' B reproduces the patch's B."""
    import numpy as np
    import scythe_jl_b200 as S
    from helpers import to_pkg
    from oracle import grids as OG
    gp = T_CASES[name]
    gp = OG.GridParameters(**{**gp.__dict__, "num_cells": 9, "xmin": 0.0, "xmax": 9.0})
    opatch = OG.createGrid(gp)
    patch = S.createGrid(to_pkg(gp), lib=emu_lib)
    tp = S.calcTileSizes(patch.params, 3, lib=emu_lib)
    otp = OG.calcTileSizes(opatch, 3)
    assert np.array_equal(tp, otp)
    rng = np.random.default_rng(2)
    u = rng.standard_normal((patch.N, patch.V))
    patch.physical[:, :, 0] = u
    S.spectralTransform(patch)
    shared = np.zeros_like(patch.spectral)
    pts = np.concatenate([[0], np.cumsum(tp[4]).astype(np.int64)])
    prev = None
    for t in range(3):
        tile = S.createGrid(S.tile_grid_params(patch.params, tp, t), lib=emu_lib)
        tile.physical[:, :, 0] = u[pts[t]:pts[t + 1]]
        S.spectralTransform(tile)
        pmask, trows = S.calcPatchMap(patch, tile)
        hmask, hrows = S.calcHaloMap(patch, tile)
        assert pmask.sum() == trows.size * patch.V and hmask.sum() == hrows.size * patch.V
        for v in range(patch.V):
            shared[np.flatnonzero(pmask[:, v]), v] = tile.spectral[trows, v]
            if prev is not None:
                shared[np.flatnonzero(prev[0][:, v]), v] += prev[1][:, v]
        if t == 2:   # the last tile's halo goes to the master (src/semiimplicit.jl:279-282)
            for v in range(patch.V):
                shared[np.flatnonzero(hmask[:, v]), v] += tile.spectral[hrows, v]
        prev = (hmask, tile.spectral[hrows].copy())
        tile.close()
    assert np.abs(shared - patch.spectral).max() <= 1e-13 * np.abs(patch.spectral).max()
    assert S.allocateSplineBuffer(patch, patch) is None
    patch.close()


def test_overlapped_step_is_bit_identical(emu_lib, monkeypatch):
    """LinearAdvectionRLZ, fused K3+K4: ring FFTs and Chebyshev stages on two streams by ring batches == one stream."""
    from helpers import check_overlapped_step
    check_overlapped_step(M_CASES["LinearAdvectionRLZ_z16_fused"], emu_lib, 1, monkeypatch, batches=(3,))
    check_overlapped_step(M_CASES["LinearAdvectionRLZ_z16_fused"], emu_lib, 2, monkeypatch, batches=(2,), exchange="columns")
