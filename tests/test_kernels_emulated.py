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


def check_moist_error_behaviour(S, lib):
    """BF02_test / rainfall_test: the names condensation_adjustment looks up (src/microphysics.jl:141-165) must exist --
    the reference throws KeyError on the first step, here the model is refused when it is created -- and the columns must
    be in the order the equation sets index by position (src/testModels.jl:233-271, :403-447)."""
    from helpers import moist_case, to_pkg
    from oracle import grids as G
    def gp(names):
        return G.GridParameters(geometry="RZ", xmin=-1e4, xmax=1e4, num_cells=4, zmin=0, zmax=1e4, zDim=8,
                                vars={n: i + 1 for i, n in enumerate(names)})
    good = gp(["s", "xi", "mu", "u", "w", "mu_c", "qss", "mu_r"])
    ref, _ = moist_case(good, rain=False)
    sref = S.ReferenceState(ref.sbar, ref.xibar, ref.mubar, ref.mu_lbar, ref.Pxi_bar)
    def make(eq, g, r=sref):
        return S.Model(S.ModelParameters(ts=0.1, equation_set=eq, grid_params=to_pkg(g), physical_params={"K": 1.0}), ref_state=r, lib=lib)
    make("BF02_test", good).close()
    with pytest.raises(S.ScytheError, match='KeyError: key "mu_c" not found'):
        make("BF02_test", gp(["s", "xi", "mu", "u", "w", "mu_l", "qss"]))
    with pytest.raises(S.ScytheError, match="expects vars"):
        make("rainfall_test", good)                       # BF02 order handed to rainfall_test
    with pytest.raises(S.ScytheError, match="needs a reference state"):
        make("rainfall_test", gp(["s", "xi", "mu", "u", "w", "mu_c", "mu_r", "qss"]), r=None)


def test_moist_error_behaviour(emu_lib):
    import scythe_jl_b200 as S
    check_moist_error_behaviour(S, emu_lib)


def check_reference_state_files(S, lib, tmp_path):
    """createModelTile's reference-state set-up (src/semiimplicit.jl:62-73) through the product: a sounding file through
    interpolate_reference_file (src/reference_state.jl:17-136) and an exact file through exact_reference_state (:159-199),
    against the oracle's restatement; then a model created from `ref_state_file` alone steps like one handed the object."""
    import numpy as np
    from helpers import model_cases, pkg_model, to_pkg, rel_err
    from oracle import chebyshev as och
    from oracle import model as OM
    case = model_cases()["rainfall_test_semiimplicit"]
    ogp = case["gp"]
    gp = to_pkg(ogp)
    z = och.mish_points(och.ChebyshevParameters(ogp.zmin, ogp.zmax, ogp.zDim, ogp.b_zDim))
    # a sounding on its own levels: surface line, then altitude / theta / q_v[g/kg]; one model level coincides with a sounding level
    alts = np.concatenate([np.linspace(250.0, 12000.0, 48), [float(z[3])]])
    alts.sort()
    lines = ["1005.0 299.5 16.0"] + [f"{float(a)!r} {float(299.5 + 4.2e-3 * a)!r} {float(16.0 * np.exp(-a / 2600.0))!r}" for a in alts]
    snd = tmp_path / "sounding.txt"
    snd.write_text("\n".join(lines) + "\n\n999 1 1\n")          # anything after the first blank line is not read
    mp = S.ModelParameters(ts=0.5, equation_set="rainfall_test", grid_params=gp, physical_params={"K": 50.0},
                           ref_state_file=str(snd), options={"semiimplicit": True, "exact_reference_state": False})
    got = S.interpolate_reference_file(mp, z, lib=lib)
    want = OM.interpolate_reference_file(ogp, snd.read_text(), z)
    for k in ("sbar", "xibar", "mubar", "mu_lbar"):
        assert rel_err(getattr(got, k), getattr(want, k)) <= 1e-12, k
    assert abs(got.Pxi_bar / want.Pxi_bar - 1) <= 1e-12
    # exact file: z sbar xibar mubar mu_lbar per level
    ref = case["ref"]
    # levels as the library computes them: the file is matched against them number for number (the reference compares text)
    zc = S.Chebyshev1D(S.ChebyshevParameters(zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, bDim=gp.b_zDim), lib=lib).mishPoints.copy()
    assert np.abs(zc - och.mish_points(och.ChebyshevParameters(ogp.zmin, ogp.zmax, ogp.zDim, ogp.b_zDim))).max() <= 1e-11 * ogp.zmax
    Tk = 300.0 - 6.5e-3 * zc
    raw = np.stack([zc, OM.entropy(Tk, 1.1 * np.exp(-zc / 8e3), 0.01 * np.exp(-zc / 2500)), np.log(1.1 * np.exp(-zc / 8e3) / OM.rho_d0),
                    OM.bhyp(0.01 * np.exp(-zc / 2500)), OM.bhyp(1e-4 * np.exp(-((zc - 3e3) / 1e3) ** 2))], 1)
    ex = tmp_path / "exact.txt"
    ex.write_text("\n".join(" ".join(repr(float(v)) for v in row) for row in raw) + "\n")
    mp.ref_state_file, mp.options = str(ex), {"semiimplicit": True, "exact_reference_state": True}
    got = S.exact_reference_state(mp, zc, lib=lib)
    want = OM.exact_reference_state_from_profiles(ogp, raw[:, 1], raw[:, 2], raw[:, 3], raw[:, 4])
    for k in ("sbar", "xibar", "mubar", "mu_lbar"):
        assert rel_err(getattr(got, k), getattr(want, k)) <= 1e-12, k
    assert abs(got.Pxi_bar / want.Pxi_bar - 1) <= 1e-12
    bad = tmp_path / "bad.txt"
    bad.write_text(ex.read_text().replace(repr(float(zc[2])), repr(float(zc[2]) + 1.0), 1))
    mp.ref_state_file = str(bad)
    with pytest.raises(S.DomainError, match="Model level does not match reference level"):
        S.exact_reference_state(mp, zc, lib=lib)
    # a model built from the file alone == a model handed the ReferenceState object
    mp.ref_state_file = str(ex)
    mp.ts = case["ts"]
    a = S.Model(mp, num_tiles=1, lib=lib)
    b = S.Model(mp, num_tiles=1, ref_state=got, lib=lib)
    for m in (a, b):
        m.initialize(case["ic"])
        m.run(2)
    assert np.array_equal(a.state(0, "var_np1"), b.state(0, "var_np1"))
    a.close(); b.close()


def test_reference_state_files(emu_lib, tmp_path):
    import scythe_jl_b200 as S
    check_reference_state_files(S, emu_lib, tmp_path)


def check_passive_history(S, lib, case, ntiles=1, n0=3, n2=3):
    """LinearAdvectionRLZ writes a tendency for h only (src/testModels.jl:93): the fused K3+K4 kernel leaves the u / v
    history arrays alone while they still hold their initial zeros.  (1) the state equals the all-slots path bit for bit,
    history arrays included; (2) once somebody stores a non-zero history for u (sb_model_set_state), the kernel must
    read it again: still bit-identical to the all-slots path, which always does; (3) storing zeros keeps the fast path."""
    import numpy as np
    from helpers import pkg_model
    runs = {}
    for mode in ("all", "fused"):
        m = pkg_model(case, ntiles, lib)
        m.set_k3_slots(mode)
        m.initialize(case["ic"])
        m.run(n0)     # on the device (n0 = 9) the last six steps replay CUDA graphs of the step: they must not outlive the guarantee
        snap = [[m.state(i, k).copy() for k in ("var_np1", "expdot_nm1", "expdot_nm2")] for i in range(ntiles)]
        for i in range(ntiles):
            h = m.state(i, "expdot_nm1")
            assert not h[:, 1:].any()                       # u, v: zeros
            m.lib.check(m.lib.sb_model_set_state(m.handle, i, 2, S.api._ptr(np.asfortranarray(h))))   # zeros for u, v: fast path kept
        m.run(1)
        snap2 = [[m.state(i, k).copy() for k in ("var_np1", "expdot_nm1", "expdot_nm2")] for i in range(ntiles)]
        for i in range(ntiles):
            h = m.state(i, "expdot_nm1")
            h[:, 1] = 1e-3 * np.cos(np.arange(h.shape[0]))  # a history for u from outside
            m.lib.check(m.lib.sb_model_set_state(m.handle, i, 2, S.api._ptr(np.asfortranarray(h))))
        m.run(n2)
        snap3 = [[m.state(i, k).copy() for k in ("var_np1", "expdot_nm1", "expdot_nm2")] for i in range(ntiles)]
        runs[mode] = (snap, snap2, snap3)
        m.close()
    for a, b in zip(runs["all"], runs["fused"]):
        for ta, tb in zip(a, b):
            for x, y in zip(ta, tb):
                assert np.array_equal(x, y)
    assert np.abs(runs["fused"][2][0][0][:, 1] - runs["fused"][1][0][0][:, 1]).max() > 1e-4     # the injected history acted on u


def test_passive_history_of_tendency_free_variables(emu_lib):
    import scythe_jl_b200 as S
    check_passive_history(S, emu_lib, M_CASES["LinearAdvectionRLZ_z16_fused"])


@pytest.mark.parametrize("by_mask", ["0", "1"])
def test_k3_passes_by_mask_and_by_union_leave_the_same_state(by_mask, emu_lib, monkeypatch):
    """In-step tileTransform! passes either group consecutive variables with identical slot masks (large grids) or take all
    variables with the union of their masks (launch-bound grids): the boundary-layer set (three different masks, one of them
    empty) must leave the all-slots state bit for bit either way, with the unread slots poisoned."""
    from helpers import check_needed_slots
    monkeypatch.setenv("SB_K3_BY_MASK", by_mask)
    check_needed_slots(M_CASES["Oneway_ShallowWater_HeightResolvedBL_z16"], emu_lib)
