"""CPU (-m "not gpu") check of the *kernel sources* themselves: the .cu files are compiled for the
host against the test-only thread emulation shim (csrc/cuda_emu.h) and compared with the oracle.
This catches indexing / maths errors before GPU time is spent; the GPU parity tests proper are in
test_gpu_parity.py.  The emulated library is never reachable from the product package."""
import pytest

from helpers import STATE_TOL, TRANSFORM_TOL, check_model, check_transforms, model_cases, transform_cases

T_CASES = transform_cases()
M_CASES = model_cases()
FAST_MODELS = ["LinearAdvection1D", "LinearShallowWater1D", "LinearAdvectionRL_K0", "LinearAdvectionRZ",
               "Euler_test_semiimplicit", "LinearAdvectionRLZ"]


@pytest.mark.parametrize("name", sorted(T_CASES))
def test_transforms_match_oracle(name, emu_lib):
    eB, eP = check_transforms(T_CASES[name], emu_lib)
    assert eB <= TRANSFORM_TOL, f"spectralTransform rel err {eB}"
    assert max(eP) <= TRANSFORM_TOL, f"gridTransform rel err per slot {eP}"


@pytest.mark.parametrize("name", FAST_MODELS)
def test_timestep_matches_oracle(name, emu_lib):
    case = dict(M_CASES[name])
    case["tiles"] = case["tiles"][-1:]   # the multi-tile variant only (CPU time)
    case["n"] = min(case["n"], 3)
    assert check_model(case, emu_lib) <= STATE_TOL


@pytest.mark.parametrize("name,ntiles", [("LinearAdvection1D", 3), ("LinearAdvectionRLZ", 2), ("Euler_test_semiimplicit", 2)])
def test_plane_distributed_solve_single_process(name, ntiles, emu_lib):
    """exchange="columns": B stays tile-local, the z-mode-plane owner overlap-adds, solves and hands every tile
    its own slice of A (here one process owns every plane) -- same answer as the shared-array scheme."""
    case = dict(M_CASES[name])
    case["tiles"] = (ntiles,)
    case["n"] = min(case["n"], 3)
    assert check_model(case, emu_lib, exchange="columns") <= STATE_TOL


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
