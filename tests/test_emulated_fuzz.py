"""The seeded random grid configurations of tests/test_gpu_fuzz.py (geometry, cells, levels, boundary conditions, tile
offset, variables) through the TEST-ONLY CPU emulation of the kernel sources vs the oracle, for the configurations small
enough for the emulation: kernel indexing at the boundaries between kernel variants is checked
before any GPU time is spent.  The GPU suite runs the same 60 seeds on the product library."""
import pytest

from helpers import TRANSFORM_TOL, check_transforms
from oracle import grids as G
from test_gpu_fuzz import random_case

MAX_POINTS = 400000


@pytest.mark.parametrize("seed", range(60))
def test_random_grid_transforms_match_oracle_emulated(seed, emu_lib):
    gp = random_case(seed)
    og = G.createGrid(gp)
    if og.physical.shape[0] > MAX_POINTS:
        pytest.skip("too many points for the CPU emulation; covered by tests/test_gpu_fuzz.py")
    eB, eP = check_transforms(gp, emu_lib, seed=seed)
    assert eB <= TRANSFORM_TOL, (gp, eB)
    tol = [TRANSFORM_TOL] * og.D
    if gp.geometry in ("RL", "RLZ"):
        tol[4] = max(TRANSFORM_TOL, 1e-15 * og.kDim ** 2)
    for d, (e, t) in enumerate(zip(eP, tol)):
        assert e <= t, (gp, d, e)
