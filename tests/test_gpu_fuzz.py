"""Seeded random grid configurations (geometry, cells, levels, boundary conditions, tile offset, variables):
spectralTransform! + gridTransform! through the C ABI vs the oracle.  Exercises the boundaries between kernel
variants (Bluestein length classes, aligned / unaligned Chebyshev tiles, vertical BCs on / off, chunk edges)."""
import numpy as np
import pytest

from helpers import TRANSFORM_TOL, check_transforms
from oracle import chebyshev as ch
from oracle import grids as G
from oracle import splines as spl

pytestmark = pytest.mark.gpu

SPL = [spl.R0, spl.R1T0, spl.R1T1, spl.R1T2, spl.R2T10, spl.R2T20, spl.R3]
CHB = [ch.R0, ch.R1T0, ch.R1T1, ch.R1T2]


def random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    geom = ["R", "RL", "RZ", "RLZ"][seed % 4]
    nv = int(rng.integers(1, 4))
    names = [f"v{i}" for i in range(nv)]
    kw = dict(geometry=geom, vars={n: i + 1 for i, n in enumerate(names)})
    if geom in ("R", "RZ"):
        nc = int(rng.integers(6, 140))        # >= 6: two rank-3 BCs leave free coefficients
        kw.update(xmin=float(rng.uniform(-5, 5)), num_cells=nc)
        kw["xmax"] = kw["xmin"] + float(rng.uniform(1, 50))
        kw["BCL"] = {n: SPL[int(rng.integers(0, 7))] for n in names}
        kw["BCR"] = {n: SPL[int(rng.integers(0, 7))] for n in names}
    else:
        # a tile somewhere inside a patch: rings from (sil-1)*3+1; sizes straddle the 64/65, 128/129, 256/257 class edges
        nc = int(rng.integers(1, 6))
        sil = int(rng.choice([1, 1, 2, 20, 21, 22, 42, 43, 85, 86, 170]))
        kw.update(xmin=float(sil - 1), xmax=float(sil - 1 + nc), num_cells=nc, spectralIndexL=sil)
        if sil == 1:
            nc = max(nc, 2)                # a one-cell patch with a folded BC has fewer than 4 free coefficients
            kw.update(xmax=float(nc), num_cells=nc)
            kw["BCL"] = {n: [spl.R1T0, spl.R1T1][int(rng.integers(0, 2))] for n in names}
    if geom in ("RZ", "RLZ"):
        zd = int(rng.choice([5, 8, 12, 16, 17, 32, 33, 64]))
        kw.update(zmin=0.0, zmax=float(rng.uniform(1, 20)), zDim=zd)
        if rng.random() < 0.5:
            kw["BCB"] = {n: CHB[int(rng.integers(0, 4))] for n in names}
            kw["BCT"] = {n: CHB[int(rng.integers(0, 4))] for n in names}
    return G.GridParameters(**kw)


@pytest.mark.parametrize("seed", range(60))
def test_random_grid_transforms_match_oracle(seed, gpu_lib):
    gp = random_case(seed)
    eB, eP = check_transforms(gp, gpu_lib, seed=seed)
    assert eB <= TRANSFORM_TOL, (gp, eB)
    og = G.createGrid(gp)
    # second derivatives of white noise amplify round-off by kDim^2 (lambda) / zDim^4 (z): scale their tolerance
    tol = [TRANSFORM_TOL] * og.D
    if gp.geometry in ("RL", "RLZ"):
        tol[4] = max(TRANSFORM_TOL, 1e-15 * og.kDim ** 2)
    for d, (e, t) in enumerate(zip(eP, tol)):
        assert e <= t, (gp, d, e)


def test_overconstrained_spline_is_rejected(gpu_lib):
    import scythe_jl_b200 as S
    gp = S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=3, BCL={"u": S.CubicBSpline.R3}, BCR={"u": S.CubicBSpline.R3},
                          vars={"u": 1})
    with pytest.raises(S.ScytheError, match="too few cells"):
        S.createGrid(gp, lib=gpu_lib)
