"""Oracle: R / RL / RZ / RLZ spectral grids, transforms and the tile/patch decomposition.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the transform algebra (Springsteel.jl absent,
SURVEY App. A); the *surface* follows the reference call sites:

* ``GridParameters`` fields + derived dims   -- /root/reference/src/spectralGrid.jl:20-45
* ``createGrid`` factory / DomainError        -- /root/reference/src/spectralGrid.jl:63-94
* ``spectralTransform!``, ``gridTransform!``    -- /root/reference/src/semiimplicit.jl:135-136,734
* ``splineTransform!(patchSplines, patchSpectral, gp, sharedSpectral, tile)`` -- :237,285
* ``tileTransform!(patchSplines, patchSpectral, gp, tile, splineBuffer)``     -- :241,252,290,305
* ``calcTileSizes`` rows [xmin; xmax; num_cells; spectralIndexL; npts]       -- :141-144,155-169
* ``calcPatchMap`` / ``calcHaloMap``            -- :79-86, used :320-329
* ``getGridpoints`` shapes, ``num_columns``     -- :59,308; /root/reference/src/shallowWaterModels.jl:366-368
* physical[N,V,D] slot order, z fastest (C2), spectral layout (App. A.3)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np
import scipy.linalg as sla

from . import chebyshev as cheb
from . import fourier
from . import splines as spl

GEOMETRIES = ("R", "RZ", "RL", "RLZ")


class DomainError(ValueError):
    """Mirror of Julia's DomainError used by createGrid / calcTileSizes."""


@dataclass
class GridParameters:
    geometry: str = "R"
    xmin: float = 0.0
    xmax: float = 0.0
    num_cells: int = 0
    l_q: float = 2.0
    BCL: dict = field(default_factory=lambda: dict(spl.R0))
    BCR: dict = field(default_factory=lambda: dict(spl.R0))
    zmin: float = 0.0
    zmax: float = 0.0
    zDim: int = 0
    b_zDim: int = -1
    BCB: dict = field(default_factory=lambda: dict(cheb.R0))
    BCT: dict = field(default_factory=lambda: dict(cheb.R0))
    vars: dict = field(default_factory=lambda: {"u": 1})
    spectralIndexL: int = 1
    tile_num: int = 0

    def __post_init__(self):
        if self.b_zDim < 0:
            self.b_zDim = cheb.default_b_zDim(self.zDim) if self.zDim > 0 else 0

    # derived (src/spectralGrid.jl:24-43)
    @property
    def rDim(self) -> int:
        return self.num_cells * spl.MUBAR

    @property
    def b_rDim(self) -> int:
        return self.num_cells + 3

    @property
    def spectralIndexR(self) -> int:
        return self.spectralIndexL + self.b_rDim - 1

    @property
    def patchOffsetL(self) -> int:
        return (self.spectralIndexL - 1) * 3

    @property
    def patchOffsetR(self) -> int:
        return self.patchOffsetL + self.rDim

    def var_bc(self, which: str, name: str) -> dict:
        d = getattr(self, which)
        if name in d and isinstance(d[name], dict):
            return d[name]
        return d  # a single BC dict shared by all variables

    def var_names(self) -> list[str]:
        return [k for k, _ in sorted(self.vars.items(), key=lambda kv: kv[1])]


class Grid:
    """One spectral grid (a patch or a tile) of any supported geometry."""

    def __init__(self, gp: GridParameters):
        if gp.geometry == "Z":
            raise DomainError("Z column model not implemented yet")
        if gp.geometry not in GEOMETRIES:
            raise DomainError("Unknown geometry")
        self.params = gp
        self.has_l = gp.geometry in ("RL", "RLZ")
        self.has_z = gp.geometry in ("RZ", "RLZ")
        self.V = len(gp.vars)
        self.rDim = gp.rDim
        self.b_rDim = gp.b_rDim
        self.zDim = gp.zDim if self.has_z else 1
        self.b_zDim = gp.b_zDim if self.has_z else 1
        self.kDim = (gp.rDim + gp.patchOffsetL) if self.has_l else 0
        self.ncolp = 1 + 2 * self.kDim
        self.D = 3 + (2 if self.has_l else 0) + (2 if self.has_z else 0)
        self.ri = np.arange(1, self.rDim + 1) + gp.patchOffsetL
        self.ring_n = (4 + 4 * self.ri) if self.has_l else np.ones(self.rDim, dtype=np.int64)
        self.ring_off = np.concatenate([[0], np.cumsum(self.ring_n)])  # in horizontal points
        self.lDim = int(self.ring_off[-1]) if self.has_l else 0
        self.hpoints = int(self.ring_off[-1])
        self.N = self.hpoints * self.zDim
        self.S = self.b_zDim * self.b_rDim * self.ncolp
        names = gp.var_names()
        self.splines = [spl.Spline1D(spl.SplineParameters(
            xmin=gp.xmin, xmax=gp.xmax, num_cells=gp.num_cells, l_q=gp.l_q,
            BCL=gp.var_bc("BCL", n), BCR=gp.var_bc("BCR", n))) for n in names]
        self.columns = None
        if self.has_z:
            self.columns = [cheb.Chebyshev1D(cheb.ChebyshevParameters(
                zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, bDim=gp.b_zDim,
                BCB=gp.var_bc("BCB", n), BCT=gp.var_bc("BCT", n))) for n in names]
        self.physical = np.zeros((self.N, self.V, self.D))
        self.spectral = np.zeros((self.S, self.V))
        self.workers = 1

    # ---------------------------------------------------------------- geometry helpers
    def num_columns(self) -> int:
        return self.hpoints if self.has_z else 0

    def getGridpoints(self) -> np.ndarray:
        r = self.splines[0].mishPoints
        if self.params.geometry == "R":
            return r.copy()
        rr = np.repeat(r, self.ring_n)
        cols = [rr]
        if self.has_l:
            cols.append(np.concatenate([fourier.ring_lambdas(int(ri)) for ri in self.ri]))
        if self.has_z:
            z = self.columns[0].mishPoints
            cols = [np.repeat(c, self.zDim) for c in cols]
            cols.append(np.tile(z, self.hpoints))
        return np.stack(cols, axis=1)

    # ---------------------------------------------------------------- forward (K1)
    def _forward_rings(self, u: np.ndarray, v: int) -> np.ndarray:
        """u [N] -> F [rDim, ncolp, b_zDim]: z-modes and retained ring coefficients per radius."""
        cols = u.reshape(self.hpoints, self.zDim)
        if self.has_z:
            zb = self.columns[v].CBtransform(cols.T, workers=self.workers).T  # [hpoints, bz]
        else:
            zb = cols
        F = np.zeros((self.rDim, self.ncolp, self.b_zDim))
        if not self.has_l:
            F[:, 0, :] = zb
            return F
        for r in range(self.rDim):
            ri = int(self.ri[r])
            c = fourier.ring_forward(zb[self.ring_off[r]:self.ring_off[r + 1]], ri, workers=self.workers)
            F[r, 0] = c[0].real
            F[r, 1:2 * ri:2] = c[1:].real
            F[r, 2:2 * ri + 1:2] = c[1:].imag
        return F

    def spectralTransform(self, physical: np.ndarray | None = None, spectral: np.ndarray | None = None):
        """physical[:, v, 0] -> spectral[:, v]  (B: spline inner products, no solve)."""
        physical = self.physical if physical is None else physical
        spectral = self.spectral if spectral is None else spectral
        for v in range(self.V):
            F = self._forward_rings(physical[:, v, 0], v)
            B = self.splines[v].SBtransform(F.reshape(self.rDim, -1))  # [b_rDim, ncolp*bz]
            B = B.reshape(self.b_rDim, self.ncolp, self.b_zDim)
            spectral[:, v] = B.transpose(2, 1, 0).reshape(-1)
        return spectral

    # ---------------------------------------------------------------- A-solve (K2)
    def spline_solve(self, B: np.ndarray, v: int) -> np.ndarray:
        """B [b_rDim, ncols] -> A [b_rDim, ncols] with variable v's BCs and filter."""
        return self.splines[v].SAtransform(B)

    # ---------------------------------------------------------------- inverse (K3)
    def _inverse_from_A(self, A: np.ndarray, patch_splines, tile: "Grid", physical: np.ndarray):
        """A: patch spectral [S_patch, V] (self is the PATCH) -> tile.physical[:, :, :]."""
        t = tile
        off = t.params.patchOffsetL - self.params.patchOffsetL
        rows = slice(off, off + t.rDim)
        for v in range(self.V):
            Av = A[:, v].reshape(self.b_zDim, self.ncolp, self.b_rDim)[:, :t.ncolp, :]
            Av = Av.transpose(2, 1, 0).reshape(self.b_rDim, -1)  # [m, p*bz]
            fields = [patch_splines[v].SItransform(Av, d, rows).reshape(t.rDim, t.ncolp, t.b_zDim)
                      for d in range(3)]  # value, d/dr, d2/dr2 as ring/z spectra
            outs = t._inverse_rings(fields, v)
            for d, o in enumerate(outs):
                physical[:, v, d] = o

    def _inverse_rings(self, fields, v: int):
        """fields: 3 x [rDim, ncolp, bz] -> list of D arrays [N] in slot order."""
        if self.has_l:
            hp = [np.zeros((self.hpoints, self.b_zDim)) for _ in range(5)]
            for r in range(self.rDim):
                ri = int(self.ri[r])
                s = slice(self.ring_off[r], self.ring_off[r + 1])
                cs = []
                for f in fields:
                    c = np.zeros((ri + 1, self.b_zDim), dtype=np.complex128)
                    c[0] = f[r, 0]
                    c[1:] = f[r, 1:2 * ri:2] + 1j * f[r, 2:2 * ri + 1:2]
                    cs.append(c)
                hp[0][s] = fourier.ring_inverse(cs[0], ri, 0, self.workers)
                hp[1][s] = fourier.ring_inverse(cs[1], ri, 0, self.workers)
                hp[2][s] = fourier.ring_inverse(cs[2], ri, 0, self.workers)
                hp[3][s] = fourier.ring_inverse(cs[0], ri, 1, self.workers)
                hp[4][s] = fourier.ring_inverse(cs[0], ri, 2, self.workers)
        else:
            hp = [f[:, 0, :] for f in fields]
        if not self.has_z:
            return [h.reshape(-1) for h in hp]
        col = self.columns[v]
        outs = []
        a_val = None
        for i, h in enumerate(hp):
            a = col.CAtransform(h.T)  # [zDim, hpoints]
            if i == 0:
                a_val = a
            outs.append(col.CItransform(a).T.reshape(-1))
        outs.append(col.CIxtransform(a_val).T.reshape(-1))
        outs.append(col.CIxxtransform(a_val).T.reshape(-1))
        return outs

    def gridTransform(self):
        """spectral (B) -> A (BCs + filter) -> physical[:, :, 0:D]  (patch only).

        ``spectral`` keeps B: run_model copies patch.spectral into sharedSpectral AFTER
        gridTransform! and still treats it as B (/root/reference/src/semiimplicit.jl:135-136,233-237).
        """
        A = np.empty_like(self.spectral)
        for v in range(self.V):
            Bv = self.spectral[:, v].reshape(-1, self.b_rDim).T
            A[:, v] = self.spline_solve(Bv, v).T.reshape(-1)
        self._inverse_from_A(A, self.splines, self, self.physical)
        return self.physical


def createGrid(gp: GridParameters) -> Grid:
    return Grid(gp)


def spectralTransform(grid: Grid):
    return grid.spectralTransform()


def gridTransform(grid: Grid):
    return grid.gridTransform()


def splineTransform(patchSplines, patchSpectral: np.ndarray, pp: GridParameters,
                    sharedSpectral: np.ndarray, tile: Grid | None = None):
    """B (shared, patch sized) -> A (patchSpectral), every spline column of the patch."""
    b_rDim = pp.b_rDim
    for v in range(patchSpectral.shape[1]):
        Bv = sharedSpectral[:, v].reshape(-1, b_rDim).T
        patchSpectral[:, v] = patchSplines[v].SAtransform(Bv).T.reshape(-1)
    return patchSpectral


def tileTransform(patchSplines, patchSpectral: np.ndarray, pp: GridParameters, tile: Grid,
                  splineBuffer=None, patch: Grid | None = None):
    """patch A -> tile.physical evaluated at the tile's own points."""
    if patch is None:
        patch = _PatchView(pp, patchSpectral.shape[1])
    patch._inverse_from_A(patchSpectral, patchSplines, tile, tile.physical)
    return tile.physical


class _PatchView(Grid):
    """Dimension-only view of the patch (no arrays/splines) used by tileTransform."""

    def __init__(self, gp: GridParameters, V: int):  # noqa: D401 - lightweight ctor
        self.params = gp
        self.has_l = gp.geometry in ("RL", "RLZ")
        self.has_z = gp.geometry in ("RZ", "RLZ")
        self.V = V
        self.rDim = gp.rDim
        self.b_rDim = gp.b_rDim
        self.zDim = gp.zDim if self.has_z else 1
        self.b_zDim = gp.b_zDim if self.has_z else 1
        self.kDim = (gp.rDim + gp.patchOffsetL) if self.has_l else 0
        self.ncolp = 1 + 2 * self.kDim


def allocateSplineBuffer(patch: Grid, tile: Grid):
    return np.zeros((tile.rDim, 3))


# -------------------------------------------------------------------- tiles
def _points_per_cell(patch: Grid) -> np.ndarray:
    n = patch.ring_n.reshape(patch.params.num_cells, spl.MUBAR).sum(axis=1)
    return n * patch.zDim


def calcTileSizes(patch: Grid, num_tiles: int) -> np.ndarray:
    """5 x num_tiles matrix [xmin; xmax; num_cells; spectralIndexL; n gridpoints].

    Cuts on cell boundaries, balancing GRID POINTS (not cells); every tile gets >= 3 cells.
    """
    gp = patch.params
    nc = gp.num_cells
    if num_tiles < 1 or nc < 3 * num_tiles:
        raise DomainError("Too many tiles for this grid (need at least 3 cells per tile)")
    ppc = _points_per_cell(patch)
    total = int(ppc.sum())
    cum = np.concatenate([[0], np.cumsum(ppc)])
    DX = (gp.xmax - gp.xmin) / nc
    out = np.zeros((5, num_tiles))
    start = 0
    for t in range(num_tiles):
        remaining_tiles = num_tiles - t - 1
        if remaining_tiles == 0:
            end = nc
        else:
            target = total * (t + 1) / num_tiles
            end = int(np.searchsorted(cum, target, side="left"))
            # choose the boundary closest to the target
            if end > 0 and abs(cum[end - 1] - target) <= abs(cum[min(end, nc)] - target):
                end -= 1
            end = max(end, start + 3)
            end = min(end, nc - 3 * remaining_tiles)
        out[0, t] = gp.xmin + start * DX
        out[1, t] = gp.xmin + end * DX
        out[2, t] = end - start
        out[3, t] = gp.spectralIndexL + start
        out[4, t] = cum[end] - cum[start]
        start = end
    return out


def tile_params(patch: Grid, tp: np.ndarray, t: int) -> GridParameters:
    """GridParameters of tile t (0-based) as built at /root/reference/src/semiimplicit.jl:155-169."""
    gp = patch.params
    names = gp.var_names()
    return replace(
        gp, xmin=float(tp[0, t]), xmax=float(tp[1, t]), num_cells=int(tp[2, t]),
        BCL={k: dict(spl.R0) for k in names}, BCR={k: dict(spl.R0) for k in names},
        spectralIndexL=int(tp[3, t]), tile_num=t + 2)


def _block_rows(grid: Grid, ncolp: int, m0: int, m1: int, shift: int = 0) -> np.ndarray:
    zb = np.arange(grid.b_zDim)[:, None, None]
    p = np.arange(ncolp)[None, :, None]
    m = np.arange(m0, m1)[None, None, :]
    return ((zb * grid.ncolp + p) * grid.b_rDim + m + shift).reshape(-1)


def calcPatchMap(patch: Grid, tile: Grid):
    """(patch rows, tile rows): the tile's OWNED block (all but its last 3 coefficients)."""
    off = tile.params.spectralIndexL - patch.params.spectralIndexL
    nown = tile.b_rDim - 3
    return (_block_rows(patch, tile.ncolp, 0, nown, off), _block_rows(tile, tile.ncolp, 0, nown))


def calcHaloMap(patch: Grid, tile: Grid):
    """(patch rows, tile rows): the tile's last 3 coefficients = next tile's first 3."""
    off = tile.params.spectralIndexL - patch.params.spectralIndexL
    return (_block_rows(patch, tile.ncolp, tile.b_rDim - 3, tile.b_rDim, off),
            _block_rows(tile, tile.ncolp, tile.b_rDim - 3, tile.b_rDim))
