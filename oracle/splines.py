"""Oracle: cubic B-spline transform (Ooyama 2002), the ``CubicBSpline`` surface Scythe uses.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: restates the
published algorithm that Springsteel.jl's CubicBSpline module implements; constrained by
the reference call sites:

* BC dictionaries and names  -- /root/reference/models/cha_bell2024/Oneway_ShallowWater_Slab.jl:13-26,
  /root/reference/notebooks/LinearAdvection_example.ipynb:43 (``Dict("PERIODIC"=>0)``, ``Dict("R0"=>0)``)
* mubar = 3 mish points per cell, b_rDim = num_cells+3, l_q = 2.0
  -- /root/reference/src/spectralGrid.jl:24-27
* mish point values -- /root/reference/notebooks/LinearAdvection_example.ipynb:91-116
* ``spectral`` holds B (pre-solve inner products); BCs/filter only in the A-solve
  -- /root/reference/src/semiimplicit.jl:135,233-237,285
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

MUBAR = 3
SQRT35 = math.sqrt(3.0 / 5.0)
GAUSS_POINTS = np.array([-SQRT35, 0.0, SQRT35])
GAUSS_WEIGHTS = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])  # times DX

# Homogeneous boundary-condition descriptors (Ooyama 2002, table of rank-r types)
R0 = {"R0": 0}
R1T0 = {"α1": -4.0, "β1": -1.0}   # u = 0
R1T1 = {"α1": 0.0, "β1": 1.0}     # u' = 0
R1T2 = {"α1": 2.0, "β1": -1.0}    # u'' = 0
R2T10 = {"α2": 1.0, "β2": -0.5}   # u = u' = 0
R2T20 = {"α2": -1.0, "β2": 0.0}   # u = u'' = 0
R3 = {"R3": 0}                    # u = u' = u'' = 0
PERIODIC = {"PERIODIC": 0}

BC_BY_NAME = {"R0": R0, "R1T0": R1T0, "R1T1": R1T1, "R1T2": R1T2,
              "R2T10": R2T10, "R2T20": R2T20, "R3": R3, "PERIODIC": PERIODIC}


def bc_name(bc: dict) -> str:
    for name, d in BC_BY_NAME.items():
        if d == bc:
            return name
    raise ValueError(f"unknown spline BC {bc}")


def bc_rank(bc: dict) -> int:
    if "α1" in bc:
        return 1
    if "α2" in bc:
        return 2
    if "R3" in bc:
        return 3
    return 0


@dataclass
class SplineParameters:
    xmin: float = 0.0
    xmax: float = 0.0
    num_cells: int = 1
    l_q: float = 2.0
    BCL: dict = field(default_factory=lambda: R0)
    BCR: dict = field(default_factory=lambda: R0)

    @property
    def DX(self) -> float:
        return (self.xmax - self.xmin) / self.num_cells

    @property
    def bDim(self) -> int:
        return self.num_cells + 3

    @property
    def mishDim(self) -> int:
        return self.num_cells * MUBAR


def basis(spp: SplineParameters, m: int, x, derivative: int = 0):
    """Cubic B-spline centred on node x_m = xmin + m*DX (m = -1 .. num_cells+1)."""
    x = np.asarray(x, dtype=np.float64)
    DXr = 1.0 / spp.DX
    xm = spp.xmin + m * spp.DX
    delta = (x - xm) * DXr
    z = np.abs(delta)
    sgn = np.where(delta > 0, -1.0, 1.0)
    z2 = 2.0 - z
    z1 = np.maximum(1.0 - z, 0.0)
    inside = z < 2.0
    if derivative == 0:
        b = (z2 ** 3 - 4.0 * z1 ** 3) / 6.0
    elif derivative == 1:
        b = sgn * 3.0 * DXr * (z2 ** 2 - 4.0 * z1 ** 2) / 6.0
    elif derivative == 2:
        b = DXr * DXr * (z2 - 4.0 * z1)
    elif derivative == 3:
        b = sgn * DXr ** 3 * np.where(z > 1.0, 1.0, np.where(z < 1.0, -3.0, 0.0))
    else:
        raise ValueError("derivative must be 0..3")
    return np.where(inside, b, 0.0)


def mish_points(spp: SplineParameters) -> np.ndarray:
    """3 Gauss-Legendre points per cell: centre +- sqrt(3/5)*DX/2."""
    c = spp.xmin + (np.arange(spp.num_cells) + 0.5) * spp.DX
    return (c[:, None] + 0.5 * spp.DX * GAUSS_POINTS[None, :]).reshape(-1)


def mish_weights(spp: SplineParameters) -> np.ndarray:
    return np.tile(spp.DX * GAUSS_WEIGHTS, spp.num_cells)


def basis_matrix(spp: SplineParameters, x: np.ndarray, derivative: int = 0) -> sp.csr_matrix:
    """Sparse [len(x), bDim] matrix of basis(-derivative) values; column j is node m=j-1."""
    x = np.asarray(x, dtype=np.float64)
    cell = np.clip(np.floor((x - spp.xmin) / spp.DX).astype(np.int64), 0, spp.num_cells - 1)
    rows, cols, vals = [], [], []
    idx = np.arange(len(x))
    for o in range(4):  # nodes cell-1 .. cell+2  -> columns cell .. cell+3
        m = cell - 1 + o
        col = m + 1
        xm = spp.xmin + m * spp.DX
        # evaluate with per-row node: inline the formula (vectorised over differing m)
        delta = (x - xm) / spp.DX
        z = np.abs(delta)
        sgn = np.where(delta > 0, -1.0, 1.0)
        z2 = 2.0 - z
        z1 = np.maximum(1.0 - z, 0.0)
        DXr = 1.0 / spp.DX
        if derivative == 0:
            b = (z2 ** 3 - 4.0 * z1 ** 3) / 6.0
        elif derivative == 1:
            b = sgn * 3.0 * DXr * (z2 ** 2 - 4.0 * z1 ** 2) / 6.0
        elif derivative == 2:
            b = DXr * DXr * (z2 - 4.0 * z1)
        else:
            b = sgn * DXr ** 3 * np.where(z > 1.0, 1.0, np.where(z < 1.0, -3.0, 0.0))
        b = np.where(z < 2.0, b, 0.0)
        rows.append(idx)
        cols.append(col)
        vals.append(b)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(len(x), spp.bDim))


def mish_basis_matrix(spp: SplineParameters, derivative: int = 0) -> sp.csr_matrix:
    """basis_matrix at the mish points, evaluated from the exact in-cell Gauss offsets
    ``delta = t_mu + 1 - j`` instead of ``(x - x_m)/DX``: the latter loses ~ (x/DX)*eps in absolute
    terms (1e-13 at the outer radii of the 334-cell grid), which would dominate oracle-vs-GPU
    comparisons.  Mathematically identical to Springsteel's ``basis(sp, m, x, derivative)``."""
    nc = spp.num_cells
    t = 0.5 + 0.5 * GAUSS_POINTS                       # position of the 3 mish points inside a cell
    DXr = 1.0 / spp.DX
    rows, cols, vals = [], [], []
    cell = np.repeat(np.arange(nc), MUBAR)
    idx = np.arange(nc * MUBAR)
    for o in range(4):
        delta = np.tile(t + 1.0 - o, nc)
        z = np.abs(delta)
        sgn = np.where(delta > 0, -1.0, 1.0)
        z2 = 2.0 - z
        z1 = np.maximum(1.0 - z, 0.0)
        if derivative == 0:
            b = (z2 ** 3 - 4.0 * z1 ** 3) / 6.0
        elif derivative == 1:
            b = sgn * 3.0 * DXr * (z2 ** 2 - 4.0 * z1 ** 2) / 6.0
        elif derivative == 2:
            b = DXr * DXr * (z2 - 4.0 * z1)
        else:
            b = sgn * DXr ** 3 * np.where(z > 1.0, 1.0, np.where(z < 1.0, -3.0, 0.0))
        rows.append(idx)
        cols.append(cell + o)
        vals.append(np.where(z < 2.0, b, 0.0))
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(nc * MUBAR, spp.bDim))


def gamma_matrix(spp: SplineParameters) -> np.ndarray:
    """BC fold Gamma [(M - rank), M]:  a = Gamma^T a_free  (rank reduction, Ooyama 2002).

    Left rank 1:  a_{-1} = α1 a_0 + β1 a_1.   Left rank 2:  a_{-1} = α2 a_1, a_0 = β2 a_1.
    Rank 3: a_{-1}=a_0=a_1=0.  Right side mirrored.  PERIODIC: a_{-1}=a_{M-1}, a_M=a_0,
    a_{M+1}=a_1 with M = num_cells free coefficients.
    """
    M = spp.bDim
    if spp.BCL == PERIODIC or spp.BCR == PERIODIC:
        if not (spp.BCL == PERIODIC and spp.BCR == PERIODIC):
            raise ValueError("PERIODIC must be set on both ends")
        nc = spp.num_cells
        G = np.zeros((nc, M))
        for m in range(-1, nc + 2):
            G[m % nc, m + 1] = 1.0
        return G
    rL, rR = bc_rank(spp.BCL), bc_rank(spp.BCR)
    nfree = M - rL - rR
    G = np.zeros((nfree, M))
    for j in range(nfree):
        G[j, j + rL] = 1.0
    if rL == 1:
        G[0, 0] = spp.BCL["α1"]
        G[1, 0] = spp.BCL["β1"]
    elif rL == 2:
        G[0, 0] = spp.BCL["α2"]
        G[0, 1] = spp.BCL["β2"]
    if rR == 1:
        G[nfree - 1, M - 1] = spp.BCR["α1"]
        G[nfree - 2, M - 1] = spp.BCR["β1"]
    elif rR == 2:
        G[nfree - 1, M - 1] = spp.BCR["α2"]
        G[nfree - 1, M - 2] = spp.BCR["β2"]
    return G


def pq_matrix(spp: SplineParameters) -> np.ndarray:
    """P + Q:  P = sum_i w_i phi_m phi_m' (mish quadrature);  Q = eps_q sum_i w_i phi'''_m phi'''_m',
    eps_q = (l_q DX / 2 pi)^6  (sixth-order low-pass, cutoff wavelength l_q*DX)."""
    w = mish_weights(spp)
    B0 = mish_basis_matrix(spp, 0)
    B3 = mish_basis_matrix(spp, 3)
    W = sp.diags(w)
    eps_q = (spp.l_q * spp.DX / (2.0 * math.pi)) ** 6
    PQ = (B0.T @ W @ B0) + eps_q * (B3.T @ W @ B3)
    return np.asarray(PQ.todense())


class Spline1D:
    """One spline column object: SB (inner product), SA (solve), SI/SIx/SIxx (evaluate)."""

    def __init__(self, spp: SplineParameters):
        self.params = spp
        self.mishPoints = mish_points(spp)
        self.weights = mish_weights(spp)
        self.gammaBC = gamma_matrix(spp)
        self.pq = pq_matrix(spp)
        self.pq_folded = self.gammaBC @ self.pq @ self.gammaBC.T
        self.pqFactor = sla.cho_factor(self.pq_folded, lower=True)
        self.B = [mish_basis_matrix(spp, d) for d in range(3)]
        self.uMish = np.zeros(spp.mishDim)
        self.b = np.zeros(spp.bDim)
        self.a = np.zeros(spp.bDim)

    # --- functional forms; all accept [dim] or [dim, ncols] arrays -------------------
    def SBtransform(self, u: np.ndarray) -> np.ndarray:
        if u.ndim == 1:
            return self.B[0].T @ (self.weights * u)
        return self.B[0].T @ (self.weights[:, None] * u)

    def SAtransform(self, b: np.ndarray) -> np.ndarray:
        return self.gammaBC.T @ sla.cho_solve(self.pqFactor, self.gammaBC @ b)

    def SItransform(self, a: np.ndarray, derivative: int = 0, rows: slice | None = None) -> np.ndarray:
        Bm = self.B[derivative]
        if rows is not None:
            Bm = Bm[rows]
        return Bm @ a
