"""Oracle: vertical Chebyshev columns (the ``Chebyshev`` surface Scythe calls).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (Springsteel.jl absent).
Reference call sites that constrain the conventions:

* column API ``CBtransform!/CAtransform!/CItransform!/CIxtransform/CIInttransform``
  -- /root/reference/src/semiimplicit.jl:569-574,593,596; /root/reference/src/shallowWaterModels.jl:424-429,480-482
* ``Chebyshev.dct_matrix(nz)``, ``dct_1st_derivative(nz,L)``, ``dct_2nd_derivative(nz,L)`` map
  the coefficient vector ``a`` to values / d/dz / d2/dz2 at the nz mish points, row 1 =
  bottom, row nz = top -- /root/reference/src/semiimplicit.jl:768-781
* level 1 is the bottom, z fastest -- /root/reference/src/semiimplicit.jl:337-338,
  /root/reference/src/shallowWaterModels.jl:463-465
* ``b_zDim = min(zDim, floor((2 zDim - 1)/3) + 1)`` -- /root/reference/src/spectralGrid.jl:35

Conventions (FFTW REDFT00): Gauss-Lobatto points ``z_j = zmin + (zmax-zmin)/2 (1 - cos(pi j/(N-1)))``;
CB: ``b_k = REDFT00(u)_k / (2(N-1))`` truncated to ``b_zDim``; CA: BC projection + zero fill;
CI: ``u_j = a_0 + (-1)^j a_{N-1} + 2 sum_{k=1}^{N-2} a_k cos(pi j k/(N-1))``.
Vertical BCs (R1T0 u=0, R1T1 u'=0, R1T2 u''=0, per end) are imposed by the minimal-norm
"global coefficient adjustment" inside the retained modes:
``a = b - C^T (C C^T)^-1 C b`` with C the BC functionals restricted to the retained modes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft as sfft
from numpy.polynomial import chebyshev as npcheb

R0 = {"R0": 0}
R1T0 = {"α0": 0.0}
R1T1 = {"α1": 0.0}
R1T2 = {"α2": 0.0}
BC_BY_NAME = {"R0": R0, "R1T0": R1T0, "R1T1": R1T1, "R1T2": R1T2}


def bc_name(bc: dict) -> str:
    for name, d in BC_BY_NAME.items():
        if d == bc:
            return name
    raise ValueError(f"unknown Chebyshev BC {bc}")


def default_b_zDim(zDim: int) -> int:
    return min(zDim, (2 * zDim - 1) // 3 + 1)


@dataclass
class ChebyshevParameters:
    zmin: float = 0.0
    zmax: float = 0.0
    zDim: int = 0
    bDim: int = 0
    BCB: dict = field(default_factory=lambda: R0)
    BCT: dict = field(default_factory=lambda: R0)


def mish_points(cp: ChebyshevParameters) -> np.ndarray:
    j = np.arange(cp.zDim)
    return cp.zmin + 0.5 * (cp.zmax - cp.zmin) * (1.0 - np.cos(math.pi * j / (cp.zDim - 1)))


def _xi(nz: int) -> np.ndarray:
    return np.cos(math.pi * np.arange(nz) / (nz - 1))  # xi_0 = +1 at the bottom


def _scale(nz: int) -> np.ndarray:
    s = np.full(nz, 2.0)
    s[0] = 1.0
    s[nz - 1] = 1.0
    return s


def _cheb_eval_matrix(nz: int, deriv: int) -> np.ndarray:
    """M[j,k] = s_k * d^deriv T_k / d xi^deriv (xi_j)."""
    xi = _xi(nz)
    s = _scale(nz)
    M = np.zeros((nz, nz))
    for k in range(nz):
        c = np.zeros(k + 1)
        c[k] = 1.0
        if deriv:
            c = npcheb.chebder(c, deriv)
        M[:, k] = s[k] * npcheb.chebval(xi, c) if len(c) else 0.0
    return M


def dct_matrix(nz: int) -> np.ndarray:
    j = np.arange(nz)[:, None]
    k = np.arange(nz)[None, :]
    return _scale(nz)[None, :] * np.cos(math.pi * j * k / (nz - 1))


def dct_1st_derivative(nz: int, length: float) -> np.ndarray:
    return (-2.0 / length) * _cheb_eval_matrix(nz, 1)


def dct_2nd_derivative(nz: int, length: float) -> np.ndarray:
    return (4.0 / (length * length)) * _cheb_eval_matrix(nz, 2)


def dct_integral(nz: int, length: float) -> np.ndarray:
    """M[j,k]: contribution of a_k to int_{zmin}^{z_j} u dz  (= 0 at the bottom)."""
    xi = _xi(nz)
    s = _scale(nz)
    M = np.zeros((nz, nz))
    for k in range(nz):
        c = np.zeros(k + 1)
        c[k] = 1.0
        ci = npcheb.chebint(c)
        # int_{zmin}^{z} u dz = (L/2) * int_{xi}^{1} u dxi
        M[:, k] = s[k] * 0.5 * length * (npcheb.chebval(1.0, ci) - npcheb.chebval(xi, ci))
    return M


def bc_rows(cp: ChebyshevParameters) -> np.ndarray:
    """BC functionals on the coefficient vector, restricted to the retained modes."""
    L = cp.zmax - cp.zmin
    mats = {"R1T0": dct_matrix(cp.zDim), "R1T1": dct_1st_derivative(cp.zDim, L),
            "R1T2": dct_2nd_derivative(cp.zDim, L)}
    rows = []
    nb, nt = bc_name(cp.BCB), bc_name(cp.BCT)
    if nb != "R0":
        rows.append(mats[nb][0, : cp.bDim])
    if nt != "R0":
        rows.append(mats[nt][cp.zDim - 1, : cp.bDim])
    return np.array(rows).reshape(len(rows), cp.bDim)


def gamma_matrix(cp: ChebyshevParameters) -> np.ndarray:
    """G [bDim,bDim] with a[:bDim] = (I + G) b."""
    C = bc_rows(cp)
    if C.shape[0] == 0:
        return np.zeros((cp.bDim, cp.bDim))
    return -C.T @ np.linalg.solve(C @ C.T, C)


class Chebyshev1D:
    def __init__(self, cp: ChebyshevParameters):
        if cp.bDim == 0:
            cp = ChebyshevParameters(cp.zmin, cp.zmax, cp.zDim, default_b_zDim(cp.zDim), cp.BCB, cp.BCT)
        self.params = cp
        self.mishPoints = mish_points(cp)
        self.gammaBC = gamma_matrix(cp)
        L = cp.zmax - cp.zmin
        self.dct = dct_matrix(cp.zDim)
        self.dct1 = dct_1st_derivative(cp.zDim, L)
        self.dct2 = dct_2nd_derivative(cp.zDim, L)
        self.dctint = dct_integral(cp.zDim, L)
        self.uMish = np.zeros(cp.zDim)
        self.b = np.zeros(cp.bDim)
        self.a = np.zeros(cp.zDim)

    # functional forms; z is axis 0, extra axes are batch
    def CBtransform(self, u: np.ndarray, workers: int = 1) -> np.ndarray:
        nz = self.params.zDim
        return sfft.dct(u, type=1, axis=0, workers=workers)[: self.params.bDim] / (2.0 * (nz - 1))

    def CAtransform(self, b: np.ndarray) -> np.ndarray:
        a = np.zeros((self.params.zDim,) + b.shape[1:])
        a[: self.params.bDim] = b + np.tensordot(self.gammaBC, b, axes=(1, 0))
        return a

    def CItransform(self, a: np.ndarray) -> np.ndarray:
        return np.tensordot(self.dct, a, axes=(1, 0))

    def CIxtransform(self, a: np.ndarray) -> np.ndarray:
        return np.tensordot(self.dct1, a, axes=(1, 0))

    def CIxxtransform(self, a: np.ndarray) -> np.ndarray:
        return np.tensordot(self.dct2, a, axes=(1, 0))

    def CIInttransform(self, a: np.ndarray, C0: float = 0.0) -> np.ndarray:
        return np.tensordot(self.dctint, a, axes=(1, 0)) + C0
