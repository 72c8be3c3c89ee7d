"""Oracle: azimuthal Fourier rings (the ``Fourier`` surface behind the RL/RLZ grids).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (Springsteel.jl absent).
Conventions restated (FFTW R2HC / HC2R):

* ring index ``ri`` (1-based over the PATCH's mish radii) has ``yDim = 4 + 4*ri`` points,
  uniform in lambda, first point at ``ymin = 0.5*dl*(ri-1)``, ``dl = 2*pi/yDim``;
  retained wavenumbers ``k = 0..ri``  (BASELINE.json "~1000 radial x ~4000 max azimuthal").
* forward (FB+FA): ``c_k = (1/yDim) * sum_j u_j exp(-i k lambda_j)``  -- i.e. the R2HC
  coefficient divided by yDim and phase-rotated to the absolute lambda origin.
* inverse (FI): ``u_j = c_0 + 2 * sum_{k>=1} Re(c_k exp(+i k lambda_j))`` (HC2R of the
  zero-padded, un-rotated spectrum).  FIx / FIxx multiply by ``ik`` / ``-k^2``; the
  angular derivative is RAW (equation sets divide by r themselves,
  /root/reference/src/shallowWaterModels.jl:72).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as sfft


def ring_points(ri: int) -> int:
    return 4 + 4 * ri


def ring_ymin(ri: int) -> float:
    n = ring_points(ri)
    return 0.5 * (2.0 * math.pi / n) * (ri - 1)


def ring_lambdas(ri: int) -> np.ndarray:
    n = ring_points(ri)
    return ring_ymin(ri) + (2.0 * math.pi / n) * np.arange(n)


def ring_phase(ri: int) -> np.ndarray:
    """exp(-i k ymin), k = 0..ri, with the angle reduced exactly: k*ymin = pi * (k (ri-1) mod 2n) / n."""
    n = ring_points(ri)
    q = (np.arange(ri + 1, dtype=np.int64) * (ri - 1)) % (2 * n)
    return np.exp(-1j * math.pi * q / n)


def ring_forward(u: np.ndarray, ri: int, workers: int = 1) -> np.ndarray:
    """u: [yDim, ...] real -> c: [ri+1, ...] complex (FB then FA)."""
    n = ring_points(ri)
    assert u.shape[0] == n
    X = sfft.rfft(u, axis=0, workers=workers)[: ri + 1] / n
    return X * ring_phase(ri).reshape((-1,) + (1,) * (u.ndim - 1))


def ring_inverse(c: np.ndarray, ri: int, derivative: int = 0, workers: int = 1) -> np.ndarray:
    """c: [ri+1, ...] complex -> u^(derivative): [yDim, ...] real."""
    n = ring_points(ri)
    k = np.arange(ri + 1).reshape((-1,) + (1,) * (c.ndim - 1))
    Y = c * np.conj(ring_phase(ri)).reshape(k.shape)
    Y[0] = Y[0].real  # wavenumber 0 is real by construction
    if derivative == 1:
        Y = Y * (1j * k)
    elif derivative == 2:
        Y = Y * (-(k.astype(np.float64) ** 2))
    full = np.zeros((n // 2 + 1,) + c.shape[1:], dtype=np.complex128)
    full[: ri + 1] = Y
    return sfft.irfft(full, n=n, axis=0, workers=workers) * n
