"""Reference-state set-up of the RZ thermodynamic equation sets (Euler_test, BF02_test, rainfall_test): the host side of
`createModelTile` (/root/reference/src/semiimplicit.jl:62-73) -- start-up work, not the hot path.  The scalar closure is
evaluated with NumPy on the `zDim` model levels; the vertical filtering / derivatives / hydrostatic integral go through
the library's Chebyshev column API on the device (`Chebyshev1D`), exactly the calls the reference makes.

* `interpolate_reference_file(model, z)`  -- src/reference_state.jl:17-136: sounding file (first line: surface pressure
  [hPa], theta [K], q_v [g/kg]; then altitude [m], theta, q_v per line), linear interpolation to the model levels,
  hydrostatic integration, re-integration through the Chebyshev column (`CIInttransform`), entropy variables.
* `exact_reference_state(model, z)`       -- src/reference_state.jl:159-199: one line per model level
  ``z sbar xibar mubar mu_lbar``, already in balance.
* `transform_reference_state(model, ref)` -- src/reference_state.jl:138-157: value, d/dz, d2/dz2 without BCs.
* `reference_state_for(model, z)`         -- the choice made at src/semiimplicit.jl:63-73.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .api import Chebyshev, Chebyshev1D, ChebyshevParameters, DomainError, ReferenceState

# constants of src/thermodynamics.jl:2-17
Rd, Rv = 287.04, 461.50
Eps = Rd / Rv
Cvd, Cvv = 716.96, 1410.0
Cpd, Cpv = Cvd + Rd, Cvv + Rv
Cl = 4186.0
gravity = 9.81
L_v0 = 2.501e6
T_0, p_0, q0 = 273.16, 1000.0, 1.0e-7
rho_d0 = 100.0 * p_0 / (T_0 * Rd)
rho_v0 = 100.0 * (6.112 * np.exp(17.67 * (T_0 - 273.15) / ((T_0 - 273.15) + 243.5))) / (T_0 * Rv)


def _L_v(Tk):
    return L_v0 + ((Cpv - Cl) * (Tk - T_0))


def vapor_pressure(p, q_v):                                    # :96-101
    return (p * q_v) / (Eps + q_v)


def entropy(Tk, rho_d, q_v):                                   # :44-54
    q_v = np.asarray(q_v, dtype=np.float64)
    safe = np.where(q_v != 0.0, q_v * rho_d / rho_v0, 1.0)
    qfactor = np.where(q_v != 0.0, q_v * (Rv * np.log(safe) - (_L_v(T_0) / T_0)), 0.0)
    return ((Cvd + (q_v * Cvv)) * np.log(Tk / T_0)) - (Rd * np.log(rho_d / rho_d0)) - qfactor


def bhyp(q_v):                                                 # :196-200
    return 0.5 * ((q_v + q0) - (q0 * q0 / (q_v + q0)))


def ahyp(mu):                                                  # :202-210
    mu = np.asarray(mu, dtype=np.float64)
    return np.where(mu < 0.0, 0.0, np.sqrt(mu * mu + q0 * q0) + mu - q0)


def _P_xi_from_s(s, xi, mu):                                   # :232-236 via thermodynamic_tuple :248-257
    q_v = ahyp(mu)
    rho_d = rho_d0 * np.exp(xi)
    Cfactor = Cvd + (q_v * Cvv)
    safe = np.where(q_v != 0.0, rho_d * q_v / rho_v0, 1.0)
    qfactor = np.where(q_v != 0.0, safe ** ((q_v * Rv) / Cfactor), 1.0)
    Tk = T_0 * np.exp((s - (q_v * _L_v(T_0) / T_0)) / Cfactor) * (rho_d / rho_d0) ** (Rd / Cfactor) * qfactor
    P_s = Tk * ((rho_d * Rd) + (q_v * rho_d * Rv)) / Cfactor
    return (Rd + (q_v * rho_d * Rv)) * ((rho_d * Tk) + P_s)


def _column(model, lib):
    gp = model.grid_params
    return Chebyshev1D(ChebyshevParameters(zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, bDim=gp.b_zDim,
                                           BCB=dict(Chebyshev.R0), BCT=dict(Chebyshev.R0)), lib=lib)


def transform_reference_state(model, ref: np.ndarray, lib=None) -> np.ndarray:
    col = _column(model, lib)
    a = col.CAtransform(col.CBtransform(ref[:, 0]))
    ref[:, 0], ref[:, 1], ref[:, 2] = col.CItransform(a), col.CIxtransform(a), col.CIxxtransform(a)
    return ref


def _finish(model, sbar, xibar, mubar, mu_lbar, lib, transform_liquid: bool) -> ReferenceState:
    for prof in (sbar, xibar, mubar) + ((mu_lbar,) if transform_liquid else ()):
        transform_reference_state(model, prof, lib)
    Pxi = _P_xi_from_s(sbar[:, 0], xibar[:, 0], mubar[:, 0])
    rho_bar = rho_d0 * np.exp(xibar[:, 0])
    q_bar = ahyp(mubar[:, 0])
    return ReferenceState(sbar, xibar, mubar, mu_lbar, float(np.mean(Pxi / (rho_bar * (1.0 + q_bar)))))


def exact_reference_state(model, z: np.ndarray, lib=None) -> ReferenceState:
    n = len(z)
    prof = [np.zeros((n, 3)) for _ in range(4)]
    with open(model.ref_state_file) as f:
        for i in range(n):
            parts = f.readline().split()
            # the reference compares the text with string(z[i]) (:179); a file it accepts has the same number here
            if len(parts) < 5 or float(parts[0]) != float(z[i]):
                raise DomainError(_lib.SB_EDOMAIN, f"DomainError with {i + 1}:\nModel level does not match reference level")
            for k in range(4):
                prof[k][i, 0] = float(parts[1 + k])
    return _finish(model, *prof, lib, transform_liquid=True)


def interpolate_reference_file(model, z: np.ndarray, lib=None) -> ReferenceState:
    with open(model.ref_state_file) as f:
        lines = f.read().split("\n")
    first = lines[0].split()
    sfc_pressure = float(first[0])
    alt, theta_in, q_in = [0.0], [float(first[1])], [float(first[2])]
    for ln in lines[1:]:
        if not ln.strip():                    # `isempty(level)`: the first blank line ends the sounding (:35-37)
            break
        a, th, q = ln.split()[:3]
        alt.append(float(a)); theta_in.append(float(th)); q_in.append(float(q))
    n = len(z)
    theta, q_v = np.zeros(n), np.zeros(n)
    theta[0], q_v[0] = theta_in[0], q_in[0]   # "Assumes first level in both cases is the surface" (:49-51)
    for i in range(1, n):
        found = False
        for j in range(1, len(alt)):          # every matching pair is applied, the last one stands (:55-68)
            if alt[j - 1] < z[i] and alt[j] > z[i]:
                w = (z[i] - alt[j - 1])
                theta[i] = theta_in[j - 1] + w * (theta_in[j] - theta_in[j - 1]) / (alt[j] - alt[j - 1])
                q_v[i] = q_in[j - 1] + w * (q_in[j] - q_in[j - 1]) / (alt[j] - alt[j - 1])
                found = True
            elif alt[j] == z[i]:
                theta[i], q_v[i] = theta_in[j], q_in[j]
                found = True
        if not found:
            raise DomainError(_lib.SB_EDOMAIN, f"DomainError with {i + 1}:\nCan't find an interpolating level for reference state")
    q_v = q_v * 1.0e-3
    Tk, p, rho_d, rho_t = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    p[0] = sfc_pressure
    e = vapor_pressure(p[0], q_v[0])
    Tk[0] = theta[0] / (p_0 / p[0]) ** (Rd / Cpd)
    rho_d[0] = 100.0 * (p[0] - e) / (Tk[0] * Rd)
    rho_t[0] = rho_d[0] * (1.0 + q_v[0])
    dlnpdz = -gravity * rho_t[0] / (p[0] * 100.0)
    for i in range(1, n):                     # first guess: piecewise hydrostatic integration (:85-94)
        p[i] = np.exp(np.log(p[i - 1]) + (dlnpdz * (z[i] - z[i - 1])))
        Tk[i] = theta[i] / (p_0 / p[i]) ** (Rd / Cpd)
        e = vapor_pressure(p[i], q_v[i])
        rho_d[i] = 100.0 * (p[i] - e) / (Tk[i] * Rd)
        rho_t[i] = rho_d[i] * (1.0 + q_v[i])
        dlnpdz = -gravity * rho_t[i] / (p[i] * 100.0)
    col = _column(model, lib)                 # re-integrate with the Chebyshev column to adjust T (:96-109)
    a = col.CAtransform(col.CBtransform(-gravity * rho_t))
    p_new = col.CIInttransform(a, sfc_pressure * 100.0) / 100.0
    Tk = theta / (p_0 / p_new) ** (Rd / Cpd)
    e = vapor_pressure(p_new, q_v)
    rho_d = 100.0 * (p_new - e) / (Tk * Rd)
    sbar, xibar, mubar, mu_lbar = (np.zeros((n, 3)) for _ in range(4))
    sbar[:, 0] = entropy(Tk, rho_d, q_v)
    xibar[:, 0] = np.log(rho_d / rho_d0)
    mubar[:, 0] = bhyp(q_v)
    return _finish(model, sbar, xibar, mubar, mu_lbar, lib, transform_liquid=False)


def reference_state_for(model, z: np.ndarray, lib=None) -> ReferenceState | None:
    """src/semiimplicit.jl:62-73: none without a file; the exact file when options[:exact_reference_state], else the sounding."""
    if not model.ref_state_file:
        return None
    exact = bool(model.options.get("exact_reference_state", model.options.get(":exact_reference_state", False)))
    return exact_reference_state(model, z, lib) if exact else interpolate_reference_file(model, z, lib)
