"""Host-side mirror of the Scythe.jl / Springsteel.jl API for the semi-spectral hot path.

Julia is not installed in this image, so the host side above the C ABI is written in Python
with the reference's names, argument meaning and error behaviour (the ``!`` of the mutating
Julia functions is dropped):

==============================  ===========================================================
reference (Julia)               here
==============================  ===========================================================
``GridParameters(; ...)``        :class:`GridParameters`          src/spectralGrid.jl:20-45
``CubicBSpline.R0 ...``          :class:`CubicBSpline` constants  models/cha_bell2024/*.jl
``createGrid(gp)``               :func:`createGrid`               src/semiimplicit.jl:130
``spectralTransform!(grid)``     :func:`spectralTransform`        src/semiimplicit.jl:135,734
``gridTransform!(grid)``         :func:`gridTransform`            src/semiimplicit.jl:136
``splineTransform!(...)``        :func:`splineTransform`          src/semiimplicit.jl:237,285
``tileTransform!(...)``          :func:`tileTransform`            src/semiimplicit.jl:241,305
``calcTileSizes(patch, n)``      :func:`calcTileSizes`            src/semiimplicit.jl:141
``getGridpoints(grid)``          :func:`getGridpoints`            src/semiimplicit.jl:59
``num_columns(grid)``            :func:`num_columns`              src/semiimplicit.jl:308
``ModelParameters(; ...)``       :class:`ModelParameters`         src/Scythe.jl:8-21
``integrate_model(model)``       :func:`integrate_model`          src/Scythe.jl:37-62
``initialize_model/run_model``   :class:`Model`                   src/semiimplicit.jl:126-299
==============================  ===========================================================

Arrays are NumPy float64 in Fortran (column-major) order so that ``grid.physical[i, v, d]``
and ``grid.spectral[s, v]`` index exactly like the Julia arrays (0-based).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field, replace

import numpy as np

from . import _lib
from ._lib import DomainError, ScytheError, UnsupportedError  # noqa: F401  (re-exported)


class CubicBSpline:
    """Radial boundary-condition descriptors, same dictionaries as Springsteel's CubicBSpline."""
    R0 = {"R0": 0}
    R1T0 = {"α1": -4.0, "β1": -1.0}
    R1T1 = {"α1": 0.0, "β1": 1.0}
    R1T2 = {"α1": 2.0, "β1": -1.0}
    R2T10 = {"α2": 1.0, "β2": -0.5}
    R2T20 = {"α2": -1.0, "β2": 0.0}
    R3 = {"R3": 0}
    PERIODIC = {"PERIODIC": 0}
    mubar = 3
    _names = ("R0", "R1T0", "R1T1", "R1T2", "R2T10", "R2T20", "R3", "PERIODIC")

    @classmethod
    def code(cls, bc: dict) -> int:
        for n in cls._names:
            if getattr(cls, n) == bc:
                return _lib.SPLINE_BC[n]
        raise ValueError(f"unknown CubicBSpline boundary condition {bc!r}")


class Chebyshev:
    """Vertical boundary-condition descriptors (Springsteel's Chebyshev module)."""
    R0 = {"R0": 0}
    R1T0 = {"α0": 0.0}
    R1T1 = {"α1": 0.0}
    R1T2 = {"α2": 0.0}
    _names = ("R0", "R1T0", "R1T1", "R1T2")

    @classmethod
    def code(cls, bc: dict) -> int:
        for n in cls._names:
            if getattr(cls, n) == bc:
                return _lib.CHEB_BC[n]
        raise ValueError(f"unknown Chebyshev boundary condition {bc!r}")


@dataclass
class ChebyshevParameters:
    """Springsteel ChebyshevParameters(zmin, zmax, zDim, bDim, BCB, BCT) (src/reference_state.jl:97-104)."""
    zmin: float = 0.0
    zmax: float = 0.0
    zDim: int = 0
    bDim: int = 0
    BCB: dict = field(default_factory=lambda: dict(Chebyshev.R0))
    BCT: dict = field(default_factory=lambda: dict(Chebyshev.R0))


class Chebyshev1D:
    """One vertical column object of the reference (col.uMish / col.b / col.a); the transforms are functional and
    batched over axis 1, and run on the device (sb_cheb_columns).  CBtransform!, CAtransform!, CItransform!,
    CIxtransform, CIxxtransform, CIInttransform: src/semiimplicit.jl:569-574,593,596."""

    def __init__(self, cp: ChebyshevParameters, lib=None, device: int = 0):
        self.lib = lib or _lib.load()
        self.device = device
        self.params = cp
        bdim = cp.bDim if cp.bDim > 0 else min(cp.zDim, (2 * cp.zDim - 1) // 3 + 1)
        self.bDim = bdim
        self._c = _lib.sb_cheb_params(zmin=cp.zmin, zmax=cp.zmax, zDim=cp.zDim, b_zDim=bdim, BCB=Chebyshev.code(cp.BCB),
                                      BCT=Chebyshev.code(cp.BCT))
        self.mishPoints = np.empty(cp.zDim)
        self.lib.check(self.lib.sb_cheb_mish_points(C.byref(self._c), _ptr(self.mishPoints)))
        self.uMish, self.b, self.a = np.zeros(cp.zDim), np.zeros(bdim), np.zeros(cp.zDim)

    def _op(self, op: int, x: np.ndarray, n_in: int, n_out: int, C0: float = 0.0) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        one = x.ndim == 1
        xin = np.asfortranarray(x.reshape(n_in, -1))
        out = np.empty((n_out, xin.shape[1]), order="F")
        self.lib.check(self.lib.sb_cheb_columns(C.byref(self._c), op, _ptr(xin), _ptr(out), xin.shape[1], float(C0), self.device))
        return out[:, 0].copy() if one else out

    def CBtransform(self, u):
        return self._op(0, u, self.params.zDim, self.bDim)

    def CAtransform(self, b):
        return self._op(1, b, self.bDim, self.params.zDim)

    def CItransform(self, a):
        return self._op(2, a, self.params.zDim, self.params.zDim)

    def CIxtransform(self, a):
        return self._op(3, a, self.params.zDim, self.params.zDim)

    def CIxxtransform(self, a):
        return self._op(4, a, self.params.zDim, self.params.zDim)

    def CIInttransform(self, a, C0: float = 0.0):
        return self._op(5, a, self.params.zDim, self.params.zDim, C0)


def _cheb_matrix(which: int, nz: int, length: float, lib=None) -> np.ndarray:
    lib = lib or _lib.load()
    cp = _lib.sb_cheb_params(zmin=0.0, zmax=float(length), zDim=nz, b_zDim=nz, BCB=0, BCT=0)
    mats = [np.empty((nz, nz), order="F") for _ in range(3)]
    lib.check(lib.sb_cheb_matrices(C.byref(cp), *(_ptr(m) for m in mats)))
    return mats[which]


def dct_matrix(nz: int, lib=None) -> np.ndarray:
    """Chebyshev.dct_matrix(nz) (src/semiimplicit.jl:772): coefficients -> values at the mish points, row 1 = bottom."""
    return _cheb_matrix(0, nz, 1.0, lib)


def dct_1st_derivative(nz: int, length: float, lib=None) -> np.ndarray:
    return _cheb_matrix(1, nz, length, lib)


def dct_2nd_derivative(nz: int, length: float, lib=None) -> np.ndarray:
    return _cheb_matrix(2, nz, length, lib)


@dataclass
class GridParameters:
    """src/spectralGrid.jl:20-45; derived fields are properties."""
    geometry: str = "R"
    xmin: float = 0.0
    xmax: float = 0.0
    num_cells: int = 0
    l_q: float = 2.0
    BCL: dict = field(default_factory=lambda: dict(CubicBSpline.R0))
    BCR: dict = field(default_factory=lambda: dict(CubicBSpline.R0))
    zmin: float = 0.0
    zmax: float = 0.0
    zDim: int = 0
    b_zDim: int = -1
    BCB: dict = field(default_factory=lambda: dict(Chebyshev.R0))
    BCT: dict = field(default_factory=lambda: dict(Chebyshev.R0))
    vars: dict = field(default_factory=lambda: {"u": 1})
    spectralIndexL: int = 1
    tile_num: int = 0

    def __post_init__(self):
        if self.b_zDim < 0:
            self.b_zDim = min(self.zDim, (2 * self.zDim - 1) // 3 + 1) if self.zDim > 0 else 0

    @property
    def rDim(self):
        return self.num_cells * CubicBSpline.mubar

    @property
    def b_rDim(self):
        return self.num_cells + 3

    @property
    def spectralIndexR(self):
        return self.spectralIndexL + self.b_rDim - 1

    @property
    def patchOffsetL(self):
        return (self.spectralIndexL - 1) * 3

    @property
    def patchOffsetR(self):
        return self.patchOffsetL + self.rDim

    def var_names(self):
        return [k for k, _ in sorted(self.vars.items(), key=lambda kv: kv[1])]

    def _bc(self, which, name):
        d = getattr(self, which)
        return d[name] if (name in d and isinstance(d[name], dict)) else d


def _c_grid_params(gp: GridParameters):
    """GridParameters -> (sb_grid_params, keep-alive objects)."""
    if gp.geometry == "Z":
        raise DomainError(_lib.SB_EDOMAIN, "Z column model not implemented yet")
    if gp.geometry not in _lib.GEOM:
        raise DomainError(_lib.SB_EDOMAIN, "Unknown geometry")
    names = gp.var_names()
    V = len(names)
    arr = lambda codes: (C.c_int32 * V)(*codes)  # noqa: E731
    bcl = arr([CubicBSpline.code(gp._bc("BCL", n)) for n in names])
    bcr = arr([CubicBSpline.code(gp._bc("BCR", n)) for n in names])
    bcb = arr([Chebyshev.code(gp._bc("BCB", n)) for n in names])
    bct = arr([Chebyshev.code(gp._bc("BCT", n)) for n in names])
    p = _lib.sb_grid_params(
        geometry=_lib.GEOM[gp.geometry], nvars=V, xmin=gp.xmin, xmax=gp.xmax, num_cells=gp.num_cells, l_q=gp.l_q,
        zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, b_zDim=gp.b_zDim if gp.b_zDim > 0 else 0,
        spectralIndexL=gp.spectralIndexL, tile_num=gp.tile_num,
        BCL=C.cast(bcl, _lib.c_i32p), BCR=C.cast(bcr, _lib.c_i32p), BCB=C.cast(bcb, _lib.c_i32p),
        BCT=C.cast(bct, _lib.c_i32p))
    return p, (bcl, bcr, bcb, bct)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_lib.c_f64p)


class Grid:
    """A spectral grid (patch or tile) whose transforms run on the GPU.

    ``physical`` [N,V,D] and ``spectral`` [S,V] are host mirrors (what Julia code indexes);
    the transform functions copy the slots they consume to the device, run the kernels and
    copy the result back -- the reference's host-array semantics.  ``resident=True`` skips the
    copies for callers that drive the device buffers themselves (:class:`Model`, bench).
    """

    def __init__(self, gp: GridParameters, device: int = 0, lib: _lib.Library | None = None, handle=None):
        self.lib = lib or _lib.load()
        self.params = gp
        self._own = handle is None
        if handle is None:
            cp, keep = _c_grid_params(gp)
            h = _lib.grid_t()
            self.lib.check(self.lib.sb_grid_create(C.byref(cp), device, None, C.byref(h)))
            handle = h
        self.handle = handle
        info = _lib.sb_grid_info()
        self.lib.check(self.lib.sb_grid_get_info(self.handle, C.byref(info)))
        self.info = info
        for n, _ in info._fields_:
            setattr(self, n, int(getattr(info, n)))
        self.lDim_h = self.lDim
        self._physical = None
        self._spectral = None
        self.resident = False

    # host mirrors, allocated on first touch
    @property
    def physical(self) -> np.ndarray:
        if self._physical is None:
            self._physical = np.zeros((self.N, self.V, self.D), order="F")
        return self._physical

    @property
    def spectral(self) -> np.ndarray:
        if self._spectral is None:
            self._spectral = np.zeros((self.S, self.V), order="F")
        return self._spectral

    def close(self):
        if self._own and self.handle is not None:
            self.lib.sb_grid_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # explicit device <-> host movement
    def upload_physical(self, slot0=0, nslots=1):
        a = np.asfortranarray(self.physical[:, :, slot0:slot0 + nslots])
        self.lib.check(self.lib.sb_grid_set_physical(self.handle, _ptr(a), slot0, nslots))

    def download_physical(self, slot0=0, nslots=None):
        nslots = self.D - slot0 if nslots is None else nslots
        a = np.empty((self.N, self.V, nslots), order="F")
        self.lib.check(self.lib.sb_grid_get_physical(self.handle, _ptr(a), slot0, nslots))
        self.physical[:, :, slot0:slot0 + nslots] = a
        return self.physical

    def upload_spectral(self, which=0, array=None):
        a = np.asfortranarray(self.spectral if array is None else array)
        self.lib.check(self.lib.sb_grid_set_spectral(self.handle, which, _ptr(a)))

    def download_spectral(self, which=0, out=None):
        a = self.spectral if out is None else out
        assert a.flags.f_contiguous
        self.lib.check(self.lib.sb_grid_get_spectral(self.handle, which, _ptr(a)))
        return a

    def sync(self):
        self.lib.check(self.lib.sb_grid_sync(self.handle))


def createGrid(gp: GridParameters, device: int = 0, lib=None) -> Grid:
    return Grid(gp, device=device, lib=lib)


def getGridpoints(grid: Grid) -> np.ndarray:
    out = np.empty((grid.N, grid.ndims), order="F")
    grid.lib.check(grid.lib.sb_grid_get_gridpoints(grid.handle, _ptr(out), out.size))
    return out[:, 0].copy() if grid.ndims == 1 else out


def num_columns(grid: Grid) -> int:
    return grid.num_columns


def spectralTransform(grid: Grid):
    """physical[:, :, 0] -> spectral (B coefficients)."""
    if not grid.resident:
        grid.upload_physical(0, 1)
    grid.lib.check(grid.lib.sb_spectral_transform(grid.handle))
    if not grid.resident:
        grid.download_spectral(0)
    return grid.spectral


def gridTransform(grid: Grid):
    """spectral (B) -> A -> physical[:, :, 0:D]; ``spectral`` keeps B."""
    if not grid.resident:
        grid.upload_spectral(0)
    grid.lib.check(grid.lib.sb_grid_transform(grid.handle))
    if not grid.resident:
        grid.download_physical(0)
    return grid.physical


def splineTransform(patch: Grid, patchSpectral: np.ndarray | None, sharedSpectral: np.ndarray | None):
    """B (sharedSpectral, patch-sized) -> A (patchSpectral).  Arrays may be None when resident."""
    if sharedSpectral is not None:
        patch.upload_spectral(0, sharedSpectral)
    patch.lib.check(patch.lib.sb_spline_transform(patch.handle, patch.handle))
    if patchSpectral is not None:
        patch.download_spectral(1, patchSpectral)
    return patchSpectral


def tileTransform(patch: Grid, patchSpectral: np.ndarray | None, tile: Grid):
    """patch A -> tile.physical at the tile's own points."""
    if patchSpectral is not None:
        patch.upload_spectral(1, patchSpectral)
    patch.lib.check(patch.lib.sb_tile_transform(patch.handle, tile.handle))
    if not tile.resident:
        tile.download_physical(0)
    return tile.physical


def calcTileSizes(patch, num_tiles: int, lib=None) -> np.ndarray:
    gp = patch.params if isinstance(patch, Grid) else patch
    lib = lib or (patch.lib if isinstance(patch, Grid) else _lib.load())
    cp, keep = _c_grid_params(gp)
    out = np.zeros((5, num_tiles), order="F")
    lib.check(lib.sb_calc_tile_sizes(C.byref(cp), num_tiles, _ptr(out)))
    return out


def _block_rows(grid: Grid, ncolp: int, m0: int, m1: int, shift: int = 0) -> np.ndarray:
    bz = max(grid.b_zDim, 1)
    gcol = 1 + 2 * grid.kDim
    zb = np.arange(bz)[:, None, None]
    p = np.arange(ncolp)[None, :, None]
    m = np.arange(m0, m1)[None, None, :]
    return ((zb * gcol + p) * grid.b_rDim + m + shift).reshape(-1)


def calcPatchMap(patch: Grid, tile: Grid):
    """src/semiimplicit.jl:79-81: (BitMatrix over patch.spectral, view into tile.spectral) of the tile's OWNED block
    (all but its last 3 coefficients of every spline column).  Returned as a boolean [S_patch, V] mask and the
    matching row indices of tile.spectral (``tile.spectral[rows]`` is the reference's tileView).  The device path
    never builds these maps (k_assemble / k_extract apply them on the fly); they are here for host code."""
    off = tile.params.spectralIndexL - patch.params.spectralIndexL
    tcol = 1 + 2 * tile.kDim
    mask = np.zeros((patch.S, patch.V), dtype=bool, order="F")
    mask[_block_rows(patch, tcol, 0, tile.b_rDim - 3, off), :] = True
    return mask, _block_rows(tile, tcol, 0, tile.b_rDim - 3)


def calcHaloMap(patch: Grid, tile: Grid):
    """src/semiimplicit.jl:84-86: the tile's last 3 coefficients of every spline column = the next tile's first 3."""
    off = tile.params.spectralIndexL - patch.params.spectralIndexL
    tcol = 1 + 2 * tile.kDim
    mask = np.zeros((patch.S, patch.V), dtype=bool, order="F")
    mask[_block_rows(patch, tcol, tile.b_rDim - 3, tile.b_rDim, off), :] = True
    return mask, _block_rows(tile, tcol, tile.b_rDim - 3, tile.b_rDim)


def allocateSplineBuffer(patch: Grid, tile: Grid):
    """src/semiimplicit.jl:90: scratch of tileTransform!.  The device keeps its own scratch; nothing to allocate."""
    return None


def tile_grid_params(gp: GridParameters, tile_params: np.ndarray, t: int) -> GridParameters:
    """GridParameters of tile t (0-based), as at src/semiimplicit.jl:155-169."""
    names = gp.var_names()
    return replace(gp, xmin=float(tile_params[0, t]), xmax=float(tile_params[1, t]), num_cells=int(tile_params[2, t]),
                   BCL={k: dict(CubicBSpline.R0) for k in names}, BCR={k: dict(CubicBSpline.R0) for k in names},
                   spectralIndexL=int(tile_params[3, t]), tile_num=t + 2)


def checkCFL(grid: Grid):
    var = C.c_int32(-1)
    idx = C.c_int64(-1)
    rc = grid.lib.sb_check_cfl(grid.handle, C.byref(var), C.byref(idx))
    if rc == _lib.SB_ENAN:
        names = grid.params.var_names()
        raise ScytheError(rc, f"NaN found in variable {names[var.value]} at index{idx.value + 1} ! "
                              "CFL condition likely violated")
    grid.lib.check(rc)


# ---------------------------------------------------------------------------------- model
@dataclass
class ModelParameters:
    """src/Scythe.jl:8-21."""
    ts: float = 0.0
    integration_time: float = 1.0
    output_interval: float = 1.0
    equation_set: str = "LinearAdvection1D"
    initial_conditions: str = "ic.csv"
    output_dir: str = "./output/"
    ref_state_file: str = ""
    grid_params: GridParameters = None
    physical_params: dict = field(default_factory=dict)
    options: dict = field(default_factory=lambda: {"semiimplicit": False, "exact_reference_state": False})


@dataclass
class ReferenceState:
    """src/reference_state.jl:4-10; arrays are [zDim, 3] (value, d/dz, d2/dz2)."""
    sbar: np.ndarray
    xibar: np.ndarray
    mubar: np.ndarray
    mu_lbar: np.ndarray | None = None
    Pxi_bar: float = 0.0


class Model:
    """initialize_model + run_model for the tiles this process owns (one tile per reference worker).

    ``num_tiles`` is the total tile count (the reference's number of workers); this process owns
    ``tile_count`` consecutive tiles starting at ``tile_first`` (all of them by default = several
    tiles emulated on one device).  With ``torch.distributed`` initialised and ``distributed=True``
    each rank owns ``num_tiles / world_size`` tiles and the shared spectral sum is an all-reduce.
    """

    def __init__(self, model: ModelParameters, num_tiles: int = 1, tile_first: int = 0, tile_count: int | None = None,
                 device: int = 0, ref_state: ReferenceState | None = None, lib=None, distributed: bool = False,
                 exchange: str | None = None):
        self.lib = lib or _lib.load()
        self.model = model
        self.num_tiles = num_tiles
        self.dist = None
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            if num_tiles % self.world:
                raise ValueError("num_tiles must be a multiple of the world size")
            per = num_tiles // self.world
            tile_first, tile_count = self.rank * per, per
        tile_count = num_tiles - tile_first if tile_count is None else tile_count
        self.tile_first, self.tile_count = tile_first, tile_count
        gp = model.grid_params
        cgp, keep = _c_grid_params(gp)
        names = gp.var_names()
        cnames = (C.c_char_p * len(names))(*[n.encode() for n in names])
        pk = list(model.physical_params.keys())
        pnames = (C.c_char_p * max(len(pk), 1))(*[str(k).lstrip(":").encode() for k in pk])
        pvals = (C.c_double * max(len(pk), 1))(*[float(model.physical_params[k]) for k in pk])
        semi = bool(model.options.get("semiimplicit", model.options.get(":semiimplicit", False)))
        mp = _lib.sb_model_params(ts=model.ts, integration_time=model.integration_time,
                                  output_interval=model.output_interval, equation_set=model.equation_set.encode(),
                                  grid=C.pointer(cgp), var_names=cnames, n_physical_params=len(pk), param_names=pnames,
                                  param_values=pvals, semiimplicit=int(semi))
        if ref_state is None and model.ref_state_file and gp.geometry in ("RZ", "RLZ"):
            # createModelTile (src/semiimplicit.jl:62-73): the reference state comes from model.ref_state_file, evaluated on
            # the model levels (the reference reads them from tilepoints[1:zDim, 2], i.e. the RZ layout: SURVEY App. E.6)
            from . import reference_state as _rs
            zcol = Chebyshev1D(ChebyshevParameters(zmin=gp.zmin, zmax=gp.zmax, zDim=gp.zDim, bDim=gp.b_zDim), lib=self.lib)
            ref_state = _rs.reference_state_for(model, zcol.mishPoints, lib=self.lib)
        self.ref_state = ref_state
        if ref_state is not None:
            self._ref = [np.asfortranarray(np.asarray(a, dtype=np.float64)) for a in
                         (ref_state.sbar, ref_state.xibar, ref_state.mubar)]
            mp.ref_sbar, mp.ref_xibar, mp.ref_mubar = (_ptr(a) for a in self._ref)
            if ref_state.mu_lbar is not None:
                self._ref.append(np.asfortranarray(np.asarray(ref_state.mu_lbar, dtype=np.float64)))
                mp.ref_mu_lbar = _ptr(self._ref[-1])
            mp.Pxi_bar = float(ref_state.Pxi_bar)
        h = _lib.model_t()
        self.lib.check(self.lib.sb_model_create(C.byref(mp), num_tiles, tile_first, tile_count, device, None, C.byref(h)))
        self.handle = h
        ph = _lib.grid_t()
        self.lib.check(self.lib.sb_model_patch(h, C.byref(ph)))
        self.patch = Grid(gp, lib=self.lib, handle=ph)
        self.patch.resident = True
        self.tile_params = calcTileSizes(gp, num_tiles, lib=self.lib)
        self.tiles = []
        for i in range(tile_count):
            th = _lib.grid_t()
            self.lib.check(self.lib.sb_model_tile(h, i, C.byref(th)))
            g = Grid(tile_grid_params(gp, self.tile_params, tile_first + i), lib=self.lib, handle=th)
            g.resident = True
            self.tiles.append(g)
        self.t = 0
        # exchange: "columns" = plane-distributed spline solve, messages moved by torch.distributed P2P
        #           "columns-native" = same, messages moved by the library's own NCCL communicator
        #           "torch" / "native" = all-reduce of the whole shared B + replicated solve (the reference's scheme)
        default = "torch"
        if self.dist is not None and self.world > 1:
            default = "columns-p2p" if (self.dist.get_backend() == "nccl" and self.world <= 8) else "columns"
        self.exchange = exchange or os.environ.get("SB_EXCHANGE", default)
        self._p2p_is_default = exchange is None and "SB_EXCHANGE" not in os.environ
        self._shared_tensor = None
        self._views = {}
        #           "columns-p2p" / "columns-p2p-native" = same solve, but no messages: the kernels store into the other
        #           GPUs' buffers through CUDA-IPC peer mappings (NVLink); only two rendezvous per step remain
        self.columns = self.exchange in ("columns", "columns-native", "columns-p2p", "columns-p2p-native")
        self.p2p = self.exchange in ("columns-p2p", "columns-p2p-native")
        if self.columns:
            self.lib.check(self.lib.sb_model_colsolve_init(self.handle, self.rank, self.world))
        if self.dist is not None and self.world > 1 and self.exchange in ("native", "columns-native", "columns-p2p-native"):
            self._init_native_comm()
        if self.p2p:
            # Two phases: every rank first maps its peers' buffers, then ALL ranks agree on the outcome, and only
            # then is the peer-store path switched on in the library (sb_model_p2p_enable cannot be undone).  A rank whose
            # cudaIpcOpenMemHandle failed while the others enabled would be written to by kernels it knows nothing about.
            ok, err = True, None
            try:
                self._open_p2p()
            except ScytheError as e:
                ok, err = False, e
            if self.dist is not None and self.world > 1:
                ok = self._all_agree(ok)
            fallback_ok = self._p2p_is_default or os.environ.get("SB_P2P_FALLBACK") == "1"
            if ok:
                self.lib.check(self.lib.sb_model_p2p_enable(self.handle))
            elif fallback_ok:
                # the peer mappings are an optimisation of the default: if CUDA IPC is unavailable on ANY rank, every
                # rank drops to the message form of the same exchange (still NCCL over NVLink, same results)
                self.exchange, self.p2p = "columns", False
            else:
                raise err or ScytheError(_lib.SB_ECUDA, "CUDA IPC peer mapping failed on another rank")

    # -- multi-process plumbing -------------------------------------------------------
    def _init_native_comm(self):
        import torch
        uid = (C.c_ubyte * 128)()
        if self.rank == 0:
            self.lib.check(self.lib.sb_comm_unique_id(uid))
        t = torch.tensor(list(uid), dtype=torch.uint8)
        backend = self.dist.get_backend()
        if backend == "nccl":
            t = t.cuda()
        self.dist.broadcast(t, 0)
        uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.lib.check(self.lib.sb_model_comm_init(self.handle, uid, self.rank, self.world))

    def _all_agree(self, ok: bool) -> bool:
        """True only if `ok` on every rank (MIN all-reduce on the backend's device)."""
        import torch
        dev = f"cuda:{torch.cuda.current_device()}" if self.dist.get_backend() == "nccl" else "cpu"
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        return bool(flag.item() >= 1.0)

    def _open_p2p(self):
        """exchange CUDA IPC handles of the receive buffers / tile A arrays and map the peers' memory (nothing is switched
        on yet: see the two-phase comment in __init__)"""
        lib = self.lib
        self._bar = None
        if self.dist is not None and self.world > 1:
            def handle(what, tile):
                h = (C.c_ubyte * 64)()
                lib.check(lib.sb_model_ipc_handle(self.handle, what, tile, h))
                return bytes(h)
            try:
                mine = (self.rank, handle(0, 0), {t: handle(1, t) for t in range(self.tile_first, self.tile_first + self.tile_count)})
            except ScytheError:
                mine = (self.rank, None, {})        # still take part in the gather so that nobody hangs
            everyone = [None] * self.world
            self.dist.all_gather_object(everyone, mine)
            if any(h is None for _, h, _ in everyone):
                raise ScytheError(_lib.SB_ECUDA, "CUDA IPC handles are not available on every rank")
            if os.environ.get("SB_TEST_P2P_FAIL_RANK") == str(self.rank):     # test hook: an asymmetric mapping failure
                raise ScytheError(_lib.SB_ECUDA, "cudaIpcOpenMemHandle failure injected on this rank (SB_TEST_P2P_FAIL_RANK)")
            for r, hrecv, tiles in everyone:
                if r == self.rank:
                    continue
                lib.check(lib.sb_model_ipc_open(self.handle, 0, r, (C.c_ubyte * 64).from_buffer_copy(hrecv)))
                for t, h in tiles.items():
                    lib.check(lib.sb_model_ipc_open(self.handle, 1, t, (C.c_ubyte * 64).from_buffer_copy(h)))

    def _barrier(self):
        """stream-ordered rendezvous (a one-element all-reduce on the compute stream)"""
        if self.dist is None or self.world == 1:
            return
        import torch
        if self.dist.get_backend() == "nccl" and torch.cuda.current_stream().cuda_stream != 0:
            # the library works on the legacy default stream (sb_model_create(stream = NULL)); the rendezvous only orders the
            # peer stores before the owner's solve if torch enqueues it on that same stream
            raise ScytheError(_lib.SB_EINVAL, "Model exchange must be called with torch's default CUDA stream current")
        if self._bar is None:
            dev = f"cuda:{torch.cuda.current_device()}" if self.dist.get_backend() == "nccl" else "cpu"
            self._bar = torch.zeros(1, dtype=torch.float64, device=dev)
        self.dist.all_reduce(self._bar)

    def _as_tensor(self, ptr: int, n: int):
        """zero-copy torch view of a library buffer (device memory; host memory in the CPU emulation build)"""
        import torch
        if self.dist.get_backend() == "nccl":
            class _Wrap:
                __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
            if n == 0:
                return torch.empty(0, dtype=torch.float64, device=f"cuda:{torch.cuda.current_device()}")
            return torch.as_tensor(_Wrap(), device=f"cuda:{torch.cuda.current_device()}")
        if n == 0:
            return torch.empty(0, dtype=torch.float64)
        buf = (C.c_double * n).from_address(ptr)
        return torch.from_numpy(np.ctypeslib.as_array(buf))

    def _shared_as_tensor(self):
        """torch view of the device-resident shared B buffer (zero copy)."""
        if self._shared_tensor is None:
            ptr, n = C.c_void_p(), C.c_int64()
            self.lib.check(self.lib.sb_grid_device_ptr(self.patch.handle, 1, C.byref(ptr), C.byref(n)))
            self._shared_tensor = self._as_tensor(ptr.value, n.value)
        return self._shared_tensor

    def _msg(self, what: int, tile: int, v: int, peer: int):
        key = (what, tile, v, peer)
        if key not in self._views:
            ptr, n = C.c_void_p(), C.c_int64()
            self.lib.check(self.lib.sb_model_colsolve_buffer(self.handle, what, tile, v, peer, C.byref(ptr), C.byref(n)))
            self._views[key] = self._as_tensor(ptr.value, n.value)
        return self._views[key]

    def _p2p(self, direction: int):
        """one message set of the plane-distributed solve: 0 = tiles' B chunks to the plane owners,
        1 = solved chunks back (torch.distributed batch_isend_irecv = one NCCL group)."""
        dist, V, per = self.dist, self.patch.V, self.num_tiles // self.world
        ops = []
        for t in range(self.num_tiles):
            owner = t // per
            for v in range(V):
                if owner == self.rank:
                    for k in range(self.world):
                        if k == self.rank:
                            continue
                        buf = self._msg(3 if direction else 0, t, v, k)
                        if buf.numel() == 0:
                            continue
                        ops.append(dist.P2POp(dist.irecv if direction else dist.isend, buf, k))
                else:
                    buf = self._msg(2 if direction else 1, t, v, self.rank)
                    if buf.numel() == 0:
                        continue
                    ops.append(dist.P2POp(dist.isend if direction else dist.irecv, buf, owner))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _gather_patch_A(self):
        """output only: every owner's solved planes -> the replicated patch A"""
        self.lib.check(self.lib.sb_model_colsolve_publish(self.handle))
        if self.dist is None or self.world == 1:
            return
        for k in range(self.world):
            for v in range(self.patch.V):
                buf = self._msg(5, 0, v, k)
                if buf.numel():
                    self.dist.broadcast(buf, src=k)

    # -- reference driver ---------------------------------------------------------------
    def initialize(self, ic: np.ndarray):
        """initialize_model: ic is patch.physical[:, :, 0] as [N_patch, V]."""
        ic = np.asfortranarray(np.asarray(ic, dtype=np.float64).reshape(self.patch.N, self.patch.V, order="F"))
        self.lib.check(self.lib.sb_model_initialize(self.handle, _ptr(ic)))
        self.t = 0

    def step(self):
        self.t += 1
        if self.dist is None or self.world == 1:
            self.lib.check(self.lib.sb_model_step(self.handle, self.t))
            return
        if self.exchange in ("columns-native", "columns-p2p-native"):
            self.lib.check(self.lib.sb_model_step(self.handle, self.t))
            return
        self.lib.check(self.lib.sb_model_advance_tiles(self.handle, self.t))
        self._exchange()
        if not self.columns:
            self.lib.check(self.lib.sb_model_spline_transform(self.handle))

    def run(self, nsteps: int):
        if self.dist is None or self.world == 1:
            self.lib.check(self.lib.sb_model_run(self.handle, self.t + 1, nsteps))
            self.t += nsteps
        else:
            for _ in range(nsteps):
                self.step()

    def output(self, to_host: bool = True) -> np.ndarray | None:
        """patch.spectral <- A; tileTransform!(patch); checkCFL  (src/semiimplicit.jl:289-291)."""
        if self.columns:
            self._gather_patch_A()
        if to_host:
            out = np.empty((self.patch.N, self.patch.V, self.patch.D), order="F")
            rc = self.lib.sb_model_output(self.handle, _ptr(out))
        else:
            out, rc = None, self.lib.sb_model_output(self.handle, None)
        if rc == _lib.SB_ENAN:
            raise ScytheError(rc, (self.lib.sb_last_error() or b"").decode())
        self.lib.check(rc)
        return out

    def state(self, tile: int, which: str) -> np.ndarray:
        idx = {"var_np1": 0, "expdot_n": 1, "expdot_nm1": 2, "expdot_nm2": 3,
               "impdot_n": 4, "impdot_nm1": 5, "impdot_nm2": 6}[which]
        g = self.tiles[tile]
        out = np.empty((g.N, g.V), order="F")
        self.lib.check(self.lib.sb_model_get_state(self.handle, tile, idx, _ptr(out)))
        return out

    def set_state(self, tile: int, var_np1: np.ndarray):
        """var_np1 of local tile <- host [N_tile, V] (restart / host-driven stepping)."""
        g = self.tiles[tile]
        a = var_np1 if (var_np1.flags.f_contiguous and var_np1.dtype == np.float64) else np.asfortranarray(var_np1, dtype=np.float64)
        assert a.size == g.N * g.V
        self.lib.check(self.lib.sb_model_set_state(self.handle, tile, 0, _ptr(a)))

    # -- checkpoint / restart (the reference's restart drops the AB3 history) -----------------
    _HIST = {"var_np1": 0, "expdot_nm1": 2, "expdot_nm2": 3, "impdot_nm1": 5, "impdot_nm2": 6}

    def checkpoint_path(self, path) -> str:
        """File this rank writes / reads: ``path`` itself for a single process, ``<stem>.tiles<first>-<last>.npz`` per
        rank in a distributed run (every rank calls checkpoint() with the same path)."""
        path = str(path)
        stem = path[:-4] if path.endswith(".npz") else path
        if self.world > 1:
            stem += f".tiles{self.tile_first}-{self.tile_first + self.tile_count - 1}"
        return stem + ".npz"

    def checkpoint(self, path):
        """Write step counter + var_np1 + AB3 (and semi-implicit) history of the local tiles to ``path`` (.npz), with the
        grid / equation-set identity that restore() checks."""
        semi = bool(self.model.options.get("semiimplicit", self.model.options.get(":semiimplicit", False)))
        keys = [k for k in self._HIST if semi or not k.startswith("impdot")]
        data = {"t": np.int64(self.t), "tile_first": np.int64(self.tile_first), "tile_count": np.int64(self.tile_count),
                "num_tiles": np.int64(self.num_tiles), "equation_set": np.array(self.model.equation_set),
                "V": np.int64(self.patch.V), "N_patch": np.int64(self.patch.N),
                "tile_N": np.array([g.N for g in self.tiles], dtype=np.int64)}
        for i in range(len(self.tiles)):
            for k in keys:
                data[f"{k}_{i}"] = self.state(i, k)
        out = self.checkpoint_path(path)
        np.savez(out, **data)
        return out

    def restore(self, path):
        """Exact restart from :meth:`checkpoint`: the history goes back in, the spectral state (B -> A) is
        rebuilt from var_np1 exactly as the step that produced it did, and stepping continues at t+1."""
        data = np.load(self.checkpoint_path(path))
        if int(data["tile_first"]) != self.tile_first:
            raise ValueError("checkpoint belongs to a different rank / tile range")
        if "num_tiles" in data:     # identity written since round 2
            want = dict(num_tiles=self.num_tiles, tile_count=self.tile_count, V=self.patch.V, N_patch=self.patch.N)
            for k, v in want.items():
                if int(data[k]) != int(v):
                    raise ValueError(f"checkpoint does not match this model: {k} = {int(data[k])}, expected {int(v)}")
            if str(data["equation_set"]) != self.model.equation_set:
                raise ValueError(f"checkpoint is of equation set {data['equation_set']}, this model runs {self.model.equation_set}")
            if [int(n) for n in data["tile_N"]] != [g.N for g in self.tiles]:
                raise ValueError("checkpoint tiles have different sizes from this model's tiles")
        for i in range(len(self.tiles)):
            for k, which in self._HIST.items():
                if f"{k}_{i}" in data:
                    a = np.asfortranarray(data[f"{k}_{i}"], dtype=np.float64)
                    if a.shape != (self.tiles[i].N, self.tiles[i].V):
                        raise ValueError(f"checkpoint array {k}_{i} has shape {a.shape}")
                    self.lib.check(self.lib.sb_model_set_state(self.handle, i, which, _ptr(a)))
        self.lib.check(self.lib.sb_model_tendency(self.handle))
        if self.world == 1 and self.columns:
            self.lib.check(self.lib.sb_model_colsolve_solve(self.handle))
        else:
            self._exchange()
        if not self.columns:
            self.lib.check(self.lib.sb_model_spline_transform(self.handle))
        self.t = int(data["t"])

    def get_state_into(self, tile: int, out: np.ndarray):
        self.lib.check(self.lib.sb_model_get_state(self.handle, tile, 0, _ptr(out)))

    # -- host-driven stepping, pipelined (sb_model_stage_*) ---------------------------------
    def stage_in(self, tile: int, var_np1: np.ndarray):
        """Asynchronous set_state: host [N_tile, V] (Fortran order, float64, ideally page-locked) -> var_np1 through a
        copy stream.  The array must stay alive and unmodified until :meth:`drain`."""
        g = self.tiles[tile]
        if not (var_np1.flags.f_contiguous and var_np1.dtype == np.float64 and var_np1.size == g.N * g.V):
            raise ValueError("stage_in needs a Fortran-contiguous float64 [N_tile, V] array (no hidden copy: the transfer is asynchronous)")
        self.lib.check(self.lib.sb_model_stage_in(self.handle, tile, _ptr(var_np1)))

    def stage_out(self, tile: int, out: np.ndarray):
        """Asynchronous get_state: var_np1 -> host ``out`` [N_tile, V]; valid after :meth:`drain`."""
        g = self.tiles[tile]
        if not (out.flags.f_contiguous and out.flags.writeable and out.dtype == np.float64 and out.size == g.N * g.V):
            raise ValueError("stage_out needs a writable Fortran-contiguous float64 [N_tile, V] array")
        self.lib.check(self.lib.sb_model_stage_out(self.handle, tile, _ptr(out)))

    def drain(self, block: bool = True):
        """Compute stream waits for every staged copy; with ``block`` the host waits too."""
        self.lib.check(self.lib.sb_model_stage_drain(self.handle, int(block)))

    def cycle_host(self, tile_in, tile_out):
        """One model_loop iteration with the state in HOST memory on both sides: every local tile's var_np1 comes from
        ``tile_in[i]`` and the stepped state goes to ``tile_out[i]``.  Returns at once; successive calls overlap their
        copies with each other's kernels.  Results are valid after :meth:`drain`."""
        for i, a in enumerate(tile_in):
            self.stage_in(i, a)
        self.cycle()
        for i, a in enumerate(tile_out):
            self.stage_out(i, a)

    def _exchange(self):
        if self.dist is None or self.world == 1:
            return
        if self.exchange in ("native", "columns-native", "columns-p2p-native"):
            self.lib.check(self.lib.sb_model_exchange(self.handle))
        elif self.p2p:
            self._barrier()
            self.lib.check(self.lib.sb_model_colsolve_solve(self.handle))
            self._barrier()
        elif self.columns:
            self._p2p(0)
            self.lib.check(self.lib.sb_model_colsolve_solve(self.handle))
            self._p2p(1)
        else:
            self.dist.all_reduce(self._shared_as_tensor())

    def cycle(self):
        """One model_loop iteration entered at calcTendency (see sb_model_cycle)."""
        self.t += 1
        if self.dist is None or self.world == 1 or self.exchange in ("native", "columns-native", "columns-p2p-native"):
            self.lib.check(self.lib.sb_model_cycle(self.handle, self.t))
            return
        self.lib.check(self.lib.sb_model_tendency(self.handle))
        self._exchange()
        if not self.columns:
            self.lib.check(self.lib.sb_model_spline_transform(self.handle))
        self.lib.check(self.lib.sb_model_physics(self.handle, self.t))

    def initialize_tiles(self, tile_ics):
        """Tile-parallel initialize_model: each local tile gets its own slice of the initial state
        ([N_tile, V]); K1 per tile, shared sum (all-reduce across ranks), K2.  Equivalent to
        spectralTransform!(patch) because the forward transform is additive across tiles."""
        assert len(tile_ics) == len(self.tiles)
        for i, ic in enumerate(tile_ics):
            self.set_state(i, ic)
        self.lib.check(self.lib.sb_model_tendency(self.handle))
        if self.world == 1 and self.columns:
            self.lib.check(self.lib.sb_model_colsolve_solve(self.handle))
        else:
            self._exchange()
        if not self.columns:
            self.lib.check(self.lib.sb_model_spline_transform(self.handle))
        self.t = 0

    def set_k3_slots(self, mode: str | int = "fused"):
        """Slots the in-step tileTransform! produces: "fused" (default: what the equation-set kernel reads, with the
        equation set + time step fused into the last transform stage where such a kernel is built), "needed" (the
        same slots, separate kernels), "all" (every slot of every variable, the reference's dataflow,
        src/semiimplicit.jl:305) or "needed-poisoned" (test hook: every other slot is NaN).  The state is
        bit-identical in all modes."""
        modes = {"fused": 0, "all": 1, "needed-poisoned": 2, "needed": 3}
        self.lib.check(self.lib.sb_model_set_k3_slots(self.handle, modes.get(mode, mode)))

    def profile(self, on: bool):
        self.lib.check(self.lib.sb_model_profile(self.handle, int(on)))

    def profile_report(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self.lib.check(self.lib.sb_model_profile_report(self.handle, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = {"launches": int(n), "ms": float(ms)}
        return out

    def sync(self):
        self.lib.check(self.lib.sb_model_sync(self.handle))

    def launch_count(self) -> int:
        return int(self.lib.sb_model_launch_count(self.handle))

    def close(self):
        if self.handle is not None:
            for g in self.tiles + [self.patch]:
                g.handle = None
            self.lib.sb_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------- CSV + driver
def read_physical_grid(path: str, grid: Grid):
    """Fill grid.physical[:, v, 0] from a CSV with coordinate columns then one column per variable."""
    with open(path) as f:
        header = f.readline().strip().split(",")
    data = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
    for name, v in grid.params.vars.items():
        grid.physical[:, v - 1, 0] = data[:, header.index(name)]


def write_grid(grid: Grid, output_dir: str, tag: str, physical: np.ndarray | None = None):
    physical = grid.physical if physical is None else physical
    pts = getGridpoints(grid)
    pts = pts.reshape(grid.N, -1)
    coord = ["r", "l", "z"] if grid.params.geometry == "RLZ" else (["r", "z"] if grid.params.geometry == "RZ" else ["r", "l"])
    names = grid.params.var_names()
    cols = [pts[:, i] for i in range(pts.shape[1])] + [physical[:, i, 0] for i in range(len(names))]
    os.makedirs(output_dir, exist_ok=True)
    np.savetxt(os.path.join(output_dir, f"physical_out_{tag}.csv"), np.stack(cols, axis=1), delimiter=",",
               header=",".join(coord[:pts.shape[1]] + names), comments="", fmt="%.17g")


def _write_output(grid: Grid, model: ModelParameters, t: float, physical: np.ndarray):
    """write_output (src/io.jl:3-13): ``physical_out_<t>.csv`` with ``<t> = string(round(t; digits=2))``; with
    ``options[:output_format]`` = "netcdf" / "both" the record is (also) appended to ``<output_dir>/scythe_out.nc``."""
    fmt = str(model.options.get("output_format", model.options.get(":output_format", "csv"))).lower()
    if fmt not in ("csv", "netcdf", "both"):
        raise ScytheError(_lib.SB_EINVAL, f"unknown output_format {fmt!r} (csv, netcdf, both)")
    if fmt in ("csv", "both"):
        write_grid(grid, model.output_dir, str(round(t, 2)), physical)
    if fmt in ("netcdf", "both"):
        from .ncio import write_grid_netcdf
        write_grid_netcdf(grid, os.path.join(model.output_dir, "scythe_out.nc"), t, physical,
                          attrs={"equation_set": model.equation_set, "ts": float(model.ts)})


def integrate_model(model: ModelParameters, num_tiles: int = 1, ic: np.ndarray | None = None,
                    ref_state: ReferenceState | None = None, write: bool = False, device: int = 0, lib=None,
                    distributed: bool = False, checkpoint: str | None = None, checkpoint_interval: float = 0.0,
                    restart: str | None = None):
    """src/Scythe.jl:37-62 + model_loop (src/semiimplicit.jl:258-299).  Returns the final patch.physical.

    ``distributed=True`` (under ``torch.distributed``): one tile per rank (the reference's one tile per worker), every rank
    runs this function, rank 0 writes.  ``checkpoint`` / ``checkpoint_interval`` / ``restart``: exact restart files with the
    AB3 history (`Model.checkpoint`), which the reference's CSV restart loses."""
    if num_tiles < 1:
        raise ScytheError(_lib.SB_EINVAL, "Need to add at least 1 worker process")
    m = Model(model, num_tiles=num_tiles, device=device, ref_state=ref_state, lib=lib, distributed=distributed)
    writer = write and m.rank == 0
    if ic is None:
        if str(model.initial_conditions).endswith(".nc"):
            from .ncio import read_physical_grid_netcdf
            read_physical_grid_netcdf(model.initial_conditions, m.patch)
        else:
            read_physical_grid(model.initial_conditions, m.patch)
        ic = m.patch.physical[:, :, 0]
    m.initialize(ic)
    num_ts = int(round(model.integration_time / model.ts))
    output_int = max(int(round(model.output_interval / model.ts)), 1)
    ckpt_int = max(int(round(checkpoint_interval / model.ts)), 1) if (checkpoint and checkpoint_interval > 0.0) else 0
    t = 0
    if restart:
        m.restore(restart)
        t = m.t
    elif write:
        out0 = m.output()
        if writer:
            _write_output(m.patch, model, 0.0, out0)
    while t < num_ts:
        n = min(output_int - (t % output_int), num_ts - t)
        if ckpt_int:
            n = min(n, ckpt_int - (t % ckpt_int))
        m.run(n)
        t += n
        if t % output_int == 0:
            out = m.output(to_host=write)
            if writer:
                _write_output(m.patch, model, t * model.ts, out)
        if ckpt_int and t % ckpt_int == 0:
            m.checkpoint(checkpoint)
    out = m.output()
    if writer and num_ts % output_int != 0:
        # finalize_model (src/semiimplicit.jl:351-355) writes the final time again; only needed when the loop has not
        _write_output(m.patch, model, model.integration_time, out)
    if checkpoint and not ckpt_int:
        m.checkpoint(checkpoint)
    m.close()
    return out
