"""CF-style NetCDF output / input of a spectral grid (the reference lists it under "Future plans",
/root/reference/README.md:40-41, and today writes CSV only: src/io.jl:3-13).

Host-side I/O, not part of the compute path: the file is NetCDF-3 (64-bit offset) written with ``scipy.io.netcdf_file``
(no netCDF4 / HDF5 in the image).  The mish points of the R / RL / RZ / RLZ grids are not a tensor product (ring ``ri``
has ``4 + 4 ri`` points), so the horizontal part is stored the CF way for unstructured data: one ``point`` dimension with
auxiliary coordinate variables ``r`` (, ``lambda``), and ``z`` as a true dimension when the grid has levels:

    dimensions:  time = UNLIMITED, point = hpoints, z = zDim
    variables:   time(time); r(point); lambda(point) [RL, RLZ]; z(z) [RZ, RLZ]; ring(point) [RL, RLZ]
                 <var>(time, point[, z]) for every model variable (attribute ``coordinates``)
                 <var>_<slot>(time, point[, z]) for the derivative slots when ``derivatives=True`` (r, rr, l, ll, z, zz)

``write_grid_netcdf`` appends a record when the file exists (one file per run) -- ``write_output``'s per-time CSV files
(`physical_out_<t>.csv`) keep their reference names in `api.write_grid`.
"""
from __future__ import annotations

import os

import numpy as np
from scipy.io import netcdf_file

from . import api

_SLOTS = {"R": ["", "r", "rr"], "RL": ["", "r", "rr", "l", "ll"], "RZ": ["", "r", "rr", "z", "zz"],
          "RLZ": ["", "r", "rr", "l", "ll", "z", "zz"]}


def _layout(grid):
    geom = grid.params.geometry
    has_z = geom in ("RZ", "RLZ")
    nz = grid.params.zDim if has_z else 1
    pts = api.getGridpoints(grid).reshape(grid.N, -1)
    hp = grid.N // nz
    return geom, has_z, nz, hp, pts


def write_grid_netcdf(grid, path: str, time: float, physical: np.ndarray | None = None, derivatives: bool = False,
                      attrs: dict | None = None) -> str:
    """Append the state at ``time`` to ``path`` (created with the coordinates on the first call)."""
    physical = grid.physical if physical is None else physical
    geom, has_z, nz, hp, pts = _layout(grid)
    names = grid.params.var_names()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    new = not os.path.exists(path)
    f = netcdf_file(path, "w" if new else "a", version=2)
    try:
        if new:
            f.Conventions = "CF-1.8"
            f.title = "Scythe semi-spectral model output"
            f.source = "scythe_jl_b200 (B200 implementation of the Scythe.jl hot path)"
            f.geometry = geom
            f.num_cells = np.int32(grid.params.num_cells)
            f.xmin, f.xmax = float(grid.params.xmin), float(grid.params.xmax)
            for k, v in (attrs or {}).items():
                setattr(f, k, v)
            f.createDimension("time", None)
            f.createDimension("point", hp)
            tv = f.createVariable("time", "f8", ("time",))
            tv.units, tv.long_name, tv.axis = "s", "model time", "T"
            col0 = pts[::nz]                                   # one row per column (z is the fastest index of a column)
            rv = f.createVariable("r", "f8", ("point",))
            rv[:] = col0[:, 0]
            rv.units, rv.long_name = "m", "radius of the mish point"
            coords = "r"
            if geom in ("RL", "RLZ"):
                lv = f.createVariable("lambda", "f8", ("point",))
                lv[:] = col0[:, 1]
                lv.units, lv.long_name = "radian", "azimuth of the mish point"
                r = col0[:, 0]
                ring = np.concatenate([[0], np.cumsum(r[1:] != r[:-1])]).astype(np.int32)
                gv = f.createVariable("ring", "i4", ("point",))
                gv[:] = ring
                gv.long_name = "ring index (ring ri has 4 + 4 ri points)"
                coords = "r lambda"
            if has_z:
                f.createDimension("z", nz)
                zv = f.createVariable("z", "f8", ("z",))
                zv[:] = pts[:nz, -1]
                zv.units, zv.long_name, zv.axis, zv.positive = "m", "height of the Chebyshev level", "Z", "up"
                f.zmin, f.zmax = float(grid.params.zmin), float(grid.params.zmax)
            dims = ("time", "point", "z") if has_z else ("time", "point")
            for n in names:
                for s in (_SLOTS[geom] if derivatives else [""]):
                    v = f.createVariable(n + ("_" + s if s else ""), "f8", dims)
                    v.coordinates = coords
                    v.long_name = n if not s else f"d{len(s)}{n}/d{s[0]}{len(s) if len(s) > 1 else ''}"
        rec = f.variables["time"].shape[0]
        f.variables["time"][rec] = float(time)
        for i, n in enumerate(names):
            for d, s in enumerate(_SLOTS[geom] if derivatives else [""]):
                key = n + ("_" + s if s else "")
                if key in f.variables:
                    a = np.ascontiguousarray(physical[:, i, d])
                    f.variables[key][rec] = a.reshape(hp, nz) if has_z else a
    finally:
        f.close()
    return path


def read_physical_grid_netcdf(path: str, grid, record: int = -1) -> float:
    """Fill ``grid.physical[:, v, 0]`` from record ``record`` of a file written by `write_grid_netcdf`; returns its time."""
    f = netcdf_file(path, "r", mmap=False)
    try:
        t = float(f.variables["time"][record])
        for name, v in grid.params.vars.items():
            grid.physical[:, v - 1, 0] = np.asarray(f.variables[name][record], dtype=np.float64).reshape(-1)
    finally:
        f.close()
    return t
