"""CPU checks of the drop-in boundary: the sm_100a library loads, exports every symbol that
include/scythe_b200.h declares, and fails loudly (no CPU fallback) when no device is present."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import scythe_jl_b200 as S
from scythe_jl_b200 import _lib, build
from oracle import grids as G

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "scythe_b200.h").read_text()


@pytest.fixture(scope="module")
def lib():
    build.build()          # nvcc cross-compiles for sm_100a without a GPU
    return _lib.load()


def declared_symbols():
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", HEADER)))


def test_header_and_bindings_agree():
    assert set(declared_symbols()) == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib.dll, name), f"{name} is declared in the header but not exported"
    assert b"sm_100a" in lib.sb_version()


def test_every_entry_point_cites_the_reference():
    # each declaration block in the header names the Scythe.jl interface it replaces
    assert HEADER.count("src/semiimplicit.jl") >= 15 and "src/Scythe.jl" in HEADER and "src/spectralGrid.jl" in HEADER


def test_no_cpu_fallback_without_a_device(lib):
    if lib.sb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(S.ScytheError, match="no CPU fallback"):
        S.createGrid(S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=4), lib=lib)
    mp = S.ModelParameters(ts=1.0, equation_set="LinearAdvection1D", physical_params={"c_0": 1.0, "K": 0.0},
                           grid_params=S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=8))
    with pytest.raises(S.ScytheError, match="no CPU fallback"):
        S.Model(mp, lib=lib)


def test_product_loader_never_finds_the_emulation_build():
    assert _lib.LIB_PATH.name == "libscythe_b200.so" and "_emu" not in str(_lib.LIB_PATH)
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.Library(ROOT / "scythe_jl_b200" / "does_not_exist.so")


@pytest.mark.parametrize("geom,nc,zDim,ntiles", [("R", 100, 0, 2), ("RL", 100, 0, 8), ("RZ", 40, 16, 3),
                                                 ("RLZ", 334, 64, 8), ("RLZ", 24, 8, 8)])
def test_calc_tile_sizes_matches_oracle(lib, geom, nc, zDim, ntiles):
    """calcTileSizes is host-only arithmetic: callable without a GPU."""
    gp = G.GridParameters(geometry=geom, xmin=0.0, xmax=3e5, num_cells=nc, zmin=0, zmax=1e4, zDim=zDim)
    ref = G.calcTileSizes(G._PatchView.__new__(G._PatchView) if False else _dims_only(gp), ntiles)
    got = S.calcTileSizes(S.GridParameters(geometry=geom, xmin=0.0, xmax=3e5, num_cells=nc, zmin=0, zmax=1e4, zDim=zDim),
                          ntiles, lib=lib)
    assert np.array_equal(got[2:], ref[2:])
    assert np.allclose(got[:2], ref[:2], rtol=1e-15, atol=0)
    assert got[2].sum() == nc and (got[2] >= 3).all()
    if geom == "RLZ" and nc == 334:   # equal-gridpoint tiles, inner tiles radially wide
        assert list(got[2].astype(int)) == [118, 49, 38, 32, 27, 26, 23, 21] or abs(int(got[2][0]) - 118) <= 1


def _dims_only(gp):
    """oracle Grid without the O(N) arrays (C4 would allocate GBs)."""
    g = G.Grid.__new__(G.Grid)
    g.params = gp
    g.has_l = gp.geometry in ("RL", "RLZ")
    g.zDim = gp.zDim if gp.geometry in ("RZ", "RLZ") else 1
    ri = np.arange(1, gp.rDim + 1) + gp.patchOffsetL
    g.ring_n = (4 + 4 * ri) if g.has_l else np.ones(gp.rDim, dtype=np.int64)
    return g


def test_too_many_tiles_is_a_domain_error(lib):
    with pytest.raises(S.DomainError):
        S.calcTileSizes(S.GridParameters(geometry="R", xmin=0, xmax=1, num_cells=8), 3, lib=lib)


# ---------------------------------------------------------------- the Julia shim (shim/src/ScytheB200.jl) vs the header
SHIM = (ROOT / "shim" / "src" / "ScytheB200.jl").read_text()


def _split_top(argstr):
    """split a comma list at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in argstr:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _header_arity():
    ar = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(sb_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", HEADER, re.S):
        args = m.group(2).strip()
        ar[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return ar


def test_julia_shim_binds_only_exported_symbols(lib):
    """Julia is not in the image, so the shim cannot run here; what can be checked on CPU is that every `ccall` names a
    symbol the library exports and passes as many arguments as the header declares, and that the three mirrored structs
    have the header's field counts."""
    arity = _header_arity()
    calls = re.findall(r"ccall\(sym\(:(sb_[a-z_0-9]+)\),\s*\w+,\s*\(([^)]*(?:\{[^}]*\}[^)]*)*)\)", SHIM)
    assert len(calls) >= 25
    for name, types in calls:
        assert hasattr(lib.dll, name), f"shim calls {name}, which the library does not export"
        n = len([t for t in _split_top(types) if t])
        assert n == arity[name], f"shim passes {n} arguments to {name}, header declares {arity[name]}"

    def julia_fields(struct):
        body = re.search(r"struct " + struct + r"\b[^\n]*\n(.*?)\nend", SHIM, re.S).group(1)
        return len([ln for ln in body.splitlines() if "::" in ln])

    def c_fields(struct):
        body = re.search(r"typedef struct " + struct + r" \{(.*?)\} " + struct + ";", HEADER, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        return sum(len(stmt.split(",")) for stmt in body.split(";") if stmt.strip())
    assert julia_fields("GridParamsC") == c_fields("sb_grid_params") == 16
    assert julia_fields("GridInfoC") == c_fields("sb_grid_info") == 13
    assert julia_fields("ModelParamsC") == c_fields("sb_model_params") == 15
