"""ctypes binding of libscythe_b200.so -- one prototype per symbol of include/scythe_b200.h.

The product path has NO fallback: if the sm_100a library has not been built, importing the
symbols raises, and every compute call fails with SB_ECUDA when no CUDA device is visible.
Tests may pass an explicit ``path`` (the CPU emulation build under tests/_emu) -- nothing in
this package ever looks for that file on its own.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libscythe_b200.so"

SB_OK, SB_EINVAL, SB_EDOMAIN, SB_ECUDA, SB_ENAN, SB_EUNSUPPORTED, SB_ECOMM = 0, -1, -2, -3, -4, -5, -6
GEOM = {"R": 0, "RZ": 1, "RL": 2, "RLZ": 3}
SPLINE_BC = {"R0": 0, "R1T0": 1, "R1T1": 2, "R1T2": 3, "R2T10": 4, "R2T20": 5, "R3": 6, "PERIODIC": 7}
CHEB_BC = {"R0": 0, "R1T0": 1, "R1T1": 2, "R1T2": 3}

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class sb_grid_params(C.Structure):
    _fields_ = [("geometry", C.c_int32), ("nvars", C.c_int32), ("xmin", C.c_double), ("xmax", C.c_double),
                ("num_cells", C.c_int64), ("l_q", C.c_double), ("zmin", C.c_double), ("zmax", C.c_double),
                ("zDim", C.c_int64), ("b_zDim", C.c_int64), ("spectralIndexL", C.c_int64), ("tile_num", C.c_int64),
                ("BCL", c_i32p), ("BCR", c_i32p), ("BCB", c_i32p), ("BCT", c_i32p)]


class sb_grid_info(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("N", "V", "D", "S", "rDim", "b_rDim", "zDim", "b_zDim", "kDim", "lDim",
                                         "num_columns", "patchOffsetL", "ndims")]


class sb_model_params(C.Structure):
    _fields_ = [("ts", C.c_double), ("integration_time", C.c_double), ("output_interval", C.c_double),
                ("equation_set", C.c_char_p), ("grid", C.POINTER(sb_grid_params)),
                ("var_names", C.POINTER(C.c_char_p)), ("n_physical_params", C.c_int32),
                ("param_names", C.POINTER(C.c_char_p)), ("param_values", c_f64p), ("semiimplicit", C.c_int32),
                ("ref_sbar", c_f64p), ("ref_xibar", c_f64p), ("ref_mubar", c_f64p), ("Pxi_bar", C.c_double),
                ("ref_mu_lbar", c_f64p)]


class sb_cheb_params(C.Structure):
    _fields_ = [("zmin", C.c_double), ("zmax", C.c_double), ("zDim", C.c_int64), ("b_zDim", C.c_int64),
                ("BCB", C.c_int32), ("BCT", C.c_int32)]


grid_t = C.c_void_p
model_t = C.c_void_p

# every symbol declared in include/scythe_b200.h: name -> (restype, argtypes)
PROTOTYPES = {
    "sb_last_error": (C.c_char_p, []),
    "sb_version": (C.c_char_p, []),
    "sb_device_count": (C.c_int, []),
    "sb_grid_create": (C.c_int, [C.POINTER(sb_grid_params), C.c_int, C.c_void_p, C.POINTER(grid_t)]),
    "sb_grid_destroy": (C.c_int, [grid_t]),
    "sb_grid_get_info": (C.c_int, [grid_t, C.POINTER(sb_grid_info)]),
    "sb_grid_get_gridpoints": (C.c_int, [grid_t, c_f64p, C.c_int64]),
    "sb_grid_set_physical": (C.c_int, [grid_t, c_f64p, C.c_int32, C.c_int32]),
    "sb_grid_get_physical": (C.c_int, [grid_t, c_f64p, C.c_int32, C.c_int32]),
    "sb_grid_set_spectral": (C.c_int, [grid_t, C.c_int32, c_f64p]),
    "sb_grid_get_spectral": (C.c_int, [grid_t, C.c_int32, c_f64p]),
    "sb_spectral_transform": (C.c_int, [grid_t]),
    "sb_grid_transform": (C.c_int, [grid_t]),
    "sb_spline_transform": (C.c_int, [grid_t, grid_t]),
    "sb_tile_transform": (C.c_int, [grid_t, grid_t]),
    "sb_calc_tile_sizes": (C.c_int, [C.POINTER(sb_grid_params), C.c_int32, c_f64p]),
    "sb_shared_clear": (C.c_int, [grid_t]),
    "sb_shared_assemble": (C.c_int, [grid_t, grid_t, grid_t, C.c_int32]),
    "sb_check_cfl": (C.c_int, [grid_t, c_i32p, C.POINTER(C.c_int64)]),
    "sb_grid_sync": (C.c_int, [grid_t]),
    "sb_grid_device_ptr": (C.c_int, [grid_t, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "sb_model_create": (C.c_int, [C.POINTER(sb_model_params), C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_void_p,
                                  C.POINTER(model_t)]),
    "sb_model_destroy": (C.c_int, [model_t]),
    "sb_model_initialize": (C.c_int, [model_t, c_f64p]),
    "sb_model_patch": (C.c_int, [model_t, C.POINTER(grid_t)]),
    "sb_model_tile": (C.c_int, [model_t, C.c_int32, C.POINTER(grid_t)]),
    "sb_model_advance_tiles": (C.c_int, [model_t, C.c_int64]),
    "sb_model_exchange": (C.c_int, [model_t]),
    "sb_model_spline_transform": (C.c_int, [model_t]),
    "sb_model_step": (C.c_int, [model_t, C.c_int64]),
    "sb_model_run": (C.c_int, [model_t, C.c_int64, C.c_int64]),
    "sb_model_output": (C.c_int, [model_t, c_f64p]),
    "sb_model_get_state": (C.c_int, [model_t, C.c_int32, C.c_int32, c_f64p]),
    "sb_model_set_state": (C.c_int, [model_t, C.c_int32, C.c_int32, c_f64p]),
    "sb_model_stage_in": (C.c_int, [model_t, C.c_int32, c_f64p]),
    "sb_model_stage_out": (C.c_int, [model_t, C.c_int32, c_f64p]),
    "sb_model_stage_drain": (C.c_int, [model_t, C.c_int32]),
    "sb_model_tendency": (C.c_int, [model_t]),
    "sb_model_cycle": (C.c_int, [model_t, C.c_int64]),
    "sb_model_physics": (C.c_int, [model_t, C.c_int64]),
    "sb_model_set_k3_slots": (C.c_int, [model_t, C.c_int32]),
    "sb_model_profile": (C.c_int, [model_t, C.c_int32]),
    "sb_model_profile_report": (C.c_int, [model_t, C.c_char_p, C.c_int64]),
    "sb_model_sync": (C.c_int, [model_t]),
    "sb_model_launch_count": (C.c_int64, [model_t]),
    "sb_cheb_mish_points": (C.c_int, [C.POINTER(sb_cheb_params), c_f64p]),
    "sb_cheb_matrices": (C.c_int, [C.POINTER(sb_cheb_params), c_f64p, c_f64p, c_f64p]),
    "sb_cheb_columns": (C.c_int, [C.POINTER(sb_cheb_params), C.c_int32, c_f64p, c_f64p, C.c_int64, C.c_double, C.c_int]),
    "sb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sb_model_colsolve_init": (C.c_int, [model_t, C.c_int32, C.c_int32]),
    "sb_model_colsolve_planes": (C.c_int, [model_t, c_i32p, C.c_int32]),
    "sb_model_colsolve_buffer": (C.c_int, [model_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_int64)]),
    "sb_model_colsolve_solve": (C.c_int, [model_t]),
    "sb_model_colsolve_publish": (C.c_int, [model_t]),
    "sb_model_ipc_handle": (C.c_int, [model_t, C.c_int32, C.c_int32, C.c_void_p]),
    "sb_model_ipc_open": (C.c_int, [model_t, C.c_int32, C.c_int32, C.c_void_p]),
    "sb_model_p2p_enable": (C.c_int, [model_t]),
    "sb_model_comm_init": (C.c_int, [model_t, C.c_void_p, C.c_int32, C.c_int32]),
    "sb_timer_start": (C.c_int, [grid_t]),
    "sb_timer_stop": (C.c_int, [grid_t, C.POINTER(C.c_float)]),
}


class ScytheError(RuntimeError):
    """ErrorException equivalent (src/semiimplicit.jl:745, src/Scythe.jl:40)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class DomainError(ScytheError, ValueError):
    """Julia DomainError (src/spectralGrid.jl:88,91; calcTileSizes)."""


class UnsupportedError(ScytheError, NotImplementedError):
    """The equation set / BC has no CUDA kernel; there is no CPU fallback."""


class Library:
    def __init__(self, path):
        path = Path(path)
        if not path.exists():
            raise ImportError(
                f"{path} is missing: build the sm_100a library with `python -m scythe_jl_b200.build` "
                "(nvcc required).  scythe_jl_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(str(path), mode=C.RTLD_LOCAL)   # LOCAL: the test-only emulation build exports the same C++ symbols
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.dll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, rc: int):
        if rc == SB_OK:
            return
        msg = (self.sb_last_error() or b"").decode()
        if rc == SB_EDOMAIN:
            raise DomainError(rc, msg)
        if rc == SB_EUNSUPPORTED:
            raise UnsupportedError(rc, msg)
        raise ScytheError(rc, msg)


_default: Library | None = None


def load(path=None) -> Library:
    """Load (once) the product library, or an explicitly given build (tests only)."""
    global _default
    if path is not None:
        return Library(path)
    if _default is None:
        _default = Library(LIB_PATH)
    return _default
