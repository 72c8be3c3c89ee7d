"""scythe_jl_b200 -- B200 (sm_100a) implementation of the Scythe.jl semi-spectral transform and
time-step hot path behind the reference's API names.  See DESIGN.md / INTEGRATION.md.

Importing the package does not load the CUDA library; the first grid/model creation does, and it
fails loudly if ``libscythe_b200.so`` has not been built (there is no CPU fallback).
"""
from .api import (Chebyshev, Chebyshev1D, ChebyshevParameters, CubicBSpline, DomainError, Grid, GridParameters, Model, ModelParameters,  # noqa: F401
                  ReferenceState, ScytheError, UnsupportedError, allocateSplineBuffer, calcHaloMap, calcPatchMap, calcTileSizes, checkCFL, createGrid, dct_1st_derivative, dct_2nd_derivative, dct_matrix,
                  getGridpoints, gridTransform, integrate_model, num_columns, read_physical_grid,
                  spectralTransform, splineTransform, tileTransform, tile_grid_params, write_grid)
from .modelfile import ModelFileError, load_model_file, parse_model_text  # noqa: F401,E402
from .ncio import read_physical_grid_netcdf, write_grid_netcdf  # noqa: F401,E402
from .reference_state import exact_reference_state, interpolate_reference_file, transform_reference_state  # noqa: F401,E402
