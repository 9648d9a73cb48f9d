"""roms_b200 -- B200-native (sm_100a) kernels for the ROMS nonlinear main3d hot path.

The product is the C-ABI shared library ``libroms_b200.so`` (include/roms_b200.h);
this package is the thin host-side binding used by tests, bench.py and the
Python mirror of the ROMS_initialize / ROMS_run / ROMS_finalize driver surface.
"""
from .lib import (Lib, Bounds, Params, Context, Config, Driver, default_config, FIELD_NAMES, APP_UPWELLING,  # noqa: F401
                  APP_BENCHMARK,
                  tile_bounds, library_path, comm_unique_id)
