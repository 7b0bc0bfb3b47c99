"""exsaddle_b200 -- B200-native implementation of exSaddle's Q2-Q1 Stokes/Lame solve path.

The product is the C-ABI shared library `libexsaddle_b200.so` (include/exsaddle_b200.h, csrc/*.cu,
hand-written sm_100a CUDA).  This package is the thin Python host mirror (ctypes) used by tests and
bench.py; it contains no numerical code and no CPU fallback: without the built library or without a
CUDA device every compute call raises.
"""
from .api import ExSaddle, XsbError, lib, library_path, device_available, pattern_row, prealloc_total, bc_list, \
    mg_level_dims, slab_range, pdist_range, dmda_grid, asm_subdomain, grad_line_tables, slab_layout, comm_unique_id, write_petsc_mat, write_petsc_vec, read_petsc_binary, write_vts, read_vts, MAT_A, MAT_A00, MAT_A01, MAT_A10, MAT_A11, MAT_MP, MAT_A00_MF, MAT_A01_MF, MAT_A10_MF, MAT_MG_LEVEL0  # noqa: F401
from .driver import run_exsaddle, monitor_short  # noqa: F401
