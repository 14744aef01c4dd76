"""hala_b200 — B200-native (sm_100a) backend for LIBHALA/hala's sparse iterative-solve hot path.

The product is libhalab200.so (hand-written CUDA behind the C ABI in include/halab200.h) plus the C++ header set
hala_b200/gpu/ that replaces the reference's gpu/ directory. This Python package is the host-side mirror used by the
tests and bench.py. Importing it without the built library fails loudly: there is no CPU fallback.
"""
from . import matgen  # noqa: F401  (numpy only)
from .capi import HalaB200Error, LIB_PATH  # noqa: F401
from .engine import (gpu_device_count, gpu_engine, gpu_vector, gpu_sparse_matrix, make_sparse_matrix,  # noqa: F401
                     gpu_triangular_matrix, make_triangular_matrix, sparse_trsv, sparse_trsm, gpu_ilu, make_ilu,
                     vcopy, axpy, scal, dot, dotu, norm2, asum, vswap, iamax, rot, rotg, rotm, rotmg, gemv, geam, dgmm, tbsv, sparse_gemv,
                     solve_cg, solve_gmres)
