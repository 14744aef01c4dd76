"""Python host-side mirror of the reference's L4 interface for the hot path, over the C ABI (include/halab200.h).

Names, argument order and semantics follow the reference so that tests read like its own
(tests/cuda_blas1_tests.hpp, tests/cuda_sparse_tests.hpp, tests/solvers_tests.hpp):

    gpu_engine(device)                      gpu/hala_gpu_engine.hpp:47-307   load / unload / vector / new_vector / synchronize
    gpu_vector                              gpu/hala_gpu_vector.hpp:49-171   size / resize (contents lost) / load / unload / fill
    gpu_sparse_matrix, make_sparse_matrix   gpu/hala_cuda_sparse_general.hpp:191-401   gemv(trans, alpha, x, beta, y)
    vcopy axpy scal dot dotu norm2 gemv     gpu/hala_gpu_blas1.hpp, gpu/hala_gpu_blas2.hpp:39-62
    sparse_gemv(engine, trans, M, N, ...)   gpu/hala_cuda_sparse_general.hpp:407-419
    solve_cg / solve_gmres                  hex/solvers/hala_solvers_cg.hpp:232-246, hala_solvers_gmres.hpp:127-230 (identity preconditioner)

The C++ twin of this file is the header set hala_b200/gpu/ (source-compatible with the reference's gpu/ directory).
This module is plumbing for tests and bench.py; all arithmetic happens in libhalab200.so.
"""
import ctypes as C
import numpy as np
from . import capi
from .capi import lib, check

_CODE = {np.dtype(np.float32): capi.HB_F32, np.dtype(np.float64): capi.HB_F64,
         np.dtype(np.complex64): capi.HB_C32, np.dtype(np.complex128): capi.HB_C64}
_REAL = {np.dtype(np.float32): np.float32, np.dtype(np.float64): np.float64,
         np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}


def gpu_device_count():
    n = C.c_int(0)
    check(lib.hb_device_count(C.byref(n)), "hb_device_count")
    return n.value


def _host_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class gpu_vector:
    """Owning device array (float32/float64/complex64/complex128/int32)."""

    def __init__(self, engine, dtype, n=0):
        self.engine = engine
        self.dtype = np.dtype(dtype)
        self.ptr = C.c_void_p(None)
        self.num = 0
        if n:
            self._alloc(n)

    def _alloc(self, n):
        p = C.c_void_p()
        check(lib.hb_malloc(self.engine.ctx, int(n) * self.dtype.itemsize, C.byref(p)), "hb_malloc")
        self.ptr, self.num = p, int(n)

    def clear(self):
        if self.ptr and self.ptr.value:
            lib.hb_free(self.engine.ctx, self.ptr)
        self.ptr, self.num = C.c_void_p(None), 0

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    def size(self):
        return self.num

    def __len__(self):
        return self.num

    def device(self):
        return self.engine.device()

    def resize(self, n):
        """Reference semantics: no-op if equal, else free + malloc, contents lost (gpu_vector.hpp:97-101)."""
        if n == self.num:
            return
        self.clear()
        if n:
            self._alloc(n)

    def load(self, cpu):
        cpu = np.ascontiguousarray(cpu, dtype=self.dtype)
        self.resize(cpu.size)
        check(lib.hb_memcpy(self.engine.ctx, self.ptr, _host_ptr(cpu), cpu.nbytes, capi.HB_H2D), "hb_memcpy(H2D)")
        return self

    def unload(self):
        out = np.empty(self.num, dtype=self.dtype)
        if self.num:
            check(lib.hb_memcpy(self.engine.ctx, _host_ptr(out), self.ptr, out.nbytes, capi.HB_D2H), "hb_memcpy(D2H)")
        return out

    def fill(self, value):
        v = np.array([value], dtype=self.dtype)
        code = -1 if self.dtype == np.int32 else _CODE[self.dtype]
        check(lib.hb_fill(self.engine.ctx, code, self.num, _host_ptr(v), self.ptr), "hb_fill")

    def offset(self, elements):
        """Raw device pointer `data() + elements` (the reference passes offset pointers, e.g. gmres:212)."""
        return C.c_void_p(self.ptr.value + elements * self.dtype.itemsize)


class gpu_engine:
    def __init__(self, deviceid=0):
        self.ctx = C.c_void_p()
        check(lib.hb_ctx_create(int(deviceid), C.byref(self.ctx)), "hb_ctx_create")
        self._dev = int(deviceid)

    def __del__(self):
        try:
            if self.ctx:
                lib.hb_ctx_destroy(self.ctx)
        except Exception:
            pass

    def device(self):
        return self._dev

    def synchronize(self):
        check(lib.hb_ctx_sync(self.ctx), "hb_ctx_sync")

    def set_stream(self, stream):
        check(lib.hb_ctx_set_stream(self.ctx, C.c_void_p(int(stream))), "hb_ctx_set_stream")

    def load(self, cpu):
        cpu = np.ascontiguousarray(cpu)
        return gpu_vector(self, cpu.dtype).load(cpu)

    def unload(self, v):
        return v.unload()

    def vector(self, n, value, dtype=np.float64):
        v = gpu_vector(self, dtype, n)
        v.fill(value)
        return v

    def new_vector(self, dtype, n=0):
        return gpu_vector(self, dtype, n)

    def timer_start(self):
        check(lib.hb_timer_start(self.ctx), "hb_timer_start")

    def timer_stop(self):
        """Milliseconds since timer_start(), measured with CUDA events on the engine's stream (synchronises)."""
        ms = C.c_float(0)
        check(lib.hb_timer_stop(self.ctx, C.byref(ms)), "hb_timer_stop")
        return ms.value

    def launch_count(self):
        c = C.c_longlong(0)
        check(lib.hb_ctx_launch_count(self.ctx, C.byref(c)), "hb_ctx_launch_count")
        return c.value


def _dptr(v):
    return v.ptr if hasattr(v, "ptr") else v


def _sc(value, dtype):
    return np.array([value], dtype=dtype)


def _default_n(x, incx, n):
    return (1 + (x.size() - 1) // incx if x.size() else 0) if n < 0 else n


# ---------------------------------------------------------------- BLAS-1 (gpu/hala_gpu_blas1.hpp + gpu_overloads.hpp:75-116)
def vcopy(engine, x, y, incx=1, incy=1, N=-1):
    n = _default_n(x, incx, N)
    if y.size() < 1 + (n - 1) * incy:
        y.resize(1 + (n - 1) * incy)       # pure output is resized, as check_set_size does
    check(lib.hb_copy(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, y.ptr, incy), "hb_copy")


def axpy(engine, alpha, x, y, incx=1, incy=1, N=-1):
    n = _default_n(x, incx, N)
    a = _sc(alpha, x.dtype)
    check(lib.hb_axpy(engine.ctx, _CODE[x.dtype], n, _host_ptr(a), x.ptr, incx, y.ptr, incy), "hb_axpy")


def scal(engine, alpha, x, incx=1, N=-1):
    n = _default_n(x, incx, N)
    a = _sc(alpha, x.dtype)
    check(lib.hb_scal(engine.ctx, _CODE[x.dtype], n, _host_ptr(a), x.ptr, incx), "hb_scal")


def dot(engine, x, y, incx=1, incy=1, N=-1, conjugate=True):
    n = _default_n(x, incx, N)
    res = np.zeros(1, dtype=x.dtype)
    check(lib.hb_dot(engine.ctx, _CODE[x.dtype], 1 if conjugate else 0, n, x.ptr, incx, y.ptr, incy, _host_ptr(res)), "hb_dot")
    return res[0]


def dotu(engine, x, y, incx=1, incy=1, N=-1):
    return dot(engine, x, y, incx, incy, N, conjugate=False)


def norm2(engine, x, incx=1, N=-1):
    n = _default_n(x, incx, N)
    res = np.zeros(1, dtype=_REAL[x.dtype])
    check(lib.hb_nrm2(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, _host_ptr(res)), "hb_nrm2")
    return res[0]


def asum(engine, x, incx=1, N=-1):
    n = _default_n(x, incx, N)
    res = np.zeros(1, dtype=_REAL[x.dtype])
    check(lib.hb_asum(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, _host_ptr(res)), "hb_asum")
    return res[0]


def vswap(engine, x, y, incx=1, incy=1, N=-1):
    n = _default_n(x, incx, N)
    check(lib.hb_swap(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, y.ptr, incy), "hb_swap")


def iamax(engine, x, incx=1, N=-1):
    """0-based index of the first entry with the largest |re| + |im| (reference returns cublas' 1-based result minus one)"""
    n = _default_n(x, incx, N)
    res = C.c_int(0)
    check(lib.hb_iamax(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, C.byref(res)), "hb_iamax")
    return res.value - 1


def rotg(engine, a, b, dtype=np.float64):
    """Givens rotation of (a, b): returns (r, z, c, s) with netlib semantics"""
    dt = np.dtype(dtype)
    va, vb, vs = _sc(a, dt), _sc(b, dt), np.zeros(1, dtype=dt)
    vc = np.zeros(1, dtype=_REAL[dt])
    check(lib.hb_rotg(engine.ctx, _CODE[dt], _host_ptr(va), _host_ptr(vb), _host_ptr(vc), _host_ptr(vs)), "hb_rotg")
    return va[0], vb[0], vc[0], vs[0]


def rot(engine, x, y, c, s, incx=1, incy=1, N=-1):
    n = _default_n(x, incx, N)
    real_s = np.iscomplexobj(np.zeros(1, x.dtype)) and not np.iscomplexobj(s)
    vc = _sc(c, _REAL[x.dtype])
    vs = _sc(s, _REAL[x.dtype] if real_s else x.dtype)
    check(lib.hb_rot(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, y.ptr, incy, _host_ptr(vc), _host_ptr(vs), 1 if real_s else 0), "hb_rot")


def rotmg(engine, d1, d2, x1, y1, param, dtype=np.float64):
    """modified Givens setup; param: numpy array of 5 (updated in place); returns (d1, d2, x1)"""
    dt = np.dtype(dtype)
    v = [_sc(t, dt) for t in (d1, d2, x1, y1)]
    check(lib.hb_rotmg(engine.ctx, _CODE[dt], _host_ptr(v[0]), _host_ptr(v[1]), _host_ptr(v[2]), _host_ptr(v[3]), _host_ptr(param)), "hb_rotmg")
    return v[0][0], v[1][0], v[2][0]


def rotm(engine, x, y, param, incx=1, incy=1, N=-1):
    n = _default_n(x, incx, N)
    prm = np.ascontiguousarray(param, dtype=x.dtype)
    check(lib.hb_rotm(engine.ctx, _CODE[x.dtype], n, x.ptr, incx, y.ptr, incy, _host_ptr(prm)), "hb_rotm")


def gemv(engine, trans, M, N, alpha, A, x, beta, y, lda=-1, incx=1, incy=1):
    lda = M if lda < 0 else lda
    dt = A.dtype
    ny = M if trans in "Nn" else N
    if beta == 0 and y.size() < 1 + (ny - 1) * incy:
        y.resize(1 + (ny - 1) * incy)
    a, b = _sc(alpha, dt), _sc(beta, dt)
    check(lib.hb_gemv(engine.ctx, _CODE[dt], trans.encode(), M, N, _host_ptr(a), _dptr(A), lda, _dptr(x), incx,
                      _host_ptr(b), _dptr(y), incy), "hb_gemv")


def geam(engine, transa, transb, M, N, alpha, A, lda, beta, B, ldb, Cm, ldc):
    if Cm.size() < ldc * N:
        Cm.resize(ldc * N)
    a, b = _sc(alpha, A.dtype), _sc(beta, A.dtype)
    check(lib.hb_geam(engine.ctx, _CODE[A.dtype], transa.encode(), transb.encode(), M, N, _host_ptr(a), A.ptr, lda, _host_ptr(b), B.ptr, ldb,
                      Cm.ptr, ldc), "hb_geam")


def dgmm(engine, side, M, N, A, lda, x, incx, Cm, ldc):
    if Cm.size() < ldc * N:
        Cm.resize(ldc * N)
    check(lib.hb_dgmm(engine.ctx, _CODE[A.dtype], side.encode(), M, N, A.ptr, lda, x.ptr, incx, Cm.ptr, ldc), "hb_dgmm")


def tbsv(engine, uplo, trans, diag, N, k, A, lda, x, incx=1):
    check(lib.hb_tbsv(engine.ctx, _CODE[A.dtype], uplo.encode(), trans.encode(), diag.encode(), N, k, A.ptr, lda, x.ptr, incx), "hb_tbsv")


# ---------------------------------------------------------------- sparse (gpu/hala_cuda_sparse_general.hpp)
class gpu_sparse_matrix:
    """Non-owning CSR view + one-time analysis (the vectors must outlive the matrix, as in the reference)."""

    def __init__(self, engine, rows, cols, nnz, pntr, indx, vals):
        self.engine, self.rows, self.cols, self.nnz = engine, rows, cols, nnz
        self.dtype = vals.dtype
        self._keep = (pntr, indx, vals)
        self.h = C.c_void_p()
        check(lib.hb_csr_create(engine.ctx, _CODE[vals.dtype], rows, cols, nnz, pntr.ptr, indx.ptr, vals.ptr,
                                C.byref(self.h)), "hb_csr_create")

    def __del__(self):
        try:
            if self.h:
                lib.hb_csr_destroy(self.h)
        except Exception:
            pass

    def set_variant(self, variant):
        check(lib.hb_csr_set_variant(self.h, int(variant)), "hb_csr_set_variant")

    def set_transpose_mode(self, mode):
        """how op 'T'/'C' products treat the cached CSR of A^T: 'checked' (default), 'frozen', 'scatter' (halab200.h)"""
        code = {"scatter": 0, "checked": 1, "frozen": 2}.get(mode, mode)
        check(lib.hb_csr_set_transpose_mode(self.h, int(code)), "hb_csr_set_transpose_mode")

    def values_changed(self):
        """tell a 'frozen' matrix that the caller rewrote the value array"""
        check(lib.hb_csr_values_changed(self.h), "hb_csr_values_changed")

    def transpose_info(self):
        m, b, n = C.c_int(0), C.c_int(0), C.c_size_t(0)
        check(lib.hb_csr_transpose_info(self.h, C.byref(m), C.byref(b), C.byref(n)), "hb_csr_transpose_info")
        return {"mode": ("scatter", "checked", "frozen")[m.value], "built": bool(b.value), "bytes": n.value}

    def max_row_nnz(self):
        m = C.c_int(0)
        check(lib.hb_csr_info(self.h, None, None, None, None, C.byref(m)), "hb_csr_info")
        return m.value

    def gemv_buffer_size(self, trans="N"):
        b = C.c_size_t(0)
        check(lib.hb_spmv_buffer_size(self.h, trans.encode(), C.byref(b)), "hb_spmv_buffer_size")
        return b.value

    def gemm(self, transa, transb, b_rows, b_cols, alpha, B, ldb, beta, Cm, ldc, work=None):
        """C = alpha op(A) op(B) + beta C (reference gpu_sparse_matrix::gemm, :302-332)"""
        N = b_cols if transb in "Nn" else b_rows
        if beta == 0 and Cm.size() < ldc * N:
            Cm.resize(ldc * N)
        a, b = _sc(alpha, self.dtype), _sc(beta, self.dtype)
        check(lib.hb_spmm(self.engine.ctx, self.h, transa.encode(), transb.encode(), b_rows, b_cols, _host_ptr(a), B.ptr, ldb,
                          _host_ptr(b), Cm.ptr, ldc), "hb_spmm")

    def gemv(self, trans, alpha, x, beta, y, work=None):
        ny = self.rows if trans in "Nn" else self.cols
        if beta == 0 and y.size() != ny:
            y.resize(ny)                    # pntr_check_set_size: y is pure output when beta == 0
        a, b = _sc(alpha, self.dtype), _sc(beta, self.dtype)
        check(lib.hb_spmv(self.engine.ctx, self.h, trans.encode(), _host_ptr(a), _dptr(x), _host_ptr(b), _dptr(y)), "hb_spmv")


# ---------------------------------------------------------------- triangular solves / ILU(0) (gpu/hala_cuda_sparse_triangular.hpp, gpu/hala_gpu_ilu.hpp)
class gpu_triangular_matrix:
    """Non-owning view of a CSR of which only the `uplo` triangle is used (reference :38-110); analysis cached per direction."""

    def __init__(self, engine, uplo, diag, pntr, indx, vals, policy="N"):
        self.engine, self.uplo, self.diag = engine, uplo, diag
        self.dtype = vals.dtype
        self.nrows, self.nz = pntr.size() - 1, indx.size()
        self._keep = (pntr, indx, vals)
        self.h = C.c_void_p()
        check(lib.hb_tri_create(engine.ctx, _CODE[vals.dtype], uplo.encode(), diag.encode(), self.nrows, self.nz, pntr.ptr, indx.ptr, vals.ptr,
                                C.byref(self.h)), "hb_tri_create")

    def __del__(self):
        try:
            if self.h:
                lib.hb_tri_destroy(self.h)
        except Exception:
            pass

    def rows(self):
        return self.nrows

    def nnz(self):
        return self.nz

    def levels(self):
        m = C.c_int(0)
        check(lib.hb_tri_info(self.h, None, None, C.byref(m)), "hb_tri_info")
        return m.value

    def trsv_buffer_size(self, *a):
        return 0

    def trsv(self, trans, alpha, b, x, work=None):
        if x.size() != self.nrows:
            x.resize(self.nrows)                # pure output is resized (check_set_size, reference :418)
        a = _sc(alpha, self.dtype)
        check(lib.hb_sptrsv(self.engine.ctx, self.h, trans.encode(), _host_ptr(a), b.ptr, 1, x.ptr, 1), "hb_sptrsv")

    def trsm(self, transa, transb, nrhs, alpha, B, ldb=-1, work=None):
        if ldb < 0:
            ldb = self.nrows if transb in "Nn" else nrhs
        a = _sc(alpha, self.dtype)
        check(lib.hb_sptrsm(self.engine.ctx, self.h, transa.encode(), transb.encode(), nrhs, _host_ptr(a), B.ptr, ldb), "hb_sptrsm")


def make_triangular_matrix(engine, uplo, diag, pntr, indx, vals, policy="N"):
    return gpu_triangular_matrix(engine, uplo, diag, pntr, indx, vals, policy)


def sparse_trsv(trans, tri, alpha, b, x):
    tri.trsv(trans, alpha, b, x)


def sparse_trsm(transa, transb, nrhs, tri, alpha, B, ldb=-1):
    tri.trsm(transa, transb, nrhs, alpha, B, ldb)


class gpu_ilu:
    """ILU(0) preconditioner (reference gpu/hala_gpu_ilu.hpp:45-199): factors in the pattern of the matrix, applied as a unit-lower
    solve followed by an upper solve on the same array."""

    def __init__(self, engine, pntr, indx, vals, policy="N"):
        self.engine, self.dtype = engine, vals.dtype
        self.num_rows, self.nnz = pntr.size() - 1, indx.size()
        self.ilu = gpu_vector(engine, vals.dtype, self.nnz)
        check(lib.hb_ilu0(engine.ctx, _CODE[vals.dtype], self.num_rows, self.nnz, pntr.ptr, indx.ptr, vals.ptr, self.ilu.ptr), "hb_ilu0")
        self.upper = gpu_triangular_matrix(engine, "U", "N", pntr, indx, self.ilu, policy)
        self.lower = gpu_triangular_matrix(engine, "L", "U", pntr, indx, self.ilu, policy)
        self._tmp = gpu_vector(engine, vals.dtype, self.num_rows)

    def factors(self):
        return self.ilu.unload()

    def buffer_size(self, *a):
        return 0

    def apply(self, x, r, num_rhs=1, work=None):
        if r.size() != num_rhs * self.num_rows:
            r.resize(num_rhs * self.num_rows)
        if num_rhs == 1:
            self.lower.trsv("N", 1.0, x, self._tmp)
            self.upper.trsv("N", 1.0, self._tmp, r)
        else:
            vcopy(self.engine, x, r)
            self.lower.trsm("N", "N", num_rhs, 1.0, r, self.num_rows)
            self.upper.trsm("N", "N", num_rhs, 1.0, r, self.num_rows)


def make_ilu(engine, pntr, indx, vals, policy="N"):
    return gpu_ilu(engine, pntr, indx, vals, policy)


def make_sparse_matrix(engine, *args):
    """make_sparse_matrix(engine, [rows,] cols, [nnz,] pntr, indx, vals) — gpu/hala_cuda_sparse_general.hpp:382-401."""
    if len(args) == 6:
        rows, cols, nnz, pntr, indx, vals = args
    else:
        cols, pntr, indx, vals = args
        rows, nnz = pntr.size() - 1, indx.size()
    return gpu_sparse_matrix(engine, rows, cols, nnz, pntr, indx, vals)


def sparse_gemv(engine, trans, M, N, alpha, pntr, indx, vals, x, beta, y):
    """One-shot SpMV (gpu/hala_cuda_sparse_general.hpp:407-419): builds a temporary matrix view per call."""
    A = make_sparse_matrix(engine, M, N, indx.size(), pntr, indx, vals)
    A.set_transpose_mode("scatter")     # a view that lives for one product: building the transposed copy cannot pay off
    A.gemv(trans, alpha, x, beta, y)


# ---------------------------------------------------------------- solvers (identity preconditioner)
def solve_cg(engine, tol, max_iter, pntr, indx, vals, b, x, matrix=None):
    """hala::solve_cg(gpu_engine, stop_criteria(tol, max_iter), pntr, indx, vals, identity, b, x).
    x shorter than the system is resized and zeroed (hex/solvers/hala_solvers_cg.hpp:193-196).
    Returns (operator applications, final ||r||_2)."""
    n = pntr.size() - 1
    if x.size() < n:
        x.resize(n)
        x.fill(0)
    A = matrix if matrix is not None else make_sparse_matrix(engine, n, pntr, indx, vals)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_cg(engine.ctx, A.h, b.ptr, x.ptr, float(tol), int(max_iter), C.byref(it), C.byref(res)), "hb_cg")
    return it.value, res.value


def solve_gmres(engine, tol, max_outer, restart, pntr, indx, vals, b, x, cproj=None, matrix=None):
    """hala::solve_gmres(gpu_engine, stop, restart, pntr, indx, vals, identity, b, x).
    cproj: conjugated Gram-Schmidt projection; default = True for complex data (the reference's 'T' is a defect there)."""
    n = pntr.size() - 1
    if x.size() < n:
        x.resize(n)
        x.fill(0)
    if cproj is None:
        cproj = np.issubdtype(vals.dtype, np.complexfloating)
    A = matrix if matrix is not None else make_sparse_matrix(engine, n, pntr, indx, vals)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_gmres(engine.ctx, A.h, b.ptr, x.ptr, float(tol), int(max_outer), int(restart), 1 if cproj else 0,
                       C.byref(it), C.byref(res)), "hb_gmres")
    return it.value, res.value
