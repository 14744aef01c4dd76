"""TEST INFRASTRUCTURE — ctypes access to the CPU oracle (liboracle.so, prefix orc_) and, when it was built
from /root/reference, the unmodified reference cpu_engine path (oracle/_ref/libhala_ref*.so, prefix ref_/refc_).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
REAL = {0: np.float32, 1: np.float64, 2: np.float32, 3: np.float64}
OPS = {"copy": 0, "axpy": 1, "scal": 2, "dot": 3, "dotu": 4, "nrm2": 5}


def build(with_ref=True):
    """Compile liboracle.so and (only where /root/reference exists) oracle/_ref/*.so. Building the checker is not using it."""
    subprocess.run(["make", "-C", HERE, "liboracle.so"] + (["ref"] if with_ref else []), check=True,
                   stdout=subprocess.DEVNULL)


def C_char(c):
    return C.c_char(c.encode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _scalar(v, dt):
    return np.array([v], dtype=dt)


class _Lib:
    def __init__(self, path, prefix, has_cproj):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.has_cproj = has_cproj
        self.version = C.cast(getattr(self.lib, prefix + "version"), C.CFUNCTYPE(C.c_char_p))().decode()

    def _f(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def spmv(self, pntr, indx, vals, x, alpha=1.0, beta=0.0, y=None, trans="N", ncols=None):
        dt = vals.dtype
        M = pntr.size - 1
        N = ncols if ncols is not None else M
        out = np.zeros(M if trans == "N" else N, dtype=dt) if y is None else np.array(y, dtype=dt, copy=True)
        a, b = _scalar(alpha, dt), _scalar(beta, dt)
        rc = self._f("spmv")(CODE[dt], C.c_char(trans.encode()), M, N, _ptr(a), int(indx.size), _ptr(pntr), _ptr(indx),
                             _ptr(vals), _ptr(np.ascontiguousarray(x, dtype=dt)), _ptr(b), _ptr(out))
        assert rc == 0, rc
        return out

    def cg(self, pntr, indx, vals, b, tol, max_iter=10**6, x0=None):
        dt = vals.dtype
        n = pntr.size - 1
        x = np.zeros(n, dtype=dt) if x0 is None else np.array(x0, dtype=dt, copy=True)
        it = C.c_int(0)
        rc = self._f("cg")(CODE[dt], n, int(indx.size), _ptr(pntr), _ptr(indx), _ptr(vals),
                           _ptr(np.ascontiguousarray(b, dtype=dt)), _ptr(x), C.c_double(tol), int(max_iter), C.byref(it))
        assert rc == 0, rc
        return x, it.value

    def gmres(self, pntr, indx, vals, b, tol, restart, max_outer=10**6, x0=None, cproj=0):
        dt = vals.dtype
        n = pntr.size - 1
        x = np.zeros(n, dtype=dt) if x0 is None else np.array(x0, dtype=dt, copy=True)
        it = C.c_int(0)
        args = [CODE[dt], n, int(indx.size), _ptr(pntr), _ptr(indx), _ptr(vals),
                _ptr(np.ascontiguousarray(b, dtype=dt)), _ptr(x), C.c_double(tol), int(max_outer), int(restart)]
        if self.has_cproj:
            args.append(int(cproj))
        rc = self._f("gmres")(*args, C.byref(it))
        assert rc == 0, rc
        return x, it.value

    def gen_stencil7(self, n, lower=-1.0, diag=6.0, upper=-1.0, row_lo=0, row_hi=None):
        """rows [row_lo, row_hi) of the n^3 7-point stencil, generated in C (oracle only): (pntr, indx, vals)"""
        f = getattr(self.lib, self.prefix + "gen_stencil7")
        f.restype = C.c_longlong
        f.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        row_hi = n ** 3 if row_hi is None else row_hi
        nnz = f(n, row_lo, row_hi, lower, diag, upper, None, None, None)
        pntr = np.empty(row_hi - row_lo + 1, dtype=np.int32)
        indx, vals = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.float64)
        f(n, row_lo, row_hi, lower, diag, upper, _ptr(pntr), _ptr(indx), _ptr(vals))
        return pntr, indx, vals

    def blas1(self, op, x, y=None, alpha=1.0, n=None, incx=1, incy=1):
        dt = x.dtype
        code = CODE[dt]
        n = (1 + (x.size - 1) // incx if x.size else 0) if n is None else n
        a = _scalar(alpha, dt)
        yy = None if y is None else np.array(y, dtype=dt, copy=True)
        res = np.zeros(1, dtype=REAL[code] if op == "nrm2" else dt)
        # scal works in place on the y slot
        xx = np.ascontiguousarray(x)
        if op == "scal":
            yy = np.array(x, dtype=dt, copy=True)
            incy = incx
        rc = self._f("blas1")(code, OPS[op], int(n), _ptr(a), _ptr(xx), int(incx),
                              _ptr(yy) if yy is not None else None, int(incy), _ptr(res))
        assert rc == 0, rc
        return res[0] if op in ("dot", "dotu", "nrm2") else yy

    def spmm(self, transa, transb, M, N, K, pntr, indx, vals, B, ldb, alpha=1.0, beta=0.0, C=None, ldc=None):
        """C (M x N, column-major, ldc) = alpha op(A) op(B) + beta C  (sparse/hala_sparse_utils.hpp:120-160); B, C flat column-major"""
        dt = vals.dtype
        ldc = M if ldc is None else ldc
        out = np.zeros(ldc * N, dtype=dt) if C is None else np.array(C, dtype=dt, copy=True)
        a, b = _scalar(alpha, dt), _scalar(beta, dt)
        ch = lambda c: C_char(c)
        rc = self._f("spmm")(CODE[dt], C_char(transa), C_char(transb), M, N, K, _ptr(a), int(indx.size), _ptr(pntr), _ptr(indx), _ptr(vals),
                             _ptr(np.ascontiguousarray(B, dtype=dt)), int(ldb), _ptr(b), _ptr(out), int(ldc))
        assert rc == 0, rc
        return out

    def batch_cg(self, pntr, indx, vals, B, nrhs, tol, max_iter=10**6):
        """reference only: hala::solve_batch_cg with the identity preconditioner; B flat column-major rows x nrhs. Returns (X, iterations)."""
        dt = vals.dtype
        n = pntr.size - 1
        X = np.zeros(n * nrhs, dtype=dt)
        it = C.c_int(0)
        rc = self._f("batch_cg")(CODE[dt], n, int(indx.size), int(nrhs), _ptr(pntr), _ptr(indx), _ptr(vals), _ptr(np.ascontiguousarray(B, dtype=dt)),
                                 _ptr(X), C.c_double(tol), int(max_iter), C.byref(it))
        assert rc == 0, rc
        return X, it.value

    def trsv(self, uplo, diag, trans, pntr, indx, vals, b, alpha=1.0, general=False):
        """x = alpha * op(T)^-1 b.  general=False: CSR holds one triangle, diagonal last (L) / first (U) — the reference's layout
        (sparse/hala_sparse_utils.hpp:283-335); general=True (oracle only): any CSR, only the `uplo` part is used."""
        dt = vals.dtype
        n = pntr.size - 1
        x = np.zeros(n, dtype=dt)
        a = _scalar(alpha, dt)
        bb = np.ascontiguousarray(b, dtype=dt)
        ch = lambda c: C.c_char(c.encode())
        if self.prefix == "orc_":
            rc = self._f("trsv")(CODE[dt], int(bool(general)), ch(uplo), ch(diag), ch(trans), n, _ptr(a), _ptr(pntr), _ptr(indx), _ptr(vals), _ptr(bb), _ptr(x))
        else:
            assert not general, "the reference's cpu_triangular_matrix expects a one-triangle CSR"
            rc = self._f("trsv")(CODE[dt], ch(uplo), ch(diag), ch(trans), n, _ptr(a), int(indx.size), _ptr(pntr), _ptr(indx), _ptr(vals), _ptr(bb), _ptr(x))
        assert rc == 0, rc
        return x

    def ilu(self, pntr, indx, vals, x):
        """ILU(0) factors in the pattern of the matrix (sorted rows, diagonal present) and U^-1 L^-1 x
        (sparse/hala_sparse_utils.hpp:228-274). Returns (factors, applied)."""
        dt = vals.dtype
        n = pntr.size - 1
        fac = np.zeros(indx.size, dtype=dt)
        xx = np.ascontiguousarray(x, dtype=dt)
        if self.prefix == "orc_":
            diag = np.zeros(n, dtype=np.int32)
            rc = self._f("ilu_factor")(CODE[dt], n, _ptr(pntr), _ptr(indx), _ptr(vals), _ptr(diag), _ptr(fac))
            assert rc == 0, rc
            r = xx.copy()
            rc = self._f("ilu_apply")(CODE[dt], n, _ptr(pntr), _ptr(indx), _ptr(diag), _ptr(fac), _ptr(r))
        else:
            r = np.zeros(n, dtype=dt)
            rc = self._f("ilu")(CODE[dt], n, int(indx.size), _ptr(pntr), _ptr(indx), _ptr(vals), _ptr(fac), _ptr(xx), _ptr(r))
        assert rc == 0, rc
        return fac, r

    def gemv(self, trans, M, N, A, x, alpha=1.0, beta=0.0, y=None, lda=None):
        dt = A.dtype
        lda = M if lda is None else lda
        out = np.zeros(M if trans == "N" else N, dtype=dt) if y is None else np.array(y, dtype=dt, copy=True)
        a, b = _scalar(alpha, dt), _scalar(beta, dt)
        rc = self._f("gemv")(CODE[dt], C.c_char(trans.encode()), M, N, _ptr(a), _ptr(np.ascontiguousarray(A)), int(lda),
                             _ptr(np.ascontiguousarray(x, dtype=dt)), _ptr(b), _ptr(out))
        assert rc == 0, rc
        return out


_cache = {}


def _preload_blas_deps():
    """The wheel's OpenBLAS needs the libgfortran/libquadmath that sit beside it but carries no RUNPATH for them."""
    import glob
    import sysconfig
    d = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
    for pat in ("libquadmath-*.so*", "libgfortran-*.so*"):
        for f in sorted(glob.glob(os.path.join(d, pat))):
            try:
                C.CDLL(f, mode=C.RTLD_GLOBAL)
            except OSError:
                pass


def oracle():
    """The plain-C restatement (always available once built)."""
    if "orc" not in _cache:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        _cache["orc"] = _Lib(path, "orc_", True)
    return _cache["orc"]


def reference(cpatch=False):
    """The unmodified reference cpu_engine path, or None when oracle/_ref was never built (no /root/reference)."""
    key = "refc" if cpatch else "ref"
    if key not in _cache:
        path = os.path.join(HERE, "_ref", "libhala_ref_cpatch.so" if cpatch else "libhala_ref.so")
        try:
            _preload_blas_deps()
            _cache[key] = _Lib(path, "refc_" if cpatch else "ref_", False) if os.path.exists(path) else None
        except OSError:
            _cache[key] = None
    return _cache[key]


class _RefGpu:
    """The unmodified reference's own gpu_engine path (cuSPARSE + cuBLAS, oracle/_ref/libhala_ref_gpu.so): bench.py's
    gpu_reference leg.  All arrays are device pointers (ints)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.refgpu_version.restype = C.c_char_p
        self.lib.refgpu_last_error.restype = C.c_char_p
        self.version = self.lib.refgpu_version().decode()

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.refgpu_last_error().decode()}")

    def spmv_us(self, code, rows, cols, nnz, pntr, indx, vals, x, y, warmup=5, reps=50):
        us = C.c_double(0)
        vp = C.c_void_p
        self._check(self.lib.refgpu_spmv(code, rows, cols, nnz, vp(pntr), vp(indx), vp(vals), vp(x), vp(y), warmup, reps, C.byref(us)), "refgpu_spmv")
        return us.value

    def cg(self, code, rows, nnz, pntr, indx, vals, b, x, tol, max_iter):
        it, sec = C.c_int(0), C.c_double(0)
        vp = C.c_void_p
        self._check(self.lib.refgpu_cg(code, rows, nnz, vp(pntr), vp(indx), vp(vals), vp(b), vp(x), C.c_double(tol), int(max_iter),
                                       C.byref(it), C.byref(sec)), "refgpu_cg")
        return it.value, sec.value

    def gmres(self, code, rows, nnz, pntr, indx, vals, b, x, tol, max_outer, restart):
        it, sec = C.c_int(0), C.c_double(0)
        vp = C.c_void_p
        self._check(self.lib.refgpu_gmres(code, rows, nnz, vp(pntr), vp(indx), vp(vals), vp(b), vp(x), C.c_double(tol), int(max_outer), int(restart),
                                          C.byref(it), C.byref(sec)), "refgpu_gmres")
        return it.value, sec.value


def reference_gpu():
    """The reference's cuSPARSE/cuBLAS path, or None when it was never built or its CUDA libraries do not load here."""
    if "refgpu" not in _cache:
        path = os.path.join(HERE, "_ref", "libhala_ref_gpu.so")
        try:
            _preload_blas_deps()
            _cache["refgpu"] = _RefGpu(path) if os.path.exists(path) else None
        except OSError:
            _cache["refgpu"] = None
    return _cache["refgpu"]
