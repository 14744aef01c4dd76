"""Shared checks. Tolerances are the ones BASELINE.json's north_star states: SpMV 1e-13 (fp64) / 1e-5 (fp32) relative
per entry, measured against sum_j |a_ij||x_j|; iteration counts +-2."""
import numpy as np

DT = ["f32", "f64", "c32", "c64"]
NP = {"f32": np.float32, "f64": np.float64, "c32": np.complex64, "c64": np.complex128}
SPMV_TOL = {"f32": 1e-5, "f64": 1e-13, "c32": 1e-5, "c64": 1e-13}
# reductions over n terms in a different order than the reference's BLAS: n * eps style bound on |x|.|y|
RED_TOL = {"f32": 2e-5, "f64": 1e-13, "c32": 2e-5, "c64": 1e-13}


def dense_from_csr(pntr, indx, vals, ncols):
    M = pntr.size - 1
    A = np.zeros((M, ncols), dtype=vals.dtype)
    for i in range(M):
        for j in range(pntr[i], pntr[i + 1]):
            A[i, indx[j]] = vals[j]
    return A


def spmv_scale(pntr, indx, vals, x, trans="N", ncols=None, alpha=1.0, beta=0.0, y0=None):
    """Per-entry magnitude sum_j |alpha||a_ij||x_j| + |beta||y0_i| used to scale the tolerance."""
    M = pntr.size - 1
    N = M if ncols is None else ncols
    rows = np.repeat(np.arange(M), np.diff(pntr))
    av, ax = np.abs(vals).astype(np.float64), np.abs(x).astype(np.float64)
    if trans == "N":
        s = np.bincount(rows, weights=av * ax[indx], minlength=M)
    else:
        s = np.bincount(indx, weights=av * ax[rows], minlength=N)
    s = abs(alpha) * s
    if y0 is not None:
        s = s + abs(beta) * np.abs(y0)
    return s


def assert_entrywise(y, yref, scale, tol, what=""):
    y, yref = np.asarray(y), np.asarray(yref)
    assert y.shape == yref.shape, (what, y.shape, yref.shape)
    assert np.all(np.isfinite(y)), what
    err = np.abs(y.astype(np.complex128) - yref.astype(np.complex128))
    bound = tol * np.maximum(scale, np.finfo(np.float64).tiny)
    worst = np.max(err / np.maximum(scale, 1e-300)) if err.size else 0.0
    assert np.all(err <= bound), f"{what}: worst scaled error {worst:.3e} > {tol:.1e}"


def assert_reduction(val, ref, x, y, dt, what="", tol=None):
    scale = float(np.sum(np.abs(x).astype(np.float64) * np.abs(y).astype(np.float64)))
    tol = RED_TOL[dt] if tol is None else tol
    assert abs(complex(val) - complex(ref)) <= tol * max(scale, 1e-300), (what, val, ref)
