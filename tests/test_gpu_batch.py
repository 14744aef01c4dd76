"""GPU parity (-m gpu) of SURVEY.md §8 row f2: SpMM (hb_spmm), the dense helpers of the batch solvers (hb_geam, hb_dgmm, hb_tbsv) and a
batch CG assembled from them the way wax/hala_blas_extensions.hpp + hex/solvers/hala_solvers_cg_batch.hpp do, against outputs of
the unmodified reference (tests/golden/ref_outputs_f2.npz), the oracle and numpy definitions."""
import numpy as np
import pytest

import hala_b200 as hb
from hala_b200 import matgen as mg
from helpers import DT, NP

pytestmark = pytest.mark.gpu
TOL = {"f32": 3e-4, "f64": 1e-11, "c32": 3e-4, "c64": 1e-11}


@pytest.mark.parametrize("dt", DT)
def test_spmm_matches_recorded_reference(engine, golden_f2, dt):
    p, i, v = mg.perturbed("convdiff7", 7, dt)
    n, N = p.size - 1, 5
    gp, gi, gv = engine.load(p), engine.load(i), engine.load(v)
    A = hb.make_sparse_matrix(engine, n, gp, gi, gv)
    for ta in "NTC":
        for tb in "NTC":
            ldb = n + 3 if tb == "N" else N + 2
            B = mg.probe_x(ldb * (N if tb == "N" else n), dt, seed=5)
            C0 = mg.probe_x((n + 1) * N, dt, seed=6)
            gC = engine.load(C0)
            A.gemm(ta, tb, n if tb == "N" else N, N if tb == "N" else n, 1.5, engine.load(B), ldb, 0.5, gC, n + 1)
            want = golden_f2[f"spmm/convdiff7:7/{dt}/{ta}{tb}"]
            assert np.abs(gC.unload() - want).max() <= TOL[dt] * np.abs(want).max(), (ta, tb)


@pytest.mark.parametrize("dt", ["f64", "c64"])
@pytest.mark.parametrize("name,n,N", [("lap3d27", 16, 7), ("lap3d7", 30, 4), ("lap2d", 100, 9)])
def test_spmm_vs_oracle_larger(engine, orc, dt, name, n, N):
    """every lanes-per-row instantiation (27-point: 8, 7-point / 5-point: 1), column counts that are not multiples of 4, beta == 0 output not read"""
    p, i, v = mg.perturbed(name, n, dt)
    rows = p.size - 1
    A = hb.make_sparse_matrix(engine, rows, engine.load(p), engine.load(i), engine.load(v))
    B = mg.probe_x(rows * N, dt, seed=5)
    gC = engine.load(np.full(rows * N, np.nan, dtype=NP[dt]))
    A.gemm("N", "N", rows, N, 2.0, engine.load(B), rows, 0.0, gC, rows)
    want = orc.spmm("N", "N", rows, N, rows, p, i, v, B, rows, alpha=2.0)
    got = gC.unload()
    assert np.all(np.isfinite(got)) and np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    Bt = np.ascontiguousarray(B.reshape(N, rows).T).reshape(-1)          # the same B stored transposed: N x rows, ldb = N
    A.gemm("N", "T", N, rows, 2.0, engine.load(Bt), N, 0.0, gC, rows)
    assert np.abs(gC.unload() - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("dt", DT)
def test_geam_dgmm_tbsv(engine, dt):
    t = NP[dt]
    rt = 1e-5 if "32" in dt else 1e-13
    M, N = 37, 45
    for ta in "NTC":
        for tb in "NTC":
            lda, ldb, ldc = (M if ta == "N" else N) + 2, (M if tb == "N" else N) + 1, M + 3
            A = mg.probe_x(lda * (N if ta == "N" else M), dt, seed=1).reshape((N if ta == "N" else M), lda)
            B = mg.probe_x(ldb * (N if tb == "N" else M), dt, seed=2).reshape((N if tb == "N" else M), ldb)
            op = lambda X, tr, ld: {"N": X[:, :M].T, "T": X[:, :N], "C": X[:, :N].conj()}[tr]     # -> M x N
            a, b = (2.0, -0.5) if not dt.startswith("c") else (2.0 - 1.0j, -0.5 + 0.25j)
            want = a * op(A, ta, lda) + b * op(B, tb, ldb)
            gC = engine.load(np.zeros(ldc * N, dtype=t))
            hb.geam(engine, ta, tb, M, N, a, engine.load(A.reshape(-1)), lda, b, engine.load(B.reshape(-1)), ldb, gC, ldc)
            got = gC.unload().reshape(N, ldc)[:, :M].T
            np.testing.assert_allclose(got, want.astype(t), rtol=10 * rt, atol=10 * rt)
    # transpose through geam, as wax/hala_blas_extensions.hpp:190-194 does: geam('T','T', N, M, 1, A, lda, 0, A, lda, At, ldat)
    A = mg.probe_x((M + 1) * N, dt, seed=3).reshape(N, M + 1)
    gAt = engine.load(np.zeros(N * M, dtype=t))
    hb.geam(engine, "T", "T", N, M, 1.0, engine.load(A.reshape(-1)), M + 1, 0.0, engine.load(A.reshape(-1)), M + 1, gAt, N)
    np.testing.assert_array_equal(gAt.unload().reshape(M, N), A[:, :M].T)
    # dgmm both sides
    A = mg.probe_x(M * N, dt, seed=4).reshape(N, M)
    xl, xr = mg.probe_x(M, dt, seed=5), mg.probe_x(2 * N, dt, seed=6)
    gC = engine.new_vector(t)
    hb.dgmm(engine, "L", M, N, engine.load(A.reshape(-1)), M, engine.load(xl), 1, gC, M)
    np.testing.assert_allclose(gC.unload().reshape(N, M), A * xl[None, :], rtol=10 * rt)
    hb.dgmm(engine, "R", M, N, engine.load(A.reshape(-1)), M, engine.load(xr), 2, gC, M)
    np.testing.assert_allclose(gC.unload().reshape(N, M), A * xr[::2][:, None], rtol=10 * rt)
    # tbsv: bandwidth 0 (vdivide) and a general band against a dense solve
    d = mg.probe_x(N, dt, seed=7) + t(3)
    x = mg.probe_x(N, dt, seed=8)
    gx = engine.load(x)
    hb.tbsv(engine, "U", "N", "N", N, 0, engine.load(d), 1, gx)
    np.testing.assert_allclose(gx.unload(), x / d, rtol=10 * rt)
    k, n = 3, 29
    for uplo in "UL":
        band = mg.probe_x((k + 1) * n, dt, seed=9).reshape(n, k + 1)
        band[:, k if uplo == "U" else 0] += t(4)
        dense = np.zeros((n, n), dtype=t)
        for j in range(n):
            for r in range(k + 1):
                i_ = j - k + r if uplo == "U" else j + r
                if 0 <= i_ < n:
                    dense[i_, j] = band[j, r]
        for tr, opm in (("N", dense), ("T", dense.T), ("C", dense.conj().T)):
            for diag in "NU":
                m = opm.copy()
                if diag == "U":
                    np.fill_diagonal(m, 1)
                rhs = mg.probe_x(n, dt, seed=10)
                gx = engine.load(rhs)
                hb.tbsv(engine, uplo, tr, diag, n, k, engine.load(band.reshape(-1)), k + 1, gx)
                np.testing.assert_allclose(gx.unload(), np.linalg.solve(m.astype(np.complex128), rhs.astype(np.complex128)).astype(t),
                                           rtol=2e-3 if "32" in dt else 1e-10, atol=2e-3 if "32" in dt else 1e-10)


@pytest.mark.parametrize("dt", DT)
def test_batch_cg_from_primitives_matches_recorded_reference(engine, golden_f2, dt):
    """solve_batch_cg restated on the device with exactly the primitives the wax templates use: SpMM, dgmm + axpy (batch_axpy),
    dgmm + gemv('T') with a vector of ones (batch_dot), tbsv with bandwidth 0 (vdivide), dgmm + vcopy (batch_scal)."""
    t = NP[dt]
    tol = 1e-4 if "32" in dt else 1e-9
    p, i, v = mg.GENERATORS["lap2d"](20, dtype=dt)
    n, nrhs = p.size - 1, 4
    Bh = mg.probe_x(n * nrhs, dt, seed=8)
    A = hb.make_sparse_matrix(engine, n, engine.load(p), engine.load(i), engine.load(v))
    e = engine
    ones = e.load(np.ones(n, dtype=t))
    nv = lambda: e.new_vector(t, n * nrhs)
    ns = lambda: e.new_vector(t, nrhs)

    def bdot(x, y, out):                    # conj(x) . y per column
        xc = x
        if dt.startswith("c"):
            xc = nv()
            hb.geam(e, "C", "C", n * nrhs, 1, 1.0, x, 1, 0.0, x, 1, xc, n * nrhs)
        prod = nv()
        hb.dgmm(e, "L", n * nrhs, 1, xc, n * nrhs, y, 1, prod, n * nrhs)
        hb.gemv(e, "T", n, nrhs, 1.0, prod, ones, 0.0, out)

    def baxpy(sign, a, x, y):               # y += sign * x diag(a)
        sx = nv()
        hb.dgmm(e, "R", n, nrhs, x, n, a, 1, sx, n)
        hb.axpy(e, sign, sx, y)

    X, R, P, AP = e.load(np.zeros(n * nrhs, dtype=t)), e.load(Bh), nv(), nv()
    hb.vcopy(e, R, P)
    zr, pap, alpha, zr2, beta = ns(), ns(), ns(), ns(), ns()
    bdot(R, R, zr)
    its = 1
    while True:
        A.gemm("N", "N", n, nrhs, 1.0, P, n, 0.0, AP, n)
        its += 1
        bdot(P, AP, pap)
        hb.vcopy(e, zr, alpha); hb.tbsv(e, "U", "N", "N", nrhs, 0, pap, 1, alpha)
        baxpy(1.0, alpha, P, X)
        baxpy(-1.0, alpha, AP, R)
        bdot(R, R, zr2)
        if np.sqrt(np.abs(zr2.unload()).max()) < tol or its > 2000:
            break
        hb.vcopy(e, zr2, beta); hb.tbsv(e, "U", "N", "N", nrhs, 0, zr, 1, beta)
        tmp = nv()
        hb.dgmm(e, "R", n, nrhs, P, n, beta, 1, tmp, n)
        hb.vcopy(e, tmp, P)
        hb.axpy(e, 1.0, R, P)
        zr, zr2 = zr2, zr
    want = golden_f2[f"batch_cg/lap2d:20/{dt}/x"]
    assert abs(its - int(golden_f2[f"batch_cg/lap2d:20/{dt}/iters"][0])) <= 2, its
    assert np.abs(X.unload() - want).max() <= 50 * tol * max(1.0, np.abs(want).max())
