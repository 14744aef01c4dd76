"""CPU tests of the oracle itself (no GPU): the C restatement (oracle/hb_oracle.c) against
  (1) vectors the reference's own tests hold (tests/golden/ref_tests.json, file:line inside),
  (2) outputs of the unmodified reference recorded in tests/golden/ref_outputs.npz,
  (3) known-answer iteration counts of the reference,
  (4) the reference itself, live, when oracle/_ref was built in this container."""
import numpy as np
import pytest

from hala_b200 import matgen as mg
from helpers import DT, NP, SPMV_TOL, RED_TOL, assert_entrywise, assert_reduction, dense_from_csr, spmv_scale


def _fixture(ref_tests, name, dt):
    f = ref_tests[name]
    return (np.array(f["pntr"], dtype=np.int32), np.array(f["indx"], dtype=np.int32), np.array(f["vals"], dtype=NP[dt]), f)


@pytest.mark.parametrize("dt", DT)
def test_reference_test_sparse_gemv(orc, ref_tests, dt):
    """tests/sparse_tests.hpp:166-191: 5x5 tridiagonal, alpha = 2, beta = 0, ops N/T/C against the dense product."""
    p, i, v, f = _fixture(ref_tests, "tridiag5", dt)
    x = (np.array(f["x_real"], dtype=NP[dt]) if dt in ("f32", "f64")
         else np.array([complex(a, b) for a, b in f["x_complex"]], dtype=NP[dt]))
    A = dense_from_csr(p, i, v, 5)
    for tr, op in (("N", A), ("T", A.T), ("C", A.conj().T)):
        y = orc.spmv(p, i, v, x, alpha=2.0, beta=0.0, trans=tr)
        np.testing.assert_allclose(y, 2.0 * (op @ x), rtol=1e-5 if "32" in dt else 1e-14)


def test_reference_post_install_smoke(orc, ref_tests):
    """cmake/post_install_test.sh:24-38: exact {4, 10, 18}."""
    f = ref_tests["post_install"]
    y = orc.spmv(np.array(f["pntr"], dtype=np.int32), np.array(f["indx"], dtype=np.int32), np.array(f["vals"]), np.array(f["x"]))
    assert y.tolist() == f["y"]


@pytest.mark.parametrize("dt", DT)
def test_reference_rect_matrix(orc, ref_tests, dt):
    """tests/sparse_tests.hpp:140-152 fixture (5x6), as gemv N and T/C."""
    p, i, v, f = _fixture(ref_tests, "rect5x6", dt)
    A = dense_from_csr(p, i, v, 6)
    x6, x5 = mg.probe_x(6, dt), mg.probe_x(5, dt)
    rt = 1e-5 if "32" in dt else 1e-14
    np.testing.assert_allclose(orc.spmv(p, i, v, x6, trans="N", ncols=6), A @ x6, rtol=rt, atol=rt)
    np.testing.assert_allclose(orc.spmv(p, i, v, x5, trans="T", ncols=6), A.T @ x5, rtol=rt, atol=rt)
    np.testing.assert_allclose(orc.spmv(p, i, v, x5, trans="C", ncols=6), A.conj().T @ x5, rtol=rt, atol=rt)


@pytest.mark.parametrize("dt", DT)
def test_reference_solver_fixtures(orc, ref_tests, dt):
    """tests/solvers_tests.hpp matrices and xref = {1..5} (the reference solves them with ILU; here identity
    preconditioner, same tolerance, same expected solution)."""
    tol = 1e-4 if "32" in dt else 1e-9
    xtol = 2e-3 if "32" in dt else 1e-7
    for name in ("tridiag5", "cyclic5_spd"):
        p, i, v, f = _fixture(ref_tests, name, dt)
        xref = np.array(f["xref"], dtype=NP[dt])
        b = dense_from_csr(p, i, v, 5) @ xref
        x, it = orc.cg(p, i, v, b, tol, x0=f.get("x0"))
        assert it <= 8
        np.testing.assert_allclose(x, xref, atol=xtol)
    p, i, v, f = _fixture(ref_tests, "nonsym5", dt)
    xref = np.array(f["xref"], dtype=NP[dt])
    b = dense_from_csr(p, i, v, 5) @ xref
    x, it = orc.gmres(p, i, v, b, tol, f["restart"])
    np.testing.assert_allclose(x, xref, atol=xtol)


def test_known_answer_iteration_counts(orc, ref_tests):
    """Iteration counts of the unmodified reference — exact for the serial restatement."""
    ka = ref_tests["known_answers"]
    for key in ("lap2d:256", "lap3d7:64", "lap3d27:64"):
        name, n = key.split(":")
        p, i, v = mg.GENERATORS[name](int(n))
        _, it = orc.cg(p, i, v, mg.rhs(p.size - 1), 1e-8)
        assert abs(it - ka["cg_iterations"][key]) <= 2, (key, it)
    for key in ("convdiff7:24", "convdiff7:48"):
        name, n = key.split(":")
        p, i, v = mg.GENERATORS[name](int(n))
        _, it = orc.gmres(p, i, v, mg.rhs(p.size - 1), 1e-8, 50)
        assert abs(it - ka["gmres50_iterations"][key]) <= 2, (key, it)
    p, i, v = mg.lap2d(1024)
    b = mg.rhs(1024 * 1024)
    y = orc.spmv(p, i, v, b)
    assert y[0] == ka["spmv_spot_lap2d_1024_x_eq_b"]["y0"] and y[512 * 1024] == ka["spmv_spot_lap2d_1024_x_eq_b"]["y_half"]


def test_oracle_matches_recorded_reference_spmv(orc, golden):
    keys = [k for k in golden.files if k.startswith("spmv/")]
    assert len(keys) >= 40
    for k in keys:
        _, mat, dt, tr = k.split("/")
        name, n = mat.split(":")
        if name == "powerlaw":
            p, i, v = mg.powerlaw(N=int(n), lmax=700 if dt == "f64" else 500, dtype=dt)
            x = mg.probe_x(int(n), dt)
            y = orc.spmv(p, i, v, x)
            scale = spmv_scale(p, i, v, x)
        else:
            p, i, v = mg.GENERATORS[name](int(n), dtype=dt)
            N = p.size - 1
            x, y0 = mg.probe_x(N, dt), mg.probe_x(N, dt, seed=99)
            y = orc.spmv(p, i, v, x, alpha=2.0, beta=0.5, y=y0, trans=tr)
            scale = spmv_scale(p, i, v, x, tr, alpha=2.0, beta=0.5, y0=y0)
        assert_entrywise(y, golden[k], scale, SPMV_TOL[dt], k)


@pytest.mark.parametrize("dt", DT)
def test_oracle_matches_recorded_reference_blas(orc, golden, dt):
    n = 1003
    x, y = mg.probe_x(n * 3, dt, seed=3), mg.probe_x(n * 3, dt, seed=4)
    a = 1.5 if dt in ("f32", "f64") else 1.5 - 0.5j
    rt = 1e-5 if "32" in dt else 1e-13
    for incx, incy in ((1, 1), (2, 3)):
        tag = f"{dt}/{incx}{incy}"
        np.testing.assert_allclose(orc.blas1("axpy", x, y, alpha=a, n=n, incx=incx, incy=incy), golden[f"axpy/{tag}"], rtol=rt, atol=rt)
        np.testing.assert_array_equal(orc.blas1("copy", x, y, n=n, incx=incx, incy=incy), golden[f"copy/{tag}"])
        np.testing.assert_allclose(orc.blas1("scal", x, alpha=a, n=n, incx=incx), golden[f"scal/{tag}"], rtol=rt, atol=rt)
        xs, ys = x[::incx][:n], y[::incy][:n]
        assert_reduction(orc.blas1("dot", x, y, n=n, incx=incx, incy=incy), golden[f"dot/{tag}"][0], xs, ys, dt, "dot")
        assert_reduction(orc.blas1("dotu", x, y, n=n, incx=incx, incy=incy), golden[f"dotu/{tag}"][0], xs, ys, dt, "dotu")
        np.testing.assert_allclose(orc.blas1("nrm2", x, n=n, incx=incx), golden[f"nrm2/{tag}"][0], rtol=RED_TOL[dt])
    M, K = 777, 7
    A = mg.probe_x(M * K, dt, seed=21)
    xm, xk = mg.probe_x(M, dt, seed=22), mg.probe_x(K, dt, seed=23)
    gt = 5e-4 if "32" in dt else 1e-12
    np.testing.assert_allclose(orc.gemv("T", M, K, A, xm), golden[f"gemv/{dt}/T"], rtol=gt, atol=gt)
    np.testing.assert_allclose(orc.gemv("C", M, K, A, xm), golden[f"gemv/{dt}/C"], rtol=gt, atol=gt)
    np.testing.assert_allclose(orc.gemv("N", M, K, A, xk, alpha=-1.0, beta=1.0, y=xm), golden[f"gemv/{dt}/N"], rtol=gt, atol=gt)


def test_oracle_matches_recorded_reference_solvers(orc, golden):
    for k in [k for k in golden.files if k.startswith("cg/") and k.endswith("/iters")]:
        _, mat, dt, _ = k.split("/")
        name, n = mat.split(":")
        p, i, v = mg.GENERATORS[name](int(n), dtype=dt)
        tol = 1e-4 if "32" in dt else 1e-8
        x, it = orc.cg(p, i, v, mg.rhs(p.size - 1, dt), tol)
        assert abs(it - int(golden[k][0])) <= 2, (k, it, golden[k][0])
        np.testing.assert_allclose(x, golden[k.replace("/iters", "/x")], atol=(50 * tol))
    for k in [k for k in golden.files if k.startswith("gmres/") and k.endswith("/iters")]:
        _, mat, dt, m, _ = k.split("/")
        name, n = mat.split(":")
        p, i, v = mg.GENERATORS[name](int(n), dtype=dt)
        tol = 1e-4 if "32" in dt else 1e-8
        x, it = orc.gmres(p, i, v, mg.rhs(p.size - 1, dt), tol, int(m[1:]), cproj=1 if dt.startswith("c") else 0)
        assert abs(it - int(golden[k][0])) <= 2, (k, it, golden[k][0])
        np.testing.assert_allclose(x, golden[k.replace("/iters", "/x")], atol=(50 * tol))


def test_oracle_against_live_reference(orc):
    """Only where oracle/_ref exists (built from /root/reference in the build container; shipped to the GPU box)."""
    from oracle import binding
    ref = binding.reference()
    if ref is None:
        pytest.skip("oracle/_ref not built here")
    for dt in DT:
        p, i, v = mg.lap3d27(7, dtype=dt)
        x = mg.probe_x(p.size - 1, dt)
        assert_entrywise(orc.spmv(p, i, v, x), ref.spmv(p, i, v, x), spmv_scale(p, i, v, x), SPMV_TOL[dt], dt)
    p, i, v = mg.lap2d(96)
    b = mg.rhs(96 * 96)
    assert abs(orc.cg(p, i, v, b, 1e-8)[1] - ref.cg(p, i, v, b, 1e-8)[1]) <= 2
    p, i, v = mg.convdiff7(14)
    b = mg.rhs(14 ** 3)
    assert abs(orc.gmres(p, i, v, b, 1e-8, 20)[1] - ref.gmres(p, i, v, b, 1e-8, 20)[1]) <= 2


def test_generators_are_well_formed():
    for name, n in (("lap2d", 17), ("lap3d7", 6), ("lap3d27", 5), ("convdiff7", 5), ("helmholtz7", 4)):
        p, i, v = mg.GENERATORS[name](n)
        N = mg.grid_rows(name, n)
        assert p.size == N + 1 and p[0] == 0 and p[-1] == i.size == v.size
        for r in range(N):
            cols = i[p[r]:p[r + 1]]
            assert np.all(np.diff(cols) > 0) and r in cols
    assert mg.lap2d(1024)[1].size == 5 * 1024 * 1024 - 4 * 1024
    assert mg.lap3d27(16)[1].size == (3 * 16 - 2) ** 3
    p, i, v = mg.powerlaw(N=5000, lmax=900)
    lens = np.diff(p)
    assert lens.min() >= 1 and lens.max() <= 900 and lens.max() > 100
    for r in (0, 17, 4999):
        cols = i[p[r]:p[r + 1]]
        assert np.all(np.diff(cols) > 0) and r in cols
    A = dense_from_csr(*mg.powerlaw(N=300, lmax=64), 300)
    assert np.all(np.abs(np.diag(A)) > np.sum(np.abs(A), axis=1) - np.abs(np.diag(A)))


# ---------------------------------------------------------------- triangular solves and ILU(0)
F1_TOL = {"f32": 2e-4, "f64": 1e-11, "c32": 2e-4, "c64": 1e-11}


@pytest.mark.parametrize("dt", DT)
def test_oracle_trsv_ilu_match_recorded_reference(orc, golden_f1, dt):
    for name, n in mg.F1_CASES:
        p, i, v = mg.perturbed(name, n, dt)
        b = mg.probe_x(p.size - 1, dt, seed=31)
        for uplo in "LU":
            tp, ti, tv = mg.split_triangle(p, i, v, uplo)
            for diag in "NU":
                for tr in "NTC":
                    want = golden_f1[f"trsv/{name}:{n}/{dt}/{uplo}{diag}{tr}"]
                    scale = np.abs(want).max()
                    got = orc.trsv(uplo, diag, tr, tp, ti, tv, b, alpha=0.5)
                    assert np.abs(got - want).max() <= F1_TOL[dt] * scale, (name, uplo, diag, tr)
                    # the general-CSR form (cuSPARSE fill-mode semantics, what gpu_ilu relies on) on the FULL matrix
                    got = orc.trsv(uplo, diag, tr, p, i, v, b, alpha=0.5, general=True)
                    assert np.abs(got - want).max() <= F1_TOL[dt] * scale, (name, uplo, diag, tr, "general")
        fac, r = orc.ilu(p, i, v, b)
        assert np.abs(fac - golden_f1[f"ilu/{name}:{n}/{dt}/factors"]).max() <= F1_TOL[dt] * np.abs(fac).max()
        want = golden_f1[f"ilu/{name}:{n}/{dt}/apply"]
        assert np.abs(r - want).max() <= F1_TOL[dt] * np.abs(want).max()


@pytest.mark.parametrize("dt", DT)
def test_oracle_trsv_ilu_reference_test_vectors(orc, dt):
    """tests/sparse_tests.hpp:259-311 (trsv: X = 0.5 * op(T)^-1 (2 op(A) Xref) == Xref) and :423-440 (ilu)"""
    t = NP[dt]
    A = np.array([[1, 0, 0], [4, 2, 0], [5, 0, 3]], dtype=t)
    xref = mg.probe_x(3, dt, seed=4)
    pl, il, vl = np.array([0, 1, 3, 5], np.int32), np.array([0, 0, 1, 0, 2], np.int32), np.array([1, 4, 2, 5, 3], dtype=t)
    for tr, op in (("N", A), ("T", A.T), ("C", A.conj().T)):
        np.testing.assert_allclose(orc.trsv("L", "N", tr, pl, il, vl, 2 * op @ xref, alpha=0.5), xref, rtol=1e-5 if "32" in dt else 1e-13)
    Au = np.array([[1, 0, 0], [1, 1, 0], [2, 0, 1]], dtype=t)
    vu = np.array([2, 1, 3, 2, 4], dtype=t)
    np.testing.assert_allclose(orc.trsv("L", "U", "N", pl, il, vu, 2 * Au @ xref, alpha=0.5), xref, rtol=1e-5 if "32" in dt else 1e-13)
    pu, iu, vv = np.array([0, 3, 4, 5], np.int32), np.array([0, 1, 2, 1, 2], np.int32), np.array([1, 4, 5, 2, 3], dtype=t)
    np.testing.assert_allclose(orc.trsv("U", "N", "N", pu, iu, vv, 2 * A.T @ xref, alpha=0.5), xref, rtol=1e-5 if "32" in dt else 1e-13)
    np.testing.assert_allclose(orc.trsv("U", "N", "T", pu, iu, vv, 2 * A @ xref, alpha=0.5), xref, rtol=1e-5 if "32" in dt else 1e-13)
    p, i = np.array([0, 3, 6, 9], np.int32), np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], np.int32)
    fac, r = orc.ilu(p, i, np.array([3, 2, 1, 2, 4, 3, 2, 1, 6], dtype=t), np.array([1, 2, 3], dtype=t))
    np.testing.assert_allclose(fac, [3, 2, 1, 2 / 3, 8 / 3, 7 / 3, 2 / 3, -0.125, 5.625], rtol=1e-5 if "32" in dt else 1e-13)
    np.testing.assert_allclose(r, [1 / 9, 1 / 9, 4 / 9], rtol=1e-5 if "32" in dt else 1e-13)


# ---------------------------------------------------------------- SpMM
@pytest.mark.parametrize("dt", DT)
def test_oracle_spmm_matches_recorded_reference(orc, golden_f2, dt):
    p, i, v = mg.perturbed("convdiff7", 7, dt)
    n, N = p.size - 1, 5
    for ta in "NTC":
        for tb in "NTC":
            ldb = n + 3 if tb == "N" else N + 2
            B = mg.probe_x(ldb * (N if tb == "N" else n), dt, seed=5)
            C0 = mg.probe_x((n + 1) * N, dt, seed=6)
            got = orc.spmm(ta, tb, n, N, n, p, i, v, B, ldb, alpha=1.5, beta=0.5, C=C0, ldc=n + 1)
            want = golden_f2[f"spmm/convdiff7:7/{dt}/{ta}{tb}"]
            assert np.abs(got - want).max() <= F1_TOL[dt] * np.abs(want).max(), (ta, tb)


def test_c_stencil_generator_equals_matgen(orc):
    """bench.py's reference arm builds the 512^3 matrix with the oracle's C generator: bit-identical to hala_b200.matgen."""
    for n in (3, 7, 12):
        N = n ** 3
        for args, gen in (((-1.0, 6.0, -1.0), mg.lap3d7), ((-1.5, 6.0, -0.5), mg.convdiff7)):
            for lo, hi in ((0, N), (n * n, N - 5)):
                a = orc.gen_stencil7(n, *args, row_lo=lo, row_hi=hi)
                b = gen(n, row_lo=lo, row_hi=hi)
                assert all(np.array_equal(u, v) and u.dtype == v.dtype for u, v in zip(a, b)), (n, lo, hi)
