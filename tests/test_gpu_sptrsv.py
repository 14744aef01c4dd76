"""GPU parity (-m gpu) of SURVEY.md §8 row f1: sparse triangular solves (hb_sptrsv / hb_sptrsm) and ILU(0) (hb_ilu0) through the C ABI,
against the vectors of the reference's own tests (tests/sparse_tests.hpp:259-370, :423-440), outputs of the unmodified
reference recorded in tests/golden/ref_outputs_f1.npz, the CPU oracle at larger sizes, and size-independent properties."""
import numpy as np
import pytest

import hala_b200 as hb
from hala_b200 import matgen as mg
from helpers import DT, NP

pytestmark = pytest.mark.gpu
TOL = {"f32": 3e-4, "f64": 1e-11, "c32": 3e-4, "c64": 1e-11}


def load(e, *arrs):
    return [e.load(a) for a in arrs]


@pytest.mark.parametrize("dt", DT)
def test_reference_trsv_vectors(engine, dt):
    """tests/sparse_tests.hpp:259-311: X = 0.5 * op(T)^-1 (2 op(A) Xref) gives back Xref, L / unit-L / U, op N / T / C"""
    t = NP[dt]
    rt = 1e-5 if "32" in dt else 1e-13
    A = np.array([[1, 0, 0], [4, 2, 0], [5, 0, 3]], dtype=t)
    xref = mg.probe_x(3, dt, seed=4)
    gp, gi, gv = load(engine, np.array([0, 1, 3, 5], np.int32), np.array([0, 0, 1, 0, 2], np.int32), np.array([1, 4, 2, 5, 3], dtype=t))
    tri = hb.make_triangular_matrix(engine, "L", "N", gp, gi, gv)
    for tr, op in (("N", A), ("T", A.T), ("C", A.conj().T)):
        x = engine.new_vector(t)
        hb.sparse_trsv(tr, tri, 0.5, engine.load((2 * op @ xref).astype(t)), x)
        np.testing.assert_allclose(x.unload(), xref, rtol=rt)
    Au = np.array([[1, 0, 0], [1, 1, 0], [2, 0, 1]], dtype=t)
    gvu = engine.load(np.array([2, 1, 3, 2, 4], dtype=t))
    tri = hb.make_triangular_matrix(engine, "L", "U", gp, gi, gvu)
    x = engine.new_vector(t)
    hb.sparse_trsv("N", tri, 0.5, engine.load((2 * Au @ xref).astype(t)), x)
    np.testing.assert_allclose(x.unload(), xref, rtol=rt)
    gp, gi, gv = load(engine, np.array([0, 3, 4, 5], np.int32), np.array([0, 1, 2, 1, 2], np.int32), np.array([1, 4, 5, 2, 3], dtype=t))
    tri = hb.make_triangular_matrix(engine, "U", "N", gp, gi, gv)
    for tr, op in (("N", A.T), ("T", A)):
        x = engine.new_vector(t)
        hb.sparse_trsv(tr, tri, 0.5, engine.load((2 * op @ xref).astype(t)), x)
        np.testing.assert_allclose(x.unload(), xref, rtol=rt)


@pytest.mark.parametrize("dt", DT)
def test_reference_trsm_vectors(engine, dt):
    """tests/sparse_tests.hpp:313-370: six right-hand sides in place, B as columns (transb N) and as rows (transb T / C)"""
    t = NP[dt]
    rt = 1e-5 if "32" in dt else 1e-13
    M, nrhs = 3, 6
    A = np.array([[1, 0, 0], [4, 2, 0], [5, 0, 3]], dtype=t)
    Xref = mg.probe_x(M * nrhs, dt, seed=4).reshape(nrhs, M).T          # column-major M x nrhs
    gp, gi, gv = load(engine, np.array([0, 1, 3, 5], np.int32), np.array([0, 0, 1, 0, 2], np.int32), np.array([1, 4, 2, 5, 3], dtype=t))
    tri = hb.make_triangular_matrix(engine, "L", "N", gp, gi, gv)
    for ta, op in (("N", A), ("T", A.T), ("C", A.conj().T)):
        B = engine.load(np.ascontiguousarray((2 * op @ Xref).T).reshape(-1).astype(t))      # column-major storage
        hb.sparse_trsm(ta, "N", nrhs, tri, 0.5, B)
        np.testing.assert_allclose(B.unload().reshape(nrhs, M).T, Xref, rtol=rt)
    for tb in ("T", "C"):                                                # B = (2 A Xref)^T stored column-major: nrhs x M, ldb = nrhs
        Bt = (2 * A @ Xref).T
        B = engine.load(np.ascontiguousarray(Bt.T).reshape(-1).astype(t))
        hb.sparse_trsm("N", tb, nrhs, tri, 0.5, B)
        np.testing.assert_allclose(B.unload().reshape(M, nrhs).T, Xref.T, rtol=rt)
    gp, gi, gv = load(engine, np.array([0, 3, 4, 5], np.int32), np.array([0, 1, 2, 1, 2], np.int32), np.array([1, 4, 5, 2, 3], dtype=t))
    tri = hb.make_triangular_matrix(engine, "U", "N", gp, gi, gv)
    for ta, op in (("N", A.T), ("T", A)):
        B = engine.load(np.ascontiguousarray((2 * op @ Xref).T).reshape(-1).astype(t))
        hb.sparse_trsm(ta, "N", nrhs, tri, 0.5, B)
        np.testing.assert_allclose(B.unload().reshape(nrhs, M).T, Xref, rtol=rt)


@pytest.mark.parametrize("dt", DT)
def test_reference_ilu_vectors(engine, dt):
    """tests/sparse_tests.hpp:423-440"""
    t = NP[dt]
    rt = 1e-5 if "32" in dt else 1e-13
    gp, gi, gv = load(engine, np.array([0, 3, 6, 9], np.int32), np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], np.int32),
                      np.array([3, 2, 1, 2, 4, 3, 2, 1, 6], dtype=t))
    ilu = hb.make_ilu(engine, gp, gi, gv)
    np.testing.assert_allclose(ilu.factors(), np.array([3, 2, 1, 2 / 3, 8 / 3, 7 / 3, 2 / 3, -0.125, 5.625]), rtol=rt)
    r = engine.new_vector(t)
    ilu.apply(engine.load(np.array([1, 2, 3], dtype=t)), r)
    np.testing.assert_allclose(r.unload(), [1 / 9, 1 / 9, 4 / 9], rtol=rt)


@pytest.mark.parametrize("dt", DT)
def test_trsv_ilu_match_recorded_reference(engine, golden_f1, dt):
    for name, n in mg.F1_CASES:
        p, i, v = mg.perturbed(name, n, dt)
        N = p.size - 1
        b = mg.probe_x(N, dt, seed=31)
        gb = engine.load(b)
        gp, gi, gv = load(engine, p, i, v)
        for uplo in "LU":
            tp, ti, tv = mg.split_triangle(p, i, v, uplo)
            gtp, gti, gtv = load(engine, tp, ti, tv)
            for diag in "NU":
                split = hb.make_triangular_matrix(engine, uplo, diag, gtp, gti, gtv)     # the reference's one-triangle layout
                full = hb.make_triangular_matrix(engine, uplo, diag, gp, gi, gv)         # full matrix, triangle selected by uplo
                for tr in "NTC":
                    want = golden_f1[f"trsv/{name}:{n}/{dt}/{uplo}{diag}{tr}"]
                    for tri in (split, full):
                        x = engine.new_vector(NP[dt])
                        tri.trsv(tr, 0.5, gb, x)
                        assert np.abs(x.unload() - want).max() <= TOL[dt] * np.abs(want).max(), (name, uplo, diag, tr)
        ilu = hb.make_ilu(engine, gp, gi, gv)
        want = golden_f1[f"ilu/{name}:{n}/{dt}/factors"]
        assert np.abs(ilu.factors() - want).max() <= TOL[dt] * np.abs(want).max(), name
        r = engine.new_vector(NP[dt])
        ilu.apply(gb, r)
        want = golden_f1[f"ilu/{name}:{n}/{dt}/apply"]
        assert np.abs(r.unload() - want).max() <= TOL[dt] * np.abs(want).max(), name
        # two right-hand sides through the trsm pair
        r2 = engine.new_vector(NP[dt])
        ilu.apply(engine.load(np.concatenate([b, 2 * b])), r2, 2)
        got = r2.unload()
        assert np.abs(got[:N] - want).max() <= TOL[dt] * np.abs(want).max() and np.abs(got[N:] - 2 * want).max() <= 2 * TOL[dt] * np.abs(want).max()


@pytest.mark.parametrize("dt", ["f64", "c64"])
@pytest.mark.parametrize("name,n", [("lap3d7", 24), ("lap3d27", 14), ("lap2d", 150)])
def test_trsv_ilu_vs_oracle_larger(engine, orc, dt, name, n):
    """sizes where the level structure matters (thousands of rows per level, hundreds of levels)"""
    p, i, v = mg.perturbed(name, n, dt)
    N = p.size - 1
    b = mg.probe_x(N, dt, seed=31)
    gb = engine.load(b)
    gp, gi, gv = load(engine, p, i, v)
    vs = (v / 16).astype(v.dtype)                # unit-diagonal solves need off-diagonals that sum below one to stay bounded
    gvs = engine.load(vs)
    for uplo in "LU":
        for diag in "NU":
            tri = hb.make_triangular_matrix(engine, uplo, diag, gp, gi, gv if diag == "N" else gvs)
            for tr in "NTC":
                want = orc.trsv(uplo, diag, tr, p, i, v if diag == "N" else vs, b, alpha=1.5, general=True)
                x = engine.new_vector(NP[dt])
                tri.trsv(tr, 1.5, gb, x)
                assert np.abs(x.unload() - want).max() <= 1e-10 * np.abs(want).max(), (uplo, diag, tr)
            assert tri.levels() > 1
    fac_o, r_o = orc.ilu(p, i, v, b)
    ilu = hb.make_ilu(engine, gp, gi, gv)
    assert np.abs(ilu.factors() - fac_o).max() <= 1e-11 * np.abs(fac_o).max()
    r = engine.new_vector(NP[dt])
    ilu.apply(gb, r)
    assert np.abs(r.unload() - r_o).max() <= 1e-10 * np.abs(r_o).max()
    # property: (L U) (U^-1 L^-1 b) == b with L, U read off the factors
    rows = np.repeat(np.arange(N), np.diff(p))
    fac = ilu.factors()
    y = r.unload()
    up = i >= rows
    Uy = np.zeros(N, dtype=fac.dtype)
    np.add.at(Uy, rows[up], fac[up] * y[i[up]])
    LUy = Uy.copy()
    lo = i < rows
    np.add.at(LUy, rows[lo], fac[lo] * Uy[i[lo]])
    assert np.abs(LUy - b).max() <= 1e-10 * np.abs(b).max()


def test_ilu_rejects_missing_diagonal(engine):
    gp, gi, gv = load(engine, np.array([0, 2, 3], np.int32), np.array([0, 1, 0], np.int32), np.array([1.0, 2.0, 3.0]))
    with pytest.raises(hb.HalaB200Error):
        hb.make_ilu(engine, gp, gi, gv)


def test_trsv_empty_and_diagonal_only(engine):
    gp, gi, gv = load(engine, np.arange(6, dtype=np.int32), np.arange(5, dtype=np.int32), np.array([1.0, 2.0, 4.0, 8.0, 16.0]))
    tri = hb.make_triangular_matrix(engine, "L", "N", gp, gi, gv)
    x = engine.new_vector(np.float64)
    tri.trsv("N", 2.0, engine.load(np.ones(5)), x)
    np.testing.assert_allclose(x.unload(), 2.0 / np.array([1.0, 2.0, 4.0, 8.0, 16.0]))
    tri.trsv("T", 2.0, engine.load(np.ones(5)), x)
    np.testing.assert_allclose(x.unload(), 2.0 / np.array([1.0, 2.0, 4.0, 8.0, 16.0]))
