"""GPU parity tests of the op 'T' / 'C' products on the cached CSR of the transpose (reference:
cusparseSpMV with a transposed operation behind gpu_sparse_matrix::gemv, gpu/hala_cuda_sparse_general.hpp:264-277, pinned by
tests/sparse_tests.hpp:184-190; CPU twin sparse/hala_sparse_utils.hpp:110-117).  Checker: the CPU oracle (oracle/hb_oracle.c),
which scatters row by row exactly as the reference does.  Tolerances: helpers.SPMV_TOL (1e-13 fp64 / 1e-5 fp32, per entry,
relative to sum |a_ij||x_i|)."""
import numpy as np
import pytest

import hala_b200 as hb
from hala_b200 import matgen as mg
from helpers import DT, NP, SPMV_TOL, assert_entrywise, spmv_scale

pytestmark = pytest.mark.gpu

MODES = ["checked", "frozen", "scatter"]


def ragged_csr(rng, rows, cols, dt, long_cols=()):
    """random rectangular CSR with empty rows, empty columns and a few very long columns (entries of a long column come from
    every row, so the per-column sort of the build has to reorder them: short sort <= 32 < block-wide sort)"""
    lens = rng.integers(0, 9, rows)
    lens[rng.choice(rows, rows // 5, replace=False)] = 0
    usable = np.setdiff1d(np.arange(cols), np.arange(3, cols, 7))          # every 7th column stays empty
    pieces = []
    for r in range(rows):
        c = rng.choice(usable, lens[r], replace=False)
        extra = [lc for lc, every in long_cols if r % every == 0]
        pieces.append(np.unique(np.concatenate([c, np.array(extra, dtype=c.dtype)])))
    pntr = np.zeros(rows + 1, dtype=np.int32)
    pntr[1:] = np.cumsum([p.size for p in pieces])
    indx = np.concatenate(pieces).astype(np.int32)
    vals = rng.uniform(-1, 1, indx.size)
    if dt in ("c32", "c64"):
        vals = vals + 1j * rng.uniform(-1, 1, indx.size)
    return pntr, indx, vals.astype(NP[dt])


def check_product(engine, orc, A, p, i, v, ncols, tr, dt, what, alpha=1.5, beta=-0.5):
    M = p.size - 1
    x = mg.probe_x(M, dt, seed=11)
    y0 = mg.probe_x(ncols, dt, seed=12)
    gy = engine.load(y0)
    A.gemv(tr, alpha, engine.load(x), beta, gy)
    ref = orc.spmv(p, i, v, x, alpha=alpha, beta=beta, y=y0, trans=tr, ncols=ncols)
    assert_entrywise(gy.unload(), ref, spmv_scale(p, i, v, x, tr, ncols=ncols, alpha=alpha, beta=beta, y0=y0), SPMV_TOL[dt], what)
    # beta == 0: y holds NaN and must not be read (cuSPARSE semantics the reference relies on)
    gy = engine.load(np.full(ncols, np.nan, dtype=NP[dt]))
    A.gemv(tr, alpha, engine.load(x), 0.0, gy)
    ref = orc.spmv(p, i, v, x, alpha=alpha, trans=tr, ncols=ncols)
    assert_entrywise(gy.unload(), ref, spmv_scale(p, i, v, x, tr, ncols=ncols, alpha=alpha), SPMV_TOL[dt], what + " beta=0")


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mode", MODES)
def test_transposed_products_all_modes(engine, orc, dt, mode):
    rng = np.random.default_rng(3)
    cases = [("lap3d27:12",) + mg.lap3d27(12, dtype=dt) + (12 ** 3,),
             ("convdiff7:14",) + mg.convdiff7(14, dtype=dt) + (14 ** 3,),
             ("powerlaw:3000",) + mg.powerlaw(N=3000, lmax=600, dtype=dt) + (3000,)]
    p, i, v = ragged_csr(rng, 900, 500, dt, long_cols=((5, 1), (6, 3), (250, 20)))     # column 5: 900 entries, 6: 300, 250: 45
    cases.append(("ragged 900x500", p, i, v, 500))
    p, i, v = ragged_csr(rng, 40, 2000, dt)
    cases.append(("ragged 40x2000", p, i, v, 2000))
    for name, p, i, v, ncols in cases:
        gp, gi, gv = engine.load(p), engine.load(i), engine.load(v)
        A = hb.make_sparse_matrix(engine, p.size - 1, ncols, i.size, gp, gi, gv)
        A.set_transpose_mode(mode)
        for tr in ("T", "C", "T"):          # alternating T / C re-gathers (conjugated copy) on complex data
            check_product(engine, orc, A, p, i, v, ncols, tr, dt, f"{name} {dt} {mode} {tr}")
        info = A.transpose_info()
        assert info["mode"] == mode
        assert info["built"] == (mode != "scatter")
        if mode != "scatter":
            assert info["bytes"] == 4 * (ncols + 1) + i.size * (8 + v.itemsize)
        # op 'N' is untouched by all of this
        x = mg.probe_x(ncols, dt, seed=5)
        gy = engine.new_vector(NP[dt])
        A.gemv("N", 1.0, engine.load(x), 0.0, gy)
        assert_entrywise(gy.unload(), orc.spmv(p, i, v, x, ncols=ncols), spmv_scale(p, i, v, x, "N", ncols=ncols), SPMV_TOL[dt], name + " N")


@pytest.mark.parametrize("dt", DT)
def test_value_changes_are_seen(engine, orc, dt):
    """The matrix is a non-owning view (reference :186-190): the caller may rewrite the values between two products.
    checked: noticed by the fingerprint; frozen: served from the cached copy until hb_csr_values_changed()."""
    p, i, v = mg.lap3d27(10, dtype=dt)
    n = p.size - 1
    gp, gi, gv = engine.load(p), engine.load(i), engine.load(v)
    A = hb.make_sparse_matrix(engine, n, gp, gi, gv)
    check_product(engine, orc, A, p, i, v, n, "T", dt, "before")
    rng = np.random.default_rng(17)
    # (a) every value changes
    v2 = (v * rng.uniform(0.5, 1.5, v.size)).astype(NP[dt])
    gv.load(v2)
    check_product(engine, orc, A, p, i, v2, n, "T", dt, "all values changed")
    # (b) one value changes in its last bit; (c) two values swap places
    v3 = v2.copy()
    k = v3.size // 3
    v3[k] = np.nextafter(v3[k].real, np.inf).astype(v3.real.dtype) + 1j * v3[k].imag if np.iscomplexobj(v3) else np.nextafter(v3[k], np.inf)
    gv.load(v3)
    x = np.zeros(n, dtype=NP[dt])
    row = int(np.searchsorted(p, k, side="right") - 1)
    x[row] = 1.0                                              # y[col] = v[k] exactly: shows the one-ulp change
    gy = engine.new_vector(NP[dt])
    A.gemv("T", 1.0, engine.load(x), 0.0, gy)
    assert gy.unload()[i[k]] == v3[k]
    v4 = v3.copy()
    a, b = p[row], p[row] + 1
    v4[a], v4[b] = v3[b], v3[a] + NP[dt](0.25)
    gv.load(v4)
    check_product(engine, orc, A, p, i, v4, n, "C", dt, "two values changed")
    # frozen: the old copy is served until the caller says so
    A.set_transpose_mode("frozen")
    check_product(engine, orc, A, p, i, v4, n, "T", dt, "frozen")
    gv.load(v)
    gy = engine.new_vector(NP[dt])
    xx = mg.probe_x(n, dt, seed=11)
    A.gemv("T", 1.0, engine.load(xx), 0.0, gy)
    assert_entrywise(gy.unload(), orc.spmv(p, i, v4, xx, trans="T"), spmv_scale(p, i, v4, xx, "T"), SPMV_TOL[dt], "frozen serves the copy")
    A.values_changed()
    check_product(engine, orc, A, p, i, v, n, "T", dt, "after values_changed")
    # back to checked, then to scatter (gives the memory back), then to checked again (rebuilds)
    A.set_transpose_mode("checked")
    check_product(engine, orc, A, p, i, v, n, "C", dt, "checked again")
    A.set_transpose_mode("scatter")
    assert not A.transpose_info()["built"]
    check_product(engine, orc, A, p, i, v, n, "T", dt, "scatter")
    A.set_transpose_mode("checked")
    check_product(engine, orc, A, p, i, v, n, "T", dt, "rebuilt")
    assert A.transpose_info()["built"]


def test_transposed_product_is_reproducible_and_ordered(engine, orc):
    """Entries of a column are kept in row order, so the cached product sums a column in the order of the reference's CPU
    scatter; two independent builds give identical bits (the fill uses atomics, the per-column sort removes their order)."""
    p, i, v = mg.powerlaw(N=20000, lmax=3000, dtype="f64")
    n = p.size - 1
    gp, gi, gv = engine.load(p), engine.load(i), engine.load(v)
    x = mg.probe_x(n, "f64", seed=3)
    outs = []
    for _ in range(3):
        A = hb.make_sparse_matrix(engine, n, gp, gi, gv)
        gy = engine.new_vector(np.float64)
        A.gemv("T", 1.0, engine.load(x), 0.0, gy)
        outs.append(gy.unload())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert_entrywise(outs[0], orc.spmv(p, i, v, x, trans="T"), spmv_scale(p, i, v, x, "T"), 1e-13, "powerlaw T")


@pytest.mark.parametrize("dt", ["f64", "c64"])
def test_transposed_spmm_through_the_cached_copy(engine, orc, dt):
    """gpu_sparse_matrix::gemm with op(A) = A^T / A^H (reference :302-332 -> cusparseSpMM): the multi right-hand-side kernel on the
    cached transpose."""
    p, i, v = mg.convdiff7(12, dtype=dt)
    n = p.size - 1
    gp, gi, gv = engine.load(p), engine.load(i), engine.load(v)
    A = hb.make_sparse_matrix(engine, n, gp, gi, gv)
    nrhs = 6
    B = np.stack([mg.probe_x(n, dt, seed=20 + k) for k in range(nrhs)], axis=1)      # n x nrhs, column-major below
    Bf = np.asfortranarray(B)
    for ta in ("T", "C"):
        gC = engine.new_vector(NP[dt])
        A.gemm(ta, "N", n, nrhs, 2.0, engine.load(Bf.ravel(order="F")), n, 0.0, gC, n)
        out = gC.unload().reshape((n, nrhs), order="F")
        for k in range(nrhs):
            ref = orc.spmv(p, i, v, B[:, k], alpha=2.0, trans=ta)
            assert_entrywise(out[:, k], ref, spmv_scale(p, i, v, B[:, k], ta, alpha=2.0), SPMV_TOL[dt], f"spmm {ta} col {k}")
    assert A.transpose_info()["built"]


def test_transposed_product_at_size(engine):
    """configs[1] size (27-point 128^3, symmetric): A^T x == A x entry for entry up to summation order; timing is bench material
    (scripts/next_rows_probe.py), here only the property."""
    from hala_b200 import devgen
    n = 128
    N = n ** 3
    tensors = devgen.stencil_slab("lap3d27", n, 0, N, device="cuda:0")
    gp, gi, gv = (devgen.torch_view(engine, t) for t in tensors)
    A = hb.make_sparse_matrix(engine, N, gp, gi, gv)
    x = engine.load(mg.probe_x(N, "f64", seed=9))
    yn, yt = engine.new_vector(np.float64), engine.new_vector(np.float64)
    A.gemv("N", 1.0, x, 0.0, yn)
    A.gemv("T", 1.0, x, 0.0, yt)
    a, b = yn.unload(), yt.unload()
    assert np.max(np.abs(a - b)) <= 1e-13 * 52.0
    info = A.transpose_info()
    assert info["built"] and info["bytes"] == 4 * (N + 1) + 16 * gi.size()
