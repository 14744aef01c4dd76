"""GPU parity (-m gpu) of the widened BLAS rows: the rest of BLAS-1 (swap, iamax, rot, rotm, rotg, rotmg)
against numpy restatements of the netlib definitions and the vectors of the reference's own tests
(tests/blas1_tests.hpp:30-41 swap, :118-130 iamax, :132-163 rotate), and the Gram-Schmidt pair at GMRES-sized column
counts (chunked multi-dot, 128-bit and element-wise paths)."""
import ctypes as C

import numpy as np
import pytest

import hala_b200 as hb
from hala_b200 import matgen as mg
from helpers import DT, NP

pytestmark = pytest.mark.gpu
TOL = {"f32": 1e-5, "f64": 1e-13, "c32": 1e-5, "c64": 1e-13}


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 5, 1000, 100003])
def test_swap(engine, dt, n):
    x, y = mg.probe_x(n, dt, seed=1), mg.probe_x(2 * n, dt, seed=2)
    gx, gy = engine.load(x), engine.load(y)
    hb.vswap(engine, gx, gy, 1, 2)                      # tests/blas1_tests.hpp:38: strides (1, 2)
    rx, ry = x.copy(), y.copy()
    rx[:], ry[::2] = y[::2], x
    assert np.array_equal(gx.unload(), rx) and np.array_equal(gy.unload(), ry)
    a, b = mg.probe_x(n, dt, seed=3), mg.probe_x(n, dt, seed=4)
    ga, gb = engine.load(a), engine.load(b)
    hb.vswap(engine, ga, gb)                            # unit stride: the 128-bit path
    assert np.array_equal(ga.unload(), b) and np.array_equal(gb.unload(), a)


@pytest.mark.parametrize("dt", DT)
def test_iamax_reference_vectors(engine, dt):
    """tests/blas1_tests.hpp:118-130: growing entries -> last index; stride 2 -> 2; then x[4] = 1 -> 3"""
    x = np.arange(1, 6).astype(NP[dt])
    if dt.startswith("c"):
        x = x * (1 + 0.5j)
    g = engine.load(x)
    assert hb.iamax(engine, g) == 4
    assert hb.iamax(engine, g, 2) == 2
    x[4] = 1.0
    assert hb.iamax(engine, engine.load(x)) == 3


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("n", [1, 33, 4097, 1 << 20])
def test_iamax_sizes(engine, dt, n):
    x = mg.probe_x(n, dt, seed=9)
    mag = np.abs(x.real) + np.abs(x.imag) if dt.startswith("c") else np.abs(x)
    assert hb.iamax(engine, engine.load(x)) == int(np.argmax(mag))
    if n > 40:                                          # ties: the FIRST of equal maxima wins
        x[7] = x[n - 3] = NP[dt](5.0)
        assert hb.iamax(engine, engine.load(x)) == 7
    assert hb.iamax(engine, engine.load(x), 1, 0) == -1  # n == 0 -> cublas returns 0, minus one


@pytest.mark.parametrize("dt", DT)
def test_rotg_rot(engine, dt):
    """tests/blas1_tests.hpp:135-143: rotg(2, 1) -> c = 2/sqrt5, s = 1/sqrt5; rot of (2, 1) vectors -> (sqrt5, 0)"""
    r, z, c, s = hb.rotg(engine, 2.0, 1.0, NP[dt])
    assert abs(c - 2 / np.sqrt(5)) < 10 * TOL[dt] and abs(s - 1 / np.sqrt(5)) < 10 * TOL[dt]
    gx, gy = engine.load(np.full(5, 2.0, NP[dt])), engine.load(np.full(5, 1.0, NP[dt]))
    hb.rot(engine, gx, gy, c, s)
    np.testing.assert_allclose(gx.unload(), np.full(5, np.sqrt(5.0)), rtol=10 * TOL[dt])
    np.testing.assert_allclose(gy.unload(), np.zeros(5), atol=10 * TOL[dt])
    # general values, strides, real s on complex vectors (cublasCsrot / cublasZdrot)
    n = 10007
    x, y = mg.probe_x(n, dt, seed=5), mg.probe_x(3 * n, dt, seed=6)
    cc, ss = 0.6, (0.8 if not dt.startswith("c") else 0.48 - 0.64j)
    for sval in ([ss] if not dt.startswith("c") else [ss, 0.8]):
        gx, gy = engine.load(x), engine.load(y)
        hb.rot(engine, gx, gy, cc, sval, 1, 3)
        rx = cc * x + sval * y[::3]
        ry = y.copy()
        ry[::3] = cc * y[::3] - np.conj(sval) * x
        np.testing.assert_allclose(gx.unload(), rx.astype(NP[dt]), rtol=20 * TOL[dt], atol=20 * TOL[dt])
        np.testing.assert_allclose(gy.unload(), ry.astype(NP[dt]), rtol=20 * TOL[dt], atol=20 * TOL[dt])
        a, b = mg.probe_x(n, dt, seed=7), mg.probe_x(n, dt, seed=8)
        ga, gb = engine.load(a), engine.load(b)
        hb.rot(engine, ga, gb, cc, sval)                # unit stride: 128-bit path
        np.testing.assert_allclose(ga.unload(), (cc * a + sval * b).astype(NP[dt]), rtol=20 * TOL[dt], atol=20 * TOL[dt])
        np.testing.assert_allclose(gb.unload(), (cc * b - np.conj(sval) * a).astype(NP[dt]), rtol=20 * TOL[dt], atol=20 * TOL[dt])


def _rotg_ref(a, b):
    roe = a if abs(a) > abs(b) else b
    scale = abs(a) + abs(b)
    if scale == 0:
        return 0.0, 0.0, 1.0, 0.0
    r = scale * np.sqrt((a / scale) ** 2 + (b / scale) ** 2) * (1 if roe >= 0 else -1)
    c, s = a / r, b / r
    z = s if abs(a) > abs(b) else (1 / c if c != 0 else 1.0)
    return r, z, c, s


@pytest.mark.parametrize("ab", [(3.0, 4.0), (-3.0, 4.0), (4.0, -3.0), (0.0, 2.0), (2.0, 0.0), (0.0, 0.0), (1e-200, 1e200)])
def test_rotg_real_cases(engine, ab):
    got = hb.rotg(engine, ab[0], ab[1], np.float64)
    np.testing.assert_allclose(got, _rotg_ref(*ab), rtol=1e-14, atol=1e-300)


def test_rotg_complex_annihilates(engine):
    a, b = 1.5 - 0.5j, -0.25 + 2.0j
    r, _, c, s = hb.rotg(engine, a, b, np.complex128)
    assert abs(c * a + s * b - r) < 1e-14 and abs(c * b - np.conj(s) * a) < 1e-14 and abs(c * c + abs(s) ** 2 - 1) < 1e-14


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_rotmg_rotm(engine, dt):
    """tests/blas1_tests.hpp:145-163: rotmg(1,1,1,0) sets flag -2 and leaves H; rotm with that param is the identity"""
    param = np.array([-1, 0, 0, 0, 0], dtype=NP[dt])
    hb.rotmg(engine, 1.0, 1.0, 1.0, 0.0, param, NP[dt])
    assert param.tolist() == [-2.0, 0.0, 0.0, 0.0, 0.0]
    x, y = mg.probe_x(3, dt, seed=4), mg.probe_x(3, dt, seed=5)
    gx, gy = engine.load(x), engine.load(y)
    hb.rotm(engine, gx, gy, param)
    assert np.array_equal(gx.unload(), x) and np.array_equal(gy.unload(), y)
    # the three non-trivial flag forms against the definition
    n = 5001
    x, y = mg.probe_x(n, dt, seed=6), mg.probe_x(n, dt, seed=7)
    for flag, H in ((-1.0, (0.5, -0.25, 2.0, 1.5)), (0.0, (1.0, -0.25, 2.0, 1.0)), (1.0, (0.5, -1.0, 1.0, 1.5))):
        h11, h21, h12, h22 = H
        gx, gy = engine.load(x), engine.load(y)
        hb.rotm(engine, gx, gy, np.array([flag, 0.5, -0.25, 2.0, 1.5]))
        np.testing.assert_allclose(gx.unload(), h11 * x + h12 * y, rtol=20 * TOL[dt], atol=20 * TOL[dt])
        np.testing.assert_allclose(gy.unload(), h21 * x + h22 * y, rtol=20 * TOL[dt], atol=20 * TOL[dt])
    # rotmg produces a transformation that zeroes the second component: H (sqrt(d1) x1, sqrt(d2) y1) -> (., 0)
    for d1, d2, x1, y1 in ((2.0, 3.0, 1.5, 0.5), (1.0, 4.0, 0.5, 2.0), (3.0, 1.0, -2.0, 0.25)):
        prm = np.zeros(5, dtype=NP[dt])
        nd1, nd2, nx1 = hb.rotmg(engine, d1, d2, x1, y1, prm, NP[dt])
        flag = prm[0]
        h11, h21, h12, h22 = prm[1], prm[2], prm[3], prm[4]
        if flag == 0:
            h11 = h22 = 1.0
        elif flag == 1:
            h12, h21 = 1.0, -1.0
        assert abs(h21 * x1 + h22 * y1) < 50 * TOL[dt]
        assert abs((h11 * x1 + h12 * y1) - nx1) < 50 * TOL[dt] * max(1.0, abs(nx1))
        assert abs(nd1 * nx1 * nx1 - (d1 * x1 * x1 + d2 * y1 * y1)) < 100 * TOL[dt] * (d1 * x1 * x1 + d2 * y1 * y1)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(4099, 50, 4099), (4096, 23, 4112), (1001, 17, 1003), (70001, 64, 70016)])
def test_gram_schmidt_pair_gmres_sizes(engine, dt, shape):
    """multi-dot (chunks of 16 / 8 columns) and multi-axpy+norm at restart-sized k; (rows, k, ldw): aligned and unaligned ldw"""
    from hala_b200.capi import lib, check
    rows, k, ldw = shape
    code = {"f32": 0, "f64": 1, "c32": 2, "c64": 3}[dt]
    W = mg.probe_x(ldw * k, dt, seed=61).reshape(k, ldw)
    r = mg.probe_x(rows, dt, seed=62)
    gW, gr, gh = engine.load(W.reshape(-1)), engine.load(r), engine.new_vector(NP[dt], k + 1)
    tol = 2e-4 if "32" in dt else 1e-11
    scale = np.abs(W[:, :rows]).astype(np.float64) @ np.abs(r).astype(np.float64)
    for conj in (0, 1):
        check(lib.hb_multi_dot(engine.ctx, code, conj, rows, k, gW.ptr, ldw, gr.ptr, gh.ptr))
        href = ((np.conj(W[:, :rows]) if conj else W[:, :rows]).astype(np.complex128 if dt.startswith("c") else np.float64) @ r)
        assert np.all(np.abs(gh.unload()[:k] - href) <= tol * scale), (dt, shape, conj)
    check(lib.hb_multi_axpy_nrm2(engine.ctx, code, rows, k, gW.ptr, ldw, gh.ptr, gr.ptr, gh.offset(k)))
    h = gh.unload()[:k]
    rref = r - W[:, :rows].T.astype(h.dtype) @ h
    bound = tol * (np.abs(r) + np.abs(W[:, :rows]).T.astype(np.float64) @ np.abs(h))
    assert np.all(np.abs(gr.unload() - rref) <= bound), (dt, shape)
    np.testing.assert_allclose(gh.unload()[k].real, np.vdot(rref, rref).real, rtol=50 * tol)
