"""Small launches of every kernel family for compute-sanitizer (racecheck / memcheck / synccheck): the streaming SpMV in its general and
virtual-row forms (shared-memory ring written by bulk copies, products written back by the consumers), fused CG, preconditioned CG,
GMRES (TMA-staged Gram-Schmidt), BLAS-1.   compute-sanitizer --tool racecheck python scripts/sanitize_probe.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hala_b200 as hb
from hala_b200 import matgen as mg
from hala_b200.capi import lib, check, PRECON_FN

e = hb.gpu_engine(0)
vp = C.c_void_p
def csr(p, i, v):
    g = (e.load(p), e.load(i), e.load(v))
    return g, hb.make_sparse_matrix(e, p.size - 1, *g)

for dt in ("f64", "c64"):
    p, i, v = mg.powerlaw(N=5000, lmax=1200, dtype=dt)            # VS form
    g, A = csr(p, i, v)
    x = e.load(mg.probe_x(5000, dt)); y = e.new_vector(v.dtype, 5000); d = e.new_vector(v.dtype, 4)
    A.gemv("N", 1.0, x, 0.0, y)
    check(lib.hb_spmv_dot(e.ctx, A.h, x.ptr, y.ptr, d.ptr))
    A.gemv("T", 1.0, x, 0.0, y)
p, i, v = mg.lap3d27(14)                                           # general form, TPR 4
g, A = csr(p, i, v)
N = p.size - 1
x = e.load(mg.probe_x(N)); y = e.new_vector(np.float64, N)
A.gemv("N", 2.0, x, 0.0, y)
gx = e.new_vector(np.float64)
print("cg", hb.solve_cg(e, 1e-8, 1000, *g, e.load(mg.rhs(N)), gx, matrix=A))
def precon(user, vin, vout):
    check(lib.hb_memcpy_async(e.ctx, vp(vout), vp(vin), N * 8, 2)); return 0
cb = PRECON_FN(precon)
it, res = C.c_int(0), C.c_double(0)
gx2 = e.new_vector(np.float64, N); gx2.fill(0.0)
gb2 = e.load(mg.rhs(N))                                            # (kept alive: a temporary would be freed before the solve reads it)
check(lib.hb_pcg(e.ctx, A.h, gb2.ptr, gx2.ptr, 1e-8, 1000, C.cast(cb, vp), None, C.byref(it), C.byref(res)))
print("pcg", it.value, res.value)
pc, ic, vc = mg.convdiff7(12)
gc, Ac = csr(pc, ic, vc)
gxc = e.new_vector(np.float64)
print("gmres", hb.solve_gmres(e, 1e-8, 100, 20, *gc, e.load(mg.rhs(12 ** 3)), gxc))
a, b = e.load(mg.probe_x(100003)), e.load(mg.probe_x(100003, seed=9))
hb.axpy(e, 1.5, a, b); hb.scal(e, 0.5, b); print("dot", hb.dot(e, a, b), hb.norm2(e, a))
e.synchronize()
print("sanitize probe done")
