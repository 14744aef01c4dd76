"""Short launch list of every hot kernel at HEAD for one `ncu --set full` capture (round-2 evidence):
   spmv_pipe_kernel (27-point 128^3, plain) | CG on the 7-point 512^3 matrix, 3 iterations (spmv+dot, cg_update, cg_direction, cg_setup) |
   GMRES(24) one cycle on convdiff7 256^3 (gs_pipe<0>, gs_pipe<1>, scale_copy) | BLAS-1 on 2^26 doubles (copy, axpy, scal, dot, nrm2) |
   preconditioned CG, 2 iterations (pcg_dot) | power-law 2^22 (VS form + combine).
   ncu --set full --clock-control none -k regex:'spmv_pipe|cg_|gs_pipe|scale_copy|_vec|reduce_kernel|pcg_dot|vsplit' python scripts/ncu_hot_kernels.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg
from hala_b200.capi import lib, check, PRECON_FN

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["spmv", "cg", "gmres", "blas1", "pcg", "pl"]
e = hb.gpu_engine(0)
dev = "cuda:0"

def csr(name, n, dt="f64"):
    N = n ** 3
    tp, ti, tv = devgen.stencil_slab(name, n, 0, N, dtype=dt, device=dev)
    views = tuple(devgen.torch_view(e, t) for t in (tp, ti, tv))
    return N, (tp, ti, tv), views, hb.make_sparse_matrix(e, N, *views)

if "spmv" in which:
    N, keep, views, A = csr("lap3d27", 128)
    x = torch.from_numpy(mg.probe_x(N)).to(dev); y = torch.empty_like(x)
    for _ in range(2): A.gemv("N", 1.0, devgen.torch_view(e, x), 0.0, devgen.torch_view(e, y))
    e.synchronize(); del A, keep, views, x, y
if "cg" in which:
    N, keep, views, A = csr("lap3d7", int(os.environ.get("CG_GRID", "512")))
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_cg(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, 4, C.byref(it), C.byref(res)), "hb_cg")
    e.synchronize(); del A, keep, views, b, x
    torch.cuda.empty_cache(); lib.hb_ctx_trim(e.ctx)
if "gmres" in which:
    N, keep, views, A = csr("convdiff7", 256)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_gmres(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, 1, 24, 0, C.byref(it), C.byref(res)), "hb_gmres")
    e.synchronize(); del A, keep, views, b, x
    torch.cuda.empty_cache(); lib.hb_ctx_trim(e.ctx)
if "blas1" in which:
    n = 1 << 26
    x = e.load(mg.probe_x(n)); y = e.load(mg.probe_x(n, seed=5))
    hb.vcopy(e, x, y); hb.axpy(e, 1.5, x, y); hb.scal(e, 0.75, y); hb.dot(e, x, y); hb.norm2(e, x)
    e.synchronize(); del x, y
if "pcg" in which:
    N, keep, views, A = csr("lap3d7", 256)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    def precon(user, vin, vout):
        check(lib.hb_memcpy_async(e.ctx, C.c_void_p(vout), C.c_void_p(vin), N * 8, 2))
        return 0
    cb = PRECON_FN(precon)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_pcg(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, 3, C.cast(cb, C.c_void_p), None, C.byref(it), C.byref(res)), "hb_pcg")
    e.synchronize(); del A, keep, views, b, x
if "pl" in which:
    N = 1 << 22
    p, i, v = mg.powerlaw(N=N, dtype="f64")
    tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
    x = torch.from_numpy(mg.probe_x(N)).to(dev); y = torch.empty_like(x)
    A = hb.make_sparse_matrix(e, N, *(devgen.torch_view(e, t) for t in (tp, ti, tv)))
    for _ in range(2): A.gemv("N", 1.0, devgen.torch_view(e, x), 0.0, devgen.torch_view(e, y))
    e.synchronize()
