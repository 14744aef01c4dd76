"""Minimal launch list of every hot kernel at HEAD for one `ncu --set full` capture (round-2 evidence): each kernel ONCE or twice.
   spmv   : spmv_pipe_kernel plain, 27-point 128^3 (configs[1])
   cg     : hb_cg, two iterations (run with HB_CG_BATCH=1) on the 7-point CG_GRID^3 matrix: spmv_pipe (A x0), cg_setup, spmv_pipe DOT, cg_update, cg_direction
   gs     : hb_multi_dot + hb_multi_axpy_nrm2 on a 256^3 x 32 basis (gs_pipe_kernel<0/1>), scale_copy through one GMRES(2) cycle
   blas1  : copy, axpy, scal, dot, nrm2 on 2^26 doubles
   pcg    : hb_pcg, one iteration (pcg_dot_kernel; the preconditioner is a device-to-device copy)
   pl     : power-law 2^22 rows (VS form of spmv_pipe_kernel + vsplit_combine)
On the GPU box:  ncu --set full --clock-control none -k regex:'spmv_pipe|cg_|gs_pipe|scale_copy|_vec|reduce_kernel|pcg_dot|vsplit' -o /tmp/r2_hot
                 python scripts/ncu_hot_kernels.py ; python scripts/ncu_digest.py /tmp/r2_hot.ncu-rep gpurun_out/r2_hot   (the report stays there)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg
from hala_b200.capi import lib, check, PRECON_FN

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["spmv", "cg", "gs", "blas1", "pcg", "pl"]
e = hb.gpu_engine(0)
dev = "cuda:0"
vp = C.c_void_p

def csr(name, n, dt="f64"):
    N = n ** 3
    tp, ti, tv = devgen.stencil_slab(name, n, 0, N, dtype=dt, device=dev)
    views = tuple(devgen.torch_view(e, t) for t in (tp, ti, tv))
    return N, (tp, ti, tv), views, hb.make_sparse_matrix(e, N, *views)

def cleanup():
    torch.cuda.empty_cache(); lib.hb_ctx_trim(e.ctx)

if "spmv" in which:
    N, keep, views, A = csr("lap3d27", 128)
    x = torch.from_numpy(mg.probe_x(N)).to(dev); y = torch.empty_like(x)
    A.gemv("N", 1.0, devgen.torch_view(e, x), 0.0, devgen.torch_view(e, y))
    e.synchronize(); del A, keep, views, x, y; cleanup()
if "cg" in which:
    N, keep, views, A = csr("lap3d7", int(os.environ.get("CG_GRID", "256")))
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_cg(e.ctx, A.h, vp(b.data_ptr()), vp(x.data_ptr()), 0.0, 3, C.byref(it), C.byref(res)), "hb_cg")     # 2 iterations: the first is a full one
    e.synchronize(); del A, keep, views, b, x; cleanup()
if "gs" in which:
    n, k = 256 ** 3, 32
    ldw = n + 262144 + 256                 # column stride as hb_gmres pads it (an odd number of 2 MiB pages)
    W = torch.rand(ldw * k, dtype=torch.float64, device=dev); r = torch.rand(n, dtype=torch.float64, device=dev)
    h = torch.zeros(k + 2, dtype=torch.float64, device=dev)
    check(lib.hb_multi_dot(e.ctx, 1, 0, n, k, vp(W.data_ptr()), ldw, vp(r.data_ptr()), vp(h.data_ptr())), "hb_multi_dot")
    h.mul_(1e-3)
    check(lib.hb_multi_axpy_nrm2(e.ctx, 1, n, k, vp(W.data_ptr()), ldw, vp(h.data_ptr()), vp(r.data_ptr()), vp(h.data_ptr() + 8 * k)), "hb_multi_axpy_nrm2")
    e.synchronize(); del W, r, h; cleanup()
    N, keep, views, A = csr("convdiff7", 128)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_gmres(e.ctx, A.h, vp(b.data_ptr()), vp(x.data_ptr()), 0.0, 1, 2, 0, C.byref(it), C.byref(res)), "hb_gmres")
    e.synchronize(); del A, keep, views, b, x; cleanup()
if "blas1" in which:
    n = 1 << 26
    x = e.load(mg.probe_x(n)); y = e.load(mg.probe_x(n, seed=5))
    hb.vcopy(e, x, y); hb.axpy(e, 1.5, x, y); hb.scal(e, 0.75, y); hb.dot(e, x, y); hb.norm2(e, x)
    e.synchronize(); del x, y; cleanup()
if "pcg" in which:
    N, keep, views, A = csr("lap3d7", 256)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    def precon(user, vin, vout):
        check(lib.hb_memcpy_async(e.ctx, vp(vout), vp(vin), N * 8, 2))
        return 0
    cb = PRECON_FN(precon)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_pcg(e.ctx, A.h, vp(b.data_ptr()), vp(x.data_ptr()), 0.0, 2, C.cast(cb, vp), None, C.byref(it), C.byref(res)), "hb_pcg")
    e.synchronize(); del A, keep, views, b, x; cleanup()
if "pl" in which:
    N = 1 << 22
    p, i, v = mg.powerlaw(N=N, dtype="f64")
    tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
    x = torch.from_numpy(mg.probe_x(N)).to(dev); y = torch.empty_like(x)
    A = hb.make_sparse_matrix(e, N, *(devgen.torch_view(e, t) for t in (tp, ti, tv)))
    A.gemv("N", 1.0, devgen.torch_view(e, x), 0.0, devgen.torch_view(e, y))
    e.synchronize()
