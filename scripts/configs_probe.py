"""BASELINE configs[3] and [4] on one B200: GMRES(50) on convdiff7-256, SpMV + GMRES(30) on helmholtz7-192 (complex<double>),
SpMV on the power-law matrix (row binning stress).  One JSON line per measurement."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg
from hala_b200.capi import lib, check

PEAK = 6542.7
e = hb.gpu_engine(0)
dev = "cuda:0"
def emit(**kw): print(json.dumps(kw), flush=True)

def spmv_time(A, x, y, reps=50):
    for _ in range(3): A.gemv("N", 1.0, x, 0.0, y)
    e.timer_start()
    for _ in range(reps): A.gemv("N", 1.0, x, 0.0, y)
    return e.timer_stop() / reps

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c4", "c5a", "c5b"]
if "c4" in which:
    n = int(os.environ.get("C4_N", "256")); N = n ** 3
    tp, ti, tv = devgen.stencil_slab("convdiff7", n, 0, N, device=dev)
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    A = hb.make_sparse_matrix(e, N, gp, gi, gv)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    gb, gx = devgen.torch_view(e, b), devgen.torch_view(e, x)
    it, res = C.c_int(0), C.c_double(0)
    check(lib.hb_gmres(e.ctx, A.h, gb.ptr, gx.ptr, 1e-8, 2, 50, 0, C.byref(it), C.byref(res)))      # warm-up: 2 cycles
    x.zero_()
    t0 = time.perf_counter()
    check(lib.hb_gmres(e.ctx, A.h, gb.ptr, gx.ptr, 1e-8, int(os.environ.get("C4_OUTER", "40")), 50, 0, C.byref(it), C.byref(res)))
    dt = time.perf_counter() - t0
    r = b.clone(); gr = devgen.torch_view(e, r)
    A.gemv("N", -1.0, gx, 1.0, gr)
    # bytes per inner iteration at basis size k: matrix + 2Ns + (2k+1+3) N s ; average k ~ 25.5
    nnz = ti.numel(); kavg = 25.5
    bytes_it = 12 * nnz + 4 * (N + 1) + 8 * N * (2 + 2 * kavg + 4)
    emit(op="gmres50", config="c4 convdiff7-%d" % n, iters=it.value, est_res=res.value, true_res=float(torch.linalg.norm(r)), wall_s=dt,
         its_per_s=it.value / dt, gbs=bytes_it * it.value / dt / 1e9, frac_measured_peak=bytes_it * it.value / dt / 1e9 / PEAK)
    del A, tp, ti, tv, b, x, r; lib.hb_ctx_trim(e.ctx); torch.cuda.empty_cache()
if "c5a" in which:
    n = int(os.environ.get("C5_N", "192")); N = n ** 3
    tp, ti, tv = devgen.stencil_slab("helmholtz7", n, 0, N, dtype="c64", device=dev)
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    A = hb.make_sparse_matrix(e, N, gp, gi, gv)
    xs = torch.from_numpy(mg.probe_x(N, "c64")).to(dev); ys = torch.empty_like(xs)
    ms = spmv_time(A, devgen.torch_view(e, xs), devgen.torch_view(e, ys))
    B = mg.spmv_bytes(N, ti.numel(), 16)
    emit(op="spmv", config="c5a helmholtz7-%d c64" % n, us=ms * 1e3, gbs=B / ms / 1e6, frac_measured_peak=B / ms / 1e6 / PEAK, gflops=8 * ti.numel() / ms / 1e6)
    b = torch.from_numpy(mg.rhs(N, "c64")).to(dev); x = torch.zeros_like(b)
    gb, gx = devgen.torch_view(e, b), devgen.torch_view(e, x)
    it, res = C.c_int(0), C.c_double(0)
    t0 = time.perf_counter()
    check(lib.hb_gmres(e.ctx, A.h, gb.ptr, gx.ptr, 1e-8, 1000, 30, 1, C.byref(it), C.byref(res)))
    dt = time.perf_counter() - t0
    r = b.clone(); gr = devgen.torch_view(e, r)
    A.gemv("N", -1.0, gx, 1.0, gr)
    emit(op="gmres30", config="c5a helmholtz7-%d c64" % n, iters=it.value, est_res=res.value, true_res=float(torch.linalg.norm(r)), wall_s=dt, its_per_s=it.value / dt)
    del A, tp, ti, tv, b, x, r, xs, ys; lib.hb_ctx_trim(e.ctx); torch.cuda.empty_cache()
if "c5b" in which:
    for dt_name in ("f64", "c64"):
        t0 = time.time()
        N = 1 << int(os.environ.get("C5B_LOG2N", "22"))
        p, i, v = mg.powerlaw(N=N, dtype=dt_name)
        gen = time.time() - t0
        gp, gi, gv = e.load(p), e.load(i), e.load(v)
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        x = e.load(mg.probe_x(N, dt_name)); y = e.new_vector(v.dtype, N)
        B = mg.spmv_bytes(N, i.size, v.dtype.itemsize)
        for variant in (1, 2, 3):
            A.set_variant(variant)
            ms = spmv_time(A, x, y, reps=20)
            emit(op="spmv", config="c5b powerlaw 2^22 " + dt_name, variant=variant, nnz=int(i.size), max_row=A.max_row_nnz(), us=ms * 1e3, gbs=B / ms / 1e6,
                 frac_measured_peak=B / ms / 1e6 / PEAK, gen_s=gen)
        A.set_variant(0)
        # parity on a sample of rows (the longest ones included) against left-to-right numpy sums
        yh = y.unload(); xh = mg.probe_x(N, dt_name)
        lens = np.diff(p); rows = np.concatenate([np.argsort(lens)[-5:], np.arange(0, N, N // 200)])
        worst = 0.0
        for r in rows:
            sl = slice(p[r], p[r + 1])
            ref = np.sum(v[sl] * xh[i[sl]]); scale = np.sum(np.abs(v[sl]) * np.abs(xh[i[sl]]))
            worst = max(worst, abs(yh[r] - ref) / scale)
        emit(op="spmv_parity", config="c5b " + dt_name, worst_scaled_err=worst)
        b = e.load(mg.rhs(N, dt_name)); gx = e.new_vector(v.dtype)
        t0 = time.perf_counter()
        it, res = hb.solve_gmres(e, 1e-8, 1000, 30, gp, gi, gv, b, gx, matrix=A)
        emit(op="gmres30", config="c5b " + dt_name, iters=it, est_res=res, wall_s=time.perf_counter() - t0)
