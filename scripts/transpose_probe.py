"""SURVEY.md §8 row f3 on one B200: op 'T' / 'C' products through the cached CSR of the transpose (hb_transpose.cu) against the atomic
scatter, per transpose mode, on the BASELINE matrices; plus the cost of a device allocation through the pooled allocator, which is
what the load -> operation -> unload pattern of mixed_engine pays per call.  Device-timed, one JSON line per measurement.
Not the contract bench."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                  # noqa: E402
import hala_b200 as hb                        # noqa: E402
from hala_b200 import devgen, matgen as mg    # noqa: E402

PEAK = 6542.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timeit(e, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e.timer_start()
    for _ in range(reps):
        fn()
    return e.timer_stop() / reps


def products(e, label, N, nnz, A, gv_tensor, dt, trans):
    es = np.dtype(dt).itemsize
    x, y = e.load(mg.probe_x(N, {"float64": "f64", "complex128": "c64"}[np.dtype(dt).name])), e.new_vector(dt, N)
    B1 = mg.spmv_bytes(N, nnz, es)
    ms = timeit(e, lambda: A.gemv("N", 1.0, x, 0.0, y))
    emit(op="spmv N", matrix=label, us=ms * 1e3, gbs=B1 / ms / 1e6, frac_measured_peak=B1 / ms / 1e6 / PEAK)
    A.set_transpose_mode("scatter")
    ms = timeit(e, lambda: A.gemv(trans, 1.0, x, 0.0, y))
    emit(op=f"spmv {trans} scatter", matrix=label, us=ms * 1e3, gbs=B1 / ms / 1e6, frac_measured_peak=B1 / ms / 1e6 / PEAK)
    A.set_transpose_mode("checked")
    e.synchronize()
    t0 = time.perf_counter()
    A.gemv(trans, 1.0, x, 0.0, y)                 # builds the transposed copy
    e.synchronize()
    emit(op="build transposed copy (first product)", matrix=label, ms=(time.perf_counter() - t0) * 1e3, bytes=A.transpose_info()["bytes"])
    ms = timeit(e, lambda: A.gemv(trans, 1.0, x, 0.0, y))
    # bytes of a checked product: the product itself plus one more read of the value array (fingerprint)
    emit(op=f"spmv {trans} checked (values unchanged)", matrix=label, us=ms * 1e3, gbs=B1 / ms / 1e6, frac_measured_peak=B1 / ms / 1e6 / PEAK,
         gbs_moved=(B1 + nnz * es) / ms / 1e6, frac_moved=(B1 + nnz * es) / ms / 1e6 / PEAK)

    def changed():
        gv_tensor[0] = gv_tensor[0] * 1.0000001   # one value rewritten by the caller -> fingerprint differs -> values re-gathered
        A.gemv(trans, 1.0, x, 0.0, y)
    ms = timeit(e, changed, reps=10)
    emit(op=f"spmv {trans} checked (values rewritten before every product)", matrix=label, us=ms * 1e3, gbs=B1 / ms / 1e6)
    A.set_transpose_mode("frozen")
    ms = timeit(e, lambda: A.gemv(trans, 1.0, x, 0.0, y))
    emit(op=f"spmv {trans} frozen", matrix=label, us=ms * 1e3, gbs=B1 / ms / 1e6, frac_measured_peak=B1 / ms / 1e6 / PEAK)
    if dt == np.float64:
        nrhs = 4
        Bm, Cm = e.load(mg.probe_x(N * nrhs)), e.new_vector(dt, N * nrhs)
        ms = timeit(e, lambda: A.gemm("T", "N", N, nrhs, 1.0, Bm, N, 0.0, Cm, N), reps=10)
        emit(op="spmm T,N frozen", matrix=label, nrhs=nrhs, us=ms * 1e3)


def main():
    quick, alloc_only = "--quick" in sys.argv, "--alloc-only" in sys.argv          # --quick: first matrix only (for ncu)
    e = hb.gpu_engine(0)
    dev = "cuda:0"
    cases = (("lap3d27", 128, np.float64), ("lap3d7", 256, np.float64), ("helmholtz7", 192, np.complex128))
    for name, n, dt in (() if alloc_only else cases[:1] if quick else cases):
        N = n ** 3
        tp, ti, tv = devgen.stencil_slab(name, n, 0, N, dtype="c64" if dt == np.complex128 else "f64", device=dev)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        products(e, f"{name}:{n}", N, ti.numel(), A, tv, dt, "C" if dt == np.complex128 else "T")
        del A, gp, gi, gv, tp, ti, tv
        torch.cuda.empty_cache()
    if quick:
        return
    if not alloc_only:
        # irregular rows: power-law 2^20 (the 2^22 matrix of configs[4] takes minutes to build on the host)
        p, i, v = mg.powerlaw(N=1 << 20, dtype="f64")
        tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        A = hb.make_sparse_matrix(e, p.size - 1, gp, gi, gv)
        products(e, "powerlaw:2^20", p.size - 1, i.size, A, tv, np.float64, "T")
        del A
    # allocation cost seen by mixed_engine-style callers: new_vector + free, and load -> axpy -> unload of 1 M doubles from pageable memory
    for nbytes in (4096, 8 << 20, 256 << 20):
        reps = 200
        e.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            vtmp = e.new_vector(np.float64, nbytes // 8)
            vtmp.clear()
        e.synchronize()
        emit(op="malloc+free", bytes=nbytes, us=(time.perf_counter() - t0) / reps * 1e6, pool=os.environ.get("HB_MEM_POOL", "1"))
    hx, hy = mg.probe_x(1 << 20), mg.probe_x(1 << 20, seed=3)
    reps = 50
    e.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        gx, gy = e.load(hx), e.load(hy)
        hb.axpy(e, 2.0, gx, gy)
        out = gy.unload()
        gx.clear(); gy.clear()
    emit(op="mixed-style axpy (load x, load y, axpy, unload y, free) 2^20 doubles", us=(time.perf_counter() - t0) / reps * 1e6,
         pool=os.environ.get("HB_MEM_POOL", "1"))


if __name__ == "__main__":
    main()
