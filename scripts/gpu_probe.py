"""Development probe (GPU box): times SpMV variants, BLAS-1 streams and the fused CG iteration on the BASELINE configs
with CUDA events, prints one JSON line per measurement. Not the contract bench (that is bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hala_b200 as hb                      # noqa: E402
from hala_b200 import matgen as mg          # noqa: E402

PEAK = 6542.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


REPS, WARM = 20, 3


def timeit(e, fn, reps=None, warm=None):
    reps = REPS if reps is None else reps
    warm = WARM if warm is None else warm
    for _ in range(warm):
        fn()
    e.timer_start()
    for _ in range(reps):
        fn()
    return e.timer_stop() / reps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2")
    ap.add_argument("--dtypes", default="f64")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--no-blas", action="store_true")
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--variants", default="1,2,3")
    args = ap.parse_args()
    global REPS, WARM
    REPS, WARM = args.reps, args.warm
    e = hb.gpu_engine(0)
    # BLAS-1 streams, 64M doubles
    n = 1 << 26 if not args.no_blas else 1 << 10
    x, y = e.vector(n, 1.0), e.vector(n, 2.0)
    ms = timeit(e, lambda: hb.vcopy(e, x, y)); emit(op="copy", n=n, ms=ms, gbs=16 * n / ms / 1e6, frac=16 * n / ms / 1e6 / PEAK)
    ms = timeit(e, lambda: hb.axpy(e, 0.5, x, y)); emit(op="axpy", n=n, ms=ms, gbs=24 * n / ms / 1e6, frac=24 * n / ms / 1e6 / PEAK)
    ms = timeit(e, lambda: hb.scal(e, 1.0001, y)); emit(op="scal", n=n, ms=ms, gbs=16 * n / ms / 1e6, frac=16 * n / ms / 1e6 / PEAK)
    ms = timeit(e, lambda: hb.dot(e, x, y)); emit(op="dot(host sync)", n=n, ms=ms, gbs=16 * n / ms / 1e6, frac=16 * n / ms / 1e6 / PEAK)
    ms = timeit(e, lambda: hb.norm2(e, x)); emit(op="nrm2(host sync)", n=n, ms=ms, gbs=8 * n / ms / 1e6, frac=8 * n / ms / 1e6 / PEAK)
    del x, y
    cfgs = {"c1": ("lap2d", 1024), "c2": ("lap3d27", 128), "c2s": ("lap3d27", 64), "c3s": ("lap3d7", 256), "c4s": ("convdiff7", 128),
            "c5a": ("helmholtz7", 96)}
    for c in args.configs.split(","):
        for dt in args.dtypes.split(","):
            name, size = cfgs[c]
            if name == "helmholtz7" and not dt.startswith("c"):
                continue
            t0 = time.time()
            p, i, v = mg.GENERATORS[name](size, dtype=dt)
            N, nnz = p.size - 1, i.size
            gen_s = time.time() - t0
            gp, gi, gv = e.load(p), e.load(i), e.load(v)
            A = hb.make_sparse_matrix(e, N, gp, gi, gv)
            gx, gy = e.load(mg.probe_x(N, dt)), e.new_vector(v.dtype, N)
            B = mg.spmv_bytes(N, nnz, v.dtype.itemsize)
            for variant in [int(t) for t in args.variants.split(",")]:
                A.set_variant(variant)
                ms = timeit(e, lambda: A.gemv("N", 1.0, gx, 0.0, gy))
                emit(op="spmv", config=c, matrix=f"{name}:{size}", dtype=dt, variant=variant, N=N, nnz=nnz, ms=ms, gbs=B / ms / 1e6,
                     frac_measured_peak=B / ms / 1e6 / PEAK, frac_8tbs=B / ms / 1e6 / 8000.0, gflops=2 * nnz / ms / 1e6, gen_s=gen_s)
            A.set_variant(0)
            if args.no_cg:
                continue
            gb = e.load(mg.rhs(N, dt))
            for iters in (200,):
                gxx = e.new_vector(v.dtype)
                hb.solve_cg(e, 0.0, 20, gp, gi, gv, gb, gxx, matrix=A)
                gxx = e.new_vector(v.dtype)
                e.timer_start()
                it, res = hb.solve_cg(e, 0.0, iters + 1, gp, gi, gv, gb, gxx, matrix=A)
                ms = e.timer_stop()
                Bcg = mg.cg_iter_bytes(N, nnz, v.dtype.itemsize)
                emit(op="cg_fixed_iters", config=c, dtype=dt, iters=it - 1, ms_total=ms, its_per_s=(it - 1) / ms * 1e3,
                     gbs=Bcg * (it - 1) / ms / 1e6, frac_measured_peak=Bcg * (it - 1) / ms / 1e6 / PEAK)
            if name != "helmholtz7":
                gxx = e.new_vector(v.dtype)
                t0 = time.time()
                it, res = hb.solve_cg(e, 1e-8 if "64" in dt else 1e-4, 10 ** 6, gp, gi, gv, gb, gxx, matrix=A)
                emit(op="cg_to_tol", config=c, dtype=dt, iters=it, res=res, wall_s=time.time() - t0, its_per_s=it / (time.time() - t0))


if __name__ == "__main__":
    main()
