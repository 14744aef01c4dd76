"""SURVEY.md §8 rows f2 / f3 on one B200: SpMM (hb_spmm) against nrhs separate SpMVs, and the transposed SpMV, on the BASELINE
matrices; device-timed, one JSON line per measurement.  Not the contract bench."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                  # noqa: E402
import hala_b200 as hb                        # noqa: E402
from hala_b200 import devgen, matgen as mg    # noqa: E402

PEAK = 6542.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
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


def main():
    e = hb.gpu_engine(0)
    dev = "cuda:0"
    for name, n in (("lap3d27", 128), ("lap3d7", 256)):
        N = n ** 3
        tp, ti, tv = devgen.stencil_slab(name, n, 0, N, device=dev)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        nnz = ti.numel()
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        x, y = e.load(mg.probe_x(N)), e.new_vector(np.float64, N)
        ms1 = timeit(e, lambda: A.gemv("N", 1.0, x, 0.0, y))
        mst = timeit(e, lambda: A.gemv("T", 1.0, x, 0.0, y))
        B1 = mg.spmv_bytes(N, nnz, 8)
        emit(op="spmv N", matrix=f"{name}:{n}", us=ms1 * 1e3, gbs=B1 / ms1 / 1e6, frac_measured_peak=B1 / ms1 / 1e6 / PEAK)
        emit(op="spmv T (atomic scatter)", matrix=f"{name}:{n}", us=mst * 1e3, gbs=B1 / mst / 1e6, frac_measured_peak=B1 / mst / 1e6 / PEAK)
        for nrhs in (4, 8, 16):
            Bm, Cm = e.load(mg.probe_x(N * nrhs)), e.new_vector(np.float64, N * nrhs)
            ms = timeit(e, lambda: A.gemm("N", "N", N, nrhs, 1.0, Bm, N, 0.0, Cm, N), reps=10)
            # algorithmic bytes: the matrix once per 4 columns (the kernel's pass structure) is NOT counted: ideal = matrix once + B + C
            ideal = nnz * 12 + 4 * (N + 1) + 2 * 8 * N * nrhs
            emit(op="spmm N,N", matrix=f"{name}:{n}", nrhs=nrhs, us=ms * 1e3, speedup_vs_spmv_loop=ms1 * nrhs / ms, gbs_ideal=ideal / ms / 1e6,
                 frac_measured_peak=ideal / ms / 1e6 / PEAK, gflops=2.0 * nnz * nrhs / ms / 1e6)
            del Bm, Cm
        del A, x, y, gp, gi, gv, tp, ti, tv
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
