"""SURVEY.md §8 row f2 on one B200: hb_spmm (multi right-hand-side CSR product) on one matrix, device-timed, one JSON line per
measurement; small enough to sit under ncu (`--few` = five calls only).  Not the contract bench.
   python scripts/spmm_probe.py lap3d27 128 4 8 16 [--few]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hala_b200 as hb                        # noqa: E402
from hala_b200 import devgen, matgen as mg    # noqa: E402

PEAK = 6542.7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    few = "--few" in sys.argv
    name, n = args[0], int(args[1])
    cols = [int(a) for a in args[2:]] or [4]
    e = hb.gpu_engine(0)
    N = n ** 3
    tp, ti, tv = devgen.stencil_slab(name, n, 0, N, device="cuda:0")
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    nnz = ti.numel()
    A = hb.make_sparse_matrix(e, N, gp, gi, gv)
    x, y = e.load(mg.probe_x(N)), e.new_vector(np.float64, N)
    reps, warm = (2, 1) if few else (20, 3)

    def timeit(fn):
        for _ in range(warm):
            fn()
        e.timer_start()
        for _ in range(reps):
            fn()
        return e.timer_stop() / reps
    ms1 = timeit(lambda: A.gemv("N", 1.0, x, 0.0, y))
    print(json.dumps({"op": "spmv N", "matrix": f"{name}:{n}", "us": ms1 * 1e3}), flush=True)
    for nrhs in cols:
        Bm, Cm = e.load(mg.probe_x(N * nrhs)), e.new_vector(np.float64, N * nrhs)
        ms = timeit(lambda: A.gemm("N", "N", N, nrhs, 1.0, Bm, N, 0.0, Cm, N))
        ideal = nnz * 12 + 4 * (N + 1) + 2 * 8 * N * nrhs          # the matrix once + B + C
        print(json.dumps({"op": "spmm N,N", "matrix": f"{name}:{n}", "nrhs": nrhs, "us": ms * 1e3, "speedup_vs_spmv_loop": ms1 * nrhs / ms,
                          "gbs_ideal": ideal / ms / 1e6, "frac_measured_peak": ideal / ms / 1e6 / PEAK, "gflops": 2.0 * nnz * nrhs / ms / 1e6}), flush=True)
        del Bm, Cm


if __name__ == "__main__":
    main()
