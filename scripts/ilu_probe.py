"""SURVEY.md §8 row f1 on one B200: ILU(0) factorisation and the two triangular solves of its application on the 7-point
Laplacian n^3 (the preconditioner of the reference's headline entry points solve_cg_ilu / solve_gmres_ilu), device-timed,
with the reference's own CPU path (oracle/_ref: factorize_ilu + cpu_ilu::apply) timed beside it on a bounded size.
One JSON line per measurement.  Not the contract bench."""
import json
import os
import sys
import time

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


def main():
    e = hb.gpu_engine(0)
    dev = "cuda:0"
    for n in [int(t) for t in os.environ.get("ILU_N", "128,256").split(",")]:
        N = n ** 3
        tp, ti, tv = devgen.stencil_slab("lap3d7", n, 0, N, device=dev)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        nnz = ti.numel()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ilu = hb.make_ilu(e, gp, gi, gv)
        e.synchronize()
        t_factor = time.perf_counter() - t0
        b = e.load(mg.probe_x(N))
        r = e.new_vector(np.float64, N)
        t0 = time.perf_counter()
        ilu.apply(b, r)                                   # first application: includes both level analyses
        e.synchronize()
        t_first = time.perf_counter() - t0
        reps = 20
        e.timer_start()
        for _ in range(reps):
            ilu.apply(b, r)
        ms = e.timer_stop() / reps
        # bytes of one application: both triangles of the factor array are walked over the FULL rows (12 B per stored entry, twice),
        # row pointers twice, b / tmp / x read and written, flags read per dependency and written per row
        nlow = (nnz - N) // 2
        bytes_apply = 2 * (12 * nnz + 4 * (N + 1)) + 2 * (8 * N * 2) + 2 * 8 * nlow + 2 * (4 * nlow + 4 * N) + 2 * 4 * N
        emit(op="ilu0_factor+analysis", n=n, rows=N, nnz=nnz, seconds=t_factor)
        emit(op="ilu_apply_first(with analyses)", n=n, seconds=t_first, levels_L=ilu.lower.levels(), levels_U=ilu.upper.levels())
        emit(op="ilu_apply", n=n, ms=ms, gbs=bytes_apply / ms / 1e6, frac_measured_peak=bytes_apply / ms / 1e6 / PEAK,
             rows_per_level=N / max(ilu.lower.levels(), 1))
        # residual check of the application: (L U) r == b
        fac = None
        del ilu, r, b
        torch.cuda.empty_cache()
    # CPU reference on a bounded size (its factorisation is O(N^2): hala_sparse_utils.hpp:230-253)
    try:
        from oracle import binding
        ref = binding.reference() or binding.oracle()
        nc = int(os.environ.get("ILU_CPU_N", "20"))
        p, i, v = mg.lap3d7(nc)
        x = mg.probe_x(nc ** 3)
        t0 = time.perf_counter()
        ref.ilu(p, i, v, x)
        emit(op="cpu_reference ilu factor+apply", n=nc, rows=nc ** 3, seconds=time.perf_counter() - t0, cores=1)
    except Exception as ex:      # noqa: BLE001
        emit(op="cpu_reference", error=str(ex))


if __name__ == "__main__":
    main()
