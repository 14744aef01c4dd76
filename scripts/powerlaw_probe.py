"""configs[4b] (power-law row lengths) SpMV on one B200: the virtual-row split (default), the same with contiguous pieces, the general
kernel (HB_VSPLIT=0), and the reference's cuSPARSE path; parity of every variant against left-to-right numpy sums on sampled rows
(the longest rows included).  One JSON line per measurement.   usage: powerlaw_probe.py [log2N=22] [dtypes=f64,c64]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg

LOG2N = int(sys.argv[1]) if len(sys.argv) > 1 else 22
DTYPES = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f64", "c64"]
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6548.2
e = hb.gpu_engine(0)
dev = "cuda:0"
def emit(**kw): print(json.dumps(kw), flush=True)

try:
    from oracle import binding          # bench/test infrastructure: the reference's own GPU path
    refgpu = binding.reference_gpu()
except Exception:
    refgpu = None

for dt in DTYPES:
    N = 1 << LOG2N
    t0 = time.time()
    p, i, v = mg.powerlaw(N=N, dtype=dt)
    gen = time.time() - t0
    tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
    xh = mg.probe_x(N, dt)
    x = torch.from_numpy(xh).to(dev); y = torch.empty_like(x)
    gp, gi, gv, gx, gy = (devgen.torch_view(e, t) for t in (tp, ti, tv, x, y))
    B = mg.spmv_bytes(N, i.size, v.dtype.itemsize)
    lens = np.diff(p)
    rows = np.concatenate([np.argsort(lens)[-8:], np.arange(0, N, max(N // 400, 1))])
    ref = np.array([np.sum(v[p[r]:p[r + 1]] * xh[i[p[r]:p[r + 1]]]) for r in rows])
    scale = np.array([np.sum(np.abs(v[p[r]:p[r + 1]]) * np.abs(xh[i[p[r]:p[r + 1]]])) for r in rows])
    for tag, env in (("vsplit round-robin", {}), ("vsplit round-robin, 2 lanes per row", {"HB_VS_TPR": "2"}), ("vsplit contiguous", {"HB_PIPE_MAP": "c"}),
                     ("general kernel (HB_VSPLIT=0)", {"HB_VSPLIT": "0"}),
                     ("vsplit, warp rows from 16", {"HB_VS_WARPROW": "16"}), ("vsplit, warp rows from 24", {"HB_VS_WARPROW": "24"}),
                     ("vsplit, warp rows from 48", {"HB_VS_WARPROW": "48"}), ("vsplit, warp rows from 96", {"HB_VS_WARPROW": "96"}),
                     ("PROBE no row sums (wrong results)", {"HB_VS_PROBE": "1"}), ("PROBE no gathers (wrong results)", {"HB_VS_PROBE": "2"}),
                     ("PROBE neither (wrong results)", {"HB_VS_PROBE": "3"})):
        for k in ("HB_PIPE_MAP", "HB_VSPLIT", "HB_VS_TPR", "HB_VS_PROBE", "HB_VS_WARPROW"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = time.perf_counter()
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        e.synchronize()
        create_ms = (time.perf_counter() - t0) * 1e3
        for _ in range(3): A.gemv("N", 1.0, gx, 0.0, gy)
        e.timer_start()
        for _ in range(30): A.gemv("N", 1.0, gx, 0.0, gy)
        us = e.timer_stop() / 30 * 1e3
        yh = y.cpu().numpy()
        err = float(np.max(np.abs(yh[rows] - ref) / scale))
        # alpha / beta path and the fused dot through the same kernels
        y.copy_(x)
        A.gemv("N", 2.0, gx, -0.5, gy)
        err2 = float(np.max(np.abs(y.cpu().numpy()[rows] - (2.0 * ref - 0.5 * xh[rows])) / (2 * scale + np.abs(xh[rows]))))
        emit(op="spmv", config=f"powerlaw 2^{LOG2N} {dt}", variant=tag, nnz=int(i.size), max_row=int(lens.max()), us=us, gbs=B / us / 1e3,
             frac_measured_peak=B / us / 1e3 / PEAK, create_ms=create_ms, worst_scaled_err=err, worst_scaled_err_alpha_beta=err2)
        del A
    for k in ("HB_PIPE_MAP", "HB_VSPLIT", "HB_VS_TPR", "HB_VS_PROBE", "HB_VS_WARPROW"):
        os.environ.pop(k, None)
    if refgpu is not None:
        code = 3 if dt == "c64" else 1
        us = refgpu.spmv_us(code, N, N, int(i.size), tp.data_ptr(), ti.data_ptr(), tv.data_ptr(), x.data_ptr(), y.data_ptr(), 5, 30)
        emit(op="spmv", config=f"powerlaw 2^{LOG2N} {dt}", variant="reference gpu_engine (cusparseSpMV ALG_DEFAULT)", us=us, gbs=B / us / 1e3, frac_measured_peak=B / us / 1e3 / PEAK)
    del tp, ti, tv, x, y
    torch.cuda.empty_cache()
