"""Shared-memory carve-out against L1 for the x gather: the streaming SpMV kernels with 1..n resident CTAs per SM (the carve-out is
sized for exactly that many; the rest of the 256 KB array is L1).  HB_PIPE_CTAS / HB_VS_CTAS are read at hb_csr_create.
One JSON line per measurement.   usage: carveout_probe.py [cases=pl,c2,c3,c5a]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg

cases = sys.argv[1].split(",") if len(sys.argv) > 1 else ["pl", "c2", "c3", "c5a"]
e = hb.gpu_engine(0)
dev = "cuda:0"
PEAK = 6548.2
def emit(**kw): print(json.dumps(kw), flush=True)

def sweep(tag, tp, ti, tv, dt, var, settings, reps=50):
    N, nnz = tp.numel() - 1, ti.numel()
    x = torch.from_numpy(mg.probe_x(N, dt)).to(dev); y = torch.empty_like(x)
    gp, gi, gv, gx, gy = (devgen.torch_view(e, t) for t in (tp, ti, tv, x, y))
    B = mg.spmv_bytes(N, nnz, 16 if dt == "c64" else 8)
    ref = None
    for n in settings:
        os.environ[var] = str(n)
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        for _ in range(5): A.gemv("N", 1.0, gx, 0.0, gy)
        e.timer_start()
        for _ in range(reps): A.gemv("N", 1.0, gx, 0.0, gy)
        us = e.timer_stop() / reps * 1e3
        yh = y.clone()
        if ref is None: ref = yh
        same = bool(torch.equal(ref, yh))
        emit(case=tag, knob=var, ctas_per_sm=n, us=us, gbs=B / us / 1e3, frac_measured_peak=B / us / 1e3 / PEAK, same_bits_as_first=same)
        del A
    os.environ.pop(var, None)

if "pl" in cases:
    p, i, v = mg.powerlaw(N=1 << 22, dtype="f64")
    tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
    sweep("powerlaw 2^22 f64 (VS kernel)", tp, ti, tv, "f64", "HB_VS_CTAS", [int(q) for q in os.environ.get("VS_SWEEP", "0,5,4,3,2,1").split(",")], reps=20)
    del tp, ti, tv
if "c2" in cases:
    tp, ti, tv = devgen.stencil_slab("lap3d27", 128, 0, 128 ** 3, device=dev)
    sweep("lap3d27-128 f64", tp, ti, tv, "f64", "HB_PIPE_CTAS", [0, 3, 2, 1], reps=100)
    del tp, ti, tv
if "c3" in cases:
    tp, ti, tv = devgen.stencil_slab("lap3d7", 256, 0, 256 ** 3, device=dev)
    sweep("lap3d7-256 f64", tp, ti, tv, "f64", "HB_PIPE_CTAS", [0, 6, 5, 4, 3, 2], reps=50)
    del tp, ti, tv
if "c5a" in cases:
    tp, ti, tv = devgen.stencil_slab("helmholtz7", 192, 0, 192 ** 3, dtype="c64", device=dev)
    sweep("helmholtz7-192 c64", tp, ti, tv, "c64", "HB_PIPE_CTAS", [0, 6, 5, 4, 3, 2], reps=50)
    del tp, ti, tv
