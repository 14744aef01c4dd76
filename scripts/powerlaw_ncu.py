"""Minimal launch list for an ncu capture of the heavy-tailed SpMV: 3 products of ours (virtual-row split) and 3 of the reference's
cuSPARSE path on the power-law matrix.   ncu --set full -k regex:'spmv|csr|vsplit' python scripts/powerlaw_ncu.py [log2N] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dt = sys.argv[2] if len(sys.argv) > 2 else "f64"
e = hb.gpu_engine(0)
N = 1 << log2n
p, i, v = mg.powerlaw(N=N, dtype=dt)
tp, ti, tv = (torch.from_numpy(a).to("cuda:0") for a in (p, i, v))
x = torch.from_numpy(mg.probe_x(N, dt)).to("cuda:0"); y = torch.empty_like(x)
gp, gi, gv, gx, gy = (devgen.torch_view(e, t) for t in (tp, ti, tv, x, y))
A = hb.make_sparse_matrix(e, N, gp, gi, gv)
for _ in range(3):
    A.gemv("N", 1.0, gx, 0.0, gy)
e.synchronize()
try:
    from oracle import binding
    ref = binding.reference_gpu()
    if ref is not None:
        ref.spmv_us(3 if dt == "c64" else 1, N, N, int(i.size), tp.data_ptr(), ti.data_ptr(), tv.data_ptr(), x.data_ptr(), y.data_ptr(), 1, 2)
except Exception as ex:
    print("no reference gpu path:", ex)
