"""27-point 128^3 SpMV: 128- vs 256-thread CTAs of the streaming kernel (HB_PIPE_CFG) x resident CTAs the carve-out is sized for (HB_PIPE_CTAS)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hala_b200 as hb
from hala_b200 import devgen, matgen as mg
e = hb.gpu_engine(0); dev = "cuda:0"
tp, ti, tv = devgen.stencil_slab("lap3d27", 128, 0, 128 ** 3, device=dev)
N, nnz = tp.numel() - 1, ti.numel()
x = torch.from_numpy(mg.probe_x(N)).to(dev); y = torch.empty_like(x)
gp, gi, gv, gx, gy = (devgen.torch_view(e, t) for t in (tp, ti, tv, x, y))
B = mg.spmv_bytes(N, nnz, 8)
for cfg in ("1", "0"):
    for ctas in ("0", "6", "5", "4", "3"):
        os.environ["HB_PIPE_CFG"] = cfg; os.environ["HB_PIPE_CTAS"] = ctas
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        for _ in range(5): A.gemv("N", 1.0, gx, 0.0, gy)
        e.timer_start()
        for _ in range(100): A.gemv("N", 1.0, gx, 0.0, gy)
        us = e.timer_stop() / 100 * 1e3
        print(json.dumps({"case": "lap3d27-128", "HB_PIPE_CFG": cfg, "HB_PIPE_CTAS": ctas, "us": us, "gbs": B / us / 1e3}), flush=True)
        del A
