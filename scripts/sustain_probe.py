"""Sustained-load probe: long back-to-back loops of SpMV / SpMV+dot / CG with SM clock sampling (nvidia-smi)."""
import ctypes as C, json, os, subprocess, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hala_b200 as hb
from hala_b200 import matgen as mg
from hala_b200.capi import lib, check

class Clocks:
    def __init__(self): self.samples, self.stop = [], False
    def run(self):
        while not self.stop:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                                  "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
            self.samples.append((time.time(), out))
    def __enter__(self):
        self.t = threading.Thread(target=self.run); self.t.start(); return self
    def __exit__(self, *a):
        self.stop = True; self.t.join()

name, size = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
e = hb.gpu_engine(0)
p, i, v = mg.GENERATORS[name](size)
N, nnz = p.size - 1, i.size
gp, gi, gv = e.load(p), e.load(i), e.load(v)
A = hb.make_sparse_matrix(e, N, gp, gi, gv)
gx, gy, gs = e.load(mg.probe_x(N)), e.new_vector(np.float64, N), e.new_vector(np.float64, 4)
B = mg.spmv_bytes(N, nnz, 8)
def loop(fn, n):
    for _ in range(5): fn()
    e.timer_start()
    for _ in range(n): fn()
    return e.timer_stop() / n
for label, fn in (("spmv", lambda: A.gemv("N", 1.0, gx, 0.0, gy)),
                  ("spmv_dot", lambda: check(lib.hb_spmv_dot(e.ctx, A.h, gx.ptr, gy.ptr, gs.ptr)))):
    for n in (20, reps, 4 * reps):
        with Clocks() as c:
            ms = loop(fn, n)
        print(json.dumps({"op": label, "reps": n, "ms": ms, "gbs": B / ms / 1e6, "clocks": [s[1] for s in c.samples][-3:]}), flush=True)
gb = e.load(mg.rhs(N))
for iters in (50, reps, 4 * reps):
    gxx = e.new_vector(np.float64)
    with Clocks() as c:
        e.timer_start()
        it, res = hb.solve_cg(e, 0.0, iters + 1, gp, gi, gv, gb, gxx, matrix=A)
        ms = e.timer_stop()
    print(json.dumps({"op": "cg", "iters": it - 1, "ms_per_iter": ms / (it - 1), "its_per_s": (it - 1) / ms * 1e3,
                      "gbs": mg.cg_iter_bytes(N, nnz, 8) * (it - 1) / ms / 1e6, "clocks": [s[1] for s in c.samples][-3:]}), flush=True)
