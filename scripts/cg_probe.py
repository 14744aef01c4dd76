"""Runs a short fixed-iteration CG on one config (for ncu launch lists)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hala_b200 as hb
from hala_b200 import matgen as mg
name, size, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
e = hb.gpu_engine(0)
p, i, v = mg.GENERATORS[name](size)
N = p.size - 1
gp, gi, gv = e.load(p), e.load(i), e.load(v)
A = hb.make_sparse_matrix(e, N, gp, gi, gv)
gb, gx = e.load(mg.rhs(N)), e.new_vector(np.float64)
it, res = hb.solve_cg(e, 0.0, iters + 1, gp, gi, gv, gb, gx, matrix=A)
print("iters", it, "res", res)
