"""BASELINE configs[3] on N GPUs: GMRES(50) on the nonsymmetric convection-diffusion matrix 256^3, row-partitioned
(launched by torch.distributed.run, or plainly for N = 1).  Prints one JSON line: inner iterations per second over a fixed
number of restart cycles, device-timed, max over ranks.  Not the contract bench (that is bench.py)."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import hala_b200 as hb
    from hala_b200 import dist as hbdist, devgen
    from hala_b200.capi import lib, check
    n = int(os.environ.get("C4_N", "256"))
    cycles = int(os.environ.get("C4_OUTER", "6"))
    restart = int(os.environ.get("C4_RESTART", "50"))
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    e = hb.gpu_engine(local)
    N = n ** 3
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        comm = hbdist.Communicator(e, rank, world)
        prob = hbdist.build_local_problem(e, comm, "convdiff7", n, dev)
        A, n_owned = prob["A"], prob["n_owned"]
    else:
        tp, ti, tv = devgen.stencil_slab("convdiff7", n, 0, N, device=dev)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        A, n_owned = hb.make_sparse_matrix(e, N, gp, gi, gv), N
    b = torch.full((n_owned,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
    x = torch.zeros(n_owned, dtype=torch.float64, device=dev)

    def solve(outer):
        x.zero_()
        it, res = C.c_int(0), C.c_double(0)
        if world > 1:
            return comm.gmres(A, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 1e-30, outer, restart)
        check(lib.hb_gmres(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 1e-30, outer, restart, 0, C.byref(it), C.byref(res)))
        return it.value, res.value

    solve(1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e.timer_start()
    it, res = solve(cycles)
    ms = e.timer_stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"op": f"gmres{restart}", "config": f"c4 convdiff7-{n} fp64", "n_gpus": world, "matvecs": it, "cycles": cycles,
                          "ms_total": ms, "its_per_s": it / ms * 1e3, "est_res": res}), flush=True)
    if world > 1:
        dist.barrier()
        del comm
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
