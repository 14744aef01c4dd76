"""Multi-GPU host layer: one process per GPU, torch.distributed for the rendezvous, libhalab200's own NCCL communicator for
the data path (halo exchange + scalar all-reduce inside hb_dist_cg).  Used by bench.py --gpus N and the multi-rank tests."""
import ctypes as C
import json
import os

import numpy as np

from . import capi, partition
from .capi import lib, check

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_pi, _pvp = C.POINTER(C.c_int), C.POINTER(C.c_void_p)
DIST_SIGNATURES = {
    "hb_dist_unique_id": (_i, [_vp]),
    "hb_dist_create": (_i, [_vp, _i, _i, _vp, _pvp]),
    "hb_dist_destroy": (_i, [_vp]),
    "hb_dist_info": (_i, [_vp, _pi, _pi]),
    "hb_dist_transport": (_i, [_vp, _pi]),
    "hb_dist_set_plan": (_i, [_vp, _i, _i, _i, _pi, _pi, _pi, _vp]),
    "hb_dist_halo_exchange": (_i, [_vp, _i, _vp]),
    "hb_dist_allreduce_sum": (_i, [_vp, _i, _vp, _i]),
    "hb_dist_halo_exchange_nccl": (_i, [_vp, _i, _vp]),
    "hb_dist_allreduce_sum_nccl": (_i, [_vp, _i, _vp, _i]),
    "hb_dist_prepare_transport": (_i, [_vp, _i]),
    "hb_dist_finish_transport": (_i, [_vp, _pi]),
    "hb_dist_debug_info": (_i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), _pi, _pi]),
    "hb_dist_cg": (_i, [_vp, _vp, _vp, _vp, _d, _i, _pi, C.POINTER(_d)]),
    "hb_dist_spmv": (_i, [_vp, _vp, _vp, _vp]),
    "hb_dist_gmres": (_i, [_vp, _vp, _vp, _vp, _d, _i, _i, _i, _pi, C.POINTER(_d)]),
}
for _name, (_res, _args) in DIST_SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


class Communicator:
    """hb_dist handle of this rank; the NCCL unique id travels over torch.distributed (any backend)."""

    def __init__(self, engine, rank, world, group=None):
        import torch
        import torch.distributed as dist
        self.engine, self.rank, self.world = engine, rank, world
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            check(lib.hb_dist_unique_id(buf), "hb_dist_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            ident = ident.cuda()
        dist.broadcast(ident, src=0, group=group)
        raw = bytes(ident.cpu().tolist())
        self.h = C.c_void_p()
        check(lib.hb_dist_create(engine.ctx, rank, world, raw, C.byref(self.h)), "hb_dist_create")
        self._keep = None

    def set_plan(self, n_owned, n_ghost, plan):
        k = len(plan["neigh"])
        arr = lambda v: (C.c_int * max(k, 1))(*v) if k else (C.c_int * 1)()
        self._keep = plan["send_idx"]
        check(lib.hb_dist_set_plan(self.h, n_owned, n_ghost, k, arr(plan["neigh"]), arr(plan["send_count"]), arr(plan["recv_count"]),
                                   C.c_void_p(plan["send_idx"].data_ptr() if plan["send_idx"].numel() else 0)), "hb_dist_set_plan")

    def cg(self, csr, b_ptr, x_ptr, tol, max_iter):
        it, res = C.c_int(0), C.c_double(0)
        check(lib.hb_dist_cg(self.h, csr.h, b_ptr, x_ptr, float(tol), int(max_iter), C.byref(it), C.byref(res)), "hb_dist_cg")
        return it.value, res.value

    def transport(self):
        """'peer' (NVLink peer memory, no collective call per iteration) or 'nccl' — what the last cg() on this plan used"""
        t = C.c_int(0)
        check(lib.hb_dist_transport(self.h, C.byref(t)), "hb_dist_transport")
        return "peer" if t.value == 1 else "nccl"

    def debug_info(self):
        """sequence numbers of the peer protocol (equal on all ranks between solves), solves redone over NCCL after a peer
        time-out, start-of-solve agreements that found the ranks' sequence numbers different"""
        ep, vep, fb, rep = C.c_ulonglong(0), C.c_ulonglong(0), C.c_int(0), C.c_int(0)
        check(lib.hb_dist_debug_info(self.h, C.byref(ep), C.byref(vep), C.byref(fb), C.byref(rep)), "hb_dist_debug_info")
        return {"epoch": ep.value, "vepoch": vep.value, "peer_fallbacks": fb.value, "epoch_repairs": rep.value}

    def gmres(self, csr, b_ptr, x_ptr, tol, max_outer, restart, cproj=0):
        it, res = C.c_int(0), C.c_double(0)
        check(lib.hb_dist_gmres(self.h, csr.h, b_ptr, x_ptr, float(tol), int(max_outer), int(restart), int(cproj), C.byref(it), C.byref(res)), "hb_dist_gmres")
        return it.value, res.value

    def spmv(self, csr, x_ext_ptr, y_ptr):
        check(lib.hb_dist_spmv(self.h, csr.h, x_ext_ptr, y_ptr), "hb_dist_spmv")

    def __del__(self):
        try:
            if self.h:
                lib.hb_dist_destroy(self.h)
        except Exception:
            pass


def build_local_problem(engine, comm, name, n, device, group=None):
    """This rank's row slab of the n^3 stencil `name`, columns renumbered to [owned | ghosts], exchange plan installed."""
    import torch
    from . import devgen
    import hala_b200 as hb
    N = n ** (2 if name == "lap2d" else 3)
    lo, hi = partition.block_range(N, comm.world, comm.rank)
    tp, ti, tv = devgen.stencil_slab(name, n, lo, hi, device=device)
    ti_local, ghosts = partition.build_ghost_map(ti, lo, hi)
    del ti
    plan = partition.exchange_plan(ghosts, N, comm.world, comm.rank, lo, group=group, device=device)
    n_owned, n_ghost = hi - lo, int(ghosts.numel())
    comm.set_plan(n_owned, n_ghost, plan)
    gp, gi, gv = (devgen.torch_view(engine, t) for t in (tp, ti_local, tv))
    A = hb.gpu_sparse_matrix(engine, n_owned, n_owned + n_ghost, ti_local.numel(), gp, gi, gv)
    return {"A": A, "N": N, "lo": lo, "hi": hi, "n_owned": n_owned, "n_ghost": n_ghost, "nnz_local": ti_local.numel(),
            "tensors": (tp, ti_local, tv), "plan": plan}


def run_bench(args, slab, ClockSampler, measured_peak, kernel_profile):
    """bench.py --gpus N (N > 1), launched by torch.distributed.run: strong scaling of CG on the n^3 7-point Laplacian
    (or, with --workload gmres, of GMRES(50) on the convection-diffusion matrix)."""
    import torch
    import torch.distributed as dist
    import hala_b200 as hb
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    e = hb.gpu_engine(local)
    comm = Communicator(e, rank, world)
    try:
        body = _run_bench_gmres if args.workload == "gmres" else _run_bench_cg
        body(args, e, comm, rank, world, local, dev, ClockSampler, measured_peak, kernel_profile)
    except BaseException as ex:
        # what bench.py prints for a failing rank: transport and sequence numbers of the peer protocol included
        try:
            ex.hb_diag = dict(comm.debug_info(), transport=comm.transport())
        except Exception:
            pass
        raise
    dist.barrier()
    del comm
    dist.destroy_process_group()


def _parallelism(comm, world):
    if comm.transport() == "peer":
        return (f"{world} ranks, 1-D row blocks; transport=peer: halo entries and the scalar partials of every iteration are stored into the "
                "peers' memory over NVLink by the iteration kernels (no collective call); interior rows run while the halo is in flight")
    return f"{world} ranks, 1-D row blocks; transport=nccl: ghost halo (ncclSend/Recv) + scalar all-reduces per iteration"


def _run_bench_cg(args, e, comm, rank, world, local, dev, ClockSampler, measured_peak, kernel_profile):
    import torch
    import torch.distributed as dist
    from . import matgen as mg
    peak, peak_src = measured_peak()
    n = args.grid
    prob = build_local_problem(e, comm, "lap3d7", n, dev)
    N, n_owned = prob["N"], prob["n_owned"]
    nnz_total = torch.tensor([prob["nnz_local"]], dtype=torch.int64, device=dev)
    dist.all_reduce(nnz_total)
    nnz = int(nnz_total.item())
    b = torch.full((n_owned,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
    x = torch.zeros(n_owned, dtype=torch.float64, device=dev)

    def solve(iters):
        x.zero_()
        return comm.cg(prob["A"], C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, iters + 1)

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local) as clk:
        solve(max(args.warmup, 3))
        fence()
        l0 = e.launch_count()
        fence()
        clk.mark_begin()
        e.timer_start()
        it, res = solve(args.steps)
        ms = e.timer_stop()
        clk.mark_end()
        fence()
    launches = e.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    its = args.steps / ms * 1e3
    Bcg = mg.cg_iter_bytes(N, nnz, 8)

    # ---- one roofline entry per kernel of the iteration, in context: the same solve with CUDA events around every launch.  Over peer
    # memory the waits for the neighbours' halo flags and for the other ranks' partial sums sit INSIDE these kernels, so their times
    # include what the rank spent waiting; the stand-alone figure of the SpMV below does not.
    check(lib.hb_ctx_profile(e.ctx, 1), "hb_ctx_profile")
    fence()
    solve(args.steps)
    check(lib.hb_ctx_profile(e.ctx, 0), "hb_ctx_profile")
    Bk = mg.spmv_bytes(n_owned, prob["nnz_local"], 8)
    kernels = kernel_profile(e, ["spmv_pipe_kernel<double,...,DOT> (+ halo-flag wait, partial publish) on this rank's slab",
                                 "pcg_update_kernel<double> (waits for the ranks' <p,Ap>; r -= a Ap, ||r||^2, partial publish)",
                                 "pcg_direction_kernel<double> (waits for the ranks' ||r||^2; halo push of the new p; x += a p, p = r + b p)"],
                             [Bk, 3 * 8 * n_owned, 5 * 8 * n_owned], peak)
    for k in kernels:           # max over ranks of each kernel's mean time
        tk = torch.tensor([k["us_per_launch"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        k["us_per_launch"] = float(tk.item())
        k["achieved"] = k["algorithmic_bytes_per_launch"] / k["us_per_launch"] / 1e3
        k["frac"] = k["achieved"] / peak
        k["over"] = "max over ranks of the per-rank mean; bytes = rank 0's slab"

    # dominant kernel alone on this rank's slab: halo-free SpMV+dot launches, CUDA events, max over ranks
    p_ext = torch.rand(n_owned + prob["n_ghost"], dtype=torch.float64, device=dev)
    q = torch.empty(n_owned, dtype=torch.float64, device=dev)
    slot = torch.zeros(4, dtype=torch.float64, device=dev)
    args_k = (e.ctx, prob["A"].h, C.c_void_p(p_ext.data_ptr()), C.c_void_p(q.data_ptr()), C.c_void_p(slot.data_ptr()))
    for _ in range(3):
        check(lib.hb_spmv_dot(*args_k))
    e.timer_start()
    for _ in range(30):
        check(lib.hb_spmv_dot(*args_k))
    kms = torch.tensor([e.timer_stop() / 30], dtype=torch.float64, device=dev)
    dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    kms = float(kms.item())
    halo = torch.tensor([prob["n_ghost"]], dtype=torch.int64, device=dev)
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    del p_ext, q

    # ---- e2e: every rank's slab (CSR with local column numbers + b) comes from pinned host memory, x goes back; all timed
    e2e = None
    if not args.no_e2e:
        tp, ti, tv = prob["tensors"]
        hp, hi, hv = tp.cpu().pin_memory(), ti.cpu().pin_memory(), tv.cpu().pin_memory()
        hb_, hx = b.cpu().pin_memory(), torch.empty(n_owned, dtype=torch.float64).pin_memory()
        dp, di, dv, db, dx = torch.empty_like(tp), torch.empty_like(ti), torch.empty_like(tv), torch.empty_like(b), torch.empty_like(x)
        h2d = sum(t.numel() * t.element_size() for t in (hp, hi, hv, hb_))
        d2h = hx.numel() * hx.element_size()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stream = torch.cuda.default_stream()

        def e2e_solve(iters):
            ev[0].record(stream)
            for d_, h_ in ((dp, hp), (di, hi), (dv, hv), (db, hb_)):
                check(lib.hb_memcpy_async(e.ctx, C.c_void_p(d_.data_ptr()), C.c_void_p(h_.data_ptr()), h_.numel() * h_.element_size(), 0))
            check(lib.hb_memset_zero(e.ctx, C.c_void_p(dx.data_ptr()), n_owned * 8))
            ev[1].record(stream)
            Ah = C.c_void_p()
            check(lib.hb_csr_create(e.ctx, 1, n_owned, n_owned + prob["n_ghost"], prob["nnz_local"], C.c_void_p(dp.data_ptr()), C.c_void_p(di.data_ptr()),
                                    C.c_void_p(dv.data_ptr()), C.byref(Ah)))
            it_, rs_ = C.c_int(0), C.c_double(0)
            check(lib.hb_dist_cg(comm.h, Ah, C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), 0.0, iters + 1, C.byref(it_), C.byref(rs_)), "hb_dist_cg")
            ev[2].record(stream)
            check(lib.hb_memcpy(e.ctx, C.c_void_p(hx.data_ptr()), C.c_void_p(dx.data_ptr()), n_owned * 8, 1))
            ev[3].record(stream)
            lib.hb_csr_destroy(Ah)
            return it_.value - 1

        e2e_solve(3)
        fence()
        e.timer_start()
        d_it = e2e_solve(args.steps)
        torch.cuda.synchronize()
        parts = torch.tensor([e.timer_stop(), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], dtype=torch.float64, device=dev)
        dist.all_reduce(parts, op=dist.ReduceOp.MAX)
        vol = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(vol)
        ems = float(parts[0].item())
        e2e = {"value": d_it / ems * 1e3, "unit": "iterations/s", "h2d_bytes_per_step": float(vol[0].item()) / args.steps,
               "d2h_bytes_per_step": float(vol[1].item()) / args.steps, "ms_total": ems, "h2d_ms": float(parts[1].item()),
               "solve_ms": float(parts[2].item()), "d2h_ms": float(parts[3].item()),
               "what": f"per rank: local CSR slab + b H2D from pinned host memory ({float(vol[0].item()) / 1e9:.2f} GB over all ranks), hb_csr_create, "
                       f"{args.steps} CG iterations (hb_dist_cg), x D2H; all inside the timed region, max over ranks. The upload is paid once per "
                       "solve: h2d_ms / solve_ms / d2h_ms split the figure"}
    if rank == 0:
        roof = dict(kernels[0]) if kernels else {"kernel": "spmv_pipe_kernel<double,...,DOT> on one rank's slab", "bound": "hbm", "us_per_launch": kms * 1e3,
                                                 "algorithmic_bytes_per_launch": Bk, "achieved": Bk / kms / 1e6, "peak": peak, "unit": "GB/s",
                                                 "frac": Bk / kms / 1e6 / peak, "traffic": None}
        roof.update({"peak_source": peak_src, "share_of_step": roof["us_per_launch"] / 1e3 / (ms / args.steps), "us_per_launch_alone": kms * 1e3,
                     "frac_alone": Bk / kms / 1e6 / peak,
                     "how": "CUDA events on the launching stream around every launch inside the solver loop (a second run of the same iterations), max over "
                            "ranks; *_alone = 30 back-to-back halo-free launches per rank, max over ranks"})
        line = {"metric": "cg_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"lap3d7-{n} fp64 unpreconditioned CG, b=1/sqrt(N), x0=0 (BASELINE configs[2])", "rows": N, "nnz": nnz,
                           "parallelism": _parallelism(comm, world),
                           "l2": "inputs larger than L2; no flush", "step": "one CG iteration", "max_ghosts_per_rank": int(halo.item())},
                "gbs": Bcg * its / 1e9, "frac_of_measured_peak": Bcg * its / 1e9 / (peak * world), "algorithmic_bytes_per_step": Bcg, "final_residual": res,
                "roofline": roof, "roofline_kernels": kernels, "cpu_baseline": None,
                "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(),
                "transport_diag": dict(comm.debug_info(), transport=comm.transport())}
        print(json.dumps(line), flush=True)


def _run_bench_gmres(args, e, comm, rank, world, local, dev, ClockSampler, measured_peak, kernel_profile):
    """--workload gmres: BASELINE configs[3], GMRES(50) on the convection-diffusion matrix n^3 (n = 256), row-partitioned; a step is one
    inner iteration; fixed budget of ceil(steps / 50) restart cycles with tolerance 0."""
    import torch
    import torch.distributed as dist
    from . import matgen as mg
    peak, peak_src = measured_peak()
    n, restart = args.grid, 50
    prob = build_local_problem(e, comm, "convdiff7", n, dev)
    N, n_owned = prob["N"], prob["n_owned"]
    nnz_total = torch.tensor([prob["nnz_local"]], dtype=torch.int64, device=dev)
    dist.all_reduce(nnz_total)
    nnz = int(nnz_total.item())
    b = torch.full((n_owned,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
    x = torch.zeros(n_owned, dtype=torch.float64, device=dev)
    cycles = max(1, (args.steps + restart - 1) // restart)

    def solve(cyc):
        x.zero_()
        return comm.gmres(prob["A"], C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, cyc, restart)

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    with ClockSampler(local) as clk:
        solve(1)
        fence()
        l0 = e.launch_count()
        fence()
        clk.mark_begin()
        e.timer_start()
        it, res = solve(cycles)
        ms = e.timer_stop()
        clk.mark_end()
        fence()
    launches = e.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    its = it / ms * 1e3
    Bit = mg.gmres_iter_bytes(N, nnz, 8, restart)
    if rank == 0:
        line = {"metric": "gmres_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": world, "steps": it, "warmup": restart + 1,
                "ms_per_step": ms / it, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"convdiff7-{n} fp64 GMRES({restart}), b=1/sqrt(N), x0=0, {cycles} restart cycle(s), tolerance 0 (BASELINE configs[3])",
                           "rows": N, "nnz": nnz, "parallelism": _parallelism(comm, world), "l2": "inputs larger than L2; no flush",
                           "step": "one inner iteration: halo, SpMV, multi-dot, all-reduce, multi-axpy + norm, all-reduce, normalise-and-append"},
                "gbs": Bit * its / 1e9, "frac_of_measured_peak": Bit * its / 1e9 / (peak * world), "algorithmic_bytes_per_step": Bit, "estimated_residual": res,
                "roofline": {"bound": "hbm", "kernel": "gs_pipe_kernel<double> + spmv_pipe_kernel, whole inner iteration, all ranks", "achieved": Bit * its / 1e9 / world,
                             "peak": peak, "unit": "GB/s", "frac": Bit * its / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src,
                             "how": "algorithmic bytes of an inner iteration at the mean basis size / measured time per iteration, per rank"},
                "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clk.summary(),
                "transport_diag": dict(comm.debug_info(), transport=comm.transport())}
        print(json.dumps(line), flush=True)
