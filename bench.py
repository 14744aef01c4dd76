#!/usr/bin/env python
"""bench.py — the contract benchmark of hala_b200.

    python bench.py --gpus N --steps K --warmup W            our arm (libhalab200 on B200)
    python bench.py --impl reference --gpus N ...            the reference's own CPU path (oracle/_ref) on the host cores

Workload (all N): BASELINE configs[2] — fp64 3-D 7-point Laplacian 512^3 (134,217,728 rows, 937,951,232 non-zeros),
unpreconditioned CG, b = 1/sqrt(N), x0 = 0, row-partitioned over the N ranks (strong scaling).  A "step" is one CG
iteration (SpMV fused with <p,Ap>, fused x/r update + ||r||^2, direction update); `value` = iterations per second with
everything resident in HBM.  At N = 1 the same line also carries BASELINE configs[1] — CSR SpMV GB/s on the fp64 27-point
Laplacian 128^3 — under "spmv", because BASELINE.json's metric names both.
`e2e` = the same K iterations through the host-buffer entry point (mixed-engine semantics of the reference,
hex/solvers/hala_solvers_cg.hpp:250-264): CSR + b copied H2D from pinned memory, solve, x copied back, all timed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 512          # configs[2]
SPMV_GRID = 128     # configs[1]


def ncu_traffic(tag):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/traffic_r*.json, written here by
    scripts/ncu_summary.py from the .ncu-rep); None when the workload is not the profiled one."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")), reverse=True):
        try:
            with open(f) as fh:
                d = json.load(fh)
            if tag in d:
                return d[tag]["dram_bytes_per_launch"]
        except Exception:
            pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        """nvidia-smi needs a few hundred ms to start: enter the context before the warm-up, mark the timed window here"""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        rows = [r for t, r in self.rows if (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1 + 0.1)]
        if not rows and self.rows:      # window shorter than one sampling period: take the sample closest to it
            mid = 0.5 * ((self.t0 or self.rows[0][0]) + (self.t1 or self.rows[-1][0]))
            rows = [min(self.rows, key=lambda tr: abs(tr[0] - mid))[1]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v == "Active":
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def slab(N, P, r):
    """1-D contiguous row blocks: rank r owns [floor(rN/P), floor((r+1)N/P))  (SURVEY.md §8e)."""
    return (r * N) // P, ((r + 1) * N) // P


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation (hala::solve_cg on cpu_engine, unmodified headers, built into
    oracle/_ref/libhala_ref.so) on a bounded sample of the same workload: a 512 x 512 x S slab of the 7-point Laplacian,
    K iterations, scaled by S/512 to the full problem (the work per iteration is linear in the rows)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding
    from hala_b200 import matgen as mg
    lib, kind = binding.reference(), "reference"
    if lib is None:
        lib, kind = binding.oracle(), "port"
    S = args.ref_planes
    n = GRID
    rows = n * n * S
    p, i, v = mg._stencil((S, n, n), mg._offsets(3, False), [-1.0] * 7, 6.0, np.float64)
    b = np.full(rows, 1.0 / np.sqrt(n ** 3))
    lib.cg(p, i, v, b, 0.0, max_iter=args.warmup + 1)
    t0 = time.perf_counter()
    _, it = lib.cg(p, i, v, b, 0.0, max_iter=args.steps + 1)
    dt = time.perf_counter() - t0
    its = (it - 1) / dt * (S / n)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "cg_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lap3d7-{n} fp64 CG (BASELINE configs[2]); CPU sample = {n}x{n}x{S} slab, scaled by {S}/{n}"},
            "cpu_baseline": {"value": its, "unit": "iterations/s", "cores": cores, "kind": kind,
                             "sample": f"{n}x{n}x{S} slab ({rows} rows, {i.size} nnz), {it - 1} CG iterations in {dt:.2f} s; SpMV is serial "
                                       f"by construction (sparse/hala_sparse_utils.hpp:103-118), BLAS-1 on OpenBLAS with {cores} threads"},
            "e2e": {"value": its, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import hala_b200 as hb
    from hala_b200 import devgen, matgen as mg
    from hala_b200.capi import lib, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from hala_b200 import dist as hbdist
        return hbdist.run_bench(args, slab, ClockSampler, measured_peak)
    torch.cuda.set_device(local)
    e = hb.gpu_engine(local)
    peak, peak_src = measured_peak()
    n = args.grid
    N = n ** 3
    dev = f"cuda:{local}"

    # ---- configs[1]: SpMV on the 27-point Laplacian 128^3 (inputs 711 MB > L2, no flush needed)
    tp, ti, tv = devgen.stencil_slab("lap3d27", SPMV_GRID, 0, SPMV_GRID ** 3, device=dev)
    N1, nnz1 = SPMV_GRID ** 3, ti.numel()
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    A1 = hb.make_sparse_matrix(e, N1, gp, gi, gv)
    x1 = torch.from_numpy(mg.probe_x(N1)).to(dev)
    y1 = torch.empty_like(x1)
    gx1, gy1 = devgen.torch_view(e, x1), devgen.torch_view(e, y1)
    for _ in range(5):
        A1.gemv("N", 1.0, gx1, 0.0, gy1)
    e.timer_start()
    reps = args.spmv_reps
    for _ in range(reps):
        A1.gemv("N", 1.0, gx1, 0.0, gy1)
    ms1 = e.timer_stop() / reps
    B1 = mg.spmv_bytes(N1, nnz1, 8)
    spmv = {"workload": "lap3d27-128 fp64 CSR SpMV (BASELINE configs[1])", "rows": N1, "nnz": nnz1, "us": ms1 * 1e3,
            "gbs": B1 / ms1 / 1e6, "gflops": 2 * nnz1 / ms1 / 1e6, "frac_of_measured_peak": B1 / ms1 / 1e6 / peak,
            "frac_of_8tbs_nominal": B1 / ms1 / 1e6 / 8000.0, "algorithmic_bytes": B1}
    del A1, gp, gi, gv, tp, ti, tv, x1, y1
    torch.cuda.empty_cache()

    # ---- configs[2]: CG on the 7-point Laplacian n^3, device resident
    tp, ti, tv = devgen.stencil_slab("lap3d7", n, 0, N, device=dev)
    nnz = ti.numel()
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    A = hb.make_sparse_matrix(e, N, gp, gi, gv)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
    x = torch.zeros(N, dtype=torch.float64, device=dev)
    gb, gx = devgen.torch_view(e, b), devgen.torch_view(e, x)

    def solve(iters):
        x.zero_()
        it, res = C.c_int(0), C.c_double(0)
        check(lib.hb_cg(e.ctx, A.h, gb.ptr, gx.ptr, 0.0, iters + 1, C.byref(it), C.byref(res)), "hb_cg")
        return it.value - 1, res.value

    with ClockSampler(local) as clk:
        solve(max(args.warmup, 3))
        torch.cuda.synchronize()
        l0 = e.launch_count()
        clk.mark_begin()
        e.timer_start()
        done, res = solve(args.steps)
        ms = e.timer_stop()
        clk.mark_end()
    launches = e.launch_count() - l0
    assert done == args.steps, (done, args.steps)
    its = args.steps / ms * 1e3
    Bcg = mg.cg_iter_bytes(N, nnz, 8)

    # ---- roofline of the dominant kernel (SpMV fused with the dot), timed alone on the same matrix right after the run
    p_like = torch.from_numpy(mg.probe_x(N)).to(dev) if N <= (1 << 27) else torch.rand(N, dtype=torch.float64, device=dev)
    q = torch.empty(N, dtype=torch.float64, device=dev)
    slot = torch.zeros(4, dtype=torch.float64, device=dev)
    for _ in range(3):
        check(lib.hb_spmv_dot(e.ctx, A.h, C.c_void_p(p_like.data_ptr()), C.c_void_p(q.data_ptr()), C.c_void_p(slot.data_ptr())))
    e.timer_start()
    kreps = 30
    for _ in range(kreps):
        check(lib.hb_spmv_dot(e.ctx, A.h, C.c_void_p(p_like.data_ptr()), C.c_void_p(q.data_ptr()), C.c_void_p(slot.data_ptr())))
    kms = e.timer_stop() / kreps
    Bk = mg.spmv_bytes(N, nnz, 8)
    roof = {"bound": "hbm", "kernel": "spmv_pipe_kernel<double,...,DOT> (CSR SpMV fused with <p,Ap>)", "achieved": Bk / kms / 1e6, "peak": peak,
            "unit": "GB/s", "frac": Bk / kms / 1e6 / peak, "traffic": ncu_traffic(f"spmv_pipe_dot_lap3d7_{n}"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": Bk,
            "us_per_launch": kms * 1e3, "share_of_step": kms / (ms / args.steps),
            "how": "CUDA events around 30 back-to-back launches on the bench matrix, same process, right after the timed region"}
    del p_like, q

    # ---- e2e: host buffers in, host buffer out (pinned), every copy inside the timed region
    e2e = None
    if not args.no_e2e:
        hp, hi, hv = tp.cpu().pin_memory(), ti.cpu().pin_memory(), tv.cpu().pin_memory()
        hb_, hx = b.cpu().pin_memory(), torch.empty(N, dtype=torch.float64).pin_memory()
        dp, di, dv = torch.empty_like(tp), torch.empty_like(ti), torch.empty_like(tv)
        db, dx = torch.empty_like(b), torch.empty_like(x)
        h2d = sum(t.numel() * t.element_size() for t in (hp, hi, hv, hb_))
        d2h = hx.numel() * hx.element_size()

        def e2e_solve(iters):
            for d, h in ((dp, hp), (di, hi), (dv, hv), (db, hb_)):
                check(lib.hb_memcpy_async(e.ctx, C.c_void_p(d.data_ptr()), C.c_void_p(h.data_ptr()), h.numel() * h.element_size(), 0))
            check(lib.hb_memset_zero(e.ctx, C.c_void_p(dx.data_ptr()), N * 8))
            Ah = C.c_void_p()
            check(lib.hb_csr_create(e.ctx, 1, N, N, nnz, C.c_void_p(dp.data_ptr()), C.c_void_p(di.data_ptr()), C.c_void_p(dv.data_ptr()), C.byref(Ah)))
            it, rs = C.c_int(0), C.c_double(0)
            check(lib.hb_cg(e.ctx, Ah, C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), 0.0, iters + 1, C.byref(it), C.byref(rs)))
            check(lib.hb_memcpy(e.ctx, C.c_void_p(hx.data_ptr()), C.c_void_p(dx.data_ptr()), N * 8, 1))
            lib.hb_csr_destroy(Ah)
            return it.value - 1

        e2e_solve(3)
        t0 = time.perf_counter()
        e.timer_start()
        d_it = e2e_solve(args.steps)
        ems = e.timer_stop()
        wall = time.perf_counter() - t0
        e2e = {"value": d_it / ems * 1e3, "unit": "iterations/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "ms_total": ems, "wall_ms": wall * 1e3,
               "what": f"CSR + b H2D from pinned host memory ({h2d / 1e9:.2f} GB), hb_csr_create, {args.steps} CG iterations, x D2H; all inside the timed region"}

    # ---- CPU baseline: the reference's own cpu_engine CG (oracle/_ref) on a bounded slab of the same matrix
    cpu = None
    if not args.no_cpu:
        from oracle import binding
        ref, kind = binding.reference(), "reference"
        if ref is None:
            ref, kind = binding.oracle(), "port"
        S = args.ref_planes
        cp, ci, cv = mg._stencil((S, n, n), mg._offsets(3, False), [-1.0] * 7, 6.0, np.float64)
        cb = np.full(n * n * S, 1.0 / np.sqrt(N))
        cit = max(4, min(args.steps, 40))
        ref.cg(cp, ci, cv, cb, 0.0, max_iter=3)
        t0 = time.perf_counter()
        _, it = ref.cg(cp, ci, cv, cb, 0.0, max_iter=cit + 1)
        dt = time.perf_counter() - t0
        cpu = {"value": (it - 1) / dt * (S / n), "unit": "iterations/s", "cores": os.cpu_count(), "kind": kind,
               "sample": f"{n}x{n}x{S} slab of the same matrix ({n * n * S} rows, {ci.size} nnz), {it - 1} iterations in {dt:.2f} s, scaled by {S}/{n}; "
                         "SpMV serial by construction, BLAS-1 on OpenBLAS threads"}

    line = {"metric": "cg_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lap3d7-{n} fp64 unpreconditioned CG, b=1/sqrt(N), x0=0 (BASELINE configs[2]); spmv = lap3d27-128 (configs[1])",
                       "rows": N, "nnz": nnz, "parallelism": "1 rank", "l2": "inputs larger than L2 (matrix 11.3 GB per iteration); no flush",
                       "step": "one CG iteration = 3 kernels (spmv+dot, update+nrm2, direction)"},
            "gbs": Bcg * its / 1e9, "frac_of_measured_peak": Bcg * its / 1e9 / peak, "algorithmic_bytes_per_step": Bcg, "final_residual": res,
            "spmv": spmv, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=GRID, help="edge of the 7-point problem (default 512 = BASELINE configs[2])")
    ap.add_argument("--ref-planes", type=int, default=24, help="z-planes of the slab the CPU reference is timed on")
    ap.add_argument("--spmv-reps", type=int, default=200, help="timed launches of the configs[1] SpMV (lower it under ncu)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        run_ours(args)
    except BaseException as ex:      # a failing rank must say why on stdout before torchrun tears the others down
        import traceback
        diag = {"bench_error": repr(ex), "rank": int(os.environ.get("RANK", "0")), "world": int(os.environ.get("WORLD_SIZE", "1")),
                "traceback": traceback.format_exc().splitlines()[-6:]}
        try:
            from hala_b200.capi import lib
            diag["hb_last_error"] = lib.hb_last_error().decode()
        except Exception as ex2:
            diag["hb_last_error"] = f"unavailable: {ex2!r}"
        diag.update(getattr(ex, "hb_diag", {}))
        print(json.dumps(diag), flush=True)
        sys.stderr.write(json.dumps(diag) + "\n")
        sys.stderr.flush()
        raise SystemExit(1)


if __name__ == "__main__":
    main()
