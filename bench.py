#!/usr/bin/env python
"""bench.py — the contract benchmark of hala_b200.

    python bench.py --gpus N --steps K --warmup W            our arm (libhalab200 on B200)
    python bench.py --impl reference --gpus N ...            the reference's own CPU path (oracle/_ref) on the host cores
    python bench.py --workload gmres ...                     BASELINE configs[3] instead of configs[2]

Workload (all N): BASELINE configs[2] — fp64 3-D 7-point Laplacian 512^3 (134,217,728 rows, 937,951,232 non-zeros),
unpreconditioned CG, b = 1/sqrt(N), x0 = 0, row-partitioned over the N ranks (strong scaling).  A "step" is one CG
iteration (SpMV fused with <p,Ap>, fused r update + ||r||^2, x and direction update); `value` = iterations per second with
everything resident in HBM.  At N = 1 the same line also carries BASELINE configs[1] — CSR SpMV GB/s on the fp64 27-point
Laplacian 128^3 — under "spmv" (BASELINE.json's metric names both), the reference's own GPU path (cuSPARSE + cuBLAS, built
unmodified into oracle/_ref/libhala_ref_gpu.so) on the same matrices under "vs_cusparse", and the same iteration through
the hala:: C++ template API under "template_api".
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
GMRES_GRID = 256    # configs[3]
GMRES_RESTART = 50


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


def kernel_profile(e, names, bytes_per_launch, peak, traffic_tags=None):
    """per-kernel in-context times of the solve that just ran with hb_ctx_profile on: CUDA events on the launching stream around
    every launch, inside the solver loop (hb_prof_mark) -> one roofline entry per kernel"""
    from hala_b200.capi import lib, check
    out = []
    for slot, (name, nbytes) in enumerate(zip(names, bytes_per_launch)):
        ms, cnt = C.c_double(0), C.c_longlong(0)
        check(lib.hb_ctx_profile_read(e.ctx, slot, C.byref(ms), C.byref(cnt)), "hb_ctx_profile_read")
        if cnt.value == 0:
            continue
        us = ms.value / cnt.value * 1e3
        out.append({"kernel": name, "bound": "hbm", "us_per_launch": us, "launches_timed": cnt.value, "algorithmic_bytes_per_launch": nbytes,
                    "achieved": nbytes / us / 1e3, "peak": peak, "unit": "GB/s", "frac": nbytes / us / 1e3 / peak,
                    "traffic": ncu_traffic(traffic_tags[slot]) if traffic_tags else None})
    return out


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation (hala::solve_cg / solve_gmres on cpu_engine, unmodified headers, built into
    oracle/_ref/libhala_ref.so) on the FULL bench matrix, generated on the host by the oracle's C generator: real iterations of
    the named configuration, no extrapolation.  The CPU needs ~2 s per iteration at 512^3, so the number of timed iterations is
    bounded (--ref-steps, default min(steps, 6)); what was timed is stated in the line.  Nothing of the product is loaded here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding
    lib, kind = binding.reference(), "reference"
    if lib is None:
        lib, kind = binding.oracle(), "port"
    gmres = args.workload == "gmres"
    n = args.grid
    t0 = time.perf_counter()
    p, i, v = binding.oracle().gen_stencil7(n, -1.5, 6.0, -0.5) if gmres else binding.oracle().gen_stencil7(n)
    gen_s = time.perf_counter() - t0
    rows = n ** 3
    b = np.full(rows, 1.0 / np.sqrt(rows))
    cores = os.cpu_count()
    if gmres:
        cyc = max(1, min((args.steps + GMRES_RESTART - 1) // GMRES_RESTART, args.ref_steps or 1))
        t0 = time.perf_counter()
        _, it = lib.gmres(p, i, v, b, 0.0, GMRES_RESTART, max_outer=cyc)
        dt = time.perf_counter() - t0
        done, metric, wl = it, "gmres_iters_per_s", f"convdiff7-{n} fp64 GMRES({GMRES_RESTART}) (BASELINE configs[3])"
        sample = f"the full matrix ({rows} rows, {i.size} nnz), {cyc} restart cycle(s) = {it} operator applications in {dt:.2f} s"
    else:
        k = args.ref_steps or max(2, min(args.steps, 6))
        lib.cg(p, i, v, b, 0.0, max_iter=2)                                  # warm-up: 1 iteration (page faults, BLAS threads)
        t0 = time.perf_counter()
        _, it = lib.cg(p, i, v, b, 0.0, max_iter=k + 1)
        dt = time.perf_counter() - t0
        done, metric, wl = it - 1, "cg_iters_per_s", f"lap3d7-{n} fp64 unpreconditioned CG, b=1/sqrt(N), x0=0 (BASELINE configs[2])"
        sample = f"the full matrix ({rows} rows, {i.size} nnz), {done} CG iterations in {dt:.2f} s after 1 warm-up iteration"
    its = done / dt
    line = {"impl": "reference", "metric": metric, "value": its, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "steps_timed": done, "same_config": True,
            "config": {"workload": wl, "rows": rows, "nnz": int(i.size), "host_generation_s": gen_s,
                       "note": "the reference's CPU SpMV is serial by construction (sparse/hala_sparse_utils.hpp:103-118); BLAS-1 runs on OpenBLAS "
                               f"with up to {cores} threads; the number of timed iterations is bounded so that the arm ends within minutes"},
            "cpu_baseline": {"value": its, "unit": "iterations/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": its, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- the reference's GPU path
def gpu_reference_leg(e, dev, args):
    """The unmodified reference built with -DHALA_ENABLE_CUDA (cuSPARSE cusparseSpMV ALG_DEFAULT + cuBLAS level 1) on the same box and
    the same device arrays: SpMV on configs[1], [4a], [4b]; hala::solve_cg(gpu_engine) on configs[0] and [2].
    Returns the `vs_cusparse` block (None when the library is absent)."""
    import torch
    import hala_b200 as hb
    from hala_b200 import devgen, matgen as mg
    from hala_b200.capi import lib, check
    from oracle import binding
    ref = binding.reference_gpu()
    if ref is None:
        return None
    out = {"reference": ref.version, "spmv": [], "cg": []}

    def spmv_pair(tag, tp, ti, tv, code, es, reps):
        N, nnz = tp.numel() - 1, ti.numel()
        x = torch.from_numpy(mg.probe_x(N, "c64" if code == 3 else "f64")).to(dev)
        y = torch.empty_like(x)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        gx, gy = devgen.torch_view(e, x), devgen.torch_view(e, y)
        for _ in range(5):
            A.gemv("N", 1.0, gx, 0.0, gy)
        e.timer_start()
        for _ in range(reps):
            A.gemv("N", 1.0, gx, 0.0, gy)
        us_ours = e.timer_stop() / reps * 1e3
        y_ours = y.clone()
        us_ref = ref.spmv_us(code, N, N, nnz, tp.data_ptr(), ti.data_ptr(), tv.data_ptr(), x.data_ptr(), y.data_ptr(), 5, reps)
        scale = float(torch.max(torch.abs(y)).item()) or 1.0
        B = mg.spmv_bytes(N, nnz, es)
        out["spmv"].append({"workload": tag, "rows": N, "nnz": nnz, "ours_us": us_ours, "cusparse_us": us_ref, "ours_gbs": B / us_ours / 1e3,
                            "cusparse_gbs": B / us_ref / 1e3, "speedup": us_ref / us_ours,
                            "max_abs_diff_over_max_abs_y": float(torch.max(torch.abs(y - y_ours)).item()) / scale})
        del A

    tp, ti, tv = devgen.stencil_slab("lap3d27", SPMV_GRID, 0, SPMV_GRID ** 3, device=dev)
    spmv_pair("lap3d27-128 fp64 (configs[1])", tp, ti, tv, 1, 8, 100)
    tp, ti, tv = devgen.stencil_slab("helmholtz7", 192, 0, 192 ** 3, dtype="c64", device=dev)
    spmv_pair("helmholtz7-192 complex<double> (configs[4a])", tp, ti, tv, 3, 16, 50)
    if not args.quick:
        p, i, v = mg.powerlaw(N=1 << 22, dtype="f64")
        tp, ti, tv = (torch.from_numpy(a).to(dev) for a in (p, i, v))
        spmv_pair("powerlaw 2^22 rows fp64, max row 65536 (configs[4b])", tp, ti, tv, 1, 8, 30)
        del p, i, v
    del tp, ti, tv
    torch.cuda.empty_cache()

    def cg_pair(tag, name, n, tol, max_iter):
        N = n ** (2 if name == "lap2d" else 3)
        tp, ti, tv = devgen.stencil_slab(name, n, 0, N, device=dev)
        nnz = ti.numel()
        b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
        x = torch.zeros(N, dtype=torch.float64, device=dev)
        gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
        A = hb.make_sparse_matrix(e, N, gp, gi, gv)
        it, res = C.c_int(0), C.c_double(0)
        res_ = {}
        for rep in range(2):                                            # first pass warms both libraries up
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            check(lib.hb_cg(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), float(tol), int(max_iter), C.byref(it), C.byref(res)), "hb_cg")
            torch.cuda.synchronize()
            res_["ours"] = (it.value, time.perf_counter() - t0)
            x.zero_()
            torch.cuda.synchronize()
            res_["ref"] = ref.cg(1, N, nnz, tp.data_ptr(), ti.data_ptr(), tv.data_ptr(), b.data_ptr(), x.data_ptr(), float(tol), int(max_iter))
        o_its, r_its = (res_["ours"][0] - 1) / res_["ours"][1], (res_["ref"][0] - 1) / res_["ref"][1]
        out["cg"].append({"workload": tag, "rows": N, "nnz": nnz, "tol": tol, "ours_operator_applications": res_["ours"][0],
                          "reference_gpu_operator_applications": res_["ref"][0], "ours_its": o_its, "reference_gpu_its": r_its, "speedup": o_its / r_its,
                          "what": "whole solve, wall clock around the call with a device synchronisation on both sides; ours = hb_cg, reference = "
                                  "hala::solve_cg(gpu_engine) of the stock headers (9 cuBLAS/cuSPARSE calls and 3 host round trips per iteration)"})
        del A, tp, ti, tv, b, x
        torch.cuda.empty_cache()

    cg_pair("lap2d5-1024 fp64 CG to 1e-8 (configs[0])", "lap2d", 1024, 1e-8, 10 ** 6)
    k = min(args.steps, 50)
    cg_pair(f"lap3d7-{args.grid} fp64 CG, {k} iterations (configs[2])", "lap3d7", args.grid, 0.0, k + 1)
    return out


def template_api_leg(args):
    """The same CG through the hala:: C++ template API (tests/_bin/template_bench, built against the reference's headers in the build
    container): hb_cg vs hala::solve_cg(gpu_engine, ...) with the identity tag / a copy lambda / the reference's own BLAS-1 loop."""
    exe = os.path.join(ROOT, "tests", "_bin", "template_bench")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ)
    import sysconfig
    env["LD_LIBRARY_PATH"] = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    try:
        r = subprocess.run([exe, str(args.grid), str(min(args.steps, 50))], capture_output=True, text=True, timeout=600, env=env)
        rows = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
        return rows if r.returncode == 0 and rows else {"error": (r.stdout + r.stderr)[-500:]}
    except Exception as ex:
        return {"error": repr(ex)}


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import hala_b200 as hb
    from hala_b200 import devgen, matgen as mg
    from hala_b200.capi import lib, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from hala_b200 import dist as hbdist
        return hbdist.run_bench(args, slab, ClockSampler, measured_peak, kernel_profile)
    torch.cuda.set_device(local)
    e = hb.gpu_engine(local)
    peak, peak_src = measured_peak()
    dev = f"cuda:{local}"
    if args.workload == "gmres":
        return run_gmres_single(args, e, dev, peak, peak_src)
    n = args.grid
    N = n ** 3

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
            "frac_of_8tbs_nominal": B1 / ms1 / 1e6 / 8000.0, "algorithmic_bytes": B1, "traffic": ncu_traffic("spmv_pipe_lap3d27_128")}
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

    # ---- roofline, one entry per kernel of the iteration: the same solve again with CUDA events around every launch (hb_ctx_profile)
    check(lib.hb_ctx_profile(e.ctx, 1), "hb_ctx_profile")
    e.timer_start()
    solve(args.steps)
    ms_prof = e.timer_stop()
    check(lib.hb_ctx_profile(e.ctx, 0), "hb_ctx_profile")
    Bk = mg.spmv_bytes(N, nnz, 8)
    kernels = kernel_profile(e, ["spmv_pipe_kernel<double,...,DOT> (CSR SpMV fused with <p,Ap>)", "cg_update_kernel<double,1> (r -= a Ap, ||r||^2, stop test)",
                                 "cg_direction_kernel<double,1> (x += a p, p = r + b p)"], [Bk, 3 * 8 * N, 5 * 8 * N], peak,
                             [f"spmv_pipe_dot_lap3d7_{n}", f"cg_update_lap3d7_{n}", f"cg_direction_lap3d7_{n}"])
    # the dominant kernel alone as well (30 back-to-back launches), the figure round 1 reported
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
    if kernels:
        roof = dict(kernels[0])
    else:
        roof = {"kernel": "spmv_pipe_kernel<double,...,DOT>", "bound": "hbm", "us_per_launch": kms * 1e3, "algorithmic_bytes_per_launch": Bk,
                "achieved": Bk / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": Bk / kms / 1e6 / peak, "traffic": None}
    roof.update({"peak_source": peak_src, "share_of_step": roof["us_per_launch"] / 1e3 / (ms / args.steps),
                 "us_per_launch_alone": kms * 1e3, "frac_alone": Bk / kms / 1e6 / peak,
                 "how": "CUDA events on the launching stream around every launch of this kernel inside the solver loop, over a second run of the same "
                        f"{args.steps} iterations (that run: {ms_prof / args.steps:.4f} ms/iteration with the events in); *_alone = 30 back-to-back launches"})
    del p_like, q

    # ---- the reference's own GPU path on the same box
    vs_cusparse = None
    if not args.no_gpu_ref:
        try:
            vs_cusparse = gpu_reference_leg(e, dev, args)
        except Exception as ex:
            vs_cusparse = {"error": repr(ex)}

    # ---- e2e: host buffers in, host buffer out (pinned), every copy inside the timed region
    e2e = None
    cpu = None
    if not args.no_e2e:
        hp, hi, hv = tp.cpu().pin_memory(), ti.cpu().pin_memory(), tv.cpu().pin_memory()
        hb_, hx = b.cpu().pin_memory(), torch.empty(N, dtype=torch.float64).pin_memory()
        dp, di, dv = torch.empty_like(tp), torch.empty_like(ti), torch.empty_like(tv)
        db, dx = torch.empty_like(b), torch.empty_like(x)
        h2d = sum(t.numel() * t.element_size() for t in (hp, hi, hv, hb_))
        d2h = hx.numel() * hx.element_size()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stream = torch.cuda.default_stream()        # the context runs on the legacy default stream; torch's default stream is the same one

        def e2e_solve(iters):
            ev[0].record(stream)
            for d, h in ((dp, hp), (di, hi), (dv, hv), (db, hb_)):
                check(lib.hb_memcpy_async(e.ctx, C.c_void_p(d.data_ptr()), C.c_void_p(h.data_ptr()), h.numel() * h.element_size(), 0))
            check(lib.hb_memset_zero(e.ctx, C.c_void_p(dx.data_ptr()), N * 8))
            ev[1].record(stream)
            Ah = C.c_void_p()
            check(lib.hb_csr_create(e.ctx, 1, N, N, nnz, C.c_void_p(dp.data_ptr()), C.c_void_p(di.data_ptr()), C.c_void_p(dv.data_ptr()), C.byref(Ah)))
            it, rs = C.c_int(0), C.c_double(0)
            check(lib.hb_cg(e.ctx, Ah, C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), 0.0, iters + 1, C.byref(it), C.byref(rs)))
            ev[2].record(stream)
            check(lib.hb_memcpy(e.ctx, C.c_void_p(hx.data_ptr()), C.c_void_p(dx.data_ptr()), N * 8, 1))
            ev[3].record(stream)
            lib.hb_csr_destroy(Ah)
            return it.value - 1

        e2e_solve(3)
        t0 = time.perf_counter()
        e.timer_start()
        d_it = e2e_solve(args.steps)
        ems = e.timer_stop()
        wall = time.perf_counter() - t0
        torch.cuda.synchronize()
        h2d_ms, solve_ms, d2h_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
        e2e = {"value": d_it / ems * 1e3, "unit": "iterations/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "ms_total": ems, "wall_ms": wall * 1e3, "h2d_ms": h2d_ms, "solve_ms": solve_ms, "d2h_ms": d2h_ms, "h2d_gbs": h2d / h2d_ms / 1e6,
               "what": f"CSR + b H2D from pinned host memory ({h2d / 1e9:.2f} GB), hb_csr_create, {args.steps} CG iterations, x D2H; all inside the timed region. "
                       "The upload is paid once per solve, so this figure grows with the step count: h2d_ms / solve_ms / d2h_ms split it"}
        del dp, di, dv, db, dx

        # ---- CPU baseline: the reference's own cpu_engine CG (oracle/_ref) on the SAME matrix (the pinned host copy), a few real iterations
        if not args.no_cpu:
            from oracle import binding
            ref, kind = binding.reference(), "reference"
            if ref is None:
                ref, kind = binding.oracle(), "port"
            cp, ci, cv = hp.numpy(), hi.numpy(), hv.numpy()
            cb = np.full(N, 1.0 / np.sqrt(N))
            cit = args.ref_steps or max(2, min(args.steps, 4))
            ref.cg(cp, ci, cv, cb, 0.0, max_iter=2)
            t0 = time.perf_counter()
            _, it = ref.cg(cp, ci, cv, cb, 0.0, max_iter=cit + 1)
            dt = time.perf_counter() - t0
            cpu = {"value": (it - 1) / dt, "unit": "iterations/s", "cores": os.cpu_count(), "kind": kind, "same_config": True,
                   "sample": f"the full bench matrix ({N} rows, {nnz} nnz), {it - 1} CG iterations in {dt:.2f} s after 1 warm-up iteration; "
                             "SpMV serial by construction (sparse/hala_sparse_utils.hpp:103-118), BLAS-1 on OpenBLAS threads"}
        del hp, hi, hv, hb_, hx
    del A, gp, gi, gv, tp, ti, tv, b, x, gb, gx
    torch.cuda.empty_cache()
    lib.hb_ctx_trim(e.ctx)

    template_api = None if args.no_template else template_api_leg(args)

    line = {"metric": "cg_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lap3d7-{n} fp64 unpreconditioned CG, b=1/sqrt(N), x0=0 (BASELINE configs[2]); spmv = lap3d27-128 (configs[1])",
                       "rows": N, "nnz": nnz, "parallelism": "1 rank", "l2": "inputs larger than L2 (matrix 11.3 GB per iteration); no flush",
                       "step": "one CG iteration = 3 kernels (spmv+dot, update+nrm2, direction)"},
            "gbs": Bcg * its / 1e9, "frac_of_measured_peak": Bcg * its / 1e9 / peak, "algorithmic_bytes_per_step": Bcg, "final_residual": res,
            "spmv": spmv, "roofline": roof, "roofline_kernels": kernels, "cpu_baseline": cpu, "e2e": e2e, "vs_cusparse": vs_cusparse,
            "template_api": template_api, "gpu_launches": launches, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)


def run_gmres_single(args, e, dev, peak, peak_src):
    """--workload gmres at N = 1: BASELINE configs[3], GMRES(50) on the 7-point convection-diffusion matrix 256^3; a step is one inner
    iteration (operator application + Gram-Schmidt against the basis); fixed budget of ceil(steps / 50) restart cycles (tolerance 0)."""
    import torch
    import hala_b200 as hb
    from hala_b200 import devgen, matgen as mg
    from hala_b200.capi import lib, check
    n = args.grid
    N = n ** 3
    tp, ti, tv = devgen.stencil_slab("convdiff7", n, 0, N, device=dev)
    nnz = ti.numel()
    gp, gi, gv = (devgen.torch_view(e, t) for t in (tp, ti, tv))
    A = hb.make_sparse_matrix(e, N, gp, gi, gv)
    b = torch.full((N,), 1.0 / np.sqrt(N), dtype=torch.float64, device=dev)
    x = torch.zeros(N, dtype=torch.float64, device=dev)
    cycles = max(1, (args.steps + GMRES_RESTART - 1) // GMRES_RESTART)

    def solve(cyc):
        x.zero_()
        it, res = C.c_int(0), C.c_double(0)
        check(lib.hb_gmres(e.ctx, A.h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), 0.0, cyc, GMRES_RESTART, 0, C.byref(it), C.byref(res)), "hb_gmres")
        return it.value, res.value

    with ClockSampler(int(dev.split(":")[1])) as clk:
        solve(1)
        torch.cuda.synchronize()
        l0 = e.launch_count()
        clk.mark_begin()
        e.timer_start()
        done, res = solve(cycles)
        ms = e.timer_stop()
        clk.mark_end()
    launches = e.launch_count() - l0
    its = done / ms * 1e3
    Bit = mg.gmres_iter_bytes(N, nnz, 8, GMRES_RESTART)
    line = {"metric": "gmres_iters_per_s", "value": its, "unit": "iterations/s", "n_gpus": 1, "steps": done, "warmup": GMRES_RESTART + 1,
            "ms_per_step": ms / done, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"convdiff7-{n} fp64 GMRES({GMRES_RESTART}), b=1/sqrt(N), x0=0, {cycles} restart cycle(s), tolerance 0 (BASELINE configs[3])",
                       "rows": N, "nnz": nnz, "parallelism": "1 rank", "l2": "inputs larger than L2; no flush",
                       "step": "one inner iteration: SpMV, multi-dot, multi-axpy + norm, normalise-and-append"},
            "gbs": Bit * its / 1e9, "frac_of_measured_peak": Bit * its / 1e9 / peak, "algorithmic_bytes_per_step": Bit, "estimated_residual": res,
            "roofline": {"bound": "hbm", "kernel": "gs_pipe_kernel<double> (multi-dot + multi-axpy over the Krylov basis) + spmv_pipe_kernel", "achieved": Bit * its / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": Bit * its / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "how": "whole inner iteration: algorithmic bytes at the mean basis size / measured time per iteration"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cg", choices=["cg", "gmres"], help="cg: BASELINE configs[2] (the headline); gmres: configs[3]")
    ap.add_argument("--grid", type=int, default=0, help="edge of the 7-point problem (default 512 = BASELINE configs[2]; gmres: 256 = configs[3])")
    ap.add_argument("--ref-steps", type=int, default=0, help="iterations the CPU reference is timed on (default: a few; ~2 s each at 512^3)")
    ap.add_argument("--spmv-reps", type=int, default=200, help="timed launches of the configs[1] SpMV (lower it under ncu)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the reference's cuSPARSE/cuBLAS leg")
    ap.add_argument("--no-template", action="store_true", help="skip the C++ template-API leg")
    ap.add_argument("--quick", action="store_true", help="skip the legs that need long host-side generation (power-law matrix)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if not args.grid:
        args.grid = GMRES_GRID if args.workload == "gmres" else GRID
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        run_ours(args)
    except BaseException as ex:      # a failing rank must say why on stdout before torchrun tears the others down
        if isinstance(ex, SystemExit) and not ex.code:
            raise
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
