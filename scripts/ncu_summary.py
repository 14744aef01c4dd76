"""Turns an .ncu-rep (read here, on the CPU box, with `ncu -i`) into the small tracked summaries under profiles/:
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x            -> x.metrics.csv (+ x.stalls.csv with --source)
   python scripts/ncu_summary.py --launches gpurun_out/launches.csv profiles/y -> y.launches.csv (per-kernel totals and shares)
Nothing here runs on the GPU box."""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def metrics(rep, dst):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(dst + ".metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel"] + [f"{m} [{units[i]}]" for m, i in cols] + ["dram_traffic_bytes", "dram_GBps_from_traffic"])
        for n, r in enumerate(rows[2:]):
            def val(name):
                i = hdr.index(name)
                v = float(r[i].replace(",", ""))
                u = units[i]
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0}.get(u, 1.0)
                return v * scale
            traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            t = val("gpu__time_duration.sum")
            w.writerow([n, r[ki][:70]] + [r[i] for _, i in cols] + [f"{traffic:.0f}", f"{traffic / t / 1e9:.1f}"])
    print("wrote", dst + ".metrics.csv")


def stalls(rep, dst, top=25):
    """per-source-line sample counts of the first kernel in the report (needs -lineinfo + --import-source on)"""
    rows = ncu_csv(rep, "source")
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address" and "Source" in r:
            hdr, start = r, i + 1
            break
    if hdr is None:
        print("no source page")
        return
    si, ci = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    reasons = [(c, i) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    body, by_reason = [], collections.Counter()
    for r in rows[start:]:
        if len(r) <= ci or r[0] in ("Address", "Kernel Name"):
            break                                   # next kernel of the report: only the first one is summarised
        s = float(r[ci].replace(",", "") or 0)
        top_reason = max(reasons, key=lambda ci_: float(r[ci_[1]].replace(",", "") or 0))[0] if s else ""
        body.append((s, r[si].strip(), top_reason))
        for name, i in reasons:
            by_reason[name] += float(r[i].replace(",", "") or 0)
    tot = sum(b[0] for b in body) or 1.0
    with open(dst + ".stalls.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", rows[0][1][:100] if rows and rows[0] and rows[0][0] == "Kernel Name" else ""])
        w.writerow(["stall_reason", "samples", "share_pct"])
        for name, s in by_reason.most_common(8):
            w.writerow([name, int(s), f"{100 * s / tot:.1f}"])
        w.writerow(["sass_instruction", "samples", "share_pct", "dominant_reason"])
        for s, src, why in sorted(body, key=lambda b: -b[0])[:top]:
            w.writerow([src[:120], int(s), f"{100 * s / tot:.1f}", why])
    print("wrote", dst + ".stalls.csv")


def launches(src, dst, ours=("spmv", "cg_", "dcg_", "csr_", "vec_", "reduce", "multi_", "pack_", "scale_", "axpy", "xpby", "gemv", "peer_", "sptrsv", "ilu", "spmm", "batch_")):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ms": 1e3, "ns": 1e-3, "us": 1.0, "s": 1e6, "second": 1e6}.get(r[ui], 1.0)
        name = r[ki]
        short = name.split("(")[0].replace("void ", "")[:80]
        a = agg.setdefault(short, [0, 0.0, 0, 0.0])
        a[0] += 1
        a[1] += v
        if v >= 10.0:                                 # launches that did work (iterations past the stop return at once: ~4 us)
            a[2] += 1
            a[3] += v
    is_ours = lambda k: any(o in k for o in ours) and not k.startswith(("at::", "at_cuda", "cub::", "thrust::", "nccl"))
    tot_ours = sum(a[1] for k, a in agg.items() if is_ours(k))
    with open(dst + ".launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_our_kernels_pct", "ours", "launches_over_10us", "mean_us_over_10us"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            mine = is_ours(k)
            w.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1] / a[0]:.2f}", f"{100 * a[1] / tot_ours:.1f}" if mine and tot_ours else "", int(mine),
                        a[2], f"{a[3] / a[2]:.2f}" if a[2] else ""])
    print("wrote", dst + ".launches.csv")


if __name__ == "__main__":
    a = sys.argv[1:]
    if a[0] == "--launches":
        launches(a[1], a[2])
    else:
        src = "--source" in a
        a = [x for x in a if x != "--source"]
        metrics(a[0], a[1])
        if src:
            stalls(a[0], a[1])
