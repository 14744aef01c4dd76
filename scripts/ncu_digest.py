"""Digest of an .ncu-rep into one small JSON (run where the report lies — on the GPU box the report itself is too big to bring back):
   python scripts/ncu_digest.py /tmp/x.ncu-rep gpurun_out/x      -> gpurun_out/x.digest.json
Per profiled kernel instance: the launch shape, the memory / pipe metrics the round's notes quote, the stall reasons (share of all
warp-state samples) and the SASS instructions that collected the most samples."""
import collections, csv, io, json, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__shared_mem_config_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0}


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, dst):
    raw = page(rep, "raw")
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ki = hdr.index("Kernel Name")
    kernels = []
    for r in rows:
        m = {}
        for name in METRICS:
            if name in hdr:
                i = hdr.index(name)
                try:
                    m[name] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    m[name] = r[i]
        t = m.get("gpu__time_duration.sum", 0) or 1e-30
        m["dram_traffic_bytes"] = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        m["dram_GBps"] = m["dram_traffic_bytes"] / t / 1e9
        kernels.append({"kernel": r[ki][:160], "metrics": m})
    src = page(rep, "source", ("--print-source", "sass"))
    blocks, cur, pending_name = [], None, ""
    for r in src:
        if r and r[0] == "Kernel Name":
            pending_name = r[1] if len(r) > 1 else ""
            cur = None
        elif r and r[0] == "Address" and "Source" in r:
            cur = {"hdr": r, "rows": [], "name": pending_name}
            blocks.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    base = lambda n: n.replace("void ", "").split("<")[0].split("(")[0].strip()
    if len(blocks) == 2 * len(kernels):             # this ncu prints every kernel's table twice
        blocks = blocks[::2]
    if len(blocks) != len(kernels):
        print("warning: source page does not line up with the raw page; stall data dropped", len(blocks), len(kernels))
        blocks = []
    else:                                            # same count: the pages list the instances in the same order
        bad = [(k["kernel"][:40], b["name"][:40]) for k, b in zip(kernels, blocks) if base(k["kernel"])[-12:] != base(b["name"])[-12:]]
        if bad:
            print("note: kernel names differ between the pages (matched by position):", bad[:3])
    for k, b in zip(kernels, blocks):
        h = b["hdr"]
        si, ci = h.index("Source"), h.index("# Samples")
        reasons = [(c, i) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = collections.Counter()
        body = []
        for r in b["rows"]:
            s = float(r[ci].replace(",", "") or 0)
            if s:
                top = max(reasons, key=lambda ci_: float(r[ci_[1]].replace(",", "") or 0))[0]
                body.append((s, r[si].strip()[:110], top))
            for name, i in reasons:
                tot[name] += float(r[i].replace(",", "") or 0)
        allsamp = sum(tot.values()) or 1.0
        k["stall_share_pct"] = {n: round(100 * v / allsamp, 1) for n, v in tot.most_common(7)}
        k["top_sass"] = [{"sass": s_, "share_pct": round(100 * s / allsamp, 1), "reason": why} for s, s_, why in sorted(body, key=lambda x: -x[0])[:10]]
        k["sass_instructions"] = len(b["rows"])
    with open(dst + ".digest.json", "w") as f:
        json.dump(kernels, f, indent=1)
    print("wrote", dst + ".digest.json", len(kernels), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
