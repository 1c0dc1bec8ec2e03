#!/usr/bin/env python3
"""Turn the scratch ncu outputs under gpurun_out/ into the small, committed summaries under profiles/.

  python tools/ncu_summary.py <tag> [--rep name=gpurun_out/x.ncu-rep ...] [--launches gpurun_out/launches.csv]
writes profiles/<tag>_<name>.json (selected raw metrics of the first profiled launch), profiles/<tag>_launches.md
(per-kernel launch count / total / average / share from the gpu__time_duration pass) and updates
profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def rep_summary(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = {"value": float(vals[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[w] = {"value": vals[i], "unit": units[i]}
        res.append(d)
    return res


def launches_summary(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > vi:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
            d[r[ki]].append(v * scale)
    tot = sum(sum(v) for v in d.values()) or 1.0
    lines = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        lines.append("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k[:110], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    return "\n".join(lines)


def main():
    tag = sys.argv[1]
    args = sys.argv[2:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    i = 0
    while i < len(args):
        if args[i] == "--rep":
            name, path = args[i + 1].split("=", 1)
            s = rep_summary(path)
            safe = "".join(ch if ch.isalnum() else "_" for ch in name).strip("_")
            json.dump(s, open(os.path.join(ROOT, "profiles", "%s_%s.json" % (tag, safe)), "w"), indent=1)
            k = s[0]
            rd, wr = k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum")
            if rd and wr:
                traffic[name] = rd["value"] * UNIT.get(rd["unit"], 1.0) + wr["value"] * UNIT.get(wr["unit"], 1.0)
            i += 2
        elif args[i] == "--launches":
            cmd = args[i + 2] if i + 2 < len(args) and not args[i + 2].startswith("--") else ""
            md = "# ncu launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` -- cold-cache, serialised: compare SHARES.\n\n%s\n\n%s\n" % (
                tag, ("command: `%s`" % cmd) if cmd else "", launches_summary(args[i + 1]))
            open(os.path.join(ROOT, "profiles", "%s_launches.md" % tag), "w").write(md)
            i += 3 if cmd else 2
        else:
            i += 1
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
