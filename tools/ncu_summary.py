#!/usr/bin/env python
"""Turn an ncu report into the small CSV summaries kept under profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep  profiles/<tag>_ncu_full.csv [--traffic profiles/wilson_dslash_traffic.json]
  python tools/ncu_summary.py list   gpurun_out/launches.csv  profiles/<tag>_launch_list_summary.csv
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def full(rep, out, traffic=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i + 1}" for i in range(len(data))])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [d[i] for d in data])
    if traffic:
        d = data[0]
        tot = sum(float(d[hdr.index(k)]) * UNIT.get(units[hdr.index(k)], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        json.dump({"kernel": d[hdr.index("Kernel Name")], "dram_bytes_per_launch": tot, "source": out}, open(traffic, "w"))


def launch_list(src, out):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "share_pct"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(t, 1), round(t / n, 2), round(100 * t / tot, 2)])


if __name__ == "__main__":
    mode, src, out = sys.argv[1:4]
    if mode == "full":
        full(src, out, sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None)
    else:
        launch_list(src, out)
