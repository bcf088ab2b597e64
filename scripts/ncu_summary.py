#!/usr/bin/env python
"""Summarises an `ncu --set full` capture of one frame's kernels into a markdown table (profiles/) and
extracts the DRAM traffic of the integrate kernel for bench.py's roofline.traffic.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_frame_kernels.md [profiles/integrate_traffic.json]
"""
import csv
import json
import re
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SC = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
stall = [h for h in hdr if re.match(r"smsp__pcsamp_warps_issue_stalled_", h) and "not_issued" not in h]


def val(r, c):
    try:
        return float(r[idx[c]].replace(",", ""))
    except Exception:
        return 0.0


lines = ["| kernel | us | DRAM rd MB | DRAM wr MB | warp inst | issue active % | grid | regs | achieved occ % | thr/inst | top stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
total = 0.0
integ = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    t = val(r, "gpu__time_duration.sum")
    u = units[idx["gpu__time_duration.sum"]]
    t = t / 1000 if u == "ns" else (t * 1000 if u == "ms" else t)
    total += t
    rd = val(r, "dram__bytes_read.sum") * SC.get(units[idx["dram__bytes_read.sum"]], 1)
    wr = val(r, "dram__bytes_write.sum") * SC.get(units[idx["dram__bytes_write.sum"]], 1)
    s = sum(val(r, h) for h in stall) or 1
    top = sorted(((val(r, h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall), reverse=True)[:3]
    lines.append("| %s | %.1f | %.2f | %.2f | %.0f | %.1f | %.0f | %.0f | %.1f | %.1f | %s |" % (
        name, t, rd, wr, val(r, "smsp__inst_executed.sum"), val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        val(r, "launch__grid_size"), val(r, "launch__registers_per_thread"),
        val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        ", ".join("%s %.0f%%" % (k, v / s * 100) for v, k in top)))
    if "k_integrate" in name:
        integ.append((rd + wr) * 1e6)
lines.append("")
lines.append("sum of kernel durations: %.1f us (ncu serialises launches and flushes caches between replays: compare shares, not absolutes)" % total)
open(out_md, "w").write("\n".join(lines) + "\n")
if out_json and integ:
    json.dump({"dram_bytes_per_launch": sum(integ) / len(integ), "launches": len(integ), "source": rep.split("/")[-1],
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the integrate kernel, ncu --set full"},
              open(out_json, "w"), indent=1)
print("\n".join(lines))
