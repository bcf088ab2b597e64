#!/usr/bin/env python
"""Hottest SASS instructions of one kernel in an ncu report (needs --import-source on / -lineinfo builds).
usage: python scripts/ncu_hot.py report.ncu-rep kernel_name [top=30] [launch_index=0]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the dump repeats a header block per launch; take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
ia, isrc, iall, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
def num(x):
    try: return int(x)
    except Exception: return 0
tot = sum(num(r[iall]) for r in data); totex = sum(num(r[iex]) for r in data)
print(f"{kern}: {len(data)} SASS instructions, {tot} stall samples, {totex} warp-instructions executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for k, r in enumerate(data):
    r.append(k)
for r in sorted(data, key=lambda r: -num(r[iall]))[:top_n]:
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"#{r[-1]:5d} samp {num(r[iall]):6d} ({100.0*num(r[iall])/max(tot,1):4.1f}%) exec {num(r[iex]):8d}  {r[isrc][:64]:64s} {st}")
