#!/usr/bin/env python
"""Warp-instructions executed / stall samples per SOURCE line of one kernel: joins the SASS page of an ncu report with the
line table of the object file (nvdisasm -g), instruction by instruction. The object must be the build that was profiled.
usage: python scripts/ncu_lines.py report.ncu-rep kernel_name object.o [top=40] [symbol_substring_in_object=kernel_name]"""
import csv, os, re, subprocess, sys, tempfile
rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sym = sys.argv[5] if len(sys.argv) > 5 else kern
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    data.append(r)
iex, ism = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], ("?", 0), False
for ln in dis:
    if ln.startswith("//---------------------"):
        on = (".text." in ln) and (sym in ln)
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln): lines.append(cur)
if len(lines) != len(data):
    print(f"warning: {len(lines)} instructions in the object, {len(data)} in the report (different build?)")
agg = {}
for k, r in enumerate(data[:len(lines)]):
    a = agg.setdefault(lines[k], [0, 0])
    try: a[0] += int(r[iex] or 0); a[1] += int(r[ism] or 0)
    except ValueError: pass
te, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{kern}: {te} warp-instructions, {ts} stall samples")
src_cache = {}
def src(f, n):
    for d in ("dynslam_b200/csrc", "."):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:100] if 0 < n <= len(src_cache[p]) else ""
    return ""
for (f, n), (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{f}:{n:<4d} exec {ex:9d} ({100.0*ex/max(te,1):4.1f}%)  samp {sm:5d} ({100.0*sm/max(ts,1):4.1f}%)  {src(f, n)}")
