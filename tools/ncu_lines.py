#!/usr/bin/env python3
"""Join an ncu SASS-level source page with nvdisasm line info: per source line, the share of
executed warp instructions, of stall samples, and the average active lanes.
usage: ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [top]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
linemap = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            infn = kname in ln
            continue
        if not infn: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
        if m: linemap[int(m.group(1), 16)] = (cur, m.group(2).strip())
    if linemap: break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
col = {n: hdr.index(n) for n in ["Address", "# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_branch_resolving", "stall_short_sb"] if n in hdr}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[col["Address"]], 16)
    if base is None: base = a
    loc = linemap.get(a - base, (None, ""))[0]
    g = agg[loc]
    g[0] += int(r[col["Instructions Executed"]]); g[1] += int(r[col["Thread Instructions Executed"]])
    g[2] += int(r[col["# Samples"]]); g[3] += int(r[col["stall_no_inst"]]); g[4] += int(r[col["stall_long_sb"]])
ti = sum(g[0] for g in agg.values()); ts = sum(g[2] for g in agg.values())
print(f"total warp instructions {ti:.4g}, samples {ts}")
print(" inst%  smp%  lanes  noinst% longsb%  location")
for loc, g in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*g[0]/ti:6.2f} {100*g[2]/ts:5.1f} {g[1]/max(g[0],1):6.1f} {100*g[3]/max(g[2],1):7.1f} {100*g[4]/max(g[2],1):7.1f}  {loc}")
