#!/usr/bin/env python3
"""Static code size of one kernel by source function (nvdisasm line info), no GPU needed:
   python tools/sass_by_function.py raymarchcl_b200/libraymarch_b200.so 'k_render_persistILb0ELi6ELi256ELi5E' [top lines]
The default kernel is bound by instruction fetch as much as by issue: this is the map used for the code diet."""
import collections, os, re, subprocess, sys, tempfile

lib, kname = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "../raymarchcl_b200/csrc/rm_scene_fused.cuh")).read().split("\n")


def fn_of(f, line):
    if f != "rm_scene_fused.cuh":
        return f
    for i in range(line - 1, 0, -1):
        m = re.match(r"^(?:RM_DEV|RM_SHARED_FN|RM_FUSED_\w+)\s+[\w:<>]+\s+(\w+)\(", src[i - 1])
        if m:
            return m.group(1)
    return "?"


for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    infn, cur, sect = False, None, None
    per_fn, per_line, per_sect, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for l in txt:
        if l.startswith("//---") and ".text." in l:
            infn = kname in l
            sect = l.split(".text.")[1].split()[0]
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            k = fn_of(*cur) if cur else None
            per_fn[k] += 1
            per_line[cur] += 1
            per_sect[sect] += 1
            op = m.group(2).strip().split(" ")
            op = op[1] if op[0].startswith("@") else op[0]
            ops[op.split(".")[0]] += 1
    if not per_fn:
        continue
    tot = sum(per_fn.values())
    print(f"{f}: {tot} instructions = {tot * 16 / 1024:.1f} KiB")
    for s, n in per_sect.most_common():
        print(f"  section {s[:90]:90s} {n:5d}")
    for k, n in per_fn.most_common(40):
        print(f"  {str(k):28s} {n:5d} {100 * n / tot:5.1f} %")
    print("  opcodes:", ", ".join(f"{o} {n}" for o, n in ops.most_common(24)))
    if len(sys.argv) > 3:
        for k, n in per_line.most_common(int(sys.argv[3])):
            print(f"  {str(k):40s} {n:5d}")
