import csv, re, collections, subprocess, sys, os, tempfile
rep, lib, kname = sys.argv[1:4]
tmp=tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",os.path.abspath(lib)],cwd=tmp,capture_output=True)
lm={}
for f in os.listdir(tmp):
    if not f.endswith('.cubin') or 'persist' not in f: continue
    txt=subprocess.run(["nvdisasm","-g","-c",os.path.join(tmp,f)],capture_output=True,text=True).stdout.split('\n')
    infn=False; cur=None
    for l in txt:
        if l.startswith('//---') and '.text.' in l:
            infn = kname in l; continue
        if not infn: continue
        m=re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
        m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
        if m: lm[int(m.group(1),16)]=(cur, m.group(2).strip())
src=open('/root/repo/raymarchcl_b200/csrc/rm_scene_fused.cuh').read().split('\n')
def fn_of(f, line):
    if f!='rm_scene_fused.cuh': return f
    for i in range(line-1,0,-1):
        m=re.match(r'^(?:RM_DEV|RM_SHARED_FN|RM_FUSED_\w+)\s+[\w:<>]+\s+(\w+)\(', src[i-1])
        if m: return m.group(1)
    return '?'
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=next(i for i,r in enumerate(rows) if 'Address' in r and 'Source' in r)
hdr=rows[h]; col={n:i for i,n in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0,0,0,0]); lines=collections.defaultdict(lambda:[0,0,0,0,0])
base=None; mism=0
for r in rows[h+1:]:
    if len(r)<len(hdr): continue
    a=int(r[col['Address']],16)
    if base is None: base=a
    loc,sass=lm.get(a-base,(None,''))
    if sass.split(' ')[0] != r[col['Source']].strip().split(' ')[0]: mism+=1
    k=fn_of(*loc) if loc else None
    for g in (agg[k], lines[loc]):
        g[0]+=int(r[col['Instructions Executed']]); g[1]+=int(r[col['Thread Instructions Executed']]); g[2]+=int(r[col['# Samples']]); g[3]+=int(r[col['stall_no_inst']]); g[4]+=int(r[col['stall_long_sb']])
print("mismatched opcodes:", mism)
ti=sum(g[0] for g in agg.values()); ts=sum(g[2] for g in agg.values())
print(f"total warp instr {ti/1e9:.2f}e9 samples {ts}")
print(f"{'function':28s} inst%  smp%  lanes noinst% longsb%")
for k,g in sorted(agg.items(), key=lambda kv:-kv[1][2])[:22]:
    print(f"{str(k):28s} {100*g[0]/ti:5.1f} {100*g[2]/ts:5.1f} {g[1]/max(g[0],1):5.1f} {100*g[3]/max(g[2],1):6.1f} {100*g[4]/max(g[2],1):6.1f}")
if len(sys.argv)>4:
    print("--- top lines")
    for k,g in sorted(lines.items(), key=lambda kv:-kv[1][2])[:int(sys.argv[4])]:
        print(f"{str(k):36s} {100*g[0]/ti:5.1f} {100*g[2]/ts:5.1f} {g[1]/max(g[0],1):5.1f} {100*g[3]/max(g[2],1):6.1f} {100*g[4]/max(g[2],1):6.1f}")
