"""Developer tool: break the instructions attributed to one line of digitise_group (typically the call of an inlined
helper such as describe()) down by the source lines of the callee, from an ncu source page and nvdisasm -gi line info.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_digitise > src.csv
    python tools/ncu_callee_lines.py src.csv pytestbeam_b200/_lib/libptb_b200.so k_digitiseINS_12PhiloxSourceELi128ELi0 <line> [depth]
"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, pat, target = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
inside=False; chain=[]; fresh=True; m_off={}
for ln in txt:
    if ln.startswith("\t.section\t.text."):
        inside = pat in ln; continue
    if ln.startswith("\t.section"): inside=False
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh: chain=[]; fresh=False
        chain.append((m.group(1).split('/')[-1], int(m.group(2)))); continue
    m=re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S+)", ln)
    if m:
        m_off[int(m.group(1),16)]=(list(chain), m.group(2)); fresh=True
lines=open(src_csv).read().split('\n')
starts=[i for i,l in enumerate(lines) if l.startswith('"Kernel Name"')]+[len(lines)]
rows=list(csv.reader(lines[starts[0]:starts[1]])); hdr=rows[1]
ia,ii,it=(hdr.index(k) for k in ('Address','Instructions Executed','Thread Instructions Executed'))
base=int(rows[2][ia],16)
agg=collections.defaultdict(lambda:[0,0,0]); tot=0; ops=collections.Counter()
for r in rows[2:]:
    if len(r)<len(hdr): continue
    ch,op=m_off.get(int(r[ia],16)-base,([],''))
    n=int(r[ii] or 0); t=int(r[it] or 0); tot+=n
    # find frame index of target line in ptb_kernels.cu
    idx=[i for i,f in enumerate(ch) if f==('ptb_kernels.cu',target)]
    if not idx: continue
    i=idx[0]
    key=ch[max(i-depth,0)] if i>0 else ('self',target)
    a=agg[key]; a[0]+=n; a[1]+=t; a[2]+=1
    ops[op.split('.')[0]]+=n
print("total",tot)
s=0
for k,a in sorted(agg.items(), key=lambda kv:(kv[0][0],kv[0][1])):
    print(f"{k[0]}:{k[1]:<5d} sass {a[2]:4d} inst% {100*a[0]/tot:6.2f} lanes {a[1]/max(a[0],1):5.1f}"); s+=a[0]
print("sum%",100*s/tot)
print(ops.most_common(25))
