"""Developer tool: executed instructions per internal subroutine label (slow paths) of a kernel."""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, so, pat = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
inside, label, amap = False, "main", {}
for ln in txt.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = re.search(pat, ln) is not None
        label = "main"
        continue
    if ln.startswith("\t.section"):
        inside = False
    if not inside:
        continue
    m = re.match(r"(\$[\w$]+|[\w$.]+):", ln)
    if m and ("__internal" in m.group(1) or "__cuda" in m.group(1)):
        label = m.group(1)
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", ln)
    if m:
        amap[int(m.group(1), 16)] = label
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    lab = amap.get(int(r[ia], 16) - base, "?")
    n = int(r[ii] or 0)
    agg[lab][0] += n; agg[lab][1] += int(r[isamp] or 0); agg[lab][2] += 1
    tot += n
for lab, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{lab[:70]:70s} sass={a[2]:5d} inst%={100 * a[0] / tot:6.2f}")
