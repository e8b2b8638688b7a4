"""Developer tool: share of executed warp instructions per phase of k_digitise (markers in the source)."""
import collections, csv, importlib.util, os, sys
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('nl', os.path.join(here, 'ncu_lines.py')); nl = importlib.util.module_from_spec(spec); spec.loader.exec_module(nl)
src_csv, so, pat = sys.argv[1:4]
lm = nl.line_map(so, pat)
src = open(os.path.join(here, '..', 'pytestbeam_b200', 'csrc', 'ptb_kernels.cu')).read().split('\n')
def find(txt):
    for i, l in enumerate(src):
        if txt in l: return i + 1
marks = [('prolog', find('k_digitise(const __grid_constant__')), ('descriptors', find('// ---- cluster descriptors')),
         ('trim', find('// A box pixel can only be kept')), ('descr. store', find('int ncT = max(cb - ca + 1, 0)')),
         ('rounds setup', find('// Pass 0 stages the hits')), ('phase A', find('// ---- phase A')),
         ('phase B', find('// ---- phase B')), ('emit', find('// kept pixels of my particle so far')),
         ('flush', find('// ---- reserve room in the plane')), ('epilog', find('// ---- per-group counters for k_scan'))]
rows = list(csv.reader(open(src_csv))); hdr = rows[1]
ia, ii = hdr.index('Address'), hdr.index('Instructions Executed'); base = int(rows[2][ia], 16)
agg = collections.OrderedDict((m[0], 0) for m in marks); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    key = lm.get(int(r[ia], 16) - base); n = int(r[ii] or 0); tot += n
    if not key or key[0] != 'ptb_kernels.cu': continue
    name = None
    for m, ln in marks:
        if ln and key[1] >= ln: name = m
    if name: agg[name] += n
for k, v in agg.items(): print(f"{k:14s} {100 * v / tot:6.2f}%")
print("total warp instructions", tot)
