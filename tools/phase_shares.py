"""Developer tool: share of executed warp instructions, active lanes per instruction and stall samples per phase of
k_digitise (markers in the source), from an ncu source page joined with nvdisasm line info (tools/ncu_lines.py).

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_digitise > src.csv
    python tools/phase_shares.py src.csv pytestbeam_b200/_lib/libptb_b200.so k_digitiseINS_12PhiloxSourceELi128ELb0 [block]
"""
import collections, csv, importlib.util, os, sys
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location('nl', os.path.join(here, 'ncu_lines.py')); nl = importlib.util.module_from_spec(spec); spec.loader.exec_module(nl)
src_csv, so, pat = sys.argv[1:4]
blk = int(sys.argv[4]) if len(sys.argv) > 4 else -1          # which kernel instance of the csv (default: the last)
lm = nl.line_map(so, pat)
src = open(os.path.join(here, '..', 'pytestbeam_b200', 'csrc', 'ptb_kernels.cu')).read().split('\n')
def find(txt):
    for i, l in enumerate(src):
        if txt in l: return i + 1
# (code inlined from describe() / profile_entry() is attributed to its call site in the kernel)
marks = [('prolog', find('__device__ __forceinline__ void digitise_group')), ('fetch', find('cp_async_wait_all();   //')),
         ('descriptors', find('// ---- cluster descriptors')), ('census+reserve', find('{   // census, one packed word')),
         ('work list', find('// ---- the work list')), ('stage 1', find('// ---- stage 1')),
         ('stage 2', find('// ---- stage 2')), ('emit', find('// kept pixels of my particle up to my slot')),
         ('plane end', find('if (!more) break;')), ('epilog', find('// ---- per-group counters for k_scan')),
         ('ticket loop', find('k_digitise(const __grid_constant__')),
         # (markers of the pooled version, before the two-stage pixel walk)
         ('rounds setup', find('// inclusive prefixes of the profile entries')), ('phase A', find('// ---- phase A')),
         ('phase B', find('// ---- phase B')), ('emit (pooled)', find('// kept pixels of my particle in this step'))]
marks = [m for m in marks if m[1]]
marks.sort(key=lambda m: m[1])
lines = open(src_csv).read().split('\n')
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
blk = blk % (len(starts) - 1)
rows = list(csv.reader(lines[starts[blk]:starts[blk + 1]])); hdr = rows[1]
ia, ii, it, isamp = (hdr.index(k) for k in ('Address', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = int(rows[2][ia], 16)
agg = collections.OrderedDict((m[0], [0, 0, 0]) for m in marks); tot = [0, 0, 0]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    key = lm.get(int(r[ia], 16) - base); n, t, s = int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)
    tot[0] += n; tot[1] += t; tot[2] += s
    if not key or key[0] != 'ptb_kernels.cu': continue
    name = None
    for m, ln in marks:
        if ln and key[1] >= ln: name = m
    if name:
        a = agg[name]; a[0] += n; a[1] += t; a[2] += s
print(f"{'phase':16s} {'inst%':>7s} {'lanes':>6s} {'samp%':>7s}")
for k, (n, t, s) in agg.items(): print(f"{k:16s} {100 * n / tot[0]:7.2f} {t / max(n, 1):6.1f} {100 * s / tot[2]:7.2f}")
print(f"total warp instructions {tot[0]}, active lanes per instruction {tot[1] / tot[0]:.2f}")
