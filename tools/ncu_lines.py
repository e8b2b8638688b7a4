"""Developer tool: join an ncu SASS-level source page (CSV) with nvdisasm line info and aggregate executed
instructions / stall samples per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv pytestbeam_b200/_lib/libptb_b200.so 'k_simulateINS_12Philox' [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def transparent_lines():
    """Source lines whose only content is the call of an inlined group body: instructions are attributed to the line
    inside that body (the next frame of the inline chain), not to the call."""
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pytestbeam_b200", "csrc", "ptb_kernels.cu")
    return {i + 1 for i, l in enumerate(open(src).read().split("\n")) if "digitise_group<Src, PC, MODE>(cfg" in l}


def line_map(so, kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    skip = transparent_lines()
    out, cur, inside, chain, fresh = {}, None, False, [], True
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = re.search(kernel_pat, ln) is not None
            continue
        if ln.startswith("\t.section"):
            inside = False
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), m.group(3).strip()))
            # the chain runs from the innermost frame to the outermost; the outermost wins unless it is a transparent call
            cur = chain[-1]
            if cur[0] == "ptb_kernels.cu" and cur[1] in skip and len(chain) > 1:
                cur = chain[-2]
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", ln)
        if m:
            out[int(m.group(1), 16)] = cur
            fresh = True
    return out


def main():
    src_csv, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(so, pat)
    lines = open(src_csv).read().split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
    rows = list(csv.reader(lines[starts[0]:starts[1]]))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ino = hdr.index("stall_no_inst")
    base = int(rows[2][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        off = int(r[ia], 16) - base
        key = lm.get(off) or ("?", 0, "")
        n, s, ni = int(r[ii] or 0), int(r[isamp] or 0), int(r[ino] or 0)
        a = agg[key[:2]]
        a[0] += n; a[1] += s; a[2] += ni; a[3] += 1
        tot[0] += n; tot[1] += s; tot[2] += ni
    print(f"total warp-instr {tot[0]:.4g}, samples {tot[1]}, no_inst samples {tot[2]}")
    print(f"{'file:line':32s} {'sass':>5s} {'inst%':>7s} {'samp%':>7s} {'noinst%':>8s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{key[0] + ':' + str(key[1]):32s} {a[3]:5d} {100 * a[0] / tot[0]:7.2f} {100 * a[1] / tot[1]:7.2f} "
              f"{100 * a[2] / max(tot[2], 1):8.2f}")


if __name__ == "__main__":
    main()
