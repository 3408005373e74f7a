#!/usr/bin/env python
"""Executed warp-instructions and stall samples per SOURCE line: joins `ncu --page source --print-source sass --csv`
(per-instruction counts, in function order) with `nvdisasm -g -c` of the same cubin (per-instruction file:line markers).
usage: sass_lines.py ncu_sass.csv nvdisasm.txt mangled_substr kernel_substr [top]"""
import collections
import csv
import re
import sys


def ncu_segment(path, filt):
    rows = list(csv.reader(open(path)))
    segs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            if filt in ",".join(r[1:]):
                segs.append(cur)
            continue
        if cur is not None:
            cur.append(r)
    return segs[-1]


def dis_lines(path, mangled):
    out, on, cur = [], False, ("?", 0)
    for ln in open(path):
        if ln.startswith("//---") and ".text." in ln:
            on = mangled in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main(ncu_csv, dis, mangled, kfilt, top=40):
    seg = ncu_segment(ncu_csv, kfilt)
    hdr, data = seg[0], [r for r in seg[1:] if len(r) > 5]
    ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    dl = dis_lines(dis, mangled)
    if len(dl) != len(data):
        print("warning: %d SASS rows in ncu vs %d in nvdisasm" % (len(data), len(dl)))
    byline, samples = collections.Counter(), collections.Counter()
    tot = tots = 0
    for r, (loc, _) in zip(data, dl):
        n = int(r[ie]) if r[ie].isdigit() else 0
        s = int(r[ss]) if r[ss].isdigit() else 0
        byline[loc] += n
        samples[loc] += s
        tot += n
        tots += s
    print("total warp-instructions %d, stall samples %d" % (tot, tots))
    for loc, n in byline.most_common(top):
        print("%-22s %5d  instr %9d %.3f   samples %6d %.3f" % (loc[0], loc[1], n, n / tot, samples[loc],
                                                                  samples[loc] / max(tots, 1)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 40)
