#!/usr/bin/env python
"""Opcode histogram of executed SASS from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import collections
import csv
import sys


def main(path, filt=None, top=40):
    rows = list(csv.reader(open(path)))
    segs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            if filt is None or filt in ",".join(r[1:]):
                segs.append(cur)
            continue
        if cur is not None:
            cur.append(r)
    best = None
    for seg in segs:
        hdr, data = seg[0], seg[1:]
        ie, si = hdr.index("Instructions Executed"), hdr.index("Source")
        tot, byop = 0, collections.Counter()
        for r in data:
            try:
                n = int(r[ie])
            except Exception:
                continue
            toks = r[si].strip().split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            byop[op.split(".")[0]] += n
            tot += n
        if best is None or tot > best[0]:
            best = (tot, byop, len(data))
    tot, byop, nins = best
    print("executed warp-instructions: %d  (static SASS instructions: %d)" % (tot, nins))
    for op, n in byop.most_common(top):
        print("%-12s %12d %.3f" % (op, n, n / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
