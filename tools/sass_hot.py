#!/usr/bin/env python
"""Hot straight-line regions of executed SASS (ncu --page source --print-source sass --csv)."""
import csv
import sys


def main(path, top=10, show=30):
    rows = list(csv.reader(open(path)))
    segs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            segs.append(cur)
            continue
        if cur is not None:
            cur.append(r)

    def total(s):
        ie = s[0].index("Instructions Executed")
        return sum(int(r[ie]) for r in s[1:] if r[ie].isdigit())
    seg = max(segs, key=total)
    hdr, data = seg[0], seg[1:]
    ie, si = hdr.index("Instructions Executed"), hdr.index("Source")
    prev, start, acc, out = None, 0, [], []
    for i, r in enumerate(data):
        n = int(r[ie]) if r[ie].isdigit() else 0
        if prev is None or abs(n - prev) > 0.02 * max(n, prev, 1):
            if acc:
                out.append((start, i - 1, prev, len(acc), acc))
            start, acc = i, []
        acc.append(r[si].strip())
        prev = n
    out.append((start, len(data) - 1, prev, len(acc), acc))
    tot = sum(a[2] * a[3] for a in out)
    for a in sorted(out, key=lambda a: -a[2] * a[3])[:top]:
        print("sass %4d-%4d exec/instr=%9d ninstr=%3d share=%.3f" % (a[0], a[1], a[2], a[3], a[2] * a[3] / tot))
        for ln in a[4][:show]:
            print("       ", ln[:100])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
