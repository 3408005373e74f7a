#!/usr/bin/env python
"""Per-kernel SASS opcode histograms of a built library -> JSON (kernel name -> {opcode.with.modifiers: count});
`sass_all_hist.py a.json b.json` diffs two of them.  Used to show that a source-level refactor (moving inline functions
between headers, host/device qualifiers) left every kernel's generated instruction mix unchanged."""
import collections
import json
import re
import subprocess
import sys


def dump(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    res, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]+", r"_ANON_\1", m.group(1))
            cur = res.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur is not None:
            t = m.group(1).split()
            cur[t[1] if t[0].startswith("@") else t[0]] += 1
    return {k: dict(v) for k, v in res.items()}


if __name__ == "__main__":
    if sys.argv[1].endswith(".json"):
        a, b = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
        bad = 0
        for k in sorted(set(a) | set(b)):
            if a.get(k) != b.get(k):
                bad += 1
                da, db = a.get(k, {}), b.get(k, {})
                diff = {o: (da.get(o, 0), db.get(o, 0)) for o in set(da) | set(db) if da.get(o, 0) != db.get(o, 0)}
                print("DIFF", k[:90], "instrs %d -> %d" % (sum(da.values()), sum(db.values())), dict(list(diff.items())[:6]))
        print("%d kernels, %d differ" % (len(set(a) | set(b)), bad))
    else:
        json.dump(dump(sys.argv[1]), open(sys.argv[2], "w"))
        print("wrote", sys.argv[2])
