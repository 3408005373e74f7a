#!/usr/bin/env python
"""Opcode-sequence fingerprint of one kernel's SASS in a built library (addresses, register numbers and immediates that
encode code offsets stripped): used to show that a source refactor left the generated device code unchanged.
usage: sass_fingerprint.py lib.so kernel_name_substring"""
import hashlib
import re
import subprocess
import sys


def fingerprint(lib, sub):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    on, ops = False, []
    for ln in out.splitlines():
        if "Function :" in ln:
            on = sub in ln
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            ins = m.group(1).strip()
            ins = re.sub(r"0x[0-9a-f]+", "IMM", ins)
            ops.append(ins)
    return len(ops), hashlib.sha256("\n".join(ops).encode()).hexdigest()[:16], ops


if __name__ == "__main__":
    n, h, ops = fingerprint(sys.argv[1], sys.argv[2])
    print(n, h)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join(ops) + "\n")
