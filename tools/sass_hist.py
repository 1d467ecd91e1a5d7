#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel in an object file / shared library.
Usage: sass_hist.py FILE KERNEL_SUBSTRING [top]   (e.g. 'analyze_kernelILi4096ELb1ELb1ELb0')"""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, ops, n = None, collections.Counter(), 0
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(2).split(".")[0]
            if op == "NOP":
                continue
            ops[op] += 1
            n += 1
    print(f"{pat}: {n} instructions = {n * 16 / 1024:.1f} KB")
    for op, c in ops.most_common(top):
        print(f"  {op:10s} {c:6d}  {100.0 * c / n:5.1f}%")


if __name__ == "__main__":
    main()
