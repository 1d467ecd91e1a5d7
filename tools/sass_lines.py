#!/usr/bin/env python
"""Static SASS attribution per source line for one kernel (needs -lineinfo).
Usage: sass_lines.py OBJECT KERNEL_SUBSTRING [--regions file:lo-hi=name,...] [--top N]
Prints instruction counts per (file, line) and per named region.  For the straight-line main loop of the
analysis kernel the static count is the per-warp dynamic count per frame."""
import argparse
import collections
import glob
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regions", default="")
    ap.add_argument("--dump", action="store_true", help="print the annotated SASS of the kernel")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=td, capture_output=True)
        cubins = glob.glob(os.path.join(td, "*.cubin"))
        text = subprocess.run(["nvdisasm", "-g", "-c"] + cubins, capture_output=True, text=True).stdout
    lines = text.splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith("//---") and ".text." in l and a.kernel in l:
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found")
    per = collections.Counter()
    ops_per = collections.defaultdict(collections.Counter)
    cur = ("?", 0)
    n = 0
    for l in lines[start + 1:]:
        if l.startswith("//---"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(2).split(".")[0]
            if op == "NOP":
                continue
            if a.dump:
                print(f"{cur[0]}:{cur[1]:<5d} {l.strip()}")
            per[cur] += 1
            ops_per[cur][op] += 1
            n += 1
    print(f"{a.kernel}: {n} instructions ({n * 16 / 1024:.1f} KB)")
    if a.regions:
        print("== regions")
        for spec in a.regions.split(","):
            rng, name = spec.split("=")
            f, lh = rng.split(":")
            lo, hi = (int(x) for x in lh.split("-"))
            tot = sum(c for (ff, ll), c in per.items() if ff == f and lo <= ll <= hi)
            ops = collections.Counter()
            for (ff, ll), c in ops_per.items():
                if ff == f and lo <= ll <= hi:
                    ops.update(c)
            print(f"  {name:24s} {tot:6d}   " + " ".join(f"{o}:{c}" for o, c in ops.most_common(10)))
    print("== top lines")
    for (f, ln), c in per.most_common(a.top):
        print(f"  {f}:{ln:<5d} {c:5d}  " + " ".join(f"{o}:{k}" for o, k in ops_per[(f, ln)].most_common(6)))


if __name__ == "__main__":
    main()
