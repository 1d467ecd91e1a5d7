#!/usr/bin/env python
"""Executed warp-instructions of the analysis kernel per frame, split at the block barriers / hand-off arrivals of
the main-warp frame loop (SASS order), with the opcode mix of each region.
Usage: ncu_regions.py REPORT FRAMES [NW]"""
import collections, csv, io, re, subprocess, sys
rep, frames = sys.argv[1], float(sys.argv[2])
nw = int(sys.argv[3]) if len(sys.argv) > 3 else 8
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; ci = {x: i for i, x in enumerate(h)}
regions = []; cur = {"n": 0, "ops": collections.Counter(), "samples": 0, "static": 0, "end": "start"}
for r in rows[2:]:
    if len(r) < 10: continue
    src = r[ci["Source"]].strip()
    e = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = mm.group(2).split(".")[0] if mm else "?"
    cur["n"] += e; cur["ops"][op] += e; cur["samples"] += s; cur["static"] += 1
    if op == "BAR" or "PHASECHK" in src or op == "EXIT":
        cur["end"] = src[:50]
        regions.append(cur)
        cur = {"n": 0, "ops": collections.Counter(), "samples": 0, "static": 0, "end": ""}
regions.append(cur)
tot_s = sum(r["samples"] for r in regions)
for i, r in enumerate(regions):
    if r["n"] < 0.5 * frames: continue
    top = " ".join(f"{k}:{v / frames:.0f}" for k, v in r["ops"].most_common(9))
    print(f"{i:3d} static {r['static']:5d}  {r['n'] / frames:8.1f}/frame ({r['n'] / frames / nw:6.1f}/main warp)  samples {100 * r['samples'] / tot_s:5.1f}%  -> {r['end']}\n      {top}")
