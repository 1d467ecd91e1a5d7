#!/usr/bin/env python
"""Split the executed warp-instructions of the analysis kernel into main-warp and tail-warp work.
An instruction executed by every main warp runs NW x frames times; tail code runs <= 1 x frames.
Usage: ncu_roles.py REPORT FRAMES NW"""
import csv, io, subprocess, sys, collections, re
rep, frames, nw = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; ci = {x: i for i, x in enumerate(h)}
main = tail = 0; mainn = tailn = 0; ms = ts = 0
tail_ops = collections.Counter(); main_ops = collections.Counter()
for r in rows[2:]:
    if len(r) < 10: continue
    e = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    if e == 0: continue
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
    op = mm.group(2).split(".")[0] if mm else "?"
    if e > 1.6 * frames:
        main += e; mainn += 1; ms += s; main_ops[op] += e
    else:
        tail += e; tailn += 1; ts += s; tail_ops[op] += e
print(f"main warps: {main / frames:8.0f} warp-instr/frame ({main / frames / nw:6.0f} per warp), {mainn} hot SASS instr = {mainn * 16 / 1024:.1f} KB, samples {ms}")
print(f"tail warps: {tail / frames:8.0f} warp-instr/frame, {tailn} executed SASS instr = {tailn * 16 / 1024:.1f} KB, samples {ts}")
print("main ops:", " ".join(f"{o}:{c / frames:.0f}" for o, c in main_ops.most_common(22)))
print("tail ops:", " ".join(f"{o}:{c / frames:.0f}" for o, c in tail_ops.most_common(22)))
