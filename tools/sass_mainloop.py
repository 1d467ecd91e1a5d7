#!/usr/bin/env python
"""Static size of the main-warp frame loop of the analysis kernel: SASS instructions in layout order from the
hop mbarrier wait (SYNCS.PHASECHK) to the FULL hand-off (BAR.ARV), split at the block barriers.
Usage: sass_mainloop.py OBJECT KERNEL_SUBSTRING"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, ins = None, []
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and sys.argv[2] in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m: ins.append(m.group(2).strip())
start = next(i for i, s in enumerate(ins) if "SYNCS.PHASECHK" in s)
end = max(i for i, s in enumerate(ins) if "BAR.ARV" in s and ("0x2" in s or "0x3" in s) and i > start and i < start + 3000)
seg = ins[start:end + 1]
print(f"frame loop: {len(seg)} instructions ({len(seg) * 16 / 1024:.1f} KB)")
prev = 0
for i, s in enumerate(seg):
    if "BAR.SYNC" in s or "BAR.ARV" in s:
        print(f"  +{i - prev:4d}  -> {s[:60]}")
        prev = i
