"""Extract the numbers the REAL crate printed in /root/reference/output.log into tests/golden/ref_log_onsets.json.

output.log is a run of the reference's live onset detector (onset.rs:413-449 log lines): one calibration onset
(`beat_pos`, `transport beat`, `target_samples`, `event_samples`, residual) and the fired onsets that followed
(`onset @ beat B (raw offset R, flux=F, burst=X/C)`).  The audio that produced them is not in the repository, so these
lines cannot pin magnitudes -- but they are outputs of the reference itself, and they pin what can be checked without
the input: the gating constants (onset.rs:356, :67-83), the window-centre offset lattice (onset.rs:386-387 with the
256 / 64 geometry and 1024-sample slots) and the f64 arithmetic of MusicalTransport::stamp_onset (timing.rs:311-337)
to the last printed digit.  Only numbers are copied; run it in the build container (the GPU box has no /root/reference):

    python tests/golden/extract_ref_log.py
"""
import json
import os
import re

SRC = "/root/reference/output.log"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_log_onsets.json")


def main():
    cal = {}
    onsets = []
    for line in open(SRC, encoding="utf-8", errors="replace"):
        m = re.search(r"beat_pos: ([0-9.eE+-]+), transport beat: ([0-9.eE+-]+), target_samples: (-?\d+), "
                      r"event_samples: (-?\d+)", line)
        if m:
            cal.update(beat_pos=m.group(1), transport_beat=m.group(2), target_samples=int(m.group(3)),
                       event_samples=int(m.group(4)))
            continue
        m = re.search(r"onset calibration: residual=([0-9.]+)ms \((-?\d+) samples\) at target frame (-?\d+)", line)
        if m:
            cal.update(residual_ms=m.group(1), residual_samples=int(m.group(2)), target_frame=int(m.group(3)))
            continue
        m = re.search(r"onset @ beat ([0-9.]+) \(raw offset (-?\d+), flux=([0-9.]+), burst=([0-9.]+)/(\d+)\)", line)
        if m:
            ts = re.match(r"\d{4}-\d\d-\d\dT(\d\d):(\d\d):(\d\d\.\d+)Z", line)      # wall clock of the log call
            t = int(ts.group(1)) * 3600 + int(ts.group(2)) * 60 + float(ts.group(3))
            onsets.append({"t": round(t, 6), "beat": float(m.group(1)), "raw_offset": int(m.group(2)),
                           "flux": float(m.group(3)), "max_excess": float(m.group(4)), "burst_count": int(m.group(5))})
    # the f64 values are kept as the strings Rust's `{}` printed (shortest round-trip representation)
    json.dump({"source": "reference output.log (onset.rs:413-449 log lines), numbers only",
               "calibration": cal, "onsets": onsets}, open(OUT, "w"), indent=1)
    print(f"{len(onsets)} onsets, calibration keys {sorted(cal)} -> {OUT}")


if __name__ == "__main__":
    main()
