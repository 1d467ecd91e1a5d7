#!/bin/bash
python - <<'P'
import sys, importlib, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import signals
from oracle import aa_oracle_py as O
aa = importlib.import_module("audio-analyzer-rs_b200")
x = np.stack([signals.multitone(3, 44100.0, 30000), signals.note_sequence(4, 44100.0, 30000)])
for feats in (2, 6, 15):
    for rep in range(2):
        an = aa.Analyzer(aa.Config(n=2048, sample_rate=44100.0, noise_floor_db=-96.0, features=feats))
        res = an.analyze_host(x, want_dbg=True)
        res2 = an.analyze_host(x, want_dbg=False)
        cfg = O.make_config(2048, 512, 44100.0, features=feats)
        for c in range(2):
            ref = O.analyze_clip(cfg, mags_in=res["mags"][c])
            for name, r in (("dbg", res), ("prod", res2)):
                g, o = r["features"][c], ref["features"]
                for k in ("energy", "flux", "centroid", "max_excess", "burst_count"):
                    rel = np.abs(g[k].astype(np.float64) - o[k]) / np.maximum(np.abs(o[k].astype(np.float64)), 1e-30)
                    bad = np.argwhere(rel > 1e-5).ravel()
                    if len(bad):
                        print(f"features={feats} rep={rep} {name} clip={c} {k}: bad frames {bad[:12].tolist()} n={len(bad)} gpu={g[k][bad[:3]]} ref={o[k][bad[:3]]}")
print("diag done")
P
