"""Small end-to-end case for compute-sanitizer: batch (all sizes), streaming, FFT, notes."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals
aa = importlib.import_module("audio-analyzer-rs_b200")
for n, sr in ((256, 48000.0), (1024, 48000.0), (2048, 44100.0), (4096, 48000.0)):
    clips = np.stack([signals.multitone(i, sr, 6 * n + 3 * (n // 4)) for i in range(3)])
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    r = an.analyze_host(clips, want_dbg=(n == 2048))
    print(n, r["T"], int(r["features"]["n_pitches"].sum()), float(r["mags"].max()))
    f = aa.FftProcessor(n)
    s = f.process_forward(clips[:, :n])
    b = f.process_inverse(s)
    print("   fft", float(np.abs(b / n - clips[:, :n]).max()))
st = aa.Stream(aa.Config(n=1024, sample_rate=48000.0))
x = signals.note_sequence(1, 48000.0, 20000)
tot = 0
for i in range(0, 19 * 1024, 1024):
    st.push(x[i:i + 1024]); tot += len(st.poll())
print("stream frames", tot)
print("notes", aa.notes_from_stable(r["stable"])["n"].sum())
# conditioning chain: the cluster pipeline (two full groups + a ragged one), with and without AGC, carried state
cl = np.stack([signals.note_sequence(40 + i, 48000.0, 1024 * 12, n_notes=3, noise_db=-90.0) for i in range(70)]).astype(np.float32)
for agc in (False, True):
    y, dyn = aa.Conditioner(48000.0, 1024, agc=agc).process_host(cl)
    print("cond agc", agc, float(np.abs(y).max()), dyn.shape)
cc = aa.Conditioner(48000.0, 1024, agc=True, carry=True)
a, _ = cc.process_host(cl[:5, : 1024 * 6])
b, _ = cc.process_host(cl[:5, 1024 * 6:])
print("cond carried", float(np.abs(a).max()), float(np.abs(b).max()))
# time-sliced host pipeline (forced) against the clip-group pipeline
os.environ["AA_HOST_SLICES"] = "3"
an = aa.Analyzer(aa.Config(n=1024, sample_rate=48000.0))
xs = np.stack([signals.multitone(i, 48000.0, 1024 + 256 * 70) for i in range(5)]).astype(np.float32)
r1 = an.analyze_host(xs)
os.environ["AA_HOST_SLICES"] = "0"
r0 = an.analyze_host(xs)
print("sliced == grouped", r1["features"].tobytes() == r0["features"].tobytes(), r1["mags"].tobytes() == r0["mags"].tobytes())
