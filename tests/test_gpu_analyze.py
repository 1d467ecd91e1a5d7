"""GPU parity of the fused frame-analysis kernel (aa_analyze_*) against the oracle.

Two kinds of checks (DESIGN.md "Parity"):
  end-to-end      signal -> GPU  vs  signal -> oracle: spectra within 1e-4 of the frame maximum
                  (north_star tolerance); discrete outputs equal except documented near-ties.
  stage-isolated  the GPU's own magnitudes fed to the oracle's feature stage: the recurrences use
                  the same op-by-op f32 arithmetic, so floors, peak masks, burst counts and
                  max_excess are BIT-EXACT; log-derived floats agree to ~1 ulp; pitch lists are
                  identical unless the oracle itself flags the frame as a near-tie.
"""
import numpy as np
import pytest

import parity
import signals
import util

pytestmark = pytest.mark.gpu


def run_gpu(aa, clips, n, sr, features=15, db=-96.0, onset_in=None, dbg=True):
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr, noise_floor_db=db, features=features))
    res = an.analyze_host(np.atleast_2d(clips), want_dbg=dbg, onset_in=onset_in)
    res["launches"] = an.last_launches
    return res


def check_stage_isolated(O, res, c, n, sr, db=-96.0, features=15, onset_in=None, label=""):
    """Feed the GPU magnitudes of clip c through the oracle feature stage and compare."""
    cfg = O.make_config(n, n // 4, sr, noise_floor_db=db, features=features)
    iso = O.analyze_clip(cfg, mags_in=res["mags"][c], onset_in=onset_in, want_floor=True, want_peaks=True,
                         want_diag=True)
    g, o = res["features"][c], iso["features"]
    T = iso["T"]
    if features & 1:
        assert np.array_equal(res["dbg_floor"][c], iso["floor"]), f"{label}: pitch floor not bit-exact"
        assert np.array_equal(res["dbg_peaks"][c], iso["peaks"]), f"{label}: peak mask differs"
        bad, ties = util.compare_pitch_records(g, o, iso["diag"])
        assert len(bad) == 0, f"{label}: pitch lists differ on non-tie frames {bad[:10]}"
        assert len(ties) <= max(2, 0.05 * T), f"{label}: too many near-tie mismatches {len(ties)}/{T}"
        if features & 8:
            ok_frames = np.ones(T, bool)
            if len(ties):
                ok_frames[ties.min():] = False      # a tie perturbs the tracker from there on
            badst = util.compare_stable(res["stable"][c][ok_frames], iso["stable"][ok_frames])
            assert len(badst) == 0, f"{label}: PitchTracker output differs at {badst[:10]}"
    if features & 2:
        assert np.array_equal(g["burst_count"], o["burst_count"]), f"{label}: burst_count"
        assert np.array_equal(g["max_excess"], o["max_excess"]), f"{label}: max_excess not bit-exact"
        assert util.ulp_close(g["energy"], o["energy"], 1e-5).all(), f"{label}: energy"
        assert util.ulp_close(g["flux"], o["flux"], 2e-5, 1e-6).all(), f"{label}: flux"
        assert util.ulp_close(g["energy_ema"], o["energy_ema"], 1e-5).all(), f"{label}: ema"
        nflag = int((g["flags"] != o["flags"]).sum())
        assert nflag <= max(1, 0.01 * T), f"{label}: {nflag} onset-flag mismatches"
    if features & 4:
        assert util.ulp_close(g["centroid"], o["centroid"], 1e-5, 1e-3).all(), f"{label}: centroid"
    return iso


def check_end_to_end(O, res, c, n, sr, x=None, mags_a=None, db=-96.0, features=15, onset_in=None, label=""):
    """signal -> GPU against signal -> oracle (or against the golden float64 magnitudes `mags_a`), strictly:
    every differing peak bit, burst bit and pitch-bin list must be a magnitude-level near-tie or lie downstream
    of one through the recurrent per-bin state (tests/parity.py) -- ZERO unexplained differences.
    Run B of the explanation is the oracle's feature stage on the GPU's own magnitudes, and the GPU's
    records are first shown to BE run B (peak masks / burst counts bit-exact, pitch lists equal except the
    oracle-flagged 1-ulp log near-ties), so A vs B is signal -> oracle vs signal -> GPU."""
    cfg = O.make_config(n, n // 4, sr, noise_floor_db=db, features=features)
    kw = dict(onset_in=onset_in, want_floor=True, want_peaks=True, want_diag=True, want_state=True)
    A = O.analyze_clip(cfg, x, **kw) if mags_a is None else O.analyze_clip(cfg, mags_in=mags_a, **kw)
    B = O.analyze_clip(cfg, mags_in=res["mags"][c], **kw)
    g = res["features"][c]
    if features & 2:
        assert np.array_equal(g["burst_count"], B["features"]["burst_count"]), f"{label}: burst_count (stage-isolated)"
    if features & 1:
        if res.get("dbg_peaks") is not None:
            assert np.array_equal(res["dbg_peaks"][c], B["peaks"]), f"{label}: peak mask (stage-isolated)"
        bad, _ = util.compare_pitch_records(g, B["features"], B["diag"])
        assert len(bad) == 0, f"{label}: pitch lists differ from the oracle on the GPU's own magnitudes at {bad[:10]}"
    rep = parity.explain(O, cfg, A, B, label)
    print(parity.summarize(rep))
    assert rep["unexplained"] == 0, (parity.summarize(rep), rep["detail"])
    assert rep["d_rel_max"] <= 2e-6, rep["d_rel_max"]
    return rep


@pytest.mark.parametrize("path", util.golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fixtures(aa, O, torch_cuda, path):
    g = util.load_golden(path)
    res = run_gpu(aa, g["samples"], g["n"], g["sr"], db=g["db"], onset_in=g["onset_in"])
    assert res["T"] == g["mags"].shape[0]
    err = util.mag_err(res["mags"][0], g["mags"])
    assert err.max() <= util.MAG_TOL, err.max()
    assert err.max() <= 2e-6            # what the kernel actually achieves against float64
    check_stage_isolated(O, res, 0, g["n"], g["sr"], g["db"], onset_in=g["onset_in"], label=path[-24:])
    # end to end against the golden records (the oracle on the golden float64 magnitudes reproduces them bit for
    # bit, tests/test_oracle.py): every difference must be an explained near-tie
    check_end_to_end(O, res, 0, g["n"], g["sr"], mags_a=g["mags"], db=g["db"], onset_in=g["onset_in"],
                     label="golden " + path[-24:-4])
    check_end_to_end(O, res, 0, g["n"], g["sr"], x=g["samples"], db=g["db"], onset_in=g["onset_in"],
                     label="oracle " + path[-24:-4])


def test_cfg1_sine_440(aa, O, torch_cuda):
    """BASELINE configs[0]: 440 Hz sine, 44.1 kHz mono 10 s, 2048-pt Hann STFT hop 512 + pitch."""
    x = signals.sine(440.0, 44100.0, 441000)
    res = run_gpu(aa, x, 2048, 44100.0)
    assert res["T"] == 858
    ref = O.analyze_clip(O.make_config(2048, 512, 44100.0), x)
    assert util.mag_err(res["mags"][0], ref["mags"]).max() <= util.MAG_TOL
    assert (res["mags"][0].argmax(axis=1) == 20).all()
    f = res["features"][0]
    assert (f["n_pitches"] == 1).all()
    assert np.allclose(f["pitch"]["freq"][:, 0], 440.196, atol=2e-3)
    assert np.allclose(f["pitch"]["score"][:300, 0], 0.522, atol=1e-3)
    st = res["stable"][0]
    assert st["n"][0] == 0 and (st["n"][1:] == 1).all()
    check_stage_isolated(O, res, 0, 2048, 44100.0, label="cfg1")
    rep = check_end_to_end(O, res, 0, 2048, 44100.0, x=x, label="cfg1")
    # the raw pitch BIN lists of the two runs are identical on every one of the 858 frames
    assert rep["pitch_lists_differ"] == 0


@pytest.mark.parametrize("n,sr", [(256, 48000.0), (512, 22050.0), (1024, 48000.0), (2048, 44100.0),
                                  (4096, 48000.0), (4096, 16000.0)])
def test_all_window_sizes_multiclip(aa, O, torch_cuda, n, sr):
    clips = np.stack([signals.multitone(100 + i, sr, 10 * n + 3 * (n // 4)) for i in range(5)]
                     + [signals.note_sequence(7, sr, 10 * n + 3 * (n // 4))])
    res = run_gpu(aa, clips, n, sr)
    cfg = O.make_config(n, n // 4, sr)
    for c in range(clips.shape[0]):
        ref = O.analyze_clip(cfg, clips[c])
        assert util.mag_err(res["mags"][c], ref["mags"]).max() <= 2e-6
        check_stage_isolated(O, res, c, n, sr, label=f"n={n} clip {c}")
        check_end_to_end(O, res, c, n, sr, x=clips[c], label=f"n={n} clip {c}")


@pytest.mark.parametrize("features", [0, 1, 2, 4, 1 | 8, 2 | 4, 1 | 2, 15])
def test_feature_subsets(aa, O, torch_cuda, features):
    x = np.stack([signals.multitone(3, 44100.0, 30000), signals.note_sequence(4, 44100.0, 30000)])
    res = run_gpu(aa, x, 2048, 44100.0, features=features)
    full = run_gpu(aa, x, 2048, 44100.0, features=15)
    assert np.array_equal(res["mags"], full["mags"])
    cfg = O.make_config(2048, 512, 44100.0, features=features)
    for c in range(2):
        check_stage_isolated(O, res, c, 2048, 44100.0, features=features, label=f"features={features}")
        ref = O.analyze_clip(cfg, mags_in=res["mags"][c])
        if not features & 1:
            assert (res["features"][c]["n_pitches"] == 0).all() and (res["stable"][c]["n"] == 0).all()
        if not features & 2:
            assert (res["features"][c]["flux"] == 0).all() and (res["features"][c]["flags"] == 0).all()
        if not features & 4:
            assert (res["features"][c]["centroid"] == 0).all()
        if not features & 8:
            assert (res["stable"][c]["n"] == 0).all()


def test_edge_lengths_and_silence(aa, O, torch_cuda):
    n, sr = 1024, 48000.0
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    # shorter than one window -> zero frames, no launch
    r = an.analyze_host(np.zeros((3, n - 4), np.float32))
    assert r["T"] == 0 and an.last_launches == 0
    # exactly one window / ragged tails that do not fill a hop
    for extra in (0, 4, 252, 256, 260):
        x = signals.multitone(9, sr, n + extra)[None, :]
        r = an.analyze_host(x)
        assert r["T"] == 1 + extra // 256
        ref = O.analyze_clip(O.make_config(n, n // 4, sr), x[0])
        assert util.mag_err(r["mags"][0], ref["mags"]).max() <= 2e-6
    # digital silence and DC
    r = an.analyze_host(np.zeros((2, 8192), np.float32))
    assert (r["mags"] == 0).all() and (r["features"]["n_pitches"] == 0).all()
    assert (r["features"]["flux"] == 0).all() and (r["features"]["burst_count"] == 0).all()
    r = an.analyze_host(np.full((1, 8192), 0.25, np.float32))
    assert (r["features"]["n_pitches"] == 0).all()
    # full-scale square wave (maximum magnitudes), clipping-level input
    sq = np.sign(np.sin(2 * np.pi * 1000.0 * np.arange(16384) / sr)).astype(np.float32)[None, :]
    r = an.analyze_host(sq, want_dbg=True)
    check_stage_isolated(O, r, 0, n, sr, label="square")


def test_candidate_overflow_and_general_selection_paths(aa, O, torch_cuda):
    """More than 256 scoring candidates per frame (the shared-memory candidate list spills to the HBM
    scratch) and more than 32 survivors of the cutoff (general selection path instead of the
    lane-resident one): a comb of 300 equal tones with every bin inside the pitch range."""
    n, sr = 4096, 16000.0
    rng = np.random.default_rng(42)
    t = np.arange(n + 12 * (n // 4), dtype=np.float64)
    x = np.zeros_like(t)
    for j in range(300):
        b = 20 + 6 * j + rng.uniform(-0.3, 0.3)
        x += 0.003 * np.sin(2 * np.pi * b * t / n + rng.uniform(0, 2 * np.pi))
    x = x.astype(np.float32)[None, :]
    res = run_gpu(aa, x, n, sr)
    iso = check_stage_isolated(O, res, 0, n, sr, label="overflow")
    assert iso["diag"]["n_scored"].max() > 256, iso["diag"]["n_scored"].max()
    assert (res["features"][0]["n_pitches"] == 8).all()
    # white noise loud enough that most local maxima are candidates
    y = (0.2 * rng.standard_normal(x.shape[1])).astype(np.float32)[None, :]
    res = run_gpu(aa, y, n, sr)
    iso = check_stage_isolated(O, res, 0, n, sr, label="noise")
    assert iso["diag"]["n_scored"].max() > 256 and iso["diag"]["n_candidates"].max() > 32


def test_noise_floor_db_and_freq_range_parameters(aa, O, torch_cuda):
    x = signals.multitone(21, 44100.0, 40000)[None, :]
    for db, fmin, fmax in [(-60.0, 24.0, 10000.0), (-96.0, 200.0, 2000.0), (-30.0, 24.0, 10000.0)]:
        an = aa.Analyzer(aa.Config(n=2048, sample_rate=44100.0, noise_floor_db=db, min_freq=fmin, max_freq=fmax))
        res = an.analyze_host(x, want_dbg=True)
        cfg = O.make_config(2048, 512, 44100.0, fmin, fmax, db)
        iso = O.analyze_clip(cfg, mags_in=res["mags"][0], want_floor=True, want_peaks=True, want_diag=True)
        assert np.array_equal(res["dbg_floor"][0], iso["floor"])
        assert np.array_equal(res["dbg_peaks"][0], iso["peaks"])
        bad, _ = util.compare_pitch_records(res["features"][0], iso["features"], iso["diag"])
        assert len(bad) == 0


@pytest.mark.parametrize("n,sr", [(4096, 48000.0), (2048, 44100.0), (1024, 48000.0), (256, 48000.0)])
def test_production_kernel_variants_equal_the_tap_build(aa, O, torch_cuda, n, sr):
    """The parity checks above run the instantiation with the debug taps, which keeps every bin's floor state.
    The production instantiations drop the pitch-floor recurrence of bin groups above max_bin (dead state) and
    are templated on how many group slots per warp can be live: max_freq picks the variant (few live slots /
    all slots).  Every record must be byte-identical to the tap build, and the tap build is checked against
    the oracle for the same parameters."""
    clips = np.stack([signals.multitone(300 + i, sr, 24 * n) for i in range(5)])
    clips[4] *= 0.0
    for fmax in (900.0, 10000.0, 20000.0):
        cfg = aa.Config(n=n, sample_rate=sr, max_freq=fmax)
        tap = aa.Analyzer(cfg).analyze_host(clips, want_dbg=True)
        prod = aa.Analyzer(cfg).analyze_host(clips, want_dbg=False)
        for k in ("features", "stable", "mags", "summaries"):
            assert tap[k].tobytes() == prod[k].tobytes(), f"n={n} max_freq={fmax}: {k} differs"
        ocfg = O.make_config(n, n // 4, sr, 24.0, fmax, -96.0)
        iso = O.analyze_clip(ocfg, mags_in=tap["mags"][0], want_floor=True, want_peaks=True, want_diag=True)
        assert np.array_equal(tap["dbg_floor"][0], iso["floor"]) and np.array_equal(tap["dbg_peaks"][0], iso["peaks"])
        bad, _ = util.compare_pitch_records(prod["features"][0], iso["features"], iso["diag"])
        assert len(bad) == 0


def test_onset_in_drives_the_tracker(aa, O, torch_cuda):
    x = signals.multitone(33, 44100.0, 60000)[None, :]
    T = (60000 - 2048) // 512 + 1
    onset = np.zeros((1, T), np.uint8)
    onset[0, [20, 21, 50]] = 1
    res = run_gpu(aa, x, 2048, 44100.0, onset_in=onset)
    check_stage_isolated(O, res, 0, 2048, 44100.0, onset_in=onset[0], label="onset_in")
    res0 = run_gpu(aa, x, 2048, 44100.0)
    assert res["stable"].tobytes() != res0["stable"].tobytes()


def test_chunked_stream_with_halo(aa, O, torch_cuda):
    """cfg4 shape: hop-aligned chunks with a one-window halo, expressed as overlapping clips.
    Stateless outputs equal the unchunked run bit for bit; stateful ones restart per chunk (see the exact and
    warm-up modes below)."""
    import importlib

    sh = importlib.import_module("audio-analyzer-rs_b200.sharding")
    n, hop, sr = 2048, 512, 48000.0
    x = signals.chord_vibrato(0xA0D14, sr, 400000)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    full = an.analyze_host(x[None, :])
    nch, clen, stride, tail = sh.uniform_chunks(len(x), n, hop, 96)
    ch = an.analyze_host(x, clip_len=clen, clip_stride=stride)
    assert ch["T"] == 96 and ch["mags"].shape[0] == nch
    Tc = nch * 96
    assert np.array_equal(ch["mags"].reshape(Tc, -1), full["mags"][0, :Tc])
    assert np.array_equal(ch["features"]["energy"].reshape(-1), full["features"]["energy"][0, :Tc])
    assert np.array_equal(ch["features"]["centroid"].reshape(-1), full["features"]["centroid"][0, :Tc])
    # the first chunk is the start of the stream: everything equal there
    assert ch["features"][0].tobytes() == full["features"][0, :96].tobytes()


def test_chunked_stream_exact_mode_is_byte_identical(aa, O, torch_cuda):
    """cfg4 exact mode ("stateful features need serial state hand-off"): consecutive chunks alternate
    between TWO analyzer handles (the two ranks of a 2-GPU run; tools/run_multigpu_cases.py does it over NCCL) and
    the analyzer state block -- floors, volatility, previous magnitudes, FluxTracker / EMA scalars, PitchTracker
    tracks, what stft.rs:209-212 / onset.rs:149-200 keep between frames -- travels with the stream as one message.
    Every output byte equals the unchunked run, with onset_in driving the tracker across chunk boundaries."""
    import importlib

    torch = torch_cuda
    sh = importlib.import_module("audio-analyzer-rs_b200.sharding")
    for n, sr, total, n_chunks in ((2048, 48000.0, 400000, 7), (4096, 48000.0, 300000, 5), (256, 48000.0, 40000, 9)):
        hop, half = n // 4, n // 2 + 1
        x = signals.chord_vibrato(0xA0D14, sr, total) if n != 256 else signals.note_sequence(3, sr, total)
        T = sh.total_frames(total, n, hop)
        onset = (np.arange(T) % 41 == 7).astype(np.uint8)
        cfg = aa.Config(n=n, sample_rate=sr)
        full = aa.Analyzer(cfg).analyze_host(x[None, :], onset_in=onset[None, :])
        handles = [aa.Analyzer(cfg), aa.Analyzer(cfg)]
        xd = torch.from_numpy(x).cuda()
        od = torch.from_numpy(onset).cuda()
        feat = torch.zeros(T, 96, device="cuda", dtype=torch.uint8)
        stab = torch.zeros(T, 136, device="cuda", dtype=torch.uint8)
        mags = torch.zeros(T, half, device="cuda", dtype=torch.float32)
        nfl = handles[0].state_floats
        assert nfl == 4 * half + 104
        states = [torch.zeros(nfl, device="cuda"), torch.zeros(nfl, device="cuda")]    # one block per "rank"
        plan = sh.chunk_plan(total, n, hop, n_chunks)
        s = torch.cuda.current_stream().cuda_stream

        def run_chunk(_stream, c, state):
            ch = plan[c]
            handles[c % 2].analyze_device_carry(
                xd.data_ptr() + 4 * ch.start, 1, ch.length, (ch.length + 3) & ~3, state.data_ptr(),
                mags=mags.data_ptr() + 4 * half * ch.first_frame, features=feat.data_ptr() + 96 * ch.first_frame,
                stable=stab.data_ptr() + 136 * ch.first_frame, onset_in=od.data_ptr() + ch.first_frame, stream=s)

        # the two "ranks" interleaved in wavefront order; the hand-off is a copy from one rank's block to the other's
        for c in range(len(plan)):
            if c:
                states[c % 2].copy_(states[(c - 1) % 2])
            run_chunk(0, c, states[c % 2])
        torch.cuda.synchronize()
        assert feat.cpu().numpy().tobytes() == full["features"][0].tobytes(), f"n={n}: features differ"
        assert stab.cpu().numpy().tobytes() == full["stable"][0].tobytes(), f"n={n}: stable pitches differ"
        assert np.array_equal(mags.cpu().numpy(), full["mags"][0]), f"n={n}: magnitudes differ"
        # the same through the generic driver on one rank (states never leave the device)
        feat2 = torch.zeros_like(feat)
        mags_keep, feat_keep = mags, feat
        feat = feat2
        log = sh.chain_chunks(1, len(plan), 0, 1, run_chunk, lambda: torch.zeros(nfl, device="cuda"))
        torch.cuda.synchronize()
        assert len(log) == len(plan) and feat2.cpu().numpy().tobytes() == full["features"][0].tobytes()


def test_chunked_stream_warmup_mode_converges(aa, O, torch_cuda):
    """cfg4 warm-up mode: chunks start `warmup_frames` early with a fresh analyzer and drop those frames.  Not
    exact by construction; the mismatch against the unchunked run must shrink with the warm-up length and the
    stateless outputs stay bit-identical (the table for the 1-hour stream is in DESIGN.md 4)."""
    import importlib

    sh = importlib.import_module("audio-analyzer-rs_b200.sharding")
    n, hop, sr = 2048, 512, 48000.0
    x = signals.chord_vibrato(0xA0D14, sr, 600000)
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    full = an.analyze_host(x[None, :], want_mags=False)
    ff = full["features"][0]
    mism = {}
    for W in (0, 16, 128, 400):
        feats = []
        for ch in sh.chunk_plan(len(x), n, hop, 6, warmup_frames=W):
            r = an.analyze_host(x[None, ch.start: ch.start + ch.length], want_mags=False)
            assert r["T"] == ch.n_frames + ch.warmup
            feats.append(r["features"][0][ch.warmup:])
        f = np.concatenate(feats)
        assert np.array_equal(f["energy"], ff["energy"]) and np.array_equal(f["centroid"], ff["centroid"])
        bad, _ = util.compare_pitch_records(f, ff)
        mism[W] = (len(bad), int((f["burst_count"] != ff["burst_count"]).sum()))
    print("warm-up mismatches (pitch-list frames, burst-count frames) of", len(ff), "frames:", mism)
    assert mism[400][0] <= mism[0][0] and mism[400][1] <= mism[0][1]
    assert mism[400][1] <= 0.01 * len(ff)


def test_device_api_many_clips_and_summaries(aa, O, torch_cuda):
    """Device-resident path: synthetic clips generated on the GPU (SURVEY 8d generator), more clips
    than SM slots so several waves run; a sample of clips is checked against the oracle and the
    per-clip summaries against numpy."""
    torch = torch_cuda
    n, sr, clip_len, n_clips = 2048, 44100.0, 44100, 700
    cfg = aa.Config(n=n, sample_rate=sr)
    an = aa.Analyzer(cfg)
    T = an.num_frames(clip_len)
    half = n // 2 + 1
    clips = torch.empty(n_clips, clip_len, device="cuda", dtype=torch.float32)
    aa.synth_clips_device(clips.data_ptr(), n_clips, clip_len, clip_len, sr, 0xA0D15)
    mags = torch.empty(n_clips, T, half, device="cuda", dtype=torch.float32)
    feat = torch.zeros(n_clips, T, 96, device="cuda", dtype=torch.uint8)
    stab = torch.zeros(n_clips, T, 136, device="cuda", dtype=torch.uint8)
    summ = torch.zeros(n_clips, 32, device="cuda", dtype=torch.uint8)
    s = torch.cuda.current_stream()
    an.analyze_device(clips.data_ptr(), n_clips, clip_len, clip_len, mags=mags.data_ptr(),
                      features=feat.data_ptr(), stable=stab.data_ptr(), summaries=summ.data_ptr(),
                      stream=s.cuda_stream)
    torch.cuda.synchronize()
    assert an.last_launches == 2
    h_clips = clips.cpu().numpy()
    assert np.isfinite(h_clips).all() and 0.01 < np.abs(h_clips).max() <= 0.51
    h_feat = feat.cpu().numpy().view(aa.FEATURES_DTYPE).reshape(n_clips, T)
    h_stab = stab.cpu().numpy().view(aa.STABLE_DTYPE).reshape(n_clips, T)
    h_summ = summ.cpu().numpy().view(aa.SUMMARY_DTYPE).reshape(n_clips)
    ocfg = O.make_config(n, n // 4, sr)
    for c in (0, 1, 147, 148, 333, 699):
        ref = O.analyze_clip(ocfg, h_clips[c])
        gm = mags[c].cpu().numpy()
        assert util.mag_err(gm, ref["mags"]).max() <= 2e-6
        iso = O.analyze_clip(ocfg, mags_in=gm, want_diag=True)
        bad, _ = util.compare_pitch_records(h_feat[c], iso["features"], iso["diag"])
        assert len(bad) == 0
        assert np.array_equal(h_feat[c]["burst_count"], iso["features"]["burst_count"])
    # summaries
    assert (h_summ["n_frames"] == T).all()
    assert np.array_equal(h_summ["n_pitched"], (h_feat["n_pitches"] > 0).sum(axis=1))
    assert np.array_equal(h_summ["n_onsets"], ((h_feat["flags"] & 4) != 0).sum(axis=1))
    assert np.allclose(h_summ["mean_centroid"], h_feat["centroid"].astype(np.float64).mean(axis=1), rtol=1e-6)
    assert np.allclose(h_summ["mean_energy"], h_feat["energy"].astype(np.float64).mean(axis=1), rtol=1e-6)
    assert np.array_equal(h_summ["max_energy"], h_feat["energy"].max(axis=1))
    # determinism: same input twice -> identical bytes
    feat2 = torch.zeros_like(feat)
    an.analyze_device(clips.data_ptr(), n_clips, clip_len, clip_len, features=feat2.data_ptr(),
                      stream=s.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(feat, feat2)
    # host path == device path
    hres = an.analyze_host(h_clips[:300], want_mags=False)
    assert hres["features"].tobytes() == h_feat[:300].tobytes()
    assert hres["stable"].tobytes() == h_stab[:300].tobytes()
    assert hres["summaries"].tobytes() == h_summ[:300].tobytes()


def test_full_size_cfg2_properties(aa, O, torch_cuda):
    """BASELINE configs[1] at full size: 1024 clips x 30 s @ 48 kHz, 4096-pt / hop 1024, all features.
    Size-independent properties: duplicated clips give identical records (independent of the CTA /
    SM they ran on), three sampled clips match the oracle, summaries are consistent."""
    torch = torch_cuda
    n, sr, clip_len, n_clips = 4096, 48000.0, 1440000, 1024
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    T = an.num_frames(clip_len)
    assert T == 1403
    clips = torch.empty(n_clips, clip_len, device="cuda", dtype=torch.float32)
    aa.synth_clips_device(clips.data_ptr(), n_clips // 2, clip_len, clip_len, sr, 0xA0D10)
    clips[n_clips // 2:] = clips[: n_clips // 2]          # second half duplicates the first
    feat = torch.zeros(n_clips, T, 96, device="cuda", dtype=torch.uint8)
    summ = torch.zeros(n_clips, 32, device="cuda", dtype=torch.uint8)
    an.analyze_device(clips.data_ptr(), n_clips, clip_len, clip_len, features=feat.data_ptr(),
                      summaries=summ.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(feat[: n_clips // 2], feat[n_clips // 2:])
    assert torch.equal(summ[: n_clips // 2], summ[n_clips // 2:])
    h_feat = feat.cpu().numpy().view(aa.FEATURES_DTYPE).reshape(n_clips, T)
    assert (h_feat["n_pitches"] > 0).mean() > 0.9          # every clip is tonal
    # three sampled clips, strictly: re-run them alone with the magnitude / floor / peak taps (records must be
    # byte-identical to what the full batch produced for them), then explain every end-to-end difference
    sample = [0, 311, 1023]
    xs = np.stack([clips[c].cpu().numpy() for c in sample])
    tap = aa.Analyzer(aa.Config(n=n, sample_rate=sr)).analyze_host(xs, want_dbg=True)
    for i, c in enumerate(sample):
        assert tap["features"][i].tobytes() == h_feat[c].tobytes(), f"clip {c}: batch records != single-clip records"
        check_stage_isolated(O, tap, i, n, sr, label=f"cfg2 clip {c}")
        check_end_to_end(O, tap, i, n, sr, x=xs[i], label=f"cfg2 clip {c}")


def test_time_segments_equal_whole_clips(aa, O, torch_cuda):
    """Batches with more clips than resident CTAs are cut into time segments that hand the analyzer state
    (per-bin floors, previous magnitudes, FluxTracker / EMA scalars, PitchTracker tracks) through HBM.
    Every output must be byte-identical to the same clips analysed whole: sub-batches of at most one clip
    per resident CTA take the static, unsegmented path."""
    torch = torch_cuda
    for n, sr, n_clips, clip_len, n_distinct in ((1024, 48000.0, 1200, 1024 + 256 * 299, 5),
                                                 (4096, 48000.0, 500, 4096 + 1024 * 200, 4)):
        an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
        T = an.num_frames(clip_len)
        half = n // 2 + 1
        base = np.stack([signals.multitone(40 + c, sr, clip_len) for c in range(n_distinct)])
        clips = torch.from_numpy(base).cuda()[torch.arange(n_clips, device="cuda") % n_distinct].contiguous()
        onset = (torch.arange(n_clips * T, device="cuda") % 37 == 5).to(torch.uint8)   # exercises the tracker's onset branch

        def run(first, count):
            feat = torch.zeros(count, T, 96, device="cuda", dtype=torch.uint8)
            stab = torch.zeros(count, T, 136, device="cuda", dtype=torch.uint8)
            mags = torch.zeros(count, T, half, device="cuda", dtype=torch.float32)
            an.analyze_device(clips[first:first + count].data_ptr(), count, clip_len, clip_len,
                              onset_in=onset[first * T:].data_ptr(), features=feat.data_ptr(),
                              stable=stab.data_ptr(), mags=mags.data_ptr(),
                              stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            return feat, stab, mags

        feat, stab, mags = run(0, n_clips)                 # queue + segments (n_clips > resident CTAs)
        sub = 148                                          # <= one clip per CTA on any B200 configuration
        for first in range(0, n_clips, sub):
            cnt = min(sub, n_clips - first)
            f2, s2, m2 = run(first, cnt)
            assert torch.equal(feat[first:first + cnt], f2), (n, first)
            assert torch.equal(stab[first:first + cnt], s2), (n, first)
            assert torch.equal(mags[first:first + cnt], m2), (n, first)
        # and the whole-clip path agrees with the oracle on one clip (stage-isolated, pitch counts)
        h_feat = feat[1].cpu().numpy().view(aa.FEATURES_DTYPE).reshape(T)
        iso = O.analyze_clip(O.make_config(n, n // 4, sr), mags_in=mags[1].cpu().numpy(),
                             onset_in=onset[T:2 * T].cpu().numpy())
        assert (iso["features"]["burst_count"] == h_feat["burst_count"]).all()
        iso = O.analyze_clip(O.make_config(n, n // 4, sr), mags_in=mags[1].cpu().numpy(),
                             onset_in=onset[T:2 * T].cpu().numpy(), want_diag=True)
        bad, ties = util.compare_pitch_records(h_feat, iso["features"], iso["diag"])
        assert len(bad) == 0 and len(ties) <= 0.05 * T, (bad[:10], len(ties))
        # ... and end to end (signal -> oracle) with every difference explained
        res1 = {"mags": mags[1:2].cpu().numpy(), "features": h_feat[None, :]}
        check_end_to_end(O, res1, 0, n, sr, x=base[1], onset_in=onset[T:2 * T].cpu().numpy(), label=f"segments n={n}")


def test_time_sliced_host_pipeline_equals_clip_groups(aa, O, torch_cuda, monkeypatch):
    """aa_analyze_host cuts large batches in TIME (every clip's frames [f0, f1) per slice, analyzer state carried from
    slice to slice in HBM, records written at their whole-clip positions) so that the copies hide behind the
    kernels; small batches go clip group by clip group.  AA_HOST_SLICES forces either path: every output byte
    must be the same, for float32 and interleaved 16-bit stereo input, with onset_in, for slice counts that do and
    do not divide the frame count, and with more clips than one wave of resident CTAs."""
    for n, sr, n_clips, length, slices in ((2048, 44100.0, 6, 2048 + 512 * 100 + 40, 3), (4096, 48000.0, 3, 120000, 7),
                                           (256, 48000.0, 1500, 256 + 64 * 63, 4)):
        rng = np.random.default_rng(n)
        if n_clips <= 8:
            x = np.stack([signals.note_sequence(100 + c, sr, length) for c in range(n_clips)]).astype(np.float32)
        else:
            base = np.stack([signals.multitone(200 + c, sr, length) for c in range(8)]).astype(np.float32)
            x = base[rng.integers(0, 8, n_clips)] * rng.uniform(0.1, 1.0, (n_clips, 1)).astype(np.float32)
        an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
        T = an.num_frames(length)
        onset = (rng.random((n_clips, T)) < 0.03).astype(np.uint8)
        monkeypatch.setenv("AA_HOST_SLICES", "0")
        ref = an.analyze_host(x, onset_in=onset)
        monkeypatch.setenv("AA_HOST_SLICES", str(slices))
        got = an.analyze_host(x, onset_in=onset)
        for k in ("mags", "features", "stable", "summaries"):
            assert got[k].tobytes() == ref[k].tobytes(), f"n={n}: {k} differs between the sliced and the grouped pipeline"
        # records only (no magnitudes), the shape the headline bench uses
        got = an.analyze_host(x, onset_in=onset, want_mags=False, want_stable=False)
        assert got["features"].tobytes() == ref["features"].tobytes()
        assert got["summaries"].tobytes() == ref["summaries"].tobytes()
        # 16-bit interleaved stereo (down-mixed on the device, reference io/mod.rs)
        if n_clips <= 8:
            L4 = length - length % 4
            pcm = np.empty((n_clips, L4, 2), np.int16)
            pcm[..., 0] = np.round(x[:, :L4] * 20000.0)
            pcm[..., 1] = np.round(x[:, :L4] * 9000.0)
            pcm = pcm.reshape(n_clips, 2 * L4)
            monkeypatch.setenv("AA_HOST_SLICES", "0")
            ref = an.analyze_host_pcm(pcm, aa.PCM_I16, 2)
            monkeypatch.setenv("AA_HOST_SLICES", str(slices))
            got = an.analyze_host_pcm(pcm, aa.PCM_I16, 2)
            for k in ("mags", "features", "stable", "summaries"):
                assert got[k].tobytes() == ref[k].tobytes(), f"n={n} pcm16x2: {k} differs"
    monkeypatch.delenv("AA_HOST_SLICES")


def test_state_carry_with_more_streams_than_resident_ctas(aa, O, torch_cuda):
    """aa_analyze_device_carry with thousands of streams in one call (the clips are dealt from the device queue):
    two chained halves equal the whole-clip launch, byte for byte."""
    torch = torch_cuda
    n, sr, n_clips, T = 512, 22050.0, 2600, 48
    hop, half = n // 4, n // 2 + 1
    length = (T - 1) * hop + n
    rng = np.random.default_rng(5)
    base = np.stack([signals.note_sequence(300 + c, sr, length) for c in range(8)]).astype(np.float32)
    x = torch.from_numpy(base[rng.integers(0, 8, n_clips)] * rng.uniform(0.1, 1.0, (n_clips, 1)).astype(np.float32)).cuda()
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    s = torch.cuda.current_stream().cuda_stream
    ref = torch.zeros(n_clips, T, 96, device="cuda", dtype=torch.uint8)
    an.analyze_device(x.data_ptr(), n_clips, length, length, features=ref.data_ptr(), stream=s)
    state = torch.zeros(n_clips, an.state_floats, device="cuda")
    parts = []
    for f0, f1 in ((0, 20), (20, T)):
        out = torch.zeros(n_clips, f1 - f0, 96, device="cuda", dtype=torch.uint8)
        an.analyze_device_carry(x.data_ptr() + 4 * f0 * hop, n_clips, (f1 - f0 - 1) * hop + n, length, state.data_ptr(),
                                features=out.data_ptr(), stream=s)
        parts.append(out)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(parts, dim=1), ref)


def test_whole_clip_queue_path(torch_cuda):
    """Batches of 16 or more clips per resident CTA (and AA_SEG_MIN=0) deal whole clips from the device-wide queue.
    The segment planner reads AA_SEG_MIN once per process, so the comparison of the queue path with the static
    path runs in a child process with segmentation switched off."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AA_SEG_MIN="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "-m", "gpu",
                        "tests/test_gpu_analyze.py::test_time_segments_equal_whole_clips"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_note_records_match_reference_from_freq(aa, O, torch_cuda):
    """NEXT row f2: Note::from_freq (theory.rs:195-209) on the device for every stable pitch."""
    x = np.stack([signals.sine(440.0, 44100.0, 30000), signals.multitone(3, 44100.0, 30000)])
    res = run_gpu(aa, x, 2048, 44100.0, dbg=False)
    notes = aa.notes_from_stable(res["stable"])
    assert notes.shape == res["stable"].shape
    assert np.array_equal(notes["n"], res["stable"]["n"])
    a4 = notes[0][5]
    assert a4["n"] == 1 and aa.NOTE_NAMES[a4["note"][0]["semis"]] == "A" and a4["note"][0]["octave"] == 4
    assert abs(a4["note"][0]["cents"] - 0.77) < 0.05        # 440.196 Hz is +0.77 cents
    checked = 0
    for c in range(2):
        for t in range(res["T"]):
            for i in range(int(res["stable"][c][t]["n"])):
                f = float(res["stable"][c][t]["pitch"][i]["freq"])
                _, octave, semis, cents = O.note_from_freq(f)
                g = notes[c][t]["note"][i]
                if abs(abs(cents) - 50.0) < 0.01:
                    continue                                  # semitone boundary: log2 ulp near-tie
                assert (int(g["octave"]), int(g["semis"])) == (octave, semis)
                assert abs(float(g["cents"]) - cents) < 2e-3  # 1 ulp of log2 * 1200 at ~5000 cents
                checked += 1
    assert checked > 100
    # unused slots are zero
    assert (notes["note"]["cents"][res["stable"]["n"] == 0] == 0).all()
    # other base frequency
    n2 = aa.notes_from_stable(res["stable"][0:1, 5:6], base_freq=415.0)
    assert aa.NOTE_NAMES[n2[0, 0]["note"][0]["semis"]] == "A#"


def test_yin_lag_search_matches_oracle(aa, O, torch_cuda):
    """a14 (NEW, self-defined): YIN-style lag search on the GPU vs the float64 oracle: lag indices equal
    except near-ties (flat d' around the threshold / the minimum), d'(lag) within 1e-4."""
    sr, n, hop = 48000.0, 2048, 512
    rng = np.random.default_rng(3)
    clips = np.stack([signals.sine(f0, sr, 20000, 0.5) + 0.3 * signals.sine(2 * f0, sr, 20000, 0.5, 1.0)
                      for f0 in (110.0, 220.0, 329.6, 441.0, 987.8)]
                     + [signals.multitone(8, sr, 20000), (0.1 * rng.standard_normal(20000)).astype(np.float32)])
    min_lag, max_lag = 24, 1024
    lag, cm = aa.yin_host(clips, n, hop, min_lag, max_lag, 0.1)
    T = (20000 - n) // hop + 1
    assert lag.shape == (len(clips), T)
    mism = 0
    for c in range(len(clips)):
        for t in range(0, T, 3):
            fr = clips[c, t * hop: t * hop + n]
            want, dp = O.yin_lag(fr, min_lag, max_lag, 0.1)
            got = int(lag[c, t])
            if got != want:
                # a near-tie: the oracle's d' at the two lags (or at the threshold) is within 1e-4
                assert abs(dp[got] - dp[want]) < 2e-4 or abs(dp[min(got, want)] - 0.1) < 2e-4, (c, t, got, want)
                mism += 1
            else:
                assert abs(float(cm[c, t]) - dp[want]) < 1e-4 + 1e-3 * dp[want]
    assert mism <= 3
    # pure tones: lag == round(sr / f0)
    assert abs(int(lag[0, 5]) - round(sr / 110.0)) <= 1 and abs(int(lag[3, 5]) - round(sr / 441.0)) <= 1
    with pytest.raises(aa.AAError):
        aa.yin_host(clips, n, hop, 0, 1024)


def test_tuner_records_match_the_reference_branches(aa, O, torch_cuda):
    """aa_tuner_from_stable_* vs the oracle restatement of Tuner::run / Interval::new (tuner.rs:148-193,
    theory.rs:306-382; the oracle replays the reference's own Interval tests): kind, indices and interval type
    are exact; accuracy / cents go through logf / log2f (<= 1 ulp apart): 1e-4 absolute cents."""
    rng = np.random.default_rng(11)
    n = 4000
    st = np.zeros(n, aa.STABLE_DTYPE)
    st["n"] = rng.integers(0, 5, n)
    st["n"][:50] = 2
    f = np.exp(rng.uniform(np.log(50.0), np.log(4000.0), (n, aa.STABLE_DTYPE["pitch"].shape[0]))).astype(np.float32)
    st["pitch"]["freq"] = f
    st["pitch"]["score"] = rng.choice([0.25, 0.5, 0.75, 1.0], f.shape).astype(np.float32)     # many score ties
    st["pitch"]["freq"][:13, 1] = st["pitch"]["freq"][:13, 0] * (np.float32(2.0) ** (np.arange(13, dtype=np.float32) / 12))
    st["pitch"]["freq"][13, 0] = 0.0
    for system in (0, 1, 2):
        for single in (False, True):
            got = aa.tuner_from_stable(st, 440.0, system, single)
            for i in range(n):
                k = int(st["n"][i])
                pairs = np.stack([st["pitch"]["freq"][i, :k], st["pitch"]["score"][i, :k]], axis=1) if k else np.zeros((0, 2))
                r = O.tuner_frame(pairs, system, single)
                g = got[i]
                assert (g["kind"], g["best"], g["lo"], g["hi"]) == (r["kind"], r["best"], r["lo"], r["hi"]), (i, system, single)
                if r["kind"] == 2:
                    near_tie = False
                    if g["interval"] != r["interval"]:      # |ratio - r_i| == |ratio - r_j| to the last ulp: not expected
                        near_tie = True
                    assert not near_tie
                    assert abs(g["accuracy"] - r["accuracy"]) <= 1e-4 * max(1.0, abs(r["accuracy"]))
                    assert g["cents"] == g["accuracy"]
                if r["kind"] == 1:
                    cents = O.note_from_freq(float(st["pitch"]["freq"][i, r["best"]]), 440.0)[3]
                    assert abs(g["cents"] - cents) <= 2e-3
    assert aa.INT_TYPES[aa.tuner_from_stable(st[:13], 440.0, 0, False)["interval"][7]] == "Per5"
    with pytest.raises(aa.AAError):
        aa.tuner_from_stable(st, 440.0, 3, False)


def test_offline_onset_events(aa, O, torch_cuda):
    """aa_onset_events_* (SURVEY 8f rank 3): fired frames compacted into OnsetEvent lists; bit-exact against the
    oracle restatement of onset.rs:386-390 + timing.rs:311-337 (velocity in f32, beat position in f64)."""
    sr, n = 48000.0, 1024
    hop = n // 4
    clips = np.stack([signals.note_sequence(60 + i, sr, 96000, n_notes=10) for i in range(9)])
    clips[8] *= 0.0
    cfg = aa.Config(n=n, sample_rate=sr)
    res = aa.Analyzer(cfg).analyze_host(clips, want_mags=False)
    ev, cnt = aa.onset_events(cfg, res["features"], clips.shape[1], bpm=97.5, max_events=64)
    fired = (res["features"]["flags"] & aa.FLAG_ONSET_FIRED) != 0
    assert np.array_equal(cnt, fired.sum(axis=1)) and cnt[8] == 0 and cnt[:8].min() >= 3
    for c in range(len(clips)):
        want, total = O.onset_events(res["features"][c], n, hop, sr, 97.5, 64)
        assert total == cnt[c]
        assert ev[c][: cnt[c]].tobytes() == want.tobytes()
        assert np.array_equal(ev[c]["frame"][: cnt[c]], np.nonzero(fired[c])[0])
        assert np.all(ev[c]["sample_position"][: cnt[c]] == ev[c]["frame"][: cnt[c]] * hop + n // 2)
    # truncation: counts keep the total, only max_events are written (the earliest ones)
    ev2, cnt2 = aa.onset_events(cfg, res["features"], clips.shape[1], bpm=97.5, max_events=2)
    assert np.array_equal(cnt2, cnt) and ev2[0].tobytes() == ev[0][:2].tobytes()
    with pytest.raises(aa.AAError):
        aa.onset_events(cfg, res["features"], clips.shape[1], bpm=0.0)
