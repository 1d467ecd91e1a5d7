"""GPU parity tests of the input conditioning chain through the C ABI
(aa_condition_host / aa_condition_device) against the CPU oracle (oracle/aa_oracle_cond.c).

Bars: filters + gate are exact f32 recurrences in the reference's operation order -> BIT-EXACT;
the DynamicsTracker goes through log10 / powf (libm vs CUDA differ by <= 2 ulp), so gains and dB
values are compared to 1e-5 relative / 2e-4 dB and discrete decisions may differ at documented
near-ties (an rms_db within 1e-3 dB of a threshold), bounded at 1 % of the slots."""
import numpy as np
import pytest

import signals

pytestmark = pytest.mark.gpu


def make_batch(sr, n, n_clips, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    clips = []
    for c in range(n_clips):
        kind = c % 4
        if kind == 0:       # notes with silences in between (gate opens / holds / closes, playing vs silence)
            x = signals.note_sequence(seed * 100 + c, sr, n, n_notes=5, noise_db=-95.0)
        elif kind == 1:     # steady multitone
            x = signals.multitone(seed * 100 + c, sr, n, noise_db=-70.0)
        elif kind == 2:     # broadband noise around -50 dBFS with a loud tone burst in the middle
            x = 3e-3 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 660 * t) * ((t > 0.4 * t[-1]) & (t < 0.6 * t[-1]))
        else:               # decaying burst into digital silence
            x = 0.2 * np.exp(-t / 0.05) * np.sin(2 * np.pi * 220 * t) + 0.0
        clips.append(np.asarray(x, np.float32))
    return np.stack(clips)


def oracle_batch(O, clips, sr, slot_len, agc):
    ys, ds = [], []
    for x in clips:
        y, d = O.condition_clip(x, sr, slot_len, agc=agc)
        ys.append(y)
        ds.append(d)
    return np.stack(ys), np.stack(ds)


@pytest.mark.parametrize("sr,slot_len,n_clips", [(48000.0, 1024, 37), (44100.0, 1024, 32), (48000.0, 256, 37),
                                                 (96000.0, 128, 65), (32000.0, 192, 5), (48000.0, 1024, 1)])
def test_filters_and_gate_are_bit_exact(aa, O, torch_cuda, sr, slot_len, n_clips):
    """Slot lengths that 128 divides go through the cluster pipeline (two SMs per 32 clips: full groups, a ragged last
    group, a single clip), the others through the one-thread-per-clip kernel."""
    n = slot_len * (96 if slot_len == 1024 else 200) + 100          # + a partial slot that must stay untouched
    n -= n % 4
    clips = make_batch(sr, n, n_clips, seed=1)
    cond = aa.Conditioner(sr, slot_len, agc=False)
    got, _ = cond.process_host(clips)
    ref, _ = oracle_batch(O, clips, sr, slot_len, agc=False)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    full = (n // slot_len) * slot_len
    assert np.array_equal(got[:, full:], clips[:, full:])
    assert not np.array_equal(got[:, :full], clips[:, :full])


def test_gate_states_are_all_exercised(aa, O, torch_cuda):
    """Quiet passages between loud ones at several levels around the gate threshold: samples with the gate open, held
    and closing (gain = (envelope / threshold)^4) all occur, the hold counter expires and is re-armed, and the output
    is still the oracle's, bit for bit."""
    sr, L = 48000.0, 1024
    n = L * 150
    rng = np.random.default_rng(21)
    t = np.arange(n) / sr
    clips = []
    for c in range(40):
        level = 10.0 ** rng.uniform(-4.5, -1.0)                   # around the -50 dBFS gate threshold
        period = rng.uniform(0.05, 0.9)
        duty = rng.uniform(0.1, 0.9)
        on = ((t / period) % 1.0) < duty
        x = level * np.sin(2 * np.pi * rng.uniform(100, 3000) * t) * on + 10.0 ** rng.uniform(-6.0, -3.5) * rng.standard_normal(n)
        clips.append(x.astype(np.float32))
    clips = np.stack(clips)
    got, _ = aa.Conditioner(sr, L, agc=False).process_host(clips)
    ref, _ = oracle_batch(O, clips, sr, L, agc=False)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # the gate did close somewhere (output far below the input level) and did stay open elsewhere
    win = 4096
    rin = np.sqrt((clips[:, : n // win * win].reshape(40, -1, win) ** 2).mean(-1))
    rout = np.sqrt((ref[:, : n // win * win].reshape(40, -1, win) ** 2).mean(-1))
    assert (rout < 0.05 * rin).any() and (rout > 0.5 * rin).any()


def check_dynamics(got, ref, max_flip=0.01):
    assert got.shape == ref.shape
    assert np.allclose(got["rms_db"], ref["rms_db"], atol=2e-4)
    assert np.allclose(got["noise_floor_db"], ref["noise_floor_db"], atol=2e-4)
    flip = got["flags"] != ref["flags"]
    assert flip.mean() <= max_flip
    # a flipped activity decision changes what enters the histories: compare the rest on clips without flips
    ok = ~flip.any(axis=1)
    assert ok.mean() > 0.9
    g, r = got[ok], ref[ok]
    assert np.allclose(g["effective_gain"], r["effective_gain"], rtol=1e-5)
    assert np.allclose(g["session_median_db"], r["session_median_db"], atol=2e-4)
    assert np.allclose(g["gain_db"], r["gain_db"], atol=2e-4)
    assert (g["level"] != r["level"]).mean() <= max_flip
    return ok


def test_full_chain_matches_the_oracle(aa, O, torch_cuda):
    sr, L = 48000.0, 1024
    n = L * 280
    clips = make_batch(sr, n, 24, seed=2)
    cond = aa.Conditioner(sr, L, agc=True)
    got, dyn = cond.process_host(clips)
    ref, rdyn = oracle_batch(O, clips, sr, L, agc=True)
    ok = check_dynamics(dyn, rdyn)
    assert np.allclose(got[ok], ref[ok], rtol=2e-5, atol=0)
    assert {0, 5} <= set(np.unique(dyn["level"]))
    assert (dyn["flags"] & 2).any()                         # the broadband branch was exercised
    # stage-isolated: given the GPU's own gains, the output is exactly gated * gain
    gated, _ = aa.Conditioner(sr, L, agc=False).process_host(clips)
    want = (gated.reshape(len(clips), -1, L) * dyn["effective_gain"][:, :, None]).reshape(len(clips), -1)
    assert np.array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))


def test_history_rings_wrap(aa, O, torch_cuda):
    """More than 256 quiet and more than 5000 active slots: both percentile windows replace their oldest entry."""
    sr, L = 48000.0, 64
    n_slots = 6200
    n = L * n_slots
    t = np.arange(n) / sr
    rng = np.random.default_rng(7)
    lvl = 0.02 + 0.3 * np.abs(np.sin(2 * np.pi * 0.37 * t)) * (1 + 0.3 * np.sin(2 * np.pi * 0.05 * t))
    x = lvl * np.sin(2 * np.pi * 440 * t)
    x[: L * 400] = 1e-4 * rng.standard_normal(L * 400)       # 400 quiet slots first
    x[L * 3000: L * 3300] = 1e-4 * rng.standard_normal(L * 300)
    clips = np.stack([x.astype(np.float32), (0.5 * x[::-1]).astype(np.float32)])
    cond = aa.Conditioner(sr, L, agc=True)
    got, dyn = cond.process_host(clips)
    ref, rdyn = oracle_batch(O, clips, sr, L, agc=True)
    assert (rdyn["flags"] & 4).sum(axis=1).min() > 5000 and ((rdyn["flags"] & 1) == 0).sum(axis=1).min() > 256
    ok = check_dynamics(dyn, rdyn, max_flip=0.01)
    assert ok.all()
    assert np.allclose(got, ref, rtol=2e-5, atol=0)


def test_carried_state_equals_one_pass(aa, O, torch_cuda):
    sr, L = 44100.0, 1024
    n = L * 120
    clips = make_batch(sr, n, 8, seed=3)
    one, dyn_one = aa.Conditioner(sr, L, agc=True).process_host(clips)
    cond = aa.Conditioner(sr, L, agc=True, carry=True)
    cut = L * 50
    a, da = cond.process_host(np.ascontiguousarray(clips[:, :cut]))
    b, db = cond.process_host(np.ascontiguousarray(clips[:, cut:]))
    assert np.array_equal(np.concatenate([a, b], axis=1).view(np.uint32), one.view(np.uint32))
    assert np.array_equal(np.concatenate([da, db], axis=1), dyn_one)
    cond.reset()
    a2, _ = cond.process_host(np.ascontiguousarray(clips[:, :cut]))
    assert np.array_equal(a2.view(np.uint32), a.view(np.uint32))


def test_device_entry_point_and_edge_cases(aa, O, torch_cuda):
    torch = torch_cuda
    sr, L = 48000.0, 1024
    cond = aa.Conditioner(sr, L, agc=True)
    # device buffers with a stride larger than the clip, conditioned in place
    n, stride, n_clips = L * 40, L * 40 + 64, 5
    clips = make_batch(sr, n, n_clips, seed=4)
    buf = torch.zeros(n_clips * stride, dtype=torch.float32, device="cuda")
    buf.view(n_clips, stride)[:, :n] = torch.from_numpy(clips).cuda()
    dyn = torch.zeros(n_clips * 40 * 8, dtype=torch.int32, device="cuda")
    cond.process_device(buf.data_ptr(), n_clips, n, stride, dyn.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    host, hdyn = cond.process_host(clips)
    assert np.array_equal(buf.view(n_clips, stride)[:, :n].cpu().numpy().view(np.uint32), host.view(np.uint32))
    assert np.array_equal(dyn.cpu().numpy().view(aa.DYNAMICS_DTYPE).reshape(n_clips, 40), hdyn)
    assert float(buf.view(n_clips, stride)[:, n:].abs().max()) == 0.0      # padding between clips untouched
    # a clip shorter than one slot has no slots: nothing happens
    short = np.ones((2, 512), np.float32)
    y, d = cond.process_host(short)
    assert np.array_equal(y, short) and d.shape == (2, 0)
    assert cond.num_slots(441000) == 430                                   # mod.rs:799-803
    # argument checks
    with pytest.raises(aa.AAError):
        aa.Conditioner(sr, 1023)
    with pytest.raises(aa.AAError):
        aa.Conditioner(0.0, 1024)
    with pytest.raises(aa.AAError):
        cond.process_device(buf.data_ptr() + 4, 1, n, stride)


def test_pcm_ingest_is_exact_and_analyze_host_pcm_equals_the_f32_path(aa, O, torch_cuda):
    """mod.rs:765-792 on the device: bit-exact against the oracle for every format / channel count, and
    aa_analyze_host_pcm(i16) == aa_analyze_host(the same samples as f32), byte for byte."""
    torch = torch_cuda
    rng = np.random.default_rng(12)
    n_clips, clip_len = 5, 4096 + 8
    for fmt, dt in ((aa.PCM_I16, np.int16), (aa.PCM_U16, np.uint16), (aa.PCM_F32, np.float32)):
        for ch in (1, 2, 3):
            if fmt == aa.PCM_F32:
                pcm = rng.standard_normal((n_clips, clip_len * ch)).astype(np.float32)
            else:
                info = np.iinfo(dt)
                pcm = rng.integers(info.min, info.max + 1, (n_clips, clip_len * ch)).astype(dt)
            src = torch.from_numpy(pcm.view(np.uint8).reshape(-1).copy()).cuda()
            out = torch.zeros(n_clips * clip_len, dtype=torch.float32, device="cuda")
            aa.ingest_device(src.data_ptr(), fmt, ch, n_clips, clip_len, clip_len, clip_len, out.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            want = np.stack([O.ingest(row, fmt, ch) for row in pcm])
            assert np.array_equal(out.cpu().numpy().reshape(n_clips, clip_len).view(np.uint32), want.view(np.uint32)), (fmt, ch)
    # end to end: 16-bit stereo PCM through the host pipeline == the mixed-down f32 clips through aa_analyze_host
    sr, n = 48000.0, 2048
    clips = np.stack([signals.multitone(40 + i, sr, 20 * n) for i in range(6)])
    l = np.clip(np.round(clips * 20000), -32768, 32767).astype(np.int16)
    r = np.clip(np.round(clips[::-1] * 9000), -32768, 32767).astype(np.int16)
    pcm = np.stack([l, r], axis=2).reshape(len(clips), -1)
    mono = np.stack([O.ingest(row, aa.PCM_I16, 2) for row in pcm])
    an = aa.Analyzer(aa.Config(n=n, sample_rate=sr))
    a = an.analyze_host_pcm(pcm, aa.PCM_I16, 2)
    b = an.analyze_host(mono)
    for k in ("features", "stable", "mags", "summaries"):
        assert a[k].tobytes() == b[k].tobytes(), k
    with pytest.raises(aa.AAError):
        an.analyze_host_pcm(pcm, 7, 2)
