"""GPU tests of the streaming boundary (aa_stream_*): the SlotPool -> ring hand-off of
stft.rs:240-266 / onset.rs:216-237 replaced by a pinned-host + device ring."""
import numpy as np
import pytest

import signals
import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,slot", [(2048, 1024), (1024, 256), (256, 1024), (1024, 1000), (4096, 4096)])
def test_stream_equals_batch_bit_for_bit(aa, O, torch_cuda, n, slot):
    """Pushing 1024-sample slots (mod.rs:126-128) -- or any other size -- yields exactly the frames
    of the offline run: same kernel, state carried across pushes in device memory."""
    sr = 48000.0
    x = signals.note_sequence(11, sr, 40 * n)
    cfg = aa.Config(n=n, sample_rate=sr)
    batch = aa.Analyzer(cfg).analyze_host(x[None, :], want_mags=False)
    st = aa.Stream(cfg)
    frames = []
    for i in range(0, len(x) - slot + 1, slot):
        st.push(x[i:i + slot])
        got = st.poll(256)
        if len(got):
            frames.append(got)
    frames = np.concatenate(frames)
    used = (len(x) // slot) * slot
    T = (used - n) // (n // 4) + 1
    assert len(frames) == T
    assert np.array_equal(frames["frame_index"], np.arange(T))
    assert frames["features"].tobytes() == batch["features"][0, :T].tobytes()
    assert frames["stable"].tobytes() == batch["stable"][0, :T].tobytes()


def test_stream_long_run_compaction_and_reset(aa, torch_cuda):
    """Enough pushes to wrap the device buffer several times."""
    n, sr = 1024, 48000.0
    cfg = aa.Config(n=n, sample_rate=sr)
    x = signals.multitone(5, sr, 300000)
    batch = aa.Analyzer(cfg).analyze_host(x[None, :], want_mags=False)
    st = aa.Stream(cfg)
    out = []
    for i in range(0, len(x) - 1023, 1024):
        st.push(x[i:i + 1024])
        out.append(st.poll(64))
    fr = np.concatenate(out)
    assert fr["features"].tobytes() == batch["features"][0, : len(fr)].tobytes()
    st.reset()
    st.push(x[:4096])
    again = st.poll(64)
    assert again["features"].tobytes() == batch["features"][0, : len(again)].tobytes()
    assert again["frame_index"][0] == 0


def test_stream_onset_signal_and_noise_floor(aa, O, torch_cuda):
    n, sr = 2048, 44100.0
    x = signals.multitone(33, sr, 60000)
    cfg = aa.Config(n=n, sample_rate=sr)
    st = aa.Stream(cfg)
    T = (60000 // 1024 * 1024 - n) // 512 + 1
    onset = np.zeros(T, np.uint8)
    frames = []
    produced = 0
    for i in range(0, len(x) - 1023, 1024):
        if i // 1024 == 20:
            st.signal_onset()              # consumed by the next processed frame (stft.rs:387)
            onset[produced] = 1
        st.push(x[i:i + 1024])
        got = st.poll()
        produced += len(got)
        frames.append(got)
    fr = np.concatenate(frames)
    # compare with the batch path given the same per-frame onset flags
    Tb = (len(x) - n) // 512 + 1
    onset_b = np.zeros((1, Tb), np.uint8)
    onset_b[0, : len(onset)] = onset
    ref = aa.Analyzer(cfg).analyze_host(x[None, :], onset_in=onset_b, want_mags=False)
    assert fr["stable"].tobytes() == ref["stable"][0, : len(fr)].tobytes()
    # noise floor change takes effect on later frames
    st2 = aa.Stream(cfg)
    st2.set_noise_floor_db(-40.0)
    st2.push(x[:8192])
    a = st2.poll()
    ref2 = aa.Analyzer(aa.Config(n=n, sample_rate=sr, noise_floor_db=-40.0)).analyze_host(x[None, :8192], want_mags=False)
    assert a["features"].tobytes() == ref2["features"][0, : len(a)].tobytes()


def test_stream_errors(aa, torch_cuda):
    st = aa.Stream(aa.Config(n=1024, sample_rate=48000.0))
    with pytest.raises(aa.AAError) as e:
        st.push(np.zeros(100000, np.float32))
    assert e.value.code == -5
    st.push(np.zeros(0, np.float32))
    assert len(st.poll()) == 0


def test_stream_push_without_polling_fails_cleanly_and_can_be_retried(aa, torch_cuda):
    """A caller that keeps pushing without polling fills the result ring: the push that would overflow it fails
    with AA_ERR_OVERFLOW BEFORE anything is consumed, so polling and pushing the same samples again loses and
    duplicates nothing -- the frames still equal the offline run bit for bit (and no device buffer is overrun,
    however often the failing push is repeated)."""
    n, sr, slot = 1024, 48000.0, 1024
    cfg = aa.Config(n=n, sample_rate=sr)
    x = signals.multitone(17, sr, 400 * slot)
    batch = aa.Analyzer(cfg).analyze_host(x[None, :], want_mags=False)
    st = aa.Stream(cfg)
    frames, refused, i = [], 0, 0
    while i + slot <= len(x):
        try:
            st.push(x[i:i + slot])
            i += slot
        except aa.AAError as e:
            assert e.code == -5                       # AA_ERR_OVERFLOW
            refused += 1
            for _ in range(20):                       # hammering the full ring must not corrupt anything
                with pytest.raises(aa.AAError):
                    st.push(x[i:i + slot])
            got = st.poll(4096)
            assert len(got) > 0
            frames.append(got)
    frames.append(st.poll(4096))
    fr = np.concatenate(frames)
    T = (len(x) // slot * slot - n) // (n // 4) + 1
    assert refused >= 2, "the result ring never filled: the test did not exercise the overflow path"
    assert len(fr) == T and np.array_equal(fr["frame_index"], np.arange(T))
    assert fr["features"].tobytes() == batch["features"][0, :T].tobytes()
    assert fr["stable"].tobytes() == batch["stable"][0, :T].tobytes()
