"""GPU parity of the FftProcessor replacement (aa_fft_*): reference src/dsp/fft.rs:14-41."""
import numpy as np
import pytest

import signals
import util

pytestmark = pytest.mark.gpu
SIZES = [256, 512, 1024, 2048, 4096]


@pytest.mark.parametrize("n", SIZES)
def test_forward_matches_f64_and_oracle(aa, O, torch_cuda, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((37, n)).astype(np.float32)
    x[0] = signals.sine(440.0, 44100.0, n) * O.hann_window(n)
    x[1] = 0.0
    x[2] = 1.0
    x[3, :] = 0.0
    x[3, 5] = 1.0        # shifted impulse
    fft = aa.FftProcessor(n)
    got = fft.process_forward(x)
    assert got.shape == (37, n // 2 + 1) and got.dtype == np.complex64
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    scale = np.maximum(np.abs(ref).max(axis=1), 1e-30)
    err = np.abs(got - ref).max(axis=1) / scale
    assert err[np.abs(ref).max(axis=1) > 0].max() < 1e-6, err.max()
    assert np.all(got[1] == 0)
    assert np.all(got[:, 0].imag == 0) and np.all(got[:, -1].imag == 0)
    # same tolerance class as the f32 oracle
    orc = np.stack([O.rfft_f32(r) for r in x[:8]])
    assert (np.abs(got[:8] - orc).max(axis=1) / scale[:8].clip(1e-30))[[0, 2, 3, 4, 5, 6, 7]].max() < 1e-6


@pytest.mark.parametrize("batch", [1, 2, 147, 1185, 5000])
def test_forward_ragged_batches(aa, torch_cuda, batch):
    n = 1024
    rng = np.random.default_rng(batch)
    x = rng.standard_normal((batch, n)).astype(np.float32)
    got = aa.FftProcessor(n).process_forward(x)
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert (np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-6


def test_single_frame_api_shape(aa, torch_cuda):
    fft = aa.FftProcessor(2048)
    out = fft.process_forward(np.ones(2048, np.float32))
    assert out.shape == (1025,) and abs(out[0] - 2048) < 1e-3 and np.abs(out[1:]).max() < 1e-3
    with pytest.raises(aa.AAError):
        fft.process_forward(np.ones(2047, np.float32))     # reference: unwrap() panic, fft.rs:69


@pytest.mark.parametrize("n", SIZES)
def test_inverse_roundtrip_and_linearity(aa, torch_cuda, n):
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal((9, n)).astype(np.float32)
    fft = aa.FftProcessor(n)
    spec = fft.process_forward(x)
    back = fft.process_inverse(spec)
    assert np.abs(back / n - x).max() < 2e-6 * np.sqrt(n)     # realfft: inverse(forward(x)) = n x
    # linearity
    a, b = x[0], x[1]
    s = fft.process_forward(np.stack([a, b, a + 2 * b]))
    assert np.abs(s[2] - (s[0] + 2 * s[1])).max() / np.abs(s[2]).max() < 2e-6


def test_forward_device_pointers_and_stream(aa, torch_cuda):
    torch = torch_cuda
    n, batch = 4096, 300
    x = torch.randn(batch, n, device="cuda", dtype=torch.float32)
    out = torch.empty(batch, n // 2 + 1, 2, device="cuda", dtype=torch.float32)
    fft = aa.FftProcessor(n)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fft.forward_device(x.data_ptr(), batch, out.data_ptr(), s.cuda_stream)
    s.synchronize()
    ref = torch.fft.rfft(x.double(), dim=1)
    got = torch.view_as_complex(out).to(torch.complex128)
    err = ((got - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)).max().item()
    assert err < 1e-6
