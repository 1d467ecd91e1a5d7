"""The C++ mirror of the reference's Rust types (host/audio_engine_gpu.hpp) compiles against
include/aa_gpu.h, links libaa_gpu.so, and behaves: without a device every constructor fails
loudly (CPU test); on a B200 the FftProcessor / STFT / OnsetDetector workers produce results."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(aa, tmp_path):
    exe = str(tmp_path / "test_mirror")
    libdir = os.path.dirname(aa.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"),
           "-L" + libdir, "-laa_gpu", "-Wl,-rpath," + libdir, "-lpthread"]
    subprocess.check_call(cmd)
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_device(aa, tmp_path):
    exe = _build(aa, tmp_path)
    if aa.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    out = subprocess.run([exe, "nodevice"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "no CPU fallback" in out.stdout
    # host-side onset stamping (no device needed): the offsets and the f64 beat the real crate logged
    assert "stamp_onset_at reproduces the logged beat 0.6653333333333337 / sample 15968" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(aa, torch_cuda, tmp_path):
    exe = _build(aa, tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "STFT emitted 76 frames" in out.stdout and "OnsetDetector emitted" in out.stdout
    assert "Reducer level mf" in out.stdout and "Tuner labels A4 / Per5" in out.stdout
