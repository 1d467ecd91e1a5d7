"""CPU checks of bench.py's command line: the BASELINE workload presets and the config dict both arms print."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _parse(m, argv, world=1):
    old, env = sys.argv, os.environ.get("WORLD_SIZE")
    sys.argv = ["bench.py"] + argv
    os.environ["WORLD_SIZE"] = str(world)
    try:
        return m.parse()
    finally:
        sys.argv = old
        if env is None:
            os.environ.pop("WORLD_SIZE", None)
        else:
            os.environ["WORLD_SIZE"] = env


def test_default_is_baseline_cfg2():
    m = _bench()
    a = _parse(m, [])
    assert (a.workload, a.clips, a.seconds, a.n, a.sr, a.features, a.no_mags) == ("cfg2", 1024, 30.0, 4096, 48000.0, 15, False)
    assert (a.gpus, a.steps, a.warmup) == (1, 5, 3)
    assert m.bytes_per_frame(4096, 15, True) == 4 * 1024 + 4 * 2049 + 96 + 136


def test_other_workloads():
    m = _bench()
    a = _parse(m, ["--workload", "cfg1"])
    assert (a.clips, a.n, a.sr, a.seconds) == (1, 2048, 44100.0, 10.0)              # stft.rs:169-170
    x = m.sine440(a, 1)
    assert x.shape == (1, 441000) and abs(float(x.max()) - 0.5) < 1e-6
    a = _parse(m, ["--workload", "cfg5"], world=8)
    assert a.clips == 65536 // 8 and a.n == 2048 and a.no_mags
    a = _parse(m, ["--workload", "cfg5"])
    assert a.clips == 65536
    a = _parse(m, ["--workload", "cfg3"])
    assert (a.n, a.n // 4) == (1024, 256)
    a = _parse(m, ["--workload", "cfg5", "--clips", "64", "--n", "1024"])          # explicit flags win
    assert a.clips == 64 and a.n == 1024


def test_both_arms_print_the_same_config():
    m = _bench()
    a = _parse(m, [])
    T = (m.clip_len_of(a) - a.n) // (a.n // 4) + 1
    assert T == 1403
    ours = m.config_dict(a, 1, a.clips * T)
    ref = m.config_dict(a, 1, a.clips * T)
    assert ours == ref and ours["frames_per_gpu_per_step"] == 1436672 and "cfg2" in ours["workload"]
