"""Build libaa_gpu.so (in-tree) with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libaa_gpu.so")
SOURCES = ["aa_analyze.cu", "aa_fft_batch.cu", "aa_misc.cu", "aa_yin.cu", "aa_cond.cu", "aa_api.cu"]
HEADERS = ["aa_fft.cuh", "aa_tma.cuh", "aa_internal.h", os.path.join("..", "..", "include", "aa_gpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    nvcc = _nvcc()
    extra = []
    # AA_SO_OUT=path: build an experiment variant next to the in-tree library without touching it.
    # Compile-time experiment knobs (env AA_DEF_<NAME>=v -> -DAA_<NAME>=v) are honoured ONLY for such a
    # variant: the in-tree library is always built from the sources alone, so a leftover environment
    # variable (or a run-time knob like AA_SEG_MIN) can never change the shipped kernel.
    so_out = os.environ.get("AA_SO_OUT") or SO
    if so_out != SO:
        for macro in sorted(os.environ):
            if macro.startswith("AA_DEF_") and os.environ[macro] != "":
                extra.append(f"-DAA_{macro[len('AA_DEF_'):]}={os.environ[macro]}")
                force = True
        if os.environ.get("AA_NVCC_EXTRA"):      # extra nvcc flags of an experiment variant, e.g. --extra-device-vectorization
            extra += os.environ["AA_NVCC_EXTRA"].split()
            force = True
    objdir = os.path.join(HERE, "build") if so_out == SO else os.path.join(HERE, "build", "variant")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose or ptxas_info:
            sys.stdout.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(so_out, objs):
        cmd = [nvcc, "-shared", "-o", so_out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return so_out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print("built", SO)
