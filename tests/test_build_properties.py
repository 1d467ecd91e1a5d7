"""Static properties of the compiled analysis kernel (cuobjdump on the in-tree object, no GPU needed):
the things DESIGN.md 3.1 claims about the SASS and that a refactor can silently lose."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "audio-analyzer-rs_b200", "build", "aa_analyze.o")
KERNELS = {   # production instantiations of the two headline window sizes (pitch + onset, LIVE = 2, one
    # sub-block per CTA; the packed form, SUBS > 1, is an opt-in experiment build)
    (4096, 1): "_ZN2aa14analyze_kernelILi4096ELb1ELb1ELb0ELi2ELi1EEEvNS_13AnalyzeParamsE",
    (2048, 1): "_ZN2aa14analyze_kernelILi2048ELb1ELb1ELb0ELi2ELi1EEEvNS_13AnalyzeParamsE",
}


def _sass(kernel):
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(OBJ):
        import importlib
        importlib.import_module("audio-analyzer-rs_b200.build").build()
    out = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, OBJ], capture_output=True, text=True).stdout
    ins = [m.group(1).strip() for m in re.finditer(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", out, re.M)]
    assert len(ins) > 1000, "kernel not found in the object"
    return ins


@pytest.mark.parametrize("n,subs", sorted(KERNELS))
def test_frame_loop_of_the_analysis_kernel(n, subs):
    ins = _sass(KERNELS[(n, subs)])
    start = next(i for i, s in enumerate(ins) if "SYNCS.PHASECHK" in s)           # hop mbarrier wait
    end = next(i for i, s in enumerate(ins) if i > start and "BAR.ARV" in s)       # FULL hand-off to the tail
    loop = ins[start:end + 1]
    # one TMA bulk copy per frame feeds the hop ring (UBLKCP), packed f32x2 arithmetic carries the FFT
    assert any("UBLKCP" in s for s in ins), "no TMA bulk copy in the kernel"
    assert sum(s.startswith(("FFMA2", "FADD2", "FMUL2")) for s in loop) > 300
    # five block barriers per frame (four FFT exchanges + the magnitude hand-over)
    assert sum("BAR.SYNC" in s for s in loop) == 5
    # no spill stores inside the frame loop, and only the few reloads of hoisted loop invariants
    assert sum(re.match(r"(@!?U?P\d+\s+)?STL", s) is not None for s in loop) == 0
    assert sum(re.match(r"(@!?U?P\d+\s+)?LDL", s) is not None for s in loop) <= 12
    # the hot loop stays below the instruction-cache cliff measured in tools/microbench (32 KB with the tail)
    assert len(loop) * 16 <= 31 * 1024
