"""CPU: the beam cut as the kernel formulates it (sentinel-padded heap, clamped speculative child address, loser cut)
against the reference's sort_token_upward/_downward loop, on random inputs with routine exact ties
(tools/heapsim.cpp is the statement-by-statement model of heap_pad_sentinels + heap_extract_fast in beam.cu)."""
import os
import shutil
import subprocess

import pytest

from util import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host C++ compiler")
@pytest.mark.parametrize("divisor", ["300", "7", "1"])       # many exact ties ... hardly any
def test_kernel_formulation_of_the_heap_select_equals_the_reference_loop(divisor, tmp_path):
    exe = str(tmp_path / "heapsim")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tools", "heapsim.cpp")], check=True)
    p = subprocess.run([exe, divisor, "600"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "mismatches 0" in p.stdout
