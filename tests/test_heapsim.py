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


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host C++ compiler")
@pytest.mark.parametrize("distinct", ["20000", "300", "8", "2"])       # hardly any exact ties ... nothing but ties
def test_closed_form_with_relocations_equals_the_reference_loop(distinct, tmp_path):
    """tools/heapdyn.cpp: the closed form of the upward beam cut with the exact treatment of re-inserted elements, in three
    forms (subtree queries, forward scans, and a lane-by-lane emulation of closed_relocate in csrc/beam.cu -- same packed keys,
    chunks, ballots, done masks and shifts), each against sort_token_upward's extraction loop (beam.c:1342-1384)."""
    exe = str(tmp_path / "heapdyn")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(ROOT, "tools", "heapdyn.cpp")], check=True)
    p = subprocess.run([exe, "7", "400", distinct], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "mismatches 0" in p.stdout
