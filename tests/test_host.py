"""CPU: host-side logic -- blob container round trip, descriptor structs, C-ABI symbols."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from julius_b200 import capi, desc, refdump, synth
from util import GOLDEN, ROOT, Golden


def test_blob_roundtrip(tmp_path):
    g = Golden("tiny")
    p = tmp_path / "x.jb2m"
    refdump.save_blob(str(p), g.blob)
    b2 = refdump.load_blob(str(p))
    assert list(b2) == list(g.blob)
    for k in g.blob:
        assert b2[k].dtype == g.blob[k].dtype and np.array_equal(b2[k], g.blob[k])


def test_descriptor_fields():
    g = Golden("small_b100")
    t = g.ds.tree
    assert t.n_nodes == len(g.blob["tree.self_a"])
    assert t.n_iso + t.n_shared == t.n_start
    assert t.beam_width == 100
    assert g.ds.gmm.n_gauss == g.blob["gmm.state_off"][-1]
    assert C.sizeof(desc.TreeDesc) % 8 == 0


def test_capi_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIBPATH), "libjb200.so not built (python -m julius_b200.build)"
    L = C.CDLL(capi.LIBPATH)
    hdr = open(os.path.join(ROOT, "include", "julius_b200.h")).read()
    names = set(re.findall(r"\b(jb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in julius_b200.h but not exported: {missing}"


def test_synth_is_seeded(tmp_path):
    a = synth.SynthModel(synth.SynthConfig.preset("tiny"))
    b = synth.SynthModel(synth.SynthConfig.preset("tiny"))
    assert np.array_equal(a.mean, b.mean) and a.words == b.words and a.bigrams == b.bigrams
    x, _ = a.sample_utterance(np.random.default_rng(3), 120)
    assert x.shape == (120, 39) and x.dtype == np.float32
    p = tmp_path / "f.mfc"
    synth.write_htk_param(str(p), x)
    y, kind = synth.read_htk_param(str(p))
    assert np.array_equal(x, y) and kind == synth.PARMKIND_MFCC_E_D_A


def test_descriptor_layouts_match_the_c_header(tmp_path):
    """The ctypes mirrors in julius_b200/desc.py must lay the descriptors out exactly as include/jb200_model.h does
    (the product library, the plugin and the oracle all read them through that header)."""
    import ctypes as C
    import shutil
    import subprocess
    from julius_b200 import desc
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    fields = {"jb200_tree_desc": (desc.TreeDesc, ["n_nodes", "lm_unk_num_log", "self_a", "bi_prob", "lm_type", "penalty1", "init_word", "cp_allowed"]),
              "jb200_gmm_desc": (desc.GmmDesc, ["n_states", "state_off", "valid", "cd_states"]),
              "jb200_dnn_desc": (desc.DnnDesc, ["n_layers", "layer_out", "w", "state_prior"])}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "jb200_model.h"', 'int main(void) {']
    for st, (_, names) in fields.items():
        src.append(f'  printf("{st} %zu", sizeof({st}));')
        for n in names:
            src.append(f'  printf(" %zu", offsetof({st}, {n}));')
        src.append('  printf("\\n");')
    src.append('  return 0; }')
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(cfile)], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    for line in out:
        if not line.strip():
            continue
        parts = line.split()
        cls, names = fields[parts[0]]
        want = [C.sizeof(cls)] + [getattr(cls, n).offset for n in names]
        assert [int(x) for x in parts[1:]] == want, parts[0]


def test_product_path_fails_loudly_without_a_device():
    """No CPU fallback: on a machine without an sm_100 GPU every create call of the C-ABI must return an error (and the
    Python mirror raise), never hand back a handle that computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = Golden("tiny")
    with pytest.raises(capi.Jb200Error):
        capi.GmmScorer(g.ds, mode=capi.GMM_EXACT)
    d = Golden("small_dnn")
    with pytest.raises(capi.Jb200Error):
        capi.DnnScorer(d.ds)
    assert capi.lib().jb200_device_count() <= 0 or True      # the call itself must not crash


def test_nothing_in_the_product_imports_the_oracle():
    """oracle/ is test infrastructure: no module of julius_b200/ and none of the C/CUDA sources may reference it."""
    pkg = os.path.join(ROOT, "julius_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".inc", ".c", ".h")):
                continue
            txt = open(os.path.join(dirpath, fn), errors="ignore").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "liboracle" in txt or "oracle/restate" in txt or '#include "oracle' in txt:
                bad.append(os.path.relpath(os.path.join(dirpath, fn), ROOT))
    assert not bad, bad


def test_user_lm_export_limits_are_loud(tmp_path):
    """-userlm is tabulated densely over the dictionary by the export step: a dictionary above the limit must make the
    start-up fail with a message, not produce a model that silently ignores the user functions."""
    import os
    import pytest
    from oracle import ffi, fixtures
    if not ffi.have_ref():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path)
    with pytest.raises(RuntimeError):
        fixtures.make_fixture("small", d, n_utts=1, n_frames=50, extra_args=["-userlm", "-b", "60"],
                              env_extra={"JREF_USERLM": "1", "JB200_USERLM_MAXWORDS": "100"})
    # within the limit the same call exports (402 words)
    m, files, dump, out = fixtures.make_fixture("small", d, n_utts=1, n_frames=50, extra_args=["-userlm", "-b", "60"],
                                                env_extra={"JREF_USERLM": "1"})
    assert os.path.getsize(os.path.join(d, "model.jb2m")) > 402 * 402 * 8
