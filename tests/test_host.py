"""CPU: host-side logic -- blob container round trip, descriptor structs, C-ABI symbols."""
import ctypes as C
import os
import re

import numpy as np

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
