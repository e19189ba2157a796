import json
import os

import numpy as np

from julius_b200 import desc, refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
# small_userlm: user-defined LM functions (-userlm) on top of the N-gram, tabulated by the exporter
CASES = ["tiny", "small_b100", "small_safe", "small_mp", "small_iwsp", "small_userlm"]
DNN_CASES = ["small_dnn", "small_dnn_iwsp"]
# pinned on the CPU only so far (the GPU suite does not run them yet)
ORACLE_ONLY_CASES = ["small_tr", "small_tm", "small_dfa"]


class Golden:
    def __init__(self, name):
        d = os.path.join(GOLDEN, name)
        self.dir = d
        self.blob = refdump.load_blob(os.path.join(d, "model.jb2m"))
        self.ds = desc.Descriptors(self.blob)
        self.utts = refdump.load_refdump(os.path.join(d, "out.jrf"))
        z = np.load(os.path.join(d, "feats.npz"))
        self.feats = [z[f"u{i}"] for i in range(len(self.utts))]
        self.meta = json.load(open(os.path.join(d, "meta.json")))


def atoms_equal(a, b):
    """bit-exact comparison of two structured atom arrays (wid, begin, end, backscore, lscore, last)."""
    if len(a) != len(b):
        return False, f"atom count {len(a)} != {len(b)}"
    for k in ("wid", "begin", "end", "last"):
        if not np.array_equal(a[k], b[k]):
            i = int(np.nonzero(a[k] != b[k])[0][0])
            return False, f"field {k} differs first at atom {i}: {a[i]} vs {b[i]}"
    for k in ("backscore", "lscore"):
        if not np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)):
            i = int(np.nonzero(a[k].view(np.uint32) != b[k].view(np.uint32))[0][0])
            return False, f"field {k} differs (bits) first at atom {i}: {a[i]} vs {b[i]}"
    return True, ""


def rel_err(a, b, floor=1.0):
    """|a-b| / max(|a|,|b|,floor): the relative criterion with an absolute floor (hard parts)."""
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def full_dnn_blob(seed=3, in_dim=528, hidden=2048, layers=7, n_out=3000):
    """A random-init DNN of the BASELINE configs[3] shape (528 = 48 x 11 inputs, 7 x 2048 logistic, n_out states) as a
    flattened-model blob dict, with the initialisation of julius_b200.synth.write_dnn (W ~ N(0, 1.5/sqrt(in)), output
    layer 3/sqrt(in), b ~ N(0, 0.1), Dirichlet priors stored as log10, calc_dnn.c:699-703)."""
    rng = np.random.default_rng(seed)
    dims = [in_dim] + [hidden] * layers + [n_out]
    b = {"dnn.n_layers": np.array([layers + 1], np.int32), "dnn.in_dim": np.array([in_dim], np.int32),
         "dnn.out_dim": np.array([n_out], np.int32), "gmm.n_states": np.array([n_out], np.int32)}
    for i in range(layers + 1):
        scale = (3.0 if i == layers else 1.5) / np.sqrt(dims[i])
        b[f"dnn.l{i}.in"] = np.array([dims[i]], np.int32)
        b[f"dnn.l{i}.out"] = np.array([dims[i + 1]], np.int32)
        b[f"dnn.l{i}.w"] = (rng.standard_normal((dims[i + 1], dims[i])) * scale).astype(np.float32).ravel()
        b[f"dnn.l{i}.b"] = (rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32)
    prior = rng.dirichlet(np.full(n_out, 5.0))
    b["dnn.state_prior"] = np.log10(prior).astype(np.float32)
    return b
