"""ctypes mirrors of the flattened-model descriptors in include/jb200_model.h.

Host-side plumbing only: builds ``jb200_gmm_desc`` / ``jb200_dnn_desc`` /
``jb200_tree_desc`` structures whose pointers alias numpy arrays taken from a
JB2M blob (julius_b200.refdump.load_blob).  The returned object keeps the arrays
alive.  Used by the product's C-ABI wrapper (julius_b200.capi) and, in tests,
by the oracle wrapper -- the descriptors are the shared boundary format.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

F = C.POINTER(C.c_float)
I = C.POINTER(C.c_int32)
U8 = C.POINTER(C.c_uint8)
DNN_MAX_LAYERS = 16


class GmmDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_states", "dim", "n_gauss", "max_mix", "gprune_method", "gprune_num",
        "iwcd_method", "iwcd_nbest", "n_cdsets", "n_cdset_states")] + [
        ("state_off", I), ("mean", F), ("ivar", F), ("gconst", F), ("lnweight", F),
        ("valid", U8), ("cd_off", I), ("cd_states", I)]


class DnnDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("layer_in", C.c_int32 * DNN_MAX_LAYERS), ("layer_out", C.c_int32 * DNN_MAX_LAYERS),
                ("w", F * DNN_MAX_LAYERS), ("b", F * DNN_MAX_LAYERS), ("state_prior", F)]


_TREE_INTS = ("n_nodes", "n_arcs", "n_words", "n_start", "n_iso", "n_shared", "n_fscore", "n_scword",
              "n_rset", "n_ctx", "head_silwid", "tail_silwid", "multipath", "beam_width",
              "lm_nvocab", "lm_nbigram", "lm_mode", "lm_unk_id")
_TREE_FLOATS = ("lm_unk_num_log", "lm_weight", "lm_penalty", "lm_penalty_trans", "score_pruning_width")
_TREE_PTRS = (("self_a", F), ("next_a", F), ("arc_off", I), ("arc_to", I), ("arc_a", F), ("stend", I),
              ("scid", I), ("outstyle", U8), ("out_ref", I), ("rset_ctx", I), ("word_ctx", I),
              ("iso_node", I), ("iso_word", I), ("iso_id", I), ("shared_node", I),
              ("wordend_a", F), ("wordend", I), ("wordbegin", I), ("is_transparent", U8), ("wton", I),
              ("cprob", F), ("fscore", F), ("scword", I), ("uni_prob", F), ("uni_bow", F),
              ("bi_bgn", I), ("bi_num", I), ("bi_wid", I), ("bi_prob", F))


# grammar (DFA) mode tail of jb200_tree_desc (absent from blobs of N-gram models)
_TREE_DFA_PTRS = (("init_word", I), ("init_node", I), ("init_lscore", F), ("cp_allowed", U8))


class TreeDesc(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in _TREE_INTS] + [(n, C.c_float) for n in _TREE_FLOATS]
                + list(_TREE_PTRS)
                + [("lm_type", C.c_int32), ("n_init", C.c_int32), ("penalty1", C.c_float), ("reserved_", C.c_int32)]
                + list(_TREE_DFA_PTRS))


def _ptr(keep: list, a: np.ndarray, ctype):
    a = np.ascontiguousarray(a)
    keep.append(a)
    if a.size == 0:
        a = np.zeros(1, dtype=a.dtype)
        keep.append(a)
    return a.ctypes.data_as(ctype)


def _sc(blob, name, default=0):
    return blob[name][0].item() if name in blob else default


class Descriptors:
    """Owns the arrays + ctypes structs for one flattened model."""

    def __init__(self, blob: dict):
        self.blob = blob
        self._keep = []
        self.gmm = self._make_gmm() if "gmm.mean" in blob else None
        self.dnn = self._make_dnn() if "dnn.n_layers" in blob else None
        self.tree = self._make_tree() if "tree.self_a" in blob else None
        self.n_states = _sc(blob, "gmm.n_states")

    # -- override helpers (tests switch gprune / iwcd without regenerating the blob)
    def _make_gmm(self) -> GmmDesc:
        b, k = self.blob, self._keep
        g = GmmDesc()
        g.n_states = _sc(b, "gmm.n_states"); g.dim = _sc(b, "gmm.dim"); g.n_gauss = _sc(b, "gmm.n_gauss")
        g.max_mix = _sc(b, "gmm.max_mix"); g.gprune_method = _sc(b, "gmm.gprune_method")
        g.gprune_num = _sc(b, "gmm.gprune_num"); g.iwcd_method = _sc(b, "am.iwcd_method", 2)
        g.iwcd_nbest = _sc(b, "am.iwcd_nbest", 3); g.n_cdsets = _sc(b, "am.n_cdsets")
        g.n_cdset_states = _sc(b, "am.n_cdset_states")
        g.state_off = _ptr(k, b["gmm.state_off"], I); g.mean = _ptr(k, b["gmm.mean"], F)
        g.ivar = _ptr(k, b["gmm.ivar"], F); g.gconst = _ptr(k, b["gmm.gconst"], F)
        g.lnweight = _ptr(k, b["gmm.lnweight"], F); g.valid = _ptr(k, b["gmm.valid"], U8)
        g.cd_off = _ptr(k, b.get("am.cd_off", np.zeros(1, np.int32)), I)
        g.cd_states = _ptr(k, b.get("am.cd_states", np.zeros(1, np.int32)), I)
        return g

    def _make_dnn(self) -> DnnDesc:
        b, k = self.blob, self._keep
        d = DnnDesc()
        d.n_layers = _sc(b, "dnn.n_layers"); d.in_dim = _sc(b, "dnn.in_dim"); d.out_dim = _sc(b, "dnn.out_dim")
        for i in range(d.n_layers):
            d.layer_in[i] = _sc(b, f"dnn.l{i}.in"); d.layer_out[i] = _sc(b, f"dnn.l{i}.out")
            d.w[i] = _ptr(k, b[f"dnn.l{i}.w"], F); d.b[i] = _ptr(k, b[f"dnn.l{i}.b"], F)
        d.state_prior = _ptr(k, b["dnn.state_prior"], F)
        return d

    def _make_tree(self) -> TreeDesc:
        b, k = self.blob, self._keep
        t = TreeDesc()
        for n in _TREE_INTS:
            setattr(t, n, int(_sc(b, "tree." + n)))
        for n in _TREE_FLOATS:
            setattr(t, n, float(_sc(b, "tree." + n, -1.0 if n == "score_pruning_width" else 0.0)))
        for n, ct in _TREE_PTRS:
            setattr(t, n, _ptr(k, b["tree." + n], ct))
        t.lm_type = int(_sc(b, "tree.lm_type", 0)); t.n_init = int(_sc(b, "tree.n_init", 0))
        t.penalty1 = float(_sc(b, "tree.penalty1", 0.0))
        np_of = {I: np.int32, F: np.float32, U8: np.uint8}
        for n, ct in _TREE_DFA_PTRS:
            setattr(t, n, _ptr(k, b.get("tree." + n, np.zeros(1, np_of[ct])), ct))
        return t

    def cd_only_gmm(self) -> GmmDesc:
        """A GmmDesc carrying only the state/cd-set layout (DNN models have no Gaussians)."""
        b, k = self.blob, self._keep
        g = GmmDesc()
        g.n_states = _sc(b, "gmm.n_states"); g.iwcd_method = _sc(b, "am.iwcd_method", 2)
        g.iwcd_nbest = _sc(b, "am.iwcd_nbest", 3); g.n_cdsets = _sc(b, "am.n_cdsets")
        g.n_cdset_states = _sc(b, "am.n_cdset_states")
        g.cd_off = _ptr(k, b.get("am.cd_off", np.zeros(1, np.int32)), I)
        g.cd_states = _ptr(k, b.get("am.cd_states", np.zeros(1, np.int32)), I)
        return g
