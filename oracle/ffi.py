"""ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement) -- TEST INFRASTRUCTURE.

Also: helpers that run the compiled reference (oracle/_ref/jref) on generated fixtures.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from julius_b200 import desc as D

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
JREF = os.path.join(HERE, "_ref", "jref")
JREF_GPU = os.path.join(HERE, "_ref", "jref_gpu")
PLUGDIR = os.path.join(HERE, "_ref")

ATOM_DT = np.dtype([("wid", "<i4"), ("begin", "<i4"), ("end", "<i4"),
                    ("backscore", "<f4"), ("lscore", "<f4"), ("last", "<i4")])
_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oracle_gmm_score.argtypes = [C.POINTER(D.GmmDesc), D.F, C.c_int, D.F]
        _lib.oracle_cdset_score.argtypes = [C.POINTER(D.GmmDesc), D.F, C.c_int, D.F]
        _lib.oracle_addlog_table.argtypes = [D.F]
        _lib.oracle_beam_decode.argtypes = [C.POINTER(D.TreeDesc), C.POINTER(D.GmmDesc), D.F, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, D.I, D.I, D.F, D.I, D.I]
        _lib.oracle_beam_decode.restype = C.c_int
        if hasattr(_lib, "oracle_dnn_score"):
            _lib.oracle_dnn_score.argtypes = [C.POINTER(D.DnnDesc), D.F, C.c_int, D.F]
    return _lib


def _f(a):
    return a.ctypes.data_as(D.F)


def gmm_score(ds: D.Descriptors, feats: np.ndarray) -> np.ndarray:
    feats = np.ascontiguousarray(feats, np.float32)
    T = feats.shape[0]
    out = np.empty((T, ds.gmm.n_states), np.float32)
    lib().oracle_gmm_score(C.byref(ds.gmm), _f(feats), T, _f(out))
    return out


def cdset_score(ds: D.Descriptors, st: np.ndarray, gmm=None) -> np.ndarray:
    g = gmm if gmm is not None else ds.gmm
    st = np.ascontiguousarray(st, np.float32)
    T = st.shape[0]
    out = np.empty((T, max(1, g.n_cdsets)), np.float32)
    lib().oracle_cdset_score(C.byref(g), _f(st), T, _f(out))
    return out[:, :g.n_cdsets]


def dnn_score(ds: D.Descriptors, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    T = x.shape[0]
    out = np.empty((T, ds.dnn.out_dim), np.float32)
    lib().oracle_dnn_score(C.byref(ds.dnn), _f(x), T, _f(out))
    return out


def addlog_table() -> np.ndarray:
    t = np.empty(500000, np.float32)
    lib().oracle_addlog_table(_f(t))
    return t


def beam_decode(ds: D.Descriptors, st: np.ndarray, gmm=None, max_atoms: int = 1 << 20, trace: bool = False):
    """Returns dict(atoms=structured array, words=list, score=float, status=int[, trace=[T,2]])."""
    g = gmm if gmm is not None else (ds.gmm if ds.gmm is not None else ds.cd_only_gmm())
    st = np.ascontiguousarray(st, np.float32)
    T, S = st.shape
    atoms = np.zeros(max_atoms, ATOM_DT)
    words = np.zeros(160, np.int32)
    nbest = C.c_int32(0); score = C.c_float(0); status = C.c_int32(0)
    tr = np.zeros((max(T, 1), 2), np.int32) if trace else None
    n = lib().oracle_beam_decode(C.byref(ds.tree), C.byref(g), _f(st), T, S,
                                 atoms.ctypes.data_as(C.c_void_p), max_atoms,
                                 words.ctypes.data_as(D.I), C.byref(nbest), C.byref(score), C.byref(status),
                                 tr.ctypes.data_as(D.I) if trace else None)
    if n < 0:
        raise RuntimeError(f"oracle_beam_decode failed: {n}")
    out = dict(atoms=atoms[:n].copy(), words=words[:nbest.value].tolist(), score=score.value, status=status.value)
    if trace:
        out["trace"] = tr
    return out


# ------------------------------------------------------------------ compiled reference
def have_ref() -> bool:
    return os.path.exists(JREF)


def run_ref(workdir: str, filelist: list, extra_args: list = (), dump: str = "out.jrf", export: str | None = None,
            tokens: bool = False, am_args: list | None = None, quiet: bool = True, timeout: int = 3600,
            binary: str | None = None, env_extra: dict | None = None, two_pass: bool = False,
            outprobout: bool = True, lm_args: list | None = None):
    """Run the compiled reference on HTK parameter files; returns (dump path, stdout).
    two_pass: also run the stack-decoding pass 2 and print its sentences (JREF_RESULT lines).
    outprobout: pass "-outprobout /dev/null", which makes the host evaluate the complete [T x S] score
    matrix (what the dump's outprob section needs); it requires a populated score cache.
    lm_args: language-model options instead of the N-gram default, e.g. ["-dfa", "g.dfa", "-v", "g.dict"]."""
    env = dict(os.environ)
    if quiet:
        env["JREF_QUIET"] = "1"
    if tokens:
        env["JREF_TOKENS"] = "1"
    if export:
        env["JB200_EXPORT"] = export
    if two_pass:
        env["JREF_RESULT"] = "1"
    if env_extra:
        env.update(env_extra)
    args = [binary or JREF, "-dump", os.path.join(workdir, dump), "-plugindir", PLUGDIR]
    args += am_args if am_args is not None else ["-h", "hmmdefs", "-hlist", "hmmlist"]
    args += (lm_args if lm_args is not None else ["-v", "dict", "-nlr", "lm.arpa"]) + ["-input", "mfcfile"] + ([] if two_pass else ["-1pass"]) + (["-outprobout", "/dev/null"] if outprobout else [])
    args += list(extra_args)
    p = subprocess.run(args, input="\n".join(filelist) + "\n", text=True, cwd=workdir, env=env,
                       capture_output=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"jref failed ({p.returncode}): {p.stdout[-2000:]} {p.stderr[-2000:]}")
    return os.path.join(workdir, dump), p.stdout
