"""Benchmark workloads (BASELINE.json configs) as flattened model blobs + seeded synthetic input.

A workload directory  workloads/<name>/  holds
    model.jb2m                      flattened model (what the GPU path loads)
    hmmdefs hmmlist dict lm.arpa    the same model in the reference's file formats (CPU baseline leg)
    meta.json
The model text files come from julius_b200.synth; the blob is produced by letting the HOST
(Julius itself, i.e. the reference build under oracle/_ref, with our export plugin loaded) read
them -- exactly what happens in deployment, where startup(Recog*) flattens the live models.
That step needs the compiled reference and is therefore done by __graft_entry__.build() in the
build container; the GPU box only reads the prepared files.
"""
from __future__ import annotations

import json
import os
import subprocess
import time

import numpy as np

from . import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WDIR = os.path.join(ROOT, "workloads")

WORKLOADS = {
    # name: (synth preset, jconf-style options)
    "mono100": ("mono100", []),                       # BASELINE configs[0]
    "tri20k": ("tri20k", []),                         # configs[1]: 3k states x 16 mix, 20k words, beam 800 (auto)
    "tri20k_gbeam": ("tri20k", ["-gprune", "beam"]),  # configs[2]: same with -gprune beam
    "tri20k_mp": ("tri20k", ["-multipath"]),          # configs[1] with the multipath tree (non-emitting word begin/end nodes)
    "dnn20k": ("tri20k", ["-dnnconf", "@DNN@"]),      # configs[3]: DNN-HMM 528 -> 7x2048 -> 3000, 20k words
}
# DNN shapes (BASELINE configs[3]: ENVR-v5.4 shape 7x2048 sigmoid, 48x11 input)
DNN_SHAPES = {"dnn20k": dict(in_dim=528, feature_len=48, context_len=11, hidden=2048, layers=7, seed=9)}


def path(name: str, *parts) -> str:
    return os.path.join(WDIR, name, *parts)


def model_dir(name: str) -> str:
    """text model files are shared between workloads of the same preset"""
    return path(WORKLOADS[name][0] if name in WORKLOADS else name)


def ready(name: str) -> bool:
    return os.path.exists(path(name, "model.jb2m")) and os.path.exists(path(name, "meta.json"))


def synth_model(name: str) -> synth.SynthModel:
    return synth.SynthModel(synth.SynthConfig.preset(WORKLOADS[name][0]))


def ensure(name: str, verbose: bool = False) -> bool:
    """Create the workload if the compiled reference is available; returns ready(name)."""
    if ready(name):
        return True
    jref = os.path.join(ROOT, "oracle", "_ref", "jref")
    if not os.path.exists(jref):
        return False
    preset, opts = WORKLOADS[name]
    t0 = time.time()
    mdir = model_dir(name)
    m = synth.SynthModel(synth.SynthConfig.preset(preset))
    if not os.path.exists(os.path.join(mdir, "lm.arpa")):
        m.write_all(mdir)
    os.makedirs(path(name), exist_ok=True)
    is_dnn = name in DNN_SHAPES
    if is_dnn:
        dc = synth.DnnConfig(**DNN_SHAPES[name])
        if not os.path.exists(path(name, "dnnconf")):
            synth.write_dnn(path(name), m.cfg.n_states, dc)
        opts = [o if o != "@DNN@" else path(name, "dnnconf") for o in opts]
    # one short utterance is enough to make the host load everything and call startup()
    rng = np.random.default_rng(5)
    fn = path(name, "probe.mfc")
    if is_dnn:
        synth.write_htk_param(fn, synth.sample_dnn_input(rng, 60, dc.in_dim), parmkind=synth.PARMKIND_USER)
    else:
        synth.write_htk_param(fn, m.sample_utterance(rng, 60)[0])
    env = dict(os.environ, JREF_QUIET="1", JB200_EXPORT=path(name, "model.jb2m"))
    args = [jref, "-dump", path(name, "probe.jrf"), "-plugindir", os.path.join(ROOT, "oracle", "_ref"),
            "-h", os.path.join(mdir, "hmmdefs"), "-hlist", os.path.join(mdir, "hmmlist"),
            "-v", os.path.join(mdir, "dict"), "-nlr", os.path.join(mdir, "lm.arpa"),
            "-input", "mfcfile", "-1pass"] + opts
    p = subprocess.run(args, input=fn + "\n", text=True, capture_output=True, env=env)
    if p.returncode != 0 or not os.path.exists(path(name, "model.jb2m")):
        raise RuntimeError(f"workload {name}: host run failed: {p.stdout[-1000:]} {p.stderr[-1000:]}")
    with open(path(name, "meta.json"), "w") as f:
        json.dump({"preset": preset, "options": opts, "model_dir": os.path.relpath(mdir, ROOT),
                   "built_sec": round(time.time() - t0, 1)}, f, indent=1)
    if verbose:
        print(f"workload {name}: built in {time.time() - t0:.1f}s")
    return True


def ensure_all(verbose: bool = False) -> None:
    for name in WORKLOADS:
        try:
            ensure(name, verbose=verbose)
        except Exception as e:   # a missing workload is reported by bench.py when it is asked for
            print(f"workload {name}: {e}")


def ref_args(name: str) -> list:
    """jconf-style options for running the reference on this workload."""
    preset, opts = WORKLOADS[name]
    mdir = model_dir(name)
    opts = [o if o != "@DNN@" else path(name, "dnnconf") for o in opts]
    return ["-h", os.path.join(mdir, "hmmdefs"), "-hlist", os.path.join(mdir, "hmmlist"),
            "-v", os.path.join(mdir, "dict"), "-nlr", os.path.join(mdir, "lm.arpa"),
            "-input", "mfcfile", "-1pass"] + opts


def is_dnn(name: str) -> bool:
    return name in DNN_SHAPES


def sample_inputs(name: str, m: synth.SynthModel, n_utts: int, n_frames: int, seed: int):
    """Feature matrices in the layout the workload's acoustic model takes."""
    if name in DNN_SHAPES:
        rng = np.random.default_rng(seed)
        return [synth.sample_dnn_input(rng, n_frames, DNN_SHAPES[name]["in_dim"]) for _ in range(n_utts)]
    return sample_batch(m, n_utts, n_frames, seed)


def write_input(name: str, fn: str, x) -> None:
    synth.write_htk_param(fn, x, parmkind=synth.PARMKIND_USER if name in DNN_SHAPES else synth.PARMKIND_MFCC_E_D_A)


def sample_batch(m: synth.SynthModel, n_utts: int, n_frames: int, seed: int):
    rng = np.random.default_rng(seed)
    return [m.sample_utterance(rng, n_frames)[0] for _ in range(n_utts)]
