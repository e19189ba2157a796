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
import shutil
import subprocess
import tempfile
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
    # configs[4]: DNN-HMM, 60k-word tree, the ENVR-v5.4 search options (/root/reference/README.md:120-144) minus -no_ccd:
    # without cross-word context handling the host wants boundary biphone/monophone models the synthetic triphone set
    # does not define ("CDSET phoneme exist in monophone?"), and with it -iwcd1 max is actually exercised
    "dnn60k_mp": ("tri60k", ["-dnnconf", "@DNN@", "-multipath", "-iwsp", "-iwcd1", "max", "-b", "4000"]),
}
# DNN shapes (BASELINE configs[3]: ENVR-v5.4 shape 7x2048 sigmoid, 48x11 input)
_ENVR = dict(in_dim=528, feature_len=48, context_len=11, hidden=2048, layers=7, seed=9, w_scale=8.0, prototype_output=True,
             cache_dir=os.path.join(WDIR, "_dnn"))
DNN_SHAPES = {"dnn20k": dict(_ENVR), "dnn60k_mp": dict(_ENVR)}


def path(name: str, *parts) -> str:
    return os.path.join(WDIR, name, *parts)


def model_dir(name: str) -> str:
    """text model files are shared between workloads of the same preset"""
    return path(WORKLOADS[name][0] if name in WORKLOADS else name)


def ready(name: str) -> bool:
    return os.path.exists(path(name, "model.jb2m")) and os.path.exists(path(name, "meta.json"))


def synth_model(name: str) -> synth.SynthModel:
    return synth.SynthModel(synth.SynthConfig.preset(WORKLOADS[name][0]))


def dnn_config(name: str) -> synth.DnnConfig:
    return synth.DnnConfig(**DNN_SHAPES[name])


def load_model(name: str) -> dict:
    """The flattened model of a workload as a blob dict.  DNN workloads keep only the tree / state layout on disk
    (the 130 MB of random-init weights are a pure function of the seed): the dnn.* entries are regenerated here,
    bit-identical to what the export plugin produces from the reference's loader (checked when the workload is built)."""
    from . import refdump
    blob = refdump.load_blob(path(name, "model.jb2m"))
    if name in DNN_SHAPES and "dnn.n_layers" not in blob:
        blob.update(synth.dnn_blob_entries(int(blob["gmm.n_states"][0]), dnn_config(name)))
    return blob


def ensure(name: str, verbose: bool = False) -> bool:
    """Create the workload if the compiled reference is available; returns ready(name)."""
    if ready(name):
        return True
    jref = os.path.join(ROOT, "oracle", "_ref", "jref")
    if not os.path.exists(jref):
        return False
    from . import refdump
    preset, opts = WORKLOADS[name]
    t0 = time.time()
    mdir = model_dir(name)
    m = synth.SynthModel(synth.SynthConfig.preset(preset))
    if not os.path.exists(os.path.join(mdir, "lm.arpa")):
        m.write_all(mdir)
    os.makedirs(path(name), exist_ok=True)
    is_dnn = name in DNN_SHAPES
    tmp = None
    if is_dnn:
        dc = dnn_config(name)
        tmp = tempfile.mkdtemp(prefix="jb200_dnn_")
        synth.write_dnn(tmp, m.cfg.n_states, dc)
        opts = [o if o != "@DNN@" else os.path.join(tmp, "dnnconf") for o in opts]
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
    if tmp:
        shutil.rmtree(tmp, ignore_errors=True)
    if p.returncode != 0 or not os.path.exists(path(name, "model.jb2m")):
        raise RuntimeError(f"workload {name}: host run failed: {p.stdout[-1000:]} {p.stderr[-1000:]}")
    if is_dnn:
        # keep the tree / state layout only; the weights are regenerated by load_model -- after checking that the
        # regenerated entries are what the host exported, bit for bit
        blob = refdump.load_blob(path(name, "model.jb2m"))
        gen = synth.dnn_blob_entries(int(blob["gmm.n_states"][0]), dc)
        for k, v in gen.items():
            if k not in blob or blob[k].dtype != v.dtype or not np.array_equal(blob[k].view(np.uint8), v.view(np.uint8)):
                raise RuntimeError(f"workload {name}: regenerated DNN entry {k} differs from the host's export")
        refdump.save_blob(path(name, "model.jb2m"), {k: v for k, v in blob.items() if not k.startswith("dnn.")})
    with open(path(name, "meta.json"), "w") as f:
        json.dump({"preset": preset, "options": WORKLOADS[name][1], "model_dir": os.path.relpath(mdir, ROOT),
                   "dnn_weights": "regenerated from the seed by workload.load_model" if is_dnn else None,
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


_DNN_TMP = {}


def ref_args(name: str) -> list:
    """jconf-style options for running the reference on this workload (DNN workloads: the .npy / dnnconf files the
    reference reads are written to a temporary directory on first use)."""
    preset, opts = WORKLOADS[name]
    mdir = model_dir(name)
    if name in DNN_SHAPES:
        if name not in _DNN_TMP:
            d = tempfile.mkdtemp(prefix="jb200_dnn_")
            synth.write_dnn(d, synth.SynthConfig.preset(preset).n_states, dnn_config(name))
            _DNN_TMP[name] = d
        opts = [o if o != "@DNN@" else os.path.join(_DNN_TMP[name], "dnnconf") for o in opts]
    return ["-h", os.path.join(mdir, "hmmdefs"), "-hlist", os.path.join(mdir, "hmmlist"),
            "-v", os.path.join(mdir, "dict"), "-nlr", os.path.join(mdir, "lm.arpa"),
            "-input", "mfcfile", "-1pass"] + opts


def is_dnn(name: str) -> bool:
    return name in DNN_SHAPES


def dnn_input_table(name: str) -> np.ndarray:
    """[n_states, in_dim] the per-state prototype inputs of a DNN workload (synth.DnnConfig.prototype_output)."""
    return synth.dnn_arrays(synth.SynthConfig.preset(WORKLOADS[name][0]).n_states, dnn_config(name))["proto"]


def sample_inputs(name: str, m: synth.SynthModel, n_utts: int, n_frames: int, seed: int):
    """Feature matrices in the layout the workload's acoustic model takes: GMM workloads draw MFCC-like frames from the
    Gaussians along a random <s> w.. </s> path; DNN workloads emit, along the same kind of path, the input vector under
    which the network favours the path's state (the state's prototype, synth.DnnConfig) plus a little jitter."""
    if name in DNN_SHAPES:
        rng = np.random.default_rng(seed)
        table = dnn_input_table(name)
        out = []
        for _ in range(n_utts):
            st, _words = m.sample_state_path(rng, n_frames)
            out.append((table[st] + np.float32(0.05) * rng.standard_normal((n_frames, table.shape[1]), dtype=np.float32)).astype(np.float32))
        return out
    return sample_batch(m, n_utts, n_frames, seed)


def write_input(name: str, fn: str, x) -> None:
    synth.write_htk_param(fn, x, parmkind=synth.PARMKIND_USER if name in DNN_SHAPES else synth.PARMKIND_MFCC_E_D_A)


def sample_batch(m: synth.SynthModel, n_utts: int, n_frames: int, seed: int):
    rng = np.random.default_rng(seed)
    return [m.sample_utterance(rng, n_frames)[0] for _ in range(n_utts)]
