"""Test tooling: how many 32-byte sectors / 128-byte lines of the per-node arrays (16-byte arrival slots, 32-byte node
records) does one frame of pass 1 touch, under the host's node numbering and under candidate renumberings of the tree?

    python tools/node_locality.py tri20k [n_utts] [n_frames]

Decodes a few synthetic utterances with the CPU restatement (oracle/), records the node of every token of every frame
(oracle_set_node_dump) and counts distinct sectors / lines per frame for each numbering.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.path.insert(0, os.path.join(ROOT, "tools"))
import renumber_proto as renumber
from julius_b200 import desc as D, refdump, workload
from oracle import ffi


def frames_of(path):
    a = np.fromfile(path, np.int32)
    out = []
    i = 0
    while i < len(a):
        n = a[i]
        out.append(a[i + 1:i + 1 + n])
        i += 1 + n
    return out


def touched(frames, perm, rec_bytes):
    """mean distinct 32-B sectors and 128-B lines per frame of an array of rec_bytes records indexed by perm[node]"""
    sec = lin = 0
    for f in frames:
        idx = perm[f].astype(np.int64) * rec_bytes
        sec += len(np.unique(idx >> 5))
        lin += len(np.unique(idx >> 7))
    return sec / len(frames), lin / len(frames)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tri20k"
    n_utts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    blob = refdump.load_blob(workload.path(name, "model.jb2m"))
    ds = D.Descriptors(blob)
    m = workload.synth_model(name)
    feats = workload.sample_inputs(name, m, n_utts, T, seed=4242)
    lib = ffi.lib()
    lib.oracle_set_node_dump.argtypes = [C.c_char_p]
    dump = "/tmp/node_dump.bin"
    lib.oracle_set_node_dump(dump.encode())
    for x in feats:
        st = ffi.dnn_score(ds, x) if workload.is_dnn(name) else ffi.gmm_score(ds, x)
        ffi.beam_decode(ds, st)
    lib.oracle_set_node_dump(b"")
    frames = frames_of(dump)
    n = int(blob["tree.n_nodes"][0])
    print(f"{name}: {n} nodes, {len(frames)} frames, {np.mean([len(f) for f in frames]):.0f} tokens per frame")
    perms = {"host": np.arange(n, dtype=np.int32)}
    for mode in renumber.MODES:
        perms[mode] = renumber.permutation(blob, mode)
    for k, perm in perms.items():
        assert sorted(perm.tolist()) == list(range(n)), k
        s16, l16 = touched(frames, perm, 16)
        s32, l32 = touched(frames, perm, 32)
        print(f"  {k:12s} slots(16 B): {s16:7.0f} sectors {l16:7.0f} lines   node records(32 B): {s32:7.0f} sectors {l32:7.0f} lines")


if __name__ == "__main__":
    main()
