"""Readers for the harness file formats: JB2M model blobs (include/jb200_model.h) and
JRF1 reference dumps (oracle/ref_driver.c).  Pure numpy; no reference, no oracle."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

_DT = {0: np.float32, 1: np.int32, 2: np.uint8}


def load_blob(path: str) -> dict:
    """Read a JB2M container into {name: ndarray}; 1-element arrays stay arrays."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"JB2M":
        raise ValueError(f"{path}: not a JB2M blob")
    ver, n, _ = struct.unpack_from("<iii", data, 4)
    if ver != 1:
        raise ValueError(f"{path}: unsupported blob version {ver}")
    pos = 16
    for _ in range(n):
        name = data[pos:pos + 48].split(b"\0", 1)[0].decode()
        dtype, _pad, count = struct.unpack_from("<iiq", data, pos + 48)
        pos += 64
        dt = np.dtype(_DT[dtype])
        nbytes = count * dt.itemsize
        out[name] = np.frombuffer(data, dtype=dt, count=count, offset=pos).copy()
        pos += nbytes + ((16 - nbytes % 16) % 16)
    return out


def save_blob(path: str, arrays: dict) -> None:
    inv = {np.dtype(np.float32): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2}
    with open(path, "wb") as f:
        f.write(b"JB2M")
        f.write(struct.pack("<iii", 1, len(arrays), 0))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()[:47]
            f.write(nb + b"\0" * (48 - len(nb)))
            f.write(struct.pack("<iiq", inv[a.dtype], 0, a.size))
            raw = a.tobytes()
            f.write(raw)
            if len(raw) % 16:
                f.write(b"\0" * (16 - len(raw) % 16))


def scalar(blob: dict, name: str, default=None):
    if name not in blob:
        return default
    return blob[name][0].item()


@dataclass
class RefUtterance:
    index: int
    n_frames: int
    decode_sec: float
    outprob: np.ndarray | None          # [T, S] float32 or None
    atoms: np.ndarray                   # structured: wid, begin, end, backscore, lscore, last
    status: int
    words: list                         # pass-1 best, in the reference's stored (reverse) order
    score: float
    tokens: list = field(default_factory=list)   # per frame: structured arrays of survivors

ATOM_DT = np.dtype([("wid", "<i4"), ("begin", "<i4"), ("end", "<i4"),
                    ("backscore", "<f4"), ("lscore", "<f4"), ("last", "<i4")])
TOKEN_DT = np.dtype([("node", "<i4"), ("score", "<f4"), ("tre_wid", "<i4"), ("tre_end", "<i4"),
                     ("cword", "<i4"), ("lscore", "<f4")])


def load_refdump(path: str) -> list:
    """Parse a JRF1 dump written by oracle/_ref/jref (one record per utterance)."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    utts = []
    pending_tokens = []
    while pos < len(data):
        (tag,) = struct.unpack_from("<i", data, pos)
        pos += 4
        if tag == 0x544F4B31:       # TOK1
            frame, tnum, nsurv = struct.unpack_from("<iii", data, pos)
            pos += 12
            arr = np.frombuffer(data, dtype=TOKEN_DT, count=nsurv, offset=pos).copy()
            pos += nsurv * TOKEN_DT.itemsize
            pending_tokens.append((frame, tnum, arr))
        elif tag == 0x4A524631:     # JRF1
            idx, T = struct.unpack_from("<ii", data, pos)
            pos += 8
            (dt,) = struct.unpack_from("<f", data, pos)
            pos += 4
            (S,) = struct.unpack_from("<i", data, pos)
            pos += 4
            outprob = None
            if S > 0:
                outprob = np.frombuffer(data, dtype="<f4", count=T * S, offset=pos).reshape(T, S).copy()
                pos += T * S * 4
            (n,) = struct.unpack_from("<i", data, pos)
            pos += 4
            atoms = np.frombuffer(data, dtype=ATOM_DT, count=n, offset=pos).copy()
            pos += n * ATOM_DT.itemsize
            status, wnum = struct.unpack_from("<ii", data, pos)
            pos += 8
            words = list(struct.unpack_from(f"<{wnum}i", data, pos)) if wnum else []
            pos += 4 * wnum
            (score,) = struct.unpack_from("<f", data, pos)
            pos += 4
            utts.append(RefUtterance(idx, T, dt, outprob, atoms, status, words, score, pending_tokens))
            pending_tokens = []
        else:
            raise ValueError(f"{path}: bad record tag {tag:#x} at {pos - 4}")
    return utts
