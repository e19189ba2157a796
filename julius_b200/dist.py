"""Multi-GPU plumbing for the utterance-parallel hot path.

Utterances are independent, so the job shards by utterance with NO data-path collective; the only
collective is the init-time broadcast of the flattened model from the rank that loaded it.
Works with NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of item indices for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_by_length(lengths, world: int):
    """Length-balanced assignment (longest-processing-time first); returns a list of index lists."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    for r in range(world):
        out[r].sort()
    return out


def _pack(blob: dict):
    manifest = [(k, str(v.dtype), int(v.size)) for k, v in blob.items()]
    total = sum(np.dtype(dt).itemsize * n + (-(np.dtype(dt).itemsize * n) % 16) for _, dt, n in manifest)
    host = np.zeros(total, np.uint8)
    pos = 0
    for k, dt, n in manifest:
        raw = np.ascontiguousarray(blob[k]).tobytes()
        host[pos:pos + len(raw)] = np.frombuffer(raw, np.uint8)
        pos += len(raw) + (-len(raw) % 16)
    return manifest, host


def _unpack(manifest, host: np.ndarray) -> dict:
    out, pos = {}, 0
    for k, dt, n in manifest:
        nb = np.dtype(dt).itemsize * n
        out[k] = host[pos:pos + nb].view(dt).copy()
        pos += nb + (-nb % 16)
    return out


def broadcast_blob(blob, rank: int, world: int, device=None, src: int = 0) -> dict:
    """One broadcast of the flattened model (all arrays packed into a single byte tensor)."""
    if world == 1:
        return blob
    import torch
    import torch.distributed as dist
    meta = [None]
    host = None
    if rank == src:
        manifest, host = _pack(blob)
        meta[0] = manifest
    dist.broadcast_object_list(meta, src=src)
    manifest = meta[0]
    total = sum(np.dtype(dt).itemsize * n + (-(np.dtype(dt).itemsize * n) % 16) for _, dt, n in manifest)
    dev = device if device is not None else torch.device("cpu")
    flat = torch.empty(total, dtype=torch.uint8, device=dev)
    if rank == src:
        flat.copy_(torch.from_numpy(host))
    dist.broadcast(flat, src=src)
    return _unpack(manifest, flat.cpu().numpy())
