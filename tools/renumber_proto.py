"""Prototype (analysis only, used by tools/node_locality.py): renumberings of the lexicon tree.
  bfs / depth_lm : keep every `next_a` chain (node, node+1, ...) consecutive -- on the synthetic 20k-word tree a chain is
                   a whole word, so these cannot separate the tree's levels and gain nothing;
  bfs_free       : plain breadth-first order from the roots (needs an explicit successor in the node record) -- what
                   jb200_decoder_create does (csrc/beam.cu)."""
from __future__ import annotations

import numpy as np

LOG_ZERO = -1000000.0
MODES = ("bfs", "depth_lm", "bfs_free")


def chains(blob):
    next_a = blob["tree.next_a"]
    n = len(next_a)
    linked = next_a != np.float32(LOG_ZERO)          # node -> node+1
    head = np.ones(n, bool)
    head[1:] = ~linked[:-1]
    heads = np.flatnonzero(head)
    chain_of = np.cumsum(head) - 1
    length = np.diff(np.append(heads, n))
    return heads, chain_of, length


def bfs_free(blob) -> np.ndarray:
    n = int(blob["tree.n_nodes"][0])
    next_a, arc_off, arc_to = blob["tree.next_a"], blob["tree.arc_off"], blob["tree.arc_to"]
    seen = np.zeros(n, bool)
    order = []
    def push(x):
        if not seen[x]:
            seen[x] = True; order.append(x)
    for r in list(blob["tree.iso_node"]) + list(blob["tree.shared_node"]):
        push(int(r))
    i = 0
    while i < len(order):
        x = order[i]; i += 1
        if next_a[x] != np.float32(LOG_ZERO) and x + 1 < n:
            push(x + 1)
        for k in arc_to[arc_off[x]:arc_off[x + 1]].tolist():
            push(k)
    for x in range(n):
        push(x)
    perm = np.empty(n, np.int32)
    perm[np.array(order)] = np.arange(n, dtype=np.int32)
    return perm


def permutation(blob, mode: str) -> np.ndarray:
    n = int(blob["tree.n_nodes"][0])
    if mode == "bfs_free":
        return bfs_free(blob)
    heads, chain_of, length = chains(blob)
    nc = len(heads)
    arc_off, arc_to = blob["tree.arc_off"], blob["tree.arc_to"]
    # chain graph
    kids = [[] for _ in range(nc)]
    src = np.repeat(np.arange(n), np.diff(arc_off))
    for s, d in zip(chain_of[src].tolist(), chain_of[arc_to].tolist()):
        if s != d:
            kids[s].append(d)
    roots = [int(chain_of[x]) for x in list(blob["tree.iso_node"]) + list(blob["tree.shared_node"])]
    depth = np.full(nc, -1, np.int64)
    order = []
    q = []
    for r in roots:
        if depth[r] < 0:
            depth[r] = 0; q.append(r)
    i = 0
    while i < len(q):
        c = q[i]; i += 1
        order.append(c)
        for k in kids[c]:
            if depth[k] < 0:
                depth[k] = depth[c] + 1; q.append(k)
    rest = [c for c in range(nc) if depth[c] < 0]
    for c in rest:
        depth[c] = 1 << 20
    order += rest
    if mode == "depth_lm":
        # best factoring value in the subtree of each chain (max over descendants), bottom-up over the BFS order
        scid = blob["tree.scid"]; fscore = blob["tree.fscore"]
        uni = blob["tree.uni_prob"]; wton = blob["tree.wton"]; scword = blob["tree.scword"]; cprob = blob["tree.cprob"]
        val = np.full(nc, -1e30)
        for node in range(n):
            s = int(scid[node])
            if s < 0:
                v = float(fscore[-s])
            elif s > 0:
                w = int(scword[s]); v = float(uni[wton[w]] + cprob[w])
            else:
                continue
            c = chain_of[node]
            if v > val[c]:
                val[c] = v
        sub = val.copy()
        for c in reversed(order):
            for k in kids[c]:
                if sub[k] > sub[c]:
                    sub[c] = sub[k]
        # inherit: chains without an own value take the parent's on the way down
        order = sorted(range(nc), key=lambda c: (depth[c], -sub[c]))
    perm = np.empty(n, np.int32)
    pos = 0
    for c in order:
        h, L = int(heads[c]), int(length[c])
        perm[h:h + L] = np.arange(pos, pos + L, dtype=np.int32)
        pos += L
    assert pos == n
    return perm
