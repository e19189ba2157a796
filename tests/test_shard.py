"""CPU, world_size 2 over gloo: the N>1 host path -- model broadcast + utterance sharding."""
import os
import subprocess
import sys

import numpy as np

from julius_b200 import dist as jd
from util import ROOT

WORKER = r'''
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from julius_b200 import dist as jd, refdump
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
blob = refdump.load_blob(os.path.join(sys.argv[1], "tests", "golden", "tiny", "model.jb2m")) if rank == 0 else None
got = jd.broadcast_blob(blob, rank, world)
ref = refdump.load_blob(os.path.join(sys.argv[1], "tests", "golden", "tiny", "model.jb2m"))
same = list(got) == list(ref) and all(got[k].dtype == ref[k].dtype and np.array_equal(got[k], ref[k]) for k in ref)
b, e = jd.shard_range(11, rank, world)
mine = torch.tensor([e - b], dtype=torch.int64)
dist.all_reduce(mine)
print(json.dumps({"rank": rank, "same": bool(same), "range": [b, e], "total": int(mine.item())}))
dist.destroy_process_group()
'''


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 512, 4096):
        for w in (1, 2, 4, 8):
            got = [jd.shard_range(n, r, w) for r in range(w)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_length_balances():
    rng = np.random.default_rng(0)
    lens = rng.integers(100, 2000, size=97)
    parts = jd.shard_by_length(lens, 8)
    assert sorted(i for p in parts for i in p) == list(range(97))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= lens.max()


def test_broadcast_and_shard_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    import json
    res = []
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
        res.append(json.loads(o.strip().splitlines()[-1]))
    assert all(r["same"] for r in res)
    assert res[0]["range"] == [0, 6] and res[1]["range"] == [6, 11]
    assert all(r["total"] == 11 for r in res)
