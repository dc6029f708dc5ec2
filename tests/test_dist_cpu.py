"""CPU: the multi-process path (world_size 2, gloo): scene broadcast, ray sharding, hit gather, max-over-ranks."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["TRQ_ROOT"])
from tracer_b200 import dist as D, harness as H, layout as L
rank, local_rank, world = D.init(backend="gloo")
assert world == 2
prim = H.scene_soup(300, seed=1, extent=0.2) if rank == 0 else None
prim = D.replicate_primitive(prim, src=0)
want = H.scene_soup(300, seed=1, extent=0.2)
for k in ("triList", "idxList", "bvhList"):
    assert np.array_equal(getattr(prim, k).view(np.uint8), getattr(want, k).view(np.uint8)), k
cs = D.checksum_primitive(prim)
assert D.max_over_ranks(float(cs % (1 << 52))) == -D.max_over_ranks(-float(cs % (1 << 52)))
n = 1001
lo, hi = D.shard_range(n, rank, world)
assert (lo, hi) == ((0, 500) if rank == 0 else (500, 1001))
rays = H.random_rays(hi - lo, seed=2, first=lo)                    # each rank generates exactly its shard
fake_hits = torch.from_numpy(rays.view(np.float32).reshape(-1, 8).copy())   # stand-in for per-rank hit tensors
parts = D.gather_hits(fake_hits)
assert [p.shape[0] for p in parts] == [500, 501]
whole = H.random_rays(n, seed=2)
assert np.array_equal(torch.cat(parts).numpy().view(np.uint8), whole.view(np.float32).reshape(-1, 8).view(np.uint8))
assert D.max_over_ranks(float(rank + 1)) == 2.0 and D.sum_over_ranks(float(hi - lo)) == float(n)
# the handle exchange of the peer-memory gather (HitGather): rank order, any backend
allh = D.exchange_handles(bytes([rank + 1]) * 64)
assert allh == bytes([1]) * 64 + bytes([2]) * 64
D.barrier()
print("rank", rank, "ok")
'''


def test_two_ranks_gloo(built, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), TRQ_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o}"
        assert f"rank {rank} ok" in o
