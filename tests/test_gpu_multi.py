"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): ray sharding and the peer-memory hit gather
(trq_trace_gather: the resolve kernel stores every record into every rank's buffer over NVLink)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["TRQ_ROOT"])
from tracer_b200 import Scene, dist as D, harness as H, layout as L, hits_to_numpy, rays_to_torch
rank, local_rank, world = D.init()
dev = f"cuda:{local_rank}"
prim = H.scene_soup(200_000, seed=1, extent=0.01) if rank == 0 else None
prim = D.replicate_primitive(prim, src=0)
scene = Scene(prim, local_rank)
cap = 300_000
g = D.HitGather(scene, cap)
# step sizes differ per rank and per step (ragged shards, an empty one, a full one); 5 steps exercise both parities
sizes = lambda step: [[cap, 1000 + 777 * r, 0 if r == 1 else 50_000, cap - r, 12345][step] for r in range(world)]
for step in range(5):
    ns = sizes(step)
    mine = H.random_rays(ns[rank], seed=10 + step, first=sum(ns[:rank]))
    g.trace(rays_to_torch(mine, dev) if ns[rank] else torch.empty((0, 8), dtype=torch.float32, device=dev), any=(step == 3))
    hits_all, counts = g.wait()
    torch.cuda.synchronize(); g.status()
    assert counts.tolist() == ns, (counts.tolist(), ns)
    for r in range(world):                                        # every rank re-traces every shard locally and compares bytes
        if ns[r] == 0:
            continue
        theirs = H.random_rays(ns[r], seed=10 + step, first=sum(ns[:r]))
        want = scene.hit(rays_to_torch(theirs, dev), any=(step == 3))
        assert torch.equal(hits_all[r, :ns[r]].view(torch.int32), want.view(torch.int32)), f"step {step}: slot {r} on rank {rank}"
    if step == 0:                                                 # and the NCCL gather of the same records agrees
        parts = D.gather_hits(scene.hit(rays_to_torch(mine, dev)))
        for r in range(world):
            assert torch.equal(parts[r].view(torch.int32), hits_all[r, :ns[r]].view(torch.int32))
g.close()
print("rank", rank, "ok")
'''


def test_peer_memory_gather_two_ranks(built, tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TRQ_ROOT=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0, out[-4000:]
    for r in range(world):
        assert f"rank {r} ok" in out


def test_single_process_multi_gpu_helper(built):
    """trq_mgpu_*: one process, one scene per device, contiguous host-ray shards; identical bytes to a single-device trace
    for ragged batch sizes (including batches smaller than the device count)."""
    import numpy as np
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tracer_b200 import MultiGpuScene, Scene, harness as H
    prim = H.scene_soup(100_000, seed=3, extent=0.02)
    one = Scene(prim, 0)
    m = MultiGpuScene(prim)
    assert m.n_devices == torch.cuda.device_count()
    assert [m.shard(10, k) for k in range(2)][0][0] == 0 and m.shard(10, m.n_devices - 1)[1] == 10
    for n, any_hit in ((1_000_003, False), (257, True), (1, False), (m.n_devices - 1, False)):
        rays = H.random_rays(n, seed=40 + n % 7)
        want = one.hit(rays, any=any_hit)
        got = m.hit(rays, any=any_hit)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), n
    m.close(); one.close()
