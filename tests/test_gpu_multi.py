"""Multi-GPU: ray sharding and the peer-memory hit gather (trq_trace_gather: the traversal kernel stores every finished
record into every rank's buffer over NVLink as the ray retires, then publishes (count, step)).

The gather tests run with one process per rank. On a box with one GPU both ranks share device 0 (CUDA IPC works between
processes on one device; the rendezvous is gloo because NCCL refuses two ranks on one GPU), so the driver's single-GPU
test run exercises the whole protocol; with >= 2 GPUs the same worker also runs one rank per GPU over NCCL."""
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
shared = os.environ["TRQ_SHARE_DEVICE"] == "1"
rank, local_rank, world = D.init(backend="gloo" if shared else None)
local_rank = 0 if shared else local_rank
torch.cuda.set_device(local_rank)
dev = f"cuda:{local_rank}"
prim = H.scene_soup(200_000, seed=1, extent=0.01) if rank == 0 else None      # triangle-only scene (C3 / soup shape)
prim = D.replicate_primitive(prim, src=0)
scene = Scene(prim, local_rank)
cap = 300_000
# Round-1 regression: hundreds of plain launches BEFORE the first gather (more than the old ring of 256 queue heads), so
# that a head left dirty by any earlier launch would make the gather's trace skip rays.
warm = rays_to_torch(H.random_rays(20_000, seed=99), dev)
tmp = torch.empty((20_000, 8), dtype=torch.float32, device=dev)
for k in range(320):
    scene.hit(warm, any=bool(k & 1), out=tmp)
g = D.HitGather(scene, cap)
# step sizes differ per rank and per step (ragged shards, an empty one, a full one); 7 steps cycle the three phases twice
sizes = lambda step: [[cap, 1000 + 777 * r, 0 if r == 1 else 50_000, cap - r, 12345, 70_000, 256][step] for r in range(world)]
prev = None
for step in range(7):
    ns = sizes(step)
    any_hit, h16 = step == 3, step in (4, 5)
    mine = H.random_rays(ns[rank], seed=10 + step, first=sum(ns[:rank]))
    g.trace(rays_to_torch(mine, dev) if ns[rank] else torch.empty((0, 8), dtype=torch.float32, device=dev), any=any_hit, hit16=h16)
    hits_all, counts = g.wait()
    torch.cuda.synchronize(); g.status()
    assert counts.tolist() == ns, (counts.tolist(), ns)
    for r in range(world):                                        # every rank re-traces every shard locally and compares bytes
        if ns[r] == 0:
            continue
        theirs = H.random_rays(ns[r], seed=10 + step, first=sum(ns[:r]))
        want = scene.hit(rays_to_torch(theirs, dev), any=any_hit, hit16=h16)
        assert torch.equal(hits_all[r, :ns[r]].contiguous().view(torch.int32), want.view(torch.int32)), f"step {step}: slot {r} on rank {rank}"
    if prev is not None:                                          # the previous step's result is still intact (three phases)
        p_all, p_want = prev
        for r in range(world):
            assert torch.equal(p_all[r, :p_want[r].shape[0]].contiguous().view(torch.int32), p_want[r].view(torch.int32)), f"step {step}: previous slot {r}"
    prev = (hits_all, [scene.hit(rays_to_torch(H.random_rays(ns[r], seed=10 + step, first=sum(ns[:r])), dev), any=any_hit, hit16=h16)
                       if ns[r] else torch.empty((0, 4 if h16 else 8), device=dev) for r in range(world)])
    if step == 0 and not shared:                                  # and the NCCL gather of the same records agrees
        parts = D.gather_hits(scene.hit(rays_to_torch(mine, dev)))
        for r in range(world):
            assert torch.equal(parts[r].view(torch.int32), hits_all[r, :ns[r]].view(torch.int32))
    D.barrier()                                                   # nobody starts step k+1 before everybody checked step k-1
g.close()
print("rank", rank, "ok")
'''


def _run(tmp_path, world, shared):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TRQ_ROOT=ROOT, TRQ_SHARE_DEVICE="1" if shared else "0")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = p.stdout.decode()
    assert p.returncode == 0, out[-4000:]
    for r in range(world):
        assert f"rank {r} ok" in out


def test_peer_memory_gather_two_ranks_on_one_gpu(built, tmp_path):
    """Runs everywhere there is a GPU: two processes share device 0."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _run(tmp_path, 2, shared=True)


def test_peer_memory_gather_more_ranks(built, tmp_path):
    """One rank per GPU over NCCL where the box has several GPUs (up to 4); on a one-GPU box three ranks share the device,
    so that a world size with an odd number of peers is exercised everywhere."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.device_count() >= 2:
        _run(tmp_path, min(torch.cuda.device_count(), 4), shared=False)
    else:
        _run(tmp_path, 3, shared=True)


def test_single_process_multi_gpu_helper(built):
    """trq_mgpu_*: one process, one scene per device, contiguous host-ray shards; identical bytes to a single-device trace
    for ragged batch sizes (including batches smaller than the device count). With one GPU the helper degenerates to one
    shard, which still goes through the same code."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tracer_b200 import MultiGpuScene, Scene, harness as H
    prim = H.scene_soup(100_000, seed=3, extent=0.02)
    one = Scene(prim, 0)
    m = MultiGpuScene(prim)
    assert m.n_devices == torch.cuda.device_count()
    assert m.shard(10, 0)[0] == 0 and m.shard(10, m.n_devices - 1)[1] == 10
    for n, any_hit in ((1_000_003, False), (257, True), (1, False), (max(1, m.n_devices - 1), False)):
        rays = H.random_rays(n, seed=40 + n % 7)
        want = one.hit(rays, any=any_hit)
        got = m.hit(rays, any=any_hit)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), n
        from tracer_b200 import layout as L
        small = m.hit(rays, any=any_hit, hit16=True)                  # 16-byte records: shard offsets in record units
        assert np.array_equal(small.view(np.uint8), L.pack_hit16(want).view(np.uint8)), n
    m.close(); one.close()
