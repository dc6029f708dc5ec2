"""Multi-GPU plumbing for the ray query: one process per GPU over torch.distributed.

The path shards by ray batch: the scene is read-only and replicated, rays are
independent, so there is NO collective on the data path. The only communication is
  * once per scene: broadcast of the six reference-layout arrays from rank 0 (NCCL over NVLink when the
    backend is nccl; gloo on CPU for the tests), after which every rank calls trq_scene_create;
  * optionally, after tracing: all-gather of the 32-byte hit records for a consumer that wants them whole --
    either NCCL (gather_hits / trace_and_gather) or, fused into the kernel that produces the records, stores into
    every rank's buffer over NVLink peer memory (HitGather -> trq_trace_gather).
"""
import os

import numpy as np

from . import layout as L
from .scene import Primitive

_FIELDS = (("sphereList", L.sphere_dtype), ("squareList", L.square_dtype), ("cubeList", L.cube_dtype),
           ("triList", L.vertex_dtype), ("idxList", np.dtype("<u4")), ("bvhList", L.bvh_dtype))


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n, rank, world):
    """Contiguous ray range [r*N/R, (r+1)*N/R) of rank r."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _comm_device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def replicate_primitive(prim, src=0):
    """Broadcast the six reference-layout arrays from rank `src`; returns a Primitive on every rank.
    `prim` is only read on the source rank (pass None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return prim
    dev = _comm_device()
    rank = dist.get_rank()
    sizes = torch.zeros(len(_FIELDS), dtype=torch.int64, device=dev)
    if rank == src:
        sizes = torch.tensor([getattr(prim, k).nbytes for k, _ in _FIELDS], dtype=torch.int64, device=dev)
    dist.broadcast(sizes, src)
    out = {}
    for (k, dt), nbytes in zip(_FIELDS, sizes.tolist()):
        if nbytes == 0:
            out[k] = np.zeros(0, dtype=dt)
            continue
        if rank == src:
            buf = torch.from_numpy(getattr(prim, k).view(np.uint8).reshape(-1)).to(dev)
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dist.broadcast(buf, src)
        out[k] = getattr(prim, k) if rank == src else buf.cpu().numpy().view(dt).reshape(-1).copy()
    return Primitive(**out)


def checksum_primitive(prim):
    """Order-sensitive 64-bit checksum of the scene bytes (to verify the broadcast)."""
    h = np.uint64(1469598103934665603)
    for k, _ in _FIELDS:
        a = getattr(prim, k).view(np.uint8).reshape(-1)
        pad = (-a.size) % 8
        w = np.concatenate([a, np.zeros(pad, dtype=np.uint8)]).view(np.uint64)
        with np.errstate(over="ignore"):
            h = h ^ (np.bitwise_xor.reduce(w * (np.arange(w.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1))) if w.size else np.uint64(0))
            h = h * np.uint64(1099511628211)
    return int(h)


def gather_hits(hits):
    """All-gather per-rank hit tensors ((n_r, 8) float32, possibly different n_r) -> list of tensors, rank order."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [hits]
    world = dist.get_world_size()
    dev = hits.device
    n = torch.tensor([hits.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    m = max(ns)
    padded = torch.zeros((m, 8), dtype=torch.float32, device=dev)
    padded[: hits.shape[0]] = hits
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded)
    return [o[:k] for o, k in zip(outs, ns)]


def max_over_ranks(value):
    """MAX all-reduce of a python float (device timing rule: multi-GPU time = max over ranks)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def trace_and_gather(scene, rays, hits_local, hits_all, any=False, sort=False, chunks=4, comm_stream=None):
    """Trace this rank's rays chunk by chunk and all-gather each chunk's 32-byte hit records while the next chunk is
    being traced ("chunked and overlapped with tracing on a second stream").

    rays (n, 8) and hits_local (n, 8) are this rank's CUDA tensors; hits_all is (world, n, 8): after the call
    hits_all[r] holds rank r's hits (every rank must pass the same n). Returns the last NCCL work handle (already waited).
    """
    import torch
    import torch.distributed as dist
    n = rays.shape[0]
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        scene.hit(rays, any=any, out=hits_local, sort=sort)
        hits_all[0].copy_(hits_local)
        return None
    comm_stream = comm_stream or torch.cuda.Stream(device=rays.device)
    bounds = [(n * c) // chunks for c in range(chunks + 1)]
    works = []
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi == lo:
            continue
        scene.hit(rays[lo:hi], any=any, out=hits_local[lo:hi], sort=sort)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ev)
            outs = [hits_all[r, lo:hi] for r in range(world)]
            works.append(dist.all_gather(outs, hits_local[lo:hi], async_op=True))
    for w in works:
        w.wait()
    torch.cuda.current_stream(rays.device).wait_stream(comm_stream)
    return works[-1] if works else None


def exchange_handles(handle):
    """All ranks contribute a bytes object; returns their concatenation in rank order (any backend)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(handle)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, bytes(handle))
    if any(len(p) != len(handle) for p in parts):
        raise RuntimeError("exchange_handles: ranks contributed handles of different sizes")
    return b"".join(parts)


class _DevArray:
    """Raw device memory as a CUDA array for torch.as_tensor (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class HitGather:
    """Hit all-gather overlapped with the traversal (trq_trace_gather): every rank's records land in every rank's
    (world, capacity, 8) buffer through NVLink peer stores, tile by tile while the trace runs. One process per GPU on one node
    (two ranks may share one device for tests: CUDA IPC works between processes on the same GPU).

        g = HitGather(scene, capacity)          # collective: allocates, exchanges CUDA-IPC handles, connects
        g.trace(rays); hits_all, counts = g.wait()      # every rank, every step; hits_all[r, :counts[r]]
        g.close()                               # collective (barrier inside)
    """

    def __init__(self, scene, capacity):
        import ctypes as C
        from ._lib import lib
        from ._lib import check
        rank, _, world = env_world()
        import torch.distributed as dist
        if dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        self.scene, self.rank, self.world, self.capacity = scene, rank, world, int(capacity)
        self._lib, self._check = lib, check
        self._h = C.c_void_p()
        handle = C.create_string_buffer(64)
        # every step below is collective: a rank that fails must not leave the others waiting, so the outcome of each
        # local call is agreed on (max over ranks) before anyone proceeds or raises
        rc = lib.trq_gather_create(scene._h, rank, world, self.capacity, C.byref(self._h), handle)
        self._agree(rc, "trq_gather_create")
        allh = exchange_handles(handle.raw)
        rc = lib.trq_gather_connect(self._h, allh)
        self._agree(rc, "trq_gather_connect")

    def _agree(self, rc, what):
        msg = self._lib.trq_last_error_string().decode() if rc != 0 else ""
        if max_over_ranks(float(rc != 0)) != 0.0:
            if self._h:
                self._lib.trq_gather_destroy(self._h)
                self._h = None
            raise RuntimeError(f"{what} failed on {'this rank: ' + msg if rc != 0 else 'another rank'}")

    def trace(self, rays, any=False, sort=False, stream=None, hit16=False):
        flags = (L.TRACE_ANY if any else 0) | (L.SORT_RAYS if sort else 0) | (L.HIT16 if hit16 else 0)
        self._hit16 = bool(hit16)
        self._check(self._lib.trq_trace_gather(self.scene._h, self._h, rays.data_ptr(), rays.shape[0], flags, self.scene._stream(stream)),
                    "trq_trace_gather")

    def wait(self, stream=None):
        """Enqueues the wait for every rank's records of the last trace(); returns ((world, capacity, 8) float32 view of the
        local buffer, (world,) int64 view of the per-rank counts). Both are valid until the next-but-one trace() executes
        (three buffer phases). After trace(hit16=True) the view is (world, capacity, 4): trq_hit16 rows."""
        import ctypes as C
        import torch
        hp, cp = C.c_void_p(), C.c_void_p()
        self._check(self._lib.trq_gather_wait(self._h, self.scene._stream(stream), C.byref(hp), C.byref(cp)), "trq_gather_wait")
        dev = torch.device("cuda", self.scene.device)
        hits = torch.as_tensor(_DevArray(hp.value, (self.world, self.capacity, 8), "<f4"), device=dev)
        if getattr(self, "_hit16", False):                       # 16-byte records at the front of each 32-byte-stride slot
            hits = hits.view(self.world, self.capacity * 2, 4)[:, : self.capacity]
        counts = torch.as_tensor(_DevArray(cp.value, (self.world,), "<i8"), device=dev)
        return hits, counts

    def status(self):
        self._check(self._lib.trq_gather_status(self._h), "trq_gather_status")

    def close(self):
        if self._h:
            import torch
            torch.cuda.synchronize(self.scene.device)
            barrier()
            self._lib.trq_gather_destroy(self._h)
            self._h = None
