"""Host-side mirror of the reference's ray-query interface, over the C-ABI.

  reference                                             here
  ----------------------------------------------------  ---------------------------------
  struct Primitive  (Render.hh:122-130)                 Primitive (six numpy arrays, reference layouts)
  BVH::buildNode / BVH::buildTree (BVH.hh:246-314)      BVHBuilder.buildNode / buildNodesTriangles / buildTree
  Scene { primitives }; scene.hit(ray, rec, test_t, any) (Render.hh:132-252)
                                                        Scene(primitives).hit(rays, any=False)  -- a BATCH of rays

Rays / hits live in device memory as torch tensors of shape (n, 8) float32 (32-byte trq_ray /
trq_hit rows), or in host memory as numpy structured arrays (`ray_dtype` / `hit_dtype`), in which
case the library stages the copies itself (TRQ_HOST_PTRS). All compute happens in the CUDA
kernels behind trq_trace; nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import layout as L
from ._lib import Camera, SceneDesc, SceneInfo, check, lib


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data


class Primitive:
    """The six arrays of `struct Primitive` (Render.hh:122-130), reference byte layouts."""

    def __init__(self, sphereList=None, squareList=None, cubeList=None, triList=None, idxList=None, bvhList=None):
        def arr(a, dt):
            if a is None:
                return np.zeros(0, dtype=dt)
            a = np.ascontiguousarray(a)
            if a.dtype != dt:
                raise TypeError(f"expected dtype {dt}, got {a.dtype}")
            return a
        self.sphereList = arr(sphereList, L.sphere_dtype)
        self.squareList = arr(squareList, L.square_dtype)
        self.cubeList = arr(cubeList, L.cube_dtype)
        self.triList = arr(triList, L.vertex_dtype)
        self.idxList = arr(None if idxList is None else np.asarray(idxList).reshape(-1), np.dtype("<u4"))
        self.bvhList = arr(bvhList, L.bvh_dtype)

    @property
    def nTri(self):
        return self.idxList.size // 3

    def desc(self):
        d = SceneDesc()
        d.sphereList, d.nSphere = _ptr(self.sphereList), self.sphereList.size
        d.squareList, d.nSquare = _ptr(self.squareList), self.squareList.size
        d.cubeList, d.nCube = _ptr(self.cubeList), self.cubeList.size
        d.triList, d.nVert = _ptr(self.triList), self.triList.size
        d.idxList, d.nTri = _ptr(self.idxList), self.nTri
        d.bvhList, d.nNode = _ptr(self.bvhList), self.bvhList.size
        return d

    def nbytes(self):
        return sum(a.nbytes for a in (self.sphereList, self.squareList, self.cubeList, self.triList, self.idxList, self.bvhList))


class DevicePrimitive:
    """`struct Primitive` whose six arrays already live on a GPU: torch uint8 CUDA tensors (or None) holding the reference
    byte layouts. Scene(DevicePrimitive) goes through trq_scene_create_device: nothing crosses PCIe."""

    _sizes = (("sphereList", L.sphere_dtype.itemsize), ("squareList", L.square_dtype.itemsize), ("cubeList", L.cube_dtype.itemsize),
              ("triList", L.vertex_dtype.itemsize), ("idxList", 12), ("bvhList", L.bvh_dtype.itemsize))

    def __init__(self, sphereList=None, squareList=None, cubeList=None, triList=None, idxList=None, bvhList=None):
        self.arrays = dict(sphereList=sphereList, squareList=squareList, cubeList=cubeList, triList=triList, idxList=idxList, bvhList=bvhList)
        for k, item in self._sizes:
            t = self.arrays[k]
            if t is not None and (not t.is_cuda or not t.is_contiguous() or t.numel() * t.element_size() % item):
                raise TypeError(f"{k}: expected a contiguous CUDA tensor holding whole {item}-byte records")

    @classmethod
    def from_host(cls, prim, device):
        """Uploads a host Primitive (for tests / for callers that build on the host but create on the device)."""
        import torch
        def up(a):
            return None if a.size == 0 else torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(device)
        return cls(up(prim.sphereList), up(prim.squareList), up(prim.cubeList), up(prim.triList), up(prim.idxList), up(prim.bvhList))

    def count(self, k):
        t = self.arrays[k]
        return 0 if t is None else t.numel() * t.element_size() // dict(self._sizes)[k]

    def desc(self):
        d = SceneDesc()
        ptr = lambda k: None if self.arrays[k] is None else self.arrays[k].data_ptr()
        d.sphereList, d.nSphere = ptr("sphereList"), self.count("sphereList")
        d.squareList, d.nSquare = ptr("squareList"), self.count("squareList")
        d.cubeList, d.nCube = ptr("cubeList"), self.count("cubeList")
        d.triList, d.nVert = ptr("triList"), self.count("triList")
        d.idxList, d.nTri = ptr("idxList"), self.count("idxList")
        d.bvhList, d.nNode = ptr("bvhList"), self.count("bvhList")
        return d


class BVHBuilder:
    """BVH::buildNode / BVH::buildTree (BVH.hh:246-314): leaves are appended, then the tree is built."""

    def __init__(self):
        self._chunks = []

    def leaves(self):
        return np.concatenate(self._chunks) if self._chunks else np.zeros(0, dtype=L.bvh_dtype)

    def buildNode(self, box_min, box_max, model_matrix, pType, pIndex):
        node = np.zeros(1, dtype=L.bvh_dtype)
        lo = np.ascontiguousarray(box_min, dtype=np.float32)
        hi = np.ascontiguousarray(box_max, dtype=np.float32)
        m = None if model_matrix is None else np.ascontiguousarray(model_matrix, dtype=np.float32).reshape(16)
        check(lib.trq_bvh_build_node(lo.ctypes.data, hi.ctypes.data, None if m is None else m.ctypes.data,
                                     int(pType), int(pIndex), node.ctypes.data), "trq_bvh_build_node")
        self._chunks.append(node)

    def buildNodesTriangles(self, triList, idxList, pIndexBase=0):
        idx = np.ascontiguousarray(np.asarray(idxList).reshape(-1), dtype="<u4")
        n = idx.size // 3
        nodes = np.zeros(n, dtype=L.bvh_dtype)
        if n:
            check(lib.trq_bvh_build_nodes_triangles(triList.ctypes.data, idx.ctypes.data, n, int(pIndexBase),
                                                    nodes.ctypes.data), "trq_bvh_build_nodes_triangles")
        self._chunks.append(nodes)

    def buildTree(self, gpu=None):
        """BVH::buildTree. gpu=None: host builder; gpu=<device index>: the same tree built on that GPU."""
        leaves = np.concatenate(self._chunks) if self._chunks else np.zeros(0, dtype=L.bvh_dtype)
        n = leaves.size
        if n == 0:
            raise ValueError("buildTree: no leaves")
        nodes = np.zeros(2 * n - 1, dtype=L.bvh_dtype)
        nodes[:n] = leaves
        nNode, depth = C.c_uint32(0), C.c_uint32(0)
        if gpu is None:
            check(lib.trq_bvh_build_tree(nodes.ctypes.data, n, C.byref(nNode), C.byref(depth)), "trq_bvh_build_tree")
        else:
            check(lib.trq_bvh_build_tree_gpu(nodes.ctypes.data, n, int(gpu), C.byref(nNode), C.byref(depth)), "trq_bvh_build_tree_gpu")
        self.maxDepth = depth.value
        return nodes[: nNode.value]

    def buildTreeDevice(self, device=0):
        """BVH::buildTree with the node array left ON the GPU (trq_bvh_build_tree_device): returns a torch uint8 CUDA
        tensor holding the 2n-1 64-byte nodes, ready for DevicePrimitive(bvhList=...)."""
        import torch
        leaves = np.concatenate(self._chunks) if self._chunks else np.zeros(0, dtype=L.bvh_dtype)
        n = leaves.size
        if n == 0:
            raise ValueError("buildTree: no leaves")
        d = torch.zeros((2 * n - 1) * L.bvh_dtype.itemsize, dtype=torch.uint8, device=f"cuda:{device}")
        d[: n * L.bvh_dtype.itemsize] = torch.from_numpy(leaves.view(np.uint8).reshape(-1)).to(d.device)
        nNode, depth = C.c_uint32(0), C.c_uint32(0)
        check(lib.trq_bvh_build_tree_device(d.data_ptr(), n, int(device), C.byref(nNode), C.byref(depth)), "trq_bvh_build_tree_device")
        self.maxDepth = depth.value
        return d


class Scene:
    """`Scene { primitives }` (Render.hh:132-134), resident on one GPU."""

    def __init__(self, primitives, device=0):
        self.primitives = primitives
        self._h = C.c_void_p(None)
        d = primitives.desc()
        if isinstance(primitives, DevicePrimitive):
            check(lib.trq_scene_create_device(C.byref(d), int(device), C.byref(self._h)), "trq_scene_create_device")
        else:
            check(lib.trq_scene_create(C.byref(d), int(device), C.byref(self._h)), "trq_scene_create")
        self.device = int(device)
        info = SceneInfo()
        check(lib.trq_scene_info(self._h, C.byref(info)), "trq_scene_info")
        self.info = {k: getattr(info, k) for k, _ in SceneInfo._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.trq_scene_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    # ---- Scene::hit over a batch ------------------------------------------------------------
    def hit(self, rays, any=False, out=None, reflayout=False, stream=None, sort=None, hit16=False):
        """rays: torch CUDA float32 tensor (n, 8) -> returns torch CUDA float32 tensor (n, 8) holding
        trq_hit rows (view as int32 for the id fields); or numpy `ray_dtype` array -> numpy `hit_dtype`.
        hit16=True: the 16-byte trq_hit16 records instead ((n, 4) float32 tensor / `hit16_dtype` array).
        sort: True = TRQ_SORT_RAYS (the batch is incoherent), False = TRQ_NO_SORT, None = the library decides (large trees only)."""
        flags = (L.TRACE_ANY if any else 0) | (L.KERNEL_REFLAYOUT if reflayout else 0) | _sort_flag(sort) | (L.HIT16 if hit16 else 0)
        width = 4 if hit16 else 8
        if isinstance(rays, np.ndarray):
            if rays.dtype != L.ray_dtype:
                raise TypeError("host rays must have ray_dtype")
            rays = np.ascontiguousarray(rays)
            hits = out if out is not None else np.empty(rays.size, dtype=L.hit16_dtype if hit16 else L.hit_dtype)
            check(lib.trq_trace(self._h, rays.ctypes.data, rays.size, flags | L.HOST_PTRS, hits.ctypes.data, None), "trq_trace")
            return hits
        import torch
        if not (rays.is_cuda and rays.dtype == torch.float32 and rays.dim() == 2 and rays.shape[1] == 8 and rays.is_contiguous()):
            raise TypeError("device rays must be a contiguous CUDA float32 tensor of shape (n, 8)")
        if rays.device.index != self.device:
            raise ValueError("rays are on a different device than the scene")
        n = rays.shape[0]
        hits = out if out is not None else torch.empty((n, width), dtype=torch.float32, device=rays.device)
        st = stream if stream is not None else torch.cuda.current_stream(rays.device).cuda_stream
        check(lib.trq_trace(self._h, rays.data_ptr(), n, flags, hits.data_ptr(), C.c_void_p(st)), "trq_trace")
        return hits

    def update_vertices(self, triList):
        """Refit after the vertices moved (same topology): numpy `vertex_dtype` array or a torch uint8 CUDA tensor."""
        if isinstance(triList, np.ndarray):
            a = np.ascontiguousarray(triList)
            check(lib.trq_scene_update_vertices(self._h, a.ctypes.data, a.size, L.HOST_PTRS), "trq_scene_update_vertices")
        else:
            n = triList.numel() * triList.element_size() // L.vertex_dtype.itemsize
            check(lib.trq_scene_update_vertices(self._h, triList.data_ptr(), n, 0), "trq_scene_update_vertices")

    @staticmethod
    def kernel_configs():
        """Names of the traversal kernel's launch configurations ("<CTA size>x<CTAs per SM><top of tree staged>")."""
        return [lib.trq_kernel_config_name(i).decode() for i in range(lib.trq_kernel_config_count())]

    def set_kernel_config(self, cfg=-1):
        """Selects a launch configuration (index; -1 = the library's choice). Returns the number of top-of-tree nodes that
        configuration stages in shared memory for this scene. Results are identical in every configuration."""
        staged = C.c_uint32(0)
        check(lib.trq_scene_set_kernel_config(self._h, int(cfg), C.byref(staged)), "trq_scene_set_kernel_config")
        return staged.value

    def kernel_config(self):
        """Name of the launch configuration this scene's traces use now."""
        return self.kernel_configs()[lib.trq_scene_kernel_config(self._h)]

    def host_sync(self):
        check(lib.trq_host_sync(self._h), "trq_host_sync")

    def hit_host(self, rays_ptr, n, hits_ptr, any=False, sort=False, asynchronous=False, hit16=False):
        """Raw host-pointer call (pinned buffers owned by the caller); used by the e2e bench."""
        flags = (L.TRACE_ANY if any else 0) | L.HOST_PTRS | (L.SORT_RAYS if sort else 0) | (L.HOST_ASYNC if asynchronous else 0) | (L.HIT16 if hit16 else 0)
        check(lib.trq_trace(self._h, rays_ptr, n, flags, hits_ptr, None), "trq_trace")

    # ---- wavefront callers (device tensors only) ----------------------------------------------
    def _stream(self, stream=None):
        import torch
        return C.c_void_p(stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream)

    def cast_rays(self, look_from, look_at, view_up, vfov, W, H, aperture=0.0, focus_dist=10.0, out=None, stream=None):
        """castRay for every pixel (Camera.hh:59-69, Render.metal:523-527) -> (W*H, 8) CUDA tensor of trq_ray."""
        import torch
        cam = Camera()
        cam.lookFrom[:], cam.lookAt[:], cam.viewUp[:] = look_from, look_at, view_up
        cam.vfov, cam.aspect, cam.aperture, cam.focus_dist = float(vfov), float(np.float32(W) / np.float32(H)), aperture, focus_dist
        rays = out if out is not None else torch.empty((W * H, 8), dtype=torch.float32, device=f"cuda:{self.device}")
        check(lib.trq_cast_rays(self._h, C.byref(cam), W, H, rays.data_ptr(), self._stream(stream)), "trq_cast_rays")
        return rays

    def hit_indirect(self, rays, count, any=False, out=None, sort=False, stream=None, hit16=False):
        """Scene::hit over min(*count, len(rays)) rays; `count` is a 1-element int64 CUDA tensor written by spawn_*."""
        import torch
        flags = (L.TRACE_ANY if any else 0) | (L.SORT_RAYS if sort else 0) | (L.HIT16 if hit16 else 0)
        hits = out if out is not None else torch.empty((rays.shape[0], 4 if hit16 else 8), dtype=torch.float32, device=rays.device)
        check(lib.trq_trace_indirect(self._h, rays.data_ptr(), count.data_ptr(), rays.shape[0], flags, hits.data_ptr(),
                                     self._stream(stream)), "trq_trace_indirect")
        return hits

    def _spawn(self, fn, rays, hits, count_in, seed_base, pixel_of, rng_state, extra, out, src, count, stream):
        import torch
        n = rays.shape[0]
        dev = rays.device
        out = out if out is not None else torch.empty((n, 8), dtype=torch.float32, device=dev)
        src = src if src is not None else torch.empty(n, dtype=torch.int32, device=dev)
        count = count if count is not None else torch.zeros(1, dtype=torch.int64, device=dev)
        args = [self._h, rays.data_ptr(), hits.data_ptr(), n, None if count_in is None else count_in.data_ptr(), int(seed_base),
                None if pixel_of is None else pixel_of.data_ptr(), None if rng_state is None else rng_state.data_ptr()]
        args += list(extra) + [out.data_ptr(), src.data_ptr(), count.data_ptr(), self._stream(stream)]
        check(fn(*args), fn.__name__)
        return out, src, count

    def rng_frame_begin(self, rng_state, stream=None):
        """Once per frame, before the first wave: the reference's toRNG entry conversion (halves of every texel swap)."""
        check(lib.trq_rng_frame_begin(self._h, rng_state.data_ptr(), rng_state.shape[0], self._stream(stream)), "trq_rng_frame_begin")

    def spawn_bounce(self, rays, hits, seed_base=0, count_in=None, out=None, src=None, count=None, stream=None,
                     pixel_of=None, rng_state=None):
        """Diffuse bounce rays of the hits (Render.metal:447-475), compacted. -> (rays_out, srcIndex, count tensor).
        rng_state: (pixels, 4) int32 CUDA tensor in the reference's RNG texture format (Render.hh:96-120), advanced in
        place; pixel_of: int32 tensor mapping input ray -> pixel (then srcIndex holds pixels too)."""
        return self._spawn(lib.trq_spawn_bounce_rng, rays, hits, count_in, seed_base, pixel_of, rng_state, (), out, src, count, stream)

    def spawn_shadow(self, rays, hits, light_a, light_b, seed_base=0, count_in=None, out=None, src=None, count=None, stream=None,
                     pixel_of=None, rng_state=None):
        """NEE shadow rays toward squareList[light_a|light_b] (Render.metal:313-337), compacted."""
        return self._spawn(lib.trq_spawn_shadow_rng, rays, hits, count_in, seed_base, pixel_of, rng_state,
                           (int(light_a), int(light_b)), out, src, count, stream)

    def profile(self, on=True):
        check(lib.trq_profile_enable(self._h, 1 if on else 0), "trq_profile_enable")

    def profile_read(self):
        """-> (launches, trace_kernel_ms_sum, resolve_kernel_ms_sum) since the last read."""
        n, a, b = C.c_uint32(0), C.c_float(0), C.c_float(0)
        check(lib.trq_profile_read(self._h, C.byref(n), C.byref(a), C.byref(b)), "trq_profile_read")
        return n.value, a.value, b.value

    def expand(self, rays, hits, out=None, stream=None):
        """trq_expand_hits: HitRecord fields (p, gn, sn, uv, f, material) for the hits of `rays`."""
        if isinstance(rays, np.ndarray):
            recs = out if out is not None else np.empty(rays.size, dtype=L.record_dtype)
            check(lib.trq_expand_hits(self._h, rays.ctypes.data, hits.ctypes.data, rays.size, L.HOST_PTRS,
                                      recs.ctypes.data, None), "trq_expand_hits")
            return recs
        import torch
        n = rays.shape[0]
        recs = out if out is not None else torch.empty((n, 16), dtype=torch.float32, device=rays.device)
        st = stream if stream is not None else torch.cuda.current_stream(rays.device).cuda_stream
        check(lib.trq_expand_hits(self._h, rays.data_ptr(), hits.data_ptr(), n, 0, recs.data_ptr(), C.c_void_p(st)),
              "trq_expand_hits")
        return recs


def _sort_flag(sort):
    return 0 if sort is None else (L.SORT_RAYS if sort else L.NO_SORT)


class MultiGpuScene:
    """trq_mgpu_*: one process, one scene per device, host rays sharded contiguously across the devices (no collective)."""

    def __init__(self, primitives, devices=None):
        self.primitives = primitives
        self._h = C.c_void_p(None)
        d = primitives.desc()
        if devices is None:
            check(lib.trq_mgpu_create(C.byref(d), None, 0, C.byref(self._h)), "trq_mgpu_create")
        else:
            arr = (C.c_int * len(devices))(*devices)
            check(lib.trq_mgpu_create(C.byref(d), arr, len(devices), C.byref(self._h)), "trq_mgpu_create")
        self.n_devices = lib.trq_mgpu_device_count(self._h)

    def shard(self, n, k):
        lo, hi = C.c_uint64(0), C.c_uint64(0)
        check(lib.trq_mgpu_shard(self._h, n, k, C.byref(lo), C.byref(hi)), "trq_mgpu_shard")
        return lo.value, hi.value

    def hit(self, rays, any=False, out=None, sort=False, hit16=False):
        """numpy `ray_dtype` array -> numpy `hit_dtype` (`hit16_dtype`) array (pinned buffers make the copies asynchronous)."""
        rays = np.ascontiguousarray(rays)
        hits = out if out is not None else np.empty(rays.size, dtype=L.hit16_dtype if hit16 else L.hit_dtype)
        flags = (L.TRACE_ANY if any else 0) | (L.SORT_RAYS if sort else 0) | (L.HIT16 if hit16 else 0)
        check(lib.trq_mgpu_trace(self._h, rays.ctypes.data, rays.size, flags, hits.ctypes.data), "trq_mgpu_trace")
        return hits

    def hit_ptr(self, rays_ptr, n, hits_ptr, any=False, sort=False):
        flags = (L.TRACE_ANY if any else 0) | (L.SORT_RAYS if sort else 0)
        check(lib.trq_mgpu_trace(self._h, rays_ptr, n, flags, hits_ptr), "trq_mgpu_trace")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.trq_mgpu_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close


def hits_to_numpy(hits):
    """torch (n, 8) float32 hit tensor -> numpy structured `hit_dtype` array (host copy)."""
    return hits.detach().cpu().numpy().view(L.hit_dtype).reshape(-1)


def rays_to_torch(rays, device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1, 8)).to(device)


def probe_bandwidth(device=0, which="l2"):
    """Measured read bandwidth in GB/s: "l2" (32 MB working set) or "hbm" (2 GB)."""
    g = C.c_double(0)
    check(lib.trq_probe_bandwidth(int(device), 0 if which == "l2" else 1, C.byref(g)), "trq_probe_bandwidth")
    return g.value


def launch_count():
    return int(lib.trq_launch_count())
