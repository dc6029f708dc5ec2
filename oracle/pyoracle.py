"""ctypes front-end of the two oracle libraries -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  Port       oracle/liboracle_rq.so        plain-C restatement, instrumented (oracle_rq.c)
  Reference  oracle/_ref/libtracer_ref.so  the reference's Render.hh compiled verbatim (ref_scene_hit.cpp)

Both take the same reference-layout arrays as tracer_b200.Primitive.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from tracer_b200 import layout as L

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(_HERE, "liboracle_rq.so")
REF_PATH = os.path.join(_HERE, "_ref", "libtracer_ref.so")

counters_dtype = np.dtype([(k, "<u4") for k in
                           ("n_fp", "n_fc", "n_disp", "n_tri", "n_sph", "n_sq", "n_cube", "n_tie", "n_quirk", "max_level")])
TOTAL_KEYS = [k for k, _ in counters_dtype.descr]


def build(which=("port", "ref")):
    subprocess.run(["make", "-s", "-C", _HERE] + list(which), check=True)


class _Prims(C.Structure):
    _fields_ = [("sphereList", C.c_void_p), ("squareList", C.c_void_p), ("cubeList", C.c_void_p),
                ("triList", C.c_void_p), ("idxList", C.c_void_p), ("bvhList", C.c_void_p)]


def _prims(p):
    q = _Prims()
    for k in ("sphereList", "squareList", "cubeList", "triList", "idxList", "bvhList"):
        a = getattr(p, k)
        setattr(q, k, a.ctypes.data if a.size else None)
    return q


def _fp(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Port:
    """The C restatement (kind "port")."""

    def __init__(self):
        if not os.path.exists(PORT_PATH):
            build(("port",))
        self.lib = C.CDLL(PORT_PATH)
        self.lib.orq_algorithmic_bytes.restype = C.c_uint64

    def trace(self, prim, rays, any=False, nthreads=1, records=False, counters=False):
        rays = np.ascontiguousarray(rays)
        n = rays.size
        hits = np.zeros(n, dtype=L.hit_dtype)
        recs = np.zeros(n, dtype=L.record_dtype) if records else None
        cnts = np.zeros(n, dtype=counters_dtype) if counters else None
        totals = np.zeros(10, dtype=np.uint64)
        q = _prims(prim)
        self.lib.orq_trace(C.byref(q), C.c_void_p(rays.ctypes.data), C.c_uint64(n), C.c_int(1 if any else 0), C.c_int(nthreads),
                           C.c_void_p(hits.ctypes.data), C.c_void_p(recs.ctypes.data if records else None),
                           C.c_void_p(cnts.ctypes.data if counters else None), C.c_void_p(totals.ctypes.data))
        tot = dict(zip(TOTAL_KEYS, (int(x) for x in totals)))
        tot["bytes"] = int(self.lib.orq_algorithmic_bytes(C.c_void_p(totals.ctypes.data), C.c_uint64(n)))
        tot["n_rays"] = n
        return {"hits": hits, "records": recs, "counters": cnts, "totals": tot}

    def aabb_hit_t(self, box8, o, d, rng):
        box8, o, d, rng = _fp(box8), _fp(o), _fp(d), _fp(rng)
        t = C.c_float(0)
        h = self.lib.orq_aabb_hit_t(C.c_void_p(box8.ctypes.data), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                                    C.c_void_p(rng.ctypes.data), C.byref(t))
        return bool(h), np.float32(t.value)

    def aabb_hit(self, box8, o, d, rng):
        box8, o, d, rng = _fp(box8), _fp(o), _fp(d), _fp(rng)
        return bool(self.lib.orq_aabb_hit(C.c_void_p(box8.ctypes.data), C.c_void_p(o.ctypes.data),
                                          C.c_void_p(d.ctypes.data), C.c_void_p(rng.ctypes.data)))

    def _leaf(self, fn, obj, o, d, rng, extra=None):
        o, d = _fp(o), _fp(d)
        rng = _fp(rng).copy()
        rec = np.zeros(1, dtype=L.record_dtype)
        args = [C.c_void_p(obj.ctypes.data)] + ([] if extra is None else [C.c_void_p(extra.ctypes.data)]) + \
               [C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(rng.ctypes.data), C.c_void_p(rec.ctypes.data)]
        return fn, args, rng, rec

    def triangle_hit(self, triList, abc, o, d, rng):
        abc = np.ascontiguousarray(abc, dtype=np.uint32)
        fn, args, rng, rec = self._leaf(self.lib.orq_triangle_hit, triList, o, d, rng, extra=abc)
        bary = np.zeros(2, dtype=np.float32)
        h = fn(*args, C.c_void_p(bary.ctypes.data))
        return bool(h), rng, rec[0], bary

    def sphere_hit(self, sphere, o, d, rng):
        fn, args, rng, rec = self._leaf(self.lib.orq_sphere_hit, np.ascontiguousarray(sphere), o, d, rng)
        return bool(fn(*args)), rng, rec[0]

    def square_hit(self, square, o, d, rng):
        fn, args, rng, rec = self._leaf(self.lib.orq_square_hit, np.ascontiguousarray(square), o, d, rng)
        return bool(fn(*args)), rng, rec[0]

    def cube_hit(self, cube, o, d, rng):
        fn, args, rng, rec = self._leaf(self.lib.orq_cube_hit, np.ascontiguousarray(cube), o, d, rng)
        return bool(fn(*args)), rng, rec[0]

    def offset_ray(self, p, n):
        p, n = _fp(p), _fp(n)
        out = np.zeros(3, dtype=np.float32)
        self.lib.orq_offset_ray(C.c_void_p(p.ctypes.data), C.c_void_p(n.ctypes.data), C.c_void_p(out.ctypes.data))
        return out


class Reference:
    """The reference's own headers compiled verbatim (kind "reference"). Only where oracle/_ref was built."""

    @staticmethod
    def available():
        return os.path.exists(REF_PATH)

    def __init__(self):
        if not os.path.exists(REF_PATH):
            if os.path.isdir("/root/reference/RT_Metal/Metal"):
                build(("ref",))
            else:
                raise FileNotFoundError(REF_PATH)
        self.lib = C.CDLL(REF_PATH)
        self.lib.ref_next_float_up.restype = C.c_float
        self.lib.ref_next_float_up.argtypes = [C.c_float]
        self.lib.ref_next_float_down.restype = C.c_float
        self.lib.ref_next_float_down.argtypes = [C.c_float]

    def sizes(self):
        out = (C.c_uint32 * 32)()
        self.lib.ref_sizes(out)
        return list(out)

    def trace(self, prim, rays, any=False, nthreads=1):
        rays = np.ascontiguousarray(rays)
        recs = np.zeros(rays.size, dtype=L.record_dtype)
        q = _prims(prim)
        self.lib.ref_scene_hit(C.byref(q), C.c_void_p(rays.ctypes.data), C.c_uint64(rays.size), C.c_int(1 if any else 0),
                               C.c_int(nthreads), C.c_void_p(recs.ctypes.data))
        return recs

    def trace_lite(self, prim, rays, any=False, nthreads=1):
        rays = np.ascontiguousarray(rays)
        hit = np.zeros(rays.size, dtype=np.uint32)
        t = np.zeros(rays.size, dtype=np.float32)
        q = _prims(prim)
        self.lib.ref_scene_hit_lite(C.byref(q), C.c_void_p(rays.ctypes.data), C.c_uint64(rays.size), C.c_int(1 if any else 0),
                                    C.c_int(nthreads), C.c_void_p(hit.ctypes.data), C.c_void_p(t.ctypes.data))
        return hit, t

    def aabb_hit_t(self, box8, o, d, rng):
        box8, o, d, rng = _fp(box8), _fp(o), _fp(d), _fp(rng)
        t = C.c_float(0)
        h = self.lib.ref_aabb_hit_t(C.c_void_p(box8.ctypes.data), C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data),
                                    C.c_void_p(rng.ctypes.data), C.byref(t))
        return bool(h), np.float32(t.value)

    def aabb_hit(self, box8, o, d, rng):
        box8, o, d, rng = _fp(box8), _fp(o), _fp(d), _fp(rng)
        return bool(self.lib.ref_aabb_hit(C.c_void_p(box8.ctypes.data), C.c_void_p(o.ctypes.data),
                                          C.c_void_p(d.ctypes.data), C.c_void_p(rng.ctypes.data)))

    def _leaf(self, fn, obj, o, d, rng, extra=None):
        o, d = _fp(o), _fp(d)
        rng = _fp(rng).copy()
        rec = np.zeros(1, dtype=L.record_dtype)
        args = [C.c_void_p(obj.ctypes.data)] + ([] if extra is None else [C.c_void_p(extra.ctypes.data)]) + \
               [C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(rng.ctypes.data), C.c_void_p(rec.ctypes.data)]
        return bool(fn(*args)), rng, rec[0]

    def triangle_hit(self, triList, abc, o, d, rng):
        return self._leaf(self.lib.ref_triangle_hit, triList, o, d, rng, extra=np.ascontiguousarray(abc, dtype=np.uint32))

    def sphere_hit(self, sphere, o, d, rng):
        return self._leaf(self.lib.ref_sphere_hit, np.ascontiguousarray(sphere), o, d, rng)

    def square_hit(self, square, o, d, rng):
        return self._leaf(self.lib.ref_square_hit, np.ascontiguousarray(square), o, d, rng)

    def cube_hit(self, cube, o, d, rng):
        return self._leaf(self.lib.ref_cube_hit, np.ascontiguousarray(cube), o, d, rng)

    def offset_ray(self, p, n):
        p, n = _fp(p), _fp(n)
        out = np.zeros(3, dtype=np.float32)
        self.lib.ref_offset_ray(C.c_void_p(p.ctypes.data), C.c_void_p(n.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def ray_ctor(self, o, d):
        o, d = _fp(o), _fp(d)
        oo, dd = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
        self.lib.ref_ray_ctor(C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(oo.ctypes.data), C.c_void_p(dd.ctypes.data))
        return oo, dd

    def coordinate_system(self, a):
        a = _fp(a)
        b, c = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
        self.lib.ref_coordinate_system(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(c.ctypes.data))
        return b, c

    def cosine_sample_hemisphere(self, u):
        u = _fp(u)
        out = np.zeros(3, dtype=np.float32)
        self.lib.ref_cosine_sample_hemisphere(C.c_void_p(u.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def cast_ray(self, cam18, s, t, len_radius=0.0, seed=1, seq=1):
        """castRay (Camera.hh:59-69) -> (origin, direction)."""
        c = _fp(cam18)
        o, d = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
        self.lib.ref_cast_ray(C.c_void_p(c.ctypes.data), C.c_float(len_radius), C.c_float(s), C.c_float(t), C.c_uint64(seed), C.c_uint64(seq),
                              C.c_void_p(o.ctypes.data), C.c_void_p(d.ctypes.data))
        return o, d

    def square_sample(self, square, uu, pos):
        """Square::sample (Square.hh:40-58) -> (lsr.p, lsr.n)."""
        sq = np.ascontiguousarray(square).reshape(1)
        u, p = _fp(uu), _fp(pos)
        po, no = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
        self.lib.ref_square_sample(C.c_void_p(sq.ctypes.data), C.c_void_p(u.ctypes.data), C.c_void_p(p.ctypes.data),
                                   C.c_void_p(po.ctypes.data), C.c_void_p(no.ctypes.data))
        return po, no

    def pcg32(self, initstate, initseq, n):
        """pcg32_srandom_r + n x pcg32_random_r / randomF (Random.metal:3-26) -> (uint32 array, float32 array)."""
        u = np.zeros(n, dtype=np.uint32)
        f = np.zeros(n, dtype=np.float32)
        self.lib.ref_pcg32_fill(C.c_uint64(initstate), C.c_uint64(initseq), C.c_uint32(n), C.c_void_p(u.ctypes.data), C.c_void_p(f.ctypes.data))
        return u, f

    def sample2d(self, initstate, initseq):
        """RandomSampler::sample2D (RandomSampler.hh:16-21) of a freshly seeded stream."""
        out = np.zeros(2, dtype=np.float32)
        self.lib.ref_sample2d(C.c_uint64(initstate), C.c_uint64(initseq), C.c_void_p(out.ctypes.data))
        return out

    def to_rng(self, rgba):
        """toRNG (Render.hh:96-107): 4 x uint32 texel -> (inc, state)."""
        t = np.ascontiguousarray(rgba, dtype=np.uint32)
        out = np.zeros(2, dtype=np.uint64)
        self.lib.ref_to_rng(C.c_void_p(t.ctypes.data), C.c_void_p(out.ctypes.data))
        return int(out[0]), int(out[1])

    def ex_rng(self, inc, state):
        """exRNG (Render.hh:109-120): (inc, state) -> 4 x uint32 texel."""
        a = np.array([inc, state], dtype=np.uint64)
        out = np.zeros(4, dtype=np.uint32)
        self.lib.ref_ex_rng(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data))
        return out


REF_BUILDER_PATH = os.path.join(_HERE, "_ref", "libtracer_ref_builder.so")


class ReferenceBuilder:
    """The reference's own host BVH builder (BVH.hh:30-314) compiled from its source (oracle/ref_builder.cpp)."""

    @staticmethod
    def available():
        return os.path.exists(REF_BUILDER_PATH)

    def __init__(self):
        if not os.path.exists(REF_BUILDER_PATH):
            if os.path.isdir("/root/reference/RT_Metal/Metal"):
                build(("ref",))
            else:
                raise FileNotFoundError(REF_BUILDER_PATH)
        self.lib = C.CDLL(REF_BUILDER_PATH)
        self.lib.refb_build_tree.restype = C.c_uint32

    def build_tree(self, leaves):
        """BVH::buildTree over a copy of `leaves` (numpy bvh_dtype); returns the 2n-1 nodes. The two padding words of a node
        are uninitialised in the reference (struct BVH has none of its own): zeroed here."""
        leaves = np.ascontiguousarray(leaves)
        n = leaves.size
        nodes = np.zeros(max(2 * n - 1, 1), dtype=leaves.dtype)
        nodes[:n] = leaves
        k = self.lib.refb_build_tree(C.c_void_p(nodes.ctypes.data), C.c_uint32(n))
        assert k == nodes.size, (k, nodes.size)
        nodes["pad"] = 0
        return nodes

    def build_node(self, box_min, box_max, model, p_type, p_index, dtype):
        lo, hi = _fp(box_min), _fp(box_max)
        m = None if model is None else np.ascontiguousarray(model, dtype=np.float32).reshape(16)
        out = np.zeros(1, dtype=dtype)
        self.lib.refb_build_node(C.c_void_p(lo.ctypes.data), C.c_void_p(hi.ctypes.data),
                                 C.c_void_p(m.ctypes.data if m is not None else None), C.c_int32(p_type), C.c_uint32(p_index),
                                 C.c_void_p(out.ctypes.data))
        out["pad"] = 0
        return out[0]


REF_SETUP_PATH = os.path.join(_HERE, "_ref", "libtracer_ref_setup.so")


class ReferenceSetup:
    """The reference's scene set-up code (RT_Metal/Tracer/Tracer.mm) compiled from its source (oracle/ref_scene_setup.cpp)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SETUP_PATH)

    def __init__(self):
        if not os.path.exists(REF_SETUP_PATH):
            if os.path.isdir("/root/reference/RT_Metal/Tracer"):
                build(("ref",))
            else:
                raise FileNotFoundError(REF_SETUP_PATH)
        self.lib = C.CDLL(REF_SETUP_PATH)
        for fn in (self.lib.refs_cornell_squares, self.lib.refs_cubes, self.lib.refs_spheres, self.lib.refs_material_count):
            fn.restype = C.c_uint32

    def _fetch(self, fn, dtype):
        n = fn(None, 0)
        a = np.zeros(n, dtype=dtype)
        fn(C.c_void_p(a.ctypes.data), n)
        return a

    def squares(self):
        return self._fetch(self.lib.refs_cornell_squares, L.square_dtype)

    def cubes(self):
        return self._fetch(self.lib.refs_cubes, L.cube_dtype)

    def spheres(self):
        return self._fetch(self.lib.refs_spheres, L.sphere_dtype)

    def prepare_camera(self, view_w, view_h):
        out = np.zeros(18, dtype=np.float32)
        self.lib.refs_prepare_camera(C.c_float(view_w), C.c_float(view_h), C.c_void_p(out.ctypes.data))
        return out

    def make_camera(self, look_from, look_at, view_up, aperture, aspect, vfov, focus):
        a, b, c = _fp(look_from), _fp(look_at), _fp(view_up)
        out = np.zeros(18, dtype=np.float32)
        self.lib.refs_make_camera(C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(c.ctypes.data), C.c_float(aperture),
                                  C.c_float(aspect), C.c_float(vfov), C.c_float(focus), C.c_void_p(out.ctypes.data))
        return out


NEXTWEEK_PATH = os.path.join(_HERE, "libnextweek_bvh.so")


class Nextweek:
    """RT_Nextweek's CPU BVH (oracle/nextweek_bvh.c): config C1's named CPU baseline. Parity unpinned (no Swift)."""

    def __init__(self, spheres, seed=42, seq=55):
        if not os.path.exists(NEXTWEEK_PATH):
            build(("port",))
        self.lib = C.CDLL(NEXTWEEK_PATH)
        self.lib.nw_build.restype = C.c_void_p
        self.lib.nw_node_count.restype = C.c_uint32
        cr = np.zeros((len(spheres), 4), dtype=np.float32)
        cr[:, :3], cr[:, 3] = spheres["center"], spheres["radius"]
        self.n = len(spheres)
        self.h = C.c_void_p(self.lib.nw_build(C.c_void_p(cr.ctypes.data), C.c_uint32(self.n), C.c_uint64(seed), C.c_uint64(seq)))

    def trace(self, rays, t_min=0.001, nthreads=1):
        rays = np.ascontiguousarray(rays)
        ids = np.zeros(rays.size, dtype=np.uint32)
        t = np.zeros(rays.size, dtype=np.float32)
        self.lib.nw_trace(self.h, C.c_void_p(rays.ctypes.data), C.c_uint64(rays.size), C.c_float(t_min), C.c_int(nthreads),
                          C.c_void_p(ids.ctypes.data), C.c_void_p(t.ctypes.data))
        return ids, t

    def node_count(self):
        return int(self.lib.nw_node_count(self.h))

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.nw_free.argtypes = [C.c_void_p]
            self.lib.nw_free(self.h)
            self.h = C.c_void_p(None)
