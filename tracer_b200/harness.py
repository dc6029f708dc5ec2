"""Workload builders for the BASELINE configs: scenes in the reference's byte
layouts + the ray batches the reference's integrators would issue.

Scene constants restate RT_Metal/Tracer/Tracer.mm (MakeSquare :127-153, MakeCube :155-163,
MakeSphere :165-172, prepareCubeList :174-243, prepareCornellBox :245-304, prepareSphereList
:306-369, prepareCamera :371-411) and the mesh placement of AAPLRenderer.mm:513-573. Ray producers
are the C++ restatements in csrc/host/harness.cpp. Nothing here is on the timed path.
"""
import ctypes as C
import math
import os

import numpy as np

from . import layout as L
from ._lib import lib
from .scene import BVHBuilder, Primitive

_MESH_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "meshes")

f32 = np.float32
SquarePadding = f32(1.0 / 512.0)                                   # Square.hh:7-9


# ------------------------------------------------------------------ matrices (Tracer.mm:3-35), math convention M[r][c]
def scale4x4(sx, sy, sz):
    return np.diag([sx, sy, sz, 1.0]).astype(np.float32)


def translation4x4(tx, ty, tz):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (tx, ty, tz)
    return m


def rotation4x4(radians, axis):
    axis = np.asarray(axis, dtype=np.float32)
    axis = axis / np.linalg.norm(axis)
    ct, st = f32(math.cos(radians)), f32(math.sin(radians))
    ci = f32(1) - ct
    x, y, z = axis
    cols = np.array([
        [ct + x * x * ci, y * x * ci + z * st, z * x * ci - y * st, 0],
        [x * y * ci - z * st, ct + y * y * ci, z * y * ci + x * st, 0],
        [x * z * ci + y * st, y * z * ci - x * st, ct + z * z * ci, 0],
        [0, 0, 0, 1]], dtype=np.float32)
    return cols.T.copy()


def colmajor(m):
    return np.ascontiguousarray(m.T, dtype=np.float32).reshape(16)


# ------------------------------------------------------------------ primitives
def make_square(axis_i, range_i, axis_j, range_j, axis_k, k, material=0):
    s = np.zeros(1, dtype=L.square_dtype)[0]
    s["axis_i"], s["axis_j"], s["axis_k"] = axis_i, axis_j, axis_k
    s["range_i"], s["range_j"], s["value_k"] = range_i, range_j, k
    a = np.zeros(3, dtype=np.float32); b = np.zeros(3, dtype=np.float32)
    a[axis_i], a[axis_j], a[axis_k] = range_i[0], range_j[0], f32(k) - SquarePadding
    b[axis_i], b[axis_j], b[axis_k] = range_i[1], range_j[1], f32(k) + SquarePadding
    s["box_mini"], s["box_maxi"] = np.minimum(a, b), np.maximum(a, b)
    s["model"] = s["normal"] = s["inverse"] = L.IDENTITY4
    s["material"] = material
    return s


def cornell_squares():
    """prepareCornellBox (Tracer.mm:245-304): order left, right, top, back, bottom, light, little light.
    Material indices as the application numbers them: one shared list filled by prepareCubeList (3 materials), then
    prepareCornellBox (light, red, green, white), then prepareSphereList (AAPLRenderer.mm:216,222,228)."""
    light, red, green, white = 3, 4, 5, 6
    return np.array([
        make_square(1, (0, 555), 2, (0, 555), 0, -245, green),
        make_square(1, (0, 555), 2, (0, 555), 0, 800, red),
        make_square(0, (-245, 800), 2, (0, 555), 1, 555, white),
        make_square(0, (-245, 800), 1, (0, 555), 2, 555, white),
        make_square(0, (-245, 800), 2, (0, 555), 1, 0, white),
        make_square(0, (400, 555), 2, (200, 355), 1, 555 - 0.1, light),
        make_square(1, (200, 300), 2, (200, 300), 0, -300, light),
    ], dtype=L.square_dtype)


def make_cube(model, material):
    c = np.zeros(1, dtype=L.cube_dtype)[0]
    inv = np.linalg.inv(model.astype(np.float64)).astype(np.float32)
    c["model"], c["inverse"], c["normal"] = colmajor(model), colmajor(inv), colmajor(inv.T)
    c["box_mini"], c["box_maxi"] = (0, 0, 0), (1, 1, 1)
    c["material"] = material
    return c


def cornell_cubes():
    """prepareCubeList (Tracer.mm:174-243): bigger (material 0: the first material the application creates), smaller
    (the literal 19); the density-volume third cube is not in the BVH."""
    bigger = translation4x4(265, 1, 295) @ rotation4x4(math.pi * 15 / 180, (0, 1, 0)) @ scale4x4(165, 330, 165)
    smaller = translation4x4(130, 1, 65) @ rotation4x4(-0.1 * math.pi, (0, 1, 0)) @ scale4x4(165, 165, 165)
    return np.array([make_cube(bigger, 0), make_cube(smaller, 19)], dtype=L.cube_dtype)


def make_sphere(r, c, material=0):
    """MakeSphere (Tracer.mm:165-172): stored radius r + 0.0001, bounding box from r."""
    s = np.zeros(1, dtype=L.sphere_dtype)[0]
    s["radius"] = f32(r) + f32(0.0001)
    s["center"] = c
    c = np.asarray(c, dtype=np.float32)
    s["box_mini"], s["box_maxi"] = c - f32(r), c + f32(r)
    s["model"] = s["normal"] = s["inverse"] = L.IDENTITY4
    s["material"] = material
    return s


def cornell_spheres():
    """prepareSphereList (Tracer.mm:306-369)."""
    out = [make_sphere(64, (200, 250, 200), 7)]                  # materials 7..18 (after 3 cube + 4 Cornell materials)
    out += [make_sphere(40, (100 * (5 - i), 50, 50), 8 + i) for i in range(6)]
    out += [make_sphere(40, (-10 + 150 * i, 500, 400), 14 + i) for i in range(5)]
    return np.array(out, dtype=L.sphere_dtype)


# ------------------------------------------------------------------ triangle meshes
def make_vertices(pos, tris, normals=None, uv=None):
    """TriangleVertex array (32 B: v, n, uv) with area-weighted vertex normals unless given."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    tris = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    if normals is None:
        fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
        normals = np.zeros_like(pos)
        for k in range(3):
            np.add.at(normals, tris[:, k], fn)
        ln = np.linalg.norm(normals, axis=1, keepdims=True)
        normals = np.where(ln > 0, normals / np.maximum(ln, 1e-30), np.array([0, 0, 1], dtype=np.float32))
    v = np.zeros(len(pos), dtype=L.vertex_dtype)
    v["v"], v["n"] = pos, normals.astype(np.float32)
    if uv is not None:
        v["uv"] = uv
    return v


def square_triangles(sq):
    ai, aj, ak = int(sq["axis_i"]), int(sq["axis_j"]), int(sq["axis_k"])
    p = np.zeros((4, 3), dtype=np.float32)
    corners = [(0, 0), (1, 0), (1, 1), (0, 1)]
    for n, (a, b) in enumerate(corners):
        p[n, ai], p[n, aj], p[n, ak] = sq["range_i"][a], sq["range_j"][b], sq["value_k"]
    return p, np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int64), np.array(corners, dtype=np.float32)


def cube_triangles(cube):
    m = np.asarray(cube["model"], dtype=np.float32).reshape(4, 4).T          # back to M[r][c]
    corners = np.array([[x, y, z, 1] for x in (0, 1) for y in (0, 1) for z in (0, 1)], dtype=np.float32)
    p = (corners @ m.T)[:, :3].astype(np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = [(q[0], q[1], q[2]) for q in quads] + [(q[0], q[2], q[3]) for q in quads]
    return p, np.array(tris, dtype=np.int64)


def load_mesh(name):
    z = np.load(os.path.join(_MESH_DIR, name + ".npz"))
    return z["v"].astype(np.float32), z["tris"].astype(np.int64)


def place_mesh(pos, x_shift=-200.0):
    """Mesh normalisation of AAPLRenderer.mm:513-573 (scale to 300, stand on y=180 ... , flip z)."""
    pos = pos.astype(np.float32)
    lo, hi = pos.min(0), pos.max(0)
    centroid = lo + (hi - lo) / f32(2)
    d = hi - lo
    max_axis = 0 if (d[0] > d[1] and d[0] > d[2]) else (1 if d[1] > d[2] else 2)
    scale = f32(300.0) / d[max_axis]
    off = np.full(3, 278, dtype=np.float32) - centroid
    off[1] = f32(180) - lo[1] * scale
    out = np.empty_like(pos)
    out[:, 0] = pos[:, 0] * scale + off[0] + f32(x_shift)
    out[:, 1] = pos[:, 1] * scale + off[1]
    out[:, 2] = pos[:, 2] * (-scale) + off[2]
    return out


def subdivide(pos, tris, levels=1):
    """1 -> 4 midpoint subdivision with shared edge midpoints."""
    pos = np.asarray(pos, dtype=np.float32)
    tris = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    for _ in range(levels):
        e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
        e.sort(axis=1)
        key = e[:, 0] * np.int64(len(pos)) + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // len(pos), uniq % len(pos)
        mid = ((pos[a] + pos[b]) * f32(0.5)).astype(np.float32)
        m = inv.reshape(3, -1) + len(pos)                    # midpoint ids of edges 01, 12, 20
        v0, v1, v2 = tris[:, 0], tris[:, 1], tris[:, 2]
        tris = np.concatenate([
            np.stack([v0, m[0], m[2]], 1), np.stack([v1, m[1], m[0]], 1),
            np.stack([v2, m[2], m[1]], 1), np.stack([m[0], m[1], m[2]], 1)])
        pos = np.concatenate([pos, mid])
    return pos, tris


class MeshSoup:
    """Accumulates indexed triangle meshes into one triList / idxList pair."""

    def __init__(self):
        self.verts, self.idx, self.nv = [], [], 0

    def add(self, pos, tris, uv=None):
        v = make_vertices(pos, tris, uv=uv)
        self.verts.append(v)
        self.idx.append((np.asarray(tris, dtype=np.int64) + self.nv).astype(np.uint32))
        self.nv += len(v)

    def arrays(self):
        if not self.verts:
            return np.zeros(0, dtype=L.vertex_dtype), np.zeros(0, dtype=np.uint32)
        return np.concatenate(self.verts), np.concatenate(self.idx).reshape(-1)


def build_primitive(triList=None, idxList=None, spheres=None, squares=None, cubes=None,
                    sphere_leaves=None, square_leaves=None, cube_leaves=None):
    """Leaf creation order of AAPLRenderer.mm:454-468,546-591: (spheres,) cubes, squares, triangles; then buildTree."""
    b = BVHBuilder()
    if spheres is not None:
        for i in (range(len(spheres)) if sphere_leaves is None else sphere_leaves):
            b.buildNode(spheres[i]["box_mini"], spheres[i]["box_maxi"], spheres[i]["model"], L.SPHERE, i)
    if cubes is not None:
        for i in (range(len(cubes)) if cube_leaves is None else cube_leaves):
            b.buildNode(cubes[i]["box_mini"], cubes[i]["box_maxi"], cubes[i]["model"], L.CUBE, i)
    if squares is not None:
        for i in (range(len(squares)) if square_leaves is None else square_leaves):
            b.buildNode(squares[i]["box_mini"], squares[i]["box_maxi"], squares[i]["model"], L.SQUARE, i)
    if triList is not None and len(triList):
        b.buildNodesTriangles(triList, idxList, 0)
    bvh = b.buildTree()
    return Primitive(sphereList=spheres, squareList=squares, cubeList=cubes, triList=triList, idxList=idxList, bvhList=bvh)


# ------------------------------------------------------------------ BASELINE scenes
def cornell_triangle_soup(with_teapot=True, mesh_levels=0, with_coatball=False):
    """C2 / C3 geometry: Cornell squares and cubes as triangles (+ teapot, + coatball), optionally subdivided."""
    ms = MeshSoup()
    for sq in cornell_squares():
        p, t, uv = square_triangles(sq)
        ms.add(p, t, uv)
    for cb in cornell_cubes():
        p, t = cube_triangles(cb)
        ms.add(p, t)
    if with_teapot:
        p, t = load_mesh("teapot")
        p = place_mesh(p, x_shift=(250.0 if with_coatball else -200.0))
        p, t = subdivide(p, t, mesh_levels)
        ms.add(p, t)
    if with_coatball:
        p, t = load_mesh("coatball")
        p = place_mesh(p, x_shift=-200.0)
        p, t = subdivide(p, t, mesh_levels)
        ms.add(p, t)
    return ms.arrays()


def scene_c2():
    """C2: Cornell box triangles + teapot (38 + 15,704 triangles)."""
    tri, idx = cornell_triangle_soup(with_teapot=True)
    return build_primitive(tri, idx)


def scene_c3(levels=2):
    """C3: 'meshes' scene, coatball + teapot each subdivided `levels` times inside the Cornell box (~1.0 M tris at 2)."""
    tri, idx = cornell_triangle_soup(with_teapot=True, mesh_levels=levels, with_coatball=True)
    return build_primitive(tri, idx)


def scene_c4(levels=2, n_random=10000, seed=7):
    """C4: C3 geometry + the 12 spheres of prepareSphereList + n_random spheres r in [2,10] in the box."""
    tri, idx = cornell_triangle_soup(with_teapot=True, mesh_levels=levels, with_coatball=True)
    sph = list(cornell_spheres())
    xi = pcg32_floats(seed, 0, 4 * n_random).reshape(-1, 4)
    for k in range(n_random):
        c = (f32(-245) + f32(1045) * xi[k, 0], f32(555) * xi[k, 1], f32(555) * xi[k, 2])
        sph.append(make_sphere(f32(2) + f32(8) * xi[k, 3], c, 30))
    spheres = np.array(sph, dtype=L.sphere_dtype)
    # squareList is uploaded for the light samples (lights 5 / 6) but gets no leaves: the walls are triangles here
    return build_primitive(tri, idx, spheres=spheres, squares=cornell_squares(), square_leaves=[])


def scene_c4_lights(prim):
    return prim.squareList[5:6], prim.squareList[6:7]


def scene_soup(n_tri, seed=1, extent=0.004):
    """C5-style random triangle soup in the unit cube."""
    tri = np.zeros(3 * n_tri, dtype=L.vertex_dtype)
    idx = np.zeros(3 * n_tri, dtype=np.uint32)
    lib.trqh_make_soup(n_tri, seed, extent, tri.ctypes.data, idx.ctypes.data)
    return build_primitive(tri, idx)


def scene_reference_cornell(with_spheres=True):
    """The reference's actual leaf mix: Cube + Square leaves (AAPLRenderer.mm:459-468) and, re-enabled as in
    the commented code (:454-457), Sphere leaves; plus the C2 triangles of the teapot."""
    ms = MeshSoup()
    p, t = load_mesh("teapot")
    ms.add(place_mesh(p), t)
    tri, idx = ms.arrays()
    return build_primitive(tri, idx, spheres=cornell_spheres() if with_spheres else None,
                           squares=cornell_squares(), cubes=cornell_cubes())


def scene_c1(seed=42, seq=54):
    """C1: RT_Nextweek randomScene (Render.swift:83-130) as Sphere leaves in the RT_Metal SAH BVH."""
    u32 = np.empty(12 * 21 * 21, dtype=np.uint32)
    lib.trqh_pcg32_fill_u32(seed, seq, u32.size, u32.ctypes.data)
    xi = u32.astype(np.float32) / f32(0xFFFFFFFF)               # randomFloat()  Random.swift:3-6 (arc4random -> PCG32)
    k = 0

    def draw():
        nonlocal k
        k += 1
        return xi[k - 1]

    sph = [make_sphere(1000, (0, -1000, 0), 0), make_sphere(1.0, (0, 1, 0), 2),
           make_sphere(1.0, (-4, 1, 0), 3), make_sphere(1.0, (4, 1, 0), 4)]
    for a in range(-10, 11):
        for b in range(-10, 11):
            mat = draw()
            c = np.array([f32(a) + f32(0.9) * draw(), 0.2, f32(b) + f32(0.9) * draw()], dtype=np.float32)
            if np.linalg.norm(c - np.array([4, 0.2, 0], dtype=np.float32)) > 0.9:
                # material draws consumed by the reference (MovingSphere frozen at centerS, t = 0)
                k += 7 if mat < 0.8 else (4 if mat < 0.95 else 0)
                sph.append(make_sphere(0.2, c, 1))
    spheres = np.array(sph, dtype=L.sphere_dtype)
    # RT_Nextweek spheres have no +0.0001 radius pad (Sphere.swift): undo MakeSphere's
    spheres["radius"] = np.where(np.arange(len(sph)) == 0, f32(1000), np.where(np.arange(len(sph)) < 4, f32(1.0), f32(0.2)))
    return build_primitive(spheres=spheres)


# ------------------------------------------------------------------ rays
def pcg32_floats(seed, seq, n):
    out = np.empty(n, dtype=np.float32)
    lib.trqh_pcg32_fill_f32(seed, seq, n, out.ctypes.data)
    return out


def cornell_camera_rays(W, H):
    """prepareCamera + castRay (Tracer.mm:371-411, Camera.hh:59-69): aperture 0, s = x/W, t = y/H."""
    rays = np.empty(W * H, dtype=L.ray_dtype)
    frm = np.array([278, 278, -800], dtype=np.float32)
    at = np.array([278, 278, 278], dtype=np.float32)
    up = np.array([0, 1, 0], dtype=np.float32)
    vfov = f32(45 * (math.pi / 180))
    lib.trqh_gen_camera_rays(frm.ctypes.data, at.ctypes.data, up.ctypes.data, vfov, f32(W) / f32(H), f32(10), W, H,
                             rays.ctypes.data)
    return rays


def camera_rays(look_from, look_at, vfov_rad, W, H, up=(0, 1, 0), focus=10.0):
    rays = np.empty(W * H, dtype=L.ray_dtype)
    frm = np.array(look_from, dtype=np.float32); at = np.array(look_at, dtype=np.float32); u = np.array(up, dtype=np.float32)
    lib.trqh_gen_camera_rays(frm.ctypes.data, at.ctypes.data, u.ctypes.data, f32(vfov_rad), f32(W) / f32(H), f32(focus), W, H,
                             rays.ctypes.data)
    return rays


def random_rays(n, seed=2, lo=(0, 0, 0), hi=(1, 1, 1), tmax=L.FLT_MAX, first=0):
    rays = np.empty(n, dtype=L.ray_dtype)
    lo = np.array(lo, dtype=np.float32); hi = np.array(hi, dtype=np.float32)
    lib.trqh_gen_random_rays(first, n, seed, lo.ctypes.data, hi.ctypes.data, f32(tmax), rays.ctypes.data)
    return rays


def _opt(a, dtype):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data


def rng_frame_begin(rng_state):
    """Host twin of trq_rng_frame_begin (in place on a contiguous (pixels, 4) uint32 array)."""
    assert rng_state.dtype == np.uint32 and rng_state.flags["C_CONTIGUOUS"]
    lib.trqh_rng_frame_begin(rng_state.ctypes.data, rng_state.shape[0])


def bounce_rays(records, seed_base=0, pixel_of=None, rng_state=None):
    """Diffuse bounce rays spawned from hit records (Render.metal:447-475); compacted. rng_state: (pixels, 4) uint32
    array in the reference's RNG texture format (Render.hh:96-120), advanced IN PLACE when it is a contiguous uint32 array."""
    records = np.ascontiguousarray(records)
    rays = np.empty(records.size, dtype=L.ray_dtype)
    src = np.empty(records.size, dtype=np.uint32)
    _, pp = _opt(pixel_of, np.uint32)
    _, rp = _opt(rng_state, np.uint32)
    k = lib.trqh_gen_bounce_rays_rng(records.ctypes.data, records.size, seed_base, pp, rp, rays.ctypes.data, src.ctypes.data)
    return rays[:k].copy(), src[:k].copy()


def shadow_rays(records, light_a, light_b, seed_base=0, pixel_of=None, rng_state=None):
    """NEE shadow rays toward two light squares (Render.metal:313-337); compacted."""
    records = np.ascontiguousarray(records)
    la = np.ascontiguousarray(light_a); lb = np.ascontiguousarray(light_b)
    rays = np.empty(records.size, dtype=L.ray_dtype)
    src = np.empty(records.size, dtype=np.uint32)
    _, pp = _opt(pixel_of, np.uint32)
    _, rp = _opt(rng_state, np.uint32)
    k = lib.trqh_gen_shadow_rays_rng(records.ctypes.data, records.size, seed_base, pp, rp, la.ctypes.data, lb.ctypes.data,
                                     rays.ctypes.data, src.ctypes.data)
    return rays[:k].copy(), src[:k].copy()
