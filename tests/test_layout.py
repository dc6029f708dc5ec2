"""Byte layouts: numpy dtypes (tracer_b200/layout.py) and the C structs (csrc/host/layout.h, static_asserts)
against sizeof/offsetof of the reference's own headers compiled verbatim."""
from tracer_b200 import layout as L


def test_dtype_sizes():
    assert L.bvh_dtype.itemsize == 64
    assert L.vertex_dtype.itemsize == 32
    assert L.sphere_dtype.itemsize == 272
    assert L.square_dtype.itemsize == 272
    assert L.cube_dtype.itemsize == 240
    assert L.ray_dtype.itemsize == 32 and L.hit_dtype.itemsize == 32 and L.record_dtype.itemsize == 64


def test_against_reference_headers(reference):
    s = reference.sizes()
    assert s[:8] == [64, 32, 272, 272, 240, 32, 48, 224]        # BVH AABB Sphere Square Cube TriangleVertex Ray HitRecord
    off = lambda dt, f: dt.fields[f][1]
    assert (s[8], s[9], s[10]) == (off(L.bvh_dtype, "pType"), off(L.bvh_dtype, "pIndex"), off(L.bvh_dtype, "mini"))
    assert (s[11], s[12], s[22]) == (off(L.sphere_dtype, "center"), off(L.sphere_dtype, "material"), off(L.sphere_dtype, "box_mini"))
    assert (s[19], s[20], s[21], s[13]) == (off(L.square_dtype, "range_i"), off(L.square_dtype, "range_j"),
                                            off(L.square_dtype, "axis_k"), off(L.square_dtype, "value_k"))
    assert (s[14], s[15], s[23]) == (off(L.square_dtype, "model"), off(L.square_dtype, "material"), off(L.square_dtype, "box_mini"))
    assert (s[16], s[17], s[18]) == (off(L.cube_dtype, "inverse"), off(L.cube_dtype, "box_mini"), off(L.cube_dtype, "material"))
