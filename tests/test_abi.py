"""CPU: the C-ABI shared library loads, exports every symbol include/*.h declares, validates scene layouts on the
host, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tracer_b200 import harness as H, layout as L
from tracer_b200 import _lib
from tracer_b200._lib import ERR_INVALID, ERR_LAYOUT, ERR_NO_DEVICE, ERR_DEPTH, OK, SceneDesc, lib
from tracer_b200.scene import Primitive

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trqh?_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    names = declared("tracer_rq.h") + declared("tracer_rq_harness.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported by libtracer_rq.so"
    assert set(_lib.ABI_SYMBOLS) <= set(declared("tracer_rq.h"))
    assert set(_lib.HARNESS_SYMBOLS) <= set(declared("tracer_rq_harness.h"))
    assert lib.trq_version() == 200


def test_struct_sizes_match_header(built):
    assert C.sizeof(SceneDesc) == 6 * 16
    assert C.sizeof(_lib.SceneInfo) == 56


def _create(prim):
    h = C.c_void_p(None)
    d = prim.desc()
    return lib.trq_scene_create(C.byref(d), 0, C.byref(h)), h


def test_layout_validation_runs_on_the_host(built):
    prim = H.scene_soup(50, seed=1, extent=0.2)
    bad = Primitive(triList=prim.triList, idxList=prim.idxList, bvhList=prim.bvhList.copy())
    bad.bvhList["left"][0] = bad.bvhList.size + 5                        # child index out of range
    rc, _ = _create(bad)
    assert rc == ERR_LAYOUT and b"out of range" in lib.trq_last_error_string()
    bad = Primitive(triList=prim.triList, idxList=prim.idxList, bvhList=prim.bvhList.copy())
    bad.bvhList["right"][0] = bad.bvhList["left"][0]                     # not a tree
    assert _create(bad)[0] == ERR_LAYOUT
    bad = Primitive(triList=prim.triList, idxList=prim.idxList, bvhList=prim.bvhList.copy())
    leaf = np.nonzero(bad.bvhList["pType"] == L.TRIANGLE)[0][0]
    bad.bvhList["pIndex"][leaf] = 10 ** 6                                # primitive index out of range
    assert _create(bad)[0] == ERR_LAYOUT
    rc = lib.trq_scene_create(None, 0, None)
    assert rc == ERR_INVALID


def test_depth_limit_of_the_32_bit_trail(built):
    """A degenerate 40-deep chain must be refused: Scene::hit's trail has 32 bits (Render.hh:140,172)."""
    n = 41
    bvh = np.zeros(2 * n - 1, dtype=L.bvh_dtype)
    bvh["pType"] = L.TRIANGLE
    # chain: interior k (index 2k) -> leaf (2k+1) and interior k+1 (2k+2); root is index 0
    for k in range(n - 1):
        i = 2 * k
        bvh[i]["pType"], bvh[i]["left"], bvh[i]["right"] = L.BVH, i + 1, i + 2
        bvh[i + 1]["parent"] = bvh[i + 2]["parent"] = i
    tri = H.make_vertices(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32), [[0, 1, 2]])
    prim = Primitive(triList=tri, idxList=np.array([0, 1, 2], dtype=np.uint32), bvhList=bvh)
    rc, _ = _create(prim)
    assert rc == ERR_DEPTH and b"trail" in lib.trq_last_error_string()


def test_no_device_means_error_not_fallback(built):
    if lib.trq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    prim = H.scene_soup(50, seed=1, extent=0.2)
    rc, h = _create(prim)
    assert rc == ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.trq_last_error_string()
    from tracer_b200 import Scene
    with pytest.raises(_lib.TrqError):
        Scene(prim, 0)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under tracer_b200/ or include/ may reference it."""
    for base in ("tracer_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                    src = open(os.path.join(dirpath, f), errors="replace").read()
                    for needle in ("from oracle", "import oracle", "pyoracle", "liboracle", "libtracer_ref", "oracle_rq.h", "orq_"):
                        assert needle not in src, f"{dirpath}/{f} references the oracle ({needle})"


def test_gather_argument_validation_needs_no_gpu(built):
    """The peer-memory gather refuses bad arguments with a status and a message (no exception, no crash)."""
    import ctypes as C
    from tracer_b200._lib import lib
    g = C.c_void_p()
    handle = C.create_string_buffer(64)
    assert lib.trq_gather_create(None, 0, 2, 1000, C.byref(g), handle) != 0
    assert b"NULL" in lib.trq_last_error_string()
    assert lib.trq_gather_connect(None, None) != 0
    assert lib.trq_trace_gather(None, None, None, 0, 0, None) != 0
    assert lib.trq_gather_wait(None, None, None, None) != 0
    assert lib.trq_gather_status(None) != 0
    assert lib.trq_gather_destroy(None) == 0


def test_mgpu_helper_fails_loudly_without_a_gpu(built):
    import ctypes as C
    import torch
    from tracer_b200 import harness as H
    from tracer_b200._lib import ERR_INVALID, lib
    assert lib.trq_mgpu_create(None, None, 0, None) == ERR_INVALID
    assert lib.trq_mgpu_device_count(None) == 0 and lib.trq_mgpu_destroy(None) == 0
    if not torch.cuda.is_available():
        from tracer_b200._lib import ERR_NO_DEVICE
        prim = H.scene_soup(50, seed=1, extent=0.2)
        h = C.c_void_p()
        d = prim.desc()
        assert lib.trq_mgpu_create(C.byref(d), None, 0, C.byref(h)) == ERR_NO_DEVICE
        assert b"no CPU fallback" in lib.trq_last_error_string()


def test_round2_entry_points_fail_loudly_without_a_device_or_arguments(built):
    """The device-resident entry points validate their arguments on the host and report TRQ_ERR_NO_DEVICE on a box
    without a GPU: nothing falls back to the CPU."""
    import torch
    prim = H.scene_soup(20, seed=1, extent=0.2)
    d = prim.desc()
    h = C.c_void_p(None)
    assert lib.trq_scene_create_device(None, 0, C.byref(h)) == ERR_INVALID
    empty = SceneDesc()
    assert lib.trq_scene_create_device(C.byref(empty), 0, C.byref(h)) == ERR_INVALID and b"empty bvhList" in lib.trq_last_error_string()
    assert lib.trq_scene_update_vertices(None, None, 0, 0) == ERR_INVALID
    assert lib.trq_bvh_build_tree_device(None, 0, 0, None, None) == ERR_INVALID
    assert lib.trq_scene_set_kernel_config(None, 0, None) == ERR_INVALID
    assert lib.trq_scene_kernel_config(None) == -1
    assert lib.trq_kernel_config_count() >= 2 and lib.trq_kernel_config_name(0).startswith(b"256x5")
    assert lib.trq_kernel_config_name(99) is None
    if not torch.cuda.is_available():
        # host pointers are not device pointers, but the device check comes first: no GPU, no work
        assert lib.trq_scene_create_device(C.byref(d), 0, C.byref(h)) == ERR_NO_DEVICE
        nodes = np.zeros(2 * 20 - 1, dtype=L.bvh_dtype)
        assert lib.trq_bvh_build_tree_device(nodes.ctypes.data, 20, 0, None, None) == ERR_NO_DEVICE
        assert b"no CUDA device" in lib.trq_last_error_string()
