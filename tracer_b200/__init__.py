"""tracer_b200 -- B200-native ray queries for iaomw/Tracer scenes (the RT_Metal Scene::hit hot path).

Public surface (host mirror of the reference interface, over the C-ABI in include/tracer_rq.h):
    Primitive, Scene, MultiGpuScene, BVHBuilder   tracer_b200.scene
    layout dtypes / flags                 tracer_b200.layout
    workload builders (bench, tests)      tracer_b200.harness
    multi-GPU sharding helpers            tracer_b200.dist
"""
from . import layout
from .scene import BVHBuilder, DevicePrimitive, MultiGpuScene, Primitive, Scene, hits_to_numpy, launch_count, probe_bandwidth, rays_to_torch

__all__ = ["layout", "Primitive", "DevicePrimitive", "Scene", "MultiGpuScene", "BVHBuilder", "hits_to_numpy", "rays_to_torch", "launch_count", "probe_bandwidth"]
