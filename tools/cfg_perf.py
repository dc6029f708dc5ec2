#!/usr/bin/env python
"""Developer tool: Mrays/s of every launch configuration of the traversal kernel (trq_scene_set_kernel_config) on a few
workloads, device-resident rays, CUDA events, median of --iters launches.
Usage: python tools/cfg_perf.py [c1 c2 c3 c4 soup1m soup10m ...] [--iters N] [--hit16] [--top N]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tracer_b200 import Scene, harness as H, layout as L, rays_to_torch  # noqa: E402


def workloads(name):
    """-> (prim, scene, {tag: (rays, any, sort)})"""
    if name == "c1":
        prim = H.scene_c1(); scene = Scene(prim, 0)
        return prim, scene, {"primary": (H.camera_rays((13, 2, 3), (0, 0, 0), np.float32(20 * np.pi / 180), 1280, 720), False, False)}
    if name in ("c2", "c3", "c4", "refcornell"):
        prim = {"c2": H.scene_c2, "c3": lambda: H.scene_c3(2), "c4": lambda: H.scene_c4(2), "refcornell": H.scene_reference_cornell}[name]()
        scene = Scene(prim, 0)
        W, Hh = (1920, 1080) if name == "c2" else (3840, 2160)
        primary = H.cornell_camera_rays(W, Hh)
        d = rays_to_torch(primary, "cuda:0")
        recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
        if name == "c4":
            la, lb = H.scene_c4_lights(prim)
            return prim, scene, {"shadow_any": (H.shadow_rays(recs, la, lb, 0)[0], True, False)}
        return prim, scene, {"primary": (primary, False, False), "bounce": (H.bounce_rays(recs)[0], False, False)}
    if name.startswith("soup"):
        n = {"soup1m": 1_000_000, "soup10m": 10_000_000, "soup100k": 100_000, "soup4m": 4_000_000}[name]
        prim = H.scene_soup(n, seed=1, extent=0.004 if n >= 10_000_000 else 0.01)
        scene = Scene(prim, 0)
        rays = H.random_rays(8_000_000, seed=2)
        sets = {"random": (rays, False, False)}
        if n >= 10_000_000:
            sets["random_sorted"] = (rays, False, True)
            sets["random_auto"] = (rays, False, None)                # no hint: the library examines the batch (large tree)
            cam = H.camera_rays((0.5, 0.5, -1.5), (0.5, 0.5, 0.5), np.float32(0.7), 3840, 2160)
            sets["camera"] = (cam, False, False)
            sets["camera_auto"] = (cam, False, None)
        return prim, scene, sets
    raise SystemExit(f"unknown workload {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=["c3"])
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--hit16", action="store_true")
    ap.add_argument("--cfgs", default="")
    ap.add_argument("--tops", default="", help="comma-separated TRQ_TOP_NODES values: the scene is created once per value")
    a = ap.parse_args()
    tops = [int(x) for x in a.tops.split(",")] if a.tops else [None]
    names = Scene.kernel_configs()
    pick = [int(x) for x in a.cfgs.split(",")] if a.cfgs else list(range(len(names)))
    for name in a.names:
        prim, scene, sets = workloads(name)
        print(f"# {name}: {scene.info}", file=sys.stderr, flush=True)
        for tag, (rays, any_hit, sort) in sets.items():
            d = rays_to_torch(rays, "cuda:0")
            out = torch.empty((d.shape[0], 4 if a.hit16 else 8), dtype=torch.float32, device="cuda:0")
            base = None
            for top, c in [(t, c) for t in tops for c in pick]:
                if top is not None:
                    if not names[c].split()[1].endswith("true") and top != tops[0]:
                        continue                                # configurations without staging: once
                    os.environ["TRQ_TOP_NODES"] = str(top)
                    scene.close()
                    scene = Scene(prim, 0)
                try:
                    staged = scene.set_kernel_config(c)
                except Exception as e:
                    print(json.dumps({"workload": f"{name}/{tag}", "cfg": names[c], "error": str(e)[:80]}), flush=True)
                    continue
                for _ in range(2):
                    scene.hit(d, any=any_hit, out=out, sort=sort, hit16=a.hit16)
                torch.cuda.synchronize()
                ts = []
                for _ in range(a.iters):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); scene.hit(d, any=any_hit, out=out, sort=sort, hit16=a.hit16); e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                chk = int(out.view(torch.int32).sum(dtype=torch.int64).item())      # every configuration must agree
                base = chk if base is None else base
                print(json.dumps({"workload": f"{name}/{tag}", "cfg": names[c], "staged_nodes": staged, "rays": int(d.shape[0]),
                                  "ms": round(ms, 4), "mrays_s": round(d.shape[0] / ms / 1e3, 1), "same_result": chk == base}), flush=True)
        scene.close()


if __name__ == "__main__":
    main()
