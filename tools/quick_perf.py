#!/usr/bin/env python
"""Developer tool: time trq_trace on a few workloads (device-resident rays, CUDA events).
Usage: python tools/quick_perf.py [c2|c3|soup1m|soup10m ...] [--reflayout] [--iters N]
Prints one JSON line per workload. Not the contract bench (that is bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tracer_b200 import Scene, harness as H, layout as L, rays_to_torch  # noqa: E402


def workloads(name):
    if name == "c2":
        prim = H.scene_c2()
        scene = Scene(prim, 0)
        prim_rays = H.cornell_camera_rays(1920, 1080)
        d = rays_to_torch(prim_rays, "cuda:0")
        recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
        bounce, _ = H.bounce_rays(recs)
        return prim, scene, {"primary": prim_rays, "bounce": bounce}
    if name == "refcornell":
        prim = H.scene_reference_cornell()
        scene = Scene(prim, 0)
        prim_rays = H.cornell_camera_rays(3840, 2160)
        d = rays_to_torch(prim_rays, "cuda:0")
        recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
        bounce, _ = H.bounce_rays(recs)
        return prim, scene, {"primary": prim_rays, "bounce": bounce}
    if name.startswith("c3"):
        levels = 2 if name == "c3" else int(name[2:])
        t = time.time(); prim = H.scene_c3(levels); tb = time.time() - t
        scene = Scene(prim, 0)
        print(f"# c3 build {tb:.1f}s, {prim.nTri} tris, depth {scene.info['maxDepth']}", file=sys.stderr)
        prim_rays = H.cornell_camera_rays(3840, 2160)
        d = rays_to_torch(prim_rays, "cuda:0")
        recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
        bounce, _ = H.bounce_rays(recs)
        return prim, scene, {"primary": prim_rays, "bounce": bounce}
    if name.startswith("soup"):
        n = {"soup1m": 1_000_000, "soup10m": 10_000_000, "soup100k": 100_000, "soup4m": 4_000_000}[name]
        t = time.time(); prim = H.scene_soup(n, seed=1, extent=0.004 if n >= 10_000_000 else 0.01); tb = time.time() - t
        scene = Scene(prim, 0)
        print(f"# soup build {tb:.1f}s depth {scene.info['maxDepth']}", file=sys.stderr)
        return prim, scene, {"random": H.random_rays(8_000_000, seed=2)}
    raise SystemExit(f"unknown workload {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=["c2"])
    ap.add_argument("--reflayout", action="store_true")
    ap.add_argument("--any", action="store_true")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--counters", action="store_true", help="also run the CPU oracle on a 1%% sample for bytes/ray")
    a = ap.parse_args()
    for name in a.names:
        prim, scene, sets = workloads(name)
        for tag, rays in sets.items():
            d = rays_to_torch(rays, "cuda:0")
            out = torch.empty((d.shape[0], 8), dtype=torch.float32, device="cuda:0")
            for _ in range(2):
                scene.hit(d, any=a.any, out=out, reflayout=a.reflayout)
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); scene.hit(d, any=a.any, out=out, reflayout=a.reflayout); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            res = {"workload": f"{name}/{tag}", "rays": int(d.shape[0]), "ms": round(ms, 3),
                   "mrays_s": round(d.shape[0] / ms / 1e3, 1), "kernel": "reflayout" if a.reflayout else "packed",
                   "any": a.any, "hit_frac": round(float((out[:, 7].view(torch.int32) & 1).float().mean()), 4)}
            if a.counters:
                from oracle.pyoracle import Port
                sub = rays[:: max(1, rays.size // 50000)]
                tot = Port().trace(prim, sub, any=a.any, nthreads=os.cpu_count())["totals"]
                bpr = tot["bytes"] / tot["n_rays"]
                res.update({"bytes_per_ray": round(bpr, 1), "n_fp": round(tot["n_fp"] / tot["n_rays"], 2),
                            "n_tri": round(tot["n_tri"] / tot["n_rays"], 2), "alg_GBs": round(bpr * d.shape[0] / ms / 1e6, 1)})
            print(json.dumps(res), flush=True)
        scene.close()


if __name__ == "__main__":
    main()
