#!/usr/bin/env python
"""Developer tool (needs a -DTRQ_STATS build of the library): lane utilisation of the packed kernel's phases."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tracer_b200 import Scene, harness as H, layout as L, rays_to_torch
from tracer_b200._lib import lib
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
if name == "c3":
    prim = H.scene_c3(2); scene = Scene(prim, 0)
    d = rays_to_torch(H.cornell_camera_rays(3840, 2160), "cuda:0")
    recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
    rays = H.bounce_rays(recs)[0]
else:
    prim = H.scene_soup(1_000_000, 1, 0.01); scene = Scene(prim, 0); rays = H.random_rays(8_000_000, seed=2)
d = rays_to_torch(rays, "cuda:0")
out = (C.c_ulonglong * 8)()
scene.hit(d); torch.cuda.synchronize(); lib.trq_debug_stats(out, 1)
scene.hit(d); torch.cuda.synchronize(); lib.trq_debug_stats(out, 1)
s = list(out)
print(f"{name}: rays {rays.size}; interior warp-steps {s[0]} lanes/step {s[1]/max(1,s[0]):.2f} ({s[1]/rays.size:.1f} lane-steps/ray); "
      f"leaf warp-steps {s[2]} lanes/step {s[3]/max(1,s[2]):.2f} ({s[3]/rays.size:.2f}/ray); refills {s[4]} lanes/refill {s[5]/max(1,s[4]):.2f}")
