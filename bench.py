#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-query hot path (contract: see DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c4|c5|soup1m] [--impl reference]

A "step" is one pass of the hot path (trq_trace, closest-hit) over one batch of synthetic rays of the
named BASELINE config. Default workload = C3 (the config the north_star target is quoted on): the
~1.0 M-triangle "meshes" scene (coatball + teapot subdivided twice inside the Cornell box) traced with
the incoherent diffuse-bounce rays spawned from the 3840x2160 primary hits.

  value   whole-job Mrays/s, rays and hit buffers resident in HBM, CUDA events, max over ranks
  e2e     the same metric through the C-ABI with HOST buffers (pinned), H2D + D2H inside the timed region
  roofline  traversal kernel only: algorithmic bytes (instrumented-oracle step counts, SURVEY 8d formula)
            / its CUDA-event duration, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own Scene::hit (oracle/_ref, compiled verbatim) on the host cores, on a
            bounded strided sample of the same rays (N=1, rank 0 only)

`--impl reference` times that CPU reference alone (rank 0; other ranks exit).
Multi-GPU: one process per GPU (torchrun), scene broadcast from rank 0 over NCCL, rays sharded by batch
(each rank traces its own bounce-ray batch: weak scaling), no collective on the data path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec closest-hit"
UNIT = "Mrays/s"
HBM_FALLBACK_GBS = 6650.0

WORKLOADS = {
    "c2": "C2: RT_Metal Cornell box triangles + teapot (15.7k tris), 1920x1080 primary + 1 diffuse bounce, device-resident wavefront (cast, trace, spawn, trace)",
    "c2bounce": "C2 geometry, only the diffuse bounce batch spawned from the 1920x1080 primary hits",
    "c3path": "C3 geometry (1.0M tris), 3840x2160 primary + 1 diffuse bounce, device-resident wavefront (cast, trace, spawn, trace)",
    "c3": "C3: RT_Metal meshes scene (coatball+teapot subdivided x16, 1.0M tris), incoherent diffuse bounce rays from 3840x2160 primary hits",
    "c4": "C4: C3 geometry + 10,012 spheres, any-hit shadow rays toward light squares 5/6 from 3840x2160 primary hits",
    "c5": "C5: 10M-triangle random soup, 8M uniform incoherent rays per GPU",
    "soup1m": "1M-triangle random soup, 8M uniform incoherent rays",
    "c1": "C1: RT_Nextweek randomScene (442 spheres) as Sphere leaves of the RT_Metal SAH BVH, 1280x720 primary rays (camera0)",
}
# algorithmic bytes per ray measured once by the instrumented oracle; used only if the
# oracle cannot be run in this process. Recomputed live on the cpu_baseline sample otherwise.
# device-resident wavefronts: name -> image size (primary rays per step); the Cornell camera of Tracer.mm:371-411
PATH_WORKLOADS = {"c2": (1920, 1080), "c3path": (3840, 2160)}


def L_ray_dtype():
    from tracer_b200 import layout
    return layout.ray_dtype


BYTES_PER_RAY_FALLBACK = {"c3path": 1400.0, "c2bounce": 1094.0, "c1": 700.0, "c2": 1094.0, "c3": 1778.0, "c4": 2887.0, "c5": 9184.0, "soup1m": 7039.0}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- workloads
def build_scene(name):
    from tracer_b200 import harness as H
    if name in ("c2", "c2bounce"):
        return H.scene_c2()
    if name in ("c3", "c3path"):
        return H.scene_c3(2)
    if name == "c4":
        return H.scene_c4(2)
    if name == "c5":
        return H.scene_soup(10_000_000, seed=1, extent=0.004)
    if name == "soup1m":
        return H.scene_soup(1_000_000, seed=1, extent=0.01)
    if name == "c1":
        return H.scene_c1()
    raise SystemExit(f"unknown workload {name}")


def c1_rays():
    """RT_Nextweek camera0 (Render.swift:35-57): from (13,2,3) to the origin, vfov 20 deg, 1280x720."""
    from tracer_b200 import harness as H
    return H.camera_rays((13, 2, 3), (0, 0, 0), np.float32(20 * np.pi / 180), 1280, 720)


def make_rays_gpu(name, prim, scene, rank, device):
    """Ray batch of this rank for the named workload (generation is not timed)."""
    from tracer_b200 import harness as H, layout as L, rays_to_torch
    if name in ("c5", "soup1m"):
        n = 8_000_000
        return H.random_rays(n, seed=2, first=rank * n), False
    if name == "c1":
        return c1_rays(), False
    W, Hh = (1920, 1080) if name == "c2bounce" else (3840, 2160)
    primary = H.cornell_camera_rays(W, Hh)
    d = rays_to_torch(primary, device)
    recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
    seed_base = rank << 32
    if name == "c4":
        la, lb = H.scene_c4_lights(prim)
        rays, _ = H.shadow_rays(recs, la, lb, seed_base)
        return rays, True
    rays, _ = H.bounce_rays(recs, seed_base)
    return rays, False


def make_rays_cpu_sample(name, prim, ref_trace_records, stride):
    """Strided sample of the same workload generated WITHOUT the GPU (reference arm): primary pixels are
    subsampled by `stride`, their hits come from the CPU reference, bounce/shadow rays are spawned from those."""
    from tracer_b200 import harness as H
    if name in ("c5", "soup1m"):
        return H.random_rays(8_000_000 // stride, seed=2), False
    if name == "c1":
        return c1_rays(), False
    W, Hh = (1920, 1080) if name in ("c2", "c2bounce") else (3840, 2160)
    primary = H.cornell_camera_rays(W, Hh)[::stride].copy()
    recs = ref_trace_records(primary)
    if name == "c4":
        la, lb = H.scene_c4_lights(prim)
        return H.shadow_rays(recs, la, lb, 0)[0], True
    bounce = H.bounce_rays(recs, 0)[0]
    if name in PATH_WORKLOADS:                                  # both waves of the wavefront
        return np.concatenate([primary, bounce]), False
    return bounce, False


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up until after the timed region; samples are kept when
    their timestamp falls inside [mark_start, mark_end] (the timed region), widened to the whole loaded span
    (warm-up + timed) when the timed region is too short to catch any."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="trq_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None
        self.t_load = time.time()
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": None}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(p[1]), float(p[2]), [v.lower().startswith("active") for v in p[5:9]]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for window, lo, hi in (("timed region", self.t0, self.t1), ("warm-up + timed region", self.t_load, self.t1)):
            if lo is None or hi is None:
                continue
            sel = [r for r in rows if lo - 0.02 <= r[0] <= hi + 0.02]
            if sel:
                reasons = sorted({n for r in sel for n, on in zip(names, r[3]) if on})
                out.update({"sm_mhz": float(np.median([r[1] for r in sel])), "sm_max_mhz": max(r[2] for r in sel),
                            "reasons": reasons, "samples": len(sel), "window": window})
                break
        return out


# ----------------------------------------------------------------------------------------------- CPU reference
def cpu_reference(prim, rays, any_hit, budget_s, nthreads):
    """Time the reference's Scene::hit (verbatim build if present, else the C port) on `rays` subsampled so that one
    pass takes about budget_s. Returns (Mrays/s, kind, cores, sample description, bytes_per_ray or None)."""
    from oracle.pyoracle import Port, Reference
    port = Port()
    use_ref = Reference.available()
    ref = Reference() if use_ref else None

    def run(sub):
        t = time.perf_counter()
        if use_ref:
            ref.trace_lite(prim, sub, any=any_hit, nthreads=nthreads)
        else:
            port.trace(prim, sub, any=any_hit, nthreads=nthreads)
        return time.perf_counter() - t

    pilot = rays[:: max(1, rays.size // 40000)]
    run(pilot[: max(1, pilot.size // 4)])                       # page in
    dt = run(pilot)
    rate = pilot.size / max(dt, 1e-6)
    want = int(min(rays.size, max(pilot.size, rate * budget_s)))
    stride = max(1, rays.size // want)
    sub = np.ascontiguousarray(rays[::stride])
    passes, dt = 0, 0.0
    while passes < 12 and (passes == 0 or dt < 0.8 * budget_s):        # ~budget_s of CPU work even when one pass is short
        dt += run(sub)
        passes += 1
    tot = port.trace(prim, pilot, any=any_hit, nthreads=nthreads)["totals"]      # step counters (not timed)
    bpr = tot["bytes"] / max(1, tot["n_rays"])
    kind = "reference" if use_ref else "port"
    sample = (f"every ray of the batch" if stride == 1 else f"every {stride}th ray of the batch") + \
             f" ({sub.size} rays) x {passes} passes, {dt:.1f} s of CPU work"
    return sub.size * passes / dt / 1e6, kind, nthreads, sample, bpr, sub


# ----------------------------------------------------------------------------------------------- arms
def run_reference_arm(a):
    rank, _, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    if rank != 0:
        return 0
    from oracle.pyoracle import Port, Reference
    prim = build_scene(a.workload)
    nthreads = os.cpu_count() or 1
    use_ref = Reference.available()
    eng = Reference() if use_ref else Port()
    rec_fn = (lambda r: eng.trace(prim, r, nthreads=nthreads)) if use_ref else (lambda r: eng.trace(prim, r, records=True, nthreads=nthreads)["records"])
    rays, any_hit = make_rays_cpu_sample(a.workload, prim, rec_fn, stride=16)

    def run(sub):
        t = time.perf_counter()
        if use_ref:
            eng.trace_lite(prim, sub, any=any_hit, nthreads=nthreads)
        else:
            eng.trace(prim, sub, any=any_hit, nthreads=nthreads)
        return time.perf_counter() - t

    pilot = rays[:: max(1, rays.size // 20000)]
    rate = pilot.size / max(run(pilot), 1e-6)
    budget = min(3.0, 120.0 / max(1, a.steps + a.warmup))
    stride = max(1, int(rays.size / max(pilot.size, rate * budget)))
    sub = np.ascontiguousarray(rays[::stride])
    for _ in range(a.warmup):
        run(sub)
    t = sum(run(sub) for _ in range(a.steps))
    v = sub.size * a.steps / t / 1e6
    kind = "reference" if use_ref else "port"
    sample = f"{sub.size} rays per step: every {16 * stride}th ray of the batch (primary pixels subsampled x16, then stride {stride})"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(t / a.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(a.workload),
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def config_of(workload, extra=None):
    c = {"workload": WORKLOADS[workload], "query": "any-hit" if workload == "c4" else "closest-hit",
         "sharding": "ray batch per rank, scene replicated"}
    if extra:
        c.update(extra)
    return c


def run_gpu_arm(a):
    # stdout carries exactly ONE line (the JSON); anything libraries print there (e.g. "NCCL version ...") goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch

    from tracer_b200 import Scene, dist as D, launch_count, rays_to_torch
    rank, local_rank, world = D.init()
    if world != a.gpus:
        log(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"

    t0 = time.time()
    prim = build_scene(a.workload) if rank == 0 else None
    t_build = time.time() - t0
    prim = D.replicate_primitive(prim, src=0)
    scene = Scene(prim, local_rank)
    # TRQ_SORT_RAYS hint: only for the batch that is incoherent AND whose scene cannot live in L2 (C5)
    sort = (a.workload == "c5") if a.sort is None else bool(a.sort)
    path = a.workload in PATH_WORKLOADS
    launches_per_step = 1
    if path:
        # device-resident wavefront (BASELINE configs[1]: "primary + 1 diffuse bounce"): castRay for every pixel,
        # trace, spawn the diffuse bounce of every hit (compacted, count stays on the device), trace that too.
        W, Hh = PATH_WORKLOADS[a.workload]
        cam = ((278, 278, -800), (278, 278, 278), (0, 1, 0), np.float32(45 * (np.pi / 180)))
        r0 = torch.empty((W * Hh, 8), dtype=torch.float32, device=device); h0 = torch.empty_like(r0)
        r1 = torch.empty_like(r0); h1 = torch.empty_like(r0)
        s1 = torch.empty(W * Hh, dtype=torch.int32, device=device); c1 = torch.zeros(1, dtype=torch.int64, device=device)
        seed_base = rank << 32
        any_hit, launches_per_step = False, 2

        def step():
            scene.cast_rays(*cam, W, Hh, out=r0)
            scene.hit(r0, out=h0)
            scene.spawn_bounce(r0, h0, seed_base=seed_base, out=r1, src=s1, count=c1)
            scene.hit_indirect(r1, c1, out=h1)

        step(); torch.cuda.synchronize()
        n1 = int(c1.item())
        n = W * Hh + n1
        rays = np.concatenate([r0.cpu().numpy().view(L_ray_dtype()).reshape(-1), r1[:n1].cpu().numpy().view(L_ray_dtype()).reshape(-1)])
        d_hits = torch.cat([h0, h1[:n1]])
    else:
        rays, any_hit = make_rays_gpu(a.workload, prim, scene, rank, device)
        n = rays.size
        d_rays = rays_to_torch(rays, device)
        d_hits = torch.empty((n, 8), dtype=torch.float32, device=device)

        def step():
            scene.hit(d_rays, any=any_hit, out=d_hits, sort=sort)
    if rank == 0:
        log(f"# scene {scene.info}, build {t_build:.1f}s, {n} rays/step/rank")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if a.flush_l2 else None

    sampler = ClockSampler(local_rank)                  # runs from before the warm-up to the end of the timed region
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device events, barrier + synchronize on both sides
    scene.profile(True)
    launches0 = launch_count()
    D.barrier(); torch.cuda.synchronize()
    sampler.mark_start()
    total_ms = 0.0
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        total_ms = e0.elapsed_time(e1)
    else:
        evs = []
        for _ in range(a.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        total_ms = sum(x.elapsed_time(y) for x, y in evs)
    sampler.mark_end()
    D.barrier()
    launches = launch_count() - launches0
    clocks = sampler.stop()
    nl, trace_ms, resolve_ms = scene.profile_read()
    scene.profile(False)
    total_ms = D.max_over_ranks(total_ms)
    total_rays = D.sum_over_ranks(n)
    value = total_rays * a.steps / total_ms / 1e3

    # ---- e2e: host (pinned) rays in, host hits out, through the C-ABI host-pointer path
    h_hits = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(a.steps, 20))
    if path:
        # inputs are the camera (52 bytes); both hit buffers come back to pinned host memory every step
        def e2e_step():
            step()
            h_hits[: W * Hh].copy_(h0, non_blocking=True)
            h_hits[W * Hh:].copy_(h1[:n1], non_blocking=True)
            torch.cuda.synchronize()
        h2d_bytes, d2h_bytes = 52, int(n * 32)
    else:
        h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).pin_memory()

        def e2e_step():
            scene.hit_host(h_rays.data_ptr(), n, h_hits.data_ptr(), any=any_hit, sort=sort)
        h2d_bytes, d2h_bytes = int(n * 32), int(n * 32)
    for _ in range(2):
        e2e_step()
    D.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t)
    e2e_value = total_rays * e2e_steps / e2e_s / 1e6
    # same traffic, but the K calls are queued back to back (TRQ_HOST_ASYNC, two alternating pinned result buffers)
    # so that one call's D2H overlaps the next call's H2D; collected once at the end
    e2e_pipe = None
    if not path:
        h_hits2 = torch.empty((n, 8), dtype=torch.float32).pin_memory()
        outs = (h_hits, h_hits2)
        for k in range(2):
            scene.hit_host(h_rays.data_ptr(), n, outs[k].data_ptr(), any=any_hit, sort=sort, asynchronous=True)
        scene.host_sync()
        D.barrier(); torch.cuda.synchronize()
        t = time.perf_counter()
        for k in range(e2e_steps):
            scene.hit_host(h_rays.data_ptr(), n, outs[k & 1].data_ptr(), any=any_hit, sort=sort, asynchronous=True)
        scene.host_sync()
        pipe_s = D.max_over_ranks(time.perf_counter() - t)
        e2e_pipe = {"value": round(total_rays * e2e_steps / pipe_s / 1e6, 2), "unit": UNIT, "steps": e2e_steps,
                    "note": "K TRQ_HOST_ASYNC calls queued back to back + one trq_host_sync; same bytes per step as e2e",
                    "equals_device_path": bool(torch.equal(h_hits2, d_hits.cpu()))}
    # the host path must produce the same bytes as the device path
    if path:
        d_hits = torch.cat([h0, h1[:n1]])
    same = bool(torch.equal(h_hits, d_hits.cpu()))
    hit_frac = float((d_hits[:, 7].view(torch.int32) & 1).float().mean())

    # ---- multi-GPU, second figure (SURVEY 8d "with and without hit all-gather"): every rank ends up with every rank's
    # hit records; the all-gather runs chunk by chunk on a second stream under the tracing of the next chunk.
    with_gather = None
    if world > 1 and not path:
        n_common = int(-D.max_over_ranks(-float(n)))                      # ranks trace slightly different batch sizes
        g_rays = d_rays[:n_common].contiguous()
        g_local = torch.empty((n_common, 8), dtype=torch.float32, device=device)
        g_all = torch.empty((world, n_common, 8), dtype=torch.float32, device=device)
        g_steps = max(3, min(a.steps, 50))
        for _ in range(3):
            D.trace_and_gather(scene, g_rays, g_local, g_all, any=any_hit, sort=sort)
        D.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(g_steps):
            D.trace_and_gather(scene, g_rays, g_local, g_all, any=any_hit, sort=sort)
        e1.record(); torch.cuda.synchronize()
        g_ms = D.max_over_ranks(e0.elapsed_time(e1))
        ok = bool(torch.equal(g_all[rank], d_hits[:n_common]))
        nccl = {"value": round(n_common * world * g_steps / g_ms / 1e3, 2), "ms_per_step": round(g_ms / g_steps, 4),
                "chunks": 4, "own_shard_matches": ok, "how": "dist.trace_and_gather: NCCL all-gather per chunk on a second stream"}
        del g_all, g_local
        # the same result through trq_trace_gather: the resolve kernel stores every record into every rank's buffer
        # over NVLink peer memory and publishes (count, step); no NCCL on the data path
        try:
            hg = D.HitGather(scene, n_common)
            for _ in range(3):
                hg.trace(g_rays, any=any_hit, sort=sort); hg.wait()
            D.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(g_steps):
                hg.trace(g_rays, any=any_hit, sort=sort)
                p_all, p_counts = hg.wait()
            e1.record(); torch.cuda.synchronize()
            published = True
            try:
                hg.status()
            except RuntimeError:                                  # a peer never published: reported, and no rank leaves the collectives below
                published = False
            p_ms = D.max_over_ranks(e0.elapsed_time(e1))
            others = [r for r in range(world) if r != rank]
            full = published and bool((p_counts == n_common).all()) and bool(torch.equal(p_all[rank].view(torch.int32), d_hits[:n_common].view(torch.int32)))
            # slot r must hold rank r's records: compare a slice with an NCCL gather of the same slice
            m = min(n_common, 1 << 18)
            ref = D.gather_hits(d_hits[:m].contiguous())
            full = full and all(bool(torch.equal(p_all[r, :m].view(torch.int32), ref[r].view(torch.int32))) for r in others)
            full = bool(-D.max_over_ranks(-float(full)) == 1.0)
            with_gather = {"value": round(n_common * world * g_steps / p_ms / 1e3, 2), "unit": UNIT, "steps": g_steps,
                           "ms_per_step": round(p_ms / g_steps, 4), "gathered_bytes_per_rank_per_step": int(n_common * world * 32),
                           "how": "trq_trace_gather: all-gather fused into the resolve kernel (peer stores over NVLink + release/acquire flags)",
                           "all_slots_match": full, "nccl_overlapped": nccl}
            hg.close()
        except RuntimeError as e:                                 # raised on every rank alike (HitGather agrees on failures)
            with_gather = dict(nccl, unit=UNIT, steps=g_steps, gathered_bytes_per_rank_per_step=int(n_common * world * 32),
                               peer_gather_error=str(e))

    # ---- multi-GPU: scene checksum agreement + gathered hit count (not timed)
    gathered = None
    if world > 1:
        cs = D.checksum_primitive(prim)
        lo, hi = D.max_over_ranks(float(cs % (1 << 52))), -D.max_over_ranks(-float(cs % (1 << 52)))
        assert lo == hi, "scene broadcast checksum differs across ranks"
        parts = D.gather_hits(d_hits[: min(n, 1 << 20)])
        gathered = int(sum(p.shape[0] for p in parts))

    if rank != 0:
        D.barrier()
        _shutdown()
        return 0

    # ---- CPU baseline + algorithmic bytes (rank 0, N=1 only for the baseline)
    peak, peak_src = peaks()
    bpr, cpu = BYTES_PER_RAY_FALLBACK[a.workload], None
    try:
        if world == 1 and not a.no_cpu_baseline:
            v, kind, cores, sample, bpr, sub = cpu_reference(prim, rays, any_hit, a.cpu_budget, os.cpu_count() or 1)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
            # parity spot-check on the timed sample: GPU hits of the sample rays vs the oracle
            from oracle.pyoracle import Port
            chk = sub[:: max(1, sub.size // 8_000_000)]          # the whole timed sample (for C3: every ray of the batch)
            want = Port().trace(prim, chk, any=any_hit, nthreads=cores)["hits"]
            got = scene.hit(rays_to_torch(chk, device), any=any_hit).cpu().numpy().view(want.dtype).reshape(-1)
            ids_ok = all(np.array_equal(got[k], want[k]) for k in ("flags", "pType", "pIndex", "leafNode"))
            t_ok = np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
            tri = want["pType"] == 3                                  # barycentrics: bit-exact for triangle hits
            uv_ok = all(np.array_equal(got[k][tri].view(np.uint32), want[k][tri].view(np.uint32)) for k in ("u", "v"))
            cpu["parity_on_sample"] = {"rays": int(chk.size), "ids_bit_exact": bool(ids_ok), "t_bit_exact": bool(t_ok),
                                       "barycentrics_bit_exact": bool(uv_ok)}
            if a.workload == "c1":
                # the config's NAMED baseline: RT_Nextweek's own CPU BVH (restated in C, oracle/nextweek_bvh.c)
                from oracle.pyoracle import Nextweek
                nw = Nextweek(prim.sphereList)
                nw.trace(rays[:50000], nthreads=cores)
                t0 = time.perf_counter(); ids, _ = nw.trace(rays, nthreads=cores); dt = time.perf_counter() - t0
                mine = scene.hit(rays_to_torch(rays, device)).cpu().numpy().view(want.dtype).reshape(-1)
                hit = (mine["flags"] & 1) == 1
                cpu["nextweek_bvh"] = {"value": round(rays.size / dt / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                       "sample": f"all {rays.size} rays, {dt:.1f} s",
                                       "same_sphere_as_gpu": round(float(np.mean(ids[hit] == mine["pIndex"][hit])), 6)}
        else:
            from oracle.pyoracle import Port
            pilot = rays[:: max(1, n // 40000)]
            tot = Port().trace(prim, pilot, any=any_hit, nthreads=os.cpu_count() or 1)["totals"]
            bpr = tot["bytes"] / tot["n_rays"]
    except Exception as e:  # the oracle is test infrastructure: never let it take the GPU numbers down
        log(f"# cpu baseline unavailable: {e!r}")

    # the memory system as measured on this GPU right now (L2 has no entry in MEASURED_PEAKS.json)
    try:
        from tracer_b200 import probe_bandwidth
        l2_gbs, hbm_read_gbs = probe_bandwidth(local_rank, "l2"), probe_bandwidth(local_rank, "hbm")
    except Exception as e:
        log(f"# bandwidth probe failed: {e!r}")
        l2_gbs = hbm_read_gbs = None
    trace_ms_avg = trace_ms / max(1, nl) * launches_per_step          # traversal-kernel time per step
    achieved = bpr * n / (trace_ms_avg * 1e-3) / 1e9 if trace_ms_avg > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(a.workload)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": round(total_ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(a.workload, {
            "rays_per_step_per_gpu": int(n), "triangles": int(prim.nTri), "bvh_nodes": int(prim.bvhList.size),
            "hit_fraction": round(hit_frac, 4),
            "ray_ordering": "TRQ_SORT_RAYS (origin cell x direction octant, inside the timed step)" if sort else "as given",
            "l2": ("flushed between steps (256 MB write)" if a.flush_l2 else
                   f"no flush: rays+hits stream {2 * n * 32 / 1e6:.0f} MB per step (> 126 MB L2)"),
        }),
        "roofline": {"bound": "hbm", "kernel": "trace_packed_kernel", "achieved": None if achieved is None else round(achieved, 1),
                     "peak": peak, "unit": "GB/s", "frac": None if achieved is None else round(achieved / peak, 4),
                     "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
                     "algorithmic_gb_per_launch": round(bpr * n / 1e9, 3), "peak_source": peak_src, "algorithmic_bytes_per_ray": round(bpr, 1),
                     "kernel_ms": round(trace_ms_avg, 4), "resolve_kernel_ms": round(resolve_ms / max(1, nl) * launches_per_step, 4),
                     "trace_launches_per_step": launches_per_step,
                     "l2_read_peak_gbs_measured": None if l2_gbs is None else round(l2_gbs, 1),
                     "frac_of_l2_peak": None if (l2_gbs is None or achieved is None) else round(achieved / l2_gbs, 4),
                     "hbm_read_gbs_measured": None if hbm_read_gbs is None else round(hbm_read_gbs, 1),
                     "kernel_share_of_step": round(trace_ms_avg / max(1e-9, total_ms / a.steps), 4)},
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps, "host_path_equals_device_path": same, "timing": "wall clock around K synchronous C-ABI calls, max over ranks"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e_pipe is not None:
        line["e2e_pipelined"] = e2e_pipe
    if gathered is not None:
        line["gathered_hits_checked"] = gathered
    if with_gather is not None:
        line["with_hit_allgather"] = with_gather
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    D.barrier()
    _shutdown()
    return 0


def _shutdown():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--sort", type=int, default=None, help="force the TRQ_SORT_RAYS hint on (1) / off (0); default: on for c5 only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_gpu_arm(a)


if __name__ == "__main__":
    sys.exit(main())
