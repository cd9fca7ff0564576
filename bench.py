#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Mrays/s at 1080p, primary + shadow + AO rays,
on the synthetic 100 M-triangle scene (BASELINE.json configs[4], SURVEY 8d "C5"), N GPUs of one
node, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W                (our arm, CUDA through the C ABI)
  python bench.py --impl reference ...                         (CPU arm on the host cores: the reference's own Embree 3.6.1 from
                                                                oracle/_ref doing every nearest-hit query of the full scene inside
                                                                the oracle's restatement of the ISPC glue; the scalar oracle port
                                                                only if that library is absent)
  torchrun ... bench.py --gpus N ...                           (N>1: spatial partitions + NCCL)

A step is one frame: generation -> trace waves -> shading/secondary rays -> classify -> ray
forwarding -> additive framebuffer.  value = (primary + shadow + AO rays of the frame) / frame time
(re-traces of forwarded rays are not counted twice).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # frames in flight: one hardware queue per stream (before torch creates the context)

from galaxy_b200 import scenes  # noqa: E402

W, H = 1920, 1080
EPS = 0.001
# ncu --set full, one frame of this workload on one B200 (profiles/r01_e_trace_kernels_full.txt):
# primary_trace_kernel 480.3 MB read + 10.6 MB written, fused_secondary_kernel 501.2 MB read + 10.9 MB written
NCU_TRAFFIC_BYTES_PER_FRAME = 1.0030e9
# volume workloads at 1024^3 (profiles/r01_e_volume_full.txt): dram bytes of the march launches of one frame
NCU_VOLUME_TRAFFIC = {"c3": 4.296e9, "c4": 6.868e9}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (samples are time-stamped
    on arrival; only those inside [mark_begin, mark_end] are used)."""

    def __init__(self, index):
        self.index, self.proc, self.samples = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        inside = [l for (t, l) in self.samples if self.t0 is not None and self.t0 <= t <= self.t1]
        note = None
        if not inside and self.samples:  # region shorter than the sampling period: nearest samples
            mid = 0.5 * (self.t0 + self.t1)
            inside = [l for (t, l) in sorted(self.samples, key=lambda s: abs(s[0] - mid))[:3]]
            note = "timed region shorter than the sampling period; nearest samples used"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def c5_alg_bytes_per_ray(n_prims_in_partition, hit_fraction):
    """B_alg = 52 + 24 + 28 h (ray I/O) + L*80 (L = ceil(log8(N/4)) 8-wide nodes) + 144 (one 4-triangle leaf)."""
    import math
    L = max(1, math.ceil(math.log(max(n_prims_in_partition, 8) / 4.0, 8)))
    return 76.0 + 28.0 * hit_fraction + 80.0 * L + 144.0, L


def cpu_sample_scene(oracle_mod, n_lat, n_lon):
    ds, _ = scenes.c5_partition_mesh(n_lat, n_lon, 1, 0)
    vis = scenes.c5_vis()
    parts = scenes.build_partitions(oracle_mod, vis, {"mesh": ds}, 1)
    return parts, vis, len(ds.indices)


CMAP = [[0.0, 1.0, 0.5, 0.5], [0.25, 0.5, 1.0, 0.5], [0.5, 0.5, 0.5, 1.0], [0.75, 1.0, 1.0, 0.5], [1.0, 1.0, 0.5, 1.0]]


def synth_volume(n, device="cuda"):
    """C3/C4 throughput volume: v = eightBalls(p) + 0.15 * fbm(8p), 5-octave value noise on a PCG32-hashed lattice (seed 7),
    n^3 float32 on [-1,1]^3; every partition's brick is synthesised on demand, slab by slab, on the device (scenes.noise_volume)."""
    import torch
    return scenes.noise_volume(n, device if (device != "cpu" and torch.cuda.is_available()) else "cpu")


def volume_case(which):
    """(visualization, camera) of C3 = DVR only (examples/noise.state) and C4 = isosurface + shadow rays
    (examples/noise_isovalue.state), SURVEY 8d."""
    if which == "c3":
        op = dict(type="VolumeVis", dataset="v", colormap=CMAP, opacitymap=[[0.0, 0.05], [0.2, 0.02], [0.21, 0.0], [1.0, 0.0]], data_range=None,
                  slices=[], isovalues=[], volume_render=True)
        return (dict(annotation="", lighting=scenes.parse_lighting({}), operators=[op]),
                scenes.parse_camera({"viewpoint": [0, 0, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}))
    op = dict(type="VolumeVis", dataset="v", colormap=CMAP, opacitymap=[[0, 1], [1, 1]], data_range=None, slices=[], isovalues=[0.6],
              volume_render=False)
    return (dict(annotation="", lighting=scenes.parse_lighting({"Sources": [[1, 1, -2, 0]], "shadows": True, "Ka": 0.4, "Kd": 0.6, "ao count": 0}),
                 operators=[op]),
            scenes.parse_camera({"viewpoint": [0, 0, -3], "viewdirection": [0, 0, 1], "viewup": [0, 1, 0], "aov": 30}))  # examples/noise_isovalue.state:25-30


PL_LINES, PL_VERTS = 20000, 41    # --workload pl: 20 000 helical poly-lines x 40 segments = 800 000 round Bezier segments


def pathlines_case():
    """(visualization, camera) of the PathLines workload: thin data-mapped tubes, primary + 1 shadow ray"""
    op = dict(type="PathLinesVis", dataset="lines", colormap=[[0.0, 0.0, 1.0, 0.0], [1.2, 1.0, 0.0, 1.0]], opacitymap=[[0, 1], [1, 1]], data_range=None,
              radius0=0.002, radius1=0.008, value0=0.0, value1=1.2)
    return (dict(annotation="", lighting=scenes.parse_lighting({"Sources": [[1, 2, -3, 0]], "shadows": True, "Ka": 0.4, "Kd": 0.6, "ao count": 0}),
                 operators=[op]),
            scenes.parse_camera({"viewpoint": [1.5, 1.0, -3.0], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 35}))


def pathlines_alg_bytes_per_ray(n_segments, hit_fraction):
    """the C5 model with curve leaves: ray I/O + L 80-byte nodes + one leaf of 3 segments (48-byte record + 64 bytes of control points each)"""
    import math
    L = max(1, math.ceil(math.log(max(n_segments, 8) / 3.0, 8)))
    return 76.0 + 28.0 * hit_fraction + 80.0 * L + 3 * (48.0 + 64.0), L


def run_cpu_baseline_pathlines(steps, warmup, lines_div=20, sample_div=4):
    """The oracle on a bounded sample of the PathLines workload: 1/lines_div of the lines, (1080p / sample_div^2) window."""
    from oracle import oracle
    vis, cam = pathlines_case()
    ds = scenes.helix_pathlines(PL_LINES // lines_div, PL_VERTS)
    parts = scenes.build_partitions(oracle, vis, {"lines": ds}, 1)
    w, h = W // sample_div, H // sample_div
    cores = os.cpu_count() or 1
    times, rays = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fb, st = oracle.render(parts, cam, vis["lighting"], w, h, EPS, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    t = float(np.mean(times))
    sample = "oracle port; same camera/lighting, %dx%d window, %d segments (1/%d of the lines); %d rays/frame" % (
        w, h, (PL_LINES // lines_div) * (PL_VERTS - 1), lines_div, rays)
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}, t * 1e3


def run_cpu_baseline_volume(which, steps, warmup, n=192, sample_div=4):
    """The oracle on a bounded sample of the volume workload: n^3 volume, (1080p / sample_div^2) window."""
    from oracle import oracle
    vis, cam = volume_case(which)
    vol = synth_volume(n, device="cpu")
    parts = scenes.build_partitions(oracle, vis, {"v": vol}, 1)
    w, h = W // sample_div, H // sample_div
    cores = os.cpu_count() or 1
    times, rays = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fb, st = oracle.render(parts, cam, vis["lighting"], w, h, EPS, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    t = float(np.mean(times))
    sample = "oracle port; same camera/lighting/transfer function, %dx%d window, %d^3 volume; %d rays/frame" % (w, h, n, rays)
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}, t * 1e3


def run_cpu_baseline(steps, warmup, sample_div=1, tess_div=4):
    """The oracle (CPU restatement of the reference's algorithm) on a bounded sample of the workload (only where oracle/_ref's
    Embree is absent)."""
    from oracle import oracle
    n_lat, n_lon = scenes.C5_FULL[0] // tess_div, scenes.C5_FULL[1] // tess_div
    parts, vis, ntri = cpu_sample_scene(oracle, n_lat, n_lon)
    cam = scenes.c5_camera()
    w, h = W // sample_div, H // sample_div
    cores = os.cpu_count() or 1
    times, rays = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fb, st = oracle.render(parts, cam, vis["lighting"], w, h, EPS, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    t = float(np.mean(times))
    sample = "oracle port; same camera/lighting, %dx%d window (1/%d of the 1080p pixels), %d triangles (1/%d tessellation of the 100M scene); %d rays/frame" % (
        w, h, sample_div * sample_div, ntri, tess_div * tess_div, rays)
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}, t * 1e3


_REF_SCENE = {}


def reference_scene(tess_div):
    """The C5 scene on the host for the CPU arm: one partition, its triangles committed to the reference's own Embree 3.6.1
    (oracle/_ref, built from /root/reference by oracle/embree.mk: binned-SAH BVH8 of Triangle4 leaves, as Galaxy builds it)."""
    if tess_div not in _REF_SCENE:
        from oracle import embree_scene
        n_lat, n_lon = scenes.C5_FULL[0] // tess_div, scenes.C5_FULL[1] // tess_div
        t0 = time.perf_counter()
        ds, _ = scenes.c5_partition_mesh(n_lat, n_lon, 1, 0)
        vis = scenes.c5_vis()
        parts = scenes.build_partitions(embree_scene.oracle_backend(), vis, {"mesh": ds}, 1)
        _REF_SCENE[tess_div] = (parts, vis, len(ds.indices), parts[0].embree.build_seconds, time.perf_counter() - t0)
    return _REF_SCENE[tess_div]


def run_cpu_reference(steps, warmup, tess_div=1, budget_s=120.0):
    """CPU arm, kind "reference": the FULL workload (all triangles, 1080p, primary + shadow + 8 AO rays) with the reference's own
    Embree doing what it does in Galaxy -- BVH build and rtcIntersect8 packet traversal (ospray Model.ih:54-70) -- on all host
    cores; the ISPC glue around it (TraceRays.ispc, lighting, classification, framebuffer) cannot be compiled here (no ispc) and is
    the oracle's scalar C++ restatement, threaded.  Also timed: Embree alone on the very rays of the frame (traversal_only), the
    ceiling of what the reference's ISPC-SIMD glue could reach on these cores.  Frames stop early once budget_s is spent."""
    from oracle import oracle, embree_scene
    parts, vis, ntri, build_s, setup_s = reference_scene(tess_div)
    cam = scenes.c5_camera()
    cores = os.cpu_count() or 1
    times, rays = [], 0
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        fb, st = oracle.render(parts, cam, vis["lighting"], W, H, EPS, nthreads=cores)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
        if time.perf_counter() - t_start > budget_s and times:
            break
    t = float(np.mean(times))
    # Embree alone on the frame's own ray lists: the primaries of the camera, then the AO + shadow rays their hits spawn
    L = oracle.resolve_lights(vis["lighting"], cam)
    rl, n = parts[0].generate_rays(cam, W, H)
    org, d, t0_, t1_ = embree_scene.raylist_columns(rl, n)
    es = parts[0].embree
    es.intersect(org, d, t0_, t1_, want=False)
    _, _, s_p = es.intersect(org, d, t0_, t1_, want=False)
    sec, ns, _ = parts[0].trace_raylist(L, rl, n, EPS)
    s_s = 0.0
    if ns:
        org2, d2, t02, t12 = embree_scene.raylist_columns(sec, ns)
        es.intersect(org2, d2, t02, t12, want=False)
        _, _, s_s = es.intersect(org2, d2, t02, t12, want=False)
    trav = (n + ns) / max(1e-9, s_p + s_s) / 1e6
    sample = ("reference's Embree 3.6.1 (BVH8/Triangle4 SAH build %.1f s, rtcIntersect8 packets, AVX2) for every nearest-hit query + the oracle's "
              "threaded C++ restatement of the ISPC glue; the full workload: %d triangles, %dx%d, %d rays/frame, %d frames timed; "
              "traversal_only = Embree alone on the frame's %d primary + %d secondary rays") % (build_s, ntri, W, H, rays, len(times), n, ns)
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
            "traversal_only": {"value": trav, "unit": "Mrays/s", "primary_mrays_s": n / max(1e-9, s_p) / 1e6,
                               "secondary_mrays_s": (ns / s_s / 1e6) if s_s > 0 else None},
            "embree_build_s": build_s, "scene_setup_s": setup_s, "frames_timed": len(times)}, t * 1e3


def parity_after_timing(gpu, ctx, world, rank, dist):
    """Driver-visible parity of exactly the code path that was timed (frames in flight, spatial partitions, peer exchange): after the
    timed region every rank builds ITS partition of a small version of the scene (1/16 tessellation per axis) and two cameras are
    rendered on two frame slots; rank 0 renders the same partitions with the CPU oracle and compares image and ray statistics."""
    import torch
    n_lat, n_lon, w, h = scenes.C5_FULL[0] // 16, scenes.C5_FULL[1] // 16, 480, 270
    vis = scenes.c5_vis()
    cams = [scenes.c5_camera(), scenes.parse_camera({"viewpoint": [-3, 1, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})]
    ds, _ = scenes.c5_partition_mesh(n_lat, n_lon, world, rank)
    part = scenes.build_partitions(gpu, vis, {"mesh": ds}, world, only_rank=rank, ctx=ctx)[0]
    keys = ["primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays"]
    for k, cam in enumerate(cams):
        gpu.render_submit([part], cam, vis["lighting"], w, h, EPS, k)
    got = []
    for k in range(len(cams)):
        st = gpu.render_wait([part], k)
        t = torch.tensor([st[key] for key in keys], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        got.append((dict(zip(keys, t.tolist())), part.download_rgba32f(w, h) if rank == 0 else None))
    if rank != 0:
        return None
    from oracle import oracle
    full, _ = scenes.c5_partition_mesh(n_lat, n_lon, 1, 0)
    o_parts = scenes.build_partitions(oracle, vis, {"mesh": full}, world)
    fracs, same = [], True
    for k, cam in enumerate(cams):
        fb_o, st_o = oracle.render(o_parts, cam, vis["lighting"], w, h, EPS)
        st_g, fb_g = got[k]
        fracs.append(float((np.abs(fb_g[..., :3] - fb_o[..., :3]).max(-1) <= 1.0 / 255).mean()))
        same = same and all(st_g[key] == st_o[key] for key in keys)
    return {"fraction": min(fracs), "stats_equal": bool(same), "frames": len(cams), "checked_against": "CPU oracle, same %d partitions" % world,
            "scene": "C5 at 1/16 tessellation per axis (%d triangles), %dx%d, 2 cameras on 2 frame slots" % (len(full.indices), w, h),
            "tolerance": "1/255 per channel on the float framebuffer; primary/shadow/AO/forwarded/terminated ray counts equal"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--tess-div", type=int, default=1, help="divide the C5 tessellation (debug only; 1 = the 100M-triangle workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c5", choices=["c5", "c3", "c4", "pl"],
                    help="c5 (default, the headline): 100M-triangle mesh; c3: volume DVR; c4: volume isosurface + shadow rays; "
                         "pl: PathLines (800 000 round Bezier segments), primary + shadow")
    ap.add_argument("--volume-n", type=int, default=1024, help="c3/c4: voxels per axis of the synthetic volume")
    ap.add_argument("--in-flight", type=int, default=None,
                    help="frames of the RenderingSet kept in flight (gxy_render_submit/wait); default 8 for the geometry workloads, 1 otherwise")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    volume = args.workload in ("c3", "c4")
    pathlines = args.workload == "pl"
    if pathlines:
        workload = "PathLines: %d helical poly-lines, %d round Bezier segments of radius 0.002-0.008, 1920x1080, primary + shadow (1 light), spatial partitions=%d" % (
            PL_LINES, PL_LINES * (PL_VERTS - 1), n_gpus)
        metric = "Mrays/s, 1080p PathLines primary+shadow"
    elif volume:
        workload = ("C3 noise volume %d^3 float32, 1920x1080, DVR only (no secondary rays), spatial partitions=%d" if args.workload == "c3" else
                    "C4 noise volume %d^3 float32, 1920x1080, isosurface 0.6 + shadow rays (1 light), spatial partitions=%d") % (args.volume_n, n_gpus)
        metric = "Mrays/s, 1080p volume march (%s)" % ("DVR" if args.workload == "c3" else "isosurface + shadow")
    else:
        workload = "C5 eightBalls-100M: %d triangles, 1920x1080, primary + shadow (1 light) + 8 AO rays, Triangles vis, spatial partitions=%d" % (
            8 * 2 * (scenes.C5_FULL[0] // args.tess_div) * (scenes.C5_FULL[1] // args.tess_div), n_gpus)
        metric = "Mrays/s, 1080p primary+shadow+AO"
    config = {"workload": workload, "width": W, "height": H, "partitions": n_gpus,
              "timing": ("a RenderingSet of frames in flight (frames_in_flight; gxy_render_submit/gxy_render_wait): every step renders and delivers its own "
                         "complete image, steps overlap on the device, ms_per_step = (last frame's end - first frame's start, CUDA events on the frames' "
                         "streams, max over ranks) / steps, so pipeline fill and drain are inside the timed region.  No explicit L2 flush: inputs larger "
                         "than L2 (a frame touches ~1 GB of a >= 4 GB scene, 8x the 126 MB L2).  e2e: the same with every image converted to RGBA8 and "
                         "copied to pinned host memory inside the timed region.  With --in-flight 1 (and for the volume / PathLines workloads): one "
                         "frame at a time, 256 MB L2 flush between frames, ms_per_step = mean per-frame device time")}

    if pathlines:
        config["timing"] = ("value: 256 MB L2 flush between frames; e2e: no explicit flush and the scene (about 90 MB of records and control points) "
                            "fits the 126 MB L2, image k downloads while frame k+1 renders")
    if args.impl == "reference":
        if rank != 0:
            return
        if pathlines:
            cb, ms = run_cpu_baseline_pathlines(max(1, args.steps), max(0, min(args.warmup, 1)))
        elif volume:
            cb, ms = run_cpu_baseline_volume(args.workload, max(1, args.steps), max(0, min(args.warmup, 1)))
        else:
            from oracle import embree_scene
            if embree_scene.available():
                cb, ms = run_cpu_reference(max(1, args.steps), max(0, min(args.warmup, 1)), args.tess_div)
            else:
                cb, ms = run_cpu_baseline(max(1, args.steps), max(0, min(args.warmup, 1)))
        line = {"metric": metric, "value": cb["value"], "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from galaxy_b200 import gpu
    if not torch.cuda.is_available() or gpu.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; galaxy_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == n_gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    nparts = world

    # ---- scene: this rank's spatial partition -------------------------------------------------
    ctx = gpu.Context(local_rank)
    if world > 1:
        uid = [gpu.comm_unique_id()] if rank == 0 else [None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    t0 = time.perf_counter()
    if pathlines:
        ds = scenes.helix_pathlines(PL_LINES, PL_VERTS)
        vis, cam = pathlines_case()
        dsets, n_tris_local = {"lines": ds}, 0
    elif volume:
        ds = synth_volume(args.volume_n)
        vis, cam = volume_case(args.workload)
        dsets, n_tris_local = {"v": ds}, 0
    else:
        n_lat, n_lon = scenes.C5_FULL[0] // args.tess_div, scenes.C5_FULL[1] // args.tess_div
        ds, _ = scenes.c5_partition_mesh(n_lat, n_lon, nparts, rank)
        vis, cam = scenes.c5_vis(), scenes.c5_camera()
        dsets, n_tris_local = {"mesh": ds}, len(ds.indices)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    part = scenes.build_partitions(gpu, vis, dsets, nparts, only_rank=rank, ctx=ctx)[0]
    t_commit = time.perf_counter() - t0
    info = part.build_info()
    del ds, dsets

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # frames in flight: a RenderingSet of `depth` frames is kept on the device at any time (the reference keeps every Rendering
    # of a set in flight, gxywriter.cpp:196-264).  Every step still delivers its own complete image; steps are counted as they
    # complete.  Geometry workloads only: the volume / PathLines schedules are synchronous per frame (depth 1).
    # measured (tools/flight_sweep.py, ms per frame at 1/2/4/8 GPUs): 4 in flight 1.10/0.85/0.63/0.57, 8 in flight 1.10/0.81/0.52/0.42
    # volumes across ranks: frames in flight only with the device-side-length schedule (GXY_VOLUME_FLIGHTS=1), else one frame at a time
    vol_flights = volume and world > 1 and os.environ.get("GXY_VOLUME_FLIGHTS", "1") != "0" and os.environ.get("GXY_PEER", "1") != "0"
    # (volumes: 4 in flight -- enough to keep the bricks along a ray busy; more concurrent marches only fight for the caches)
    depth = args.in_flight if args.in_flight is not None else (8 if not (volume or pathlines) else 4 if vol_flights else 1)
    depth = max(1, min(depth, gpu.max_slots(), max(1, args.steps)))
    use_flush = depth == 1   # depth 1: 256 MB L2 flush between frames; depth > 1: the frames' inputs are >> L2 (config.timing)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def submit(slot):
        gpu.render_submit([part], cam, vis["lighting"], W, H, EPS, slot)

    def run_pipeline(n_steps, on_frame=None):
        """n_steps frames, `depth` in flight; returns the per-frame stats in completion order"""
        out = []
        for k in range(min(depth, n_steps)):
            if use_flush:
                flush.zero_()
                torch.cuda.synchronize()
            submit(k % depth)
        for k in range(n_steps):
            st = gpu.render_wait([part], k % depth)
            out.append(st)
            if on_frame is not None:
                on_frame(k)
            if k + depth < n_steps:
                if use_flush:
                    flush.zero_()
                    torch.cuda.synchronize()
                submit(k % depth)
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    run_pipeline(max(3, args.warmup) + depth - 1)
    # ---- timed region: K frames, device-timed: CUDA events of the library on the frames' own streams, stamps relative to ctx.mark()
    barrier()
    ctx.mark()
    barrier()
    sampler.mark_begin()
    t_wall0 = time.perf_counter()
    frames = run_pipeline(args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.mark_end()
    clocks = sampler.stop()
    if use_flush:   # frames run one after the other with a flush in between: the sum of their device times
        ms_local = float(np.sum([f["device_ms"] for f in frames]))
    else:           # frames overlap: first start to last end on the device
        ms_local = max(f["t_end_ms"] for f in frames) - min(f["t_begin_ms"] for f in frames)
    trace_ms = float(np.sum([f["trace_ms"] for f in frames]))
    launches = int(np.sum([f["kernel_launches"] for f in frames]))
    traced = int(np.sum([f["traced_rays"] for f in frames]))
    dequeued = int(np.sum([f["dequeued_rays"] for f in frames]))
    samples = int(np.sum([f["volume_samples"] for f in frames]))
    latency_ms = float(np.mean([f["device_ms"] for f in frames]))
    st = frames[-1]
    rays_local = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    hit_local = st["shadow_rays"]  # one shadow ray per surface-hit primary (1 light)

    # ---- e2e: through the public calls with host buffers: camera/lights H2D, RGBA8 image D2H into pinned memory ------
    # Every frame's image is delivered to the host inside the timed region; the transfer of frame k (copy stream, one of
    # two pinned buffers) overlaps the frames behind it, as a viewer or an image writer thread would run it.
    imgs = [gpu.pinned_array((H, W, 4), np.uint8) for _ in range(2)] if rank == 0 else None

    def deliver(k):
        if rank == 0:
            part.download_wait()  # image k-1 (it travelled while frame k rendered)
            part.download_rgba8_async(imgs[k & 1])

    run_pipeline(2, deliver)
    if rank == 0:
        part.download_wait()
    barrier()
    t0 = time.perf_counter()
    run_pipeline(args.steps, deliver)   # no explicit L2 flush here (config.timing)
    if rank == 0:
        part.download_wait()
    barrier()
    t_e2e = time.perf_counter() - t0

    parity = None
    if not (volume or pathlines):
        parity = parity_after_timing(gpu, ctx, world, rank, dist)

    if world > 1:
        t = torch.tensor([ms_local, t_e2e, trace_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max, t_e2e, trace_ms_max = t.tolist()
        c = torch.tensor([rays_local, launches, traced, hit_local, st["primary_rays"], dequeued, latency_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays_total, launches_total, traced_total, hits_total, prim_total, dequeued_total, latency_sum = c.tolist()
        latency_ms = latency_sum / world
    else:
        ms_max, trace_ms_max = ms_local, trace_ms
        rays_total, launches_total, traced_total, hits_total, prim_total, dequeued_total = rays_local, launches, traced, hit_local, st["primary_rays"], dequeued

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = ms_max / args.steps
    value = rays_total / (ms_per_step * 1e-3) / 1e6
    e2e_value = rays_total / (t_e2e / args.steps) / 1e6
    peaks, peak_kind = measured_peaks()
    hbm = float(peaks["hbm_gbs"])
    hfrac = hits_total / max(1.0, prim_total)
    b_alg, levels = c5_alg_bytes_per_ray(n_tris_local, hfrac)
    # dominant kernels = the persistent trace launches.  Units: the rays those launches DEQUEUE (primaries the generation kernel
    # finishes itself -- they can reach no primitive -- are not counted), times the algorithmic bytes per ray of SURVEY 8(d).
    # Time: depth 1 -> the summed CUDA-event durations of the trace launches; frames in flight -> the launches of different
    # frames overlap, so the time they occupy the device is the timed region itself (first start to last end on this rank,
    # which also contains the ~3 % of generation/shading kernels: a lower bound for the trace kernels alone).
    kernel_ms = trace_ms if use_flush else ms_local
    achieved = dequeued * b_alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    traffic = NCU_TRAFFIC_BYTES_PER_FRAME if n_gpus == 1 and args.tess_div == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                "kernel": "gxy::primary_trace_kernel + gxy::fused_secondary_kernel (+ gxy::inbox_trace_kernel across ranks): the persistent trace launches",
                "traffic_note": "dram__bytes_read+write of the trace launches of one frame, ncu --set full, profiles/r01_e_trace_kernels_full.txt",
                "dram_util": (traffic / (ms_per_step * 1e-3) / 1e9 / hbm) if traffic else None,
                "peak_kind": peak_kind + " (burst copy figure)", "alg_bytes_per_ray": b_alg, "bvh_levels_model": levels,
                "units_per_frame": dequeued / args.steps, "units": "rays dequeued by the trace launches of this rank (culled primaries excluded)",
                "kernel_time": "sum of trace-launch CUDA-event durations" if use_flush else "timed region of this rank (launches of frames in flight overlap)",
                "trace_share_of_step": trace_ms_max / max(1e-9, ms_max) if use_flush else None}
    if pathlines:
        b_alg, levels = pathlines_alg_bytes_per_ray(info["n_prims"], hfrac)
        achieved = traced * b_alg / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                    "kernel": "gxy::trace_kernel<0,true,false,true> (per-lane wide-BVH traversal with the round-Bezier curve test)",
                    "peak_kind": peak_kind + " (burst copy figure)", "alg_bytes_per_ray": b_alg, "bvh_levels_model": levels,
                    "trace_share_of_step": trace_ms_max / max(1e-9, ms_max)}
    if volume:
        # B_alg = min(16 B x samples, 4 B x voxels in the frustum): 16 algorithmic bytes per trilinear sample (4 new
        # float voxels per step) when rays are >= 1 voxel apart, but never more than reading every voxel of this rank's brick once
        vox_rank = float(args.volume_n) ** 3 / max(1, nparts)
        spf = samples / args.steps
        alg_frame = min(16.0 * spf, 4.0 * vox_rank)
        vol_ms = trace_ms if use_flush else ms_local   # frames in flight: the march launches of different frames overlap (see the C5 note)
        achieved = alg_frame * args.steps / (vol_ms * 1e-3) / 1e9 if vol_ms > 0 else 0.0
        vtraffic = NCU_VOLUME_TRAFFIC.get(args.workload) if (args.volume_n == 1024 and n_gpus == 1) else None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                    "traffic": vtraffic,
                    "kernel": "gxy::trace_kernel<1,false,true> (volume march: trilinear sample + transfer function + compositing / isosurface search)",
                    "traffic_note": "dram__bytes_read+write per frame at 1024^3, ncu --set full, profiles/r01_e_volume_full.txt",
                    "dram_util": (vtraffic / (ms_per_step * 1e-3) / 1e9 / hbm) if vtraffic else None,
                    "peak_kind": peak_kind + " (burst copy figure)", "alg_bytes_per_frame": alg_frame,
                    "alg_bytes_rule": "min(16 B x samples, 4 B x voxels of this rank's brick)", "samples_per_frame": spf,
                    "gsamples_per_s": samples / (vol_ms * 1e-3) / 1e9 if vol_ms > 0 else 0.0,
                    "trace_share_of_step": trace_ms_max / max(1e-9, ms_max) if use_flush else None}
    line = {"metric": metric, "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 40 + 276, "d2h_bytes_per_step": W * H * 4},
            "gpu_launches": int(launches_total), "roofline": roofline, "clocks": clocks,
            "rays_per_frame": int(rays_total), "traced_rays_per_frame": int(traced_total / args.steps),
            "dequeued_rays_per_frame": int(dequeued_total / args.steps), "wall_ms_per_step": t_wall / args.steps * 1e3,
            "frames_in_flight": depth, "frame_latency_ms": latency_ms,
            "scene": {"triangles_this_rank": n_tris_local, "bvh_nodes": info["n_nodes"], "bvh_build_ms": info["build_ms"], "bvh_build_alloc_host_ms": info.get("alloc_host_ms"), "mesh_gen_s": t_gen,
                      "commit_s": t_commit}}
    if parity is not None:
        line["parity"] = parity
    if n_gpus == 1 and not args.no_cpu_baseline:
        if pathlines:
            cb, _ = run_cpu_baseline_pathlines(2, 0)
        elif volume:
            cb, _ = run_cpu_baseline_volume(args.workload, 2, 0)
        else:
            from oracle import embree_scene
            # bounded sample: 3 frames of the full workload (about 10-30 s of CPU work after the scene is built)
            cb, _ = run_cpu_reference(3, 0, args.tess_div, budget_s=40.0) if embree_scene.available() else run_cpu_baseline(3, 0)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
