"""Throughput of the two "next" rows built this round (SURVEY 8f2 PathLines, 8f3 Sampler) on one GPU, in a few seconds:
   python tools/f_rows_bench.py [n_lines] [volume_n]          (GPU, through the C ABI)
   python tools/f_rows_bench.py [n_lines] [volume_n] --cpu    (the oracle port on the host cores, 480x270: a reported baseline)
PathLines: n_lines helical poly-lines of 40 segments, 1920x1080, primary + 1 shadow ray (list path, trace_kernel<0,true,false,CURVES>).
Sampler: IsoSampler on the radial eightBalls volume, 1920x1080 camera rays (sampler_trace_kernel + classify + re-queue)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from galaxy_b200 import scenes  # noqa: E402


def lines_dataset(n_lines, seed=3):
    rng = np.random.default_rng(seed)
    pts, lines, k = [], [], 0
    for _ in range(n_lines):
        t = np.linspace(0, 4.0, 41)
        c = rng.uniform(-.6, .6, 3)
        r = rng.uniform(0.1, 0.3)
        p = np.stack([c[0] + r * np.cos(t), c[1] + r * np.sin(t), c[2] + 0.1 * t - 0.2], 1)
        pts.append(p); lines.append(list(range(k, k + 41))); k += 41
    pts = np.concatenate(pts).astype(np.float32)
    return scenes.PathLinesDataset(pts, np.linalg.norm(pts, axis=1), lines)


def cpu_leg(n_lines, vol_n, cam, vis, svis):
    """the same two workloads on the oracle (all host threads) at a quarter of the resolution per axis"""
    import os
    from oracle import oracle
    w, h = 480, 270
    out = {"cores": os.cpu_count(), "kind": "port", "sample": "%dx%d of the 1920x1080 frame" % (w, h)}
    parts = scenes.build_partitions(oracle, vis, {"lines": lines_dataset(n_lines)}, 1)
    t0 = time.time()
    _, st = oracle.render(parts, cam, vis["lighting"], w, h, 0.001)
    dt = time.time() - t0
    rays = st["primary_rays"] + st["shadow_rays"]
    out["pathlines"] = dict(rays=rays, seconds=dt, mrays_per_s=rays / dt / 1e6)
    sp = scenes.build_partitions(oracle, svis, {"v": scenes.radial_volume("eightBalls", vol_n)}, 1)
    t0 = time.time()
    samp, st = oracle.sample(sp, cam, w, h)
    dt = time.time() - t0
    out["sampler"] = dict(traced_rays=st["traced_rays"], samples=int(len(samp[0])), seconds=dt, mrays_per_s=st["traced_rays"] / dt / 1e6)
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_lines = int(args[0]) if len(args) > 0 else 1000
    vol_n = int(args[1]) if len(args) > 1 else 128
    out = {}
    cam = dict(eye=[1.5, 1.0, -3.0], dir=[-1.5, -1.0, 3.0], up=[0.0, 1.0, 0.0], aov=35.0)
    vis = dict(annotation="", lighting=dict(lights=[[1.0, 2.0, -3.0]], types=[2], n_ao=0, ao_radius=0.5, shadows=True, Ka=0.4, Kd=0.6),
               operators=[dict(type="PathLinesVis", dataset="lines", colormap=[[0.0, 0.0, 1.0, 0.0], [1.2, 1.0, 0.0, 1.0]], opacitymap=[[0, 1], [1, 1]],
                               data_range=None, radius0=0.004, radius1=0.012, value0=0.0, value1=1.2)])
    svis = dict(annotation="", lighting=scenes.parse_lighting(None), operators=[scenes.parse_operator({"type": "IsoSampler", "dataset": "v", "isovalue": 0.25})])
    if "--cpu" in sys.argv:
        print(json.dumps(cpu_leg(n_lines, vol_n, cam, vis, svis)))
        return
    from galaxy_b200 import gpu
    t0 = time.time()
    parts = scenes.build_partitions(gpu, vis, {"lines": lines_dataset(n_lines)}, 1)
    ms = []
    for it in range(6):
        st = gpu.render_device(parts, cam, vis["lighting"], 1920, 1080, 0.001)
        ms.append(st["device_ms"])
    rays = st["primary_rays"] + st["shadow_rays"]
    best = float(np.median(ms[2:]))
    out["pathlines"] = dict(segments=n_lines * 40, build=parts[0].build_info(), rays_per_frame=rays, ms_per_frame=best, trace_ms=st["trace_ms"],
                            mrays_per_s=rays / best / 1e3, hit_fraction=st["shadow_rays"] / max(1, st["primary_rays"]), wall_s=time.time() - t0)
    t0 = time.time()
    vol = scenes.radial_volume("eightBalls", vol_n)
    sp = scenes.build_partitions(gpu, svis, {"v": vol}, 1)
    import os
    for mode, key in (("0", "sampler"), ("1", "sampler_loop_mode")):   # GXY_SAMPLER_LOOP: one launch per crossing / per visit of a partition
        os.environ["GXY_SAMPLER_LOOP"] = mode
        ms = []
        for it in range(4):
            samp, st = gpu.sample(sp, cam, 1920, 1080)
            ms.append(st["device_ms"])
        best = float(np.median(ms[1:]))
        out[key] = dict(volume=vol_n, samples=int(len(samp[0])), traced_rays=st["traced_rays"], waves=st["waves"], ms_per_frame=best,
                        mrays_per_s=st["traced_rays"] / best / 1e3, wall_s=time.time() - t0)
    os.environ.pop("GXY_SAMPLER_LOOP", None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
