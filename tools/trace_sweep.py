#!/usr/bin/env python
"""Tuning sweep of the trace kernel variants on the headline (C5) scene, one process, one scene build.
usage: python tools/trace_sweep.py [tess_div] [frames] [variant ...]   (GPU box only)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

tess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
variants = sys.argv[3:] or ["old", "f8p1b8", "c8b8", "c8b8L2", "c4b8", "c12b8", "c8b6", "c8b10"]
W, H = 1920, 1080
ctx = gpu.Context(0)
t0 = time.time()
ds, _ = scenes.c5_partition_mesh(scenes.C5_FULL[0] // tess, scenes.C5_FULL[1] // tess, 1, 0)
vis, cam = scenes.c5_vis(), scenes.c5_camera()
part = scenes.build_partitions(gpu, vis, {"mesh": ds}, 1, only_rank=0, ctx=ctx)[0]
print("scene", len(ds.indices), "tris", part.build_info(), "setup %.1fs" % (time.time() - t0), flush=True)
del ds
ref_img = None
fetch = [(int(a), int(b)) for a, b in (x.split(",") for x in os.environ.get("GXY_FETCH_SWEEP", "").split(";") if x)] or [None]
for v in [(v, f) for v in variants for f in fetch]:
    v, f = v
    if f:
        os.environ["GXY_FETCH_P"], os.environ["GXY_FETCH_S"] = str(f[0]), str(f[1])
    if v == "old":
        os.environ["GXY_TRACE_PERSISTENT"] = "0"
    else:
        os.environ["GXY_TRACE_PERSISTENT"] = "1"
        os.environ["GXY_TRACE_VARIANT"] = v
    ms, tr = [], []
    for it in range(frames + 2):
        st = gpu.render_device([part], cam, vis["lighting"], W, H, 0.001)
        if it >= 2:
            ms.append(st["device_ms"])
            tr.append(st["trace_ms"])
    img = part.download_rgba8(W, H)
    if ref_img is None:
        ref_img = img
    rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    print(json.dumps({"variant": v, "fetch": f, "ms": round(float(np.median(ms)), 4), "trace_ms": round(float(np.median(tr)), 4), "min_ms": round(float(np.min(ms)), 4),
                      "Mrays/s": round(rays / np.median(ms) / 1e3, 1), "rays": rays, "nodes/ray": round(st["nodes_visited"] / max(1, st["traced_rays"]), 2),
                      "prims/ray": round(st["prims_tested"] / max(1, st["traced_rays"]), 2), "img_equal_first": bool(np.array_equal(img, ref_img)),
                      "img_maxdiff": int(np.abs(img.astype(int) - ref_img.astype(int)).max())}), flush=True)
