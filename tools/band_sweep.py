#!/usr/bin/env python
"""Sweep of the band pipeline of the single-GPU fused frame path (GXY_BANDS x GXY_BAND_STREAMS) on the headline scene.
usage: python tools/band_sweep.py "bands,streams;bands,streams;..." [frames]   (GPU box only)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

combos = [tuple(int(x) for x in c.split(",")) for c in (sys.argv[1] if len(sys.argv) > 1 else "1,1;2,2").split(";")]
combos = [c if len(c) >= 3 else c + (8,) for c in combos]  # bands, streams, CTAs per SM per launch [, fetch threshold primary, secondary]
combos = [c if len(c) == 5 else c[:3] + (12, 12) for c in combos]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = gpu.Context(0)
ds, _ = scenes.c5_partition_mesh(scenes.C5_FULL[0], scenes.C5_FULL[1], 1, 0)
vis, cam = scenes.c5_vis(), scenes.c5_camera()
part = scenes.build_partitions(gpu, vis, {"mesh": ds}, 1, only_rank=0, ctx=ctx)[0]
del ds
ref = None
for bands, streams, bps, fp, fs in combos:
    os.environ["GXY_BANDS"], os.environ["GXY_BAND_STREAMS"], os.environ["GXY_FUSED_BLOCKS_PER_SM"] = str(bands), str(streams), str(bps)
    os.environ["GXY_FETCH_P"], os.environ["GXY_FETCH_S"] = str(fp), str(fs)
    ms = []
    for it in range(frames + 3):
        st = gpu.render_device([part], cam, vis["lighting"], 1920, 1080, 0.001)
        if it >= 3:
            ms.append(st["device_ms"])
    img = part.download_rgba8(1920, 1080)
    ref = img if ref is None else ref
    rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    print(json.dumps({"bands": bands, "streams": streams, "ctas_per_sm": bps, "fetch": [fp, fs], "ms_median": round(float(np.median(ms)), 4), "ms_min": round(float(np.min(ms)), 4),
                      "Mrays/s": round(rays / np.median(ms) / 1e3, 1), "rays": rays, "img_maxdiff_vs_first": int(np.abs(img.astype(int) - ref.astype(int)).max())}),
          flush=True)
