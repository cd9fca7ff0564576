#!/usr/bin/env python
"""A few single C5 frames on one GPU for ncu (one band, one frame at a time): python tools/prof_frame.py [tess_div] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GXY_BANDS", "1")
from galaxy_b200 import gpu, scenes  # noqa: E402

tess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = gpu.Context(0)
ds, _ = scenes.c5_partition_mesh(scenes.C5_FULL[0] // tess, scenes.C5_FULL[1] // tess, 1, 0)
vis, cam = scenes.c5_vis(), scenes.c5_camera()
part = scenes.build_partitions(gpu, vis, {"mesh": ds}, 1, ctx=ctx)[0]
del ds
for k in range(frames):
    st = gpu.render_device([part], cam, vis["lighting"], 1920, 1080, 0.001)
rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
print({"tess": tess, "device_ms": st["device_ms"], "rays": rays, "dequeued": st["dequeued_rays"], "nodes_visited": st["nodes_visited"],
       "prims_tested": st["prims_tested"], "nodes_per_dequeued_ray": st["nodes_visited"] / max(1, st["dequeued_rays"]),
       "prims_per_dequeued_ray": st["prims_tested"] / max(1, st["dequeued_rays"])})
