#!/usr/bin/env python
"""BVH build of the full C5 scene, committed several times in one process: build_ms (CUDA events around the whole build) beside the host
time the driver spent inside the build's cudaMalloc / cudaFree calls (gxy_vis_build_times), and the wall clock of the commit.
  python tools/bvh_build_probe.py [tess_div] [repeats]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401  (as in bench.py: the build shares the process with torch's context)

from galaxy_b200 import gpu, scenes  # noqa: E402

tess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = gpu.Context(0)
ds, _ = scenes.c5_partition_mesh(scenes.C5_FULL[0] // tess, scenes.C5_FULL[1] // tess, 1, 0)
vis = scenes.c5_vis()
t0 = time.perf_counter()
part = scenes.build_partitions(gpu, vis, {"mesh": ds}, 1, only_rank=0, ctx=ctx)[0]
first_wall = time.perf_counter() - t0
info = part.build_info()
print(json.dumps({"commit": 0, "wall_s_incl_upload": round(first_wall, 3), **info}), flush=True)
for k in range(1, reps):
    t0 = time.perf_counter()
    part.commit()
    wall = time.perf_counter() - t0
    print(json.dumps({"commit": k, "wall_s": round(wall, 3), **part.build_info()}), flush=True)
