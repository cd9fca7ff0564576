#!/usr/bin/env python
"""Volume configs of BASELINE.json on one GPU (not the bench.py headline): C3 = N^3 noise-like volume, DVR
only; C4 = isosurface + shadow rays.  The volume is synthesised on the GPU with torch (eightBalls distance
field + a sine lattice: closed form, seeded by nothing) -- it is a throughput workload, parity is covered
by the tests on the reference's own volumes.
usage: python tools/volume_bench.py N [frames] [c3|c4|both]"""
import json
import sys
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
which = sys.argv[3] if len(sys.argv) > 3 else "both"
W, H = 1920, 1080
t0 = time.time()
c = torch.linspace(-1.0, 1.0, N, device="cuda", dtype=torch.float32)
data = np.empty((N, N, N), np.float32)
Y, X = torch.meshgrid(c, c, indexing="ij")
for k0 in range(0, N, 64):
    Z = c[k0:k0 + 64].view(-1, 1, 1)
    eb = torch.sqrt((X.abs() - .5) ** 2 + (Y.abs() - .5) ** 2 + (Z.abs() - .5) ** 2)
    v = eb + 0.15 * (torch.sin(37.0 * X) * torch.sin(41.0 * Y) * torch.sin(43.0 * Z) * 0.5 + 0.5)
    data[k0:k0 + 64] = v.cpu().numpy()
sp = 2.0 / (N - 1)
vol = scenes.VolumeDataset([-1.0, -1.0, -1.0], (N, N, N), [sp, sp, sp], data)
print("volume %d^3 synthesised in %.1fs" % (N, time.time() - t0), flush=True)
cmap = [[0.0, 1.0, 0.5, 0.5], [0.25, 0.5, 1.0, 0.5], [0.5, 0.5, 0.5, 1.0], [0.75, 1.0, 1.0, 0.5], [1.0, 1.0, 0.5, 1.0]]
omap = [[0.0, 0.05], [0.2, 0.02], [0.21, 0.0], [1.0, 0.0]]
cases = {
    "c3": (dict(type="VolumeVis", dataset="v", colormap=cmap, opacitymap=omap, data_range=None, slices=[], isovalues=[], volume_render=True),
           scenes.parse_lighting({}), scenes.parse_camera({"viewpoint": [0, 0, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})),
    "c4": (dict(type="VolumeVis", dataset="v", colormap=cmap, opacitymap=[[0, 1], [1, 1]], data_range=None, slices=[], isovalues=[0.35], volume_render=False),
           scenes.parse_lighting({"Sources": [[1, 1, -2, 0]], "shadows": True, "Ka": 0.4, "Kd": 0.6, "ao count": 0}),
           scenes.parse_camera({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})),
}
for name, (op, lighting, cam) in cases.items():
    if which not in ("both", name):
        continue
    vis = dict(annotation="", lighting=lighting, operators=[op])
    part = scenes.build_partitions(gpu, vis, {"v": vol}, 1)[0]
    ms, st = [], None
    for it in range(frames + 2):
        st = gpu.render_device([part], cam, lighting, W, H, 0.001)
        if it >= 2:
            ms.append(st["device_ms"])
    rays = st["primary_rays"] + st["shadow_rays"] + st["ao_rays"]
    m = float(np.median(ms))
    print(json.dumps({"case": name, "N": N, "ms": round(m, 3), "trace_ms": round(st["trace_ms"], 3), "rays": rays, "Mrays/s": round(rays / m / 1e3, 1),
                      "samples": st["volume_samples"], "staged": round(st["staged_samples"] / max(1, st["volume_samples"]), 3), "Gsamples/s": round(st["volume_samples"] / m / 1e6, 2),
                      "alg_GB/s_16B_per_sample": round(st["volume_samples"] * 16 / m / 1e6, 1), "waves": st["waves"]}), flush=True)
    del part
