#!/usr/bin/env python
"""Host phase timer (GXY_PROFILE=1) of the volume frame loop across ranks: torchrun ... tools/vol_profile.py [c3|c4] [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from galaxy_b200 import gpu, scenes  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.cuda.set_device(local)
ctx = gpu.Context(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [gpu.comm_unique_id()] if rank == 0 else [None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
vis, cam = bench.volume_case(which)
part = scenes.build_partitions(gpu, vis, {"v": bench.synth_volume(n)}, world, only_rank=rank, ctx=ctx)[0]
for it in range(6):
    if it == 4:
        os.environ["GXY_PROFILE"] = "1"
    st = gpu.render_device([part], cam, vis["lighting"], 1920, 1080, 0.001)
    if it >= 4 and rank == 0:
        print("frame", it, {k: st[k] for k in ("device_ms", "trace_ms", "waves", "kernel_launches", "forwarded_rays", "traced_rays")}, flush=True)
if world > 1:
    dist.destroy_process_group()
