#!/usr/bin/env python
"""Per-phase host timing of the multi-process frame loop (GXY_PROFILE=1) on the headline scene.
torchrun ... tools/mp_profile.py [tess_div]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
tess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(local)
ctx = gpu.Context(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [gpu.comm_unique_id()] if rank == 0 else [None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
ds, _ = scenes.c5_partition_mesh(scenes.C5_FULL[0] // tess, scenes.C5_FULL[1] // tess, world, rank)
vis, cam = scenes.c5_vis(), scenes.c5_camera()
part = scenes.build_partitions(gpu, vis, {"mesh": ds}, world, only_rank=rank, ctx=ctx)[0]
del ds
for it in range(6):
    if it == 4:
        os.environ["GXY_PROFILE"] = "1"
    st = gpu.render_device([part], cam, vis["lighting"], 1920, 1080, 0.001)
    if it >= 4 and rank == 0:
        print("frame", it, "device_ms %.3f trace_ms %.3f" % (st["device_ms"], st["trace_ms"]), flush=True)
if world > 1:
    dist.destroy_process_group()
