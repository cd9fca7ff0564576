#!/usr/bin/env python
"""The Sampler across processes (one process per GPU, gxy_sample with a communicator): rank r owns partition r of the volume, rays
that leave a partition travel to the neighbour's process over NCCL; rank 0 compares every partition's SET of samples and the summed
ray statistics with the CPU oracle sampling the same partitions in one process.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mp_sampler_parity.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

CAM = dict(eye=[2.0, 1.5, -3.0], dir=[-2.0, -1.5, 3.0], up=[0.0, 1.0, 0.0], aov=35.0)
KEYS = ["primary_rays", "traced_rays", "forwarded_rays"]


def sampler_vis(kind, param):
    key = "tolerance" if kind == "GradientSampler" else "isovalue"
    return dict(annotation="", lighting=scenes.parse_lighting(None), operators=[scenes.parse_operator({"type": kind, "dataset": "v", key: param})])


def sorted_rows(a):
    a = np.ascontiguousarray(a, np.float32)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gpu.Context(local)
    uid = [gpu.comm_unique_id()] if rank == 0 else [None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    vol = scenes.radial_volume("eightBalls", 64)
    w, h = 160, 120
    ok = True
    for kind, param in (("IsoSampler", 0.25), ("GradientSampler", 0.9)):
        vis = sampler_vis(kind, param)
        part = scenes.build_partitions(gpu, vis, {"v": vol}, world, only_rank=rank, ctx=ctx)[0]
        for loop in ("1", "0"):
            os.environ["GXY_SAMPLER_LOOP"] = loop
            samples, st = gpu.sample([part], CAM, w, h)
            t = torch.tensor([st[k] for k in KEYS], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            mine = samples[0]
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            if rank == 0:
                from oracle import oracle
                o_parts = scenes.build_partitions(oracle, vis, {"v": vol}, world)
                so, st_o = oracle.sample(o_parts, CAM, w, h)
                sets_equal = all(sorted_rows(gathered[r]).shape == sorted_rows(so[r]).shape
                                 and np.array_equal(sorted_rows(gathered[r]).view(np.uint32), sorted_rows(so[r]).view(np.uint32)) for r in range(world))
                stats = dict(zip(KEYS, t.tolist()))
                # the one-launch mode counts the same passes; both must equal the oracle's totals
                same = all(stats[k] == st_o[k] for k in KEYS)
                print(json.dumps({"mode": "sampler %s loop=%s" % (kind, loop), "world": world, "sample_sets_equal": bool(sets_equal), "stats_equal": same,
                                  "samples_per_partition": [int(len(g)) for g in gathered], "gpu": stats, "oracle": {k: st_o[k] for k in KEYS},
                                  "waves": st["waves"], "device_ms": st["device_ms"]}), flush=True)
                ok = ok and sets_equal and same
    os.environ.pop("GXY_SAMPLER_LOOP", None)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
