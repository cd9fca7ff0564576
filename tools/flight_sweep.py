#!/usr/bin/env python
"""Throughput of the C5 frame with D frames in flight (gxy_render_submit / gxy_render_wait), the scene built once.
  python tools/flight_sweep.py [tess_div] [steps]                  (one GPU)
  torchrun --nproc-per-node N ... tools/flight_sweep.py [tess_div] (one process per GPU, peer arenas)
Environment sweeps: FLIGHT_DEPTHS="1,2,4,8"  FLIGHT_ENVS="GXY_BANDS=4;GXY_BANDS=2;GXY_BANDS=1"
FLIGHT_TIMELINE=<file>: append one JSON line per (rank, frame) of the last repetition: begin / end of the frame on that rank's device. """
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from galaxy_b200 import gpu, scenes  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
tess = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
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
W, H = 1920, 1080


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


HOST = {"submit_s": 0.0, "n": 0}


def submit(slot):
    t0 = time.perf_counter()
    gpu.render_submit([part], cam, vis["lighting"], W, H, 0.001, slot)
    HOST["submit_s"] += time.perf_counter() - t0
    HOST["n"] += 1


def pipeline(n, depth):
    out = []
    for k in range(min(depth, n)):
        submit(k % depth)
    for k in range(n):
        out.append(gpu.render_wait([part], k % depth))
        if k + depth < n:
            submit(k % depth)
    return out


depths = [int(x) for x in os.environ.get("FLIGHT_DEPTHS", "1,2,3,4,6,8").split(",")]
envs = [e for e in os.environ.get("FLIGHT_ENVS", "").split(";")] if os.environ.get("FLIGHT_ENVS") else [""]
for env in envs:
    sets = [kv.split("=") for kv in env.split(",") if "=" in kv]
    for k, v in sets:
        os.environ[k] = v
    for depth in depths:
        pipeline(depth + 3, depth)
        HOST["submit_s"], HOST["n"] = 0.0, 0
        res = []
        for rep in range(3):
            barrier()
            ctx.mark()
            barrier()
            t0 = time.perf_counter()
            fr = pipeline(steps, depth)
            barrier()
            wall = time.perf_counter() - t0
            span = max(f["t_end_ms"] for f in fr) - min(f["t_begin_ms"] for f in fr)
            t = torch.tensor([span, wall * 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rays = fr[-1]["primary_rays"] + fr[-1]["shadow_rays"] + fr[-1]["ao_rays"]
            busy = float(np.mean([f["device_ms"] for f in fr]))
            c = torch.tensor([rays, busy, fr[-1]["dequeued_rays"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(c)
            pr = torch.zeros(world, dtype=torch.float64, device="cuda")
            pr[rank] = fr[-1]["dequeued_rays"]
            if world > 1:
                dist.all_reduce(pr)
            res.append((t[0].item() / steps, t[1].item() / steps, c[1].item() / world, c[0].item(), c[2].item(), [int(x) for x in pr.tolist()]))
        tl_path = os.environ.get("FLIGHT_TIMELINE")
        if tl_path:
            # per-rank timeline of the last repetition: when each frame's first kernel started and its last one ended on this rank's
            # device, relative to the mark all ranks set behind a barrier (so the origins agree to within the barrier's skew)
            mine = [dict(frame=k, slot=k % depth, t_begin_ms=round(f["t_begin_ms"], 4), t_end_ms=round(f["t_end_ms"], 4),
                         trace_ms=round(f["trace_ms"], 4), dequeued_rays=int(f["dequeued_rays"]), waves=int(f["waves"])) for k, f in enumerate(fr)]
            allr = [None] * world
            if world > 1:
                dist.all_gather_object(allr, mine)
            else:
                allr = [mine]
            if rank == 0:
                with open(tl_path, "a") as fh:
                    for r, rows in enumerate(allr):
                        for row in rows:
                            fh.write(json.dumps(dict(world=world, depth=depth, env=env, rank=r, **row)) + "\n")
        if rank == 0:
            best = min(res)
            print(json.dumps({"world": world, "env": env, "depth": depth, "ms_per_frame_device": round(best[0], 4), "ms_per_frame_wall": round(best[1], 4),
                              "frame_latency_ms": round(best[2], 4), "Mrays/s": round(best[3] / best[0] / 1e3, 1), "rays": int(best[3]),
                              "dequeued": int(best[4]), "dequeued_per_rank": best[5], "host_submit_ms_per_frame": round(HOST["submit_s"] / max(1, HOST["n"]) * 1e3, 4), "all": [round(r[0], 4) for r in res]}), flush=True)
    for k, v in sets:
        os.environ.pop(k, None)
if world > 1:
    dist.destroy_process_group()
