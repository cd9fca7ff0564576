#!/usr/bin/env python
"""Multi-process parity check of the spatially partitioned frame path (one process per GPU, NCCL ray
exchange + framebuffer reduce): run under torchrun, rank r owns partition r; rank 0 compares the image
and the ray statistics with the CPU oracle rendering the same partitions in one process.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mp_parity.py
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gpu.Context(local)
    uid = [gpu.comm_unique_id()] if rank == 0 else [None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    n_lat, n_lon, w, h = 60, 120, 320, 240
    vis, cam = scenes.c5_vis(), scenes.c5_camera()
    ds, _ = scenes.c5_partition_mesh(n_lat, n_lon, world, rank)
    part = scenes.build_partitions(gpu, vis, {"mesh": ds}, world, only_rank=rank, ctx=ctx)[0]
    keys = ["primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays"]
    results = {}
    only = os.environ.get("MP_PARITY_ONLY", "")   # e.g. "volumes": just the nineBalls modes (a short multi-GPU check)
    for mode in (("fused", "lists") if not only else ("fused",)):
        os.environ["GXY_FUSED"] = "1" if mode == "fused" else "0"
        st = gpu.render_device([part], cam, vis["lighting"], w, h, 0.001)
        t = torch.tensor([st[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        results[mode] = (dict(zip(keys, t.tolist())), part.download_rgba32f(w, h) if rank == 0 else None)
    os.environ.pop("GXY_FUSED", None)
    # frames in flight over the peer arenas (gxy_render_submit / gxy_render_wait): three cameras on three slots, submitted before
    # the first wait; every slot must deliver the frame of ITS camera
    fl_cams = [scenes.parse_camera({"viewpoint": vp, "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}) for vp in ([3, 2, -4], [-3, 1, -4], [0.5, 3, -4.5], [4, -1, 2])]
    depth = 3
    for k in range(depth):
        gpu.render_submit([part], fl_cams[k], vis["lighting"], w, h, 0.001, k)
    flights = []
    for k in range(len(fl_cams)):
        st = gpu.render_wait([part], k % depth)
        t = torch.tensor([st[key] for key in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        flights.append((dict(zip(keys, t.tolist())), part.download_rgba32f(w, h) if rank == 0 else None))
        if k + depth < len(fl_cams):
            gpu.render_submit([part], fl_cams[k + depth], vis["lighting"], w, h, 0.001, k % depth)
    # BASELINE.json configs[1]: tests/nineBalls.state (two volumes: DVR + isosurfaces + slice, shadows), volume bricks over the ranks
    st9 = scenes.parse_state(json.load(open(os.path.join(ROOT, "tests", "golden", "states", "nineBalls.state"))))
    ds9 = scenes.load_datasets(st9, scenes.default_data_provider(n=96))
    vis9, cam9, w9, h9 = st9["visualizations"][0], st9["cameras"][1], 256, 256
    part9 = scenes.build_partitions(gpu, vis9, ds9, world, only_rank=rank, ctx=ctx)[0]
    os.environ["GXY_VOLUME_FLIGHTS"] = "0"   # this mode: the synchronous NCCL list loop
    s9 = gpu.render_device([part9], cam9, vis9["lighting"], w9, h9, st9["epsilon"])
    t = torch.tensor([s9[k] for k in keys], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    results["nineBalls"] = (dict(zip(keys, t.tolist())), part9.download_rgba32f(w9, h9) if rank == 0 else None)
    # the same volumes with frames in flight: the list kernels with device-side lengths over the peer arenas (GXY_VOLUME_FLIGHTS=1),
    # three cameras of the state file on two frame slots
    os.environ["GXY_VOLUME_FLIGHTS"] = "1"
    vf = []
    cams9 = [st9["cameras"][k] for k in (1, 0, 2)]
    for k in range(2):
        gpu.render_submit([part9], cams9[k], vis9["lighting"], w9, h9, st9["epsilon"], k)
    for k in range(len(cams9)):
        s = gpu.render_wait([part9], k % 2)
        t = torch.tensor([s[key] for key in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        vf.append((dict(zip(keys, t.tolist())), part9.download_rgba32f(w9, h9) if rank == 0 else None))
        if k + 2 < len(cams9):
            gpu.render_submit([part9], cams9[k + 2], vis9["lighting"], w9, h9, st9["epsilon"], k % 2)
    os.environ.pop("GXY_VOLUME_FLIGHTS", None)
    if not only:
        # PathLines (round Bezier curves): the NCCL list path with the curve kernels, poly-lines cut at the partition planes
        pl = scenes.helix_pathlines(24)
        vis_pl = dict(annotation="", lighting=dict(lights=[[1.0, 2.0, -3.0]], types=[2], n_ao=2, ao_radius=0.5, shadows=True, Ka=0.4, Kd=0.6),
                      operators=[dict(type="PathLinesVis", dataset="lines", colormap=[[0.0, 0.0, 1.0, 0.0], [1.2, 1.0, 0.0, 1.0]],
                                      opacitymap=[[0, 1], [1, 1]], data_range=None, radius0=0.01, radius1=0.05, value0=0.0, value1=1.2)])
        cam_pl = dict(eye=[1.5, 1.0, -3.0], dir=[-1.5, -1.0, 3.0], up=[0.0, 1.0, 0.0], aov=35.0)
        part_pl = scenes.build_partitions(gpu, vis_pl, {"lines": pl}, world, only_rank=rank, ctx=ctx)[0]
        spl = gpu.render_device([part_pl], cam_pl, vis_pl["lighting"], w, h, 0.001)
        t = torch.tensor([spl[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        results["pathlines"] = (dict(zip(keys, t.tolist())), part_pl.download_rgba32f(w, h) if rank == 0 else None)
    ok = True
    if rank == 0:
        from oracle import oracle
        if not only:
            o_pl = scenes.build_partitions(oracle, vis_pl, {"lines": pl}, world)
            fb_opl, st_opl = oracle.render(o_pl, cam_pl, vis_pl["lighting"], w, h, 0.001)
        full, _ = scenes.c5_partition_mesh(n_lat, n_lon, 1, 0)
        o_parts = scenes.build_partitions(oracle, vis, {"mesh": full}, world)
        fb_o, st_o = oracle.render(o_parts, cam, vis["lighting"], w, h, 0.001)
        o9 = scenes.build_partitions(oracle, vis9, ds9, world)
        fb_o9, st_o9 = oracle.render(o9, cam9, vis9["lighting"], w9, h9, st9["epsilon"])
        for k, (st, fb) in enumerate(vf):
            fb_f, st_f = oracle.render(o9, cams9[k], vis9["lighting"], w9, h9, st9["epsilon"])
            frac = float((np.abs(fb[..., :3] - fb_f[..., :3]).max(-1) <= 1.0 / 255).mean())
            same = all(st[key] == st_f[key] for key in st)
            print(json.dumps({"mode": "nineBalls volume flight %d (slot %d)" % (k, k % 2), "world": world, "fraction_within_1_255": frac,
                              "stats_equal": same, "gpu": st, "oracle": {key: st_f[key] for key in st}}), flush=True)
            ok = ok and same and frac >= 0.999
        for k, (st, fb) in enumerate(flights):
            fb_f, st_f = oracle.render(o_parts, fl_cams[k], vis["lighting"], w, h, 0.001)
            frac = float((np.abs(fb[..., :3] - fb_f[..., :3]).max(-1) <= 1.0 / 255).mean())
            same = all(st[key] == st_f[key] for key in st)
            print(json.dumps({"mode": "flight %d (slot %d)" % (k, k % depth), "world": world, "fraction_within_1_255": frac, "stats_equal": same,
                              "gpu": st, "oracle": {key: st_f[key] for key in st}}), flush=True)
            ok = ok and same and frac >= 0.999
        for mode, (st, fb) in results.items():
            ref_fb, ref_st = (fb_o9, st_o9) if mode == "nineBalls" else (fb_opl, st_opl) if mode == "pathlines" else (fb_o, st_o)
            frac = float((np.abs(fb[..., :3] - ref_fb[..., :3]).max(-1) <= 1.0 / 255).mean())
            same = all(st[k] == ref_st[k] for k in st)
            print(json.dumps({"mode": mode, "world": world, "fraction_within_1_255": frac, "stats_equal": same, "gpu": st,
                              "oracle": {k: ref_st[k] for k in st}}), flush=True)
            ok = ok and same and frac >= 0.999
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
