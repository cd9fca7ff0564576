"""The reference's distributed ray loop, list by list, on top of the per-RayList entry points.

This is the host-side protocol of `processRays_task::work` + `SendRaysMsg` + `SendPixelsMsg` + the
RenderingSet termination check (src/renderer/Renderer.cpp:535-643, 732-836; RenderingSet.cpp:483-531)
restated bulk-synchronously over `torch.distributed`: every rank owns one spatial partition (a Scene of
either backend: `galaxy_b200.gpu` = the C ABI of the CUDA library, `oracle.oracle` = the CPU checker),
and per wave
    trace the local list (TraceRays::Trace)  ->  classify (Renderer::Classify/AssignDestinations)  ->
    add TERMINATED rays to the partial framebuffer (Rendering::AddLocalPixels)  ->
    exchange the rays whose classification is a rank (all_to_all of the live columns)  ->
    all_reduce the number of rays left (termination)
and one reduce of the partial framebuffers to rank 0 at the end.  It is what INTEGRATION.md section 4
describes (the reference's own loop running unchanged on the new trace call), it is how the N>1 path is
covered on CPU (gloo, world_size 2), and it is NOT the fast path: `gxy_render` keeps everything on the
devices and exchanges with NCCL.  `sample_distributed` is the same loop for the Sampler (src/sampler): the way
a sampling Visualization runs with one process per partition today (gxy_sample is single-process).
"""
import numpy as np
import torch
import torch.distributed as dist

COLS = {n: i for i, n in enumerate(["ox", "oy", "oz", "dx", "dy", "dz", "nx", "ny", "nz", "sample", "r", "g", "b", "o", "sr", "sg", "sb", "so",
                                    "t", "tMax", "x", "y", "type", "term", "classification"])}
TERMINATED = -1
KEEP_HERE = -3


def _icol(rays, name, n):
    return rays[COLS[name], :n].view(np.int32)


def _pack(cols):
    n = cols.shape[1]
    al = max(16, (n + 15) & ~15)
    out = np.zeros((25, al), np.float32)
    out[:, :n] = cols
    return out, n


def render_distributed(scene, resolve_lights, camera, lighting, w, h, epsilon=0.001, group=None):
    """Render one frame with one partition per rank.  `scene` is this rank's Scene; `resolve_lights` the
    backend's resolve_lights function.  Returns (fb on rank 0 else None, stats dict summed over ranks)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    L = resolve_lights(lighting, camera)
    fb = np.zeros((h, w, 4), np.float32)
    stats = dict(primary_rays=0, shadow_rays=0, ao_rays=0, forwarded_rays=0, terminated_rays=0, traced_rays=0, waves=0)
    rays, n = scene.generate_rays(camera, w, h)
    stats["primary_rays"] = n
    pending = [(rays, n)] if n else []
    while True:
        keep = []  # live columns of the rays that stay for the next wave (spawned here or KEEP_HERE)
        outgoing = [[] for _ in range(world)]
        for rays, n in pending:
            sec, nsec, _ = scene.trace_raylist(L, rays, n, epsilon)
            stats["traced_rays"] += n
            stats["waves"] += 1
            if nsec:
                typ = _icol(sec, "type", nsec)
                stats["shadow_rays"] += int((typ == 2).sum())
                stats["ao_rays"] += int((typ == 4).sum())
                keep.append(sec[:, :nsec].copy())
            scene.classify(rays, n)
            cls = _icol(rays, "classification", n)
            term = cls == TERMINATED
            if term.any():  # Rendering::AddLocalPixels
                x, y = _icol(rays, "x", n)[term], _icol(rays, "y", n)[term]
                ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
                for c, name in enumerate("rgbo"):
                    np.add.at(fb[..., c], (y[ok], x[ok]), rays[COLS[name], :n][term][ok])
                stats["terminated_rays"] += int(term.sum())
            for dst in range(world):
                m = cls == dst
                if m.any():
                    outgoing[dst].append(rays[:, :n][:, m].copy())
                    stats["forwarded_rays"] += int(m.sum())
            m = cls == KEEP_HERE
            if m.any():
                keep.append(rays[:, :n][:, m].copy())
        # ---- exchange (SendRaysMsg): counts first, then the columns
        send = [np.concatenate(o, axis=1) if o else np.zeros((25, 0), np.float32) for o in outgoing]
        counts = torch.tensor([s.shape[1] for s in send], dtype=torch.int64)
        all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        recv_counts = [int(all_counts[src][rank]) for src in range(world)]
        recv = [torch.zeros((25, c), dtype=torch.float32) for c in recv_counts]
        reqs = []
        for peer in range(world):
            if peer == rank:
                continue
            if send[peer].shape[1]:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send[peer])), peer, group=group))
            if recv_counts[peer]:
                reqs.append(dist.irecv(recv[peer], peer, group=group))
        for r in reqs:
            r.wait()
        if send[rank].shape[1]:
            keep.append(send[rank])
        for peer in range(world):
            if peer != rank and recv_counts[peer]:
                keep.append(recv[peer].numpy())
        pending = []
        if keep:
            cols = np.concatenate(keep, axis=1)
            for a in range(0, cols.shape[1], 1000000):  # max_rays_per_packet (Renderer.cpp:134)
                pending.append(_pack(cols[:, a:a + 1000000]))
        left = torch.tensor([sum(n for _, n in pending)], dtype=torch.int64)
        dist.all_reduce(left, group=group)  # RenderingSet::SynchronousCheckMsg (RenderingSet.cpp:483-531)
        if int(left) == 0:
            break
    t = torch.from_numpy(fb)
    dist.reduce(t, 0, group=group)  # SendPixelsMsg towards the image owner
    keys = sorted(stats)
    s = torch.tensor([stats[k] for k in keys], dtype=torch.int64)
    dist.all_reduce(s, group=group)
    return (t.numpy() if rank == 0 else None), dict(zip(keys, s.tolist()))


def _exchange(outgoing, rank, world, group):
    """SendRaysMsg: counts first, then the columns; returns the lists received (incl. the ones addressed to this rank itself)"""
    send = [np.concatenate(o, axis=1) if o else np.zeros((25, 0), np.float32) for o in outgoing]
    counts = torch.tensor([s.shape[1] for s in send], dtype=torch.int64)
    all_counts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    recv_counts = [int(all_counts[src][rank]) for src in range(world)]
    recv = [torch.zeros((25, c), dtype=torch.float32) for c in recv_counts]
    reqs = []
    for peer in range(world):
        if peer == rank:
            continue
        if send[peer].shape[1]:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send[peer])), peer, group=group))
        if recv_counts[peer]:
            reqs.append(dist.irecv(recv[peer], peer, group=group))
    for r in reqs:
        r.wait()
    got = [send[rank]] if send[rank].shape[1] else []
    got += [recv[peer].numpy() for peer in range(world) if peer != rank and recv_counts[peer]]
    return got


def sample_distributed(scene, camera, w, h, group=None):
    """The Sampler (src/sampler/Sampler.cpp:52-133) with one partition per rank: per wave SamplerTraceRays on the local lists,
    Classify, one sample per ray whose term has RAY_SURFACE (kept by THIS rank, as every rank keeps its own Particles), the
    KEEP_HERE rays (those that left a sample) stay, the BOUNDARY rays move to their neighbour.  `scene` is this rank's
    sampling Scene of either backend.  Returns (this rank's samples (n,3), stats dict summed over ranks)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    stats = dict(primary_rays=0, forwarded_rays=0, traced_rays=0, waves=0, samples=0)
    rays, n = scene.generate_rays(camera, w, h)
    stats["primary_rays"] = n
    pending = [(rays, n)] if n else []
    samples = []
    while True:
        keep = []
        outgoing = [[] for _ in range(world)]
        for rays, n in pending:
            scene.sample_raylist(rays, n)
            stats["traced_rays"] += n
            stats["waves"] += 1
            scene.classify(rays, n)
            hit = (_icol(rays, "term", n) & 1) != 0  # RAY_SURFACE
            if hit.any():
                t = rays[COLS["t"], :n][hit]
                samples.append(np.stack([rays[COLS["ox"], :n][hit] + t * rays[COLS["dx"], :n][hit],
                                         rays[COLS["oy"], :n][hit] + t * rays[COLS["dy"], :n][hit],
                                         rays[COLS["oz"], :n][hit] + t * rays[COLS["dz"], :n][hit]], 1).astype(np.float32))
            cls = _icol(rays, "classification", n)
            for dst in range(world):
                m = cls == dst
                if m.any():
                    outgoing[dst].append(rays[:, :n][:, m].copy())
                    stats["forwarded_rays"] += int(m.sum())
            m = cls == KEEP_HERE
            if m.any():
                keep.append(rays[:, :n][:, m].copy())
        keep += _exchange(outgoing, rank, world, group)
        pending = []
        if keep:
            cols = np.concatenate(keep, axis=1)
            for a in range(0, cols.shape[1], 1000000):
                pending.append(_pack(cols[:, a:a + 1000000]))
        left = torch.tensor([sum(n for _, n in pending)], dtype=torch.int64)
        dist.all_reduce(left, group=group)
        if int(left) == 0:
            break
    mine = np.concatenate(samples) if samples else np.zeros((0, 3), np.float32)
    stats["samples"] = len(mine)
    keys = sorted(stats)
    s = torch.tensor([stats[k] for k in keys], dtype=torch.int64)
    dist.all_reduce(s, group=group)
    return mine, dict(zip(keys, s.tolist()))
