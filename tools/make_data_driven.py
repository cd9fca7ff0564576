#!/usr/bin/env python
"""Writes closed-form stand-ins for the datasets of the reference's tests/data-driven.state, so that this state file -- a volume with
slices and DVR, path lines, particles and a triangle mesh in one Visualization -- renders unchanged with galaxy_b200/gxywriter.

The reference makes them with VTK (tests/create_data_driven_datasets.vpy): stream lines of the field (-y, x, 0.1) from 5 seeds on the
segment (-0.7,-0.7,-0.9)..(0.7,0.7,-0.9) with the scalar oneBall = |p|; the isosurface oneBall = 1.4 as a mesh with the scalar
eightBalls; 2000 points of the isosurface eightBalls = 0.3 with the scalar oneBall.  Without VTK the same objects are written in
closed form: the stream lines of that field are helices, the isosurfaces are spheres (so the images resemble the reference's gold,
but are not comparable with it pixel by pixel: VTK's contouring and its Runge-Kutta steps are not reproduced).

  python tools/make_data_driven.py [-o outdir] [-P nparts] [-n 256]
  cp tests/golden/states/data-driven.state outdir/ && cd outdir && <repo>/galaxy_b200/gxywriter -P nparts data-driven.state
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from galaxy_b200 import scenes  # noqa: E402


def eightballs(p):
    return np.sqrt(((np.abs(p) - 0.5) ** 2).sum(-1))


def make_datasets():
    """(PathLinesDataset, ParticlesDataset, TrianglesDataset) as described above, float32"""
    # stream lines: p(t) = (r cos(a + t), r sin(a + t), z0 + 0.1 t) until the line leaves [-1,1]^3
    pts, lines, k = [], [], 0
    for s in np.linspace(0.0, 1.0, 5):
        x0, y0, z0 = -0.7 + 1.4 * s, -0.7 + 1.4 * s, -0.9
        r, a = np.hypot(x0, y0), np.arctan2(y0, x0)
        t = np.arange(0.0, 19.0 + 1e-9, 0.05)
        p = np.stack([r * np.cos(a + t), r * np.sin(a + t), z0 + 0.1 * t], 1) if r > 0 else np.stack([0 * t, 0 * t, z0 + 0.1 * t], 1)
        inside = np.all(np.abs(p) <= 1.0, axis=1)
        n = int(np.argmin(inside)) if not inside.all() else len(p)      # up to the first point outside
        if n >= 2:
            pts.append(p[:n]); lines.append(list(range(k, k + n))); k += n
    pts = np.concatenate(pts).astype(np.float32)
    lines_ds = scenes.PathLinesDataset(pts, np.linalg.norm(pts, axis=1), lines)
    # particles: 250 Fibonacci points on each of the 8 spheres |p - c| = 0.3
    i = np.arange(250) + 0.5
    phi, theta = np.arccos(1 - 2 * i / 250), np.pi * (1 + 5 ** 0.5) * i
    unit = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1)
    cs = np.array([[sx, sy, sz] for sz in (-.5, .5) for sy in (-.5, .5) for sx in (-.5, .5)])
    pp = np.concatenate([c + 0.3 * unit for c in cs]).astype(np.float32)
    parts_ds = scenes.ParticlesDataset(pp, np.linalg.norm(pp, axis=1))
    # mesh: the sphere |p| = 1.4 inside the cube (its eight corner caps), analytic normals, scalar = eightBalls
    n_lat, n_lon = 384, 768
    la, lo = np.meshgrid(np.linspace(0, np.pi, n_lat + 1), np.linspace(0, 2 * np.pi, n_lon, endpoint=False), indexing="ij")
    nrm = np.stack([np.sin(la) * np.cos(lo), np.sin(la) * np.sin(lo), np.cos(la)], -1).reshape(-1, 3)
    v = 1.4 * nrm
    idx = lambda a, b: a * n_lon + (b % n_lon)
    tris = []
    for a in range(n_lat):
        for b in range(n_lon):
            q = [idx(a, b), idx(a + 1, b), idx(a + 1, b + 1), idx(a, b + 1)]
            tris += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
    tris = np.asarray(tris, np.int64)
    keep = np.all(np.abs(v[tris]) <= 1.0, axis=(1, 2))
    tris = tris[keep]
    used = np.zeros(len(v), bool)
    used[tris.ravel()] = True
    remap = np.cumsum(used) - 1
    v32 = v[used].astype(np.float32)
    mesh_ds = scenes.TrianglesDataset(v32, nrm[used].astype(np.float32), eightballs(v32.astype(np.float64)).astype(np.float32), remap[tris].astype(np.int32))
    return lines_ds, parts_ds, mesh_ds


def write(outdir, nparts, n=256, mode="appended-raw"):
    from tests.vtu_writer import write_vtu
    os.makedirs(outdir, exist_ok=True)
    lines, parts, mesh = make_datasets()
    ext, _ = scenes.geometry_extents(nparts)
    docs = {"streamlines": [], "eightBalls-points": [], "oneBall-mesh": []}
    for r in range(nparts):
        l = lines if nparts == 1 else scenes.clip_pathlines(lines, ext[r])
        p = parts if nparts == 1 else scenes.clip_particles(parts, ext[r])
        m = mesh if nparts == 1 else scenes.clip_triangles(mesh, ext[r])
        write_vtu(os.path.join(outdir, "streamlines-%d.vtu" % r), l.points, scalars=l.data, scalars_name="oneBall", mode=mode, polylines=[list(x) for x in l.lines])
        write_vtu(os.path.join(outdir, "eightBalls-points-%d.vtu" % r), p.centers, None, None, p.data, scalars_name="oneBall", mode=mode)
        write_vtu(os.path.join(outdir, "oneBall-mesh-%d.vtu" % r), m.verts, m.indices, m.normals, m.data, scalars_name="eightBalls", mode=mode)
        for name in docs:
            docs[name].append({"filename": "%s-%d.vtu" % (name, r), "extent": [float(x) for x in ext[r]]})
    for name, parts_doc in docs.items():
        json.dump({"parts": parts_doc}, open(os.path.join(outdir, name + ".part"), "w"))
    vol = scenes.radial_volume("eightBalls", n)
    base = os.path.join(outdir, "radial-eightBalls")
    with open(base + ".vol", "w") as f:
        f.write("float\n%f %f %f\n%d %d %d\n%f %f %f\nradial-eightBalls.raw\n" % (*[float(x) for x in vol.origin], *vol.counts, *[float(x) for x in vol.deltas]))
    vol.data.astype(np.float32).tofile(base + ".raw")
    return lines, parts, mesh, vol


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", default=".")
    ap.add_argument("-P", type=int, default=1)
    ap.add_argument("-n", type=int, default=256)
    a = ap.parse_args()
    l, p, m, v = write(a.o, a.P, a.n)
    print("wrote %d stream lines (%d vertices), %d particles, %d triangles, radial-eightBalls %s in %s for %d partition(s)" % (
        len(l.lines), len(l.points), len(p.centers), len(m.indices), v.counts, a.o, a.P))
