#!/usr/bin/env python
"""Writes closed-form stand-ins for the datasets of the reference's tests/data-driven.state, so that this state file -- a volume with
slices and DVR, path lines, particles and a triangle mesh in one Visualization -- renders unchanged with galaxy_b200/gxywriter.

The reference makes them with VTK (tests/create_data_driven_datasets.vpy): stream lines of the field (-y, x, 0.1) from 5 seeds on the
segment (-0.7,-0.7,-0.9)..(0.7,0.7,-0.9) with the scalar oneBall = |p|; the isosurface oneBall = 1.4 as a mesh with the scalar
eightBalls; 2000 points of the isosurface eightBalls = 0.3 with the scalar oneBall -- all on a 65^3 grid.  Without VTK the same objects
are regenerated from their definitions: the stream lines of that field are helices (exact, instead of the tracer's Runge-Kutta steps);
the contours are rebuilt as vtkContourFilter builds them on image data as far as a renderer can tell -- the same points on the grid
edges in the same order (the particle subset depends on it), gradient normals, interpolated scalars.  The oracle's render of the
UNCHANGED state file on these datasets is within 1/255 of the reference's gold on 99.99 % of the pixels.

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


def _grid_fields(n=65):
    """the 65^3 image data of tests/create_data_driven_datasets.vpy: oneBall = |p|, eightBalls, float32 on [-1,1]^3"""
    c = -1 + (2.0 / (n - 1)) * np.arange(n)
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    one = np.sqrt(X * X + Y * Y + Z * Z).astype(np.float32)
    eight = np.sqrt((np.abs(X) - .5) ** 2 + (np.abs(Y) - .5) ** 2 + (np.abs(Z) - .5) ** 2).astype(np.float32)
    return one, eight


def contour(field, value, other, origin=-1.0):
    """What vtkContourFilter leaves of an isosurface on image data, as far as the renderer can see it: one point per grid edge the
    value crosses, at t = (value - s0)/(s1 - s0), in the filter's order (slice by slice, row by row, the x, y, z edge of every grid
    point); per point the normalised central-difference gradient and `other` interpolated along the edge; one polygon per cell,
    fan-triangulated (VTK's case tables may cut a cell's polygon along another diagonal: a sub-pixel difference for these spheres).
    field/other: [k, j, i] float32.  Returns (points, normals, data, triangles)."""
    n = field.shape[0]
    h = 2.0 / (n - 1)
    f, g_other = field.astype(np.float64), other.astype(np.float64)
    v = f >= value
    gz, gy, gx = np.gradient(f, h)
    vid = -np.ones((3, n, n, n), np.int64)
    K, J, I = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    keys, recs = [], []
    for axis, (dk, dj, di) in enumerate(((0, 0, 1), (0, 1, 0), (1, 0, 0))):      # x edge, y edge, z edge of a grid point
        sl0 = (slice(0, n - dk), slice(0, n - dj), slice(0, n - di))
        sl1 = (slice(dk, n), slice(dj, n), slice(di, n))
        cross = v[sl0] != v[sl1]
        s0, s1 = f[sl0][cross], f[sl1][cross]
        t = (value - s0) / (s1 - s0)
        k, j, i = K[sl0][cross], J[sl0][cross], I[sl0][cross]
        p = np.stack([origin + h * (i + (t if axis == 0 else 0)), origin + h * (j + (t if axis == 1 else 0)), origin + h * (k + (t if axis == 2 else 0))], 1)
        g0 = np.stack([gx[sl0][cross], gy[sl0][cross], gz[sl0][cross]], 1)
        g1 = np.stack([gx[sl1][cross], gy[sl1][cross], gz[sl1][cross]], 1)
        recs.append((np.full(len(k), axis), k, j, i, p, g0 + t[:, None] * (g1 - g0), g_other[sl0][cross] + t * (g_other[sl1][cross] - g_other[sl0][cross])))
        keys.append(((k * n + j) * n + i) * 3 + axis)
    o = np.argsort(np.concatenate(keys), kind="stable")
    ax, k_, j_, i_, P, G, D = [np.concatenate([r[q] for r in recs])[o] for q in range(7)]
    vid[ax, k_, j_, i_] = np.arange(len(P))
    N = G / np.linalg.norm(G, axis=1, keepdims=True)
    edges = [(0, 0, 0, 0), (0, 0, 1, 0), (0, 1, 0, 0), (0, 1, 1, 0), (1, 0, 0, 0), (1, 0, 0, 1), (1, 1, 0, 0), (1, 1, 0, 1),
             (2, 0, 0, 0), (2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1)]
    any_edge = np.zeros((n - 1, n - 1, n - 1), bool)
    for (a, dk, dj, di) in edges:
        any_edge |= vid[a, dk:n - 1 + dk, dj:n - 1 + dj, di:n - 1 + di] >= 0
    tris = []
    for (k, j, i) in np.argwhere(any_edge):
        ids = [int(vid[a, k + dk, j + dj, i + di]) for (a, dk, dj, di) in edges]
        ids = [x for x in ids if x >= 0]
        if len(ids) < 3:
            continue
        pts = P[ids]
        c, nm = pts.mean(0), N[ids].mean(0)
        nm /= np.linalg.norm(nm)
        u = np.cross(nm, [1.0, 0.0, 0.0]) if abs(nm[0]) < 0.9 else np.cross(nm, [0.0, 1.0, 0.0])
        u /= np.linalg.norm(u)
        w = np.cross(nm, u)
        ring = [ids[q] for q in np.argsort(np.arctan2((pts - c) @ w, (pts - c) @ u))]
        tris += [[ring[0], ring[q], ring[q + 1]] for q in range(1, len(ring) - 1)]
    return P.astype(np.float32), N.astype(np.float32), D.astype(np.float32), np.asarray(tris, np.int32)


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
    one, eight = _grid_fields()
    # particles (do_particles): every pSkip-th point of the contour eightBalls = 0.3, 2000 of them, scalar oneBall
    cp, _, cd, _ = contour(eight, 0.3, one)
    pskip = max(1, int(float(len(cp)) / 2000))
    sel = np.arange(min(2000, len(cp))) * pskip
    parts_ds = scenes.ParticlesDataset(cp[sel], cd[sel])
    # mesh (do_mesh): the contour oneBall = 1.4 (the corner caps of that sphere inside the cube), scalar eightBalls
    mv, mn, md, mt = contour(one, 1.4, eight)
    mesh_ds = scenes.TrianglesDataset(mv, mn, md, mt)
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
