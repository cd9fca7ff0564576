"""Helpers shared by the parity tests."""
import json
import os

import numpy as np

from galaxy_b200 import scenes

COLS = ["ox", "oy", "oz", "dx", "dy", "dz", "nx", "ny", "nz", "sample", "r", "g", "b", "o", "sr", "sg", "sb", "so", "t", "tMax",
        "x", "y", "type", "term", "classification"]
CI = {n: i for i, n in enumerate(COLS)}


def load_state(golden_dir, name, provider):
    st = scenes.parse_state(json.load(open(os.path.join(golden_dir, "states", name + ".state"))))
    return st, scenes.load_datasets(st, provider)


def icol(rays, name, n):
    return rays[CI[name], :n].view(np.int32)


def fcol(rays, name, n):
    return rays[CI[name], :n]


def image_fraction(a_rgba8, b_rgba8, tol=1):
    d = np.abs(a_rgba8[..., :3].astype(int) - b_rgba8[..., :3].astype(int)).max(-1)
    return float((d <= tol).mean())


def fb_fraction(fa, fb, tol=1.0 / 255):
    d = np.abs(fa[..., :3].astype(np.float64) - fb[..., :3].astype(np.float64)).max(-1)
    return float((d <= tol).mean())


def random_soup(n_tris, n_spheres, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.9, 0.9, (n_tris, 1, 3))
    verts = (c + rng.normal(0, 0.06, (n_tris, 3, 3))).reshape(-1, 3).astype(np.float32)
    idx = np.arange(n_tris * 3, dtype=np.int32).reshape(-1, 3)
    nrm = rng.normal(0, 1, (n_tris * 3, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    data = np.linalg.norm(verts, axis=1).astype(np.float32)
    tri = scenes.TrianglesDataset(verts, nrm, data, idx)
    pc = rng.uniform(-0.9, 0.9, (n_spheres, 3)).astype(np.float32)
    par = scenes.ParticlesDataset(pc, rng.uniform(0, 1, n_spheres).astype(np.float32))
    return tri, par


def soup_vis(with_particles=True, lighting=None):
    ops = [dict(type="TrianglesVis", dataset="tris", colormap=[[0.0, 1.0, 0.2, 0.2], [1.6, 0.2, 0.2, 1.0]], opacitymap=[[0, 1], [1, 1]],
                data_range=None)]
    if with_particles:
        ops.append(dict(type="ParticlesVis", dataset="parts", colormap=[[0.0, 0.2, 1.0, 0.2], [1.0, 1.0, 1.0, 0.2]], opacitymap=[[0, 1], [1, 1]],
                        data_range=None, radius0=0.02, radius1=0.06, value0=0.0, value1=1.0))
    lighting = lighting or dict(lights=[[1.0, 2.0, -3.0]], types=[2], n_ao=4, ao_radius=0.5, shadows=True, Ka=0.3, Kd=0.7)
    return dict(annotation="", lighting=lighting, operators=ops)


def random_rays(n, seed):
    rng = np.random.default_rng(seed)
    org = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-0.8, 0.8, (n, 3)).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return org, d.astype(np.float32)
