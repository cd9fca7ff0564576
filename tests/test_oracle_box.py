"""The oracle's Box restatements against the reference's own src/data/Box.cpp, compiled from where it lies into oracle/_ref/
(make -C oracle ref; the .so travels to the GPU box): exit_face, intersect and the (origin, counts, deltas) constructor, bit for
bit, on random rays and on the edge cases the reference's thresholds create (|d| around 1e-4, origins on faces, rays that miss,
zero direction components).  Also pins the Python partition boxes (scenes.volume_boxes) that both back ends are fed."""
import ctypes as C
import os

import numpy as np
import pytest

from galaxy_b200 import scenes
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libgxy_box_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libgxy_box_ref.so not built (needs /root/reference at build time)")

fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


def _f(a):
    return a.ctypes.data_as(fp)


def _i(a):
    return a.ctypes.data_as(ip)


def cases(n, seed):
    rng = np.random.default_rng(seed)
    lo = rng.uniform(-1.5, 0.5, (n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0.05, 2.0, (n, 3))).astype(np.float32)
    boxes = np.concatenate([lo, hi], 1)
    org = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    inside = rng.random(n) < 0.4  # exit_face is called with origins inside or on the box, intersect with the camera outside
    org[inside] = (lo[inside] + (hi[inside] - lo[inside]) * rng.random((inside.sum(), 3))).astype(np.float32)
    d = rng.normal(0, 1, (n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # edge cases: components around the +-1e-4 threshold of exit_face, exact zeros, origins exactly on a face
    k = n // 10
    d[:k, 0] = rng.choice(np.array([1e-4, -1e-4, 0.99e-4, 1.01e-4, -1.01e-4, 0.0], np.float32), k)
    d[k:2 * k, 1] = 0.0
    d[2 * k:3 * k, 2] = rng.choice(np.array([1e-4, -1e-4, 0.0], np.float32), k)
    org[3 * k:4 * k, 0] = lo[3 * k:4 * k, 0]
    org[4 * k:5 * k, 1] = hi[4 * k:5 * k, 1]
    rays = np.ascontiguousarray(np.concatenate([org, d.astype(np.float32)], 1))
    return np.ascontiguousarray(boxes), rays


def test_exit_face_and_intersect_bit_exact():
    ref = C.CDLL(REF)
    olib = oracle.lib()
    n = 200000
    boxes, rays = cases(n, 7)
    f_ref, f_or = np.zeros(n, np.int32), np.zeros(n, np.int32)
    ref.gxref_exit_face(n, _f(boxes), _f(rays), _i(f_ref))
    olib.gxo_exit_face(n, _f(boxes), _f(rays), _i(f_or))
    assert np.array_equal(f_ref, f_or)
    assert set(np.unique(f_ref)) == {0, 1, 2, 3, 4, 5}
    h_ref, h_or = np.zeros(n, np.int32), np.zeros(n, np.int32)
    t_ref, t_or = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
    with np.errstate(all="ignore"):
        ref.gxref_box_intersect(n, _f(boxes), _f(rays), _i(h_ref), _f(t_ref))
        olib.gxo_box_intersect(n, _f(boxes), _f(rays), _i(h_or), _f(t_or))
    assert np.array_equal(h_ref, h_or)
    assert 0.1 < h_ref.mean() < 0.95
    hit = h_ref == 1
    assert np.array_equal(t_ref[hit].view(np.int32), t_or[hit].view(np.int32))  # bit for bit where the result is defined


def test_partition_boxes_match_reference_box_constructor():
    """Volume::local_import builds global_box = Box(origin + delta, counts - 2, deltas) and local_box = Box(origin + offset * delta,
    local counts, deltas) (Volume.cpp:379-390); scenes.volume_boxes restates that in numpy float32."""
    ref = C.CDLL(REF)
    for n, nparts in ((256, 1), (256, 8), (130, 4), (67, 2), (1024, 8)):
        sp = float("%f" % (2.0 / (n - 1)))
        vol = scenes.VolumeDataset([-1.0, -1.0, -1.0], (n, n, n), [sp, sp, sp], np.zeros((1, 1, 1), np.float32).repeat(n, 0).repeat(n, 1).repeat(n, 2)
                                   if n <= 130 else np.broadcast_to(np.zeros((1, 1, 1), np.float32), (n, n, n)))
        fac = scenes.factor(nparts)
        for part in scenes.partition(fac, vol.counts):
            gmin, gmax, lmin, lmax = scenes.volume_boxes(vol, part)
            out = np.zeros(6, np.float32)
            go = (vol.origin + vol.deltas).astype(np.float32)
            gc = np.array([c - 2 for c in vol.counts], np.int32)
            ref.gxref_box_from_grid(_f(np.ascontiguousarray(go)), _i(gc), _f(np.ascontiguousarray(vol.deltas)), _f(out))
            assert np.array_equal(out[:3], gmin) and np.array_equal(out[3:], gmax)
            lo = np.array([vol.origin[a] + np.float32(part["offsets"][a]) * vol.deltas[a] for a in range(3)], np.float32)
            lc = np.array(part["counts"], np.int32)
            ref.gxref_box_from_grid(_f(lo), _i(lc), _f(np.ascontiguousarray(vol.deltas)), _f(out))
            assert np.array_equal(out[:3], lmin) and np.array_equal(out[3:], lmax)
