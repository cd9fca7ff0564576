"""The interactive / asynchronous frame path: frame-stamped accumulation (Rendering::AddLocalPixels +
ACCUMULATE_PIXEL without GXY_WRITE_IMAGES, src/renderer/Rendering.cpp:104-153).  CPU: the oracle applies the reference's
per-contribution rule literally; its consequences are checked.  GPU: the one-pass merge of the CUDA path against it."""
import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util

CAM_A = dict(eye=[3.0, 2.0, -4.0], dir=[-3.0, -2.0, 4.0], up=[0.0, 1.0, 0.0], aov=30.0)
CAM_B = dict(eye=[1.2, 0.4, -2.2], dir=[-0.2, -0.4, 2.2], up=[0.0, 1.0, 0.0], aov=30.0)      # moved closer and to the side
W, H = 160, 120


def scene(backend, nparts):
    tri, par = util.random_soup(400, 150, 3)
    vis = util.soup_vis(with_particles=True)
    return scenes.build_partitions(backend, vis, {"tris": tri, "parts": par}, nparts), vis


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def test_stamped_accumulation_rules(oracle):
    parts, vis = scene(oracle, 1)
    L = vis["lighting"]
    plain_a, _ = oracle.render(parts, CAM_A, L, W, H)
    plain_b, _ = oracle.render(parts, CAM_B, L, W, H)
    r = oracle.ProgressiveRendering(W, H)
    # frame 0 on zeroed stamps: plain accumulation
    fb0, _ = r.render(parts, CAM_A, L, 0)
    assert np.array_equal(fb0, plain_a) and r.frame[0] == 0
    # the same frame number again: contributions add on top (no pixel has an older stamp)
    fb0b, _ = r.render(parts, CAM_A, L, 0)
    assert np.allclose(fb0b, 2 * plain_a, rtol=1e-6, atol=1e-7)
    # a newer frame from another camera: pixels it writes are replaced, all others keep showing the old image
    fb1, _ = r.render(parts, CAM_B, L, 1)
    rays, n = parts[0].generate_rays(CAM_B, W, H)
    mask = np.zeros((H, W), bool)
    mask[util.icol(rays, "y", n), util.icol(rays, "x", n)] = True
    assert 0.1 < mask.mean() < 1.0 or mask.all()
    assert np.array_equal(fb1[mask], plain_b[mask]) and np.array_equal(fb1[~mask], fb0b[~mask])
    assert np.array_equal(r.kbuffer[mask], np.ones(mask.sum(), np.int32)) and (r.kbuffer[~mask] == 0).all() and r.frame[0] == 1
    # a stale frame is dropped entirely
    fb_stale, st = r.render(parts, CAM_A, L, 0)
    assert np.array_equal(fb_stale, fb1) and st["terminated_rays"] == 0 and r.frame[0] == 1
