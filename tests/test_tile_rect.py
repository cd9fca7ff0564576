"""CPU: the screen rectangle over which a rank generates primary rays across processes (gxy_debug_tile_rect = peer_tile_rect, host
arithmetic of libgxy_b200) must contain EVERY pixel whose ray touches the rank's box -- otherwise that rank would silently never
originate those rays.  Checked against the oracle's Camera::generate_initial_rays restatement (which keeps a pixel iff its ray hits
the box when the box is the whole data), for perspective and orthographic cameras, boxes on and off the axis, near and far."""
import numpy as np
import pytest

from galaxy_b200 import gpu, scenes
from oracle import oracle


def pixels_hitting_box(cam, w, h, lo, hi):
    s = oracle.Scene()
    s.set_partition(lo, hi, lo, hi, [-1] * 6)
    s.commit()
    rays, n = s.generate_rays(cam, w, h)
    return rays[20, :n].view(np.int32).copy(), rays[21, :n].view(np.int32).copy()


CASES = [
    ({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [-1, -1, -1], [0, 0, 0]),
    ({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [0, 0, 0], [1, 1, 1]),
    ({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [0, -1, -1], [1, 0, 0]),
    ({"viewpoint": [-3, 1, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [-1, 0, 0], [0, 1, 1]),
    ({"viewpoint": [0.5, 3, -4.5], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 45}, [-1, -1, 0], [1, 0, 1]),
    ({"viewpoint": [0, 0, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [-1, -1, -1], [1, 1, 1]),
    ({"viewpoint": [4, -1, 2], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 20}, [0.2, -0.9, -0.3], [0.5, -0.1, 0.4]),
    ({"viewpoint": [0, 0, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 0}, [-0.5, -0.25, -1], [0.75, 0.5, 1]),     # orthographic
    ({"viewpoint": [0, 0, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}, [5, 5, -1], [6, 6, 1]),                  # off screen
]


@pytest.mark.parametrize("k", range(len(CASES)))
@pytest.mark.parametrize("size", [(320, 180), (257, 131)])
def test_rect_contains_every_pixel_whose_ray_touches_the_box(k, size):
    cam_doc, lo, hi = CASES[k]
    cam = scenes.parse_camera(cam_doc)
    w, h = size
    rc, (x0, y0, nx, ny) = gpu.debug_tile_rect(cam, w, h, lo, hi)
    xs, ys = pixels_hitting_box(cam, w, h, lo, hi)
    if rc == 0:
        return  # whole image: trivially complete
    assert rc == 1
    if len(xs) == 0:
        return
    assert nx > 0 and ny > 0
    assert xs.min() >= 8 * x0 and xs.max() < 8 * (x0 + nx), (xs.min(), xs.max(), x0, nx)
    assert ys.min() >= 4 * y0 and ys.max() < 4 * (y0 + ny), (ys.min(), ys.max(), y0, ny)
    # and it is not vacuous: for a box that covers part of the screen the rectangle is smaller than the image
    tiles = ((w + 7) // 8) * ((h + 3) // 4)
    print(cam_doc["viewpoint"], lo, hi, "pixels", len(xs), "rect tiles", nx * ny, "of", tiles)


def test_eye_inside_or_behind_the_box_scans_everything():
    cam = scenes.parse_camera({"viewpoint": [0.1, 0.1, 0.1], "viewcenter": [1, 0, 0], "viewup": [0, 1, 0], "aov": 60})
    rc, _ = gpu.debug_tile_rect(cam, 64, 64, [-1, -1, -1], [1, 1, 1])
    assert rc == 0
    assert gpu.lib().gxy_debug_tile_rect(None, 64, 64, None, None, None) == -1
