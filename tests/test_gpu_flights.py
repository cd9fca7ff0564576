"""-m gpu: frames in flight (gxy_render_submit / gxy_render_wait) -- a RenderingSet whose frames overlap on the device, as the
reference keeps every Rendering of a set in flight (src/apps/gxywriter.cpp:196-264).  Every slot must deliver its own frame:
images and ray statistics are compared with the oracle per camera."""
import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    return g


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def cameras():
    return [scenes.parse_camera({"viewpoint": vp, "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})
            for vp in ([3, 2, -4], [-3, 1, -4], [0.5, 3, -4.5], [4, -1, 2], [-2, -2, -4])]


def test_frames_in_flight_deliver_their_own_images(gpu, oracle):
    """5 cameras on 3 slots, submitted ahead of the waits: every frame equals the oracle's frame for ITS camera."""
    tri = scenes.eightballs_mesh(24, 48)
    vis = scenes.c5_vis()
    g = scenes.build_partitions(gpu, vis, {"mesh": tri}, 1)
    o = scenes.build_partitions(oracle, vis, {"mesh": tri}, 1)
    cams = cameras()
    w, h, depth = 320, 180, 3
    got = []
    for k in range(min(depth, len(cams))):
        gpu.render_submit(g, cams[k], vis["lighting"], w, h, 0.001, k % depth)
    for k in range(len(cams)):
        st = gpu.render_wait(g, k % depth)
        got.append((g[0].download_rgba32f(w, h), st))
        if k + depth < len(cams):
            gpu.render_submit(g, cams[k + depth], vis["lighting"], w, h, 0.001, k % depth)
    for k, cam in enumerate(cams):
        fb_o, st_o = oracle.render(o, cam, vis["lighting"], w, h, 0.001)
        fb_g, st_g = got[k]
        for key in ("primary_rays", "shadow_rays", "ao_rays", "terminated_rays", "traced_rays"):
            assert st_g[key] == st_o[key], (k, key, st_g, st_o)
        frac = util.fb_fraction(fb_g, fb_o)
        print("camera", k, "fraction %.6f" % frac, "dequeued", st_g["dequeued_rays"], "of", st_g["traced_rays"], "t %.3f..%.3f ms" % (st_g["t_begin_ms"], st_g["t_end_ms"]))
        assert frac >= 0.999
        assert 0 < st_g["dequeued_rays"] <= st_g["traced_rays"]
        assert st_g["t_end_ms"] > st_g["t_begin_ms"] >= 0.0


def test_slot_protocol_errors(gpu):
    tri = scenes.eightballs_mesh(8, 16)
    vis = scenes.c5_vis()
    g = scenes.build_partitions(gpu, vis, {"mesh": tri}, 1)
    cam = scenes.c5_camera()
    with pytest.raises(gpu.GxyError):
        gpu.render_wait(g, 5)                      # never submitted
    gpu.render_submit(g, cam, vis["lighting"], 64, 64, 0.001, 1)
    with pytest.raises(gpu.GxyError):
        gpu.render_submit(g, cam, vis["lighting"], 64, 64, 0.001, 1)   # still in flight
    gpu.render_wait(g, 1)
    with pytest.raises(gpu.GxyError):
        gpu.render_wait(g, 1)                      # already waited
    with pytest.raises(gpu.GxyError):
        gpu.render_submit(g, cam, vis["lighting"], 64, 64, 0.001, gpu.max_slots())


def test_volume_frames_go_through_the_same_calls(gpu, oracle, golden_dir, provider):
    """A volume Visualization renders synchronously inside the submit; two slots still hold two different images until waited for."""
    st, ds = util.load_state(golden_dir, "nineBalls", provider)
    vis = st["visualizations"][0]
    g = scenes.build_partitions(gpu, vis, ds, 1)
    o = scenes.build_partitions(oracle, vis, ds, 1)
    for k in (0, 1):
        gpu.render_submit(g, st["cameras"][k], vis["lighting"], 128, 128, st["epsilon"], k)
    for k in (0, 1):
        gpu.render_wait(g, k)
        fb_g = g[0].download_rgba32f(128, 128)
        fb_o, _ = oracle.render(o, st["cameras"][k], vis["lighting"], 128, 128, st["epsilon"])
        assert util.fb_fraction(fb_g, fb_o) >= 0.999
