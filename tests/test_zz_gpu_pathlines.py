"""-m gpu: PathLines (round Bezier curves) on the CUDA path, through the C ABI, against the oracle.
(Named test_zz_* so that it runs after the established parity files.)"""
import json
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from galaxy_b200 import scenes
from tests import util
from tests.test_curve_host import CAM, pathlines_scene, pathlines_vis
from tests.test_oracle_curves import build, rays_at
from tests.test_gxywriter_geometry import EXE, stage_pathlines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def mixed_scene():
    """the operator mix of tests/data-driven.state: a volume with slices + DVR, path lines, particles, a mesh"""
    vol = scenes.radial_volume("eightBalls", 64)
    tri, par = util.random_soup(300, 150, 9)
    tri.verts *= 0.6; par.centers *= 0.6
    ds = {"volume": vol, "lines": pathlines_scene(31, 8), "parts": par, "tris": tri}
    vis = dict(annotation="", lighting=dict(lights=[[-1.0, -2.0, -4.0]], types=[2], n_ao=0, ao_radius=1.0, shadows=True, Ka=0.5, Kd=0.5),
               operators=[
                   dict(type="VolumeVis", dataset="volume", colormap=[[0.0, 1.0, 0.5, 0.5], [0.4, 0.5, 0.5, 1.0], [0.9, 1.0, 0.5, 1.0]],
                        opacitymap=[[0.0, 0.2], [0.3, 0.03], [0.31, 0.0], [1.0, 0.0]], data_range=None, isovalues=[],
                        slices=[[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]], volume_render=True),
                   pathlines_vis()["operators"][0],
                   dict(type="ParticlesVis", dataset="parts", colormap=[[0.0, 0.2, 1.0, 0.2], [1.0, 1.0, 1.0, 0.2]], opacitymap=[[0, 1], [1, 1]],
                        data_range=None, radius0=0.01, radius1=0.04, value0=0.0, value1=1.0),
                   dict(type="TrianglesVis", dataset="tris", colormap=[[0.0, 1.0, 0.2, 0.2], [1.6, 0.2, 0.2, 1.0]], opacitymap=[[0, 1], [1, 1]],
                        data_range=None)])
    return ds, vis


def test_curve_nearest_hit_bit_exact(gpu, oracle):
    """gxy_intersect on a PathLines scene: (geomID, primID) and (t, u) bit-identical to the oracle's search (the device
    source compiled for the CPU already is, tests/test_curve_host.py; this is the same code through the wide BVH)."""
    ds, vis = {"lines": pathlines_scene()}, pathlines_vis()
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    o = scenes.build_partitions(oracle, vis, ds, 1)[0]
    v, d, c = ds["lines"].to_arrays()
    cp = build(oracle.lib(), v, d, c, 0.01, 0.05, 0.0, 1.2)
    org, dr, tn, tf = rays_at(cp, 100000, 5, 0.03)
    ig, tg = g.intersect(org, dr, tn, tf)
    io, to = o.intersect(org, dr, tn, tf)
    hit = io[:, 0] >= 0
    print("segments", len(cp), "hit fraction %.3f" % hit.mean(), "id mismatches", int((ig != io).any(1).sum()), g.build_info())
    assert hit.mean() > 0.2
    assert np.array_equal(ig, io)
    assert np.array_equal(tg[hit, :2].view(np.uint32), to[hit, :2].view(np.uint32))
    # bounded interval
    tn2, tf2 = np.full_like(tn, 2.5), np.full_like(tf, 3.5)
    ig, tg = g.intersect(org, dr, tn2, tf2)
    io, to = o.intersect(org, dr, tn2, tf2)
    assert np.array_equal(ig, io) and np.array_equal(tg[io[:, 0] >= 0, 0], to[io[:, 0] >= 0, 0])


@pytest.mark.parametrize("nparts,n_ao", [(1, 0), (1, 4), (2, 2)])
def test_pathlines_render_matches_oracle(gpu, oracle, nparts, n_ao):
    ds, vis = {"lines": pathlines_scene()}, pathlines_vis(n_ao=n_ao)
    g = scenes.build_partitions(gpu, vis, ds, nparts)
    o = scenes.build_partitions(oracle, vis, ds, nparts)
    fb_g, st_g = gpu.render(g, CAM, vis["lighting"], 256, 192, 0.001)
    fb_o, st_o = oracle.render(o, CAM, vis["lighting"], 256, 192, 0.001)
    for k in ("primary_rays", "shadow_rays", "ao_rays", "terminated_rays", "forwarded_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    frac = util.fb_fraction(fb_g, fb_o, 1.0 / 255)
    print("pathlines", nparts, n_ao, "fraction %.6f" % frac, st_g)
    assert (fb_o[..., :3].max(-1) > 0).mean() > 0.03
    assert frac >= 0.999


def test_pathlines_trace_raylist_matches_oracle(gpu, oracle):
    ds, vis = {"lines": pathlines_scene()}, pathlines_vis(n_ao=2)
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    o = scenes.build_partitions(oracle, vis, ds, 1)[0]
    rg, ng = g.generate_rays(CAM, 160, 120)
    ro, no = o.generate_rays(CAM, 160, 120)
    assert ng == no and ng > 0
    L = oracle.resolve_lights(vis["lighting"], CAM)
    sg, nsg, hg = g.trace_raylist(L, rg, ng, 0.001, want_hits=True)
    so, nso, ho = o.trace_raylist(L, ro, no, 0.001, want_hits=True)
    assert nsg == nso and nso > 0
    assert np.array_equal(hg, ho)
    assert np.array_equal(util.icol(rg, "term", ng), util.icol(ro, "term", no))
    hit = (util.icol(ro, "term", no) & 1) != 0
    for c in ("r", "g", "b", "o", "t"):
        assert np.allclose(util.fcol(rg, c, ng), util.fcol(ro, c, no), rtol=1e-4, atol=1e-5), c
    for c in ("sr", "sg", "sb", "nx", "ny", "nz"):
        assert np.allclose(util.fcol(rg, c, ng)[hit], util.fcol(ro, c, no)[hit], rtol=1e-4, atol=1e-5), c
    for c in ("x", "y", "type", "term"):
        assert np.array_equal(util.icol(sg, c, nsg), util.icol(so, c, nso)), c


@pytest.mark.parametrize("nparts", [1, 2])
def test_data_driven_operator_mix_matches_oracle(gpu, oracle, nparts):
    ds, vis = mixed_scene()
    cam = dict(eye=[3.0, 4.0, 3.0], dir=[-3.0, -4.0, -3.0], up=[0.0, 0.0, 1.0], aov=35.0)     # tests/data-driven.state
    g = scenes.build_partitions(gpu, vis, ds, nparts)
    o = scenes.build_partitions(oracle, vis, ds, nparts)
    fb_g, st_g = gpu.render(g, cam, vis["lighting"], 256, 256, 0.001)
    fb_o, st_o = oracle.render(o, cam, vis["lighting"], 256, 256, 0.001)
    for k in ("primary_rays", "shadow_rays", "terminated_rays", "forwarded_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    frac = util.fb_fraction(fb_g, fb_o, 1.0 / 255)
    print("operator mix", nparts, "fraction %.6f" % frac, st_g)
    assert frac >= 0.999


def test_pathlines_api_errors(gpu):
    ctx = gpu.Context.default(0)
    L = gpu.lib()
    import ctypes as C
    v = np.zeros((3, 3), np.float32)
    bad = np.array([2], np.int32)
    h = C.c_void_p()
    rc = L.gxy_pathlines_create(ctx.h, 3, v.ctypes.data_as(C.POINTER(C.c_float)), None, 1, bad.ctypes.data_as(C.POINTER(C.c_int)), C.byref(h))
    assert rc != 0 and b"segment 0" in L.gxy_last_error()


def test_pathlines_state_renders_like_the_python_binding(gpu, tmp_path):
    """gxywriter on a PathLines state (1 and 2 partitions): the PNG equals the image of the Python binding on the same data;
    a PathLines dataset under another Vis type is refused with the reference's kind of message."""
    for nparts in (1, 2):
        tmp = str(tmp_path / ("p%d" % nparts))
        os.makedirs(tmp)
        path, state, pieces = stage_pathlines(tmp, nparts)
        out = subprocess.run([EXE, "-s", "256", "192", "-P", str(nparts), "-o", os.path.join(tmp, "img"), path], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        png = np.asarray(Image.open(os.path.join(tmp, "img_00000.png")).convert("RGBA"))
        st = scenes.parse_state(state)
        vis, cam = st["visualizations"][0], st["cameras"][0]
        if nparts == 1:
            parts = scenes.build_partitions(gpu, vis, {"pathlines": pieces[0]}, 1)
            gpu.render(parts, cam, vis["lighting"], 256, 192, st["epsilon"])
            assert util.image_fraction(png, parts[0].download_rgba8(256, 192), 0) == 1.0
        assert (png[..., :3].max(-1) > 0).mean() > 0.03
    path, state, _ = stage_pathlines(str(tmp_path), 1)
    state["Visualizations"][0]["operators"][0]["type"] = "Particles"
    json.dump(state, open(path, "w"))
    out = subprocess.run([EXE, "-s", "32", "32", "-o", str(tmp_path / "x"), path], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "on a PathLines dataset" in out.stderr
