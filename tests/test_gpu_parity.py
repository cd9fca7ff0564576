"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle and the reference's golds."""
import os

import numpy as np
import pytest
from PIL import Image

from galaxy_b200 import scenes
from tests import util

pytestmark = pytest.mark.gpu

from tests.test_oracle_golds import GOLDS, check_gold_fraction  # 99.9 % for every gold; nineBalls_0/_1: documented deviation -> xfail


@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


@pytest.mark.parametrize("name,cam_idx,min_gold", GOLDS)
def test_state_files_match_oracle_and_golds(gpu, oracle, golden_dir, provider, name, cam_idx, min_gold):
    """Every reproducible reference test state: GPU image within 1/255 of the oracle image on
    >= 99.9 % of pixels and of the reference's gold on the same fraction the oracle reaches."""
    st, ds = util.load_state(golden_dir, name, provider)
    vis, cam = st["visualizations"][0], st["cameras"][cam_idx]
    g_parts = scenes.build_partitions(gpu, vis, ds, 1)
    o_parts = scenes.build_partitions(oracle, vis, ds, 1)
    fb_g, st_g = gpu.render(g_parts, cam, vis["lighting"], 512, 512, st["epsilon"])
    fb_o, st_o = oracle.render(o_parts, cam, vis["lighting"], 512, 512, st["epsilon"])
    for k in ("primary_rays", "shadow_rays", "ao_rays", "terminated_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    frac_fb = util.fb_fraction(fb_g, fb_o, 1.0 / 255)
    img_g = g_parts[0].download_rgba8(512, 512)
    img_o = oracle.fb_to_rgba8(fb_o)
    gold = np.asarray(Image.open(os.path.join(golden_dir, "golds", "%s_%05d.png" % (name, cam_idx))).convert("RGBA"))
    frac_img = util.image_fraction(img_g, img_o)
    frac_gold = util.image_fraction(img_g, gold)
    rel = np.abs(fb_g - fb_o).max() / max(1e-12, np.abs(fb_o).max())
    print(name, cam_idx, "fb<=1/255: %.6f img: %.6f gold: %.6f max rel fb err %.3g" % (frac_fb, frac_img, frac_gold, rel), st_g)
    assert frac_fb >= 0.999 and frac_img >= 0.999
    assert rel <= 1e-3  # volume-integrated colour/opacity within 1e-3 relative (north star)
    check_gold_fraction(name, cam_idx, frac_gold, min_gold)  # last: an expected failure must not hide the asserts above


@pytest.mark.parametrize("name", ["oneBall", "nineBalls", "xyz", "camera-shadow"])
def test_trace_raylist_matches_oracle(gpu, oracle, golden_dir, provider, name):
    """TraceRays::Trace on one RayList: generated rays bit-identical, traced columns and the
    spawned SECONDARY list compared column by column (term/type/x/y exact, floats <= 1e-5 abs)."""
    st, ds = util.load_state(golden_dir, name, provider)
    vis, cam = st["visualizations"][0], st["cameras"][0]
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    o = scenes.build_partitions(oracle, vis, ds, 1)[0]
    w = h = 160
    rg, ng = g.generate_rays(cam, w, h)
    ro, no = o.generate_rays(cam, w, h)
    assert ng == no and ng > 0
    assert np.array_equal(rg[:, :ng].view(np.int32)[[0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 18, 19, 20, 21, 22]],
                          ro[:, :no].view(np.int32)[[0, 1, 2, 3, 4, 5, 10, 11, 12, 13, 18, 19, 20, 21, 22]])
    L = oracle.resolve_lights(vis["lighting"], cam)
    Lg = gpu.resolve_lights(vis["lighting"], cam)
    assert L == Lg
    sg, nsg, _ = g.trace_raylist(L, rg, ng, st["epsilon"])
    so, nso, _ = o.trace_raylist(L, ro, no, st["epsilon"])
    assert nsg == nso
    assert np.array_equal(util.icol(rg, "term", ng), util.icol(ro, "term", no))
    exact = 0
    for c in ("r", "g", "b", "o", "t"):
        a, b = util.fcol(rg, c, ng), util.fcol(ro, c, no)
        assert np.allclose(a, b, rtol=1e-4, atol=1e-5), c
        exact += int(np.array_equal(a, b))
    hit = (util.icol(ro, "term", no) & 1) != 0
    for c in ("sr", "sg", "sb", "nx", "ny", "nz"):
        assert np.allclose(util.fcol(rg, c, ng)[hit], util.fcol(ro, c, no)[hit], rtol=1e-4, atol=1e-5), c
    print(name, "rays", ng, "secondary", nsg, "bit-exact float columns of 5:", exact)
    if nso:
        for c in ("x", "y", "type", "term"):
            assert np.array_equal(util.icol(sg, c, nsg), util.icol(so, c, nso)), c
        for c in ("ox", "oy", "oz", "dx", "dy", "dz", "r", "g", "b", "o", "t", "tMax"):
            assert np.allclose(util.fcol(sg, c, nsg), util.fcol(so, c, nso), rtol=1e-4, atol=1e-6, equal_nan=True), c
        # second wave: trace the secondaries, classify
        sg2, sg2n, _ = g.trace_raylist(L, sg, nsg, st["epsilon"])
        so2, so2n, _ = o.trace_raylist(L, so, nso, st["epsilon"])
        assert sg2n == so2n == 0
        mism = (util.icol(sg, "term", nsg) != util.icol(so, "term", nso)).mean()
        assert mism <= 1e-4, mism
        g.classify(sg, nsg)
        o.classify(so, nso)
        assert (util.icol(sg, "classification", nsg) != util.icol(so, "classification", nso)).mean() <= 1e-4


@pytest.mark.parametrize("n_tris,n_spheres,seed", [(1, 0, 1), (3, 2, 2), (500, 200, 3), (40000, 5000, 4), (300000, 0, 5)])
def test_nearest_hit_ids_bit_exact(gpu, oracle, n_tris, n_spheres, seed):
    """K2/K4: nearest-hit (geomID, primID) bit-exact against the oracle's brute-force-equivalent
    search; t/u/v bit-exact (same op order); the near-tie fraction is reported."""
    tri, par = util.random_soup(n_tris, n_spheres, seed)
    ds = {"tris": tri}
    if n_spheres:
        ds["parts"] = par
    vis = util.soup_vis(with_particles=n_spheres > 0)
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    o = scenes.build_partitions(oracle, vis, ds, 1)[0]
    n = 200000
    org, d = util.random_rays(n, seed + 100)
    tn, tf = np.zeros(n, np.float32), np.full(n, np.inf, np.float32)
    ig, tg = g.intersect(org, d, tn, tf)
    io, to = o.intersect(org, d, tn, tf)
    mism = (ig != io).any(1)
    print("prims", n_tris + n_spheres, "hit fraction %.3f" % (io[:, 0] >= 0).mean(), "id mismatches", int(mism.sum()), "info", g.build_info())
    # mismatching ids are only allowed on exact-t ties; with the deterministic tie rule there are none
    assert mism.sum() == 0
    assert np.array_equal(tg.view(np.int32), to.view(np.int32))


@pytest.mark.parametrize("nparts", [2, 4, 8])
def test_partitioned_volume_render_matches_oracle(gpu, oracle, golden_dir, provider, nparts):
    """Spatial partitions with ray forwarding (several partitions on one device exchange rays by
    device copies): same image as the oracle at the same partition count, and same ray counts."""
    st, ds = util.load_state(golden_dir, "nineBalls", provider)
    vis, cam = st["visualizations"][0], st["cameras"][1]
    g_parts = scenes.build_partitions(gpu, vis, ds, nparts)
    o_parts = scenes.build_partitions(oracle, vis, ds, nparts)
    fb_g, st_g = gpu.render(g_parts, cam, vis["lighting"], 384, 384, st["epsilon"])
    fb_o, st_o = oracle.render(o_parts, cam, vis["lighting"], 384, 384, st["epsilon"])
    print(nparts, st_g, st_o)
    for k in ("primary_rays", "shadow_rays", "forwarded_rays", "terminated_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    assert util.fb_fraction(fb_g, fb_o) >= 0.999
    fb_1, _ = oracle.render(scenes.build_partitions(oracle, vis, ds, 1), cam, vis["lighting"], 384, 384, st["epsilon"])
    print("vs 1 partition:", util.fb_fraction(fb_g, fb_1))


@pytest.mark.parametrize("nparts", [1, 2, 8])
def test_partitioned_geometry_render_matches_oracle(gpu, oracle, nparts):
    """Triangles + particles, shadows + AO, partitioned by the createPartitionDoc extents."""
    tri = scenes.eightballs_mesh(60, 120)
    _, par = util.random_soup(1, 3000, 9)
    ds = {"tris": tri, "parts": par}
    vis = util.soup_vis(True)
    cam = dict(eye=[3.0, 2.0, -4.0], dir=[-3.0, -2.0, 4.0], up=[0.0, 1.0, 0.0], aov=30.0)
    g_parts = scenes.build_partitions(gpu, vis, ds, nparts)
    o_parts = scenes.build_partitions(oracle, vis, ds, nparts)
    fb_g, st_g = gpu.render(g_parts, cam, vis["lighting"], 320, 240, 0.001)
    fb_o, st_o = oracle.render(o_parts, cam, vis["lighting"], 320, 240, 0.001)
    print(nparts, st_g, st_o)
    for k in ("primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    assert util.fb_fraction(fb_g, fb_o) >= 0.999


def test_edge_cases(gpu, oracle, golden_dir, provider):
    st, ds = util.load_state(golden_dir, "xyz", provider)
    vis, cam = st["visualizations"][0], st["cameras"][0]
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    L = gpu.resolve_lights(vis["lighting"], cam)
    # empty list
    rays = np.zeros((25, 16), np.float32)
    out, n, _ = g.trace_raylist(L, rays, 0)
    assert out is None and n == 0
    # rays that miss the box entirely: BOUNDARY/TIMEOUT semantics equal the oracle's
    o = scenes.build_partitions(oracle, vis, ds, 1)[0]
    rays[0:3, :4] = np.array([[5, 5, 5, 5], [5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    rays[3:6, :4] = np.array([[1, 0, 0, -1], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32)
    rays[19, :4] = np.float32(3.4e38)
    rays[22, :4] = np.array([1, 1, 2, 4], np.int32).view(np.float32)
    r2 = rays.copy()
    g.trace_raylist(L, rays, 4)
    o.trace_raylist(L, r2, 4)
    assert np.array_equal(util.icol(rays, "term", 4), util.icol(r2, "term", 4))
    assert np.array_equal(util.fcol(rays, "t", 4).view(np.int32), util.fcol(r2, "t", 4).view(np.int32))
    # camera inside the volume
    cam2 = dict(cam, eye=[0.1, 0.2, 0.3], dir=[0.0, 0.0, 1.0])
    fb_g, sg = gpu.render([g], cam2, vis["lighting"], 96, 64, st["epsilon"])
    fb_o, so = oracle.render([o], cam2, vis["lighting"], 96, 64, st["epsilon"])
    assert sg["primary_rays"] == so["primary_rays"]
    assert util.fb_fraction(fb_g, fb_o) >= 0.999
    # tone map parity incl. out-of-range values
    img = g.download_rgba8(96, 64)
    assert np.array_equal(img, oracle.fb_to_rgba8(fb_g))
