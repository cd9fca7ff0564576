"""-m gpu: configurations the reference supports that its own test states do not exercise -- every one through the C ABI
against the oracle on the same inputs: uchar volumes, an orthographic camera (aov = 0), the three light types together,
more volume operators than the specialised kernels cover (the catch-all instantiation), volumes + triangles + particles in
one Visualization, a partition without any primitive, and the tiled / row-major primary-ray orders of the frame path."""
import os

import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util

pytestmark = pytest.mark.gpu

CMAP = [[0.0, 1.0, 0.5, 0.5], [0.25, 0.5, 1.0, 0.5], [0.5, 0.5, 0.5, 1.0], [0.75, 1.0, 1.0, 0.5], [1.0, 1.0, 0.5, 1.0]]
OMAP = [[0.0, 0.05], [0.2, 0.02], [0.21, 0.0], [1.0, 0.0]]
CAM = dict(eye=[1.0, 3.0, -3.0], dir=[-1.0, -3.0, 3.0], up=[0.0, 1.0, 0.0], aov=30.0, annotation="")


@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def volvis(dataset, iso=(), slices=(), dvr=False, rng=None, omap=None):
    return dict(type="VolumeVis", dataset=dataset, colormap=CMAP, opacitymap=omap or [[0.0, 1.0], [1.0, 1.0]], data_range=rng,
                slices=[list(s) for s in slices], isovalues=list(iso), volume_render=dvr)


def both(gpu, oracle, vis, ds, cam, w, h, nparts=1, eps=0.001, min_frac=0.999):
    g = scenes.build_partitions(gpu, vis, ds, nparts)
    o = scenes.build_partitions(oracle, vis, ds, nparts)
    fb_g, st_g = gpu.render(g, cam, vis["lighting"], w, h, eps)
    fb_o, st_o = oracle.render(o, cam, vis["lighting"], w, h, eps)
    for k in ("primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    frac = util.fb_fraction(fb_g, fb_o)
    rel = np.abs(fb_g - fb_o).max() / max(1e-12, np.abs(fb_o).max())
    print("fraction within 1/255: %.6f, max rel err %.3g" % (frac, rel), st_g)
    assert frac >= min_frac
    return fb_g, st_g


def test_uchar_volume(gpu, oracle):
    """type "uchar" volumes (Volume.cpp:207; SSV_sample_uint8): values 0..255 sampled as float."""
    v = scenes.radial_volume("eightBalls", 96)
    q = np.clip(v.data * 160.0, 0, 255).astype(np.uint8)
    ds = {"v": scenes.VolumeDataset(v.origin, v.counts, v.deltas, q)}
    L = scenes.parse_lighting({"Sources": [[1, 1, -2, 0]], "shadows": True, "Ka": 0.4, "Kd": 0.6})
    vis = dict(annotation="", lighting=L, operators=[volvis("v", iso=[40.0], dvr=True, rng=[0.0, 255.0], omap=[[0.0, 0.02], [0.3, 0.0], [1.0, 0.0]])])
    both(gpu, oracle, vis, ds, CAM, 200, 160)
    both(gpu, oracle, vis, ds, CAM, 200, 160, nparts=4)


def test_orthographic_camera(gpu, oracle):
    """aov = 0 selects the orthographic branch of Camera::generate_initial_rays (Camera.cpp:548-556)."""
    ds = {"v": scenes.radial_volume("oneBall", 64)}
    L = scenes.parse_lighting({"Sources": [[0, 0, -4, 0]], "shadows": False})
    vis = dict(annotation="", lighting=L, operators=[volvis("v", iso=[0.5], slices=[[0, 1, 0, 0.1]])])
    cam = dict(eye=[0.5, 0.4, -3.0], dir=[-0.1, -0.1, 1.0], up=[0.0, 1.0, 0.0], aov=0.0, annotation="")
    fb, st = both(gpu, oracle, vis, ds, cam, 160, 128)
    assert st["primary_rays"] > 0


def test_three_light_types_with_shadows_and_ao(gpu, oracle):
    """directional (0), camera-relative (1) and point (2) sources together (Lighting.cpp:59-115, Rendering.cpp:157-216)."""
    ds = {"v": scenes.radial_volume("eightBalls", 80)}
    L = scenes.parse_lighting({"Sources": [[1, 1, -2, 0], [0.5, 0.5, 0.0, 1], [2.0, 2.0, -2.0, 2]], "shadows": True, "Ka": 0.3, "Kd": 0.7,
                               "ao count": 5, "ao radius": 0.4})
    vis = dict(annotation="", lighting=L, operators=[volvis("v", iso=[0.3])])
    fb, st = both(gpu, oracle, vis, ds, CAM, 192, 144)
    assert st["shadow_rays"] == 3 * (st["ao_rays"] // 5) and st["ao_rays"] > 0


def test_more_volume_operators_than_specialised_kernels(gpu, oracle):
    """4 VolumeVis operators: the catch-all kernel instantiation (GXY_MAX_VOLUME_VIS) with fewer operators than its width."""
    ds = {"a": scenes.radial_volume("oneBall", 48), "b": scenes.radial_volume("eightBalls", 48), "x": scenes.radial_volume("xramp", 48),
          "y": scenes.radial_volume("yramp", 48)}
    L = scenes.parse_lighting({"Sources": [[1, 2, -3, 0]], "shadows": True})
    ops = [volvis("a", iso=[0.8]), volvis("b", dvr=True, omap=OMAP), volvis("x", slices=[[1, 0, 0, 0.2]]), volvis("y", iso=[-0.5, 0.6])]
    vis = dict(annotation="", lighting=L, operators=ops)
    both(gpu, oracle, vis, ds, CAM, 160, 128)
    both(gpu, oracle, vis, ds, CAM, 160, 128, nparts=2)


def test_volume_and_geometry_in_one_visualization(gpu, oracle):
    """VolumeVis + TrianglesVis + ParticlesVis: slices/isosurfaces/DVR clipped by nearest surface hits (TraceRays.ispc:420-470)."""
    tri, par = util.random_soup(400, 150, 21)
    ds = {"v": scenes.radial_volume("oneBall", 64), "tris": tri, "parts": par}
    geo = util.soup_vis(True)
    L = scenes.parse_lighting({"Sources": [[1, 2, -3, 0]], "shadows": True, "Ka": 0.3, "Kd": 0.7, "ao count": 2, "ao radius": 0.5})
    vis = dict(annotation="", lighting=L, operators=[volvis("v", iso=[0.9], dvr=True, omap=OMAP)] + geo["operators"])
    both(gpu, oracle, vis, ds, CAM, 192, 144)


def test_partition_without_primitives(gpu, oracle):
    """8 partitions, all geometry in one octant: seven partitions hold no primitive at all (no tree to walk) and only forward."""
    tri = scenes.eightballs_mesh_single(20, 40)
    verts = (tri["pts"] + np.array([0.5, 0.5, 0.5])).astype(np.float32)
    data = np.linalg.norm(verts, axis=1).astype(np.float32)
    ds = {"tris": scenes.TrianglesDataset(verts, tri["nrm"], data, tri["tris"])}
    vis = util.soup_vis(False)
    cam = dict(eye=[-3.0, -2.0, -4.0], dir=[3.0, 2.0, 4.0], up=[0.0, 1.0, 0.0], aov=30.0, annotation="")
    fb, st = both(gpu, oracle, vis, ds, cam, 200, 150, nparts=8)
    assert st["forwarded_rays"] > 0 and st["shadow_rays"] > 0


@pytest.mark.parametrize("tiled", ["0", "1"])
def test_primary_ray_order_does_not_change_the_image(gpu, oracle, tiled, golden_dir, provider):
    """GXY_TILED: the frame path's 16x8-tile generation order against the reference's row-major order, at a window size that
    is not a multiple of the tile (edge tiles are partly outside)."""
    st, ds = util.load_state(golden_dir, "nineBalls", provider)
    vis, cam = st["visualizations"][0], st["cameras"][2]
    old = os.environ.get("GXY_TILED")
    os.environ["GXY_TILED"] = tiled
    try:
        both(gpu, oracle, vis, ds, cam, 251, 173)
    finally:
        if old is None:
            del os.environ["GXY_TILED"]
        else:
            os.environ["GXY_TILED"] = old


def test_async_download_equals_blocking_download(gpu, oracle, golden_dir, provider):
    """gxy_frame_download_rgba8_async + _wait: two images in flight while later frames render, each equal to the blocking
    download of the same frame."""
    st, ds = util.load_state(golden_dir, "xyz", provider)
    vis = st["visualizations"][0]
    g = scenes.build_partitions(gpu, vis, ds, 1)[0]
    cams = [dict(st["cameras"][0]), dict(st["cameras"][0], eye=[1.5, 2.0, -3.0], dir=[-1.5, -2.0, 3.0]), dict(st["cameras"][0], eye=[-2.0, 1.0, -3.0], dir=[2.0, -1.0, 3.0])]
    w, h = 200, 120
    want = []
    for c in cams:
        gpu.render_device([g], c, vis["lighting"], w, h, st["epsilon"])
        want.append(g.download_rgba8(w, h).copy())
    bufs = [gpu.pinned_array((h, w, 4), np.uint8) for _ in cams]
    for c, b in zip(cams, bufs):
        gpu.render_device([g], c, vis["lighting"], w, h, st["epsilon"])
        g.download_rgba8_async(b)  # not waited for: the next render overlaps the copy
    g.download_wait()
    for a, b in zip(want, bufs):
        assert np.array_equal(a, b)
    assert not np.array_equal(want[0], want[1])


@pytest.mark.parametrize("bands,streams", [("1", "1"), ("3", "3"), ("4", "2"), ("7", "7")])
def test_band_pipeline_renders_the_same_frame(gpu, oracle, bands, streams):
    """GXY_BANDS / GXY_BAND_STREAMS: the single-GPU fused frame split into interleaved bands of tile rows on several streams
    (gxy_render) against the oracle, at a window whose tile rows do not divide evenly among the bands."""
    tri = scenes.eightballs_mesh(40, 80)
    ds = {"tris": tri}
    vis = util.soup_vis(False)
    cam = dict(eye=[3.0, 2.0, -4.0], dir=[-3.0, -2.0, 4.0], up=[0.0, 1.0, 0.0], aov=30.0, annotation="")
    old = {k: os.environ.get(k) for k in ("GXY_BANDS", "GXY_BAND_STREAMS")}
    os.environ["GXY_BANDS"], os.environ["GXY_BAND_STREAMS"] = bands, streams
    try:
        fb, st = both(gpu, oracle, vis, ds, cam, 333, 217)
        assert st["ao_rays"] > 0 and st["shadow_rays"] > 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# min_staged: share of ALL samples of the frame that must come from staged boxes (oneBall: 32 AO rays per hit dominate and only
# primaries are staged; xyz: slices only, nothing to march)
@pytest.mark.parametrize("name,cam_idx,min_staged", [("oneBall", 0, 0.01), ("xyz", 0, 0.0), ("camera-shadow", 0, 0.3)])
def test_tma_staged_march_matches_oracle(gpu, oracle, golden_dir, provider, name, cam_idx, min_staged):
    """GXY_MARCH_TMA=1: the primaries of a one-volume scene go through the TMA-staged march (gxy_march_tma.cu: 3-D boxes of
    voxels fetched by cp.async.bulk.tensor into shared memory one stage ahead, global gathers for whatever a box does not
    cover).  Same image and ray counts as the oracle, and the staged boxes must actually serve samples."""
    st, ds = util.load_state(golden_dir, name, provider)
    vis, cam = st["visualizations"][0], st["cameras"][cam_idx]
    old = os.environ.get("GXY_MARCH_TMA")
    os.environ["GXY_MARCH_TMA"] = "1"
    try:
        fb, sg = both(gpu, oracle, vis, ds, cam, 400, 300, eps=st["epsilon"])
        print("staged fraction", sg["staged_samples"] / max(1, sg["volume_samples"]))
        assert sg["staged_samples"] >= min_staged * sg["volume_samples"]
    finally:
        if old is None:
            del os.environ["GXY_MARCH_TMA"]
        else:
            os.environ["GXY_MARCH_TMA"] = old
