"""-m gpu: the interactive / asynchronous frame path on the CUDA path against the oracle's literal
frame-stamped accumulation.  (Green on the B200 since round 1's driver run; round 2 added the frames-in-flight form.)"""
import numpy as np
import pytest

from tests import util
from tests.test_progressive import CAM_A, CAM_B, H, W, scene


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 2])
def test_gpu_progressive_frames_match_oracle(gpu, oracle, nparts):
    g, vis = scene(gpu, nparts)
    o, _ = scene(oracle, nparts)
    L = vis["lighting"]
    r = oracle.ProgressiveRendering(W, H)
    gpu.progressive_reset(g[0])
    for frame, cam in ((0, CAM_A), (0, CAM_A), (1, CAM_B), (0, CAM_A), (3, CAM_A), (4, CAM_B)):
        fb_o, st_o = r.render(o, cam, L, frame)
        fb_g, st_g = gpu.render_progressive(g, cam, L, W, H, frame)
        frac = util.fb_fraction(fb_g, fb_o, 1.0 / 255)
        assert frac >= 0.999, (frame, frac)
        assert st_g["terminated_rays"] == st_o["terminated_rays"], (frame, st_g, st_o)
    # a new window size re-allocates: plain first frame again
    fb_g, _ = gpu.render_progressive(g, CAM_A, L, W // 2, H // 2, 0)
    fb_o, _ = oracle.render(o, CAM_A, L, W // 2, H // 2)
    assert util.fb_fraction(fb_g, fb_o, 1.0 / 255) >= 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 2])
def test_gpu_progressive_frames_in_flight(gpu, oracle, nparts):
    """frames 1, 2, 3 are submitted to three frame slots before the first wait (gxyviewer keeps rendering while frames are on their
    way) and waited for in the order 1, 3, 2: frame 2 arrives after frame 3 was merged, so it is stale and dropped -- the oracle,
    which applies the reference's rule per contribution (Rendering.cpp:104-153), does the same with that arrival order."""
    g, vis = scene(gpu, nparts)
    o, _ = scene(oracle, nparts)
    L = vis["lighting"]
    r = oracle.ProgressiveRendering(W, H)
    gpu.progressive_reset(g[0])
    frames = [(1, CAM_A, 0), (2, CAM_B, 1), (3, CAM_A, 2)]
    for frame, cam, slot in frames:
        gpu.render_progressive_submit(g, cam, L, W, H, frame, slot)
    for k, expect_merged in ((0, True), (2, True), (1, False)):
        frame, cam, slot = frames[k]
        fb_o, _ = r.render(o, cam, L, frame)
        fb_g, st_g, merged = gpu.render_progressive_wait(g, W, H, slot)
        assert merged == expect_merged, (frame, merged)
        assert util.fb_fraction(fb_g, fb_o, 1.0 / 255) >= 0.999, frame
    # the same frame number again adds (ACCUMULATE_PIXEL), also through a slot
    gpu.render_progressive_submit(g, CAM_A, L, W, H, 3, 0)
    fb_o, _ = r.render(o, CAM_A, L, 3)
    fb_g, _, merged = gpu.render_progressive_wait(g, W, H, 0)
    assert merged and util.fb_fraction(fb_g, fb_o, 1.0 / 255) >= 0.999


@pytest.mark.gpu
def test_gxywriter_samples_a_sampling_visualization(gpu, tmp_path):
    """gxywriter on a state whose Visualization holds only sampler operators: the .samples file holds the same set of points
    the Python binding collects (C++ gxy::Sampler on top of gxy_sample)."""
    import json
    import os
    import subprocess

    from galaxy_b200 import scenes
    from tests.test_gxywriter import EXE, write_vol
    tmp = str(tmp_path)
    vol = scenes.radial_volume("eightBalls", 48)
    write_vol(os.path.join(tmp, "radial-eightBalls.vol"), vol)
    doc = {"Datasets": [{"name": "v", "type": "Volume", "filename": "radial-eightBalls.vol"}],
           "Visualizations": [{"operators": [{"type": "IsoSampler", "dataset": "v", "isovalue": 0.25}]}],
           "Cameras": [{"aov": 35.0, "viewpoint": [2.0, 1.5, -3.0], "viewcenter": [0.0, 0.0, 0.0], "viewup": [0.0, 1.0, 0.0]}]}
    state = os.path.join(tmp, "sampling.state")
    json.dump(doc, open(state, "w"))
    for nparts in (1, 2):
        base = os.path.join(tmp, "s%d" % nparts)
        out = subprocess.run([EXE, "-s", "96", "64", "-P", str(nparts), "-o", base, state], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        got = np.fromfile(base + "_00000.samples", np.float32).reshape(-1, 3)
        st = scenes.parse_state(doc, base_dir=tmp)
        ds = scenes.load_datasets(st, scenes.default_data_provider(data_dir=tmp))
        parts = scenes.build_partitions(gpu, st["visualizations"][0], ds, nparts)
        want, _ = gpu.sample(parts, st["cameras"][0], 96, 64)
        want = np.concatenate(want)
        a, b = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
        assert len(a) > 0 and a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.gpu
def test_gpu_sampler_with_two_operators_matches_oracle(gpu, oracle):
    """two sampler operators on two volumes in one sampling Visualization (the first that fires ends the step, the others are
    not evaluated in it, SamplerTraceRays.ispc:195-204): same samples as the oracle at 1 and 2 partitions"""
    from galaxy_b200 import scenes
    from tests.test_sampler import CAM2, sorted_rows
    vis = dict(annotation="", lighting=scenes.parse_lighting(None),
               operators=[scenes.parse_operator({"type": "IsoSampler", "dataset": "a", "isovalue": 0.25}),
                          scenes.parse_operator({"type": "GradientSampler", "dataset": "b", "tolerance": 0.6})])
    ds = {"a": scenes.radial_volume("eightBalls", 48), "b": scenes.radial_volume("oneBall", 48)}
    for nparts in (1, 2):
        sg, st_g = gpu.sample(scenes.build_partitions(gpu, vis, ds, nparts), CAM2, 120, 90)
        so, st_o = oracle.sample(scenes.build_partitions(oracle, vis, ds, nparts), CAM2, 120, 90)
        assert st_g["traced_rays"] == st_o["traced_rays"] and st_g["forwarded_rays"] == st_o["forwarded_rays"]
        for r in range(nparts):
            a, b = sorted_rows(sg[r]), sorted_rows(so[r])
            assert len(b) > 0 and a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nparts, r)


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 2, 8])
def test_gpu_sampler_loop_mode_matches_oracle(gpu, oracle, nparts, monkeypatch):
    """GXY_SAMPLER_LOOP=1: a ray makes all its passes through a partition inside one launch (and the sample buffer grows by a
    second run when 4 samples per ray were not enough): the same sample sets and the same pass count as the oracle"""
    from galaxy_b200 import scenes
    from tests.test_sampler import CAM2, sampler_vis, sorted_rows
    monkeypatch.setenv("GXY_SAMPLER_LOOP", "1")
    for kind, param, name in (("IsoSampler", 0.25, "eightBalls"), ("GradientSampler", 0.9995, "oneBall")):   # the second: > 4 samples per ray
        vol = scenes.radial_volume(name, 48)
        vis = sampler_vis(kind, param)
        sg, st_g = gpu.sample(scenes.build_partitions(gpu, vis, {"v": vol}, nparts), CAM2, 96, 72)
        so, st_o = oracle.sample(scenes.build_partitions(oracle, vis, {"v": vol}, nparts), CAM2, 96, 72)
        for k in ("primary_rays", "traced_rays", "forwarded_rays"):
            assert st_g[k] == st_o[k], (kind, k, st_g, st_o)
        for r in range(nparts):
            a, b = sorted_rows(sg[r]), sorted_rows(so[r])
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (kind, r)
        assert st_g["waves"] < st_o["traced_rays"]


@pytest.mark.gpu
def test_data_driven_state_renders_unchanged(gpu, oracle, tmp_path):
    """the reference's tests/data-driven.state VERBATIM on the stand-in datasets of tools/make_data_driven.py: gxywriter's PNG at 1 and
    2 partitions, the Python binding and the oracle agree; and the picture resembles the reference's gold"""
    import json
    import os
    import shutil
    import subprocess

    from PIL import Image

    from galaxy_b200 import scenes
    from tests.test_gxywriter import EXE
    from tools.make_data_driven import make_datasets, write
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    state_src = os.path.join(root, "tests", "golden", "states", "data-driven.state")
    st = scenes.parse_state(json.load(open(state_src)))
    vis, cam = st["visualizations"][0], st["cameras"][0]
    lines, parts, mesh = make_datasets()
    write(str(tmp_path / "p1"), 1, n=96)
    # the volume as the .vol file describes it (its header rounds origin and spacing to 6 digits, scripts/vti2vol:70-80)
    ds = {"volume": scenes.load_vol_file(str(tmp_path / "p1" / "radial-eightBalls.vol")), "pathlines": lines, "particles": parts, "tmesh": mesh}
    o = scenes.build_partitions(oracle, vis, ds, 1)
    fb_o, st_o = oracle.render(o, cam, vis["lighting"], 256, 256, st["epsilon"])
    g = scenes.build_partitions(gpu, vis, ds, 1)
    fb_g, st_g = gpu.render(g, cam, vis["lighting"], 256, 256, st["epsilon"])
    for k in ("primary_rays", "shadow_rays", "terminated_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    assert util.fb_fraction(fb_g, fb_o, 1.0 / 255) >= 0.999
    img_g = g[0].download_rgba8(256, 256)
    for nparts in (1, 2):
        tmp = str(tmp_path / ("p%d" % nparts))
        if nparts > 1:
            write(tmp, nparts, n=96)
        shutil.copy(state_src, tmp)
        out = subprocess.run([EXE, "-s", "256", "256", "-P", str(nparts), "-o", os.path.join(tmp, "dd"), os.path.join(tmp, "data-driven.state")],
                             capture_output=True, text=True, timeout=180)
        assert out.returncode == 0, out.stderr
        png = np.asarray(Image.open(os.path.join(tmp, "dd_00000.png")).convert("RGBA"))
        assert util.image_fraction(png, img_g, 1) >= (1.0 if nparts == 1 else 0.97)     # 2 partitions: cut stream lines end in doubled end points, bricks restart the march


@pytest.mark.gpu
def test_gpu_data_driven_gold(gpu, golden_dir):
    """the 12th gold on the CUDA path: tests/data-driven.state unchanged, datasets regenerated by tools/make_data_driven.py,
    512 x 512: within 1/255 of the reference's gold on >= 99.9 % of the pixels (the oracle reaches 99.99 %)"""
    import json
    import os

    from galaxy_b200 import scenes
    from tests.test_oracle_golds import data_driven_datasets, gold_fraction
    st = scenes.parse_state(json.load(open(os.path.join(golden_dir, "states", "data-driven.state"))))
    vis, cam = st["visualizations"][0], st["cameras"][0]
    g = scenes.build_partitions(gpu, vis, data_driven_datasets(), 1)
    gpu.render(g, cam, vis["lighting"], 512, 512, st["epsilon"])
    frac = gold_fraction(golden_dir, g[0].download_rgba8(512, 512), "data-driven", 0)
    print("data-driven on the GPU: fraction within 1/255 of the gold:", frac)
    assert frac >= 0.999
