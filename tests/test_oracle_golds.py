"""Pins the CPU oracle against the reference's own golden vectors: the gold PNGs of
tests/image-gold-tests.sh.  CPU only."""
import json
import os

import numpy as np
import pytest
from PIL import Image

from galaxy_b200 import scenes
from oracle import oracle

# (state, camera index) -> minimum fraction of pixels within 1/255 of the gold, per channel.
# 9 of 11 reproduce the gold on >= 99.9 % of pixels with the reference's CURRENT source order.
# nineBalls_0/_1 reach 99.78 / 99.82 %: their golds were produced by a build that integrated the
# DVR interval before searching it for an iso crossing (proved by test_ninballs_gold_provenance).
GOLDS = [("infinite", 0, 0.999), ("infinite-shadow", 0, 0.999), ("absolute", 0, 0.999), ("absolute-shadow", 0, 0.999),
         ("camera", 0, 0.999), ("camera-shadow", 0, 0.999), ("xyz", 0, 0.999), ("oneBall", 0, 0.999),
         ("nineBalls", 0, 0.997), ("nineBalls", 1, 0.998), ("nineBalls", 2, 0.999)]


def render_oracle(golden_dir, provider, name, cam_idx, size=512):
    st = scenes.parse_state(json.load(open(os.path.join(golden_dir, "states", name + ".state"))))
    ds = scenes.load_datasets(st, provider)
    vis, cam = st["visualizations"][0], st["cameras"][cam_idx]
    parts = scenes.build_partitions(oracle, vis, ds, 1)
    fb, stats = oracle.render(parts, cam, vis["lighting"], size, size, st["epsilon"])
    return oracle.fb_to_rgba8(fb), stats


def gold_fraction(golden_dir, img, name, cam_idx):
    gold = np.asarray(Image.open(os.path.join(golden_dir, "golds", "%s_%05d.png" % (name, cam_idx))).convert("RGBA"))
    assert gold.shape == img.shape
    diff = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(-1)
    return float((diff <= 1).mean())


@pytest.mark.parametrize("name,cam_idx,min_frac", GOLDS)
def test_oracle_matches_gold(golden_dir, provider, name, cam_idx, min_frac):
    img, stats = render_oracle(golden_dir, provider, name, cam_idx)
    frac = gold_fraction(golden_dir, img, name, cam_idx)
    print(name, cam_idx, "fraction within 1/255:", frac, stats)
    assert frac >= min_frac
    assert (img[..., 3] == 255).all()
    assert stats["orphan_pixels"] == 0


def test_nineballs_gold_provenance(golden_dir, provider):
    """With the DVR interval integrated BEFORE the iso search (diagnostic switch), all three
    nineBalls golds are reproduced on >= 99.99 % of pixels: evidence that those golds predate the
    reference's current TraceRays.ispc:483-506 ordering."""
    oracle.lib().gxo_set_option(b"dvr_before_iso", 1)
    try:
        for cam_idx in range(3):
            img, _ = render_oracle(golden_dir, provider, "nineBalls", cam_idx)
            assert gold_fraction(golden_dir, img, "nineBalls", cam_idx) >= 0.9999
    finally:
        oracle.lib().gxo_set_option(b"dvr_before_iso", 0)


def test_data_driven_state_resembles_its_gold(golden_dir):
    """The 12th gold, tests/data-driven.state (volume slices + DVR, path lines, particles, a mesh): its datasets come out of VTK
    (contour filter, stream tracer) in the reference and cannot be regenerated here; tools/make_data_driven.py writes closed-form
    stand-ins (helical stream lines of the same field from the same seeds, spheres for the isosurfaces).  The oracle's render of the
    UNCHANGED state file on them agrees with the reference's gold within 1/255 on > 92 % of the pixels (the particles sit at other
    points of their spheres; the stream-line tubes, slices, DVR and mesh caps coincide).  Not a pin -- the inputs differ -- but it
    shows the PathLines radius / colour rules and the four-operator mix against the reference's own picture."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from tools.make_data_driven import make_datasets
    st = scenes.parse_state(json.load(open(os.path.join(golden_dir, "states", "data-driven.state"))))
    lines, parts, mesh = make_datasets()
    ds = {"volume": scenes.radial_volume("eightBalls", 256), "pathlines": lines, "particles": parts, "tmesh": mesh}
    vis, cam = st["visualizations"][0], st["cameras"][0]
    o = scenes.build_partitions(oracle, vis, ds, 1)
    fb, stats = oracle.render(o, cam, vis["lighting"], 512, 512, st["epsilon"])
    img = oracle.fb_to_rgba8(fb)
    frac = gold_fraction(golden_dir, img, "data-driven", 0)
    # ... and outside the footprint of the eight particle balls and their shadows (8 spheres of radius 0.345 in place of the points,
    # against no particles at all) the agreement is at the level of the other golds: the stream-line tubes are the reference's
    import copy
    cs = np.array([[sx, sy, sz] for sz in (-.5, .5) for sy in (-.5, .5) for sx in (-.5, .5)], np.float32)
    vb = copy.deepcopy(vis)
    vb["operators"][2].update(radius0=0.345, radius1=0.345, value0=0.0, value1=0.0)
    vn = copy.deepcopy(vis)
    del vn["operators"][2]
    fb_b, _ = oracle.render(scenes.build_partitions(oracle, vb, dict(ds, particles=scenes.ParticlesDataset(cs, np.ones(8, np.float32))), 1),
                            cam, vis["lighting"], 512, 512, st["epsilon"])
    ds_n = {k: v for k, v in ds.items() if k != "particles"}
    fb_n, _ = oracle.render(scenes.build_partitions(oracle, vn, ds_n, 1), cam, vis["lighting"], 512, 512, st["epsilon"])
    foot = np.abs(oracle.fb_to_rgba8(fb_b)[..., :3].astype(int) - oracle.fb_to_rgba8(fb_n)[..., :3].astype(int)).max(-1) > 0
    gold = np.asarray(Image.open(os.path.join(golden_dir, "golds", "data-driven_00000.png")).convert("RGBA"))
    d = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(-1)
    outside = float((d[~foot] <= 1).mean())
    print("data-driven stand-in: fraction within 1/255 of the gold: %.4f; outside the particle balls (%.1f %% of the image): %.4f" % (
        frac, 100 * (1 - foot.mean()), outside), stats)
    assert frac >= 0.92 and foot.mean() < 0.12 and outside >= 0.997
