"""The Sampler path (src/sampler): rays leave sample points where a sampler operator fires.
CPU part: the oracle's restatement against closed-form properties (the reference has no golden data for this path and its
kernels are ISPC: parity of this row is pinned by these properties only).  GPU part: the CUDA kernel against the oracle."""
import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util

CAM = dict(eye=[0.0, 0.0, -4.0], dir=[0.0, 0.0, 4.0], up=[0.0, 1.0, 0.0], aov=30.0)
CAM2 = dict(eye=[2.0, 1.5, -3.0], dir=[-2.0, -1.5, 3.0], up=[0.0, 1.0, 0.0], aov=35.0)


def sampler_vis(kind, param):
    key = "tolerance" if kind == "GradientSampler" else "isovalue"
    return dict(annotation="", lighting=scenes.parse_lighting(None), operators=[scenes.parse_operator({"type": kind, "dataset": "v", key: param})])


def sorted_rows(a):
    a = np.ascontiguousarray(a, np.float32)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


def test_operator_keys():
    g = scenes.parse_operator({"type": "GradientSampler", "dataset": "scalar", "tolerance": 0.1, "volume rendering": False})   # examples/noise.state
    assert g["type"] == "GradientSamplerVis" and g["tolerance"] == float(np.float32(0.1))
    i = scenes.parse_operator({"type": "IsoSampler", "dataset": "s"})
    assert i["type"] == "IsoSamplerVis" and i["isovalue"] == 0.0


def test_iso_sampler_samples_lie_on_the_isosurface(oracle):
    """oneBall is f = |p|: every sample is 0.001 (in t) behind a crossing of |p| = 0.6, two crossings per ray through the ball."""
    vol = scenes.radial_volume("oneBall", 64)
    vis = sampler_vis("IsoSampler", 0.6)
    parts = scenes.build_partitions(oracle, vis, {"v": vol}, 1)
    samp, st = oracle.sample(parts, CAM, 96, 96)
    p = samp[0]
    r = np.linalg.norm(p, axis=1)
    assert len(p) > 1000 and np.all(np.abs(r - 0.6) < 2.5e-3)
    # rays through the ball: |pixel ray passes within 0.6 of the origin| -> 2 samples each
    rays, n = parts[0].generate_rays(CAM, 96, 96)
    o = rays[0:3, :n].T.astype(np.float64); d = rays[3:6, :n].T.astype(np.float64)
    dist = np.linalg.norm(np.cross(o, d), axis=1) / np.linalg.norm(d, axis=1)
    through = int((dist < 0.6 - 0.02).sum())
    assert 2 * through <= len(p) <= 2 * int((dist < 0.6 + 0.02).sum())
    assert st["primary_rays"] == n and st["traced_rays"] >= n + len(p)       # every sample costs one more pass (KEEP_HERE)
    # front and back crossings: z < 0 and z > 0 in equal numbers
    assert abs(int((p[:, 2] < 0).sum()) - int((p[:, 2] > 0).sum())) <= 0.05 * len(p)   # perspective: grazing rays cross twice in front


def test_sample_raylist_sets_t_and_term(oracle):
    vol = scenes.radial_volume("oneBall", 64)
    part = scenes.build_partitions(oracle, sampler_vis("IsoSampler", 0.6), {"v": vol}, 1)[0]
    rays, n = part.generate_rays(CAM, 64, 64)
    before = rays.copy()
    part.sample_raylist(rays, n)
    term = util.icol(rays, "term", n).copy()
    assert set(np.unique(term)) <= {1, 4} and (term == 1).any() and (term == 4).any()       # RAY_SURFACE / RAY_BOUNDARY
    changed = [c for c in range(25) if not np.array_equal(rays[c, :n].view(np.int32), before[c, :n].view(np.int32))]
    assert set(changed) <= {util.CI["t"], util.CI["term"]}
    t = util.fcol(rays, "t", n).copy()
    hitp = rays[0:3, :n].T + t[:, None] * rays[3:6, :n].T
    assert np.all(np.abs(np.linalg.norm(hitp[term == 1], axis=1) - 0.6) < 2.5e-3)
    # a second pass continues behind the sample: the same rays now find the far crossing
    part.sample_raylist(rays, n)
    t2 = util.fcol(rays, "t", n)
    again = (term == 1) & (util.icol(rays, "term", n) == 1)
    assert again.sum() > 0.9 * (term == 1).sum() and np.all(t2[again] > t[again])


def test_partitions_collect_the_same_samples_up_to_the_brick_seams(oracle):
    vol = scenes.radial_volume("eightBalls", 64)
    vis = sampler_vis("IsoSampler", 0.25)
    one, _ = oracle.sample(scenes.build_partitions(oracle, vis, {"v": vol}, 1), CAM2, 80, 60)
    for nparts in (2, 8):
        parts = scenes.build_partitions(oracle, vis, {"v": vol}, nparts)
        many, st = oracle.sample(parts, CAM2, 80, 60)
        assert st["forwarded_rays"] > 0
        # every sample lies in the box of the partition that collected it (forwarded rays restart at the seam)
        ext = [scenes.volume_boxes(vol, scenes.partition(scenes.factor(nparts), vol.counts)[r]) for r in range(nparts)]
        for r in range(nparts):
            lo, hi = ext[r][2], ext[r][3]
            assert np.all(many[r] >= lo - 2e-3) and np.all(many[r] <= hi + 2e-3)
        total = sum(len(m) for m in many)
        assert abs(total - len(one[0])) <= 0.03 * len(one[0])          # crossings inside the step that straddles a seam are lost / doubled


def test_gradient_sampler_fires_where_the_gradient_turns(oracle):
    """f = |p|: the gradient is the unit radial vector; dot(g(t), g(t - step)) drops below the tolerance only near the centre,
    where the direction turns quickly.  tolerance 0.5 -> the samples sit within a few steps of the closest approach."""
    vol = scenes.radial_volume("oneBall", 64)
    vis = sampler_vis("GradientSampler", 0.5)
    parts = scenes.build_partitions(oracle, vis, {"v": vol}, 1)
    samp, st = oracle.sample(parts, CAM, 64, 64)
    p = samp[0]
    assert 0 < len(p) < st["primary_rays"]
    assert np.all(np.linalg.norm(p, axis=1) < 0.12)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    assert g.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return g


@pytest.mark.gpu
@pytest.mark.parametrize("kind,param,name", [("IsoSampler", 0.6, "oneBall"), ("IsoSampler", 0.25, "eightBalls"), ("GradientSampler", 0.9, "eightBalls")])
def test_gpu_sample_raylist_matches_oracle(gpu, oracle, kind, param, name):
    vol = scenes.radial_volume(name, 64)
    vis = sampler_vis(kind, param)
    g = scenes.build_partitions(gpu, vis, {"v": vol}, 1)[0]
    o = scenes.build_partitions(oracle, vis, {"v": vol}, 1)[0]
    rg, ng = g.generate_rays(CAM2, 160, 120)
    ro, no = o.generate_rays(CAM2, 160, 120)
    assert ng == no and ng > 0
    for _ in range(3):                       # three passes: every ray continues behind its sample
        g.sample_raylist(rg, ng)
        o.sample_raylist(ro, no)
        assert np.array_equal(util.icol(rg, "term", ng), util.icol(ro, "term", no))
        assert np.array_equal(util.fcol(rg, "t", ng).view(np.uint32), util.fcol(ro, "t", no).view(np.uint32))      # bit-exact
    assert (util.icol(ro, "term", no) == 1).any()


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 2, 8])
def test_gpu_sampler_frame_matches_oracle(gpu, oracle, nparts):
    vol = scenes.radial_volume("eightBalls", 64)
    vis = sampler_vis("IsoSampler", 0.25)
    g = scenes.build_partitions(gpu, vis, {"v": vol}, nparts)
    o = scenes.build_partitions(oracle, vis, {"v": vol}, nparts)
    sg, st_g = gpu.sample(g, CAM2, 160, 120)
    so, st_o = oracle.sample(o, CAM2, 160, 120)
    for k in ("primary_rays", "traced_rays", "forwarded_rays"):
        assert st_g[k] == st_o[k], (k, st_g, st_o)
    for r in range(nparts):                  # the same SET of samples per partition (order is unspecified in the reference too)
        a, b = sorted_rows(sg[r]), sorted_rows(so[r])
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), r
    print("sampler", nparts, "samples", sum(len(s) for s in sg), st_g)


@pytest.mark.gpu
def test_gpu_samples_render_as_particles_without_leaving_the_device(gpu, oracle):
    vol = scenes.radial_volume("oneBall", 64)
    vis = sampler_vis("IsoSampler", 0.6)
    g = scenes.build_partitions(gpu, vis, {"v": vol}, 1)
    sg, _ = gpu.sample(g, CAM, 64, 64)
    colors, opac = scenes.resample_tf([[0.0, 1.0, 0.5, 0.2], [1.0, 1.0, 0.5, 0.2]], [[0, 1], [1, 1]])
    lighting = dict(lights=[[1.0, 2.0, -3.0]], types=[2], n_ao=0, ao_radius=1.0, shadows=False, Ka=0.4, Kd=0.6)
    sc = gpu.Scene()
    sc.set_partition([-1, -1, -1], [1, 1, 1], [-1, -1, -1], [1, 1, 1], [-1] * 6)
    sc.add_particles_from_samples(g[0], 0.01, 0.0, 0.0, 0.0, colors, opac, 0.0, 1.0)
    sc.commit()
    fb_g, _ = gpu.render([sc], CAM, lighting, 128, 128, 0.001)
    so = oracle.Scene()
    so.set_partition([-1, -1, -1], [1, 1, 1], [-1, -1, -1], [1, 1, 1], [-1] * 6)
    so.add_particles_vis(sg[0], np.zeros(len(sg[0]), np.float32), 0.01, 0.0, 0.0, 0.0, colors, opac, 0.0, 1.0)
    so.commit()
    fb_o, _ = oracle.render([so], CAM, lighting, 128, 128, 0.001)
    assert (fb_o[..., :3].max(-1) > 0).mean() > 0.05
    assert util.fb_fraction(fb_g, fb_o, 1.0 / 255) >= 0.999
