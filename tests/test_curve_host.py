"""CPU checks of the product's PathLines code where no GPU is needed:
 * galaxy_b200/csrc/gxy_curve.cuh -- the DEVICE source of the ray/round-Bezier test -- is compiled as plain C++
   (tests/curve_host_harness.cpp, g++ -ffp-contract=off) and must agree BIT FOR BIT with the oracle's independent
   restatement (which tests/test_oracle_curves.py pins against Embree's own intersector);
 * gxy_build_curves (host code of the library, DataDrivenPathLines::finalize) against the oracle's, bit for bit;
 * the state-file front end: PathLinesVis keys and defaults, PathLines::load_from_vtkPointSet's vertex duplication,
   the poly-line partitioning of scripts/partitionVTUs.vpy;
 * an oracle render of a PathLines scene (tube colours = transfer function of the radius mapped back to data)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from galaxy_b200 import gpu, scenes
from oracle import oracle
from tests.test_oracle_curves import _p, build, helices, intersect, rays_at

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_curve(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("curve_host") / "libcurve_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-mavx2", "-mfma", "-Wall", "-shared",
                           "-o", so, os.path.join(ROOT, "tests", "curve_host_harness.cpp")])
    return C.CDLL(so)


@pytest.mark.parametrize("seed,radii", [(7, (0.002, 0.06, 0.0, 1.7)), (8, (0.03, 0.03, 0.0, 0.0)), (9, (0.05, 0.01, 0.2, 1.0)),
                                       (10, (-1.0, 1.0, 0.0, 1.0))])
def test_device_source_matches_oracle_bit_for_bit(host_curve, seed, radii):
    O = oracle.lib()
    v, d, c = helices(seed)
    cp = build(O, v, d, c, *radii)
    org, dr, tn, tf = rays_at(cp, 20000, seed + 100, 0.03)
    for per_curve in (0, 1):
        for tn_, tf_ in ((tn, tf), (np.full_like(tn, 2.5), np.full_like(tf, 3.5))):
            a = intersect(O, "gxo_curve_intersect", cp, org, dr, tn_, tf_, per_curve)
            b = intersect(host_curve, "gxc_curve_intersect", cp, org, dr, tn_, tf_, per_curve)
            assert (a[0] >= (1 if per_curve else 0)).sum() > 100
            assert np.array_equal(a[0], b[0])
            assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))      # t, u
            assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))      # Ng


def test_cull_is_conservative_on_grazing_rays(host_curve):
    """The device path first asks curve_precull (line-to-line distance against the bound stored with the BVH record) and only
    then runs the sub-division.  Rays that graze the tubes -- through surface points, perpendicular to the normal, shifted by
    -1e-3 .. +1e-3 of the radius scale along it -- must get the oracle's answer bit for bit; the cull must also do its job."""
    O = oracle.lib()
    v, d, c = helices(12, nlines=8)
    cp = build(O, v, d, c, 0.004, 0.05, 0.0, 1.7)
    org, dr, tn, tf = rays_at(cp, 20000, 77, 0.03)
    prim, tu, ng = intersect(O, "gxo_curve_intersect", cp, org, dr, tn, tf, 0)
    hit = prim >= 0
    P = org[hit] + tu[hit, :1] * dr[hit]
    N = ng[hit] / np.linalg.norm(ng[hit], axis=1, keepdims=True)
    rng = np.random.default_rng(3)
    T = np.cross(N, rng.normal(size=N.shape))
    T /= np.linalg.norm(T, axis=1, keepdims=True)
    orgs, dirs = [], []
    for delta in (-1e-3, -1e-4, -1e-5, 0.0, 1e-5, 1e-4, 1e-3):
        orgs.append(P + N * (delta * 0.05) - T * rng.uniform(0.5, 3.0, (len(P), 1)))
        dirs.append(T * rng.uniform(0.5, 2.0, (len(P), 1)))
    org2 = np.ascontiguousarray(np.concatenate(orgs), np.float32)
    dr2 = np.ascontiguousarray(np.concatenate(dirs), np.float32)
    tn2, tf2 = np.zeros(len(org2), np.float32), np.full(len(org2), 3.4e38, np.float32)
    c0, t0 = C.c_long(), C.c_long()
    host_curve.gxc_cull_stats(C.byref(c0), C.byref(t0))
    a = intersect(O, "gxo_curve_intersect", cp, org2, dr2, tn2, tf2, 1)
    b = intersect(host_curve, "gxc_curve_intersect", cp, org2, dr2, tn2, tf2, 1)
    c1, t1 = C.c_long(), C.c_long()
    host_curve.gxc_cull_stats(C.byref(c1), C.byref(t1))
    assert 0.2 < a[0].reshape(len(org2), -1).any(1).mean() < 0.9         # really grazing: a good part hits, a good part misses
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    culled = (c1.value - c0.value) / (t1.value - t0.value)
    print("cull fraction on all ray x segment pairs: %.3f" % culled)
    assert culled > 0.8                                                 # all-pairs: nearly every pair is a miss the cull sees


def test_device_source_with_cull_against_embree_itself(host_curve):
    """the device source (cull + test) against the REFERENCE'S OWN compiled intersector (oracle/_ref, skipped where absent):
    the decisions agree but for a few grazing pairs per 10 000 hits (Embree's rcp/rsqrt are approximations), and no pair is lost to
    the cull: whatever Embree hits and the device source misses, the oracle (no cull) misses too"""
    so = os.path.join(ROOT, "oracle", "_ref", "libgxy_embree_curve_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libgxy_embree_curve_ref.so not built (needs /root/reference: make -C oracle ref)")
    R, O = C.CDLL(so), oracle.lib()
    hits = lost = extra = lost_to_cull = 0
    for seed in range(200, 206):
        rng = np.random.default_rng(seed)
        radii = (float(10 ** rng.uniform(-3, -1)), float(10 ** rng.uniform(-3, -1)), 0.0, float(rng.uniform(0.5, 2)))
        v, d, c = helices(seed, nlines=int(rng.integers(2, 10)))
        cp = build(O, v, d, c, *radii)
        org, dr, tn, tf = rays_at(cp, 5000, seed + 7, float(10 ** rng.uniform(-3, -1)))
        a = intersect(R, "gxr_curve_intersect", cp, org, dr, tn, tf, 1)
        b = intersect(host_curve, "gxc_curve_intersect", cp, org, dr, tn, tf, 1)
        o = intersect(O, "gxo_curve_intersect", cp, org, dr, tn, tf, 1)
        hits += int(a[0].sum()); lost += int(((a[0] == 1) & (b[0] == 0)).sum()); extra += int(((a[0] == 0) & (b[0] == 1)).sum())
        lost_to_cull += int(((a[0] == 1) & (b[0] == 0) & (o[0] == 1)).sum())
    assert hits > 10000 and lost + extra <= max(2, hits // 2000) and lost_to_cull == 0, (hits, lost, extra, lost_to_cull)


def test_device_source_degenerate_inputs(host_curve):
    """zero-length segments, zero radius, rays along the axis, zero direction components: same answers, no hangs."""
    O = oracle.lib()
    cp = np.array([
        [[0, 0, 0, .1], [0, 0, 0, .1], [0, 0, 0, .1], [0, 0, 0, .1]],                     # a point
        [[0, 0, 0, 0], [1 / 3, 0, 0, 0], [2 / 3, 0, 0, 0], [1, 0, 0, 0]],                 # zero radius
        [[0, 0, 1, .05], [0, 0, 1, .05], [1, 0, 1, .05], [1, 0, 1, .05]],                 # doubled ends (a Galaxy end segment)
        [[0, 1, 0, .2], [.3, 1, 0, .01], [.6, 1, 0, .2], [1, 1, 0, .01]],                 # strongly varying radius
    ], np.float32)
    org = np.array([[-2, 0, 0], [.5, 0, -1], [.5, 0, 3], [-1, 1, 0], [.5, 1, -2], [0.5, 0.5, 0.5]], np.float32)
    d = np.array([[1, 0, 0], [0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 0, 1], [0, 0, 0]], np.float32)
    tn, tf = np.zeros(len(org), np.float32), np.full(len(org), 3.4e38, np.float32)
    a = intersect(O, "gxo_curve_intersect", cp, org, d, tn, tf, 1)
    b = intersect(host_curve, "gxc_curve_intersect", cp, org, d, tn, tf, 1)
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


def test_library_curve_builder_matches_oracle_bit_for_bit():
    """gxy_build_curves is host code of libgxy_b200.so (the reference builds the curves on the host too): no device needed."""
    L, O = gpu.lib(), oracle.lib()
    for seed, radii in ((3, (0.002, 0.02, 0.0, 1.7)), (4, (0.05, 0.05, 1.0, 1.0)), (5, (-1.0, 1.0, 0.0, 1.0)), (6, (0.04, 0.01, 0.3, 0.9))):
        v, d, c = helices(seed, nlines=9)
        want = build(O, v, d, c, *radii)
        got = np.zeros_like(want)
        assert L.gxy_build_curves(len(v), _p(v), _p(d), len(c), _p(c), *radii, _p(got)) == 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        # data == NULL means 0 everywhere
        z = np.zeros_like(d)
        want0 = build(O, v, z, c, *radii)
        assert L.gxy_build_curves(len(v), _p(v), None, len(c), _p(c), *radii, _p(got)) == 0
        assert np.array_equal(got.view(np.uint32), want0.view(np.uint32))
    bad = np.array([len(v) - 1], np.int32)
    assert L.gxy_build_curves(len(v), _p(v), _p(d), 1, _p(bad), 0.1, 0.1, 0.0, 0.0, _p(got)) != 0
    assert b"segment 0" in L.gxy_last_error()


def test_pathlines_vis_keys_and_vertex_duplication():
    op = scenes.parse_operator({"dataset": "pathlines", "type": "PathLinesVis", "colormap": [[0.0, 0.0, 1.0, 0.0], [2.0, 1.0, 0.0, 1.0]],
                                "radius0": 0.002, "radius1": 0.02, "value0": 0.0, "value1": 1.7})       # tests/data-driven.state
    assert op["type"] == "PathLinesVis" and op["radius0"] == float(np.float32(0.002)) and op["value1"] == float(np.float32(1.7))
    dflt = scenes.parse_operator({"dataset": "p", "type": "PathLines"})                                  # "Vis" is appended
    assert dflt["type"] == "PathLinesVis" and (dflt["radius0"], dflt["radius1"], dflt["value0"], dflt["value1"]) == (-1.0, 1.0, 0.0, 1.0)
    pts = np.arange(15, dtype=np.float32).reshape(5, 3)
    ds = scenes.PathLinesDataset(pts, np.arange(5), [[0, 1, 2], [2, 3], [4, 0, 1, 3]])
    v, d, c = ds.to_arrays()
    assert len(v) == 9 and d.tolist() == [0, 1, 2, 2, 3, 4, 0, 1, 3]
    assert c.tolist() == [0, 1, 3, 5, 6, 7]                       # k - cells segments, first vertex of each


def test_polyline_partitioning_follows_the_reference_script():
    # one line crossing x = 0 twice: inside runs get one extra vertex at either end (partitionVTUs.vpy:116-147)
    x = np.array([-0.9, -0.6, -0.3, -0.15, 0.2, 0.5, 0.8, 0.4, -0.12, -0.5], np.float32)
    pts = np.stack([x, np.zeros_like(x), np.zeros_like(x)], 1)
    ds = scenes.PathLinesDataset(pts, np.arange(len(x)), [list(range(len(x)))])
    left = scenes.clip_pathlines(ds, [-1, 0, -1, 1, -1, 1], ghost=0.1)        # x <= 0.1
    assert [len(l) for l in left.lines] == [5, 3]                              # 0..3 (+4), then (7+) 8,9
    assert left.points[:, 0].tolist() == x[[0, 1, 2, 3, 4, 7, 8, 9]].tolist()
    right = scenes.clip_pathlines(ds, [0, 1, -1, 1, -1, 1], ghost=0.1)        # x >= -0.1
    assert [len(l) for l in right.lines] == [6]                                # (3+) 4..7 (+8)
    assert right.data.tolist() == [3, 4, 5, 6, 7, 8]
    none = scenes.clip_pathlines(ds, [5, 6, 5, 6, 5, 6])
    assert none.lines == [] and len(none.points) == 0


def pathlines_scene(seed=21, nlines=12):
    rng = np.random.default_rng(seed)
    pts, data, lines, k = [], [], [], 0
    for _ in range(nlines):
        n = int(rng.integers(4, 14))
        t = np.linspace(0, 3 + rng.uniform(0, 2), n)
        c = rng.uniform(-.4, .4, 3)
        p = np.stack([c[0] + 0.45 * np.cos(t), c[1] + 0.45 * np.sin(t), c[2] + 0.15 * t - 0.3], 1)
        pts.append(p); data.append(np.linalg.norm(p, axis=1)); lines.append(list(range(k, k + n))); k += n
    return scenes.PathLinesDataset(np.concatenate(pts), np.concatenate(data), lines)


def pathlines_vis(shadows=True, n_ao=0):
    return dict(annotation="", lighting=dict(lights=[[1.0, 2.0, -3.0]], types=[2], n_ao=n_ao, ao_radius=0.5, shadows=shadows, Ka=0.4, Kd=0.6),
                operators=[dict(type="PathLinesVis", dataset="lines", colormap=[[0.0, 0.0, 1.0, 0.0], [1.2, 1.0, 0.0, 1.0]],
                                opacitymap=[[0, 1], [1, 1]], data_range=None, radius0=0.01, radius1=0.05, value0=0.0, value1=1.2)])


CAM = dict(eye=[1.5, 1.0, -3.0], dir=[-1.5, -1.0, 3.0], up=[0.0, 1.0, 0.0], aov=35.0)


def test_oracle_renders_pathlines_scene():
    ds = pathlines_scene()
    vis = pathlines_vis()
    sc = scenes.build_partitions(oracle, vis, {"lines": ds}, 1)
    fb, st = oracle.render(sc, CAM, vis["lighting"], 192, 128, 0.001)
    lit = fb[..., :3].max(-1) > 0
    assert 0.03 < lit.mean() < 0.6                      # tubes cover part of the image
    assert st["shadow_rays"] > 0
    # hit colours lie on the colormap's line blue->red (green = 1 - red... here g ramps 1 -> 0, r 0 -> 1), scaled by the lighting
    rgb = fb[..., :3][lit]
    s = rgb[:, 0] + rgb[:, 1]
    assert np.all(np.abs(rgb[:, 2] / np.maximum(s, 1e-6)) <= 1.0 + 1e-3)
    # nearest-hit ids through the oracle's BVH equal a brute-force scan over the built curves
    v, d, c = ds.to_arrays()
    cp = build(oracle.lib(), v, d, c, 0.01, 0.05, 0.0, 1.2)
    org, dr, tn, tf = rays_at(cp, 4000, 5, 0.03)
    brute = intersect(oracle.lib(), "gxo_curve_intersect", cp, org, dr, tn, tf, 0)
    gp, tuv = sc[0].intersect(org, dr, tn, tf)
    assert np.array_equal(gp[:, 1], brute[0]) and np.array_equal(tuv[:, 0][brute[0] >= 0], brute[1][:, 0][brute[0] >= 0])
    # 2 partitions: runs, forwards rays, and the image stays close (cut lines end in doubled end points)
    sc2 = scenes.build_partitions(oracle, vis, {"lines": ds}, 2)
    fb2, st2 = oracle.render(sc2, CAM, vis["lighting"], 192, 128, 0.001)
    assert st2["forwarded_rays"] > 0
    assert float((np.abs(fb2[..., :3] - fb[..., :3]).max(-1) <= 2.0 / 255).mean()) > 0.97
