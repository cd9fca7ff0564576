"""Nearest-hit ids against the reference's own Embree TRAVERSAL (not only its triangle arithmetic): the vendored Embree 3.6.1 of
/root/reference built whole by oracle/embree.mk (binned-SAH BVH8 with Triangle4 leaves, rtcIntersect8 packets = what ISPC's
rtcIntersectV runs, ospray/common/Model.ih:54-70; scene set-up as Model.cpp:49-107 / TriangleMesh.cpp:129-136).

  * the oracle vs COMMITTED Embree answers (tests/golden/embree_traversal_fixtures.npz) -- runs everywhere;
  * the oracle vs the live library, where oracle/_ref/libgxy_embree_scene_ref.so exists (it travels to the GPU box);
  * -m gpu: the CUDA traversal (gxy_intersect through the C ABI) vs the live library or the committed answers.
north star: primitive ids bit-exact except on near-tie rays, tie fraction reported.  A near tie = the two candidates' t differ by
<= 2 ulp: Embree's own winner there depends on its rcp+Newton rounding and BVH order; ours is the lowest id."""
import importlib.util
import os

import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util
from tests.test_oracle_embree import ulps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_embree_traversal_fixtures", os.path.join(ROOT, "tests", "golden", "make_embree_traversal_fixtures.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
FX = np.load(os.path.join(ROOT, "tests", "golden", "embree_traversal_fixtures.npz"))


def check_against_embree(prim_e, tuv_e, ids_o, tuv_o, what, t_ulps=4):
    """hit/miss decisions identical; primID identical except on near ties (the two winners' t within t_ulps ulp); t within t_ulps ulp
    (IEEE divide here, rcp + one Newton step in Embree: up to 3 ulp seen over ~10^5 hits), u/v within 2^-21 absolute."""
    hit_e, hit_o = prim_e >= 0, ids_o[:, 1] >= 0
    assert np.array_equal(hit_e, hit_o), (what, "hit/miss decisions differ on %d rays" % int((hit_e != hit_o).sum()))
    same = prim_e == ids_o[:, 1]
    diff = hit_e & ~same
    if diff.any():
        assert ulps(tuv_e[diff, 0], tuv_o[diff, 0]).max(initial=0) <= t_ulps, (what, "primID differs away from a tie")
    agree = hit_e & same
    assert ulps(tuv_e[agree, 0], tuv_o[agree, 0]).max(initial=0) <= t_ulps, what
    assert np.abs(tuv_e[agree, 1:] - tuv_o[agree, 1:]).max(initial=0) <= 2.0 ** -21, what
    exact_t = float((tuv_e[agree, 0].view(np.int32) == tuv_o[agree, 0].view(np.int32)).mean()) if agree.any() else 1.0
    return int(diff.sum()), int(hit_e.sum()), exact_t


def _oracle_scene(tri):
    from oracle import oracle
    return scenes.build_partitions(oracle, util.soup_vis(False), {"tris": tri}, 1)[0]


@pytest.mark.parametrize("k", range(len(gen.CASES)))
def test_oracle_ids_match_committed_embree_traversal(k):
    tri, org, d, tn, tf = gen.case(*gen.CASES[k])
    ids_o, tuv_o = _oracle_scene(tri).intersect(org, d, tn, tf)
    nd, nh, ex = check_against_embree(FX["prim%d" % k], FX["tuv%d" % k], ids_o, tuv_o, "fixture %d" % k)
    print(gen.CASES[k], "hits", nh, "near-tie id differences", nd, "tie fraction %.2e" % (nd / max(1, nh)), "bit-exact t %.3f" % ex)
    assert nh > 1000 and nd <= max(2, nh // 500)
    # Embree's packet and single-ray traversals gave the same ids when the fixture was made
    assert np.array_equal(FX["prim%d" % k], FX["prim1_%d" % k])


def _live():
    from oracle import embree_scene
    if not embree_scene.available():
        pytest.skip("oracle/_ref/libgxy_embree_scene_ref.so not built (needs /root/reference: make -C oracle -f embree.mk embree)")
    return embree_scene


def test_oracle_ids_match_live_embree_traversal_on_the_headline_mesh():
    """the bumpy eight-sphere mesh at 1/16 of the headline tessellation per axis (390 k triangles), camera-like and AO-like rays"""
    es_mod = _live()
    tri = scenes.eightballs_mesh(scenes.C5_FULL[0] // 16, scenes.C5_FULL[1] // 16)
    n = 60000
    org, d = util.random_rays(n, 31)
    tn = np.full(n, 0.001, np.float32)
    tf = np.where(np.arange(n) % 2 == 0, np.float32(50.0), np.float32(0.3)).astype(np.float32)
    es = es_mod.EmbreeScene(tri.verts, tri.indices)
    prim_e, tuv_e, _ = es.intersect(org, d, tn, tf, packet=8)
    ids_o, tuv_o = _oracle_scene(tri).intersect(org, d, tn, tf)
    nd, nh, ex = check_against_embree(prim_e, tuv_e, ids_o, tuv_o, "live c5/16")
    print("triangles", len(tri.indices), "hits", nh, "near-tie id differences", nd, "tie fraction %.2e" % (nd / max(1, nh)), "Embree build %.2f s" % es.build_seconds)
    assert nh > 5000 and nd <= max(2, nh // 500)


def test_embree_scene_is_deterministic_and_restores_the_fp_state():
    """two commits of the same mesh answer alike; the caller's MXCSR (flush-to-zero / denormals) is what it was before"""
    es_mod = _live()
    tri, org, d, tn, tf = gen.case(*gen.CASES[0])
    a = es_mod.EmbreeScene(tri.verts, tri.indices).intersect(org, d, tn, tf, threads=1)
    b = es_mod.EmbreeScene(tri.verts, tri.indices, threads=2).intersect(org, d, tn, tf, threads=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.int32), b[1].view(np.int32))
    assert np.array_equal(a[0], FX["prim0"])
    tiny = np.float32(1e-40)
    assert tiny * np.float32(1.0) != 0.0   # denormals still alive in this thread


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(gen.CASES)))
def test_gpu_ids_match_committed_embree_traversal(k):
    from galaxy_b200 import gpu
    tri, org, d, tn, tf = gen.case(*gen.CASES[k])
    g = scenes.build_partitions(gpu, util.soup_vis(False), {"tris": tri}, 1)[0]
    ids_g, tuv_g = g.intersect(org, d, tn, tf)
    nd, nh, ex = check_against_embree(FX["prim%d" % k], FX["tuv%d" % k], ids_g, tuv_g, "gpu fixture %d" % k)
    print(gen.CASES[k], "hits", nh, "near-tie id differences", nd, "tie fraction %.2e" % (nd / max(1, nh)))
    assert nh > 1000 and nd <= max(2, nh // 500)


@pytest.mark.gpu
def test_gpu_ids_match_live_embree_traversal_on_a_large_mesh():
    """1/4 tessellation per axis of the headline scene (6.25 M triangles), 400 k rays: the device LBVH traversal and Embree's SAH BVH8
    traversal must name the same triangle for every ray (ids bit-exact away from near ties)"""
    from galaxy_b200 import gpu
    es_mod = _live()
    tri = scenes.eightballs_mesh(scenes.C5_FULL[0] // 4, scenes.C5_FULL[1] // 4)
    n = 400000
    org, d = util.random_rays(n, 57)
    tn = np.full(n, 0.001, np.float32)
    tf = np.where(np.arange(n) % 2 == 0, np.float32(50.0), np.float32(0.3)).astype(np.float32)
    es = es_mod.EmbreeScene(tri.verts, tri.indices)
    prim_e, tuv_e, _ = es.intersect(org, d, tn, tf, packet=8)
    g = scenes.build_partitions(gpu, util.soup_vis(False), {"tris": tri}, 1)[0]
    ids_g, tuv_g = g.intersect(org, d, tn, tf)
    nd, nh, ex = check_against_embree(prim_e, tuv_e, ids_g, tuv_g, "gpu live c5/4")
    print("triangles", len(tri.indices), "hits", nh, "near-tie id differences", nd, "tie fraction %.2e" % (nd / max(1, nh)), "bit-exact t %.3f" % ex)
    assert nh > 50000 and nd <= max(2, nh // 500)
