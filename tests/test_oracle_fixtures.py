"""The oracle against COMMITTED answers of the reference's own compiled code (tests/golden/embree_box_fixtures.npz, made by
tests/golden/make_embree_box_fixtures.py in the build container from oracle/_ref): Embree's packet triangle intersector and
src/data/Box.cpp.  Runs everywhere, also where /root/reference and oracle/_ref do not exist (the live comparisons in
tests/test_oracle_embree.py / test_oracle_box.py are skipped there)."""
import importlib.util
import os

import numpy as np

from galaxy_b200 import scenes
from oracle import oracle
from tests import test_oracle_box as tb
from tests import test_oracle_embree as te
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_embree_box_fixtures", os.path.join(ROOT, "tests", "golden", "make_embree_box_fixtures.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
FX = np.load(os.path.join(ROOT, "tests", "golden", "embree_box_fixtures.npz"))


def test_triangle_hits_match_committed_embree_answers():
    for k, (kind, n_tris, seed, n_rays) in enumerate(gen.TRI_CASES):
        tri, org, d, tn, tf = gen.tri_case(kind, n_tris, seed, n_rays)
        o = scenes.build_partitions(oracle, util.soup_vis(False), {"tris": tri}, 1)[0]
        ids_o, tuv_o = o.intersect(org, d, tn, tf)
        nd, nh, _ = te.check_against_embree(FX["tri_prim%d" % k], FX["tri_tuv%d" % k], ids_o, tuv_o, "fixture %s %d" % (kind, n_tris))
        assert nh > 0 and nd <= max(2, nh // 500)


def test_box_answers_match_committed_reference_answers():
    boxes, rays = tb.cases(gen.BOX_N, gen.BOX_SEED)
    olib = oracle.lib()
    f = np.zeros(gen.BOX_N, np.int32)
    olib.gxo_exit_face(gen.BOX_N, tb._f(boxes), tb._f(rays), tb._i(f))
    assert np.array_equal(f, FX["box_face"].astype(np.int32))
    h, t = np.zeros(gen.BOX_N, np.int32), np.zeros((gen.BOX_N, 2), np.float32)
    with np.errstate(all="ignore"):
        olib.gxo_box_intersect(gen.BOX_N, tb._f(boxes), tb._f(rays), tb._i(h), tb._f(t))
    assert np.array_equal(h, FX["box_hit"].astype(np.int32))
    hit = h == 1
    assert np.array_equal(t[hit].view(np.int32), FX["box_t"][hit].view(np.int32))
