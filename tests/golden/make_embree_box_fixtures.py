"""Generates tests/golden/embree_box_fixtures.npz from the REFERENCE'S OWN compiled code (oracle/_ref, `make -C oracle ref`):
 * Embree 3.6.1's MoellerTrumboreIntersectorK<4,8> nearest hits (libgxy_embree_ref.so) for seeded triangle soups and the
   bumpy-sphere mesh of the headline workload;
 * src/data/Box.cpp exit_face / intersect answers (libgxy_box_ref.so) for seeded boxes and rays incl. the edge cases.
Only the answers are stored: the inputs are regenerated from the seeds by the same functions the live tests use.  Keeps the two
pins alive where oracle/_ref is absent (tests/test_oracle_fixtures.py).

  python tests/golden/make_embree_box_fixtures.py     (in the build container)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from galaxy_b200 import scenes  # noqa: E402
from tests import test_oracle_box as tb  # noqa: E402
from tests import test_oracle_embree as te  # noqa: E402
from tests import util  # noqa: E402

TRI_CASES = [("soup", 7, 2, 4000), ("soup", 500, 3, 4000), ("soup", 20000, 4, 3000), ("c5mesh", 24 * 48, 77, 4000)]
BOX_N, BOX_SEED = 20000, 7


def tri_case(kind, n_tris, seed, n_rays):
    tri = scenes.eightballs_mesh(24, 48) if kind == "c5mesh" else util.random_soup(n_tris, 0, seed)[0]
    org, d = util.random_rays(n_rays, seed + 50)
    tn = np.full(n_rays, 0.001 if kind == "c5mesh" else 0.0, np.float32)
    tf = np.full(n_rays, 50.0 if kind == "c5mesh" else np.inf, np.float32)
    return tri, org, d, tn, tf


def main():
    L = C.CDLL(te.SO)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.gxr_intersect_packet8.argtypes = [C.c_int, fp, ip, C.c_int, fp, fp, fp, fp, ip, fp, fp, C.c_int]
    out = {}
    for k, (kind, n_tris, seed, n_rays) in enumerate(TRI_CASES):
        tri, org, d, tn, tf = tri_case(kind, n_tris, seed, n_rays)
        prim, tuv, _ = te.embree_packet(L, tri, org, d, tn, tf)
        out["tri_prim%d" % k], out["tri_tuv%d" % k] = prim, tuv
    B = C.CDLL(tb.REF)
    boxes, rays = tb.cases(BOX_N, BOX_SEED)
    f = np.zeros(BOX_N, np.int32)
    B.gxref_exit_face(BOX_N, tb._f(boxes), tb._f(rays), tb._i(f))
    h, t = np.zeros(BOX_N, np.int32), np.zeros((BOX_N, 2), np.float32)
    with np.errstate(all="ignore"):
        B.gxref_box_intersect(BOX_N, tb._f(boxes), tb._f(rays), tb._i(h), tb._f(t))
    t[h != 1] = 0
    out.update(box_face=f.astype(np.int8), box_hit=h.astype(np.int8), box_t=t)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "embree_box_fixtures.npz"), **out)
    print("written", sum(v.nbytes for v in out.values()), "bytes of arrays")


if __name__ == "__main__":
    main()
