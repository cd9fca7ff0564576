"""Generates tests/golden/embree_traversal_fixtures.npz: nearest-hit answers (primID, t, u, v) of the reference's own Embree 3.6.1
-- binned-SAH BVH8/Triangle4 build + rtcIntersect8 packet TRAVERSAL, compiled from /root/reference by oracle/embree.mk -- for seeded
rays on seeded meshes.  Run in the build container (needs oracle/_ref/libgxy_embree_scene_ref.so):  python tests/golden/make_embree_traversal_fixtures.py
The cases are regenerated from their seeds by the tests; only Embree's answers are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from galaxy_b200 import scenes  # noqa: E402
from tests import util  # noqa: E402

# (kind, a, b, ray seed, n rays): "balls" = the bumpy eight-sphere mesh of the headline workload at n_lat=a, n_lon=b (shared edges:
# genuine exact ties), "soup" = a random triangles with seed b
CASES = [("balls", 24, 48, 101, 30000), ("balls", 60, 120, 102, 30000), ("soup", 20000, 7, 103, 20000)]


def case(kind, a, b, seed, n):
    tri = scenes.eightballs_mesh(a, b) if kind == "balls" else util.random_soup(a, 0, b)[0]
    org, d = util.random_rays(n, seed)
    # half of the rays from outside towards the data (camera-like), half born on a sphere of radius 0.45 around a ball centre (AO-like,
    # short interval)
    rng = np.random.default_rng(seed + 1)
    tn = np.full(n, 0.001, np.float32)
    tf = np.where(rng.uniform(size=n) < 0.5, np.float32(50.0), np.float32(0.3)).astype(np.float32)
    return tri, org, d, tn, tf


def main():
    from oracle import embree_scene
    out = {}
    for k, c in enumerate(CASES):
        tri, org, d, tn, tf = case(*c)
        es = embree_scene.EmbreeScene(tri.verts, tri.indices)
        prim, tuv, _ = es.intersect(org, d, tn, tf, packet=8)
        prim1, tuv1, _ = es.intersect(org, d, tn, tf, packet=1)
        out["prim%d" % k], out["tuv%d" % k] = prim, tuv
        out["prim1_%d" % k] = prim1
        print(c, "hits", int((prim >= 0).sum()), "packet/single id differences", int(((prim != prim1) & (prim >= 0)).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "embree_traversal_fixtures.npz"), **out)


if __name__ == "__main__":
    main()
