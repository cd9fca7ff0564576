"""Generates tests/golden/curve_fixtures.npz from the REFERENCE'S OWN compiled code (oracle/_ref, built from /root/reference by
`make -C oracle ref`): control points from ospray::DataDrivenPathLines::finalize (libgxy_ddpathlines_ref.so) and nearest hits from
Embree 3.6.1's round-Bezier sweep intersector (libgxy_embree_curve_ref.so) on them.  /root/reference does not exist on the GPU
box and a fresh checkout has no oracle/_ref: the committed fixture keeps the two PathLines pins alive there
(tests/test_oracle_curves.py::test_committed_reference_fixtures).

  python tests/golden/make_curve_fixtures.py          (in the build container, after `make -C oracle ref`)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_oracle_curves import _p, helices, intersect, rays_at  # noqa: E402


def main():
    B = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgxy_ddpathlines_ref.so"))
    E = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgxy_embree_curve_ref.so"))
    out = {}
    for k, (seed, radii) in enumerate(((7, (0.002, 0.06, 0.0, 1.7)), (11, (0.03, 0.03, 0.0, 0.0)), (12, (0.05, 0.01, 0.2, 1.0)))):
        verts, data, conn = helices(seed)
        cp = np.zeros((len(conn), 4, 4), np.float32)
        rc = B.gxr_build_curves(C.c_int(len(verts)), _p(verts), _p(data), C.c_int(len(conn)), _p(conn), *[C.c_float(x) for x in radii], _p(cp))
        assert rc == 0
        org, d, tn, tf = rays_at(cp, 3000, seed + 50, 0.03)
        prim, tu, ng = intersect(E, "gxr_curve_intersect", cp, org, d, tn, tf, 0)
        out.update({"verts%d" % k: verts, "data%d" % k: data, "conn%d" % k: conn, "radii%d" % k: np.float32(radii), "cp%d" % k: cp,
                    "org%d" % k: org, "dir%d" % k: d, "prim%d" % k: prim, "tu%d" % k: tu, "ng%d" % k: ng})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "curve_fixtures.npz"), **out)
    print("written", sum(v.nbytes for v in out.values()), "bytes of arrays")


if __name__ == "__main__":
    main()
