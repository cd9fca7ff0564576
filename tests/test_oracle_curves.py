"""Pins the oracle's PathLines arithmetic against the REFERENCE'S OWN code: Embree 3.6.1's round-Bezier
sweep intersector (kernels/geometry/curve_intersector_sweep.h, what rtcIntersectV runs for
RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE) compiled from /root/reference into oracle/_ref (oracle/embree_curve_ref.cpp +
`make -C oracle ref`).  CPU only.  Where the .so is absent the Embree comparisons are skipped; the closed-form checks
of the curve builder and of the intersector run everywhere."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libgxy_embree_curve_ref.so")
FMAX = np.float32(3.4028234e38)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int if a.dtype == np.int32 else C.c_float))


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle.lib()


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libgxy_embree_curve_ref.so not built (needs /root/reference: make -C oracle ref)")
    L = C.CDLL(SO)
    L.gxr_curve_describe.restype = C.c_char_p
    return L


def intersect(L, fn, cp, org, d, tn, tf, per_curve):
    nc, nr = len(cp), len(org)
    n = nr * nc if per_curve else nr
    prim, tu, ng = np.zeros(n, np.int32), np.zeros((n, 2), np.float32), np.zeros((n, 3), np.float32)
    getattr(L, fn)(C.c_int(nc), _p(cp), C.c_int(nr), _p(org), _p(d), _p(tn), _p(tf), _p(prim), _p(tu), _p(ng), C.c_int(per_curve))
    return prim, tu, ng


def build(L, verts, data, conn, r0, r1, v0, v1):
    cp = np.zeros((len(conn), 4, 4), np.float32)
    rc = L.gxo_build_curves(C.c_int(len(verts)), _p(verts), _p(data), C.c_int(len(conn)), _p(conn), C.c_float(r0), C.c_float(r1),
                            C.c_float(v0), C.c_float(v1), _p(cp))
    assert rc == 0
    return cp


def helices(seed, nlines=6):
    rng = np.random.default_rng(seed)
    verts, data, conn, k = [], [], [], 0
    for _ in range(nlines):
        n = int(rng.integers(2, 12))
        t = np.linspace(0, 2 + rng.uniform(0, 3), n)
        c = rng.uniform(-.5, .5, 3)
        v = np.stack([c[0] + 0.5 * np.cos(t), c[1] + 0.5 * np.sin(t), c[2] + 0.2 * t], 1)
        for i in range(n):
            verts.append(v[i]); data.append(np.linalg.norm(v[i]))
            if i < n - 1:
                conn.append(k)
            k += 1
    return np.array(verts, np.float32), np.array(data, np.float32), np.array(conn, np.int32)


def rays_at(cp, n, seed, spread):
    rng = np.random.default_rng(seed)
    org = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    tgt = cp[rng.integers(0, len(cp), n), rng.integers(0, 4, n), :3] + rng.normal(scale=spread, size=(n, 3))
    d = (tgt - org).astype(np.float32)
    d[::2] /= np.linalg.norm(d[::2], axis=1, keepdims=True)      # half normalised (primaries), half not (shadow rays)
    return org, np.ascontiguousarray(d), np.zeros(n, np.float32), np.full(n, FMAX, np.float32)


def compare(a, b, per_curve):
    hit_a, hit_b = (a[0] >= (1 if per_curve else 0)), (b[0] >= (1 if per_curve else 0))
    n = len(hit_a)
    # hit/miss decisions and nearest primitive: identical but for grazing rays (Embree's rcp/rsqrt are rcpss/rsqrtss +
    # one Newton step, the oracle's are IEEE); the fraction is asserted tiny and reported
    disagree = int((hit_a != hit_b).sum()) + int((hit_a & hit_b & (a[0] != b[0])).sum())
    assert disagree <= max(2, n // 50000), disagree
    both = hit_a & hit_b & (a[0] == b[0])
    assert both.sum() > 1000
    ta, tb = a[1][both, 0], b[1][both, 0]
    rel = np.abs(ta - tb) / np.abs(ta)
    assert np.quantile(rel, 0.999) < 1e-5 and rel.max() < 1e-3
    assert np.quantile(np.abs(a[1][both, 1] - b[1][both, 1]), 0.999) < 1e-4          # curve parameter u
    na = a[2][both] / np.linalg.norm(a[2][both], axis=1, keepdims=True)
    nb = b[2][both] / np.linalg.norm(b[2][both], axis=1, keepdims=True)
    assert np.quantile((na * nb).sum(1), 0.001) > 0.9999
    return float((ta == tb).mean())


def test_curve_builder_closed_form(orc):
    """DataDrivenPathLines.cpp:103-156 on a 4-vertex line + a 2-vertex line: shared joints, doubled ends, tangents."""
    verts = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 0], [5, 5, 5], [5, 5, 6]], np.float32)
    data = np.array([0, 1, 2, 3, 1, 1], np.float32)
    conn = np.array([0, 1, 2, 4], np.int32)
    cp = build(orc, verts, data, conn, 0.1, 0.4, 0.0, 3.0)
    r = lambda d: np.float32(0.1) + np.float32(d / 3.0) * np.float32(0.3)
    # line ends repeat their end point and radius
    assert (cp[0, 0] == cp[0, 1]).all() and (cp[0, 0, :3] == verts[0]).all() and cp[0, 0, 3] == np.float32(0.1)
    assert (cp[2, 2] == cp[2, 3]).all() and (cp[2, 3, :3] == verts[3]).all() and abs(cp[2, 3, 3] - 0.4) < 1e-7
    # a middle segment's 4th control point IS the next segment's first (Embree reads 4 consecutive vertices)
    assert (cp[0, 3] == cp[1, 0]).all() and (cp[1, 3] == cp[2, 0]).all()
    assert (cp[1, 0, :3] == verts[1]).all() and abs(cp[1, 0, 3] - r(1)) < 1e-7
    # tangents: delta = (next-start)/3, r = |seg|/(|seg|+|next seg|); C1 across the joint
    delta = (verts[2] - verts[0]) / 3
    rr = 1.0 / (1.0 + np.sqrt(2.0))
    np.testing.assert_allclose(cp[0, 2, :3], verts[1] - rr * delta, atol=1e-6)
    np.testing.assert_allclose(cp[1, 1, :3], verts[1] + (1 - rr) * delta, atol=1e-6)
    # radii of the inner control points: lerp(1/3), lerp(2/3) of the segment's end radii
    np.testing.assert_allclose(cp[0, 2, 3], (1 / 3) * 0.1 + (2 / 3) * r(1), atol=1e-6)
    np.testing.assert_allclose(cp[1, 1, 3], (2 / 3) * r(1) + (1 / 3) * r(2), atol=1e-6)
    # the separate line is a straight doubled-end segment
    assert (cp[3, 0] == cp[3, 1]).all() and (cp[3, 2] == cp[3, 3]).all() and (cp[3, 3, :3] == verts[5]).all()
    # value0 == value1 -> radius0 everywhere (MAP_RADIUS)
    cp2 = build(orc, verts, data, conn, 0.25, 0.9, 1.0, 1.0)
    assert (cp2[..., 3] == np.float32(0.25)).all()
    # bad connectivity is refused
    bad = np.array([5], np.int32)
    out = np.zeros((1, 4, 4), np.float32)
    assert orc.gxo_build_curves(C.c_int(6), _p(verts), _p(data), C.c_int(1), _p(bad), C.c_float(.1), C.c_float(.1), C.c_float(0),
                                C.c_float(0), _p(out)) < 0


def test_straight_tube_closed_form(orc):
    """A straight constant-radius segment is a capped cylinder (flat caps: the half-planes of :165-172)."""
    cp = np.array([[[0, 0, 0, .1], [1 / 3, 0, 0, .1], [2 / 3, 0, 0, .1], [1, 0, 0, .1]]], np.float32)
    org = np.array([[.5, 0, -1], [.5, .05, -1], [.5, .2, -1], [-.05, 0, -1], [.25, 0, 2]], np.float32)
    d = np.array([[0, 0, 1]] * 4 + [[0, 0, -2]], np.float32)
    prim, tu, ng = intersect(orc, "gxo_curve_intersect", cp, org, d, np.zeros(5, np.float32), np.full(5, FMAX, np.float32), 0)
    assert prim.tolist() == [0, 0, -1, -1, 0]
    np.testing.assert_allclose(tu[0], [0.9, 0.5], atol=2e-6)
    np.testing.assert_allclose(tu[1], [1 - np.sqrt(0.01 - 0.0025), 0.5], atol=2e-6)
    np.testing.assert_allclose(tu[4], [0.95, 0.25], atol=2e-6)            # unnormalised direction: t in units of |dir|
    n1 = ng[1] / np.linalg.norm(ng[1])
    np.testing.assert_allclose(n1, [0, 0.5, -np.sqrt(0.75)], atol=1e-5)
    # the interval is open at both ends: a hit at t is not found again with tnear = t or tfar = t
    t0 = tu[0, 0]
    prim2, tu2, _ = intersect(orc, "gxo_curve_intersect", cp, org[:1], d[:1], np.array([t0], np.float32), np.full(1, FMAX, np.float32), 0)
    assert prim2[0] == 0 and abs(tu2[0, 0] - 1.1) < 2e-6                 # the exit point instead
    prim3, _, _ = intersect(orc, "gxo_curve_intersect", cp, org[:1], d[:1], np.zeros(1, np.float32), np.array([t0], np.float32), 0)
    assert prim3[0] == -1


@pytest.mark.parametrize("per_curve", [0, 1])
def test_random_curves_match_embree(orc, ref, per_curve):
    rng = np.random.default_rng(5)
    nc = 40
    cp = np.zeros((nc, 4, 4), np.float32)
    for c in range(nc):
        p, d = rng.uniform(-1, 1, 3), rng.normal(size=3)
        d /= np.linalg.norm(d)
        for k in range(4):
            cp[c, k, :3] = p
            p = p + 0.25 * (d + 0.5 * rng.normal(size=3))
        cp[c, :, 3] = rng.uniform(0.01, 0.08) * (1 + np.array([0, *rng.uniform(-.3, .3, 3)]))
    org, d, tn, tf = rays_at(cp, 20000, 11, 0.05)
    a = intersect(ref, "gxr_curve_intersect", cp, org, d, tn, tf, per_curve)
    b = intersect(orc, "gxo_curve_intersect", cp, org, d, tn, tf, per_curve)
    exact = compare(a, b, per_curve)
    assert exact > 0.5      # most t are bit-identical; the rest is rcp/rsqrt


@pytest.mark.parametrize("radii", [(0.002, 0.06, 0.0, 1.7), (0.03, 0.03, 0.0, 0.0), (0.05, 0.01, 0.2, 1.0)])
def test_built_pathlines_match_embree(orc, ref, radii):
    """Curves as Galaxy builds them (doubled end points: zero end tangents) with data-mapped radii."""
    verts, data, conn = helices(7)
    cp = build(orc, verts, data, conn, *radii)
    org, d, tn, tf = rays_at(cp, 30000, 13, 0.03)
    a = intersect(ref, "gxr_curve_intersect", cp, org, d, tn, tf, 0)
    b = intersect(orc, "gxo_curve_intersect", cp, org, d, tn, tf, 0)
    compare(a, b, 0)
    # bounded intervals: tnear/tfar inside the scene
    tn2 = np.full(len(org), 2.5, np.float32)
    tf2 = np.full(len(org), 3.5, np.float32)
    a = intersect(ref, "gxr_curve_intersect", cp, org, d, tn2, tf2, 0)
    b = intersect(orc, "gxo_curve_intersect", cp, org, d, tn2, tf2, 0)
    compare(a, b, 0)


def test_curve_builder_matches_the_references_own_finalize(orc):
    """gxo_build_curves against ospray::DataDrivenPathLines::finalize ITSELF (src/ospray/DataDrivenPathLines.cpp compiled from the
    reference tree into oracle/_ref behind oracle/ddpathlines_ref.cpp, whose stand-in for the ISPC export setCurve captures the
    curve vertices Embree would be handed): bit for bit, for data-mapped, constant, negative-default and shrinking radii, single
    segments and lines that share no vertices."""
    so = os.path.join(ROOT, "oracle", "_ref", "libgxy_ddpathlines_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libgxy_ddpathlines_ref.so not built (needs /root/reference: make -C oracle ref)")
    R = C.CDLL(so)
    cases = [(helices(seed, nlines=9), radii) for seed, radii in ((3, (0.002, 0.02, 0.0, 1.7)), (4, (0.05, 0.05, 1.0, 1.0)), (5, (-1.0, 1.0, 0.0, 1.0)),
                                                                   (6, (0.04, 0.01, 0.3, 0.9)))]
    v = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 0], [5, 5, 5], [5, 5, 6]], np.float32)
    cases.append(((v, np.array([0, 1, 2, 3, 1, 1], np.float32), np.array([0, 1, 2, 4], np.int32)), (0.1, 0.4, 0.0, 3.0)))
    cases.append(((v, np.array([0, 1, 2, 3, 1, 1], np.float32), np.array([3], np.int32)), (0.1, 0.4, 0.0, 3.0)))          # one segment
    cases.append(((v, np.array([9, -1, 2, 7, 1, 1], np.float32), np.array([0, 2, 4], np.int32)), (0.1, 0.4, 0.0, 3.0)))     # no segment continues
    for (verts, data, conn), radii in cases:
        want = build(orc, verts, data, conn, *radii)
        got = np.zeros_like(want)
        rc = R.gxr_build_curves(C.c_int(len(verts)), _p(verts), _p(data), C.c_int(len(conn)), _p(conn), *[C.c_float(x) for x in radii], _p(got))
        assert rc == 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_committed_reference_fixtures(orc):
    """tests/golden/curve_fixtures.npz holds what the reference's own compiled code answered in the build container
    (tests/golden/make_curve_fixtures.py: DataDrivenPathLines::finalize control points, Embree sweep-intersector hits).  It keeps
    both PathLines pins alive where oracle/_ref does not exist: control points bit for bit, hit decisions and nearest segments
    identical but for grazing rays, t within 1e-5."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "curve_fixtures.npz"))
    for k in range(3):
        verts, data, conn, radii = fx["verts%d" % k], fx["data%d" % k], fx["conn%d" % k], [float(x) for x in fx["radii%d" % k]]
        cp = build(orc, verts, data, conn, *radii)
        assert np.array_equal(cp.view(np.uint32), fx["cp%d" % k].view(np.uint32))
        org, d = fx["org%d" % k], fx["dir%d" % k]
        tn, tf = np.zeros(len(org), np.float32), np.full(len(org), FMAX, np.float32)
        b = intersect(orc, "gxo_curve_intersect", cp, org, d, tn, tf, 0)
        compare((fx["prim%d" % k], fx["tu%d" % k], fx["ng%d" % k]), b, 0)
