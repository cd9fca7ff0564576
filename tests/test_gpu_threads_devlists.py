"""-m gpu: the boundary as the reference uses it: TraceRays::Trace is called from 5 pool threads at once on ONE shared
Visualization (src/framework/Application.cpp:75, src/renderer/Renderer.cpp:504-556), and a RayList may stay on the device between
Trace, Classify and the next Trace (device-resident list handles)."""
import threading

import numpy as np
import pytest

from galaxy_b200 import scenes
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from galaxy_b200 import gpu as g
    return g


def _scene(gpu, backend):
    tri, par = util.random_soup(3000, 200, 5)
    vis = util.soup_vis(True)
    return scenes.build_partitions(backend, vis, {"tris": tri, "parts": par}, 1)[0], vis


def _diff(a, b):
    """which columns of two (columns, n) arrays differ, in how many rays, and the first such ray"""
    bad = a.view(np.int32) != b.view(np.int32)
    cols = np.nonzero(bad.any(axis=1))[0]
    rays = np.nonzero(bad.any(axis=0))[0]
    return {"columns": cols.tolist(), "n_rays": int(rays.size), "first_ray": int(rays[0]) if rays.size else -1, "of": int(a.shape[1])}


def test_five_threads_trace_on_one_visualization(gpu):
    """5 threads x 6 calls each, every thread its own RayList (different cameras and sizes): every result equals the result of the
    same call made alone (bit for bit: the trace is deterministic per list), nothing leaks between the lanes."""
    g, vis = _scene(gpu, gpu)
    cams = [scenes.parse_camera({"viewpoint": vp, "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 35})
            for vp in ([3, 2, -4], [-3, 1, -4], [0.5, 3, -4.5], [4, -1, 2], [-2, -2, -4])]
    sizes = [(96, 64), (128, 80), (64, 64), (160, 96), (80, 120)]
    lists, alone = [], []
    for cam, (w, h) in zip(cams, sizes):
        rays, n = g.generate_rays(cam, w, h)
        L = gpu.resolve_lights(vis["lighting"], cam)
        lists.append((rays, n, L))
        r = rays.copy()
        sec, ns, hits = g.trace_raylist(L, r, n, 0.001, want_hits=True)
        g.classify(r, n)
        alone.append((r, sec, ns, hits))
    errors = []

    def worker(k):
        try:
            rays, n, L = lists[k]
            for _ in range(6):
                r = rays.copy()
                sec, ns, hits = g.trace_raylist(L, r, n, 0.001, want_hits=True)
                g.classify(r, n)
                ref = alone[k]
                assert ns == ref[2], (k, ns, ref[2])
                assert np.array_equal(hits, ref[3]), (k, "hit ids", _diff(hits.T, ref[3].T))
                assert np.array_equal(r[:, :n].view(np.int32), ref[0][:, :n].view(np.int32)), (k, "rays", _diff(r[:, :n], ref[0][:, :n]))
                if ns:
                    assert np.array_equal(sec[:, :ns].view(np.int32), ref[1][:, :ns].view(np.int32)), (k, "secondaries",
                                                                                                         _diff(sec[:, :ns], ref[1][:, :ns]))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(5)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_device_resident_lists_match_the_host_lists(gpu):
    """upload -> trace (in place, secondary list stays on the device) -> classify -> trace the secondaries -> download: every
    column equals what the host-list calls return for the same rays"""
    from oracle import oracle
    g, vis = _scene(gpu, gpu)
    o, _ = _scene(gpu, oracle)
    cam = scenes.c5_camera()
    rays, n = g.generate_rays(cam, 128, 96)
    L = gpu.resolve_lights(vis["lighting"], cam)
    # host-list path
    r_host = rays.copy()
    sec_host, ns_host, _ = g.trace_raylist(L, r_host, n, 0.001)
    g.classify(r_host, n)
    s_host = sec_host.copy()
    sec2_host, ns2_host, _ = g.trace_raylist(L, s_host, ns_host, 0.001)
    g.classify(s_host, ns_host)
    # device-list path
    d = g.upload_raylist(rays, n)
    assert d.n == n
    dsec = d.trace(L, 0.001)
    d.classify()
    assert dsec is not None and dsec.n == ns_host
    dsec2 = dsec.trace(L, 0.001)          # AO/shadow rays spawn nothing
    dsec.classify()
    assert dsec2 is None and ns2_host == 0
    r_dev, n_dev = d.download()
    s_dev, ns_dev = dsec.download()
    assert n_dev == n and ns_dev == ns_host
    assert np.array_equal(r_dev[:, :n].view(np.int32), r_host[:, :n].view(np.int32))
    assert np.array_equal(s_dev[:, :ns_dev].view(np.int32), s_host[:, :ns_host].view(np.int32))
    # and the oracle agrees on what matters (term, classification, t)
    ro, no = o.generate_rays(cam, 128, 96)
    so, nso, _ = o.trace_raylist(L, ro, no, 0.001)
    o.classify(ro, no)
    assert no == n and nso == ns_host
    assert np.array_equal(util.icol(ro, "term", no), util.icol(r_dev, "term", n))
    assert np.array_equal(util.icol(ro, "classification", no), util.icol(r_dev, "classification", n))
    # an empty list and a download into a list that is too small
    e = g.upload_raylist(rays, 0)
    assert e.n == 0 and e.trace(L, 0.001) is None
    small = np.zeros((25, 16), np.float32)
    with pytest.raises(gpu.GxyError):
        gpu.check(gpu.lib().gxy_raylist_download(d.h, gpu.RayListView(small.ctypes.data_as(gpu.C.POINTER(gpu.C.c_float)), 0, 16)))
