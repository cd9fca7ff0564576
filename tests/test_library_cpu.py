"""CPU-only checks of the product library: it loads, exports every symbol include/gxy_gpu.h
declares, refuses to compute without a GPU, and its host-side helpers agree with the oracle."""
import os
import re

import numpy as np
import pytest

from galaxy_b200 import gpu, scenes
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gxy_gpu.h")).read()
    names = sorted(set(re.findall(r"\b(gxy_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    L = gpu.lib()
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    if gpu.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(gpu.GxyError):
        gpu.Context(0)


def test_partition_helpers_match_reference_restatements():
    for n in (1, 2, 3, 4, 6, 8, 12, 16):
        assert gpu.factor(n) == oracle.factor(n) == scenes.factor(n)
        for grid in ((256, 256, 256), (130, 67, 41)):
            f = gpu.factor(n)
            a, b = gpu.partition(n, f, grid), oracle.partition(n, f, grid)
            assert np.array_equal(a, b)
            py = scenes.partition(f, grid)
            for r in range(n):
                assert list(a[r][3:6]) == py[r]["offsets"] and list(a[r][6:9]) == py[r]["counts"]
    assert scenes.factor(2) == (1, 1, 2) and scenes.factor(4) == (1, 2, 2) and scenes.factor(8) == (2, 2, 2)


def test_transfer_function_resampling_bit_exact():
    cmap = [[0.0, 1.0, 0.5, 0.5], [0.25, 0.5, 1.0, 0.5], [0.5, 0.5, 0.5, 1.0], [0.75, 1.0, 1.0, 0.5], [1.0, 1.0, 0.5, 1.0]]
    omap = [[0.0, 0.05], [0.2, 0.02], [0.21, 0.0], [1.0, 0.0]]
    cg, og = gpu.resample_tf(cmap, omap)
    co, oo = oracle.resample_tf(cmap, omap)
    cp, op = scenes.resample_tf(cmap, omap)
    assert np.array_equal(cg, co) and np.array_equal(og, oo)
    assert np.array_equal(cp, co) and np.array_equal(op, oo)


def test_resolve_lights_bit_exact():
    cam = dict(eye=[1.0, 3.0, -3.0], dir=[-1.0, -3.0, 3.0], up=[0.0, 1.0, 0.0], aov=30.0)
    L = dict(lights=[[1.0, 1.0, 0.0], [0.0, 0.0, 1.0], [8.0, 0.0, 0.0]], types=[1, 0, 2], n_ao=3, ao_radius=1.0, shadows=True, Ka=0.4, Kd=0.6)
    assert gpu.resolve_lights(L, cam) == oracle.resolve_lights(L, cam)


def test_halton_tables_match_reference_when_present():
    path = "/root/reference/src/renderer/UV.ih"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    blocks = re.findall(r"\{([^}]*)\}", open(path).read())
    tabs = [np.array([float(x) for x in b.replace("\n", " ").split(",") if x.strip()], np.float32) for b in blocks[:2]]
    for b, ref in zip((2, 3), tabs):
        got = []
        for i in range(256):
            inv = np.float32(1.0) / np.float32(b)
            f, r, k = inv, np.float32(0), i
            while k > 0:
                r = np.float32(r + f * np.float32(k % b))
                f = np.float32(f * inv)
                k //= b
            got.append(np.float32(float("%g" % float(r))))
        assert np.array_equal(np.array(got, np.float32), ref)
