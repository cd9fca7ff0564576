"""The synthetic volume of the C3/C4 workloads: v = eightBalls(p) + 0.15 fbm(8p), 5-octave value noise on a PCG32-hashed
lattice (seed 7).  The torch generator that bench.py runs slab by slab on the device must be the numpy definition (scenes.fbm3)."""
import numpy as np

from galaxy_b200 import scenes


def _reference(n):
    c = -1.0 + np.arange(n) * (2.0 / (n - 1))
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")
    eb = np.sqrt((np.abs(X) - .5) ** 2 + (np.abs(Y) - .5) ** 2 + (np.abs(Z) - .5) ** 2)
    return (eb + 0.15 * scenes.fbm3(8.0 * np.stack([X, Y, Z], -1), scenes.NOISE_SEED)).astype(np.float32)


def test_torch_generator_equals_the_numpy_definition():
    n = 20
    ds = scenes.noise_volume(n)
    full = ds.data[0:n, 0:n, 0:n]
    ref = _reference(n)
    assert full.dtype == np.float32 and full.shape == (n, n, n)
    assert np.abs(full - ref).max() <= 2e-6
    # value distribution as the survey states it: float32 in [0, ~1.9]
    assert 0.0 <= full.min() and full.max() < 2.0
    # a partition's brick is the same voxels, whatever block it is cut into
    assert np.array_equal(ds.data[3:9, 2:20, 5:6], full[3:9, 2:20, 5:6])
    assert ds.counts == (n, n, n) and abs(float(ds.deltas[0]) - 2.0 / (n - 1)) < 1e-7


def test_partitions_of_the_lazy_volume_render_like_the_materialised_one():
    """build_partitions cuts bricks out of the lazy volume exactly as out of an array (2 partitions, ghost shells included)"""
    from oracle import oracle
    n = 24
    lazy = scenes.noise_volume(n)
    solid = scenes.VolumeDataset([-1, -1, -1], (n, n, n), lazy.deltas, lazy.data[0:n, 0:n, 0:n])
    vis = dict(annotation="", lighting=scenes.parse_lighting({}),
               operators=[dict(type="VolumeVis", dataset="v", colormap=[[0.0, 1.0, 0.5, 0.5], [1.0, 0.5, 0.5, 1.0]], opacitymap=[[0.0, 0.05], [1.0, 0.0]],
                               data_range=None, slices=[], isovalues=[0.6], volume_render=True)])
    cam = scenes.parse_camera({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})
    a, _ = oracle.render(scenes.build_partitions(oracle, vis, {"v": lazy}, 2), cam, vis["lighting"], 48, 48, 0.001)
    b, _ = oracle.render(scenes.build_partitions(oracle, vis, {"v": solid}, 2), cam, vis["lighting"], 48, 48, 0.001)
    assert np.array_equal(a, b) and a.max() > 0
