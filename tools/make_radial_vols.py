#!/usr/bin/env python
"""Writes the closed-form test volumes of the reference (src/apps/radial.cpp:208-231: oneBall, eightBalls, xramp, yramp, zramp) as
.vol + .raw files in the layout scripts/vti2vol produces, so that the reference's own tests/*.state files can be rendered with
galaxy_b200/gxywriter exactly as tests/image-gold-tests.sh renders them with the reference.

  python tools/make_radial_vols.py [-n 256] [-o outdir] [name ...]
  cp tests/golden/states/oneBall.state outdir/ && cd outdir && ../galaxy_b200/gxywriter -s 512 512 oneBall.state"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from galaxy_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=256, help="grid points per axis (the reference's tests use 256)")
ap.add_argument("-o", default=".", help="output directory")
ap.add_argument("names", nargs="*", default=["oneBall", "eightBalls", "xramp", "yramp", "zramp"])
a = ap.parse_args()
os.makedirs(a.o, exist_ok=True)
for name in a.names:
    vol = scenes.radial_volume(name, a.n)
    base = os.path.join(a.o, "radial-%s" % name)
    with open(base + ".vol", "w") as f:
        f.write("float\n%f %f %f\n%d %d %d\n%f %f %f\nradial-%s.raw\n" % (*[float(x) for x in vol.origin], *vol.counts, *[float(x) for x in vol.deltas], name))
    vol.data.astype(np.float32).tofile(base + ".raw")
    print("wrote", base + ".vol", vol.counts)
