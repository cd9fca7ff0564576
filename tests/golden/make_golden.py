"""Copies the reference's parity fixtures (state files + gold PNGs) into tests/golden/.
Run in the build container only: python tests/golden/make_golden.py"""
import glob
import os
import shutil

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))
for sub, pat in (("states", "*.state"), ("golds", "golds/*.png")):
    os.makedirs(os.path.join(HERE, sub), exist_ok=True)
    for f in glob.glob(os.path.join(REF, pat)):
        dst = os.path.join(HERE, sub, os.path.basename(f))
        shutil.copyfile(f, dst)
        os.chmod(dst, 0o644)
        print("copied", f)
