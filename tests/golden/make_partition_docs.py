#!/usr/bin/env python
"""Generates tests/golden/partition_docs.json by running the REFERENCE's own scripts/createPartitionDoc.py (unmodified, from
/root/reference) for the partition counts and grids the tests use.  Run in the build container only (the reference tree does not
exist on the GPU box); the output is the committed fixture that pins galaxy_b200.scenes.geometry_extents / factor.
usage: python tests/golden/make_partition_docs.py"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/scripts/createPartitionDoc.py"
CASES = [dict(origin=-1.0, counts=256, spacing=2.0 / 255), dict(origin=-1.0, counts=101, spacing=0.02), dict(origin=0.0, counts=67, spacing=0.125)]
NPARTS = [1, 2, 3, 4, 5, 6, 8, 12, 16, 27]

out = []
for c in CASES:
    for n in NPARTS:
        r = subprocess.run([sys.executable, SCRIPT, "-o", repr(c["origin"]), "-c", str(c["counts"]), "-s", repr(c["spacing"]), str(n)],
                           capture_output=True, text=True, check=True)
        out.append(dict(c, nparts=n, doc=json.loads(r.stdout)))
json.dump(out, open(os.path.join(HERE, "partition_docs.json"), "w"), indent=0)
print("wrote", len(out), "partition documents")
