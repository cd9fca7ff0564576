"""Geometry datasets through the C++ host: partition documents (.part) + .vtu pieces read without VTK (galaxy_b200/host/gxy_vtu.cpp),
TrianglesVis / ParticlesVis operators, against the Python front end on the same data.  CPU part: every supported DataArray
encoding yields bit-identical arrays (hashes), boxes and neighbours equal the reference's Geometry::get_partitioning restatement.
GPU part: gxywriter's PNG equals the image rendered through the Python binding."""
import json
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from galaxy_b200 import scenes
from tests import util
from tests.vtu_writer import write_vtu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "galaxy_b200", "gxywriter")
ENCODINGS = [("ascii", False, False), ("binary", False, False), ("binary", True, False), ("binary", True, True), ("appended-raw", False, False),
             ("appended-raw", True, True), ("appended-base64", False, False), ("appended-base64", True, False)]


def fnv(a):
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return str(h)


def stage_geometry(tmp, nparts, mode="appended-raw", compressed=False, header64=False, n_lat=24, n_lon=48):
    """the eightBalls mesh + a particle cloud, partitioned as scripts/partitionVTUs.vpy does, as .part documents + .vtu pieces"""
    mesh = scenes.eightballs_mesh(n_lat, n_lon)
    _, par = util.random_soup(1, 400, 5)
    ext, _ = scenes.geometry_extents(nparts)
    tparts, pparts = [], []
    for r in range(nparts):
        t = mesh if nparts == 1 else scenes.clip_triangles(mesh, ext[r])
        p = par if nparts == 1 else scenes.clip_particles(par, ext[r])
        write_vtu(os.path.join(tmp, "mesh-%d.vtu" % r), t.verts, t.indices, t.normals, t.data, mode=mode, compressed=compressed, header64=header64)
        write_vtu(os.path.join(tmp, "parts-%d.vtu" % r), p.centers, None, None, p.data.astype(np.float64), scalars_name="temperature", mode=mode,
                  compressed=compressed, header64=header64)
        tparts.append(t)
        pparts.append(p)
    for name in ("mesh", "parts"):
        json.dump({"parts": [{"filename": "%s-%d.vtu" % (name, r), "extent": [float(x) for x in ext[r]]} for r in range(nparts)]},
                  open(os.path.join(tmp, name + ".part"), "w"))
    state = {
        "Datasets": [{"name": "mesh", "type": "Triangles", "filename": "mesh.part"}, {"name": "cloud", "type": "Particles", "filename": "parts.part"}],
        "Renderer": {"epsilon": 0.001},
        "Visualizations": [{"Lighting": {"Sources": [[1, 2, -3, 0]], "shadows": True, "Ka": 0.3, "Kd": 0.7, "ao count": 4, "ao radius": 0.5},
                            "operators": [{"type": "Triangles", "dataset": "mesh", "colormap": [[0.0, 1.0, 0.2, 0.2], [1.6, 0.2, 0.2, 1.0]]},
                                          {"type": "Particles", "dataset": "cloud", "radius0": 0.02, "radius1": 0.06, "value0": 0.0, "value1": 1.0,
                                           "colormap": [[0.0, 0.2, 1.0, 0.2], [1.0, 1.0, 1.0, 0.2]]}]}],
        "Cameras": [{"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30}],
    }
    path = os.path.join(tmp, "geometry.state")
    json.dump(state, open(path, "w"))
    return path, state, ext, tparts, pparts


@pytest.mark.parametrize("mode,compressed,header64", ENCODINGS)
def test_vtu_encodings_read_bit_identically(tmp_path, mode, compressed, header64):
    assert os.path.exists(EXE), "galaxy_b200/gxywriter is not built (run __graft_entry__.build())"
    state, _, ext, tparts, pparts = stage_geometry(str(tmp_path), 2, mode, compressed, header64)
    out = subprocess.run([EXE, "--describe", "-P", "2", state], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout, parse_float=lambda t: float(np.float32(t)))
    gm = {g["name"]: g for g in got["geometries"]}
    assert gm["mesh"]["type"] == "Triangles" and gm["cloud"]["type"] == "Particles"
    for r in range(2):
        t, p = gm["mesh"]["parts"][r], gm["cloud"]["parts"][r]
        assert t["loaded"] and p["loaded"]
        assert t["n_vertices"] == len(tparts[r].verts) and t["n_connectivity"] == tparts[r].indices.size
        assert t["hash_vertices"] == fnv(tparts[r].verts) and t["hash_normals"] == fnv(tparts[r].normals)
        assert t["hash_data"] == fnv(tparts[r].data) and t["hash_connectivity"] == fnv(tparts[r].indices.astype(np.int32))
        assert p["n_vertices"] == len(pparts[r].centers) and p["hash_vertices"] == fnv(pparts[r].centers)
        assert p["hash_data"] == fnv(pparts[r].data)  # written as Float64, converted like Particles.cpp:124
        for part in (t, p):
            assert part["lmin"] == [float(ext[r][0]), float(ext[r][2]), float(ext[r][4])]
            assert part["lmax"] == [float(ext[r][1]), float(ext[r][3]), float(ext[r][5])]
            assert part["gmin"] == [float(ext[:, 0].min()), float(ext[:, 2].min()), float(ext[:, 4].min())]
            assert part["gmax"] == [float(ext[:, 1].max()), float(ext[:, 3].max()), float(ext[:, 5].max())]
            assert part["neighbors"] == scenes.geometry_neighbors(ext, r)


def test_geometry_defaults_and_errors(tmp_path):
    """no normals -> (1,0,0) with the reference's message, no scalars -> 0 (Triangles.cpp:107-132); a piece that is not
    triangles, a partition document of the wrong length and a missing piece are reported, not fatal."""
    tmp = str(tmp_path)
    mesh = scenes.eightballs_mesh(6, 12)
    write_vtu(os.path.join(tmp, "m-0.vtu"), mesh.verts, mesh.indices, None, None, mode="binary")
    json.dump({"parts": [{"filename": "m-0.vtu", "extent": [-1, 1, -1, 1, -1, 1]}]}, open(os.path.join(tmp, "m.part"), "w"))
    st = {"Datasets": [{"name": "m", "type": "Triangles", "filename": "m.part"}], "Cameras": [], "Visualizations": []}
    json.dump(st, open(os.path.join(tmp, "a.state"), "w"))
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "a.state")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "triangle set has no normals" in r.stderr
    g = json.loads(r.stdout)["geometries"][0]["parts"][0]
    nv = len(mesh.verts)
    assert g["hash_normals"] == fnv(np.tile(np.array([1, 0, 0], np.float32), nv)) and g["hash_data"] == fnv(np.zeros(nv, np.float32))
    # "Normals_" is accepted when "Normals" is absent; an undeclared scalar array is found by the name "data"
    write_vtu(os.path.join(tmp, "m-0.vtu"), mesh.verts, mesh.indices, mesh.normals, mesh.data, mode="ascii", normals_name="Normals_", declare_scalars=False)
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "a.state")], capture_output=True, text=True, timeout=60)
    g = json.loads(r.stdout)["geometries"][0]["parts"][0]
    assert g["hash_normals"] == fnv(mesh.normals) and g["hash_data"] == fnv(mesh.data)
    # missing piece
    os.remove(os.path.join(tmp, "m-0.vtu"))
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "a.state")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "error reading" in r.stderr and json.loads(r.stdout)["geometries"][0]["parts"][0]["loaded"] is False
    # an unknown dataset type is refused with the reference's message (Datasets.cpp:96-155); a missing partition document too
    st["Datasets"] = [{"name": "p", "type": "Hexahedra", "filename": "p.part"}]
    json.dump(st, open(os.path.join(tmp, "b.state"), "w"))
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "b.state")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "invalid Dataset type" in r.stderr
    st["Datasets"] = [{"name": "p", "type": "PathLines", "filename": "missing.part"}]
    json.dump(st, open(os.path.join(tmp, "b.state"), "w"))
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "b.state")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("nparts", [1, 8])
def test_gxywriter_geometry_matches_python_binding(tmp_path, nparts):
    """Same data, same state: the PNG gxywriter writes from .part/.vtu files equals the image of the Python binding, which the
    parity tests pin to the oracle."""
    from galaxy_b200 import gpu
    tmp = str(tmp_path)
    state, doc, ext, tparts, pparts = stage_geometry(tmp, nparts, "appended-raw", True, False, n_lat=40, n_lon=80)
    r = subprocess.run([EXE, "-s", "320", "240", "-P", str(nparts), state], capture_output=True, text=True, timeout=300, cwd=tmp)
    assert r.returncode == 0, r.stderr + r.stdout
    img = np.asarray(Image.open(os.path.join(tmp, "image_00000.png")).convert("RGBA"))
    st = scenes.parse_state(doc)
    vis, cam = st["visualizations"][0], st["cameras"][0]
    mesh = scenes.eightballs_mesh(40, 80)
    _, par = util.random_soup(1, 400, 5)
    parts = scenes.build_partitions(gpu, vis, {"mesh": mesh, "cloud": par}, nparts)
    gpu.render_device(parts, cam, vis["lighting"], 320, 240, st["epsilon"])
    ref = parts[0].download_rgba8(320, 240)
    frac = util.image_fraction(img, ref, tol=0)
    print("partitions", nparts, "identical pixels:", frac)
    assert frac >= 0.9999


# ---- PathLines datasets ---------------------------------------------------------------------------------
def stage_pathlines(tmp, nparts, mode="appended-raw", compressed=False):
    from tests.test_curve_host import pathlines_scene
    ds = pathlines_scene(21, 10)
    ext, _ = scenes.geometry_extents(nparts)
    pieces = []
    for r in range(nparts):
        p = ds if nparts == 1 else scenes.clip_pathlines(ds, ext[r])
        write_vtu(os.path.join(tmp, "lines-%d.vtu" % r), p.points, scalars=p.data, mode=mode, compressed=compressed, polylines=[list(l) for l in p.lines])
        pieces.append(p)
    json.dump({"parts": [{"filename": "lines-%d.vtu" % r, "extent": [float(x) for x in ext[r]]} for r in range(nparts)]},
              open(os.path.join(tmp, "lines.part"), "w"))
    state = {
        "Datasets": [{"name": "pathlines", "type": "PathLines", "filename": "lines.part"}],
        "Renderer": {"epsilon": 0.001},
        "Visualizations": [{"Lighting": {"Sources": [[1, 2, -3, 0]], "shadows": True, "Ka": 0.4, "Kd": 0.6, "ao count": 0},
                            "operators": [{"type": "PathLinesVis", "dataset": "pathlines", "colormap": [[0.0, 0.0, 1.0, 0.0], [1.2, 1.0, 0.0, 1.0]],
                                           "radius0": 0.01, "radius1": 0.05, "value0": 0.0, "value1": 1.2}]}],
        "Cameras": [{"viewpoint": [1.5, 1.0, -3.0], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 35}],
    }
    path = os.path.join(tmp, "pathlines.state")
    json.dump(state, open(path, "w"))
    return path, state, pieces


@pytest.mark.parametrize("nparts,mode,compressed", [(1, "ascii", False), (2, "appended-raw", True), (8, "binary", False)])
def test_pathlines_pieces_load_as_the_reference_loads_them(tmp_path, nparts, mode, compressed):
    """PathLines::load_from_vtkPointSet through the C++ host: vertices duplicated per poly-line in cell order, data beside
    them, connectivity = first vertex of every segment -- bit-identical to the Python front end's arrays."""
    path, state, pieces = stage_pathlines(str(tmp_path), nparts, mode, compressed)
    out = subprocess.run([EXE, "--describe", "-P", str(nparts), path], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    g = d["geometries"][0]
    assert g["type"] == "PathLines" and len(g["parts"]) == nparts
    op = d["visualizations"][0]["operators"][0]
    assert op["type"] == "PathLinesVis" and np.array_equal(np.float32(op["radii"]), np.float32([0.01, 0.05, 0.0, 1.2]))
    for r in range(nparts):
        v, dat, c = pieces[r].to_arrays()
        pr = g["parts"][r]
        assert pr["loaded"] and pr["n_vertices"] == len(v) and pr["n_connectivity"] == len(c)
        assert pr["hash_vertices"] == fnv(v.astype(np.float32)) and pr["hash_data"] == fnv(dat.astype(np.float32))
        assert pr["hash_connectivity"] == fnv(c.astype(np.int32))


def test_reference_data_driven_state_loads_verbatim(tmp_path):
    """tests/data-driven.state of the reference, unchanged, on the stand-in datasets of tools/make_data_driven.py (8 partitions):
    all four datasets load through the C++ host, the operators parse with the state's values"""
    import shutil
    import sys
    sys.path.insert(0, ROOT)
    from tools.make_data_driven import write
    tmp = str(tmp_path)
    lines, parts, mesh, vol = write(tmp, 8, n=34)
    shutil.copy(os.path.join(ROOT, "tests", "golden", "states", "data-driven.state"), tmp)
    out = subprocess.run([EXE, "--describe", "-P", "8", os.path.join(tmp, "data-driven.state")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    assert [o["type"] for o in d["visualizations"][0]["operators"]] == ["VolumeVis", "PathLinesVis", "ParticlesVis", "TrianglesVis"]
    assert np.array_equal(np.float32(d["visualizations"][0]["operators"][1]["radii"]), np.float32([0.002, 0.02, 0.0, 1.7]))
    geo = {g["name"]: g for g in d["geometries"]}
    assert set(geo) == {"pathlines", "particles", "tmesh"} and all(len(g["parts"]) == 8 and all(p["loaded"] for p in g["parts"]) for g in geo.values())
    assert sum(p["n_vertices"] for p in geo["particles"]["parts"]) >= len(parts.centers)          # ghost zones duplicate some
    assert sum(p["n_connectivity"] for p in geo["tmesh"]["parts"]) >= 3 * len(mesh.indices)
    assert sum(p["n_connectivity"] for p in geo["pathlines"]["parts"]) > 0


@pytest.mark.parametrize("header64", [False, True])
@pytest.mark.parametrize("compressed", [False, True])
def test_corrupt_size_fields_are_errors_not_crashes(tmp_path, header64, compressed):
    """Size and count fields of an appended-raw .vtu are file content: a huge byte count, a huge block count, a huge block size and a
    truncated file must all come back as 'error reading' (Geometry.cpp:176-257 prints and goes on), never as a crash or an
    out-of-bounds read (arithmetic on them used to wrap around)."""
    import struct
    tmp = str(tmp_path)
    mesh = scenes.eightballs_mesh(6, 12)
    good = os.path.join(tmp, "good.vtu")
    write_vtu(good, mesh.verts, mesh.indices, mesh.normals, mesh.data, mode="appended-raw", compressed=compressed, header64=header64)
    blob = open(good, "rb").read()
    start = blob.index(b"_", blob.index(b"<AppendedData")) + 1
    hw = 8 if header64 else 4
    fmt = "<Q" if header64 else "<I"
    big = (1 << 64) - 16 if header64 else (1 << 32) - 16
    variants = {"huge first word": blob[:start] + struct.pack(fmt, big) + blob[start + hw:],
                "truncated": blob[:start + 3 * hw + 5],
                "cut in the middle": blob[:start + (len(blob) - start) // 2]}
    if compressed:
        variants["huge block size"] = blob[:start + hw] + struct.pack(fmt, big) + blob[start + 2 * hw:]
        variants["huge compressed size"] = blob[:start + 3 * hw] + struct.pack(fmt, big) + blob[start + 4 * hw:]
    json.dump({"parts": [{"filename": "m-0.vtu", "extent": [-1, 1, -1, 1, -1, 1]}]}, open(os.path.join(tmp, "m.part"), "w"))
    st = {"Datasets": [{"name": "m", "type": "Triangles", "filename": "m.part"}], "Cameras": [], "Visualizations": []}
    json.dump(st, open(os.path.join(tmp, "a.state"), "w"))
    for what, data in variants.items():
        open(os.path.join(tmp, "m-0.vtu"), "wb").write(data)
        r = subprocess.run([EXE, "--describe", os.path.join(tmp, "a.state")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, (what, r.returncode, r.stderr[-400:])
        assert "error reading" in r.stderr, (what, r.stderr[-400:])
        assert json.loads(r.stdout)["geometries"][0]["parts"][0]["loaded"] is False, what
    # and the untouched file still loads
    open(os.path.join(tmp, "m-0.vtu"), "wb").write(blob)
    r = subprocess.run([EXE, "--describe", os.path.join(tmp, "a.state")], capture_output=True, text=True, timeout=60)
    assert json.loads(r.stdout)["geometries"][0]["parts"][0]["loaded"] is True
