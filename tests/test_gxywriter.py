"""The C++ host side above the C ABI (galaxy_b200/host: Datasets / Volume / Camera / Lighting / Visualization /
Rendering + the gxywriter driver): an existing reference .state file, with its .vol datasets on disk, renders
unchanged.  CPU part: `gxywriter --describe` (no GPU) parses every reference test state to exactly what the Python
front end (galaxy_b200.scenes, itself pinned by the gold images through the oracle) parses.  GPU part: the PNGs
gxywriter writes match the reference's gold images."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest
from PIL import Image

from galaxy_b200 import scenes
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "galaxy_b200", "gxywriter")
STATES = ["oneBall", "nineBalls", "xyz", "camera", "camera-shadow", "absolute", "absolute-shadow", "infinite", "infinite-shadow"]


def write_vol(path, vol):
    """scripts/vti2vol:70-80: text header (type / origin / counts / deltas / raw file name, %f) + raw file."""
    raw = os.path.basename(path)[:-4] + ".raw"
    with open(path, "w") as f:
        f.write("float\n%f %f %f\n%d %d %d\n%f %f %f\n%s\n" % (*[float(x) for x in vol.origin], *vol.counts, *[float(x) for x in vol.deltas], raw))
    vol.data.astype(np.float32).tofile(os.path.join(os.path.dirname(path), raw))


def stage(tmp, name, n):
    """state file + the radial .vol files it names, in one directory (what tests/image-gold-tests.sh sets up)."""
    src = os.path.join(ROOT, "tests", "golden", "states", name + ".state")
    shutil.copy(src, os.path.join(tmp, name + ".state"))
    doc = json.load(open(src))
    for ds in doc["Datasets"]:
        fn = ds["filename"]
        p = os.path.join(tmp, fn)
        if not os.path.exists(p):
            write_vol(p, scenes.radial_volume(fn[len("radial-"):-len(".vol")], n))
    return os.path.join(tmp, name + ".state"), doc


def f32list(x):
    return [float(np.float32(v)) for v in np.asarray(x, np.float64).ravel()]


@pytest.mark.parametrize("name", STATES)
@pytest.mark.parametrize("nparts", [1, 8])
def test_describe_matches_python_front_end(tmp_path, name, nparts):
    assert os.path.exists(EXE), "galaxy_b200/gxywriter is not built (run __graft_entry__.build())"
    state, doc = stage(str(tmp_path), name, 20)
    out = subprocess.run([EXE, "--describe", "-P", str(nparts), state], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout, parse_float=lambda t: float(np.float32(t)))  # %.9g text -> the float32 it denotes
    st = scenes.parse_state(doc)
    assert got["epsilon"] == float(np.float32(st["epsilon"]))
    assert len(got["cameras"]) == len(st["cameras"])
    for g, c in zip(got["cameras"], st["cameras"]):
        assert g["eye"] == f32list(c["eye"]) and g["dir"] == f32list(c["dir"]) and g["up"] == f32list(c["up"])
        assert g["aov"] == float(np.float32(c["aov"])) and g["annotation"] == c["annotation"]
    assert len(got["visualizations"]) == len(st["visualizations"])
    for g, v in zip(got["visualizations"], st["visualizations"]):
        L = v["lighting"]
        assert g["annotation"] == v["annotation"]
        assert [f32list(x) for x in L["lights"]] == g["lighting"]["lights"] and L["types"] == g["lighting"]["types"]
        assert g["lighting"]["n_ao"] == L["n_ao"] and g["lighting"]["shadows"] == L["shadows"]
        assert g["lighting"]["ao_radius"] == f32list([L["ao_radius"]]) and g["lighting"]["Ka"] == f32list([L["Ka"]])
        assert g["lighting"]["Kd"] == f32list([L["Kd"]])
        assert len(g["operators"]) == len(v["operators"])
        for go, op in zip(g["operators"], v["operators"]):
            assert go["type"] == op["type"] and go["dataset"] == op["dataset"]
            assert go["isovalues"] == f32list(op["isovalues"]) and go["slices"] == f32list(op["slices"])
            assert go["volume_render"] == op["volume_render"]
            lo, hi = op["data_range"] if op["data_range"] is not None else (op["colormap"][0][0], op["colormap"][-1][0])
            assert go["range"] == f32list([lo, hi])
            col, opac = scenes.resample_tf(op["colormap"], op["opacitymap"])
            assert np.array_equal(np.asarray(go["tf_colors"], np.float32), col.ravel())
            assert np.array_equal(np.asarray(go["tf_opacities"], np.float32), opac.ravel())
    # datasets: header values as the reference reads them, partition boxes / neighbours as Volume.cpp computes them
    fac = scenes.factor(nparts)
    for g in got["datasets"]:
        fn = [d["filename"] for d in doc["Datasets"] if d.get("name", d["filename"]) == g["name"]][0]
        vol = scenes.load_vol_file(os.path.join(str(tmp_path), fn))
        assert g["counts"] == list(vol.counts) and g["origin"] == f32list(vol.origin) and g["deltas"] == f32list(vol.deltas)
        parts = scenes.partition(fac, vol.counts)
        for gp, p in zip(g["partitions"], parts):
            gmin, gmax, lmin, lmax = scenes.volume_boxes(vol, p)
            assert gp["gmin"] == f32list(gmin) and gp["gmax"] == f32list(gmax) and gp["lmin"] == f32list(lmin) and gp["lmax"] == f32list(lmax)
            assert gp["neighbors"] == scenes.neighbors_of(p["ijk"], fac)
            assert gp["goffsets"] == p["goffsets"] and gp["gcounts"] == p["gcounts"]


def test_errors_are_reported_not_fatal(tmp_path):
    """The reference's loaders print and return false; the driver turns that into exit status 1."""
    bad = tmp_path / "bad.state"
    bad.write_text('{"Datasets": [{"name": "x", "type": "Volume", "filename": "missing.vol"}], "Cameras": [], "Visualizations": []}')
    r = subprocess.run([EXE, "--describe", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "unable to open volfile" in r.stderr
    bad.write_text("{ not json")
    r = subprocess.run([EXE, "--describe", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Bad state file" in r.stderr
    geo = tmp_path / "geo.state"
    geo.write_text('{"Datasets": [{"name": "m", "type": "Triangles", "filename": "m.part"}], "Cameras": [], "Visualizations": []}')
    r = subprocess.run([EXE, "--describe", str(geo)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "partition document" in r.stderr  # the .part file does not exist


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_images", [("oneBall", 1), ("nineBalls", 3), ("xyz", 1), ("camera-shadow", 1)])
def test_gxywriter_renders_reference_states_to_gold(tmp_path, golden_dir, name, n_images):
    from tests.test_oracle_golds import check_gold_fraction  # 99.9 % for every gold; nineBalls_0/_1: documented deviation -> xfail
    state, _ = stage(str(tmp_path), name, 256)
    r = subprocess.run([EXE, "-s", "512", "512", state], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr + r.stdout
    assert "TIMING total" in r.stdout
    fracs = []
    for k in range(n_images):
        img = np.asarray(Image.open(os.path.join(str(tmp_path), "image_%05d.png" % k)).convert("RGBA"))
        gold = np.asarray(Image.open(os.path.join(golden_dir, "golds", "%s_%05d.png" % (name, k))).convert("RGBA"))
        assert img.shape == gold.shape
        fracs.append(util.image_fraction(img, gold))
        print(name, k, "fraction within 1/255 of the gold:", fracs[-1])
    for k in reversed(range(n_images)):   # (the documented deviations, nineBalls_0/_1, last: an xfail ends the test)
        check_gold_fraction(name, k, fracs[k])


@pytest.mark.gpu
def test_gxywriter_partitions_match_single(tmp_path):
    """-P 8: same state rendered on 8 spatial partitions equals the 1-partition image to within the tolerance."""
    state, _ = stage(str(tmp_path), "nineBalls", 256)
    for P in (1, 8):
        r = subprocess.run([EXE, "-s", "384", "384", "-P", str(P), "-o", "p%d" % P, state], capture_output=True, text=True, timeout=300,
                           cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr + r.stdout
    a = np.asarray(Image.open(os.path.join(str(tmp_path), "p1_00001.png")).convert("RGBA"))
    b = np.asarray(Image.open(os.path.join(str(tmp_path), "p8_00001.png")).convert("RGBA"))
    assert util.image_fraction(a, b) >= 0.995


def test_camera_and_colormap_files(tmp_path):
    """"Cameras": ["cam.json"] (a ParaView camera configuration in JSON form, Camera.cpp:168-232) and "colormap": "cmap.json"
    (a ParaView colormap export, MappedVis.cpp:104-166): same values from the C++ host and the Python front end."""
    tmp = str(tmp_path)
    write_vol(os.path.join(tmp, "radial-oneBall.vol"), scenes.radial_volume("oneBall", 12))
    cam = {"PVCameraConfiguration": {"Proxy": {"Property": [
        {"@name": "CameraPosition", "Element": [{"@value": "1.25"}, {"@value": "-2.5"}, {"@value": "3.1"}]},
        {"@name": "CameraFocalPoint", "Element": [{"@value": "0.1"}, {"@value": "0.2"}, {"@value": "0.3"}]},
        {"@name": "CameraViewUp", "Element": [{"@value": "0"}, {"@value": "0"}, {"@value": "1"}]},
        {"@name": "CameraViewAngle", "Element": {"@value": "27.5"}},
        {"@name": "Ignored", "Element": {"@value": "1"}}]}}}
    json.dump(cam, open(os.path.join(tmp, "cam.json"), "w"))
    cmap = [{"Name": "test", "RGBPoints": [0.0, 0.1, 0.2, 0.3, 0.5, 0.9, 0.8, 0.7, 1.5, 1.0, 1.0, 0.0],
             "Points": [0.0, 0.0, 0.5, 0.0, 0.7, 0.25, 0.5, 0.0, 1.5, 1.0, 0.5, 0.0]}]
    json.dump(cmap, open(os.path.join(tmp, "cmap.json"), "w"))
    json.dump({"RGBPoints": [0.0, 1.0, 0.0, 0.0, 2.0, 0.0, 0.0, 1.0]}, open(os.path.join(tmp, "cmap2.json"), "w"))
    doc = {"Datasets": [{"name": "oneBall", "type": "Volume", "filename": "radial-oneBall.vol"}],
           "Visualizations": [{"operators": [{"type": "Volume", "dataset": "oneBall", "volume rendering": True, "colormap": "cmap.json",
                                              "opacitymap": [[0, 0.5], [1, 0.5]]},
                                             {"type": "Volume", "dataset": "oneBall", "isovalues": [0.4], "transfer function": "cmap2.json"}]}],
           "Cameras": ["cam.json"]}
    state = os.path.join(tmp, "files.state")
    json.dump(doc, open(state, "w"))
    out = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout, parse_float=lambda t: float(np.float32(t)))
    st = scenes.parse_state(doc, base_dir=tmp)
    c, g = st["cameras"][0], got["cameras"][0]
    assert g["eye"] == f32list(c["eye"]) and g["dir"] == f32list(c["dir"]) and g["up"] == f32list(c["up"]) and g["aov"] == float(np.float32(c["aov"]))
    assert g["eye"] == f32list([1.25, -2.5, 3.1]) and g["aov"] == 27.5
    for go, op in zip(got["visualizations"][0]["operators"], st["visualizations"][0]["operators"]):
        col, opac = scenes.resample_tf(op["colormap"], op["opacitymap"])
        assert np.array_equal(np.asarray(go["tf_colors"], np.float32), col.ravel())
        assert np.array_equal(np.asarray(go["tf_opacities"], np.float32), opac.ravel())
        assert go["range"] == f32list([op["colormap"][0][0], op["colormap"][-1][0]])
    # the file's own opacity points win over an "opacitymap" next to a colormap file; no "Points" -> opacity 1
    assert st["visualizations"][0]["operators"][0]["opacitymap"] == [[0.0, 0.0], [0.7, 0.25], [1.5, 1.0]]
    assert st["visualizations"][0]["operators"][1]["opacitymap"] == [[0.0, 1.0], [1.0, 1.0]]
    # a missing file is reported
    doc["Cameras"] = ["nope.json"]
    json.dump(doc, open(state, "w"))
    r = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "error loading camera" in r.stderr


def test_pvcc_camera_file(tmp_path):
    """"Cameras": ["cam.pvcc"]: ParaView's own XML camera configuration (Camera::LoadFromPVCC, Camera.cpp:118-163; tried when the
    file is not JSON, :182-191) -- same camera from the C++ host and the Python front end; a broken file is refused."""
    tmp = str(tmp_path)
    write_vol(os.path.join(tmp, "radial-oneBall.vol"), scenes.radial_volume("oneBall", 12))
    open(os.path.join(tmp, "cam.pvcc"), "w").write("""<?xml version="1.0"?>
<!-- saved by ParaView -->
<PVCameraConfiguration description="ParaView camera configuration" version="1.0">
  <Proxy group="views" type="RenderView" id="4875" servers="21">
    <Property name="CameraPosition" id="4875.CameraPosition" number_of_elements="3">
      <Element index="0" value="1.2345678901234"/>
      <Element index="1" value="-2.5"/>
      <Element index="2" value="3.1e0"/>
    </Property>
    <Property name="CameraFocalPoint" id="4875.CameraFocalPoint" number_of_elements="3">
      <Element index="2" value="0.3"/>
      <Element index="0" value="0.1"/>
      <Element index="1" value="0.2"/>
    </Property>
    <Property name="CameraViewUp" id="4875.CameraViewUp" number_of_elements="3">
      <Element index="0" value="0"/> <Element index="1" value="0"/> <Element index="2" value="1"/>
    </Property>
    <Property name="CameraViewAngle" id="4875.CameraViewAngle" number_of_elements="1">
      <Element index="0" value="27.5"/>
    </Property>
    <Property name="CameraParallelProjection" id="4875.CameraParallelProjection" number_of_elements="1">
      <Element index="0" value="0"/>
      <Domain name="bool" id="4875.CameraParallelProjection.bool"/>
    </Property>
  </Proxy>
</PVCameraConfiguration>
""")
    doc = {"Datasets": [{"name": "oneBall", "type": "Volume", "filename": "radial-oneBall.vol"}],
           "Visualizations": [{"operators": [{"type": "Volume", "dataset": "oneBall", "isovalues": [0.4]}]}],
           "Cameras": ["cam.pvcc"]}
    state = os.path.join(tmp, "pvcc.state")
    json.dump(doc, open(state, "w"))
    out = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout, parse_float=lambda t: float(np.float32(t)))
    st = scenes.parse_state(doc, base_dir=tmp)
    c, g = st["cameras"][0], got["cameras"][0]
    assert g["eye"] == f32list(c["eye"]) and g["dir"] == f32list(c["dir"]) and g["up"] == f32list(c["up"]) and g["aov"] == float(np.float32(c["aov"]))
    assert g["eye"] == f32list([1.2345678901234, -2.5, 3.1]) and g["up"] == [0.0, 0.0, 1.0] and g["aov"] == 27.5
    assert g["dir"] == f32list([np.float32(0.1) - np.float32(1.2345678901234), np.float32(0.2) - np.float32(-2.5), np.float32(0.3) - np.float32(3.1)])
    # neither JSON nor a camera configuration
    open(os.path.join(tmp, "cam.pvcc"), "w").write("<NotACamera><Proxy/></NotACamera>")
    out = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "error loading camera from" in out.stderr
    os.remove(os.path.join(tmp, "cam.pvcc"))
    out = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "unable to open" in out.stderr


def test_sampler_and_pathlines_operators_parse(tmp_path):
    """examples/noise.state of the reference asks for a GradientSampler; IsoSampler and PathLines keys and defaults
    (GradientSamplerVis.cpp:77-85, IsoSamplerVis.cpp:77-85, PathLinesVis.cpp:46-55,105-114): C++ host == Python front end."""
    tmp = str(tmp_path)
    write_vol(os.path.join(tmp, "radial-oneBall.vol"), scenes.radial_volume("oneBall", 12))
    doc = {"Datasets": [{"name": "scalar", "type": "Volume", "filename": "radial-oneBall.vol"}],
           "Visualizations": [{"annotation": "", "operators": [{"type": "GradientSampler", "dataset": "scalar", "tolerance": 0.1, "volume rendering": False},
                                                                {"type": "IsoSampler", "dataset": "scalar", "isovalue": 0.35},
                                                                {"type": "IsoSampler", "dataset": "scalar"}]}],
           "Cameras": [{"aov": 30.0, "viewpoint": [0.0, 0.0, -20.0], "viewdirection": [0.0, 0.0, 1.0], "viewup": [0.0, 1.0, 0.0]}]}
    state = os.path.join(tmp, "noise.state")
    json.dump(doc, open(state, "w"))
    out = subprocess.run([EXE, "--describe", state], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout, parse_float=lambda t: float(np.float32(t)))
    st = scenes.parse_state(doc, base_dir=tmp)
    ops_g, ops_p = got["visualizations"][0]["operators"], st["visualizations"][0]["operators"]
    assert [o["type"] for o in ops_g] == [o["type"] for o in ops_p] == ["GradientSamplerVis", "IsoSamplerVis", "IsoSamplerVis"]
    assert ops_g[0]["tolerance"] == [ops_p[0]["tolerance"]] == [float(np.float32(0.1))]
    assert ops_g[1]["isovalue"] == [ops_p[1]["isovalue"]] == [float(np.float32(0.35))] and ops_g[2]["isovalue"] == [0.0]
