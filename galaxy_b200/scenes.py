"""State-file front end + synthetic datasets used by the tests and bench.py.

This is the Python mirror of the reference's host-side set-up for the hot path: it parses a
Galaxy `.state` JSON (docs/state_files.md; parsers cited below), builds the Datasets, partitions
them the way the reference does, and instantiates one backend Scene per partition.  It is
backend-agnostic: `backend` is either `galaxy_b200.gpu` (the product's C-ABI) or, in tests
only, `oracle.oracle` (the CPU checker) -- both expose the same Scene/render interface.

All float arithmetic that decides box planes is done in float32 in the reference's association
order so partition planes are bit-identical across backends.
"""
import json
import math
import os

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------
# Datasets
class VolumeDataset:
    """src/data/Volume.cpp: global grid; data[z, y, x] float32 or uint8."""

    kind = "Volume"

    def __init__(self, origin, counts, deltas, data):
        self.origin = np.asarray(origin, f32)
        self.counts = tuple(int(c) for c in counts)  # (nx, ny, nz)
        self.deltas = np.asarray(deltas, f32)
        assert data.shape == (self.counts[2], self.counts[1], self.counts[0])
        self.data = data


class TrianglesDataset:
    """src/data/Triangles.cpp: float3 vertices, float3 normals, per-vertex data, int3 connectivity."""

    kind = "Triangles"

    def __init__(self, verts, normals, data, indices):
        self.verts = np.ascontiguousarray(verts, f32)
        self.normals = np.ascontiguousarray(normals, f32)
        self.data = np.ascontiguousarray(data, f32)
        self.indices = np.ascontiguousarray(indices, np.int32)


class ParticlesDataset:
    """src/data/Particles.cpp: float3 centres + per-particle data."""

    kind = "Particles"

    def __init__(self, centers, data):
        self.centers = np.ascontiguousarray(centers, f32)
        self.data = np.ascontiguousarray(data, f32)


class PathLinesDataset:
    """src/data/PathLines.cpp: poly-lines over a shared point array (a .vtu with VTK_POLY_LINE cells)."""

    kind = "PathLines"

    def __init__(self, points, data, lines):
        self.points = np.ascontiguousarray(points, dtype=f32).reshape(-1, 3)
        self.data = np.ascontiguousarray(data, dtype=f32).reshape(-1)
        self.lines = [np.asarray(l, dtype=np.int64) for l in lines]

    def to_arrays(self):
        """PathLines::load_from_vtkPointSet (PathLines.cpp:73-125): vertices duplicated per cell, in cell order;
        connectivity[j] = index of the first vertex of every segment."""
        ids = np.concatenate(self.lines) if self.lines else np.zeros(0, np.int64)
        conn, k = [], 0
        for l in self.lines:
            conn.extend(range(k, k + len(l) - 1))
            k += len(l)
        return self.points[ids], self.data[ids], np.asarray(conn, np.int32)


def radial_volume(name, n=256):
    """Closed-form fixtures of src/apps/radial.cpp:208-231 as written to .vol by scripts/vti2vol:70-80
    (note the %f rounding of origin/spacing in the header: spacing 0.007843 for n=256)."""
    d = f32(2.0 / (n - 1))
    c = (f32(-1) + np.arange(n, dtype=f32) * d).astype(f32)  # X = -1 + (k+sx)*d in fp32
    Z, Y, X = np.meshgrid(c, c, c, indexing="ij")

    def length(a, b, cc):  # sqrt(a*a + b*b + c*c), fp32, left to right
        return np.sqrt(((a * a).astype(f32) + (b * b).astype(f32)).astype(f32) + (cc * cc).astype(f32)).astype(f32)

    if name == "oneBall":
        v = length(X, Y, Z)
    elif name == "eightBalls":
        h = f32(0.5)
        v = length((np.abs(X) - h).astype(f32), (np.abs(Y) - h).astype(f32), (np.abs(Z) - h).astype(f32))
    elif name in ("xramp", "yramp", "zramp"):
        v = {"xramp": X, "yramp": Y, "zramp": Z}[name].astype(f32)
    else:
        raise KeyError(name)
    sp = float("%f" % (2.0 / (n - 1)))
    return VolumeDataset([-1.0, -1.0, -1.0], (n, n, n), [sp, sp, sp], np.ascontiguousarray(v))


def load_vol_file(path):
    """Volume::local_import, `.vol` / `.json` headers (src/data/Volume.cpp:195-297)."""
    d = os.path.dirname(path)
    if path.endswith(".vol"):
        tok = open(path).read().split()
        typ = tok[0]
        origin = [float(x) for x in tok[1:4]]
        counts = [int(x) for x in tok[4:7]]
        deltas = [float(x) for x in tok[7:10]]
        raw = tok[10]
    else:
        doc = json.load(open(path))
        typ, origin, counts, deltas, raw = doc["type"], doc["origin"], doc["counts"], doc["delta"], doc["rawdata"]
    raw = raw if raw.startswith("/") else os.path.join(d, raw)
    dt = np.float32 if typ == "float" else np.uint8
    data = np.fromfile(raw, dtype=dt).reshape(counts[2], counts[1], counts[0])
    return VolumeDataset(origin, counts, deltas, data)


def default_data_provider(data_dir=None, n=256):
    cache = {}

    def provider(ds):
        fn = ds["filename"]
        if fn in cache:
            return cache[fn]
        if data_dir and os.path.exists(os.path.join(data_dir, fn)):
            v = load_vol_file(os.path.join(data_dir, fn))
        elif fn.startswith("radial-") and fn.endswith(".vol"):
            v = radial_volume(fn[len("radial-"):-len(".vol")], n)
        else:
            raise FileNotFoundError(fn)
        cache[fn] = v
        return v

    return provider


# ------------------------------------------------------------------------------------------------
# state file -> description
def parse_lighting(v):
    """Lighting::LoadStateFromValue (src/renderer/Lighting.cpp:59-115); defaults Lighting.ispc:29-44."""
    L = dict(lights=[[1.0, 1.0, 1.0]], types=[2], n_ao=0, ao_radius=1.0, shadows=False, Ka=0.5, Kd=0.5)
    if v is None:
        return L
    if "Sources" in v:
        L["lights"], L["types"] = [], []
        for s in v["Sources"]:
            x, y, z = f32(s[0]), f32(s[1]), f32(s[2])
            t = 0 if len(s) == 3 else int(s[3])
            if t == 0:
                d = f32(np.sqrt(f32(f32(x * x + y * y) + z * z)))
                if d == 0:
                    x = y = z = f32(0.577350)
                else:
                    x, y, z = f32(x / d), f32(y / d), f32(z / d)
            L["lights"].append([float(x), float(y), float(z)])
            L["types"].append(t)
    if "shadows" in v:
        L["shadows"] = bool(v["shadows"])
    L["n_ao"] = int(v.get("ao count", 0))
    L["ao_radius"] = float(f32(v.get("ao radius", 1)))
    L["Ka"] = float(f32(v.get("Ka", 0.5)))
    L["Kd"] = float(f32(v.get("Kd", 0.5)))
    return L


def _resolve(name, base_dir):
    if name.startswith("/") or not base_dir:
        return name
    p = os.path.join(base_dir, name)
    return p if os.path.exists(p) else name


def _strtof(s):
    """the float a C++ stream extraction / strtof reads (no double rounding)"""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    return float(libc.strtof(s.strip().encode(), None))


def _camera_from_pvcc(path):
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    proxy = root.find("Proxy") if root.tag == "PVCameraConfiguration" else root.find("PVCameraConfiguration/Proxy")
    if proxy is None:
        raise ValueError("error loading camera from " + path)
    cam = dict(eye=[0.0] * 3, up=[0.0, 1.0, 0.0], aov=30.0, annotation="")
    center = [f32(0)] * 3
    for prop in proxy.findall("Property"):
        values = [f32(-1e32)] * 3
        for el in prop.findall("Element"):
            values[int(el.get("index"))] = f32(_strtof(el.get("value")))
        name = prop.get("name")
        if name == "CameraPosition":
            cam["eye"] = [float(x) for x in values]
        elif name == "CameraFocalPoint":
            center = values
        elif name == "CameraViewUp":
            cam["up"] = [float(x) for x in values]
        elif name == "CameraViewAngle":
            cam["aov"] = float(values[0])
    cam["dir"] = [float(f32(center[k] - f32(cam["eye"][k]))) for k in range(3)]
    return cam


def parse_camera(v, base_dir=""):
    """Camera::LoadFromJSON (src/renderer/Camera.cpp:165-281): an object, or the name of a ParaView camera file in JSON form."""
    if isinstance(v, str):
        try:
            doc = json.load(open(_resolve(v, base_dir)))
        except ValueError:  # Camera.cpp:182-191: not JSON -> ParaView's own XML form (.pvcc), Camera::LoadFromPVCC (:118-163)
            return _camera_from_pvcc(_resolve(v, base_dir))
        cam = dict(eye=[0.0] * 3, up=[0.0, 1.0, 0.0], aov=30.0, annotation="")
        center = [f32(0)] * 3
        for p in doc["PVCameraConfiguration"]["Proxy"]["Property"]:
            name = p.get("@name")
            el = p.get("Element")
            if name == "CameraPosition":
                cam["eye"] = [float(f32(float(el[k]["@value"]))) for k in range(3)]
            elif name == "CameraFocalPoint":
                center = [f32(float(el[k]["@value"])) for k in range(3)]
            elif name == "CameraViewUp":
                cam["up"] = [float(f32(float(el[k]["@value"]))) for k in range(3)]
            elif name == "CameraViewAngle":
                cam["aov"] = float(f32(float(el["@value"])))
        cam["dir"] = [float(f32(center[k] - f32(cam["eye"][k]))) for k in range(3)]
        return cam
    eye = np.asarray(v["viewpoint"], f32)
    if "viewdirection" in v:
        d = np.asarray(v["viewdirection"], f32)
    else:
        d = (np.asarray(v["viewcenter"], np.float64) - eye.astype(np.float64)).astype(f32)
    return dict(eye=[float(x) for x in eye], dir=[float(x) for x in d], up=[float(f32(x)) for x in v["viewup"]], aov=float(f32(v["aov"])),
                annotation=v.get("annotation", ""))


def resample_tf(cmap, omap):
    """MappedVis::local_commit (src/renderer/MappedVis.cpp:277-338): 256-entry tables; x in double."""
    cmap = np.asarray(cmap, f32).reshape(-1, 4)
    omap = np.asarray(omap, f32).reshape(-1, 2)

    def table(xs, ys):
        out = np.zeros((256,) + ys.shape[1:], f32)
        i0, i1 = 0, 1
        xmin, xmax = xs[0], xs[-1]
        for i in range(256):
            x = f32(float(xmin) + (i / 255.0) * float(f32(xmax - xmin)))
            if x > xmax:
                x = xmax
            while xs[i1] < x:
                i0 += 1
                i1 += 1
            d = f32(f32(x - xs[i0]) / f32(xs[i1] - xs[i0]))
            out[i] = (ys[i0] + d * (ys[i1] - ys[i0]).astype(f32)).astype(f32)
        return out

    return table(cmap[:, 0], cmap[:, 1:4]), table(omap[:, 0], omap[:, 1])


def parse_operator(v, base_dir=""):
    """Vis/MappedVis/VolumeVis/ParticlesVis::LoadFromJSON (Vis.cpp:126-146, MappedVis.cpp:86-203,
    VolumeVis.cpp:119-163, ParticlesVis.cpp:104-118)."""
    t = v["type"]
    if not t.endswith("Vis"):
        t += "Vis"  # Visualization.cpp:318-320
    op = dict(type=t, dataset=v["dataset"])
    cmap = [[0.0, 0.4, 0.4, 0.4], [1.0, 1.0, 1.0, 1.0]]  # MappedVis.cpp:58-62
    omap = [[0.0, 1.0], [1.0, 1.0]]
    m = v.get("transfer function", v.get("colormap"))
    if isinstance(m, str) and m not in ("", "default"):  # MappedVis.cpp:104-166: ParaView colormap export
        doc = json.load(open(_resolve(m, base_dir)))
        cm = doc[0] if isinstance(doc, list) else doc
        pts = cm.get("Points")
        omap = [[float(pts[i]), float(pts[i + 1])] for i in range(0, len(pts) - 1, 4)] if pts is not None else [[0.0, 1.0], [1.0, 1.0]]
        rgb = cm["RGBPoints"]
        cmap = [[float(x) for x in rgb[i:i + 4]] for i in range(0, len(rgb) - 3, 4)]
    elif m is not None and not isinstance(m, str):
        cmap = [[float(x) for x in row[:4]] for row in m]
        if "opacitymap" in v:
            omap = [[float(x) for x in row[:2]] for row in v["opacitymap"]]
    op["colormap"], op["opacitymap"] = cmap, omap
    op["data_range"] = [float(f32(x)) for x in v["data range"]] if "data range" in v else None
    if t == "VolumeVis":
        op["isovalues"] = [float(f32(x)) for x in v.get("isovalues", [])]
        if "slices" in v:
            op["slices"] = [[float(f32(x)) for x in s[:4]] for s in v["slices"]]
        elif "plane" in v:
            op["slices"] = [[float(f32(x)) for x in v["plane"][:4]]]
        else:
            op["slices"] = []
        op["volume_render"] = bool(v.get("volume rendering", False))
    elif t == "GradientSamplerVis":
        # src/sampler/GradientSamplerVis.cpp:77-85 (initialize: GradientSamplerVis.ispc:74-84 -> 0)
        op["tolerance"] = float(f32(v.get("tolerance", 0.0)))
    elif t == "IsoSamplerVis":
        # src/sampler/IsoSamplerVis.cpp:77-85
        op["isovalue"] = float(f32(v.get("isovalue", 0.0)))
    elif t == "PathLinesVis":
        # PathLinesVis::initialize / LoadFromJSON (PathLinesVis.cpp:46-55,105-114)
        op["radius0"] = float(f32(v.get("radius0", -1.0)))
        op["radius1"] = float(f32(v.get("radius1", 1.0)))
        op["value0"] = float(f32(v.get("value0", 0.0)))
        op["value1"] = float(f32(v.get("value1", 1.0)))
    elif t == "ParticlesVis":
        # ParticlesVis::initialize / LoadFromJSON (ParticlesVis.cpp:46-55,104-118)
        op["radius0"] = float(v.get("radius0", 0.025))
        op["radius1"] = float(v.get("radius1", 0.0))
        op["value0"] = float(v.get("value0", 0.0))
        op["value1"] = float(v.get("value1", 0.0))
        if "radius" in v:
            op["radius0"], op["radius1"], op["value0"], op["value1"] = float(v["radius"]), 0.0, 0.0, 0.0
    return op


def parse_state(doc, base_dir=""):
    if isinstance(doc, str):
        doc = json.loads(doc)
    st = dict(datasets=doc.get("Datasets", []), epsilon=float(doc.get("Renderer", {}).get("epsilon", 0.001)))
    vs = doc.get("Visualization", doc.get("Visualizations"))
    if not isinstance(vs, list):
        vs = [vs]
    st["visualizations"] = [dict(annotation=v.get("annotation", ""), lighting=parse_lighting(v.get("Lighting", v.get("lighting"))),
                                 operators=[parse_operator(o, base_dir) for o in v["operators"]]) for v in vs]
    cs = doc.get("Cameras", doc.get("Camera"))
    if not isinstance(cs, list):
        cs = [cs]
    st["cameras"] = [parse_camera(c, base_dir) for c in cs]
    return st


# ------------------------------------------------------------------------------------------------
# partitioning
def factor(ijk):
    """src/data/Volume.cpp:88-122."""
    if ijk == 1:
        return (1, 1, 1)
    f, mm = (1, 1, ijk), ijk + 3
    for i in range(1, (ijk >> 1) + 1):
        jk = ijk // i
        if ijk == i * jk:
            for j in range(1, (jk >> 1) + 1):
                k = jk // j
                if jk == j * k and i + j + k < mm:
                    mm, f = i + j + k, (i, j, k)
    return f


def partition(factors, grid):
    """src/data/Volume.cpp:133-172 -> list of dict(ijk, offsets, counts, goffsets, gcounts), rank order."""
    n = [g - 2 for g in grid]
    d = [n[a] // factors[a] for a in range(3)]
    parts = []
    for k in range(factors[2]):
        for j in range(factors[1]):
            for i in range(factors[0]):
                ijk = (i, j, k)
                off = [1 + ijk[a] * d[a] for a in range(3)]
                cnt = [1 + ((n[a] - off[a]) if ijk[a] == factors[a] - 1 else d[a]) for a in range(3)]
                parts.append(dict(ijk=ijk, offsets=off, counts=cnt, goffsets=[o - 1 for o in off], gcounts=[c + 2 for c in cnt]))
    return parts


def neighbors_of(ijk, factors):
    """src/data/Volume.cpp:358-377."""
    def rank(i, j, k):
        return i + j * factors[0] + k * factors[0] * factors[1]
    i, j, k = ijk
    return [rank(i - 1, j, k) if i > 0 else -1, rank(i + 1, j, k) if i < factors[0] - 1 else -1,
            rank(i, j - 1, k) if j > 0 else -1, rank(i, j + 1, k) if j < factors[1] - 1 else -1,
            rank(i, j, k - 1) if k > 0 else -1, rank(i, j, k + 1) if k < factors[2] - 1 else -1]


def volume_boxes(vol, part):
    """global/local Box of a partition (src/data/Volume.cpp:379-390, Box.cpp:69-80), fp32."""
    o, d = vol.origin, vol.deltas
    go = (o + d).astype(f32)
    gc = [c - 2 for c in vol.counts]
    gmax = np.array([go[a] + f32(gc[a] - 1) * d[a] for a in range(3)], f32)
    lo = np.array([o[a] + f32(part["offsets"][a]) * d[a] for a in range(3)], f32)
    lmax = np.array([lo[a] + f32(part["counts"][a] - 1) * d[a] for a in range(3)], f32)
    return go, gmax, lo, lmax


def geometry_extents(nparts, origin=-1.0, counts=256, spacing=None):
    """Partition extents for geometry datasets: scripts/createPartitionDoc.py:76-106 (float64 there;
    stored as fp32 by Geometry::get_partitioning, src/data/Geometry.cpp:283-292)."""
    if spacing is None:
        spacing = 2.0 / (counts - 1)
    fac = factor(nparts)
    n = [int((counts - 2) / f) for f in fac]
    IJK = [[1 + i * n[j] for i in range(fac[j])] + [counts - 2] for j in range(3)]
    IJK = [list(zip(a[:-1], a[1:])) for a in IJK]
    out = []
    for kz in IJK[2]:
        for jy in IJK[1]:
            for ix in IJK[0]:
                out.append([origin + ix[0] * spacing, origin + ix[1] * spacing, origin + jy[0] * spacing, origin + jy[1] * spacing,
                            origin + kz[0] * spacing, origin + kz[1] * spacing])
    return np.asarray(out, f32), fac


def geometry_neighbors(extents, r):
    """src/data/Geometry.cpp:296-344."""
    l = extents[r]
    nb = [-1] * 6
    for i, e in enumerate(extents):
        LAST = lambda a: e[2 * a + 1] == l[2 * a]
        NEXT = lambda a: e[2 * a] == l[2 * a + 1]
        EQ = lambda a: e[2 * a] == l[2 * a]
        if LAST(0) and EQ(1) and EQ(2): nb[0] = i
        if NEXT(0) and EQ(1) and EQ(2): nb[1] = i
        if EQ(0) and LAST(1) and EQ(2): nb[2] = i
        if EQ(0) and NEXT(1) and EQ(2): nb[3] = i
        if EQ(0) and EQ(1) and LAST(2): nb[4] = i
        if EQ(0) and EQ(1) and NEXT(2): nb[5] = i
    return nb


def clip_triangles(ds, ext, ghost=0.1):
    """Partition a mesh: keep triangles with any vertex inside the ghost-extended extent
    (scripts/partitionVTUs.vpy:35,66-67), compacting the vertex arrays."""
    lo = np.array([ext[0], ext[2], ext[4]], f32) - f32(ghost)
    hi = np.array([ext[1], ext[3], ext[5]], f32) + f32(ghost)
    inside = np.all((ds.verts >= lo) & (ds.verts <= hi), axis=1)
    keep = inside[ds.indices].any(axis=1)
    tri = ds.indices[keep]
    used = np.zeros(len(ds.verts), bool)
    used[tri.ravel()] = True
    remap = np.cumsum(used, dtype=np.int64) - 1
    return TrianglesDataset(ds.verts[used], ds.normals[used], ds.data[used], remap[tri].astype(np.int32))


def clip_particles(ds, ext, ghost=0.1):
    lo = np.array([ext[0], ext[2], ext[4]], f32) - f32(ghost)
    hi = np.array([ext[1], ext[3], ext[5]], f32) + f32(ghost)
    inside = np.all((ds.centers >= lo) & (ds.centers <= hi), axis=1)
    return ParticlesDataset(ds.centers[inside], ds.data[inside])


def helix_pathlines(n_lines, n_verts=12, seed=21):
    """A synthetic PathLines dataset: helical poly-lines inside [-1,1]^3, data = |p| (what a stream tracer through a
    rotating field leaves; tests/create_data_driven_datasets.vpy makes the reference's with vtkStreamTracer)."""
    rng = np.random.default_rng(seed)
    pts, lines, k = [], [], 0
    for _ in range(n_lines):
        t = np.linspace(0, 3 + rng.uniform(0, 2), n_verts)
        c = rng.uniform(-.4, .4, 3)
        pts.append(np.stack([c[0] + 0.45 * np.cos(t), c[1] + 0.45 * np.sin(t), c[2] + 0.15 * t - 0.3], 1))
        lines.append(list(range(k, k + n_verts)))
        k += n_verts
    pts = np.concatenate(pts).astype(f32)
    return PathLinesDataset(pts, np.linalg.norm(pts, axis=1), lines)


def clip_pathlines(ds, ext, ghost=0.1):
    """Partition poly-lines as scripts/partitionVTUs.vpy:116-147 does: every run of vertices inside the ghost-extended
    extent, plus one vertex at either end, becomes a poly-line of the piece; points are compacted in order of first use."""
    lower = np.array([ext[0], ext[2], ext[4]], np.float64) - ghost
    upper = np.array([ext[1], ext[3], ext[5]], np.float64) + ghost
    pmap, order, lines = {}, [], []
    for cell in ds.lines:
        pts = ds.points[cell].astype(np.float64)
        insiders = np.all(np.hstack((pts <= upper, pts >= lower)), axis=1)
        linelen, offset, kk = len(insiders), 0, insiders
        while len(kk) > 0:
            s = int(np.argmax(kk))
            if not kk[s]:
                break
            n = int(np.argmax(np.logical_not(kk[s:])))
            if n == 0:
                n = len(kk[s:]) + 1
            n = s + n - 1
            s, n = offset + s, offset + n
            if s > 0:
                s -= 1
            if n < linelen - 1:
                n += 1
            seg = cell[s:n + 1]
            offset = n + 1
            kk = insiders[offset:]
            for i in seg:
                if int(i) not in pmap:
                    pmap[int(i)] = len(order)
                    order.append(int(i))
            if len(seg) >= 2:
                lines.append([pmap[int(i)] for i in seg])
    order = np.asarray(order, np.int64)
    return PathLinesDataset(ds.points[order], ds.data[order], lines)


# ------------------------------------------------------------------------------------------------
def build_partitions(backend, vis, datasets, nparts, geom_extents=None, only_rank=None, **scene_kw):
    """Instantiate one backend Scene per partition for Visualization `vis` (a parse_state entry).
    datasets: dict name -> *Dataset.  Returns list of scenes (or [scene] for only_rank)."""
    fac = factor(nparts)
    ds_ids = {name: i for i, name in enumerate(sorted(datasets))}
    scenes = []
    first = datasets[vis["operators"][0]["dataset"]]
    if geom_extents is None and first.kind != "Volume":
        geom_extents, _ = geometry_extents(nparts)
    for r in range(nparts):
        if only_rank is not None and r != only_rank:
            continue
        sc = backend.Scene(**scene_kw)
        boxes_set = False
        for op in vis["operators"]:
            ds = datasets[op["dataset"]]
            colors, opac = resample_tf(op["colormap"], op["opacitymap"])
            if op["data_range"] is not None:
                lo, hi = op["data_range"]
            else:
                lo, hi = op["colormap"][0][0], op["colormap"][-1][0]
            if ds.kind == "Volume":
                part = partition(fac, ds.counts)[r]
                if not boxes_set:  # Visualization::local_commit: boxes/neighbours of the FIRST vis (Visualization.cpp:139-160)
                    gmin, gmax, lmin, lmax = volume_boxes(ds, part)
                    sc.set_partition(gmin, gmax, lmin, lmax, neighbors_of(part["ijk"], fac))
                    boxes_set = True
                go, gc = part["goffsets"], part["gcounts"]
                brick = np.ascontiguousarray(ds.data[go[2]:go[2] + gc[2], go[1]:go[1] + gc[1], go[0]:go[0] + gc[0]])
                origin = np.array([ds.origin[a] + f32(go[a]) * ds.deltas[a] for a in range(3)], f32)  # Volume.h:124-129
                if op["type"] in ("GradientSamplerVis", "IsoSamplerVis"):  # a sampling Visualization (src/sampler)
                    kind = op["type"][:-3]
                    sc.add_sampler_vis(ds_ids[op["dataset"]], gc, origin, ds.deltas, brick, kind,
                                       op["tolerance"] if kind == "GradientSampler" else op["isovalue"])
                    continue
                sc.add_volume_vis(ds_ids[op["dataset"]], gc, origin, ds.deltas, brick, op["slices"], op["isovalues"], op["volume_render"],
                                  colors, opac, lo, hi)
            else:
                if geom_extents is None:  # geometry after a volume operator: its own partition document
                    geom_extents, _ = geometry_extents(nparts)
                ext = geom_extents[r]
                if not boxes_set:
                    g = geom_extents
                    gmin = [g[:, 0].min(), g[:, 2].min(), g[:, 4].min()]
                    gmax = [g[:, 1].max(), g[:, 3].max(), g[:, 5].max()]
                    sc.set_partition(gmin, gmax, [ext[0], ext[2], ext[4]], [ext[1], ext[3], ext[5]], geometry_neighbors(geom_extents, r))
                    boxes_set = True
                if ds.kind == "Triangles":
                    p = ds if nparts == 1 else clip_triangles(ds, ext)
                    sc.add_triangles_vis(p.verts, p.normals, p.data, p.indices, colors, opac, lo, hi)
                elif ds.kind == "Particles":
                    p = ds if nparts == 1 else clip_particles(ds, ext)
                    sc.add_particles_vis(p.centers, p.data, op.get("radius0", 0.025), op.get("radius1", 0.0), op.get("value0", 0.0),
                                         op.get("value1", 0.0), colors, opac, lo, hi)
                elif ds.kind == "PathLines":
                    p = ds if nparts == 1 else clip_pathlines(ds, ext)
                    pv, pd, pc = p.to_arrays()
                    sc.add_pathlines_vis(pv, pd, pc, op.get("radius0", -1.0), op.get("radius1", 1.0), op.get("value0", 0.0),
                                         op.get("value1", 1.0), colors, opac, lo, hi)
                else:
                    raise NotImplementedError(ds.kind)
        sc.commit()
        scenes.append(sc)
    return scenes


def load_datasets(state, provider):
    out = {}
    for ds in state["datasets"]:
        if ds["type"] != "Volume":
            raise NotImplementedError("file-based %s datasets" % ds["type"])
        out[ds["name"]] = provider(ds)
    return out


# ------------------------------------------------------------------------------------------------
# synthetic benchmark scenes (generators are ours, closed-form, seeded)
def _pcg_hash(x):
    x = (x.astype(np.uint64) * np.uint64(747796405) + np.uint64(2891336453)) & np.uint64(0xFFFFFFFF)
    w = (((x >> ((x >> np.uint64(28)) + np.uint64(4))) ^ x) * np.uint64(277803737)) & np.uint64(0xFFFFFFFF)
    return ((w >> np.uint64(22)) ^ w) & np.uint64(0xFFFFFFFF)


def value_noise3(p, seed):
    """Trilinear value noise on the integer lattice; p (...,3) float64 -> [0,1)."""
    pf = np.floor(p)
    fr = p - pf
    fr = fr * fr * (3 - 2 * fr)
    ip = pf.astype(np.int64)
    out = 0.0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                h = (ip[..., 0] + dx) * 73856093 ^ (ip[..., 1] + dy) * 19349663 ^ (ip[..., 2] + dz) * 83492791 ^ seed * 2654435761
                v = _pcg_hash(h & 0xFFFFFFFF).astype(np.float64) / 4294967296.0
                w = (fr[..., 0] if dx else 1 - fr[..., 0]) * (fr[..., 1] if dy else 1 - fr[..., 1]) * (fr[..., 2] if dz else 1 - fr[..., 2])
                out = out + w * v
    return out


def fbm3(p, seed, octaves=5):
    a, f, s, norm = 0.5, 1.0, 0.0, 0.0
    for o in range(octaves):
        s = s + a * value_noise3(p * f, seed + o)
        norm += a
        a *= 0.5
        f *= 2.0
    return s / norm


# ---- the noise volume of the volume workloads C3 / C4 ---------------------------------------------------------------
NOISE_SEED = 7


def noise_volume_block(n, z0, z1, y0, y1, x0, x1, device="cpu"):
    """voxels [z0:z1, y0:y1, x0:x1] of the n^3 noise volume: v = eightBalls(p) + 0.15 * fbm(8 p), p in [-1,1]^3 (origin -1, spacing
    2/(n-1)), fbm = 5 octaves of trilinear value noise whose lattice values are PCG32 hashes (seed 7): exactly fbm3() above, written
    with torch so that a 2048^3 volume is synthesised slab by slab on the device.  float64 arithmetic, float32 result."""
    import torch
    dev = torch.device(device)
    sp = 2.0 / (n - 1)
    cz = (-1.0 + torch.arange(z0, z1, device=dev, dtype=torch.float64) * sp).view(-1, 1, 1)
    cy = (-1.0 + torch.arange(y0, y1, device=dev, dtype=torch.float64) * sp).view(1, -1, 1)
    cx = (-1.0 + torch.arange(x0, x1, device=dev, dtype=torch.float64) * sp).view(1, 1, -1)
    eb = torch.sqrt((cx.abs() - 0.5) ** 2 + (cy.abs() - 0.5) ** 2 + (cz.abs() - 0.5) ** 2)
    M = 0xFFFFFFFF

    def pcg(x):
        x = (x * 747796405 + 2891336453) & M
        w = (((x >> ((x >> 28) + 4)) ^ x) * 277803737) & M
        return ((w >> 22) ^ w) & M

    def axis(c, f):
        p = c * f
        pf = torch.floor(p)
        fr = p - pf
        return pf.to(torch.int64), fr * fr * (3 - 2 * fr)

    s, a, f, norm = 0.0, 0.5, 1.0, 0.0
    for o in range(5):
        (ix, fx), (iy, fy), (iz, fz) = axis(cx, 8.0 * f), axis(cy, 8.0 * f), axis(cz, 8.0 * f)
        val = 0.0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    h = ((ix + dx) * 73856093) ^ ((iy + dy) * 19349663) ^ ((iz + dz) * 83492791) ^ ((NOISE_SEED + o) * 2654435761)
                    v = pcg(h & M).to(torch.float64) / 4294967296.0
                    w = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                    val = val + w * v
        s = s + a * val
        norm += a
        a *= 0.5
        f *= 2.0
    return (eb + 0.15 * (s / norm)).to(torch.float32)


class _LazyNoiseData:
    """data[z, y, x] of a noise volume, synthesised on demand block by block (what build_partitions slices out per partition)"""

    def __init__(self, n, device):
        self.n, self.device, self.shape = n, device, (n, n, n)

    def __getitem__(self, key):
        zs, ys, xs = [k.indices(self.n) for k in key]
        out = np.empty((zs[1] - zs[0], ys[1] - ys[0], xs[1] - xs[0]), np.float32)
        for z in range(zs[0], zs[1], 32):   # slabs of 32 planes: ~1 GB of float64 temporaries at 2048^2 per plane
            z1 = min(z + 32, zs[1])
            out[z - zs[0]:z1 - zs[0]] = noise_volume_block(self.n, z, z1, ys[0], ys[1], xs[0], xs[1], self.device).cpu().numpy()
        return out


def noise_volume(n, device="cpu"):
    """the C3/C4 volume as a VolumeDataset whose voxels are generated when a partition asks for its brick"""
    sp = 2.0 / (n - 1)
    ds = VolumeDataset.__new__(VolumeDataset)
    ds.origin = np.asarray([-1.0, -1.0, -1.0], f32)
    ds.counts = (n, n, n)
    ds.deltas = np.asarray([sp, sp, sp], f32)
    ds.data = _LazyNoiseData(n, device)
    return ds


def eightballs_mesh(n_lat, n_lon, seed=11, bump=0.05):
    """C5 mesh: 8 UV-spheres centred (+-.5,+-.5,+-.5), radius 0.3*(1+bump*fbm(6*dir)),
    n_lat x n_lon quads each split in two (polar quads keep one zero-area triangle), data = |p|."""
    th = np.linspace(0.0, math.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * math.pi, n_lon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1)
    d[:, -1] = d[:, 0]  # close the seam exactly
    r = 0.3 * (1.0 + bump * (2.0 * fbm3(6.0 * d + 17.0, seed) - 1.0))
    r[0, :] = r[0, 0]
    r[-1, :] = r[-1, 0]
    pts = (d * r[..., None]).reshape(-1, 3)
    nrm = d.reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n_lat), np.arange(n_lon), indexing="ij")
    a = (i * (n_lon + 1) + j).ravel()
    b = a + 1
    c = a + (n_lon + 1)
    e = c + 1
    tris = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, e], 1)], 0).astype(np.int64)
    V, N, I = [], [], []
    for k, (sx, sy, sz) in enumerate([(x, y, z) for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)]):
        V.append(pts + np.array([sx, sy, sz]))
        N.append(nrm)
        I.append(tris + k * len(pts))
    V = np.concatenate(V).astype(f32)
    N = np.concatenate(N).astype(f32)
    I = np.concatenate(I).astype(np.int32)
    D = np.sqrt((V.astype(np.float64) ** 2).sum(1)).astype(f32)
    return TrianglesDataset(V, N, D, I)


# ------------------------------------------------------------------------------------------------
# C5: "eightBalls-100M" triangle scene, its Visualization, lighting and camera
C5_FULL = (1768, 3536)  # n_lat x n_lon -> 8 x 12,503,296 = 100,026,368 triangles
C5_CENTERS = [(x, y, z) for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)]


def c5_vis():
    return dict(annotation="", lighting=parse_lighting({"Sources": [[1, 2, -3, 0]], "shadows": True, "Ka": 0.3, "Kd": 0.7, "ao count": 8,
                                                        "ao radius": 0.3}),
                operators=[dict(type="TrianglesVis", dataset="mesh", colormap=[[0.3, 0.0, 0.0, 1.0], [1.0, 1.0, 1.0, 0.0]],
                                opacitymap=[[0.0, 1.0], [1.0, 1.0]], data_range=None)])


def c5_camera():
    return parse_camera({"viewpoint": [3, 2, -4], "viewcenter": [0, 0, 0], "viewup": [0, 1, 0], "aov": 30})


def eightballs_mesh_single(n_lat, n_lon, seed=11, bump=0.05):
    th = np.linspace(0.0, math.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * math.pi, n_lon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1)
    d[:, -1] = d[:, 0]
    r = 0.3 * (1.0 + bump * (2.0 * fbm3(6.0 * d + 17.0, seed) - 1.0))
    r[0, :] = r[0, 0]
    r[-1, :] = r[-1, 0]
    pts = (d * r[..., None]).reshape(-1, 3)
    nrm = d.reshape(-1, 3).astype(f32)
    i, j = np.meshgrid(np.arange(n_lat, dtype=np.int32), np.arange(n_lon, dtype=np.int32), indexing="ij")
    a = (i * np.int32(n_lon + 1) + j).ravel()
    b = a + np.int32(1)
    c = a + np.int32(n_lon + 1)
    e = c + np.int32(1)
    tris = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, e], 1)], 0).astype(np.int32)
    return dict(pts=pts, nrm=nrm, tris=tris)


def c5_partition_mesh(n_lat, n_lon, nparts, rank, seed=11, ghost=0.1):
    """The triangles of partition `rank`: every bumpy sphere whose box touches the ghost-extended
    extent (no sphere straddles a partition plane for nparts in 1,2,4,8, so this equals the
    reference's cell-level clipping, scripts/partitionVTUs.vpy:35,66-67).  Returns (dataset, extents)."""
    one = eightballs_mesh_single(n_lat, n_lon, seed)
    ext, _ = geometry_extents(nparts)
    e = ext[rank]
    lo = np.array([e[0], e[2], e[4]]) - ghost
    hi = np.array([e[1], e[3], e[5]]) + ghost
    V, N, I, D = [], [], [], []
    nv = 0
    for c in C5_CENTERS:
        c = np.asarray(c)
        if nparts > 1 and (np.any(c + 0.34 < lo) or np.any(c - 0.34 > hi)):
            continue
        v = (one["pts"] + c).astype(f32)
        V.append(v)
        N.append(one["nrm"])
        I.append(one["tris"] + np.int32(nv))
        D.append(np.sqrt((v.astype(np.float64) ** 2).sum(1)).astype(f32))
        nv += len(v)
    if not V:
        return TrianglesDataset(np.zeros((0, 3), f32), np.zeros((0, 3), f32), np.zeros((0,), f32), np.zeros((0, 3), np.int32)), ext
    return TrianglesDataset(np.concatenate(V), np.concatenate(N), np.concatenate(D), np.concatenate(I)), ext
