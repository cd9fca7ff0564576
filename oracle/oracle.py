"""ctypes binding of the CPU oracle (oracle/libgxy_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/gxy_oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
product package galaxy_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_LIGHTS = 16


class Lighting(C.Structure):
    _fields_ = [("n_lights", C.c_int), ("lights", (C.c_float * 3) * MAX_LIGHTS), ("types", C.c_int * MAX_LIGHTS),
                ("n_ao", C.c_int), ("ao_radius", C.c_float), ("shadows", C.c_int), ("Ka", C.c_float), ("Kd", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3), ("aov", C.c_float)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays",
                                            "traced_rays", "volume_samples", "orphan_pixels", "waves")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    so = os.path.join(_HERE, "libgxy_oracle.so")
    src = os.path.join(_HERE, "gxy_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgxy_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.gxo_scene_create.restype = vp
        L.gxo_scene_destroy.argtypes = [vp]
        L.gxo_scene_set_partition.argtypes = [vp, fp, fp, fp, fp, ip]
        L.gxo_scene_add_volume_vis.argtypes = [vp, C.c_int, ip, fp, fp, C.c_int, vp, C.c_int, fp, C.c_int, fp, C.c_int, fp, fp,
                                               C.c_float, C.c_float]
        L.gxo_scene_add_triangles_vis.argtypes = [vp, C.c_int, fp, fp, fp, C.c_int, ip, fp, fp, C.c_float, C.c_float]
        L.gxo_scene_add_particles_vis.argtypes = [vp, C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp,
                                                  C.c_float, C.c_float]
        L.gxo_scene_add_pathlines_vis.argtypes = [vp, C.c_int, fp, fp, C.c_int, ip, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp,
                                                  C.c_float, C.c_float]
        L.gxo_build_curves.argtypes = [C.c_int, fp, fp, C.c_int, ip, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        L.gxo_render_progressive.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float,
                                             C.c_int, C.c_int, fp, ip, ip, C.POINTER(Stats)]
        L.gxo_scene_add_sampler_vis.argtypes = [vp, C.c_int, ip, fp, fp, C.c_int, vp, C.c_int, C.c_float]
        L.gxo_sample_raylist.argtypes = [vp, fp, C.c_int, C.c_int]
        L.gxo_sample.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Stats)]
        L.gxo_scene_samples.argtypes = [vp, C.POINTER(fp)]
        L.gxo_scene_samples.restype = C.c_longlong
        L.gxo_scene_commit.argtypes = [vp]
        L.gxo_resample_tf.argtypes = [C.c_int, fp, C.c_int, fp, fp, fp]
        L.gxo_resolve_lights.argtypes = [C.POINTER(Lighting), C.POINTER(Camera), C.POINTER(Lighting)]
        L.gxo_trace_raylist.argtypes = [vp, C.POINTER(Lighting), fp, C.c_int, C.c_int, C.c_float, ip]
        L.gxo_trace_raylist.restype = C.c_longlong
        L.gxo_fetch_secondary.argtypes = [vp, fp, C.c_int]
        L.gxo_classify.argtypes = [vp, fp, C.c_int, C.c_int]
        L.gxo_generate_rays.argtypes = [vp, C.POINTER(Camera), C.c_int, C.c_int, fp, C.c_int]
        L.gxo_render.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float,
                                 C.c_int, C.c_int, fp, C.POINTER(Stats)]
        L.gxo_fb_to_rgba8.argtypes = [fp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
        L.gxo_factor.argtypes = [C.c_int, ip]
        L.gxo_partition.argtypes = [C.c_int, ip, ip, ip]
        L.gxo_intersect.argtypes = [vp, C.c_int, fp, fp, fp, fp, ip, fp]
        L.gxo_scene_set_intersector.argtypes = [vp, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def make_lighting(d):
    """d: dict(lights=[[x,y,z],..], types=[..], n_ao, ao_radius, shadows, Ka, Kd)"""
    L = Lighting()
    L.n_lights = len(d["lights"])
    for i, (l, t) in enumerate(zip(d["lights"], d["types"])):
        for k in range(3):
            L.lights[i][k] = l[k]
        L.types[i] = t
    L.n_ao, L.ao_radius, L.shadows, L.Ka, L.Kd = d["n_ao"], d["ao_radius"], int(d["shadows"]), d["Ka"], d["Kd"]
    return L


def lighting_to_dict(L):
    return dict(lights=[[L.lights[i][k] for k in range(3)] for i in range(L.n_lights)], types=[L.types[i] for i in range(L.n_lights)],
                n_ao=L.n_ao, ao_radius=L.ao_radius, shadows=bool(L.shadows), Ka=L.Ka, Kd=L.Kd)


def make_camera(d):
    c = Camera()
    for k in range(3):
        c.eye[k], c.dir[k], c.up[k] = d["eye"][k], d["dir"][k], d["up"][k]
    c.aov = d["aov"]
    return c


class Scene:
    """One partition's Visualization (mirrors galaxy_b200.gpu.Scene)."""

    def __init__(self):
        self.h = lib().gxo_scene_create()
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().gxo_scene_destroy(self.h)
            self.h = None

    def set_partition(self, gmin, gmax, lmin, lmax, neighbors):
        a = [_f32(x) for x in (gmin, gmax, lmin, lmax)]
        n = np.ascontiguousarray(neighbors, dtype=np.int32)
        lib().gxo_scene_set_partition(self.h, _f(a[0]), _f(a[1]), _f(a[2]), _f(a[3]), _i(n))

    def add_volume_vis(self, dataset_id, dims, origin, spacing, voxels, slices, isovalues, volume_render, colors, opacities, lo, hi):
        voxels = np.ascontiguousarray(voxels)
        assert voxels.dtype in (np.float32, np.uint8)
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        o, s = _f32(origin), _f32(spacing)
        sl = _f32(np.asarray(slices, dtype=np.float32).reshape(-1, 4)) if len(slices) else np.zeros((0, 4), np.float32)
        iso = _f32(isovalues) if len(isovalues) else np.zeros((0,), np.float32)
        col, op = _f32(colors), _f32(opacities)
        self._keep += [voxels]
        return lib().gxo_scene_add_volume_vis(self.h, dataset_id, _i(dims), _f(o), _f(s), 0 if voxels.dtype == np.float32 else 1,
                                              voxels.ctypes.data_as(C.c_void_p), len(sl), _f(sl), len(iso), _f(iso),
                                              int(volume_render), _f(col), _f(op), lo, hi)

    def add_triangles_vis(self, verts, normals, data, indices, colors, opacities, lo, hi):
        verts, normals, data = _f32(verts), _f32(normals), _f32(data)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        col, op = _f32(colors), _f32(opacities)
        self._keep += [verts, normals, data, indices]
        return lib().gxo_scene_add_triangles_vis(self.h, len(verts), _f(verts), _f(normals), _f(data), len(indices), _i(indices),
                                                 _f(col), _f(op), lo, hi)

    def add_particles_vis(self, centers, data, radius0, radius1, value0, value1, colors, opacities, lo, hi):
        centers, data = _f32(centers), _f32(data)
        col, op = _f32(colors), _f32(opacities)
        self._keep += [centers, data]
        return lib().gxo_scene_add_particles_vis(self.h, len(centers), _f(centers), _f(data), radius0, radius1, value0, value1,
                                                 _f(col), _f(op), lo, hi)

    def add_pathlines_vis(self, verts, data, connectivity, radius0, radius1, value0, value1, colors, opacities, lo, hi):
        verts, data = _f32(verts), _f32(data)
        conn = np.ascontiguousarray(connectivity, dtype=np.int32)
        col, op = _f32(colors), _f32(opacities)
        rc = lib().gxo_scene_add_pathlines_vis(self.h, len(verts), _f(verts), _f(data), len(conn), _i(conn), radius0, radius1, value0,
                                               value1, _f(col), _f(op), lo, hi)
        if rc < 0:
            raise ValueError("pathlines: bad connectivity")
        return rc

    def add_sampler_vis(self, dataset_id, dims, origin, spacing, voxels, kind, param):
        """kind: "GradientSampler" (param = tolerance) | "IsoSampler" (param = isovalue)"""
        voxels = np.ascontiguousarray(voxels)
        assert voxels.dtype in (np.float32, np.uint8)
        dims = np.ascontiguousarray(dims, dtype=np.int32)
        o, s = _f32(origin), _f32(spacing)
        self._keep += [voxels]
        rc = lib().gxo_scene_add_sampler_vis(self.h, dataset_id, _i(dims), _f(o), _f(s), 0 if voxels.dtype == np.float32 else 1,
                                             voxels.ctypes.data_as(C.c_void_p), {"GradientSampler": 0, "IsoSampler": 1}[kind], param)
        assert rc >= 0
        return rc

    def sample_raylist(self, rays, n):
        assert rays.dtype == np.float32 and rays.flags.c_contiguous and rays.shape[0] == 25
        lib().gxo_sample_raylist(self.h, _f(rays), n, rays.shape[1])

    def samples(self):
        p = C.POINTER(C.c_float)()
        n = lib().gxo_scene_samples(self.h, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n, 3)).copy() if n else np.zeros((0, 3), np.float32)

    def commit(self):
        return lib().gxo_scene_commit(self.h)

    # -- per-list entry points ---------------------------------------------------------------
    def trace_raylist(self, lighting, rays, n, epsilon=0.001, want_hits=False):
        """rays: float32 array (25, aligned_n) in RayList column order (ints bit-cast); traced in
        place.  Returns (secondary (25, aligned_out) or None, n_out, hit_ids or None)."""
        assert rays.dtype == np.float32 and rays.flags.c_contiguous and rays.shape[0] == 25
        hits = np.empty((n, 2), np.int32) if want_hits else None
        L = make_lighting(lighting)
        nout = lib().gxo_trace_raylist(self.h, C.byref(L), _f(rays), n, rays.shape[1], epsilon, _i(hits))
        out = None
        if nout > 0:
            al = max(16, (nout + 15) & ~15)
            out = np.zeros((25, al), np.float32)
            lib().gxo_fetch_secondary(self.h, _f(out), al)
        return out, int(nout), hits

    def classify(self, rays, n):
        lib().gxo_classify(self.h, _f(rays), n, rays.shape[1])

    def generate_rays(self, camera, w, h):
        al = max(16, (w * h + 15) & ~15)
        rays = np.zeros((25, al), np.float32)
        cam = make_camera(camera)
        n = lib().gxo_generate_rays(self.h, C.byref(cam), w, h, _f(rays), al)
        return rays, n

    def intersect(self, org, dir, tnear, tfar):
        org, dir, tnear, tfar = _f32(org), _f32(dir), _f32(tnear), _f32(tfar)
        n = len(org)
        ids = np.empty((n, 2), np.int32)
        tuv = np.empty((n, 3), np.float32)
        lib().gxo_intersect(self.h, n, _f(org), _f(dir), _f(tnear), _f(tfar), _i(ids), _f(tuv))
        return ids, tuv


def render(parts, camera, lighting, w, h, epsilon=0.001, max_rays_per_packet=1000000, nthreads=0):
    """Full frame over the list of partition Scenes.  Returns (fb float32 (h,w,4) y-up, stats dict)."""
    fb = np.zeros((h, w, 4), np.float32)
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, L, st = make_camera(camera), make_lighting(lighting), Stats()
    rc = lib().gxo_render(len(parts), arr, C.byref(cam), C.byref(L), w, h, epsilon, max_rays_per_packet, nthreads, _f(fb), C.byref(st))
    assert rc == 0
    return fb, st.as_dict()


class ProgressiveRendering:
    """the state a Rendering keeps on the interactive path: framebuffer, per-pixel frame stamps, current frame (Rendering.cpp:55-58,218-256)"""

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.fb = np.zeros((h, w, 4), np.float32)
        self.kbuffer = np.zeros((h, w), np.int32)
        self.frame = np.array([-1], np.int32)

    def render(self, parts, camera, lighting, frame, epsilon=0.001, nthreads=0):
        arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
        cam, L, st = make_camera(camera), make_lighting(lighting), Stats()
        rc = lib().gxo_render_progressive(len(parts), arr, C.byref(cam), C.byref(L), self.w, self.h, epsilon, nthreads, frame, _f(self.fb),
                                          _i(self.kbuffer), _i(self.frame), C.byref(st))
        assert rc == 0
        return self.fb.copy(), st.as_dict()


def sample(parts, camera, w, h, max_rays_per_packet=1000000, nthreads=0):
    """Sampler over a frame of camera rays.  Returns ([samples (n,3) per partition], stats dict)."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, st = make_camera(camera), Stats()
    rc = lib().gxo_sample(len(parts), arr, C.byref(cam), w, h, max_rays_per_packet, nthreads, C.byref(st))
    assert rc == 0
    return [p.samples() for p in parts], st.as_dict()


def resolve_lights(lighting, camera):
    L, cam, out = make_lighting(lighting), make_camera(camera), Lighting()
    lib().gxo_resolve_lights(C.byref(L), C.byref(cam), C.byref(out))
    return lighting_to_dict(out)


def resample_tf(cmap, omap):
    cmap, omap = _f32(np.asarray(cmap).reshape(-1, 4)), _f32(np.asarray(omap).reshape(-1, 2))
    col, op = np.empty((256, 3), np.float32), np.empty((256,), np.float32)
    lib().gxo_resample_tf(len(cmap), _f(cmap), len(omap), _f(omap), _f(col), _f(op))
    return col, op


def fb_to_rgba8(fb):
    h, w, _ = fb.shape
    fb = _f32(fb)
    out = np.empty((h, w, 4), np.uint8)
    lib().gxo_fb_to_rgba8(_f(fb), w, h, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out


def factor(n):
    f = np.zeros(3, np.int32)
    lib().gxo_factor(n, _i(f))
    return tuple(int(x) for x in f)


def partition(n, factors, grid):
    out = np.zeros((n, 15), np.int32)
    f, g = np.asarray(factors, np.int32), np.asarray(grid, np.int32)
    lib().gxo_partition(n, _i(f), _i(g), _i(out))
    return out
