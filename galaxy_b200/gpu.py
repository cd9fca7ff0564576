"""ctypes binding of the product library libgxy_b200.so (include/gxy_gpu.h).

Mirrors the reference's host-side objects for the hot path (Visualization / Lighting / Camera /
RayList / Renderer::render) one to one on top of the C ABI.  There is NO fallback: if the
library is not built or no CUDA device is usable every compute call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
# Frames in flight put up to 8 x 17 streams on the device.  With the default of 8 hardware queues several streams share a queue and a
# kernel that waits at a cross-rank flag barrier can hold up an unrelated stream behind it; 32 queues (the maximum) before the CUDA
# context exists.  The library does the same in gxy_context_create; both are no-ops when the caller has set the variable.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
MAX_LIGHTS = 16


class GxyError(RuntimeError):
    pass


class Lighting(C.Structure):
    _fields_ = [("n_lights", C.c_int), ("lights", (C.c_float * 3) * MAX_LIGHTS), ("types", C.c_int * MAX_LIGHTS),
                ("n_ao", C.c_int), ("ao_radius", C.c_float), ("shadows", C.c_int), ("Ka", C.c_float), ("Kd", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("dir", C.c_float * 3), ("up", C.c_float * 3), ("aov", C.c_float)]


class TransferFunction(C.Structure):
    _fields_ = [("colors", (C.c_float * 3) * 256), ("opacities", C.c_float * 256), ("range_lo", C.c_float), ("range_hi", C.c_float)]


class RayListView(C.Structure):
    _fields_ = [("base", C.POINTER(C.c_float)), ("n", C.c_int), ("aligned_n", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("primary_rays", "shadow_rays", "ao_rays", "forwarded_rays", "terminated_rays", "traced_rays",
                                            "waves", "kernel_launches")] + [("device_ms", C.c_float), ("trace_ms", C.c_float),
                                                                                ("nodes_visited", C.c_longlong), ("prims_tested", C.c_longlong), ("volume_samples", C.c_longlong), ("staged_samples", C.c_longlong),
                                                                                ("dequeued_rays", C.c_longlong), ("t_begin_ms", C.c_float), ("t_end_ms", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def library_path():
    # GXY_LIB selects a diagnostic build (e.g. libgxy_b200_counters.so, `make counters`)
    return os.environ.get("GXY_LIB") or os.path.join(_HERE, "libgxy_b200.so")


def build(force=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    so = library_path()
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "gxy_gpu.h"))
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        if not os.path.exists("/usr/local/cuda/bin/nvcc") and os.path.exists(so):
            return so
        subprocess.check_call(["make", "-C", src_dir, "-j4"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = library_path()
        if not os.path.exists(so):
            raise GxyError("libgxy_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(so)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.gxy_last_error.restype = C.c_char_p
        L.gxy_version.restype = C.c_char_p
        L.gxy_context_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.gxy_context_destroy.argtypes = [vp]
        L.gxy_context_synchronize.argtypes = [vp]
        L.gxy_volume_create.argtypes = [vp, ip, fp, fp, C.c_int, vp, C.POINTER(vp)]
        L.gxy_volume_destroy.argtypes = [vp]
        L.gxy_triangles_create.argtypes = [vp, C.c_int, fp, fp, fp, C.c_int, ip, C.POINTER(vp)]
        L.gxy_triangles_destroy.argtypes = [vp]
        L.gxy_particles_create.argtypes = [vp, C.c_int, fp, fp, C.POINTER(vp)]
        L.gxy_particles_destroy.argtypes = [vp]
        L.gxy_pathlines_create.argtypes = [vp, C.c_int, fp, fp, C.c_int, ip, C.POINTER(vp)]
        L.gxy_pathlines_destroy.argtypes = [vp]
        L.gxy_build_curves.argtypes = [C.c_int, fp, fp, C.c_int, ip, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        L.gxy_vis_add_pathlines.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(TransferFunction)]
        L.gxy_render_progressive.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float,
                                             C.c_int, C.POINTER(Stats)]
        L.gxy_render_progressive_submit.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float,
                                                    C.c_int, C.c_int]
        L.gxy_render_progressive_wait.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(Stats), ip]
        L.gxy_progressive_download_rgba32f.argtypes = [vp, fp]
        L.gxy_progressive_download_rgba8.argtypes = [vp, C.POINTER(C.c_ubyte)]
        L.gxy_progressive_reset.argtypes = [vp]
        L.gxy_vis_add_sampler.argtypes = [vp, vp, C.c_int, C.c_float]
        L.gxy_sample_raylist.argtypes = [vp, RayListView]
        L.gxy_sample.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.c_int, C.c_int, C.POINTER(Stats)]
        L.gxy_vis_sample_count.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.gxy_vis_download_samples.argtypes = [vp, fp]
        L.gxy_particles_from_samples.argtypes = [vp, C.POINTER(vp)]
        L.gxy_vis_create.argtypes = [vp, C.POINTER(vp)]
        L.gxy_vis_destroy.argtypes = [vp]
        L.gxy_vis_set_partition.argtypes = [vp, fp, fp, fp, fp, ip]
        L.gxy_vis_add_volume.argtypes = [vp, vp, C.c_int, fp, C.c_int, fp, C.c_int, C.POINTER(TransferFunction)]
        L.gxy_vis_add_triangles.argtypes = [vp, vp, C.POINTER(TransferFunction)]
        L.gxy_vis_add_particles.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(TransferFunction)]
        L.gxy_vis_commit.argtypes = [vp]
        L.gxy_vis_build_info.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), fp]
        L.gxy_vis_build_times.argtypes = [vp, fp, fp]
        L.gxy_resolve_lights.argtypes = [C.POINTER(Lighting), C.POINTER(Camera), C.POINTER(Lighting)]
        L.gxy_resample_transfer_function.argtypes = [C.c_int, fp, C.c_int, fp, C.POINTER(TransferFunction)]
        L.gxy_factor.argtypes = [C.c_int, ip]
        L.gxy_partition.argtypes = [C.c_int, ip, ip, ip]
        L.gxy_trace_raylist.argtypes = [vp, C.POINTER(Lighting), RayListView, C.c_float, C.POINTER(vp), ip]
        L.gxy_raylist_get_view.argtypes = [vp, C.POINTER(RayListView)]
        L.gxy_raylist_free.argtypes = [vp]
        L.gxy_classify.argtypes = [vp, RayListView]
        L.gxy_raylist_upload.argtypes = [vp, RayListView, C.POINTER(vp)]
        L.gxy_raylist_size.argtypes = [vp, ip]
        L.gxy_raylist_download.argtypes = [vp, RayListView]
        L.gxy_dev_raylist_free.argtypes = [vp]
        L.gxy_trace_raylist_dev.argtypes = [vp, C.POINTER(Lighting), vp, C.c_float, C.POINTER(vp)]
        L.gxy_classify_dev.argtypes = [vp, vp]
        L.gxy_generate_rays.argtypes = [vp, C.POINTER(Camera), C.c_int, C.c_int, RayListView, ip]
        L.gxy_intersect.argtypes = [vp, C.c_int, fp, fp, fp, fp, ip, fp]
        L.gxy_render.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float, C.POINTER(Stats)]
        L.gxy_render_submit.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(Camera), C.POINTER(Lighting), C.c_int, C.c_int, C.c_float, C.c_int]
        L.gxy_render_wait.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(Stats)]
        L.gxy_context_mark.argtypes = [vp]
        L.gxy_debug_tile_rect.argtypes = [C.POINTER(Camera), C.c_int, C.c_int, fp, fp, ip]
        L.gxy_frame_download_rgba32f.argtypes = [vp, fp]
        L.gxy_frame_download_rgba8.argtypes = [vp, C.POINTER(C.c_ubyte)]
        L.gxy_frame_download_rgba8_async.argtypes = [vp, C.POINTER(C.c_ubyte)]
        L.gxy_frame_download_wait.argtypes = [vp]
        L.gxy_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.gxy_host_free.argtypes = [vp]
        L.gxy_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
        L.gxy_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
        L.gxy_comm_destroy.argtypes = [vp]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise GxyError(lib().gxy_last_error().decode())


def device_count():
    return lib().gxy_device_count()


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def make_lighting(d):
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


def make_tf(colors, opacities, lo, hi):
    tf = TransferFunction()
    col, op = _f32(colors).reshape(256, 3), _f32(opacities).reshape(256)
    C.memmove(tf.colors, col.ctypes.data, 256 * 3 * 4)
    C.memmove(tf.opacities, op.ctypes.data, 256 * 4)
    tf.range_lo, tf.range_hi = lo, hi
    return tf


class Context:
    _default = {}

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().gxy_context_create(device, C.byref(self.h)))
        self.device = device

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = Context(device)
        return cls._default[device]

    def mark(self):
        """device idle + origin of the t_begin_ms / t_end_ms frame stamps (gxy_context_mark)"""
        check(lib().gxy_context_mark(self.h))

    def synchronize(self):
        check(lib().gxy_context_synchronize(self.h))

    def comm_init(self, rank, nranks, uid):
        buf = (C.c_ubyte * 128)(*uid)
        check(lib().gxy_comm_init(self.h, rank, nranks, buf))


def comm_unique_id():
    buf = (C.c_ubyte * 128)()
    check(lib().gxy_comm_unique_id(buf))
    return bytes(buf)


class Scene:
    """One partition's Visualization on one GPU (gxy_vis + the datasets it owns)."""

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or Context.default(device)
        self.h = C.c_void_p()
        check(lib().gxy_vis_create(self.ctx.h, C.byref(self.h)))
        self._volumes = {}
        self._owned = []

    def __del__(self):
        try:
            L = lib()
            if getattr(self, "h", None):
                L.gxy_vis_destroy(self.h)
                self.h = None
            for kind, h in getattr(self, "_owned", []):
                getattr(L, "gxy_%s_destroy" % kind)(h)
            self._owned = []
        except Exception:
            pass

    def set_partition(self, gmin, gmax, lmin, lmax, neighbors):
        a = [_f32(x) for x in (gmin, gmax, lmin, lmax)]
        n = np.ascontiguousarray(neighbors, dtype=np.int32)
        check(lib().gxy_vis_set_partition(self.h, _f(a[0]), _f(a[1]), _f(a[2]), _f(a[3]), _i(n)))

    def add_sampler_vis(self, dataset_id, dims, origin, spacing, voxels, kind, param):
        """kind: "GradientSampler" (param = tolerance) | "IsoSampler" (param = isovalue)"""
        if dataset_id not in self._volumes:
            voxels = np.ascontiguousarray(voxels)
            assert voxels.dtype in (np.float32, np.uint8)
            d = np.ascontiguousarray(dims, dtype=np.int32)
            o, s = _f32(origin), _f32(spacing)
            h = C.c_void_p()
            check(lib().gxy_volume_create(self.ctx.h, _i(d), _f(o), _f(s), 0 if voxels.dtype == np.float32 else 1,
                                          voxels.ctypes.data_as(C.c_void_p), C.byref(h)))
            self._volumes[dataset_id] = h
            self._owned.append(("volume", h))
        check(lib().gxy_vis_add_sampler(self.h, self._volumes[dataset_id], {"GradientSampler": 0, "IsoSampler": 1}[kind], param))

    def sample_raylist(self, rays, n):
        assert rays.dtype == np.float32 and rays.flags.c_contiguous and rays.shape[0] == 25
        check(lib().gxy_sample_raylist(self.h, RayListView(_f(rays), n, rays.shape[1])))

    def samples(self):
        n = C.c_longlong()
        check(lib().gxy_vis_sample_count(self.h, C.byref(n)))
        out = np.zeros((n.value, 3), np.float32)
        check(lib().gxy_vis_download_samples(self.h, _f(out)))
        return out

    def add_particles_from_samples(self, sampling_scene, radius0, radius1, value0, value1, colors, opacities, lo, hi):
        """a ParticlesVis on the samples another (sampling) Scene collected; they never leave the device"""
        h = C.c_void_p()
        check(lib().gxy_particles_from_samples(sampling_scene.h, C.byref(h)))
        self._owned.append(("particles", h))
        tf = make_tf(colors, opacities, lo, hi)
        check(lib().gxy_vis_add_particles(self.h, h, radius0, radius1, value0, value1, C.byref(tf)))

    def add_volume_vis(self, dataset_id, dims, origin, spacing, voxels, slices, isovalues, volume_render, colors, opacities, lo, hi):
        if dataset_id not in self._volumes:
            voxels = np.ascontiguousarray(voxels)
            assert voxels.dtype in (np.float32, np.uint8)
            d = np.ascontiguousarray(dims, dtype=np.int32)
            o, s = _f32(origin), _f32(spacing)
            h = C.c_void_p()
            check(lib().gxy_volume_create(self.ctx.h, _i(d), _f(o), _f(s), 0 if voxels.dtype == np.float32 else 1,
                                          voxels.ctypes.data_as(C.c_void_p), C.byref(h)))
            self._volumes[dataset_id] = h
            self._owned.append(("volume", h))
        sl = _f32(np.asarray(slices, dtype=np.float32).reshape(-1, 4)) if len(slices) else np.zeros((0, 4), np.float32)
        iso = _f32(isovalues) if len(isovalues) else np.zeros((0,), np.float32)
        tf = make_tf(colors, opacities, lo, hi)
        check(lib().gxy_vis_add_volume(self.h, self._volumes[dataset_id], len(sl), _f(sl), len(iso), _f(iso), int(volume_render), C.byref(tf)))

    def add_triangles_vis(self, verts, normals, data, indices, colors, opacities, lo, hi):
        verts, normals, data = _f32(verts), _f32(normals), _f32(data)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        h = C.c_void_p()
        check(lib().gxy_triangles_create(self.ctx.h, len(verts), _f(verts), _f(normals), _f(data), len(indices), _i(indices), C.byref(h)))
        self._owned.append(("triangles", h))
        tf = make_tf(colors, opacities, lo, hi)
        check(lib().gxy_vis_add_triangles(self.h, h, C.byref(tf)))

    def add_particles_vis(self, centers, data, radius0, radius1, value0, value1, colors, opacities, lo, hi):
        centers, data = _f32(centers), _f32(data)
        h = C.c_void_p()
        check(lib().gxy_particles_create(self.ctx.h, len(centers), _f(centers), _f(data), C.byref(h)))
        self._owned.append(("particles", h))
        tf = make_tf(colors, opacities, lo, hi)
        check(lib().gxy_vis_add_particles(self.h, h, radius0, radius1, value0, value1, C.byref(tf)))

    def add_pathlines_vis(self, verts, data, connectivity, radius0, radius1, value0, value1, colors, opacities, lo, hi):
        verts, data = _f32(verts), _f32(data)
        conn = np.ascontiguousarray(connectivity, dtype=np.int32)
        h = C.c_void_p()
        check(lib().gxy_pathlines_create(self.ctx.h, len(verts), _f(verts), _f(data), len(conn), _i(conn), C.byref(h)))
        self._owned.append(("pathlines", h))
        tf = make_tf(colors, opacities, lo, hi)
        check(lib().gxy_vis_add_pathlines(self.h, h, radius0, radius1, value0, value1, C.byref(tf)))

    def commit(self):
        check(lib().gxy_vis_commit(self.h))

    def build_info(self):
        a, b, ms = C.c_longlong(), C.c_longlong(), C.c_float()
        check(lib().gxy_vis_build_info(self.h, C.byref(a), C.byref(b), C.byref(ms)))
        al = C.c_float()
        check(lib().gxy_vis_build_times(self.h, None, C.byref(al)))
        return dict(n_prims=a.value, n_nodes=b.value, build_ms=ms.value, alloc_host_ms=al.value)

    # -- per-list entry points (TraceRays::Trace etc.) ------------------------------------------
    def trace_raylist(self, lighting, rays, n, epsilon=0.001, want_hits=False):
        assert rays.dtype == np.float32 and rays.flags.c_contiguous and rays.shape[0] == 25
        hits = np.empty((n, 2), np.int32) if want_hits else None
        L = make_lighting(lighting)
        view = RayListView(_f(rays), n, rays.shape[1])
        out_h = C.c_void_p()
        check(lib().gxy_trace_raylist(self.h, C.byref(L), view, epsilon, C.byref(out_h), _i(hits)))
        out, nout = None, 0
        if out_h:
            ov = RayListView()
            check(lib().gxy_raylist_get_view(out_h, C.byref(ov)))
            nout = ov.n
            out = np.ctypeslib.as_array(ov.base, shape=(25, ov.aligned_n)).copy()
            lib().gxy_raylist_free(out_h)
        return out, nout, hits

    def classify(self, rays, n):
        check(lib().gxy_classify(self.h, RayListView(_f(rays), n, rays.shape[1])))

    def upload_raylist(self, rays, n):
        """A device-resident copy of the first n rays of a (25, aligned_n) RayList array (gxy_raylist_upload)."""
        h = C.c_void_p()
        check(lib().gxy_raylist_upload(self.h, RayListView(_f(rays), n, rays.shape[1]), C.byref(h)))
        return DevRayList(self, h)

    def generate_rays(self, camera, w, h):
        al = max(16, (w * h + 15) & ~15)
        rays = np.zeros((25, al), np.float32)
        cam = make_camera(camera)
        n = C.c_int()
        check(lib().gxy_generate_rays(self.h, C.byref(cam), w, h, RayListView(_f(rays), 0, al), C.byref(n)))
        return rays, n.value

    def intersect(self, org, dir, tnear, tfar):
        org, dir, tnear, tfar = _f32(org), _f32(dir), _f32(tnear), _f32(tfar)
        n = len(org)
        ids = np.empty((n, 2), np.int32)
        tuv = np.empty((n, 3), np.float32)
        check(lib().gxy_intersect(self.h, n, _f(org), _f(dir), _f(tnear), _f(tfar), _i(ids), _f(tuv)))
        return ids, tuv

    def download_rgba32f(self, w, h):
        fb = np.empty((h, w, 4), np.float32)
        check(lib().gxy_frame_download_rgba32f(self.h, _f(fb)))
        return fb

    def download_rgba8(self, w, h, out=None):
        """RGBA8 image of the last frame; `out` may be a reusable (pinned) buffer from pinned_array()."""
        if out is None:
            out = np.empty((h, w, 4), np.uint8)
        check(lib().gxy_frame_download_rgba8(self.h, out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out

    def download_rgba8_async(self, out):
        """Start the D2H of the last frame into `out` (a pinned_array); the next render overlaps it.  download_wait() before reading."""
        check(lib().gxy_frame_download_rgba8_async(self.h, out.ctypes.data_as(C.POINTER(C.c_ubyte))))

    def download_wait(self):
        check(lib().gxy_frame_download_wait(self.h))


class DevRayList:
    """A RayList that stays on the device between Trace / Classify calls (gxy_dev_raylist)."""

    def __init__(self, scene, h):
        self.scene, self.h = scene, h

    def __del__(self):
        try:
            if self.h:
                lib().gxy_dev_raylist_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n(self):
        n = C.c_int()
        check(lib().gxy_raylist_size(self.h, C.byref(n)))
        return n.value

    def trace(self, lighting, epsilon=0.001):
        """TraceRays::Trace in place; returns the spawned SECONDARY list as another DevRayList, or None"""
        L = make_lighting(lighting)
        out = C.c_void_p()
        check(lib().gxy_trace_raylist_dev(self.scene.h, C.byref(L), self.h, epsilon, C.byref(out)))
        return DevRayList(self.scene, out) if out else None

    def classify(self):
        check(lib().gxy_classify_dev(self.scene.h, self.h))

    def download(self):
        """-> ((25, aligned_n) float32 array, n)"""
        n = self.n
        al = max(16, (n + 15) & ~15)
        rays = np.zeros((25, al), np.float32)
        check(lib().gxy_raylist_download(self.h, RayListView(_f(rays), n, al)))
        return rays, n


class _Pinned:
    def __init__(self, nbytes):
        self.p = C.c_void_p()
        check(lib().gxy_host_alloc(nbytes, C.byref(self.p)))

    def __del__(self):
        try:
            lib().gxy_host_free(self.p)
        except Exception:
            pass


def pinned_array(shape, dtype):
    """numpy array over page-locked host memory (gxy_host_alloc); freed with the array."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    owner = _Pinned(n)
    buf = (C.c_ubyte * n).from_address(owner.p.value)
    buf._gxy_owner = owner  # the ctypes buffer is the base object of the array: the allocation lives exactly as long as any view of it
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def render_progressive(parts, camera, lighting, w, h, frame, epsilon=0.001):
    """The interactive frame path (gxy_render_progressive): the displayed image lives on parts[0] across calls.
    Returns (displayed image float32 (h,w,4) y-up, stats dict)."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, L, st = make_camera(camera), make_lighting(lighting), Stats()
    check(lib().gxy_render_progressive(len(parts), arr, C.byref(cam), C.byref(L), w, h, epsilon, frame, C.byref(st)))
    fb = np.empty((h, w, 4), np.float32)
    check(lib().gxy_progressive_download_rgba32f(parts[0].h, _f(fb)))
    return fb, st.as_dict()


def render_progressive_submit(parts, camera, lighting, w, h, frame, slot, epsilon=0.001):
    """frame `frame` of the interactive path onto frame slot `slot` (gxy_render_progressive_submit)"""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, L = make_camera(camera), make_lighting(lighting)
    check(lib().gxy_render_progressive_submit(len(parts), arr, C.byref(cam), C.byref(L), w, h, epsilon, frame, slot))


def render_progressive_wait(parts, w, h, slot):
    """-> (displayed image float32 (h,w,4) y-up, stats, merged: False if the frame was superseded while in flight and dropped)"""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    st, merged = Stats(), C.c_int()
    check(lib().gxy_render_progressive_wait(len(parts), arr, slot, C.byref(st), C.byref(merged)))
    fb = np.empty((h, w, 4), np.float32)
    check(lib().gxy_progressive_download_rgba32f(parts[0].h, _f(fb)))
    return fb, st.as_dict(), bool(merged.value)


def progressive_reset(part):
    check(lib().gxy_progressive_reset(part.h))


def sample(parts, camera, w, h):
    """Sampler over a frame of camera rays (gxy_sample).  Returns ([samples (n,3) per partition], stats dict)."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, st = make_camera(camera), Stats()
    check(lib().gxy_sample(len(parts), arr, C.byref(cam), w, h, C.byref(st)))
    return [p.samples() for p in parts], st.as_dict()


def render_submit(parts, camera, lighting, w, h, epsilon=0.001, slot=0):
    """Enqueue one frame of a RenderingSet on frame slot `slot` (gxy_render_submit); returns without waiting for the device."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, L = make_camera(camera), make_lighting(lighting)
    check(lib().gxy_render_submit(len(parts), arr, C.byref(cam), C.byref(L), w, h, epsilon, slot))


def render_wait(parts, slot=0):
    """Wait for the frame on `slot`; it becomes the last frame of parts[0] for the download calls.  Returns its stats."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    st = Stats()
    check(lib().gxy_render_wait(len(parts), arr, slot, C.byref(st)))
    return st.as_dict()


def debug_tile_rect(camera, w, h, lo, hi):
    """(valid, (x0, y0, nx, ny)) in 8x4-pixel tiles: where a rank with partition box [lo, hi] generates primaries (host arithmetic)"""
    cam = make_camera(camera)
    lo, hi, rect = _f32(lo), _f32(hi), np.zeros(4, np.int32)
    rc = lib().gxy_debug_tile_rect(C.byref(cam), w, h, _f(lo), _f(hi), _i(rect))
    return rc, tuple(int(x) for x in rect)


def max_slots():
    return lib().gxy_render_max_slots()


def render_device(parts, camera, lighting, w, h, epsilon=0.001):
    """Renderer::render for one frame; the framebuffer stays on the device of parts[0]."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    cam, L, st = make_camera(camera), make_lighting(lighting), Stats()
    check(lib().gxy_render(len(parts), arr, C.byref(cam), C.byref(L), w, h, epsilon, C.byref(st)))
    return st.as_dict()


def render(parts, camera, lighting, w, h, epsilon=0.001, **_ignored):
    """Same interface as oracle.render: (fb float32 (h,w,4) y-up, stats)."""
    st = render_device(parts, camera, lighting, w, h, epsilon)
    return parts[0].download_rgba32f(w, h), st


def resolve_lights(lighting, camera):
    L, cam, out = make_lighting(lighting), make_camera(camera), Lighting()
    check(lib().gxy_resolve_lights(C.byref(L), C.byref(cam), C.byref(out)))
    return lighting_to_dict(out)


def resample_tf(cmap, omap):
    cmap, omap = _f32(np.asarray(cmap).reshape(-1, 4)), _f32(np.asarray(omap).reshape(-1, 2))
    tf = TransferFunction()
    check(lib().gxy_resample_transfer_function(len(cmap), _f(cmap), len(omap), _f(omap), C.byref(tf)))
    return np.ctypeslib.as_array(tf.colors).reshape(256, 3).copy(), np.ctypeslib.as_array(tf.opacities).copy()


def factor(n):
    f = np.zeros(3, np.int32)
    lib().gxy_factor(n, _i(f))
    return tuple(int(x) for x in f)


def partition(n, factors, grid):
    out = np.zeros((n, 15), np.int32)
    f, g = np.asarray(factors, np.int32), np.asarray(grid, np.int32)
    lib().gxy_partition(n, _i(f), _i(g), _i(out))
    return out
