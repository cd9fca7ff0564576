"""TEST INFRASTRUCTURE / CPU baseline arm.  ctypes binding of oracle/_ref/libgxy_embree_scene_ref.so: the reference's own vendored
Embree 3.6.1 (BVH8/Triangle4 SAH build + rtcIntersect8 packet traversal), compiled from /root/reference by `make -C oracle -f embree.mk
embree` (oracle/embree.mk, oracle/embree_scene_ref.cpp).  Only tests/ and bench.py's CPU legs may import this; the product never does."""
import ctypes as C
import os

import numpy as np

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libgxy_embree_scene_ref.so")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        L.gxy_embree_scene_create.restype = C.c_void_p
        L.gxy_embree_scene_create.argtypes = [fp, C.c_size_t, up, C.c_size_t, C.c_int]
        L.gxy_embree_scene_build_seconds.restype = C.c_double
        L.gxy_embree_scene_build_seconds.argtypes = [C.c_void_p]
        L.gxy_embree_scene_destroy.argtypes = [C.c_void_p]
        L.gxy_embree_intersect.restype = C.c_double
        L.gxy_embree_intersect.argtypes = [C.c_void_p, C.c_size_t, fp, fp, fp, fp, ip, ip, fp, fp, fp, fp, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class EmbreeScene:
    """A committed static triangle scene (the way Galaxy hands a Triangles dataset to Embree through OSPRay)."""

    def __init__(self, verts, indices, threads=0):
        v = np.ascontiguousarray(verts, np.float32)
        i = np.ascontiguousarray(indices, np.int32).view(np.uint32)
        self.h = lib().gxy_embree_scene_create(_p(v, C.c_float), len(v), _p(i, C.c_uint32), len(i), threads)
        if not self.h:
            raise RuntimeError("gxy_embree_scene_create failed")
        self.build_seconds = lib().gxy_embree_scene_build_seconds(self.h)
        self.n_tris = len(i)

    def __del__(self):
        if getattr(self, "h", None):
            lib().gxy_embree_scene_destroy(self.h)
            self.h = None

    def intersect(self, org, d, tnear, tfar, packet=8, threads=0, want=True):
        """-> (prim int32 (n,), tuv float32 (n,3), seconds).  want=False: timing only (no outputs stored)."""
        org, d = np.ascontiguousarray(org, np.float32), np.ascontiguousarray(d, np.float32)
        tn, tf = np.ascontiguousarray(tnear, np.float32), np.ascontiguousarray(tfar, np.float32)
        n = len(org)
        if want:
            prim, t, u, v = np.empty(n, np.int32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
            s = lib().gxy_embree_intersect(self.h, n, _p(org, C.c_float), _p(d, C.c_float), _p(tn, C.c_float), _p(tf, C.c_float), None,
                                           _p(prim, C.c_int32), _p(t, C.c_float), _p(u, C.c_float), _p(v, C.c_float), None, packet, threads)
            return prim, np.stack([t, u, v], 1), s
        s = lib().gxy_embree_intersect(self.h, n, _p(org, C.c_float), _p(d, C.c_float), _p(tn, C.c_float), _p(tf, C.c_float), None, None, None,
                                       None, None, None, packet, threads)
        return None, None, s


def attach_to_oracle_scene(oracle_scene, verts, indices, threads=0):
    """bench.py's CPU arm "reference": the oracle Scene `oracle_scene` (one TrianglesVis, not yet committed) takes its nearest hits
    from a committed EmbreeScene over the same triangles (rtcIntersect8 packets).  Returns the EmbreeScene (keep it alive)."""
    from oracle import oracle
    es = EmbreeScene(verts, indices, threads)
    fn = C.cast(lib().gxy_embree_intersect_cb, C.c_void_p)
    oracle.lib().gxo_scene_set_intersector(oracle_scene.h, fn, C.c_void_p(es.h))
    return es


def oracle_backend(threads=0):
    """A `backend` for scenes.build_partitions: oracle Scenes whose TrianglesVis is traced by the reference's Embree (one geometry
    operator, one partition: what bench.py's CPU arm renders).  Everything but the nearest-hit search stays the oracle's."""
    from oracle import oracle

    class Scene(oracle.Scene):
        def add_triangles_vis(self, verts, normals, data, indices, *rest):
            self._embree_mesh = (verts, indices)
            return super().add_triangles_vis(verts, normals, data, indices, *rest)

        def commit(self):
            mesh = getattr(self, "_embree_mesh", None)
            if mesh is not None:
                self.embree = attach_to_oracle_scene(self, mesh[0], mesh[1], threads)
            return super().commit()

    class Backend:
        pass

    Backend.Scene = Scene
    return Backend


def raylist_columns(rays, n):
    """(org (n,3), dir (n,3), t, tMax) of a 25-column RayList array as the oracle / the product hand it out"""
    org = np.ascontiguousarray(rays[0:3, :n].T)
    d = np.ascontiguousarray(rays[3:6, :n].T)
    return org, d, np.ascontiguousarray(rays[18, :n]), np.ascontiguousarray(rays[19, :n])
