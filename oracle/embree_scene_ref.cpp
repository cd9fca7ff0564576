// TEST INFRASTRUCTURE + CPU baseline arm (never on the product path).
// A plain C interface over the reference's own vendored Embree 3.6.1, compiled from its sources by oracle/embree.mk:
//   commit a static triangle scene exactly as Galaxy does through OSPRay -- rtcNewScene with no flags (Model.cpp:65-77 with the
//   default params: static, not compact, not robust => BVH8/Triangle4 binned-SAH on AVX2, embree/kernels/common/scene.cpp:123-137),
//   RTC_GEOMETRY_TYPE_TRIANGLE with FLOAT3 vertices (12-byte stride) and UINT3 indices
//   (src/ospray/DataDrivenTriangleMesh.cpp:129-139, ospray/geometry/TriangleMesh.cpp:129-136) --
//   and intersect a ray array with rtcIntersect8 packets (what ISPC's rtcIntersectV runs at programCount 8,
//   ospray/common/Model.ih:54-70) or rtcIntersect1, on T host threads.
// Used by tests/test_embree_traversal.py (pins nearest-hit ids of the CUDA traversal and of the oracle against the reference's
// own BVH build + traversal) and by bench.py's cpu_baseline / --impl reference legs (kind "reference").
#include <embree3/rtcore.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <xmmintrin.h>
#include <pmmintrin.h>

namespace {
struct Scene
{
    RTCDevice device = nullptr;
    RTCScene scene = nullptr;
    double build_seconds = 0;
};
double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

extern "C" {

// verts: nv x 3 floats; idx: nt x 3 uint32.  Buffers are copied (Embree reads 16 bytes per float3, so a shared buffer
// would need padding the caller does not promise).  threads <= 0: all cores.
void *gxy_embree_scene_create(const float *verts, size_t nv, const uint32_t *idx, size_t nt, int threads)
{
    Scene *s = new Scene;
    std::string cfg = threads > 0 ? "threads=" + std::to_string(threads) : std::string();
    s->device = rtcNewDevice(cfg.empty() ? nullptr : cfg.c_str());
    if (!s->device) { delete s; return nullptr; }
    s->scene = rtcNewScene(s->device);
    RTCGeometry g = rtcNewGeometry(s->device, RTC_GEOMETRY_TYPE_TRIANGLE);
    void *vb = rtcSetNewGeometryBuffer(g, RTC_BUFFER_TYPE_VERTEX, 0, RTC_FORMAT_FLOAT3, 3 * sizeof(float), nv);
    void *ib = rtcSetNewGeometryBuffer(g, RTC_BUFFER_TYPE_INDEX, 0, RTC_FORMAT_UINT3, 3 * sizeof(uint32_t), nt);
    if (!vb || !ib) { rtcReleaseGeometry(g); rtcReleaseScene(s->scene); rtcReleaseDevice(s->device); delete s; return nullptr; }
    std::memcpy(vb, verts, nv * 3 * sizeof(float));
    std::memcpy(ib, idx, nt * 3 * sizeof(uint32_t));
    double t0 = now();
    rtcCommitGeometry(g);
    rtcAttachGeometry(s->scene, g);
    rtcReleaseGeometry(g);
    rtcCommitScene(s->scene);
    s->build_seconds = now() - t0;
    return s;
}

double gxy_embree_scene_build_seconds(void *h) { return static_cast<Scene *>(h)->build_seconds; }

void gxy_embree_scene_destroy(void *h)
{
    Scene *s = static_cast<Scene *>(h);
    if (!s) return;
    rtcReleaseScene(s->scene);
    rtcReleaseDevice(s->device);
    delete s;
}

// Nearest hit of n rays in (tnear, tfar].  ng3 (may be null): Embree's unnormalised geometric normal, 3 floats per ray.  org/dir: n x 3 floats.  Outputs (any may be null): geomID/primID (-1 on a miss),
// t (tfar on a miss), u, v.  packet = 8: rtcIntersect8 over groups of 8 consecutive rays; 1: rtcIntersect1.
// Returns the wall-clock seconds of the intersect loop (threads joined).
double gxy_embree_intersect(void *h, size_t n, const float *org, const float *dir, const float *tnear, const float *tfar,
                            int32_t *geom_id, int32_t *prim_id, float *t, float *u, float *v, float *ng3, int packet, int threads)
{
    Scene *s = static_cast<Scene *>(h);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    const size_t chunk = 4096; // multiple of 8
    std::atomic<size_t> next(0);
    auto work = [&]() {
        const unsigned mxcsr = _mm_getcsr();                  // restored below: the calling thread may be the caller's own
        _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);           // as Embree asks of its callers
        _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
        for (;;) {
            size_t b = next.fetch_add(chunk);
            if (b >= n) break;
            size_t e = b + chunk < n ? b + chunk : n;
            if (packet == 8) {
                for (size_t i = b; i < e; i += 8) {
                    alignas(32) RTCRayHit8 rh;
                    alignas(32) int valid[8];
                    for (int k = 0; k < 8; k++) {
                        size_t j = i + k;
                        bool ok = j < e;
                        size_t q = ok ? j : i;
                        valid[k] = ok ? -1 : 0;
                        rh.ray.org_x[k] = org[3 * q]; rh.ray.org_y[k] = org[3 * q + 1]; rh.ray.org_z[k] = org[3 * q + 2];
                        rh.ray.dir_x[k] = dir[3 * q]; rh.ray.dir_y[k] = dir[3 * q + 1]; rh.ray.dir_z[k] = dir[3 * q + 2];
                        rh.ray.tnear[k] = tnear[q]; rh.ray.tfar[k] = tfar[q];
                        rh.ray.time[k] = 0.f; rh.ray.mask[k] = 0xffffffffu; rh.ray.id[k] = 0; rh.ray.flags[k] = 0;
                        rh.hit.geomID[k] = RTC_INVALID_GEOMETRY_ID; rh.hit.primID[k] = RTC_INVALID_GEOMETRY_ID;
                        rh.hit.instID[0][k] = RTC_INVALID_GEOMETRY_ID;
                        rh.hit.u[k] = 0.f; rh.hit.v[k] = 0.f;
                    }
                    RTCIntersectContext ctx;
                    rtcInitIntersectContext(&ctx);
                    rtcIntersect8(valid, s->scene, &ctx, &rh);
                    for (int k = 0; k < 8 && i + k < e; k++) {
                        size_t j = i + k;
                        if (geom_id) geom_id[j] = (int32_t)rh.hit.geomID[k];
                        if (prim_id) prim_id[j] = (int32_t)rh.hit.primID[k];
                        if (t) t[j] = rh.ray.tfar[k];
                        if (u) u[j] = rh.hit.u[k];
                        if (v) v[j] = rh.hit.v[k];
                        if (ng3) { ng3[3 * j] = rh.hit.Ng_x[k]; ng3[3 * j + 1] = rh.hit.Ng_y[k]; ng3[3 * j + 2] = rh.hit.Ng_z[k]; }
                    }
                }
            } else {
                for (size_t j = b; j < e; j++) {
                    RTCRayHit rh;
                    rh.ray.org_x = org[3 * j]; rh.ray.org_y = org[3 * j + 1]; rh.ray.org_z = org[3 * j + 2];
                    rh.ray.dir_x = dir[3 * j]; rh.ray.dir_y = dir[3 * j + 1]; rh.ray.dir_z = dir[3 * j + 2];
                    rh.ray.tnear = tnear[j]; rh.ray.tfar = tfar[j];
                    rh.ray.time = 0.f; rh.ray.mask = 0xffffffffu; rh.ray.id = 0; rh.ray.flags = 0;
                    rh.hit.geomID = RTC_INVALID_GEOMETRY_ID; rh.hit.primID = RTC_INVALID_GEOMETRY_ID;
                    rh.hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
                    rh.hit.u = rh.hit.v = 0.f;
                    RTCIntersectContext ctx;
                    rtcInitIntersectContext(&ctx);
                    rtcIntersect1(s->scene, &ctx, &rh);
                    if (geom_id) geom_id[j] = (int32_t)rh.hit.geomID;
                    if (prim_id) prim_id[j] = (int32_t)rh.hit.primID;
                    if (t) t[j] = rh.ray.tfar;
                    if (u) u[j] = rh.hit.u;
                    if (v) v[j] = rh.hit.v;
                    if (ng3) { ng3[3 * j] = rh.hit.Ng_x; ng3[3 * j + 1] = rh.hit.Ng_y; ng3[3 * j + 2] = rh.hit.Ng_z; }
                }
            }
        }
        _mm_setcsr(mxcsr);
    };
    double t0 = now();
    std::vector<std::thread> pool;
    for (int k = 1; k < threads; k++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    return now() - t0;
}

// the same as a callback of the oracle (gxo_intersect_fn, oracle/gxy_oracle.h): user = the scene; rtcIntersect8 packets on the
// calling thread (the oracle's trace loop is already one thread per chunk of the RayList)
void gxy_embree_intersect_cb(void *user, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar,
                             int *geom_id, int *prim_id, float *t, float *u, float *v, float *ng3)
{
    gxy_embree_intersect(user, (size_t)n, org3, dir3, tnear, tfar, geom_id, prim_id, t, u, v, ng3, 8, 1);
}

} // extern "C"
