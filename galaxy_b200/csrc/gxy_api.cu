// gxy_api.cu -- the C ABI of include/gxy_gpu.h: handles, commit, the per-RayList entry points
// (TraceRays::Trace, Classify, SpawnRays) and the frame-level wave loop that replaces
// Renderer::local_render / processRays_task / RayQManager / SendRaysMsg / SendPixelsMsg
// (src/renderer/Renderer.cpp:179-269,504-656,732-836; Rendering.cpp:125-153) on the device.
#include "gxy_internal.h"

#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace gxy;

#include <chrono>
// GXY_PROFILE=1: named CUDA events after the launches of a frame (device timeline, printed per rank at the end of the frame)
static std::vector<std::pair<const char *, cudaEvent_t>> g_timeline;
static bool timeline_on() { const char *e = getenv("GXY_PROFILE"); return e && atoi(e); }
namespace gxy {
void gxy_timeline_mark(const char *name, cudaStream_t st) {
  if (!timeline_on()) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_timeline.push_back(std::make_pair(name, e));
}
}  // namespace gxy
static void timeline_print(int rank) {
  if (g_timeline.empty()) return;
  std::string line = "[timeline rank " + std::to_string(rank) + "]";
  float prev = 0.f;
  for (size_t i = 1; i < g_timeline.size(); i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_timeline[0].second, g_timeline[i].second);
    char buf[96];
    snprintf(buf, sizeof buf, " %s %.3f(+%.3f)", g_timeline[i].first, ms, ms - prev);
    line += buf;
    prev = ms;
  }
  fprintf(stderr, "%s\n", line.c_str());
  for (auto &e : g_timeline) cudaEventDestroy(e.second);
  g_timeline.clear();
}

// GXY_TILED=0 restores the reference's row-major primary-ray order on the frame path (tuning / A-B runs)
static bool tiled_order() { const char *e = getenv("GXY_TILED"); return !(e && atoi(e) == 0); }

static bool march_tma_on() { const char *e = getenv("GXY_MARCH_TMA"); return e && atoi(e) != 0; }

struct PhaseTimer {  // GXY_PROFILE=1: host wall-clock per phase of gxy_render (each mark synchronises nothing)
  bool on;
  std::chrono::steady_clock::time_point t0;
  std::vector<std::pair<std::string, double>> marks;
  PhaseTimer() : on(getenv("GXY_PROFILE") && atoi(getenv("GXY_PROFILE"))), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *name) {
    if (!on) return;
    auto t = std::chrono::steady_clock::now();
    marks.push_back(std::make_pair(std::string(name), std::chrono::duration<double, std::milli>(t - t0).count()));
    t0 = t;
  }
  void report(int rank) {
    if (!on) return;
    std::map<std::string, double> agg;
    std::vector<std::string> order;
    for (auto &m : marks) { if (!agg.count(m.first)) order.push_back(m.first); agg[m.first] += m.second; }
    std::string line = "[gxy_render rank " + std::to_string(rank) + "]";
    char buf[64];
    for (auto &k : order) { snprintf(buf, sizeof buf, " %s=%.3f", k.c_str(), agg[k]); line += buf; }
    fprintf(stderr, "%s\n", line.c_str());
  }
};

// ------------------------------------------------------------------------------------------------
static thread_local char g_error[1024] = "";
void gxy_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}
#define GXY_CHECK(cond, ...)          \
  do {                                \
    if (!(cond)) {                    \
      gxy_set_error(__VA_ARGS__);     \
      return 1;                       \
    }                                 \
  } while (0)

// ---- NCCL through dlopen (no link-time dependency; the single-GPU path never touches it) ---------
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  GXY_CHECK(g_nccl.lib, "cannot dlopen libnccl.so.2: %s", dlerror());
#define LOADSYM(field, name)                                        \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);              \
  GXY_CHECK(g_nccl.field, "libnccl: missing symbol %s", name);
  LOADSYM(GetUniqueId, "ncclGetUniqueId")
  LOADSYM(CommInitRank, "ncclCommInitRank")
  LOADSYM(CommDestroy, "ncclCommDestroy")
  LOADSYM(Send, "ncclSend")
  LOADSYM(Recv, "ncclRecv")
  LOADSYM(AllGather, "ncclAllGather")
  LOADSYM(Reduce, "ncclReduce")
  LOADSYM(GroupStart, "ncclGroupStart")
  LOADSYM(GroupEnd, "ncclGroupEnd")
  LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
  return 0;
}
#define GXY_NCCL(call)                                                                              \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess) {                                                                        \
      gxy_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));  \
      return 1;                                                                                     \
    }                                                                                               \
  } while (0)

// ------------------------------------------------------------------------------------------------
// one rank's peer arena and the arenas of all other ranks mapped through CUDA IPC (PeerTable, gxy_internal.h)
struct PeerArena {
  char *base = nullptr;
  size_t bytes = 0;
  PeerTable T;
  void *opened[GXY_MAX_RANKS];
  unsigned epoch = 0;
  bool disabled = false;  // IPC mapping failed on some rank (or GXY_PEER=0): the NCCL exchange is used instead
  PeerArena() { memset(&T, 0, sizeof T); memset(opened, 0, sizeof opened); }
};

struct gxy_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // D2H of finished images while the next frame renders
  cudaStream_t lanes[16] = {};  // extra lanes of the band pipeline of the fused frame path ([0] unused: the main stream)
  cudaEvent_t ev_fork = nullptr, ev_join[16] = {};
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  PeerArena arena;
  cudaEvent_t ev_ref = nullptr;  // origin of gxy_stats::t_begin_ms / t_end_ms (context creation, or the last gxy_context_mark)
};

struct gxy_volume {
  gxy_context *ctx;
  DevVolume dv;
  void *d_vox;
  int last_tf;
};
struct gxy_triangles {
  gxy_context *ctx;
  int nv, nt;
  float *d_verts, *d_normals, *d_data;
  int *d_idx;
};
struct gxy_particles {
  gxy_context *ctx;
  int n;
  float *d_centers, *d_data;
};
struct gxy_pathlines {
  gxy_context *ctx;
  std::vector<float> verts, data;   // host copies: the curves depend on the Vis (radius mapping), built at commit
  std::vector<int> conn;
  bool has_data;
};
struct gxy_raylist {
  float *base;
  int n, aligned_n;
};

// growable device ray list
struct RayBuf {
  float *base = nullptr;
  size_t cap = 0;
  Rays v;
  int reserve(size_t n, bool keep, cudaStream_t st) {
    if (n <= cap) return 0;
    size_t ncap = std::max(n, cap + cap / 2);
    ncap = (ncap + 63) & ~(size_t)63;
    float *nb = nullptr;
    GXY_CUDA(cudaMalloc(&nb, sizeof(float) * GXY_RAYLIST_COLUMNS * ncap));
    if (keep && base && cap) {
      GXY_CUDA(cudaMemcpy2DAsync(nb, ncap * 4, base, cap * 4, cap * 4, GXY_RAYLIST_COLUMNS, cudaMemcpyDeviceToDevice, st));
      GXY_CUDA(cudaStreamSynchronize(st));
    }
    if (base) cudaFree(base);
    base = nb;
    cap = ncap;
    v = rays_view(base, cap);
    return 0;
  }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
  }
};

template <typename T>
struct Scratch {
  T *p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return 0;
    const size_t ncap = (n + n / 4 + 255) & ~(size_t)255;
    T *np = nullptr;
    if (cudaMalloc(&np, sizeof(T) * ncap) != cudaSuccess) {  // the old buffer goes too: nothing may write through a stale size
      gxy_set_error("cudaMalloc of %zu bytes failed: %s", sizeof(T) * ncap, cudaGetErrorString(cudaGetLastError()));
      release();
      return 1;
    }
    if (p) cudaFree(p);
    p = np;
    cap = ncap;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};


// One frame in flight on the fused frame paths (geometry-only Visualizations): its own streams, queues, hit records,
// partial framebuffer, error flag and -- one process per GPU -- its own peer arena, so that several frames of a
// RenderingSet overlap on the device the way the reference keeps all Renderings of a set in flight at once
// (src/apps/gxywriter.cpp:196-264 starts every Rendering before the first wait; RayQManager.cpp:69-82 interleaves their lists).
#define GXY_MAX_FLIGHTS 16
struct Flight {
  bool pending = false, peer = false, sync_done = false;
  cudaStream_t st = nullptr, lanes[16] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[16] = {}, ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace_ev;  // pool, reused by every frame of this flight
  size_t n_trace_ev = 0;
  RayBuf hits, next, cur;
  Scratch<unsigned long long> fq;
  Scratch<unsigned> rawhits;
  Scratch<float> fb;
  Scratch<unsigned char> proxies;
  Scratch<int> err;  // this flight's device error flag (SceneParams::error_flag of its launches)
  PeerArena arena;
  struct Tail { FusedQueues q[16]; unsigned long long trav[2]; int error, pad; } *h_tail = nullptr;  // page-locked: end-of-frame counters
  float *fb_result = nullptr;
  int w = 0, h = 0, n_bands = 1, n_sec_per_hit = 0, peer_k = 0;
  float epsilon = 0.f;
  gxy_lighting lights;
  DevLights L;
  gxy_stats S;
  bool volume = false;             // peer frame of a Visualization with volumes: list kernels with device-side lengths (flight_enqueue_volume)
  RayBuf vl[2];                    // ... its two device lists
  Scratch<int> v_hit_index, v_block_sums;
  Scratch<unsigned long long> vq;  // VolQueues
  int v_cap = 0, v_k = 0;
  bool graph_frame = false;        // this frame was submitted as one CUDA graph launch (GXY_GRAPH=1, flight_submit_peer)
  cudaGraphExec_t gexec = nullptr;
  int prog_frame = 0;     // gxy_render_progressive_submit: the frame number and camera of the frame on this slot
  gxy_camera prog_cam;
};


// Scratch of ONE call of a per-RayList entry point (gxy_trace_raylist, gxy_classify, gxy_sample_raylist, gxy_generate_rays,
// gxy_intersect and their device-list forms).  The reference calls TraceRays::Trace from GXY_NTHREADS pool threads at once, each with
// its own TraceRays object, all sharing one Visualization (src/framework/Application.cpp:75, src/renderer/Renderer.cpp:504-556): a
// call takes a free lane of the Visualization (own stream, own lists, own error flag), concurrent calls run side by side on the
// device, and more callers than lanes wait for one.  The committed scene (SceneParams, BVH, volumes) is read-only during calls.
#define GXY_LIST_LANES 8
struct ListLane {
  bool busy = false;
  cudaStream_t st = nullptr;
  RayBuf cur, next;
  Scratch<int> hit_index, block_sums, small, io_i, err;
  Scratch<float> io_f;
};
struct gxy_dev_raylist {  // a RayList that lives on the device between calls (SURVEY 8b "Ownership")
  gxy_context *ctx;
  RayBuf buf;
  int n;
};

struct VolOp {
  gxy_volume *vol;
  std::vector<float> slices, iso;
  int volume_render;
  gxy_transfer_function tf;
};
struct GeomOp {
  int kind;
  gxy_triangles *tri;
  gxy_particles *par;
  gxy_pathlines *pl;
  float radius0, radius1, value0, value1;
  gxy_transfer_function tf;
};

struct SamplerOp {
  gxy_volume *vol;
  int kind;
  float param;
};

struct gxy_vis {
  gxy_context *ctx;
  float gmin[3], gmax[3], lmin[3], lmax[3];
  int neighbors[6];
  std::vector<VolOp> vols;
  std::vector<GeomOp> geoms;
  // interactive frame path (gxy_render_progressive): displayed image, per-pixel frame stamps, pixels touched by the last frame
  float *prog_image = nullptr;
  int *prog_kbuffer = nullptr;
  unsigned char *prog_touched = nullptr;
  int prog_w = 0, prog_h = 0, prog_frame = -1;
  std::vector<SamplerOp> samplers;  // a sampling Visualization (src/sampler) holds only these
  SamplerParams SP;
  float *d_samples = nullptr;       // xyz of the samples collected by gxy_sample
  unsigned long long samples_cap = 0, n_samples = 0;
  unsigned long long *d_sample_count = nullptr;
  bool committed = false;
  SceneParams P;
  DevTF *d_tfs = nullptr;
  DevGeom *d_geoms = nullptr;
  int *d_error = nullptr;
  BvhResult bvh;
  std::vector<float *> d_curves;   // control points of the PathLines operators (device), one buffer per operator
  bool has_dvr = false;
  // work buffers
  RayBuf cur, next, send, recv, hits;
  Scratch<unsigned long long> fq;  // FusedQueues of the fused frame path
  Scratch<unsigned> rawhits;
  Scratch<unsigned char> proxies;  // peer path: the PartProxy of every rank (launch_proxy_gather)
  Scratch<int> hit_index, block_sums, small;  // small: nhit, counts, offsets, cursor ...
  Scratch<unsigned long long> counters;       // [0] terminated [1] samples
  Scratch<float> fb, fb_tmp;
  float *fb_result = nullptr;  // where the last frame's image is (fb.p, or a buffer inside the peer arena)
  Scratch<unsigned char> rgba8;
  struct FrameTailT { unsigned long long counters[4], trav[2]; int error, pad; } *h_tail = nullptr;  // page-locked: end-of-frame counters
  Scratch<unsigned char> rgba8_async[2];  // staging images of gxy_frame_download_rgba8_async
  cudaEvent_t async_ready[2] = {nullptr, nullptr}, async_done[2] = {nullptr, nullptr};
  int async_slot = 0;
  Scratch<float> io_f;
  Scratch<int> io_i;
  int fb_w = 0, fb_h = 0;
  Flight *flights[GXY_MAX_FLIGHTS] = {};
  std::mutex lanes_mu;
  std::condition_variable lanes_cv;
  ListLane *list_lanes[GXY_LIST_LANES] = {};  // frames in flight (gxy_render_submit / gxy_render_wait); [0] also serves gxy_render
};

static int use_device(gxy_context *c) {
  GXY_CUDA(cudaSetDevice(c->device));
  return 0;
}

template <typename T>
static int upload(T **dst, const T *src, size_t n) {
  *dst = nullptr;
  if (!src || !n) return 0;
  GXY_CUDA(cudaMalloc(dst, sizeof(T) * n));
  GXY_CUDA(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char *gxy_last_error(void) { return g_error; }
const char *gxy_version(void) { return "galaxy_b200 0.1 (sm_100a)"; }

int gxy_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gxy_context_create(int device, gxy_context **out) {
  // frames in flight use many streams; effective only if the CUDA context does not exist yet (see gxy_render_submit)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int n = gxy_device_count();
  GXY_CHECK(n > 0, "no CUDA device available: galaxy_b200 has no CPU fallback");
  GXY_CHECK(device >= 0 && device < n, "invalid device %d (have %d)", device, n);
  GXY_CUDA(cudaSetDevice(device));
  gxy_context *c = new gxy_context();
  c->device = device;
  GXY_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  GXY_CUDA(cudaEventCreate(&c->ev_ref));
  GXY_CUDA(cudaEventRecord(c->ev_ref, c->stream));
  *out = c;
  return 0;
}
void gxy_context_destroy(gxy_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (int k = 0; k < 16; k++) {
    if (c->lanes[k]) cudaStreamDestroy(c->lanes[k]);
    if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_ref) cudaEventDestroy(c->ev_ref);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
int gxy_context_synchronize(gxy_context *c) {
  if (use_device(c)) return 1;
  GXY_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int gxy_context_mark(gxy_context *c) {
  GXY_CHECK(c, "NULL context");
  if (use_device(c)) return 1;
  GXY_CUDA(cudaDeviceSynchronize());
  GXY_CUDA(cudaEventRecord(c->ev_ref, c->stream));
  GXY_CUDA(cudaEventSynchronize(c->ev_ref));
  return 0;
}

// ---- datasets --------------------------------------------------------------------------------
int gxy_volume_create(gxy_context *c, const int dims[3], const float origin[3], const float spacing[3], int type, const void *voxels,
                      gxy_volume **out) {
  if (use_device(c)) return 1;
  GXY_CHECK(type == 0 || type == 1, "volume type must be 0 (float) or 1 (uchar)");
  GXY_CHECK(dims[0] >= 2 && dims[1] >= 2 && dims[2] >= 2, "volume dims must be >= 2");
  gxy_volume *v = new gxy_volume();
  v->ctx = c;
  v->last_tf = -1;
  DevVolume &d = v->dv;
  for (int k = 0; k < 3; k++) d.dims[k] = dims[k];
  d.type = type;
  d.origin = make_float3(origin[0], origin[1], origin[2]);
  d.spacing = make_float3(spacing[0], spacing[1], spacing[2]);
  d.rcp = make_float3(1.0f / spacing[0], 1.0f / spacing[1], 1.0f / spacing[2]);  // rcp(gridSpacing), StructuredVolume.ispc:208-211
  d.upper = make_float3(nextafterf((float)(dims[0] - 1), 0.f), nextafterf((float)(dims[1] - 1), 0.f),
                        nextafterf((float)(dims[2] - 1), 0.f));                    // StructuredVolume.ispc:223
  d.samplingStep = fminf(fminf(spacing[0], spacing[1]), spacing[2]);              // StructuredVolume.ispc:250
  d.samplingRate = 1.0f;                                                          // OsprayVolume.cpp:45
  d.nx = (unsigned long long)dims[0];
  d.nxy = (unsigned long long)dims[0] * dims[1];
  d.tf = 0;
  d.pad = 0;
  const size_t bytes = (size_t)dims[0] * dims[1] * dims[2] * (type == 0 ? 4 : 1);
  if (cudaMalloc(&v->d_vox, bytes) != cudaSuccess) { delete v; gxy_set_error("cudaMalloc(%zu) failed for volume", bytes); return 1; }
  GXY_CUDA(cudaMemcpy(v->d_vox, voxels, bytes, cudaMemcpyHostToDevice));
  d.vox = v->d_vox;
  *out = v;
  return 0;
}
void gxy_volume_destroy(gxy_volume *v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  cudaFree(v->d_vox);
  delete v;
}

int gxy_triangles_create(gxy_context *c, int nv, const float *verts, const float *normals, const float *data, int nt, const int *indices,
                         gxy_triangles **out) {
  if (use_device(c)) return 1;
  GXY_CHECK(nv >= 0 && nt >= 0 && (nt == 0 || (verts && indices)), "triangle mesh needs vertices and indices");
  gxy_triangles *t = new gxy_triangles();
  t->ctx = c; t->nv = nv; t->nt = nt;
  if (upload(&t->d_verts, verts, (size_t)nv * 3) || upload(&t->d_normals, normals, (size_t)nv * 3) || upload(&t->d_data, data, (size_t)nv) ||
      upload(&t->d_idx, indices, (size_t)nt * 3)) {
    delete t;
    return 1;
  }
  *out = t;
  return 0;
}
void gxy_triangles_destroy(gxy_triangles *t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaFree(t->d_verts); cudaFree(t->d_normals); cudaFree(t->d_data); cudaFree(t->d_idx);
  delete t;
}

int gxy_particles_create(gxy_context *c, int n, const float *centers, const float *data, gxy_particles **out) {
  if (use_device(c)) return 1;
  GXY_CHECK(n >= 0 && (n == 0 || centers), "particles need centres");
  gxy_particles *p = new gxy_particles();
  p->ctx = c; p->n = n;
  if (upload(&p->d_centers, centers, (size_t)n * 3) || upload(&p->d_data, data, (size_t)n)) { delete p; return 1; }
  *out = p;
  return 0;
}
void gxy_particles_destroy(gxy_particles *p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaFree(p->d_centers); cudaFree(p->d_data);
  delete p;
}

int gxy_pathlines_create(gxy_context *c, int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity,
                         gxy_pathlines **out) {
  if (use_device(c)) return 1;
  GXY_CHECK(n_verts >= 0 && n_segments >= 0 && (n_verts == 0 || verts) && (n_segments == 0 || connectivity), "pathlines need vertices and connectivity");
  for (int i = 0; i < n_segments; i++)
    GXY_CHECK(connectivity[i] >= 0 && connectivity[i] + 1 < n_verts, "pathlines: segment %d starts at vertex %d of %d", i, connectivity[i], n_verts);
  gxy_pathlines *p = new gxy_pathlines();
  p->ctx = c;
  p->verts.assign(verts, verts + 3 * (size_t)n_verts);
  p->has_data = data != nullptr;
  if (data) p->data.assign(data, data + n_verts);
  else p->data.assign((size_t)n_verts, 0.f);
  p->conn.assign(connectivity, connectivity + n_segments);
  *out = p;
  return 0;
}
void gxy_pathlines_destroy(gxy_pathlines *p) { delete p; }

// DataDrivenPathLines.cpp:28-37 (MAP_RADIUS; its `R <= 0.0` / `R >= 1.0` compare in double, exact for a float)
static inline float map_radius(float d, float radius0, float radius1, float value0, float value1) {
  if (value0 == value1) return radius0;
  const float R = (d - value0) / (value1 - value0);
  return (R <= 0.0f) ? radius0 : (R >= 1.0f) ? radius1 : radius0 + R * (radius1 - radius0);
}
// ospcommon lerp(factor, a, b) (ospmath.h:196-200)
static inline float os_lerp(float factor, float a, float b) { return (1.f - factor) * a + factor * b; }

int gxy_build_curves(int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity, float radius0,
                     float radius1, float value0, float value1, float *cp_out) {
  GXY_CHECK(n_segments == 0 || (verts && connectivity && cp_out), "gxy_build_curves: NULL argument");
  for (int i = 0; i < n_segments; i++)
    GXY_CHECK(connectivity[i] >= 0 && connectivity[i] + 1 < n_verts, "pathlines: segment %d starts at vertex %d of %d", i, connectivity[i], n_verts);
  // DataDrivenPathLines.cpp:103-156.  vertexCurve/indexCurve: a segment that continues pushes 3 control points, its 4th is
  // the first of the next segment (Embree reads 4 consecutive vertices from indexCurve[i]); a line end repeats its end point.
  struct P4 { float x, y, z, r; };
  std::vector<P4> vc;
  std::vector<size_t> ic((size_t)n_segments);
  vc.reserve(3 * (size_t)n_segments + 4);
  bool middle = false;
  float tx = 0.f, ty = 0.f, tz = 0.f;   // tangent
  for (int i = 0; i < n_segments; i++) {
    const int idx = connectivity[i];
    const float *s = verts + 3 * (size_t)idx, *e = s + 3;
    const float sx = s[0] - e[0], sy = s[1] - e[1], sz = s[2] - e[2];
    const float lengthSegment = sqrtf(sx * sx + sy * sy + sz * sz);
    const float startRadius = map_radius(data ? data[idx] : 0.f, radius0, radius1, value0, value1);
    const float endRadius = map_radius(data ? data[idx + 1] : 0.f, radius0, radius1, value0, value1);
    ic[i] = vc.size();
    vc.push_back(P4{s[0], s[1], s[2], startRadius});
    if (middle) vc.push_back(P4{s[0] + tx, s[1] + ty, s[2] + tz, os_lerp(1.f / 3, startRadius, endRadius)});
    else vc.push_back(P4{s[0], s[1], s[2], startRadius});
    middle = i + 1 < n_segments && connectivity[i + 1] == idx + 1;
    if (middle) {
      const float *n = e + 3;
      const float dx = (1.f / 3) * (n[0] - s[0]), dy = (1.f / 3) * (n[1] - s[1]), dz = (1.f / 3) * (n[2] - s[2]);
      const float nx = n[0] - e[0], ny = n[1] - e[1], nz = n[2] - e[2];
      const float b = sqrtf(nx * nx + ny * ny + nz * nz);
      const float r = lengthSegment / (lengthSegment + b);
      vc.push_back(P4{e[0] - r * dx, e[1] - r * dy, e[2] - r * dz, os_lerp(2.f / 3, startRadius, endRadius)});
      tx = (1.f - r) * dx; ty = (1.f - r) * dy; tz = (1.f - r) * dz;
    } else {
      vc.push_back(P4{e[0], e[1], e[2], endRadius});
      vc.push_back(P4{e[0], e[1], e[2], endRadius});
    }
  }
  for (int i = 0; i < n_segments; i++) memcpy(cp_out + 16 * (size_t)i, &vc[ic[i]], 16 * sizeof(float));
  return 0;
}

// ---- Visualization -----------------------------------------------------------------------------
int gxy_vis_create(gxy_context *c, gxy_vis **out) {
  if (use_device(c)) return 1;
  gxy_vis *v = new gxy_vis();
  v->ctx = c;
  for (int k = 0; k < 3; k++) v->gmin[k] = v->gmax[k] = v->lmin[k] = v->lmax[k] = 0.f;
  for (int k = 0; k < 6; k++) v->neighbors[k] = -1;
  memset(&v->P, 0, sizeof v->P);
  *out = v;
  return 0;
}

static void flight_destroy(gxy_vis *v, Flight *F);
static void vis_free_commit(gxy_vis *v) {
  if (v->d_tfs) cudaFree(v->d_tfs);
  if (v->d_geoms) cudaFree(v->d_geoms);
  if (v->d_error) cudaFree(v->d_error);
  if (v->bvh.nodes) cudaFree(v->bvh.nodes);
  if (v->bvh.prims) cudaFree(v->bvh.prims);
  for (float *p : v->d_curves) cudaFree(p);
  v->d_curves.clear();
  v->d_tfs = nullptr; v->d_geoms = nullptr; v->d_error = nullptr;
  v->bvh = BvhResult();
  v->committed = false;
}

void gxy_vis_destroy(gxy_vis *v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  vis_free_commit(v);
  v->cur.release(); v->next.release(); v->send.release(); v->recv.release(); v->hits.release(); v->fq.release(); v->rawhits.release();
  v->hit_index.release(); v->block_sums.release(); v->small.release(); v->counters.release();
  if (v->d_samples) cudaFree(v->d_samples);
  if (v->d_sample_count) cudaFree(v->d_sample_count);
  if (v->prog_image) cudaFree(v->prog_image);
  if (v->prog_kbuffer) cudaFree(v->prog_kbuffer);
  if (v->prog_touched) cudaFree(v->prog_touched);
  v->fb.release(); v->fb_tmp.release(); v->rgba8.release(); v->io_f.release(); v->io_i.release();
  if (v->ctx->copy_stream) cudaStreamSynchronize(v->ctx->copy_stream);
  for (int s = 0; s < 2; s++) {
    v->rgba8_async[s].release();
    if (v->async_ready[s]) cudaEventDestroy(v->async_ready[s]);
    if (v->async_done[s]) cudaEventDestroy(v->async_done[s]);
  }
  v->proxies.release();
  if (v->h_tail) cudaFreeHost(v->h_tail);
  for (int k = 0; k < GXY_MAX_FLIGHTS; k++) flight_destroy(v, v->flights[k]);
  for (int k = 0; k < GXY_LIST_LANES; k++)
    if (ListLane *l = v->list_lanes[k]) {
      if (l->st) { cudaStreamSynchronize(l->st); cudaStreamDestroy(l->st); }
      l->cur.release(); l->next.release(); l->hit_index.release(); l->block_sums.release(); l->small.release(); l->io_i.release();
      l->err.release(); l->io_f.release();
      delete l;
    }
  delete v;
}

int gxy_vis_set_partition(gxy_vis *v, const float gmin[3], const float gmax[3], const float lmin[3], const float lmax[3],
                          const int neighbors[6]) {
  for (int k = 0; k < 3; k++) { v->gmin[k] = gmin[k]; v->gmax[k] = gmax[k]; v->lmin[k] = lmin[k]; v->lmax[k] = lmax[k]; }
  for (int k = 0; k < 6; k++) v->neighbors[k] = neighbors[k];
  v->committed = false;
  return 0;
}

int gxy_vis_add_volume(gxy_vis *v, gxy_volume *vol, int n_slices, const float *slices4, int n_iso, const float *isovalues, int volume_render,
                       const gxy_transfer_function *tf) {
  GXY_CHECK(vol && tf, "gxy_vis_add_volume: NULL argument");
  GXY_CHECK((int)v->vols.size() < GXY_MAX_VOLUME_VIS, "more than %d volume operators in one Visualization", GXY_MAX_VOLUME_VIS);
  GXY_CHECK(n_slices <= GXY_MAX_SLICES && n_iso <= GXY_MAX_ISOVALUES, "too many slices (%d > %d) or isovalues (%d > %d)", n_slices,
            GXY_MAX_SLICES, n_iso, GXY_MAX_ISOVALUES);
  GXY_CHECK(vol->ctx == v->ctx, "volume belongs to another context");
  VolOp op;
  op.vol = vol;
  op.slices.assign(slices4, slices4 + 4 * (size_t)n_slices);
  op.iso.assign(isovalues, isovalues + n_iso);
  op.volume_render = volume_render;
  op.tf = *tf;
  v->vols.push_back(op);
  v->committed = false;
  return 0;
}

int gxy_vis_add_triangles(gxy_vis *v, gxy_triangles *t, const gxy_transfer_function *tf) {
  GXY_CHECK(t && tf, "gxy_vis_add_triangles: NULL argument");
  GXY_CHECK(t->ctx == v->ctx, "mesh belongs to another context");
  GeomOp g;
  memset(&g, 0, sizeof g);
  g.kind = 0; g.tri = t; g.tf = *tf;
  v->geoms.push_back(g);
  v->committed = false;
  return 0;
}

int gxy_vis_add_particles(gxy_vis *v, gxy_particles *p, float radius0, float radius1, float value0, float value1,
                          const gxy_transfer_function *tf) {
  GXY_CHECK(p && tf, "gxy_vis_add_particles: NULL argument");
  GXY_CHECK(p->ctx == v->ctx, "particles belong to another context");
  GeomOp g;
  memset(&g, 0, sizeof g);
  g.kind = 1; g.par = p; g.tf = *tf;
  g.radius0 = radius0; g.radius1 = radius1; g.value0 = value0; g.value1 = value1;
  v->geoms.push_back(g);
  v->committed = false;
  return 0;
}

int gxy_vis_add_sampler(gxy_vis *v, gxy_volume *vol, int kind, float param) {
  GXY_CHECK(v && vol, "gxy_vis_add_sampler: NULL argument");
  GXY_CHECK(kind == GXY_SAMPLER_GRADIENT || kind == GXY_SAMPLER_ISO, "unknown sampler kind %d", kind);
  GXY_CHECK((int)v->samplers.size() < GXY_MAX_VOLUME_VIS, "more than %d sampler operators in one Visualization", GXY_MAX_VOLUME_VIS);
  GXY_CHECK(vol->ctx == v->ctx, "volume belongs to another context");
  SamplerOp op;
  op.vol = vol; op.kind = kind; op.param = param;
  v->samplers.push_back(op);
  v->committed = false;
  return 0;
}

int gxy_vis_add_pathlines(gxy_vis *v, gxy_pathlines *p, float radius0, float radius1, float value0, float value1,
                          const gxy_transfer_function *tf) {
  GXY_CHECK(p && tf, "gxy_vis_add_pathlines: NULL argument");
  GXY_CHECK(p->ctx == v->ctx, "pathlines belong to another context");
  GeomOp g;
  memset(&g, 0, sizeof g);
  g.kind = 2; g.pl = p; g.tf = *tf;
  g.radius0 = radius0; g.radius1 = radius1; g.value0 = value0; g.value1 = value1;
  v->geoms.push_back(g);
  v->committed = false;
  return 0;
}

static void pack_tf(const gxy_transfer_function &in, DevTF &out) {
  for (int i = 0; i < 256; i++) out.e[i] = make_float4(in.colors[i][0], in.colors[i][1], in.colors[i][2], in.opacities[i]);
  out.lo = in.range_lo; out.hi = in.range_hi; out.pad0 = out.pad1 = 0.f;
}

int gxy_vis_commit(gxy_vis *v) {
  if (use_device(v->ctx)) return 1;
  GXY_CHECK(v->samplers.empty() || (v->vols.empty() && v->geoms.empty()),
            "a sampling Visualization holds only sampler operators (SamplerTraceRays calls every operator as a SamplerVis)");
  vis_free_commit(v);
  // SamplerTraceRays.ispc:136-150: step = min samplingStep*samplingRate over the operators
  memset(&v->SP, 0, sizeof v->SP);
  v->SP.n_ops = (int)v->samplers.size();
  v->SP.lmin = make_float3(v->lmin[0], v->lmin[1], v->lmin[2]); v->SP.lmax = make_float3(v->lmax[0], v->lmax[1], v->lmax[2]);
  for (size_t m = 0; m < v->samplers.size(); m++) {
    v->SP.op[m].kind = v->samplers[m].kind; v->SP.op[m].param = v->samplers[m].param; v->SP.op[m].vol = v->samplers[m].vol->dv;
    const float s = v->samplers[m].vol->dv.samplingStep * v->samplers[m].vol->dv.samplingRate;
    if (m == 0 || s < v->SP.step) v->SP.step = s;
  }
  SceneParams &P = v->P;
  memset(&P, 0, sizeof P);
  P.gmin = make_float3(v->gmin[0], v->gmin[1], v->gmin[2]); P.gmax = make_float3(v->gmax[0], v->gmax[1], v->gmax[2]);
  P.lmin = make_float3(v->lmin[0], v->lmin[1], v->lmin[2]); P.lmax = make_float3(v->lmax[0], v->lmax[1], v->lmax[2]);
  for (int k = 0; k < 6; k++) P.neighbors[k] = v->neighbors[k];
  P.n_volvis = (int)v->vols.size();
  P.n_geoms = (int)v->geoms.size();
  std::vector<DevTF> tfs(v->vols.size() + v->geoms.size());
  // TraceRays.ispc:342-361
  P.integrate = 0;
  P.step = -1.f;
  v->has_dvr = false;
  for (size_t m = 0; m < v->vols.size(); m++) {
    const VolOp &op = v->vols[m];
    pack_tf(op.tf, tfs[m]);
    op.vol->last_tf = (int)m;  // MappedVis.cpp:206-212: last committed Vis owns the volume object's TF
    if (op.volume_render) { P.integrate = 1; v->has_dvr = true; }
    if (!op.iso.empty()) P.integrate = 1;
    const float s = op.vol->dv.samplingStep * op.vol->dv.samplingRate;
    if (P.step < 0 || P.step > s) P.step = s;
  }
  for (size_t m = 0; m < v->vols.size(); m++) {
    const VolOp &op = v->vols[m];
    DevVolVis &d = P.vv[m];
    d.n_slices = (int)op.slices.size() / 4;
    d.n_iso = (int)op.iso.size();
    d.volume_render = op.volume_render;
    d.tf = (int)m;
    for (int k = 0; k < d.n_slices; k++) d.slices[k] = make_float4(op.slices[4 * k], op.slices[4 * k + 1], op.slices[4 * k + 2], op.slices[4 * k + 3]);
    for (int k = 0; k < d.n_iso; k++) d.iso[k] = op.iso[k];
    d.vol = op.vol->dv;
    d.vol.tf = op.vol->last_tf;
  }
  std::vector<DevGeom> dg(v->geoms.size());
  std::vector<GeomBuildInput> bi(v->geoms.size());
  for (size_t k = 0; k < v->geoms.size(); k++) {
    const GeomOp &g = v->geoms[k];
    pack_tf(g.tf, tfs[v->vols.size() + k]);
    DevGeom &d = dg[k];
    memset(&d, 0, sizeof d);
    GeomBuildInput &b = bi[k];
    memset(&b, 0, sizeof b);
    d.kind = g.kind;
    d.tf = (int)(v->vols.size() + k);
    b.kind = g.kind;
    b.geom_id = (int)k;
    if (g.kind == 0) {
      d.idx = g.tri->d_idx; d.normals = g.tri->d_normals; d.data = g.tri->d_data;
      b.n_prims = g.tri->nt; b.verts = g.tri->d_verts; b.idx = g.tri->d_idx;
    } else if (g.kind == 2) {
      // DataDrivenPathLines::finalize runs at commit in the reference too (the radii depend on this Vis)
      const int nseg = (int)g.pl->conn.size();
      std::vector<float> cp(16 * (size_t)nseg);
      if (gxy_build_curves((int)(g.pl->verts.size() / 3), g.pl->verts.data(), g.pl->has_data ? g.pl->data.data() : nullptr, nseg,
                           g.pl->conn.data(), g.radius0, g.radius1, g.value0, g.value1, cp.data()))
        return 1;
      float *d_cp = nullptr;
      GXY_CUDA(cudaMalloc(&d_cp, sizeof(float) * (cp.empty() ? 16 : cp.size())));
      v->d_curves.push_back(d_cp);
      if (!cp.empty()) GXY_CUDA(cudaMemcpy(d_cp, cp.data(), sizeof(float) * cp.size(), cudaMemcpyHostToDevice));
      d.centers = d_cp;
      d.radius0 = g.radius0; d.radius1 = g.radius1; d.value0 = g.value0; d.value1 = g.value1;
      b.n_prims = nseg; b.centers = d_cp;
      P.n_curves += nseg;
    } else {
      d.centers = g.par->d_centers; d.data = g.par->d_data;
      d.radius0 = g.radius0; d.radius1 = g.radius1; d.value0 = g.value0; d.value1 = g.value1;
      // DataDrivenSpheres.ispc:211-220
      float eps = logf(g.radius0);
      if (eps < 0.f) eps = -1.f / eps;
      if (eps > (float)(g.radius0 / 100.0)) eps = (float)(g.radius0 / 100.0);
      d.epsilon = eps;
      b.n_prims = g.par->n; b.centers = g.par->d_centers; b.data = g.par->d_data;
      b.radius0 = g.radius0; b.radius1 = g.radius1; b.value0 = g.value0; b.value1 = g.value1; b.epsilon = eps;
    }
  }
  if (!tfs.empty()) {
    GXY_CUDA(cudaMalloc(&v->d_tfs, sizeof(DevTF) * tfs.size()));
    GXY_CUDA(cudaMemcpy(v->d_tfs, tfs.data(), sizeof(DevTF) * tfs.size(), cudaMemcpyHostToDevice));
  }
  if (!dg.empty()) {
    GXY_CUDA(cudaMalloc(&v->d_geoms, sizeof(DevGeom) * dg.size()));
    GXY_CUDA(cudaMemcpy(v->d_geoms, dg.data(), sizeof(DevGeom) * dg.size(), cudaMemcpyHostToDevice));
  }
  // [0] error flag (int) [1] ray-queue head (unsigned) [2..5] traversal counters (2 x u64)
  GXY_CUDA(cudaMalloc(&v->d_error, 32));
  GXY_CUDA(cudaMemset(v->d_error, 0, 32));
  if (!bi.empty()) {
    if (build_bvh(bi.data(), (int)bi.size(), &v->bvh, v->ctx->stream)) return 1;
  }
  P.tfs = v->d_tfs;
  P.geoms = v->d_geoms;
  P.nodes = v->bvh.nodes;
  P.prims = v->bvh.prims;
  P.n_prims = v->bvh.n_prims;
  P.error_flag = v->d_error;
  P.work_counter = reinterpret_cast<unsigned *>(v->d_error) + 1;
  P.trav_counters = reinterpret_cast<unsigned long long *>(v->d_error + 2);
  v->committed = true;
  return 0;
}

int gxy_vis_build_info(gxy_vis *v, long long *n_prims, long long *n_nodes, float *build_ms) {
  if (n_prims) *n_prims = v->bvh.n_prims;
  if (n_nodes) *n_nodes = v->bvh.n_nodes;
  if (build_ms) *build_ms = v->bvh.build_ms;
  return 0;
}
int gxy_vis_build_times(gxy_vis *v, float *build_ms, float *alloc_host_ms) {
  GXY_CHECK(v, "NULL visualization");
  if (build_ms) *build_ms = v->bvh.build_ms;
  if (alloc_host_ms) *alloc_host_ms = v->bvh.alloc_host_ms;
  return 0;
}

// ---- host helpers ------------------------------------------------------------------------------
struct H3 { float x, y, z; };
static inline void h_normalize(H3 &a) {  // src/data/dtypes.h normalize(vec3f&)
  float d = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  if (d != 0) { d = (float)(1.0 / d); a.x *= d; a.y *= d; a.z *= d; }
}
static inline H3 h_cross(H3 a, H3 b) { return H3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

int gxy_resolve_lights(const gxy_lighting *in, const gxy_camera *cam, gxy_lighting *out) {
  // Rendering::resolve_lights (Rendering.cpp:157-216)
  GXY_CHECK(in && cam && out, "gxy_resolve_lights: NULL argument");
  GXY_CHECK(in->n_lights >= 0 && in->n_lights <= GXY_MAX_LIGHTS, "too many lights");
  *out = *in;
  H3 viewpoint{cam->eye[0], cam->eye[1], cam->eye[2]}, viewup{cam->up[0], cam->up[1], cam->up[2]}, viewdir{cam->dir[0], cam->dir[1], cam->dir[2]};
  h_normalize(viewup);
  h_normalize(viewdir);
  H3 right = h_cross(viewdir, viewup);
  h_normalize(right);
  H3 up = h_cross(viewdir, right);
  for (int i = 0; i < in->n_lights; i++)
    if (in->types[i] == 1) {
      const float lx = in->lights[i][0], ly = in->lights[i][1], lz = in->lights[i][2];
      // viewpoint + scale1(x,right) + scale1(y,up) + scale1(z,viewdir), left to right
      out->lights[i][0] = ((viewpoint.x + lx * right.x) + ly * up.x) + lz * viewdir.x;
      out->lights[i][1] = ((viewpoint.y + lx * right.y) + ly * up.y) + lz * viewdir.y;
      out->lights[i][2] = ((viewpoint.z + lx * right.z) + ly * up.z) + lz * viewdir.z;
      out->types[i] = 2;
    }
  return 0;
}

int gxy_resample_transfer_function(int n, const float *cmap, int m, const float *omap, gxy_transfer_function *out) {
  // MappedVis::local_commit (MappedVis.cpp:277-338); x is evaluated in double there
  GXY_CHECK(n >= 2 && m >= 2 && cmap && omap && out, "transfer function needs >= 2 control points");
  int i0 = 0, i1 = 1;
  float xmin = cmap[0], xmax = cmap[4 * (n - 1)];
  for (int i = 0; i < 256; i++) {
    float x = (float)(xmin + (i / (255.0)) * (xmax - xmin));
    if (x > xmax) x = xmax;
    while (cmap[4 * i1] < x) i0++, i1++;
    const float d = (x - cmap[4 * i0]) / (cmap[4 * i1] - cmap[4 * i0]);
    for (int c = 0; c < 3; c++) out->colors[i][c] = cmap[4 * i0 + 1 + c] + d * (cmap[4 * i1 + 1 + c] - cmap[4 * i0 + 1 + c]);
  }
  i0 = 0, i1 = 1;
  xmin = omap[0]; xmax = omap[2 * (m - 1)];
  for (int i = 0; i < 256; i++) {
    float x = (float)(xmin + (i / (255.0)) * (xmax - xmin));
    if (x > xmax) x = xmax;
    while (omap[2 * i1] < x) i0++, i1++;
    const float d = (x - omap[2 * i0]) / (omap[2 * i1] - omap[2 * i0]);
    out->opacities[i] = omap[2 * i0 + 1] + d * (omap[2 * i1 + 1] - omap[2 * i0 + 1]);
  }
  return 0;
}

void gxy_factor(int ijk, int factors[3]) {
  // Volume.cpp:88-122
  factors[0] = factors[1] = 1; factors[2] = ijk;
  if (ijk == 1) { factors[2] = 1; return; }
  int mm = ijk + 3;
  for (int i = 1; i <= ijk >> 1; i++) {
    const int jk = ijk / i;
    if (ijk == (i * jk))
      for (int j = 1; j <= jk >> 1; j++) {
        const int k = jk / j;
        if (jk == (j * k)) {
          const int m = i + j + k;
          if (m < mm) { mm = m; factors[0] = i; factors[1] = j; factors[2] = k; }
        }
      }
  }
}

void gxy_partition(int n, const int factors[3], const int grid[3], int *out) {
  // Volume.cpp:133-172
  (void)n;
  const int ni = grid[0] - 2, nj = grid[1] - 2, nk = grid[2] - 2;
  const int di = ni / factors[0], dj = nj / factors[1], dk = nk / factors[2];
  int *p = out;
  for (int k = 0; k < factors[2]; k++)
    for (int j = 0; j < factors[1]; j++)
      for (int i = 0; i < factors[0]; i++, p += 15) {
        p[0] = i; p[1] = j; p[2] = k;
        p[3] = 1 + i * di; p[4] = 1 + j * dj; p[5] = 1 + k * dk;
        p[6] = 1 + ((i == (factors[0] - 1)) ? ni - p[3] : di);
        p[7] = 1 + ((j == (factors[1] - 1)) ? nj - p[4] : dj);
        p[8] = 1 + ((k == (factors[2] - 1)) ? nk - p[5] : dk);
        p[9] = p[3] - 1; p[10] = p[4] - 1; p[11] = p[5] - 1;
        p[12] = p[6] + 2; p[13] = p[7] + 2; p[14] = p[8] + 2;
      }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
static DevLights make_dev_lights(const gxy_lighting &L) {
  DevLights d;
  memset(&d, 0, sizeof d);
  d.n_lights = L.n_lights; d.n_ao = L.n_ao; d.shadows = L.shadows;
  d.ao_radius = L.ao_radius; d.Ka = L.Ka; d.Kd = L.Kd;
  for (int i = 0; i < L.n_lights; i++) {
    for (int k = 0; k < 3; k++) d.lights[i][k] = L.lights[i][k];
    d.types[i] = L.types[i];
  }
  return d;
}

static DevCamera make_dev_camera(const gxy_camera &c, int width, int height) {
  // Camera::generate_initial_rays set-up (Camera.cpp:528-582); tan() in double as the reference
  DevCamera f;
  H3 veye{c.eye[0], c.eye[1], c.eye[2]}, vu{c.up[0], c.up[1], c.up[2]}, vdir{c.dir[0], c.dir[1], c.dir[2]};
  H3 center;
  if (c.aov == 0.0f) {
    center = H3{veye.x + vdir.x, veye.y + vdir.y, veye.z + vdir.z};
    h_normalize(vdir);
  } else {
    const float d = (float)(1.0 / tan(2 * 3.1415926 * ((double)c.aov / 2.0) / 360.0));
    h_normalize(vdir);
    center = H3{veye.x + vdir.x * d, veye.y + vdir.y * d, veye.z + vdir.z * d};
  }
  H3 vr = h_cross(vdir, vu);
  h_normalize(vr);
  vu = h_cross(vr, vdir);
  h_normalize(vu);
  const float pixel_scaling = (float)((((width < height) ? width : height) - 1.0) / 2.0);
  f.off_x = (float)((width - 1) / 2.0);
  f.off_y = (float)((height - 1) / 2.0);
  f.scaling = (float)(1.0 / pixel_scaling);  // Camera.h:66
  f.veye = make_float3(veye.x, veye.y, veye.z); f.vdir = make_float3(vdir.x, vdir.y, vdir.z);
  f.vr = make_float3(vr.x, vr.y, vr.z); f.vu = make_float3(vu.x, vu.y, vu.z);
  f.center = make_float3(center.x, center.y, center.z);
  f.ortho = (c.aov == 0.0f) ? 1 : 0;
  return f;
}

static int check_vis(gxy_vis *v) {
  GXY_CHECK(v, "NULL visualization");
  GXY_CHECK(v->committed, "visualization not committed (call gxy_vis_commit)");
  return use_device(v->ctx);
}

static int check_error_flag(gxy_vis *v) {
  int e = 0;
  GXY_CUDA(cudaMemcpyAsync(&e, v->d_error, sizeof(int), cudaMemcpyDeviceToHost, v->ctx->stream));
  GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
  if (e != 0) GXY_CUDA(cudaMemsetAsync(v->d_error, 0, sizeof(int), v->ctx->stream));  // reported once
  GXY_CHECK(e == 0, "device error flag %d (1: BVH traversal stack overflow, 3: ray list / inbox capacity, 4: peer barrier timeout, 6: TMA copy did not complete, 7: translucent surface on the fused path)", e);
  return 0;
}

static int h2d_rays(cudaStream_t st, RayBuf &buf, gxy_raylist_view r) {
  if (buf.reserve((size_t)std::max(r.n, 1), false, st)) return 1;
  if (r.n > 0)
    GXY_CUDA(cudaMemcpy2DAsync(buf.base, buf.cap * 4, r.base, (size_t)r.aligned_n * 4, (size_t)r.n * 4, GXY_RAYLIST_COLUMNS,
                               cudaMemcpyHostToDevice, st));
  return 0;
}
static int d2h_rays(cudaStream_t st, RayBuf &buf, float *base, int n, int aligned_n, int first_col = 0, int ncols = GXY_RAYLIST_COLUMNS) {
  if (n > 0)
    GXY_CUDA(cudaMemcpy2DAsync(base + (size_t)first_col * aligned_n, (size_t)aligned_n * 4, buf.base + (size_t)first_col * buf.cap, buf.cap * 4,
                               (size_t)n * 4, ncols, cudaMemcpyDeviceToHost, st));
  return 0;
}

// a free lane of the Visualization for the duration of one call (waits if all GXY_LIST_LANES are taken)
struct LaneGuard {
  gxy_vis *v;
  ListLane *l = nullptr;
  explicit LaneGuard(gxy_vis *vis) : v(vis) {
    std::unique_lock<std::mutex> lk(v->lanes_mu);
    for (;;) {
      for (int k = 0; k < GXY_LIST_LANES && !l; k++) {
        if (!v->list_lanes[k]) v->list_lanes[k] = new ListLane();
        if (!v->list_lanes[k]->busy) l = v->list_lanes[k];
      }
      if (l) break;
      v->lanes_cv.wait(lk);
    }
    l->busy = true;
  }
  ~LaneGuard() {
    {
      std::lock_guard<std::mutex> lk(v->lanes_mu);
      l->busy = false;
    }
    v->lanes_cv.notify_one();
  }
  // stream and error flag are created by the first call that uses the lane (the device is current: check_vis)
  int prepare() {
    if (!l->st) GXY_CUDA(cudaStreamCreateWithFlags(&l->st, cudaStreamNonBlocking));
    if (!l->err.p) {
      if (l->err.reserve(8)) return 1;
      GXY_CUDA(cudaMemsetAsync(l->err.p, 0, sizeof(int) * 8, l->st));
    }
    return 0;
  }
  // the scene as this call's kernels see it: the lane's own error flag and its own ray-queue head (the persistent trace kernel
  // pulls rays through SceneParams::work_counter; two concurrent launches must not share one)
  SceneParams params() const {
    SceneParams P = v->P;
    P.error_flag = l->err.p;
    P.work_counter = reinterpret_cast<unsigned *>(l->err.p + 4);
    return P;
  }
  int check_error() {
    int e = 0;
    GXY_CUDA(cudaMemcpyAsync(&e, l->err.p, sizeof(int), cudaMemcpyDeviceToHost, l->st));
    GXY_CUDA(cudaStreamSynchronize(l->st));
    if (e != 0) GXY_CUDA(cudaMemsetAsync(l->err.p, 0, sizeof(int), l->st));  // reported once
    GXY_CHECK(e == 0, "device error flag %d (1: BVH traversal stack overflow, 3: ray list / inbox capacity, 4: peer barrier timeout, 6: TMA copy did not complete, 7: translucent surface on the fused path)", e);
    return 0;
  }
};

// TraceRays::Trace on a list that is already on the device (lane stream): trace in place, hit scan, AO/shadow spawn into `sec`
// (reserved here); *n_sec = rays spawned.  d_hits may be NULL.
static int trace_list_on_lane(gxy_vis *v, LaneGuard &G, const gxy_lighting *lights, RayBuf &rays, int n, float epsilon, int *d_hits, RayBuf &sec,
                              long long *n_sec) {
  ListLane &l = *G.l;
  cudaStream_t st = l.st;
  const SceneParams P = G.params();
  if (launch_trace(P, rays.v, n, epsilon, d_hits, false, nullptr, st)) return 1;
  if (l.hit_index.reserve((size_t)2 * n) || l.block_sums.reserve((size_t)n / 1024 + 2) || l.small.reserve(64)) return 1;
  if (launch_hit_scan(rays.v, n, l.hit_index.p, l.block_sums.p, l.small.p, st)) return 1;
  int nhit = 0;
  GXY_CUDA(cudaMemcpyAsync(&nhit, l.small.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaStreamSynchronize(st));
  const long long n_out = (long long)nhit * (lights->n_ao + (lights->shadows ? lights->n_lights : 0));
  GXY_CHECK(n_out < (1ll << 31), "secondary ray list too large (%lld)", n_out);
  const DevLights L = make_dev_lights(*lights);
  if (sec.reserve((size_t)std::max<long long>(n_out, 1), false, st)) return 1;
  // the spawn kernels write the 16 columns generateAORays / the shadow-ray loop set (TraceRays.ispc:645-857); the others (normal,
  // surface colour, sample, classification) are undefined in the reference's fresh RayList -- zero here, so that a list handed
  // back to the caller is a function of its inputs alone and carries no stale device memory
  if (n_out > 0) GXY_CUDA(cudaMemset2DAsync(sec.base, sec.cap * 4, 0, (size_t)n_out * 4, GXY_RAYLIST_COLUMNS, st));
  if (launch_shade_spawn(L, rays.v, n, l.hit_index.p, l.small.p, sec.v, epsilon, st)) return 1;
  *n_sec = n_out;
  (void)v;
  return 0;
}

extern "C" {

int gxy_trace_raylist(gxy_vis *v, const gxy_lighting *lights, gxy_raylist_view rays, float epsilon, gxy_raylist **out, int *hit_ids) {
  if (check_vis(v)) return 1;
  GXY_CHECK(lights && rays.base && rays.n >= 0 && rays.aligned_n >= rays.n, "gxy_trace_raylist: bad arguments");
  GXY_CHECK(lights->n_lights >= 1 && lights->n_lights <= GXY_MAX_LIGHTS, "lighting needs 1..%d lights", GXY_MAX_LIGHTS);
  if (out) *out = nullptr;
  const int n = rays.n;
  if (n == 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  ListLane &l = *G.l;
  cudaStream_t st = l.st;
  if (h2d_rays(st, l.cur, rays)) return 1;
  int *d_hits = nullptr;
  if (hit_ids) {
    if (l.io_i.reserve((size_t)2 * n)) return 1;
    d_hits = l.io_i.p;
  }
  long long n_out = 0;
  if (trace_list_on_lane(v, G, lights, l.cur, n, epsilon, d_hits, l.next, &n_out)) return 1;
  if (d2h_rays(st, l.cur, rays.base, n, rays.aligned_n)) return 1;
  if (hit_ids) GXY_CUDA(cudaMemcpyAsync(hit_ids, d_hits, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, st));
  if (n_out > 0 && out) {
    gxy_raylist *o = new gxy_raylist();
    o->n = (int)n_out;
    o->aligned_n = std::max(16, (o->n + 15) & ~15);  // Rays.cpp:57-58
    o->base = (float *)calloc((size_t)GXY_RAYLIST_COLUMNS * o->aligned_n, sizeof(float));
    if (!o->base) { delete o; gxy_set_error("out of host memory for the secondary list"); return 1; }
    if (d2h_rays(st, l.next, o->base, o->n, o->aligned_n)) { free(o->base); delete o; return 1; }
    *out = o;
  }
  GXY_CUDA(cudaStreamSynchronize(st));
  return G.check_error();
}

int gxy_raylist_get_view(gxy_raylist *l, gxy_raylist_view *view) {
  GXY_CHECK(l && view, "NULL raylist");
  view->base = l->base; view->n = l->n; view->aligned_n = l->aligned_n;
  return 0;
}
void gxy_raylist_free(gxy_raylist *l) {
  if (!l) return;
  free(l->base);
  delete l;
}

// ---- device-resident RayLists ---------------------------------------------------------------------
int gxy_raylist_upload(gxy_vis *v, gxy_raylist_view rays, gxy_dev_raylist **out) {
  if (check_vis(v)) return 1;
  GXY_CHECK(out && rays.n >= 0 && rays.aligned_n >= rays.n && (rays.base || rays.n == 0), "gxy_raylist_upload: bad arguments");
  LaneGuard G(v);
  if (G.prepare()) return 1;
  std::unique_ptr<gxy_dev_raylist> d(new gxy_dev_raylist());
  d->ctx = v->ctx;
  d->n = rays.n;
  if (h2d_rays(G.l->st, d->buf, rays)) { d->buf.release(); return 1; }
  GXY_CUDA(cudaStreamSynchronize(G.l->st));
  *out = d.release();
  return 0;
}
int gxy_raylist_size(gxy_dev_raylist *d, int *n) {
  GXY_CHECK(d && n, "gxy_raylist_size: NULL argument");
  *n = d->n;
  return 0;
}
int gxy_raylist_download(gxy_dev_raylist *d, gxy_raylist_view rays) {
  GXY_CHECK(d && rays.aligned_n >= d->n && (rays.base || d->n == 0), "gxy_raylist_download: the host list is too small (%d rays)", d ? d->n : 0);
  if (use_device(d->ctx)) return 1;
  if (d->n > 0)  // (the default stream of the library's context: the list is at rest between calls)
    GXY_CUDA(cudaMemcpy2D(rays.base, (size_t)rays.aligned_n * 4, d->buf.base, d->buf.cap * 4, (size_t)d->n * 4, GXY_RAYLIST_COLUMNS,
                          cudaMemcpyDeviceToHost));
  return 0;
}
void gxy_dev_raylist_free(gxy_dev_raylist *d) {
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  d->buf.release();
  delete d;
}
int gxy_trace_raylist_dev(gxy_vis *v, const gxy_lighting *lights, gxy_dev_raylist *rays, float epsilon, gxy_dev_raylist **secondary) {
  if (check_vis(v)) return 1;
  GXY_CHECK(lights && rays, "gxy_trace_raylist_dev: bad arguments");
  GXY_CHECK(lights->n_lights >= 1 && lights->n_lights <= GXY_MAX_LIGHTS, "lighting needs 1..%d lights", GXY_MAX_LIGHTS);
  if (secondary) *secondary = nullptr;
  if (rays->n == 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  std::unique_ptr<gxy_dev_raylist> sec(new gxy_dev_raylist());
  sec->ctx = v->ctx;
  long long n_out = 0;
  if (trace_list_on_lane(v, G, lights, rays->buf, rays->n, epsilon, nullptr, sec->buf, &n_out)) { sec->buf.release(); return 1; }
  GXY_CUDA(cudaStreamSynchronize(G.l->st));
  if (G.check_error()) { sec->buf.release(); return 1; }
  sec->n = (int)n_out;
  if (n_out > 0 && secondary) *secondary = sec.release();
  else sec->buf.release();
  return 0;
}
int gxy_classify_dev(gxy_vis *v, gxy_dev_raylist *rays) {
  if (check_vis(v)) return 1;
  GXY_CHECK(rays, "gxy_classify_dev: NULL list");
  if (rays->n == 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  if (launch_classify(G.params(), rays->buf.v, rays->n, G.l->st)) return 1;
  GXY_CUDA(cudaStreamSynchronize(G.l->st));
  return 0;
}

int gxy_classify(gxy_vis *v, gxy_raylist_view rays) {
  if (check_vis(v)) return 1;
  if (rays.n == 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  if (h2d_rays(G.l->st, G.l->cur, rays)) return 1;
  if (launch_classify(G.params(), G.l->cur.v, rays.n, G.l->st)) return 1;
  if (d2h_rays(G.l->st, G.l->cur, rays.base, rays.n, rays.aligned_n, 24, 1)) return 1;
  GXY_CUDA(cudaStreamSynchronize(G.l->st));
  return 0;
}

int gxy_sample_raylist(gxy_vis *v, gxy_raylist_view rays) {
  if (check_vis(v)) return 1;
  if (rays.n == 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  if (h2d_rays(G.l->st, G.l->cur, rays)) return 1;
  if (launch_sampler_trace(v->SP, G.l->cur.v, rays.n, nullptr, nullptr, 0, nullptr, false, G.l->st)) return 1;
  if (d2h_rays(G.l->st, G.l->cur, rays.base, rays.n, rays.aligned_n)) return 1;
  GXY_CUDA(cudaStreamSynchronize(G.l->st));
  return 0;
}

// room for `need` samples, keeping the ones already collected
static int reserve_samples(gxy_vis *v, unsigned long long need, cudaStream_t st) {
  if (need <= v->samples_cap) return 0;
  unsigned long long ncap = std::max(need, v->samples_cap + v->samples_cap / 2);
  ncap = (ncap + 1023ull) & ~1023ull;
  float *nb = nullptr;
  GXY_CUDA(cudaMalloc(&nb, sizeof(float) * 3 * ncap));
  if (v->d_samples && v->n_samples) {
    GXY_CUDA(cudaMemcpyAsync(nb, v->d_samples, sizeof(float) * 3 * v->n_samples, cudaMemcpyDeviceToDevice, st));
    GXY_CUDA(cudaStreamSynchronize(st));
  }
  if (v->d_samples) cudaFree(v->d_samples);
  v->d_samples = nb;
  v->samples_cap = ncap;
  return 0;
}

// Sampler over a frame (Sampler.cpp:52-133 on top of Renderer::local_render / processRays): per wave and partition
//   sample-trace (+ sample points) -> Classify -> counting sort by destination, KEEP_HERE rays (the ones that left a sample)
//   back into the partition's own next list, BOUNDARY rays to the neighbour; until no partition has rays left.
// One exchange step of the list path across processes (one process per GPU): all-gather of the per-destination counts, then the
// 24 ray columns of every non-empty (source, destination) pair as NCCL send/recv on the context's stream, and the received rays
// plus the ones that stay (destination = this rank) appended to v->next at *n_next.  *global_pending = rays alive anywhere after it.
static int nccl_exchange_rays(gxy_vis *v, int nranks, int me, const int *send_counts, const int *send_offsets, int *n_next, long long *global_pending,
                              gxy_stats &S) {
  cudaStream_t st = v->ctx->stream;
  if (v->io_i.reserve((size_t)nranks * (nranks + 1))) return 1;
  int *d_mine = v->io_i.p, *d_all = v->io_i.p + nranks;
  GXY_CUDA(cudaMemcpyAsync(d_mine, send_counts, sizeof(int) * nranks, cudaMemcpyHostToDevice, st));
  GXY_NCCL(g_nccl.AllGather(d_mine, d_all, nranks, ncclInt32, v->ctx->comm, st));
  std::vector<int> all((size_t)nranks * nranks, 0);
  GXY_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(int) * nranks * nranks, cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaStreamSynchronize(st));
  std::vector<int> recv_from(nranks, 0);
  long long pending = 0;
  int n_recv = 0;
  for (int r = 0; r < nranks; r++) {
    for (int d = 0; d < nranks; d++) pending += all[(size_t)r * nranks + d];
    recv_from[r] = all[(size_t)r * nranks + me];
    if (r != me) n_recv += recv_from[r];
  }
  for (int d = 0; d < nranks; d++)
    if (d != me) S.forwarded_rays += send_counts[d];
  if (n_recv && v->recv.reserve((size_t)n_recv, false, st)) return 1;
  bool any = false;
  for (int r = 0; r < nranks; r++) any = any || (r != me && (recv_from[r] || send_counts[r]));
  if (any) {
    GXY_NCCL(g_nccl.GroupStart());
    int roff = 0;
    for (int r = 0; r < nranks; r++) {
      if (r == me) continue;
      if (send_counts[r])
        for (int c = 0; c < 24; c++)
          GXY_NCCL(g_nccl.Send(v->send.base + (size_t)c * v->send.cap + send_offsets[r], send_counts[r], ncclFloat32, r, v->ctx->comm, st));
      if (recv_from[r])
        for (int c = 0; c < 24; c++)
          GXY_NCCL(g_nccl.Recv(v->recv.base + (size_t)c * v->recv.cap + roff, recv_from[r], ncclFloat32, r, v->ctx->comm, st));
      roff += recv_from[r];
    }
    GXY_NCCL(g_nccl.GroupEnd());
  }
  const int total = n_recv + recv_from[me];
  if (total) {
    if (v->next.reserve((size_t)*n_next + total, true, st)) return 1;
    if (recv_from[me]) {
      if (launch_copy_rays(v->next.v, (size_t)*n_next, v->send.v, (size_t)send_offsets[me], recv_from[me], st)) return 1;
      *n_next += recv_from[me];
      S.kernel_launches += 1;
    }
    int roff = 0;
    for (int r = 0; r < nranks; r++) {
      if (r == me || !recv_from[r]) continue;
      if (launch_copy_rays(v->next.v, (size_t)*n_next, v->recv.v, (size_t)roff, recv_from[r], st)) return 1;
      *n_next += recv_from[r];
      roff += recv_from[r];
      S.kernel_launches += 1;
    }
  }
  *global_pending = pending;
  return 0;
}

int gxy_sample(int nparts, gxy_vis *const *parts, const gxy_camera *cam, int w, int h, gxy_stats *stats) {
  GXY_CHECK(nparts >= 1 && parts && cam && w > 0 && h > 0, "gxy_sample: bad arguments");
  for (int p = 0; p < nparts; p++) {
    if (check_vis(parts[p])) return 1;
    GXY_CHECK(!parts[p]->samplers.empty(), "gxy_sample: partition %d holds no sampler operator", p);
  }
  // one process per GPU (gxy_comm_init): this rank's partition is parts[0], rays that cross into a neighbour travel as NCCL
  // send/recv pairs of their 24 columns, every rank takes part in every wave until no ray is left anywhere (Sampler.cpp:52-133
  // runs on the Renderer's ray queues and message layer; src/sampler/Sampler.h)
  const bool multi_proc = parts[0]->ctx->comm != nullptr;
  GXY_CHECK(!multi_proc || nparts == 1, "gxy_sample: with a communicator every process passes its one partition");
  for (int p = 0; p < nparts; p++) GXY_CHECK((parts[p]->ctx->comm != nullptr) == multi_proc, "gxy_sample: partitions with and without a communicator");
  const int nd = multi_proc ? parts[0]->ctx->nranks : nparts;  // destinations of a forwarded ray
  const int me = multi_proc ? parts[0]->ctx->rank : 0;
  const DevCamera C = make_dev_camera(*cam, w, h);
  const int npix = w * h;
  gxy_stats S;
  memset(&S, 0, sizeof S);
  // all passes of a ray through a partition inside one launch (default; measured on one B200, IsoSampler on eightBalls 128^3, 1080p:
  // 10 waves x 5 kernels with a host synchronisation each 2.87 ms -> 1 wave 1.29 ms, 1.69 G ray passes/s; same sample sets and pass
  // counts, tests/test_sampler.py).  GXY_SAMPLER_LOOP=0: one launch per crossing
  const char *le = getenv("GXY_SAMPLER_LOOP");
  const bool loop_mode = !(le && atoi(le) == 0);
  gxy_context *ctx0 = parts[0]->ctx;
  struct EventPair {  // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  if (use_device(ctx0)) return 1;
  GXY_CUDA(cudaEventCreate(&ev.a));
  GXY_CUDA(cudaEventCreate(&ev.b));
  cudaEvent_t ev0 = ev.a, ev1 = ev.b;
  GXY_CUDA(cudaEventRecord(ev0, ctx0->stream));
  std::vector<int> n_cur(nparts, 0);
  for (int p = 0; p < nparts; p++) {
    gxy_vis *v = parts[p];
    if (use_device(v->ctx)) return 1;
    cudaStream_t st = v->ctx->stream;
    const int npix_pad = ((w + 15) / 16) * ((h + 7) / 8) * 128;
    if (v->block_sums.reserve((size_t)std::max(npix_pad, 1 << 20) / 1024 + 2) || v->small.reserve(64 + 4 * (size_t)nd)) return 1;
    if (!v->d_sample_count) GXY_CUDA(cudaMalloc(&v->d_sample_count, 2 * sizeof(unsigned long long)));  // [0] samples [1] passes (loop mode)
    GXY_CUDA(cudaMemsetAsync(v->d_sample_count, 0, 2 * sizeof(unsigned long long), st));
    v->n_samples = 0;
    if (v->cur.reserve(npix, false, st)) return 1;
    if (launch_generate(v->P, C, w, h, tiled_order(), v->cur.v, nullptr, v->block_sums.p, v->small.p, st)) return 1;
    S.kernel_launches += 3;
    GXY_CUDA(cudaMemcpyAsync(&n_cur[p], v->small.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GXY_CUDA(cudaStreamSynchronize(st));
    S.primary_rays += n_cur[p];
  }
  std::vector<std::vector<int>> send_counts(nparts, std::vector<int>(nd, 0)), send_offsets(nparts, std::vector<int>(nd + 1, 0));
  for (int wave = 0; wave < 1000000; wave++) {
    long long pending = 0;
    for (int p = 0; p < nparts; p++) pending += n_cur[p];
    if (pending == 0 && !multi_proc) break;  // (across processes the exchange below decides: a rank with nothing to trace may still receive)
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      std::fill(send_counts[p].begin(), send_counts[p].end(), 0);
      const int n = n_cur[p];
      if (n == 0) continue;
      if (use_device(v->ctx)) return 1;
      cudaStream_t st = v->ctx->stream;
      if (!loop_mode) {
        if (reserve_samples(v, v->n_samples + (unsigned long long)n, st)) return 1;  // at most one sample per ray and pass
        if (launch_sampler_trace(v->SP, v->cur.v, n, v->d_samples, v->d_sample_count, v->samples_cap, nullptr, false, st)) return 1;
      } else {
        // GXY_SAMPLER_LOOP=1: all passes of a ray inside one launch.  The number of samples is not bounded by n any more: start with
        // room for 4 per ray; if the kernel counted more than fit, grow to the exact count, restore t and run the launch again
        if (v->io_f.reserve((size_t)n)) return 1;
        GXY_CUDA(cudaMemcpyAsync(v->io_f.p, v->cur.v.t, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        const unsigned long long before = v->n_samples;
        if (reserve_samples(v, before + 4ull * (unsigned long long)n, st)) return 1;
        for (int attempt = 0; attempt < 2; attempt++) {
          GXY_CUDA(cudaMemsetAsync(v->d_sample_count + 1, 0, sizeof(unsigned long long), st));
          if (launch_sampler_trace(v->SP, v->cur.v, n, v->d_samples, v->d_sample_count, v->samples_cap, v->d_sample_count + 1, true, st)) return 1;
          unsigned long long cnt[2] = {0, 0};
          GXY_CUDA(cudaMemcpyAsync(cnt, v->d_sample_count, sizeof cnt, cudaMemcpyDeviceToHost, st));
          GXY_CUDA(cudaStreamSynchronize(st));
          if (cnt[0] <= v->samples_cap) {
            S.traced_rays += (long long)cnt[1] - n;  // the passes beyond the first (the first is counted below, as in the other mode)
            break;
          }
          GXY_CHECK(attempt == 0, "gxy_sample: sample buffer still too small after growing to the counted size");
          if (reserve_samples(v, cnt[0], st)) return 1;
          GXY_CUDA(cudaMemcpyAsync(v->d_sample_count, &before, sizeof before, cudaMemcpyHostToDevice, st));
          GXY_CUDA(cudaMemcpyAsync(v->cur.v.t, v->io_f.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
          GXY_CUDA(cudaStreamSynchronize(st));
        }
      }
      if (launch_classify(v->P, v->cur.v, n, st)) return 1;
      if (v->send.reserve((size_t)n, false, st)) return 1;
      int *d_counts = v->small.p + 8, *d_offsets = d_counts + nd, *d_cursor = d_offsets + nd + 1;
      if (launch_partition_by_destination(v->cur.v, n, nd, multi_proc ? me : p, v->send.v, d_counts, d_offsets, d_cursor, st)) return 1;
      GXY_CUDA(cudaMemcpyAsync(send_counts[p].data(), d_counts, sizeof(int) * nd, cudaMemcpyDeviceToHost, st));
      GXY_CUDA(cudaMemcpyAsync(send_offsets[p].data(), d_offsets, sizeof(int) * (nd + 1), cudaMemcpyDeviceToHost, st));
      GXY_CUDA(cudaMemcpyAsync(&v->n_samples, v->d_sample_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      S.kernel_launches += 5;
      S.traced_rays += n;
      S.waves++;
    }
    for (int p = 0; p < nparts; p++) {
      if (n_cur[p] == 0) continue;
      if (use_device(parts[p]->ctx)) return 1;
      GXY_CUDA(cudaStreamSynchronize(parts[p]->ctx->stream));
    }
    std::vector<int> n_next(nparts, 0);
    if (multi_proc) {
      long long global_pending = 0;
      if (nccl_exchange_rays(parts[0], nd, me, send_counts[0].data(), send_offsets[0].data(), &n_next[0], &global_pending, S)) return 1;
      GXY_CUDA(cudaStreamSynchronize(parts[0]->ctx->stream));
      std::swap(parts[0]->cur, parts[0]->next);
      n_cur[0] = n_next[0];
      if (global_pending == 0) break;
      continue;
    }
    for (int src = 0; src < nparts; src++)
      for (int dst = 0; dst < nparts; dst++) {
        const int cnt = send_counts[src][dst];
        if (!cnt) continue;
        gxy_vis *vs = parts[src], *vd = parts[dst];
        if (src != dst) S.forwarded_rays += cnt;
        if (use_device(vd->ctx)) return 1;
        cudaStream_t st = vd->ctx->stream;
        GXY_CHECK(vs->ctx->device == vd->ctx->device, "gxy_sample: all partitions of one call live on one device");
        if (vd->next.reserve((size_t)n_next[dst] + cnt, true, st)) return 1;
        if (launch_copy_rays(vd->next.v, (size_t)n_next[dst], vs->send.v, (size_t)send_offsets[src][dst], cnt, st)) return 1;
        S.kernel_launches += 1;
        n_next[dst] += cnt;
      }
    for (int p = 0; p < nparts; p++) {
      if (use_device(parts[p]->ctx)) return 1;
      GXY_CUDA(cudaStreamSynchronize(parts[p]->ctx->stream));
      std::swap(parts[p]->cur, parts[p]->next);
      n_cur[p] = n_next[p];
    }
  }
  if (use_device(ctx0)) return 1;
  GXY_CUDA(cudaEventRecord(ev1, ctx0->stream));
  GXY_CUDA(cudaEventSynchronize(ev1));
  GXY_CUDA(cudaEventElapsedTime(&S.device_ms, ev0, ev1));
  for (int p = 0; p < nparts; p++)
    if (check_error_flag(parts[p])) return 1;
  if (stats) *stats = S;
  return 0;
}

int gxy_vis_sample_count(gxy_vis *v, long long *n) {
  GXY_CHECK(v && n, "gxy_vis_sample_count: NULL argument");
  *n = (long long)v->n_samples;
  return 0;
}

int gxy_vis_download_samples(gxy_vis *v, float *xyz) {
  if (check_vis(v)) return 1;
  GXY_CHECK(xyz || v->n_samples == 0, "gxy_vis_download_samples: NULL buffer");
  if (v->n_samples) GXY_CUDA(cudaMemcpy(xyz, v->d_samples, sizeof(float) * 3 * v->n_samples, cudaMemcpyDeviceToHost));
  return 0;
}

int gxy_particles_from_samples(gxy_vis *v, gxy_particles **out) {
  if (check_vis(v)) return 1;
  GXY_CHECK(out, "gxy_particles_from_samples: NULL argument");
  gxy_particles *p = new gxy_particles();
  p->ctx = v->ctx; p->n = (int)v->n_samples;
  p->d_centers = nullptr; p->d_data = nullptr;
  const size_t n = (size_t)std::max<unsigned long long>(v->n_samples, 1);
  if (cudaMalloc(&p->d_centers, sizeof(float) * 3 * n) != cudaSuccess || cudaMalloc(&p->d_data, sizeof(float) * n) != cudaSuccess) {
    cudaFree(p->d_centers);
    delete p;
    gxy_set_error("gxy_particles_from_samples: out of device memory");
    return 1;
  }
  cudaError_t e = v->n_samples ? cudaMemcpy(p->d_centers, v->d_samples, sizeof(float) * 3 * v->n_samples, cudaMemcpyDeviceToDevice) : cudaSuccess;
  if (e == cudaSuccess) e = cudaMemset(p->d_data, 0, sizeof(float) * n);   // newsample.u.value = 0.0 (Sampler.cpp:83)
  if (e != cudaSuccess) {
    gxy_set_error("gxy_particles_from_samples: %s", cudaGetErrorString(e));
    gxy_particles_destroy(p);
    return 1;
  }
  *out = p;
  return 0;
}

int gxy_generate_rays(gxy_vis *v, const gxy_camera *cam, int w, int h, gxy_raylist_view rays, int *n_out) {
  if (check_vis(v)) return 1;
  GXY_CHECK(cam && rays.base && w > 0 && h > 0 && rays.aligned_n >= w * h, "gxy_generate_rays: bad arguments");
  LaneGuard G(v);
  if (G.prepare()) return 1;
  ListLane &l = *G.l;
  cudaStream_t st = l.st;
  const int npix = w * h;
  if (l.cur.reserve(npix, false, st) || l.block_sums.reserve((size_t)npix / 1024 + 2) || l.small.reserve(64)) return 1;
  const DevCamera C = make_dev_camera(*cam, w, h);
  if (launch_generate(G.params(), C, w, h, false, l.cur.v, nullptr, l.block_sums.p, l.small.p, st)) return 1;
  int n = 0;
  GXY_CUDA(cudaMemcpyAsync(&n, l.small.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaStreamSynchronize(st));
  if (d2h_rays(st, l.cur, rays.base, n, rays.aligned_n)) return 1;
  GXY_CUDA(cudaStreamSynchronize(st));
  *n_out = n;
  return 0;
}

int gxy_intersect(gxy_vis *v, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar, int *geom_prim2,
                  float *tuv3) {
  if (check_vis(v)) return 1;
  if (n <= 0) return 0;
  LaneGuard G(v);
  if (G.prepare()) return 1;
  ListLane &l = *G.l;
  cudaStream_t st = l.st;
  if (l.io_f.reserve((size_t)11 * n) || l.io_i.reserve((size_t)2 * n)) return 1;
  float *d_org = l.io_f.p, *d_dir = d_org + 3 * (size_t)n, *d_tn = d_dir + 3 * (size_t)n, *d_tf = d_tn + n, *d_tuv = d_tf + n;
  GXY_CUDA(cudaMemcpyAsync(d_org, org3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st));
  GXY_CUDA(cudaMemcpyAsync(d_dir, dir3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st));
  GXY_CUDA(cudaMemcpyAsync(d_tn, tnear, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  GXY_CUDA(cudaMemcpyAsync(d_tf, tfar, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  if (launch_intersect(G.params(), n, d_org, d_dir, d_tn, d_tf, l.io_i.p, d_tuv, st)) return 1;
  GXY_CUDA(cudaMemcpyAsync(geom_prim2, l.io_i.p, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaMemcpyAsync(tuv3, d_tuv, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaStreamSynchronize(st));
  return G.check_error();
}

// ------------------------------------------------------------------------------------------------
// Peer arenas (one process per GPU).  Collective over the communicator: every rank must call with the same
// sizes (they follow from w, h and the lighting, which are the same everywhere).  NCCL is only the bootstrap
// here: it carries the 64-byte IPC handles and the "did it work everywhere" vote.
static void peer_arena_unmap(PeerArena &A, int rank) {
  for (int r = 0; r < GXY_MAX_RANKS; r++)
    if (A.opened[r] && r != rank) { cudaIpcCloseMemHandle(A.opened[r]); A.opened[r] = nullptr; }
}

static int peer_allgather_bytes(gxy_context *c, const void *mine, void *all, size_t bytes_each, Scratch<unsigned char> &dev) {
  const size_t n = (size_t)c->nranks;
  if (dev.reserve(bytes_each * (n + 1))) return 1;
  GXY_CUDA(cudaMemcpyAsync(dev.p, mine, bytes_each, cudaMemcpyHostToDevice, c->stream));
  GXY_NCCL(g_nccl.AllGather(dev.p, dev.p + bytes_each, bytes_each, ncclChar, c->comm, c->stream));
  GXY_CUDA(cudaMemcpyAsync(all, dev.p + bytes_each, bytes_each * n, cudaMemcpyDeviceToHost, c->stream));
  GXY_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static int ensure_peer_arena(gxy_context *c, PeerArena &A, unsigned npix, unsigned inbox_cap) {
  if (A.disabled) return 0;
  if (A.base && A.T.npix >= npix && A.T.inbox_cap >= inbox_cap) return 0;
  GXY_CHECK(c->nranks <= GXY_MAX_RANKS, "peer arenas support up to %d ranks", GXY_MAX_RANKS);
  Scratch<unsigned char> dev;
  // 1. nobody may still have the old arenas mapped when they are freed
  peer_arena_unmap(A, c->rank);
  char token = 1, tokens[GXY_MAX_RANKS];
  if (peer_allgather_bytes(c, &token, tokens, 1, dev)) return 1;
  if (A.base) { cudaFree(A.base); A.base = nullptr; }
  // 2. allocate and clear the new arena
  PeerTable T;
  memset(&T, 0, sizeof T);
  T.rank = c->rank; T.nranks = c->nranks; T.inbox_cap = inbox_cap; T.npix = npix;
  auto align = [](unsigned long long x) { return (x + 255ull) & ~255ull; };
  T.off_proxy = sizeof(PeerCtrl);
  T.off_fb = align(T.off_proxy + sizeof(PartProxy));
  T.off_final = align(T.off_fb + (unsigned long long)npix * 16ull);
  T.off_inbox[0] = align(T.off_final + (unsigned long long)npix * 16ull);
  T.off_inbox[1] = align(T.off_inbox[0] + (unsigned long long)inbox_cap * 64ull);
  const size_t bytes = (size_t)align(T.off_inbox[1] + (unsigned long long)inbox_cap * 64ull);
  int ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (cudaMalloc(&A.base, bytes) != cudaSuccess) { cudaGetLastError(); A.base = nullptr; ok = 0; }
  if (ok && cudaMemsetAsync(A.base, 0, sizeof(PeerCtrl), c->stream) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, A.base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  // 3. exchange handles (the all-gather also orders every rank's memset before any peer access) and map
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; } msg, all[GXY_MAX_RANKS];
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memset(&msg, 0, sizeof msg);
  msg.h = mine; msg.ok = ok;
  if (peer_allgather_bytes(c, &msg, all, sizeof(Msg), dev)) return 1;
  for (int r = 0; r < c->nranks; r++) ok = ok && all[r].ok;
  if (ok) {
    for (int r = 0; r < c->nranks && ok; r++) {
      if (r == c->rank) { T.base[r] = A.base; continue; }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        fprintf(stderr, "[galaxy_b200 rank %d] cudaIpcOpenMemHandle(rank %d) failed: %s\n", c->rank, r, cudaGetErrorString(cudaGetLastError()));
        ok = 0;
        break;
      }
      A.opened[r] = ptr;
      T.base[r] = (char *)ptr;
    }
  }
  // 4. all or nothing
  int vote = ok, votes[GXY_MAX_RANKS];
  if (peer_allgather_bytes(c, &vote, votes, sizeof(int), dev)) return 1;
  for (int r = 0; r < c->nranks; r++) ok = ok && votes[r];
  dev.release();
  if (!ok) {
    peer_arena_unmap(A, c->rank);
    if (A.base) cudaFree(A.base);
    A.base = nullptr;
    A.disabled = true;
    if (c->rank == 0) fprintf(stderr, "[galaxy_b200] peer arenas unavailable (CUDA IPC); using the NCCL ray exchange\n");
    return 0;
  }
  A.T = T;
  A.bytes = bytes;
  A.epoch = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Frames in flight.  A Flight owns everything one frame of a geometry-only Visualization mutates, so a frame is
// SUBMITTED (all its launches enqueued on the flight's streams, no host round trip) and later WAITED for; in between the
// host submits other flights.  Two frame schedules are submitted this way:
//   single  one partition without neighbours (the single-GPU frame): the band pipeline
//   peer    one process per GPU: a fixed schedule of waves over the peer arenas (below)
// Everything else (volumes, several partitions in one process, the NCCL list exchange) renders synchronously inside the
// submit call and only hands its image over at the wait.
static int flight_get(gxy_vis *v, int slot, Flight **out) {
  GXY_CHECK(slot >= 0 && slot < GXY_MAX_FLIGHTS, "frame slot %d out of range (0..%d)", slot, GXY_MAX_FLIGHTS - 1);
  Flight *F = v->flights[slot];
  if (!F) {
    F = new Flight();
    v->flights[slot] = F;
    memset(&F->S, 0, sizeof F->S);
    GXY_CUDA(cudaStreamCreateWithFlags(&F->st, cudaStreamNonBlocking));
    GXY_CUDA(cudaEventCreate(&F->ev0));
    GXY_CUDA(cudaEventCreate(&F->ev1));
    GXY_CUDA(cudaEventCreateWithFlags(&F->ev_fork, cudaEventDisableTiming));
    GXY_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&F->h_tail), sizeof(Flight::Tail), cudaHostAllocPortable));
    memset(F->h_tail, 0, sizeof(Flight::Tail));
    if (F->err.reserve(4)) return 1;
    GXY_CUDA(cudaMemsetAsync(F->err.p, 0, sizeof(int) * 4, F->st));
  }
  *out = F;
  return 0;
}

static void flight_destroy(gxy_vis *v, Flight *F) {
  if (!F) return;
  if (F->st) cudaStreamSynchronize(F->st);
  for (int k = 0; k < 16; k++) {
    if (F->lanes[k]) { cudaStreamSynchronize(F->lanes[k]); cudaStreamDestroy(F->lanes[k]); }
    if (F->ev_join[k]) cudaEventDestroy(F->ev_join[k]);
  }
  for (auto &e : F->trace_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (F->gexec) cudaGraphExecDestroy(F->gexec);
  if (F->ev_fork) cudaEventDestroy(F->ev_fork);
  if (F->ev0) cudaEventDestroy(F->ev0);
  if (F->ev1) cudaEventDestroy(F->ev1);
  if (F->st) cudaStreamDestroy(F->st);
  F->hits.release(); F->next.release(); F->cur.release(); F->fq.release(); F->rawhits.release(); F->fb.release();
  F->vl[0].release(); F->vl[1].release(); F->v_hit_index.release(); F->v_block_sums.release(); F->vq.release();
  F->proxies.release(); F->err.release();
  peer_arena_unmap(F->arena, v->ctx->rank);
  if (F->arena.base) cudaFree(F->arena.base);
  if (F->h_tail) cudaFreeHost(F->h_tail);
  delete F;
}

static int flight_lane(Flight &F, int k, cudaStream_t *out) {
  if (!F.lanes[k]) {
    GXY_CUDA(cudaStreamCreateWithFlags(&F.lanes[k], cudaStreamNonBlocking));
    GXY_CUDA(cudaEventCreateWithFlags(&F.ev_join[k], cudaEventDisableTiming));
  }
  *out = F.lanes[k];
  return 0;
}
static int flight_trace_begin(Flight &F, cudaStream_t st) {
  if (F.graph_frame) return 0;  // (no timing events inside a captured frame: gxy_stats::trace_ms is 0 for it)
  if (F.n_trace_ev == F.trace_ev.size()) {
    cudaEvent_t ta, tb;
    GXY_CUDA(cudaEventCreate(&ta));
    GXY_CUDA(cudaEventCreate(&tb));
    F.trace_ev.push_back(std::make_pair(ta, tb));
  }
  GXY_CUDA(cudaEventRecord(F.trace_ev[F.n_trace_ev].first, st));
  return 0;
}
static int flight_trace_end(Flight &F, cudaStream_t st) {
  if (F.graph_frame) return 0;
  GXY_CUDA(cudaEventRecord(F.trace_ev[F.n_trace_ev].second, st));
  F.n_trace_ev++;
  return 0;
}
// the frame's counters and error flag travel to page-locked memory behind the end-of-frame event; the flag is cleared for
// the next frame of this flight (a reported error does not poison later frames)
static int flight_tail(gxy_vis *v, Flight &F, int n_queues) {
  cudaStream_t st = F.st;
  GXY_CUDA(cudaMemcpyAsync(F.h_tail->q, F.fq.p, (size_t)n_queues * sizeof(FusedQueues), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaMemcpyAsync(&F.h_tail->error, F.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaMemsetAsync(F.err.p, 0, sizeof(int), st));
  (void)v;
  return 0;
}

// ---- single partition without neighbours: the band pipeline (one GPU) --------------------------------
// The window is split into interleaved bands of tile rows, each band runs gen -> primary trace -> shade -> secondary
// trace with its own queues on its own stream.  A persistent trace kernel ends with a drain phase -- the last rays fetched
// still need their ~30 dependent node visits while most warps have nothing left (ncu: the primary kernel issues on 60 % of
// its active cycles but only 31 % of all cycles) -- and a frame has two of them; with bands the CTAs of another band's
// kernel move into the SMs a draining kernel vacates, and primary and secondary phases of different bands overlap.
// measured on C5 (tools/band_sweep.py; bands on as many streams): 1: 1.69 ms, 2: 1.61, 3: 1.53, 4: 1.43, 6: 1.44-1.51,
// 8: 1.43-1.53, 16: 1.81; fewer streams than bands is worse than no bands at all (4 bands on 2 streams: 1.99 ms)
static int flight_submit_single(gxy_vis *v, Flight &F, const DevCamera &C, int w, int h, float epsilon) {
  const int npix = w * h;
  const int n_sec_per_hit = F.n_sec_per_hit;
  // alone on the device a frame fills the drain phases of its trace kernels with its own bands (4); behind other frames in flight
  // their kernels do that, and one band per frame is faster (measured on C5, tools/flight_sweep.py: 4 in flight, 4 bands 1.215 ms,
  // 2 bands 1.160, 1 band 1.102 per frame; 1 in flight: 1.447 / 1.635 / 1.703)
  bool others = false;
  for (int k = 0; k < GXY_MAX_FLIGHTS; k++) others = others || (v->flights[k] && v->flights[k] != &F && v->flights[k]->pending);
  int n_bands = others ? 1 : 4;
  if (const char *e = getenv("GXY_BANDS")) n_bands = std::max(1, std::min(16, atoi(e)));
  int n_lanes = n_bands;
  if (const char *e = getenv("GXY_BAND_STREAMS")) n_lanes = std::max(1, std::min(16, atoi(e)));
  n_lanes = std::min(n_lanes, n_bands);
  const int tiles_x8 = (w + 7) / 8, tiles_y4 = (h + 3) / 4;
  const size_t qcap = (size_t)tiles_x8 * tiles_y4 * 32;  // queue slots of all bands together (>= npix)
  cudaStream_t st = F.st;
  SceneParams P = v->P;
  P.error_flag = F.err.p;
  if (F.hits.reserve(qcap, false, st) || F.fq.reserve((size_t)n_bands * sizeof(FusedQueues) / 8) || F.rawhits.reserve((size_t)6 * qcap) ||
      F.next.reserve(qcap, false, st) || F.cur.reserve(64, false, st) || F.fb.reserve((size_t)npix * 4))
    return 1;
  F.n_bands = n_bands;
  F.fb_result = F.fb.p;
  F.n_trace_ev = 0;
  gxy_stats &S = F.S;
  GXY_CUDA(cudaEventRecord(F.ev0, st));
  GXY_CUDA(cudaMemsetAsync(F.fb.p, 0, sizeof(float) * 4 * npix, st));
  GXY_CUDA(cudaMemsetAsync(F.fq.p, 0, (size_t)n_bands * sizeof(FusedQueues), st));
  auto rays_offset = [](Rays r, size_t off) {
    float **fcol = &r.ox;
    for (int k = 0; k < 20; k++) fcol[k] += off;
    int **icol = &r.x;
    for (int k = 0; k < 5; k++) icol[k] += off;
    return r;
  };
  if (flight_trace_begin(F, st)) return 1;
  GXY_CUDA(cudaEventRecord(F.ev_fork, st));
  for (int k = 1; k < n_lanes; k++) {
    cudaStream_t sk;
    if (flight_lane(F, k, &sk)) return 1;
    GXY_CUDA(cudaStreamWaitEvent(sk, F.ev_fork, 0));
  }
  size_t off = 0;
  for (int b = 0; b < n_bands; b++) {
    cudaStream_t sb = (b % n_lanes) ? F.lanes[b % n_lanes] : st;
    const int rows = (tiles_y4 - b + n_bands - 1) / n_bands;
    if (rows <= 0) continue;
    const size_t nq = (size_t)tiles_x8 * rows * 32;
    FusedQueues *qb = reinterpret_cast<FusedQueues *>(F.fq.p) + b;
    if (launch_fused_primary(P, C, F.L, w, h, F.fb.p, rays_offset(F.next.v, off), F.rawhits.p + off, (unsigned)qcap, rays_offset(F.hits.v, off),
                             F.cur.v, 0u, qb, epsilon, nullptr, nullptr, b, n_bands, sb))
      return 1;
    S.kernel_launches += 3;
    if (n_sec_per_hit > 0) {
      if (launch_fused_secondary(P, F.L, w, h, n_sec_per_hit, (long long)nq * n_sec_per_hit, F.fb.p, rays_offset(F.hits.v, off), F.cur.v, 0u, qb,
                                 epsilon, !v->has_dvr, nullptr, 0, sb))
        return 1;
      S.kernel_launches += 1;
    }
    off += nq;
  }
  for (int k = 1; k < n_lanes; k++) {
    GXY_CUDA(cudaEventRecord(F.ev_join[k], F.lanes[k]));
    GXY_CUDA(cudaStreamWaitEvent(st, F.ev_join[k], 0));
  }
  if (flight_trace_end(F, st)) return 1;
  S.waves += n_sec_per_hit > 0 ? 2 : 1;
  if (flight_tail(v, F, n_bands)) return 1;
  GXY_CUDA(cudaEventRecord(F.ev1, st));
  return 0;
}

// ---- one process per GPU: the frame on one rank ------------------------------------------------------
// A fixed schedule of kernel launches with no host round trip:
//   wave 0     generate -> trace primaries -> shade hits; rays that leave go to peer inboxes[0]
//   wave k>=1  AO/shadow rays of the hits of wave k-1 -> trace inbox[(k-1)&1] -> shade the new hits; leavers go to inboxes[k&1]
// each wave ends with the flag barrier.  The secondaries of a wave's hits run one wave later on purpose: the rays a rank
// forwards while tracing reach its neighbours one barrier earlier, so a back partition traces the forwarded primaries
// while the front partition is busy with its AO/shadow rays.  A ray crosses at most H = sum(grid_i - 1) partition faces
// and so does each of its secondaries: 2H + 1 waves after wave 0 suffice for a regular grid; the global count of
// outstanding work (rays in flight + hits without secondaries) the last barrier leaves behind is checked at the wait and
// further waves run until it is zero.  The barriers cost latency, not throughput: while a rank waits in frame f its SMs
// trace the frames submitted behind it.
static int peer_wave(gxy_vis *v, Flight &F, const SceneParams &P, int w, int h, bool spawn, bool overlap) {
  const PeerTable &T = F.arena.T;
  FusedQueues *q = reinterpret_cast<FusedQueues *>(F.fq.p);
  float *fb = reinterpret_cast<float *>(F.arena.base + T.off_fb);
  cudaStream_t st = F.st;
  const int npix = w * h;
  gxy_stats &S = F.S;
  const int k = ++F.peer_k;
  const int parity_in = (k - 1) & 1;
  // waves 1 and 2 carry the bulk (the secondaries of the local hits; what the neighbours forwarded and its secondaries): full grids.
  // Later waves only move the few rays that cross a second or third partition face: small grids, the CTAs are persistent anyway
  // and a launch of 1184 CTAs that find an empty queue is not free.
  int late = 2;
  if (const char *e = getenv("GXY_LATE_BLOCKS")) late = std::max(1, std::min(8, atoi(e)));
  const int bps = k <= 2 ? 0 : late;
  if (flight_trace_begin(F, st)) return 1;
  // AO/shadow rays of the hits the previous wave found (leavers go to inboxes[k&1]) on a second stream ...
  cudaStream_t s2 = st;
  if (spawn && overlap) {
    if (flight_lane(F, 1, &s2)) return 1;
    GXY_CUDA(cudaEventRecord(F.ev_fork, st));
    GXY_CUDA(cudaStreamWaitEvent(s2, F.ev_fork, 0));
  }
  if (spawn && launch_fused_secondary(P, F.L, w, h, F.n_sec_per_hit, (long long)npix * F.n_sec_per_hit, fb, F.hits.v, F.cur.v, 0u, q, F.epsilon,
                                      !v->has_dvr, &T, k & 1, s2, bps))
    return 1;
  gxy_timeline_mark("secondary", s2);
  // ... while this stream traces the rays the neighbours sent during the previous wave (the two only share atomic counters)
  if (launch_inbox_wave(P, F.L, T, parity_in, w, h, fb, F.rawhits.p, (unsigned)npix, F.hits.v, q, F.epsilon, !v->has_dvr, st, bps)) return 1;
  if (s2 != st) {
    GXY_CUDA(cudaEventRecord(F.ev_join[1], s2));
    GXY_CUDA(cudaStreamWaitEvent(st, F.ev_join[1], 0));
  }
  if (flight_trace_end(F, st)) return 1;
  gxy_timeline_mark("inbox+shade", st);
  if (launch_wave_epilogue(T, q, ++F.arena.epoch, parity_in, spawn, F.err.p, st)) return 1;
  gxy_timeline_mark("barrier", st);
  S.kernel_launches += 3 + (spawn ? 1 : 0);
  S.waves++;
  return 0;
}
static int peer_finish(gxy_vis *v, Flight &F) {
  const PeerTable &T = F.arena.T;
  FusedQueues *q = reinterpret_cast<FusedQueues *>(F.fq.p);
  cudaStream_t st = F.st;
  // the queue counters as the last wave's barrier left them (global_pending is the termination test of the wait)
  if (flight_tail(v, F, 1)) return 1;
  // framebuffer: every rank sums its slice of all partial images into the owner's final image
  if (launch_fb_gather(T, st)) return 1;
  gxy_timeline_mark("fb_gather", st);
  if (launch_wave_epilogue(T, q, ++F.arena.epoch, -1, false, F.err.p, st)) return 1;
  gxy_timeline_mark("barrier_end", st);
  F.S.kernel_launches += 2;
  return 0;
}

// The rectangle of 8x4-pixel tiles that the box [lo - margin, hi + margin] projects into (x0, y0, nx, ny), for the generation
// kernel of one rank: a pixel whose ray touches the box lies inside the bounding rectangle of the projected corners (the box is
// convex, the projection maps lines to lines) as long as every corner is in front of the eye; 2 pixels are added for the rounding of
// the per-pixel ray arithmetic.  Returns false (whole image) when a corner is at or behind the eye plane.
static bool peer_tile_rect(const DevCamera &C, const float lo[3], const float hi[3], int w, int h, int rect[4]) {
  const double vr[3] = {C.vr.x, C.vr.y, C.vr.z}, vu[3] = {C.vu.x, C.vu.y, C.vu.z}, vd[3] = {C.vdir.x, C.vdir.y, C.vdir.z};
  const double eye[3] = {C.veye.x, C.veye.y, C.veye.z}, cen[3] = {C.center.x, C.center.y, C.center.z};
  auto dot = [](const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  const double rr = dot(vr, vr), uu = dot(vu, vu), ru = dot(vr, vu);
  const double det = rr * uu - ru * ru;
  if (!(det > 1e-20) || !(C.scaling > 0.f)) return false;
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  double c[3] = {cen[0] - eye[0], cen[1] - eye[1], cen[2] - eye[2]};  // eye -> image plane (perspective)
  const double cc = dot(c, c);
  for (int k = 0; k < 8; k++) {
    double m[3], p[3], rel[3];
    for (int a = 0; a < 3; a++) {
      m[a] = 1e-3 * ((double)hi[a] - (double)lo[a]) + 1e-4;
      p[a] = (k >> a) & 1 ? (double)hi[a] + 2.0 * m[a] : (double)lo[a] - 2.0 * m[a];
    }
    if (C.ortho) {
      for (int a = 0; a < 3; a++) rel[a] = p[a] - cen[a];
    } else {
      const double q[3] = {p[0] - eye[0], p[1] - eye[1], p[2] - eye[2]};
      const double qc = dot(q, c);
      if (!(cc > 0.0) || !(qc > 1e-6 * cc)) return false;  // at or behind the eye plane
      const double sc = cc / qc;
      for (int a = 0; a < 3; a++) rel[a] = eye[a] + sc * q[a] - cen[a];
    }
    const double br = dot(rel, vr), bu = dot(rel, vu);
    const double fx = (br * uu - bu * ru) / det, fy = (bu * rr - br * ru) / det;
    const double x = fx / (double)C.scaling + (double)C.off_x, y = fy / (double)C.scaling + (double)C.off_y;
    xmin = std::min(xmin, x); xmax = std::max(xmax, x); ymin = std::min(ymin, y); ymax = std::max(ymax, y);
  }
  if (!(xmin <= xmax) || !(ymin <= ymax)) return false;
  const int x0 = (int)std::max(0.0, std::floor(xmin) - 2.0), x1 = (int)std::min((double)(w - 1), std::ceil(xmax) + 2.0);
  const int y0 = (int)std::max(0.0, std::floor(ymin) - 2.0), y1 = (int)std::min((double)(h - 1), std::ceil(ymax) + 2.0);
  if (x1 < x0 || y1 < y0) { rect[0] = rect[1] = rect[2] = rect[3] = 0; return true; }  // the box is off screen: nothing to generate
  rect[0] = x0 / 8; rect[1] = y0 / 4;
  rect[2] = x1 / 8 - rect[0] + 1; rect[3] = y1 / 4 - rect[1] + 1;
  return true;
}

static int flight_enqueue_peer(gxy_vis *v, Flight &F, const DevCamera &C, int w, int h, float epsilon) {
  gxy_context *c = v->ctx;
  PeerArena &A = F.arena;
  const PeerTable &T = A.T;
  cudaStream_t st = F.st;
  const int npix = w * h;
  gxy_stats &S = F.S;
  SceneParams P = v->P;
  P.error_flag = F.err.p;
  float *fb = reinterpret_cast<float *>(A.base + T.off_fb);
  FusedQueues *q = reinterpret_cast<FusedQueues *>(F.fq.p);
  F.n_bands = 1;
  F.fb_result = c->rank == 0 ? reinterpret_cast<float *>(A.base + T.off_final) : fb;
  F.n_trace_ev = 0;
  F.peer_k = 0;
  gxy_timeline_mark("start", st);
  GXY_CUDA(cudaMemsetAsync(fb, 0, sizeof(float) * 4 * npix, st));
  GXY_CUDA(cudaMemsetAsync(q, 0, sizeof(FusedQueues), st));
  gxy_timeline_mark("memset", st);
  // ---- every rank publishes its partition proxy (box, neighbours, top of the BVH) and collects everybody's
  const bool spawn = F.n_sec_per_hit > 0;
  const bool peer_overlap = !(getenv("GXY_PEER_OVERLAP") && atoi(getenv("GXY_PEER_OVERLAP")) == 0);
  PartProxy *proxies = reinterpret_cast<PartProxy *>(F.proxies.p);
  if (launch_proxy_publish(P, T, st)) return 1;
  if (launch_wave_epilogue(T, q, ++A.epoch, -1, false, F.err.p, st)) return 1;
  if (launch_proxy_gather(T, proxies, st)) return 1;
  gxy_timeline_mark("proxies", st);
  S.kernel_launches += 3;
  // ---- wave 0: generation, trace of the primaries, shading of the hits (their AO/shadow rays follow in wave 1)
  if (flight_trace_begin(F, st)) return 1;
  int rect[4];
  const bool have_rect = peer_tile_rect(C, v->lmin, v->lmax, w, h, rect) && !(getenv("GXY_GEN_RECT") && atoi(getenv("GXY_GEN_RECT")) == 0);
  if (launch_fused_primary(P, C, F.L, w, h, fb, F.next.v, F.rawhits.p, (unsigned)npix, F.hits.v, F.cur.v, 0u, q, epsilon, &T, proxies, 0, 1, st,
                           have_rect ? rect : nullptr))
    return 1;
  if (flight_trace_end(F, st)) return 1;
  S.kernel_launches += 3;
  if (launch_wave_epilogue(T, q, ++A.epoch, -1, spawn, F.err.p, st)) return 1;
  gxy_timeline_mark("barrier0", st);
  S.kernel_launches += 1;
  S.waves = 1;
  // ---- waves 1..bound
  int f[3];
  gxy_factor(c->nranks, f);
  const int bound = 2 * ((f[0] - 1) + (f[1] - 1) + (f[2] - 1)) + (spawn ? 1 : 0);
  for (int b = 0; b < bound; b++)
    if (peer_wave(v, F, P, w, h, spawn, peer_overlap)) return 1;
  return peer_finish(v, F);
}

// ---- one process per GPU, Visualization with volumes: the frame on one rank ----------------------------
// The list kernels of the synchronous loop (render_sync) on two device lists whose lengths stay on the device, one wave = trace ->
// hit scan -> AO/shadow spawn into the next list -> classify -> framebuffer -> forward (records into the neighbours' inboxes) ->
// flag barrier -> the own inbox is appended to the next list.  A ray and each of its secondaries cross at most H partition faces:
// 2H + 2 waves, and whatever the last barrier still reports is traced at the wait.  No host round trip, so frames overlap across
// ranks: the brick behind marches frame f while the brick in front marches frame f + 1 -- the only way a DVR frame whose bricks
// depend on each other front to back can use more than one GPU at a time.
static int vol_wave(gxy_vis *v, Flight &F, const SceneParams &P, int w, int h) {
  const PeerTable &T = F.arena.T;
  VolQueues *q = reinterpret_cast<VolQueues *>(F.vq.p);
  float *fb = reinterpret_cast<float *>(F.arena.base + T.off_fb);
  cudaStream_t st = F.st;
  const int k = F.v_k++;
  const int cur = k & 1, cap = F.v_cap;
  const int n_ao = F.lights.n_ao, n_sh = F.lights.shadows ? F.lights.n_lights : 0;
  Rays L0 = F.vl[cur].v, L1 = F.vl[cur ^ 1].v;
  if (flight_trace_begin(F, st)) return 1;
  if (launch_trace(P, L0, cap, F.epsilon, nullptr, !v->has_dvr, &q->samples, st, &q->cnt[cur])) return 1;
  if (flight_trace_end(F, st)) return 1;
  if (launch_hit_scan(L0, cap, F.v_hit_index.p, F.v_block_sums.p, &q->nhit, st, &q->cnt[cur])) return 1;
  if (launch_vol_wave_counts(q, cur, n_ao, n_sh, k == 0, st)) return 1;
  if (launch_shade_spawn(F.L, L0, cap, F.v_hit_index.p, &q->nhit, L1, F.epsilon, st, w * h)) return 1;
  if (launch_classify(P, L0, cap, st, &q->cnt[cur])) return 1;
  if (launch_accumulate(L0, cap, fb, w, h, &q->terminated, st, &q->cnt[cur])) return 1;
  if (launch_vol_forward(L0, cap, L1, cap, q, cur, T, k & 1, F.err.p, st)) return 1;
  if (launch_vol_epilogue(T, q, cur, ++F.arena.epoch, F.err.p, st)) return 1;
  if (launch_vol_unpack(T, k & 1, L1, cap, q, cur, F.err.p, st)) return 1;
  F.S.kernel_launches += 13 + (n_ao > 0 ? 1 : 0);
  F.S.waves++;
  return 0;
}
static int vol_finish(gxy_vis *v, Flight &F) {
  const PeerTable &T = F.arena.T;
  cudaStream_t st = F.st;
  VolQueues *q = reinterpret_cast<VolQueues *>(F.vq.p);
  GXY_CUDA(cudaMemcpyAsync(F.h_tail->q, q, sizeof(VolQueues), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaMemcpyAsync(&F.h_tail->error, F.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GXY_CUDA(cudaMemsetAsync(F.err.p, 0, sizeof(int), st));
  if (launch_fb_gather(T, st)) return 1;
  // (the end barrier only needs an all-rank rendezvous: reuse the wave barrier with an empty next list's count)
  if (launch_vol_epilogue(T, q, F.v_k & 1, ++F.arena.epoch, F.err.p, st)) return 1;
  F.S.kernel_launches += 2;
  (void)v;
  return 0;
}
static int flight_enqueue_volume(gxy_vis *v, Flight &F, const DevCamera &C, int w, int h, float epsilon) {
  gxy_context *c = v->ctx;
  PeerArena &A = F.arena;
  const PeerTable &T = A.T;
  cudaStream_t st = F.st;
  const int npix = w * h;
  SceneParams P = v->P;
  P.error_flag = F.err.p;
  float *fb = reinterpret_cast<float *>(A.base + T.off_fb);
  VolQueues *q = reinterpret_cast<VolQueues *>(F.vq.p);
  F.n_bands = 1;
  F.fb_result = c->rank == 0 ? reinterpret_cast<float *>(A.base + T.off_final) : fb;
  F.n_trace_ev = 0;
  F.v_k = 0;
  (void)epsilon;
  GXY_CUDA(cudaMemsetAsync(fb, 0, sizeof(float) * 4 * npix, st));
  GXY_CUDA(cudaMemsetAsync(q, 0, sizeof(VolQueues), st));
  // a rendezvous before anybody writes into anybody's inbox for this frame: the previous frame on this slot is over everywhere
  if (launch_vol_epilogue(T, q, 0, ++A.epoch, F.err.p, st)) return 1;
  // Camera::generate_initial_rays: the rays whose first brick is this one, in 16x8-pixel tile order; the count stays in q->cnt[0]
  if (launch_generate(P, C, w, h, tiled_order(), F.vl[0].v, nullptr, F.v_block_sums.p, &q->cnt[0], st)) return 1;
  F.S.kernel_launches += 4;
  int f[3];
  gxy_factor(c->nranks, f);
  const int H = (f[0] - 1) + (f[1] - 1) + (f[2] - 1);
  const bool spawn = F.n_sec_per_hit > 0;
  const int waves = (spawn ? 2 : 1) * H + (spawn ? 2 : 1);
  for (int k = 0; k < waves; k++)
    if (vol_wave(v, F, P, w, h)) return 1;
  return vol_finish(v, F);
}

// Submission of a peer frame.  Direct: ~75 runtime calls (37 launches on two streams, their fork/join events, memsets, the tail
// copies).  As a graph: the same calls are CAPTURED into a CUDA graph, the flight's executable graph is updated in place
// (cudaGraphExecUpdate: same topology, new kernel parameters -- camera, tile rectangle, barrier epochs) and launched with one call;
// the host then spends its time on the next frame's capture instead of on driver submissions (default; GXY_GRAPH=0 submits
// directly, as does GXY_PROFILE=1, whose timeline events cannot live in a capture).  The first peer frame of a process
// always goes direct (constant tables are uploaded and streams created on first use, which a capture must not contain).
static bool g_peer_warm = false;
static int flight_submit_peer(gxy_vis *v, Flight &F, const DevCamera &C, int w, int h, float epsilon) {
  gxy_context *c = v->ctx;
  cudaStream_t st = F.st;
  const int npix = w * h;
  if (F.volume) {
    const size_t cap = (size_t)F.v_cap;
    if (F.vl[0].reserve(cap, false, st) || F.vl[1].reserve(cap, false, st) || F.v_hit_index.reserve(2 * cap) ||
        F.v_block_sums.reserve(cap / 1024 + 2) || F.vq.reserve((sizeof(VolQueues) + 7) / 8))
      return 1;
  } else if (F.hits.reserve(npix, false, st) || F.fq.reserve(sizeof(FusedQueues) / 8) || F.rawhits.reserve((size_t)6 * npix) ||
             F.next.reserve(npix, false, st) || F.cur.reserve(64, false, st) || F.proxies.reserve(sizeof(PartProxy) * (size_t)c->nranks))
    return 1;
  cudaStream_t s2;
  if (flight_lane(F, 1, &s2)) return 1;
  // measured on 8 GPUs (tools/flight_sweep.py, 8 frames in flight): host time per submitted frame 0.271 ms direct, 0.067 ms as a graph;
  // frame 0.482 -> 0.389 ms (the direct submission was host bound); on 2 GPUs 0.156 -> 0.045 ms of host time, frame unchanged
  static bool g_vol_warm = false;
  const bool graph = (F.volume ? g_vol_warm : g_peer_warm) && !timeline_on() && !(getenv("GXY_GRAPH") && atoi(getenv("GXY_GRAPH")) == 0);
  F.graph_frame = graph;
  GXY_CUDA(cudaEventRecord(F.ev0, st));
  if (!graph) {
    if (F.volume ? flight_enqueue_volume(v, F, C, w, h, epsilon) : flight_enqueue_peer(v, F, C, w, h, epsilon)) return 1;
    (F.volume ? g_vol_warm : g_peer_warm) = true;
  } else {
    GXY_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = F.volume ? flight_enqueue_volume(v, F, C, w, h, epsilon) : flight_enqueue_peer(v, F, C, w, h, epsilon);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc || ce != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      if (!rc) gxy_set_error("stream capture of the frame failed: %s", cudaGetErrorString(ce));
      cudaGetLastError();
      return 1;
    }
    bool ok = false;
    if (F.gexec) {
      cudaGraphExecUpdateResultInfo info;
      ok = cudaGraphExecUpdate(F.gexec, g, &info) == cudaSuccess;
      if (!ok) { cudaGetLastError(); cudaGraphExecDestroy(F.gexec); F.gexec = nullptr; }
    }
    if (!ok) {
      const cudaError_t ie = cudaGraphInstantiate(&F.gexec, g, 0);
      if (ie != cudaSuccess) { cudaGraphDestroy(g); F.gexec = nullptr; gxy_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); return 1; }
    }
    cudaGraphDestroy(g);
    GXY_CUDA(cudaGraphLaunch(F.gexec, st));
  }
  GXY_CUDA(cudaEventRecord(F.ev1, st));
  return 0;
}

static VolQueues tail_vq(const Flight &F) {  // the end-of-frame counters of a volume frame (copied into the page-locked tail by vol_finish)
  VolQueues q;
  memcpy(&q, F.h_tail->q, sizeof q);
  return q;
}

static int flight_wait(gxy_vis *v, Flight &F, gxy_stats *stats) {
  gxy_context *c = v->ctx;
  GXY_CHECK(F.pending, "no frame was submitted to this slot");
  F.pending = false;
  if (!F.sync_done) {
    GXY_CUDA(cudaEventSynchronize(F.ev1));
    int err = F.h_tail->error;
    if (F.peer && F.volume && err == 0) {
      SceneParams P = v->P;
      P.error_flag = F.err.p;
      while (tail_vq(F).global_pending != 0u && F.h_tail->error == 0) {
        GXY_CHECK(F.v_k < 4096, "volume wave loop does not terminate (%u units of work in flight)", tail_vq(F).global_pending);
        F.graph_frame = false;
        if (vol_wave(v, F, P, F.w, F.h) || vol_finish(v, F)) return 1;
        GXY_CUDA(cudaEventRecord(F.ev1, F.st));
        GXY_CUDA(cudaEventSynchronize(F.ev1));
      }
      err = F.h_tail->error;
    } else if (F.peer && err == 0) {
      // anything still in flight after the scheduled waves (cannot happen on a regular grid): one wave at a time, every rank alike
      SceneParams P = v->P;
      P.error_flag = F.err.p;
      const bool peer_overlap = !(getenv("GXY_PEER_OVERLAP") && atoi(getenv("GXY_PEER_OVERLAP")) == 0);
      while (F.h_tail->q[0].global_pending != 0u && F.h_tail->error == 0) {
        GXY_CHECK(F.peer_k < 4096, "peer wave loop does not terminate (%u units of work in flight)", F.h_tail->q[0].global_pending);
        F.graph_frame = false;
        if (peer_wave(v, F, P, F.w, F.h, F.n_sec_per_hit > 0, peer_overlap) || peer_finish(v, F)) return 1;
        GXY_CUDA(cudaEventRecord(F.ev1, F.st));
        GXY_CUDA(cudaEventSynchronize(F.ev1));
      }
      err = F.h_tail->error;
    }
    GXY_CHECK(err == 0, "device error flag %d (1: BVH traversal stack overflow, 3: ray list / inbox capacity, 4: peer barrier timeout, 6: TMA copy did not complete, 7: translucent surface on the fused path)", err);
    gxy_stats &S = F.S;
    FusedQueues t = F.h_tail->q[0];
    if (F.volume) memset(&t, 0, sizeof t);
    for (int b = 1; b < F.n_bands; b++) {
      const FusedQueues &o = F.h_tail->q[b];
      t.n_generated += o.n_generated; t.n_hits += o.n_hits; t.n_terminated += o.n_terminated; t.n_primary32 += o.n_primary32;
      t.nodes += o.nodes; t.prims += o.prims; t.n_spill += o.n_spill;
    }
    const int nsec = F.n_sec_per_hit;
    S.primary_rays = (long long)t.n_generated;
    S.ao_rays = (long long)t.n_hits * F.lights.n_ao;
    S.shadow_rays = (long long)t.n_hits * (F.lights.shadows ? F.lights.n_lights : 0);
    S.terminated_rays = (long long)t.n_terminated;
    S.nodes_visited = (long long)t.nodes;
    S.prims_tested = (long long)t.prims;
    if (F.volume) {
      const VolQueues vq = tail_vq(F);
      S.primary_rays = (long long)vq.generated; S.ao_rays = (long long)vq.ao; S.shadow_rays = (long long)vq.shadow;
      S.forwarded_rays = (long long)vq.forwarded; S.terminated_rays = (long long)vq.terminated;
      S.traced_rays = S.dequeued_rays = (long long)vq.traced;
      S.volume_samples = (long long)vq.samples;
    } else if (F.peer) {
      S.forwarded_rays = (long long)t.n_spill + (long long)t.n_virtual;
      S.traced_rays = (long long)t.n_generated + (long long)t.n_hits * nsec + (long long)t.n_inbox;
      S.dequeued_rays = (long long)t.n_primary32 + (long long)t.n_hits * nsec + (long long)t.n_inbox;
    } else {
      S.traced_rays = (long long)t.n_generated + (long long)t.n_hits * nsec;
      S.dequeued_rays = (long long)t.n_primary32 + (long long)t.n_hits * nsec;
    }
    cudaEventElapsedTime(&S.device_ms, F.ev0, F.ev1);
    cudaEventElapsedTime(&S.t_begin_ms, c->ev_ref, F.ev0);
    cudaEventElapsedTime(&S.t_end_ms, c->ev_ref, F.ev1);
    S.trace_ms = 0.f;
    for (size_t k = 0; k < F.n_trace_ev; k++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, F.trace_ev[k].first, F.trace_ev[k].second);
      S.trace_ms += ms;
    }
#ifdef GXY_TRAV_COUNTERS
    {
      unsigned long long trav[2] = {0ull, 0ull};
      GXY_CUDA(cudaMemcpy(trav, v->P.trav_counters, sizeof trav, cudaMemcpyDeviceToHost));
      GXY_CUDA(cudaMemset(v->P.trav_counters, 0, sizeof trav));
      S.nodes_visited += (long long)trav[0];
      S.prims_tested += (long long)trav[1];
    }
#endif
    if (F.peer) timeline_print(c->rank);
  }
  v->fb_w = F.w; v->fb_h = F.h;
  v->fb_result = F.fb_result;
  if (stats) *stats = F.S;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Frame-level wave loop.
//
// A wave on one partition: trace the current list (mixed types) -> scan surface hits -> spawn AO /
// shadow rays into the next list + light the primaries -> classify -> add TERMINATED rays to this
// partition's partial framebuffer -> counting-sort the BOUNDARY rays by destination rank.  Between
// waves the sorted segments are handed to their destinations (device copy in-process, NCCL
// send/recv across processes) and appended to the destination's next list.  The loop ends when no
// partition has rays left (global sum), which replaces the reference's busy/idle tree +
// MPI_Allreduce termination protocol (RenderingSet.cpp:289-589).
static int render_sync(int nparts, gxy_vis *const *parts, const gxy_camera *cam, const gxy_lighting *lights_in, int w, int h, float epsilon,
                       gxy_stats *stats) {
  GXY_CHECK(nparts >= 1 && parts && cam && lights_in && w > 0 && h > 0, "gxy_render: bad arguments");
  for (int p = 0; p < nparts; p++)
    if (check_vis(parts[p])) return 1;
  gxy_context *ctx0 = parts[0]->ctx;
  const bool multi_proc = ctx0->comm != nullptr;
  GXY_CHECK(!multi_proc || nparts == 1, "with a communicator attached each process drives exactly one partition");
  const int nranks = multi_proc ? ctx0->nranks : nparts;
  const int rank0 = multi_proc ? ctx0->rank : 0;
  gxy_lighting lights;
  if (gxy_resolve_lights(lights_in, cam, &lights)) return 1;
  GXY_CHECK(lights.n_lights >= 1, "lighting needs at least one light");
  const DevLights L = make_dev_lights(lights);
  const DevCamera C = make_dev_camera(*cam, w, h);
  const int npix = w * h;
  const int n_sec_per_hit = lights.n_ao + (lights.shadows ? lights.n_lights : 0);
  gxy_stats S;
  memset(&S, 0, sizeof S);
  PhaseTimer PT;

  struct FrameEvents {  // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace;
    ~FrameEvents()
    {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
      for (auto &e : trace) {
        if (e.first) cudaEventDestroy(e.first);
        if (e.second) cudaEventDestroy(e.second);
      }
    }
    int new_pair(cudaEvent_t *ta, cudaEvent_t *tb)  // owned from creation on
    {
      trace.push_back(std::make_pair((cudaEvent_t) nullptr, (cudaEvent_t) nullptr));
      GXY_CUDA(cudaEventCreate(&trace.back().first));
      GXY_CUDA(cudaEventCreate(&trace.back().second));
      *ta = trace.back().first;
      *tb = trace.back().second;
      return 0;
    }
  } FE;
  if (use_device(ctx0)) return 1;
  GXY_CUDA(cudaEventCreate(&FE.a));
  GXY_CUDA(cudaEventCreate(&FE.b));
  cudaEvent_t &ev0 = FE.a, &ev1 = FE.b;
  GXY_CUDA(cudaEventRecord(ev0, ctx0->stream));

  std::vector<int> n_cur(nparts, 0);
  // Geometry-only Visualizations take the fused path for the rays born in this frame (gxy_fused.cu);
  // only rays that cross into a neighbour partition come back as lists ("spill") for the wave loop.
  bool fused = true;
  // (decided from the operator list, which is the same on every rank, not from the clipped primitive count)
  for (int p = 0; p < nparts; p++) fused = fused && parts[p]->P.n_volvis == 0 && !parts[p]->geoms.empty();
  if (!(getenv("GXY_FUSED_CURVES") && atoi(getenv("GXY_FUSED_CURVES")) != 0))  // PathLines on the list-path kernels (frame_kind)
    for (int p = 0; p < nparts; p++)
      for (const GeomOp &g : parts[p]->geoms) fused = fused && g.kind != 2;
  if (const char *e = getenv("GXY_FUSED")) fused = fused && atoi(e) != 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &trace_events = FE.trace;
  for (int p = 0; p < nparts; p++) {
    gxy_vis *v = parts[p];
    if (use_device(v->ctx)) return 1;
    cudaStream_t st = v->ctx->stream;
    const int npix_pad = ((w + 15) / 16) * ((h + 7) / 8) * 128;  // slots of the tiled generation order
    if (v->block_sums.reserve((size_t)std::max(npix_pad, 1 << 20) / 1024 + 2) || v->small.reserve(64 + 4 * (size_t)nranks) || v->counters.reserve(4) ||
        v->fb.reserve((size_t)npix * 4))
      return 1;
    v->fb_w = w; v->fb_h = h;
    v->fb_result = v->fb.p;
    GXY_CUDA(cudaMemsetAsync(v->fb.p, 0, sizeof(float) * 4 * npix, st));
    GXY_CUDA(cudaMemsetAsync(v->counters.p, 0, sizeof(unsigned long long) * 4, st));
  }
  if (fused) {
    static_assert(sizeof(FusedQueues) == 96, "FusedQueues layout");
    std::vector<FusedQueues> fq(nparts);
    std::vector<bool> can_spill(nparts, false);
    // (a single partition without neighbours never gets here: it is a Flight, flight_submit_single)
    const int n_bands = 1;
    const int tiles_x8 = (w + 7) / 8, tiles_y4 = (h + 3) / 4;
    const size_t qcap = (size_t)tiles_x8 * tiles_y4 * 32;  // queue slots of all bands together (>= npix)
    // ---- primary rays: generate -> trace -> light -> framebuffer, hit records for the secondaries
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      cudaStream_t st = v->ctx->stream;
      for (int f = 0; f < 6; f++) can_spill[p] = can_spill[p] || v->neighbors[f] >= 0;
      if (v->hits.reserve(qcap, false, st) || v->fq.reserve((size_t)n_bands * sizeof(FusedQueues) / 8) || v->rawhits.reserve((size_t)6 * qcap)) return 1;
      if (v->next.reserve(qcap, false, st)) return 1;  // the generated primaries
      if (v->cur.reserve(can_spill[p] ? (size_t)npix : 64, false, st)) return 1;
      GXY_CUDA(cudaMemsetAsync(v->fq.p, 0, (size_t)n_bands * sizeof(FusedQueues), st));
      cudaEvent_t ta, tb;
      if (FE.new_pair(&ta, &tb)) return 1;
      GXY_CUDA(cudaEventRecord(ta, st));
      {
        if (launch_fused_primary(v->P, C, L, w, h, v->fb.p, v->next.v, v->rawhits.p, (unsigned)qcap, v->hits.v, v->cur.v,
                                 can_spill[p] ? (unsigned)v->cur.cap : 0u, reinterpret_cast<FusedQueues *>(v->fq.p), epsilon, nullptr, nullptr, 0, 1,
                                 st))
          return 1;
        S.kernel_launches += 3;
        S.waves++;
      }
      GXY_CUDA(cudaEventRecord(tb, st));
    }
    PT.mark("launchP");
    // ---- secondary rays.  A partition without neighbours cannot spill: no host round trip at all.
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (n_sec_per_hit == 0) continue;
      if (use_device(v->ctx)) return 1;
      cudaStream_t st = v->ctx->stream;
      long long max_rays = (long long)npix * n_sec_per_hit;
      if (can_spill[p]) {  // size the spill list for the worst case: every secondary ray leaves
        GXY_CUDA(cudaMemcpyAsync(&fq[p], v->fq.p, sizeof(FusedQueues), cudaMemcpyDeviceToHost, st));
        GXY_CUDA(cudaStreamSynchronize(st));
        max_rays = (long long)fq[p].n_hits * n_sec_per_hit;
        GXY_CHECK(max_rays + fq[p].n_spill < (1ll << 31) - (1ll << 24), "secondary ray list too large (%lld rays)", max_rays);
        if (v->cur.reserve((size_t)(max_rays + fq[p].n_spill), true, st)) return 1;
      }
      cudaEvent_t ta, tb;
      if (FE.new_pair(&ta, &tb)) return 1;
      GXY_CUDA(cudaEventRecord(ta, st));
      if (launch_fused_secondary(v->P, L, w, h, n_sec_per_hit, max_rays, v->fb.p, v->hits.v, v->cur.v, can_spill[p] ? (unsigned)v->cur.cap : 0u,
                                 reinterpret_cast<FusedQueues *>(v->fq.p), epsilon, !v->has_dvr, nullptr, 0, st))
        return 1;
      GXY_CUDA(cudaEventRecord(tb, st));
      S.kernel_launches += 1;
      S.waves++;
    }
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      GXY_CUDA(cudaMemcpyAsync(&fq[p], v->fq.p, sizeof(FusedQueues), cudaMemcpyDeviceToHost, v->ctx->stream));
    }
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
      S.primary_rays += (long long)fq[p].n_generated;
      S.ao_rays += (long long)fq[p].n_hits * lights.n_ao;
      S.shadow_rays += (long long)fq[p].n_hits * (lights.shadows ? lights.n_lights : 0);
      S.traced_rays += (long long)fq[p].n_generated + (long long)fq[p].n_hits * n_sec_per_hit;
      S.dequeued_rays += (long long)fq[p].n_primary32 + (long long)fq[p].n_hits * n_sec_per_hit;
      S.terminated_rays += (long long)fq[p].n_terminated;
      S.nodes_visited += (long long)fq[p].nodes;
      S.prims_tested += (long long)fq[p].prims;
      n_cur[p] = (int)fq[p].n_spill;  // rays bound for a neighbour partition, classification already set
    }
    PT.mark("fusedPS");
  } else {
    // ---- generation (Camera::generate_initial_rays; every partition scans the full window) ----
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      if (v->cur.reserve(npix, false, v->ctx->stream)) return 1;
      if (launch_generate(v->P, C, w, h, tiled_order(), v->cur.v, nullptr, v->block_sums.p, v->small.p, v->ctx->stream)) return 1;
      S.kernel_launches += 3;
    }
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      GXY_CUDA(cudaMemcpyAsync(&n_cur[p], v->small.p, sizeof(int), cudaMemcpyDeviceToHost, v->ctx->stream));
      GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
      S.primary_rays += n_cur[p];
    }
  }

  // small[] layout (ints): [0] nhit/ngen  [8 .. 8+nranks) counts  [8+nranks .. 8+2nranks] offsets  then cursor
  std::vector<std::vector<int>> send_counts(nparts, std::vector<int>(nranks, 0)), send_offsets(nparts, std::vector<int>(nranks + 1, 0));
  std::vector<int> n_spawn(nparts, 0);
  std::vector<int> all_counts;  // multi-process: nranks x (nranks+1)
  const int wave_cap = 100000;
  int wave = 0;
  for (; wave < wave_cap; wave++) {
    // the spill lists of the fused kernels enter the loop at its exchange step
    const bool lists_classified = fused && wave == 0;
    long long pending_local = 0;
    for (int p = 0; p < nparts; p++) pending_local += n_cur[p];
    if (!multi_proc && pending_local == 0) break;
    // ---- trace + scan (async per partition) ----
    for (int p = 0; p < nparts && !lists_classified; p++) {
      gxy_vis *v = parts[p];
      const int n = n_cur[p];
      if (n == 0) continue;
      if (use_device(v->ctx)) return 1;
      cudaStream_t st = v->ctx->stream;
      cudaEvent_t ta, tb;
      if (FE.new_pair(&ta, &tb)) return 1;
      GXY_CUDA(cudaEventRecord(ta, st));
      // GXY_MARCH_TMA=1: the primaries (compact 16x8-pixel beams) of a one-volume scene go through the TMA-staged march
      const bool tma = wave == 0 && march_tma_on() && march_tma_eligible(v->P);
      if (tma) {
        const float ax = fabsf(C.vdir.x), ay = fabsf(C.vdir.y), az = fabsf(C.vdir.z);
        const int axis = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
        if (launch_march_tma(v->P, v->cur.v, n, epsilon, axis, v->counters.p + 1, v->counters.p + 2, st)) return 1;
      } else if (launch_trace(v->P, v->cur.v, n, epsilon, nullptr, !v->has_dvr, v->counters.p + 1, st)) return 1;
      GXY_CUDA(cudaEventRecord(tb, st));
      if (v->hit_index.reserve((size_t)2 * n) || v->block_sums.reserve((size_t)n / 1024 + 2)) return 1;
      if (launch_hit_scan(v->cur.v, n, v->hit_index.p, v->block_sums.p, v->small.p, st)) return 1;
      S.kernel_launches += 4;
      S.traced_rays += n;
      S.dequeued_rays += n;
      S.waves++;
    }
    // ---- read hit counts, size the next lists, shade + spawn, classify, accumulate, sort ----
    std::vector<int> nhit(nparts, 0);
    for (int p = 0; p < nparts && !lists_classified; p++) {
      if (n_cur[p] == 0) continue;
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      GXY_CUDA(cudaMemcpyAsync(&nhit[p], v->small.p, sizeof(int), cudaMemcpyDeviceToHost, v->ctx->stream));
    }
    for (int p = 0; p < nparts; p++) {
      gxy_vis *v = parts[p];
      n_spawn[p] = 0;
      std::fill(send_counts[p].begin(), send_counts[p].end(), 0);
      const int n = n_cur[p];
      if (n == 0) continue;
      if (use_device(v->ctx)) return 1;
      cudaStream_t st = v->ctx->stream;
      GXY_CUDA(cudaStreamSynchronize(st));
      if (!lists_classified) {
        const long long nsp = (long long)nhit[p] * n_sec_per_hit;
        GXY_CHECK(nsp < (1ll << 31) - (1ll << 24), "secondary ray list too large (%lld rays)", nsp);
        n_spawn[p] = (int)nsp;
        S.ao_rays += (long long)nhit[p] * lights.n_ao;
        S.shadow_rays += (long long)nhit[p] * (lights.shadows ? lights.n_lights : 0);
        if (v->next.reserve((size_t)std::max<long long>(nsp, 1), false, st)) return 1;
        if (launch_shade_spawn(L, v->cur.v, n, v->hit_index.p, v->small.p, v->next.v, epsilon, st)) return 1;
        if (launch_classify(v->P, v->cur.v, n, st)) return 1;
        if (launch_accumulate(v->cur.v, n, v->fb.p, w, h, v->counters.p, st)) return 1;
        S.kernel_launches += 3 + (lights.n_ao > 0 ? 1 : 0);
      }
      if (nranks > 1) {
        if (v->send.reserve((size_t)n, false, st)) return 1;
        int *d_counts = v->small.p + 8, *d_offsets = d_counts + nranks, *d_cursor = d_offsets + nranks + 1;
        if (launch_partition_by_destination(v->cur.v, n, nranks, multi_proc ? rank0 : p, v->send.v, d_counts, d_offsets, d_cursor, st)) return 1;
        S.kernel_launches += 3;
        GXY_CUDA(cudaMemcpyAsync(send_counts[p].data(), d_counts, sizeof(int) * nranks, cudaMemcpyDeviceToHost, st));
        GXY_CUDA(cudaMemcpyAsync(send_offsets[p].data(), d_offsets, sizeof(int) * (nranks + 1), cudaMemcpyDeviceToHost, st));
      }
    }
    for (int p = 0; p < nparts; p++) {
      if (n_cur[p] == 0) continue;
      if (use_device(parts[p]->ctx)) return 1;
      GXY_CUDA(cudaStreamSynchronize(parts[p]->ctx->stream));
    }
    PT.mark("wave_compute");
    // ---- exchange ----
    std::vector<int> n_next(nparts, 0);
    for (int p = 0; p < nparts; p++) n_next[p] = n_spawn[p];
    if (!multi_proc) {
      for (int src = 0; src < nparts; src++)
        for (int dst = 0; dst < nparts; dst++) {
          const int cnt = send_counts[src][dst];
          if (!cnt) continue;
          gxy_vis *vs = parts[src], *vd = parts[dst];
          S.forwarded_rays += cnt;
          if (use_device(vd->ctx)) return 1;
          cudaStream_t st = vd->ctx->stream;
          if (vd->next.reserve((size_t)n_next[dst] + cnt, true, st)) return 1;
          if (vs->ctx->device == vd->ctx->device) {
            if (launch_copy_rays(vd->next.v, (size_t)n_next[dst], vs->send.v, (size_t)send_offsets[src][dst], cnt, st)) return 1;
          } else {
            // stage through the destination's recv list with peer copies, column by column
            if (vd->recv.reserve((size_t)cnt, false, st)) return 1;
            for (int c = 0; c < 24; c++)
              GXY_CUDA(cudaMemcpyPeerAsync(vd->recv.base + (size_t)c * vd->recv.cap, vd->ctx->device,
                                           vs->send.base + (size_t)c * vs->send.cap + send_offsets[src][dst], vs->ctx->device,
                                           sizeof(float) * cnt, st));
            if (launch_copy_rays(vd->next.v, (size_t)n_next[dst], vd->recv.v, 0, cnt, st)) return 1;
          }
          S.kernel_launches += 1;
          n_next[dst] += cnt;
        }
      if (nparts > 1)  // the send lists are rewritten by the next wave: wait for every consumer
        for (int p = 0; p < nparts; p++) {
          if (use_device(parts[p]->ctx)) return 1;
          GXY_CUDA(cudaStreamSynchronize(parts[p]->ctx->stream));
        }
    } else {
      // one process per GPU: exchange counts, then the ray columns, with NCCL over NVLink
      gxy_vis *v = parts[0];
      cudaStream_t st = v->ctx->stream;
      const int me = rank0;
      std::vector<int> mine(nranks + 1, 0);
      for (int d = 0; d < nranks; d++) mine[d] = send_counts[0][d];
      mine[nranks] = n_spawn[0];
      if (v->io_i.reserve((size_t)(nranks + 1) * (nranks + 1))) return 1;
      int *d_mine = v->io_i.p, *d_all = v->io_i.p + (nranks + 1);
      GXY_CUDA(cudaMemcpyAsync(d_mine, mine.data(), sizeof(int) * (nranks + 1), cudaMemcpyHostToDevice, st));
      GXY_NCCL(g_nccl.AllGather(d_mine, d_all, nranks + 1, ncclInt32, v->ctx->comm, st));
      all_counts.resize((size_t)nranks * (nranks + 1));
      GXY_CUDA(cudaMemcpyAsync(all_counts.data(), d_all, sizeof(int) * nranks * (nranks + 1), cudaMemcpyDeviceToHost, st));
      GXY_CUDA(cudaStreamSynchronize(st));
      long long global_pending = 0;
      int n_recv_total = 0;
      std::vector<int> recv_from(nranks, 0);
      for (int r = 0; r < nranks; r++) {
        for (int d = 0; d < nranks; d++) global_pending += all_counts[(size_t)r * (nranks + 1) + d];
        global_pending += all_counts[(size_t)r * (nranks + 1) + nranks];
        recv_from[r] = all_counts[(size_t)r * (nranks + 1) + me];
        n_recv_total += recv_from[r];
      }
      for (int d = 0; d < nranks; d++) S.forwarded_rays += send_counts[0][d];
      if (n_recv_total) {
        if (v->recv.reserve((size_t)n_recv_total, false, st)) return 1;
      }
      bool any = false;
      for (int r = 0; r < nranks; r++) any = any || recv_from[r] || send_counts[0][r];
      if (any) {
        GXY_NCCL(g_nccl.GroupStart());
        int roff = 0;
        for (int r = 0; r < nranks; r++) {
          if (r != me && send_counts[0][r])
            for (int c = 0; c < 24; c++)
              GXY_NCCL(g_nccl.Send(v->send.base + (size_t)c * v->send.cap + send_offsets[0][r], send_counts[0][r], ncclFloat32, r, v->ctx->comm, st));
          if (r != me && recv_from[r])
            for (int c = 0; c < 24; c++)
              GXY_NCCL(g_nccl.Recv(v->recv.base + (size_t)c * v->recv.cap + roff, recv_from[r], ncclFloat32, r, v->ctx->comm, st));
          roff += recv_from[r];
        }
        GXY_NCCL(g_nccl.GroupEnd());
        if (v->next.reserve((size_t)n_next[0] + n_recv_total, true, st)) return 1;
        roff = 0;
        for (int r = 0; r < nranks; r++) {
          if (!recv_from[r]) continue;
          if (r == me) {
            if (launch_copy_rays(v->next.v, (size_t)n_next[0], v->send.v, (size_t)send_offsets[0][me], recv_from[r], st)) return 1;
          } else {
            if (launch_copy_rays(v->next.v, (size_t)n_next[0], v->recv.v, (size_t)roff, recv_from[r], st)) return 1;
          }
          S.kernel_launches += 1;
          n_next[0] += recv_from[r];
          roff += recv_from[r];
        }
      }
      if (global_pending == 0) { n_cur[0] = 0; break; }
    }
    PT.mark("exchange");
    for (int p = 0; p < nparts; p++) {
      std::swap(parts[p]->cur, parts[p]->next);
      n_cur[p] = n_next[p];
    }
  }

  GXY_CHECK(wave < wave_cap, "the wave loop did not terminate after %d waves", wave_cap);
  PT.mark("exchange_tail");
  // ---- framebuffer: partial sums -> owner (SendPixelsMsg / AddLocalPixels) ----
  if (!multi_proc) {
    gxy_vis *v0 = parts[0];
    for (int p = 1; p < nparts; p++) {
      gxy_vis *v = parts[p];
      if (use_device(v->ctx)) return 1;
      GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
      if (use_device(v0->ctx)) return 1;
      const float *src = v->fb.p;
      if (v->ctx->device != v0->ctx->device) {
        if (v0->fb_tmp.reserve((size_t)npix * 4)) return 1;
        GXY_CUDA(cudaMemcpyPeerAsync(v0->fb_tmp.p, v0->ctx->device, v->fb.p, v->ctx->device, sizeof(float) * 4 * npix, v0->ctx->stream));
        src = v0->fb_tmp.p;
      }
      if (launch_fb_add(v0->fb.p, src, (size_t)npix * 4, v0->ctx->stream)) return 1;
      S.kernel_launches += 1;
    }
  } else {
    gxy_vis *v = parts[0];
    GXY_NCCL(g_nccl.Reduce(v->fb.p, v->fb.p, (size_t)npix * 4, ncclFloat32, ncclSum, 0, v->ctx->comm, v->ctx->stream));
  }
  if (use_device(ctx0)) return 1;
  GXY_CUDA(cudaEventRecord(ev1, ctx0->stream));
  // the frame's counters and error flags travel to page-locked memory behind the end-of-frame event: one host wait per
  // partition instead of four blocking copies
  for (int p = 0; p < nparts; p++) {
    gxy_vis *v = parts[p];
    if (use_device(v->ctx)) return 1;
    if (!v->h_tail) GXY_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&v->h_tail), sizeof(gxy_vis::FrameTailT), cudaHostAllocPortable));
    cudaStream_t st = v->ctx->stream;
    GXY_CUDA(cudaMemcpyAsync(v->h_tail->counters, v->counters.p, sizeof v->h_tail->counters, cudaMemcpyDeviceToHost, st));
    GXY_CUDA(cudaMemcpyAsync(v->h_tail->trav, v->P.trav_counters, sizeof v->h_tail->trav, cudaMemcpyDeviceToHost, st));
    GXY_CUDA(cudaMemsetAsync(v->P.trav_counters, 0, sizeof v->h_tail->trav, st));
    GXY_CUDA(cudaMemcpyAsync(&v->h_tail->error, v->d_error, sizeof(int), cudaMemcpyDeviceToHost, st));
    GXY_CUDA(cudaMemsetAsync(v->d_error, 0, sizeof(int), st));  // reported once: a transient error does not poison the Visualization
  }
  if (use_device(ctx0)) return 1;
  GXY_CUDA(cudaEventSynchronize(ev1));
  PT.mark("fb_reduce");
  PT.report(rank0);
  if (PT.on) fprintf(stderr, "[gxy_render rank %d] waves=%lld traced=%lld forwarded=%lld\n", rank0, S.waves, S.traced_rays, S.forwarded_rays);
  cudaEventElapsedTime(&S.device_ms, ev0, ev1);
  for (auto &e : trace_events) {
    float ms = 0.f;
    cudaEventSynchronize(e.second);
    cudaEventElapsedTime(&ms, e.first, e.second);
    S.trace_ms += ms;
  }
  for (int p = 0; p < nparts; p++) {
    gxy_vis *v = parts[p];
    if (use_device(v->ctx)) return 1;
    GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
    const gxy_vis::FrameTailT &t = *v->h_tail;
    S.terminated_rays += (long long)t.counters[0];
    S.volume_samples += (long long)t.counters[1];
    S.staged_samples += (long long)t.counters[2];
    S.nodes_visited += (long long)t.trav[0];
    S.prims_tested += (long long)t.trav[1];
    GXY_CHECK(t.error == 0, "device error flag %d (1: BVH traversal stack overflow, 3: ray list / inbox capacity, 4: peer barrier timeout, 6: TMA copy did not complete, 7: translucent surface on the fused path)", t.error);
  }
  if (stats) *stats = S;
  return 0;
}

// which frame schedule a set of partitions takes: 0 = synchronous list/fused loop (render_sync), 1 = single-partition band
// pipeline as a Flight, 2 = one process per GPU over peer arenas as a Flight
static int frame_kind(int nparts, gxy_vis *const *parts) {
  bool fused = true;
  for (int p = 0; p < nparts; p++) fused = fused && parts[p]->P.n_volvis == 0 && !parts[p]->geoms.empty();
  // PathLines take the list-path kernels (one thread per ray).  GXY_FUSED_CURVES=1: the CURVES instantiations of the persistent frame
  // kernels (round-Bezier test inside the cooperative primitive passes, 140 registers, 3 CTAs per SM) -- parity-green on the B200 but
  // SLOWER: 40 000 segments 12.6 -> 18.1 ms per frame, 800 000 segments 18.1 -> 24.7 ms (profiles/r02_f_*): a warp-wide pass whose
  // lanes each run a divergent sub-division loop costs more than it saves
  if (!(getenv("GXY_FUSED_CURVES") && atoi(getenv("GXY_FUSED_CURVES")) != 0))
    for (int p = 0; p < nparts; p++)
      for (const GeomOp &g : parts[p]->geoms) fused = fused && g.kind != 2;
  if (const char *e = getenv("GXY_FUSED")) fused = fused && atoi(e) != 0;
  if (!fused) {
    // one process per GPU, a Visualization with volumes: the list kernels with device-side lengths over the peer arenas
    // (flight_enqueue_volume); GXY_VOLUME_FLIGHTS=0 / GXY_PEER=0: the synchronous NCCL list loop
    if (parts[0]->ctx->comm && nparts == 1 && parts[0]->P.n_volvis >= 1 && parts[0]->samplers.empty() &&
        !(getenv("GXY_PEER") && atoi(getenv("GXY_PEER")) == 0) && !(getenv("GXY_VOLUME_FLIGHTS") && atoi(getenv("GXY_VOLUME_FLIGHTS")) == 0))
      return 3;
    return 0;
  }
  if (parts[0]->ctx->comm) {
    if (const char *e = getenv("GXY_PEER"))
      if (atoi(e) == 0) return 0;
    return 2;
  }
  bool alone = nparts == 1;
  for (int f = 0; f < 6 && alone; f++) alone = parts[0]->neighbors[f] < 0;
  return alone ? 1 : 0;
}

int gxy_render_submit(int nparts, gxy_vis *const *parts, const gxy_camera *cam, const gxy_lighting *lights_in, int w, int h, float epsilon,
                      int slot) {
  GXY_CHECK(nparts >= 1 && parts && cam && lights_in && w > 0 && h > 0, "gxy_render: bad arguments");
  for (int p = 0; p < nparts; p++)
    if (check_vis(parts[p])) return 1;
  gxy_vis *v = parts[0];
  gxy_context *c = v->ctx;
  GXY_CHECK(!c->comm || nparts == 1, "with a communicator attached each process drives exactly one partition");
  if (use_device(c)) return 1;
  Flight *Fp = nullptr;
  if (flight_get(v, slot, &Fp)) return 1;
  Flight &F = *Fp;
  GXY_CHECK(!F.pending, "frame slot %d is still in flight (call gxy_render_wait first)", slot);
  gxy_lighting lights;
  if (gxy_resolve_lights(lights_in, cam, &lights)) return 1;
  GXY_CHECK(lights.n_lights >= 1, "lighting needs at least one light");
  memset(&F.S, 0, sizeof F.S);
  F.lights = lights;
  F.L = make_dev_lights(lights);
  F.n_sec_per_hit = lights.n_ao + (lights.shadows ? lights.n_lights : 0);
  F.w = w; F.h = h; F.epsilon = epsilon;
  F.sync_done = false;
  F.peer = false;
  const DevCamera C = make_dev_camera(*cam, w, h);
  // an image still being converted for a download reads the framebuffer this frame is about to clear
  for (int s = 0; s < 2; s++)
    if (v->async_ready[s]) GXY_CUDA(cudaStreamWaitEvent(F.st, v->async_ready[s], 0));
  int kind = frame_kind(nparts, parts);
  F.volume = false;
  if (kind == 2 || kind == 3) {
    unsigned long long cap = (unsigned long long)w * h * (unsigned long long)(1 + F.n_sec_per_hit);
    if (kind == 3) {  // the lists hold the tile-ordered primaries (generation slots) or the secondaries + what the neighbours sent
      const unsigned long long slots = (unsigned long long)((w + 15) / 16) * ((h + 7) / 8) * 128ull;
      cap = std::max(slots, (unsigned long long)w * h) * (unsigned long long)(1 + F.n_sec_per_hit) + 1024ull;
    }
    GXY_CHECK(cap < (1ull << 31) - (1ull << 24), "peer inbox too large (%llu records)", cap);
    if (c->arena.disabled) F.arena.disabled = true;  // (decided once, by flight 0, on every rank alike)
    if (ensure_peer_arena(c, F.arena, (unsigned)(w * h), (unsigned)cap)) return 1;
    if (F.arena.disabled) { c->arena.disabled = true; kind = 0; }
    F.v_cap = (int)cap;
  }
  if (kind == 1) {
    if (flight_submit_single(v, F, C, w, h, epsilon)) return 1;
  } else if (kind == 2 || kind == 3) {
    F.peer = true;
    F.volume = kind == 3;
    if (flight_submit_peer(v, F, C, w, h, epsilon)) return 1;
  } else {
    // synchronous schedules render here; the image is kept in the flight's own buffer until the wait
    if (render_sync(nparts, parts, cam, lights_in, w, h, epsilon, &F.S)) return 1;
    if (c->rank == 0 || !c->comm) {
      if (F.fb.reserve((size_t)w * h * 4)) return 1;
      GXY_CUDA(cudaMemcpyAsync(F.fb.p, v->fb_result, sizeof(float) * 4 * (size_t)w * h, cudaMemcpyDeviceToDevice, c->stream));
      GXY_CUDA(cudaStreamSynchronize(c->stream));
      F.fb_result = F.fb.p;
    } else F.fb_result = v->fb_result;
    F.sync_done = true;
  }
  F.pending = true;
  return 0;
}

int gxy_render_wait(int nparts, gxy_vis *const *parts, int slot, gxy_stats *stats) {
  GXY_CHECK(nparts >= 1 && parts && parts[0], "gxy_render_wait: bad arguments");
  gxy_vis *v = parts[0];
  GXY_CHECK(slot >= 0 && slot < GXY_MAX_FLIGHTS && v->flights[slot], "frame slot %d was never submitted", slot);
  if (use_device(v->ctx)) return 1;
  return flight_wait(v, *v->flights[slot], stats);
}

int gxy_render_max_slots(void) { return GXY_MAX_FLIGHTS; }

int gxy_debug_tile_rect(const gxy_camera *cam, int w, int h, const float lo[3], const float hi[3], int rect[4]) {
  if (!cam || !lo || !hi || !rect || w <= 0 || h <= 0) return -1;
  const DevCamera C = make_dev_camera(*cam, w, h);
  return peer_tile_rect(C, lo, hi, w, h, rect) ? 1 : 0;
}

int gxy_render(int nparts, gxy_vis *const *parts, const gxy_camera *cam, const gxy_lighting *lights_in, int w, int h, float epsilon,
               gxy_stats *stats) {
  GXY_CHECK(nparts >= 1 && parts && cam && lights_in && w > 0 && h > 0, "gxy_render: bad arguments");
  for (int p = 0; p < nparts; p++)
    if (check_vis(parts[p])) return 1;
  if (frame_kind(nparts, parts) == 0 || (parts[0]->ctx->comm && parts[0]->ctx->arena.disabled)) return render_sync(nparts, parts, cam, lights_in, w, h, epsilon, stats);
  if (gxy_render_submit(nparts, parts, cam, lights_in, w, h, epsilon, 0)) return 1;
  return gxy_render_wait(nparts, parts, 0, stats);
}

// ---- interactive frame path (Rendering.cpp:104-153) -----------------------------------------------
int gxy_progressive_reset(gxy_vis *v) {
  GXY_CHECK(v, "NULL visualization");
  if (use_device(v->ctx)) return 1;
  if (v->prog_image) {
    const size_t npix = (size_t)v->prog_w * v->prog_h;
    GXY_CUDA(cudaMemsetAsync(v->prog_image, 0, sizeof(float) * 4 * npix, v->ctx->stream));
    GXY_CUDA(cudaMemsetAsync(v->prog_kbuffer, 0, sizeof(int) * npix, v->ctx->stream));
    GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
  }
  v->prog_frame = -1;
  return 0;
}

static int progressive_check(int nparts, gxy_vis *const *parts) {
  for (int p = 0; p < nparts; p++) {
    if (check_vis(parts[p])) return 1;
    GXY_CHECK(parts[p]->ctx->comm == nullptr, "gxy_render_progressive drives all partitions from one process (no communicator)");
    GXY_CHECK(parts[p]->ctx->device == parts[0]->ctx->device, "gxy_render_progressive: all partitions of one call live on one device");
  }
  return use_device(parts[0]->ctx);
}
// Rendering::local_commit (:218-238): the displayed image, its frame stamps and the touched mask for a w x h window
static int progressive_prepare(gxy_vis *own, int w, int h) {
  const size_t npix = (size_t)w * h;
  if (!own->prog_image || own->prog_w != w || own->prog_h != h) {
    if (own->prog_image) cudaFree(own->prog_image);
    if (own->prog_kbuffer) cudaFree(own->prog_kbuffer);
    if (own->prog_touched) cudaFree(own->prog_touched);
    own->prog_image = nullptr; own->prog_kbuffer = nullptr; own->prog_touched = nullptr;
    own->prog_w = 0; own->prog_h = 0;
    float *img = nullptr;
    int *kb = nullptr;
    unsigned char *tc = nullptr;
    if (cudaMalloc(&img, sizeof(float) * 4 * npix) != cudaSuccess || cudaMalloc(&kb, sizeof(int) * npix) != cudaSuccess ||
        cudaMalloc(&tc, npix) != cudaSuccess) {
      gxy_set_error("out of device memory for the progressive image (%d x %d): %s", w, h, cudaGetErrorString(cudaGetLastError()));
      if (img) cudaFree(img);
      if (kb) cudaFree(kb);
      if (tc) cudaFree(tc);
      return 1;
    }
    own->prog_image = img; own->prog_kbuffer = kb; own->prog_touched = tc;
    own->prog_w = w; own->prog_h = h;
    if (gxy_progressive_reset(own)) return 1;
  }
  return 0;
}
// ACCUMULATE_PIXEL for a whole finished frame (own->fb_result): the pixels this frame wrote are those for which some partition
// originated a primary ray; each takes the new sum if its stamp is older, adds it if the frame number repeats
static int progressive_merge(int nparts, gxy_vis *const *parts, const gxy_camera *cam, int w, int h, int frame) {
  gxy_vis *own = parts[0];
  cudaStream_t st = own->ctx->stream;
  const size_t npix = (size_t)w * h;
  const DevCamera C = make_dev_camera(*cam, w, h);
  GXY_CUDA(cudaMemsetAsync(own->prog_touched, 0, npix, st));
  for (int p = 0; p < nparts; p++) {
    gxy_vis *v = parts[p];
    if (v->cur.reserve(npix, false, st) || v->block_sums.reserve(npix / 1024 + 2) || v->small.reserve(64)) return 1;
    if (launch_generate(v->P, C, w, h, false, v->cur.v, nullptr, v->block_sums.p, v->small.p, v->ctx->stream)) return 1;
    int n = 0;
    GXY_CUDA(cudaMemcpyAsync(&n, v->small.p, sizeof(int), cudaMemcpyDeviceToHost, v->ctx->stream));
    GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
    if (launch_mark_touched(v->cur.v, n, w, own->prog_touched, st)) return 1;
    GXY_CUDA(cudaStreamSynchronize(st));
  }
  GXY_CHECK(own->fb_result, "gxy_render_progressive: no frame on the image owner");
  if (launch_merge_stamped(own->fb_result, own->prog_image, own->prog_kbuffer, own->prog_touched, (int)npix, frame, st)) return 1;
  GXY_CUDA(cudaStreamSynchronize(st));
  if (frame > own->prog_frame) own->prog_frame = frame;
  return 0;
}

int gxy_render_progressive(int nparts, gxy_vis *const *parts, const gxy_camera *cam, const gxy_lighting *lights, int w, int h, float epsilon,
                           int frame, gxy_stats *stats) {
  GXY_CHECK(nparts >= 1 && parts && cam && lights && w > 0 && h > 0, "gxy_render_progressive: bad arguments");
  if (progressive_check(nparts, parts)) return 1;
  gxy_vis *own = parts[0];
  if (progressive_prepare(own, w, h)) return 1;
  if (stats) memset(stats, 0, sizeof *stats);
  if (frame < own->prog_frame) return 0;  // AddLocalPixels :138: the pixels of a stale frame are dropped
  if (gxy_render(nparts, parts, cam, lights, w, h, epsilon, stats)) return 1;
  if (progressive_merge(nparts, parts, cam, w, h, frame)) return 1;
  if (stats) stats->kernel_launches += 4 * nparts + 1;
  return 0;
}

// The same with frames in flight (gxyviewer keeps rendering while older frames are still on their way, Rendering.cpp:104-153):
// a frame is submitted to a frame slot and merged into the displayed image when it is waited for.  Frames may be waited for in any
// order: one that arrives after a newer frame was merged is stale and dropped, exactly as AddLocalPixels drops the pixels of a
// superseded frame (:138-152).
int gxy_render_progressive_submit(int nparts, gxy_vis *const *parts, const gxy_camera *cam, const gxy_lighting *lights, int w, int h,
                                  float epsilon, int frame, int slot) {
  GXY_CHECK(nparts >= 1 && parts && cam && lights && w > 0 && h > 0, "gxy_render_progressive_submit: bad arguments");
  if (progressive_check(nparts, parts)) return 1;
  gxy_vis *own = parts[0];
  if (progressive_prepare(own, w, h)) return 1;
  if (gxy_render_submit(nparts, parts, cam, lights, w, h, epsilon, slot)) return 1;
  Flight &F = *own->flights[slot];
  F.prog_frame = frame;
  F.prog_cam = *cam;
  return 0;
}
int gxy_render_progressive_wait(int nparts, gxy_vis *const *parts, int slot, gxy_stats *stats, int *merged) {
  GXY_CHECK(nparts >= 1 && parts && parts[0], "gxy_render_progressive_wait: bad arguments");
  if (progressive_check(nparts, parts)) return 1;
  gxy_vis *own = parts[0];
  if (merged) *merged = 0;
  if (gxy_render_wait(nparts, parts, slot, stats)) return 1;
  Flight &F = *own->flights[slot];
  GXY_CHECK(own->prog_image && own->prog_w == F.w && own->prog_h == F.h, "the progressive image was re-allocated while this frame was in flight");
  if (F.prog_frame < own->prog_frame) return 0;  // superseded while in flight: dropped
  if (progressive_merge(nparts, parts, &F.prog_cam, F.w, F.h, F.prog_frame)) return 1;
  if (merged) *merged = 1;
  return 0;
}

int gxy_progressive_download_rgba32f(gxy_vis *v, float *fb) {
  GXY_CHECK(v && fb, "gxy_progressive_download_rgba32f: NULL argument");
  GXY_CHECK(v->prog_image, "no progressive frame yet (call gxy_render_progressive)");
  if (use_device(v->ctx)) return 1;
  GXY_CUDA(cudaMemcpy(fb, v->prog_image, sizeof(float) * 4 * (size_t)v->prog_w * v->prog_h, cudaMemcpyDeviceToHost));
  return 0;
}

int gxy_progressive_download_rgba8(gxy_vis *v, unsigned char *rgba) {
  GXY_CHECK(v && rgba, "gxy_progressive_download_rgba8: NULL argument");
  GXY_CHECK(v->prog_image, "no progressive frame yet (call gxy_render_progressive)");
  if (use_device(v->ctx)) return 1;
  const size_t npix = (size_t)v->prog_w * v->prog_h;
  if (v->rgba8.reserve(npix * 4)) return 1;
  if (launch_tonemap(v->prog_image, v->prog_w, v->prog_h, v->rgba8.p, v->ctx->stream)) return 1;
  GXY_CUDA(cudaMemcpyAsync(rgba, v->rgba8.p, npix * 4, cudaMemcpyDeviceToHost, v->ctx->stream));
  GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

int gxy_frame_download_rgba32f(gxy_vis *v, float *fb) {
  if (check_vis(v)) return 1;
  GXY_CHECK(v->fb_w > 0 && v->fb_result && fb, "no frame rendered yet");
  GXY_CUDA(cudaMemcpyAsync(fb, v->fb_result, sizeof(float) * 4 * (size_t)v->fb_w * v->fb_h, cudaMemcpyDeviceToHost, v->ctx->stream));
  GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

int gxy_frame_download_rgba8(gxy_vis *v, unsigned char *rgba) {
  if (check_vis(v)) return 1;
  GXY_CHECK(v->fb_w > 0 && v->fb_result && rgba, "no frame rendered yet");
  const size_t n = (size_t)v->fb_w * v->fb_h * 4;
  if (v->rgba8.reserve(n)) return 1;
  if (launch_tonemap(v->fb_result, v->fb_w, v->fb_h, v->rgba8.p, v->ctx->stream)) return 1;
  GXY_CUDA(cudaMemcpyAsync(rgba, v->rgba8.p, n, cudaMemcpyDeviceToHost, v->ctx->stream));
  GXY_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

int gxy_frame_download_rgba8_async(gxy_vis *v, unsigned char *rgba) {
  if (check_vis(v)) return 1;
  GXY_CHECK(v->fb_w > 0 && v->fb_result && rgba, "no frame rendered yet");
  gxy_context *c = v->ctx;
  if (!c->copy_stream) GXY_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  const int s = v->async_slot;
  v->async_slot ^= 1;
  if (!v->async_ready[s]) {
    GXY_CUDA(cudaEventCreateWithFlags(&v->async_ready[s], cudaEventDisableTiming));
    GXY_CUDA(cudaEventCreateWithFlags(&v->async_done[s], cudaEventDisableTiming));
  } else {
    GXY_CUDA(cudaStreamWaitEvent(c->stream, v->async_done[s], 0));  // the copy that last used this staging image
  }
  const size_t n = (size_t)v->fb_w * v->fb_h * 4;
  if (v->rgba8_async[s].reserve(n)) return 1;
  if (launch_tonemap(v->fb_result, v->fb_w, v->fb_h, v->rgba8_async[s].p, c->stream)) return 1;
  GXY_CUDA(cudaEventRecord(v->async_ready[s], c->stream));
  GXY_CUDA(cudaStreamWaitEvent(c->copy_stream, v->async_ready[s], 0));
  GXY_CUDA(cudaMemcpyAsync(rgba, v->rgba8_async[s].p, n, cudaMemcpyDeviceToHost, c->copy_stream));
  GXY_CUDA(cudaEventRecord(v->async_done[s], c->copy_stream));
  return 0;
}

int gxy_frame_download_wait(gxy_vis *v) {
  if (check_vis(v)) return 1;
  if (v->ctx->copy_stream) GXY_CUDA(cudaStreamSynchronize(v->ctx->copy_stream));
  return 0;
}

int gxy_host_alloc(size_t bytes, void **out) {
  GXY_CHECK(out, "gxy_host_alloc: NULL argument");
  GXY_CHECK(gxy_device_count() > 0, "no CUDA device available: galaxy_b200 has no CPU fallback");
  GXY_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return 0;
}
void gxy_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- multi-process -------------------------------------------------------------------------------
int gxy_comm_unique_id(unsigned char id[128]) {
  if (load_nccl()) return 1;
  ncclUniqueId u;
  GXY_NCCL(g_nccl.GetUniqueId(&u));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &u, 128);
  return 0;
}

int gxy_comm_init(gxy_context *c, int rank, int nranks, const unsigned char id[128]) {
  if (load_nccl()) return 1;
  if (use_device(c)) return 1;
  GXY_CHECK(!c->comm, "communicator already initialised");
  ncclUniqueId u;
  memcpy(&u, id, 128);
  GXY_NCCL(g_nccl.CommInitRank(&c->comm, nranks, u, rank));
  c->rank = rank;
  c->nranks = nranks;
  return 0;
}

int gxy_comm_destroy(gxy_context *c) {
  if (c->comm) {
    if (use_device(c)) return 1;
    cudaStreamSynchronize(c->stream);
    peer_arena_unmap(c->arena, c->rank);
    if (c->arena.base) cudaFree(c->arena.base);
    c->arena = PeerArena();
    GXY_NCCL(g_nccl.CommDestroy(c->comm));
    c->comm = nullptr;
    c->rank = 0;
    c->nranks = 1;
  }
  return 0;
}

}  // extern "C"
