// gxy_bvh.cu -- device-side BVH build over all geometry operators of a Visualization.
// Replaces the Embree build that runs inside ospCommit(model) (src/renderer/Visualization.cpp:219-277
// -> ospray/common/Model.cpp:50-101 -> rtcCommitScene; Embree: binned-SAH BVH8.Triangle4,
// embree/kernels/common/scene.cpp:123-137).
//
// B200 design: everything on the device, no host round trips proportional to N.
//   1. primitive boxes + centroid bounds            (one pass, HBM-bound)
//   2. 63-bit Morton keys, radix sort (cub::DeviceRadixSort; build-time only, not on the frame path)
//   3. Karras 2012 binary radix tree + bottom-up refit with atomic flags
//   4. top-down collapse, level by level, into 8-wide nodes: repeatedly open the child with the
//      largest surface area; subtrees of <= LEAF_MAX primitives become leaves; child boxes
//      quantised to 8 bits in a per-node power-of-two frame (conservative), children placed in
//      octant order; the internal children of a node get consecutive node ids (in slot order)
//      and its leaf primitives consecutive record positions (compressed-wide-BVH addressing)
//   5. primitive records (48 B: v0,e1,e2 | centre,radius + ids) written in node order
#include <chrono>
#include "gxy_internal.h"
#include "gxy_curve.cuh"

#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace gxy {

#define LEAF_MAX 3  // 8 slots x 3 primitives = the 24 primitive bits of a hit mask
#define MAX_BUILD_GEOMS 16

struct BuildGeoms {
  int n;
  long long offset[MAX_BUILD_GEOMS + 1];
  GeomBuildInput g[MAX_BUILD_GEOMS];
};

__device__ __forceinline__ float sphere_radius(const GeomBuildInput &g, long long i) {
  // DataDrivenSpheres.ispc:65-88
  if (g.data && g.value0 != g.value1) {
    const float dataval = g.data[i];
    const float d = (dataval - g.value0) / (g.value1 - g.value0);
    if (d > 1) return g.radius1;
    else if (d < 0) return g.radius0;
    else return g.radius0 + d * (g.radius1 - g.radius0);
  }
  return g.radius0;
}

__device__ __forceinline__ int find_geom(const BuildGeoms &B, long long p) {
  int k = 0;
  while (k + 1 < B.n && p >= B.offset[k + 1]) k++;
  return k;
}

__device__ __forceinline__ unsigned f2ord(float f) {  // order-preserving float -> uint
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// bounds[0..2] = min centroid (ordered uint), [3..5] = max centroid
__global__ void __launch_bounds__(256)
    prim_bounds_kernel(const __grid_constant__ BuildGeoms B, long long N, float4 *__restrict__ lo, float4 *__restrict__ hi,
                       unsigned *__restrict__ bounds) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (p < N) {
    const int gi = find_geom(B, p);
    const GeomBuildInput &g = B.g[gi];
    const long long i = p - B.offset[gi];
    if (g.kind == 0) {
      for (int j = 0; j < 3; j++) {
        const float *v = g.verts + 3 * (size_t)g.idx[3 * i + j];
        for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], v[k]); mx[k] = fmaxf(mx[k], v[k]); }
      }
    } else if (g.kind == 1) {
      const float r = sphere_radius(g, i);
      for (int k = 0; k < 3; k++) { mn[k] = g.centers[3 * i + k] - r; mx[k] = g.centers[3 * i + k] + r; }
    } else {
      // round Bezier segment: the swept surface lies in the convex hull of the control points grown by the largest
      // control radius (position and radius are Bernstein combinations of the control values)
      const float *c = g.centers + 16 * (size_t)i;
      const float r = fmaxf(fmaxf(fabsf(c[3]), fabsf(c[7])), fmaxf(fabsf(c[11]), fabsf(c[15])));
      for (int k = 0; k < 3; k++) {
        mn[k] = fminf(fminf(c[k], c[4 + k]), fminf(c[8 + k], c[12 + k])) - r;
        mx[k] = fmaxf(fmaxf(c[k], c[4 + k]), fmaxf(c[8 + k], c[12 + k])) + r;
      }
    }
    lo[p] = make_float4(mn[0], mn[1], mn[2], 0.f);
    hi[p] = make_float4(mx[0], mx[1], mx[2], 0.f);
  }
  // block reduce centroid bounds
  __shared__ unsigned s[6];
  if (threadIdx.x < 3) s[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) s[threadIdx.x] = 0u;
  __syncthreads();
  if (p < N)
    for (int k = 0; k < 3; k++) {
      const float c = 0.5f * mn[k] + 0.5f * mx[k];
      atomicMin(&s[k], f2ord(c));
      atomicMax(&s[3 + k], f2ord(c));
    }
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(&bounds[threadIdx.x], s[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(&bounds[threadIdx.x], s[threadIdx.x]);
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void __launch_bounds__(256)
    morton_kernel(long long N, const float4 *__restrict__ lo, const float4 *__restrict__ hi, const unsigned *__restrict__ bounds,
                  unsigned long long *__restrict__ keys, unsigned *__restrict__ vals) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const float4 a = lo[p], b = hi[p];
  const float c[3] = {0.5f * a.x + 0.5f * b.x, 0.5f * a.y + 0.5f * b.y, 0.5f * a.z + 0.5f * b.z};
  unsigned long long q[3];
  for (int k = 0; k < 3; k++) {
    const float mn = ord2f(bounds[k]), mx = ord2f(bounds[3 + k]);
    const float ext = mx - mn;
    float u = ext > 0.f ? (c[k] - mn) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    q[k] = (unsigned long long)fminf(u * 2097152.0f, 2097151.0f);
  }
  keys[p] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
  vals[p] = (unsigned)p;
}

__global__ void __launch_bounds__(256)
    gather_boxes_kernel(long long N, const unsigned *__restrict__ vals, const float4 *__restrict__ lo, const float4 *__restrict__ hi,
                        float4 *__restrict__ slo, float4 *__restrict__ shi) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const unsigned p = vals[s];
  slo[s] = lo[p];
  shi[s] = hi[p];
}

// ---- Karras 2012 ---------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

// children: >= 0 internal node id, < 0 leaf ~sorted_position
__global__ void __launch_bounds__(256)
    karras_kernel(int n, const unsigned long long *__restrict__ keys, int2 *__restrict__ child, int2 *__restrict__ range,
                  int *__restrict__ parent_int, int *__restrict__ parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  int2 c;
  if (lo == gamma) { c.x = ~gamma; parent_leaf[gamma] = i; }
  else { c.x = gamma; parent_int[gamma] = i; }
  if (hi == gamma + 1) { c.y = ~(gamma + 1); parent_leaf[gamma + 1] = i; }
  else { c.y = gamma + 1; parent_int[gamma + 1] = i; }
  child[i] = c;
  range[i] = make_int2(lo, hi);
  if (i == 0) parent_int[0] = -1;
}

__global__ void __launch_bounds__(256)
    refit_kernel(int n, const int2 *__restrict__ child, const int *__restrict__ parent_int, const int *__restrict__ parent_leaf,
                 const float4 *__restrict__ slo, const float4 *__restrict__ shi, float4 *__restrict__ nlo, float4 *__restrict__ nhi,
                 int *__restrict__ flags) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n) return;
  int node = parent_leaf[leaf];
  while (node >= 0) {
    if (atomicAdd(&flags[node], 1) == 0) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    const int2 c = child[node];
    const float4 al = c.x < 0 ? slo[~c.x] : nlo[c.x], ah = c.x < 0 ? shi[~c.x] : nhi[c.x];
    const float4 bl = c.y < 0 ? slo[~c.y] : nlo[c.y], bh = c.y < 0 ? shi[~c.y] : nhi[c.y];
    nlo[node] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
    nhi[node] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    __threadfence();
    node = parent_int[node];
  }
}

// ---- wide collapse ---------------------------------------------------------------------------------
struct BinView {
  int n;  // leaves
  const int2 *child, *range;
  const float4 *slo, *shi, *nlo, *nhi;
};
__device__ __forceinline__ int bin_count(const BinView &b, int c) { return c < 0 ? 1 : b.range[c].y - b.range[c].x + 1; }
__device__ __forceinline__ int bin_first(const BinView &b, int c) { return c < 0 ? ~c : b.range[c].x; }
__device__ __forceinline__ void bin_box(const BinView &b, int c, float4 &lo, float4 &hi) {
  if (c < 0) { lo = b.slo[~c]; hi = b.shi[~c]; }
  else { lo = b.nlo[c]; hi = b.nhi[c]; }
}
__device__ __forceinline__ float box_area(const float4 &lo, const float4 &hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}

// exponent byte e such that 255 * 2^(e-127) >= ext
__device__ __forceinline__ int quant_exp(float ext) {
  if (!(ext > 0.f)) return 1;
  int e;
  frexpf(ext / 255.0f, &e);  // ext/255 = m * 2^e, m in [0.5,1)  => 2^e >= ext/255
  int eb = e + 127;
  return max(1, min(eb, 254));
}

__global__ void __launch_bounds__(128)
    collapse_kernel(const __grid_constant__ BinView B, int begin, int end, int *__restrict__ wide_bin, WideNode *__restrict__ nodes,
                    int *__restrict__ counter, int capacity, unsigned *__restrict__ prim_counter, unsigned *__restrict__ dest_of,
                    int *__restrict__ err) {
  const int wid = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (wid >= end) return;
  const int root = wide_bin[wid];
  int ch[8];
  int k = 0;
  if (root < 0 || bin_count(B, root) <= LEAF_MAX) {
    ch[k++] = root;  // degenerate: the whole (sub)tree is one leaf
  } else {
    const int2 c = B.child[root];
    ch[0] = c.x; ch[1] = c.y; k = 2;
    while (k < 8) {
      int best = -1;
      float best_a = -1.f;
      for (int m = 0; m < k; m++) {
        if (ch[m] < 0 || bin_count(B, ch[m]) <= LEAF_MAX) continue;
        float4 lo, hi;
        bin_box(B, ch[m], lo, hi);
        const float a = box_area(lo, hi);
        if (a > best_a) { best_a = a; best = m; }
      }
      if (best < 0) break;
      const int2 c2 = B.child[ch[best]];
      ch[best] = c2.x;
      ch[k++] = c2.y;
    }
  }
  // union box
  float4 lo[8], hi[8];
  float ulo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, uhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int m = 0; m < k; m++) {
    bin_box(B, ch[m], lo[m], hi[m]);
    ulo[0] = fminf(ulo[0], lo[m].x); ulo[1] = fminf(ulo[1], lo[m].y); ulo[2] = fminf(ulo[2], lo[m].z);
    uhi[0] = fmaxf(uhi[0], hi[m].x); uhi[1] = fmaxf(uhi[1], hi[m].y); uhi[2] = fmaxf(uhi[2], hi[m].z);
  }
  // slots in octant order relative to the node centre
  const float cx = 0.5f * ulo[0] + 0.5f * uhi[0], cy = 0.5f * ulo[1] + 0.5f * uhi[1], cz = 0.5f * ulo[2] + 0.5f * uhi[2];
  int slot_of[8];
  unsigned used = 0;
  for (int m = 0; m < k; m++) {
    const float mx = 0.5f * lo[m].x + 0.5f * hi[m].x, my = 0.5f * lo[m].y + 0.5f * hi[m].y, mz = 0.5f * lo[m].z + 0.5f * hi[m].z;
    const int pref = (mx > cx ? 1 : 0) | (my > cy ? 2 : 0) | (mz > cz ? 4 : 0);
    int bests = -1, bestd = 99;
    for (int s = 0; s < 8; s++) {
      if (used & (1u << s)) continue;
      const int dist = __popc(s ^ pref);
      if (dist < bestd) { bestd = dist; bests = s; }
    }
    slot_of[m] = bests;
    used |= 1u << bests;
  }
  // slot -> child
  int child_in_slot[8];
  for (int sl = 0; sl < 8; sl++) child_in_slot[sl] = -1;
  for (int m = 0; m < k; m++) child_in_slot[slot_of[m]] = m;
  // consecutive node ids for the internal children, consecutive record positions for the leaf primitives
  int n_int = 0, n_leaf_prims = 0;
  for (int m = 0; m < k; m++) {
    if (ch[m] < 0 || bin_count(B, ch[m]) <= LEAF_MAX) n_leaf_prims += bin_count(B, ch[m]);
    else n_int++;
  }
  int base = 0;
  if (n_int) {
    base = atomicAdd(counter, n_int);
    if (base + n_int > capacity) { *err = 2; return; }
  }
  unsigned pbase = 0;
  if (n_leaf_prims) pbase = atomicAdd(prim_counter, (unsigned)n_leaf_prims);
  WideNode nd;
  nd.ox = ulo[0]; nd.oy = ulo[1]; nd.oz = ulo[2];
  nd.imask = 0;
  nd.child_base = (unsigned)base;
  nd.prim_base = pbase;
  for (int sl = 0; sl < 8; sl++) {
    nd.meta[sl] = 0;
    nd.qlox[sl] = nd.qloy[sl] = nd.qloz[sl] = 255;  // empty slot: inverted box
    nd.qhix[sl] = nd.qhiy[sl] = nd.qhiz[sl] = 0;
  }
  // quantisation frame per axis; verified against the exact expression the traversal uses
  int eb[3];
  for (int a = 0; a < 3; a++) {
    eb[a] = quant_exp(uhi[a] - ulo[a]);
    for (int attempt = 0; attempt < 4; attempt++) {
      const float scale = __uint_as_float((unsigned)eb[a] << 23);
      bool ok = true;
      for (int m = 0; m < k && ok; m++) {
        const float chh = a == 0 ? hi[m].x : a == 1 ? hi[m].y : hi[m].z;
        if (ceilf((chh - ulo[a]) / scale) > 255.0f) ok = false;
      }
      if (ok) break;
      eb[a] = min(eb[a] + 1, 254);
    }
  }
  nd.ex = (unsigned char)eb[0]; nd.ey = (unsigned char)eb[1]; nd.ez = (unsigned char)eb[2];
  int next_int = 0, poff = 0;
  for (int sl = 0; sl < 8; sl++) {
    const int m = child_in_slot[sl];
    if (m < 0) continue;
    unsigned char *ql[3] = {nd.qlox, nd.qloy, nd.qloz}, *qh[3] = {nd.qhix, nd.qhiy, nd.qhiz};
    for (int a = 0; a < 3; a++) {
      const float scale = __uint_as_float((unsigned)eb[a] << 23);
      const float cl = a == 0 ? lo[m].x : a == 1 ? lo[m].y : lo[m].z;
      const float chh = a == 0 ? hi[m].x : a == 1 ? hi[m].y : hi[m].z;
      float q0 = floorf((cl - ulo[a]) / scale), q1 = ceilf((chh - ulo[a]) / scale);
      q0 = fminf(fmaxf(q0, 0.f), 255.f);
      q1 = fminf(fmaxf(q1, 0.f), 255.f);
      while (q0 > 0.f && __fmaf_rn(q0, scale, ulo[a]) > cl) q0 -= 1.f;
      while (q1 < 255.f && __fmaf_rn(q1, scale, ulo[a]) < chh) q1 += 1.f;
      ql[a][sl] = (unsigned char)q0;
      qh[a][sl] = (unsigned char)q1;
    }
    if (ch[m] < 0 || bin_count(B, ch[m]) <= LEAF_MAX) {
      const int cnt = bin_count(B, ch[m]), first = bin_first(B, ch[m]);
      nd.meta[sl] = (unsigned char)((((1u << cnt) - 1u) << 5) | (unsigned)poff);  // count in unary | record offset
      for (int j = 0; j < cnt; j++) dest_of[first + j] = pbase + (unsigned)(poff + j);
      poff += cnt;
    } else {
      const int id = base + next_int++;
      nd.meta[sl] = (unsigned char)(0x20u | (24u + (unsigned)sl));
      nd.imask |= (unsigned char)(1u << sl);
      wide_bin[id] = ch[m];
    }
  }
  nodes[wid] = nd;
}

__global__ void __launch_bounds__(256)
    count_internal_kernel(int n, const int2 *__restrict__ range, unsigned long long *__restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool big = i < n - 1 && (range[i].y - range[i].x + 1) > LEAF_MAX;
  const unsigned b = __ballot_sync(0xffffffffu, big);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

__global__ void __launch_bounds__(256)
    emit_prims_kernel(const __grid_constant__ BuildGeoms B, long long N, const unsigned *__restrict__ vals,
                      const unsigned *__restrict__ dest_of, PrimRec *__restrict__ out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const long long p = vals[s];
  const int gi = find_geom(B, p);
  const GeomBuildInput &g = B.g[gi];
  const long long i = p - B.offset[gi];
  PrimRec r;
  if (g.kind == 0) {
    const float *a = g.verts + 3 * (size_t)g.idx[3 * i], *b = g.verts + 3 * (size_t)g.idx[3 * i + 1], *c = g.verts + 3 * (size_t)g.idx[3 * i + 2];
    // embree/kernels/geometry/triangle.h:52-53: e1 = v0 - v1, e2 = v2 - v0
    r.a = make_float4(a[0], a[1], a[2], a[0] - b[0]);
    r.b = make_float4(a[1] - b[1], a[2] - b[2], c[0] - a[0], c[1] - a[1]);
    r.c = make_float4(c[2] - a[2], __uint_as_float((unsigned)g.geom_id), __uint_as_float((unsigned)i), 0.f);
  } else if (g.kind == 1) {
    r.a = make_float4(g.centers[3 * i], g.centers[3 * i + 1], g.centers[3 * i + 2], sphere_radius(g, i));
    r.b = make_float4(g.epsilon, 0.f, 0.f, 0.f);
    r.c = make_float4(0.f, __uint_as_float((unsigned)g.geom_id | (1u << 24)), __uint_as_float((unsigned)i), 0.f);
  } else {
    const unsigned long long cp = (unsigned long long)(g.centers + 16 * (size_t)i);   // 64-byte aligned (cudaMalloc + 64 i)
    r.a = make_float4(__uint_as_float((unsigned)cp), __uint_as_float((unsigned)(cp >> 32)), gxc::curve_bound_radius(g.centers + 16 * (size_t)i), 0.f);
    r.b = make_float4(0.f, 0.f, 0.f, 0.f);
    r.c = make_float4(0.f, __uint_as_float((unsigned)g.geom_id | (2u << 24)), __uint_as_float((unsigned)i), 0.f);
  }
  out[dest_of[s]] = r;
}

// The build allocates and releases its multi-GB buffers one by one (peak memory stays near the result's size); the host time
// the driver spends in those calls lies inside the build's event pair, so it is accounted for separately (BvhResult::alloc_host_ms).
static thread_local double t_alloc_ms = 0.0;
struct AllocClock {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~AllocClock() { t_alloc_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
template <typename T>
static cudaError_t timed_malloc(T **p, size_t bytes) {
  AllocClock c;
  return cudaMalloc(p, bytes);
}
static void timed_free(void *p) {
  AllocClock c;
  cudaFree(p);
}
template <typename T>
struct DevBuf {
  T *p = nullptr;
  ~DevBuf() { if (p) timed_free(p); }
  int alloc(size_t n) { return timed_malloc(&p, sizeof(T) * (n ? n : 1)) == cudaSuccess ? 0 : 1; }
};

// bump allocator over one device allocation (sizing pass with base == nullptr, then the real one)
struct BuildArena {
  char *base = nullptr;
  size_t off = 0;
  ~BuildArena() { if (base) timed_free(base); }
  static size_t up(size_t b) { return (b + 255) & ~(size_t)255; }
  template <typename T>
  T *take(size_t count) {
    T *q = reinterpret_cast<T *>(reinterpret_cast<uintptr_t>(base) + off);  // base == nullptr: the sizing pass
    off += up(sizeof(T) * (count ? count : 1));
    return q;
  }
};

int build_bvh(const GeomBuildInput *geoms, int n_geoms, BvhResult *out, cudaStream_t st) {
  *out = BvhResult();
  if (n_geoms > MAX_BUILD_GEOMS) { gxy_set_error("too many geometry operators (%d > %d)", n_geoms, MAX_BUILD_GEOMS); return 1; }
  BuildGeoms B;
  B.n = n_geoms;
  long long N = 0;
  for (int k = 0; k < n_geoms; k++) { B.g[k] = geoms[k]; B.offset[k] = N; N += geoms[k].n_prims; }
  B.offset[n_geoms] = N;
  if (N == 0) return 0;
  if (N >= (1ll << 28)) { gxy_set_error("too many primitives for one partition (%lld >= 2^28)", N); return 1; }
  cudaEvent_t e0, e1;
  GXY_CUDA(cudaEventCreate(&e0));
  GXY_CUDA(cudaEventCreate(&e1));
  GXY_CUDA(cudaEventRecord(e0, st));
  t_alloc_ms = 0.0;
  const int n = (int)N;
  const unsigned gridN = (unsigned)((N + 255) / 256);

  // One arena for all the scratch whose size follows from N alone (bump allocation, released once at the end): the driver's time for
  // ~25 separate multi-GB cudaMalloc / cudaFree calls inside the build varied between 50 and 500 ms from run to run against ~40 ms
  // of kernels (profiles/r02_m_bvh_build_probe_before.txt).  nlo/nhi (node boxes, first written by the refit) take the place of
  // lo/hi (primitive boxes in input order, dead once gathered into sorted order).
  size_t tmp_bytes = 0;
  {
    cub::DoubleBuffer<unsigned long long> qk(nullptr, nullptr);
    cub::DoubleBuffer<unsigned> qv(nullptr, nullptr);
    GXY_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, qk, qv, n, 0, 63, st));
  }
  BuildArena arena;
  float4 *lo_p = nullptr, *hi_p = nullptr, *slo_p = nullptr, *shi_p = nullptr;
  unsigned *bounds_p = nullptr, *vals_p = nullptr, *vals2_p = nullptr, *dest_of_p = nullptr, *prim_counter_p = nullptr;
  unsigned long long *keys_p = nullptr, *keys2_p = nullptr, *cnt_p = nullptr;
  unsigned char *tmp_p = nullptr;
  int2 *child_p = nullptr, *range_p = nullptr;
  int *parent_int_p = nullptr, *parent_leaf_p = nullptr, *flags_p = nullptr, *counter_p = nullptr, *err_p = nullptr;
  for (int pass = 0; pass < 2; pass++) {
    arena.off = 0;
    lo_p = arena.take<float4>(N); hi_p = arena.take<float4>(N);
    slo_p = arena.take<float4>(N); shi_p = arena.take<float4>(N);
    bounds_p = arena.take<unsigned>(6);
    keys_p = arena.take<unsigned long long>(N); keys2_p = arena.take<unsigned long long>(N);
    vals_p = arena.take<unsigned>(N); vals2_p = arena.take<unsigned>(N);
    tmp_p = arena.take<unsigned char>(tmp_bytes);
    dest_of_p = arena.take<unsigned>(N); prim_counter_p = arena.take<unsigned>(1);
    child_p = arena.take<int2>(N); range_p = arena.take<int2>(N);
    parent_int_p = arena.take<int>(N); parent_leaf_p = arena.take<int>(N); flags_p = arena.take<int>(N);
    cnt_p = arena.take<unsigned long long>(1); counter_p = arena.take<int>(1); err_p = arena.take<int>(1);
    if (pass == 0 && timed_malloc(&arena.base, arena.off) != cudaSuccess) {
      cudaGetLastError();
      gxy_set_error("BVH build: out of device memory (%zu bytes of scratch for %lld primitives)", arena.off, N);
      return 1;
    }
  }
  struct { float4 *p; } lo{lo_p}, hi{hi_p}, slo{slo_p}, shi{shi_p}, nlo{lo_p}, nhi{hi_p};
  struct { unsigned *p; } bounds{bounds_p}, vals{vals_p}, vals2{vals2_p}, dest_of{dest_of_p}, prim_counter{prim_counter_p};
  struct { unsigned long long *p; } keys{keys_p}, keys2{keys2_p}, cnt{cnt_p};
  struct { int2 *p; } child{child_p}, range{range_p};
  struct { int *p; } parent_int{parent_int_p}, parent_leaf{parent_leaf_p}, flags{flags_p}, counter{counter_p}, err{err_p};
  DevBuf<int> wide_bin;
  const unsigned init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  GXY_CUDA(cudaMemcpyAsync(bounds.p, init_bounds, sizeof init_bounds, cudaMemcpyHostToDevice, st));
  prim_bounds_kernel<<<gridN, 256, 0, st>>>(B, N, lo.p, hi.p, bounds.p);
  morton_kernel<<<gridN, 256, 0, st>>>(N, lo.p, hi.p, bounds.p, keys.p, vals.p);
  GXY_CUDA(cudaGetLastError());
  {
    cub::DoubleBuffer<unsigned long long> dk(keys.p, keys2.p);
    cub::DoubleBuffer<unsigned> dv(vals.p, vals2.p);
    GXY_CUDA(cub::DeviceRadixSort::SortPairs(tmp_p, tmp_bytes, dk, dv, n, 0, 63, st));
    GXY_CUDA(cudaStreamSynchronize(st));
    if (dk.Current() != keys.p) std::swap(keys.p, keys2.p);
    if (dv.Current() != vals.p) std::swap(vals.p, vals2.p);
  }
  gather_boxes_kernel<<<gridN, 256, 0, st>>>(N, vals.p, lo.p, hi.p, slo.p, shi.p);
  GXY_CUDA(cudaStreamSynchronize(st));

  PrimRec *prims = nullptr;
  GXY_CUDA(timed_malloc(&prims, sizeof(PrimRec) * N));
  // dest_of: sorted position -> record position (node order)
  GXY_CUDA(cudaMemsetAsync(prim_counter.p, 0, sizeof(unsigned), st));

  WideNode *nodes = nullptr;
  long long n_nodes = 0;
  int depth = 0;
  if (n <= LEAF_MAX) {
    // a single leaf under one wide node
    WideNode nd;
    memset(&nd, 0, sizeof nd);
    std::vector<float4> hl(n), hh(n);
    GXY_CUDA(cudaMemcpy(hl.data(), slo.p, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    GXY_CUDA(cudaMemcpy(hh.data(), shi.p, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    float ulo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, uhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int m = 0; m < n; m++) {
      const float l3[3] = {hl[m].x, hl[m].y, hl[m].z}, h3[3] = {hh[m].x, hh[m].y, hh[m].z};
      for (int a = 0; a < 3; a++) { ulo[a] = fminf(ulo[a], l3[a]); uhi[a] = fmaxf(uhi[a], h3[a]); }
    }
    nd.ox = ulo[0]; nd.oy = ulo[1]; nd.oz = ulo[2];
    unsigned char *ex[3] = {&nd.ex, &nd.ey, &nd.ez};
    for (int a = 0; a < 3; a++) {
      int e = 0;
      const float ext = uhi[a] - ulo[a];
      if (ext > 0.f) frexpf(ext / 255.0f, &e); else e = -126;
      *ex[a] = (unsigned char)std::max(1, std::min(e + 1 + 127, 254));  // one extra octave of slack
    }
    nd.imask = 0; nd.child_base = 0; nd.prim_base = 0;
    for (int s = 0; s < 8; s++) { nd.meta[s] = 0; nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255; nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0; }
    nd.qlox[0] = nd.qloy[0] = nd.qloz[0] = 0;
    nd.qhix[0] = nd.qhiy[0] = nd.qhiz[0] = 255;
    nd.meta[0] = (unsigned char)((((1u << n) - 1u) << 5) | 0u);
    const unsigned ident[LEAF_MAX] = {0u, 1u, 2u};
    GXY_CUDA(cudaMemcpyAsync(dest_of.p, ident, sizeof(unsigned) * n, cudaMemcpyHostToDevice, st));
    GXY_CUDA(cudaStreamSynchronize(st));
    GXY_CUDA(timed_malloc(&nodes, sizeof(WideNode)));
    GXY_CUDA(cudaMemcpy(nodes, &nd, sizeof nd, cudaMemcpyHostToDevice));
    n_nodes = 1;
    depth = 1;
  } else {
    karras_kernel<<<(n - 1 + 255) / 256, 256, 0, st>>>(n, keys.p, child.p, range.p, parent_int.p, parent_leaf.p);
    GXY_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int) * n, st));
    refit_kernel<<<gridN, 256, 0, st>>>(n, child.p, parent_int.p, parent_leaf.p, slo.p, shi.p, nlo.p, nhi.p, flags.p);
    GXY_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
    count_internal_kernel<<<gridN, 256, 0, st>>>(n, range.p, cnt.p);
    unsigned long long h_cnt = 0;
    GXY_CUDA(cudaMemcpyAsync(&h_cnt, cnt.p, sizeof h_cnt, cudaMemcpyDeviceToHost, st));
    GXY_CUDA(cudaStreamSynchronize(st));
    const int capacity = (int)h_cnt + 1;
    WideNode *tmp_nodes = nullptr;
    GXY_CUDA(timed_malloc(&tmp_nodes, sizeof(WideNode) * (size_t)capacity));
    if (wide_bin.alloc(capacity)) { gxy_set_error("BVH build: out of device memory"); return 1; }
    const int zero = 0, one = 1;
    GXY_CUDA(cudaMemcpyAsync(wide_bin.p, &zero, sizeof(int), cudaMemcpyHostToDevice, st));  // wide 0 <- binary root 0
    GXY_CUDA(cudaMemcpyAsync(counter.p, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    GXY_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), st));
    BinView bv;
    bv.n = n; bv.child = child.p; bv.range = range.p; bv.slo = slo.p; bv.shi = shi.p; bv.nlo = nlo.p; bv.nhi = nhi.p;
    int begin = 0, end = 1;
    while (begin < end) {
      collapse_kernel<<<(end - begin + 127) / 128, 128, 0, st>>>(bv, begin, end, wide_bin.p, tmp_nodes, counter.p, capacity, prim_counter.p,
                                                                 dest_of.p, err.p);
      int h_counter = 0;
      GXY_CUDA(cudaMemcpyAsync(&h_counter, counter.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      GXY_CUDA(cudaStreamSynchronize(st));
      begin = end;
      end = h_counter;
      depth++;
      if (depth > 4096) break;
    }
    int h_err = 0;
    GXY_CUDA(cudaMemcpy(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (h_err) { timed_free(tmp_nodes); timed_free(prims); gxy_set_error("BVH build: wide-node capacity exceeded"); return 1; }
    n_nodes = end;
    GXY_CUDA(timed_malloc(&nodes, sizeof(WideNode) * (size_t)n_nodes));
    GXY_CUDA(cudaMemcpyAsync(nodes, tmp_nodes, sizeof(WideNode) * (size_t)n_nodes, cudaMemcpyDeviceToDevice, st));
    GXY_CUDA(cudaStreamSynchronize(st));
    timed_free(tmp_nodes);
  }
  emit_prims_kernel<<<gridN, 256, 0, st>>>(B, N, vals.p, dest_of.p, prims);
  GXY_CUDA(cudaGetLastError());
  GXY_CUDA(cudaEventRecord(e1, st));
  GXY_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  out->nodes = nodes; out->prims = prims; out->n_nodes = n_nodes; out->n_prims = N; out->max_depth = depth; out->build_ms = ms; out->alloc_host_ms = (float)t_alloc_ms;
  return 0;
}

}  // namespace gxy
