// gxy_common.cuh -- shared device/host structures and exact-order fp32 helpers.
//
// fp convention of the whole library (DESIGN.md "numerics"): the translation units are compiled
// with -fmad=false, IEEE divide/sqrt, no fast-math.  Every expression on the ray/shading path is
// written in the association order of the reference source it replaces; __fmaf_rn appears only
// (a) where Embree's AVX2 triangle test itself uses FMA and (b) in conservative
// BVH box tests, which never decide a result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "../../include/gxy_gpu.h"

// RayFlags.h:28-36, Renderer.cpp:78-82
#define RAY_PRIMARY 1
#define RAY_SHADOW 2
#define RAY_AO 4
#define RAY_SURFACE 1
#define RAY_OPAQUE 2
#define RAY_BOUNDARY 4
#define RAY_TIMEOUT 8
#define CLS_TERMINATED (-1)
#define CLS_DROP_ON_FLOOR (-2)
#define CLS_KEEP_HERE (-3)
#define CLS_UNDETERMINED (-4)

#define GXY_SM_COUNT 148

namespace gxy {

// ---- RayList on the device: the reference's 25-column SoA (Rays.ih:20-47) --------------------
struct Rays {
  float *ox, *oy, *oz, *dx, *dy, *dz, *nx, *ny, *nz, *sample, *r, *g, *b, *o, *sr, *sg, *sb, *so, *t, *tMax;
  int *x, *y, *type, *term, *classification;
};
__host__ __device__ inline Rays rays_view(float *base, size_t aligned) {
  Rays v;
  float **f = &v.ox;
  for (int k = 0; k < 20; k++) f[k] = base + (size_t)k * aligned;
  int **ip = &v.x;
  for (int k = 0; k < 5; k++) ip[k] = (int *)(base + (size_t)(20 + k) * aligned);
  return v;
}

// ---- transfer function: 256 x (r,g,b,opacity) -------------------------------------------------
struct DevTF {
  float4 e[256];
  float lo, hi, pad0, pad1;
};

struct DevVolume {
  int dims[3];
  int type;  // 0 float, 1 uchar
  float3 origin, spacing, rcp, upper;
  const void *vox;
  float samplingStep, samplingRate;
  int tf;  // the volume object's TF (last Vis added on this volume, MappedVis.cpp:206-212)
  int pad;
  unsigned long long nx, nxy;
};

struct DevVolVis {
  int n_slices, n_iso, volume_render, tf;
  float4 slices[GXY_MAX_SLICES];
  float iso[GXY_MAX_ISOVALUES];
  DevVolume vol;
};

// geometry operator table entry (post-intersect data)
struct DevGeom {
  int kind;  // 0 triangles, 1 spheres, 2 round Bezier curves (PathLines)
  int tf;
  const int *idx;        // triangles: int3 per prim
  const float *normals;  // float3 per vertex or NULL
  const float *data;     // per vertex (triangles) / per particle (spheres) or NULL
  const float *centers;  // spheres: float3 per particle; curves: 4 control points (x,y,z,r) = 16 floats per segment
  float radius0, radius1, value0, value1, epsilon;
  int pad;
};

// 8-wide quantised BVH node, 80 bytes (5 x 16 B loads), after Ylitie/Karras/Laine 2017 ("compressed
// wide BVH"): children of a node are stored contiguously (internal children from child_base in slot
// order, primitive records of its leaf children from prim_base), so a child is addressed by a
// bit position instead of a 32-bit reference and the traversal stack holds one entry per NODE.
//   origin (3 f32) | ex ey ez imask (4 u8) | child_base prim_base (2 u32) | meta (8 u8) | qlo/qhi per axis (6 x 8 u8)
// imask: bit s set = slot s holds an internal node.  meta[s]: 0 = empty slot;
//   internal: 0b001_11sss (low 5 bits = 24 + s); leaf: high 3 bits = primitive count in unary
//   (001, 011, 111), low 5 bits = offset of its first record from prim_base (<= 21).
// Slots are filled in octant order (bit0 = x-high, bit1 = y-high, bit2 = z-high half).
struct __align__(16) WideNode {
  float ox, oy, oz;
  unsigned char ex, ey, ez, imask;
  unsigned int child_base, prim_base;
  unsigned char meta[8];
  unsigned char qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

// primitive record in leaf order, 48 bytes (3 x 16 B loads)
//   triangle: a=(v0.xyz, e1.x) b=(e1.yz, e2.xy) c=(e2.z, bits(geom|kind<<24), bits(prim), 0)
//   sphere:   a=(c.xyz, radius) b=(epsilon,0,0,0) c=(0, bits(geom|1<<24), bits(prim), 0)
//   curve:    a=(address of its 4 control points: low, high word; bound radius for the cull, gxy_curve.cuh; 0) c=(0, bits(geom|2<<24), bits(prim), 0)
struct __align__(16) PrimRec {
  float4 a, b, c;
};

struct SceneParams {
  float3 gmin, gmax, lmin, lmax;
  int neighbors[6];
  int n_volvis, n_geoms;
  int integrate;  // any volume_render || isovalues (TraceRays.ispc:348-352)
  float step;     // min samplingStep*samplingRate (TraceRays.ispc:354-360)
  int n_curves;   // round Bezier segments (PathLines) among the primitives; > 0 selects the CURVES kernels
  int pad_curves; // (these two words fill what was alignment padding in front of vv: no offset changes)
  DevVolVis vv[GXY_MAX_VOLUME_VIS];
  const DevTF *tfs;
  const DevGeom *geoms;
  const WideNode *nodes;
  const PrimRec *prims;
  long long n_prims;
  int *error_flag;  // set to 1 by a kernel on traversal-stack overflow
  unsigned *work_counter;  // ray queue head of the persistent trace kernel (zeroed before each launch)
  unsigned long long *trav_counters;  // [0] nodes visited [1] primitives tested (only with -DGXY_TRAV_COUNTERS)
};

struct DevLights {
  int n_lights, n_ao, shadows, pad;
  float ao_radius, Ka, Kd, pad1;
  float lights[GXY_MAX_LIGHTS][3];
  int types[GXY_MAX_LIGHTS];
};

// camera frame computed on the host in double/float exactly as Camera.cpp:528-582 does
struct DevCamera {
  float3 veye, vdir, vr, vu, center;
  float scaling, off_x, off_y;
  int ortho;
};

// ---- exact-order math ------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 neg3(float3 a) { return f3(-a.x, -a.y, -a.z); }
// ospray/math/vec.ih dot()
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// vec.ih:575-576 / 581-582
__device__ __forceinline__ float3 normalize_isp(float3 v) { return v * (1.f / sqrtf(dot3(v, v))); }
__device__ __forceinline__ float3 safe_normalize(float3 v) { return v * (1.f / sqrtf(fmaxf(FLT_MIN, dot3(v, v)))); }
// src/data/dtypes.h normalize(vec3f&)
__device__ __forceinline__ void normalize_gxy(float3 &a) {
  float d = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  if (d != 0) {
    d = 1.0f / d;
    a.x *= d;
    a.y *= d;
    a.z *= d;
  }
}
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }

// LinearTransferFunction.ispc:19-97 on the packed 256 x float4 table
__device__ __forceinline__ float3 tf_color(const DevTF *__restrict__ tf, float value) {
  if (isnan(value)) return f3(0.f, 0.f, 0.f);
  float lo = tf->lo, hi = tf->hi;
  if (value <= lo) {
    float4 c = __ldg(&tf->e[0]);
    return f3(c.x, c.y, c.z);
  }
  if (value >= hi) {
    float4 c = __ldg(&tf->e[255]);
    return f3(c.x, c.y, c.z);
  }
  value = (value - lo) / (hi - lo) * 255.0f;
  int index = (int)floorf(value);
  float rem = value - (float)index;
  float4 a = __ldg(&tf->e[index]), b = __ldg(&tf->e[min(index + 1, 255)]);
  float om = 1.0f - rem;
  return f3(om * a.x + rem * b.x, om * a.y + rem * b.y, om * a.z + rem * b.z);
}
__device__ __forceinline__ float tf_opacity(const DevTF *__restrict__ tf, float value) {
  if (isnan(value)) return 0.0f;
  float lo = tf->lo, hi = tf->hi;
  if (value <= lo) return __ldg(&tf->e[0]).w;
  if (value >= hi) return __ldg(&tf->e[255]).w;
  float remapped = (value - lo) / (hi - lo) * 255.0f;
  int index = (int)floorf(remapped);
  float rem = remapped - (float)index;
  return (1.0f - rem) * __ldg(&tf->e[index]).w + rem * __ldg(&tf->e[min(index + 1, 255)]).w;
}
// colour and opacity of the same value in one lookup (DVR branch reads both)
__device__ __forceinline__ float4 tf_both(const DevTF *__restrict__ tf, float value) {
  if (isnan(value)) return make_float4(0.f, 0.f, 0.f, 0.f);
  float lo = tf->lo, hi = tf->hi;
  if (value <= lo) return __ldg(&tf->e[0]);
  if (value >= hi) return __ldg(&tf->e[255]);
  float remapped = (value - lo) / (hi - lo) * 255.0f;
  int index = (int)floorf(remapped);
  float rem = remapped - (float)index;
  float4 a = __ldg(&tf->e[index]), b = __ldg(&tf->e[min(index + 1, 255)]);
  float om = 1.0f - rem;
  return make_float4(om * a.x + rem * b.x, om * a.y + rem * b.y, om * a.z + rem * b.z, om * a.w + rem * b.w);
}

// SharedStructuredVolume.ispc:131-191, StructuredVolume.ispc:208-211
template <int TYPE>
__device__ __forceinline__ float voxel_at(const void *__restrict__ vox, unsigned long long o) {
  if (TYPE == 0) return __ldg((const float *)vox + o);
  return (float)__ldg((const unsigned char *)vox + o);
}
__device__ __forceinline__ float vol_sample(const DevVolume &v, float3 p) {
  float lx = v.rcp.x * (p.x - v.origin.x), ly = v.rcp.y * (p.y - v.origin.y), lz = v.rcp.z * (p.z - v.origin.z);
  float cx = fmaxf(0.0f, fminf(lx, v.upper.x)), cy = fmaxf(0.0f, fminf(ly, v.upper.y)), cz = fmaxf(0.0f, fminf(lz, v.upper.z));
  int ix = (int)cx, iy = (int)cy, iz = (int)cz;
  float fx = cx - (float)ix, fy = cy - (float)iy, fz = cz - (float)iz;
  unsigned long long o = (unsigned long long)ix + (unsigned long long)iy * v.nx + (unsigned long long)iz * v.nxy;
  float v000, v001, v010, v011, v100, v101, v110, v111;
  if (v.type == 0) {
    const float *__restrict__ d = (const float *)v.vox + o;
    v000 = __ldg(d); v001 = __ldg(d + 1);
    v010 = __ldg(d + v.nx); v011 = __ldg(d + v.nx + 1);
    v100 = __ldg(d + v.nxy); v101 = __ldg(d + v.nxy + 1);
    v110 = __ldg(d + v.nxy + v.nx); v111 = __ldg(d + v.nxy + v.nx + 1);
  } else {
    const unsigned char *__restrict__ d = (const unsigned char *)v.vox + o;
    v000 = (float)__ldg(d); v001 = (float)__ldg(d + 1);
    v010 = (float)__ldg(d + v.nx); v011 = (float)__ldg(d + v.nx + 1);
    v100 = (float)__ldg(d + v.nxy); v101 = (float)__ldg(d + v.nxy + 1);
    v110 = (float)__ldg(d + v.nxy + v.nx); v111 = (float)__ldg(d + v.nxy + v.nx + 1);
  }
  const float v00 = v000 + fx * (v001 - v000);
  const float v01 = v010 + fx * (v011 - v010);
  const float v10 = v100 + fx * (v101 - v100);
  const float v11 = v110 + fx * (v111 - v110);
  const float v0 = v00 + fy * (v01 - v00);
  const float v1 = v10 + fy * (v11 - v10);
  return v0 + fz * (v1 - v0);
}
// StructuredVolume.ispc:70-113
__device__ __forceinline__ float3 vol_gradient(const DevVolume &v, float3 p) {
  float s = vol_sample(v, p);
  float gx = vol_sample(v, p + f3(v.spacing.x, 0.f, 0.f)) - s;
  float gy = vol_sample(v, p + f3(0.f, v.spacing.y, 0.f)) - s;
  float gz = vol_sample(v, p + f3(0.f, 0.f, v.spacing.z)) - s;
  return f3(gx / v.spacing.x, gy / v.spacing.y, gz / v.spacing.z);
}

}  // namespace gxy

// error plumbing shared by the translation units
void gxy_set_error(const char *fmt, ...);
#define GXY_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      gxy_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
