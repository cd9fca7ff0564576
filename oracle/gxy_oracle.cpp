/*
 * gxy_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code; see gxy_oracle.h).
 *
 * A plain C++ restatement of Galaxy's "trace a RayList against a Visualization" path.
 * Every function cites the reference file:line it follows (paths relative to the reference
 * tree).  Written independently of galaxy_b200/csrc; the CUDA path is checked against this.
 *
 * fp convention: compile with -ffp-contract=off; every expression below is evaluated in fp32
 * left-to-right exactly as written unless a double is written explicitly; fmaf() appears only
 * where Embree's AVX2 triangle test uses madd/msub.
 */
#include "gxy_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// RayFlags.h:28-36 and Renderer.cpp:78-82
enum { RAY_PRIMARY = 1, RAY_SHADOW = 2, RAY_AO = 4, RAY_EMPTY = 8 };
enum { RAY_SURFACE = 1, RAY_OPAQUE = 2, RAY_BOUNDARY = 4, RAY_TIMEOUT = 8 };
enum { TERMINATED = -1, DROP_ON_FLOOR = -2, KEEP_HERE = -3, UNDETERMINED = -4 };

struct V3 { float x, y, z; };
static inline V3 mk(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// ospray/math/vec.ih dot(): a.x*b.x + a.y*b.y + a.z*b.z
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// ospray/math/vec.ih:575-576
static inline V3 normalize_isp(V3 v) { return v * (1.f / sqrtf(dot(v, v))); }
// ospray/math/vec.ih:581-582
static inline V3 safe_normalize(V3 v) { return v * (1.f / sqrtf(std::max(FLT_MIN, dot(v, v)))); }
// src/data/dtypes.h normalize(vec3f&): d = len; if (d != 0) { d = 1.0/d; a *= d }
static inline void normalize_gxy(V3 &a) {
  float d = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
  if (d != 0) { d = (float)(1.0 / d); a.x *= d; a.y *= d; a.z *= d; }
}
static inline float min3(float a, float b, float c) { return std::min(std::min(a, b), c); }

// ---------------------------------------------------------------------------------------------
// Transfer function: ospray/transferFunction/LinearTransferFunction.ispc:19-97 (256 entries)
struct TF {
  float color[256][3];
  float opacity[256];
  float lo, hi;
};

static inline V3 tf_color(const TF &tf, float value) {
  if (std::isnan(value)) return mk(0.f, 0.f, 0.f);
  if (value <= tf.lo) return mk(tf.color[0][0], tf.color[0][1], tf.color[0][2]);
  if (value >= tf.hi) return mk(tf.color[255][0], tf.color[255][1], tf.color[255][2]);
  value = (value - tf.lo) / (tf.hi - tf.lo) * (256 - 1.0f);
  int index = (int)floorf(value);
  float remainder = value - index;
  int i1 = std::min(index + 1, 255);
  const float *a = tf.color[index], *b = tf.color[i1];
  return mk((1.0f - remainder) * a[0] + remainder * b[0], (1.0f - remainder) * a[1] + remainder * b[1],
            (1.0f - remainder) * a[2] + remainder * b[2]);
}

static inline float tf_opacity(const TF &tf, float value) {
  if (std::isnan(value)) return 0.0f;
  if (value <= tf.lo) return tf.opacity[0];
  if (value >= tf.hi) return tf.opacity[255];
  const float remapped = (value - tf.lo) / (tf.hi - tf.lo) * (256 - 1.0f);
  int index = (int)floorf(remapped);
  float remainder = remapped - index;
  return (1.0f - remainder) * tf.opacity[index] + remainder * tf.opacity[std::min(index + 1, 255)];
}

// ---------------------------------------------------------------------------------------------
// Volume: src/ospray/OsprayVolume.cpp:25-52 + ospray/volume/structured/**
struct VolumeData {
  int dims[3];
  V3 origin, spacing, rcp_spacing, upper;   // upper = nextafter(dims-1, 0), StructuredVolume.ispc:223
  int type;                                 // 0 float, 1 uchar
  const void *voxels;
  float samplingStep;                       // reduce_min(gridSpacing), StructuredVolume.ispc:250
  float samplingRate;                       // 1.0, OsprayVolume.cpp:45
  const TF *tf;                             // the volume object's TF (last committed Vis wins)
};

static inline float voxel(const VolumeData &v, size_t ofs) {
  return v.type == 0 ? ((const float *)v.voxels)[ofs] : (float)((const unsigned char *)v.voxels)[ofs];
}

// SharedStructuredVolume.ispc:131-191 (SSV_sample_*), transform StructuredVolume.ispc:208-211
static inline float vol_sample(const VolumeData &v, V3 p) {
  V3 lc = mk(v.rcp_spacing.x * (p.x - v.origin.x), v.rcp_spacing.y * (p.y - v.origin.y),
             v.rcp_spacing.z * (p.z - v.origin.z));
  // clamp(v, lo, hi) = max(lo, min(v, hi))  (ospray/math/math.ih:109)
  V3 c = mk(std::max(0.0f, std::min(lc.x, v.upper.x)), std::max(0.0f, std::min(lc.y, v.upper.y)),
            std::max(0.0f, std::min(lc.z, v.upper.z)));
  int ix = (int)c.x, iy = (int)c.y, iz = (int)c.z;
  float fx = c.x - (float)ix, fy = c.y - (float)iy, fz = c.z - (float)iz;
  size_t nx = v.dims[0], nxy = (size_t)v.dims[0] * v.dims[1];
  size_t o = ix + iy * nx + iz * nxy;
  const float v000 = voxel(v, o), v001 = voxel(v, o + 1);
  const float v00 = v000 + fx * (v001 - v000);
  const float v010 = voxel(v, o + nx), v011 = voxel(v, o + nx + 1);
  const float v01 = v010 + fx * (v011 - v010);
  const float v100 = voxel(v, o + nxy), v101 = voxel(v, o + nxy + 1);
  const float v10 = v100 + fx * (v101 - v100);
  const float v110 = voxel(v, o + nxy + nx), v111 = voxel(v, o + nxy + nx + 1);
  const float v11 = v110 + fx * (v111 - v110);
  const float v0 = v00 + fy * (v01 - v00);
  const float v1 = v10 + fy * (v11 - v10);
  return v0 + fz * (v1 - v0);
}

// StructuredVolume.ispc:70-113 (forward differences)
static inline V3 vol_gradient(const VolumeData &v, V3 p) {
  float s = vol_sample(v, p);
  V3 g;
  g.x = vol_sample(v, p + mk(v.spacing.x, 0.f, 0.f)) - s;
  g.y = vol_sample(v, p + mk(0.f, v.spacing.y, 0.f)) - s;
  g.z = vol_sample(v, p + mk(0.f, 0.f, v.spacing.z)) - s;
  return mk(g.x / v.spacing.x, g.y / v.spacing.y, g.z / v.spacing.z);
}

struct VolumeVisOp {          // src/renderer/VolumeVis.ih:25-37
  int dataset_id;
  VolumeData *vol;
  TF tf;
  std::vector<float> slices;  // 4 per plane
  std::vector<float> isovalues;
  bool volume_render;
};

// ---------------------------------------------------------------------------------------------
// Geometry
struct GeomOp {
  int kind;                   // 0 triangles, 1 spheres, 2 round Bezier curves (PathLines)
  TF tf;
  // triangles (src/ospray/OsprayTriangles.cpp:25-57)
  int nv, nt;
  const float *verts, *normals, *data;
  const int *idx;
  // spheres (src/ospray/OsprayParticles.cpp:34-40, ParticlesVis.cpp:136-144)
  int n;
  const float *centers;
  float radius0, radius1, value0, value1, epsilon;
  // curves: 4 control points (x,y,z,r) per segment, built by build_curves (owned by the scene)
  int ncurves;
  const float *cp;
};

struct Prim { int geom, prim; };
struct BNode { float lo[3], hi[3]; int left, right, first, count; };

struct Hit1 { int geomID, primID; float t, u, v; V3 Ng; };

static inline float sphere_radius(const GeomOp &g, int i) {
  // DataDrivenSpheres.ispc:65-88,100-112
  if (g.data && g.value0 != g.value1) {
    float dataval = g.data[i];
    float d = (dataval - g.value0) / (g.value1 - g.value0);
    if (d > 1) return g.radius1;
    else if (d < 0) return g.radius0;
    else return g.radius0 + d * (g.radius1 - g.radius0);
  }
  return g.radius0;
}

// Embree triangle test, AVX2 op order:
// embree/kernels/geometry/triangle_intersector_moeller.h:210-266, triangle.h:52-53,
// common/math/vec3.h:216,221 (dot/cross with madd/msub)
static inline float edot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
static inline V3 ecross(V3 a, V3 b) {
  return mk(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}

static inline bool tri_test(V3 v0, V3 v1, V3 v2, V3 org, V3 dir, float tnear, float tfar,
                            float &t, float &u, float &v, V3 &Ng) {
  const V3 e1 = v0 - v1, e2 = v2 - v0;
  Ng = ecross(e2, e1);
  const V3 C = v0 - org;
  const V3 R = ecross(C, dir);
  const float den = edot(Ng, dir);
  const float absDen = fabsf(den);
  const float sgn = den < 0.f || (den == 0.f && std::signbit(den)) ? -1.f : 1.f;
  const float U = edot(e2, R) * sgn;   // xor with sign mask == multiply by +-1 (exact)
  if (!(U >= 0.0f)) return false;
  const float V = edot(e1, R) * sgn;
  if (!(V >= 0.0f)) return false;
  const float W = absDen - U - V;
  if (!(W >= 0.0f)) return false;
  const float T = edot(Ng, C) * sgn;
  if (!((absDen * tnear < T) && (T <= absDen * tfar))) return false;
  if (!(den != 0.f)) return false;
  t = T / absDen;      // Embree: T*rcp(absDen) (rcp+Newton, ISA dependent) -- policy: IEEE divide
  u = U / absDen;
  v = V / absDen;
  return true;
}

// DataDrivenSpheres.ispc:90-155
static inline bool sphere_test(const GeomOp &g, int prim, V3 org, V3 dir, float t0, float tcur,
                               float &t, V3 &Ng) {
  V3 center = mk(g.centers[3 * prim], g.centers[3 * prim + 1], g.centers[3 * prim + 2]);
  float radius = sphere_radius(g, prim);
  const float approxDist = dot(center - org, dir);
  const V3 closeOrg = org + approxDist * dir;
  const V3 A = center - closeOrg;
  const float a = dot(dir, dir);
  const float b = 2.f * dot(dir, A);
  const float c = dot(A, A) - radius * radius;
  const float radical = b * b - 4.f * a * c;
  if (radical < 0.f) return false;
  const float srad = sqrtf(radical);
  const float t_in = (b - srad) * (1.f / (2.f * a)) + approxDist;
  const float t_out = (b + srad) * (1.f / (2.f * a)) + approxDist;
  bool hit = false;
  if (t_in > t0 && t_in < tcur) { hit = true; t = t_in; }
  else if (t_out > (t0 + g.epsilon) && t_out < tcur) { hit = true; t = t_out; }
  if (hit) Ng = org + t * dir - center;
  return hit;
}


// ---------------------------------------------------------------------------------------------
// PathLines: round cubic Bezier segments.  Galaxy turns every poly-line segment into 4 control points (x,y,z,r)
// (src/ospray/DataDrivenPathLines.cpp:103-156) and hands them to Embree as RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE
// (DataDrivenPathLines.ispc:319-324), which intersects them with SweepCurve1Intersector{1,K}<BezierCurve3fa>
// (embree/kernels/geometry/curve_intersector_virtual.cpp:256-266; curve_intersector_sweep.h:55-241).  The restatement
// below follows that file for the AVX/AVX2 build (VSIZEX = 8 lanes = 7 sub-segments per level, numBezierSubdivisions = 2)
// lane by lane; madd/msub are fmaf as in the AVX2 build, rcp/rsqrt (rcpss/rsqrtss + Newton in Embree) are IEEE.
struct V4 { float x, y, z, w; };
static inline V4 mk4(float x, float y, float z, float w) { V4 v = {x, y, z, w}; return v; }
static inline V3 xyz(V4 a) { return mk(a.x, a.y, a.z); }
static inline V4 operator+(V4 a, V4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline V4 operator-(V4 a, V4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline V4 operator*(float s, V4 a) { return mk4(s * a.x, s * a.y, s * a.z, s * a.w); }
static inline float e_rcp(float x) { return 1.0f / x; }
static inline float e_rsqrt(float x) { return 1.0f / sqrtf(x); }
static inline float dot_fa(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }   // Vec3fa dot = _mm_dp_ps(a,b,0x7F), vec3fa.h:289
static inline V4 lerp4(V4 a, V4 b, float t) {   // vec4.h:191 / vec3fa.h:341: madd(1-t, v0, t*v1)
  const float s = 1.0f - t;
  return mk4(fmaf(s, a.x, t * b.x), fmaf(s, a.y, t * b.y), fmaf(s, a.z, t * b.z), fmaf(s, a.w, t * b.w));
}
// CubicBezierCurve::eval(t,p,dp) / veval(t,p,dp) (bezier_curve.h:322-339, 476-492): de Casteljau
static inline void bezier_eval(const V4 cp[4], float t, V4 &p, V4 &dp) {
  const V4 p10 = lerp4(cp[0], cp[1], t), p11 = lerp4(cp[1], cp[2], t), p12 = lerp4(cp[2], cp[3], t);
  const V4 p20 = lerp4(p10, p11, t), p21 = lerp4(p11, p12, t);
  p = lerp4(p20, p21, t);
  dp = 3.0f * (p21 - p20);
}
// eval_dudu (bezier_curve.h:382-386) with BezierBasis::derivative2 (:52-61)
static inline V4 bezier_dudu(const V4 cp[4], float t1) {
  const float t0 = 1.0f - t1;
  const float b0 = 6.0f * t0, b1 = 6.0f * fmaf(-2.0f, t0, t1), b2 = 6.0f * fmaf(-2.0f, t1, t0), b3 = 6.0f * t1;
  V4 r;
  r.x = fmaf(b0, cp[0].x, fmaf(b1, cp[1].x, fmaf(b2, cp[2].x, b3 * cp[3].x)));
  r.y = fmaf(b0, cp[0].y, fmaf(b1, cp[1].y, fmaf(b2, cp[2].y, b3 * cp[3].y)));
  r.z = fmaf(b0, cp[0].z, fmaf(b1, cp[1].z, fmaf(b2, cp[2].z, b3 * cp[3].z)));
  r.w = fmaf(b0, cp[0].w, fmaf(b1, cp[1].w, fmaf(b2, cp[2].w, b3 * cp[3].w)));
  return r;
}

struct CurveRay { V3 dir; float tnear, tfar; float u; V3 Ng; };   // org is 0 after the shift by dt (:236-238)
static const float E_ULP = 1.1920929e-07f;   // embree ulp = numeric_limits<float>::epsilon()

// intersect_bezier_iterative_jacobian (curve_intersector_sweep.h:69-120); the epilog (Intersect1Epilog1 without filter,
// intersector_epilog.h:72-80) shortens the ray
static bool curve_jacobian(CurveRay &ray, float dt, const V4 cp[4], float u, float t) {
  const V3 dir = ray.dir;
  const float length_ray_dir = sqrtf(dot_fa(dir, dir));
  for (int i = 0; i < 5; i++) {
    const V3 Q = mk(fmaf(t, dir.x, 0.0f), fmaf(t, dir.y, 0.0f), fmaf(t, dir.z, 0.0f));
    V4 P4, dP4; bezier_eval(cp, u, P4, dP4);
    const V4 ddP4 = bezier_dudu(cp, u);
    const V3 P = xyz(P4), dPdu = xyz(dP4), ddPdu = xyz(ddP4);
    const V3 R = Q - P;
    const V3 dRdu = neg(dPdu);
    const V3 dRdt = dir;
    const V3 T = dPdu * e_rsqrt(dot_fa(dPdu, dPdu));
    const float pp = dot_fa(dPdu, dPdu), pdp = dot_fa(dPdu, ddPdu);   // dnormalize, vec3fa.h:321-326
    const V3 dTdu = ((pp * ddPdu - pdp * dPdu) * e_rcp(pp)) * e_rsqrt(pp);
    const float f = dot_fa(R, T);
    const float dfdu = dot_fa(dRdu, T) + dot_fa(R, dTdu);
    const float dfdt = dot_fa(dRdt, T);
    const float K = dot_fa(R, R) - f * f;
    const float dKdu = dot_fa(R, dRdu) - f * dfdu;
    const float dKdt = dot_fa(R, dRdt) - f * dfdt;
    const float rsqrt_K = e_rsqrt(K);
    const float g = sqrtf(K) - P4.w;
    const float dgdu = dKdu * rsqrt_K - dP4.w;
    const float dgdt = dKdt * rsqrt_K;
    // rcp(J)*Vec2f(f,g), J = LinearSpace2f(dfdu,dfdt,dgdu,dgdt): adjoint()/det() (linearspace2.h:49-55), then b.x*vx + b.y*vy
    const float det = dfdu * dgdt - dgdu * dfdt;
    const float ixx = dgdt / det, ixy = -dgdu / det, iyx = -dfdt / det, iyy = dfdu / det;   // vx = (ixx,ixy), vy = (iyx,iyy)
    const float du = f * ixx + g * iyx, dtt = f * ixy + g * iyy;
    u = u - du; t = t - dtt;
    const bool converged_u = fabsf(f) < 16.0f * E_ULP * std::max(std::max(fabsf(dPdu.x), fabsf(dPdu.y)), fabsf(dPdu.z));
    const bool converged_t = fabsf(g) < 16.0f * E_ULP * length_ray_dir;
    if (converged_u && converged_t) {
      t += dt;
      if (!(t > ray.tnear && t < ray.tfar)) return false;
      if (!(u >= 0.0f && u <= 1.0f)) return false;
      const V3 QP = Q - P;
      const V3 Rn = QP * e_rsqrt(dot_fa(QP, QP));
      const V3 U = mk(fmaf(dP4.w, Rn.x, dPdu.x), fmaf(dP4.w, Rn.y, dPdu.y), fmaf(dP4.w, Rn.z, dPdu.z));
      const V3 V = ecross(dPdu, Rn);
      ray.tfar = t; ray.u = u; ray.Ng = ecross(V, U);
      return true;
    }
  }
  return false;
}

static inline float vdot(V3 a, V3 b) { return edot(a, b); }   // Vec3<vfloat> dot, vec3.h:216

// CylinderN<8>::intersect, one lane (cylinder.h:162-225).  Returns the lane's valid bit; t0/t1 = +inf/-inf when invalid.
static inline bool cylinder_lane(V3 p0, V3 p1, float r, V3 dir, float &t_lo, float &t_up, float &u0, V3 &Ng0, float &u1, V3 &Ng1) {
  const float rr = r * r;
  const V3 d01 = p1 - p0;
  const float rl = e_rsqrt(vdot(d01, d01));
  const V3 dP = d01 * rl;
  const V3 O = mk(0.f, 0.f, 0.f) - p0, dO = dir;
  const float dOdO = vdot(dO, dO), OdO = vdot(dO, O), OO = vdot(O, O), dOz = vdot(dP, dO), Oz = vdot(dP, O);
  const float A = dOdO - dOz * dOz;
  const float B = 2.0f * (OdO - dOz * Oz);
  const float C = OO - Oz * Oz - rr;
  const float D = B * B - 4.0f * A * C;
  bool valid = D >= 0.0f;
  const float Q = sqrtf(D);
  const float rcp_2A = e_rcp(2.0f * A);
  const float t0 = (-B - Q) * rcp_2A, t1 = (-B + Q) * rcp_2A;
  u0 = fmaf(t0, dOz, Oz) * rl;
  { const V3 Pr = t0 * dir; const V3 Pl = mk(fmaf(u0, d01.x, p0.x), fmaf(u0, d01.y, p0.y), fmaf(u0, d01.z, p0.z)); Ng0 = Pr - Pl; }
  u1 = fmaf(t1, dOz, Oz) * rl;
  { const V3 Pr = t1 * dir; const V3 Pl = mk(fmaf(u1, d01.x, p0.x), fmaf(u1, d01.y, p0.y), fmaf(u1, d01.z, p0.z)); Ng1 = Pr - Pl; }
  t_lo = valid ? t0 : INFINITY;
  t_up = valid ? t1 : -INFINITY;
  const float eps = 16.0f * E_ULP * std::max(fabsf(dOdO), fabsf(dOz * dOz));
  if (valid && fabsf(A) < eps) {   // ray parallel to the cylinder
    const bool inside = C <= 0.0f;
    t_lo = inside ? -INFINITY : INFINITY;
    t_up = inside ? INFINITY : -INFINITY;
    valid = inside;
  }
  return valid;
}
// HalfPlaneN::intersect, one lane, ray origin 0 (plane.h:60-72)
static inline void halfplane_lane(V3 P, V3 N, V3 dir, float &lower, float &upper) {
  const V3 O = mk(0.f, 0.f, 0.f) - P;
  const float ON = vdot(O, N), DN = vdot(dir, N);
  const bool eps = fabsf(DN) < 1E-18f;
  const float t = -ON * e_rcp(DN);
  lower = (eps || DN < 0.0f) ? -INFINITY : t;
  upper = (eps || DN > 0.0f) ? INFINITY : t;
}
static inline float e_min(float a, float b) { return a < b ? a : b; }   // _mm_min_ps(a,b): a < b ? a : b (NaN -> b)
static inline float e_max(float a, float b) { return a > b ? a : b; }
// select_min (vfloat8_avx.h:669-674): lowest lane holding the minimum, else lowest valid lane
static inline int select_min_lane(unsigned valid, const float *v) {
  float m = INFINITY;
  for (int i = 0; i < 8; i++) if (valid >> i & 1) m = e_min(v[i], m);
  unsigned vm = 0;
  for (int i = 0; i < 8; i++) if ((valid >> i & 1) && v[i] == m) vm |= 1u << i;
  if (!vm) vm = valid;
  return __builtin_ctz(vm);
}

// intersect_bezier_recursive_jacobian (curve_intersector_sweep.h:122-221)
static bool curve_recursive(CurveRay &ray, float dt, const V4 cp[4], float u0, float u1, int depth) {
  const int maxDepth = 2;   // numBezierSubdivisions with __AVX__
  const V3 dir = ray.dir;
  const float dscale = (u1 - u0) * (1.0f / (3.0f * 7));
  float vu0[8]; V4 P0[8], dP0du[8];
  for (int i = 0; i < 8; i++) {
    vu0[i] = fmaf((float)i * (1.0f / 7), u1 - u0, u0);   // lerp(u0,u1,step*(1/7)) = madd(t,b-a,a), vfloat8_avx.h:477
    bezier_eval(cp, vu0[i], P0[i], dP0du[i]);
    dP0du[i] = mk4(dP0du[i].x * dscale, dP0du[i].y * dscale, dP0du[i].z * dscale, dP0du[i].w * dscale);
  }
  unsigned valid = 0, valid0 = 0, valid1 = 0, unstable0 = 0, unstable1 = 0;
  float tp0_lo[8], tp1_lo[8], tp1_up[8], uo0[8], uo1[8];
  for (int i = 0; i < 7; i++) {
    const V4 P3 = P0[i + 1], dP3du = dP0du[i + 1];
    const V4 P1 = P0[i] + dP0du[i], P2 = P3 - dP3du;
    const V3 p0 = xyz(P0[i]), p3 = xyz(P3), d0 = xyz(dP0du[i]), d3 = xyz(dP3du), chord = p3 - p0;
    // bounding cylinders (:147-157); sqr_point_to_line_distance(PmQ0,Q1mQ0), vec3.h:253-258
    const V3 n1 = ecross(d0, chord), n2 = ecross(d3, chord);
    const float rcd = e_rcp(vdot(chord, chord));
    const float rr1 = vdot(n1, n1) * rcd, rr2 = vdot(n2, n2) * rcd;
    const float maxr12 = sqrtf(e_max(rr1, rr2));
    const float one_plus_ulp = 1.0f + 2.0f * E_ULP, one_minus_ulp = 1.0f - 2.0f * E_ULP;
    float r_outer = e_max(e_max(P0[i].w, P1.w), e_max(P2.w, P3.w)) + maxr12;
    float r_inner = e_min(e_min(P0[i].w, P1.w), e_min(P2.w, P3.w)) - maxr12;
    r_outer = one_plus_ulp * r_outer;
    r_inner = e_max(0.0f, one_minus_ulp * r_inner);
    float to_lo, to_up, u_outer0, u_outer1; V3 Ngo0, Ngo1;
    bool v = cylinder_lane(p0, p3, r_outer, dir, to_lo, to_up, u_outer0, Ngo0, u_outer1, Ngo1);
    // cap planes (:165-172); intersect(BBox,BBox) = (max(lower), min(upper))
    float tp_lo = e_max(ray.tnear - dt, to_lo), tp_up = e_min(ray.tfar - dt, to_up);
    float hl, hu;
    halfplane_lane(p0, d0, dir, hl, hu);       tp_lo = e_max(tp_lo, hl); tp_up = e_min(tp_up, hu);
    halfplane_lane(p3, neg(d3), dir, hl, hu);  tp_lo = e_max(tp_lo, hl); tp_up = e_min(tp_up, hu);
    v = v && (tp_lo <= tp_up);
    // u of the outer hits (:176-179)
    u_outer0 = e_min(e_max(u_outer0, 0.0f), 1.0f);   // clamp(x,lo,hi) = min(max(x,lo),hi)
    u_outer1 = e_min(e_max(u_outer1, 0.0f), 1.0f);
    uo0[i] = fmaf(((float)i + u_outer0) * (1.0f / 8.0f), u1 - u0, u0);   // (step+u)*(1/float(VSIZEX)): Embree's own 1/8
    uo1[i] = fmaf(((float)i + u_outer1) * (1.0f / 8.0f), u1 - u0, u0);
    // inner cylinder (:181-188)
    float ti_lo, ti_up, ui0, ui1; V3 Ngi0, Ngi1;
    const bool valid_inner = cylinder_lane(p0, p3, r_inner, dir, ti_lo, ti_up, ui0, Ngi0, ui1, Ngi1);
    const V3 nd = dir * e_rsqrt(dot_fa(dir, dir));   // normalize(ray.dir) on Vec3fa, then broadcast
    const V3 nn0 = Ngi0 * e_rsqrt(vdot(Ngi0, Ngi0)), nn1 = Ngi1 * e_rsqrt(vdot(Ngi1, Ngi1));
    const bool un0 = !valid_inner || (fabsf(vdot(nd, nn0)) < 0.3f);
    const bool un1 = !valid_inner || (fabsf(vdot(nd, nn1)) < 0.3f);
    // subtract the inner interval (:190-195; bbox.h:166-172)
    tp0_lo[i] = tp_lo; const float tp0_up = e_min(tp_up, ti_lo);
    tp1_lo[i] = e_max(tp_lo, ti_up); tp1_up[i] = tp_up;
    if (v) valid |= 1u << i;
    if (v && tp0_lo[i] <= tp0_up) valid0 |= 1u << i;
    if (v && tp1_lo[i] <= tp1_up[i]) valid1 |= 1u << i;
    if (un0) unstable0 |= 1u << i;
    if (un1) unstable1 |= 1u << i;
  }
  if (!valid) return false;
  if (!(valid0 | valid1)) return false;
  bool found = false;
  while (valid0) {   // first hits front to back (:198-208)
    const int i = select_min_lane(valid0, tp0_lo); valid0 &= ~(1u << i);
    const int termDepth = (unstable0 >> i & 1) ? maxDepth + 1 : maxDepth;
    if (depth >= termDepth) found = curve_jacobian(ray, dt, cp, uo0[i], tp0_lo[i]) | found;
    else                    found = curve_recursive(ray, dt, cp, vu0[i], vu0[i + 1], depth + 1) | found;
    for (int k = 0; k < 7; k++) if (!(tp0_lo[k] + dt <= ray.tfar)) valid0 &= ~(1u << k);
  }
  for (int k = 0; k < 7; k++) if (!(tp1_lo[k] + dt <= ray.tfar)) valid1 &= ~(1u << k);
  while (valid1) {   // second hits front to back (:211-219)
    const int i = select_min_lane(valid1, tp1_lo); valid1 &= ~(1u << i);
    const int termDepth = (unstable1 >> i & 1) ? maxDepth + 1 : maxDepth;
    if (depth >= termDepth) found = curve_jacobian(ray, dt, cp, uo1[i], tp1_up[i]) | found;
    else                    found = curve_recursive(ray, dt, cp, vu0[i], vu0[i + 1], depth + 1) | found;
    for (int k = 0; k < 7; k++) if (!(tp1_lo[k] + dt <= ray.tfar)) valid1 &= ~(1u << k);
  }
  return found;
}

// SweepCurve1Intersector1::intersect (curve_intersector_sweep.h:224-241).  Nearest hit of ONE segment in (tnear, tfar).
static bool curve_test(const float *cp16, V3 org, V3 dir, float tnear, float tfar, float &t, float &u, V3 &Ng) {
  V4 cp[4];
  for (int k = 0; k < 4; k++) cp[k] = mk4(cp16[4 * k], cp16[4 * k + 1], cp16[4 * k + 2], cp16[4 * k + 3]);
  const V4 c4 = 0.25f * (((cp[0] + cp[1]) + cp[2]) + cp[3]);   // center()
  const float dt = dot_fa(xyz(c4) - org, dir) * e_rcp(dot_fa(dir, dir));
  const V3 ref = mk(fmaf(dt, dir.x, org.x), fmaf(dt, dir.y, org.y), fmaf(dt, dir.z, org.z));
  for (int k = 0; k < 4; k++) { cp[k].x -= ref.x; cp[k].y -= ref.y; cp[k].z -= ref.z; }   // curve0 - ref, ref.w = 0
  CurveRay ray; ray.dir = dir; ray.tnear = tnear; ray.tfar = tfar; ray.u = 0.f; ray.Ng = mk(0.f, 0.f, 0.f);
  if (!curve_recursive(ray, dt, cp, 0.0f, 1.0f, 1)) return false;
  t = ray.tfar; u = ray.u; Ng = ray.Ng;
  return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
namespace {
// DataDrivenPathLines::finalize (src/ospray/DataDrivenPathLines.cpp:28-37 MAP_RADIUS, :103-156): one cubic Bezier
// segment per poly-line segment; `connectivity[i]` is the index of the segment's first vertex (PathLines.cpp:110-122),
// consecutive segments of one line are joined with Catmull-Rom-like tangents, line ends repeat their end point.
static inline float map_radius(float d, float radius0, float radius1, float value0, float value1) {
  if (value0 == value1) return radius0;
  const float R = (d - value0) / (value1 - value0);
  // `R <= 0.0` / `R >= 1.0` compare in double, which is exact for a float
  return (R <= 0.0f) ? radius0 : (R >= 1.0f) ? radius1 : radius0 + R * (radius1 - radius0);
}
static inline float lerp_os(float factor, float a, float b) { return (1.f - factor) * a + factor * b; }   // ospcommon lerp(factor,a,b), math.h
static void build_curves(int nseg, const int *indices, const float *verts, const float *data, float radius0, float radius1,
                         float value0, float value1, std::vector<float> &out) {
  std::vector<float> vc;          // vertexCurve: a middle segment pushes 3 points, its 4th is the next segment's first
  std::vector<size_t> ic(nseg);   // indexCurve
  bool middleSegment = false;
  V3 tangent = mk(0.f, 0.f, 0.f);
  auto push = [&](V3 p, float r) { vc.push_back(p.x); vc.push_back(p.y); vc.push_back(p.z); vc.push_back(r); };
  auto vert = [&](int i) { return mk(verts[3 * (size_t)i], verts[3 * (size_t)i + 1], verts[3 * (size_t)i + 2]); };
  for (int i = 0; i < nseg; i++) {
    const int idx = indices[i];
    const V3 start = vert(idx), end = vert(idx + 1);
    const V3 se = start - end;
    const float lengthSegment = sqrtf(dot(se, se));
    const float startRadius = map_radius(data ? data[idx] : 0.f, radius0, radius1, value0, value1);
    const float endRadius = map_radius(data ? data[idx + 1] : 0.f, radius0, radius1, value0, value1);
    ic[i] = vc.size() / 4;
    if (middleSegment) {
      push(start, startRadius);
      push(start + tangent, lerp_os(1.f / 3, startRadius, endRadius));
    } else {
      push(start, startRadius);
      push(start, startRadius);
    }
    middleSegment = i + 1 < nseg && indices[i + 1] == idx + 1;
    if (middleSegment) {
      const V3 next = vert(idx + 2);
      const V3 delta = (1.f / 3) * (next - start);
      const V3 ne = next - end;
      const float b = sqrtf(dot(ne, ne));
      const float r = lengthSegment / (lengthSegment + b);
      push(end - r * delta, lerp_os(2.f / 3, startRadius, endRadius));
      tangent = (1.f - r) * delta;
    } else {
      push(end, endRadius);
      push(end, endRadius);
    }
  }
  // what Embree gathers per primitive: the 4 consecutive vertices from indexCurve[i] (rtcSetSharedGeometryBuffer,
  // DataDrivenPathLines.ispc:319-322), expanded here to 16 floats per segment
  out.resize((size_t)16 * nseg);
  for (int i = 0; i < nseg; i++) memcpy(&out[16 * (size_t)i], &vc[4 * ic[i]], 16 * sizeof(float));
}
}  // namespace

struct gxo_scene {
  V3 gmin, gmax, lmin, lmax;
  std::vector<std::unique_ptr<std::vector<float>>> curve_store;
  // Sampler (src/sampler): sampler operators of a sampling Visualization and the samples this partition collected
  struct SamplerOp { int kind; float param; VolumeData *vol; };   // kind 0 GradientSampler (tolerance), 1 IsoSampler (isovalue)
  std::vector<SamplerOp> svis;
  std::vector<float> samples;   // xyz per sample (Sampler::HandleTerminatedRays, Sampler.cpp:52-92)
  int neighbors[6];
  std::vector<VolumeVisOp> vvis;
  std::map<int, std::unique_ptr<VolumeData>> volumes;
  std::vector<GeomOp> geoms;
  // accel
  std::vector<Prim> prims;
  std::vector<BNode> nodes;
  // last secondary list
  std::vector<float> secondary;
  int secondary_n = 0, secondary_aligned = 0;
  std::atomic<long long> sample_count{0};
  // external nearest-hit provider for the triangles of geometry operator 0 (the reference's own Embree, oracle/embree_scene_ref.cpp):
  // when set, trace_kernel takes the nearest hit of every ray from it (in blocks, so that it can run 8-ray packets) instead of walking
  // the median-split tree below, and commit does not build that tree.  bench.py's CPU arm "reference"; never set by a parity test.
  gxo_intersect_fn ext_fn = nullptr;
  void *ext_user = nullptr;
};

namespace {

// ---------------------------------------------------------------------------------------------
// oracle-private BVH (median split over prim boxes); traversal visits every node whose slightly
// enlarged box overlaps the ray interval, so the nearest hit equals a brute-force scan.
static void prim_box(const gxo_scene &s, const Prim &p, float lo[3], float hi[3]) {
  const GeomOp &g = s.geoms[p.geom];
  if (g.kind == 0) {
    for (int k = 0; k < 3; k++) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
    for (int j = 0; j < 3; j++) {
      const float *v = g.verts + 3 * (size_t)g.idx[3 * (size_t)p.prim + j];
      for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
    }
  } else if (g.kind == 1) {
    float r = sphere_radius(g, p.prim);
    for (int k = 0; k < 3; k++) { lo[k] = g.centers[3 * (size_t)p.prim + k] - r; hi[k] = g.centers[3 * (size_t)p.prim + k] + r; }
  } else {
    // the swept surface lies in the convex hull of the control points grown by the largest control radius
    const float *c = g.cp + 16 * (size_t)p.prim;
    const float r = std::max(std::max(fabsf(c[3]), fabsf(c[7])), std::max(fabsf(c[11]), fabsf(c[15])));
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(std::min(c[k], c[4 + k]), std::min(c[8 + k], c[12 + k])) - r;
      hi[k] = std::max(std::max(c[k], c[4 + k]), std::max(c[8 + k], c[12 + k])) + r;
    }
  }
}

static int build_node(gxo_scene &s, std::vector<float> &boxes, std::vector<int> &order, int first, int count) {
  BNode n;
  for (int k = 0; k < 3; k++) { n.lo[k] = FLT_MAX; n.hi[k] = -FLT_MAX; }
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = first; i < first + count; i++) {
    const float *b = &boxes[6 * (size_t)order[i]];
    for (int k = 0; k < 3; k++) {
      n.lo[k] = std::min(n.lo[k], b[k]); n.hi[k] = std::max(n.hi[k], b[3 + k]);
      float c = 0.5f * (b[k] + b[3 + k]);
      clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c);
    }
  }
  n.left = n.right = -1; n.first = first; n.count = count;
  int id = (int)s.nodes.size();
  s.nodes.push_back(n);
  if (count > 4) {
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    int mid = first + count / 2;
    std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count, [&](int a, int b) {
      float ca = boxes[6 * (size_t)a + axis] + boxes[6 * (size_t)a + 3 + axis];
      float cb = boxes[6 * (size_t)b + axis] + boxes[6 * (size_t)b + 3 + axis];
      return ca < cb || (ca == cb && a < b);
    });
    int l = build_node(s, boxes, order, first, mid - first);
    int r = build_node(s, boxes, order, mid, first + count - mid);
    s.nodes[id].left = l; s.nodes[id].right = r; s.nodes[id].count = 0;
  }
  return id;
}

static void build_accel(gxo_scene &s) {
  s.prims.clear(); s.nodes.clear();
  std::vector<Prim> all;
  for (int g = 0; g < (int)s.geoms.size(); g++) {
    int np = s.geoms[g].kind == 0 ? s.geoms[g].nt : s.geoms[g].kind == 1 ? s.geoms[g].n : s.geoms[g].ncurves;
    for (int i = 0; i < np; i++) all.push_back(Prim{g, i});
  }
  if (all.empty()) return;
  std::vector<float> boxes(6 * all.size());
  for (size_t i = 0; i < all.size(); i++) prim_box(s, all[i], &boxes[6 * i], &boxes[6 * i + 3]);
  std::vector<int> order(all.size());
  for (size_t i = 0; i < all.size(); i++) order[i] = (int)i;
  s.nodes.reserve(all.size());
  build_node(s, boxes, order, 0, (int)all.size());
  s.prims.resize(all.size());
  for (size_t i = 0; i < all.size(); i++) s.prims[i] = all[order[i]];
}

static inline bool box_overlap(const BNode &n, V3 org, V3 dir, double tnear, double tfar) {
  // conservative double-precision slab test with a relative pad; never rejects a true overlap
  double t0 = tnear, t1 = tfar;
  const float o[3] = {org.x, org.y, org.z}, d[3] = {dir.x, dir.y, dir.z};
  for (int k = 0; k < 3; k++) {
    double pad = 1e-5 * (1.0 + std::fabs((double)n.lo[k]) + std::fabs((double)n.hi[k]));
    double lo = (double)n.lo[k] - pad, hi = (double)n.hi[k] + pad;
    if (d[k] == 0.f) { if (o[k] < lo || o[k] > hi) return false; continue; }
    double a = (lo - o[k]) / d[k], b = (hi - o[k]) / d[k];
    if (a > b) std::swap(a, b);
    t0 = std::max(t0, a); t1 = std::min(t1, b);
    if (t0 > t1 * (1 + 1e-6) + 1e-6) return false;
  }
  return true;
}

// nearest hit in (tnear, tfar]; smallest t wins, ties -> lowest (geomID, primID).
static bool nearest_hit(const gxo_scene &s, V3 org, V3 dir, float tnear, float tfar, Hit1 &best) {
  best.geomID = -1; best.primID = -1; best.t = tfar;
  if (s.nodes.empty()) return false;
  int stack[128]; int sp = 0; stack[sp++] = 0;
  bool found = false;
  while (sp) {
    const BNode &n = s.nodes[stack[--sp]];
    if (!box_overlap(n, org, dir, tnear, best.t)) continue;
    if (n.left >= 0) { stack[sp++] = n.left; stack[sp++] = n.right; continue; }
    for (int i = n.first; i < n.first + n.count; i++) {
      const Prim &p = s.prims[i];
      const GeomOp &g = s.geoms[p.geom];
      float t, u = 0, v = 0; V3 Ng;
      bool h;
      if (g.kind == 0) {
        const int *ix = g.idx + 3 * (size_t)p.prim;
        V3 v0 = mk(g.verts[3 * (size_t)ix[0]], g.verts[3 * (size_t)ix[0] + 1], g.verts[3 * (size_t)ix[0] + 2]);
        V3 v1 = mk(g.verts[3 * (size_t)ix[1]], g.verts[3 * (size_t)ix[1] + 1], g.verts[3 * (size_t)ix[1] + 2]);
        V3 v2 = mk(g.verts[3 * (size_t)ix[2]], g.verts[3 * (size_t)ix[2] + 1], g.verts[3 * (size_t)ix[2] + 2]);
        // candidates are accepted against the ORIGINAL interval, the nearest is picked after
        h = tri_test(v0, v1, v2, org, dir, tnear, tfar, t, u, v, Ng);
      } else if (g.kind == 1) {
        // spheres: strict t < tfar (DataDrivenSpheres.ispc:135-141)
        h = sphere_test(g, p.prim, org, dir, tnear, tfar, t, Ng);
      } else {
        // curves: tnear < t < tfar, u in [0,1] (curve_intersector_sweep.h:108-109); u is the curve parameter
        h = curve_test(g.cp + 16 * (size_t)p.prim, org, dir, tnear, tfar, t, u, Ng);
      }
      if (!h) continue;
      bool better = !found || t < best.t ||
                    (t == best.t && (p.geom < best.geomID || (p.geom == best.geomID && p.prim < best.primID)));
      if (better) { found = true; best.geomID = p.geom; best.primID = p.prim; best.t = t; best.u = u; best.v = v; best.Ng = Ng; }
    }
  }
  return found;
}

// ---------------------------------------------------------------------------------------------
// RayList view (Rays.ih:20-47 column order)
struct RL {
  float *ox, *oy, *oz, *dx, *dy, *dz, *nx, *ny, *nz, *sample, *r, *g, *b, *o, *sr, *sg, *sb, *so, *t, *tMax;
  int *x, *y, *type, *term, *classification;
  int n, aligned;
};
static RL view(float *base, int n, int aligned) {
  RL v; float **f = &v.ox;
  for (int k = 0; k < 20; k++) f[k] = base + (size_t)k * aligned;
  int **ip = &v.x;
  for (int k = 0; k < 5; k++) ip[k] = (int *)(base + (size_t)(20 + k) * aligned);
  v.n = n; v.aligned = aligned;
  return v;
}
static inline int align16(int n) { return (n + 15) & ~15; }   // Rays.cpp:57-58
struct OwnedRL {
  std::vector<float> mem; RL v; int kind;   // kind 0 PRIMARY list, 1 SECONDARY (RayList::RayListType)
  OwnedRL(int n, int kind_) : mem((size_t)25 * std::max(16, align16(n))), kind(kind_) { v = view(mem.data(), n, std::max(16, align16(n))); }
};
static void copy_ray(const RL &s, int i, RL &d, int j) {      // RayList::CopyRay Rays.h:64-91
  float *const *sf = &s.ox; float **df = &d.ox;
  for (int k = 0; k < 20; k++) df[k][j] = sf[k][i];
  int *const *si = &s.x; int **di = &d.x;
  for (int k = 0; k < 5; k++) di[k][j] = si[k][i];
}

// AO direction tables: src/renderer/UV.ih:21-58 hold a base-2 / base-3 radical inverse printed
// with 6 significant digits.  Regenerated here (fp32 accumulate, multiply by fp32 1/b, "%g").
static void halton_tables(float U[256], float V[256]) {
  for (int pass = 0; pass < 2; pass++) {
    int b = pass ? 3 : 2;
    for (int i = 0; i < 256; i++) {
      float inv = 1.f / (float)b, f = inv, r = 0.f;
      for (int k = i; k > 0; k /= b) { r = r + f * (float)(k % b); f = f * inv; }
      char buf[64]; snprintf(buf, sizeof buf, "%g", (double)r);
      (pass ? V : U)[i] = strtof(buf, nullptr);
    }
  }
}

static void parallel_for(int n, int nthreads, const std::function<void(int, int)> &fn) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (n < 4096 || nthreads == 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  int chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int a = t * chunk, b = std::min(n, a + chunk);
    if (a >= b) break;
    th.emplace_back(fn, a, b);
  }
  for (auto &t : th) t.join();
}

// Diagnostic switch (tests only): integrate the DVR interval BEFORE searching it for an
// isosurface crossing, i.e. with the un-clipped step.  The reference's nineBalls golds match this
// order (see DESIGN.md "gold provenance"); the reference's current code (and the default here)
// searches first (TraceRays.ispc:483-506).
static int g_dvr_before_iso = 0;

struct Hit {   // TraceRays.ispc:79-88
  float sample, t; V3 point, normal, color; float opacity;
};

// ---------------------------------------------------------------------------------------------
// TraceRays_TraceRays, one ray (TraceRays.ispc:326-623)
static void trace_one(gxo_scene &S, RL &R, int i, bool integrate, float step, float epsilon,
                      long long &nsamples, int *hit_ids, const Hit1 *pre = nullptr) {
  const int nvv = (int)S.vvis.size();
  bool shadeFlag = R.type[i] == RAY_PRIMARY;
  V3 org = mk(R.ox[i], R.oy[i], R.oz[i]);
  V3 dir = mk(R.dx[i], R.dy[i], R.dz[i]);
  if (dir.x == 0.f) dir.x = 1e-6f;                           // :377-379
  if (dir.y == 0.f) dir.y = 1e-6f;
  if (dir.z == 0.f) dir.z = 1e-6f;
  float ray_t0 = R.t[i], ray_t = R.tMax[i];
  float tTimeout = ray_t;
  float color[4] = {R.r[i], R.g[i], R.b[i], R.o[i]};

  // MyIntersectBox :90-106 with rcp(dir) := 1.0f/dir
  float tEntry, tExitVolume;
  {
    V3 rd = mk(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    V3 mins = mk((S.lmin.x - org.x) * rd.x, (S.lmin.y - org.y) * rd.y, (S.lmin.z - org.z) * rd.z);
    V3 maxs = mk((S.lmax.x - org.x) * rd.x, (S.lmax.y - org.y) * rd.y, (S.lmax.z - org.z) * rd.z);
    tEntry = std::max(std::min(mins.x, maxs.x), std::max(std::min(mins.y, maxs.y), std::min(mins.z, maxs.z)));
    tExitVolume = std::min(std::max(mins.x, maxs.x), std::min(std::max(mins.y, maxs.y), std::max(mins.z, maxs.z)));
  }
  if (tEntry < ray_t0) tEntry = ray_t0;                      // :412-413
  else if (tEntry > ray_t0) ray_t0 = tEntry;
  ray_t = std::min(ray_t, tExitVolume);                      // :418

  Hit hit; memset(&hit, 0, sizeof hit);
  bool surface_hit = false;

  // LookForSliceHit :143-250
  {
    bool h = false; int vid = -1;
    for (int major = 0; major < nvv; major++) {
      const VolumeVisOp &vv = S.vvis[major];
      for (size_t minor = 0; minor < vv.slices.size() / 4; minor++) {
        const float *pl = &vv.slices[4 * minor];
        V3 pnorm = mk(pl[0], pl[1], pl[2]);
        float d = pl[3];
        float denom = dot(dir, pnorm);
        if (fabsf(denom) > 0.0001f) {
          float t = (d - dot(org, pnorm)) / denom;
          if (t >= ray_t0 && t <= ray_t) {
            hit.normal = denom > 0 ? neg(pnorm) : pnorm;
            hit.opacity = 1.0f; hit.t = t; vid = major; ray_t = hit.t; h = true;
          }
        }
      }
    }
    if (h && shadeFlag) {
      hit.point = org + hit.t * dir;
      const VolumeVisOp &vv = S.vvis[vid];
      hit.sample = vol_sample(*vv.vol, hit.point); nsamples++;
      hit.color = tf_color(vv.tf, hit.sample);
      hit.opacity = 1.0f;
    }
    surface_hit = h;
  }

  // LookForGeometryHit :108-141 (model is always non-NULL, Visualization.cpp:219-277)
  if (hit_ids) { hit_ids[2 * i] = -1; hit_ids[2 * i + 1] = -1; }
  if (!S.geoms.empty()) {
    Hit1 h1;
    bool found1;
    if (pre) { h1 = *pre; found1 = h1.geomID >= 0; }
    else found1 = nearest_hit(S, org, dir, ray_t0, ray_t, h1);
    if (found1) {
      ray_t = h1.t;
      if (hit_ids) { hit_ids[2 * i] = h1.geomID; hit_ids[2 * i + 1] = h1.primID; }
      if (shadeFlag) {
        // postIntersect: ospray/common/Model.ih:97-187 + geometry-specific part
        const GeomOp &g = S.geoms[h1.geomID];
        V3 Ng = h1.Ng, Ns = h1.Ng;
        float cr = 1.f, cg = 1.f, cb = 1.f, ca = 1.f;
        if (g.kind == 0) {  // DataDrivenTriangleMesh.ispc:34-121
          const int *ix = g.idx + 3 * (size_t)h1.primID;
          V3 bary = mk(1.0f - h1.u - h1.v, h1.u, h1.v);
          if (g.normals) {
            V3 a = mk(g.normals[3 * (size_t)ix[0]], g.normals[3 * (size_t)ix[0] + 1], g.normals[3 * (size_t)ix[0] + 2]);
            V3 b = mk(g.normals[3 * (size_t)ix[1]], g.normals[3 * (size_t)ix[1] + 1], g.normals[3 * (size_t)ix[1] + 2]);
            V3 c = mk(g.normals[3 * (size_t)ix[2]], g.normals[3 * (size_t)ix[2] + 1], g.normals[3 * (size_t)ix[2] + 2]);
            Ns = bary.x * a + bary.y * b + bary.z * c;      // interpolate(): vec.ih:723-726
          }
          if (g.data) {
            float d = bary.x * g.data[ix[0]] + bary.y * g.data[ix[1]] + bary.z * g.data[ix[2]];
            V3 c = tf_color(g.tf, d);
            cr = c.x; cg = c.y; cb = c.z; ca = 1.0f;
          }
        } else if (g.kind == 1) {   // DataDrivenSpheres.ispc:46-63
          V3 c = tf_color(g.tf, g.data ? g.data[h1.primID] : 0.f);
          cr = c.x; cg = c.y; cb = c.z; ca = 1.0f;
        } else {            // DataDrivenPathLines_postIntersect, DataDrivenPathLines.ispc:213-277: Ng = Ns = ray.Ng; the
          // radius between the segment's FIRST TWO control points (vertices[indices[primID]], [+1]) is mapped back to a data value
          const float *c4 = g.cp + 16 * (size_t)h1.primID;
          const float radius = ((1.f - h1.u) * c4[3]) + (h1.u * c4[7]);
          float dataval;
          if (g.radius0 == g.radius1) dataval = g.value0;
          else if (g.radius0 < g.radius1 && radius < g.radius0) dataval = g.value0;
          else if (g.radius0 < g.radius1 && radius > g.radius1) dataval = g.value1;
          else if (g.radius0 > g.radius1 && radius < g.radius1) dataval = g.value1;
          else if (g.radius0 > g.radius1 && radius > g.radius0) dataval = g.value0;
          else dataval = g.value0 + ((radius - g.radius0) / (g.radius1 - g.radius0)) * (g.value1 - g.value0);
          V3 c = tf_color(g.tf, dataval);
          cr = c.x; cg = c.y; cb = c.z; ca = 1.0f;
        }
        V3 ffnng = normalize_isp(Ng);
        Ng = ffnng;
        Ns = normalize_isp(Ns);
        const bool flip = dot(dir, Ng) >= 0.f;
        if (flip) Ng = neg(Ng);
        if (dot(Ng, Ns) < 0.f) Ns = neg(Ns);
        hit.color = mk(cr, cg, cb);
        hit.opacity = ca;
        hit.normal = Ns;
        hit.t = ray_t;
      }
      surface_hit = true;
    }
  }

  float tTermination = ray_t;

  if (integrate) {   // :446-570
    float tLast, tThis;
    float sLast[100], sThis[100];
    bool hit_isosurface = false;
    tLast = tEntry + epsilon;
    bool opaque = (min3(color[0], color[1], color[2]) >= 1.0f || color[3] > 0.999f);

    for (tThis = tEntry; tThis <= tTermination && !opaque && !hit_isosurface;
         tThis = (tThis == tEntry) ? (tEntry + epsilon)
                                   : (((tThis + step) > tTermination) && (tThis < tTermination)) ? tTermination : tThis + step) {
      V3 coord = org + tThis * dir;
      for (int m = 0; m < nvv; m++) sThis[m] = vol_sample(*S.vvis[m].vol, coord);   // SampleVolumes :312-324
      nsamples += nvv;

      if (tThis > tEntry && tLast >= epsilon) {
        auto integrate_dvr = [&](bool update_last) {
          for (int major = 0; major < nvv; major++) {   // :512-542
            const VolumeVisOp &vv = S.vvis[major];
            if (vv.volume_render) {
              const VolumeData *vol = vv.vol;
              const TF &tf = *vol->tf;
              float sVolume = (sLast[major] + sThis[major]) / 2;
              float sampleOpacity = tf_opacity(tf, sVolume);
              if (sampleOpacity > 0) {
                float cl = std::max(0.0f, std::min(sampleOpacity / vol->samplingRate, 1.0f));
                if (shadeFlag) {
                  V3 sc = tf_color(tf, sVolume);
                  float wo = cl;
                  float w4[4] = {wo * sc.x, wo * sc.y, wo * sc.z, wo * 1.0f};
                  float om = 1.0f - color[3];
                  for (int k = 0; k < 4; k++) color[k] = color[k] + om * w4[k];
                } else {
                  float weightedOpacity = ((tThis - tLast) / step) * cl;
                  float f = 1 - weightedOpacity;
                  for (int k = 0; k < 4; k++) color[k] = color[k] * f;
                }
              }
            }
            if (update_last) sLast[major] = sThis[major];
          }
        };
        if (g_dvr_before_iso) integrate_dvr(false);
        // LookForIsoHit :252-310
        bool h = false; int vid = -1;
        for (int major = 0; major < nvv; major++) {
          float sl = sLast[major], st = sThis[major];
          const VolumeVisOp &vv = S.vvis[major];
          for (size_t minor = 0; minor < vv.isovalues.size(); minor++) {
            float isoval = vv.isovalues[minor];
            if (((isoval >= sl) && (isoval < st)) || ((isoval <= sl) && (isoval > st))) {
              h = true; vid = major;
              hit.t = tLast + ((isoval - sl) / (st - sl)) * (tThis - tLast);
              hit.sample = isoval;
            }
          }
        }
        if (h) {
          hit.point = org + hit.t * dir;
          if (shadeFlag) {
            const VolumeVisOp &vv = S.vvis[vid];
            hit.normal = safe_normalize(vol_gradient(*vv.vol, hit.point)); nsamples += 4;
            if (dot(dir, hit.normal) > 0) hit.normal = neg(hit.normal);
            hit.color = tf_color(vv.tf, hit.sample);
            hit.opacity = 1.0f;
          }
          tTermination = hit.t; tThis = hit.t;
          surface_hit = true; hit_isosurface = true;
          V3 c2 = org + tThis * dir;
          for (int m = 0; m < nvv; m++) sThis[m] = vol_sample(*S.vvis[m].vol, c2);
          nsamples += nvv;
        }
        if (!g_dvr_before_iso) integrate_dvr(true);
      }
      for (int m = 0; m < nvv; m++) sLast[m] = sThis[m];
      opaque = (min3(color[0], color[1], color[2]) >= 1.0f || color[3] > 0.999f);
      if (opaque) tTermination = tThis;
      tLast = tThis;
    }
    if (tThis > tTermination) tThis = tTermination;
    ray_t = tTermination;
  }

  R.r[i] = color[0]; R.g[i] = color[1]; R.b[i] = color[2]; R.o[i] = color[3];
  int term = (min3(color[0], color[1], color[2]) >= 1.0f || color[3] > 0.999f) ? RAY_OPAQUE : 0;
  R.t[i] = ray_t;
  if (surface_hit) {
    term |= RAY_SURFACE;
    // for non-shaded rays hit.opacity is uninitialised in the reference unless a slice set it;
    // every surface kind sets opacity 1 when shaded (:133,:195,:240,:304).  Treat as 1.
    if (!shadeFlag || hit.opacity > 0.999f) term |= RAY_OPAQUE;
    if (shadeFlag) {
      R.sr[i] = hit.color.x; R.sg[i] = hit.color.y; R.sb[i] = hit.color.z; R.so[i] = 1.0f;
      R.nx[i] = hit.normal.x; R.ny[i] = hit.normal.y; R.nz[i] = hit.normal.z;
    }
  } else if (tTermination == tExitVolume) term |= RAY_BOUNDARY;
  else if (tTermination == tTimeout) term |= RAY_TIMEOUT;
  R.term[i] = term;
}

static void trace_setup(const gxo_scene &S, float global_epsilon, bool &integrate, float &step, float &epsilon) {
  // TraceRays.ispc:342-361
  integrate = false; step = -1; epsilon = 0;
  for (const VolumeVisOp &vv : S.vvis) {
    if (vv.volume_render) integrate = true;
    if (!vv.isovalues.empty()) integrate = true;
    float s = vv.vol->samplingStep * vv.vol->samplingRate;
    if (step < 0 || step > s) { step = s; epsilon = global_epsilon * s; }
  }
}

static void trace_kernel(gxo_scene &S, RL &R, float global_epsilon, int nthreads, int *hit_ids) {
  bool integrate; float step, epsilon;
  trace_setup(S, global_epsilon, integrate, step, epsilon);
  std::atomic<long long> total{0};
  const bool ext = S.ext_fn && S.vvis.empty() && S.geoms.size() == 1 && S.geoms[0].kind == 0;
  parallel_for(R.n, nthreads, [&](int a, int b) {
    long long ns = 0;
    if (!ext) {
      for (int i = a; i < b; i++) trace_one(S, R, i, integrate, step, epsilon, ns, hit_ids);
    } else {
      // the interval of LookForGeometryHit for a scene without slices: the clip of trace_one (:377-418), block by block
      const int B = 2048;
      std::vector<float> org(3 * B), dir(3 * B), tn(B), tf(B), t(B), u(B), v(B), ng(3 * B);
      std::vector<int> gid(B), pid(B);
      for (int a0 = a; a0 < b; a0 += B) {
        const int m = std::min(B, b - a0);
        for (int k = 0; k < m; k++) {
          const int i = a0 + k;
          V3 o = mk(R.ox[i], R.oy[i], R.oz[i]), d = mk(R.dx[i], R.dy[i], R.dz[i]);
          if (d.x == 0.f) d.x = 1e-6f;
          if (d.y == 0.f) d.y = 1e-6f;
          if (d.z == 0.f) d.z = 1e-6f;
          float ray_t0 = R.t[i], ray_t = R.tMax[i];
          V3 rd = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
          V3 mins = mk((S.lmin.x - o.x) * rd.x, (S.lmin.y - o.y) * rd.y, (S.lmin.z - o.z) * rd.z);
          V3 maxs = mk((S.lmax.x - o.x) * rd.x, (S.lmax.y - o.y) * rd.y, (S.lmax.z - o.z) * rd.z);
          float tEntry = std::max(std::min(mins.x, maxs.x), std::max(std::min(mins.y, maxs.y), std::min(mins.z, maxs.z)));
          float tExit = std::min(std::max(mins.x, maxs.x), std::min(std::max(mins.y, maxs.y), std::max(mins.z, maxs.z)));
          if (tEntry < ray_t0) tEntry = ray_t0; else if (tEntry > ray_t0) ray_t0 = tEntry;
          ray_t = std::min(ray_t, tExit);
          org[3 * k] = o.x; org[3 * k + 1] = o.y; org[3 * k + 2] = o.z;
          dir[3 * k] = d.x; dir[3 * k + 1] = d.y; dir[3 * k + 2] = d.z;
          tn[k] = ray_t0; tf[k] = ray_t;
        }
        S.ext_fn(S.ext_user, m, org.data(), dir.data(), tn.data(), tf.data(), gid.data(), pid.data(), t.data(), u.data(), v.data(), ng.data());
        for (int k = 0; k < m; k++) {
          Hit1 h;
          h.geomID = (pid[k] >= 0 && tn[k] <= tf[k]) ? 0 : -1; h.primID = pid[k]; h.t = t[k]; h.u = u[k]; h.v = v[k];
          h.Ng = mk(ng[3 * k], ng[3 * k + 1], ng[3 * k + 2]);
          trace_one(S, R, a0 + k, integrate, step, epsilon, ns, hit_ids, &h);
        }
      }
    }
    total += ns;
  });
  S.sample_count += total.load();
}

// ---------------------------------------------------------------------------------------------
// SamplerTraceRays_SamplerTraceRays (src/sampler/SamplerTraceRays.ispc:128-222) with the two sampler operators
// GradientSamplerVis (GradientSamplerVis.ispc:36-72) and IsoSamplerVis (IsoSamplerVis.ispc:36-67).  The reference keeps
// sLast/tLast/tHit as varying members of the (shared) operator struct; they are per-ray state here.  rcp(dir) := 1/dir.
static void sampler_trace(const gxo_scene &S, RL &R, int nthreads) {
  const int nv = (int)S.svis.size();
  if (nv < 1) return;   // :136: nothing is touched without a sampler operator
  float step = S.svis[0].vol->samplingStep * S.svis[0].vol->samplingRate;
  for (int m = 1; m < nv; m++) {
    const float s = S.svis[m].vol->samplingStep * S.svis[m].vol->samplingRate;
    if (s < step) step = s;
  }
  parallel_for(R.n, nthreads, [&](int a, int b) {
    std::vector<V3> gLast(nv); std::vector<float> sLast(nv), tLastV(nv);
    for (int i = a; i < b; i++) {
      const V3 org = mk(R.ox[i], R.oy[i], R.oz[i]);
      V3 dir = mk(R.dx[i], R.dy[i], R.dz[i]);
      const float ray_t = R.t[i];
      if (dir.x == 0.f) dir.x = 1e-6f;
      if (dir.y == 0.f) dir.y = 1e-6f;
      if (dir.z == 0.f) dir.z = 1e-6f;
      // EntryT / ExitT (:62-86)
      const V3 rd = mk(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
      const V3 mins = mk((S.lmin.x - org.x) * rd.x, (S.lmin.y - org.y) * rd.y, (S.lmin.z - org.z) * rd.z);
      const V3 maxs = mk((S.lmax.x - org.x) * rd.x, (S.lmax.y - org.y) * rd.y, (S.lmax.z - org.z) * rd.z);
      float tEntry = std::max(std::min(mins.x, maxs.x), std::max(std::min(mins.y, maxs.y), std::min(mins.z, maxs.z)));
      const float tExit = std::min(std::max(mins.x, maxs.x), std::min(std::max(mins.y, maxs.y), std::max(mins.z, maxs.z)));
      if (tEntry < ray_t) tEntry = ray_t;
      float tThis = tEntry + step;
      int hit = -1;
      for (int m = 0; m < nv; m++) {   // init (:186-190)
        const V3 coord = org + tEntry * dir;
        if (S.svis[m].kind == 0) gLast[m] = vol_gradient(*S.svis[m].vol, coord);
        else sLast[m] = vol_sample(*S.svis[m].vol, coord);
        tLastV[m] = tEntry;
      }
      while (tThis <= tExit && hit == -1) {
        hit = -1;
        for (int m = 0; m < nv && hit == -1; m++) {   // check_interval
          const V3 coord = org + tThis * dir;
          bool h = false; float tHit = 0.f;
          if (S.svis[m].kind == 0) {
            const V3 gThis = vol_gradient(*S.svis[m].vol, coord);
            const float dotValue = dot(gThis, gLast[m]);
            if (dotValue < S.svis[m].param) { tHit = (tLastV[m] + tThis) / 2.0f; h = true; }
            gLast[m] = gThis;
          } else {
            const float iso = S.svis[m].param, sThis = vol_sample(*S.svis[m].vol, coord);
            if (((sLast[m] < iso) && (sThis >= iso)) || ((sLast[m] > iso) && sThis <= iso)) {
              tHit = tLastV[m] + (((iso - sLast[m]) / (sThis - sLast[m])) * (tThis - tLastV[m]));
              h = true;
            }
            sLast[m] = sThis;
          }
          tLastV[m] = tThis;
          if (h) { tThis = tHit; hit = m; }
        }
        if (hit != -1 || tThis == tExit) break;
        tThis = tThis + step;
        if (tThis > tExit) tThis = tExit;
      }
      R.t[i] = (hit != -1) ? tThis + 0.001f : tThis;   // ISPC literals without a suffix are float
      R.term[i] = (hit != -1) ? RAY_SURFACE : RAY_BOUNDARY;
    }
  });
}

// ---------------------------------------------------------------------------------------------
// TraceRays::Trace (TraceRays.cpp:68-146) with GXY_REVERSE_LIGHTING (CMakeLists.txt:66)
static std::unique_ptr<OwnedRL> trace_and_spawn(gxo_scene &S, const gxo_lighting &L, RL &R, float eps, int nthreads,
                                                int *hit_ids, long long *n_ao_out, long long *n_sh_out) {
  trace_kernel(S, R, eps, nthreads, hit_ids);
  const int n = R.n;
  std::vector<int> ao_offsets(n), shadow_offsets(n, -1);
  int nAO = L.n_ao; bool do_shadows = L.shadows != 0; int nLights = L.n_lights;
  int nOut = 0;
  for (int i = 0; i < n; i++) {
    ao_offsets[i] = nOut;
    if (R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) nOut += nAO;
  }
  int ao_knt = nOut;
  if (do_shadows)
    for (int i = 0; i < n; i++) {
      if (R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) { shadow_offsets[i] = nOut; nOut += nLights; }
      else shadow_offsets[i] = -1;
    }
  int sh_knt = nOut - ao_knt;
  if (n_ao_out) *n_ao_out = ao_knt;
  if (n_sh_out) *n_sh_out = sh_knt;
  std::unique_ptr<OwnedRL> out;
  if (nOut) out.reset(new OwnedRL(nOut, 1));

  // (the four loops below are independent per ray -- own columns, own precomputed output slots -- and run threaded; the reference runs
  // them serially inside one pool thread per RayList, Renderer.cpp:504-556)
  // ambientLighting :735-761
  parallel_for(n, nthreads, [&](int lo_, int hi_) {
  for (int i = lo_; i < hi_; i++)
    if (R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) {
      float ambient_scale = L.Ka * (1.0f - R.o[i]);
      R.r[i] += ambient_scale * R.sr[i]; R.g[i] += ambient_scale * R.sg[i]; R.b[i] += ambient_scale * R.sb[i];
    }
  });

  // generateAORays :625-733
  if (ao_knt) {
    static float U[256], V[256]; static bool init = false;
    if (!init) { halton_tables(U, V); init = true; }
    RL &O = out->v;
    float Ka = -L.Ka / L.n_ao;
    const float epsilon = eps;
    parallel_for(n, nthreads, [&](int lo_, int hi_) {
    for (int i = lo_; i < hi_; i++)
      if (R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) {
        V3 sn = mk(R.nx[i], R.ny[i], R.nz[i]);
        float ambient_scale = Ka * (1.0f - R.o[i]);
        float ar = ambient_scale * R.sr[i], ag = ambient_scale * R.sg[i], ab = ambient_scale * R.sb[i];
        V3 b0 = mk(1.0f, 0.0f, 0.0f);
        if (fabsf(dot(b0, sn)) > 0.95f) b0 = mk(0.0f, 1.0f, 0.0f);
        V3 b1 = normalize_isp(cross(b0, sn));
        b0 = normalize_isp(cross(b1, sn));
        float ox = R.ox[i] + R.t[i] * R.dx[i], oy = R.oy[i] + R.t[i] * R.dy[i], oz = R.oz[i] + R.t[i] * R.dz[i];
        ox = ox + epsilon * sn.x; oy = oy + epsilon * sn.y; oz = oz + epsilon * sn.z;
        for (int j = 0; j < nAO; j++) {
          int offset = ao_offsets[i] + j;
          int r = ((R.x[i] * 9949 + R.y[i] * 9613 + j * 9151) >> 8) & 0xff;
          const float r0 = U[r], r1 = V[r];
          const float w = sqrtf(1.f - r1);
          const float x = cosf((2.f * (float)M_PI) * r0) * w;
          const float y = sinf((2.f * (float)M_PI) * r0) * w;
          const float z = sqrtf(r1) + epsilon;
          V3 rd = x * b0 + y * b1 + z * sn;
          O.ox[offset] = ox; O.oy[offset] = oy; O.oz[offset] = oz;
          O.dx[offset] = rd.x; O.dy[offset] = rd.y; O.dz[offset] = rd.z;
          O.r[offset] = ar; O.g[offset] = ag; O.b[offset] = ab; O.o[offset] = 0.0f;
          O.t[offset] = 0.0f; O.tMax[offset] = L.ao_radius;
          O.x[offset] = R.x[i]; O.y[offset] = R.y[i]; O.type[offset] = RAY_AO; O.term[offset] = 0;
        }
      }
    });
  }

  // diffuseLighting :859-923
  {
    float Kd = L.Kd / L.n_lights;
    parallel_for(n, nthreads, [&](int lo_, int hi_) {
    for (int i = lo_; i < hi_; i++)
      if (R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) {
        V3 sn = mk(R.nx[i], R.ny[i], R.nz[i]);
        float tr = 0, tg = 0, tb = 0;
        for (int k = 0; k < L.n_lights; k++) {
          V3 lvec; V3 lt = mk(L.lights[k][0], L.lights[k][1], L.lights[k][2]);
          if (L.types[k]) {
            float t = R.t[i];
            V3 sp = mk(R.ox[i] + t * R.dx[i], R.oy[i] + t * R.dy[i], R.oz[i] + t * R.dz[i]);
            lvec = safe_normalize(lt - sp);
          } else lvec = neg(lt);
          float d = dot(sn, lvec);
          if (d > 0) {
            float dff = (1.0f - R.o[i]) * d;
            tr += dff * R.sr[i]; tg += dff * R.sg[i]; tb += dff * R.sb[i];
          }
        }
        R.r[i] = R.r[i] + Kd * (1 - R.o[i]) * tr;
        R.g[i] = R.g[i] + Kd * (1 - R.o[i]) * tg;
        R.b[i] = R.b[i] + Kd * (1 - R.o[i]) * tb;
        R.o[i] = R.o[i] + Kd * (1 - R.o[i]) * R.o[i];
      }
    });
  }

  // generateShadowRays :763-857
  if (sh_knt) {
    RL &O = out->v;
    float Kd = -L.Kd / L.n_lights;
    const float epsilon = eps;
    parallel_for(n, nthreads, [&](int lo_, int hi_) {
    for (int i = lo_; i < hi_; i++) {
      int offset = shadow_offsets[i];
      if (offset == -1) continue;
      V3 sn = mk(R.nx[i], R.ny[i], R.nz[i]);
      for (int k = 0; k < L.n_lights; k++) {
        V3 sp = mk(R.ox[i] + R.t[i] * R.dx[i] + epsilon * R.nx[i], R.oy[i] + R.t[i] * R.dy[i] + epsilon * R.ny[i],
                   R.oz[i] + R.t[i] * R.dz[i] + epsilon * R.nz[i]);
        V3 lvec; V3 lt = mk(L.lights[k][0], L.lights[k][1], L.lights[k][2]);
        if (L.types[k]) lvec = safe_normalize(lt - sp);
        else lvec = neg(lt);
        lvec = safe_normalize(lvec);
        float d = dot(sn, lvec);
        if (d < 0) d = 0;
        float dff = (1.0f - R.o[i]) * Kd * d;
        O.ox[offset] = sp.x; O.oy[offset] = sp.y; O.oz[offset] = sp.z;
        O.dx[offset] = lvec.x; O.dy[offset] = lvec.y; O.dz[offset] = lvec.z;
        O.r[offset] = dff * R.sr[i]; O.g[offset] = dff * R.sg[i]; O.b[offset] = dff * R.sb[i]; O.o[offset] = 0.0f;
        O.t[offset] = 0.0f; O.tMax[offset] = std::numeric_limits<float>::infinity();
        O.x[offset] = R.x[i]; O.y[offset] = R.y[i]; O.type[offset] = RAY_SHADOW; O.term[offset] = 0;
        offset++;
      }
    }
    });
  }
  return out;
}

// Box::exit_face (Box.cpp:84-97)
static int exit_face(V3 mn, V3 mx, float x, float y, float z, float dx, float dy, float dz) {
  float tx = (dx > 0.0001f) ? ((mx.x - x) / dx) : (dx < -0.0001f) ? ((mn.x - x) / dx) : FLT_MAX;
  float ty = (dy > 0.0001f) ? ((mx.y - y) / dy) : (dy < -0.0001f) ? ((mn.y - y) / dy) : FLT_MAX;
  float tz = (dz > 0.0001f) ? ((mx.z - z) / dz) : (dz < -0.0001f) ? ((mn.z - z) / dz) : FLT_MAX;
  if (tx < 0) tx = FLT_MAX;
  if (ty < 0) ty = FLT_MAX;
  if (tz < 0) tz = FLT_MAX;
  if (tx < ty && tx < tz) return (dx < 0) ? 0 : 1;
  else if (ty < tz) return (dy < 0) ? 2 : 3;
  else return (dz < 0) ? 4 : 5;
}

// Renderer::Classify + AssignDestinations (Renderer.cpp:304-454), REVERSE_LIGHTING
static void classify(const gxo_scene &S, RL &R) {
  for (int i = 0; i < R.n; i++) {
    int typ = R.type[i], term = R.term[i];
    int c = UNDETERMINED;
    if (typ == RAY_PRIMARY) {
      if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
      else if ((term & RAY_OPAQUE) | (term & RAY_TIMEOUT)) c = TERMINATED;
      else c = KEEP_HERE;
    } else if (typ == RAY_SHADOW) {
      if ((term & RAY_OPAQUE) | (term & RAY_SURFACE)) c = TERMINATED;
      else if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
      else c = DROP_ON_FLOOR;
    } else if (typ == RAY_AO) {
      if ((term & RAY_OPAQUE) | (term & RAY_SURFACE)) c = TERMINATED;
      else if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
      else if (term & RAY_TIMEOUT) c = DROP_ON_FLOOR;
      else c = DROP_ON_FLOOR;
    }
    if (c == RAY_BOUNDARY) {
      int f = exit_face(S.lmin, S.lmax, R.ox[i], R.oy[i], R.oz[i], R.dx[i], R.dy[i], R.dz[i]);
      if (S.neighbors[f] >= 0) c = S.neighbors[f];
      else c = (typ == RAY_SHADOW || typ == RAY_AO) ? DROP_ON_FLOOR : TERMINATED;
    }
    R.classification[i] = c;
  }
}

// Box::intersect (Box.cpp:107-150)
static bool box_intersect(V3 mn, V3 mx, V3 org, V3 dir, float &tmin, float &tmax) {
  tmin = (mn.x - org.x) / dir.x;
  tmax = (mx.x - org.x) / dir.x;
  if (tmin > tmax) std::swap(tmax, tmin);
  if (tmax < 0) return false;
  float tymin = (mn.y - org.y) / dir.y, tymax = (mx.y - org.y) / dir.y;
  if (tymin > tymax) std::swap(tymax, tymin);
  if (tymax < 0) return false;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  float tzmin = (mn.z - org.z) / dir.z, tzmax = (mx.z - org.z) / dir.z;
  if (tzmin > tzmax) std::swap(tzmin, tzmax);
  if (tzmax < 0) return false;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  if (tmin < 0) tmin = 0;
  return true;
}

struct CamFrame { V3 veye, vdir, vr, vu, center; float scaling, off_x, off_y; bool ortho; };

// Camera::generate_initial_rays set-up (Camera.cpp:528-582)
static CamFrame camera_frame(const gxo_camera &c, int width, int height) {
  CamFrame f;
  V3 veye = mk(c.eye[0], c.eye[1], c.eye[2]), vu = mk(c.up[0], c.up[1], c.up[2]), vdir = mk(c.dir[0], c.dir[1], c.dir[2]);
  V3 center;
  if (c.aov == 0.0f) { center = veye + vdir; normalize_gxy(vdir); }
  else {
    float d = (float)(1.0 / tan(2 * 3.1415926 * ((double)c.aov / 2.0) / 360.0));
    normalize_gxy(vdir);
    center = veye + (vdir * d);
  }
  V3 vr = cross(vdir, vu); normalize_gxy(vr);
  vu = cross(vr, vdir); normalize_gxy(vu);
  float pixel_scaling = (float)((((width < height) ? width : height) - 1.0) / 2.0);
  f.off_x = (float)((width - 1) / 2.0);
  f.off_y = (float)((height - 1) / 2.0);
  f.scaling = (float)(1.0 / pixel_scaling);   // Camera.h:66
  f.veye = veye; f.vdir = vdir; f.vr = vr; f.vu = vu; f.center = center; f.ortho = (c.aov == 0.0f);
  return f;
}

// Camera::SpawnRays per pixel (Camera.cpp:379-493); returns 1 if kept, 0 if not; *global_hit says
// whether the ray hits the global box at all (for the orphan count).
static inline int spawn_pixel(const gxo_scene &S, const CamFrame &a, int x, int y, V3 &vorigin, V3 &vray, bool *global_hit) {
  float fx = (x - a.off_x) * a.scaling;
  float fy = (y - a.off_y) * a.scaling;
  V3 xy;
  xy.x = a.center.x + fx * a.vr.x + fy * a.vu.x;
  xy.y = a.center.y + fx * a.vr.y + fy * a.vu.y;
  xy.z = a.center.z + fx * a.vr.z + fy * a.vu.z;
  if (a.ortho) { vorigin = xy - a.vdir; vray = a.vdir; }
  else { vorigin = a.veye; vray = xy - a.veye; normalize_gxy(vray); }
  float gmin, gmax, lmin = 0, lmax = 0;
  bool hit = box_intersect(S.gmin, S.gmax, vorigin, vray, gmin, gmax);
  if (global_hit) *global_hit = hit;
  if (hit) hit = box_intersect(S.lmin, S.lmax, vorigin, vray, lmin, lmax);
  float d = fabsf(lmin) - fabsf(gmin);
  return (hit && (lmax >= 0) && (d < 0.000001f) && (d > -0.000001f)) ? 1 : 0;
}

static int generate_range(const gxo_scene &S, const CamFrame &a, int w, int start, int count, RL &R) {
  int dst = 0;
  for (int i = 0; i < count; i++) {
    int p = i + start;
    int x = p % w, y = p / w;
    V3 o, d;
    if (spawn_pixel(S, a, x, y, o, d, nullptr)) {
      R.x[dst] = x; R.y[dst] = y;
      R.ox[dst] = o.x; R.oy[dst] = o.y; R.oz[dst] = o.z;
      R.dx[dst] = d.x; R.dy[dst] = d.y; R.dz[dst] = d.z;
      R.r[dst] = 0; R.g[dst] = 0; R.b[dst] = 0; R.o[dst] = 0; R.t[dst] = 0; R.tMax[dst] = FLT_MAX;
      R.type[dst] = RAY_PRIMARY;
      dst++;
    }
  }
  return dst;
}

}  // namespace

// =============================================================================================
extern "C" {

gxo_scene *gxo_scene_create(void) {
  gxo_scene *s = new gxo_scene();
  s->gmin = s->lmin = mk(0, 0, 0); s->gmax = s->lmax = mk(0, 0, 0);
  for (int i = 0; i < 6; i++) s->neighbors[i] = -1;
  return s;
}
void gxo_scene_destroy(gxo_scene *s) { delete s; }

void gxo_scene_set_partition(gxo_scene *s, const float gmin[3], const float gmax[3], const float lmin[3],
                             const float lmax[3], const int neighbors[6]) {
  s->gmin = mk(gmin[0], gmin[1], gmin[2]); s->gmax = mk(gmax[0], gmax[1], gmax[2]);
  s->lmin = mk(lmin[0], lmin[1], lmin[2]); s->lmax = mk(lmax[0], lmax[1], lmax[2]);
  for (int i = 0; i < 6; i++) s->neighbors[i] = neighbors[i];
}

static VolumeData *scene_volume_impl(gxo_scene *s, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                                     int type, const void *voxels);
static inline VolumeData *scene_volume(gxo_scene *s, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                                       int type, const void *voxels) {
  return scene_volume_impl(s, dataset_id, dims, origin, spacing, type, voxels);
}
static void set_tf(TF &tf, const float *colors, const float *opac, float lo, float hi) {
  memcpy(tf.color, colors, sizeof tf.color); memcpy(tf.opacity, opac, sizeof tf.opacity);
  tf.lo = lo; tf.hi = hi;
}

int gxo_scene_add_volume_vis(gxo_scene *s, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                             int type, const void *voxels, int n_slices, const float *slices4, int n_iso,
                             const float *isovalues, int volume_render, const float *colors, const float *opacities,
                             float lo, float hi) {
  if ((int)s->vvis.size() >= 100) return -1;   // TraceRays.ispc:455 sLast[100]
  VolumeData *vd = scene_volume(s, dataset_id, dims, origin, spacing, type, voxels);
  s->vvis.emplace_back();
  VolumeVisOp &op = s->vvis.back();
  op.dataset_id = dataset_id; op.vol = vd;
  set_tf(op.tf, colors, opacities, lo, hi);
  op.slices.assign(slices4, slices4 + 4 * n_slices);
  op.isovalues.assign(isovalues, isovalues + n_iso);
  op.volume_render = volume_render != 0;
  return (int)s->vvis.size() - 1;
}

/* GradientSamplerVis ("tolerance", GradientSamplerVis.cpp:77-85) / IsoSamplerVis ("isovalue", IsoSamplerVis.cpp:77-85) on a brick */
int gxo_scene_add_sampler_vis(gxo_scene *s, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                              int type, const void *voxels, int kind, float param) {
  if (kind != 0 && kind != 1) return -1;
  gxo_scene::SamplerOp op;
  op.kind = kind; op.param = param; op.vol = scene_volume(s, dataset_id, dims, origin, spacing, type, voxels);
  s->svis.push_back(op);
  return (int)s->svis.size() - 1;
}

static VolumeData *scene_volume_impl(gxo_scene *s, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                                     int type, const void *voxels) {
  auto it = s->volumes.find(dataset_id);
  if (it == s->volumes.end()) {
    std::unique_ptr<VolumeData> v(new VolumeData());
    for (int k = 0; k < 3; k++) v->dims[k] = dims[k];
    v->origin = mk(origin[0], origin[1], origin[2]);
    v->spacing = mk(spacing[0], spacing[1], spacing[2]);
    v->rcp_spacing = mk(1.0f / spacing[0], 1.0f / spacing[1], 1.0f / spacing[2]);
    v->upper = mk(nextafterf((float)(dims[0] - 1), 0.f), nextafterf((float)(dims[1] - 1), 0.f), nextafterf((float)(dims[2] - 1), 0.f));
    v->type = type; v->voxels = voxels;
    v->samplingStep = min3(spacing[0], spacing[1], spacing[2]);
    v->samplingRate = 1.0f;
    v->tf = nullptr;
    it = s->volumes.emplace(dataset_id, std::move(v)).first;
  }
  return it->second.get();
}

int gxo_scene_add_triangles_vis(gxo_scene *s, int nv, const float *verts, const float *normals, const float *data, int nt,
                                const int *indices, const float *colors, const float *opacities, float lo, float hi) {
  GeomOp g; memset(&g, 0, sizeof g);
  g.kind = 0; g.nv = nv; g.nt = nt; g.verts = verts; g.normals = normals; g.data = data; g.idx = indices;
  set_tf(g.tf, colors, opacities, lo, hi);
  s->geoms.push_back(g);
  return (int)s->geoms.size() - 1;
}

int gxo_scene_add_particles_vis(gxo_scene *s, int n, const float *centers, const float *data, float radius0, float radius1,
                                float value0, float value1, const float *colors, const float *opacities, float lo, float hi) {
  GeomOp g; memset(&g, 0, sizeof g);
  g.kind = 1; g.n = n; g.centers = centers; g.data = data;
  g.radius0 = radius0; g.radius1 = radius1; g.value0 = value0; g.value1 = value1;
  // DataDrivenSpheres.ispc:211-220
  g.epsilon = logf(radius0);
  if (g.epsilon < 0.f) g.epsilon = -1.f / g.epsilon;
  if (g.epsilon > (float)(radius0 / 100.0)) g.epsilon = (float)(radius0 / 100.0);
  set_tf(g.tf, colors, opacities, lo, hi);
  s->geoms.push_back(g);
  return (int)s->geoms.size() - 1;
}

int gxo_scene_add_pathlines_vis(gxo_scene *s, int n_verts, const float *verts, const float *data, int n_segments,
                                const int *connectivity, float radius0, float radius1, float value0, float value1,
                                const float *colors, const float *opacities, float lo, float hi) {
  for (int i = 0; i < n_segments; i++)
    if (connectivity[i] < 0 || connectivity[i] + 1 >= n_verts) return -1;
  GeomOp g; memset(&g, 0, sizeof g);
  g.kind = 2; g.data = data;
  g.radius0 = radius0; g.radius1 = radius1; g.value0 = value0; g.value1 = value1;
  s->curve_store.emplace_back(new std::vector<float>());
  build_curves(n_segments, connectivity, verts, data, radius0, radius1, value0, value1, *s->curve_store.back());
  g.ncurves = n_segments; g.cp = s->curve_store.back()->data();
  set_tf(g.tf, colors, opacities, lo, hi);
  s->geoms.push_back(g);
  return (int)s->geoms.size() - 1;
}

int gxo_build_curves(int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity,
                     float radius0, float radius1, float value0, float value1, float *cp_out) {
  for (int i = 0; i < n_segments; i++)
    if (connectivity[i] < 0 || connectivity[i] + 1 >= n_verts) return -1;
  std::vector<float> cp;
  build_curves(n_segments, connectivity, verts, data, radius0, radius1, value0, value1, cp);
  memcpy(cp_out, cp.data(), cp.size() * sizeof(float));
  return 0;
}

int gxo_curve_intersect(int n_curves, const float *cp, int n_rays, const float *org3, const float *dir3, const float *tnear,
                        const float *tfar, int *prim_out, float *tu_out, float *ng_out, int per_curve) {
  for (int r = 0; r < n_rays; r++) {
    const V3 org = mk(org3[3 * r], org3[3 * r + 1], org3[3 * r + 2]), dir = mk(dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]);
    int best = -1; float bt = tfar[r], bu = 0.f; V3 bn = mk(0.f, 0.f, 0.f);
    for (int p = 0; p < n_curves; p++) {
      float t, u; V3 Ng;
      const bool h = curve_test(cp + 16 * (size_t)p, org, dir, tnear[r], tfar[r], t, u, Ng);
      if (per_curve) {
        const size_t o = (size_t)r * n_curves + p;
        prim_out[o] = h ? 1 : 0;
        tu_out[2 * o] = h ? t : tfar[r]; tu_out[2 * o + 1] = h ? u : 0.f;
        if (ng_out) { ng_out[3 * o] = h ? Ng.x : 0.f; ng_out[3 * o + 1] = h ? Ng.y : 0.f; ng_out[3 * o + 2] = h ? Ng.z : 0.f; }
      } else if (h && (best < 0 || t < bt)) { best = p; bt = t; bu = u; bn = Ng; }
    }
    if (!per_curve) {
      prim_out[r] = best; tu_out[2 * r] = bt; tu_out[2 * r + 1] = bu;
      if (ng_out) { ng_out[3 * r] = bn.x; ng_out[3 * r + 1] = bn.y; ng_out[3 * r + 2] = bn.z; }
    }
  }
  return 0;
}

int gxo_scene_commit(gxo_scene *s) {
  // MappedVis::SetTheOsprayDataObject (MappedVis.cpp:206-212): the volume object's TF is the
  // one of the last Vis committed on that dataset.
  for (VolumeVisOp &op : s->vvis) op.vol->tf = &op.tf;
  if (!s->ext_fn) build_accel(*s);
  return 0;
}

void gxo_scene_set_intersector(gxo_scene *s, gxo_intersect_fn fn, void *user) {
  s->ext_fn = fn;
  s->ext_user = user;
}

void gxo_resample_tf(int n, const float *cmap, int m, const float *omap, float *colors_out, float *opac_out) {
  // MappedVis::local_commit (MappedVis.cpp:277-338); note the double arithmetic in x.
  int i0 = 0, i1 = 1;
  float xmin = cmap[0], xmax = cmap[4 * (n - 1)];
  for (int i = 0; i < 256; i++) {
    float x = (float)(xmin + (i / (255.0)) * (xmax - xmin));
    if (x > xmax) x = xmax;
    while (cmap[4 * i1] < x) i0++, i1++;
    float d = (x - cmap[4 * i0]) / (cmap[4 * i1] - cmap[4 * i0]);
    colors_out[3 * i + 0] = cmap[4 * i0 + 1] + d * (cmap[4 * i1 + 1] - cmap[4 * i0 + 1]);
    colors_out[3 * i + 1] = cmap[4 * i0 + 2] + d * (cmap[4 * i1 + 2] - cmap[4 * i0 + 2]);
    colors_out[3 * i + 2] = cmap[4 * i0 + 3] + d * (cmap[4 * i1 + 3] - cmap[4 * i0 + 3]);
  }
  i0 = 0, i1 = 1;
  xmin = omap[0]; xmax = omap[2 * (m - 1)];
  for (int i = 0; i < 256; i++) {
    float x = (float)(xmin + (i / (255.0)) * (xmax - xmin));
    if (x > xmax) x = xmax;
    while (omap[2 * i1] < x) i0++, i1++;
    float d = (x - omap[2 * i0]) / (omap[2 * i1] - omap[2 * i0]);
    opac_out[i] = omap[2 * i0 + 1] + d * (omap[2 * i1 + 1] - omap[2 * i0 + 1]);
  }
}

void gxo_resolve_lights(const gxo_lighting *in, const gxo_camera *cam, gxo_lighting *out) {
  // Rendering::resolve_lights (Rendering.cpp:157-216)
  *out = *in;
  V3 viewpoint = mk(cam->eye[0], cam->eye[1], cam->eye[2]);
  V3 viewup = mk(cam->up[0], cam->up[1], cam->up[2]);
  V3 viewdir = mk(cam->dir[0], cam->dir[1], cam->dir[2]);
  normalize_gxy(viewup); normalize_gxy(viewdir);
  V3 right = cross(viewdir, viewup); normalize_gxy(right);
  V3 up = cross(viewdir, right);
  for (int i = 0; i < in->n_lights; i++) {
    if (in->types[i] == 1) {
      V3 p = viewpoint + in->lights[i][0] * right + in->lights[i][1] * up + in->lights[i][2] * viewdir;
      out->lights[i][0] = p.x; out->lights[i][1] = p.y; out->lights[i][2] = p.z;
      out->types[i] = 2;
    }
  }
}

long long gxo_trace_raylist(gxo_scene *s, const gxo_lighting *lights, float *base, int n, int aligned_n, float epsilon, int *hit_ids) {
  RL R = view(base, n, aligned_n);
  std::unique_ptr<OwnedRL> out = trace_and_spawn(*s, *lights, R, epsilon, 0, hit_ids, nullptr, nullptr);
  if (out) {
    s->secondary = std::move(out->mem); s->secondary_n = out->v.n; s->secondary_aligned = out->v.aligned;
  } else { s->secondary.clear(); s->secondary_n = 0; s->secondary_aligned = 0; }
  return s->secondary_n;
}

int gxo_fetch_secondary(gxo_scene *s, float *out_base, int aligned_n) {
  if (aligned_n < s->secondary_n) return -1;
  for (int k = 0; k < 25; k++)
    memcpy(out_base + (size_t)k * aligned_n, s->secondary.data() + (size_t)k * s->secondary_aligned, sizeof(float) * s->secondary_n);
  return s->secondary_n;
}

int gxo_classify(gxo_scene *s, float *base, int n, int aligned_n) {
  RL R = view(base, n, aligned_n);
  classify(*s, R);
  return 0;
}

int gxo_generate_rays(gxo_scene *s, const gxo_camera *cam, int w, int h, float *base, int aligned_n) {
  RL R = view(base, w * h, aligned_n);
  CamFrame a = camera_frame(*cam, w, h);
  return generate_range(*s, a, w, 0, w * h, R);
}

}  // extern "C"

// sampler = false: Renderer (frame into fb).  sampler = true: Sampler (src/sampler/Sampler.cpp): the same loop with
// Sampler::Trace -> SamplerTraceRays and Sampler::HandleTerminatedRays -> one particle per ray whose term has RAY_SURFACE;
// such a PRIMARY ray is KEEP_HERE (Renderer::Classify: neither BOUNDARY nor OPAQUE) and is traced again from its new t,
// so a ray leaves one sample per crossing until it reaches the boundary of the partition and moves on.
// kbuffer != NULL: the interactive / asynchronous frame path (Rendering.cpp:104-153 without GXY_WRITE_IMAGES): fb and kbuffer
// persist across frames, nothing is cleared up front; a batch of pixels of frame f is dropped when f < *rendering_frame
// (AddLocalPixels :140-152), else every contribution first resets its pixel if the pixel's stamp is older (ACCUMULATE_PIXEL).
static int render_impl(bool sampler, int nparts, gxo_scene **parts, const gxo_camera *cam, const gxo_lighting *lights_in, int w, int h,
                       float epsilon, int max_rays_per_packet, int nthreads, float *fb, gxo_stats *stats, int *kbuffer = nullptr,
                       int frame = 0, int *rendering_frame = nullptr) {
  gxo_stats st; memset(&st, 0, sizeof st);
  if (fb && !kbuffer) memset(fb, 0, sizeof(float) * 4 * (size_t)w * h);
  if (sampler) for (int p = 0; p < nparts; p++) parts[p]->samples.clear();
  gxo_lighting L;
  gxo_resolve_lights(lights_in, cam, &L);
  CamFrame a = camera_frame(*cam, w, h);
  if (max_rays_per_packet <= 0) max_rays_per_packet = 1000000;   // Renderer.cpp:134
  for (int p = 0; p < nparts; p++) parts[p]->sample_count = 0;

  std::deque<std::pair<int, std::unique_ptr<OwnedRL>>> q;

  // Camera::generate_initial_rays chunking (Camera.cpp:789-825): the window is always the full
  // image (the projected-bbox cull is dead code, Camera.cpp:643-653)
  for (int p = 0; p < nparts; p++) {
    int total = w * h;
    for (int i = 0; i < total; i += max_rays_per_packet) {
      int kthis = (i + max_rays_per_packet) > total ? total - i : max_rays_per_packet;
      std::unique_ptr<OwnedRL> rl(new OwnedRL(kthis, 0));
      // generation in parallel chunks, order-preserving
      int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
      if (nt < 1) nt = 1;
      std::vector<std::unique_ptr<OwnedRL>> pieces(nt);
      std::vector<int> counts(nt, 0);
      int chunk = (kthis + nt - 1) / nt;
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) {
        int s0 = t * chunk, c = std::min(chunk, kthis - s0);
        if (c <= 0) break;
        pieces[t].reset(new OwnedRL(c, 0));
        th.emplace_back([&, t, s0, c]() { counts[t] = generate_range(*parts[p], a, w, i + s0, c, pieces[t]->v); });
      }
      for (auto &t : th) t.join();
      int dst = 0;
      for (int t = 0; t < nt; t++) {
        if (!pieces[t]) continue;
        for (int k = 0; k < counts[t]; k++) copy_ray(pieces[t]->v, k, rl->v, dst++);
      }
      if (dst) { rl->v.n = dst; st.primary_rays += dst; q.emplace_back(p, std::move(rl)); }
    }
  }
  // orphan pixels: hit the global box but claimed by no partition
  {
    std::vector<unsigned char> claimed((size_t)w * h, 0), ghit((size_t)w * h, 0);
    for (int p = 0; p < nparts; p++)
      for (int pix = 0; pix < w * h; pix++) {
        V3 o, d; bool gh;
        if (spawn_pixel(*parts[p], a, pix % w, pix / w, o, d, &gh)) claimed[pix]++;
        if (gh) ghit[pix] = 1;
      }
    for (int pix = 0; pix < w * h; pix++) if (ghit[pix] && !claimed[pix]) st.orphan_pixels++;
  }

  // RayQManager FIFO + processRays_task::work (Renderer.cpp:535-643)
  while (!q.empty()) {
    int p = q.front().first;
    std::unique_ptr<OwnedRL> rl = std::move(q.front().second);
    q.pop_front();
    st.waves++;
    gxo_scene &S = *parts[p];
    while (rl) {
      RL &R = rl->v;
      st.traced_rays += R.n;
      long long nao = 0, nsh = 0;
      std::unique_ptr<OwnedRL> out;
      if (sampler) sampler_trace(S, R, nthreads);
      else out = trace_and_spawn(S, L, R, epsilon, nthreads, nullptr, &nao, &nsh);
      st.ao_rays += nao; st.shadow_rays += nsh;
      if (out) {
        // Renderer::Trace (Renderer.cpp:658-691): split to <= max list size, enqueue locally
        int n = out->v.n;
        if (n > max_rays_per_packet) {
          for (int s0 = 0; s0 < n; s0 += max_rays_per_packet) {
            int c = std::min(max_rays_per_packet, n - s0);
            std::unique_ptr<OwnedRL> piece(new OwnedRL(c, 1));
            for (int k = 0; k < c; k++) copy_ray(out->v, s0 + k, piece->v, k);
            q.emplace_back(p, std::move(piece));
          }
        } else q.emplace_back(p, std::move(out));
      }
      classify(S, R);
      if (sampler)   // Sampler::HandleTerminatedRays (Sampler.cpp:52-92): every ray of the list whose term has RAY_SURFACE
        for (int i = 0; i < R.n; i++)
          if (R.term[i] & RAY_SURFACE) {
            S.samples.push_back(R.ox[i] + R.t[i] * R.dx[i]);
            S.samples.push_back(R.oy[i] + R.t[i] * R.dy[i]);
            S.samples.push_back(R.oz[i] + R.t[i] * R.dz[i]);
          }
      // HandleTerminatedRays (Renderer.cpp:456-502) + AddLocalPixels (Rendering.cpp:125-153)
      std::vector<int> knts(nparts, 0); int nKeepers = 0;
      for (int i = 0; i < R.n; i++) {
        int c = R.classification[i];
        if (c == TERMINATED && !sampler) {
          const size_t offset = (size_t)R.y[i] * w + R.x[i];
          float *ptr = fb + (offset << 2);
          if (kbuffer) {
            if (!(frame >= *rendering_frame)) continue;          // a stale frame's pixels are dropped (:140)
            if (frame > *rendering_frame) *rendering_frame = frame;
            if (kbuffer[offset] < frame) { ptr[0] = ptr[1] = ptr[2] = ptr[3] = 0.f; kbuffer[offset] = frame; }
          }
          ptr[0] += R.r[i]; ptr[1] += R.g[i]; ptr[2] += R.b[i]; ptr[3] += R.o[i];
          st.terminated_rays++;
        } else if (c == KEEP_HERE) nKeepers++;
        else if (c >= 0 && c < nparts) knts[c]++;
      }
      std::vector<std::unique_ptr<OwnedRL>> lists(nparts);
      std::unique_ptr<OwnedRL> keepers;
      if (nKeepers) keepers.reset(new OwnedRL(nKeepers, rl->kind));
      for (int d = 0; d < nparts; d++) if (knts[d]) { lists[d].reset(new OwnedRL(knts[d], rl->kind)); knts[d] = 0; }
      nKeepers = 0;
      for (int i = 0; i < R.n; i++) {
        int c = R.classification[i];
        if (c >= 0 && c < nparts) copy_ray(R, i, lists[c]->v, knts[c]++);
        else if (c == KEEP_HERE) copy_ray(R, i, keepers->v, nKeepers++);
      }
      for (int d = 0; d < nparts; d++) if (lists[d]) { st.forwarded_rays += lists[d]->v.n; q.emplace_back(d, std::move(lists[d])); }
      rl = std::move(keepers);
    }
  }
  for (int p = 0; p < nparts; p++) st.volume_samples += parts[p]->sample_count.load();
  if (stats) *stats = st;
  return 0;
}

extern "C" {

int gxo_render(int nparts, gxo_scene **parts, const gxo_camera *cam, const gxo_lighting *lights_in, int w, int h, float epsilon,
               int max_rays_per_packet, int nthreads, float *fb, gxo_stats *stats) {
  return render_impl(false, nparts, parts, cam, lights_in, w, h, epsilon, max_rays_per_packet, nthreads, fb, stats);
}

int gxo_render_progressive(int nparts, gxo_scene **parts, const gxo_camera *cam, const gxo_lighting *lights_in, int w, int h, float epsilon,
                           int nthreads, int frame, float *fb_inout, int *kbuffer_inout, int *rendering_frame_inout, gxo_stats *stats) {
  return render_impl(false, nparts, parts, cam, lights_in, w, h, epsilon, 0, nthreads, fb_inout, stats, kbuffer_inout, frame,
                     rendering_frame_inout);
}

int gxo_sample(int nparts, gxo_scene **parts, const gxo_camera *cam, int w, int h, int max_rays_per_packet, int nthreads,
               gxo_stats *stats) {
  gxo_lighting L; memset(&L, 0, sizeof L);
  L.n_lights = 1; L.lights[0][0] = L.lights[0][1] = L.lights[0][2] = 1.f; L.types[0] = 2; L.Ka = L.Kd = 0.5f;   // unused by the sampler
  return render_impl(true, nparts, parts, cam, &L, w, h, 0.001f, max_rays_per_packet, nthreads, nullptr, stats);
}

long long gxo_scene_samples(gxo_scene *s, const float **xyz) {
  if (xyz) *xyz = s->samples.data();
  return (long long)(s->samples.size() / 3);
}

int gxo_sample_raylist(gxo_scene *s, float *base, int n, int aligned_n) {
  RL R = view(base, n, aligned_n);
  sampler_trace(*s, R, 0);
  return 0;
}

void gxo_fb_to_rgba8(const float *fb, int w, int h, unsigned char *out) {
  // ImageWriter.cpp:30-48; (unsigned char)(255*f) is UB out of range in C++; defined here as
  // x86 cvttss2si (0x80000000 on overflow/NaN) followed by truncation to the low byte.
  auto cvt = [](float f) -> unsigned char {
    float v = 255 * f;
    int32_t iv;
    if (!(v > -2147483904.0f && v < 2147483648.0f)) iv = INT32_MIN; else iv = (int32_t)v;
    return (unsigned char)(iv & 0xff);
  };
  const float *p = fb;
  for (int y = 0; y < h; y++) {
    unsigned char *b = out + (size_t)((h - 1) - y) * w * 4;
    for (int x = 0; x < w; x++) {
      *b++ = cvt(*p++); *b++ = cvt(*p++); *b++ = cvt(*p++); *b++ = 0xff; p++;
    }
  }
}

void gxo_factor(int ijk, int factors[3]) {
  // Volume.cpp:88-122
  if (ijk == 1) { factors[0] = factors[1] = factors[2] = 1; return; }
  int mm = ijk + 3;
  for (int i = 1; i <= ijk >> 1; i++) {
    int jk = ijk / i;
    if (ijk == (i * jk))
      for (int j = 1; j <= jk >> 1; j++) {
        int k = jk / j;
        if (jk == (j * k)) {
          int m = i + j + k;
          if (m < mm) { mm = m; factors[0] = i; factors[1] = j; factors[2] = k; }
        }
      }
  }
}

void gxo_partition(int n, const int factors[3], const int grid[3], int *out) {
  // Volume.cpp:133-172
  int ni = grid[0] - 2, nj = grid[1] - 2, nk = grid[2] - 2;
  int di = ni / factors[0], dj = nj / factors[1], dk = nk / factors[2];
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
  (void)n;
}

void gxo_set_option(const char *name, int value) {
  if (!strcmp(name, "dvr_before_iso")) g_dvr_before_iso = value;
}

int gxo_intersect(gxo_scene *s, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar,
                  int *geom_prim2, float *tuv3) {
  parallel_for(n, 0, [&](int a, int b) {
    for (int i = a; i < b; i++) {
      Hit1 h;
      V3 o = mk(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2]), d = mk(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]);
      bool f = nearest_hit(*s, o, d, tnear[i], tfar[i], h);
      geom_prim2[2 * i] = f ? h.geomID : -1; geom_prim2[2 * i + 1] = f ? h.primID : -1;
      tuv3[3 * i] = f ? h.t : tfar[i]; tuv3[3 * i + 1] = f ? h.u : 0.f; tuv3[3 * i + 2] = f ? h.v : 0.f;
    }
  });
  return 0;
}

/* the box helpers on their own, so that they can be pinned against the reference's compiled Box.cpp (oracle/box_ref.cpp,
 * tests/test_oracle_box.py): boxes6 = n x (minx miny minz maxx maxy maxz), rays6 = n x (x y z dx dy dz) */
void gxo_exit_face(int n, const float *boxes6, const float *rays6, int *faces) {
  for (int i = 0; i < n; i++) {
    const float *b = boxes6 + 6 * i, *r = rays6 + 6 * i;
    faces[i] = exit_face(mk(b[0], b[1], b[2]), mk(b[3], b[4], b[5]), r[0], r[1], r[2], r[3], r[4], r[5]);
  }
}
void gxo_box_intersect(int n, const float *boxes6, const float *rays6, int *hit, float *t2) {
  for (int i = 0; i < n; i++) {
    const float *b = boxes6 + 6 * i, *r = rays6 + 6 * i;
    float tmin = 0.f, tmax = 0.f;
    hit[i] = box_intersect(mk(b[0], b[1], b[2]), mk(b[3], b[4], b[5]), mk(r[0], r[1], r[2]), mk(r[3], r[4], r[5]), tmin, tmax) ? 1 : 0;
    t2[2 * i] = tmin; t2[2 * i + 1] = tmax;
  }
}

}  // extern "C"
