// gxy_curve.cuh -- ray / round cubic Bezier segment: the primitive test of PathLines.
//
// Replaces what rtcIntersectV runs for RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE (src/ospray/DataDrivenPathLines.ispc:
// 319-324): Embree 3.6.1's sweep intersector, embree/kernels/geometry/curve_intersector_sweep.h
//   SweepCurve1Intersector1::intersect                   :224-241  -> curve_test
//   intersect_bezier_recursive_jacobian                  :122-221  -> curve_level_setup + the loop in curve_test
//   intersect_bezier_iterative_jacobian                  : 69-120  -> curve_newton
//   CylinderN::intersect (cylinder.h:162-225), HalfPlaneN::intersect (plane.h:60-72)
// for the AVX/AVX2 build Galaxy ships (8 SIMD lanes = 7 sub-segments per level, numBezierSubdivisions = 2, a third
// level where the inner cylinder is missed or grazed).
//
// Embree walks the levels by recursion over 8-wide SIMD values; here one thread walks them with an explicit
// three-entry level array in local memory and plain loops over the 7 sub-segments.  The arithmetic keeps
// Embree's association order: its madd/msub (FMA in the AVX2 build) are fmaf, its rcp/rsqrt (rcpss/rsqrtss +
// one Newton step, ISA dependent) are IEEE divide / 1/sqrt, Vec3fa dot products are (xx+yy)+zz as _mm_dp_ps
// adds them, SIMD min/max are `a<b?a:b` / `a>b?a:b` as the instructions define them for NaNs.
//
// The file is self-contained (no CUDA types) and compiles as plain C++ as well, so that tests/ can run THIS
// source on the CPU against the oracle without a GPU (tests/test_curve_host.py); the product only ever uses
// the device build.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GXC_FN __device__ __forceinline__
#define GXC_NOINLINE static __device__ __noinline__
#else
#define GXC_FN static inline
#define GXC_NOINLINE static
#endif

namespace gxc {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct CurveHit { float t, u; V3 Ng; };

GXC_FN V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GXC_FN V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GXC_FN V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
GXC_FN V3 xyz(V4 a) { return v3(a.x, a.y, a.z); }
GXC_FN float rcp_ieee(float x) { return 1.0f / x; }
GXC_FN float rsqrt_ieee(float x) { return 1.0f / sqrtf(x); }
GXC_FN float smin(float a, float b) { return a < b ? a : b; }   // minps(a,b)
GXC_FN float smax(float a, float b) { return a > b ? a : b; }   // maxps(a,b)
// Vec3fa dot: _mm_dp_ps(a,b,0x7F) (common/math/vec3fa.h:289)
GXC_FN float dot_a(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// Vec3<vfloat> dot / cross (common/math/vec3.h:216,221)
GXC_FN float dot_v(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
GXC_FN V3 cross_v(V3 a, V3 b) {
  return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
// lerp(v0,v1,t) = madd(1-t, v0, t*v1) (vec4.h:191, vec3fa.h:341)
GXC_FN V4 lerp4(V4 a, V4 b, float t) {
  const float s = 1.0f - t;
  V4 r;
  r.x = fmaf(s, a.x, t * b.x); r.y = fmaf(s, a.y, t * b.y); r.z = fmaf(s, a.z, t * b.z); r.w = fmaf(s, a.w, t * b.w);
  return r;
}
// CubicBezierCurve::eval / veval: position and first derivative by de Casteljau (bezier_curve.h:322-339, 476-492)
GXC_FN void bezier_eval(const V4 *cp, float t, V4 &p, V4 &dp) {
  const V4 p10 = lerp4(cp[0], cp[1], t), p11 = lerp4(cp[1], cp[2], t), p12 = lerp4(cp[2], cp[3], t);
  const V4 p20 = lerp4(p10, p11, t), p21 = lerp4(p11, p12, t);
  p = lerp4(p20, p21, t);
  dp.x = 3.0f * (p21.x - p20.x); dp.y = 3.0f * (p21.y - p20.y); dp.z = 3.0f * (p21.z - p20.z); dp.w = 3.0f * (p21.w - p20.w);
}
// eval_dudu (bezier_curve.h:382-386) over BezierBasis::derivative2 (:52-61); only xyz is used
GXC_FN V3 bezier_dudu(const V4 *cp, float t1) {
  const float t0 = 1.0f - t1;
  const float b0 = 6.0f * t0, b1 = 6.0f * fmaf(-2.0f, t0, t1), b2 = 6.0f * fmaf(-2.0f, t1, t0), b3 = 6.0f * t1;
  return v3(fmaf(b0, cp[0].x, fmaf(b1, cp[1].x, fmaf(b2, cp[2].x, b3 * cp[3].x))),
            fmaf(b0, cp[0].y, fmaf(b1, cp[1].y, fmaf(b2, cp[2].y, b3 * cp[3].y))),
            fmaf(b0, cp[0].z, fmaf(b1, cp[1].z, fmaf(b2, cp[2].z, b3 * cp[3].z))));
}

#define GXC_ULP 1.1920929e-07f   // embree `ulp` = numeric_limits<float>::epsilon()
#define GXC_INF (__builtin_huge_valf())

// the ray in the frame shifted by dt along it (origin 0) and the running result; tfar shrinks with every hit
// exactly as the epilog does (Intersect1Epilog1 without filters, intersector_epilog.h:72-80)
struct CurveRay {
  V3 dir;
  float tnear, tfar, dt, len_dir;
  CurveHit hit;
};

// Newton iteration on (u,t) from a cylinder hit (:69-120)
GXC_FN bool curve_newton(CurveRay &ray, const V4 *cp, float u, float t) {
  const V3 dir = ray.dir;
  for (int it = 0; it < 5; it++) {
    const V3 Q = v3(fmaf(t, dir.x, 0.0f), fmaf(t, dir.y, 0.0f), fmaf(t, dir.z, 0.0f));
    V4 P4, dP4;
    bezier_eval(cp, u, P4, dP4);
    const V3 P = xyz(P4), dPdu = xyz(dP4), ddPdu = bezier_dudu(cp, u);
    const V3 R = sub(Q, P);
    const V3 dRdu = v3(-dPdu.x, -dPdu.y, -dPdu.z);
    const float pp = dot_a(dPdu, dPdu), pdp = dot_a(dPdu, ddPdu);
    const V3 T = scale(dPdu, rsqrt_ieee(pp));
    // dnormalize (vec3fa.h:321-326): (pp*dp - pdp*p) * rcp(pp) * rsqrt(pp)
    const V3 dTdu = scale(scale(sub(scale(ddPdu, pp), scale(dPdu, pdp)), rcp_ieee(pp)), rsqrt_ieee(pp));
    const float f = dot_a(R, T);
    const float dfdu = dot_a(dRdu, T) + dot_a(R, dTdu);
    const float dfdt = dot_a(dir, T);
    const float K = dot_a(R, R) - f * f;
    const float dKdu = dot_a(R, dRdu) - f * dfdu;
    const float dKdt = dot_a(R, dir) - f * dfdt;
    const float rsqrt_K = rsqrt_ieee(K);
    const float g = sqrtf(K) - P4.w;
    const float dgdu = dKdu * rsqrt_K - dP4.w;
    const float dgdt = dKdt * rsqrt_K;
    // rcp(LinearSpace2f(dfdu,dfdt,dgdu,dgdt)) * Vec2f(f,g): adjoint()/det() (linearspace2.h:49-55), then b.x*vx + b.y*vy
    const float det = dfdu * dgdt - dgdu * dfdt;
    const float du = f * (dgdt / det) + g * (-dfdt / det);
    const float dtt = f * (-dgdu / det) + g * (dfdu / det);
    u = u - du;
    t = t - dtt;
    const bool converged_u = fabsf(f) < 16.0f * GXC_ULP * smax(smax(fabsf(dPdu.x), fabsf(dPdu.y)), fabsf(dPdu.z));
    const bool converged_t = fabsf(g) < 16.0f * GXC_ULP * ray.len_dir;
    if (converged_u && converged_t) {
      t += ray.dt;
      if (!(t > ray.tnear && t < ray.tfar)) return false;   // open interval; rejects NaNs
      if (!(u >= 0.0f && u <= 1.0f)) return false;
      const V3 Rn = scale(R, rsqrt_ieee(dot_a(R, R)));
      const V3 U = v3(fmaf(dP4.w, Rn.x, dPdu.x), fmaf(dP4.w, Rn.y, dPdu.y), fmaf(dP4.w, Rn.z, dPdu.z));
      const V3 V = cross_v(dPdu, Rn);
      ray.tfar = t;
      ray.hit.t = t; ray.hit.u = u; ray.hit.Ng = cross_v(V, U);
      return true;
    }
  }
  return false;
}

// one infinite cylinder (p0,p1,r) against the ray from the origin (cylinder.h:162-225, one lane)
struct CylHit { float lo, up, u0, u1; V3 Ng0, Ng1; bool valid; };
GXC_FN void cylinder_hit(V3 p0, V3 p1, float r, V3 dir, CylHit &h) {
  const float rr = r * r;
  const V3 ax = sub(p1, p0);
  const float rl = rsqrt_ieee(dot_v(ax, ax));
  const V3 dP = scale(ax, rl);
  const V3 O = sub(v3(0.0f, 0.0f, 0.0f), p0);
  const float dOdO = dot_v(dir, dir), OdO = dot_v(dir, O), OO = dot_v(O, O), dOz = dot_v(dP, dir), Oz = dot_v(dP, O);
  const float A = dOdO - dOz * dOz;
  const float B = 2.0f * (OdO - dOz * Oz);
  const float C = OO - Oz * Oz - rr;
  const float D = B * B - 4.0f * A * C;
  bool valid = D >= 0.0f;
  const float Q = sqrtf(D);
  const float rcp_2A = rcp_ieee(2.0f * A);
  const float t0 = (-B - Q) * rcp_2A, t1 = (-B + Q) * rcp_2A;
  h.u0 = fmaf(t0, dOz, Oz) * rl;
  h.Ng0 = sub(scale(dir, t0), v3(fmaf(h.u0, ax.x, p0.x), fmaf(h.u0, ax.y, p0.y), fmaf(h.u0, ax.z, p0.z)));
  h.u1 = fmaf(t1, dOz, Oz) * rl;
  h.Ng1 = sub(scale(dir, t1), v3(fmaf(h.u1, ax.x, p0.x), fmaf(h.u1, ax.y, p0.y), fmaf(h.u1, ax.z, p0.z)));
  h.lo = valid ? t0 : GXC_INF;
  h.up = valid ? t1 : -GXC_INF;
  const float eps = 16.0f * GXC_ULP * smax(fabsf(dOdO), fabsf(dOz * dOz));
  if (valid && fabsf(A) < eps) {   // ray parallel to the axis: inside -> everything, outside -> nothing
    const bool inside = C <= 0.0f;
    h.lo = inside ? -GXC_INF : GXC_INF;
    h.up = inside ? GXC_INF : -GXC_INF;
    valid = inside;
  }
  h.valid = valid;
}
// half space {x : (x-P).N >= 0} against the ray from the origin (plane.h:60-72, one lane)
GXC_FN void halfplane_clip(V3 P, V3 N, V3 dir, float &lo, float &up) {
  const V3 O = sub(v3(0.0f, 0.0f, 0.0f), P);
  const float ON = dot_v(O, N), DN = dot_v(dir, N);
  const bool eps = fabsf(DN) < 1E-18f;   // min_rcp_input
  const float t = -ON * rcp_ieee(DN);
  lo = smax(lo, (eps || DN < 0.0f) ? -GXC_INF : t);
  up = smin(up, (eps || DN > 0.0f) ? GXC_INF : t);
}

// what one call of intersect_bezier_recursive_jacobian keeps across its two hit loops
struct CurveLevel {
  float vu[8];                  // parameter of the 8 sub-division points
  float t0_lo[7];               // first-hit intervals: lower end (the start value of t)
  float t1_lo[7], t1_up[7];     // second-hit intervals
  float uo0[7], uo1[7];         // start values of u
  unsigned valid0, valid1, unstable0, unstable1;
  int phase;
};

// select_min(valid, v) (simd/vfloat8_avx.h:669-674): the lowest lane holding the minimum, else the lowest valid lane
GXC_FN int select_min7(unsigned valid, const float *v) {
  float m = GXC_INF;
  for (int i = 0; i < 7; i++)
    if ((valid >> i) & 1u) m = smin(v[i], m);
  int first_valid = -1;
  for (int i = 0; i < 7; i++)
    if ((valid >> i) & 1u) {
      if (v[i] == m) return i;
      if (first_valid < 0) first_valid = i;
    }
  return first_valid;
}
// valid &= (lower + dt <= ray.tfar)
GXC_FN unsigned prune7(unsigned valid, const float *lower, float dt, float tfar) {
  for (int i = 0; i < 7; i++)
    if (!(lower[i] + dt <= tfar)) valid &= ~(1u << i);
  return valid;
}

// the first half of intersect_bezier_recursive_jacobian (:131-196): bounding cylinders and cap planes of the 7
// sub-segments of [u0,u1].  false: nothing to visit at this level.
GXC_FN bool curve_level_setup(const CurveRay &ray, const V4 *cp, float u0, float u1, CurveLevel &L) {
  const V3 dir = ray.dir;
  const float dscale = (u1 - u0) * (1.0f / (3.0f * 7));
  const V3 ndir = scale(dir, rsqrt_ieee(dot_a(dir, dir)));   // normalize(ray.dir) on Vec3fa
  L.valid0 = L.valid1 = L.unstable0 = L.unstable1 = 0u;
  L.phase = 0;
  unsigned valid = 0u;
  V4 Pa, dPa;
  L.vu[0] = fmaf(0.0f, u1 - u0, u0);   // lerp(u0,u1,step*(1/7)) = madd(t, b-a, a) (vfloat8_avx.h:477)
  bezier_eval(cp, L.vu[0], Pa, dPa);
  dPa.x *= dscale; dPa.y *= dscale; dPa.z *= dscale; dPa.w *= dscale;
  for (int i = 0; i < 7; i++) {
    V4 Pb, dPb;
    L.vu[i + 1] = fmaf((float)(i + 1) * (1.0f / 7), u1 - u0, u0);
    bezier_eval(cp, L.vu[i + 1], Pb, dPb);
    dPb.x *= dscale; dPb.y *= dscale; dPb.z *= dscale; dPb.w *= dscale;
    // control radii of the sub-segment: P0.w, P1.w = P0.w + dP0.w, P2.w = P3.w - dP3.w, P3.w
    const float w1 = Pa.w + dPa.w, w2 = Pb.w - dPb.w;
    const V3 p0 = xyz(Pa), p3 = xyz(Pb), d0 = xyz(dPa), d3 = xyz(dPb), chord = sub(p3, p0);
    // sqr_point_to_line_distance(PmQ0, Q1mQ0) (vec3.h:253-258) of the two inner control points from the chord
    const V3 n1 = cross_v(d0, chord), n2 = cross_v(d3, chord);
    const float rcd = rcp_ieee(dot_v(chord, chord));
    const float maxr12 = sqrtf(smax(dot_v(n1, n1) * rcd, dot_v(n2, n2) * rcd));
    float r_outer = smax(smax(Pa.w, w1), smax(w2, Pb.w)) + maxr12;
    float r_inner = smin(smin(Pa.w, w1), smin(w2, Pb.w)) - maxr12;
    r_outer = (1.0f + 2.0f * GXC_ULP) * r_outer;
    r_inner = smax(0.0f, (1.0f - 2.0f * GXC_ULP) * r_inner);
    CylHit co, ci;
    cylinder_hit(p0, p3, r_outer, dir, co);
    float lo = smax(ray.tnear - ray.dt, co.lo), up = smin(ray.tfar - ray.dt, co.up);
    halfplane_clip(p0, d0, dir, lo, up);
    halfplane_clip(p3, v3(-d3.x, -d3.y, -d3.z), dir, lo, up);
    const bool v = co.valid && (lo <= up);
    if (!v) {
      // Embree computes the inner cylinder for all 8 lanes at once; a lane the outer cylinder or the cap planes reject can set
      // neither valid0 nor valid1, and only those lanes' u, interval and stability values are ever read: skip the rest
      L.t0_lo[i] = L.t1_lo[i] = GXC_INF;
      L.t1_up[i] = -GXC_INF;
      L.uo0[i] = L.uo1[i] = 0.0f;
      Pa = Pb; dPa = dPb;
      continue;
    }
    const float c0 = smin(smax(co.u0, 0.0f), 1.0f), c1 = smin(smax(co.u1, 0.0f), 1.0f);
    L.uo0[i] = fmaf(((float)i + c0) * (1.0f / 8.0f), u1 - u0, u0);   // Embree's own (step+u)*(1/VSIZEX)
    L.uo1[i] = fmaf(((float)i + c1) * (1.0f / 8.0f), u1 - u0, u0);
    cylinder_hit(p0, p3, r_inner, dir, ci);
    const V3 nn0 = scale(ci.Ng0, rsqrt_ieee(dot_v(ci.Ng0, ci.Ng0))), nn1 = scale(ci.Ng1, rsqrt_ieee(dot_v(ci.Ng1, ci.Ng1)));
    const bool un0 = !ci.valid || (fabsf(dot_v(ndir, nn0)) < 0.3f);
    const bool un1 = !ci.valid || (fabsf(dot_v(ndir, nn1)) < 0.3f);
    // subtract(tp, tc_inner, tp0, tp1) (bbox.h:166-172)
    L.t0_lo[i] = lo;
    const float t0_up = smin(up, ci.lo);
    L.t1_lo[i] = smax(lo, ci.up);
    L.t1_up[i] = up;
    if (v) valid |= 1u << i;
    if (v && L.t0_lo[i] <= t0_up) L.valid0 |= 1u << i;
    if (v && L.t1_lo[i] <= L.t1_up[i]) L.valid1 |= 1u << i;
    if (un0) L.unstable0 |= 1u << i;
    if (un1) L.unstable1 |= 1u << i;
    Pa = Pb; dPa = dPb;
  }
  return valid != 0u && (L.valid0 | L.valid1) != 0u;
}

// ---- conservative cull in front of the test (ours, not Embree's: it only ever rejects rays the test would reject) ----
// Every point of the swept surface lies within `bound` of the LINE through the segment's end points: the centre curve
// stays in the convex hull of the control points (distance to a line is convex, the two inner control points are at most
// rho from it) and the radius is a Bernstein combination of the control radii (<= max |w|).  Computed once per segment
// when the BVH records are written; < 0 = no cull (degenerate chord).
GXC_FN float curve_bound_radius(const float *cp16) {
  const V3 p0 = v3(cp16[0], cp16[1], cp16[2]), p3 = v3(cp16[12], cp16[13], cp16[14]);
  const V3 chord = sub(p3, p0);
  const float cc = dot_a(chord, chord);
  if (!(cc > 1e-30f)) return -1.0f;
  const V3 n1 = cross_v(sub(v3(cp16[4], cp16[5], cp16[6]), p0), chord), n2 = cross_v(sub(v3(cp16[8], cp16[9], cp16[10]), p0), chord);
  const float rho = sqrtf(smax(dot_a(n1, n1), dot_a(n2, n2)) / cc);
  const float wmax = smax(smax(fabsf(cp16[3]), fabsf(cp16[7])), smax(fabsf(cp16[11]), fabsf(cp16[15])));
  const float b = (rho + wmax) * 1.004f + 1e-6f * (sqrtf(dot_a(p0, p0)) + sqrtf(cc));
  return (b == b && b < GXC_INF) ? b : -1.0f;
}
// true: the ray's line passes farther than `bound` from the chord's line, so the ray cannot touch the segment.
// Margins: 0.4 % of the bound at build time, here the rounding of (p0-org).n against |p0-org| and a refusal to decide
// for nearly parallel lines (sin^2 < 1e-6, where dir x chord cancels); a hit reported by curve_test lies on the surface
// to within 16 ulp of |dir| (:104-105), far inside these margins.
GXC_FN bool curve_precull(V3 p0, V3 p3, float bound, V3 org, V3 dir) {
  if (!(bound >= 0.0f)) return false;
  const V3 chord = sub(p3, p0);
  const V3 n = cross_v(dir, chord);
  const float nn = dot_a(n, n), dd = dot_a(dir, dir), cc = dot_a(chord, chord);
  if (!(nn > 1e-6f * dd * cc)) return false;
  const V3 po = sub(p0, org);
  const float s = dot_a(po, n);
  const float tol = bound + 4e-6f * sqrtf(dot_a(po, po));
  return s * s > tol * tol * nn;
}

// Nearest hit of ONE segment (cp = 4 control points x,y,z,r) in the open interval (tnear, tfar).
GXC_NOINLINE bool curve_test(const float *cp16, float ox, float oy, float oz, float dx, float dy, float dz, float tnear, float tfar,
                             CurveHit &out) {
  V4 cp[4];
  for (int k = 0; k < 4; k++) { cp[k].x = cp16[4 * k]; cp[k].y = cp16[4 * k + 1]; cp[k].z = cp16[4 * k + 2]; cp[k].w = cp16[4 * k + 3]; }
  CurveRay ray;
  ray.dir = v3(dx, dy, dz);
  ray.tnear = tnear; ray.tfar = tfar;
  // move the ray closer to the curve (:233-238): centre = 0.25*(v0+v1+v2+v3)
  const V3 c = v3(0.25f * (((cp[0].x + cp[1].x) + cp[2].x) + cp[3].x), 0.25f * (((cp[0].y + cp[1].y) + cp[2].y) + cp[3].y),
                  0.25f * (((cp[0].z + cp[1].z) + cp[2].z) + cp[3].z));
  const float dd = dot_a(ray.dir, ray.dir);
  ray.dt = dot_a(sub(c, v3(ox, oy, oz)), ray.dir) * rcp_ieee(dd);
  ray.len_dir = sqrtf(dd);
  const V3 ref = v3(fmaf(ray.dt, dx, ox), fmaf(ray.dt, dy, oy), fmaf(ray.dt, dz, oz));
  for (int k = 0; k < 4; k++) { cp[k].x -= ref.x; cp[k].y -= ref.y; cp[k].z -= ref.z; }
  ray.hit.t = tfar; ray.hit.u = 0.0f; ray.hit.Ng = v3(0.0f, 0.0f, 0.0f);

  // the recursion of :198-219 with an explicit level array: level d is Embree's depth d+1; a sub-segment is
  // refined while depth < termDepth (2, or 3 where its inner cylinder is missed or grazed), else iterated
  CurveLevel lv[3];
  bool found = false;
  int d = 0;
  if (!curve_level_setup(ray, cp, 0.0f, 1.0f, lv[0])) return false;
  while (d >= 0) {
    CurveLevel &L = lv[d];
    const bool second = L.phase != 0;
    unsigned &valid = second ? L.valid1 : L.valid0;
    const float *lower = second ? L.t1_lo : L.t0_lo;
    if (valid == 0u) {
      if (!second) {   // first hits done: the second hits start with a prune (:209)
        L.valid1 = prune7(L.valid1, L.t1_lo, ray.dt, ray.tfar);
        L.phase = 1;
      } else {         // level done: back in the parent's loop, whose iteration ends with a prune (:207,:218)
        d--;
        if (d >= 0) {
          CurveLevel &Pl = lv[d];
          if (Pl.phase == 0) Pl.valid0 = prune7(Pl.valid0, Pl.t0_lo, ray.dt, ray.tfar);
          else Pl.valid1 = prune7(Pl.valid1, Pl.t1_lo, ray.dt, ray.tfar);
        }
      }
      continue;
    }
    const int i = select_min7(valid, lower);
    valid &= ~(1u << i);
    const int termDepth = (((second ? L.unstable1 : L.unstable0) >> i) & 1u) ? 3 : 2;
    if (d + 1 >= termDepth) {
      if (curve_newton(ray, cp, second ? L.uo1[i] : L.uo0[i], second ? L.t1_up[i] : L.t0_lo[i])) found = true;
      valid = prune7(valid, lower, ray.dt, ray.tfar);
    } else if (curve_level_setup(ray, cp, L.vu[i], L.vu[i + 1], lv[d + 1])) {
      d++;
    } else {
      valid = prune7(valid, lower, ray.dt, ray.tfar);
    }
  }
  if (found) out = ray.hit;
  return found;
}

}  // namespace gxc
