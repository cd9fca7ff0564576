// gxy_traverse.cuh -- nearest-hit traversal of the 8-wide quantised BVH + the two primitive tests.
// Replaces rtcIntersectV (ospray/common/Model.ih:54-70) i.e. Embree's BVH8 traversal with
// MoellerTrumboreIntersectorK (embree/kernels/geometry/triangle_intersector_moeller.h:210-266)
// and the DataDrivenSpheres user-geometry callback (src/ospray/DataDrivenSpheres.ispc:90-155).
#pragma once
#include "gxy_common.cuh"

namespace gxy {

struct Hit1 {
  float t, u, v;
  int geom, prim;
  float3 Ng;
};

#define GXY_STACK_SMEM 24   // entries per thread kept in shared memory
#define GXY_STACK_LOCAL 40  // overflow entries in local memory
#define GXY_TRACE_THREADS 128

// Embree AVX2 op order (common/math/vec3.h:216,221): dot = madd(x,x, madd(y,y, z*z)),
// cross.x = msub(a.y,b.z, a.z*b.y)
__device__ __forceinline__ float edot(float3 a, float3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z))); }
__device__ __forceinline__ float3 ecross(float3 a, float3 b) {
  return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
            __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// one triangle record against the ray; candidate interval is the ORIGINAL (tnear, tfar]
__device__ __forceinline__ bool tri_test(const float4 ra, const float4 rb, const float4 rc, float3 org, float3 dir, float tnear,
                                         float tfar, float &t, float &u, float &v, float3 &Ng) {
  const float3 v0 = f3(ra.x, ra.y, ra.z), e1 = f3(ra.w, rb.x, rb.y), e2 = f3(rb.z, rb.w, rc.x);
  Ng = ecross(e2, e1);
  const float3 C = v0 - org;
  const float3 R = ecross(C, dir);
  const float den = edot(Ng, dir);
  const float absDen = fabsf(den);
  const unsigned sgn = __float_as_uint(den) & 0x80000000u;
  const float U = __uint_as_float(__float_as_uint(edot(e2, R)) ^ sgn);
  if (!(U >= 0.0f)) return false;
  const float V = __uint_as_float(__float_as_uint(edot(e1, R)) ^ sgn);
  if (!(V >= 0.0f)) return false;
  const float W = absDen - U - V;
  if (!(W >= 0.0f)) return false;
  const float T = __uint_as_float(__float_as_uint(edot(Ng, C)) ^ sgn);
  if (!((absDen * tnear < T) && (T <= absDen * tfar))) return false;
  if (!(den != 0.f)) return false;
  t = T / absDen;  // Embree: T*rcp(absDen), rcp+Newton (ISA dependent); policy: IEEE divide
  u = U / absDen;
  v = V / absDen;
  return true;
}

// DataDrivenSpheres.ispc:114-141 with (t0, tfar) the ORIGINAL interval
__device__ __forceinline__ bool sphere_test(const float4 ra, const float4 rb, float3 org, float3 dir, float t0, float tfar, float &t,
                                            float3 &Ng) {
  const float3 center = f3(ra.x, ra.y, ra.z);
  const float radius = ra.w, geps = rb.x;
  const float approxDist = dot3(center - org, dir);
  const float3 closeOrg = org + approxDist * dir;
  const float3 A = center - closeOrg;
  const float a = dot3(dir, dir);
  const float b = 2.f * dot3(dir, A);
  const float c = dot3(A, A) - radius * radius;
  const float radical = b * b - 4.f * a * c;
  if (radical < 0.f) return false;
  const float srad = sqrtf(radical);
  const float t_in = (b - srad) * (1.f / (2.f * a)) + approxDist;
  const float t_out = (b + srad) * (1.f / (2.f * a)) + approxDist;
  bool hit = false;
  if (t_in > t0 && t_in < tfar) { hit = true; t = t_in; }
  else if (t_out > (t0 + geps) && t_out < tfar) { hit = true; t = t_out; }
  if (hit) Ng = org + t * dir - center;
  return hit;
}

__device__ __forceinline__ float byte_f(unsigned w, int k) { return (float)((w >> (8 * k)) & 0xffu); }

// Nearest hit in (tnear, tfar]: smallest t wins; equal t -> lowest (geomID, primID)  [deterministic,
// independent of traversal order; Embree's own tie rule is BVH-order dependent, SURVEY A.7].
// ANYHIT: stop at the first accepted candidate (occlusion rays when nothing integrates along t).
// stack: shared-memory array [GXY_STACK_SMEM][blockDim.x] of uint2, this thread uses column threadIdx.x.
template <bool ANYHIT>
__device__ __forceinline__ bool traverse(const SceneParams &P, float3 org, float3 dir, float tnear, float tfar, Hit1 &best,
                                         uint2 *__restrict__ stack) {
  best.geom = -1;
  best.prim = -1;
  best.t = tfar;
  best.u = best.v = 0.f;
  best.Ng = f3(0.f, 0.f, 0.f);
  if (P.n_prims == 0) return false;
  bool found = false;
  const WideNode *__restrict__ nodes = P.nodes;
  const PrimRec *__restrict__ prims = P.prims;
  // box tests never decide a result: guarded reciprocal, FMA allowed
  const float gx = fabsf(dir.x) > 1e-30f ? dir.x : copysignf(1e-30f, dir.x);
  const float gy = fabsf(dir.y) > 1e-30f ? dir.y : copysignf(1e-30f, dir.y);
  const float gz = fabsf(dir.z) > 1e-30f ? dir.z : copysignf(1e-30f, dir.z);
  const float idx = 1.0f / gx, idy = 1.0f / gy, idz = 1.0f / gz;
  const bool o1 = dir.x < 0.f, o2 = dir.y < 0.f, o4 = dir.z < 0.f;
  const unsigned sel = o1 ? (o2 ? 0x0123u : 0x2301u) : (o2 ? 0x1032u : 0x3210u);
  uint2 local_stack[GXY_STACK_LOCAL];
  int sp = 0;
  const int tid = threadIdx.x, stride = blockDim.x;
  unsigned cur = 0;  // root node ref
  float cur_t = tnear;
  bool have = true;
  while (true) {
    if (!have) {
      if (sp == 0) break;
      --sp;
      uint2 e = sp < GXY_STACK_SMEM ? stack[sp * stride + tid] : local_stack[sp - GXY_STACK_SMEM];
      cur = e.x;
      cur_t = __uint_as_float(e.y);
      if (cur_t > best.t) continue;  // strict: equal-t candidates are still examined (tie rule)
    }
    have = false;
    if (cur & 0x80000000u) {
      // ---- leaf: up to 8 primitive records
      unsigned first = (cur & 0x7fffffffu) >> 3;
      int count = (int)(cur & 7u) + 1;
      for (int k = 0; k < count; k++) {
        const float4 *rec = reinterpret_cast<const float4 *>(prims + first + k);
        const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rc = __ldg(rec + 2);
        const unsigned gk = __float_as_uint(rc.y);
        const int geom = (int)(gk & 0xffffffu), prim = (int)__float_as_uint(rc.z);
        float t, u = 0.f, v = 0.f;
        float3 Ng;
        bool h;
        if ((gk >> 24) == 0) h = tri_test(ra, rb, rc, org, dir, tnear, tfar, t, u, v, Ng);
        else h = sphere_test(ra, rb, org, dir, tnear, tfar, t, Ng);
        if (h && (!found || t < best.t || (t == best.t && (geom < best.geom || (geom == best.geom && prim < best.prim))))) {
          found = true;
          best.t = t; best.u = u; best.v = v; best.geom = geom; best.prim = prim; best.Ng = Ng;
          if (ANYHIT) return true;
        }
      }
      continue;
    }
    // ---- inner node: test the 8 quantised child boxes
    const uint4 *np = reinterpret_cast<const uint4 *>(nodes + cur);
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4), n5 = __ldg(np + 5);
    const float sx = __uint_as_float((n0.w & 0xffu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23),
                sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
    const float ax = sx * idx, ay = sy * idy, az = sz * idz;
    const float bx = (__uint_as_float(n0.x) - org.x) * idx, by = (__uint_as_float(n0.y) - org.y) * idy,
                bz = (__uint_as_float(n0.z) - org.z) * idz;
    // permute slots so that position p holds slot p^oct (no dynamic register indexing):
    // bytes inside a word by PRMT with a per-ray selector, words and refs by conditional swaps
    unsigned r0 = n1.x, r1 = n1.y, r2 = n1.z, r3 = n1.w, r4 = n2.x, r5 = n2.y, r6 = n2.z, r7 = n2.w;
#define GXY_CSWAP(c, a_, b_) { unsigned ta = (c) ? b_ : a_; b_ = (c) ? a_ : b_; a_ = ta; }
    GXY_CSWAP(o1, r0, r1) GXY_CSWAP(o1, r2, r3) GXY_CSWAP(o1, r4, r5) GXY_CSWAP(o1, r6, r7)
    GXY_CSWAP(o2, r0, r2) GXY_CSWAP(o2, r1, r3) GXY_CSWAP(o2, r4, r6) GXY_CSWAP(o2, r5, r7)
    GXY_CSWAP(o4, r0, r4) GXY_CSWAP(o4, r1, r5) GXY_CSWAP(o4, r2, r6) GXY_CSWAP(o4, r3, r7)
    const unsigned refs[8] = {r0, r1, r2, r3, r4, r5, r6, r7};
    // qlox qloy | qloz qhix | qhiy qhiz   (8 bytes each)
    unsigned qlox[2] = {__byte_perm(n3.x, 0, sel), __byte_perm(n3.y, 0, sel)}, qloy[2] = {__byte_perm(n3.z, 0, sel), __byte_perm(n3.w, 0, sel)};
    unsigned qloz[2] = {__byte_perm(n4.x, 0, sel), __byte_perm(n4.y, 0, sel)}, qhix[2] = {__byte_perm(n4.z, 0, sel), __byte_perm(n4.w, 0, sel)};
    unsigned qhiy[2] = {__byte_perm(n5.x, 0, sel), __byte_perm(n5.y, 0, sel)}, qhiz[2] = {__byte_perm(n5.z, 0, sel), __byte_perm(n5.w, 0, sel)};
    GXY_CSWAP(o4, qlox[0], qlox[1]) GXY_CSWAP(o4, qloy[0], qloy[1]) GXY_CSWAP(o4, qloz[0], qloz[1])
    GXY_CSWAP(o4, qhix[0], qhix[1]) GXY_CSWAP(o4, qhiy[0], qhiy[1]) GXY_CSWAP(o4, qhiz[0], qhiz[1])
#undef GXY_CSWAP
    const float tb = best.t;
    // push far-to-near in octant order so the nearest octant is popped first
#pragma unroll
    for (int p = 7; p >= 0; p--) {
      const unsigned ref = refs[p];
      if (ref == 0) continue;
      const int w = p >> 2, k = p & 3;
      float x0 = byte_f(qlox[w], k), x1 = byte_f(qhix[w], k);
      float y0 = byte_f(qloy[w], k), y1 = byte_f(qhiy[w], k);
      float z0 = byte_f(qloz[w], k), z1 = byte_f(qhiz[w], k);
      float tx0 = __fmaf_rn(x0, ax, bx), tx1 = __fmaf_rn(x1, ax, bx);
      float ty0 = __fmaf_rn(y0, ay, by), ty1 = __fmaf_rn(y1, ay, by);
      float tz0 = __fmaf_rn(z0, az, bz), tz1 = __fmaf_rn(z1, az, bz);
      float tmin = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tnear));
      float tmax = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tb));
      // conservative slack: a few ulps of the magnitudes involved
      float slack = 4e-6f * fmaxf(fabsf(tmin), fabsf(tmax)) + 1e-30f;
      if (tmin - tmax <= slack) {
        if (have) {  // previously selected (farther) child goes to the stack
          uint2 e = make_uint2(cur, __float_as_uint(cur_t));
          if (sp < GXY_STACK_SMEM) stack[sp * stride + tid] = e;
          else if (sp < GXY_STACK_SMEM + GXY_STACK_LOCAL) local_stack[sp - GXY_STACK_SMEM] = e;
          if (sp < GXY_STACK_SMEM + GXY_STACK_LOCAL) sp++;
          else *P.error_flag = 1;
        }
        cur = ref;
        cur_t = fminf(tmin, tmax) - slack;
        have = true;
      }
    }
  }
  return found;
}

}  // namespace gxy
