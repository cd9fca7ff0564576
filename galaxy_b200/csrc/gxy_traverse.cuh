// gxy_traverse.cuh -- nearest-hit traversal of the compressed 8-wide BVH + the two primitive tests.
// Replaces rtcIntersectV (ospray/common/Model.ih:54-70) i.e. Embree's BVH8 traversal with
// MoellerTrumboreIntersectorK (embree/kernels/geometry/triangle_intersector_moeller.h:210-266)
// and the DataDrivenSpheres user-geometry callback (src/ospray/DataDrivenSpheres.ispc:90-155).
//
// Traversal scheme (after Ylitie, Karras, Laine 2017): the per-ray state is a NODE GROUP
// (child_base, hit bits of the still-unvisited internal children | imask) and a PRIMITIVE GROUP
// (prim_base, hit bits of up to 24 leaf primitives).  One step = pop the nearest child of the node
// group in octant order (highest set bit), test its 8 quantised child boxes branch-free into a new
// hit mask, then test the primitive group.  The stack holds one 8-byte entry per node whose
// children are only partly visited; the first GXY_STACK_SMEM entries live in shared memory.
// The persistent kernel runs the two phases warp-synchronously: lanes holding a primitive group
// wait until PRIM_T lanes of the warp have one (or nobody can do node work), so that the
// primitive tests run with many lanes instead of one or two (gxy_kernels.cu).
//
// Exactness: child boxes are conservative (quantised outwards at build time, rounding of the slab
// arithmetic bounded per node and added to the interval), every primitive is tested against the
// ORIGINAL (tnear, tfar] with the reference's arithmetic, and the winner is the smallest t with
// ties broken on the lowest (geomID, primID) -- independent of traversal order.
#pragma once
#include "gxy_common.cuh"
#include "gxy_curve.cuh"

namespace gxy {

struct Hit1 {
  float t, u, v;
  int geom, prim;
  float3 Ng;
};

#define GXY_STACK_SMEM 10   // entries per thread kept in shared memory
#define GXY_STACK_LOCAL 54  // overflow entries in local memory
#define GXY_TRACE_THREADS 128
#define GXY_NO_HIT 0xffffffffu

// Embree AVX2 op order (common/math/vec3.h:216,221): dot = madd(x,x, madd(y,y, z*z)),
// cross.x = msub(a.y,b.z, a.z*b.y)
__device__ __forceinline__ float edot(float3 a, float3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z))); }
__device__ __forceinline__ float3 ecross(float3 a, float3 b) {
  return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
            __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

// one triangle record against the ray; candidate interval is the ORIGINAL (tnear, tfar]
__device__ __forceinline__ bool tri_test(const float4 ra, const float4 rb, const float4 rc, float3 org, float3 dir, float tnear,
                                         float tfar, float &t, float &u, float &v) {
  const float3 v0 = f3(ra.x, ra.y, ra.z), e1 = f3(ra.w, rb.x, rb.y), e2 = f3(rb.z, rb.w, rc.x);
  const float3 Ng = ecross(e2, e1);
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
__device__ __forceinline__ bool sphere_test(const float4 ra, const float4 rb, float3 org, float3 dir, float t0, float tfar, float &t) {
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
  return hit;
}

// one round Bezier segment (PathLines) against the ray: the record holds the address of its 4 control points.
// Embree's sweep intersector (gxy_curve.cuh); hits in the OPEN interval (tnear, tfar), u = curve parameter.
__device__ __forceinline__ bool curve_rec_test(const float4 ra, float3 org, float3 dir, float tnear, float tfar, gxc::CurveHit &h) {
  const float4 *cp = reinterpret_cast<const float4 *>(((unsigned long long)__float_as_uint(ra.y) << 32) | (unsigned long long)__float_as_uint(ra.x));
  const float4 q0 = __ldg(cp), q3 = __ldg(cp + 3);
  // conservative cull (ra.z = the segment's bound radius, written with the record): most rays that reach a leaf box pass the
  // thin tube by; they skip the sub-division
  if (gxc::curve_precull(gxc::v3(q0.x, q0.y, q0.z), gxc::v3(q3.x, q3.y, q3.z), ra.z, gxc::v3(org.x, org.y, org.z), gxc::v3(dir.x, dir.y, dir.z)))
    return false;
  float c[16];
  c[0] = q0.x; c[1] = q0.y; c[2] = q0.z; c[3] = q0.w;
  c[12] = q3.x; c[13] = q3.y; c[14] = q3.z; c[15] = q3.w;
#pragma unroll
  for (int k = 1; k < 3; k++) {
    const float4 q = __ldg(cp + k);
    c[4 * k] = q.x; c[4 * k + 1] = q.y; c[4 * k + 2] = q.z; c[4 * k + 3] = q.w;
  }
  return gxc::curve_test(c, org.x, org.y, org.z, dir.x, dir.y, dir.z, tnear, tfar, h);
}

// ---- per-ray constants and traversal state ------------------------------------------------------
struct RayCtx {
  float3 org, dir;
  float idx, idy, idz;  // guarded reciprocals (box tests only; they never decide a result)
  float tnear, tfar;    // the ORIGINAL interval
  unsigned octinv4;     // (7 ^ negative-direction mask) replicated in 4 bytes
};

struct TravState {
  uint2 ng, tg;       // node group (child_base, hits<<24 | imask), primitive group (prim_base, 24 hit bits)
  int sp;
  float best_t, best_u, best_v;
  unsigned best_key;  // geomID << 28 | primID; GXY_NO_HIT = none (the builder guarantees primID < 2^28, geomID < 16)
  unsigned best_rec;  // index of the winning primitive record
#ifdef GXY_TRAV_COUNTERS
  unsigned n_nodes, n_prims;
#endif
};

// rcp: 1.0f/dir per component if the caller has it already (the box clip of TraceRays computes it), else NULL
__device__ __forceinline__ void ray_ctx_init(RayCtx &rc, float3 org, float3 dir, float tnear, float tfar, const float3 *rcp = nullptr) {
  rc.org = org; rc.dir = dir; rc.tnear = tnear; rc.tfar = tfar;
  const bool sx = fabsf(dir.x) > 1e-30f, sy = fabsf(dir.y) > 1e-30f, sz = fabsf(dir.z) > 1e-30f;
  if (rcp && sx && sy && sz) {
    rc.idx = rcp->x; rc.idy = rcp->y; rc.idz = rcp->z;
  } else {
    const float gx = sx ? dir.x : copysignf(1e-30f, dir.x);
    const float gy = sy ? dir.y : copysignf(1e-30f, dir.y);
    const float gz = sz ? dir.z : copysignf(1e-30f, dir.z);
    rc.idx = 1.0f / gx; rc.idy = 1.0f / gy; rc.idz = 1.0f / gz;
  }
  // octant from the sign of the reciprocal actually used, so that near/far plane selection and
  // child order always agree with the slab arithmetic (also for -0.0 components)
  const unsigned rs = (rc.idx < 0.f ? 1u : 0u) | (rc.idy < 0.f ? 2u : 0u) | (rc.idz < 0.f ? 4u : 0u);
  rc.octinv4 = (7u ^ rs) * 0x01010101u;
}

__device__ __forceinline__ void trav_init(TravState &s, const RayCtx &rc) {
  s.ng = make_uint2(0u, 0x80000000u);  // root = "child 0 of a virtual group": bit 31, imask 0
  s.tg = make_uint2(0u, 0u);
  s.sp = 0;
  s.best_t = rc.tfar; s.best_u = 0.f; s.best_v = 0.f;
  s.best_key = GXY_NO_HIT; s.best_rec = 0u;
#ifdef GXY_TRAV_COUNTERS
  s.n_nodes = 0; s.n_prims = 0;
#endif
}

// Quantised plane byte -> float without the conversion unit.  I2F.U8 runs on the quarter-rate XU pipe (measured on
// B200: 16 conversions/clk/SM, tools/ubench/f32x2.cu) and a node test needs 48 of them; instead one PRMT drops the
// byte into mantissa bits 8..15 of 128.0f: f = 128 + q/256 exactly, and the slab equation t = q*a + b becomes
// t = f*(256a) + (b - 32768a), one FFMA per plane (the constant folding costs 3 FFMA per node; its rounding,
// <= ulp(32768|a|)/2 = |a|/512, goes into the conservative widening E).
//   GXY_BOX_CONV 0: I2F + FFMA   1: PRMT + FFMA   2: PRMT + FFMA2 (near/far plane of an axis as one packed f32x2 op)
#ifndef GXY_BOX_CONV
#define GXY_BOX_CONV 0
#endif
__device__ __forceinline__ float byte_f(unsigned w, int k) {
#if GXY_BOX_CONV == 0
  return (float)((w >> (8 * k)) & 0xffu);
#else
  return __uint_as_float(__byte_perm(w, 0x43000000u, 0x7404u + 0x10u * (unsigned)k));
#endif
}

// the 4 children of one half of a node: returns their contribution to the hit mask
__device__ __forceinline__ unsigned test_half(unsigned meta4, unsigned octinv4, unsigned nx4, unsigned ny4, unsigned nz4, unsigned fx4,
                                              unsigned fy4, unsigned fz4, float ax, float ay, float az, float bnx, float bny, float bnz,
                                              float bfx, float bfy, float bfz, float tnear, float tbest) {
  const unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
  const unsigned inner_mask4 = (is_inner4 >> 4) * 0xffu;
  const unsigned bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
  const unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
  unsigned hitmask = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
#if GXY_BOX_CONV == 2
    const float2 tx = __ffma2_rn(make_float2(byte_f(nx4, k), byte_f(fx4, k)), make_float2(ax, ax), make_float2(bnx, bfx));
    const float2 ty = __ffma2_rn(make_float2(byte_f(ny4, k), byte_f(fy4, k)), make_float2(ay, ay), make_float2(bny, bfy));
    const float2 tz = __ffma2_rn(make_float2(byte_f(nz4, k), byte_f(fz4, k)), make_float2(az, az), make_float2(bnz, bfz));
    const float tnx = tx.x, tfx = tx.y, tny = ty.x, tfy = ty.y, tnz = tz.x, tfz = tz.y;
#else
    const float tnx = __fmaf_rn(byte_f(nx4, k), ax, bnx), tfx = __fmaf_rn(byte_f(fx4, k), ax, bfx);
    const float tny = __fmaf_rn(byte_f(ny4, k), ay, bny), tfy = __fmaf_rn(byte_f(fy4, k), ay, bfy);
    const float tnz = __fmaf_rn(byte_f(nz4, k), az, bnz), tfz = __fmaf_rn(byte_f(fz4, k), az, bfz);
#endif
    const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tnear));
    const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, tbest));
    if (tmin <= tmax) hitmask |= ((child_bits4 >> (8 * k)) & 0xffu) << ((bit_index4 >> (8 * k)) & 0xffu);
  }
  return hitmask;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void stack_push(const SceneParams &P, TravState &s, uint2 e, uint2 *__restrict__ stack, uint2 *__restrict__ lstack) {
  if (s.sp < GXY_STACK_SMEM) stack[s.sp * blockDim.x + threadIdx.x] = e;
  else if (s.sp < GXY_STACK_SMEM + GXY_STACK_LOCAL) lstack[s.sp - GXY_STACK_SMEM] = e;
  if (s.sp < GXY_STACK_SMEM + GXY_STACK_LOCAL) s.sp++;
  else *P.error_flag = 1;
}

// Test the 8 quantised child boxes of one node against the ray: `bit` (24..31) selects a hit child of the node
// group (child_base, hits_imask); tbest = current upper end of the interval.  Returns the node group and the
// primitive group of that node.  No traversal state is touched (also used by the generation kernel's cull).
__device__ __forceinline__ void node_eval(const WideNode *__restrict__ nodes, const RayCtx &rc, unsigned child_base, unsigned hits_imask, int bit,
                                          float tb, uint2 &ng_out, uint2 &tg_out) {
  const unsigned slot = (unsigned)(bit - 24) ^ (rc.octinv4 & 7u);
  const unsigned rel = __popc(hits_imask & ~(0xffffffffu << slot));
  const uint4 *np = reinterpret_cast<const uint4 *>(nodes + (child_base + rel));
  const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
  // slab coefficients in the node's quantisation frame: t = q * a + b
  const float ax0 = __uint_as_float((n0.w & 0xffu) << 23) * rc.idx, ay0 = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * rc.idy,
              az0 = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * rc.idz;
  const float bx0 = (__uint_as_float(n0.x) - rc.org.x) * rc.idx, by0 = (__uint_as_float(n0.y) - rc.org.y) * rc.idy,
              bz0 = (__uint_as_float(n0.z) - rc.org.z) * rc.idz;
#if GXY_BOX_CONV == 0
  const float ax = ax0, ay = ay0, az = az0, bx = bx0, by = by0, bz = bz0;
  // rounding of (q*a + b) against the exact plane distance is below 2^-24 * (4|b| + 510|a|)
  // (reciprocal, difference, product, fma): widen the interval by 4e-7 * (|b| + 255|a|) per axis
  const float Ex = 4e-7f * __fmaf_rn(255.f, fabsf(ax), fabsf(bx)), Ey = 4e-7f * __fmaf_rn(255.f, fabsf(ay), fabsf(by)),
              Ez = 4e-7f * __fmaf_rn(255.f, fabsf(az), fabsf(bz));
  const float bnx = bx - Ex, bny = by - Ey, bnz = bz - Ez, bfx = bx + Ex, bfy = by + Ey, bfz = bz + Ez;
#else
  // byte_f() yields f = 128 + q/256: t = f*(256 a) + (b - 32768 a).  256a is exact; the folded constant rounds once more,
  // by at most ulp(|b| + 32768|a|)/2 <= 6e-8|b| + |a|/512.  Total: (2.4e-7 + 6e-8)|b| + (3.1e-5 + 1/512)|a| < 5e-7|b| + 0.0021|a|,
  // applied with directed rounding because it is of the order of one ulp of the folded constant.
  const float ax = 256.f * ax0, ay = 256.f * ay0, az = 256.f * az0;
  const float bx = __fmaf_rn(-32768.f, ax0, bx0), by = __fmaf_rn(-32768.f, ay0, by0), bz = __fmaf_rn(-32768.f, az0, bz0);
  const float Ex = __fmaf_rn(0.0021f, fabsf(ax0), 5e-7f * fabsf(bx0)), Ey = __fmaf_rn(0.0021f, fabsf(ay0), 5e-7f * fabsf(by0)),
              Ez = __fmaf_rn(0.0021f, fabsf(az0), 5e-7f * fabsf(bz0));
  const float bnx = __fadd_rd(bx, -Ex), bny = __fadd_rd(by, -Ey), bnz = __fadd_rd(bz, -Ez);
  const float bfx = __fadd_ru(bx, Ex), bfy = __fadd_ru(by, Ey), bfz = __fadd_ru(bz, Ez);
#endif
  const bool negx = rc.idx < 0.f, negy = rc.idy < 0.f, negz = rc.idz < 0.f;
  // n2 = qlox[8] qloy[8]   n3 = qloz[8] qhix[8]   n4 = qhiy[8] qhiz[8]
  const unsigned nx0 = negx ? n3.z : n2.x, nx1 = negx ? n3.w : n2.y, fx0 = negx ? n2.x : n3.z, fx1 = negx ? n2.y : n3.w;
  const unsigned ny0 = negy ? n4.x : n2.z, ny1 = negy ? n4.y : n2.w, fy0 = negy ? n2.z : n4.x, fy1 = negy ? n2.w : n4.y;
  const unsigned nz0 = negz ? n4.z : n3.x, nz1 = negz ? n4.w : n3.y, fz0 = negz ? n3.x : n4.z, fz1 = negz ? n3.y : n4.w;
  unsigned hitmask = test_half(n1.z, rc.octinv4, nx0, ny0, nz0, fx0, fy0, fz0, ax, ay, az, bnx, bny, bnz, bfx, bfy, bfz, rc.tnear, tb);
  hitmask |= test_half(n1.w, rc.octinv4, nx1, ny1, nz1, fx1, fy1, fz1, ax, ay, az, bnx, bny, bnz, bfx, bfy, bfz, rc.tnear, tb);
  ng_out = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
  tg_out = make_uint2(n1.y, hitmask & 0x00ffffffu);
}

// Conservative cull (generation kernel): false only if the ray (interval of rc) cannot reach any primitive -- no child
// box of the root is hit, or none of the boxes LEVELS levels below those.  Box tests are conservative (see above), so a
// culled ray has no candidate and the full traversal would return "no hit".
// nodes: the tree itself, or (LEVELS = 1) a 9-node excerpt of it (root with child_base = 1, followed by its internal children).
template <int D>
__device__ __forceinline__ bool cull_group(const WideNode *__restrict__ nodes, const RayCtx &rc, uint2 g) {
  unsigned pending = g.y;
  while (pending > 0x00ffffffu) {
    const int bit = 31 - __clz((int)pending);
    pending &= ~(1u << bit);
    uint2 ng, tg;
    node_eval(nodes, rc, g.x, g.y, bit, rc.tfar, ng, tg);
    if (tg.y != 0u) return true;
    if (ng.y > 0x00ffffffu) {
      if constexpr (D <= 1) return true;
      else if (cull_group<D - 1>(nodes, rc, ng)) return true;
    }
  }
  return false;
}
template <int LEVELS = 1>
__device__ __forceinline__ bool may_hit_anything(const WideNode *__restrict__ nodes, const RayCtx &rc) {
  uint2 ng, tg;
  node_eval(nodes, rc, 0u, 0x80000000u, 31, rc.tfar, ng, tg);
  if (tg.y != 0u) return true;
  return cull_group<LEVELS>(nodes, rc, ng);
}

// Node phase: pop the nearest unvisited internal child of the node group (requires s.ng.y > 0x00ffffff)
// and test its 8 children -> new node group + primitive group.
// PREFETCH: 0 none; 1 = L2 prefetch of the primitive records of the new group and of the node that
// will be visited next (overlaps their DRAM latency with the primitive phase); 2 = same into L1
template <int PREFETCH>
__device__ __forceinline__ void node_step(const SceneParams &P, const RayCtx &rc, TravState &s, uint2 *__restrict__ stack,
                                          uint2 *__restrict__ lstack) {
  const unsigned hits_imask = s.ng.y;
  const int bit = 31 - __clz((int)hits_imask);
  s.ng.y &= ~(1u << bit);
  if (s.ng.y > 0x00ffffffu) stack_push(P, s, s.ng, stack, lstack);  // siblings left: the group goes to the stack
#ifdef GXY_TRAV_COUNTERS
  s.n_nodes++;
#endif
  node_eval(P.nodes, rc, s.ng.x, hits_imask, bit, s.best_t, s.ng, s.tg);
  if (PREFETCH) {
    if (s.tg.y) {
      const char *first = reinterpret_cast<const char *>(P.prims + (s.tg.x + (unsigned)(__ffs((int)s.tg.y) - 1)));
      const char *last = reinterpret_cast<const char *>(P.prims + (s.tg.x + (unsigned)(31 - __clz((int)s.tg.y))));
      if (PREFETCH == 1) { prefetch_l2(first); prefetch_l2(first + 32); prefetch_l2(last + 16); prefetch_l2(last + 32); }
      else { prefetch_l1(first); prefetch_l1(first + 32); prefetch_l1(last + 16); prefetch_l1(last + 32); }
    }
    if (s.ng.y > 0x00ffffffu) {
      const int nbit = 31 - __clz((int)s.ng.y);
      const unsigned nslot = (unsigned)(nbit - 24) ^ (rc.octinv4 & 7u);
      const char *nn = reinterpret_cast<const char *>(P.nodes + (s.ng.x + __popc(s.ng.y & ~(0xffffffffu << nslot))));
      if (PREFETCH == 1) { prefetch_l2(nn); prefetch_l2(nn + 32); prefetch_l2(nn + 64); }
      else { prefetch_l1(nn); prefetch_l1(nn + 32); prefetch_l1(nn + 64); }
    }
  }
}

// Primitive phase: test every primitive of the group.  Returns true if the traversal is over
// (anyhit and a candidate was accepted).  CURVES: the scene holds PathLines (kind 2 records); only the list-path
// kernels of such scenes are instantiated with it, every other kernel is compiled exactly as before.
template <bool CURVES = false>
__device__ __forceinline__ bool prim_step(const SceneParams &P, const RayCtx &rc, TravState &s, const bool anyhit) {
  while (s.tg.y != 0u) {
    const int b = __ffs((int)s.tg.y) - 1;
    s.tg.y &= s.tg.y - 1u;
    const unsigned ri = s.tg.x + (unsigned)b;
    const float4 *rec = reinterpret_cast<const float4 *>(P.prims + ri);
    const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rcq = __ldg(rec + 2);
#ifdef GXY_TRAV_COUNTERS
    s.n_prims++;
#endif
    const unsigned gk = __float_as_uint(rcq.y);
    const unsigned key = ((gk & 0xffffffu) << 28) | __float_as_uint(rcq.z);
    float t, u = 0.f, v = 0.f;
    bool h;
    if ((gk >> 24) == 0) h = tri_test(ra, rb, rcq, rc.org, rc.dir, rc.tnear, rc.tfar, t, u, v);
    else if (!CURVES || (gk >> 24) == 1) h = sphere_test(ra, rb, rc.org, rc.dir, rc.tnear, rc.tfar, t);
    else {
      gxc::CurveHit ch;
      h = curve_rec_test(ra, rc.org, rc.dir, rc.tnear, rc.tfar, ch);
      if (h) { t = ch.t; u = ch.u; }
    }
    if (h && (t < s.best_t || (t == s.best_t && key < s.best_key))) {
      s.best_t = t; s.best_u = u; s.best_v = v; s.best_key = key; s.best_rec = ri;
      if (anyhit) return true;
    }
  }
  return false;
}

// Cooperative primitive passes of one warp (all 32 lanes must call this, converged): lane 4g+k tests
// the k-th pending primitive of the g-th lane that holds a primitive group, the owner then takes the
// (t, geomID, primID)-smallest candidate of its helpers.  On return no lane has pending primitives.
// owner_slot: 8 bytes of shared memory private to the warp.
template <bool CURVES = false>
__device__ __forceinline__ void coop_prim_passes(const SceneParams &P, const RayCtx &rc, TravState &st, bool &trav, const bool anyhit,
                                                 unsigned char *owner_slot, const unsigned lane, const unsigned lt_mask) {
  const unsigned FULL = 0xffffffffu;
  unsigned owners = __ballot_sync(FULL, trav && st.tg.y != 0u);
  while (owners != 0u) {
    const bool own = trav && st.tg.y != 0u;
    const unsigned r = (unsigned)__popc(owners & lt_mask);  // rank among the owning lanes
    if (own && r < 8u) owner_slot[r] = (unsigned char)lane;
    __syncwarp();
    const unsigned g = lane >> 2, k = lane & 3u;
    const bool gvalid = g < (unsigned)__popc(owners);
    const unsigned o = gvalid ? (unsigned)owner_slot[g] : lane;
    const unsigned obits = __shfl_sync(FULL, st.tg.y, o), obase = __shfl_sync(FULL, st.tg.x, o);
    unsigned b = obits;  // drop the k lowest set bits: the k-th pending primitive of the owner
    if (k >= 1u) b &= b - 1u;
    if (k >= 2u) b &= b - 1u;
    if (k >= 3u) b &= b - 1u;
    const bool tvalid = gvalid && b != 0u;
    const float3 oorg = f3(__shfl_sync(FULL, rc.org.x, o), __shfl_sync(FULL, rc.org.y, o), __shfl_sync(FULL, rc.org.z, o));
    const float3 odir = f3(__shfl_sync(FULL, rc.dir.x, o), __shfl_sync(FULL, rc.dir.y, o), __shfl_sync(FULL, rc.dir.z, o));
    const float otn = __shfl_sync(FULL, rc.tnear, o), otf = __shfl_sync(FULL, rc.tfar, o);
    float ct = __int_as_float(0x7f800000), cu = 0.f, cv = 0.f;
    unsigned ckey = GXY_NO_HIT, crec = 0u;
    if (tvalid) {
      crec = obase + (unsigned)(__ffs((int)b) - 1);
      const float4 *rec = reinterpret_cast<const float4 *>(P.prims + crec);
      const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rcq = __ldg(rec + 2);
#ifdef GXY_TRAV_COUNTERS
      atomicAdd(P.trav_counters + 1, 1ull);
#endif
      const unsigned gk = __float_as_uint(rcq.y);
      float t, u = 0.f, v = 0.f;
      bool h;
      if ((gk >> 24) == 0) h = tri_test(ra, rb, rcq, oorg, odir, otn, otf, t, u, v);
      else if (!CURVES || (gk >> 24) == 1) h = sphere_test(ra, rb, oorg, odir, otn, otf, t);
      else {  // a round Bezier segment of a PathLines operator (only the CURVES instantiations of the frame kernels get here)
        gxc::CurveHit ch;
        h = curve_rec_test(ra, oorg, odir, otn, otf, ch);
        if (h) { t = ch.t; u = ch.u; }
      }
      if (h) { ct = t; cu = u; cv = v; ckey = ((gk & 0xffffffu) << 28) | __float_as_uint(rcq.z); }
    }
    // the owner (rank r < 8) picks the best of its helpers, lanes 4r .. 4r+3
    const unsigned h0 = (4u * r) & 31u;
    int hb = -1;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const float ht = __shfl_sync(FULL, ct, h0 + kk);
      const unsigned hk = __shfl_sync(FULL, ckey, h0 + kk);
      if (own && r < 8u && hk != GXY_NO_HIT && (ht < st.best_t || (ht == st.best_t && hk < st.best_key))) {
        st.best_t = ht; st.best_key = hk; hb = kk;
      }
    }
    const unsigned hsrc = h0 + (unsigned)(hb < 0 ? 0 : hb);
    const float hu = __shfl_sync(FULL, cu, hsrc), hv = __shfl_sync(FULL, cv, hsrc);
    const unsigned hrec = __shfl_sync(FULL, crec, hsrc);
    if (own && r < 8u) {
      if (hb >= 0) { st.best_u = hu; st.best_v = hv; st.best_rec = hrec; }
      unsigned nb = st.tg.y;  // the 4 lowest pending primitives have been tested
      nb &= nb - 1u; nb &= nb - 1u; nb &= nb - 1u; nb &= nb - 1u;
      st.tg.y = nb;
      if (hb >= 0 && anyhit) { trav = false; st.tg.y = 0u; }
    }
    __syncwarp();
    owners = __ballot_sync(FULL, trav && st.tg.y != 0u);
  }
}

// After a phase: make sure the lane has a node group with unvisited children (or pending primitives);
// returns false when the traversal is complete.
__device__ __forceinline__ bool trav_advance(TravState &s, const uint2 *__restrict__ stack, const uint2 *__restrict__ lstack) {
  if (s.tg.y == 0u && s.ng.y <= 0x00ffffffu) {
    if (s.sp == 0) return false;
    --s.sp;
    s.ng = s.sp < GXY_STACK_SMEM ? stack[s.sp * blockDim.x + threadIdx.x] : lstack[s.sp - GXY_STACK_SMEM];
  }
  return true;
}

// One per-lane traversal step (node, then its primitives).  Returns false when the traversal is complete.
template <bool CURVES = false>
__device__ __forceinline__ bool trav_step(const SceneParams &P, const RayCtx &rc, TravState &s, const bool anyhit,
                                          uint2 *__restrict__ stack, uint2 *__restrict__ lstack) {
  if (s.ng.y > 0x00ffffffu) node_step<0>(P, rc, s, stack, lstack);
  if (prim_step<CURVES>(P, rc, s, anyhit)) return false;
  return trav_advance(s, stack, lstack);
}

// geomID / primID / Ng of the winner, recomputed from its record (saves registers in the loop)
template <bool CURVES = false>
__device__ __forceinline__ void trav_fetch_hit(const SceneParams &P, const RayCtx &rc, const TravState &s, Hit1 &best) {
  best.t = s.best_t; best.u = s.best_u; best.v = s.best_v;
  best.geom = (int)(s.best_key >> 28);
  best.prim = (int)(s.best_key & 0x0fffffffu);
  const float4 *rec = reinterpret_cast<const float4 *>(P.prims + s.best_rec);
  const float4 ra = __ldg(rec), rb = __ldg(rec + 1), rcq = __ldg(rec + 2);
  if ((__float_as_uint(rcq.y) >> 24) == 0) {
    const float3 e1 = f3(ra.w, rb.x, rb.y), e2 = f3(rb.z, rb.w, rcq.x);
    best.Ng = ecross(e2, e1);  // embree triangle.h:133-136
  } else if (!CURVES || (__float_as_uint(rcq.y) >> 24) == 1) {
    best.Ng = rc.org + s.best_t * rc.dir - f3(ra.x, ra.y, ra.z);  // DataDrivenSpheres.ispc:143-150
  } else {
    // the same test on the same interval finds the same nearest hit of this segment again, now for its normal
    // (curve_intersector_sweep.h:110-113); keeps Ng out of the traversal state
    gxc::CurveHit ch;
    ch.Ng.x = ch.Ng.y = ch.Ng.z = 0.f;
    curve_rec_test(ra, rc.org, rc.dir, rc.tnear, rc.tfar, ch);
    best.Ng = f3(ch.Ng.x, ch.Ng.y, ch.Ng.z);
  }
}

// Nearest hit in (tnear, tfar], run to completion.  ANYHIT: stop at the first accepted candidate
// (occlusion rays when nothing integrates along t).  stack: shared-memory array
// [GXY_STACK_SMEM][blockDim.x] of uint2, this thread uses column threadIdx.x.
template <bool ANYHIT, bool CURVES = false>
__device__ __forceinline__ bool traverse(const SceneParams &P, float3 org, float3 dir, float tnear, float tfar, Hit1 &best,
                                         uint2 *__restrict__ stack) {
  best.geom = -1; best.prim = -1; best.t = tfar; best.u = best.v = 0.f;
  best.Ng = f3(0.f, 0.f, 0.f);
  if (P.n_prims == 0) return false;
  RayCtx rc;
  ray_ctx_init(rc, org, dir, tnear, tfar);
  TravState s;
  trav_init(s, rc);
  uint2 lstack[GXY_STACK_LOCAL];
  while (trav_step<CURVES>(P, rc, s, ANYHIT, stack, lstack)) {}
  if (s.best_key == GXY_NO_HIT) return false;
  trav_fetch_hit<CURVES>(P, rc, s, best);
  return true;
}

}  // namespace gxy
