// gxy_shade.cuh -- per-ray device code shared by the list-based kernels (gxy_kernels.cu) and the fused
// frame kernels (gxy_fused.cu): ray generation, box clipping, postIntersect shading, lighting and
// secondary-ray construction, classification.  Every function restates the reference lines it cites.
#pragma once
#include <mutex>
#include "gxy_traverse.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>

namespace gxy {

// postIntersect of a geometry hit (Model.ih:97-187 + DataDrivenTriangleMesh.ispc:34-121 /
// DataDrivenSpheres.ispc:46-63): colour from the transfer function, shading normal normalised and
// faced towards the ray
template <bool CURVES = false>
__device__ __forceinline__ void shade_geometry_hit(const SceneParams &P, const Hit1 &h1, float3 dir, float3 &col, float &ca, float3 &Ns) {
  const DevGeom g = P.geoms[h1.geom];
  float3 Ng = h1.Ng;
  Ns = h1.Ng;
  col = f3(1.f, 1.f, 1.f);
  ca = 1.f;
  if (g.kind == 0) {  // DataDrivenTriangleMesh.ispc:34-121
    const int i0 = __ldg(g.idx + 3 * (size_t)h1.prim), i1 = __ldg(g.idx + 3 * (size_t)h1.prim + 1), i2 = __ldg(g.idx + 3 * (size_t)h1.prim + 2);
    const float3 bary = f3(1.0f - h1.u - h1.v, h1.u, h1.v);
    if (g.normals) {
      const float3 a = f3(__ldg(g.normals + 3 * (size_t)i0), __ldg(g.normals + 3 * (size_t)i0 + 1), __ldg(g.normals + 3 * (size_t)i0 + 2));
      const float3 b = f3(__ldg(g.normals + 3 * (size_t)i1), __ldg(g.normals + 3 * (size_t)i1 + 1), __ldg(g.normals + 3 * (size_t)i1 + 2));
      const float3 c = f3(__ldg(g.normals + 3 * (size_t)i2), __ldg(g.normals + 3 * (size_t)i2 + 1), __ldg(g.normals + 3 * (size_t)i2 + 2));
      Ns = bary.x * a + bary.y * b + bary.z * c;  // interpolate(), vec.ih:723-726
    }
    if (g.data) {
      const float d = bary.x * __ldg(g.data + i0) + bary.y * __ldg(g.data + i1) + bary.z * __ldg(g.data + i2);
      col = tf_color(P.tfs + g.tf, d);
      ca = 1.0f;
    }
  } else if (!CURVES || g.kind == 1) {  // DataDrivenSpheres.ispc:46-63
    col = tf_color(P.tfs + g.tf, g.data ? __ldg(g.data + h1.prim) : 0.f);
    ca = 1.0f;
  } else {  // DataDrivenPathLines_postIntersect (DataDrivenPathLines.ispc:213-277): Ng = Ns = ray.Ng; the radius between
    // the segment's FIRST TWO control points at the curve parameter u is mapped back to a data value
    const float w0 = __ldg(g.centers + 16 * (size_t)h1.prim + 3), w1 = __ldg(g.centers + 16 * (size_t)h1.prim + 7);
    const float radius = ((1.f - h1.u) * w0) + (h1.u * w1);
    float dataval;
    if (g.radius0 == g.radius1) dataval = g.value0;
    else if (g.radius0 < g.radius1 && radius < g.radius0) dataval = g.value0;
    else if (g.radius0 < g.radius1 && radius > g.radius1) dataval = g.value1;
    else if (g.radius0 > g.radius1 && radius < g.radius1) dataval = g.value1;
    else if (g.radius0 > g.radius1 && radius > g.radius0) dataval = g.value0;
    else dataval = g.value0 + ((radius - g.radius0) / (g.radius1 - g.radius0)) * (g.value1 - g.value0);
    col = tf_color(P.tfs + g.tf, dataval);
    ca = 1.0f;
  }
  Ng = normalize_isp(Ng);
  Ns = normalize_isp(Ns);
  if (dot3(dir, Ng) >= 0.f) Ng = neg3(Ng);
  if (dot3(Ng, Ns) < 0.f) Ns = neg3(Ns);
}


// ------------------------------------------------------------------------------------------------
// A ray in traversal, as the persistent kernels carry it between setup and finish
struct PendingRay {
  int ray;        // index in the list, -1 = none
  float tExit;    // tExitVolume of MyIntersectBox
  bool anyhit;
  bool opaque;    // the ray's own colour already makes it OPAQUE (TraceRays.ispc:573)
};

// clip to the local box (TraceRays.ispc:377-418) and start a traversal; returns false if the interval is empty
__device__ __forceinline__ bool setup_ray_box(float3 bmin, float3 bmax, bool has_prims, int i, bool shadeFlag, float3 org, float3 dir,
                                              float ray_t0, float ray_t, int anyhit_secondary, RayCtx &rc, TravState &st, PendingRay &pr) {
  if (dir.x == 0.f) dir.x = 1e-6f;  // :377-379
  if (dir.y == 0.f) dir.y = 1e-6f;
  if (dir.z == 0.f) dir.z = 1e-6f;
  float tEntry, tExitVolume;
  float3 rcp;
  {
    const float rx = 1.0f / dir.x, ry = 1.0f / dir.y, rz = 1.0f / dir.z;
    rcp = f3(rx, ry, rz);
    const float mnx = (bmin.x - org.x) * rx, mny = (bmin.y - org.y) * ry, mnz = (bmin.z - org.z) * rz;
    const float mxx = (bmax.x - org.x) * rx, mxy = (bmax.y - org.y) * ry, mxz = (bmax.z - org.z) * rz;
    tEntry = fmaxf(fminf(mnx, mxx), fmaxf(fminf(mny, mxy), fminf(mnz, mxz)));
    tExitVolume = fminf(fmaxf(mnx, mxx), fminf(fmaxf(mny, mxy), fmaxf(mnz, mxz)));
  }
  if (tEntry < ray_t0) tEntry = ray_t0;  // :412-413
  else if (tEntry > ray_t0) ray_t0 = tEntry;
  ray_t = fminf(ray_t, tExitVolume);  // :418
  ray_ctx_init(rc, org, dir, ray_t0, ray_t, &rcp);
  trav_init(st, rc);
  pr.ray = i;
  pr.tExit = tExitVolume;
  pr.anyhit = !shadeFlag && anyhit_secondary;
  pr.opaque = false;
  // an empty interval cannot accept any candidate (both primitive tests need tnear < t <= tfar); a partition
  // whose clipped geometry is empty has no tree to walk
  return ray_t0 <= ray_t && has_prims;
}
__device__ __forceinline__ bool setup_ray_values(const SceneParams &P, int i, bool shadeFlag, float3 org, float3 dir, float ray_t0, float ray_t,
                                                 int anyhit_secondary, RayCtx &rc, TravState &st, PendingRay &pr) {
  return setup_ray_box(P.lmin, P.lmax, P.n_prims > 0, i, shadeFlag, org, dir, ray_t0, ray_t, anyhit_secondary, rc, st, pr);
}


__device__ __forceinline__ bool setup_geom_ray(const SceneParams &P, const Rays &R, int i, int anyhit_secondary, RayCtx &rc, TravState &st,
                                               PendingRay &pr) {
  return setup_ray_values(P, i, R.type[i] == RAY_PRIMARY, f3(R.ox[i], R.oy[i], R.oz[i]), f3(R.dx[i], R.dy[i], R.dz[i]), R.t[i], R.tMax[i],
                          anyhit_secondary, rc, st, pr);
}

// ------------------------------------------------------------------------------------------------
// AO direction tables (src/renderer/UV.ih:21-58 + TraceRays.ispc:692-701), filled by the host:
// x = cos(2*pi*r0)*sqrt(1-r1), y = sin(2*pi*r0)*sqrt(1-r1), z = sqrt(r1)  for the 256 (r0,r1) pairs
static __constant__ float c_ao_x[256], c_ao_y[256], c_ao_z[256];  // one copy per translation unit

static void halton_tables(float U[256], float V[256]) {
  // UV.ih holds base-2 / base-3 radical inverses accumulated in fp32 and printed with "%g"
  for (int pass = 0; pass < 2; pass++) {
    const int b = pass ? 3 : 2;
    for (int i = 0; i < 256; i++) {
      float inv = 1.f / (float)b, f = inv, r = 0.f;
      for (int k = i; k > 0; k /= b) { r = r + f * (float)(k % b); f = f * inv; }
      char buf[64];
      snprintf(buf, sizeof buf, "%g", (double)r);
      (pass ? V : U)[i] = strtof(buf, nullptr);
    }
  }
}

static int ensure_ao_tables() {
  // __constant__ memory is per device: one upload for each device this process drives (a process may hold several contexts)
  static std::mutex mu;
  static bool done_on[64] = {};
  int dev = 0;
  GXY_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  bool &done = done_on[dev & 63];
  if (done) return 0;
  float U[256], V[256], x[256], y[256], z[256];
  halton_tables(U, V);
  for (int r = 0; r < 256; r++) {
    const float r0 = U[r], r1 = V[r];
    const float w = sqrtf(1.f - r1);
    x[r] = cosf((2.f * (float)M_PI) * r0) * w;
    y[r] = sinf((2.f * (float)M_PI) * r0) * w;
    z[r] = sqrtf(r1);
  }
  GXY_CUDA(cudaMemcpyToSymbol(c_ao_x, x, sizeof x));
  GXY_CUDA(cudaMemcpyToSymbol(c_ao_y, y, sizeof y));
  GXY_CUDA(cudaMemcpyToSymbol(c_ao_z, z, sizeof z));
  done = true;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Lighting of a surface-hit PRIMARY ray and the secondary rays it spawns (GXY_REVERSE_LIGHTING,
// TraceRays.ispc:625-923).  HitPoint = what those kernels read from the traced primary ray.
struct HitPoint {
  float ox, oy, oz, dx, dy, dz, t;  // the primary ray and its hit distance
  float3 sn;                        // shading normal (nx ny nz)
  float sr, sg, sb;                 // surface colour
  float o;                          // accumulated opacity BEFORE diffuseLighting
  int px, py;
};
struct SecRay {
  float3 org, dir;
  float r, g, b, tMax;
  int type;
};
__device__ __forceinline__ HitPoint load_hit_point(const Rays &R, int i) {
  HitPoint h;
  h.ox = R.ox[i]; h.oy = R.oy[i]; h.oz = R.oz[i]; h.dx = R.dx[i]; h.dy = R.dy[i]; h.dz = R.dz[i]; h.t = R.t[i];
  h.sn = f3(R.nx[i], R.ny[i], R.nz[i]);
  h.sr = R.sr[i]; h.sg = R.sg[i]; h.sb = R.sb[i]; h.o = R.o[i];
  h.px = R.x[i]; h.py = R.y[i];
  return h;
}
// TraceRays_generateAORays (:645-731), split in the part that depends on the hit only (origin and the
// tangent basis) and the part that depends on the ray number j.  tx/ty/tz: the AO direction tables
// (constant memory in the list kernels, a shared-memory copy in the persistent one).
__device__ __forceinline__ void ao_basis(const HitPoint &h, float epsilon, float3 &org, float3 &b0, float3 &b1) {
  const float3 sn = h.sn;
  b0 = f3(1.0f, 0.0f, 0.0f);
  if (fabsf(dot3(b0, sn)) > 0.95f) b0 = f3(0.0f, 1.0f, 0.0f);
  b1 = normalize_isp(cross3(b0, sn));
  b0 = normalize_isp(cross3(b1, sn));
  const float t = h.t;
  float ox = h.ox + t * h.dx, oy = h.oy + t * h.dy, oz = h.oz + t * h.dz;
  ox = ox + epsilon * sn.x; oy = oy + epsilon * sn.y; oz = oz + epsilon * sn.z;
  org = f3(ox, oy, oz);
}
__device__ __forceinline__ SecRay ao_ray_from_basis(const DevLights &L, float3 sn, float3 org, float3 b0, float3 b1, float sr, float sg,
                                                    float sb, float o, int px, int py, int j, float epsilon, const float *tx, const float *ty,
                                                    const float *tz) {
  const int nAO = L.n_ao;
  const float Ka = -L.Ka / nAO;  // GXY_REVERSE_LIGHTING
  const float ambient_scale = Ka * (1.0f - o);
  const int r = ((px * 9949 + py * 9613 + j * 9151) >> 8) & 0xff;
  const float x = tx[r], y = ty[r], z = tz[r] + epsilon;
  SecRay s;
  s.org = org;
  s.dir = x * b0 + y * b1 + z * sn;
  s.r = ambient_scale * sr; s.g = ambient_scale * sg; s.b = ambient_scale * sb;
  s.tMax = L.ao_radius;
  s.type = RAY_AO;
  return s;
}
__device__ __forceinline__ SecRay make_ao_ray(const DevLights &L, const HitPoint &h, int j, float epsilon) {
  float3 org, b0, b1;
  ao_basis(h, epsilon, org, b0, b1);
  return ao_ray_from_basis(L, h.sn, org, b0, b1, h.sr, h.sg, h.sb, h.o, h.px, h.py, j, epsilon, c_ao_x, c_ao_y, c_ao_z);
}
// TraceRays_ambientLighting + _diffuseLighting on (r,g,b,o) of the primary (:735-761, :859-923)
__device__ __forceinline__ void light_primary(const DevLights &L, const HitPoint &h, float &r, float &g, float &b, float &o) {
  const int nL = L.n_lights;
  {
    const float ambient_scale = L.Ka * (1.0f - o);
    r += ambient_scale * h.sr; g += ambient_scale * h.sg; b += ambient_scale * h.sb;
  }
  {
    const float Kd = L.Kd / nL;
    float tr = 0, tg = 0, tb = 0;
    for (int k = 0; k < nL; k++) {
      const float3 lt = f3(L.lights[k][0], L.lights[k][1], L.lights[k][2]);
      float3 lvec;
      if (L.types[k]) {
        const float3 sp = f3(h.ox + h.t * h.dx, h.oy + h.t * h.dy, h.oz + h.t * h.dz);
        lvec = safe_normalize(lt - sp);
      } else lvec = neg3(lt);
      const float d = dot3(h.sn, lvec);
      if (d > 0) {
        const float dff = (1.0f - o) * d;
        tr += dff * h.sr; tg += dff * h.sg; tb += dff * h.sb;
      }
    }
    r = r + Kd * (1 - o) * tr;
    g = g + Kd * (1 - o) * tg;
    b = b + Kd * (1 - o) * tb;
    o = o + Kd * (1 - o) * o;
  }
}
// TraceRays_generateShadowRays, light k (:813-854); o_lit = the opacity AFTER diffuseLighting
__device__ __forceinline__ SecRay make_shadow_ray(const DevLights &L, const HitPoint &h, int k, float epsilon, float o_lit) {
  const int nL = L.n_lights;
  const float Kd = -L.Kd / nL;  // GXY_REVERSE_LIGHTING
  const float3 sn = h.sn;
  const float3 sp = f3(h.ox + h.t * h.dx + epsilon * sn.x, h.oy + h.t * h.dy + epsilon * sn.y, h.oz + h.t * h.dz + epsilon * sn.z);
  const float3 lt = f3(L.lights[k][0], L.lights[k][1], L.lights[k][2]);
  float3 lvec;
  if (L.types[k]) lvec = safe_normalize(lt - sp);
  else lvec = neg3(lt);
  lvec = safe_normalize(lvec);
  float d = dot3(sn, lvec);
  if (d < 0) d = 0;
  const float dff = (1.0f - o_lit) * Kd * d;
  SecRay s;
  s.org = sp; s.dir = lvec;
  s.r = dff * h.sr; s.g = dff * h.sg; s.b = dff * h.sb;
  s.tMax = __int_as_float(0x7f800000);
  s.type = RAY_SHADOW;
  return s;
}

// ------------------------------------------------------------------------------------------------
// Box::exit_face (src/data/Box.cpp:84-97)
__device__ __forceinline__ int exit_face(float3 mn, float3 mx, float x, float y, float z, float dx, float dy, float dz) {
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

__device__ __forceinline__ int classify_box(float3 bmin, float3 bmax, const int *__restrict__ neighbors, int typ, int term, float ox, float oy,
                                            float oz, float dx, float dy, float dz) {
  int c = CLS_UNDETERMINED;
  if (typ == RAY_PRIMARY) {
    if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
    else if ((term & RAY_OPAQUE) | (term & RAY_TIMEOUT)) c = CLS_TERMINATED;
    else c = CLS_KEEP_HERE;
  } else if (typ == RAY_SHADOW) {
    if ((term & RAY_OPAQUE) | (term & RAY_SURFACE)) c = CLS_TERMINATED;
    else if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
    else c = CLS_DROP_ON_FLOOR;
  } else if (typ == RAY_AO) {
    if ((term & RAY_OPAQUE) | (term & RAY_SURFACE)) c = CLS_TERMINATED;
    else if (term & RAY_BOUNDARY) c = RAY_BOUNDARY;
    else c = CLS_DROP_ON_FLOOR;  // TIMEOUT or unknown
  }
  if (c == RAY_BOUNDARY) {
    const int f = exit_face(bmin, bmax, ox, oy, oz, dx, dy, dz);
    const int nb = neighbors[f];
    if (nb >= 0) c = nb;
    else c = (typ == RAY_SHADOW || typ == RAY_AO) ? CLS_DROP_ON_FLOOR : CLS_TERMINATED;
  }
  return c;
}
__device__ __forceinline__ int classify_values(const SceneParams &P, int typ, int term, float ox, float oy, float oz, float dx, float dy,
                                               float dz) {
  return classify_box(P.lmin, P.lmax, P.neighbors, typ, term, ox, oy, oz, dx, dy, dz);
}

__device__ __forceinline__ int classify_ray(const SceneParams &P, const Rays &R, int i) {
  return classify_values(P, R.type[i], R.term[i], R.ox[i], R.oy[i], R.oz[i], R.dx[i], R.dy[i], R.dz[i]);
}

// ------------------------------------------------------------------------------------------------
// Box::intersect (src/data/Box.cpp:107-150)
__device__ __forceinline__ bool box_intersect(float3 mn, float3 mx, float3 org, float3 dir, float &tmin, float &tmax) {
  tmin = (mn.x - org.x) / dir.x;
  tmax = (mx.x - org.x) / dir.x;
  if (tmin > tmax) { float s = tmax; tmax = tmin; tmin = s; }
  if (tmax < 0) return false;
  float tymin = (mn.y - org.y) / dir.y, tymax = (mx.y - org.y) / dir.y;
  if (tymin > tymax) { float s = tymax; tymax = tymin; tymin = s; }
  if (tymax < 0) return false;
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  float tzmin = (mn.z - org.z) / dir.z, tzmax = (mx.z - org.z) / dir.z;
  if (tzmin > tzmax) { float s = tzmax; tzmax = tzmin; tzmin = s; }
  if (tzmax < 0) return false;
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  if (tmin < 0) tmin = 0;
  return true;
}

// Camera::SpawnRays per pixel (Camera.cpp:403-441)
__device__ __forceinline__ bool spawn_pixel(const SceneParams &P, const DevCamera &a, int x, int y, float3 &vorigin, float3 &vray) {
  const float fx = ((float)x - a.off_x) * a.scaling;
  const float fy = ((float)y - a.off_y) * a.scaling;
  float3 xy;
  xy.x = a.center.x + fx * a.vr.x + fy * a.vu.x;
  xy.y = a.center.y + fx * a.vr.y + fy * a.vu.y;
  xy.z = a.center.z + fx * a.vr.z + fy * a.vu.z;
  if (a.ortho) { vorigin = xy - a.vdir; vray = a.vdir; }
  else { vorigin = a.veye; vray = xy - a.veye; normalize_gxy(vray); }
  float gmin, gmax, lmin = 0, lmax = 0;
  bool hit = box_intersect(P.gmin, P.gmax, vorigin, vray, gmin, gmax);
  if (hit) hit = box_intersect(P.lmin, P.lmax, vorigin, vray, lmin, lmax);
  const float d = fabsf(lmin) - fabsf(gmin);
  return hit && (lmax >= 0) && (d < 0.000001f) && (d > -0.000001f);
}
// the two halves of spawn_pixel for callers that test one ray against several partition boxes: the pixel's ray and its
// entry into the global box (false: the ray misses the data altogether) ...
__device__ __forceinline__ bool camera_ray(const SceneParams &P, const DevCamera &a, int x, int y, float3 &vorigin, float3 &vray, float &gmin) {
  const float fx = ((float)x - a.off_x) * a.scaling;
  const float fy = ((float)y - a.off_y) * a.scaling;
  float3 xy;
  xy.x = a.center.x + fx * a.vr.x + fy * a.vu.x;
  xy.y = a.center.y + fx * a.vr.y + fy * a.vu.y;
  xy.z = a.center.z + fx * a.vr.z + fy * a.vu.z;
  if (a.ortho) { vorigin = xy - a.vdir; vray = a.vdir; }
  else { vorigin = a.veye; vray = xy - a.veye; normalize_gxy(vray); }
  float gmax;
  return box_intersect(P.gmin, P.gmax, vorigin, vray, gmin, gmax);
}
// ... and "is this partition the one the ray enters first" (Camera.cpp:431-441)
__device__ __forceinline__ bool first_brick(float3 bmin, float3 bmax, float3 vorigin, float3 vray, float gmin) {
  float lmin = 0, lmax = 0;
  const bool hit = box_intersect(bmin, bmax, vorigin, vray, lmin, lmax);
  const float d = fabsf(lmin) - fabsf(gmin);
  return hit && (lmax >= 0) && (d < 0.000001f) && (d > -0.000001f);
}


}  // namespace gxy
