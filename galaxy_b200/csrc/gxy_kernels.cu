// gxy_kernels.cu -- the per-ray kernels of the hot path (sm_100a).
//
//  trace_kernel        K1  TraceRays_TraceRays             src/renderer/TraceRays.ispc:326-623
//                      K2/K4/K5 traversal + prim tests + postIntersect (gxy_traverse.cuh)
//                      K6/K7 volume sample / gradient / TF  (gxy_common.cuh)
//  ao_spawn_kernel     K9  TraceRays_generateAORays        TraceRays.ispc:625-733
//  light_shadow_kernel K10/K8 ambient+diffuse lighting, generateShadowRays  :735-923
//  classify_kernel     K11 Renderer::Classify/AssignDestinations  src/renderer/Renderer.cpp:304-454
//  accumulate_kernel   K13 Rendering::AddLocalPixels       src/renderer/Rendering.cpp:125-153
//  generate_*          K12 Camera::SpawnRays               src/renderer/Camera.cpp:379-493
//  tonemap_kernel      K14 ColorImageWriter::Write         src/renderer/ImageWriter.cpp:30-48
#include "gxy_internal.h"
#include "gxy_shade.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace gxy {

// ------------------------------------------------------------------------------------------------
struct SurfHit {  // TraceRays.ispc:79-88
  float t, opacity;
  float3 normal, color;
};

// Per-volume constants of the march loop, held in registers for the whole ray (the loop is issue bound: every
// constant-bank reload and every 64-bit address operation inside it costs an issue slot per sample).
struct VolK {
  float ox, oy, oz, rx, ry, rz, ux, uy, uz;
  unsigned nx, nxy;      // IDX32 only: row / slice pitch in elements
  const char *vox;
  int type;
};
#ifndef GXY_MARCH_PIN
#define GXY_MARCH_PIN 0
#endif
// keep a loop invariant in a register: ptxas re-materialises kernel parameters from the constant bank inside the loop
// (one LDC/LDCU issue slot each per sample) unless it cannot see where the value comes from -- a shuffle from the own lane
__device__ __forceinline__ void pin_rt(float &x) { x = __shfl_sync(0xffffffffu, x, threadIdx.x & 31u); }
__device__ __forceinline__ void pin_rt(unsigned &x) { x = __shfl_sync(0xffffffffu, x, threadIdx.x & 31u); }
#if GXY_MARCH_PIN
__device__ __forceinline__ void pin(float &x) { pin_rt(x); }
__device__ __forceinline__ void pin(unsigned &x) { pin_rt(x); }
#else
__device__ __forceinline__ void pin(float &x) {}
__device__ __forceinline__ void pin(unsigned &x) {}
#endif
__device__ __forceinline__ void volk_load(VolK &k, const DevVolume &v) {
  k.ox = v.origin.x; k.oy = v.origin.y; k.oz = v.origin.z;
  k.rx = v.rcp.x; k.ry = v.rcp.y; k.rz = v.rcp.z;
  k.ux = v.upper.x; k.uy = v.upper.y; k.uz = v.upper.z;
  k.nx = (unsigned)v.nx; k.nxy = (unsigned)v.nxy;
  k.vox = (const char *)v.vox; k.type = v.type;
  pin(k.ox); pin(k.oy); pin(k.oz); pin(k.rx); pin(k.ry); pin(k.rz); pin(k.ux); pin(k.uy); pin(k.uz); pin(k.nx); pin(k.nxy);
}

// SSV_sample_float_32 (SharedStructuredVolume.ispc:131-191): the arithmetic of vol_sample() (gxy_common.cuh), split
// into its three stages so that the march loop can run them one sample apart (software pipeline):
//   tap    cell index + interpolation weights; 32-bit element indices (volumes below 2^31 voxels): 2 IMAD instead
//          of 64-bit index math
//   fetch  the 8 voxel loads (or 4 L1 prefetches of the rows that hold them)
//   lerp   x -> y -> z as a + f*(b - a)
struct VolTap {
  unsigned o0;
  float fx, fy, fz;
};
struct Vox8 {
  float v000, v001, v010, v011, v100, v101, v110, v111;
};
__device__ __forceinline__ void vol_tap_k32(const VolK &v, float3 p, VolTap &t) {
  const float lx = v.rx * (p.x - v.ox), ly = v.ry * (p.y - v.oy), lz = v.rz * (p.z - v.oz);
  const float cx = fmaxf(0.0f, fminf(lx, v.ux)), cy = fmaxf(0.0f, fminf(ly, v.uy)), cz = fmaxf(0.0f, fminf(lz, v.uz));
  const int ix = (int)cx, iy = (int)cy, iz = (int)cz;
  t.fx = cx - (float)ix; t.fy = cy - (float)iy; t.fz = cz - (float)iz;
  t.o0 = (unsigned)ix + (unsigned)iy * v.nx + (unsigned)iz * v.nxy;
}
__device__ __forceinline__ void vol_fetch_k32(const VolK &v, const VolTap &t, Vox8 &x) {
  const unsigned o0 = t.o0, o1 = o0 + v.nx, o2 = o0 + v.nxy, o3 = o2 + v.nx;
  if (v.type == 0) {
    const float *__restrict__ b = (const float *)v.vox;
    const float *q0 = b + o0, *q1 = b + o1, *q2 = b + o2, *q3 = b + o3;
    x.v000 = __ldg(q0); x.v001 = __ldg(q0 + 1);
    x.v010 = __ldg(q1); x.v011 = __ldg(q1 + 1);
    x.v100 = __ldg(q2); x.v101 = __ldg(q2 + 1);
    x.v110 = __ldg(q3); x.v111 = __ldg(q3 + 1);
  } else {
    const unsigned char *__restrict__ b = (const unsigned char *)v.vox;
    const unsigned char *q0 = b + o0, *q1 = b + o1, *q2 = b + o2, *q3 = b + o3;
    x.v000 = (float)__ldg(q0); x.v001 = (float)__ldg(q0 + 1);
    x.v010 = (float)__ldg(q1); x.v011 = (float)__ldg(q1 + 1);
    x.v100 = (float)__ldg(q2); x.v101 = (float)__ldg(q2 + 1);
    x.v110 = (float)__ldg(q3); x.v111 = (float)__ldg(q3 + 1);
  }
}
__device__ __forceinline__ void vol_prefetch_k32(const VolK &v, const VolTap &t) {
  const unsigned o0 = t.o0, o1 = o0 + v.nx, o2 = o0 + v.nxy, o3 = o2 + v.nx;
  const unsigned sh = v.type == 0 ? 2u : 0u;
  const char *b = v.vox;
  asm volatile("prefetch.global.L1 [%0];" ::"l"(b + ((size_t)o0 << sh)));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(b + ((size_t)o1 << sh)));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(b + ((size_t)o2 << sh)));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(b + ((size_t)o3 << sh)));
}
__device__ __forceinline__ float vol_lerp8(const Vox8 &x, const VolTap &t) {
  const float v00 = x.v000 + t.fx * (x.v001 - x.v000);
  const float v01 = x.v010 + t.fx * (x.v011 - x.v010);
  const float v10 = x.v100 + t.fx * (x.v101 - x.v100);
  const float v11 = x.v110 + t.fx * (x.v111 - x.v110);
  const float v0 = v00 + t.fy * (v01 - v00);
  const float v1 = v10 + t.fy * (v11 - v10);
  return v0 + t.fz * (v1 - v0);
}
__device__ __forceinline__ float vol_sample_k32(const VolK &v, float3 p) {
  VolTap t;
  Vox8 x;
  vol_tap_k32(v, p, t);
  vol_fetch_k32(v, t, x);
  return vol_lerp8(x, t);
}

// tf_both / tf_opacity (gxy_common.cuh; LinearTransferFunction.ispc:19-97) with the value range held in registers
__device__ __forceinline__ float4 tf_both_k(const DevTF *__restrict__ tf, float lo, float hi, float d, float value) {
  if (isnan(value)) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (value <= lo) return __ldg(&tf->e[0]);
  if (value >= hi) return __ldg(&tf->e[255]);
  const float remapped = (value - lo) / d * 255.0f;
  const float fl = floorf(remapped);
  const int index = (int)fl;
  const float rem = remapped - fl;  // == remapped - (float)index
  const float4 a = __ldg(&tf->e[index]), b = __ldg(&tf->e[min(index + 1, 255)]);
  const float om = 1.0f - rem;
  return make_float4(om * a.x + rem * b.x, om * a.y + rem * b.y, om * a.z + rem * b.z, om * a.w + rem * b.w);
}
__device__ __forceinline__ float tf_opacity_k(const DevTF *__restrict__ tf, float lo, float hi, float d, float value) {
  if (isnan(value)) return 0.0f;
  if (value <= lo) return __ldg(&tf->e[0]).w;
  if (value >= hi) return __ldg(&tf->e[255]).w;
  const float remapped = (value - lo) / d * 255.0f;
  const float fl = floorf(remapped);
  const int index = (int)fl;
  const float rem = remapped - fl;
  return (1.0f - rem) * __ldg(&tf->e[index]).w + rem * __ldg(&tf->e[min(index + 1, 255)]).w;
}

template <int NV, bool IDX32>
__device__ __forceinline__ void sample_volumes(const SceneParams &P, const VolK *vk, int nvv, float3 coord, float *s) {
#pragma unroll
  for (int m = 0; m < NV; m++)
    if (m < nvv) s[m] = IDX32 ? vol_sample_k32(vk[m], coord) : vol_sample(P.vv[m].vol, coord);
}

// march-loop software pipeline (IDX32 volumes): 0 = none, 1 = L1 prefetch of the next sample's rows,
// 2 = the next sample's 8 voxels are loaded into registers while the current sample is processed
#ifndef GXY_MARCH_PIPE
#define GXY_MARCH_PIPE 0
#endif
#ifndef GXY_MARCH_BLOCKS
#define GXY_MARCH_BLOCKS 8
#endif

template <int NV, bool HAS_GEOM, bool IDX32, bool CURVES = false>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, (NV <= 2 && !HAS_GEOM) ? GXY_MARCH_BLOCKS : 1)
    trace_kernel(const __grid_constant__ SceneParams P, Rays R, int n, float global_epsilon, int *__restrict__ hit_ids,
                 int anyhit_secondary, unsigned long long *__restrict__ sample_counter, const int *__restrict__ n_dev) {
  __shared__ uint2 stack[HAS_GEOM ? GXY_STACK_SMEM * GXY_TRACE_THREADS : 1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // n_dev != NULL: the list's length lives on the device (frames in flight: no host round trip between the waves of a frame);
  // the launch is then sized for the list's capacity n
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  // launch_trace instantiates NV = n_volvis for up to 3 volume operators; only the catch-all has fewer than NV
  const int nvv = (NV <= 3) ? NV : (NV < P.n_volvis ? NV : P.n_volvis);
  float step = P.step;
  float epsilon = global_epsilon * step;  // TraceRays.ispc:359
#if GXY_MARCH_PIN
  pin_rt(step); pin_rt(epsilon);
#endif
  unsigned nsamples = 0;

  const bool shadeFlag = R.type[i] == RAY_PRIMARY;
  const float3 org = f3(R.ox[i], R.oy[i], R.oz[i]);
  float3 dir = f3(R.dx[i], R.dy[i], R.dz[i]);
  if (dir.x == 0.f) dir.x = 1e-6f;  // :377-379
  if (dir.y == 0.f) dir.y = 1e-6f;
  if (dir.z == 0.f) dir.z = 1e-6f;
  float ray_t0 = R.t[i], ray_t = R.tMax[i];
  const float tTimeout = ray_t;
  float cr = R.r[i], cg = R.g[i], cb = R.b[i], co = R.o[i];

  // MyIntersectBox :90-106, rcp(dir) := 1.0f/dir
  float tEntry, tExitVolume;
  {
    const float rx = 1.0f / dir.x, ry = 1.0f / dir.y, rz = 1.0f / dir.z;
    const float mnx = (P.lmin.x - org.x) * rx, mny = (P.lmin.y - org.y) * ry, mnz = (P.lmin.z - org.z) * rz;
    const float mxx = (P.lmax.x - org.x) * rx, mxy = (P.lmax.y - org.y) * ry, mxz = (P.lmax.z - org.z) * rz;
    tEntry = fmaxf(fminf(mnx, mxx), fmaxf(fminf(mny, mxy), fminf(mnz, mxz)));
    tExitVolume = fminf(fmaxf(mnx, mxx), fminf(fmaxf(mny, mxy), fmaxf(mnz, mxz)));
  }
  if (tEntry < ray_t0) tEntry = ray_t0;  // :412-413
  else if (tEntry > ray_t0) ray_t0 = tEntry;
  ray_t = fminf(ray_t, tExitVolume);  // :418

  SurfHit hit;
  hit.t = 0.f; hit.opacity = 0.f; hit.normal = f3(0.f, 0.f, 0.f); hit.color = f3(0.f, 0.f, 0.f);
  bool surface_hit = false;

  // LookForSliceHit :143-250
  if (NV > 0) {
    int vid = -1;
    for (int major = 0; major < nvv; major++) {
      const int ns = P.vv[major].n_slices;
      for (int minor = 0; minor < ns; minor++) {
        const float4 pl = P.vv[major].slices[minor];
        const float3 pnorm = f3(pl.x, pl.y, pl.z);
        const float denom = dot3(dir, pnorm);
        if (fabsf(denom) > 0.0001f) {
          const float t = (pl.w - dot3(org, pnorm)) / denom;
          if (t >= ray_t0 && t <= ray_t) {
            hit.normal = denom > 0 ? neg3(pnorm) : pnorm;
            hit.opacity = 1.0f; hit.t = t; vid = major; ray_t = t; surface_hit = true;
          }
        }
      }
    }
    if (surface_hit && shadeFlag) {
      const float3 point = org + hit.t * dir;
      const float s = vol_sample(P.vv[vid].vol, point);
      nsamples++;
      hit.color = tf_color(P.tfs + P.vv[vid].tf, s);
      hit.opacity = 1.0f;
    }
  }

  // LookForGeometryHit :108-141 + postIntersect (Model.ih:97-187)
  if (HAS_GEOM) {
    Hit1 h1;
    bool found;
    if (!shadeFlag && anyhit_secondary) found = traverse<true, CURVES>(P, org, dir, ray_t0, ray_t, h1, stack);
    else found = traverse<false, CURVES>(P, org, dir, ray_t0, ray_t, h1, stack);
    if (hit_ids) { hit_ids[2 * i] = found ? h1.geom : -1; hit_ids[2 * i + 1] = found ? h1.prim : -1; }
    if (found) {
      ray_t = h1.t;
      if (shadeFlag) {
        float3 col, Ns;
        float ca;
        shade_geometry_hit<CURVES>(P, h1, dir, col, ca, Ns);
        hit.color = col; hit.opacity = ca; hit.normal = Ns; hit.t = ray_t;
      }
      surface_hit = true;
    }
  } else if (hit_ids) {
    hit_ids[2 * i] = -1; hit_ids[2 * i + 1] = -1;
  }

  float tTermination = ray_t;

  if (NV > 0 && P.integrate) {  // :446-570
    constexpr int NVA = NV > 0 ? NV : 1;
    float tLast, tThis;
    float sLast[NVA], sThis[NVA];
    // loop invariants, fetched once per ray
    VolK vk[NVA];
    const DevTF *tf_vol[NVA];
    float tf_lo[NVA], tf_hi[NVA], tf_d[NVA];  // valueRange of the DVR transfer function and hi - lo
    float rate[NVA];
    bool dvr[NVA];
    bool any_iso = false;
#pragma unroll
    for (int m = 0; m < NV; m++)
      if (m < nvv) {
        if (IDX32) volk_load(vk[m], P.vv[m].vol);
        tf_vol[m] = P.tfs + P.vv[m].vol.tf;
        tf_lo[m] = __ldg(&tf_vol[m]->lo); tf_hi[m] = __ldg(&tf_vol[m]->hi);
        tf_d[m] = tf_hi[m] - tf_lo[m];
        rate[m] = P.vv[m].vol.samplingRate;
        dvr[m] = P.vv[m].volume_render != 0;
        any_iso = any_iso || P.vv[m].n_iso > 0;
      }
    unsigned iters = 0;
    bool hit_isosurface = false;
    tLast = tEntry + epsilon;
    bool opaque = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f);
    constexpr int PIPE = IDX32 ? GXY_MARCH_PIPE : 0;
    // pipeline state: taps (and voxels) of the sample at tSpec, the position the loop will visit next unless it ends
    VolTap tap[NVA];
    Vox8 vox[NVA];
    float tSpec = tEntry;
    if (PIPE) {
#pragma unroll
      for (int m = 0; m < NV; m++)
        if (m < nvv) {
          vol_tap_k32(vk[m], org + tSpec * dir, tap[m]);
          if (PIPE == 2) vol_fetch_k32(vk[m], tap[m], vox[m]);
        }
    }
    for (tThis = tEntry; tThis <= tTermination && !opaque && !hit_isosurface;
         tThis = (tThis == tEntry) ? (tEntry + epsilon)
                                   : (((tThis + step) > tTermination) && (tThis < tTermination)) ? tTermination : tThis + step) {
      if (PIPE) {
        // the loop update only depends on tThis and tTermination; when neither an isosurface hit nor opacity ends
        // the loop, the next position is the one predicted here (clamped addresses: any position is safe to load)
        const float tNext = (tThis == tEntry) ? (tEntry + epsilon)
                                              : (((tThis + step) > tTermination) && (tThis < tTermination)) ? tTermination : tThis + step;
        VolTap ntap[NVA];
        Vox8 nvox[NVA];
#pragma unroll
        for (int m = 0; m < NV; m++)
          if (m < nvv) {
            if (tThis != tSpec) {  // never taken by construction; keeps the pipeline honest
              vol_tap_k32(vk[m], org + tThis * dir, tap[m]);
              if (PIPE == 2) vol_fetch_k32(vk[m], tap[m], vox[m]);
            }
            vol_tap_k32(vk[m], org + tNext * dir, ntap[m]);
            if (PIPE == 2) vol_fetch_k32(vk[m], ntap[m], nvox[m]);
            else {
              vol_prefetch_k32(vk[m], ntap[m]);
              vol_fetch_k32(vk[m], tap[m], vox[m]);
            }
            sThis[m] = vol_lerp8(vox[m], tap[m]);
            tap[m] = ntap[m];
            if (PIPE == 2) vox[m] = nvox[m];
          }
        tSpec = tNext;
      } else {
        sample_volumes<NV, IDX32>(P, vk, nvv, org + tThis * dir, sThis);
      }
      iters++;
      if (tThis > tEntry && tLast >= epsilon) {
        if (any_iso) {
          // LookForIsoHit :252-310
          bool h = false;
          int vid = -1;
          float hsample = 0.f;
#pragma unroll
          for (int major = 0; major < NV; major++)
            if (major < nvv) {
              const float sl = sLast[major], st = sThis[major];
              const int ni = P.vv[major].n_iso;
              for (int minor = 0; minor < ni; minor++) {
                const float isoval = P.vv[major].iso[minor];
                if (((isoval >= sl) && (isoval < st)) || ((isoval <= sl) && (isoval > st))) {
                  h = true; vid = major;
                  hit.t = tLast + ((isoval - sl) / (st - sl)) * (tThis - tLast);
                  hsample = isoval;
                }
              }
            }
          if (h) {
            const float3 point = org + hit.t * dir;
            if (shadeFlag) {
              hit.normal = safe_normalize(vol_gradient(P.vv[vid].vol, point));
              nsamples += 4;
              if (dot3(dir, hit.normal) > 0) hit.normal = neg3(hit.normal);
              hit.color = tf_color(P.tfs + P.vv[vid].tf, hsample);
              hit.opacity = 1.0f;
            }
            tTermination = hit.t; tThis = hit.t;
            surface_hit = true; hit_isosurface = true;
            sample_volumes<NV, IDX32>(P, vk, nvv, org + tThis * dir, sThis);
            iters++;
          }
        }
        // DVR :512-542
#pragma unroll
        for (int major = 0; major < NV; major++)
          if (major < nvv) {
            if (dvr[major]) {
              const DevTF *tf = tf_vol[major];
              const float sVolume = (sLast[major] + sThis[major]) / 2;
              if (shadeFlag) {
                const float4 ca = tf_both_k(tf, tf_lo[major], tf_hi[major], tf_d[major], sVolume);
                if (ca.w > 0) {
                  // x / 1.0f == x for every x: the division is only executed for a sampling rate other than 1
                  const float wo = fmaxf(0.0f, fminf(rate[major] == 1.0f ? ca.w : ca.w / rate[major], 1.0f));
                  const float om = 1.0f - co;
                  cr = cr + om * (wo * ca.x); cg = cg + om * (wo * ca.y); cb = cb + om * (wo * ca.z); co = co + om * (wo * 1.0f);
                }
              } else {
                const float sampleOpacity = tf_opacity_k(tf, tf_lo[major], tf_hi[major], tf_d[major], sVolume);
                if (sampleOpacity > 0) {
                  const float so1 = rate[major] == 1.0f ? sampleOpacity : sampleOpacity / rate[major];
                  const float weightedOpacity = ((tThis - tLast) / step) * fmaxf(0.0f, fminf(so1, 1.0f));
                  const float f = 1 - weightedOpacity;
                  cr = cr * f; cg = cg * f; cb = cb * f; co = co * f;
                }
              }
            }
          }
      }
#pragma unroll
      for (int m = 0; m < NV; m++) sLast[m] = sThis[m];
      opaque = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f);
      if (opaque) tTermination = tThis;
      tLast = tThis;
    }
    nsamples += iters * (unsigned)nvv;
    ray_t = tTermination;
  }

  R.r[i] = cr; R.g[i] = cg; R.b[i] = cb; R.o[i] = co;
  int term = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f) ? RAY_OPAQUE : 0;
  R.t[i] = ray_t;
  if (surface_hit) {
    term |= RAY_SURFACE;
    // non-shaded rays: the reference reads an uninitialised hit.opacity (:591); every surface kind
    // sets opacity 1 when shaded, and Classify treats SURFACE like OPAQUE for SHADOW/AO rays.
    if (!shadeFlag || hit.opacity > 0.999f) term |= RAY_OPAQUE;
    if (shadeFlag) {
      R.sr[i] = hit.color.x; R.sg[i] = hit.color.y; R.sb[i] = hit.color.z; R.so[i] = 1.0f;
      R.nx[i] = hit.normal.x; R.ny[i] = hit.normal.y; R.nz[i] = hit.normal.z;
    }
  } else if (tTermination == tExitVolume) term |= RAY_BOUNDARY;
  else if (tTermination == tTimeout) term |= RAY_TIMEOUT;
  R.term[i] = term;

  if (sample_counter) {
    unsigned tot = nsamples;
    const unsigned mask = __activemask();
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_down_sync(mask, tot, off);
    // lanes that exited early make this an over/under count only if the leader is inactive
    if ((threadIdx.x & 31) == (__ffs(mask) - 1)) atomicAdd(sample_counter, (unsigned long long)tot);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent-warp variant of K1 for Visualizations with geometry only (no volume operators): the
// same per-ray arithmetic as trace_kernel<0, true>, but each lane fetches a new ray from a global
// queue as soon as FETCH_T lanes of its warp have finished theirs (Aila/Laine-style dynamic fetch),
// so that a warp is not held by its longest ray.  Per lane: setup (clip to the local box,
// TraceRays.ispc:377-418) -> trav_step()* -> finish (postIntersect + term, :563-610).
__device__ __forceinline__ void finish_geom_ray(const SceneParams &P, const Rays &R, const RayCtx &rc, const TravState &st,
                                                const PendingRay &pr, int *__restrict__ hit_ids) {
  const int i = pr.ray;
  const bool shadeFlag = R.type[i] == RAY_PRIMARY;
  const bool found = st.best_key != GXY_NO_HIT;
  const float cr = R.r[i], cg = R.g[i], cb = R.b[i], co = R.o[i];  // unchanged without volumes
  const float tTimeout = R.tMax[i];
  float ray_t = rc.tfar;
  int term = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f) ? RAY_OPAQUE : 0;
  if (found) {
    ray_t = st.best_t;
    term |= RAY_SURFACE;
    float opacity = 1.0f;
    if (shadeFlag) {
      Hit1 h1;
      trav_fetch_hit(P, rc, st, h1);
      float3 col, Ns;
      shade_geometry_hit(P, h1, rc.dir, col, opacity, Ns);
      R.sr[i] = col.x; R.sg[i] = col.y; R.sb[i] = col.z; R.so[i] = 1.0f;
      R.nx[i] = Ns.x; R.ny[i] = Ns.y; R.nz[i] = Ns.z;
    }
    if (!shadeFlag || opacity > 0.999f) term |= RAY_OPAQUE;  // see trace_kernel
  } else if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
  else if (ray_t == tTimeout) term |= RAY_TIMEOUT;
  R.t[i] = ray_t;
  R.term[i] = term;
  if (hit_ids) {
    hit_ids[2 * i] = found ? (int)(st.best_key >> 28) : -1;
    hit_ids[2 * i + 1] = found ? (int)(st.best_key & 0x0fffffffu) : -1;
  }
#ifdef GXY_TRAV_COUNTERS
  atomicAdd(P.trav_counters, (unsigned long long)st.n_nodes);
  atomicAdd(P.trav_counters + 1, (unsigned long long)st.n_prims);
#endif
}

template <int FETCH_T, int PRIM_T, int MIN_BLOCKS, int PREFETCH>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, MIN_BLOCKS)
    trace_geom_kernel(const __grid_constant__ SceneParams P, Rays R, int n, int *__restrict__ hit_ids, int anyhit_secondary) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  uint2 lstack[GXY_STACK_LOCAL];
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  RayCtx rc;
  TravState st;
  PendingRay pr;
  pr.ray = -1; pr.tExit = 0.f; pr.anyhit = false;
  st.tg.y = 0u; st.ng.y = 0u;
  bool trav = false, exhausted = false;
  while (true) {
    // ---- refill: finished lanes write their result and fetch the next ray of the queue
    const unsigned m_idle = __ballot_sync(FULL, !trav);
    if (m_idle == FULL || (!exhausted && __popc(m_idle) >= FETCH_T)) {
      bool ex_local = false;
      if (!trav) {
        if (pr.ray >= 0) {
          finish_geom_ray(P, R, rc, st, pr, hit_ids);
          pr.ray = -1;
        }
        if (!exhausted) {
          const int leader = __ffs((int)m_idle) - 1;
          const unsigned cnt = (unsigned)__popc(m_idle);
          unsigned base = 0;
          if ((int)lane == leader) base = atomicAdd(P.work_counter, cnt);
          base = __shfl_sync(m_idle, base, leader);
          const unsigned my = base + (unsigned)__popc(m_idle & ((1u << lane) - 1u));
          ex_local = base + cnt >= (unsigned)n;
          if (my < (unsigned)n) trav = setup_geom_ray(P, R, (int)my, anyhit_secondary, rc, st, pr);
        }
      }
      exhausted = exhausted || __any_sync(FULL, ex_local);
      if (__ballot_sync(FULL, pr.ray >= 0) == 0u) break;
    }
    // ---- one warp-synchronous phase: primitives when enough lanes hold a group (or nobody can do
    //      node work), else one node step for every lane that has a node to visit
    const bool has_prims = trav && st.tg.y != 0u;
    const bool can_node = trav && !has_prims;  // invariant: then st.ng.y > 0x00ffffff
    const unsigned mp = __ballot_sync(FULL, has_prims), mn = __ballot_sync(FULL, can_node);
    if (mp != 0u && (__popc(mp) >= PRIM_T || mn == 0u)) {
      if (has_prims) {
        if (prim_step(P, rc, st, pr.anyhit)) trav = false;
        else trav = trav_advance(st, stack, lstack);
      }
    } else if (can_node) {
      node_step<PREFETCH>(P, rc, st, stack, lstack);
      trav = trav_advance(st, stack, lstack);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cooperative variant: the node phase is per lane, but the primitive tests of a warp are spread over
// ALL its lanes.  After a node step ~4 lanes of a warp hold a primitive group of 1..8 records; tested
// by their owners alone that is a loop of up to 8 dependent DRAM round trips at 2 lanes of 32.  Here
// lane 4g+k tests the k-th pending primitive of the g-th owning lane (ray fetched with shuffles), so
// one pass tests up to 4 primitives of each of 8 owners with independent loads, and the owner picks
// the (t, geomID, primID)-smallest candidate of its 4 helpers with shuffles.
template <int FETCH_T, int MIN_BLOCKS, int PREFETCH>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, MIN_BLOCKS)
    trace_geom_coop_kernel(const __grid_constant__ SceneParams P, Rays R, int n, int *__restrict__ hit_ids, int anyhit_secondary) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  __shared__ unsigned char owner_of[GXY_TRACE_THREADS / 32][8];
  uint2 lstack[GXY_STACK_LOCAL];
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  RayCtx rc;
  TravState st;
  PendingRay pr;
  pr.ray = -1; pr.tExit = 0.f; pr.anyhit = false;
  st.tg.y = 0u; st.ng.y = 0u;
  rc.org = f3(0.f, 0.f, 0.f); rc.dir = f3(1.f, 1.f, 1.f); rc.tnear = 0.f; rc.tfar = 0.f;
  st.best_t = 0.f; st.best_key = GXY_NO_HIT;
  bool trav = false, exhausted = false;
  while (true) {
    // ---- refill: finished lanes write their result and fetch the next ray of the queue
    const unsigned m_idle = __ballot_sync(FULL, !trav);
    if (m_idle == FULL || (!exhausted && __popc(m_idle) >= FETCH_T)) {
      bool ex_local = false;
      if (!trav) {
        if (pr.ray >= 0) {
          finish_geom_ray(P, R, rc, st, pr, hit_ids);
          pr.ray = -1;
        }
        if (!exhausted) {
          const int leader = __ffs((int)m_idle) - 1;
          const unsigned cnt = (unsigned)__popc(m_idle);
          unsigned base = 0;
          if ((int)lane == leader) base = atomicAdd(P.work_counter, cnt);
          base = __shfl_sync(m_idle, base, leader);
          const unsigned my = base + (unsigned)__popc(m_idle & lt_mask);
          ex_local = base + cnt >= (unsigned)n;
          if (my < (unsigned)n) trav = setup_geom_ray(P, R, (int)my, anyhit_secondary, rc, st, pr);
        }
      }
      exhausted = exhausted || __any_sync(FULL, ex_local);
      if (__ballot_sync(FULL, pr.ray >= 0) == 0u) break;
    }
    // ---- node phase (invariant: a traversing lane has st.ng.y > 0x00ffffff and no pending primitives here)
    if (trav) node_step<PREFETCH>(P, rc, st, stack, lstack);
    // ---- cooperative primitive passes
    coop_prim_passes(P, rc, st, trav, pr.anyhit, owner_of[warp], lane, lt_mask);
    // ---- next node group (stack pop) or end of traversal
    if (trav) trav = trav_advance(st, stack, lstack);
  }
}

typedef void (*trace_geom_fn)(const SceneParams, Rays, int, int *, int);
struct TraceVariant {
  const char *name;
  trace_geom_fn fn;
  int min_blocks;
};
// tuning variants (GXY_TRACE_VARIANT=<name>): f = fetch threshold, p = primitive-phase threshold,
// b = resident blocks per SM the kernel is compiled for, n = no prefetch; the first one is the default
static const TraceVariant g_trace_variants[] = {
    {"c8b8", trace_geom_coop_kernel<8, 8, 0>, 8},      {"c8b8L2", trace_geom_coop_kernel<8, 8, 1>, 8},  {"c4b8", trace_geom_coop_kernel<4, 8, 0>, 8},
    {"c12b8", trace_geom_coop_kernel<12, 8, 0>, 8},    {"c8b6", trace_geom_coop_kernel<8, 6, 0>, 6},    {"c8b10", trace_geom_coop_kernel<8, 10, 0>, 10},
    {"f8p1b8", trace_geom_kernel<8, 1, 8, 0>, 8},      {"f8p1b8L2", trace_geom_kernel<8, 1, 8, 1>, 8},  {"f8p8b8", trace_geom_kernel<8, 8, 8, 0>, 8},
};

static int launch_trace_geom(const SceneParams &P, Rays R, int n, int *hit_ids, bool anyhit, cudaStream_t st) {
  // the environment is consulted on every launch so that a tuning script can sweep variants in one process
  const TraceVariant *variant = &g_trace_variants[0];
  if (const char *e = getenv("GXY_TRACE_VARIANT"))
    for (const TraceVariant &v : g_trace_variants)
      if (!strcmp(v.name, e)) variant = &v;
  int blocks_per_sm = variant->min_blocks;
  if (const char *e = getenv("GXY_TRACE_BLOCKS_PER_SM")) blocks_per_sm = atoi(e) > 0 ? atoi(e) : blocks_per_sm;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = GXY_SM_COUNT;
  }
  const int needed = (n + GXY_TRACE_THREADS - 1) / GXY_TRACE_THREADS;
  const int blocks = needed < sms * blocks_per_sm ? needed : sms * blocks_per_sm;
  GXY_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned), st));
  variant->fn<<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, hit_ids, anyhit ? 1 : 0);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

template <int NV>
static int launch_trace_nv(const SceneParams &P, Rays R, int n, float eps, int *hit_ids, bool anyhit, unsigned long long *sc,
                           cudaStream_t st, const int *n_dev) {
  const int blocks = (n + GXY_TRACE_THREADS - 1) / GXY_TRACE_THREADS;
  // 32-bit element indices in the sampler when every volume of the Visualization has fewer than 2^31 voxels
  bool idx32 = NV > 0;
  for (int m = 0; m < P.n_volvis && m < NV; m++) {
    const DevVolume &v = P.vv[m].vol;
    idx32 = idx32 && (unsigned long long)v.dims[0] * (unsigned long long)v.dims[1] * (unsigned long long)v.dims[2] < (1ull << 31);
  }
  if (P.n_prims > 0 && P.n_curves > 0) {
    // Visualizations with PathLines: the per-lane traversal with the curve test; 64-bit sampler indices (one instantiation per NV)
    trace_kernel<NV, true, false, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, eps, hit_ids, anyhit ? 1 : 0, sc, n_dev);
  } else if (P.n_prims > 0) {
    if (idx32) trace_kernel<NV, true, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, eps, hit_ids, anyhit ? 1 : 0, sc, n_dev);
    else trace_kernel<NV, true, false><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, eps, hit_ids, anyhit ? 1 : 0, sc, n_dev);
  } else {
    if (idx32) trace_kernel<NV, false, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, eps, hit_ids, 0, sc, n_dev);
    else trace_kernel<NV, false, false><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, R, n, eps, hit_ids, 0, sc, n_dev);
  }
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_trace(const SceneParams &P, Rays R, int n, float global_epsilon, int *hit_ids, bool anyhit_secondary,
                 unsigned long long *sample_counter, cudaStream_t st, const int *n_dev) {
  if (n <= 0) return 0;
  const char *pe = getenv("GXY_TRACE_PERSISTENT");
  const bool persistent = !(pe && atoi(pe) == 0);
  // (the persistent geometry kernel pulls rays through a host-zeroed queue head and knows its list length on the host)
  if (P.n_volvis == 0 && P.n_prims > 0 && P.n_curves == 0 && persistent && !n_dev) return launch_trace_geom(P, R, n, hit_ids, anyhit_secondary, st);
  switch (P.n_volvis) {
    case 0: return launch_trace_nv<0>(P, R, n, global_epsilon, hit_ids, anyhit_secondary, sample_counter, st, n_dev);
    case 1: return launch_trace_nv<1>(P, R, n, global_epsilon, hit_ids, anyhit_secondary, sample_counter, st, n_dev);
    case 2: return launch_trace_nv<2>(P, R, n, global_epsilon, hit_ids, anyhit_secondary, sample_counter, st, n_dev);
    case 3: return launch_trace_nv<3>(P, R, n, global_epsilon, hit_ids, anyhit_secondary, sample_counter, st, n_dev);
    default: return launch_trace_nv<GXY_MAX_VOLUME_VIS>(P, R, n, global_epsilon, hit_ids, anyhit_secondary, sample_counter, st, n_dev);
  }
}

// ------------------------------------------------------------------------------------------------
// nearest-hit only (gxy_intersect)
template <bool CURVES>
__global__ void __launch_bounds__(GXY_TRACE_THREADS)
    intersect_kernel(const __grid_constant__ SceneParams P, int n, const float *__restrict__ org3, const float *__restrict__ dir3,
                     const float *__restrict__ tnear, const float *__restrict__ tfar, int *__restrict__ gp, float *__restrict__ tuv) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Hit1 h;
  const float3 o = f3(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2]), d = f3(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]);
  const bool f = traverse<false, CURVES>(P, o, d, tnear[i], tfar[i], h, stack);
  gp[2 * i] = f ? h.geom : -1;
  gp[2 * i + 1] = f ? h.prim : -1;
  tuv[3 * i] = f ? h.t : tfar[i];
  tuv[3 * i + 1] = f ? h.u : 0.f;
  tuv[3 * i + 2] = f ? h.v : 0.f;
}

int launch_intersect(const SceneParams &P, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar,
                     int *geom_prim2, float *tuv3, cudaStream_t st) {
  if (n <= 0) return 0;
  const int blocks = (n + GXY_TRACE_THREADS - 1) / GXY_TRACE_THREADS;
  if (P.n_curves > 0) intersect_kernel<true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, n, org3, dir3, tnear, tfar, geom_prim2, tuv3);
  else intersect_kernel<false><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, n, org3, dir3, tnear, tfar, geom_prim2, tuv3);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ordered compaction helpers: 3 passes (block counts, scan of block counts, ranked write)
#define SCAN_THREADS 256
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  // v: this thread's count; returns exclusive prefix within the block
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
  for (int off = 1; off < 32; off <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    for (int off = 1; off < 32; off <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const int base = wid ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void scan_block_sums_kernel(int *__restrict__ sums, int nblocks, int *__restrict__ total_out) {
  // single block: exclusive scan of sums[0..nblocks) in place, total to *total_out
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int idx = base + threadIdx.x;
    const int v = idx < nblocks ? sums[idx] : 0;
    int tot;
    const int ex = block_exclusive_scan(v, &tot);
    const int c = carry;
    if (idx < nblocks) sums[idx] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// ---- hit scan (TraceRays.cpp:89-117): flag = PRIMARY && SURFACE
__device__ __forceinline__ int hit_flag(const Rays &R, int i, int n) {
  return (i < n && R.type[i] == RAY_PRIMARY && (R.term[i] & RAY_SURFACE)) ? 1 : 0;
}
__global__ void __launch_bounds__(SCAN_THREADS) hit_count_kernel(Rays R, int n, int *__restrict__ block_sums, const int *__restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) c += hit_flag(R, base + k, n);
  int tot;
  block_exclusive_scan(c, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_THREADS)
    hit_index_kernel(Rays R, int n, const int *__restrict__ block_sums, int *__restrict__ hit_index, int *__restrict__ hit_list,
                     const int *__restrict__ n_dev) {
  if (n_dev) n = min(n, *n_dev);
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int f[SCAN_ITEMS], c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { f[k] = hit_flag(R, base + k, n); c += f[k]; }
  int pos = block_sums[blockIdx.x] + block_exclusive_scan(c, nullptr);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) {
      hit_index[base + k] = f[k] ? pos : -1;
      if (f[k]) { hit_list[pos] = base + k; pos++; }
    }
}

int launch_hit_scan(Rays R, int n, int *d_hit_index, int *d_block_sums, int *d_nhit, cudaStream_t st, const int *n_dev) {
  // d_hit_index holds 2*n ints: [0,n) hit_index, [n,2n) hit_list
  const int nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
  hit_count_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(R, n, d_block_sums, n_dev);
  scan_block_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(d_block_sums, nblocks, d_nhit);
  hit_index_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(R, n, d_block_sums, d_hit_index, d_hit_index + n, n_dev);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// one thread per AO ray (TraceRays.ispc:645-731); runs BEFORE light_shadow_kernel (uses o before
// diffuseLighting updates it, as the reference's call order does, TraceRays.cpp:124-131)
__global__ void __launch_bounds__(256)
    ao_spawn_kernel(const __grid_constant__ DevLights L, Rays R, const int *__restrict__ hit_list, const int *__restrict__ d_nhit, Rays O,
                    float epsilon) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nAO = L.n_ao;
  const long long total = (long long)(*d_nhit) * nAO;
  if (k >= total) return;
  const int h = (int)(k / nAO), j = (int)(k - (long long)h * nAO);
  const HitPoint hp = load_hit_point(R, hit_list[h]);
  const SecRay s = make_ao_ray(L, hp, j, epsilon);
  O.ox[k] = s.org.x; O.oy[k] = s.org.y; O.oz[k] = s.org.z;
  O.dx[k] = s.dir.x; O.dy[k] = s.dir.y; O.dz[k] = s.dir.z;
  O.r[k] = s.r; O.g[k] = s.g; O.b[k] = s.b; O.o[k] = 0.0f;
  O.t[k] = 0.0f; O.tMax[k] = s.tMax;
  O.x[k] = hp.px; O.y[k] = hp.py; O.type[k] = RAY_AO; O.term[k] = 0;
}

// one thread per surface-hit primary: ambient (:735-761), diffuse (:859-923), shadow rays (:763-857)
__global__ void __launch_bounds__(256)
    light_shadow_kernel(const __grid_constant__ DevLights L, Rays R, const int *__restrict__ hit_list, const int *__restrict__ d_nhit,
                        Rays O, float epsilon) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  const int nhit = *d_nhit;
  if (h >= nhit) return;
  const int i = hit_list[h];
  const int nL = L.n_lights;
  const HitPoint hp = load_hit_point(R, i);
  float r = R.r[i], g = R.g[i], b = R.b[i], o = hp.o;
  light_primary(L, hp, r, g, b, o);
  R.r[i] = r; R.g[i] = g; R.b[i] = b; R.o[i] = o;
  // generateShadowRays (uses the o updated by diffuseLighting, as the reference does)
  if (L.shadows) {
    long long offset = (long long)nhit * L.n_ao + (long long)h * nL;
    for (int k = 0; k < nL; k++) {
      const SecRay s = make_shadow_ray(L, hp, k, epsilon, o);
      O.ox[offset] = s.org.x; O.oy[offset] = s.org.y; O.oz[offset] = s.org.z;
      O.dx[offset] = s.dir.x; O.dy[offset] = s.dir.y; O.dz[offset] = s.dir.z;
      O.r[offset] = s.r; O.g[offset] = s.g; O.b[offset] = s.b; O.o[offset] = 0.0f;
      O.t[offset] = 0.0f; O.tMax[offset] = s.tMax;
      O.x[offset] = hp.px; O.y[offset] = hp.py; O.type[offset] = RAY_SHADOW; O.term[offset] = 0;
      offset++;
    }
  }
}

int launch_shade_spawn(const DevLights &L, Rays R, int n, const int *d_hit_index, const int *d_nhit, Rays out, float epsilon,
                       cudaStream_t st, int max_hits) {
  // upper bounds for the grids (the exact hit count stays on the device); max_hits > 0: a tighter bound than the list's size
  // (a list of capacity n never holds more PRIMARY rays than the image has pixels)
  if (n <= 0) return 0;
  if (ensure_ao_tables()) return 1;
  const int *hit_list = d_hit_index + n;
  const int nh = max_hits > 0 && max_hits < n ? max_hits : n;
  if (L.n_ao > 0) {
    const long long total = (long long)nh * L.n_ao;
    const long long blocks = (total + 255) / 256;
    ao_spawn_kernel<<<(unsigned)blocks, 256, 0, st>>>(L, R, hit_list, d_nhit, out, epsilon);
  }
  light_shadow_kernel<<<(nh + 255) / 256, 256, 0, st>>>(L, R, hit_list, d_nhit, out, epsilon);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) classify_kernel(const __grid_constant__ SceneParams P, Rays R, int n, const int *__restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  R.classification[i] = classify_ray(P, R, i);
}

int launch_classify(const SceneParams &P, Rays R, int n, cudaStream_t st, const int *n_dev) {
  if (n <= 0) return 0;
  classify_kernel<<<(n + 255) / 256, 256, 0, st>>>(P, R, n, n_dev);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
    accumulate_kernel(Rays R, int n, float *__restrict__ fb, int w, int h, unsigned long long *__restrict__ d_terminated,
                      const int *__restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  const bool term = i < n && R.classification[i] == CLS_TERMINATED;
  if (term) {
    const int x = R.x[i], y = R.y[i];
    if (x >= 0 && x < w && y >= 0 && y < h) {
      float4 *p = reinterpret_cast<float4 *>(fb) + ((size_t)y * w + x);
      atomicAdd(p, make_float4(R.r[i], R.g[i], R.b[i], R.o[i]));  // red.global.add.v4.f32 (sm_90+)
    }
  }
  if (d_terminated) {
    const unsigned b = __ballot_sync(0xffffffffu, term);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(d_terminated, (unsigned long long)__popc(b));
  }
}

int launch_accumulate(Rays R, int n, float *fb, int w, int h, unsigned long long *d_terminated, cudaStream_t st, const int *n_dev) {
  if (n <= 0) return 0;
  accumulate_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, n, fb, w, h, d_terminated, n_dev);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// counting sort by destination (Renderer.cpp:561-618).  KEEP_HERE rays are re-queued locally by
// the caller mapping them to its own rank before this runs (keep_rank).
__global__ void __launch_bounds__(256) dest_count_kernel(Rays R, int n, int nranks, int keep_rank, int *__restrict__ counts) {
  extern __shared__ int s_counts[];
  for (int k = threadIdx.x; k < nranks; k += blockDim.x) s_counts[k] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int c = R.classification[i];
    if (c == CLS_KEEP_HERE) c = keep_rank;
    if (c >= 0 && c < nranks) atomicAdd(&s_counts[c], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nranks; k += blockDim.x)
    if (s_counts[k]) atomicAdd(&counts[k], s_counts[k]);
}
__global__ void dest_offsets_kernel(const int *__restrict__ counts, int nranks, int *__restrict__ offsets, int *__restrict__ cursor) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < nranks; k++) { offsets[k] = acc; cursor[k] = acc; acc += counts[k]; }
    offsets[nranks] = acc;
  }
}
__global__ void __launch_bounds__(256) dest_scatter_kernel(Rays R, int n, int nranks, int keep_rank, Rays O, int *__restrict__ cursor) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = R.classification[i];
  if (c == CLS_KEEP_HERE) c = keep_rank;
  if (c < 0 || c >= nranks) return;
  const int j = atomicAdd(&cursor[c], 1);
  // live columns only: surface/normal/sample/classification are dead across a hop
  O.ox[j] = R.ox[i]; O.oy[j] = R.oy[i]; O.oz[j] = R.oz[i];
  O.dx[j] = R.dx[i]; O.dy[j] = R.dy[i]; O.dz[j] = R.dz[i];
  O.r[j] = R.r[i]; O.g[j] = R.g[i]; O.b[j] = R.b[i]; O.o[j] = R.o[i];
  O.t[j] = R.t[i]; O.tMax[j] = R.tMax[i];
  O.x[j] = R.x[i]; O.y[j] = R.y[i]; O.type[j] = R.type[i]; O.term[j] = R.term[i];
}

int launch_partition_by_destination(Rays R, int n, int nranks, int keep_rank, Rays out, int *d_counts, int *d_offsets, int *d_cursor,
                                    cudaStream_t st) {
  GXY_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * nranks, st));
  if (n > 0) dest_count_kernel<<<(n + 255) / 256, 256, sizeof(int) * nranks, st>>>(R, n, nranks, keep_rank, d_counts);
  dest_offsets_kernel<<<1, 32, 0, st>>>(d_counts, nranks, d_offsets, d_cursor);
  if (n > 0) dest_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, n, nranks, keep_rank, out, d_cursor);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// Pixel of queue slot p.  tiles_x == 0: the reference's order (row-major, Camera.cpp:403-441).  tiles_x > 0: 16x8-pixel
// tiles of 128 slots (one trace CTA), each made of four 8x4 warp tiles: the rays of a warp / CTA stay a compact beam, so
// the voxels they sample at each march step share cache lines and the CTA re-uses them from L1 on the following steps
// (an oblique view in row-major order re-read the volume 3.7x from DRAM: ncu, profiles/).
__device__ __forceinline__ bool slot_pixel(int p, int w, int h, int tiles_x, int &x, int &y) {
  if (tiles_x == 0) { x = p % w; y = p / w; return true; }
  const int tile = p >> 7, i = p & 127, wq = i >> 5, lane = i & 31;
  x = (tile % tiles_x) * 16 + (wq & 1) * 8 + (lane & 7);
  y = (tile / tiles_x) * 8 + (wq >> 1) * 4 + (lane >> 3);
  return x < w && y < h;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    generate_count_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevCamera C, int w, int h, int tiles_x, int npix,
                          int *__restrict__ block_sums) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int p = base + k;
    float3 o, d;
    int x, y;
    if (p < npix && slot_pixel(p, w, h, tiles_x, x, y) && spawn_pixel(P, C, x, y, o, d)) c++;
  }
  int tot;
  block_exclusive_scan(c, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_THREADS)
    generate_write_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevCamera C, int w, int h, int tiles_x, int npix,
                          const int *__restrict__ block_sums, Rays O) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  float3 o[SCAN_ITEMS], d[SCAN_ITEMS];
  int px[SCAN_ITEMS], py[SCAN_ITEMS];
  bool f[SCAN_ITEMS];
  int c = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int p = base + k;
    f[k] = p < npix && slot_pixel(p, w, h, tiles_x, px[k], py[k]) && spawn_pixel(P, C, px[k], py[k], o[k], d[k]);
    c += f[k] ? 1 : 0;
  }
  int dst = block_sums[blockIdx.x] + block_exclusive_scan(c, nullptr);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (f[k]) {
      O.x[dst] = px[k]; O.y[dst] = py[k];
      O.ox[dst] = o[k].x; O.oy[dst] = o[k].y; O.oz[dst] = o[k].z;
      O.dx[dst] = d[k].x; O.dy[dst] = d[k].y; O.dz[dst] = d[k].z;
      O.r[dst] = 0; O.g[dst] = 0; O.b[dst] = 0; O.o[dst] = 0; O.t[dst] = 0; O.tMax[dst] = FLT_MAX;
      O.type[dst] = RAY_PRIMARY; O.term[dst] = 0;
      dst++;
    }
}

int launch_generate(const SceneParams &P, const DevCamera &C, int w, int h, bool tiled, Rays out, int *d_flags_scan, int *d_block_sums,
                    int *d_count, cudaStream_t st) {
  (void)d_flags_scan;
  const int tiles_x = tiled ? (w + 15) / 16 : 0;
  const int npix = tiled ? tiles_x * ((h + 7) / 8) * 128 : w * h;
  const int nblocks = (npix + SCAN_TILE - 1) / SCAN_TILE;
  generate_count_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(P, C, w, h, tiles_x, npix, d_block_sums);
  scan_block_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(d_block_sums, nblocks, d_count);
  generate_write_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(P, C, w, h, tiles_x, npix, d_block_sums, out);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tonemap_kernel(const float4 *__restrict__ fb, int w, int h, uchar4 *__restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= w * h) return;
  const int x = p % w, y = p / w;
  const float4 c = fb[p];
  // (unsigned char)(255*f): x86 cvttss2si (0x80000000 when out of range / NaN) then low byte
  auto cvt = [](float f) -> unsigned char {
    const float v = 255 * f;
    int iv;
    if (!(v > -2147483904.0f && v < 2147483648.0f)) iv = (int)0x80000000;
    else iv = (int)v;  // cvt.rzi
    return (unsigned char)(iv & 0xff);
  };
  out[(size_t)((h - 1) - y) * w + x] = make_uchar4(cvt(c.x), cvt(c.y), cvt(c.z), 0xff);
}
int launch_tonemap(const float *fb, int w, int h, unsigned char *rgba, cudaStream_t st) {
  tonemap_kernel<<<(w * h + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4 *>(fb), w, h, reinterpret_cast<uchar4 *>(rgba));
  GXY_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) fb_add_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = dst[i];
  const float4 b = src[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  dst[i] = a;
}
int launch_fb_add(float *dst, const float *src, size_t n, cudaStream_t st) {
  const size_t n4 = n / 4;
  fb_add_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4 *>(dst), reinterpret_cast<const float4 *>(src), n4);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) copy_rays_kernel(Rays D, size_t doff, Rays S, size_t soff, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = blockIdx.y;  // column
  float *const *df = &D.ox;
  float *const *sf = &S.ox;
  // the 5 int columns follow the 20 float columns in the struct; copy them as raw 32-bit words
  if (c < 20) df[c][doff + i] = sf[c][soff + i];
  else {
    int *const *di = &D.x;
    int *const *si = &S.x;
    di[c - 20][doff + i] = si[c - 20][soff + i];
  }
}
int launch_copy_rays(Rays dst, size_t dst_off, Rays src, size_t src_off, int n, cudaStream_t st) {
  if (n <= 0) return 0;
  dim3 grid((n + 255) / 256, 24);  // classification (column 24) is dead across a hop
  copy_rays_kernel<<<grid, 256, 0, st>>>(dst, dst_off, src, src_off, n);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gxy
