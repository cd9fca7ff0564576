// gxy_march_tma.cu -- TMA-staged variant of the volume march (TraceRays_TraceRays for a Visualization with ONE float volume
// operator and no geometry; src/renderer/TraceRays.ispc:326-623, SharedStructuredVolume.ispc:131-191).
//
// One CTA = 128 rays (on the frame path a 16x8-pixel tile of primaries).  The CTA advances its rays in lock step, K samples
// per stage.  For every stage the voxels the rays are going to need -- the bounding box, in voxel indices, of the segments
// their next K (+2, see below) steps cover -- are fetched as ONE 3-D box by the Tensor Memory Accelerator
// (cp.async.bulk.tensor.3d, completion on an mbarrier) into one of two shared-memory buffers, one stage ahead of the
// arithmetic; the trilinear sampler then reads its 8 voxels with LDS at compile-time offsets instead of 8 global gathers.
// The staged box is a CACHE, never a precondition: every sample tests whether its cell lies inside the box of the current
// stage and falls back to the global gathers of the plain kernel if not (rays of other directions in the list, beams wider
// than the box, oblique views), so the results are bit-identical to trace_kernel's by construction.
//
// Why: the plain kernel is bound by the latency of its gathers (8 warps per scheduler cannot cover ~1800 cycles per sample
// iteration); a bulk copy that runs ahead decouples that latency from the warps that do the arithmetic.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include "gxy_internal.h"

namespace gxy {

namespace {

struct SurfHitT {
  float t, opacity;
  float3 normal, color;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded: a bulk copy that never completes (a bad tensor map) must not hang the device; false after ~2^22 polls
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity) {
  for (unsigned tries = 0; tries < (1u << 22); tries++) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
               : "memory");
}

// min/max over the CTA of per-thread voxel-index boxes; threads without a ray pass an empty box (lo = INT_MAX, hi = INT_MIN)
struct BoxRed {
  int lo[3], hi[3], n_active;
};

}  // namespace

// BX x BY x BZ floats per stage, K samples per stage.  The box of stage s+1 is issued at the start of stage s from the rays'
// positions then: a ray at t = T takes its stage-(s+1) samples inside [T + (K-2) step, T + (2K-1) step] (every increment of the
// march is <= step, and at most one of them -- the first -- is the tiny epsilon), so the box covers K+1 steps along the rays.
template <int BX, int BY, int BZ, int K>
__global__ void __launch_bounds__(128)
    march_tma_kernel(const __grid_constant__ SceneParams P, const CUtensorMap *__restrict__ tmap, Rays R, int n, float global_epsilon,
                     unsigned long long *__restrict__ sample_counter, unsigned long long *__restrict__ staged_counter) {
  __shared__ __align__(128) float buf[2][BZ * BY * BX];
  __shared__ __align__(8) unsigned long long mbar[2];
  __shared__ int red_lo[4][3], red_hi[4][3], red_n[4];
  __shared__ int box_org[2][3];
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DevVolume &V = P.vv[0].vol;
  const float step = P.step;
  const float epsilon = global_epsilon * step;  // TraceRays.ispc:359
  const bool integrate = P.integrate != 0;
  // loop invariants
  const float vox_ox = V.origin.x, vox_oy = V.origin.y, vox_oz = V.origin.z;
  const float vox_rx = V.rcp.x, vox_ry = V.rcp.y, vox_rz = V.rcp.z;
  const float vox_ux = V.upper.x, vox_uy = V.upper.y, vox_uz = V.upper.z;
  const unsigned nx = (unsigned)V.nx, nxy = (unsigned)V.nxy;
  const float *__restrict__ vox = (const float *)V.vox;
  const DevTF *tf_vol = P.tfs + V.tf;
  const float tf_lo = __ldg(&tf_vol->lo), tf_hi = __ldg(&tf_vol->hi), tf_d = tf_hi - tf_lo;
  const float rate = V.samplingRate;
  const bool dvr = P.vv[0].volume_render != 0;
  const int n_iso = P.vv[0].n_iso;

  if (tid == 0) {
    mbar_init(&mbar[0], 1u);
    mbar_init(&mbar[1], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the barrier init must be visible to the async proxy (TMA)
  }

  // ---- the beginning of TraceRays_TraceRays for this thread's ray (:377-441), as in trace_kernel
  const int i = blockIdx.x * blockDim.x + tid;
  bool active = false, have_ray = i < n;
  bool shadeFlag = false, surface_hit = false, hit_isosurface = false, opaque = false;
  float3 org = f3(0.f, 0.f, 0.f), dir = f3(1.f, 1.f, 1.f);
  float tEntry = 0.f, tExitVolume = 0.f, tTimeout = 0.f, tTermination = 0.f, tThis = 0.f, tLast = 0.f;
  float cr = 0.f, cg = 0.f, cb = 0.f, co = 0.f, sLast = 0.f, sThis = 0.f;
  SurfHitT hit;
  hit.t = 0.f; hit.opacity = 0.f; hit.normal = f3(0.f, 0.f, 0.f); hit.color = f3(0.f, 0.f, 0.f);
  unsigned nsamples = 0, nstaged = 0;
  if (have_ray) {
    shadeFlag = R.type[i] == RAY_PRIMARY;
    org = f3(R.ox[i], R.oy[i], R.oz[i]);
    dir = f3(R.dx[i], R.dy[i], R.dz[i]);
    if (dir.x == 0.f) dir.x = 1e-6f;  // :377-379
    if (dir.y == 0.f) dir.y = 1e-6f;
    if (dir.z == 0.f) dir.z = 1e-6f;
    float ray_t0 = R.t[i], ray_t = R.tMax[i];
    tTimeout = ray_t;
    cr = R.r[i]; cg = R.g[i]; cb = R.b[i]; co = R.o[i];
    {  // MyIntersectBox :90-106
      const float rx = 1.0f / dir.x, ry = 1.0f / dir.y, rz = 1.0f / dir.z;
      const float mnx = (P.lmin.x - org.x) * rx, mny = (P.lmin.y - org.y) * ry, mnz = (P.lmin.z - org.z) * rz;
      const float mxx = (P.lmax.x - org.x) * rx, mxy = (P.lmax.y - org.y) * ry, mxz = (P.lmax.z - org.z) * rz;
      tEntry = fmaxf(fminf(mnx, mxx), fmaxf(fminf(mny, mxy), fminf(mnz, mxz)));
      tExitVolume = fminf(fmaxf(mnx, mxx), fminf(fmaxf(mny, mxy), fmaxf(mnz, mxz)));
    }
    if (tEntry < ray_t0) tEntry = ray_t0;  // :412-413
    else if (tEntry > ray_t0) ray_t0 = tEntry;
    ray_t = fminf(ray_t, tExitVolume);  // :418
    {  // LookForSliceHit :143-250
      const int ns = P.vv[0].n_slices;
      for (int minor = 0; minor < ns; minor++) {
        const float4 pl = P.vv[0].slices[minor];
        const float3 pnorm = f3(pl.x, pl.y, pl.z);
        const float denom = dot3(dir, pnorm);
        if (fabsf(denom) > 0.0001f) {
          const float t = (pl.w - dot3(org, pnorm)) / denom;
          if (t >= ray_t0 && t <= ray_t) {
            hit.normal = denom > 0 ? neg3(pnorm) : pnorm;
            hit.opacity = 1.0f; hit.t = t; ray_t = t; surface_hit = true;
          }
        }
      }
      if (surface_hit && shadeFlag) {
        const float sv = vol_sample(V, org + hit.t * dir);
        nsamples++;
        hit.color = tf_color(P.tfs + P.vv[0].tf, sv);
        hit.opacity = 1.0f;
      }
    }
    tTermination = ray_t;
    tLast = tEntry + epsilon;
    opaque = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f);
    tThis = tEntry;
    active = integrate;
  }

  // voxel-index box of the ray's samples for t in [ta, tb] (cells ix..ix+1 etc.), clamped like the sampler clamps
  auto ray_box = [&](float ta, float tb, int lo[3], int hi[3]) {
    const float ax = fmaxf(0.0f, fminf(vox_rx * ((org.x + ta * dir.x) - vox_ox), vox_ux)), bx = fmaxf(0.0f, fminf(vox_rx * ((org.x + tb * dir.x) - vox_ox), vox_ux));
    const float ay = fmaxf(0.0f, fminf(vox_ry * ((org.y + ta * dir.y) - vox_oy), vox_uy)), by = fmaxf(0.0f, fminf(vox_ry * ((org.y + tb * dir.y) - vox_oy), vox_uy));
    const float az = fmaxf(0.0f, fminf(vox_rz * ((org.z + ta * dir.z) - vox_oz), vox_uz)), bz = fmaxf(0.0f, fminf(vox_rz * ((org.z + tb * dir.z) - vox_oz), vox_uz));
    lo[0] = (int)fminf(ax, bx); hi[0] = (int)fmaxf(ax, bx) + 1;
    lo[1] = (int)fminf(ay, by); hi[1] = (int)fmaxf(ay, by) + 1;
    lo[2] = (int)fminf(az, bz); hi[2] = (int)fmaxf(az, bz) + 1;
  };
  // CTA-wide union; every thread returns the same result.  One barrier inside; the caller's next barrier frees red_* again.
  auto cta_box = [&](bool mine, const int lo[3], const int hi[3]) {
    BoxRed r;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int l = __reduce_min_sync(FULL, mine ? lo[a] : 0x7fffffff);
      const int h = __reduce_max_sync(FULL, mine ? hi[a] : (int)0x80000000);
      if (lane == 0) { red_lo[warp][a] = l; red_hi[warp][a] = h; }
    }
    const int cnt = __popc(__ballot_sync(FULL, mine));
    if (lane == 0) red_n[warp] = cnt;
    __syncthreads();
    r.n_active = red_n[0] + red_n[1] + red_n[2] + red_n[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      r.lo[a] = min(min(red_lo[0][a], red_lo[1][a]), min(red_lo[2][a], red_lo[3][a]));
      r.hi[a] = max(max(red_hi[0][a], red_hi[1][a]), max(red_hi[2][a], red_hi[3][a]));
    }
    return r;  // red_* are written again only after the caller's next CTA barrier (end of the stage / after the first issue)
  };
  constexpr unsigned STAGE_BYTES = (unsigned)(BX * BY * BZ * sizeof(float));
  // thread 0: box origin = low corner of the union (a beam larger than the box is cut off).  The innermost coordinate of a
  // tiled TMA load must be a multiple of 16 bytes (measured: any other x raises "illegal instruction", tools/ubench/tma1d.cu),
  // so the box starts at x rounded down to a multiple of 4 voxels.
  auto issue = [&](int b, const BoxRed &r) {
    const int x0 = r.lo[0] & ~3;
    box_org[b][0] = x0; box_org[b][1] = r.lo[1]; box_org[b][2] = r.lo[2];
    mbar_expect_tx(&mbar[b], STAGE_BYTES);
    tma_load_3d(&buf[b][0], tmap, x0, r.lo[1], r.lo[2], &mbar[b]);
  };

  __syncthreads();  // mbarriers initialised
  int lo[3], hi[3];
  // ---- box of stage 0
  ray_box(tThis, tThis + (K - 1) * step, lo, hi);
  BoxRed br = cta_box(active, lo, hi);
  unsigned phase[2] = {0u, 0u};
  bool pending[2] = {false, false};
  if (br.n_active > 0) {
    if (tid == 0) issue(0, br);
    pending[0] = true;
  }
  __syncthreads();  // box_org[0] visible

  for (int s = 0; br.n_active > 0; s++) {
    const int b = s & 1;
    // ---- issue the box of stage s+1 into the other buffer (everybody finished reading it at the end of stage s-1)
    ray_box(tThis + (K - 2) * step, tThis + (2 * K - 1) * step, lo, hi);
    const BoxRed nxt = cta_box(active, lo, hi);
    if (nxt.n_active > 0) {
      if (tid == 0) issue(b ^ 1, nxt);
      pending[b ^ 1] = true;
    }
    // ---- wait for this stage's voxels
    if (!mbar_wait(&mbar[b], phase[b])) *P.error_flag = 6;
    phase[b] ^= 1u;
    pending[b] = false;
    const int ox = box_org[b][0], oy = box_org[b][1], oz = box_org[b][2];
    const float *__restrict__ sb = &buf[b][0];
    // ---- K iterations of the march loop (:462-560), identical to trace_kernel's except for where the 8 voxels come from
#pragma unroll 1
    for (int k = 0; k < K && active; k++) {
      if (!(tThis <= tTermination && !opaque && !hit_isosurface)) { active = false; break; }
      {
        const float3 p = org + tThis * dir;
        const float lx = vox_rx * (p.x - vox_ox), ly = vox_ry * (p.y - vox_oy), lz = vox_rz * (p.z - vox_oz);
        const float cx = fmaxf(0.0f, fminf(lx, vox_ux)), cy = fmaxf(0.0f, fminf(ly, vox_uy)), cz = fmaxf(0.0f, fminf(lz, vox_uz));
        const int ix = (int)cx, iy = (int)cy, iz = (int)cz;
        const float fx = cx - (float)ix, fy = cy - (float)iy, fz = cz - (float)iz;
        float v000, v001, v010, v011, v100, v101, v110, v111;
        const unsigned ux = (unsigned)(ix - ox), uy = (unsigned)(iy - oy), uz = (unsigned)(iz - oz);
        if (ux < (unsigned)(BX - 1) && uy < (unsigned)(BY - 1) && uz < (unsigned)(BZ - 1)) {
          const float *q = sb + ((uz * BY + uy) * BX + ux);
          v000 = q[0]; v001 = q[1]; v010 = q[BX]; v011 = q[BX + 1];
          v100 = q[BY * BX]; v101 = q[BY * BX + 1]; v110 = q[BY * BX + BX]; v111 = q[BY * BX + BX + 1];
          nstaged++;
        } else {
          const unsigned o0 = (unsigned)ix + (unsigned)iy * nx + (unsigned)iz * nxy;
          const float *q0 = vox + o0, *q1 = q0 + nx, *q2 = q0 + nxy, *q3 = q2 + nx;
          v000 = __ldg(q0); v001 = __ldg(q0 + 1); v010 = __ldg(q1); v011 = __ldg(q1 + 1);
          v100 = __ldg(q2); v101 = __ldg(q2 + 1); v110 = __ldg(q3); v111 = __ldg(q3 + 1);
        }
        const float v00 = v000 + fx * (v001 - v000);
        const float v01 = v010 + fx * (v011 - v010);
        const float v10 = v100 + fx * (v101 - v100);
        const float v11 = v110 + fx * (v111 - v110);
        const float v0 = v00 + fy * (v01 - v00);
        const float v1 = v10 + fy * (v11 - v10);
        sThis = v0 + fz * (v1 - v0);
      }
      nsamples++;
      if (tThis > tEntry && tLast >= epsilon) {
        if (n_iso > 0) {  // LookForIsoHit :252-310
          bool h = false;
          float hsample = 0.f;
          for (int minor = 0; minor < n_iso; minor++) {
            const float isoval = P.vv[0].iso[minor];
            if (((isoval >= sLast) && (isoval < sThis)) || ((isoval <= sLast) && (isoval > sThis))) {
              h = true;
              hit.t = tLast + ((isoval - sLast) / (sThis - sLast)) * (tThis - tLast);
              hsample = isoval;
            }
          }
          if (h) {
            const float3 point = org + hit.t * dir;
            if (shadeFlag) {
              hit.normal = safe_normalize(vol_gradient(V, point));
              nsamples += 4;
              if (dot3(dir, hit.normal) > 0) hit.normal = neg3(hit.normal);
              hit.color = tf_color(P.tfs + P.vv[0].tf, hsample);
              hit.opacity = 1.0f;
            }
            tTermination = hit.t; tThis = hit.t;
            surface_hit = true; hit_isosurface = true;
            sThis = vol_sample(V, org + tThis * dir);
            nsamples++;
          }
        }
        if (dvr) {  // :512-542
          const float sVolume = (sLast + sThis) / 2;
          if (shadeFlag) {
            const float4 ca = tf_both(tf_vol, sVolume);
            if (ca.w > 0) {
              const float wo = fmaxf(0.0f, fminf(rate == 1.0f ? ca.w : ca.w / rate, 1.0f));
              const float om = 1.0f - co;
              cr = cr + om * (wo * ca.x); cg = cg + om * (wo * ca.y); cb = cb + om * (wo * ca.z); co = co + om * (wo * 1.0f);
            }
          } else {
            const float sampleOpacity = tf_opacity(tf_vol, sVolume);
            if (sampleOpacity > 0) {
              const float so1 = rate == 1.0f ? sampleOpacity : sampleOpacity / rate;
              const float weightedOpacity = ((tThis - tLast) / step) * fmaxf(0.0f, fminf(so1, 1.0f));
              const float f = 1 - weightedOpacity;
              cr = cr * f; cg = cg * f; cb = cb * f; co = co * f;
            }
          }
        }
      }
      sLast = sThis;
      opaque = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f);
      if (opaque) tTermination = tThis;
      tLast = tThis;
      tThis = (tThis == tEntry) ? (tEntry + epsilon)
                                : (((tThis + step) > tTermination) && (tThis < tTermination)) ? tTermination : tThis + step;
    }
    // everybody is done with buf[b]: order these generic-proxy reads before the async-proxy write of the next fill
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    br = nxt;
  }
  // a box that was issued but never consumed must land before the CTA's shared memory goes away
  for (int b = 0; b < 2; b++)
    if (pending[b]) (void)mbar_wait(&mbar[b], phase[b]);

  // ---- the end of TraceRays_TraceRays (:563-610)
  if (have_ray) {
    (void)tf_lo; (void)tf_hi; (void)tf_d;
    R.r[i] = cr; R.g[i] = cg; R.b[i] = cb; R.o[i] = co;
    int term = (min3f(cr, cg, cb) >= 1.0f || co > 0.999f) ? RAY_OPAQUE : 0;
    R.t[i] = tTermination;
    if (surface_hit) {
      term |= RAY_SURFACE;
      if (!shadeFlag || hit.opacity > 0.999f) term |= RAY_OPAQUE;  // see trace_kernel
      if (shadeFlag) {
        R.sr[i] = hit.color.x; R.sg[i] = hit.color.y; R.sb[i] = hit.color.z; R.so[i] = 1.0f;
        R.nx[i] = hit.normal.x; R.ny[i] = hit.normal.y; R.nz[i] = hit.normal.z;
      }
    } else if (tTermination == tExitVolume) term |= RAY_BOUNDARY;
    else if (tTermination == tTimeout) term |= RAY_TIMEOUT;
    R.term[i] = term;
  }
  if (sample_counter) {
    unsigned tot = nsamples, st = nstaged;
    for (int off = 16; off > 0; off >>= 1) { tot += __shfl_down_sync(FULL, tot, off); st += __shfl_down_sync(FULL, st, off); }
    if (lane == 0) {
      atomicAdd(sample_counter, (unsigned long long)tot);
      if (staged_counter) atomicAdd(staged_counter, (unsigned long long)st);
    }
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

bool march_tma_eligible(const SceneParams &P) {
  if (P.n_volvis != 1 || P.n_prims > 0) return false;
  const DevVolume &v = P.vv[0].vol;
  if (v.type != 0) return false;
  if ((unsigned long long)v.dims[0] * (unsigned long long)v.dims[1] * (unsigned long long)v.dims[2] >= (1ull << 31)) return false;
  if (v.dims[0] % 4 != 0) return false;  // TMA: global strides must be multiples of 16 bytes
  if ((reinterpret_cast<uintptr_t>(v.vox) & 15u) != 0) return false;
  return encode_tiled() != nullptr;
}

template <int BX, int BY, int BZ, int K>
static int launch_variant(const SceneParams &P, Rays R, int n, float eps, unsigned long long *sc, unsigned long long *staged, cudaStream_t st) {
  const DevVolume &v = P.vv[0].vol;
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)v.dims[0], (cuuint64_t)v.dims[1], (cuuint64_t)v.dims[2]};
  const cuuint64_t gstride[2] = {(cuuint64_t)v.dims[0] * 4u, (cuuint64_t)v.dims[0] * (cuuint64_t)v.dims[1] * 4u};
  const cuuint32_t box[3] = {BX, BY, BZ};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(v.vox), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gxy_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 1;
  }
  // the descriptor lives in global memory (a ring of slots: launches on different streams must not overwrite each other's):
  // SceneParams alone fills most of the 4 KB of classic kernel-parameter space
  static CUtensorMap *ring = nullptr;
  static unsigned slot = 0;
  constexpr unsigned RING = 64;
  if (!ring) GXY_CUDA(cudaMalloc(&ring, sizeof(CUtensorMap) * RING));
  CUtensorMap *d_map = ring + (slot++ % RING);
  GXY_CUDA(cudaMemcpyAsync(d_map, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
  const int blocks = (n + 127) / 128;
  march_tma_kernel<BX, BY, BZ, K><<<blocks, 128, 0, st>>>(P, d_map, R, n, eps, sc, staged);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

// axis: the volume axis the beams mostly run along (0 x, 1 y, 2 z): the staged box is thin along it
int launch_march_tma(const SceneParams &P, Rays R, int n, float eps, int axis, unsigned long long *sample_counter, unsigned long long *staged_counter,
                     cudaStream_t st) {
  if (n <= 0) return 0;
  switch (axis) {
    case 0: return launch_variant<8, 16, 32, 4>(P, R, n, eps, sample_counter, staged_counter, st);
    case 1: return launch_variant<32, 8, 16, 4>(P, R, n, eps, sample_counter, staged_counter, st);
    default: {
      // box shape / samples per stage for beams along z, measured on C3-1024 (plain gather kernel: 6.99 ms):
      //   0: 32x16x8, K=4   8.38 ms (96 % of the samples from staged boxes)
      //   1: 32x14x10, K=6  7.78 ms (96 %)      <- default
      //   2: 28x14x12, K=8  7.99 ms (91 %)
      int variant = 1;
      if (const char *e = getenv("GXY_TMA_VARIANT")) variant = atoi(e);
      if (variant == 0) return launch_variant<32, 16, 8, 4>(P, R, n, eps, sample_counter, staged_counter, st);
      if (variant == 2) return launch_variant<28, 14, 12, 8>(P, R, n, eps, sample_counter, staged_counter, st);
      return launch_variant<32, 14, 10, 6>(P, R, n, eps, sample_counter, staged_counter, st);
    }
  }
}

}  // namespace gxy
