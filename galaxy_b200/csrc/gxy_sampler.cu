// gxy_sampler.cu -- the Sampler's ray kernel: rays march through the volume bricks of a sampling
// Visualization and stop wherever a sampler operator fires; the hit points become Particles.
//
//   sampler_trace_kernel   SamplerTraceRays_SamplerTraceRays   src/sampler/SamplerTraceRays.ispc:128-222
//     gradient operator    GradientSamplerVis_init / _check_interval   src/sampler/GradientSamplerVis.ispc:36-72
//     iso operator         IsoSamplerVis_init / _check_interval        src/sampler/IsoSamplerVis.ispc:36-67
//     sample points        Sampler::HandleTerminatedRays               src/sampler/Sampler.cpp:52-92
//
// One thread per ray; a CTA of the frame path is a 16x8-pixel tile (launch_generate's tiled order), so the 8 (iso) or 32
// (gradient) voxel gathers of a step share cache lines across the warp as in the march kernel.  The reference keeps
// sLast / tLast / tHit as varying members of the operator struct shared by all threads; they are per-ray registers here.
// The hit point is appended to the partition's sample buffer by the same kernel (one atomicAdd per hit), fused with the
// trace instead of a second pass over the list under a mutex.  rcp(dir) := 1/dir, no FMA contraction (-fmad=false).
#include "gxy_internal.h"

namespace gxy {

// LOOP = false: one pass per ray, as the reference's kernel.  LOOP = true (GXY_SAMPLER_LOOP=1, opt-in): a ray that left a sample is
// KEEP_HERE in Renderer::Classify and would come straight back with t = the sample's t; the thread does that next pass itself,
// with the same arithmetic (the march restarts from max(EntryT, t)), until the ray reaches the boundary -- one launch per visit
// of a partition instead of one per crossing.  passes: number of passes (= the reference's traced-ray count).
template <bool LOOP>
__global__ void __launch_bounds__(128)
    sampler_trace_kernel(const __grid_constant__ SamplerParams SP, Rays R, int n, float *__restrict__ samples,
                         unsigned long long *__restrict__ sample_count, unsigned long long sample_cap,
                         unsigned long long *__restrict__ passes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nv = SP.n_ops;
  const float step = SP.step;
  const float3 org = f3(R.ox[i], R.oy[i], R.oz[i]);
  const float3 dir0 = f3(R.dx[i], R.dy[i], R.dz[i]);
  float3 dir = dir0;
  float ray_t = R.t[i];
  unsigned my_passes = 0;
  if (dir.x == 0.f) dir.x = 1e-6f;  // :163-165
  if (dir.y == 0.f) dir.y = 1e-6f;
  if (dir.z == 0.f) dir.z = 1e-6f;
  // EntryT / ExitT (:62-86)
  const float rx = 1.0f / dir.x, ry = 1.0f / dir.y, rz = 1.0f / dir.z;
  const float mnx = (SP.lmin.x - org.x) * rx, mny = (SP.lmin.y - org.y) * ry, mnz = (SP.lmin.z - org.z) * rz;
  const float mxx = (SP.lmax.x - org.x) * rx, mxy = (SP.lmax.y - org.y) * ry, mxz = (SP.lmax.z - org.z) * rz;
  const float tEntryBox = fmaxf(fminf(mnx, mxx), fmaxf(fminf(mny, mxy), fminf(mnz, mxz)));
  const float tExit = fminf(fmaxf(mnx, mxx), fminf(fmaxf(mny, mxy), fmaxf(mnz, mxz)));
  for (;;) {  // one iteration = one pass of the reference's kernel over this ray
  my_passes++;
  float tEntry = tEntryBox;
  if (tEntry < ray_t) tEntry = ray_t;  // :176-179
  float tThis = tEntry + step;
  int hit = -1;

  float3 gLast[GXY_MAX_VOLUME_VIS];
  float sLast[GXY_MAX_VOLUME_VIS], tLast[GXY_MAX_VOLUME_VIS];
  for (int m = 0; m < nv; m++) {  // init (:186-190)
    const float3 coord = org + tEntry * dir;
    if (SP.op[m].kind == 0) gLast[m] = vol_gradient(SP.op[m].vol, coord);
    else sLast[m] = vol_sample(SP.op[m].vol, coord);
    tLast[m] = tEntry;
  }
  while (tThis <= tExit && hit == -1) {  // :192-211
    for (int m = 0; m < nv && hit == -1; m++) {
      const float3 coord = org + tThis * dir;
      bool h = false;
      float tHit = 0.f;
      if (SP.op[m].kind == 0) {
        const float3 gThis = vol_gradient(SP.op[m].vol, coord);
        const float dotValue = dot3(gThis, gLast[m]);
        if (dotValue < SP.op[m].param) { tHit = (tLast[m] + tThis) / 2.0f; h = true; }
        gLast[m] = gThis;
      } else {
        const float iso = SP.op[m].param, sThis = vol_sample(SP.op[m].vol, coord);
        if (((sLast[m] < iso) && (sThis >= iso)) || ((sLast[m] > iso) && sThis <= iso)) {
          tHit = tLast[m] + (((iso - sLast[m]) / (sThis - sLast[m])) * (tThis - tLast[m]));
          h = true;
        }
        sLast[m] = sThis;
      }
      tLast[m] = tThis;
      if (h) { tThis = tHit; hit = m; }
    }
    if (hit != -1 || tThis == tExit) break;
    tThis = tThis + step;
    if (tThis > tExit) tThis = tExit;
  }
  const float t_out = (hit != -1) ? tThis + 0.001f : tThis;  // :214-217
  if (hit != -1 && samples) {  // Sampler.cpp:74-86: position from the list's own (unpatched) direction
    const unsigned long long k = atomicAdd(sample_count, 1ull);
    if (k < sample_cap) {
      samples[3 * k] = org.x + t_out * dir0.x;
      samples[3 * k + 1] = org.y + t_out * dir0.y;
      samples[3 * k + 2] = org.z + t_out * dir0.z;
    }
  }
  // RAY_SURFACE only: KEEP_HERE (Renderer.cpp:304-421) -> the next pass starts behind the sample.  Only while t advances (at
  // |t| >= 2^14 the +0.001 is below one ulp): a ray that does not move goes back to the host loop like in the default mode.
  if (LOOP && hit != -1 && t_out > ray_t && my_passes < (1u << 20)) {
    ray_t = t_out;
    continue;
  }
  R.t[i] = t_out;
  R.term[i] = (hit != -1) ? RAY_SURFACE : RAY_BOUNDARY;
  break;
  }
  if (LOOP && passes) atomicAdd(passes, (unsigned long long)my_passes);
}

int launch_sampler_trace(const SamplerParams &SP, Rays R, int n, float *samples, unsigned long long *sample_count,
                         unsigned long long sample_cap, unsigned long long *passes, bool loop, cudaStream_t st) {
  if (n <= 0 || SP.n_ops < 1) return 0;  // SamplerTraceRays.ispc:136: nothing is touched without a sampler operator
  if (loop) sampler_trace_kernel<true><<<(n + 127) / 128, 128, 0, st>>>(SP, R, n, samples, sample_count, sample_cap, passes);
  else sampler_trace_kernel<false><<<(n + 127) / 128, 128, 0, st>>>(SP, R, n, samples, sample_count, sample_cap, passes);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gxy
