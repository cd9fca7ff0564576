// gxy_progressive.cu -- the interactive / asynchronous frame path: frame-stamped accumulation.
//
// Reference: Rendering::AddLocalPixels + ACCUMULATE_PIXEL without GXY_WRITE_IMAGES (src/renderer/Rendering.cpp:104-153):
// the framebuffer is never cleared between frames; every pixel carries the stamp of the last frame that wrote it
// (`kbuffer`); the first contribution of a newer frame resets the pixel, later ones add; pixels of a frame older than the
// rendering's current one are dropped.  gxyviewer shows the image while it fills in.
//
// Here a frame is produced by the bulk-synchronous frame path into a cleared buffer (gxy_render), so the stamp logic
// collapses into ONE pass over the image: a pixel that received at least one contribution of frame f -- exactly the
// pixels for which some partition originated a primary ray, since every primary ray ends as a TERMINATED contribution
// somewhere -- takes the new sum if its stamp is older (reset + adds) or adds it (same frame again); all other pixels
// keep what they show.  mark_touched_kernel derives that set from the generated ray lists.
#include "gxy_internal.h"

namespace gxy {

__global__ void __launch_bounds__(256) mark_touched_kernel(Rays R, int n, int w, unsigned char *__restrict__ touched) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) touched[(size_t)R.y[i] * w + R.x[i]] = 1;
}

__global__ void __launch_bounds__(256) merge_stamped_kernel(const float4 *__restrict__ frame_sum, float4 *__restrict__ image,
                                                            int *__restrict__ kbuffer, const unsigned char *__restrict__ touched,
                                                            int npix, int frame) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix || !touched[p]) return;
  const float4 s = frame_sum[p];
  if (kbuffer[p] < frame) {  // ACCUMULATE_PIXEL: reset, stamp, then the frame's contributions
    image[p] = s;
    kbuffer[p] = frame;
  } else {                   // the same frame again: contributions add on top
    float4 o = image[p];
    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    image[p] = o;
  }
}

int launch_mark_touched(Rays R, int n, int w, unsigned char *touched, cudaStream_t st) {
  if (n <= 0) return 0;
  mark_touched_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, n, w, touched);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_merge_stamped(const float *frame_sum, float *image, int *kbuffer, const unsigned char *touched, int npix, int frame,
                         cudaStream_t st) {
  merge_stamped_kernel<<<(npix + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4 *>(frame_sum), reinterpret_cast<float4 *>(image),
                                                           kbuffer, touched, npix, frame);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gxy
