// gxy_internal.h -- host-side declarations shared by the translation units of libgxy_b200.so
#pragma once
#include "gxy_common.cuh"

namespace gxy {

// ---- kernels (gxy_kernels.cu) ---------------------------------------------------------------
// K1 TraceRays_TraceRays on n rays (TraceRays.ispc:326-623).  hit_ids may be NULL.
// anyhit_secondary: occlusion-only traversal for non-PRIMARY rays (valid when no DVR integrates).
// n_dev (here and below; may be NULL): the list's length on the device -- the launch is sized for n (the list's capacity) and every
// kernel takes min(n, *n_dev); the waves of a frame in flight never bring a count to the host
int launch_trace(const SceneParams &P, Rays R, int n, float global_epsilon, int *hit_ids, bool anyhit_secondary,
                 unsigned long long *sample_counter, cudaStream_t st, const int *n_dev = nullptr);
// offsets for the secondary list (TraceRays.cpp:89-117): per ray "1 if PRIMARY&&SURFACE" flags are
// scanned; returns device pointers inside `scratch`.  n_hit is written to *d_nhit (device).
int launch_hit_scan(Rays R, int n, int *d_hit_index /*n*/, int *d_block_sums, int *d_nhit, cudaStream_t st, const int *n_dev = nullptr);
// ambientLighting + generateAORays + diffuseLighting + generateShadowRays (TraceRays.ispc:625-923)
int launch_shade_spawn(const DevLights &L, Rays R, int n, const int *d_hit_index, const int *d_nhit, Rays out, float epsilon,
                       cudaStream_t st, int max_hits = 0);
// Renderer::Classify + AssignDestinations (Renderer.cpp:304-454)
int launch_classify(const SceneParams &P, Rays R, int n, cudaStream_t st, const int *n_dev = nullptr);
// HandleTerminatedRays + AddLocalPixels (Renderer.cpp:456-502, Rendering.cpp:125-153): fb += rgba
int launch_accumulate(Rays R, int n, float *fb, int w, int h, unsigned long long *d_terminated, cudaStream_t st, const int *n_dev = nullptr);
// counting-sort of rays with classification >= 0 into per-destination segments of `out`
// (Renderer.cpp:561-618).  d_counts: nranks ints (device, zeroed by the call), d_offsets nranks+1.
int launch_partition_by_destination(Rays R, int n, int nranks, int keep_rank, Rays out, int *d_counts, int *d_offsets, int *d_cursor,
                                    cudaStream_t st);
// Camera::SpawnRays (Camera.cpp:379-493): ordered compaction of the kept pixels; tiled = false: the reference's pixel order,
// true: 16x8-pixel CTA tiles of 8x4 warp tiles (frame path: compact beams for the march kernel)
int launch_generate(const SceneParams &P, const DevCamera &C, int w, int h, bool tiled, Rays out, int *d_flags_scan, int *d_block_sums,
                    int *d_count, cudaStream_t st);
// ColorImageWriter::Write (ImageWriter.cpp:30-48)
int launch_tonemap(const float *fb, int w, int h, unsigned char *rgba, cudaStream_t st);
int launch_fb_add(float *dst, const float *src, size_t n, cudaStream_t st);
// copy n rays (25 columns) between lists: dst[dst_off + i] = src[src_off + i]
int launch_copy_rays(Rays dst, size_t dst_off, Rays src, size_t src_off, int n, cudaStream_t st);
int launch_intersect(const SceneParams &P, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar,
                     int *geom_prim2, float *tuv3, cudaStream_t st);

// ---- Sampler (gxy_sampler.cu) -----------------------------------------------------------------------
// operators of a sampling Visualization (src/sampler): kind 0 GradientSamplerVis (param = tolerance), 1 IsoSamplerVis (isovalue)
struct SamplerOpDev {
  int kind;
  float param;
  DevVolume vol;
};
struct SamplerParams {
  int n_ops;
  float step;  // min samplingStep*samplingRate over the operators (SamplerTraceRays.ispc:141-150)
  float3 lmin, lmax;
  SamplerOpDev op[GXY_MAX_VOLUME_VIS];
};
// SamplerTraceRays::Trace on n rays: t and term rewritten in place; samples != NULL: the hit points (xyz) are appended at
// *sample_count (device counter, incremented for every hit; points beyond sample_cap are counted but not written)
// loop (opt-in, GXY_SAMPLER_LOOP=1): a ray that left a sample does its next pass inside the kernel until it reaches the boundary;
// *passes (device counter) += the passes made, the reference's traced-ray count.  Needs samples != NULL.
int launch_sampler_trace(const SamplerParams &SP, Rays R, int n, float *samples, unsigned long long *sample_count,
                         unsigned long long sample_cap, unsigned long long *passes, bool loop, cudaStream_t st);

// ---- frame-stamped accumulation of the interactive frame path (gxy_progressive.cu; Rendering.cpp:104-153) ----------
// touched[y*w+x] = 1 for every ray of the list
int launch_mark_touched(Rays R, int n, int w, unsigned char *touched, cudaStream_t st);
// per touched pixel: stamp older than `frame` -> image = frame_sum, stamp = frame; else image += frame_sum
int launch_merge_stamped(const float *frame_sum, float *image, int *kbuffer, const unsigned char *touched, int npix, int frame,
                         cudaStream_t st);

// ---- TMA-staged volume march (gxy_march_tma.cu) ---------------------------------------------------------
// one float volume operator, no geometry, < 2^31 voxels, x dimension a multiple of 4 (TMA strides are multiples of 16 bytes)
bool march_tma_eligible(const SceneParams &P);
// same contract as launch_trace for such a Visualization; axis = volume axis the beams mostly run along (0 x, 1 y, 2 z);
// staged_counter (may be NULL) receives the number of samples served from the staged boxes
int launch_march_tma(const SceneParams &P, Rays R, int n, float global_epsilon, int axis, unsigned long long *sample_counter,
                     unsigned long long *staged_counter, cudaStream_t st);

// ---- fused frame kernels for geometry-only Visualizations (gxy_fused.cu) ---------------------------
struct FusedQueues {  // device memory, zeroed at the start of a frame
  unsigned pixel_head, n_hits, n_spill, sec_head;
  unsigned n_primary32, hits_done, inbox_head, spill_done;
  unsigned long long n_terminated, nodes, prims, n_inbox;
  unsigned global_pending, n_generated;  // n_generated: kept primaries (n_primary32 = those queued for the trace kernel)
  unsigned sec_lo, sec_hi;               // peer path: hit records [sec_lo, sec_hi) get their AO/shadow rays in the current wave
  unsigned n_virtual, pad[3];            // peer path: forwards decided at generation from the partition proxies (no record sent)
};

// ---- peer arenas: the one-process-per-GPU exchange without NCCL on the data path -------------------
// Every rank owns one device allocation ("arena") that all other ranks map with CUDA IPC.  A ray that
// leaves a partition is written by the trace kernel itself, as one 64-byte record, straight into the
// destination rank's inbox over NVLink (slot from a system-scope atomicAdd on the destination's
// counter); waves are separated by a device-side flag barrier; the partial framebuffers are summed tile
// by tile with peer loads.  Replaces SendRaysMsg / SendPixelsMsg over MPI (Renderer.cpp:732-836,
// Renderer.h:246-340) and the busy/idle termination protocol (RenderingSet.cpp:289-589).
#define GXY_MAX_RANKS 16
struct PeerCtrl {  // first 256 bytes of every arena
  unsigned flags[GXY_MAX_RANKS];       // flags[r]: last epoch rank r has signalled to this rank
  unsigned pending[2][GXY_MAX_RANKS];  // pending[epoch & 1][r]: rays rank r sent anywhere in that epoch's wave
  unsigned inbox_count[2];             // records appended to this rank's inbox[parity]
  unsigned pad[14];
};
static_assert(sizeof(PeerCtrl) == 256, "PeerCtrl layout");
// What a rank publishes about its partition so that every other rank can tell, without communication, that a primary
// ray cannot hit anything there: the partition box, its neighbour table and the top two levels of its BVH (root, patched
// to child_base = 1, followed by its internal children in slot order).  See gen_primary_peer_kernel (gxy_fused.cu).
struct PartProxy {
  float3 lmin, lmax;
  int neighbors[6];
  int has_prims;
  int pad[3];
  WideNode nodes[9];
};
static_assert(sizeof(PartProxy) == 64 + 9 * 80, "PartProxy layout");
struct PeerTable {  // kernel parameter: every rank's arena as mapped in THIS process (base[rank] = own)
  int rank, nranks;
  unsigned inbox_cap;  // 64-byte records per inbox
  unsigned npix;
  unsigned long long off_fb, off_final, off_inbox[2];  // byte offsets inside an arena, the same on every rank
  unsigned long long off_proxy;                        // this rank's PartProxy
  char *base[GXY_MAX_RANKS];
};
// inbox record: 4 x float4 = (ox oy oz dx) (dy dz t tMax) (r g b o) (x y type term)

// generation (1 kernel) -> trace + classify of the misses (persistent kernel) -> shade/light/framebuffer of
// the hits (1 kernel); hit records go to `hits` (columns ox..dz t nx..nz sr..sb o x y), rays bound for
// a neighbour to `spill` (peer == NULL) or to the neighbour's inbox[0] (peer != NULL)
// prim: list for the generated rays (>= w*h), raw: 6*w*h words of scratch, hits: >= w*h records
// peer != NULL: proxies = the PartProxy of every rank in local memory (launch_proxy_gather); rays are then originated by the
// first partition on their path whose proxy they may hit instead of being forwarded through the ones they cannot
// band / n_bands: only the tile rows band, band + n_bands, ... of the window (gxy_render pipelines bands on two streams);
// raw_stride: words per plane of `raw` (>= the number of rays this launch can queue)
int launch_fused_primary(const SceneParams &P, const DevCamera &C, const DevLights &L, int w, int h, float *fb, Rays prim, unsigned *raw,
                         unsigned raw_stride, Rays hits, Rays spill, unsigned spill_cap, FusedQueues *q, float epsilon, const PeerTable *peer,
                         const PartProxy *proxies, int band, int n_bands, cudaStream_t st, const int *tile_rect = nullptr);
// write this rank's PartProxy into its arena; after the next flag barrier ...
int launch_proxy_publish(const SceneParams &P, const PeerTable &T, cudaStream_t st);
// ... copy every rank's published PartProxy (peer loads) into the local table `out` (nranks entries)
int launch_proxy_gather(const PeerTable &T, PartProxy *out, cudaStream_t st);
// AO + shadow rays of the hit records [q->hits_done, q->n_hits): generate -> trace -> classify -> framebuffer
int launch_fused_secondary(const SceneParams &P, const DevLights &L, int w, int h, int nsec, long long max_rays, float *fb, Rays hits,
                           Rays spill, unsigned spill_cap, FusedQueues *q, float epsilon, bool anyhit, const PeerTable *peer, int parity_out,
                           cudaStream_t st, int blocks_per_sm = 0);
// one wave >= 1 of the peer path: trace the records of inbox[parity_in]; PRIMARY hits -> raw records -> shading
// (hit records appended to `hits`); everything else is classified: framebuffer, dropped, or the next inbox
int launch_inbox_wave(const SceneParams &P, const DevLights &L, const PeerTable &T, int parity_in, int w, int h, float *fb, unsigned *raw,
                      unsigned raw_stride, Rays hits, FusedQueues *q, float epsilon, bool anyhit, cudaStream_t st, int blocks_per_sm = 0);
// end of a wave: queue bookkeeping, reset of the consumed inbox, flag barrier across ranks, global pending count
int launch_wave_epilogue(const PeerTable &T, FusedQueues *q, unsigned epoch, int parity_consumed, bool hits_spawn, int *error_flag,
                         cudaStream_t st);
// sum of all partial framebuffers for this rank's slice of the image, written to the image owner (rank 0)
int launch_fb_gather(const PeerTable &T, cudaStream_t st);

// ---- frames in flight for Visualizations with volumes (one process per GPU): list kernels with device-side lengths + peer inboxes ----
struct VolQueues {  // device memory, zeroed at the start of a frame
  int cnt[2];                   // rays in the two device lists of the frame (the current one and the one being assembled)
  int nhit;                     // PRIMARY && SURFACE rays of the current list (launch_hit_scan)
  unsigned away_wave, kept_wave;  // this wave: records pushed to other ranks / rays appended to the own next list
  unsigned global_pending;      // after a wave's barrier: work left on all ranks
  unsigned long long generated, traced, ao, shadow, forwarded, terminated, samples, inbox;
};
// after the hit scan of list `cur`: statistics, and the next list opens with the nhit * (n_ao + n_sh) rays about to be spawned
int launch_vol_wave_counts(VolQueues *q, int cur, int n_ao, int n_sh, bool first, cudaStream_t st);
// classified list `cur` (capacity cap): rays for other ranks -> records in their inbox[parity]; rays that stay -> appended to `next`
int launch_vol_forward(Rays R, int cap, Rays next, int next_cap, VolQueues *q, int cur, const PeerTable &T, int parity, int *error_flag,
                       cudaStream_t st);
// flag barrier of the wave; q->global_pending = work left anywhere
int launch_vol_epilogue(const PeerTable &T, VolQueues *q, int cur, unsigned epoch, int *error_flag, cudaStream_t st);
// after the barrier: the records of the own inbox[parity] are appended to `next`, its length is final, the inbox is handed back
int launch_vol_unpack(const PeerTable &T, int parity, Rays next, int next_cap, VolQueues *q, int cur, int *error_flag, cudaStream_t st);

// GXY_PROFILE=1: device timeline of a frame -- a named CUDA event after a launch; printed by the frame function
void gxy_timeline_mark(const char *name, cudaStream_t st);

// ---- BVH build (gxy_bvh.cu) -----------------------------------------------------------------
struct GeomBuildInput {
  int kind;  // 0 triangles, 1 spheres, 2 round Bezier curves (centers = 16 floats per segment: 4 control points x,y,z,r)
  int geom_id;
  long long n_prims;
  const float *verts;  // device
  const int *idx;      // device
  const float *centers;
  const float *data;
  float radius0, radius1, value0, value1, epsilon;
};
struct BvhResult {
  WideNode *nodes = nullptr;
  PrimRec *prims = nullptr;
  long long n_nodes = 0, n_prims = 0;
  int max_depth = 0;
  float build_ms = 0.f;       // CUDA events around the whole build (kernels + the waits for the host between them)
  float alloc_host_ms = 0.f;  // host time spent inside cudaMalloc / cudaFree of the build's buffers within that span
};
int build_bvh(const GeomBuildInput *geoms, int n_geoms, BvhResult *out, cudaStream_t st);

}  // namespace gxy
