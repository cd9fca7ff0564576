/*
 * gxy_oracle.h -- C API of the CPU ORACLE for Galaxy's ray-rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / CPU baseline.  The product path (galaxy_b200/csrc, include/gxy_gpu.h)
 * never links, imports or calls this library.
 *
 * The oracle is a plain C++ restatement of the reference's algorithm for this path (file:line
 * citations are in gxy_oracle.cpp, relative to the reference tree).  Parity pin: the oracle's
 * renders of the reference's tests/ state files are compared against the reference's gold
 * PNGs (tests/golden/, see tests/test_oracle_golds.py) and against the reference's own
 * vendored Embree 3.6.1 compiled into oracle/_ref (tests/test_oracle_embree.py for triangles,
 * tests/test_oracle_curves.py for the round Bezier curves of PathLines) and its own Box.cpp
 * (tests/test_oracle_box.py).
 * PARITY UNPINNED for the Sampler part (gxo_sample*, src/sampler): the reference holds no golden
 * data for it and its kernels are ISPC, which cannot be compiled here; that part is checked
 * against closed-form properties only (tests/test_sampler.py).  PathLines: the curve builder
 * (gxo_build_curves) is pinned bit for bit by the reference's own DataDrivenPathLines::finalize
 * and the curve intersector by Embree's own header, both compiled into oracle/_ref.
 *
 * Floating point convention (shared with the CUDA path so both can be compared tightly):
 * IEEE fp32, round-to-nearest, true divides and square roots, NO contraction of a*b+c into
 * FMA except where the reference's Embree triangle test itself uses FMA (AVX2 madd/msub).
 */
#ifndef GXY_ORACLE_H
#define GXY_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define GXO_MAX_LIGHTS 16

typedef struct gxo_scene gxo_scene;

/* Lighting block of a Visualization (src/renderer/Lighting.cpp:59-115; Lighting.ih:23-32). */
typedef struct {
  int   n_lights;
  float lights[GXO_MAX_LIGHTS][3];
  int   types[GXO_MAX_LIGHTS];   /* 0 directional, 1 camera-relative, 2 point */
  int   n_ao;
  float ao_radius;
  int   shadows;
  float Ka, Kd;
} gxo_lighting;

/* Camera (src/renderer/Camera.h:199-204). */
typedef struct {
  float eye[3];
  float dir[3];
  float up[3];
  float aov;
} gxo_camera;

typedef struct {
  long long primary_rays;      /* originated (Camera.cpp:475)                       */
  long long shadow_rays;       /* spawned (TraceRays.cpp:119-122)                    */
  long long ao_rays;
  long long forwarded_rays;    /* rays sent to a neighbour partition                 */
  long long terminated_rays;   /* contributions added to the framebuffer             */
  long long traced_rays;       /* rays passed through TraceRays (incl. re-traces)    */
  long long volume_samples;    /* SampleVolumes calls x volumes                      */
  long long orphan_pixels;     /* pixels whose ray hits the global box but no rank claims */
  long long waves;
} gxo_stats;

gxo_scene *gxo_scene_create(void);
void gxo_scene_destroy(gxo_scene *);

/* global/local box + the six face neighbours of this partition (-1 = none)
 * (Visualization.cpp:139-160, Volume.cpp:358-390). */
void gxo_scene_set_partition(gxo_scene *, const float gmin[3], const float gmax[3],
                             const float lmin[3], const float lmax[3], const int neighbors[6]);

/* A VolumeVis operator on a (ghosted) brick.  voxels is shared, not copied (must outlive the
 * scene), type 0 = float32, 1 = uint8.  dataset_id identifies the underlying dataset so that
 * two operators on one dataset share the volume object's transfer function as the reference
 * does (MappedVis.cpp:206-212: last committed Vis wins; used by the DVR branch only).
 * colors = 256x3 floats, opacities = 256 floats (already resampled, MappedVis.cpp:277-338). */
int gxo_scene_add_volume_vis(gxo_scene *, int dataset_id, const int dims[3], const float origin[3],
                             const float spacing[3], int type, const void *voxels,
                             int n_slices, const float *slices4, int n_iso, const float *isovalues,
                             int volume_render, const float *colors, const float *opacities,
                             float range_lo, float range_hi);

/* A TrianglesVis operator: float3 vertices / normals, per-vertex data, int3 indices (shared). */
int gxo_scene_add_triangles_vis(gxo_scene *, int n_verts, const float *verts, const float *normals,
                                const float *data, int n_tris, const int *indices,
                                const float *colors, const float *opacities,
                                float range_lo, float range_hi);

/* A ParticlesVis operator: sphere centres + per-particle data (shared). */
int gxo_scene_add_particles_vis(gxo_scene *, int n, const float *centers, const float *data,
                                float radius0, float radius1, float value0, float value1,
                                const float *colors, const float *opacities,
                                float range_lo, float range_hi);

/* A PathLinesVis operator: poly-line vertices + per-vertex data, connectivity[i] = index of the first vertex of
 * segment i (PathLines.cpp:110-122).  Built into round cubic Bezier segments as DataDrivenPathLines::finalize does
 * (src/ospray/DataDrivenPathLines.cpp:103-156); the control points are owned by the scene.  <0: bad connectivity. */
int gxo_scene_add_pathlines_vis(gxo_scene *, int n_verts, const float *verts, const float *data, int n_segments,
                                const int *connectivity, float radius0, float radius1, float value0, float value1,
                                const float *colors, const float *opacities, float range_lo, float range_hi);
/* the control points alone: cp_out = n_segments x 4 x (x,y,z,r) */
int gxo_build_curves(int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity,
                     float radius0, float radius1, float value0, float value1, float *cp_out);
/* Embree's round-Bezier sweep intersector restated (curve_intersector_sweep.h:55-241), brute force over n_curves
 * segments; same signature as oracle/embree_curve_ref.cpp's gxr_curve_intersect, against which it is pinned. */
int gxo_curve_intersect(int n_curves, const float *cp, int n_rays, const float *org3, const float *dir3, const float *tnear,
                        const float *tfar, int *prim_out, float *tu_out, float *ng_out, int per_curve);

/* Builds the acceleration structure (the oracle's own simple BVH). */
int gxo_scene_commit(gxo_scene *);

/* Resample control points to a 256-entry table (MappedVis.cpp:277-338).  cmap = n x 4 (x,r,g,b),
 * omap = m x 2 (x,o).  colors_out 768 floats, opac_out 256 floats. */
void gxo_resample_tf(int n, const float *cmap, int m, const float *omap,
                     float *colors_out, float *opac_out);

/* Rendering::resolve_lights (Rendering.cpp:157-216). */
void gxo_resolve_lights(const gxo_lighting *in, const gxo_camera *cam, gxo_lighting *out);

/* RayList memory layout of the reference (Rays.cpp:42-204): 25 columns of aligned_n 4-byte
 * entries each, in the order ox oy oz dx dy dz nx ny nz sample r g b o sr sg sb so t tMax
 * (float) x y type term classification (int).  `base` points at column 0. */

/* TraceRays::Trace (TraceRays.cpp:68-146): traces the n rays in place and returns the number of
 * spawned secondary rays (kept inside the scene until fetched), or <0 on error.  `lights` must
 * already be resolved.  nearest-hit ids (geomID,primID or -1) are written to hit_ids (2*n ints)
 * if non-NULL. */
long long gxo_trace_raylist(gxo_scene *, const gxo_lighting *lights, float *base, int n, int aligned_n,
                            float epsilon, int *hit_ids);
/* copy the secondary list produced by the last gxo_trace_raylist into out_base */
int gxo_fetch_secondary(gxo_scene *, float *out_base, int aligned_n);

/* Renderer::Classify + AssignDestinations (Renderer.cpp:304-454) on a traced list. */
int gxo_classify(gxo_scene *, float *base, int n, int aligned_n);

/* Camera::generate_initial_rays/SpawnRays (Camera.cpp:379-493,528-829) for this partition.
 * Writes up to w*h rays into base (aligned_n >= w*h); returns the count. */
int gxo_generate_rays(gxo_scene *, const gxo_camera *, int w, int h, float *base, int aligned_n);

/* Whole frame over nparts partitions simulated in one process: generation, trace, secondary
 * rays, classify, forwarding between partitions, additive framebuffer (Renderer.cpp:179-269,
 * 504-656; Rendering.cpp:125-153).  fb = w*h*4 floats (y up), zeroed by the call.
 * lights are NOT yet resolved (resolve_lights is applied).  nthreads<=0: all cores. */
int gxo_render(int nparts, gxo_scene **parts, const gxo_camera *, const gxo_lighting *,
               int w, int h, float epsilon, int max_rays_per_packet, int nthreads,
               float *fb, gxo_stats *stats);

/* ---- Sampler (src/sampler): rays leave a sample point wherever a sampler operator fires ---------------------------
 * A sampling Visualization holds only sampler operators (SamplerTraceRays.ispc:128-222 calls every volumeVis through the
 * SamplerVis function table): kind 0 = GradientSamplerVis, param = "tolerance" (fires when dot(grad_this, grad_last) <
 * tolerance, sample at the interval's midpoint); kind 1 = IsoSamplerVis, param = "isovalue" (fires when the value crosses
 * it, sample at the linear crossing).  Volume arguments as gxo_scene_add_volume_vis. */
int gxo_scene_add_sampler_vis(gxo_scene *, int dataset_id, const int dims[3], const float origin[3], const float spacing[3],
                              int type, const void *voxels, int kind, float param);
/* SamplerTraceRays::Trace on one RayList: t and term are rewritten in place (term = RAY_SURFACE on a sample, else RAY_BOUNDARY) */
int gxo_sample_raylist(gxo_scene *, float *base, int n, int aligned_n);
/* Sampler over a frame of camera rays (Renderer::local_render with Sampler::Trace / Sampler::HandleTerminatedRays,
 * Sampler.cpp:52-133): every partition collects its samples; stats: primary_rays, traced_rays, forwarded_rays, waves. */
int gxo_sample(int nparts, gxo_scene **parts, const gxo_camera *, int w, int h, int max_rays_per_packet, int nthreads, gxo_stats *stats);
/* the samples of one partition after gxo_sample: returns their number, *xyz = 3 floats per sample (owned by the scene) */
long long gxo_scene_samples(gxo_scene *, const float **xyz);

/* The interactive / asynchronous frame path (Rendering::AddLocalPixels + ACCUMULATE_PIXEL without GXY_WRITE_IMAGES,
 * Rendering.cpp:104-153): the framebuffer (w*h*4 floats), the per-pixel frame stamps `kbuffer` (w*h ints, 0 after
 * Rendering::local_commit / local_reset :233-254) and the rendering's current frame (initially -1, :57) persist across calls.  Nothing is cleared: a
 * contribution of frame f resets its pixel first if the pixel's stamp is older; contributions of a frame older than the
 * rendering's current one are dropped. */
int gxo_render_progressive(int nparts, gxo_scene **parts, const gxo_camera *, const gxo_lighting *, int w, int h, float epsilon,
                           int nthreads, int frame, float *fb_inout, int *kbuffer_inout, int *rendering_frame_inout, gxo_stats *stats);

/* ColorImageWriter::Write (ImageWriter.cpp:30-48): float RGBA (y up) -> RGBA8 rows top-down,
 * truncating (unsigned char)(255*f) with x86 cvttss2si + low-byte semantics. */
void gxo_fb_to_rgba8(const float *fb, int w, int h, unsigned char *out);

/* Volume partitioning (Volume.cpp:88-172): factors for n ranks and the part table. */
void gxo_factor(int n, int factors[3]);
/* out: per part 15 ints: ijk[3] offsets[3] counts[3] goffsets[3] gcounts[3] */
void gxo_partition(int n, const int factors[3], const int grid[3], int *out);

/* Diagnostic switches for tests ("dvr_before_iso": see gxy_oracle.cpp). */
/* CPU arm "reference" of bench.py only: take the nearest hit of geometry operator 0's triangles from an external provider (the
 * reference's own Embree 3.6.1, oracle/embree_scene_ref.cpp) instead of the oracle's own tree.  fn gets n rays (org/dir 3 floats each,
 * interval (tnear, tfar]) and fills primID (-1: miss; geomID is ignored), t, u, v and the unnormalised geometric normal.  Set before
 * gxo_scene_commit (which then builds no tree).  Never used by the parity tests: they check the oracle's own arithmetic. */
typedef void (*gxo_intersect_fn)(void *user, int n, const float *org3, const float *dir3, const float *tnear, const float *tfar,
                                 int *geom_id, int *prim_id, float *t, float *u, float *v, float *ng3);
void gxo_scene_set_intersector(gxo_scene *, gxo_intersect_fn fn, void *user);
void gxo_set_option(const char *name, int value);
/* Box::exit_face / Box::intersect restatements on arrays (pinned against the reference's compiled Box.cpp) */
void gxo_exit_face(int n, const float *boxes6, const float *rays6, int *faces);
void gxo_box_intersect(int n, const float *boxes6, const float *rays6, int *hit, float *t2);

/* Nearest-hit query only (K2/K4): for n rays (org,dir,tnear,tfar) report geomID, primID, t, u, v.
 * Used for primID parity tests. */
int gxo_intersect(gxo_scene *, int n, const float *org3, const float *dir3, const float *tnear,
                  const float *tfar, int *geom_prim2, float *tuv3);

#ifdef __cplusplus
}
#endif
#endif
