/*
 * gxy_gpu.h -- C ABI of the B200-native (sm_100a) implementation of Galaxy's ray-rendering
 * hot path: "trace a RayList against a Visualization" (BVH traversal + ray/triangle +
 * ray/sphere, volume march with isosurface/slice detection, lighting + secondary rays,
 * classify/forward between spatial partitions, additive framebuffer).
 *
 * This is the drop-in boundary.  In the reference the C++ host reaches native code for this path
 * only through the ISPC `export` functions (extern "C", void* handles) and the OSPRay C API;
 * every entry point below names the reference interface it replaces (file:line relative to the
 * TACC/Galaxy tree).  Plain pointers and sizes only; no torch / CUDA types in the signatures.
 * INTEGRATION.md shows the reference-side bindings a Galaxy maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; gxy_last_error() gives the text.
 *    Nothing ever calls exit() (the reference prints and exit(1)s, e.g. Renderer.cpp:377).
 *  - host buffers passed in are copied to the device at create/commit time (the reference shares
 *    them, OSP_DATA_SHARED_BUFFER, OsprayVolume.cpp:36-37); the device owns its copy.
 *  - there is NO CPU fallback: if no CUDA device is usable every compute entry fails loudly.
 *  - kernels report through a device error flag that the entry point reads before it returns ("device error flag N" in
 *    gxy_last_error(); the flag is cleared once reported): 1 BVH traversal stack overflow, 3 ray list / inbox capacity
 *    exceeded, 4 peer barrier timeout, 6 TMA copy did not complete, 7 translucent surface (opacity <= 0.999) on the
 *    fused frame path, which has no keeper list (no shader produces one today).
 */
#ifndef GXY_GPU_H
#define GXY_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GXY_MAX_LIGHTS     16
#define GXY_MAX_VOLUME_VIS 8    /* reference: 100 (TraceRays.ispc:455) */
#define GXY_MAX_SLICES     16
#define GXY_MAX_ISOVALUES  16
#define GXY_RAYLIST_COLUMNS 25

typedef struct gxy_context   gxy_context;    /* one CUDA device + stream + memory pools          */
typedef struct gxy_volume    gxy_volume;     /* src/data/Volume + src/ospray/OsprayVolume         */
typedef struct gxy_triangles gxy_triangles;  /* src/data/Triangles + OsprayTriangles              */
typedef struct gxy_particles gxy_particles;  /* src/data/Particles + OsprayParticles              */
typedef struct gxy_pathlines gxy_pathlines;  /* src/data/PathLines + OsprayPathLines              */
typedef struct gxy_vis       gxy_vis;        /* src/renderer/Visualization (one partition)        */
typedef struct gxy_raylist   gxy_raylist;    /* src/renderer/Rays.h RayList, library-owned        */
typedef struct gxy_dev_raylist gxy_dev_raylist; /* ... a RayList that stays on the device between calls */

/* Lighting_ispc (src/renderer/Lighting.ih:23-32; Lighting_Set{Lights,K,AO,ShadowFlag},
 * Lighting.ispc:54-134). */
typedef struct {
  int   n_lights;
  float lights[GXY_MAX_LIGHTS][3];
  int   types[GXY_MAX_LIGHTS];      /* 0 directional, 1 camera-relative, 2 point (Lighting.cpp:59-115) */
  int   n_ao;
  float ao_radius;
  int   shadows;
  float Ka, Kd;
} gxy_lighting;

/* Camera (src/renderer/Camera.h:199-204). */
typedef struct {
  float eye[3];
  float dir[3];
  float up[3];
  float aov;                         /* degrees; 0 => orthographic (Camera.cpp:548-556) */
} gxy_camera;

/* transfer function as handed to OSPRay by MappedVis::local_commit (MappedVis.cpp:277-338):
 * 256 colours, 256 opacities, valueRange. */
typedef struct {
  float colors[256][3];
  float opacities[256];
  float range_lo, range_hi;
} gxy_transfer_function;

/* RayList_ispc (src/renderer/Rays.ih:20-47; layout Rays.cpp:42-204): 25 columns of aligned_n
 * 4-byte entries, order ox oy oz dx dy dz nx ny nz sample r g b o sr sg sb so t tMax (float),
 * x y type term classification (int).  base points at column 0 (host memory). */
typedef struct {
  float *base;
  int    n;
  int    aligned_n;
} gxy_raylist_view;

typedef struct {
  long long primary_rays;     /* originated (Camera.cpp:475)                                  */
  long long shadow_rays;      /* spawned (TraceRays.cpp:98-122)                               */
  long long ao_rays;
  long long forwarded_rays;   /* sent to a neighbour partition (Renderer.cpp:620-634)         */
  long long terminated_rays;  /* contributions added to the framebuffer (Rendering.cpp:125-153) */
  long long traced_rays;      /* rays passed through the trace kernel, incl. re-traces        */
  long long waves;            /* trace launches                                               */
  long long kernel_launches;  /* kernels of this library launched for the frame               */
  float     device_ms;        /* CUDA-event time, generation -> last framebuffer add          */
  float     trace_ms;         /* sum of the trace-kernel launch durations (CUDA events)       */
  long long nodes_visited;    /* BVH nodes / primitives tested by the trace kernel; 0 unless  */
  long long prims_tested;     /* the library was compiled with -DGXY_TRAV_COUNTERS            */
  long long volume_samples;   /* trilinear volume samples taken by the march (SampleVolumes x volumes) */
  long long staged_samples;   /* ... of which served from TMA-staged shared-memory boxes (GXY_MARCH_TMA=1)      */
  long long dequeued_rays;    /* rays the trace launches actually took from a queue or list: traced_rays minus the
                                 primaries the generation kernel finished itself (they can reach no primitive)    */
  float     t_begin_ms;       /* frames in flight: start and end of this frame on the device, relative to the    */
  float     t_end_ms;         /* last gxy_context_mark (or the creation of the context)                          */
} gxy_stats;

/* ---- library ---------------------------------------------------------------------------- */
const char *gxy_last_error(void);
const char *gxy_version(void);
/* number of usable CUDA devices (0 => every compute entry point will fail) */
int gxy_device_count(void);

/* ---- context ---------------------------------------------------------------------------- */
int  gxy_context_create(int device, gxy_context **out);
void gxy_context_destroy(gxy_context *);
int  gxy_context_synchronize(gxy_context *);
/* waits for the device to go idle and makes "now" the origin of gxy_stats::t_begin_ms / t_end_ms */
int  gxy_context_mark(gxy_context *);

/* ---- datasets (partition-local data, copied H2D once) -------------------------------------- */
/* replaces ospNewVolume("shared_structured_volume") + ospSet* in OsprayVolume::OsprayVolume
 * (src/ospray/OsprayVolume.cpp:25-52).  dims/origin are those of the GHOSTED brick
 * (Volume.h:124-129), voxels x-fastest, type 0 = float32, 1 = uint8. */
int  gxy_volume_create(gxy_context *, const int dims[3], const float origin[3], const float spacing[3],
                       int type, const void *voxels, gxy_volume **out);
void gxy_volume_destroy(gxy_volume *);

/* replaces ospNewGeometry("ddtriangles") + ospSetData (src/ospray/OsprayTriangles.cpp:25-57):
 * float3 vertices / normals (normals may be NULL), per-vertex data (may be NULL), int3 indices. */
int  gxy_triangles_create(gxy_context *, int n_verts, const float *verts, const float *normals,
                          const float *data, int n_tris, const int *indices, gxy_triangles **out);
void gxy_triangles_destroy(gxy_triangles *);

/* replaces ospNewGeometry("ddspheres") (src/ospray/OsprayParticles.cpp:34-40): float3 centres,
 * per-particle data (may be NULL). */
int  gxy_particles_create(gxy_context *, int n, const float *centers, const float *data, gxy_particles **out);
void gxy_particles_destroy(gxy_particles *);

/* replaces ospNewGeometry("ddpathlines") (src/ospray/OsprayPathLines.cpp:25-71): poly-line vertices (float3),
 * per-vertex data (may be NULL = 0) and the connectivity of src/data/PathLines.cpp:110-122: connectivity[i] = index
 * of the first vertex of segment i, its second vertex is the next one.  The arrays are copied (host side): the
 * curves are only built when a PathLinesVis maps data to radii (gxy_vis_commit). */
int  gxy_pathlines_create(gxy_context *, int n_verts, const float *verts, const float *data, int n_segments,
                          const int *connectivity, gxy_pathlines **out);
void gxy_pathlines_destroy(gxy_pathlines *);
/* DataDrivenPathLines::finalize (src/ospray/DataDrivenPathLines.cpp:28-37,103-156) on the host, as in the
 * reference: one round cubic Bezier segment per poly-line segment, cp_out = n_segments x 4 control points
 * (x,y,z,radius) as Embree gathers them (4 consecutive vertices from indexCurve[i]).  Needs no device. */
int  gxy_build_curves(int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity,
                      float radius0, float radius1, float value0, float value1, float *cp_out);

/* ---- Visualization ------------------------------------------------------------------------ */
int  gxy_vis_create(gxy_context *, gxy_vis **out);
void gxy_vis_destroy(gxy_vis *);
/* Visualization_commit boxes (src/renderer/Visualization.ispc:48-93) + neighbours
 * (Visualization.cpp:139-160; Volume.cpp:358-377; -1 = no neighbour on that face). */
int  gxy_vis_set_partition(gxy_vis *, const float gmin[3], const float gmax[3], const float lmin[3],
                           const float lmax[3], const int neighbors[6]);
/* VolumeVis operator: VolumeVis_SetSlices/SetIsovalues/SetVolumeRenderFlag (VolumeVis.ispc:57-91)
 * + MappedVis_set_transferFunction.  Operators on the same gxy_volume share the volume object's
 * transfer function for DVR (last added wins, MappedVis.cpp:206-212). */
int  gxy_vis_add_volume(gxy_vis *, gxy_volume *, int n_slices, const float *slices4, int n_isovalues,
                        const float *isovalues, int volume_render, const gxy_transfer_function *);
/* TrianglesVis operator (geomID = order among geometry operators, Visualization.cpp:270-273) */
int  gxy_vis_add_triangles(gxy_vis *, gxy_triangles *, const gxy_transfer_function *);
/* ParticlesVis operator; radius0/1 value0/1 per ParticlesVis.cpp:136-144 */
int  gxy_vis_add_particles(gxy_vis *, gxy_particles *, float radius0, float radius1, float value0,
                           float value1, const gxy_transfer_function *);
/* PathLinesVis operator; radius0/1 value0/1 per PathLinesVis.cpp:105-125,133-144 (the data value of a vertex is
 * mapped to the tube radius there; the hit colour is the transfer function of the radius mapped back).  The
 * segments are traced as Embree's round Bezier curves (RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE,
 * DataDrivenPathLines.ispc:319-324). */
int  gxy_vis_add_pathlines(gxy_vis *, gxy_pathlines *, float radius0, float radius1, float value0,
                           float value1, const gxy_transfer_function *);
/* Visualization::SetOsprayObjects -> ospCommit(model) (Visualization.cpp:207-285): builds the
 * BVH over all geometry operators on the device. */
int  gxy_vis_commit(gxy_vis *);
/* build statistics of the last commit: primitives, wide nodes, build milliseconds */
int  gxy_vis_build_info(gxy_vis *, long long *n_prims, long long *n_nodes, float *build_ms);
/* build_ms again (CUDA events around the whole build) and the part of that span the host spent inside cudaMalloc / cudaFree of the
 * build's buffers: the build allocates and releases ~25 buffers of up to 4.8 GB one by one, and the driver's time for that varies
 * from run to run (tens to hundreds of milliseconds) while the kernels' does not */
int  gxy_vis_build_times(gxy_vis *, float *build_ms, float *alloc_host_ms);

/* ---- host-side helpers that the reference keeps in C++ ------------------------------------ */
/* Rendering::resolve_lights (src/renderer/Rendering.cpp:157-216) */
int  gxy_resolve_lights(const gxy_lighting *in, const gxy_camera *, gxy_lighting *out);
/* MappedVis::local_commit resampling (src/renderer/MappedVis.cpp:277-338); cmap n x (x,r,g,b),
 * omap m x (x,o); range is NOT touched. */
int  gxy_resample_transfer_function(int n, const float *cmap4, int m, const float *omap2,
                                    gxy_transfer_function *out);
/* Volume partitioning (src/data/Volume.cpp:88-172) */
void gxy_factor(int n, int factors[3]);
/* out: per part 15 ints ijk[3] offsets[3] counts[3] goffsets[3] gcounts[3], rank order */
void gxy_partition(int n, const int factors[3], const int grid[3], int *out);

/* ---- per-RayList entry points (the reference's hot call) ------------------------------------ */
/* Threading: the reference calls TraceRays::Trace from GXY_NTHREADS pool threads (default 5) at once, every thread with its
 * own TraceRays object, all sharing one Visualization and Lighting that are read-only while rendering
 * (src/framework/Application.cpp:75, src/renderer/Renderer.cpp:504-556).  The entry points of this section may be called concurrently
 * from any number of host threads on ONE gxy_vis: each call takes one of the Visualization's 8 list lanes (own CUDA stream, own device
 * lists, own error flag) and concurrent calls overlap on the device; a ninth caller waits for a free lane.  gxy_vis_commit and the
 * gxy_vis_add_* calls must not run at the same time (commits happen while rendering is quiescent, as in the reference).
 * gxy_last_error() is per thread. */
/* TraceRays::Trace (src/renderer/TraceRays.cpp:68-146), i.e. ispc::TraceRays_TraceRays +
 * _ambientLighting + _generateAORays + _diffuseLighting + _generateShadowRays
 * (TraceRays.ispc:326,625,735,763,859).  rays (host) are traced in place; *out receives the
 * SECONDARY list (NULL if no ray was spawned), library-owned.  lights must be resolved.
 * hit_ids (may be NULL): 2 ints per ray, nearest-hit (geomID, primID) or (-1,-1).
 * Of the secondary list the 16 columns generateAORays / the shadow-ray loop set are defined (ox..dz r g b o t tMax x y type term);
 * the reference leaves the other 9 uninitialised, here they are 0, so the list is a function of the call's inputs alone. */
int  gxy_trace_raylist(gxy_vis *, const gxy_lighting *lights, gxy_raylist_view rays, float epsilon,
                       gxy_raylist **out, int *hit_ids);
int  gxy_raylist_get_view(gxy_raylist *, gxy_raylist_view *view);
void gxy_raylist_free(gxy_raylist *);
/* Renderer::Classify + AssignDestinations (src/renderer/Renderer.cpp:304-454) */
int  gxy_classify(gxy_vis *, gxy_raylist_view rays);
/* Device-resident RayLists (SURVEY 8b "Ownership": the reference's RayList is one refcounted smem block that travels between
 * Renderer::Trace, Classify and the message layer, Renderer.h:223; here the list may stay in HBM between those calls).
 * upload copies the 25 columns H2D once; trace / classify then work on the device copy -- the secondary list of a trace is itself a
 * device list, so the waves of a frame driven list by list (Renderer::ProcessRays) never cross PCIe; download copies all columns
 * back (host list with aligned_n >= gxy_raylist_size). */
int  gxy_raylist_upload(gxy_vis *, gxy_raylist_view rays, gxy_dev_raylist **out);
int  gxy_raylist_size(gxy_dev_raylist *, int *n);
int  gxy_raylist_download(gxy_dev_raylist *, gxy_raylist_view rays);
void gxy_dev_raylist_free(gxy_dev_raylist *);
/* TraceRays::Trace on a device list: rays traced in place; *secondary (may be NULL) receives the spawned AO/shadow list or NULL */
int  gxy_trace_raylist_dev(gxy_vis *, const gxy_lighting *lights, gxy_dev_raylist *rays, float epsilon, gxy_dev_raylist **secondary);
int  gxy_classify_dev(gxy_vis *, gxy_dev_raylist *rays);
/* Camera::generate_initial_rays + SpawnRays (src/renderer/Camera.cpp:379-493,528-829) for this
 * partition; rays.aligned_n >= w*h; *n_out = number of rays kept, in pixel order. */
int  gxy_generate_rays(gxy_vis *, const gxy_camera *, int w, int h, gxy_raylist_view rays, int *n_out);
/* nearest-hit only (rtcIntersect through Model.ih:54-70): org/dir 3 floats per ray */
int  gxy_intersect(gxy_vis *, int n, const float *org3, const float *dir3, const float *tnear,
                   const float *tfar, int *geom_prim2, float *tuv3);

/* ---- interactive / asynchronous frame path ------------------------------------------------ */
/* Rendering::AddLocalPixels + ACCUMULATE_PIXEL as built without GXY_WRITE_IMAGES (src/renderer/Rendering.cpp:104-153),
 * what gxyviewer displays: the image and a per-pixel frame stamp live on parts[0] across calls and are never cleared.
 * Frame `frame` is rendered as gxy_render renders it; then every pixel that received a contribution takes the new value
 * if its stamp is older (reset + adds) or adds it (the same frame number again), every other pixel keeps what it shows.
 * A frame older than the newest one seen is dropped (AddLocalPixels :138).  A change of w x h re-allocates (local_commit).
 * One process; parts as for gxy_render. */
int  gxy_render_progressive(int nparts, gxy_vis *const *parts, const gxy_camera *, const gxy_lighting *, int w, int h,
                            float epsilon, int frame, gxy_stats *stats);
/* The same with frames in flight: frame `frame` goes onto frame slot `slot` (see gxy_render_submit) and is merged into the displayed
 * image when it is waited for; *merged = 0 if a newer frame was merged while this one was in flight (it is dropped, as AddLocalPixels
 * drops the pixels of a superseded frame, :138-152).  Slots may be waited for in any order. */
int  gxy_render_progressive_submit(int nparts, gxy_vis *const *parts, const gxy_camera *, const gxy_lighting *, int w, int h,
                                   float epsilon, int frame, int slot);
int  gxy_render_progressive_wait(int nparts, gxy_vis *const *parts, int slot, gxy_stats *stats, int *merged);
/* the displayed image: float RGBA, y up (w*h*4 floats) / RGBA8 rows top-down as ColorImageWriter writes them */
int  gxy_progressive_download_rgba32f(gxy_vis *, float *fb);
int  gxy_progressive_download_rgba8(gxy_vis *, unsigned char *rgba);
/* Rendering::local_reset (:240-256): image and stamps to zero, current frame to -1 */
int  gxy_progressive_reset(gxy_vis *);

/* ---- Sampler (src/sampler) ---------------------------------------------------------------- */
/* A sampling Visualization holds only sampler operators (SamplerTraceRays.ispc:128-222 calls every volumeVis through the
 * SamplerVis function table).  kind 0 = GradientSamplerVis, param = "tolerance" (GradientSamplerVis.cpp:77-85): a sample
 * where dot(gradient here, gradient one step back) < tolerance; kind 1 = IsoSamplerVis, param = "isovalue"
 * (IsoSamplerVis.cpp:77-85): a sample where the value crosses it.  Replaces GradientSamplerVis_* / IsoSamplerVis_* (ispc). */
#define GXY_SAMPLER_GRADIENT 0
#define GXY_SAMPLER_ISO      1
int  gxy_vis_add_sampler(gxy_vis *, gxy_volume *, int kind, float param);
/* replaces ispc::SamplerTraceRays_SamplerTraceRays (SamplerTraceRays.cpp:48-53): t and term of every ray are rewritten
 * in place (term = RAY_SURFACE where an operator fired, else RAY_BOUNDARY). */
int  gxy_sample_raylist(gxy_vis *, gxy_raylist_view rays);
/* Sampler over a frame of camera rays: Renderer::local_render with Sampler::Trace and Sampler::HandleTerminatedRays
 * (src/sampler/Sampler.cpp:52-133) on the device.  A ray leaves one sample per firing and continues behind it until it
 * reaches the partition's boundary, then moves to the neighbour.  Every partition keeps its samples on the device
 * (as every rank keeps its own Particles in the reference).  Either one process drives all partitions (nparts >= 1, no
 * communicator), or one process per GPU (gxy_comm_init) passes its one partition: rays that cross into a neighbour then
 * travel as NCCL send/recv pairs and every process must make the call (it is collective; the stats are this process's).
 * stats: primary_rays, traced_rays, forwarded_rays, waves, kernel_launches, device_ms. */
int  gxy_sample(int nparts, gxy_vis *const *parts, const gxy_camera *, int w, int h, gxy_stats *stats);
/* the samples of one partition after gxy_sample: their number; their positions (3 floats each, order unspecified as in
 * the reference, where threads append under a lock); or as a Particles dataset (value 0, Sampler.cpp:83) without
 * leaving the device, ready for gxy_vis_add_particles. */
int  gxy_vis_sample_count(gxy_vis *, long long *n);
int  gxy_vis_download_samples(gxy_vis *, float *xyz);
int  gxy_particles_from_samples(gxy_vis *, gxy_particles **out);

/* ---- frame level ------------------------------------------------------------------------- */
/* Renderer::local_render + the processRays loop + Rendering::AddLocalPixels, all on device
 * (src/renderer/Renderer.cpp:179-269,504-656; Rendering.cpp:125-153).  parts[0..nparts) are the
 * partitions driven by THIS process (normally 1; several partitions on one device are allowed
 * and exchange rays by device copies).  If a communicator is attached (gxy_comm_init) the
 * partitions of all ranks take part and rays are exchanged with NCCL send/recv.
 * lights are unresolved (resolve_lights is applied).  The float framebuffer (w*h*4, y up) stays
 * on the device of parts[0] (image owner = rank 0); fetch it with gxy_frame_download_*. */
int  gxy_render(int nparts, gxy_vis *const *parts, const gxy_camera *, const gxy_lighting *, int w, int h,
                float epsilon, gxy_stats *stats);
/* A RenderingSet in flight.  The reference starts every Rendering of a set before it waits for any of them
 * (src/apps/gxywriter.cpp:196-264; RenderingSet::WaitForDone, RenderingSet.cpp:289-589) and its ray queue interleaves their
 * lists (src/renderer/RayQManager.cpp:69-82), so a rank that has nothing left to do for one Rendering works on the next.
 * gxy_render_submit enqueues one frame on frame slot `slot` (0 .. gxy_render_max_slots()-1) and returns without waiting for
 * the device; gxy_render_wait blocks until that frame is complete, fills `stats` and makes it "the last frame" of parts[0]
 * for gxy_frame_download_*.  Every slot owns its queues, framebuffer and (one process per GPU) peer arena: frames on
 * different slots overlap on the device, and a barrier a rank waits at in one frame costs latency, not throughput.
 * With a communicator attached the calls are collective: every rank submits and waits for the same slots in the same order.
 * Enqueued without any host round trip: geometry-only Visualizations (one GPU, or one process per GPU) and, with one process per
 * GPU, Visualizations with volumes; everything else (volumes on one GPU, PathLines, several partitions in one process) renders
 * synchronously inside the submit call and hands its image over at the wait.
 * gxy_render == submit + wait on slot 0. */
int  gxy_render_submit(int nparts, gxy_vis *const *parts, const gxy_camera *, const gxy_lighting *, int w, int h,
                       float epsilon, int slot);
int  gxy_render_wait(int nparts, gxy_vis *const *parts, int slot, gxy_stats *stats);
int  gxy_render_max_slots(void);
/* Diagnostic (host arithmetic only, no device): the rectangle of 8x4-pixel tiles (x0, y0, nx, ny) over which a rank whose partition box
 * is [lo, hi] generates primary rays on the multi-process frame path -- every pixel whose ray touches the box must lie inside it.
 * Returns 1 (rect valid), 0 (a corner of the box is at or behind the eye plane: the whole image is scanned), -1 (bad arguments). */
int  gxy_debug_tile_rect(const gxy_camera *, int w, int h, const float lo[3], const float hi[3], int rect[4]);
/* D2H of the last frame: float RGBA (y up) ... */
int  gxy_frame_download_rgba32f(gxy_vis *owner, float *fb);
/* ... or RGBA8 rows top-down exactly as ColorImageWriter::Write does (ImageWriter.cpp:30-48) */
int  gxy_frame_download_rgba8(gxy_vis *owner, unsigned char *rgba);
/* The same without blocking: the image of the last frame is converted on the device and copied to `rgba` (page-locked
 * memory from gxy_host_alloc, else the copy is not asynchronous) on a separate copy stream, so that the next gxy_render of
 * this owner overlaps the transfer.  `rgba` is valid after gxy_frame_download_wait, which blocks until every pending
 * asynchronous download of this owner has landed.  Two downloads may be in flight (two device-side staging images);
 * a third one waits for the first.  The reference has no counterpart: its image writer runs after WaitForDone
 * (src/apps/gxywriter.cpp:262-264). */
int  gxy_frame_download_rgba8_async(gxy_vis *owner, unsigned char *rgba);
int  gxy_frame_download_wait(gxy_vis *owner);

/* Page-locked host memory for the buffers handed to gxy_frame_download_* / gxy_trace_raylist: with it
 * the D2H/H2D copies run at PCIe speed without a staging copy (any host pointer is accepted; pageable
 * ones go through the driver's staging path).  The reference allocates RayLists and framebuffers with
 * malloc (framework/smem.cpp:54-79); this is the CUDA equivalent of that allocator. */
int  gxy_host_alloc(size_t bytes, void **out);
void gxy_host_free(void *p);

/* ---- multi-process (one process per GPU) --------------------------------------------------- */
/* replaces MessageManager's MPI transport for SendRaysMsg / SendPixelsMsg
 * (src/framework/MessageManager.cpp:321-434; Renderer.cpp:732-836) inside one NVLink box. */
int  gxy_comm_unique_id(unsigned char id[128]);
int  gxy_comm_init(gxy_context *, int rank, int nranks, const unsigned char id[128]);
int  gxy_comm_destroy(gxy_context *);

#ifdef __cplusplus
}
#endif
#endif
