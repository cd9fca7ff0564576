// gxy_fused.cu -- the frame path for geometry-only Visualizations: two persistent kernels that keep
// a ray in registers from its creation to its framebuffer contribution.
//
//   fused_primary_kernel    Camera::SpawnRays (Camera.cpp:379-493) -> TraceRays_TraceRays
//                           (TraceRays.ispc:326-623) -> ambient/diffuseLighting (:735-761, :859-923) ->
//                           Renderer::Classify/AssignDestinations (Renderer.cpp:304-454) ->
//                           Rendering::AddLocalPixels (Rendering.cpp:125-153) + a compact record per
//                           surface hit
//   fused_secondary_kernel  TraceRays_generateAORays/_generateShadowRays (:625-733, :763-857) from
//                           those records -> TraceRays_TraceRays -> Classify -> AddLocalPixels
//
// The list-based kernels (gxy_kernels.cu) materialise every ray of a wave in a 100-B/ray RayList and
// run generate / trace / hit-scan / spawn / classify / accumulate as separate launches with host
// round trips between them.  On one partition no ray ever needs to exist in memory: only rays that
// leave through an internal face are written out ("spilled") as ordinary RayList entries and then
// take the list path (exchange, re-trace on the neighbour).  Arithmetic per ray is the same device
// code as the list path (gxy_shade.cuh, gxy_traverse.cuh); only the order of the framebuffer
// additions differs, which is unordered in the reference too (Rendering.cpp:91-153).
#include "gxy_internal.h"
#include "gxy_shade.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace gxy {

#define FULLMASK 0xffffffffu
// BVH levels below the root's children that the generation kernel tests before it queues a primary ray for the trace kernel
// experiment switches of the persistent trace kernels (variant builds: make VARIANT=.. EXTRA=-D.. variant)
#ifndef GXY_NODE_PREFETCH
#define GXY_NODE_PREFETCH 0   // node_step<>: 1 = L2 prefetch of the next node + the new primitive records, 2 = the same into L1
#endif
#ifndef GXY_MIN_BLOCKS
#define GXY_MIN_BLOCKS 8      // __launch_bounds__(128, N): 8 -> 64 registers; 10 -> 48; 12 -> 40 (spills)
#endif
#ifndef GXY_CURVE_BLOCKS
#define GXY_CURVE_BLOCKS 3    // resident CTAs of the CURVES instantiations (PathLines: the round-Bezier test needs ~150 registers)
#endif
#ifndef GXY_CULL_LEVELS
#define GXY_CULL_LEVELS 3  // measured on C5: 1 -> 1.449 ms, 2 -> 1.449, 3 -> 1.416, 4 -> 1.425 per frame
#endif

// pixel order of the primary queue: 8x4 tiles, so that the 32 lanes of a warp start as a compact beam
// band / n_bands: the queue covers the tile rows band, band + n_bands, ... (frame split into interleaved bands that are
// pipelined on two streams, gxy_render)
__device__ __forceinline__ void tile_pixel(unsigned idx, int tiles_x, int band, int n_bands, int &x, int &y) {
  const unsigned tile = idx >> 5, in = idx & 31u;
  x = (int)(tile % (unsigned)tiles_x) * 8 + (int)(in & 7u);
  y = ((int)(tile / (unsigned)tiles_x) * n_bands + band) * 4 + (int)(in >> 3);
}

__device__ __forceinline__ void write_spill(const Rays &S, unsigned j, float3 org, float3 dir, float r, float g, float b, float o, float t,
                                            float tMax, int x, int y, int type, int term, int cls) {
  S.ox[j] = org.x; S.oy[j] = org.y; S.oz[j] = org.z;
  S.dx[j] = dir.x; S.dy[j] = dir.y; S.dz[j] = dir.z;
  S.r[j] = r; S.g[j] = g; S.b[j] = b; S.o[j] = o;
  S.t[j] = t; S.tMax[j] = tMax;
  S.x[j] = x; S.y[j] = y; S.type[j] = type; S.term[j] = term; S.classification[j] = cls;
}

// position of each flagged lane in a global append list; all lanes of `group` must call (converged)
__device__ __forceinline__ unsigned group_append(unsigned group, unsigned lane, bool flag, unsigned *counter) {
  const unsigned m = __ballot_sync(group, flag);
  if (m == 0u) return 0u;
  const int leader = __ffs((int)m) - 1;
  unsigned base = 0u;
  if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
  base = __shfl_sync(group, base, leader);
  return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}

// ---- peer inboxes (PeerTable, gxy_internal.h) -----------------------------------------------------
__device__ __forceinline__ PeerCtrl *peer_ctrl(const PeerTable &T, int r) { return reinterpret_cast<PeerCtrl *>(T.base[r]); }
__device__ __forceinline__ float4 *peer_inbox(const PeerTable &T, int r, int parity) {
  return reinterpret_cast<float4 *>(T.base[r] + T.off_inbox[parity]);
}
// Append one 64-byte record per flagged lane to the inbox[parity] of its destination rank.  All lanes of
// `group` must call (converged).  Lanes bound for the same rank share one system-scope atomicAdd on that
// rank's counter (an NVLink round trip) and write neighbouring records.
__device__ __forceinline__ void peer_push(const PeerTable &T, int parity, unsigned group, unsigned lane, bool flag, int dest, float3 org,
                                          float3 dir, float r, float g, float b, float o, float t, float tMax, int x, int y, int type, int term,
                                          int *error_flag) {
  if (dest < 0 || dest >= T.nranks) flag = false;
  unsigned todo = __ballot_sync(group, flag);
  const unsigned lt = (1u << lane) - 1u;
  while (todo != 0u) {
    const int leader = __ffs((int)todo) - 1;
    const int d = __shfl_sync(group, dest, leader);
    const unsigned same = __ballot_sync(group, flag && dest == d);
    unsigned base = 0u;
    if ((int)lane == leader) base = atomicAdd_system(&peer_ctrl(T, d)->inbox_count[parity], (unsigned)__popc(same));
    base = __shfl_sync(group, base, leader);
    if (flag && dest == d) {
      const unsigned pos = base + (unsigned)__popc(same & lt);
      if (pos < T.inbox_cap) {
        float4 *rec = peer_inbox(T, d, parity) + 4 * (size_t)pos;
        rec[0] = make_float4(org.x, org.y, org.z, dir.x);
        rec[1] = make_float4(dir.y, dir.z, t, tMax);
        rec[2] = make_float4(r, g, b, o);
        rec[3] = make_float4(__int_as_float(x), __int_as_float(y), __int_as_float(type), __int_as_float(term));
      } else *error_flag = 3;
    }
    todo &= ~same;
  }
}

// Camera::SpawnRays for the whole window (Camera.cpp:379-493): one thread per pixel in tile order.  A kept ray is
// clipped to the local box and tested against the top two levels of the BVH (may_hit_anything): rays that can
// reach no primitive are finished here at full SIMD width exactly as the trace kernel finishes a ray without a
// hit (term BOUNDARY/TIMEOUT -> Classify: TERMINATED with colour 0, or on to the neighbour partition); the
// others are appended (unordered) to `out` (columns ox..dz t x y; tMax = FLT_MAX, type PRIMARY implied).
__global__ void __launch_bounds__(256)
    gen_primary_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevCamera C, int w, int h, int tiles_x, int band,
                       int n_bands, unsigned n_queue, Rays out, Rays spill, unsigned spill_cap, FusedQueues *__restrict__ q) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  bool kept = false, queued = false;
  int x = 0, y = 0;
  float3 o3 = f3(0.f, 0.f, 0.f), d3 = f3(0.f, 0.f, 0.f);
  if (idx < n_queue) {
    tile_pixel(idx, tiles_x, band, n_bands, x, y);
    kept = x < w && y < h && spawn_pixel(P, C, x, y, o3, d3);
  }
  bool do_spill = false, terminated = false;
  int term = 0, cls = CLS_UNDETERMINED;
  float ray_t = 0.f;
  if (kept) {
    RayCtx rc;
    TravState st;
    PendingRay pr;
    queued = setup_ray_values(P, 0, true, o3, d3, 0.f, FLT_MAX, 0, rc, st, pr) && may_hit_anything<GXY_CULL_LEVELS>(P.nodes, rc);
    if (!queued) {
      ray_t = rc.tfar;
      if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
      else if (ray_t == FLT_MAX) term |= RAY_TIMEOUT;
      cls = classify_values(P, RAY_PRIMARY, term, o3.x, o3.y, o3.z, d3.x, d3.y, d3.z);
      terminated = cls == CLS_TERMINATED;  // carries (0,0,0,0): nothing to add
      do_spill = cls >= 0;
    }
  }
  const unsigned km = __ballot_sync(FULLMASK, kept), tm = __ballot_sync(FULLMASK, terminated);
  if (lane == 0u) {
    if (km) atomicAdd(&q->n_generated, (unsigned)__popc(km));
    if (tm) atomicAdd(&q->n_terminated, (unsigned long long)__popc(tm));
  }
  const unsigned pos = group_append(FULLMASK, lane, queued, &q->n_primary32);
  if (queued) {
    out.ox[pos] = o3.x; out.oy[pos] = o3.y; out.oz[pos] = o3.z;
    out.dx[pos] = d3.x; out.dy[pos] = d3.y; out.dz[pos] = d3.z;
    out.t[pos] = 0.f;
    out.x[pos] = x; out.y[pos] = y;
  }
  const unsigned sp = group_append(FULLMASK, lane, do_spill, &q->n_spill);
  if (do_spill) {
    if (sp < spill_cap) write_spill(spill, sp, o3, d3, 0.f, 0.f, 0.f, 0.f, ray_t, FLT_MAX, x, y, RAY_PRIMARY, term, cls);
    else *P.error_flag = 3;
  }
}

// The same for one rank of a multi-process frame.  Every rank holds the PartProxy of every partition (box, neighbours,
// top two BVH levels), so it can follow a pixel's ray through the partitions exactly as generation -> trace -> Classify
// -> forward would (Camera.cpp:431-441, TraceRays.ispc:377-418, Renderer.cpp:304-454), for as long as the ray cannot hit
// anything: the partition that the ray enters first originates it (statistics), every partition it leaves without a
// possible hit forwards it (statistics), and the first partition whose proxy it MAY hit takes it into its trace queue
// with t = the exit distance of the partition before, which is what the forwarded record would have carried.  All ranks
// evaluate the same arithmetic on the same proxies, so exactly one rank queues (or terminates) each ray and nothing
// crosses NVLink for the ~80 % of the primaries that only pass through.  Rays that may hit a partition but do not are
// forwarded for real by the trace kernel.
__global__ void __launch_bounds__(256)
    gen_primary_peer_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevCamera C, int w, int h, int tiles_x,
                            unsigned n_queue, Rays out, unsigned queue_cap, FusedQueues *__restrict__ q, int me, int nranks,
                            const PartProxy *__restrict__ prox, int skip_far, int tile_x0, int tile_y0) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  int x = 0, y = 0;
  float3 o3 = f3(0.f, 0.f, 0.f), d3 = f3(0.f, 0.f, 0.f);
  float gmin = 0.f;
  bool in_data = false;
  if (idx < n_queue) {
    // tiles_x tiles per row of the launch's window, which starts at tile (tile_x0, tile_y0): the screen rectangle this rank's box
    // projects into (peer_tile_rect), or the whole image
    tile_pixel(idx, tiles_x, 0, 1, x, y);
    x += tile_x0 * 8; y += tile_y0 * 4;
    in_data = x < w && y < h && camera_ray(P, C, x, y, o3, d3, gmin);
  }
  if (in_data && skip_far) {
    // A ray that stays clear of this rank's box cannot be originated, forwarded, terminated or queued HERE (all of that needs the
    // ray inside the box, at worst grazing a face): it is somebody else's pixel.  Plain slab test against the box grown by a
    // margin that is orders of magnitude above the rounding of the box arithmetic below, so grazing rays still take the full path.
    const float3 lo = prox[me].lmin, hi = prox[me].lmax;
    const float mx = 1e-3f * (hi.x - lo.x) + 1e-4f, my = 1e-3f * (hi.y - lo.y) + 1e-4f, mz = 1e-3f * (hi.z - lo.z) + 1e-4f;
    float t0 = 0.f, t1 = FLT_MAX;
    const float dd[3] = {d3.x, d3.y, d3.z}, oo[3] = {o3.x, o3.y, o3.z}, l3[3] = {lo.x - mx, lo.y - my, lo.z - mz}, h3[3] = {hi.x + mx, hi.y + my, hi.z + mz};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (fabsf(dd[a]) < 1e-12f) {
        if (oo[a] < l3[a] || oo[a] > h3[a]) t1 = -1.f;
      } else {
        const float r = 1.0f / dd[a];
        const float ta = (l3[a] - oo[a]) * r, tb = (h3[a] - oo[a]) * r;
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
      }
    }
    if (t0 > t1) in_data = false;
  }
  unsigned n_gen = 0u, n_fwd = 0u, n_term = 0u;
  bool queued = false;
  float t_queued = 0.f;
  if (in_data) {
    for (int s = 0; s < nranks; s++) {
      if (!first_brick(prox[s].lmin, prox[s].lmax, o3, d3, gmin)) continue;
      if (s == me) n_gen++;
      int cur = s;
      float t0 = 0.f;
      for (int hops = 0;; hops++) {
        const PartProxy &pp = prox[cur];
        RayCtx rc;
        TravState st;
        PendingRay pr;
        const bool may = setup_ray_box(pp.lmin, pp.lmax, pp.has_prims != 0, 0, true, o3, d3, t0, FLT_MAX, 0, rc, st, pr) &&
                         may_hit_anything(pp.nodes, rc);
        if (may || hops >= 2 * nranks) {  // (the hop bound only guards against a cyclic neighbour table)
          if (cur == me && !queued) { queued = true; t_queued = t0; }
          else if (cur == me) {  // a pixel that two partitions both originate (boundary tie) and both rays end up here
            const unsigned pos2 = atomicAdd(&q->n_primary32, 1u);
            if (pos2 >= queue_cap) { *P.error_flag = 3; break; }
            out.ox[pos2] = o3.x; out.oy[pos2] = o3.y; out.oz[pos2] = o3.z;
            out.dx[pos2] = d3.x; out.dy[pos2] = d3.y; out.dz[pos2] = d3.z;
            out.t[pos2] = t0;
            out.x[pos2] = x; out.y[pos2] = y;
          }
          break;
        }
        const float ray_t = rc.tfar;
        int term = 0;
        if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
        else if (ray_t == FLT_MAX) term |= RAY_TIMEOUT;
        const int cls = classify_box(pp.lmin, pp.lmax, pp.neighbors, RAY_PRIMARY, term, o3.x, o3.y, o3.z, d3.x, d3.y, d3.z);
        if (cls == CLS_TERMINATED) { if (cur == me) n_term++; break; }
        if (cls < 0 || cls >= nranks) break;  // KEEP_HERE without a hit: the reference drops it too
        if (cur == me) n_fwd++;
        t0 = ray_t;
        cur = cls;
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    n_gen += __shfl_down_sync(FULLMASK, n_gen, off);
    n_fwd += __shfl_down_sync(FULLMASK, n_fwd, off);
    n_term += __shfl_down_sync(FULLMASK, n_term, off);
  }
  if (lane == 0u) {
    if (n_gen) atomicAdd(&q->n_generated, n_gen);
    if (n_fwd) atomicAdd(&q->n_virtual, n_fwd);
    if (n_term) atomicAdd(&q->n_terminated, (unsigned long long)n_term);
  }
  const unsigned pos = group_append(FULLMASK, lane, queued, &q->n_primary32);
  if (queued && pos >= queue_cap) { *P.error_flag = 3; queued = false; }
  if (queued) {
    out.ox[pos] = o3.x; out.oy[pos] = o3.y; out.oz[pos] = o3.z;
    out.dx[pos] = d3.x; out.dy[pos] = d3.y; out.dz[pos] = d3.z;
    out.t[pos] = t_queued;
    out.x[pos] = x; out.y[pos] = y;
  }
}

// this rank's PartProxy, written into its own arena (read by every rank after the next flag barrier)
__global__ void __launch_bounds__(32) proxy_publish_kernel(const __grid_constant__ SceneParams P, PartProxy *__restrict__ dst) {
  const unsigned lane = threadIdx.x;
  if (lane == 0u) {
    dst->lmin = P.lmin; dst->lmax = P.lmax;
    for (int f = 0; f < 6; f++) dst->neighbors[f] = P.neighbors[f];
    dst->has_prims = P.n_prims > 0 ? 1 : 0;
    dst->pad[0] = dst->pad[1] = dst->pad[2] = 0;
  }
  if (P.n_prims <= 0) return;
  const WideNode root = P.nodes[0];
  const unsigned n_inner = (unsigned)__popc((unsigned)root.imask);
  if (lane == 0u) {
    WideNode r = root;
    r.child_base = 1u;
    dst->nodes[0] = r;
  } else if (lane <= 8u && lane - 1u < n_inner) {
    dst->nodes[lane] = P.nodes[root.child_base + (lane - 1u)];
  }
}

__global__ void __launch_bounds__(256) proxy_gather_kernel(const __grid_constant__ PeerTable T, PartProxy *__restrict__ out) {
  constexpr unsigned W = sizeof(PartProxy) / 16u;
  for (unsigned i = threadIdx.x; i < W * (unsigned)T.nranks; i += blockDim.x) {
    const unsigned r = i / W, k = i - r * W;
    reinterpret_cast<uint4 *>(out + r)[k] = reinterpret_cast<const uint4 *>(T.base[r] + T.off_proxy)[k];
  }
}

// A persistent trace launch is sized for the worst case (every pixel hits, every secondary ray exists); the real length of its queue
// is only known on the device, and it is final when the kernel starts.  CTAs that the queue cannot give rays_per_thread rays per thread
// leave at once, so a launch with little work holds few SM slots and the launches of OTHER frames in flight run beside it instead of
// behind it: across ranks a wave often carries a few ten thousand rays, and 1184 resident CTAs of such a launch would block the SMs for
// the whole latency-bound lifetime of those rays.  rays_per_thread = 0: every CTA stays (the behaviour before).
__device__ __forceinline__ bool surplus_cta(unsigned n_queue, int rays_per_thread) {
  return rays_per_thread > 0 && blockIdx.x > 0u &&
         (unsigned long long)blockIdx.x * (unsigned long long)(GXY_TRACE_THREADS * rays_per_thread) >= (unsigned long long)n_queue;
}

// One iteration of a persistent trace warp (all 32 lanes, converged): a node step for every lane that has no primitives pending, then
// the cooperative primitive passes -- but only once prim_t lanes hold a primitive group, or no lane is left that could do a node
// step.  A pass costs the whole warp ~330 instructions whether it serves one owner or eight; with a node step per lane and
// iteration only a lane or two of a warp reach a leaf, so deferring the passes (prim_t > 1) looked like a saving.  MEASURED on C5 with 4
// frames in flight (tools/flight_sweep.py, profiles/r02_b_primt_sweep.jsonl): prim_t 1: 1.111 ms per frame, 2: 1.116, 3: 1.120, 4: 1.132,
// 6: 1.155, 8: 1.180, 12: 1.224 -- the lanes that wait lose more node steps than the fuller passes save, so the default stays 1
// (GXY_PRIM_T).  Lanes that wait keep their node group; the order in which primitives are tested never decides a result (smallest t,
// ties on the lowest ids), and the parity tests pass with prim_t = 4.
template <bool CURVES>
__device__ __forceinline__ void trace_iteration(const SceneParams &P, const RayCtx &rc, TravState &st, bool &trav, const bool anyhit,
                                                uint2 *__restrict__ stack, uint2 *__restrict__ lstack, unsigned char *owner_slot,
                                                const unsigned lane, const unsigned lt_mask, const int prim_t) {
  const bool node_ready = trav && st.tg.y == 0u;
  if (node_ready) node_step<GXY_NODE_PREFETCH>(P, rc, st, stack, lstack);
  const unsigned owners = __ballot_sync(FULLMASK, trav && st.tg.y != 0u);
  if (owners != 0u) {
    // (second ballot: lanes that will still be able to take a node step next iteration without a primitive pass)
    if (prim_t <= 1 || __popc(owners) >= prim_t ||
        __ballot_sync(FULLMASK, trav && st.tg.y == 0u && (st.ng.y > 0x00ffffffu || st.sp > 0)) == 0u)
      coop_prim_passes<CURVES>(P, rc, st, trav, anyhit, owner_slot, lane, lt_mask);
  }
  if (trav) trav = trav_advance(st, stack, lstack);
}

// Trace of the generated primaries (persistent warps, dynamic fetch, cooperative primitive tests).
// A surface hit leaves a 6-word raw record (ray, t, u, v, ids, record) for shade_hits_kernel; a miss is
// classified here: TERMINATED (adds nothing: its colour is 0) or spilled towards a neighbour.
template <int FETCH_T, int MIN_BLOCKS, bool PEER, bool CURVES = false>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, CURVES ? GXY_CURVE_BLOCKS : MIN_BLOCKS)
    primary_trace_kernel(const __grid_constant__ SceneParams P, Rays R, unsigned *__restrict__ raw, unsigned raw_stride, Rays spill,
                         unsigned spill_cap, FusedQueues *__restrict__ q, const __grid_constant__ PeerTable T, int prim_t, int rays_per_thread) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  __shared__ unsigned char owner_of[GXY_TRACE_THREADS / 32][8];
  uint2 lstack[GXY_STACK_LOCAL];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned n_queue = q->n_primary32;
  if (surplus_cta(n_queue, rays_per_thread)) return;
  RayCtx rc;
  TravState st;
  PendingRay pr;
  pr.ray = -1; pr.tExit = 0.f; pr.anyhit = false; pr.opaque = false;
  st.tg.y = 0u; st.ng.y = 0u;
  rc.org = f3(0.f, 0.f, 0.f); rc.dir = f3(1.f, 1.f, 1.f); rc.tnear = 0.f; rc.tfar = 0.f;
  st.best_t = 0.f; st.best_u = 0.f; st.best_v = 0.f; st.best_key = GXY_NO_HIT; st.best_rec = 0u;
  bool trav = false, exhausted = false;
  while (true) {
    const unsigned m_idle = __ballot_sync(FULLMASK, !trav);
    if (m_idle == FULLMASK || (!exhausted && __popc(m_idle) >= FETCH_T)) {
      bool ex_local = false;
      if (!trav) {
        const bool fin = pr.ray >= 0;
        const bool has_hit = fin && st.best_key != GXY_NO_HIT;
        bool do_spill = false, terminated = false;
        int term = 0, cls = CLS_UNDETERMINED;
        if (fin && !has_hit) {
          const float ray_t = rc.tfar;
          if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
          else if (ray_t == FLT_MAX) term |= RAY_TIMEOUT;
          cls = classify_values(P, RAY_PRIMARY, term, rc.org.x, rc.org.y, rc.org.z, R.dx[pr.ray], R.dy[pr.ray], R.dz[pr.ray]);
          terminated = cls == CLS_TERMINATED;  // carries (0,0,0,0): nothing to add
          do_spill = cls >= 0;
        }
#ifdef GXY_TRAV_COUNTERS
        if (fin) atomicAdd(&q->nodes, (unsigned long long)st.n_nodes);
#endif
        // all appends and the fetch of this refill are issued back to back by one lane
        const unsigned hm = __ballot_sync(m_idle, has_hit), sm = __ballot_sync(m_idle, do_spill), tm = __ballot_sync(m_idle, terminated);
        const int leader = __ffs((int)m_idle) - 1;
        const unsigned cnt = (unsigned)__popc(m_idle);
        unsigned hbase = 0u, sbase = 0u, fbase = 0u;
        if ((int)lane == leader) {
          if (hm) hbase = atomicAdd(&q->n_hits, (unsigned)__popc(hm));
          if (sm) sbase = atomicAdd(&q->n_spill, (unsigned)__popc(sm));
          if (tm) atomicAdd(&q->n_terminated, (unsigned long long)__popc(tm));
          if (!exhausted) fbase = atomicAdd(&q->pixel_head, cnt);
        }
        hbase = __shfl_sync(m_idle, hbase, leader);
        sbase = __shfl_sync(m_idle, sbase, leader);
        fbase = __shfl_sync(m_idle, fbase, leader);
        if (has_hit) {
          const unsigned hp = hbase + (unsigned)__popc(hm & lt_mask);
          raw[hp] = (unsigned)pr.ray;
          raw[hp + raw_stride] = __float_as_uint(st.best_t);
          raw[hp + 2u * raw_stride] = __float_as_uint(st.best_u);
          raw[hp + 3u * raw_stride] = __float_as_uint(st.best_v);
          raw[hp + 4u * raw_stride] = st.best_key;
          raw[hp + 5u * raw_stride] = st.best_rec;
        }
        if (PEER) {
          float3 d0 = f3(0.f, 0.f, 0.f);
          int px = 0, py = 0;
          if (do_spill) { d0 = f3(R.dx[pr.ray], R.dy[pr.ray], R.dz[pr.ray]); px = R.x[pr.ray]; py = R.y[pr.ray]; }
          peer_push(T, 0, m_idle, lane, do_spill, cls, rc.org, d0, 0.f, 0.f, 0.f, 0.f, rc.tfar, FLT_MAX, px, py, RAY_PRIMARY, term, P.error_flag);
        } else if (do_spill) {
          const unsigned sp = sbase + (unsigned)__popc(sm & lt_mask);
          if (sp < spill_cap)
            write_spill(spill, sp, rc.org, f3(R.dx[pr.ray], R.dy[pr.ray], R.dz[pr.ray]), 0.f, 0.f, 0.f, 0.f, rc.tfar, FLT_MAX, R.x[pr.ray],
                        R.y[pr.ray], RAY_PRIMARY, term, cls);
          else *P.error_flag = 3;
        }
        pr.ray = -1;
        if (!exhausted) {
          const unsigned my = fbase + (unsigned)__popc(m_idle & lt_mask);
          ex_local = fbase + cnt >= n_queue;
          if (my < n_queue)
            trav = setup_ray_values(P, (int)my, true, f3(R.ox[my], R.oy[my], R.oz[my]), f3(R.dx[my], R.dy[my], R.dz[my]), R.t[my], FLT_MAX, 0,
                                    rc, st, pr);
        }
      }
      exhausted = exhausted || __any_sync(FULLMASK, ex_local);
      if (exhausted && __ballot_sync(FULLMASK, pr.ray >= 0) == 0u) break;
    }
    trace_iteration<CURVES>(P, rc, st, trav, false, stack, lstack, owner_of[warp], lane, lt_mask, prim_t);
  }
}

// postIntersect + ambient/diffuse lighting + framebuffer add for the surface hits [q->hits_done, q->n_hits), one
// thread per hit (full SIMD efficiency, unlike a finish inside the persistent kernel), and the record the
// secondary kernel generates AO/shadow rays from.  The traced ray is read from the primary list (AOS = false)
// or from the 64-byte records of an inbox (AOS = true; the ray's own colour is 0 on this path).
template <bool AOS, bool CURVES = false>
__global__ void __launch_bounds__(256)
    shade_hits_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevLights L, Rays R, const float4 *__restrict__ inbox,
                      const unsigned *__restrict__ raw, unsigned raw_stride, int w, float4 *__restrict__ fb, Rays hits,
                      FusedQueues *__restrict__ q, float epsilon) {
  const unsigned n_hits = q->n_hits;
  unsigned n_term = 0u;
  for (unsigned hidx = q->hits_done + blockIdx.x * blockDim.x + threadIdx.x; hidx < n_hits; hidx += gridDim.x * blockDim.x) {
    const unsigned i = raw[hidx];
    RayCtx rc;
    float3 dir0;
    int px, py;
    if (AOS) {
      const float4 a = inbox[4 * (size_t)i], b = inbox[4 * (size_t)i + 1], d = inbox[4 * (size_t)i + 3];
      rc.org = f3(a.x, a.y, a.z);
      dir0 = f3(a.w, b.x, b.y);
      px = __float_as_int(d.x); py = __float_as_int(d.y);
    } else {
      rc.org = f3(R.ox[i], R.oy[i], R.oz[i]);
      dir0 = f3(R.dx[i], R.dy[i], R.dz[i]);
      px = R.x[i]; py = R.y[i];
    }
    rc.dir = dir0;
    if (rc.dir.x == 0.f) rc.dir.x = 1e-6f;  // TraceRays.ispc:377-379
    if (rc.dir.y == 0.f) rc.dir.y = 1e-6f;
    if (rc.dir.z == 0.f) rc.dir.z = 1e-6f;
    TravState st;
    if (CURVES) {
      // the normal of a curve hit is recomputed by running the segment's test again (trav_fetch_hit<true>): it needs the very
      // interval the trace used, i.e. the same box clip
      PendingRay pr0;
      float t0 = 0.f, t1 = FLT_MAX;
      if (AOS) { const float4 b = inbox[4 * (size_t)i + 1]; t0 = b.z; t1 = b.w; }
      else t0 = R.t[i];
      setup_ray_values(P, 0, true, rc.org, dir0, t0, t1, 0, rc, st, pr0);
    }
    st.best_t = __uint_as_float(raw[hidx + raw_stride]);
    st.best_u = __uint_as_float(raw[hidx + 2u * raw_stride]);
    st.best_v = __uint_as_float(raw[hidx + 3u * raw_stride]);
    st.best_key = raw[hidx + 4u * raw_stride];
    st.best_rec = raw[hidx + 5u * raw_stride];
    Hit1 h1;
    trav_fetch_hit<CURVES>(P, rc, st, h1);
    float3 col, Ns;
    float opacity;
    shade_geometry_hit<CURVES>(P, h1, rc.dir, col, opacity, Ns);
    int term = RAY_SURFACE;
    if (opacity > 0.999f) term |= RAY_OPAQUE;
    HitPoint hp;
    hp.ox = rc.org.x; hp.oy = rc.org.y; hp.oz = rc.org.z; hp.dx = dir0.x; hp.dy = dir0.y; hp.dz = dir0.z; hp.t = st.best_t;
    hp.sn = Ns; hp.sr = col.x; hp.sg = col.y; hp.sb = col.z; hp.o = 0.f; hp.px = px; hp.py = py;
    float r = 0.f, g = 0.f, b = 0.f, o = 0.f;
    light_primary(L, hp, r, g, b, o);
    const int cls = classify_values(P, RAY_PRIMARY, term, hp.ox, hp.oy, hp.oz, hp.dx, hp.dy, hp.dz);
    if (cls == CLS_TERMINATED) {
      n_term++;
      atomicAdd(fb + ((size_t)hp.py * w + hp.px), make_float4(r, g, b, o));
    } else if (cls == CLS_KEEP_HERE) {
      // a translucent surface (opacity <= 0.999): the reference re-queues the primary behind the hit (Renderer.cpp:304-421); this
      // path has no keeper list, and no shader here produces one today (shade_geometry_hit: opacity 1) -- refuse rather than lose it
      *P.error_flag = 7;
    }
    hits.ox[hidx] = hp.ox; hits.oy[hidx] = hp.oy; hits.oz[hidx] = hp.oz;
    hits.dx[hidx] = hp.dx; hits.dy[hidx] = hp.dy; hits.dz[hidx] = hp.dz;
    hits.t[hidx] = hp.t;
    hits.nx[hidx] = Ns.x; hits.ny[hidx] = Ns.y; hits.nz[hidx] = Ns.z;
    hits.sr[hidx] = col.x; hits.sg[hidx] = col.y; hits.sb[hidx] = col.z;
    hits.o[hidx] = 0.f;  // opacity before diffuseLighting (no volumes on this path)
    hits.x[hidx] = hp.px; hits.y[hidx] = hp.py;
    // the j-independent part of generateAORays, once per hit: origin -> r g b, b0 -> sample so tMax,
    // b1 -> type term classification (spare columns of the record)
    float3 aorg, b0, b1;
    ao_basis(hp, epsilon, aorg, b0, b1);
    hits.r[hidx] = aorg.x; hits.g[hidx] = aorg.y; hits.b[hidx] = aorg.z;
    hits.sample[hidx] = b0.x; hits.so[hidx] = b0.y; hits.tMax[hidx] = b0.z;
    hits.type[hidx] = __float_as_int(b1.x); hits.term[hidx] = __float_as_int(b1.y); hits.classification[hidx] = __float_as_int(b1.z);
  }
  for (int off = 16; off > 0; off >>= 1) n_term += __shfl_down_sync(FULLMASK, n_term, off);
  if (n_term != 0u && (threadIdx.x & 31u) == 0u) atomicAdd(&q->n_terminated, (unsigned long long)n_term);
}

// queue order of the secondary rays: all AO rays (hit-major), then all shadow rays (hit-major), as the
// reference lays out its SECONDARY list (TraceRays.cpp:89-122): keeps the long, parallel shadow rays of
// neighbouring hits together in a warp instead of mixing them with the short AO rays
__device__ __forceinline__ void secondary_index(unsigned k, unsigned n_hits, int n_ao, int n_sh, int &hi, int &j) {
  const unsigned n_ao_rays = n_hits * (unsigned)n_ao;
  if (k < n_ao_rays) { hi = (int)(k / (unsigned)n_ao); j = (int)(k - (unsigned)hi * (unsigned)n_ao); }
  else { const unsigned k2 = k - n_ao_rays; hi = (int)(k2 / (unsigned)n_sh); j = n_ao + (int)(k2 - (unsigned)hi * (unsigned)n_sh); }
}

// secondary ray j of hit record hi: AO rays first (from the prepared basis), then one shadow ray per light
__device__ __forceinline__ SecRay make_secondary(const DevLights &L, const Rays &hits, int hi, int j, float epsilon, const float *tx,
                                                 const float *ty, const float *tz) {
  if (j < L.n_ao) {
    const float3 sn = f3(hits.nx[hi], hits.ny[hi], hits.nz[hi]);
    const float3 org = f3(hits.r[hi], hits.g[hi], hits.b[hi]);
    const float3 b0 = f3(hits.sample[hi], hits.so[hi], hits.tMax[hi]);
    const float3 b1 = f3(__int_as_float(hits.type[hi]), __int_as_float(hits.term[hi]), __int_as_float(hits.classification[hi]));
    return ao_ray_from_basis(L, sn, org, b0, b1, hits.sr[hi], hits.sg[hi], hits.sb[hi], hits.o[hi], hits.x[hi], hits.y[hi], j, epsilon, tx, ty,
                             tz);
  }
  const HitPoint hp = load_hit_point(hits, hi);
  const float Kd = L.Kd / L.n_lights;
  const float o_lit = hp.o + Kd * (1 - hp.o) * hp.o;  // the o diffuseLighting leaves behind (:918-920)
  return make_shadow_ray(L, hp, j - L.n_ao, epsilon, o_lit);
}

template <int FETCH_T, int MIN_BLOCKS, bool PEER, bool CURVES = false>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, CURVES ? GXY_CURVE_BLOCKS : MIN_BLOCKS)
    fused_secondary_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ DevLights L, int w, int h, int nsec,
                           float4 *__restrict__ fb, Rays hits, Rays spill, unsigned spill_cap, FusedQueues *__restrict__ q, float epsilon,
                           int anyhit_secondary, const __grid_constant__ PeerTable T, int parity_out, int prim_t, int rays_per_thread) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  __shared__ unsigned char owner_of[GXY_TRACE_THREADS / 32][8];
  __shared__ float ao_tab[3][256];  // divergent indices: shared memory, not the constant cache
  uint2 lstack[GXY_STACK_LOCAL];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) { ao_tab[0][k] = c_ao_x[k]; ao_tab[1][k] = c_ao_y[k]; ao_tab[2][k] = c_ao_z[k]; }
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  // hit records of this launch: the wave's own [hits_done, n_hits) on the single-process path; on the peer path the
  // records of the PREVIOUS wave [sec_lo, sec_hi) (wave_epilogue_kernel), so that the rays a rank forwards while
  // tracing become visible to its neighbours one barrier before the rank starts on its own secondaries
  const unsigned h0 = PEER ? q->sec_lo : q->hits_done;
  const unsigned n_hits = (PEER ? q->sec_hi : q->n_hits) - h0;
  if (n_hits == 0u) return;
  const unsigned long long total64 = (unsigned long long)n_hits * (unsigned)nsec;
  const unsigned n_queue = total64 > 0xfffffff0ull ? 0xfffffff0u : (unsigned)total64;
  if (surplus_cta(n_queue, rays_per_thread)) return;
  RayCtx rc;
  TravState st;
  PendingRay pr;
  pr.ray = -1; pr.tExit = 0.f; pr.anyhit = false; pr.opaque = false;
  st.tg.y = 0u; st.ng.y = 0u;
  rc.org = f3(0.f, 0.f, 0.f); rc.dir = f3(1.f, 1.f, 1.f); rc.tnear = 0.f; rc.tfar = 0.f;
  st.best_t = 0.f; st.best_key = GXY_NO_HIT;
  bool trav = false, exhausted = false;
  while (true) {
    const unsigned m_idle = __ballot_sync(FULLMASK, !trav);
    if (m_idle == FULLMASK || (!exhausted && __popc(m_idle) >= FETCH_T)) {
      bool ex_local = false;
      if (!trav) {
        bool do_spill = false, terminated = false;
        SecRay s;
        s.org = f3(0.f, 0.f, 0.f); s.dir = f3(0.f, 0.f, 0.f); s.r = s.g = s.b = 0.f; s.tMax = 0.f; s.type = 0;
        float ray_t = 0.f;
        int px = 0, py = 0, term = 0, cls = CLS_UNDETERMINED;
        if (pr.ray >= 0) {
          const bool found = st.best_key != GXY_NO_HIT;
          ray_t = found ? st.best_t : rc.tfar;
          // only rays that were occluded or reached the box need their colour / exact direction again
          if (found || ray_t == pr.tExit || pr.opaque) {
            int hi, j;
            secondary_index((unsigned)pr.ray, n_hits, L.n_ao, nsec - L.n_ao, hi, j);
            hi += (int)h0;
            s = make_secondary(L, hits, hi, j, epsilon, ao_tab[0], ao_tab[1], ao_tab[2]);
            px = hits.x[hi]; py = hits.y[hi];
            term = (min3f(s.r, s.g, s.b) >= 1.0f) ? RAY_OPAQUE : 0;
            if (found) term |= RAY_SURFACE | RAY_OPAQUE;  // non-shaded ray on a surface, see trace_kernel
            else if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
            else if (ray_t == s.tMax) term |= RAY_TIMEOUT;
            cls = classify_values(P, s.type, term, s.org.x, s.org.y, s.org.z, s.dir.x, s.dir.y, s.dir.z);
            if (cls == CLS_TERMINATED) {
              terminated = true;
              atomicAdd(fb + ((size_t)py * w + px), make_float4(s.r, s.g, s.b, 0.f));
            } else if (cls >= 0) do_spill = true;
          }
          // else: TIMEOUT (AO ray that reached ao_radius) or nothing: DROP_ON_FLOOR (Renderer.cpp:394-419)
#ifdef GXY_TRAV_COUNTERS
          atomicAdd(&q->nodes, (unsigned long long)st.n_nodes);
#endif
        }
        // the spill append, the statistics and the fetch of this refill are issued back to back by one lane
        const unsigned sm = __ballot_sync(m_idle, do_spill), tm = __ballot_sync(m_idle, terminated);
        const int leader = __ffs((int)m_idle) - 1;
        const unsigned cnt = (unsigned)__popc(m_idle);
        unsigned sbase = 0u, fbase = 0u;
        if ((int)lane == leader) {
          if (sm) sbase = atomicAdd(&q->n_spill, (unsigned)__popc(sm));
          if (tm) atomicAdd(&q->n_terminated, (unsigned long long)__popc(tm));
          if (!exhausted) fbase = atomicAdd(&q->sec_head, cnt);
        }
        sbase = __shfl_sync(m_idle, sbase, leader);
        fbase = __shfl_sync(m_idle, fbase, leader);
        if (PEER) {
          peer_push(T, parity_out, m_idle, lane, do_spill, cls, s.org, s.dir, s.r, s.g, s.b, 0.f, ray_t, s.tMax, px, py, s.type, term,
                    P.error_flag);
        } else if (do_spill) {
          const unsigned sp = sbase + (unsigned)__popc(sm & lt_mask);
          if (sp < spill_cap) write_spill(spill, sp, s.org, s.dir, s.r, s.g, s.b, 0.f, ray_t, s.tMax, px, py, s.type, term, cls);
          else *P.error_flag = 3;
        }
        pr.ray = -1;
        if (!exhausted) {
          const unsigned my = fbase + (unsigned)__popc(m_idle & lt_mask);
          ex_local = fbase + cnt >= n_queue;
          if (my < n_queue) {
            int hi, j;
            secondary_index(my, n_hits, L.n_ao, nsec - L.n_ao, hi, j);
            hi += (int)h0;
            const SecRay sr = make_secondary(L, hits, hi, j, epsilon, ao_tab[0], ao_tab[1], ao_tab[2]);
            trav = setup_ray_values(P, (int)my, false, sr.org, sr.dir, 0.f, sr.tMax, anyhit_secondary, rc, st, pr);
            pr.opaque = min3f(sr.r, sr.g, sr.b) >= 1.0f;
          }
        }
      }
      exhausted = exhausted || __any_sync(FULLMASK, ex_local);
      if (exhausted && __ballot_sync(FULLMASK, pr.ray >= 0) == 0u) break;
    }
    trace_iteration<CURVES>(P, rc, st, trav, pr.anyhit, stack, lstack, owner_of[warp], lane, lt_mask, prim_t);
  }
}

// ------------------------------------------------------------------------------------------------
// Wave >= 1 of the peer path: the rays other ranks wrote into this rank's inbox[parity_in] (PRIMARY rays
// that found nothing in their first partition, SHADOW/AO rays on their way to the light / out of the AO
// sphere).  Same persistent-warp traversal; per ray TraceRays_TraceRays restricted to geometry
// (TraceRays.ispc:377-441, 563-610), then Renderer::Classify (Renderer.cpp:304-454): a PRIMARY hit leaves a
// raw record for shade_hits_kernel<true>; everything else adds to the framebuffer, is dropped, or is
// written on into the next partition's inbox[parity_in ^ 1].
template <int FETCH_T, int MIN_BLOCKS, bool CURVES = false>
__global__ void __launch_bounds__(GXY_TRACE_THREADS, CURVES ? GXY_CURVE_BLOCKS : MIN_BLOCKS)
    inbox_trace_kernel(const __grid_constant__ SceneParams P, const __grid_constant__ PeerTable T, int parity_in, int w,
                       float4 *__restrict__ fb, unsigned *__restrict__ raw, unsigned raw_stride, FusedQueues *__restrict__ q,
                       int anyhit_secondary, int prim_t, int rays_per_thread) {
  __shared__ uint2 stack[GXY_STACK_SMEM * GXY_TRACE_THREADS];
  __shared__ unsigned char owner_of[GXY_TRACE_THREADS / 32][8];
  uint2 lstack[GXY_STACK_LOCAL];
  const float4 *__restrict__ inbox = peer_inbox(T, T.rank, parity_in);
  const unsigned n_in = peer_ctrl(T, T.rank)->inbox_count[parity_in];
  const unsigned n_queue = n_in < T.inbox_cap ? n_in : T.inbox_cap;
  if (n_queue == 0u || surplus_cta(n_queue, rays_per_thread)) return;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  RayCtx rc;
  TravState st;
  PendingRay pr;
  pr.ray = -1; pr.tExit = 0.f; pr.anyhit = false; pr.opaque = false;
  st.tg.y = 0u; st.ng.y = 0u;
  rc.org = f3(0.f, 0.f, 0.f); rc.dir = f3(1.f, 1.f, 1.f); rc.tnear = 0.f; rc.tfar = 0.f;
  st.best_t = 0.f; st.best_u = 0.f; st.best_v = 0.f; st.best_key = GXY_NO_HIT; st.best_rec = 0u;
  bool trav = false, exhausted = false;
  while (true) {
    const unsigned m_idle = __ballot_sync(FULLMASK, !trav);
    if (m_idle == FULLMASK || (!exhausted && __popc(m_idle) >= FETCH_T)) {
      bool ex_local = false;
      if (!trav) {
        bool has_hit = false, do_spill = false, terminated = false;
        float3 d0 = f3(0.f, 0.f, 0.f);
        float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
        float ray_t = 0.f, tMax = 0.f;
        int px = 0, py = 0, type = 0, term = 0, cls = CLS_UNDETERMINED;
        if (pr.ray >= 0) {
          const float4 a = inbox[4 * (size_t)pr.ray], b = inbox[4 * (size_t)pr.ray + 1], d = inbox[4 * (size_t)pr.ray + 3];
          col = inbox[4 * (size_t)pr.ray + 2];
          d0 = f3(a.w, b.x, b.y);
          tMax = b.w;
          px = __float_as_int(d.x); py = __float_as_int(d.y); type = __float_as_int(d.z);
          const bool found = st.best_key != GXY_NO_HIT;
          ray_t = found ? st.best_t : rc.tfar;
          term = (min3f(col.x, col.y, col.z) >= 1.0f || col.w > 0.999f) ? RAY_OPAQUE : 0;  // TraceRays.ispc:573
          if (type == RAY_PRIMARY && found) has_hit = true;  // shaded, lit and classified by shade_hits_kernel
          else {
            if (found) term |= RAY_SURFACE | RAY_OPAQUE;  // non-shaded ray on a surface, see trace_kernel
            else if (ray_t == pr.tExit) term |= RAY_BOUNDARY;
            else if (ray_t == tMax) term |= RAY_TIMEOUT;
            cls = classify_values(P, type, term, rc.org.x, rc.org.y, rc.org.z, d0.x, d0.y, d0.z);
            if (cls == CLS_TERMINATED) {
              terminated = true;
              if (col.x != 0.f || col.y != 0.f || col.z != 0.f || col.w != 0.f) atomicAdd(fb + ((size_t)py * w + px), col);
            } else if (cls >= 0) do_spill = true;
          }
#ifdef GXY_TRAV_COUNTERS
          atomicAdd(&q->nodes, (unsigned long long)st.n_nodes);
#endif
        }
        const unsigned hm = __ballot_sync(m_idle, has_hit), sm = __ballot_sync(m_idle, do_spill), tm = __ballot_sync(m_idle, terminated);
        const int leader = __ffs((int)m_idle) - 1;
        const unsigned cnt = (unsigned)__popc(m_idle);
        unsigned hbase = 0u, fbase = 0u;
        if ((int)lane == leader) {
          if (hm) hbase = atomicAdd(&q->n_hits, (unsigned)__popc(hm));
          if (sm) atomicAdd(&q->n_spill, (unsigned)__popc(sm));
          if (tm) atomicAdd(&q->n_terminated, (unsigned long long)__popc(tm));
          if (!exhausted) fbase = atomicAdd(&q->inbox_head, cnt);
        }
        hbase = __shfl_sync(m_idle, hbase, leader);
        fbase = __shfl_sync(m_idle, fbase, leader);
        if (has_hit) {
          const unsigned hp = hbase + (unsigned)__popc(hm & lt_mask);
          if (hp < raw_stride) {
            raw[hp] = (unsigned)pr.ray;
            raw[hp + raw_stride] = __float_as_uint(st.best_t);
            raw[hp + 2u * raw_stride] = __float_as_uint(st.best_u);
            raw[hp + 3u * raw_stride] = __float_as_uint(st.best_v);
            raw[hp + 4u * raw_stride] = st.best_key;
            raw[hp + 5u * raw_stride] = st.best_rec;
          } else *P.error_flag = 3;
        }
        peer_push(T, parity_in ^ 1, m_idle, lane, do_spill, cls, rc.org, d0, col.x, col.y, col.z, col.w, ray_t, tMax, px, py, type, term,
                  P.error_flag);
        pr.ray = -1;
        if (!exhausted) {
          const unsigned my = fbase + (unsigned)__popc(m_idle & lt_mask);
          ex_local = fbase + cnt >= n_queue;
          if (my < n_queue) {
            const float4 a = inbox[4 * (size_t)my], b = inbox[4 * (size_t)my + 1], d = inbox[4 * (size_t)my + 3];
            trav = setup_ray_values(P, (int)my, __float_as_int(d.z) == RAY_PRIMARY, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), b.z, b.w,
                                    anyhit_secondary, rc, st, pr);
          }
        }
      }
      exhausted = exhausted || __any_sync(FULLMASK, ex_local);
      if (exhausted && __ballot_sync(FULLMASK, pr.ray >= 0) == 0u) break;
    }
    trace_iteration<CURVES>(P, rc, st, trav, pr.anyhit, stack, lstack, owner_of[warp], lane, lt_mask, prim_t);
  }
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Flag barrier over all ranks (one warp): lane r publishes (work this rank leaves behind for a later wave, epoch) in rank r's arena and
// waits for rank r's flag in its own.  Everything the kernels before this one wrote into peer arenas is complete (stream order) before
// the flags go out.  Returns the sum of every rank's `sent` on every lane 0 (valid on lane 0).  All 32 lanes must call.
__device__ __forceinline__ unsigned peer_barrier(const PeerTable &T, unsigned epoch, unsigned sent, unsigned long long timeout_ns,
                                                 int *__restrict__ error_flag) {
  const unsigned lane = threadIdx.x & 31u;
  PeerCtrl *mine = peer_ctrl(T, T.rank);
  sent = __shfl_sync(FULLMASK, sent, 0);
  __threadfence_system();
  __syncwarp();
  unsigned theirs = 0u;
  if ((int)lane < T.nranks) {
    PeerCtrl *pc = peer_ctrl(T, (int)lane);
    *reinterpret_cast<volatile unsigned *>(&pc->pending[epoch & 1u][T.rank]) = sent;
    __threadfence_system();
    st_release_sys(&pc->flags[T.rank], epoch);
    const unsigned long long t0 = global_timer_ns();
    bool ok = true;
    while ((int)(ld_acquire_sys(&mine->flags[lane]) - epoch) < 0) {
      if (global_timer_ns() - t0 > timeout_ns) { ok = false; break; }
      __nanosleep(200);
    }
    if (!ok) *error_flag = 4;
    theirs = *reinterpret_cast<volatile unsigned *>(&mine->pending[epoch & 1u][lane]);
  }
  for (int off = 16; off > 0; off >>= 1) theirs += __shfl_down_sync(FULLMASK, theirs, off);
  return theirs;
}

// End of a wave on this rank (one warp): advance the queue bookkeeping, hand the consumed inbox back, then a
// flag barrier over all ranks -- lane r publishes (rays sent this wave, epoch) in rank r's arena and waits for
// rank r's flag in its own.  Everything the kernels before this one wrote into peer arenas is complete (stream
// order) before the flags go out.  After the barrier every rank holds the same global count of rays in flight.
// Replaces the per-wave count all-gather + host synchronisation of the NCCL path and, over a frame, the
// reference's busy/idle termination tree (RenderingSet.cpp:289-589).
__global__ void __launch_bounds__(32)
    wave_epilogue_kernel(const __grid_constant__ PeerTable T, FusedQueues *__restrict__ q, unsigned epoch, int parity_consumed,
                         int hits_spawn, unsigned long long timeout_ns, int *__restrict__ error_flag) {
  const unsigned lane = threadIdx.x;
  PeerCtrl *mine = peer_ctrl(T, T.rank);
  unsigned sent = 0u;
  if (lane == 0u) {
    // work this wave leaves behind for a later one: rays sent to other ranks, and surface hits whose AO/shadow rays
    // are generated in the next wave
    sent = q->n_spill - q->spill_done;
    if (hits_spawn) sent += q->n_hits - q->hits_done;
    q->spill_done = q->n_spill;
    q->sec_lo = q->sec_hi;
    q->sec_hi = q->n_hits;
    q->hits_done = q->n_hits;
    q->sec_head = 0u;
    q->inbox_head = 0u;
    if (parity_consumed >= 0) {
      q->n_inbox += (unsigned long long)mine->inbox_count[parity_consumed];
      mine->inbox_count[parity_consumed] = 0u;
    }
  }
  const unsigned theirs = peer_barrier(T, epoch, sent, timeout_ns, error_flag);
  if (lane == 0u) q->global_pending = theirs;
}

// Rendering::AddLocalPixels over all ranks (Rendering.cpp:125-153) for this rank's slice of the image: the sum of
// every rank's partial framebuffer (peer loads over NVLink, rank order) goes to the image owner's final buffer.
__global__ void __launch_bounds__(256) fb_gather_kernel(const __grid_constant__ PeerTable T) {
  const unsigned per = (T.npix + (unsigned)T.nranks - 1u) / (unsigned)T.nranks;
  const unsigned lo = per * (unsigned)T.rank, hi = min(T.npix, lo + per);
  float4 *__restrict__ dst = reinterpret_cast<float4 *>(T.base[0] + T.off_final);
  // all peer loads of a pixel are issued before the first is used (8 ranks: 8 x 16 bytes in flight per thread); the launch is one
  // CTA per SM on purpose: enough loads in flight to fill NVLink, while the SMs stay with the trace kernels of the frames behind
  for (unsigned i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
    float4 v[GXY_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < GXY_MAX_RANKS; r++)
      if (r < T.nranks) v[r] = __ldcs(reinterpret_cast<const float4 *>(T.base[r] + T.off_fb) + i);
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < GXY_MAX_RANKS; r++)
      if (r < T.nranks) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    dst[i] = acc;
  }
}


// ------------------------------------------------------------------------------------------------
// Frames in flight for Visualizations with volumes (one process per GPU): the list kernels of gxy_kernels.cu run on two
// device lists whose lengths never leave the device (VolQueues::cnt), rays that leave the brick travel as 64-byte inbox records --
// they carry what a ray needs to go on in the next brick: origin, direction, t, tMax, the colour and opacity accumulated so far,
// pixel, type -- and the flag barrier separates the waves, exactly as on the geometry path.
__global__ void vol_wave_counts_kernel(VolQueues *__restrict__ q, int cur, int n_ao, int n_sh, int first) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int n = q->cnt[cur];
  if (first) q->generated = (unsigned long long)n;
  q->traced += (unsigned long long)n;
  const long long nh = q->nhit;
  q->ao += (unsigned long long)(nh * n_ao);
  q->shadow += (unsigned long long)(nh * n_sh);
  q->cnt[cur ^ 1] = (int)(nh * (n_ao + n_sh));  // the spawned rays open the next list; keepers and the inbox are appended behind them
  q->away_wave = 0u;
  q->kept_wave = 0u;
}

// Renderer::SendRays for one list (Renderer.cpp:620-634): every ray classified for another rank becomes a record in that rank's
// inbox[parity]; a ray that stays (KEEP_HERE, or its own rank as destination) is appended to the next list
__global__ void __launch_bounds__(256)
    vol_forward_kernel(Rays R, int cap, Rays next, int next_cap, VolQueues *__restrict__ q, int cur, const __grid_constant__ PeerTable T, int parity,
                       int *__restrict__ error_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const int n = min(cap, q->cnt[cur]);
  int cls = CLS_TERMINATED;
  if (i < n) cls = R.classification[i];
  if (cls == CLS_KEEP_HERE) cls = T.rank;
  const bool away = cls >= 0 && cls != T.rank && cls < T.nranks;
  const bool keep = cls == T.rank;
  float3 org = f3(0.f, 0.f, 0.f), dir = f3(0.f, 0.f, 0.f);
  float r = 0.f, g = 0.f, b = 0.f, o = 0.f, t = 0.f, tMax = 0.f;
  int x = 0, y = 0, type = 0, term = 0;
  if (away || keep) {
    org = f3(R.ox[i], R.oy[i], R.oz[i]); dir = f3(R.dx[i], R.dy[i], R.dz[i]);
    r = R.r[i]; g = R.g[i]; b = R.b[i]; o = R.o[i]; t = R.t[i]; tMax = R.tMax[i];
    x = R.x[i]; y = R.y[i]; type = R.type[i]; term = R.term[i];
  }
  peer_push(T, parity, FULLMASK, lane, away, cls, org, dir, r, g, b, o, t, tMax, x, y, type, term, error_flag);
  if (keep) {
    const int pos = atomicAdd(&q->cnt[cur ^ 1], 1);
    if (pos < next_cap) {
      next.ox[pos] = org.x; next.oy[pos] = org.y; next.oz[pos] = org.z; next.dx[pos] = dir.x; next.dy[pos] = dir.y; next.dz[pos] = dir.z;
      next.r[pos] = r; next.g[pos] = g; next.b[pos] = b; next.o[pos] = o; next.t[pos] = t; next.tMax[pos] = tMax;
      next.x[pos] = x; next.y[pos] = y; next.type[pos] = type; next.term[pos] = term;
    } else *error_flag = 3;
  }
  const unsigned ma = __ballot_sync(FULLMASK, away), mk = __ballot_sync(FULLMASK, keep);
  if (lane == 0u) {
    if (ma) atomicAdd(&q->away_wave, (unsigned)__popc(ma));
    if (mk) atomicAdd(&q->kept_wave, (unsigned)__popc(mk));
  }
}

// end of a wave: what this rank leaves for a later wave = the records it pushed to other ranks + the rays of its next list so far
// (spawned AO/shadow rays and kept rays); barrier; the global sum is the termination test
__global__ void __launch_bounds__(32)
    vol_epilogue_kernel(const __grid_constant__ PeerTable T, VolQueues *__restrict__ q, int cur, unsigned epoch, unsigned long long timeout_ns,
                        int *__restrict__ error_flag) {
  const unsigned lane = threadIdx.x;
  unsigned sent = 0u;
  if (lane == 0u) {
    sent = q->away_wave + (unsigned)max(q->cnt[cur ^ 1], 0);
    q->forwarded += (unsigned long long)q->away_wave + (unsigned long long)q->kept_wave;
  }
  const unsigned theirs = peer_barrier(T, epoch, sent, timeout_ns, error_flag);
  if (lane == 0u) q->global_pending = theirs;
}

// after the barrier: the records the other ranks wrote into inbox[parity] during this wave are appended to the next list
__global__ void __launch_bounds__(256)
    vol_unpack_kernel(const __grid_constant__ PeerTable T, int parity, Rays next, int next_cap, const VolQueues *__restrict__ q, int cur,
                      int *__restrict__ error_flag) {
  const float4 *__restrict__ inbox = peer_inbox(T, T.rank, parity);
  const unsigned n_in = min(peer_ctrl(T, T.rank)->inbox_count[parity], T.inbox_cap);
  const int base = q->cnt[cur ^ 1];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += gridDim.x * blockDim.x) {
    const long long pos = (long long)base + i;
    if (pos >= next_cap) { *error_flag = 3; continue; }
    const float4 a = inbox[4 * (size_t)i], b = inbox[4 * (size_t)i + 1], c = inbox[4 * (size_t)i + 2], d = inbox[4 * (size_t)i + 3];
    next.ox[pos] = a.x; next.oy[pos] = a.y; next.oz[pos] = a.z; next.dx[pos] = a.w; next.dy[pos] = b.x; next.dz[pos] = b.y;
    next.t[pos] = b.z; next.tMax[pos] = b.w;
    next.r[pos] = c.x; next.g[pos] = c.y; next.b[pos] = c.z; next.o[pos] = c.w;
    next.x[pos] = __float_as_int(d.x); next.y[pos] = __float_as_int(d.y); next.type[pos] = __float_as_int(d.z); next.term[pos] = __float_as_int(d.w);
  }
}
// ... and the list's length and the inbox counter follow once every record is in place (next launch on the stream)
__global__ void vol_unpack_commit_kernel(const __grid_constant__ PeerTable T, int parity, VolQueues *__restrict__ q, int cur, int next_cap) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  PeerCtrl *mine = peer_ctrl(T, T.rank);
  const unsigned n_in = min(mine->inbox_count[parity], T.inbox_cap);
  q->inbox += (unsigned long long)n_in;
  const long long total = (long long)q->cnt[cur ^ 1] + n_in;
  q->cnt[cur ^ 1] = (int)min(total, (long long)next_cap);
  q->cnt[cur] = 0;
  q->away_wave = 0u;  // (consumed by the wave's epilogue; the rendezvous at the end of a frame must not count them again)
  q->kept_wave = 0u;
  mine->inbox_count[parity] = 0u;
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = GXY_SM_COUNT;
  }
  return sms;
}

static PeerTable no_peers() {
  PeerTable T;
  memset(&T, 0, sizeof T);
  return T;
}

// resident CTAs per SM a persistent trace launch asks for (8 = all the kernel is compiled for; the band pipeline of
// gxy_render may ask for fewer so that the kernels of several bands share the SMs from the start)
static unsigned fused_blocks_per_sm() {
  unsigned b = (unsigned)GXY_MIN_BLOCKS;
  if (const char *e = getenv("GXY_FUSED_BLOCKS_PER_SM")) b = (unsigned)std::max(1, std::min(GXY_MIN_BLOCKS, atoi(e)));
  return b;
}

// lanes of a warp that must hold a primitive group before the cooperative primitive passes run (trace_iteration)
static int prim_threshold() {
  int t = 1;
  if (const char *e = getenv("GXY_PRIM_T")) t = std::max(1, std::min(32, atoi(e)));
  return t;
}

// rays a queue must hold per thread of a CTA for that CTA to stay (surplus_cta); 0 = off
static int rays_per_thread() {
  int r = 4;
  if (const char *e = getenv("GXY_RAYS_PER_THREAD")) r = std::max(0, std::min(64, atoi(e)));
  return r;
}

static int fetch_threshold(const char *env) {
  int ft = 12;  // measured optimum on the 100M-triangle scene (tools/trace_sweep.py, GXY_FETCH_SWEEP)
  if (const char *e = getenv(env)) ft = atoi(e);
  return ft;
}

int launch_fused_primary(const SceneParams &P, const DevCamera &C, const DevLights &L, int w, int h, float *fb, Rays prim, unsigned *raw,
                         unsigned raw_stride, Rays hits, Rays spill, unsigned spill_cap, FusedQueues *q, float epsilon, const PeerTable *peer,
                         const PartProxy *proxies, int band, int n_bands, cudaStream_t st, const int *tile_rect) {
  if (ensure_ao_tables()) return 1;
  const int tiles_x = (w + 7) / 8, tiles_y = (h + 3) / 4;
  const int rows = (tiles_y - band + n_bands - 1) / n_bands;  // tile rows band, band + n_bands, ...
  if (rows <= 0) return 0;
  unsigned n_queue = (unsigned)tiles_x * (unsigned)rows * 32u;
  const PeerTable T = peer ? *peer : no_peers();
  const int skip_far = !(getenv("GXY_GEN_SKIP_FAR") && atoi(getenv("GXY_GEN_SKIP_FAR")) == 0);
  if (peer) {
    // generation only over the tiles this rank's box can project into (tile_rect = x0, y0, nx, ny in 8x4 tiles; NULL: all)
    int tx0 = 0, ty0 = 0, ntx = tiles_x, nty = tiles_y;
    if (tile_rect && skip_far) { tx0 = tile_rect[0]; ty0 = tile_rect[1]; ntx = tile_rect[2]; nty = tile_rect[3]; }
    const unsigned n_gen = (unsigned)ntx * (unsigned)nty * 32u;
    if (n_gen > 0u)
      gen_primary_peer_kernel<<<(n_gen + 255) / 256, 256, 0, st>>>(P, C, w, h, ntx, n_gen, prim, raw_stride, q, T.rank, T.nranks, proxies, skip_far,
                                                                    tx0, ty0);
  }
  else gen_primary_kernel<<<(n_queue + 255) / 256, 256, 0, st>>>(P, C, w, h, tiles_x, band, n_bands, n_queue, prim, spill, spill_cap, q);
  gxy_timeline_mark("gen", st);
  const unsigned needed = (n_queue + GXY_TRACE_THREADS - 1) / GXY_TRACE_THREADS;
  const bool curves = P.n_curves > 0;  // PathLines: the CURVES instantiations (round-Bezier test in the cooperative passes; 3 CTAs per SM)
  const unsigned blocks = std::min<unsigned>(needed, (unsigned)sm_count() * (curves ? (unsigned)GXY_CURVE_BLOCKS : fused_blocks_per_sm()));
  if (curves) {
    if (peer) primary_trace_kernel<12, GXY_MIN_BLOCKS, true, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, prim, raw, raw_stride, spill, spill_cap, q, T, prim_threshold(), rays_per_thread());
    else primary_trace_kernel<12, GXY_MIN_BLOCKS, false, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, prim, raw, raw_stride, spill, spill_cap, q, T, prim_threshold(), rays_per_thread());
    gxy_timeline_mark("primary", st);
    const unsigned sb = std::min<unsigned>((n_queue + 255) / 256, (unsigned)sm_count() * 4u);
    shade_hits_kernel<false, true><<<sb, 256, 0, st>>>(P, L, prim, nullptr, raw, raw_stride, w, reinterpret_cast<float4 *>(fb), hits, q, epsilon);
    GXY_CUDA(cudaGetLastError());
    return 0;
  }
#define GXY_LAUNCH_P(FT)                                                                                                                 \
  do {                                                                                                                                   \
    if (peer) primary_trace_kernel<FT, GXY_MIN_BLOCKS, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, prim, raw, raw_stride, spill, spill_cap, q, T, prim_threshold(), rays_per_thread()); \
    else primary_trace_kernel<FT, GXY_MIN_BLOCKS, false><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, prim, raw, raw_stride, spill, spill_cap, q, T, prim_threshold(), rays_per_thread());      \
  } while (0)
  switch (fetch_threshold("GXY_FETCH_P")) {
    case 4: GXY_LAUNCH_P(4); break;
    case 8: GXY_LAUNCH_P(8); break;
    case 16: GXY_LAUNCH_P(16); break;
    case 24: GXY_LAUNCH_P(24); break;
    default: GXY_LAUNCH_P(12); break;
  }
#undef GXY_LAUNCH_P
  gxy_timeline_mark("primary", st);
  // grid-stride over the hits the trace found (far fewer than pixels across ranks): no more CTAs than the device holds at once
  const unsigned shade_blocks = std::min<unsigned>((n_queue + 255) / 256, (unsigned)sm_count() * 8u);
  shade_hits_kernel<false><<<shade_blocks, 256, 0, st>>>(P, L, prim, nullptr, raw, raw_stride, w, reinterpret_cast<float4 *>(fb), hits, q, epsilon);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_fused_secondary(const SceneParams &P, const DevLights &L, int w, int h, int nsec, long long max_rays, float *fb, Rays hits,
                           Rays spill, unsigned spill_cap, FusedQueues *q, float epsilon, bool anyhit, const PeerTable *peer, int parity_out,
                           cudaStream_t st, int blocks_per_sm) {
  if (nsec <= 0 || max_rays <= 0) return 0;
  if (ensure_ao_tables()) return 1;
  const long long needed = (max_rays + GXY_TRACE_THREADS - 1) / GXY_TRACE_THREADS;
  const unsigned bps = blocks_per_sm > 0 ? (unsigned)std::min(8, blocks_per_sm) : fused_blocks_per_sm();
  const unsigned blocks = (unsigned)std::min<long long>(needed, (long long)sm_count() * bps);
  const PeerTable T = peer ? *peer : no_peers();
  if (P.n_curves > 0) {
    const unsigned cb = (unsigned)std::min<long long>(needed, (long long)sm_count() * std::min<unsigned>(bps, (unsigned)GXY_CURVE_BLOCKS));
    if (peer)
      fused_secondary_kernel<12, GXY_MIN_BLOCKS, true, true><<<cb, GXY_TRACE_THREADS, 0, st>>>(P, L, w, h, nsec, reinterpret_cast<float4 *>(fb), hits, spill, spill_cap, q, epsilon,
                                                                                             anyhit ? 1 : 0, T, parity_out, prim_threshold(), rays_per_thread());
    else
      fused_secondary_kernel<12, GXY_MIN_BLOCKS, false, true><<<cb, GXY_TRACE_THREADS, 0, st>>>(P, L, w, h, nsec, reinterpret_cast<float4 *>(fb), hits, spill, spill_cap, q, epsilon,
                                                                                              anyhit ? 1 : 0, T, parity_out, prim_threshold(), rays_per_thread());
    GXY_CUDA(cudaGetLastError());
    return 0;
  }
#define GXY_LAUNCH_S(FT)                                                                                                              \
  do {                                                                                                                                \
    if (peer)                                                                                                                         \
      fused_secondary_kernel<FT, GXY_MIN_BLOCKS, true><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, L, w, h, nsec, reinterpret_cast<float4 *>(fb), hits, spill, \
                                                                                spill_cap, q, epsilon, anyhit ? 1 : 0, T, parity_out, prim_threshold(), rays_per_thread());     \
    else                                                                                                                              \
      fused_secondary_kernel<FT, GXY_MIN_BLOCKS, false><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, L, w, h, nsec, reinterpret_cast<float4 *>(fb), hits, spill, \
                                                                                 spill_cap, q, epsilon, anyhit ? 1 : 0, T, parity_out, prim_threshold(), rays_per_thread());    \
  } while (0)
  switch (fetch_threshold("GXY_FETCH_S")) {
    case 4: GXY_LAUNCH_S(4); break;
    case 8: GXY_LAUNCH_S(8); break;
    case 16: GXY_LAUNCH_S(16); break;
    case 24: GXY_LAUNCH_S(24); break;
    default: GXY_LAUNCH_S(12); break;
  }
#undef GXY_LAUNCH_S
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_inbox_wave(const SceneParams &P, const DevLights &L, const PeerTable &T, int parity_in, int w, int h, float *fb, unsigned *raw,
                      unsigned raw_stride, Rays hits, FusedQueues *q, float epsilon, bool anyhit, cudaStream_t st, int blocks_per_sm) {
  (void)h;
  const unsigned bps = blocks_per_sm > 0 ? (unsigned)std::min(8, blocks_per_sm) : 8u;
  const unsigned blocks = (unsigned)sm_count() * bps;
  if (P.n_curves > 0) {
    const unsigned cb = (unsigned)sm_count() * std::min<unsigned>(bps, (unsigned)GXY_CURVE_BLOCKS);
    inbox_trace_kernel<12, GXY_MIN_BLOCKS, true><<<cb, GXY_TRACE_THREADS, 0, st>>>(P, T, parity_in, w, reinterpret_cast<float4 *>(fb), raw, raw_stride, q, anyhit ? 1 : 0,
                                                                                  prim_threshold(), rays_per_thread());
    gxy_timeline_mark("inbox", st);
    const float4 *inbox_c = reinterpret_cast<const float4 *>(T.base[T.rank] + T.off_inbox[parity_in]);
    shade_hits_kernel<true, true><<<(unsigned)sm_count() * std::min(4u, bps), 256, 0, st>>>(P, L, hits, inbox_c, raw, raw_stride, w, reinterpret_cast<float4 *>(fb), hits, q,
                                                                                           epsilon);
    GXY_CUDA(cudaGetLastError());
    return 0;
  }
  inbox_trace_kernel<12, GXY_MIN_BLOCKS><<<blocks, GXY_TRACE_THREADS, 0, st>>>(P, T, parity_in, w, reinterpret_cast<float4 *>(fb), raw, raw_stride, q,
                                                                   anyhit ? 1 : 0, prim_threshold(), rays_per_thread());
  gxy_timeline_mark("inbox", st);
  const float4 *inbox = reinterpret_cast<const float4 *>(T.base[T.rank] + T.off_inbox[parity_in]);
  shade_hits_kernel<true><<<(unsigned)sm_count() * std::min(4u, bps), 256, 0, st>>>(P, L, hits, inbox, raw, raw_stride, w, reinterpret_cast<float4 *>(fb),
                                                                                  hits, q, epsilon);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_wave_epilogue(const PeerTable &T, FusedQueues *q, unsigned epoch, int parity_consumed, bool hits_spawn, int *error_flag,
                         cudaStream_t st) {
  unsigned long long timeout_ns = 30ull * 1000000000ull;
  if (const char *e = getenv("GXY_PEER_TIMEOUT_MS")) timeout_ns = (unsigned long long)atoll(e) * 1000000ull;
  wave_epilogue_kernel<<<1, 32, 0, st>>>(T, q, epoch, parity_consumed, hits_spawn ? 1 : 0, timeout_ns, error_flag);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_vol_wave_counts(VolQueues *q, int cur, int n_ao, int n_sh, bool first, cudaStream_t st) {
  vol_wave_counts_kernel<<<1, 32, 0, st>>>(q, cur, n_ao, n_sh, first ? 1 : 0);
  GXY_CUDA(cudaGetLastError());
  return 0;
}
int launch_vol_forward(Rays R, int cap, Rays next, int next_cap, VolQueues *q, int cur, const PeerTable &T, int parity, int *error_flag,
                       cudaStream_t st) {
  vol_forward_kernel<<<(cap + 255) / 256, 256, 0, st>>>(R, cap, next, next_cap, q, cur, T, parity, error_flag);
  GXY_CUDA(cudaGetLastError());
  return 0;
}
int launch_vol_epilogue(const PeerTable &T, VolQueues *q, int cur, unsigned epoch, int *error_flag, cudaStream_t st) {
  unsigned long long timeout_ns = 30ull * 1000000000ull;
  if (const char *e = getenv("GXY_PEER_TIMEOUT_MS")) timeout_ns = (unsigned long long)atoll(e) * 1000000ull;
  vol_epilogue_kernel<<<1, 32, 0, st>>>(T, q, cur, epoch, timeout_ns, error_flag);
  GXY_CUDA(cudaGetLastError());
  return 0;
}
int launch_vol_unpack(const PeerTable &T, int parity, Rays next, int next_cap, VolQueues *q, int cur, int *error_flag, cudaStream_t st) {
  vol_unpack_kernel<<<(unsigned)sm_count() * 4u, 256, 0, st>>>(T, parity, next, next_cap, q, cur, error_flag);
  vol_unpack_commit_kernel<<<1, 32, 0, st>>>(T, parity, q, cur, next_cap);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_proxy_publish(const SceneParams &P, const PeerTable &T, cudaStream_t st) {
  proxy_publish_kernel<<<1, 32, 0, st>>>(P, reinterpret_cast<PartProxy *>(T.base[T.rank] + T.off_proxy));
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_proxy_gather(const PeerTable &T, PartProxy *out, cudaStream_t st) {
  proxy_gather_kernel<<<1, 256, 0, st>>>(T, out);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

int launch_fb_gather(const PeerTable &T, cudaStream_t st) {
  unsigned per_sm = 1u;
  if (const char *e = getenv("GXY_GATHER_BLOCKS")) per_sm = (unsigned)std::max(1, std::min(8, atoi(e)));
  fb_gather_kernel<<<(unsigned)sm_count() * per_sm, 256, 0, st>>>(T);
  GXY_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gxy
