/*
 * embree_curve_ref.cpp -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * A thin driver around the REFERENCE'S OWN ray / round-Bezier-curve arithmetic.  Galaxy hands its
 * PathLines to Embree as RTC_GEOMETRY_TYPE_ROUND_BEZIER_CURVE (src/ospray/DataDrivenPathLines.ispc:
 * 319-324); Embree 3.6.1 intersects those with SweepCurve1Intersector1 / SweepCurve1IntersectorK
 * <BezierCurveT<Vec3fa>> (kernels/geometry/curve_intersector_virtual.cpp:256-266), both of which run
 * intersect_bezier_recursive_jacobian ray by ray (kernels/geometry/curve_intersector_sweep.h:122-221).
 * This file includes that header from where it lies under /root/reference and calls
 *     SweepCurve1Intersector1<BezierCurve3fa>::intersect       (:224-241)
 * over ALL curve segments in primitive order (no BVH, as embree_tri_ref.cpp does for triangles),
 * with an epilog that does what Intersect1Epilog1 does without filters and masks
 * (intersector_epilog.h:72-80: tfar = t, Ng, u, v, primID).  Nothing of Embree is copied.
 *
 * Built only where /root/reference exists (make -C oracle ref); the .so lands in oracle/_ref/.
 * Used by tests/test_oracle_curves.py to pin oracle/gxy_oracle.cpp's curve test.
 */
#include "kernels/subdiv/bezier_curve.h"
#include "kernels/geometry/curve_intersector_sweep.h"

using namespace embree;
using namespace embree::isa;

namespace {
struct NearestEpilog1 {
  RayHit &ray;
  unsigned primID;
  NearestEpilog1(RayHit &ray, unsigned primID) : ray(ray), primID(primID) {}
  template <typename Hit>
  __forceinline bool operator()(Hit &hit) const {
    hit.finalize();
    ray.tfar = hit.t;
    ray.Ng = hit.Ng;
    ray.u = hit.u;
    ray.v = hit.v;
    ray.primID = primID;
    ray.geomID = 0;
    return true;
  }
};
}  // namespace

extern "C" {

const char *gxr_curve_describe(void) {
  return "Embree 3.6.1 SweepCurve1Intersector1<BezierCurve3fa> (AVX2+FMA, VSIZEX=8, 2 subdivisions), brute force in primitive order";
}

/* nearest hit of each ray over all n_curves cubic Bezier segments (cp = 4 control points x (x,y,z,r) per
 * segment).  prim_out[n] (-1 = miss), tu_out[2n] = (t,u) (t = tfar on a miss), ng_out[3n].
 * per_curve != 0: every (ray, curve) pair is tested on its own against the ORIGINAL interval and the
 * outputs are n_rays x n_curves (prim_out = 0/1 hit flag). */
int gxr_curve_intersect(int n_curves, const float *cp, int n_rays, const float *org3, const float *dir3, const float *tnear,
                        const float *tfar, int *prim_out, float *tu_out, float *ng_out, int per_curve) {
  SweepCurve1Intersector1<BezierCurve3fa> isec;
  for (int r = 0; r < n_rays; r++) {
    const Vec3fa org(org3[3 * r], org3[3 * r + 1], org3[3 * r + 2]);
    const Vec3fa dir(dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]);
    RayHit ray(org, dir, tnear[r], tfar[r]);
    ray.primID = (unsigned)-1;
    ray.u = ray.v = 0.f;
    ray.Ng = Vec3fa(0.f);
    CurvePrecalculations1 pre(ray, nullptr);
    for (int p = 0; p < n_curves; p++) {
      const float *c = cp + 16 * (size_t)p;
      Vec3fa v0(c[0], c[1], c[2]), v1(c[4], c[5], c[6]), v2(c[8], c[9], c[10]), v3(c[12], c[13], c[14]);
      v0.w = c[3]; v1.w = c[7]; v2.w = c[11]; v3.w = c[15];
      if (per_curve) {
        RayHit one(org, dir, tnear[r], tfar[r]);
        one.primID = (unsigned)-1; one.u = one.v = 0.f; one.Ng = Vec3fa(0.f);
        const bool h = isec.intersect(pre, one, nullptr, (unsigned)p, v0, v1, v2, v3, NearestEpilog1(one, (unsigned)p));
        const size_t o = (size_t)r * n_curves + p;
        prim_out[o] = h ? 1 : 0;
        tu_out[2 * o] = one.tfar; tu_out[2 * o + 1] = one.u;
        if (ng_out) { ng_out[3 * o] = one.Ng.x; ng_out[3 * o + 1] = one.Ng.y; ng_out[3 * o + 2] = one.Ng.z; }
      } else {
        isec.intersect(pre, ray, nullptr, (unsigned)p, v0, v1, v2, v3, NearestEpilog1(ray, (unsigned)p));
      }
    }
    if (!per_curve) {
      prim_out[r] = (int)ray.primID;
      tu_out[2 * r] = ray.tfar; tu_out[2 * r + 1] = ray.u;
      if (ng_out) { ng_out[3 * r] = ray.Ng.x; ng_out[3 * r + 1] = ray.Ng.y; ng_out[3 * r + 2] = ray.Ng.z; }
    }
  }
  return 0;
}
}
