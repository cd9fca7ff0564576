// TEST HARNESS (not product code): compiles the product's device source galaxy_b200/csrc/gxy_curve.cuh as plain C++
// so that the arithmetic the GPU kernels run can be compared with the oracle on a machine without a GPU
// (tests/test_curve_host.py builds this with g++ -ffp-contract=off into a temporary directory).
#include "../galaxy_b200/csrc/gxy_curve.cuh"

static long g_culled = 0, g_tested = 0;
extern "C" void gxc_cull_stats(long *culled, long *tested) { *culled = g_culled; *tested = g_tested; }

extern "C" int gxc_curve_intersect(int n_curves, const float *cp, int n_rays, const float *org3, const float *dir3, const float *tnear,
                                   const float *tfar, int *prim_out, float *tu_out, float *ng_out, int per_curve) {
  for (int r = 0; r < n_rays; r++) {
    int best = -1;
    float bt = tfar[r], bu = 0.f, bn[3] = {0.f, 0.f, 0.f};
    for (int p = 0; p < n_curves; p++) {
      gxc::CurveHit h;
      // as the device does (curve_rec_test, gxy_traverse.cuh): the conservative cull with the bound the BVH builder stores, then the test
      const float *c = cp + 16 * (long)p;
      const bool culled = gxc::curve_precull(gxc::v3(c[0], c[1], c[2]), gxc::v3(c[12], c[13], c[14]), gxc::curve_bound_radius(c),
                                             gxc::v3(org3[3 * r], org3[3 * r + 1], org3[3 * r + 2]), gxc::v3(dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]));
      g_culled += culled ? 1 : 0;
      g_tested += 1;
      const bool hit = !culled && gxc::curve_test(c, org3[3 * r], org3[3 * r + 1], org3[3 * r + 2], dir3[3 * r], dir3[3 * r + 1],
                                                  dir3[3 * r + 2], tnear[r], tfar[r], h);
      if (per_curve) {
        const long o = (long)r * n_curves + p;
        prim_out[o] = hit ? 1 : 0;
        tu_out[2 * o] = hit ? h.t : tfar[r]; tu_out[2 * o + 1] = hit ? h.u : 0.f;
        if (ng_out) { ng_out[3 * o] = hit ? h.Ng.x : 0.f; ng_out[3 * o + 1] = hit ? h.Ng.y : 0.f; ng_out[3 * o + 2] = hit ? h.Ng.z : 0.f; }
      } else if (hit && (best < 0 || h.t < bt)) {
        best = p; bt = h.t; bu = h.u; bn[0] = h.Ng.x; bn[1] = h.Ng.y; bn[2] = h.Ng.z;
      }
    }
    if (!per_curve) {
      prim_out[r] = best; tu_out[2 * r] = bt; tu_out[2 * r + 1] = bu;
      if (ng_out) { ng_out[3 * r] = bn[0]; ng_out[3 * r + 1] = bn[1]; ng_out[3 * r + 2] = bn[2]; }
    }
  }
  return 0;
}
