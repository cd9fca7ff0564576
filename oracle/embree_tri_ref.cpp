/*
 * embree_tri_ref.cpp -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * A thin driver around the REFERENCE'S OWN ray/triangle arithmetic: it includes, from where they
 * lie under /root/reference, the vendored Embree 3.6.1 headers
 *     third-party/embree/kernels/geometry/triangle_intersector_moeller.h
 *       MoellerTrumboreIntersectorK<M,K>::intersectK  (:210-266, the packet path rtcIntersectV runs)
 *       MoellerTrumboreIntersector1<M>::intersect     (:62-111, the single-ray path)
 *       MoellerTrumboreHitK::operator()               (:188-200, t = T*rcp(absDen), rcp+Newton)
 * and calls them over ALL triangles of a mesh in primitive order (no BVH: Embree's BVH builder and
 * traversal need the whole library, which is built with cmake + generated headers and is therefore
 * treated as unbuildable here, DESIGN.md "oracle").  Nothing of Embree is copied into this repo:
 * the two configuration headers cmake would generate (kernels/config.h, rtcore_config.h) are
 * produced by oracle/Makefile from the reference's own *.h.in templates into oracle/_ref/gen/.
 *
 * Nearest-hit bookkeeping is Embree's: a candidate is accepted iff |den|*tnear < T <= |den|*tfar
 * with tfar already replaced by the t of the previously accepted hit (IntersectKEpilogM,
 * intersector_epilog.h:573-640), so among exact ties the LATER primitive wins, here "later in
 * primitive order".
 *
 * Built only where /root/reference exists; the .so lands in oracle/_ref/ (git-ignored, shipped to
 * the GPU box).  Used by tests/test_oracle_embree.py to pin oracle/gxy_oracle.cpp's triangle test.
 */
#include "kernels/geometry/triangle_intersector_moeller.h"

#include <cstring>
#include <thread>
#include <vector>

using namespace embree;
using namespace embree::isa;

namespace {

template <int K>
struct NearestEpilogK {  // what IntersectKEpilogM does without filters / masks
  vfloat<K> &tfar, &u, &v, &t;
  vint<K> &prim;
  Vec3vf<K> &Ng;
  int primID;
  NearestEpilogK(vfloat<K> &tfar, vfloat<K> &u, vfloat<K> &v, vfloat<K> &t, vint<K> &prim, Vec3vf<K> &Ng, int primID)
      : tfar(tfar), u(u), v(v), t(t), prim(prim), Ng(Ng), primID(primID) {}
  template <typename Hit>
  __forceinline vbool<K> operator()(const vbool<K> &valid, const Hit &hit) const {
    vfloat<K> hu, hv, ht;
    Vec3vf<K> hNg;
    std::tie(hu, hv, ht, hNg) = hit();
    tfar = select(valid, ht, tfar);
    t = select(valid, ht, t);
    u = select(valid, hu, u);
    v = select(valid, hv, v);
    prim = select(valid, vint<K>(primID), prim);
    Ng.x = select(valid, hNg.x, Ng.x);
    Ng.y = select(valid, hNg.y, Ng.y);
    Ng.z = select(valid, hNg.z, Ng.z);
    return valid;
  }
};

template <int K>
void packet_range(int nt, const float *verts, const int *idx, int r0, int r1, const float *org3, const float *dir3, const float *tnear,
                  const float *tfar, int *prim_out, float *tuv_out, float *ng_out) {
  MoellerTrumboreIntersectorK<4, K> isec(vbool<K>(true), *(RayK<K> *)nullptr);
  for (int base = r0; base < r1; base += K) {
    float o[3][K], d[3][K], tn[K], tf[K];
    int act[K];
    for (int l = 0; l < K; l++) {
      const int r = base + l < r1 ? base + l : r1 - 1;
      act[l] = base + l < r1 ? -1 : 0;
      for (int a = 0; a < 3; a++) { o[a][l] = org3[3 * r + a]; d[a][l] = dir3[3 * r + a]; }
      tn[l] = tnear[r]; tf[l] = tfar[r];
    }
    const Vec3vf<K> O(vfloat<K>::loadu(o[0]), vfloat<K>::loadu(o[1]), vfloat<K>::loadu(o[2]));
    const Vec3vf<K> D(vfloat<K>::loadu(d[0]), vfloat<K>::loadu(d[1]), vfloat<K>::loadu(d[2]));
    const vfloat<K> TN = vfloat<K>::loadu(tn);
    vfloat<K> TF = vfloat<K>::loadu(tf);
    const vbool<K> valid0 = vint<K>::loadu(act) != vint<K>(0);
    vfloat<K> u(0.f), v(0.f), t = TF;
    vint<K> prim(-1);
    Vec3vf<K> Ng(vfloat<K>(0.f), vfloat<K>(0.f), vfloat<K>(0.f));
    for (int p = 0; p < nt; p++) {
      const float *a = verts + 3 * (size_t)idx[3 * p], *b = verts + 3 * (size_t)idx[3 * p + 1], *c = verts + 3 * (size_t)idx[3 * p + 2];
      // embree/kernels/geometry/triangle.h:52-53 (TriangleM ctor): e1 = v0 - v1, e2 = v2 - v0; Ng = cross(e2, e1) (:133-136)
      const Vec3fa v0(a[0], a[1], a[2]), v1(b[0], b[1], b[2]), v2(c[0], c[1], c[2]);
      const Vec3fa e1 = v0 - v1, e2 = v2 - v0;
      const Vec3vf<K> tv0(vfloat<K>(v0.x), vfloat<K>(v0.y), vfloat<K>(v0.z));
      const Vec3vf<K> te1(vfloat<K>(e1.x), vfloat<K>(e1.y), vfloat<K>(e1.z));
      const Vec3vf<K> te2(vfloat<K>(e2.x), vfloat<K>(e2.y), vfloat<K>(e2.z));
      const Vec3vf<K> tNg = cross(te2, te1);
      isec.intersectK(valid0, O, D, TN, TF, tv0, te1, te2, tNg, NearestEpilogK<K>(TF, u, v, t, prim, Ng, p));
    }
    for (int l = 0; l < K && base + l < r1; l++) {
      const int r = base + l;
      prim_out[r] = prim[l];
      tuv_out[3 * r] = t[l]; tuv_out[3 * r + 1] = u[l]; tuv_out[3 * r + 2] = v[l];
      if (ng_out) { ng_out[3 * r] = Ng.x[l]; ng_out[3 * r + 1] = Ng.y[l]; ng_out[3 * r + 2] = Ng.z[l]; }
    }
  }
}

}  // namespace

extern "C" {

const char *gxr_describe(void) {
  return "Embree 3.6.1 MoellerTrumboreIntersectorK<4,8>::intersectK (AVX2+FMA), brute force in primitive order";
}

/* nearest hit of each ray over all triangles, Embree's packet arithmetic (K = 8, the AVX2 width).
 * prim_out[n] (-1 = miss), tuv_out[3n] = (t,u,v) (t = tfar on a miss), ng_out[3n] may be NULL. */
int gxr_intersect_packet8(int n_tris, const float *verts, const int *idx, int n_rays, const float *org3, const float *dir3,
                          const float *tnear, const float *tfar, int *prim_out, float *tuv_out, float *ng_out, int nthreads) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  const int chunk = ((n_rays + nthreads - 1) / nthreads + 7) & ~7;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    const int r0 = t * chunk, r1 = std::min(n_rays, r0 + chunk);
    if (r0 >= r1) break;
    th.emplace_back([=]() { packet_range<8>(n_tris, verts, idx, r0, r1, org3, dir3, tnear, tfar, prim_out, tuv_out, ng_out); });
  }
  for (auto &t : th) t.join();
  return 0;
}

/* single-ray path (MoellerTrumboreIntersector1<4>, Triangle4 blocks in primitive order; inside a
 * block the first lane of minimal t wins as select_min does, Intersect1EpilogM :242-248). */
int gxr_intersect_single(int n_tris, const float *verts, const int *idx, int n_rays, const float *org3, const float *dir3,
                         const float *tnear, const float *tfar, int *prim_out, float *tuv_out) {
  MoellerTrumboreIntersector1<4> isec;
  for (int r = 0; r < n_rays; r++) {
    Ray ray(Vec3fa(org3[3 * r], org3[3 * r + 1], org3[3 * r + 2]), Vec3fa(dir3[3 * r], dir3[3 * r + 1], dir3[3 * r + 2]), tnear[r], tfar[r]);
    int best = -1;
    float bu = 0.f, bv = 0.f;
    for (int p0 = 0; p0 < n_tris; p0 += 4) {
      float x[9][4];
      int ok[4];
      for (int l = 0; l < 4; l++) {
        const int p = p0 + l < n_tris ? p0 + l : n_tris - 1;
        ok[l] = p0 + l < n_tris ? -1 : 0;
        for (int k = 0; k < 3; k++)
          for (int a = 0; a < 3; a++) x[3 * k + a][l] = verts[3 * (size_t)idx[3 * p + k] + a];
      }
      const Vec3vf<4> v0(vfloat4::loadu(x[0]), vfloat4::loadu(x[1]), vfloat4::loadu(x[2]));
      const Vec3vf<4> v1(vfloat4::loadu(x[3]), vfloat4::loadu(x[4]), vfloat4::loadu(x[5]));
      const Vec3vf<4> v2(vfloat4::loadu(x[6]), vfloat4::loadu(x[7]), vfloat4::loadu(x[8]));
      const Vec3vf<4> e1 = v0 - v1, e2 = v2 - v0;
      const Vec3vf<4> Ng = cross(e2, e1);
      MoellerTrumboreHitM<4> hit;
      const vbool4 valid0 = vint4::loadu(ok) != vint4(0);
      if (isec.intersect(valid0, ray, v0, e1, e2, Ng, hit)) {
        hit.finalize();
        const size_t i = select_min(hit.valid, hit.vt);
        ray.tfar = hit.vt[i];
        bu = hit.vu[i]; bv = hit.vv[i];
        best = p0 + (int)i;
      }
    }
    prim_out[r] = best;
    tuv_out[3 * r] = ray.tfar; tuv_out[3 * r + 1] = bu; tuv_out[3 * r + 2] = bv;
  }
  return 0;
}
}
