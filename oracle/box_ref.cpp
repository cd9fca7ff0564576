// box_ref.cpp (ours) -- C entry points over the reference's own gxy::Box (src/data/Box.cpp, compiled from where it lies under
// /root/reference by `make -C oracle ref`; no reference source is copied).  Used by tests/test_oracle_box.py to pin the oracle's
// restatements of Box::exit_face / Box::intersect / Box(origin, counts, deltas) bit for bit.  Test infrastructure only.
#include "Box.h"

extern "C" {

// boxes6: n x (minx miny minz maxx maxy maxz); rays6: n x (x y z dx dy dz)
void gxref_exit_face(int n, const float *boxes6, const float *rays6, int *faces) {
  for (int i = 0; i < n; i++) {
    const float *b = boxes6 + 6 * i, *r = rays6 + 6 * i;
    gxy::Box box(b[0], b[1], b[2], b[3], b[4], b[5]);
    faces[i] = box.exit_face(r[0], r[1], r[2], r[3], r[4], r[5]);
  }
}

// hit[i] = Box::intersect(o, d, tmin, tmax); t2 = n x (tmin tmax) as left by the call
void gxref_box_intersect(int n, const float *boxes6, const float *rays6, int *hit, float *t2) {
  for (int i = 0; i < n; i++) {
    const float *b = boxes6 + 6 * i, *r = rays6 + 6 * i;
    gxy::Box box(b[0], b[1], b[2], b[3], b[4], b[5]);
    gxy::vec3f o, d;
    o.x = r[0]; o.y = r[1]; o.z = r[2];
    d.x = r[3]; d.y = r[4]; d.z = r[5];
    float tmin = 0.f, tmax = 0.f;
    hit[i] = box.intersect(o, d, tmin, tmax) ? 1 : 0;
    t2[2 * i] = tmin; t2[2 * i + 1] = tmax;
  }
}

// Box(float *o, int *n, float *d): min = o, max = o + (n-1)*d  (Box.cpp:69-80)
void gxref_box_from_grid(const float *origin3, const int *counts3, const float *deltas3, float *minmax6) {
  float o[3] = {origin3[0], origin3[1], origin3[2]}, d[3] = {deltas3[0], deltas3[1], deltas3[2]};
  int n[3] = {counts3[0], counts3[1], counts3[2]};
  gxy::Box box(o, n, d);
  minmax6[0] = box.xyz_min.x; minmax6[1] = box.xyz_min.y; minmax6[2] = box.xyz_min.z;
  minmax6[3] = box.xyz_max.x; minmax6[4] = box.xyz_max.y; minmax6[5] = box.xyz_max.z;
}
}
