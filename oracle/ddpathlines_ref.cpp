/*
 * ddpathlines_ref.cpp -- ORACLE SUPPORT (test infrastructure, NOT product code).
 *
 * Drives the REFERENCE'S OWN curve builder, ospray::DataDrivenPathLines::finalize (src/ospray/DataDrivenPathLines.cpp:50-168),
 * compiled from where it lies under /root/reference together with the two OSPRay base-class sources it needs
 * (ospray/common/Managed.cpp, components/ospcommon/utility/ParameterizedObject.cpp).  finalize() reads its
 * parameters ("vertices", "data", "indices", "radius0/1", "value0/1", "transferFunction"), builds vertexCurve /
 * indexCurve and hands them to the ISPC side through ispc::DataDrivenPathLines_setCurve -- the ISPC export is
 * where this file sits: its stand-in below captures exactly what Embree would be given and expands it to the 4
 * control points per segment Embree gathers (4 consecutive vertices from indexCurve[i]).
 *
 * What is OURS here (stubs, no reference code copied): the ISPC exports (create / setCurve / delete_uniform), the
 * handful of ospray::Geometry / ospray::Data members and ospray::postStatusMsg that the reference objects link
 * against (OSPRay's own Geometry.cpp, Data.cpp and OSPCommon.cpp pull in the device, the ISPC runtime and
 * Embree's scene: not needed to run finalize()).  The two headers a build would generate (OSPConfig.h from its
 * .h.in template, the *_ispc.h export headers) are produced by oracle/Makefile into oracle/_ref/gen.
 *
 * Built only where /root/reference exists (make -C oracle ref) into oracle/_ref/libgxy_ddpathlines_ref.so; used by
 * tests/test_oracle_curves.py to pin gxo_build_curves (and through it the product's gxy_build_curves) bit for bit.
 */
#include <cstring>
#include <vector>

#include "DataDrivenPathLines.h"
#include "common/Data.h"

namespace {
std::vector<float> g_vertices;          // vertexCurve as handed to Embree (x,y,z,r per vertex)
std::vector<unsigned> g_indices;        // indexCurve
float g_params[4];
}

// ---- the ISPC side (stand-ins for the generated exports) ------------------------------------------------------------
namespace ispc {
extern "C" {
void *DataDrivenPathLines_create(void *) { return nullptr; }
void *DataDrivenPathLines_setCurve(void *, void *, const float *vertices, int32_t numVertices, const uint32_t *indices, int32_t numSegments,
                                   void *, float r0, float r1, float v0, float v1) {
  g_vertices.assign(vertices, vertices + 4 * (size_t)numVertices);
  g_indices.assign(indices, indices + numSegments);
  g_params[0] = r0; g_params[1] = r1; g_params[2] = v0; g_params[3] = v1;
  return nullptr;
}
void delete_uniform(void *) {}
}
}  // namespace ispc

// ---- the OSPRay pieces DataDrivenPathLines.o / Managed.o link against, reduced to what finalize() touches ------------
namespace ospray {
Geometry::Geometry() { managedObjectType = OSP_GEOMETRY; }
void Geometry::setMaterial(Material *) {}
void Geometry::setMaterialList(Data *) {}
Material *Geometry::getMaterial() const { return nullptr; }
std::string Geometry::toString() const { return "ospray::Geometry"; }
void Geometry::finalize(Model *) {}
Data::Data(size_t n, OSPDataType t, const void *init, int f) : data(const_cast<void *>(init)), numItems(n), numBytes(0), flags(f), type(t) {
  managedObjectType = OSP_DATA;
}
Data::~Data() {}
void Data::commit() {}
std::string Data::toString() const { return "ospray::Data"; }
void postStatusMsg(const std::string &, uint32_t) {}
void postStatusMsg(const std::stringstream &, uint32_t) {}
StatusMsgStream postStatusMsg(uint32_t level) { return StatusMsgStream(level); }
}  // namespace ospray

extern "C" {

const char *gxr_ddpathlines_describe(void) {
  return "ospray::DataDrivenPathLines::finalize (src/ospray/DataDrivenPathLines.cpp) compiled from the reference tree";
}

/* cp_out: n_segments x 4 x (x,y,z,r).  Returns 0, or -1 if finalize() threw. */
int gxr_build_curves(int n_verts, const float *verts, const float *data, int n_segments, const int *connectivity, float radius0, float radius1,
                     float value0, float value1, float *cp_out) {
  using namespace ospray;
  try {
    DataDrivenPathLines *g = new DataDrivenPathLines();
    Data *v = new Data((size_t)n_verts, OSP_FLOAT3, verts, OSP_DATA_SHARED_BUFFER);
    Data *d = new Data((size_t)n_verts, OSP_FLOAT, data, OSP_DATA_SHARED_BUFFER);
    Data *i = new Data((size_t)n_segments, OSP_INT, connectivity, OSP_DATA_SHARED_BUFFER);
    Data *tf = new Data(0, OSP_FLOAT, nullptr, 0);   // any managed object: finalize() only asks it for its (null) ISPC handle
    g->setParam<ManagedObject *>("vertices", v);
    g->setParam<ManagedObject *>("data", d);
    g->setParam<ManagedObject *>("indices", i);
    g->setParam<ManagedObject *>("transferFunction", tf);
    g->setParam<float>("radius", 0.1f);              // OsprayPathLines.cpp:43-44
    g->setParam<float>("radius0", radius0);          // PathLinesVis::SetTheOsprayDataObject (PathLinesVis.cpp:133-144)
    g->setParam<float>("radius1", radius1);
    g->setParam<float>("value0", value0);
    g->setParam<float>("value1", value1);
    g_vertices.clear(); g_indices.clear();
    g->finalize(reinterpret_cast<Model *>(tf));      // a Model is only asked for its ISPC handle as well
    if ((int)g_indices.size() != n_segments) return -1;
    for (int s = 0; s < n_segments; s++) memcpy(cp_out + 16 * (size_t)s, &g_vertices[4 * (size_t)g_indices[s]], 16 * sizeof(float));
    return 0;                                         // (the five small objects are left to the process: test driver)
  } catch (...) {
    return -1;
  }
}
}
