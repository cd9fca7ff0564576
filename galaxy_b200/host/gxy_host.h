// gxy_host.h -- C++ host side above the C ABI (include/gxy_gpu.h): the state-file-driven objects of the
// reference, with the reference's names, JSON keys and error behaviour (bool returns + a message on
// stderr), so that an existing .state file renders unchanged.  Everything that touches a ray goes through
// the C ABI; nothing here computes on the hot path.
//
//   reference class (file)                                   here
//   Datasets (src/data/Datasets.cpp:96-155)                  gxy::Datasets
//   Volume   (src/data/Volume.cpp:88-172,195-392)            gxy::Volume        (.vol / .json headers + raw)
//   Camera   (src/renderer/Camera.cpp:165-281)               gxy::Camera
//   Lighting (src/renderer/Lighting.cpp:59-115)              gxy::Lighting
//   Vis / MappedVis / VolumeVis (Vis.cpp:126-146,            gxy::Vis (one struct, `type` says which)
//     MappedVis.cpp:86-203, VolumeVis.cpp:119-163)
//   Visualization (src/renderer/Visualization.cpp:139-340)   gxy::Visualization  (Commit -> one gxy_vis per partition)
//   Rendering / RenderingSet (Rendering.cpp:91-293)          gxy::Rendering      (Render, SaveImage)
//   Renderer (src/renderer/Renderer.cpp:125,289-296)         gxy::Renderer       (epsilon)
//   ColorImageWriter / write_png (ImageWriter.cpp:30-67)     gxy::write_png      (zlib; libpng is not needed)
//
//   Geometry / Triangles / Particles / PathLines (src/data/Geometry.cpp:176-344,  gxy::Geometry  (partition document + .vtu/.vtp pieces
//     Triangles.cpp:79-146, Particles.cpp:90-129, PathLines.cpp:73-125)               through the VTK-free reader of gxy_vtu.h)
//   PathLinesVis (src/renderer/PathLinesVis.cpp:46-55,105-114)                    gxy::Vis with type "PathLinesVis"
#pragma once
#include <string>
#include <vector>

#include "../../include/gxy_gpu.h"
#include "gxy_json.h"

namespace gxy {

struct vec3i { int x, y, z; };

// src/data/Volume.cpp:124-172
struct VolumePart {
  int ijk[3], offsets[3], counts[3], goffsets[3], gcounts[3];
};

class Volume {
 public:
  bool Import(const std::string &filename);  // header only (Volume::local_import up to the raw read)
  // ghosted brick of partition `rank` of `nparts`, rows read from the raw file exactly as Volume.cpp:333-348
  bool LoadBrick(int nparts, int rank, std::vector<unsigned char> &samples, VolumePart &part) const;
  // global / local Box of a partition in fp32 (Volume.cpp:379-390; Box.cpp:69-80)
  void Boxes(const VolumePart &part, float gmin[3], float gmax[3], float lmin[3], float lmax[3]) const;
  static void Neighbors(const VolumePart &part, const int factors[3], int neighbors[6]);  // Volume.cpp:358-377

  std::string name, filename, raw_filename;
  bool is_float = true;
  float origin[3] = {0, 0, 0}, deltas[3] = {1, 1, 1};
  int counts[3] = {0, 0, 0};
  int number_of_components = 1;
};

// One partition of a geometry dataset as the reference holds it after load_from_vtkPointSet
struct GeometryPiece {
  std::vector<float> vertices, normals, data;  // 3, 3, 1 per vertex
  std::vector<int> connectivity;               // Triangles: 3 per triangle; PathLines: first vertex of every segment
};

// Geometry::local_import / get_partitioning (src/data/Geometry.cpp:176-344): the dataset file is a partition document
// {"parts": [{"filename": "piece.vtu", "extent": [x0,x1,y0,y1,z0,z1]}, ...]} with one part per rank
class Geometry {
 public:
  bool Import(const std::string &filename);
  int NumberOfParts() const { return (int)part_files.size(); }
  bool LoadPiece(int rank, GeometryPiece &out) const;  // Triangles / Particles / PathLines ::load_from_vtkPointSet
  void Boxes(int rank, float gmin[3], float gmax[3], float lmin[3], float lmax[3], int neighbors[6]) const;
  std::string name, type, filename;                    // type: "Triangles" | "Particles" | "PathLines"
  std::vector<std::string> part_files;
  std::vector<float> extents;                          // 6 per part, stored as float like the reference
};

class Datasets {
 public:
  bool LoadFromJSON(const json::Value &doc, const std::string &state_dir);
  const Volume *FindVolume(const std::string &name) const;
  const Geometry *FindGeometry(const std::string &name) const;
  std::vector<Volume> volumes;
  std::vector<Geometry> geometries;
};

class Camera {
 public:
  // v: a camera object, or the name of a ParaView camera file, in JSON form or as ParaView writes it (.pvcc, XML) (Camera.cpp:118-232)
  bool LoadFromJSON(const json::Value &v, const std::string &base_dir = "");
  bool LoadFromPVCC(const std::string &filename);
  static bool LoadCamerasFromJSON(const json::Value &doc, std::vector<Camera> &out, const std::string &base_dir = "");
  gxy_camera AsABI() const;
  float eye[3] = {0, 0, 0}, dir[3] = {0, 0, 1}, up[3] = {0, 1, 0}, aov = 30.f;
  int width = 512, height = 512;  // Camera.h:203-204
  std::string annotation;
};

class Lighting {
 public:
  Lighting();
  bool LoadStateFromValue(const json::Value &v);
  gxy_lighting abi;
};

struct Vis {
  std::string type;     // "VolumeVis" | "TrianglesVis" | "ParticlesVis" | "PathLinesVis" | "GradientSamplerVis" | "IsoSamplerVis"
  std::string dataset;
  std::vector<float> colormap;    // n x (x,r,g,b)
  std::vector<float> opacitymap;  // m x (x,o)
  bool has_range = false;
  float range[2] = {0, 1};
  std::vector<float> isovalues, slices;  // slices: k x (a,b,c,d)
  bool volume_render = false;
  // ParticlesVis defaults (ParticlesVis.cpp:46-55,104-118); a PathLinesVis starts from -1, 1, 0, 1 (PathLinesVis.cpp:46-55)
  float radius0 = 0.025f, radius1 = 0.f, value0 = 0.f, value1 = 0.f;
  float tolerance = 0.f, isovalue = 0.f;  // GradientSamplerVis.cpp:77-85 / IsoSamplerVis.cpp:77-85 (a sampling Visualization, src/sampler)
  // base_dir: where a "colormap" / "transfer function" given as a file name (ParaView JSON, MappedVis.cpp:104-166) is looked up
  bool LoadFromJSON(const json::Value &v, const std::string &base_dir = "");
};

class Visualization {
 public:
  ~Visualization();
  bool LoadFromJSON(const json::Value &v, const std::string &base_dir = "");
  static bool LoadVisualizationsFromJSON(const json::Value &doc, std::vector<Visualization> &out, const std::string &base_dir = "");
  // builds one gxy_vis per partition (all on `ctx`'s device) from the datasets
  bool Commit(gxy_context *ctx, const Datasets &datasets, int nparts);
  void Release();
  std::string annotation;
  Lighting lighting;
  std::vector<Vis> operators;
  std::vector<gxy_vis *> parts;
  std::vector<gxy_volume *> owned_volumes;
  std::vector<gxy_triangles *> owned_triangles;
  std::vector<gxy_particles *> owned_particles;
  std::vector<gxy_pathlines *> owned_pathlines;
};

// Sampler (src/sampler/Sampler.h:34-72): a Renderer whose rays leave Particles where a sampler operator fires.  The
// Visualization must hold only GradientSampler / IsoSampler operators; the samples stay on the device, per partition.
class Sampler {
 public:
  // one frame of camera rays over all partitions of `visualization` (Sampler::Trace + HandleTerminatedRays, Sampler.cpp:52-133)
  bool Sample(const Camera &camera, Visualization &visualization, int width, int height);
  // Sampler::GetSamples: xyz of the samples of partition r (host copy) / as a device Particles dataset for a ParticlesVis
  bool GetSamples(const Visualization &visualization, int r, std::vector<float> &xyz) const;
  gxy_stats stats;
};

class Renderer {
 public:
  bool LoadStateFromDocument(const json::Value &doc);
  float epsilon = 0.001f;  // Renderer.cpp:125
};

class Rendering {
 public:
  Camera *camera = nullptr;
  Visualization *visualization = nullptr;
  int width = 512, height = 512;
  std::vector<unsigned char> rgba8;  // rows top-down
  gxy_stats stats;
  bool Render(const Renderer &renderer);
  // frame-stamped accumulation of the interactive path (Rendering.cpp:104-153): nothing is cleared between frames
  bool RenderProgressive(const Renderer &renderer, int frame);
  // Rendering::SaveImage (Rendering.cpp:272-293): <base>_%05d<vis annotation><camera annotation>.png, or
  // <base><annotations>.png when an annotation is present
  std::string ImageName(const std::string &base, int index) const;
  bool SaveImage(const std::string &base, int index) const;
};

bool write_png(const std::string &path, int w, int h, const unsigned char *rgba);

// description of the parsed state for tests / --describe (JSON text)
std::string describe_state(const Renderer &r, const std::vector<Camera> &cams, const std::vector<Visualization> &vis, const Datasets &ds,
                           int nparts);

}  // namespace gxy
