// gxy_host.cpp -- see gxy_host.h.  Compiled with -ffp-contract=off: the few float expressions here (light
// normalisation, partition boxes) decide planes and directions that must be bit-identical to the oracle's.
#include "gxy_host.h"
#include "gxy_vtu.h"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace gxy {

static std::string dir_of(const std::string &path) {
  const size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string("") : path.substr(0, k + 1);
}

// ---- Volume --------------------------------------------------------------------------------------
// Volume::local_import, header part (src/data/Volume.cpp:176-297)
bool Volume::Import(const std::string &fname) {
  filename = fname;
  const std::string dir = dir_of(fname);
  const size_t dot = fname.find_last_of('.');
  const std::string ext = dot == std::string::npos ? std::string("") : fname.substr(dot + 1);
  std::string data_fname;
  if (ext == "vol") {
    std::ifstream in(fname.c_str());
    if (in.fail()) {
      std::cerr << "ERROR: unable to open volfile: " << fname << std::endl;
      return false;
    }
    std::string type_string;
    in >> type_string;
    is_float = type_string == "float";
    in >> origin[0] >> origin[1] >> origin[2];
    in >> counts[0] >> counts[1] >> counts[2];
    in >> deltas[0] >> deltas[1] >> deltas[2];
    in >> data_fname;
    if (in.fail()) {
      std::cerr << "ERROR: malformed volfile header: " << fname << std::endl;
      return false;
    }
  } else if (ext == "json") {
    json::Value doc;
    try {
      doc = json::ParseFile(fname);
    } catch (const std::exception &e) {
      std::cerr << "JSON parse error in " << fname << ": " << e.what() << "\n";
      return false;
    }
    const char *need[] = {"type", "origin", "counts", "delta", "rawdata"};
    for (const char *k : need)
      if (!doc.HasMember(k)) {
        std::cerr << "volume JSON has no " << k << " field: " << fname << "\n";
        return false;
      }
    is_float = doc["type"].GetString() == std::string("float");
    for (int a = 0; a < 3; a++) {
      origin[a] = (float)doc["origin"][a].GetDouble();
      counts[a] = doc["counts"][a].GetInt();
      deltas[a] = (float)doc["delta"][a].GetDouble();
    }
    data_fname = doc["rawdata"].GetString();
    number_of_components = doc.HasMember("number of components") ? doc["number of components"].GetInt() : 1;
  } else {
    std::cerr << "Volume::local_import: unrecognized file extension (" << ext << ")\n";
    return false;
  }
  if (number_of_components != 1) {
    std::cerr << "Volume: only scalar volumes can be rendered (number of components = " << number_of_components << ")\n";
    return false;
  }
  raw_filename = (!data_fname.empty() && data_fname[0] == '/') ? data_fname : dir + data_fname;
  return true;
}

bool Volume::LoadBrick(int nparts, int rank, std::vector<unsigned char> &samples, VolumePart &part) const {
  int f[3];
  gxy_factor(nparts, f);
  std::vector<int> table((size_t)nparts * 15);
  gxy_partition(nparts, f, counts, table.data());
  const int *t = &table[(size_t)rank * 15];
  for (int a = 0; a < 3; a++) {
    part.ijk[a] = t[a]; part.offsets[a] = t[3 + a]; part.counts[a] = t[6 + a];
    part.goffsets[a] = t[9 + a]; part.gcounts[a] = t[12 + a];
  }
  const size_t sample_sz = is_float ? 4 : 1;
  const size_t row_sz = (size_t)part.gcounts[0] * sample_sz;
  samples.resize(row_sz * (size_t)part.gcounts[1] * (size_t)part.gcounts[2]);
  std::ifstream raw(raw_filename.c_str(), std::ios::in | std::ios::binary);
  if (raw.fail()) {
    std::cerr << "ERROR: unable to open raw volume data: " << raw_filename << std::endl;
    return false;
  }
  char *dst = reinterpret_cast<char *>(samples.data());
  for (int z = 0; z < part.gcounts[2]; z++)
    for (int y = 0; y < part.gcounts[1]; y++) {
      const std::streamoff src = (std::streamoff)((((long long)(part.goffsets[2] + z) * ((long long)counts[1] * counts[0])) +
                                                   ((long long)(part.goffsets[1] + y) * counts[0]) + part.goffsets[0]) *
                                                  (long long)sample_sz);
      raw.seekg(src, std::ios_base::beg);
      raw.read(dst, (std::streamsize)row_sz);
      if (raw.fail()) {
        std::cerr << "ERROR: short read from " << raw_filename << std::endl;
        return false;
      }
      dst += row_sz;
    }
  return true;
}

void Volume::Boxes(const VolumePart &part, float gmin[3], float gmax[3], float lmin[3], float lmax[3]) const {
  for (int a = 0; a < 3; a++) {
    const float go = origin[a] + deltas[a];  // the global box excludes the outermost shell (Volume.cpp:379-381)
    const int gc = counts[a] - 2;
    gmin[a] = go;
    gmax[a] = go + (float)(gc - 1) * deltas[a];  // Box(origin, counts, deltas), Box.cpp:69-80
    const float lo = origin[a] + (float)part.offsets[a] * deltas[a];
    lmin[a] = lo;
    lmax[a] = lo + (float)(part.counts[a] - 1) * deltas[a];
  }
}

void Volume::Neighbors(const VolumePart &p, const int f[3], int nb[6]) {
  auto rank = [&](int i, int j, int k) { return i + j * f[0] + k * f[0] * f[1]; };
  const int i = p.ijk[0], j = p.ijk[1], k = p.ijk[2];
  nb[0] = i > 0 ? rank(i - 1, j, k) : -1;
  nb[1] = i < f[0] - 1 ? rank(i + 1, j, k) : -1;
  nb[2] = j > 0 ? rank(i, j - 1, k) : -1;
  nb[3] = j < f[1] - 1 ? rank(i, j + 1, k) : -1;
  nb[4] = k > 0 ? rank(i, j, k - 1) : -1;
  nb[5] = k < f[2] - 1 ? rank(i, j, k + 1) : -1;
}

// ---- Datasets ------------------------------------------------------------------------------------
static bool load_typed(const json::Value &v, const std::string &state_dir, Datasets &out) {
  if (!v.HasMember("filename") || !v.HasMember("type")) {
    std::cerr << "Dataset must have name and type\n";
    return false;
  }
  std::string name = v["filename"].GetString();
  const std::string type = v["type"].GetString();
  if (type == "Volume") {
    Volume vol;
    std::string fn = v["filename"].GetString();
    if (!fn.empty() && fn[0] != '/') fn = state_dir + fn;  // the reference resolves against the working directory
    if (!vol.Import(fn)) return false;
    if (v.HasMember("name")) name = v["name"].GetString();
    vol.name = name;
    out.volumes.push_back(vol);
    return true;
  }
  if (type == "Particles" || type == "Triangles" || type == "PathLines") {
    Geometry g;
    g.type = type;
    std::string fn = v["filename"].GetString();
    if (!fn.empty() && fn[0] != '/') fn = state_dir + fn;
    if (!g.Import(fn)) return false;
    if (v.HasMember("name")) name = v["name"].GetString();
    g.name = name;
    out.geometries.push_back(g);
    return true;
  }
  std::cerr << "invalid Dataset type: " << type << "\n";
  return false;
}

bool Datasets::LoadFromJSON(const json::Value &doc, const std::string &state_dir) {
  if (!doc.HasMember("Datasets")) {
    std::cerr << "JSON has no Datasets clause\n";
    return false;
  }
  const json::Value &ds = doc["Datasets"];
  if (ds.IsArray()) {
    for (size_t i = 0; i < ds.Size(); i++)
      if (!load_typed(ds[i], state_dir, *this)) return false;
    return true;
  }
  return load_typed(ds, state_dir, *this);
}

const Geometry *Datasets::FindGeometry(const std::string &name) const {
  for (const Geometry &g : geometries)
    if (g.name == name) return &g;
  return nullptr;
}

// ---- Geometry ------------------------------------------------------------------------------------
bool Geometry::Import(const std::string &fname) {
  filename = fname;
  json::Value doc;
  try {
    doc = json::ParseFile(fname);
  } catch (const std::exception &e) {
    std::cerr << "parse error in partition document: " << fname << " (" << e.what() << ")\n";
    return false;
  }
  if (!doc.HasMember("parts")) {
    std::cerr << "partition document does not have parts\n";  // Geometry.cpp:283-288
    return false;
  }
  const json::Value &parts = doc["parts"];
  if (!parts.IsArray() || parts.Size() == 0) {
    std::cerr << "invalid partition document\n";
    return false;
  }
  const std::string dir = dir_of(fname);
  try {
    for (size_t i = 0; i < parts.Size(); i++) {
      if (!parts[i].HasMember("filename")) {
        std::cerr << "partition document: only file parts are supported (part " << i << " has no filename)\n";
        return false;
      }
      std::string f = parts[i]["filename"].GetString();
      if (!f.empty() && f[0] != '/') f = dir + f;
      part_files.push_back(f);
      for (int j = 0; j < 6; j++) extents.push_back((float)parts[i]["extent"][j].GetDouble());
    }
  } catch (const std::exception &e) {
    std::cerr << "invalid partition document: " << e.what() << "\n";
    return false;
  }
  return true;
}

void Geometry::Boxes(int rank, float gmin[3], float gmax[3], float lmin[3], float lmax[3], int neighbors[6]) const {
  float g[6] = {3.402823466e+38f, -3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f};
  const float *l = &extents[(size_t)rank * 6];
  for (int i = 0; i < 6; i++) neighbors[i] = -1;
  for (int i = 0; i < NumberOfParts(); i++) {
    const float *e = &extents[(size_t)i * 6];
    for (int a = 0; a < 3; a++) {
      if (e[2 * a] < g[2 * a]) g[2 * a] = e[2 * a];
      if (e[2 * a + 1] > g[2 * a + 1]) g[2 * a + 1] = e[2 * a + 1];
    }
    auto LAST = [&](int a) { return e[2 * a + 1] == l[2 * a]; };
    auto NEXT = [&](int a) { return e[2 * a] == l[2 * a + 1]; };
    auto EQ = [&](int a) { return e[2 * a] == l[2 * a]; };
    if (LAST(0) && EQ(1) && EQ(2)) neighbors[0] = i;
    if (NEXT(0) && EQ(1) && EQ(2)) neighbors[1] = i;
    if (EQ(0) && LAST(1) && EQ(2)) neighbors[2] = i;
    if (EQ(0) && NEXT(1) && EQ(2)) neighbors[3] = i;
    if (EQ(0) && EQ(1) && LAST(2)) neighbors[4] = i;
    if (EQ(0) && EQ(1) && NEXT(2)) neighbors[5] = i;
  }
  for (int a = 0; a < 3; a++) { gmin[a] = g[2 * a]; gmax[a] = g[2 * a + 1]; lmin[a] = l[2 * a]; lmax[a] = l[2 * a + 1]; }
}

bool Geometry::LoadPiece(int rank, GeometryPiece &out) const {
  VtuData d;
  std::string err;
  if (!read_vtu(part_files[(size_t)rank], d, err)) {
    std::cerr << "error reading " << part_files[(size_t)rank] << ": " << err << "\n";
    return false;
  }
  const size_t nv = (size_t)d.n_points;
  out.vertices = d.points;
  if (type == "Triangles") {
    if (d.normals.empty()) {  // Triangles.cpp:107-118
      if (nv) std::cerr << "triangle set has no normals\n";
      out.normals.resize(3 * nv);
      for (size_t i = 0; i < nv; i++) { out.normals[3 * i] = 1.0f; out.normals[3 * i + 1] = 0.0f; out.normals[3 * i + 2] = 0.0f; }
    } else {
      out.normals = d.normals;
    }
    long long prev = 0;
    for (long long o : d.offsets) {
      if (o - prev != 3) {
        std::cerr << part_files[(size_t)rank] << ": a Triangles dataset may only hold triangle cells\n";
        return false;
      }
      prev = o;
    }
    if (d.connectivity.size() % 3) {
      std::cerr << part_files[(size_t)rank] << ": connectivity is not a list of triangles\n";
      return false;
    }
    out.connectivity = d.connectivity;
  }
  if (type == "PathLines") {
    // PathLines::load_from_vtkPointSet (PathLines.cpp:73-125): the points of every cell are copied in cell order (a point
    // shared by two lines is duplicated), data follows them, connectivity[j] = index of the first vertex of each segment
    out.vertices.clear();
    out.data.clear();
    long long prev = 0;
    for (long long o : d.offsets) {
      if (o < prev || (size_t)o > d.connectivity.size()) {
        std::cerr << part_files[(size_t)rank] << ": bad cell offsets\n";
        return false;
      }
      for (long long l = prev; l < o; l++) {
        const int id = d.connectivity[(size_t)l];
        if (id < 0 || (size_t)id >= nv) {
          std::cerr << part_files[(size_t)rank] << ": point id " << id << " out of range\n";
          return false;
        }
        const int k = (int)(out.vertices.size() / 3);
        for (int a = 0; a < 3; a++) out.vertices.push_back(d.points[3 * (size_t)id + a]);
        out.data.push_back(d.scalars.empty() ? 0.0f : d.scalars[(size_t)id]);
        if (l < o - 1) out.connectivity.push_back(k);
      }
      prev = o;
    }
    return true;
  }
  out.data = d.scalars;
  if (out.data.empty()) out.data.assign(nv, 0.0f);  // Triangles.cpp:127-132 / Particles.cpp:124
  return true;
}

const Volume *Datasets::FindVolume(const std::string &name) const {
  for (const Volume &v : volumes)
    if (v.name == name) return &v;
  return nullptr;
}

// ---- Camera --------------------------------------------------------------------------------------
static std::string resolve_path(const std::string &name, const std::string &base_dir) {
  if (name.empty() || name[0] == '/' || base_dir.empty()) return name;
  std::ifstream probe((base_dir + name).c_str());
  return probe ? base_dir + name : name;  // the reference resolves against the working directory
}

// Camera::LoadFromPVCC (Camera.cpp:118-163): ParaView's camera configuration as XML,
//   <PVCameraConfiguration><Proxy><Property name="CameraPosition"><Element index="0" value="1.5"/>...
// read with boost::property_tree there; here a scan for the Property / Element tags below PVCameraConfiguration.Proxy
// (attribute values as the stream extraction to float reads them: strtof).
static bool xml_attr(const std::string &tag, const char *key, std::string &out) {
  size_t p = 0;
  const std::string k(key);
  while ((p = tag.find(k, p)) != std::string::npos) {
    const bool start_ok = p > 0 && isspace((unsigned char)tag[p - 1]);
    size_t q = p + k.size();
    while (q < tag.size() && isspace((unsigned char)tag[q])) q++;
    if (start_ok && q < tag.size() && tag[q] == '=') {
      q++;
      while (q < tag.size() && isspace((unsigned char)tag[q])) q++;
      if (q < tag.size() && (tag[q] == '"' || tag[q] == '\'')) {
        const size_t e = tag.find(tag[q], q + 1);
        if (e == std::string::npos) return false;
        out = tag.substr(q + 1, e - q - 1);
        return true;
      }
    }
    p += k.size();
  }
  return false;
}

bool Camera::LoadFromPVCC(const std::string &filename) {
  std::ifstream ifs(filename.c_str());
  if (!ifs) return false;
  std::stringstream ss;
  ss << ifs.rdbuf();
  const std::string s = ss.str();
  float center[3] = {0, 0, 0};
  bool in_config = false, in_proxy = false, seen_proxy = false;
  std::string property;
  float values[3] = {-1e32f, -1e32f, -1e32f};
  auto assign = [&]() {
    if (property == "CameraPosition") for (int i = 0; i < 3; i++) eye[i] = values[i];
    else if (property == "CameraFocalPoint") for (int i = 0; i < 3; i++) center[i] = values[i];
    else if (property == "CameraViewUp") for (int i = 0; i < 3; i++) up[i] = values[i];
    else if (property == "CameraViewAngle") aov = values[0];
    property.clear();
  };
  size_t pos = 0;
  while (true) {
    const size_t lt = s.find('<', pos);
    if (lt == std::string::npos) break;
    if (s.compare(lt, 4, "<!--") == 0) {
      const size_t e = s.find("-->", lt);
      if (e == std::string::npos) return false;
      pos = e + 3;
      continue;
    }
    const size_t gt = s.find('>', lt);
    if (gt == std::string::npos) return false;
    const std::string tag = s.substr(lt + 1, gt - lt - 1);
    pos = gt + 1;
    if (tag.empty() || tag[0] == '?' || tag[0] == '!') continue;
    size_t n = 0;
    while (n < tag.size() && !isspace((unsigned char)tag[n]) && tag[n] != '/') n++;
    const bool closing = tag[0] == '/';
    std::string name = closing ? tag.substr(1) : tag.substr(0, n);
    if (closing) {
      const size_t ws = name.find_first_of(" \t\r\n");
      if (ws != std::string::npos) name.erase(ws);
    }
    const bool self_closing = !closing && tag.back() == '/';
    if (name == "PVCameraConfiguration") in_config = !closing;
    else if (name == "Proxy" && in_config) { in_proxy = !closing && !self_closing; seen_proxy = true; }
    else if (name == "Property" && in_proxy) {
      if (closing) assign();
      else {
        if (!xml_attr(tag, "name", property)) return false;   // get<std::string>("<xmlattr>.name") throws -> catch(...) -> false
        values[0] = values[1] = values[2] = -1e32f;
        if (self_closing) assign();
      }
    } else if (name == "Element" && in_proxy && !property.empty() && !closing) {
      std::string si, sv;
      if (!xml_attr(tag, "index", si) || !xml_attr(tag, "value", sv)) return false;
      char *end = nullptr;
      const long indx = strtol(si.c_str(), &end, 10);
      if (end == si.c_str() || indx < 0 || indx > 2) return false;   // the reference writes values[indx] unchecked
      const float val = strtof(sv.c_str(), &end);
      if (end == sv.c_str()) return false;
      values[indx] = val;
    }
  }
  if (!seen_proxy) return false;   // get_child("PVCameraConfiguration.Proxy") throws
  for (int k = 0; k < 3; k++) dir[k] = center[k] - eye[k];
  return true;
}

bool Camera::LoadFromJSON(const json::Value &v, const std::string &base_dir) {
  if (v.IsString()) {  // Camera.cpp:168-232: a ParaView camera configuration, converted to JSON or as ParaView writes it (.pvcc, XML)
    const std::string fname = resolve_path(v.GetString(), base_dir);
    {
      std::ifstream probe(fname.c_str());
      if (!probe) {
        std::cerr << "unable to open " << v.GetString() << "\n";
        return false;
      }
    }
    json::Value doc;
    try {
      doc = json::ParseFile(fname);
    } catch (const std::exception &e) {
      if (!LoadFromPVCC(fname)) {  // Camera.cpp:182-191: not JSON -> try the XML form
        std::cerr << "error loading camera from " << v.GetString() << "\n";
        return false;
      }
      return true;
    }
    try {
      if (!doc.HasMember("PVCameraConfiguration") || !doc["PVCameraConfiguration"].HasMember("Proxy") ||
          !doc["PVCameraConfiguration"]["Proxy"].HasMember("Property")) {
        std::cerr << "invalid Paraview camera file: " << v.GetString() << "\n";
        return false;
      }
      const json::Value &props = doc["PVCameraConfiguration"]["Proxy"]["Property"];
      float center[3] = {0, 0, 0};
      for (size_t i = 0; i < props.Size(); i++) {
        const json::Value &p = props[i];
        if (!p.HasMember("@name")) continue;
        const std::string name = p["@name"].GetString();
        auto elem = [&](int k) { return (float)atof(p["Element"][(size_t)k]["@value"].GetString().c_str()); };
        if (name == "CameraPosition") for (int k = 0; k < 3; k++) eye[k] = elem(k);
        else if (name == "CameraFocalPoint") for (int k = 0; k < 3; k++) center[k] = elem(k);
        else if (name == "CameraViewUp") for (int k = 0; k < 3; k++) up[k] = elem(k);
        else if (name == "CameraViewAngle") aov = (float)atof(p["Element"]["@value"].GetString().c_str());
      }
      for (int k = 0; k < 3; k++) dir[k] = center[k] - eye[k];
    } catch (const std::exception &e) {
      std::cerr << "invalid Paraview camera file: " << v.GetString() << " (" << e.what() << ")\n";
      return false;
    }
    return true;
  }
  try {
    if (v.HasMember("annotation")) annotation = v["annotation"].GetString();
    for (int a = 0; a < 3; a++) eye[a] = (float)v["viewpoint"][a].GetDouble();
    if (v.HasMember("viewdirection")) {
      for (int a = 0; a < 3; a++) dir[a] = (float)v["viewdirection"][a].GetDouble();
    } else if (v.HasMember("viewcenter")) {
      for (int a = 0; a < 3; a++) dir[a] = (float)(v["viewcenter"][a].GetDouble() - (double)eye[a]);  // Camera.cpp:254-256
    } else {
      std::cerr << "need either viewdirection or viewcenter\n";
      return false;
    }
    if (v.HasMember("dimensions")) {
      width = v["dimensions"][0].GetInt();
      height = v["dimensions"][0].GetInt();  // sic: Camera.cpp:267-271 reads [0] for both
    }
    for (int a = 0; a < 3; a++) up[a] = (float)v["viewup"][a].GetDouble();
    aov = (float)v["aov"].GetDouble();
  } catch (const std::exception &e) {
    std::cerr << "error loading camera: " << e.what() << "\n";
    return false;
  }
  return true;
}

bool Camera::LoadCamerasFromJSON(const json::Value &doc, std::vector<Camera> &out, const std::string &base_dir) {
  const json::Value *c = doc.Find("Cameras");
  if (!c) c = doc.Find("Camera");
  if (!c) {
    std::cerr << "JSON has no Camera or Cameras clause\n";
    return false;
  }
  if (c->IsArray()) {
    for (size_t i = 0; i < c->Size(); i++) {
      Camera cam;
      if (!cam.LoadFromJSON((*c)[i], base_dir)) return false;
      out.push_back(cam);
    }
  } else {
    Camera cam;
    if (!cam.LoadFromJSON(*c, base_dir)) return false;
    out.push_back(cam);
  }
  return true;
}

gxy_camera Camera::AsABI() const {
  gxy_camera c;
  for (int a = 0; a < 3; a++) { c.eye[a] = eye[a]; c.dir[a] = dir[a]; c.up[a] = up[a]; }
  c.aov = aov;
  return c;
}

// ---- Lighting ------------------------------------------------------------------------------------
Lighting::Lighting() {  // Lighting.ispc:29-44 / Lighting.cpp:43-57: one point light at (1,1,1), no AO, no shadows, Ka = Kd = 0.5
  memset(&abi, 0, sizeof abi);
  abi.n_lights = 1;
  abi.lights[0][0] = abi.lights[0][1] = abi.lights[0][2] = 1.0f;
  abi.types[0] = 2;
  abi.n_ao = 0;
  abi.ao_radius = 1.0f;
  abi.shadows = 0;
  abi.Ka = 0.5f;
  abi.Kd = 0.5f;
}

bool Lighting::LoadStateFromValue(const json::Value &v) {
  try {
    if (v.HasMember("Sources")) {
      const json::Value &s = v["Sources"];
      if (s.Size() > GXY_MAX_LIGHTS) {
        std::cerr << "at most " << GXY_MAX_LIGHTS << " light sources are supported\n";
        return false;
      }
      abi.n_lights = (int)s.Size();
      for (size_t i = 0; i < s.Size(); i++) {
        float x = (float)s[i][0].GetDouble(), y = (float)s[i][1].GetDouble(), z = (float)s[i][2].GetDouble();
        const int t = s[i].Size() == 3 ? 0 : s[i][3].GetInt();
        if (t == 0) {
          const float d = sqrtf(x * x + y * y + z * z);
          if (d == 0.0f) {
            std::cerr << "WARNING: Directional light souce cannot be (0,0,0) so using (1,1,1)" << std::endl;
            x = y = z = 0.577350f;
          } else {
            x = x / d; y = y / d; z = z / d;
          }
        }
        abi.lights[i][0] = x; abi.lights[i][1] = y; abi.lights[i][2] = z;
        abi.types[i] = t;
      }
    }
    if (v.HasMember("shadows")) abi.shadows = v["shadows"].GetBool() ? 1 : 0;
    abi.n_ao = v.HasMember("ao count") ? v["ao count"].GetInt() : 0;
    abi.ao_radius = v.HasMember("ao radius") ? (float)v["ao radius"].GetDouble() : 1.0f;
    abi.Ka = v.HasMember("Ka") ? (float)v["Ka"].GetDouble() : 0.5f;
    abi.Kd = v.HasMember("Kd") ? (float)v["Kd"].GetDouble() : 0.5f;
  } catch (const std::exception &e) {
    std::cerr << "error loading Lighting: " << e.what() << "\n";
    return false;
  }
  return true;
}

// ---- Vis -----------------------------------------------------------------------------------------
bool Vis::LoadFromJSON(const json::Value &v, const std::string &base_dir) {
  try {
    type = v["type"].GetString();
    if (type.size() < 3 || type.compare(type.size() - 3, 3, "Vis")) type += "Vis";  // Visualization.cpp:318-320
    if (v.HasMember("dataset")) dataset = v["dataset"].GetString();
    else {
      std::cerr << "Vis needs a dataset\n";  // Vis.cpp:126-146 (the key form needs the reference's object registry)
      return false;
    }
    // MappedVis defaults (MappedVis.cpp:58-62)
    colormap = {0.0f, 0.4f, 0.4f, 0.4f, 1.0f, 1.0f, 1.0f, 1.0f};
    opacitymap = {0.0f, 1.0f, 1.0f, 1.0f};
    const json::Value *m = v.Find("transfer function");
    if (!m) m = v.Find("colormap");
    if (m && m->IsString()) {  // MappedVis.cpp:104-166: a ParaView colormap export ("RGBPoints", optional "Points")
      const std::string fname = m->GetString();
      if (!fname.empty() && fname != "default") {
        json::Value doc;
        try {
          doc = json::ParseFile(resolve_path(fname, base_dir));
        } catch (const std::exception &e) {
          std::cerr << "unable to open transfer function file: " << fname << " (" << e.what() << ")\n";
          return false;
        }
        const json::Value &cmap = doc.IsArray() ? doc[0] : doc;
        opacitymap.clear();
        if (cmap.HasMember("Points")) {
          const json::Value &oa = cmap["Points"];
          for (size_t i = 0; i + 1 < oa.Size(); i += 4) {
            opacitymap.push_back((float)oa[i].GetDouble());
            opacitymap.push_back((float)oa[i + 1].GetDouble());
          }
        } else {
          opacitymap = {0.0f, 1.0f, 1.0f, 1.0f};
        }
        colormap.clear();
        const json::Value &rgba = cmap["RGBPoints"];
        for (size_t i = 0; i + 3 < rgba.Size(); i += 4)
          for (int k = 0; k < 4; k++) colormap.push_back((float)rgba[i + k].GetDouble());
      }
    } else if (m) {
      colormap.clear();
      for (size_t i = 0; i < m->Size(); i++)
        for (int k = 0; k < 4; k++) colormap.push_back((float)(*m)[i][k].GetDouble());
      if (v.HasMember("opacitymap")) {
        opacitymap.clear();
        const json::Value &o = v["opacitymap"];
        for (size_t i = 0; i < o.Size(); i++)
          for (int k = 0; k < 2; k++) opacitymap.push_back((float)o[i][k].GetDouble());
      }
    }
    if (v.HasMember("data range")) {
      has_range = true;
      range[0] = (float)v["data range"][0].GetDouble();
      range[1] = (float)v["data range"][1].GetDouble();
    }
    if (type == "VolumeVis") {  // VolumeVis.cpp:119-163
      if (v.HasMember("isovalues"))
        for (size_t i = 0; i < v["isovalues"].Size(); i++) isovalues.push_back((float)v["isovalues"][i].GetDouble());
      if (v.HasMember("slices")) {
        for (size_t i = 0; i < v["slices"].Size(); i++)
          for (int k = 0; k < 4; k++) slices.push_back((float)v["slices"][i][k].GetDouble());
      } else if (v.HasMember("plane")) {
        for (int k = 0; k < 4; k++) slices.push_back((float)v["plane"][k].GetDouble());
      }
      volume_render = v.HasMember("volume rendering") ? v["volume rendering"].GetBool() : false;
    } else if (type == "GradientSamplerVis") {  // src/sampler/GradientSamplerVis.cpp:77-85
      if (v.HasMember("tolerance")) tolerance = (float)v["tolerance"].GetDouble();
    } else if (type == "IsoSamplerVis") {  // src/sampler/IsoSamplerVis.cpp:77-85
      if (v.HasMember("isovalue")) isovalue = (float)v["isovalue"].GetDouble();
    } else if (type == "PathLinesVis") {  // PathLinesVis.cpp:46-55 (initialize), 105-114 (LoadFromJSON)
      radius0 = -1.0f; radius1 = 1.0f; value0 = 0.0f; value1 = 1.0f;
      if (v.HasMember("radius0")) radius0 = (float)v["radius0"].GetDouble();
      if (v.HasMember("radius1")) radius1 = (float)v["radius1"].GetDouble();
      if (v.HasMember("value0")) value0 = (float)v["value0"].GetDouble();
      if (v.HasMember("value1")) value1 = (float)v["value1"].GetDouble();
    } else if (type == "ParticlesVis") {  // ParticlesVis.cpp:104-118
      if (v.HasMember("radius0")) radius0 = (float)v["radius0"].GetDouble();
      if (v.HasMember("radius1")) radius1 = (float)v["radius1"].GetDouble();
      if (v.HasMember("value0")) value0 = (float)v["value0"].GetDouble();
      if (v.HasMember("value1")) value1 = (float)v["value1"].GetDouble();
      if (v.HasMember("radius")) { radius0 = (float)v["radius"].GetDouble(); radius1 = 0.f; value0 = 0.f; value1 = 0.f; }
    }
  } catch (const std::exception &e) {
    std::cerr << "error loading a Vis: " << e.what() << "\n";
    return false;
  }
  return true;
}

// ---- Visualization -------------------------------------------------------------------------------
Visualization::~Visualization() { Release(); }

void Visualization::Release() {
  for (gxy_vis *p : parts) gxy_vis_destroy(p);
  parts.clear();
  for (gxy_volume *v : owned_volumes) gxy_volume_destroy(v);
  owned_volumes.clear();
  for (gxy_triangles *t : owned_triangles) gxy_triangles_destroy(t);
  owned_triangles.clear();
  for (gxy_particles *p : owned_particles) gxy_particles_destroy(p);
  owned_particles.clear();
  for (gxy_pathlines *p : owned_pathlines) gxy_pathlines_destroy(p);
  owned_pathlines.clear();
}

bool Visualization::LoadFromJSON(const json::Value &v, const std::string &base_dir) {
  try {
    if (v.HasMember("annotation")) annotation = v["annotation"].GetString();
    const json::Value *l = v.Find("Lighting");
    if (!l) l = v.Find("lighting");
    if (l && !lighting.LoadStateFromValue(*l)) return false;
    if (!v.HasMember("operators")) {
      std::cerr << "Visualization has no operators\n";
      return false;
    }
    const json::Value &ops = v["operators"];
    for (size_t i = 0; i < ops.Size(); i++) {
      Vis op;
      if (!op.LoadFromJSON(ops[i], base_dir)) return false;
      operators.push_back(op);
    }
  } catch (const std::exception &e) {
    std::cerr << "error loading a Visualization: " << e.what() << "\n";
    return false;
  }
  return true;
}

bool Visualization::LoadVisualizationsFromJSON(const json::Value &doc, std::vector<Visualization> &out, const std::string &base_dir) {
  const json::Value *v = doc.Find("Visualization");
  if (!v) v = doc.Find("Visualizations");
  if (!v) {
    std::cerr << "JSON has no Visualization or Visualizations clause\n";
    return false;
  }
  const size_t n = v->IsArray() ? v->Size() : 1;
  out.resize(n);  // in place: a Visualization owns device handles and is not copyable in spirit
  for (size_t i = 0; i < n; i++)
    if (!out[i].LoadFromJSON(v->IsArray() ? (*v)[i] : *v, base_dir)) return false;
  return true;
}

static bool check_abi(int rc, const char *what) {
  if (rc) std::cerr << what << ": " << gxy_last_error() << "\n";
  return rc == 0;
}

bool Visualization::Commit(gxy_context *ctx, const Datasets &datasets, int nparts) {
  Release();
  int f[3];
  gxy_factor(nparts, f);
  for (int r = 0; r < nparts; r++) {
    gxy_vis *vis = nullptr;
    if (!check_abi(gxy_vis_create(ctx, &vis), "gxy_vis_create")) return false;
    parts.push_back(vis);
    bool boxes_set = false;
    std::vector<std::pair<std::string, gxy_volume *>> vols;  // one device volume per dataset and partition
    for (const Vis &op : operators) {
      if (op.type == "TrianglesVis" || op.type == "ParticlesVis" || op.type == "PathLinesVis") {
        const Geometry *geo = datasets.FindGeometry(op.dataset);
        if (!geo) {
          std::cerr << "Unable to find data using name: " << op.dataset << "\n";
          return false;
        }
        if (geo->type + "Vis" != op.type) {
          std::cerr << op.type << " on a " << geo->type << " dataset (" << op.dataset << ")\n";
          return false;
        }
        if (geo->NumberOfParts() != nparts) {
          std::cerr << "invalid partition document: " << geo->filename << " has " << geo->NumberOfParts() << " parts for " << nparts
                    << " partitions\n";  // Geometry.cpp:291-296
          return false;
        }
        if (!boxes_set) {
          float gmin[3], gmax[3], lmin[3], lmax[3];
          int nb[6];
          geo->Boxes(r, gmin, gmax, lmin, lmax, nb);
          if (!check_abi(gxy_vis_set_partition(vis, gmin, gmax, lmin, lmax, nb), "gxy_vis_set_partition")) return false;
          boxes_set = true;
        }
        GeometryPiece piece;
        if (!geo->LoadPiece(r, piece)) return false;
        gxy_transfer_function gtf;
        if (!check_abi(gxy_resample_transfer_function((int)op.colormap.size() / 4, op.colormap.data(), (int)op.opacitymap.size() / 2,
                                                      op.opacitymap.data(), &gtf),
                       "gxy_resample_transfer_function"))
          return false;
        gtf.range_lo = op.has_range ? op.range[0] : op.colormap[0];
        gtf.range_hi = op.has_range ? op.range[1] : op.colormap[op.colormap.size() - 4];
        const int nv = (int)(piece.vertices.size() / 3);
        if (op.type == "TrianglesVis") {
          gxy_triangles *gt = nullptr;
          if (!check_abi(gxy_triangles_create(ctx, nv, piece.vertices.data(), piece.normals.data(), piece.data.data(),
                                              (int)(piece.connectivity.size() / 3), piece.connectivity.data(), &gt),
                         "gxy_triangles_create"))
            return false;
          owned_triangles.push_back(gt);
          if (!check_abi(gxy_vis_add_triangles(vis, gt, &gtf), "gxy_vis_add_triangles")) return false;
        } else if (op.type == "PathLinesVis") {
          gxy_pathlines *gl = nullptr;
          if (!check_abi(gxy_pathlines_create(ctx, nv, piece.vertices.data(), piece.data.data(), (int)piece.connectivity.size(),
                                              piece.connectivity.data(), &gl),
                         "gxy_pathlines_create"))
            return false;
          owned_pathlines.push_back(gl);
          if (!check_abi(gxy_vis_add_pathlines(vis, gl, op.radius0, op.radius1, op.value0, op.value1, &gtf), "gxy_vis_add_pathlines")) return false;
        } else {
          gxy_particles *gp = nullptr;
          if (!check_abi(gxy_particles_create(ctx, nv, piece.vertices.data(), piece.data.data(), &gp), "gxy_particles_create")) return false;
          owned_particles.push_back(gp);
          if (!check_abi(gxy_vis_add_particles(vis, gp, op.radius0, op.radius1, op.value0, op.value1, &gtf), "gxy_vis_add_particles")) return false;
        }
        continue;
      }
      const bool sampler_op = op.type == "GradientSamplerVis" || op.type == "IsoSamplerVis";
      if (op.type != "VolumeVis" && !sampler_op) {
        std::cerr << op.type << " is not supported by this driver\n";
        return false;
      }
      const Volume *vol = datasets.FindVolume(op.dataset);
      if (!vol) {
        std::cerr << "Unable to find data using name: " << op.dataset << "\n";  // Vis.cpp:83-96
        return false;
      }
      gxy_volume *dv = nullptr;
      for (auto &kv : vols)
        if (kv.first == op.dataset) dv = kv.second;
      VolumePart part;
      if (!dv) {
        std::vector<unsigned char> samples;
        if (!vol->LoadBrick(nparts, r, samples, part)) return false;
        float org[3];
        for (int a = 0; a < 3; a++) org[a] = vol->origin[a] + (float)part.goffsets[a] * vol->deltas[a];  // Volume.h:124-129
        if (!check_abi(gxy_volume_create(ctx, part.gcounts, org, vol->deltas, vol->is_float ? 0 : 1, samples.data(), &dv), "gxy_volume_create"))
          return false;
        owned_volumes.push_back(dv);
        vols.push_back(std::make_pair(op.dataset, dv));
      } else {
        std::vector<unsigned char> unused;
        int ff[3];
        gxy_factor(nparts, ff);
        std::vector<int> table((size_t)nparts * 15);
        gxy_partition(nparts, ff, vol->counts, table.data());
        const int *t = &table[(size_t)r * 15];
        for (int a = 0; a < 3; a++) { part.ijk[a] = t[a]; part.offsets[a] = t[3 + a]; part.counts[a] = t[6 + a]; part.goffsets[a] = t[9 + a]; part.gcounts[a] = t[12 + a]; }
      }
      if (!boxes_set) {  // boxes and neighbours come from the first Vis's dataset (Visualization.cpp:139-160)
        float gmin[3], gmax[3], lmin[3], lmax[3];
        int nb[6];
        vol->Boxes(part, gmin, gmax, lmin, lmax);
        Volume::Neighbors(part, f, nb);
        if (!check_abi(gxy_vis_set_partition(vis, gmin, gmax, lmin, lmax, nb), "gxy_vis_set_partition")) return false;
        boxes_set = true;
      }
      if (sampler_op) {  // a sampling Visualization (src/sampler): no transfer function, one parameter
        if (!check_abi(gxy_vis_add_sampler(vis, dv, op.type == "IsoSamplerVis" ? GXY_SAMPLER_ISO : GXY_SAMPLER_GRADIENT,
                                           op.type == "IsoSamplerVis" ? op.isovalue : op.tolerance),
                       "gxy_vis_add_sampler"))
          return false;
        continue;
      }
      gxy_transfer_function tf;
      if (!check_abi(gxy_resample_transfer_function((int)op.colormap.size() / 4, op.colormap.data(), (int)op.opacitymap.size() / 2,
                                                    op.opacitymap.data(), &tf),
                     "gxy_resample_transfer_function"))
        return false;
      // valueRange: "data range" if given, else the colormap's x-range (MappedVis.cpp:277-338)
      tf.range_lo = op.has_range ? op.range[0] : op.colormap[0];
      tf.range_hi = op.has_range ? op.range[1] : op.colormap[op.colormap.size() - 4];
      if (!check_abi(gxy_vis_add_volume(vis, dv, (int)op.slices.size() / 4, op.slices.data(), (int)op.isovalues.size(), op.isovalues.data(),
                                        op.volume_render ? 1 : 0, &tf),
                     "gxy_vis_add_volume"))
        return false;
    }
    if (!check_abi(gxy_vis_commit(vis), "gxy_vis_commit")) return false;
  }
  return true;
}

// ---- Sampler -------------------------------------------------------------------------------------
bool Sampler::Sample(const Camera &camera, Visualization &visualization, int width, int height) {
  if (visualization.parts.empty()) {
    std::cerr << "Sampler::Sample: the Visualization is not committed\n";
    return false;
  }
  const gxy_camera cam = camera.AsABI();
  return check_abi(gxy_sample((int)visualization.parts.size(), visualization.parts.data(), &cam, width, height, &stats), "gxy_sample");
}

bool Sampler::GetSamples(const Visualization &visualization, int r, std::vector<float> &xyz) const {
  if (r < 0 || (size_t)r >= visualization.parts.size()) return false;
  long long n = 0;
  if (!check_abi(gxy_vis_sample_count(visualization.parts[(size_t)r], &n), "gxy_vis_sample_count")) return false;
  xyz.resize(3 * (size_t)n);
  return check_abi(gxy_vis_download_samples(visualization.parts[(size_t)r], xyz.data()), "gxy_vis_download_samples");
}

// ---- Renderer / Rendering ------------------------------------------------------------------------
bool Renderer::LoadStateFromDocument(const json::Value &doc) {  // Renderer.cpp:289-296
  const json::Value *r = doc.Find("Renderer");
  if (r && r->HasMember("epsilon")) epsilon = (float)(*r)["epsilon"].GetDouble();
  return true;
}

bool Rendering::Render(const Renderer &renderer) {
  if (!camera || !visualization || visualization->parts.empty()) {
    std::cerr << "Rendering: camera and a committed visualization are needed\n";
    return false;
  }
  const gxy_camera cam = camera->AsABI();
  memset(&stats, 0, sizeof stats);
  if (!check_abi(gxy_render((int)visualization->parts.size(), visualization->parts.data(), &cam, &visualization->lighting.abi, width, height,
                            renderer.epsilon, &stats),
                 "gxy_render"))
    return false;
  rgba8.resize((size_t)width * height * 4);
  return check_abi(gxy_frame_download_rgba8(visualization->parts[0], rgba8.data()), "gxy_frame_download_rgba8");
}

// The interactive viewer's frame (Rendering::AddLocalPixels without GXY_WRITE_IMAGES, Rendering.cpp:104-153): the image owner keeps
// the displayed image and the per-pixel frame stamps across calls; rgba8 receives what is displayed after frame `frame`.
bool Rendering::RenderProgressive(const Renderer &renderer, int frame) {
  if (!camera || !visualization || visualization->parts.empty()) {
    std::cerr << "Rendering: camera and a committed visualization are needed\n";
    return false;
  }
  const gxy_camera cam = camera->AsABI();
  memset(&stats, 0, sizeof stats);
  if (!check_abi(gxy_render_progressive((int)visualization->parts.size(), visualization->parts.data(), &cam, &visualization->lighting.abi, width,
                                        height, renderer.epsilon, frame, &stats),
                 "gxy_render_progressive"))
    return false;
  rgba8.resize((size_t)width * height * 4);
  return check_abi(gxy_progressive_download_rgba8(visualization->parts[0], rgba8.data()), "gxy_progressive_download_rgba8");
}

std::string Rendering::ImageName(const std::string &base, int index) const {
  const std::string va = visualization ? visualization->annotation : std::string(""), ca = camera ? camera->annotation : std::string("");
  if (!va.empty() || !ca.empty()) return base + va + ca + ".png";
  char istr[16];
  snprintf(istr, sizeof istr, "%05d", index);
  return base + '_' + istr + va + ca + ".png";
}

bool Rendering::SaveImage(const std::string &base, int index) const {
  return write_png(ImageName(base, index), width, height, rgba8.data());
}

// ---- PNG -----------------------------------------------------------------------------------------
// RGBA8, rows top-down (what ColorImageWriter hands to write_png, ImageWriter.cpp:30-67; mypng.cpp:44-95)
bool write_png(const std::string &path, int w, int h, const unsigned char *rgba) {
  std::vector<unsigned char> raw((size_t)h * ((size_t)w * 4 + 1));
  for (int y = 0; y < h; y++) {
    raw[(size_t)y * ((size_t)w * 4 + 1)] = 0;  // filter type None
    memcpy(&raw[(size_t)y * ((size_t)w * 4 + 1) + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) {
    std::cerr << "cannot write " << path << "\n";
    return false;
  }
  auto be32 = [](unsigned char *p, unsigned v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; };
  auto chunk = [&](const char *tag, const unsigned char *data, size_t n) {
    unsigned char hdr[8];
    be32(hdr, (unsigned)n);
    memcpy(hdr + 4, tag, 4);
    fwrite(hdr, 1, 8, f);
    if (n) fwrite(data, 1, n, f);
    uLong c = crc32(0L, reinterpret_cast<const Bytef *>(tag), 4);
    if (n) c = crc32(c, data, (uInt)n);
    unsigned char tail[4];
    be32(tail, (unsigned)c);
    fwrite(tail, 1, 4, f);
  };
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  fwrite(sig, 1, 8, f);
  unsigned char ihdr[13];
  be32(ihdr, (unsigned)w);
  be32(ihdr + 4, (unsigned)h);
  ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;  // 8 bit, RGBA
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", z.data(), (size_t)zlen);
  chunk("IEND", nullptr, 0);
  const bool ok = !ferror(f);
  fclose(f);
  return ok;
}

// ---- description ---------------------------------------------------------------------------------
static void put_floats(std::ostringstream &o, const float *v, size_t n) {
  o << "[";
  for (size_t i = 0; i < n; i++) {
    char b[40];
    snprintf(b, sizeof b, "%.9g", (double)v[i]);
    o << (i ? ", " : "") << b;
  }
  o << "]";
}
static void put_ints(std::ostringstream &o, const int *v, size_t n) {
  o << "[";
  for (size_t i = 0; i < n; i++) o << (i ? ", " : "") << v[i];
  o << "]";
}
static std::string quoted(const std::string &s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') o.push_back('\\');
    o.push_back(c);
  }
  return o + "\"";
}

std::string describe_state(const Renderer &r, const std::vector<Camera> &cams, const std::vector<Visualization> &vis, const Datasets &ds,
                           int nparts) {
  std::ostringstream o;
  char b[40];
  snprintf(b, sizeof b, "%.9g", (double)r.epsilon);
  o << "{\"epsilon\": " << b << ", \"cameras\": [";
  for (size_t i = 0; i < cams.size(); i++) {
    const Camera &c = cams[i];
    o << (i ? ", " : "") << "{\"eye\": ";
    put_floats(o, c.eye, 3);
    o << ", \"dir\": ";
    put_floats(o, c.dir, 3);
    o << ", \"up\": ";
    put_floats(o, c.up, 3);
    snprintf(b, sizeof b, "%.9g", (double)c.aov);
    o << ", \"aov\": " << b << ", \"annotation\": " << quoted(c.annotation) << "}";
  }
  o << "], \"visualizations\": [";
  for (size_t i = 0; i < vis.size(); i++) {
    const Visualization &v = vis[i];
    const gxy_lighting &L = v.lighting.abi;
    o << (i ? ", " : "") << "{\"annotation\": " << quoted(v.annotation) << ", \"lighting\": {\"lights\": [";
    for (int k = 0; k < L.n_lights; k++) {
      o << (k ? ", " : "");
      put_floats(o, L.lights[k], 3);
    }
    o << "], \"types\": ";
    put_ints(o, L.types, (size_t)L.n_lights);
    o << ", \"n_ao\": " << L.n_ao << ", \"ao_radius\": ";
    put_floats(o, &L.ao_radius, 1);
    o << ", \"shadows\": " << (L.shadows ? "true" : "false") << ", \"Ka\": ";
    put_floats(o, &L.Ka, 1);
    o << ", \"Kd\": ";
    put_floats(o, &L.Kd, 1);
    o << "}, \"operators\": [";
    for (size_t k = 0; k < v.operators.size(); k++) {
      const Vis &op = v.operators[k];
      gxy_transfer_function tf;
      gxy_resample_transfer_function((int)op.colormap.size() / 4, op.colormap.data(), (int)op.opacitymap.size() / 2, op.opacitymap.data(), &tf);
      o << (k ? ", " : "") << "{\"type\": " << quoted(op.type) << ", \"dataset\": " << quoted(op.dataset) << ", \"isovalues\": ";
      put_floats(o, op.isovalues.data(), op.isovalues.size());
      o << ", \"slices\": ";
      put_floats(o, op.slices.data(), op.slices.size());
      o << ", \"tolerance\": ";
      put_floats(o, &op.tolerance, 1);
      o << ", \"isovalue\": ";
      put_floats(o, &op.isovalue, 1);
      o << ", \"radii\": ";
      const float rr[4] = {op.radius0, op.radius1, op.value0, op.value1};
      put_floats(o, rr, 4);
      o << ", \"volume_render\": " << (op.volume_render ? "true" : "false") << ", \"range\": ";
      const float rg[2] = {op.has_range ? op.range[0] : op.colormap[0], op.has_range ? op.range[1] : op.colormap[op.colormap.size() - 4]};
      put_floats(o, rg, 2);
      o << ", \"tf_colors\": ";
      put_floats(o, &tf.colors[0][0], 256 * 3);
      o << ", \"tf_opacities\": ";
      put_floats(o, tf.opacities, 256);
      o << "}";
    }
    o << "]}";
  }
  o << "], \"datasets\": [";
  for (size_t i = 0; i < ds.volumes.size(); i++) {
    const Volume &v = ds.volumes[i];
    o << (i ? ", " : "") << "{\"name\": " << quoted(v.name) << ", \"type\": " << quoted(v.is_float ? "float" : "uchar") << ", \"origin\": ";
    put_floats(o, v.origin, 3);
    o << ", \"counts\": ";
    put_ints(o, v.counts, 3);
    o << ", \"deltas\": ";
    put_floats(o, v.deltas, 3);
    o << ", \"partitions\": [";
    int f[3];
    gxy_factor(nparts, f);
    std::vector<int> table((size_t)nparts * 15);
    gxy_partition(nparts, f, v.counts, table.data());
    for (int r2 = 0; r2 < nparts; r2++) {
      VolumePart p;
      const int *t = &table[(size_t)r2 * 15];
      for (int a = 0; a < 3; a++) { p.ijk[a] = t[a]; p.offsets[a] = t[3 + a]; p.counts[a] = t[6 + a]; p.goffsets[a] = t[9 + a]; p.gcounts[a] = t[12 + a]; }
      float gmin[3], gmax[3], lmin[3], lmax[3];
      int nb[6];
      v.Boxes(p, gmin, gmax, lmin, lmax);
      Volume::Neighbors(p, f, nb);
      o << (r2 ? ", " : "") << "{\"gmin\": ";
      put_floats(o, gmin, 3);
      o << ", \"gmax\": ";
      put_floats(o, gmax, 3);
      o << ", \"lmin\": ";
      put_floats(o, lmin, 3);
      o << ", \"lmax\": ";
      put_floats(o, lmax, 3);
      o << ", \"neighbors\": ";
      put_ints(o, nb, 6);
      o << ", \"goffsets\": ";
      put_ints(o, p.goffsets, 3);
      o << ", \"gcounts\": ";
      put_ints(o, p.gcounts, 3);
      o << "}";
    }
    o << "]}";
  }
  o << "], \"geometries\": [";
  auto fnv = [](const void *data, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *b = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
  };
  for (size_t i = 0; i < ds.geometries.size(); i++) {
    const Geometry &g = ds.geometries[i];
    o << (i ? ", " : "") << "{\"name\": " << quoted(g.name) << ", \"type\": " << quoted(g.type) << ", \"parts\": [";
    for (int r2 = 0; r2 < g.NumberOfParts(); r2++) {
      float gmin[3], gmax[3], lmin[3], lmax[3];
      int nb[6];
      g.Boxes(r2, gmin, gmax, lmin, lmax, nb);
      GeometryPiece piece;
      const bool ok = g.LoadPiece(r2, piece);
      o << (r2 ? ", " : "") << "{\"loaded\": " << (ok ? "true" : "false") << ", \"gmin\": ";
      put_floats(o, gmin, 3);
      o << ", \"gmax\": ";
      put_floats(o, gmax, 3);
      o << ", \"lmin\": ";
      put_floats(o, lmin, 3);
      o << ", \"lmax\": ";
      put_floats(o, lmax, 3);
      o << ", \"neighbors\": ";
      put_ints(o, nb, 6);
      o << ", \"n_vertices\": " << piece.vertices.size() / 3 << ", \"n_connectivity\": " << piece.connectivity.size();
      o << ", \"hash_vertices\": \"" << fnv(piece.vertices.data(), piece.vertices.size() * 4) << "\"";
      o << ", \"hash_normals\": \"" << fnv(piece.normals.data(), piece.normals.size() * 4) << "\"";
      o << ", \"hash_data\": \"" << fnv(piece.data.data(), piece.data.size() * 4) << "\"";
      o << ", \"hash_connectivity\": \"" << fnv(piece.connectivity.data(), piece.connectivity.size() * 4) << "\"}";
    }
    o << "]}";
  }
  o << "]}";
  return o.str();
}

}  // namespace gxy
