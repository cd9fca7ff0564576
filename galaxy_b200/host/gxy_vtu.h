// gxy_vtu.h -- a VTK-free reader for the VTK XML UnstructuredGrid / PolyData files (.vtu / .vtp) that Galaxy's geometry
// datasets are stored in.  The reference reads them with vtkXMLUnstructuredGridReader and then copies, from the vtkPointSet,
// the float points, the point-data arrays "Normals"/"Normals_", the active scalars (or the array named "data") and the cell
// connectivity (src/data/Geometry.cpp:176-257, Triangles.cpp:79-146, Particles.cpp:90-129).  This reader extracts exactly
// those arrays.  Supported encodings of a <DataArray>: format="ascii", format="binary" (inline base64) and format="appended"
// (<AppendedData encoding="raw" | "base64">), uncompressed or compressor="vtkZLibDataCompressor", header_type UInt32 or
// UInt64, byte_order LittleEndian; value types Float32/Float64 and Int8..Int64/UInt8..UInt64.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace gxy {

struct VtuData {
  std::vector<float> points;         // 3 per point
  std::vector<float> normals;        // 3 per point, empty if the file has none
  std::vector<float> scalars;        // 1 per point (doubles converted as the reference does), empty if none
  std::vector<int> connectivity;     // point ids of all cells, concatenated
  std::vector<long long> offsets;    // end offset of every cell in connectivity
  long long n_points = 0, n_cells = 0;
};

// returns false and fills `error` on failure; never exits
bool read_vtu(const std::string &path, VtuData &out, std::string &error);

}  // namespace gxy
