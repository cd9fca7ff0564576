// gxy_vtu.cpp -- see gxy_vtu.h.  File layout per "VTK File Formats" (XML formats, <DataArray> encodings): binary blocks are
// [header][data]; header = one word (byte count) when uncompressed, [#blocks][block size][last block size][c-size ...] when
// compressed with vtkZLibDataCompressor; inline/appended-base64 encode the header as its own base64 unit.
#include "gxy_vtu.h"

#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <new>

namespace gxy {
namespace {

struct Tag {
  std::string name;                          // "DataArray", "/DataArray", "Points", ...
  std::map<std::string, std::string> attrs;
  bool self_closing = false;
  size_t begin = 0, end = 0;                 // [begin, end) of the tag text in the file
};

// next tag at or after pos; false at end of text
bool next_tag(const std::string &s, size_t pos, size_t limit, Tag &t) {
  while (true) {
    const size_t lt = s.find('<', pos);
    if (lt == std::string::npos || lt >= limit) return false;
    if (s.compare(lt, 4, "<!--") == 0) {
      const size_t e = s.find("-->", lt);
      if (e == std::string::npos) return false;
      pos = e + 3;
      continue;
    }
    if (lt + 1 < s.size() && (s[lt + 1] == '?' || s[lt + 1] == '!')) {
      const size_t e = s.find('>', lt);
      if (e == std::string::npos) return false;
      pos = e + 1;
      continue;
    }
    const size_t gt = s.find('>', lt);
    if (gt == std::string::npos) return false;
    t = Tag();
    t.begin = lt;
    t.end = gt + 1;
    size_t p = lt + 1;
    while (p < gt && !isspace((unsigned char)s[p]) && s[p] != '/' ) p++;
    if (s[lt + 1] == '/') {  // closing tag
      t.name = s.substr(lt + 1, gt - lt - 1);
      while (!t.name.empty() && isspace((unsigned char)t.name.back())) t.name.pop_back();
      return true;
    }
    t.name = s.substr(lt + 1, p - lt - 1);
    t.self_closing = gt > lt && s[gt - 1] == '/';
    // attributes
    while (p < gt) {
      while (p < gt && (isspace((unsigned char)s[p]) || s[p] == '/')) p++;
      if (p >= gt) break;
      size_t eq = s.find('=', p);
      if (eq == std::string::npos || eq >= gt) break;
      std::string key = s.substr(p, eq - p);
      while (!key.empty() && isspace((unsigned char)key.back())) key.pop_back();
      size_t q = eq + 1;
      while (q < gt && isspace((unsigned char)s[q])) q++;
      if (q >= gt || (s[q] != '"' && s[q] != '\'')) break;
      const char quote = s[q];
      const size_t qe = s.find(quote, q + 1);
      if (qe == std::string::npos || qe > gt) break;
      t.attrs[key] = s.substr(q + 1, qe - q - 1);
      p = qe + 1;
    }
    return true;
  }
}

int b64_value(unsigned char c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}

// decode one base64 unit starting at text[pos]: stops after a padded quantum or when max_bytes are produced or at a
// non-base64 character; advances pos past what it consumed
std::vector<unsigned char> b64_decode(const std::string &text, size_t &pos, size_t end, size_t max_bytes) {
  std::vector<unsigned char> out;
  while (pos < end && out.size() < max_bytes) {
    int v[4], n = 0, pad = 0;
    size_t p = pos;
    while (n < 4 && p < end) {
      const unsigned char c = (unsigned char)text[p];
      if (isspace(c)) { p++; continue; }
      if (c == '=') { v[n++] = 0; pad++; p++; continue; }
      const int b = b64_value(c);
      if (b < 0) break;
      v[n++] = b;
      p++;
    }
    if (n < 4) break;
    pos = p;
    out.push_back((unsigned char)((v[0] << 2) | (v[1] >> 4)));
    if (pad < 2) out.push_back((unsigned char)(((v[1] & 15) << 4) | (v[2] >> 2)));
    if (pad < 1) out.push_back((unsigned char)(((v[2] & 3) << 6) | v[3]));
    if (pad) break;
  }
  if (out.size() > max_bytes) out.resize(max_bytes);
  return out;
}

size_t b64_chars(size_t nbytes) { return (nbytes + 2) / 3 * 4; }

struct Ctx {
  bool header64 = false, compressed = false;
  std::string error;
};

unsigned long long header_word(const unsigned char *p, bool h64) {
  if (h64) { uint64_t v; memcpy(&v, p, 8); return v; }
  uint32_t v; memcpy(&v, p, 4); return v;
}

bool inflate_blocks(const unsigned char *blocks, size_t avail, const std::vector<unsigned long long> &csizes, unsigned long long block_size,
                    unsigned long long last_size, std::vector<unsigned char> &out, std::string &err) {
  size_t off = 0;
  out.clear();
  // sizes come from the file: compare by subtraction (no wrap-around) and bound what a block may inflate to (zlib cannot expand
  // beyond ~1032x; VTK writes blocks of 32 KiB .. a few MiB)
  const unsigned long long max_block = 1ull << 30;
  if (block_size > max_block || last_size > max_block) { err = "compressed block size field out of range"; return false; }
  for (size_t k = 0; k < csizes.size(); k++) {
    if (csizes[k] > avail - off) { err = "compressed block runs past the end of the data"; return false; }
    const unsigned long long usz = (k + 1 == csizes.size() && last_size) ? last_size : block_size;
    if (usz / 1100 > csizes[k] + 1) { err = "compressed block claims an impossible inflated size"; return false; }
    const size_t at = out.size();
    out.resize(at + (size_t)usz);
    uLongf dl = (uLongf)usz;
    if (uncompress(out.data() + at, &dl, blocks + off, (uLong)csizes[k]) != Z_OK) { err = "zlib: corrupt block"; return false; }
    out.resize(at + dl);
    off += (size_t)csizes[k];
  }
  return true;
}

// raw bytes of a binary DataArray stored as base64 text [pos, end)
bool bytes_from_base64(const std::string &text, size_t pos, size_t end, const Ctx &c, std::vector<unsigned char> &out, std::string &err) {
  const size_t hw = c.header64 ? 8 : 4;
  while (pos < end && isspace((unsigned char)text[pos])) pos++;
  if (!c.compressed) {
    size_t p = pos;
    std::vector<unsigned char> first = b64_decode(text, p, end, (size_t)-1);  // up to the first padded quantum or the end
    if (first.size() < hw) { err = "truncated base64 data"; return false; }
    const unsigned long long nbytes = header_word(first.data(), c.header64);
    if (first.size() == hw && nbytes > 0) {  // the header was encoded as its own unit: the data follow
      out = b64_decode(text, p, end, (size_t)-1);
    } else {
      out.assign(first.begin() + hw, first.end());
    }
    if (out.size() < nbytes) { err = "base64 data shorter than its header says"; return false; }
    out.resize((size_t)nbytes);
    return true;
  }
  size_t p = pos;
  std::vector<unsigned char> h1 = b64_decode(text, p, end, hw);
  if (h1.size() < hw) { err = "truncated compressed header"; return false; }
  const unsigned long long nblocks = header_word(h1.data(), c.header64);
  if (nblocks > (end - pos) / hw) { err = "compressed header claims more blocks than the data can hold"; return false; }
  const size_t hbytes = hw * (3 + (size_t)nblocks);
  p = pos;
  std::vector<unsigned char> hdr = b64_decode(text, p, std::min(end, pos + b64_chars(hbytes)), hbytes);
  if (hdr.size() < hbytes) { err = "truncated compressed header"; return false; }
  std::vector<unsigned long long> cs(nblocks);
  for (size_t k = 0; k < nblocks; k++) cs[k] = header_word(hdr.data() + hw * (3 + k), c.header64);
  p = pos + b64_chars(hbytes);
  std::vector<unsigned char> blocks = b64_decode(text, p, end, (size_t)-1);
  return inflate_blocks(blocks.data(), blocks.size(), cs, header_word(hdr.data() + hw, c.header64), header_word(hdr.data() + 2 * hw, c.header64), out,
                        err);
}

bool bytes_from_raw(const std::string &file, size_t pos, const Ctx &c, std::vector<unsigned char> &out, std::string &err) {
  const size_t hw = c.header64 ? 8 : 4;
  const unsigned char *base = reinterpret_cast<const unsigned char *>(file.data());
  if (pos > file.size() || hw > file.size() - pos) { err = "appended data offset past the end of the file"; return false; }
  if (!c.compressed) {
    const unsigned long long nbytes = header_word(base + pos, c.header64);
    if (nbytes > file.size() - pos - hw) { err = "appended array runs past the end of the file"; return false; }
    out.assign(base + pos + hw, base + pos + hw + nbytes);
    return true;
  }
  const unsigned long long nblocks = header_word(base + pos, c.header64);
  if (nblocks > (file.size() - pos) / hw) { err = "appended compressed header past the end of the file"; return false; }
  const size_t hbytes = hw * (3 + (size_t)nblocks);
  if (hbytes > file.size() - pos) { err = "appended compressed header past the end of the file"; return false; }
  std::vector<unsigned long long> cs(nblocks);
  for (size_t k = 0; k < nblocks; k++) cs[k] = header_word(base + pos + hw * (3 + k), c.header64);
  return inflate_blocks(base + pos + hbytes, file.size() - pos - hbytes, cs, header_word(base + pos + hw, c.header64),
                        header_word(base + pos + 2 * hw, c.header64), out, err);
}

// numeric array of a DataArray as doubles or 64-bit ints, whatever its stored type
struct Numbers {
  std::vector<double> f;
  std::vector<long long> i;
  bool is_float = false;
};

bool numbers_from_bytes(const std::vector<unsigned char> &b, const std::string &type, Numbers &n, std::string &err) {
  const unsigned char *p = b.data();
  auto conv = [&](auto tag, bool is_float) {
    using T = decltype(tag);
    const size_t cnt = b.size() / sizeof(T);
    n.is_float = is_float;
    if (is_float) n.f.resize(cnt); else n.i.resize(cnt);
    for (size_t k = 0; k < cnt; k++) {
      T v;
      memcpy(&v, p + k * sizeof(T), sizeof(T));
      if (is_float) n.f[k] = (double)v; else n.i[k] = (long long)v;
    }
  };
  if (type == "Float32") conv(float(), true);
  else if (type == "Float64") conv(double(), true);
  else if (type == "Int8") conv(int8_t(), false);
  else if (type == "UInt8") conv(uint8_t(), false);
  else if (type == "Int16") conv(int16_t(), false);
  else if (type == "UInt16") conv(uint16_t(), false);
  else if (type == "Int32") conv(int32_t(), false);
  else if (type == "UInt32") conv(uint32_t(), false);
  else if (type == "Int64") conv(int64_t(), false);
  else if (type == "UInt64") conv(uint64_t(), false);
  else { err = "unsupported DataArray type " + type; return false; }
  return true;
}

bool numbers_from_ascii(const std::string &text, size_t pos, size_t end, const std::string &type, Numbers &n) {
  n.is_float = type == "Float32" || type == "Float64";
  const char *p = text.c_str() + pos, *e = text.c_str() + end;
  while (p < e) {
    while (p < e && isspace((unsigned char)*p)) p++;
    if (p >= e) break;
    char *q = nullptr;
    if (n.is_float) {
      // Float32 text is parsed as float so that the stored value is what a float reader sees
      const double v = type == "Float32" ? (double)strtof(p, &q) : strtod(p, &q);
      if (q == p) break;
      n.f.push_back(v);
    } else {
      const long long v = strtoll(p, &q, 10);
      if (q == p) break;
      n.i.push_back(v);
    }
    p = q;
  }
  return true;
}

}  // namespace

static bool read_vtu_impl(const std::string &path, VtuData &out, std::string &error) {
  std::ifstream in(path.c_str(), std::ios::in | std::ios::binary);
  if (!in) { error = "cannot open " + path; return false; }
  std::stringstream ss;
  ss << in.rdbuf();
  const std::string file = ss.str();
  out = VtuData();

  // the XML ends (for scanning purposes) where the appended blob begins: raw data may contain '<'
  size_t xml_limit = file.size(), appended_data = std::string::npos;
  bool appended_raw = true;
  {
    const size_t ad = file.find("<AppendedData");
    if (ad != std::string::npos) {
      Tag t;
      if (!next_tag(file, ad, file.size(), t)) { error = "malformed <AppendedData>"; return false; }
      appended_raw = t.attrs.count("encoding") ? t.attrs["encoding"] == "raw" : true;
      const size_t us = file.find('_', t.end);
      if (us == std::string::npos) { error = "<AppendedData> without '_' marker"; return false; }
      appended_data = us + 1;
      xml_limit = ad;
    }
  }
  size_t appended_end = file.size();
  if (appended_data != std::string::npos && !appended_raw) {
    const size_t close = file.find("</AppendedData>", appended_data);
    if (close != std::string::npos) appended_end = close;
  }

  Ctx ctx;
  std::string section, active_scalars;
  int pieces = 0;
  size_t pos = 0;
  Tag t;
  bool have_points = false;
  while (next_tag(file, pos, xml_limit, t)) {
    pos = t.end;
    if (t.name == "VTKFile") {
      const std::string type = t.attrs.count("type") ? t.attrs["type"] : "";
      if (type != "UnstructuredGrid" && type != "PolyData") { error = "VTKFile type " + type + " is not UnstructuredGrid or PolyData"; return false; }
      if (t.attrs.count("byte_order") && t.attrs["byte_order"] != "LittleEndian") { error = "only LittleEndian files are supported"; return false; }
      ctx.header64 = t.attrs.count("header_type") && t.attrs["header_type"] == "UInt64";
      if (t.attrs.count("compressor")) {
        if (t.attrs["compressor"] != "vtkZLibDataCompressor") { error = "unsupported compressor " + t.attrs["compressor"]; return false; }
        ctx.compressed = true;
      }
    } else if (t.name == "Piece") {
      if (++pieces > 1) { error = "files with more than one <Piece> are not supported"; return false; }
      if (t.attrs.count("NumberOfPoints")) out.n_points = atoll(t.attrs["NumberOfPoints"].c_str());
      if (t.attrs.count("NumberOfCells")) out.n_cells = atoll(t.attrs["NumberOfCells"].c_str());
      if (t.attrs.count("NumberOfPolys")) out.n_cells = atoll(t.attrs["NumberOfPolys"].c_str());
    } else if (t.name == "Points" || t.name == "Cells" || t.name == "Polys" || t.name == "CellData" || t.name == "Verts" || t.name == "Lines" ||
               t.name == "Strips" || t.name == "FieldData") {
      section = t.self_closing ? "" : t.name;
    } else if (t.name == "PointData") {
      section = t.self_closing ? "" : "PointData";
      active_scalars = t.attrs.count("Scalars") ? t.attrs["Scalars"] : "";
    } else if (!t.name.empty() && t.name[0] == '/') {
      if (t.name == "/" + section) section = "";
    } else if (t.name == "DataArray") {
      // content of an inline array: up to its closing tag
      size_t content_begin = t.end, content_end = t.end;
      if (!t.self_closing) {
        const size_t close = file.find("</DataArray>", t.end);
        if (close == std::string::npos || close > xml_limit) { error = "unterminated <DataArray>"; return false; }
        content_end = close;
        pos = close + 12;
      }
      const std::string name = t.attrs.count("Name") ? t.attrs["Name"] : "";
      const bool want = section == "Points" || (section == "PointData") || ((section == "Cells" || section == "Polys") && (name == "connectivity" || name == "offsets"));
      if (!want) continue;
      const std::string type = t.attrs.count("type") ? t.attrs["type"] : "Float32";
      const std::string format = t.attrs.count("format") ? t.attrs["format"] : "ascii";
      Numbers nums;
      if (format == "ascii") {
        numbers_from_ascii(file, content_begin, content_end, type, nums);
      } else {
        std::vector<unsigned char> bytes;
        if (format == "binary") {
          if (!bytes_from_base64(file, content_begin, content_end, ctx, bytes, error)) return false;
        } else if (format == "appended") {
          if (appended_data == std::string::npos) { error = "appended DataArray without <AppendedData>"; return false; }
          const size_t off = t.attrs.count("offset") ? (size_t)atoll(t.attrs["offset"].c_str()) : 0;
          if (appended_raw) {
            if (!bytes_from_raw(file, appended_data + off, ctx, bytes, error)) return false;
          } else if (!bytes_from_base64(file, appended_data + off, appended_end, ctx, bytes, error)) {
            return false;
          }
        } else {
          error = "unsupported DataArray format " + format;
          return false;
        }
        if (!numbers_from_bytes(bytes, type, nums, error)) return false;
      }
      auto to_float = [&](std::vector<float> &dst) {
        if (nums.is_float) { dst.resize(nums.f.size()); for (size_t k = 0; k < nums.f.size(); k++) dst[k] = (float)nums.f[k]; }
        else { dst.resize(nums.i.size()); for (size_t k = 0; k < nums.i.size(); k++) dst[k] = (float)nums.i[k]; }
      };
      if (section == "Points") {
        if (type != "Float32") { error = "can only handle float points"; return false; }  // Triangles.cpp:92-97
        to_float(out.points);
        have_points = true;
      } else if (section == "PointData") {
        const int ncomp = t.attrs.count("NumberOfComponents") ? atoi(t.attrs["NumberOfComponents"].c_str()) : 1;
        if ((name == "Normals" || (name == "Normals_" && out.normals.empty())) && ncomp == 3 && type == "Float32") to_float(out.normals);
        else if (ncomp == 1 && ((!active_scalars.empty() && name == active_scalars) || (active_scalars.empty() && name == "data"))) {
          if (type == "Float32" || type == "Float64") to_float(out.scalars);  // other types: the reference leaves data = 0
        }
      } else if (name == "connectivity") {
        out.connectivity.resize(nums.i.size());
        for (size_t k = 0; k < nums.i.size(); k++) out.connectivity[k] = (int)nums.i[k];
      } else if (name == "offsets") {
        out.offsets = nums.i;
      }
    }
  }
  if (!pieces) { error = "no <Piece> in " + path; return false; }
  if (!have_points && out.n_points > 0) { error = "no <Points> array in " + path; return false; }
  if ((long long)out.points.size() != 3 * out.n_points) { error = "point array size does not match NumberOfPoints"; return false; }
  if (!out.normals.empty() && out.normals.size() != out.points.size()) { error = "normal array size does not match the points"; return false; }
  if (!out.scalars.empty() && (long long)out.scalars.size() != out.n_points) { error = "scalar array size does not match the points"; return false; }
  for (int id : out.connectivity)
    if (id < 0 || id >= out.n_points) { error = "connectivity refers to a point that does not exist"; return false; }
  return true;
}

// a truncated or malformed dataset is an error return ("print and return false", Geometry.cpp:176-257), never a crash: size fields
// of the file that slip through the range checks can still ask for more memory than there is
bool read_vtu(const std::string &path, VtuData &out, std::string &error) {
  try {
    return read_vtu_impl(path, out, error);
  } catch (const std::bad_alloc &) {
    error = "out of memory while reading " + path + " (corrupt size field?)";
  } catch (const std::length_error &) {
    error = "impossible array size in " + path;
  } catch (const std::out_of_range &) {
    error = "malformed " + path;
  }
  out = VtuData();
  return false;
}

}  // namespace gxy
