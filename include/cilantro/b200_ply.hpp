// PLY passthrough for PointCloud3f (host-only I/O; the reference delegates to the vendored tinyply:
// utilities/ply_io.hpp:43-106, utilities/point_cloud.hpp:502-543). Self-contained reader / writer of the
// subset cilantro uses: element "vertex" with scalar properties x y z [nx ny nz] [red green blue]; ascii,
// binary_little_endian and binary_big_endian; any scalar property type (cast to float like
// vectorSetFromPLYDataBuffer); other elements and list properties are parsed and skipped.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cilantro {
namespace b200 {
namespace ply {

struct Property {
  std::string name;
  int type = -1;        // index into kTypes
  bool is_list = false;
  int count_type = -1;  // list length type
};
struct Element {
  std::string name;
  size_t count = 0;
  std::vector<Property> props;
};

inline int type_index(const std::string& t) {
  static const char* names[8][2] = {{"char", "int8"},   {"uchar", "uint8"}, {"short", "int16"}, {"ushort", "uint16"},
                                    {"int", "int32"},   {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
  for (int i = 0; i < 8; i++)
    if (t == names[i][0] || t == names[i][1]) return i;
  throw std::runtime_error("PLY: unknown property type '" + t + "'");
}
inline size_t type_size(int t) {
  static const size_t s[8] = {1, 1, 2, 2, 4, 4, 4, 8};
  return s[t];
}

inline double decode(const unsigned char* p, int t, bool swap) {
  unsigned char b[8];
  const size_t n = type_size(t);
  for (size_t i = 0; i < n; i++) b[i] = swap ? p[n - 1 - i] : p[i];
  switch (t) {
    case 0: { int8_t v; std::memcpy(&v, b, 1); return v; }
    case 1: { uint8_t v; std::memcpy(&v, b, 1); return v; }
    case 2: { int16_t v; std::memcpy(&v, b, 2); return v; }
    case 3: { uint16_t v; std::memcpy(&v, b, 2); return v; }
    case 4: { int32_t v; std::memcpy(&v, b, 4); return v; }
    case 5: { uint32_t v; std::memcpy(&v, b, 4); return v; }
    case 6: { float v; std::memcpy(&v, b, 4); return v; }
    default: { double v; std::memcpy(&v, b, 8); return v; }
  }
}

// Reads the vertex element: fills the 3 x N arrays (packed xyz) that are present in the file.
inline void read(const std::string& file_name, std::vector<float>& points, std::vector<float>& normals,
                 std::vector<float>& colors) {
  points.clear();
  normals.clear();
  colors.clear();
  std::ifstream in(file_name, std::ios::binary);
  if (!in) throw std::runtime_error("PLY: cannot open '" + file_name + "'");
  std::string line;
  std::getline(in, line);
  if (line.substr(0, 3) != "ply") throw std::runtime_error("PLY: missing magic in '" + file_name + "'");
  int format = -1;  // 0 ascii, 1 little endian, 2 big endian
  std::vector<Element> elements;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string tok;
    ls >> tok;
    if (tok == "format") {
      ls >> tok;
      format = tok == "ascii" ? 0 : (tok == "binary_little_endian" ? 1 : (tok == "binary_big_endian" ? 2 : -1));
    } else if (tok == "element") {
      Element e;
      ls >> e.name >> e.count;
      elements.push_back(e);
    } else if (tok == "property") {
      if (elements.empty()) throw std::runtime_error("PLY: property before element");
      Property p;
      ls >> tok;
      if (tok == "list") {
        std::string ct, vt;
        ls >> ct >> vt >> p.name;
        p.is_list = true;
        p.count_type = type_index(ct);
        p.type = type_index(vt);
      } else {
        p.type = type_index(tok);
        ls >> p.name;
      }
      elements.back().props.push_back(p);
    } else if (tok == "end_header") {
      break;
    }
  }
  if (format < 0) throw std::runtime_error("PLY: unsupported or missing format line");
  uint16_t one = 1;
  const bool host_little = *reinterpret_cast<unsigned char*>(&one) == 1;
  const bool swap = (format == 1 && !host_little) || (format == 2 && host_little);
  for (const Element& e : elements) {
    const bool is_vertex = e.name == "vertex";
    int slot[9];  // property index of x y z nx ny nz red green blue
    for (int& s : slot) s = -1;
    static const char* want[9] = {"x", "y", "z", "nx", "ny", "nz", "red", "green", "blue"};
    if (is_vertex)
      for (size_t k = 0; k < e.props.size(); k++)
        for (int w = 0; w < 9; w++)
          if (!e.props[k].is_list && e.props[k].name == want[w]) slot[w] = (int)k;
    const bool has_p = slot[0] >= 0 && slot[1] >= 0 && slot[2] >= 0;
    const bool has_n = slot[3] >= 0 && slot[4] >= 0 && slot[5] >= 0;
    const bool has_c = slot[6] >= 0 && slot[7] >= 0 && slot[8] >= 0;
    if (is_vertex) {
      if (has_p) points.resize(3 * e.count);
      if (has_n) normals.resize(3 * e.count);
      if (has_c) colors.resize(3 * e.count);
    }
    std::vector<double> row(e.props.size());
    for (size_t r = 0; r < e.count; r++) {
      for (size_t k = 0; k < e.props.size(); k++) {
        const Property& p = e.props[k];
        if (format == 0) {
          if (p.is_list) {
            double cnt = 0, v;
            in >> cnt;
            for (long i = 0; i < (long)cnt; i++) in >> v;
          } else {
            in >> row[k];
          }
        } else {
          unsigned char buf[8];
          if (p.is_list) {
            in.read((char*)buf, (std::streamsize)type_size(p.count_type));
            const long cnt = (long)decode(buf, p.count_type, swap);
            in.ignore((std::streamsize)(cnt * (long)type_size(p.type)));
          } else {
            in.read((char*)buf, (std::streamsize)type_size(p.type));
            row[k] = decode(buf, p.type, swap);
          }
        }
      }
      if (!in) throw std::runtime_error("PLY: unexpected end of data in '" + file_name + "'");
      if (is_vertex) {
        for (int a = 0; a < 3; a++) {
          if (has_p) points[3 * r + a] = (float)row[slot[a]];
          if (has_n) normals[3 * r + a] = (float)row[slot[3 + a]];
          if (has_c) colors[3 * r + a] = (1.0f / 255.0f) * (float)row[slot[6 + a]];  // point_cloud.hpp:513
        }
      }
    }
  }
}

// Writes float x y z [nx ny nz] and uchar red green blue (255 * colour, truncated) — point_cloud.hpp:520-543.
inline void write(const std::string& file_name, bool binary, size_t n, const float* points, const float* normals,
                  const float* colors) {
  std::ofstream out(file_name, std::ios::binary);
  if (!out) throw std::runtime_error("PLY: cannot create '" + file_name + "'");
  uint16_t one = 1;
  const bool host_little = *reinterpret_cast<unsigned char*>(&one) == 1;
  out << "ply\nformat " << (binary ? (host_little ? "binary_little_endian" : "binary_big_endian") : "ascii") << " 1.0\n";
  out << "element vertex " << n << "\n";
  out << "property float x\nproperty float y\nproperty float z\n";
  if (normals) out << "property float nx\nproperty float ny\nproperty float nz\n";
  if (colors) out << "property uchar red\nproperty uchar green\nproperty uchar blue\n";
  out << "end_header\n";
  if (!binary) out.precision(9);
  for (size_t i = 0; i < n; i++) {
    unsigned char rgb[3] = {0, 0, 0};
    if (colors)
      for (int a = 0; a < 3; a++) rgb[a] = static_cast<unsigned char>(255.0f * colors[3 * i + a]);
    if (binary) {
      out.write((const char*)(points + 3 * i), 12);
      if (normals) out.write((const char*)(normals + 3 * i), 12);
      if (colors) out.write((const char*)rgb, 3);
    } else {
      out << points[3 * i] << ' ' << points[3 * i + 1] << ' ' << points[3 * i + 2];
      if (normals) out << ' ' << normals[3 * i] << ' ' << normals[3 * i + 1] << ' ' << normals[3 * i + 2];
      if (colors) out << ' ' << (int)rgb[0] << ' ' << (int)rgb[1] << ' ' << (int)rgb[2];
      out << '\n';
    }
  }
  if (!out) throw std::runtime_error("PLY: write failed for '" + file_name + "'");
}

}  // namespace ply
}  // namespace b200
}  // namespace cilantro
