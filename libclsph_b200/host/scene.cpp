// scene.cpp -- Wavefront OBJ loading and host-side face normals for the collision scene.
//
// Behaviour follows the reference's libclsph/scene.cpp:9-67 on top of its vendored loader
// (util/tinyobj, v0.9.6), restated rather than reused:
//   * faces are fan-triangulated; a vertex of the output mesh is one distinct (v, vt, vn) index
//     triple, numbered in order of first use inside its shape (tiny_obj_loader.cc:155-194);
//   * a new shape starts at every `o` / `g` record that follows faces; like the reference, the
//     scene keeps the LAST shape's vertices / indices while normals accumulate over all shapes
//     (erratum E10; every shipped scene has exactly one shape);
//   * normal = (v1 - v0) x (v2 - v0), normalised with an fp32 divide by the length.
// The collision kernels only observe triangles (corner positions, order) and normals.
#include "scene.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <tuple>

namespace {

struct obj_corner {
  int v, vt, vn;
  bool operator<(const obj_corner& o) const { return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn); }
};

struct obj_shape {
  std::vector<float> positions;
  std::vector<unsigned int> indices;
};

// "7", "7/3", "7//2", "7/3/2"; negative = relative to the end. Returns 0-based indices, -1 if absent.
obj_corner parse_corner(const std::string& tok, int n_v, int n_vt, int n_vn) {
  obj_corner c = {-1, -1, -1};
  int* slot[3] = {&c.v, &c.vt, &c.vn};
  const int count[3] = {n_v, n_vt, n_vn};
  size_t at = 0;
  for (int k = 0; k < 3 && at <= tok.size(); ++k) {
    size_t slash = tok.find('/', at);
    std::string part = tok.substr(at, slash == std::string::npos ? std::string::npos : slash - at);
    if (!part.empty()) {
      int idx = std::atoi(part.c_str());
      *slot[k] = idx > 0 ? idx - 1 : (idx < 0 ? count[k] + idx : -1);
    }
    if (slash == std::string::npos) break;
    at = slash + 1;
  }
  return c;
}

void flush_shape(std::vector<std::vector<obj_corner>>& faces, const std::vector<float>& v, std::vector<obj_shape>& out) {
  if (faces.empty()) return;
  obj_shape s;
  std::map<obj_corner, unsigned int> seen;
  auto vertex_of = [&](const obj_corner& c) -> unsigned int {
    auto it = seen.find(c);
    if (it != seen.end()) return it->second;
    unsigned int id = static_cast<unsigned int>(s.positions.size() / 3);
    for (int k = 0; k < 3; ++k) s.positions.push_back(v[3 * static_cast<size_t>(c.v) + k]);
    seen[c] = id;
    return id;
  };
  for (const auto& f : faces)
    for (size_t k = 2; k < f.size(); ++k) {  // fan: (0, k-1, k)
      s.indices.push_back(vertex_of(f[0]));
      s.indices.push_back(vertex_of(f[k - 1]));
      s.indices.push_back(vertex_of(f[k]));
    }
  out.push_back(std::move(s));
  faces.clear();
}

bool read_obj(const std::string& path, std::vector<obj_shape>& shapes, std::string& err) {
  std::ifstream in(path.c_str());
  if (!in) {
    err = "Cannot open file [" + path + "]";
    return false;
  }
  std::vector<float> v;
  int n_vt = 0, n_vn = 0;
  std::vector<std::vector<obj_corner>> faces;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tag;
    if (!(ls >> tag) || tag[0] == '#') continue;
    if (tag == "v") {
      float x = 0.f, y = 0.f, z = 0.f;
      ls >> x >> y >> z;
      v.push_back(x); v.push_back(y); v.push_back(z);
    } else if (tag == "vt") {
      ++n_vt;
    } else if (tag == "vn") {
      ++n_vn;
    } else if (tag == "f") {
      std::vector<obj_corner> face;
      std::string tok;
      while (ls >> tok) {
        obj_corner c = parse_corner(tok, static_cast<int>(v.size() / 3), n_vt, n_vn);
        if (c.v < 0 || static_cast<size_t>(c.v) >= v.size() / 3) {
          err = "Face refers to a vertex that does not exist in [" + path + "]";
          return false;
        }
        face.push_back(c);
      }
      if (face.size() >= 3) faces.push_back(face);
    } else if (tag == "o" || tag == "g") {
      flush_shape(faces, v, shapes);
    }
  }
  flush_shape(faces, v, shapes);
  return true;
}

}  // namespace

bool scene::load(std::string filename) {
  std::vector<obj_shape> shapes;
  std::string err;
  if (!read_obj(std::string("scenes/") + filename, shapes, err)) {
    std::cerr << err << std::endl;
    return false;
  }
  std::cout << "Scene Loading - number of shapes in file [" << filename << "]: " << shapes.size() << std::endl;

  for (size_t s = 0; s < shapes.size(); ++s) {
    indices = shapes[s].indices;
    if (indices.size() % 3 != 0) {
      std::cerr << "Meshes must be made of triangles only" << std::endl;
      return false;
    }
    face_count = static_cast<unsigned int>(indices.size() / 3);
    vertices = shapes[s].positions;
    for (size_t f = 0; f < face_count; ++f) {
      const float* a = &vertices[3 * indices[3 * f + 0]];
      const float* b = &vertices[3 * indices[3 * f + 1]];
      const float* c = &vertices[3 * indices[3 * f + 2]];
      const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
      const float e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      const float nx = e1[1] * e2[2] - e1[2] * e2[1];
      const float ny = e1[2] * e2[0] - e1[0] * e2[2];
      const float nz = e1[0] * e2[1] - e1[1] * e2[0];
      const float len = static_cast<float>(std::sqrt(static_cast<double>(nx * nx + ny * ny + nz * nz)));
      face_normals.push_back(nx / len);
      face_normals.push_back(ny / len);
      face_normals.push_back(nz / len);
    }
  }
  return true;
}
