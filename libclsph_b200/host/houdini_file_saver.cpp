// houdini_file_saver.cpp -- ASCII Houdini .geo ("PGEOMETRY V5") particle frames.
//
// Produces the same bytes as the reference's libclsph/file_save_delegates/houdini_file_saver.cpp
// :25-92 through util/houdini_geo/HoudiniFileDumpHelper.cpp:19-90: header, per point
// "x y z 0 (vx vy vz<TAB>r g b<TAB>mass)", the single particle primitive, trailer. Numbers use
// the default iostream float format (= printf %g, 6 significant digits). Written with one
// pre-sized buffer and snprintf instead of per-field stream inserts: frame export is the largest
// wall-time term once the step runs on the GPU.
#include "file_save_delegates/houdini_file_saver.h"

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

namespace {

// Frame number as 7 digits: 9-wide zero padding cut to its last 7 characters (reference :15-23).
std::string frame_suffix(int num) {
  std::ostringstream ss;
  ss << std::setw(9) << std::setfill('0') << num;
  std::string s = ss.str();
  return s.size() > 7 ? s.substr(s.size() - 7) : s;
}

// Density -> colour ramp of the reference (:47-60).
void density_colour(float rho, float* r, float* g, float* b) {
  *r = (rho > 1000.f && rho <= 2000.f) ? (rho - 1000.f) / 1000.f : 0.f;
  *g = (rho >= 0.f && rho < 1000.f) ? 1.f - rho / 1000.f : 0.f;
  *b = (rho >= 500.f && rho <= 1000.f) ? (rho - 500.f) / 500.f
       : (rho >= 1000.f && rho <= 1500.f) ? 1.f - (rho - 1000.f) / 500.f
                                          : 0.f;
}

}  // namespace

int houdini_file_saver::writeFrameToFile(particle* particles, const simulation_parameters& parameters) {
  const std::string file_name = frames_folder_prefix + "frames/frame" + frame_suffix(++frame_count) + ".geo";
  const unsigned int n = parameters.particles_count;

  std::string out;
  out.reserve(static_cast<size_t>(n) * 96 + 512);
  char line[256];
  out += "PGEOMETRY V5\n";
  std::snprintf(line, sizeof(line), "NPoints %d NPrims 1\n", static_cast<int>(n));
  out += line;
  out += "NPointGroups 0 NPrimGroups 1\n";
  out += "NPointAttrib 3 NVertexAttrib 0 NPrimAttrib 2 NAttrib 0\n";
  out += "PointAttrib\nv 3 float 1 1 1\ncolor 3 float 1 1 1\nmass 1 float 1\n";
  for (unsigned int i = 0; i < n; ++i) {
    const particle& q = particles[i];
    float r, g, b;
    density_colour(q.density, &r, &g, &b);
    const int len = std::snprintf(line, sizeof(line), "%g %g %g %g (%g %g %g\t%g %g %g\t%g)\n",
                                  (double)q.position.s[0], (double)q.position.s[1], (double)q.position.s[2], 0.0,
                                  (double)q.velocity.s[0], (double)q.velocity.s[1], (double)q.velocity.s[2], (double)r,
                                  (double)g, (double)b, (double)parameters.particle_mass);
    out.append(line, static_cast<size_t>(len));
  }
  out += "PrimitiveAttrib\ngenerator 1 index 1 location1\ndopobject 1 index 1 /obj/AutoDopNetwork:1\n";
  std::snprintf(line, sizeof(line), "Part %d", static_cast<int>(n));
  out += line;
  for (unsigned int i = 0; i < n; ++i) {
    const int len = std::snprintf(line, sizeof(line), " %d", static_cast<int>(i));
    out.append(line, static_cast<size_t>(len));
  }
  out += " [0\t0]\nbox_object1 unordered\n1 1\nbeginExtra\nendExtra\n";

  std::ofstream file(file_name.c_str(), std::ios::out | std::ios::trunc | std::ios::binary);
  if (file.is_open()) {
    file.write(out.data(), static_cast<std::streamsize>(out.size()));
  } else {
    std::cerr << "Error while writing to " << file_name << std::endl;
  }
  return 0;
}
