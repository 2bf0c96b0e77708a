// host_capi.cpp -- C entry points onto the C++ host classes, so the Python test-suite and
// bench.py can drive sph_simulation / scene / houdini_file_saver exactly as a C++ user would.
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "file_save_delegates/houdini_file_saver.h"
#include "sph_simulation.h"

namespace {
struct quiet_cout {
  std::ostringstream sink;
  std::streambuf* saved;
  quiet_cout() : saved(std::cout.rdbuf(sink.rdbuf())) {}
  ~quiet_cout() { std::cout.rdbuf(saved); }
};
}  // namespace

extern "C" {

// sph_simulation::load_settings. Returns 0, or 1 with `error` (size 256) filled.
int clsph_host_load_settings(const char* fluid_json, const char* sim_json, simulation_parameters* p,
                             precomputed_kernel_values* t, float* initial_volume, int* write_all_frames, int* serialize,
                             char* error) {
  try {
    sph_simulation sim;
    sim.quiet = true;
    sim.load_settings(fluid_json, sim_json);
    *p = sim.parameters;
    *t = sim.precomputed_terms;
    *initial_volume = sim.initial_volume;
    *write_all_frames = sim.write_intermediate_frames;
    *serialize = sim.serialize;
    return 0;
  } catch (const std::exception& e) {
    if (error) std::strncpy(error, e.what(), 255);
    return 1;
  }
}

// scene::load (relative to the current working directory). Call with null arrays for the sizes.
int clsph_host_scene_load(const char* name, unsigned int* face_count, size_t* n_vertex_floats, size_t* n_indices,
                          float* face_normals, float* vertices, unsigned int* indices) {
  quiet_cout q;
  scene s;
  if (!s.load(name)) return 1;
  *face_count = s.face_count;
  *n_vertex_floats = s.vertices.size();
  *n_indices = s.indices.size();
  if (face_normals) std::memcpy(face_normals, s.face_normals.data(), sizeof(float) * s.face_normals.size());
  if (vertices) std::memcpy(vertices, s.vertices.data(), sizeof(float) * s.vertices.size());
  if (indices) std::memcpy(indices, s.indices.data(), sizeof(unsigned int) * s.indices.size());
  return 0;
}

// houdini_file_saver::writeFrameToFile, `frames` times in a row (frame numbering included).
int clsph_host_write_frames(const char* prefix, particle* particles, const simulation_parameters* p, int frames) {
  houdini_file_saver saver(prefix);
  for (int k = 0; k < frames; ++k) saver.writeFrameToFile(particles, *p);
  return 0;
}

// The same with the format chosen (0 geo, 1 bgeo) and, with packed != 0, through writeFramePoints from the seven
// floats per particle (position, velocity, density) that clsph_frame_begin packs.
int clsph_host_write_frames_as(const char* prefix, particle* particles, const simulation_parameters* p, int frames, int format,
                               int packed) {
  houdini_file_saver saver(prefix);
  saver.format = format ? houdini_file_saver::bgeo : houdini_file_saver::geo;
  std::vector<float> points;
  if (packed) {
    points.resize(static_cast<size_t>(p->particles_count) * 7);
    for (unsigned int i = 0; i < p->particles_count; ++i) {
      float* o = &points[static_cast<size_t>(i) * 7];
      for (int k = 0; k < 3; ++k) o[k] = particles[i].position.s[k], o[3 + k] = particles[i].velocity.s[k];
      o[6] = particles[i].density;
    }
  }
  for (int k = 0; k < frames; ++k) {
    if (packed) saver.writeFramePoints(points.data(), p->particles_count, p->particle_mass, p->h);
    else saver.writeFrameToFile(particles, *p);
  }
  return 0;
}

// sph_simulation::simulate with the given settings and scene for `frames` frames; the lattice or
// last_frame.bin in the working directory is the start state. policy: 0 substep, 1 full frames,
// 2 never; callbacks: 0 none installed, 1 counting callbacks installed. states_out (nullable)
// receives the array seen by the last full-frame post_frame call.
int clsph_host_simulate(const simulation_parameters* p, const precomputed_kernel_values* t, float initial_volume,
                        const float* face_normals, const float* vertices, size_t n_vertex_floats,
                        const unsigned int* indices, unsigned int face_count, int frames, int policy, int callbacks,
                        particle* states_out, simulation_parameters* params_out, long* callback_calls) {
  quiet_cout q;
  sph_simulation sim;
  sim.quiet = true;
  sim.parameters = *p;
  sim.precomputed_terms = *t;
  sim.initial_volume = initial_volume;
  sim.current_scene.face_count = face_count;
  sim.current_scene.face_normals.assign(face_normals, face_normals + 3 * (size_t)face_count);
  sim.current_scene.vertices.assign(vertices, vertices + n_vertex_floats);
  sim.current_scene.indices.assign(indices, indices + 3 * (size_t)face_count);
  sim.host_sync = static_cast<sph_simulation::host_sync_policy>(policy);
  long calls = 0;
  if (callbacks) {
    sim.pre_frame = [&](particle*, const simulation_parameters&, bool) { ++calls; };
    sim.post_frame = [&](particle* parts, const simulation_parameters& prm, bool full) {
      ++calls;
      if (full && states_out) std::memcpy(states_out, parts, sizeof(particle) * prm.particles_count);
      if (full && params_out) *params_out = prm;
    };
  }
  try {
    sim.simulate(frames);
  } catch (const std::exception& e) {  // e.g. a truncated last_frame.bin (init_particles)
    std::cerr << "clsph_host_simulate: " << e.what() << std::endl;
    return 2;
  }
  if (!callbacks) {
    if (states_out) std::memcpy(states_out, sim.final_particles(), sizeof(particle) * p->particles_count);
    if (params_out) *params_out = sim.parameters;
  }
  if (callback_calls) *callback_calls = calls;
  return 0;
}

}  // extern "C"
