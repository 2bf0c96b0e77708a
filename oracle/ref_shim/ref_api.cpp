/*
 * ref_api.cpp -- C entry points onto the UNMODIFIED reference host code (TEST
 * INFRASTRUCTURE, see oracle/oracle.h).
 *
 * Links against libclsph/sph_simulation.cpp, libclsph/scene.cpp, util/cl_boilerplate.cpp and
 * util/tinyobj/tiny_obj_loader.cc compiled straight from the reference tree against the
 * OpenCL shim in this directory. Everything here goes through the reference's public API
 * (sph_simulation::load_settings / ::simulate / callbacks, scene::load) except the
 * clsph_ref_kernel_* functions, which enqueue a single reference kernel through the shim's
 * cl:: objects so that one stage can be observed in isolation.
 */
#include <omp.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <stdexcept>

#include "sph_simulation.h" /* the reference's, via -I<reference>/libclsph */
#include "file_save_delegates/houdini_file_saver.h" /* the reference's */

namespace {

struct stop_simulation {};

/* RAII chdir: the reference resolves kernels/, scenes/ and last_frame.bin against the CWD
 * (libclsph/sph_simulation.cpp:14, 60; libclsph/scene.cpp:13). */
class scoped_cwd {
 public:
  explicit scoped_cwd(const char* dir) : ok_(false) {
    if (!getcwd(saved_, sizeof(saved_))) return;
    ok_ = (dir == nullptr) || (chdir(dir) == 0);
  }
  ~scoped_cwd() {
    if (chdir(saved_) != 0) std::perror("clsph_ref: chdir back");
  }
  bool ok() const { return ok_; }

 private:
  char saved_[4096];
  bool ok_;
};

/* The reference prints a lot to std::cout; keep test logs readable. */
class quiet_cout {
 public:
  quiet_cout() : saved_(std::cout.rdbuf(sink_.rdbuf())) {}
  ~quiet_cout() { std::cout.rdbuf(saved_); }

 private:
  std::ostringstream sink_;
  std::streambuf* saved_;
};

double now_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

/* sph_simulation::load_settings (libclsph/sph_simulation.cpp:405-506) on real JSON files. */
int clsph_ref_load_settings(const char* fluid_json, const char* sim_json, simulation_parameters* p,
                            precomputed_kernel_values* t, float* initial_volume,
                            int* write_all_frames, int* serialize) {
  try {
    quiet_cout q;
    sph_simulation sim;
    sim.load_settings(fluid_json, sim_json);
    *p = sim.parameters;
    *t = sim.precomputed_terms;
    *initial_volume = sim.initial_volume;
    *write_all_frames = sim.write_intermediate_frames ? 1 : 0;
    *serialize = sim.serialize ? 1 : 0;
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "clsph_ref_load_settings: %s\n", e.what());
    return 1;
  }
}

/* scene::load (libclsph/scene.cpp:9-67). `root_dir` must contain scenes/<name>. Call once
 * with null arrays to get the sizes, then again with storage. */
int clsph_ref_scene_load(const char* root_dir, const char* name, uint32_t* face_count,
                         size_t* n_vertex_floats, size_t* n_indices, float* face_normals,
                         float* vertices, uint32_t* indices) {
  scoped_cwd cwd(root_dir);
  if (!cwd.ok()) return 2;
  quiet_cout q;
  scene s;
  if (!s.load(name)) return 1;
  *face_count = s.face_count;
  *n_vertex_floats = s.vertices.size();
  *n_indices = s.indices.size();
  if (face_normals) std::memcpy(face_normals, s.face_normals.data(), sizeof(float) * s.face_normals.size());
  if (vertices) std::memcpy(vertices, s.vertices.data(), sizeof(float) * s.vertices.size());
  if (indices) std::memcpy(indices, s.indices.data(), sizeof(uint32_t) * s.indices.size());
  return 0;
}

/*
 * sph_simulation::simulate (libclsph/sph_simulation.cpp:346-403) for exactly `substeps`
 * sub-steps. `work_dir` must contain kernels/sph.cl (only its existence matters to the shim).
 * `initial` (nullable) is injected the way the reference allows: as last_frame.bin.
 *   states_out      [substeps][N] (record_all != 0) or [N] (last state only); nullable
 *   params_out      parameters as seen by the callback after the last sub-step
 *   seconds_out     [substeps] wall time of each simulate_single_frame call; nullable
 */
int clsph_ref_simulate(const char* work_dir, const simulation_parameters* p,
                       const precomputed_kernel_values* t, float initial_volume,
                       const float* face_normals, const float* vertices, size_t n_vertex_floats,
                       const uint32_t* indices, uint32_t face_count, const particle* initial,
                       int substeps, int record_all, particle* states_out,
                       simulation_parameters* params_out, double* seconds_out) {
  scoped_cwd cwd(work_dir);
  if (!cwd.ok()) return 2;
  quiet_cout q;
  const size_t n = p->particles_count;
  std::remove("last_frame.bin");
  if (initial) {
    std::FILE* f = std::fopen("last_frame.bin", "wb");
    if (!f || std::fwrite(initial, sizeof(particle), n, f) != n) return 3;
    std::fclose(f);
  }

  sph_simulation sim;
  sim.parameters = *p;
  sim.precomputed_terms = *t;
  sim.initial_volume = initial_volume;
  sim.serialize = false;
  sim.current_scene.face_count = face_count;
  sim.current_scene.face_normals.assign(face_normals, face_normals + 3 * size_t(face_count));
  sim.current_scene.vertices.assign(vertices, vertices + n_vertex_floats);
  sim.current_scene.indices.assign(indices, indices + 3 * size_t(face_count));

  int done = 0;
  double t0 = 0.0;
  sim.pre_frame = [&](particle*, const simulation_parameters&, bool full_frame) {
    if (!full_frame) t0 = now_seconds();
  };
  sim.post_frame = [&](particle* parts, const simulation_parameters& prm, bool full_frame) {
    if (full_frame) return;
    double t1 = now_seconds();
    if (seconds_out) seconds_out[done] = t1 - t0;
    if (states_out && (record_all || done == substeps - 1))
      std::memcpy(states_out + (record_all ? size_t(done) * n : 0), parts, sizeof(particle) * n);
    if (params_out) *params_out = prm;
    if (++done == substeps) throw stop_simulation();
  };

  int rc = 0;
  try {
    const int per_frame = int(1.f / p->simulation_scale) + 1;
    sim.simulate(substeps / (per_frame > 1 ? per_frame - 1 : 1) + 2);
    rc = 4; /* ran out of frames before `substeps` */
  } catch (const stop_simulation&) {
    rc = 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "clsph_ref_simulate: %s\n", e.what());
    rc = 5;
  }
  std::remove("last_frame.bin");
  return rc;
}

/* houdini_file_saver::writeFrameToFile (libclsph/file_save_delegates/houdini_file_saver.cpp:25-92),
 * `frames` times in a row: writes <prefix>frames/frame000000K.geo. */
int clsph_ref_write_frames(const char* prefix, particle* particles, const simulation_parameters* p, int frames) {
  houdini_file_saver saver = houdini_file_saver(std::string(prefix));
  for (int k = 0; k < frames; ++k) saver.writeFrameToFile(particles, *p);
  return 0;
}

/* ---- single reference kernels through the shim's cl:: objects ---------------------------- */

static cl::Buffer make_buffer(cl::Context& ctx, cl::CommandQueue& q, const void* src, size_t bytes) {
  cl::Buffer b(ctx, CL_MEM_READ_WRITE, bytes ? bytes : 16);
  if (src && bytes) q.enqueueWriteBuffer(b, CL_TRUE, 0, bytes, src);
  return b;
}

static unsigned int work_group_size(unsigned int n) {
  /* libclsph/sph_simulation.cpp:181-185 */
  unsigned int wg = cl::shim::kMaxWorkGroupSize;
  while (n % wg != 0) wg /= 2;
  return wg;
}

/* kernels/grid.cl:43-67 */
int clsph_ref_kernel_locate_in_grid(const particle* in, particle* out, const simulation_parameters* p) {
  cl::Context ctx;
  cl::CommandQueue q;
  const size_t bytes = sizeof(particle) * p->particles_count;
  cl::Buffer bin = make_buffer(ctx, q, in, bytes), bout = make_buffer(ctx, q, nullptr, bytes);
  cl::Kernel k(cl::Program(), "locate_in_grid");
  k.setArg(0, bin);
  k.setArg(1, bout);
  k.setArg(2, *p);
  cl_int rc = q.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(p->particles_count),
                                     cl::NDRange(work_group_size(p->particles_count)));
  q.enqueueReadBuffer(bout, CL_TRUE, 0, bytes, out);
  return rc;
}

/* kernels/sph.cl:9-40 */
int clsph_ref_kernel_density_pressure(const particle* in, particle* out, const simulation_parameters* p,
                                      const precomputed_kernel_values* t, const uint32_t* cell_table) {
  cl::Context ctx;
  cl::CommandQueue q;
  const size_t bytes = sizeof(particle) * p->particles_count;
  const unsigned int wg = work_group_size(p->particles_count);
  cl::Buffer bin = make_buffer(ctx, q, in, bytes), bout = make_buffer(ctx, q, nullptr, bytes);
  cl::Buffer btab = make_buffer(ctx, q, cell_table, sizeof(uint32_t) * p->grid_cell_count);
  cl::Kernel k(cl::Program(), "density_pressure");
  k.setArg(0, bin);
  k.setArg(1, wg * sizeof(particle), nullptr);
  k.setArg(2, bout);
  k.setArg(3, *p);
  k.setArg(4, *t);
  k.setArg(5, btab);
  cl_int rc = q.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(p->particles_count), cl::NDRange(wg));
  q.enqueueReadBuffer(bout, CL_TRUE, 0, bytes, out);
  return rc;
}

/* kernels/sph.cl:42-62 */
int clsph_ref_kernel_forces(const particle* in, particle* out, const simulation_parameters* p,
                            const precomputed_kernel_values* t, const uint32_t* cell_table) {
  cl::Context ctx;
  cl::CommandQueue q;
  const size_t bytes = sizeof(particle) * p->particles_count;
  cl::Buffer bin = make_buffer(ctx, q, in, bytes), bout = make_buffer(ctx, q, nullptr, bytes);
  cl::Buffer btab = make_buffer(ctx, q, cell_table, sizeof(uint32_t) * p->grid_cell_count);
  cl::Kernel k(cl::Program(), "forces");
  k.setArg(0, bin);
  k.setArg(1, bout);
  k.setArg(2, *p);
  k.setArg(3, *t);
  k.setArg(4, btab);
  cl_int rc = q.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(p->particles_count),
                                     cl::NDRange(work_group_size(p->particles_count)));
  q.enqueueReadBuffer(bout, CL_TRUE, 0, bytes, out);
  return rc;
}

/* kernels/sph.cl:64-112 */
int clsph_ref_kernel_advection_collision(const particle* in, particle* out, const simulation_parameters* p,
                                         const precomputed_kernel_values* t, const float* face_normals,
                                         const float* vertices, size_t n_vertex_floats,
                                         const uint32_t* indices, uint32_t face_count) {
  cl::Context ctx;
  cl::CommandQueue q;
  const size_t bytes = sizeof(particle) * p->particles_count;
  cl::Buffer bin = make_buffer(ctx, q, in, bytes), bout = make_buffer(ctx, q, nullptr, bytes);
  cl::Buffer btab = make_buffer(ctx, q, nullptr, 16);
  cl::Buffer bn = make_buffer(ctx, q, face_normals, sizeof(float) * 3 * face_count);
  cl::Buffer bv = make_buffer(ctx, q, vertices, sizeof(float) * n_vertex_floats);
  cl::Buffer bi = make_buffer(ctx, q, indices, sizeof(uint32_t) * 3 * face_count);
  cl::Kernel k(cl::Program(), "advection_collision");
  k.setArg(0, bin);
  k.setArg(1, bout);
  k.setArg(2, *p);
  k.setArg(3, *t);
  k.setArg(4, btab);
  k.setArg(5, bn);
  k.setArg(6, bv);
  k.setArg(7, bi);
  k.setArg(8, face_count);
  cl_int rc = q.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(p->particles_count),
                                     cl::NDRange(work_group_size(p->particles_count)));
  q.enqueueReadBuffer(bout, CL_TRUE, 0, bytes, out);
  return rc;
}

int clsph_ref_num_threads(void) { return omp_get_max_threads(); }
void clsph_ref_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

}  // extern "C"
