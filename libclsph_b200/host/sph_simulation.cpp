// sph_simulation.cpp -- host side of the drop-in libclsph API on top of the CUDA C ABI.
//
// Mirrors the reference's libclsph/sph_simulation.cpp for everything a caller can observe:
// settings and derived constants (:405-506), the initial lattice or last_frame.bin resume
// (:48-94), the frame / sub-step loop with its callbacks (:346-403). The per-step work itself
// (:107-344: transfers, host bounds, host radix scan, host cell table, six OpenCL kernels) is
// replaced by calls into include/clsph_cuda.h and stays on the GPU.
#include "sph_simulation.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "clsph_cuda.h"
#include "file_save_delegates/houdini_file_saver.h"
#include "mini_json.h"

namespace {

const double kPi = 3.14159265358979323846;
const unsigned int kPreferredWorkGroupSizeMultiple = 64;

// Same contract as the reference's check_cl_error / exit_on_error (util/cl_boilerplate.h:28-34):
// report where it happened and terminate.
#define check_cuda_abi(ctx, call)                                                                      \
  do {                                                                                                 \
    int rc__ = (call);                                                                                 \
    if (rc__ != CLSPH_OK) {                                                                            \
      std::cerr << "A CUDA ABI error occured (" << __FILE__ << ":" << __LINE__ << ")-> " << rc__ << " " \
                << clsph_last_error(ctx) << std::endl;                                                 \
      std::exit(-1);                                                                                   \
    }                                                                                                  \
  } while (0)

float number_of(const clsph_host::json_value& obj, const char* key) {
  return static_cast<float>(obj.at(key).as_number(key));
}

clsph_host::json_value read_json_file(const std::string& path) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("Cannot open settings file: " + path);
  return clsph_host::parse_json_stream(in);
}

}  // namespace

struct sph_simulation::impl {
  clsph_context* ctx = nullptr;
  std::vector<particle> last_state;
};

sph_simulation::sph_simulation()
    : parameters(), precomputed_terms(), write_intermediate_frames(false), serialize(false), initial_volume(0.f),
      host_sync(sync_every_substep), cuda_device(0), quiet(false), frame_saver(nullptr), impl_(new impl) {}

sph_simulation::~sph_simulation() {
  if (impl_->ctx) clsph_destroy(impl_->ctx);
  delete impl_;
}

const particle* sph_simulation::final_particles() const {
  return impl_->last_state.empty() ? nullptr : impl_->last_state.data();
}

// Restores last_frame.bin if present, else places the particles on a cubic lattice centred on
// x = z = 0 and growing upward from y = 0 (reference :48-94). acceleration / grid_index, which the
// reference leaves uninitialised, are zeroed.
void sph_simulation::init_particles(particle* buffer, const simulation_parameters& params) {
  const int per_side = static_cast<int>(std::ceil(cbrtf(static_cast<float>(params.particles_count))));
  const float side_length = cbrtf(initial_volume);
  const float spacing = side_length / static_cast<float>(per_side);
  if (!quiet)
    std::cout << "volume: " << initial_volume << " side_length: " << side_length << " spacing: " << spacing << std::endl;

  std::ifstream resume("last_frame.bin", std::ios::in | std::ios::binary);
  if (resume) {
    // The reference reads the checkpoint with cereal's loadBinary (:59-67), which throws on a short read; a
    // truncated file must not leave the tail particles zeroed and coincident at the origin.
    const std::streamsize want = static_cast<std::streamsize>(sizeof(particle)) * params.particles_count;
    resume.read(reinterpret_cast<char*>(buffer), want);
    if (resume.gcount() != want)
      throw std::runtime_error("last_frame.bin: failed to read " + std::to_string(want) + " bytes from the checkpoint, read " +
                               std::to_string(resume.gcount()) + " (particles_count does not match the file)");
    return;
  }
  const unsigned int ups = static_cast<unsigned int>(per_side);
  const float half = side_length / 2.f;
  for (unsigned int i = 0; i < params.particles_count; ++i) {
    particle q = particle();
    q.position.s[0] = static_cast<float>(i % ups) * spacing - half;
    q.position.s[1] = static_cast<float>((i / ups) % ups) * spacing;
    q.position.s[2] = static_cast<float>(i / (ups * ups)) * spacing - half;
    buffer[i] = q;
  }
}

void sph_simulation::simulate(int frame_count) {
  if (frame_count == 0) frame_count = static_cast<int>(std::ceil(parameters.simulation_time * parameters.target_fps));
  const unsigned int n = parameters.particles_count;

  if (impl_->ctx) {
    clsph_destroy(impl_->ctx);
    impl_->ctx = nullptr;
  }
  {
    int rc = clsph_create(&impl_->ctx, cuda_device, n, 0);
    if (rc != CLSPH_OK) {
      std::cerr << "A CUDA ABI error occured (" << __FILE__ << ":" << __LINE__ << ")-> " << rc << " "
                << clsph_last_error(nullptr) << std::endl;
      std::exit(-1);
    }
  }
  clsph_context* ctx = impl_->ctx;
  for (size_t k = 0; k < device_options.size(); ++k)
    check_cuda_abi(ctx, clsph_set_option(ctx, device_options[k].first.c_str(), device_options[k].second));
  check_cuda_abi(ctx, clsph_set_scene(ctx, current_scene.face_normals.data(), current_scene.vertices.data(),
                                      current_scene.vertices.size(), current_scene.indices.data(), current_scene.face_count));
  check_cuda_abi(ctx, clsph_set_parameters(ctx, &parameters, &precomputed_terms));

  impl_->last_state.assign(n, particle());
  particle* particles = impl_->last_state.data();
  init_particles(particles, parameters);
  check_cuda_abi(ctx, clsph_upload_particles(ctx, particles, n));

  const bool any_callback = static_cast<bool>(pre_frame) || static_cast<bool>(post_frame);
  const host_sync_policy policy = any_callback ? host_sync : sync_never;

  // frame export by the library itself: two page-locked buffers, the copy of frame i overlaps its sub-steps
  const bool export_frames = frame_saver != nullptr && policy != sync_every_substep && frame_count > 0;
  float* frame_points[2] = {nullptr, nullptr};
  bool frame_in_flight = false;
  if (export_frames)
    for (int b = 0; b < 2; ++b) {
      void* mem = nullptr;
      check_cuda_abi(nullptr, clsph_host_alloc(&mem, sizeof(float) * 7u * static_cast<size_t>(n)));
      frame_points[b] = static_cast<float*>(mem);
    }
  auto collect_frame = [&](int buffer) {
    check_cuda_abi(ctx, clsph_frame_end(ctx));
    frame_saver->writeFramePoints(frame_points[buffer], n, parameters.particle_mass, parameters.h);
    frame_in_flight = false;
  };

  for (int i = 0; i < frame_count; ++i) {
    if (export_frames) {
      if (frame_in_flight) collect_frame((i + 1) & 1);
      check_cuda_abi(ctx, clsph_frame_begin(ctx, frame_points[i & 1], n));
      frame_in_flight = true;
    }
    if (pre_frame) pre_frame(particles, parameters, true);

    for (int j = 0; static_cast<float>(j) < (1.f / parameters.simulation_scale); ++j) {
      if (pre_frame) pre_frame(particles, parameters, false);

      if (policy == sync_every_substep) {
        // the reference's simulate_single_frame(particles, particles): callbacks may have edited the array
        check_cuda_abi(ctx, clsph_simulate_single_frame(ctx, particles, particles, &parameters, nullptr));
      } else {
        check_cuda_abi(ctx, clsph_step(ctx, 1));
      }

      if (post_frame) post_frame(particles, parameters, false);
    }

    if (policy == sync_full_frames) {
      check_cuda_abi(ctx, clsph_download_particles(ctx, particles));
      check_cuda_abi(ctx, clsph_get_parameters(ctx, &parameters));
    }
    if (post_frame) post_frame(particles, parameters, true);
  }

  if (export_frames) {
    if (frame_in_flight) collect_frame((frame_count + 1) & 1);
    frame_saver->wait();  // the saver copies what it is given, but leave nothing behind that points into the buffers
    clsph_host_free(frame_points[0]);
    clsph_host_free(frame_points[1]);
  }
  if (policy != sync_every_substep) {
    check_cuda_abi(ctx, clsph_download_particles(ctx, particles));
    check_cuda_abi(ctx, clsph_get_parameters(ctx, &parameters));
  }
  clsph_destroy(impl_->ctx);
  impl_->ctx = nullptr;
}

void sph_simulation::load_settings(std::string fluid_file_name, std::string parameters_file_name) {
  int particles_inside_influence_radius = 0;
  {
    const clsph_host::json_value fluid = read_json_file(fluid_file_name);
    parameters.fluid_density = number_of(fluid, "fluid_density");
    parameters.dynamic_viscosity = number_of(fluid, "dynamic_viscosity");
    parameters.restitution = number_of(fluid, "restitution");
    if (parameters.restitution < 0 || parameters.restitution > 1) throw std::runtime_error("Restitution has an invalid value!");
    parameters.K = number_of(fluid, "k");
    parameters.surface_tension_threshold = number_of(fluid, "surface_tension_threshold");
    parameters.surface_tension = number_of(fluid, "surface_tension");
    particles_inside_influence_radius = static_cast<int>(fluid.at("particles_inside_influence_radius").as_number("particles_inside_influence_radius"));
  }
  {
    const clsph_host::json_value sim = read_json_file(parameters_file_name);
    parameters.particles_count = static_cast<unsigned int>(sim.at("particles_count").as_number("particles_count"));
    if (parameters.particles_count % kPreferredWorkGroupSizeMultiple != 0 && !quiet) {
      std::cout << std::endl
                << "\033[1;31m You should choose a number of particles that is divisble by the preferred work group size.\033[0m";
      std::cout << std::endl << "\033[1;31m Performances will be sub-optimal.\033[0m" << std::endl;
    }
    parameters.particle_mass = number_of(sim, "particle_mass");
    parameters.simulation_time = number_of(sim, "simulation_time");
    parameters.target_fps = number_of(sim, "target_fps");
    parameters.simulation_scale = number_of(sim, "simulation_scale");
    const clsph_host::json_value& g = sim.at("constant_acceleration");
    parameters.constant_acceleration.s[0] = number_of(g, "x");
    parameters.constant_acceleration.s[1] = number_of(g, "y");
    parameters.constant_acceleration.s[2] = number_of(g, "z");
    write_intermediate_frames = sim.at("write_all_frames").as_bool("write_all_frames");
    serialize = sim.at("serialize").as_bool("serialize");
  }

  // Derived constants. Evaluation types follow the reference (:490-505): fp32 products, one double
  // division by 4*pi before cbrtf, and the smoothing constants entirely in double.
  parameters.total_mass = static_cast<float>(parameters.particles_count) * parameters.particle_mass;
  initial_volume = parameters.total_mass / parameters.fluid_density;
  const float volume_per_particle = initial_volume / static_cast<float>(parameters.particles_count);
  const float three_n_v = 3.f * (static_cast<float>(particles_inside_influence_radius) * volume_per_particle);
  parameters.h = cbrtf(static_cast<float>(static_cast<double>(three_n_v) / (static_cast<double>(4.f) * kPi)));
  parameters.time_delta = 1.f / parameters.target_fps;
  parameters.max_velocity = 0.8f * parameters.h / parameters.time_delta;

  const double h = static_cast<double>(parameters.h);
  const double h6 = std::pow(h, 6.0), h9 = std::pow(h, 9.0);
  precomputed_terms.poly_6 = static_cast<float>(315.0 / (64.0 * kPi * h9));
  precomputed_terms.poly_6_gradient = static_cast<float>(-945.0 / (32.0 * kPi * h9));
  precomputed_terms.poly_6_laplacian = static_cast<float>(-945.0 / (32.0 * kPi * h9));
  precomputed_terms.spiky = static_cast<float>(-45.0 / (kPi * h6));
  precomputed_terms.viscosity = static_cast<float>(45.0 / (kPi * h6));
}
