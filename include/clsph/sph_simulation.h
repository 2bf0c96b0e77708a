/* sph_simulation.h -- the libclsph simulation object on top of the CUDA C ABI.
 *
 * Public surface = the reference's libclsph/sph_simulation.h:8-27 (simulate, parameters,
 * precomputed_terms, pre_frame / post_frame, load_settings, write_intermediate_frames,
 * serialize, initial_volume, current_scene), so code written against libclsph compiles
 * unchanged. What differs is underneath: no cl:: members, no run-time kernel build, the state
 * lives on the GPU between sub-steps (include/clsph_cuda.h). */
#ifndef CLSPH_SPH_SIMULATION_H_
#define CLSPH_SPH_SIMULATION_H_

#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "common/structures.h"
#include "scene.h"

class houdini_file_saver;

class sph_simulation {
 public:
  sph_simulation();
  ~sph_simulation();
  sph_simulation(const sph_simulation&) = delete;
  sph_simulation& operator=(const sph_simulation&) = delete;

  /* Runs `frame_count` frames (0 = ceil(simulation_time * target_fps)), each made of
   * 1 / simulation_scale sub-steps, calling pre_frame / post_frame around every sub-step
   * (full_frame = false) and every frame (full_frame = true). */
  void simulate(int frame_count = 0);

  simulation_parameters parameters;
  precomputed_kernel_values precomputed_terms;

  std::function<void(particle*, const simulation_parameters&, bool)> pre_frame;
  std::function<void(particle*, const simulation_parameters&, bool)> post_frame;

  /* Reads the two JSON files and derives h, time_delta, max_velocity and the smoothing
   * constants. Throws std::runtime_error on a missing key or an invalid restitution. */
  void load_settings(std::string fluid_file_name, std::string parameters_file_name);

  bool write_intermediate_frames;
  bool serialize;
  float initial_volume;
  scene current_scene;

  /* ---- additions (defaults reproduce the reference's behaviour) ------------------------- */

  /* When the host array handed to the callbacks is refreshed from the GPU and written back.
   *   sync_every_substep   the reference's semantics: callbacks may read AND modify the array
   *                        around every sub-step (upload + download per sub-step);
   *   sync_full_frames     the array is current only when full_frame == true, and callbacks
   *                        must not modify it: one download per frame, no uploads;
   *   sync_never           callbacks get a stale array (progress bars, timers).
   * With no callback installed the state never leaves the GPU until simulate() returns. */
  enum host_sync_policy { sync_every_substep = 0, sync_full_frames = 1, sync_never = 2 };
  host_sync_policy host_sync;

  int cuda_device;          /* which GPU (default 0)                                        */
  /* clsph_set_option(name, value) pairs applied to the device context before the scene and the
   * particles are handed over, e.g. {"sub_cell_order", 1}, {"face_grid", 1} (include/clsph_cuda.h). */
  std::vector<std::pair<std::string, long long> > device_options;
  bool quiet;               /* suppress the reference's console chatter (default false)     */
  /* Frame export off the critical path: when set (and host_sync != sync_every_substep), simulate() itself hands the
   * state at the start of every frame -- the moment example/particles.cpp:93-96 writes it from its pre_frame
   * callback -- to this saver, packed on the GPU to the 28 bytes per particle a frame file needs and copied while
   * the frame's sub-steps already run (clsph_frame_begin / clsph_frame_end). Files are byte for byte those of
   * writeFrameToFile on the downloaded array. */
  houdini_file_saver* frame_saver;

  /* State after the last simulate() call, in the reference's output order. */
  const particle* final_particles() const;

 private:
  struct impl;
  impl* impl_;
  void init_particles(particle* buffer, const simulation_parameters&);
};

#endif
