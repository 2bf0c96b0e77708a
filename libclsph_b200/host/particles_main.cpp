// particles_main.cpp -- `clsphparticles <fluid> <simulation_properties> <scene> <frames_prefix>`.
//
// Command-line driver with the behaviour of the reference's example/particles.cpp: loads the two
// JSON files and the scene (working-directory relative: fluid_properties/, simulation_properties/,
// scenes/, frames/), writes a .geo frame per full frame (or per sub-step with write_all_frames),
// optionally serialises last_frame.bin, resumes from it when its size matches, prints the settings
// and a progress bar. Extra flag: --yes skips the "press a key" prompt; --frames N limits the run;
// --sync full|substep chooses when the host array is refreshed (default: substep if
// write_all_frames or serialize, else full). --frame-export device|host: who prepares a frame -- `device` (default
// with --sync full): the GPU packs the 28 bytes per particle a frame needs and copies them while the next sub-steps
// run (sph_simulation::frame_saver), the 80-byte array is not downloaded at all; `host`: the reference's way, the
// pre_frame callback writes the downloaded array. Same files, byte for byte.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <cstdlib>
#include <string>
#include <vector>

#include "file_save_delegates/houdini_file_saver.h"
#include "sph_simulation.h"

int main(int argc, char** argv) {
  std::string positional[4];
  int n_positional = 0, frames = 0;
  bool assume_yes = false;
  std::string sync = "", frame_export = "", format = "";
  std::vector<std::string> device_options;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "--yes")) assume_yes = true;
    else if (!std::strcmp(argv[a], "--frames") && a + 1 < argc) frames = std::atoi(argv[++a]);
    else if (!std::strcmp(argv[a], "--sync") && a + 1 < argc) sync = argv[++a];
    else if (!std::strcmp(argv[a], "--frame-export") && a + 1 < argc) frame_export = argv[++a];
    else if (!std::strcmp(argv[a], "--format") && a + 1 < argc) format = argv[++a];  // geo | bgeo
    else if (!std::strcmp(argv[a], "--option") && a + 1 < argc) device_options.push_back(argv[++a]);  // name=value
    else if (n_positional < 4) positional[n_positional++] = argv[a];
  }
  if (n_positional < 4) {
    std::cout << "Too few arguments" << std::endl
              << "Usage: ./sph <fluid_name> <simulation_properties_name> <scene_name> <frames_folder_prefix>" << std::endl;
    return -1;
  }

  sph_simulation simulation;
  for (size_t k = 0; k < device_options.size(); ++k) {
    const std::string::size_type eq = device_options[k].find('=');
    if (eq == std::string::npos) {
      std::cerr << "--option expects name=value, got " << device_options[k] << std::endl;
      return -1;
    }
    simulation.device_options.push_back(std::make_pair(device_options[k].substr(0, eq), std::atoll(device_options[k].c_str() + eq + 1)));
  }
  houdini_file_saver saver = houdini_file_saver(positional[3]);
  if (format == "bgeo") saver.format = houdini_file_saver::bgeo;
  else if (format == "geo") saver.format = houdini_file_saver::geo;
  else if (!format.empty()) {
    std::cerr << "--format expects geo or bgeo, got " << format << std::endl;
    return -1;
  }
  try {
    simulation.load_settings("fluid_properties/" + positional[0] + ".json", "simulation_properties/" + positional[1] + ".json");
  } catch (const std::exception& ex) {
    std::cerr << ex.what() << std::endl;
    return -1;
  }

  const bool every_substep = simulation.write_intermediate_frames || simulation.serialize;
  simulation.host_sync = (sync == "substep" || (sync.empty() && every_substep)) ? sph_simulation::sync_every_substep
                                                                                 : sph_simulation::sync_full_frames;
  // frames straight from the device: only when the array is not needed on the host between frames
  const bool device_frames = simulation.host_sync != sph_simulation::sync_every_substep && frame_export != "host";
  if (device_frames) {
    simulation.frame_saver = &saver;
    simulation.host_sync = sph_simulation::sync_never;  // the callback below only draws the progress bar
  }
  int calls = 0;
  bool array_is_current = true;  // the initial state is on the host before the first callback
  simulation.pre_frame = [&](particle* particles, const simulation_parameters& params, bool full_frame) {
    const bool current = simulation.host_sync == sph_simulation::sync_every_substep || full_frame || array_is_current;
    if (!device_frames && simulation.write_intermediate_frames != full_frame && current) saver.writeFrameToFile(particles, params);
    if (simulation.serialize && current) {
      std::ofstream file_out("last_frame.bin", std::ios::binary);
      file_out.write(reinterpret_cast<const char*>(particles), static_cast<std::streamsize>(sizeof(particle)) * params.particles_count);
    }
    array_is_current = false;
    ++calls;
    const int num_frames = static_cast<int>(params.simulation_time / (params.time_delta * params.simulation_scale));
    const int progress = num_frames > 0 ? calls * 80 / num_frames : 80;
    std::string bar(80, ' ');
    for (int j = 0; j < 80 && j < progress; ++j) bar[j] = '-';
    std::cout << "[" << bar << "] " << calls << "/" << num_frames << std::endl;
  };

  const simulation_parameters& p = simulation.parameters;
  std::cout << std::endl
            << "Loaded parameters          " << std::endl
            << "-----------------          " << std::endl
            << "Simulation time:           " << p.simulation_time << std::endl
            << "Target FPS:                " << p.target_fps << std::endl
            << "Time delta:                " << p.time_delta << std::endl
            << "Simulation scale:          " << p.simulation_scale << std::endl
            << "Write intermediate frames: " << (simulation.write_intermediate_frames ? "true" : "false") << std::endl
            << "Serialize frames:          " << (simulation.serialize ? "true" : "false") << std::endl
            << std::endl
            << "Particle count:            " << p.particles_count << std::endl
            << "Particle mass:             " << p.particle_mass << std::endl
            << "Total mass:                " << p.total_mass << std::endl
            << "Initial volume:            " << simulation.initial_volume << std::endl
            << std::endl
            << "Fluid density:             " << p.fluid_density << std::endl
            << "Dynamic viscosity:         " << p.dynamic_viscosity << std::endl
            << "Surface tension threshold: " << p.surface_tension_threshold << std::endl
            << "Surface tension:           " << p.surface_tension << std::endl
            << "Stiffness (k):             " << p.K << std::endl
            << "Restitution:               " << p.restitution << std::endl
            << std::endl
            << "Kernel support radius (h): " << p.h << std::endl
            << std::endl
            << "Saving to folder:          " << saver.frames_folder_prefix + "frames/" << std::endl;

  if (!simulation.current_scene.load(positional[2])) {
    std::cerr << "Unable to load scene: " << positional[2] << std::endl;
    return -1;
  }

  // A checkpoint of the wrong size means a different particle count or an interrupted write.
  std::ifstream checkpoint("last_frame.bin", std::ios::in | std::ios::binary | std::ios::ate);
  if (checkpoint) {
    const size_t file_size = static_cast<size_t>(checkpoint.tellg());
    if (file_size == static_cast<size_t>(p.particles_count) * sizeof(particle)) {
      std::cout << std::endl << "\033[1;32m Serialized frame found.  Simulation will pick up where last run left off.\033[0m";
      std::cout << std::endl << "\033[1;32m To start a new simulation, delete last_frame.bin. \033[0m" << std::endl;
    } else {
      std::cout << std::endl
                << "\033[1;31m Serialized frame of incorrect size found. Revert to last know settings or delete it, then try again. \033[0m"
                << std::endl;
      return 0;
    }
  }

  char response = 'y';
  if (!assume_yes) {
    std::cout << std::endl << "Revise simulation parameters.  Press q to quit, any other key to proceed with simulation" << std::endl;
    std::cin >> response;
  }
  if (response != 'q') simulation.simulate(frames);
  return 0;
}
