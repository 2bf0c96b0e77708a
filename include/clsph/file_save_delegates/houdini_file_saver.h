/* houdini_file_saver.h -- per-frame Houdini particle files: ASCII .geo, or binary .bgeo (same public
 * surface as the reference's libclsph/file_save_delegates/houdini_file_saver.h:8-20, plus
 * `asynchronous`, `format` and wait()). */
#ifndef CLSPH_HOUDINI_FILE_SAVER_H_
#define CLSPH_HOUDINI_FILE_SAVER_H_

#include <string>

#include "common/structures.h"

class houdini_file_saver {
 public:
  houdini_file_saver(std::string frames_folder_prefix);
  houdini_file_saver(const houdini_file_saver& other);            /* copies settings, not pending frames */
  houdini_file_saver& operator=(const houdini_file_saver& other);
  ~houdini_file_saver();                                          /* waits for pending frames */

  /* Writes <prefix>frames/frameNNNNNNN.geo ("PGEOMETRY V5": position, v, colour ramp from the
   * density, mass), byte for byte what the reference writes, and returns 0; prints to stderr if
   * the file cannot be opened. With `asynchronous` (default) the call only copies what the frame
   * needs and returns; formatting (all host cores) and the write happen on a background thread,
   * at most two frames in flight. */
  int writeFrameToFile(particle* particles, const simulation_parameters& parameters);

  /* The same frame from what the file actually needs of each particle: n records of seven floats (position,
   * velocity, density) in the order of the particle array -- what clsph_frame_begin packs on the GPU (28 instead of
   * 80 bytes per particle over the host link). sph_simulation::frame_saver uses it. Same files, byte for byte. */
  int writeFramePoints(const float* points, unsigned int count, float particle_mass, float support_radius = 0.f);

  /* Blocks until every frame handed to writeFrameToFile / writeFramePoints is on disk. */
  void wait();

  /* geo: "PGEOMETRY V5" text (the reference's default build). bgeo: <prefix>frames/frameNNNNNNN.bgeo, the binary
   * "Bgeo V5" file the reference writes through libpartio when it is compiled with USE_PARTIO (position, velocity,
   * color, id, mass, pscale = the support radius; houdini_file_saver.cpp:78-88, util/partio/PartioFunctions.h:5-65).
   * Compiling this library with -DUSE_PARTIO makes bgeo the default, as in the reference. */
  enum frame_format { geo = 0, bgeo = 1 };

  std::string frames_folder_prefix;
  bool asynchronous;
  frame_format format;

 private:
  struct writer;
  int submit_job(void* frame_job);
  int frame_count;
  writer* writer_;
};

#endif
