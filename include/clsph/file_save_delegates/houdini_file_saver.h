/* houdini_file_saver.h -- per-frame ASCII Houdini .geo writer (same public surface as the
 * reference's libclsph/file_save_delegates/houdini_file_saver.h:8-20). */
#ifndef CLSPH_HOUDINI_FILE_SAVER_H_
#define CLSPH_HOUDINI_FILE_SAVER_H_

#include <string>

#include "common/structures.h"

class houdini_file_saver {
 public:
  houdini_file_saver(std::string frames_folder_prefix)
      : frames_folder_prefix(frames_folder_prefix), frame_count(0) {}

  /* Writes <prefix>frames/frameNNNNNNN.geo ("PGEOMETRY V5": position, v, colour ramp from the
   * density, mass) and returns 0; prints to stderr if the file cannot be opened. */
  int writeFrameToFile(particle* particles, const simulation_parameters& parameters);

  std::string frames_folder_prefix;

 private:
  int frame_count;
};

#endif
