/* scene.h -- triangle scene the fluid collides with (same public surface as the reference's
 * libclsph/scene.h:7-15: load(), face_count, face_normals, vertices, indices). */
#ifndef CLSPH_SCENE_H_
#define CLSPH_SCENE_H_

#include <string>
#include <vector>

class scene {
 public:
  scene() : face_count(0) {}

  /* Reads scenes/<filename> (Wavefront OBJ, relative to the working directory like the
   * reference, libclsph/scene.cpp:13) and computes one unit normal per triangle on the host.
   * Returns false, with a message on stderr, if the file cannot be read or holds a
   * non-triangle mesh after fan triangulation. */
  bool load(std::string filename);

  unsigned int face_count;
  std::vector<float> face_normals;    /* 3 per face                    */
  std::vector<float> vertices;        /* 3 per vertex                  */
  std::vector<unsigned int> indices;  /* 3 per face, into `vertices`   */
};

#endif
