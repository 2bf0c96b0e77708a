/* common/structures.h -- drop-in include path for libclsph user code.
 *
 * libclsph sources say `#include "common/structures.h"` and expect `particle`,
 * `simulation_parameters`, `precomputed_kernel_values` and the cl_* host types
 * (reference: libclsph/common/structures.h). Here they come from clsph_types.h, with the same
 * layouts and no OpenCL dependency. */
#ifndef CLSPH_COMMON_STRUCTURES_H_
#define CLSPH_COMMON_STRUCTURES_H_
#include "../clsph_types.h"
#ifdef __cplusplus
/* The reference's structures.h pulls in util/cl_boilerplate.h, and with it these standard headers
 * (util/cl_boilerplate.h:4-8). libclsph user code relies on that: example/particles.cpp uses
 * std::ofstream and std::filebuf without including <fstream> itself. */
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#endif
#define COLLISION_VOLUMES_COUNT 3
#endif
