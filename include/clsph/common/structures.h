/* common/structures.h -- drop-in include path for libclsph user code.
 *
 * libclsph sources say `#include "common/structures.h"` and expect `particle`,
 * `simulation_parameters`, `precomputed_kernel_values` and the cl_* host types
 * (reference: libclsph/common/structures.h). Here they come from clsph_types.h, with the same
 * layouts and no OpenCL dependency. */
#ifndef CLSPH_COMMON_STRUCTURES_H_
#define CLSPH_COMMON_STRUCTURES_H_
#include "../clsph_types.h"
#define COLLISION_VOLUMES_COUNT 3
#endif
