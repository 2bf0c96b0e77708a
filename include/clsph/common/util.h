/* common/util.h -- 10-bit-per-axis Morton (Z-curve) cell index, host side.
 *
 * Same function names and results as the reference's libclsph/common/util.h:21-62
 * (get_grid_index_z_curve, get_cell_coords_z_curve, uninterleave), written in log-step form.
 * `static inline` so several translation units may include it. */
#ifndef CLSPH_COMMON_UTIL_H_
#define CLSPH_COMMON_UTIL_H_
#include "../clsph_types.h"

/* Moves bits 0..9 of v to positions 0, 3, 6, ..., 27. */
static inline cl_uint clsph_spread_bits_by_3(cl_uint v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

/* Inverse of the above: collects every third bit into the low 10 bits. */
static inline cl_uint uninterleave(cl_uint v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}

static inline cl_uint get_grid_index_z_curve(cl_uint x, cl_uint y, cl_uint z) {
  return clsph_spread_bits_by_3(x) | (clsph_spread_bits_by_3(y) << 1) | (clsph_spread_bits_by_3(z) << 2);
}

static inline cl_uint3 get_cell_coords_z_curve(cl_uint index) {
  cl_uint3 c;
  c.s[0] = uninterleave(index);
  c.s[1] = uninterleave(index >> 1);
  c.s[2] = uninterleave(index >> 2);
  c.s[3] = 0;
  return c;
}
#endif
