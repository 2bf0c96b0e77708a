#!/usr/bin/env bash
# build_ref.sh -- compile the reference's own hot path into oracle/_ref/libclsph_ref.so
# (TEST INFRASTRUCTURE, see oracle/oracle.h). Needs the reference tree (this container only;
# the built .so travels to the GPU box, the tree does not).
#
# What gets compiled, straight from $REFERENCE with no edits:
#   libclsph/sph_simulation.cpp  libclsph/scene.cpp  util/cl_boilerplate.cpp
#   util/tinyobj/tiny_obj_loader.cc            (+ vendored picojson / cereal headers)
#   libclsph/file_save_delegates/houdini_file_saver.cpp  util/houdini_geo/HoudiniFileDumpHelper.cpp
# against oracle/ref_shim/CL/cl.hpp (an in-process OpenCL stand-in), and the kernel program
#   libclsph/kernels/*.cl + libclsph/common/{structures,util}.h
# compiled as C++ behind oracle/ref_shim/cl_device.h. The kernel files need three one-line
# patches to run on ANY 64-bit host; they are applied to a generated copy under
# oracle/_ref/gen/ (git-ignored, never committed):
#   E1  sort.cl:42   `global size_t* start_indices` -> `global unsigned int*` (the buffer holds
#                    32-bit counters; as shipped the kernel only works where size_t is 32 bit)
#   E2  sph.cl:49    `particle output_particle;` is stored whole while only .acceleration was
#                    set; OpenCL compilers drop the undefined stores, g++ does not
#                    -> initialise it from the input particle (the intended semantics)
#   SYN sort.cl:19   `(uint2)(a, b)` is an OpenCL vector literal, not C++ -> `uint2(a, b)`
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REFERENCE="${REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
# NB: this image exports CXX=/opt/gcc/bin/g++, a wrapper without libgomp; use the system g++.
CXX="${CLSPH_CXX:-g++}"

if [ ! -f "$REFERENCE/libclsph/kernels/sph.cl" ]; then
  echo "build_ref.sh: reference tree not found at $REFERENCE (nothing to do)" >&2
  exit 3
fi

mkdir -p "$GEN/kernels" "$GEN/common" "$OUT/obj"
for f in "$REFERENCE"/libclsph/kernels/*.cl; do
  sed -e 's/global size_t\* start_indices/global unsigned int* start_indices/' \
      -e 's/return (uint2)(start_index, end_index);/return uint2(start_index, end_index);/' \
      -e 's/^  particle output_particle;$/  particle output_particle = input_data[current_particle_index];/' \
      "$f" > "$GEN/kernels/$(basename "$f")"
done
cp "$REFERENCE/libclsph/common/structures.h" "$REFERENCE/libclsph/common/util.h" "$GEN/common/"
# every patch must have landed exactly once
grep -q 'global unsigned int\* start_indices' "$GEN/kernels/sort.cl"
grep -q 'return uint2(start_index, end_index);' "$GEN/kernels/sort.cl"
[ "$(grep -c 'particle output_particle = input_data\[current_particle_index\];' "$GEN/kernels/sph.cl")" = 2 ]

# -ffp-contract=off: evaluate the reference's expressions as written (oracle.h contract).
FLAGS="-O2 -std=c++11 -g0 -fPIC -fopenmp -mfma -ffp-contract=off -w"
INC="-I$HERE/ref_shim -I$REFERENCE -I$REFERENCE/libclsph -I$HERE/../include"

$CXX $FLAGS $INC -c "$REFERENCE/libclsph/sph_simulation.cpp"     -o "$OUT/obj/sph_simulation.o"
$CXX $FLAGS $INC -c "$REFERENCE/libclsph/scene.cpp"              -o "$OUT/obj/scene.o"
$CXX $FLAGS $INC -c "$REFERENCE/util/cl_boilerplate.cpp"         -o "$OUT/obj/cl_boilerplate.o"
$CXX $FLAGS $INC -c "$REFERENCE/util/tinyobj/tiny_obj_loader.cc" -o "$OUT/obj/tiny_obj_loader.o"
$CXX $FLAGS $INC -c "$REFERENCE/libclsph/file_save_delegates/houdini_file_saver.cpp" -o "$OUT/obj/houdini_file_saver.o"
$CXX $FLAGS $INC -c "$REFERENCE/util/houdini_geo/HoudiniFileDumpHelper.cpp" -o "$OUT/obj/HoudiniFileDumpHelper.o"
$CXX $FLAGS $INC -DREF_KERNEL_PROGRAM="\"$GEN/kernels/sph.cl\"" \
                 -c "$HERE/ref_shim/cl_runtime.cpp"              -o "$OUT/obj/cl_runtime.o"
$CXX $FLAGS $INC -c "$HERE/ref_shim/ref_api.cpp"                 -o "$OUT/obj/ref_api.o"
$CXX -shared -fopenmp -o "$OUT/libclsph_ref.so" "$OUT"/obj/*.o
echo "built $OUT/libclsph_ref.so"
