#!/bin/bash
# Builds oracle/_ref/libref3dworld.so from the UNMODIFIED reference sources under /root/reference (read-only) plus the
# driver/stub TUs in this directory. Same flags as the reference makefile:11 (-O3 -fopenmp, no -march). Outputs only into oracle/_ref/.
set -e
R=${REFERENCE_ROOT:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../_ref
mkdir -p "$OUT" "$HERE/shim/GL"
: > "$HERE/shim/GL/gl.h"; : > "$HERE/shim/GL/glu.h"
INC="-I $HERE/shim -I $R/dependencies/glew-2.0.0/include -I $R/dependencies/freeglut-3.3.2/include -I $R/src -I $R/src/texture_tile_blend -I $R/Targa -I $R/dependencies/gli -I $R/dependencies/glm -I $R/dependencies/meshoptimizer/src -I $R/dependencies/stb"
FLAGS="-O3 -fopenmp -w -fPIC -ffunction-sections -fdata-sections"
for f in mesh_gen erosion upsurface heightmap; do
  [ "$OUT/$f.o" -nt "$R/src/$f.cpp" ] || g++ $FLAGS $INC -c "$R/src/$f.cpp" -o "$OUT/$f.o"
done
for f in ref_driver ref_stubs ref_glm; do
  g++ $FLAGS $INC -c "$HERE/$f.cpp" -o "$OUT/$f.o"
done
g++ -shared -fopenmp -Wl,--gc-sections -Wl,--no-undefined -Wl,--version-script="$HERE/exports.map" -o "$OUT/libref3dworld.so" \
  "$OUT"/ref_driver.o "$OUT"/ref_stubs.o "$OUT"/ref_glm.o "$OUT"/mesh_gen.o "$OUT"/erosion.o "$OUT"/upsurface.o "$OUT"/heightmap.o
echo "built $OUT/libref3dworld.so"
