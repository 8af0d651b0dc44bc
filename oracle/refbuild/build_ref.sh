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
# tiled_mesh.cpp cannot be linked as a whole (it pulls in the engine), but the functions of the path can be compiled on their own: cut them out
# of the read-only source AT BUILD TIME into a generated TU (oracle/_ref is git-ignored: no reference source is committed) between a prelude and
# a harness of ours. Cut by function signature, not by line number: the file's header block (includes, constants, extern declarations = everything
# before its first function definition), get_max_sea_level, get_xy_scale .. tile_t::create_zvals (src/tiled_mesh.cpp:447-546) and
# tile_t::calc_mesh_ao_lighting (:586-662).
T=$R/src/tiled_mesh.cpp
mkdir -p "$OUT/gen"
{
  echo "// GENERATED at build time by oracle/refbuild/build_ref.sh from $T - do not commit"
  cat "$HERE/ref_tiled_prelude.inc"
  awk '/^[a-z].*\) \{\r?$/ {exit} {print}' "$T"
  echo 'float get_water_z_height(); bool using_hmap_with_detail(); extern terrain_hmap_manager_t terrain_hmap_manager; // declarations the cut-out functions need (defined in the harness / ref_stubs.cpp)'
  grep -E '^#define BILINEAR_INTERP' "$T"
  grep -E '^float get_max_sea_level +\(\)' "$T"
  awk '/^float get_xy_scale\(\) \{/ {p=1} p {print} p && /^bool tile_t::create_zvals/ {f=1} f && /^}/ {exit}' "$T"
  awk '/^void tile_t::calc_mesh_ao_lighting\(\) \{/ {p=1} p {print} p && /^}/ {exit}' "$T"
  # terrain weights texture: get_texture_ixs + tile_t::create_texture (src/tiled_mesh.cpp:1049-1352) and what it calls in
  # src/Textures.cpp - update_lttex_ix + get_tids (:1289-1312) with TEXTURE_SMOOTH (:12); h_dirt / clip_hd1 are that file's globals, defined here
  echo 'float h_dirt[NTEX_DIRT], clip_hd1; extern bool water_is_lava; extern int DISABLE_WATER; extern float vegetation; extern ttex lttex_dirt[NTEX_DIRT];'
  grep -E '^float const TEXTURE_SMOOTH' "$R/src/Textures.cpp"
  awk '/^void update_lttex_ix\(int &ix\)/ {p=1} p {print} p && /^void get_tids\(/ {f=1} f && /^}/ {exit}' "$R/src/Textures.cpp"
  echo 'extern float glaciate_exp;'
  awk '/^void gen_tex_height_tables\(\) \{/ {p=1} p {print} p && /^}/ {exit}' "$R/src/Textures.cpp"   # :1757-1761, called by init_terrain_mesh (mesh_gen.o)
  awk '/^void get_texture_ixs\(/ {p=1} p {print} p && /^}/ {exit}' "$T"
  awk '/^bool check_region_int\(/ {p=1} p {print} p && /^}/ {exit}' "$T"
  awk '/^void tile_t::create_texture\(mesh_xy_grid_cache_t &height_gen\) \{/ {p=1} p {print} p && /^}/ {exit}' "$T"
  cat "$HERE/ref_tiled_harness.inc"
} > "$OUT/gen/tiled_extract.cpp"
g++ $FLAGS $INC -I "$HERE" -c "$OUT/gen/tiled_extract.cpp" -o "$OUT/tiled_extract.o"
# The same for the voxel path: voxels.cpp cannot be linked either (renderer, collision objects), but
# voxel_manager itself is a plain container, so its member functions are cut out by signature: the include/constant block at the top of the file,
# voxel_grid<V>::init_grid/init, voxel_manager::clear .. create_procedural (src/voxels.cpp:271-346), atten_at_edges .. remove_unconnected_outside()
# (:403-610: the attenuation passes, interpolate_pt, add_triangles_for_voxel, val_is_outside, calc_outside_val, determine_voxels_outside) and
# FLOOD_FILL_INNER .. make_voxel_inside (:729-868: the flood fills).
V=$R/src/voxels.cpp
{
  echo "// GENERATED at build time by oracle/refbuild/build_ref.sh from $V - do not commit"
  cat "$HERE/ref_tiled_prelude.inc"
  awk '/^voxel_params_t global_voxel_params;/ {exit} {print}' "$V"
  echo 'extern int dynamic_mesh_scroll, rand_gen_index, scrolling, display_mode, mesh_gen_mode, mesh_freq_filter; void gen_rx_ry(float &rx, float &ry); // declarations the cut-out functions need'
  awk '/^template<typename V> void voxel_grid<V>::init_grid/ {p=1} /^\/\/ Note: assumes mesh is centered/ {exit} p {print}' "$V"
  awk '/^void voxel_manager::clear\(\)/ {p=1} /^void voxel_manager::create_from_cobjs/ {exit} p {print}' "$V"
  awk '/^void voxel_manager::atten_at_edges/ {p=1} p {print} p && /^void voxel_manager::remove_unconnected_outside\(\)/ {f=1} f && /^}/ {exit}' "$V"
  awk '/^#define FLOOD_FILL_INNER/ {p=1} /^bool voxel_manager::point_inside_volume/ {exit} p {print}' "$V"
  cat "$HERE/ref_voxels_harness.inc"
} > "$OUT/gen/voxels_extract.cpp"
g++ $FLAGS $INC -I "$HERE" -c "$OUT/gen/voxels_extract.cpp" -o "$OUT/voxels_extract.o"
# Mesh shadows: class mesh_shadow_gen + calc_mesh_shadows from src/visibility.cpp (:411-517) and the line clip they call
# (TEST_CLIP_T + do_line_clip, src/Math3d.cpp:1029-1034,1070-1086), cut out by signature; both files pull in the engine as a whole.
{
  echo "// GENERATED at build time by oracle/refbuild/build_ref.sh from $R/src/visibility.cpp and $R/src/Math3d.cpp - do not commit"
  echo '#include "3DWorld.h"'
  echo '#include "mesh.h"'
  echo 'extern float zmin, zmax; extern bool combined_gu; extern int XY_SUM_SIZE;'
  awk '/^#define TEST_CLIP_T/ {p=1} p {print} p && !/\\\r?$/ {exit}' "$R/src/Math3d.cpp"
  awk '/^bool do_line_clip\(point &v1, point &v2, float const d\[3\]\[2\]\) \{/ {p=1} p {print} p && /^}/ {exit}' "$R/src/Math3d.cpp"
  awk '/^class mesh_shadow_gen \{/ {p=1} /^void calc_visibility/ {exit} p {print}' "$R/src/visibility.cpp"
  cat "$HERE/ref_shadow_harness.inc"
} > "$OUT/gen/shadow_extract.cpp"
# Built WITHOUT -fopenmp: mesh_shadow_gen::run (src/visibility.cpp:497-503) runs run_x() and run_y() as two concurrent OpenMP sections that both store into
# sh_out_x/sh_out_y (last writer wins, unsynchronised), so the reference's own result is timing dependent in the cells both passes reach. The pinned semantics are
# the serial ones (run_x then run_y) -- one of the outcomes the reference can produce, and the only reproducible one.
g++ ${FLAGS/-fopenmp/} $INC -I "$HERE" -c "$OUT/gen/shadow_extract.cpp" -o "$OUT/shadow_extract.o"
g++ -shared -fopenmp -Wl,--gc-sections -Wl,--no-undefined -Wl,--version-script="$HERE/exports.map" -o "$OUT/libref3dworld.so" \
  "$OUT"/ref_driver.o "$OUT"/ref_stubs.o "$OUT"/ref_glm.o "$OUT"/mesh_gen.o "$OUT"/erosion.o "$OUT"/upsurface.o "$OUT"/heightmap.o "$OUT"/tiled_extract.o "$OUT"/voxels_extract.o "$OUT"/shadow_extract.o
# the generated translation units contain reference source text: they exist only for the duration of this build (TW_KEEP_GEN=1 keeps them for debugging);
# what stays in oracle/_ref/ (git-ignored) are objects and the shared library
if [ -z "$TW_KEEP_GEN" ]; then rm -rf "$OUT/gen"; fi
echo "built $OUT/libref3dworld.so"
