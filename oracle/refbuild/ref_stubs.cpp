// TEST INFRASTRUCTURE ONLY (oracle/_ref): the handful of engine globals/functions that the reference objects
// mesh_gen.o / erosion.o / upsurface.o reference but that live in translation units we cannot link (GL, textures, UI).
// Initial values mirror the reference defaults (src/3DWorld.cpp:89-116, src/matrix_ops.cpp:9-24); the GL shader methods
// are no-ops (never reached: the driver routes GPU gen modes to the CPU get_noise_zval path); get_bare_ls_tid follows
// src/Textures.cpp:1284-1287, rgen_core_t::randd follows src/gen_object.cpp:377-381, set_scene_constants_stub follows
// src/matrix_ops.cpp:57-86 for the globals this library owns.
#include "3DWorld.h"
#include "mesh.h"
#include "shaders.h"
#include "textures.h"

extern float zmin, zmax; // defined in mesh_gen.o

int MESH_X_SIZE(128), MESH_Y_SIZE(128), MESH_Z_SIZE(1);
float X_SCENE_SIZE(4.0), Y_SCENE_SIZE(4.0), Z_SCENE_SIZE(4.0);
float MESH_HEIGHT(0), XY_SCENE_SIZE(0);
float DX_VAL(0), DY_VAL(0), HALF_DXY(0), DX_VAL_INV(0), DY_VAL_INV(0);
int mesh_seed(0), mesh_rgen_index(0);
float water_plane_z(0.0), water_h_off(0.0), water_h_off_rel(0.0), custom_glaciate_exp(0.0), erode_amount(1.0);
float relh_adj_tex_stub(0.0), clip_hd1_stub(0.0);

extern "C" void set_scene_constants_stub() {
	MESH_HEIGHT   = 0.10f*Z_SCENE_SIZE;
	XY_SCENE_SIZE = 0.5f*(X_SCENE_SIZE + Y_SCENE_SIZE);
	DX_VAL        = (2.0f*X_SCENE_SIZE)/(float)MESH_X_SIZE;
	DY_VAL        = (2.0f*Y_SCENE_SIZE)/(float)MESH_Y_SIZE;
	HALF_DXY      = 0.5f*(DX_VAL + DY_VAL);
	DX_VAL_INV    = 1.0f/DX_VAL;
	DY_VAL_INV    = 1.0f/DY_VAL;
}

int get_bare_ls_tid(float zval) {
	float const relh(relh_adj_tex_stub + (zval - zmin)/(zmax - zmin));
	return ((relh > clip_hd1_stub) ? (int)ROCK_TEX : (int)DIRT_TEX);
}

double rgen_core_t::randd() {
	double rand_num;
	randome_int(rand_num);
	return rand_num/2147483563.;
}

void register_timing_value(const char *, int, bool) {}
extern "C" int glutGet(unsigned) {return 0;}
void free_texture(unsigned &tid) {tid = 0;}

void shader_t::enable () {}
void shader_t::disable() {}
void shader_t::set_prefix(char const *const, unsigned) {}
bool shader_t::add_uniform_float(char const *const, float) const {return 1;}
void compute_shader_t::begin() {}
void compute_shader_t::end_shader() {}
void compute_shader_t::setup_and_run(unsigned &, bool, bool, bool) {}
void compute_shader_t::prep_for_read_pixels(bool) {}
void compute_shader_t::read_float_vals(vector<float> &, bool, bool) {}

// ---- additional engine globals referenced by gen_mesh() and its helpers (src/3DWorld.cpp:89-133, src/matrix_ops.cpp:20-45); ground mode,
// no scrolling, no heightmap files, camera in the air (camera_mode 0 => update_temperature() returns early) ----
#include "heightmap.h"
float **mesh_height = nullptr;
unsigned char **mesh_draw = nullptr;
int xoff2(0), yoff2(0), world_mode(WMODE_GROUND), scrolling(0), read_landscape(0), read_heightmap(0), do_read_mesh(0), invert_mh_image(0);
int mesh_scale_change(0), camera_mode(0), MESH_SIZE[3] = {0}, XY_MULT_SIZE(0), XY_SUM_SIZE(0);
unsigned erosion_iters(0);
bool combined_gu(0);
float LARGE_ZVAL(0), CLOUD_CEILING(0), mesh_file_scale(1.0), mesh_file_tz(0.0), read_mesh_zmm(0.0), disabled_mesh_z(FAR_DISTANCE);
float temperature(20.0), univ_temp(20.0), init_temperature(20.0);
char *mh_filename(nullptr), *mesh_file(nullptr);
point mesh_origin, camera_pos, camera_origin, surface_pos;
rand_gen_t global_rand_gen;

extern "C" void set_scene_constants_stub2() { // the rest of set_scene_constants() + alloc_matrices() for mesh_height (src/matrix_ops.cpp:57-100)
	MESH_SIZE[0] = MESH_X_SIZE; MESH_SIZE[1] = MESH_Y_SIZE; MESH_SIZE[2] = MESH_Z_SIZE;
	XY_MULT_SIZE = MESH_X_SIZE*MESH_Y_SIZE; XY_SUM_SIZE = MESH_X_SIZE + MESH_Y_SIZE;
	CLOUD_CEILING = CLOUD_CEILING0*Z_SCENE_SIZE;
	LARGE_ZVAL    = 100.0f*CLOUD_CEILING;
	if (mesh_height) {delete [] mesh_height[0]; delete [] mesh_height; mesh_height = nullptr;}
	matrix_gen_2d(mesh_height);
}
void checked_fclose(FILE *fp) {if (fp) fclose(fp);}
bool open_file(FILE *&fp, char const *const fn, std::string const &file_type, char const *const mode) {fp = nullptr; return 0;}
void texture_t::resize(int, int) {}
void texture_t::load(int, bool, bool, bool) {}
void texture_t::gl_delete() {}
void texture_t::free_client_mem() {delete [] data; data = nullptr;} // client-memory part of src/Textures.cpp:512-518
void texture_t::alloc() {free_client_mem(); data = new unsigned char[(size_t)width*height*ncolors];} // src/Textures.cpp:486-490 without the GL state

// ---- get_exact_zval(): procedural branch only (no tiled-terrain heightmap texture loaded) ----
char *mh_filename_tt(nullptr);
bool using_tiled_terrain_hmap_tex() {return 0;}
bool using_hmap_with_detail() {return 0;}
float get_tiled_terrain_height_tex(float, float, bool) {return 0.0;}
unsigned hmap_filter_width(0); // src/3DWorld.cpp: config value "hmap_filter_width", default 0 (only used for 8-bit heightmaps)
