// TEST INFRASTRUCTURE ONLY (oracle/_ref): C-ABI driver around the UNMODIFIED reference
// translation units src/mesh_gen.cpp, src/erosion.cpp, src/upsurface.cpp of fegennari/3DWorld,
// compiled where they lie under /root/reference by oracle/refbuild/build_ref.sh.
// Nothing here is product code; no reference source is copied - the reference headers are
// included from /root/reference at build time and the reference objects are linked as-is.
#include "3DWorld.h"     // must come first (it renames the libc timer_t around its std includes)
#include <cstring>
#include <unordered_map>
#include <unordered_set>
#include <memory>
#include <list>
#include <fstream>
#include "collision_detect.h"
#define class struct     // test-only: mesh_xy_grid_cache_t's default-private members become visible, so the driver can select
#include "mesh.h"        // gen_mode 3/4 (whose reference implementation needs a GL shader) on the CPU get_noise_zval() path
#undef class
#include "upsurface.h"
#include "heightmap.h"
#include "sinf.h"
#include <omp.h>

// ---- symbols defined by the reference objects we link ----
extern int start_eval_sin, GLACIATE, mesh_gen_mode, mesh_gen_shape, mesh_freq_filter;
extern float zmax, zmin, zmax_est, zbottom, ztop, mesh_scale, mesh_scale_z, mesh_scale_z_inv, mesh_height_scale, zmax_est2, zmax_est2_inv;
extern float glaciate_exp, glaciate_exp_inv;
extern float sinTable[][5];
extern hmap_params_t hmap_params;
extern float MESH_START_MAG, MESH_START_FREQ, MESH_MAG_MULT, MESH_FREQ_MULT;
void create_sin_table();
void gen_rand_sine_table_entries(float scaled_height);
void compute_scale();
void set_zmax_est(float zval);
float get_noise_zval(float xval, float yval, int mode, int shape);
float eval_mesh_sin_terms(float xv, float yv);
void apply_glaciate(float &zval);
float get_water_z_height();
float eval_mesh_sin_terms_scaled(float xval, float yval, float xy_scale);
float get_exact_zval(float xval_in, float yval_in, bool no_xyoff);
extern int xoff2, yoff2;
extern float mesh_file_scale, mesh_file_tz;
void apply_erosion(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters);
void gen_rx_ry(float &rx, float &ry);

// ---- globals the reference objects expect from the rest of the engine (stubs) ----
extern int mesh_seed, mesh_rgen_index;
extern float erode_amount, water_plane_z, custom_glaciate_exp, MESH_HEIGHT;
extern float relh_adj_tex_stub, clip_hd1_stub;

extern "C" {

struct ref_scene_t {
	int mesh_x, mesh_y, mesh_z;           // mesh_size
	float xss, yss, zss;                   // scene_size
	float mesh_scale, mesh_height_scale;   // mesh_scale, mesh_height
	int gen_mode, gen_shape, freq_filter, seed, rgen_index, glaciate;
	float custom_glaciate_exp;
	float hmap[14];                        // hmap_params_t in declaration order
	float zmax_est;                        // if >0: set_zmax_est(zmax_est)
	float mesh_scale_z;                    // mesh_scale_z (1 => inv 1)
};

void set_scene_constants_stub(); // defined in ref_stubs.cpp (restates matrix_ops.cpp:59-85 for the globals we own)

void ref_setup(const ref_scene_t *s, int gen_sine_table) {
	MESH_X_SIZE = s->mesh_x; MESH_Y_SIZE = s->mesh_y; MESH_Z_SIZE = s->mesh_z;
	X_SCENE_SIZE = s->xss; Y_SCENE_SIZE = s->yss; Z_SCENE_SIZE = s->zss;
	set_scene_constants_stub();
	mesh_scale = s->mesh_scale; mesh_height_scale = s->mesh_height_scale;
	mesh_scale_z = s->mesh_scale_z; mesh_scale_z_inv = 1.0/mesh_scale_z;
	mesh_gen_mode = s->gen_mode; mesh_gen_shape = s->gen_shape; mesh_freq_filter = s->freq_filter;
	mesh_seed = s->seed; mesh_rgen_index = s->rgen_index; GLACIATE = s->glaciate;
	custom_glaciate_exp = s->custom_glaciate_exp;
	memcpy(&hmap_params, s->hmap, 14*sizeof(float));
	create_sin_table();
	compute_scale();
	if (gen_sine_table) {gen_rand_sine_table_entries(MESH_HEIGHT*mesh_height_scale);}
	if (s->zmax_est > 0.0f) {set_zmax_est(s->zmax_est);}
	glaciate_exp     = ((custom_glaciate_exp == 0.0) ? 3.0 : custom_glaciate_exp);
	glaciate_exp_inv = 1.0/glaciate_exp;
}

int   ref_get_start_eval_sin() {return start_eval_sin;}
void  ref_set_start_eval_sin(int v) {start_eval_sin = v;}
float ref_get_dx() {return DX_VAL;}
float ref_get_dy() {return DY_VAL;}
float ref_get_mesh_height() {return MESH_HEIGHT;}
float ref_get_half_dxy() {return HALF_DXY;}
void  ref_get_sin_table(float *out) {memcpy(out, sin_table.data(), 2*TSIZE*sizeof(float));}
void  ref_get_sine_params(float *out) {memcpy(out, sinTable, 90*5*sizeof(float));}
void  ref_set_sine_params(const float *in) {memcpy(sinTable, in, 90*5*sizeof(float));}
void  ref_get_rx_ry(float *rx, float *ry) {gen_rx_ry(*rx, *ry);}
void  ref_set_zmax_est(float v) {set_zmax_est(v);}
float ref_get_zmax_est() {return zmax_est;}
void  ref_set_threads(int n) {omp_set_num_threads(n);}
int   ref_get_max_threads() {return omp_get_max_threads();}

// mesh_xy_grid_cache_t::build_arrays + [enable_glaciate] + eval_index over the grid, as heightmap_t::proc_gen / tile_t::create_zvals do
void ref_heightgen(float x0, float y0, float dx, float dy, unsigned nx, unsigned ny, int cache_values, int glaciate, int min_start_sin, float *out) {
	int const mode(mesh_gen_mode);
	mesh_xy_grid_cache_t hg;
	if (mode >= MGEN_SIMPLEX_GPU) {mesh_gen_mode = MGEN_SIMPLEX;} // no GL here: build as a CPU mode, then select the CPU restatement of mode 3/4 (get_noise_zval) below
	hg.build_arrays(x0, y0, dx, dy, nx, ny, (cache_values && mode < MGEN_SIMPLEX_GPU));
	mesh_gen_mode = mode;
	hg.gen_mode   = mode;
	if (glaciate) {hg.enable_glaciate();}
#pragma omp parallel for schedule(static,1)
	for (int y = 0; y < (int)ny; ++y) {
		for (unsigned x = 0; x < nx; ++x) {out[(size_t)y*nx + x] = hg.eval_index(x, y, min_start_sin);}
	}
}

float ref_get_noise_zval(float xval, float yval, int mode, int shape) {return get_noise_zval(xval, yval, mode, shape);}
float ref_eval_mesh_sin_terms(float xv, float yv) {return eval_mesh_sin_terms(xv, yv);}
float ref_get_water_z_height() {return get_water_z_height();}
// point queries: eval_mesh_sin_terms_scaled / get_exact_zval, procedural branch (no heightmap texture, no landscape file)
float ref_eval_mesh_sin_terms_scaled(float xval, float yval, float xy_scale) {return eval_mesh_sin_terms_scaled(xval, yval, xy_scale);}
// tile_t::upload_normal_texture's per-cell arithmetic (src/tiled_mesh.cpp:865-880) with the reference's own vector3d::get_norm(), min() and
// UNROLL_3X from 3DWorld.h; the loop/indexing is the driver's (tiled_mesh.cpp itself cannot be linked), get_norm_not_normalized as src/tiled_mesh.h:281-283
void ref_tile_normals(const float *zvals, unsigned zvsize, float dx_val, float dy_val, unsigned char *normal_data, float *min_normal_z_out) {
	unsigned const stride(zvsize - 1);
	float const dxdy_(dx_val*dy_val);
	float min_normal_z(1.0);
	for (unsigned y = 0; y < stride; ++y) {
		for (unsigned x = 0; x < stride; ++x) {
			unsigned const ix(y*stride + x), ix2(y*zvsize + x), ix_off(4*ix);
			vector3d const norm(vector3d(dy_val*(zvals[ix2] - zvals[ix2 + 1]), dx_val*(zvals[ix2] - zvals[ix2 + zvsize]), dxdy_).get_norm());
			min_normal_z = min(min_normal_z, norm.z);
			UNROLL_3X(normal_data[ix_off+i_] = (unsigned char)(127.0*(norm[i_] + 1.0)););
			normal_data[ix_off+3] = 0;
		}
	}
	*min_normal_z_out = min_normal_z;
}
// terrain_hmap_manager_t::get_clamped_height (src/heightmap.cpp:385-402) over a zvsize^2 tile at (x1, y1), as tile_t::create_zvals' heightmap
// branch calls it (src/tiled_mesh.cpp:500); the manager's protected heightmap_t is filled with a caller-provided 16-bit image
struct ref_hmap_mgr_t : public terrain_hmap_manager_t {
	void set_image(const unsigned char *data16, int w, int h) {hmap.width = w; hmap.height = h; hmap.ncolors = 2; hmap.alloc(); memcpy(hmap.get_data(), data16, 2*(size_t)w*h);}
	void clear_image() {hmap.free_client_mem();}
};
void ref_hmap_sample_tile(const unsigned char *data16, int w, int h, int x1, int y1, unsigned zvsize, float mesh_scale_, float mesh_file_scale_, float mesh_file_tz_, float *out) {
	float const ms(mesh_scale), mfs(mesh_file_scale), mft(mesh_file_tz);
	mesh_scale = mesh_scale_; mesh_file_scale = mesh_file_scale_; mesh_file_tz = mesh_file_tz_;
	{
		ref_hmap_mgr_t mgr;
		mgr.set_image(data16, w, h);
		for (unsigned y = 0; y < zvsize; ++y) {
			for (unsigned x = 0; x < zvsize; ++x) {out[y*zvsize + x] = mgr.get_clamped_height((x1 + x), (y1 + y));}
		}
		mgr.clear_image();
	}
	mesh_scale = ms; mesh_file_scale = mfs; mesh_file_tz = mft;
}
// kind 0/1/2 = eval_mesh_sin_terms / eval_mesh_sin_terms_scaled / get_exact_zval for n points (the loop is the driver's; each value is the reference's)
void ref_eval_points(int kind, const float *xy, size_t n, float xy_scale, int no_xyoff, int xo, int yo, float *out) {
	int const xs(xoff2), ys(yoff2);
	xoff2 = xo; yoff2 = yo;
	for (size_t i = 0; i < n; ++i) {
		float const x(xy[2*i]), y(xy[2*i+1]);
		out[i] = (kind == 0) ? eval_mesh_sin_terms(x, y) : ((kind == 1) ? eval_mesh_sin_terms_scaled(x, y, xy_scale) : get_exact_zval(x, y, (no_xyoff != 0)));
	}
	xoff2 = xs; yoff2 = ys;
}
float ref_get_exact_zval(float xval, float yval, int no_xyoff, int xo, int yo) {
	int const xs(xoff2), ys(yoff2);
	xoff2 = xo; yoff2 = yo;
	float const z(get_exact_zval(xval, yval, (no_xyoff != 0)));
	xoff2 = xs; yoff2 = ys;
	return z;
}
float ref_glm_simplex2(float x, float y);
float ref_glm_perlin2 (float x, float y);
float ref_glm_simplex3(float x, float y, float z);
float ref_glm_perlin3 (float x, float y, float z);

struct ref_erosion_t {float erode_amount, water_plane_z, half_dxy, zmin, zmax, relh_adj_tex, clip_hd1;};

void ref_apply_erosion(float *hmap, int xsize, int ysize, float min_zval, unsigned num_iters, const ref_erosion_t *p) {
	erode_amount = p->erode_amount; water_plane_z = p->water_plane_z; HALF_DXY = p->half_dxy;
	zmin = p->zmin; zmax = p->zmax; relh_adj_tex_stub = p->relh_adj_tex; clip_hd1_stub = p->clip_hd1;
	apply_erosion(hmap, xsize, ysize, min_zval, num_iters);
}

// noise_gen_3d + the create_procedural() loop of src/voxels.cpp:278-346 (voxels.cpp itself cannot be linked; its loop body only calls
// noise_gen_3d / glm which ARE the reference's code)
void ref_noise3d_rdata(int rs1, int rs2, float mag, float freq, float *rdata420) {
	noise_gen_3d ngen; ngen.set_rand_seeds(rs1, rs2); ngen.gen_sines(mag, freq);
	memcpy(rdata420, ngen.rdata, sizeof(ngen.rdata));
}
float ref_noise3d_point(int rs1, int rs2, float mag, float freq, float x, float y, float z) {
	noise_gen_3d ngen; ngen.set_rand_seeds(rs1, rs2); ngen.gen_sines(mag, freq);
	return ngen.get_val(point(x, y, z));
}
void ref_voxel_fill(unsigned nx, unsigned ny, unsigned nz, const float lo_pos[3], const float vsz[3], const float offset[3],
	float mag, float freq, int normalize_to_1, int rs1, int rs2, int gen_mode, float zscale, float *out)
{
	unsigned const xyz_num[3] = {nx, ny, nz};
	vector<float> xyz_vals[3];
	noise_gen_3d ngen;
	float rx(0.0), ry(0.0);
	point const lo(lo_pos[0], lo_pos[1], lo_pos[2]);
	vector3d const off(offset[0], offset[1], offset[2]), vs(vsz[0], vsz[1], vsz[2]);
	if (gen_mode == MGEN_SINE) {
		ngen.set_rand_seeds(rs1, rs2);
		ngen.gen_sines(mag, freq);
		ngen.gen_xyz_vals((lo + off), vs, xyz_num, xyz_vals);
	}
	else {gen_rx_ry(rx, ry);}
#pragma omp parallel for schedule(static,1)
	for (int y = 0; y < (int)ny; ++y) {
		for (unsigned x = 0; x < nx; ++x) {
			for (unsigned z = 0; z < nz; ++z) {
				float val(0.0);
				if (gen_mode == MGEN_SINE) {val = ngen.get_val(x, y, z, xyz_vals);}
				else {
					point const pos((point(x, y, z)*vs + lo) + off);
					float nmag(mag), nfreq(0.25*freq);
					float const lacunarity(1.92), gain(0.5);
					for (int n = 0; n < max(1, ((int)MAX_FREQ_BINS - mesh_freq_filter)); ++n) {
						float const nvx(nfreq*pos.x + rx), nvy(nfreq*pos.y + ry), nvz(nfreq*pos.z + (rx-ry));
						val   += nmag*((gen_mode == MGEN_PERLIN) ? ref_glm_perlin3(nvx, nvy, nvz) : ref_glm_simplex3(nvx, nvy, nvz));
						nmag  *= gain;
						nfreq *= lacunarity;
					}
				}
				val += z*zscale;
				if (normalize_to_1) {val = CLIP_TO_pm1(val);}
				out[z + (x + (size_t)y*nx)*nz] = val;
			}
		}
	}
}

} // extern "C"

// ---- ground-mode driver: the reference's own gen_mesh() (src/mesh_gen.cpp:257-355) = sine table + estimate_zminmax + glaciate + apply_erosion ----
void gen_mesh(int surface_type, int keep_sin_table, int update_zvals);
extern float **mesh_height;
extern "C" {
void set_scene_constants_stub2();
// call after ref_setup(.., gen_sine_table=0): gen_mesh() regenerates the sine table itself (keep_sin_table=0), exactly as gen_scene() does
void ref_gen_mesh(unsigned erosion_iters_, const ref_erosion_t *p, float *out, float *zvals6) { // out: MESH_Y_SIZE*MESH_X_SIZE floats; zvals6: zmin, zmax, zmax_est, zbottom, ztop, water_plane_z
	extern unsigned erosion_iters;
	set_scene_constants_stub2();
	erosion_iters = erosion_iters_;
	erode_amount = p->erode_amount; relh_adj_tex_stub = p->relh_adj_tex; clip_hd1_stub = p->clip_hd1; // water_plane_z, zmin, zmax are set by gen_mesh itself (set_zvals)
	gen_mesh(0, 0, 1);
	memcpy(out, mesh_height[0], (size_t)MESH_X_SIZE*MESH_Y_SIZE*sizeof(float));
	zvals6[0] = zmin; zvals6[1] = zmax; zvals6[2] = zmax_est; zvals6[3] = zbottom; zvals6[4] = ztop; zvals6[5] = water_plane_z;
}
}
