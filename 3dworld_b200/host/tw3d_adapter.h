// tw3d_adapter.h - C++ host adapter: the reference's own call surface for the terrain path, implemented on the C ABI (include/tw3d.h).
//
//   tw3d::mesh_xy_grid_cache_t   <->  mesh_xy_grid_cache_t                    src/mesh.h:22-45, src/mesh_gen.cpp:588-650,754-792
//   tw3d::apply_erosion          <->  apply_erosion(float*,int,int,float,unsigned)   src/function_registry.h:354, src/erosion.cpp:14
//   tw3d::noise_gen_3d           <->  noise_gen_3d::{set_rand_seeds,gen_sines}        src/upsurface.h:39-50
//   tw3d::create_procedural      <->  voxel_manager::create_procedural               src/voxels.h:196, src/voxels.cpp:278-346
//   tw3d::create_zvals_batch     <->  the height fill + erosion of tile_t::create_zvals for many tiles   src/tiled_mesh.cpp:467-515
//
// The reference reads ~20 globals on this path; here they are one explicit struct (scene_globals) set once per scene
// with set_globals(). Same names, argument meaning and error behaviour as the reference: argument errors assert/abort like the
// reference's assert()s (define TW3D_NO_ABORT to get a tw3d::error exception instead). Header-only; link with -l3dworld_b200.
// No CPU fallback: all grid evaluation happens on the GPU through the C ABI; without a device ctx() fails.
#pragma once
#include <tw3d.h>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace tw3d {

struct error : std::runtime_error {
	int status;
	error(int s, std::string const &m) : std::runtime_error(m), status(s) {}
};

// every reference global the path reads, under the reference's own names
struct scene_globals {
	int   mesh_gen_mode = TW_MGEN_SINE, mesh_gen_shape = 0, start_eval_sin = 0, GLACIATE = 1, mesh_seed = 0, mesh_rgen_index = 0;
	float mesh_scale = 1.0f, mesh_scale_z_inv = 1.0f, DX_VAL_INV = 16.0f, DY_VAL_INV = 16.0f, MESH_HEIGHT = 0.4f, mesh_height_scale = 1.0f;
	float zmax_est = 1.0f, custom_glaciate_exp = 0.0f;
	tw_hmap_params hmap_params = {1000.0f, 0, 0, 0, 1000.0f, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	int   MESH_X_SIZE = 128, MESH_Y_SIZE = 128;
	float X_SCENE_SIZE = 4.0f, Y_SCENE_SIZE = 4.0f; // get_exact_zval (src/mesh_gen.cpp:818-819)
	int   xoff2 = 0, yoff2 = 0;                      // current mesh scroll offset (src/mesh_gen.cpp:826-829)
	float mesh_file_scale = 1.0f, mesh_file_tz = 0.0f; // scale_mh_texture_val (src/mesh_gen.cpp:120), heightmap-texture tiles
	// erosion (src/erosion.cpp:11,98; src/Textures.cpp:1284-1287)
	float erode_amount = 1.0f, water_plane_z = 0.0f, HALF_DXY = 0.0625f, zmin = -1.0f, zmax = 1.0f, relh_adj_tex = 0.0f, clip_hd1 = 0.5f;
	float zbottom = 0.0f, ztop = 0.0f;                  // set_zvals (src/mesh_gen.cpp:494-504); gen_mesh leaves them alone on a flat mesh, as the reference does
	// sine-table recurrence of gen_mesh (src/mesh_gen.cpp:34,226-231; config keywords mesh_start_mag / mesh_start_freq / mesh_mag_mult / mesh_freq_mult)
	float MESH_START_MAG = 0.02f, MESH_START_FREQ = 240.0f, MESH_MAG_MULT = 2.0f, MESH_FREQ_MULT = 0.5f;
};

namespace detail {
	inline void fail(int status, const char *what, tw_ctx *c) {
		std::string msg = std::string(what) + ": " + (c ? tw_last_error(c) : "no context");
#ifdef TW3D_NO_ABORT
		throw error(status, msg);
#else
		if (status == TW_ERR_ARG) {fprintf(stderr, "tw3d: assertion failed: %s\n", msg.c_str()); abort();} // the reference asserts
		throw error(status, msg);
#endif
	}
	struct state_t {
		scene_globals g;
		std::vector<float> sin_table, sine_params;
		unsigned generation = 0; // bumped by set_globals WHEN IT IS GIVEN A TABLE, so that thread-local contexts re-upload the tables (scalars are read per call)
	};
	inline state_t &state() {static state_t s; return s;}
	struct tls_ctx {
		tw_ctx *c = nullptr; unsigned generation = ~0u;
		~tls_ctx() {if (c) tw_destroy(c);}
	};
}

// thread-local context (a tw_ctx is not re-entrant); device from $TW3D_DEVICE (default 0)
inline tw_ctx *ctx() {
	static thread_local detail::tls_ctx t;
	detail::state_t &s = detail::state();
	if (!t.c) {
		const char *dev = getenv("TW3D_DEVICE");
		int const rc = tw_create(dev ? atoi(dev) : 0, &t.c);
		if (rc != TW_OK) {t.c = nullptr; throw error(rc, "tw_create failed: no CUDA device (lib3dworld_b200 has no CPU fallback)");}
	}
	if (t.generation != s.generation) {
		int rc = tw_set_sin_table(t.c, s.sin_table.empty() ? nullptr : s.sin_table.data());
		if (rc == TW_OK && !s.sine_params.empty()) {rc = tw_set_sine_params(t.c, s.sine_params.data());}
		if (rc != TW_OK) {detail::fail(rc, "table upload", t.c);}
		t.generation = s.generation;
	}
	return t.c;
}

// sin_table: the reference's sin_table.data() (2*TSIZE floats) or nullptr to build it; sinTable: &sinTable[0][0] (90*5 floats) or nullptr
inline void set_globals(scene_globals const &g, const float *sin_table = nullptr, const float *sinTable = nullptr) {
	detail::state_t &s = detail::state();
	s.g = g;
	if (sin_table) {s.sin_table.assign(sin_table, sin_table + TW_SIN_TABLE_SIZE);}
	if (sinTable)  {s.sine_params.assign(sinTable, sinTable + TW_F_TABLE_SIZE*5);}
	if (sin_table || sinTable) {++s.generation;}
}
inline scene_globals const &globals() {return detail::state().g;}

inline tw_height_params height_params_from_globals(int gen_mode, int gen_shape) {
	scene_globals const &g = globals();
	tw_height_params p;
	memset(&p, 0, sizeof(p));
	p.gen_mode = gen_mode; p.gen_shape = gen_shape; p.start_eval_sin = g.start_eval_sin; p.glaciate = g.GLACIATE;
	p.mesh_scale = g.mesh_scale; p.mesh_scale_z_inv = g.mesh_scale_z_inv; p.dx_val_inv = g.DX_VAL_INV; p.dy_val_inv = g.DY_VAL_INV;
	p.mesh_height = g.MESH_HEIGHT; p.mesh_height_scale = g.mesh_height_scale; p.zmax_est = g.zmax_est; p.custom_glaciate_exp = g.custom_glaciate_exp;
	tw_gen_rx_ry(g.mesh_seed, g.mesh_rgen_index, gen_mode, &p.rx, &p.ry); // gen_rx_ry(), src/mesh_gen.cpp:581-586
	p.hmap = g.hmap_params;
	return p;
}
inline tw_erosion_params erosion_params_from_globals() {
	scene_globals const &g = globals();
	tw_erosion_params e = {g.erode_amount, g.water_plane_z, g.HALF_DXY, g.zmin, g.zmax, g.relh_adj_tex, g.clip_hd1};
	return e;
}

// ------------------------------------------------------------------------------------------------ mesh_xy_grid_cache_t
// Same interface and call order as the reference (build_arrays, then optionally enable_glaciate, then eval_index per cell). The grid is
// evaluated on the GPU as a whole the first time a value is needed (or asynchronously when no_wait is set, mirroring the GLSL path:
// build_arrays returns 0 while the job is in flight, src/mesh_gen.cpp:597-603) and eval_index reads the host copy.
class mesh_xy_grid_cache_t {
	mutable std::vector<float> vals;
	mutable std::mutex mtx;
	mutable bool have_vals = false, job_running = false;
	mutable tw_ctx *job_ctx = nullptr;      // the context the pending job was launched on: a job launched on the main thread (build_arrays no_wait +
	                                        // enable_glaciate) is collected on THAT context even when eval_index runs on an OpenMP worker, whose own
	                                        // thread-local context has nothing pending and would report TW_OK at once
	mutable int vals_min_start = 0; mutable bool vals_glaciate = false;
	tw_grid2d grid = {0, 0, 1, 1, 0, 0};
	int gen_mode = TW_MGEN_SINE, gen_shape = 0;
	bool do_glaciate = false, async_requested = false;

	void launch(bool glaciate, int min_start_sin, bool wait) const {
		tw_height_params const p = height_params_from_globals(gen_mode, gen_shape);
		vals.resize((size_t)grid.nx*grid.ny);
		tw_ctx *c = ctx();
		int rc = tw_heightgen_2d_launch(c, &grid, &p, glaciate, min_start_sin, vals.data(), nullptr);
		if (rc != TW_OK) {detail::fail(rc, "build_arrays", c);}
		job_running = true; job_ctx = c; vals_glaciate = glaciate; vals_min_start = min_start_sin;
		if (wait) {collect(true);}
	}
	bool collect(bool wait) const { // always called with mtx held
		tw_ctx *c = job_ctx ? job_ctx : ctx();
		int const rc = tw_heightgen_2d_poll(c, wait ? 1 : 0);
		if (rc == TW_ERR_NOT_READY) return false;
		if (rc != TW_OK) {detail::fail(rc, "eval_index", c);}
		job_running = false; have_vals = true; job_ctx = nullptr;
		return true;
	}
public:
	bool build_arrays(float x0, float y0, float dx, float dy, unsigned nx, unsigned ny, bool cache_values=0, bool force_sine_mode=0, bool no_wait=0) {
		assert(nx > 0 && ny > 0); // src/mesh_gen.cpp:589
		std::lock_guard<std::mutex> lock(mtx);
		scene_globals const &g = globals();
		tw_grid2d const ng = {x0, y0, dx, dy, nx, ny};
		int const mode = (force_sine_mode ? (int)TW_MGEN_SINE : g.mesh_gen_mode), shape = (force_sine_mode ? 0 : g.mesh_gen_shape);
		bool const same = (memcmp(&ng, &grid, sizeof(grid)) == 0 && mode == gen_mode && shape == gen_shape);
		if (job_running && !same) {collect(true);} // a different grid was in flight: drain it
		if (!same) {have_vals = false;}
		grid = ng; gen_mode = mode; gen_shape = shape;
		do_glaciate = 0; // must call enable_glaciate() after this call if needed (src/mesh_gen.cpp:594)
		async_requested = false;
		if (gen_mode >= TW_MGEN_SIMPLEX_GPU && no_wait) { // GPU modes: launch now or report progress (src/mesh_gen.cpp:597-603)
			if (job_running) {return collect(false);}
			if (have_vals && vals_glaciate && vals_min_start == 0) return 1;
			async_requested = true; // launched by enable_glaciate(), which setup_height_gen_async always calls next (src/tiled_mesh.cpp:462)
			return 0;
		}
		// cache_values: the reference fills cached_vals with un-glaciated values here (src/mesh_gen.cpp:627-636); every caller then calls enable_glaciate()
		// and reads through eval_index, which applies the glaciation per cell - so the grid is evaluated once, lazily, in the state eval_index asks for,
		// instead of once un-glaciated now and once more glaciated on the first eval_index
		(void)cache_values;
		return 1;
	}
	void enable_glaciate() {
		std::lock_guard<std::mutex> lock(mtx);
		do_glaciate = 1;
		if (async_requested && !job_running) {have_vals = false; launch(true, 0, false); async_requested = false;}
	}
	float eval_index(unsigned x, unsigned y, int min_start_sin=0, bool use_cache=1) const {
		assert(x < grid.nx && y < grid.ny); // src/mesh_gen.cpp:756
		(void)use_cache;
		int const mss = (gen_mode == TW_MGEN_SINE) ? min_start_sin : 0;
		bool ready;
		{std::lock_guard<std::mutex> lock(mtx); ready = (have_vals && !job_running && vals_glaciate == do_glaciate && vals_min_start == mss);} // flags are only read under the mutex
		if (!ready) {
			std::lock_guard<std::mutex> lock(mtx);
			if (job_running) {collect(true);}
			if (!(have_vals && vals_glaciate == do_glaciate && vals_min_start == mss)) {have_vals = false; launch(do_glaciate, mss, true);}
		}
		return vals[(size_t)y*grid.nx + x];
	}
	// whole-grid accessors (what the OpenMP eval_index loops of the reference's callers produce)
	void get_grid(float *out, int min_start_sin=0) const {
		(void)eval_index(0, 0, min_start_sin);
		memcpy(out, vals.data(), vals.size()*sizeof(float));
	}
	void clear_context() {std::lock_guard<std::mutex> lock(mtx); if (job_running) {collect(true);} have_vals = false; vals.clear();}
	void free_cshader() {}
	~mesh_xy_grid_cache_t() {if (job_running) {try {collect(true);} catch (...) {}}}
};

// apply_erosion(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters), src/function_registry.h:354
inline void apply_erosion(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters) {
	tw_erosion_params const e = erosion_params_from_globals();
	if (num_iters == 0 || e.erode_amount <= 0.0) return; // erosion disabled (src/erosion.cpp:16)
	tw_ctx *c = ctx();
	int const rc = tw_erode(c, heightmap, xsize, ysize, min_zval, num_iters, &e);
	if (rc != TW_OK) {detail::fail(rc, "apply_erosion", c);}
}

// ------------------------------------------------------------------------------------------------ point queries (batched)
// float get_exact_zval(float xval, float yval, bool no_xyoff=0) / eval_mesh_sin_terms(xv, yv) / eval_mesh_sin_terms_scaled(xval, yval, xy_scale)
// (src/function_registry.h:340-341, src/mesh_gen.cpp:797-847) for n points at once: xy = n (x, y) pairs. The single-point forms below cost a
// kernel launch per call - callers that place many objects (buildings, scenery) should collect their points and use the batch forms.
inline void eval_points(int kind, const float *xy, size_t n, float *out, float xy_scale = 1.0f, bool no_xyoff = false) {
	scene_globals const &g = globals();
	tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_point_query q;
	q.kind = kind; q.xy_scale = xy_scale; q.mesh_x_size = g.MESH_X_SIZE; q.mesh_y_size = g.MESH_Y_SIZE;
	q.x_scene_size = g.X_SCENE_SIZE; q.y_scene_size = g.Y_SCENE_SIZE; q.xoff2 = g.xoff2; q.yoff2 = g.yoff2; q.no_xyoff = no_xyoff;
	tw_ctx *c = ctx();
	int const rc = tw_eval_points(c, xy, n, &p, &q, out);
	if (rc != TW_OK) {detail::fail(rc, "eval_points", c);}
}
inline void get_exact_zvals(const float *xy, size_t n, float *zvals_out, bool no_xyoff = false) {eval_points(TW_PQ_EXACT_ZVAL, xy, n, zvals_out, 1.0f, no_xyoff);}
inline float get_exact_zval(float xval, float yval, bool no_xyoff = false) {
	float const xy[2] = {xval, yval}; float z = 0.0f;
	get_exact_zvals(xy, 1, &z, no_xyoff);
	return z;
}
inline float eval_mesh_sin_terms(float xv, float yv) {
	float const xy[2] = {xv, yv}; float z = 0.0f;
	eval_points(TW_PQ_SIN_TERMS, xy, 1, &z);
	return z;
}
inline float eval_mesh_sin_terms_scaled(float xval, float yval, float xy_scale) {
	float const xy[2] = {xval, yval}; float z = 0.0f;
	eval_points(TW_PQ_SIN_TERMS_SCALED, xy, 1, &z, xy_scale);
	return z;
}

// the same with the reference's OpenMP semantics (`#pragma omp parallel for schedule(dynamic,1)`, src/erosion.cpp:66): num_threads droplets in
// flight on the one heightmap, order-dependent result like the reference's; num_threads = 1 equals apply_erosion(), 0 = fill the GPU
inline void apply_erosion_parallel(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, unsigned num_threads = 0) {
	tw_erosion_params const e = erosion_params_from_globals();
	if (num_iters == 0 || e.erode_amount <= 0.0) return;
	tw_ctx *c = ctx();
	int const rc = tw_erode_parallel(c, heightmap, xsize, ysize, min_zval, num_iters, &e, num_threads);
	if (rc != TW_OK) {detail::fail(rc, "apply_erosion_parallel", c);}
}

// Height fill + per-tile erosion of tile_t::create_zvals for a batch of tiles (origins = tile x1,y1 pairs; zvals_out = ntiles*zvsize^2 floats)
inline void create_zvals_batch(const int32_t *origins_xy, unsigned ntiles, unsigned zvsize, float dx, float dy, unsigned erosion_iters_tt, float *zvals_out, tw_minmax *mm = nullptr) {
	scene_globals const &g = globals();
	tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_erosion_params const e = erosion_params_from_globals();
	tw_ctx *c = ctx();
	// apply_erosion(zvals.data(), zvsize, zvsize, zmin, erosion_iters_tt): min_zval is the global zmin (src/tiled_mesh.cpp:515)
	int const rc = tw_create_zvals_batch(c, origins_xy, ntiles, g.MESH_X_SIZE, g.MESH_Y_SIZE, dx, dy, zvsize, &p, erosion_iters_tt, &e, g.zmin, zvals_out, mm);
	if (rc != TW_OK) {detail::fail(rc, "create_zvals_batch", c);}
}

// heightmap-texture mode of tile_t::create_zvals (src/tiled_mesh.cpp:498-501): zvals[y*zvsize + x] = terrain_hmap_manager.get_clamped_height(x1 + x, y1 + y)
// for a batch of tiles, from the 16-bit image hmap16 (width*height*2 bytes, the layout heightmap_t keeps); TEX_EDGE_MODE 2 = mirror as compiled in the reference
inline void create_zvals_from_heightmap(const uint8_t *hmap16, int width, int height, const int32_t *origins_xy, unsigned ntiles, unsigned zvsize, float *zvals_out, int tex_edge_mode = TW_HMAP_EDGE_MIRROR) {
	scene_globals const &g = globals();
	tw_hmap_sampler hs;
	hs.width = width; hs.height = height; hs.edge_mode = tex_edge_mode; hs.mesh_scale = g.mesh_scale;
	hs.h_scale = 0.0008f*g.mesh_height_scale; // READ_MESH_H_SCALE*mesh_height_scale (src/mesh_gen.cpp:22,120)
	hs.mesh_file_scale = g.mesh_file_scale; hs.mesh_file_tz = g.mesh_file_tz; hs.mesh_scale_z_inv = g.mesh_scale_z_inv;
	tw_ctx *c = ctx();
	int const rc = tw_heightmap_sample_tiles(c, hmap16, &hs, origins_xy, ntiles, zvsize, zvals_out);
	if (rc != TW_OK) {detail::fail(rc, "create_zvals_from_heightmap", c);}
}

// tile_t::upload_normal_texture (src/tiled_mesh.cpp:865-880, minus the GL upload) and tile_t::calc_mesh_ao_lighting (:586-662) for a batch of
// finished tiles: normal_data = ntiles*stride^2*4 bytes (RGBA, alpha 0), ao_lighting = ntiles*stride^2 bytes, stride = zvsize-1
inline void tile_normals(const float *zvals, unsigned ntiles, unsigned zvsize, float dx_val, float dy_val, unsigned char *normal_data, float *min_normal_z = nullptr) {
	tw_ctx *c = ctx();
	int const rc = tw_tile_normals_batch(c, zvals, ntiles, zvsize, dx_val, dy_val, normal_data, min_normal_z);
	if (rc != TW_OK) {detail::fail(rc, "tile_normals", c);}
}
// tile_t::create_zvals + calc_mesh_ao_lighting with enable_tiled_mesh_ao for a batch: in the GPU gen modes the (stride+72)^2 context is generated once, zvals are
// its interior (src/tiled_mesh.cpp:479-487,505) and the AO rays test the un-eroded context (:604)
inline void create_zvals_with_ao(const int32_t *origins_xy, unsigned ntiles, unsigned zvsize, float dx, float dy, unsigned erosion_iters_tt, float *zvals_out, unsigned char *ao_lighting, tw_minmax *mm = nullptr) {
	scene_globals const &g = globals();
	tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_erosion_params const e = erosion_params_from_globals();
	tw_ctx *c = ctx();
	int const rc = tw_create_zvals_ao_batch(c, origins_xy, ntiles, g.MESH_X_SIZE, g.MESH_Y_SIZE, dx, dy, zvsize, &p, erosion_iters_tt, &e, g.zmin, g.HALF_DXY, zvals_out, ao_lighting, mm);
	if (rc != TW_OK) {detail::fail(rc, "create_zvals_with_ao", c);}
}
inline void tile_ao_lighting(const float *zvals, const int32_t *origins_xy, unsigned ntiles, unsigned zvsize, float dx, float dy, unsigned char *ao_lighting) {
	scene_globals const &g = globals();
	tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_ctx *c = ctx();
	int const rc = tw_tile_ao_batch(c, zvals, origins_xy, ntiles, g.MESH_X_SIZE, g.MESH_Y_SIZE, dx, dy, zvsize, &p, g.HALF_DXY, ao_lighting);
	if (rc != TW_OK) {detail::fail(rc, "tile_ao_lighting", c);}
}

// calc_mesh_shadows(l, lpos, mh, smask, xsize, ysize, sh_in_x, sh_in_y, sh_out_x, sh_out_y) (src/visibility.cpp:508-517) for ALL tiles a light change dirties, in
// one call: tile_t::calc_shadows_for_light's chain (src/tiled_mesh.cpp:664-692: each tile starts from the sh_out of its neighbours toward the light) becomes dependency
// waves on the device. tile_xy = (x1/size, y1/size) per tile; smask = ntiles*zvsize^2 bytes (0 / MESH_SHADOW); sh_out_* optional (ntiles*zvsize floats each).
// no_shadow = (l == LIGHT_MOON && combined_gu). Globals read: X/Y_SCENE_SIZE, DX/DY_VAL(+_INV), XY_SUM_SIZE = MESH_X_SIZE + MESH_Y_SIZE, zmin, zmax.
inline void calc_mesh_shadows(const float lpos[3], const float *zvals, const int32_t *tile_xy, unsigned ntiles, unsigned zvsize, float dx_val, float dy_val,
                              unsigned char *smask, float *sh_out_x = nullptr, float *sh_out_y = nullptr, bool no_shadow = false) {
	scene_globals const &g = globals();
	tw_shadow_params sp;
	for (int d = 0; d < 3; ++d) {sp.lpos[d] = lpos[d];}
	sp.x_scene_size = g.X_SCENE_SIZE; sp.y_scene_size = g.Y_SCENE_SIZE;
	sp.dx_val = dx_val; sp.dy_val = dy_val; sp.dx_val_inv = 1.0f/dx_val; sp.dy_val_inv = 1.0f/dy_val; // set_scene_constants: DX_VAL_INV = 1.0/DX_VAL
	sp.xy_sum_size = g.MESH_X_SIZE + g.MESH_Y_SIZE;
	sp.zmin = g.zmin; sp.zmax = g.zmax; sp.no_shadow = no_shadow ? 1 : 0;
	tw_ctx *c = ctx();
	int const rc = tw_tile_shadows_batch(c, zvals, tile_xy, ntiles, zvsize, &sp, smask, sh_out_x, sh_out_y);
	if (rc != TW_OK) {detail::fail(rc, "calc_mesh_shadows", c);}
}

// tile_t::create_texture's terrain part (src/tiled_mesh.cpp:1071-1248) for a batch of tiles: mesh_weight_data (RGBA = {sand, dirt, grass, rock}, stride^2 texels per
// tile) and has_any_grass. The caller passes what the reference reads from engine tables: h_dirt[] and lttex_dirt[].id (as TW_TEX_* classes), sthresh, the biome corners
// params[y][x].{grass, dirt} of every tile (grass[4] then dirt[4]), get_water_z_height(), vegetation, relh_adj_tex, mesh_gen_shape / mesh_scale_z (noise_scale),
// water_is_lava || DISABLE_WATER == 2. Texels inside cities / over tunnels / under buildings and the tree pass stay with the caller (it overwrites them afterwards).
// (h_dirt / tex_class: the engine's own tables, or tw_gen_tex_height_tables(water_h_off_rel, temperature, glaciate_exp, ...) = init_terrain_mesh + gen_tex_height_tables)
struct weight_tables {float h_dirt[5]; int tex_class[5]; float sthresh[2][2]; float water_level, vegetation; bool snow_to_rock; int mesh_gen_shape; float mesh_scale_z;};
inline void create_texture_weights(const float *zvals, const int32_t *origins_xy, unsigned ntiles, unsigned zvsize, float dx_val, float dy_val, weight_tables const &wt,
                                   const float *tile_params, unsigned char *mesh_weight_data, unsigned char *has_any_grass = nullptr) {
	scene_globals const &g = globals();
	tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_weight_params W;
	memset(&W, 0, sizeof(W));
	for (int i = 0; i < 5; ++i) {W.h_dirt[i] = wt.h_dirt[i]; W.tex_class[i] = wt.tex_class[i];}
	for (int a = 0; a < 2; ++a) {for (int b = 0; b < 2; ++b) {W.sthresh[a][b] = wt.sthresh[a][b];}}
	W.zmin = g.zmin; W.zmax = g.zmax; W.relh_adj_tex = g.relh_adj_tex; W.water_level = wt.water_level;
	float const MESH_NOISE_SCALE = 0.003;
	W.noise_scale = ((wt.mesh_gen_shape == 2) ? 2.0 : 1.0)*MESH_NOISE_SCALE*wt.mesh_scale_z;          // src/tiled_mesh.cpp:1085-1088, same types
	float const SQRT2 = sqrt(2.0);                                                                     // src/3DWorld.h:132
	W.vnz_scale = (g.mesh_gen_mode == TW_MGEN_DWARP_GPU) ? SQRT2 : 1.0;
	W.vegetation = wt.vegetation; W.snow_to_rock = wt.snow_to_rock ? 1 : 0;
	W.dx_val = dx_val; W.dy_val = dy_val; W.dxdy = dx_val*dy_val;
	W.xy_mult = 1.0/float(zvsize - 2);                                                                 // size = zvsize - 2
	tw_ctx *c = ctx();
	int const rc = tw_tile_weights_batch(c, zvals, origins_xy, ntiles, g.MESH_X_SIZE, g.MESH_Y_SIZE, dx_val, dy_val, zvsize, &p, &W, tile_params, mesh_weight_data, has_any_grass);
	if (rc != TW_OK) {detail::fail(rc, "create_texture_weights", c);}
}

// ------------------------------------------------------------------------------------------------ gen_mesh (ground mode)
// gen_mesh(surface_type=0, keep_sin_table=0, update_zvals=1) for WMODE_GROUND (src/mesh_gen.cpp:257-355): regenerates the sine table from
// the function-static generator state (pass the same tw_rng across calls), fills mesh_height, estimates zmax_est from a 128x128 probe of the
// equation (estimate_zminmax, :447-485), sets the z globals (set_zvals, :494-504), glaciates (:388-404) and erodes (:443).
// The derived globals are written back into scene_globals (zmax_est, zmin, zmax, water_plane_z) exactly as the reference leaves them.
struct gen_mesh_result {float zmin, zmax, zmax_est, zbottom, ztop, water_plane_z;};

inline gen_mesh_result gen_mesh(float *mesh_height /* MESH_Y_SIZE x MESH_X_SIZE, row-major */, tw_rng &sine_rng, float x_scene_size, float y_scene_size,
	unsigned erosion_iters, int xoff2 = 0, int yoff2 = 0, float dx_val = 0.0f, float dy_val = 0.0f, float water_h_off = 0.0f, float water_h_off_rel = 0.0f)
{
	scene_globals g = globals();
	int const MX = g.MESH_X_SIZE, MY = g.MESH_Y_SIZE;
	if (dx_val == 0.0f) {dx_val = 1.0f/g.DX_VAL_INV;}
	if (dy_val == 0.0f) {dy_val = 1.0f/g.DY_VAL_INV;}
	std::vector<float> sinTable(TW_F_TABLE_SIZE*5);
	tw_gen_sine_params(&sine_rng, g.MESH_HEIGHT*g.mesh_height_scale, MX, MY, x_scene_size, y_scene_size, g.mesh_seed, g.mesh_rgen_index, g.mesh_gen_mode,
	                   g.MESH_START_MAG, g.MESH_START_FREQ, g.MESH_MAG_MULT, g.MESH_FREQ_MULT, sinTable.data());
	set_globals(g, nullptr, sinTable.data());
	tw_ctx *c = ctx();
	tw_height_params p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
	tw_grid2d const grid = {(float)(xoff2 - MX/2), (float)(yoff2 - MY/2), dx_val, dy_val, (uint32_t)MX, (uint32_t)MY}; // gen_mesh_sine_table, :201-210
	tw_minmax mm;
	int rc = tw_heightgen_2d(c, &grid, &p, 0, 0, mesh_height, &mm);
	if (rc != TW_OK) {detail::fail(rc, "gen_mesh", c);}
	float zmin = mm.zmin, zmax = mm.zmax;                        // calc_zminmax
	float zmax_est = (zmax < -zmin) ? -zmin : zmax;              // set_zmax_est(max(zmax, -zmin))
	gen_mesh_result r;
	if (zmax == zmin) { // flat mesh: estimate_zminmax returns BEFORE set_zvals (src/mesh_gen.cpp:463-466): zmin/zmax stay the measured pair, zbottom/ztop/water_plane_z keep their old values
		zmax_est = zmax_est + 1.0E-6;
		r.zbottom = g.zbottom; r.ztop = g.ztop; r.zmin = zmin; r.zmax = zmax; r.zmax_est = zmax_est; r.water_plane_z = g.water_plane_z;
	}
	else {
		float const XY_SCENE_SIZE(0.5f*(x_scene_size + y_scene_size));
		float const rm_scale(1000.0*XY_SCENE_SIZE/g.mesh_scale);
		tw_grid2d const probe = {0.0f, 0.0f, rm_scale, rm_scale, 128, 128};   // EST_RAND_PARAM
		std::vector<float> h(128*128);
		rc = tw_heightgen_2d(c, &probe, &p, 0, 0, h.data(), nullptr);
		if (rc != TW_OK) {detail::fail(rc, "estimate_zminmax", c);}
		for (float v : h) {float const a(std::fabs(v)); zmax_est = (zmax_est < a) ? a : zmax_est;}
		if (g.mesh_gen_mode != TW_MGEN_SINE) {zmax_est *= 1.2;}
		zmax_est = 1.1*zmax_est;
		r.zbottom = zmin; r.ztop = zmax;                              // set_zvals
		r.zmin = -zmax_est; r.zmax = zmax_est; r.zmax_est = zmax_est;
		r.water_plane_z = tw_water_z_height(zmax_est, g.GLACIATE, g.custom_glaciate_exp, water_h_off, water_h_off_rel);
	}
	g.zmax_est = zmax_est; g.zmin = r.zmin; g.zmax = r.zmax; g.water_plane_z = r.water_plane_z; g.zbottom = r.zbottom; g.ztop = r.ztop;
	set_globals(g);
	p.zmax_est = zmax_est;
	if (g.GLACIATE) { // gen_terrain_map -> glaciate()
		rc = tw_glaciate_mesh(c, mesh_height, MX, MY, xoff2, yoff2, MX, MY, &p, &mm);
		if (rc != TW_OK) {detail::fail(rc, "glaciate", c);}
		r.zbottom = mm.zmin; r.ztop = mm.zmax;                    // glaciate() recomputes them (calc_zminmax + the zbottom/ztop update, :399-403)
		g.zbottom = r.zbottom; g.ztop = r.ztop; set_globals(g);
	}
	apply_erosion(mesh_height, MX, MY, r.zbottom, erosion_iters);
	return r;
}

// tail of tile_t::create_zvals for a batch of finished tiles (sub_zmin/sub_zmax, mzmin/mzmax, mesh_dz, radius, water bbox), src/tiled_mesh.cpp:517-541
inline void tile_bounds(const float *zvals, unsigned ntiles, unsigned zvsize, float wpz_max, float dx_val, float dy_val, unsigned size, tw_tile_bounds *out) {
	tw_ctx *c = ctx();
	int const rc = tw_tile_bounds_batch(c, zvals, ntiles, zvsize, wpz_max, dx_val, dy_val, size, out);
	if (rc != TW_OK) {detail::fail(rc, "tile_bounds", c);}
}

// noise_gen_3d: the table-generation half of the reference class (src/upsurface.h:39-50); grid evaluation goes through create_procedural
class noise_gen_3d {
	int rs1 = 1, rs2 = 1;
public:
	unsigned num_sines = 0;
	float rdata[TW_N3D_RDATA] = {0};
	void set_rand_seeds(int rs1_, int rs2_) {rs1 = rs1_; rs2 = rs2_;}
	void gen_sines(float mag, float freq) {
		assert(mag > 0.0 && freq > 0.0); // src/upsurface.cpp:19
		tw_noise3d_gen_sines(rs1, rs2, mag, freq, rdata);
		num_sines = TW_N3D_SINES;
	}
};

// the part of voxel_grid<float> that create_procedural touches (src/voxels.h:100-160)
struct voxel_grid_view {
	unsigned nx, ny, nz;
	float vsz[3], lo_pos[3];
	std::vector<float> *data; // resized to nx*ny*nz, index z + (x + y*nx)*nz
};

// voxel_manager::create_procedural(mag, freq, offset, normalize_to_1, rseed1, rseed2, gen_mode, verbose) (src/voxels.cpp:278);
// zscale = (params.invert ? -1.0 : 1.0)*params.z_gradient/(nz-1), mesh_freq_filter from the scene
inline void create_procedural(voxel_grid_view const &v, float mag, float freq, const float offset[3], bool normalize_to_1, int rseed1, int rseed2,
	int gen_mode, float zscale, int mesh_freq_filter)
{
	scene_globals const &g = globals();
	tw_voxel_params vp;
	memset(&vp, 0, sizeof(vp));
	vp.nx = v.nx; vp.ny = v.ny; vp.nz = v.nz;
	for (int d = 0; d < 3; ++d) {vp.lo_pos[d] = v.lo_pos[d]; vp.vsz[d] = v.vsz[d]; vp.offset[d] = offset[d];}
	vp.mag = mag; vp.freq = freq; vp.gen_mode = gen_mode; vp.normalize_to_1 = normalize_to_1; vp.rseed1 = rseed1; vp.rseed2 = rseed2;
	vp.octaves = (5 - mesh_freq_filter > 1) ? (5 - mesh_freq_filter) : 1; // max(1, MAX_FREQ_BINS - mesh_freq_filter), src/voxels.cpp:333
	if (gen_mode != TW_MGEN_SINE) {tw_gen_rx_ry(g.mesh_seed, g.mesh_rgen_index, gen_mode, &vp.rx, &vp.ry);}
	vp.zscale = zscale;
	v.data->resize((size_t)v.nx*v.ny*v.nz);
	tw_ctx *c = ctx();
	int const rc = tw_voxel_fill(c, &vp, nullptr, v.data->data());
	if (rc != TW_OK) {detail::fail(rc, "create_procedural", c);}
}

// voxel_model::build after the fill (src/voxels.cpp:1523-1530 + create_block :1077-1108): determine_voxels_outside, remove_unconnected_outside
// (+ remove_interior_holes for remove_unconnected > 2) and the marching-cubes triangles of the whole grid, on the device. `outside` gets the
// reference's flag bytes; zix_xy (optional) = the per-column under-mesh index the reference derives from z_min_matrix (:596-600); the case tables are
// voxel_detail::edge_table / tri_table / edge_to_vals of src/marching_cubes.h. Returns the unwelded triangle soup (9 floats per triangle).
inline std::vector<float> voxel_build(voxel_grid_view const &v, std::vector<unsigned char> &outside, float isolevel, bool invert, bool make_closed_surface,
	unsigned remove_unconnected, bool keep_at_edge, bool sphere_mode_or_no_mesh, bool skip_under_mesh, const uint32_t *zix_xy,
	const unsigned *edge_table, const int *tri_table, const unsigned *edge_to_vals)
{
	tw_voxel_post_params vp;
	memset(&vp, 0, sizeof(vp));
	vp.nx = v.nx; vp.ny = v.ny; vp.nz = v.nz;
	for (int d = 0; d < 3; ++d) {vp.lo_pos[d] = v.lo_pos[d]; vp.vsz[d] = v.vsz[d];}
	vp.isolevel = isolevel; vp.invert = invert; vp.make_closed_surface = make_closed_surface; vp.remove_unconnected = (int)remove_unconnected;
	vp.keep_at_edge = keep_at_edge; vp.centre_seed = sphere_mode_or_no_mesh; vp.skip_under_mesh = skip_under_mesh;
	outside.resize(v.data->size());
	tw_ctx *c = ctx();
	int rc = tw_voxel_outside(c, v.data->data(), &vp, zix_xy, outside.data());
	if (rc == TW_OK) {rc = tw_voxel_remove_unconnected(c, v.data->data(), outside.data(), &vp, nullptr);}
	uint64_t n = 0;
	if (rc == TW_OK) {rc = tw_voxel_triangles(c, v.data->data(), outside.data(), &vp, edge_table, tri_table, edge_to_vals, nullptr, 0, &n);}
	std::vector<float> tris((size_t)n*9);
	if (rc == TW_OK && n) {rc = tw_voxel_triangles(c, v.data->data(), outside.data(), &vp, edge_table, tri_table, edge_to_vals, tris.data(), n, &n);}
	if (rc != TW_OK) {detail::fail(rc, "voxel_build", c);}
	return tris;
}

// ------------------------------------------------------------------------------------------------ all GPUs of the box (include/tw3d.h "Multi-GPU")
// One object per process: per-device contexts with the current tables, NUMA-local pinned output bands, the tile loop of tile_draw_t::update
// (src/tiled_mesh.cpp:2367-2417) dealt out over the devices, and the global z range (get_heightmap_z_range, src/map_view.cpp:399-407) reduced with NCCL.
class multi_gpu {
	tw_multi *m = nullptr;
	std::vector<void *> host_bands;
	void check(int rc, const char *what) {if (rc != TW_OK) {throw error(rc, std::string(what) + ": " + (m ? tw_multi_last_error(m) : "tw_multi_create failed"));}}
public:
	explicit multi_gpu(int ndev, const int *devices = nullptr) {
		check(tw_multi_create(devices, ndev, &m), "tw_multi_create");
		detail::state_t &s = detail::state();
		if (!s.sine_params.empty()) {check(tw_multi_set_sine_params(m, s.sine_params.data()), "tw_multi_set_sine_params");}
	}
	~multi_gpu() {for (void *p : host_bands) {tw_multi_free_host(m, p);} tw_multi_destroy(m);}
	multi_gpu(multi_gpu const &) = delete;
	multi_gpu &operator=(multi_gpu const &) = delete;
	int size() const {return tw_multi_size(m);}
	tw_multi *handle() {return m;}
	// pinned host memory on device i's NUMA node for its band of `ntiles` tiles of zvsize^2 floats (owned by this object)
	float *alloc_band(int i, uint32_t ntiles, uint32_t zvsize) {
		uint32_t a, b;
		tw_multi_range(ntiles, size(), i, &a, &b);
		void *p = nullptr;
		check(tw_multi_alloc_host(m, i, (size_t)(b - a)*zvsize*zvsize*sizeof(float), &p), "tw_multi_alloc_host");
		host_bands.push_back(p);
		return (float *)p;
	}
	// tile_t::create_zvals for all tiles, every device on its band; returns the global z range
	tw_minmax create_zvals(const int32_t *origins_xy, uint32_t ntiles, uint32_t zvsize, float dx, float dy, unsigned erosion_iters_tt, float *const *bands, tw_minmax *mm = nullptr) {
		scene_globals const &g = globals();
		tw_height_params const p = height_params_from_globals(g.mesh_gen_mode, g.mesh_gen_shape);
		tw_erosion_params const e = erosion_params_from_globals();
		tw_minmax zr = {0, 0};
		check(tw_create_zvals_sharded(m, origins_xy, ntiles, g.MESH_X_SIZE, g.MESH_Y_SIZE, dx, dy, zvsize, &p, erosion_iters_tt, &e, g.zmin, bands, mm, &zr), "tw_create_zvals_sharded");
		return zr;
	}
	// heightmap_t::run_erosion on a map held as row bands, coherent across the devices (the batched variant, see tw_erode_sweeps in tw3d.h)
	uint64_t erode_sweeps(float *const *bands, int xsize, int ysize, float min_zval, unsigned num_iters, unsigned sweep = 8192, int halo = 64) {
		tw_erosion_params const e = erosion_params_from_globals();
		uint64_t moves = 0;
		check(tw_erode_sweeps_sharded(m, bands, xsize, ysize, min_zval, num_iters, &e, sweep, halo, &moves), "tw_erode_sweeps_sharded");
		return moves;
	}
};

} // namespace tw3d
