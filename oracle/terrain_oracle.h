/* TEST INFRASTRUCTURE ONLY - CPU restatement ("port") of the reference algorithm for the terrain hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (3dworld_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference has NO tests or golden vectors for this path, so this oracle is pinned
 * against outputs of the reference itself - the unmodified reference objects built into oracle/_ref/libref3dworld.so by
 * oracle/refbuild/build_ref.sh - bit for bit (tests/test_oracle_vs_reference.py), and against the committed fixtures under
 * tests/golden/ that were generated from that library (tests/golden/make_golden.py).
 *
 * Plain C, fp32, compiled with -ffp-contract=off so every a*b+c is two roundings exactly like the reference build
 * (makefile:11 "-O3 -fopenmp", no -march => no FMA). Each function cites the reference file:line it follows.
 */
#ifndef TERRAIN_ORACLE_H
#define TERRAIN_ORACLE_H
#include "../include/tw3d.h"   /* POD parameter structs only (types, no code) */

#ifdef __cplusplus
extern "C" {
#endif

/* rand_gen_t, src/rand_gen.h:22-26,66-70,86,90; src/gen_object.cpp:377-381 */
void   to_rng_set(tw_rng *r, long s1, long s2);
int    to_rng_rand(tw_rng *r);
double to_rng_randd(tw_rng *r);
float  to_rng_rand_float(tw_rng *r);
float  to_rng_rand_uniform(tw_rng *r, float a, float b);

void  to_build_sin_table(float *tab65536);                                   /* src/mesh_gen.cpp:72-81 */
float to_sinf_lut(const float *tab, float v);                                /* SINF, src/sinf.h:13-14 */
float to_cosf_lut(const float *tab, float v);                                /* COSF, src/sinf.h:15 */
int   to_compute_scale(float mesh_scale, int mesh_freq_filter);              /* src/mesh_gen.cpp:544-548 */
void  to_gen_sine_params(tw_rng *rgen, float scaled_height, int mesh_x_size, int mesh_y_size, float x_scene_size, float y_scene_size,
                         int mesh_seed, int mesh_rgen_index, int mesh_gen_mode, float mesh_start_mag, float mesh_start_freq,
                         float mesh_mag_mult, float mesh_freq_mult, float *sine_params450); /* src/mesh_gen.cpp:213-254 */
void  to_gen_rx_ry(int mesh_seed, int mesh_rgen_index, int mesh_gen_mode, float *rx, float *ry); /* src/mesh_gen.cpp:581-586 */
float to_water_z_height(float zmax_est, int glaciate, float custom_glaciate_exp, float water_h_off, float water_h_off_rel); /* :507-512 */

/* GLM 0.9.9.1 gtc/noise, dependencies/glm/glm/gtc/noise.inl:25-62,66-133,592-646,649-721 */
float to_simplex2(float x, float y);
float to_perlin2(float x, float y);
float to_simplex3(float x, float y, float z);
float to_perlin3(float x, float y, float z);

/* get_noise_zval, src/mesh_gen.cpp:734-751 (rx, ry hoisted) */
float to_get_noise_zval(float xval, float yval, const tw_height_params *p);
/* eval_mesh_sin_terms, src/mesh_gen.cpp:797-805 */
float to_eval_mesh_sin_terms(float xv, float yv, const float *sin_table, const float *sine_params450, int start_eval_sin);

/* build_arrays + [enable_glaciate] + eval_index(x,y,min_start_sin) for every cell; out[y*nx+x]. nthreads<=0: all cores. */
void to_heightgen_2d(const tw_grid2d *g, const tw_height_params *p, const float *sin_table, const float *sine_params450,
                     int enable_glaciate, int min_start_sin, float *out, int nthreads);

/* glaciate() of the ground-mode mesh (src/mesh_gen.cpp:388-404: apply_glaciate + apply_mesh_sine per cell, zbottom/ztop), in place */
void to_glaciate_mesh(float *mesh, int nx, int ny, int xoff2, int yoff2, int mesh_x_size, int mesh_y_size, const tw_height_params *p,
                      const float *sin_table, float *zbottom, float *ztop);
/* gen_mesh(surface_type=0, keep_sin_table=0, update_zvals=1) in ground mode (src/mesh_gen.cpp:257-355): sine-table entries, mesh fill,
 * estimate_zminmax (:447-485), set_zvals (:494-504), glaciate, apply_erosion. p_inout->zmax_est is an output. zvals6: zmin, zmax,
 * zmax_est, zbottom, ztop, water_plane_z as the globals stand afterwards. ep_partial: erode_amount, relh_adj_tex, clip_hd1, half_dxy are
 * inputs; water_plane_z/zmin/zmax are filled from set_zvals. */
void to_gen_mesh(tw_rng *sine_rng, tw_height_params *p_inout, int mesh_x_size, int mesh_y_size, float x_scene_size, float y_scene_size,
                 int mesh_seed, int mesh_rgen_index, int xoff2, int yoff2, float dx_val, float dy_val, float water_h_off, float water_h_off_rel,
                 unsigned erosion_iters, tw_erosion_params *ep_partial, const float *sin_table, float *sine_params450_out, float *mesh_out, float *zvals6);
/* 4x4 sub-block z ranges + water bbox of tile_t::create_zvals (src/tiled_mesh.cpp:517-540); out: ntiles x tw_tile_bounds */
void to_tile_bounds(const float *zvals, unsigned ntiles, unsigned zvsize, float wpz_max, float dx_val, float dy_val, unsigned size, tw_tile_bounds *out);

/* apply_erosion, src/erosion.cpp:14-164 (serial droplet order) ; returns total droplet steps */
/* tile_t::upload_normal_texture / calc_mesh_ao_lighting (src/tiled_mesh.cpp:865-880, 586-662); czv = generated context grids (stride+72)^2 per tile */
void to_tile_normals(const float *zvals, unsigned ntiles, unsigned zvsize, float dx_val, float dy_val, unsigned char *rgba, float *min_normal_z);
void to_tile_ao(const float *zvals, const float *czv, unsigned ntiles, unsigned zvsize, float half_dxy, int use_ao_zvals, unsigned char *ao);
/* heightmap-texture mode of tile_t::create_zvals: terrain_hmap_manager_t::get_clamped_height over tiles (src/heightmap.cpp:385-402) */
void to_hmap_sample_tiles(const unsigned char *data16, const tw_hmap_sampler *H, const int *origins_xy, unsigned ntiles, unsigned zvsize, float *out);
/* eval_mesh_sin_terms / eval_mesh_sin_terms_scaled / get_exact_zval (procedural branch) for n points, src/mesh_gen.cpp:797-847 */
void to_eval_points(const float *xy, size_t n, const tw_height_params *p, const tw_point_query *q, const float *sin_table, const float *sine_params450, float *out);
unsigned long long to_apply_erosion(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *p);
unsigned long long to_erode_sweeps(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *p, unsigned sweep, int halo);

/* the protocol of the product's speculative serial-order erosion (M_SPEC) as a sequential model: must equal to_apply_erosion for every window / cap / tile size / schedule */
unsigned long long to_erode_spec_model(float *heightmap, int xsize, int ysize, float min_zval, unsigned num_iters, const tw_erosion_params *ep,
                                       unsigned B, unsigned cap, int tile_shift, unsigned sched_seed, unsigned long long *stats);

/* noise_gen_3d, src/upsurface.cpp:16-85 */
void  to_noise3d_gen_sines(int rs1, int rs2, float mag, float freq, float *rdata420);
float to_noise3d_get_val_pt(const float *rdata420, const float *sin_table, float x, float y, float z);
/* voxel_manager::create_procedural fill loop (+ optional atten pass), src/voxels.cpp:278-346,403-482 */
void  to_voxel_fill(const tw_voxel_params *vp, const float *rdata420, const float *sin_table, float *out, int nthreads);

/* heightmap_t::from_floats / to_floats 16-bit, src/heightmap.cpp:191-215 + src/Textures.cpp:1889-1893; returns count of out-of-range values */
size_t to_from_floats_u16(const float *vals, size_t n, float val_mult, float val_add, unsigned char *out2n);
void   to_to_floats_u16(const unsigned char *data2n, size_t n, float val_mult, float val_add, float *vals);

/* mesh shadows, ref: src/visibility.cpp:411-517, src/Math3d.cpp:1029-1086, src/tiled_mesh.cpp:664-692 */
void to_calc_mesh_shadows(const tw_shadow_params *sp, const float *mh, unsigned char *smask, int xsize, int ysize, const float *sh_in_x, const float *sh_in_y,
                          float *sh_out_x, float *sh_out_y);
void to_tile_shadows_batch(const float *zvals, const int *tile_xy, unsigned ntiles, unsigned zvsize, const tw_shadow_params *sp, unsigned char *smask,
                           float *sh_out_x, float *sh_out_y);
/* terrain weights texture, ref: src/tiled_mesh.cpp:1071-1248 (terrain part), src/Textures.cpp:1289-1312; rand = the un-scaled jitter noise grids */
void to_tile_weights(const float *zvals, const float *rand, unsigned ntiles, unsigned zvsize, const float *tile_params, const tw_weight_params *wp, unsigned char *rgba,
                     unsigned char *has_any_grass);
/* voxel post-processing, ref: src/voxels.cpp:485-610,739-868 */
void to_voxel_outside(const float *vals, const tw_voxel_post_params *vp, const unsigned *zix_xy, unsigned char *outside);
unsigned long long to_voxel_remove_unconnected(float *vals, unsigned char *outside, const tw_voxel_post_params *vp);
unsigned long long to_voxel_triangles(const float *vals, const unsigned char *outside, const tw_voxel_post_params *vp, const unsigned *edge_table, const int *tri_table,
                                      const unsigned *edge_to_vals, float *tris, unsigned long long capacity);

#ifdef __cplusplus
}
#endif
#endif
