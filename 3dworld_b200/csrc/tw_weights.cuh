// tw_weights.cuh - one texel of the terrain weights texture: the per-cell body of tile_t::create_texture (src/tiled_mesh.cpp:1140-1248) for terrain
// without cities / tunnels / buildings / trees (those branches read engine state - road networks, building footprints, the tree map - and stay with the
// caller, who overwrites the texels they cover; include/tw3d.h). The reference mixes float and double arithmetic (unsuffixed literals); every operation
// below has the type the C++ expression has there, and the file is compiled without FMA contraction. Usable from the host (tests/cpp/test_weights.cpp
// compiles it with g++ and compares it with the oracle without a GPU) and from tw_tiles.cu's kernel.
#pragma once
#include "../../include/tw3d.h"
#include <math.h>
#ifdef __CUDACC__
#define TW_HD __host__ __device__ __forceinline__
#else
#define TW_HD inline
#endif

namespace tww {

TW_HD float clip01(float x) {return fmaxf(0.0f, fminf(1.0f, x));} // CLIP_TO_01: max(0.0f, min(1.0f, x)); the arguments here are never NaN-vs-number ambiguous (std::min/max pick an operand either way)

TW_HD void update_lttex_ix(int &ix, const tw_weight_params &W) { // src/Textures.cpp:1289-1292
	if (W.snow_to_rock && W.tex_class[ix] == TW_TEX_SNOW) {--ix;}
	if (W.vegetation == 0.0f && W.tex_class[ix] == TW_TEX_GROUND) {++ix;}
}
TW_HD void get_tids(float relh, int &k1, int &k2, float *t, const tw_weight_params &W) { // src/Textures.cpp:1294-1312; TEXTURE_SMOOTH = 0.01 (:12)
	float const TEXTURE_SMOOTH = 0.01f;
	if      (relh < W.h_dirt[0]) {k1 = 0;}
	else if (relh < W.h_dirt[1]) {k1 = 1;}
	else if (relh < W.h_dirt[2]) {k1 = 2;}
	else if (relh < W.h_dirt[3]) {k1 = 3;}
	else                         {k1 = 4;}
	if (k1 < 4 && (W.h_dirt[k1] - relh) < TEXTURE_SMOOTH) {
		if (t) {*t = (float)(1.0 - (double)((W.h_dirt[k1] - relh)/TEXTURE_SMOOTH));}
		k2 = k1 + 1;
		update_lttex_ix(k1, W);
		update_lttex_ix(k2, W);
	}
	else {
		update_lttex_ix(k1, W);
		k2 = k1;
	}
}
TW_HD float bilinear(const float c[4], float x, float y) { // BILINEAR_INTERP (src/tiled_mesh.cpp:189); c = {[0][0], [0][1], [1][0], [1][1]} = [y][x]
	return (y*(x*c[3] + (1.0f - x)*c[2]) + (1.0f - y)*(x*c[1] + (1.0f - x)*c[0]));
}

// zv: the tile's zvals (zvsize = size + 2 per row); x, y < stride = size + 1; rand_offset = noise_scale*eval_index(x, y, 50) of the force-sine-mode grid;
// tile_params = the tile's biome corners: grass[4] then dirt[4] ([y][x] order); rgba = {sand, dirt, grass, rock}; returns "this texel has grass" (has_any_grass)
TW_HD bool weights_texel(const float *zv, unsigned zvsize, unsigned x, unsigned y, float rand_offset, const float *tile_params, const tw_weight_params &W, unsigned char rgba[4]) {
	unsigned const ix = y*zvsize + x;
	float const dz_inv = 1.0f/(W.zmax - W.zmin);
	float const steep_mult_grass = 1.0f/(W.sthresh[0][1] - W.sthresh[0][0]), steep_mult_snow = 1.0f/(W.sthresh[1][1] - W.sthresh[1][0]);
	float const steep_mult_rock = 1.0f/(0.8f*W.sthresh[0][0] - 0.5f*W.sthresh[0][0]);
	float weights[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
	float const mh00 = zv[ix], mh01 = zv[ix + 1], mh10 = zv[ix + zvsize], mh11 = zv[ix + zvsize + 1];
	float const mhmin = fminf(fminf(mh00, mh01), fminf(mh10, mh11)), mhmax = fmaxf(fmaxf(mh00, mh01), fmaxf(mh10, mh11));
	float const relh1 = W.relh_adj_tex + (mhmin - W.zmin)*dz_inv + rand_offset, relh2 = W.relh_adj_tex + (mhmax - W.zmin)*dz_inv + rand_offset;
	int k1, k2, k3, k4;
	get_tids(relh1, k1, k2, nullptr, W);
	get_tids(relh2, k3, k4, nullptr, W);
	bool const same_tid = (k1 == k4);
	float t = 0.0f;
	k2 = k4;
	if (!same_tid) {
		float const relh = W.relh_adj_tex + (mh00 - W.zmin)*dz_inv;
		get_tids(relh, k1, k2, &t, W);
	}
	float weight_scale = 1.0f;
	bool const grass = (W.tex_class[k1] == TW_TEX_GROUND || W.tex_class[k2] == TW_TEX_GROUND), snow = (W.tex_class[k2] == TW_TEX_SNOW);
	int const sand_ix = W.class_ix[TW_TEX_SAND], dirt_ix = W.class_ix[TW_TEX_DIRT], grass_ix = W.class_ix[TW_TEX_GROUND], rock_ix = W.class_ix[TW_TEX_ROCK];
	if (grass || snow) {
		const float *sti = W.sthresh[snow ? 1 : 0];
		float const nx = W.dy_val*(zv[ix] - zv[ix + 1]), ny = W.dx_val*(zv[ix] - zv[ix + zvsize]), nz = W.dxdy; // get_norm_not_normalized (src/tiled_mesh.h:281-283)
		float vnz = W.vnz_scale*nz/sqrtf(nx*nx + ny*ny + nz*nz);
		if (grass && vnz > sti[1]) {vnz = clip01(1.0f + 20.0f*rand_offset);} // dry patches of dirt and sand in the grass
		if (vnz < sti[1]) { // steep slopes
			if (grass) {
				float rock_weight = (W.tex_class[k1] == TW_TEX_GROUND || W.tex_class[k2] == TW_TEX_ROCK) ? t : 0.0f;
				float const steepness = (float)(1.0 - (double)clip01((vnz - 0.5f*sti[0])*steep_mult_rock));
				rock_weight  = (float)((double)rock_weight*(1.0 - (double)steepness) + (double)steepness);
				weight_scale = clip01((vnz - sti[0])*steep_mult_grass);
				weights[rock_ix] = (float)((double)weights[rock_ix] + (1.0 - (double)weight_scale)*(double)rock_weight);
				weights[dirt_ix] = (float)((double)weights[dirt_ix] + (1.0 - (double)weight_scale)*(1.0 - (double)rock_weight));
			}
			else { // snow
				weight_scale = clip01(2.0f*(vnz - sti[0])*steep_mult_snow);
				weights[rock_ix] = (float)((double)weights[rock_ix] + (1.0 - (double)weight_scale));
			}
		}
	}
	weights[k2] += weight_scale*t;
	weights[k1] = (float)((double)weights[k1] + (double)weight_scale*(1.0 - (double)t));
	float const xv = (float)x*W.xy_mult, yv = (float)y*W.xy_mult;
	if (W.vegetation > 0.0f) { // convert dirt to sand only when there is vegetation
		float const dirt_scale = bilinear(tile_params + 4, xv, yv);
		if (dirt_scale < 1.0f) {
			weights[sand_ix] = (float)((double)weights[sand_ix] + (1.0 - (double)dirt_scale)*(double)weights[dirt_ix]);
			weights[dirt_ix] *= dirt_scale;
		}
	}
	if (grass) {
		float const grass_scale = (mhmin < W.water_level) ? 0.0f : bilinear(tile_params, xv, yv); // no grass under water
		if (grass_scale < 1.0f) { // convert grass to sand
			float const gscale = clip01(2.5f*(grass_scale - 0.5f) + 0.5f);
			weights[sand_ix]  = (float)((double)weights[sand_ix] + (1.0 - (double)gscale)*(double)weights[grass_ix]);
			weights[grass_ix] *= gscale;
		}
	}
	for (int i = 0; i < 4; ++i) { // weights sum to 1: the fifth (snow) is implied
		rgba[i] = ((double)weights[i] <= 0.01) ? (unsigned char)0 : (((double)weights[i] >= 0.99) ? (unsigned char)255 : (unsigned char)(255.0*(double)weights[i]));
	}
	return grass;
}

} // namespace tww
