// Multi-GPU through the C++ adapter (tw3d::multi_gpu): the tile loop of tile_draw_t::update dealt out over all visible devices must give exactly what one
// device gives for the whole batch, and the NCCL-reduced z range must equal the min/max over all tiles. usage: test_multi [ndev]   (0 = all devices)
#define TW3D_NO_ABORT
#include "tw3d_adapter.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime_api.h>

int main(int argc, char **argv) {
	try {
		int ndev = (argc > 1) ? atoi(argv[1]) : 0, have = 0;
		if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {printf("no device\n"); return 0;}
		if (ndev <= 0 || ndev > have) ndev = have;
		tw3d::scene_globals g;
		g.mesh_gen_mode = 4; g.mesh_seed = 1; g.start_eval_sin = tw_compute_scale(1.0f, 1); g.zmax_est = 2.3f;
		g.MESH_X_SIZE = g.MESH_Y_SIZE = 64; g.DX_VAL_INV = g.DY_VAL_INV = 8.0f; g.HALF_DXY = 0.125f;
		g.hmap_params.sine_mag = 5.0f; g.hmap_params.sine_freq = 0.001f; g.hmap_params.sine_bias = -4.0f;
		g.water_plane_z = -1.0f; g.zmin = -9.0f; g.zmax = 6.0f; g.clip_hd1 = 0.4f;
		tw3d::set_globals(g);
		unsigned const S = 64, zv = 66, side = 9, nt = side*side, iters = 200;
		float const dx = 0.125f, dy = 0.125f;
		std::vector<int32_t> org(2*nt);
		for (unsigned t = 0; t < nt; ++t) {org[2*t] = (int)(t % side)*(int)S*5 - 700; org[2*t+1] = (int)(t / side)*(int)S*3 + 90;}
		std::vector<float> one((size_t)nt*zv*zv);
		std::vector<tw_minmax> mm1(nt), mm(nt);
		tw3d::create_zvals_batch(org.data(), nt, zv, dx, dy, iters, one.data(), mm1.data());
		tw3d::multi_gpu M(ndev);
		std::vector<float *> bands(ndev);
		for (int i = 0; i < ndev; ++i) {bands[i] = M.alloc_band(i, nt, zv);}
		tw_minmax const zr = M.create_zvals(org.data(), nt, zv, dx, dy, iters, bands.data(), mm.data());
		size_t bad = 0;
		float lo = one[0], hi = one[0];
		for (float v : one) {lo = (v < lo) ? v : lo; hi = (v > hi) ? v : hi;}
		for (int i = 0; i < ndev; ++i) {
			uint32_t a, b;
			tw_multi_range(nt, ndev, i, &a, &b);
			bad += (memcmp(bands[i], one.data() + (size_t)a*zv*zv, (size_t)(b - a)*zv*zv*sizeof(float)) != 0);
		}
		bad += (memcmp(mm.data(), mm1.data(), nt*sizeof(tw_minmax)) != 0);
		bad += !(zr.zmin == lo && zr.zmax == hi);
		printf("devices %d tiles %u: %s (z range %.6f .. %.6f)\n", ndev, nt, bad ? "MISMATCH" : "identical to one device", zr.zmin, zr.zmax);
		return bad ? 3 : 0;
	}
	catch (tw3d::error const &e) {fprintf(stderr, "tw3d error %d: %s\n", e.status, e.what()); return 2;}
}
