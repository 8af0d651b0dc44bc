// Drives the C++ adapter (3dworld_b200/host/tw3d_adapter.h) exactly the way the reference's callers drive the reference classes:
//   heightmap_t::proc_gen       (src/heightmap.cpp:130-151): build_arrays(cache_values=1) + enable_glaciate + eval_index loop + apply_erosion
//   tile_t::create_zvals        (src/tiled_mesh.cpp:467-515): build_arrays(no_wait) / enable_glaciate / eval_index, per-tile erosion
//   voxel_manager::create_procedural (src/voxels.cpp:278)
// and dumps the results as raw fp32 for tests/test_cpp_adapter.py to compare with the oracle.
// usage: test_adapter <mode> <out.bin>      (mode = mesh_gen_mode 0..4);  "test_adapter probe" only checks that the library loads
#define TW3D_NO_ABORT
#include "tw3d_adapter.h"
#include <cstdio>
#include <cstdlib>

static void dump(FILE *f, std::vector<float> const &v) {fwrite(v.data(), sizeof(float), v.size(), f);}

int main(int argc, char **argv) {
	if (argc >= 2 && std::string(argv[1]) == "probe") {
		printf("abi %d\n", tw_abi_version());
		try {tw3d::ctx(); printf("device ok\n");}
		catch (tw3d::error const &e) {printf("no device: status %d\n", e.status); return (e.status == TW_ERR_NO_DEVICE) ? 0 : 2;}
		return 0;
	}
	if (argc >= 4 && std::string(argv[1]) == "shadows") { // test_adapter shadows <in.bin> <out.bin>: tests/golden/shadows.npz through tw3d::calc_mesh_shadows
		// in: int32 ntiles, zvsize, nlights; float32 params[7] (X/Y_SCENE_SIZE, DX/DY_VAL, XY_SUM_SIZE, zmin, zmax); int32 tile_xy[2*ntiles]; float32 lights[3*nlights]; float32 tiles
		FILE *in = fopen(argv[2], "rb"), *out = fopen(argv[3], "wb");
		if (!in || !out) return 1;
		int32_t hdr[3]; float prm[7];
		if (fread(hdr, 4, 3, in) != 3 || fread(prm, 4, 7, in) != 7) return 1;
		size_t const nt = hdr[0], zv = hdr[1], nl = hdr[2];
		std::vector<int32_t> txy(2*nt); std::vector<float> lights(3*nl), tiles(nt*zv*zv), ox(nt*zv), oy(nt*zv);
		std::vector<unsigned char> smask(nt*zv*zv);
		if (fread(txy.data(), 4, txy.size(), in) != txy.size() || fread(lights.data(), 4, lights.size(), in) != lights.size() || fread(tiles.data(), 4, tiles.size(), in) != tiles.size()) return 1;
		try {
			tw3d::scene_globals g;
			g.X_SCENE_SIZE = prm[0]; g.Y_SCENE_SIZE = prm[1]; g.MESH_X_SIZE = g.MESH_Y_SIZE = (int)prm[4]/2; g.zmin = prm[5]; g.zmax = prm[6];
			tw3d::set_globals(g);
			for (size_t l = 0; l < nl; ++l) {
				tw3d::calc_mesh_shadows(&lights[3*l], tiles.data(), txy.data(), (unsigned)nt, (unsigned)zv, prm[2], prm[3], smask.data(), ox.data(), oy.data());
				fwrite(smask.data(), 1, smask.size(), out); fwrite(ox.data(), 4, ox.size(), out); fwrite(oy.data(), 4, oy.size(), out);
			}
		}
		catch (tw3d::error const &e) {fprintf(stderr, "error %d: %s\n", e.status, e.what()); return 3;}
		fclose(in); fclose(out);
		return 0;
	}
	if (argc < 3) {fprintf(stderr, "usage: test_adapter <mode> <out.bin>\n"); return 1;}
	int const mode = atoi(argv[1]);
	FILE *f = fopen(argv[2], "wb");
	if (!f) return 1;
	try {
		// scene: mesh 128, scene 4 => DX_VAL 0.0625; mesh_freq_filter 1 => start_eval_sin 10; seed 1; hmap sine as scene_config/config.txt:76
		tw3d::scene_globals g;
		g.mesh_gen_mode = mode; g.mesh_seed = 1; g.start_eval_sin = tw_compute_scale(1.0f, 1); g.zmax_est = 2.3f;
		g.hmap_params.sine_mag = 5.0f; g.hmap_params.sine_freq = 0.001f; g.hmap_params.sine_bias = -4.0f;
		std::vector<float> sinTable(450);
		tw_rng rng = {1, 1};
		tw_gen_sine_params(&rng, g.MESH_HEIGHT*g.mesh_height_scale, 128, 128, 4.0f, 4.0f, g.mesh_seed, g.mesh_rgen_index, mode, 0.02f, 240.0f, 2.0f, 0.5f, sinTable.data());
		tw3d::set_globals(g, nullptr, sinTable.data());
		float const DX = 0.0625f, DY = 0.0625f;

		// --- heightmap_t::proc_gen ---
		int const W = 160, H = 96;
		std::vector<float> vals((size_t)W*H);
		{
			tw3d::mesh_xy_grid_cache_t height_gen;
			height_gen.build_arrays(-0.5*W, -0.5*H, DX, DY, W, H, 1); // cache_values=1
			height_gen.enable_glaciate();
			for (int i = 0; i < H; ++i) {for (int j = 0; j < W; ++j) {vals[(size_t)W*i + j] = height_gen.eval_index(j, i);}}
		}
		dump(f, vals);
		float min_z = vals[0], max_z = vals[0];
		for (float v : vals) {min_z = (v < min_z) ? v : min_z; max_z = (v > max_z) ? v : max_z;}
		g.zmin = min_z - 0.1f; g.zmax = max_z + 0.1f; g.water_plane_z = min_z - 10.0f; g.clip_hd1 = 0.5f;
		tw3d::set_globals(g);
		tw3d::apply_erosion(vals.data(), W, H, min_z, 700); // run_erosion
		dump(f, vals);

		// --- tile_t::create_zvals with the async (no_wait) protocol of the GPU gen modes ---
		unsigned const zvsize = 66;
		int const x1 = 5*64, y1 = -3*64;
		std::vector<float> zvals((size_t)zvsize*zvsize);
		{
			tw3d::mesh_xy_grid_cache_t height_gen;
			int frames = 0;
			for (;;) { // setup_height_gen_async every frame until the results are ready (src/tiled_mesh.cpp:2393-2402)
				bool const ready = height_gen.build_arrays((x1 - 128/2), (y1 - 128/2), DX, DY, zvsize, zvsize, 0, 0, /*no_wait=*/1);
				height_gen.enable_glaciate();
				++frames;
				if (ready) break;
				if (frames > 100000000) {fprintf(stderr, "async job never finished\n"); return 3;}
			}
			for (unsigned y = 0; y < zvsize; ++y) {for (unsigned x = 0; x < zvsize; ++x) {zvals[y*zvsize + x] = height_gen.eval_index(x, y);}}
			printf("tile ready after %d frame(s)\n", frames);
		}
		dump(f, zvals);

		// --- batched tiles: height + 300 droplets each ---
		int32_t const origins[6] = {0, 0, 64, 0, 0, 64};
		std::vector<float> batch((size_t)3*zvsize*zvsize);
		g.MESH_X_SIZE = g.MESH_Y_SIZE = 128;
		tw3d::set_globals(g);
		tw3d::create_zvals_batch(origins, 3, zvsize, DX, DY, 300, batch.data());
		dump(f, batch);

		// --- voxel_manager::create_procedural ---
		std::vector<float> vox;
		tw3d::voxel_grid_view v = {24, 10, 30, {0.4f, 0.65f, 0.11f}, {-7.9f, -7.8f, -1.5f}, &vox};
		float const offset[3] = {0.5f, -0.25f, 0.0f};
		tw3d::create_procedural(v, 1.0f, 1.0f, offset, true, 123, 456, (mode >= 3 ? 1 : mode), 0.0f, 2);
		dump(f, vox);

		// --- point queries: get_exact_zval for a scrolled scene (batch and single-point forms), eval_mesh_sin_terms ---
		{
			g.xoff2 = 100; g.yoff2 = -40;
			tw3d::set_globals(g);
			float const xy[10] = {0.0f, 0.0f, 1.5f, -2.25f, -3.9f, 3.9f, 100.0f, 250.0f, 0.03125f, 0.0625f};
			std::vector<float> pz(5);
			tw3d::get_exact_zvals(xy, 5, pz.data());
			pz.push_back(tw3d::get_exact_zval(1.5f, -2.25f, true));
			pz.push_back(tw3d::eval_mesh_sin_terms(0.3f, -7.0f));
			pz.push_back(tw3d::eval_mesh_sin_terms_scaled(70.0f, 12.5f, 16.0f));
			dump(f, pz);
		}

		// --- gen_mesh(): BASELINE config 1 (128x128 sine mesh, mesh_seed 6, glaciate, mesh_freq_filter 2, mesh_height 0.7) + 2000 droplets ---
		if (mode == 0) {
			tw3d::scene_globals g1;
			g1.mesh_gen_mode = 0; g1.mesh_seed = 6; g1.start_eval_sin = tw_compute_scale(1.0f, 2); g1.mesh_height_scale = 0.7f;
			g1.hmap_params.sine_mag = 5.0f; g1.hmap_params.sine_freq = 0.001f; g1.hmap_params.sine_bias = -4.0f;
			g1.clip_hd1 = 0.5f; g1.relh_adj_tex = 0.0f;
			tw3d::set_globals(g1);
			tw_rng srng = {1, 1};
			std::vector<float> mesh(128*128);
			tw3d::gen_mesh_result const r = tw3d::gen_mesh(mesh.data(), srng, 4.0f, 4.0f, 2000);
			dump(f, mesh);
			std::vector<float> z6 = {r.zmin, r.zmax, r.zmax_est, r.zbottom, r.ztop, r.water_plane_z};
			dump(f, z6);
		}
	}
	catch (tw3d::error const &e) {fprintf(stderr, "tw3d error %d: %s\n", e.status, e.what()); fclose(f); return 2;}
	fclose(f);
	return 0;
}
