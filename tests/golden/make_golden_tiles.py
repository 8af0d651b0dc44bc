"""Generates tests/golden/tiles.npz. Build container only; fixtures committed.
  normals_*  : the UNMODIFIED reference's vector3d::get_norm()/byte quantisation over a reference-generated tile (oracle/_ref ref_tile_normals)
  ao_*       : the reference's OWN tile_t::calc_mesh_ao_lighting (function body cut out of src/tiled_mesh.cpp at build time, oracle/refbuild/build_ref.sh)
               on reference-generated heights. m0/m1: the CPU-gen-mode flow (it builds the context itself, zvals inside the tile). m4: the GPU-gen-mode
               flow of tile_t::create_zvals + calc_mesh_ao_lighting with enable_tiled_mesh_ao: context generated once, tile_m4c = its interior,
               tile_m4e = apply_erosion(tile_m4c) (ero_m4 = min_zval, iters, erosion params), ao_m4c / ao_m4e = AO against the UN-eroded context
    python tests/golden/make_golden_tiles.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
import refapi as R  # noqa: E402
from cases import HM_CFG  # noqa: E402

RL = R.lib()
d = {}
S, zv, ray = 64, 66, 36
for mode in (0, 1, 4):
    R.setup(mode=mode, freq_filter=1, seed=1, zmax_est=2.3, hmap=HM_CFG, mesh=(64, 64, 1))
    dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
    x1, y1 = 3 * S, -5 * S
    tile = R.heightgen(x1 - 32, y1 - 32, dx, dy, zv, zv, 0, 1)                      # setup_height_gen_async(x1, y1, zvsize, zvsize): x0 = x1 - MESH_X_SIZE/2
    csz = zv - 1 + 2 * ray
    context = R.heightgen(x1 - ray - 32, y1 - ray - 32, dx, dy, csz, csz, 0, 1)
    n = "m%d" % mode
    d["args_" + n] = np.array([mode, x1, y1, S, zv, dx, dy, RL.ref_get_half_dxy()], np.float64)
    if mode == 0:
        d["sp_" + n] = R.sine_params()
    d["tile_" + n] = tile
    rgba, mnz = R.tile_normals(tile, dx, dy)
    d["normals_" + n], d["min_normal_z_" + n] = rgba, np.float32(mnz)
    hd = RL.ref_get_half_dxy()
    if mode < 3:
        d["ao_" + n] = R.tile_ao_lighting(S, x1, y1, tile, None, hd)
        assert np.array_equal(d["ao_" + n], O.tile_ao(tile[None], context[None], hd)[0])
    else:
        d["ao_" + n] = R.tile_ao_lighting(S, x1, y1, tile, context, hd)          # given zvals are only the ray origins
        cut = np.ascontiguousarray(context[ray:ray + zv, ray:ray + zv])
        lo, hi = float(cut.min()), float(cut.max())
        ero = [lo, 500, 1.0, lo + 0.25 * (hi - lo), hd, lo - 0.1, hi + 0.1, 0.0, 0.5]     # min_zval, iters, erode_amount, water_plane_z, half_dxy, zmin, zmax, relh_adj_tex, clip_hd1
        RL.ref_set_threads(1)
        eroded = R.apply_erosion(cut, ero[0], int(ero[1]), *ero[2:])
        RL.ref_set_threads(8)
        assert (eroded != cut).sum() > 100
        d["tile_" + n + "c"], d["tile_" + n + "e"], d["ero_" + n] = cut, eroded, np.array(ero, np.float64)
        d["ao_" + n + "c"] = R.tile_ao_lighting(S, x1, y1, cut, context, hd)
        d["ao_" + n + "e"] = R.tile_ao_lighting(S, x1, y1, eroded, context, hd)
        assert (d["ao_" + n + "e"] != O.tile_ao(eroded[None], context[None], hd)[0]).any()   # the CPU-mode flow would give a different map
# heightmap-texture tiles (N2): the reference's terrain_hmap_manager_t::get_clamped_height over a 16-bit image (mirror edges)
rng = np.random.default_rng(77)
R.setup(mode=0, freq_filter=1, seed=1, mesh_height_scale=1.5, mesh_scale_z=2.0)
img = rng.integers(0, 256, (40, 56, 2), dtype=np.uint8)
d["hmap_img"] = img
cases = [(1.0, 1.0, 0.0, -20, -15), (0.5, 2.5, -0.75, -60, 30), (2.0, 0.8, 0.0, 100, -90), (0.37, 1.0, 0.1, -1000, 999)]
d["hmap_cases"] = np.array(cases, np.float64)
d["hmap_out"] = np.stack([R.hmap_sample_tile(img, int(x1), int(y1), 34, ms, mfs, tz) for ms, mfs, tz, x1, y1 in cases])
np.savez_compressed(os.path.join(HERE, "tiles.npz"), **d)
print("wrote tiles.npz", {k: v.shape for k, v in d.items() if k.startswith(("ao", "normals"))}, [float(d["ao_m%d" % m].mean()) for m in (0, 1, 4)])
