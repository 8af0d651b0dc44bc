"""Generates tests/golden/tiles.npz. Build container only; fixtures committed.
  normals_*  : the UNMODIFIED reference's vector3d::get_norm()/byte quantisation over a reference-generated tile (oracle/_ref ref_tile_normals)
  ao_*       : tile_t::calc_mesh_ao_lighting restated (oracle to_tile_ao; tiled_mesh.cpp cannot be linked) on heights and context grid that the
               reference's mesh_xy_grid_cache_t produced
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
    d["ao_" + n] = O.tile_ao(tile[None], context[None], RL.ref_get_half_dxy())[0]
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
