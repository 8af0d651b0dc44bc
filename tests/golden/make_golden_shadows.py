"""Generates tests/golden/shadows.npz. Build container only; fixture committed.
The reference's OWN calc_mesh_shadows / mesh_shadow_gen / do_line_clip (function bodies cut out of src/visibility.cpp and src/Math3d.cpp at build time,
oracle/refbuild/build_ref.sh; serial run_x-then-run_y semantics) over a 3x3 block of reference-generated tiles, chained by hand the way
tile_t::calc_shadows_for_light does (src/tiled_mesh.cpp:664-692): tiles toward the light first, sh_in = the neighbours' sh_out.
    python tests/golden/make_golden_shadows.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refapi as R  # noqa: E402
from cases import HM_CFG  # noqa: E402

RL = R.lib()
assert R.has_shadow_extract()
S, zv = 32, 34
R.setup(mesh=(S, S, 1), mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
txy = [(tx, ty) for ty in range(-1, 2) for tx in range(2, 5)]
tiles = np.stack([R.heightgen(tx * S - S // 2, ty * S - S // 2, dx, dy, zv, zv, 0, 1) for tx, ty in txy])
tiles = ((tiles - np.float32(tiles.mean())) * np.float32(3.0)).astype(np.float32)
zlo, zhi = float(tiles.min()) - 0.5, float(tiles.max()) + 0.5
lights = [(3.0, 2.0, 0.4), (-4.0, 1.0, 0.3), (1.0, -5.0, 0.5), (-2.0, -3.0, 0.6), (0.0, 0.0, 5.0), (2.0, 1.0, zlo - 1.0)]
d = {"tiles": tiles, "tile_xy": np.array(txy, np.int32), "lights": np.array(lights, np.float32),
     "params": np.array([4.0, 4.0, dx, dy, 2 * S, zlo, zhi], np.float64)}       # X/Y_SCENE_SIZE, DX/DY_VAL, XY_SUM_SIZE, zmin, zmax
idx = {t: i for i, t in enumerate(txy)}
for li, lp in enumerate(lights):
    sx, sy = (-1 if lp[0] < 0 else 1), (-1 if lp[1] < 0 else 1)
    done, masks = {}, {}
    for i in sorted(range(len(txy)), key=lambda i: -(sx * txy[i][0] + sy * txy[i][1])):
        nbx, nby = idx.get((txy[i][0] + sx, txy[i][1])), idx.get((txy[i][0], txy[i][1] + sy))
        six = done[nby][0] if nby is not None else None
        siy = done[nbx][1] if nbx is not None else None
        m, ox, oy = R.calc_mesh_shadows(lp, tiles[i], zlo, zhi, six, siy)
        done[i], masks[i] = (ox, oy), m
    d["smask_%d" % li] = np.stack([masks[i] for i in range(len(txy))])
    d["sh_out_x_%d" % li] = np.stack([done[i][0] for i in range(len(txy))])
    d["sh_out_y_%d" % li] = np.stack([done[i][1] for i in range(len(txy))])
np.savez_compressed(os.path.join(HERE, "shadows.npz"), **d)
print("wrote shadows.npz", [float((d["smask_%d" % i] == 2).mean()) for i in range(len(lights))])
