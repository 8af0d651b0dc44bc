"""Generates tests/golden/weights.npz (terrain weights texture). Build container only; fixture committed.
The reference's OWN tile_t::create_texture (cut out of src/tiled_mesh.cpp at build time with get_tids / update_lttex_ix from src/Textures.cpp,
oracle/refbuild/build_ref.sh; terrain-only path) on reference-generated tiles, for several parameter sets.
    python tests/golden/make_golden_weights.py"""
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
assert R.has_texture_extract()
RL.ref_set_threads(1)
S, zv = 64, 66
d = {}
ids = R.tex_ids()
for mode in (1, 4):
    R.setup(mesh=(S, S, 1), mode=mode, freq_filter=1, seed=1, zmax_est=2.3, hmap=HM_CFG)
    dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
    origins = [(3 * S, -5 * S), (4 * S, -5 * S), (-2 * S, 7 * S)]
    tiles = np.stack([R.heightgen(x1 - S // 2, y1 - S // 2, dx, dy, zv, zv, 0, 1) for x1, y1 in origins])
    tiles = ((tiles - np.float32(tiles.mean())) * np.float32(0.4 / max(1e-6, float(tiles.std())))).astype(np.float32)
    lo, hi = float(tiles.min()), float(tiles.max())
    n = "m%d" % mode
    d["tiles_" + n], d["origins_" + n] = tiles, np.array(origins, np.int32)
    d["sine_params_" + n] = R.sine_params()
    cases = [dict(h_dirt=[0.40, 0.44, 0.60, 0.75, 1.0], order=[0, 1, 2, 3, 4], relh=0.0, veg=1.0, s2r=0), dict(h_dirt=[0.2, 0.5, 0.62, 0.7, 1.0], order=[1, 0, 2, 4, 3], relh=0.03, veg=1.0, s2r=1),
             dict(h_dirt=[0.40, 0.44, 0.60, 0.75, 1.0], order=[0, 1, 2, 3, 4], relh=-0.05, veg=0.0, s2r=0)]
    rng = np.random.default_rng(3)
    corners = rng.uniform(-0.1, 1.2, (len(origins), 8)).astype(np.float32)         # API layout: grass[4] then dirt[4] per tile
    d["corners_" + n] = corners
    zmin, zmax = lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo)
    d["scal_" + n] = np.array([mode, S, dx, dy, zmin, zmax, RL.ref_get_water_z_height(), RL.ref_get_start_eval_sin(), RL.ref_get_mesh_height()], np.float64)
    for ci, c in enumerate(cases):
        d["case_%s_%d" % (n, ci)] = np.array(c["h_dirt"] + c["order"] + [c["relh"], c["veg"], c["s2r"]], np.float64)
        outs, flags = [], []
        for t, (x1, y1) in enumerate(origins):
            cr = [v for i in range(4) for v in (corners[t][i], corners[t][4 + i])]
            w, hag = R.tile_create_texture(S, x1, y1, tiles[t], cr, c["h_dirt"], [ids[k] for k in c["order"]], c["veg"], c["relh"], zmin, zmax, c["s2r"])
            outs.append(w)
            flags.append(hag)
        d["weights_%s_%d" % (n, ci)], d["grass_%s_%d" % (n, ci)] = np.stack(outs), np.array(flags, np.uint8)
np.savez_compressed(os.path.join(HERE, "weights.npz"), **d)
print("wrote weights.npz", {k: v.shape for k, v in d.items() if k.startswith("weights")})
