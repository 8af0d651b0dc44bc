"""Generates tests/golden/points.npz from the UNMODIFIED reference (oracle/_ref/libref3dworld.so): eval_mesh_sin_terms,
eval_mesh_sin_terms_scaled and get_exact_zval (src/mesh_gen.cpp:797-847) at random points - SURVEY.md section 8(a) row a9.
Same conventions as make_golden.py (build container only; fixtures committed).
    python tests/golden/make_golden_points.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refapi as R  # noqa: E402
from cases import HM_ALL, HM_CFG  # noqa: E402

rng = np.random.default_rng(4321)
d = {}
for mode in (0, 1, 2, 3, 4):
    for shape, ff, hmap, gl in ((0, 1, HM_CFG, 1), (1, 2, HM_ALL, 1), (2, 0, {}, 0)):
        R.setup(mode=mode, shape=shape, freq_filter=ff, seed=1, glaciate=gl, hmap=hmap, zmax_est=2.3)
        n = 96 if mode == 4 else 256
        name = "p_m%d_s%d" % (mode, shape)
        d[name + "_args"] = np.array([mode, shape, ff, gl], np.float64)
        d[name + "_hmap"] = np.array([hmap.get(k, dflt) for k, dflt in zip(R.HMAP_FIELDS, R.HMAP_DEFAULT)], np.float32)
        if mode == 0 or shape == 0:
            d[name + "_sp"] = R.sine_params()
        # (kind, span, xy_scale, no_xyoff, xoff2, yoff2)
        for qi, (kind, span, xy_scale, no_xyoff, xo, yo) in enumerate(((0, 30.0, 1.0, 0, 0, 0), (1, 500.0, 16.0, 0, 0, 0), (2, 4.0, 1.0, 0, 640, -1280), (2, 40.0, 1.0, 1, 5, 5))):
            if kind == 0 and not (mode == 0 or shape == 0):
                continue
            xy = rng.uniform(-span, span, (n, 2)).astype(np.float32)
            xy[: n // 8] = np.round(xy[: n // 8])
            d["%s_q%d_xy" % (name, qi)] = xy
            d["%s_q%d_query" % (name, qi)] = np.array([kind, xy_scale, 128, 128, 4.0, 4.0, xo, yo, no_xyoff], np.float64)
            d["%s_q%d_out" % (name, qi)] = R.eval_points(kind, xy, xy_scale, no_xyoff, xo, yo)
np.savez_compressed(os.path.join(HERE, "points.npz"), **d)
print("wrote points.npz:", len(d), "arrays")
