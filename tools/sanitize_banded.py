"""compute-sanitizer target: the band decomposition of the sweep erosion with middle bands, all on one device (small case).
    compute-sanitizer --tool memcheck python tools/sanitize_banded.py [nbands] [sharded|banded] [nx ny iters sweep halo]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
tw = importlib.import_module("3dworld_b200")
scene = importlib.import_module("3dworld_b200.scene")
from cases import HM_CFG  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sharded = len(sys.argv) > 2 and sys.argv[2] == "sharded"   # one band per GPU through tw_erode_sweeps_sharded (NCCL) instead of all bands on device 0
cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0)
ctx = tw.Context(0)
ep = cfg.erosion_params()
nx, ny, iters, sweep, halo = 200, 100 * nb, 4000, 512, 44
if len(sys.argv) > 7:
    nx, ny, iters, sweep, halo = (int(v) for v in sys.argv[3:8])
z = ctx.heightgen_2d(cfg.heightmap_grid(nx, ny), cfg.height_params())
zmin = float(z.min())
one = z.copy()
m1 = ctx.erode_sweeps(one, zmin, iters, ep, sweep, halo)
bands = [z[a:b].copy() for a, b in (tw.multi_range(ny, nb, i) for i in range(nb))]
if sharded:
    M = tw.Multi(list(range(nb)))
    m2 = M.erode_sweeps_sharded(bands, nx, ny, zmin, iters, ep, sweep, halo)
    M.close()
else:
    m2 = ctx.erode_sweeps_banded(bands, nx, ny, zmin, iters, ep, sweep, halo)
got = np.concatenate(bands)
print("moves", m1, m2, "differing cells", int((got.view(np.uint32) != one.view(np.uint32)).sum()))
