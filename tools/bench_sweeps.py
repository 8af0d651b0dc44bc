"""tw_erode_sweeps (coherent batched erosion) throughput on one 8192^2 map; profiling aid."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

tw = importlib.import_module("3dworld_b200")
scene = importlib.import_module("3dworld_b200.scene")
ctx = tw.Context(0)
N = 8192
cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0), zmax_est=2.3)
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_, (zmin, zmax) = ctx.heightgen_2d(cfg.heightmap_grid(N, N), cfg.height_params(), out=d, want_minmax=True)
ep = cfg.erosion_params()
for iters, sweep in ((100000, 1024), (100000, 8192), (1000000, 8192), (1000000, 65536)):
    best = 1e9
    for _ in range(2):
        w = d.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moves = ctx.erode_sweeps(w, zmin, iters, ep, sweep, 64)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("sweeps: 8192^2, %7d droplets, sweep %6d, halo 64: %.4f s  %.3e droplets/s  %.3e moves/s (%.1f moves/droplet)" % (iters, sweep, best, iters / best, moves / best, moves / iters), flush=True)
