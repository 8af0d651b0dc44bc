"""Smallest driver for ncu captures of the serial droplet walk: one 8192^2 (or --size) simplex map, --iters droplets through tw_erode."""
import argparse
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

tw = importlib.import_module("3dworld_b200")
scene = importlib.import_module("3dworld_b200.scene")
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=8192)
ap.add_argument("--iters", type=int, default=1000)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
ctx = tw.Context(0)
cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0), zmax_est=2.3)
d = torch.empty((a.size, a.size), dtype=torch.float32, device="cuda")
_, (zmin, zmax) = ctx.heightgen_2d(cfg.heightmap_grid(a.size, a.size), cfg.height_params(), out=d, want_minmax=True)
ep = cfg.erosion_params()
for _ in range(a.reps):
    w = d.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.erode(w, zmin, a.iters, ep)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print("map %d^2, %d droplets: %.4f s, %.3f us/move, %d moves" % (a.size, a.iters, dt, 1e6 * dt / ctx.last_erosion_steps, ctx.last_erosion_steps))
