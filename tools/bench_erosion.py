"""Erosion throughput sweep (tiles x size): droplets/s and steps/s of tw_erode_tiles with the tiles resident in HBM. Profiling aid."""
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
HM_CFG = dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, nargs="+", default=[2048, 8192, 16384])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    ctx = tw.Context(0)
    S, zv = args.size, args.size + 2
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    hp = cfg.height_params()
    ep = cfg.erosion_params()
    for nt in args.tiles:
        side = int(nt ** 0.5 + 0.999)
        origins = [((t % side) * S, (t // side) * S) for t in range(nt)]
        tiles = torch.empty((nt, zv, zv), dtype=torch.float32, device="cuda")
        ctx.heightgen_tiles(origins, cfg.mesh_size, float(cfg.dx_val), float(cfg.dy_val), zv, hp, out=tiles)
        zmin, zmax = ctx.minmax(tiles)
        for rep in range(args.reps + 1):
            work = tiles.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.erode_tiles(work, args.iters, ep, min_zval_all=zmin)
            dt = time.perf_counter() - t0
            if rep > 0 or args.reps == 0:
                print("tiles %6d size %d iters %d: %.4f s  %.3e droplets/s  %.3e steps/s  (%.1f steps/droplet)" %
                      (nt, zv, args.iters, dt, nt * args.iters / dt, ctx.last_erosion_steps / dt, ctx.last_erosion_steps / (nt * args.iters)), flush=True)
        del tiles, work


if __name__ == "__main__":
    main()
