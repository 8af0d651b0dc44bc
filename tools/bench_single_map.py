"""Single-heightmap erosion (BASELINE config 3 shape): serial order (tw_erode) vs the reference's OpenMP mode (tw_erode_parallel). Profiling aid."""
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
    ap.add_argument("--size", type=int, nargs="+", default=[1024, 8192])
    ap.add_argument("--iters", type=int, nargs="+", default=[10000, 100000])
    ap.add_argument("--direct-max", type=int, default=100000)
    args = ap.parse_args()
    ctx = tw.Context(0)
    for N in args.size:
        cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
        d = torch.empty((N, N), dtype=torch.float32, device="cuda")
        _, (zmin, zmax) = ctx.heightgen_2d(cfg.heightmap_grid(N, N), cfg.height_params(), out=d, want_minmax=True)
        ep = cfg.erosion_params()
        for iters in args.iters:
            for mode in ("serial", "omp-1", "omp-64", "omp-4096", "omp-auto"):
                if mode in ("serial", "omp-1") and iters > args.direct_max:
                    continue
                for rep in range(2):
                    work = d.clone()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    if mode == "serial":
                        ctx.erode(work, zmin, iters, ep)
                    else:
                        ctx.erode_parallel(work, zmin, iters, ep, 0 if mode == "omp-auto" else int(mode[4:]))
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                print("map %d^2 droplets %7d %-8s: %.4f s  %.3e droplets/s  %.3e moves/s (%.1f moves/droplet)" %
                      (N, iters, mode, dt, iters / dt, ctx.last_erosion_steps / dt, ctx.last_erosion_steps / iters), flush=True)


if __name__ == "__main__":
    main()
