"""Erosion mode sweep (profiling aid): single map and tile batches through the global / window / whole modes of droplet_kernel.
Settings are passed as environment variables that the library re-reads per call."""
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
KEYS = ("TW_EROSION_MODE", "TW_EROSION_LANES", "TW_EROSION_SMEM_LANES", "TW_EROSION_WIN", "TW_EROSION_WIN_MIN_MOVES", "TW_EROSION_WHOLE_MAX",
        "TW_EROSION_WINDOW_ALL", "TW_EROSION_HEAVY", "TW_PIPE_CHUNKS")


def setenv(env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)


def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def single_map(ctx, N, iters_list, envs):
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
    d = torch.empty((N, N), dtype=torch.float32, device="cuda")
    _, (zmin, zmax) = ctx.heightgen_2d(cfg.heightmap_grid(N, N), cfg.height_params(), out=d, want_minmax=True)
    ep = cfg.erosion_params()
    work = torch.empty_like(d)
    for iters in iters_list:
        for name, env in envs:
            setenv(env)

            def run():
                work.copy_(d)
                torch.cuda.synchronize()
                run.t0 = time.perf_counter()
                ctx.erode(work, zmin, iters, ep)
                torch.cuda.synchronize()
                run.dt = time.perf_counter() - run.t0
            best = 1e9
            for _ in range(2):
                run()
                best = min(best, run.dt)
            mv = ctx.last_erosion_steps
            print("map %5d^2 droplets %7d %-28s %.4f s  %.3e droplets/s  %.3e moves/s  %.3f us/move (%.1f moves/droplet)" %
                  (N, iters, name, best, iters / best, mv / best, 1e6 * best / mv, mv / iters), flush=True)


def tiles(ctx, S, nt, iters, envs, fused=False):
    zv = S + 2
    cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    hp, ep = cfg.height_params(), cfg.erosion_params()
    side = int(nt ** 0.5 + 0.999)
    origins = [((t % side) * S, (t // side) * S) for t in range(nt)]
    base = torch.empty((nt, zv, zv), dtype=torch.float32, device="cuda")
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    tg = timed(lambda: ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp, out=base))
    work = torch.empty_like(base)
    for name, env in envs:
        setenv(env)
        best = 1e9
        for _ in range(2):
            work.copy_(base)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.erode_tiles(work, iters, ep, min_zval_all=ep.zmin)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        mv = ctx.last_erosion_steps
        line = "tiles %6d x %d^2 droplets %d %-34s erode %.4f s  %.3e droplets/s  %.3e moves/s (%.1f moves/droplet; gen %.4f s)" % (
            nt, zv, iters, name, best, nt * iters / best, mv / best, mv / (nt * iters), tg)
        if fused:
            tf = timed(lambda: ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, iters, ep, ep.zmin, out=work))
            line += "  fused %.4f s" % tf
        print(line, flush=True)
    del base, work


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", nargs="+", default=["map", "t130", "t258"])
    ap.add_argument("--big", action="store_true")
    args = ap.parse_args()
    ctx = tw.Context(0)
    G = dict(TW_EROSION_MODE="global", TW_EROSION_LANES="32")
    if "map" in args.what:
        envs = [("global G32", G)]
        for w in (16, 32, 48, 64):
            for k in (0, 2):
                envs.append(("window W%d K%d" % (w, k), dict(TW_EROSION_MODE="window", TW_EROSION_WIN=str(w), TW_EROSION_WIN_MIN_MOVES=str(k))))
        envs.append(("window W32 K0 G16", dict(TW_EROSION_MODE="window", TW_EROSION_WIN="32", TW_EROSION_WIN_MIN_MOVES="0", TW_EROSION_SMEM_LANES="16")))
        envs.append(("auto", {}))
        single_map(ctx, 8192, [1000, 20000], envs)
        single_map(ctx, 128, [20000], [("global G32", G), ("window W32 K0", dict(TW_EROSION_MODE="window", TW_EROSION_WIN_MIN_MOVES="0")),
                                       ("whole G32", dict(TW_EROSION_MODE="whole")), ("whole G16", dict(TW_EROSION_MODE="whole", TW_EROSION_SMEM_LANES="16")), ("auto", {})])
    if "t130" in args.what:
        for nt in (16, 444, 1776, 8192) + ((32768,) if args.big else ()):
            envs = [("global auto-G", dict(TW_EROSION_MODE="global")), ("whole", dict(TW_EROSION_MODE="whole")),
                    ("window W32 K2", dict(TW_EROSION_MODE="window")), ("auto", {})]
            tiles(ctx, 128, nt, 1000, envs)
    if "t258" in args.what:
        for nt in (16, 1024, 8192) + ((16384, 65536) if args.big else ()):
            envs = [("global auto-G", dict(TW_EROSION_MODE="global")), ("window W32 K2", dict(TW_EROSION_MODE="window")),
                    ("window W48 K2", dict(TW_EROSION_MODE="window", TW_EROSION_WIN="48")),
                    ("split heavy 592", dict(TW_EROSION_WINDOW_ALL="0", TW_EROSION_HEAVY="592")),
                    ("split heavy 1184", dict(TW_EROSION_WINDOW_ALL="0", TW_EROSION_HEAVY="1184")),
                    ("split heavy 2368", dict(TW_EROSION_WINDOW_ALL="0", TW_EROSION_HEAVY="2368")), ("auto", {})]
            if nt <= 1024:
                envs = [e for e in envs if not e[0].startswith("split")]
            tiles(ctx, 256, nt, 1000, envs, fused=(nt >= 8192))


if __name__ == "__main__":
    main()
