"""Fused tile pipeline (tw_create_zvals_batch) vs separate calls; profiling aid. Env: TW_PIPE_CHUNKS, TW_PIPE_NOPRIO."""
import argparse, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
tw = importlib.import_module("3dworld_b200"); scene = importlib.import_module("3dworld_b200.scene")
ap = argparse.ArgumentParser(); ap.add_argument("--tiles", type=int, default=16384); ap.add_argument("--mode", type=int, default=4); ap.add_argument("--iters", type=int, default=1000)
args = ap.parse_args()
cfg = scene.SceneConfig(mesh_gen_mode=args.mode, mesh_freq_filter=1, mesh_seed=1, hmap=dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0), zmax_est=2.3, mesh_size=(256, 256, 1))
hp, ep = cfg.height_params(), cfg.erosion_params()
ctx = tw.Context(0)
nt, zv = args.tiles, 258
side = int(nt ** 0.5 + 0.999)
origins = [((t % side) * 256, (t // side) * 256) for t in range(nt)]
tiles = torch.empty((nt, zv, zv), dtype=torch.float32, device="cuda")
dx, dy = float(cfg.dx_val), float(cfg.dy_val)
ctx.create_zvals_batch(origins[:4096], cfg.mesh_size, dx, dy, zv, hp, 50, ep, ep.zmin, out=tiles[:4096])
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp, out=tiles); t1 = time.perf_counter()
    ctx.erode_tiles(tiles, args.iters, ep, min_zval_all=ep.zmin); t2 = time.perf_counter()
    ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, args.iters, ep, ep.zmin, out=tiles); t3 = time.perf_counter()
print("chunks=%s noprio=%s tiles %d mode %d: gen %.3f s  erode %.3f s (%.1f moves/droplet)  separate %.3f s  fused %.3f s" %
      (os.environ.get("TW_PIPE_CHUNKS", "4"), os.environ.get("TW_PIPE_NOPRIO", "0"), nt, args.mode, t1 - t0, t2 - t1, ctx.last_erosion_steps / (nt * float(args.iters)), t2 - t0, t3 - t2), flush=True)
ctx.close()
