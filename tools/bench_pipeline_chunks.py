"""tw_create_zvals_batch: chunk-count / reorder sweep on the BASELINE config-5 shape (258^2 tiles, mode 4, 1000 droplets); profiling aid."""
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
ap.add_argument("--tiles", type=int, nargs="+", default=[8192, 65536])
ap.add_argument("--chunks", type=int, nargs="+", default=[1, 2, 3, 4, 8, 16])
a = ap.parse_args()
ctx = tw.Context(0)
cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, hmap=dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0), zmax_est=2.3, mesh_size=(256, 256, 1))
hp, ep = cfg.height_params(), cfg.erosion_params()
dx, dy, zv = float(cfg.dx_val), float(cfg.dy_val), 258
for nt in a.tiles:
    side = 256
    origins = [((t % side) * 256, (t // side) * 256) for t in range(nt)]
    out = torch.empty((nt, zv, zv), dtype=torch.float32, device="cuda")
    for reorder in (1, 0):
        for ch in a.chunks:
            if ch == 1 and not reorder:
                continue
            os.environ["TW_PIPE_CHUNKS"] = str(ch)
            if reorder:
                os.environ.pop("TW_PIPE_NO_REORDER", None)
            else:
                os.environ["TW_PIPE_NO_REORDER"] = "1"
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 1000, ep, ep.zmin, out=out)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            print("tiles %6d chunks %2d reorder %d: %.4f s" % (nt, ch, reorder, best), flush=True)
    del out
