"""One big map, the reference's serial droplet order: M_SPEC (speculative parallel walks, in-order commit; the default) vs the plain serial walk
(TW_EROSION_MODE=global), bit-compared. BASELINE config 3 shape.   python tools/bench_spec.py [--size 8192] [--iters 1000 100000 1000000]"""
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


def run(ctx, d, zmin, iters, ep, mode, window=None):
    for k in ("TW_EROSION_MODE", "TW_SPEC_WINDOW"):
        os.environ.pop(k, None)
    if mode != "auto":
        os.environ["TW_EROSION_MODE"] = mode
    if window:
        os.environ["TW_SPEC_WINDOW"] = str(window)
    best, out = 1e30, None
    for rep in range(2):
        work = d.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.erode(work, zmin, iters, ep)
        torch.cuda.synchronize()
        best, out = min(best, time.perf_counter() - t0), work
    return best, out, ctx.last_erosion_steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs="+", default=[8192])
    ap.add_argument("--iters", type=int, nargs="+", default=[1000, 100000])
    ap.add_argument("--serial-max", type=int, default=100000)
    ap.add_argument("--windows", type=int, nargs="+", default=[2048])
    args = ap.parse_args()
    ctx = tw.Context(0)
    for N in args.size:
        cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
        d = torch.empty((N, N), dtype=torch.float32, device="cuda")
        _, (zmin, zmax) = ctx.heightgen_2d(cfg.heightmap_grid(N, N), cfg.height_params(), out=d, want_minmax=True)
        ep = cfg.erosion_params()
        for iters in args.iters:
            ref = None
            if iters <= args.serial_max:
                t, ref, steps = run(ctx, d, zmin, iters, ep, "global")
                print("map %d^2 droplets %8d serial walk      : %.5f s  %.3e droplets/s (%.1f moves/droplet)" % (N, iters, t, iters / t, steps / iters), flush=True)
            for w in args.windows:
                t, out, steps2 = run(ctx, d, zmin, iters, ep, "spec", w)
                same = "" if ref is None else ("  identical to the serial walk: %s, moves equal: %s" % (bool(torch.equal(out.view(torch.int32), ref.view(torch.int32))), steps2 == steps))
                print("map %d^2 droplets %8d speculative B=%-5d: %.5f s  %.3e droplets/s (%.1f moves/droplet)%s" % (N, iters, w, t, iters / t, steps2 / iters, same), flush=True)


if __name__ == "__main__":
    main()
