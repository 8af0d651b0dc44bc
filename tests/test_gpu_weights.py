"""GPU parity of tw_tile_weights_batch (terrain weights texture): the committed reference fixture (the reference's own tile_t::create_texture),
and the CPU oracle on generated tiles with random parameter sets - every RGBA byte and the has_any_grass flags. The kernel's arithmetic is also checked on the
host without a GPU (tests/test_weights_host.py)."""
import os

import numpy as np
import pytest

from cases import convert, HM_CFG
from test_oracle_golden import weights_golden_case
from test_weights_host import weight_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_weights_golden(tw, scene, ctx):
    import torch
    g = np.load(os.path.join(GOLD, "weights.npz"))
    for n in ("m1", "m4"):
        ctx.set_sine_params(g["sine_params_" + n])
        for ci in range(3):
            hmap = scene.SceneConfig(hmap=HM_CFG).height_params().hmap
            hp, wp, S, dx, dy = weights_golden_case(tw.HeightParams, tw.WeightParams, hmap, g, n, ci)
            tiles, org, corners = g["tiles_" + n], g["origins_" + n], g["corners_" + n]
            w, flags = ctx.tile_weights(tiles, org, (S, S, 1), dx, dy, hp, wp, corners)
            exp = g["weights_%s_%d" % (n, ci)]
            assert np.array_equal(w, exp), (n, ci, int((w != exp).sum()))
            assert np.array_equal(flags, g["grass_%s_%d" % (n, ci)])
            if ci == 0:                # device-resident inputs and output
                dw = torch.empty(exp.shape, dtype=torch.uint8, device="cuda")
                ctx.tile_weights(torch.from_numpy(tiles).cuda(), org, (S, S, 1), dx, dy, hp, wp, torch.from_numpy(corners).cuda(), out=dw, want_grass_flags=False)
                assert np.array_equal(dw.cpu().numpy(), exp)


@pytest.mark.parametrize("mode,S,side", [(1, 64, 3), (0, 32, 4), (4, 128, 2), (2, 17, 5)])
def test_weights_vs_oracle(tw, scene, oracle, ctx, mode, S, side):
    rng = np.random.default_rng(S)
    zv = S + 2
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    ctx.set_sine_params(cfg.sine_params())
    origins = [(tx * S - 5 * S, ty * S + 2 * S) for ty in range(side) for tx in range(side)]
    hp = cfg.height_params()
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    tiles = ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp)
    tiles = ((tiles - np.float32(tiles.mean())) * np.float32(0.4 / max(1e-6, float(tiles.std())))).astype(np.float32)
    corners = rng.uniform(-0.2, 1.3, (len(origins), 8)).astype(np.float32)
    rand = oracle.weights_noise(convert(hp, oracle.HeightParams), cfg.sine_params(), origins, (S, S), dx, dy, S + 1)
    for wp in weight_cases(tw.WeightParams, rng, float(tiles.min()), float(tiles.max()), S, dx, dy):
        exp, eflags = oracle.tile_weights(tiles, rand, corners, convert(wp, oracle.WeightParams))
        w, flags = ctx.tile_weights(tiles, origins, cfg.mesh_size, dx, dy, hp, wp, corners)
        assert np.array_equal(w, exp), int((w != exp).sum())
        assert np.array_equal(flags, eflags)
    assert len(np.unique(exp.reshape(-1, 4), axis=0)) > 30


def test_weights_argument_errors(tw, scene, ctx):
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(32, 32, 1))
    ctx.set_sine_params(cfg.sine_params())
    z = np.zeros((1, 34, 34), np.float32)
    wp = weight_cases(tw.WeightParams, np.random.default_rng(0), -1.0, 1.0, 32, 0.25, 0.25, n=1)[0]
    wp.tex_class[1] = 0                                         # sand twice, dirt never: the reference asserts (get_texture_ixs)
    with pytest.raises(tw.TwError):
        ctx.tile_weights(z, [(0, 0)], (32, 32, 1), 0.25, 0.25, cfg.height_params(), wp, np.zeros((1, 8), np.float32))
