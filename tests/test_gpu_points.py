"""GPU parity for the batched point queries: tw_eval_points vs the reference's golden outputs and vs the oracle."""
import os

import numpy as np
import pytest

from cases import convert, HM_ALL, HM_CFG
from test_oracle_golden import point_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_point_queries_match_reference_golden(tw, ctx, beq):
    h = np.load(os.path.join(GOLD, "points.npz"))
    n = 0
    for key, hp, pq, xy, sp, exp in point_cases(tw, h):
        if sp is not None:
            ctx.set_sine_params(sp)
        assert beq(ctx.eval_points(xy, hp, pq), exp) == 0, key
        n += 1
    assert n >= 50


@pytest.mark.parametrize("mode", [0, 1, 2, 4])
def test_point_queries_large_batch_vs_oracle(tw, scene, oracle, ctx, beq, mode):
    """A batch the size of a city's building list, host and device pointers; scrolled get_exact_zval with every post-process switched on."""
    import torch
    rng = np.random.default_rng(mode)
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_gen_shape=mode % 3, mesh_freq_filter=1, mesh_seed=3, hmap=HM_ALL, zmax_est=2.3)
    hp = cfg.height_params()
    sp = cfg.sine_params()
    ctx.set_sine_params(sp)
    n = 20000 if mode != 4 else 4000
    xy = rng.uniform(-60.0, 60.0, (n, 2)).astype(np.float32)
    for kind, xy_scale, no_xyoff in ((tw.PQ_EXACT_ZVAL, 1.0, 0), (tw.PQ_SIN_TERMS_SCALED, 16.0, 0), (tw.PQ_SIN_TERMS, 1.0, 0)):
        pq = tw.PointQuery(kind, xy_scale, 128, 128, 4.0, 4.0, -3000, 777, no_xyoff)
        exp = oracle.eval_points(xy, convert(hp, oracle.HeightParams), convert(pq, oracle.PointQuery), sp)
        assert beq(ctx.eval_points(xy, hp, pq), exp) == 0
        d_xy, d_out = torch.from_numpy(xy).cuda(), torch.empty(n, dtype=torch.float32, device="cuda")
        ctx.eval_points(d_xy, hp, pq, out=d_out)
        assert beq(d_out.cpu().numpy(), exp) == 0


def test_point_query_agrees_with_grid(tw, scene, ctx, beq):
    """get_exact_zval at the cell positions of a glaciated grid equals mesh_xy_grid_cache_t::eval_index there in the noise modes (the
    reference computes both from get_noise_zval; src/mesh_gen.cpp:762 vs :808,818) when the index-space coordinates coincide."""
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap={}, zmax_est=2.3)
    hp = cfg.height_params()
    n = 64
    # grid cell (i, j) of build_arrays(x0 = -64, dx = DX_VAL) evaluates get_noise_zval(i - 64, ...); eval_mesh_sin_terms_scaled(xval, ., 1) evaluates
    # get_noise_zval(xval - 64, ...): xval = i
    z = ctx.heightgen_2d(cfg.heightmap_grid(n, n), hp, enable_glaciate=0)
    g = cfg.heightmap_grid(n, n)
    jj, ii = np.mgrid[0:n, 0:n]
    xval = (ii + g.x0 + 64).astype(np.float32)
    yval = (jj + g.y0 + 64).astype(np.float32)
    xy = np.stack([xval.ravel(), yval.ravel()], 1)
    pq = tw.PointQuery(tw.PQ_SIN_TERMS_SCALED, 1.0, 128, 128, 4.0, 4.0, 0, 0, 0)
    assert beq(ctx.eval_points(xy, hp, pq).reshape(n, n), z) == 0


def test_point_query_errors(tw, ctx):
    hp = tw.HeightParams()
    with pytest.raises(tw.TwError):
        ctx.eval_points(np.zeros((4, 2), np.float32), hp, tw.PointQuery(7, 1.0, 128, 128, 4.0, 4.0, 0, 0, 0))
    assert ctx.eval_points(np.zeros((0, 2), np.float32), hp, tw.PointQuery(0, 1.0, 128, 128, 4.0, 4.0, 0, 0, 0)).shape == (0,)
