"""GPU parity of tw_tile_shadows_batch vs the CPU oracle, which tests/test_oracle_vs_reference.py pins against the reference's own
calc_mesh_shadows / mesh_shadow_gen / do_line_clip: masks and outgoing shadow heights, bit for bit, single tiles and chained blocks of tiles."""
import numpy as np
import pytest

from cases import convert, HM_CFG

pytestmark = pytest.mark.gpu


def _params(tw, cfg, S, zlo, zhi, lp):
    sp = tw.ShadowParams()
    sp.x_scene_size, sp.y_scene_size = cfg.scene_size[0], cfg.scene_size[1]
    sp.dx_val, sp.dy_val = float(cfg.dx_val), float(cfg.dy_val)
    sp.dx_val_inv, sp.dy_val_inv = 1.0 / np.float32(cfg.dx_val), 1.0 / np.float32(cfg.dy_val)
    sp.xy_sum_size, sp.zmin, sp.zmax, sp.no_shadow = 2 * S, zlo, zhi, 0
    for d in range(3):
        sp.lpos[d] = lp[d]
    return sp


@pytest.mark.parametrize("S,side", [(64, 1), (32, 4), (128, 3), (17, 5)])
def test_tile_shadows_vs_oracle(tw, scene, oracle, ctx, beq, S, side):
    import torch
    zv = S + 2
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0, mesh_size=(S, S, 1))
    txy = [(tx - 1, ty + 3) for ty in range(side) for tx in range(side)]
    if side == 5:
        txy = [t for i, t in enumerate(txy) if i % 4 != 1]          # a batch with holes: missing neighbours mean "no incoming heights"
    origins = [(tx * S, ty * S) for tx, ty in txy]
    tiles = ctx.heightgen_tiles(origins, cfg.mesh_size, float(cfg.dx_val), float(cfg.dy_val), zv, cfg.height_params())
    tiles = ((tiles - np.float32(tiles.mean())) * np.float32(3.0)).astype(np.float32)      # around 0: the rays run at z = 0 and are clipped against [zmin, zmax]
    zlo, zhi = float(tiles.min()) - 0.5, float(tiles.max()) + 0.5
    assert zlo < 0.0 < zhi
    for lp in ((3.0, 2.0, 0.4), (-4.0, 1.0, 0.3), (1.0, -5.0, 0.5), (-2.0, -3.0, 2.0), (0.2, 6.0, 0.15), (5.0, 0.0, 1.0), (0.0, 0.0, 5.0), (2.0, 1.0, zlo - 1.0)):
        sp = _params(tw, cfg, S, zlo, zhi, lp)
        mo, oxo, oyo = oracle.tile_shadows_batch(tiles, txy, convert(sp, oracle.ShadowParams))
        m, ox, oy = ctx.tile_shadows(tiles, txy, sp)
        assert np.array_equal(m, mo), (lp, int((m != mo).sum()))
        assert beq(ox, oxo) == 0 and beq(oy, oyo) == 0, lp
    sp = _params(tw, cfg, S, zlo, zhi, (3.0, 2.0, 0.4))
    mo, _, _ = oracle.tile_shadows_batch(tiles, txy, convert(sp, oracle.ShadowParams))
    assert 0 < (mo == 2).mean() < 1
    dz = torch.from_numpy(np.ascontiguousarray(tiles)).cuda()
    dm = torch.empty(tiles.shape, dtype=torch.uint8, device="cuda")
    ctx.tile_shadows(dz, txy, sp, out=dm)
    assert np.array_equal(dm.cpu().numpy(), mo)
    sp.no_shadow = 1
    assert not ctx.tile_shadows(tiles, txy, sp)[0].any()


def test_tile_shadows_golden(tw, ctx, beq):
    """tests/golden/shadows.npz: the reference's own calc_mesh_shadows over a chained 3x3 block of tiles (made by tests/golden/make_golden_shadows.py)."""
    import os
    from test_oracle_golden import shadow_params
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shadows.npz"))
    txy = [tuple(int(v) for v in t) for t in g["tile_xy"]]
    shadowed = 0
    for li, lp in enumerate(g["lights"]):
        m, ox, oy = ctx.tile_shadows(g["tiles"], txy, shadow_params(tw.ShadowParams, g["params"], lp))
        assert np.array_equal(m, g["smask_%d" % li]), li
        assert beq(ox, g["sh_out_x_%d" % li]) == 0 and beq(oy, g["sh_out_y_%d" % li]) == 0, li
        shadowed += int((m == 2).sum())
    assert shadowed > 1000
