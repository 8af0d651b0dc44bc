"""GPU parity for the streaming passes (heightmap 16-bit pack/unpack, min/max) vs the oracle."""
import numpy as np
import pytest

from cases import convert

pytestmark = pytest.mark.gpu


def test_from_to_floats_u16(tw, oracle, ctx):
    rng = np.random.default_rng(5)
    for n in (1, 7, 4096, 100003):
        vals = rng.uniform(-3.0, 5.0, n).astype(np.float32)
        mn, mx = float(vals.min()), float(vals.max())
        mult, add = np.float32(max(1e-12, mx - mn) / 255.0), np.float32(mn)   # set_mesh_height_scales_for_zval_range(min_z, dz/255)
        exp, bad = oracle.from_floats_u16(vals, float(mult), float(add))
        assert bad == 0
        got = ctx.from_floats_u16(vals, float(mult), float(add))
        assert np.array_equal(got, exp)
        back = ctx.to_floats_u16(got, float(mult), float(add))
        assert np.array_equal(back, oracle.to_floats_u16(exp, float(mult), float(add)))
        assert np.abs(back - vals).max() <= float(mult) / 256.0 * 1.01 + 1e-6          # round trip within one 16-bit step


def test_from_floats_out_of_range_is_an_error(tw, ctx):
    vals = np.array([0.0, 300.0, 1.0], np.float32)
    with pytest.raises(tw.TwError) as e:
        ctx.from_floats_u16(vals, 1.0, 0.0)
    assert e.value.status == tw.TW_ERR_ARG


def test_minmax(tw, ctx):
    rng = np.random.default_rng(6)
    for n in (1, 33, 1 << 20):
        v = rng.standard_normal(n).astype(np.float32)
        assert ctx.minmax(v) == (float(v.min()), float(v.max()))


def test_argument_errors(tw, scene, ctx):
    cfg = scene.SceneConfig(mesh_gen_mode=1)
    hp = cfg.height_params()
    with pytest.raises(tw.TwError):
        ctx.heightgen_2d(tw.Grid2D(0, 0, 1, 1, 0, 4), hp, out=np.empty((4, 1), np.float32))   # nx == 0 (the reference asserts)
    hp.gen_mode = 9
    with pytest.raises(tw.TwError):
        ctx.heightgen_2d(tw.Grid2D(0, 0, 1, 1, 4, 4), hp)


def test_heightmap_texture_tiles(tw, oracle, ctx, beq):
    """N2: heightmap-texture mode of tile_t::create_zvals (get_clamped_height) - golden outputs of the linked reference, then a batch with all
    three edge modes, host and device pointers, vs the oracle."""
    import os
    import torch
    from test_oracle_golden import hmap_cases
    h = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiles.npz"))
    for img, hs, org, exp in hmap_cases(tw, h):
        assert beq(ctx.heightmap_sample_tiles(img, hs, [org], 34)[0], exp) == 0
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (300, 257, 2), dtype=np.uint8)
    origins = [(int(x), int(y)) for x, y in rng.integers(-900, 900, (40, 2))]
    for edge in (0, 1, 2):
        for ms in (1.0, 0.45, 2.2):
            hs = tw.HmapSampler(257, 300, edge, ms, 0.0012, 1.7, -0.3, 0.8)
            exp = oracle.hmap_sample_tiles(img, convert(hs, oracle.HmapSampler), origins, 66)
            assert beq(ctx.heightmap_sample_tiles(img, hs, origins, 66), exp) == 0
            d_out = torch.empty((len(origins), 66, 66), dtype=torch.float32, device="cuda")
            ctx.heightmap_sample_tiles(torch.from_numpy(img).cuda(), hs, origins, 66, out=d_out)
            assert beq(d_out.cpu().numpy(), exp) == 0
