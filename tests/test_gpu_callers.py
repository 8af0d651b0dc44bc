"""GPU parity for the callers' pieces next to the generators: gen_mesh() ground-mode flow (BASELINE config 1) and the tail of
tile_t::create_zvals (sub-block z ranges, water bbox), through the C ABI, vs the oracle / golden fixtures of the reference's gen_mesh()."""
import os

import numpy as np
import pytest

from cases import convert, HM_CFG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["gm_cfg1", "gm_cfg1_eroded", "gm_simplex"])
def test_gen_mesh_matches_reference_golden(tw, scene, ctx, beq, name):
    h = np.load(os.path.join(GOLD, "height.npz"))
    mode, seed, ff, iters, mhs, smag, sfreq, sbias = h[name + "_args"]
    hmap = dict(sine_mag=float(smag), sine_freq=float(sfreq), sine_bias=float(sbias)) if smag else {}
    cfg = scene.SceneConfig(mesh_gen_mode=int(mode), mesh_freq_filter=int(ff), mesh_seed=int(seed), mesh_height=float(mhs), hmap=hmap)
    cfg.clip_hd1 = lambda: 0.5           # the fixture was generated with clip_hd1 = 0.5, relh_adj_tex = 0
    mesh, z, _ = scene.gen_mesh(ctx, cfg, erosion_iters=int(iters))
    assert beq(mesh, h[name]) == 0
    got = np.array([z[k] for k in ("zmin", "zmax", "zmax_est", "zbottom", "ztop", "water_plane_z")], np.float32)
    assert beq(got, h[name + "_zvals"]) == 0


def test_gen_mesh_matches_oracle_with_offsets(tw, scene, oracle, ctx, beq):
    for mode, seed, xoff2, yoff2 in ((0, 6, 37, -512), (2, 9, -1000, 250)):
        cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=seed, hmap=dict(HM_CFG, volcano_width=300.0, volcano_height=2.0))
        cfg.clip_hd1 = lambda: 0.4
        mesh, z, sp = scene.gen_mesh(ctx, cfg, erosion_iters=300, xoff2=xoff2, yoff2=yoff2)
        cfg2 = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=seed, hmap=dict(HM_CFG, volcano_width=300.0, volcano_height=2.0))
        hp = convert(cfg2.height_params(), oracle.HeightParams)
        mo, zo, spo = oracle.gen_mesh(hp, seed=seed, xoff2=xoff2, yoff2=yoff2, erosion_iters=300, ep=oracle.ErosionParams(1.0, 0.0, 0.0625, 0.0, 0.0, 0.0, 0.4))
        assert beq(sp, spo) == 0 and beq(mesh, mo) == 0
        assert all(np.float32(z[k]) == np.float32(zo[k]) for k in zo)


def test_tile_bounds(tw, scene, oracle, ctx):
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(128, 128, 1))
    hp = cfg.height_params()
    for S in (128, 64):
        zv = S + 2
        origins = [(tx * S * 3, ty * S * 5 - 2000) for ty in range(4) for tx in range(5)]
        tiles = ctx.heightgen_tiles(origins, (S, S), float(cfg.dx_val), float(cfg.dy_val), zv, hp)
        wpz_max = float(np.median(tiles))        # get_max_sea_level(): about half the cells under water
        got = ctx.tile_bounds(tiles, wpz_max, float(cfg.dx_val), float(cfg.dy_val), S)
        exp = oracle.tile_bounds(tiles, wpz_max, float(cfg.dx_val), float(cfg.dy_val), S)
        for g, e in zip(got, exp):
            assert bytes(g) == bytes(e)
        import torch
        got_dev = ctx.tile_bounds(torch.from_numpy(tiles).cuda(), wpz_max, float(cfg.dx_val), float(cfg.dy_val), S)
        assert all(bytes(g) == bytes(e) for g, e in zip(got_dev, exp))


def _oracle_proc_gen(oracle, cfg, hp_o, w, h, iters, ep_o):
    """heightmap_t::proc_gen restated from the pinned oracle pieces (src/heightmap.cpp:130-215, src/mesh_gen.cpp:120-131)."""
    f32 = np.float32
    vals = oracle.heightgen_2d(oracle.Grid2D(-0.5 * w, -0.5 * h, float(cfg.dx_val), float(cfg.dy_val), w, h), hp_o, None, 1, 0)
    moves = 0
    if iters:
        vals, moves = oracle.apply_erosion(vals, float(vals.min()), iters, ep_o)
    min_z, max_z = f32(vals.min()), f32(vals.max())
    dz = max(f32(1.0E-12), f32(max_z - min_z))
    dz255 = f32(float(dz) / 255.0)
    mhs, mszi, R = f32(hp_o.mesh_height_scale), f32(hp_o.mesh_scale_z_inv), f32(0.0008)
    mfs = f32(dz255 / f32(f32(R * mhs) * mszi))
    mtz = f32(min_z / mszi)
    mult, add = f32(f32(f32(R * mhs) * mfs) * mszi), f32(mtz * mszi)
    img, bad = oracle.from_floats_u16(vals, float(mult), float(add))
    assert bad == 0
    return img, vals, (float(min_z), float(max_z), float(mult), float(add)), moves


@pytest.mark.parametrize("mode,iters,w,h", [(1, 0, 320, 200), (4, 800, 256, 256), (0, 500, 130, 258)])
def test_proc_gen_heightmap(tw, scene, oracle, ctx, beq, mode, iters, w, h):
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
    hp, ep = cfg.height_params(), cfg.erosion_params()
    if mode == 0:
        ctx.set_sine_params(cfg.sine_params())
        hp_sp = cfg.sine_params()
    img_o, vals_o, (mn, mx, mult, add), moves = _oracle_proc_gen_sp(oracle, cfg, hp, w, h, iters, ep, cfg.sine_params() if mode == 0 else None)
    vals = np.empty((h, w), np.float32)
    img, info, _ = ctx.proc_gen_heightmap(w, h, float(cfg.dx_val), float(cfg.dy_val), hp, iters, ep, vals=vals)
    assert beq(vals, vals_o) == 0
    assert (info.min_z, info.max_z, info.val_mult, info.val_add) == (mn, mx, mult, add)
    assert info.erosion_moves == moves
    assert np.array_equal(img, img_o)
    import torch
    d_img = torch.empty(2 * w * h, dtype=torch.uint8, device="cuda")
    ctx.proc_gen_heightmap(w, h, float(cfg.dx_val), float(cfg.dy_val), hp, iters, ep, data16=d_img)
    assert np.array_equal(d_img.cpu().numpy(), img_o)


def _oracle_proc_gen_sp(oracle, cfg, hp, w, h, iters, ep, sp):
    hp_o, ep_o = convert(hp, oracle.HeightParams), convert(ep, oracle.ErosionParams)
    if sp is None:
        return _oracle_proc_gen(oracle, cfg, hp_o, w, h, iters, ep_o)
    orig = oracle.heightgen_2d
    try:
        oracle.heightgen_2d = lambda g, p, _sp, gl, mss: orig(g, p, sp, gl, mss)
        return _oracle_proc_gen(oracle, cfg, hp_o, w, h, iters, ep_o)
    finally:
        oracle.heightgen_2d = orig


@pytest.mark.parametrize("mode", [0, 1, 4])
def test_tile_normals_and_ao_golden(tw, ctx, mode):
    """N1: normal map (the reference's own get_norm arithmetic) and AO (restated ray march on reference heights) of the golden tile."""
    from test_oracle_golden import tile_case
    h = np.load(os.path.join(GOLD, "tiles.npz"))
    n = "m%d" % mode
    hp, sp, a = tile_case(tw, h, mode)
    if sp is not None:
        ctx.set_sine_params(sp)
    x1, y1, S, zv, dx, dy, half_dxy = int(a[1]), int(a[2]), int(a[3]), int(a[4]), float(a[5]), float(a[6]), float(a[7])
    tile = h["tile_" + n]
    rgba, mnz = ctx.tile_normals(tile[None], dx, dy)
    assert np.array_equal(rgba[0], h["normals_" + n]) and mnz[0] == h["min_normal_z_" + n]
    ao = ctx.tile_ao(tile[None], [(x1, y1)], (S, S), dx, dy, hp, half_dxy)
    assert np.array_equal(ao[0], h["ao_" + n])           # = the reference's own calc_mesh_ao_lighting (tests/golden/make_golden_tiles.py)


def test_create_zvals_with_ao_gpu_mode_golden(tw, ctx, beq):
    """enable_tiled_mesh_ao + mesh_gen_mode 4: tile_t::create_zvals generates the (stride+72)^2 context once, cuts zvals out of it and erodes them;
    calc_mesh_ao_lighting tests the rays against the UN-eroded context. Golden = the reference's own functions."""
    from test_oracle_golden import tile_case
    h = np.load(os.path.join(GOLD, "tiles.npz"))
    hp, _, a = tile_case(tw, h, 4)
    x1, y1, S, zv, dx, dy, half_dxy = int(a[1]), int(a[2]), int(a[3]), int(a[4]), float(a[5]), float(a[6]), float(a[7])
    e = h["ero_m4"]
    ep = tw.ErosionParams(*[float(v) for v in e[2:]])
    z, ao = ctx.create_zvals_ao_batch([(x1, y1)], (S, S), dx, dy, zv, hp, 0, ep, float(e[0]), half_dxy)
    assert beq(z[0], h["tile_m4c"]) == 0 and np.array_equal(ao[0], h["ao_m4c"])
    z, ao, mm = ctx.create_zvals_ao_batch([(x1, y1)], (S, S), dx, dy, zv, hp, int(e[1]), ep, float(e[0]), half_dxy, want_minmax=True)
    assert beq(z[0], h["tile_m4e"]) == 0 and np.array_equal(ao[0], h["ao_m4e"])
    assert (mm[0, 0], mm[0, 1]) == (h["tile_m4e"].min(), h["tile_m4e"].max())
    # the two-call route gives the same AO when handed the eroded zvals (context regenerated inside tw_tile_ao_batch)
    assert np.array_equal(ctx.tile_ao(h["tile_m4e"][None], [(x1, y1)], (S, S), dx, dy, hp, half_dxy)[0], h["ao_m4e"])


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_create_zvals_with_ao_batch_vs_oracle(tw, scene, oracle, ctx, beq, mode):
    """tw_create_zvals_ao_batch over a batch, every gen mode, host and device outputs, vs the oracle's restatement of the two reference flows
    (which tests/test_oracle_vs_reference.py pins against the reference's own function bodies)."""
    import torch
    S, zv = 64, 66
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    hp, ep = cfg.height_params(), cfg.erosion_params()
    sp = None
    if mode == 0:
        sp = cfg.sine_params()
        ctx.set_sine_params(sp)
    dx, dy, hd = float(cfg.dx_val), float(cfg.dy_val), 0.5 * float(cfg.dx_val + cfg.dy_val)
    origins = [(tx * S * 9 - 1500, ty * S * 7 + 200) for ty in range(2) for tx in range(4)]
    hp_o, ep_o = convert(hp, oracle.HeightParams), convert(ep, oracle.ErosionParams)
    csz = zv - 1 + 72
    contexts = np.stack([oracle.heightgen_2d(oracle.Grid2D(x1 - 36 - S // 2, y1 - 36 - S // 2, dx, dy, csz, csz), hp_o, sp, 1, 0) for x1, y1 in origins])
    if mode >= 3:
        raw = np.ascontiguousarray(contexts[:, 36:36 + zv, 36:36 + zv])
    else:
        raw = np.stack([oracle.heightgen_2d(oracle.Grid2D(x1 - S // 2, y1 - S // 2, dx, dy, zv, zv), hp_o, sp, 1, 0) for x1, y1 in origins])
    exp_z = np.stack([oracle.apply_erosion(t, ep.zmin, 300, ep_o)[0] for t in raw])
    exp_ao = oracle.tile_ao(exp_z, contexts, hd, use_ao_zvals=(mode >= 3))
    z, ao = ctx.create_zvals_ao_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 300, ep, ep.zmin, hd)
    assert beq(z, exp_z) == 0 and np.array_equal(ao, exp_ao)
    assert (exp_z != raw).any() and ao.min() < 255
    dz = torch.empty((len(origins), zv, zv), dtype=torch.float32, device="cuda")
    dao = torch.empty((len(origins), zv - 1, zv - 1), dtype=torch.uint8, device="cuda")
    ctx.create_zvals_ao_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 300, ep, ep.zmin, hd, out=dz, ao=dao)
    assert beq(dz.cpu().numpy(), exp_z) == 0 and np.array_equal(dao.cpu().numpy(), exp_ao)
    assert np.array_equal(ctx.tile_ao(exp_z, origins, cfg.mesh_size, dx, dy, hp, hd), exp_ao)


def test_tile_normals_and_ao_batch_vs_oracle(tw, scene, oracle, ctx, beq):
    """A batch of eroded tiles, host and device pointers, vs the oracle (context grids from the oracle's heightgen)."""
    import torch
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(128, 128, 1))
    hp, ep = cfg.height_params(), cfg.erosion_params()
    S, zv = 128, 130
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    origins = [(tx * S, ty * S - 640) for ty in range(2) for tx in range(3)]
    tiles = ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 200, ep, ep.zmin)
    hp_o = convert(hp, oracle.HeightParams)
    csz = zv - 1 + 72
    contexts = np.stack([oracle.heightgen_2d(oracle.Grid2D(x1 - 36 - S // 2, y1 - 36 - S // 2, dx, dy, csz, csz), hp_o, None, 1, 0) for x1, y1 in origins])
    exp_n, exp_m = oracle.tile_normals(tiles, dx, dy)
    exp_ao = oracle.tile_ao(tiles, contexts, 0.0625)
    rgba, mnz = ctx.tile_normals(tiles, dx, dy)
    assert np.array_equal(rgba, exp_n) and beq(mnz, exp_m) == 0
    ao = ctx.tile_ao(tiles, origins, cfg.mesh_size, dx, dy, hp, 0.0625)
    assert np.array_equal(ao, exp_ao)
    assert ao.min() < ao.max() == 255
    d_tiles = torch.from_numpy(tiles).cuda()
    d_rgba = torch.empty((len(origins), zv - 1, zv - 1, 4), dtype=torch.uint8, device="cuda")
    d_ao = torch.empty((len(origins), zv - 1, zv - 1), dtype=torch.uint8, device="cuda")
    ctx.tile_normals(d_tiles, dx, dy, out=d_rgba)
    ctx.tile_ao(d_tiles, origins, cfg.mesh_size, dx, dy, hp, 0.0625, out=d_ao)
    assert np.array_equal(d_rgba.cpu().numpy(), exp_n) and np.array_equal(d_ao.cpu().numpy(), exp_ao)
