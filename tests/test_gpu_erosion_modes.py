"""GPU parity of every shipped erosion code path ("parity gaps"): all lane-group sizes G of droplet_kernel, the three
height-residency modes (global / shared-memory window / whole map in shared memory), the heavy/light split of a batch, the multi-chunk
fused pipeline, the BASELINE config-5 shape, and the committed golden fixtures - all bit-exact against the CPU oracle (which is pinned
against the unmodified reference objects, tests/test_oracle_vs_reference.py). The selection knobs are environment variables the library
re-reads on every call (TW_EROSION_*, TW_PIPE_CHUNKS)."""
import os

import numpy as np
import pytest

from cases import convert, HM_CFG

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ENV_KEYS = ("TW_EROSION_MODE", "TW_EROSION_LANES", "TW_EROSION_SMEM_LANES", "TW_EROSION_WIN", "TW_EROSION_WIN_MIN_MOVES", "TW_EROSION_WHOLE_MAX",
            "TW_EROSION_WINDOW_ALL", "TW_EROSION_HEAVY", "TW_PIPE_CHUNKS", "TW_SPEC_WINDOW", "TW_SPEC_LOG", "TW_SPEC_TILES", "TW_SPEC_VIEWS")


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ENV_KEYS:
        monkeypatch.delenv(k, raising=False)


def _terrain(scene, ctx, n, m, mode=1, seed=1):
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=seed, hmap=HM_CFG, zmax_est=2.0)
    return cfg, ctx.heightgen_2d(cfg.heightmap_grid(n, m), cfg.height_params())


def _eparams(tw, z):
    zmin, zmax = float(z.min()), float(z.max())
    return zmin, [tw.ErosionParams(1.0, zmin - 10, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 0.5),                      # no ocean: long walks, border exits
                  tw.ErosionParams(1.0, zmin + 0.2 * (zmax - zmin), 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 2.0)]  # ocean stop, all dirt (NaN hazard path)


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_global_mode_every_group_size(tw, scene, oracle, ctx, beq, monkeypatch, lanes):
    """pick_group() dispatches G = 8 / 16 for >= 65536 / 32768 maps (BASELINE config 5 at N = 1, 2): every G must equal the serial order."""
    monkeypatch.setenv("TW_EROSION_MODE", "global")
    monkeypatch.setenv("TW_EROSION_LANES", str(lanes))
    cfg, z = _terrain(scene, ctx, 130, 97)
    zmin, eps = _eparams(tw, z)
    for ep in eps:
        zc, steps = oracle.apply_erosion(z, zmin, 700, convert(ep, oracle.ErosionParams))
        assert beq(ctx.erode(z.copy(), zmin, 700, ep), zc) == 0
        assert ctx.last_erosion_steps == steps
        assert beq(ctx.erode_parallel(z.copy(), zmin, 700, ep, num_threads=1), zc) == 0      # M_ATOMIC with one group = serial order
    # tile batch (several maps per warp for G < 32)
    S, zv = 32, 34
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0, mesh_size=(S, S, 1))
    origins = [(tx * S * 7, ty * S * 5) for ty in range(5) for tx in range(7)]
    tiles = ctx.heightgen_tiles(origins, cfg.mesh_size, float(cfg.dx_val), float(cfg.dy_val), zv, cfg.height_params())
    tzmin = float(tiles.min())
    ep = tw.ErosionParams(1.0, tzmin + 0.3 * float(tiles.max() - tiles.min()), 0.0625, tzmin - 0.1, float(tiles.max()) + 0.1, 0.0, 0.5)
    exp = np.stack([oracle.apply_erosion(t, tzmin, 150, convert(ep, oracle.ErosionParams))[0] for t in tiles])
    assert beq(ctx.erode_tiles(tiles.copy(), 150, ep, min_zval_all=tzmin), exp) == 0


@pytest.mark.parametrize("mode,win,kmin,lanes", [("window", 32, 0, 32), ("window", 32, 2, 32), ("window", 8, 0, 32), ("window", 16, 3, 16), ("window", 64, 0, 32),
                                                 ("window", 32, 0, 8), ("window", 24, 1, 1), ("whole", 0, 0, 32), ("whole", 0, 0, 16), ("whole", 0, 0, 4), ("whole", 0, 0, 1)])
@pytest.mark.parametrize("n,m,iters", [(130, 130, 1000), (64, 200, 500), (17, 9, 200), (200, 150, 1500)])
def test_shared_memory_modes_single_map(tw, scene, oracle, ctx, beq, monkeypatch, mode, win, kmin, lanes, n, m, iters):
    """M_WINDOW (sliding shared-memory window, write-through) and M_WHOLE (padded map built and walked in shared memory) == serial order,
    for every window size / minimum-moves setting / lane-group size, including maps smaller than the window and border exits."""
    if mode == "whole" and (n + 8) * (m + 8) * 4 > 227 * 1024:
        pytest.skip("map does not fit in shared memory")
    monkeypatch.setenv("TW_EROSION_MODE", mode)
    monkeypatch.setenv("TW_EROSION_SMEM_LANES", str(lanes))
    if win:
        monkeypatch.setenv("TW_EROSION_WIN", str(win))
        monkeypatch.setenv("TW_EROSION_WIN_MIN_MOVES", str(kmin))
    cfg, z = _terrain(scene, ctx, n, m)
    zmin, eps = _eparams(tw, z)
    for ep in eps:
        zc, steps = oracle.apply_erosion(z, zmin, iters, convert(ep, oracle.ErosionParams))
        zg = ctx.erode(z.copy(), zmin, iters, ep)
        assert beq(zg, zc) == 0, "max abs diff %g" % np.nanmax(np.abs(zg - zc))
        assert ctx.last_erosion_steps == steps


@pytest.mark.parametrize("mode", ["window", "whole", "auto"])
def test_flat_and_clamp_cases(tw, oracle, ctx, beq, monkeypatch, mode):
    if mode != "auto":
        monkeypatch.setenv("TW_EROSION_MODE", mode)
    z = np.zeros((100, 100), np.float32)
    z[50:, :] = 0.001
    ep = tw.ErosionParams(1.0, -5.0, 0.0625, -1.0, 1.0, 0.0, 0.5)
    zc, _ = oracle.apply_erosion(z, -1.0, 300, convert(ep, oracle.ErosionParams))   # flat ground: random-direction fallback
    assert beq(ctx.erode(z.copy(), -1.0, 300, ep), zc) == 0
    zc, _ = oracle.apply_erosion(z, 0.0005, 3, convert(ep, oracle.ErosionParams))   # min_zval clamp applies to untouched cells too
    assert beq(ctx.erode(z.copy(), 0.0005, 3, ep), zc) == 0


@pytest.mark.parametrize("mode", ["window", "whole", "auto"])
def test_tile_batches_in_shared_memory_modes(tw, scene, oracle, ctx, beq, monkeypatch, mode):
    """tile_t::create_zvals semantics (every tile eroded alone with droplets 0..N-1) through the latency modes, host and device pointers,
    per-tile min_zval; 130^2 tiles = the reference's default mesh_size 128."""
    import torch
    if mode != "auto":
        monkeypatch.setenv("TW_EROSION_MODE", mode)
    S, zv = 128, 130
    cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    hp, ep = cfg.height_params(), cfg.erosion_params()
    origins = [(tx * S * 3 - 2000, ty * S * 3 + 300) for ty in range(4) for tx in range(5)]
    tiles = ctx.heightgen_tiles(origins, cfg.mesh_size, float(cfg.dx_val), float(cfg.dy_val), zv, hp)
    exp = np.stack([oracle.apply_erosion(t, ep.zmin, 1000, convert(ep, oracle.ErosionParams))[0] for t in tiles])
    moves = sum(oracle.apply_erosion(t, ep.zmin, 1000, convert(ep, oracle.ErosionParams))[1] for t in tiles)
    assert beq(ctx.erode_tiles(tiles.copy(), 1000, ep, min_zval_all=ep.zmin), exp) == 0
    assert ctx.last_erosion_steps == moves
    dev = torch.from_numpy(tiles).cuda()
    ctx.erode_tiles(dev, 1000, ep, min_zval_all=ep.zmin)
    assert beq(dev.cpu().numpy(), exp) == 0
    mz = np.linspace(ep.zmin, ep.zmin + 0.5, len(origins)).astype(np.float32)
    exp = np.stack([oracle.apply_erosion(t, float(mz[i]), 80, convert(ep, oracle.ErosionParams))[0] for i, t in enumerate(tiles)])
    assert beq(ctx.erode_tiles(tiles.copy(), 80, ep, min_zvals=mz), exp) == 0
    fused = ctx.create_zvals_batch(origins, cfg.mesh_size, float(cfg.dx_val), float(cfg.dy_val), zv, hp, 1000, ep, ep.zmin)
    exp = np.stack([oracle.apply_erosion(t, ep.zmin, 1000, convert(ep, oracle.ErosionParams))[0] for t in tiles])
    assert beq(fused, exp) == 0


def _many_small_tiles(scene, ctx, side, mode=1):
    S, zv = 16, 18
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1), scene_size=(0.5, 0.5, 4.0))
    origins = [(tx * S * 40 - 3000, ty * S * 40 + 500) for ty in range(side) for tx in range(side)]
    return cfg, origins, zv


@pytest.mark.parametrize("env", [dict(TW_EROSION_WHOLE_MAX="0", TW_EROSION_WINDOW_ALL="0", TW_EROSION_HEAVY="300"),                      # heavy window launch + light global launch, forked streams
                                 dict(TW_EROSION_WHOLE_MAX="0", TW_EROSION_WINDOW_ALL="0", TW_EROSION_HEAVY="1000", TW_EROSION_LANES="8"),
                                 dict(TW_EROSION_WHOLE_MAX="0", TW_EROSION_WINDOW_ALL="100000"),                                         # everything through windows, scheduled
                                 dict(TW_EROSION_WHOLE_MAX="100000"),                                                                     # everything whole-map
                                 dict(TW_EROSION_MODE="global", TW_EROSION_LANES="16")])
def test_batch_split_heavy_light(tw, scene, oracle, ctx, beq, monkeypatch, env):
    """4900 tiles: the heaviest-first schedule, the window launch for its first slots on the forked high-priority stream and the global launch
    for the rest must together erode every tile exactly once, in the serial droplet order of each tile."""
    cfg, origins, zv = _many_small_tiles(scene, ctx, 70)
    hp, ep = cfg.height_params(), cfg.erosion_params()
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    raw = ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp)
    monkeypatch.setenv("TW_EROSION_MODE", "global")
    monkeypatch.setenv("TW_EROSION_LANES", "32")
    base = ctx.erode_tiles(raw.copy(), 60, ep, min_zval_all=ep.zmin)            # the round-1 path (tested against the oracle in test_gpu_erosion.py)
    steps = ctx.last_erosion_steps
    monkeypatch.delenv("TW_EROSION_MODE"); monkeypatch.delenv("TW_EROSION_LANES")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = ctx.erode_tiles(raw.copy(), 60, ep, min_zval_all=ep.zmin)
    assert beq(got, base) == 0
    assert ctx.last_erosion_steps == steps
    rng = np.random.default_rng(5)
    for t in rng.choice(len(origins), 24, replace=False):
        zc, _ = oracle.apply_erosion(raw[t], ep.zmin, 60, convert(ep, oracle.ErosionParams))
        assert beq(got[t], zc) == 0


@pytest.mark.parametrize("chunks", [2, 3, 5])
def test_fused_pipeline_multi_chunk(tw, scene, oracle, ctx, beq, monkeypatch, chunks):
    """tw_create_zvals_batch with several chunks: generation of chunk k+1 on the main stream overlaps the erosion of chunk k on the two
    auxiliary streams, each with its own scratch half and heavy-stream fork (the path BASELINE config 5 runs at N = 1)."""
    cfg, origins, zv = _many_small_tiles(scene, ctx, 70, mode=4)
    hp, ep = cfg.height_params(), cfg.erosion_params()
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    raw = ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp)
    one, mm1 = ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 60, ep, ep.zmin, want_minmax=True)
    steps = ctx.last_erosion_steps
    monkeypatch.setenv("TW_PIPE_CHUNKS", str(chunks))
    monkeypatch.setenv("TW_EROSION_WHOLE_MAX", "0")
    monkeypatch.setenv("TW_EROSION_WINDOW_ALL", "0")
    monkeypatch.setenv("TW_EROSION_HEAVY", "200")
    got, mm = ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 60, ep, ep.zmin, want_minmax=True)
    assert beq(got, one) == 0 and np.array_equal(mm, mm1)
    assert ctx.last_erosion_steps == steps
    rng = np.random.default_rng(chunks)
    for t in rng.choice(len(origins), 16, replace=False):
        zc, _ = oracle.apply_erosion(raw[t], ep.zmin, 60, convert(ep, oracle.ErosionParams))
        assert beq(got[t], zc) == 0


def test_config5_shape_sampled_against_oracle(tw, scene, oracle, ctx, beq):
    """The real BASELINE config-5 per-rank shape at N = 2: 32768 tiles of 258^2, 8-octave domain warp + 1000 droplets per tile through the
    fused call with its default settings (two chunks of 16384, G = 16 global launch + window launch for the heaviest tiles); 32 sampled
    tiles - the 8 heaviest by move count among them - are compared with the oracle."""
    import torch
    cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(256, 256, 1))
    hp, ep = cfg.height_params(), cfg.erosion_params()
    side_x, side_y, zv = 256, 128, 258
    origins = [((t % side_x) * 256, (t // side_x) * 256) for t in range(side_x * side_y)]
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    raw = torch.empty((len(origins), zv, zv), dtype=torch.float32, device="cuda")
    ctx.heightgen_tiles(origins, cfg.mesh_size, dx, dy, zv, hp, out=raw)
    out = torch.empty_like(raw)
    ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 1000, ep, ep.zmin, out=out)
    land = (raw > (ep.water_plane_z - ep.half_dxy)).sum(dim=(1, 2))                  # the schedule's work estimate: cells above the ocean-stop level
    heavy = torch.argsort(land, descending=True)[:8].cpu().numpy()
    rng = np.random.default_rng(11)
    sample = np.concatenate([heavy, rng.choice(len(origins), 24, replace=False)])
    for t in sample:
        zc, _ = oracle.apply_erosion(raw[int(t)].cpu().numpy(), ep.zmin, 1000, convert(ep, oracle.ErosionParams))
        assert beq(out[int(t)].cpu().numpy(), zc) == 0, int(t)
    changed = (out != raw).sum().item()
    assert changed > 100000


@pytest.mark.parametrize("mode", ["auto", "global", "window", "whole"])
def test_golden_erosion_fixture_on_gpu(tw, ctx, beq, monkeypatch, mode):
    """tests/golden/erosion.npz = outputs of the unmodified reference apply_erosion (1 thread): the synthetic 96^2 cases and the BASELINE
    config-1 fixture mapx/mesh128.txt, straight against the GPU."""
    if mode != "auto":
        monkeypatch.setenv("TW_EROSION_MODE", mode)
    e = np.load(os.path.join(GOLD, "erosion.npz"))
    for key_in, keys in (("in0", ["0_%d" % i for i in range(3)]), ("mesh128_in", ["mesh128"])):
        for k in keys:
            a = e["args" + k] if k != "mesh128" else e["mesh128_args"]
            ep = tw.ErosionParams(*[float(v) for v in a[2:]])
            out = ctx.erode(e[key_in].copy(), float(a[0]), int(a[1]), ep)
            exp = e["out" + k] if k != "mesh128" else e["mesh128_out"]
            assert beq(out, exp) == 0, k


def test_golden_voxel_fixture_on_gpu(tw, ctx, beq):
    """tests/golden/voxel.npz = outputs of the reference's noise_gen_3d / GLM 3-D loops, straight against tw_voxel_fill."""
    v = np.load(os.path.join(GOLD, "voxel.npz"))
    geo = v["geom"]
    for mode in (0, 1, 2):
        for name, norm, zs in (("v%d" % mode, 1, 0.0), ("v%d_unclamped" % mode, 0, 0.01)):
            vp = tw.VoxelParams()
            vp.nx, vp.ny, vp.nz = 20, 12, 28
            for d in range(3):
                vp.lo_pos[d], vp.vsz[d], vp.offset[d] = geo[d], geo[3 + d], geo[6 + d]
            vp.mag = vp.freq = 1.0
            vp.gen_mode, vp.normalize_to_1, vp.rseed1, vp.rseed2, vp.octaves = mode, norm, 123, 456, 3
            vp.rx, vp.ry = (float(x) for x in v["v%d_rxry" % mode])
            vp.zscale = zs
            assert beq(ctx.voxel_fill(vp), v[name]) == 0, name


@pytest.mark.parametrize("n,m,iters,env", [
    (130, 97, 700, {}),                                                  # a small map: almost every droplet conflicts with an earlier one -> constant re-walking, short prefixes
    (300, 200, 3000, dict(TW_SPEC_WINDOW="256")),
    (512, 512, 5000, {}),
    (258, 258, 1000, dict(TW_SPEC_WINDOW="64")),
    (200, 260, 2000, dict(TW_SPEC_LOG="64", TW_SPEC_TILES="16")),       # logs / tile lists that overflow all the time: outsized droplets walked in place as the window's head
    (200, 260, 2000, dict(TW_SPEC_VIEWS="2")),                            # at most two views per droplet: long walks go in place
    (20, 30, 300, {}),                                                   # smaller than the private view
    (1000, 700, 20000, dict(TW_SPEC_WINDOW="4096")),
])
def test_speculative_serial_order_equals_serial(tw, scene, oracle, ctx, beq, monkeypatch, n, m, iters, env):
    """M_SPEC (droplets walked speculatively in parallel, committed in the reference's order) == the serial reference order, bit for bit,
    moves included - on maps small enough that the speculation machinery (conflicts, re-walks, in-place heads, view reloads with the droplet's own
    log laid back on top) is exercised constantly."""
    monkeypatch.setenv("TW_EROSION_MODE", "spec")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg, z = _terrain(scene, ctx, n, m)
    zmin, eps = _eparams(tw, z)
    for ep in eps:
        zc, steps = oracle.apply_erosion(z, zmin, iters, convert(ep, oracle.ErosionParams))
        got = ctx.erode(z.copy(), zmin, iters, ep)
        assert beq(got, zc) == 0, "%d cells differ" % beq(got, zc)
        assert ctx.last_erosion_steps == steps


def test_speculative_is_the_default_for_one_big_map(tw, scene, oracle, ctx, beq, monkeypatch):
    """No environment override: a 1024^2 map with >= 64 droplets takes the M_SPEC path (spec_eligible) and equals the serial order; the forced global mode agrees."""
    cfg, z = _terrain(scene, ctx, 1024, 1024)
    zmin, eps = _eparams(tw, z)
    ep = eps[1]
    zc, steps = oracle.apply_erosion(z, zmin, 20000, convert(ep, oracle.ErosionParams))
    assert beq(ctx.erode(z.copy(), zmin, 20000, ep), zc) == 0 and ctx.last_erosion_steps == steps
    monkeypatch.setenv("TW_EROSION_MODE", "global")
    assert beq(ctx.erode(z.copy(), zmin, 20000, ep), zc) == 0 and ctx.last_erosion_steps == steps
