"""CPU: the plain-C oracle reproduces the committed golden fixtures (outputs of the unmodified reference code) bit for bit."""
import ctypes as C
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def hp_from_args(O, args, hmap, seed=1):
    mode, shape, ff, gl = (int(v) for v in args[:4])
    hp = O.HeightParams()
    hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = mode, shape, O.compute_scale(1.0, ff), gl
    hp.mesh_scale, hp.mesh_scale_z_inv = 1.0, 1.0
    hp.dx_val_inv = hp.dy_val_inv = 16.0          # mesh 128, scene 4 => DX_VAL = 0.0625
    hp.mesh_height, hp.mesh_height_scale = np.float32(0.1) * np.float32(4.0), 1.0
    hp.zmax_est, hp.custom_glaciate_exp = 2.3, 0.0
    hp.rx, hp.ry = O.gen_rx_ry(seed, 0, mode)
    hp.hmap = O.HmapParams(*[float(v) for v in hmap])
    return hp


def test_kat_values_from_survey(oracle, beq):
    k = load("kat.npz")
    # eval_index(3,5) = 1.15601587 / -3.95350838 / 1.79742229, noise_gen_3d get_val = 1.25381982
    assert abs(float(k["eval_index_3_5_mode0"]) - 1.15601587) < 5e-8
    assert abs(float(k["eval_index_3_5_mode1"]) + 3.95350838) < 5e-7
    assert abs(float(k["eval_index_3_5_mode2"]) - 1.79742229) < 5e-7
    assert abs(float(k["noise3d_point"]) - 1.25381982) < 5e-7
    sp = oracle.gen_sine_params(np.float32(0.4), seed=0, mode=0, rng=oracle.Rng(1, 1))
    assert beq(sp, k["sine_params_fresh_process"]) == 0
    for mode in (0, 1, 2):
        hp = hp_from_args(oracle, [mode, 0, 0, 0], [1000.0, 0, 0, 0, 1000.0] + [0] * 9, seed=0)
        z = oracle.heightgen_2d(oracle.Grid2D(0, 0, 0.0625, 0.0625, 8, 8), hp, sp, 0, 0)
        assert np.float32(z[5, 3]) == k["eval_index_3_5_mode%d" % mode]
    rd = oracle.noise3d_gen_sines(123, 456, 1.0, 1.0)
    v = oracle.lib().to_noise3d_get_val_pt(rd.ctypes.data_as(C.c_void_p), oracle.sin_table().ctypes.data_as(C.c_void_p), 0.1, 0.2, 0.3)
    assert np.float32(v) == k["noise3d_point"]


def test_glm_noise(oracle, beq):
    g = load("glm_noise.npz")
    L = oracle.lib()
    for name in ("simplex2", "perlin2", "simplex3", "perlin3"):
        f = getattr(L, "to_" + name)
        got = np.array([f(*[float(v) for v in p]) for p in g[name + "_in"]], np.float32)
        assert beq(got, g[name + "_out"]) == 0, name


def test_host_tables(oracle, beq):
    t = load("host_tables.npz")
    assert beq(oracle.sin_table(), t["sin_table"]) == 0
    for i in range(4):
        mode, seed, idx, mx, my, sx, sy, mhs = t["sp%d_args" % i]
        sh = np.float32(0.1) * np.float32(4.0) * np.float32(mhs)
        sp = oracle.gen_sine_params(sh, mesh=(int(mx), int(my)), scene=(float(sx), float(sy)), seed=int(seed), rgen_index=int(idx), mode=int(mode))
        if int(seed) != 0 or int(mode) != 0:   # otherwise the reference's function-static rng state depends on call history
            assert beq(sp, t["sp%d" % i]) == 0
        assert beq(np.array(oracle.gen_rx_ry(int(seed), int(idx), int(mode)), np.float32), t["rxry%d" % i]) == 0
    assert beq(oracle.noise3d_gen_sines(123, 456, 1.0, 1.0), t["rdata_123_456"]) == 0
    assert beq(oracle.noise3d_gen_sines(7, 9, 2.5, 0.3), t["rdata_7_9"]) == 0


def test_height_grids(oracle, beq):
    h = load("height.npz")
    for mode in range(5):
        for shape in range(3):
            n = "h_m%d_s%d" % (mode, shape)
            a = h[n + "_args"]
            hp = hp_from_args(oracle, a, h[n + "_hmap"])
            g = oracle.Grid2D(a[4], a[5], 0.0625, 0.0625, int(a[6]), int(a[7]))
            z = oracle.heightgen_2d(g, hp, h[n + "_sp"] if mode == 0 else None, 1, 0)
            assert beq(z, h[n]) == 0, n
    hp = hp_from_args(oracle, [0, 0, 2, 1], [1000.0, 0, 0, 0, 1000.0] + [0] * 9)
    hp.mesh_height_scale, hp.zmax_est = 0.7, 0.5
    z = oracle.heightgen_2d(oracle.Grid2D(-64, -64, 0.0625, 0.0625, 128, 128), hp, h["cfg1_sp"], 1, 0)
    assert beq(z, h["cfg1"]) == 0


def point_cases(mod, h):
    """Yields (name, HeightParams, PointQuery, xy, sine_params, expected) for every query stored in points.npz (mod = oracle or product module)."""
    for mode in range(5):
        for shape in range(3):
            n = "p_m%d_s%d" % (mode, shape)
            hp = hp_from_args(mod, h[n + "_args"], h[n + "_hmap"])
            sp = h[n + "_sp"] if (n + "_sp") in h else None
            for qi in range(4):
                key = "%s_q%d" % (n, qi)
                if key + "_xy" not in h:
                    continue
                q = h[key + "_query"]
                pq = mod.PointQuery(int(q[0]), float(q[1]), int(q[2]), int(q[3]), float(q[4]), float(q[5]), int(q[6]), int(q[7]), int(q[8]))
                yield key, hp, pq, h[key + "_xy"], sp, h[key + "_out"]


def test_point_queries(oracle, beq):
    for key, hp, pq, xy, sp, exp in point_cases(oracle, load("points.npz")):
        assert beq(oracle.eval_points(xy, hp, pq, sp), exp) == 0, key


def tile_case(mod, h, mode):
    """(HeightParams, sine_params, args) of a tiles.npz case: mesh 64x64, scene 4 => DX_VAL 0.125, mesh_freq_filter 1, seed 1, HM_CFG, zmax_est 2.3."""
    n = "m%d" % mode
    hp = hp_from_args(mod, [mode, 0, 1, 1], [1000.0, 0, 0, 0, 1000.0, 0, 0, 0, 0, 5.0, 0.001, -4.0, 0, 0])
    hp.dx_val_inv = hp.dy_val_inv = 8.0
    return hp, (h["sp_" + n] if mode == 0 else None), h["args_" + n]


def test_tile_normals_and_ao(oracle, beq):
    h = load("tiles.npz")
    for mode in (0, 1, 4):
        n = "m%d" % mode
        hp, sp, a = tile_case(oracle, h, mode)
        x1, y1, S, zv, dx, dy, half_dxy = int(a[1]), int(a[2]), int(a[3]), int(a[4]), float(a[5]), float(a[6]), float(a[7])
        tile = oracle.heightgen_2d(oracle.Grid2D(x1 - S // 2, y1 - S // 2, dx, dy, zv, zv), hp, sp, 1, 0)
        assert beq(tile, h["tile_" + n]) == 0
        rgba, mnz = oracle.tile_normals(tile[None], dx, dy)
        assert np.array_equal(rgba[0], h["normals_" + n]) and mnz[0] == h["min_normal_z_" + n]
        csz = zv - 1 + 72
        context = oracle.heightgen_2d(oracle.Grid2D(x1 - 36 - S // 2, y1 - 36 - S // 2, dx, dy, csz, csz), hp, sp, 1, 0)
        assert np.array_equal(oracle.tile_ao(tile[None], context[None], half_dxy, use_ao_zvals=(mode >= 3))[0], h["ao_" + n])
        if mode >= 3:   # create_zvals + calc_mesh_ao_lighting in a GPU gen mode: zvals = interior of the context, erosion, AO against the un-eroded context
            cut = np.ascontiguousarray(context[36:36 + zv, 36:36 + zv])
            assert beq(cut, h["tile_" + n + "c"]) == 0
            e = h["ero_" + n]
            eroded, _ = oracle.apply_erosion(cut, float(e[0]), int(e[1]), oracle.ErosionParams(*[float(v) for v in e[2:]]))
            assert beq(eroded, h["tile_" + n + "e"]) == 0
            assert np.array_equal(oracle.tile_ao(cut[None], context[None], half_dxy, use_ao_zvals=True)[0], h["ao_" + n + "c"])
            assert np.array_equal(oracle.tile_ao(eroded[None], context[None], half_dxy, use_ao_zvals=True)[0], h["ao_" + n + "e"])


def hmap_cases(mod, h):
    f32 = np.float32
    for (ms, mfs, tz, x1, y1), exp in zip(h["hmap_cases"], h["hmap_out"]):
        img = h["hmap_img"]
        hs = mod.HmapSampler(img.shape[1], img.shape[0], 2, float(ms), float(f32(0.0008) * f32(1.5)), float(mfs), float(tz), 0.5)
        yield img, hs, (int(x1), int(y1)), exp


def test_heightmap_texture_tiles(oracle, beq):
    for img, hs, org, exp in hmap_cases(oracle, load("tiles.npz")):
        assert beq(oracle.hmap_sample_tiles(img, hs, [org], 34)[0], exp) == 0


def test_erosion(oracle, beq):
    e = load("erosion.npz")
    for key_in, keys in (("in0", ["0_%d" % i for i in range(3)]), ("mesh128_in", ["mesh128"])):
        for k in keys:
            a = e["args" + k] if k != "mesh128" else e["mesh128_args"]
            ep = oracle.ErosionParams(*[float(v) for v in a[2:]])
            out, steps = oracle.apply_erosion(e[key_in], float(a[0]), int(a[1]), ep)
            exp = e["out" + k] if k != "mesh128" else e["mesh128_out"]
            assert beq(out, exp) == 0, k
            assert steps > 0


def test_voxels(oracle, beq):
    v = load("voxel.npz")
    geo = v["geom"]
    for mode in (0, 1, 2):
        for name, norm, zs in (("v%d" % mode, 1, 0.0), ("v%d_unclamped" % mode, 0, 0.01)):
            vp = oracle.VoxelParams()
            vp.nx, vp.ny, vp.nz = 20, 12, 28
            for d in range(3):
                vp.lo_pos[d], vp.vsz[d], vp.offset[d] = geo[d], geo[3 + d], geo[6 + d]
            vp.mag = vp.freq = 1.0
            vp.gen_mode, vp.normalize_to_1, vp.rseed1, vp.rseed2, vp.octaves = mode, norm, 123, 456, 3
            vp.rx, vp.ry = (float(x) for x in v["v%d_rxry" % mode])
            vp.zscale = zs
            assert beq(oracle.voxel_fill(vp), v[name]) == 0, name


def test_u16_pack_roundtrip(oracle):
    rng = np.random.default_rng(3)
    vals = rng.uniform(10.0, 20.0, 1000).astype(np.float32)
    mult, add = np.float32(10.0 / 255.0), np.float32(10.0)
    packed, bad = oracle.from_floats_u16(vals, float(mult), float(add))
    assert bad == 0
    back = oracle.to_floats_u16(packed, float(mult), float(add))
    assert np.abs(back - vals).max() <= float(mult) / 256.0 * 1.01 + 1e-6
    _, bad = oracle.from_floats_u16(np.array([0.0], np.float32), 1.0, 1.0)   # v = -1: out of range
    assert bad == 1


def _gen_mesh_case(O, args):
    mode, seed, ff, iters, mhs, smag, sfreq, sbias = args
    hp = O.HeightParams()
    hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = int(mode), 0, O.compute_scale(1.0, int(ff)), 1
    hp.mesh_scale = hp.mesh_scale_z_inv = 1.0
    hp.dx_val_inv = hp.dy_val_inv = 16.0
    hp.mesh_height, hp.mesh_height_scale, hp.zmax_est = np.float32(0.1) * np.float32(4.0), float(mhs), 1.0
    hp.rx, hp.ry = O.gen_rx_ry(int(seed), 0, int(mode))
    hp.hmap = O.hmap_params(sine_mag=float(smag), sine_freq=float(sfreq), sine_bias=float(sbias))
    return hp, int(seed), int(iters)


def test_gen_mesh_ground_mode(oracle, beq):
    """BASELINE config 1 (128x128 sine mesh, seed 6, glaciate, freq filter 2, mesh_height 0.7) through the whole gen_mesh() flow:
    sine table, mesh fill, estimate_zminmax probe, set_zvals, glaciate(), apply_erosion - against the reference's own gen_mesh()."""
    h = load("height.npz")
    for name in ("gm_cfg1", "gm_cfg1_eroded", "gm_simplex"):
        hp, seed, iters = _gen_mesh_case(oracle, h[name + "_args"])
        mesh, z6, _ = oracle.gen_mesh(hp, seed=seed, erosion_iters=iters, ep=oracle.ErosionParams(1.0, 0.0, 0.0625, 0.0, 0.0, 0.0, 0.5))
        assert beq(mesh, h[name]) == 0, name
        assert beq(np.array([z6[k] for k in ("zmin", "zmax", "zmax_est", "zbottom", "ztop", "water_plane_z")], np.float32), h[name + "_zvals"]) == 0


def test_tile_bounds_small_case(oracle):
    z = np.arange(36, dtype=np.float32).reshape(1, 6, 6)           # zvsize 6 -> block_size 1: sub-block (yy,xx) covers rows yy..yy+1, cols xx..xx+1
    b = oracle.tile_bounds(z, 7.5, 0.0625, 0.0625, 4)[0]
    assert list(b.sub_zmin)[:4] == [0.0, 1.0, 2.0, 3.0] and list(b.sub_zmax)[:4] == [7.0, 8.0, 9.0, 10.0]
    assert (b.mzmin, b.mzmax, b.mesh_dz) == (0.0, 28.0, 7.0)     # cells beyond index 4 are never visited (last row/column is not rendered)
    assert (b.wx1, b.wy1, b.wx2, b.wy2) == (0, 0, 4, 1)            # z < 7.5: rows 0 (cols 0..4) and row 1 (cols 0,1)


def test_voxel_post_processing(oracle, beq):
    """tests/golden/voxel_post.npz = outputs of the reference's own voxel_manager functions (N3): flags, flood fills, triangle soup."""
    g = load("voxel_post.npz")
    tables = (g["edge_table"], g["tri_table"], g["edge_to_vals"])
    for name in ("sine", "inv", "mesh"):
        a = g[name + "_params"]
        p = oracle.VoxelPostParams()
        p.nx, p.ny, p.nz = int(a[0]), int(a[1]), int(a[2])
        for d in range(3):
            p.lo_pos[d], p.vsz[d] = float(a[3 + d]), float(a[6 + d])
        p.isolevel, p.invert, p.make_closed_surface, p.remove_unconnected, p.keep_at_edge, p.centre_seed, p.skip_under_mesh = float(a[9]), int(a[10]), int(a[11]), int(a[12]), int(a[13]), int(a[14]), int(a[15])
        zix = g[name + "_zix"] if (name + "_zix") in g.files else None
        out = oracle.voxel_outside(g[name + "_vals"], p, zix)
        assert np.array_equal(out, g[name + "_outside"])
        v2, o2, _ = oracle.voxel_remove_unconnected(g[name + "_vals"], out, p)
        assert np.array_equal(o2, g[name + "_outside2"]) and beq(v2, g[name + "_vals2"]) == 0
        assert beq(oracle.voxel_triangles(v2, o2, p, tables), g[name + "_tris"]) == 0


def shadow_params(P, a, lp):
    """tests/golden/shadows.npz 'params' = X/Y_SCENE_SIZE, DX/DY_VAL, XY_SUM_SIZE, zmin, zmax."""
    sp = P()
    sp.x_scene_size, sp.y_scene_size, sp.dx_val, sp.dy_val = float(a[0]), float(a[1]), float(a[2]), float(a[3])
    sp.dx_val_inv, sp.dy_val_inv = 1.0 / np.float32(a[2]), 1.0 / np.float32(a[3])
    sp.xy_sum_size, sp.zmin, sp.zmax, sp.no_shadow = int(a[4]), float(a[5]), float(a[6]), 0
    for d in range(3):
        sp.lpos[d] = float(lp[d])
    return sp


def test_mesh_shadows(oracle, beq):
    """tests/golden/shadows.npz = the reference's own calc_mesh_shadows over a 3x3 block of tiles chained as tile_t::calc_shadows_for_light chains them (N4)."""
    g = load("shadows.npz")
    txy = [tuple(int(v) for v in t) for t in g["tile_xy"]]
    for li, lp in enumerate(g["lights"]):
        m, ox, oy = oracle.tile_shadows_batch(g["tiles"], txy, shadow_params(oracle.ShadowParams, g["params"], lp))
        assert np.array_equal(m, g["smask_%d" % li]), li
        assert beq(ox, g["sh_out_x_%d" % li]) == 0 and beq(oy, g["sh_out_y_%d" % li]) == 0, li


def weights_golden_case(HP, WP, hmap_params, g, n, ci):
    """HeightParams / WeightParams of one case of tests/golden/weights.npz (the scene tests/golden/make_golden_weights.py set up in the reference)."""
    mode, S, dx, dy, zmin, zmax, water, start, mesh_height = [float(v) for v in g["scal_" + n]]
    c = g["case_%s_%d" % (n, ci)]
    hp = HP()
    hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = int(mode), 0, int(start), 1
    hp.mesh_scale, hp.mesh_scale_z_inv = 1.0, 1.0
    hp.dx_val_inv, hp.dy_val_inv = 1.0 / np.float32(dx), 1.0 / np.float32(dy)
    hp.mesh_height, hp.mesh_height_scale, hp.zmax_est, hp.custom_glaciate_exp = mesh_height, 1.0, 2.3, 0.0
    hp.hmap = hmap_params
    wp = WP()
    for i in range(5):
        wp.h_dirt[i], wp.tex_class[i] = float(c[i]), int(c[5 + i])
    wp.sthresh[0][0], wp.sthresh[0][1], wp.sthresh[1][0], wp.sthresh[1][1] = 0.68, 0.86, 0.48, 0.72
    wp.zmin, wp.zmax, wp.relh_adj_tex, wp.vegetation, wp.snow_to_rock = zmin, zmax, float(c[10]), float(c[11]), int(c[12])
    wp.water_level = water
    wp.noise_scale = np.float32(1.0 * float(np.float32(0.003)) * 1.0)
    wp.vnz_scale = float(np.float32(np.sqrt(2.0))) if int(mode) == 4 else 1.0          # SQRT2 = sqrt(2.0) as a float (src/3DWorld.h:132)
    wp.dx_val, wp.dy_val, wp.dxdy = dx, dy, float(np.float32(dx) * np.float32(dy))
    wp.xy_mult = np.float32(1.0 / float(np.float32(S)))
    return hp, wp, int(S), dx, dy


def test_terrain_weights_texture(oracle):
    """tests/golden/weights.npz = the reference's own tile_t::create_texture (terrain part) on reference tiles, gen modes 1 and 4, three parameter sets (N4)."""
    from cases import HM_CFG
    g = load("weights.npz")
    for n in ("m1", "m4"):
        for ci in range(3):
            hp, wp, S, dx, dy = weights_golden_case(oracle.HeightParams, oracle.WeightParams, oracle.hmap_params(**HM_CFG), g, n, ci)
            rand = oracle.weights_noise(hp, g["sine_params_" + n], [tuple(int(v) for v in o) for o in g["origins_" + n]], (S, S), dx, dy, S + 1)
            w, flags = oracle.tile_weights(g["tiles_" + n], rand, g["corners_" + n], wp)
            assert np.array_equal(w, g["weights_%s_%d" % (n, ci)]), (n, ci, int((w != g["weights_%s_%d" % (n, ci)]).sum()))
            assert np.array_equal(flags, g["grass_%s_%d" % (n, ci)])


def test_spec_protocol_model(oracle, beq):
    """The protocol of the product's speculative serial-order erosion (M_SPEC) as a sequential model in the oracle: whatever the window, the moves
    allowed per round, the conflict-tile size and the order in which the walkers (and the in-place head among them) run, the result is the serial reference order's, bit
    for bit, moves included - on maps so small that almost every droplet conflicts with an earlier one."""
    g = load("erosion.npz")
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:57, 0:83].astype(np.float32)
    z = (np.sin(xx * 0.21) * np.cos(yy * 0.17) * 0.8 + 0.3 * np.sin(xx * 0.05 + yy * 0.09) + 0.05 * rng.standard_normal((57, 83))).astype(np.float32)
    zmin, zmax = float(z.min()), float(z.max())
    total_wasted = total_inplace = 0
    for ep in (oracle.ErosionParams(1.0, zmin - 10, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 0.5), oracle.ErosionParams(1.0, zmin + 0.2 * (zmax - zmin), 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 2.0)):
        want, steps = oracle.apply_erosion(z, zmin, 400, ep)
        for window, cap, shift, seed in ((1, 1000000, 2, 0), (7, 3, 2, 1), (32, 1, 2, 2), (32, 16, 3, 3), (64, 64, 2, 4), (256, 8, 4, 5), (16, 1000000, 2, 6), (500, 5, 2, 7)):
            got, moves, (rounds, walks, wasted, inplace) = oracle.erode_spec_model(z, zmin, 400, ep, window, cap, shift, seed)
            assert moves == steps and beq(got, want) == 0, (window, cap, shift, seed, moves, steps)
            assert walks == 400 + wasted and rounds >= 1
            total_wasted, total_inplace = total_wasted + wasted, total_inplace + inplace
    assert total_wasted > 200 and total_inplace > 200            # conflicts and in-place heads did occur
