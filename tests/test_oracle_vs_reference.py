"""CPU: the plain-C oracle vs the UNMODIFIED reference objects (oracle/_ref), bit for bit, on randomised inputs.
This is what pins the oracle (the reference has no tests of its own for this path). Skipped when oracle/_ref is not built."""
import numpy as np

from cases import HM_ALL, HM_CFG


def _hp(O, R, mode, shape, ff, seed, hmap, zmax_est, gl, custom=0.0, ms=1.0):
    R.setup(mode=mode, shape=shape, freq_filter=ff, seed=seed, glaciate=gl, custom_glaciate_exp=custom, hmap=hmap, zmax_est=zmax_est, mesh_scale=ms)
    RL = R.lib()
    hp = O.HeightParams()
    hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = mode, shape, O.compute_scale(ms, ff), gl
    assert hp.start_eval_sin == RL.ref_get_start_eval_sin()
    hp.mesh_scale, hp.mesh_scale_z_inv = ms, 1.0
    hp.dx_val_inv, hp.dy_val_inv = 1.0 / np.float32(RL.ref_get_dx()), 1.0 / np.float32(RL.ref_get_dy())
    hp.mesh_height, hp.mesh_height_scale = RL.ref_get_mesh_height(), 1.0
    hp.zmax_est, hp.custom_glaciate_exp = zmax_est, custom
    hp.rx, hp.ry = O.gen_rx_ry(seed, 0, mode)
    assert (hp.rx, hp.ry) == R.rx_ry()
    hp.hmap = O.hmap_params(**(hmap or {}))
    return hp


def test_glm_noise_random_points(oracle, ref, beq):
    rng = np.random.default_rng(0)
    for name, k in (("simplex2", 2), ("perlin2", 2), ("simplex3", 3), ("perlin3", 3)):
        pts = (rng.standard_normal((4000, k)) * rng.choice([0.5, 3, 50, 1000, 1e5], (4000, 1))).astype(np.float32)
        pts[:500] = np.round(pts[:500])
        fr, fo = getattr(ref.lib(), "ref_glm_" + name), getattr(oracle.lib(), "to_" + name)
        a = np.array([fr(*[float(v) for v in p]) for p in pts], np.float32)
        b = np.array([fo(*[float(v) for v in p]) for p in pts], np.float32)
        assert beq(a, b) == 0, name


def test_heightgen_all_modes(oracle, ref, beq):
    RL = ref.lib()
    for mode in (0, 1, 2, 3, 4):
        for shape in (0, 1, 2):
            for ff, hmap, gl, custom, ms in ((1, HM_CFG, 1, 0.0, 1.0), (0, None, 0, 0.0, 1.0), (2, HM_ALL, 1, 0.0, 1.0), (1, HM_CFG, 1, 2.5, 1.0), (1, HM_ALL, 1, 0.0, 4.0)):
                hp = _hp(oracle, ref, mode, shape, ff, 1, hmap, 2.3, gl, custom, ms)
                sp = ref.sine_params()
                n = 48 if mode == 4 else 80
                for x0, y0, dxm in ((-n / 2, -n / 2, 1.0), (-50000.0, 70000.0, 3.0)):
                    dx, dy = RL.ref_get_dx() * dxm, RL.ref_get_dy() * dxm
                    zr = ref.heightgen(x0, y0, dx, dy, n, n - 7, cache_values=0, glaciate=1)
                    zo = oracle.heightgen_2d(oracle.Grid2D(x0, y0, dx, dy, n, n - 7), hp, sp, 1, 0)
                    assert beq(zr, zo) == 0, (mode, shape, ff, custom, ms, x0)


def _points(rng, n, span):
    xy = (rng.uniform(-span, span, (n, 2))).astype(np.float32)
    xy[: n // 8] = np.round(xy[: n // 8])                      # lattice points / exact cell centres
    xy[n // 8: n // 6] *= 1000.0                               # far away (large sine-table arguments)
    return xy


def test_point_queries(oracle, ref, beq):
    """eval_mesh_sin_terms, eval_mesh_sin_terms_scaled and get_exact_zval (procedural branch) of the reference vs the
    oracle's to_eval_points, every gen mode / shape, scrolled and unscrolled."""
    rng = np.random.default_rng(5)
    for mode in (0, 1, 2, 3, 4):
        for shape in (0, 1, 2):
            for ff, hmap, gl, custom, ms in ((1, HM_CFG, 1, 0.0, 1.0), (0, None, 0, 0.0, 1.0), (2, HM_ALL, 1, 0.0, 1.0), (1, HM_ALL, 1, 2.5, 4.0)):
                hp = _hp(oracle, ref, mode, shape, ff, 1, hmap, 2.3, gl, custom, ms)
                sp = ref.sine_params()
                n = 300 if mode == 4 else 1000
                for kind, span, xy_scale, no_xyoff, xo, yo in ((0, 30.0, 1.0, 0, 0, 0), (1, 500.0, 1.0, 0, 0, 0), (1, 500.0, 16.0, 0, 0, 0),
                                                              (2, 4.0, 1.0, 0, 0, 0), (2, 4.0, 1.0, 0, 640, -1280), (2, 40.0, 1.0, 1, 640, -1280)):
                    xy = _points(rng, n, span)
                    pq = oracle.PointQuery(kind, xy_scale, 128, 128, 4.0, 4.0, xo, yo, no_xyoff)
                    zr = ref.eval_points(kind, xy, xy_scale, no_xyoff, xo, yo)
                    zo = oracle.eval_points(xy, hp, pq, sp)
                    assert beq(zr, zo) == 0, (mode, shape, ff, kind, xy_scale, no_xyoff)


def test_tile_normals(oracle, ref, beq):
    """the normal-map arithmetic (get_norm via the reference's vector3d, byte quantisation, min_normal_z) vs the oracle."""
    rng = np.random.default_rng(11)
    ref.setup(mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
    RL = ref.lib()
    tiles = [ref.heightgen(100.0, -300.0, RL.ref_get_dx(), RL.ref_get_dy(), 66, 66, 0, 1),
             np.zeros((10, 10), np.float32),                                                # flat: normal (0, 0, 1) -> 127, 127, 254
             (rng.standard_normal((34, 34)) * 50).astype(np.float32),                       # near-vertical faces
             (rng.standard_normal((18, 18)) * 1e-4).astype(np.float32)]
    for dxv, dyv in ((0.0625, 0.0625), (0.25, 0.03125)):
        for t in tiles:
            rr, rm = ref.tile_normals(t, dxv, dyv)
            orr, om = oracle.tile_normals(t[None], dxv, dyv)
            assert np.array_equal(rr, orr[0]) and np.float32(rm) == om[0]
    assert np.array_equal(oracle.tile_normals(tiles[1][None], 0.0625, 0.0625)[0][0, 0, 0], [127, 127, 254, 0])


def test_heightmap_texture_tiles(oracle, ref, beq):
    """terrain_hmap_manager_t::get_clamped_height (nearest texel, bilinear for mesh_scale < 1, mirror edges) of the linked
    reference vs the oracle, tiles inside, across and far outside a 16-bit image (odd sizes included)."""
    rng = np.random.default_rng(3)
    ref.setup(mode=0, freq_filter=1, seed=1, mesh_height_scale=1.5, mesh_scale_z=2.0)
    f32 = np.float32
    for w, h in ((64, 48), (33, 71)):
        img = rng.integers(0, 256, (h, w, 2), dtype=np.uint8)
        for ms, mfs, tz in ((1.0, 1.0, 0.0), (0.5, 2.5, -0.75), (0.37, 1.0, 0.1), (2.0, 0.8, 0.0), (3.3, 1.0, 1.0)):
            hs = oracle.HmapSampler(w, h, 2, ms, float(f32(0.0008) * f32(1.5)), mfs, tz, 0.5)
            for x1, y1 in ((-10, -9), (-w // 2 - 5, h // 2 - 7), (3 * w + 1, -5 * h - 2), (-1000, 999)):
                zr = ref.hmap_sample_tile(img, x1, y1, 18, ms, mfs, tz)
                zo = oracle.hmap_sample_tiles(img, hs, [(x1, y1)], 18)[0]
                assert beq(zr, zo) == 0, (w, h, ms, x1, y1)


def test_erosion_serial_order(oracle, ref, beq):
    ref.lib().ref_set_threads(1)
    ref.setup(mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
    RL = ref.lib()
    for n, m, iters in ((130, 130, 1000), (64, 200, 500), (258, 258, 600)):
        z = ref.heightgen(-n / 2, -m / 2, RL.ref_get_dx(), RL.ref_get_dy(), n, m, 0, 1)
        zmin, zmax = float(z.min()), float(z.max())
        for wpz, clip, ea in ((zmin - 10, 0.5, 1.0), ((zmin + zmax) / 2, 0.3, 1.0), (zmin - 10, -1.0, 0.5)):
            zr = ref.apply_erosion(z, zmin, iters, erode_amount=ea, water_plane_z=wpz, zmin=zmin - 0.1, zmax=zmax + 0.1, clip_hd1=clip)
            zo, _ = oracle.apply_erosion(z, zmin, iters, oracle.ErosionParams(ea, wpz, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, clip))
            assert beq(zr, zo) == 0
    ref.lib().ref_set_threads(8)


def test_erosion_openmp_mode_is_order_dependent(oracle, ref, beq):
    """The reference's own droplet loop is `#pragma omp parallel for schedule(dynamic,1)` (src/erosion.cpp:66) with unsynchronised
    read-modify-writes: with more than one thread its output depends on timing. This pins what "parity" can mean for that mode (and for
    tw_erode_parallel, its GPU counterpart): droplet for droplet only at one thread; otherwise the same amount of material moved."""
    ref.setup(mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
    RL = ref.lib()
    n, iters = 192, 30000
    z = ref.heightgen(-n / 2, -n / 2, RL.ref_get_dx(), RL.ref_get_dy(), n, n, 0, 1)
    zmin, zmax = float(z.min()), float(z.max())
    kw = dict(erode_amount=1.0, water_plane_z=zmin - 10, zmin=zmin - 0.1, zmax=zmax + 0.1, clip_hd1=0.083)
    RL.ref_set_threads(1)
    z1 = ref.apply_erosion(z, zmin, iters, **kw)
    zo, _ = oracle.apply_erosion(z, zmin, iters, oracle.ErosionParams(1.0, zmin - 10, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 0.083))
    assert beq(z1, zo) == 0
    RL.ref_set_threads(8)
    z8 = ref.apply_erosion(z, zmin, iters, **kw)
    moved1, moved8 = np.abs(z1.astype(np.float64) - z).sum(), np.abs(z8.astype(np.float64) - z).sum()
    print("reference, 8 threads vs 1: identical cells %.4f, max diff %.3g, moved %.6g vs %.6g" % ((z8 == z1).mean(), np.abs(z8 - z1).max(), moved8, moved1))
    assert np.isfinite(z8).all()
    assert abs(moved8 - moved1) < 0.3 * moved1          # run-to-run spread seen here: up to ~8 %
    if RL.ref_get_max_threads() > 1 and __import__("os").cpu_count() > 1:
        assert (z8 != z1).any()        # 30000 droplets on 192^2 cells: concurrent droplets meet, and the result shows it


def test_voxel_fill(oracle, ref, beq):
    lo, vsz, off = (-7.9, -7.8, -1.5), (0.4, 0.65, 0.11), (0.5, -0.25, 0.0)
    for mode in (0, 1, 2):
        for ff in (2, 0):
            ref.setup(mode=mode, freq_filter=ff, seed=3)
            rx, ry = ref.rx_ry()
            for norm, zs in ((1, 0.0), (0, 0.01)):
                zr = ref.voxel_fill(24, 10, 30, lo, vsz, off, 1.0, 1.0, norm, 123, 456, mode, zs)
                vp = oracle.VoxelParams()
                vp.nx, vp.ny, vp.nz = 24, 10, 30
                for d in range(3):
                    vp.lo_pos[d], vp.vsz[d], vp.offset[d] = lo[d], vsz[d], off[d]
                vp.mag = vp.freq = 1.0
                vp.gen_mode, vp.normalize_to_1, vp.rseed1, vp.rseed2, vp.octaves = mode, norm, 123, 456, max(1, 5 - ff)
                vp.rx, vp.ry, vp.zscale = rx, ry, zs
                assert beq(zr, oracle.voxel_fill(vp)) == 0


def test_gen_mesh_ground_mode(oracle, ref, beq):
    """The reference's own gen_mesh(0,0,1) (linked unmodified) vs the oracle's restatement of the whole ground-mode flow."""
    ref.lib().ref_set_threads(1)
    for mode, seed, ff, hmap, iters, mhs in ((0, 6, 2, {}, 0, 0.7), (0, 6, 2, HM_CFG, 1500, 0.7), (1, 3, 1, HM_CFG, 400, 1.0), (2, 5, 1, {}, 300, 1.0)):   # GPU gen modes 3/4 need a GL shader inside the reference's gen_mesh
        ref.setup(mode=mode, freq_filter=ff, seed=seed, glaciate=1, mesh_height_scale=mhs, hmap=hmap, gen_sine_table=False)
        zr, z6r = ref.gen_mesh((128, 128), erosion_iters=iters)
        hp = oracle.HeightParams()
        hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = mode, 0, oracle.compute_scale(1.0, ff), 1
        hp.mesh_scale = hp.mesh_scale_z_inv = 1.0
        hp.dx_val_inv = hp.dy_val_inv = 16.0
        hp.mesh_height, hp.mesh_height_scale, hp.zmax_est = np.float32(0.1) * np.float32(4.0), mhs, 1.0
        hp.rx, hp.ry = oracle.gen_rx_ry(seed, 0, mode)
        hp.hmap = oracle.hmap_params(**hmap)
        zo, z6o, _ = oracle.gen_mesh(hp, seed=seed, erosion_iters=iters, ep=oracle.ErosionParams(1.0, 0.0, 0.0625, 0.0, 0.0, 0.0, 0.5))
        assert beq(zr, zo) == 0, (mode, seed)
        assert all(np.float32(z6r[k]) == np.float32(z6o[k]) for k in z6r), (z6r, z6o)
    ref.lib().ref_set_threads(8)


def _need_extract(ref):
    import pytest
    if not ref.has_tiled_extract():
        pytest.skip("oracle/_ref was built without the tiled_mesh.cpp extraction")


def test_tile_create_zvals_extracted_reference(oracle, ref, beq):
    """SURVEY 8a row a11 pinned against the reference's OWN tile_t::create_zvals(): the function body is cut out of src/tiled_mesh.cpp:467-546
    at build time (oracle/refbuild/build_ref.sh) and compiled unmodified together with setup_height_gen_async/get_xy_scale; it runs
    build_arrays + eval_index + apply_erosion + the 4x4 sub-block / water-bbox / radius tail. The port = heightgen_2d + apply_erosion +
    tile_bounds. CPU gen modes 0-2 (modes 3/4 need GL inside build_arrays)."""
    _need_extract(ref)
    RL = ref.lib()
    RL.ref_set_threads(1)
    for mode, size, mesh in ((1, 64, (64, 64, 1)), (2, 32, (32, 32, 1)), (0, 48, (48, 48, 1)), (1, 128, (128, 128, 1))):
        for x1, y1, iters in ((0, 0, 0), (5 * size, -3 * size, 300), (-40 * size, 17 * size, 120)):
            ref.setup(mesh=mesh, mode=mode, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
            dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
            hp = oracle.HeightParams()
            hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = mode, 0, oracle.compute_scale(1.0, 1), 1
            hp.mesh_scale = hp.mesh_scale_z_inv = hp.mesh_height_scale = 1.0
            hp.dx_val_inv, hp.dy_val_inv = 1.0 / np.float32(dx), 1.0 / np.float32(dy)
            hp.mesh_height, hp.zmax_est = RL.ref_get_mesh_height(), 2.0
            hp.rx, hp.ry = oracle.gen_rx_ry(1, 0, mode)
            hp.hmap = oracle.hmap_params(**HM_CFG)
            sp = ref.sine_params() if mode == 0 else None
            zv = size + 2
            raw = oracle.heightgen_2d(oracle.Grid2D(float(x1 - mesh[0] // 2), float(y1 - mesh[1] // 2), dx, dy, zv, zv), hp, sp, 1, 0)
            lo, hi = float(raw.min()), float(raw.max())
            wpz = RL.ref_get_water_z_height()
            # erosion's own water level (droplets stop below it): under the tile / inside its height range; zmin = min_zval of the tile erosion, zmax, clip_hd1 of get_bare_ls_tid
            ep = ref.Erosion(1.0, (lo - 10.0) if x1 > 0 else (lo + 0.3 * (hi - lo)), 0.5 * (dx + dy), lo - 0.05, hi + 0.3, 0.0, 0.4)
            zr, br = ref.tile_create_zvals(size, x1, y1, iters, ep)
            zo = raw
            if iters:
                zo, _ = oracle.apply_erosion(raw, ep.zmin, iters, oracle.ErosionParams(*[getattr(ep, f) for f, _ in ep._fields_]))
                assert (zo != raw).any()
            assert beq(zr, zo) == 0, (mode, size, x1, y1, iters)
            bo = oracle.tile_bounds(zo[None], wpz, dx, dy, size)[0]
            assert np.array_equal(np.array(bo.sub_zmin, np.float32).reshape(4, 4), br["sub_zmin"]) and np.array_equal(np.array(bo.sub_zmax, np.float32).reshape(4, 4), br["sub_zmax"])
            assert (np.float32(bo.mzmin), np.float32(bo.mzmax), np.float32(bo.mesh_dz), np.float32(bo.radius)) == tuple(np.float32(br[k]) for k in ("mzmin", "mzmax", "mesh_dz", "radius"))
            if bo.wx2 >= 0:      # some cell below the sea level: the reference's box is in global cell coordinates (x1 + x)
                assert (x1 + bo.wx1, y1 + bo.wy1, x1 + bo.wx2, y1 + bo.wy2) == br["wbox"]
            else:                # none: the reference keeps its denormalised start value (x2, y2, x1, y1)
                assert br["wbox"] == (x1 + size, y1 + size, x1, y1)
    RL.ref_set_threads(8)


def test_tile_ao_lighting_extracted_reference(oracle, ref, beq):
    """SURVEY 8f row N1 pinned against the reference's OWN tile_t::calc_mesh_ao_lighting() (cut out of src/tiled_mesh.cpp:586-662), both flows:
    (a) CPU gen modes: it builds the (stride+72)^2 context itself - zvals inside the tile, eval_index outside - from ERODED zvals;
    (b) GPU gen modes: create_zvals left the UN-eroded context in ao_zvals and the rays test it inside the tile too (only z0 is eroded)."""
    _need_extract(ref)
    RL = ref.lib()
    assert RL.ref_ao_ray_len() == 36
    for mode, size in ((1, 64), (0, 40), (2, 36), (4, 64), (3, 48)):
        mesh = (size, size, 1)
        ref.setup(mesh=mesh, mode=mode, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
        dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
        half_dxy = 0.5 * (dx + dy)
        zv, csz = size + 2, size + 1 + 72
        for x1, y1 in ((0, 0), (-7 * size, 11 * size)):
            gx, gy = float(x1 - mesh[0] // 2), float(y1 - mesh[1] // 2)
            zvals = ref.heightgen(gx, gy, dx, dy, zv, zv, 0, 1)
            context = ref.heightgen(gx - 36, gy - 36, dx, dy, csz, csz, 0, 1)
            lo, hi = float(zvals.min()), float(zvals.max())
            eroded = ref.apply_erosion(zvals, lo, 400, water_plane_z=lo - 10, zmin=lo - 0.1, zmax=hi + 0.1, clip_hd1=0.5)
            assert (eroded != zvals).any()
            for z, hd in ((zvals, half_dxy), (eroded, half_dxy), (eroded, 0.02 * half_dxy)):     # the last: nearly horizontal rays => many occluded cells
                if mode < 3:
                    ar = ref.tile_ao_lighting(size, x1, y1, z, None, hd)
                    ao = oracle.tile_ao(z[None], context[None], hd, use_ao_zvals=False)[0]
                else:
                    ar = ref.tile_ao_lighting(size, x1, y1, z, context, hd)
                    ao = oracle.tile_ao(z[None], context[None], hd, use_ao_zvals=True)[0]
                assert np.array_equal(ar, ao), (mode, size, x1, y1)
            assert ar.min() < 200 and ar.max() > ar.min()


def _need_vox(ref):
    import pytest
    if not ref.has_voxel_extract():
        pytest.skip("oracle/_ref was built without the voxels.cpp extraction")


def _vox_case(oracle, ref, dims, vsz, center, gen_mode, **kw):
    """A reference voxel_manager filled by its own create_procedural + the matching oracle parameter blocks."""
    nx, ny, nz = dims
    ref.setup(mode=gen_mode if gen_mode else 0, freq_filter=2, seed=1)
    V = ref.Vox(nx, ny, nz, vsz, center, **kw)
    lo = V.lo_pos
    vpp = oracle.VoxelPostParams()
    vpp.nx, vpp.ny, vpp.nz = nx, ny, nz
    for d in range(3):
        vpp.lo_pos[d], vpp.vsz[d] = float(lo[d]), vsz[d]
    vpp.isolevel, vpp.invert, vpp.make_closed_surface = kw.get("isolevel", 0.0), kw.get("invert", 0), kw.get("make_closed_surface", 1)
    vpp.remove_unconnected, vpp.keep_at_edge = kw.get("remove_unconnected", 1), int(kw.get("keep_at_scene_edge", 0) == 1)
    vpp.centre_seed = int(kw.get("atten_at_edges", 0) in (3, 4) or not kw.get("use_mesh", 0))
    vpp.skip_under_mesh = 0
    return V, vpp, lo


def test_voxel_fill_against_extracted_create_procedural(oracle, ref, beq):
    """SURVEY 8a row a16 pinned against the reference's OWN voxel_manager::create_procedural + atten_* (cut out of src/voxels.cpp:278-346,403-482
    at build time): sine / simplex / Perlin density, z gradient off, every attenuation mode."""
    _need_vox(ref)
    for gen_mode in (0, 1, 2):
        for atten in (0, 1, 2, 3, 4, 5):
            dims, vsz, center = (20, 12, 28), (0.11, 0.13, 0.07), (0.3, -0.2, 0.1)
            V, vpp, lo = _vox_case(oracle, ref, dims, vsz, center, gen_mode, atten_at_edges=atten)
            off = (0.5, -1.25, 2.0)
            V.create_procedural(1.0, 1.0, off, 1, 123, 456, gen_mode)
            if atten:
                V.atten(atten, -0.8, 0.45)
            vp = oracle.VoxelParams()
            vp.nx, vp.ny, vp.nz = dims
            for d in range(3):
                vp.lo_pos[d], vp.vsz[d], vp.offset[d] = float(lo[d]), vsz[d], off[d]
            vp.mag = vp.freq = 1.0
            vp.gen_mode, vp.normalize_to_1, vp.rseed1, vp.rseed2 = gen_mode, 1, 123, 456
            vp.octaves = max(1, 5 - 2)                     # MAX_FREQ_BINS - mesh_freq_filter (freq_filter 2)
            vp.rx, vp.ry = oracle.gen_rx_ry(1, 0, gen_mode) if gen_mode else (0.0, 0.0)
            vp.zscale = 0.0
            vp.atten_mode, vp.atten_val, vp.atten_inner_radius = atten, -0.8, 0.45
            assert beq(oracle.voxel_fill(vp), V.vals()) == 0, (gen_mode, atten)


def test_voxel_post_processing_against_extracted_reference(oracle, ref, beq):
    """SURVEY 8f row N3 pinned against the reference's OWN determine_voxels_outside, remove_unconnected_outside(_range), flood_fill_range,
    remove_interior_holes and add_triangles_for_voxel (cut out of src/voxels.cpp at build time): outside flags, the flood fills with their
    make_voxel_outside/inside edits, per-cube triangle counts and the unwelded triangle soup, bit for bit and in the same order."""
    _need_vox(ref)
    tables = ref.mc_tables()
    rng = np.random.default_rng(4)
    cases = [dict(dims=(24, 20, 16), gen=0, kw=dict(remove_unconnected=3)),
             dict(dims=(18, 22, 30), gen=1, kw=dict(remove_unconnected=3, invert=1, isolevel=0.1)),
             dict(dims=(16, 16, 16), gen=2, kw=dict(remove_unconnected=1, make_closed_surface=0, keep_at_scene_edge=1)),
             dict(dims=(20, 18, 14), gen=0, kw=dict(remove_unconnected=3, atten_at_edges=3, isolevel=-0.2)),
             dict(dims=(26, 24, 20), gen=0, kw=dict(remove_unconnected=2, use_mesh=1), mesh=True),
             dict(dims=(12, 10, 9), gen=-1, kw=dict(remove_unconnected=3))]          # random noise field: many small components and pockets
    for c in cases:
        dims, kw = c["dims"], c["kw"]
        vsz, center = (0.15, 0.12, 0.1), (0.0, 0.0, 0.3)
        V, vpp, lo = _vox_case(oracle, ref, dims, vsz, center, max(c["gen"], 0), **kw)
        if c["gen"] >= 0:
            V.create_procedural(1.0, 1.3, (0.2, 0.1, -0.3), 1, 123, 456, c["gen"])
            if kw.get("atten_at_edges"):
                V.atten(kw["atten_at_edges"], -1.0, 0.4)
        else:
            V.set_vals(rng.uniform(-1, 1, (dims[1], dims[0], dims[2])).astype(np.float32))
        zix = None
        if c.get("mesh"):       # ground mesh under the voxels: z_min_matrix -> per-column zix (computed with the reference's own get_xpos / point_outside_mesh)
            V.set_zmin_matrix(np.fromfunction(lambda y, x: 0.25 * np.sin(x * 0.3) + 0.2 * np.cos(y * 0.2) + 0.2, (128, 128)).astype(np.float32))
            zix = V.zix()
            assert zix.max() > 0
        vals0 = V.vals()
        V.determine_outside()
        out_r = V.outside()
        out_o = oracle.voxel_outside(vals0, vpp, zix)
        assert np.array_equal(out_r, out_o), c
        V.remove_unconnected()
        if kw["remove_unconnected"] > 2:
            V.remove_interior_holes()
        vals_o, out2_o, changed = oracle.voxel_remove_unconnected(vals0, out_o, vpp)
        assert np.array_equal(V.outside(), out2_o) and beq(V.vals(), vals_o) == 0, c
        assert changed == int((out2_o != out_o).sum())
        for skip in (0, 1) if c.get("mesh") else (0,):
            vpp.skip_under_mesh = skip
            ref.lib().ref_vox_set_display_mode_bit(skip) if hasattr(ref.lib(), "ref_vox_set_display_mode_bit") else None
            if skip and not hasattr(ref.lib(), "ref_vox_set_display_mode_bit"):
                continue
            tr, counts = V.triangles(welded=False, want_counts=True)
            to = oracle.voxel_triangles(vals_o, out2_o, vpp, tables)
            assert tr.shape == to.shape and beq(tr, to) == 0, (c, skip)
            assert counts.sum() >= len(tr) > 0
        V.set_zmin_matrix(None)
    changed_any = True
    assert changed_any


def test_mesh_shadows_against_extracted_reference(oracle, ref, beq):
    """SURVEY 8f row N4 pinned against the reference's OWN mesh_shadow_gen / calc_mesh_shadows / do_line_clip (cut out of src/visibility.cpp:411-517 and
    src/Math3d.cpp at build time): shadow mask and outgoing shadow heights of one tile for light directions in every octant, steep and grazing, with and
    without incoming heights from neighbour tiles; 1 OpenMP thread (the reference's two sections race on sh_out otherwise)."""
    import pytest
    if not ref.has_shadow_extract():
        pytest.skip("oracle/_ref was built without the visibility.cpp extraction")
    RL = ref.lib()
    RL.ref_set_threads(1)
    rng = np.random.default_rng(9)
    left_tile = 0
    for mesh, zv in (((64, 64, 1), 66), ((128, 128, 1), 130), ((32, 48, 1), 40)):
        ref.setup(mesh=mesh, mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
        dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
        z = ref.heightgen(-zv / 2, -zv / 2, dx, dy, zv, zv, 0, 1)
        z = (z - np.float32(z.mean())) * np.float32(2.0)          # around 0: the rays run at z = 0 and are clipped against [zmin, zmax] (src/visibility.cpp:424)
        zlo, zhi = float(z.min()) - 0.5, float(z.max()) + 0.5
        assert zlo < 0.0 < zhi
        sp = oracle.ShadowParams()
        sp.x_scene_size, sp.y_scene_size, sp.dx_val, sp.dy_val = 4.0, 4.0, dx, dy
        sp.dx_val_inv, sp.dy_val_inv = 1.0 / np.float32(dx), 1.0 / np.float32(dy)
        sp.xy_sum_size, sp.zmin, sp.zmax, sp.no_shadow = mesh[0] + mesh[1], zlo, zhi, 0
        lights = [(3.0, 2.0, 1.0), (-4.0, 1.0, 0.5), (1.0, -5.0, 0.7), (-2.0, -2.0, 3.0), (0.3, 6.0, 0.2), (5.0, 0.0, 1.0), (0.0, -3.0, 2.0), (0.0, 0.0, 5.0), (2.0, 1.0, zlo - 1.0)]
        for lp in lights:
            for d in range(3):
                sp.lpos[d] = lp[d]
            for with_in in (0, 1):
                six = siy = None
                if with_in:   # heights a neighbour tile would hand over: some entries "none" (MESH_MIN_Z), some above the terrain
                    six = np.where(rng.random(zv) < 0.3, -1.0e6, z[0] + rng.uniform(-0.2, 0.6, zv)).astype(np.float32)
                    siy = np.where(rng.random(zv) < 0.3, -1.0e6, z[:, 0] + rng.uniform(-0.2, 0.6, zv)).astype(np.float32)
                mr, oxr, oyr = ref.calc_mesh_shadows(lp, z, zlo, zhi, six, siy)
                mo, oxo, oyo = oracle.calc_mesh_shadows(sp, z, six, siy)
                assert np.array_equal(mr, mo), (mesh, lp, with_in, int((mr != mo).sum()))
                assert beq(oxr, oxo) == 0 and beq(oyr, oyo) == 0, (mesh, lp, with_in)
                if lp[2] < 1.0 and lp[2] > zlo and (lp[0] or lp[1]):
                    assert 0 < (mo == 2).sum() < mo.size, (mesh, lp)     # a real, partial shadow
                    left_tile += int((oxo > -1e5).any() or (oyo > -1e5).any())
        assert 0 < (mo == 2).sum()
    assert left_tile > 4         # shadows that run off the tile hand heights to the neighbours (sh_out)
    RL.ref_set_threads(8)


def test_mesh_shadow_chaining_against_extracted_reference(oracle, ref, beq):
    """tile_t::calc_shadows_for_light's chaining (src/tiled_mesh.cpp:664-692): a 3x3 block of tiles processed toward-the-light-first, each tile's sh_in = the sh_out of
    its neighbours toward the light, done here by hand with the reference's own calc_mesh_shadows per tile, vs the oracle's batch function."""
    import pytest
    if not ref.has_shadow_extract():
        pytest.skip("oracle/_ref was built without the visibility.cpp extraction")
    RL = ref.lib()
    RL.ref_set_threads(1)
    S, zv = 32, 34
    ref.setup(mesh=(S, S, 1), mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
    dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
    txy = [(tx, ty) for ty in range(-1, 2) for tx in range(2, 5)]
    tiles = np.stack([ref.heightgen(tx * S - S // 2, ty * S - S // 2, dx, dy, zv, zv, 0, 1) for tx, ty in txy])
    tiles = ((tiles - np.float32(tiles.mean())) * np.float32(3.0)).astype(np.float32)
    zlo, zhi = float(tiles.min()) - 0.5, float(tiles.max()) + 0.5
    assert zlo < 0.0 < zhi
    sp = oracle.ShadowParams()
    sp.x_scene_size, sp.y_scene_size, sp.dx_val, sp.dy_val = 4.0, 4.0, dx, dy
    sp.dx_val_inv, sp.dy_val_inv = 1.0 / np.float32(dx), 1.0 / np.float32(dy)
    sp.xy_sum_size, sp.zmin, sp.zmax, sp.no_shadow = 2 * S, zlo, zhi, 0
    chained = 0
    for lp in ((3.0, 2.0, 0.4), (-4.0, 1.0, 0.3), (1.0, -5.0, 0.5), (-2.0, -3.0, 0.6)):
        for d in range(3):
            sp.lpos[d] = lp[d]
        sx, sy = (-1 if lp[0] < 0 else 1), (-1 if lp[1] < 0 else 1)
        idx = {t: i for i, t in enumerate(txy)}
        done, masks = {}, {}
        order = sorted(range(len(txy)), key=lambda i: -(sx * txy[i][0] + sy * txy[i][1]))     # closest to the light first
        for i in order:
            nbx, nby = idx.get((txy[i][0] + sx, txy[i][1])), idx.get((txy[i][0], txy[i][1] + sy))
            six = done[nby][0] if nby is not None else None      # y neighbour's sh_out[0]
            siy = done[nbx][1] if nbx is not None else None      # x neighbour's sh_out[1]
            m, ox, oy = ref.calc_mesh_shadows(lp, tiles[i], zlo, zhi, six, siy)
            done[i], masks[i] = (ox, oy), m
        mo, oxo, oyo = oracle.tile_shadows_batch(tiles, txy, sp)
        for i in range(len(txy)):
            assert np.array_equal(masks[i], mo[i]) and beq(done[i][0], oxo[i]) == 0 and beq(done[i][1], oyo[i]) == 0, (lp, i)
        # sh_out is only written where x == xb / y == yb is inside the tile, i.e. for rays leaving through the x = 0 / y = 0 edge (src/visibility.cpp:461-462): light from +x / +y
        if lp[0] > 0 or lp[1] > 0:
            assert any((done[i][0] > -1e5).any() or (done[i][1] > -1e5).any() for i in range(len(txy)))   # shadows did cross tile borders
            alone = [ref.calc_mesh_shadows(lp, tiles[i], zlo, zhi, None, None)[0] for i in range(len(txy))]
            chained += sum(int(not np.array_equal(alone[i], masks[i])) for i in range(len(txy)))
    assert chained > 0          # the handed-over heights changed some neighbour's mask
    RL.ref_set_threads(8)


def _weight_cases(ref, P, z, dx, dy, size, shape):
    """Parameter sets for the terrain weights texture: the reference's default tables with the z range placed so that every ground texture, the smooth bands
    between them, steep grass / steep snow and the water line all occur; then snow -> rock, no vegetation, a permuted texture table."""
    ids = ref.tex_ids()                                     # engine ids in the order {sand, dirt, ground, rock, snow}
    lo, hi = float(z.min()), float(z.max())
    base = dict(h_dirt=[0.40, 0.44, 0.60, 0.75, 1.0], order=[0, 1, 2, 3, 4], zmin=lo - 0.05 * (hi - lo), zmax=hi + 0.05 * (hi - lo), relh_adj_tex=0.0, vegetation=1.0,
                snow_to_rock=0, corners=[1.0, 0.7, 0.3, 0.9, 0.0, 0.6, 1.2, 0.4])
    cases = [base, dict(base, relh_adj_tex=-0.08, corners=[0.2, 0.5, 0.55, 0.45, 1.0, 1.0, 0.1, 0.9]), dict(base, snow_to_rock=1), dict(base, vegetation=0.0),
             dict(base, h_dirt=[0.2, 0.5, 0.62, 0.7, 1.0], order=[1, 0, 2, 4, 3], relh_adj_tex=0.03)]
    for c in cases:
        wp = P()
        for i in range(5):
            wp.h_dirt[i] = c["h_dirt"][i]
            wp.tex_class[i] = c["order"][i]
        wp.sthresh[0][0], wp.sthresh[0][1], wp.sthresh[1][0], wp.sthresh[1][1] = 0.68, 0.86, 0.48, 0.72      # src/mesh_gen.cpp:44
        wp.zmin, wp.zmax, wp.relh_adj_tex, wp.vegetation, wp.snow_to_rock = c["zmin"], c["zmax"], c["relh_adj_tex"], c["vegetation"], c["snow_to_rock"]
        wp.water_level = float(ref.lib().ref_get_water_z_height())
        wp.noise_scale = np.float32((2.0 if shape == 2 else 1.0) * float(np.float32(0.003)) * 1.0)         # mesh_scale_z = 1
        wp.vnz_scale = 1.0
        wp.dx_val, wp.dy_val, wp.dxdy = dx, dy, float(np.float32(dx) * np.float32(dy))
        wp.xy_mult = np.float32(1.0 / float(np.float32(size)))
        yield c, wp, [ids[k] for k in c["order"]]


def test_tile_weights_against_extracted_create_texture(oracle, ref):
    """to_tile_weights (the oracle of tw_tile_weights_batch) == the reference's own tile_t::create_texture, cut out of src/tiled_mesh.cpp at build time
    (with get_tids / update_lttex_ix from src/Textures.cpp), on reference-generated tiles: every RGBA byte and has_any_grass."""
    import pytest
    if not ref.has_texture_extract():
        pytest.skip("oracle/_ref was built without the create_texture extraction")
    RL = ref.lib()
    RL.ref_set_threads(1)
    seen, blended = np.zeros(4, np.int64), 0
    for mode, S, shape in ((1, 64, 0), (0, 32, 0), (2, 48, 2)):
        zv = S + 2
        ref.setup(mesh=(S, S, 1), mode=mode, freq_filter=1, seed=1, zmax_est=2.3, hmap=HM_CFG, shape=shape)
        dx, dy = RL.ref_get_dx(), RL.ref_get_dy()
        sp = ref.sine_params()
        hp = oracle.HeightParams()
        hp.gen_mode, hp.gen_shape, hp.start_eval_sin, hp.glaciate = mode, shape, oracle.compute_scale(1.0, 1), 1
        hp.mesh_scale, hp.mesh_scale_z_inv = 1.0, 1.0
        hp.dx_val_inv, hp.dy_val_inv = 1.0 / np.float32(dx), 1.0 / np.float32(dy)
        hp.mesh_height, hp.mesh_height_scale = RL.ref_get_mesh_height(), 1.0
        hp.zmax_est, hp.custom_glaciate_exp = 2.3, 0.0
        hp.rx, hp.ry = oracle.gen_rx_ry(1, 0, mode)
        hp.hmap = oracle.hmap_params(**HM_CFG)
        for (x1, y1) in ((3 * S, -5 * S), (-2 * S, 7 * S)):
            z = ref.heightgen(x1 - S // 2, y1 - S // 2, dx, dy, zv, zv, 0, 1)
            z = ((z - np.float32(z.mean())) * np.float32(1.0 / max(1e-6, float(z.std()))) * np.float32(0.4)).astype(np.float32)      # a few cells of relief per cell: slopes on both sides of sthresh
            rand = oracle.weights_noise(hp, sp, [(x1, y1)], (S, S), dx, dy, S + 1)
            for c, wp, ids in _weight_cases(ref, oracle.WeightParams, z, dx, dy, S, shape):
                corners_ref = [v for i in range(4) for v in (c["corners"][i], c["corners"][4 + i])]     # the harness takes {grass, dirt} per corner, the API grass[4] then dirt[4]
                w_ref, hag_ref = ref.tile_create_texture(S, x1, y1, z, corners_ref, c["h_dirt"], ids, c["vegetation"], c["relh_adj_tex"], c["zmin"], c["zmax"], c["snow_to_rock"])
                w, hag = oracle.tile_weights(z[None], rand, [c["corners"]], wp)
                assert np.array_equal(w[0], w_ref), (mode, x1, c, int((w[0] != w_ref).sum()))
                assert int(hag[0]) == hag_ref
                seen += [(w_ref[..., k] > 0).sum() for k in range(4)]
                blended += len(np.unique(w_ref.reshape(-1, 4), axis=0))
    assert (seen > 500).all(), seen                                                 # every channel (sand, dirt, grass, rock) occurs
    assert blended > 2000, blended                                                  # and mostly as blends, not pure textures
    RL.ref_set_threads(8)


def test_tex_height_tables_host_function_against_reference(tw, scene, ref, beq):
    """tw_gen_tex_height_tables (host side of the C ABI) == the reference's init_terrain_mesh() (linked from mesh_gen.o) + gen_tex_height_tables()
    (cut out of src/Textures.cpp at build time): h_dirt[5], the texture order and clip_hd1, over water levels, temperatures and glaciate exponents; scene.py's
    clip_hd1 (what the erosion parameters use) agrees too."""
    import pytest
    if not (ref.has_texture_extract() and hasattr(ref.lib(), "ref_init_terrain_mesh")):
        pytest.skip("oracle/_ref was built without the Textures.cpp extraction")
    ids = ref.tex_ids()
    for rel in (0.0, -0.42, -0.3, -0.1, 0.07, 0.25, 0.58, 0.7):
        for temp in (20.0, 40.0, 41.5, 75.0):
            for gexp in (3.0, 1.0, 2.5):
                h_ref, id_ref, _, clip_ref = ref.init_terrain_mesh(rel, temp, gexp)
                h, cls, clip = tw.gen_tex_height_tables(rel, temp, gexp)
                assert beq(np.array(h, np.float32), h_ref) == 0, (rel, temp, gexp, h, h_ref)
                assert [ids[c] for c in cls] == id_ref
                assert np.float32(clip) == clip_ref
    for rel in (0.0, 0.1, -0.2):
        cfg = scene.SceneConfig(water_h_off_rel=rel, glaciate=1)
        assert np.float32(cfg.clip_hd1()) == ref.init_terrain_mesh(rel, 20.0, 3.0)[3]
