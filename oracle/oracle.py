"""TEST INFRASTRUCTURE ONLY. ctypes wrapper around oracle/_build/libterrain_oracle.so (the plain-C restatement of the
reference terrain hot path, see terrain_oracle.h). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product package never does."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libterrain_oracle.so")


# ---- POD mirrors of include/tw3d.h (types only) ----
class HmapParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("plat_bot", "plat_h", "plat_s", "plat_max", "crat_h", "crat_s", "crack_lo", "crack_hi",
                                         "crack_d", "sine_mag", "sine_freq", "sine_bias", "volcano_width", "volcano_height")]


class HeightParams(C.Structure):
    _fields_ = [("gen_mode", C.c_int), ("gen_shape", C.c_int), ("start_eval_sin", C.c_int), ("glaciate", C.c_int),
                ("mesh_scale", C.c_float), ("mesh_scale_z_inv", C.c_float), ("dx_val_inv", C.c_float), ("dy_val_inv", C.c_float),
                ("mesh_height", C.c_float), ("mesh_height_scale", C.c_float), ("zmax_est", C.c_float),
                ("custom_glaciate_exp", C.c_float), ("rx", C.c_float), ("ry", C.c_float), ("hmap", HmapParams)]


class Grid2D(C.Structure):
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("dx", C.c_float), ("dy", C.c_float), ("nx", C.c_uint32), ("ny", C.c_uint32)]


class MinMax(C.Structure):
    _fields_ = [("zmin", C.c_float), ("zmax", C.c_float)]


class ErosionParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("erode_amount", "water_plane_z", "half_dxy", "zmin", "zmax", "relh_adj_tex", "clip_hd1")]


class VoxelParams(C.Structure):
    _fields_ = [("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
                ("lo_pos", C.c_float * 3), ("vsz", C.c_float * 3), ("offset", C.c_float * 3),
                ("mag", C.c_float), ("freq", C.c_float), ("gen_mode", C.c_int), ("normalize_to_1", C.c_int),
                ("rseed1", C.c_int), ("rseed2", C.c_int), ("octaves", C.c_int), ("rx", C.c_float), ("ry", C.c_float),
                ("zscale", C.c_float), ("atten_mode", C.c_int), ("atten_val", C.c_float), ("atten_inner_radius", C.c_float)]


class TileBounds(C.Structure):
    _fields_ = [("sub_zmin", C.c_float * 16), ("sub_zmax", C.c_float * 16), ("mzmin", C.c_float), ("mzmax", C.c_float), ("mesh_dz", C.c_float),
                ("radius", C.c_float), ("wx1", C.c_int32), ("wy1", C.c_int32), ("wx2", C.c_int32), ("wy2", C.c_int32)]


class VoxelPostParams(C.Structure):
    _fields_ = [("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32), ("lo_pos", C.c_float * 3), ("vsz", C.c_float * 3), ("isolevel", C.c_float),
                ("invert", C.c_int), ("make_closed_surface", C.c_int), ("remove_unconnected", C.c_int), ("keep_at_edge", C.c_int), ("centre_seed", C.c_int),
                ("skip_under_mesh", C.c_int)]


class ShadowParams(C.Structure):
    _fields_ = [("lpos", C.c_float * 3), ("x_scene_size", C.c_float), ("y_scene_size", C.c_float), ("dx_val", C.c_float), ("dy_val", C.c_float), ("dx_val_inv", C.c_float),
                ("dy_val_inv", C.c_float), ("xy_sum_size", C.c_int), ("zmin", C.c_float), ("zmax", C.c_float), ("no_shadow", C.c_int)]


class WeightParams(C.Structure):
    _fields_ = [("h_dirt", C.c_float * 5), ("tex_class", C.c_int * 5), ("class_ix", C.c_int * 5), ("sthresh", (C.c_float * 2) * 2), ("zmin", C.c_float), ("zmax", C.c_float),
                ("relh_adj_tex", C.c_float), ("water_level", C.c_float), ("noise_scale", C.c_float), ("vnz_scale", C.c_float), ("vegetation", C.c_float), ("snow_to_rock", C.c_int),
                ("dx_val", C.c_float), ("dy_val", C.c_float), ("dxdy", C.c_float), ("xy_mult", C.c_float)]


class PointQuery(C.Structure):
    _fields_ = [("kind", C.c_int), ("xy_scale", C.c_float), ("mesh_x_size", C.c_int), ("mesh_y_size", C.c_int), ("x_scene_size", C.c_float),
                ("y_scene_size", C.c_float), ("xoff2", C.c_int), ("yoff2", C.c_int), ("no_xyoff", C.c_int)]


class HmapSampler(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("edge_mode", C.c_int), ("mesh_scale", C.c_float), ("h_scale", C.c_float),
                ("mesh_file_scale", C.c_float), ("mesh_file_tz", C.c_float), ("mesh_scale_z_inv", C.c_float)]


class Rng(C.Structure):
    _fields_ = [("rseed1", C.c_int64), ("rseed2", C.c_int64)]


def hmap_params(**kw):
    h = HmapParams(1000.0, 0, 0, 0, 1000.0, 0, 0, 0, 0, 0, 0, 0, 0, 0)  # hmap_params_t defaults, src/mesh.h:85-88
    for k, v in kw.items():
        setattr(h, k, v)
    return h


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "terrain_oracle.c")):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.to_rng_rand.argtypes = [C.POINTER(Rng)]
        L.to_rng_rand.restype = C.c_int
        L.to_rng_randd.argtypes = [C.POINTER(Rng)]
        L.to_rng_randd.restype = C.c_double
        L.to_rng_rand_float.argtypes = [C.POINTER(Rng)]
        L.to_rng_rand_float.restype = C.c_float
        L.to_build_sin_table.argtypes = [vp]
        L.to_compute_scale.argtypes = [C.c_float, C.c_int]
        L.to_compute_scale.restype = C.c_int
        L.to_gen_sine_params.argtypes = [C.POINTER(Rng), C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_float, C.c_float, vp]
        L.to_gen_rx_ry.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.to_water_z_height.argtypes = [C.c_float, C.c_int, C.c_float, C.c_float, C.c_float]
        L.to_water_z_height.restype = C.c_float
        for n, k in (("to_simplex2", 2), ("to_perlin2", 2), ("to_simplex3", 3), ("to_perlin3", 3)):
            getattr(L, n).argtypes = [C.c_float] * k
            getattr(L, n).restype = C.c_float
        L.to_get_noise_zval.argtypes = [C.c_float, C.c_float, C.POINTER(HeightParams)]
        L.to_get_noise_zval.restype = C.c_float
        L.to_eval_mesh_sin_terms.argtypes = [C.c_float, C.c_float, vp, vp, C.c_int]
        L.to_eval_mesh_sin_terms.restype = C.c_float
        L.to_heightgen_2d.argtypes = [C.POINTER(Grid2D), C.POINTER(HeightParams), vp, vp, C.c_int, C.c_int, vp, C.c_int]
        L.to_glaciate_mesh.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(HeightParams), vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.to_gen_mesh.argtypes = [C.POINTER(Rng), C.POINTER(HeightParams), C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint, C.POINTER(ErosionParams), vp, vp, vp, vp]
        L.to_tile_bounds.argtypes = [vp, C.c_uint, C.c_uint, C.c_float, C.c_float, C.c_float, C.c_uint, vp]
        L.to_tile_normals.argtypes = [vp, C.c_uint, C.c_uint, C.c_float, C.c_float, vp, vp]
        L.to_tile_ao.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_float, C.c_int, vp]
        L.to_hmap_sample_tiles.argtypes = [vp, C.POINTER(HmapSampler), vp, C.c_uint, C.c_uint, vp]
        L.to_eval_points.argtypes = [vp, C.c_size_t, C.POINTER(HeightParams), C.POINTER(PointQuery), vp, vp, vp]
        L.to_apply_erosion.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_uint, C.POINTER(ErosionParams)]
        L.to_apply_erosion.restype = C.c_ulonglong
        L.to_erode_sweeps.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_uint, C.POINTER(ErosionParams), C.c_uint, C.c_int]
        L.to_erode_sweeps.restype = C.c_ulonglong
        L.to_calc_mesh_shadows.argtypes = [C.POINTER(ShadowParams), vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]
        L.to_tile_shadows_batch.argtypes = [vp, vp, C.c_uint, C.c_uint, C.POINTER(ShadowParams), vp, vp, vp]
        L.to_voxel_outside.argtypes = [vp, C.POINTER(VoxelPostParams), vp, vp]
        L.to_voxel_remove_unconnected.argtypes = [vp, vp, C.POINTER(VoxelPostParams)]
        L.to_voxel_remove_unconnected.restype = C.c_ulonglong
        L.to_voxel_triangles.argtypes = [vp, vp, C.POINTER(VoxelPostParams), vp, vp, vp, vp, C.c_ulonglong]
        L.to_voxel_triangles.restype = C.c_ulonglong
        L.to_noise3d_gen_sines.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, vp]
        L.to_noise3d_get_val_pt.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float]
        L.to_noise3d_get_val_pt.restype = C.c_float
        L.to_voxel_fill.argtypes = [C.POINTER(VoxelParams), vp, vp, vp, C.c_int]
        L.to_from_floats_u16.argtypes = [vp, C.c_size_t, C.c_float, C.c_float, vp]
        L.to_from_floats_u16.restype = C.c_size_t
        L.to_to_floats_u16.argtypes = [vp, C.c_size_t, C.c_float, C.c_float, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_SIN_TABLE = None


def sin_table():
    global _SIN_TABLE
    if _SIN_TABLE is None:
        t = np.empty(65536, np.float32)
        lib().to_build_sin_table(_p(t))
        _SIN_TABLE = t
    return _SIN_TABLE


def compute_scale(mesh_scale, freq_filter):
    return lib().to_compute_scale(mesh_scale, freq_filter)


def gen_sine_params(scaled_height, mesh=(128, 128), scene=(4.0, 4.0), seed=0, rgen_index=0, mode=0, rng=None,
                    start_mag=0.02, start_freq=240.0, mag_mult=2.0, freq_mult=0.5):
    rng = rng if rng is not None else Rng(1, 1)
    out = np.empty((90, 5), np.float32)
    lib().to_gen_sine_params(C.byref(rng), scaled_height, mesh[0], mesh[1], scene[0], scene[1], seed, rgen_index, mode,
                             start_mag, start_freq, mag_mult, freq_mult, _p(out))
    return out


def gen_rx_ry(seed, rgen_index, mode):
    rx, ry = C.c_float(), C.c_float()
    lib().to_gen_rx_ry(seed, rgen_index, mode, C.byref(rx), C.byref(ry))
    return rx.value, ry.value


def heightgen_2d(grid, hp, sine_params=None, enable_glaciate=1, min_start_sin=0, nthreads=0):
    out = np.empty((grid.ny, grid.nx), np.float32)
    sp = np.ascontiguousarray(sine_params if sine_params is not None else np.zeros((90, 5)), np.float32)
    lib().to_heightgen_2d(C.byref(grid), C.byref(hp), _p(sin_table()), _p(sp), int(enable_glaciate), int(min_start_sin), _p(out), nthreads)
    return out


def gen_mesh(hp, mesh=(128, 128), scene=(4.0, 4.0), seed=0, rgen_index=0, xoff2=0, yoff2=0, dx=0.0625, dy=0.0625, erosion_iters=0, ep=None, rng=None,
             water_h_off=0.0, water_h_off_rel=0.0):
    """to_gen_mesh: returns (mesh, zvals6 dict, sine_params); hp.zmax_est and ep.water_plane_z/zmin/zmax are outputs."""
    rng = rng if rng is not None else Rng(1, 1)
    ep = ep if ep is not None else ErosionParams(1.0, 0.0, 0.0625, -1.0, 1.0, 0.0, 0.5)
    out = np.empty((mesh[1], mesh[0]), np.float32)
    sp = np.empty((90, 5), np.float32)
    z6 = np.empty(6, np.float32)
    lib().to_gen_mesh(C.byref(rng), C.byref(hp), mesh[0], mesh[1], scene[0], scene[1], seed, rgen_index, xoff2, yoff2, dx, dy, water_h_off, water_h_off_rel,
                      erosion_iters, C.byref(ep), _p(sin_table()), _p(sp), _p(out), _p(z6))
    return out, dict(zip(("zmin", "zmax", "zmax_est", "zbottom", "ztop", "water_plane_z"), (float(v) for v in z6))), sp


def tile_bounds(tiles, wpz_max, dx_val, dy_val, size):
    tiles = np.ascontiguousarray(tiles, np.float32)
    nt, zv = tiles.shape[0], tiles.shape[1]
    out = (TileBounds * nt)()
    lib().to_tile_bounds(_p(tiles), nt, zv, wpz_max, dx_val, dy_val, size, C.cast(out, C.c_void_p))
    return out


def tile_normals(tiles, dx_val, dy_val):
    tiles = np.ascontiguousarray(tiles, np.float32)
    nt, zv = tiles.shape[0], tiles.shape[1]
    rgba = np.empty((nt, zv - 1, zv - 1, 4), np.uint8)
    mnz = np.empty(nt, np.float32)
    lib().to_tile_normals(_p(tiles), nt, zv, dx_val, dy_val, _p(rgba), _p(mnz))
    return rgba, mnz


def tile_ao(tiles, contexts, half_dxy, use_ao_zvals=False):
    """contexts: [nt, stride+72, stride+72] heights generated at origin (x1-36, y1-36) (heightgen_2d of the shifted, enlarged grid).
    use_ao_zvals: the reference's GPU-gen-mode flow (rays test the un-eroded context inside the tile too)."""
    tiles = np.ascontiguousarray(tiles, np.float32)
    contexts = np.ascontiguousarray(contexts, np.float32)
    nt, zv = tiles.shape[0], tiles.shape[1]
    assert contexts.shape == (nt, zv - 1 + 72, zv - 1 + 72)
    ao = np.empty((nt, zv - 1, zv - 1), np.uint8)
    lib().to_tile_ao(_p(tiles), _p(contexts), nt, zv, half_dxy, int(use_ao_zvals), _p(ao))
    return ao


def hmap_sample_tiles(data16, hs, origins_xy, zvsize):
    data16 = np.ascontiguousarray(data16, np.uint8)
    org = np.ascontiguousarray(origins_xy, np.int32).reshape(-1, 2)
    out = np.empty((org.shape[0], zvsize, zvsize), np.float32)
    lib().to_hmap_sample_tiles(_p(data16), C.byref(hs), _p(org), org.shape[0], zvsize, _p(out))
    return out


def eval_points(xy, hp, pq, sine_params=None):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty(xy.shape[0], np.float32)
    sp = np.ascontiguousarray(sine_params if sine_params is not None else np.zeros((90, 5)), np.float32)
    lib().to_eval_points(_p(xy), xy.shape[0], C.byref(hp), C.byref(pq), _p(sin_table()), _p(sp), _p(out))
    return out


def apply_erosion(h, min_zval, num_iters, ep):
    h = np.array(h, np.float32, order="C", copy=True)
    ys, xs = h.shape
    steps = lib().to_apply_erosion(_p(h), xs, ys, min_zval, num_iters, C.byref(ep))
    return h, int(steps)


def erode_sweeps(h, min_zval, num_iters, ep, sweep, halo):
    """The coherent batched erosion of tw_erode_sweeps (frozen map per sweep, fixed-point deltas, halo rule); returns (map, moves)."""
    h = np.array(h, np.float32, order="C", copy=True)
    ys, xs = h.shape
    steps = lib().to_erode_sweeps(_p(h), xs, ys, min_zval, num_iters, C.byref(ep), sweep, halo)
    return h, int(steps)


def erode_spec_model(h, min_zval, num_iters, ep, window, cap, tile_shift=2, sched_seed=0):
    """to_erode_spec_model: the sequential model of the product's speculative serial-order erosion; returns (map, moves, (rounds, walks, wasted walks, in-place walks))."""
    h = np.array(h, np.float32, order="C", copy=True)
    ys, xs = h.shape
    stats = (C.c_ulonglong * 4)()
    L = lib()
    L.to_erode_spec_model.restype = C.c_ulonglong
    L.to_erode_spec_model.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_uint, C.c_void_p]
    steps = L.to_erode_spec_model(h.ctypes.data, xs, ys, min_zval, num_iters, C.cast(C.byref(ep), C.c_void_p), window, cap, tile_shift, sched_seed, C.cast(stats, C.c_void_p))
    return h, int(steps), tuple(int(v) for v in stats)


def noise3d_gen_sines(rs1, rs2, mag, freq):
    out = np.empty(420, np.float32)
    lib().to_noise3d_gen_sines(rs1, rs2, mag, freq, _p(out))
    return out


def voxel_fill(vp, rdata=None, nthreads=0):
    out = np.empty((vp.ny, vp.nx, vp.nz), np.float32)
    rd = None if rdata is None else np.ascontiguousarray(rdata, np.float32)
    lib().to_voxel_fill(C.byref(vp), None if rd is None else _p(rd), _p(sin_table()), _p(out), nthreads)
    return out


def calc_mesh_shadows(sp, mh, sh_in_x=None, sh_in_y=None):
    """calc_mesh_shadows of one tile: returns (smask, sh_out_x, sh_out_y); sh_out start at MESH_MIN_Z."""
    mh = np.ascontiguousarray(mh, np.float32)
    ys, xs = mh.shape
    smask = np.empty((ys, xs), np.uint8)
    ox, oy = np.full(xs, -1.0e6, np.float32), np.full(ys, -1.0e6, np.float32)
    six = None if sh_in_x is None else np.ascontiguousarray(sh_in_x, np.float32)
    siy = None if sh_in_y is None else np.ascontiguousarray(sh_in_y, np.float32)
    lib().to_calc_mesh_shadows(C.byref(sp), _p(mh), _p(smask), xs, ys, None if six is None else _p(six), None if siy is None else _p(siy), _p(ox), _p(oy))
    return smask, ox, oy


def tile_shadows_batch(tiles, tile_xy, sp):
    tiles = np.ascontiguousarray(tiles, np.float32)
    nt, zv = tiles.shape[0], tiles.shape[1]
    txy = np.ascontiguousarray(tile_xy, np.int32).reshape(-1, 2)
    smask = np.empty((nt, zv, zv), np.uint8)
    ox, oy = np.empty((nt, zv), np.float32), np.empty((nt, zv), np.float32)
    lib().to_tile_shadows_batch(_p(tiles), _p(txy), nt, zv, C.byref(sp), _p(smask), _p(ox), _p(oy))
    return smask, ox, oy


def voxel_outside(vals, vpp, zix_xy=None):
    vals = np.ascontiguousarray(vals, np.float32)
    out = np.empty(vals.shape, np.uint8)
    z = None if zix_xy is None else np.ascontiguousarray(zix_xy, np.uint32)
    lib().to_voxel_outside(_p(vals), C.byref(vpp), None if z is None else _p(z), _p(out))
    return out


def voxel_remove_unconnected(vals, outside, vpp):
    vals = np.array(vals, np.float32, order="C", copy=True)
    outside = np.array(outside, np.uint8, order="C", copy=True)
    changed = lib().to_voxel_remove_unconnected(_p(vals), _p(outside), C.byref(vpp))
    return vals, outside, int(changed)


def voxel_triangles(vals, outside, vpp, tables):
    vals = np.ascontiguousarray(vals, np.float32)
    outside = np.ascontiguousarray(outside, np.uint8)
    e, t, v = (np.ascontiguousarray(tables[0], np.uint32), np.ascontiguousarray(tables[1], np.int32), np.ascontiguousarray(tables[2], np.uint32))
    n = lib().to_voxel_triangles(_p(vals), _p(outside), C.byref(vpp), _p(e), _p(t), _p(v), None, 0)
    tris = np.empty((n, 3, 3), np.float32)
    lib().to_voxel_triangles(_p(vals), _p(outside), C.byref(vpp), _p(e), _p(t), _p(v), _p(tris), n)
    return tris


def from_floats_u16(vals, val_mult, val_add):
    vals = np.ascontiguousarray(vals, np.float32).ravel()
    out = np.empty(2 * vals.size, np.uint8)
    bad = lib().to_from_floats_u16(_p(vals), vals.size, val_mult, val_add, _p(out))
    return out, int(bad)


def to_floats_u16(data, val_mult, val_add):
    data = np.ascontiguousarray(data, np.uint8).ravel()
    out = np.empty(data.size // 2, np.float32)
    lib().to_to_floats_u16(_p(data), out.size, val_mult, val_add, _p(out))
    return out


def weights_noise(hp, sine_params, origins_xy, mesh_size, dx, dy, stride):
    """The jitter noise grids of tile_t::create_texture: build_arrays(x1 - MESH_X_SIZE/2, y1 - MESH_Y_SIZE/2, 80*DX_VAL, 80*DY_VAL, stride, stride, 0, force_sine_mode=1),
    eval_index(x, y, 50) - un-scaled, one [stride, stride] grid per tile."""
    import copy
    h = copy.copy(hp)
    h.gen_mode, h.gen_shape = 0, 0                           # force_sine_mode: gen_mode = MGEN_SINE, gen_shape = 0 (src/mesh_gen.cpp:592-593)
    fx, fy = np.float32(80.0) * np.float32(dx), np.float32(80.0) * np.float32(dy)
    out = []
    for x1, y1 in origins_xy:
        g = Grid2D(float(x1 - mesh_size[0] // 2), float(y1 - mesh_size[1] // 2), float(fx), float(fy), stride, stride)
        out.append(heightgen_2d(g, h, sine_params, 0, 50))
    return np.stack(out)


def tile_weights(zvals, rand, tile_params, wp):
    """to_tile_weights: zvals [nt, zv, zv], rand [nt, zv-1, zv-1] (weights_noise), tile_params [nt, 8] -> (rgba [nt, zv-1, zv-1, 4] uint8, has_any_grass [nt] uint8)."""
    z = np.ascontiguousarray(zvals, np.float32)
    nt, zv = z.shape[0], z.shape[1]
    r = np.ascontiguousarray(rand, np.float32)
    tp = np.ascontiguousarray(tile_params, np.float32).reshape(nt, 8)
    out = np.empty((nt, zv - 1, zv - 1, 4), np.uint8)
    flags = np.empty(nt, np.uint8)
    lib().to_tile_weights(_p(z), _p(r), nt, zv, _p(tp), C.byref(wp), _p(out), _p(flags))
    return out, flags

