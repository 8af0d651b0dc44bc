"""TEST INFRASTRUCTURE ONLY. ctypes wrapper around oracle/_ref/libref3dworld.so, i.e. the UNMODIFIED reference
objects (src/mesh_gen.cpp, src/erosion.cpp, src/upsurface.cpp, vendored GLM) linked with the driver in oracle/refbuild/.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref3dworld.so")

HMAP_DEFAULT = [1000.0, 0, 0, 0, 1000.0, 0, 0, 0, 0, 0, 0, 0, 0, 0]  # hmap_params_t defaults, src/mesh.h:85-88
HMAP_FIELDS = ["plat_bot", "plat_h", "plat_s", "plat_max", "crat_h", "crat_s", "crack_lo", "crack_hi", "crack_d",
               "sine_mag", "sine_freq", "sine_bias", "volcano_width", "volcano_height"]


class Scene(C.Structure):
    _fields_ = [("mesh_x", C.c_int), ("mesh_y", C.c_int), ("mesh_z", C.c_int),
                ("xss", C.c_float), ("yss", C.c_float), ("zss", C.c_float),
                ("mesh_scale", C.c_float), ("mesh_height_scale", C.c_float),
                ("gen_mode", C.c_int), ("gen_shape", C.c_int), ("freq_filter", C.c_int), ("seed", C.c_int),
                ("rgen_index", C.c_int), ("glaciate", C.c_int), ("custom_glaciate_exp", C.c_float),
                ("hmap", C.c_float * 14), ("zmax_est", C.c_float), ("mesh_scale_z", C.c_float)]


class Erosion(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("erode_amount", "water_plane_z", "half_dxy", "zmin", "zmax", "relh_adj_tex", "clip_hd1")]


class VoxParams(C.Structure):
    _fields_ = [("nx", C.c_uint), ("ny", C.c_uint), ("nz", C.c_uint), ("vsz", C.c_float * 3), ("center", C.c_float * 3), ("isolevel", C.c_float),
                ("invert", C.c_int), ("make_closed_surface", C.c_int), ("remove_unconnected", C.c_int), ("keep_at_scene_edge", C.c_int), ("atten_at_edges", C.c_int),
                ("remove_under_mesh", C.c_int), ("use_mesh", C.c_int), ("radius_val", C.c_float), ("atten_thresh", C.c_float)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.ref_setup.argtypes = [C.POINTER(Scene), C.c_int]
        L.ref_get_start_eval_sin.restype = C.c_int
        for n in ("ref_get_dx", "ref_get_dy", "ref_get_mesh_height", "ref_get_half_dxy", "ref_get_zmax_est", "ref_get_water_z_height"):
            getattr(L, n).restype = C.c_float
        L.ref_set_zmax_est.argtypes = [C.c_float]
        L.ref_set_start_eval_sin.argtypes = [C.c_int]
        L.ref_get_max_threads.restype = C.c_int
        L.ref_heightgen.argtypes = [C.c_float] * 4 + [C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_get_noise_zval.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int]
        L.ref_get_noise_zval.restype = C.c_float
        L.ref_eval_mesh_sin_terms.argtypes = [C.c_float, C.c_float]
        L.ref_eval_mesh_sin_terms.restype = C.c_float
        for n, k in (("ref_glm_simplex2", 2), ("ref_glm_perlin2", 2), ("ref_glm_simplex3", 3), ("ref_glm_perlin3", 3)):
            getattr(L, n).argtypes = [C.c_float] * k
            getattr(L, n).restype = C.c_float
        L.ref_apply_erosion.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint, C.POINTER(Erosion)]
        L.ref_noise3d_rdata.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ref_noise3d_point.argtypes = [C.c_int, C.c_int] + [C.c_float] * 5
        L.ref_noise3d_point.restype = C.c_float
        L.ref_voxel_fill.argtypes = [C.c_uint] * 3 + [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.ref_get_rx_ry.argtypes = [fp, fp]
        L.ref_gen_mesh.argtypes = [C.c_uint, C.POINTER(Erosion), C.c_void_p, C.c_void_p]
        if hasattr(L, "ref_tile_create_zvals"):   # functions cut out of src/tiled_mesh.cpp at build time (oracle/refbuild/build_ref.sh)
            L.ref_tile_create_zvals.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_uint, C.POINTER(Erosion), C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_tile_ao_lighting.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
            L.ref_ao_ray_len.restype = C.c_uint
        if hasattr(L, "ref_vox_init"):   # voxel_manager member functions cut out of src/voxels.cpp at build time
            L.ref_vox_init.argtypes = [C.POINTER(VoxParams)]
            L.ref_vox_create_procedural.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            L.ref_vox_atten.argtypes = [C.c_int, C.c_float, C.c_float]
            L.ref_vox_triangles.argtypes = [C.c_int, C.c_void_p, C.c_ulonglong, C.c_void_p]
            L.ref_vox_triangles.restype = C.c_ulonglong
            for n in ("ref_vox_get_lo_pos", "ref_vox_set_vals", "ref_vox_get_vals", "ref_vox_get_outside", "ref_vox_set_zmin_matrix", "ref_vox_zix"):
                getattr(L, n).argtypes = [C.c_void_p]
            L.ref_vox_tables.argtypes = [C.c_void_p] * 3
            L.ref_vox_set_display_mode_bit.argtypes = [C.c_int]
        if hasattr(L, "ref_calc_mesh_shadows"):   # mesh_shadow_gen / calc_mesh_shadows cut out of src/visibility.cpp at build time
            L.ref_calc_mesh_shadows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def setup(mesh=(128, 128, 1), scene=(4.0, 4.0, 4.0), mesh_scale=1.0, mesh_height_scale=1.0, mode=0, shape=0, freq_filter=2,
          seed=0, rgen_index=0, glaciate=1, custom_glaciate_exp=0.0, hmap=None, zmax_est=0.0, mesh_scale_z=1.0, gen_sine_table=True):
    s = Scene()
    s.mesh_x, s.mesh_y, s.mesh_z = mesh
    s.xss, s.yss, s.zss = scene
    s.mesh_scale, s.mesh_height_scale = mesh_scale, mesh_height_scale
    s.gen_mode, s.gen_shape, s.freq_filter, s.seed, s.rgen_index, s.glaciate = mode, shape, freq_filter, seed, rgen_index, glaciate
    s.custom_glaciate_exp = custom_glaciate_exp
    h = list(HMAP_DEFAULT)
    if hmap:
        for k, v in hmap.items():
            h[HMAP_FIELDS.index(k)] = v
    for i, v in enumerate(h):
        s.hmap[i] = v
    s.zmax_est, s.mesh_scale_z = zmax_est, mesh_scale_z
    lib().ref_setup(C.byref(s), int(gen_sine_table))
    return s


def sin_table():
    out = np.empty(65536, np.float32)
    lib().ref_get_sin_table(out.ctypes.data_as(C.c_void_p))
    return out


def sine_params():
    out = np.empty((90, 5), np.float32)
    lib().ref_get_sine_params(out.ctypes.data_as(C.c_void_p))
    return out


def set_sine_params(p):
    p = np.ascontiguousarray(p, np.float32)
    assert p.shape == (90, 5)
    lib().ref_set_sine_params(p.ctypes.data_as(C.c_void_p))


def rx_ry():
    rx, ry = C.c_float(), C.c_float()
    lib().ref_get_rx_ry(C.byref(rx), C.byref(ry))
    return rx.value, ry.value


def heightgen(x0, y0, dx, dy, nx, ny, cache_values=0, glaciate=1, min_start_sin=0):
    out = np.empty((ny, nx), np.float32)
    lib().ref_heightgen(x0, y0, dx, dy, nx, ny, cache_values, glaciate, min_start_sin, out.ctypes.data_as(C.c_void_p))
    return out


def tile_normals(tile, dx_val, dy_val):
    tile = np.ascontiguousarray(tile, np.float32)
    zv = tile.shape[0]
    rgba = np.empty((zv - 1, zv - 1, 4), np.uint8)
    mnz = C.c_float(0.0)
    lib().ref_tile_normals(tile.ctypes.data_as(C.c_void_p), zv, C.c_float(dx_val), C.c_float(dy_val), rgba.ctypes.data_as(C.c_void_p), C.byref(mnz))
    return rgba, mnz.value


def hmap_sample_tile(data16, x1, y1, zvsize, mesh_scale=1.0, mesh_file_scale=1.0, mesh_file_tz=0.0):
    """terrain_hmap_manager_t::get_clamped_height of the reference over one tile (TEX_EDGE_MODE 2 = mirror, compile-time)."""
    data16 = np.ascontiguousarray(data16, np.uint8)
    h, w = data16.shape[0], data16.shape[1]
    out = np.empty((zvsize, zvsize), np.float32)
    lib().ref_hmap_sample_tile(data16.ctypes.data_as(C.c_void_p), w, h, int(x1), int(y1), zvsize, C.c_float(mesh_scale), C.c_float(mesh_file_scale),
                               C.c_float(mesh_file_tz), out.ctypes.data_as(C.c_void_p))
    return out


def eval_points(kind, xy, xy_scale=1.0, no_xyoff=0, xoff2=0, yoff2=0):
    """kind 0/1/2 = eval_mesh_sin_terms / eval_mesh_sin_terms_scaled / get_exact_zval of the reference for every (x, y) row."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.empty(xy.shape[0], np.float32)
    lib().ref_eval_points(int(kind), xy.ctypes.data_as(C.c_void_p), C.c_size_t(xy.shape[0]), C.c_float(xy_scale), int(no_xyoff), int(xoff2), int(yoff2),
                          out.ctypes.data_as(C.c_void_p))
    return out


def apply_erosion(h, min_zval, num_iters, erode_amount=1.0, water_plane_z=0.0, half_dxy=0.0625, zmin=-1.0, zmax=1.0, relh_adj_tex=0.0, clip_hd1=0.5):
    h = np.array(h, np.float32, order="C", copy=True)
    ys, xs = h.shape
    p = Erosion(erode_amount, water_plane_z, half_dxy, zmin, zmax, relh_adj_tex, clip_hd1)
    lib().ref_apply_erosion(h.ctypes.data_as(C.c_void_p), xs, ys, min_zval, num_iters, C.byref(p))
    return h


def noise3d_rdata(rs1, rs2, mag, freq):
    out = np.empty(420, np.float32)
    lib().ref_noise3d_rdata(rs1, rs2, mag, freq, out.ctypes.data_as(C.c_void_p))
    return out


def voxel_fill(nx, ny, nz, lo_pos, vsz, offset, mag, freq, normalize_to_1, rs1, rs2, gen_mode, zscale):
    out = np.empty((ny, nx, nz), np.float32)
    a = [np.asarray(v, np.float32) for v in (lo_pos, vsz, offset)]
    lib().ref_voxel_fill(nx, ny, nz, *[v.ctypes.data_as(C.c_void_p) for v in a], mag, freq, int(normalize_to_1), rs1, rs2, gen_mode, zscale,
                         out.ctypes.data_as(C.c_void_p))
    return out


def gen_mesh(mesh_xy, erosion_iters=0, erode_amount=1.0, relh_adj_tex=0.0, clip_hd1=0.5):
    """The reference's own gen_mesh(0,0,1) (call setup(..., gen_sine_table=False) first: gen_mesh regenerates the table itself)."""
    out = np.empty((mesh_xy[1], mesh_xy[0]), np.float32)
    z6 = np.empty(6, np.float32)
    p = Erosion(erode_amount, 0.0, 0.0, 0.0, 0.0, relh_adj_tex, clip_hd1)
    lib().ref_gen_mesh(erosion_iters, C.byref(p), out.ctypes.data_as(C.c_void_p), z6.ctypes.data_as(C.c_void_p))
    return out, dict(zip(("zmin", "zmax", "zmax_est", "zbottom", "ztop", "water_plane_z"), (float(v) for v in z6)))


def has_tiled_extract():
    return available() and hasattr(lib(), "ref_tile_create_zvals")


def tile_create_zvals(size, x1, y1, erosion_iters=0, ep=None):
    """The reference's own tile_t::create_zvals() (function body cut out of src/tiled_mesh.cpp:467-546 at build time), CPU gen modes 0-2:
    returns (zvals [zvsize, zvsize], dict(sub_zmin[4,4], sub_zmax[4,4], mzmin, mzmax, mesh_dz, radius, wbox=(wx1, wy1, wx2, wy2)))."""
    zv = size + 2
    z = np.empty((zv, zv), np.float32)
    b = np.empty(36, np.float32)
    w = np.empty(4, np.int32)
    rc = lib().ref_tile_create_zvals(size, int(x1), int(y1), int(erosion_iters), None if ep is None else C.byref(ep), z.ctypes.data_as(C.c_void_p),
                                     b.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return z, dict(sub_zmin=b[:16].reshape(4, 4).copy(), sub_zmax=b[16:32].reshape(4, 4).copy(), mzmin=float(b[32]), mzmax=float(b[33]), mesh_dz=float(b[34]),
                   radius=float(b[35]), wbox=tuple(int(v) for v in w))


def tile_ao_lighting(size, x1, y1, zvals, ao_context=None, half_dxy=0.0625):
    """The reference's own tile_t::calc_mesh_ao_lighting() (cut out of src/tiled_mesh.cpp:586-662). ao_context=None: it builds the context
    itself (zvals inside the tile, eval_index outside; CPU gen modes); otherwise the GPU-mode flow with the stored un-eroded ao_zvals."""
    zvals = np.ascontiguousarray(zvals, np.float32)
    assert zvals.shape == (size + 2, size + 2)
    ao = np.empty((size + 1, size + 1), np.uint8)
    cz = None
    if ao_context is not None:
        cz = np.ascontiguousarray(ao_context, np.float32)
        assert cz.shape == (size + 1 + 72, size + 1 + 72)
    rc = lib().ref_tile_ao_lighting(size, int(x1), int(y1), zvals.ctypes.data_as(C.c_void_p), None if cz is None else cz.ctypes.data_as(C.c_void_p),
                                    C.c_float(half_dxy), ao.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return ao


# ---- the reference's own voxel_manager (member functions cut out of src/voxels.cpp at build time) ----
def has_voxel_extract():
    return available() and hasattr(lib(), "ref_vox_init")


class Vox:
    """One voxel_manager of the reference: init(grid + voxel_params_t) then the steps of voxel_model::build (src/voxels.cpp:1496-1530)."""

    def __init__(self, nx, ny, nz, vsz, center, isolevel=0.0, invert=0, make_closed_surface=1, remove_unconnected=1, keep_at_scene_edge=0, atten_at_edges=0,
                 remove_under_mesh=0, use_mesh=0, radius_val=0.5, atten_thresh=1.0):
        p = VoxParams()
        p.nx, p.ny, p.nz = nx, ny, nz
        for d in range(3):
            p.vsz[d], p.center[d] = vsz[d], center[d]
        p.isolevel, p.invert, p.make_closed_surface, p.remove_unconnected = isolevel, invert, make_closed_surface, remove_unconnected
        p.keep_at_scene_edge, p.atten_at_edges, p.remove_under_mesh, p.use_mesh, p.radius_val, p.atten_thresh = keep_at_scene_edge, atten_at_edges, remove_under_mesh, use_mesh, radius_val, atten_thresh
        self.p, self.shape = p, (ny, nx, nz)
        lib().ref_vox_init(C.byref(p))

    @property
    def lo_pos(self):
        lo = np.empty(3, np.float32)
        lib().ref_vox_get_lo_pos(lo.ctypes.data_as(C.c_void_p))
        return lo

    def set_vals(self, v):
        v = np.ascontiguousarray(v, np.float32)
        assert v.shape == self.shape
        lib().ref_vox_set_vals(v.ctypes.data_as(C.c_void_p))

    def vals(self):
        v = np.empty(self.shape, np.float32)
        lib().ref_vox_get_vals(v.ctypes.data_as(C.c_void_p))
        return v

    def outside(self):
        o = np.empty(self.shape, np.uint8)
        lib().ref_vox_get_outside(o.ctypes.data_as(C.c_void_p))
        return o

    def create_procedural(self, mag, freq, offset, normalize_to_1, rs1, rs2, gen_mode):
        off = np.asarray(offset, np.float32)
        lib().ref_vox_create_procedural(mag, freq, off.ctypes.data_as(C.c_void_p), int(normalize_to_1), rs1, rs2, gen_mode)

    def atten(self, mode, val, radius=0.5):
        lib().ref_vox_atten(mode, val, radius)

    def set_zmin_matrix(self, mesh):
        lib().ref_vox_set_zmin_matrix(None if mesh is None else np.ascontiguousarray(mesh, np.float32).ctypes.data_as(C.c_void_p))

    def zix(self):
        z = np.empty((self.shape[0], self.shape[1]), np.uint32)
        lib().ref_vox_zix(z.ctypes.data_as(C.c_void_p))
        return z

    def determine_outside(self):
        lib().ref_vox_determine_outside()

    def remove_unconnected(self):
        lib().ref_vox_remove_unconnected()

    def remove_interior_holes(self):
        lib().ref_vox_remove_interior_holes()

    def triangles(self, welded=False, want_counts=False):
        """add_triangles_for_voxel over the grid: (tris [n, 3, 3] float32, per-voxel count_only counts or None)."""
        counts = np.empty(self.shape, np.uint32) if want_counts else None
        n = lib().ref_vox_triangles(int(welded), None, 0, None if counts is None else counts.ctypes.data_as(C.c_void_p))
        tris = np.empty((n, 3, 3), np.float32)
        lib().ref_vox_triangles(int(welded), tris.ctypes.data_as(C.c_void_p), n, None)
        return tris, counts


def mc_tables():
    e, t, v = np.empty(256, np.uint32), np.empty((256, 16), np.int32), np.empty((12, 2), np.uint32)
    lib().ref_vox_tables(e.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    return e, t, v


def has_shadow_extract():
    return available() and hasattr(lib(), "ref_calc_mesh_shadows")


def calc_mesh_shadows(lpos, mh, zmin, zmax, sh_in_x=None, sh_in_y=None):
    """The reference's own calc_mesh_shadows (LIGHT_SUN) on one tile, scene constants from setup(); run with ref_set_threads(1). Returns (smask, sh_out_x, sh_out_y)."""
    mh = np.ascontiguousarray(mh, np.float32)
    ys, xs = mh.shape
    smask = np.empty((ys, xs), np.uint8)
    ox, oy = np.full(xs, -1.0e6, np.float32), np.full(ys, -1.0e6, np.float32)
    lp = np.asarray(lpos, np.float32)
    six = None if sh_in_x is None else np.ascontiguousarray(sh_in_x, np.float32)
    siy = None if sh_in_y is None else np.ascontiguousarray(sh_in_y, np.float32)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    lib().ref_calc_mesh_shadows(p(lp), p(mh), p(smask), xs, ys, zmin, zmax, p(six), p(siy), p(ox), p(oy))
    return smask, ox, oy


def has_texture_extract():
    return available() and hasattr(lib(), "ref_tile_create_texture")


def tex_ids():
    """The engine's texture ids in the order {sand, dirt, ground (grass), rock, snow}."""
    out = (C.c_int * 5)()
    lib().ref_tex_ids(out)
    return [int(v) for v in out]


def tile_create_texture(size, x1, y1, zvals, params8, h_dirt, tex_id_order, vegetation, relh_adj_tex, zmin, zmax, snow_to_rock=0):
    """The reference's own tile_t::create_texture (terrain-only path: no cities / buildings / tunnels / trees) on caller-provided zvals; scene constants
    (DX_VAL, mesh_gen_mode / shape, sine tables, water level) from setup(). params8 = biome corners [y][x]{grass, dirt}; h_dirt / tex_id_order = the
    texture-height table. Returns (weights [stride, stride, 4] uint8 = {sand, dirt, grass, rock}, has_any_grass)."""
    L = lib()
    zv = np.ascontiguousarray(zvals, np.float32)
    assert zv.shape == (size + 2, size + 2)
    st = size + 1
    out = np.empty((st, st, 4), np.uint8)
    p8 = np.ascontiguousarray(params8, np.float32).reshape(8)
    hd = np.ascontiguousarray(h_dirt, np.float32).reshape(5)
    ids = (C.c_int * 5)(*[int(v) for v in tex_id_order])
    hag = C.c_int(0)
    L.ref_tile_create_texture.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    rc = L.ref_tile_create_texture(size, x1, y1, zv.ctypes.data, p8.ctypes.data, hd.ctypes.data, C.cast(ids, C.c_void_p), vegetation, relh_adj_tex, zmin, zmax, int(snow_to_rock), out.ctypes.data, C.cast(C.byref(hag), C.c_void_p))
    if rc != 0:
        raise RuntimeError("ref_tile_create_texture failed: %d" % rc)
    return out, int(hag.value)


def init_terrain_mesh(water_h_off_rel=0.0, temperature=20.0, glaciate_exp=3.0):
    """The reference's init_terrain_mesh() (linked) + gen_tex_height_tables() (cut out of src/Textures.cpp at build time): (h_dirt[5], texture ids[5], lttex zvals[5], clip_hd1)."""
    L = lib()
    h, ids, zv, clip = (C.c_float * 5)(), (C.c_int * 5)(), (C.c_float * 5)(), C.c_float()
    L.ref_init_terrain_mesh.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_init_terrain_mesh(water_h_off_rel, temperature, glaciate_exp, C.cast(h, C.c_void_p), C.cast(ids, C.c_void_p), C.cast(zv, C.c_void_p), C.cast(C.byref(clip), C.c_void_p))
    return np.array(h, np.float32), [int(v) for v in ids], np.array(zv, np.float32), np.float32(clip.value)
