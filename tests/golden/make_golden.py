"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref3dworld.so, built by oracle/refbuild/build_ref.sh from
/root/reference). Run in the build container only (needs /root/reference for mapx/mesh128.txt and the _ref build); the fixtures are
committed so the GPU box and later rounds can check the oracle and the CUDA path without the reference tree.
    python tests/golden/make_golden.py
The reference publishes no golden vectors for this path, so these are outputs of the reference code itself."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refapi as R  # noqa: E402
from cases import HM_ALL, HM_CFG  # noqa: E402

RL = R.lib()
RL.ref_set_threads(1)
rng = np.random.default_rng(1234)

# ---- KATs named in SURVEY.md section 8(c) ----
kat = {}
for mode in (0, 1, 2):
    R.setup(mode=mode, freq_filter=0, seed=0, glaciate=0)
    z = R.heightgen(0, 0, RL.ref_get_dx(), RL.ref_get_dy(), 128, 128, cache_values=1, glaciate=0)
    kat["eval_index_3_5_mode%d" % mode] = np.float32(z[5, 3])
    if mode == 0:
        kat["sine_params_fresh_process"] = R.sine_params()   # function-static rng in its initial (1,1) state
kat["noise3d_point"] = np.float32(RL.ref_noise3d_point(123, 456, 1.0, 1.0, 0.1, 0.2, 0.3))
np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)

# ---- GLM noise point samples ----
d = {}
for name, k in (("simplex2", 2), ("perlin2", 2), ("simplex3", 3), ("perlin3", 3)):
    pts = (rng.standard_normal((512, k)) * rng.choice([0.5, 3, 50, 1000, 1e5], (512, 1))).astype(np.float32)
    pts[:64] = np.round(pts[:64])
    pts[64:128] = np.round(pts[64:128] * 2) / 2
    f = getattr(RL, "ref_glm_" + name)
    d[name + "_in"] = pts
    d[name + "_out"] = np.array([f(*[float(v) for v in p]) for p in pts], np.float32)
np.savez_compressed(os.path.join(HERE, "glm_noise.npz"), **d)

# ---- host-side generators ----
d = {}
for i, (mode, seed, idx, mesh, scene, mhs) in enumerate([(0, 6, 0, (128, 128, 1), (4., 4., 4.), 0.7), (1, 0, 3, (128, 128, 1), (4., 4., 4.), 1.0),
                                                       (2, 1, 0, (256, 128, 1), (8., 4., 4.), 1.0), (4, 1, 0, (128, 128, 1), (4., 4., 4.), 1.0)]):
    R.setup(mesh=mesh, scene=scene, mode=mode, seed=seed, rgen_index=idx, mesh_height_scale=mhs)
    d["sp%d_args" % i] = np.array([mode, seed, idx, mesh[0], mesh[1], scene[0], scene[1], mhs], np.float64)
    d["sp%d" % i] = R.sine_params()
    d["rxry%d" % i] = np.array(R.rx_ry(), np.float32)
d["rdata_123_456"] = R.noise3d_rdata(123, 456, 1.0, 1.0)
d["rdata_7_9"] = R.noise3d_rdata(7, 9, 2.5, 0.3)
d["sin_table"] = R.sin_table()
np.savez_compressed(os.path.join(HERE, "host_tables.npz"), **d)

# ---- height grids (BASELINE config 1/2 parameters at fixture size) ----
d = {}
names = []
for mode in (0, 1, 2, 3, 4):
    for shape, ff, hmap, gl in ((0, 1, HM_CFG, 1), (1, 2, HM_ALL, 1), (2, 0, {}, 0)):
        R.setup(mode=mode, shape=shape, freq_filter=ff, seed=1, glaciate=gl, hmap=hmap, zmax_est=2.3)
        nx, ny = (40, 28) if mode == 4 else (56, 44)
        x0, y0 = -4096.0 + 100 * mode, -4096.0 + 37 * shape
        z = R.heightgen(x0, y0, RL.ref_get_dx(), RL.ref_get_dy(), nx, ny, cache_values=0, glaciate=1)
        n = "h_m%d_s%d" % (mode, shape)
        names.append(n)
        d[n] = z
        d[n + "_args"] = np.array([mode, shape, ff, gl, x0, y0, nx, ny], np.float64)
        d[n + "_hmap"] = np.array([hmap.get(k, dflt) for k, dflt in zip(R.HMAP_FIELDS, R.HMAP_DEFAULT)], np.float32)
        if mode == 0:
            d[n + "_sp"] = R.sine_params()
# config 1: 128x128 sine mesh, mesh_seed 6, glaciate, mesh_freq_filter 2, mesh_height 0.7
R.setup(mode=0, freq_filter=2, seed=6, glaciate=1, mesh_height_scale=0.7, zmax_est=0.5)
d["cfg1_sp"] = R.sine_params()
d["cfg1"] = R.heightgen(-64, -64, RL.ref_get_dx(), RL.ref_get_dy(), 128, 128, cache_values=0, glaciate=1)
# the reference's own gen_mesh() (ground mode): BASELINE config 1 = 128x128 sine mesh, mesh_seed 6, glaciate, mesh_freq_filter 2, mesh_height 0.7,
# with and without erosion; plus a simplex-mode mesh. zvals = zmin, zmax, zmax_est, zbottom, ztop, water_plane_z after the call.
for name, (mode, seed, ff, hmap, iters, mhs) in (("gm_cfg1", (0, 6, 2, {}, 0, 0.7)), ("gm_cfg1_eroded", (0, 6, 2, HM_CFG, 2000, 0.7)), ("gm_simplex", (1, 3, 1, HM_CFG, 500, 1.0))):
    R.setup(mode=mode, freq_filter=ff, seed=seed, glaciate=1, mesh_height_scale=mhs, hmap=hmap, gen_sine_table=False)
    m, z6 = R.gen_mesh((128, 128), erosion_iters=iters)
    d[name] = m
    d[name + "_zvals"] = np.array([z6[k] for k in ("zmin", "zmax", "zmax_est", "zbottom", "ztop", "water_plane_z")], np.float32)
    d[name + "_args"] = np.array([mode, seed, ff, iters, mhs, hmap.get("sine_mag", 0.0), hmap.get("sine_freq", 0.0), hmap.get("sine_bias", 0.0)], np.float64)
np.savez_compressed(os.path.join(HERE, "height.npz"), **d)

# ---- erosion ----
d = {}
R.setup(mode=1, freq_filter=1, seed=1, zmax_est=2.0, hmap=HM_CFG)
z = R.heightgen(-48, -48, RL.ref_get_dx(), RL.ref_get_dy(), 96, 96, 0, 1)
zmin, zmax = float(z.min()), float(z.max())
d["in0"] = z
for i, (wpz, clip, ea, iters) in enumerate(((zmin - 10, 0.5, 1.0, 600), ((zmin + zmax) / 2, 0.3, 1.0, 600), (zmin - 10, -1.0, 0.5, 300))):
    kw = dict(erode_amount=ea, water_plane_z=wpz, half_dxy=0.0625, zmin=zmin - 0.1, zmax=zmax + 0.1, relh_adj_tex=0.0, clip_hd1=clip)
    d["out0_%d" % i] = R.apply_erosion(z, zmin, iters, **kw)
    d["args0_%d" % i] = np.array([zmin, iters, ea, wpz, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, clip], np.float64)
txt = open("/root/reference/mapx/mesh128.txt").read().split()   # stored 128x128 heightfield (BASELINE config 1 erosion input)
nx, ny = int(txt[0]), int(txt[1])
m = np.array(txt[2:2 + nx * ny], np.float32).reshape(ny, nx)
mn, mx = float(m.min()), float(m.max())
wpz = mn + 0.3 * (mx - mn)
d["mesh128_in"] = m
d["mesh128_out"] = R.apply_erosion(m, mn, 5000, water_plane_z=wpz, zmin=mn, zmax=mx, clip_hd1=0.5)
d["mesh128_args"] = np.array([mn, 5000, 1.0, wpz, 0.0625, mn, mx, 0.0, 0.5], np.float64)
np.savez_compressed(os.path.join(HERE, "erosion.npz"), **d)

# ---- voxels ----
d = {}
lo, vsz, off = (-7.9, -7.8, -1.5), (0.4, 0.65, 0.11), (0.5, -0.25, 0.0)
for mode in (0, 1, 2):
    R.setup(mode=mode, freq_filter=2, seed=3)
    d["v%d_rxry" % mode] = np.array(R.rx_ry(), np.float32)
    d["v%d" % mode] = R.voxel_fill(20, 12, 28, lo, vsz, off, 1.0, 1.0, 1, 123, 456, mode, 0.0)
    d["v%d_unclamped" % mode] = R.voxel_fill(20, 12, 28, lo, vsz, off, 1.0, 1.0, 0, 123, 456, mode, 0.01)
d["geom"] = np.array(lo + vsz + off, np.float32)
np.savez_compressed(os.path.join(HERE, "voxel.npz"), **d)
print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
