"""CPU: the product's host-side code (no GPU needed): C-ABI exports, loud failure without a device, host table generators vs the golden
fixtures, the scene-config mirror, and the rank sharding used by bench.py."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(tw):
    hdr = open(os.path.join(ROOT, "include", "tw3d.h")).read()
    declared = sorted(set(re.findall(r"TW_API[^;(]*?\b(tw_\w+)\s*\(", hdr)))
    assert declared == sorted(tw.ABI_SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", tw.LIB_PATH], text=True)
    exported = set(re.findall(r" T (tw_\w+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert tw.lib.tw_abi_version() == 1


def test_no_cpu_fallback(tw):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert tw.lib.tw_create(0, C.byref(h)) == tw.TW_ERR_NO_DEVICE and not h.value
    with pytest.raises(tw.TwError):
        tw.Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "3dworld_b200")):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "terrain_oracle" not in src and "refapi" not in src and "oracle." not in src.replace("oracle.py", ""), os.path.join(dirpath, f)


def test_host_generators_match_golden(tw, beq):
    t = np.load(os.path.join(GOLD, "host_tables.npz"))
    assert beq(tw.build_sin_table(), t["sin_table"]) == 0
    for i in range(4):
        mode, seed, idx, mx, my, sx, sy, mhs = t["sp%d_args" % i]
        sh = np.float32(0.1) * np.float32(4.0) * np.float32(mhs)
        sp = tw.gen_sine_params(sh, mesh=(int(mx), int(my)), scene=(float(sx), float(sy)), seed=int(seed), rgen_index=int(idx), mode=int(mode))
        assert beq(sp, t["sp%d" % i]) == 0
        assert beq(np.array(tw.gen_rx_ry(int(seed), int(idx), int(mode)), np.float32), t["rxry%d" % i]) == 0
    assert beq(tw.noise3d_gen_sines(123, 456, 1.0, 1.0), t["rdata_123_456"]) == 0
    assert beq(tw.noise3d_gen_sines(7, 9, 2.5, 0.3), t["rdata_7_9"]) == 0
    k = np.load(os.path.join(GOLD, "kat.npz"))
    rng = tw.Rng(1, 1)   # the reference's function-static generator in a fresh process
    assert beq(tw.gen_sine_params(np.float32(0.4), rng=rng), k["sine_params_fresh_process"]) == 0
    assert (rng.rseed1, rng.rseed2) != (1, 1)   # state advances: a second call yields a different table, as in the reference
    assert beq(tw.gen_sine_params(np.float32(0.4), rng=rng), k["sine_params_fresh_process"]) != 0


def test_host_generators_match_oracle(tw, oracle, beq):
    for ms in (0.3, 1.0, 2.0, 4.0, 64.0, 1000.0):
        for ff in range(-2, 8):
            assert tw.compute_scale(ms, ff) == oracle.compute_scale(ms, ff)
    assert tw.compute_scale(1.0, 1) == 10     # 8 octaves
    for seed, idx, mode in ((0, 0, 0), (0, 5, 1), (7, 0, 4), (-3, 2, 2)):
        assert tw.gen_rx_ry(seed, idx, mode) == oracle.gen_rx_ry(seed, idx, mode)
    for args in ((1.0, 1, 0.0, 0.0, 0.0), (2.3, 1, 2.5, 0.1, 0.05), (0.5, 0, 0.0, 0.0, 0.7)):
        assert tw.water_z_height(*args) == oracle.lib().to_water_z_height(*args)


def test_scene_config(tw, scene):
    cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, zmax_est=2.3)
    hp = cfg.height_params()
    assert float(cfg.dx_val) == 0.0625 and hp.dx_val_inv == 16.0 and hp.start_eval_sin == 10
    assert abs(hp.mesh_height - 0.4) < 1e-7
    g = cfg.heightmap_grid(8192, 8192)
    assert (g.x0, g.y0, g.nx, g.ny) == (-4096.0, -4096.0, 8192, 8192)
    ep = cfg.erosion_params()
    assert ep.zmin == pytest.approx(-2.3) and 0.0 < ep.clip_hd1 < 1.0
    vp = scene.voxel_landscape_params(scene.SceneConfig(scene_size=(16.0, 16.0, 4.0), mesh_size=(128, 128, 64)), 512, 512, 512)
    assert vp.nx == 512 and vp.rseed2 == 456 and vp.vsz[0] > 0


def test_division_free_forms():
    """csrc/tw_noise.cuh replaces glm::mod(a,289)'s true division and perlin's i/41 by division-free sequences; they must equal the IEEE
    results for EVERY input in their guarded range (numpy fp32 arithmetic is IEEE round-to-nearest; a - q*289 is exact so no fma is needed)."""
    f32 = np.float32
    a = np.arange(-(1 << 22) + 1, 1 << 22, dtype=np.int64).astype(np.float32)
    q = np.floor(a * (f32(1.0) / f32(289.0)))
    r = a - q * f32(289.0)
    r = np.where(r >= f32(289.0), r - f32(289.0), r)
    ref = a - f32(289.0) * np.floor(a / f32(289.0))
    assert np.array_equal(r, ref) and not np.signbit(r).any()
    from fractions import Fraction

    def fma(x, y, z):   # correctly rounded fp32 fma via exact rationals
        v = Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z))
        c = f32(float(v))
        cands = [np.nextafter(c, f32(-np.inf)), c, np.nextafter(c, f32(np.inf))]
        return min(cands, key=lambda t: (abs(Fraction(float(t)) - v), int(f32(t).view(np.uint32)) & 1))
    c41 = f32(1.0) / f32(41.0)
    for i in range(289):
        fi = f32(i)
        q0 = fi * c41
        assert fma(fma(-q0, f32(41.0), fi), c41, q0) == fi / f32(41.0)


def test_hash_table_arguments_are_exact():
    """csrc/tw_noise2.cuh tabulates the Ashima hash: `permute(k)` is looked up for k <= 290, the gradient entry is indexed by the argument of the
    next permute (k <= 578), and mod(i, 289) is left 'lazy' (289 may stand for 0). All of that rests on three facts about the reference's fp32
    arithmetic, checked here exhaustively with numpy (IEEE round-to-nearest fp32):
      1. permute(k) = mod289((34k + 1)k) in fp32 equals the exact integer (34k^2 + k) mod 289 for every k in [0, 580): no rounding anywhere;
      2. hence permute(k + 289) == permute(k), so a lazy 289 (or 290) hashes like 0 (or 1);
      3. the lazy remainder a - 289*floor(a*RN(1/289)) lies in [0, 289] for |a| < 2^22 and is 289 only for multiples of 289;
         the 3-D kernels use the same multiply form directly (glm mod289) and guard |a| < 2^20."""
    f32 = np.float32
    k = np.arange(0, 580, dtype=np.int64)
    kf = k.astype(f32)
    prod = (kf * f32(34.0) + f32(1.0)) * kf                                   # two roundings in the reference; exact because < 2^24
    assert np.array_equal(prod.astype(np.int64), (34 * k + 1) * k) and prod.max() < 2 ** 24
    t = np.floor(prod * (f32(1.0) / f32(289.0)))
    perm = prod - t * f32(289.0)                                              # fma(-t, 289, prod) == this: both exact
    assert np.array_equal(perm.astype(np.int64), ((34 * k + 1) * k) % 289)
    assert np.array_equal(perm[289:578], perm[0:289])                         # permute(k + 289) == permute(k)
    assert perm.max() <= 288 and perm.min() >= 0
    # arguments reachable in the 2-D kernels: q + ix + 1 with q <= 288, ix <= 289 (lazy) -> <= 578 < table size 580
    assert 288 + 289 + 1 < 580
    a = np.arange(-(1 << 22) + 1, 1 << 22, dtype=np.int64)
    af = a.astype(f32)
    r = af - np.floor(af * (f32(1.0) / f32(289.0))) * f32(289.0)
    assert r.min() >= 0 and r.max() <= 289
    assert np.array_equal(r == 289, (a % 289 == 0) & (r != 0))
    assert np.array_equal(r.astype(np.int64) % 289, a % 289)
    # table offsets: k*128 + 1.5*2^23 is exact in fp32 and its bit pattern is 0x4B400000 + 128k
    off = (kf * f32(128.0) + f32(12582912.0)).view(np.uint32).astype(np.int64)
    assert np.array_equal(off, 0x4B400000 + 128 * k)


def test_ctypes_mirrors_match_the_header_layout(tw, oracle, tmp_path):
    """The ctypes structures of the binding (and of the oracle's wrapper) must have the size and field offsets the C compiler gives the PODs
    of include/tw3d.h - a drifted mirror would silently shift every parameter after it."""
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pods = {"tw_hmap_params": "HmapParams", "tw_height_params": "HeightParams", "tw_grid2d": "Grid2D", "tw_minmax": "MinMax", "tw_erosion_params": "ErosionParams",
            "tw_voxel_params": "VoxelParams", "tw_tile_bounds": "TileBounds", "tw_heightmap_info": "HeightmapInfo", "tw_rng": "Rng",
            "tw_point_query": "PointQuery", "tw_hmap_sampler": "HmapSampler"}
    lines = ['#include <tw3d.h>', '#include <stdio.h>', '#include <stddef.h>', 'int main(void) {']
    for c_name, py_name in pods.items():
        cls = getattr(tw, py_name)
        lines.append('printf("%s %%zu", sizeof(%s));' % (c_name, c_name))
        for fname, _ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (c_name, fname))
        lines.append('printf("\\n");')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    out = subprocess.check_output([exe], text=True).strip().splitlines()
    assert len(out) == len(pods)
    for line, (c_name, py_name) in zip(out, pods.items()):
        nums = [int(v) for v in line.split()[1:]]
        for mod in (tw, oracle):
            cls = getattr(mod, py_name, None)
            if cls is None:
                continue
            assert C.sizeof(cls) == nums[0], (c_name, mod.__name__)
            assert [getattr(cls, f).offset for f, _ in cls._fields_] == nums[1:], (c_name, mod.__name__)


def test_null_context_is_an_error_not_a_crash(tw):
    """Status codes instead of assert()/exit(): every entry point checks its context first (include/tw3d.h: TW_ERR_ARG = -3)."""
    import ctypes as C
    L = tw.lib
    g, hp = tw.Grid2D(0, 0, 1, 1, 4, 4), tw.HeightParams()
    assert L.tw_heightgen_2d(None, C.byref(g), C.byref(hp), 1, 0, None, None) == tw.TW_ERR_ARG
    assert L.tw_heightgen_2d_poll(None, 0) == tw.TW_ERR_ARG
    assert L.tw_erode(None, None, 4, 4, 0.0, 10, None) == tw.TW_ERR_ARG
    assert L.tw_erode_parallel(None, None, 4, 4, 0.0, 10, None, 0) == tw.TW_ERR_ARG
    assert L.tw_eval_points(None, None, 0, None, None, None) == tw.TW_ERR_ARG
    assert L.tw_tile_normals_batch(None, None, 0, 0, 0.0, 0.0, None, None) == tw.TW_ERR_ARG
    assert L.tw_tile_ao_batch(None, None, None, 0, 0, 0, 0.0, 0.0, 0, None, 0.0, None) == tw.TW_ERR_ARG
    assert L.tw_heightmap_sample_tiles(None, None, None, None, 0, 0, None) == tw.TW_ERR_ARG
    assert L.tw_voxel_fill(None, None, None, None) == tw.TW_ERR_ARG
    assert L.tw_sync(None) == tw.TW_ERR_ARG


def test_table_driven_simplex_equals_glm_on_cpu(oracle):
    """The shipped noise kernel evaluates glm::simplex(vec2) through a 580-entry table (csrc/tw_noise2.cuh, level 3): gradient entry indexed by
    permute(iy') + ix', lazy mod289, first hash from the table. This numpy fp32 emulation of exactly that data flow must agree bit for bit with
    the oracle's literal restatement (pinned against the reference's GLM) - a CPU check of the algorithm, independent of the device tests."""
    f32 = np.float32
    Cx, Cy, Cz, Cw = f32(0.211324865405187), f32(0.366025403784439), f32(-0.577350269189626), f32(0.024390243902439)

    def permute(x):
        v = (x * f32(34.0) + f32(1.0)) * x
        return v - np.floor(v * (f32(1.0) / f32(289.0))) * f32(289.0)

    k = np.arange(580, dtype=np.int64).astype(f32)
    pk = permute(k)
    t = pk * Cw
    X = (t - np.floor(t)) * f32(2.0) - f32(1.0)
    H = np.abs(X) - f32(0.5)
    A0 = X - np.floor(X + f32(0.5))
    Nn = f32(1.79284291400159) - f32(0.85373472095314) * (A0 * A0 + H * H)

    rng = np.random.default_rng(2)
    n = 400000
    v = (rng.standard_normal((n, 2)) * rng.choice([0.5, 3.0, 50.0, 1000.0, 2.0e5], (n, 1))).astype(f32)
    v[:2000] = (np.round(v[:2000] / 289.0) * 289.0).astype(f32)           # lattice points on multiples of 289: the lazy remainder returns 289
    vx, vy = v[:, 0].copy(), v[:, 1].copy()
    s = vx * Cy + vy * Cy
    ix, iy = np.floor(vx + s), np.floor(vy + s)
    tt = ix * Cx + iy * Cx
    x0x, x0y = vx - ix + tt, vy - iy + tt
    gt = x0x > x0y
    i1x, i1y = gt.astype(f32), (~gt).astype(f32)
    x12x, x12y, x12z, x12w = (x0x + Cx) - i1x, (x0y + Cx) - i1y, x0x + Cz, x0y + Cz
    lazy = lambda a: a - np.floor(a * (f32(1.0) / f32(289.0))) * f32(289.0)   # noqa: E731
    ixm, iym = lazy(ix), lazy(iy)
    assert ixm.min() >= 0 and ixm.max() <= 289
    j0 = iym.astype(np.int64)
    q0, q1, q2 = pk[j0], pk[j0 + (~gt)], pk[j0 + 1]                        # first hash from the table (entries 289/290 repeat 0/1)
    k0, k1, k2 = (q0 + ixm).astype(np.int64), (q1 + ixm + i1x).astype(np.int64), (q2 + ixm + f32(1.0)).astype(np.int64)
    assert max(k0.max(), k1.max(), k2.max()) <= 578

    def m4(a, b):
        m = f32(0.5) - (a * a + b * b)
        m = np.where(m < 0, f32(0.0), m).astype(f32)
        m = m * m
        return m * m
    m0, m1, m2 = m4(x0x, x0y) * Nn[k0], m4(x12x, x12y) * Nn[k1], m4(x12z, x12w) * Nn[k2]
    gx = A0[k0] * x0x + H[k0] * x0y
    gy = A0[k1] * x12x + H[k1] * x12y
    gz = A0[k2] * x12z + H[k2] * x12w
    got = (((m0 * gx) + (m1 * gy)) + (m2 * gz)) * f32(130.0)
    f = oracle.lib().to_simplex2
    exp = np.array([f(float(a), float(b)) for a, b in zip(vx[:60000], vy[:60000])], f32)
    assert np.array_equal(got[:60000].view(np.uint32), exp.view(np.uint32))


def test_table_driven_simplex3_equals_glm_on_cpu(oracle):
    """Same check for the voxel kernels' glm::simplex(vec3) (csrc/tw_noise2.cuh simplex3_lut): lattice indices from glm's multiply-form mod289
    (which yields exactly 289 for multiples of 289), first permute from the table, second computed, third folded into the gradient entry."""
    f32 = np.float32
    c289 = f32(1.0) / f32(289.0)
    mod289 = lambda x: x - np.floor(x * c289) * f32(289.0)                     # noqa: E731
    permute = lambda x: mod289((x * f32(34.0) + f32(1.0)) * x)                 # noqa: E731
    k = np.arange(580, dtype=np.int64).astype(f32)
    pk = permute(k)
    n_ = f32(0.142857142857)
    nsx, nsy, nsz = n_ * f32(2.0) - f32(0.0), n_ * f32(0.5) - f32(1.0), n_ * f32(1.0) - f32(0.0)
    j = pk - f32(49.0) * np.floor(pk * nsz * nsz)
    x_ = np.floor(j * nsz)
    y_ = np.floor(j - f32(7.0) * x_)
    X, Y = x_ * nsx + nsy, y_ * nsx + nsy
    H = f32(1.0) - np.abs(X) - np.abs(Y)
    sh = -(~(f32(0.0) < H)).astype(f32)                                        # -step(h, 0): glm::step(edge = h, x = 0) = (0 < h) ? 0 : 1
    Px, Py, Pz = X + (np.floor(X) * f32(2.0) + f32(1.0)) * sh, Y + (np.floor(Y) * f32(2.0) + f32(1.0)) * sh, H
    nn = f32(1.79284291400159) - f32(0.85373472095314) * (Px * Px + Py * Py + Pz * Pz)
    Px, Py, Pz = Px * nn, Py * nn, Pz * nn

    rng = np.random.default_rng(4)
    n = 30000
    v = (rng.standard_normal((n, 3)) * rng.choice([0.5, 3.0, 50.0, 1000.0, 5.0e4], (n, 1))).astype(f32)
    v[:1500] = (np.round(v[:1500] / 289.0) * 289.0 + rng.uniform(0, 1, (1500, 3))).astype(f32)   # lattice indices on multiples of 289
    vx, vy, vz = v[:, 0].copy(), v[:, 1].copy(), v[:, 2].copy()
    Cx, Cy = f32(1.0 / 6.0), f32(1.0 / 3.0)
    s = vx * Cy + vy * Cy + vz * Cy
    i0, i1, i2 = np.floor(vx + s), np.floor(vy + s), np.floor(vz + s)
    t = i0 * Cx + i1 * Cx + i2 * Cx
    x0x, x0y, x0z = vx - i0 + t, vy - i1 + t, vz - i2 + t
    step = lambda edge, x: (~(x < edge)).astype(f32)                           # noqa: E731  glm::step
    gx, gy, gz = step(x0y, x0x), step(x0z, x0y), step(x0x, x0z)
    lx, ly, lz = f32(1.0) - gx, f32(1.0) - gy, f32(1.0) - gz
    i1x, i1y, i1z = np.minimum(gx, lz), np.minimum(gy, lx), np.minimum(gz, ly)
    i2x, i2y, i2z = np.maximum(gx, lz), np.maximum(gy, lx), np.maximum(gz, ly)
    x1 = (x0x - i1x + Cx, x0y - i1y + Cx, x0z - i1z + Cx)
    x2 = (x0x - i2x + Cy, x0y - i2y + Cy, x0z - i2z + Cy)
    x3 = (x0x - f32(0.5), x0y - f32(0.5), x0z - f32(0.5))
    i0, i1, i2 = mod289(i0), mod289(i1), mod289(i2)
    assert min(i0.min(), i1.min(), i2.min()) >= 0 and max(i0.max(), i1.max(), i2.max()) <= 289 and (i2 == 289).any()
    tab = lambda a: pk[a.astype(np.int64)]                                     # noqa: E731  first permute from the table (index <= 290)
    q0, q1, q2, q3 = tab(i2), tab(i2 + i1z), tab(i2 + i2z), tab(i2 + f32(1.0))
    k0 = permute(q0 + i1) + i0
    k1 = permute(q1 + i1 + i1y) + i0 + i1x
    k2 = permute(q2 + i1 + i2y) + i0 + i2x
    k3 = permute(q3 + i1 + f32(1.0)) + i0 + f32(1.0)
    ks = [a.astype(np.int64) for a in (k0, k1, k2, k3)]
    assert max(a.max() for a in ks) <= 578

    def corner(kk, xx):
        m = f32(0.6) - (xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2])
        m = np.where(m < 0, f32(0.0), m).astype(f32)
        m = m * m
        d = Px[kk] * xx[0] + Py[kk] * xx[1] + Pz[kk] * xx[2]
        return (m * m) * d
    c0, c1, c2, c3 = corner(ks[0], (x0x, x0y, x0z)), corner(ks[1], x1), corner(ks[2], x2), corner(ks[3], x3)
    got = f32(42.0) * ((c0 + c1) + (c2 + c3))
    f = oracle.lib().to_simplex3
    exp = np.array([f(float(a), float(b), float(c)) for a, b, c in zip(vx, vy, vz)], f32)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_table_driven_perlin2_equals_glm_on_cpu(oracle):
    """And for glm::perlin(vec2) (perlin2_lut): lazy mod289 of the four lattice indices, first permute from the table, gradient entry
    {gx*n, gy*n} indexed by permute(ix') + iy' (second permute folded in)."""
    f32 = np.float32
    c289 = f32(1.0) / f32(289.0)
    lazy = lambda a: a - np.floor(a * c289) * f32(289.0)                       # noqa: E731
    permute = lambda x: lazy((x * f32(34.0) + f32(1.0)) * x)                   # noqa: E731
    k = np.arange(580, dtype=np.int64).astype(f32)
    pk = permute(k)
    q = pk / f32(41.0)                                                         # the device's division-free form equals the IEEE quotient (test_division_free_forms)
    g = (q - np.floor(q)) * f32(2.0) - f32(1.0)
    GY = np.abs(g) - f32(0.5)
    GX = g - np.floor(g + f32(0.5))
    nn = f32(1.79284291400159) - f32(0.85373472095314) * (GX * GX + GY * GY)
    GX, GY = GX * nn, GY * nn
    rng = np.random.default_rng(6)
    n = 40000
    P = (rng.standard_normal((n, 2)) * rng.choice([0.5, 3.0, 50.0, 1000.0, 2.0e5], (n, 1))).astype(f32)
    P[:2000] = (np.round(P[:2000] / 289.0) * 289.0 + rng.uniform(-1, 1, (2000, 2))).astype(f32)
    Px, Py = P[:, 0].copy(), P[:, 1].copy()
    flx, fly = np.floor(Px), np.floor(Py)
    frx, fry = Px - flx, Py - fly
    Pfz, Pfw = frx - f32(1.0), fry - f32(1.0)
    Pix, Piy, Piz, Piw = lazy(flx), lazy(fly), lazy(flx + f32(1.0)), lazy(fly + f32(1.0))
    assert max(Pix.max(), Piz.max(), Piy.max(), Piw.max()) <= 289 and (Piz == 289).any()
    qx, qz = pk[Pix.astype(np.int64)], pk[Piz.astype(np.int64)]
    k00, k10, k01, k11 = [(a + b).astype(np.int64) for a, b in ((qx, Piy), (qz, Piy), (qx, Piw), (qz, Piw))]
    assert max(k00.max(), k10.max(), k01.max(), k11.max()) <= 577
    n00 = GX[k00] * frx + GY[k00] * fry
    n10 = GX[k10] * Pfz + GY[k10] * fry
    n01 = GX[k01] * frx + GY[k01] * Pfw
    n11 = GX[k11] * Pfz + GY[k11] * Pfw
    fade = lambda t: (t * t * t) * (t * (t * f32(6.0) - f32(15.0)) + f32(10.0))   # noqa: E731
    mix = lambda x, y, a: x + a * (y - x)                                      # noqa: E731
    fdx, fdy = fade(frx), fade(fry)
    got = mix(mix(n00, n10, fdx), mix(n01, n11, fdx), fdy) * f32(2.3)
    f = oracle.lib().to_perlin2
    exp = np.array([f(float(a), float(b)) for a, b in zip(Px, Py)], f32)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_denormal_table_addressing_is_exact():
    """The noise kernels turn a small integer k (held in a float) into the shared-memory address of table entry k with ONE fused multiply-add on denormals
    (csrc/tw_noise2.cuh, TW_LUT_DENORM): fma(k, 128*2^-149, A*2^-149) must have the bit pattern 128*k + A for every reachable k (<= 578; 3-D tables the same)
    and every base address A below 2^18 (the lane's copy of entry 0). Checked here in IEEE fp32 on the host: the product k*128 (< 2^17) and the sum with A (< 2^19)
    are exact multiples of 2^-149 below 2^23*2^-149, i.e. representable denormals, so the fused and the unfused evaluation agree and nothing rounds."""
    k = np.arange(0, 600, dtype=np.float32)
    entry = np.array([128], np.uint32).view(np.float32)[0]                  # 128 * 2^-149
    for A in (0, 16, 112, 1024 + 48, 74 * 1024 + 96, (1 << 18) - 16):
        a = np.array([A], np.uint32).view(np.float32)[0]
        unfused = (k * entry + a).astype(np.float32)                        # two roundings, both exact
        fused = np.array([np.float32(np.float64(kk) * np.float64(entry) + np.float64(a)) for kk in k], np.float32)   # exact product and sum in fp64, one rounding
        want = (128 * k.astype(np.uint32) + A).astype(np.uint32)
        assert np.array_equal(unfused.view(np.uint32), want) and np.array_equal(fused.view(np.uint32), want), A
    step = np.float32(1.0) * entry                                          # the middle corner: address of entry k + i1.y = fma(i1.y, 128*2^-149, address of entry k)
    base = (np.float32(17.0) * entry + np.array([4096 + 32], np.uint32).view(np.float32)[0]).astype(np.float32)
    assert (base + step).astype(np.float32).view(np.uint32) == 128 * 18 + 4096 + 32 and (base + np.float32(0.0) * entry).astype(np.float32).view(np.uint32) == 128 * 17 + 4096 + 32
