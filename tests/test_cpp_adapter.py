"""The C++ host adapter (3dworld_b200/host/tw3d_adapter.h: reference class/function signatures on top of the C ABI).
CPU: it compiles, links against lib3dworld_b200.so and fails loudly without a device. GPU: driven like the reference's callers
(heightmap_t::proc_gen, tile_t::create_zvals incl. the no_wait protocol, create_procedural) and compared bit-for-bit with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from cases import convert, HM_CFG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_adapter")


@pytest.fixture(scope="module")
def exe(tw):
    src = os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp")
    hdr = os.path.join(ROOT, "3dworld_b200", "host", "tw3d_adapter.h")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(tw.LIB_PATH)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "3dworld_b200", "host"),
                               src, "-L" + os.path.join(ROOT, "3dworld_b200"), "-l3dworld_b200", "-Wl,-rpath," + os.path.join(ROOT, "3dworld_b200"), "-o", EXE])
    return EXE


def test_adapter_builds_and_refuses_without_device(exe):
    import torch
    out = subprocess.run([exe, "probe"], capture_output=True, text=True)
    assert out.returncode == 0 and "abi 1" in out.stdout
    assert ("device ok" in out.stdout) == torch.cuda.is_available()
    if not torch.cuda.is_available():
        assert "no device: status -1" in out.stdout          # TW_ERR_NO_DEVICE: no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2, 4])
def test_adapter_matches_oracle(exe, tw, scene, oracle, ctx, beq, tmp_path, mode):
    path = str(tmp_path / "out.bin")
    r = subprocess.run([exe, str(mode), path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = np.fromfile(path, np.float32)
    W, H, zv = 160, 96, 66
    sizes = [W * H, W * H, zv * zv, 3 * zv * zv, 24 * 10 * 30, 8] + ([128 * 128, 6] if mode == 0 else [])
    assert data.size == sum(sizes)
    parts = np.split(data, np.cumsum(sizes)[:-1])
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
    hp = convert(cfg.height_params(), oracle.HeightParams)
    sp = cfg.sine_params()
    f32 = np.float32
    # heightmap_t::proc_gen
    z = oracle.heightgen_2d(oracle.Grid2D(-0.5 * W, -0.5 * H, 0.0625, 0.0625, W, H), hp, sp, 1, 0)
    assert beq(parts[0], z) == 0
    mn, mx = f32(z.min()), f32(z.max())
    ep = oracle.ErosionParams(1.0, float(mn - f32(10.0)), 0.0625, float(mn - f32(0.1)), float(mx + f32(0.1)), 0.0, 0.5)
    ze, _ = oracle.apply_erosion(z, float(mn), 700, ep)
    assert beq(parts[1], ze) == 0
    # tile_t::create_zvals (async protocol)
    zt = oracle.heightgen_2d(oracle.Grid2D(float(5 * 64 - 64), float(-3 * 64 - 64), 0.0625, 0.0625, zv, zv), hp, sp, 1, 0)
    assert beq(parts[2], zt) == 0
    # batched tiles + per-tile erosion with min_zval = zmin global
    exp = []
    for x1, y1 in ((0, 0), (64, 0), (0, 64)):
        t = oracle.heightgen_2d(oracle.Grid2D(float(x1 - 64), float(y1 - 64), 0.0625, 0.0625, zv, zv), hp, sp, 1, 0)
        exp.append(oracle.apply_erosion(t, float(mn - f32(0.1)), 300, ep)[0])
    assert beq(parts[3], np.stack(exp)) == 0
    # create_procedural
    vp = oracle.VoxelParams()
    vp.nx, vp.ny, vp.nz = 24, 10, 30
    for d, (lo, vs, off) in enumerate(zip((-7.9, -7.8, -1.5), (0.4, 0.65, 0.11), (0.5, -0.25, 0.0))):
        vp.lo_pos[d], vp.vsz[d], vp.offset[d] = lo, vs, off
    vmode = 1 if mode >= 3 else mode
    vp.mag = vp.freq = 1.0
    vp.gen_mode, vp.normalize_to_1, vp.rseed1, vp.rseed2, vp.octaves = vmode, 1, 123, 456, 3
    vp.rx, vp.ry = oracle.gen_rx_ry(1, 0, vmode) if vmode != 0 else (0.0, 0.0)
    assert beq(parts[4], oracle.voxel_fill(vp)) == 0
    # point queries (get_exact_zval scrolled by xoff2/yoff2 = 100/-40, single-point forms)
    xy = np.array([[0.0, 0.0], [1.5, -2.25], [-3.9, 3.9], [100.0, 250.0], [0.03125, 0.0625]], np.float32)
    PQ = oracle.PointQuery
    exp = list(oracle.eval_points(xy, hp, PQ(2, 1.0, 128, 128, 4.0, 4.0, 100, -40, 0), sp))
    exp += list(oracle.eval_points(xy[1:2], hp, PQ(2, 1.0, 128, 128, 4.0, 4.0, 100, -40, 1), sp))
    exp += list(oracle.eval_points(np.array([[0.3, -7.0]], np.float32), hp, PQ(0, 1.0, 128, 128, 4.0, 4.0, 100, -40, 0), sp))
    exp += list(oracle.eval_points(np.array([[70.0, 12.5]], np.float32), hp, PQ(1, 16.0, 128, 128, 4.0, 4.0, 100, -40, 0), sp))
    assert beq(parts[5], np.array(exp, np.float32)) == 0
    if mode == 0:   # tw3d::gen_mesh == the reference's own gen_mesh() (golden fixture gm_cfg1_eroded)
        h = np.load(os.path.join(ROOT, "tests", "golden", "height.npz"))
        assert beq(parts[6], h["gm_cfg1_eroded"]) == 0
        assert beq(parts[7], h["gm_cfg1_eroded_zvals"]) == 0


@pytest.fixture(scope="module")
def exe_multi(tw):
    src = os.path.join(ROOT, "tests", "cpp", "test_multi.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "test_multi")
    hdr = os.path.join(ROOT, "3dworld_b200", "host", "tw3d_adapter.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(tw.LIB_PATH)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "3dworld_b200", "host"),
                               "-I", "/usr/local/cuda/include", src, "-L" + os.path.join(ROOT, "3dworld_b200"), "-l3dworld_b200", "-L/usr/local/cuda/lib64", "-lcudart",
                               "-Wl,-rpath," + os.path.join(ROOT, "3dworld_b200"), "-Wl,-rpath,/usr/local/cuda/lib64", "-o", out])
    return out


@pytest.mark.gpu
def test_adapter_mesh_shadows_golden(exe, tmp_path, beq):
    """tw3d::calc_mesh_shadows (the reference's calc_mesh_shadows signature, batched over tiles) on the committed reference fixture."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "shadows.npz"))
    tiles, txy, lights, prm = g["tiles"], g["tile_xy"], g["lights"], g["params"].astype(np.float32)
    nt, zv = tiles.shape[0], tiles.shape[1]
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(src, "wb") as f:
        f.write(np.array([nt, zv, len(lights)], np.int32).tobytes() + prm.tobytes() + txy.astype(np.int32).tobytes() + lights.astype(np.float32).tobytes() + tiles.tobytes())
    r = subprocess.run([exe, "shadows", src, dst], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(dst, "rb").read()
    per = nt * zv * zv + 2 * 4 * nt * zv
    assert len(raw) == per * len(lights)
    for li in range(len(lights)):
        blk = raw[li * per:(li + 1) * per]
        m = np.frombuffer(blk[:nt * zv * zv], np.uint8).reshape(nt, zv, zv)
        o = np.frombuffer(blk[nt * zv * zv:], np.float32).reshape(2, nt, zv)
        assert np.array_equal(m, g["smask_%d" % li]), li
        assert beq(o[0], g["sh_out_x_%d" % li]) == 0 and beq(o[1], g["sh_out_y_%d" % li]) == 0, li


def test_multi_gpu_adapter_builds(exe_multi):
    assert os.path.exists(exe_multi)


@pytest.mark.gpu
def test_multi_gpu_adapter_all_devices(exe_multi):
    """tw3d::multi_gpu over every visible device (1 on a single-GPU box: the sharded call then runs without NCCL) == one device."""
    r = subprocess.run([exe_multi, "0"], capture_output=True, text=True)
    assert r.returncode == 0 and "identical to one device" in r.stdout, r.stdout + r.stderr
