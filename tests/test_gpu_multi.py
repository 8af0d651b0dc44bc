"""Multi-GPU host layer (include/tw3d.h "Multi-GPU", csrc/tw_multi.cu): the sharded calls produce exactly what one device produces for
the whole batch (tiles / rows are pure functions of global coordinates), the global z range comes out of the library's own ncclAllReduce,
and the one-process-per-GPU variant reduces across ranks. Cases that need two devices skip on a one-GPU box (gpurun --gpus 2 runs them)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import HM_CFG

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch
    return torch.cuda.device_count()


def _case(scene, mode=4, S=64, side=6):
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(S, S, 1))
    origins = [(tx * S * 5 - 900, ty * S * 3 + 100) for ty in range(side) for tx in range(side)]
    return cfg, origins, S + 2


@pytest.mark.parametrize("ndev", [1, 2, 4])
def test_create_zvals_sharded_equals_one_device(tw, scene, ctx, beq, ndev):
    import torch
    if _ndev() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    cfg, origins, zv = _case(scene)
    hp, ep = cfg.height_params(), cfg.erosion_params()
    dx, dy = float(cfg.dx_val), float(cfg.dy_val)
    ref, mm_ref = ctx.create_zvals_batch(origins, cfg.mesh_size, dx, dy, zv, hp, 150, ep, ep.zmin, want_minmax=True)
    m = tw.Multi(list(range(ndev)))
    try:
        ranges = [tw.multi_range(len(origins), ndev, i) for i in range(ndev)]
        assert ranges[0][0] == 0 and ranges[-1][1] == len(origins) and all(ranges[i][1] == ranges[i + 1][0] for i in range(ndev - 1))
        # (a) NUMA-local pinned host bands
        bands, ptrs = zip(*[m.alloc_host(i, (b - a, zv, zv)) for i, (a, b) in enumerate(ranges)])
        mm, zr = m.create_zvals_sharded(origins, cfg.mesh_size, dx, dy, zv, hp, 150, ep, ep.zmin, list(bands), want_minmax=True)
        got = np.concatenate(bands)
        assert beq(got, ref) == 0 and np.array_equal(mm, mm_ref)
        assert zr == (float(ref.min()), float(ref.max()))
        for p in ptrs:
            m.free_host(p)
        # (b) device-resident bands, no per-tile min/max requested: the range still comes from the devices
        dbands = [torch.empty((b - a, zv, zv), dtype=torch.float32, device="cuda:%d" % i) for i, (a, b) in enumerate(ranges)]
        _, zr = m.create_zvals_sharded(origins, cfg.mesh_size, dx, dy, zv, hp, 150, ep, ep.zmin, dbands)
        assert beq(np.concatenate([d.cpu().numpy() for d in dbands]), ref) == 0 and zr == (float(ref.min()), float(ref.max()))
    finally:
        m.close()


@pytest.mark.parametrize("ndev", [1, 2])
def test_heightgen_2d_sharded_row_bands(tw, scene, ctx, beq, ndev):
    if _ndev() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    cfg = scene.SceneConfig(mesh_gen_mode=4, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3)
    hp = cfg.height_params()
    g = cfg.heightmap_grid(700, 333)
    ref = ctx.heightgen_2d(g, hp)
    m = tw.Multi(list(range(ndev)))
    try:
        ranges = [tw.multi_range(g.ny, ndev, i) for i in range(ndev)]
        bands = [np.empty((b - a, g.nx), np.float32) for a, b in ranges]
        zr = m.heightgen_2d_sharded(g, hp, bands)
        assert beq(np.concatenate(bands), ref) == 0 and zr == (float(ref.min()), float(ref.max()))
    finally:
        m.close()


def test_dist_allreduce_two_ranks(tw):
    """One process per GPU: ranks exchange the NCCL id through a file, each reduces its own (zmin, zmax) with tw_dist_allreduce_minmax."""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    code = r'''
import importlib, os, sys, time
sys.path.insert(0, %r)
tw = importlib.import_module("3dworld_b200")
rank, path = int(sys.argv[1]), sys.argv[2]
if rank == 0:
    uid = tw.dist_unique_id()
    open(path + ".tmp", "wb").write(uid); os.replace(path + ".tmp", path)
else:
    while not os.path.exists(path): time.sleep(0.05)
    uid = open(path, "rb").read()
ctx = tw.Context(rank)
ctx.dist_init(2, rank, uid)
lo, hi = ctx.dist_allreduce_minmax(-1.0 - rank, 5.0 + 10 * rank)
assert (lo, hi) == (-2.0, 15.0), (lo, hi)
lo, hi = ctx.dist_allreduce_minmax(3.0 - 7 * rank, 4.0)
assert (lo, hi) == (-4.0, 4.0), (lo, hi)
ctx.close()
print("rank", rank, "ok")
''' % ROOT
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        idf = os.path.join(d, "id")
        procs = [subprocess.Popen([sys.executable, "-c", code, str(r), idf], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


@pytest.mark.parametrize("n,m,iters,sweep,halo", [(300, 200, 4000, 256, 48), (512, 512, 20000, 1000, 64), (130, 97, 777, 50, 44), (200, 260, 3000, 3000, 100), (20, 30, 500, 7, 44)])
def test_erode_sweeps_single_device_equals_oracle(tw, scene, oracle, ctx, beq, n, m, iters, sweep, halo):
    """The coherent batched erosion (frozen map per sweep, 64-bit fixed-point deltas, halo rule) on one GPU == the CPU oracle of the same
    algorithm, bit for bit (integer accumulation makes it independent of the order in which the GPU walks the droplets of a sweep)."""
    from cases import convert
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0)
    z = ctx.heightgen_2d(cfg.heightmap_grid(n, m), cfg.height_params())
    zmin, zmax = float(z.min()), float(z.max())
    for ep in (tw.ErosionParams(1.0, zmin - 10, 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 0.5), tw.ErosionParams(1.0, zmin + 0.2 * (zmax - zmin), 0.0625, zmin - 0.1, zmax + 0.1, 0.0, 2.0)):
        zc, moves = oracle.erode_sweeps(z, zmin, iters, convert(ep, oracle.ErosionParams), sweep, halo)
        zg = z.copy()
        got_moves = ctx.erode_sweeps(zg, zmin, iters, ep, sweep, halo)
        assert beq(zg, zc) == 0, "max abs diff %g" % np.nanmax(np.abs(zg - zc))
        assert got_moves == moves
        assert (zc != z).sum() > 100
    assert np.isfinite(zc).all()     # NaN deltas are dropped by the fixed-point accumulation (the reference's serial order can poison cells)


@pytest.mark.parametrize("nbands", [2, 3, 4, 5, 8])
def test_erode_sweeps_banded_equals_single_band(tw, scene, ctx, beq, nbands):
    """The band decomposition of tw_erode_sweeps_sharded (halo copies, per-sweep exchange of the border deltas, middle bands with two neighbours) with all
    bands on ONE device, so that every band count is covered by the single-GPU suite: same bits as the undivided run."""
    import torch
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0)
    ep = cfg.erosion_params()
    # the 700-wide case has droplets whose own writes make their next position NaN: they must end (rule of the algorithm), not read the map's row 0 from a band
    for (nx, ny, iters, sweep, halo) in ((300, 100 * nbands, 20000, 1024, 44), (333, 140 * nbands + 3, 20000, 4096, 64), (700, max(640, 100 * nbands), 30000, 2048, 44)):
        z = ctx.heightgen_2d(cfg.heightmap_grid(nx, ny), cfg.height_params())
        zmin = float(z.min())
        one = z.copy()
        moves1 = ctx.erode_sweeps(one, zmin, iters, ep, sweep, halo)
        ranges = [tw.multi_range(ny, nbands, i) for i in range(nbands)]
        bands = [z[a:b].copy() for a, b in ranges]
        assert ctx.erode_sweeps_banded(bands, nx, ny, zmin, iters, ep, sweep, halo) == moves1
        assert beq(np.concatenate(bands), one) == 0
        dbands = [torch.from_numpy(z[a:b].copy()).cuda() for a, b in ranges]
        ctx.erode_sweeps_banded(dbands, nx, ny, zmin, iters, ep, sweep, halo)
        assert beq(np.concatenate([d.cpu().numpy() for d in dbands]), one) == 0
        for a, _ in ranges[1:]:
            assert (one[a - 8:a + 8] != z[a - 8:a + 8]).any()       # erosion happened across every border


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_erode_sweeps_sharded_equals_single_device(tw, scene, oracle, ctx, beq, ndev):
    """Row bands over ndev GPUs with one grouped ncclSend/ncclRecv of the border deltas per sweep == the one-GPU run, bit for bit; droplets that
    cross band borders and reach the halo rule are included (halo 44 = the minimum, view 32 + 12)."""
    import torch
    if _ndev() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.0)
    ep = cfg.erosion_params()
    m = tw.Multi(list(range(ndev)))
    try:
        for (nx, ny, iters, sweep, halo) in ((700, max(640, 100 * ndev), 30000, 2048, 44), (1024, max(1536, 140 * ndev), 60000, 8192, 64)):     # bands >= 2*halo + 8 rows
            z = ctx.heightgen_2d(cfg.heightmap_grid(nx, ny), cfg.height_params())
            zmin = float(z.min())
            one = z.copy()
            moves1 = ctx.erode_sweeps(one, zmin, iters, ep, sweep, halo)
            ranges = [tw.multi_range(ny, ndev, i) for i in range(ndev)]
            host_bands = [z[a:b].copy() for a, b in ranges]            # copies: a row slice of a contiguous array is already contiguous, i.e. a view
            moves = m.erode_sweeps_sharded(host_bands, nx, ny, zmin, iters, ep, sweep, halo)
            assert beq(np.concatenate(host_bands), one) == 0 and moves == moves1
            dev_bands = [torch.from_numpy(z[a:b].copy()).to("cuda:%d" % i) for i, (a, b) in enumerate(ranges)]
            m.erode_sweeps_sharded(dev_bands, nx, ny, zmin, iters, ep, sweep, halo)
            assert beq(np.concatenate([d.cpu().numpy() for d in dev_bands]), one) == 0
            border = ranges[0][1]
            assert (one[border - 8:border + 8] != z[border - 8:border + 8]).any()      # erosion did happen across the first border
    finally:
        m.close()
