"""CPU, world_size 2 over gloo: the rank sharding of the tile grid + the z-range all-reduce (the only collective of the path).
The per-rank compute is injected; here it is the CPU oracle standing in for Context.heightgen_tiles (no GPU in this container)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dworld_b200.shard")
    scene = importlib.import_module("3dworld_b200.scene")
    import oracle as O
    from cases import convert, HM_CFG
    cfg = scene.SceneConfig(mesh_gen_mode=1, mesh_freq_filter=1, mesh_seed=1, hmap=HM_CFG, zmax_est=2.3, mesh_size=(32, 32, 1))
    hp = convert(cfg.height_params(), O.HeightParams)
    S, zv = 32, 34

    def gen(origins):
        tiles = np.stack([O.heightgen_2d(O.Grid2D(float(x1 - S // 2), float(y1 - S // 2), float(cfg.dx_val), float(cfg.dy_val), zv, zv), hp, None, 1, 0, 1)
                          for x1, y1 in origins])
        return tiles, np.stack([tiles.min(axis=(1, 2)), tiles.max(axis=(1, 2))], axis=1)

    org, tiles, zr = shard.generate_sharded(gen, 3, 5, S, rank, world, dist)
    q.put((rank, org, tiles, zr))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_exact():
    import importlib
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("3dworld_b200.shard")
    for n in (1, 5, 8, 256):
        for w in (1, 2, 3, 8):
            spans = [shard.tile_rows_for_rank(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert shard.tile_rows_for_rank(256, 3, 8) == (96, 128)     # BASELINE config 5: 256 tile rows over 8 GPUs = 32 rows each


def test_two_ranks_cover_the_grid_and_agree_on_z_range():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, org0, t0, zr0), (r1, org1, t1, zr1) = res
    assert len(org0) == 9 and len(org1) == 6                      # 5 tile rows -> 3 + 2
    allorg = np.concatenate([org0, org1])
    assert len({tuple(o) for o in allorg}) == 15 and allorg[:, 1].max() == 4 * 32
    assert zr0 == zr1                                             # every rank sees the same global z range
    lo = min(t0.min(), t1.min())
    hi = max(t0.max(), t1.max())
    assert zr0 == (pytest.approx(float(lo)), pytest.approx(float(hi)))
    # neighbouring rows owned by different ranks overlap by two cells and agree bit for bit: no halo exchange is needed
    assert np.array_equal(t0[6][32:34, :], t1[0][0:2, :])
