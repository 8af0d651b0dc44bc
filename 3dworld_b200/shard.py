"""Multi-GPU sharding of the terrain path. Generation and per-tile erosion need NO data exchange: a tile is a pure function
of its global origin + seed, and the reference erodes every tile on its own (src/tiled_mesh.cpp:515). Ranks therefore own contiguous
blocks of tile rows; the only collective is a 2-float min/max all-reduce for the global z range (get_heightmap_z_range,
src/map_view.cpp:399-407). Works with any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
import numpy as np


def tile_rows_for_rank(n_tile_rows, rank, world):
    """Contiguous block partition of tile rows; the first (n % world) ranks get one extra row."""
    base, rem = divmod(n_tile_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_origins(n_tiles_x, row_lo, row_hi, tile_size, x_origin=0, y_origin=0):
    """(x1, y1) of every tile this rank owns, row-major (tile_t coordinates: x1 = tx*size, src/tiled_mesh.cpp:295-300)."""
    return np.array([(x_origin + tx * tile_size, y_origin + ty * tile_size) for ty in range(row_lo, row_hi) for tx in range(n_tiles_x)], np.int32)


def global_z_range(local_min, local_max, dist=None, device=None):
    """All-reduce of the per-rank z range. dist: torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_min), float(local_max)
    import torch
    t = torch.tensor([-float(local_min), float(local_max)], dtype=torch.float32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return -float(t[0]), float(t[1])


def generate_sharded(gen_tiles, n_tiles_x, n_tiles_y, tile_size, rank, world, dist=None, device=None):
    """Run gen_tiles(origins) -> (tiles[nt, zv, zv], minmax[nt, 2]) on this rank's rows and return (origins, tiles, global z range).
    gen_tiles is the per-rank compute (Context.heightgen_tiles [+ erode_tiles] on the GPU)."""
    lo, hi = tile_rows_for_rank(n_tiles_y, rank, world)
    org = tile_origins(n_tiles_x, lo, hi, tile_size)
    if len(org) == 0:
        return org, None, global_z_range(np.inf, -np.inf, dist, device)
    tiles, mm = gen_tiles(org)
    zr = global_z_range(np.min(mm[:, 0]), np.max(mm[:, 1]), dist, device)
    return org, tiles, zr
