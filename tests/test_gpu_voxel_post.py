"""GPU parity of the voxel post-processing: tw_voxel_outside, tw_voxel_remove_unconnected, tw_voxel_triangles vs the committed
golden outputs of the reference's own voxel_manager functions and vs the CPU oracle (itself pinned against them, tests/test_oracle_vs_reference.py)."""
import os

import numpy as np
import pytest

from cases import convert

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _params(mod, a):
    p = mod.VoxelPostParams()
    p.nx, p.ny, p.nz = int(a[0]), int(a[1]), int(a[2])
    for d in range(3):
        p.lo_pos[d], p.vsz[d] = float(a[3 + d]), float(a[6 + d])
    p.isolevel, p.invert, p.make_closed_surface, p.remove_unconnected, p.keep_at_edge, p.centre_seed, p.skip_under_mesh = float(a[9]), int(a[10]), int(a[11]), int(a[12]), int(a[13]), int(a[14]), int(a[15])
    return p


@pytest.mark.parametrize("name", ["sine", "inv", "mesh"])
def test_voxel_post_golden(tw, ctx, beq, name):
    import torch
    g = np.load(os.path.join(GOLD, "voxel_post.npz"))
    tables = (g["edge_table"], g["tri_table"], g["edge_to_vals"])
    vpp = _params(tw, g[name + "_params"])
    vals = g[name + "_vals"]
    zix = g[name + "_zix"] if (name + "_zix") in g.files else None
    out = ctx.voxel_outside(vals, vpp, zix)
    assert np.array_equal(out, g[name + "_outside"])
    v2, o2 = vals.copy(), out.copy()
    changed = ctx.voxel_remove_unconnected(v2, o2, vpp)
    assert np.array_equal(o2, g[name + "_outside2"]) and beq(v2, g[name + "_vals2"]) == 0
    assert changed == int((g[name + "_outside2"] != g[name + "_outside"]).sum())
    tris = ctx.voxel_triangles(v2, o2, vpp, tables)
    assert tris.shape == g[name + "_tris"].shape and beq(tris, g[name + "_tris"]) == 0     # same triangles, same order
    # device-resident chain (nothing goes back to the host between the steps)
    dv, dz = torch.from_numpy(vals).cuda(), (None if zix is None else torch.from_numpy(zix.astype(np.int32)).cuda())
    do = torch.empty(vals.shape, dtype=torch.uint8, device="cuda")
    ctx.voxel_outside(dv, vpp, dz, out=do)
    ctx.voxel_remove_unconnected(dv, do, vpp)
    dt = torch.empty((len(tris), 3, 3), dtype=torch.float32, device="cuda")
    _, n = ctx.voxel_triangles(dv, do, vpp, tables, out=dt)
    assert n == len(tris) and beq(dt.cpu().numpy(), g[name + "_tris"]) == 0
    small = torch.zeros((7, 3, 3), dtype=torch.float32, device="cuda")               # capacity smaller than the result: the count is still reported
    _, n = ctx.voxel_triangles(dv, do, vpp, tables, out=small)
    assert n == len(tris) and beq(small.cpu().numpy(), g[name + "_tris"][:7]) == 0


@pytest.mark.parametrize("dims,seed,kw", [((40, 33, 29), 1, dict(remove_unconnected=3)), ((64, 64, 64), 2, dict(remove_unconnected=3, invert=1, isolevel=0.2, make_closed_surface=0)),
                                          ((130, 70, 50), 3, dict(remove_unconnected=1, keep_at_edge=1, centre_seed=0)), ((17, 19, 23), 4, dict(remove_unconnected=3, centre_seed=0, skip_under_mesh=1))])
def test_voxel_post_vs_oracle_random_fields(tw, oracle, ctx, beq, dims, seed, kw):
    """Smoothed random fields (many components, pockets, long thin connections => deep flood fills) and random under-mesh heights."""
    g = np.load(os.path.join(GOLD, "voxel_post.npz"))
    tables = (g["edge_table"], g["tri_table"], g["edge_to_vals"])
    nx, ny, nz = dims
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((ny, nx, nz)).astype(np.float32)
    for ax in range(3):
        f = (f + np.roll(f, 1, ax) + np.roll(f, -1, ax)) / 3
    vals = np.ascontiguousarray(f * 3, np.float32)
    vpp = tw.VoxelPostParams()
    vpp.nx, vpp.ny, vpp.nz = nx, ny, nz
    for d in range(3):
        vpp.lo_pos[d], vpp.vsz[d] = (-1.0, 0.5, 0.25)[d], (0.05, 0.07, 0.04)[d]
    vpp.isolevel, vpp.invert, vpp.make_closed_surface = kw.get("isolevel", 0.0), kw.get("invert", 0), kw.get("make_closed_surface", 1)
    vpp.remove_unconnected, vpp.keep_at_edge, vpp.centre_seed, vpp.skip_under_mesh = kw["remove_unconnected"], kw.get("keep_at_edge", 0), kw.get("centre_seed", 1), kw.get("skip_under_mesh", 0)
    zix = None if vpp.centre_seed else rng.integers(0, nz // 2, (ny, nx)).astype(np.uint32)
    po = convert(vpp, oracle.VoxelPostParams)
    out_o = oracle.voxel_outside(vals, po, zix)
    out = ctx.voxel_outside(vals, vpp, zix)
    assert np.array_equal(out, out_o)
    v_o, o_o, ch_o = oracle.voxel_remove_unconnected(vals, out_o, po)
    v2, o2 = vals.copy(), out.copy()
    ch = ctx.voxel_remove_unconnected(v2, o2, vpp)
    assert np.array_equal(o2, o_o) and beq(v2, v_o) == 0 and ch == ch_o
    assert ch_o > 0
    t_o = oracle.voxel_triangles(v_o, o_o, po, tables)
    t = ctx.voxel_triangles(v2, o2, vpp, tables)
    assert t.shape == t_o.shape and beq(t, t_o) == 0 and len(t) > 100
