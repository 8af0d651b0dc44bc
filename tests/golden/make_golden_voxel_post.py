"""Generates tests/golden/voxel_post.npz. Build container only; fixtures committed.
Every array is the output of the reference's OWN voxel_manager member functions (create_procedural, determine_voxels_outside,
remove_unconnected_outside, remove_interior_holes, add_triangles_for_voxel - cut out of src/voxels.cpp at build time by
oracle/refbuild/build_ref.sh) plus the marching-cubes tables of src/marching_cubes.h that the product takes as an input.
    python tests/golden/make_golden_voxel_post.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refapi as R  # noqa: E402

d = {}
e, t, v = R.mc_tables()
d["edge_table"], d["tri_table"], d["edge_to_vals"] = e, t, v
cases = {"sine": dict(dims=(28, 24, 20), gen=0, kw=dict(remove_unconnected=3)),
         "inv": dict(dims=(18, 22, 26), gen=1, kw=dict(remove_unconnected=3, invert=1, isolevel=0.1, keep_at_scene_edge=1)),
         "mesh": dict(dims=(26, 24, 20), gen=0, kw=dict(remove_unconnected=2, use_mesh=1))}
for name, c in cases.items():
    R.setup(mode=c["gen"], freq_filter=2, seed=1)
    nx, ny, nz = c["dims"]
    vsz, center = (0.15, 0.12, 0.1), (0.0, 0.0, 0.3)
    V = R.Vox(nx, ny, nz, vsz, center, **c["kw"])
    V.create_procedural(1.0, 1.3, (0.2, 0.1, -0.3), 1, 123, 456, c["gen"])
    kw = c["kw"]
    params = [nx, ny, nz, *V.lo_pos, *vsz, kw.get("isolevel", 0.0), kw.get("invert", 0), kw.get("make_closed_surface", 1), kw["remove_unconnected"],
              int(kw.get("keep_at_scene_edge", 0) == 1), int(not kw.get("use_mesh", 0)), 0]
    d[name + "_params"] = np.array(params, np.float64)
    d[name + "_vals"] = V.vals()
    if kw.get("use_mesh"):
        V.set_zmin_matrix(np.fromfunction(lambda y, x: 0.25 * np.sin(x * 0.3) + 0.2 * np.cos(y * 0.2) + 0.2, (128, 128)).astype(np.float32))
        d[name + "_zix"] = V.zix()
    V.determine_outside()
    d[name + "_outside"] = V.outside()
    V.remove_unconnected()
    if kw["remove_unconnected"] > 2:
        V.remove_interior_holes()
    d[name + "_outside2"], d[name + "_vals2"] = V.outside(), V.vals()
    d[name + "_tris"], _ = V.triangles(welded=False)
    V.set_zmin_matrix(None)
    print(name, d[name + "_tris"].shape, int((d[name + "_outside2"] != d[name + "_outside"]).sum()), "flags changed")
np.savez_compressed(os.path.join(HERE, "voxel_post.npz"), **d)
