"""CPU: the arithmetic of the CUDA kernel behind tw_tile_weights_batch (terrain weights texture). The kernel's body is one
`__host__ __device__` function (3dworld_b200/csrc/tw_weights.cuh); tests/cpp/test_weights.cpp compiles it with g++ (no FMA contraction, like the library) and this
test compares it bit for bit with the oracle, which tests/test_oracle_vs_reference.py pins against the reference's own tile_t::create_texture. The GPU test
(tests/test_gpu_weights.py) then only has to show that the device build of the same function agrees."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_weights")


def weight_cases(P, rng, zlo, zhi, size, dx, dy, n=6):
    """Random but plausible parameter sets (shared with the GPU test)."""
    out = []
    for k in range(n):
        wp = P()
        hd = np.sort(rng.uniform(0.15, 0.9, 4)).tolist() + [1.0] if k else [0.40, 0.44, 0.60, 0.75, 1.0]
        order = rng.permutation(5).tolist() if k >= 3 else [0, 1, 2, 3, 4]
        for i in range(5):
            wp.h_dirt[i], wp.tex_class[i] = hd[i], order[i]
            wp.class_ix[order[i]] = i
        wp.sthresh[0][0], wp.sthresh[0][1], wp.sthresh[1][0], wp.sthresh[1][1] = 0.68, 0.86, 0.48, 0.72
        pad = 0.05 * (zhi - zlo)
        wp.zmin, wp.zmax, wp.relh_adj_tex = zlo - pad, zhi + pad, float(rng.uniform(-0.1, 0.1)) if k else 0.0
        wp.water_level = zlo + float(rng.uniform(0.0, 0.3)) * (zhi - zlo)
        wp.noise_scale = np.float32((2.0 if k == 2 else 1.0) * float(np.float32(0.003)))
        wp.vnz_scale = float(np.float32(np.sqrt(2.0))) if k == 4 else 1.0
        wp.vegetation, wp.snow_to_rock = (0.0 if k == 5 else 1.0), int(k == 1)
        wp.dx_val, wp.dy_val, wp.dxdy = dx, dy, float(np.float32(dx) * np.float32(dy))
        wp.xy_mult = np.float32(1.0 / float(np.float32(size)))
        out.append(wp)
    return out


def test_kernel_body_on_the_host_equals_oracle(oracle, tmp_path):
    src = os.path.join(ROOT, "tests", "cpp", "test_weights.cpp")
    hdr = os.path.join(ROOT, "3dworld_b200", "csrc", "tw_weights.cuh")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", src, "-o", EXE])
    rng = np.random.default_rng(11)
    nt, S = 5, 40
    zv, st = S + 2, S + 1
    yy, xx = np.mgrid[0:zv, 0:zv].astype(np.float32)
    z = np.stack([(np.sin(xx * (0.11 + 0.03 * t)) * np.cos(yy * 0.07) * (0.5 + 0.2 * t) + 0.02 * rng.standard_normal((zv, zv))).astype(np.float32) for t in range(nt)])
    z[3, 5:9, 5:9] = z[3, 5, 5]                                  # a flat patch (zero slope)
    rand = rng.uniform(-9.0, 9.0, (nt, st, st)).astype(np.float32)
    tp = rng.uniform(-0.2, 1.3, (nt, 8)).astype(np.float32)
    dx = dy = 0.125
    for wp in weight_cases(oracle.WeightParams, rng, float(z.min()), float(z.max()), S, dx, dy):
        exp, flags = oracle.tile_weights(z, rand, tp, wp)
        inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(inp, "wb") as f:
            f.write(np.array([nt, zv], np.uint32).tobytes() + bytes(wp) + z.tobytes() + rand.tobytes() + tp.tobytes())
        subprocess.check_call([EXE, inp, outp])
        raw = np.fromfile(outp, np.uint8)
        got, gflags = raw[:exp.size].reshape(exp.shape), raw[exp.size:]
        assert np.array_equal(got, exp), int((got != exp).sum())
        assert np.array_equal(gflags, flags)
        assert len(np.unique(exp.reshape(-1, 4), axis=0)) > 50
