"""GPU parity: tw_voxel_fill (CUDA through the C ABI) vs the CPU oracle - bit-exact, sine and GLM modes, with and without attenuation."""
import numpy as np
import pytest

from cases import convert

pytestmark = pytest.mark.gpu


def _vp(tw, scene, mode, ff, nx, ny, nz, norm=1, zs=0.0, atten=0):
    cfg = scene.SceneConfig(mesh_gen_mode=mode, mesh_freq_filter=ff, mesh_seed=3, scene_size=(16.0, 16.0, 4.0), mesh_size=(128, 128, 64), zmax_est=1.0)
    vp = scene.voxel_landscape_params(cfg, nx, ny, nz, normalize_to_1=norm)
    vp.zscale = zs
    vp.atten_mode, vp.atten_val, vp.atten_inner_radius = atten, 0.7, 0.4
    vp.offset[0], vp.offset[1] = 0.5, -0.25
    return vp


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("dims", [(40, 24, 36), (9, 5, 130), (70, 3, 2)])
def test_voxel_fill_bit_exact(tw, scene, oracle, ctx, beq, mode, dims):
    for ff in (2, 0):
        for norm, zs in ((1, 0.0), (0, 0.01)):
            vp = _vp(tw, scene, mode, ff, *dims, norm=norm, zs=zs)
            got = ctx.voxel_fill(vp)
            exp = oracle.voxel_fill(convert(vp, oracle.VoxelParams))
            assert beq(got, exp) == 0, "max abs diff %g" % np.abs(got - exp).max()


@pytest.mark.parametrize("mode", [1, 2])
def test_voxel_glm_table_domain_edges(tw, scene, oracle, ctx, beq, mode):
    """The GLM kernels read hash/gradient tables while the lattice stays below 2^20 and use the literal arithmetic beyond: grids that sit on
    multiples of 289 (mod289 returns exactly 289 there), straddle the switch-over, or lie far outside it must all match the oracle."""
    for freq, off in ((1.0, (289.0 * 4, -289.0 * 8, 289.0 * 12)), (64.0, (16000.0, 0.0, 0.0)), (1.0, (3.0e6, -2.0e6, 1.0e6)), (1000.0, (0.0, 0.0, 0.0))):
        vp = _vp(tw, scene, mode, 0, 24, 6, 140)
        vp.freq = freq
        for d in range(3):
            vp.offset[d] = off[d]
        got = ctx.voxel_fill(vp)
        exp = oracle.voxel_fill(convert(vp, oracle.VoxelParams))
        assert beq(got, exp) == 0, (freq, off)


@pytest.mark.parametrize("atten", [1, 2, 3, 4, 5])
def test_voxel_atten_modes(tw, scene, oracle, ctx, beq, atten):
    vp = _vp(tw, scene, 0, 2, 33, 20, 48, atten=atten)
    assert beq(ctx.voxel_fill(vp), oracle.voxel_fill(convert(vp, oracle.VoxelParams))) == 0


def test_voxel_explicit_rdata_and_device_out(tw, scene, oracle, ctx, beq):
    import torch
    vp = _vp(tw, scene, 0, 2, 64, 32, 64)
    rd = tw.noise3d_gen_sines(77, 99, 2.0, 0.5)
    out = torch.empty((32, 64, 64), dtype=torch.float32, device="cuda")
    ctx.voxel_fill(vp, rdata=rd, out=out)
    assert beq(out.cpu().numpy(), oracle.voxel_fill(convert(vp, oracle.VoxelParams), rdata=rd)) == 0


def test_voxel_matches_linked_reference(tw, scene, ref, ctx, beq):
    for mode in (0, 1, 2):
        ref.setup(mode=mode, freq_filter=2, seed=3)
        vp = _vp(tw, scene, mode, 2, 40, 24, 36)
        zr = ref.voxel_fill(40, 24, 36, list(vp.lo_pos), list(vp.vsz), list(vp.offset), 1.0, 1.0, 1, vp.rseed1, vp.rseed2, mode, 0.0)
        assert beq(ctx.voxel_fill(vp), zr) == 0


def test_voxel_512_cube_properties(tw, scene, oracle, ctx, beq):
    """BASELINE config 4 (512^3 sine): a z-column slab recomputed as a smaller grid at the same positions matches; an oracle-sized block matches."""
    import torch
    vp = _vp(tw, scene, 0, 2, 512, 512, 512)
    out = torch.empty((512, 512, 512), dtype=torch.float32, device="cuda")
    ctx.voxel_fill(vp, out=out)
    assert torch.isfinite(out).all() and out.abs().max().item() <= 1.0
    sub = _vp(tw, scene, 0, 2, 24, 16, 512)
    for d in range(3):
        sub.lo_pos[d], sub.vsz[d] = vp.lo_pos[d], vp.vsz[d]
    exp = oracle.voxel_fill(convert(sub, oracle.VoxelParams))
    assert beq(out[:16, :24, :].cpu().numpy(), exp) == 0
