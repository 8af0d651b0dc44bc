"""Shared parameter sets for the parity tests (BASELINE.json configs at oracle-sized grids + edge cases)."""
import ctypes as C

HM_CFG = dict(sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0)   # scene_config/config.txt:76
HM_ALL = dict(plat_bot=0.1, plat_h=0.2, plat_s=5.0, plat_max=0.3, crat_h=0.5, crat_s=1.0, crack_lo=-0.2, crack_hi=-0.1, crack_d=0.5,
              sine_mag=5.0, sine_freq=0.001, sine_bias=-4.0, volcano_width=2000.0, volcano_height=3.0)


def convert(obj, cls):
    """Reinterpret one ctypes POD as the layout-identical class of another module (oracle <-> product mirrors of include/tw3d.h)."""
    assert C.sizeof(obj) == C.sizeof(cls), (C.sizeof(obj), C.sizeof(cls))
    return cls.from_buffer_copy(bytes(obj))


def height_cases():
    """(name, SceneConfig kwargs, (x0, y0, dx_mult), (nx, ny))"""
    out = []
    for mode in (0, 1, 2, 3, 4):
        for shape in (0, 1, 2):
            for tag, ff, hmap, gl, custom, ms in (("cfg", 1, HM_CFG, 1, 0.0, 1.0), ("plain", 0, {}, 0, 0.0, 1.0), ("all", 2, HM_ALL, 1, 0.0, 1.0),
                                                  ("zoom", 1, HM_ALL, 1, 0.0, 4.0)):
                kw = dict(mesh_gen_mode=mode, mesh_gen_shape=shape, mesh_freq_filter=ff, hmap=hmap, glaciate=gl, custom_glaciate_exp=custom,
                          mesh_scale=ms, mesh_seed=1, zmax_est=2.3)
                n = 70 if mode == 4 else 130
                for oi, org in enumerate(((-n / 2, -n / 2, 1.0), (1000.0, -3000.0, 1.0), (-50000.0, 70000.0, 3.0))):
                    if tag != "cfg" and oi == 1:
                        continue
                    out.append(("m%d_s%d_%s_o%d" % (mode, shape, tag, oi), kw, org, (n, n - 7)))
    return out
