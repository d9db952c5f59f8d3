"""Generates tests/golden/golden_small.npz: small seeded inputs and the ORACLE's outputs for every pass.

The reference cannot run here (D3D12 + runtime-compiled HLSL) and ships no golden
vectors (§4), so these fixtures pin the oracle against drift and give the -m gpu tests a committed
target that does not depend on rebuilding the oracle. kat.json holds the analytic known answers of
SURVEY.md §8(c) plus the constants the reference's own A_CPU FidelityFX build produced in this container.

    python tests/golden/make_golden.py        # rewrites golden_small.npz (kat.json is hand-derived)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import ctypes as C
    import oracle_lib as orc
    from vqengine_b200 import synth
    from envmaps import small_env
    g = {}
    env = small_env()
    w, h = 24, 14
    planes = synth.gbuffer(w, h, seed=123, emissive=True)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=123, n_point=3, n_spot=2, casters=True, hdri_offset=0.3)
    g["fwd_planes"] = np.stack(planes)
    g["fwd_per_frame"] = np.frombuffer(bytes(pf), dtype=np.uint8).copy()
    g["fwd_per_view"] = np.frombuffer(bytes(pv), dtype=np.uint8).copy()
    g["env_diff"], g["env_spec"], g["env_lut"] = env["diff"], env["spec"], env["lut"]
    g["env_dims"] = np.array([env["diff_res"], env["spec_res"], env["spec_mips"]], np.int32)
    g["fwd_out"] = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"],
                                        env["spec_mips"], env["lut"])
    img = synth.hdr_image(40, 24, seed=5)
    g["post_in"] = img
    g["blur_x"] = orc.gaussian_blur(img, False)
    g["blur_xy"] = orc.gaussian_blur(g["blur_x"], True)
    tm = synth.default_tonemapper()
    g["tonemap_srgb"] = orc.tonemap(tm, img)
    from vqengine_b200 import TonemapperParams
    g["tonemap_pq"] = orc.tonemap(TonemapperParams(0, 1, 200.0, 0, 1.0), img)
    ldr = g["tonemap_srgb"]
    g["cas_08"] = orc.cas(orc.cas_setup(0.8, 40, 24, 40, 24), ldr)
    g["easu_2x_wrap"] = orc.fsr_easu(orc.fsr_easu_con(40, 24, 40, 24, 80, 48), ldr, 80, 48, 0)
    g["easu_2x_clamp"] = orc.fsr_easu(orc.fsr_easu_con(40, 24, 40, 24, 80, 48), ldr, 80, 48, 1)
    g["rcas_02"] = orc.fsr_rcas(orc.fsr_rcas_con(0.2), ldr)
    simg = synth.hdr_image(72, 40, seed=6)
    g["spd_in"] = simg
    for i, m in enumerate(orc.spd_downsample(simg, 5), start=1):
        g[f"spd_l{i}"] = m
    hd = synth.hdri(64, 32, seed=9)
    levels = 6
    pyr = orc.hdri_build_mips(hd, levels)
    g["hdri"] = hd
    g["hdri_pyr"] = pyr
    g["diffuse_8"] = orc.diffuse_irradiance(pyr, 64, 32, levels, 8, n_phi=16, n_theta=8, src_mip=1)
    g["diffuse_4_step05"] = orc.diffuse_irradiance(pyr, 64, 32, levels, 4, step=0.05, src_mip=2)
    g["spec_8x3"] = orc.specular_prefilter(pyr, 64, 32, levels, 8, 3, num_samples=64)
    g["lut_16"] = orc.brdf_integration_lut(16, 16, samples=256)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **g)
    print("wrote golden_small.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
