"""Small IBL map sets for the forward-pass parity tests, built with the ORACLE on the CPU (test input
generation; the maps are inputs to both the CUDA kernel and the oracle)."""
import functools

import numpy as np


@functools.lru_cache(maxsize=4)
def small_env(hdri_w=128, hdri_h=64, diff_res=16, spec_res=32, spec_mips=5, lut=32, seed=77):
    import oracle_lib as orc
    from vqengine_b200 import synth
    levels = orc.lib().orc_mip_level_count(hdri_w, hdri_h)
    pyr = orc.hdri_build_mips(synth.hdri(hdri_w, hdri_h, seed=seed), levels)
    diff = orc.diffuse_irradiance(pyr, hdri_w, hdri_h, levels, diff_res, n_phi=16, n_theta=8, src_mip=1)
    # blurred per face like EnvironmentMapRendering.cpp:279-373
    faces = diff.reshape(6, diff_res, diff_res, 4)
    diff_b = np.stack([orc.gaussian_blur(orc.gaussian_blur(f, False), True) for f in faces]).reshape(-1, 4)
    spec = orc.specular_prefilter(pyr, hdri_w, hdri_h, levels, spec_res, spec_mips, num_samples=64)
    lut_img = orc.brdf_integration_lut(lut, lut, samples=128)
    return dict(pyr=pyr, levels=levels, hdri_w=hdri_w, hdri_h=hdri_h, diff=np.ascontiguousarray(diff_b), diff_res=diff_res,
                spec=spec, spec_res=spec_res, spec_mips=spec_mips, lut=lut_img)
