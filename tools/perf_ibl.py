"""GPU-box helper: time the IBL integrals (K2, K3, K11) at the BASELINE sizes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench
ctx = vq.Context(0)
for hw, hh in ((2048, 1024), (4096, 2048)):
    levels = vq.mip_level_count(hw, hh)
    pyr_t = torch.zeros((vq.pyramid_texel_count(hw, hh, levels), 4), dtype=torch.float32, device="cuda")
    pyr_t[: hw * hh] = torch.from_numpy(synth.hdri(hw, hh)).cuda().reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, hw, hh, levels)
    ms = bench.time_gpu(torch, lambda: ctx.hdri_build_mips(pyr), 5, warmup=2)
    print(f"{hw}x{hh}: hdri_build_mips (+ sampling copy) {ms*1e3:8.1f} us")
    res, nm = 512, 9
    spec = torch.empty((vq.cubemap_texel_count(res, nm), 4), dtype=torch.float32, device="cuda")
    ms = bench.time_gpu(torch, lambda: ctx.specular_prefilter(pyr, vq.cubemap_of(spec, res, nm), 512), 5, warmup=2)
    print(f"{hw}x{hh}: specular prefilter 512^2 x9, 512 samples {ms:8.3f} ms  ({vq.cubemap_texel_count(res, nm) / ms / 1e3:.0f} Mtexels/s)")
    diff = torch.empty((6 * 64 * 64, 4), dtype=torch.float32, device="cuda")
    ms = bench.time_gpu(torch, lambda: ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 64, 1), n_phi=64, n_theta=16, src_mip=3), 5, warmup=2)
    print(f"{hw}x{hh}: diffuse irradiance 64^2, 64x16 grid {ms*1e3:8.1f} us")
    ms = bench.time_gpu(torch, lambda: ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 64, 1), step=0.01, src_mip=3), 2, warmup=1)
    print(f"{hw}x{hh}: diffuse irradiance 64^2, engine step 0.010 {ms:8.3f} ms")
