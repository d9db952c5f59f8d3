"""GPU-box helper for round 2: timings of the SURVEY 8(f).4 kernels (vq_shadow.cu), which have not run on a GPU yet.
4K forward pass with 2 point casters + 1 spot caster + a shadowing directional light against the unshadowed K1, and the MIN
depth pyramid of a 4K depth buffer. Prints one JSON object; run tests first: `python -m pytest tests -m gpu_next`."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench

W, H = bench.W4K, bench.H4K
ctx = vq.Context(0)
envk = bench.build_env_maps_gpu(ctx, vq, torch)
peak, _ = bench.hbm_peak()
planes = synth.gbuffer(W, H, seed=synth.SEED_BASE + 3)
pf, pv = synth.scene_constants(W, H, envk["spec_mips"], n_point=2, n_spot=1, casters=True)
L = pf.Lights
m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0
for sc in range(L.numSpotCasters):
    for k in range(16): L.shadowViews[sc].m[k] = float(m[k])
for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
L.directional.shadowing = 1
res_pt, res_2d = 1024, 2048
pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = float(res_2d)
pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = float(res_2d)
dpl = [torch.from_numpy(p).cuda() for p in planes[:3]]
gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(5)
cubes = torch.rand((max(L.numPointCasters, 1), 6, res_pt, res_pt), device="cuda", generator=g) * 1.2
spots = torch.rand((max(L.numSpotCasters, 1), res_2d, res_2d), device="cuda", generator=g) * 0.4 + 0.3
dmap = torch.rand((res_2d, res_2d), device="cuda", generator=g) * 0.4 + 0.3
r = {}
ms0 = bench.time_gpu(torch, lambda: ctx.forward_lighting(pf, pv, gb, envk["env"], out), 20)
ms1 = bench.time_gpu(torch, lambda: ctx.forward_lighting_shadowed(pf, pv, gb, envk["env"], out, cubes, spots, dmap), 20)
nb = W * H * 64
r["forward_4k_casters_unshadowed"] = {"ms": round(ms0, 4), "note": "K1 alone, casters lit with factor 1"}
r["forward_4k_casters_shadowed"] = {"ms": round(ms1, 4), "casters": f"{L.numPointCasters} point (20-tap cube PCF, {res_pt}^2 faces) + "
                                    f"{L.numSpotCasters} spot + directional (5x5 PCF, {res_2d}^2)",
                                    "algorithmic_GBps": round((nb + W * H * 80) / ms1 / 1e6, 1), "shadow_pass_ms": round(ms1 - ms0, 4)}
depth = torch.rand((H, W), device="cuda", generator=g)
n = vq.depth_pyramid_level_count(W, H)
levels = torch.empty((vq.depth_pyramid_texel_count(W, H, n),), dtype=torch.float32, device="cuda")
ms = bench.time_gpu(torch, lambda: ctx.depth_min_pyramid(depth, levels), 20)
nb = W * H * 4 * 2 + int(W * H * 4 * (1 / 3 + 2 / 3))          # copy (read + write) + every level written once, padded level read once
r["depth_min_pyramid_4k"] = {"ms": round(ms, 4), "levels": n, "algorithmic_GBps": round(nb / ms / 1e6, 1), "hbm_frac": round(nb / ms / 1e6 / peak, 3)}
print(json.dumps(r, indent=1))
