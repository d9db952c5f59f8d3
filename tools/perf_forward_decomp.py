"""GPU-box helper: where does K1's time go? Times the 4K forward pass with parts of the work switched off through
runtime parameters (same library, same launch): no lights, diffuse-only IBL, constant roughness (coherent specular
mip / LUT row), smooth normals (coherent cube taps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench
ctx = vq.Context(0)
envk = bench.build_env_maps_gpu(ctx, vq, torch)
W, H = 3840, 2160
planes = synth.gbuffer(W, H)
out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")

def run(tag, planes_, n_point=4, directional=True, diffuse_only=False):
    pf, pv = synth.scene_constants(W, H, envk["spec_mips"], n_point=n_point, directional=directional)
    pv.EnvironmentMapDiffuseOnlyIllumination = int(diffuse_only)
    dpl = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in planes_]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    ms = bench.time_gpu(torch, lambda: ctx.forward_lighting(pf, pv, gb, envk["env"], out), 30, warmup=5)
    print(f"{tag:40s} {ms*1e3:8.1f} us", flush=True)

run("full (4 point + dir + IBL)", planes)
run("no lights (IBL only)", planes, n_point=0, directional=False)
run("1 point light + IBL", planes, n_point=1, directional=False)
run("lights + diffuse-only IBL", planes, diffuse_only=True)
run("no lights + diffuse-only IBL", planes, n_point=0, directional=False, diffuse_only=True)
p2 = [p.copy() for p in planes]; p2[1][..., 3] = 0.5
run("full, constant roughness 0.5", p2)
p3 = [p.copy() for p in planes]; p3[1][..., 0] = 0.0; p3[1][..., 1] = 1.0; p3[1][..., 2] = 0.0
run("full, constant normal (0,1,0)", p3)
p4 = [p.copy() for p in p3]; p4[1][..., 3] = 0.5
run("full, constant normal + roughness", p4)
run("no lights, constant normal + roughness", p4, n_point=0, directional=False)
