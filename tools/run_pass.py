"""GPU-box helper for profiling: builds the 4K inputs and runs the chosen pass a few times.
usage: python tools/run_pass.py forward|post|ibl|frame [iters]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "forward"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = vq.Context(0)
W, H = 3840, 2160
if which == "forward":
    envk = bench.build_env_maps_gpu(ctx, vq, torch)
    planes = synth.gbuffer(W, H)
    pf, pv = synth.scene_constants(W, H, envk["spec_mips"])
    dpl = [torch.from_numpy(p).cuda() for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    for _ in range(iters):
        ctx.forward_lighting(pf, pv, gb, envk["env"], out)
    torch.cuda.synchronize()
    print("nan:", int(torch.isnan(out).sum()))
elif which == "post":
    envk = None
    print(bench.extra_passes.__name__)
    img = torch.from_numpy(synth.hdr_image(W, H)).cuda()
    a, b, t, c = (torch.empty_like(img) for _ in range(4))
    e = torch.empty((2 * H, 2 * W, 4), dtype=torch.float32, device="cuda"); r = torch.empty_like(e)
    (dx, dy), sc = vq.spd_setup(W, H)
    mips = [torch.empty((H >> l, W >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, sc.mips + 1)]
    tm = synth.default_tonemapper()
    for _ in range(iters):
        ctx.spd_downsample(sc, img, mips)
        ctx.gaussian_blur(img, a, False); ctx.gaussian_blur(a, b, True)
        ctx.tonemap(tm, b, t)
        ctx.cas(vq.cas_setup(0.8, W, H, W, H), t, c)
        ctx.fsr_easu(vq.fsr_easu_con(W, H, W, H, 2 * W, 2 * H), c, e)
        ctx.fsr_rcas(vq.fsr_rcas_con(0.2), e, r)
    torch.cuda.synchronize()
elif which == "frame":
    envk = bench.build_env_maps_gpu(ctx, vq, torch)
    src = torch.from_numpy(synth.hdri(4096, 2048)).cuda()
    data = ctx.hdr_save_host(src)
    half = torch.empty((1024, 2048, 4), dtype=torch.float32, device="cuda")
    scene = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    refl = torch.rand((H, W, 4), dtype=torch.float32, device="cuda")
    _, inv = synth.sky_view_proj(0.7, 0.1, 1.0, W / H)
    for _ in range(iters):
        img, lum = ctx.hdr_decode(data)
        ctx.hdr_encode_rgbe(src)
        ctx.image_resize(src, half)
        ctx.skydome(inv.astype(np.float32).reshape(16), envk["pyr"], scene)
        ctx.apply_reflections(scene, refl)
    torch.cuda.synchronize()
else:
    for _ in range(iters):
        envk = bench.build_env_maps_gpu(ctx, vq, torch)
print("done", which)
