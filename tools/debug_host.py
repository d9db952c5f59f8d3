"""GPU-box debug helper: device path vs host-buffer path of K1 at 4K, mismatch statistics + raw kernel timing."""
import os, sys, time
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
pf, pv = synth.scene_constants(W, H, envk["spec_mips"])
dpl = [torch.from_numpy(p).cuda() for p in planes]
gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
out = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
out2 = torch.zeros_like(out)
ctx.forward_lighting(pf, pv, gb, envk["env"], out)
ctx.forward_lighting(pf, pv, gb, envk["env"], out2)
torch.cuda.synchronize()
print("device run-to-run equal:", torch.equal(out, out2), "nan:", int(torch.isnan(out).sum()), "inf:", int(torch.isinf(out).sum()),
      "max:", float(out[torch.isfinite(out)].max()))
ms = bench.time_gpu(torch, lambda: ctx.forward_lighting(pf, pv, gb, envk["env"], out), 20)
print(f"forward 4K: {ms*1e3:.1f} us  {W*H/ms/1e3:.0f} Mpx/s  {64*W*H/ms/1e6:.0f} GB/s algorithmic")
hpl = [torch.from_numpy(p).pin_memory() for p in planes]
hgb = vq.GBuffer(vq.image_of(hpl[0]), vq.image_of(hpl[1]), vq.image_of(hpl[2]), vq.null_image())
hout = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
ctx.forward_lighting_host(pf, pv, hgb, envk["env"], hout)
d = hout.cuda()
ne = (d != out) & ~(torch.isnan(d) & torch.isnan(out))
print("host vs device mismatching values:", int(ne.sum()))
if int(ne.sum()):
    rows = torch.nonzero(ne.any(dim=2).any(dim=1)).flatten()
    print("rows with mismatches:", rows[:20].tolist(), "... count", len(rows))
    idx = torch.nonzero(ne)[:5]
    for i in idx.tolist():
        print(i, float(d[tuple(i)]), float(out[tuple(i)]))
t0 = time.perf_counter()
for _ in range(5):
    ctx.forward_lighting_host(pf, pv, hgb, envk["env"], hout)
dt = (time.perf_counter() - t0) / 5
print(f"host path: {dt*1e3:.2f} ms per 4K frame -> {W*H/dt/1e6:.0f} Mpx/s ; PCIe {(4*W*H*16)/dt/1e9:.1f} GB/s both directions")
