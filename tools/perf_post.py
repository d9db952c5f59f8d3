"""GPU-box helper: time the post-chain kernels at 4K with warm clocks."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench
ctx = vq.Context(0)
peak, _ = bench.hbm_peak()
W, H = 3840, 2160
px = W * H
img = torch.from_numpy(synth.hdr_image(W, H)).cuda()
a, b, t, c = (torch.empty_like(img) for _ in range(4))
e = torch.empty((2 * H, 2 * W, 4), dtype=torch.float32, device="cuda"); r = torch.empty_like(e)
(dx, dy), sc = vq.spd_setup(W, H)
mips = [torch.empty((H >> l, W >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, sc.mips + 1)]
tm = synth.default_tonemapper()
ctx.tonemap(tm, img, t)
cas_c, easu_c, rcas_c = vq.cas_setup(0.8, W, H, W, H), vq.fsr_easu_con(W, H, W, H, 2 * W, 2 * H), vq.fsr_rcas_con(0.2)
which = sys.argv[1:] or ["spd", "blur_x", "blur_y", "tonemap", "cas", "easu", "rcas"]
P = {"spd": (lambda: ctx.spd_downsample(sc, img, mips), px * (16 + 16 / 3)), "blur_x": (lambda: ctx.gaussian_blur(img, a, False), px * 32),
     "blur_y": (lambda: ctx.gaussian_blur(img, b, True), px * 32), "tonemap": (lambda: ctx.tonemap(tm, img, t), px * 32),
     "cas": (lambda: ctx.cas(cas_c, t, c), px * 32), "easu": (lambda: ctx.fsr_easu(easu_c, t, e), px * 16 + 4 * px * 16),
     "rcas": (lambda: ctx.fsr_rcas(rcas_c, e, r), 4 * px * 32)}
for k in which:
    fn, nb = P[k]
    ms = bench.time_gpu(torch, fn, 20)
    print(f"{k:8s} {ms*1e3:8.1f} us  {nb/ms/1e6:7.0f} GB/s  {nb/ms/1e6/peak:5.3f} of measured HBM peak")
