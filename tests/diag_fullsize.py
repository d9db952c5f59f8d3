"""GPU-box diagnostic (test infrastructure: it runs the oracle, so it lives under tests/): where do the full-size parity tests (tests/test_fullsize_gpu.py) disagree with the oracle?
Prints the worst texels/pixels of the C5 specular face and of the C3 forward frame with their inputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth
import bench
import oracle_lib as orc

ctx = vq.Context(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def spec():
    hw, hh, res, mips = 4096, 2048, 512, 9
    levels = vq.mip_level_count(hw, hh)
    pyr_t = torch.zeros((vq.pyramid_texel_count(hw, hh, levels), 4), dtype=torch.float32, device="cuda")
    pyr_t[: hw * hh] = dev(synth.hdri(hw, hh, seed=synth.SEED_BASE + 5)).reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, hw, hh, levels)
    ctx.hdri_build_mips(pyr)
    n = vq.cubemap_texel_count(res, mips)
    cube = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    ctx.specular_prefilter(pyr, vq.cubemap_of(cube, res, mips), 512)
    torch.cuda.synchronize()
    got = cube.cpu().numpy(); hp = pyr_t.cpu().numpy()
    for m, f in ((1, 3), (1, 2), (1, 0), (2, 3), (0, 3)):
        nn = res >> m
        a = vq.cubemap_offset(res, m, f)
        ids = np.arange(a, a + nn * nn, dtype=np.int64)
        ref = orc.specular_prefilter_texels(hp, hw, hh, levels, res, mips, ids)
        d = np.abs(got[ids].astype(np.float64) - ref).max(axis=1)
        bad = np.argsort(-d)[:12]
        print(f"== spec mip {m} face {f}: max {d.max():.3e}, frac>1e-4 {(d > 1e-4).mean():.5f}")
        py, px = np.divmod(np.arange(nn * nn), nn)
        r = np.hypot(px - (nn - 1) / 2, py - (nn - 1) / 2) / nn       # distance from the face centre (pole for faces 2,3)
        for b in bad:
            print(f"   texel ({px[b]:4d},{py[b]:4d}) r={r[b]:.3f} err {d[b]:.3e} ref {ref[b][:3]} got {got[ids][b][:3]}")
        for lo, hi in ((0, .02), (.02, .05), (.05, .1), (.1, .2), (.2, .8)):
            msk = (r >= lo) & (r < hi)
            if msk.any():
                print(f"   r in [{lo},{hi}): n={msk.sum()} max err {d[msk].max():.3e} frac>1e-4 {(d[msk] > 1e-4).mean():.5f}")


def forward():
    envk = bench.build_env_maps_gpu(ctx, vq, torch)
    e = {k: envk[k].cpu().numpy() for k in ("diff", "spec", "lut")}
    w, h = 1920, 1080
    planes = synth.gbuffer(w, h, seed=synth.SEED_BASE + 3)
    pf, pv = synth.scene_constants(w, h, envk["spec_mips"])
    dpl = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, envk["env"], out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ref = orc.forward_lighting(pf, pv, planes, e["diff"], envk["diff_res"], e["spec"], envk["spec_res"], envk["spec_mips"], e["lut"])
    d = np.abs(got.astype(np.float64) - ref)
    sc = (d / np.maximum(1.0, np.abs(ref))).max(axis=2)
    print(f"== forward 1080p: max scaled {sc.max():.3e}; pixels > 1e-4: {(sc > 1e-4).sum()}, > 5e-5: {(sc > 5e-5).sum()}")
    ys, xs = np.unravel_index(np.argsort(-sc, axis=None)[:20], sc.shape)
    cam = np.array([pv.CameraPosition.x, pv.CameraPosition.y, pv.CameraPosition.z])
    L = pf.Lights
    for y, x in zip(ys, xs):
        P = planes[0][y, x, :3]; N = planes[1][y, x, :3]; rough = planes[1][y, x, 3]; alb = planes[2][y, x]
        V = (cam - P) / np.linalg.norm(cam - P); Nn = N / np.linalg.norm(N)
        info = []
        for i in range(L.numPointLights):
            l = L.point_lights[i]
            Lv = np.array([l.position.x, l.position.y, l.position.z]) - P
            Wi = Lv / np.linalg.norm(Lv); H = (V + Wi) / np.linalg.norm(V + Wi)
            nh = float(np.dot(Nn, H)); a2 = float(rough) ** 4
            info.append(f"L{i}: nh={nh:.7f} t={nh * nh * (a2 - 1) + 1:.3e} nl={np.dot(Nn, Wi):.3f}")
        print(f"   ({x},{y}) scaled {sc[y, x]:.3e} ref {ref[y, x, :3]} got {got[y, x, :3]} rough {rough:.4f} a2 {float(rough) ** 4:.3e} metal {alb[3]:.2f} nv {np.dot(Nn, V):.4f} | " + " ; ".join(info))


if __name__ == "__main__":
    which = sys.argv[1:] or ["spec", "forward"]
    if "spec" in which: spec()
    if "forward" in which: forward()
