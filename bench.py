#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 shading backend (contract: task brief + BASELINE.json).

A "step" is ONE forward-PBR lighting pass (K1) over a 3840x2160 synthetic G-buffer with 4 point lights,
1 directional light and IBL (BASELINE.json metric "Mpixels/s forward-PBR @4K"), inputs resident in HBM.
  value      = Mpixels/s over all ranks (weak scaling: every rank shades its own 3840x2160 row tile)
  e2e        = the same pass through the blocking host-buffer C-ABI call (pinned host G-buffer in,
               host image out; H2D/D2H inside the timed region)
  roofline   = algorithmic 64 B/pixel / kernel time, against the measured HBM copy peak
  cpu_baseline = the scalar oracle (CPU port of the HLSL) on this box's host cores, bounded row sample
  extra      = per-kernel timings for the other SURVEY.md §8 rows (post chain @4K, IBL integrals)
`--impl reference` times the reference's shader text compiled for the CPU (oracle/_ref/libhlslref.so), or the CPU oracle when that
library is absent (the reference's D3D12/HLSL path itself cannot run here).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

W4K, H4K = 3840, 2160
BYTES_PER_PX = 64          # 3 x float4 in + 1 x float4 out
METRIC = "forward_pbr_4k_mpixels_per_s"
UNIT = "Mpixels/s"
WORKLOAD = "forward-PBR 3840x2160 G-buffer (3 float4 planes), 4 point + 1 directional + IBL (64^2 diffuse, 512^2 x9 specular, 1024^2 LUT)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def build_env_maps_gpu(ctx, vq, torch, hdri_w=2048, hdri_h=1024, diff_res=64, spec_res=512, spec_mips=9, lut=1024):
    """IBL inputs of the forward pass, produced by OUR kernels from a synthetic HDRI (setup, untimed)."""
    from vqengine_b200 import synth
    levels = vq.mip_level_count(hdri_w, hdri_h)
    pyr_t = torch.zeros((vq.pyramid_texel_count(hdri_w, hdri_h, levels), 4), dtype=torch.float32, device="cuda")
    pyr_t[: hdri_w * hdri_h] = torch.from_numpy(synth.hdri(hdri_w, hdri_h)).cuda().reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, hdri_w, hdri_h, levels)
    ctx.hdri_build_mips(pyr)
    diff = torch.zeros((6 * diff_res * diff_res, 4), dtype=torch.float32, device="cuda")
    ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, diff_res, 1), n_phi=64, n_theta=16, src_mip=3)
    # per-face Gaussian blur (EnvironmentMapRendering.cpp:279-373)
    faces = diff.view(6, diff_res, diff_res, 4)
    tmp = torch.empty_like(faces[0]); blurred = torch.empty_like(faces)
    for f in range(6):
        ctx.gaussian_blur(faces[f], tmp, False)
        ctx.gaussian_blur(tmp, blurred[f], True)
    spec = torch.zeros((vq.cubemap_texel_count(spec_res, spec_mips), 4), dtype=torch.float32, device="cuda")
    ctx.specular_prefilter(pyr, vq.cubemap_of(spec, spec_res, spec_mips), 512)
    lut_t = torch.zeros((lut, lut, 2), dtype=torch.float32, device="cuda")
    ctx.brdf_integration_lut(lut_t, 2048)
    torch.cuda.synchronize()
    keep = dict(pyr_t=pyr_t, pyr=pyr, diff=blurred.reshape(-1, 4).contiguous(), spec=spec, lut=lut_t,
                diff_res=diff_res, spec_res=spec_res, spec_mips=spec_mips, levels=levels, hdri_w=hdri_w, hdri_h=hdri_h)
    keep["env"] = vq.EnvironmentMaps(vq.cubemap_of(keep["diff"], diff_res, 1), vq.cubemap_of(spec, spec_res, spec_mips),
                                     vq.image_of(lut_t, 2))
    # the RENDER_TARGET -> SHADER_RESOURCE transition after prefiltering: bordered sampling copies, built once
    ctx.environment_prepare(keep["env"])
    torch.cuda.synchronize()
    return keep


def time_gpu(torch, fn, iters, warmup=3, min_warm_ms=30.0, min_timed_ms=20.0):
    """CUDA-event timing on the current stream. Warm-up runs at least `warmup` launches AND `min_warm_ms` of GPU work (the
    SM clock needs a few ms of load to leave its idle state after host-side input generation); the timed region is at
    least `iters` launches and about `min_timed_ms` long."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); n = 0
    while True:
        fn(); n += 1
        if n >= warmup:
            e1.record(); torch.cuda.synchronize()
            if e0.elapsed_time(e1) >= min_warm_ms or n >= 2000:
                break
    per = max(e0.elapsed_time(e1) / n, 1e-3)
    iters = max(iters, min(int(min_timed_ms / per), 2000))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters   # ms


def extra_passes(ctx, vq, torch, envk, peak):
    """Per-kernel numbers for the other SURVEY.md §8 rows (single GPU; not part of `value`)."""
    from vqengine_b200 import synth
    out = {}
    w, h = W4K, H4K
    px = w * h
    img = torch.from_numpy(synth.hdr_image(w, h)).cuda()
    a, b, t, c = (torch.empty_like(img) for _ in range(4))
    e = torch.empty((2 * h, 2 * w, 4), dtype=torch.float32, device="cuda")
    r = torch.empty_like(e)
    (dx, dy), sc = vq.spd_setup(w, h)
    mips = [torch.empty((h >> l, w >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, sc.mips + 1)]
    tm = synth.default_tonemapper()
    cas_c, easu_c, rcas_c = vq.cas_setup(0.8, w, h, w, h), vq.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), vq.fsr_rcas_con(0.2)
    passes = [
        ("spd", lambda: ctx.spd_downsample(sc, img, mips), px * (16 + 16 / 3)),
        ("blur_x", lambda: ctx.gaussian_blur(img, a, False), px * 32),
        ("blur_y", lambda: ctx.gaussian_blur(a, b, True), px * 32),
        ("tonemap", lambda: ctx.tonemap(tm, b, t), px * 32),
        ("cas", lambda: ctx.cas(cas_c, t, c), px * 32),
        ("fsr_easu_2x", lambda: ctx.fsr_easu(easu_c, c, e), px * 16 + 4 * px * 16),
        ("fsr_rcas_8k", lambda: ctx.fsr_rcas(rcas_c, e, r), 4 * px * 32),
    ]
    chain_ms, chain_bytes = 0.0, 0.0
    for name, fn, nbytes in passes:
        ms = time_gpu(torch, fn, 10)
        gbs = nbytes / ms / 1e6
        out[name] = {"ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1), "hbm_frac": round(gbs / peak, 3)}
        chain_ms += ms; chain_bytes += nbytes
    out["post_chain_4k"] = {"ms": round(chain_ms, 4), "input_Mpixels_per_s": round(px / chain_ms / 1e3, 1),
                            "algorithmic_GBps": round(chain_bytes / chain_ms / 1e6, 1),
                            "hbm_frac": round(chain_bytes / chain_ms / 1e6 / peak, 3)}
    del img, a, b, t, c, e, r, mips
    # IBL integrals (bound: SFU/FP32 + L1/L2, not HBM; texels/s and samples/s are the honest figures)
    pyr = envk["pyr"]
    res, nm = envk["spec_res"], envk["spec_mips"]
    spec = torch.empty_like(envk["spec"])
    ms = time_gpu(torch, lambda: ctx.specular_prefilter(pyr, vq.cubemap_of(spec, res, nm), 512), 3, warmup=1)
    texels = vq.cubemap_texel_count(res, nm)
    out["ibl_specular_prefilter"] = {"config": f"{envk['hdri_w']}x{envk['hdri_h']} HDRI -> {res}^2 x6 x{nm} mips, 512 samples",
                                     "ms": round(ms, 3), "texels_per_s": round(texels / ms * 1e3),
                                     "samples_per_s": round(texels * 512 / ms * 1e3)}
    diff = torch.empty((6 * 64 * 64, 4), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 64, 1), n_phi=64, n_theta=16, src_mip=3), 5, warmup=1)
    out["ibl_diffuse_irradiance"] = {"config": "2048x1024 HDRI -> 64^2 x6, 64x16 = 1024 samples (BASELINE config 2)",
                                     "ms": round(ms, 4), "texels_per_s": round(6 * 64 * 64 / ms * 1e3),
                                     "samples_per_s": round(6 * 64 * 64 * 1024 / ms * 1e3)}
    ms = time_gpu(torch, lambda: ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 64, 1), step=0.01, src_mip=3), 1, warmup=1)
    out["ibl_diffuse_irradiance_reference_step"] = {"config": "step 0.010 -> 629x158 = 99382 samples/texel (engine default)",
                                                    "ms": round(ms, 3), "samples_per_s": round(6 * 64 * 64 * 99382 / ms * 1e3)}
    lut = torch.empty((1024, 1024, 2), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: ctx.brdf_integration_lut(lut, 2048), 2, warmup=1)
    out["brdf_lut"] = {"config": "1024^2, 2048 samples", "ms": round(ms, 3), "samples_per_s": round(1024 * 1024 * 2048 / ms * 1e3)}
    ms = time_gpu(torch, lambda: ctx.hdri_build_mips(pyr), 5, warmup=1)
    nb = envk["hdri_w"] * envk["hdri_h"] * (16 * 4 / 3 + 16 / 3)
    out["hdri_min_pyramid"] = {"ms": round(ms, 4), "algorithmic_GBps": round(nb / ms / 1e6, 1)}
    out.update(surface_producer_pass(ctx, vq, torch, peak))
    out.update(frame_format_passes(ctx, vq, torch, envk, peak))
    return out


def frame_format_passes(ctx, vq, torch, envk, peak):
    """.hdr decode / encode at BASELINE config 5's HDRI size, skydome + ApplyReflections at 4K"""
    import time
    from vqengine_b200 import synth
    out = {}
    hw, hh = 4096, 2048
    src = torch.from_numpy(synth.hdri(hw, hh)).cuda()
    data = ctx.hdr_save_host(src)                               # warm-up (allocations, worker threads, lazy module load)
    t0 = time.perf_counter(); data = ctx.hdr_save_host(src); t_save = time.perf_counter() - t0
    info, offs = vq.hdr_parse(data)
    n = len(data)
    dfile = torch.zeros(((n + 15) // 16 * 16 + 16,), dtype=torch.uint8, device="cuda")
    dfile[:n] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    doffs = torch.from_numpy(offs.view(np.int64)).cuda()
    img = torch.empty((hh, hw, 4), dtype=torch.float32, device="cuda")
    lum = torch.zeros((1,), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: vq._check(vq.lib.vq_hdr_decode(ctx._h, dfile.data_ptr(), n, C.byref(info), doffs.data_ptr(),
                                                                vq.image_of(img), lum.data_ptr(), vq._stream_ptr(None))), 10)
    nb = n + hw * hh * 16
    out["hdr_decode_4096x2048"] = {"config": f".hdr file {n / 1e6:.1f} MB (RLE, {n / hw / hh:.2f} B/texel) -> RGBA32F + max luminance, bit-exact vs stbi_loadf",
                                   "ms": round(ms, 4), "Mtexels_per_s": round(hw * hh / ms / 1e3, 1),
                                   "algorithmic_GBps": round(nb / ms / 1e6, 1), "hbm_frac": round(nb / ms / 1e6 / peak, 3)}
    ctx.hdr_load_host(data, img)                                # warm-up
    t0 = time.perf_counter(); ctx.hdr_load_host(data, img); t_load = time.perf_counter() - t0
    out["hdr_decode_4096x2048"]["e2e_host_file_to_device_image_ms"] = round(t_load * 1e3, 2)
    ms = time_gpu(torch, lambda: ctx.hdr_encode_rgbe(src), 10)
    nb = hw * hh * 20
    out["hdr_encode_rgbe_4096x2048"] = {"ms": round(ms, 4), "algorithmic_GBps": round(nb / ms / 1e6, 1), "hbm_frac": round(nb / ms / 1e6 / peak, 3),
                                        "e2e_device_image_to_host_file_ms": round(t_save * 1e3, 2)}
    half = torch.empty((hh // 2, hw // 2, 4), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: ctx.image_resize(src, half), 10)
    nb = hw * hh * 16 + (hw // 2) * hh * 32 + (hw // 2) * (hh // 2) * 16      # read in, write + read the intermediate, write out
    out["image_resize_4096x2048_to_2048x1024"] = {"config": "stbir_resize_float (Mitchell, separable, edge clamp), bit-exact; gather tables cached in the context, horizontal taps staged in shared memory",
                                                   "ms": round(ms, 4), "algorithmic_GBps": round(nb / ms / 1e6, 1), "hbm_frac": round(nb / ms / 1e6 / peak, 3)}
    w, h = W4K, H4K
    _, inv = synth.sky_view_proj(0.7, 0.1, 1.0, w / h)
    scene = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: ctx.skydome(inv.astype(np.float32).reshape(16), envk["pyr"], scene), 10)
    out["skydome_4k"] = {"config": "every pixel background (no mask), 2048x1024 HDRI level 0", "ms": round(ms, 4),
                         "algorithmic_GBps": round(w * h * 16 / ms / 1e6, 1), "hbm_frac": round(w * h * 16 / ms / 1e6 / peak, 3)}
    refl = torch.rand((h, w, 4), dtype=torch.float32, device="cuda")
    ms = time_gpu(torch, lambda: ctx.apply_reflections(scene, refl), 10)
    out["apply_reflections_4k"] = {"ms": round(ms, 4), "algorithmic_GBps": round(w * h * 48 / ms / 1e6, 1),
                                   "hbm_frac": round(w * h * 48 / ms / 1e6 / peak, 3)}
    return out


def surface_scene_gpu(ctx, vq, torch, w, h, n_materials=4, tex_res=1024):
    """SURVEY 8(f).1 workload: n materials x up to 6 RGBA8 maps of tex_res^2 (mip chains built on the GPU by
    vq_texture_build_mips) + the three interpolant planes and the SSAO plane of a w x h frame"""
    from vqengine_b200 import synth
    mats, texs = synth.materials(n_materials, tex_res, uniform=True)
    keep, mts, tex_bytes = [], [], 0
    for t in texs:
        mt = vq.MaterialTextures()
        for slot, lvl0 in t.items():
            if lvl0 is None:
                continue
            th, tw = lvl0.shape[:2]
            levels = vq.mip_level_count(tw, th)
            buf = torch.zeros(vq.pyramid_texel_count(tw, th, levels) * 4, dtype=torch.uint8, device="cuda")
            buf[: tw * th * 4] = torch.from_numpy(lvl0.reshape(-1)).cuda()
            desc = vq.texture_of(buf, tw, th, levels)
            ctx.texture_build_mips(desc)
            setattr(mt, slot, desc)
            keep.append(buf); tex_bytes += buf.numel()
        mts.append(mt)
    table = ctx.material_table(mats, mts)
    planes = [torch.from_numpy(p).cuda() for p in synth.surface_inputs(w, h, n_materials)]
    si = vq.SurfaceInputs(vq.image_of(planes[0]), vq.image_of(planes[1]), vq.image_of(planes[2]), vq.image_of(planes[3], 1))
    return {"table": table, "inputs": si, "keep": keep + planes, "texture_bytes": tex_bytes, "n_materials": n_materials,
            "tex_res": tex_res}


def surface_producer_pass(ctx, vq, torch, peak):
    w, h = W4K, H4K
    sc = surface_scene_gpu(ctx, vq, torch, w, h)
    g = [torch.empty((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
    gb = vq.GBuffer(*(vq.image_of(t) for t in g))
    ms = time_gpu(torch, lambda: ctx.gbuffer_from_materials(sc["inputs"], sc["table"], 0.3, gb), 10)
    nbytes = w * h * (48 + 4 + 64)          # 3 float4 interpolant planes + SSAO in, 4 float4 G-buffer planes out
    out = {"surface_producer_4k": {
        "config": f"{sc['n_materials']} materials (separate maps / ORM / constants / tiled non-pow2), {sc['tex_res']}^2 RGBA8 "
                  f"maps = {sc['texture_bytes'] / 1e6:.1f} MB (L2-resident side data), SSAO + emissive planes",
        "ms": round(ms, 4), "Mpixels_per_s": round(w * h / ms / 1e3, 1), "algorithmic_bytes_per_px": 116,
        "algorithmic_GBps": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / peak, 3)}}
    tw = 4096
    levels = vq.mip_level_count(tw, tw)
    buf = torch.randint(0, 256, (vq.pyramid_texel_count(tw, tw, levels) * 4,), dtype=torch.uint8, device="cuda")
    desc = vq.texture_of(buf, tw, tw, levels)
    ms = time_gpu(torch, lambda: ctx.texture_build_mips(desc), 10)
    nb = tw * tw * 4 * (4 / 3 + 1 / 3)      # every level read once (except the last), every level but 0 written once
    out["texture_box_mips_4096"] = {"ms": round(ms, 4), "algorithmic_GBps": round(nb / ms / 1e6, 1),
                                    "hbm_frac": round(nb / ms / 1e6 / peak, 3), "launches": (levels - 1 + 5) // 6}
    sc["table"].close()
    return out


def ibl_specular_strong_scaling(ctx, vq, torch, dist, rank, world, hdri_w=4096, hdri_h=2048, res=512, mips=9, iters=3):
    """BASELINE config 5 (IBL half): 4096x2048 HDRI -> 512^2 x6 x9-mip specular prefilter, STRONG scaling: the flattened
    (mip, face, row) space is cut into `world` cost-balanced contiguous ranges, every rank prefilters its range from a
    replicated HDRI pyramid, then ONE all-gather assembles the packed cubemap on every rank (inside the timed region)."""
    from vqengine_b200 import synth, distributed as vd
    levels = vq.mip_level_count(hdri_w, hdri_h)
    pyr_t = torch.zeros((vq.pyramid_texel_count(hdri_w, hdri_h, levels), 4), dtype=torch.float32, device="cuda")
    pyr_t[: hdri_w * hdri_h] = torch.from_numpy(synth.hdri(hdri_w, hdri_h)).cuda().reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, hdri_w, hdri_h, levels)
    ctx.hdri_build_mips(pyr)
    cube_t = torch.zeros((vq.cubemap_texel_count(res, mips), 4), dtype=torch.float32, device="cuda")
    cube = vq.cubemap_of(cube_t, res, mips)
    plan = vd.InterleavedSpecularPlan(res, mips, world)
    my_rows = plan.row_ranges(rank)

    def compute():
        for rb, re in my_rows:
            ctx.specular_prefilter(pyr, cube, 512, rb, re)

    def step():
        compute()
        if dist is not None and world > 1:
            plan.gather(cube_t, rank)

    step(); torch.cuda.synchronize()
    if dist: dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    for _ in range(iters):
        compute()
    e2.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters, e1.elapsed_time(e2) / iters], dtype=torch.float64, device="cuda")
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_compute = float(t[0]), float(t[1])
    n_tex = vq.cubemap_texel_count(res, mips)
    # ---- fused compute + gather: the prefilter kernels store every texel straight into all ranks' cubemaps over NVLink
    #      (peer pointers from torch symmetric memory); one device-side barrier ends the step ----
    fused = None
    if dist is not None and world > 1:
        try:
            import torch.distributed._symmetric_memory as symm
            sym_t = symm.empty((n_tex, 4), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            hdl = symm.rendezvous(sym_t, dist.group.WORLD)
            peers = [hdl.get_buffer(r, (n_tex, 4), torch.float32) for r in range(world)]
            order = [rank] + [r for r in range(world) if r != rank]
            cubes = [vq.cubemap_of(peers[r], res, mips) for r in order]
            split_rows = [(r0 + rank * k, r0 + (rank + 1) * k) for (_, r0, k, _) in plan.split]
            repl_rows = [(r0, r0 + rows) for (_, r0, rows, _) in plan.replicated]

            def fused_step():
                for rb, re in split_rows:
                    ctx.specular_prefilter_multi(pyr, cubes, 512, rb, re)       # my block of every split mip -> all ranks
                for rb, re in repl_rows:
                    ctx.specular_prefilter(pyr, cubes[0], 512, rb, re)          # tiny mips: every rank computes its own copy
                hdl.barrier()

            fused_step(); torch.cuda.synchronize(); dist.barrier()
            same = bool(torch.equal(sym_t, cube_t))
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(iters):
                fused_step()
            f1.record(); torch.cuda.synchronize()
            tf = torch.tensor([f0.elapsed_time(f1) / iters], dtype=torch.float64, device="cuda")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            fused = {"ms": round(float(tf[0]), 3), "texels_per_s": round(n_tex / float(tf[0]) * 1e3), "equals_nccl_allgather": same,
                     "how": "vq_specular_prefilter_multi: every prefiltered texel is stored into all ranks' cubemaps (symmetric-memory peer pointers, NVLink P2P) while the SMs integrate; one device-side barrier"}
            del sym_t, peers
        except Exception as ex:   # symmetric memory unavailable on this box: keep the NCCL numbers
            fused = {"error": repr(ex)[:300]}
    del pyr_t, cube_t
    return {"config": f"{hdri_w}x{hdri_h} HDRI -> {res}^2 x6 x{mips} mips, 512 samples; strong scaling over {world} GPU(s): every mip split into {world} equal row blocks (tiny mips replicated) + ONE all-gather of the 33.5 MB cubemap",
            "ms": round(ms, 3), "texels_per_s": round(n_tex / ms * 1e3), "ms_compute_only": round(ms_compute, 3),
            "rows_per_rank": sum(b - a for a, b in my_rows), "replicated_mips": [m for (m, _, _, _) in plan.replicated],
            "fused_p2p": fused}


# ------------------------------------------------------------------------------------------------
def cpu_reference_forward(planes, rows_target_s=12.0, env=None, threads=None):
    """cpu_baseline of the GPU arm: the scalar oracle over a bounded row sample of the 4K workload. Preferred form: one process per
    host core, run in a FRESH interpreter (`bench.py --impl cpu-port`: no fork of this process, which holds a CUDA context) —
    on the pool's boxes processes scale where the threads of one process do not. Fallback: std::thread row split in this process."""
    try:
        env_ = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-port", "--steps", "4", "--warmup", "1"],
                           capture_output=True, text=True, timeout=420, env=env_)
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        px = j["rows"] * j["reps"] * W4K
        return {"value": round(j["value"], 4), "unit": UNIT, "cores": j["procs"], "kind": "port",
                "sample": f"4 x {j['reps']} x ({j['rows']} of {H4K} rows x {W4K} px of the same 4K G-buffer) = {4 * px} px at {j['ms']:.1f} ms per pass, scalar C++ "
                          f"oracle, one forked process per host core ({j['one_process_mpx_s']} Mpixels/s per process)"}
    except Exception as ex:
        print(f"# multi-process cpu_baseline unavailable ({ex!r}); using threads", file=sys.stderr)
    return _cpu_reference_forward_threads(planes, rows_target_s, env, threads)


def _cpu_reference_forward_threads(planes, rows_target_s=12.0, env=None, threads=None):
    """The scalar oracle on `threads` host threads over a bounded row sample of the 4K workload."""
    import oracle_lib as orc
    from vqengine_b200 import synth
    threads = threads or orc.cpu_threads()
    probe_rows = 16
    pf, pv = synth.scene_constants(W4K, H4K, env["spec_mips"])
    args = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    t0 = time.perf_counter()
    orc.forward_lighting(pf, pv, planes, *args, 0, probe_rows, threads)
    rate = probe_rows * W4K / (time.perf_counter() - t0)
    rows = int(min(planes[0].shape[0], max(probe_rows, rows_target_s * rate / W4K)))
    # repeat the (bounded) sample until about rows_target_s of CPU work has been timed
    t0 = time.perf_counter(); reps = 0
    while True:
        orc.forward_lighting(pf, pv, planes, *args, 0, rows, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= rows_target_s or reps >= 200:
            break
    return {"value": round(reps * rows * W4K / dt / 1e6, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x ({rows} of {H4K} rows x {W4K} px of the same 4K G-buffer) = {reps * rows * W4K} px in {dt:.1f} s, scalar C++ oracle, std::thread row split"}


def cpu_reference_image_class(hw=4096, hh=2048):
    """cpu_baseline for the SURVEY 8(f).2 rows, kind "reference": the engine's OWN Image class (Libs/VQUtils/Source/Image.cpp
    compiled unmodified into oracle/_ref/libvqimageref.so) timed on this box's host cores, single-threaded as in the engine —
    Image::LoadFromFile (stbi_loadf + CalculateMaxLuminance), Image::CreateResizedImage (stbir_resize_float), Image::SaveToDisk
    (stbi_write_hdr) on the same 4096x2048 HDRI the GPU rows are measured on. Returns None when the library is not present."""
    import tempfile
    import oracle_lib as orc
    from vqengine_b200 import synth
    if orc.image_ref() is None:
        return None
    src = synth.hdri(hw, hh)
    with tempfile.TemporaryDirectory() as td:
        path_in, path_out = os.path.join(td, "in.hdr"), os.path.join(td, "out.hdr")
        open(path_in, "wb").write(orc.hdr_encode(src))
        t0 = time.perf_counter(); texels, lum = orc.ref_image_load(path_in); t_load = time.perf_counter() - t0
        if texels is None:
            return None
        t0 = time.perf_counter(); half = orc.ref_image_resize(texels, hw // 2, hh // 2); t_resize = time.perf_counter() - t0
        t0 = time.perf_counter(); ok = orc.ref_image_save(path_out, texels); t_save = time.perf_counter() - t0
        if not ok:
            return None
    return {"kind": "reference", "cores": 1, "unit": "ms",
            "sample": f"one {hw}x{hh} equirect HDRI through the reference's Image class (oracle/_ref/libvqimageref.so), file on tmpfs/disk cache",
            "image_load_from_file_ms": round(t_load * 1e3, 1), "image_create_resized_half_ms": round(t_resize * 1e3, 1),
            "image_save_to_disk_ms": round(t_save * 1e3, 1)}


def cpu_env():
    """IBL maps for the CPU arm, built by the ORACLE at reduced sizes (the full-size maps would cost the scalar CPU
    code minutes; map size does not change the per-pixel work of the forward pass)."""
    from envmaps import small_env
    return small_env(hdri_w=256, hdri_h=128, diff_res=16, spec_res=64, spec_mips=6, lut=64, seed=77)


_REF_JOB = {}


def _cpu_rows_worker(job):
    """forked worker: rows [rb,re) of the band, `reps` times over, through the reference's shader text compiled as C++ ('text') or the
    scalar port ('port')"""
    import oracle_lib as orc
    j = _REF_JOB
    rb, re, reps = job
    acc = 0.0
    for _ in range(reps):
        if j["kind"] == "text":
            orc.hlsl_forward_gbuffer(j["pf"], j["pv"], j["planes"], *j["env"], row_begin=rb, row_end=re, out=j["out"])
            acc += float(j["out"][rb:re, :, :3].sum())
        else:
            acc += float(orc.forward_lighting(j["pf"], j["pv"], j["planes"], *j["env"], rb, re, 1, out=j["out"])[rb:re, :, :3].sum())
    return acc


def _cpu_arm(kind, steps, warmup, planes, pf, pv, a, budget_s=60.0):
    """`kind` over a bounded band of the 4K G-buffer, one FORKED PROCESS per host core (the compiled shader's cbuffers are process
    globals; and processes were measured to scale where threads of one process did not). A step = the whole band `reps` times,
    every process shading its own rows, `reps` sized from a probe so that a step lasts about budget_s / (steps + warmup) seconds
    (long enough for the task hand-off not to matter). Returns dict(value Mpx/s, ms, rows, reps, procs, one_process_mpx_s)."""
    import multiprocessing as mp
    import numpy as np
    import oracle_lib as orc
    procs = orc.cpu_threads()
    band = planes[0].shape[0]
    budget_s = float(os.environ.get("VQ_CPU_ARM_BUDGET_S", budget_s))     # tests shorten it
    _REF_JOB.update(kind=kind, pf=pf, pv=pv, planes=planes, env=a, out=np.zeros((band, W4K, 4), np.float32))
    t0 = time.perf_counter()
    _cpu_rows_worker((0, 4, 1))
    one = 4 * W4K / (time.perf_counter() - t0)                           # px/s of one process
    n = min(procs, band)
    cuts = [(band * i // n, band * (i + 1) // n) for i in range(n)]
    with mp.get_context("fork").Pool(procs) as pool:
        def step(reps):
            return sum(pool.map(_cpu_rows_worker, [(rb, re, reps) for rb, re in cuts], chunksize=1))
        step(1)                                                           # first touch of every process's pages
        t0 = time.perf_counter(); step(1); dt = time.perf_counter() - t0
        rate = band * W4K / dt                                            # measured parallel rate (a box may grant fewer CPUs than it lists)
        reps = int(max(1, round(budget_s / (steps + warmup) * rate / (band * W4K))))
        for _ in range(warmup):
            step(reps)
        t0 = time.perf_counter()
        work = 0.0
        for _ in range(steps):
            work += step(reps)
        dt = (time.perf_counter() - t0) / steps
    assert work != 0.0
    return {"value": band * reps * W4K / dt / 1e6, "ms": dt * 1e3, "rows": band, "reps": reps, "procs": procs,
            "one_process_mpx_s": round(one / 1e6, 3)}


def _cpu_workload():
    from vqengine_b200 import synth
    env = cpu_env()
    planes = synth.gbuffer(W4K, 512, seed=synth.SEED_BASE + 3)   # a 512-row band of the 4K G-buffer
    pf, pv = synth.scene_constants(W4K, H4K, env["spec_mips"])
    a = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    return planes, pf, pv, a


def run_cpu_port(args):
    """hidden helper (`--impl cpu-port`): the scalar port, one process per core, as one JSON line. The GPU arm runs this in a fresh
    interpreter for its cpu_baseline leg (no fork of a process that holds a CUDA context)."""
    planes, pf, pv, a = _cpu_workload()
    r = _cpu_arm("port", args.steps, args.warmup, planes, pf, pv, a, budget_s=15.0)
    print(json.dumps({"impl": "cpu-port", **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import oracle_lib as orc
    planes, pf, pv, a = _cpu_workload()
    port = text = None
    try:
        port = _cpu_arm("port", max(2, args.steps // 4), 1, planes, pf, pv, a, budget_s=10.0)
    except Exception as ex:
        print(f"# multi-process port unavailable ({ex!r})", file=sys.stderr)
    try:
        if orc.hlsl_ref() is not None:
            # agreement with the port on a few rows (PSMain renormalises the interpolated normal; the G-buffer pass takes it as is)
            chk = orc.hlsl_forward_gbuffer(pf, pv, planes, *a, row_begin=0, row_end=2)[:2]
            want = orc.forward_lighting(pf, pv, planes, *a, 0, 2, 1)[:2]
            rel = float(np.max(np.abs(chk - want) / np.maximum(1.0, np.abs(want))))
            text = _cpu_arm("text", args.steps, args.warmup, planes, pf, pv, a)
            text["max_scaled_delta_vs_port"] = rel
    except Exception as ex:                                        # library missing / fork unavailable: time the port instead
        print(f"# reference shader text arm unavailable ({ex!r}); timing the CPU port", file=sys.stderr)
    port_note = (f"the scalar C++ port of the same math on the same box: {port['value']:.2f} Mpixels/s ({port['procs']} processes, "
                 f"{port['one_process_mpx_s']} per process)") if port else "scalar port not timed"
    if text is not None:
        v, ms, kind, cores = text["value"], text["ms"], "reference", text["procs"]
        sample = (f"each step = {text['reps']} x ({text['rows']} rows x {W4K} px of the 4K G-buffer) through the reference's own ForwardLighting.hlsl PSMain "
                  f"(+ Lighting/BRDF/ShadingMath.hlsl) compiled as C++ (oracle/_ref/libhlslref.so), {cores} forked processes "
                  f"({text['one_process_mpx_s']} Mpixels/s per process); texture fetches served by the oracle's samplers; max scaled "
                  f"|delta| vs the port {text['max_scaled_delta_vs_port']:.1e}; {port_note}")
        note = "the reference's D3D12/HLSL path needs Windows; this arm runs its shader text compiled for the CPU"
    else:
        threads = orc.cpu_threads()
        if port is None:                                           # last resort: threads of this process
            t0 = time.perf_counter(); orc.forward_lighting(pf, pv, planes, *a, 0, 16, threads)
            rate = 16 * W4K / (time.perf_counter() - t0)
            rows = int(min(512, max(16, 60.0 * rate / W4K / (args.steps + args.warmup))))
            for _ in range(args.warmup):
                orc.forward_lighting(pf, pv, planes, *a, 0, rows, threads)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                orc.forward_lighting(pf, pv, planes, *a, 0, rows, threads)
            dt = (time.perf_counter() - t0) / args.steps
            port = {"value": rows * W4K / dt / 1e6, "ms": dt * 1e3, "rows": rows, "procs": threads}
        v, ms, kind, cores = port["value"], port["ms"], "port", port["procs"]
        sample = f"each step = {port.get('reps', 1)} x ({port['rows']} rows x {W4K} px of the 4K G-buffer) through the scalar C++ oracle on {cores} host cores"
        note = "the reference's D3D12/HLSL path needs Windows; this arm is the CPU port (oracle) of the identical math"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "note": note},
                      "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                      "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-port"])
    ap.add_argument("--no-extra", action="store_true", help="skip the per-kernel extra section")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "cpu-port":
        return run_cpu_port(args)

    import numpy as np
    import torch
    import vqengine_b200 as vq
    from vqengine_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak, peak_src = hbm_peak()
    ctx = vq.Context(local)
    envk = build_env_maps_gpu(ctx, vq, torch)

    # this rank's row tile of the 3840 x (2160*world) frame: weak scaling, fixed work per GPU
    planes = synth.gbuffer(W4K, H4K, seed=synth.SEED_BASE + 3 + rank)
    pf, pv = synth.scene_constants(W4K, H4K, envk["spec_mips"])
    dpl = [torch.from_numpy(p).cuda() for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    out = torch.zeros((H4K, W4K, 4), dtype=torch.float32, device="cuda")
    step = lambda: ctx.forward_lighting(pf, pv, gb, envk["env"], out)

    launches0 = vq.launch_count()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lt0 = vq.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    timed_launches = vq.launch_count() - lt0
    ms_total = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so that the clock sampler sees the load even for short runs
    t_end = time.time() + 0.4
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    px_all = W4K * H4K * world
    value = px_all / ms_step / 1e3   # Mpixels/s

    # ---- e2e: the blocking host-buffer call, pinned host memory, H2D + D2H inside the timed region ----
    hpl = [torch.from_numpy(p).pin_memory() for p in planes]
    hgb = vq.GBuffer(vq.image_of(hpl[0]), vq.image_of(hpl[1]), vq.image_of(hpl[2]), vq.null_image())
    hout = torch.zeros((H4K, W4K, 4), dtype=torch.float32).pin_memory()
    ctx.resize(W4K, H4K)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.forward_lighting_host(pf, pv, hgb, envk["env"], hout)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.forward_lighting_host(pf, pv, hgb, envk["env"], hout)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = px_all / float(t.item()) / 1e6
    assert torch.equal(hout.cuda(), out), "host-buffer path and device path disagree"
    h2d, d2h = 3 * W4K * H4K * 16, W4K * H4K * 16

    # ---- multi-GPU: one all-gather of the shaded tiles over NVLink (reported separately) ----
    gather = None
    if dist:
        full = torch.empty((world * H4K, W4K, 4), dtype=torch.float32, device="cuda")
        for _ in range(2):
            dist.all_gather_into_tensor(full, out)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            step()
            dist.all_gather_into_tensor(full, out)
        g1.record(); torch.cuda.synchronize()
        tg = torch.tensor([g0.elapsed_time(g1) / 5], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"ms_per_step_with_allgather": round(float(tg.item()), 4),
                  "value_with_allgather": round(px_all / float(tg.item()) / 1e3, 1),
                  "allgather_bytes_per_rank_in": (world - 1) * H4K * W4K * 16}
        # reference result of the NCCL path for the fused check below
        step(); dist.all_gather_into_tensor(full, out); torch.cuda.synchronize()
        # ---- fused compute + gather: the forward kernel stores every pixel straight into all ranks' frames over NVLink
        #      (peer pointers from torch symmetric memory), so the transfer overlaps the shading ----
        try:
            import torch.distributed._symmetric_memory as symm
            frame = symm.empty((world * H4K, W4K, 4), dtype=torch.float32, device=torch.device("cuda", local))
            hdl = symm.rendezvous(frame, dist.group.WORLD)
            frame.zero_(); torch.cuda.synchronize(); dist.barrier()
            # local frame first, then peers in a rotated order so that the ranks do not all write to the same peer at once
            imgs = [vq.Image(int(hdl.buffer_ptrs[(rank + k) % world]), W4K, world * H4K, W4K * 16) for k in range(world)]
            fused = lambda: ctx.forward_lighting_multi(pf, pv, gb, envk["env"], imgs, rank * H4K)
            for _ in range(2):
                fused(); hdl.barrier()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(5):
                fused(); hdl.barrier()
            f1.record(); torch.cuda.synchronize()
            tf = torch.tensor([f0.elapsed_time(f1) / 5], dtype=torch.float64, device="cuda")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            same = torch.tensor([1.0 if torch.equal(frame, full) else 0.0], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            gather["fused_p2p"] = {"ms_per_step": round(float(tf.item()), 4), "value": round(px_all / float(tf.item()) / 1e3, 1),
                                   "equals_nccl_allgather": bool(same.item() == 1.0),
                                   "how": "vq_forward_lighting_multi: one kernel shades the local tile and stores it into every rank's frame (symmetric-memory peer pointers, NVLink P2P), then one device-side barrier"}
            del frame
        except Exception as ex:   # symmetric memory unavailable on this box: keep the NCCL numbers
            gather["fused_p2p"] = {"error": repr(ex)[:300]}
        del full

    line = None
    if rank == 0:
        achieved = BYTES_PER_PX * W4K * H4K / (ms_step * 1e-3) / 1e9   # per-GPU GB/s of the dominant (only) kernel
        traffic = None
        tp = os.path.join(ROOT, "profiles", "forward_traffic.json")
        if os.path.exists(tp):
            try: traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception: pass
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "tile_per_gpu": f"{W4K}x{H4K}", "frame": f"{W4K}x{H4K * world}",
                           "parallelism": f"row-tiles x{world}", "l2_policy": "inputs (398 MB) + output (133 MB) per step exceed the 126 MB L2; no flush needed"},
                "roofline": {"bound": "hbm", "kernel": "forward_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": BYTES_PER_PX * W4K * H4K},
                "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "vq_forward_lighting_host (pinned host buffers, 16 row chunks pipelined over 3 streams)"},
                "gpu_launches": int(timed_launches), "clocks": clocks}
        if gather: line["allgather"] = gather
    ibl_strong = None
    if not args.no_extra:
        del hpl, hout
        hpl = hout = None
        ibl_strong = ibl_specular_strong_scaling(ctx, vq, torch, dist, rank, world)
    if rank == 0 and ibl_strong is not None:
        line["ibl_specular_prefilter_strong"] = ibl_strong
    if world == 1 and rank == 0:
        if not args.no_extra:
            line["extra"] = extra_passes(ctx, vq, torch, envk, peak)
            line["extra"]["note"] = "per-kernel figures for the other SURVEY.md section-8 rows; not part of `value`"
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_reference_forward(planes, env=cpu_env())
            try:                     # the (f).2 rows have a compilable reference: time it beside the kernels
                ref_img = cpu_reference_image_class()
                if ref_img is not None:
                    line["cpu_baseline_image_class"] = ref_img
            except Exception as ex:  # never let the side measurement cost the bench line
                line["cpu_baseline_image_class"] = {"error": repr(ex)[:200]}
    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
