"""GPU-box helper for compute-sanitizer (memcheck / racecheck / initcheck): one small launch of every kernel.
   compute-sanitizer --tool memcheck  python tools/sanitize_small.py
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
from vqengine_b200 import synth

ctx = vq.Context(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
w, h = 100, 70
img = dev(synth.hdr_image(w, h)); a = torch.empty_like(img); b = torch.empty_like(img)
ctx.gaussian_blur(img, a, False); ctx.gaussian_blur(a, b, True)
ctx.tonemap(synth.default_tonemapper(), b, a)
ctx.cas(vq.cas_setup(0.8, w, h, w, h), a, b)
e = torch.empty((2 * h, 2 * w, 4), device="cuda"); r = torch.empty_like(e)
ctx.fsr_easu(vq.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), b, e)
ctx.fsr_easu(vq.fsr_easu_con(w, h, w, h, 80, 50), b, torch.empty((50, 80, 4), device="cuda"))
ctx.fsr_rcas(vq.fsr_rcas_con(0.2), e, r)
(dx, dy), c = vq.spd_setup(200, 140); c.mips = 7
ctx.spd_downsample(c, dev(synth.hdr_image(200, 140)), [torch.empty((140 >> l, 200 >> l, 4), device="cuda") for l in range(1, 8)])
hw, hh = 128, 64
levels = vq.mip_level_count(hw, hh)
pyr_t = torch.zeros((vq.pyramid_texel_count(hw, hh, levels), 4), device="cuda"); pyr_t[: hw * hh] = dev(synth.hdri(hw, hh)).reshape(-1, 4)
pyr = vq.pyramid_of(pyr_t, hw, hh, levels); ctx.hdri_build_mips(pyr)
diff = torch.zeros((6 * 8 * 8, 4), device="cuda"); ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 8, 1), n_phi=16, n_theta=8, src_mip=1)
ctx.diffuse_irradiance(pyr, vq.cubemap_of(diff, 8, 1), step=0.05, src_mip=1)
spec = torch.zeros((vq.cubemap_texel_count(16, 4), 4), device="cuda"); ctx.specular_prefilter(pyr, vq.cubemap_of(spec, 16, 4), 64)
lut = torch.zeros((16, 16, 2), device="cuda"); ctx.brdf_integration_lut(lut, 64)
planes = [dev(p) for p in synth.gbuffer(64, 40, emissive=True)]
pf, pv = synth.scene_constants(64, 40, 4, n_spot=2, casters=True)
gb = vq.GBuffer(*[vq.image_of(p) for p in planes])
em = vq.EnvironmentMaps(vq.cubemap_of(diff, 8, 1), vq.cubemap_of(spec, 16, 4), vq.image_of(lut, 2))
out = torch.zeros((40, 64, 4), device="cuda")
ctx.forward_lighting(pf, pv, gb, em, out)
ctx.environment_prepare(em); ctx.forward_lighting(pf, pv, gb, em, out)
# wider frame: several column tiles + ragged last tile + more rows than the persistent grid has row groups
planes2 = [dev(p) for p in synth.gbuffer(300, 9)]
pf2, pv2 = synth.scene_constants(300, 9, 4)
out2 = torch.zeros((9, 300, 4), device="cuda")
ctx.forward_lighting(pf2, pv2, vq.GBuffer(vq.image_of(planes2[0]), vq.image_of(planes2[1]), vq.image_of(planes2[2]), vq.null_image()), em, out2)
ctx.forward_lighting_multi(pf, pv, gb, em, [vq.image_of(out), vq.image_of(torch.zeros_like(out))], 0)
ctx.specular_prefilter_multi(pyr, [vq.cubemap_of(spec, 16, 4), vq.cubemap_of(torch.zeros_like(spec), 16, 4)], 64)
# round 2: the persistent range-list launch (two destinations, cuts through mip 0 and the middle of faces) and the in-kernel
# rendezvous of both fused kernels with the flag arrays of a 2-"rank" world living on this one GPU (rank 0's peer already at epoch 1)
ctx.specular_prefilter_ranges(pyr, [vq.cubemap_of(spec, 16, 4), vq.cubemap_of(torch.zeros_like(spec), 16, 4)], [(0, 5), (5, 37), (40, 150)], 64)
flags = torch.zeros((2, 2), dtype=torch.int32, device="cuda"); flags[0, 1] = 1          # [rank][word]: word 1 of rank 0's array set by "rank 1"
sig = vq.PeerSignal(); sig.flags[0] = flags[0].data_ptr(); sig.flags[1] = flags[1].data_ptr(); sig.n_ranks, sig.my_index, sig.epoch = 2, 0, 1
ctx.specular_prefilter_ranges(pyr, [vq.cubemap_of(spec, 16, 4), vq.cubemap_of(torch.zeros_like(spec), 16, 4)], [(0, 96)], 64, signal=sig)
ctx.forward_lighting_multi(pf, pv, gb, em, [vq.image_of(out), vq.image_of(torch.zeros_like(out))], 0, signal=sig)
torch.cuda.synchronize(); assert int(flags[1, 0]) == 1, "the rendezvous did not signal the peer"
# the shadowed pass and the depth pyramid
pfs, pvs = synth.scene_constants(64, 40, 4, n_point=2, n_spot=1, casters=True)
pfs.Lights.directional.shadowing = 1
for sc_ in range(pfs.Lights.numSpotCasters):
    pfs.Lights.shadowViews[sc_].m[0] = pfs.Lights.shadowViews[sc_].m[5] = 0.04; pfs.Lights.shadowViews[sc_].m[14] = 0.5; pfs.Lights.shadowViews[sc_].m[15] = 1.0
pfs.Lights.shadowViewDirectional.m[0] = pfs.Lights.shadowViewDirectional.m[5] = 0.04; pfs.Lights.shadowViewDirectional.m[14] = 0.5; pfs.Lights.shadowViewDirectional.m[15] = 1.0
pfs.f2SpotLightShadowMapDimensions.x = pfs.f2SpotLightShadowMapDimensions.y = 16.0
pfs.f2DirectionalLightShadowMapDimensions.x = pfs.f2DirectionalLightShadowMapDimensions.y = 16.0
gb3 = vq.GBuffer(vq.image_of(planes[0]), vq.image_of(planes[1]), vq.image_of(planes[2]), vq.null_image())
ctx.forward_lighting_shadowed(pfs, pvs, gb3, em, out, torch.rand((max(pfs.Lights.numPointCasters, 1), 6, 8, 8), device="cuda"),
                              torch.rand((max(pfs.Lights.numSpotCasters, 1), 16, 16), device="cuda"), torch.rand((16, 16), device="cuda"))
depth = torch.rand((37, 53), device="cuda"); nlev = vq.depth_pyramid_level_count(53, 37)
ctx.depth_min_pyramid(depth, torch.empty((vq.depth_pyramid_texel_count(53, 37, nlev),), device="cuda"))
depth2 = torch.rand((131, 203), device="cuda"); nlev2 = vq.depth_pyramid_level_count(203, 131)      # 8 levels: two launches, padded level 6 in between
ctx.depth_min_pyramid(depth2, torch.empty((vq.depth_pyramid_texel_count(203, 131, nlev2),), device="cuda"))
# SURVEY 8(f).1 (bench.surface_scene_gpu builds the mip chains with vq_texture_build_mips and a material table)
import bench
sc = bench.surface_scene_gpu(ctx, vq, torch, 64, 38, n_materials=3, tex_res=64)
g4 = [torch.empty((38, 64, 4), device="cuda") for _ in range(4)]
ctx.gbuffer_from_materials(sc["inputs"], sc["table"], 0.3, vq.GBuffer(*(vq.image_of(t) for t in g4)))      # texel records
os.environ["VQ_SURFACE_RECORDS"] = "0"
t_maps = ctx.material_table(sc["mats"], sc["mts"]); os.environ.pop("VQ_SURFACE_RECORDS")
ctx.gbuffer_from_materials(sc["inputs"], t_maps, 0.3, vq.GBuffer(*(vq.image_of(t) for t in g4)), alpha_mask=True)   # map by map
# SURVEY 8(f).2 / (f).3
src = dev(synth.hdri(96, 20)); src8 = dev(synth.hdri(6, 5))
for im in (src, src8):
    data = ctx.hdr_save_host(im)
    dec, lum = ctx.hdr_decode(data)
    ctx.hdr_load_host(data, torch.zeros_like(im))
small = torch.zeros((7, 41, 4), device="cuda"); ctx.image_resize(src, small)
_, inv = synth.sky_view_proj(0.3, 0.1, 1.0, 64 / 40)
ctx.skydome(inv.astype(np.float32).reshape(16), pyr, out, normal_mask=planes[1])
ctx.skydome(inv.astype(np.float32).reshape(16), pyr, out)
ctx.apply_reflections(out, torch.rand_like(out)); ctx.apply_reflections(out, torch.rand_like(out), torch.rand_like(out))
torch.cuda.synchronize()
print("sanitize_small: all kernels launched, finite:", bool(torch.isfinite(out).all()))
