"""-m gpu parity: K1 forward PBR lighting against the scalar oracle, through the C-ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import dev, host, assert_scaled, report, TOL
from envmaps import small_env

pytestmark = pytest.mark.gpu


def _run(ctx, vq, orc, w, h, *, n_point=4, n_spot=0, directional=True, emissive=False, offset=0.0, casters=False,
         diffuse_only=False, point_range=50.0, seed=3, rows=None):
    from vqengine_b200 import synth
    env = small_env()
    planes = synth.gbuffer(w, h, seed=seed, emissive=emissive)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=n_point, n_spot=n_spot,
                                   directional=directional, hdri_offset=offset, casters=casters, point_range=point_range)
    pv.EnvironmentMapDiffuseOnlyIllumination = int(diffuse_only)
    dplanes = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dplanes[0]), vq.image_of(dplanes[1]), vq.image_of(dplanes[2]),
                    vq.image_of(dplanes[3]) if emissive else vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]),
                            vq.image_of(dl, 2))
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    rb, re = rows if rows else (0, h)
    ctx.forward_lighting(pf, pv, gb, em, out, rb, re)
    ref = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"],
                               env["spec_mips"], env["lut"], rb, re)
    return host(out), ref


@pytest.mark.parametrize("w,h", [(64, 36), (1, 1), (257, 3), (480, 270)])
def test_forward_config3_shape(ctx, vq, orc, w, h):
    """4 point + 1 directional + IBL (BASELINE config 3 lighting) at small sizes."""
    got, ref = _run(ctx, vq, orc, w, h)
    r = assert_scaled(f"forward{w}x{h}", got, ref)
    assert (got[..., 3] == ref[..., 3]).all()      # alpha = roughness passthrough
    print(r)


def test_forward_prepared_environment_equals_per_call_padding(ctx, vq, orc):
    """vq_environment_prepare (bordered cube copies built once) must not change a single bit of the result,
    and a stale registration must not be used for different maps."""
    from vqengine_b200 import synth
    env = small_env()
    w, h = 96, 54
    planes = synth.gbuffer(w, h, seed=5)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=5)
    dplanes = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dplanes[0]), vq.image_of(dplanes[1]), vq.image_of(dplanes[2]), vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    a, b, c = (torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(3))
    ctx.environment_invalidate()
    ctx.forward_lighting(pf, pv, gb, em, a)
    ctx.environment_prepare(em)
    ctx.forward_lighting(pf, pv, gb, em, b)
    assert np.array_equal(host(a), host(b))
    # different maps at different addresses: the registration does not match -> padded per call -> different result
    ds2 = dev(env["spec"] * 0.5)
    em2 = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds2, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    ctx.forward_lighting(pf, pv, gb, em2, c)
    assert not np.array_equal(host(c), host(b))
    ctx.environment_invalidate()


def test_forward_all_light_types(ctx, vq, orc):
    got, ref = _run(ctx, vq, orc, 160, 90, n_point=7, n_spot=3, emissive=True, offset=0.7, casters=True)
    assert_scaled("forward_all", got, ref)


def test_forward_no_lights_ibl_only(ctx, vq, orc):
    got, ref = _run(ctx, vq, orc, 96, 54, n_point=0, directional=False)
    assert_scaled("forward_ibl", got, ref)


def test_forward_diffuse_only_flag(ctx, vq, orc):
    got, ref = _run(ctx, vq, orc, 96, 54, diffuse_only=True)
    assert_scaled("forward_diffonly", got, ref)


def test_forward_range_cut(ctx, vq, orc):
    """`D < l.range` is a discontinuity: a range that cuts through the field must give the same mask."""
    got, ref = _run(ctx, vq, orc, 320, 180, point_range=14.0, directional=False)
    assert_scaled("forward_range", got, ref)


def test_forward_max_lights(ctx, vq, orc):
    got, ref = _run(ctx, vq, orc, 48, 27, n_point=100, n_spot=20)
    assert_scaled("forward_maxlights", got, ref)


def test_forward_row_tiles(ctx, vq, orc):
    """rows [a,b) only: untouched rows stay zero, touched rows equal the full-frame result."""
    full, _ = _run(ctx, vq, orc, 64, 40)
    part, ref = _run(ctx, vq, orc, 64, 40, rows=(13, 29))
    assert np.array_equal(part[13:29], full[13:29])
    assert (part[:13] == 0).all() and (part[29:] == 0).all()


def test_single_pixel_kat(ctx, vq):
    """BASELINE config 1 / SURVEY.md §8(c): N=V=Wi=(0,0,1), albedo .5, rough .5, metal 0, white point light
    at distance 2, brightness 10 -> 0.50927037 per channel (ambient 0, IBL maps = 0)."""
    pf = vq.PerFrameData(); pv = vq.PerViewLightingData()
    pf.Lights.numPointLights = 1
    l = pf.Lights.point_lights[0]
    l.position.z = 2.0; l.range = 100.0; l.brightness = 10.0
    l.color.x = l.color.y = l.color.z = 1.0
    pv.CameraPosition.z = 5.0; pv.MaxEnvMapLODLevels = 2.0
    pos = dev(np.array([[[0, 0, 0, 0]]], np.float32)); nrm = dev(np.array([[[0, 0, 1, 0.5]]], np.float32))
    alb = dev(np.array([[[0.5, 0.5, 0.5, 0.0]]], np.float32))
    zeros_d = torch.zeros((6 * 4, 4), device="cuda"); zeros_s = torch.zeros((6 * 4 + 6, 4), device="cuda")
    lut = torch.zeros((2, 2, 2), device="cuda")
    gb = vq.GBuffer(vq.image_of(pos), vq.image_of(nrm), vq.image_of(alb), vq.null_image())
    em = vq.EnvironmentMaps(vq.cubemap_of(zeros_d, 2, 1), vq.cubemap_of(zeros_s, 2, 2), vq.image_of(lut, 2))
    out = torch.zeros((1, 1, 4), device="cuda")
    ctx.forward_lighting(pf, pv, gb, em, out)
    o = host(out)[0, 0]
    assert np.allclose(o[:3], 0.50927037, atol=3e-7), o
    assert o[3] == 0.5


def test_forward_host_entry(ctx, vq, orc):
    """the blocking host-buffer call (uploads, shades, downloads) equals the device call."""
    from vqengine_b200 import synth
    env = small_env()
    w, h = 200, 300
    planes = synth.gbuffer(w, h, seed=8)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=8)
    hp = [torch.from_numpy(p).pin_memory() for p in planes]
    gbh = vq.GBuffer(vq.image_of(hp[0]), vq.image_of(hp[1]), vq.image_of(hp[2]), vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]),
                            vq.image_of(dl, 2))
    hout = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
    ctx.forward_lighting_host(pf, pv, gbh, em, hout)
    dplanes = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dplanes[0]), vq.image_of(dplanes[1]), vq.image_of(dplanes[2]), vq.null_image())
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, em, out)
    assert np.array_equal(hout.numpy(), host(out))


def test_forward_host_entry_rejects_mismatched_planes(ctx, vq):
    """every host plane is copied with the OUTPUT's width/height: a smaller G-buffer plane or a bad emissive descriptor must
    be refused before any copy is enqueued (it would read past the end of the caller's host memory)"""
    from vqengine_b200 import synth
    env = small_env()
    w, h = 64, 48
    planes = synth.gbuffer(w, h, seed=8, emissive=True)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=8)
    hp = [torch.from_numpy(p) for p in planes]
    small = torch.zeros((h // 2, w, 4), dtype=torch.float32)
    narrow = torch.zeros((h, w // 2, 4), dtype=torch.float32)
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    hout = torch.zeros((h, w, 4), dtype=torch.float32)
    for bad in (vq.GBuffer(vq.image_of(hp[0]), vq.image_of(small), vq.image_of(hp[2]), vq.null_image()),
                vq.GBuffer(vq.image_of(narrow), vq.image_of(hp[1]), vq.image_of(hp[2]), vq.null_image()),
                vq.GBuffer(vq.image_of(hp[0]), vq.image_of(hp[1]), vq.image_of(hp[2]), vq.image_of(small))):
        with pytest.raises(vq.VqError) as e:
            ctx.forward_lighting_host(pf, pv, bad, em, hout)
        assert e.value.code == vq.VQ_ERR_INVALID_ARG
    badpitch = vq.image_of(hp[3]); badpitch.pitch_bytes = w * 16 - 16
    with pytest.raises(vq.VqError):
        ctx.forward_lighting_host(pf, pv, vq.GBuffer(vq.image_of(hp[0]), vq.image_of(hp[1]), vq.image_of(hp[2]), badpitch), em, hout)
    ctx.forward_lighting_host(pf, pv, vq.GBuffer(*(vq.image_of(t) for t in hp)), em, hout)     # the good call still works
    assert np.isfinite(hout.numpy()).all() and hout.numpy().any()


def test_forward_pitched_planes_and_output(ctx, vq, orc):
    from vqengine_b200 import synth
    env = small_env()
    w, h = 70, 33
    planes = synth.gbuffer(w, h, seed=12)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=12)
    bigs = []
    views = []
    for p in planes:
        big = torch.full((h, w + 6, 4), 3.0, dtype=torch.float32, device="cuda"); big[:, :w] = dev(p)
        bigs.append(big); views.append(big[:, :w])
    gb = vq.GBuffer(vq.image_of(views[0]), vq.image_of(views[1]), vq.image_of(views[2]), vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    obig = torch.full((h, w + 4, 4), -1.0, dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, em, obig[:, :w])
    ref = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    assert_scaled("forward_pitched", host(obig[:, :w].contiguous()), ref)
    assert (host(obig)[:, w:] == -1.0).all()


def test_forward_empty_row_range_is_a_noop(ctx, vq, orc):
    got, _ = _run(ctx, vq, orc, 32, 16, rows=(7, 7))
    assert (got == 0).all()


def test_forward_rejects_bad_arguments(ctx, vq):
    pf = vq.PerFrameData(); pv = vq.PerViewLightingData()
    pf.Lights.numPointLights = 101           # beyond the cbuffer array
    t = torch.zeros((4, 4, 4), device="cuda")
    gb = vq.GBuffer(vq.image_of(t), vq.image_of(t), vq.image_of(t), vq.null_image())
    c = torch.zeros((6 * 4, 4), device="cuda"); lut = torch.zeros((2, 2, 2), device="cuda")
    em = vq.EnvironmentMaps(vq.cubemap_of(c, 2, 1), vq.cubemap_of(torch.zeros((6 * 4 + 6, 4), device="cuda"), 2, 2), vq.image_of(lut, 2))
    with pytest.raises(vq.VqError) as e:
        ctx.forward_lighting(pf, pv, gb, em, torch.zeros((4, 4, 4), device="cuda"))
    assert e.value.code == vq.VQ_ERR_INVALID_ARG
    pf.Lights.numPointLights = 0
    with pytest.raises(vq.VqError):
        ctx.forward_lighting(pf, pv, gb, em, torch.zeros((5, 4, 4), device="cuda"))     # size mismatch


def test_forward_multi_destination_store(ctx, vq, orc):
    """vq_forward_lighting_multi: the tile is written to row dst_row_offset+y of EVERY destination frame (on the GPU box
    the extra destinations are peer-GPU frames over NVLink; here two local frames stand in for them)."""
    from vqengine_b200 import synth
    env = small_env()
    w, h = 80, 24
    planes = synth.gbuffer(w, h, seed=21)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=21)
    dplanes = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dplanes[0]), vq.image_of(dplanes[1]), vq.image_of(dplanes[2]), vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    single = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, em, single)
    frames = [torch.zeros((3 * h, w, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
    ctx.forward_lighting_multi(pf, pv, gb, em, [vq.image_of(f) for f in frames], dst_row_offset=h)
    for f in frames:
        got = host(f)
        assert np.array_equal(got[h:2 * h], host(single)) and (got[:h] == 0).all() and (got[2 * h:] == 0).all()
    with pytest.raises(vq.VqError):      # destination too small for offset + tile
        ctx.forward_lighting_multi(pf, pv, gb, em, [vq.image_of(frames[0])], dst_row_offset=2 * h + 1)


def test_forward_full_size_4k_properties(ctx, vq, orc):
    """BASELINE full size (3840x2160, 4 point + 1 directional + IBL at 64^2 / 512^2 x9 / 1024^2): size-independent
    properties — row tiling is bit-invariant (the multi-GPU partition), alpha passes roughness through, everything is
    finite — plus parity against the oracle on two 12-row bands (prepared sampling copies, persistent grid with all
    30 x 29 CTAs in play)."""
    import bench
    from vqengine_b200 import synth
    envk = bench.build_env_maps_gpu(ctx, vq, torch)
    w, h = 3840, 2160
    planes = synth.gbuffer(w, h)
    pf, pv = synth.scene_constants(w, h, envk["spec_mips"])
    dpl = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    whole = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, envk["env"], whole)
    tiled = torch.zeros_like(whole)
    for rb, re in ((0, 541), (541, 1080), (1080, 1081), (1081, 2160)):
        ctx.forward_lighting(pf, pv, gb, envk["env"], tiled, rb, re)
    torch.cuda.synchronize()
    assert torch.equal(whole, tiled)
    assert bool(torch.isfinite(whole).all())
    assert torch.equal(whole[..., 3], dpl[1][..., 3])
    got = host(whole)
    env_np = {k: host(envk[k]) for k in ("diff", "spec", "lut")}
    for r0 in (7, 1500):
        ref = orc.forward_lighting(pf, pv, planes, env_np["diff"], envk["diff_res"], env_np["spec"], envk["spec_res"],
                                   envk["spec_mips"], env_np["lut"], r0, r0 + 12)
        assert_scaled(f"forward4k rows {r0}", got[r0:r0 + 12], ref[r0:r0 + 12])
    ctx.environment_invalidate()
