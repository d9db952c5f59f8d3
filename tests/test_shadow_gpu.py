"""-m gpu: SURVEY §8(f).4 — vq_forward_lighting_shadowed and vq_depth_min_pyramid against the oracle.

First run on a B200 at the start of round 2 (all green, memcheck clean: profiles/r02_shadow_first_run.txt). The oracle side
of every comparison is pinned against the reference's shader text (tests/test_hlsl_ref.py: shadowed PSMain, DownsampleDepth.hlsl)."""
import numpy as np
import pytest
import torch

from gpu_util import dev, host, report, TOL
from envmaps import small_env

pytestmark = pytest.mark.gpu


def _scene(w, h, seed):
    from vqengine_b200 import synth
    env = small_env()
    planes = synth.gbuffer(w, h, seed=seed)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=2, n_spot=1, casters=True)
    L = pf.Lights
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0; m[1] = 0.01; m[4] = -0.02; m[12] = 0.1
    for sc in range(L.numSpotCasters):
        for k in range(16): L.shadowViews[sc].m[k] = float(m[k]) * (1.0 + 0.3 * sc)
    for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
    L.directional.shadowing = 1
    pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = 16.0
    pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = 16.0
    return env, planes, pf, pv


def _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap, rows=None):
    h, w = planes[0].shape[:2]
    dplanes = [dev(p) for p in planes[:3]]
    gb = vq.GBuffer(vq.image_of(dplanes[0]), vq.image_of(dplanes[1]), vq.image_of(dplanes[2]), vq.null_image())
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]), vq.image_of(dl, 2))
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    rb, re = rows if rows else (0, h)
    ctx.forward_lighting_shadowed(pf, pv, gb, em, out,
                                  point_cubes=dev(cubes) if cubes is not None else None,
                                  spot_maps=dev(spots) if spots is not None else None,
                                  directional_map=dev(dmap) if dmap is not None else None, row_begin=rb, row_end=re)
    return host(out)


def _oracle_pass(orc, env, planes, pf, pv, cubes, spots, dmap):
    return orc.forward_lighting_shadowed(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"],
                                         env["lut"], point_cubes=cubes, point_res=cubes.shape[2] if cubes is not None else 0,
                                         spot_maps=spots, dir_map=dmap)


def _assert_pcf(name, got, ref, max_flip_frac=2e-3):
    """|delta| <= TOL*max(1,|ref|) except where a PCF tap flipped on the last ulp of tan/acos/pow: those pixels are few and
    differ by whole multiples of one tap's weight; bound their number, not their size."""
    assert np.isfinite(got).all(), f"{name}: non-finite output"
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bad = (d > TOL * np.maximum(1.0, np.abs(ref))).any(axis=-1)
    assert bad.mean() <= max_flip_frac, f"{name}: {bad.sum()} of {bad.size} pixels outside tolerance ({report(name, got, ref)})"
    return bad.mean()


@pytest.mark.parametrize("w,h", [(96, 54), (33, 7), (1, 1)])
def test_shadowed_pass_matches_oracle(ctx, vq, orc, w, h):
    env, planes, pf, pv = _scene(w, h, 6)
    L = pf.Lights
    rng = np.random.default_rng(110)
    cubes = rng.uniform(0.0, 1.2, (L.numPointCasters, 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (L.numSpotCasters, 16, 16)).astype(np.float32)
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    got = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap)
    ref = _oracle_pass(orc, env, planes, pf, pv, cubes, spots, dmap)
    print(_assert_pcf(f"shadowed{w}x{h}", got, ref))
    assert (got[..., 3] == ref[..., 3]).all()


def test_shadowed_pass_full_caster_lists(ctx, vq, orc):
    """5 point + 5 spot casters + the directional light: all eleven 5-bit slots of the per-pixel PCF record, non-square
    non-power-of-two spot maps, a rotated HDRI (the ROT instantiation of the shadowed kernel) and a ragged row tile"""
    from shadow_util import fill_casters
    w, h = 300, 41
    env, planes, pf, pv = _scene(w, h, 21)
    fill_casters(pf, 5, 5, seed=2)
    pf.fHDRIOffsetInRadians = 0.7
    rng = np.random.default_rng(7)
    cubes = rng.uniform(0.0, 1.2, (5, 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (5, 12, 20)).astype(np.float32)
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    pf.f2SpotLightShadowMapDimensions.x, pf.f2SpotLightShadowMapDimensions.y = 20.0, 12.0
    got = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap)
    ref = _oracle_pass(orc, env, planes, pf, pv, cubes, spots, dmap)
    print(_assert_pcf("shadowed_full_lists", got, ref))
    lit = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    assert np.abs(lit - ref).max() > 1e-2                                  # the shadow tests bite


def test_nothing_occluded_equals_the_unshadowed_kernel(ctx, vq, orc):
    """all maps at 'nothing in front': the caster terms carry factor 1, so the result must agree with K1's own (factor 1) path"""
    w, h = 64, 36
    env, planes, pf, pv = _scene(w, h, 3)
    L = pf.Lights
    cubes = np.ones((L.numPointCasters, 6, 8, 8), np.float32)
    spots = np.ones((L.numSpotCasters, 16, 16), np.float32)
    dmap = np.ones((16, 16), np.float32)
    got = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap)
    ref = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    _assert_pcf("unoccluded", got, ref, max_flip_frac=0.0)
    # no maps bound at all: the same
    got2 = _device_pass(ctx, vq, env, planes, pf, pv, None, None, None)
    _assert_pcf("no_maps", got2, ref, max_flip_frac=0.0)


def test_everything_occluded_equals_pass_without_casters(ctx, vq, orc):
    w, h = 64, 36
    env, planes, pf, pv = _scene(w, h, 3)
    L = pf.Lights
    cubes = np.zeros((L.numPointCasters, 6, 8, 8), np.float32)
    spots = np.zeros((L.numSpotCasters, 16, 16), np.float32)
    dmap = np.zeros((16, 16), np.float32)
    got = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap)
    npc, nsc = L.numPointCasters, L.numSpotCasters
    L.numPointCasters = 0; L.numSpotCasters = 0; L.directional.enabled = 0
    ref = orc.forward_lighting(pf, pv, planes, env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    L.numPointCasters = npc; L.numSpotCasters = nsc; L.directional.enabled = 1
    _assert_pcf("occluded", got, ref, max_flip_frac=0.0)


def test_row_range_only_touches_its_rows(ctx, vq, orc):
    w, h = 48, 20
    env, planes, pf, pv = _scene(w, h, 9)
    L = pf.Lights
    rng = np.random.default_rng(111)
    cubes = rng.uniform(0.0, 1.2, (L.numPointCasters, 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (L.numSpotCasters, 16, 16)).astype(np.float32)
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    full = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap)
    part = _device_pass(ctx, vq, env, planes, pf, pv, cubes, spots, dmap, rows=(5, 13))
    assert np.array_equal(part[5:13], full[5:13])
    assert (part[:5] == 0).all() and (part[13:] == 0).all()


@pytest.mark.parametrize("w,h", [(128, 128), (200, 120), (65, 33), (31, 70), (256, 16), (5, 3), (1, 1), (1920, 1080), (3840, 2160),
                                 (4097, 3), (8192, 66), (63, 129), (64, 64)])
def test_depth_min_pyramid_bit_exact(ctx, vq, orc, w, h):
    rng = np.random.default_rng(160 + w)
    depth = rng.uniform(0.05, 1.0, (h, w)).astype(np.float32)
    want = orc.depth_min_pyramid(depth)
    n = vq.depth_pyramid_level_count(w, h)
    assert n == len(want)
    total = vq.depth_pyramid_texel_count(w, h, n)
    assert total == sum(lv.size for lv in want)
    d = torch.from_numpy(depth).cuda()
    levels = torch.full((total,), -1.0, dtype=torch.float32, device="cuda")
    ctx.depth_min_pyramid(d, levels)
    got = host(levels)
    o = 0
    for l, lv in enumerate(want):
        g = got[o:o + lv.size].reshape(lv.shape)
        o += lv.size
        assert np.array_equal(g.view(np.uint32), lv.view(np.uint32)), l


def test_depth_min_pyramid_strided_rows_and_partial_levels(ctx, vq, orc):
    w, h = 100, 60
    rng = np.random.default_rng(170)
    big = torch.from_numpy(rng.uniform(0.05, 1.0, (h, 128)).astype(np.float32)).cuda()
    view = big[:, :w]                                                # pitch 128 floats
    want = orc.depth_min_pyramid(view.cpu().numpy())
    n = 3
    levels = torch.zeros((vq.depth_pyramid_texel_count(w, h, n),), dtype=torch.float32, device="cuda")
    ctx.depth_min_pyramid(view, levels, n_levels=n)
    got = host(levels)
    o = 0
    for lv in want[:n]:
        assert np.array_equal(got[o:o + lv.size].reshape(lv.shape), lv)
        o += lv.size
