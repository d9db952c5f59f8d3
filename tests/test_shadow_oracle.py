"""CPU (-m "not gpu"): SURVEY A25 / §8(f).4 groundwork — the oracle's shadow tests (oracle/oracle_shadow.cpp restating
Lighting.hlsl:79-272 and the caster section of PSMain, ForwardLighting.hlsl:321-377). The product path still lights casters
with shadow factor 1; these tests fix the semantics the CUDA path has to match once shadow maps become an input:
closed forms of the three PCF footprints and the two limits of the shadowed PSMain (nothing occluded == the unshadowed
pass bit for bit; everything occluded == the pass without casters)."""
import ctypes as C

import numpy as np
import pytest

from envmaps import small_env


def _scene(w=24, h=12, seed=3):
    from vqengine_b200 import synth
    env = small_env()
    planes = synth.gbuffer(w, h, seed=seed)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=2, n_spot=1, casters=True)
    L = pf.Lights
    # light-space transforms that keep the whole height field inside the frustum: x,y in [-20,20] -> [-0.8,0.8], depth 0.5
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0
    for sc in range(L.numSpotCasters):
        for k in range(16): L.shadowViews[sc].m[k] = float(m[k])
    for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
    L.directional.shadowing = 1
    pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = 16.0
    pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = 16.0
    return env, planes, pf, pv


def _args(env):
    return (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])


def test_nothing_occluded_equals_unshadowed_pass(orc):
    env, planes, pf, pv = _scene()
    L = pf.Lights
    assert L.numPointCasters >= 1 and L.numSpotCasters >= 1
    cubes = np.ones((L.numPointCasters, 6, 8, 8), np.float32)         # stored distance / range = 1: nothing closer than the range
    spots = np.ones((L.numSpotCasters, 16, 16), np.float32)           # light-space depth 1: nothing in front
    dmap = np.ones((16, 16), np.float32)
    lit = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), point_cubes=cubes, point_res=8, spot_maps=spots, dir_map=dmap)
    ref = orc.forward_lighting(pf, pv, planes, *_args(env))
    assert np.array_equal(lit.view(np.uint32), ref.view(np.uint32))
    # no maps bound at all == the same
    assert np.array_equal(orc.forward_lighting_shadowed(pf, pv, planes, *_args(env)).view(np.uint32), ref.view(np.uint32))


def test_everything_occluded_equals_pass_without_casters(orc):
    env, planes, pf, pv = _scene()
    L = pf.Lights
    cubes = np.zeros((L.numPointCasters, 6, 8, 8), np.float32)
    spots = np.zeros((L.numSpotCasters, 16, 16), np.float32)
    dmap = np.zeros((16, 16), np.float32)
    dark = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), point_cubes=cubes, point_res=8, spot_maps=spots, dir_map=dmap)
    npc, nsc = L.numPointCasters, L.numSpotCasters
    L.numPointCasters = 0; L.numSpotCasters = 0; L.directional.enabled = 0
    ref = orc.forward_lighting(pf, pv, planes, *_args(env))
    L.numPointCasters = npc; L.numSpotCasters = nsc; L.directional.enabled = 1
    # the occluded terms are added as exact zeros (x * 0), so the sums agree bit for bit
    assert np.array_equal(dark.view(np.uint32), ref.view(np.uint32))
    lit = orc.forward_lighting(pf, pv, planes, *_args(env))
    assert (lit[..., :3] >= dark[..., :3]).all() and (lit[..., :3] > dark[..., :3]).any()


def test_spot_outside_frustum_is_shadowed(orc):
    env, planes, pf, pv = _scene()
    L = pf.Lights
    for k in (0, 5):
        L.shadowViews[0].m[k] = 10.0                                  # every pixel lands outside [-1,1]^2
    spots = np.ones((L.numSpotCasters, 16, 16), np.float32)
    out = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), spot_maps=spots)
    nsc = L.numSpotCasters
    L.numSpotCasters = 0
    ref = orc.forward_lighting(pf, pv, planes, *_args(env))
    L.numSpotCasters = nsc
    inside = (np.abs(planes[0][..., 0]) * 10.0 <= 1.0) & (np.abs(planes[0][..., 1]) * 10.0 <= 1.0)
    assert np.array_equal(out[~inside].view(np.uint32), ref[~inside].view(np.uint32))


def test_pcf_footprint_fractions(orc):
    """5x5 POINT taps one shadow-map texel apart; a vertical occluder edge through the footprint shadows whole columns"""
    n = 16
    m = np.ones((n, n), np.float32); m[:, :8] = 0.0                   # left half holds an occluder at depth 0
    for col, want in [(2, 0.0), (12, 1.0), (8, 0.6), (9, 0.8), (7, 0.4), (6, 0.2), (10, 1.0), (5, 0.0)]:
        u = (col + 0.5) / n; v = 0.5
        lsp = (2 * u - 1, 1 - 2 * v, 0.5, 1.0)                        # clip space of texel centre (col, 8)
        got = orc.shadow_test_pcf(lsp, 0.0, 1.0, m)
        assert got == pytest.approx(want, abs=1e-6), (col, got, want)
        assert orc.shadow_test_pcf(lsp, 0.0, 1.0, m, directional=True) == pytest.approx(want, abs=1e-6)
    # slope-scaled bias (spot): depth 0.5 against a map of 0.45 is lit once bias*tan(acos(NdotL)) exceeds 0.05
    m2 = np.full((n, n), 0.45, np.float32)
    assert orc.shadow_test_pcf((0, 0, 0.5, 1), 0.01, 0.5, m2) == 0.0                       # tan(60 deg)*0.01 = 0.017 < 0.05
    assert orc.shadow_test_pcf((0, 0, 0.5, 1), 0.05, 0.5, m2) == 1.0                       # 0.087 > 0.05
    assert orc.shadow_test_pcf((0, 0, 0.5, 1), 0.04, 0.5, m2, directional=True) == 0.0     # constant bias 0.04 < 0.05
    assert orc.shadow_test_pcf((0, 0, 0.5, 1), 0.06, 0.5, m2, directional=True) == 1.0
    # outside the frustum, and WRAP addressing of the taps at the border
    assert orc.shadow_test_pcf((1.5, 0, 0.5, 1), 0.0, 1.0, np.ones((n, n), np.float32)) == 0.0
    assert orc.shadow_test_pcf((0, 0, 1.5, 1), 0.0, 1.0, np.ones((n, n), np.float32)) == 0.0
    edge = orc.shadow_test_pcf((2 * (15.5 / n) - 1, 0.0, 0.5, 1.0), 0.0, 1.0, m)          # taps at columns 13..17 -> 16,17 wrap to 0,1 (occluded)
    assert edge == pytest.approx(0.6, abs=1e-6)


def test_omnidirectional_tap_fractions(orc):
    res, far = 8, 10.0
    lit = np.ones((6, res, res), np.float32)
    assert orc.shadow_test_omni((0, 0, 5.0), 0.0, 0.0, far, lit) == 1.0       # stored depth = range: nothing occludes
    assert orc.shadow_test_omni((0, 0, 5.0), 0.0, 0.0, far, lit * 0.3) == 0.0  # occluder at 3 < 5 everywhere
    assert orc.shadow_test_omni((0, 0, 5.0), 0.0, 0.0, far, lit * 0.5) == 1.0  # exactly at the pixel's distance: 5 > 5 + 0.001 is false
    # the sample vectors are -(Lw + dir*radius): for Lw = (0,0,5) all 20 land on the -Z face; occlude only the half with x' < 0
    half = lit.copy()
    half[5, :, : res // 2] = 0.0
    got = orc.shadow_test_omni((0, 0, 5.0), 0.0, 0.0, far, half)
    assert min(abs(got - k / 20) for k in range(21)) < 1e-6 and 0.0 < got < 1.0
    # disk radius grows with the view distance: (1 + d/far)/8
    far_taps = orc.shadow_test_omni((0, 0, 0.3), 0.0, 70.0, far, half)
    assert min(abs(far_taps - k / 20) for k in range(21)) < 1e-6
