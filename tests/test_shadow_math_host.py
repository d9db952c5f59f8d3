"""CPU (-m "not gpu"): SURVEY §8(f).4 — the product's per-pixel shadow math, compiled for the HOST, against the oracle.

vqengine_b200/csrc/vq_shadow_math.cuh holds everything the CUDA kernels of vq_shadow.cu compute per pixel / per texel (BRDF and
light functions with individually rounded operations, the three PCF tests, cube / 2-D point taps, the caster loop of PSMain,
the MIN-pyramid texel) plus the launcher's set-up code (light block, per-frame copy without casters, level plan). The same
header builds with g++ when VQ_HOST_CHECK is defined (tests/host_check/shadow_math_host.cpp), so the restatement is checked
here bit for bit without a GPU; what remains for the GPU run (tests/test_shadow_gpu.py) is the launch code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as orc
from envmaps import small_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck") / "libshadowhost.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", str(out),
                           os.path.join(ROOT, "tests", "host_check", "shadow_math_host.cpp")])
    return C.CDLL(str(out))


def _scene(w, h, seed, n_point=2, n_spot=1):
    import vqengine_b200 as vq
    from vqengine_b200 import synth
    env = small_env()
    planes = synth.gbuffer(w, h, seed=seed)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=n_point, n_spot=n_spot, casters=True)
    L = pf.Lights
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0; m[1] = 0.01; m[4] = -0.02; m[12] = 0.1
    for sc in range(L.numSpotCasters):
        for k in range(16): L.shadowViews[sc].m[k] = float(m[k]) * (1.0 + 0.3 * sc)
    for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
    L.directional.shadowing = 1
    pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = 16.0
    pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = 16.0
    return vq, env, planes, pf, pv


def _args(env):
    return (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])


def _host_pass(hostlib, vq, env, planes, pf, pv, cubes, spots, dmap):
    base_pf = vq.PerFrameData()
    hostlib.hostcheck_per_frame_without_casters(C.byref(pf), C.byref(base_pf))
    assert base_pf.Lights.numPointCasters == 0 and base_pf.Lights.numSpotCasters == 0 and base_pf.Lights.directional.enabled == 0
    assert base_pf.Lights.numPointLights == pf.Lights.numPointLights and base_pf.fAmbientLightingFactor == pf.fAmbientLightingFactor
    base = orc.forward_lighting(base_pf, pv, planes, *_args(env))           # stands in for K1 (its own parity tests are -m gpu)
    sm = vq.ShadowMaps()
    keep = [np.ascontiguousarray(a, np.float32) if a is not None else None for a in (cubes, spots, dmap)]
    if keep[0] is not None: sm.point_cubes, sm.point_res = keep[0].ctypes.data, keep[0].shape[2]
    if keep[1] is not None: sm.spot_maps, sm.spot_width, sm.spot_height = keep[1].ctypes.data, keep[1].shape[2], keep[1].shape[1]
    if keep[2] is not None: sm.directional_map, sm.directional_width, sm.directional_height = keep[2].ctypes.data, keep[2].shape[1], keep[2].shape[0]
    h, w = planes[0].shape[:2]
    out = np.zeros((h, w, 4), np.float32)
    p = [np.ascontiguousarray(a, np.float32) for a in planes[:3]]
    hostlib.hostcheck_shade_casters(C.byref(pf), C.byref(pv), C.byref(sm), orc._p(p[0]), orc._p(p[1]), orc._p(p[2]), orc._p(base),
                                    C.c_int(w * h), orc._p(out))
    return out


@pytest.mark.parametrize("seed,n_point,n_spot", [(6, 2, 1), (9, 5, 5), (12, 0, 0)])
def test_caster_math_equals_oracle_bit_for_bit(hostlib, seed, n_point, n_spot):
    w, h = 96, 54
    vq, env, planes, pf, pv = _scene(w, h, seed, n_point, n_spot)
    L = pf.Lights
    assert L.directional.enabled
    rng = np.random.default_rng(110 + seed)
    cubes = rng.uniform(0.0, 1.2, (max(L.numPointCasters, 1), 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (max(L.numSpotCasters, 1), 16, 16)).astype(np.float32)
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    got = _host_pass(hostlib, vq, env, planes, pf, pv, cubes, spots, dmap)
    want = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), point_cubes=cubes, point_res=8, spot_maps=spots, dir_map=dmap)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.argwhere(got.view(np.uint32) != want.view(np.uint32))[:5]
    lit = orc.forward_lighting(pf, pv, planes, *_args(env))
    assert not np.array_equal(lit, want)                                      # the shadow tests bite


def test_caster_math_full_caster_lists_bit_for_bit(hostlib):
    """5 point casters + 5 spot casters + the directional light: every slot of the per-pixel record in use"""
    from shadow_util import fill_casters
    w, h = 80, 45
    vq, env, planes, pf, pv = _scene(w, h, 21, 1, 1)
    fill_casters(pf, 5, 5, seed=2)
    rng = np.random.default_rng(7)
    cubes = rng.uniform(0.0, 1.2, (5, 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (5, 12, 20)).astype(np.float32)        # non-square, not a power of two
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    pf.f2SpotLightShadowMapDimensions.x, pf.f2SpotLightShadowMapDimensions.y = 20.0, 12.0
    got = _host_pass(hostlib, vq, env, planes, pf, pv, cubes, spots, dmap)
    want = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), point_cubes=cubes, point_res=8, spot_maps=spots, dir_map=dmap)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the records the device stores: 5 bits per slot, counts within their tap numbers, something in every slot somewhere
    sm = vq.ShadowMaps()
    sm.point_cubes, sm.point_res = cubes.ctypes.data, 8
    sm.spot_maps, sm.spot_width, sm.spot_height = spots.ctypes.data, 20, 12
    sm.directional_map, sm.directional_width, sm.directional_height = dmap.ctypes.data, 16, 16
    rec = np.zeros(w * h, np.uint64)
    p = [np.ascontiguousarray(a, np.float32) for a in planes[:2]]
    hostlib.hostcheck_pcf_records(C.byref(pf), C.byref(pv), C.byref(sm), orc._p(p[0]), orc._p(p[1]), C.c_int(w * h),
                                  rec.ctypes.data_as(C.c_void_p))
    assert (rec >> np.uint64(55)).max() == 0
    for slot in range(11):
        c = (rec >> np.uint64(5 * slot)) & np.uint64(31)
        assert c.max() <= (20 if slot < 5 else 25) and c.max() > 0, slot


def test_caster_math_without_maps_equals_unshadowed_pass(hostlib):
    vq, env, planes, pf, pv = _scene(64, 36, 3)
    got = _host_pass(hostlib, vq, env, planes, pf, pv, None, None, None)
    want = orc.forward_lighting(pf, pv, planes, *_args(env))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_caster_math_directional_disabled_or_unshadowed(hostlib):
    vq, env, planes, pf, pv = _scene(48, 27, 4)
    dmap = np.full((16, 16), 0.4, np.float32)
    for enabled, shadowing in ((0, 1), (1, 0)):
        pf.Lights.directional.enabled, pf.Lights.directional.shadowing = enabled, shadowing
        got = _host_pass(hostlib, vq, env, planes, pf, pv, None, None, dmap)
        want = orc.forward_lighting_shadowed(pf, pv, planes, *_args(env), dir_map=dmap)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (enabled, shadowing)


@pytest.mark.parametrize("w,h", [(128, 128), (200, 120), (65, 33), (31, 70), (256, 16), (5, 3), (1, 1), (640, 360)])
def test_depth_pyramid_plan_and_texel_equal_oracle(hostlib, w, h):
    rng = np.random.default_rng(160 + w)
    pitch = w + 3
    big = rng.uniform(0.05, 1.0, (h, pitch)).astype(np.float32)
    want = orc.depth_min_pyramid(np.ascontiguousarray(big[:, :w]))
    total = sum(lv.size for lv in want)
    levels = np.full(total, -1.0, np.float32)
    n = hostlib.hostcheck_depth_min_pyramid(orc._p(big), C.c_int(pitch), C.c_int(w), C.c_int(h), orc._p(levels), C.c_int(len(want)))
    assert n == len(want)
    o = 0
    for l, lv in enumerate(want):
        assert np.array_equal(levels[o:o + lv.size].reshape(lv.shape).view(np.uint32), lv.view(np.uint32)), l
        o += lv.size
    assert hostlib.hostcheck_depth_min_pyramid(orc._p(big), C.c_int(pitch), C.c_int(w), C.c_int(h), orc._p(levels), C.c_int(len(want) + 1)) == -1
