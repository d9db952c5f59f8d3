"""CPU (-m "not gpu"): the oracle against the reference's OWN SHADER TEXT.

oracle/_ref/libhlslref.so is Shaders/ForwardLighting.hlsl (with Lighting.hlsl, BRDF.hlsl, ShadingMath.hlsl,
LightingConstantBufferData.h), Tonemapper.hlsl + HDR.hlsl, GaussianBlur.hlsl, CubemapConvolution.hlsl, Skydome.hlsl and
ApplyReflections.hlsl of /root/reference compiled with g++: oracle/ref_shim/hlsl_to_cpp.py respells the handful of tokens C++
cannot parse (build time, into the git-ignored oracle/_ref/obj/), oracle/ref_shim/hlsl_compat/hlsl_compat.h supplies the HLSL
types and intrinsics with the semantics oracle/hlsl_math.h documents, and every texture fetch is served by the oracle's
samplers. What these tests pin is therefore the oracle's restatement of the shaders — expression structure, operand order,
constants, branches, loop trip counts, data flow — bit for bit; what they cannot pin is D3D's unspecified intrinsic / filter
rounding, which is a documented decision on both sides.

Skipped when oracle/_ref was not built (no /root/reference at build time and no prebuilt copy)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as orc
from envmaps import small_env

pytestmark = pytest.mark.skipif(orc.hlsl_ref() is None, reason="oracle/_ref/libhlslref.so not built (needs /root/reference at build time)")

f32 = C.c_float


def _v(*x):
    return (f32 * len(x))(*[float(a) for a in x])


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(_bits(a), _bits(b)) or bool(np.all((_bits(a) == _bits(b)) | (np.isnan(a) & np.isnan(b))))


def _unit(rng, n=3):
    v = rng.normal(size=n).astype(np.float32)
    return v / np.float32(np.linalg.norm(v))


# ---- BRDF.hlsl / ShadingMath.hlsl / Lighting.hlsl, function by function -----------------------------------------------------
def test_brdf_terms_bit_exact():
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(101)
    for i in range(4000):
        N, V, L, H = _unit(rng), _unit(rng), _unit(rng), _unit(rng)
        F0 = rng.uniform(0, 1, 3).astype(np.float32)
        rough = np.float32(rng.uniform(0, 1)) if i % 50 else np.float32(0.0)
        ndh = np.float32(rng.uniform(0, 1)) if i % 37 else np.float32(1.0)
        assert _same(o.orc_ndf_ggx(f32(ndh), f32(rough)), r.hlslref_ndf_ggx(f32(ndh), f32(rough)))
        assert _same(o.orc_geometry_smith(_v(*N), _v(*V), _v(*L), f32(rough)), r.hlslref_geometry_smith(_v(*N), _v(*V), _v(*L), f32(rough)))
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        o.orc_fresnel_schlick(_v(*H), _v(*V), _v(*F0), orc._p(a)); r.hlslref_fresnel_schlick(_v(*H), _v(*V), _v(*F0), orc._p(b))
        assert _same(a, b)
        o.orc_fresnel_gaussian(_v(*H), _v(*V), _v(*F0), orc._p(a)); r.hlslref_fresnel_gaussian(_v(*H), _v(*V), _v(*F0), orc._p(b))
        assert _same(a, b)


def test_brdf_and_environment_brdf_bit_exact():
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(102)
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    for i in range(4000):
        N = _unit(rng) * np.float32(rng.uniform(0.5, 1.5))            # BRDF() renormalises N and V itself
        V = _unit(rng) * np.float32(rng.uniform(0.5, 1.5))
        Wi = _unit(rng)
        alb = rng.uniform(0, 1, 3).astype(np.float32)
        rough, metal = np.float32(rng.uniform(0, 1)), np.float32(rng.uniform(0, 1))
        if i % 97 == 0: rough = np.float32(0)
        if i % 89 == 0: metal = np.float32(1)
        o.orc_brdf(_v(*N), _v(*V), _v(*Wi), _v(*alb), f32(rough), f32(metal), orc._p(a))
        r.hlslref_brdf(_v(*N), _v(*V), _v(*Wi), _v(*alb), f32(rough), f32(metal), orc._p(b))
        assert _same(a, b), (i, a, b)
        irr, spec = rng.uniform(0, 4, 3).astype(np.float32), rng.uniform(0, 8, 3).astype(np.float32)
        sb = rng.uniform(0, 1, 2).astype(np.float32)
        ndv = np.float32(rng.uniform(0, 1))
        o.orc_environment_brdf(f32(ndv), f32(rough), f32(metal), _v(*alb), _v(*irr), _v(*spec), _v(*sb), orc._p(a))
        r.hlslref_environment_brdf(f32(ndv), f32(rough), f32(metal), _v(*alb), _v(*irr), _v(*spec), _v(*sb), orc._p(b))
        assert _same(a, b), (i, a, b)


def test_light_functions_bit_exact(vq):
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(103)
    a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
    for i in range(3000):
        P = rng.uniform(-10, 10, 3).astype(np.float32)
        N, V = _unit(rng), _unit(rng)
        alb = rng.uniform(0, 1, 3).astype(np.float32)
        rough, metal = np.float32(rng.uniform(0.02, 1)), np.float32(rng.uniform(0, 1))
        pl = vq.PointLight()
        pl.position.x, pl.position.y, pl.position.z = [float(x) for x in rng.uniform(-12, 12, 3)]
        pl.color.x, pl.color.y, pl.color.z = [float(x) for x in rng.uniform(0, 1, 3)]
        pl.range = float(rng.uniform(1, 30)); pl.brightness = float(rng.uniform(0, 500))
        o.orc_point_light(C.byref(pl), _v(*P), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(a))
        r.hlslref_point_light(C.byref(pl), _v(*P), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(b))
        assert _same(a, b), (i, a, b)
        sl = vq.SpotLight()
        sl.position.x, sl.position.y, sl.position.z = [float(x) for x in rng.uniform(-12, 12, 3)]
        sl.color.x, sl.color.y, sl.color.z = [float(x) for x in rng.uniform(0, 1, 3)]
        d = _unit(rng) * np.float32(rng.uniform(0.5, 2))
        sl.spotDir.x, sl.spotDir.y, sl.spotDir.z = [float(x) for x in d]
        sl.innerConeAngle = float(rng.uniform(0.1, 0.6)); sl.outerConeAngle = sl.innerConeAngle + float(rng.uniform(0.05, 0.8))
        sl.brightness = float(rng.uniform(0, 500)); sl.range = float(rng.uniform(1, 30))
        assert _same(o.orc_spotlight_intensity(C.byref(sl), _v(*P)), r.hlslref_spotlight_intensity(C.byref(sl), _v(*P)))
        o.orc_spot_light(C.byref(sl), _v(*P), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(a))
        r.hlslref_spot_light(C.byref(sl), _v(*P), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(b))
        assert _same(a, b), (i, a, b)
        dl = vq.DirectionalLight()
        d = _unit(rng) * np.float32(rng.uniform(0.5, 2))
        dl.lightDirection.x, dl.lightDirection.y, dl.lightDirection.z = [float(x) for x in d]
        dl.color.x, dl.color.y, dl.color.z = [float(x) for x in rng.uniform(0, 1, 3)]
        dl.brightness = float(rng.uniform(0, 20)); dl.enabled = 1
        o.orc_directional_light(C.byref(dl), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(a))
        r.hlslref_directional_light(C.byref(dl), _v(*N), _v(*V), _v(*alb), f32(rough), f32(metal), orc._p(b))
        assert _same(a, b), (i, a, b)


def test_sampling_math_bit_exact():
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(104)
    a2, b2 = np.zeros(2, np.float32), np.zeros(2, np.float32)
    a3, b3 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    for i in list(range(0, 3000)) + [2 ** 31, 2 ** 32 - 1, 0x55555555, 0xAAAAAAAA]:
        n = int(rng.integers(1, 5000))
        o.orc_hammersley(C.c_uint32(i), C.c_uint32(n), orc._p(a2)); r.hlslref_hammersley(C.c_uint32(i), C.c_uint32(n), orc._p(b2))
        assert _same(a2, b2)
    for i in range(3000):
        Xi = rng.uniform(0, 1, 2).astype(np.float32)
        N = _unit(rng) if i % 10 else np.array([0, 0, 1], np.float32)       # both branches of the tangent-frame choice
        rough = np.float32(rng.uniform(0, 1))
        o.orc_importance_sample_ggx(_v(*Xi), _v(*N), f32(rough), orc._p(a3)); r.hlslref_importance_sample_ggx(_v(*Xi), _v(*N), f32(rough), orc._p(b3))
        assert _same(a3, b3)
        d = rng.normal(size=3).astype(np.float32)
        d = d / np.float32(np.linalg.norm(d))
        o.orc_direction_to_equirect_uv(_v(*d), orc._p(a2)); r.hlslref_direction_to_equirect_uv(_v(*d), orc._p(b2))
        assert _same(a2, b2)
        s, n_, t = rng.uniform(0, 1, 3).astype(np.float32), _unit(rng), _unit(rng)
        b3[:] = 0
        r.hlslref_unpack_normal(_v(*s), _v(*n_), _v(*t), orc._p(b3))
        assert _same(orc.unpack_normal(s, n_, t), b3)
    for ndv, rough, n in ((0.5, 0.5, 64), (0.031, 0.97, 128), (0.999, 0.015, 96), (1.0, 0.0, 16), (0.25, 1.0, 2048)):
        o.orc_integrate_brdf(f32(ndv), f32(rough), C.c_int(n), orc._p(a2)); r.hlslref_integrate_brdf(f32(ndv), f32(rough), C.c_int(n), orc._p(b2))
        assert _same(a2, b2), (ndv, rough, n, a2, b2)


# ---- Tonemapper.hlsl + HDR.hlsl --------------------------------------------------------------------------------------------
def test_tonemapper_csmain_bit_exact(vq):
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(105)
    a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
    n = 0
    for curve in (0, 1, 2, 3):
        for space in (0, 1):
            for gamma in (0, 1):
                p = vq.TonemapperParams()
                p.ContentColorSpace, p.OutputDisplayCurve, p.ToggleGammaCorrection = space, curve, gamma
                p.DisplayReferenceBrightnessLevel = 200.0
                for i in range(400):
                    px = (rng.uniform(0, 1, 4) ** 3 * 20).astype(np.float32)
                    if i % 40 == 0: px[:3] = 0
                    if i % 41 == 0: px[0] = 0.0031308 / (1 - 0.0031308)         # Reinhard output right at the sRGB knee
                    o.orc_tonemap_pixel(C.byref(p), _v(*px), orc._p(a)); r.hlslref_tonemap_pixel(C.byref(p), _v(*px), orc._p(b))
                    assert _same(a, b), (curve, space, gamma, px, a, b)
                    n += 1
    assert n == 4 * 2 * 2 * 400


def test_hdr_curves_bit_exact():
    """HDR.hlsl:76-119, all five colour-space / transfer functions (the tonemapper only reaches three of them)"""
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(111)
    A = [np.zeros(3, np.float32) for _ in range(5)]
    B = [np.zeros(3, np.float32) for _ in range(5)]
    for i in range(3000):
        c = (rng.uniform(0, 1, 3) ** 4 * (12 if i % 3 else 1)).astype(np.float32)
        if i % 100 == 0: c[:] = (0.0031308, 0.04045, 0.0)                          # the knees
        if i % 7 == 0: c[1] = -c[1]                                                 # abs() inside the curves
        o.orc_hdr_curves(_v(*c), *[orc._p(x) for x in A]); r.hlslref_hdr_curves(_v(*c), *[orc._p(x) for x in B])
        for k in range(5):
            assert _same(A[k], B[k]), (i, k, c, A[k], B[k])


# ---- GaussianBlur.hlsl -----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(40, 24), (7, 5), (1, 1), (33, 2)])
def test_gaussian_blur_bit_exact(w, h):
    rng = np.random.default_rng(106)
    img = (rng.uniform(0, 1, (h, w, 4)) ** 2 * 6).astype(np.float32)
    for vertical in (False, True):
        ref = np.zeros_like(img)
        orc.hlsl_ref().hlslref_gaussian_blur(orc._p(img), orc._p(ref), C.c_int(w), C.c_int(h), C.c_int(int(vertical)))
        assert _same(orc.gaussian_blur(img, vertical), ref)


# ---- CubemapConvolution.hlsl ------------------------------------------------------------------------------------------------
def test_specular_prefilter_texels_bit_exact():
    env = small_env()
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(107)
    a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
    w, h, levels, pyr = env["hdri_w"], env["hdri_h"], env["levels"], orc._f(env["pyr"])
    for i in range(24):
        d = rng.normal(size=3).astype(np.float32)                    # PSMain normalises the interpolated direction
        rough = np.float32([0.0, 0.25, 0.5, 0.75, 1.0, 0.125][i % 6])
        # the shader's sample count is the constant 512 (CubemapConvolution.hlsl:178)
        o.orc_specular_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(rough), f32(w), f32(h), C.c_int(512), orc._p(a))
        r.hlslref_specular_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(rough), f32(w), f32(h), orc._p(b))
        assert _same(a, b), (i, d, rough, a, b)


def test_diffuse_irradiance_texels_bit_exact():
    """the shader's float-accumulated angle loops (step 0.010 -> 629 x 158 taps, source mip 3) against the oracle's"""
    env = small_env()
    o, r = orc.lib(), orc.hlsl_ref()
    rng = np.random.default_rng(108)
    a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
    w, h, levels, pyr = env["hdri_w"], env["hdri_h"], env["levels"], orc._f(env["pyr"])
    nphi, nth = C.c_int(0), C.c_int(0)
    assert o.orc_diffuse_angle_counts(f32(0.010), 0, 0, C.byref(nphi), C.byref(nth)) == 629 * 158
    for i in range(6):
        d = rng.normal(size=3).astype(np.float32)
        o.orc_diffuse_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(0.010), 0, 0, 3, orc._p(a))
        r.hlslref_diffuse_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), orc._p(b))
        assert _same(a, b), (i, d, a, b)


def test_brdf_lut_texels_bit_exact():
    r = orc.hlsl_ref()
    b = np.zeros(2, np.float32)
    for x, y in ((0, 0), (1023, 1023), (512, 17), (3, 900), (777, 333)):
        r.hlslref_brdf_lut_texel(C.c_int(x), C.c_int(y), orc._p(b))
        row = orc.brdf_integration_lut(1024, 1024, samples=2048, row_begin=y, row_end=y + 1, threads=1)
        assert _same(row[y, x], b), (x, y, row[y, x], b)


# ---- Skydome.hlsl / ApplyReflections.hlsl -----------------------------------------------------------------------------------
def test_skydome_psmain_bit_exact(vq):
    from vqengine_b200 import synth
    env = small_env()
    o, r = orc.lib(), orc.hlsl_ref()
    w, h = 48, 27
    inv = synth.sky_view_proj(0.7, -0.2, 1.0, w / h)[1].astype(np.float32)
    m = vq.Matrix()
    for k in range(16): m.m[k] = float(inv.reshape(-1)[k])
    scene = np.zeros((h, w, 4), np.float32)
    got = orc.skydome(env["pyr"], env["hdri_w"], env["hdri_h"], env["levels"], inv, scene)
    pyr = orc._f(env["pyr"])
    d, b = np.zeros(3, np.float32), np.zeros(4, np.float32)
    for y in range(h):
        for x in range(w):
            o.orc_skydome_look_direction(C.byref(m), x, y, w, h, orc._p(d))        # VSMain + rasteriser stand-in
            r.hlslref_skydome_pixel(orc._p(pyr), env["hdri_w"], env["hdri_h"], env["levels"], _v(*d), orc._p(b))
            assert _same(got[y, x], b), (x, y)


@pytest.mark.parametrize("bv", [False, True])
def test_apply_reflections_bit_exact(bv):
    rng = np.random.default_rng(109)
    w, h = 37, 19
    scene = (rng.uniform(0, 1, (h, w, 4)) * 5).astype(np.float32)
    refl = (rng.uniform(0, 1, (h, w, 4)) * 2).astype(np.float32)
    vol = rng.uniform(0, 1, (h, w, 4)).astype(np.float32) if bv else None
    want = orc.apply_reflections(scene.copy(), refl, vol)
    got = scene.copy()
    orc.hlsl_ref().hlslref_apply_reflections(orc._p(got), orc._p(refl), orc._p(vol) if bv else None, C.c_int(w), C.c_int(h))
    assert _same(want, got)


# ---- ForwardLighting.hlsl PSMain, whole pixel -------------------------------------------------------------------------------
def _pixel_scene(w, h, seed, casters, diffuse_only=False, uniform=False):
    from surface_util import material_set
    from vqengine_b200 import synth
    env = small_env()
    mats, texs, chains = material_set(4, 32, uniform=uniform)
    planes = synth.surface_inputs(w, h, 4, seed=seed, uv_scale=0.07)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=3, n_spot=2, casters=casters)
    pv.EnvironmentMapDiffuseOnlyIllumination = int(diffuse_only)
    return env, mats, chains, planes, pf, pv


def _oracle_pixels(env, mats, chains, planes, pf, pv, alpha_mask=False, **shadow):
    g = orc.gbuffer_from_materials(planes, mats, chains, pf.fAmbientLightingFactor, alpha_mask=alpha_mask, emissive=True,
                                   init=[np.full(planes[0].shape, np.nan, np.float32)] * 4 if alpha_mask else None)
    args = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    if shadow:
        return orc.forward_lighting_shadowed_e(pf, pv, g, *args, **shadow), g
    return orc.forward_lighting(pf, pv, g, *args), g


@pytest.mark.parametrize("diffuse_only", [False, True])
@pytest.mark.parametrize("seed", [5, 21])
def test_forward_psmain_unshadowed_bit_exact(diffuse_only, seed):
    """material sampling + Has*Map selection + normal mapping + SSAO + ambient/emissive + IBL + point/spot/directional lights:
    oracle (surface producer -> forward pass) == ForwardLighting.hlsl PSMain, every pixel, every bit"""
    env, mats, chains, planes, pf, pv = _pixel_scene(160, 72, seed, casters=False, diffuse_only=diffuse_only)
    pf.Lights.directional.shadowing = 0
    want, _ = _oracle_pixels(env, mats, chains, planes, pf, pv)
    got, disc = orc.hlsl_forward_image(pf, pv, planes, mats, chains, env)
    assert not disc.any()
    assert _same(want, got), np.argwhere(_bits(want) != _bits(got))[:5]


def test_forward_psmain_shadowed_bit_exact():
    """+ the caster lists and the shadowing directional light with their PCF tests (Lighting.hlsl:79-272)"""
    env, mats, chains, planes, pf, pv = _pixel_scene(128, 64, 6, casters=True)
    L = pf.Lights
    assert L.numPointCasters >= 1 and L.numSpotCasters >= 1 and L.directional.enabled
    L.directional.shadowing = 1
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0; m[1] = 0.01; m[4] = -0.02; m[12] = 0.1
    for sc in range(L.numSpotCasters):
        for k in range(16): L.shadowViews[sc].m[k] = float(m[k]) * (1.0 + 0.3 * sc)
    for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
    pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = 16.0
    pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = 16.0
    rng = np.random.default_rng(110)
    cubes = rng.uniform(0.0, 1.2, (L.numPointCasters, 6, 8, 8)).astype(np.float32)
    spots = rng.uniform(0.3, 0.7, (L.numSpotCasters, 16, 16)).astype(np.float32)
    dmap = rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32)
    shadow = dict(point_cubes=cubes, point_res=8, spot_maps=spots, dir_map=dmap)
    want, _ = _oracle_pixels(env, mats, chains, planes, pf, pv, **shadow)
    got, disc = orc.hlsl_forward_image(pf, pv, planes, mats, chains, env, **shadow)
    assert not disc.any()
    assert _same(want, got), np.argwhere(_bits(want) != _bits(got))[:5]
    # the shadow tests actually bite: differs from the unshadowed pass
    lit, _ = _oracle_pixels(env, mats, chains, planes, pf, pv)
    assert not _same(lit, want)


def test_forward_psmain_alpha_mask_permutation():
    """ENABLE_ALPHA_MASK=1: the same pixels are discarded, the survivors are identical"""
    env, mats, chains, planes, pf, pv = _pixel_scene(96, 48, 7, casters=False)
    pf.Lights.directional.shadowing = 0
    want, g = _oracle_pixels(env, mats, chains, planes, pf, pv, alpha_mask=True)
    got, disc = orc.hlsl_forward_image(pf, pv, planes, mats, chains, env, alpha_mask=True)
    dropped = np.isnan(g[0][..., 0])
    assert np.array_equal(dropped, disc)
    assert _same(want[~dropped], got[~dropped])


# ---- AMDFidelityFX.hlsl: CAS, FSR1 EASU, FSR1 RCAS (the GPU code path of ffx_a.h / ffx_cas.h / ffx_fsr1.h, A_GPU + A_HLSL) -----
def _post_image(w, h, seed, peak=1.0):
    rng = np.random.default_rng(seed)
    img = (rng.uniform(0, 1, (h, w, 4)) ** 2 * peak).astype(np.float32)
    img[h // 3: h // 3 + 2, :, :3] = 0.0                     # flat black rows: the zero-direction / zero-max branches
    img[:, w // 2, :3] = peak
    img[..., 3] = 1.0
    return img


@pytest.mark.parametrize("w,h", [(64, 32), (37, 21), (16, 16), (5, 3)])
@pytest.mark.parametrize("sharp", [0.0, 0.6, 1.0])
def test_cas_csmain_bit_exact(w, h, sharp):
    img = _post_image(w, h, 120 + w)
    con = orc.cas_setup(sharp, w, h, w, h)
    got = np.zeros_like(img); got[..., 3] = 1.0
    orc.hlsl_ref().hlslref_cas(con, orc._p(img), orc._p(got), C.c_int(w), C.c_int(h))
    assert _same(orc.cas(con, img), got)


@pytest.mark.parametrize("iw,ih,ow,oh", [(48, 27, 96, 54), (40, 30, 52, 39), (33, 17, 77, 41), (8, 8, 16, 16)])
def test_fsr_easu_csmain_bit_exact(iw, ih, ow, oh):
    """the shader's four Gather4 per channel (D3D footprint, CLAMP) against the oracle's twelve integer texel fetches"""
    img = _post_image(iw, ih, 130 + iw, peak=4.0)
    con = orc.fsr_easu_con(iw, ih, iw, ih, ow, oh)
    got = np.zeros((oh, ow, 4), np.float32); got[..., 3] = 1.0
    orc.hlsl_ref().hlslref_fsr_easu(con, orc._p(img), C.c_int(iw), C.c_int(ih), orc._p(got), C.c_int(ow), C.c_int(oh))
    assert _same(orc.fsr_easu(con, img, ow, oh, address_mode=1), got)


@pytest.mark.parametrize("w,h", [(64, 32), (37, 21), (5, 3)])
@pytest.mark.parametrize("stops", [0.0, 0.2, 2.0])
def test_fsr_rcas_csmain_bit_exact(w, h, stops):
    img = _post_image(w, h, 140 + w)
    con = orc.fsr_rcas_con(stops)
    got = np.zeros_like(img); got[..., 3] = 1.0
    orc.hlsl_ref().hlslref_fsr_rcas(con, orc._p(img), orc._p(got), C.c_int(w), C.c_int(h))
    assert _same(orc.fsr_rcas(con, img), got)


# ---- FidelityFX SPD (ffx_spd.h, LDS path) through AMDFidelityFX.hlsl's callbacks: one OS thread per lane, real barriers ------
def _mip_dims(w, h, n):
    return [(max(1, w >> l), max(1, h >> l)) for l in range(n)]


@pytest.mark.parametrize("w,h", [(128, 128), (256, 64), (64, 64), (192, 128), (512, 256), (100, 60), (70, 130), (33, 17)])
def test_spd_downsample_bit_exact(w, h):
    """every mip the 256-lane workgroups produce (LDS reductions, the last-workgroup tail from mip 6 on) == the oracle's
    level-by-level restatement, including its claim about which levels reduce column-major"""
    rng = np.random.default_rng(150 + w)
    src = (rng.uniform(0, 1, (h, w, 4)) ** 2 * 3).astype(np.float32)
    (dx, dy), (ox, oy), (nwg, mips) = orc.spd_setup((0, 0, w, h))
    want = orc.spd_downsample(src, mips)
    # the product (and the oracle) stop where a dimension would drop below 1: the D3D mip chain goes on with
    # max(1, .) sizes and SPD fills those levels from out-of-range (zero) texels; they are left unbound here (stores dropped)
    n_levels = len(want) + 1
    dims = _mip_dims(w, h, n_levels)
    levels = np.zeros((sum(a * b for a, b in dims), 4), np.float32)
    levels[: w * h] = src.reshape(-1, 4)
    orc.hlsl_ref().hlslref_spd_downsample(orc._p(levels), C.c_int(w), C.c_int(h), C.c_int(n_levels), C.c_uint32(mips), C.c_uint32(nwg),
                                          C.c_uint32(ox), C.c_uint32(oy), C.c_uint32(dx), C.c_uint32(dy))
    o = w * h
    assert 1 <= len(want) <= mips
    for l, lvl in enumerate(want, start=1):
        lw, lh = dims[l]
        got = levels[o: o + lw * lh].reshape(lh, lw, 4)
        o += lw * lh
        assert lvl.shape[:2] == (lh, lw)
        assert _same(lvl, got), (l, np.argwhere(_bits(lvl) != _bits(got))[:4])


@pytest.mark.parametrize("w,h", [(128, 128), (64, 64), (200, 120), (65, 33), (31, 70), (256, 16), (5, 3), (1, 1)])
def test_depth_min_pyramid_bit_exact(w, h):
    """DownsampleDepth.hlsl CSMain (level 0 copy + SPD with the MIN reduction, D3D mip-chain sizes, zero padding) == oracle"""
    rng = np.random.default_rng(160 + w)
    depth = rng.uniform(0.05, 1.0, (h, w)).astype(np.float32)
    want = orc.depth_min_pyramid(depth)
    dims = [lv.shape[::-1] for lv in want]
    assert dims[0] == (w, h) and dims[-1] == (1, 1) and len(want) == 1 + int(np.floor(np.log2(max(w, h))))
    levels = np.full(sum(a * b for a, b in dims), -1.0, np.float32)
    orc.hlsl_ref().hlslref_depth_pyramid(orc._p(depth), C.c_int(w), C.c_int(h), orc._p(levels), C.c_int(len(dims)))
    o = 0
    for l, lv in enumerate(want):
        lw, lh = dims[l]
        got = levels[o:o + lw * lh].reshape(lh, lw)
        o += lw * lh
        assert _same(lv, got), (l, lv, got)
    # closed form where no dimension has been clamped yet: plain 2x2 MIN of the level above
    for l in range(1, len(want)):
        if (w >> l) >= 1 and (h >> l) >= 1:
            up = want[l - 1]
            lh, lw = want[l].shape
            ref = np.minimum(np.minimum(up[0:2 * lh:2, 0:2 * lw:2], up[0:2 * lh:2, 1:2 * lw:2]),
                             np.minimum(up[1:2 * lh:2, 0:2 * lw:2], up[1:2 * lh:2, 1:2 * lw:2]))
            assert np.array_equal(ref, want[l])


def test_psmain_driven_from_a_gbuffer_equals_the_forward_oracle():
    """the K1 workload through the unmodified PSMain text (material constants = G-buffer texel): what bench.py --impl reference
    times. PSMain renormalises the interpolated normal, the G-buffer path takes it as is: equal up to that rounding."""
    from vqengine_b200 import synth
    env = small_env()
    w, h = 96, 40
    for emissive in (False, True):
        planes = synth.gbuffer(w, h, seed=8, emissive=emissive)
        pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=8, n_point=4, n_spot=2, casters=False)
        pf.Lights.directional.shadowing = 0
        args = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
        want = orc.forward_lighting(pf, pv, planes, *args)
        got = orc.hlsl_forward_gbuffer(pf, pv, planes, *args)
        assert np.allclose(got, want, rtol=2e-5, atol=2e-6), np.abs(got - want).max()
        assert (got[..., 3] == want[..., 3]).all()
        part = orc.hlsl_forward_gbuffer(pf, pv, planes, *args, row_begin=7, row_end=19)
        assert np.array_equal(part[7:19], got[7:19]) and not part[:7].any() and not part[19:].any()
