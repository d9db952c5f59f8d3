"""Deterministic cases shared by tests/golden/make_shader_golden.py (runs them through the REFERENCE's shader text compiled as C++,
oracle/_ref/libhlslref.so, and commits the outputs) and tests/test_shader_golden.py (runs them through the ORACLE and requires the
committed outputs bit for bit). Each case: name -> (reference_fn, oracle_fn), both returning one float32 array.
Inputs that are not produced by this repository's own seeded generators (vqengine_b200/synth.py) are stored in the golden file too."""
import ctypes as C

import numpy as np

import oracle_lib as orc
from envmaps import small_env

f32 = C.c_float


def _v(*x):
    return (f32 * len(x))(*[float(a) for a in x])


def _unit(rng):
    v = rng.normal(size=3).astype(np.float32)
    return v / np.float32(np.linalg.norm(v))


def _pixel_scene(w, h, seed, casters):
    from surface_util import material_set
    from vqengine_b200 import synth
    env = small_env()
    mats, texs, chains = material_set(4, 32)
    planes = synth.surface_inputs(w, h, 4, seed=seed, uv_scale=0.07)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=seed, n_point=3, n_spot=2, casters=casters)
    return env, mats, chains, planes, pf, pv


def _shadow_setup(pf, seed):
    L = pf.Lights
    L.directional.shadowing = 1
    m = np.zeros(16, np.float32); m[0] = 1 / 25; m[5] = 1 / 25; m[14] = 0.5; m[15] = 1.0; m[1] = 0.01; m[4] = -0.02; m[12] = 0.1
    for sc in range(L.numSpotCasters):
        for k in range(16): L.shadowViews[sc].m[k] = float(m[k]) * (1.0 + 0.3 * sc)
    for k in range(16): L.shadowViewDirectional.m[k] = float(m[k])
    pf.f2SpotLightShadowMapDimensions.x = pf.f2SpotLightShadowMapDimensions.y = 16.0
    pf.f2DirectionalLightShadowMapDimensions.x = pf.f2DirectionalLightShadowMapDimensions.y = 16.0
    rng = np.random.default_rng(seed)
    return dict(point_cubes=rng.uniform(0.0, 1.2, (L.numPointCasters, 6, 8, 8)).astype(np.float32), point_res=8,
                spot_maps=rng.uniform(0.3, 0.7, (L.numSpotCasters, 16, 16)).astype(np.float32),
                dir_map=rng.uniform(0.3, 0.7, (16, 16)).astype(np.float32))


def _oracle_pixels(env, mats, chains, planes, pf, pv, alpha_mask=False, **shadow):
    g = orc.gbuffer_from_materials(planes, mats, chains, pf.fAmbientLightingFactor, alpha_mask=alpha_mask, emissive=True,
                                   init=[np.full(planes[0].shape, np.nan, np.float32)] * 4 if alpha_mask else None)
    args = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    out = orc.forward_lighting_shadowed_e(pf, pv, g, *args, **shadow) if shadow else orc.forward_lighting(pf, pv, g, *args)
    if alpha_mask:
        out[np.isnan(g[0][..., 0])] = np.nan                         # discarded pixels
    return out


def _ref_pixels(env, mats, chains, planes, pf, pv, alpha_mask=False, **shadow):
    out, disc = orc.hlsl_forward_image(pf, pv, planes, mats, chains, env, alpha_mask=alpha_mask, **shadow)
    if alpha_mask:
        out[disc] = np.nan
    return out


def cases():
    import vqengine_b200 as vq
    o = orc.lib()
    r = orc.hlsl_ref          # called lazily: None when oracle/_ref is not built
    out = {}

    # ---- whole PSMain -------------------------------------------------------------------------------------------------
    def psmain(kind):
        def build():
            w, h, seed = (96, 48, 7) if kind == "alpha_mask" else (48, 24, 5)         # the larger scene has pixels the mask discards
            env, mats, chains, planes, pf, pv = _pixel_scene(w, h, seed, casters=(kind == "shadowed"))
            kw = {}
            if kind == "shadowed":
                kw = _shadow_setup(pf, 110)
            else:
                pf.Lights.directional.shadowing = 0
            if kind == "diffuse_only":
                pv.EnvironmentMapDiffuseOnlyIllumination = 1
            return (env, mats, chains, planes, pf, pv), dict(alpha_mask=(kind == "alpha_mask"), **kw)
        return (lambda: (lambda a, k: _ref_pixels(*a, **k))(*build())), (lambda: (lambda a, k: _oracle_pixels(*a, **k))(*build()))
    for kind in ("unshadowed", "diffuse_only", "shadowed", "alpha_mask"):
        out["psmain_" + kind] = psmain(kind)

    # ---- BRDF / light scalars -----------------------------------------------------------------------------------------
    def brdf(fn_name):
        def run(lib, prefix):
            rng = np.random.default_rng(201)
            res = np.zeros((300, 3), np.float32)
            for i in range(300):
                N, V, Wi = _unit(rng) * np.float32(rng.uniform(0.5, 1.5)), _unit(rng) * np.float32(rng.uniform(0.5, 1.5)), _unit(rng)
                alb = rng.uniform(0, 1, 3).astype(np.float32)
                rough, metal = np.float32(rng.uniform(0, 1)), np.float32(rng.uniform(0, 1))
                getattr(lib, prefix + "brdf")(_v(*N), _v(*V), _v(*Wi), _v(*alb), f32(rough), f32(metal), orc._p(res[i]))
            return res
        return (lambda: run(r(), "hlslref_")), (lambda: run(o, "orc_"))
    out["brdf"] = brdf("brdf")

    def integrate():
        def run(lib, prefix):
            res = np.zeros((5, 2), np.float32)
            for i, (ndv, rough, n) in enumerate(((0.5, 0.5, 64), (0.031, 0.97, 128), (0.999, 0.015, 96), (1.0, 0.0, 16), (0.25, 1.0, 2048))):
                getattr(lib, prefix + "integrate_brdf")(f32(ndv), f32(rough), C.c_int(n), orc._p(res[i]))
            return res
        return (lambda: run(r(), "hlslref_")), (lambda: run(o, "orc_"))
    out["integrate_brdf"] = integrate()

    # ---- tonemapper + curves ------------------------------------------------------------------------------------------
    def tonemap():
        def run(lib, prefix):
            rng = np.random.default_rng(202)
            res = []
            for curve in (0, 1, 2, 3):
                for space in (0, 1):
                    for gamma in (0, 1):
                        p = vq.TonemapperParams()
                        p.ContentColorSpace, p.OutputDisplayCurve, p.ToggleGammaCorrection = space, curve, gamma
                        p.DisplayReferenceBrightnessLevel = 200.0
                        for i in range(24):
                            px = (rng.uniform(0, 1, 4) ** 3 * 20).astype(np.float32)
                            b = np.zeros(4, np.float32)
                            getattr(lib, prefix + "tonemap_pixel")(C.byref(p), _v(*px), orc._p(b))
                            res.append(b)
            return np.stack(res)
        return (lambda: run(r(), "hlslref_")), (lambda: run(o, "orc_"))
    out["tonemap"] = tonemap()

    # ---- image passes -------------------------------------------------------------------------------------------------
    def img(w, h, seed, peak=1.0):
        rng = np.random.default_rng(seed)
        a = (rng.uniform(0, 1, (h, w, 4)) ** 2 * peak).astype(np.float32)
        a[..., 3] = 1.0
        return a

    def blur(vertical):
        src = img(24, 16, 203, 6.0)
        def ref():
            b = np.zeros_like(src); r().hlslref_gaussian_blur(orc._p(src), orc._p(b), C.c_int(24), C.c_int(16), C.c_int(int(vertical))); return b
        return ref, (lambda: orc.gaussian_blur(src, vertical))
    out["blur_x"], out["blur_y"] = blur(False), blur(True)

    def cas():
        src = img(37, 21, 204)
        con = orc.cas_setup(0.6, 37, 21, 37, 21)
        def ref():
            b = np.zeros_like(src); b[..., 3] = 1.0; r().hlslref_cas(con, orc._p(src), orc._p(b), C.c_int(37), C.c_int(21)); return b
        return ref, (lambda: orc.cas(con, src))
    out["cas"] = cas()

    def easu():
        src = img(33, 17, 205, 4.0)
        con = orc.fsr_easu_con(33, 17, 33, 17, 77, 41)
        def ref():
            b = np.zeros((41, 77, 4), np.float32); b[..., 3] = 1.0
            r().hlslref_fsr_easu(con, orc._p(src), C.c_int(33), C.c_int(17), orc._p(b), C.c_int(77), C.c_int(41)); return b
        return ref, (lambda: orc.fsr_easu(con, src, 77, 41, address_mode=1))
    out["easu"] = easu()

    def rcas():
        src = img(37, 21, 206)
        con = orc.fsr_rcas_con(0.2)
        def ref():
            b = np.zeros_like(src); b[..., 3] = 1.0; r().hlslref_fsr_rcas(con, orc._p(src), orc._p(b), C.c_int(37), C.c_int(21)); return b
        return ref, (lambda: orc.fsr_rcas(con, src))
    out["rcas"] = rcas()

    def spd():
        w, h = 64, 32
        src = img(w, h, 207, 3.0)
        (dx, dy), (ox, oy), (nwg, mips) = orc.spd_setup((0, 0, w, h))
        def orc_levels():
            return np.concatenate([l.reshape(-1, 4) for l in orc.spd_downsample(src, mips)])
        def ref():
            n = len(orc.spd_downsample(src, mips)) + 1
            dims = [(max(1, w >> l), max(1, h >> l)) for l in range(n)]
            lv = np.zeros((sum(a * b for a, b in dims), 4), np.float32); lv[: w * h] = src.reshape(-1, 4)
            r().hlslref_spd_downsample(orc._p(lv), C.c_int(w), C.c_int(h), C.c_int(n), C.c_uint32(mips), C.c_uint32(nwg), C.c_uint32(ox),
                                       C.c_uint32(oy), C.c_uint32(dx), C.c_uint32(dy))
            return lv[w * h:]
        return ref, orc_levels
    out["spd"] = spd()

    def depth():
        w, h = 65, 33
        d = np.random.default_rng(208).uniform(0.05, 1.0, (h, w)).astype(np.float32)
        def ref():
            want = orc.depth_min_pyramid(d)
            lv = np.full(sum(x.size for x in want), -1.0, np.float32)
            r().hlslref_depth_pyramid(orc._p(d), C.c_int(w), C.c_int(h), orc._p(lv), C.c_int(len(want)))
            return lv
        return ref, (lambda: np.concatenate([x.reshape(-1) for x in orc.depth_min_pyramid(d)]))
    out["depth_pyramid"] = depth()

    # ---- IBL integrals ------------------------------------------------------------------------------------------------
    def ibl(kind):
        def run(ref):
            env = small_env()
            w, h, levels, pyr = env["hdri_w"], env["hdri_h"], env["levels"], orc._f(env["pyr"])
            rng = np.random.default_rng(209)
            n = 12 if kind == "specular" else 2
            res = np.zeros((n, 4), np.float32)
            for i in range(n):
                d = rng.normal(size=3).astype(np.float32)
                if kind == "specular":
                    rough = np.float32([0.0, 0.25, 0.5, 0.75, 1.0, 0.125][i % 6])
                    if ref:
                        r().hlslref_specular_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(rough), f32(w), f32(h), orc._p(res[i]))
                    else:
                        o.orc_specular_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(rough), f32(w), f32(h), C.c_int(512), orc._p(res[i]))
                else:
                    if ref:
                        r().hlslref_diffuse_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), orc._p(res[i]))
                    else:
                        o.orc_diffuse_irradiance_texel(orc._p(pyr), w, h, levels, _v(*d), f32(0.010), 0, 0, 3, orc._p(res[i]))
            return res
        return (lambda: run(True)), (lambda: run(False))
    out["specular_prefilter_texels"], out["diffuse_irradiance_texels"] = ibl("specular"), ibl("diffuse")

    def lut():
        pts = ((0, 0), (1023, 1023), (512, 17), (3, 900))
        def ref():
            res = np.zeros((len(pts), 2), np.float32)
            for i, (x, y) in enumerate(pts):
                r().hlslref_brdf_lut_texel(C.c_int(x), C.c_int(y), orc._p(res[i]))
            return res
        def orc_():
            return np.stack([orc.brdf_integration_lut(1024, 1024, samples=2048, row_begin=y, row_end=y + 1, threads=1)[y, x] for x, y in pts])
        return ref, orc_
    out["brdf_lut_texels"] = lut()
    return out
