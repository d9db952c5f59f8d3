"""CPU (-m "not gpu"): pins the oracle.
The reference ships no golden vectors for this path and its HLSL cannot run here, so the pins are
(1) the reference's own FidelityFX A_CPU setup functions compiled in place (oracle/_ref), (2) the analytic known
answers of SURVEY.md §8(c) (tests/golden/kat.json), (3) the committed oracle fixtures (drift guard)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
f32 = C.c_float
V3 = f32 * 3


def hexes(a):
    return [f"{x:08x}" for x in a]


def test_ref_easu_con(orc):
    want = KAT["reference_a_cpu"]["FsrEasuCon_1920x1080_to_3840x2160"]
    assert hexes(orc.fsr_easu_con(1920, 1080, 1920, 1080, 3840, 2160))[:8] == want
    if orc.ref():
        assert hexes(orc.fsr_easu_con(1920, 1080, 1920, 1080, 3840, 2160, which="ref"))[:8] == want


def test_ref_rcas_con(orc):
    want = KAT["reference_a_cpu"]["FsrRcasCon_0p2"]
    assert hexes(orc.fsr_rcas_con(0.2))[:2] == want
    if orc.ref():
        assert hexes(orc.fsr_rcas_con(0.2, which="ref"))[:2] == want


def test_ref_cas_setup(orc):
    k = KAT["reference_a_cpu"]["CasSetup_0p8_3840x2160"]
    c = hexes(orc.cas_setup(0.8, 3840, 2160, 3840, 2160))
    assert c[0] == k["const0_0"] and c[2] == k["const0_2"] and c[4] == k["const1_0"] and c[5] == k["const1_1"]


def test_ref_spd_setup(orc):
    k = KAT["reference_a_cpu"]["SpdSetup_0_0_3840_2160"]
    d, o, n = orc.spd_setup((0, 0, 3840, 2160))
    assert d == k["dispatch"] and n == [k["numWorkGroups"], k["mips"]] and o == [0, 0]


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libffxref.so")), reason="oracle/_ref not built")
def test_setup_functions_random_sweep_vs_reference_build(orc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        iw, ih = int(rng.integers(16, 4096)), int(rng.integers(16, 2304))
        s = float(rng.uniform(1.0, 4.0))
        ow, oh = int(iw * s), int(ih * s)
        assert list(orc.fsr_easu_con(iw, ih, iw + 7, ih + 3, ow, oh)) == list(orc.fsr_easu_con(iw, ih, iw + 7, ih + 3, ow, oh, which="ref"))
        sh = float(rng.uniform(-0.5, 1.5))
        assert list(orc.cas_setup(sh, iw, ih, ow, oh)) == list(orc.cas_setup(sh, iw, ih, ow, oh, which="ref"))
        st = float(rng.uniform(0, 8))
        assert list(orc.fsr_rcas_con(st)) == list(orc.fsr_rcas_con(st, which="ref"))
        rect = (int(rng.integers(0, 512)), int(rng.integers(0, 512)), iw, ih)
        assert orc.spd_setup(rect) == orc.spd_setup(rect, which="ref")
        assert orc.spd_setup(rect, 5) == orc.spd_setup(rect, 5, which="ref")


def test_single_pixel_brdf_kat(orc):
    """BASELINE config 1: 1x1 pixel, 1 point light, scalar BRDF on the CPU."""
    k = KAT["analytic"]["single_pixel_point_light"]
    L = orc.lib()
    assert abs(L.orc_ndf_ggx(f32(1.0), f32(0.5)) - k["D"]) < 1e-5
    assert abs(L.orc_geometry_smith(V3(0, 0, 1), V3(0, 0, 1), V3(0, 0, 1), f32(0.5)) - k["G"]) < 1e-6
    out = V3()
    L.orc_fresnel_schlick(V3(0, 0, 1), V3(0, 0, 1), V3(.04, .04, .04), out)
    assert abs(out[0] - k["F"]) < 1e-7
    L.orc_brdf(V3(0, 0, 1), V3(0, 0, 1), V3(0, 0, 1), V3(.5, .5, .5), f32(.5), f32(0), out)
    assert abs(out[0] - k["brdf"]) < 1e-6
    import vqengine_b200 as vq
    l = vq.PointLight()
    l.position.z = 2.0; l.range = 100.0; l.brightness = 10.0
    l.color.x = l.color.y = l.color.z = 1.0
    L.orc_point_light(C.byref(l), V3(0, 0, 0), V3(0, 0, 1), V3(0, 0, 1), V3(.5, .5, .5), f32(.5), f32(0), out)
    assert all(abs(out[i] - k["contribution"]) < 2e-7 for i in range(3))
    l.range = 1.5                          # D < range fails -> no contribution (Lighting.hlsl:318)
    L.orc_point_light(C.byref(l), V3(0, 0, 0), V3(0, 0, 1), V3(0, 0, 1), V3(.5, .5, .5), f32(.5), f32(0), out)
    assert list(out) == [0, 0, 0]


def test_random_brdf_tuples_are_finite_and_reciprocal(orc):
    """64 random (N,V,L,albedo,rough,metal) tuples: finite, non-negative, Helmholtz-reciprocal in the
    specular lobe up to the F(H,V) asymmetry being symmetric for H = normalize(V+L)."""
    rng = np.random.default_rng(1)
    L = orc.lib()
    for _ in range(64):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        def hemi():
            v = rng.normal(size=3); v /= np.linalg.norm(v)
            return v if v @ n > 0.05 else hemi()
        v, l = hemi(), hemi()
        alb = rng.uniform(0.02, 0.9, 3); r = float(rng.uniform(0.04, 1)); m = float(rng.uniform(0, 1))
        a, b = V3(), V3()
        L.orc_brdf(V3(*n), V3(*v), V3(*l), V3(*alb), f32(r), f32(m), a)
        L.orc_brdf(V3(*n), V3(*l), V3(*v), V3(*alb), f32(r), f32(m), b)
        assert np.isfinite(list(a)).all() and min(a) >= 0
        assert np.allclose(list(a), list(b), rtol=2e-4, atol=1e-6)


def test_tonemapper_kat(orc):
    import vqengine_b200 as vq
    o = (f32 * 4)()
    orc.lib().orc_tonemap_pixel(C.byref(vq.TonemapperParams(0, 0, 200.0, 1, 1.0)), (f32 * 4)(1, 1, 1, 0.25), o)
    assert abs(o[0] - KAT["analytic"]["tonemap_srgb_of_1"]) < 1e-7 and o[3] == 0.25
    orc.lib().orc_tonemap_pixel(C.byref(vq.TonemapperParams(0, 9, 200.0, 1, 1.0)), (f32 * 4)(1, 1, 1, 1), o)
    assert list(o)[:3] == [1.0, 1.0, 0.0]            # unknown curve paints yellow (Tonemapper.hlsl:143-145)
    orc.lib().orc_tonemap_pixel(C.byref(vq.TonemapperParams(0, 2, 200.0, 1, 1.0)), (f32 * 4)(3, 2, 1, 1), o)
    assert list(o)[:3] == [3.0, 2.0, 1.0]            # linear passthrough
    orc.lib().orc_tonemap_pixel(C.byref(vq.TonemapperParams(1, 1, 10000.0, 0, 1.0)), (f32 * 4)(1, 1, 1, 1), o)
    assert abs(o[0] - 1.0) < 1e-6                    # PQ(1.0 of 10000 nits) = 1


def test_diffuse_loop_trip_counts(orc):
    a, b = C.c_int(), C.c_int()
    for key, step in [("0.01", 0.01), ("0.025", 0.025), ("0.05", 0.05), ("pi_over_32", float(np.float32(np.pi / 32)))]:
        n = orc.lib().orc_diffuse_angle_counts(f32(step), 0, 0, C.byref(a), C.byref(b))
        assert [a.value, b.value] == KAT["analytic"]["diffuse_trip_counts"][key] and n == a.value * b.value
    assert orc.lib().orc_diffuse_angle_counts(f32(0.0), 64, 16, C.byref(a), C.byref(b)) == 1024   # BASELINE config 2


def test_hammersley(orc):
    h = (f32 * 2)()
    for i, want in enumerate(KAT["analytic"]["hammersley_512_first5"]):
        orc.lib().orc_hammersley(i, 512, h)
        assert list(h) == want


def test_mip_level_counts(orc):
    for k, v in KAT["analytic"]["mip_level_count"].items():
        w, h = map(int, k.split("x"))
        assert orc.lib().orc_mip_level_count(w, h) == v


def test_aprx_bit_hacks(orc):
    """APrx* are integer operations on the float's bits (ffx_a.h:1842-1845)."""
    L = orc.lib()
    for x in [0.001, 0.3, 1.0, 2.5, 17.0]:
        bits = np.float32(x).view(np.uint32)
        assert np.float32(L.orc_aprx(0, f32(x))).view(np.uint32) == np.uint32((int(bits) >> 1) + 0x1fbc4639)
        assert np.float32(L.orc_aprx(1, f32(x))).view(np.uint32) == np.uint32(0x7ef07ebb - int(bits))
        assert np.float32(L.orc_aprx(3, f32(x))).view(np.uint32) == np.uint32(0x5f347d74 - (int(bits) >> 1))
        assert abs(L.orc_aprx(2, f32(x)) * x - 1.0) < 4e-3           # medium-precision reciprocal
        assert abs(L.orc_aprx(0, f32(x)) / np.sqrt(x) - 1.0) < 0.07  # low-precision sqrt


def test_golden_fixture_matches_oracle(orc):
    """drift guard: the committed oracle outputs are reproduced bit-for-bit by the current oracle build"""
    from vqengine_b200 import synth, TonemapperParams
    g = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
    img = g["post_in"]
    assert np.array_equal(orc.gaussian_blur(img, False), g["blur_x"])
    assert np.array_equal(orc.gaussian_blur(g["blur_x"], True), g["blur_xy"])
    # libm (powf/exp2f/...) may differ in the last bit between glibc builds: allow 2 ulp there
    assert np.allclose(orc.tonemap(synth.default_tonemapper(), img), g["tonemap_srgb"], rtol=3e-7, atol=1e-7)
    assert np.allclose(orc.tonemap(TonemapperParams(0, 1, 200.0, 0, 1.0), img), g["tonemap_pq"], rtol=3e-6, atol=1e-7)
    ldr = g["tonemap_srgb"]
    assert np.array_equal(orc.cas(orc.cas_setup(0.8, 40, 24, 40, 24), ldr), g["cas_08"])
    assert np.array_equal(orc.fsr_easu(orc.fsr_easu_con(40, 24, 40, 24, 80, 48), ldr, 80, 48, 0), g["easu_2x_wrap"])
    assert np.array_equal(orc.fsr_easu(orc.fsr_easu_con(40, 24, 40, 24, 80, 48), ldr, 80, 48, 1), g["easu_2x_clamp"])
    assert np.array_equal(orc.fsr_rcas(orc.fsr_rcas_con(0.2), ldr), g["rcas_02"])
    for i, m in enumerate(orc.spd_downsample(g["spd_in"], 5), start=1):
        assert np.array_equal(m, g[f"spd_l{i}"])
    assert np.array_equal(orc.hdri_build_mips(g["hdri"], 6), g["hdri_pyr"])
    assert np.allclose(orc.diffuse_irradiance(g["hdri_pyr"], 64, 32, 6, 8, n_phi=16, n_theta=8, src_mip=1), g["diffuse_8"], rtol=1e-5, atol=1e-6)
    assert np.allclose(orc.specular_prefilter(g["hdri_pyr"], 64, 32, 6, 8, 3, num_samples=64), g["spec_8x3"], rtol=1e-5, atol=1e-6)
    assert np.allclose(orc.brdf_integration_lut(16, 16, samples=256), g["lut_16"], rtol=1e-5, atol=1e-7)
