"""CPU (-m "not gpu"): size-independent properties of the oracle's passes and of the sampling semantics it fixes
(SURVEY.md §8(c), §9): these hold for the reference's math by construction and pin the restatement's structure."""
import ctypes as C

import numpy as np
import pytest

f32 = C.c_float
V3 = f32 * 3


def test_gaussian_blur_constant_and_linearity(orc):
    rng = np.random.default_rng(0)
    c = np.full((20, 30, 4), 0.75, np.float32)
    out = orc.gaussian_blur(orc.gaussian_blur(c, False), True)
    assert np.abs(out[..., :3] - 0.75).max() < 1e-6 and (out[..., 3] == 1).all()
    a, b = rng.uniform(0, 4, (16, 40, 4)).astype(np.float32), rng.uniform(0, 4, (16, 40, 4)).astype(np.float32)
    lhs = orc.gaussian_blur(a + b, False)[..., :3]
    rhs = (orc.gaussian_blur(a, False) + orc.gaussian_blur(b, False))[..., :3]
    assert np.abs(lhs - rhs).max() < 5e-6
    # clamp-to-edge: a 1-pixel-wide image is its own blur
    col = rng.uniform(0, 1, (9, 1, 4)).astype(np.float32)
    assert np.allclose(orc.gaussian_blur(col, False)[..., :3], col[..., :3], atol=1e-6)


def test_spd_is_plain_2x2_average(orc):
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 8, (48, 80, 4)).astype(np.float32)
    lv = orc.spd_downsample(img, 4)
    cur = img.astype(np.float64)
    for l in lv:
        h, w = cur.shape[0] // 2, cur.shape[1] // 2
        cur = cur[: 2 * h, : 2 * w].reshape(h, 2, w, 2, 4).mean(axis=(1, 3))
        assert l.shape == (h, w, 4) and np.abs(l - cur).max() < 1e-5
    # checksum of checksums: the mean is preserved down the chain for power-of-two sizes
    p2 = rng.uniform(0, 1, (64, 64, 4)).astype(np.float32)
    assert abs(orc.spd_downsample(p2, 6)[-1].mean() - p2.mean()) < 1e-5


def test_min_pyramid(orc):
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 16, (32, 64, 4)).astype(np.float32)
    pyr = orc.hdri_build_mips(img, 6)
    l1 = pyr[64 * 32: 64 * 32 + 32 * 16].reshape(16, 32, 4)
    want = img[:, :, :3].reshape(16, 2, 32, 2, 3).min(axis=(1, 3))
    assert np.array_equal(l1[..., :3], want) and (l1[..., 3] == 1).all()
    assert pyr[-2:, :3].min() >= img[..., :3].min() - 0    # last level 2x1 holds minima of minima
    assert np.isclose(pyr[-2:, :3].min(), img[..., :3].min())


def test_specular_prefilter_constant_radiance(orc):
    L = np.array([3.0, 0.25, 1.5, 1.0], np.float32)
    pyr = orc.hdri_build_mips(np.tile(L, (32, 64, 1)), 6)
    out = orc.specular_prefilter(pyr, 64, 32, 6, 8, 3, num_samples=128)
    assert np.allclose(out[:, :3], L[:3], rtol=3e-6) and (out[:, 3] == 1).all()


def test_diffuse_irradiance_constant_radiance_reference_step(orc):
    L = np.array([2.0, 1.0, 0.5, 1.0], np.float32)
    pyr = orc.hdri_build_mips(np.tile(L, (16, 32, 1)), 5)
    out = orc.diffuse_irradiance(pyr, 32, 16, 5, 2, step=0.01, src_mip=3)
    assert np.allclose(out[:, 0] / 2.0, 0.99415, atol=2e-4) and (out[:, 3] == 1).all()


def test_brdf_lut_corner_and_range(orc):
    lut = orc.brdf_integration_lut(32, 32, samples=512)
    assert np.isfinite(lut).all() and lut.min() >= 0 and lut.max() <= 1.05
    assert abs(lut[0, 31, 0] - 1.0) < 0.05 and lut[0, 31, 1] < 0.05       # roughness->0, NdotV->1: (~1, ~0)


def test_cube_direction_roundtrip(orc):
    L = orc.lib()
    d = V3(); face = C.c_int(); sx = f32(); sy = f32()
    for res in (2, 8, 64):
        for f in range(6):
            for (px, py) in [(0, 0), (res - 1, 0), (res // 2, res - 1), (res - 1, res - 1)]:
                L.orc_cube_texel_direction(f, px, py, res, d)
                L.orc_direction_to_cube_face(d, C.byref(face), C.byref(sx), C.byref(sy))
                assert face.value == f
                assert abs((sx.value * 0.5 + 0.5) * res - 0.5 - px) < 1e-4 and abs((0.5 - sy.value * 0.5) * res - 0.5 - py) < 1e-4
    # D3D face order / axes (CubemapUtility.h:31-41)
    for f, axis in enumerate([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]):
        L.orc_direction_to_cube_face(V3(*axis), C.byref(face), C.byref(sx), C.byref(sy))
        assert face.value == f and sx.value == 0 and sy.value == 0


def test_cube_edge_tap_resolution_is_symmetric(orc):
    """stepping off face A lands on the adjacent edge texel of face B, and stepping back off B lands on A's edge texel"""
    L = orc.lib()
    out = (C.c_int * 3)(); back = (C.c_int * 3)()
    for N in (2, 4, 16, 64, 512):
        for f in range(6):
            for t in range(N):
                for (i, j) in [(-1, t), (N, t), (t, -1), (t, N)]:
                    L.orc_cube_resolve_edge_tap(N, f, i, j, out)
                    f2, i2, j2 = list(out)
                    assert f2 != f and 0 <= i2 < N and 0 <= j2 < N
                    assert i2 in (0, N - 1) or j2 in (0, N - 1)          # an edge texel of the neighbour
                    # step from that texel outward across the same edge: must return to (f, clamp(i), clamp(j))
                    cands = [(-1, j2), (N, j2), (i2, -1), (i2, N)]
                    hits = []
                    for (a, b) in cands:
                        if (a in (-1, N)) != (b in (-1, N)):
                            L.orc_cube_resolve_edge_tap(N, f2, a, b, back)
                            hits.append(tuple(back))
                    assert (f, min(max(i, 0), N - 1), min(max(j, 0), N - 1)) in hits


def test_cube_sampling_is_continuous_across_edges_and_corners(orc):
    rng = np.random.default_rng(3)
    N = 8
    cube = rng.uniform(0, 4, (6 * N * N, 4)).astype(np.float32)
    L = orc.lib()
    a, b = (f32 * 4)(), (f32 * 4)()
    p = cube.ctypes.data_as(C.POINTER(f32))
    eps = 1e-4
    for _ in range(300):
        # a point on a cube edge (two coordinates of magnitude 1) or corner, nudged to either side
        d = rng.uniform(-1, 1, 3)
        k = rng.permutation(3)
        d[k[0]] = rng.choice([-1.0, 1.0]); d[k[1]] = rng.choice([-1.0, 1.0])
        if rng.uniform() < 0.2:
            d[k[2]] = rng.choice([-1.0, 1.0])
        d1, d2 = d.copy(), d.copy()
        d1[k[0]] *= (1 + eps); d2[k[1]] *= (1 + eps)
        L.orc_sample_cube(p, N, 1, V3(*d1), 0, a)
        L.orc_sample_cube(p, N, 1, V3(*d2), 0, b)
        assert np.abs(np.array(a) - np.array(b)).max() < 4 * 4 * N * eps, (d, list(a), list(b))


def test_equirect_sampling_wraps_and_interpolates(orc):
    rng = np.random.default_rng(4)
    w, h = 16, 8
    img = rng.uniform(0, 1, (h, w, 4)).astype(np.float32)
    pyr = orc.hdri_build_mips(img, 4)
    L = orc.lib(); o = (f32 * 4)(); o2 = (f32 * 4)()
    p = pyr.ctypes.data_as(C.POINTER(f32))
    L.orc_sample_equirect(p, w, h, 4, f32((3 + 0.5) / w), f32((2 + 0.5) / h), f32(0), o)       # texel centre -> texel
    assert np.allclose(list(o), img[2, 3], atol=1e-6)
    L.orc_sample_equirect(p, w, h, 4, f32(0.0), f32((2 + 0.5) / h), f32(0), o)                 # u = 0: wraps to last column
    assert np.allclose(list(o), 0.5 * (img[2, 0] + img[2, w - 1]), atol=1e-6)
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(0), o)
    L.orc_sample_equirect(p, w, h, 4, f32(1.3), f32(-0.6), f32(0), o2)                          # periodic in u and v
    assert np.allclose(list(o), list(o2), atol=2e-5)
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(0.5), o2)                         # trilinear midpoint
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(1.0), o)
    m1 = np.array(o)
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(0.0), o)
    assert np.allclose(list(o2), 0.5 * (np.array(o) + m1), atol=1e-6)
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(99.0), o)                         # lod clamps to the last level
    L.orc_sample_equirect(p, w, h, 4, f32(0.3), f32(0.4), f32(3.0), o2)
    assert list(o) == list(o2)


def test_easu_identity_and_flat(orc):
    """flat image stays flat; 1:1 'upscale' of an image reproduces it within the ringing clamp"""
    flat = np.full((12, 16, 4), 0.4, np.float32)
    out = orc.fsr_easu(orc.fsr_easu_con(16, 12, 16, 12, 32, 24), flat, 32, 24, 0)
    assert np.abs(out[..., :3] - 0.4).max() < 1e-6 and (out[..., 3] == 1).all()
    rng = np.random.default_rng(5)
    img = rng.uniform(0, 1, (12, 16, 4)).astype(np.float32)
    same = orc.fsr_easu(orc.fsr_easu_con(16, 12, 16, 12, 16, 12), img, 16, 12, 1)
    # at scale 1 the resolve position is the texel centre: the result lies inside the 2x2 min/max clamp around it
    assert np.isfinite(same).all() and same[..., :3].min() >= img[..., :3].min() - 1e-6 and same[..., :3].max() <= img[..., :3].max() + 1e-6


def test_cas_and_rcas_flat_and_bounds(orc):
    flat = np.full((9, 11, 4), 0.3, np.float32)
    c = orc.cas(orc.cas_setup(0.8, 11, 9, 11, 9), flat)
    assert np.abs(c[2:-2, 2:-2, :3] - 0.3).max() < 2e-3          # APrxMedRcp is a ~1e-3 approximation
    r = orc.fsr_rcas(orc.fsr_rcas_con(0.2), flat)
    assert np.abs(r[2:-2, 2:-2, :3] - 0.3).max() < 2e-3
    rng = np.random.default_rng(6)
    img = rng.uniform(0, 1, (20, 24, 4)).astype(np.float32)
    c = orc.cas(orc.cas_setup(1.0, 24, 20, 24, 20), img)
    assert c.min() >= 0 and c.max() <= 1                          # CAS saturates its output


def test_forward_row_ranges_and_thread_invariance(orc):
    from vqengine_b200 import synth
    from envmaps import small_env
    env = small_env()
    planes = synth.gbuffer(20, 12, seed=2)
    pf, pv = synth.scene_constants(20, 12, env["spec_mips"], seed=2, n_spot=1)
    a = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    full = orc.forward_lighting(pf, pv, planes, *a, threads=1)
    assert np.array_equal(full, orc.forward_lighting(pf, pv, planes, *a, threads=5))
    part = orc.forward_lighting(pf, pv, planes, *a, 3, 9, threads=2)
    assert np.array_equal(part[3:9], full[3:9]) and (part[:3] == 0).all() and (part[9:] == 0).all()
    assert np.isfinite(full).all() and (full[..., 3] == planes[1][..., 3]).all()
