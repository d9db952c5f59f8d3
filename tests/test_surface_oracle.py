"""CPU: the §8(f).1 oracle (oracle/oracle_surface.cpp) against independent restatements and known answers.
The reference holds no golden vectors for this path: parity unpinned by the reference itself; the pins
here are (1) an independent numpy restatement of DXGIUtils.cpp:263-287 and (2) analytic known answers. (The restatement of
PSMain's surface half is additionally pinned against the reference's shader text compiled as C++: tests/test_hlsl_ref.py.)"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as orc
import vqengine_b200 as vq
from vqengine_b200 import synth
from surface_util import material_set


def np_box_mip(src):
    """(a+b+c+d)/4 truncating over complete 2x2 blocks, DXGIUtils.cpp:263-287"""
    h, w = src.shape[:2]
    s = src[: (h // 2) * 2, : (w // 2) * 2].astype(np.uint32)
    return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2]) // 4).astype(np.uint8)


@pytest.mark.parametrize("w,h", [(64, 64), (48, 20), (37, 19), (1, 8), (130, 3)])
def test_box_mip_chain_matches_numpy(w, h):
    rng = np.random.default_rng(w * 1000 + h)
    lvl0 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    levels = vq.mip_level_count(w, h)
    chain = orc.texture_mip_chain(lvl0, levels)
    cur = lvl0
    for l in range(levels):
        o = vq.pyramid_offset(w, h, l) * 4
        lw, lh = w >> l, h >> l
        assert cur.shape[:2] == (lh, lw)
        assert np.array_equal(chain[o:o + lw * lh * 4].reshape(lh, lw, 4), cur), f"level {l}"
        cur = np_box_mip(cur)


def test_box_mip_truncates():
    lvl0 = np.array([[[0, 1, 2, 255]] * 2, [[0, 1, 3, 255], [3, 0, 0, 254]]], np.uint8)     # sums 3, 3, 7, 1019
    chain = orc.texture_mip_chain(lvl0, 2)
    assert list(chain[16:20]) == [0, 0, 1, 254]


def test_lod_known_answers():
    w = h = 64
    levels = vq.mip_level_count(w, h)
    chain = orc.texture_mip_chain(np.full((h, w, 4), 200, np.uint8), levels)
    for k, want in ((1.0, 0.0), (2.0, 1.0), (8.0, 3.0), (0.25, 0.0), (1e6, levels - 1.0)):
        _, lod = orc.sample_texture8(chain, w, h, levels, 0.3, 0.7, ddx=(k / w, 0.0), ddy=(0.0, 0.5 / h))
        assert abs(lod - want) < 1e-6, (k, lod)
    # the longer of the two footprint axes decides; non-square sizes scale u and v separately; bias adds
    _, lod = orc.sample_texture8(chain, w, h, levels, 0, 0, ddx=(1 / w, 0), ddy=(3 / w, 4 / h))
    assert abs(lod - np.log2(5.0)) < 1e-6
    _, lod = orc.sample_texture8(chain, w, h, levels, 0, 0, ddx=(2 / w, 0), ddy=(0, 0), bias=0.75)
    assert abs(lod - 1.75) < 1e-6
    # zero derivatives (flat quad) -> lod 0, constant texture -> the constant at any lod
    val, lod = orc.sample_texture8(chain, w, h, levels, 5.3, -2.2)
    assert lod == 0.0 and np.allclose(val, 200 / 255.0, atol=1e-7)
    # null SRV reads 0
    val, _ = orc.sample_texture8(None, 0, 0, 0, 0.5, 0.5)
    assert (val == 0).all()


def test_sampling_known_answers():
    w, h = 8, 4
    lvl0 = np.zeros((h, w, 4), np.uint8)
    lvl0[..., 0] = (np.arange(w)[None, :] * 30)            # ramp in x
    lvl0[..., 1] = ((np.arange(w)[None, :] + np.arange(h)[:, None]) % 2) * 255   # checker
    levels = vq.mip_level_count(w, h)
    chain = orc.texture_mip_chain(lvl0, levels)
    # texel centre -> the texel
    val, _ = orc.sample_texture8(chain, w, h, levels, (3 + 0.5) / w, (2 + 0.5) / h)
    assert np.allclose(val[:2], [90 / 255.0, 255 / 255.0], atol=1e-7)
    # half-way between texel 3 and 4 in x
    val, _ = orc.sample_texture8(chain, w, h, levels, 4.0 / w, 2.5 / h)
    assert abs(val[0] - 105 / 255.0) < 1e-6 and abs(val[1] - 0.5) < 1e-6
    # WRAP: u = 0 blends texel w-1 and texel 0; whole-number uv offsets change nothing
    val, _ = orc.sample_texture8(chain, w, h, levels, 0.0, 0.5 / h)
    assert abs(val[0] - (210 + 0) / 2 / 255.0) < 1e-6
    a, _ = orc.sample_texture8(chain, w, h, levels, 0.3, 0.6)
    b, _ = orc.sample_texture8(chain, w, h, levels, 0.3 + 2.0, 0.6 - 1.0)
    assert np.allclose(a, b, atol=2e-6)
    # lod 1: the checker averages to 127 (truncated), trilinear at lod 0.5 is the mean of both levels
    val, lod = orc.sample_texture8(chain, w, h, levels, 0.5, 0.5, ddx=(2.0 / w, 0), ddy=(0, 0))
    assert lod == 1.0 and abs(val[1] - 127 / 255.0) < 1e-6
    v0, _ = orc.sample_texture8(chain, w, h, levels, 0.3, 0.3)
    v1, _ = orc.sample_texture8(chain, w, h, levels, 0.3, 0.3, ddx=(2.0 / w, 0))
    vh, lod = orc.sample_texture8(chain, w, h, levels, 0.3, 0.3, ddx=(np.sqrt(2.0) / w, 0))
    assert abs(lod - 0.5) < 1e-6 and np.allclose(vh, 0.5 * (v0 + v1), atol=1e-6)


def test_unpack_normal_known_answers():
    n = np.array([0.0, 1.0, 0.0], np.float32)
    t = np.array([1.0, 0.0, 0.0], np.float32)
    # flat tangent-space normal (0.5,0.5,1) -> the surface normal
    assert np.allclose(orc.unpack_normal([0.5, 0.5, 1.0], n, t), n, atol=1e-6)
    # +x in tangent space -> T; +y -> B = normalize(cross(T, N))
    assert np.allclose(orc.unpack_normal([1.0, 0.5, 0.5], n, t), t, atol=1e-6)
    assert np.allclose(orc.unpack_normal([0.5, 1.0, 0.5], n, t), np.cross(t, n), atol=1e-6)
    # a tangent that is not orthogonal to N is Gram-Schmidt-ed first (ShadingMath.hlsl:47)
    t2 = np.array([0.8, 0.6, 0.0], np.float32)
    assert np.allclose(orc.unpack_normal([1.0, 0.5, 0.5], n, t2), [1, 0, 0], atol=1e-6)


def test_constant_material_known_answers():
    """all SRVs null -> G-buffer = material constants; N = normalize(interpolated normal); ao = ambient * ssao(x+1,y+1)"""
    w, h = 40, 22
    mats, texs, chains = material_set(3, 16)
    planes = synth.surface_inputs(w, h, 3)
    planes[2][..., 3] = 2.0                                 # material 2 = constants only
    out = orc.gbuffer_from_materials(planes, mats, chains, 0.25)
    m = mats[2]
    assert np.array_equal(out[0][..., :3], planes[0][..., :3])
    ss = np.roll(planes[3][..., 0], (-1, -1), axis=(0, 1))
    assert np.allclose(out[0][..., 3], np.float32(0.25) * ss, rtol=0, atol=1e-7)
    nn = planes[1][..., :3] / np.sqrt((planes[1][..., :3] ** 2).sum(-1, keepdims=True))
    assert np.allclose(out[1][..., :3], nn, atol=2e-7)
    assert np.allclose(out[1][..., 3], m.roughness) and np.allclose(out[2][..., 3], m.metalness)
    assert np.allclose(out[2][..., :3], [m.diffuse.x, m.diffuse.y, m.diffuse.z])
    assert np.allclose(out[3], [m.emissiveColor.x, m.emissiveColor.y, m.emissiveColor.z, m.emissiveIntensity])


def test_textured_material_properties():
    w, h = 96, 54
    mats, texs, chains = material_set(4, 64)
    planes = synth.surface_inputs(w, h, 4, uv_scale=0.02)
    out = orc.gbuffer_from_materials(planes, mats, chains, 0.3)
    assert all(np.isfinite(o).all() for o in out)
    ln = np.sqrt((out[1][..., :3] ** 2).sum(-1))
    assert np.abs(ln - 1.0).max() < 1e-5                    # TBN is orthonormal, sampled normal is normalised
    mid = planes[2][..., 3].astype(int)
    for i, m in enumerate(mats):                            # sRGB->linear darkens: textured albedo <= constant tint
        sel = mid == i
        if int(m.textureConfig) & vq.TEXCFG_DIFFUSE:
            assert (out[2][sel][:, :3] <= np.array([m.diffuse.x, m.diffuse.y, m.diffuse.z]) + 1e-6).all()
        if int(m.textureConfig) & (vq.TEXCFG_ROUGHNESS | vq.TEXCFG_ORM):
            assert (out[1][sel][:, 3] <= m.roughness + 1e-6).all()
    # ENABLE_ALPHA_MASK: discarded pixels keep the previous G-buffer contents, the others are unchanged
    init = [np.full((h, w, 4), -7.0, np.float32) for _ in range(4)]
    masked = orc.gbuffer_from_materials(planes, mats, chains, 0.3, alpha_mask=True, init=init)
    gone = (masked[2] == -7.0).all(-1)
    assert 0.02 < gone.mean() < 0.6
    for a, b in zip(masked, out):
        assert np.array_equal(a[~gone], b[~gone]) and (a[gone] == -7.0).all()
    assert not (gone & (mid == 2)).any()                    # no diffuse map -> never discarded
    # row tiling is exact (derivative quads are aligned to absolute rows, also for an odd split)
    top = orc.gbuffer_from_materials(planes, mats, chains, 0.3, row_begin=0, row_end=17)
    bot = orc.gbuffer_from_materials(planes, mats, chains, 0.3, row_begin=17, row_end=h)
    for t, b, o in zip(top, bot, out):
        assert np.array_equal(np.concatenate([t[:17], b[17:]]), o)


def test_uv_derivatives_follow_the_quad():
    """mip selection uses the 2x2-quad differences of uv: doubling the uv pitch raises the lod by exactly 1"""
    w, h = 32, 16
    mats, texs, chains = material_set(2, 64)
    a = synth.surface_inputs(w, h, 2, uv_scale=0.02, ssao=False)
    for p in a:
        p[..., :3] = p[0:1, 0:1, :3]                        # constant geometry, only uv varies
    a[2][..., 3] = 1.0                                      # glTF-style material (albedo + normal + ORM)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    tex_w = 64
    m = mats[1]
    a[0][..., 3] = xs * (1.0 / tex_w / m.uvScaleOffset.x)   # exactly one texel per pixel in x, none in y -> lod 0
    a[1][..., 3] = 0.0
    b = [p.copy() for p in a]
    b[0][..., 3] *= 2.0                                     # two texels per pixel -> lod 1
    oa = orc.gbuffer_from_materials(a, mats, chains, 1.0)
    ob = orc.gbuffer_from_materials(b, mats, chains, 1.0)
    ch, tw, th, tl = chains[1]["occl_rough_metal"]
    u, v = float(a[0][3, 5, 3]) * m.uvScaleOffset.x + m.uvScaleOffset.z, m.uvScaleOffset.w
    s0, lod0 = orc.sample_texture8(ch, tw, th, tl, u, v, ddx=(1.0 / tex_w, 0.0))
    assert lod0 == 0.0 and abs(oa[1][3, 5, 3] - m.roughness * s0[1]) < 1e-6
    u = float(b[0][3, 5, 3]) * m.uvScaleOffset.x + m.uvScaleOffset.z
    s1, lod1 = orc.sample_texture8(ch, tw, th, tl, u, v, ddx=(2.0 / tex_w, 0.0))
    assert lod1 == 1.0 and abs(ob[1][3, 5, 3] - m.roughness * s1[1]) < 1e-6
