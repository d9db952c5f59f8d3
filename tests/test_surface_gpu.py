"""-m gpu parity: SURVEY §8(f).1 — RGBA8 mip chain (bit-exact) and the surface producer (G-buffer from material textures,
|delta| <= 1e-4) against the scalar oracle, through the C-ABI."""
import numpy as np
import pytest
import torch

from gpu_util import dev, host, assert_abs, assert_scaled, TOL
from envmaps import small_env
from surface_util import material_set

pytestmark = pytest.mark.gpu


def _upload_materials(ctx, vq, mats, chains):
    keep, tex = [], []
    for d in chains:
        mt = vq.MaterialTextures()
        for slot in vq.MATERIAL_TEXTURE_SLOTS:
            e = d.get(slot)
            if e is not None:
                buf, w, h, levels = e
                t = torch.from_numpy(buf).cuda()
                keep.append(t)
                setattr(mt, slot, vq.texture_of(t, w, h, levels))
        tex.append(mt)
    return ctx.material_table(mats, tex), keep


def _run(ctx, vq, orc, w, h, *, n=4, tex_res=64, uv_scale=0.02, ssao=True, emissive=True, alpha_mask=False, rows=None,
         seed=12, ambient=0.3, uniform=False):
    from vqengine_b200 import synth
    mats, texs, chains = material_set(n, tex_res, uniform=uniform)
    planes = synth.surface_inputs(w, h, n, seed=seed, ssao=ssao, uv_scale=uv_scale)
    table, keep = _upload_materials(ctx, vq, mats, chains)
    dp = [dev(p) for p in planes]
    si = vq.SurfaceInputs(vq.image_of(dp[0]), vq.image_of(dp[1]), vq.image_of(dp[2]),
                          vq.image_of(dp[3], 1) if ssao else vq.null_image())
    init = [np.full((h, w, 4), -7.0, np.float32) for _ in range(4 if emissive else 3)]
    outs = [dev(i) for i in init]
    gb = vq.GBuffer(vq.image_of(outs[0]), vq.image_of(outs[1]), vq.image_of(outs[2]),
                    vq.image_of(outs[3]) if emissive else vq.null_image())
    rb, re = rows if rows else (0, h)
    ctx.gbuffer_from_materials(si, table, ambient, gb, alpha_mask=alpha_mask, row_begin=rb, row_end=re)
    got = [host(o) for o in outs]
    table.close()
    ref = orc.gbuffer_from_materials(planes, mats, chains, ambient, alpha_mask=alpha_mask, emissive=emissive, init=init,
                                     row_begin=rb, row_end=re)
    return got, ref, planes, mats


NAMES = ("position_ao", "normal_roughness", "albedo_metalness", "emissive")


def _check(tag, got, ref):
    for name, g, r in zip(NAMES, got, ref):
        assert np.array_equal(g == -7.0, r == -7.0), f"{tag}/{name}: different pixels written"
        print(assert_abs(f"{tag}/{name}", g, r))
    assert np.array_equal(got[0][..., :3], ref[0][..., :3])          # P is a pass-through


@pytest.mark.parametrize("w,h", [(64, 64), (48, 20), (37, 19), (1, 8), (130, 3), (1024, 512)])
def test_texture_mip_chain_bit_exact(ctx, vq, orc, w, h):
    rng = np.random.default_rng(w * 7 + h)
    lvl0 = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    levels = vq.mip_level_count(w, h)
    ref = orc.texture_mip_chain(lvl0, levels)
    buf = np.zeros_like(ref)
    buf[: w * h * 4] = lvl0.reshape(-1)
    t = torch.from_numpy(buf).cuda()
    ctx.texture_build_mips(vq.texture_of(t, w, h, levels))
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy(), ref)


@pytest.mark.parametrize("w,h,uv_scale", [(96, 54, 0.02), (96, 54, 0.1), (96, 54, 0.004), (251, 141, 0.03), (33, 7, 0.5), (1, 1, 0.1),
                                          (257, 9, 0.05), (640, 5, 0.02), (128, 2, 0.1)])
def test_surface_parity(ctx, vq, orc, w, h, uv_scale):
    """every texture configuration (separate maps / ORM / constants only / tiled non-pow2), magnified to heavily minified"""
    got, ref, _, _ = _run(ctx, vq, orc, w, h, uv_scale=uv_scale)
    _check(f"surface{w}x{h}@{uv_scale}", got, ref)


@pytest.mark.parametrize("uv_scale", [0.02, 0.2])
def test_surface_parity_uniform_map_sizes(ctx, vq, orc, uv_scale):
    """all maps of a material the same size: consecutive Sample() calls share one sampler state in the kernel"""
    got, ref, _, _ = _run(ctx, vq, orc, 192, 108, tex_res=128, uv_scale=uv_scale, uniform=True)
    _check(f"surface_uniform@{uv_scale}", got, ref)


def _record_material_set(vq, w=100, h=60, seed=77):
    """two materials whose maps all share one (non-power-of-two) size, so both take the texel-record path: every texture flag at
    once (separate roughness/metalness/AO maps AND an ORM map), uv tiling, a biased normal map on one, a null emissive SRV with
    the flag set on the other"""
    from vqengine_b200 import synth
    from vqengine_b200.shader_data import (TEXCFG_DIFFUSE, TEXCFG_NORMAL, TEXCFG_AO, TEXCFG_ROUGHNESS, TEXCFG_METALLIC,
                                           TEXCFG_EMISSIVE, TEXCFG_ORM)
    base, _ = synth.materials(2, 16)
    kinds = {"diffuse": "albedo", "normals": "normal", "emissive": "emissive", "metalness": "scalar", "roughness": "scalar",
             "occl_rough_metal": "orm", "local_ao": "scalar"}
    mats, chains = [], []
    for i, m in enumerate(base):
        m.textureConfig = float(TEXCFG_DIFFUSE | TEXCFG_NORMAL | TEXCFG_AO | TEXCFG_ROUGHNESS | TEXCFG_METALLIC | TEXCFG_EMISSIVE | TEXCFG_ORM)
        m.emissiveIntensity = 1.5
        m.normalMapMipBias = 0.75 if i == 0 else 0.0
        m.uvScaleOffset.x, m.uvScaleOffset.y, m.uvScaleOffset.z, m.uvScaleOffset.w = (2.5, 1.5, 0.2, -0.3) if i == 0 else (1.0, 1.0, 0.0, 0.0)
        levels = vq.mip_level_count(w, h)
        d = {}
        for k, (slot, kind) in enumerate(kinds.items()):
            if i == 1 and slot == "emissive":
                d[slot] = None                                        # flag set, SRV null: reads 0
                continue
            lvl0 = synth.material_texture(kind, w, h, seed + 10 * i + k)
            import oracle_lib
            d[slot] = (oracle_lib.texture_mip_chain(lvl0, levels), w, h, levels)
        mats.append(m); chains.append(d)
    return mats, chains


def _run_custom(ctx, vq, mats, chains, planes, w, h, alpha_mask=False):
    table, keep = _upload_materials(ctx, vq, mats, chains)
    dp = [dev(p) for p in planes]
    si = vq.SurfaceInputs(vq.image_of(dp[0]), vq.image_of(dp[1]), vq.image_of(dp[2]), vq.image_of(dp[3], 1))
    init = [np.full((h, w, 4), -7.0, np.float32) for _ in range(4)]
    outs = [dev(i) for i in init]
    ctx.gbuffer_from_materials(si, table, 0.3, vq.GBuffer(*(vq.image_of(o) for o in outs)), alpha_mask=alpha_mask)
    got = [host(o) for o in outs]
    table.close()
    return got, init


@pytest.mark.parametrize("uv_scale,alpha_mask", [(0.02, False), (0.3, False), (0.004, True)])
def test_surface_record_path_every_flag(ctx, vq, orc, uv_scale, alpha_mask):
    """texel records (one 16-byte record per texel holding all seven maps): non-pow2 maps, all flags, biased normal, null SRV"""
    from vqengine_b200 import synth
    w, h = 160, 90
    mats, chains = _record_material_set(vq)
    planes = synth.surface_inputs(w, h, 2, uv_scale=uv_scale)
    got, init = _run_custom(ctx, vq, mats, chains, planes, w, h, alpha_mask)
    ref = orc.gbuffer_from_materials(planes, mats, chains, 0.3, alpha_mask=alpha_mask, init=init)
    _check(f"surface_records@{uv_scale}", got, ref)


@pytest.mark.parametrize("uniform", [True, False])
def test_surface_record_path_equals_map_by_map(ctx, vq, orc, uniform, monkeypatch):
    """the record path performs the map-by-map path's operations in the same order: bit-identical G-buffers"""
    from vqengine_b200 import synth
    w, h = 192, 108
    if uniform:
        mats, texs, chains = material_set(4, 128, uniform=True)
        planes = synth.surface_inputs(w, h, 4, uv_scale=0.05)
    else:
        mats, chains = _record_material_set(vq)
        planes = synth.surface_inputs(w, h, 2, uv_scale=0.05)
    a, _ = _run_custom(ctx, vq, mats, chains, planes, w, h)
    monkeypatch.setenv("VQ_SURFACE_RECORDS", "0")
    b, _ = _run_custom(ctx, vq, mats, chains, planes, w, h)
    for name, x, y in zip(NAMES, a, b):
        assert np.array_equal(x, y), f"{name}: records differ from map-by-map sampling"


def test_surface_without_optional_planes(ctx, vq, orc):
    got, ref, _, _ = _run(ctx, vq, orc, 96, 54, ssao=False, emissive=False)
    _check("surface_no_ssao_no_emissive", got, ref)


def test_surface_alpha_mask_discard(ctx, vq, orc):
    got, ref, _, _ = _run(ctx, vq, orc, 160, 90, uv_scale=0.01, alpha_mask=True)
    gone = (ref[2] == -7.0).all(-1)
    assert 0.02 < gone.mean() < 0.7
    _check("surface_alpha_mask", got, ref)


@pytest.mark.parametrize("rows", [(0, 17), (17, 54), (5, 6), (53, 54), (20, 20)])
def test_surface_row_ranges(ctx, vq, orc, rows):
    """row tiles (multi-GPU sharding): quads stay aligned to absolute even rows, rows outside the range are untouched"""
    got, ref, _, _ = _run(ctx, vq, orc, 96, 54, rows=rows)
    _check(f"surface_rows{rows}", got, ref)
    for g in got:
        assert (g[: rows[0]] == -7.0).all() and (g[rows[1]:] == -7.0).all()


def test_surface_feeds_forward_lighting(ctx, vq, orc):
    """producer -> K1: the G-buffer layout is the contract between them. Stage-wise strict: the GPU G-buffer is within 1e-4 of
    the oracle's, and K1 on the GPU G-buffer is within 1e-4 (scaled) of the oracle's K1 on the SAME G-buffer. The end-to-end
    difference (oracle chain vs GPU chain) is reported: K1 amplifies 1e-5-level G-buffer differences through the GGX lobe."""
    from vqengine_b200 import synth
    w, h = 96, 54
    env = small_env()
    mats, texs, chains = material_set(4, 64)
    planes = synth.surface_inputs(w, h, 4, uv_scale=0.02)
    pf, pv = synth.scene_constants(w, h, env["spec_mips"], seed=3)
    pf.fAmbientLightingFactor = 0.07
    table, keep = _upload_materials(ctx, vq, mats, chains)
    dp = [dev(p) for p in planes]
    si = vq.SurfaceInputs(vq.image_of(dp[0]), vq.image_of(dp[1]), vq.image_of(dp[2]), vq.image_of(dp[3], 1))
    g = [torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
    gb = vq.GBuffer(*(vq.image_of(t) for t in g))
    ctx.gbuffer_from_materials(si, table, pf.fAmbientLightingFactor, gb)
    dd, ds, dl = dev(env["diff"]), dev(env["spec"]), dev(env["lut"])
    em = vq.EnvironmentMaps(vq.cubemap_of(dd, env["diff_res"], 1), vq.cubemap_of(ds, env["spec_res"], env["spec_mips"]),
                            vq.image_of(dl, 2))
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, em, out)
    got = host(out)
    gpu_g = [host(t) for t in g]
    table.close()
    rg = orc.gbuffer_from_materials(planes, mats, chains, pf.fAmbientLightingFactor)
    for name, a, b in zip(NAMES, gpu_g, rg):
        assert_abs(f"chain/{name}", a, b)
    envargs = (env["diff"], env["diff_res"], env["spec"], env["spec_res"], env["spec_mips"], env["lut"])
    print(assert_scaled("forward(gpu gbuffer)", got, orc.forward_lighting(pf, pv, gpu_g, *envargs)))
    from gpu_util import report
    r = report("surface->forward end to end", got, orc.forward_lighting(pf, pv, rg, *envargs))
    print(r)
    assert r["max_scaled"] <= 2e-3 and r["frac_abs_le_tol"] > 0.99


def test_surface_full_size_properties(ctx, vq, orc):
    """3840x2160 (BASELINE full size): tiling invariance, finiteness, unit normals, and parity on sampled rows"""
    from vqengine_b200 import synth
    w, h = 3840, 2160
    mats, texs, chains = material_set(4, 256)
    planes = synth.surface_inputs(w, h, 4)
    table, keep = _upload_materials(ctx, vq, mats, chains)
    dp = [dev(p) for p in planes]
    si = vq.SurfaceInputs(vq.image_of(dp[0]), vq.image_of(dp[1]), vq.image_of(dp[2]), vq.image_of(dp[3], 1))
    full = [torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
    tiled = [torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
    ctx.gbuffer_from_materials(si, table, 0.3, vq.GBuffer(*(vq.image_of(t) for t in full)))
    gbt = vq.GBuffer(*(vq.image_of(t) for t in tiled))
    for rb, re in ((0, 701), (701, 1400), (1400, 2160)):
        ctx.gbuffer_from_materials(si, table, 0.3, gbt, row_begin=rb, row_end=re)
    torch.cuda.synchronize()
    for a, b in zip(full, tiled):
        assert torch.equal(a, b)
    assert all(bool(torch.isfinite(t).all()) for t in full)
    ln = full[1][..., :3].norm(dim=-1)
    assert float((ln - 1).abs().max()) < 1e-5
    rows = (0, 1, 777, 1079, 1080, 2158, 2159)
    for y in rows:
        ref = orc.gbuffer_from_materials(planes, mats, chains, 0.3, row_begin=y, row_end=y + 1)
        for name, g, r in zip(NAMES, full, ref):
            assert_abs(f"surface4k/row{y}/{name}", g[y].cpu().numpy(), r[y])
    table.close()


def test_surface_argument_errors(ctx, vq):
    from vqengine_b200 import synth
    mats, texs, chains = material_set(2, 16)
    bad = vq.MaterialTextures()
    bad.diffuse = vq.Texture2D(1234, 16, 16, 9)                       # more levels than a 16x16 texture has
    with pytest.raises(vq.VqError):
        ctx.material_table(mats[:1], [bad])
    table, keep = _upload_materials(ctx, vq, mats, chains)
    a = torch.zeros((8, 8, 4), device="cuda"); b = torch.zeros((8, 9, 4), device="cuda")
    si = vq.SurfaceInputs(vq.image_of(a), vq.image_of(a), vq.image_of(b), vq.null_image())
    gb = vq.GBuffer(vq.image_of(a), vq.image_of(a), vq.image_of(a), vq.null_image())
    with pytest.raises(vq.VqError):
        ctx.gbuffer_from_materials(si, table, 1.0, gb)                # planes differ in size
    si = vq.SurfaceInputs(vq.image_of(a), vq.image_of(a), vq.image_of(a), vq.null_image())
    with pytest.raises(vq.VqError):
        ctx.gbuffer_from_materials(si, table, 1.0, gb, row_begin=3, row_end=9)
    table.close()
