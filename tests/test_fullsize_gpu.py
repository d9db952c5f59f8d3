"""-m gpu parity at the FULL sizes of BASELINE.json's configs 2-5 ("full image <= C4; C5: one full
face x mip plus 1 % random texels"), through the C-ABI, against the scalar oracle on the box's host cores.

  C2  2048x1024 HDRI -> 6 x 64^2 diffuse irradiance, 64 x 16 sample grid (1024 samples), whole cube, strict 1e-4
  C3  1920x1080 forward PBR, 4 point + 1 directional + IBL (64^2 / 512^2 x 9 / 1024^2 LUT), whole frame
  C4  3840x2160 -> SPD (11 mips, bit-exact) -> blur X,Y -> tonemap -> CAS -> EASU 2x (7680x4320) -> RCAS, whole images,
      each stage against the oracle on the stage's own input, strict 1e-4
  C5  4096x2048 HDRI -> 512^2 x 6 x 9 mips, 512 samples: one full face of one mip + 1 % of all texels chosen at random;
      and a 7680x4320 forward frame: two row tiles of the 8-rank partition (540 rows each) against the oracle

HDR-valued outputs (forward radiance, prefiltered texels up to the HDRI peak) are held to 1e-4 * max(1, |ref|) AND,
explicitly, to the strict absolute 1e-4 wherever |ref| <= 1 (gpu_util.assert_scaled)."""
import numpy as np
import pytest
import torch

from gpu_util import dev, host, assert_abs, assert_scaled, report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def envk(ctx, vq):
    import bench
    k = bench.build_env_maps_gpu(ctx, vq, torch)
    k["np"] = {n: host(k[n]) for n in ("diff", "spec", "lut")}
    yield k
    ctx.environment_invalidate()


def test_c2_diffuse_irradiance_full_cube(ctx, vq, orc):
    from vqengine_b200 import synth
    w, h, res = 2048, 1024, 64
    levels = vq.mip_level_count(w, h)
    pyr_t = torch.zeros((vq.pyramid_texel_count(w, h, levels), 4), dtype=torch.float32, device="cuda")
    img = synth.hdri(w, h, seed=synth.SEED_BASE + 2)
    pyr_t[: w * h] = dev(img).reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, w, h, levels)
    ctx.hdri_build_mips(pyr)
    ref_pyr = orc.hdri_build_mips(img, levels)
    assert np.array_equal(host(pyr_t), ref_pyr)                     # K11 at the config size: bit-exact
    out = torch.zeros((6 * res * res, 4), dtype=torch.float32, device="cuda")
    ctx.diffuse_irradiance(pyr, vq.cubemap_of(out, res, 1), n_phi=64, n_theta=16, src_mip=3)
    ref = orc.diffuse_irradiance(ref_pyr, w, h, levels, res, n_phi=64, n_theta=16, src_mip=3)
    print(assert_abs("C2 diffuse 6x64^2", host(out), ref))


def test_c3_forward_1080p_whole_frame(ctx, vq, orc, envk):
    from vqengine_b200 import synth
    w, h = 1920, 1080
    planes = synth.gbuffer(w, h, seed=synth.SEED_BASE + 3)
    pf, pv = synth.scene_constants(w, h, envk["spec_mips"])
    dpl = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ctx.forward_lighting(pf, pv, gb, envk["env"], out)
    e = envk["np"]
    ref = orc.forward_lighting(pf, pv, planes, e["diff"], envk["diff_res"], e["spec"], envk["spec_res"], envk["spec_mips"], e["lut"])
    got = host(out)
    r = assert_scaled("C3 forward 1920x1080", got, ref)
    assert (got[..., 3] == ref[..., 3]).all()
    assert r["frac_abs_le_tol"] >= 0.9999, r
    print(r)


def test_c4_post_chain_full_size(ctx, vq, orc):
    from vqengine_b200 import synth
    w, h = 3840, 2160
    img = synth.hdr_image(w, h, seed=synth.SEED_BASE + 4)
    d = dev(img)
    # SPD: 11 mips in one launch, bit-exact
    (dx, dy), c = vq.spd_setup(w, h)
    mips = min(c.mips, int(np.floor(np.log2(min(w, h)))))
    c.mips = mips
    dsts = [torch.zeros((h >> l, w >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, mips + 1)]
    ctx.spd_downsample(c, d, dsts)
    for l, (g, r) in enumerate(zip(dsts, orc.spd_downsample(img, mips)), start=1):
        assert np.array_equal(host(g), r), f"C4 SPD level {l}"
    del dsts
    a, b = torch.empty_like(d), torch.empty_like(d)
    tm = synth.default_tonemapper()
    ctx.gaussian_blur(d, a, False)
    ha = host(a)
    assert_abs("C4 blur_x", ha, orc.gaussian_blur(img, False))
    ctx.gaussian_blur(a, b, True)
    hb = host(b)
    assert_abs("C4 blur_y", hb, orc.gaussian_blur(ha, True))
    del ha
    ctx.tonemap(tm, b, a)
    ht = host(a)
    assert_abs("C4 tonemap", ht, orc.tonemap(tm, hb))
    del hb
    ctx.cas(vq.cas_setup(0.8, w, h, w, h), a, b)
    hc = host(b)
    assert_abs("C4 cas", hc, orc.cas(orc.cas_setup(0.8, w, h, w, h), ht))
    del ht
    e = torch.empty((2 * h, 2 * w, 4), dtype=torch.float32, device="cuda")
    ctx.fsr_easu(vq.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), b, e)
    he = host(e)
    assert_abs("C4 easu 7680x4320", he, orc.fsr_easu(orc.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), hc, 2 * w, 2 * h, 0))
    del hc
    r = torch.empty_like(e)
    ctx.fsr_rcas(vq.fsr_rcas_con(0.2), e, r)
    assert_abs("C4 rcas 7680x4320", host(r), orc.fsr_rcas(orc.fsr_rcas_con(0.2), he))


def test_c5_specular_full_face_and_random_texels(ctx, vq, orc):
    from vqengine_b200 import synth
    hw, hh, res, mips = 4096, 2048, 512, 9
    levels = vq.mip_level_count(hw, hh)
    pyr_t = torch.zeros((vq.pyramid_texel_count(hw, hh, levels), 4), dtype=torch.float32, device="cuda")
    pyr_t[: hw * hh] = dev(synth.hdri(hw, hh, seed=synth.SEED_BASE + 5)).reshape(-1, 4)
    pyr = vq.pyramid_of(pyr_t, hw, hh, levels)
    ctx.hdri_build_mips(pyr)
    n = vq.cubemap_texel_count(res, mips)
    cube = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    ctx.specular_prefilter(pyr, vq.cubemap_of(cube, res, mips), 512)
    got = host(cube)
    hp = host(pyr_t)
    assert np.isfinite(got).all() and (got[:, 3] == 1.0).all()
    def check(name, ids):
        """|delta| <= 1e-4 * max(1,|ref|) + 2 * S, S = how far the ORACLE's own texel moves when (a) its look direction is tilted by
        4.8e-7 rad (2^-21: a few ulps) or (b) the length of its un-normalised sample vectors is off by 1-4 ulps
        (oracle_capi.cpp: orc_specular_prefilter_sensitivity). S is ~1e-6 except where an importance sample lands within a few
        texels of a POLE of the equirect map: the reference feeds reflect(-V, H) — unit length only up to rounding — into
        v = asin(-L.y)/pi + 0.5, whose slope is infinite at the poles, and its WRAP-in-v sampler then blends in the OPPOSITE pole's
        row: one ulp of |L| moves the blend weight by ~0.15 and the texel by 4e-4 (measured: profiles/r02_diag_fullsize.txt, up
        to 2e-3 on the 5 texels around each pole of mips 1-2). Two correct fp32 evaluations cannot agree better than the
        reference agrees with itself; everywhere else the bound is the strict one, and independently of S at least 99.5 % of the
        texels of every set must meet the strict bound (measured: 99.93-99.99 %)."""
        ref = orc.specular_prefilter_texels(hp, hw, hh, levels, res, mips, ids)
        sens = orc.specular_prefilter_sensitivity(hp, hw, hh, levels, res, mips, ids)
        g = got[ids]
        assert np.isfinite(g).all()
        d = np.abs(g.astype(np.float64) - ref).max(axis=1)
        scale = np.maximum(1.0, np.abs(ref).max(axis=1))
        bound = 1e-4 * scale + 2.0 * sens
        relaxed = sens > 2e-5 * scale                 # a relaxation worth more than a fifth of the strict bound
        r = report(name, g, ref)
        r.update(max_sensitivity=float(sens.max()), frac_relaxed=float(relaxed.mean()),
                 max_scaled_where_strict=float((d / scale)[~relaxed].max()), frac_within_strict=float((d <= 1e-4 * scale).mean()))
        print(r)
        assert (d <= bound).all(), f"{name}: {int((d > bound).sum())} texels outside the bound, worst {float((d - bound).max()):.3e} ({r})"
        # the relaxation must stay an exception: whatever S says, at least 99.5 % of the texels meet the STRICT bound
        assert r["frac_within_strict"] >= 0.995, f"{name}: {r}"

    # one full face of one mip: mip 1 (256^2, roughness 1/8); face 3 (-Y) holds a pole, face 0 does not
    for m, f in ((1, 3), (1, 0)):
        nn = res >> m
        a = vq.cubemap_offset(res, m, f)
        check(f"C5 specular mip{m} face{f} ({nn}^2)", np.arange(a, a + nn * nn, dtype=np.int64))
    # 1 % of all texels, uniformly at random over the packed cube (so mostly mips 0-2, like the work itself)
    rng = np.random.default_rng(0x5EED0005)
    check("C5 specular 1% random texels", np.sort(rng.choice(n, size=n // 100, replace=False).astype(np.int64)))
    # and every texel of the small mips (3..8), where one texel integrates a wide lobe
    check("C5 specular mips 3..8 complete", np.arange(vq.cubemap_offset(res, 3, 0), n, dtype=np.int64))


def test_c5_forward_8k_row_tiles(ctx, vq, orc, envk):
    """BASELINE config 5's forward half: 7680x4320 split into the 8-rank row partition (540 rows per rank). Ranks 0 and 5
    are shaded as a rank would (its own row range of the full-size G-buffer) and compared with the oracle over the whole tile."""
    from vqengine_b200 import synth, distributed as vd
    w, h = 7680, 4320
    planes = synth.gbuffer(w, h, seed=synth.SEED_BASE + 5)
    pf, pv = synth.scene_constants(w, h, envk["spec_mips"])
    dpl = [dev(p) for p in planes]
    gb = vq.GBuffer(vq.image_of(dpl[0]), vq.image_of(dpl[1]), vq.image_of(dpl[2]), vq.null_image())
    out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    e = envk["np"]
    for rank in (0, 5):
        rb, re = vd.equal_tiles(h, 8)[rank]
        assert re - rb == 540
        ctx.forward_lighting(pf, pv, gb, envk["env"], out, rb, re)
        ref = orc.forward_lighting(pf, pv, planes, e["diff"], envk["diff_res"], e["spec"], envk["spec_res"], envk["spec_mips"],
                                   e["lut"], rb, re)
        got = host(out[rb:re])
        print(assert_scaled(f"C5 forward 7680x4320 rank {rank} rows {rb}..{re}", got, ref[rb:re]))
    assert bool((out[540:2700] == 0).all())                         # rows of other ranks untouched
