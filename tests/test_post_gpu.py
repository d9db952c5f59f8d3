"""-m gpu parity: post-chain kernels (K5-K10) against the scalar oracle, through the C-ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import dev, host, assert_abs, assert_scaled, report

pytestmark = pytest.mark.gpu

SIZES = [(64, 48), (333, 129), (1, 1), (17, 5), (640, 360)]


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("curve,gamma,cs", [(0, 1, 0), (0, 0, 0), (1, 0, 0), (1, 0, 1), (2, 0, 0), (7, 0, 0)])
def test_tonemap(ctx, vq, orc, w, h, curve, gamma, cs):
    from vqengine_b200 import synth
    img = synth.hdr_image(w, h, seed=11 + w)
    p = vq.TonemapperParams(cs, curve, 200.0, gamma, 1.0)
    out = torch.empty_like(dev(img))
    ctx.tonemap(p, dev(img), out)
    assert_abs(f"tonemap{w}x{h}", host(out), orc.tonemap(p, img))


def test_tonemap_kat(ctx, vq):
    # in 1.0, sRGB curve, gamma on -> Reinhard 0.5 -> 0.7353569
    img = np.ones((4, 4, 4), np.float32)
    out = torch.empty_like(dev(img))
    ctx.tonemap(vq.TonemapperParams(0, 0, 200.0, 1, 1.0), dev(img), out)
    assert np.allclose(host(out)[..., :3], 0.7353569, atol=2e-7)
    assert (host(out)[..., 3] == 1.0).all()


@pytest.mark.parametrize("w,h", SIZES + [(1030, 70)])
def test_gaussian_blur(ctx, orc, w, h):
    from vqengine_b200 import synth
    img = synth.hdr_image(w, h, seed=5)
    d = dev(img)
    tmp, out = torch.empty_like(d), torch.empty_like(d)
    ctx.gaussian_blur(d, tmp, vertical=False)
    ctx.gaussian_blur(tmp, out, vertical=True)
    rx = orc.gaussian_blur(img, False)
    assert_abs(f"blurX{w}x{h}", host(tmp), rx, tol=1e-5)
    assert_abs(f"blurXY{w}x{h}", host(out), orc.gaussian_blur(rx, True), tol=1e-5)
    assert (host(out)[..., 3] == 1.0).all()


def test_gaussian_blur_constant_fixed_point(ctx):
    # weights sum to 1.000000 -> a constant image is a fixed point up to rounding
    img = np.full((40, 50, 4), 0.625, np.float32)
    d = dev(img)
    a, b = torch.empty_like(d), torch.empty_like(d)
    ctx.gaussian_blur(d, a, False)
    ctx.gaussian_blur(a, b, True)
    assert np.abs(host(b)[..., :3] - 0.625).max() < 2e-6


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("sharp", [0.0, 0.8, 1.0])
def test_cas(ctx, vq, orc, w, h, sharp):
    from vqengine_b200 import synth
    img = orc.tonemap(synth.default_tonemapper(), synth.hdr_image(w, h, seed=9))   # CAS expects [0,1]
    con = vq.cas_setup(sharp, w, h, w, h)
    out = torch.empty_like(dev(img))
    ctx.cas(con, dev(img), out)
    assert_abs(f"cas{w}x{h}", host(out), orc.cas(con, img), tol=1e-5)


@pytest.mark.parametrize("iw,ih,ow,oh", [(64, 48, 128, 96), (100, 60, 130, 78), (37, 21, 74, 42), (8, 8, 8, 8),
                                         (320, 180, 640, 360), (77, 43, 100, 56)])
@pytest.mark.parametrize("mode", [0, 1])
def test_fsr_easu(ctx, vq, orc, iw, ih, ow, oh, mode):
    from vqengine_b200 import synth
    img = orc.tonemap(synth.default_tonemapper(), synth.hdr_image(iw, ih, seed=21))
    con = vq.fsr_easu_con(iw, ih, iw, ih, ow, oh)
    out = torch.empty((oh, ow, 4), dtype=torch.float32, device="cuda")
    ctx.fsr_easu(con, dev(img), out, address_mode=mode)
    assert_abs(f"easu{iw}x{ih}->{ow}x{oh}", host(out), orc.fsr_easu(con, img, ow, oh, mode), tol=2e-5)


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("stops", [0.0, 0.2, 2.0])
def test_fsr_rcas(ctx, vq, orc, w, h, stops):
    from vqengine_b200 import synth
    img = orc.tonemap(synth.default_tonemapper(), synth.hdr_image(w, h, seed=31))
    con = vq.fsr_rcas_con(stops)
    out = torch.empty_like(dev(img))
    ctx.fsr_rcas(con, dev(img), out)
    assert_abs(f"rcas{w}x{h}", host(out), orc.fsr_rcas(con, img), tol=1e-5)


@pytest.mark.parametrize("w,h", [(64, 64), (128, 64), (200, 120), (960, 540), (65, 33), (4096, 64), (1000, 1000)])
def test_spd(ctx, vq, orc, w, h):
    from vqengine_b200 import synth
    img = synth.hdr_image(w, h, seed=41)
    (dx, dy), c = vq.spd_setup(w, h)
    mips = min(c.mips, int(np.floor(np.log2(min(w, h)))))   # floor-sized chain ends at min dim 1
    c.mips = mips
    dsts = [torch.zeros((h >> l, w >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, mips + 1)]
    for rep in range(2):   # twice: the global ticket counter must reset itself
        ctx.spd_downsample(c, dev(img), dsts)
    ref = orc.spd_downsample(img, mips)
    assert len(ref) == mips
    for l, (g, r) in enumerate(zip(dsts, ref), start=1):
        assert np.array_equal(host(g), r), f"spd level {l} of {w}x{h}: not bit-exact, max diff {np.abs(host(g) - r).max()}"


def test_spd_two_streams_in_flight(ctx, vq, orc):
    """vqcuda.h: calls are thread-safe on distinct streams. Two SPD pyramids with a last-workgroup tail (mips > 6) are kept in
    flight on two streams of ONE context, many times over: every launch owns its ticket word, so each result stays bit-exact."""
    from vqengine_b200 import synth
    sizes = [(1024, 512), (960, 540)]
    imgs = [synth.hdr_image(w, h, seed=50 + i) for i, (w, h) in enumerate(sizes)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    work = []
    for (w, h), img in zip(sizes, imgs):
        (dx, dy), c = vq.spd_setup(w, h)
        c.mips = min(c.mips, int(np.floor(np.log2(min(w, h)))))
        assert c.mips > 6
        work.append((c, dev(img), [torch.zeros((h >> l, w >> l, 4), dtype=torch.float32, device="cuda") for l in range(1, c.mips + 1)]))
    torch.cuda.synchronize()
    for rep in range(40):
        for s, (c, d, dsts) in zip(streams, work):
            ctx.spd_downsample(c, d, dsts, stream=s)
    torch.cuda.synchronize()
    for img, (c, d, dsts) in zip(imgs, work):
        for l, (g, r) in enumerate(zip(dsts, orc.spd_downsample(img, c.mips)), start=1):
            assert np.array_equal(host(g), r), f"level {l}"


def test_post_chain_config4_small(ctx, vq, orc):
    """BASELINE config 4 order at a small size: SPD -> BlurX,Y -> Tonemap -> CAS -> EASU 2x -> RCAS.
    Every stage is checked against the oracle applied to the SAME input (the kernel's previous-stage output):
    CAS/EASU/RCAS contain sqrt-like and min/max non-smooth steps that amplify 1e-7 input differences, so an
    end-to-end comparison of two independently rounded chains is reported, not asserted at 1e-4."""
    from vqengine_b200 import synth
    w, h = 480, 270
    img = synth.hdr_image(w, h, seed=4)
    d = dev(img)
    a, b, t, c = (torch.empty_like(d) for _ in range(4))
    tm = synth.default_tonemapper()
    e = torch.empty((2 * h, 2 * w, 4), dtype=torch.float32, device="cuda")
    r = torch.empty_like(e)
    ctx.gaussian_blur(d, a, False)
    assert_abs("chain.blur_x", host(a), orc.gaussian_blur(img, False), tol=1e-5)
    ctx.gaussian_blur(a, b, True)
    assert_abs("chain.blur_y", host(b), orc.gaussian_blur(host(a), True), tol=1e-5)
    ctx.tonemap(tm, b, t)
    assert_abs("chain.tonemap", host(t), orc.tonemap(tm, host(b)), tol=1e-5)
    ctx.cas(vq.cas_setup(0.8, w, h, w, h), t, c)
    assert_abs("chain.cas", host(c), orc.cas(orc.cas_setup(0.8, w, h, w, h), host(t)), tol=1e-5)
    ctx.fsr_easu(vq.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), c, e)
    assert_abs("chain.easu", host(e), orc.fsr_easu(orc.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), host(c), 2 * w, 2 * h, 0), tol=2e-5)
    ctx.fsr_rcas(vq.fsr_rcas_con(0.2), e, r)
    assert_abs("chain.rcas", host(r), orc.fsr_rcas(orc.fsr_rcas_con(0.2), host(e)), tol=1e-5)
    o = orc.gaussian_blur(orc.gaussian_blur(img, False), True)
    o = orc.tonemap(tm, o)
    o = orc.cas(orc.cas_setup(0.8, w, h, w, h), o)
    o = orc.fsr_easu(orc.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), o, 2 * w, 2 * h, 0)
    o = orc.fsr_rcas(orc.fsr_rcas_con(0.2), o)
    rep = report("chain_end_to_end", host(r), o)
    print(rep)
    assert rep["frac_abs_le_tol"] > 0.999 and rep["max_abs"] < 2e-2


def test_bad_arguments(ctx, vq):
    d = torch.zeros((8, 8, 4), device="cuda")
    with pytest.raises(vq.VqError) as e:
        ctx.gaussian_blur(d, d, False)      # in-place
    assert e.value.code == vq.VQ_ERR_INVALID_ARG
    with pytest.raises(vq.VqError):
        ctx.cas(vq.cas_setup(0.5, 8, 8, 8, 8), d, torch.zeros((4, 4, 4), device="cuda"))


def _pitched(t, pad=5):
    """a view with pitch > width*16: the left part of a wider tensor"""
    h, w, c = t.shape
    big = torch.full((h, w + pad, c), -7.0, dtype=torch.float32, device="cuda")
    big[:, :w] = t
    return big[:, :w], big


def test_pitched_images_all_post_passes(ctx, vq, orc):
    """pitch_bytes != width*16 on inputs AND outputs; pad columns must stay untouched"""
    from vqengine_b200 import synth
    w, h = 150, 67
    img = synth.hdr_image(w, h, seed=77)
    ldr = orc.tonemap(synth.default_tonemapper(), img)
    vin, _ = _pitched(dev(img)); vldr, _ = _pitched(dev(ldr))
    vout, big = _pitched(torch.zeros((h, w, 4), device="cuda"), pad=9)
    tm = synth.default_tonemapper()
    ctx.tonemap(tm, vin, vout)
    assert_abs("pitched.tonemap", host(vout.contiguous()), orc.tonemap(tm, img))
    ctx.gaussian_blur(vin, vout, False); assert_abs("pitched.blur_x", host(vout.contiguous()), orc.gaussian_blur(img, False), tol=1e-5)
    ctx.gaussian_blur(vin, vout, True); assert_abs("pitched.blur_y", host(vout.contiguous()), orc.gaussian_blur(img, True), tol=1e-5)
    ctx.cas(vq.cas_setup(0.8, w, h, w, h), vldr, vout); assert_abs("pitched.cas", host(vout.contiguous()), orc.cas(orc.cas_setup(0.8, w, h, w, h), ldr), tol=1e-5)
    ctx.fsr_rcas(vq.fsr_rcas_con(0.2), vldr, vout); assert_abs("pitched.rcas", host(vout.contiguous()), orc.fsr_rcas(orc.fsr_rcas_con(0.2), ldr), tol=1e-5)
    assert (host(big)[:, w:] == -7.0).all(), "a kernel wrote into the pitch padding"
    eo, ebig = _pitched(torch.zeros((2 * h, 2 * w, 4), device="cuda"), pad=3)
    ctx.fsr_easu(vq.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), vldr, eo)
    assert_abs("pitched.easu", host(eo.contiguous()), orc.fsr_easu(orc.fsr_easu_con(w, h, w, h, 2 * w, 2 * h), ldr, 2 * w, 2 * h, 0), tol=2e-5)
    assert (host(ebig)[:, 2 * w:] == -7.0).all()
    (dx, dy), c = vq.spd_setup(w, h)
    c.mips = 6
    dsts = [_pitched(torch.zeros((h >> l, w >> l, 4), device="cuda"), pad=l)[0] for l in range(1, 7)]
    ctx.spd_downsample(c, vin, dsts)
    for l, (g, r) in enumerate(zip(dsts, orc.spd_downsample(img, 6)), start=1):
        assert np.array_equal(host(g.contiguous()), r), l


@pytest.mark.parametrize("scale", [1.3, 1.5, 1.7, 2.0, 3.0])
def test_fsr_easu_presets(ctx, vq, orc, scale):
    """the engine's FSR1 presets (PostProcess.h:42-52: 0.77/0.67/0.58/0.50 per axis) and a 3x case: every output pixel is
    written exactly once by the per-input-texel kernel"""
    from vqengine_b200 import synth
    ow, oh = 384, 216
    iw, ih = int(round(ow / scale)), int(round(oh / scale))
    img = orc.tonemap(synth.default_tonemapper(), synth.hdr_image(iw, ih, seed=3))
    con = vq.fsr_easu_con(iw, ih, iw, ih, ow, oh)
    out = torch.full((oh, ow, 4), float("nan"), dtype=torch.float32, device="cuda")
    ctx.fsr_easu(con, dev(img), out)
    assert_abs(f"easu x{scale}", host(out), orc.fsr_easu(con, img, ow, oh, 0), tol=2e-5)


def test_fsr_easu_downscale_uses_generic_path(ctx, vq, orc):
    from vqengine_b200 import synth
    img = orc.tonemap(synth.default_tonemapper(), synth.hdr_image(120, 80, seed=3))
    con = vq.fsr_easu_con(120, 80, 120, 80, 90, 60)
    out = torch.full((60, 90, 4), float("nan"), dtype=torch.float32, device="cuda")
    ctx.fsr_easu(con, dev(img), out)
    assert_abs("easu down", host(out), orc.fsr_easu(con, img, 90, 60, 0), tol=2e-5)


def test_spd_max_size_4096(ctx, vq, orc):
    """the header's limit (4096^2, 12 mips): bit-exact, and the ticket counter survives back-to-back launches"""
    rng = np.random.default_rng(9)
    img = rng.uniform(0, 4, (4096, 4096, 4)).astype(np.float32)
    (dx, dy), c = vq.spd_setup(4096, 4096)
    assert c.mips == 12 and c.numWorkGroups == 64 * 64
    dsts = [torch.zeros((4096 >> l, 4096 >> l, 4), device="cuda") for l in range(1, 13)]
    d = dev(img)
    ctx.spd_downsample(c, d, dsts); ctx.spd_downsample(c, d, dsts)
    for l, (g, r) in enumerate(zip(dsts, orc.spd_downsample(img, 12)), start=1):
        assert np.array_equal(host(g), r), l
