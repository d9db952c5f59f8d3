"""CPU (-m "not gpu"): SURVEY §8(f).2 / (f).3 — pins the .hdr codec restatement (oracle/oracle_frame.cpp) against
(1) the committed golden files written and decoded by the reference's own stb codec (tests/golden/make_hdr_golden.py),
(2) that codec itself when oracle/_ref/libstbref.so is present, on crafted streams (runs of 1, zero-length records,
flat first scanline, header variants, corrupt and truncated data); checks the HOST half of the product's codec
(vq_hdr_parse, vq_hdr_pack_file — no GPU involved) against the oracle; and the skydome / ApplyReflections oracle
against closed forms."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "hdr_golden.json")))
HEADER = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n"


def _gold_file(name):
    return open(os.path.join(HERE, "golden", f"hdr_{name}.hdr"), "rb").read()


def _rand_image(w, h, seed=1):
    rng = np.random.default_rng(seed)
    a = (rng.random((h, w, 4), dtype=np.float32) ** 2 * 6.0).astype(np.float32)
    a[:, : max(w // 4, 1), :3] = np.float32(0.5)
    a[..., 3] = 1.0
    return a


def _rle_scanline(rng, width, style):
    """one scanline record {2,2,hi,lo} + 4 run lists in the given style; returns (bytes, planes [4,width] uint8)"""
    out = bytearray([2, 2, width >> 8, width & 255])
    planes = np.zeros((4, width), dtype=np.uint8)
    for k in range(4):
        i = 0
        while i < width:
            left = width - i
            if style == "runs_of_1":
                v = int(rng.integers(0, 256)); out += bytes([129, v]); planes[k, i] = v; i += 1
            elif style == "zero_records" and rng.random() < 0.3:
                out += bytes([0])                               # zero-length literal record: a no-op byte
            else:
                if rng.random() < 0.5:
                    n = int(min(left, rng.integers(1, 128))); v = int(rng.integers(0, 256))
                    out += bytes([128 + n, v]); planes[k, i:i + n] = v
                else:
                    n = int(min(left, rng.integers(1, 129))); vals = rng.integers(0, 256, n, dtype=np.uint8)
                    out += bytes([n]) + vals.tobytes(); planes[k, i:i + n] = vals
                i += n
    return bytes(out), planes


def _crafted(width, height, style, seed=3, header=HEADER):
    rng = np.random.default_rng(seed)
    body = bytearray(); rows = []
    for _ in range(height):
        b, p = _rle_scanline(rng, width, style)
        body += b; rows.append(p.T.copy())                     # [width,4] RGBE
    return header + f"-Y {height} +X {width}\n".encode() + bytes(body), np.stack(rows)


def _rgbe_to_float(rgbe):
    e = rgbe[..., 3].astype(np.int32)
    f = np.ldexp(np.float32(1.0), e - 136).astype(np.float32)
    out = np.ones(rgbe.shape[:-1] + (4,), dtype=np.float32)
    out[..., :3] = np.where(e[..., None] != 0, rgbe[..., :3].astype(np.float32) * f[..., None], np.float32(0.0))
    return out


# ---- (1) golden files written/decoded by the reference's stb ------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GOLD))
def test_golden_decode_and_encode(orc, name):
    g = GOLD[name]
    data = _gold_file(name)
    assert hashlib.sha256(data).hexdigest() == g["file_sha256"]
    rc, dec, lum = orc.hdr_decode(data)
    assert rc == 0 and dec.shape == (g["height"], g["width"], 4)
    assert hashlib.sha256(dec.tobytes()).hexdigest() == g["decoded_f32_sha256"]          # bit-for-bit stbi_loadf
    assert [f"{x:08x}" for x in dec.reshape(-1).view(np.uint32)[:16]] == g["decoded_first_texels_hex"]
    assert f"{np.float32(lum).view(np.uint32):08x}" == g["max_luminance_hex"]
    src = np.load(os.path.join(HERE, "golden", f"hdr_{name}_src.npy"))
    assert hashlib.sha256(src.tobytes()).hexdigest() == g["source_f32_sha256"]
    assert orc.hdr_encode(src) == data                                                    # byte-for-byte stbi_write_hdr


# ---- (2) against the reference codec itself, where it was built ---------------------------------------------------------
def _need_ref(orc):
    if orc.stb_ref() is None:
        pytest.skip("oracle/_ref/libstbref.so not built (no /root/reference here)")


@pytest.mark.parametrize("w,h", [(8, 1), (9, 5), (64, 8), (127, 3), (128, 2), (129, 2), (300, 4), (7, 3), (1, 1)])
def test_encode_decode_equal_reference(orc, w, h):
    _need_ref(orc)
    a = _rand_image(w, h, seed=w * 31 + h)
    f_o, f_r = orc.hdr_encode(a), orc.hdr_encode(a, "ref")
    assert f_o == f_r
    rc, d_o, _ = orc.hdr_decode(f_r)
    rc2, d_r, _ = orc.hdr_decode(f_r, "ref")
    assert rc == 0 and rc2 == 0 and np.array_equal(d_o.view(np.uint32), d_r.view(np.uint32))


@pytest.mark.parametrize("style", ["mixed", "runs_of_1", "zero_records"])
def test_crafted_streams_equal_reference(orc, style):
    data, rgbe = _crafted(40, 5, style)
    rc, d_o, _ = orc.hdr_decode(data)
    assert rc == 0
    assert np.array_equal(d_o.view(np.uint32), _rgbe_to_float(rgbe).view(np.uint32))      # independent numpy restatement
    if orc.stb_ref() is not None:
        rc2, d_r, _ = orc.hdr_decode(data, "ref")
        assert rc2 == 0 and np.array_equal(d_o.view(np.uint32), d_r.view(np.uint32))


def test_header_variants_and_flat_first_scanline(orc):
    rng = np.random.default_rng(9)
    px = rng.integers(1, 256, (3, 16, 4), dtype=np.uint8)
    px[0, 0] = (200, 100, 50, 130)                              # first three bytes are not {2,2,len<128}: flat file
    for header in (b"#?RGBE\nFORMAT=32-bit_rle_rgbe\n\n",
                   b"#?RADIANCE\n# comment\nEXPOSURE=1.0\nFORMAT=32-bit_rle_rgbe\nGAMMA=1\n\n",
                   b"#?RADIANCE\n" + b"x" * 2000 + b"\nFORMAT=32-bit_rle_rgbe\n\n"):
        data = header + b"-Y 3 +X 16\n" + px.tobytes()
        rc, d_o, _ = orc.hdr_decode(data)
        assert rc == 0 and np.array_equal(d_o.view(np.uint32), _rgbe_to_float(px).view(np.uint32))
        if orc.stb_ref() is not None:
            rc2, d_r, _ = orc.hdr_decode(data, "ref")
            assert rc2 == 0 and np.array_equal(d_o.view(np.uint32), d_r.view(np.uint32))


def test_decode_failures_match_reference(orc):
    good, _ = _crafted(16, 2, "mixed")
    bad = {
        "not_hdr": b"#?RADIANCF\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 16\n",
        "no_format": b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 2 +X 16\n",
        "layout": HEADER + b"+Y 2 +X 16\n",
        "layout2": HEADER + b"-Y 2 -X 16\n",
        "scanline_len": HEADER + b"-Y 2 +X 16\n" + bytes([2, 2, 0, 15]) + bytes(64),
        "run_overflow": HEADER + b"-Y 1 +X 16\n" + bytes([2, 2, 0, 16, 128 + 17, 5]) + bytes(64),
        "dump_overflow": HEADER + b"-Y 1 +X 16\n" + bytes([2, 2, 0, 16, 17]) + bytes(64),
    }
    want = {"not_hdr": 1, "no_format": 2, "layout": 3, "layout2": 3, "scanline_len": 4, "run_overflow": 5, "dump_overflow": 5}
    for k, data in bad.items():
        rc, _, _ = orc.hdr_decode(data)
        assert rc == want[k], k
        if orc.stb_ref() is not None:
            assert orc.hdr_decode(data, "ref")[0] != 0, k
    assert orc.hdr_decode(good)[0] == 0
    assert orc.hdr_decode(good[:-3])[0] == 5                    # ends inside a scanline (documented: stb would not return)


def test_linear_to_rgbe_properties(orc):
    a = np.array([[0, 0, 0, 1], [1e-33, 0, 0, 1], [1, 1, 1, 1], [0.5, 0.25, 0.125, 1], [255.9, 3, 2, 1], [1e30, 1e29, 0, 1],
                  [9.9e-33, 0, 0, 1], [1.1e-32, 0, 0, 1]], dtype=np.float32)
    r = orc.linear_to_rgbe(a)
    assert (r[0] == 0).all() and (r[1] == 0).all() and (r[6] == 0).all() and r[7][3] != 0
    assert tuple(r[2]) == (128, 128, 128, 129) and tuple(r[3]) == (128, 64, 32, 128)
    # decode(encode(x)) <= x and within 1/128 relative of the largest channel; encode is idempotent on decoded values
    rng = np.random.default_rng(5)
    x = (rng.random((4096, 4), dtype=np.float32) * np.float32(50.0)).astype(np.float32)
    q = _rgbe_to_float(orc.linear_to_rgbe(x))
    assert (q[:, :3] <= x[:, :3]).all()
    assert (x[:, :3] - q[:, :3] <= x[:, :3].max(axis=1, keepdims=True) / 128.0 + 1e-30).all()
    assert np.array_equal(orc.linear_to_rgbe(q), orc.linear_to_rgbe(x))


# ---- host half of the product codec (C-ABI, no GPU) ---------------------------------------------------------------------
def test_host_parse_index(vq, orc):
    for style in ("mixed", "runs_of_1", "zero_records"):
        data, rgbe = _crafted(33, 4, style, seed=11)
        info, offs = vq.hdr_parse(data)
        assert (info.width, info.height, info.flat) == (33, 4, 0) and offs.shape == (17,)
        assert list(offs) == sorted(offs) and offs[-1] == len(data)
        # every scanline's channel 0 starts right after its 4-byte marker
        for j in range(4):
            o = int(offs[4 * j])
            assert data[o - 4:o] == bytes([2, 2, 0, 33])
    info, offs = vq.hdr_parse(_gold_file("flat_5x4"))
    assert (info.width, info.height, info.flat) == (5, 4, 1) and offs is None
    flat_first = HEADER + b"-Y 2 +X 16\n" + bytes([200, 100, 50, 130]) + bytes(124)
    info, offs = vq.hdr_parse(flat_first)
    assert info.flat == 1 and offs is None and info.data_offset == len(HEADER) + len(b"-Y 2 +X 16\n")


def test_host_parse_failures(vq, orc):
    cases = [b"#?RADIANCF\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 16\n", b"#?RADIANCE\n\n-Y 2 +X 16\n", HEADER + b"+Y 2 +X 16\n",
             HEADER + b"-Y 2 +X 16\n" + bytes([2, 2, 0, 15]) + bytes(64),
             HEADER + b"-Y 1 +X 16\n" + bytes([2, 2, 0, 16, 128 + 17, 5]) + bytes(64),
             HEADER + b"-Y 1 +X 16\n" + bytes([2, 2, 0, 16, 17]) + bytes(64),
             _crafted(16, 2, "mixed")[0][:-3]]
    for data in cases:
        assert orc.hdr_decode(data)[0] != 0
        with pytest.raises(vq.VqError):
            vq.hdr_parse(data)


@pytest.mark.parametrize("w,h", [(8, 1), (64, 8), (300, 4), (7, 3), (129, 2)])
def test_host_pack_file_equals_oracle(vq, orc, w, h):
    a = _rand_image(w, h, seed=w + h)
    rgbe = orc.linear_to_rgbe(a)
    assert vq.hdr_pack_file(rgbe) == orc.hdr_encode(a)
    if orc.stb_ref() is not None:
        assert vq.hdr_pack_file(rgbe) == orc.hdr_encode(a, "ref")


def test_host_pack_golden(vq, orc):
    for name in GOLD:
        src = np.load(os.path.join(HERE, "golden", f"hdr_{name}_src.npy"))
        assert vq.hdr_pack_file(orc.linear_to_rgbe(src)) == _gold_file(name)


# ---- skydome / ApplyReflections oracle ----------------------------------------------------------------------------------
def test_skydome_oracle_closed_forms(orc):
    from vqengine_b200 import synth
    hw, hh = 64, 32
    const = np.zeros((hh, hw, 4), dtype=np.float32); const[...] = (0.25, 0.5, 2.0, 1.0)
    vp, inv = synth.sky_view_proj(0.3, -0.2, 1.0, 16 / 9)
    scene = np.zeros((9, 16, 4), dtype=np.float32)
    orc.skydome(const.reshape(-1, 4), hw, hh, 1, inv, scene)
    assert np.abs(scene - np.array([0.25, 0.5, 2.0, 1.0], dtype=np.float32)).max() <= 1e-6
    # the centre of an odd-sized frame looks along the camera's forward axis: +Z for yaw = pitch = 0,
    # DirectionToEquirectUV(+Z) = (0.25, 0.5)  (ShadingMath.hlsl:70-80: u = atan2(z,x)/(-2pi)+0.5)
    img = synth.smooth_hdri(hw, hh)
    vp, inv = synth.sky_view_proj(0.0, 0.0, 1.0, 1.0)
    scene = np.zeros((5, 5, 4), dtype=np.float32)
    orc.skydome(img.reshape(-1, 4), hw, hh, 1, inv, scene)
    x, y = 0.25 * hw - 0.5, 0.5 * hh - 0.5
    x0, y0 = int(np.floor(x)), int(np.floor(y)); fx, fy = x - x0, y - y0
    want = (img[y0, x0] * (1 - fx) + img[y0, x0 + 1] * fx) * (1 - fy) + (img[y0 + 1, x0] * (1 - fx) + img[y0 + 1, x0 + 1] * fx) * fy
    assert np.abs(scene[2, 2, :3] - want[:3]).max() <= 2e-5 and scene[2, 2, 3] == 1.0
    # mask: only texels whose normal is exactly zero are written
    mask = np.ones((5, 5, 4), dtype=np.float32); mask[1, 3, :3] = 0.0
    scene2 = np.full((5, 5, 4), 7.0, dtype=np.float32)
    orc.skydome(img.reshape(-1, 4), hw, hh, 1, inv, scene2, normal_mask=mask)
    assert (scene2[1, 3] == scene[1, 3]).all()
    scene2[1, 3] = 7.0
    assert (scene2 == 7.0).all()


def test_apply_reflections_oracle(orc):
    rng = np.random.default_rng(2)
    s = rng.random((6, 10, 4), dtype=np.float32); r = rng.random((6, 10, 4), dtype=np.float32); bv = rng.random((6, 10, 4), dtype=np.float32)
    o = orc.apply_reflections(s.copy(), r)
    assert np.array_equal(o[..., :3], s[..., :3] + r[..., :3]) and np.array_equal(o[..., 3], s[..., 3])
    o2 = orc.apply_reflections(s.copy(), r, bv)
    want = bv[..., :3] * bv[..., 3:4] + (s[..., :3] + r[..., :3]) * (np.float32(1.0) - bv[..., 3:4])
    assert np.array_equal(o2[..., :3], want) and np.array_equal(o2[..., 3], bv[..., 3])
