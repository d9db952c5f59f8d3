"""CPU (-m "not gpu"): SURVEY §8(f).2, the HDRI downsize (Image::CreateResizedImage -> stbir_resize_float). Pins the oracle's
restatement of stb_image_resize v0.96 (oracle/oracle_resize.cpp) against (1) golden outputs produced by the reference's own
stbir (tests/golden/resize_golden.json, made by make_hdr_golden.py), (2) that code itself where oracle/_ref/libstbref.so
exists; and checks the product's HOST half — the per-axis gather tables vq_image_resize uploads — by replaying them in
numpy with the kernels' operation order: bit-identical to the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "resize_golden.json")))
CASES = [(64, 32, 32, 16), (64, 32, 16, 8), (100, 37, 41, 13), (33, 17, 33, 9), (16, 16, 16, 16), (128, 64, 16, 8), (50, 50, 49, 1),
         (257, 3, 100, 3), (9, 9, 1, 1)]


def _img(w, h, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((h, w, 4), dtype=np.float32) * 5).astype(np.float32)


def replay_gather_tables(vq, img, ow, oh):
    """what resize_h_kernel / resize_v_kernel compute: taps in increasing order, product rounded, then added (fp32)"""
    h, w = img.shape[:2]
    sh, ch, wh = vq.resize_axis_table(w, ow)
    sv, cv, wv = vq.resize_axis_table(h, oh)
    mid = np.zeros((h, ow, 4), np.float32)
    for t in range(wh.shape[1]):
        term = (img[:, np.clip(sh + t, 0, w - 1), :] * wh[None, :, t, None]).astype(np.float32)
        mid = np.where((t < ch)[None, :, None], (mid + term).astype(np.float32), mid)
    out = np.zeros((oh, ow, 4), np.float32)
    for t in range(wv.shape[1]):
        term = (mid[np.clip(sv + t, 0, h - 1), :, :] * wv[:, t, None, None]).astype(np.float32)
        out = np.where((t < cv)[:, None, None], (out + term).astype(np.float32), out)
    return out


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_reference_golden(orc, name):
    g = GOLD[name]
    rng = np.random.default_rng(0x5EED0000 + 22)
    src = None
    for n in GOLD:                                        # the generator draws the sources in file order from one stream
        gg = GOLD[n]
        s = (rng.random((gg["h"], gg["w"], 4), dtype=np.float32) ** 2 * 9.0).astype(np.float32)
        if n == name:
            src = s
            break
    assert hashlib.sha256(src.tobytes()).hexdigest() == g["source_f32_sha256"]
    out = orc.resize_downsample(src, g["ow"], g["oh"])
    assert [f"{x:08x}" for x in out.reshape(-1).view(np.uint32)[:12]] == g["resized_first_texels_hex"]
    assert hashlib.sha256(out.tobytes()).hexdigest() == g["resized_f32_sha256"]          # bit-for-bit stbir_resize_float


@pytest.mark.parametrize("w,h,ow,oh", CASES)
def test_oracle_equals_reference_where_built(orc, w, h, ow, oh):
    if orc.stb_ref() is None:
        pytest.skip("oracle/_ref/libstbref.so not built (no /root/reference here)")
    a = _img(w, h, w * 3 + h)
    assert np.array_equal(orc.resize_downsample(a, ow, oh).view(np.uint32), orc.resize_downsample(a, ow, oh, "ref").view(np.uint32))


@pytest.mark.parametrize("w,h,ow,oh", CASES)
def test_host_gather_tables_replay_equals_oracle(vq, orc, w, h, ow, oh):
    a = _img(w, h, w + 7 * h)
    assert np.array_equal(replay_gather_tables(vq, a, ow, oh).view(np.uint32), orc.resize_downsample(a, ow, oh).view(np.uint32))


def test_gather_table_properties(vq):
    for n_in, n_out in [(8192, 4096), (8192, 1024), (4096, 2048), (1000, 333), (64, 64)]:
        start, count, w = vq.resize_axis_table(n_in, n_out)
        assert (count >= 1).all() and (np.diff(start) >= 0).all()
        assert np.abs(w.sum(axis=1) - 1.0).max() <= 4e-7               # normalised per output sample
        radius = 2.0 * n_in / n_out
        assert count.max() <= int(np.ceil(2 * radius)) + 2
        centre = (np.arange(n_out) + 0.5) * n_in / n_out                 # taps straddle the output sample's centre
        assert ((start <= centre) & (start + count >= centre)).all()
    with pytest.raises(vq.VqError):
        vq.resize_axis_table(16, 32)                                    # upsizing is not the engine's path


def test_constant_image_and_mean(orc):
    c = np.zeros((40, 80, 4), np.float32); c[...] = (0.25, 3.0, 7.5, 1.0)
    o = orc.resize_downsample(c, 20, 10)
    assert np.abs(o - c[0, 0]).max() <= 2e-6
    a = _img(128, 64, 5)
    o = orc.resize_downsample(a, 64, 32)
    assert abs(float(o.mean()) - float(a.mean())) <= 2e-2


def test_random_size_pairs(vq, orc):
    """40 random (w, h) -> (ow <= w, oh <= h) pairs, ratios from 1:1 to 300:1: host tables replayed == oracle bit for bit,
    and oracle == the reference's stbir where it is built. (A 400-pair sweep of the same generator found no difference.)"""
    rng = np.random.default_rng(5)
    for _ in range(40):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 40))
        ow, oh = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        a = (rng.random((h, w, 4), dtype=np.float32) * 8).astype(np.float32)
        o = orc.resize_downsample(a, ow, oh)
        assert np.array_equal(replay_gather_tables(vq, a, ow, oh).view(np.uint32), o.view(np.uint32)), (w, h, ow, oh)
        if orc.stb_ref() is not None:
            assert np.array_equal(o.view(np.uint32), orc.resize_downsample(a, ow, oh, "ref").view(np.uint32)), (w, h, ow, oh)
