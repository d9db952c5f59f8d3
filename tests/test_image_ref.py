"""CPU (-m "not gpu"): the reference's OWN Image class — Libs/VQUtils/Source/Image.cpp compiled unmodified and in place into
oracle/_ref/libvqimageref.so (oracle/Makefile; only the MSVC-only Log.h / utils.h are shadowed) — against the oracle:
Image::LoadFromFile (texels AND the MaxLuminance it stores), Image::CreateResizedImage, Image::SaveToDisk,
Image::CalculateMipLevelCount. This is the engine's code path end to end, one level above the stb pins."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(orc):
    if orc.image_ref() is None:
        pytest.skip("oracle/_ref/libvqimageref.so not built (no /root/reference here)")
    return orc


def _img(w, h, seed):
    rng = np.random.default_rng(seed)
    a = (rng.random((h, w, 4), dtype=np.float32) ** 2 * 7).astype(np.float32)
    a[:, : w // 3, :3] = np.float32(0.25)
    a[..., 3] = 1.0
    return a


@pytest.mark.parametrize("w,h", [(64, 8), (7, 3), (300, 5), (8, 1), (129, 2)])
def test_load_from_file_texels_and_max_luminance(ref, vq, tmp_path, w, h):
    orc = ref
    path = str(tmp_path / f"in_{w}x{h}.hdr")
    data = orc.hdr_encode(_img(w, h, w + h))
    open(path, "wb").write(data)
    texels, lum = orc.ref_image_load(path)                 # Image::LoadFromFile
    rc, mine, my_lum = orc.hdr_decode(data)
    assert rc == 0 and texels is not None
    assert np.array_equal(texels.view(np.uint32), mine.view(np.uint32))
    assert np.float32(lum) == np.float32(my_lum)           # Image::MaxLuminance == CalculateMaxLuminance restated
    info, _ = vq.hdr_parse(data)                           # the product's host parser sees the same image
    assert (info.width, info.height) == (w, h)


@pytest.mark.parametrize("w,h,ow,oh", [(64, 32, 32, 16), (100, 37, 41, 13), (128, 64, 16, 8), (33, 17, 33, 9)])
def test_create_resized_image(ref, w, h, ow, oh):
    a = _img(w, h, w * 3 + h)
    assert np.array_equal(ref.ref_image_resize(a, ow, oh).view(np.uint32), ref.resize_downsample(a, ow, oh).view(np.uint32))


@pytest.mark.parametrize("w,h", [(64, 8), (7, 3), (300, 5)])
def test_save_to_disk_is_byte_identical(ref, vq, tmp_path, w, h):
    a = _img(w, h, 11 * w + h)
    path = str(tmp_path / "out.hdr")
    assert ref.ref_image_save(path, a)                     # Image::SaveToDisk
    data = open(path, "rb").read()
    assert data == ref.hdr_encode(a)
    assert data == vq.hdr_pack_file(ref.linear_to_rgbe(a))  # the product's host packer on the oracle's RGBE texels


def test_calculate_mip_level_count(ref, vq):
    for w, h in [(2048, 1024), (4096, 2048), (4096, 4096), (512, 512), (1, 1), (3, 1000), (8192, 4096), (640, 360)]:
        want = ref.ref_mip_level_count(w, h)               # Image::CalculateMipLevelCount
        assert want == int(ref.lib().orc_mip_level_count(w, h)) == vq.mip_level_count(w, h), (w, h)


def test_engine_downsize_flow_4k_to_1k(ref, tmp_path):
    """CreateEnvironmentMapTextureFromHiResAndSaveToDisk (EnvironmentMap.cpp:142-209) through the reference's Image class:
    LoadFromFile -> CreateResizedImage -> SaveToDisk, against oracle decode -> resize -> encode: identical file"""
    from vqengine_b200 import synth
    src = synth.hdri(512, 256)
    hi = str(tmp_path / "hi.hdr"); lo = str(tmp_path / "lo.hdr")
    open(hi, "wb").write(ref.hdr_encode(src))
    texels, _ = ref.ref_image_load(hi)
    small = ref.ref_image_resize(texels, 128, 64)
    assert ref.ref_image_save(lo, small)
    rc, dec, _ = ref.hdr_decode(open(hi, "rb").read())
    assert open(lo, "rb").read() == ref.hdr_encode(ref.resize_downsample(dec, 128, 64))
