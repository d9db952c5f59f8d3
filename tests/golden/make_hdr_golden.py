"""Generates tests/golden/hdr_*.hdr + hdr_golden.json with the REFERENCE's own stb codec (oracle/_ref/libstbref.so =
/root/reference/Libs/VQUtils/Libs/stb/stb_image{,_write}.h compiled in place). Run in the build container only:
    python tests/golden/make_hdr_golden.py
The committed files pin the oracle (and through it the kernels) where /root/reference is absent."""
import hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))]
import oracle_lib as orc

assert orc.stb_ref() is not None, "oracle/_ref/libstbref.so missing: run `make -C oracle ref` where /root/reference exists"
rng = np.random.default_rng(0x5EED0000 + 21)


def image(w, h):
    a = (rng.random((h, w, 4), dtype=np.float32) ** 3 * 12.0).astype(np.float32)
    a[:, : w // 3, :3] = np.float32(0.75)                      # long runs
    a[h // 2:, w // 2:, :3] *= np.float32(1e-4)                # small exponents
    a[0, 0, :3] = 0.0                                          # exact zero texel
    a[-1, -1, :3] = np.float32(3e-33)                          # below the 1e-32 cut-off
    a[..., 3] = 1.0
    return a


meta = {}
for name, (w, h) in {"rle_48x6": (48, 6), "rle_200x3": (200, 3), "flat_5x4": (5, 4), "rle_8x2": (8, 2)}.items():
    src = image(w, h)
    data = orc.hdr_encode(src, "ref")                          # the reference's stbi_write_hdr
    rc, dec, _ = orc.hdr_decode(data, "ref")                   # the reference's stbi_loadf
    assert rc == 0
    open(os.path.join(HERE, f"hdr_{name}.hdr"), "wb").write(data)
    meta[name] = {"width": w, "height": h, "file_sha256": hashlib.sha256(data).hexdigest(),
                  "source_f32_sha256": hashlib.sha256(src.tobytes()).hexdigest(),
                  "decoded_f32_sha256": hashlib.sha256(dec.tobytes()).hexdigest(),
                  "decoded_first_texels_hex": [f"{x:08x}" for x in dec.reshape(-1).view(np.uint32)[:16]],
                  "max_luminance_hex": f"{np.float32(orc.hdr_decode(data)[2]).view(np.uint32):08x}"}
    np.save(os.path.join(HERE, f"hdr_{name}_src.npy"), src)
json.dump(meta, open(os.path.join(HERE, "hdr_golden.json"), "w"), indent=1)
print("wrote", list(meta))

# ---- stbir_resize_float golden (Image::CreateResizedImage): input regenerated from the seed, output bits from the reference ----
rz = {}
rng2 = np.random.default_rng(0x5EED0000 + 22)
for name, (w, h, ow, oh) in {"64x32_to_32x16": (64, 32, 32, 16), "100x37_to_41x13": (100, 37, 41, 13), "96x48_to_12x6": (96, 48, 12, 6)}.items():
    src = (rng2.random((h, w, 4), dtype=np.float32) ** 2 * 9.0).astype(np.float32)
    ref = orc.resize_downsample(src, ow, oh, "ref")               # the reference's stbir_resize_float
    rz[name] = {"w": w, "h": h, "ow": ow, "oh": oh, "source_f32_sha256": hashlib.sha256(src.tobytes()).hexdigest(),
                "resized_f32_sha256": hashlib.sha256(ref.tobytes()).hexdigest(),
                "resized_first_texels_hex": [f"{x:08x}" for x in ref.reshape(-1).view(np.uint32)[:12]]}
json.dump(rz, open(os.path.join(HERE, "resize_golden.json"), "w"), indent=1)
print("wrote", list(rz))
