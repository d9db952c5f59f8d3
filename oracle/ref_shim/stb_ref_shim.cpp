// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// stb_ref_shim.cpp — compiles the reference's OWN vendored stb_image.h / stb_image_write.h
// (/root/reference/Libs/VQUtils/Libs/stb, included in place, nothing copied) and exposes the two calls
// Image::LoadFromFile / Image::SaveToDisk make for .hdr files (Libs/VQUtils/Source/Image.cpp:119-121, 210-213)
// over memory buffers, so that the oracle's restatement can be pinned bit-for-bit / byte-for-byte.
#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_HDR
#define STBI_NO_STDIO
#include "stb/stb_image.h"
#include <cstdio>
#define STB_IMAGE_WRITE_IMPLEMENTATION
#define STBI_WRITE_NO_STDIO
#include "stb/stb_image_write.h"
#define STB_IMAGE_RESIZE_IMPLEMENTATION
#include "stb/stb_image_resize.h"
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

// stbi_loadf(path, &x, &y, &n, 4) on a memory image of the file; returns 0 on failure. rgba may be null (size query).
int stbref_loadf(const uint8_t* file, int n, int* w, int* h, float* rgba) {
    int comp = 0;
    float* p = stbi_loadf_from_memory(file, n, w, h, &comp, 4);
    if (!p) return 0;
    if (rgba) std::memcpy(rgba, p, (size_t)*w * (size_t)*h * 16);
    stbi_image_free(p);
    return 1;
}

static void sink(void* ctx, void* data, int size) {
    auto* v = static_cast<std::vector<uint8_t>*>(ctx);
    v->insert(v->end(), (uint8_t*)data, (uint8_t*)data + size);
}
// stbi_write_hdr(path, x, y, 4, data) into memory; returns the file size
uint64_t stbref_write_hdr(const float* rgba, int w, int h, uint8_t* file, uint64_t capacity) {
    std::vector<uint8_t> v;
    if (!stbi_write_hdr_to_func(sink, &v, w, h, 4, rgba)) return 0;
    if (file) std::memcpy(file, v.data(), v.size() < capacity ? v.size() : capacity);
    return v.size();
}

// Image::CreateResizedImage (Image.cpp:148-190): stbir_resize_float(in, w, h, 0, out, ow, oh, 0, 4); returns stb's rc (1 = ok)
int stbref_resize_float(const float* in, int w, int h, float* out, int ow, int oh) {
    return stbir_resize_float(in, w, h, 0, out, ow, oh, 0, 4);
}

}  // extern "C"
