// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// vqimage_ref_shim.cpp — C entry points over the REFERENCE'S OWN Image class: /root/reference/Libs/VQUtils/Source/Image.cpp
// is compiled unmodified and in place (oracle/Makefile; only Log.h / utils.h are shadowed by vqutils_shim/, and `__int64`
// is spelled `long long`), so Image::LoadFromFile, ::CreateResizedImage, ::SaveToDisk, ::CalculateMipLevelCount and the
// MaxLuminance it stores are the engine's code, not a restatement. Output: oracle/_ref/libvqimageref.so.
#include "Image.h"
#include <cstring>

extern "C" {
// Image::LoadFromFile(path): returns 1 and fills w/h/max_luminance; rgba may be null (size query)
int vqimg_load(const char* path, int* w, int* h, float* max_luminance, float* rgba) {
    Image img = Image::LoadFromFile(path);
    if (!img.pData) return 0;
    *w = img.Width; *h = img.Height; *max_luminance = img.MaxLuminance;
    if (rgba && img.IsHDR()) std::memcpy(rgba, img.pData, (size_t)img.Width * img.Height * 16);
    const int hdr = img.IsHDR() ? 1 : 0;
    img.Destroy();
    return 1 + hdr;      // 2 = HDR (RGBA32F), 1 = SDR
}
// Image::CreateResizedImage on RGBA32F texels
int vqimg_resize(const float* rgba, int w, int h, float* out, int ow, int oh) {
    Image src; src.Width = w; src.Height = h; src.BytesPerPixel = 16; src.pData = const_cast<float*>(rgba);
    Image dst = Image::CreateResizedImage(src, (unsigned)ow, (unsigned)oh);
    if (!dst.pData) return 0;
    std::memcpy(out, dst.pData, (size_t)ow * oh * 16);
    dst.Destroy();
    return 1;
}
// Image::SaveToDisk(path) for RGBA32F texels (.hdr)
int vqimg_save(const char* path, const float* rgba, int w, int h) {
    Image img; img.Width = w; img.Height = h; img.BytesPerPixel = 16; img.pData = const_cast<float*>(rgba);
    return img.SaveToDisk(path) ? 1 : 0;
}
int vqimg_mip_level_count(unsigned long long w, unsigned long long h) { return (int)Image::CalculateMipLevelCount(w, h); }
}
