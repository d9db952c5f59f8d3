// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// mip_ref_shim.cpp — C entry point over the REFERENCE'S OWN VQ_DXGI_UTILS::MipImage (Source/Renderer/Resources/DXGIUtils.cpp:
// 250-317, compiled unmodified and in place; only <dxgiformat.h> and Engine/GPUMarker.h are generated stand-ins): the CPU
// mip step TextureManager.cpp:714-727 runs for every level of material textures (RGBA8 box) and HDRIs (RGBA32F MIN).
// Output: oracle/_ref/libvqmipref.so. Note the reference indexes (x+1, y+1) unconditionally: only even sizes are defined.
#include "Renderer/Resources/DXGIUtils.h"

extern "C" void vqmip_image(const void* src, void* dst, unsigned width, unsigned height, unsigned bytes_per_pixel) {
    VQ_DXGI_UTILS::MipImage(src, dst, width, height, bytes_per_pixel);
}
