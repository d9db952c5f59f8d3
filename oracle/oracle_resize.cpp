// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// oracle_resize.cpp — SURVEY §8(f).2, the CPU downsize of an HDRI: Image::CreateResizedImage
// (Libs/VQUtils/Source/Image.cpp:148-190) -> stbir_resize_float(in, w, h, 0, out, W, H, 0, 4), reached from
// CreateEnvironmentMapTextureFromHiResAndSaveToDisk (Source/Engine/EnvironmentMap.cpp:142-209: 8k -> 4k/2k/1k).
// Restated from the vendored stb_image_resize.h v0.96 (Libs/VQUtils/Libs/stb): simple API = float, linear colour space,
// no alpha premultiplication, STBIR_EDGE_CLAMP, default filters; for a downsize both axes take the Mitchell-Netravali
// kernel (STBIR_DEFAULT_FILTER_DOWNSAMPLE, :433) in the "downsample" formulation: every INPUT sample is scattered to the
// output samples it influences (:1024-1193), horizontally per scanline (:1525-1640), then vertically (:1984-2201).
// Same scatter order here, so the sums round identically. PINNED bit-for-bit against oracle/_ref/libstbref.so.
#include "oracle.h"
#include <cmath>
#include <vector>

namespace orc {

namespace {
float filter_mitchell(float x) {                       // stbir__filter_mitchell (:825-837), B = C = 1/3
    x = (float)std::fabs(x);
    if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
    if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
    return 0.0f;
}

struct Contrib { int n0, n1; };
struct Axis {                                          // one axis of the transform, downsample formulation
    int in_size, out_size, margin, width, num;         // width = coefficients per contributor (4), num = in_size + 2*margin
    float scale;
    std::vector<Contrib> c; std::vector<float> k;      // k[j*width + (i - c[j].n0)]
};

Axis make_axis(int in_size, int out_size) {
    Axis a;
    a.in_size = in_size; a.out_size = out_size;
    a.scale = ((float)out_size / in_size) / 1.0f;                                       // stbir__calculate_transform (:2212-2235), s1-s0 = 1
    const float support = 2.0f;                                                         // stbir__support_two
    a.margin = (int)std::ceil(support * 2 / a.scale) / 2;                               // stbir__get_filter_pixel_margin (:893-896)
    a.width = (int)std::ceil(support * 2);                                              // stbir__get_coefficient_width (:901-907)
    a.num = in_size + a.margin * 2;                                                     // stbir__get_contributors (:909-915)
    a.c.assign(a.num, {0, -1});
    a.k.assign((size_t)a.num * a.width, 0.0f);
    const float in_pixels_radius = support / a.scale;                                   // stbir__calculate_filters (:1221)
    for (int n = 0; n < a.num; ++n) {
        const int n_adjusted = n - a.margin;
        // stbir__calculate_sample_range_downsample (:1024-1036), shift = 0
        const float in_pixel_center = (float)n_adjusted + 0.5f;
        const float lo = (in_pixel_center - in_pixels_radius) * a.scale - 0.0f;
        const float hi = (in_pixel_center + in_pixels_radius) * a.scale - 0.0f;
        const float out_center_of_in = in_pixel_center * a.scale - 0.0f;
        const int first = (int)std::floor(lo + 0.5), last = (int)std::floor(hi - 0.5);
        // stbir__calculate_coefficients_downsample (:1088-1116)
        Contrib& ct = a.c[n];
        float* kg = &a.k[(size_t)n * a.width];
        ct.n0 = first; ct.n1 = last;
        for (int i = 0; i <= last - first; ++i) {
            const float out_pixel_center = (float)(i + first) + 0.5f;
            kg[i] = filter_mitchell(out_pixel_center - out_center_of_in) * a.scale;
        }
        for (int i = last - first; i >= 0; --i) {
            if (kg[i]) break;
            ct.n1 = ct.n0 + i - 1;
        }
    }
    // stbir__normalize_downsample_coefficients (:1118-1193)
    for (int i = 0; i < out_size; ++i) {
        float total = 0;
        for (int j = 0; j < a.num; ++j) {
            if (i >= a.c[j].n0 && i <= a.c[j].n1) total += a.k[(size_t)j * a.width + (i - a.c[j].n0)];
            else if (i < a.c[j].n0) break;
        }
        const float scale = 1 / total;
        for (int j = 0; j < a.num; ++j) {
            if (i >= a.c[j].n0 && i <= a.c[j].n1) a.k[(size_t)j * a.width + (i - a.c[j].n0)] *= scale;
            else if (i < a.c[j].n0) break;
        }
    }
    for (int j = 0; j < a.num; ++j) {
        float* kg = &a.k[(size_t)j * a.width];
        int skip = 0;
        while (skip < a.width && kg[skip] == 0) skip++;       // (stb reads on past the group when it is all zero; bounded here)
        a.c[j].n0 += skip;
        while (a.c[j].n0 < 0) { a.c[j].n0++; skip++; }
        const int range = a.c[j].n1 - a.c[j].n0 + 1;
        const int mx = std::min(a.width, range);
        for (int i = 0; i < mx; ++i) {
            if (i + skip >= a.width) break;
            kg[i] = kg[i + skip];
        }
    }
    for (int j = 0; j < a.num; ++j) a.c[j].n1 = std::min(a.c[j].n1, out_size - 1);
    return a;
}
inline int edge_clamp(int n, int max) { return n < 0 ? 0 : (n >= max ? max - 1 : n); }   // stbir__edge_wrap, STBIR_EDGE_CLAMP
}  // namespace

// stbir_resize_float(in, w, h, 0, out, ow, oh, 0, 4) for ow <= w and oh <= h (the engine's downsize). Returns 0 on success.
int ResizeFloat4_Downsample(const float* in, int w, int h, float* out, int ow, int oh) {
    if (!in || !out || w <= 0 || h <= 0 || ow <= 0 || oh <= 0 || ow > w || oh > h) return 1;
    const Axis H = make_axis(w, ow), V = make_axis(h, oh);
    std::vector<float> acc((size_t)ow * oh * 4, 0.0f);                    // the ring buffer rows, all of them
    std::vector<float> hbuf((size_t)ow * 4);
    const float v_radius = 2.0f / V.scale;
    for (int y = -V.margin; y < h + V.margin; ++y) {                      // stbir__buffer_loop_downsample (:2162-2201)
        const float c = (float)y + 0.5f;
        const int of = (int)std::floor((c - v_radius) * V.scale - 0.0f + 0.5), ol = (int)std::floor((c + v_radius) * V.scale - 0.0f - 0.5);
        if (ol < 0 || of >= oh) continue;
        // decode (edge clamp in both directions, :1262-1400) + stbir__resample_horizontal_downsample (:1525-1640)
        const float* row = in + (size_t)edge_clamp(y, h) * w * 4;
        std::fill(hbuf.begin(), hbuf.end(), 0.0f);
        for (int x = 0; x < H.num; ++x) {
            const int in_x = edge_clamp(x - H.margin, w);
            const Contrib ct = H.c[x];
            for (int k = ct.n0; k <= ct.n1; ++k) {
                const float coefficient = H.k[(size_t)x * H.width + (k - ct.n0)];
                for (int ch = 0; ch < 4; ++ch) hbuf[(size_t)k * 4 + ch] += row[(size_t)in_x * 4 + ch] * coefficient;
            }
        }
        // stbir__resample_vertical_downsample (:1984-2070)
        const Contrib cv = V.c[y + V.margin];
        for (int k = cv.n0; k <= cv.n1; ++k) {
            const float coefficient = V.k[(size_t)(y + V.margin) * V.width + (k - cv.n0)];
            float* dst = &acc[(size_t)k * ow * 4];
            for (int i = 0; i < ow * 4; ++i) dst[i] += hbuf[i] * coefficient;
        }
    }
    for (size_t i = 0; i < acc.size(); ++i) out[i] = acc[i];             // float output: the encode step is a copy
    return 0;
}

}  // namespace orc
