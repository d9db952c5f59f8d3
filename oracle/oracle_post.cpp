// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// oracle_post.cpp — scalar restatement of the post chain: Tonemapper.hlsl / HDR.hlsl /
// GaussianBlur.hlsl and the FidelityFX CAS / FSR1 (EASU, RCAS) / SPD math the engine's wrappers
// (Shaders/AMDFidelityFX.hlsl) select: fp32 paths, FFXCAS_NO_UPSCALING, no CAS_SLOW, no
// FSR_RCAS_DENOISE, SPD reduction = average, LDS (non-wave) ordering.
// The constant-setup functions are pinned bit-for-bit against the reference's own A_CPU build
// (oracle/_ref/libffxref.so).
#include "oracle.h"

namespace orc {

// ------------------------------------------------------------------------------------------------
// Tonemapper.hlsl / HDR.hlsl
// ------------------------------------------------------------------------------------------------
float3 Tonemap_Reinhard(float3 c) { return c / (c + splat3(1.0f)); }      // Tonemapper.hlsl:24-27

static inline float sel(bool c, float a, float b) { return c ? a : b; }

float3 LinearToSRGB(float3 c) {                                           // HDR.hlsl:76-80
    const float e = 1.0f / 2.4f;
    return make3(sel(c.x < 0.0031308f, 12.92f * c.x, 1.055f * std::pow(std::fabs(c.x), e) - 0.055f),
                 sel(c.y < 0.0031308f, 12.92f * c.y, 1.055f * std::pow(std::fabs(c.y), e) - 0.055f),
                 sel(c.z < 0.0031308f, 12.92f * c.z, 1.055f * std::pow(std::fabs(c.z), e) - 0.055f));
}
float3 SRGBToLinear(float3 c) {                                           // HDR.hlsl:82-86
    return make3(sel(c.x < 0.04045f, c.x / 12.92f, std::pow(std::fabs(c.x + 0.055f) / 1.055f, 2.4f)),
                 sel(c.y < 0.04045f, c.y / 12.92f, std::pow(std::fabs(c.y + 0.055f) / 1.055f, 2.4f)),
                 sel(c.z < 0.04045f, c.z / 12.92f, std::pow(std::fabs(c.z + 0.055f) / 1.055f, 2.4f)));
}
static inline float3 mul33(const float m[9], float3 v) {   // mul(matrix, column-vector): rows dot v
    return make3(m[0] * v.x + m[1] * v.y + m[2] * v.z,
                 m[3] * v.x + m[4] * v.y + m[5] * v.z,
                 m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
float3 Rec709ToRec2020(float3 c) {                                        // HDR.hlsl:88-97
    static const float m[9] = {0.627402f, 0.329292f, 0.043306f,
                               0.069095f, 0.919544f, 0.011360f,
                               0.016394f, 0.088028f, 0.895578f};
    return mul33(m, c);
}
float3 Rec2020ToRec709(float3 c) {                                        // HDR.hlsl:99-108
    static const float m[9] = {1.660496f, -0.587656f, -0.072840f,
                               -0.124547f, 1.132895f, -0.008348f,
                               -0.018154f, -0.100597f, 1.118751f};
    return mul33(m, c);
}
float3 LinearToST2084(float3 color) {                                     // HDR.hlsl:110-119
    const float m1 = 2610.0f / 4096.0f / 4;
    const float m2 = 2523.0f / 4096.0f * 128;
    const float c1 = 3424.0f / 4096.0f;
    const float c2 = 2413.0f / 4096.0f * 32;
    const float c3 = 2392.0f / 4096.0f * 32;
    const float3 cp = pow3(abs3(color), m1);
    const float3 num = splat3(c1) + cp * c2;
    const float3 den = splat3(1.0f) + cp * c3;
    return pow3(num / den, m2);
}

float4 Tonemapper_CSMain(const VqTonemapperParams& p, float4 InRGBA) {   // Tonemapper.hlsl:110-151
    float3 OutRGB = splat3(0.0f);
    switch (p.OutputDisplayCurve) {
        case VQ_DISPLAY_CURVE_SRGB:
            OutRGB = Tonemap_Reinhard(xyz(InRGBA));
            if (p.ToggleGammaCorrection) OutRGB = LinearToSRGB(OutRGB);
            break;
        case VQ_DISPLAY_CURVE_ST2084: {
            const float ST2084_MAX = 10000.0f;                            // HDR.hlsl:43
            const float HDR_Scalar = p.DisplayReferenceBrightnessLevel / ST2084_MAX;
            OutRGB = xyz(InRGBA);
            if (p.ContentColorSpace == VQ_COLOR_SPACE_REC_709) OutRGB = Rec709ToRec2020(OutRGB);
            OutRGB = LinearToST2084(OutRGB * HDR_Scalar);
        } break;
        case VQ_DISPLAY_CURVE_LINEAR:
            OutRGB = xyz(InRGBA);
            break;
        default:
            OutRGB = make3(1, 1, 0);
            break;
    }
    return make4(OutRGB, InRGBA.w);
}

// ------------------------------------------------------------------------------------------------
// GaussianBlur.hlsl, KERNEL_DIMENSION 21 (the default; the PSOs pass no macro)
// ------------------------------------------------------------------------------------------------
float4 GaussianBlur_CSMain(const Image& in, int x, int y, bool vertical, int sizeX, int sizeY) { // :119-186
    static const float KERNEL_WEIGHTS[11] = {0.224716f, 0.191756f, 0.119146f, 0.053897f, 0.017746f, 0.004252f,
                                             0.000741f, 0.000094f, 0.000009f, 0.000001f, 0.0f};  // :110
    const int KERNEL_DIMENSION = 21, KERNEL_RANGE_MINUS1 = 10;
    float3 OutRGB = splat3(0.0f);
    for (int kernelIt = 0; kernelIt < KERNEL_DIMENSION; ++kernelIt) {
        const int kernelOffset = kernelIt - KERNEL_RANGE_MINUS1;
        const int kernelIndex = std::abs(kernelOffset);
        int sx = x, sy = y;
        if (vertical) sy = std::min(std::max(y + kernelOffset, 0), sizeY - 1);
        else          sx = std::min(std::max(x + kernelOffset, 0), sizeX - 1);
        const float* t = in.at(sx, sy);
        OutRGB += make3(t[0], t[1], t[2]) * KERNEL_WEIGHTS[kernelIndex];
    }
    return make4(OutRGB, 1.0f);
}

// ------------------------------------------------------------------------------------------------
// FidelityFX scalar helpers (FSR1.0/ffx_a.h; CPU definitions at :283-365, bit hacks at :1842-1845)
// ------------------------------------------------------------------------------------------------
float APrxLoSqrtF1(float a) { return asfloat((asuint(a) >> 1u) + 0x1fbc4639u); }
float APrxLoRcpF1(float a)  { return asfloat(0x7ef07ebbu - asuint(a)); }
float APrxMedRcpF1(float a) { const float b = asfloat(0x7ef19fffu - asuint(a)); return b * (-b * a + 2.0f); }
float APrxLoRsqF1(float a)  { return asfloat(0x5f347d74u - (asuint(a) >> 1u)); }

static inline float ARcpF1(float a) { return 1.0f / a; }
static inline float ASatF1(float a) { return std::fmin(1.0f, std::fmax(0.0f, a)); }
static inline float ALerpF1(float a, float b, float c) { return b * c + (-a * c + a); }   // ffx_a.h:298
static inline float AMin3F1(float x, float y, float z) { return std::fmin(x, std::fmin(y, z)); }
static inline float AMax3F1(float x, float y, float z) { return std::fmax(x, std::fmax(y, z)); }

// AU1_AH1_AF1 (ffx_a.h:482-550): truncating float -> half with denormals, +-INF/NaN -> +-65504.
// The reference stores this as two 512-entry tables indexed by sign|exponent; the tables are this rule:
static uint32_t AU1_AH1_AF1(float f) {
    const uint32_t u = asuint(f);
    const uint32_t i = u >> 23, e = i & 0xffu, sign = (i >> 8) << 15;
    uint32_t base, shift;
    if (e < 103u)       { base = 0u;                       shift = 24u; }
    else if (e < 113u)  { base = 0x0400u >> (113u - e);    shift = 126u - e; }
    else if (e <= 142u) { base = (e - 112u) << 10;         shift = 13u; }
    else                { base = 0x7bffu;                  shift = 24u; }
    return (base | sign) + ((u & 0x7fffffu) >> shift);
}
static uint32_t AU1_AH2_AF2(float a0, float a1) { return AU1_AH1_AF1(a0) + (AU1_AH1_AF1(a1) << 16); }

void CasSetup(uint32_t const0[4], uint32_t const1[4], float sharpness,
              float inX, float inY, float outX, float outY) {             // ffx_cas.h:375-394
    const0[0] = asuint(inX * ARcpF1(outX));
    const0[1] = asuint(inY * ARcpF1(outY));
    const0[2] = asuint(0.5f * inX * ARcpF1(outX) - 0.5f);
    const0[3] = asuint(0.5f * inY * ARcpF1(outY) - 0.5f);
    const float sharp = -ARcpF1(ALerpF1(8.0f, 5.0f, ASatF1(sharpness)));
    const1[0] = asuint(sharp);
    const1[1] = AU1_AH2_AF2(sharp, 0.0f);
    const1[2] = asuint(8.0f * inX * ARcpF1(outX));
    const1[3] = 0;
}

void FsrEasuCon(uint32_t con0[4], uint32_t con1[4], uint32_t con2[4], uint32_t con3[4],
                float inVpX, float inVpY, float inSzX, float inSzY, float outX, float outY) {  // ffx_fsr1.h:156-202
    con0[0] = asuint(inVpX * ARcpF1(outX));
    con0[1] = asuint(inVpY * ARcpF1(outY));
    con0[2] = asuint(0.5f * inVpX * ARcpF1(outX) - 0.5f);
    con0[3] = asuint(0.5f * inVpY * ARcpF1(outY) - 0.5f);
    con1[0] = asuint(ARcpF1(inSzX));
    con1[1] = asuint(ARcpF1(inSzY));
    con1[2] = asuint(1.0f * ARcpF1(inSzX));
    con1[3] = asuint(-1.0f * ARcpF1(inSzY));
    con2[0] = asuint(-1.0f * ARcpF1(inSzX));
    con2[1] = asuint(2.0f * ARcpF1(inSzY));
    con2[2] = asuint(1.0f * ARcpF1(inSzX));
    con2[3] = asuint(2.0f * ARcpF1(inSzY));
    con3[0] = asuint(0.0f * ARcpF1(inSzX));
    con3[1] = asuint(4.0f * ARcpF1(inSzY));
    con3[2] = con3[3] = 0;
}

void FsrRcasCon(uint32_t con[4], float sharpness) {                       // ffx_fsr1.h:662-672
    sharpness = std::exp2(-sharpness);
    con[0] = asuint(sharpness);
    con[1] = AU1_AH2_AF2(sharpness, sharpness);
    con[2] = 0;
    con[3] = 0;
}

void SpdSetup(uint32_t dispatchXY[2], uint32_t workGroupOffset[2], uint32_t numWorkGroupsAndMips[2],
              const uint32_t rectInfo[4], int mips) {                     // ffx_spd.h:327-351
    workGroupOffset[0] = rectInfo[0] / 64;
    workGroupOffset[1] = rectInfo[1] / 64;
    const uint32_t endIndexX = (rectInfo[0] + rectInfo[2] - 1) / 64;
    const uint32_t endIndexY = (rectInfo[1] + rectInfo[3] - 1) / 64;
    dispatchXY[0] = endIndexX + 1 - workGroupOffset[0];
    dispatchXY[1] = endIndexY + 1 - workGroupOffset[1];
    numWorkGroupsAndMips[0] = dispatchXY[0] * dispatchXY[1];
    if (mips >= 0) {
        numWorkGroupsAndMips[1] = (uint32_t)mips;
    } else {
        const uint32_t resolution = std::max(rectInfo[2], rectInfo[3]);
        numWorkGroupsAndMips[1] = (uint32_t)std::fmin(std::floor(std::log2((float)resolution)), 12.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// image loads with D3D semantics
// ------------------------------------------------------------------------------------------------
static inline float3 LoadZeroBorder(const Image& in, int x, int y) {      // Texture.Load out of range -> 0
    if (x < 0 || y < 0 || x >= in.width || y >= in.height) return splat3(0.0f);
    const float* t = in.at(x, y);
    return make3(t[0], t[1], t[2]);
}
static inline float3 LoadAddressed(const Image& in, int x, int y, int addressMode) {
    if (addressMode == 0) {   // WRAP
        x %= in.width;  if (x < 0) x += in.width;
        y %= in.height; if (y < 0) y += in.height;
    } else {                  // CLAMP
        x = std::min(std::max(x, 0), in.width - 1);
        y = std::min(std::max(y, 0), in.height - 1);
    }
    const float* t = in.at(x, y);
    return make3(t[0], t[1], t[2]);
}

// ------------------------------------------------------------------------------------------------
// CAS, no-scaling branch (CAS/ffx_cas.h:408-537), CasInput = identity (AMDFidelityFX.hlsl:108)
// ------------------------------------------------------------------------------------------------
float3 CasFilter_NoScaling(const Image& in, int x, int y, const uint32_t const1[4]) {
    const float3 b = LoadZeroBorder(in, x, y - 1);
    const float3 d = LoadZeroBorder(in, x - 1, y);
    const float3 e = LoadZeroBorder(in, x, y);
    const float3 f = LoadZeroBorder(in, x + 1, y);
    const float3 h = LoadZeroBorder(in, x, y + 1);
    // (a, c, g, i are loaded by the reference but only used under CAS_BETTER_DIAGONALS)
    const float mnR = AMin3F1(AMin3F1(d.x, e.x, f.x), b.x, h.x);
    const float mnG = AMin3F1(AMin3F1(d.y, e.y, f.y), b.y, h.y);
    const float mnB = AMin3F1(AMin3F1(d.z, e.z, f.z), b.z, h.z);
    const float mxR = AMax3F1(AMax3F1(d.x, e.x, f.x), b.x, h.x);
    const float mxG = AMax3F1(AMax3F1(d.y, e.y, f.y), b.y, h.y);
    const float mxB = AMax3F1(AMax3F1(d.z, e.z, f.z), b.z, h.z);
    const float rcpMR = APrxLoRcpF1(mxR);
    const float rcpMG = APrxLoRcpF1(mxG);
    const float rcpMB = APrxLoRcpF1(mxB);
    float ampR = ASatF1(std::fmin(mnR, 1.0f - mxR) * rcpMR);
    float ampG = ASatF1(std::fmin(mnG, 1.0f - mxG) * rcpMG);
    float ampB = ASatF1(std::fmin(mnB, 1.0f - mxB) * rcpMB);
    ampR = APrxLoSqrtF1(ampR);
    ampG = APrxLoSqrtF1(ampG);
    ampB = APrxLoSqrtF1(ampB);
    (void)ampR; (void)ampB;
    const float peak = asfloat(const1[0]);
    const float wG = ampG * peak;
    const float rcpWeight = APrxMedRcpF1(1.0f + 4.0f * wG);
    float3 pix;
    pix.x = ASatF1((b.x * wG + d.x * wG + f.x * wG + h.x * wG + e.x) * rcpWeight);
    pix.y = ASatF1((b.y * wG + d.y * wG + f.y * wG + h.y * wG + e.y) * rcpWeight);
    pix.z = ASatF1((b.z * wG + d.z * wG + f.z * wG + h.z * wG + e.z) * rcpWeight);
    return pix;
}

// ------------------------------------------------------------------------------------------------
// FSR1 EASU (FSR1.0/ffx_fsr1.h:239-437). The four Gather4 calls reduce to 12 integer texel
// fetches around (fx,fy) = floor(ip*con0.xy + con0.zw) (SURVEY.md §9 "EASU gathers").
// ------------------------------------------------------------------------------------------------
static void FsrEasuTapF(float3& aC, float& aW, float2 off, float2 dir, float2 len, float lob, float clp, float3 c) { // :239-272
    float2 v;
    v.x = (off.x * (dir.x)) + (off.y * dir.y);
    v.y = (off.x * (-dir.y)) + (off.y * dir.x);
    v = v * len;
    float d2 = v.x * v.x + v.y * v.y;
    d2 = std::fmin(d2, clp);
    float wB = (2.0f / 5.0f) * d2 + (-1.0f);
    float wA = lob * d2 + (-1.0f);
    wB *= wB;
    wA *= wA;
    wB = (25.0f / 16.0f) * wB + (-(25.0f / 16.0f - 1.0f));
    const float w = wB * wA;
    aC += c * w; aW += w;
}
static void FsrEasuSetF(float2& dir, float& len, float2 pp, bool biS, bool biT, bool biU, bool biV,
                        float lA, float lB, float lC, float lD, float lE) {                        // :275-313
    float w = 0.0f;
    if (biS) w = (1.0f - pp.x) * (1.0f - pp.y);
    if (biT) w = pp.x * (1.0f - pp.y);
    if (biU) w = (1.0f - pp.x) * pp.y;
    if (biV) w = pp.x * pp.y;
    const float dc = lD - lC;
    const float cb = lC - lB;
    float lenX = std::fmax(std::fabs(dc), std::fabs(cb));
    lenX = APrxLoRcpF1(lenX);
    const float dirX = lD - lB;
    dir.x += dirX * w;
    lenX = ASatF1(std::fabs(dirX) * lenX);
    lenX *= lenX;
    len += lenX * w;
    const float ec = lE - lC;
    const float ca = lC - lA;
    float lenY = std::fmax(std::fabs(ec), std::fabs(ca));
    lenY = APrxLoRcpF1(lenY);
    const float dirY = lE - lA;
    dir.y += dirY * w;
    lenY = ASatF1(std::fabs(dirY) * lenY);
    lenY *= lenY;
    len += lenY * w;
}
static inline float Luma2(float3 c) { return c.z * 0.5f + (c.x * 0.5f + c.y); }   // :356-359

float3 FsrEasuF(const Image& in, int ipx, int ipy, const uint32_t con[16], int addressMode) {      // :315-437
    float2 pp = make2((float)ipx * asfloat(con[0]) + asfloat(con[2]),
                      (float)ipy * asfloat(con[1]) + asfloat(con[3]));
    const float2 fp = make2(std::floor(pp.x), std::floor(pp.y));
    pp = pp - fp;
    const int fx = (int)fp.x, fy = (int)fp.y;
    //    b c
    //  e f g h
    //  i j k l
    //    n o
    const float3 b = LoadAddressed(in, fx, fy - 1, addressMode), c = LoadAddressed(in, fx + 1, fy - 1, addressMode);
    const float3 e = LoadAddressed(in, fx - 1, fy, addressMode), f = LoadAddressed(in, fx, fy, addressMode);
    const float3 g = LoadAddressed(in, fx + 1, fy, addressMode), h = LoadAddressed(in, fx + 2, fy, addressMode);
    const float3 i = LoadAddressed(in, fx - 1, fy + 1, addressMode), j = LoadAddressed(in, fx, fy + 1, addressMode);
    const float3 k = LoadAddressed(in, fx + 1, fy + 1, addressMode), l = LoadAddressed(in, fx + 2, fy + 1, addressMode);
    const float3 n = LoadAddressed(in, fx, fy + 2, addressMode), o = LoadAddressed(in, fx + 1, fy + 2, addressMode);
    const float bL = Luma2(b), cL = Luma2(c), eL = Luma2(e), fL = Luma2(f), gL = Luma2(g), hL = Luma2(h);
    const float iL = Luma2(i), jL = Luma2(j), kL = Luma2(k), lL = Luma2(l), nL = Luma2(n), oL = Luma2(o);
    float2 dir = make2(0.0f, 0.0f);
    float len = 0.0f;
    FsrEasuSetF(dir, len, pp, true, false, false, false, bL, eL, fL, gL, jL);
    FsrEasuSetF(dir, len, pp, false, true, false, false, cL, fL, gL, hL, kL);
    FsrEasuSetF(dir, len, pp, false, false, true, false, fL, iL, jL, kL, nL);
    FsrEasuSetF(dir, len, pp, false, false, false, true, gL, jL, kL, lL, oL);
    const float2 dir2 = dir * dir;
    float dirR = dir2.x + dir2.y;
    const bool zro = dirR < (1.0f / 32768.0f);
    dirR = APrxLoRsqF1(dirR);
    dirR = zro ? 1.0f : dirR;
    dir.x = zro ? 1.0f : dir.x;
    dir = dir * dirR;
    len = len * 0.5f;
    len *= len;
    const float stretch = (dir.x * dir.x + dir.y * dir.y) * APrxLoRcpF1(std::fmax(std::fabs(dir.x), std::fabs(dir.y)));
    const float2 len2 = make2(1.0f + (stretch - 1.0f) * len, 1.0f + (-0.5f) * len);
    const float lob = 0.5f + ((1.0f / 4.0f - 0.04f) - 0.5f) * len;
    const float clp = APrxLoRcpF1(lob);
    const float3 min4 = min3(min3(f, min3(g, j)), k);    // min(AMin3F3(f,g,j),k) with AMin3(x,y,z)=min(x,min(y,z))
    const float3 max4 = max3(max3(f, max3(g, j)), k);
    float3 aC = splat3(0.0f);
    float aW = 0.0f;
    FsrEasuTapF(aC, aW, make2(0.0f, -1.0f) - pp, dir, len2, lob, clp, b);
    FsrEasuTapF(aC, aW, make2(1.0f, -1.0f) - pp, dir, len2, lob, clp, c);
    FsrEasuTapF(aC, aW, make2(-1.0f, 1.0f) - pp, dir, len2, lob, clp, i);
    FsrEasuTapF(aC, aW, make2(0.0f, 1.0f) - pp, dir, len2, lob, clp, j);
    FsrEasuTapF(aC, aW, make2(0.0f, 0.0f) - pp, dir, len2, lob, clp, f);
    FsrEasuTapF(aC, aW, make2(-1.0f, 0.0f) - pp, dir, len2, lob, clp, e);
    FsrEasuTapF(aC, aW, make2(1.0f, 1.0f) - pp, dir, len2, lob, clp, k);
    FsrEasuTapF(aC, aW, make2(2.0f, 1.0f) - pp, dir, len2, lob, clp, l);
    FsrEasuTapF(aC, aW, make2(2.0f, 0.0f) - pp, dir, len2, lob, clp, h);
    FsrEasuTapF(aC, aW, make2(1.0f, 0.0f) - pp, dir, len2, lob, clp, g);
    FsrEasuTapF(aC, aW, make2(1.0f, 2.0f) - pp, dir, len2, lob, clp, o);
    FsrEasuTapF(aC, aW, make2(0.0f, 2.0f) - pp, dir, len2, lob, clp, n);
    return min3(max4, max3(min4, aC * ARcpF1(aW)));
}

// ------------------------------------------------------------------------------------------------
// FSR1 RCAS (ffx_fsr1.h:684-769), FsrRcasInputF = identity, no denoise, no alpha passthrough
// ------------------------------------------------------------------------------------------------
float3 FsrRcasF(const Image& in, int x, int y, const uint32_t con[4]) {
    const float FSR_RCAS_LIMIT = 0.25f - (1.0f / 16.0f);                  // :654
    const float3 b = LoadZeroBorder(in, x, y - 1);
    const float3 d = LoadZeroBorder(in, x - 1, y);
    const float3 e = LoadZeroBorder(in, x, y);
    const float3 f = LoadZeroBorder(in, x + 1, y);
    const float3 h = LoadZeroBorder(in, x, y + 1);
    const float mn4R = std::fmin(AMin3F1(b.x, d.x, f.x), h.x);
    const float mn4G = std::fmin(AMin3F1(b.y, d.y, f.y), h.y);
    const float mn4B = std::fmin(AMin3F1(b.z, d.z, f.z), h.z);
    const float mx4R = std::fmax(AMax3F1(b.x, d.x, f.x), h.x);
    const float mx4G = std::fmax(AMax3F1(b.y, d.y, f.y), h.y);
    const float mx4B = std::fmax(AMax3F1(b.z, d.z, f.z), h.z);
    const float2 peakC = make2(1.0f, -1.0f * 4.0f);
    const float hitMinR = mn4R * ARcpF1(4.0f * mx4R);
    const float hitMinG = mn4G * ARcpF1(4.0f * mx4G);
    const float hitMinB = mn4B * ARcpF1(4.0f * mx4B);
    const float hitMaxR = (peakC.x - mx4R) * ARcpF1(4.0f * mn4R + peakC.y);
    const float hitMaxG = (peakC.x - mx4G) * ARcpF1(4.0f * mn4G + peakC.y);
    const float hitMaxB = (peakC.x - mx4B) * ARcpF1(4.0f * mn4B + peakC.y);
    const float lobeR = std::fmax(-hitMinR, hitMaxR);
    const float lobeG = std::fmax(-hitMinG, hitMaxG);
    const float lobeB = std::fmax(-hitMinB, hitMaxB);
    const float lobe = std::fmax(-FSR_RCAS_LIMIT, std::fmin(AMax3F1(lobeR, lobeG, lobeB), 0.0f)) * asfloat(con[0]);
    const float rcpL = APrxMedRcpF1(4.0f * lobe + 1.0f);
    float3 pix;
    pix.x = (lobe * b.x + lobe * d.x + lobe * h.x + lobe * f.x + e.x) * rcpL;
    pix.y = (lobe * b.y + lobe * d.y + lobe * h.y + lobe * f.y + e.y) * rcpL;
    pix.z = (lobe * b.z + lobe * d.z + lobe * h.z + lobe * f.z + e.z) * rcpL;
    return pix;
}

// ------------------------------------------------------------------------------------------------
// SPD (SPD/ffx_spd.h:557-835, LDS path; SpdReduce4 = (v0+v1+v2+v3)*0.25, AMDFidelityFX.hlsl:463-466).
// Level L+1 texel (x,y) is the reduction of level L texels (2x..2x+1, 2y..2y+1); the operand order is
//   level 1 (from the source) and level 7 (from mip 6, SpdReduceLoad4):  (x,y),(x,y+1),(x+1,y),(x+1,y+1)
//   every other level (SpdReduceIntermediate / mip 7's v0..v3):          (x,y),(x+1,y),(x,y+1),(x+1,y+1)
// Destination sizes are floor-halved, so no out-of-range (zero) texel is ever averaged in.
// ------------------------------------------------------------------------------------------------
void SpdDownsampleLevel(const Image& src, const MutImage& dst, int dstLevel) {
    const bool columnMajor = (dstLevel == 1 || dstLevel == 7);
    for (int y = 0; y < dst.height; ++y)
        for (int x = 0; x < dst.width; ++x) {
            const float* p00 = src.at(2 * x, 2 * y);
            const float* p10 = src.at(2 * x + 1, 2 * y);
            const float* p01 = src.at(2 * x, 2 * y + 1);
            const float* p11 = src.at(2 * x + 1, 2 * y + 1);
            const float* v0 = p00;
            const float* v1 = columnMajor ? p01 : p10;
            const float* v2 = columnMajor ? p10 : p01;
            const float* v3 = p11;
            float* o = dst.at(x, y);
            for (int ch = 0; ch < 4; ++ch) o[ch] = (v0[ch] + v1[ch] + v2[ch] + v3[ch]) * 0.25f;
        }
}

}  // namespace orc
