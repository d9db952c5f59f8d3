// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// oracle_shading.cpp — scalar restatement of BRDF.hlsl, ShadingMath.hlsl, Lighting.hlsl and the
// lighting half of ForwardLighting.hlsl::PSMain, plus the texture-sampling semantics.
// Compile with -ffp-contract=off. pow(x,2) is written x*x (what FXC/DXC emit for a literal 2);
// every other pow is powf.
#include "oracle.h"
#include <thread>

namespace orc {

// ------------------------------------------------------------------------------------------------
// packed-resource arithmetic
// ------------------------------------------------------------------------------------------------
int mip_level_count(uint64_t w, uint64_t h) {           // Libs/VQUtils/Source/Image.cpp:231-241
    int mips = 0;
    while (w >= 1 && h >= 1) { ++mips; w >>= 1; h >>= 1; }
    return mips;
}
size_t Cubemap::offset(int mip, int face) const {
    size_t o = 0;
    for (int m = 0; m < mip; ++m) { size_t r = (size_t)(res >> m); o += 6 * r * r; }
    size_t r = (size_t)(res >> mip);
    return o + (size_t)face * r * r;
}
size_t cubemap_texel_count(int res, int mips) {
    size_t o = 0;
    for (int m = 0; m < mips; ++m) { size_t r = (size_t)(res >> m); o += 6 * r * r; }
    return o;
}
int cubemap_row_count(int res, int mips) {
    int n = 0;
    for (int m = 0; m < mips; ++m) n += 6 * (res >> m);
    return n;
}
size_t Pyramid::offset(int level) const {
    size_t o = 0;
    for (int l = 0; l < level; ++l) o += (size_t)(width >> l) * (size_t)(height >> l);
    return o;
}
size_t pyramid_texel_count(int w, int h, int levels) {
    size_t o = 0;
    for (int l = 0; l < levels; ++l) o += (size_t)(w >> l) * (size_t)(h >> l);
    return o;
}

// ------------------------------------------------------------------------------------------------
// BRDF.hlsl
// ------------------------------------------------------------------------------------------------
static constexpr float EPSILON = 0.000000000001f;   // BRDF.hlsl:22

float NormalDistributionGGX(float NdotH, float roughness) {          // BRDF.hlsl:65-79
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float nh2 = NdotH * NdotH;
    const float t = nh2 * (a2 - 1.0f) + 1.0f;
    const float denom = PI * (t * t);
    if (denom < EPSILON) return 1.0f;
    return a2 / denom;
}

float Geometry_Smiths_SchlickGGX(float3 N, float3 V, float roughness) {   // BRDF.hlsl:82-97
    const float rp1 = roughness + 1.0f;
    const float k = (rp1 * rp1) / 8.0f;
    const float NV = std::fmax(0.0f, dot(N, V));
    const float denom = (NV * (1.0f - k) + k) + 0.0001f;
    return NV / denom;
}

float Geometry_Smiths_SchlickGGX_EnvironmentMap(float3 N, float3 V, float roughness) {  // BRDF.hlsl:100-115
    const float k = (roughness * roughness) / 2.0f;
    const float NV = std::fmax(0.0f, dot(N, V));
    const float denom = (NV * (1.0f - k) + k) + 0.0001f;
    return NV / denom;
}

float Geometry_Smith(float3 N, float3 V, float3 L, float k) {             // BRDF.hlsl:118-121
    return Geometry_Smiths_SchlickGGX(N, V, k) * Geometry_Smiths_SchlickGGX(N, L, k);
}

float GeometryEnvironmentMap(float3 N, float3 V, float3 L, float k) {     // BRDF.hlsl:124-129
    const float geomNV = Geometry_Smiths_SchlickGGX_EnvironmentMap(N, V, k);
    const float geomNL = Geometry_Smiths_SchlickGGX_EnvironmentMap(N, L, k);
    return geomNV * geomNL;
}

float3 Fresnel_Schlick(float3 N, float3 V, float3 F0) {                   // BRDF.hlsl:132-136
    const float p = std::pow(1.0f - std::fmax(0.0f, dot(N, V)), 5.0f);
    return F0 + (make3(1, 1, 1) - F0) * p;
}

float3 Fresnel_Gaussian(float3 H, float3 V, float3 F0) {                  // BRDF.hlsl:140-147
    const float c0 = -5.55373f;
    const float c1 = -6.98316f;
    const float VdotH = std::fmax(0.0f, dot(V, H));
    return F0 + (make3(1, 1, 1) - F0) * std::pow(2.0f, (c0 * VdotH - c1) * VdotH);
}

float3 FresnelWithRoughness(float cosTheta, float3 F0, float roughness) { // BRDF.hlsl:152-156
    const float p = std::pow(1.0f - cosTheta, 5.0f);
    return F0 + (max3(splat3(1.0f - roughness), F0) - F0) * p;
}

float3 F_LambertDiffuse(float3 kd) { return kd / PI; }                    // BRDF.hlsl:158-161

float3 BRDF(const BRDF_Surface& s, float3 Wi, float3 V) {                 // BRDF.hlsl:163-194
    const float3 Wo = normalize(V);
    const float3 N = normalize(s.N);
    const float3 H = normalize(Wo + Wi);
    const float NdotH = saturate(dot(N, H));
    const float VdotH = saturate(dot(Wo, H)); (void)VdotH;
    const float NdotV = saturate(dot(N, Wo));
    const float NdotL = saturate(dot(N, Wi));

    const float3 albedo = s.diffuseColor;
    const float roughness = s.roughness;
    const float metalness = s.metalness;
    const float3 F0 = lerp(make3(0.04f, 0.04f, 0.04f), albedo, metalness);

    const float3 F = Fresnel_Schlick(H, V, F0);      // note: the un-renormalised V argument (line 181)
    const float G = Geometry_Smith(N, Wo, Wi, roughness);
    const float D = NormalDistributionGGX(NdotH, roughness);
    const float denom = std::fmax(4.0f * NdotV * NdotL, 0.0001f);
    const float3 specular = F * D * G / denom;       // D * F * G / denom, left to right
    const float3 Is = specular;

    const float3 kS = F;
    const float3 kD = (make3(1, 1, 1) - kS) * (1.0f - metalness);
    const float3 Id = F_LambertDiffuse(kD * albedo);
    return Id + Is;
}

float3 EnvironmentBRDF(float NdotV, float roughness, float metallic, float3 diffuseColor,
                       float3 diffuseIrradiance, float3 preFilteredSpecular, float2 F0ScaleBias) { // BRDF.hlsl:196-207
    const float3 F0 = lerp(splat3(0.04f), diffuseColor, metallic);
    const float3 Ks = FresnelWithRoughness(NdotV, F0, roughness);
    const float3 Kd = (splat3(1.0f) - Ks) * (1.0f - metallic);
    const float3 diffuse = diffuseIrradiance * diffuseColor;
    const float3 specular = preFilteredSpecular * (Ks * F0ScaleBias.x + splat3(F0ScaleBias.y));
    return Kd * diffuse + specular;
}

float3 ImportanceSampleGGX(float2 Xi, float3 N, float roughness) {        // BRDF.hlsl:217-238
    const float a = roughness * roughness;
    const float phi = 2.0f * PI * Xi.x;
    const float cosTheta = std::sqrt((1.0f - Xi.y) / (1.0f + (a * a - 1.0f) * Xi.y));
    const float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
    float3 H;
    H.x = std::cos(phi) * sinTheta;
    H.y = std::sin(phi) * sinTheta;
    H.z = cosTheta;
    const float3 up = std::fabs(N.z) < 0.999f ? make3(0, 0, 1) : make3(1, 0, 0);
    const float3 tangent = normalize(cross(up, N));
    const float3 bitangent = cross(N, tangent);
    const float3 sample = tangent * H.x + bitangent * H.y + N * H.z;
    return normalize(sample);
}

float2 IntegrateBRDF(float NdotV, float roughness, int sampleCount) {     // BRDF.hlsl:239-283
    float3 V;
    V.x = std::sqrt(1.0f - NdotV * NdotV);
    V.y = 0;
    V.z = NdotV;
    float F0Scale = 0, F0Bias = 0;
    const float3 N = make3(0, 0, 1);
    for (uint32_t i = 0; i < (uint32_t)sampleCount; ++i) {
        const float2 Xi = Hammersley(i, (uint32_t)sampleCount);
        const float3 H = ImportanceSampleGGX(Xi, N, roughness);
        const float3 L = normalize(reflect(-V, H));
        const float NdotL = std::fmax(L.z, 0.0f);
        const float NdotH = std::fmax(H.z, 0.0f);
        const float VdotH = std::fmax(dot(V, H), 0.0f);
        if (NdotL > 0.0f) {
            const float G = GeometryEnvironmentMap(N, V, L, roughness);
            const float G_Vis = std::fmax((G * VdotH) / (NdotH * NdotV), 0.0001f);
            const float Fc = std::pow(1.0f - VdotH, 5.0f);
            F0Scale += (1.0f - Fc) * G_Vis;
            F0Bias += Fc * G_Vis;
        }
    }
    return make2(F0Scale, F0Bias) / (float)sampleCount;
}

// ------------------------------------------------------------------------------------------------
// ShadingMath.hlsl
// ------------------------------------------------------------------------------------------------
float RadicalInverse_VdC(uint32_t bits) {                                 // ShadingMath.hlsl:87-95
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
float2 Hammersley(uint32_t i, uint32_t count) {                           // ShadingMath.hlsl:119-127
    return make2((float)i / (float)count, RadicalInverse_VdC(i));
}
float2 DirectionToEquirectUV(float3 v) {                                  // ShadingMath.hlsl:70-80
    float2 uv = make2(std::atan2(v.z, v.x), std::asin(-v.y));
    uv = uv / make2(-TWO_PI, PI);
    uv = uv + make2(0.5f, 0.5f);
    return uv;
}
float3 SRGBToLinear_pow22(float3 c) { return pow3(c, 2.2f); }             // ShadingMath.hlsl:65

// ------------------------------------------------------------------------------------------------
// Texture sampling. Decisions — D3D filtering hardware is not in the source:
//   * texel-space coordinate x = u*W - 0.5, fp32 fractional weights, lerp(a,b,t) = a + t*(b-a)
//   * equirect: WRAP in u AND v (RootSignatures.cpp:402), trilinear between floor(lod), floor(lod)+1
//   * cubemap: D3D face selection by major axis; bilinear; taps that fall one texel outside the face
//     are taken from the neighbouring face ("seamless", as D3D10+ hardware does); the one tap that can
//     fall outside in both directions (cube corner) is the mean of the other three taps
//   * LUT: CLAMP
// ------------------------------------------------------------------------------------------------
static inline int wrapi(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }
static inline float4 ld4(const float* p) { return {p[0], p[1], p[2], p[3]}; }

static float4 BilinearWrap(const Pyramid& tex, int level, float2 uv) {
    const int W = tex.w(level), H = tex.h(level);
    const float* base = tex.data + 4 * tex.offset(level);
    const float x = uv.x * (float)W - 0.5f;
    const float y = uv.y * (float)H - 0.5f;
    if (!(std::isfinite(x) && std::isfinite(y))) return make4(0, 0, 0, 0);
    const float x0 = std::floor(x), y0 = std::floor(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = wrapi((int)x0, W), iy0 = wrapi((int)y0, H);
    const int ix1 = wrapi(ix0 + 1, W), iy1 = wrapi(iy0 + 1, H);
    const float4 t00 = ld4(base + 4 * ((size_t)iy0 * W + ix0));
    const float4 t10 = ld4(base + 4 * ((size_t)iy0 * W + ix1));
    const float4 t01 = ld4(base + 4 * ((size_t)iy1 * W + ix0));
    const float4 t11 = ld4(base + 4 * ((size_t)iy1 * W + ix1));
    const float4 top = lerp(t00, t10, fx);
    const float4 bot = lerp(t01, t11, fx);
    return lerp(top, bot, fy);
}

float4 SampleEquirectLevel(const Pyramid& tex, float2 uv, float lod) {
    const float maxLod = (float)(tex.levels - 1);
    lod = std::fmin(std::fmax(lod, 0.0f), maxLod);
    const float l0f = std::floor(lod);
    const int l0 = (int)l0f;
    const float f = lod - l0f;
    const float4 c0 = BilinearWrap(tex, l0, uv);
    if (f == 0.0f || l0 + 1 >= tex.levels) return c0;
    const float4 c1 = BilinearWrap(tex, l0 + 1, uv);
    return lerp(c0, c1, f);
}

// A35: per-texel look direction of a cube face (CubemapUtility.cpp:40-48 view matrices + 90-degree
// projection, EnvironmentMapRendering.cpp:174; VSMain_PerFace CubemapConvolution.hlsl:63-73).
static inline float3 FaceDir(int face, float sx, float sy) {
    switch (face) {
        case 0: return make3(1.0f, sy, -sx);
        case 1: return make3(-1.0f, sy, sx);
        case 2: return make3(sx, 1.0f, -sy);
        case 3: return make3(sx, -1.0f, sy);
        case 4: return make3(sx, sy, 1.0f);
        default: return make3(-sx, sy, -1.0f);
    }
}
float3 CubeTexelDirection(int face, int px, int py, int res) {
    const float sx = 2.0f * ((float)px + 0.5f) / (float)res - 1.0f;
    const float sy = 1.0f - 2.0f * ((float)py + 0.5f) / (float)res;
    return FaceDir(face, sx, sy);
}
void DirectionToCubeFace(float3 d, int* face, float* sx, float* sy) {
    const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    if (ax >= ay && ax >= az) {
        if (d.x > 0) { *face = 0; *sx = -d.z / ax; *sy = d.y / ax; }
        else         { *face = 1; *sx =  d.z / ax; *sy = d.y / ax; }
    } else if (ay >= az) {
        if (d.y > 0) { *face = 2; *sx = d.x / ay; *sy = -d.z / ay; }
        else         { *face = 3; *sx = d.x / ay; *sy =  d.z / ay; }
    } else {
        if (d.z > 0) { *face = 4; *sx =  d.x / az; *sy = d.y / az; }
        else         { *face = 5; *sx = -d.x / az; *sy = d.y / az; }
    }
}

// Integer-only: texel (i,j) of `face` with exactly one coordinate one step outside [0,N) ->
// the texel of the neighbouring face whose centre direction is nearest (see DESIGN.md "cube edges").
void CubeResolveEdgeTap(int N, int face, int i, int j, int* of, int* oi, int* oj) {
    const int A = 2 * i + 1 - N;     // N * sx
    const int B = N - 1 - 2 * j;     // N * sy
    const int C = N;
    int dx, dy, dz;
    switch (face) {
        case 0: dx = C;  dy = B;  dz = -A; break;
        case 1: dx = -C; dy = B;  dz = A;  break;
        case 2: dx = A;  dy = C;  dz = -B; break;
        case 3: dx = A;  dy = -C; dz = B;  break;
        case 4: dx = A;  dy = B;  dz = C;  break;
        default: dx = -A; dy = B; dz = -C; break;
    }
    const int M = N + 1;             // the out-of-range component has magnitude N+1 > every other one
    int nf, nsx, nsy;
    if (dx == M)       { nf = 0; nsx = -dz; nsy = dy; }
    else if (dx == -M) { nf = 1; nsx = dz;  nsy = dy; }
    else if (dy == M)  { nf = 2; nsx = dx;  nsy = -dz; }
    else if (dy == -M) { nf = 3; nsx = dx;  nsy = dz; }
    else if (dz == M)  { nf = 4; nsx = dx;  nsy = dy; }
    else               { nf = 5; nsx = -dx; nsy = dy; }
    int ni = ((nsx + M) * N) / (2 * M);
    int nj = ((M - nsy) * N) / (2 * M);
    ni = std::min(std::max(ni, 0), N - 1);
    nj = std::min(std::max(nj, 0), N - 1);
    *of = nf; *oi = ni; *oj = nj;
}

float4 SampleCubeLevel(const Cubemap& tex, float3 dir, int mip) {
    mip = std::min(std::max(mip, 0), tex.mips - 1);
    const int N = tex.res >> mip;
    int face; float sx, sy;
    DirectionToCubeFace(dir, &face, &sx, &sy);
    const float u = sx * 0.5f + 0.5f;
    const float v = 0.5f - sy * 0.5f;
    const float x = u * (float)N - 0.5f;
    const float y = v * (float)N - 0.5f;
    if (!(std::isfinite(x) && std::isfinite(y))) return make4(0, 0, 0, 0);
    float x0 = std::floor(x), y0 = std::floor(y);
    int i0 = std::min(std::max((int)x0, -1), N - 1);
    int j0 = std::min(std::max((int)y0, -1), N - 1);
    const float fx = x - (float)i0, fy = y - (float)j0;
    float4 t[4]; bool corner[4];
    int nCorner = 0;
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + (k & 1), j = j0 + (k >> 1);
        const bool oi = (i < 0 || i >= N), oj = (j < 0 || j >= N);
        corner[k] = oi && oj;
        if (corner[k]) { ++nCorner; t[k] = make4(0, 0, 0, 0); continue; }
        if (oi || oj) {
            int f2, i2, j2;
            CubeResolveEdgeTap(N, face, i, j, &f2, &i2, &j2);
            t[k] = ld4(tex.texel(mip, f2, i2, j2));
        } else {
            t[k] = ld4(tex.texel(mip, face, i, j));
        }
    }
    if (nCorner) {   // at most one tap of a 2x2 footprint can be a cube corner (N >= 2)
        float4 s = make4(0, 0, 0, 0);
        for (int k = 0; k < 4; ++k) if (!corner[k]) s = s + t[k];
        const float4 m = s * (1.0f / 3.0f);
        for (int k = 0; k < 4; ++k) if (corner[k]) t[k] = m;
    }
    const float4 top = lerp(t[0], t[1], fx);
    const float4 bot = lerp(t[2], t[3], fx);
    return lerp(top, bot, fy);
}

float2 SampleLUT(const Image& lut, float u, float v) {
    const int W = lut.width, H = lut.height;
    const float x = u * (float)W - 0.5f;
    const float y = v * (float)H - 0.5f;
    const float x0 = std::floor(x), y0 = std::floor(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = std::min(std::max((int)x0, 0), W - 1), ix1 = std::min(std::max((int)x0 + 1, 0), W - 1);
    const int iy0 = std::min(std::max((int)y0, 0), H - 1), iy1 = std::min(std::max((int)y0 + 1, 0), H - 1);
    const float* p00 = lut.at(ix0, iy0); const float* p10 = lut.at(ix1, iy0);
    const float* p01 = lut.at(ix0, iy1); const float* p11 = lut.at(ix1, iy1);
    float2 r;
    r.x = lerp(lerp(p00[0], p10[0], fx), lerp(p01[0], p11[0], fx), fy);
    r.y = lerp(lerp(p00[1], p10[1], fx), lerp(p01[1], p11[1], fx), fy);
    return r;
}

// ------------------------------------------------------------------------------------------------
// Lighting.hlsl
// ------------------------------------------------------------------------------------------------
static inline float3 f3(const VqFloat3& v) { return {v.x, v.y, v.z}; }

float AttenuationBRDF(float dist) { return 1.0f / (dist * dist); }        // Lighting.hlsl:29-32

float SpotlightIntensity(const VqSpotLight& l, float3 worldPos) {         // Lighting.hlsl:57-73
    const float3 pixelDirectionInWorldSpace = normalize(worldPos - f3(l.position));
    const float3 spotDir = normalize(f3(l.spotDir));
    const float theta = std::acos(dot(pixelDirectionInWorldSpace, spotDir));
    if (theta > l.outerConeAngle) return 0.0f;
    if (theta <= l.innerConeAngle) return 1.0f;
    return 1.0f - (theta - l.innerConeAngle) / (l.outerConeAngle - l.innerConeAngle);
}

float3 CalculatePointLightIllumination(const VqPointLight& l, const BRDF_Surface& s, float3 P, float3 V) { // :308-322
    float3 IdIs = splat3(0.0f);
    const float3 Lw = f3(l.position);
    const float3 Wi = normalize(Lw - P);
    const float D = length(Lw - P);
    const float NdotL = saturate(dot(s.N, Wi));
    const float3 radiance = f3(l.color) * AttenuationBRDF(D) * l.brightness;
    if (D < l.range)
        IdIs += BRDF(s, Wi, V) * radiance * NdotL;
    return IdIs;
}

float3 CalculateSpotLightIllumination(const VqSpotLight& l, const BRDF_Surface& s, float3 P, float3 V) {  // :323-333
    float3 IdIs = splat3(0.0f);
    const float3 Wi = normalize(f3(l.position) - P);
    const float3 radiance = f3(l.color) * SpotlightIntensity(l, P) * l.brightness * AttenuationBRDF(length(f3(l.position) - P));
    const float NdotL = saturate(dot(s.N, Wi));
    IdIs += BRDF(s, Wi, V) * radiance * NdotL;
    return IdIs;
}

float3 CalculateDirectionalLightIllumination(const VqDirectionalLight& l, const BRDF_Surface& s, float3 V) { // :334-345
    const float3 Wi = normalize(-f3(l.lightDirection));
    const float3 radiance = f3(l.color) * l.brightness;
    const float NdotL = saturate(dot(s.N, Wi));
    return BRDF(s, Wi, V) * radiance * NdotL;
}

// GetHDRIRotationMatrix (Lighting.hlsl:348-358) applied as mul(v, m) with v a row vector.
static inline float3 RotateByHDRIOffset(float3 v, float offsetRad) {
    const float cosB = std::cos(-offsetRad);
    const float sinB = std::sin(-offsetRad);
    // m = {cosB,0,sinB; 0,1,0; -sinB,0,cosB};  (v*m).x = v.x*m00 + v.y*m10 + v.z*m20
    return make3(v.x * cosB + v.y * 0.0f + v.z * -sinB,
                 v.x * 0.0f + v.y * 1.0f + v.z * 0.0f,
                 v.x * sinB + v.y * 0.0f + v.z * cosB);
}

float3 CalculateEnvironmentMapIllumination(const BRDF_Surface& s, float3 V, int MAX_REFLECTION_LOD,
        const Cubemap& texEnvMapDiff, const Cubemap& texEnvMapSpec, const Image& lut, float fHDRIOffsetRad) { // :360-380
    const float NdotV = saturate(dot(s.N, V));
    const float3 R = RotateByHDRIOffset(reflect(-V, s.N), fHDRIOffsetRad);
    const float3 N = RotateByHDRIOffset(s.N, fHDRIOffsetRad);
    const int MIP_LEVEL = (int)(s.roughness * (float)MAX_REFLECTION_LOD);
    const float3 spec = xyz(SampleCubeLevel(texEnvMapSpec, R, MIP_LEVEL));
    const float2 F0ScaleBias = SampleLUT(lut, NdotV, s.roughness);
    const float3 diff = xyz(SampleCubeLevel(texEnvMapDiff, N, 0));
    return EnvironmentBRDF(NdotV, s.roughness, s.metalness, s.diffuseColor, diff, spec, F0ScaleBias);
}

float3 CalculateEnvironmentMapIllumination_DiffuseOnly(const BRDF_Surface& s, float3 V,
        const Cubemap& texEnvMapDiff, float fHDRIOffsetRad) {             // :382-395
    const float NdotV = saturate(dot(s.N, V));
    const float3 N = RotateByHDRIOffset(s.N, fHDRIOffsetRad);
    const float3 diff = xyz(SampleCubeLevel(texEnvMapDiff, N, 0));
    return EnvironmentBRDF(NdotV, s.roughness, s.metalness, s.diffuseColor, diff, splat3(0.0f), make2(0, 0));
}

// ------------------------------------------------------------------------------------------------
// ForwardLighting.hlsl PSMain from line 285 on (the G-buffer carries what lines 226-283 produce).
// Shadow maps do not exist headless: every shadow factor is 1 ("noShadows").
// ------------------------------------------------------------------------------------------------
float4 ForwardLighting_PSMain(const VqPerFrameData& cbPerFrame, const VqPerViewLightingData& cbPerView,
                              float4 position_ao, float4 normal_roughness, float4 albedo_metalness,
                              const float4* emissive,
                              const Cubemap& texEnvMapDiff, const Cubemap& texEnvMapSpec, const Image& lut) {
    BRDF_Surface Surface;
    Surface.N = xyz(normal_roughness);
    Surface.roughness = normal_roughness.w;
    Surface.diffuseColor = xyz(albedo_metalness);
    Surface.metalness = albedo_metalness.w;
    Surface.emissiveColor = emissive ? xyz(*emissive) : splat3(0.0f);
    Surface.emissiveIntensity = emissive ? emissive->w : 0.0f;
    const float ao = position_ao.w;

    const float3 P = xyz(position_ao);
    const float3 V = normalize(f3(cbPerView.CameraPosition) - P);                  // :285

    float3 I_total = Surface.diffuseColor * ao + Surface.emissiveColor * Surface.emissiveIntensity;  // :290-293

    if (cbPerView.EnvironmentMapDiffuseOnlyIllumination)                           // :299-306
        I_total += CalculateEnvironmentMapIllumination_DiffuseOnly(Surface, V, texEnvMapDiff, cbPerFrame.fHDRIOffsetInRadians);
    else
        I_total += CalculateEnvironmentMapIllumination(Surface, V, (int)cbPerView.MaxEnvMapLODLevels,
                        texEnvMapDiff, texEnvMapSpec, lut, cbPerFrame.fHDRIOffsetInRadians);

    const VqSceneLighting& L = cbPerFrame.Lights;
    for (int p = 0; p < L.numPointLights; ++p)                                     // :310-313
        I_total += CalculatePointLightIllumination(L.point_lights[p], Surface, P, V);
    for (int sI = 0; sI < L.numSpotLights; ++sI)                                   // :314-317
        I_total += CalculateSpotLightIllumination(L.spot_lights[sI], Surface, P, V);

    for (int pc = 0; pc < L.numPointCasters; ++pc) {                               // :321-340, shadow factor 1
        const VqPointLight& l = L.point_casters[pc];
        const float D = length(f3(l.position) - P);
        if (D < l.range)
            I_total += CalculatePointLightIllumination(l, Surface, P, V) * 1.0f;
    }
    for (int sc = 0; sc < L.numSpotCasters; ++sc)                                  // :343-356, shadow factor 1
        I_total += CalculateSpotLightIllumination(L.spot_casters[sc], Surface, P, V) * 1.0f;

    if (L.directional.enabled)                                                     // :360-377, shadow factor 1
        I_total += CalculateDirectionalLightIllumination(L.directional, Surface, V) * 1.0f;

    return make4(I_total, Surface.roughness);                                      // :380
}

// ------------------------------------------------------------------------------------------------
void ParallelRows(int n, int threads, void (*f)(int, void*), void* user) {
    if (threads <= 1 || n <= 1) { for (int r = 0; r < n; ++r) f(r, user); return; }
    threads = std::min(threads, n);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const int b = (int)((int64_t)n * t / threads), e = (int)((int64_t)n * (t + 1) / threads);
        pool.emplace_back([=]() { for (int r = b; r < e; ++r) f(r, user); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace orc
