// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// oracle.h — scalar fp32 CPU restatement of VQEngine's per-pixel shading math.
// Every function cites the reference file:line it follows. PARITY STATUS: the reference ships no
// golden vectors or numeric tests for this path — the restatement is pinned by
//   (1) the FidelityFX A_CPU setup functions compiled from the reference's own headers
//       (oracle/_ref/libffxref.so, built by oracle/Makefile; tests/test_oracle_pins.py), and
//   (2) the analytic known-answer values derived in SURVEY.md §8(c) (tests/golden/kat.json),
//   (3) the reference's own code where it compiles here: stb codec / resizer, Image.cpp, DXGIUtils.cpp MipImage,
//       PostProcess.cpp, LightingConstantBufferData.h (oracle/_ref/lib*.so, tests/test_*_ref.py), and
//   (4) the reference's SHADER TEXT compiled as C++ (oracle/_ref/libhlslref.so = Shaders/*.hlsl through
//       ref_shim/hlsl_to_cpp.py + ref_shim/hlsl_compat/hlsl_compat.h): tests/test_hlsl_ref.py requires this oracle to
//       equal it bit for bit (whole PSMain incl. shadowed casters, tonemapper, blur, IBL integrals, skydome, reflection
//       composite, CAS / EASU / RCAS). What stays a decision is D3D's unspecified intrinsic and filter rounding
//       (hlsl_math.h, the samplers below) — see DESIGN.md §5.
#pragma once
#include "hlsl_math.h"
#include "../include/vq_shader_data.h"
#include <cstddef>
#include <vector>

namespace orc {

// ---- resource views (host memory) -------------------------------------------------------------
struct Image {            // float4 texels (float2 for LUTs: `channels` = 2)
    const float* data; int width, height; size_t pitch_floats; int channels;
    const float* at(int x, int y) const { return data + (size_t)y * pitch_floats + (size_t)x * channels; }
};
struct MutImage {
    float* data; int width, height; size_t pitch_floats; int channels;
    float* at(int x, int y) const { return data + (size_t)y * pitch_floats + (size_t)x * channels; }
    Image view() const { return {data, width, height, pitch_floats, channels}; }
};
struct Cubemap {          // packed mip-major / face-minor, float4
    const float* data; int res, mips;
    size_t offset(int mip, int face) const;   // in texels
    const float* texel(int mip, int face, int x, int y) const {
        return data + 4 * (offset(mip, face) + (size_t)y * (res >> mip) + x);
    }
};
struct Pyramid {          // packed levels, float4
    const float* data; int width, height, levels;
    size_t offset(int level) const;           // in texels
    int w(int l) const { return width >> l; }
    int h(int l) const { return height >> l; }
};

int    mip_level_count(uint64_t w, uint64_t h);        // Image.cpp:231-241
size_t cubemap_texel_count(int res, int mips);
int    cubemap_row_count(int res, int mips);
size_t pyramid_texel_count(int w, int h, int levels);

// ---- §8(f).1 surface producer (oracle_surface.cpp) -----------------------------------------------
struct Texture8 {         // RGBA8 UNORM, packed levels (width>>l) x (height>>l); data == nullptr -> null SRV (reads 0)
    const uint8_t* data; int width, height, levels;
    size_t offset(int level) const;           // in texels
    int w(int l) const { return width >> l; }
    int h(int l) const { return height >> l; }
};
struct MaterialTextures8 { Texture8 diffuse, normals, emissive, metalness, roughness, occl_rough_metal, local_ao; };
struct SurfaceIn  { float3 WorldSpacePosition, WorldSpaceNormal, WorldSpaceTangent; float2 uv; };   // ForwardLighting.hlsl:42-53
struct SurfaceOut { float3 P; float ao; float3 N; float roughness; float3 diffuseColor; float metalness;
                    float3 emissiveColor; float emissiveIntensity; };
void   MipImage_Box8(const uint8_t* src, uint8_t* dst, int width, int height);                       // DXGIUtils.cpp:263-287
float  TextureLod(const Texture8& tex, float2 ddx, float2 ddy, float bias);
float4 SampleTexture8(const Texture8& tex, float2 uv, float2 ddx, float2 ddy, float bias);
float3 UnpackNormal(float3 SampledNormal, float3 worldNormal, float3 worldTangent);                  // ShadingMath.hlsl:44-52
bool   Surface_PSMain(const SurfaceIn& In, float2 uv_ddx_raw, float2 uv_ddy_raw, const VqMaterialData& mat,
                      const MaterialTextures8& tex, float fAmbientLightingFactor, float ssao, bool alphaMask,
                      SurfaceOut* o);                                                                 // ForwardLighting.hlsl:226-283

// ---- BRDF.hlsl ---------------------------------------------------------------------------------
struct BRDF_Surface {     // BRDF.hlsl:50-58
    float3 N; float roughness; float3 diffuseColor; float metalness;
    float emissiveIntensity; float3 emissiveColor;
};
float  NormalDistributionGGX(float NdotH, float roughness);                          // BRDF.hlsl:65-79
float  Geometry_Smiths_SchlickGGX(float3 N, float3 V, float roughness);              // BRDF.hlsl:82-97
float  Geometry_Smiths_SchlickGGX_EnvironmentMap(float3 N, float3 V, float roughness);// BRDF.hlsl:100-115
float  Geometry_Smith(float3 N, float3 V, float3 L, float k);                        // BRDF.hlsl:118-121
float  GeometryEnvironmentMap(float3 N, float3 V, float3 L, float k);                // BRDF.hlsl:124-129
float3 Fresnel_Schlick(float3 N, float3 V, float3 F0);                               // BRDF.hlsl:132-136
float3 Fresnel_Gaussian(float3 H, float3 V, float3 F0);                              // BRDF.hlsl:140-147
float3 FresnelWithRoughness(float cosTheta, float3 F0, float roughness);             // BRDF.hlsl:152-156
float3 F_LambertDiffuse(float3 kd);                                                  // BRDF.hlsl:158-161
float3 BRDF(const BRDF_Surface& s, float3 Wi, float3 V);                             // BRDF.hlsl:163-194
float3 EnvironmentBRDF(float NdotV, float roughness, float metallic, float3 diffuseColor,
                       float3 diffuseIrradiance, float3 preFilteredSpecular, float2 F0ScaleBias); // BRDF.hlsl:196-207
float3 ImportanceSampleGGX(float2 Xi, float3 N, float roughness);                    // BRDF.hlsl:217-238
float2 IntegrateBRDF(float NdotV, float roughness, int sampleCount);                 // BRDF.hlsl:239-283

// ---- ShadingMath.hlsl --------------------------------------------------------------------------
float  RadicalInverse_VdC(uint32_t bits);                                            // ShadingMath.hlsl:87-95
float2 Hammersley(uint32_t i, uint32_t count);                                       // ShadingMath.hlsl:119-127
float2 DirectionToEquirectUV(float3 v);                                              // ShadingMath.hlsl:70-80
float3 SRGBToLinear_pow22(float3 c);                                                 // ShadingMath.hlsl:65

// ---- texture sampling semantics (decisions, identical in the CUDA kernels) --------
float4 SampleEquirectLevel(const Pyramid& tex, float2 uv, float lod);   // trilinear, WRAP in u and v
float4 SampleCubeLevel(const Cubemap& tex, float3 dir, int mip);        // bilinear, seamless edges/corners
float2 SampleLUT(const Image& lut, float u, float v);                   // bilinear, CLAMP
// integer resolve of a one-texel-outside tap to the neighbouring face (used by SampleCubeLevel)
void   CubeResolveEdgeTap(int N, int face, int i, int j, int* of, int* oi, int* oj);
float3 CubeTexelDirection(int face, int px, int py, int res);           // A35: un-normalised direction
void   DirectionToCubeFace(float3 d, int* face, float* sx, float* sy);  // inverse of the above (ndc coords)

// ---- Lighting.hlsl -----------------------------------------------------------------------------
float  AttenuationBRDF(float dist);                                                  // Lighting.hlsl:29-32
float  SpotlightIntensity(const VqSpotLight& l, float3 worldPos);                    // Lighting.hlsl:57-73
float3 CalculatePointLightIllumination(const VqPointLight& l, const BRDF_Surface& s, float3 P, float3 V); // :308-322
float3 CalculateSpotLightIllumination(const VqSpotLight& l, const BRDF_Surface& s, float3 P, float3 V);   // :323-333
float3 CalculateDirectionalLightIllumination(const VqDirectionalLight& l, const BRDF_Surface& s, float3 V); // :334-345
float3 CalculateEnvironmentMapIllumination(const BRDF_Surface& s, float3 V, int MAX_REFLECTION_LOD,
        const Cubemap& texEnvMapDiff, const Cubemap& texEnvMapSpec, const Image& lut, float fHDRIOffsetRad); // :360-380
float3 CalculateEnvironmentMapIllumination_DiffuseOnly(const BRDF_Surface& s, float3 V,
        const Cubemap& texEnvMapDiff, float fHDRIOffsetRad);                         // :382-395

// ---- ForwardLighting.hlsl PSMain over one G-buffer texel (ForwardLighting.hlsl:285-380) ---------
float4 ForwardLighting_PSMain(const VqPerFrameData& cbPerFrame, const VqPerViewLightingData& cbPerView,
                              float4 position_ao, float4 normal_roughness, float4 albedo_metalness,
                              const float4* emissive,   // may be null
                              const Cubemap& texEnvMapDiff, const Cubemap& texEnvMapSpec, const Image& lut);

// ---- shadow tests (oracle_shadow.cpp; SURVEY A25 / §8(f).4 groundwork — the product still uses shadow factor 1) ----
struct ShadowTestPCFData { float4 lightSpacePos; float depthBias, NdotL, viewDistanceOfPixel; };     // Lighting.hlsl:79-87
struct ShadowMaps {       // linear R32F; any pointer may be null (= that light type is lit unshadowed)
    const float* pointCubes; int pointRes;        // [caster][face][y][x], value = distance / light range (Lighting.hlsl:156-158)
    const float* spotMaps; int spotW, spotH;      // [caster][y][x], value = light-space depth
    const float* dirMap; int dirW, dirH;          // [y][x]
};
float  SamplePoint2D(const float* map, int w, int h, float u, float v);                // POINT filter, WRAP (RootSignatures.cpp:148)
float  SamplePointCube(const float* cube, int res, float3 dir);
float  OmnidirectionalShadowTestPCF(const ShadowTestPCFData& pcf, const float* cube, int res, float3 lightVectorWorldSpace, float fFarPlane); // :113-165
float  ShadowTestPCF(const ShadowTestPCFData& pcf, const float* map, int w, int h, float2 shadowMapDimensions);             // :168-211
float  ShadowTestPCF_Directional(const ShadowTestPCFData& pcf, const float* map, int w, int h, float2 shadowMapDimensions); // :215-263
int    DepthMinPyramid(const float* depth, int w, int h, float* levels, int max_levels);       // DownsampleDepth.hlsl:50-119
float4 ForwardLighting_PSMain_Shadowed(const VqPerFrameData& cbPerFrame, const VqPerViewLightingData& cbPerView,
                                       float4 position_ao, float4 normal_roughness, float4 albedo_metalness, const float4* emissive,
                                       const Cubemap& texEnvMapDiff, const Cubemap& texEnvMapSpec, const Image& lut,
                                       const ShadowMaps& sm);                                                                // ForwardLighting.hlsl:285-380

// ---- CubemapConvolution.hlsl / IBL -------------------------------------------------------------
void   MipImage_MinFilter(const float* src, float* dst, int width, int height);       // DXGIUtils.cpp:289-317
// phi/theta sequences of PSMain_DiffuseIrradiance's loops (CubemapConvolution.hlsl:129-135)
void   DiffuseIrradianceAngles(float step, int n_phi, int n_theta, std::vector<float>& phis, std::vector<float>& thetas);
// f64Accum = false: the HLSL's sequential fp32 running sum. true: diagnostic variant that adds the SAME fp32 terms in
// double, to separate the reference's own summation error (O(n*eps) for 99k terms) from kernel error.
float4 DiffuseIrradiance_PSMain(const Pyramid& hdri, float3 lookDir, const std::vector<float>& phis,
                                const std::vector<float>& thetas, int srcMip, bool f64Accum = false);   // :112-163
float4 SpecularIrradiance_PSMain(const Pyramid& hdri, float3 lookDir, float Roughness,
                                 float2 TextureDimensionsLOD0, uint32_t numSamples,
                                 float sampleLengthScale = 1.0f);  // :168-223

// ---- post chain --------------------------------------------------------------------------------
float3 Tonemap_Reinhard(float3 c);                                                    // Tonemapper.hlsl:24-27
float3 LinearToSRGB(float3 c);                                                        // HDR.hlsl:76-80
float3 SRGBToLinear(float3 c);                                                        // HDR.hlsl:82-86
float3 Rec709ToRec2020(float3 c);                                                     // HDR.hlsl:88-97
float3 Rec2020ToRec709(float3 c);                                                     // HDR.hlsl:99-108
float3 LinearToST2084(float3 c);                                                      // HDR.hlsl:110-119
float4 Tonemapper_CSMain(const VqTonemapperParams& p, float4 in);                     // Tonemapper.hlsl:110-151
float4 GaussianBlur_CSMain(const Image& in, int x, int y, bool vertical, int sizeX, int sizeY); // GaussianBlur.hlsl:119-186

// FidelityFX (restated; setup functions are pinned bit-for-bit against oracle/_ref)
float  APrxLoSqrtF1(float a);    // ffx_a.h:1842
float  APrxLoRcpF1(float a);     // ffx_a.h:1843
float  APrxMedRcpF1(float a);    // ffx_a.h:1844
float  APrxLoRsqF1(float a);     // ffx_a.h:1845
void   CasSetup(uint32_t const0[4], uint32_t const1[4], float sharpness, float inX, float inY, float outX, float outY); // ffx_cas.h:375-394
void   FsrEasuCon(uint32_t con0[4], uint32_t con1[4], uint32_t con2[4], uint32_t con3[4],
                  float inVpX, float inVpY, float inSzX, float inSzY, float outX, float outY);  // ffx_fsr1.h:156-202
void   FsrRcasCon(uint32_t con[4], float sharpnessStops);                             // ffx_fsr1.h:662-672
void   SpdSetup(uint32_t dispatchXY[2], uint32_t workGroupOffset[2], uint32_t numWorkGroupsAndMips[2],
                const uint32_t rectInfo[4], int mips);                                // ffx_spd.h:327-351
float3 CasFilter_NoScaling(const Image& in, int x, int y, const uint32_t const1[4]);  // ffx_cas.h:408-537
float3 FsrEasuF(const Image& in, int x, int y, const uint32_t con[16], int addressMode); // ffx_fsr1.h:315-437
float3 FsrRcasF(const Image& in, int x, int y, const uint32_t con[4]);                // ffx_fsr1.h:684-769
// SPD: dst level L+1 from level L with the reduction order of the LDS path (SURVEY.md §9 "SPD")
void   SpdDownsampleLevel(const Image& src, const MutImage& dst, int dstLevel /*1-based*/);

// ---- §8(f).2/(f).3 on-disk formats and frame composition (oracle_frame.cpp) -------------------------
float4 HdrConvert(const uint8_t rgbe[4]);                                            // stb_image.h stbi__hdr_convert, req_comp 4
int    HdrDecode(const uint8_t* file, size_t n, int* w, int* h, std::vector<float>* rgba);   // stbi__hdr_load (Image.cpp:119-121)
float  CalculateMaxLuminance(const float* rgba, int width, int height);              // Image.cpp:43-86
void   LinearToRgbe(uint8_t rgbe[4], const float linear[3]);                         // stb_image_write.h stbiw__linear_to_rgbe
void   HdrEncode(const float* rgba, int width, int height, std::vector<uint8_t>* file);      // stbi_write_hdr_core (Image.cpp:210-213)
int    ResizeFloat4_Downsample(const float* in, int w, int h, float* out, int ow, int oh);  // stbir_resize_float (Image.cpp:148-190)
float3 SkydomeLookDirection(const VqMatrix& invViewProj, int px, int py, int width, int height);
float4 Skydome_PSMain(const Pyramid& texEquirectEnvironmentMap, const VqMatrix& invViewProj,
                      int px, int py, int width, int height);                        // Skydome.hlsl:35-56
float4 ApplyReflections_CSMain(float4 SceneRadianceAndRoughness, float4 ReflectionRadiance, const float4* bv); // ApplyReflections.hlsl:31-57

// run f(row) for rows [0,n) on `threads` std::threads (contiguous row blocks)
void   ParallelRows(int n, int threads, void (*f)(int row, void* user), void* user);

}  // namespace orc
