/*
 * vq_shader_data.h — plain-C POD parameter blocks for the shading backend.
 *
 * These are OUR restatement of the layouts the engine shares between CPU and GPU
 * (reference: Shaders/LightingConstantBufferData.h:50-186, compiled both ways
 * through Shaders/VQPlatform.h:4-26). They are byte-compatible, so the engine's
 * own `PerFrameData` / `PerViewLightingData` / `FTonemapper` / FidelityFX
 * constant blocks can be passed by pointer unchanged.
 *
 * Layout facts pinned with static_asserts below:
 *   PointLight 48 B, SpotLight 64 B, DirectionalLight 40 B,
 *   SceneLighting 7088 B (matrix starts 16-B aligned -> 8 B pad after directional),
 *   PerFrameData 7120 B, PerViewLightingData 320 B.
 */
#ifndef VQ_SHADER_DATA_H
#define VQ_SHADER_DATA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference: LightingConstantBufferData.h:40-47 */
#define VQ_NUM_LIGHTS_POINT            100
#define VQ_NUM_LIGHTS_SPOT             20
#define VQ_NUM_SHADOWING_LIGHTS_POINT  5
#define VQ_NUM_SHADOWING_LIGHTS_SPOT   5

typedef struct VqFloat2 { float x, y; } VqFloat2;
typedef struct VqFloat3 { float x, y, z; } VqFloat3;
typedef struct VqFloat4 { float x, y, z, w; } VqFloat4;
typedef struct VqMatrix { float m[16]; } VqMatrix; /* 64 B, row-major as XMMATRIX stores it */

/* reference: LightingConstantBufferData.h:50-62 */
typedef struct VqPointLight {
    VqFloat3 position;    float range;
    VqFloat3 color;       float brightness;
    VqFloat3 attenuation; float depthBias;
} VqPointLight;

/* reference: LightingConstantBufferData.h:64-79 (64 B; the "48 bytes" comment there is stale) */
typedef struct VqSpotLight {
    VqFloat3 position;  float outerConeAngle;
    VqFloat3 color;     float brightness;
    VqFloat3 spotDir;   float depthBias;
    float innerConeAngle; float range; float dummy1; float dummy2;
} VqSpotLight;

/* reference: LightingConstantBufferData.h:81-90 */
typedef struct VqDirectionalLight {
    VqFloat3 lightDirection; float brightness;
    VqFloat3 color;          float depthBias;
    int32_t shadowing;
    int32_t enabled;
} VqDirectionalLight;

/* reference: LightingConstantBufferData.h:92-109 */
typedef struct VqSceneLighting {
    int32_t numPointLights;   /* non-shadow caster counts */
    int32_t numSpotLights;
    int32_t numPointCasters;  /* shadow caster counts */
    int32_t numSpotCasters;
    VqDirectionalLight directional;
    uint32_t _pad_matrix_align[2];      /* HLSL packing == alignas(16) XMMATRIX on the CPU */
    VqMatrix shadowViewDirectional;
    VqPointLight point_lights[VQ_NUM_LIGHTS_POINT];
    VqPointLight point_casters[VQ_NUM_SHADOWING_LIGHTS_POINT];
    VqSpotLight  spot_lights[VQ_NUM_LIGHTS_SPOT];
    VqSpotLight  spot_casters[VQ_NUM_SHADOWING_LIGHTS_SPOT];
    VqMatrix     shadowViews[VQ_NUM_SHADOWING_LIGHTS_SPOT];
} VqSceneLighting;

/* reference: LightingConstantBufferData.h:164-172; filled at SceneRendering.cpp:429-450 */
typedef struct VqPerFrameData {
    VqSceneLighting Lights;
    VqFloat2 f2PointLightShadowMapDimensions;
    VqFloat2 f2SpotLightShadowMapDimensions;
    VqFloat2 f2DirectionalLightShadowMapDimensions;
    float fAmbientLightingFactor;
    float fHDRIOffsetInRadians;
} VqPerFrameData;

/* reference: LightingConstantBufferData.h:173-186; filled at SceneRendering.cpp:452-467 */
typedef struct VqPerViewLightingData {
    VqMatrix matView;
    VqMatrix matViewToWorld;
    VqMatrix matProjInverse;
    VqFloat4 WorldFrustumPlanes[6];
    VqFloat3 CameraPosition;
    float    MaxEnvMapLODLevels;
    VqFloat2 ScreenDimensions;
    int32_t  EnvironmentMapDiffuseOnlyIllumination;
    float    pad1;
} VqPerViewLightingData;

/* reference: Renderer/Rendering/HDR.h enums, Shaders/HDR.hlsl:33-40 */
enum { VQ_COLOR_SPACE_REC_709 = 0, VQ_COLOR_SPACE_REC_2020 = 1 };
enum { VQ_DISPLAY_CURVE_SRGB = 0, VQ_DISPLAY_CURVE_ST2084 = 1, VQ_DISPLAY_CURVE_LINEAR = 2 };

/* reference: PostProcess.h:84-91 (FTonemapper; the shader cbuffer Tonemapper.hlsl:98-104 reads the first 4 fields) */
typedef struct VqTonemapperParams {
    int32_t ContentColorSpace;
    int32_t OutputDisplayCurve;
    float   DisplayReferenceBrightnessLevel;
    int32_t ToggleGammaCorrection;
    float   UIHDRBrightness;
} VqTonemapperParams;

/* reference: PostProcess.h:92-96 (FBlurParams), GaussianBlur.hlsl:62-65 */
typedef struct VqBlurParams { int32_t iImageSizeX, iImageSizeY; } VqBlurParams;

/* reference: EnvironmentMapRendering.cpp:178 (cb1_t), CubemapConvolution.hlsl:49-54 */
typedef struct VqConvolutionParams {
    float ViewDimX, ViewDimY;   /* = HDRI dimensions for the specular pass (EnvironmentMapRendering.cpp:433-434) */
    float Roughness;
    int32_t MIP;
} VqConvolutionParams;

/* reference: AMDFidelityFX.hlsl:394-399 (spdConstants) */
typedef struct VqSpdConstants {
    uint32_t mips;
    uint32_t numWorkGroups;
    uint32_t workGroupOffset[2];
} VqSpdConstants;

/* reference: LightingConstantBufferData.h:126-143 (MaterialData, 80 B; textureConfig is the Has*Map bitfield
 * of Material::GetTextureConfig (Material.cpp:23-36) converted to float, Material.h:126) */
typedef struct VqMaterialData {
    VqFloat3 diffuse;       float alpha;
    VqFloat3 emissiveColor; float emissiveIntensity;
    VqFloat3 specular;      float normalMapMipBias;
    VqFloat4 uvScaleOffset;
    float roughness, metalness, displacement, textureConfig;
} VqMaterialData;

/* reference: LightingConstantBufferData.h:116-124 (Has*Map bits) */
#define VQ_TEXCFG_DIFFUSE    (1 << 0)
#define VQ_TEXCFG_NORMAL     (1 << 1)
#define VQ_TEXCFG_AO         (1 << 2)
#define VQ_TEXCFG_ALPHA_MASK (1 << 3)
#define VQ_TEXCFG_ROUGHNESS  (1 << 4)
#define VQ_TEXCFG_METALLIC   (1 << 5)
#define VQ_TEXCFG_HEIGHT     (1 << 6)
#define VQ_TEXCFG_EMISSIVE   (1 << 7)
#define VQ_TEXCFG_ORM        (1 << 8)

#ifdef __cplusplus
} /* extern "C" */
static_assert(sizeof(VqMaterialData) == 80, "MaterialData layout");
static_assert(sizeof(VqPointLight) == 48, "PointLight layout");
static_assert(sizeof(VqSpotLight) == 64, "SpotLight layout");
static_assert(sizeof(VqDirectionalLight) == 40, "DirectionalLight layout");
static_assert(offsetof(VqSceneLighting, shadowViewDirectional) == 64, "matrix must be 16-B aligned");
static_assert(offsetof(VqSceneLighting, point_lights) == 128, "point_lights offset");
static_assert(offsetof(VqSceneLighting, spot_lights) == 5168, "spot_lights offset");
static_assert(sizeof(VqSceneLighting) == 7088, "SceneLighting layout");
static_assert(sizeof(VqPerFrameData) == 7120, "PerFrameData layout");
static_assert(sizeof(VqPerViewLightingData) == 320, "PerViewLightingData layout");
static_assert(offsetof(VqPerViewLightingData, CameraPosition) == 288, "CameraPosition offset");
static_assert(sizeof(VqTonemapperParams) == 20, "FTonemapper layout");
#endif

#endif /* VQ_SHADER_DATA_H */
