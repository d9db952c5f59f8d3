// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// hlsl_ref_shim.cpp -> oracle/_ref/libhlslref.so: the reference's own shader text, compiled as C++.
//
// The *.inc files included below are generated at build time by hlsl_to_cpp.py from /root/reference/Shaders/*.hlsl
// (read in place; the generated text lives only under the git-ignored _ref/obj/hlsl/). Each shader sits in its own
// namespace on top of hlsl_compat.h, which supplies the HLSL types and intrinsics with the semantics oracle/hlsl_math.h
// documents. This file is the "runtime" around them: it fills the shaders' constant-buffer globals from the product's
// byte-compatible Vq* structs (matrices loaded the way HLSL's default column-major packing reads an XMMATRIX), binds
// the textures, and serves every Sample / Load through the ORACLE's samplers (liboracle.so) — filtering is D3D
// hardware, not shader text, so both sides of the comparison must see the same texels. tests/test_hlsl_ref.py then
// requires oracle == compiled shader text, bit for bit.
#include "hlsl_compat.h"
#include "../../include/vq_shader_data.h"
#include <pthread.h>
#include <thread>
#include <vector>

extern "C" {   // liboracle.so (oracle_capi.cpp)
void orc_sample_texture8(const uint8_t* tex, int w, int h, int levels, float u, float v, float dudx, float dvdx, float dudy,
                         float dvdy, float bias, float* out4, float* out_lod);
void orc_sample_cube(const float* cube, int res, int mips, const float d[3], int mip, float out[4]);
void orc_sample_equirect(const float* pyr, int w, int h, int levels, float u, float v, float lod, float out[4]);
void orc_sample_lut(const float* lut, int w, int h, float u, float v, float out[2]);
float orc_sample_point2d(const float* map, int w, int h, float u, float v);
float orc_sample_point_cube(const float* cube, int res, const float d[3]);
}

namespace hl {
HlTextureOps g_tex;

namespace fwd  {
#include "forward.inc"
}
namespace fwdm {   // ENABLE_ALPHA_MASK=1 permutation
#include "forward_alphamask.inc"
}
namespace tonemap {
#include "Tonemapper.inc"
}
namespace conv {
#include "CubemapConvolution.inc"
}
namespace blur {
#include "GaussianBlur.inc"
}
namespace sky {
#include "Skydome.inc"
}
namespace refl {
#include "ApplyReflections.inc"
}
namespace reflbv {   // COMPOSITE_BOUNDING_VOLUMES=1 permutation
#include "ApplyReflections_bv.inc"
}
namespace cas {      // AMDFidelityFX.hlsl, FFXCAS_CS=1 (fp32, FFXCAS_NO_UPSCALING=1: the engine's defaults)
#include "ffx_cas.inc"
}
namespace easu {     // AMDFidelityFX.hlsl, FSR_EASU_CS=1 (fp32)
#include "ffx_easu.inc"
}
namespace rcas {     // AMDFidelityFX.hlsl, FSR_RCAS_CS=1 (fp32)
#include "ffx_rcas.inc"
}
void (*g_group_barrier)() = nullptr;
namespace spd {      // AMDFidelityFX.hlsl, FFXSPD_CS=1, SPD_NO_WAVE_OPERATIONS (LDS path). The engine's PSO for it is commented
#include "ffx_spd.inc"   // out and the section lacks its own #include of SPD/ffx_a.h: the build passes it with --pre-include.
}
namespace depth {    // DownsampleDepth.hlsl (the engine's live SPD user: MIN depth pyramid), SPD_NO_WAVE_OPERATIONS
#include "depth.inc"
}
}  // namespace hl

namespace {

enum Kind { K_NULL = 0, K_TEX8, K_EQUIRECT, K_CUBE, K_LUT, K_PLANE1_POINT, K_R32_2D_ARRAY, K_R32_CUBE_ARRAY, K_IMAGE4, K_IMAGE2, K_IMAGE3IN4, K_IMAGE1, K_CONST };
struct Binding { int kind; const void* ptr; void* wptr; int w, h, levels; float constant; };
Binding g_bind[64];
float g_ddx[2], g_ddy[2];   // implicit derivatives of the scaled uv for the pixel being shaded (see forward_image)

void fetch(void*, int tid, int /*sid*/, int kind, const float c[4], float lod, float out[4]) {
    const Binding& b = g_bind[tid];
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    switch (b.kind) {
        case K_NULL: return;                                                    // null SRV reads 0
        case K_CONST: out[0] = out[1] = out[2] = out[3] = b.constant; return;
        case K_TEX8:                                                            // Sample / SampleBias, implicit derivatives
            orc_sample_texture8((const uint8_t*)b.ptr, b.w, b.h, b.levels, c[0], c[1], g_ddx[0], g_ddx[1], g_ddy[0], g_ddy[1],
                                kind == 2 ? lod : 0.0f, out, nullptr);
            return;
        case K_EQUIRECT: orc_sample_equirect((const float*)b.ptr, b.w, b.h, b.levels, c[0], c[1], lod, out); return;
        case K_CUBE: orc_sample_cube((const float*)b.ptr, b.w, b.levels, c, (int)lod, out); return;
        case K_LUT: orc_sample_lut((const float*)b.ptr, b.w, b.h, c[0], c[1], out); return;
        case K_PLANE1_POINT: out[0] = orc_sample_point2d((const float*)b.ptr, b.w, b.h, c[0], c[1]); return;
        case K_R32_2D_ARRAY:
            out[0] = orc_sample_point2d((const float*)b.ptr + (size_t)(int)c[2] * b.w * b.h, b.w, b.h, c[0], c[1]); return;
        case K_R32_CUBE_ARRAY:
            out[0] = orc_sample_point_cube((const float*)b.ptr + (size_t)(int)c[3] * 6 * b.w * b.w, b.w, c); return;
        case K_IMAGE4: {
            const float* img = (const float*)b.ptr;
            if (kind >= 4) {   // Gather R/G/B with CLAMP addressing: footprint floor(uv*size - 0.5) .. +1, order (0,1) (1,1) (1,0) (0,0)
                const int ch = kind - 4;
                const int x0 = (int)std::floor(c[0] * (float)b.w - 0.5f), y0 = (int)std::floor(c[1] * (float)b.h - 0.5f);
                auto at = [&](int x, int y) { x = x < 0 ? 0 : (x >= b.w ? b.w - 1 : x); y = y < 0 ? 0 : (y >= b.h ? b.h - 1 : y); return img[((size_t)y * b.w + x) * 4 + ch]; };
                out[0] = at(x0, y0 + 1); out[1] = at(x0 + 1, y0 + 1); out[2] = at(x0 + 1, y0); out[3] = at(x0, y0);
                return;
            }
            const int x = (int)c[0], y = (int)c[1];
            if (x < 0 || y < 0 || x >= b.w || y >= b.h) return;                  // out-of-range Load reads 0
            const float* p = img + ((size_t)y * b.w + x) * 4; out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3]; return;
        }
        case K_IMAGE1: {                                                        // R32F plane: Load (0 outside), size query
            if (kind == 7) { out[0] = (float)b.w; out[1] = (float)b.h; return; }
            const int x = (int)c[0], y = (int)c[1];
            if (x < 0 || y < 0 || x >= b.w || y >= b.h) return;
            out[0] = ((const float*)b.ptr)[(size_t)y * b.w + x]; return;
        }
        case K_IMAGE2: { const float* p = (const float*)b.ptr + ((size_t)(int)c[1] * b.w + (int)c[0]) * 2; out[0] = p[0]; out[1] = p[1]; return; }
    }
}
void store(void*, int tid, int x, int y, const float v[4]) {
    const Binding& b = g_bind[tid];
    if (x < 0 || y < 0 || x >= b.w || y >= b.h) return;                          // out-of-range UAV writes are dropped
    if (b.kind == K_IMAGE1) { ((float*)b.wptr)[(size_t)y * b.w + x] = v[0]; return; }
    const int n = b.kind == K_IMAGE2 ? 2 : (b.kind == K_IMAGE3IN4 ? 3 : 4);       // K_IMAGE3IN4: float3 UAV over an RGBA plane
    float* p = (float*)b.wptr + ((size_t)y * b.w + x) * (b.kind == K_IMAGE2 ? 2 : 4);
    for (int i = 0; i < n; ++i) p[i] = v[i];
}
void install() { hl::g_tex.user = nullptr; hl::g_tex.fetch = fetch; hl::g_tex.store = store; }
void bind(int id, int kind, const void* ptr, int w, int h, int levels, float constant = 0.0f, void* wptr = nullptr) {
    g_bind[id] = Binding{ptr || kind == K_CONST || wptr ? kind : K_NULL, ptr, wptr, w, h, levels, constant};
}

hl::float3 f3(const VqFloat3& v) { return hl::float3(v.x, v.y, v.z); }
// An XMMATRIX is uploaded as 16 floats, row after row; HLSL's `matrix` in a cbuffer is column-major by default, so the
// shader's M[i][j] is memory[j*4 + i].
hl::matrix load_matrix(const VqMatrix& m) {
    hl::matrix r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.row[i].v[j] = m.m[j * 4 + i];
    return r;
}

template <class NS_PointLight> void cp(NS_PointLight& d, const VqPointLight& s) {
    d.position = f3(s.position); d.range = s.range; d.color = f3(s.color); d.brightness = s.brightness;
    d.attenuation = f3(s.attenuation); d.depthBias = s.depthBias;
}
template <class NS_SpotLight> void cp(NS_SpotLight& d, const VqSpotLight& s) {
    d.position = f3(s.position); d.outerConeAngle = s.outerConeAngle; d.color = f3(s.color); d.brightness = s.brightness;
    d.spotDir = f3(s.spotDir); d.depthBias = s.depthBias; d.innerConeAngle = s.innerConeAngle; d.range = s.range;
    d.dummy1 = s.dummy1; d.dummy2 = s.dummy2;
}
template <class PF> void load_per_frame(PF& d, const VqPerFrameData& s) {
    d.Lights.numPointLights = s.Lights.numPointLights; d.Lights.numSpotLights = s.Lights.numSpotLights;
    d.Lights.numPointCasters = s.Lights.numPointCasters; d.Lights.numSpotCasters = s.Lights.numSpotCasters;
    auto& dl = d.Lights.directional; const auto& sl = s.Lights.directional;
    dl.lightDirection = f3(sl.lightDirection); dl.brightness = sl.brightness; dl.color = f3(sl.color); dl.depthBias = sl.depthBias;
    dl.shadowing = sl.shadowing; dl.enabled = sl.enabled;
    d.Lights.shadowViewDirectional = load_matrix(s.Lights.shadowViewDirectional);
    for (int i = 0; i < VQ_NUM_LIGHTS_POINT; ++i) cp(d.Lights.point_lights[i], s.Lights.point_lights[i]);
    for (int i = 0; i < VQ_NUM_SHADOWING_LIGHTS_POINT; ++i) cp(d.Lights.point_casters[i], s.Lights.point_casters[i]);
    for (int i = 0; i < VQ_NUM_LIGHTS_SPOT; ++i) cp(d.Lights.spot_lights[i], s.Lights.spot_lights[i]);
    for (int i = 0; i < VQ_NUM_SHADOWING_LIGHTS_SPOT; ++i) { cp(d.Lights.spot_casters[i], s.Lights.spot_casters[i]); d.Lights.shadowViews[i] = load_matrix(s.Lights.shadowViews[i]); }
    d.f2PointLightShadowMapDimensions = hl::float2(s.f2PointLightShadowMapDimensions.x, s.f2PointLightShadowMapDimensions.y);
    d.f2SpotLightShadowMapDimensions = hl::float2(s.f2SpotLightShadowMapDimensions.x, s.f2SpotLightShadowMapDimensions.y);
    d.f2DirectionalLightShadowMapDimensions = hl::float2(s.f2DirectionalLightShadowMapDimensions.x, s.f2DirectionalLightShadowMapDimensions.y);
    d.fAmbientLightingFactor = s.fAmbientLightingFactor; d.fHDRIOffsetInRadians = s.fHDRIOffsetInRadians;
}
template <class PV> void load_per_view(PV& d, const VqPerViewLightingData& s) {
    d.matView = load_matrix(s.matView); d.matViewToWorld = load_matrix(s.matViewToWorld); d.matProjInverse = load_matrix(s.matProjInverse);
    for (int i = 0; i < 6; ++i) d.WorldFrustumPlanes[i] = hl::float4(s.WorldFrustumPlanes[i].x, s.WorldFrustumPlanes[i].y, s.WorldFrustumPlanes[i].z, s.WorldFrustumPlanes[i].w);
    d.CameraPosition = f3(s.CameraPosition); d.MaxEnvMapLODLevels = s.MaxEnvMapLODLevels;
    d.ScreenDimensions = hl::float2(s.ScreenDimensions.x, s.ScreenDimensions.y);
    d.EnvironmentMapDiffuseOnlyIllumination = s.EnvironmentMapDiffuseOnlyIllumination; d.pad1 = s.pad1;
}
template <class MD> void load_material(MD& d, const VqMaterialData& s) {
    d.diffuse = f3(s.diffuse); d.alpha = s.alpha; d.emissiveColor = f3(s.emissiveColor); d.emissiveIntensity = s.emissiveIntensity;
    d.specular = f3(s.specular); d.normalMapMipBias = s.normalMapMipBias;
    d.uvScaleOffset = hl::float4(s.uvScaleOffset.x, s.uvScaleOffset.y, s.uvScaleOffset.z, s.uvScaleOffset.w);
    d.roughness = s.roughness; d.metalness = s.metalness; d.displacement = s.displacement; d.textureConfig = s.textureConfig;
}

struct OrcTexture2D { const uint8_t* ptr; int32_t width, height, levels; };   // as oracle_capi.cpp
struct OrcMaterialTextures { OrcTexture2D t[7]; };                             // diffuse, normals, emissive, alpha mask? see forward_image

// texture ids handed to the shaders' resource globals
enum { T_DIFFUSE = 1, T_NORMALS, T_EMISSIVE, T_ALPHAMASK, T_METALNESS, T_ROUGHNESS, T_ORM, T_LOCALAO, T_HEIGHT, T_SSAO,
       T_ENVDIFF, T_ENVSPEC, T_BRDFLUT, T_SHADOW_DIR, T_SHADOW_SPOT, T_SHADOW_POINT, T_IN, T_OUT, T_IN2, T_HDRI };

#define BIND_FORWARD_RESOURCES(NS)                                                                                     \
    hl::NS::texDiffuse.id = T_DIFFUSE; hl::NS::texNormals.id = T_NORMALS; hl::NS::texEmissive.id = T_EMISSIVE;            \
    hl::NS::texAlphaMask.id = T_ALPHAMASK; hl::NS::texMetalness.id = T_METALNESS; hl::NS::texRoughness.id = T_ROUGHNESS;  \
    hl::NS::texOcclRoughMetal.id = T_ORM; hl::NS::texLocalAO.id = T_LOCALAO; hl::NS::texHeightmap.id = T_HEIGHT;          \
    hl::NS::texScreenSpaceAO.id = T_SSAO; hl::NS::texEnvMapDiff.id = T_ENVDIFF; hl::NS::texEnvMapSpec.id = T_ENVSPEC;     \
    hl::NS::texBRDFIntegral.id = T_BRDFLUT; hl::NS::texDirectionalLightShadowMap.id = T_SHADOW_DIR;                      \
    hl::NS::texSpotLightShadowMaps.id = T_SHADOW_SPOT; hl::NS::texPointLightShadowMaps.id = T_SHADOW_POINT;

template <class PSInput, class PSMainFn, class CBO>
void forward_image_impl(PSMainFn psmain, CBO& cbPerObject,
                        const float* position_u, const float* normal_v, const float* tangent_m, int width, int height,
                        const VqMaterialData* materials, const OrcMaterialTextures* textures, int n_materials,
                        float* out_color, uint8_t* out_discarded) {
    auto rawuv = [&](int x, int y, float uv[2]) { const size_t o = ((size_t)y * width + x) * 4; uv[0] = position_u[o + 3]; uv[1] = normal_v[o + 3]; };
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const size_t o = ((size_t)y * width + x) * 4;
            int mi = (int)tangent_m[o + 3];
            mi = mi < 0 ? 0 : (mi >= n_materials ? n_materials - 1 : mi);
            const VqMaterialData& mat = materials[mi];
            load_material(cbPerObject.materialData, mat);
            // D3D: texture order of the material slots as the product passes them (vqcuda.h VqMaterialTextures):
            // diffuse, normals, emissive, metalness, roughness, occlusion-roughness-metalness, local AO
            const OrcMaterialTextures& mt = textures[mi];
            const int ids[7] = {T_DIFFUSE, T_NORMALS, T_EMISSIVE, T_METALNESS, T_ROUGHNESS, T_ORM, T_LOCALAO};
            for (int k = 0; k < 7; ++k) bind(ids[k], K_TEX8, mt.t[k].ptr, mt.t[k].width, mt.t[k].height, mt.t[k].levels);
            // implicit derivatives: fine differences of the scaled uv inside the aligned 2x2 quad (oracle_surface.cpp header)
            const int qx = x & ~1, qy = y & ~1;
            const int qx1 = qx + 1 < width ? qx + 1 : qx, qy1 = qy + 1 < height ? qy + 1 : qy;
            float a[2], b[2], c[2], d[2];
            rawuv(qx, y, a); rawuv(qx1, y, b); rawuv(x, qy, c); rawuv(x, qy1, d);
            g_ddx[0] = (b[0] - a[0]) * mat.uvScaleOffset.x; g_ddx[1] = (b[1] - a[1]) * mat.uvScaleOffset.y;
            g_ddy[0] = (d[0] - c[0]) * mat.uvScaleOffset.x; g_ddy[1] = (d[1] - c[1]) * mat.uvScaleOffset.y;

            PSInput In;
            In.position = hl::float4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);      // SV_Position: pixel centre
            In.WorldSpacePosition = hl::float3(position_u[o], position_u[o + 1], position_u[o + 2]);
            In.WorldSpaceNormal = hl::float3(normal_v[o], normal_v[o + 1], normal_v[o + 2]);
            In.WorldSpaceTangent = hl::float3(tangent_m[o], tangent_m[o + 1], tangent_m[o + 2]);
            float uv[2]; rawuv(x, y, uv);
            In.uv = hl::float2(uv[0], uv[1]);
            out_discarded[(size_t)y * width + x] = 0;
            try {
                const auto r = psmain(In);
                out_color[o] = r.color.x; out_color[o + 1] = r.color.y; out_color[o + 2] = r.color.z; out_color[o + 3] = r.color.w;
            } catch (const hl::Discard&) {
                out_discarded[(size_t)y * width + x] = 1;
            }
        }
}

// One workgroup of a shader that uses groupshared memory and barriers: one OS thread per lane, GroupMemoryBarrierWithGroupSync
// = a pthread barrier over all lanes. Workgroups run one after the other (any order is legal for a dispatch).
pthread_barrier_t g_bar;
void bar_wait() { pthread_barrier_wait(&g_bar); }
template <class F> void run_workgroup(int lanes, F lane) {
    pthread_barrier_init(&g_bar, nullptr, (unsigned)lanes);
    hl::g_group_barrier = bar_wait;
    std::vector<std::thread> th;
    th.reserve((size_t)lanes);
    for (int i = 0; i < lanes; ++i) th.emplace_back(lane, i);
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_bar);
}

template <class F> static void dispatch16(int w, int h, F csmain) {
    for (uint32_t gy = 0; gy < (uint32_t)((h + 15) / 16); ++gy)
        for (uint32_t gx = 0; gx < (uint32_t)((w + 15) / 16); ++gx)
            for (uint32_t t = 0; t < 64; ++t) csmain(hl::uint3(t, 0u, 0u), hl::uint3(gx, gy, 0u));
}

}  // namespace

extern "C" {

// ForwardLighting.hlsl PSMain over an image of interpolated vertex attributes (the same planes
// orc_gbuffer_from_materials takes), all textures bound, shadow maps optional (null = no casters expected).
void hlslref_forward_image(const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                           const float* position_u, const float* normal_v, const float* tangent_m, const float* ssao,
                           int width, int height, const VqMaterialData* materials, const void* textures, int n_materials,
                           int alpha_mask,
                           const float* diff_cube, int diff_res, const float* spec_cube, int spec_res, int spec_mips,
                           const float* lut, int lut_w, int lut_h,
                           const float* point_cubes, int point_res, const float* spot_maps, int spot_w, int spot_h,
                           const float* dir_map, int dir_w, int dir_h, float* out_color, uint8_t* out_discarded) {
    install();
    for (auto& b : g_bind) b = Binding{};
    if (ssao) bind(T_SSAO, K_PLANE1_POINT, ssao, width, height, 1); else bind(T_SSAO, K_CONST, nullptr, 0, 0, 0, 1.0f);
    bind(T_ENVDIFF, K_CUBE, diff_cube, diff_res, diff_res, 1);
    bind(T_ENVSPEC, K_CUBE, spec_cube, spec_res, spec_res, spec_mips);
    bind(T_BRDFLUT, K_LUT, lut, lut_w, lut_h, 1);
    bind(T_SHADOW_DIR, K_PLANE1_POINT, dir_map, dir_w, dir_h, 1);
    bind(T_SHADOW_SPOT, K_R32_2D_ARRAY, spot_maps, spot_w, spot_h, 1);
    bind(T_SHADOW_POINT, K_R32_CUBE_ARRAY, point_cubes, point_res, point_res, 1);
    const OrcMaterialTextures* tx = (const OrcMaterialTextures*)textures;
    if (alpha_mask) {
        BIND_FORWARD_RESOURCES(fwdm)
        load_per_frame(hl::fwdm::cbPerFrame, *pf); load_per_view(hl::fwdm::cbPerView, *pv);
        forward_image_impl<hl::fwdm::PSInput>(hl::fwdm::PSMain, hl::fwdm::cbPerObject, position_u, normal_v, tangent_m, width, height,
                                              materials, tx, n_materials, out_color, out_discarded);
    } else {
        BIND_FORWARD_RESOURCES(fwd)
        load_per_frame(hl::fwd::cbPerFrame, *pf); load_per_view(hl::fwd::cbPerView, *pv);
        forward_image_impl<hl::fwd::PSInput>(hl::fwd::PSMain, hl::fwd::cbPerObject, position_u, normal_v, tangent_m, width, height,
                                             materials, tx, n_materials, out_color, out_discarded);
    }
}

// ForwardLighting.hlsl PSMain driven from a G-buffer (the K1 workload): per pixel the material constants are set to the
// G-buffer texel (diffuse = albedo, roughness, metalness, emissive; textureConfig = 0, every material map a null SRV, SSAO = 1,
// fAmbientLightingFactor = the texel's ao), so the UNMODIFIED PSMain text shades exactly what K1 shades. Used by
// bench.py --impl reference as the reference's own implementation of the path on the CPU (one process per core: the
// shader's cbuffer globals are per process).
void hlslref_forward_gbuffer_rows(const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                                  const float* position_ao, const float* normal_roughness, const float* albedo_metalness,
                                  const float* emissive /* may be null */, int width, int row_begin, int row_end,
                                  const float* diff_cube, int diff_res, const float* spec_cube, int spec_res, int spec_mips,
                                  const float* lut, int lut_w, int lut_h, float* out_color) {
    install();
    for (auto& b : g_bind) b = Binding{};
    bind(T_SSAO, K_CONST, nullptr, 0, 0, 0, 1.0f);
    bind(T_ENVDIFF, K_CUBE, diff_cube, diff_res, diff_res, 1);
    bind(T_ENVSPEC, K_CUBE, spec_cube, spec_res, spec_res, spec_mips);
    bind(T_BRDFLUT, K_LUT, lut, lut_w, lut_h, 1);
    BIND_FORWARD_RESOURCES(fwd)
    load_per_frame(hl::fwd::cbPerFrame, *pf); load_per_view(hl::fwd::cbPerView, *pv);
    VqMaterialData mat{};
    mat.uvScaleOffset.x = 1.0f; mat.uvScaleOffset.y = 1.0f; mat.alpha = 1.0f;
    g_ddx[0] = g_ddx[1] = g_ddy[0] = g_ddy[1] = 0.0f;
    for (int y = row_begin; y < row_end; ++y)
        for (int x = 0; x < width; ++x) {
            const size_t o = ((size_t)y * width + x) * 4;
            mat.diffuse = VqFloat3{albedo_metalness[o], albedo_metalness[o + 1], albedo_metalness[o + 2]};
            mat.metalness = albedo_metalness[o + 3];
            mat.roughness = normal_roughness[o + 3];
            if (emissive) { mat.emissiveColor = VqFloat3{emissive[o], emissive[o + 1], emissive[o + 2]}; mat.emissiveIntensity = emissive[o + 3]; }
            load_material(hl::fwd::cbPerObject.materialData, mat);
            hl::fwd::cbPerFrame.fAmbientLightingFactor = position_ao[o + 3];
            hl::fwd::PSInput In;
            In.position = hl::float4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
            In.WorldSpacePosition = hl::float3(position_ao[o], position_ao[o + 1], position_ao[o + 2]);
            In.WorldSpaceNormal = hl::float3(normal_roughness[o], normal_roughness[o + 1], normal_roughness[o + 2]);
            In.WorldSpaceTangent = hl::float3(1.0f, 0.0f, 0.0f);
            In.uv = hl::float2(0.0f, 0.0f);
            const auto r = hl::fwd::PSMain(In);
            out_color[o] = r.color.x; out_color[o + 1] = r.color.y; out_color[o + 2] = r.color.z; out_color[o + 3] = r.color.w;
        }
}

// ---- scalar probes into BRDF.hlsl / ShadingMath.hlsl / Lighting.hlsl (compiled inside ForwardLighting.hlsl) ---------
static hl::float3 v3(const float* p) { return hl::float3(p[0], p[1], p[2]); }
static void o3(float* o, const hl::float3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static hl::fwd::BRDF_Surface surf(const float N[3], const float albedo[3], float roughness, float metalness) {
    hl::fwd::BRDF_Surface s{}; s.N = v3(N); s.diffuseColor = v3(albedo); s.roughness = roughness; s.metalness = metalness; return s;
}
float hlslref_ndf_ggx(float NdotH, float roughness) { return hl::fwd::NormalDistributionGGX(NdotH, roughness); }
float hlslref_geometry_smith(const float N[3], const float V[3], const float L[3], float k) { return hl::fwd::Geometry_Smith(v3(N), v3(V), v3(L), k); }
void hlslref_fresnel_schlick(const float H[3], const float V[3], const float F0[3], float out[3]) { o3(out, hl::fwd::Fresnel_Schlick(v3(H), v3(V), v3(F0))); }
void hlslref_fresnel_gaussian(const float H[3], const float V[3], const float F0[3], float out[3]) { o3(out, hl::fwd::Fresnel_Gaussian(v3(H), v3(V), v3(F0))); }
void hlslref_brdf(const float N[3], const float V[3], const float Wi[3], const float albedo[3], float roughness, float metalness, float out[3]) {
    o3(out, hl::fwd::BRDF(surf(N, albedo, roughness, metalness), v3(Wi), v3(V)));
}
void hlslref_environment_brdf(float NdotV, float roughness, float metallic, const float diffuseColor[3], const float diffuseIrradiance[3],
                              const float preFilteredSpecular[3], const float F0ScaleBias[2], float out[3]) {
    o3(out, hl::fwd::EnvironmentBRDF(NdotV, roughness, metallic, v3(diffuseColor), v3(diffuseIrradiance), v3(preFilteredSpecular),
                                     hl::float2(F0ScaleBias[0], F0ScaleBias[1])));
}
void hlslref_point_light(const VqPointLight* l, const float P[3], const float N[3], const float V[3], const float albedo[3],
                         float roughness, float metalness, float out[3]) {
    hl::fwd::PointLight pl; cp(pl, *l);
    o3(out, hl::fwd::CalculatePointLightIllumination(pl, surf(N, albedo, roughness, metalness), v3(P), v3(V)));
}
void hlslref_spot_light(const VqSpotLight* l, const float P[3], const float N[3], const float V[3], const float albedo[3],
                        float roughness, float metalness, float out[3]) {
    hl::fwd::SpotLight sl; cp(sl, *l);
    o3(out, hl::fwd::CalculateSpotLightIllumination(sl, surf(N, albedo, roughness, metalness), v3(P), v3(V)));
}
void hlslref_directional_light(const VqDirectionalLight* l, const float N[3], const float V[3], const float albedo[3],
                               float roughness, float metalness, float out[3]) {
    hl::fwd::DirectionalLight dl;
    dl.lightDirection = f3(l->lightDirection); dl.brightness = l->brightness; dl.color = f3(l->color); dl.depthBias = l->depthBias;
    dl.shadowing = l->shadowing; dl.enabled = l->enabled;
    o3(out, hl::fwd::CalculateDirectionalLightIllumination(dl, surf(N, albedo, roughness, metalness), v3(V)));
}
float hlslref_spotlight_intensity(const VqSpotLight* l, const float P[3]) { hl::fwd::SpotLight sl; cp(sl, *l); return hl::fwd::SpotlightIntensity(sl, v3(P)); }
void hlslref_hammersley(uint32_t i, uint32_t n, float out[2]) { const hl::float2 h = hl::fwd::Hammersley(i, n); out[0] = h.x; out[1] = h.y; }
void hlslref_importance_sample_ggx(const float Xi[2], const float N[3], float roughness, float out[3]) {
    o3(out, hl::fwd::ImportanceSampleGGX(hl::float2(Xi[0], Xi[1]), v3(N), roughness));
}
void hlslref_integrate_brdf(float NdotV, float roughness, int samples, float out[2]) {
    const hl::float2 r = hl::fwd::IntegrateBRDF(NdotV, roughness, samples); out[0] = r.x; out[1] = r.y;
}
void hlslref_direction_to_equirect_uv(const float d[3], float out[2]) { const hl::float2 r = hl::fwd::DirectionToEquirectUV(v3(d)); out[0] = r.x; out[1] = r.y; }
void hlslref_unpack_normal(const float sampled[3], const float n[3], const float t[3], float out[3]) { o3(out, hl::fwd::UnpackNormal(v3(sampled), v3(n), v3(t))); }
void hlslref_srgb_to_linear_pow22(const float c[3], float out[3]) { o3(out, hl::fwd::SRGBToLinear(v3(c))); }

// ---- Tonemapper.hlsl CSMain on one pixel ---------------------------------------------------------------------------
void hlslref_tonemap_pixel(const VqTonemapperParams* p, const float in[4], float out[4]) {
    install();
    hl::tonemap::ContentColorSpaceEnum = p->ContentColorSpace; hl::tonemap::OutputDisplayCurveEnum = p->OutputDisplayCurve;
    hl::tonemap::DisplayReferenceBrightnessLevel = p->DisplayReferenceBrightnessLevel; hl::tonemap::ToggleGammaCorrection = p->ToggleGammaCorrection;
    hl::tonemap::texColorInput.id = T_IN; hl::tonemap::texColorOutput.id = T_OUT;
    bind(T_IN, K_IMAGE4, in, 1, 1, 1); bind(T_OUT, K_IMAGE4, nullptr, 1, 1, 1, 0.0f, out);
    hl::tonemap::CSMain(hl::uint3(0u), hl::uint3(0u), hl::uint3(0u));
}
void hlslref_hdr_curves(const float c[3], float lin_to_srgb[3], float srgb_to_lin[3], float r709_to_2020[3], float r2020_to_709[3], float st2084[3]) {
    o3(lin_to_srgb, hl::tonemap::LinearToSRGB(v3(c))); o3(srgb_to_lin, hl::tonemap::SRGBToLinear(v3(c)));
    o3(r709_to_2020, hl::tonemap::Rec709ToRec2020(v3(c))); o3(r2020_to_709, hl::tonemap::Rec2020ToRec709(v3(c)));
    o3(st2084, hl::tonemap::LinearToST2084(v3(c)));
}

// ---- GaussianBlur.hlsl CSMain_X / CSMain_Y over an image ------------------------------------------------------------
void hlslref_gaussian_blur(const float* in, float* out, int w, int h, int vertical) {
    install();
    hl::blur::iImageSize = hl::int2(w, h);
    hl::blur::texColorInput.id = T_IN; hl::blur::texColorOutput.id = T_OUT;
    bind(T_IN, K_IMAGE4, in, w, h, 1); bind(T_OUT, K_IMAGE4, nullptr, w, h, 1, 0.0f, out);
    for (uint32_t y = 0; y < (uint32_t)h; ++y)
        for (uint32_t x = 0; x < (uint32_t)w; ++x) {
            const hl::uint3 id(x, y, 0u);
            if (vertical) hl::blur::CSMain_Y(hl::uint3(x & 7u, y & 7u, 0u), hl::uint3(x >> 3, y >> 3, 0u), id);
            else          hl::blur::CSMain_X(hl::uint3(x & 7u, y & 7u, 0u), hl::uint3(x >> 3, y >> 3, 0u), id);
        }
}

// ---- CubemapConvolution.hlsl --------------------------------------------------------------------------------------
void hlslref_diffuse_irradiance_texel(const float* pyramid, int w, int h, int levels, const float dir[3], float out[4]) {
    install();
    hl::conv::texEquirectEnvironmentMap.id = T_HDRI; bind(T_HDRI, K_EQUIRECT, pyramid, w, h, levels);
    hl::conv::GSOut In{}; In.CubemapLookDirection = v3(dir);
    const hl::float4 r = hl::conv::PSMain_DiffuseIrradiance(In); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void hlslref_specular_irradiance_texel(const float* pyramid, int w, int h, int levels, const float dir[3], float roughness,
                                       float dim_x, float dim_y, float out[4]) {
    install();
    hl::conv::texEquirectEnvironmentMap.id = T_HDRI; bind(T_HDRI, K_EQUIRECT, pyramid, w, h, levels);
    hl::conv::Roughness = roughness; hl::conv::TextureDimensionsLOD0 = hl::float2(dim_x, dim_y);
    hl::conv::GSOut In{}; In.CubemapLookDirection = v3(dir);
    const hl::float4 r = hl::conv::PSMain_SpecularIrradiance(In); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
// CSMain_BRDFIntegration writes texel (x, y) of the 1024^2 LUT (the size and the 2048 samples are constants in the shader)
void hlslref_brdf_lut_texel(int x, int y, float out[2]) {
    install();
    static std::vector<float> lut;                       // sparse use: only the requested texel is read back
    lut.assign((size_t)1024 * 1024 * 2, 0.0f);
    hl::conv::texBRDFLUT.id = T_OUT; bind(T_OUT, K_IMAGE2, nullptr, 1024, 1024, 1, 0.0f, lut.data());
    hl::conv::CSMain_BRDFIntegration(hl::uint3((uint32_t)x, (uint32_t)y, 0u));
    out[0] = lut[((size_t)y * 1024 + x) * 2]; out[1] = lut[((size_t)y * 1024 + x) * 2 + 1];
}

// ---- Skydome.hlsl PSMain ------------------------------------------------------------------------------------------
void hlslref_skydome_pixel(const float* pyramid, int w, int h, int levels, const float look_dir[3], float out[4]) {
    install();
    hl::sky::texEquirectEnvironmentMap.id = T_HDRI; bind(T_HDRI, K_EQUIRECT, pyramid, w, h, levels);
    hl::sky::PSInput In{}; In.CubemapLookDirection = v3(look_dir);
    const hl::float4 r = hl::sky::PSMain(In); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---- ApplyReflections.hlsl CSMain, both permutations, in place on `scene` ------------------------------------------
void hlslref_apply_reflections(float* scene, const float* reflection, const float* bounding_volumes, int w, int h) {
    install();
    bind(T_IN, K_IMAGE4, reflection, w, h, 1); bind(T_IN2, K_IMAGE4, bounding_volumes, w, h, 1);
    bind(T_OUT, K_IMAGE4, scene, w, h, 1, 0.0f, scene);
    hl::refl::TexReflectionRadiance.id = T_IN; hl::refl::TexSceneColor.id = T_OUT;
    hl::reflbv::TexReflectionRadiance.id = T_IN; hl::reflbv::TexBoundingVolumes.id = T_IN2; hl::reflbv::TexSceneColor.id = T_OUT;
    for (uint32_t y = 0; y < (uint32_t)h; ++y)
        for (uint32_t x = 0; x < (uint32_t)w; ++x) {
            const hl::uint3 id(x, y, 0u), z(0u);
            if (bounding_volumes) hl::reflbv::CSMain(z, z, id); else hl::refl::CSMain(z, z, id);
        }
}


// ---- AMDFidelityFX.hlsl: CAS / FSR1 EASU / FSR1 RCAS compute entry points, dispatched like the engine does ----------
// (PostProcess.cpp: one 64-thread group per 16x16 output tile); outputs are float3 UAVs, alpha of `out` is left as is.
void hlslref_cas(const uint32_t c[8], const float* in, float* out, int w, int h) {
    install();
    hl::cas::CASConst0 = hl::uint4(c[0], c[1], c[2], c[3]); hl::cas::CASConst1 = hl::uint4(c[4], c[5], c[6], c[7]);
    hl::cas::CASInputTexture.id = T_IN; hl::cas::CASOutputTexture.id = T_OUT;
    bind(T_IN, K_IMAGE4, in, w, h, 1); bind(T_OUT, K_IMAGE3IN4, nullptr, w, h, 1, 0.0f, out);
    dispatch16(w, h, hl::cas::CAS_CSMain);
}
void hlslref_fsr_easu(const uint32_t c[16], const float* in, int in_w, int in_h, float* out, int out_w, int out_h) {
    install();
    hl::easu::FSRConst0 = hl::uint4(c[0], c[1], c[2], c[3]); hl::easu::FSRConst1 = hl::uint4(c[4], c[5], c[6], c[7]);
    hl::easu::FSRConst2 = hl::uint4(c[8], c[9], c[10], c[11]); hl::easu::FSRConst3 = hl::uint4(c[12], c[13], c[14], c[15]);
    hl::easu::FSRInputTexture.id = T_IN; hl::easu::FSROutputTexture.id = T_OUT;
    bind(T_IN, K_IMAGE4, in, in_w, in_h, 1); bind(T_OUT, K_IMAGE3IN4, nullptr, out_w, out_h, 1, 0.0f, out);
    dispatch16(out_w, out_h, hl::easu::FSR_EASU_CSMain);
}
void hlslref_fsr_rcas(const uint32_t c[4], const float* in, float* out, int w, int h) {
    install();
    hl::rcas::RCASConst0 = hl::uint4(c[0], c[1], c[2], c[3]);
    hl::rcas::RCASInputTexture.id = T_IN; hl::rcas::RCASOutputTexture.id = T_OUT;
    bind(T_IN, K_IMAGE4, in, w, h, 1); bind(T_OUT, K_IMAGE3IN4, nullptr, w, h, 1, 0.0f, out);
    dispatch16(w, h, hl::rcas::FSR_RCAS_CSMain);
}


// ---- FidelityFX SPD through the engine's two wrappers ----------------------------------------------------------------
// levels: packed float4 mip chain, level 0 = the source (w x h), level l = max(1, w>>l) x max(1, h>>l); the shader writes
// levels 1..mips. Constants as SpdSetup computes them for the full rectangle (oracle / libffxref).
void hlslref_spd_downsample(float* levels, int w, int h, int n_levels, uint32_t mips, uint32_t num_work_groups,
                            uint32_t wg_off_x, uint32_t wg_off_y, uint32_t dispatch_x, uint32_t dispatch_y) {
    install();
    hl::spd::mips = mips; hl::spd::numWorkGroups = num_work_groups; hl::spd::workGroupOffset = hl::uint2(wg_off_x, wg_off_y);
    static hl::spd::SpdGlobalAtomicBuffer counter; counter = hl::spd::SpdGlobalAtomicBuffer{};
    hl::spd::spdGlobalAtomic.data = &counter;
    size_t off = 0;
    for (int l = 0; l < 13; ++l) {
        const int id = 32 + l;
        hl::spd::imgDst[l].id = id;
        if (l < n_levels) {
            const int lw = (w >> l) > 0 ? (w >> l) : 1, lh = (h >> l) > 0 ? (h >> l) : 1;
            bind(id, K_IMAGE4, levels + off * 4, lw, lh, 1, 0.0f, levels + off * 4);
            off += (size_t)lw * lh;
        } else g_bind[id] = Binding{};
    }
    hl::spd::imgDst6.id = 32 + 6;
    for (uint32_t gy = 0; gy < dispatch_y; ++gy)
        for (uint32_t gx = 0; gx < dispatch_x; ++gx)
            run_workgroup(256, [&](int lane) { hl::spd::SPD_CSMain(hl::uint3(gx, gy, 0u), (uint32_t)lane); });
}
// DownsampleDepth.hlsl CSMain: level 0 = copy of the depth buffer, levels 1.. = MIN pyramid. One 32x8 group per 64x64 texels.
void hlslref_depth_pyramid(const float* depth, int w, int h, float* levels, int n_levels) {
    install();
    hl::depth::uImageDimensionsXY = hl::int2(w, h);
    static uint32_t counter; counter = 0;
    hl::depth::g_global_atomic.data = &counter;
    hl::depth::g_depth_buffer.id = T_IN; bind(T_IN, K_IMAGE1, depth, w, h, 1);
    size_t off = 0;
    for (int l = 0; l < 13; ++l) {
        const int id = 32 + l;
        hl::depth::g_downsampled_depth_buffer[l].id = id;
        if (l < n_levels) {
            const int lw = (w >> l) > 0 ? (w >> l) : 1, lh = (h >> l) > 0 ? (h >> l) : 1;
            bind(id, K_IMAGE1, levels + off, lw, lh, 1, 0.0f, levels + off);
            off += (size_t)lw * lh;
        } else g_bind[id] = Binding{};
    }
    const uint32_t gxn = (uint32_t)((w + 63) / 64), gyn = (uint32_t)((h + 63) / 64);
    for (uint32_t gy = 0; gy < gyn; ++gy)
        for (uint32_t gx = 0; gx < gxn; ++gx)
            run_workgroup(256, [&](int lane) {
                const uint32_t tx = (uint32_t)lane & 31u, ty = (uint32_t)lane >> 5;
                hl::depth::CSMain(hl::uint3(gx * 32u + tx, gy * 8u + ty, 0u), hl::uint3(gx, gy, 0u), (uint32_t)lane);
            });
}

}  // extern "C"
