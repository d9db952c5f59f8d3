/*
 * vqcuda.h — C-ABI of the B200 (sm_100a) headless shading backend.
 *
 * One entry point per GPU program the engine dispatches for its per-pixel passes.
 * The reference has no FFI of its own; each function below names
 * the reference dispatch site + HLSL entry point it replaces. Conventions mirror the
 * engine's (SURVEY.md §8(b) "Conventions"):
 *   - resources are caller-owned (the engine passes integer IDs; here: device pointers
 *     described by {ptr, width, height, pitch}); the library never frees caller memory;
 *   - nothing throws across the boundary: every call returns 0 (VQ_OK) or a negative
 *     VqStatus; vq_last_error() gives a thread-local message;
 *   - calls only ENQUEUE work on `stream` (== recording into a command list); they are
 *     thread-safe when callers use distinct streams (== per-thread command lists,
 *     SceneRendering.cpp:197-207): host-side context state is guarded by a lock and every
 *     SPD launch gets its own ticket word. EXCEPTIONS — calls whose DEVICE scratch is one
 *     buffer per context and which therefore must not be in flight concurrently on two
 *     streams of the same context (use one context per thread for these, or order them):
 *       vq_forward_lighting* with an environment that is NOT registered by
 *       vq_environment_prepare (the sampling copies are rebuilt into per-context scratch),
 *       vq_environment_prepare itself, vq_image_resize, vq_depth_min_pyramid,
 *       vq_forward_lighting_shadowed (per-pixel PCF records),
 *       vq_forward_lighting_host.
 *     Synchronisation is the caller's (cudaStreamSynchronize == fence wait), except the
 *     *_host convenience calls which block.
 *   - there is NO CPU fallback: without a CUDA device every device call fails with
 *     VQ_ERR_NO_DEVICE.
 *
 * All pixel data is linear fp32: RGBA32F (`float4`, 16 B) unless stated otherwise.
 * `stream` is a cudaStream_t passed as void* so that this header needs no CUDA headers.
 */
#ifndef VQCUDA_H
#define VQCUDA_H

#include "vq_shader_data.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#  define VQ_API __declspec(dllexport)
#else
#  define VQ_API __attribute__((visibility("default")))
#endif

typedef enum VqStatus {
    VQ_OK               =  0,
    VQ_ERR_INVALID_ARG  = -1,
    VQ_ERR_CUDA         = -2,
    VQ_ERR_UNSUPPORTED  = -3,
    VQ_ERR_NO_DEVICE    = -4,
    VQ_ERR_OUT_OF_MEMORY= -5
} VqStatus;

/* ------------------------------------------------------------------------------------------
 * Resource descriptors (replace TextureID/SRV_ID/UAV_ID/RTV_ID, EnvironmentMapRendering.h:28-48)
 * ------------------------------------------------------------------------------------------ */

/* 2-D image of float4 texels (float2 where stated). pitch_bytes is a multiple of 16. */
typedef struct VqImage {
    void*   ptr;
    int32_t width, height;
    size_t  pitch_bytes;
} VqImage;

/* Cubemap with a mip chain, float4 texels, tightly packed, mip-major / face-minor:
 *   offset_texels(mip, face) = sum_{m<mip} 6*(res>>m)^2 + face*(res>>mip)^2
 * Face order = D3D / CubemapUtility.h:31-41: +X(RIGHT) -X(LEFT) +Y(UP) -Y(DOWN) +Z(FRONT) -Z(BACK). */
typedef struct VqCubemap {
    void*   ptr;
    int32_t res;   /* edge length of mip 0 */
    int32_t mips;
} VqCubemap;

/* Equirectangular HDRI with its CPU-style mip pyramid (TextureManager.cpp:714-727), float4,
 * tightly packed levels: level l is (width>>l) x (height>>l); levels <= vq_mip_level_count(). */
typedef struct VqPyramid {
    void*   ptr;
    int32_t width, height;
    int32_t levels;
} VqPyramid;

/* The surface record PSMain consumes after material-texture sampling (BRDF.hlsl:50-58,
 * ForwardLighting.hlsl:245-285) as three (optionally four) float4 planes:
 *   position_ao      = { P.xyz (world space), ao }   ao = fAmbientLightingFactor*localAO*SSAO
 *   normal_roughness = { N.xyz (unit, world space), roughness }
 *   albedo_metalness = { diffuseColor.rgb, metalness }
 *   emissive         = { emissiveColor.rgb, emissiveIntensity }   (ptr == NULL -> no emissive) */
typedef struct VqGBuffer {
    VqImage position_ao;
    VqImage normal_roughness;
    VqImage albedo_metalness;
    VqImage emissive;
} VqGBuffer;

/* The three IBL inputs PSMain binds (SceneRendering.cpp:1690-1717). brdf_lut holds float2 texels. */
typedef struct VqEnvironmentMaps {
    VqCubemap irradiance_diffuse;   /* blurred diffuse irradiance, 1 mip */
    VqCubemap irradiance_specular;  /* prefiltered specular, `mips` levels */
    VqImage   brdf_lut;             /* float2 (scale,bias), CubemapConvolution.hlsl:227-240 */
} VqEnvironmentMaps;

typedef struct VqContext VqContext;

/* ------------------------------------------------------------------------------------------
 * Lifetime  (IRenderPass::Initialize / Destroy / OnCreateWindowSizeDependentResources,
 *            RenderPass.h:44-59)
 * ------------------------------------------------------------------------------------------ */
VQ_API int  vq_ctx_create(int device, VqContext** out_ctx);
VQ_API int  vq_ctx_destroy(VqContext* ctx);
/* Pre-sizes internal scratch (SPD counters, host-call staging) for a render resolution. Optional. */
VQ_API int  vq_ctx_resize(VqContext* ctx, int width, int height);
VQ_API const char* vq_last_error(void);
VQ_API const char* vq_version(void);
/* Number of kernels this library has launched in this process (all contexts). */
VQ_API uint64_t vq_launch_count(void);

/* End-of-pass rendezvous of a multi-GPU step, run by the kernel itself (its last CTA): after all of this rank's stores
 * (local and peer) are ordered, it writes `epoch` into word `my_index` of every OTHER rank's flag array and then waits
 * until words j != my_index of its OWN array have reached `epoch` — i.e. until every peer has finished storing into this
 * rank's buffers. flags[0] = this rank's array, flags[1..n_ranks-1] = the peers' arrays mapped into this process
 * (n_ranks 32-bit words each, zero-initialised, e.g. in torch symmetric memory); epoch must increase from call to call.
 * Replaces a host-issued barrier per step: the step is ONE kernel. */
typedef struct VqPeerSignal {
    uint32_t* flags[8];
    int32_t   n_ranks;
    int32_t   my_index;
    uint32_t  epoch;
} VqPeerSignal;

/* ------------------------------------------------------------------------------------------
 * K1  Forward PBR lighting.  Replaces VQRenderer::RenderSceneColor (SceneRendering.cpp:1619-1851)
 *     + PSMain (ForwardLighting.hlsl:222-391) over a G-buffer instead of rasterised draws.
 *     Shades rows [row_begin,row_end) of the image (row tiling for multi-GPU); out = {I.rgb, roughness}.
 *     This entry point binds no shadow maps: caster lists and a shadowing directional light are lit with shadow
 *     factor 1 (no occlusion) - which differs from PSMain whenever the engine has shadow maps bound; use
 *     vq_forward_lighting_shadowed (below) for the PCF-shadowed result.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_forward_lighting(VqContext* ctx,
                               const VqPerFrameData* per_frame,
                               const VqPerViewLightingData* per_view,
                               const VqGBuffer* gbuffer,
                               const VqEnvironmentMaps* env,
                               VqImage out_color,
                               int row_begin, int row_end,
                               void* stream);

/* K1 fused with the gather of tiles (multi-GPU): shades rows [row_begin,row_end) of the local G-buffer tile and stores every
 * pixel to row (dst_row_offset + y) of EACH of the n_outs (<= 8) destination frames. Destinations other than the local
 * frame are peer-GPU buffers mapped into this process (CUDA IPC / torch symmetric memory): the stores travel over NVLink
 * while the SMs keep shading, so the all-gather of tiles overlaps the math instead of following it. The caller provides
 * the cross-rank barrier after the kernel (e.g. symmetric-memory barrier or a 1-element all-reduce). */
VQ_API int vq_forward_lighting_multi(VqContext* ctx,
                                     const VqPerFrameData* per_frame,
                                     const VqPerViewLightingData* per_view,
                                     const VqGBuffer* gbuffer,
                                     const VqEnvironmentMaps* env,
                                     const VqImage* out_frames, int n_outs, int dst_row_offset,
                                     int row_begin, int row_end,
                                     void* stream);
/* ... and with the cross-rank rendezvous run by the kernel's last CTA (VqPeerSignal above): shade + gather + barrier = ONE kernel. */
VQ_API int vq_forward_lighting_multi_signal(VqContext* ctx, const VqPerFrameData* per_frame, const VqPerViewLightingData* per_view,
                                            const VqGBuffer* gbuffer, const VqEnvironmentMaps* env, const VqImage* outs, int n_outs,
                                            int dst_row_offset, int row_begin, int row_end, const VqPeerSignal* signal, void* stream);

/* The forward pass samples the IBL cubemaps from bordered copies (each face carries a 1-texel border of its
 * neighbours, so seamless bilinear taps never leave the face). vq_environment_prepare builds those copies once and
 * registers them in the context — the analogue of the RENDER_TARGET -> PIXEL_SHADER_RESOURCE barrier the engine
 * records after prefiltering (EnvironmentMapRendering.cpp:466-472). Call it again whenever the maps' contents
 * change, or vq_environment_invalidate to drop the registration. Without it vq_forward_lighting re-pads the cubes
 * into context scratch on every call (always correct; serialise such calls per context). */
VQ_API int vq_environment_prepare(VqContext* ctx, const VqEnvironmentMaps* env, void* stream);
VQ_API int vq_environment_invalidate(VqContext* ctx);

/* ------------------------------------------------------------------------------------------
 * K11 HDRI mip pyramid, 2x2 MIN filter, alpha = 1.  Replaces VQ_DXGI_UTILS::MipImage
 *     (DXGIUtils.cpp:289-317) called from TextureManager.cpp:714-727. Level 0 must be filled.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_hdri_build_mips(VqContext* ctx, VqPyramid hdri, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  Diffuse irradiance convolution.  Replaces the per-face draws at
 *     EnvironmentMapRendering.cpp:221-240 + PSMain_DiffuseIrradiance (CubemapConvolution.hlsl:112-163).
 *     Two sampling modes:
 *       step  > 0 : the reference's float-accumulated (phi,theta) loops with that step
 *                   (INTEGRATION_STEP_DIFFUSE_IRRADIANCE; 0.010/0.025/0.050, PipelineStateObjects.cpp:1298-1306)
 *       step == 0 : an integer n_phi x n_theta grid, phi_i = i*(2pi/n_phi), theta_j = j*(pi/2/n_theta)
 *     Source level is HDRI mip `src_mip` (reference: 3). Computes texel rows [row_begin,row_end)
 *     of the flattened (face, row) space, face*res + row; out.mips must be 1.
 * ------------------------------------------------------------------------------------------ */
typedef struct VqDiffuseIrradianceParams {
    float   step;
    int32_t n_phi, n_theta;
    int32_t src_mip;
} VqDiffuseIrradianceParams;
VQ_API int vq_diffuse_irradiance(VqContext* ctx, const VqDiffuseIrradianceParams* params,
                                 VqPyramid hdri, VqCubemap out,
                                 int row_begin, int row_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  Specular prefilter (GGX importance sampling, 512 Hammersley samples, PDF-based source mip).
 *     Replaces the mip x face draw loop at EnvironmentMapRendering.cpp:413-465 +
 *     PSMain_SpecularIrradiance (CubemapConvolution.hlsl:168-223). For every mip m of `out`:
 *     Roughness = m/(mips-1), ViewDim = HDRI dims (EnvironmentMapRendering.cpp:432-434).
 *     Computes texel rows [row_begin,row_end) of the flattened (mip, face, row) space
 *     (vq_cubemap_row_count rows in total).
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_specular_prefilter(VqContext* ctx, VqPyramid hdri, VqCubemap out,
                                 int num_samples, int row_begin, int row_end, void* stream);
/* K3 fused with the gather of row blocks (multi-GPU strong scaling): every texel of rows [row_begin,row_end) is stored into
 * each of the n_outs (<= 8) cubemaps of identical shape — the local one first, then the other ranks' buffers mapped into
 * this process (CUDA IPC / torch symmetric memory). The stores travel over NVLink while the SMs keep integrating; the caller
 * provides the cross-rank barrier after the kernel(s). */
VQ_API int vq_specular_prefilter_multi(VqContext* ctx, VqPyramid hdri, const VqCubemap* outs, int n_outs,
                                       int num_samples, int row_begin, int row_end, void* stream);


/* K3 as ONE persistent launch over several row ranges (n_ranges pairs [begin,end), increasing and disjoint) of the
 * flattened (mip, face, row) space, stored into n_outs (<= 8) cubemaps, with the optional rendezvous above (NULL: none).
 * This is what one rank of the strong-scaled prefilter runs per step: its row block of every mip, the replicated tail
 * mips, the fused gather over NVLink and the cross-rank barrier, in one kernel. */
VQ_API int vq_specular_prefilter_ranges(VqContext* ctx, VqPyramid hdri, const VqCubemap* outs, int n_outs, int num_samples,
                                        const int* row_ranges, int n_ranges, const VqPeerSignal* signal, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  BRDF integration LUT (split-sum scale,bias).  Replaces ComputeBRDFIntegrationLUT
 *     (Renderer.cpp:871-909) + CSMain_BRDFIntegration (CubemapConvolution.hlsl:227-240).
 *     out holds float2 texels; NdotV = (x+.5)/W, roughness = (y+.5)/H; reference: 1024^2, 2048 samples.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_brdf_integration_lut(VqContext* ctx, VqImage out_rg, int num_samples,
                                   int row_begin, int row_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  Separable 21-tap Gaussian blur, clamp-to-edge, alpha = 1.  Replaces the dispatches at
 *     EnvironmentMapRendering.cpp:341-372 / SceneRendering.cpp:2582-2638 + CSMain_X / CSMain_Y
 *     (GaussianBlur.hlsl:119-186).
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_gaussian_blur_x(VqContext* ctx, const VqBlurParams* params, VqImage in, VqImage out, void* stream);
VQ_API int vq_gaussian_blur_y(VqContext* ctx, const VqBlurParams* params, VqImage in, VqImage out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K6  Tonemapper.  Replaces the dispatch at SceneRendering.cpp:2640-2656 + CSMain (Tonemapper.hlsl:110-151).
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_tonemap(VqContext* ctx, const VqTonemapperParams* params, VqImage in, VqImage out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7  FidelityFX CAS, sharpen-only path.  Replaces SceneRendering.cpp:2658-2693 + CAS_CSMain
 *     (AMDFidelityFX.hlsl:122-174) -> CasFilter(noScaling) (CAS/ffx_cas.h:408-537).
 *     cas_const = the 8 words CasSetup wrote (FFFXCAS::CASConstantBlock, PostProcess.h:104-112).
 *     Reads outside the image return 0 (D3D Texture.Load). Output alpha is 1.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_cas(VqContext* ctx, const uint32_t cas_const[8], VqImage in, VqImage out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K8  FSR1 EASU.  Replaces SceneRendering.cpp:2711-2734 + FSR_EASU_CSMain (AMDFidelityFX.hlsl:247-288)
 *     -> FsrEasuF (FSR1.0/ffx_fsr1.h:315-437). easu_const = FFSR1_EASU::EASUConstantBlock (16 words).
 *     Gathers use the engine's sampler addressing: VQ_ADDRESS_WRAP (RootSignatures.cpp:525) by default.
 * ------------------------------------------------------------------------------------------ */
enum { VQ_ADDRESS_WRAP = 0, VQ_ADDRESS_CLAMP = 1 };
VQ_API int vq_fsr_easu(VqContext* ctx, const uint32_t easu_const[16], int address_mode,
                       VqImage in, VqImage out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K9  FSR1 RCAS.  Replaces SceneRendering.cpp:2735-2781 + FSR_RCAS_CSMain (AMDFidelityFX.hlsl:335-376)
 *     -> FsrRcasF (ffx_fsr1.h:684-769). rcas_const = FFSR1_RCAS::RCASConstantBlock (4 words).
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_fsr_rcas(VqContext* ctx, const uint32_t rcas_const[4], VqImage in, VqImage out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K10 FidelityFX SPD: the whole mip chain in one launch.  Replaces SPD_CSMain
 *     (AMDFidelityFX.hlsl:525-545) -> SpdDownsample (SPD/ffx_spd.h:811-835), reduction =
 *     (v0+v1+v2+v3)*0.25 (AMDFidelityFX.hlsl:463-466). `mips` are the destination levels
 *     1..constants->mips (mips[i] is level i+1, floor-halved sizes); src is level 0.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_spd_downsample(VqContext* ctx, const VqSpdConstants* constants,
                             VqImage src, const VqImage* mips, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-side constant setup == the FidelityFX `A_CPU` functions the engine calls
 * (PostProcess.cpp:39-99): bit-identical results.
 * ------------------------------------------------------------------------------------------ */
/* FsrEasuCon (ffx_fsr1.h:156-202) -> con[16] = con0|con1|con2|con3 */
VQ_API void vq_fsr_easu_con(uint32_t con[16],
                            float input_viewport_w, float input_viewport_h,
                            float input_size_w, float input_size_h,
                            float output_w, float output_h);
/* FsrRcasCon (ffx_fsr1.h:662-672) */
VQ_API void vq_fsr_rcas_con(uint32_t con[4], float sharpness_stops);
/* CasSetup (ffx_cas.h:375-394) -> con[8] = const0|const1 */
VQ_API void vq_cas_setup(uint32_t con[8], float sharpness,
                         float input_w, float input_h, float output_w, float output_h);
/* SpdSetup (ffx_spd.h:327-351); rect = {left, top, width, height}; mips < 0 -> derive from rect */
VQ_API void vq_spd_setup(uint32_t dispatch_xy[2], VqSpdConstants* constants,
                         const uint32_t rect[4], int mips);
/* Image::CalculateMipLevelCount (Libs/VQUtils/Source/Image.cpp:231-241) */
VQ_API int  vq_mip_level_count(uint64_t w, uint64_t h);

/* Layout helpers for the packed descriptors above (host-side arithmetic only). */
VQ_API uint64_t vq_cubemap_texel_count(int res, int mips);             /* total float4 texels */
VQ_API uint64_t vq_cubemap_offset(int res, int mip, int face);         /* texel offset of (mip,face) */
VQ_API int      vq_cubemap_row_count(int res, int mips);               /* rows in flattened (mip,face,row) */
VQ_API uint64_t vq_pyramid_texel_count(int width, int height, int levels);
VQ_API uint64_t vq_pyramid_offset(int width, int height, int level);   /* texel offset of a level */

/* ------------------------------------------------------------------------------------------
 * SURVEY §8(f).1  Surface producer: the part of PSMain BEFORE lighting (ForwardLighting.hlsl:226-283) —
 *     material-texture sampling, sRGB->linear, normal mapping, Has*Map selection — turned into a kernel
 *     that fills the G-buffer K1 consumes from what a rasteriser interpolates per covered pixel.
 * ------------------------------------------------------------------------------------------ */

/* Material texture: RGBA8 UNORM (what Image::LoadFromFile + TextureManager upload, 4 B/texel), with the CPU-built
 * box-filter mip chain (TextureManager.cpp:714-727 -> DXGIUtils.cpp:250-287), tightly packed levels, level l is
 * (width>>l) x (height>>l), levels <= vq_mip_level_count(); texel offsets as vq_pyramid_offset().
 * ptr == NULL is the engine's null SRV (Renderer_Resources.cpp:383-387): every read returns 0. */
typedef struct VqTexture2D {
    void*   ptr;
    int32_t width, height;
    int32_t levels;
} VqTexture2D;

/* the descriptor table of one material, in the order PSMain samples them (ForwardLighting.hlsl:87-94,229-235; texAlphaMask t3 is bound but never sampled) */
typedef struct VqMaterialTextures {
    VqTexture2D diffuse, normals, emissive, metalness, roughness, occl_rough_metal, local_ao;
} VqMaterialTextures;

/* PSInput as a rasteriser would deliver it (ForwardLighting.hlsl:42-53), one record per pixel in three float4 planes:
 *   position_u = { WorldSpacePosition.xyz, uv.x }
 *   normal_v   = { WorldSpaceNormal.xyz  (interpolated, not normalised), uv.y }
 *   tangent_m  = { WorldSpaceTangent.xyz (interpolated, not normalised), material index as an exact small float }
 *   ssao       = optional R32F plane, the texScreenSpaceAO the shader point-samples (NULL -> 1) */
typedef struct VqSurfaceInputs {
    VqImage position_u, normal_v, tangent_m;
    VqImage ssao;
} VqSurfaceInputs;

typedef struct VqMaterialTable VqMaterialTable;   /* device-resident copy of the scene's materials */

/* RGBA8 mip chain, per-channel (a+b+c+d)/4 truncating: VQ_DXGI_UTILS::MipImage, 4-byte branch (DXGIUtils.cpp:250-287).
 * Level 0 must be filled. Bit-exact. */
VQ_API int vq_texture_build_mips(VqContext* ctx, VqTexture2D tex, void* stream);

/* Uploads `count` materials (constants + texture descriptors, both HOST arrays; the texel pointers inside are DEVICE
 * pointers that must outlive the table). Material::GetCBufferData + the SRV table built at AssetLoader.cpp:406-420.
 * A material whose bound maps all have the same width, height and level count (>= 2 maps) also gets a SAMPLING COPY:
 * one 16-byte record per texel holding the bytes PSMain reads of every map, so a trilinear tap is one 128-bit load for
 * all maps (like the bordered cubemap copies of vq_environment_prepare). The copy is taken HERE, synchronously (the call
 * waits for the device first, so mip chains still being built on other streams are complete): fill the textures, mips
 * included, before creating the table, and re-create the table when texels change. Costs 16 B per texel of the shared
 * size; if that allocation fails, or with VQ_SURFACE_RECORDS=0 in the environment, the maps are sampled one by one. */
VQ_API int vq_material_table_create(VqContext* ctx, const VqMaterialData* materials,
                                    const VqMaterialTextures* textures, int count, VqMaterialTable** out_table);
VQ_API int vq_material_table_destroy(VqContext* ctx, VqMaterialTable* table);

/* Fills rows [row_begin,row_end) of the G-buffer (emissive plane optional). `ambient_factor` = PerFrameData
 * fAmbientLightingFactor (ForwardLighting.hlsl:245). alpha_mask != 0 compiles in the ENABLE_ALPHA_MASK discard
 * (ForwardLighting.hlsl:237-240): such pixels are left untouched in the G-buffer. Texture LOD comes from the 2x2-quad
 * finite differences of uv, as the implicit-derivative Sample() of a pixel shader does. */
VQ_API int vq_gbuffer_from_materials(VqContext* ctx, const VqSurfaceInputs* in, const VqMaterialTable* table,
                                     float ambient_factor, int alpha_mask, const VqGBuffer* out,
                                     int row_begin, int row_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY §8(f).2  On-disk format at either end of the path: Radiance .hdr (RGBE).
 *     Load  = Image::LoadFromFile -> stbi_loadf(path, &x, &y, &n, 4) (Libs/VQUtils/Source/Image.cpp:119-121) with
 *             Image::CalculateMaxLuminance (Image.cpp:43-86) fused into the same pass; results are bit-identical to stb's.
 *     Save  = Image::SaveToDisk -> stbi_write_hdr(path, x, y, 4, data) (Image.cpp:210-213); files are byte-identical.
 *     The byte-serial work (header text, locating / emitting run lists) is host code; the per-texel work is kernels,
 *     so host<->device traffic is the file image (<= ~4 B/texel), not the fp32 image (16 B/texel).
 * ------------------------------------------------------------------------------------------ */
typedef struct VqHdrInfo {
    int32_t  width, height;
    uint64_t data_offset;   /* first byte after the resolution line (or where flat data starts) */
    int32_t  flat;          /* 1: 4 bytes per texel from data_offset on, no run-length encoding */
    int32_t  reserved;
} VqHdrInfo;
/* HOST. Parses the header of a .hdr file image. channel_offsets == NULL: header only (dims for allocation).
 * Otherwise (run-length encoded files: info->flat == 0 after the first call) it must hold 4*height+1 entries and
 * receives the file offset of every scanline's R,G,B,E run list (+ the end offset); the run headers are validated
 * (stb's "corrupt HDR" conditions -> VQ_ERR_INVALID_ARG; a stream that ends inside a scanline is an error, where stb
 * would spin on zero bytes; a resolution line with a zero dimension is rejected, where stb returns an image without texels).
 * A file whose scanlines turn out not to be run-length encoded flips info->flat to 1. */
VQ_API int vq_hdr_parse(const void* file, uint64_t size, VqHdrInfo* info, uint64_t* channel_offsets);
/* DEVICE. Expands the file image into `out` (width x height RGBA32F, alpha 1). dev_file: 16-byte aligned, allocation
 * padded to a multiple of 16 bytes. dev_channel_offsets: device copy of the index (NULL for flat files).
 * dev_max_luminance (optional): receives max over texels of 0.2126 r + 0.7152 g + 0.0722 b (Image::MaxLuminance). */
VQ_API int vq_hdr_decode(VqContext* ctx, const void* dev_file, uint64_t size, const VqHdrInfo* info,
                         const uint64_t* dev_channel_offsets, VqImage out, float* dev_max_luminance, void* stream);
/* Blocking: host file image -> device fp32 image (parse + upload + decode). */
VQ_API int vq_hdr_load_host(VqContext* ctx, const void* host_file, uint64_t size, VqImage out, float* max_luminance);
/* DEVICE. RGBA32F -> RGBE texels (4 bytes each, tightly packed rows), stbiw__linear_to_rgbe. Inputs are radiance (>= 0). */
VQ_API int vq_hdr_encode_rgbe(VqContext* ctx, VqImage in, void* dev_rgbe, void* stream);
/* HOST. RGBE texels -> the .hdr file image (text header + per-scanline run lists). file == NULL: size query. */
VQ_API int vq_hdr_pack_file(const void* host_rgbe, int width, int height, void* file, uint64_t capacity, uint64_t* size);
/* Blocking: device fp32 image -> host file image (encode + download + pack). */
VQ_API int vq_hdr_save_host(VqContext* ctx, VqImage in, void* host_file, uint64_t capacity, uint64_t* size);

/* Image::CreateResizedImage (Image.cpp:148-190) -> stbir_resize_float(in, w, h, 0, out, W, H, 0, 4): the downsize the engine
 * applies to a hi-res HDRI before saving the smaller .hdr (EnvironmentMap.cpp:142-209; 8k -> 4k/2k/1k). Separable
 * Mitchell-Netravali, edge clamp, weights normalised per output sample; bit-identical to the vendored stb_image_resize
 * v0.96. out.width <= in.width and out.height <= in.height (otherwise VQ_ERR_UNSUPPORTED). The intermediate image and the gather
 * tables live in the context (grow-only; tables are rebuilt when the size pair changes): serialise calls per context. */
VQ_API int vq_image_resize(VqContext* ctx, VqImage in, VqImage out, void* stream);
/* HOST. The per-axis gather table vq_image_resize builds (first tap, tap count, normalised weights per output sample):
 * out[i] = sum_{t<count[i]} weights[i*capacity_taps + t] * in[clamp(start[i]+t, 0, in_size-1)], taps in increasing order.
 * start == NULL: only *max_taps is written. */
VQ_API int vq_resize_axis_table(int in_size, int out_size, int* start, int* count, float* weights, int capacity_taps, int* max_taps);

/* ------------------------------------------------------------------------------------------
 * SURVEY §8(f).3  The two streaming passes that complete a headless frame.
 *   Skydome: replaces the environment-map draw at SceneRendering.cpp:1821-1850 + PSMain (Skydome.hlsl:35-56).
 *     inv_view_proj = inverse of the matViewProj the engine uploads (SceneView.EnvironmentMapViewProj = skyCam.View *
 *     skyCam.Proj, Scene.cpp:573-584; row-vector convention), computed by the caller. The pass writes {rgb, 1} to
 *     rows [row_begin,row_end) of scene_color where no surface was shaded: pixels whose normal_mask texel has
 *     xyz == 0 (the G-buffer normal plane; the engine gets the same effect from the depth test), or every pixel when
 *     normal_mask is NULL.
 *   ApplyReflections: replaces CSMain (ApplyReflections.hlsl:31-57): scene.rgb += reflection.rgb, alpha kept;
 *     with a bounding_volumes layer (COMPOSITE_BOUNDING_VOLUMES) the sum is blended under it. In place.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_skydome(VqContext* ctx, const VqMatrix* inv_view_proj, VqPyramid hdri, const VqImage* normal_mask,
                      VqImage scene_color, int row_begin, int row_end, void* stream);
VQ_API int vq_apply_reflections(VqContext* ctx, VqImage scene_color, VqImage reflection_radiance,
                                const VqImage* bounding_volumes, void* stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY §8(f).4  Shadowed casters and the hierarchical MIN depth pyramid.
 *   vq_forward_lighting_shadowed: vq_forward_lighting with the shadow maps PSMain binds (ForwardLighting.hlsl:87-89):
 *     the caster lists are multiplied by OmnidirectionalShadowTestPCF / ShadowTestPCF and a shadowing directional light
 *     by ShadowTestPCF_Directional (Lighting.hlsl:79-272; ForwardLighting.hlsl:321-377). Maps are linear R32F, tightly
 *     packed, POINT-sampled with WRAP addressing (RootSignatures.cpp:148):
 *       point_cubes      [caster][face][y][x], value = distance / light range   (TextureCubeArray, t22)
 *       spot_maps        [caster][y][x],       value = light-space depth         (Texture2DArray, t16)
 *       directional_map  [y][x]                                                   (Texture2D, t13)
 *     A NULL map lights that light type unshadowed (factor 1). shadowViews / shadowViewDirectional and the
 *     f2*ShadowMapDimensions fields of per_frame are read as the shader reads them.
 *     Two launches: a PCF kernel that stores the shadowed-tap counts of every caster per pixel (8 B/pixel of scratch owned
 *     by the context), then the forward kernel, which weights the caster lights with them. Because of that scratch, calls
 *     on ONE context must be issued on one stream at a time (like vq_depth_min_pyramid).
 *   vq_depth_min_pyramid: replaces CSMain (DownsampleDepth.hlsl:85-119 = FidelityFX SPD with a MIN reduction): level 0 is
 *     a copy of the R32F depth image, level l = max(1, w>>l) x max(1, h>>l) holds the 2x2 minimum of the zero-padded
 *     level above; `levels` receives n_levels (<= vq_depth_pyramid_level_count = 1 + floor(log2(max(w,h))), at most 13)
 *     tightly packed levels, vq_depth_pyramid_texel_count() floats in total. The padded-domain levels ping-pong through scratch
 *     owned by the context (like vq_image_resize): calls on ONE context must be issued on one stream at a time.
 *   Parity on B200: tests/test_shadow_gpu.py (15 tests against the oracle, itself pinned to the reference's shader text).
 * ------------------------------------------------------------------------------------------ */
typedef struct VqShadowMaps {
    const void* point_cubes;     int32_t point_res;
    const void* spot_maps;       int32_t spot_width, spot_height;
    const void* directional_map; int32_t directional_width, directional_height;
} VqShadowMaps;
VQ_API int vq_forward_lighting_shadowed(VqContext* ctx,
                                        const VqPerFrameData* per_frame,
                                        const VqPerViewLightingData* per_view,
                                        const VqGBuffer* gbuffer,
                                        const VqEnvironmentMaps* env,
                                        const VqShadowMaps* shadow_maps,
                                        VqImage out_color,
                                        int row_begin, int row_end,
                                        void* stream);
VQ_API int      vq_depth_pyramid_level_count(int width, int height);
VQ_API uint64_t vq_depth_pyramid_texel_count(int width, int height, int levels);
VQ_API int vq_depth_min_pyramid(VqContext* ctx, VqImage depth_r32f, void* levels, int n_levels, void* stream);

/* ------------------------------------------------------------------------------------------
 * Blocking host-buffer entry points: the same passes called with HOST pointers (what an engine
 * integration that keeps its frame data in system memory would call). Each call uploads the
 * inputs, runs the kernel(s), downloads the result and returns when the result is in `out`.
 * Transfers are chunked by rows and overlapped with compute on internal streams.
 * ------------------------------------------------------------------------------------------ */
VQ_API int vq_forward_lighting_host(VqContext* ctx,
                                    const VqPerFrameData* per_frame,
                                    const VqPerViewLightingData* per_view,
                                    const VqGBuffer* host_gbuffer,
                                    const VqEnvironmentMaps* device_env,
                                    VqImage host_out_color);

#ifdef __cplusplus
}
#endif
#endif /* VQCUDA_H */
