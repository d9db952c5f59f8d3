// vq_shadow.cu — SURVEY §8(f).4: shadowed casters and the hierarchical MIN depth pyramid.
//
//   vq_forward_lighting_shadowed   PSMain with the shadow maps bound (ForwardLighting.hlsl:321-377): the caster lists and a
//                                  shadowing directional light multiplied by their PCF tests (Lighting.hlsl:79-272)
//   vq_depth_min_pyramid           DownsampleDepth.hlsl:50-119 (FidelityFX SPD with the MIN reduction over a D3D mip chain)
//
// Checked against the oracle (oracle/oracle_shadow.cpp, itself pinned bit for bit to the reference's shader text,
// tests/test_hlsl_ref.py) on a B200: tests/test_shadow_gpu.py (-m gpu), memcheck clean (profiles/r02_shadow_first_run.txt).
//
// Structure of the shadowed pass = the oracle's: K1 (vq_forward.cu) shades everything that involves no caster — ambient,
// emissive, IBL, the non-shadowing point and spot lights — with the caster lists emptied and the directional light off;
// the kernel below then adds, in PSMain's order, point casters x OmnidirectionalShadowTestPCF, spot casters x ShadowTestPCF
// and the directional light x ShadowTestPCF_Directional onto that result. The sum order is the reference's.
//
// The per-pixel math lives in vq_shadow_math.cuh, which also compiles for the HOST: tests/test_shadow_math_host.py checks it
// against the oracle bit for bit on the CPU.
#include "vq_common.cuh"
#include "vq_shadow_math.cuh"
#include <string.h>

namespace {

struct ShadowParams {
    ImgV pos, nrm, alb, out;
    int rowBegin, rows, width;
    vqshadow::ShadowLights L;
};

__device__ __forceinline__ vqshadow::Px4 px4(float4 v) { vqshadow::Px4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }

// One thread per pixel; rows strided over the grid. Reads the K1 result (everything without casters), adds the caster terms.
__global__ void __launch_bounds__(128) shadow_casters_kernel(const __grid_constant__ ShadowParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.width) return;
    for (int r = blockIdx.y; r < P.rows; r += gridDim.y) {
        const int y = P.rowBegin + r;
        const float4 p4 = vq::ld_stream(P.pos.row(y) + x), n4 = vq::ld_stream(P.nrm.row(y) + x), a4 = vq::ld_stream(P.alb.row(y) + x);
        float4* dst = P.out.row(y) + x;
        const vqshadow::Px4 o = vqshadow::shade_casters(P.L, px4(p4), px4(n4), px4(a4), px4(*dst));
        vq::st_stream(dst, make_float4(o.x, o.y, o.z, o.w));
    }
}

// ---- MIN depth pyramid: one launch per level (each level is a quarter of the one above: HBM-bound, 1.33 x 4 B/texel) ----
// dst(x,y) = min of the zero-padded 2x2 block of src; src is the PADDED-domain level (ceil-halved sizes, see the oracle),
// kept in `pad` buffers; `store` receives the level clipped to its D3D size max(1, w>>l) x max(1, h>>l).
__global__ void __launch_bounds__(256) depth_min_level_kernel(const float* __restrict__ src, int sw, int sh,
                                                              float* __restrict__ pad, int pw, int ph,
                                                              float* __restrict__ store, int lw, int lh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= pw || y >= ph) return;
    const float m = vqshadow::depth_min_texel(src, sw, sh, x, y);
    pad[(size_t)y * pw + x] = m;
    if (x < lw && y < lh) store[(size_t)y * lw + x] = m;
}
__global__ void __launch_bounds__(256) depth_copy_kernel(const float* __restrict__ src, int pitch, float* __restrict__ dst, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w && y < h) dst[(size_t)y * w + x] = __ldg(src + (size_t)y * pitch + x);
}

int ensure_scratch(void** ptr, size_t* have, size_t need) {
    if (*have >= need && *ptr) return VQ_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *have = 0;
    if (cudaMalloc(ptr, need) != cudaSuccess) { cudaGetLastError(); vq_set_error("cudaMalloc(%zu) failed", need); return VQ_ERR_OUT_OF_MEMORY; }
    *have = need;
    return VQ_OK;
}

}  // namespace

extern "C" int vq_forward_lighting_shadowed(VqContext* ctx, const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                                            const VqGBuffer* gb, const VqEnvironmentMaps* env, const VqShadowMaps* sm,
                                            VqImage out, int row_begin, int row_end, void* stream) {
    int rc = vq_enter(ctx); if (rc) return rc;
    VQ_MARK("RenderSceneColor");
    VQ_REQUIRE(pf && pv && gb && env && sm, "null parameter block");
    const VqSceneLighting& L = pf->Lights;
    VQ_REQUIRE(L.numPointCasters >= 0 && L.numPointCasters <= VQ_NUM_SHADOWING_LIGHTS_POINT &&
               L.numSpotCasters >= 0 && L.numSpotCasters <= VQ_NUM_SHADOWING_LIGHTS_SPOT, "caster counts exceed the cbuffer arrays");
    if (sm->point_cubes) VQ_REQUIRE(sm->point_res > 0, "point shadow cubes need a positive resolution");
    if (sm->spot_maps) VQ_REQUIRE(sm->spot_width > 0 && sm->spot_height > 0, "spot shadow maps need positive dimensions");
    if (sm->directional_map) VQ_REQUIRE(sm->directional_width > 0 && sm->directional_height > 0, "directional shadow map needs positive dimensions");
    // everything without a caster: K1 with the caster lists emptied and the directional light off (oracle_shadow.cpp)
    const VqPerFrameData base = vqshadow::per_frame_without_casters(*pf);
    rc = vq_forward_launch(ctx, &base, pv, gb, env, out, row_begin, row_end, (cudaStream_t)stream); if (rc) return rc;
    if (row_begin == row_end) return VQ_OK;
    if (L.numPointCasters == 0 && L.numSpotCasters == 0 && !L.directional.enabled) return VQ_OK;

    ShadowParams P;
    memset(&P, 0, sizeof(P));
    P.pos = make_view(gb->position_ao); P.nrm = make_view(gb->normal_roughness); P.alb = make_view(gb->albedo_metalness);
    P.out = make_view(out);
    P.rowBegin = row_begin; P.rows = row_end - row_begin; P.width = gb->position_ao.width;
    vqshadow::fill_shadow_lights(P.L, *pf, *pv, *sm);

    const unsigned gx = (unsigned)((P.width + 127) / 128);
    unsigned gy = (unsigned)(ctx->sm_count * 8) / gx;
    if (gy < 1) gy = 1;
    if (gy > (unsigned)P.rows) gy = (unsigned)P.rows;
    if (gy > 65535u) gy = 65535u;
    shadow_casters_kernel<<<dim3(gx, gy), 128, 0, (cudaStream_t)stream>>>(P);
    return vq_check_launch("shadow_casters");
}

extern "C" int vq_depth_pyramid_level_count(int width, int height) { return vqshadow::depth_level_count(width, height); }
extern "C" uint64_t vq_depth_pyramid_texel_count(int width, int height, int levels) {
    uint64_t t = 0;
    for (int l = 0; l < levels; ++l) t += (uint64_t)((width >> l) > 0 ? (width >> l) : 1) * (uint64_t)((height >> l) > 0 ? (height >> l) : 1);
    return t;
}

extern "C" int vq_depth_min_pyramid(VqContext* ctx, VqImage depth, void* levels, int n_levels, void* stream_) {
    int rc = vq_enter(ctx); if (rc) return rc;
    VQ_MARK("DownsampleDepth");
    cudaStream_t stream = (cudaStream_t)stream_;
    VQ_REQUIRE(vq_image_ok(depth, 4), "bad depth image descriptor (R32F)");
    VQ_REQUIRE(levels && ((uintptr_t)levels % 4) == 0, "levels buffer is null or misaligned");
    const int W = depth.width, H = depth.height;
    VQ_REQUIRE(n_levels >= 1 && n_levels <= vq_depth_pyramid_level_count(W, H), "level count out of range");
    // two ping-pong buffers for the padded-domain levels (ceil-halved sizes): level 1 is at most ceil(W/2) x ceil(H/2)
    const size_t padBytes = (size_t)((W + 1) / 2) * ((H + 1) / 2) * 4;
    VqScratchLock lock(ctx);
    rc = ensure_scratch(&ctx->depth_pad, &ctx->depth_pad_bytes, padBytes * 2 > 16 ? padBytes * 2 : 16); if (rc) return rc;
    float* padA = (float*)ctx->depth_pad;
    float* padB = padA + padBytes / 4;
    float* out = (float*)levels;
    const dim3 blk(32, 8);
    depth_copy_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), blk, 0, stream>>>((const float*)depth.ptr, (int)(depth.pitch_bytes / 4), out, W, H);
    rc = vq_check_launch("depth_copy"); if (rc) return rc;
    vqshadow::DepthLevelPlan plan[13];
    vqshadow::depth_pyramid_plan(W, H, n_levels, plan);
    const float* src = out;                                   // level 0 doubles as its own padded-domain level
    for (int l = 1; l < n_levels; ++l) {
        const vqshadow::DepthLevelPlan& p = plan[l];
        float* pad = (l & 1) ? padA : padB;
        depth_min_level_kernel<<<dim3((p.pw + 31) / 32, (p.ph + 7) / 8), blk, 0, stream>>>(src, p.sw, p.sh, pad, p.pw, p.ph, out + p.out_offset, p.lw, p.lh);
        rc = vq_check_launch("depth_min_level"); if (rc) return rc;
        src = pad;
    }
    return VQ_OK;
}
