/* c_client.c — the drop-in boundary from plain C: nothing but include/vqcuda.h and the CUDA runtime.
 *
 * What an engine-side integration does per environment map and per frame, in ~100 lines:
 *   load a Radiance .hdr file image (Image::LoadFromFile)  -> vq_hdr_parse + vq_hdr_load_host
 *   HDRI mip pyramid, diffuse + specular prefilter, LUT     -> vq_hdri_build_mips, vq_diffuse_irradiance,
 *                                                             vq_specular_prefilter, vq_brdf_integration_lut
 *   forward PBR over a G-buffer, sky, tonemap               -> vq_forward_lighting, vq_skydome, vq_tonemap
 *   save the frame as .hdr (Image::SaveToDisk)              -> vq_hdr_save_host
 * Build: gcc -std=c99 examples/c_client.c -Iinclude -I/usr/local/cuda/include -Lvqengine_b200 -lvqcuda \
 *            -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/vqengine_b200 -o c_client
 * Run:   ./c_client environment.hdr out.hdr            (needs a CUDA device: there is no CPU fallback)
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vqcuda.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != VQ_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vq_last_error()); return 1; } } while (0)
#define CUDA(call)  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

static void* dmalloc(size_t n) { void* p = NULL; return cudaMalloc(&p, n) == cudaSuccess ? p : NULL; }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s environment.hdr out.hdr\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char* file = (unsigned char*)malloc((size_t)n);
    if (fread(file, 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    VqContext* ctx = NULL;
    CHECK(vq_ctx_create(0, &ctx));

    /* ---- environment: .hdr -> pyramid -> diffuse / specular / LUT -------------------------------------------------- */
    VqHdrInfo info;
    CHECK(vq_hdr_parse(file, (uint64_t)n, &info, NULL));
    const int levels = vq_mip_level_count((uint64_t)info.width, (uint64_t)info.height);
    VqPyramid hdri = { dmalloc(vq_pyramid_texel_count(info.width, info.height, levels) * 16), info.width, info.height, levels };
    VqImage level0 = { hdri.ptr, info.width, info.height, (size_t)info.width * 16 };
    float maxLuminance = 0.0f;
    CHECK(vq_hdr_load_host(ctx, file, (uint64_t)n, level0, &maxLuminance));
    CHECK(vq_hdri_build_mips(ctx, hdri, NULL));
    printf("%s: %d x %d, %d mips, max luminance %.3f\n", argv[1], info.width, info.height, levels, maxLuminance);

    const int diffRes = 64, specRes = 256, specMips = vq_mip_level_count(specRes, specRes) - 1, lutRes = 512;
    VqCubemap diff = { dmalloc(vq_cubemap_texel_count(diffRes, 1) * 16), diffRes, 1 };
    VqCubemap spec = { dmalloc(vq_cubemap_texel_count(specRes, specMips) * 16), specRes, specMips };
    VqImage lut = { dmalloc((size_t)lutRes * lutRes * 8), lutRes, lutRes, (size_t)lutRes * 8 };
    VqDiffuseIrradianceParams dp = { 0.0f, 64, 16, levels > 3 ? 3 : levels - 1 };
    CHECK(vq_diffuse_irradiance(ctx, &dp, hdri, diff, 0, 6 * diffRes, NULL));
    CHECK(vq_specular_prefilter(ctx, hdri, spec, 512, 0, vq_cubemap_row_count(specRes, specMips), NULL));
    CHECK(vq_brdf_integration_lut(ctx, lut, 1024, 0, lutRes, NULL));
    VqEnvironmentMaps env = { diff, spec, lut };
    CHECK(vq_environment_prepare(ctx, &env, NULL));

    /* ---- a frame: a sphere in front of the camera (G-buffer written on the host), sky elsewhere ------------------------ */
    const int W = 640, H = 360;
    const size_t planeBytes = (size_t)W * H * 16;
    float *pos = (float*)calloc(1, planeBytes), *nrm = (float*)calloc(1, planeBytes), *alb = (float*)calloc(1, planeBytes);
    const float fovY = 1.0f, aspect = (float)W / H, t = tanf(0.5f * fovY);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            /* camera at the origin looking down +Z; ray through the pixel, sphere of radius 1 at (0,0,4) */
            const float dx = (2.0f * (x + 0.5f) / W - 1.0f) * t * aspect, dy = (1.0f - 2.0f * (y + 0.5f) / H) * t, dz = 1.0f;
            const float il = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz), rx = dx * il, ry = dy * il, rz = dz * il;
            const float b = rz * 4.0f, disc = b * b - 15.0f;
            if (disc <= 0.0f) continue;                                   /* normal stays 0: no surface -> skydome */
            const float s = b - sqrtf(disc);
            float* P = pos + ((size_t)y * W + x) * 4; float* N = nrm + ((size_t)y * W + x) * 4; float* A = alb + ((size_t)y * W + x) * 4;
            P[0] = rx * s; P[1] = ry * s; P[2] = rz * s; P[3] = 0.05f;     /* ao = ambient factor */
            N[0] = P[0]; N[1] = P[1]; N[2] = P[2] - 4.0f; N[3] = 0.1f + 0.8f * (float)x / W;   /* roughness sweeps left to right */
            A[0] = 0.9f; A[1] = 0.6f; A[2] = 0.2f; A[3] = y < H / 2 ? 1.0f : 0.0f;            /* metal on the top half */
        }
    VqGBuffer gb = { { dmalloc(planeBytes), W, H, (size_t)W * 16 }, { dmalloc(planeBytes), W, H, (size_t)W * 16 },
                     { dmalloc(planeBytes), W, H, (size_t)W * 16 }, { NULL, 0, 0, 0 } };
    CUDA(cudaMemcpy(gb.position_ao.ptr, pos, planeBytes, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(gb.normal_roughness.ptr, nrm, planeBytes, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(gb.albedo_metalness.ptr, alb, planeBytes, cudaMemcpyHostToDevice));

    VqPerFrameData* pf = (VqPerFrameData*)calloc(1, sizeof(VqPerFrameData));
    VqPerViewLightingData pv; memset(&pv, 0, sizeof(pv));
    pf->fAmbientLightingFactor = 0.05f;
    pf->Lights.numPointLights = 1;
    pf->Lights.point_lights[0].position.x = 3.0f; pf->Lights.point_lights[0].position.y = 3.0f; pf->Lights.point_lights[0].position.z = 1.0f;
    pf->Lights.point_lights[0].color.x = pf->Lights.point_lights[0].color.y = pf->Lights.point_lights[0].color.z = 1.0f;
    pf->Lights.point_lights[0].brightness = 60.0f; pf->Lights.point_lights[0].range = 50.0f;
    pv.MaxEnvMapLODLevels = (float)specMips;
    pv.ScreenDimensions.x = (float)W; pv.ScreenDimensions.y = (float)H;

    VqImage scene = { dmalloc(planeBytes), W, H, (size_t)W * 16 }, ldr = { dmalloc(planeBytes), W, H, (size_t)W * 16 };
    CHECK(vq_forward_lighting(ctx, pf, &pv, &gb, &env, scene, 0, H, NULL));
    /* inverse of (identity view) x (LH perspective): x' = x/w', y' = y/h', z = 1 plane -> direction (nx*t*aspect, ny*t, 1) */
    VqMatrix inv; memset(&inv, 0, sizeof(inv));
    inv.m[0] = t * aspect; inv.m[5] = t; inv.m[11] = 0.0f; inv.m[14] = 1.0f; inv.m[15] = 1.0f; inv.m[10] = 0.0f;
    VqPyramid sky = { hdri.ptr, hdri.width, hdri.height, 1 };
    CHECK(vq_skydome(ctx, &inv, sky, &gb.normal_roughness, scene, 0, H, NULL));
    VqTonemapperParams tm; memset(&tm, 0, sizeof(tm));
    tm.ToggleGammaCorrection = 1;
    CHECK(vq_tonemap(ctx, &tm, scene, ldr, NULL));

    /* ---- save the HDR scene colour as a Radiance file ------------------------------------------------------------------ */
    const uint64_t cap = 256 + (uint64_t)W * H * 6 + (uint64_t)H * 8;
    unsigned char* out = (unsigned char*)malloc((size_t)cap);
    uint64_t outSize = 0;
    CHECK(vq_hdr_save_host(ctx, scene, out, cap, &outSize));
    f = fopen(argv[2], "wb");
    if (!f || fwrite(out, 1, (size_t)outSize, f) != (size_t)outSize) { perror(argv[2]); return 2; }
    fclose(f);
    printf("%s: %d x %d, %llu bytes; %llu kernel launches\n", argv[2], W, H, (unsigned long long)outSize, (unsigned long long)vq_launch_count());
    CUDA(cudaDeviceSynchronize());
    vq_ctx_destroy(ctx);
    return 0;
}
