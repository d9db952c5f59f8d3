// vq_forward.cu — K1 forward PBR lighting for sm_100a.
//
// Replaces VQRenderer::RenderSceneColor (SceneRendering.cpp:1619-1851) + PSMain
// (ForwardLighting.hlsl:285-380) over a G-buffer of three float4 planes (+ optional emissive):
//   64 B/pixel of HBM traffic (3 x LDG.128 + 1 x STG.128, all coalesced, streaming);
//   the light array is staged once per block into shared memory; the IBL cubemaps and the BRDF LUT
//   are read through L1/L2 (they are L2-resident side data: <= 42 MB at the reference sizes).
// The math is PSMain's with per-pixel invariants hoisted out of the light loops (N, V, N.V, the
// Smith-G term of V, F0, kD factors). Discontinuities are evaluated exactly as the oracle does:
//   * `D < l.range` uses an unfused |L-P|^2 and a per-light threshold on the squared distance that is
//     equivalent to the correctly-rounded sqrt compare;
//   * the specular mip is int(roughness * MAX_LOD) (one fp32 multiply);
//   * cubemap taps are seamless across face edges, so face selection is not a discontinuity.
#include "vq_common.cuh"

using namespace vq;

namespace {

constexpr int FWD_THREADS = 256;
constexpr int MAX_POINT = VQ_NUM_LIGHTS_POINT + VQ_NUM_SHADOWING_LIGHTS_POINT;   // casters appended (shadow factor 1)
constexpr int MAX_SPOT = VQ_NUM_LIGHTS_SPOT + VQ_NUM_SHADOWING_LIGHTS_SPOT;

// Bordered sampling copy of a cubemap: every face of every mip is stored as (N+2)x(N+2) texels, the 1-texel
// border holding the neighbouring faces' edge texels (corners: the mean of the three texels that meet there), so
// that the seamless bilinear footprint never leaves the face. mipOffset[] are texel offsets of face 0 per mip.
struct CubeV { const float4* p; int res, mips; uint32_t mipOffset[16]; };
struct LutV { const float2* p; int w, h, pitch2; };

struct FwdParams {
    VqSceneLighting lights;            // 7088 B, read once per block into shared memory
    float3 cam;
    float cosB, sinB;                  // GetHDRIRotationMatrix (Lighting.hlsl:348-358), cos/sin(-offset)
    int maxLod;                        // int(MaxEnvMapLODLevels)
    int diffuseOnly;
    int hasEmissive;
    ImgV pos, nrm, alb, emi;
    ImgV outs[8];                      // destinations: the local frame and, for the fused gather, the peers' frames (NVLink P2P)
    int nOut, dstRowOffset;            // every shaded pixel goes to row dstRowOffset+y of every destination
    CubeV diff, spec;
    LutV lut;
    int rowBegin, rows, width;
};

struct SPoint { float3 pos; float d2Limit; float3 color; float brightness; };
struct SSpot { float3 pos; float outer; float3 color; float brightness; float3 dir; float inner; float invCone; float pad[3]; };

// ---------------------------------------------------------------------------------------------
// cubemap sampling: bilinear, seamless (identical rule in oracle/oracle_shading.cpp)
// ---------------------------------------------------------------------------------------------
// D3D face selection (largest |component|; ties X > Y > Z as in the oracle) written with selects so that a warp never
// diverges on it:  face, and the face-plane coordinates sx (right), sy (up) in [-1,1]
__device__ __forceinline__ void dir_to_face(float3 d, int& face, float& sx, float& sy) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    const bool isX = ax >= ay && ax >= az;
    const bool isY = !isX && ay >= az;
    const float ma = isX ? ax : (isY ? ay : az);
    const float mv = isX ? d.x : (isY ? d.y : d.z);          // signed major component
    const bool pos = mv > 0.0f;
    const float r = rcp_fast(ma);
    // +X: (-z, y)  -X: (z, y)  +Y: (x, -z)  -Y: (x, z)  +Z: (x, y)  -Z: (-x, y)
    const float u = isX ? (pos ? -d.z : d.z) : (isY ? d.x : (pos ? d.x : -d.x));
    const float v = isY ? (pos ? -d.z : d.z) : d.y;
    sx = u * r; sy = v * r;
    face = (isX ? 0 : (isY ? 2 : 4)) + (pos ? 0 : 1);
}

// integer-only neighbour lookup for a tap one texel outside the face (see DESIGN.md "cube edges"); used only by
// the border-padding kernel, never per pixel
__device__ void cube_resolve_edge(int N, int face, int i, int j, int& of, int& oi, int& oj) {
    const int A = 2 * i + 1 - N, B = N - 1 - 2 * j, C = N;
    int dx, dy, dz;
    switch (face) {
        case 0: dx = C;  dy = B;  dz = -A; break;
        case 1: dx = -C; dy = B;  dz = A;  break;
        case 2: dx = A;  dy = C;  dz = -B; break;
        case 3: dx = A;  dy = -C; dz = B;  break;
        case 4: dx = A;  dy = B;  dz = C;  break;
        default: dx = -A; dy = B; dz = -C; break;
    }
    const int M = N + 1;
    int nsx, nsy;
    if (dx == M)       { of = 0; nsx = -dz; nsy = dy; }
    else if (dx == -M) { of = 1; nsx = dz;  nsy = dy; }
    else if (dy == M)  { of = 2; nsx = dx;  nsy = -dz; }
    else if (dy == -M) { of = 3; nsx = dx;  nsy = dz; }
    else if (dz == M)  { of = 4; nsx = dx;  nsy = dy; }
    else               { of = 5; nsx = -dx; nsy = dy; }
    oi = min(max(((nsx + M) * N) / (2 * M), 0), N - 1);
    oj = min(max(((M - nsy) * N) / (2 * M), 0), N - 1);
}

// packed cube (mip-major / face-minor, N x N faces) -> bordered sampling copy ((N+2) x (N+2) faces)
__global__ void __launch_bounds__(256) cube_pad_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                        int res, int mips, uint32_t totalPadded) {
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < totalPadded; idx += gridDim.x * 256u) {
        // locate (mip, face, pj, pi) of this padded texel
        uint32_t rem = idx, srcOff = 0; int m = 0, N = res;
        for (; m < mips; ++m) {
            N = res >> m;
            const uint32_t sz = 6u * (uint32_t)(N + 2) * (uint32_t)(N + 2);
            if (rem < sz) break;
            rem -= sz; srcOff += 6u * (uint32_t)N * (uint32_t)N;
        }
        const int P = N + 2;
        const int face = (int)(rem / (uint32_t)(P * P));
        const int r2 = (int)(rem % (uint32_t)(P * P));
        const int i = r2 % P - 1, j = r2 / P - 1;           // face-relative texel, -1..N
        const float4* sm = src + srcOff;
        auto fetch = [&](int f, int x, int y) { return __ldg(sm + (size_t)f * N * N + (size_t)y * N + x); };
        auto edge = [&](int x, int y) { int f2, i2, j2; cube_resolve_edge(N, face, x, y, f2, i2, j2); return fetch(f2, i2, j2); };
        const bool oi = (i < 0 || i >= N), oj = (j < 0 || j >= N);
        float4 v;
        if (!oi && !oj) v = fetch(face, i, j);
        else if (oi != oj) v = edge(i, j);
        else {   // cube corner: mean of the three texels meeting there (own corner + the two edge neighbours)
            const int ci = i < 0 ? 0 : N - 1, cj = j < 0 ? 0 : N - 1;
            const float4 a = fetch(face, ci, cj), b = edge(i, cj), c = edge(ci, j);
            v = make_float4((a.x + b.x + c.x) * (1.0f / 3.0f), (a.y + b.y + c.y) * (1.0f / 3.0f),
                            (a.z + b.z + c.z) * (1.0f / 3.0f), (a.w + b.w + c.w) * (1.0f / 3.0f));
        }
        dst[idx] = v;
    }
}

__device__ __forceinline__ float3 sample_cube(const CubeV& c, float3 dir, int mip) {
    mip = min(max(mip, 0), c.mips - 1);
    const int N = c.res >> mip, P = N + 2;
    int face; float sx, sy;
    dir_to_face(dir, face, sx, sy);
    const float x = fmaf(fmaf(sx, 0.5f, 0.5f), (float)N, -0.5f);
    const float y = fmaf(fmaf(-sy, 0.5f, 0.5f), (float)N, -0.5f);
    const float xf = fminf(fmaxf(floorf(x), -1.0f), (float)(N - 1)), yf = fminf(fmaxf(floorf(y), -1.0f), (float)(N - 1));
    const float fx = x - xf, fy = y - yf;
    const int i0 = (int)xf + 1, j0 = (int)yf + 1;            // bordered coordinates: 0..N
    const float4* p = c.p + (c.mipOffset[mip] + (uint32_t)(face * (P * P) + j0 * P + i0));
    const float4 t00 = __ldg(p), t10 = __ldg(p + 1), t01 = __ldg(p + P), t11 = __ldg(p + P + 1);
    const float3 top = lerp(xyz(t00), xyz(t10), fx), bot = lerp(xyz(t01), xyz(t11), fx);
    return lerp(top, bot, fy);
}

__device__ __forceinline__ float2 sample_lut(const LutV& l, float u, float v) {   // bilinear, CLAMP
    const float x = fmaf(u, (float)l.w, -0.5f), y = fmaf(v, (float)l.h, -0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = min(max((int)x0, 0), l.w - 1), ix1 = min(max((int)x0 + 1, 0), l.w - 1);
    const int iy0 = min(max((int)y0, 0), l.h - 1), iy1 = min(max((int)y0 + 1, 0), l.h - 1);
    const uint32_t r0 = (uint32_t)(iy0 * l.pitch2), r1 = (uint32_t)(iy1 * l.pitch2);
    const float2 p00 = __ldg(l.p + (r0 + ix0)), p10 = __ldg(l.p + (r0 + ix1));
    const float2 p01 = __ldg(l.p + (r1 + ix0)), p11 = __ldg(l.p + (r1 + ix1));
    return make_float2(lerp(lerp(p00.x, p10.x, fx), lerp(p01.x, p11.x, fx), fy),
                       lerp(lerp(p00.y, p10.y, fx), lerp(p01.y, p11.y, fx), fy));
}

// ---------------------------------------------------------------------------------------------
// per-pixel shading state with everything that does not depend on the light hoisted
// ---------------------------------------------------------------------------------------------
// Only what the light loop reads stays live across it (register diet: 64-80 registers decide the occupancy).
struct Px {
    float3 P, Nn, V;              // position, normalize(Ns), normalize(cam - P)
    float nsLen;                  // |Ns|: dot(Ns,Wi) = nsLen * dot(Nn,Wi)  (Lighting.hlsl:316 uses the raw s.N)
    float nv, NdotV, gV;          // dot(Nn,V), saturate, Smith-G1 of V (BRDF.hlsl:82-97)
    float a2, a2m1, k, omk;
    const float4* nrmTexel;       // where the raw normal lives: re-read by the exact slow path only
};

// ---- exact re-evaluation of N.H -------------------------------------------------------------------
// GGX's t = nh2*(a2-1)+1 cancels catastrophically near a highlight on a smooth surface (t ~ a2 ~ 1e-5),
// so D amplifies a 1-ulp difference in N.H by up to 1e4. Where t is small the kernel therefore recomputes
// N.H with the oracle's exact operation sequence (correctly rounded div/sqrt, no FMA contraction): V, Wo,
// N, Wi, H as BRDF.hlsl:166-169 / Lighting.hlsl:312 write them. T_EXACT = 0.02 bounds the fast path's
// relative error in D by 2*dt/t <= 2*5e-7/0.02 = 5e-5; the slow path runs for < 1 % of pixel-light pairs.
constexpr float T_EXACT = 0.02f;

__device__ __forceinline__ float dot_u(float3 a, float3 b) {     // (x*x' + y*y') + z*z', every op rounded
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
// v / sqrt(dot(v,v)) with correctly rounded sqrt and divisions (== the oracle's normalize)
__device__ __forceinline__ float3 normalize_u(float3 v) {
    const float l = __fsqrt_rn(dot_u(v, v));
    return f3(__fdiv_rn(v.x, l), __fdiv_rn(v.y, l), __fdiv_rn(v.z, l));
}

__device__ __noinline__ float exact_ndoth(float3 cam, float3 P, const float4* nrmTexel, float3 wiSrc, float wiLenSq) {
    const float4 nr = __ldg(nrmTexel);
    const float3 Ns = f3(nr.x, nr.y, nr.z);
    const float3 Vv = f3(__fsub_rn(cam.x, P.x), __fsub_rn(cam.y, P.y), __fsub_rn(cam.z, P.z));
    const float3 V = normalize_u(Vv);                 // ForwardLighting.hlsl:285
    const float3 Wo = normalize_u(V);                 // BRDF.hlsl:166
    const float3 N = normalize_u(Ns);                 // BRDF.hlsl:167
    const float wl = __fsqrt_rn(wiLenSq);
    const float3 Wi = f3(__fdiv_rn(wiSrc.x, wl), __fdiv_rn(wiSrc.y, wl), __fdiv_rn(wiSrc.z, wl));
    const float3 Hs = f3(__fadd_rn(Wo.x, Wi.x), __fadd_rn(Wo.y, Wi.y), __fadd_rn(Wo.z, Wi.z));
    const float3 H = normalize_u(Hs);                 // BRDF.hlsl:168
    return saturate(dot_u(N, H));
}

struct Acc { float3 a, b, c; };   // sum over lights of w*col * {(1-fc), fc*spec, spec}

// One light: accumulates BRDF(s, Wi, V) * radiance * NdotL (BRDF.hlsl:163-194, Lighting.hlsl:308-345) in the
// factored form  r = K1*(1-fc) + omF0*(fc*spec) + F0*spec  with  F = F0 + (1-F0)*fc.
//   Lv, d2 : un-normalised light vector and its squared length (Wi = Lv/sqrt(d2)),  invD = 1/sqrt(d2)
//   scale  : attenuation * brightness * spot intensity (0 when the light is out of range);  col : light colour
// H = normalize(V+Wi) is never formed: |V+Wi|^2 = 2+2c with c = V.Wi, so N.H = (N.V+N.Wi)*rh and
// H.V = (1+c)*rh with rh = rsqrt(2+2c).
// The body is branch-free (a light that does not contribute gets weight 0) so that two lights can be interleaved
// by the scheduler; only the rare exact-N.H slow path branches.
__device__ __forceinline__ void shade_light(const Px& s, Acc& acc, float3 cam, float3 Lv, float d2, float invD,
                                            float scale, float3 col) {
    const float nl = dot(s.Nn, Lv) * invD;
    const float c = dot(s.V, Lv) * invD;
    const float sq = fmaxf(fmaf(2.0f, c, 2.0f), 1e-12f);
    const float rh = rsqrt_fast(sq);
    float NdotH = saturate((s.nv + nl) * rh);
    const float HV = fmaxf(0.0f, (1.0f + c) * rh);
    const float fc = pow5(1.0f - HV);                  // Fresnel_Schlick (BRDF.hlsl:132-136)
    float t = fmaf(NdotH * NdotH, s.a2m1, 1.0f);       // NormalDistributionGGX (BRDF.hlsl:65-79)
    const float NL = fminf(fmaxf(nl, 0.0f), 1.0f);     // saturate(N.L) == max(0,N.L) for unit vectors
    const float w = saturate(s.nsLen * nl) * scale;    // NdotL of the raw s.N (Lighting.hlsl:316) * radiance scale
    if (t < T_EXACT && w > 0.0f) {
        NdotH = exact_ndoth(cam, s.P, s.nrmTexel, Lv, d2);
        t = __fadd_rn(__fmul_rn(__fmul_rn(NdotH, NdotH), s.a2m1), 1.0f);
    }
    const float dDen = PI * (t * t);
    const float gDen = fmaf(NL, s.omk, s.k) + 0.0001f; // Geometry_Smiths_SchlickGGX of L (BRDF.hlsl:82-97)
    const float sDen = fmaxf(4.0f * s.NdotV * NL, 0.0001f);
    const bool tiny = dDen < 0.000000000001f;          // `denom < EPSILON -> D = 1` (BRDF.hlsl:77)
    const float num = (tiny ? 1.0f : s.a2) * s.gV * NL;
    const float den = (tiny ? 1.0f : dDen) * gDen * sDen;
    const float spec = num * rcp_fast(den);            // D*G/denom with ONE reciprocal (den >= 1e-20)
    const float wa = (1.0f - fc) * w, wc = spec * w, wb = fc * wc;
    acc.a.x = fmaf(wa, col.x, acc.a.x); acc.a.y = fmaf(wa, col.y, acc.a.y); acc.a.z = fmaf(wa, col.z, acc.a.z);
    acc.b.x = fmaf(wb, col.x, acc.b.x); acc.b.y = fmaf(wb, col.y, acc.b.y); acc.b.z = fmaf(wb, col.z, acc.b.z);
    acc.c.x = fmaf(wc, col.x, acc.c.x); acc.c.y = fmaf(wc, col.y, acc.c.y); acc.c.z = fmaf(wc, col.z, acc.c.z);
}

// point light i (Lighting.hlsl:308-322): in range <=> d2 < d2Limit (== length(Lw-P) < l.range, exactly)
__device__ __forceinline__ void shade_point(const Px& s, Acc& acc, float3 cam, const SPoint& l) {
    const float3 Lv = l.pos - s.P;
    const float d2 = dot_u(Lv, Lv);                    // |L-P|^2 exactly as the oracle's dot()
    const float invD = rsqrt_fast(fmaxf(d2, 1e-30f));
    const float scale = d2 < l.d2Limit ? (invD * invD) * l.brightness : 0.0f;   // AttenuationBRDF = 1/D^2
    shade_light(s, acc, cam, Lv, d2, invD, scale, l.color);
}

constexpr int FWD_BX = 64, FWD_BY = 4;   // 256 threads: 64 x 4 pixel tile; grid.x covers the row, grid.y strides rows

// Tuning knobs, A/B-measured on B200 at 4K (profiles/r01_forward_variants.txt, us per frame):
//   blocks/SM, prefetch, pair, IBL-last:  3,0,0,1 -> 328 (80 regs, no spills)  | 4,0,0,1 -> 334 | 4,0,0,0 -> 344
//   3,0,1,1 -> 349 | 2,1,1,0 (116 regs) -> 358 | 3,0,1,0 -> 364 | 4,0,1,x -> 376-379 | 3,1,1,0 -> 381 | 5,0,0,1 -> 437
// Shading the lights first and the environment map last keeps the 8 gathered texels out of the light loop's live
// range; neither register prefetching nor interleaving two lights pays at this register budget.
#ifndef FWD_MIN_BLOCKS
#define FWD_MIN_BLOCKS 3
#endif
#ifndef FWD_PREFETCH
#define FWD_PREFETCH 0
#endif
#ifndef FWD_PAIR
#define FWD_PAIR 0
#endif
#ifndef FWD_IBL_LAST
#define FWD_IBL_LAST 1
#endif
__global__ void __launch_bounds__(FWD_THREADS, FWD_MIN_BLOCKS) forward_kernel(const __grid_constant__ FwdParams P) {
    __shared__ SPoint sPoint[MAX_POINT];
    __shared__ SSpot sSpot[MAX_SPOT];

    // ---- stage the light arrays (Scene::GatherLightData layout) into shared memory, once per block ----
    const VqSceneLighting& L = P.lights;
    const int nP = L.numPointLights, nPC = L.numPointCasters, nS = L.numSpotLights, nSC = L.numSpotCasters;
    const int tid = threadIdx.y * FWD_BX + threadIdx.x;
    for (int i = tid; i < nP + nPC; i += FWD_THREADS) {
        const VqPointLight& l = i < nP ? L.point_lights[i] : L.point_casters[i - nP];
        SPoint s;
        s.pos = f3(l.position.x, l.position.y, l.position.z);
        s.color = f3(l.color.x, l.color.y, l.color.z);
        s.brightness = l.brightness;
        // smallest d2 with sqrt_rn(d2) >= range: then (d2 < limit) == (sqrt_rn(d2) < range) exactly
        float lim = l.range * l.range;
        if (!(l.range > 0.0f)) lim = 0.0f;
        else {
            while (sqrtf(lim) >= l.range && lim > 0.0f) lim = __uint_as_float(__float_as_uint(lim) - 1u);
            while (sqrtf(lim) < l.range) lim = __uint_as_float(__float_as_uint(lim) + 1u);
        }
        s.d2Limit = lim;
        sPoint[i] = s;
    }
    for (int i = tid; i < nS + nSC; i += FWD_THREADS) {
        const VqSpotLight& l = i < nS ? L.spot_lights[i] : L.spot_casters[i - nS];
        SSpot s;
        s.pos = f3(l.position.x, l.position.y, l.position.z);
        s.color = f3(l.color.x, l.color.y, l.color.z);
        s.brightness = l.brightness;
        s.dir = normalize_u(f3(l.spotDir.x, l.spotDir.y, l.spotDir.z));   // normalize(l.spotDir), Lighting.hlsl:60
        s.outer = l.outerConeAngle; s.inner = l.innerConeAngle;
        s.invCone = 1.0f / (l.outerConeAngle - l.innerConeAngle);
        sSpot[i] = s;
    }
    __syncthreads();
    const int numPoint = nP + nPC, numSpot = nS + nSC;
    const bool dirEnabled = L.directional.enabled != 0;
    float3 dirWi = f3(0.0f), dirRadiance = f3(0.0f);
    if (dirEnabled) {                                            // Lighting.hlsl:334-345
        dirWi = normalize_u(f3(-L.directional.lightDirection.x, -L.directional.lightDirection.y, -L.directional.lightDirection.z));
        dirRadiance = f3(L.directional.color.x, L.directional.color.y, L.directional.color.z) * L.directional.brightness;
    }

    const int x = blockIdx.x * FWD_BX + threadIdx.x;
    if (x >= P.width) return;
    const int rowStep = gridDim.y * FWD_BY;
    int ry = blockIdx.y * FWD_BY + threadIdx.y;
    if (ry >= P.rows) return;
    // software prefetch: the next row's G-buffer texels are requested before the current pixel is shaded
    float4 pa = ld_stream(P.pos.row(P.rowBegin + ry) + x);
    float4 nr = ld_stream(P.nrm.row(P.rowBegin + ry) + x);
    float4 am = ld_stream(P.alb.row(P.rowBegin + ry) + x);
    for (; ry < P.rows; ry += rowStep) {
        const int y = P.rowBegin + ry;
        const int ryn = ry + rowStep;
        const bool more = ryn < P.rows;
        float4 paN = pa, nrN = nr, amN = am;
        if (FWD_PREFETCH && more) {
            paN = ld_stream(P.pos.row(P.rowBegin + ryn) + x);
            nrN = ld_stream(P.nrm.row(P.rowBegin + ryn) + x);
            amN = ld_stream(P.alb.row(P.rowBegin + ryn) + x);
        }

        Px s;
        s.P = xyz(pa);
        s.nrmTexel = P.nrm.row(y) + x;
        const float3 Ns = xyz(nr), albedo = xyz(am);
        const float roughness = nr.w, metalness = am.w, ao = pa.w;
        const float3 Vv = P.cam - s.P;
        s.V = Vv * rsqrt_fast(dot(Vv, Vv));                      // ForwardLighting.hlsl:285
        const float n2 = dot(Ns, Ns), rn = rsqrt_fast(n2);
        s.Nn = Ns * rn;                                          // BRDF.hlsl:167
        s.nsLen = n2 * rn;
        const float a = roughness * roughness;
        s.a2 = __fmul_rn(a, a); s.a2m1 = __fsub_rn(s.a2, 1.0f);  // no contraction: feeds the exact t
        const float rp1 = roughness + 1.0f;
        s.k = (rp1 * rp1) * 0.125f; s.omk = 1.0f - s.k;
        s.nv = dot(s.Nn, s.V);
        s.NdotV = saturate(s.nv);
        s.gV = s.NdotV * rcp_fast(fmaf(s.NdotV, s.omk, s.k) + 0.0001f);

        float3 I = albedo * ao;                                  // ForwardLighting.hlsl:290-293
        if (P.hasEmissive) {
            const float4 em = ld_stream(P.emi.row(y) + x);
            I += xyz(em) * em.w;
        }

#if !FWD_IBL_LAST
        // ---- environment map (Lighting.hlsl:360-395, BRDF.hlsl:196-207) ----
        {
            const float3 F0 = lerp(f3(0.04f), albedo, metalness);
            const float NdotVs = saturate(s.nsLen * s.nv);       // saturate(dot(s.N, V))
            const bool rot = P.sinB != 0.0f || P.cosB != 1.0f;       // uniform: yaw offset 0 is the common case
            const float3 Nr = rot ? f3(Ns.x * P.cosB - Ns.z * P.sinB, Ns.y, Ns.x * P.sinB + Ns.z * P.cosB) : Ns;
            const float3 diffIrr = sample_cube(P.diff, Nr, 0);
            float3 specCol = f3(0.0f); float2 sb = make_float2(0.0f, 0.0f);
            if (!P.diffuseOnly) {
                const float3 R0 = reflect(-s.V, Ns);
                const float3 R = rot ? f3(R0.x * P.cosB - R0.z * P.sinB, R0.y, R0.x * P.sinB + R0.z * P.cosB) : R0;
                const int mip = (int)(roughness * (float)P.maxLod);
                specCol = sample_cube(P.spec, R, mip);
                sb = sample_lut(P.lut, NdotVs, roughness);
            }
            const float fr = pow5(1.0f - NdotVs);                // FresnelWithRoughness, BRDF.hlsl:152-156
            const float omr = 1.0f - roughness;
            const float3 Ks = f3(fmaf(fmaxf(omr, F0.x) - F0.x, fr, F0.x),
                                 fmaf(fmaxf(omr, F0.y) - F0.y, fr, F0.y),
                                 fmaf(fmaxf(omr, F0.z) - F0.z, fr, F0.z));
            const float om = 1.0f - metalness;
            I.x += (1.0f - Ks.x) * om * (diffIrr.x * albedo.x) + specCol.x * fmaf(Ks.x, sb.x, sb.y);
            I.y += (1.0f - Ks.y) * om * (diffIrr.y * albedo.y) + specCol.y * fmaf(Ks.y, sb.x, sb.y);
            I.z += (1.0f - Ks.z) * om * (diffIrr.z * albedo.z) + specCol.z * fmaf(Ks.z, sb.x, sb.y);
        }

#endif
        Acc acc; acc.a = f3(0.0f); acc.b = f3(0.0f); acc.c = f3(0.0f);
        // ---- point lights, then unshadowed point casters (Lighting.hlsl:308-322; PSMain :310-313,321-340) ----
        // two at a time: the two bodies are independent, so their dependency chains interleave
        int i = 0;
        for (; FWD_PAIR && i + 1 < numPoint; i += 2) {
            const SPoint l0 = sPoint[i], l1 = sPoint[i + 1];
            shade_point(s, acc, P.cam, l0);
            shade_point(s, acc, P.cam, l1);
        }
        for (; i < numPoint; ++i) shade_point(s, acc, P.cam, sPoint[i]);
        // ---- spot lights, then unshadowed spot casters (Lighting.hlsl:57-73,323-333) ----
        for (int k = 0; k < numSpot; ++k) {
            const SSpot l = sSpot[k];
            const float3 Lv = l.pos - s.P;
            const float d2 = dot_u(Lv, Lv);
            const float invD = rsqrt_fast(fmaxf(d2, 1e-30f));
            const float theta = acosf(fminf(fmaxf(-dot(Lv, l.dir) * invD, -1.0f), 1.0f));   // pixel direction = -Wi
            float inten = 1.0f - (theta - l.inner) * l.invCone;
            inten = theta > l.outer ? 0.0f : (theta <= l.inner ? 1.0f : inten);
            shade_light(s, acc, P.cam, Lv, d2, invD, inten * l.brightness * (invD * invD), l.color);
        }
        // ---- directional (PSMain :360-377 with ShadowingFactor = 1) ----
        if (dirEnabled) shade_light(s, acc, P.cam, dirWi, 1.0f, 1.0f, 1.0f, dirRadiance);

        {   // F0 = lerp(0.04, albedo, metal) (BRDF.hlsl:177); K1 = (1-F0)*(1-metal)*albedo/PI (BRDF.hlsl:189-191)
            const float3 F0 = lerp(f3(0.04f), albedo, metalness);
            const float3 omF0 = f3(1.0f) - F0;
            const float3 K1 = omF0 * albedo * ((1.0f - metalness) * (1.0f / PI));
            I.x += fmaf(K1.x, acc.a.x, fmaf(omF0.x, acc.b.x, F0.x * acc.c.x));
            I.y += fmaf(K1.y, acc.a.y, fmaf(omF0.y, acc.b.y, F0.y * acc.c.y));
            I.z += fmaf(K1.z, acc.a.z, fmaf(omF0.z, acc.b.z, F0.z * acc.c.z));
        }
#if FWD_IBL_LAST
        // ---- environment map (Lighting.hlsl:360-395, BRDF.hlsl:196-207) ----
        {
            const float3 F0 = lerp(f3(0.04f), albedo, metalness);
            const float NdotVs = saturate(s.nsLen * s.nv);       // saturate(dot(s.N, V))
            const bool rot = P.sinB != 0.0f || P.cosB != 1.0f;       // uniform: yaw offset 0 is the common case
            const float3 Nr = rot ? f3(Ns.x * P.cosB - Ns.z * P.sinB, Ns.y, Ns.x * P.sinB + Ns.z * P.cosB) : Ns;
            const float3 diffIrr = sample_cube(P.diff, Nr, 0);
            float3 specCol = f3(0.0f); float2 sb = make_float2(0.0f, 0.0f);
            if (!P.diffuseOnly) {
                const float3 R0 = reflect(-s.V, Ns);
                const float3 R = rot ? f3(R0.x * P.cosB - R0.z * P.sinB, R0.y, R0.x * P.sinB + R0.z * P.cosB) : R0;
                const int mip = (int)(roughness * (float)P.maxLod);
                specCol = sample_cube(P.spec, R, mip);
                sb = sample_lut(P.lut, NdotVs, roughness);
            }
            const float fr = pow5(1.0f - NdotVs);                // FresnelWithRoughness, BRDF.hlsl:152-156
            const float omr = 1.0f - roughness;
            const float3 Ks = f3(fmaf(fmaxf(omr, F0.x) - F0.x, fr, F0.x),
                                 fmaf(fmaxf(omr, F0.y) - F0.y, fr, F0.y),
                                 fmaf(fmaxf(omr, F0.z) - F0.z, fr, F0.z));
            const float om = 1.0f - metalness;
            I.x += (1.0f - Ks.x) * om * (diffIrr.x * albedo.x) + specCol.x * fmaf(Ks.x, sb.x, sb.y);
            I.y += (1.0f - Ks.y) * om * (diffIrr.y * albedo.y) + specCol.y * fmaf(Ks.y, sb.x, sb.y);
            I.z += (1.0f - Ks.z) * om * (diffIrr.z * albedo.z) + specCol.z * fmaf(Ks.z, sb.x, sb.y);
        }

#endif
        {   // :380 — one STG.128 per destination; peer destinations are mapped NVLink addresses (fused compute + gather)
            const float4 o = make_float4(I.x, I.y, I.z, roughness);
            for (int k = 0; k < P.nOut; ++k) st_stream(P.outs[k].row(P.dstRowOffset + y) + x, o);
        }
        if (FWD_PREFETCH) { pa = paN; nr = nrN; am = amN; }
        else if (more) {
            pa = ld_stream(P.pos.row(P.rowBegin + ryn) + x);
            nr = ld_stream(P.nrm.row(P.rowBegin + ryn) + x);
            am = ld_stream(P.alb.row(P.rowBegin + ryn) + x);
        }
    }
}

uint64_t padded_texels(int res, int mips) {
    uint64_t n = 0;
    for (int m = 0; m < mips; ++m) { const uint64_t p = (uint64_t)(res >> m) + 2; n += 6 * p * p; }
    return n;
}
bool cube_desc_ok(const VqCubemap& c) {
    return c.ptr && c.res >= 1 && c.mips >= 1 && c.mips <= 16 && (c.res >> (c.mips - 1)) >= 1 &&
           padded_texels(c.res, c.mips) < (1ull << 31);
}
// builds the bordered sampling copy of `c` into `dst` (padded_texels() float4s) on `stream`
int pad_cube(const VqCubemap& c, float4* dst, cudaStream_t stream) {
    const uint32_t total = (uint32_t)padded_texels(c.res, c.mips);
    unsigned blocks = (total + 255u) / 256u;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    cube_pad_kernel<<<blocks, 256, 0, stream>>>((const float4*)c.ptr, dst, c.res, c.mips, total);
    return vq_check_launch("cube_pad");
}
void fill_cube_view(const VqCubemap& c, const float4* padded, CubeV& v) {
    v.p = padded; v.res = c.res; v.mips = c.mips;
    uint32_t off = 0;
    for (int m = 0; m < 16; ++m) {
        v.mipOffset[m] = m < c.mips ? off : 0u;
        if (m < c.mips) { const uint32_t p = (uint32_t)(c.res >> m) + 2u; off += 6u * p * p; }
    }
}
bool same_cube(const VqCubemap& a, const VqCubemap& b) { return a.ptr == b.ptr && a.res == b.res && a.mips == b.mips; }

int ensure_bytes(void** ptr, size_t* have, size_t need) {
    if (*have >= need && *ptr) return VQ_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *have = 0;
    if (cudaMalloc(ptr, need) != cudaSuccess) { cudaGetLastError(); vq_set_error("cudaMalloc(%zu) failed", need); return VQ_ERR_OUT_OF_MEMORY; }
    *have = need;
    return VQ_OK;
}

}  // namespace

int vq_forward_launch_multi(VqContext* ctx, const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                            const VqGBuffer* gb, const VqEnvironmentMaps* env, const VqImage* outs, int n_outs,
                            int dst_row_offset, int row_begin, int row_end, cudaStream_t stream) {
    VQ_REQUIRE(pf && pv && gb && env && outs, "null parameter block");
    VQ_REQUIRE(n_outs >= 1 && n_outs <= 8, "1..8 destinations");
    VQ_REQUIRE(vq_image_ok(gb->position_ao) && vq_image_ok(gb->normal_roughness) && vq_image_ok(gb->albedo_metalness), "bad image descriptor");
    const int W = gb->position_ao.width, H = gb->position_ao.height;
    VQ_REQUIRE(gb->normal_roughness.width == W && gb->albedo_metalness.width == W &&
               gb->normal_roughness.height == H && gb->albedo_metalness.height == H, "G-buffer planes must have the same size");
    VQ_REQUIRE(dst_row_offset >= 0, "negative destination row offset");
    for (int k = 0; k < n_outs; ++k)
        VQ_REQUIRE(vq_image_ok(outs[k]) && outs[k].width == W && outs[k].height >= dst_row_offset + H,
                   "every destination must be as wide as the G-buffer and hold dst_row_offset + height rows");
    VQ_REQUIRE(row_begin >= 0 && row_end <= H && row_begin <= row_end, "row range out of bounds");
    const VqSceneLighting& L = pf->Lights;
    VQ_REQUIRE(L.numPointLights >= 0 && L.numPointLights <= VQ_NUM_LIGHTS_POINT &&
               L.numSpotLights >= 0 && L.numSpotLights <= VQ_NUM_LIGHTS_SPOT &&
               L.numPointCasters >= 0 && L.numPointCasters <= VQ_NUM_SHADOWING_LIGHTS_POINT &&
               L.numSpotCasters >= 0 && L.numSpotCasters <= VQ_NUM_SHADOWING_LIGHTS_SPOT, "light counts exceed the cbuffer arrays");
    if (row_begin == row_end) return VQ_OK;

    FwdParams P;
    memset(&P, 0, sizeof(P));
    P.lights = L;
    P.cam = make_float3(pv->CameraPosition.x, pv->CameraPosition.y, pv->CameraPosition.z);
    P.cosB = cosf(-pf->fHDRIOffsetInRadians);
    P.sinB = sinf(-pf->fHDRIOffsetInRadians);
    P.maxLod = (int)pv->MaxEnvMapLODLevels;
    P.diffuseOnly = pv->EnvironmentMapDiffuseOnlyIllumination != 0;
    P.pos = make_view(gb->position_ao); P.nrm = make_view(gb->normal_roughness); P.alb = make_view(gb->albedo_metalness);
    for (int k = 0; k < n_outs; ++k) P.outs[k] = make_view(outs[k]);
    P.nOut = n_outs; P.dstRowOffset = dst_row_offset;
    P.hasEmissive = gb->emissive.ptr != nullptr;
    if (P.hasEmissive) {
        VQ_REQUIRE(vq_image_ok(gb->emissive) && gb->emissive.width == W && gb->emissive.height == H, "bad emissive plane");
        P.emi = make_view(gb->emissive);
    }
    VQ_REQUIRE(cube_desc_ok(env->irradiance_diffuse), "bad cubemap descriptor (irradiance_diffuse)");
    if (!P.diffuseOnly) {
        VQ_REQUIRE(cube_desc_ok(env->irradiance_specular), "bad cubemap descriptor (irradiance_specular)");
        VQ_REQUIRE(vq_image_ok(env->brdf_lut, 8), "bad BRDF LUT descriptor");
        P.lut.p = (const float2*)env->brdf_lut.ptr; P.lut.w = env->brdf_lut.width; P.lut.h = env->brdf_lut.height;
        P.lut.pitch2 = (int)(env->brdf_lut.pitch_bytes / 8);
    }
    // bordered sampling copies: from the prepared environment when it matches, else padded now on this stream
    int rc;
    const bool prepared = ctx->env_valid && same_cube(ctx->env_key.irradiance_diffuse, env->irradiance_diffuse) &&
                          (P.diffuseOnly || same_cube(ctx->env_key.irradiance_specular, env->irradiance_specular));
    if (prepared) {
        fill_cube_view(env->irradiance_diffuse, (const float4*)ctx->env_diff, P.diff);
        if (!P.diffuseOnly) fill_cube_view(env->irradiance_specular, (const float4*)ctx->env_spec, P.spec);
    } else {
        rc = ensure_bytes(&ctx->tmp_diff, &ctx->tmp_diff_bytes, padded_texels(env->irradiance_diffuse.res, env->irradiance_diffuse.mips) * 16); if (rc) return rc;
        rc = pad_cube(env->irradiance_diffuse, (float4*)ctx->tmp_diff, stream); if (rc) return rc;
        fill_cube_view(env->irradiance_diffuse, (const float4*)ctx->tmp_diff, P.diff);
        if (!P.diffuseOnly) {
            rc = ensure_bytes(&ctx->tmp_spec, &ctx->tmp_spec_bytes, padded_texels(env->irradiance_specular.res, env->irradiance_specular.mips) * 16); if (rc) return rc;
            rc = pad_cube(env->irradiance_specular, (float4*)ctx->tmp_spec, stream); if (rc) return rc;
            fill_cube_view(env->irradiance_specular, (const float4*)ctx->tmp_spec, P.spec);
        }
    }
    P.rowBegin = row_begin; P.rows = row_end - row_begin; P.width = W;

    // grid.x covers a row in 64-pixel tiles; grid.y strides 4-row groups: about 3 resident CTAs per SM, several waves
    const unsigned gx = (unsigned)((W + FWD_BX - 1) / FWD_BX);
    unsigned gy = (unsigned)((P.rows + FWD_BY - 1) / FWD_BY);
    const unsigned targetBlocks = (unsigned)ctx->sm_count * (unsigned)FWD_MIN_BLOCKS * 4u;
    const unsigned gyCap = (targetBlocks + gx - 1) / gx;
    if (gy > gyCap) gy = gyCap < 1 ? 1 : gyCap;
    forward_kernel<<<dim3(gx, gy), dim3(FWD_BX, FWD_BY), 0, stream>>>(P);
    return vq_check_launch("forward_lighting");
}

int vq_forward_launch(VqContext* ctx, const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                      const VqGBuffer* gb, const VqEnvironmentMaps* env, VqImage out,
                      int row_begin, int row_end, cudaStream_t stream) {
    VQ_REQUIRE(gb && vq_image_ok(out) && out.height == gb->position_ao.height, "G-buffer planes and output must have the same size");
    return vq_forward_launch_multi(ctx, pf, pv, gb, env, &out, 1, 0, row_begin, row_end, stream);
}

extern "C" int vq_forward_lighting(VqContext* ctx, const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                                   const VqGBuffer* gb, const VqEnvironmentMaps* env, VqImage out,
                                   int row_begin, int row_end, void* stream) {
    int rc = vq_enter(ctx); if (rc) return rc;
    return vq_forward_launch(ctx, pf, pv, gb, env, out, row_begin, row_end, (cudaStream_t)stream);
}

extern "C" int vq_forward_lighting_multi(VqContext* ctx, const VqPerFrameData* pf, const VqPerViewLightingData* pv,
                                         const VqGBuffer* gb, const VqEnvironmentMaps* env, const VqImage* outs, int n_outs,
                                         int dst_row_offset, int row_begin, int row_end, void* stream) {
    int rc = vq_enter(ctx); if (rc) return rc;
    return vq_forward_launch_multi(ctx, pf, pv, gb, env, outs, n_outs, dst_row_offset, row_begin, row_end, (cudaStream_t)stream);
}

// The IBL cubemaps are sampled from bordered copies (see CubeV). vq_environment_prepare builds them once and
// registers them in the context: the analogue of the RENDER_TARGET -> PIXEL_SHADER_RESOURCE transition the engine
// records after prefiltering (EnvironmentMapRendering.cpp:466-472). Call it again whenever the maps' contents change.
// Without it vq_forward_lighting re-pads the cubes on every call (always correct, ~15 us slower at 512^2 x 9 mips).
extern "C" int vq_environment_prepare(VqContext* ctx, const VqEnvironmentMaps* env, void* stream) {
    int rc = vq_enter(ctx); if (rc) return rc;
    VQ_REQUIRE(env, "env is null");
    VQ_REQUIRE(cube_desc_ok(env->irradiance_diffuse), "bad cubemap descriptor (irradiance_diffuse)");
    ctx->env_valid = 0;
    rc = ensure_bytes(&ctx->env_diff, &ctx->env_diff_bytes, padded_texels(env->irradiance_diffuse.res, env->irradiance_diffuse.mips) * 16); if (rc) return rc;
    rc = pad_cube(env->irradiance_diffuse, (float4*)ctx->env_diff, (cudaStream_t)stream); if (rc) return rc;
    ctx->env_key = *env;
    if (env->irradiance_specular.ptr) {
        VQ_REQUIRE(cube_desc_ok(env->irradiance_specular), "bad cubemap descriptor (irradiance_specular)");
        rc = ensure_bytes(&ctx->env_spec, &ctx->env_spec_bytes, padded_texels(env->irradiance_specular.res, env->irradiance_specular.mips) * 16); if (rc) return rc;
        rc = pad_cube(env->irradiance_specular, (float4*)ctx->env_spec, (cudaStream_t)stream); if (rc) return rc;
    }
    ctx->env_valid = 1;
    return VQ_OK;
}
extern "C" int vq_environment_invalidate(VqContext* ctx) {
    if (!ctx) { vq_set_error("null context"); return VQ_ERR_INVALID_ARG; }
    ctx->env_valid = 0;
    return VQ_OK;
}
