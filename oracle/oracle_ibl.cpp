// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// oracle_ibl.cpp — scalar restatement of the environment-map integrals
// (Shaders/CubemapConvolution.hlsl:112-223) and the CPU HDRI mip filter
// (Source/Renderer/Resources/DXGIUtils.cpp:289-317).
// Accumulation is the HLSL's: one sequential fp32 running sum per texel.
#include "oracle.h"

namespace orc {

// DXGIUtils.cpp:289-317, bytesPerPixel == 16 branch: RGB = min of the 2x2 block, A = 1.
// The reference loops `y < height; y += 2` and reads (x+1, y+1): for an odd width/height that reads (and,
// through outY, writes) out of bounds. Its HDRIs are power-of-two sized so it never happens there; the
// restatement visits only complete 2x2 blocks (destination = floor(w/2) x floor(h/2), the size the
// reference allocates), which is identical for even sizes.
void MipImage_MinFilter(const float* src, float* dst, int width, int height) {
    const int offsetsX[] = {0, 1, 0, 1};
    const int offsetsY[] = {0, 0, 1, 1};
    for (int y = 0; y + 1 < height; y += 2)
        for (int x = 0; x + 1 < width; x += 2) {
            float rgb[4][3];
            for (int smp = 0; smp < 4; ++smp)
                for (int ch = 0; ch < 3; ++ch)
                    rgb[smp][ch] = src[(size_t)(x + offsetsX[smp]) * 4 + (size_t)(y + offsetsY[smp]) * 4 * width + ch];
            float f[4];
            for (int ch = 0; ch < 3; ++ch)
                f[ch] = std::min(rgb[0][ch], std::min(rgb[1][ch], std::min(rgb[2][ch], rgb[3][ch])));
            f[3] = 1.0f;
            const int outX = x >> 1, outY = y >> 1;
            for (int ch = 0; ch < 4; ++ch)
                dst[(size_t)outX * 4 + 4 * (size_t)outY * (width >> 1) + ch] = f[ch];
        }
}

// The (phi, theta) sequences of the two loops at CubemapConvolution.hlsl:129-135.
//   step > 0 : the reference's float-accumulated loop variable (phi += step), so the trip count and
//              the values depend on fp32 rounding (0.010 -> 629 x 158)
//   step == 0: integer grid, phi_i = float(i) * (TWO_PI / n_phi), theta_j = float(j) * (PI_OVER_TWO / n_theta)
void DiffuseIrradianceAngles(float step, int n_phi, int n_theta, std::vector<float>& phis, std::vector<float>& thetas) {
    phis.clear(); thetas.clear();
    if (step > 0.0f) {
        for (float phi = 0.0f; phi < TWO_PI; phi += step) phis.push_back(phi);
        for (float theta = 0.0f; theta < PI_OVER_TWO; theta += step) thetas.push_back(theta);
    } else {
        const float dphi = TWO_PI / (float)n_phi;
        const float dtheta = PI_OVER_TWO / (float)n_theta;
        for (int i = 0; i < n_phi; ++i) phis.push_back((float)i * dphi);
        for (int j = 0; j < n_theta; ++j) thetas.push_back((float)j * dtheta);
    }
}

// CubemapConvolution.hlsl:112-163
float4 DiffuseIrradiance_PSMain(const Pyramid& hdri, float3 lookDir, const std::vector<float>& phis,
                                const std::vector<float>& thetas, int srcMip, bool f64Accum) {
    const float3 N = normalize(lookDir);
    float3 up = make3(0, 1, 0);
    const float3 right = normalize(cross(up, N));
    up = normalize(cross(N, right));

    float3 irradiance = splat3(0.0f);
    double acc64[3] = {0.0, 0.0, 0.0};   // diagnostic only (f64Accum): the same fp32 terms summed without fp32 rounding
    float numSamples = 0.0f;
    for (float phi : phis) {
        for (float theta : thetas) {
            const float sinTheta = std::sin(theta);
            const float cosTheta = std::cos(theta);
            const float sinPhi = std::sin(phi);
            const float cosPhi = std::cos(phi);
            const float3 tangentSample = make3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
            float3 sampleVec = right * tangentSample.x + up * tangentSample.y + N * tangentSample.z;
            sampleVec = normalize(sampleVec);
            const float mipLevel = (float)srcMip;
            const float3 L = xyz(SampleEquirectLevel(hdri, DirectionToEquirectUV(sampleVec), mipLevel));
            const float3 term = L * cosTheta * sinTheta;
            irradiance += term;
            acc64[0] += term.x; acc64[1] += term.y; acc64[2] += term.z;
            numSamples += 1.0f;
        }
    }
    if (f64Accum) irradiance = make3((float)acc64[0], (float)acc64[1], (float)acc64[2]);
    irradiance = irradiance * PI / numSamples;
    return make4(irradiance, 1.0f);
}

// CubemapConvolution.hlsl:168-223
// sampleLengthScale (test infrastructure, 1 = the reference): scales the UN-normalised sample vector L before it is mapped to
// the equirect uv. The reference feeds reflect(-V, H) straight into asin(-L.y): |L| = 1 only up to rounding, and near a pole
// (|L.y| -> 1, infinite slope of asin) one ulp of |L| moves v by a fraction of a texel row — with the WRAP-in-v sampler blending
// in the opposite pole's row. The conditioning probe of the full-size parity test varies exactly this degree of freedom.
float4 SpecularIrradiance_PSMain(const Pyramid& hdri, float3 lookDir, float Roughness,
                                 float2 TextureDimensionsLOD0, uint32_t NUM_SAMPLES, float sampleLengthScale) {
    const float3 N = normalize(lookDir);
    const float3 R = N;
    const float3 V = R;
    float3 prefilteredColor = splat3(0.0f);
    float totalWeight = 0.0f;
    for (uint32_t i = 0; i < NUM_SAMPLES; ++i) {
        const float2 Xi = Hammersley(i, NUM_SAMPLES);
        const float3 H = ImportanceSampleGGX(Xi, N, Roughness);
        const float3 L = reflect(-V, H);
        const float NdotL = saturate(dot(N, L));
        if (NdotL > 0.0f) {
            const float NdotH = saturate(dot(N, H));
            const float HdotV = saturate(dot(H, V));
            const float D = NormalDistributionGGX(NdotH, Roughness);
            const float pdf = (D * NdotH / (4.0f * HdotV));
            const float fOmegaS = 1.0f / (std::fmax((float)NUM_SAMPLES * pdf, 0.00001f));
            const float fOmegaP = 4.0f * PI / (6.0f * TextureDimensionsLOD0.x * TextureDimensionsLOD0.y);
            const float fMipBias = -1.0f;
            const float fMipLevel = Roughness == 0.0f ? 0.0f
                                  : std::fmax(0.5f * std::log2(fOmegaS / fOmegaP) + fMipBias, 0.0f);
            const float3 Ls = sampleLengthScale == 1.0f ? L : L * sampleLengthScale;
            prefilteredColor += xyz(SampleEquirectLevel(hdri, DirectionToEquirectUV(Ls), fMipLevel)) * NdotL;
            totalWeight += NdotL;
        }
    }
    prefilteredColor = prefilteredColor / std::fmax(totalWeight, 0.0001f);
    return make4(prefilteredColor, 1.0f);
}

}  // namespace orc
