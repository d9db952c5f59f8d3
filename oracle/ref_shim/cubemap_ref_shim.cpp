// ORACLE — TEST INFRASTRUCTURE ONLY. Links against the reference's Source/Renderer/Resources/CubemapUtility.cpp compiled in
// place (oracle/Makefile -> oracle/_ref/libvqcuberef.so; DirectXMath stand-in: ref_shim/dxmath_shim). Exposes the per-face view
// matrix the engine renders every cubemap face with (EnvironmentMapRendering.cpp:170-176: CalculateViewMatrix(face) times a
// 90-degree, aspect-1 projection) as the look direction of a face pixel, to pin oracle CubeTexelDirection.
#include "Renderer/Resources/CubemapUtility.h"

extern "C" {
// row-major XMMATRIX of face `face` (camera at the origin)
void cuberef_view_matrix(int face, float out16[16]) {
    const DirectX::XMMATRIX m = CubemapUtility::CalculateViewMatrix(face);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out16[4 * i + j] = m.r[i][j];
}
// un-normalised look direction through the centre of pixel (px, py) of an N x N face: view-space ray (ndc.x, ndc.y, 1) of the
// 90-degree projection taken back to world space with the transpose of the (orthonormal) view rotation
void cuberef_texel_direction(int face, int px, int py, int res, float out3[3]) {
    const DirectX::XMMATRIX v = CubemapUtility::CalculateViewMatrix(face);
    const float nx = 2.0f * ((float)px + 0.5f) / (float)res - 1.0f;
    const float ny = 1.0f - 2.0f * ((float)py + 0.5f) / (float)res;
    const float ray[3] = {nx, ny, 1.0f};
    for (int i = 0; i < 3; ++i) out3[i] = ray[0] * v.r[i][0] + ray[1] * v.r[i][1] + ray[2] * v.r[i][2];
}
}
