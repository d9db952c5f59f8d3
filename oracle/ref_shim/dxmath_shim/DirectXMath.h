// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for <DirectXMath.h> with just the storage types the reference's
// Shaders/VQPlatform.h maps the HLSL vector names to (public, documented layouts: XMFLOATn = n packed floats, XMMATRIX =
// four 16-byte-aligned rows), so that Shaders/LightingConstantBufferData.h compiles UNMODIFIED as the CPU side (VQ_CPU) of the
// shared CPU/GPU struct header. Used only to compare struct layouts (tests/test_struct_layout_ref.py).
#pragma once
#include <cstdint>
namespace DirectX {
struct XMFLOAT2 { float x, y; XMFLOAT2() = default; constexpr XMFLOAT2(float a, float b) : x(a), y(b) {} };
struct XMFLOAT3 { float x, y, z; XMFLOAT3() = default; constexpr XMFLOAT3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct XMFLOAT4 { float x, y, z, w; XMFLOAT4() = default; constexpr XMFLOAT4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} };
struct XMINT2 { int32_t x, y; XMINT2() = default; constexpr XMINT2(int32_t a, int32_t b) : x(a), y(b) {} };
struct XMINT3 { int32_t x, y, z; };
struct XMINT4 { int32_t x, y, z, w; };
struct XMUINT2 { uint32_t x, y; };
struct XMUINT3 { uint32_t x, y, z; };
struct XMUINT4 { uint32_t x, y, z, w; };
struct alignas(16) XMMATRIX { float r[4][4]; };
struct XMFLOAT3X3 { float m[3][3]; };
struct XMFLOAT4X3 { float m[4][3]; };
struct XMFLOAT3X4 { float m[3][4]; };
struct XMFLOAT4X4 { float m[4][4]; };

// --- the few functions Source/Renderer/Resources/CubemapUtility.cpp calls, with DirectXMath's DOCUMENTED semantics (public API
// of Microsoft's MIT-licensed library, which is not part of the reference tree): row-vector convention, v' = v * M.
constexpr float XM_PI = 3.141592654f;
constexpr float XM_PIDIV2 = 1.570796327f;
struct XMVECTOR { float f[4]; };
inline XMVECTOR XMLoadFloat3(const XMFLOAT3* p) { return XMVECTOR{{p->x, p->y, p->z, 0.0f}}; }
inline XMVECTOR operator+(XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[0] + b.f[0], a.f[1] + b.f[1], a.f[2] + b.f[2], a.f[3] + b.f[3]}}; }
inline XMMATRIX XMMatrixIdentity() { XMMATRIX m{}; for (int i = 0; i < 4; ++i) m.r[i][i] = 1.0f; return m; }
// XMMatrixLookAtLH(eye, at, up): zaxis = normalize(at - eye); xaxis = normalize(cross(up, zaxis)); yaxis = cross(zaxis, xaxis);
// columns of the rotation part are the axes, last row = (-dot(xaxis,eye), -dot(yaxis,eye), -dot(zaxis,eye), 1)
inline XMMATRIX XMMatrixLookAtLH(XMVECTOR eye, XMVECTOR at, XMVECTOR up) {
    auto sub = [](XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[0] - b.f[0], a.f[1] - b.f[1], a.f[2] - b.f[2], 0.0f}}; };
    auto dot = [](XMVECTOR a, XMVECTOR b) { return a.f[0] * b.f[0] + a.f[1] * b.f[1] + a.f[2] * b.f[2]; };
    auto cross = [](XMVECTOR a, XMVECTOR b) { return XMVECTOR{{a.f[1] * b.f[2] - a.f[2] * b.f[1], a.f[2] * b.f[0] - a.f[0] * b.f[2], a.f[0] * b.f[1] - a.f[1] * b.f[0], 0.0f}}; };
    auto norm = [&](XMVECTOR a) { const float l = __builtin_sqrtf(dot(a, a)); return XMVECTOR{{a.f[0] / l, a.f[1] / l, a.f[2] / l, 0.0f}}; };
    const XMVECTOR z = norm(sub(at, eye)), x = norm(cross(up, z)), y = cross(z, x);
    XMMATRIX m{};
    for (int i = 0; i < 3; ++i) { m.r[i][0] = x.f[i]; m.r[i][1] = y.f[i]; m.r[i][2] = z.f[i]; }
    m.r[3][0] = -dot(x, eye); m.r[3][1] = -dot(y, eye); m.r[3][2] = -dot(z, eye); m.r[3][3] = 1.0f;
    return m;
}
}
