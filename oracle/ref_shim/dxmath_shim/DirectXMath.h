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
}
