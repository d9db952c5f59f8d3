// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked into libvqcuda.
//
// hlsl_math.h — the HLSL vector types and intrinsics the reference shaders use, as scalar fp32
// C++ (SURVEY.md §9 "HLSL intrinsics"). Compile with -ffp-contract=off so that a*b+c is two
// roundings, exactly as written.
//
// Semantics fixed here (decisions, since D3D does not specify bit-exact results):
//   saturate(x)   = min(max(x,0),1)
//   lerp(a,b,t)   = a + t*(b-a)
//   reflect(i,n)  = i - 2*n*dot(i,n)
//   normalize(v)  = v / sqrt(dot(v,v))
//   pow(x,y)      = powf
//   mul(v, M)     = row-vector * matrix
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

inline float2 make2(float x, float y) { return {x, y}; }
inline float3 make3(float x, float y, float z) { return {x, y, z}; }
inline float3 splat3(float s) { return {s, s, s}; }
inline float4 make4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline float4 make4(float3 v, float w) { return {v.x, v.y, v.z, w}; }
inline float3 xyz(float4 v) { return {v.x, v.y, v.z}; }

inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator/(float3 a, float3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, float3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float3 operator/(float3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }

inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator*(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 operator*(float2 a, float s) { return {a.x * s, a.y * s}; }
inline float2 operator/(float2 a, float2 b) { return {a.x / b.x, a.y / b.y}; }
inline float2 operator/(float2 a, float s) { return {a.x / s, a.y / s}; }

inline float4 operator+(float4 a, float4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline float4 operator-(float4 a, float4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline float4 operator*(float4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

// HLSL dot on float3: x*x' + y*y' + z*z', summed left to right.
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float3 cross(float3 a, float3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(float3 v) { return std::sqrt(dot(v, v)); }
inline float3 normalize(float3 v) { return v / std::sqrt(dot(v, v)); }
inline float saturate(float x) { return std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * t; }
inline float3 reflect(float3 i, float3 n) { return i - n * (2.0f * dot(i, n)); }
inline float3 max3(float3 a, float3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline float3 min3(float3 a, float3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline float3 abs3(float3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline float3 pow3(float3 a, float e) { return {std::pow(a.x, e), std::pow(a.y, e), std::pow(a.z, e)}; }

inline uint32_t asuint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float asfloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// reference: Shaders/ShadingMath.hlsl:25-29
constexpr float PI = 3.14159265359f;
constexpr float TWO_PI = 6.28318530718f;
constexpr float PI_OVER_TWO = 1.5707963268f;

}  // namespace orc
