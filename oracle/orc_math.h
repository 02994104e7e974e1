// TEST INFRASTRUCTURE — CPU oracle for the rustracer path-tracing hot path.  Not part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// orc_math.h: GLSL-flavoured vector maths used by the restatement.  Semantics follow the GLSL 4.60
// built-ins the reference shaders call: normalize, reflect, refract (zero on total
// internal reflection), mix, clamp, smoothstep, floatBitsToInt / intBitsToFloat.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct vec4 { float x, y, z, w; float& operator[](int i) { return (&x)[i]; } float operator[](int i) const { return (&x)[i]; } };
struct uvec4 { uint32_t x, y, z, w; };

inline vec2 V2(float x, float y) { return {x, y}; }
inline vec3 V3(float x, float y, float z) { return {x, y, z}; }
inline vec3 V3(float s) { return {s, s, s}; }
inline vec4 V4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline vec4 V4(vec3 v, float w) { return {v.x, v.y, v.z, w}; }
inline vec3 xyz(vec4 v) { return {v.x, v.y, v.z}; }

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator*(float s, vec2 a) { return {a.x * s, a.y * s}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }
inline vec2 operator-(vec2 a, float s) { return {a.x - s, a.y - s}; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3 operator/(float s, vec3 a) { return {s / a.x, s / a.y, s / a.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline bool operator==(vec3 a, vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(vec3 a, vec3 b) { return !(a == b); }

inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(float s, vec4 a) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4& operator*=(vec4& a, vec4 b) { a = a * b; return a; }

inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec4 normalize(vec4 a) { float l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l, a.w / l}; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

inline float clampf(float x, float lo, float hi) { return std::min(std::max(x, lo), hi); }
inline float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
inline vec3 min3(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 max3(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline vec3 clamp3(vec3 a, float lo, float hi) { return {clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)}; }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix3(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 pow3(vec3 a, float e) { return {std::pow(a.x, e), std::pow(a.y, e), std::pow(a.z, e)}; }
inline vec3 exp3(vec3 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }
inline vec3 log3(vec3 a) { return {std::log(a.x), std::log(a.y), std::log(a.z)}; }
inline float smoothstep(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }
inline vec3 refract(vec3 I, vec3 N, float eta) {
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return V3(0.0f);
    return eta * I - (eta * d + std::sqrt(k)) * N;
}
inline int32_t floatBitsToInt(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline float intBitsToFloat(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
inline float uintBitsToFloat(uint32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

// column-major 4x4 (GLSL mat4 / nalgebra / glam storage)
struct mat4 { float m[16]; };
inline vec4 mul(const mat4& M, vec4 v) {
    // GLSL: M * v = col0*v.x + col1*v.y + col2*v.z + col3*v.w
    vec4 r;
    for (int i = 0; i < 4; ++i)
        r[i] = M.m[0 + i] * v.x + M.m[4 + i] * v.y + M.m[8 + i] * v.z + M.m[12 + i] * v.w;
    return r;
}

}  // namespace orc
