// rt_math.h — small fp32 vector helpers for device code (CUDA; also compiled by the RT_EMU test build).
#pragma once
#include "rt_platform.h"

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct u4 { uint32_t x, y, z, w; };

RT_D f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
RT_D f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_D f3 mk3(float s) { return mk3(s, s, s); }
RT_D f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
RT_D f4 mk4(f3 v, float w) { return mk4(v.x, v.y, v.z, w); }
RT_D f4 mk4(float4 v) { return mk4(v.x, v.y, v.z, v.w); }
RT_D f3 xyz(f4 v) { return mk3(v.x, v.y, v.z); }
RT_D f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
RT_D float comp(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// a / b for a finite b > 0: the same bits as the IEEE division (0 / b keeps the sign of the zero).  A zero numerator (ior 1,
// no transmission: every material of the Cornell scenes) sends the hardware division down its ~35-instruction slow path.
RT_D float div_pos(float a, float b) { return a == 0.0f ? a : a / b; }

RT_D f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_D f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_D f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_D f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
RT_D f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_D f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
RT_D f3 operator/(f3 a, float s) { const float r = 1.0f / s; return mk3(a.x * r, a.y * r, a.z * r); }
RT_D f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
RT_D f3 operator-(f3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
RT_D f3 operator-(float s, f3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
RT_D f3 operator/(float s, f3 a) { return mk3(s / a.x, s / a.y, s / a.z); }
RT_D f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_D f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
RT_D f3& operator*=(f3& a, f3 b) { a = a * b; return a; }
RT_D f3& operator*=(f3& a, float s) { a = a * s; return a; }
RT_D f3& operator/=(f3& a, float s) { a = a / s; return a; }
RT_D bool is_zero(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

RT_D f4 operator+(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
RT_D f4 operator*(f4 a, f4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
RT_D f4 operator*(f4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }
RT_D f4& operator*=(f4& a, f4 b) { a = a * b; return a; }
RT_D f2 operator+(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
RT_D f2 operator*(f2 a, float s) { return mk2(a.x * s, a.y * s); }

RT_D float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_D float dot(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
RT_D f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
RT_D float length(f3 a) { return sqrtf(dot(a, a)); }
RT_D f3 normalize(f3 a) { const float r = rt_rsqrt(dot(a, a)); return mk3(a.x * r, a.y * r, a.z * r); }
RT_D f4 normalize(f4 a) { const float r = rt_rsqrt(dot(a, a)); return mk4(a.x * r, a.y * r, a.z * r, a.w * r); }
// bit-exact counterparts used by raygen so that primary rays are identical to the oracle's (debug channels and hit
// ids of primary rays then match exactly): explicitly rounded ops, oracle operation order
RT_D f3 normalize_exact(f3 a) {
    const float l = sqrtf(rt_fadd(rt_fadd(rt_fmul(a.x, a.x), rt_fmul(a.y, a.y)), rt_fmul(a.z, a.z)));
    return mk3(rt_fdiv(a.x, l), rt_fdiv(a.y, l), rt_fdiv(a.z, l));
}
RT_D f4 mat4_mul_exact(const float* M, f4 v) {
    f4 r;
    r.x = rt_fadd(rt_fadd(rt_fadd(rt_fmul(M[0], v.x), rt_fmul(M[4], v.y)), rt_fmul(M[8], v.z)), rt_fmul(M[12], v.w));
    r.y = rt_fadd(rt_fadd(rt_fadd(rt_fmul(M[1], v.x), rt_fmul(M[5], v.y)), rt_fmul(M[9], v.z)), rt_fmul(M[13], v.w));
    r.z = rt_fadd(rt_fadd(rt_fadd(rt_fmul(M[2], v.x), rt_fmul(M[6], v.y)), rt_fmul(M[10], v.z)), rt_fmul(M[14], v.w));
    r.w = rt_fadd(rt_fadd(rt_fadd(rt_fmul(M[3], v.x), rt_fmul(M[7], v.y)), rt_fmul(M[11], v.z)), rt_fmul(M[15], v.w));
    return r;
}
RT_D float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }
RT_D float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
RT_D float saturate(float x) { return clampf(x, 0.0f, 1.0f); }
RT_D f3 min3(f3 a, f3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
RT_D f3 max3(f3 a, f3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
RT_D f3 clamp3(f3 a, float lo, float hi) { return mk3(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)); }
RT_D float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
RT_D f3 mix3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
RT_D f3 pow3(f3 a, float e) { return mk3(powf(a.x, e), powf(a.y, e), powf(a.z, e)); }
RT_D f3 exp3(f3 a) { return mk3(expf(a.x), expf(a.y), expf(a.z)); }
RT_D f3 log3(f3 a) { return mk3(logf(a.x), logf(a.y), logf(a.z)); }
RT_D float smoothstepf(float e0, float e1, float x) { float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
RT_D f3 reflect3(f3 I, f3 N) { return I - 2.0f * dot(N, I) * N; }
RT_D f3 refract3(f3 I, f3 N, float eta) {   // GLSL refract: zero vector on total internal reflection
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return mk3(0.0f);
    return eta * I - (eta * d + sqrtf(k)) * N;
}
// column-major mat4 (GLSL) times vec4
RT_D f4 mat4_mul(const float* M, f4 v) {
    return mk4(M[0] * v.x + M[4] * v.y + M[8] * v.z + M[12] * v.w, M[1] * v.x + M[5] * v.y + M[9] * v.z + M[13] * v.w,
               M[2] * v.x + M[6] * v.y + M[10] * v.z + M[14] * v.w, M[3] * v.x + M[7] * v.y + M[11] * v.z + M[15] * v.w);
}
