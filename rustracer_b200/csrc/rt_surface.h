// rt_surface.h — RNG streams, texture fetch, vertex interpolation and the any-hit alpha test.
// Semantics: crates/examples/gltf_viewer/shaders/lib/Random.glsl, lib/RayTracingCommons.glsl:65-97,
// RayTracing.rahit:44-104 (== RayTracing.shadow.rahit) of the reference; Vulkan texel addressing at LOD 0.
#pragma once
#include "rt_scene_dev.h"

// ---- lib/Random.glsl -------------------------------------------------------------------------------
RT_D uint32_t tea16(uint32_t val0, uint32_t val1) {   // InitRandomSeed :12-25
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
RT_D float lcg_float(uint32_t& seed) {   // RandomInt/RandomFloat :27-42
    seed = 1664525u * seed + 1013904223u;
    return (float)(seed & 0x00FFFFFFu) / (float)0x01000000;
}
RT_D f2 random_in_unit_disk(uint32_t& seed) {   // :44-54
    for (;;) {
        float a = lcg_float(seed), b = lcg_float(seed);
        f2 p = mk2(2.0f * a - 1.0f, 2.0f * b - 1.0f);
        if (rt_fadd(rt_fmul(p.x, p.x), rt_fmul(p.y, p.y)) < 1.0f) return p;   // same rounding as the oracle
    }
}
RT_D f3 random_in_unit_sphere(uint32_t& seed) {   // :56-66
    for (;;) {
        float a = lcg_float(seed), b = lcg_float(seed), c = lcg_float(seed);
        f3 p = mk3(2.0f * a - 1.0f, 2.0f * b - 1.0f, 2.0f * c - 1.0f);
        if (dot(p, p) < 1.0f) return p;
    }
}
RT_D float uint_to_float01(uint32_t x) { return rt_uint_as_float(0x3f800000u | (x >> 9)) - 1.0f; }   // :70-72
RT_D uint32_t pcg4d_x(u4 v) {   // :81-98, only .x is ever consumed
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w;
    return v.x;
}
RT_D float rng_next(u4& s) { s.w++; return uint_to_float01(pcg4d_x(s)); }   // rand :102-105
RT_D uint32_t jenkins_hash(uint32_t x) { x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15; return x; }
RT_D f3 hash_and_color(uint32_t i) {
    uint32_t h = jenkins_hash(i);
    return mk3(((h >> 0) & 0xFFu) / 255.0f, ((h >> 8) & 0xFFu) / 255.0f, ((h >> 16) & 0xFFu) / 255.0f);
}
// deviation D3: order-independent BLEND draw, payload stream not advanced
RT_D float blend_random(u4 s, uint32_t instance_id, uint32_t primitive_id) {
    u4 k; k.x = s.x ^ (instance_id * 0x9E3779B9u); k.y = s.y ^ (primitive_id * 0x85EBCA6Bu); k.z = s.z ^ 0xC2B2AE35u; k.w = s.w;
    return uint_to_float01(pcg4d_x(k));
}

// ---- textures ----------------------------------------------------------------------------------------
RT_D int wrap_coord(int i, int n, uint32_t mode) {
    if (mode == RT_WRAP_REPEAT) {
        if ((n & (n - 1)) == 0) return i & (n - 1);      // power-of-two sizes: the same non-negative remainder without an integer division
        int m = i % n; return m < 0 ? m + n : m;
    }
    if (mode == RT_WRAP_MIRROR) {
        int m = i % (2 * n); if (m < 0) m += 2 * n;
        int k = m - n; k = k >= 0 ? k : -(1 + k);
        return (n - 1) - k;
    }
    return min(max(i, 0), n - 1);
}
// byte / 255.0f, bit for bit (the sequence is checked for the 256 inputs in tests/test_host.py, the textured parity cases
// compare it with the oracle's division): product with the rounded reciprocal plus one residual correction.  The IEEE division it replaces takes its ~30-instruction slow path whenever the byte is 0 (a fully
// transparent texel of a MASK texture: most of the any-hit taps in foliage).
RT_D float unorm8(uint32_t b) {
    const float x = (float)b, c = 0.0039215688593685626984f;
    const float q0 = rt_fmul(x, c);
    return rt_fma(rt_fma(-q0, 255.0f, x), c, q0);
}
RT_D f4 fetch_texel(const DScene& S, const DImage& im, int x, int y) {
    const uint32_t p = rt_ld(reinterpret_cast<const uint32_t*>(im.px) + ((size_t)y * im.w + x));
    const uint32_t r = p & 0xFF, g = (p >> 8) & 0xFF, b = (p >> 16) & 0xFF, a = p >> 24;
    if (im.srgb) return mk4(rt_ld(S.srgb_lut + r), rt_ld(S.srgb_lut + g), rt_ld(S.srgb_lut + b), unorm8(a));
    return mk4(unorm8(r), unorm8(g), unorm8(b), unorm8(a));
}
RT_D f4 sample_image(const DScene& S, const DImage& im, uint32_t filter, uint32_t ws, uint32_t wt, f2 uv) {
    if (!(isfinite(uv.x) && isfinite(uv.y)) || im.w == 0) return mk4(0, 0, 0, 0);
    float u = uv.x * (float)im.w, v = uv.y * (float)im.h;
    if (filter == RT_FILTER_NEAREST)
        return fetch_texel(S, im, wrap_coord((int)floorf(u), im.w, ws), wrap_coord((int)floorf(v), im.h, wt));
    float fu = u - 0.5f, fv = v - 0.5f;
    float i0f = floorf(fu), j0f = floorf(fv);
    float a = fu - i0f, b = fv - j0f;
    int i0 = wrap_coord((int)i0f, im.w, ws), i1 = wrap_coord((int)i0f + 1, im.w, ws);
    int j0 = wrap_coord((int)j0f, im.h, wt), j1 = wrap_coord((int)j0f + 1, im.h, wt);
    f4 t00 = fetch_texel(S, im, i0, j0), t10 = fetch_texel(S, im, i1, j0), t01 = fetch_texel(S, im, i0, j1), t11 = fetch_texel(S, im, i1, j1);
    f4 top = t00 * (1.0f - a) + t10 * a;
    f4 bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
RT_D f4 texture2d(const DScene& S, int tex_index, f2 uv) {
    if (tex_index < 0 || (uint32_t)tex_index >= S.n_textures) return mk4(1, 1, 1, 1);
    const DTexture t = S.textures[tex_index];
    const DImage im = S.images[t.image];
    return sample_image(S, im, t.mag_filter, t.wrap_s, t.wrap_t, uv);
}
RT_D f3 texture_cube(const DScene& S, f3 r) {   // Vulkan major-axis face selection; see oracle note on face edges
    float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    int face; float sc, tc, ma;
    if (az >= ax && az >= ay) { ma = az; if (r.z >= 0) { face = 4; sc = r.x; tc = -r.y; } else { face = 5; sc = -r.x; tc = -r.y; } }
    else if (ay >= ax)        { ma = ay; if (r.y >= 0) { face = 2; sc = r.x; tc = r.z; } else { face = 3; sc = r.x; tc = -r.z; } }
    else                      { ma = ax; if (r.x >= 0) { face = 0; sc = -r.z; tc = -r.y; } else { face = 1; sc = r.z; tc = -r.y; } }
    f2 uv = mk2(0.5f * (sc / ma) + 0.5f, 0.5f * (tc / ma) + 0.5f);
    return xyz(sample_image(S, S.sky[face], RT_FILTER_LINEAR, RT_WRAP_CLAMP, RT_WRAP_CLAMP, uv));
}

// ---- vertex fetch / interpolation (RayTracingCommons.glsl:65-97) ---------------------------------------
RT_D f2 get_uv(f4 uv0and1, int index) {
    if (index == 0) return mk2(uv0and1.x, uv0and1.y);
    if (index == 1) return mk2(uv0and1.z, uv0and1.w);
    return mk2(0.0f, 0.0f);
}
RT_D f4 ld_f4(const float* p) { const float4 v = rt_ld(reinterpret_cast<const float4*>(p)); return mk4(v.x, v.y, v.z, v.w); }
RT_D f3 ld_f3(const float* p) { const float4 v = rt_ld(reinterpret_cast<const float4*>(p)); return mk3(v.x, v.y, v.z); }   // 16-byte aligned vec3+pad

struct TriIndices { uint32_t i0, i1, i2; };
RT_D TriIndices fetch_indices(const DScene& S, const rt_prim_info& pi, uint32_t prim) {
    const uint32_t io = pi.i_offset + 3 * prim;
    TriIndices t; t.i0 = pi.v_offset + rt_ld(S.indices + io); t.i1 = pi.v_offset + rt_ld(S.indices + io + 1); t.i2 = pi.v_offset + rt_ld(S.indices + io + 2);
    return t;
}

// ---- any-hit alpha test (RayTracing.rahit:44-104); true = ignore the candidate --------------------------
RT_D bool anyhit_ignore(const DScene& S, uint32_t instance_id, uint32_t primitive_id, uint32_t geo_id, float bu, float bv, u4 rng, unsigned long long* taps = nullptr) {
    const rt_prim_info pi = S.prim_infos[geo_id];
    const rt_material& mat = S.materials[pi.material_id];
    const uint32_t alpha_mode = mat.alpha_mode;
    if (alpha_mode == 1) return false;
    const TriIndices ti = fetch_indices(S, pi, primitive_id);
    const rt_vertex& v0 = S.vertices[ti.i0]; const rt_vertex& v1 = S.vertices[ti.i1]; const rt_vertex& v2 = S.vertices[ti.i2];
    const float b0 = 1.0f - bu - bv;
    f4 uv = mk4(v0.uv0[0], v0.uv0[1], v0.uv1[0], v0.uv1[1]) * b0 + mk4(v1.uv0[0], v1.uv0[1], v1.uv1[0], v1.uv1[1]) * bu + mk4(v2.uv0[0], v2.uv0[1], v2.uv1[0], v2.uv1[1]) * bv;
    f4 vcolor = ld_f4(v0.color) * b0 + ld_f4(v1.color) * bu + ld_f4(v2.color) * bv;
    f4 color4 = vcolor * ld_f4(mat.base_color);
    if (mat.base_color_texture.index >= 0) { if (taps) ++*taps; color4 *= texture2d(S, mat.base_color_texture.index, get_uv(uv, mat.base_color_texture.coord)); }
    float opacity = color4.w;
    if (mat.workflow == 1) {
        f4 diffuse_factor = ld_f4(mat.sg_diffuse_factor);
        if (mat.sg_diffuse_texture.index >= 0) { if (taps) ++*taps; diffuse_factor *= texture2d(S, mat.sg_diffuse_texture.index, get_uv(uv, mat.sg_diffuse_texture.coord)); }
        opacity = (vcolor * diffuse_factor).w;
    }
    if (alpha_mode == 2) return opacity < mat.alpha_cutoff;
    return opacity <= blend_random(rng, instance_id, primitive_id);
}
