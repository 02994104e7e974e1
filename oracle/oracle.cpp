// TEST INFRASTRUCTURE — CPU oracle for the rustracer path-tracing hot path.
// Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library.  The product (rustracer_b200/csrc) never links it.
//
// PARITY UNPINNED BY THE REFERENCE: the reference ships no golden vector, known-answer test or image for
// this path (3 smoke tests, none numeric — SURVEY.md §4), cannot be built here (no cargo/rustc/glslc/Vulkan)
// and is itself non-deterministic (clockARB() seeds).  This file is therefore a line-by-line CPU
// restatement of the reference's GLSL with the documented deviations D1-D3/D8:
//   RayTracing.rgen:25-167, RayTracing.rchit:28-477, RayTracing.rahit:44-110, RayTracing.rmiss:14-45,
//   RayTracing.shadow.rmiss, lib/Random.glsl, lib/PBR.glsl (live functions), lib/Material.glsl:78-89,
//   lib/PunctualLight.glsl:17-30, lib/RayTracingCommons.glsl:32-119, lib/Tonemapping.glsl, lib/Heatmap.glsl,
//   lib/Camera.glsl, AnimationCompute.comp    (paths relative to crates/examples/gltf_viewer/shaders/)
// plus the part the reference delegates to the Vulkan driver (traceRayEXT): a two-level BVH2 and the
// Woop/Benthin/Wald watertight ray-triangle test (JCGT 2013) in fp32 with a fixed operation order.
//
// Build: see oracle/Makefile (-O2 -ffp-contract=off: no fused multiply-add may be formed, the CUDA side uses
// the matching round-to-nearest intrinsics so intersection results are bit-identical).
#include "../include/rt_b200.h"
#include "orc_math.h"

#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <chrono>
#include <atomic>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

// =====================================================================================================
// RNG — lib/Random.glsl
// =====================================================================================================
// Random.glsl:12-25  (tea, 16 rounds)
uint32_t InitRandomSeed(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
// Random.glsl:27-42
uint32_t RandomInt(uint32_t& seed) { return (seed = 1664525u * seed + 1013904223u); }
float RandomFloat(uint32_t& seed) { return (float)(RandomInt(seed) & 0x00FFFFFFu) / (float)0x01000000; }
// Random.glsl:44-54
vec2 RandomInUnitDisk(uint32_t& seed) {
    for (;;) {
        float a = RandomFloat(seed), b = RandomFloat(seed);
        vec2 p = 2.0f * V2(a, b) - 1.0f;
        if (dot(p, p) < 1.0f) return p;
    }
}
// Random.glsl:56-66
vec3 RandomInUnitSphere(uint32_t& seed) {
    for (;;) {
        float a = RandomFloat(seed), b = RandomFloat(seed), c = RandomFloat(seed);
        vec3 p = 2.0f * V3(a, b, c) - 1.0f;
        if (dot(p, p) < 1.0f) return p;
    }
}
// Random.glsl:70-72
float uintToFloat(uint32_t x) { return uintBitsToFloat(0x3f800000u | (x >> 9)) - 1.0f; }
// Random.glsl:81-98
uvec4 pcg4d(uvec4 v) {
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16; v.y ^= v.y >> 16; v.z ^= v.z >> 16; v.w ^= v.w >> 16;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    return v;
}
// Random.glsl:102-105
float rnd(uvec4& s) { s.w++; return uintToFloat(pcg4d(s).x); }
// Random.glsl:108-115
uint32_t jenkinsHash(uint32_t x) {
    x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15; return x;
}
// Random.glsl:119-125
vec3 hashAndColor(uint32_t i) {
    uint32_t h = jenkinsHash(i);
    return V3(((h >> 0) & 0xFFu) / 255.0f, ((h >> 8) & 0xFFu) / 255.0f, ((h >> 16) & 0xFFu) / 255.0f);
}
// Deviation D3: RayTracing.rahit:101 draws rand(Ray.rngState) once per BLEND candidate in driver-defined
// order.  We draw an order-independent number from the payload state and the candidate's ids and do not
// advance the payload stream.
float blendRandom(uvec4 s, uint32_t instance_id, uint32_t primitive_id) {
    uvec4 k = {s.x ^ (instance_id * 0x9E3779B9u), s.y ^ (primitive_id * 0x85EBCA6Bu), s.z ^ 0xC2B2AE35u, s.w};
    return uintToFloat(pcg4d(k).x);
}

// =====================================================================================================
// Scene storage
// =====================================================================================================
struct Aabb {
    vec3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(vec3 p) { lo = min3(lo, p); hi = max3(hi, p); }
    void grow(const Aabb& b) { lo = min3(lo, b.lo); hi = max3(hi, b.hi); }
    float area() const { vec3 e = hi - lo; return e.x * e.y + e.y * e.z + e.z * e.x; }
};

struct BNode { Aabb box; uint32_t left, count; };  // count > 0: leaf over prim[left, left+count)

struct Bvh {
    std::vector<BNode> nodes;
    std::vector<uint32_t> prim;
    void build(const std::vector<Aabb>& boxes);
private:
    void subdivide(uint32_t ni, const std::vector<Aabb>& boxes, const std::vector<vec3>& cent);
};

void Bvh::build(const std::vector<Aabb>& boxes) {
    nodes.clear(); prim.resize(boxes.size());
    for (size_t i = 0; i < boxes.size(); ++i) prim[i] = (uint32_t)i;
    if (boxes.empty()) return;
    std::vector<vec3> cent(boxes.size());
    for (size_t i = 0; i < boxes.size(); ++i) cent[i] = (boxes[i].lo + boxes[i].hi) * 0.5f;
    nodes.reserve(boxes.size() * 2);
    BNode root; root.left = 0; root.count = (uint32_t)boxes.size();
    for (auto& b : boxes) root.box.grow(b);
    nodes.push_back(root);
    subdivide(0, boxes, cent);
}

void Bvh::subdivide(uint32_t ni, const std::vector<Aabb>& boxes, const std::vector<vec3>& cent) {
    // binned SAH, 16 bins; quality is irrelevant to results (traversal is conservative + tie-broken by id)
    std::vector<uint32_t> todo{ni};
    while (!todo.empty()) {
        uint32_t n = todo.back(); todo.pop_back();
        uint32_t first = nodes[n].left, count = nodes[n].count;
        if (count <= 4) continue;
        Aabb cb;
        for (uint32_t i = 0; i < count; ++i) cb.grow(cent[prim[first + i]]);
        const int NB = 16;
        int bestAxis = -1, bestSplit = 0; float bestCost = 1e30f;
        for (int ax = 0; ax < 3; ++ax) {
            float lo = cb.lo[ax], hi = cb.hi[ax];
            if (!(hi > lo)) continue;
            Aabb bb[NB]; uint32_t bc[NB] = {0};
            float scale = NB / (hi - lo);
            for (uint32_t i = 0; i < count; ++i) {
                uint32_t p = prim[first + i];
                int b = std::min(NB - 1, (int)((cent[p][ax] - lo) * scale));
                bb[b].grow(boxes[p]); bc[b]++;
            }
            float la[NB - 1], ra[NB - 1]; uint32_t lc[NB - 1], rc[NB - 1];
            Aabb l, r; uint32_t ls = 0, rs = 0;
            for (int i = 0; i < NB - 1; ++i) {
                ls += bc[i]; if (bc[i]) l.grow(bb[i]); lc[i] = ls; la[i] = ls ? l.area() : 0.0f;
                rs += bc[NB - 1 - i]; if (bc[NB - 1 - i]) r.grow(bb[NB - 1 - i]); rc[NB - 2 - i] = rs; ra[NB - 2 - i] = rs ? r.area() : 0.0f;
            }
            for (int i = 0; i < NB - 1; ++i) {
                if (!lc[i] || !rc[i]) continue;
                float c = lc[i] * la[i] + rc[i] * ra[i];
                if (c < bestCost) { bestCost = c; bestAxis = ax; bestSplit = i; }
            }
        }
        uint32_t mid;
        if (bestAxis < 0) {
            mid = first + count / 2;  // all centroids coincide: split in the middle
        } else {
            float lo = cb.lo[bestAxis], hi = cb.hi[bestAxis], scale = NB / (hi - lo);
            auto it = std::partition(prim.begin() + first, prim.begin() + first + count, [&](uint32_t p) {
                int b = std::min(NB - 1, (int)((cent[p][bestAxis] - lo) * scale));
                return b <= bestSplit;
            });
            mid = (uint32_t)(it - prim.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        BNode l, r;
        l.left = first; l.count = mid - first;
        r.left = mid; r.count = first + count - mid;
        for (uint32_t i = l.left; i < l.left + l.count; ++i) l.box.grow(boxes[prim[i]]);
        for (uint32_t i = r.left; i < r.left + r.count; ++i) r.box.grow(boxes[prim[i]]);
        uint32_t li = (uint32_t)nodes.size();
        nodes.push_back(l); nodes.push_back(r);
        nodes[n].left = li; nodes[n].count = 0;
        todo.push_back(li); todo.push_back(li + 1);
    }
}

struct Image { std::vector<uint8_t> px; uint32_t w = 0, h = 0, srgb = 0; };

struct Geometry {
    uint32_t v_offset, i_offset, v_len, i_len, opaque, material_id;
    Bvh bvh;
};

struct Instance {
    float o2w[12];   // row-major 3x4
    float w2o[12];
    uint32_t geo_id;
    // Baked instance (rule shared with the CUDA core, rt_core.h classify_instances): when the geometry is referenced
    // by exactly one instance or has at most 256 triangles, the triangles are transformed to world space once
    // (fp32, fixed operation order) and the world ray is intersected directly — no per-ray transform.
    bool baked = false;
    std::vector<vec3> wpos;   // 3 world-space vertices per triangle
    Bvh wbvh;
};

}  // namespace

// Test helper (SURVEY.md §8d ray set iv): while a recorder is attached, raygen() / castShadowRay() append every traced
// segment of bounce >= min_bounce (and every shadow ray) together with the payload RNG state the any-hit stage would see.
struct RayRecorder {
    uint32_t min_bounce = 0; size_t max_rays = 0, max_shadow = 0;
    std::vector<rt_ray> rays, shadow; std::vector<uint32_t> rng, shadow_rng;
};

struct orc_scene {
    RayRecorder* rec = nullptr;
    std::vector<rt_vertex> vertices_in;   // as uploaded
    std::vector<rt_vertex> vertices;      // after skinning (what BLASes and shading read)
    std::vector<uint32_t> indices;
    std::vector<rt_prim_info> prim_infos;
    std::vector<Geometry> geos;
    std::vector<rt_material> materials;
    std::vector<Instance> instances;
    std::vector<Image> images;
    std::vector<rt_sampler_desc> samplers;
    std::vector<rt_texture_desc> textures;
    std::vector<rt_light> dlights, plights;
    std::vector<float> skins;
    Image sky[6]; bool has_sky_faces = false;
    Bvh tlas;
    float srgb_lut[256];
    // statistics of the last render
    mutable std::atomic<uint64_t> rays_extend{0}, rays_shadow{0}, shaded{0}, tex_taps{0}, light_cands{0};
};

namespace {

thread_local std::string g_err;

// =====================================================================================================
// Instance transform inversion (double, fixed formula; the CUDA library uses the same formula)
// =====================================================================================================
void invert_3x4(const float* m, float* out) {
    double a00 = m[0], a01 = m[1], a02 = m[2], t0 = m[3];
    double a10 = m[4], a11 = m[5], a12 = m[6], t1 = m[7];
    double a20 = m[8], a21 = m[9], a22 = m[10], t2 = m[11];
    double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    double det = a00 * c00 + a01 * c01 + a02 * c02;
    double id = 1.0 / det;
    double i00 = c00 * id, i01 = (a02 * a21 - a01 * a22) * id, i02 = (a01 * a12 - a02 * a11) * id;
    double i10 = c01 * id, i11 = (a00 * a22 - a02 * a20) * id, i12 = (a02 * a10 - a00 * a12) * id;
    double i20 = c02 * id, i21 = (a01 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a01 * a10) * id;
    out[0] = (float)i00; out[1] = (float)i01; out[2] = (float)i02; out[3] = (float)(-(i00 * t0 + i01 * t1 + i02 * t2));
    out[4] = (float)i10; out[5] = (float)i11; out[6] = (float)i12; out[7] = (float)(-(i10 * t0 + i11 * t1 + i12 * t2));
    out[8] = (float)i20; out[9] = (float)i21; out[10] = (float)i22; out[11] = (float)(-(i20 * t0 + i21 * t1 + i22 * t2));
}

// row-major 3x4 times (p,1) / (d,0); fixed order ((m0*x + m1*y) + m2*z) + m3, no fma
inline vec3 xform_point(const float* m, vec3 p) {
    return V3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
              ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
              ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
inline vec3 xform_dir(const float* m, vec3 d) {
    return V3((m[0] * d.x + m[1] * d.y) + m[2] * d.z,
              (m[4] * d.x + m[5] * d.y) + m[6] * d.z,
              (m[8] * d.x + m[9] * d.y) + m[10] * d.z);
}

// =====================================================================================================
// Watertight ray/triangle intersection — Woop, Benthin, Wald, JCGT 2(1) 2013, fp32, double fallback on
// exact-zero edge functions.  The reference delegates this to traceRayEXT (RayTracing.rgen:89-92).
// =====================================================================================================
struct RayShear {
    int kx, ky, kz; float Sx, Sy, Sz; vec3 o;
    void init(vec3 org, vec3 d) {
        o = org;
        kz = 0;
        if (std::fabs(d.y) > std::fabs(d[kz])) kz = 1;
        if (std::fabs(d.z) > std::fabs(d[kz])) kz = 2;
        kx = kz + 1; if (kx == 3) kx = 0;
        ky = kx + 1; if (ky == 3) ky = 0;
        // kx/ky swap of the paper omitted: result-neutral for the two-sided test (see rt_traverse.h)
        Sx = d[kx] / d[kz]; Sy = d[ky] / d[kz]; Sz = 1.0f / d[kz];
    }
};

inline bool tri_test(const RayShear& r, vec3 v0, vec3 v1, vec3 v2, float tmin, float tmax, float& t, float& bu, float& bv) {
    const vec3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
    const float Ax = A[r.kx] - r.Sx * A[r.kz], Ay = A[r.ky] - r.Sy * A[r.kz];
    const float Bx = B[r.kx] - r.Sx * B[r.kz], By = B[r.ky] - r.Sy * B[r.kz];
    const float Cx = C[r.kx] - r.Sx * C[r.kz], Cy = C[r.ky] - r.Sy * C[r.kz];
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx; U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx; V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax; W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float Az = r.Sz * A[r.kz], Bz = r.Sz * B[r.kz], Cz = r.Sz * C[r.kz];
    const float T = (U * Az + V * Bz) + W * Cz;
    const float inv = 1.0f / det;
    const float tt = T * inv;
    if (!(tt > tmin && tt < tmax)) return false;   // open interval (Vulkan: tMin < t < tMax)
    t = tt; bu = V * inv; bv = W * inv;
    return true;
}

// conservative slab test (never rejects a box the ray touches)
inline bool box_test(const Aabb& b, vec3 o, vec3 d, float tmin, float tmax, float& tnear) {
    float t0 = tmin, t1 = tmax;
    for (int a = 0; a < 3; ++a) {
        if (d[a] == 0.0f) { if (o[a] < b.lo[a] || o[a] > b.hi[a]) return false; continue; }
        float inv = 1.0f / d[a];
        float ta = (b.lo[a] - o[a]) * inv, tb = (b.hi[a] - o[a]) * inv;
        if (ta > tb) std::swap(ta, tb);
        ta -= std::fabs(ta) * 1e-6f + 1e-30f; tb += std::fabs(tb) * 1e-6f + 1e-30f;
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) return false;
    }
    tnear = t0; return true;
}

struct Hit { float t, u, v; uint32_t inst, prim; bool valid; };

// =====================================================================================================
// Textures — Vulkan texel addressing at LOD 0: sRGB decode before filtering, bilinear in fp32
// =====================================================================================================
inline int wrap_coord(int i, int n, uint32_t mode) {
    if (mode == RT_WRAP_REPEAT) { int m = i % n; return m < 0 ? m + n : m; }
    if (mode == RT_WRAP_MIRROR) {
        int m = i % (2 * n); if (m < 0) m += 2 * n;
        int k = m - n; k = k >= 0 ? k : -(1 + k);
        return (n - 1) - k;
    }
    return std::min(std::max(i, 0), n - 1);
}
inline vec4 texel(const orc_scene& s, const Image& im, int x, int y) {
    const uint8_t* p = &im.px[((size_t)y * im.w + x) * 4];
    if (im.srgb) return V4(s.srgb_lut[p[0]], s.srgb_lut[p[1]], s.srgb_lut[p[2]], p[3] / 255.0f);
    return V4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
}
vec4 sample_image(const orc_scene& s, const Image& im, uint32_t filter, uint32_t ws, uint32_t wt, vec2 uv) {
    if (!(std::isfinite(uv.x) && std::isfinite(uv.y)) || im.w == 0) return V4(0, 0, 0, 0);
    float u = uv.x * (float)im.w, v = uv.y * (float)im.h;
    if (filter == RT_FILTER_NEAREST) {
        int i = wrap_coord((int)std::floor(u), im.w, ws), j = wrap_coord((int)std::floor(v), im.h, wt);
        return texel(s, im, i, j);
    }
    float fu = u - 0.5f, fv = v - 0.5f;
    float i0f = std::floor(fu), j0f = std::floor(fv);
    float a = fu - i0f, b = fv - j0f;
    int i0 = wrap_coord((int)i0f, im.w, ws), i1 = wrap_coord((int)i0f + 1, im.w, ws);
    int j0 = wrap_coord((int)j0f, im.h, wt), j1 = wrap_coord((int)j0f + 1, im.h, wt);
    vec4 t00 = texel(s, im, i0, j0), t10 = texel(s, im, i1, j0), t01 = texel(s, im, i0, j1), t11 = texel(s, im, i1, j1);
    vec4 top = t00 * (1.0f - a) + t10 * a;
    vec4 bot = t01 * (1.0f - a) + t11 * a;
    return top * (1.0f - b) + bot * b;
}
vec4 texture2d(const orc_scene& s, int tex_index, vec2 uv) {
    if (tex_index < 0 || (size_t)tex_index >= s.textures.size()) return V4(1, 1, 1, 1);
    s.tex_taps.fetch_add(1, std::memory_order_relaxed);   // 16 B x 4 texels per tap
    const rt_texture_desc& t = s.textures[tex_index];
    const Image& im = s.images[t.image_index];
    const rt_sampler_desc& sm = s.samplers[t.sampler_index];
    return sample_image(s, im, sm.mag_filter, sm.wrap_s, sm.wrap_t, uv);
}
// samplerCube, Vulkan major-axis selection (z wins ties, then y); default sampler = linear filter.
// Simplification (documented): texels are clamped inside the selected face instead of filtering across faces.
vec3 texture_cube(const orc_scene& s, vec3 r) {
    float ax = std::fabs(r.x), ay = std::fabs(r.y), az = std::fabs(r.z);
    int face; float sc, tc, ma;
    if (az >= ax && az >= ay) { ma = az; if (r.z >= 0) { face = 4; sc = r.x; tc = -r.y; } else { face = 5; sc = -r.x; tc = -r.y; } }
    else if (ay >= ax)        { ma = ay; if (r.y >= 0) { face = 2; sc = r.x; tc = r.z; } else { face = 3; sc = r.x; tc = -r.z; } }
    else                      { ma = ax; if (r.x >= 0) { face = 0; sc = -r.z; tc = -r.y; } else { face = 1; sc = r.z; tc = -r.y; } }
    vec2 uv = V2(0.5f * (sc / ma) + 0.5f, 0.5f * (tc / ma) + 0.5f);
    vec4 c = sample_image(s, s.sky[face], RT_FILTER_LINEAR, RT_WRAP_CLAMP, RT_WRAP_CLAMP, uv);
    return xyz(c);
}

// =====================================================================================================
// lib/RayTracingCommons.glsl
// =====================================================================================================
inline vec2 getUV(vec4 uv0And1, int index) {   // :65-72
    if (index == 0) return V2(uv0And1.x, uv0And1.y);
    if (index == 1) return V2(uv0And1.z, uv0And1.w);
    return V2(0.0f, 0.0f);
}
struct MixVertex { vec4 uv0And1; vec3 pos; vec4 color; vec3 normal; vec4 tangent; };
inline vec3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
inline vec4 ld4(const float* p) { return V4(p[0], p[1], p[2], p[3]); }
// :75-97  getMixVertexAndGeoNormal
vec3 getMixVertexAndGeoNormal(const rt_vertex& a, const rt_vertex& b, const rt_vertex& c, vec2 attrs, MixVertex& m) {
    const vec3 bc = V3(1.0f - attrs.x - attrs.y, attrs.x, attrs.y);
    vec4 ua = V4(a.uv0[0], a.uv0[1], a.uv1[0], a.uv1[1]), ub = V4(b.uv0[0], b.uv0[1], b.uv1[0], b.uv1[1]),
         uc = V4(c.uv0[0], c.uv0[1], c.uv1[0], c.uv1[1]);
    m.uv0And1 = ua * bc.x + ub * bc.y + uc * bc.z;
    m.pos = ld3(a.position) * bc.x + ld3(b.position) * bc.y + ld3(c.position) * bc.z;
    m.color = ld4(a.color) * bc.x + ld4(b.color) * bc.y + ld4(c.color) * bc.z;
    m.normal = normalize(normalize(ld3(a.normal)) * bc.x + normalize(ld3(b.normal)) * bc.y + normalize(ld3(c.normal)) * bc.z);
    m.tangent = normalize(ld4(a.tangent) * bc.x + ld4(b.tangent) * bc.y + ld4(c.tangent) * bc.z);
    // calculate_geo_normal :75-80
    vec3 p0 = ld3(a.position), p1 = ld3(b.position), p2 = ld3(c.position);
    return cross(p1 - p0, p2 - p0);
}
// :103-119  offset_ray (Ray Tracing Gems ch. 6)
vec3 offset_ray(vec3 p, vec3 n) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    int32_t of_i[3] = {(int32_t)(n.x * int_scale), (int32_t)(n.y * int_scale), (int32_t)(n.z * int_scale)};
    vec3 p_i = V3(intBitsToFloat(floatBitsToInt(p.x) + ((p.x < 0) ? -of_i[0] : of_i[0])),
                  intBitsToFloat(floatBitsToInt(p.y) + ((p.y < 0) ? -of_i[1] : of_i[1])),
                  intBitsToFloat(floatBitsToInt(p.z) + ((p.z < 0) ? -of_i[2] : of_i[2])));
    return V3(std::fabs(p.x) < origin ? p.x + float_scale * n.x : p_i.x,
              std::fabs(p.y) < origin ? p.y + float_scale * n.y : p_i.y,
              std::fabs(p.z) < origin ? p.z + float_scale * n.z : p_i.z);
}

// =====================================================================================================
// lib/PBR.glsl (live functions only; preprocessor resolves to GGX + Frostbite + height-correlated G2)
// =====================================================================================================
const float PI = 3.141592653589f;
const float ONE_OVER_PI = 1.0f / PI;
const float TWO_PI = 2.0f * PI;
enum { DIFFUSE_TYPE = 1, SPECULAR_TYPE = 2, TRANSMISSION_TYPE = 3 };

struct MaterialBrdf {   // PBR.glsl:173-193
    vec3 baseColor; float metallic, roughness, ior, transmission; bool use_spec; float specular_factor;
    vec3 specular_color_factor, dielectricSpecularF0, dielectricSpecularF90, F0, F90, c_diff; bool frontFace;
    vec3 attenuation_color; float attenuation_distance; bool volume; float t_diff;
};
void matBuild(MaterialBrdf& m) {   // :195-206
    float factor = (m.ior - 1.0f) / (m.ior + 1.0f);
    m.dielectricSpecularF0 = min3(factor * factor * m.specular_color_factor, V3(1.0f)) * m.specular_factor;
    m.dielectricSpecularF90 = m.specular_color_factor;
    m.F0 = mix3(m.dielectricSpecularF0, m.baseColor, m.metallic);
    m.F90 = mix3(m.dielectricSpecularF90, V3(1.0f), m.metallic);
    m.c_diff = mix3(m.baseColor, V3(0.0f), m.metallic);
}
float luminance(vec3 rgb) { return dot(rgb, V3(0.2126f, 0.7152f, 0.0722f)); }   // :221-224
vec3 evalFresnelSchlick(vec3 f0, float f90, float NdotS) {   // :228-231
    return f0 + (f90 - f0) * std::pow(1.0f - NdotS, 5.0f);
}
struct BRDFp { float specular, diffuse, transmission; };
float Smith_G_a(float alpha, float NdotS) {   // :248-250
    return NdotS / (std::max(0.00001f, alpha) * std::sqrt(1.0f - std::min(0.99999f, NdotS * NdotS)));
}
float Smith_G_Lambda_GGX(float a) { return (-1.0f + std::sqrt(1.0f + (1.0f / (a * a)))) * 0.5f; }   // :253-255
float Smith_G2_Height_Correlated(float alpha, float NdotL, float NdotV) {   // :258-262
    float aL = Smith_G_a(alpha, NdotL), aV = Smith_G_a(alpha, NdotV);
    return 1.0f / (1.0f + Smith_G_Lambda_GGX(aL) + Smith_G_Lambda_GGX(aV));
}
float GGX_D(float alphaSquared, float NdotH) {   // :286-289
    float b = ((alphaSquared - 1.0f) * NdotH * NdotH + 1.0f);
    return alphaSquared / (PI * b * b);
}
float shadowedF90(vec3 F90) { return std::min(1.0f, luminance(F90)); }   // :312-322

struct BrdfData {   // :13-45
    vec3 specularF0, diffuseReflectance, specularF90; float roughness, alpha, alphaSquared; vec3 F;
    vec3 V, N, H, L; float NdotL, NdotV, LdotH, NdotH, VdotH; bool Vbackfacing, Lbackfacing;
};
vec3 evalMicrofacet(const BrdfData& d) {   // :292-303
    float D = GGX_D(std::max(0.00001f, d.alphaSquared), d.NdotH);
    float G2 = Smith_G2_Height_Correlated(d.alpha, d.NdotL, d.NdotV);
    return ((d.F * G2 * D) / (4.0f * d.NdotL * d.NdotV)) * d.NdotL;
}
BRDFp getBrdfProbability(const MaterialBrdf& mat, vec3 V, vec3 shadingNormal) {   // :324-358
    float specularF0 = luminance(mat.F0);
    float diffuseReflectance = luminance(mat.c_diff);
    float Fresnel = saturate(luminance(evalFresnelSchlick(V3(specularF0), shadowedF90(mat.F90), std::max(0.0f, dot(V, shadingNormal)))));
    float specular = Fresnel * mat.specular_factor;
    float penetration = diffuseReflectance * (1.0f - mat.specular_factor * Fresnel);
    float diffuse = penetration * (1.0f - mat.transmission);
    float transmission = penetration * mat.transmission;
    float sum = std::max(0.0001f, (specular + diffuse + transmission));
    float p = specular / sum;
    p = clampf(p, 0.001f, 0.9f);
    float d = (1 - p) * (1 - mat.transmission);
    float t = (1 - p) * mat.transmission;
    sum = p + d + t;
    p /= sum; d /= sum; t /= sum;
    return {p, d, t};
}
vec4 getRotationToZAxis(vec3 v) {   // :362-368
    if (v.z < -0.99999f) return V4(1.0f, 0.0f, 0.0f, 0.0f);
    return normalize(V4(v.y, -v.x, 0.0f, 1.0f + v.z));
}
vec3 rotatePoint(vec4 q, vec3 v) {   // :373-376
    const vec3 qAxis = V3(q.x, q.y, q.z);
    return 2.0f * dot(qAxis, v) * qAxis + (q.w * q.w - dot(qAxis, qAxis)) * v + 2.0f * q.w * cross(qAxis, v);
}
vec4 invertRotation(vec4 q) { return V4(-q.x, -q.y, -q.z, q.w); }   // :580-583
vec3 sampleHemisphere(vec2 u) {   // :381-399
    float a = std::sqrt(u.x), b = TWO_PI * u.y;
    return V3(a * std::cos(b), a * std::sin(b), std::sqrt(1.0f - u.x));
}
vec3 sampleGGXVNDF(vec3 Ve, vec2 alpha2D, vec2 u) {   // :413-436
    vec3 Vh = normalize(V3(alpha2D.x * Ve.x, alpha2D.y * Ve.y, Ve.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    vec3 T1 = lensq > 0.0f ? V3(-Vh.y, Vh.x, 0.0f) * inversesqrt(lensq) : V3(1.0f, 0.0f, 0.0f);
    vec3 T2 = cross(Vh, T1);
    float r = std::sqrt(u.x), phi = TWO_PI * u.y;
    float t1 = r * std::cos(phi), t2 = r * std::sin(phi);
    float s = 0.5f * (1.0f + Vh.z);
    t2 = mixf(std::sqrt(1.0f - t1 * t1), t2, s);
    vec3 Nh = t1 * T1 + t2 * T2 + std::sqrt(std::max(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(V3(alpha2D.x * Nh.x, alpha2D.y * Nh.y, std::max(0.0f, Nh.z)));
}
float Smith_G1_GGX4(float alpha, float NdotS, float alphaSquared, float NdotSSquared) {   // :446-448
    (void)alpha; (void)NdotS;
    return 2.0f / (std::sqrt(((alphaSquared * (1.0f - NdotSSquared)) + NdotSSquared) / NdotSSquared) + 1.0f);
}
float Smith_G2_Over_G1_Height_Correlated(float alpha, float alphaSquared, float NdotL, float NdotV) {   // :470-474
    float G1V = Smith_G1_GGX4(alpha, NdotV, alphaSquared, NdotV * NdotV);
    float G1L = Smith_G1_GGX4(alpha, NdotL, alphaSquared, NdotL * NdotL);
    return G1L / (G1V + G1L - G1V * G1L);
}
float frostbiteDisneyDiffuse(const BrdfData& d) {   // :486-496
    float energyBias = 0.5f * d.roughness;
    float energyFactor = mixf(1.0f, 1.0f / 1.51f, d.roughness);
    float FD90MinusOne = energyBias + 2.0f * d.LdotH * d.LdotH * d.roughness - 1.0f;
    float FDL = 1.0f + (FD90MinusOne * std::pow(1.0f - d.NdotL, 5.0f));
    float FDV = 1.0f + (FD90MinusOne * std::pow(1.0f - d.NdotV, 5.0f));
    return FDL * FDV * energyFactor;
}
vec3 evalFrostbiteDisneyDiffuse(const BrdfData& d) {   // :516-518
    return d.diffuseReflectance * (frostbiteDisneyDiffuse(d) * ONE_OVER_PI * d.NdotL);
}
BrdfData prepareBRDFData(vec3 N, vec3 L, vec3 V, const MaterialBrdf& material) {   // :527-564
    BrdfData d;
    d.V = V; d.N = N; d.H = normalize(L + V); d.L = L;
    float NdotL = dot(N, L), NdotV = dot(N, V);
    d.Vbackfacing = (NdotV <= 0.0f); d.Lbackfacing = (NdotL <= 0.0f);
    d.NdotL = std::min(std::max(0.00001f, NdotL), 1.0f);
    d.NdotV = std::min(std::max(0.00001f, NdotV), 1.0f);
    d.LdotH = saturate(dot(L, d.H)); d.NdotH = saturate(dot(N, d.H)); d.VdotH = saturate(dot(V, d.H));
    d.specularF0 = material.F0; d.specularF90 = material.F90; d.diffuseReflectance = material.c_diff;
    d.roughness = material.roughness; d.alpha = material.roughness * material.roughness; d.alphaSquared = d.alpha * d.alpha;
    d.F = evalFresnelSchlick(d.specularF0, shadowedF90(material.F90), d.VdotH);
    return d;
}
vec3 sampleSpecularMicrofacet(vec3 Vlocal, float alpha, float alphaSquared, vec3 specularF0, vec2 u, vec3& weight, vec3 specularF90) {   // :587-616
    vec3 Hlocal;
    if (alpha == 0.0f) Hlocal = V3(0.0f, 0.0f, 1.0f);
    else Hlocal = sampleGGXVNDF(Vlocal, V2(alpha, alpha), u);
    vec3 Llocal = reflect(-Vlocal, Hlocal);
    float HdotL = std::max(0.00001f, std::min(1.0f, dot(Hlocal, Llocal)));
    const vec3 Nlocal = V3(0.0f, 0.0f, 1.0f);
    float NdotL = std::max(0.00001f, std::min(1.0f, dot(Nlocal, Llocal)));
    float NdotV = std::max(0.00001f, std::min(1.0f, dot(Nlocal, Vlocal)));
    vec3 F = evalFresnelSchlick(specularF0, shadowedF90(specularF90), HdotL);
    weight = F * Smith_G2_Over_G1_Height_Correlated(alpha, alphaSquared, NdotL, NdotV);
    return Llocal;
}
// :649-729
bool evalIndirectCombinedBRDF(vec2 u, vec3 shadingNormal, vec3 geometryNormal, vec3 V, const MaterialBrdf& material,
                              uint32_t brdfType, vec3& rayDirection, vec3& sampleWeight, float& volume_dis) {
    if (dot(geometryNormal, V) < 0.0f) return false;
    vec4 qRotationToZ = getRotationToZAxis(shadingNormal);
    vec3 Vlocal = rotatePoint(qRotationToZ, V);
    const vec3 Nlocal = V3(0.0f, 0.0f, 1.0f);
    vec3 rayDirectionLocal = V3(0.0f);
    if (brdfType == DIFFUSE_TYPE) {
        rayDirectionLocal = sampleHemisphere(u);
        const BrdfData data = prepareBRDFData(Nlocal, rayDirectionLocal, Vlocal, material);
        sampleWeight = (1.0f - material.specular_factor * data.F) * data.diffuseReflectance * frostbiteDisneyDiffuse(data);
        sampleWeight *= (1.0f - material.transmission);
    } else if (brdfType == SPECULAR_TYPE) {
        const BrdfData data = prepareBRDFData(Nlocal, V3(0.0f, 0.0f, 1.0f), Vlocal, material);
        rayDirectionLocal = sampleSpecularMicrofacet(Vlocal, data.alpha, data.alphaSquared, data.specularF0, u, sampleWeight, data.specularF90);
        sampleWeight *= material.specular_factor;
    } else if (brdfType == TRANSMISSION_TYPE) {
        if (material.volume) {
            const float refraction_ratio = material.frontFace ? 1.0f / material.ior : material.ior;
            const vec3 refracted = refract(-Vlocal, Nlocal, refraction_ratio);
            if (refracted == V3(0.0f)) { sampleWeight = V3(0.0f); return false; }
            rayDirectionLocal = refracted;
        } else {
            rayDirectionLocal = -Vlocal;
        }
        const BrdfData data = prepareBRDFData(Nlocal, rayDirectionLocal, Vlocal, material);
        sampleWeight = max3(V3(0.0f), data.diffuseReflectance * material.transmission);
        if (!material.frontFace && material.volume) {
            float dis = volume_dis;
            volume_dis = -1.0f;
            vec3 sigma = log3(material.attenuation_color) / material.attenuation_distance;
            vec3 attenuation = exp3(sigma * dis);
            sampleWeight *= min3(attenuation, V3(1.0f));
        }
    }
    if (luminance(sampleWeight) == 0.0f) return false;
    rayDirection = normalize(rotatePoint(invertRotation(qRotationToZ), rayDirectionLocal));
    return true;
}
vec3 evalCombinedBRDF(vec3 N, vec3 L, vec3 V, const MaterialBrdf& material) {   // :731-745
    const BrdfData data = prepareBRDFData(N, L, V, material);
    if (data.Vbackfacing || data.Lbackfacing) return V3(0.0f);
    vec3 specular = evalMicrofacet(data);
    vec3 diffuse = evalFrostbiteDisneyDiffuse(data);
    return diffuse + specular;
}
float Schlick(float cosine, float refractionIndex) {   // :48-53 (debug path only)
    float r0 = (1 - refractionIndex) / (1 + refractionIndex);
    r0 *= r0;
    return r0 + (1 - r0) * std::pow(1 - cosine, 5.0f);
}

// lib/Material.glsl:78-89
float convertMetallic(vec3 diffuse, vec3 specular, float maxSpecular) {
    const float c_MinRoughness = 0.04f;
    float perceivedDiffuse = std::sqrt(0.299f * diffuse.x * diffuse.x + 0.587f * diffuse.y * diffuse.y + 0.114f * diffuse.z * diffuse.z);
    float perceivedSpecular = std::sqrt(0.299f * specular.x * specular.x + 0.587f * specular.y * specular.y + 0.114f * specular.z * specular.z);
    if (perceivedSpecular < c_MinRoughness) return 0.0f;
    float a = c_MinRoughness;
    float b = perceivedDiffuse * (1.0f - maxSpecular) / (1.0f - c_MinRoughness) + perceivedSpecular - 2.0f * c_MinRoughness;
    float c = c_MinRoughness - perceivedSpecular;
    float D = std::max(b * b - 4.0f * a * c, 0.0f);
    return clampf((-b + std::sqrt(D)) / (2.0f * a), 0.0f, 1.0f);
}

// lib/PunctualLight.glsl:17-30
vec3 getLightIntensityAtPoint(const rt_light& light, float distance) {
    vec3 color = light.intensity * ld3(light.color);
    if (light.kind == 1) {
        const float radius = 0.5f, radiusSquared = radius * radius, distanceSquared = distance * distance;
        const float attenuation = 2.0f / (distanceSquared + radiusSquared + distance * std::sqrt(distanceSquared + radiusSquared));
        return color * attenuation;
    }
    return color;
}

// =====================================================================================================
// Any-hit — RayTracing.rahit:44-104 (and the identical RayTracing.shadow.rahit)
// returns true when the candidate must be ignored
// =====================================================================================================
bool testOpacityAnyHit(const orc_scene& s, uint32_t instance_id, uint32_t primitive_id, uint32_t geo_id, vec2 attrs, uvec4 rngState) {
    const rt_prim_info& pi = s.prim_infos[geo_id];
    const rt_material& mat = s.materials[pi.material_id];
    if (mat.alpha_mode == 1) return false;
    const uint32_t io = pi.i_offset + 3 * primitive_id;
    const rt_vertex& v0 = s.vertices[pi.v_offset + s.indices[io]];
    const rt_vertex& v1 = s.vertices[pi.v_offset + s.indices[io + 1]];
    const rt_vertex& v2 = s.vertices[pi.v_offset + s.indices[io + 2]];
    MixVertex mv;
    getMixVertexAndGeoNormal(v0, v1, v2, attrs, mv);
    vec4 color4 = mv.color * ld4(mat.base_color);
    if (mat.base_color_texture.index >= 0)
        color4 *= texture2d(s, mat.base_color_texture.index, getUV(mv.uv0And1, mat.base_color_texture.coord));
    float opacity = color4.w;
    if (mat.workflow == 1) {
        vec4 diffuse_factor = ld4(mat.sg_diffuse_factor);
        if (mat.sg_diffuse_texture.index >= 0)
            diffuse_factor *= texture2d(s, mat.sg_diffuse_texture.index, getUV(mv.uv0And1, mat.sg_diffuse_texture.coord));
        color4 = mv.color * diffuse_factor;
        opacity = color4.w;
    }
    if (mat.alpha_mode == 2) return opacity < mat.alpha_cutoff;
    float u = blendRandom(rngState, instance_id, primitive_id);   // D3 (reference: rand(Ray.rngState))
    return opacity <= u;
}

// =====================================================================================================
// traceRayEXT — two-level traversal.  Closest hit = lexicographic minimum of (t, instance, primitive)
// over all accepted candidates with tmin < t < tmax, hence independent of traversal order.
// =====================================================================================================
struct TraceCtx { bool ray_opaque; bool terminate_first; uvec4 rng; };

void trace_blas(const orc_scene& s, uint32_t inst_id, vec3 o, vec3 d, float tmin, float tmax, const TraceCtx& c, Hit& best, bool& done) {
    const Instance& in = s.instances[inst_id];
    const Geometry& g = s.geos[in.geo_id];
    const Bvh& bvh = in.baked ? in.wbvh : g.bvh;
    if (bvh.nodes.empty()) return;
    vec3 oo = in.baked ? o : xform_point(in.w2o, o), od = in.baked ? d : xform_dir(in.w2o, d);
    RayShear rs; rs.init(oo, od);
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    const bool need_alpha = !g.opaque && !c.ray_opaque;
    while (sp) {
        const BNode& n = bvh.nodes[stack[--sp]];
        float tn;
        if (!box_test(n.box, oo, od, tmin, best.valid ? best.t : tmax, tn)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; ++i) {
                uint32_t prim = bvh.prim[n.left + i];
                const uint32_t io = g.i_offset + 3 * prim;
                vec3 p0, p1, p2;
                if (in.baked) { p0 = in.wpos[3 * prim]; p1 = in.wpos[3 * prim + 1]; p2 = in.wpos[3 * prim + 2]; }
                else {
                    p0 = ld3(s.vertices[g.v_offset + s.indices[io]].position);
                    p1 = ld3(s.vertices[g.v_offset + s.indices[io + 1]].position);
                    p2 = ld3(s.vertices[g.v_offset + s.indices[io + 2]].position);
                }
                float t, u, v;
                if (!tri_test(rs, p0, p1, p2, tmin, tmax, t, u, v)) continue;
                if (best.valid) {
                    if (t > best.t) continue;
                    if (t == best.t && !(inst_id < best.inst || (inst_id == best.inst && prim < best.prim))) continue;
                }
                if (need_alpha && testOpacityAnyHit(s, inst_id, prim, in.geo_id, V2(u, v), c.rng)) continue;
                best = {t, u, v, inst_id, prim, true};
                if (c.terminate_first) { done = true; return; }
            }
        } else {
            if (sp + 2 > 128) { std::fprintf(stderr, "oracle: BLAS stack overflow\n"); std::abort(); }
            stack[sp++] = n.left; stack[sp++] = n.left + 1;
        }
    }
}

Hit trace(const orc_scene& s, vec3 o, vec3 d, float tmin, float tmax, const TraceCtx& c) {
    Hit best{-1.0f, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, false};
    if (s.tlas.nodes.empty()) return best;
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    bool done = false;
    while (sp && !done) {
        const BNode& n = s.tlas.nodes[stack[--sp]];
        float tn;
        if (!box_test(n.box, o, d, tmin, best.valid ? best.t : tmax, tn)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count && !done; ++i)
                trace_blas(s, s.tlas.prim[n.left + i], o, d, tmin, tmax, c, best, done);
        } else {
            if (sp + 2 > 128) { std::fprintf(stderr, "oracle: TLAS stack overflow\n"); std::abort(); }
            stack[sp++] = n.left; stack[sp++] = n.left + 1;
        }
    }
    return best;
}

// brute force over every instance and triangle (used by tests to pin the BVH traversal)
Hit trace_brute(const orc_scene& s, vec3 o, vec3 d, float tmin, float tmax, const TraceCtx& c) {
    Hit best{-1.0f, 0, 0, 0xFFFFFFFFu, 0xFFFFFFFFu, false};
    for (uint32_t ii = 0; ii < s.instances.size(); ++ii) {
        const Instance& in = s.instances[ii];
        const Geometry& g = s.geos[in.geo_id];
        vec3 oo = in.baked ? o : xform_point(in.w2o, o), od = in.baked ? d : xform_dir(in.w2o, d);
        RayShear rs; rs.init(oo, od);
        const bool need_alpha = !g.opaque && !c.ray_opaque;
        for (uint32_t prim = 0; prim < g.i_len / 3; ++prim) {
            const uint32_t io = g.i_offset + 3 * prim;
            vec3 p0, p1, p2;
            if (in.baked) { p0 = in.wpos[3 * prim]; p1 = in.wpos[3 * prim + 1]; p2 = in.wpos[3 * prim + 2]; }
            else {
                p0 = ld3(s.vertices[g.v_offset + s.indices[io]].position);
                p1 = ld3(s.vertices[g.v_offset + s.indices[io + 1]].position);
                p2 = ld3(s.vertices[g.v_offset + s.indices[io + 2]].position);
            }
            float t, u, v;
            if (!tri_test(rs, p0, p1, p2, tmin, tmax, t, u, v)) continue;
            if (best.valid && (t > best.t || (t == best.t && !(ii < best.inst || (ii == best.inst && prim < best.prim))))) continue;
            if (need_alpha && testOpacityAnyHit(s, ii, prim, in.geo_id, V2(u, v), c.rng)) continue;
            best = {t, u, v, ii, prim, true};
            if (c.terminate_first) return best;
        }
    }
    return best;
}

// =====================================================================================================
// Payload + shaders
// =====================================================================================================
struct RayPayload {   // RayTracingCommons.glsl:11-24
    vec3 hitValue{0, 0, 0}; vec3 hitPoint{0, 0, 0}; float t = 0; vec3 scatterDirection{0, 0, 0}; bool needScatter = false;
    uint32_t RandomSeed = 0; vec3 emittance{0, 0, 0}; uvec4 rngState{0, 0, 0, 0}; float volume_dis = -1.0f;
};

const float tMin = 0.001f, tMax = 10000.0f;   // lib/Camera.glsl:2-3

// RayTracing.rmiss:14-45
void miss_shader(const orc_scene& s, const rt_ubo& ubo, vec3 worldRayDirection, RayPayload& Ray) {
    vec3 light_acc = V3(0.0f);
    vec3 ray_direction = normalize(worldRayDirection);
    if (Ray.t != 0) {
        for (size_t i = 0; i < s.dlights.size(); i++) {
            const rt_light& li = s.dlights[i];
            float c = dot(normalize(ld3(li.transform)), ray_direction);
            if (c < 0.0f) light_acc += -c * ld3(li.color) * li.intensity;
        }
    }
    if (ubo.has_sky) {
        vec3 skyColor = s.has_sky_faces ? texture_cube(s, ray_direction) : V3(0.0f);
        light_acc += skyColor + light_acc;
    } else {
        light_acc += V3(0.01f);
    }
    Ray.hitValue = V3(0.0f);
    Ray.needScatter = false;
    Ray.emittance = light_acc;
    if (s.dlights.empty()) Ray.emittance = V3(0.0f);
    Ray.t = -1.0f;
}

// RayTracing.rchit:35-59
bool castShadowRay(orc_scene& s, const rt_ubo& ubo, vec3 hitPosition, vec3 directionToLight, float tmax, const RayPayload& Ray) {
    TraceCtx c; c.ray_opaque = ubo.fully_opaque != 0; c.terminate_first = true; c.rng = Ray.rngState;
    s.rays_shadow.fetch_add(1, std::memory_order_relaxed);
    if (s.rec && s.rec->shadow.size() < s.rec->max_shadow) {
        rt_ray r = {{hitPosition.x, hitPosition.y, hitPosition.z}, 0.1f, {directionToLight.x, directionToLight.y, directionToLight.z}, tmax};
        s.rec->shadow.push_back(r);
        for (uint32_t v : {c.rng.x, c.rng.y, c.rng.z, c.rng.w}) s.rec->shadow_rng.push_back(v);
    }
    Hit h = trace(s, hitPosition, directionToLight, 0.1f, tmax, c);
    return !h.valid;
}

// RayTracing.rchit:77-122 (sampleLightUniform :62-75 inlined)
bool sampleLightRIS(const orc_scene& s, uvec4& rngState, vec3 hitPosition, vec3 surfaceNormal, rt_light& selectedSample, float& lightSampleWeight) {
    uint32_t light_num = (uint32_t)s.plights.size();
    if (light_num == 0) return false;
    float totalWeights = 0.0f, samplePdfG = 0.0f;
    uint32_t candidates_num = std::min(light_num, 3u);
    for (uint32_t i = 0; i < candidates_num; i++) {
        if (luminance(ld3(s.plights[i].color) * s.plights[i].intensity) < 0.1f) continue;
        s.light_cands.fetch_add(1, std::memory_order_relaxed);
        uint32_t randomLightIndex = std::min(light_num - 1, (uint32_t)(rnd(rngState) * (float)light_num));
        const rt_light& candidate = s.plights[randomLightIndex];
        float candidateWeight = (float)light_num;
        vec3 lightVector = ld3(candidate.transform) - hitPosition;
        vec3 L = normalize(lightVector);
        if (dot(surfaceNormal, L) < 0.00001f) continue;
        float candidatePdfG = luminance(getLightIntensityAtPoint(candidate, length(lightVector)));
        const float candidateRISWeight = candidatePdfG * candidateWeight;
        totalWeights += candidateRISWeight;
        if (rnd(rngState) < (candidateRISWeight / totalWeights)) { selectedSample = candidate; samplePdfG = candidatePdfG; }
    }
    if (totalWeights == 0.0f) return false;
    lightSampleWeight = (totalWeights / 3.0f) / samplePdfG;
    return true;
}

// RayTracing.rchit:136-477
void closest_hit(orc_scene& s, const rt_ubo& ubo, const Hit& hit, vec3 worldRayDirection, RayPayload& Ray) {
    const Instance& inst = s.instances[hit.inst];
    const uint32_t customIndex = inst.geo_id;
    const rt_prim_info& primInfo = s.prim_infos[customIndex];
    const rt_material& mat = s.materials[primInfo.material_id];
    const float* O2W = inst.o2w;
    auto normal_transform = [&](vec3 n) { return normalize(xform_dir(O2W, n)); };   // :28-30

    vec3 last_hit = Ray.hitPoint;
    Ray.t = hit.t;
    const uint32_t indexOffset = primInfo.i_offset + 3 * hit.prim;
    const rt_vertex& v0 = s.vertices[primInfo.v_offset + s.indices[indexOffset]];
    const rt_vertex& v1 = s.vertices[primInfo.v_offset + s.indices[indexOffset + 1]];
    const rt_vertex& v2 = s.vertices[primInfo.v_offset + s.indices[indexOffset + 2]];

    MixVertex mix_vertex;
    const vec2 HitAttributes = V2(hit.u, hit.v);
    vec3 geo_normal = normal_transform(getMixVertexAndGeoNormal(v0, v1, v2, HitAttributes, mix_vertex));
    vec3 pos = mix_vertex.pos;
    vec3 origin = xform_point(O2W, pos);
    vec4 uv0And1 = mix_vertex.uv0And1;

    vec4 color4 = mix_vertex.color * ld4(mat.base_color);
    if (mat.base_color_texture.index >= 0)
        color4 *= texture2d(s, mat.base_color_texture.index, getUV(uv0And1, mat.base_color_texture.coord));
    vec3 color = xyz(color4);
    Ray.needScatter = false;
    Ray.hitPoint = pos;

    vec3 normal = mix_vertex.normal;
    if (mat.normal_texture.index >= 0) {
        vec3 normal_t = normalize(xyz(texture2d(s, mat.normal_texture.index, getUV(uv0And1, mat.normal_texture.coord))) * 2.0f - 1.0f);
        // getNormal :124-129
        vec3 tm = xyz(mix_vertex.tangent);
        vec3 tangent = normalize(tm - dot(tm, normal) * normal);
        vec3 b = normalize(cross(normal, tangent) * mix_vertex.tangent.w);
        normal = tangent * normal_t.x + b * normal_t.y + normal * normal_t.z;
    }
    normal = normal_transform(normal);

    const vec3 V = -normalize(worldRayDirection);
    const float cosv = dot(V, geo_normal);
    const bool frontFace = cosv >= 0.0f;
    geo_normal = frontFace ? geo_normal : -geo_normal;
    const vec3 outwardNormal = dot(geo_normal, normal) < 0.0f ? -normal : normal;

    vec3 emittance = ld3(mat.emissive_factor);
    if (mat.emissive_texture.index >= 0)
        emittance *= xyz(texture2d(s, mat.emissive_texture.index, getUV(uv0And1, mat.emissive_texture.coord)));

    float metallic = mat.metallic_factor, roughness = mat.roughness_factor;
    if (mat.metallic_roughness_texture.index >= 0) {
        vec4 mr = texture2d(s, mat.metallic_roughness_texture.index, getUV(uv0And1, mat.metallic_roughness_texture.coord));
        roughness *= mr.y; metallic *= mr.z;
    }

    vec3 specular_factor_workflow = V3(1.0f);
    if (mat.workflow == 1) {
        vec4 diffuse_factor = ld4(mat.sg_diffuse_factor);
        vec4 specular_glossiness_factor = ld4(mat.sg_specular_glossiness_factor);
        if (mat.sg_diffuse_texture.index >= 0)
            diffuse_factor *= texture2d(s, mat.sg_diffuse_texture.index, getUV(uv0And1, mat.sg_diffuse_texture.coord));
        if (mat.sg_specular_glossiness_texture.index >= 0)
            specular_glossiness_factor *= texture2d(s, mat.sg_specular_glossiness_texture.index, getUV(uv0And1, mat.sg_specular_glossiness_texture.coord));
        specular_factor_workflow = xyz(specular_glossiness_factor);
        roughness = 1.0f - specular_glossiness_factor.w;
        float maxSpecular = std::max(std::max(specular_factor_workflow.x, specular_factor_workflow.y), specular_factor_workflow.z);
        color = xyz(mix_vertex.color) * xyz(diffuse_factor);
        metallic = convertMetallic(color, specular_factor_workflow, maxSpecular);
    }

    float transmission_factor = 0.0f;
    if (mat.transmission_exist) {
        transmission_factor = mat.transmission_factor;
        if (mat.transmission_texture.index >= 0)
            transmission_factor *= texture2d(s, mat.transmission_texture.index, getUV(uv0And1, mat.transmission_texture.coord)).x;
    }

    uint32_t mapping = ubo.mapping;
    if (mat.unlit) mapping = RT_MAP_ALBEDO;
    switch (mapping) {   // :258-286
        case RT_MAP_ALBEDO: Ray.emittance = color; return;
        case RT_MAP_TRIANGLE: Ray.emittance = V3(1 - HitAttributes.x - HitAttributes.y, HitAttributes.x, HitAttributes.y); return;
        case RT_MAP_INSTANCE: Ray.emittance = hashAndColor(hit.inst); return;
        case RT_MAP_METALLIC: Ray.emittance = V3(metallic); return;
        case RT_MAP_ROUGHNESS: Ray.emittance = V3(roughness); return;
        case RT_MAP_NORMAL: Ray.emittance = (outwardNormal + 1.0f) / 2.0f; return;
        case RT_MAP_TANGENT: Ray.emittance = (normal_transform(xyz(mix_vertex.tangent)) + 1.0f) / 2.0f; return;
        case RT_MAP_TRANSMISSION: Ray.emittance = V3(transmission_factor); return;
        case RT_MAP_GEO_ID: Ray.emittance = hashAndColor(customIndex); return;
        default: break;
    }

    float spec_factor = mat.specular_factor;
    vec3 spec_color_factor = ld3(mat.specular_color_factor);
    if (mat.specular_texture.index >= 0)
        spec_factor *= texture2d(s, mat.specular_texture.index, getUV(uv0And1, mat.specular_texture.coord)).w;
    if (mat.specular_color_texture.index >= 0)
        spec_color_factor *= xyz(texture2d(s, mat.specular_color_texture.index, getUV(uv0And1, mat.specular_color_texture.coord)));
    const float ior = mat.ior;

    Ray.hitPoint = origin;
    uint32_t seed = Ray.RandomSeed;
    uvec4 rngState = Ray.rngState;

    Ray.emittance = emittance * ubo.exposure;
    Ray.needScatter = false;
    uint32_t brdfType;

    vec3 throughput = V3(1.0f);
    MaterialBrdf matbrdf;
    matbrdf.baseColor = color; matbrdf.metallic = metallic; matbrdf.roughness = roughness;
    matbrdf.ior = mat.volume_exists ? ior : 1.0f;   // :320
    matbrdf.transmission = transmission_factor;
    matbrdf.specular_factor = spec_factor; matbrdf.specular_color_factor = spec_color_factor;
    matbrdf.use_spec = mat.specular_exist != 0; matbrdf.frontFace = frontFace;
    matBuild(matbrdf);
    if (mat.workflow == 1) {
        float maxSpecular = std::max(std::max(specular_factor_workflow.x, specular_factor_workflow.y), specular_factor_workflow.z);
        matbrdf.c_diff = color * (1.0f - maxSpecular);
        matbrdf.F0 = specular_factor_workflow;
    }
    matbrdf.attenuation_color = ld3(mat.attenuation_color);
    matbrdf.attenuation_distance = mat.attenuation_distance;
    matbrdf.volume = mat.volume_exists != 0;
    float displacement = length(origin - last_hit);
    matbrdf.t_diff = displacement;

    rt_light light; float light_weight;
    if (sampleLightRIS(s, rngState, origin, geo_normal, light, light_weight)) {
        vec3 light_vec = ld3(light.transform) - origin;
        float light_distance = length(light_vec);
        light_vec = normalize(light_vec);
        if (castShadowRay(s, ubo, origin, light_vec, light_distance, Ray)) {
            Ray.emittance += evalCombinedBRDF(outwardNormal, light_vec, V, matbrdf) * light_weight * light.intensity * ld3(light.color);
        }
    }

    if (metallic == 1.0f && roughness == 0.0f) {
        brdfType = SPECULAR_TYPE;
    } else {
        BRDFp bp = getBrdfProbability(matbrdf, V, outwardNormal);
        float randfloat = rnd(rngState);
        if (randfloat < bp.specular) {
            brdfType = SPECULAR_TYPE;
            throughput /= bp.specular;
            if (Ray.volume_dis >= 0) Ray.volume_dis += displacement;
        } else if (randfloat >= bp.specular && randfloat <= bp.specular + bp.diffuse) {
            brdfType = DIFFUSE_TYPE;
            throughput /= bp.diffuse;
            if (Ray.volume_dis >= 0) Ray.volume_dis += displacement;
        } else {
            brdfType = TRANSMISSION_TYPE;
            if (mat.volume_exists) {
                if (Ray.volume_dis >= 0) Ray.volume_dis += displacement;
                else if (frontFace) Ray.volume_dis = 0;
            }
            throughput /= bp.transmission;
        }
    }

    if (brdfType == TRANSMISSION_TYPE) origin = offset_ray(origin, -geo_normal);
    vec3 brdfWeight = V3(0.0f);   // D8
    float u0 = rnd(rngState), u1 = rnd(rngState);
    vec2 u = V2(u0, u1);
    vec3 direction = V3(0.0f);    // D8
    Ray.needScatter = evalIndirectCombinedBRDF(u, outwardNormal, geo_normal, V, matbrdf, brdfType, direction, brdfWeight, Ray.volume_dis);

    throughput *= brdfWeight;
    Ray.hitPoint = origin;
    Ray.scatterDirection = direction;
    Ray.hitValue = throughput;

    if (ubo.debug == 1) {   // :437-474 legacy path on the LCG stream
        if (matbrdf.transmission > 0.0f) {
            const float refraction_ratio = frontFace ? 1 / ior : ior;
            const float cos_theta = std::fabs(cosv);
            const vec3 refracted = refract(worldRayDirection, outwardNormal, refraction_ratio);
            const float reflectProb = refracted != V3(0.0f) ? Schlick(cos_theta, refraction_ratio) : 1;
            Ray.hitValue = color;
            Ray.needScatter = true;
            if (RandomFloat(seed) < reflectProb) Ray.scatterDirection = reflect(worldRayDirection, normal);
            else Ray.scatterDirection = refracted;
        } else if (length(emittance) < 0.01f && roughness == 1.0f) {
            const bool isScattered = dot(worldRayDirection, geo_normal) < 0.0f;
            const vec3 scatter = normalize(outwardNormal + RandomInUnitSphere(seed));
            Ray.needScatter = isScattered;
            Ray.scatterDirection = scatter;
            Ray.hitValue = isScattered ? color : V3(0.0f);
        } else if (metallic > 0.0f) {
            vec3 reflected = reflect(worldRayDirection, outwardNormal);
            const bool isScattered = dot(reflected, geo_normal) > 0;
            Ray.needScatter = isScattered;
            Ray.hitValue = isScattered ? color : V3(0.0f);
            Ray.scatterDirection = reflected + 0.08f * RandomInUnitSphere(seed);
        }
    }
    Ray.RandomSeed = seed;
    Ray.rngState = rngState;
}

// lib/Tonemapping.glsl
vec3 LINEARtoSRGB(vec3 c) { return pow3(c, 1.0f / 2.2f); }
vec3 toneMapUncharted2Impl(vec3 color) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((color * (A * color + C * B) + D * E) / (color * (A * color + B) + D * F)) - E / F;
}
vec3 toneMapUncharted(vec3 color) {
    const float W = 11.2f;
    color = toneMapUncharted2Impl(color * 2.0f);
    vec3 whiteScale = 1.0f / toneMapUncharted2Impl(V3(W));
    return LINEARtoSRGB(color * whiteScale);
}
vec3 toneMapHejlRichard(vec3 color) {
    color = max3(V3(0.0f), color - V3(0.004f));
    return (color * (6.2f * color + 0.5f)) / (color * (6.2f * color + 1.7f) + 0.06f);
}
vec3 toneMapACES(vec3 color) {
    const float A = 2.51f, B = 0.03f, C = 2.43f, D = 0.59f, E = 0.14f;
    return LINEARtoSRGB(clamp3((color * (A * color + B)) / (color * (C * color + D) + E), 0.0f, 1.0f));
}
vec3 defaultToneMap(vec3 color) { color = color / (color + 1.0f); return LINEARtoSRGB(color); }
// lib/Heatmap.glsl:4-33
vec3 heatmap(float t) {
    const vec3 c[10] = {V3(0.0f / 255.0f, 2.0f / 255.0f, 91.0f / 255.0f), V3(0.0f / 255.0f, 108.0f / 255.0f, 251.0f / 255.0f),
                        V3(0.0f / 255.0f, 221.0f / 255.0f, 221.0f / 255.0f), V3(51.0f / 255.0f, 221.0f / 255.0f, 0.0f / 255.0f),
                        V3(255.0f / 255.0f, 252.0f / 255.0f, 0.0f / 255.0f), V3(255.0f / 255.0f, 180.0f / 255.0f, 0.0f / 255.0f),
                        V3(255.0f / 255.0f, 104.0f / 255.0f, 0.0f / 255.0f), V3(226.0f / 255.0f, 22.0f / 255.0f, 0.0f / 255.0f),
                        V3(191.0f / 255.0f, 0.0f / 255.0f, 83.0f / 255.0f), V3(145.0f / 255.0f, 0.0f / 255.0f, 65.0f / 255.0f)};
    const float s = t * 10.0f;
    const int cur = int(s) <= 9 ? int(s) : 9;
    const int prv = cur >= 1 ? cur - 1 : 0;
    const int nxt = cur < 9 ? cur + 1 : 9;
    const float blur = 0.8f;
    const float wc = smoothstep(float(cur) - blur, float(cur) + blur, s) * (1.0f - smoothstep(float(cur + 1) - blur, float(cur + 1) + blur, s));
    const float wp = 1.0f - smoothstep(float(cur) - blur, float(cur) + blur, s);
    const float wn = smoothstep(float(cur + 1) - blur, float(cur + 1) + blur, s);
    const vec3 r = wc * c[cur] + wp * c[prv] + wn * c[nxt];
    return clamp3(r, 0.0f, 1.0f);
}

inline uint8_t to_unorm8(float v) {   // Vulkan float -> UNORM: clamp, scale, round to nearest even
    if (!(v > 0.0f)) return 0;         // also NaN -> 0
    if (v >= 1.0f) return 255;
    return (uint8_t)std::nearbyintf(v * 255.0f);
}

// RayTracing.rgen:25-167 for one pixel.  acc (RGBA32F) is read/written, out (RGBA8) written.
void raygen(orc_scene& s, const rt_ubo& ubo, uint32_t px, uint32_t py, uint32_t W, uint32_t H, float* acc, uint8_t* out) {
    const mat4& MVI = *reinterpret_cast<const mat4*>(ubo.model_view_inverse);
    const mat4& PI_ = *reinterpret_cast<const mat4*>(ubo.projection_inverse);
    // D1: clockARB() surrogate
    const uint32_t clk = InitRandomSeed(ubo.total_number_of_samples, ubo.random_seed);
    RayPayload Ray;
    Ray.RandomSeed = InitRandomSeed(InitRandomSeed(px, py), clk);
    uvec4 rngState = {px, py, ubo.frame_count, 0};
    Ray.rngState = {px, py, clk, 0};
    uint64_t n_traces = 0;

    vec3 radiance = V3(0.0f);
    for (uint32_t smp = 0; smp < ubo.number_of_samples; ++smp) {
        vec2 pixel;
        vec2 pixelCenter = V2((float)px + 0.5f, (float)py + 0.5f);
        if (ubo.antialiasing) {
            float ox = rnd(rngState), oy = rnd(rngState);
            pixel = pixelCenter + (V2(ox, oy) - 0.5f);
        } else pixel = pixelCenter;
        const vec2 uv = (pixel / V2((float)W, (float)H)) * 2.0f - 1.0f;

        vec2 offset = ubo.aperture / 2 * RandomInUnitDisk(Ray.RandomSeed);
        vec4 origin = mul(MVI, V4(offset.x, offset.y, 0, 1));
        vec4 target = mul(PI_, V4(uv.x, uv.y, 1, 1));
        vec4 direction = mul(MVI, V4(normalize(xyz(target) * ubo.focus_distance - V3(offset.x, offset.y, 0)), 0));
        float tFar = tMax;
        if (ubo.orthographic_fov_dis > 0.0f) {
            vec2 new_uv = (1.0f + ubo.orthographic_fov_dis) * uv;
            origin = mul(MVI, V4(new_uv.x, -new_uv.y, 0, 1));
            direction = mul(MVI, V4(0, 0, -1, 0));
            tFar = 10 * tMax;
        }
        vec3 throughput = V3(1.0f);
        Ray.t = 0;
        Ray.volume_dis = -1.0f;
        for (uint32_t b = 0; b < ubo.number_of_bounces; b++) {
            TraceCtx c; c.ray_opaque = ubo.fully_opaque != 0; c.terminate_first = false; c.rng = Ray.rngState;
            if (s.rec && b >= s.rec->min_bounce && s.rec->rays.size() < s.rec->max_rays) {
                rt_ray r = {{origin.x, origin.y, origin.z}, tMin, {direction.x, direction.y, direction.z}, tFar};
                s.rec->rays.push_back(r);
                for (uint32_t v : {c.rng.x, c.rng.y, c.rng.z, c.rng.w}) s.rec->rng.push_back(v);
            }
            Hit h = trace(s, xyz(origin), xyz(direction), tMin, tFar, c);
            s.rays_extend.fetch_add(1, std::memory_order_relaxed); n_traces++;
            if (h.valid) { s.shaded.fetch_add(1, std::memory_order_relaxed); closest_hit(s, ubo, h, xyz(direction), Ray); }
            else miss_shader(s, ubo, xyz(direction), Ray);

            const vec3 hitColor = Ray.hitValue;
            const float t = Ray.t;
            const bool isScattered = Ray.needScatter;
            radiance += throughput * Ray.emittance;
            Ray.emittance = V3(0.0f);
            if (b + 1 == ubo.number_of_bounces) break;
            if (b > 3) {   // MIN_BOUNCES 3
                float rrProbability = clampf(luminance(throughput), 0.01f, 0.95f);
                float prop = rnd(rngState);
                if (rrProbability < prop) break;
                else throughput /= rrProbability;
            }
            throughput *= hitColor;
            if (!isScattered || t < 0) break;
            origin = V4(Ray.hitPoint, 1.0f);
            direction = V4(Ray.scatterDirection, 0);
        }
    }

    const size_t pi = ((size_t)py * W + px) * 4;
    const bool accumulate = ubo.number_of_samples != ubo.total_number_of_samples;
    const vec3 accumulatedColor = (accumulate ? V3(acc[pi], acc[pi + 1], acc[pi + 2]) : V3(0.0f)) + radiance;
    radiance = accumulatedColor / (float)ubo.total_number_of_samples;
    vec3 color;
    switch (ubo.tone_mapping_mode) {
        case 0: color = defaultToneMap(radiance); break;
        case 1: color = toneMapUncharted(radiance); break;
        case 2: color = toneMapHejlRichard(radiance); break;
        case 3: color = toneMapACES(radiance); break;
        default: color = LINEARtoSRGB(radiance); break;
    }
    if (ubo.mapping == RT_MAP_HEAT) {
        // D9: clockARB() delta replaced by 100000 ticks per traced path segment (deterministic)
        const float heatmapScale = 1000000.0f * ubo.heatmap_scale * ubo.heatmap_scale;
        const float deltaTimeScaled = clampf((float)(n_traces * 100000ull) / heatmapScale, 0.0f, 1.0f);
        color = heatmap(deltaTimeScaled);
    } else if (ubo.mapping == RT_MAP_DISTANCE) {
        color = V3(std::min((ubo.heatmap_scale - std::max(Ray.t, tMin)) / ubo.heatmap_scale, 1.0f));
    }
    acc[pi] = accumulatedColor.x; acc[pi + 1] = accumulatedColor.y; acc[pi + 2] = accumulatedColor.z; acc[pi + 3] = 0.0f;
    const size_t po = ((size_t)py * W + px) * 4;
    out[po] = to_unorm8(color.x); out[po + 1] = to_unorm8(color.y); out[po + 2] = to_unorm8(color.z); out[po + 3] = 255;
}

// AnimationCompute.comp:14-39
void skin_vertices(orc_scene& s) {
    const size_t n = s.vertices_in.size();
    s.vertices.resize(n);
#pragma omp parallel for schedule(static)
    for (long long gi = 0; gi < (long long)n; ++gi) {
        rt_vertex v = s.vertices_in[gi];
        int skin_index = v.skin_index;
        if (skin_index >= 0 && (size_t)skin_index * 4096 + 4096 <= s.skins.size()) {
            const float* bones = &s.skins[(size_t)skin_index * 4096];
            float M[16];
            for (int k = 0; k < 16; ++k)
                M[k] = v.weights[0] * bones[v.joints[0] * 16 + k] + v.weights[1] * bones[v.joints[1] * 16 + k] +
                       v.weights[2] * bones[v.joints[2] * 16 + k] + v.weights[3] * bones[v.joints[3] * 16 + k];
            const mat4& T = *reinterpret_cast<const mat4*>(M);
            vec4 pos = mul(T, V4(v.position[0], v.position[1], v.position[2], 1.0f));
            v.position[0] = pos.x; v.position[1] = pos.y; v.position[2] = pos.z;
            vec3 nn = normalize(xyz(mul(T, V4(v.normal[0], v.normal[1], v.normal[2], 0.0f))));
            v.normal[0] = nn.x; v.normal[1] = nn.y; v.normal[2] = nn.z;
            float w = v.tangent[3];
            vec4 tt = normalize(mul(T, V4(v.tangent[0], v.tangent[1], v.tangent[2], 0.0f)));
            v.tangent[0] = tt.x; v.tangent[1] = tt.y; v.tangent[2] = tt.z; v.tangent[3] = w;
        }
        s.vertices[gi] = v;
    }
}

void build_blas(orc_scene& s, uint32_t g) {
    Geometry& geo = s.geos[g];
    const uint32_t ntri = geo.i_len / 3;
    std::vector<Aabb> boxes(ntri);
    for (uint32_t t = 0; t < ntri; ++t)
        for (int k = 0; k < 3; ++k)
            boxes[t].grow(ld3(s.vertices[geo.v_offset + s.indices[geo.i_offset + 3 * t + k]].position));
    geo.bvh.build(boxes);
}

void build_tlas(orc_scene& s) {
    std::vector<Aabb> boxes(s.instances.size());
    std::vector<uint32_t> refs(s.geos.size(), 0);
    for (auto& in : s.instances) refs[in.geo_id]++;
    for (size_t i = 0; i < s.instances.size(); ++i) {
        Instance& in = s.instances[i];
        invert_3x4(in.o2w, in.w2o);
        const Geometry& g = s.geos[in.geo_id];
        in.baked = refs[in.geo_id] == 1 || g.i_len / 3 <= 256;
        if (in.baked) {
            const uint32_t ntri = g.i_len / 3;
            in.wpos.resize((size_t)ntri * 3);
            std::vector<Aabb> tb(ntri);
            for (uint32_t t = 0; t < ntri; ++t)
                for (int k = 0; k < 3; ++k) {
                    in.wpos[3 * t + k] = xform_point(in.o2w, ld3(s.vertices[g.v_offset + s.indices[g.i_offset + 3 * t + k]].position));
                    tb[t].grow(in.wpos[3 * t + k]);
                }
            in.wbvh.build(tb);
            if (in.wbvh.nodes.empty()) { boxes[i].lo = V3(0.0f); boxes[i].hi = V3(0.0f); }
            else { boxes[i] = in.wbvh.nodes[0].box; vec3 e = (boxes[i].hi - boxes[i].lo) * 1e-5f + V3(1e-6f); boxes[i].lo = boxes[i].lo - e; boxes[i].hi = boxes[i].hi + e; }
            continue;
        }
        if (g.bvh.nodes.empty()) { boxes[i].lo = V3(0.0f); boxes[i].hi = V3(0.0f); continue; }
        const Aabb& b = g.bvh.nodes[0].box;
        for (int c = 0; c < 8; ++c) {
            vec3 p = V3((c & 1) ? b.hi.x : b.lo.x, (c & 2) ? b.hi.y : b.lo.y, (c & 4) ? b.hi.z : b.lo.z);
            boxes[i].grow(xform_point(in.o2w, p));
        }
        // pad: the transformed corners are rounded, keep the box conservative
        vec3 e = (boxes[i].hi - boxes[i].lo) * 1e-5f + V3(1e-6f);
        boxes[i].lo = boxes[i].lo - e; boxes[i].hi = boxes[i].hi + e;
    }
    s.tlas.build(boxes);
}

int fail(const std::string& m) { g_err = m; return 1; }

}  // namespace

// =====================================================================================================
// C API (ctypes)
// =====================================================================================================
extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_scene_create(const rt_scene_desc* d, orc_scene** out) {
    if (!d || !out) return fail("null argument");
    orc_scene* s = new orc_scene();
    s->vertices_in.assign(d->vertices, d->vertices + d->n_vertices);
    s->indices.assign(d->indices, d->indices + d->n_indices);
    s->prim_infos.assign(d->prim_infos, d->prim_infos + d->n_geometries);
    s->materials.assign(d->materials, d->materials + d->n_materials);
    for (uint32_t g = 0; g < d->n_geometries; ++g) {
        Geometry geo;
        geo.v_offset = d->prim_infos[g].v_offset; geo.i_offset = d->prim_infos[g].i_offset;
        geo.material_id = d->prim_infos[g].material_id;
        geo.v_len = d->geometries[g].v_len; geo.i_len = d->geometries[g].i_len; geo.opaque = d->geometries[g].opaque;
        if (geo.i_len % 3 || (uint64_t)geo.i_offset + geo.i_len > d->n_indices || geo.material_id >= d->n_materials) {
            delete s; return fail("geometry out of range");
        }
        s->geos.push_back(std::move(geo));
    }
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        Instance in; std::memcpy(in.o2w, d->instances[i].transform, sizeof in.o2w); in.geo_id = d->instances[i].geo_id;
        if (in.geo_id >= d->n_geometries) { delete s; return fail("instance geo_id out of range"); }
        s->instances.push_back(in);
    }
    for (uint32_t i = 0; i < d->n_images; ++i) {
        Image im; im.w = d->images[i].width; im.h = d->images[i].height; im.srgb = d->images[i].srgb;
        im.px.assign(d->images[i].rgba8, d->images[i].rgba8 + (size_t)im.w * im.h * 4);
        s->images.push_back(std::move(im));
    }
    s->samplers.assign(d->samplers, d->samplers + d->n_samplers);
    s->textures.assign(d->textures, d->textures + d->n_textures);
    for (auto& t : s->textures)
        if (t.image_index >= s->images.size() || t.sampler_index >= s->samplers.size()) { delete s; return fail("texture out of range"); }
    s->dlights.assign(d->dlights, d->dlights + d->n_dlights);
    s->plights.assign(d->plights, d->plights + d->n_plights);
    if (d->skins && d->n_skins) s->skins.assign(d->skins, d->skins + (size_t)d->n_skins * 4096);
    if (d->skybox_faces[0] && d->skybox_width) {
        for (int f = 0; f < 6; ++f) {
            s->sky[f].w = d->skybox_width; s->sky[f].h = d->skybox_height; s->sky[f].srgb = d->skybox_srgb;
            s->sky[f].px.assign(d->skybox_faces[f], d->skybox_faces[f] + (size_t)d->skybox_width * d->skybox_height * 4);
        }
        s->has_sky_faces = true;
    }
    for (int i = 0; i < 256; ++i) {
        double c = i / 255.0;
        s->srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    skin_vertices(*s);
    for (uint32_t g = 0; g < s->geos.size(); ++g) build_blas(*s, g);
    build_tlas(*s);
    *out = s;
    return 0;
}

void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_scene_update_instances(orc_scene* s, const rt_instance* inst, uint32_t n) {
    if (n != s->instances.size()) return fail("instance count mismatch");
    for (uint32_t i = 0; i < n; ++i) {
        std::memcpy(s->instances[i].o2w, inst[i].transform, sizeof(float) * 12);
        s->instances[i].geo_id = inst[i].geo_id;
    }
    build_tlas(*s);
    return 0;
}

int orc_scene_update_skins(orc_scene* s, const float* mats, uint32_t n_skins) {
    s->skins.assign(mats, mats + (size_t)n_skins * 4096);
    skin_vertices(*s);
    for (uint32_t g = 0; g < s->geos.size(); ++g) build_blas(*s, g);   // reference: full rebuild (main.rs:388-395)
    build_tlas(*s);
    return 0;
}

int orc_scene_update_lights(orc_scene* s, const rt_light* d, uint32_t nd, const rt_light* p, uint32_t np) {
    s->dlights.assign(d, d + nd); s->plights.assign(p, p + np); return 0;
}

int orc_scene_read_vertices(orc_scene* s, rt_vertex* out, uint32_t n) {
    if (n > s->vertices.size()) return fail("too many vertices requested");
    std::memcpy(out, s->vertices.data(), (size_t)n * sizeof(rt_vertex)); return 0;
}

static void fill_hit(const orc_scene* s, const Hit& h, rt_hit* o) {
    if (h.valid) { o->t = h.t; o->u = h.u; o->v = h.v; o->instance_id = h.inst; o->primitive_id = h.prim; o->geo_id = s->instances[h.inst].geo_id; }
    else { o->t = -1.0f; o->u = o->v = 0; o->instance_id = o->primitive_id = o->geo_id = 0xFFFFFFFFu; }
}

int orc_trace_closest(orc_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, rt_hit* hits) {
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; ++i) {
        TraceCtx c; c.ray_opaque = (flags & RT_TRACE_OPAQUE) != 0; c.terminate_first = false;
        c.rng = rng4 ? uvec4{rng4[4 * i], rng4[4 * i + 1], rng4[4 * i + 2], rng4[4 * i + 3]} : uvec4{0, 0, 0, 0};
        Hit h = trace(*s, ld3(rays[i].origin), ld3(rays[i].direction), rays[i].tmin, rays[i].tmax, c);
        fill_hit(s, h, &hits[i]);
    }
    return 0;
}

int orc_trace_closest_brute(orc_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, rt_hit* hits) {
#pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < (long long)n; ++i) {
        TraceCtx c; c.ray_opaque = (flags & RT_TRACE_OPAQUE) != 0; c.terminate_first = false;
        c.rng = rng4 ? uvec4{rng4[4 * i], rng4[4 * i + 1], rng4[4 * i + 2], rng4[4 * i + 3]} : uvec4{0, 0, 0, 0};
        Hit h = trace_brute(*s, ld3(rays[i].origin), ld3(rays[i].direction), rays[i].tmin, rays[i].tmax, c);
        fill_hit(s, h, &hits[i]);
    }
    return 0;
}

int orc_trace_any(orc_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, uint8_t* occluded) {
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; ++i) {
        TraceCtx c; c.ray_opaque = (flags & RT_TRACE_OPAQUE) != 0; c.terminate_first = true;
        c.rng = rng4 ? uvec4{rng4[4 * i], rng4[4 * i + 1], rng4[4 * i + 2], rng4[4 * i + 3]} : uvec4{0, 0, 0, 0};
        Hit h = trace(*s, ld3(rays[i].origin), ld3(rays[i].direction), rays[i].tmin, rays[i].tmax, c);
        occluded[i] = h.valid ? 1 : 0;
    }
    return 0;
}

// one frame (RayTracing.rgen over W x H).  rows [row0,row1) only when row1 > row0 (bounded CPU-baseline samples).
int orc_render(orc_scene* s, const rt_ubo* ubo, uint32_t W, uint32_t H, float* acc, uint8_t* out, uint32_t row0, uint32_t row1, rt_stats* stats) {
    if (row1 <= row0) { row0 = 0; row1 = H; }
    if (row1 > H) row1 = H;
    if (row0 > row1) row0 = row1;
    if (ubo->total_number_of_samples == 0) return fail("total_number_of_samples must be > 0");
    s->rays_extend = 0; s->rays_shadow = 0; s->shaded = 0; s->tex_taps = 0; s->light_cands = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
    for (long long y = row0; y < (long long)row1; ++y)
        for (uint32_t x = 0; x < W; ++x) raygen(*s, *ubo, x, (uint32_t)y, W, H, acc, out);
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->ms_total = std::chrono::duration<float, std::milli>(t1 - t0).count();
        stats->rays_extend = s->rays_extend; stats->rays_shadow = s->rays_shadow; stats->shaded_hits = s->shaded;
        stats->tex_taps = s->tex_taps; stats->light_cands = s->light_cands;
        stats->pixel_samples = (uint64_t)(row1 - row0) * W * ubo->number_of_samples;
    }
    return 0;
}

// SURVEY.md §8d ray set (iv): the rays the path tracer itself generates at bounce >= min_bounce (incoherent, starting ON
// surfaces with tMin = 0.001, lib/Camera.glsl:2-3) and its shadow rays (tMin = 0.1, RayTracing.rchit:43) for one frame,
// sampled every `stride` pixels in scan order (single-threaded: deterministic order).
int orc_record_bounce_rays(orc_scene* s, const rt_ubo* ubo, uint32_t W, uint32_t H, uint32_t stride, uint32_t min_bounce,
                           rt_ray* rays, uint32_t* rng4, uint32_t max_rays, rt_ray* srays, uint32_t* srng4, uint32_t max_shadow,
                           uint32_t* n_rays, uint32_t* n_shadow) {
    if (!s || !ubo || !n_rays || !n_shadow) return fail("orc_record_bounce_rays: null argument");
    if (ubo->total_number_of_samples == 0) return fail("total_number_of_samples must be > 0");
    if (!stride) stride = 1;
    RayRecorder rec; rec.min_bounce = min_bounce; rec.max_rays = rays ? max_rays : 0; rec.max_shadow = srays ? max_shadow : 0;
    std::vector<float> acc((size_t)W * H * 4, 0.0f); std::vector<uint8_t> out((size_t)W * H * 4);
    s->rec = &rec;
    for (uint64_t p = 0; p < (uint64_t)W * H && (rec.rays.size() < rec.max_rays || rec.shadow.size() < rec.max_shadow); p += stride) {
        raygen(*s, *ubo, (uint32_t)(p % W), (uint32_t)(p / W), W, H, acc.data(), out.data());
    }
    s->rec = nullptr;
    *n_rays = (uint32_t)rec.rays.size(); *n_shadow = (uint32_t)rec.shadow.size();
    if (*n_rays) { std::memcpy(rays, rec.rays.data(), rec.rays.size() * sizeof(rt_ray)); if (rng4) std::memcpy(rng4, rec.rng.data(), rec.rng.size() * 4); }
    if (*n_shadow) { std::memcpy(srays, rec.shadow.data(), rec.shadow.size() * sizeof(rt_ray)); if (srng4) std::memcpy(srng4, rec.shadow_rng.data(), rec.shadow_rng.size() * 4); }
    return 0;
}

// Per-bounce payload trace of one pixel (golden vectors / debugging): writes up to max_records records of
// 16 floats: [b, hit, t, inst, prim, emit.xyz, hitValue.xyz, needScatter, origin.xyz(next), pad]
int orc_trace_pixel(orc_scene* s, const rt_ubo* ubo, uint32_t W, uint32_t H, uint32_t px, uint32_t py, float* records, uint32_t max_records) {
    // re-run raygen logic for sample 0 only, recording
    const mat4& MVI = *reinterpret_cast<const mat4*>(ubo->model_view_inverse);
    const mat4& PI_ = *reinterpret_cast<const mat4*>(ubo->projection_inverse);
    const uint32_t clk = InitRandomSeed(ubo->total_number_of_samples, ubo->random_seed);
    RayPayload Ray; Ray.RandomSeed = InitRandomSeed(InitRandomSeed(px, py), clk);
    uvec4 rngState = {px, py, ubo->frame_count, 0}; Ray.rngState = {px, py, clk, 0};
    vec2 pixel; vec2 pixelCenter = V2((float)px + 0.5f, (float)py + 0.5f);
    if (ubo->antialiasing) { float ox = rnd(rngState), oy = rnd(rngState); pixel = pixelCenter + (V2(ox, oy) - 0.5f); } else pixel = pixelCenter;
    const vec2 uv = (pixel / V2((float)W, (float)H)) * 2.0f - 1.0f;
    vec2 offset = ubo->aperture / 2 * RandomInUnitDisk(Ray.RandomSeed);
    vec4 origin = mul(MVI, V4(offset.x, offset.y, 0, 1));
    vec4 target = mul(PI_, V4(uv.x, uv.y, 1, 1));
    vec4 direction = mul(MVI, V4(normalize(xyz(target) * ubo->focus_distance - V3(offset.x, offset.y, 0)), 0));
    vec3 throughput = V3(1.0f); Ray.t = 0; Ray.volume_dis = -1.0f;
    uint32_t nrec = 0;
    for (uint32_t b = 0; b < ubo->number_of_bounces && nrec < max_records; b++) {
        TraceCtx c; c.ray_opaque = ubo->fully_opaque != 0; c.terminate_first = false; c.rng = Ray.rngState;
        Hit h = trace(*s, xyz(origin), xyz(direction), tMin, tMax, c);
        if (h.valid) closest_hit(*s, *ubo, h, xyz(direction), Ray); else miss_shader(*s, *ubo, xyz(direction), Ray);
        float* r = records + 16 * nrec++;
        r[0] = (float)b; r[1] = h.valid; r[2] = h.t; r[3] = (float)h.inst; r[4] = (float)h.prim;
        r[5] = Ray.emittance.x; r[6] = Ray.emittance.y; r[7] = Ray.emittance.z;
        r[8] = Ray.hitValue.x; r[9] = Ray.hitValue.y; r[10] = Ray.hitValue.z; r[11] = Ray.needScatter;
        r[12] = Ray.hitPoint.x; r[13] = Ray.hitPoint.y; r[14] = Ray.hitPoint.z; r[15] = Ray.volume_dis;
        const vec3 hitColor = Ray.hitValue; const float t = Ray.t; const bool isScattered = Ray.needScatter;
        Ray.emittance = V3(0.0f);
        if (b + 1 == ubo->number_of_bounces) break;
        if (b > 3) { float p = clampf(luminance(throughput), 0.01f, 0.95f); float q = rnd(rngState); if (p < q) break; else throughput /= p; }
        throughput *= hitColor;
        if (!isScattered || t < 0) break;
        origin = V4(Ray.hitPoint, 1.0f); direction = V4(Ray.scatterDirection, 0);
    }
    return (int)nrec;
}

// ---- known-answer helpers for the RNG / BSDF golden vectors ----
uint32_t orc_tea(uint32_t a, uint32_t b) { return InitRandomSeed(a, b); }
void orc_pcg4d(const uint32_t* in4, uint32_t* out4) { uvec4 r = pcg4d({in4[0], in4[1], in4[2], in4[3]}); out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w; }
float orc_rand(uint32_t* state4) { uvec4 s{state4[0], state4[1], state4[2], state4[3]}; float r = rnd(s); state4[3] = s.w; return r; }
float orc_lcg_float(uint32_t* seed) { return RandomFloat(*seed); }
void orc_offset_ray(const float* p, const float* n, float* out) { vec3 r = offset_ray(ld3(p), ld3(n)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_tonemap(uint32_t mode, const float* rgb, float* out) {
    vec3 c = ld3(rgb), r;
    switch (mode) { case 0: r = defaultToneMap(c); break; case 1: r = toneMapUncharted(c); break; case 2: r = toneMapHejlRichard(c); break;
                    case 3: r = toneMapACES(c); break; default: r = LINEARtoSRGB(c); }
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// evaluates BSDF sampling for a synthetic material: in = [N(3) G(3) V(3) base(3) metallic rough ior transmission volume frontFace type u0 u1 volume_dis]
// out = [ok dir(3) weight(3) volume_dis pS pD pT evalCombined(3 for L=dir)]
void orc_bsdf_sample(const float* in, float* out) {
    MaterialBrdf m{};
    vec3 N = ld3(in), G = ld3(in + 3), V = ld3(in + 6);
    m.baseColor = ld3(in + 9); m.metallic = in[12]; m.roughness = in[13]; m.ior = in[14]; m.transmission = in[15];
    m.volume = in[16] != 0; m.frontFace = in[17] != 0; m.specular_factor = 1.0f; m.specular_color_factor = V3(1.0f); m.use_spec = false;
    m.attenuation_color = V3(0.9f, 0.8f, 0.7f); m.attenuation_distance = 1.5f; m.t_diff = 0;
    matBuild(m);
    uint32_t type = (uint32_t)in[18]; vec2 u = V2(in[19], in[20]); float vd = in[21];
    vec3 dir = V3(0.0f), w = V3(0.0f);
    bool ok = evalIndirectCombinedBRDF(u, N, G, V, m, type, dir, w, vd);
    BRDFp p = getBrdfProbability(m, V, N);
    vec3 e = ok ? evalCombinedBRDF(N, dir, V, m) : V3(0.0f);
    out[0] = ok; out[1] = dir.x; out[2] = dir.y; out[3] = dir.z; out[4] = w.x; out[5] = w.y; out[6] = w.z; out[7] = vd;
    out[8] = p.specular; out[9] = p.diffuse; out[10] = p.transmission; out[11] = e.x; out[12] = e.y; out[13] = e.z;
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
