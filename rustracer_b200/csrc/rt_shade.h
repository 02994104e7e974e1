// rt_shade.h — wavefront restatement of the reference's programmable stages.
//   raygen            RayTracing.rgen:25-130 (camera ray, per-sample state)        -> raygen_path()
//   closest hit       RayTracing.rchit:136-477 + lib/PBR.glsl + lib/Material.glsl   -> shade_hit()
//   miss              RayTracing.rmiss:14-45                                         -> shade_miss()
//   bounce epilogue   RayTracing.rgen:94-129 (radiance, RR, throughput, next ray)   -> inside shade_path()
//   accumulate        RayTracing.rgen:132-166 + lib/Tonemapping.glsl + Heatmap.glsl -> accumulate_pixel()
// The megakernel's recursion (raygen -> chit -> shadow trace) becomes queues: shade_path() emits at most one
// continuation ray and one shadow ray per hit (recursion depth 2).
#pragma once
#include "rt_traverse.h"

#define RT_PI 3.141592653589f
#define RT_ONE_OVER_PI (1.0f / RT_PI)
#define RT_TWO_PI (2.0f * RT_PI)
#define RT_TMIN 0.001f
#define RT_TMAX 10000.0f
enum { RT_DIFFUSE = 1, RT_SPECULAR = 2, RT_TRANSMISSION = 3 };

RT_D float luminance(f3 c) { return dot(c, mk3(0.2126f, 0.7152f, 0.0722f)); }

// ---- lib/PBR.glsl -----------------------------------------------------------------------------------
struct Surface {           // MaterialBrdf (PBR.glsl:173-193), only the fields that are read
    f3 F0, F90, c_diff;
    float roughness, ior, transmission, specular_factor;
    f3 attenuation_color; float attenuation_distance;
    bool volume, front_face;
};
RT_D f3 fresnel_schlick(f3 f0, float f90, float NdotS) { return f0 + (f90 - f0) * pow5(1.0f - NdotS); }   // :228-231
RT_D float shadowed_f90(f3 F90) { return fminf(1.0f, luminance(F90)); }                                           // :312-322
RT_D float smith_g_a(float alpha, float NdotS) { return NdotS / (fmaxf(0.00001f, alpha) * sqrtf(1.0f - fminf(0.99999f, NdotS * NdotS))); }
RT_D float smith_lambda_ggx(float a) { return (-1.0f + sqrtf(1.0f + (1.0f / (a * a)))) * 0.5f; }
RT_D float smith_g2_height_correlated(float alpha, float NdotL, float NdotV) {                                     // :258-262
    return 1.0f / (1.0f + smith_lambda_ggx(smith_g_a(alpha, NdotL)) + smith_lambda_ggx(smith_g_a(alpha, NdotV)));
}
RT_D float ggx_d(float a2, float NdotH) { float b = ((a2 - 1.0f) * NdotH * NdotH + 1.0f); return a2 / (RT_PI * b * b); }   // :286-289
RT_D float smith_g1_ggx(float a2, float NdotS2) { return 2.0f / (sqrtf(((a2 * (1.0f - NdotS2)) + NdotS2) / NdotS2) + 1.0f); }   // :446-448
RT_D float smith_g2_over_g1(float a2, float NdotL, float NdotV) {                                                  // :470-474
    float G1V = smith_g1_ggx(a2, NdotV * NdotV), G1L = smith_g1_ggx(a2, NdotL * NdotL);
    return G1L / (G1V + G1L - G1V * G1L);
}
struct BrdfTerms { f3 F; float NdotL, NdotV, LdotH, NdotH, alpha, alpha2; bool Vback, Lback; };
RT_D BrdfTerms prepare_brdf(f3 N, f3 L, f3 V, const Surface& m) {   // prepareBRDFData :527-564
    BrdfTerms d;
    const f3 H = normalize(L + V);
    const float NdotL = dot(N, L), NdotV = dot(N, V);
    d.Vback = (NdotV <= 0.0f); d.Lback = (NdotL <= 0.0f);
    d.NdotL = fminf(fmaxf(0.00001f, NdotL), 1.0f); d.NdotV = fminf(fmaxf(0.00001f, NdotV), 1.0f);
    d.LdotH = saturate(dot(L, H)); d.NdotH = saturate(dot(N, H));
    const float VdotH = saturate(dot(V, H));
    d.alpha = m.roughness * m.roughness; d.alpha2 = d.alpha * d.alpha;
    d.F = fresnel_schlick(m.F0, shadowed_f90(m.F90), VdotH);
    return d;
}
RT_D float frostbite_diffuse(const BrdfTerms& d, float roughness) {   // :486-496
    const float energyBias = 0.5f * roughness, energyFactor = mixf(1.0f, 1.0f / 1.51f, roughness);
    const float FD90MinusOne = energyBias + 2.0f * d.LdotH * d.LdotH * roughness - 1.0f;
    const float FDL = 1.0f + (FD90MinusOne * pow5(1.0f - d.NdotL));
    const float FDV = 1.0f + (FD90MinusOne * pow5(1.0f - d.NdotV));
    return FDL * FDV * energyFactor;
}
RT_D f3 eval_combined_brdf(f3 N, f3 L, f3 V, const Surface& m) {   // evalCombinedBRDF :731-745
    const BrdfTerms d = prepare_brdf(N, L, V, m);
    if (d.Vback || d.Lback) return mk3(0.0f);
    const float D = ggx_d(fmaxf(0.00001f, d.alpha2), d.NdotH);
    const float G2 = smith_g2_height_correlated(d.alpha, d.NdotL, d.NdotV);
    const f3 specular = ((d.F * G2 * D) / (4.0f * d.NdotL * d.NdotV)) * d.NdotL;                 // evalMicrofacet :292-303
    const f3 diffuse = m.c_diff * (frostbite_diffuse(d, m.roughness) * RT_ONE_OVER_PI * d.NdotL);   // :516-518
    return diffuse + specular;
}
struct LobeProb { float specular, diffuse, transmission; };
RT_D LobeProb brdf_probability(const Surface& m, f3 V, f3 N) {   // getBrdfProbability :324-358
    const float specularF0 = luminance(m.F0), diffuseReflectance = luminance(m.c_diff);
    const float Fresnel = saturate(luminance(fresnel_schlick(mk3(specularF0), shadowed_f90(m.F90), fmaxf(0.0f, dot(V, N)))));
    const float specular = Fresnel * m.specular_factor;
    const float penetration = diffuseReflectance * (1.0f - m.specular_factor * Fresnel);
    const float diffuse = penetration * (1.0f - m.transmission), transmission = penetration * m.transmission;
    float sum = fmaxf(0.0001f, (specular + diffuse + transmission));
    float p = clampf(specular / sum, 0.001f, 0.9f);
    float d = (1.0f - p) * (1.0f - m.transmission), t = (1.0f - p) * m.transmission;
    sum = p + d + t;
    LobeProb r; r.specular = p / sum; r.diffuse = div_pos(d, sum); r.transmission = div_pos(t, sum);    // sum >= p >= 0.001
    return r;
}
RT_D f4 rotation_to_z(f3 v) {   // getRotationToZAxis :362-368
    if (v.z < -0.99999f) return mk4(1.0f, 0.0f, 0.0f, 0.0f);
    return normalize(mk4(v.y, -v.x, 0.0f, 1.0f + v.z));
}
RT_D f3 rotate_point(f4 q, f3 v) {   // :373-376
    const f3 a = mk3(q.x, q.y, q.z);
    return 2.0f * dot(a, v) * a + (q.w * q.w - dot(a, a)) * v + 2.0f * q.w * cross(a, v);
}
RT_D f3 sample_ggx_vndf(f3 Ve, float alpha, f2 u) {   // sampleGGXVNDF :413-436 (isotropic alpha2D)
    const f3 Vh = normalize(mk3(alpha * Ve.x, alpha * Ve.y, Ve.z));
    const float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    const f3 T1 = lensq > 0.0f ? mk3(-Vh.y, Vh.x, 0.0f) * (1.0f / sqrtf(lensq)) : mk3(1.0f, 0.0f, 0.0f);
    const f3 T2 = cross(Vh, T1);
    const float r = sqrtf(u.x), phi = RT_TWO_PI * u.y;
    const float t1 = r * cosf(phi); float t2 = r * sinf(phi);
    const float s = 0.5f * (1.0f + Vh.z);
    t2 = mixf(sqrtf(1.0f - t1 * t1), t2, s);
    const f3 Nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(mk3(alpha * Nh.x, alpha * Nh.y, fmaxf(0.0f, Nh.z)));
}
// evalIndirectCombinedBRDF :649-729.  Returns needScatter.
RT_D bool sample_bsdf(f2 u, f3 N, f3 geoN, f3 V, const Surface& m, int lobe, f3& dir_out, f3& weight, float& volume_dis) {
    if (dot(geoN, V) < 0.0f) return false;
    const f4 q = rotation_to_z(N);
    const f3 Vl = rotate_point(q, V);
    const f3 Nl = mk3(0.0f, 0.0f, 1.0f);
    f3 Ll = mk3(0.0f);
    if (lobe == RT_DIFFUSE) {
        const float a = sqrtf(u.x), b = RT_TWO_PI * u.y;                 // sampleHemisphere :381-394
        Ll = mk3(a * cosf(b), a * sinf(b), sqrtf(1.0f - u.x));
        const BrdfTerms d = prepare_brdf(Nl, Ll, Vl, m);
        weight = (1.0f - m.specular_factor * d.F) * m.c_diff * frostbite_diffuse(d, m.roughness);
        weight *= (1.0f - m.transmission);
    } else if (lobe == RT_SPECULAR) {
        const float alpha = m.roughness * m.roughness, alpha2 = alpha * alpha;
        const f3 Hl = (alpha == 0.0f) ? Nl : sample_ggx_vndf(Vl, alpha, u);   // sampleSpecularMicrofacet :587-616
        Ll = reflect3(-Vl, Hl);
        const float HdotL = fmaxf(0.00001f, fminf(1.0f, dot(Hl, Ll)));
        const float NdotL = fmaxf(0.00001f, fminf(1.0f, Ll.z)), NdotV = fmaxf(0.00001f, fminf(1.0f, Vl.z));
        const f3 F = fresnel_schlick(m.F0, shadowed_f90(m.F90), HdotL);
        weight = F * smith_g2_over_g1(alpha2, NdotL, NdotV);
        weight *= m.specular_factor;
    } else {
        if (m.volume) {
            const float eta = m.front_face ? 1.0f / m.ior : m.ior;
            const f3 refr = refract3(-Vl, Nl, eta);
            if (is_zero(refr)) { weight = mk3(0.0f); return false; }
            Ll = refr;
        } else {
            Ll = -Vl;
        }
        weight = max3(mk3(0.0f), m.c_diff * m.transmission);
        if (!m.front_face && m.volume) {
            const float dis = volume_dis;
            volume_dis = -1.0f;
            const f3 sigma = log3(m.attenuation_color) / m.attenuation_distance;
            weight *= min3(exp3(sigma * dis), mk3(1.0f));
        }
    }
    if (luminance(weight) == 0.0f) return false;
    const f4 qi = mk4(-q.x, -q.y, -q.z, q.w);
    dir_out = normalize(rotate_point(qi, Ll));
    return true;
}
RT_D float convert_metallic(f3 diffuse, f3 specular, float maxSpecular) {   // lib/Material.glsl:78-89
    const float kMin = 0.04f;
    const float pd = sqrtf(0.299f * diffuse.x * diffuse.x + 0.587f * diffuse.y * diffuse.y + 0.114f * diffuse.z * diffuse.z);
    const float ps = sqrtf(0.299f * specular.x * specular.x + 0.587f * specular.y * specular.y + 0.114f * specular.z * specular.z);
    if (ps < kMin) return 0.0f;
    const float a = kMin, b = pd * (1.0f - maxSpecular) / (1.0f - kMin) + ps - 2.0f * kMin, c = kMin - ps;
    const float D = fmaxf(b * b - 4.0f * a * c, 0.0f);
    return clampf((-b + sqrtf(D)) / (2.0f * a), 0.0f, 1.0f);
}
RT_D f3 light_intensity_at(const rt_light& l, float distance) {   // lib/PunctualLight.glsl:17-30
    const f3 color = l.intensity * mk3(l.color[0], l.color[1], l.color[2]);
    if (l.kind == 1u) {
        const float r2 = 0.25f, d2 = distance * distance;
        return color * (2.0f / (d2 + r2 + distance * sqrtf(d2 + r2)));
    }
    return color;
}
RT_D f3 offset_ray(f3 p, f3 n) {   // lib/RayTracingCommons.glsl:103-119
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    const int ox = (int)(n.x * int_scale), oy = (int)(n.y * int_scale), oz = (int)(n.z * int_scale);
    const f3 pi = mk3(rt_uint_as_float((uint32_t)((int)rt_float_as_uint(p.x) + ((p.x < 0) ? -ox : ox))),
                      rt_uint_as_float((uint32_t)((int)rt_float_as_uint(p.y) + ((p.y < 0) ? -oy : oy))),
                      rt_uint_as_float((uint32_t)((int)rt_float_as_uint(p.z) + ((p.z < 0) ? -oz : oz))));
    return mk3(fabsf(p.x) < origin ? p.x + float_scale * n.x : pi.x, fabsf(p.y) < origin ? p.y + float_scale * n.y : pi.y,
               fabsf(p.z) < origin ? p.z + float_scale * n.z : pi.z);
}

// ---- per-path state -----------------------------------------------------------------------------------
struct PathState {
    f3 origin, dir; float tmin, tmax;
    f3 throughput; uint32_t pixel;
    uint32_t path_w, pix_w, lens_seed; float volume_dis;
};
struct FrameParams {       // the UBO plus per-frame derived values
    rt_ubo ubo;
    uint32_t width, height;
    uint32_t clk;          // D1: tea(total_number_of_samples, random_seed) replaces uint(clockARB())
    uint32_t sample;       // index of the sample within this frame (RayTracing.rgen:49)
};
struct ShadowRay { f3 origin, dir; float tmax; f3 contrib; uint32_t pixel, path_w; };

RT_D u4 path_stream(const FrameParams& P, uint32_t pixel, uint32_t w) { u4 s; s.x = pixel % P.width; s.y = pixel / P.width; s.z = P.clk; s.w = w; return s; }
RT_D u4 pixel_stream(const FrameParams& P, uint32_t pixel, uint32_t w) { u4 s; s.x = pixel % P.width; s.y = pixel / P.width; s.z = P.ubo.frame_count; s.w = w; return s; }

// RayTracing.rgen:32-80.  For sample 0 the three streams are created; later samples continue them
// (they live outside the sample loop in the reference).
RT_D PathState raygen_path(const FrameParams& P, uint32_t pixel, uint32_t pix_w, uint32_t path_w, uint32_t lens_seed) {
    const uint32_t px = pixel % P.width, py = pixel / P.width;
    const rt_ubo& ubo = P.ubo;
    if (P.sample == 0) { pix_w = 0; path_w = 0; lens_seed = tea16(tea16(px, py), P.clk); }
    u4 pix = pixel_stream(P, pixel, pix_w);
    f2 pc = mk2((float)px + 0.5f, (float)py + 0.5f);
    if (ubo.antialiasing) { const float ox = rng_next(pix), oy = rng_next(pix); pc = mk2(rt_fadd(pc.x, rt_fsub(ox, 0.5f)), rt_fadd(pc.y, rt_fsub(oy, 0.5f))); }
    // everything up to the primary ray is evaluated with explicitly rounded operations in the oracle's order
    const f2 uv = mk2(rt_fsub(rt_fmul(rt_fdiv(pc.x, (float)P.width), 2.0f), 1.0f), rt_fsub(rt_fmul(rt_fdiv(pc.y, (float)P.height), 2.0f), 1.0f));
    const f2 disk = random_in_unit_disk(lens_seed);
    const float half_ap = rt_fdiv(ubo.aperture, 2.0f);
    const f2 offset = mk2(rt_fmul(half_ap, disk.x), rt_fmul(half_ap, disk.y));
    f4 origin = mat4_mul_exact(ubo.model_view_inverse, mk4(offset.x, offset.y, 0.0f, 1.0f));
    const f4 target = mat4_mul_exact(ubo.projection_inverse, mk4(uv.x, uv.y, 1.0f, 1.0f));
    const f3 tdir = mk3(rt_fsub(rt_fmul(target.x, ubo.focus_distance), offset.x), rt_fsub(rt_fmul(target.y, ubo.focus_distance), offset.y), rt_fsub(rt_fmul(target.z, ubo.focus_distance), 0.0f));
    f4 direction = mat4_mul_exact(ubo.model_view_inverse, mk4(normalize_exact(tdir), 0.0f));
    float tFar = RT_TMAX;
    if (ubo.orthographic_fov_dis > 0.0f) {
        const float k = rt_fadd(1.0f, ubo.orthographic_fov_dis);
        const f2 nuv = mk2(rt_fmul(k, uv.x), rt_fmul(k, uv.y));
        origin = mat4_mul_exact(ubo.model_view_inverse, mk4(nuv.x, -nuv.y, 0.0f, 1.0f));
        direction = mat4_mul_exact(ubo.model_view_inverse, mk4(0.0f, 0.0f, -1.0f, 0.0f));
        tFar = 10.0f * RT_TMAX;
    }
    PathState s;
    s.origin = xyz(origin); s.dir = xyz(direction); s.tmin = RT_TMIN; s.tmax = tFar;
    s.throughput = mk3(1.0f); s.pixel = pixel;
    s.path_w = path_w; s.pix_w = pix.w; s.lens_seed = lens_seed; s.volume_dis = -1.0f;
    return s;
}

struct ShadeOut {
    f3 emittance;          // Ray.emittance without the NEE term (that one is resolved by the shadow pass)
    f3 hit_value;          // Ray.hitValue
    bool need_scatter;
    f3 next_origin, next_dir;
    bool has_shadow; ShadowRay shadow;   // shadow.contrib excludes the path throughput
    float t;               // Ray.t
};

// RayTracing.rmiss:14-45
RT_D void shade_miss(const DScene& S, const FrameParams& P, f3 world_dir, bool primary, ShadeOut& o, RtCounters* cnt) {
    f3 acc = mk3(0.0f);
    const f3 rd = normalize(world_dir);
    if (!primary) {
        for (uint32_t i = 0; i < S.n_dlights; i++) {
            const rt_light& li = S.dlights[i];
            const float c = dot(normalize(mk3(li.transform[0], li.transform[1], li.transform[2])), rd);
            if (c < 0.0f) acc += -c * mk3(li.color[0], li.color[1], li.color[2]) * li.intensity;
        }
    }
    if (P.ubo.has_sky) {
        const f3 sky = S.has_sky_faces ? texture_cube(S, rd) : mk3(0.0f);
        acc += sky + acc;          // sic, RayTracing.rmiss:33 doubles the directional term
    } else {
        acc += mk3(0.01f);
    }
    o.hit_value = mk3(0.0f); o.need_scatter = false; o.emittance = acc;
    if (S.n_dlights == 0) o.emittance = mk3(0.0f);
    o.t = -1.0f; o.has_shadow = false;
}

// RayTracing.rchit:136-477.  `st` carries the payload fields that survive the stage (rng, volume_dis, lens seed).
// SIMPLE: the scene has no textures and no specular-glossiness material and the frame uses mapping == RENDER with
// debug == 0 (checked on the host per frame) — the same kind of specialisation as the reference's any-hit-free pipeline
// for fully opaque scenes (pipeline_res.rs:155-166).  It removes the texture / debug code from the hot kernel.
// COUNT: texture taps and light candidates of this hit are added to `cnt` (SURVEY.md §8d algorithmic bytes).
template <bool SIMPLE, bool COUNT = false>
RT_D void shade_hit(const DScene& S, const FrameParams& P, const RtHit& hit, PathState& st, ShadeOut& o, RtCounters* cnt) {
    const rt_ubo& ubo = P.ubo;
    uint32_t n_taps = 0, n_cands = 0;
#define RT_TEX(idx, uvc) (n_taps += COUNT ? 1u : 0u, texture2d(S, (idx), (uvc)))
    const float4* op = S.inst_o2w + (size_t)hit.inst * RT_O2W_F4;
    const float4 m0 = rt_ld(op), m1 = rt_ld(op + 1), m2 = rt_ld(op + 2), m3 = rt_ld(op + 3);
    rt_prim_info pinfo;     // copy of prim_infos[geo_id] kept in the instance record (one dependent load less)
    pinfo.v_offset = rt_float_as_uint(m3.x); pinfo.i_offset = rt_float_as_uint(m3.y); pinfo.material_id = rt_float_as_uint(m3.z);
    const uint32_t geo_id = rt_float_as_uint(m3.w);
    const rt_material& mat = S.materials[pinfo.material_id];
    const TriIndices ti = fetch_indices(S, pinfo, hit.prim);
    const rt_vertex& v0 = S.vertices[ti.i0]; const rt_vertex& v1 = S.vertices[ti.i1]; const rt_vertex& v2 = S.vertices[ti.i2];
    const float b0 = 1.0f - hit.u - hit.v, b1 = hit.u, b2 = hit.v;

    // getMixVertexAndGeoNormal (RayTracingCommons.glsl:82-97)
    const f3 p0 = ld_f3(v0.position), p1 = ld_f3(v1.position), p2 = ld_f3(v2.position);
    const f4 uv = mk4(v0.uv0[0], v0.uv0[1], v0.uv1[0], v0.uv1[1]) * b0 + mk4(v1.uv0[0], v1.uv0[1], v1.uv1[0], v1.uv1[1]) * b1 + mk4(v2.uv0[0], v2.uv0[1], v2.uv1[0], v2.uv1[1]) * b2;
    const f3 pos = p0 * b0 + p1 * b1 + p2 * b2;
    const f4 vcolor = ld_f4(v0.color) * b0 + ld_f4(v1.color) * b1 + ld_f4(v2.color) * b2;
    f3 normal = normalize(normalize(ld_f3(v0.normal)) * b0 + normalize(ld_f3(v1.normal)) * b1 + normalize(ld_f3(v2.normal)) * b2);
    const f4 tangent = normalize(ld_f4(v0.tangent) * b0 + ld_f4(v1.tangent) * b1 + ld_f4(v2.tangent) * b2);
#define RT_O2W_DIR(n) normalize(mk3(m0.x * (n).x + m0.y * (n).y + m0.z * (n).z, m1.x * (n).x + m1.y * (n).y + m1.z * (n).z, m2.x * (n).x + m2.y * (n).y + m2.z * (n).z))
    const f3 gn_obj = cross(p1 - p0, p2 - p0);
    f3 geo_normal = RT_O2W_DIR(gn_obj);
    f3 origin = mk3(m0.x * pos.x + m0.y * pos.y + m0.z * pos.z + m0.w, m1.x * pos.x + m1.y * pos.y + m1.z * pos.z + m1.w, m2.x * pos.x + m2.y * pos.y + m2.z * pos.z + m2.w);

    f4 color4 = vcolor * ld_f4(mat.base_color);
    if (!SIMPLE && mat.base_color_texture.index >= 0) color4 *= RT_TEX(mat.base_color_texture.index, get_uv(uv, mat.base_color_texture.coord));
    f3 color = xyz(color4);
    if (!SIMPLE && mat.normal_texture.index >= 0) {
        const f3 nt = normalize(xyz(RT_TEX(mat.normal_texture.index, get_uv(uv, mat.normal_texture.coord))) * 2.0f - 1.0f);
        const f3 tm = xyz(tangent);                                         // getNormal :124-129
        const f3 tg = normalize(tm - dot(tm, normal) * normal);
        const f3 bt = normalize(cross(normal, tg) * tangent.w);
        normal = tg * nt.x + bt * nt.y + normal * nt.z;
    }
    normal = RT_O2W_DIR(normal);

    const f3 V = -normalize(st.dir);
    const float cosv = dot(V, geo_normal);
    const bool front_face = cosv >= 0.0f;
    geo_normal = front_face ? geo_normal : -geo_normal;
    const f3 N = dot(geo_normal, normal) < 0.0f ? -normal : normal;

    f3 emissive = mk3(mat.emissive_factor[0], mat.emissive_factor[1], mat.emissive_factor[2]);
    if (!SIMPLE && mat.emissive_texture.index >= 0) emissive *= xyz(RT_TEX(mat.emissive_texture.index, get_uv(uv, mat.emissive_texture.coord)));
    float metallic = mat.metallic_factor, roughness = mat.roughness_factor;
    if (!SIMPLE && mat.metallic_roughness_texture.index >= 0) {
        const f4 mr = RT_TEX(mat.metallic_roughness_texture.index, get_uv(uv, mat.metallic_roughness_texture.coord));
        roughness *= mr.y; metallic *= mr.z;
    }
    f3 spec_wf = mk3(1.0f);
    const bool sg = !SIMPLE && mat.workflow == 1u;
    if (sg) {
        f4 diffuse_factor = ld_f4(mat.sg_diffuse_factor), sgf = ld_f4(mat.sg_specular_glossiness_factor);
        if (mat.sg_diffuse_texture.index >= 0) diffuse_factor *= RT_TEX(mat.sg_diffuse_texture.index, get_uv(uv, mat.sg_diffuse_texture.coord));
        if (mat.sg_specular_glossiness_texture.index >= 0) sgf *= RT_TEX(mat.sg_specular_glossiness_texture.index, get_uv(uv, mat.sg_specular_glossiness_texture.coord));
        spec_wf = xyz(sgf);
        roughness = 1.0f - sgf.w;
        color = xyz(vcolor) * xyz(diffuse_factor);
        metallic = convert_metallic(color, spec_wf, fmaxf(fmaxf(spec_wf.x, spec_wf.y), spec_wf.z));
    }
    float transmission = 0.0f;
    if (mat.transmission_exist) {
        transmission = mat.transmission_factor;
        if (!SIMPLE && mat.transmission_texture.index >= 0) transmission *= RT_TEX(mat.transmission_texture.index, get_uv(uv, mat.transmission_texture.coord)).x;
    }

    o.t = hit.t; o.need_scatter = false; o.has_shadow = false; o.hit_value = mk3(0.0f);
    o.next_origin = pos; o.next_dir = mk3(0.0f);
    if (COUNT && cnt && n_taps) rt_atomic_add64(&cnt->tex_taps, n_taps);   // (taps below this point are flushed at the end)
    n_taps = 0;
    uint32_t mapping = SIMPLE ? (uint32_t)RT_MAP_RENDER : ubo.mapping;
    if (mat.unlit) mapping = RT_MAP_ALBEDO;
    switch (mapping) {   // :258-286 debug channels return through Ray.emittance
        case RT_MAP_ALBEDO: o.emittance = color; return;
        case RT_MAP_TRIANGLE: o.emittance = mk3(1.0f - hit.u - hit.v, hit.u, hit.v); return;
        case RT_MAP_INSTANCE: o.emittance = hash_and_color(hit.inst); return;
        case RT_MAP_METALLIC: o.emittance = mk3(metallic); return;
        case RT_MAP_ROUGHNESS: o.emittance = mk3(roughness); return;
        case RT_MAP_NORMAL: o.emittance = (N + 1.0f) / 2.0f; return;
        case RT_MAP_TANGENT: { const f3 tx = xyz(tangent); o.emittance = (RT_O2W_DIR(tx) + 1.0f) / 2.0f; return; }
        case RT_MAP_TRANSMISSION: o.emittance = mk3(transmission); return;
        case RT_MAP_GEO_ID: o.emittance = hash_and_color(geo_id); return;
        default: break;
    }
#undef RT_O2W_DIR

    float spec_factor = mat.specular_factor;
    f3 spec_color = mk3(mat.specular_color_factor[0], mat.specular_color_factor[1], mat.specular_color_factor[2]);
    if (!SIMPLE && mat.specular_texture.index >= 0) spec_factor *= RT_TEX(mat.specular_texture.index, get_uv(uv, mat.specular_texture.coord)).w;
    if (!SIMPLE && mat.specular_color_texture.index >= 0) spec_color *= xyz(RT_TEX(mat.specular_color_texture.index, get_uv(uv, mat.specular_color_texture.coord)));

    o.emittance = emissive * ubo.exposure;
    Surface m;
    m.roughness = roughness; m.ior = mat.volume_exists ? mat.ior : 1.0f;   // :320
    m.transmission = transmission; m.specular_factor = spec_factor; m.front_face = front_face;
    {   // matBuild :195-206
        const float f = m.ior + 1.0f > 0.0f ? div_pos(m.ior - 1.0f, m.ior + 1.0f) : (m.ior - 1.0f) / (m.ior + 1.0f);
        const f3 dF0 = min3(f * f * spec_color, mk3(1.0f)) * spec_factor;
        m.F0 = mix3(dF0, color, metallic); m.F90 = mix3(spec_color, mk3(1.0f), metallic); m.c_diff = mix3(color, mk3(0.0f), metallic);
    }
    if (sg) { m.c_diff = color * (1.0f - fmaxf(fmaxf(spec_wf.x, spec_wf.y), spec_wf.z)); m.F0 = spec_wf; }
    m.attenuation_color = mk3(mat.attenuation_color[0], mat.attenuation_color[1], mat.attenuation_color[2]);
    m.attenuation_distance = mat.attenuation_distance; m.volume = mat.volume_exists != 0u;
    // last_hit == the ray origin for b > 0 (rgen sets origin = Ray.hitPoint); for b == 0 it only matters when
    // volume_dis >= 0, which cannot happen (D8)
    const float displacement = length(origin - st.origin);

    u4 rng = path_stream(P, st.pixel, st.path_w);
    const uint32_t entry_w = st.path_w;
    // NEE: sampleLightRIS :77-122 + castShadowRay :35-59 (resolved by the shadow pass)
    if (S.nee_plights) {      // (0 when no candidate slot can pass the luminance test below: the loop would draw nothing)
        float totalWeights = 0.0f, samplePdfG = 0.0f; uint32_t sel = 0;
        const uint32_t ncand = S.n_plights < 3u ? S.n_plights : 3u;
        for (uint32_t i = 0; i < ncand; i++) {
            const rt_light& li = S.plights[i];
            if (luminance(mk3(li.color[0], li.color[1], li.color[2]) * li.intensity) < 0.1f) continue;
            if (COUNT) ++n_cands;
            uint32_t k = (uint32_t)(rng_next(rng) * (float)S.n_plights);
            if (k > S.n_plights - 1u) k = S.n_plights - 1u;
            const rt_light& cand = S.plights[k];
            const f3 lv = mk3(cand.transform[0], cand.transform[1], cand.transform[2]) - origin;
            const float ld = length(lv);
            if (dot(geo_normal, normalize(lv)) < 0.00001f) continue;
            const float pdfG = luminance(light_intensity_at(cand, ld));
            const float w = pdfG * (float)S.n_plights;
            totalWeights += w;
            if (rng_next(rng) < (w / totalWeights)) { sel = k; samplePdfG = pdfG; }
        }
        if (totalWeights != 0.0f) {
            const rt_light& light = S.plights[sel];
            const float lw = (totalWeights / 3.0f) / samplePdfG;
            f3 lv = mk3(light.transform[0], light.transform[1], light.transform[2]) - origin;
            const float dist = length(lv);
            lv = normalize(lv);
            o.has_shadow = true;
            o.shadow.origin = origin; o.shadow.dir = lv; o.shadow.tmax = dist; o.shadow.pixel = st.pixel; o.shadow.path_w = entry_w;
            o.shadow.contrib = eval_combined_brdf(N, lv, V, m) * lw * light.intensity * mk3(light.color[0], light.color[1], light.color[2]);
        }
    }

    int lobe; f3 thr = mk3(1.0f);
    if (metallic == 1.0f && roughness == 0.0f) {
        lobe = RT_SPECULAR;
    } else {
        const LobeProb bp = brdf_probability(m, V, N);
        const float r = rng_next(rng);
        if (r < bp.specular) { lobe = RT_SPECULAR; thr /= bp.specular; if (st.volume_dis >= 0.0f) st.volume_dis += displacement; }
        else if (r >= bp.specular && r <= bp.specular + bp.diffuse) { lobe = RT_DIFFUSE; thr /= bp.diffuse; if (st.volume_dis >= 0.0f) st.volume_dis += displacement; }
        else {
            lobe = RT_TRANSMISSION;
            if (m.volume) { if (st.volume_dis >= 0.0f) st.volume_dis += displacement; else if (front_face) st.volume_dis = 0.0f; }
            thr /= bp.transmission;
        }
    }
    if (lobe == RT_TRANSMISSION) origin = offset_ray(origin, -geo_normal);
    f3 weight = mk3(0.0f), ndir = mk3(0.0f);
    const float u0 = rng_next(rng), u1 = rng_next(rng);
    o.need_scatter = sample_bsdf(mk2(u0, u1), N, geo_normal, V, m, lobe, ndir, weight, st.volume_dis);
    thr *= weight;
    o.next_origin = origin; o.next_dir = ndir; o.hit_value = thr;

    if (!SIMPLE && ubo.debug == 1u) {   // :437-474 legacy path on the LCG stream
        uint32_t seed = st.lens_seed;
        const f3 wd = st.dir;
        if (m.transmission > 0.0f) {
            const float eta = front_face ? 1.0f / mat.ior : mat.ior;
            const f3 refr = refract3(wd, N, eta);
            float r0 = (1.0f - eta) / (1.0f + eta); r0 *= r0;
            const float reflectProb = !is_zero(refr) ? (r0 + (1.0f - r0) * pow5(1.0f - fabsf(cosv))) : 1.0f;
            o.hit_value = color; o.need_scatter = true;
            if (lcg_float(seed) < reflectProb) o.next_dir = reflect3(wd, normal); else o.next_dir = refr;
        } else if (length(emissive) < 0.01f && roughness == 1.0f) {
            const bool sc = dot(wd, geo_normal) < 0.0f;
            o.next_dir = normalize(N + random_in_unit_sphere(seed));
            o.need_scatter = sc; o.hit_value = sc ? color : mk3(0.0f);
        } else if (metallic > 0.0f) {
            const f3 refl = reflect3(wd, N);
            const bool sc = dot(refl, geo_normal) > 0.0f;
            o.need_scatter = sc; o.hit_value = sc ? color : mk3(0.0f);
            o.next_dir = refl + 0.08f * random_in_unit_sphere(seed);
        }
        st.lens_seed = seed;
    }
    st.path_w = rng.w;
    if (COUNT && cnt) { if (n_taps) rt_atomic_add64(&cnt->tex_taps, n_taps); if (n_cands) rt_atomic_add64(&cnt->light_cands, n_cands); }
#undef RT_TEX
}

// ---- lib/Tonemapping.glsl, lib/Heatmap.glsl, RayTracing.rgen:132-166 ------------------------------------
RT_D f3 linear_to_srgb(f3 c) { return pow3(c, 1.0f / 2.2f); }
RT_D f3 uncharted2(f3 c) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((c * (A * c + C * B) + D * E) / (c * (A * c + B) + D * F)) - E / F;
}
RT_D f3 tonemap(uint32_t mode, f3 c) {
    switch (mode) {
        case 0: return linear_to_srgb(c / (c + 1.0f));
        case 1: { const f3 col = uncharted2(c * 2.0f); const f3 ws = 1.0f / uncharted2(mk3(11.2f)); return linear_to_srgb(col * ws); }
        case 2: { const f3 k = max3(mk3(0.0f), c - mk3(0.004f)); return (k * (6.2f * k + 0.5f)) / (k * (6.2f * k + 1.7f) + 0.06f); }
        case 3: return linear_to_srgb(clamp3((c * (2.51f * c + 0.03f)) / (c * (2.43f * c + 0.59f) + 0.14f), 0.0f, 1.0f));
        default: return linear_to_srgb(c);
    }
}
RT_D f3 heatmap(float t) {
    const float c[10][3] = {{0.0f / 255.0f, 2.0f / 255.0f, 91.0f / 255.0f}, {0.0f / 255.0f, 108.0f / 255.0f, 251.0f / 255.0f}, {0.0f / 255.0f, 221.0f / 255.0f, 221.0f / 255.0f},
                            {51.0f / 255.0f, 221.0f / 255.0f, 0.0f / 255.0f}, {255.0f / 255.0f, 252.0f / 255.0f, 0.0f / 255.0f}, {255.0f / 255.0f, 180.0f / 255.0f, 0.0f / 255.0f},
                            {255.0f / 255.0f, 104.0f / 255.0f, 0.0f / 255.0f}, {226.0f / 255.0f, 22.0f / 255.0f, 0.0f / 255.0f}, {191.0f / 255.0f, 0.0f / 255.0f, 83.0f / 255.0f},
                            {145.0f / 255.0f, 0.0f / 255.0f, 65.0f / 255.0f}};
    const float s = t * 10.0f;
    const int cur = (int)s <= 9 ? (int)s : 9, prv = cur >= 1 ? cur - 1 : 0, nxt = cur < 9 ? cur + 1 : 9;
    const float blur = 0.8f;
    const float wc = smoothstepf((float)cur - blur, (float)cur + blur, s) * (1.0f - smoothstepf((float)(cur + 1) - blur, (float)(cur + 1) + blur, s));
    const float wp = 1.0f - smoothstepf((float)cur - blur, (float)cur + blur, s);
    const float wn = smoothstepf((float)(cur + 1) - blur, (float)(cur + 1) + blur, s);
    const f3 r = mk3(wc * c[cur][0] + wp * c[prv][0] + wn * c[nxt][0], wc * c[cur][1] + wp * c[prv][1] + wn * c[nxt][1], wc * c[cur][2] + wp * c[prv][2] + wn * c[nxt][2]);
    return clamp3(r, 0.0f, 1.0f);
}
RT_D uint32_t to_unorm8(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 1.0f) return 255u;
    return (uint32_t)rintf(v * 255.0f);
}
// RGBA8 of an accumulated radiance sum (RayTracing.rgen:150-166 without the debug overrides): used by the multi-GPU combine
RT_D uint32_t tonemap_rgba8(const rt_ubo& ubo, f3 sum) {
    const f3 color = tonemap(ubo.tone_mapping_mode, sum / (float)ubo.total_number_of_samples);
    return to_unorm8(color.x) | (to_unorm8(color.y) << 8) | (to_unorm8(color.z) << 16) | 0xFF000000u;
}
// frame_rad: radiance gathered this frame (all samples).  last_t / n_traces feed the DISTANCE / HEAT mappings.
RT_D void accumulate_pixel(const rt_ubo& ubo, float4* acc, uint32_t* out, size_t pixel, f3 frame_rad, float last_t, uint32_t n_traces) {
    const bool accumulate = ubo.number_of_samples != ubo.total_number_of_samples;
    f3 sum = frame_rad;
    if (accumulate) { const float4 a = acc[pixel]; sum = mk3(a.x, a.y, a.z) + frame_rad; }
    const f3 radiance = sum / (float)ubo.total_number_of_samples;
    f3 color = tonemap(ubo.tone_mapping_mode, radiance);
    if (ubo.mapping == RT_MAP_HEAT) {
        // D9: clockARB() delta replaced by 100000 ticks per traced path segment
        const float scale = 1000000.0f * ubo.heatmap_scale * ubo.heatmap_scale;
        color = heatmap(clampf((float)((unsigned long long)n_traces * 100000ull) / scale, 0.0f, 1.0f));
    } else if (ubo.mapping == RT_MAP_DISTANCE) {
        color = mk3(fminf((ubo.heatmap_scale - fmaxf(last_t, RT_TMIN)) / ubo.heatmap_scale, 1.0f));
    }
    acc[pixel] = make_float4(sum.x, sum.y, sum.z, 0.0f);
    out[pixel] = to_unorm8(color.x) | (to_unorm8(color.y) << 8) | (to_unorm8(color.z) << 16) | 0xFF000000u;
}
