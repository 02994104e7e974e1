// rt_traverse.h — two-level traversal of the compressed 8-wide BVH + watertight ray/triangle test.
// Replaces what the reference delegates to the Vulkan driver: traceRayEXT at RayTracing.rgen:89-92 (closest
// hit, optional any-hit) and RayTracing.rchit:51-55 (shadow: terminate on first hit, skip closest hit).
// Semantics pinned by the oracle (oracle/oracle.cpp): open interval tmin < t < tmax, two-sided triangles,
// closest hit = lexicographic minimum of (t, instance, primitive), alpha test before commit.
#pragma once
#include "rt_surface.h"

struct RtHit { float t, u, v; uint32_t inst, prim; };

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

// (kx, ky, kz) is always a cyclic rotation of (0, 1, 2) (shear_init): v rotated so that .z is the kz component.  Bitwise
// selects on two masks: the chained ?: of comp() compiled to nine divergent branches per triangle in the cooperative
// round, where the 32 lanes test triangles against rays of different major axes.
RT_D f3 rotate_to_shear(f3 v, uint32_t m0, uint32_t m1) {      // m0 = all ones if kz == 0, m1 = all ones if kz == 1
    const uint32_t x = rt_float_as_uint(v.x), y = rt_float_as_uint(v.y), z = rt_float_as_uint(v.z);
    const uint32_t rx = (y & m0) | (((z & m1) | (x & ~m1)) & ~m0);       // kz 0: y   kz 1: z   kz 2: x
    const uint32_t ry = (z & m0) | (((x & m1) | (y & ~m1)) & ~m0);       //       z         x         y
    const uint32_t rz = (x & m0) | (((y & m1) | (z & ~m1)) & ~m0);       //       x         y         z
    return mk3(rt_uint_as_float(rx), rt_uint_as_float(ry), rt_uint_as_float(rz));
}
RT_D RayShear shear_init(f3 d) {
    RayShear r;
    int kz = 0; float m = fabsf(d.x);
    if (fabsf(d.y) > m) { kz = 1; m = fabsf(d.y); }
    if (fabsf(d.z) > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    const f3 ds = rotate_to_shear(d, kz == 0 ? 0xFFFFFFFFu : 0u, kz == 1 ? 0xFFFFFFFFu : 0u);    // (d[kx], d[ky], d[kz])
    const float dz = ds.z;
    // (the paper swaps kx/ky when d[kz] < 0 to preserve winding; for a two-sided test the swap negates U, V, W, det
    //  and T together and leaves t, u, v bit-identical, so it is omitted here and in the oracle)
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = rt_fdiv(ds.x, dz); r.Sy = rt_fdiv(ds.y, dz); r.Sz = rt_fdiv(1.0f, dz);
    return r;
}

// Woop/Benthin/Wald watertight test; every operation is an explicitly rounded IEEE op (no FMA contraction),
// so the result is bit-identical to the oracle's.
RT_D bool tri_test(const RayShear& r, f3 o, f3 v0, f3 v1, f3 v2, float tmin, float tmax, float& t, float& bu, float& bv) {
    const uint32_t m0 = r.kz == 0 ? 0xFFFFFFFFu : 0u, m1 = r.kz == 1 ? 0xFFFFFFFFu : 0u;
    const f3 A = rotate_to_shear(mk3(rt_fsub(v0.x, o.x), rt_fsub(v0.y, o.y), rt_fsub(v0.z, o.z)), m0, m1);
    const f3 B = rotate_to_shear(mk3(rt_fsub(v1.x, o.x), rt_fsub(v1.y, o.y), rt_fsub(v1.z, o.z)), m0, m1);
    const f3 C = rotate_to_shear(mk3(rt_fsub(v2.x, o.x), rt_fsub(v2.y, o.y), rt_fsub(v2.z, o.z)), m0, m1);
    const float Akz = A.z, Bkz = B.z, Ckz = C.z;
    const float Ax = rt_fsub(A.x, rt_fmul(r.Sx, Akz)), Ay = rt_fsub(A.y, rt_fmul(r.Sy, Akz));
    const float Bx = rt_fsub(B.x, rt_fmul(r.Sx, Bkz)), By = rt_fsub(B.y, rt_fmul(r.Sy, Bkz));
    const float Cx = rt_fsub(C.x, rt_fmul(r.Sx, Ckz)), Cy = rt_fsub(C.y, rt_fmul(r.Sy, Ckz));
    float U = rt_fsub(rt_fmul(Cx, By), rt_fmul(Cy, Bx));
    float V = rt_fsub(rt_fmul(Ax, Cy), rt_fmul(Ay, Cx));
    float W = rt_fsub(rt_fmul(Bx, Ay), rt_fmul(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)rt_dsub(rt_dmul((double)Cx, (double)By), rt_dmul((double)Cy, (double)Bx));
        V = (float)rt_dsub(rt_dmul((double)Ax, (double)Cy), rt_dmul((double)Ay, (double)Cx));
        W = (float)rt_dsub(rt_dmul((double)Bx, (double)Ay), rt_dmul((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = rt_fadd(rt_fadd(U, V), W);
    if (det == 0.0f) return false;
    const float Az = rt_fmul(r.Sz, Akz), Bz = rt_fmul(r.Sz, Bkz), Cz = rt_fmul(r.Sz, Ckz);
    const float T = rt_fadd(rt_fadd(rt_fmul(U, Az), rt_fmul(V, Bz)), rt_fmul(W, Cz));
    const float inv = rt_rcp(det);            // correctly rounded reciprocal == the oracle's 1.0f / det
    const float tt = rt_fmul(T, inv);
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; bu = rt_fmul(V, inv); bv = rt_fmul(W, inv);
    return true;
}

// row-major 3x4 transform with the oracle's operation order: ((m0*x + m1*y) + m2*z) + m3
RT_D f3 xform_point_exact(float4 r0, float4 r1, float4 r2, f3 p) {
    return mk3(rt_fadd(rt_fadd(rt_fadd(rt_fmul(r0.x, p.x), rt_fmul(r0.y, p.y)), rt_fmul(r0.z, p.z)), r0.w),
               rt_fadd(rt_fadd(rt_fadd(rt_fmul(r1.x, p.x), rt_fmul(r1.y, p.y)), rt_fmul(r1.z, p.z)), r1.w),
               rt_fadd(rt_fadd(rt_fadd(rt_fmul(r2.x, p.x), rt_fmul(r2.y, p.y)), rt_fmul(r2.z, p.z)), r2.w));
}
RT_D f3 xform_dir_exact(float4 r0, float4 r1, float4 r2, f3 d) {
    return mk3(rt_fadd(rt_fadd(rt_fmul(r0.x, d.x), rt_fmul(r0.y, d.y)), rt_fmul(r0.z, d.z)),
               rt_fadd(rt_fadd(rt_fmul(r1.x, d.x), rt_fmul(r1.y, d.y)), rt_fmul(r1.z, d.z)),
               rt_fadd(rt_fadd(rt_fmul(r2.x, d.x), rt_fmul(r2.y, d.y)), rt_fmul(r2.z, d.z)));
}

RT_D uint32_t byte_of(uint32_t v, int i) { return (v >> (8 * i)) & 0xFFu; }
RT_D float safe_rcp_dir(float d) { return rt_rcp_approx(fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }   // slab test is padded: 1-ulp rcp is fine
RT_D uint32_t octant_inv(f3 d) { return 7u - ((d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u)); }

// Intersects the 8 child boxes of one node.  Returns the hit mask: bits 24..31 inner children in traversal priority
// order (slot ^ octinv), bits 0..23 leaf primitives.
// Child planes are bfloat16 pairs packed as (hi << 16 | lo) relative to the node origin (rt_scene_dev.h).  The word read as
// a float is the hi plane (enlarged by < 1 bf16 ulp by the lo bits: conservative), word * 65536 is the lo plane; which of
// the two is the near plane depends on the ray direction only, so the choice is an integer multiply by a per-ray constant
// (1 or 65536): IMAD + FFMA on the fma pipe, no PRMT / SEL / I2F on the alu pipe, which the 8-bit layout saturated.
// Rounding: c and the products carry a few ulp of their magnitudes; the slab distances are padded by EPS times those
// magnitudes (2^(E-127) bounds the plane values), so the test never rejects a box the ray touches.
RT_D uint32_t node_intersect(const float4 n0, const float4 n1, const float4 x03, const float4 x47, const float4 y03, const float4 y47, const float4 z03, const float4 z47,
                             f3 o, f3 idir, uint32_t octinv, float tmin, float tmax) {
    const uint32_t n0w = rt_float_as_uint(n0.w);
    const float bound = rt_uint_as_float((n0w & 0xFFu) << 23);
    const float cx = (n0.x - o.x) * idir.x, cy = (n0.y - o.y) * idir.y, cz = (n0.z - o.z) * idir.z;
    const float EPS = 1.0e-6f;   // ~16 ulp of the magnitudes involved
    const float px = EPS * fmaf(bound, fabsf(idir.x), fabsf(cx)), py = EPS * fmaf(bound, fabsf(idir.y), fabsf(cy)), pz = EPS * fmaf(bound, fabsf(idir.z), fabsf(cz));
    const float cxn = cx - px, cxf = cx + px, cyn = cy - py, cyf = cy + py, czn = cz - pz, czf = cz + pz;
    // near / far plane multipliers from the sign bit of the direction: 1 (hi plane) / 65536 (lo plane).  Written as
    // arithmetic on the sign bit so that the products below stay integer multiply-adds on the fma pipe (a select between
    // the constants is strength-reduced to shifts on the alu pipe); recomputed per node: they cost less than the three
    // registers they would occupy in the 64-register kernel
    const uint32_t sx = rt_float_as_uint(idir.x) >> 31, sy = rt_float_as_uint(idir.y) >> 31, sz = rt_float_as_uint(idir.z) >> 31;
    uint32_t mnx = 65536u - sx * 65535u, mny = 65536u - sy * 65535u, mnz = 65536u - sz * 65535u;
    uint32_t mfx = 1u + sx * 65535u, mfy = 1u + sy * 65535u, mfz = 1u + sz * 65535u;
    rt_opaque(mnx); rt_opaque(mny); rt_opaque(mnz); rt_opaque(mfx); rt_opaque(mfy); rt_opaque(mfz);
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = rt_float_as_uint(half ? n1.w : n1.z);
        // per byte: inner children (low 5 bits >= 24) take priority slot ^ octinv, leaves keep their primitive offset
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xFFu;
        const uint32_t bit_index4 = (meta4 ^ ((octinv * 0x01010101u) & inner_mask4 & 0x07070707u)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const float4 xw = half ? x47 : x03, yw = half ? y47 : y03, zw = half ? z47 : z03;
        const uint32_t xs[4] = {rt_float_as_uint(xw.x), rt_float_as_uint(xw.y), rt_float_as_uint(xw.z), rt_float_as_uint(xw.w)};
        const uint32_t ys[4] = {rt_float_as_uint(yw.x), rt_float_as_uint(yw.y), rt_float_as_uint(yw.z), rt_float_as_uint(yw.w)};
        const uint32_t zs[4] = {rt_float_as_uint(zw.x), rt_float_as_uint(zw.y), rt_float_as_uint(zw.z), rt_float_as_uint(zw.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tnx = fmaf(rt_uint_as_float(xs[j] * mnx), idir.x, cxn), tfx = fmaf(rt_uint_as_float(xs[j] * mfx), idir.x, cxf);
            const float tny = fmaf(rt_uint_as_float(ys[j] * mny), idir.y, cyn), tfy = fmaf(rt_uint_as_float(ys[j] * mfy), idir.y, cyf);
            const float tnz = fmaf(rt_uint_as_float(zs[j] * mnz), idir.z, czn), tfz = fmaf(rt_uint_as_float(zs[j] * mfz), idir.z, czf);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            // child_bits << bit_index: one PRMT for the byte, a wrapping shift that only looks at the low 5 bits of its amount
            if (cmin <= cmax) hitmask |= rt_shl_wrap(rt_byte_of(child_bits4, j), bit_index4 >> (8 * j));
        }
    }
    return hitmask;
}

enum { RT_MODE_CLOSEST = 0, RT_MODE_ANY = 1 };
#ifdef RT_EMU_PROFILE
static unsigned long long g_emu_prof[8];   // developer probe (host emulation only): node visits per level, triangle tests per BLAS kind
extern "C" unsigned long long* emu_prof() { return g_emu_prof; }
#endif

// Resumable traversal state: init() then step() until it returns true.  The persistent extend kernel keeps one of
// these per lane and refills finished lanes with new rays (dynamic fetch); trace_ray() below simply loops.
struct Trav {
    f3 ow, dw; float tmin, tmax;          // world ray
    f3 o, idir; uint32_t octinv;          // current-level ray (world in the TLAS, object space inside a BLAS)
    RayShear sh;
    const float4* nodes; uint32_t tri_off;   // BLAS node array (rebased) and first triangle of the current BLAS
    int blas_sp;                          // stack height at BLAS entry; -1 = in the TLAS
    uint32_t cur_inst, cur_geo; bool cur_alpha;
    bool merged, identity;                // current BLAS is the merged world-space BLAS / has an identity transform
    uint2 ngroup, tgroup; int sp;
    RtHit hit; bool found;
    u4 rng;
};

RT_D void trav_set_level_ray(Trav& t, f3 o, f3 d) {
    t.o = o; t.idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z)); t.octinv = octant_inv(d);
}

template <bool SINGLE = false>
RT_D void trav_init(Trav& t, const DScene& S, f3 ow, f3 dw, float tmin, float tmax, u4 rng) {
    t.ow = ow; t.dw = dw; t.tmin = tmin; t.tmax = tmax; t.rng = rng;
    trav_set_level_ray(t, ow, dw);
    t.sh.kx = 0; t.sh.ky = 1; t.sh.kz = 2; t.sh.Sx = t.sh.Sy = t.sh.Sz = 0.0f;
    t.nodes = S.tlas_nodes; t.tri_off = 0; t.blas_sp = -1; t.cur_inst = 0; t.cur_geo = 0; t.cur_alpha = false; t.merged = false; t.identity = false;
    t.ngroup = make_uint2(0u, 0x80000000u); t.tgroup = make_uint2(0u, 0u); t.sp = 0;
    t.hit.t = tmax; t.hit.u = 0.0f; t.hit.v = 0.0f; t.hit.inst = 0xFFFFFFFFu; t.hit.prim = 0xFFFFFFFFu; t.found = false;
    if (SINGLE || S.single_merged || S.merged_first) {
        // the whole scene is the merged world-space BLAS, or it is traversed first: start inside it
        t.nodes = S.blas_nodes + (size_t)S.merged_node_off * RT_NODE_F4; t.tri_off = S.merged_tri_off;
        t.blas_sp = 0; t.identity = true; t.merged = true; t.cur_inst = S.n_instances;
        t.sh = shear_init(dw);
    }
}

// BLAS exhausted: back to world space (nothing to recompute after an identity instance)
RT_D void trav_leave_blas(Trav& t, const DScene& S) {
    t.blas_sp = -1; t.nodes = S.tlas_nodes;
    if (!t.identity) trav_set_level_ray(t, t.ow, t.dw);
    // DScene::merged_first: the merged BLAS was the start of the ray, not a TLAS entry; the TLAS root comes next
    if (t.merged && S.merged_first) { t.merged = false; t.ngroup = make_uint2(0u, 0x80000000u); }
}

// The traversal is a "while-while" loop (Aila & Laine 2009) over two kinds of steps:
//   node step : precondition tgroup empty.  Visits the next inner child of ngroup (one 128-byte node, 8 box tests) or,
//               when ngroup holds no more inner children, leaves the BLAS / pops the stack.  Returns true when the
//               ray has no work left.
//   prim step : precondition tgroup non-empty.  Takes one primitive of tgroup: a triangle (watertight test, alpha
//               test, commit) inside a BLAS, or an instance (ray transform, BLAS entry) in the TLAS.  Returns true
//               when an any-hit ray terminates.
// A warp runs all its lanes through node steps until every lane holds primitives, then through prim steps: both
// phases execute with most lanes active instead of interleaving per lane.
// MODE: closest / any (terminate on first accepted hit).  ALPHA: run the alpha test on non-opaque geometry
// (false == gl_RayFlagsOpaqueEXT / the reference's `fully_opaque` pipeline without any-hit shaders).
// visits the next inner child of ngroup (precondition: ngroup holds one): one 128-byte node, 8 box tests
template <bool COUNT, bool SINGLE = false>
RT_D void trav_visit_child(Trav& t, const DScene& S, uint2* stack, unsigned long long* c4) {
    const uint32_t hits = t.ngroup.y, imask = t.ngroup.y;
    const int child_bit = rt_bfind(hits);
    const uint32_t child_base = t.ngroup.x;
    t.ngroup.y &= ~(1u << child_bit);
    if (t.ngroup.y > 0x00FFFFFFu) stack[t.sp++] = t.ngroup;
    const uint32_t slot = (uint32_t)(child_bit - 24) ^ (t.octinv & 7u);
    const uint32_t rel = (uint32_t)rt_popc(imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
    const float4* np = (SINGLE ? S.blas_nodes + (size_t)S.merged_node_off * RT_NODE_F4 : t.nodes) + (size_t)(child_base + rel) * RT_NODE_F4;
    const float4 n0 = rt_ld(np), n1 = rt_ld(np + 1), n2 = rt_ld(np + 2), n3 = rt_ld(np + 3), n4 = rt_ld(np + 4), n5 = rt_ld(np + 5), n6 = rt_ld(np + 6), n7 = rt_ld(np + 7);
    if (COUNT) c4[0]++;
#ifdef RT_EMU_PROFILE
    g_emu_prof[t.blas_sp < 0 ? 0 : (t.merged ? 1 : 2)]++;
#endif
    const uint32_t hm = node_intersect(n0, n1, n2, n3, n4, n5, n6, n7, t.o, t.idir, t.octinv, t.tmin, t.hit.t);
    t.ngroup.x = rt_float_as_uint(n1.x); t.tgroup.x = rt_float_as_uint(n1.y);
    t.ngroup.y = (hm & 0xFF000000u) | (rt_float_as_uint(n0.w) >> 24);
    t.tgroup.y = hm & 0x00FFFFFFu;
}

template <bool COUNT, bool SINGLE = false>
RT_D bool trav_node_step(Trav& t, const DScene& S, uint2* stack, unsigned long long* c4) {
    if (t.ngroup.y > 0x00FFFFFFu) { trav_visit_child<COUNT, SINGLE>(t, S, stack, c4); return false; }
    if (t.blas_sp >= 0 && t.sp == t.blas_sp) {
        trav_leave_blas(t, S);
        if (t.ngroup.y > 0x00FFFFFFu) return false;        // merged-first scenes: the TLAS root
    }
    if (t.sp == 0) return true;
    const uint2 e = stack[--t.sp];
    if (e.y > 0x00FFFFFFu) { t.ngroup = e; } else { t.tgroup = e; t.ngroup = make_uint2(0u, 0u); }
    return false;
}

// TLAS leaf: enter the BLAS of the instance at bit `bit` of tgroup.  Remaining TLAS work goes on the stack first.
template <bool ALPHA, bool COUNT>
RT_D void trav_enter_instance(Trav& t, const DScene& S, uint2* stack, unsigned long long* c4) {
    const int bit = rt_bfind(t.tgroup.y);
    t.tgroup.y &= ~(1u << bit);
    const uint32_t inst = rt_ld(S.tlas_prims + t.tgroup.x + bit);
    if (t.tgroup.y) stack[t.sp++] = t.tgroup;
    if (t.ngroup.y > 0x00FFFFFFu) stack[t.sp++] = t.ngroup;
    const float4* ip = S.inst_w2o + (size_t)inst * RT_INST_F4;
    const float4 r0 = rt_ld(ip), r1 = rt_ld(ip + 1), r2 = rt_ld(ip + 2), meta = rt_ld(ip + 3);
    if (COUNT) c4[2]++;
#ifdef RT_EMU_PROFILE
    g_emu_prof[5]++;
#endif
    const uint32_t flags = rt_float_as_uint(meta.z);
    t.identity = (flags & RT_INST_IDENTITY) != 0u; t.merged = (flags & RT_INST_MERGED) != 0u;
    if (t.identity) {
        t.sh = shear_init(t.dw);              // world ray is used as is
    } else {
        const f3 od = xform_dir_exact(r0, r1, r2, t.dw);
        trav_set_level_ray(t, xform_point_exact(r0, r1, r2, t.ow), od);
        t.sh = shear_init(od);
    }
    t.cur_inst = inst; t.cur_geo = rt_float_as_uint(meta.y);
    t.cur_alpha = ALPHA && !(flags & RT_INST_OPAQUE);
    // node / primitive indices inside a BLAS are local to it: rebase the array pointers
    t.nodes = S.blas_nodes + (size_t)rt_float_as_uint(meta.x) * RT_NODE_F4;
    t.tri_off = rt_float_as_uint(meta.w);
    t.blas_sp = t.sp;
    // the root is entered through a virtual parent whose only inner child is node 0
    // (child_bit = 31, imask byte = 0 -> relative index 0)
    t.ngroup = make_uint2(0u, 0x80000000u); t.tgroup = make_uint2(0u, 0u);
}

// closest-hit acceptance of a geometric candidate of the current instance: lexicographic (t, instance, primitive)
RT_D bool trav_candidate_wins(const Trav& t, float tt, uint32_t inst, uint32_t prim) {
    if (!t.found) return true;
    if (tt > t.hit.t) return false;
    if (tt == t.hit.t && !(inst < t.hit.inst || (inst == t.hit.inst && prim < t.hit.prim))) return false;
    return true;
}
RT_D void trav_commit(Trav& t, float tt, float bu, float bv, uint32_t inst, uint32_t prim) {
    t.hit.t = tt; t.hit.u = bu; t.hit.v = bv; t.hit.inst = inst; t.hit.prim = prim; t.found = true;
}
// alpha-test context of a triangle: per instance record for the merged BLAS, else the current instance's
RT_D void trav_alpha_context(const DScene& S, bool merged, uint32_t inst, uint32_t cur_geo, bool cur_alpha, uint32_t& geo, bool& alpha) {
    geo = cur_geo; alpha = cur_alpha;
    if (merged) {
        const float4 meta = rt_ld(S.inst_w2o + (size_t)inst * RT_INST_F4 + 3);
        geo = rt_float_as_uint(meta.y); alpha = !(rt_float_as_uint(meta.z) & RT_INST_OPAQUE);
    }
}

template <int MODE, bool ALPHA, bool COUNT>
RT_D bool trav_prim_step(Trav& t, const DScene& S, uint2* stack, unsigned long long* c4) {
    if (t.blas_sp < 0) { trav_enter_instance<ALPHA, COUNT>(t, S, stack, c4); return false; }
    const int bit = rt_bfind(t.tgroup.y);
    t.tgroup.y &= ~(1u << bit);
    const float4* tp = S.tris + (size_t)(t.tri_off + t.tgroup.x + bit) * RT_TRI_F4;
    const float4 a = rt_ld(tp), b = rt_ld(tp + 1), c = rt_ld(tp + 2);
    if (COUNT) c4[1]++;
#ifdef RT_EMU_PROFILE
    g_emu_prof[t.merged ? 3 : 4]++;
#endif
    float tt, bu, bv;
    if (!tri_test(t.sh, t.o, xyz(a), xyz(b), xyz(c), t.tmin, t.tmax, tt, bu, bv)) return false;
    const uint32_t prim = rt_float_as_uint(a.w);
    const uint32_t inst = t.merged ? rt_float_as_uint(b.w) : t.cur_inst;
    if (!trav_candidate_wins(t, tt, inst, prim)) return false;
    if (ALPHA) {
        uint32_t geo; bool alpha;
        trav_alpha_context(S, t.merged, inst, t.cur_geo, t.cur_alpha, geo, alpha);
        if (alpha) {
            if (COUNT) c4[3]++;
            if (anyhit_ignore(S, inst, prim, geo, bu, bv, t.rng, COUNT ? c4 + 4 : nullptr)) return false;
        }
    }
    trav_commit(t, tt, bu, bv, inst, prim);
    return MODE == RT_MODE_ANY;
}

RT_D void trav_finish(Trav& t) { if (!t.found) t.hit.t = -1.0f; }

template <int MODE, bool ALPHA, bool COUNT>
RT_D bool trace_ray(const DScene& S, f3 ow, f3 dw, float tmin, float tmax, u4 rng, RtHit& hit, RtCounters* cnt) {
    uint2 stack[RT_STACK_SIZE];
    unsigned long long c4[5] = {0, 0, 0, 0, 0};
    Trav t;
    trav_init(t, S, ow, dw, tmin, tmax, rng);
    bool done = false;
    while (!done) {
        while (!done && t.tgroup.y == 0u) done = trav_node_step<COUNT>(t, S, stack, c4);
        while (!done && t.tgroup.y != 0u) done = trav_prim_step<MODE, ALPHA, COUNT>(t, S, stack, c4);
    }
    trav_finish(t);
    if (COUNT && cnt) {
        rt_atomic_add64(&cnt->nodes, c4[0]); rt_atomic_add64(&cnt->tris, c4[1]);
        rt_atomic_add64(&cnt->insts, c4[2]); rt_atomic_add64(&cnt->anyhits, c4[3]); rt_atomic_add64(&cnt->tex_taps, c4[4]);
    }
    hit = t.hit;
    return t.found;
}
