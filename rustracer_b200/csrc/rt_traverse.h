// rt_traverse.h — two-level traversal of the compressed 8-wide BVH + watertight ray/triangle test.
// Replaces what the reference delegates to the Vulkan driver: traceRayEXT at RayTracing.rgen:89-92 (closest
// hit, optional any-hit) and RayTracing.rchit:51-55 (shadow: terminate on first hit, skip closest hit).
// Semantics pinned by the oracle (oracle/oracle.cpp): open interval tmin < t < tmax, two-sided triangles,
// closest hit = lexicographic minimum of (t, instance, primitive), alpha test before commit.
#pragma once
#include "rt_surface.h"

struct RtHit { float t, u, v; uint32_t inst, prim; };

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

RT_D RayShear shear_init(f3 d) {
    RayShear r;
    int kz = 0; float m = fabsf(d.x);
    if (fabsf(d.y) > m) { kz = 1; m = fabsf(d.y); }
    if (fabsf(d.z) > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    const float dz = comp(d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = rt_fdiv(comp(d, kx), dz); r.Sy = rt_fdiv(comp(d, ky), dz); r.Sz = rt_fdiv(1.0f, dz);
    return r;
}

// Woop/Benthin/Wald watertight test; every operation is an explicitly rounded IEEE op (no FMA contraction),
// so the result is bit-identical to the oracle's.
RT_D bool tri_test(const RayShear& r, f3 o, f3 v0, f3 v1, f3 v2, float tmin, float tmax, float& t, float& bu, float& bv) {
    const f3 A = mk3(rt_fsub(v0.x, o.x), rt_fsub(v0.y, o.y), rt_fsub(v0.z, o.z));
    const f3 B = mk3(rt_fsub(v1.x, o.x), rt_fsub(v1.y, o.y), rt_fsub(v1.z, o.z));
    const f3 C = mk3(rt_fsub(v2.x, o.x), rt_fsub(v2.y, o.y), rt_fsub(v2.z, o.z));
    const float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    const float Ax = rt_fsub(comp(A, r.kx), rt_fmul(r.Sx, Akz)), Ay = rt_fsub(comp(A, r.ky), rt_fmul(r.Sy, Akz));
    const float Bx = rt_fsub(comp(B, r.kx), rt_fmul(r.Sx, Bkz)), By = rt_fsub(comp(B, r.ky), rt_fmul(r.Sy, Bkz));
    const float Cx = rt_fsub(comp(C, r.kx), rt_fmul(r.Sx, Ckz)), Cy = rt_fsub(comp(C, r.ky), rt_fmul(r.Sy, Ckz));
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
    const float tt = rt_fdiv(T, det);
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; bu = rt_fdiv(V, det); bv = rt_fdiv(W, det);
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
RT_D float safe_rcp_dir(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }
RT_D uint32_t octant_inv(f3 d) { return 7u - ((d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u)); }

// Intersects the 8 quantised child boxes of one node.  Returns the hit mask: bits 24..31 inner children in
// traversal priority order (slot ^ octinv), bits 0..23 leaf primitives.  Slab distances are padded by the
// worst-case rounding error of the de-quantisation so the test never rejects a box the ray touches.
RT_D uint32_t node_intersect(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4,
                             f3 o, f3 idir, uint32_t octinv, float tmin, float tmax) {
    const uint32_t n0w = rt_float_as_uint(n0.w);
    const float sx = rt_uint_as_float((((n0w >> 0) & 0xFFu)) << 23), sy = rt_uint_as_float((((n0w >> 8) & 0xFFu)) << 23), sz = rt_uint_as_float((((n0w >> 16) & 0xFFu)) << 23);
    const float ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
    const float ox = (n0.x - o.x) * idir.x, oy = (n0.y - o.y) * idir.y, oz = (n0.z - o.z) * idir.z;
    const float EPS = 1.0e-6f;   // ~16 ulp of the magnitudes involved
    const float px = EPS * (fabsf(ox) + 255.0f * fabsf(ax)), py = EPS * (fabsf(oy) + 255.0f * fabsf(ay)), pz = EPS * (fabsf(oz) + 255.0f * fabsf(az));
    const float oxl = ox - px, oxh = ox + px, oyl = oy - py, oyh = oy + py, ozl = oz - pz, ozh = oz + pz;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = rt_float_as_uint(half ? n1.w : n1.z);
        const uint32_t qlox = rt_float_as_uint(half ? n2.y : n2.x), qloy = rt_float_as_uint(half ? n2.w : n2.z), qloz = rt_float_as_uint(half ? n3.y : n3.x);
        const uint32_t qhix = rt_float_as_uint(half ? n3.w : n3.z), qhiy = rt_float_as_uint(half ? n4.y : n4.x), qhiz = rt_float_as_uint(half ? n4.w : n4.z);
        const uint32_t nx = idir.x < 0.0f ? qhix : qlox, fx = idir.x < 0.0f ? qlox : qhix;
        const uint32_t ny = idir.y < 0.0f ? qhiy : qloy, fy = idir.y < 0.0f ? qloy : qhiy;
        const uint32_t nz = idir.z < 0.0f ? qhiz : qloz, fz = idir.z < 0.0f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t meta = byte_of(meta4, j);
            const float tnx = fmaf((float)byte_of(nx, j), ax, oxl), tfx = fmaf((float)byte_of(fx, j), ax, oxh);
            const float tny = fmaf((float)byte_of(ny, j), ay, oyl), tfy = fmaf((float)byte_of(fy, j), ay, oyh);
            const float tnz = fmaf((float)byte_of(nz, j), az, ozl), tfz = fmaf((float)byte_of(fz, j), az, ozh);
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (cmin <= cmax) {
                const bool inner = (meta & 0x18u) == 0x18u;
                const uint32_t bit_index = (inner ? (meta ^ (octinv & 7u)) : meta) & 0x1Fu;
                hitmask |= (meta >> 5) << bit_index;
            }
        }
    }
    return hitmask;
}

enum { RT_MODE_CLOSEST = 0, RT_MODE_ANY = 1 };

// MODE: closest / any (terminate on first accepted hit).  ALPHA: run the alpha test on non-opaque geometry
// (false == gl_RayFlagsOpaqueEXT / the reference's `fully_opaque` pipeline without any-hit shaders).
template <int MODE, bool ALPHA, bool COUNT>
RT_D bool trace_ray(const DScene& S, f3 ow, f3 dw, float tmin, float tmax, u4 rng, RtHit& hit, RtCounters* cnt) {
    uint2 stack[RT_STACK_SIZE];
    int sp = 0;
    hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.inst = 0xFFFFFFFFu; hit.prim = 0xFFFFFFFFu;
    bool found = false;

    // current-level ray (world in the TLAS, object space inside a BLAS)
    f3 o = ow, d = dw;
    f3 idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    uint32_t octinv = octant_inv(d);
    RayShear sh; sh.kx = 0; sh.ky = 1; sh.kz = 2; sh.Sx = sh.Sy = sh.Sz = 0.0f;
    const float4* nodes = S.tlas_nodes;
    const float4* tris = S.tris;
    int blas_sp = -1;            // stack height at BLAS entry; -1 = in the TLAS
    uint32_t cur_inst = 0, cur_geo = 0; bool cur_alpha = false;
    unsigned long long c_nodes = 0, c_tris = 0, c_insts = 0, c_any = 0;

    uint2 ngroup = make_uint2(0u, 0x80000000u), tgroup = make_uint2(0u, 0u);
    for (;;) {
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t hits = ngroup.y, imask = ngroup.y;
            const int child_bit = rt_bfind(hits);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << child_bit);
            if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
            const uint32_t slot = (uint32_t)(child_bit - 24) ^ (octinv & 7u);
            const uint32_t rel = (uint32_t)rt_popc(imask & ~(0xFFFFFFFFu << slot) & 0xFFu);
            const float4* np = nodes + (size_t)(child_base + rel) * RT_NODE_F4;
            const float4 n0 = rt_ld(np), n1 = rt_ld(np + 1), n2 = rt_ld(np + 2), n3 = rt_ld(np + 3), n4 = rt_ld(np + 4);
            if (COUNT) c_nodes++;
            const uint32_t hm = node_intersect(n0, n1, n2, n3, n4, o, idir, octinv, tmin, hit.t);
            ngroup.x = rt_float_as_uint(n1.x); tgroup.x = rt_float_as_uint(n1.y);
            ngroup.y = (hm & 0xFF000000u) | (rt_float_as_uint(n0.w) >> 24);
            tgroup.y = hm & 0x00FFFFFFu;
        } else {
            tgroup = ngroup; ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0u) {
            const int bit = rt_bfind(tgroup.y);
            tgroup.y &= ~(1u << bit);
            if (blas_sp < 0) {
                // TLAS leaf: enter the instance's BLAS.  Remaining TLAS work goes on the stack first.
                const uint32_t inst = rt_ld(S.tlas_prims + tgroup.x + bit);
                if (tgroup.y) stack[sp++] = tgroup;
                if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
                const float4* ip = S.inst_w2o + (size_t)inst * RT_INST_F4;
                const float4 r0 = rt_ld(ip), r1 = rt_ld(ip + 1), r2 = rt_ld(ip + 2), meta = rt_ld(ip + 3);
                if (COUNT) c_insts++;
                o = xform_point_exact(r0, r1, r2, ow); d = xform_dir_exact(r0, r1, r2, dw);
                idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                octinv = octant_inv(d);
                sh = shear_init(d);
                cur_inst = inst; cur_geo = rt_float_as_uint(meta.y);
                cur_alpha = ALPHA && !(rt_float_as_uint(meta.z) & RT_INST_OPAQUE);
                // node / primitive indices inside a BLAS are local to it: rebase the array pointers
                nodes = S.blas_nodes + (size_t)rt_float_as_uint(meta.x) * RT_NODE_F4;
                tris = S.tris + (size_t)rt_float_as_uint(meta.w) * RT_TRI_F4;
                blas_sp = sp;
                // the root is entered through a virtual parent whose only inner child is node 0
                // (child_bit = 31, imask byte = 0 -> relative index 0)
                ngroup = make_uint2(0u, 0x80000000u); tgroup = make_uint2(0u, 0u);
                break;
            } else {
                const float4* tp = tris + (size_t)(tgroup.x + bit) * RT_TRI_F4;
                const float4 a = rt_ld(tp), b = rt_ld(tp + 1), c = rt_ld(tp + 2);
                if (COUNT) c_tris++;
                float t, bu, bv;
                if (!tri_test(sh, o, xyz(a), xyz(b), xyz(c), tmin, tmax, t, bu, bv)) continue;
                const uint32_t prim = rt_float_as_uint(a.w);
                if (found) {
                    if (t > hit.t) continue;
                    if (t == hit.t && !(cur_inst < hit.inst || (cur_inst == hit.inst && prim < hit.prim))) continue;
                }
                if (ALPHA && cur_alpha) {
                    if (COUNT) c_any++;
                    if (anyhit_ignore(S, cur_inst, prim, cur_geo, bu, bv, rng)) continue;
                }
                hit.t = t; hit.u = bu; hit.v = bv; hit.inst = cur_inst; hit.prim = prim; found = true;
                if (MODE == RT_MODE_ANY) goto done;
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (blas_sp >= 0 && sp == blas_sp) {
                // BLAS exhausted: back to world space
                blas_sp = -1; nodes = S.tlas_nodes; o = ow; d = dw;
                idir = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
                octinv = octant_inv(d);
            }
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
done:
    if (COUNT && cnt) {
        rt_atomic_add64(&cnt->nodes, c_nodes); rt_atomic_add64(&cnt->tris, c_tris);
        rt_atomic_add64(&cnt->insts, c_insts); rt_atomic_add64(&cnt->anyhits, c_any);
    }
    if (!found) hit.t = -1.0f;
    return found;
}
