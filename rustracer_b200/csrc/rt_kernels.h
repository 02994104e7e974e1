// rt_kernels.h — the wavefront stages as launchable kernels.
//   extend   : persistent warps fetch 32 rays at a time from the queue (dynamic load balance), while-while
//              traversal in rt_traverse.h, hit record out
//   shade    : one thread per path: miss / closest-hit shading, bounce epilogue, compaction into the next queue
//              (warp-aggregated atomics), shadow-ray emission
//   shadow   : any-hit traversal of the shadow queue, unoccluded contributions added to the frame radiance
//   raygen / accumulate / skinning / triangle packing / instance setup / refit
// Launch shapes: grids are multiples of the SM count (148 on B200); nothing here uses tensor cores — the path
// has no dense contraction.
#pragma once
#include "rt_shade.h"
#include "rt_build.h"

struct FrameBuffers {
    float4* acc;        // RGBA32F accumulation image (RayTracing.rgen:18)
    uint32_t* out;      // RGBA8 output image (RayTracing.rgen:19)
    float4* rad;        // radiance gathered during the current frame
    float2* aux;        // last Ray.t, number of traced segments (DISTANCE / HEAT mappings)
    uint4* pixrng;      // per-pixel stream positions carried between the samples of one frame
};
struct TilePart { uint32_t strip_rows, n_parts, part, width, height; };

RT_D uint32_t local_to_pixel(const TilePart& tp, uint32_t i) {
    if (tp.n_parts <= 1) return i;
    const uint32_t lr = i / tp.width, x = i % tp.width;
    const uint32_t k = lr / tp.strip_rows, r = lr % tp.strip_rows;
    const uint32_t y = (k * tp.n_parts + tp.part) * tp.strip_rows + r;
    return y * tp.width + x;
}

// Queue order of the primary rays: 8x4-pixel tiles instead of 32x1 row segments, so that the 32 rays a warp fetches together
// (and, through the order-preserving compaction, their descendants) start closer to each other.  A bijection of the local
// index space; paths carry their pixel, so nothing else depends on it.  n_rows = rows this context renders.
#ifndef RT_PRIMARY_TILES
#define RT_PRIMARY_TILES 1
#endif
RT_D uint32_t primary_order(uint32_t i, uint32_t width, uint32_t n_rows) {
#if RT_PRIMARY_TILES
    if ((width & 7u) == 0u && (n_rows & 3u) == 0u) {
        const uint32_t tiles_x = width >> 3, t = i >> 5, in = i & 31u;
        const uint32_t tx = t % tiles_x, ty = t / tiles_x;
        return (ty * 4u + (in >> 3)) * width + tx * 8u + (in & 7u);
    }
#endif
    return i;
}

RT_D void store_path(const DQueue& q, uint32_t slot, const PathState& s) {
    q.o_tmin[slot] = make_float4(s.origin.x, s.origin.y, s.origin.z, s.tmin);
    q.d_tmax[slot] = make_float4(s.dir.x, s.dir.y, s.dir.z, s.tmax);
    q.thr_pix[slot] = make_float4(s.throughput.x, s.throughput.y, s.throughput.z, rt_uint_as_float(s.pixel));
    q.rng[slot] = make_uint4(s.path_w, s.pix_w, s.lens_seed, rt_float_as_uint(s.volume_dis));
}
RT_D PathState load_path(const DQueue& q, uint32_t slot) {
    const float4 a = q.o_tmin[slot], b = q.d_tmax[slot], c = q.thr_pix[slot]; const uint4 r = q.rng[slot];
    PathState s;
    s.origin = mk3(a.x, a.y, a.z); s.tmin = a.w; s.dir = mk3(b.x, b.y, b.z); s.tmax = b.w;
    s.throughput = mk3(c.x, c.y, c.z); s.pixel = rt_float_as_uint(c.w);
    s.path_w = r.x; s.pix_w = r.y; s.lens_seed = r.z; s.volume_dis = rt_uint_as_float(r.w);
    return s;
}

// ---- per-item bodies ----------------------------------------------------------------------------------
RT_D void raygen_item(const FrameParams& P, const TilePart& tp, const FrameBuffers& fb, const DQueue& q, uint32_t i, uint32_t n_rows) {
    const uint32_t pixel = local_to_pixel(tp, primary_order(i, tp.width, n_rows));
    uint4 pr = make_uint4(0u, 0u, 0u, 0u);
    if (P.sample != 0) pr = fb.pixrng[pixel];
    else { fb.rad[pixel] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); fb.aux[pixel] = make_float2(0.0f, 0.0f); }
    const PathState s = raygen_path(P, pixel, pr.x, pr.y, pr.z);
    store_path(q, i, s);
}

template <bool ALPHA, bool COUNT>
RT_D void extend_item(const DScene& S, const FrameParams& P, const DQueue& q, const DHits& hits, uint32_t i, RtCounters* cnt) {
    const float4 a = q.o_tmin[i], b = q.d_tmax[i];
    u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
    if (ALPHA) { const uint32_t pixel = rt_float_as_uint(q.thr_pix[i].w); rng = path_stream(P, pixel, q.rng[i].x); }
    RtHit h;
    trace_ray<RT_MODE_CLOSEST, ALPHA, COUNT>(S, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, rng, h, cnt);
    hits.tuvp[i] = make_float4(h.t, h.u, h.v, rt_uint_as_float(h.prim));
    hits.inst[i] = h.inst;
}

// returns true when the path continues; slot allocation is done by the caller (warp-aggregated on the GPU)
struct ShadeResult { bool alive; PathState next; bool has_shadow; ShadowRay shadow; bool hit; };

template <bool SIMPLE, bool COUNT>
RT_D ShadeResult shade_item(const DScene& S, const FrameParams& P, const FrameBuffers& fb, const DQueue& q, const DHits& hits,
                            uint32_t i, uint32_t bounce, RtCounters* cnt) {
    PathState st = load_path(q, i);
    const float4 hv = hits.tuvp[i];
    RtHit h; h.t = hv.x; h.u = hv.y; h.v = hv.z; h.prim = rt_float_as_uint(hv.w); h.inst = hits.inst[i];
    ShadeOut so;
    if (h.t < 0.0f) shade_miss(S, P, st.dir, bounce == 0, so, cnt);
    else shade_hit<SIMPLE, COUNT>(S, P, h, st, so, cnt);

    ShadeResult r; r.alive = false; r.has_shadow = so.has_shadow; r.hit = !(h.t < 0.0f);
    if (so.has_shadow) { r.shadow = so.shadow; r.shadow.contrib = so.shadow.contrib * st.throughput; }
    // RayTracing.rgen:94-129
    const f3 add = st.throughput * so.emittance;
    if (add.x != 0.0f || add.y != 0.0f || add.z != 0.0f) {
        float4 cur = fb.rad[st.pixel];
        cur.x += add.x; cur.y += add.y; cur.z += add.z;
        fb.rad[st.pixel] = cur;
    }
    bool alive = (bounce + 1 != P.ubo.number_of_bounces);
    if (alive && bounce > 3) {   // MIN_BOUNCES 3
        const float p = clampf(luminance(st.throughput), 0.01f, 0.95f);
        u4 pix = pixel_stream(P, st.pixel, st.pix_w);
        const float qv = rng_next(pix); st.pix_w = pix.w;
        if (p < qv) alive = false; else st.throughput /= p;
    }
    if (alive) {
        st.throughput *= so.hit_value;
        if (!so.need_scatter || so.t < 0.0f) alive = false;
    }
    if (alive) {
        st.origin = so.next_origin; st.dir = so.next_dir; st.tmin = RT_TMIN;
        r.alive = true; r.next = st;
    } else {
        if (P.ubo.number_of_samples > 1) fb.pixrng[st.pixel] = make_uint4(st.pix_w, st.path_w, st.lens_seed, 0u);
        if (P.ubo.mapping == RT_MAP_DISTANCE || P.ubo.mapping == RT_MAP_HEAT) {
            float2 a = fb.aux[st.pixel]; a.x = so.t; a.y += (float)(bounce + 1); fb.aux[st.pixel] = a;
        }
    }
    return r;
}

template <bool ALPHA, bool COUNT>
RT_D void shadow_item(const DScene& S, const FrameParams& P, const FrameBuffers& fb, const DShadowQueue& sq, uint32_t i, RtCounters* cnt) {
    const float4 a = sq.o_tmax[i], b = sq.d_pix[i], c = sq.contrib[i];
    const uint32_t pixel = rt_float_as_uint(b.w);
    u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
    if (ALPHA) rng = path_stream(P, pixel, rt_float_as_uint(c.w));
    RtHit h;
    const bool occluded = trace_ray<RT_MODE_ANY, ALPHA, COUNT>(S, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), 0.1f, a.w, rng, h, cnt);   // tMin 0.1: RayTracing.rchit:43
    if (!occluded) {
        float4 cur = fb.rad[pixel];
        cur.x += c.x; cur.y += c.y; cur.z += c.z;
        fb.rad[pixel] = cur;
    }
}

RT_D void accumulate_item(const FrameParams& P, const TilePart& tp, const FrameBuffers& fb, uint32_t i, bool tonemap_only) {
    const uint32_t pixel = local_to_pixel(tp, i);
    if (tonemap_only) {
        // used after a cross-GPU reduce: acc already holds the sum for total_number_of_samples
        rt_ubo u = P.ubo; u.number_of_samples = 0; if (u.total_number_of_samples == 0) u.total_number_of_samples = 1;
        accumulate_pixel(u, fb.acc, fb.out, pixel, mk3(0.0f), 0.0f, 0u);
        return;
    }
    const float4 r = fb.rad[pixel]; const float2 a = fb.aux[pixel];
    accumulate_pixel(P.ubo, fb.acc, fb.out, pixel, mk3(r.x, r.y, r.z), a.x, (uint32_t)a.y);
}

// AnimationCompute.comp:14-39 on one vertex held as its 8 float4 (position, normal, tangent, color, weights, joints, uv0|uv1, skin_index|pad)
RT_D void skin_transform(float4 (&r)[8], const float* skins, uint32_t n_skins) {
    const int skin_index = (int)rt_float_as_uint(r[7].x);
    if (skin_index >= 0 && (uint32_t)skin_index < n_skins) {
        const float* bones = skins + (size_t)skin_index * (RT_MAX_JOINTS * 16);
        const float w[4] = {r[4].x, r[4].y, r[4].z, r[4].w};
        const uint32_t j[4] = {rt_float_as_uint(r[5].x), rt_float_as_uint(r[5].y), rt_float_as_uint(r[5].z), rt_float_as_uint(r[5].w)};
        // M = sum_k w_k * bones[joint_k], position / normal / tangent = M * v: every operation explicitly rounded, in the
        // oracle's order (oracle.cpp::skin_vertices, g++ -ffp-contract=off) -> the skinned vertices are bit-identical
        const float4* bm0 = reinterpret_cast<const float4*>(bones + (size_t)(j[0] & (RT_MAX_JOINTS - 1)) * 16);
        const float4* bm1 = reinterpret_cast<const float4*>(bones + (size_t)(j[1] & (RT_MAX_JOINTS - 1)) * 16);
        const float4* bm2 = reinterpret_cast<const float4*>(bones + (size_t)(j[2] & (RT_MAX_JOINTS - 1)) * 16);
        const float4* bm3 = reinterpret_cast<const float4*>(bones + (size_t)(j[3] & (RT_MAX_JOINTS - 1)) * 16);
        float M[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 c0 = rt_ld(bm0 + c), c1 = rt_ld(bm1 + c), c2 = rt_ld(bm2 + c), c3 = rt_ld(bm3 + c);
            M[c * 4 + 0] = rt_fadd(rt_fadd(rt_fadd(rt_fmul(w[0], c0.x), rt_fmul(w[1], c1.x)), rt_fmul(w[2], c2.x)), rt_fmul(w[3], c3.x));
            M[c * 4 + 1] = rt_fadd(rt_fadd(rt_fadd(rt_fmul(w[0], c0.y), rt_fmul(w[1], c1.y)), rt_fmul(w[2], c2.y)), rt_fmul(w[3], c3.y));
            M[c * 4 + 2] = rt_fadd(rt_fadd(rt_fadd(rt_fmul(w[0], c0.z), rt_fmul(w[1], c1.z)), rt_fmul(w[2], c2.z)), rt_fmul(w[3], c3.z));
            M[c * 4 + 3] = rt_fadd(rt_fadd(rt_fadd(rt_fmul(w[0], c0.w), rt_fmul(w[1], c1.w)), rt_fmul(w[2], c2.w)), rt_fmul(w[3], c3.w));
        }
        const f4 pos = mat4_mul_exact(M, mk4(r[0].x, r[0].y, r[0].z, 1.0f));
        const f3 nn = normalize_exact(xyz(mat4_mul_exact(M, mk4(r[1].x, r[1].y, r[1].z, 0.0f))));
        const f4 t4 = mat4_mul_exact(M, mk4(r[2].x, r[2].y, r[2].z, 0.0f));
        // normalize(vec4): the w component of M * (t, 0) takes part in the length (AnimationCompute.comp:36), w itself is restored
        const float tl = sqrtf(rt_fadd(rt_fadd(rt_fadd(rt_fmul(t4.x, t4.x), rt_fmul(t4.y, t4.y)), rt_fmul(t4.z, t4.z)), rt_fmul(t4.w, t4.w)));
        const f3 tt = mk3(rt_fdiv(t4.x, tl), rt_fdiv(t4.y, tl), rt_fdiv(t4.z, tl));
        r[0].x = pos.x; r[0].y = pos.y; r[0].z = pos.z;
        r[1].x = nn.x; r[1].y = nn.y; r[1].z = nn.z;
        r[2].x = tt.x; r[2].y = tt.y; r[2].z = tt.z;   // w (handedness) kept
    }
}
RT_D void skin_item(const rt_vertex* vin, rt_vertex* vout, const float* skins, uint32_t n_skins, uint32_t i) {
    const float4* src = reinterpret_cast<const float4*>(vin + i);
    float4 r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = src[k];
    skin_transform(r, skins, n_skins);
    float4* dst = reinterpret_cast<float4*>(vout + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = r[k];
}

RT_D DAabb tri_box_of(const rt_vertex* verts, const uint32_t* indices, uint32_t v_offset, uint32_t i_offset, uint32_t prim) {
    DAabb b;
    for (int k = 0; k < 3; ++k) {
        const float* p = verts[v_offset + indices[i_offset + 3 * prim + k]].position;
        for (int a = 0; a < 3; ++a) {
            if (k == 0) { b.lo[a] = p[a]; b.hi[a] = p[a]; }
            else { b.lo[a] = fminf(b.lo[a], p[a]); b.hi[a] = fmaxf(b.hi[a], p[a]); }
        }
    }
    return b;
}

#ifndef RT_EMU
// ---- CUDA kernels -------------------------------------------------------------------------------------
// Skinning as a streaming kernel: a CTA stages 128 vertices (16 KB) through shared memory so that every global access is
// a fully coalesced 16-byte-per-lane stream (one thread per vertex reading its own 128-byte record touches 32 cache
// lines per load instruction); each thread then transforms its vertex out of shared memory.  Rows are padded to 9
// float4 so the per-vertex accesses of a quarter warp fall into distinct banks.  The tiles are double-buffered: the
// asynchronous copies (cp.async, 16 B per lane, global -> shared without a register round trip) of the CTA's next tile are
// in flight while the current one is transformed and stored, so the load latency is no longer exposed once per tile
// (ncu, config 4: 54 % of the stall samples were long-scoreboard waits of the staging loads).
#define RT_SKIN_TILE 128
#define RT_SKIN_CTAS_PER_SM 6      // 2 x 18 KB of shared memory per CTA
RT_D void skin_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__global__ void __launch_bounds__(RT_SKIN_TILE) skin_kernel(const rt_vertex* vin, rt_vertex* vout, const float* skins, uint32_t n_skins, uint32_t n) {
    __shared__ float4 tile[2][RT_SKIN_TILE * 9];
    const float4* src = reinterpret_cast<const float4*>(vin); float4* dst = reinterpret_cast<float4*>(vout);
    const uint32_t stride = gridDim.x * RT_SKIN_TILE;
    auto stage = [&](uint32_t base, int buf) {           // issues this thread's copies of tile `base` (possibly none) as one group
        if (base < n) {
            const uint32_t cnt = n - base < RT_SKIN_TILE ? n - base : RT_SKIN_TILE;
            for (uint32_t f = threadIdx.x; f < cnt * 8u; f += RT_SKIN_TILE) skin_cp_async16(&tile[buf][(f >> 3) * 9u + (f & 7u)], src + (size_t)base * 8u + f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t base = blockIdx.x * RT_SKIN_TILE; int buf = 0;
    stage(base, 0);
    for (; base < n; base += stride, buf ^= 1) {
        stage(base + stride, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the newest group: the current tile has landed
        __syncthreads();
        const uint32_t cnt = n - base < RT_SKIN_TILE ? n - base : RT_SKIN_TILE;
        float4* t = tile[buf];
        if (threadIdx.x < cnt) {
            float4 r[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = t[threadIdx.x * 9u + k];
            skin_transform(r, skins, n_skins);
#pragma unroll
            for (int k = 0; k < 3; ++k) t[threadIdx.x * 9u + k] = r[k];      // only position / normal / tangent change
        }
        __syncthreads();
        for (uint32_t f = threadIdx.x; f < cnt * 8u; f += RT_SKIN_TILE) __stcs(dst + (size_t)base * 8u + f, t[(f >> 3) * 9u + (f & 7u)]);
        __syncthreads();                                          // the buffer is staged again at the top of the next iteration but one
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
#define RT_EXTEND_THREADS 128
#ifndef RT_EXTEND_MIN_BLOCKS
#define RT_EXTEND_MIN_BLOCKS 6   // <= 85 registers: 24 warps per SM
#endif
#ifndef RT_EXTEND_MIN_BLOCKS_SINGLE
#define RT_EXTEND_MIN_BLOCKS_SINGLE 8   // single-BLAS specialisation fits 64 registers without spills: 32 warps per SM
#endif

// Persistent traversal kernel body.
//  * per-lane dynamic fetch (Aila & Laine 2009; Ylitie et al. 2017): every lane owns a resumable traversal state;
//    finished lanes store their result and, once enough of the warp is idle, pull new rays from the queue with one
//    warp-aggregated atomic.
//  * every iteration each lane that holds a ray visits exactly one BVH8 node (popping / leaving a BLAS first if it
//    has to), so the 8-box test runs with most of the warp active.
//  * deferred, warp-cooperative triangle tests: leaf sizes vary from 1 to 24 triangles, which left ~3 of 32 lanes
//    busy when each lane tested its own leaves (profiles/r01).  Instead a lane appends (owner lane, triangle) items
//    to a per-warp queue in shared memory and keeps traversing with its (conservatively stale) hit distance; whenever
//    32 items are queued the warp tests them one per lane against the owners' rays (kept in shared memory) and
//    resolves each owner's winner (min t, then min primitive id) with shared-memory atomics.  A lane drains its
//    outstanding items before it leaves a BLAS or retires its ray.
// MODE selects closest-hit (extend) or any-hit (shadow) semantics.
#ifndef RT_REFILL_BELOW
#define RT_REFILL_BELOW 16     // refill when fewer than this many lanes still hold a ray (sweep 6..32 on config 2: flat optimum at 14-18; a refill costs ~150 instructions at few lanes)
#endif
#define RT_WARPS_PER_BLOCK (RT_EXTEND_THREADS / 32)
#ifndef RT_NODE_STEPS
#define RT_NODE_STEPS 2         // node visits per lane between two triangle rounds: fuller rounds (+1.5 % on config 2)
#endif
#ifndef RT_TQ_PUSH_MAX
#define RT_TQ_PUSH_MAX 7        // triangles a lane may queue per iteration (the rest stay parked in its tgroup)
#endif
#ifndef RT_TQ_CAP
#define RT_TQ_CAP 512u         // per-warp triangle queue capacity (power of two, >= 31 + 32 * RT_TQ_PUSH_MAX * RT_NODE_STEPS)
#endif
#define RT_TQ_TRI_BITS 27      // item = owner lane << 27 | absolute triangle index
#ifndef RT_ENTER_BATCH
#define RT_ENTER_BATCH 8        // two-level scenes: instance entries wait for this many lanes ...
#endif
#ifndef RT_ENTER_MAX_WAIT
#define RT_ENTER_MAX_WAIT 3     // ... or this many node steps
#endif

#ifdef RT_PROBE
// developer probe (variant builds only, scripts/gpu_probe.py): per-warp start / queue-exhausted / exit times, rays, iterations
__device__ unsigned long long g_probe[8192 * 6];
RT_D unsigned long long rt_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif
template <bool ALPHA>
struct CoopShared {            // one per warp; SoA over the 32 owner lanes
    float ox[32], oy[32], oz[32], Sx[32], Sy[32], Sz[32], tmin[32], tmax[32], cur_t[32], u[32], v[32];
    unsigned long long best_ip[32];          // (ordered t << 32 | rank) of the best candidate of this round
    unsigned long long win_ip[32];           // its (instance << 32 | primitive)
    uint32_t kxyz[32], done[32];
    uint32_t inst[32];                       // 0xFFFFFFFF = merged BLAS (instance id in the triangle record)
    uint32_t geo[ALPHA ? 32 : 1], alpha[ALPHA ? 32 : 1], rng[4][ALPHA ? 32 : 1];   // any-hit context, only kept when alpha tests can run
    uint32_t items[RT_TQ_CAP];
    uint32_t tail;                           // total items ever appended
};

template <bool ALPHA, bool SINGLE>
RT_D void coop_publish_ray(const Trav& tv, CoopShared<ALPHA>& sh, uint32_t lane) {
    sh.ox[lane] = tv.o.x; sh.oy[lane] = tv.o.y; sh.oz[lane] = tv.o.z;
    sh.Sx[lane] = tv.sh.Sx; sh.Sy[lane] = tv.sh.Sy; sh.Sz[lane] = tv.sh.Sz;
    sh.tmin[lane] = tv.tmin; sh.tmax[lane] = tv.tmax; sh.cur_t[lane] = tv.found ? tv.hit.t : tv.tmax;
    sh.kxyz[lane] = (uint32_t)tv.sh.kx | ((uint32_t)tv.sh.ky << 2) | ((uint32_t)tv.sh.kz << 4);
    if (!SINGLE) sh.inst[lane] = tv.merged ? 0xFFFFFFFFu : tv.cur_inst;
    if (ALPHA) {
        sh.geo[lane] = tv.cur_geo; sh.alpha[lane] = tv.cur_alpha ? 1u : 0u;
        sh.rng[0][lane] = tv.rng.x; sh.rng[1][lane] = tv.rng.y; sh.rng[2][lane] = tv.rng.z; sh.rng[3][lane] = tv.rng.w;
    }
}

// Tests up to 32 queued (owner, triangle) items, one per lane.  Returns through `outstanding` / tv the owner-side
// bookkeeping.  `usable`: this lane's ray is still live (results of a retired any-hit ray are discarded).
template <int MODE, bool ALPHA, bool COUNT, bool SINGLE>
RT_D void coop_round(Trav& tv, const DScene& S, CoopShared<ALPHA>& sh, uint32_t head, uint32_t n, uint32_t lane, uint32_t& outstanding, bool usable, bool& terminated,
                     unsigned long long* c4) {
    // winner per owner = lexicographic minimum of (t, instance, primitive).  One 64-bit shared-memory atomicMin on
    // (ordered t << 32 | rank) decides it: rank (triangle record .w) orders the triangles of a BLAS by (instance, primitive)
    sh.best_ip[lane] = ~0ull; sh.done[lane] = 0u;
    __syncwarp();
    bool hit = false; float tt = 0.0f, bu = 0.0f, bv = 0.0f; uint32_t owner = 0; unsigned long long key = 0, ip = 0;
    if (lane < n) {
        const uint32_t item = sh.items[(head + lane) & (RT_TQ_CAP - 1u)];
        owner = item >> RT_TQ_TRI_BITS;
        const float4* tp = S.tris + (size_t)(item & ((1u << RT_TQ_TRI_BITS) - 1u)) * RT_TRI_F4;
        const float4 a = rt_ld(tp), b = rt_ld(tp + 1), c = rt_ld(tp + 2);
        if (COUNT) c4[1]++;
        RayShear rs; const uint32_t kk = sh.kxyz[owner];
        rs.kx = (int)(kk & 3u); rs.ky = (int)((kk >> 2) & 3u); rs.kz = (int)((kk >> 4) & 3u);
        rs.Sx = sh.Sx[owner]; rs.Sy = sh.Sy[owner]; rs.Sz = sh.Sz[owner];
        hit = tri_test(rs, mk3(sh.ox[owner], sh.oy[owner], sh.oz[owner]), xyz(a), xyz(b), xyz(c), sh.tmin[owner], sh.tmax[owner], tt, bu, bv);
        const uint32_t prim = rt_float_as_uint(a.w);
        const bool merged = SINGLE || sh.inst[owner] == 0xFFFFFFFFu;
        const uint32_t inst = merged ? rt_float_as_uint(b.w) : sh.inst[owner];
        ip = ((unsigned long long)inst << 32) | prim;
        if (hit && tt > sh.cur_t[owner]) hit = false;            // cannot beat the owner's committed hit
        if (ALPHA && hit) {
            uint32_t geo; bool alpha;
            trav_alpha_context(S, merged, inst, sh.geo[owner], sh.alpha[owner] != 0u, geo, alpha);
            if (alpha) {
                if (COUNT) c4[3]++;
                u4 rng; rng.x = sh.rng[0][owner]; rng.y = sh.rng[1][owner]; rng.z = sh.rng[2][owner]; rng.w = sh.rng[3][owner];
                if (anyhit_ignore(S, inst, prim, geo, bu, bv, rng, COUNT ? c4 + 4 : nullptr)) hit = false;
            }
        }
        if (hit) { key = ((unsigned long long)float_to_ordered(tt) << 32) | rt_float_as_uint(c.w); atomicMin(&sh.best_ip[owner], key); }
        atomicAdd(&sh.done[owner], 1u);
    }
    __syncwarp();
    if (hit && key == sh.best_ip[owner]) { sh.u[owner] = bu; sh.v[owner] = bv; sh.win_ip[owner] = ip; }
    __syncwarp();
    const uint32_t d = sh.done[lane];
    if (d) {
        outstanding -= d;
        if (usable && sh.best_ip[lane] != ~0ull) {
            const float ct = ordered_to_float((uint32_t)(sh.best_ip[lane] >> 32));
            const uint32_t ci = (uint32_t)(sh.win_ip[lane] >> 32), cp = (uint32_t)sh.win_ip[lane];
            if (trav_candidate_wins(tv, ct, ci, cp)) {
                trav_commit(tv, ct, sh.u[lane], sh.v[lane], ci, cp);
                sh.cur_t[lane] = ct;
                if (MODE == RT_MODE_ANY) terminated = true;
            }
        }
    }
    __syncwarp();
}

// SINGLE: the scene is one merged world-space BLAS (DScene::single_merged) — no TLAS level, no instance entry / exit,
// node and triangle bases come from the kernel parameters (uniform registers) instead of per-lane state.
template <int MODE, bool ALPHA, bool COUNT, bool SINGLE, class LoadRay, class StoreHit>
RT_D void persistent_trace(const DScene& S, const uint32_t count, uint32_t* fetch, RtCounters* cnt, LoadRay load_ray, StoreHit store_hit, uint32_t min_rays_per_cta = 0u) {
    // thin queues (late bounces): the launch is sized for the largest queue; CTAs beyond count / min_rays_per_cta leave at
    // once so that the remaining warps are refilled several times instead of every warp draining a single batch of rays
    if (min_rays_per_cta && (unsigned long long)blockIdx.x * min_rays_per_cta >= count && blockIdx.x != 0u) return;
    __shared__ CoopShared<ALPHA> coop_smem[RT_WARPS_PER_BLOCK];
    CoopShared<ALPHA>& sh = coop_smem[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint2 stack[RT_STACK_SIZE];
    unsigned long long c4[5] = {0, 0, 0, 0, 0};
    Trav tv;
    tv.tgroup = make_uint2(0u, 0u); tv.ngroup = make_uint2(0u, 0u); tv.sp = 0; tv.blas_sp = -1; tv.found = false;
    bool active = false, exhausted = false;
    uint32_t idx = 0, outstanding = 0;          // outstanding: this lane's items still in the queue
    uint32_t q_head = 0, q_count = 0;           // warp-uniform queue cursor
    int enter_wait = 0;                         // two-level scenes: steps the pending instance entries have been postponed (warp-uniform)
    if (lane == 0) sh.tail = 0u;
    __syncwarp();
#ifdef RT_PROBE
    const unsigned long long probe_t0 = rt_globaltimer(); unsigned long long probe_tx = 0; uint32_t probe_rays = 0, probe_iters = 0;
#endif
    for (;;) {
        if (!exhausted) {
            const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && outstanding == 0u);
            if (need) {
                const int leader = __ffs((int)need) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(need));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (!active && outstanding == 0u) {
                    idx = base + (uint32_t)__popc(need & lt_mask);
                    if (idx < count) { load_ray(idx, tv); active = true; if (SINGLE || tv.blas_sp >= 0) coop_publish_ray<ALPHA, SINGLE>(tv, sh, lane); }
                }
                if (base + (uint32_t)__popc(need) >= count) exhausted = true;
#ifdef RT_PROBE
                probe_rays += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, active && idx >= base && idx < count && idx - base < 32u));
                if (exhausted) probe_tx = rt_globaltimer();
#endif
            }
        }
        if (!__ballot_sync(0xFFFFFFFFu, active || outstanding != 0u)) break;
        uint32_t holding;
        do {
            bool want_flush = false;
#pragma unroll 1
            for (int rep = 0; rep < RT_NODE_STEPS; ++rep) {
            uint32_t leaf_mask = 0u, leaf_base = 0u;
            if (active && !want_flush) {
                // acquire the next node group: leave the BLAS / pop until ngroup holds an inner child or instances are parked
                while (tv.ngroup.y <= 0x00FFFFFFu && tv.tgroup.y == 0u) {
                    if (!SINGLE && tv.blas_sp >= 0 && tv.sp == tv.blas_sp) {
                        // queued triangles refer to the object-space ray; after an identity BLAS the published ray stays
                        // valid, the flush is postponed to the next instance entry
                        if (outstanding && !tv.identity) { want_flush = true; break; }
                        trav_leave_blas(tv, S);
                        if (tv.ngroup.y > 0x00FFFFFFu) break;              // merged-first scenes: the TLAS root is next
                    }
                    if (tv.sp == 0) {
                        if (outstanding) { want_flush = true; break; }
                        trav_finish(tv); store_hit(idx, tv); active = false; break;
                    }
                    const uint2 e = stack[--tv.sp];
                    if (e.y > 0x00FFFFFFu) tv.ngroup = e; else { tv.tgroup = e; tv.ngroup = make_uint2(0u, 0u); }
                }
            }
            // Two-level scenes: entering an instance (ray transform, shear setup, record fetch: ~180 instructions) is per-lane
            // work that few lanes reach in the same step — alone it ran at 2-3 of 32 lanes and cost 13 % of the kernel's
            // instructions (config 3, profiles/r02_extend_config3_ncu.txt).  Lanes that reach a TLAS leaf therefore wait until
            // RT_ENTER_BATCH of them can enter together, nobody else has a node to visit, or RT_ENTER_MAX_WAIT steps have
            // passed: a waiting lane only forgoes its share of a node step (~12 warp instructions).
            bool do_work = active && !want_flush, enter = false;
            if constexpr (!SINGLE) {
                enter = do_work && tv.tgroup.y != 0u && tv.blas_sp < 0;
                if (enter && outstanding) { want_flush = true; do_work = false; enter = false; }   // entering republishes the lane's ray
                const uint32_t em = __ballot_sync(0xFFFFFFFFu, enter);
                if (em) {
                    const uint32_t others = __ballot_sync(0xFFFFFFFFu, do_work && !enter);
                    if (__popc(em) >= RT_ENTER_BATCH || others == 0u || ++enter_wait >= RT_ENTER_MAX_WAIT) enter_wait = 0;
                    else { if (enter) do_work = false; enter = false; }
                }
            }
            {
                if (do_work) {
                    if (!SINGLE && enter) {
                        // TLAS level: parked instances
                        trav_enter_instance<ALPHA, COUNT>(tv, S, stack, c4);
                        coop_publish_ray<ALPHA, SINGLE>(tv, sh, lane);
                    } else {
                        // BLAS level with leftover leaf triangles parked: queue those first; else visit the next node
                        // (the acquire loop above left an inner child in ngroup unless primitives are parked)
                        if (tv.tgroup.y == 0u) trav_visit_child<COUNT, SINGLE>(tv, S, stack, c4);
                        if ((SINGLE || tv.blas_sp >= 0) && tv.tgroup.y != 0u) {
                            leaf_base = (SINGLE ? S.merged_tri_off : tv.tri_off) + tv.tgroup.x;
                            leaf_mask = tv.tgroup.y;
                            if (__popc(leaf_mask) > RT_TQ_PUSH_MAX) {     // keep the RT_TQ_PUSH_MAX highest bits, park the rest
                                uint32_t keep = 0u, m = leaf_mask;
#pragma unroll
                                for (int q = 0; q < RT_TQ_PUSH_MAX; ++q) { const uint32_t b = 1u << (31 - __clz((int)m)); keep |= b; m &= ~b; }
                                leaf_mask = keep;
                            }
                            tv.tgroup.y &= ~leaf_mask;
                        }
                    }
                }
            }
            // append this iteration's leaf triangles to the warp queue
            const uint32_t k = (uint32_t)__popc(leaf_mask);
            // queue positions from one shared-memory atomic per pushing lane (item order inside the queue is irrelevant:
            // hit resolution is order-independent) instead of a 5-step shuffle scan
            if (k) {
                uint32_t slot = atomicAdd(&sh.tail, k) & (RT_TQ_CAP - 1u);
                outstanding += k;
                // (this loop runs with few lanes: keep it short — owner | first triangle folded into one base, a wrapped slot
                //  counter instead of masking a position; leaf_base + bit < 2^27 cannot carry into the owner bits)
                const uint32_t item_base = (lane << RT_TQ_TRI_BITS) | leaf_base;
                uint32_t off = slot * 4u;                                 // byte offset into the ring
                do {
                    uint32_t bit; asm("bfind.u32 %0, %1;" : "=r"(bit) : "r"(leaf_mask));
                    leaf_mask ^= 1u << bit;
                    *reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(sh.items) + off) = item_base + bit;
                    off = (off + 4u) & (RT_TQ_CAP * 4u - 1u);
                } while (leaf_mask);
            }
            __syncwarp();
            q_count = sh.tail - q_head;
            if (q_count + 32u * RT_TQ_PUSH_MAX + 32u > RT_TQ_CAP) break;     // (warp-uniform) no room for another round of pushes
            }   // rep
            const bool flush = __any_sync(0xFFFFFFFFu, want_flush);
            __syncwarp();
            while (q_count >= 32u || (flush && q_count)) {
                const uint32_t n = q_count < 32u ? q_count : 32u;
                bool terminated = false;
                coop_round<MODE, ALPHA, COUNT, SINGLE>(tv, S, sh, q_head, n, lane, outstanding, active, terminated, c4);
                if (MODE == RT_MODE_ANY && terminated && active) { trav_finish(tv); store_hit(idx, tv); active = false; }
                q_head += n; q_count -= n;
            }
            // a lane that only waited for its last triangles can retire now instead of spending another iteration
            if (want_flush && active && outstanding == 0u && tv.sp == 0 && tv.ngroup.y <= 0x00FFFFFFu && tv.tgroup.y == 0u &&
                (SINGLE || tv.blas_sp <= 0)) {
                trav_finish(tv); store_hit(idx, tv); active = false;
            }
            holding = __ballot_sync(0xFFFFFFFFu, active);
#ifdef RT_PROBE
            ++probe_iters;
#endif
        } while (holding && (exhausted || __popc(holding) >= RT_REFILL_BELOW));
        // lanes whose any-hit ray retired early may still own queued items: drain so they can be refilled
        // (a closest-hit lane only retires with an empty queue share: no second copy of the round in those kernels)
        if constexpr (MODE == RT_MODE_ANY) if (__any_sync(0xFFFFFFFFu, !active && outstanding != 0u)) {
            while (q_count) {
                const uint32_t n = q_count < 32u ? q_count : 32u;
                bool terminated = false;
                coop_round<MODE, ALPHA, COUNT, SINGLE>(tv, S, sh, q_head, n, lane, outstanding, active, terminated, c4);
                if (MODE == RT_MODE_ANY && terminated && active) { trav_finish(tv); store_hit(idx, tv); active = false; }
                q_head += n; q_count -= n;
            }
        }
    }
#ifdef RT_PROBE
    if (lane == 0) {
        const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (wid < 8192u) {
            unsigned long long* p = g_probe + (size_t)wid * 6;
            p[0] = probe_t0; p[1] = probe_tx; p[2] = rt_globaltimer(); p[3] = probe_rays; p[4] = probe_iters;
            uint32_t smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); p[5] = smid;
        }
    }
#endif
    if (COUNT && cnt) {
        atomicAdd(&cnt->nodes, c4[0]); atomicAdd(&cnt->tris, c4[1]); atomicAdd(&cnt->insts, c4[2]); atomicAdd(&cnt->anyhits, c4[3]); atomicAdd(&cnt->tex_taps, c4[4]);
    }
}

template <bool ALPHA, bool COUNT, bool SINGLE>
__global__ void __launch_bounds__(RT_EXTEND_THREADS, SINGLE ? RT_EXTEND_MIN_BLOCKS_SINGLE : RT_EXTEND_MIN_BLOCKS) extend_kernel(DScene S, FrameParams P, DQueue q, DHits hits, const uint32_t* count_ptr, uint32_t* fetch, RtCounters* cnt, uint32_t min_rays_per_cta) {
    persistent_trace<RT_MODE_CLOSEST, ALPHA, COUNT, SINGLE>(S, *count_ptr, fetch, cnt,
        [&](uint32_t i, Trav& tv) {
            const float4 a = q.o_tmin[i], b = q.d_tmax[i];
            u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
            if (ALPHA) { const uint32_t pixel = rt_float_as_uint(q.thr_pix[i].w); rng = path_stream(P, pixel, q.rng[i].x); }
            trav_init<SINGLE>(tv, S, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, rng);
        },
        [&](uint32_t i, const Trav& tv) {
            hits.tuvp[i] = make_float4(tv.hit.t, tv.hit.u, tv.hit.v, rt_uint_as_float(tv.hit.prim));
            hits.inst[i] = tv.hit.inst;
        }, min_rays_per_cta);
}

template <bool ALPHA, bool COUNT, bool SINGLE>
__global__ void __launch_bounds__(RT_EXTEND_THREADS, SINGLE ? RT_EXTEND_MIN_BLOCKS_SINGLE : RT_EXTEND_MIN_BLOCKS) shadow_kernel(DScene S, FrameParams P, FrameBuffers fb, DShadowQueue sq, const uint32_t* count_ptr, uint32_t* fetch, RtCounters* cnt, uint32_t min_rays_per_cta) {
    persistent_trace<RT_MODE_ANY, ALPHA, COUNT, SINGLE>(S, *count_ptr, fetch, cnt,
        [&](uint32_t i, Trav& tv) {
            const float4 a = sq.o_tmax[i], b = sq.d_pix[i];
            u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
            if (ALPHA) rng = path_stream(P, rt_float_as_uint(b.w), rt_float_as_uint(sq.contrib[i].w));
            trav_init<SINGLE>(tv, S, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), 0.1f, a.w, rng);   // tMin 0.1: RayTracing.rchit:43
        },
        [&](uint32_t i, const Trav& tv) {
            if (!tv.found) {   // unoccluded: add the light's contribution to the frame radiance of the pixel
                const float4 c = sq.contrib[i]; const uint32_t pixel = rt_float_as_uint(sq.d_pix[i].w);
                float4 cur = fb.rad[pixel];
                cur.x += c.x; cur.y += c.y; cur.z += c.z;
                fb.rad[pixel] = cur;
            }
        }, min_rays_per_cta);
}

#ifndef RT_SHADE_PREFETCH
#define RT_SHADE_PREFETCH 1   // 1 = L2, 2 = L1
#endif
#if RT_SHADE_PREFETCH == 2
RT_D void rt_prefetch(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
RT_D void rt_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS 8
#endif
#ifndef RT_SHADE_SORT
#define RT_SHADE_SORT 1       // 1: material-sorted shading (CTA-local counting sort of each 128-path tile), 0: queue order
#endif
#ifndef RT_SHADE_OCTSORT
#define RT_SHADE_OCTSORT 1    // next-bounce queue grouped by direction class inside each 128-path tile (1: octant, 2: octant x major axis, 3: two sign bits; 0: off)
#endif
#define RT_OCT_CLASSES (RT_SHADE_OCTSORT == 2 ? 24u : (RT_SHADE_OCTSORT == 3 ? 4u : 8u))
#define RT_SHADE_CLASSES 32   // class 0 = miss, 1..30 = material id (mod 30), 31 = no path (tail of the last tile)

// Material-sorted closest-hit shading (north_star: "a material-sorted ... closest-hit shading pass").  The reference's
// ubershader (RayTracing.rchit:136-477) branches on workflow / transmission / volume / texture bits per material; in queue
// order a warp holds a random mix of them (18-19 of 32 lanes active after the first bounce, profiles/r02).  Each CTA
// therefore takes a tile of 128 consecutive paths, reads only what the sort key needs (hit distance + instance ->
// material id, two coalesced loads and one L1-resident gather), counting-sorts the tile by (miss | material) in shared
// memory — warp-level match + a 128-entry block scan, ~70 instructions per path against ~1500 of shading — and thread t
// then shades the t-th path of the sorted order.  Paths are independent, so the image is bit-identical to queue order;
// only the order of the compacted next-bounce queue changes.
template <bool SIMPLE, bool COUNT>
__global__ void __launch_bounds__(128, RT_SHADE_MIN_BLOCKS) shade_kernel(DScene S, FrameParams P, FrameBuffers fb, DQueue qin, DHits hits, DQueue qout, DShadowQueue sq,
                                                    const uint32_t* count_ptr, uint32_t* out_count, uint32_t* shadow_count, uint32_t* hit_count, uint32_t bounce, RtCounters* cnt) {
    const uint32_t count = *count_ptr;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#if RT_SHADE_SORT
    __shared__ uint32_t s_hist[RT_SHADE_CLASSES * 4];     // [class][warp] counts, then exclusive offsets
    __shared__ uint32_t s_wsum[4];
    __shared__ uint8_t s_perm[128];
#endif
    // all threads of a CTA iterate together (barriers / ballots below are convergent)
    for (uint32_t tile = blockIdx.x * 128u; tile < count; tile += gridDim.x * 128u) {
        uint32_t i = tile + threadIdx.x;
#if RT_SHADE_PREFETCH
        {   // stream the next tile's path state and hit records towards the SM while this one is shaded
            const uint32_t nx = i + gridDim.x * 128u;
            if (nx < count) {
                rt_prefetch(qin.o_tmin + nx); rt_prefetch(qin.d_tmax + nx); rt_prefetch(qin.thr_pix + nx); rt_prefetch(qin.rng + nx);
                rt_prefetch(hits.tuvp + nx); rt_prefetch(hits.inst + nx);
            }
        }
#endif
#if RT_SHADE_SORT
        {
            uint32_t cls = RT_SHADE_CLASSES - 1u;
            if (i < count) {
                const float t = hits.tuvp[i].x;
                cls = 0u;
                if (!(t < 0.0f)) cls = 1u + rt_float_as_uint(rt_ld(S.inst_o2w + (size_t)hits.inst[i] * RT_O2W_F4 + 3).z) % (RT_SHADE_CLASSES - 2u);
            }
            s_hist[threadIdx.x] = 0u;
            __syncthreads();
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, cls);
            const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (rank == 0u) s_hist[cls * 4u + warp] = (uint32_t)__popc(peers);
            __syncthreads();
            // exclusive scan of the 128 (class-major, warp-minor) counts: one entry per thread
            const uint32_t v = s_hist[threadIdx.x];
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
            if (lane == 31u) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t base = 0u;
            for (uint32_t w = 0; w < warp; ++w) base += s_wsum[w];
            s_hist[threadIdx.x] = base + incl - v;
            __syncthreads();
            s_perm[s_hist[cls * 4u + warp] + rank] = (uint8_t)threadIdx.x;
            __syncthreads();
            i = tile + s_perm[threadIdx.x];      // (paths beyond count sort last: a thread with i >= count idles as before)
        }
#endif
        ShadeResult r; r.alive = false; r.has_shadow = false; r.hit = false;
        if (i < count) r = shade_item<SIMPLE, COUNT>(S, P, fb, qin, hits, i, bounce, cnt);
        {   // closest hits actually shaded (rt_stats::shaded_hits; misses run the miss stage only)
            const uint32_t hit_mask = __ballot_sync(0xFFFFFFFFu, r.hit);
            if (hit_mask && lane == 0) atomicAdd(hit_count, (uint32_t)__popc(hit_mask));
        }
#if RT_SHADE_SORT && RT_SHADE_OCTSORT
        {   // compaction of the tile's surviving paths, grouped by the direction class of the next ray: the traversal kernel
            // refills its warps with runs of consecutive rays, and rays that start close together (one 128-path tile) and
            // leave into the same octant walk the same part of the tree in the same order
            uint32_t key = RT_OCT_CLASSES;
            if (r.alive) {
                const uint32_t oct = 7u - octant_inv(r.next.dir);
#if RT_SHADE_OCTSORT == 1
                key = oct;                                                     // 8 classes: the octant
#elif RT_SHADE_OCTSORT == 2
                const float ax = fabsf(r.next.dir.x), ay = fabsf(r.next.dir.y), az = fabsf(r.next.dir.z);
                key = oct * 3u + (ax >= ay && ax >= az ? 0u : (ay >= az ? 1u : 2u));   // 24 classes: octant x major axis
#else
                key = oct >> 1;                                                // 4 classes: signs of x and y
#endif
            }
            __syncthreads();                                   // s_hist / s_wsum are free again
            s_hist[threadIdx.x] = 0u;
            __syncthreads();
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
            const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (rank == 0u) s_hist[key * 4u + warp] = (uint32_t)__popc(peers);
            __syncthreads();
            // exclusive scan of the (class-major, warp-minor) counts, one entry per thread; the dead class sorts last
            const uint32_t v = s_hist[threadIdx.x];
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
            if (lane == 31u) s_wsum[warp] = incl;
            __syncthreads();
            uint32_t base = 0u;
            for (uint32_t w = 0; w < warp; ++w) base += s_wsum[w];
            s_hist[threadIdx.x] = base + incl - v;
            __syncthreads();
            if (threadIdx.x == 0u) { const uint32_t n_alive = s_hist[RT_OCT_CLASSES * 4u]; s_wsum[0] = n_alive ? atomicAdd(out_count, n_alive) : 0u; }   // one atomic per tile
            __syncthreads();
            const uint32_t slot = s_wsum[0] + s_hist[key * 4u + warp] + rank;
            __syncthreads();                                   // the next tile's sort reuses s_hist / s_wsum
            if (r.alive) store_path(qout, slot, r.next);
        }
#else
        const uint32_t alive_mask = __ballot_sync(0xFFFFFFFFu, r.alive);
        if (alive_mask) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(out_count, (uint32_t)__popc(alive_mask));
            slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
            if (r.alive) store_path(qout, slot0 + (uint32_t)__popc(alive_mask & ((1u << lane) - 1u)), r.next);
        }
#endif
        const uint32_t sh_mask = __ballot_sync(0xFFFFFFFFu, r.has_shadow);
        if (sh_mask) {
            uint32_t slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(shadow_count, (uint32_t)__popc(sh_mask));
            slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
            if (r.has_shadow) {
                const uint32_t s = slot0 + (uint32_t)__popc(sh_mask & ((1u << lane) - 1u));
                sq.o_tmax[s] = make_float4(r.shadow.origin.x, r.shadow.origin.y, r.shadow.origin.z, r.shadow.tmax);
                sq.d_pix[s] = make_float4(r.shadow.dir.x, r.shadow.dir.y, r.shadow.dir.z, rt_uint_as_float(r.shadow.pixel));
                sq.contrib[s] = make_float4(r.shadow.contrib.x, r.shadow.contrib.y, r.shadow.contrib.z, rt_uint_as_float(r.shadow.path_w));
            }
        }
    }
}

// rt_trace_closest / rt_trace_any (default path): the caller's ray set goes through the SAME persistent traversal as the
// frame kernels above (persistent_trace + coop_round: dynamic fetch, deferred warp-cooperative triangle rounds, 64-bit
// atomicMin tie-break, SINGLE specialisation) — only the ray source and the hit sink differ from extend_kernel /
// shadow_kernel, so per-ray ID parity is established on the code the bench times.
template <int MODE, bool ALPHA, bool SINGLE>
__global__ void __launch_bounds__(RT_EXTEND_THREADS, SINGLE ? RT_EXTEND_MIN_BLOCKS_SINGLE : RT_EXTEND_MIN_BLOCKS) trace_wavefront_kernel(DScene S, const rt_ray* rays, uint32_t n, const uint32_t* rng4, rt_hit* hits, uint8_t* occluded, uint32_t* fetch) {
    persistent_trace<MODE, ALPHA, false, SINGLE>(S, n, fetch, nullptr,
        [&](uint32_t i, Trav& tv) {
            const float4 a = rt_ld(reinterpret_cast<const float4*>(rays + i)), b = rt_ld(reinterpret_cast<const float4*>(rays + i) + 1);
            u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
            if (ALPHA && rng4) { const uint4 r = rt_ld(reinterpret_cast<const uint4*>(rng4) + i); rng.x = r.x; rng.y = r.y; rng.z = r.z; rng.w = r.w; }
            trav_init<SINGLE>(tv, S, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), a.w, b.w, rng);
        },
        [&](uint32_t i, const Trav& tv) {
            if (MODE == RT_MODE_ANY) { occluded[i] = tv.found ? 1 : 0; return; }
            rt_hit o;
            if (tv.found) { o.t = tv.hit.t; o.u = tv.hit.u; o.v = tv.hit.v; o.instance_id = tv.hit.inst; o.primitive_id = tv.hit.prim; o.geo_id = rt_float_as_uint(S.inst_w2o[(size_t)tv.hit.inst * RT_INST_F4 + 3].y); }
            else { o.t = -1.0f; o.u = 0.0f; o.v = 0.0f; o.instance_id = o.primitive_id = o.geo_id = 0xFFFFFFFFu; }
            hits[i] = o;
        });
}

template <bool ALPHA, bool COUNT>
__global__ void __launch_bounds__(128) trace_rays_kernel(DScene S, const rt_ray* rays, uint32_t n, const uint32_t* rng4, rt_hit* hits, uint8_t* occluded, RtCounters* cnt) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const rt_ray r = rays[i];
        u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
        if (rng4) { rng.x = rng4[4 * i]; rng.y = rng4[4 * i + 1]; rng.z = rng4[4 * i + 2]; rng.w = rng4[4 * i + 3]; }
        RtHit h;
        if (occluded) {
            occluded[i] = trace_ray<RT_MODE_ANY, ALPHA, COUNT>(S, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, rng, h, cnt) ? 1 : 0;
        } else {
            const bool f = trace_ray<RT_MODE_CLOSEST, ALPHA, COUNT>(S, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, rng, h, cnt);
            rt_hit o;
            if (f) { o.t = h.t; o.u = h.u; o.v = h.v; o.instance_id = h.inst; o.primitive_id = h.prim; o.geo_id = rt_float_as_uint(S.inst_w2o[(size_t)h.inst * RT_INST_F4 + 3].y); }
            else { o.t = -1.0f; o.u = 0.0f; o.v = 0.0f; o.instance_id = o.primitive_id = o.geo_id = 0xFFFFFFFFu; }
            hits[i] = o;
        }
    }
}
#endif
