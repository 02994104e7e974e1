// rt_core.h — contexts, scenes, frame orchestration and the C ABI of include/rt_b200.h.
// Included by rt_api.cu (product, CUDA) and by tests/emu/emu.cpp (RT_EMU: test-only host emulation whose
// exported symbols carry an emu_ prefix and which the rustracer_b200 package never loads).
#pragma once
#include "rt_kernels.h"

#include <string>
#include <vector>
#include <cstdio>
#include <cmath>

#ifdef RT_EMU
#define RT_API(name) emu_##name
#else
#define RT_API(name) name
#endif

namespace rtcore {

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define RT_CHECK(expr, what) do { if (expr) return fail(std::string(what) + ": " + rt_platform_error()); } while (0)

static uint32_t host_tea16(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}

// object->world 3x4 (row-major) -> world->object, evaluated in double with a fixed formula.  The oracle uses the
// same formula so both sides feed identical matrices to the (bit-exact) ray transform.
static void invert_3x4(const float* m, float* out) {
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

template <class T> static int dev_alloc(T** p, size_t count) { return rt_malloc((void**)p, count * sizeof(T)); }
template <class T> static int dev_upload(T** p, const T* src, size_t count, rt_stream_t s) {
    if (dev_alloc(p, count ? count : 1)) return 1;
    if (count && rt_h2d(*p, src, count * sizeof(T), s)) return 1;
    return 0;
}

struct StageEvent { int stage; rt_timer a, b; };

}  // namespace rtcore
using namespace rtcore;

#ifndef RT_MIN_RAYS_PER_CTA
#define RT_MIN_RAYS_PER_CTA 0
#endif
#ifndef RT_MAX_FRAMES_IN_FLIGHT
#define RT_MAX_FRAMES_IN_FLIGHT 4
#endif
#define RT_MAX_SCENE_VERSIONS 4
#define RT_TICKET_RING 24      // a multiple of every slot count (1..8 with RT_MAX_FRAMES_IN_FLIGHT raised)

// One frame in flight: everything a frame writes except the shared accumulation image.  Mirrors the reference's
// InFlightFrames (app/src/lib.rs:34,329,400-401: IN_FLIGHT_FRAMES = 2, one fence per frame) and its per-swapchain-image
// storage images next to the single acc_images[0] (app/src/lib.rs:305-325).
struct FrameSlot {
    rt_stream_t stream = nullptr;        // private stream of the slot (unused while the context has a single slot)
    FrameBuffers fb{};                   // fb.acc aliases rt_context::acc
    DQueue q[2]{};
    DHits hits{};
    DShadowQueue sq{};
    uint32_t* counters = nullptr; size_t counters_cap = 0;   // qcount | scount | fetch_extend | fetch_shadow | shaded hits
    RtCounters* dev_cnt = nullptr;
    // last frame rendered on this slot
    uint32_t last_S = 0, last_B = 0; uint64_t last_pixels = 0; bool last_counted = false; bool last_valid = false;
    rt_timer ev_begin, ev_end;
    std::vector<StageEvent> stage_events; size_t stage_used = 0;
    unsigned long long launches_before = 0, launches_after = 0;
    rt_event done; bool pending = false;   // recorded after the last operation queued for this slot
    const void* scene = nullptr; uint32_t scene_version = 0;   // what the slot's last frame reads (animated scenes are multi-buffered)
};

struct rt_context {
    int device = 0;
    rt_stream_t stream = nullptr;
    rt_stream_t last_stream = nullptr;   // stream of the most recent rt_render / rt_tonemap (may be caller-provided)
    uint32_t width = 0, height = 0;
    float4* acc = nullptr;               // RGBA32F accumulation image shared by all slots (RayTracing.rgen:18)
    // multi-GPU combine (rt_combine): the accumulation image is the head of ONE allocation [acc | snap | display | sync]
    // so that a single CUDA IPC handle (or peer pointer) gives the other GPUs everything they touch:
    float4* snap = nullptr;              //   snapshot of acc a combine works on (frames keep accumulating meanwhile)
    uint32_t* display = nullptr;         //   combined, tonemapped RGBA8 image (every rank's band lands here: all-gather)
    uint32_t* sync = nullptr;            //   RT_SYNC_* words written / polled across GPUs (device-side flags, no host barrier)
    rt_event ev_combine; bool combine_pending = false; rt_stream_t combine_stream = nullptr;   // the last rt_combine (possibly on a side stream)
    uint32_t combine_epoch = 0, recv_expected = 0;   // combines issued so far / display bands expected so far (RT_SYNC_RECV)
    FrameSlot slot[RT_MAX_FRAMES_IN_FLIGHT];
    uint32_t n_slots = 1, cur = 0;       // cur: slot of the most recently submitted frame
    uint64_t frame_seq = 0;              // frames submitted so far (ticket of the next frame)
    rt_event ev_submit, ev_acc, ev_consumer; bool acc_pending = false, consumer_pending = false;
    // fences of the last RT_TICKET_RING submissions (ticket % ring).  The ring length is a multiple of every slot count, so
    // the frame that re-records an entry runs on the same slot stream as the one it replaces: waiting on it is conservative.
    rt_event ticket_done[RT_TICKET_RING];
    bool timers = false;
    std::vector<void*> ipc_opened;
    bool ipc_exported = false;          // rt_ipc_export handed out a handle of `acc`: it must not be re-allocated
};

// One BLAS: either a geometry's own object-space BLAS (needed only when a non-baked instance references it) or the
// merged world-space BLAS of all baked instances.
struct GeoRecord { uint32_t node_off, tri_off, n_tris, n_nodes, depth; bool skinned, needed; };

struct rt_scene {
    rt_context* ctx = nullptr;
    // host mirrors needed for rebuilds
    std::vector<rt_geometry> geometries; std::vector<rt_prim_info> prim_infos; std::vector<rt_instance> instances;
    std::vector<GeoRecord> geo;
    GeoRecord merged{};                      // world-space BLAS over the baked instances' triangles
    std::vector<uint8_t> baked;              // per instance
    uint2* d_bake_src = nullptr;             // per merged primitive: (instance, primitive)
    bool merged_skinned = false;
    uint32_t n_vertices = 0, n_indices = 0, n_materials = 0, n_skins = 0;
    // device arrays
    rt_vertex *d_vin = nullptr, *d_vout = nullptr; uint32_t* d_indices = nullptr; rt_prim_info* d_prim = nullptr;
    rt_material* d_mat = nullptr; float* d_skins = nullptr;
    rt_light *d_dl = nullptr, *d_pl = nullptr; uint32_t cap_dl = 0, cap_pl = 0;
    std::vector<uint8_t*> d_image_px; DImage* d_images = nullptr; DTexture* d_textures = nullptr; float* d_lut = nullptr;
    uint8_t* d_sky[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float4 *d_blas_nodes = nullptr, *d_tris = nullptr; DAabb* d_node_box = nullptr; uint32_t* d_node_parent = nullptr;
    uint32_t* d_prim_order = nullptr; DAabb* d_leaf_boxes = nullptr; DAabb* d_prim_boxes = nullptr; uint32_t* d_pending = nullptr;
    uint64_t total_nodes = 0, total_tris = 0;
    float4 *d_tlas_nodes = nullptr; uint32_t* d_tlas_prims = nullptr; DAabb* d_tlas_box = nullptr; uint32_t* d_tlas_parent = nullptr;
    DAabb* d_inst_boxes = nullptr; float4 *d_inst_w2o = nullptr, *d_inst_o2w = nullptr; uint32_t* d_inst_root = nullptr; uint32_t* d_entry_rec = nullptr;
    uint32_t tlas_nodes = 0, tlas_depth = 0, blas_depth = 0;
    BuildScratch scratch;
    DScene ds{};
    bool has_nee = false;
    bool plain_materials = false;            // no texture reference and no specular-glossiness material anywhere
    float build_ms = 0, refit_ms = 0, skin_ms = 0, tlas_ms = 0;
    // Buffers a skin update rewrites while frames may still read them: skinned vertices, packed triangles, BLAS and
    // TLAS nodes.  A skinned scene keeps n_versions copies; d_vout / d_tris / d_blas_nodes / d_tlas_nodes alias the
    // current one, a refit switches to the next and only waits for the frames that still read THAT copy, so the
    // update of frame f+1 overlaps the rendering of frame f (the reference rebuilds in place behind three fence waits,
    // gltf_viewer/src/main.rs:384-395).
    struct Version { rt_vertex* vout = nullptr; float4 *tris = nullptr, *blas_nodes = nullptr, *tlas_nodes = nullptr; };
    Version ver[RT_MAX_SCENE_VERSIONS]; uint32_t n_versions = 1, cur_ver = 0;
    // TLAS refit state (rt_scene_update_skins keeps the TLAS topology): leaf k -> entry, boxes in leaf order, counters
    uint32_t* d_tlas_leaf_entry = nullptr; DAabb* d_tlas_leaf_boxes = nullptr; uint32_t* d_tlas_pending = nullptr; uint32_t tlas_entries = 0;
    // asynchronous skin updates: frames wait for ev_updated on the device; the stage timers are read lazily
    rt_event ev_updated; bool update_pending = false;
    rt_timer upd_t[4]; bool upd_timers_pending = false, upd_timers_created = false;
};

namespace rtcore {

static void free_slot(FrameSlot* f) {
    void* ptrs[] = {f->fb.out, f->fb.rad, f->fb.aux, f->fb.pixrng, f->hits.tuvp, f->hits.inst, f->sq.o_tmax, f->sq.d_pix, f->sq.contrib,
                    f->q[0].o_tmin, f->q[0].d_tmax, f->q[0].thr_pix, f->q[0].rng, f->q[1].o_tmin, f->q[1].d_tmax, f->q[1].thr_pix, f->q[1].rng};
    for (void* p : ptrs) if (p) rt_free(p);
    f->fb = FrameBuffers{}; f->hits = DHits{}; f->sq = DShadowQueue{}; f->q[0] = DQueue{}; f->q[1] = DQueue{};
    f->last_valid = false; f->pending = false;
}
static void free_frame(rt_context* c) {
    for (uint32_t k = 0; k < RT_MAX_FRAMES_IN_FLIGHT; ++k) free_slot(&c->slot[k]);
    if (c->acc) rt_free(c->acc);
    c->acc = nullptr; c->snap = nullptr; c->display = nullptr; c->sync = nullptr;
}
// byte offsets inside the accumulation block (the same arithmetic locates a peer's snap / display / sync from its base)
static inline size_t acc_block_snap(size_t n) { return n * sizeof(float4); }
static inline size_t acc_block_display(size_t n) { return 2 * n * sizeof(float4); }
static inline size_t acc_block_sync(size_t n) { return 2 * n * sizeof(float4) + ((n * 4 + 255) & ~(size_t)255); }
static inline size_t acc_block_bytes(size_t n) { return acc_block_sync(n) + 256; }

static int alloc_slot(rt_context* c, FrameSlot* f, size_t n) {
    int e = 0;
    e |= dev_alloc(&f->fb.out, n); e |= dev_alloc(&f->fb.rad, n); e |= dev_alloc(&f->fb.aux, n); e |= dev_alloc(&f->fb.pixrng, n);
    e |= dev_alloc(&f->hits.tuvp, n); e |= dev_alloc(&f->hits.inst, n);
    e |= dev_alloc(&f->sq.o_tmax, n); e |= dev_alloc(&f->sq.d_pix, n); e |= dev_alloc(&f->sq.contrib, n);
    for (int k = 0; k < 2; ++k) { e |= dev_alloc(&f->q[k].o_tmin, n); e |= dev_alloc(&f->q[k].d_tmax, n); e |= dev_alloc(&f->q[k].thr_pix, n); e |= dev_alloc(&f->q[k].rng, n); }
    if (!f->dev_cnt) { e |= dev_alloc(&f->dev_cnt, 1); if (!e) rt_memset(f->dev_cnt, 0, sizeof(RtCounters), c->stream); }
    if (e) return 1;
    f->fb.acc = c->acc;
    rt_memset(f->fb.out, 0, n * 4, c->stream);
    rt_memset(f->fb.rad, 0, n * sizeof(float4), c->stream); rt_memset(f->fb.aux, 0, n * sizeof(float2), c->stream);
    if (!f->stream && rt_stream_create(&f->stream)) return 1;
    f->done.create(); f->ev_begin.create(); f->ev_end.create();
    return 0;
}

// (re)allocates the accumulation image and n_slots frame slots; the caller has synchronised the context
static int alloc_frame(rt_context* c, uint32_t w, uint32_t h) {
    free_frame(c);
    const size_t n = (size_t)w * h;
    if (rt_malloc((void**)&c->acc, acc_block_bytes(n))) return 1;
    c->snap = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->acc) + acc_block_snap(n));
    c->display = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(c->acc) + acc_block_display(n));
    c->sync = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(c->acc) + acc_block_sync(n));
    c->combine_epoch = 0; c->recv_expected = 0;
    rt_memset(c->acc, 0, acc_block_bytes(n), c->stream);
    for (uint32_t k = 0; k < c->n_slots; ++k) if (alloc_slot(c, &c->slot[k], n)) { free_frame(c); return 1; }
    c->width = w; c->height = h; c->cur = 0; c->acc_pending = false; c->consumer_pending = false;
    return rt_stream_sync(c->stream);
}

static void update_ds(rt_scene* s) {
    DScene& d = s->ds;
    d.tlas_nodes = s->d_tlas_nodes; d.tlas_prims = s->d_tlas_prims; d.blas_nodes = s->d_blas_nodes; d.tris = s->d_tris;
    d.inst_w2o = s->d_inst_w2o; d.inst_o2w = s->d_inst_o2w; d.n_instances = (uint32_t)s->instances.size();
    d.vertices = s->d_vout; d.indices = s->d_indices; d.prim_infos = s->d_prim; d.materials = s->d_mat;
    d.textures = s->d_textures; d.images = s->d_images; d.srgb_lut = s->d_lut;
    d.dlights = s->d_dl; d.plights = s->d_pl;
}

// ---- multi-buffered animated scenes ---------------------------------------------------------------------
static size_t ver_bytes_vout(const rt_scene* s) { return (size_t)(s->n_vertices ? s->n_vertices : 1) * sizeof(rt_vertex); }
static size_t ver_bytes_tris(const rt_scene* s) { return (size_t)(s->total_tris ? s->total_tris : 1) * RT_TRI_F4 * sizeof(float4); }
static size_t ver_bytes_blas(const rt_scene* s) { return (size_t)(s->total_nodes ? s->total_nodes : 1) * RT_NODE_F4 * sizeof(float4); }
static size_t ver_bytes_tlas(const rt_scene* s) { return (size_t)(s->instances.size() + 1) * RT_NODE_F4 * sizeof(float4); }
static void use_version(rt_scene* s, uint32_t v) {
    s->cur_ver = v;
    s->d_vout = s->ver[v].vout; s->d_tris = s->ver[v].tris; s->d_blas_nodes = s->ver[v].blas_nodes; s->d_tlas_nodes = s->ver[v].tlas_nodes;
}
// after a (synchronous) build every copy must hold the new topology: replicate the current version
static int replicate_versions(rt_scene* s) {
    rt_stream_t st = s->ctx->stream; const uint32_t c = s->cur_ver; int e = 0;
    for (uint32_t v = 0; v < s->n_versions; ++v) {
        if (v == c) continue;
        e |= rt_d2d(s->ver[v].vout, s->ver[c].vout, ver_bytes_vout(s), st); e |= rt_d2d(s->ver[v].tris, s->ver[c].tris, ver_bytes_tris(s), st);
        e |= rt_d2d(s->ver[v].blas_nodes, s->ver[c].blas_nodes, ver_bytes_blas(s), st); e |= rt_d2d(s->ver[v].tlas_nodes, s->ver[c].tlas_nodes, ver_bytes_tlas(s), st);
    }
    return e;
}
static int alloc_versions(rt_scene* s, uint32_t n) {
    for (uint32_t v = s->n_versions; v < n; ++v) {
        int e = rt_malloc((void**)&s->ver[v].vout, ver_bytes_vout(s)); e |= rt_malloc((void**)&s->ver[v].tris, ver_bytes_tris(s));
        e |= rt_malloc((void**)&s->ver[v].blas_nodes, ver_bytes_blas(s)); e |= rt_malloc((void**)&s->ver[v].tlas_nodes, ver_bytes_tlas(s));
        if (e) return 1;
    }
    return 0;
}

static int run_skinning(rt_scene* s) {
    rt_stream_t st = s->ctx->stream;
    const rt_vertex* vin = s->d_vin; rt_vertex* vout = s->d_vout; const float* skins = s->d_skins; const uint32_t ns = s->n_skins;
#ifndef RT_EMU
    if (s->n_vertices) {
        size_t blocks = ((size_t)s->n_vertices + RT_SKIN_TILE - 1) / RT_SKIN_TILE; const size_t cap = (size_t)g_rt_sm_count * RT_SKIN_CTAS_PER_SM;   // persistent: every CTA streams several double-buffered tiles
        if (blocks > cap) blocks = cap;
        skin_kernel<<<(unsigned)blocks, RT_SKIN_TILE, 0, st>>>(vin, vout, skins, ns, s->n_vertices);
        ++g_rt_launch_count;
    }
#else
    rt_launch(s->n_vertices, st, RT_LAMBDA(size_t i) { skin_item(vin, vout, skins, ns, (uint32_t)i); });
#endif
    return 0;
}

// Baking rule (must match oracle/oracle.cpp): an instance is intersected in world space through the merged BLAS when
// its geometry is referenced by exactly one instance or has at most RT_BAKE_MAX_TRIS triangles.
static void classify_instances(rt_scene* s) {
    std::vector<uint32_t> refs(s->geometries.size(), 0);
    for (auto& in : s->instances) refs[in.geo_id]++;
    s->baked.assign(s->instances.size(), 0);
    for (size_t i = 0; i < s->instances.size(); ++i) {
        const uint32_t g = s->instances[i].geo_id;
        s->baked[i] = (refs[g] == 1 || s->geometries[g].i_len / 3 <= RT_BAKE_MAX_TRIS) ? 1 : 0;
    }
}

// (re)packs the triangles of geometry g in leaf order (object space) and refreshes the per-leaf boxes
static void pack_tris(rt_scene* s, uint32_t g) {
    const GeoRecord& gr = s->geo[g]; const rt_prim_info pi = s->prim_infos[g];
    const rt_vertex* verts = s->d_vout; const uint32_t* indices = s->d_indices;
    const uint32_t* order = s->d_prim_order + gr.tri_off; float4* tris = s->d_tris + (size_t)gr.tri_off * RT_TRI_F4; DAabb* lb = s->d_leaf_boxes + gr.tri_off;
    rt_launch(gr.n_tris, s->ctx->stream, RT_LAMBDA(size_t k) {
        const uint32_t prim = order[k];
        const uint32_t io = pi.i_offset + 3 * prim;
        const float* p0 = verts[pi.v_offset + indices[io]].position; const float* p1 = verts[pi.v_offset + indices[io + 1]].position; const float* p2 = verts[pi.v_offset + indices[io + 2]].position;
        tris[k * RT_TRI_F4 + 0] = make_float4(p0[0], p0[1], p0[2], rt_uint_as_float(prim));
        tris[k * RT_TRI_F4 + 1] = make_float4(p1[0], p1[1], p1[2], 0.0f);
        tris[k * RT_TRI_F4 + 2] = make_float4(p2[0], p2[1], p2[2], rt_uint_as_float(prim));      // .w: tie-break rank inside this BLAS
        DAabb b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = fminf(p0[a], fminf(p1[a], p2[a])); b.hi[a] = fmaxf(p0[a], fmaxf(p1[a], p2[a])); }
        lb[k] = b;
    });
}

// world-space vertices of merged primitive `src` = (instance, primitive): object->world with the oracle's operation order
RT_D void baked_triangle(const rt_vertex* verts, const uint32_t* indices, const rt_prim_info* prims, const float4* inst_w2o, const float4* inst_o2w,
                         uint2 src, f3& w0, f3& w1, f3& w2) {
    const uint32_t geo = rt_float_as_uint(inst_w2o[(size_t)src.x * RT_INST_F4 + 3].y);
    const rt_prim_info pi = prims[geo];
    const float4 m0 = inst_o2w[(size_t)src.x * RT_O2W_F4], m1 = inst_o2w[(size_t)src.x * RT_O2W_F4 + 1], m2 = inst_o2w[(size_t)src.x * RT_O2W_F4 + 2];
    const uint32_t io = pi.i_offset + 3 * src.y;
    const float* p0 = verts[pi.v_offset + indices[io]].position; const float* p1 = verts[pi.v_offset + indices[io + 1]].position; const float* p2 = verts[pi.v_offset + indices[io + 2]].position;
    w0 = xform_point_exact(m0, m1, m2, mk3(p0[0], p0[1], p0[2]));
    w1 = xform_point_exact(m0, m1, m2, mk3(p1[0], p1[1], p1[2]));
    w2 = xform_point_exact(m0, m1, m2, mk3(p2[0], p2[1], p2[2]));
}

static void pack_merged(rt_scene* s) {
    const GeoRecord& gr = s->merged;
    const rt_vertex* verts = s->d_vout; const uint32_t* indices = s->d_indices; const rt_prim_info* prims = s->d_prim;
    const float4* w2o = s->d_inst_w2o; const float4* o2w = s->d_inst_o2w; const uint2* src = s->d_bake_src;
    const uint32_t* order = s->d_prim_order + gr.tri_off; float4* tris = s->d_tris + (size_t)gr.tri_off * RT_TRI_F4; DAabb* lb = s->d_leaf_boxes + gr.tri_off;
    rt_launch(gr.n_tris, s->ctx->stream, RT_LAMBDA(size_t k) {
        const uint2 sp = src[order[k]];
        f3 a, b, c; baked_triangle(verts, indices, prims, w2o, o2w, sp, a, b, c);
        tris[k * RT_TRI_F4 + 0] = make_float4(a.x, a.y, a.z, rt_uint_as_float(sp.y));
        tris[k * RT_TRI_F4 + 1] = make_float4(b.x, b.y, b.z, rt_uint_as_float(sp.x));
        tris[k * RT_TRI_F4 + 2] = make_float4(c.x, c.y, c.z, rt_uint_as_float(order[k]));   // .w: rank of (instance, primitive): bake_src is sorted by it
        DAabb bb;
        bb.lo[0] = fminf(a.x, fminf(b.x, c.x)); bb.lo[1] = fminf(a.y, fminf(b.y, c.y)); bb.lo[2] = fminf(a.z, fminf(b.z, c.z));
        bb.hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); bb.hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); bb.hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
        lb[k] = bb;
    });
}

static int build_one(rt_scene* s, GeoRecord& gr, const char* what) {
    WideOut out;
    out.nodes = s->d_blas_nodes + (size_t)gr.node_off * RT_NODE_F4; out.prim_order = s->d_prim_order + gr.tri_off;
    out.node_box = s->d_node_box + gr.node_off; out.node_parent = s->d_node_parent + gr.node_off; out.max_nodes = gr.n_tris ? gr.n_tris : 1u;
    WideBvhInfo info;
    const int e = build_wide_bvh(s->d_prim_boxes, gr.n_tris, s->scratch, out, s->ctx->stream, &info);
    if (e) return fail(std::string(what) + " build failed (code " + std::to_string(e) + ")");
    gr.n_nodes = info.n_nodes; gr.depth = info.depth;
    return 0;
}

static int build_blas(rt_scene* s, uint32_t g) {
    GeoRecord& gr = s->geo[g]; const rt_prim_info pi = s->prim_infos[g];
    const rt_vertex* verts = s->d_vout; const uint32_t* indices = s->d_indices; DAabb* pb = s->d_prim_boxes;
    rt_launch(gr.n_tris, s->ctx->stream, RT_LAMBDA(size_t k) { pb[k] = tri_box_of(verts, indices, pi.v_offset, pi.i_offset, (uint32_t)k); });
    if (build_one(s, gr, "BLAS")) return 1;
    pack_tris(s, g);
    return 0;
}

static int build_merged(rt_scene* s) {
    GeoRecord& gr = s->merged;
    const rt_vertex* verts = s->d_vout; const uint32_t* indices = s->d_indices; const rt_prim_info* prims = s->d_prim;
    const float4* w2o = s->d_inst_w2o; const float4* o2w = s->d_inst_o2w; const uint2* src = s->d_bake_src; DAabb* pb = s->d_prim_boxes;
    rt_launch(gr.n_tris, s->ctx->stream, RT_LAMBDA(size_t k) {
        f3 a, b, c; baked_triangle(verts, indices, prims, w2o, o2w, src[k], a, b, c);
        DAabb bb;
        bb.lo[0] = fminf(a.x, fminf(b.x, c.x)); bb.lo[1] = fminf(a.y, fminf(b.y, c.y)); bb.lo[2] = fminf(a.z, fminf(b.z, c.z));
        bb.hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); bb.hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); bb.hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
        pb[k] = bb;
    });
    if (build_one(s, gr, "merged BLAS")) return 1;
    pack_merged(s);
    return 0;
}

// bottom-up refit of one BLAS whose triangles / leaf boxes were just re-packed
static void refit_nodes(rt_scene* s, const GeoRecord& gr) {
    rt_stream_t st = s->ctx->stream;
    float4* nodes = s->d_blas_nodes + (size_t)gr.node_off * RT_NODE_F4; DAabb* nb = s->d_node_box + gr.node_off;
    const uint32_t* parent = s->d_node_parent + gr.node_off; uint32_t* pending = s->d_pending + gr.node_off;
    const DAabb* lb = s->d_leaf_boxes + gr.tri_off;
    rt_launch(gr.n_nodes, st, RT_LAMBDA(size_t w) { pending[w] = (uint32_t)rt_popc(rt_float_as_uint(nodes[w * RT_NODE_F4].w) >> 24); });
#ifndef RT_EMU
    if (gr.n_nodes) {
        refit_nodes_kernel<<<(unsigned)(((size_t)gr.n_nodes * 8 + 255) / 256), 256, 0, st>>>(nodes, nb, lb, parent, pending, gr.n_nodes);
        ++g_rt_launch_count;
    }
    return;
#endif
    rt_launch(gr.n_nodes, st, RT_LAMBDA(size_t w0) {
        uint32_t w = (uint32_t)w0;
        if ((rt_float_as_uint(nodes[(size_t)w * RT_NODE_F4].w) >> 24) != 0u) return;   // only nodes without inner children start
        for (;;) {
            refit_wide_node(nodes, nb, lb, w);
            const uint32_t p = parent[w];
            if (p == 0xFFFFFFFFu) return;
            rt_threadfence();
            const uint32_t old = rt_atomic_add(&pending[p], 0xFFFFFFFFu);   // decrement
            if (old != 1u) return;   // other children still pending
            rt_threadfence();
            w = p;
        }
    });
}
static void refit_blas(rt_scene* s, uint32_t g) { pack_tris(s, g); refit_nodes(s, s->geo[g]); }
static void refit_merged(rt_scene* s) { pack_merged(s); refit_nodes(s, s->merged); }

// uploads the instance records (transform inverses, BLAS offsets, flags); record n_instances is the merged BLAS
static int upload_instance_records(rt_scene* s) {
    rt_stream_t st = s->ctx->stream;
    const uint32_t n = (uint32_t)s->instances.size();
    std::vector<float> w2o((size_t)(n + 1) * 16), o2w((size_t)(n ? n : 1) * 4 * RT_O2W_F4);
    for (uint32_t i = 0; i < n; ++i) {
        const rt_instance& in = s->instances[i];
        float inv[12]; invert_3x4(in.transform, inv);
        memcpy(&w2o[(size_t)i * 16], inv, 48);
        const GeoRecord& gr = s->geo[in.geo_id];
        uint32_t meta[4] = {gr.node_off, in.geo_id, s->geometries[in.geo_id].opaque ? RT_INST_OPAQUE : 0u, gr.tri_off};
        memcpy(&w2o[(size_t)i * 16 + 12], meta, 16);
        memcpy(&o2w[(size_t)i * 4 * RT_O2W_F4], in.transform, 48);
        // 4th row: the geometry's PrimInfo, so that shading resolves indices and material without the geo -> PrimInfo hop
        const rt_prim_info& pi = s->prim_infos[in.geo_id];
        const uint32_t shade_meta[4] = {pi.v_offset, pi.i_offset, pi.material_id, in.geo_id};
        memcpy(&o2w[(size_t)i * 4 * RT_O2W_F4 + 12], shade_meta, 16);
    }
    const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(&w2o[(size_t)n * 16], ident, 48);
    uint32_t meta[4] = {s->merged.node_off, 0u, RT_INST_IDENTITY | RT_INST_MERGED, s->merged.tri_off};
    memcpy(&w2o[(size_t)n * 16 + 12], meta, 16);
    RT_CHECK(rt_h2d(s->d_inst_w2o, w2o.data(), w2o.size() * 4, st), "upload instances");
    if (n) RT_CHECK(rt_h2d(s->d_inst_o2w, o2w.data(), (size_t)n * 16 * RT_O2W_F4, st), "upload instances");
    return 0;
}

// world-space boxes of the TLAS entries (entry e -> instance record d_entry_rec[e], BLAS root d_inst_root[e])
static void instance_boxes(rt_scene* s, uint32_t ne) {
    rt_stream_t st = s->ctx->stream;
    const uint32_t n = (uint32_t)s->instances.size();
    const float4* o2wd = s->d_inst_o2w; const uint32_t* rootd = s->d_inst_root; const uint32_t* recd = s->d_entry_rec; const DAabb* nb = s->d_node_box; DAabb* ib = s->d_inst_boxes;
    rt_launch(ne, st, RT_LAMBDA(size_t e) {
        const DAabb b = nb[rootd[e]];
        const uint32_t rec = recd[e];
        DAabb w;
        if (rec == n) { w = b; }      // merged BLAS: already world space
        else {
            const float4 m0 = o2wd[(size_t)rec * RT_O2W_F4], m1 = o2wd[(size_t)rec * RT_O2W_F4 + 1], m2 = o2wd[(size_t)rec * RT_O2W_F4 + 2];
            for (int c = 0; c < 8; ++c) {
                const float x = (c & 1) ? b.hi[0] : b.lo[0], y = (c & 2) ? b.hi[1] : b.lo[1], z = (c & 4) ? b.hi[2] : b.lo[2];
                const float p[3] = {m0.x * x + m0.y * y + m0.z * z + m0.w, m1.x * x + m1.y * y + m1.z * z + m1.w, m2.x * x + m2.y * y + m2.z * z + m2.w};
                for (int a = 0; a < 3; ++a) { if (c == 0) { w.lo[a] = p[a]; w.hi[a] = p[a]; } else { w.lo[a] = fminf(w.lo[a], p[a]); w.hi[a] = fmaxf(w.hi[a], p[a]); } }
            }
        }
        // the corners were rounded and the object-space ray is rounded too: keep the world box conservative
        for (int a = 0; a < 3; ++a) {
            const float pad = (w.hi[a] - w.lo[a]) * 1e-5f + 1e-6f * fmaxf(fabsf(w.lo[a]), fabsf(w.hi[a])) + 1e-7f;
            w.lo[a] -= pad; w.hi[a] += pad;
        }
        ib[e] = w;
    });
}

// TLAS over the non-baked instances plus (when it has triangles) the merged world-space BLAS
static int build_tlas(rt_scene* s) {
    rt_stream_t st = s->ctx->stream;
    const uint32_t n = (uint32_t)s->instances.size();
    std::vector<uint32_t> entry_rec, entry_root;
    for (uint32_t i = 0; i < n; ++i) if (!s->baked[i]) { entry_rec.push_back(i); entry_root.push_back(s->geo[s->instances[i].geo_id].node_off); }
    // the merged BLAS is a TLAS entry only when it is the whole scene (SINGLE kernels); next to real instances rays start in it
    const bool only_merged = s->merged.n_tris && entry_rec.empty();
    s->ds.merged_first = (s->merged.n_tris && !only_merged && !getenv("RT_B200_NO_MERGED_FIRST")) ? 1u : 0u;   // (env: A/B knob)
    if (s->merged.n_tris && !s->ds.merged_first) { entry_rec.push_back(n); entry_root.push_back(s->merged.node_off); }
    const uint32_t ne = (uint32_t)entry_rec.size();
    if (ne) {
        RT_CHECK(rt_h2d(s->d_inst_root, entry_root.data(), (size_t)ne * 4, st), "upload TLAS entries");
        RT_CHECK(rt_h2d(s->d_entry_rec, entry_rec.data(), (size_t)ne * 4, st), "upload TLAS entries");
    }
    instance_boxes(s, ne);
    const uint32_t* recd = s->d_entry_rec; DAabb* ib = s->d_inst_boxes;
    WideOut out; out.nodes = s->d_tlas_nodes; out.prim_order = s->d_tlas_prims; out.node_box = s->d_tlas_box; out.node_parent = s->d_tlas_parent; out.max_nodes = ne ? ne : 1u;
    out.leaf_max = 1u;
    if (const char* e = getenv("RT_B200_TLAS_LEAF")) { const int v = atoi(e); out.leaf_max = (uint32_t)(v < 1 ? 1 : (v > (int)RT_LEAF_MAX ? (int)RT_LEAF_MAX : v)); }   // (env: A/B knob)
    WideBvhInfo info;
    const int e = build_wide_bvh(ib, ne, s->scratch, out, st, &info);
    if (e) return fail("TLAS build failed (code " + std::to_string(e) + ")");
    uint32_t* tp = s->d_tlas_prims; uint32_t* le = s->d_tlas_leaf_entry;
    rt_launch(ne, st, RT_LAMBDA(size_t k) { le[k] = tp[k]; tp[k] = recd[tp[k]]; });   // TLAS leaf -> entry (kept for refits) -> instance record index
    s->tlas_nodes = info.n_nodes; s->tlas_depth = info.depth; s->tlas_entries = ne;
    s->ds.single_merged = only_merged ? 1u : 0u;
    s->ds.merged_node_off = s->merged.node_off; s->ds.merged_tri_off = s->merged.tri_off;
    uint32_t bd = s->merged.depth;
    for (auto& g : s->geo) if (g.needed && g.depth > bd) bd = g.depth;
    s->blas_depth = bd;
    if (s->tlas_depth + s->blas_depth + 6 > RT_STACK_SIZE)
        return fail("BVH too deep for the traversal stack: tlas " + std::to_string(s->tlas_depth) + " + blas " + std::to_string(s->blas_depth));
    return 0;
}

// keeps the TLAS topology, refreshes the entry boxes and refits its nodes (no host synchronisation)
static int refit_tlas(rt_scene* s) {
#ifdef RT_EMU
    return build_tlas(s);
#else
    rt_stream_t st = s->ctx->stream;
    const uint32_t ne = s->tlas_entries;
    if (!ne || !s->tlas_nodes) return 0;
    instance_boxes(s, ne);
    const DAabb* ib = s->d_inst_boxes; DAabb* lbx = s->d_tlas_leaf_boxes; const uint32_t* le = s->d_tlas_leaf_entry;
    float4* nodes = s->d_tlas_nodes; uint32_t* pending = s->d_tlas_pending;
    rt_launch(ne, st, RT_LAMBDA(size_t k) { lbx[k] = ib[le[k]]; });
    rt_launch(s->tlas_nodes, st, RT_LAMBDA(size_t w) { pending[w] = (uint32_t)rt_popc(rt_float_as_uint(nodes[w * RT_NODE_F4].w) >> 24); });
    refit_nodes_kernel<<<(unsigned)(((size_t)s->tlas_nodes * 8 + 255) / 256), 256, 0, st>>>(nodes, s->d_tlas_box, lbx, s->d_tlas_parent, pending, s->tlas_nodes);
    ++g_rt_launch_count;
    return 0;
#endif
}

static void compute_has_nee(rt_scene* s, const rt_light* pl, uint32_t n) {
    // RayTracing.rchit:86: a candidate slot i < min(n,3) is skipped when luminance(color*intensity) < 0.1
    s->has_nee = false;
    for (uint32_t i = 0; i < n && i < 3; ++i) {
        const float lum = 0.2126f * pl[i].color[0] * pl[i].intensity + 0.7152f * pl[i].color[1] * pl[i].intensity + 0.0722f * pl[i].color[2] * pl[i].intensity;
        if (!(lum < 0.1f)) s->has_nee = true;
    }
    s->ds.nee_plights = s->has_nee ? n : 0u;
}

static int upload_sky(rt_scene* s, const uint8_t* const faces[6], uint32_t w, uint32_t h, uint32_t srgb) {
    for (int f = 0; f < 6; ++f) {
        if (s->d_sky[f]) { rt_free(s->d_sky[f]); s->d_sky[f] = nullptr; }
        RT_CHECK(dev_upload(&s->d_sky[f], faces[f], (size_t)w * h * 4, s->ctx->stream), "skybox upload");
        s->ds.sky[f].px = s->d_sky[f]; s->ds.sky[f].w = w; s->ds.sky[f].h = h; s->ds.sky[f].srgb = srgb; s->ds.sky[f]._pad = 0;
    }
    s->ds.has_sky_faces = 1;
    return 0;
}

static uint32_t owned_rows(const TilePart& tp) {
    if (tp.n_parts <= 1) return tp.height;
    uint32_t rows = 0;
    const uint32_t n_strips = (tp.height + tp.strip_rows - 1) / tp.strip_rows;
    for (uint32_t k = tp.part; k < n_strips; k += tp.n_parts) rows += ((k + 1) * tp.strip_rows <= tp.height) ? tp.strip_rows : (tp.height - k * tp.strip_rows);
    return rows;
}

static int sync_all(rt_context* c) {
    int e = 0;
    if (c->combine_pending) { e |= c->ev_combine.sync(); c->combine_pending = false; }    // a combine on a caller's side stream
    if (c->last_stream && c->last_stream != c->stream) e |= rt_stream_sync(c->last_stream);
    if (c->n_slots > 1) for (uint32_t k = 0; k < c->n_slots; ++k) if (c->slot[k].stream) e |= rt_stream_sync(c->slot[k].stream);
    e |= rt_stream_sync(c->stream);
    return e;
}
// device-side join: `st` waits for every frame submitted so far (no host block)
static void join_frames(rt_context* c, rt_stream_t st) {
    // (with a single slot the frame ran on the caller's stream; its `done` event was recorded there)
    for (uint32_t k = 0; k < c->n_slots; ++k) if (c->slot[k].pending && (c->n_slots == 1 || c->slot[k].stream != st)) c->slot[k].done.wait(st);
}
// a consumer of the accumulation image ran on `st`: later frames must not accumulate before it
static void consumer_ran(rt_context* c, rt_stream_t st) {
    c->ev_consumer.record(st); c->consumer_pending = true;
}

static StageEvent* stage_begin(FrameSlot* c, bool on, int stage, rt_stream_t st) {
    if (!on) return nullptr;
    if (c->stage_used == c->stage_events.size()) { StageEvent e; e.stage = stage; e.a.create(); e.b.create(); c->stage_events.push_back(e); }
    StageEvent* e = &c->stage_events[c->stage_used++];
    e->stage = stage; e->a.record(st);
    return e;
}
static void stage_end(StageEvent* e, rt_stream_t st) { if (e) e->b.record(st); }

#ifndef RT_EMU
static inline unsigned persistent_grid(int blocks_per_sm) { return (unsigned)((g_rt_sm_count > 0 ? g_rt_sm_count : 148) * blocks_per_sm); }
// traversal kernels: blocks per SM of the persistent grid (RT_B200_TRACE_BLOCKS overrides; experiments only)
static inline unsigned trace_grid() {
    static int bps = -1;
    if (bps < 0) { const char* e = getenv("RT_B200_TRACE_BLOCKS"); bps = e ? atoi(e) : 8; if (bps < 1) bps = 1; }
    return persistent_grid(bps);
}
#endif

#ifndef RT_EMU
// CTAs of a traversal launch beyond queue size / this leave immediately (0 = every CTA stays); RT_B200_MIN_RAYS_PER_CTA overrides
static inline uint32_t min_rays_per_cta() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("RT_B200_MIN_RAYS_PER_CTA"); v = e ? atoi(e) : RT_MIN_RAYS_PER_CTA; if (v < 0) v = 0; }
    return (uint32_t)v;
}
#endif

template <bool ALPHA, bool COUNT>
static void launch_extend(FrameSlot* c, const DScene& S, const FrameParams& P, const DQueue& q, const uint32_t* count, uint32_t* fetch, uint32_t max_count, rt_stream_t st) {
#ifdef RT_EMU
    const uint32_t n = *count;
    for (uint32_t i = 0; i < n; ++i) extend_item<ALPHA, COUNT>(S, P, q, c->hits, i, c->dev_cnt);
    (void)fetch; (void)max_count; (void)st;
#else
    if (S.single_merged) extend_kernel<ALPHA, COUNT, true><<<trace_grid(), RT_EXTEND_THREADS, 0, st>>>(S, P, q, c->hits, count, fetch, c->dev_cnt, min_rays_per_cta());
    else extend_kernel<ALPHA, COUNT, false><<<trace_grid(), RT_EXTEND_THREADS, 0, st>>>(S, P, q, c->hits, count, fetch, c->dev_cnt, min_rays_per_cta());
    ++g_rt_launch_count; (void)max_count;
#endif
}
template <bool ALPHA, bool COUNT>
static void launch_shadow(FrameSlot* c, const DScene& S, const FrameParams& P, const uint32_t* count, uint32_t* fetch, rt_stream_t st) {
#ifdef RT_EMU
    const uint32_t n = *count;
    for (uint32_t i = 0; i < n; ++i) shadow_item<ALPHA, COUNT>(S, P, c->fb, c->sq, i, c->dev_cnt);
    (void)fetch; (void)st;
#else
    if (S.single_merged) shadow_kernel<ALPHA, COUNT, true><<<trace_grid(), RT_EXTEND_THREADS, 0, st>>>(S, P, c->fb, c->sq, count, fetch, c->dev_cnt, min_rays_per_cta());
    else shadow_kernel<ALPHA, COUNT, false><<<trace_grid(), RT_EXTEND_THREADS, 0, st>>>(S, P, c->fb, c->sq, count, fetch, c->dev_cnt, min_rays_per_cta());
    ++g_rt_launch_count;
#endif
}
template <bool SIMPLE, bool COUNT>
static void launch_shade(FrameSlot* c, const DScene& S, const FrameParams& P, const DQueue& qin, const DQueue& qout, const uint32_t* count, uint32_t* out_count,
                         uint32_t* shadow_count, uint32_t* hit_count, uint32_t bounce, rt_stream_t st) {
#ifdef RT_EMU
    const uint32_t n = *count;
    for (uint32_t i = 0; i < n; ++i) {
        ShadeResult r = shade_item<SIMPLE, COUNT>(S, P, c->fb, qin, c->hits, i, bounce, c->dev_cnt);
        if (r.hit) ++*hit_count;
        if (r.alive) store_path(qout, (*out_count)++, r.next);
        if (r.has_shadow) {
            const uint32_t k = (*shadow_count)++;
            c->sq.o_tmax[k] = make_float4(r.shadow.origin.x, r.shadow.origin.y, r.shadow.origin.z, r.shadow.tmax);
            c->sq.d_pix[k] = make_float4(r.shadow.dir.x, r.shadow.dir.y, r.shadow.dir.z, rt_uint_as_float(r.shadow.pixel));
            c->sq.contrib[k] = make_float4(r.shadow.contrib.x, r.shadow.contrib.y, r.shadow.contrib.z, rt_uint_as_float(r.shadow.path_w));
        }
    }
    (void)st;
#else
    shade_kernel<SIMPLE, COUNT><<<persistent_grid(8), 128, 0, st>>>(S, P, c->fb, qin, c->hits, qout, c->sq, count, out_count, shadow_count, hit_count, bounce, c->dev_cnt);
    ++g_rt_launch_count;
#endif
}

template <bool ALPHA, bool COUNT>
static int render_frame(rt_context* ctx, FrameSlot* c, rt_scene* s, const FrameParams& P0, const TilePart& tp, uint32_t flags, rt_stream_t st) {
    FrameParams P = P0;
    const uint32_t S = P.ubo.number_of_samples, B = P.ubo.number_of_bounces;
    const uint32_t n_rows = owned_rows(tp), n_local = n_rows * tp.width;
    const size_t per = (size_t)(S ? S : 1) * (B + 1);
    if (per * 5 > c->counters_cap) {
        if (c->counters) rt_free(c->counters);
        c->counters_cap = per * 5 + 64;
        RT_CHECK(dev_alloc(&c->counters, c->counters_cap), "counter allocation");
    }
    uint32_t* qcount = c->counters; uint32_t* scount = c->counters + per; uint32_t* fetch_e = c->counters + 2 * per; uint32_t* fetch_s = c->counters + 3 * per; uint32_t* hcount = c->counters + 4 * per;
    rt_memset(c->counters, 0, per * 5 * sizeof(uint32_t), st);
    if (COUNT) rt_memset(c->dev_cnt, 0, sizeof(RtCounters), st);
    const bool timing = (flags & 4u) != 0;
    c->stage_used = 0;
    const bool simple = s->plain_materials && P.ubo.mapping == RT_MAP_RENDER && P.ubo.debug == 0u;
    const FrameBuffers fb = c->fb; const DScene DS = s->ds;
    for (uint32_t smp = 0; smp < S; ++smp) {
        P.sample = smp;
        const FrameParams Pk = P; const DQueue q0 = c->q[0]; uint32_t* qc0 = qcount + (size_t)smp * (B + 1);
        StageEvent* ev = stage_begin(c, timing, 0, st);
        rt_launch(n_local, st, RT_LAMBDA(size_t i) {
            if (i == 0) *qc0 = n_local;
            raygen_item(Pk, tp, fb, q0, (uint32_t)i, n_rows);
        });
        stage_end(ev, st);
        for (uint32_t b = 0; b < B; ++b) {
            const size_t idx = (size_t)smp * (B + 1) + b;
            const DQueue& qin = c->q[b & 1]; const DQueue& qout = c->q[(b + 1) & 1];
            ev = stage_begin(c, timing, 1, st);
            launch_extend<ALPHA, COUNT>(c, DS, P, qin, qcount + idx, fetch_e + idx, n_local, st);
            stage_end(ev, st);
            ev = stage_begin(c, timing, 2, st);
            if (simple) launch_shade<true, COUNT>(c, DS, P, qin, qout, qcount + idx, qcount + idx + 1, scount + idx, hcount + idx, b, st);
            else launch_shade<false, COUNT>(c, DS, P, qin, qout, qcount + idx, qcount + idx + 1, scount + idx, hcount + idx, b, st);
            stage_end(ev, st);
            if (s->has_nee) {
                ev = stage_begin(c, timing, 3, st);
                launch_shadow<ALPHA, COUNT>(c, DS, P, scount + idx, fetch_s + idx, st);
                stage_end(ev, st);
            }
        }
    }
    {
        const FrameParams Pk = P; const bool no_trace = (S == 0);
        // frames in flight trace concurrently but accumulate in submission order on the shared image
        if (ctx->n_slots > 1 && ctx->acc_pending) ctx->ev_acc.wait(st);
        StageEvent* ev = stage_begin(c, timing, 4, st);
        rt_launch(n_local, st, RT_LAMBDA(size_t i) {
            if (no_trace) { const uint32_t pixel = local_to_pixel(tp, (uint32_t)i); fb.rad[pixel] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); fb.aux[pixel] = make_float2(0.0f, 0.0f); }
            accumulate_item(Pk, tp, fb, (uint32_t)i, false);
        });
        stage_end(ev, st);
        if (ctx->n_slots > 1) { ctx->ev_acc.record(st); ctx->acc_pending = true; }
    }
    c->last_S = S; c->last_B = B; c->last_pixels = (uint64_t)n_local * S; c->last_counted = COUNT; c->last_valid = true;
    return 0;
}

}  // namespace rtcore

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* RT_API(rt_last_error)(void) { return g_err.c_str(); }
const char* RT_API(rt_version)(void) {
#ifdef RT_EMU
    return "rustracer_b200 0.1.0 (host emulation build — tests only)";
#else
    return "rustracer_b200 0.1.0 (CUDA sm_100a)";
#endif
}

void RT_API(rt_context_destroy)(rt_context* c);

int RT_API(rt_context_create)(int device, uint32_t width, uint32_t height, rt_context** out) {
    if (!out || !width || !height) return fail("rt_context_create: bad arguments");
    rt_context* c = new rt_context();
    c->device = device;
#ifndef RT_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { delete c; return fail("rt_context_create: no CUDA device (there is no CPU fallback)"); }
    if (device < 0 || device >= ndev) { delete c; return fail("rt_context_create: device ordinal out of range"); }
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return fail("rt_context_create: cudaSetDevice failed"); }
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, device); g_rt_sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail("rt_context_create: stream creation failed"); }
    c->timers = true;
#endif
    c->ev_submit.create(); c->ev_acc.create(); c->ev_consumer.create(); c->ev_combine.create();
    for (auto& e : c->ticket_done) e.create();
    if (alloc_frame(c, width, height)) { RT_API(rt_context_destroy)(c); return fail(std::string("rt_context_create: allocation failed: ") + rt_platform_error()); }
#ifndef RT_EMU
    {   // CUDA loads a kernel's code on its first launch: run the builder once over a few thousand dummy boxes so that the first
        // real scene build (and its build_ms) does not pay ~7 ms of lazy module loading.  Once per process and device.
        static bool warmed[64] = {false};
        if (device < 64 && !warmed[device]) {
            warmed[device] = true;
            const uint32_t n = 6000;       // > RT_PLOC_FINISH_MAX: per-round kernels and the finishing kernel both run
            DAabb* boxes = nullptr; BuildScratch sc; WideOut wo{}; WideBvhInfo info;
            int e = dev_alloc(&boxes, n); e |= dev_alloc(&wo.nodes, (size_t)n * RT_NODE_F4); e |= dev_alloc(&wo.prim_order, n); e |= dev_alloc(&wo.node_box, n); e |= dev_alloc(&wo.node_parent, n);
            wo.max_nodes = n;
            if (!e) {
                DAabb* b = boxes;
                rt_launch(n, c->stream, RT_LAMBDA(size_t i) { DAabb x; const float f = (float)(i % 97u), g = (float)(i / 97u); x.lo[0] = f; x.lo[1] = g; x.lo[2] = 0.0f; x.hi[0] = f + 0.9f; x.hi[1] = g + 0.9f; x.hi[2] = 0.5f; b[i] = x; });
                build_wide_bvh(boxes, n, sc, wo, c->stream, &info);
                rt_stream_sync(c->stream);
            }
            scratch_free(sc);
            rt_free(boxes); rt_free(wo.nodes); rt_free(wo.prim_order); rt_free(wo.node_box); rt_free(wo.node_parent);
        }
    }
#endif
    *out = c;
    return 0;
}

void RT_API(rt_context_destroy)(rt_context* c) {
    if (!c) return;
    sync_all(c);
    free_frame(c);
    for (FrameSlot& f : c->slot) {
        if (f.counters) rt_free(f.counters);
        if (f.dev_cnt) rt_free(f.dev_cnt);
        for (auto& e : f.stage_events) { e.a.destroy(); e.b.destroy(); }
        f.ev_begin.destroy(); f.ev_end.destroy(); f.done.destroy();
        rt_stream_destroy(f.stream);
    }
    c->ev_submit.destroy(); c->ev_acc.destroy(); c->ev_consumer.destroy(); c->ev_combine.destroy();
    for (auto& e : c->ticket_done) e.destroy();
#ifndef RT_EMU
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->stream) cudaStreamDestroy(c->stream);
#endif
    delete c;
}

int RT_API(rt_frame_resize)(rt_context* c, uint32_t width, uint32_t height) {
    if (!c || !width || !height) return fail("rt_frame_resize: bad arguments");
#ifndef RT_EMU
    cudaSetDevice(c->device);
#endif
    sync_all(c);
    if (width == c->width && height == c->height && c->acc) {
        // same size: drop the accumulation in place.  The allocation (and with it any CUDA IPC mapping a peer holds on the
        // accumulation image, rt_ipc_export) stays valid.
        const size_t n = (size_t)width * height;
        rt_memset(c->acc, 0, n * sizeof(float4), c->stream);     // (snap / display / sync keep their combine state: peers may be mid-protocol)
        for (uint32_t k = 0; k < c->n_slots; ++k) {
            FrameSlot& f = c->slot[k];
            rt_memset(f.fb.out, 0, n * 4, c->stream); rt_memset(f.fb.rad, 0, n * sizeof(float4), c->stream); rt_memset(f.fb.aux, 0, n * sizeof(float2), c->stream);
            f.last_valid = false; f.pending = false;
        }
        c->cur = 0; c->acc_pending = false; c->consumer_pending = false;
        return rt_stream_sync(c->stream) ? fail(std::string("rt_frame_resize: ") + rt_platform_error()) : 0;
    }
    if (c->ipc_exported) return fail("rt_frame_resize: the accumulation image is exported to peers (rt_ipc_export); a size change would free memory they have mapped");
    if (alloc_frame(c, width, height)) return fail(std::string("rt_frame_resize: allocation failed: ") + rt_platform_error());
    return 0;
}

int RT_API(rt_context_set_frames_in_flight)(rt_context* c, uint32_t n) {
    if (!c) return fail("rt_context_set_frames_in_flight: null context");
    if (n < 1 || n > RT_MAX_FRAMES_IN_FLIGHT || RT_TICKET_RING % n) return fail("rt_context_set_frames_in_flight: n must be 1.." + std::to_string(RT_MAX_FRAMES_IN_FLIGHT));
    if (n == c->n_slots) return 0;
#ifndef RT_EMU
    cudaSetDevice(c->device);
#endif
    if (sync_all(c)) return fail(std::string("rt_context_set_frames_in_flight: ") + rt_platform_error());
    // the accumulation image is kept; slots beyond the old count are allocated, surplus ones released
    const size_t npx = (size_t)c->width * c->height;
    if (c->cur >= n) {   // the image of the last frame must stay readable from the slot rt_readback looks at
        rt_d2d(c->slot[0].fb.out, c->slot[c->cur].fb.out, npx * 4, c->stream);
        rt_stream_sync(c->stream);
        c->slot[0].last_valid = false; c->cur = 0;
    }
    for (uint32_t k = n; k < c->n_slots; ++k) free_slot(&c->slot[k]);
    for (uint32_t k = c->n_slots; k < n; ++k)
        if (alloc_slot(c, &c->slot[k], npx)) { for (uint32_t j = c->n_slots; j <= k; ++j) free_slot(&c->slot[j]); return fail(std::string("rt_context_set_frames_in_flight: allocation failed: ") + rt_platform_error()); }
    c->n_slots = n;
    return rt_stream_sync(c->stream) ? fail(std::string("rt_context_set_frames_in_flight: ") + rt_platform_error()) : 0;
}

void RT_API(rt_scene_destroy)(rt_scene* s) {
    if (!s) return;
    sync_all(s->ctx);
    void* ptrs[] = {s->d_vin, s->d_vout, s->d_indices, s->d_prim, s->d_mat, s->d_skins, s->d_dl, s->d_pl, s->d_images, s->d_textures, s->d_lut,
                    s->d_blas_nodes, s->d_tris, s->d_node_box, s->d_node_parent, s->d_prim_order, s->d_leaf_boxes, s->d_prim_boxes, s->d_pending,
                    s->d_tlas_nodes, s->d_tlas_prims, s->d_tlas_box, s->d_tlas_parent, s->d_inst_boxes, s->d_inst_w2o, s->d_inst_o2w, s->d_inst_root, s->d_entry_rec, s->d_bake_src,
                    s->d_tlas_leaf_entry, s->d_tlas_leaf_boxes, s->d_tlas_pending};
    for (void* p : ptrs) if (p) rt_free(p);
    for (uint32_t v = 0; v < s->n_versions; ++v) {      // (the current version's buffers were freed through their aliases above)
        if (v == s->cur_ver) continue;
        rt_free(s->ver[v].vout); rt_free(s->ver[v].tris); rt_free(s->ver[v].blas_nodes); rt_free(s->ver[v].tlas_nodes);
    }
    for (uint8_t* p : s->d_image_px) if (p) rt_free(p);
    for (int f = 0; f < 6; ++f) if (s->d_sky[f]) rt_free(s->d_sky[f]);
    scratch_free(s->scratch);
    s->ev_updated.destroy();
    if (s->upd_timers_created) for (auto& t : s->upd_t) t.destroy();
    delete s;
}

int RT_API(rt_scene_create)(rt_context* c, const rt_scene_desc* d, rt_scene** out) {
    if (!c || !d || !out) return fail("rt_scene_create: null argument");
#ifndef RT_EMU
    cudaSetDevice(c->device);
#endif
    rt_stream_t st = c->stream;
    // ---- validation (the reference panics on malformed input; we return an error) ----
    if (d->n_geometries && (!d->prim_infos || !d->geometries)) return fail("rt_scene_create: geometry arrays missing");
    if ((d->n_vertices && !d->vertices) || (d->n_indices && !d->indices) || (d->n_materials && !d->materials) || (d->n_instances && !d->instances) ||
        (d->n_images && !d->images) || (d->n_samplers && !d->samplers) || (d->n_textures && !d->textures) || (d->n_dlights && !d->dlights) ||
        (d->n_plights && !d->plights) || (d->n_skins && !d->skins))
        return fail("rt_scene_create: an array pointer is null although its count is not zero");
    for (uint32_t i = 0; i < d->n_images; ++i) if (!d->images[i].rgba8 || !d->images[i].width || !d->images[i].height) return fail("rt_scene_create: empty image");
    if (!d->n_materials) return fail("rt_scene_create: at least one material is required");
    for (uint32_t g = 0; g < d->n_geometries; ++g) {
        const rt_prim_info& pi = d->prim_infos[g]; const rt_geometry& ge = d->geometries[g];
        if (ge.i_len % 3) return fail("rt_scene_create: geometry index count not a multiple of 3");
        if ((uint64_t)pi.i_offset + ge.i_len > d->n_indices || (uint64_t)pi.v_offset + ge.v_len > d->n_vertices) return fail("rt_scene_create: geometry range outside vertex/index arrays");
        if (pi.material_id >= d->n_materials) return fail("rt_scene_create: material_id out of range");
        for (uint32_t k = 0; k < ge.i_len; ++k) if (d->indices[pi.i_offset + k] >= ge.v_len) return fail("rt_scene_create: vertex index out of range");
    }
    for (uint32_t i = 0; i < d->n_instances; ++i) if (d->instances[i].geo_id >= d->n_geometries) return fail("rt_scene_create: instance geo_id out of range");
    for (uint32_t t = 0; t < d->n_textures; ++t) if (d->textures[t].image_index >= d->n_images || d->textures[t].sampler_index >= d->n_samplers) return fail("rt_scene_create: texture refers to a missing image or sampler");
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const rt_material& mt = d->materials[m];
        const rt_texture_info* tis[] = {&mt.base_color_texture, &mt.metallic_roughness_texture, &mt.normal_texture, &mt.emissive_texture, &mt.transmission_texture,
                                         &mt.specular_texture, &mt.specular_color_texture, &mt.sg_diffuse_texture, &mt.sg_specular_glossiness_texture};
        for (const rt_texture_info* ti : tis) if (ti->index >= 0 && (uint32_t)ti->index >= d->n_textures) return fail("rt_scene_create: material texture index out of range");
    }

    rt_scene* s = new rt_scene(); s->ctx = c;
    auto bail = [&](const std::string& m) { RT_API(rt_scene_destroy)(s); return fail(m); };
    s->n_vertices = d->n_vertices; s->n_indices = d->n_indices; s->n_materials = d->n_materials; s->n_skins = d->n_skins;
    s->geometries.assign(d->geometries, d->geometries + d->n_geometries);
    s->prim_infos.assign(d->prim_infos, d->prim_infos + d->n_geometries);
    s->instances.assign(d->instances, d->instances + d->n_instances);
    int e = 0;
    e |= dev_upload(&s->d_vin, d->vertices, d->n_vertices, st); e |= dev_alloc(&s->d_vout, d->n_vertices ? d->n_vertices : 1);
    e |= dev_upload(&s->d_indices, d->indices, d->n_indices, st); e |= dev_upload(&s->d_prim, d->prim_infos, d->n_geometries, st);
    e |= dev_upload(&s->d_mat, d->materials, d->n_materials, st);
    e |= dev_upload(&s->d_skins, d->skins, (size_t)d->n_skins * RT_MAX_JOINTS * 16, st);
    s->cap_dl = d->n_dlights; s->cap_pl = d->n_plights;
    e |= dev_upload(&s->d_dl, d->dlights, d->n_dlights, st); e |= dev_upload(&s->d_pl, d->plights, d->n_plights, st);
    if (e) return bail(std::string("rt_scene_create: upload failed: ") + rt_platform_error());
    s->ds.n_dlights = d->n_dlights; s->ds.n_plights = d->n_plights; compute_has_nee(s, d->plights, d->n_plights);
    // textures
    std::vector<DImage> dimg(d->n_images); s->d_image_px.assign(d->n_images, nullptr);
    for (uint32_t i = 0; i < d->n_images; ++i) {
        const rt_image_desc& im = d->images[i];
        if (dev_upload(&s->d_image_px[i], im.rgba8, (size_t)im.width * im.height * 4, st)) return bail("rt_scene_create: image upload failed");
        dimg[i].px = s->d_image_px[i]; dimg[i].w = im.width; dimg[i].h = im.height; dimg[i].srgb = im.srgb; dimg[i]._pad = 0;
    }
    std::vector<DTexture> dtex(d->n_textures);
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        const rt_sampler_desc& sm = d->samplers[d->textures[t].sampler_index];
        dtex[t].image = d->textures[t].image_index; dtex[t].mag_filter = sm.mag_filter; dtex[t].wrap_s = sm.wrap_s; dtex[t].wrap_t = sm.wrap_t;
    }
    float lut[256];
    for (int i = 0; i < 256; ++i) { const double cc = i / 255.0; lut[i] = (float)(cc <= 0.04045 ? cc / 12.92 : pow((cc + 0.055) / 1.055, 2.4)); }
    e |= dev_upload(&s->d_images, dimg.data(), dimg.size(), st); e |= dev_upload(&s->d_textures, dtex.data(), dtex.size(), st); e |= dev_upload(&s->d_lut, lut, 256, st);
    if (e) return bail("rt_scene_create: texture table upload failed");
    s->ds.n_textures = d->n_textures;
    s->plain_materials = true;
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const rt_material& mt = d->materials[m];
        const rt_texture_info* tis[] = {&mt.base_color_texture, &mt.metallic_roughness_texture, &mt.normal_texture, &mt.emissive_texture, &mt.transmission_texture,
                                         &mt.specular_texture, &mt.specular_color_texture, &mt.sg_diffuse_texture, &mt.sg_specular_glossiness_texture};
        for (const rt_texture_info* ti : tis) if (ti->index >= 0) s->plain_materials = false;
        if (mt.workflow == 1u) s->plain_materials = false;
    }
    if (d->skybox_faces[0] && d->skybox_width && upload_sky(s, d->skybox_faces, d->skybox_width, d->skybox_height, d->skybox_srgb)) return bail(g_err);

    // ---- geometry records, baking classification and BVH storage ----
    classify_instances(s);
    uint64_t node_off = 0, tri_off = 0; uint32_t max_tris = 0;
    s->geo.resize(d->n_geometries);
    for (uint32_t g = 0; g < d->n_geometries; ++g) {
        GeoRecord& gr = s->geo[g];
        gr.n_tris = d->geometries[g].i_len / 3; gr.node_off = 0; gr.tri_off = 0; gr.n_nodes = 0; gr.depth = 0; gr.skinned = false; gr.needed = false;
        for (uint32_t v = 0; v < d->geometries[g].v_len && !gr.skinned; ++v) if (d->vertices[d->prim_infos[g].v_offset + v].skin_index >= 0) gr.skinned = true;
    }
    std::vector<uint2> bake_src;
    s->merged_skinned = false;
    for (uint32_t i = 0; i < d->n_instances; ++i) {
        const uint32_t g = d->instances[i].geo_id;
        if (!s->baked[i]) { s->geo[g].needed = true; continue; }
        for (uint32_t p = 0; p < s->geo[g].n_tris; ++p) bake_src.push_back(make_uint2(i, p));
        if (s->geo[g].skinned) s->merged_skinned = true;
    }
    for (uint32_t g = 0; g < d->n_geometries; ++g) {
        GeoRecord& gr = s->geo[g];
        if (!gr.needed) continue;
        gr.node_off = (uint32_t)node_off; gr.tri_off = (uint32_t)tri_off;
        node_off += gr.n_tris ? gr.n_tris : 1; tri_off += gr.n_tris; if (gr.n_tris > max_tris) max_tris = gr.n_tris;
    }
    s->merged = GeoRecord{};
    s->merged.n_tris = (uint32_t)bake_src.size(); s->merged.node_off = (uint32_t)node_off; s->merged.tri_off = (uint32_t)tri_off; s->merged.needed = true; s->merged.skinned = s->merged_skinned;
    node_off += s->merged.n_tris ? s->merged.n_tris : 1; tri_off += s->merged.n_tris; if (s->merged.n_tris > max_tris) max_tris = s->merged.n_tris;
    if (node_off > 0xFFFFFFF0ull || tri_off >= (1ull << 27) || bake_src.size() >= (1ull << 27))
        return bail("rt_scene_create: more than 2^27 unique (baked + instanced) triangles are not supported");
    s->total_nodes = node_off; s->total_tris = tri_off;
    const uint32_t ninst = d->n_instances;
    e = 0;
    e |= dev_alloc(&s->d_blas_nodes, (size_t)(node_off ? node_off : 1) * RT_NODE_F4); e |= dev_alloc(&s->d_tris, (size_t)(tri_off ? tri_off : 1) * RT_TRI_F4);
    e |= dev_alloc(&s->d_node_box, node_off ? node_off : 1); e |= dev_alloc(&s->d_node_parent, node_off ? node_off : 1); e |= dev_alloc(&s->d_pending, node_off ? node_off : 1);
    e |= dev_alloc(&s->d_prim_order, tri_off ? tri_off : 1); e |= dev_alloc(&s->d_leaf_boxes, tri_off ? tri_off : 1); e |= dev_alloc(&s->d_prim_boxes, max_tris ? max_tris : 1);
    e |= dev_upload(&s->d_bake_src, bake_src.data(), bake_src.size(), st);
    e |= dev_alloc(&s->d_tlas_nodes, (size_t)(ninst + 1) * RT_NODE_F4); e |= dev_alloc(&s->d_tlas_prims, ninst + 1);
    e |= dev_alloc(&s->d_tlas_box, ninst + 1); e |= dev_alloc(&s->d_tlas_parent, ninst + 1); e |= dev_alloc(&s->d_inst_boxes, ninst + 1);
    e |= dev_alloc(&s->d_tlas_leaf_entry, ninst + 1); e |= dev_alloc(&s->d_tlas_leaf_boxes, ninst + 1); e |= dev_alloc(&s->d_tlas_pending, ninst + 1);
    e |= dev_alloc(&s->d_inst_w2o, (size_t)(ninst + 1) * RT_INST_F4); e |= dev_alloc(&s->d_inst_o2w, (size_t)(ninst ? ninst : 1) * RT_O2W_F4);
    e |= dev_alloc(&s->d_inst_root, ninst + 1); e |= dev_alloc(&s->d_entry_rec, ninst + 1);
    if (e) return bail(std::string("rt_scene_create: BVH allocation failed: ") + rt_platform_error());

    // builder scratch is reserved up front: the build below then runs without allocating (and build_ms times the build, not cudaMalloc)
    if (scratch_reserve(s->scratch, (size_t)(max_tris > ninst + 1 ? max_tris : ninst + 1))) return bail(std::string("rt_scene_create: scratch allocation failed: ") + rt_platform_error());
    rt_timer t0, t1, t2; t0.create(); t1.create(); t2.create();
    t0.record(st);
    run_skinning(s);   // initial ComputeUnit::dispatch (main.rs:85-91); copy-through when nothing is skinned
    if (upload_instance_records(s)) { t0.destroy(); t1.destroy(); t2.destroy(); return bail(g_err); }
    for (uint32_t g = 0; g < d->n_geometries; ++g) {
        if (!s->geo[g].needed) continue;
        if (build_blas(s, g)) { t0.destroy(); t1.destroy(); t2.destroy(); return bail(g_err); }
    }
    if (build_merged(s)) { t0.destroy(); t1.destroy(); t2.destroy(); return bail(g_err); }
    t1.record(st);
    if (build_tlas(s)) { t0.destroy(); t1.destroy(); t2.destroy(); return bail(g_err); }
    t2.record(st);
    if (rt_stream_sync(st)) { t0.destroy(); t1.destroy(); t2.destroy(); return bail(std::string("rt_scene_create: device error: ") + rt_platform_error()); }
    s->build_ms = rt_timer_ms(t0, t1); s->tlas_ms = rt_timer_ms(t1, t2);
    t0.destroy(); t1.destroy(); t2.destroy();
    // skinned scenes are double-buffered so that a skin update overlaps the frames still reading the previous pose
    s->ver[0].vout = s->d_vout; s->ver[0].tris = s->d_tris; s->ver[0].blas_nodes = s->d_blas_nodes; s->ver[0].tlas_nodes = s->d_tlas_nodes;
    s->n_versions = 1; s->cur_ver = 0;
    if (d->n_skins) {
        if (alloc_versions(s, 2)) return bail(std::string("rt_scene_create: allocation failed: ") + rt_platform_error());
        s->n_versions = 2;
        if (replicate_versions(s) || rt_stream_sync(st)) return bail(std::string("rt_scene_create: device error: ") + rt_platform_error());
    }
    update_ds(s);
    *out = s;
    return 0;
}

int RT_API(rt_scene_update_instances)(rt_scene* s, const rt_instance* inst, uint32_t n) {
    if (!s || !inst) return fail("rt_scene_update_instances: null argument");
    if (n != s->instances.size()) return fail("rt_scene_update_instances: instance count differs from the scene's");
    for (uint32_t i = 0; i < n; ++i) if (inst[i].geo_id >= s->geo.size()) return fail("rt_scene_update_instances: geo_id out of range");
    for (uint32_t i = 0; i < n; ++i) if (inst[i].geo_id != s->instances[i].geo_id) return fail("rt_scene_update_instances: geo_id of an instance may not change");
    s->instances.assign(inst, inst + n);
    sync_all(s->ctx);            // frames in flight (or on a caller's stream) still read the structures rebuilt below
    rt_timer t0, t1; t0.create(); t1.create(); t0.record(s->ctx->stream);
    if (upload_instance_records(s)) { t0.destroy(); t1.destroy(); return 1; }
    if (s->merged.n_tris) refit_merged(s);        // baked instances moved: re-bake their triangles, refit the merged BLAS
    int e = build_tlas(s);
    if (!e) e = replicate_versions(s);
    t1.record(s->ctx->stream); rt_stream_sync(s->ctx->stream); s->tlas_ms = rt_timer_ms(t0, t1); t0.destroy(); t1.destroy();
    if (e) return 1;
    update_ds(s);
    return 0;
}

// stage timers of the last skin update are read on demand (rt_scene_bvh_info): the update itself never blocks the host
static void collect_update_timers(rt_scene* s, bool block = true) {
    if (!s->upd_timers_pending) return;
    if (!block && !s->upd_t[3].ready()) return;   // the previous update is still running: its timings are simply dropped
    if (block) rt_stream_sync(s->ctx->stream);
    s->skin_ms = rt_timer_ms(s->upd_t[0], s->upd_t[1]); s->refit_ms = rt_timer_ms(s->upd_t[1], s->upd_t[2]); s->tlas_ms = rt_timer_ms(s->upd_t[2], s->upd_t[3]);
    s->upd_timers_pending = false;
}

int RT_API(rt_scene_update_skins)(rt_scene* s, const float* mats, uint32_t n_skins, int rebuild) {
    if (!s || (!mats && n_skins)) return fail("rt_scene_update_skins: null argument");
    if (n_skins != s->n_skins) return fail("rt_scene_update_skins: skin count differs from the scene's");
    rt_stream_t st = s->ctx->stream;
#ifndef RT_EMU
    cudaSetDevice(s->ctx->device);
#endif
    // Frames in flight (or on a caller's stream) still read the vertices / BVH rewritten below: the scene's stream
    // waits for them on the device.  A refit never blocks the host: skinning, BLAS refit and TLAS refit are queued,
    // later frames wait for ev_updated.  A rebuild reads builder counters back level by level and is synchronous.
    if (rebuild || s->n_versions == 1) join_frames(s->ctx, st);
    if (rebuild) sync_all(s->ctx);
    if (!rebuild && s->n_versions > 1) {
        // write the next copy; only the frames that still read that copy have to finish first
        const uint32_t next = (s->cur_ver + 1) % s->n_versions;
        rt_context* c = s->ctx;
        for (uint32_t k = 0; k < c->n_slots; ++k)
            if (c->slot[k].pending && c->slot[k].scene == (const void*)s && c->slot[k].scene_version == next) c->slot[k].done.wait(st);
        use_version(s, next);
    }
    if (!s->upd_timers_created) { for (auto& t : s->upd_t) t.create(); s->ev_updated.create(); s->upd_timers_created = true; }
    collect_update_timers(s, false);     // (the events are about to be re-recorded; never blocks the host)
    RT_CHECK(rt_h2d(s->d_skins, mats, (size_t)n_skins * RT_MAX_JOINTS * 16 * 4, st), "skin upload");
    s->upd_t[0].record(st);
    run_skinning(s);
    s->upd_t[1].record(st);
    for (uint32_t g = 0; g < s->geo.size(); ++g) {
        if (!s->geo[g].skinned || !s->geo[g].needed) continue;
        if (rebuild) { if (build_blas(s, g)) return 1; } else refit_blas(s, g);
    }
    if (s->merged_skinned && s->merged.n_tris) { if (rebuild) { if (build_merged(s)) return 1; } else refit_merged(s); }
    s->upd_t[2].record(st);
    const int e = rebuild ? build_tlas(s) : refit_tlas(s);
    s->upd_t[3].record(st);
    s->upd_timers_pending = true;
    s->ev_updated.record(st); s->update_pending = true;
    if (e) return 1;
    if (rebuild) {
        if (replicate_versions(s) || rt_stream_sync(st)) return fail(std::string("rt_scene_update_skins: ") + rt_platform_error());
    }
    update_ds(s);
    return 0;
}

int RT_API(rt_scene_set_versions)(rt_scene* s, uint32_t n) {
    if (!s) return fail("rt_scene_set_versions: null scene");
    if (n < 1 || n > RT_MAX_SCENE_VERSIONS) return fail("rt_scene_set_versions: n must be 1..4");
    if (n == s->n_versions) return 0;
    if (sync_all(s->ctx)) return fail(std::string("rt_scene_set_versions: ") + rt_platform_error());
    if (n < s->n_versions) {
        // keep the current copy as version 0
        if (s->cur_ver != 0) { std::swap(s->ver[0], s->ver[s->cur_ver]); use_version(s, 0); }
        for (uint32_t v = n; v < s->n_versions; ++v) { rt_free(s->ver[v].vout); rt_free(s->ver[v].tris); rt_free(s->ver[v].blas_nodes); rt_free(s->ver[v].tlas_nodes); s->ver[v] = rt_scene::Version{}; }
        s->n_versions = n;
    } else {
        const uint32_t old_n = s->n_versions;
        if (alloc_versions(s, n)) {
            for (uint32_t v = old_n; v < n; ++v) { rt_free(s->ver[v].vout); rt_free(s->ver[v].tris); rt_free(s->ver[v].blas_nodes); rt_free(s->ver[v].tlas_nodes); s->ver[v] = rt_scene::Version{}; }
            return fail(std::string("rt_scene_set_versions: allocation failed: ") + rt_platform_error());
        }
        s->n_versions = n;
        if (replicate_versions(s) || rt_stream_sync(s->ctx->stream)) return fail(std::string("rt_scene_set_versions: ") + rt_platform_error());
    }
    for (uint32_t k = 0; k < RT_MAX_FRAMES_IN_FLIGHT; ++k) if (s->ctx->slot[k].scene == (const void*)s) s->ctx->slot[k].scene_version = s->cur_ver;
    update_ds(s);
    return 0;
}

int RT_API(rt_scene_update_lights)(rt_scene* s, const rt_light* dl, uint32_t ndl, const rt_light* pl, uint32_t npl) {
    if (!s) return fail("rt_scene_update_lights: null scene");
    rt_stream_t st = s->ctx->stream;
    sync_all(s->ctx);
    if (ndl > s->cap_dl) { rt_free(s->d_dl); s->d_dl = nullptr; RT_CHECK(dev_alloc(&s->d_dl, ndl), "light allocation"); s->cap_dl = ndl; }
    if (npl > s->cap_pl) { rt_free(s->d_pl); s->d_pl = nullptr; RT_CHECK(dev_alloc(&s->d_pl, npl), "light allocation"); s->cap_pl = npl; }
    if (ndl) RT_CHECK(rt_h2d(s->d_dl, dl, (size_t)ndl * sizeof(rt_light), st), "light upload");
    if (npl) RT_CHECK(rt_h2d(s->d_pl, pl, (size_t)npl * sizeof(rt_light), st), "light upload");
    s->ds.n_dlights = ndl; s->ds.n_plights = npl; compute_has_nee(s, pl, npl);
    update_ds(s);
    return 0;
}

int RT_API(rt_scene_set_skybox)(rt_scene* s, const uint8_t* const faces[6], uint32_t w, uint32_t h, uint32_t srgb) {
    if (!s || !faces || !w || !h) return fail("rt_scene_set_skybox: bad arguments");
    for (int f = 0; f < 6; ++f) if (!faces[f]) return fail("rt_scene_set_skybox: null face");
    sync_all(s->ctx);
    return upload_sky(s, faces, w, h, srgb);
}

int RT_API(rt_render)(rt_context* c, rt_scene* s, const rt_ubo* ubo, const rt_render_opts* opts, void* stream) {
    if (!c || !s || !ubo) return fail("rt_render: null argument");
    if (s->ctx->device != c->device) return fail("rt_render: scene lives on another device");   // contexts of one device may share a scene (read-only here)
    if (ubo->total_number_of_samples == 0) return fail("rt_render: total_number_of_samples must be > 0");
#ifndef RT_EMU
    cudaSetDevice(c->device);
#endif
    rt_stream_t st = stream ? (rt_stream_t)stream : c->stream;
    c->last_stream = st;
    FrameParams P; P.ubo = *ubo; P.width = c->width; P.height = c->height; P.sample = 0;
    P.clk = host_tea16(ubo->total_number_of_samples, ubo->random_seed);   // D1
    TilePart tp; tp.width = c->width; tp.height = c->height; tp.strip_rows = 1; tp.n_parts = 1; tp.part = 0;
    uint32_t flags = 0;
    if (opts) {
        flags = opts->flags;
        if (opts->n_parts > 1) {
            if (!opts->strip_rows || opts->part >= opts->n_parts) return fail("rt_render: bad tile partition");
            tp.strip_rows = opts->strip_rows; tp.n_parts = opts->n_parts; tp.part = opts->part;
        }
    }
    // frames in flight: the frame runs on its slot's stream, after everything queued on `st` so far and after the
    // last consumer of the accumulation image; with a single slot it runs on `st` itself (strict stream order)
    const uint32_t k = (uint32_t)(c->frame_seq % c->n_slots);
    FrameSlot* f = &c->slot[k];
    if (c->n_slots > 1) {
        c->ev_submit.record(st); c->ev_submit.wait(f->stream);
        st = f->stream;
    }
    if (c->consumer_pending) c->ev_consumer.wait(st);   // (a no-op when the consumer ran on this very stream)
    if (s->update_pending) s->ev_updated.wait(st);     // asynchronous skin update queued on the scene's stream
#ifndef RT_EMU
    f->launches_before = g_rt_launch_count;
#endif
    if (c->timers) f->ev_begin.record(st);
    const bool alpha = !ubo->fully_opaque, count = (flags & RT_RENDER_COUNTERS) != 0;
    int e;
    if (alpha) e = count ? render_frame<true, true>(c, f, s, P, tp, flags, st) : render_frame<true, false>(c, f, s, P, tp, flags, st);
    else e = count ? render_frame<false, true>(c, f, s, P, tp, flags, st) : render_frame<false, false>(c, f, s, P, tp, flags, st);
    if (c->timers) f->ev_end.record(st);
    f->done.record(st); f->pending = true;
    c->ticket_done[c->frame_seq % RT_TICKET_RING].record(st);
    f->scene = s; f->scene_version = s->cur_ver;
    c->cur = k; ++c->frame_seq;
#ifndef RT_EMU
    f->launches_after = g_rt_launch_count;
    if (!e && cudaPeekAtLastError() != cudaSuccess) return fail(std::string("rt_render: launch failed: ") + rt_platform_error());
#endif
    return e;
}

int RT_API(rt_tonemap)(rt_context* c, const rt_ubo* ubo, void* stream) {
    if (!c || !ubo) return fail("rt_tonemap: null argument");
    rt_stream_t st = stream ? (rt_stream_t)stream : c->stream;
    c->last_stream = st;
    join_frames(c, st);
    FrameParams P; P.ubo = *ubo; P.width = c->width; P.height = c->height; P.sample = 0; P.clk = 0;
    TilePart tp; tp.width = c->width; tp.height = c->height; tp.strip_rows = 1; tp.n_parts = 1; tp.part = 0;
    const FrameBuffers fb = c->slot[c->cur].fb;
    rt_launch((size_t)c->width * c->height, st, RT_LAMBDA(size_t i) { accumulate_item(P, tp, fb, (uint32_t)i, true); });
    consumer_ran(c, st);
    return 0;
}

int RT_API(rt_synchronize)(rt_context* c) {
    if (!c) return fail("rt_synchronize: null context");
    if (sync_all(c)) return fail(std::string("rt_synchronize: ") + rt_platform_error());
    return 0;
}

int RT_API(rt_readback)(rt_context* c, float* acc, uint8_t* out) {
    if (!c) return fail("rt_readback: null context");
    const size_t n = (size_t)c->width * c->height;
    if (sync_all(c)) return fail(std::string("rt_readback: ") + rt_platform_error());
    if (acc) RT_CHECK(rt_d2h(acc, c->acc, n * 16, c->stream), "rt_readback");
    if (out) RT_CHECK(rt_d2h(out, c->slot[c->cur].fb.out, n * 4, c->stream), "rt_readback");
    if (rt_stream_sync(c->stream)) return fail(std::string("rt_readback: ") + rt_platform_error());
    return 0;
}

int RT_API(rt_upload_accumulation)(rt_context* c, const float* acc) {
    if (!c || !acc) return fail("rt_upload_accumulation: null argument");
    if (sync_all(c)) return fail(std::string("rt_upload_accumulation: ") + rt_platform_error());
    RT_CHECK(rt_h2d(c->acc, acc, (size_t)c->width * c->height * 16, c->stream), "rt_upload_accumulation");
    rt_stream_sync(c->stream);
    return 0;
}

int RT_API(rt_device_ptrs)(rt_context* c, void** acc, void** out) {
    if (!c) return fail("rt_device_ptrs: null context");
    if (acc) *acc = c->acc;
    if (out) *out = c->slot[c->cur].fb.out;
    return 0;
}

int RT_API(rt_join)(rt_context* c, void* stream) {
    if (!c) return fail("rt_join: null context");
    rt_stream_t st = stream ? (rt_stream_t)stream : c->stream;
    join_frames(c, st);
    if (c->combine_pending && c->combine_stream != st) c->ev_combine.wait(st);      // a combine queued on a side stream
    return 0;
}

int RT_API(rt_readback_async)(rt_context* c, uint8_t* out, uint64_t* ticket) {
    if (!c || !out) return fail("rt_readback_async: null argument");
    if (c->frame_seq == 0) return fail("rt_readback_async: no frame submitted yet");
    FrameSlot* f = &c->slot[c->cur];
    rt_stream_t st = c->n_slots > 1 ? f->stream : (c->last_stream ? c->last_stream : c->stream);
    RT_CHECK(rt_d2h(out, f->fb.out, (size_t)c->width * c->height * 4, st), "rt_readback_async");
    f->done.record(st); f->pending = true;
    c->ticket_done[(c->frame_seq - 1) % RT_TICKET_RING].record(st);
    if (ticket) *ticket = c->frame_seq - 1;
    return 0;
}

int RT_API(rt_frame_wait)(rt_context* c, uint64_t ticket) {
    if (!c) return fail("rt_frame_wait: null context");
    if (ticket >= c->frame_seq) return fail("rt_frame_wait: ticket of a frame that was never submitted");
    // older than the ring: its entry now belongs to a later frame of the same slot, which completes after it
    if (c->ticket_done[ticket % RT_TICKET_RING].sync()) return fail(std::string("rt_frame_wait: ") + rt_platform_error());
    return 0;
}

int RT_API(rt_last_frame_stats)(rt_context* c, rt_stats* o) {
    if (!c || !o) return fail("rt_last_frame_stats: null argument");
    memset(o, 0, sizeof *o);
    FrameSlot* f = &c->slot[c->cur];
    if (!f->last_valid) return fail("rt_last_frame_stats: no frame rendered yet");
    if (sync_all(c)) return fail(std::string("rt_last_frame_stats: ") + rt_platform_error());
    const size_t per = (size_t)(f->last_S ? f->last_S : 1) * (f->last_B + 1);
    std::vector<uint32_t> h(per * 5);
    RT_CHECK(rt_d2h(h.data(), f->counters, per * 5 * 4, c->stream), "rt_last_frame_stats");
    rt_stream_sync(c->stream);
    for (uint32_t smp = 0; smp < f->last_S; ++smp)
        for (uint32_t b = 0; b < f->last_B; ++b) { o->rays_extend += h[(size_t)smp * (f->last_B + 1) + b]; o->rays_shadow += h[per + (size_t)smp * (f->last_B + 1) + b]; o->shaded_hits += h[4 * per + (size_t)smp * (f->last_B + 1) + b]; }
    o->pixel_samples = f->last_pixels;
    if (f->last_counted) {
        RtCounters k; RT_CHECK(rt_d2h(&k, f->dev_cnt, sizeof k, c->stream), "rt_last_frame_stats"); rt_stream_sync(c->stream);
        o->nodes = k.nodes; o->tris = k.tris; o->insts = k.insts; o->anyhits = k.anyhits; o->tex_taps = k.tex_taps; o->light_cands = k.light_cands;
    }
    if (c->timers) o->ms_total = rt_timer_ms(f->ev_begin, f->ev_end);
    for (size_t i = 0; i < f->stage_used; ++i) {
        StageEvent& e = f->stage_events[i]; const float ms = rt_timer_ms(e.a, e.b);
        switch (e.stage) { case 0: o->ms_raygen += ms; break; case 1: o->ms_extend += ms; o->n_extend_launches++; break; case 2: o->ms_shade += ms; break; case 3: o->ms_shadow += ms; break; default: o->ms_accum += ms; }
    }
    o->n_kernel_launches = (uint32_t)(f->launches_after - f->launches_before);
    return 0;
}

static int trace_common(rt_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, rt_hit* hits, uint8_t* occ) {
    if (!s || (!rays && n)) return fail("rt_trace: null argument");
    if (n == 0) return 0;
    rt_context* c = s->ctx; rt_stream_t st = c->stream;
#ifndef RT_EMU
    cudaSetDevice(c->device);
#endif
    rt_ray* d_rays = nullptr; uint32_t* d_rng = nullptr; rt_hit* d_hits = nullptr; uint8_t* d_occ = nullptr;
    int e = dev_upload(&d_rays, rays, n, st);
    if (rng4) e |= dev_upload(&d_rng, rng4, (size_t)n * 4, st);
    if (hits) e |= dev_alloc(&d_hits, n); else e |= dev_alloc(&d_occ, n);
    if (e) { rt_free(d_rays); rt_free(d_rng); rt_free(d_hits); rt_free(d_occ); return fail(std::string("rt_trace: allocation failed: ") + rt_platform_error()); }
    const DScene DS = s->ds; const bool alpha = !(flags & RT_TRACE_OPAQUE);
#ifdef RT_EMU
    for (uint32_t i = 0; i < n; ++i) {
        u4 rng; rng.x = rng.y = rng.z = rng.w = 0;
        if (d_rng) { rng.x = d_rng[4 * i]; rng.y = d_rng[4 * i + 1]; rng.z = d_rng[4 * i + 2]; rng.w = d_rng[4 * i + 3]; }
        const rt_ray& r = d_rays[i]; RtHit h; bool f;
        const f3 o = mk3(r.origin[0], r.origin[1], r.origin[2]), dd = mk3(r.direction[0], r.direction[1], r.direction[2]);
        if (d_occ) { f = alpha ? trace_ray<RT_MODE_ANY, true, false>(DS, o, dd, r.tmin, r.tmax, rng, h, nullptr) : trace_ray<RT_MODE_ANY, false, false>(DS, o, dd, r.tmin, r.tmax, rng, h, nullptr); d_occ[i] = f; }
        else {
            f = alpha ? trace_ray<RT_MODE_CLOSEST, true, false>(DS, o, dd, r.tmin, r.tmax, rng, h, nullptr) : trace_ray<RT_MODE_CLOSEST, false, false>(DS, o, dd, r.tmin, r.tmax, rng, h, nullptr);
            rt_hit& oh = d_hits[i];
            if (f) { oh.t = h.t; oh.u = h.u; oh.v = h.v; oh.instance_id = h.inst; oh.primitive_id = h.prim; oh.geo_id = rt_float_as_uint(DS.inst_w2o[(size_t)h.inst * RT_INST_F4 + 3].y); }
            else { oh.t = -1.0f; oh.u = oh.v = 0.0f; oh.instance_id = oh.primitive_id = oh.geo_id = 0xFFFFFFFFu; }
        }
    }
#else
    if (flags & RT_TRACE_SCALAR) {      // one thread per ray, plain while-while loop (cross-check of the wavefront path)
        const unsigned grid = persistent_grid(8);
        if (alpha) trace_rays_kernel<true, false><<<grid, 128, 0, st>>>(DS, d_rays, n, d_rng, d_hits, d_occ, nullptr);
        else trace_rays_kernel<false, false><<<grid, 128, 0, st>>>(DS, d_rays, n, d_rng, d_hits, d_occ, nullptr);
    } else {                            // the persistent traversal the frame kernels run
        uint32_t* d_fetch = nullptr;
        if (dev_alloc(&d_fetch, 1)) { rt_free(d_rays); rt_free(d_rng); rt_free(d_hits); rt_free(d_occ); return fail(std::string("rt_trace: allocation failed: ") + rt_platform_error()); }
        rt_memset(d_fetch, 0, 4, st);
        const unsigned grid = trace_grid();
#define RT_TW(MODE, A, SG) trace_wavefront_kernel<MODE, A, SG><<<grid, RT_EXTEND_THREADS, 0, st>>>(DS, d_rays, n, d_rng, d_hits, d_occ, d_fetch)
        const bool single = DS.single_merged != 0u;
        if (d_occ) { if (alpha) { if (single) RT_TW(RT_MODE_ANY, true, true); else RT_TW(RT_MODE_ANY, true, false); } else { if (single) RT_TW(RT_MODE_ANY, false, true); else RT_TW(RT_MODE_ANY, false, false); } }
        else { if (alpha) { if (single) RT_TW(RT_MODE_CLOSEST, true, true); else RT_TW(RT_MODE_CLOSEST, true, false); } else { if (single) RT_TW(RT_MODE_CLOSEST, false, true); else RT_TW(RT_MODE_CLOSEST, false, false); } }
#undef RT_TW
        rt_stream_sync(st); rt_free(d_fetch);
    }
    ++g_rt_launch_count;
#endif
    if (hits) e = rt_d2h(hits, d_hits, (size_t)n * sizeof(rt_hit), st); else e = rt_d2h(occ, d_occ, n, st);
    e |= rt_stream_sync(st);
    rt_free(d_rays); rt_free(d_rng); rt_free(d_hits); rt_free(d_occ);
    if (e) return fail(std::string("rt_trace: device error: ") + rt_platform_error());
    return 0;
}

int RT_API(rt_trace_closest)(rt_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, rt_hit* hits) {
    if (!hits && n) return fail("rt_trace_closest: null hits");
    return trace_common(s, rays, n, flags, rng4, hits, nullptr);
}
int RT_API(rt_trace_any)(rt_scene* s, const rt_ray* rays, uint32_t n, uint32_t flags, const uint32_t* rng4, uint8_t* occluded) {
    if (!occluded && n) return fail("rt_trace_any: null output");
    return trace_common(s, rays, n, flags, rng4, nullptr, occluded);
}

int RT_API(rt_scene_read_vertices)(rt_scene* s, rt_vertex* out, uint32_t n) {
    if (!s || !out) return fail("rt_scene_read_vertices: null argument");
    if (n > s->n_vertices) return fail("rt_scene_read_vertices: more vertices requested than the scene holds");
    rt_stream_sync(s->ctx->stream);
    RT_CHECK(rt_d2h(out, s->d_vout, (size_t)n * sizeof(rt_vertex), s->ctx->stream), "rt_scene_read_vertices");
    rt_stream_sync(s->ctx->stream);
    return 0;
}

int RT_API(rt_scene_read_nodes)(rt_scene* s, int geo, float* out, uint32_t max_nodes, uint32_t* n_nodes) {
    if (!s || !n_nodes) return fail("rt_scene_read_nodes: null argument");
    if (geo >= (int)s->geo.size()) return fail("rt_scene_read_nodes: geometry index out of range");
    GeoRecord tl{}; tl.needed = true; tl.n_nodes = s->tlas_nodes;          // geo == -2: the TLAS
    const GeoRecord& gr = geo == -2 ? tl : (geo < 0 ? s->merged : s->geo[geo]);
    *n_nodes = gr.needed ? gr.n_nodes : 0u;
    if (!out || !*n_nodes) return 0;
    if (max_nodes < *n_nodes) return fail("rt_scene_read_nodes: buffer too small");
    sync_all(s->ctx);
    const float4* src = geo == -2 ? s->d_tlas_nodes : s->d_blas_nodes + (size_t)gr.node_off * RT_NODE_F4;
    RT_CHECK(rt_d2h(out, src, (size_t)gr.n_nodes * RT_NODE_F4 * sizeof(float4), s->ctx->stream), "rt_scene_read_nodes");
    rt_stream_sync(s->ctx->stream);
    return 0;
}

int RT_API(rt_scene_bvh_info)(rt_scene* s, rt_bvh_info* o) {
    if (!s || !o) return fail("rt_scene_bvh_info: null argument");
    memset(o, 0, sizeof *o);
    collect_update_timers(s);
    for (auto& g : s->geo) if (g.needed) o->blas_nodes += g.n_nodes;
    o->blas_nodes += s->merged.n_nodes;
    o->blas_tris = s->total_tris; o->tlas_nodes = s->tlas_nodes;
    o->bytes = o->blas_nodes * (uint64_t)RT_NODE_BYTES + o->blas_tris * 48ull + o->tlas_nodes * (uint64_t)RT_NODE_BYTES + (uint64_t)s->instances.size() * (64 + 48);
    o->max_depth_blas = s->blas_depth; o->max_depth_tlas = s->tlas_depth;
    o->build_ms = s->build_ms; o->refit_ms = s->refit_ms; o->skin_ms = s->skin_ms; o->tlas_ms = s->tlas_ms;
    return 0;
}

}  // extern "C"
