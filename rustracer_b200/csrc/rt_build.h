// rt_build.h — GPU BVH builder: Morton codes -> radix sort -> Karras LBVH -> bottom-up AABB fit ->
// surface-area-guided collapse to 8-wide nodes with bfloat16 child planes (128 B/node = one cache line), plus bottom-up refit.
// Replaces vkCmdBuildAccelerationStructuresKHR (crates/libs/vulkan/src/ray_tracing/acceleration_structure.rs:95-173)
// as driven by create_as / create_top_as (crates/libs/asset_loader/src/acceleration_structures.rs:79-249).
// The same builder serves BLASes (primitives = triangles) and the TLAS (primitives = instance boxes).
#pragma once
#include "rt_scene_dev.h"

struct DAabb { float lo[3], hi[3]; };

RT_D float aabb_half_area(const DAabb& b) {
    const float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
    return ex * ey + ey * ez + ez * ex;
}
RT_D DAabb aabb_union(const DAabb& a, const DAabb& b) {
    DAabb r;
    for (int k = 0; k < 3; ++k) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return r;
}
RT_D uint32_t float_to_ordered(float f) { uint32_t u = rt_float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
RT_D float ordered_to_float(uint32_t u) { return rt_uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

RT_D uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

struct BuildScratch {
    size_t capacity = 0;        // primitives
    uint64_t *keys = nullptr, *keys_tmp = nullptr;
    uint32_t *vals = nullptr, *vals_tmp = nullptr;
    int2* bin_children = nullptr;     // n-1 internal nodes: child ids (internal k -> k, leaf k -> n-1+k)
    uint32_t* bin_parent = nullptr;   // 2n-1
    DAabb* bin_box = nullptr;         // 2n-1
    uint2* bin_range = nullptr;       // n-1: [first,last] sorted-primitive range of each internal node (LBVH only)
    uint32_t* bin_count = nullptr;    // n-1: primitives below each internal node
    uint32_t *cl_a = nullptr, *cl_b = nullptr;   // PLOC: cluster (binary node id) lists, ping-pong
    uint32_t *nn = nullptr, *valid = nullptr, *pos = nullptr;   // PLOC: nearest neighbour, survives-this-round flag, scan
    uint32_t* flags = nullptr;        // n-1
    uint32_t* bounds = nullptr;       // 6 ordered-uint centroid bounds
    uint2 *q_a = nullptr, *q_b = nullptr;   // collapse work queues: (binary node, wide node)
    uint32_t* counters = nullptr;     // [0] wide nodes, [1] primitives emitted, [2] next queue size
    void* lib_tmp = nullptr; size_t lib_tmp_bytes = 0;   // radix sort / scan temporary storage: per scene (device + stream of its context)
};

inline int scratch_reserve(BuildScratch& s, size_t n) {
    if (n <= s.capacity) return 0;
    void** ptrs[] = {(void**)&s.keys, (void**)&s.keys_tmp, (void**)&s.vals, (void**)&s.vals_tmp, (void**)&s.bin_children, (void**)&s.bin_parent,
                     (void**)&s.bin_box, (void**)&s.bin_range, (void**)&s.flags, (void**)&s.bounds, (void**)&s.q_a, (void**)&s.q_b, (void**)&s.counters,
                     (void**)&s.bin_count, (void**)&s.cl_a, (void**)&s.cl_b, (void**)&s.nn, (void**)&s.valid, (void**)&s.pos};
    for (void** p : ptrs) { if (*p) rt_free(*p); *p = nullptr; }
    size_t cap = n + n / 4 + 16;
    int e = 0;
    e |= rt_malloc((void**)&s.keys, cap * 8); e |= rt_malloc((void**)&s.keys_tmp, cap * 8);
    e |= rt_malloc((void**)&s.vals, cap * 4); e |= rt_malloc((void**)&s.vals_tmp, cap * 4);
    e |= rt_malloc((void**)&s.bin_children, cap * sizeof(int2)); e |= rt_malloc((void**)&s.bin_parent, 2 * cap * 4);
    e |= rt_malloc((void**)&s.bin_box, 2 * cap * sizeof(DAabb)); e |= rt_malloc((void**)&s.bin_range, cap * sizeof(uint2));
    e |= rt_malloc((void**)&s.flags, cap * 4); e |= rt_malloc((void**)&s.bounds, 64);
    e |= rt_malloc((void**)&s.q_a, cap * sizeof(uint2)); e |= rt_malloc((void**)&s.q_b, cap * sizeof(uint2));
    e |= rt_malloc((void**)&s.counters, 64);
    e |= rt_malloc((void**)&s.bin_count, cap * 4); e |= rt_malloc((void**)&s.cl_a, cap * 4); e |= rt_malloc((void**)&s.cl_b, cap * 4);
    e |= rt_malloc((void**)&s.nn, cap * 4); e |= rt_malloc((void**)&s.valid, cap * 4); e |= rt_malloc((void**)&s.pos, cap * 4);
    s.capacity = e ? 0 : cap;
    return e;
}
inline void scratch_free(BuildScratch& s) {
    void** ptrs[] = {(void**)&s.keys, (void**)&s.keys_tmp, (void**)&s.vals, (void**)&s.vals_tmp, (void**)&s.bin_children, (void**)&s.bin_parent,
                     (void**)&s.bin_box, (void**)&s.bin_range, (void**)&s.flags, (void**)&s.bounds, (void**)&s.q_a, (void**)&s.q_b, (void**)&s.counters,
                     (void**)&s.bin_count, (void**)&s.cl_a, (void**)&s.cl_b, (void**)&s.nn, (void**)&s.valid, (void**)&s.pos};
    for (void** p : ptrs) { if (*p) rt_free(*p); *p = nullptr; }
    if (s.lib_tmp) rt_free(s.lib_tmp);
    s.lib_tmp = nullptr; s.lib_tmp_bytes = 0;
    s.capacity = 0;
}

// ---- Karras 2012 -------------------------------------------------------------------------------------
RT_D int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + rt_clz32((uint32_t)i ^ (uint32_t)j);
    return rt_clz64(a ^ b);
}
RT_D void lbvh_build_node(const uint64_t* keys, int n, int i, int2* children, uint32_t* parent, uint2* range, uint32_t* count) {
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    parent[left] = (uint32_t)i; parent[right] = (uint32_t)i;
    range[i] = make_uint2((uint32_t)lo, (uint32_t)hi);
    count[i] = (uint32_t)(hi - lo + 1);
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

// ---- wide node emission --------------------------------------------------------------------------------
struct WideOut {
    float4* nodes;          // RT_NODE_F4 per node (indices local to this BVH)
    uint32_t* prim_order;   // leaf-order -> source primitive index
    DAabb* node_box;        // unquantised box per wide node (refit / TLAS input)
    uint32_t* node_parent;  // parent wide node (0xFFFFFFFF for the root)
    uint32_t max_nodes;
    uint32_t leaf_max = RT_LEAF_MAX;   // primitives a leaf slot may hold (<= RT_LEAF_MAX).  The TLAS uses 1: entering an
                                       // instance costs ~200 instructions, so every instance gets its own box test first
};

// ---- bfloat16 child planes (rt_scene_dev.h: node layout) ----------------------------------------------------
// plane value = child coordinate - node origin, evaluated exactly in double, then rounded conservatively: lo planes towards
// -inf, hi planes towards +inf.  Inputs are >= 0 (the origin is the minimum of the child lows).
RT_D float float_round_down(double d) { float f = (float)d; if ((double)f > d) f = nextafterf(f, -3.0e38f); return f; }
RT_D float float_round_up(double d) { float f = (float)d; if ((double)f < d) f = nextafterf(f, 3.0e38f); return f; }
RT_D uint32_t bf16_lo_plane(float child_lo, float origin) { return rt_float_as_uint(float_round_down((double)child_lo - (double)origin)) >> 16; }   // truncation == floor for values >= 0
RT_D uint32_t bf16_hi_plane(float child_hi, float origin) { return (rt_float_as_uint(float_round_up((double)child_hi - (double)origin)) + 0xFFFFu) >> 16; }
#define RT_PLANE_EMPTY 0x00007F80u      // hi = 0, lo = +inf: the slab is never entered
// exponent byte E with 2^(E-127) >= every plane value the traversal can read from a node whose largest hi plane is hi_max
// (a hi word read as a float is < hi + one bf16 ulp)
RT_D uint32_t plane_bound_exponent(uint32_t hi_max_bf16) { uint32_t e = (((hi_max_bf16 + 1u) << 16) >> 23) + 1u; return e > 254u ? 254u : e; }

// Writes one wide node.  child_box[i] valid for meta[i] != 0.  Conservative: the planes the traversal reconstructs
// always contain the child box.
RT_D void write_wide_node(float4* node, const DAabb& box, const DAabb* child_box, const uint32_t* meta, uint32_t imask,
                          uint32_t child_base, uint32_t prim_base) {
    uint32_t w[3][8]; uint32_t hi_max = 0u;
    for (int i = 0; i < 8; ++i) {
        for (int a = 0; a < 3; ++a) {
            if (!meta[i]) { w[a][i] = RT_PLANE_EMPTY; continue; }
            const uint32_t lo = bf16_lo_plane(child_box[i].lo[a], box.lo[a]), hi = bf16_hi_plane(child_box[i].hi[a], box.lo[a]);
            w[a][i] = (hi << 16) | lo;
            if (hi > hi_max) hi_max = hi;
        }
    }
    const uint32_t e_imask = plane_bound_exponent(hi_max) | (imask << 24);
    node[0] = make_float4(box.lo[0], box.lo[1], box.lo[2], rt_uint_as_float(e_imask));
    node[1] = make_float4(rt_uint_as_float(child_base), rt_uint_as_float(prim_base), rt_uint_as_float(meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24)),
                          rt_uint_as_float(meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24)));
    for (int a = 0; a < 3; ++a) {
        node[2 + 2 * a] = make_float4(rt_uint_as_float(w[a][0]), rt_uint_as_float(w[a][1]), rt_uint_as_float(w[a][2]), rt_uint_as_float(w[a][3]));
        node[3 + 2 * a] = make_float4(rt_uint_as_float(w[a][4]), rt_uint_as_float(w[a][5]), rt_uint_as_float(w[a][6]), rt_uint_as_float(w[a][7]));
    }
}

// One collapse work item: binary subtree `bin` becomes wide node `wide`.
// Binary node ids: internal k -> k (root 0), leaf at sorted position k -> n-1+k.
RT_D void collapse_node(int n, const int2* bin_children, const DAabb* bin_box, const uint32_t* bin_count, const uint32_t* vals,
                        uint32_t bin, uint32_t wide, uint32_t parent_wide, WideOut out, uint32_t* counters, uint2* q_out, uint32_t* q_out_count) {
    auto count_of = [&](uint32_t c) -> uint32_t { return c >= (uint32_t)(n - 1) ? 1u : bin_count[c]; };
    uint32_t ch[8]; int nch = 0;
    if (n == 1 || count_of(bin) <= out.leaf_max) { ch[nch++] = bin; }
    else { ch[nch++] = (uint32_t)bin_children[bin].x; ch[nch++] = (uint32_t)bin_children[bin].y; }
    // greedy: open the child with the largest surface area until 8 children or nothing left to open
    while (nch < 8) {
        int best = -1; float best_area = -1.0f;
        for (int i = 0; i < nch; ++i) {
            if (count_of(ch[i]) <= out.leaf_max) continue;
            const float a = aabb_half_area(bin_box[ch[i]]);
            if (a > best_area) { best_area = a; best = i; }
        }
        if (best < 0) break;
        const int2 c = bin_children[ch[best]];
        ch[best] = (uint32_t)c.x; ch[nch++] = (uint32_t)c.y;
    }
    // spare slots are free to test (the node test always does 8 boxes): split multi-primitive leaf slots further,
    // largest first, so that each primitive sits in the tightest box the node can give it
    while (nch < 8) {
        int best = -1; float best_area = -1.0f;
        for (int i = 0; i < nch; ++i) {
            if (ch[i] >= (uint32_t)(n - 1) || count_of(ch[i]) > out.leaf_max) continue;   // single primitive / inner child
            const float a = aabb_half_area(bin_box[ch[i]]);
            if (a > best_area) { best_area = a; best = i; }
        }
        if (best < 0) break;
        const int2 c = bin_children[ch[best]];
        ch[best] = (uint32_t)c.x; ch[nch++] = (uint32_t)c.y;
    }
    const DAabb box = (n == 1) ? bin_box[0] : bin_box[bin];
    // slot assignment: slot s is visited first by rays travelling in the negative direction of the axes whose bit
    // is set in s, so it should hold the child lying furthest towards +axis on those axes (greedy max-cost matching)
    const float cx = 0.5f * (box.lo[0] + box.hi[0]), cy = 0.5f * (box.lo[1] + box.hi[1]), cz = 0.5f * (box.lo[2] + box.hi[2]);
    float cost[8][8];
    for (int i = 0; i < nch; ++i) {
        const DAabb& b = bin_box[ch[i]];
        const float dx = 0.5f * (b.lo[0] + b.hi[0]) - cx, dy = 0.5f * (b.lo[1] + b.hi[1]) - cy, dz = 0.5f * (b.lo[2] + b.hi[2]) - cz;
        for (int s = 0; s < 8; ++s) cost[i][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
    }
    int slot_child[8]; for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    bool used[8] = {false, false, false, false, false, false, false, false};
    for (int k = 0; k < nch; ++k) {
        int bi = -1, bs = -1; float bc = -3.0e38f;
        for (int i = 0; i < nch; ++i) {
            if (used[i]) continue;
            for (int s = 0; s < 8; ++s) if (slot_child[s] < 0 && cost[i][s] > bc) { bc = cost[i][s]; bi = i; bs = s; }
        }
        used[bi] = true; slot_child[bs] = bi;
    }
    // emit
    uint32_t n_inner = 0, n_prims = 0;
    for (int s = 0; s < 8; ++s) {
        if (slot_child[s] < 0) continue;
        const uint32_t c = ch[slot_child[s]], cnt = count_of(c);
        if (cnt <= out.leaf_max) n_prims += cnt; else n_inner++;
    }
    const uint32_t child_base = n_inner ? rt_atomic_add(&counters[0], n_inner) : 0u;
    const uint32_t prim_base = n_prims ? rt_atomic_add(&counters[1], n_prims) : 0u;
    uint32_t meta[8]; DAabb cbox[8]; uint32_t imask = 0, inner_rank = 0, prim_off = 0;
    for (int s = 0; s < 8; ++s) {
        meta[s] = 0;
        if (slot_child[s] < 0) continue;
        const uint32_t c = ch[slot_child[s]], cnt = count_of(c);
        cbox[s] = (n == 1) ? bin_box[0] : bin_box[c];
        if (cnt <= out.leaf_max) {
            // gather the (<= out.leaf_max <= RT_LEAF_MAX) primitives below c, left to right
            uint32_t stack[RT_LEAF_MAX]; int sp = 0; uint32_t cur = c, k = 0;
            for (;;) {
                if (cur >= (uint32_t)(n - 1) || n == 1) {
                    out.prim_order[prim_base + prim_off + k++] = vals[n == 1 ? 0u : cur - (uint32_t)(n - 1)];
                    if (!sp) break;
                    cur = stack[--sp];
                } else { const int2 cc = bin_children[cur]; stack[sp++] = (uint32_t)cc.y; cur = (uint32_t)cc.x; }
            }
            meta[s] = (((1u << cnt) - 1u) << 5) | prim_off;   // unary count | first offset
            prim_off += cnt;
        } else {
            const uint32_t w = child_base + inner_rank;
            if (w < out.max_nodes) {
                const uint32_t qi = rt_atomic_add(q_out_count, 1u);
                q_out[qi] = make_uint2(c, w);
                out.node_parent[w] = wide;
            }
            meta[s] = (1u << 5) | (24u + (uint32_t)s);
            imask |= 1u << s; inner_rank++;
        }
    }
    out.node_box[wide] = box;
    if (parent_wide == 0xFFFFFFFFu) out.node_parent[wide] = 0xFFFFFFFFu;
    write_wide_node(out.nodes + (size_t)wide * RT_NODE_F4, box, cbox, meta, imask, child_base, prim_base);
}

struct WideBvhInfo { uint32_t n_nodes = 0, n_prims = 0, depth = 0; };

#ifdef RT_EMU
#include <functional>
// EXPERIMENT (emulation only, RT_EMU_SAH=1): binned-SAH top-down build into the LBVH's binary-tree arrays.
inline void emu_sah_binary_build(const DAabb* boxes, int n, uint32_t* vals, int2* children, uint32_t* parent, uint32_t* bcount, DAabb* bin_box) {
    std::vector<uint32_t> order(n);
    for (int i = 0; i < n; ++i) order[i] = (uint32_t)i;
    int next_internal = 0;
    auto grow = [&](DAabb& a, const DAabb& b) { for (int k = 0; k < 3; ++k) { a.lo[k] = fminf(a.lo[k], b.lo[k]); a.hi[k] = fmaxf(a.hi[k], b.hi[k]); } };
    auto empty = []() { DAabb b; for (int k = 0; k < 3; ++k) { b.lo[k] = 3e38f; b.hi[k] = -3e38f; } return b; };
    std::function<uint32_t(int, int, uint32_t)> rec = [&](int first, int count, uint32_t par) -> uint32_t {
        if (count == 1) { const uint32_t id = (uint32_t)(n - 1 + first); bin_box[id] = boxes[order[first]]; parent[id] = par; return id; }
        const uint32_t id = (uint32_t)next_internal++;
        parent[id] = par;
        DAabb cb = empty(), nb = empty();
        for (int i = 0; i < count; ++i) { const DAabb& b = boxes[order[first + i]]; grow(nb, b); DAabb c; for (int k = 0; k < 3; ++k) c.lo[k] = c.hi[k] = 0.5f * (b.lo[k] + b.hi[k]); grow(cb, c); }
        const int NB = 16; int bestAxis = -1, bestSplit = 0; float bestCost = 3e38f;
        for (int ax = 0; ax < 3; ++ax) {
            const float lo = cb.lo[ax], hi = cb.hi[ax]; if (!(hi > lo)) continue;
            DAabb bb[NB]; int bc[NB]; for (int b = 0; b < NB; ++b) { bb[b] = empty(); bc[b] = 0; }
            const float scale = NB / (hi - lo);
            for (int i = 0; i < count; ++i) { const DAabb& b = boxes[order[first + i]]; int k = std::min(NB - 1, (int)((0.5f * (b.lo[ax] + b.hi[ax]) - lo) * scale)); grow(bb[k], b); bc[k]++; }
            float la[NB], ra[NB]; int lc[NB], rc[NB]; DAabb l = empty(), r = empty(); int ls = 0, rs = 0;
            for (int i = 0; i < NB - 1; ++i) { ls += bc[i]; if (bc[i]) grow(l, bb[i]); lc[i] = ls; la[i] = ls ? aabb_half_area(l) : 0.f;
                                               rs += bc[NB - 1 - i]; if (bc[NB - 1 - i]) grow(r, bb[NB - 1 - i]); rc[NB - 2 - i] = rs; ra[NB - 2 - i] = rs ? aabb_half_area(r) : 0.f; }
            for (int i = 0; i < NB - 1; ++i) { if (!lc[i] || !rc[i]) continue; const float c = lc[i] * la[i] + rc[i] * ra[i]; if (c < bestCost) { bestCost = c; bestAxis = ax; bestSplit = i; } }
        }
        int mid;
        if (bestAxis < 0) mid = first + count / 2;
        else {
            const float lo = cb.lo[bestAxis], scale = NB / (cb.hi[bestAxis] - lo);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t p) {
                const DAabb& b = boxes[p]; return std::min(NB - 1, (int)((0.5f * (b.lo[bestAxis] + b.hi[bestAxis]) - lo) * scale)) <= bestSplit; });
            mid = (int)(it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        const uint32_t l = rec(first, mid - first, id), r = rec(mid, first + count - mid, id);
        children[id] = make_int2((int)l, (int)r); bcount[id] = (uint32_t)count; bin_box[id] = nb;
        return id;
    };
    rec(0, n, 0xFFFFFFFFu);
    for (int i = 0; i < n; ++i) vals[i] = order[i];
}
#endif

// ---- PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) ---------------------------------
// Bottom-up agglomerative build over the Morton-ordered cluster list: every round each cluster finds the
// neighbour within +-RT_PLOC_RADIUS list positions whose union box has the smallest area; mutual nearest
// neighbours merge; the list is compacted order-preservingly.  ~20 % fewer node visits per ray than the LBVH
// on the Lucy stand-in.  Internal ids are handed out downwards from n-2 so that the last merge is the root, id 0.
#ifndef RT_MORTON_MAX_ASPECT
#define RT_MORTON_MAX_ASPECT 2.0f
#endif
#ifndef RT_PLOC_RADIUS
#define RT_PLOC_RADIUS 16
#endif
#ifndef RT_COLLAPSE_BATCH
#define RT_COLLAPSE_BATCH 8     // collapse levels launched between two host reads of the queue size
#endif
#ifndef RT_PLOC_BATCH
#define RT_PLOC_BATCH 16     // rounds between two host reads of the cluster count (465 k triangles: 113 rounds, ~10 ms)
#endif
inline int builder_kind() {
    static int kind = -1;
    if (kind < 0) { const char* e = getenv("RT_B200_BUILDER"); kind = (e && e[0] == 'l') ? 0 : 1; }   // "lbvh" | "ploc" (default)
    return kind;
}

#ifndef RT_EMU
// Last PLOC rounds in ONE launch: once at most RT_PLOC_FINISH_MAX clusters are left (the long tail of the algorithm: most of
// its ~110 rounds for 465 k triangles), a single 1024-thread CTA runs the remaining rounds back to back — the same three
// phases (nearest neighbour, survivor flags + exclusive scan, merge + order-preserving compaction) with __syncthreads()
// between them instead of 5 launches per round and a host read every few rounds.  Same arithmetic and the same id
// assignment as the per-round kernels below: the tree is identical.
#define RT_PLOC_FINISH_MAX 4096
__global__ void __launch_bounds__(1024) ploc_finish_kernel(int n, int2* bin_children, DAabb* bin_box, uint32_t* bin_count, uint32_t* cl_in, uint32_t* cl_out,
                                                          uint32_t* nn, uint32_t* valid, uint32_t* pos, const uint32_t* cur) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t c = cur[0], merged_before = cur[1];
    while (c > 1u) {
        const int ci = (int)c;
        for (uint32_t i = tid; i < c; i += 1024u) {
            const DAabb a = bin_box[cl_in[i]];
            const int lo = (int)i - RT_PLOC_RADIUS < 0 ? 0 : (int)i - RT_PLOC_RADIUS;
            const int hi = (int)i + RT_PLOC_RADIUS > ci - 1 ? ci - 1 : (int)i + RT_PLOC_RADIUS;
            float best = 3.0e38f; int bj = -1;
            for (int j = lo; j <= hi; ++j) {
                if (j == (int)i) continue;
                const float d = aabb_half_area(aabb_union(a, bin_box[cl_in[j]]));
                if (d < best || bj < 0) { best = d; bj = j; }
            }
            nn[i] = (uint32_t)bj;
        }
        __syncthreads();
        // survivor flags + exclusive scan: thread t owns the entries [t * per, t * per + per)
        const uint32_t per = (c + 1023u) / 1024u, first = tid * per;
        uint32_t local = 0u;
        for (uint32_t k = 0; k < per; ++k) {
            const uint32_t i = first + k;
            if (i < c) { const uint32_t j = nn[i]; const uint32_t v = (nn[j] == i && j < i) ? 0u : 1u; valid[i] = v; local += v; }
        }
        uint32_t incl = local;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
        if (lane == 31u) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0u) {
            uint32_t w = s_warp[lane], wi = w;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane >= d) wi += o; }
            s_warp[lane] = wi - w;
            if (lane == 31u) s_total = wi;
        }
        __syncthreads();
        uint32_t run = s_warp[warp] + incl - local;
        for (uint32_t k = 0; k < per; ++k) { const uint32_t i = first + k; if (i < c) { pos[i] = run; run += valid[i]; } }
        const uint32_t c_new = s_total;
        __syncthreads();
        for (uint32_t i = tid; i < c; i += 1024u) {
            if (!valid[i]) continue;
            const uint32_t j = nn[i];
            uint32_t id = cl_in[i];
            if (nn[j] == i) {
                const uint32_t rank = j - pos[j];   // absorbed entries before j == merges before this one
                const uint32_t a = id, b = cl_in[j];
                id = (uint32_t)(n - 2) - (merged_before + rank);
                bin_children[id] = make_int2((int)a, (int)b);
                bin_box[id] = aabb_union(bin_box[a], bin_box[b]);
                bin_count[id] = (a >= (uint32_t)(n - 1) ? 1u : bin_count[a]) + (b >= (uint32_t)(n - 1) ? 1u : bin_count[b]);
            }
            cl_out[pos[i]] = id;
        }
        merged_before += c - c_new; c = c_new;
        uint32_t* t = cl_in; cl_in = cl_out; cl_out = t;
        __syncthreads();
    }
}
#endif

inline int ploc_build(const DAabb* prim_boxes, int n, BuildScratch& sc, rt_stream_t stream) {
    int2* bin_children = sc.bin_children; DAabb* bin_box = sc.bin_box; uint32_t* bin_count = sc.bin_count;
    const uint32_t* vals = sc.vals; uint32_t *nn = sc.nn, *valid = sc.valid, *pos = sc.pos, *counters = sc.counters;
    uint32_t *cl_in = sc.cl_a, *cl_out = sc.cl_b;
    {
        uint32_t* cl = cl_in;
        rt_launch((size_t)n, stream, RT_LAMBDA(size_t i) {
            const uint32_t leaf = (uint32_t)(n - 1) + (uint32_t)i;
            bin_box[leaf] = prim_boxes[vals[i]];
            cl[i] = leaf;
        });
    }
    // The cluster count c and the number of merges so far live on the device (counters[4..5] and [6..7], alternating by
    // round), so that RT_PLOC_BATCH rounds run back to back with launches sized for the count at the start of the batch;
    // the host reads c once per batch instead of once per round (45 rounds for the 465 k triangles of config 2).
    uint32_t c_ub = (uint32_t)n, rounds = 0;
    { const uint32_t init[4] = {(uint32_t)n, 0u, (uint32_t)n, 0u}; if (rt_h2d(&counters[4], init, sizeof init, stream)) return 1; }
    while (c_ub > 1) {
#ifndef RT_EMU
        if (c_ub <= RT_PLOC_FINISH_MAX) {     // the long tail of the algorithm: one launch, no further host reads
            ploc_finish_kernel<<<1, 1024, 0, stream>>>(n, bin_children, bin_box, bin_count, cl_in, cl_out, nn, valid, pos, counters + 4 + 2 * (rounds & 1u));
            ++g_rt_launch_count;
            break;
        }
#endif
        // launches are sized for the cluster count at the start of a batch: short batches while the list is long (a round
        // removes 30-45 % of it), longer ones later
        const int batch = c_ub > 65536u ? 2 : (c_ub > 16384u ? 3 : RT_PLOC_BATCH);
        for (int r = 0; r < batch; ++r, ++rounds) {
            const uint32_t* cl = cl_in; uint32_t* co = cl_out;
            const uint32_t* cur = counters + 4 + 2 * (rounds & 1u); uint32_t* nxt = counters + 4 + 2 * ((rounds + 1u) & 1u);
            rt_launch(c_ub, stream, RT_LAMBDA(size_t i) {
                const int ci = (int)cur[0];
                if ((int)i >= ci || ci <= 1) return;
                const DAabb a = bin_box[cl[i]];
                const int lo = (int)i - RT_PLOC_RADIUS < 0 ? 0 : (int)i - RT_PLOC_RADIUS;
                const int hi = (int)i + RT_PLOC_RADIUS > ci - 1 ? ci - 1 : (int)i + RT_PLOC_RADIUS;
                float best = 3.0e38f; int bj = -1;
                for (int j = lo; j <= hi; ++j) {
                    if (j == (int)i) continue;
                    const float d = aabb_half_area(aabb_union(a, bin_box[cl[j]]));
                    if (d < best || bj < 0) { best = d; bj = j; }   // ties -> lowest position, which guarantees a mutual pair
                }
                nn[i] = (uint32_t)bj;
            });
            rt_launch(c_ub, stream, RT_LAMBDA(size_t i) {
                const uint32_t ci = cur[0];
                if (i >= ci || ci <= 1u) { valid[i] = (ci <= 1u && i == 0) ? 1u : 0u; return; }
                const uint32_t j = nn[i];
                valid[i] = (nn[j] == (uint32_t)i && j < (uint32_t)i) ? 0u : 1u;   // the higher half of a pair is absorbed
            });
            if (rt_exclusive_scan_u32(valid, pos, c_ub, &sc.lib_tmp, &sc.lib_tmp_bytes, stream)) return 1;
            rt_launch(c_ub, stream, RT_LAMBDA(size_t i) {
                const uint32_t ci = cur[0], merged_before = cur[1];
                if (ci <= 1u) { if (i == 0) { nxt[0] = ci; nxt[1] = merged_before; co[0] = cl[0]; } return; }
                if (i >= ci) return;
                if (i == (size_t)(ci - 1u)) { const uint32_t c_new = pos[i] + valid[i]; nxt[0] = c_new; nxt[1] = merged_before + (ci - c_new); }
                if (!valid[i]) return;
                const uint32_t j = nn[i];
                uint32_t id = cl[i];
                if (nn[j] == (uint32_t)i) {
                    const uint32_t rank = j - pos[j];   // absorbed entries before j == merges before this one
                    const uint32_t a = id, b = cl[j];
                    id = (uint32_t)(n - 2) - (merged_before + rank);
                    bin_children[id] = make_int2((int)a, (int)b);
                    bin_box[id] = aabb_union(bin_box[a], bin_box[b]);
                    bin_count[id] = (a >= (uint32_t)(n - 1) ? 1u : bin_count[a]) + (b >= (uint32_t)(n - 1) ? 1u : bin_count[b]);
                }
                co[pos[i]] = id;
            });
            uint32_t* t = cl_in; cl_in = cl_out; cl_out = t;
        }
        uint32_t c_new = 0;
        if (rt_d2h(&c_new, counters + 4 + 2 * (rounds & 1u), 4, stream)) return 1;
        if (rt_stream_sync(stream)) return 1;
        if (c_new == 0 || c_new > c_ub || (c_new == c_ub && c_ub > 1)) return 4;   // cannot happen: every round has at least one mutual pair
        c_ub = c_new;
        if (rounds > 100000u) return 4;
    }
    return 0;
}

// Builds an 8-wide BVH over n primitive boxes (device pointer).  Synchronises the stream once per tree level.
inline int build_wide_bvh(const DAabb* prim_boxes, uint32_t n, BuildScratch& sc, WideOut out, rt_stream_t stream, WideBvhInfo* info) {
    info->n_nodes = 0; info->n_prims = n; info->depth = 0;
    if (n == 0) {
        // empty BVH: one root node with no children
        DAabb* nb = out.node_box; float4* nodes = out.nodes; uint32_t* np = out.node_parent;
        rt_launch(1, stream, RT_LAMBDA(size_t) {
            DAabb b; for (int k = 0; k < 3; ++k) { b.lo[k] = 0.0f; b.hi[k] = 0.0f; }
            uint32_t meta[8] = {0, 0, 0, 0, 0, 0, 0, 0}; DAabb cb[8];
            nb[0] = b; np[0] = 0xFFFFFFFFu;
            write_wide_node(nodes, b, cb, meta, 0, 0, 0);
        });
        info->n_nodes = 1; info->depth = 1;
        return 0;
    }
    if (scratch_reserve(sc, n)) return 1;
    uint32_t* bounds = sc.bounds; uint64_t* keys = sc.keys; uint32_t* vals = sc.vals;
    const uint32_t init_bounds[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    rt_h2d(bounds, init_bounds, sizeof init_bounds, stream);
    rt_launch(n, stream, RT_LAMBDA(size_t i) {
        const DAabb b = prim_boxes[i];
        for (int k = 0; k < 3; ++k) {
            const float c = 0.5f * (b.lo[k] + b.hi[k]);
            rt_atomic_min(&bounds[k], float_to_ordered(c)); rt_atomic_max(&bounds[3 + k], float_to_ordered(c));
        }
    });
    float max_aspect = RT_MORTON_MAX_ASPECT;
    if (const char* e = getenv("RT_B200_MORTON_ASPECT")) { const float v = (float)atof(e); if (v >= 1.0f) max_aspect = v; }   // (env: A/B knob; 1e30 = per axis)
    rt_launch(n, stream, RT_LAMBDA(size_t i) {
        const DAabb b = prim_boxes[i];
        uint32_t q[3];
        // Morton cells are kept (nearly) cubic: an axis is never normalised by less than 1 / max_aspect of the largest
        // centroid extent.  Plain per-axis normalisation stretches the thin axis of a flat scene (a field of instances, a
        // layer of foliage cards) until its bits are noise that breaks the locality of the other two (config 3: twice the
        // node visits in the card BLAS, +56 % in the TLAS)
        float ext = 0.0f;
        for (int k = 0; k < 3; ++k) ext = fmaxf(ext, ordered_to_float(bounds[3 + k]) - ordered_to_float(bounds[k]));
        for (int k = 0; k < 3; ++k) {
            const float lo = ordered_to_float(bounds[k]), hi = ordered_to_float(bounds[3 + k]);
            const float c = 0.5f * (b.lo[k] + b.hi[k]);
            const float e = fmaxf(hi - lo, ext / max_aspect);
            float f = (e > 0.0f) ? (c - lo) / e : 0.0f;
            f = fminf(fmaxf(f, 0.0f), 1.0f);
            q[k] = (uint32_t)fminf(f * 2097152.0f, 2097151.0f);
        }
        keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
        vals[i] = (uint32_t)i;
    });
    if (rt_sort_pairs_u64(sc.keys, sc.vals, sc.keys_tmp, sc.vals_tmp, n, &sc.lib_tmp, &sc.lib_tmp_bytes, stream)) return 1;
    int2* bin_children = sc.bin_children; uint32_t* bin_parent = sc.bin_parent; uint2* bin_range = sc.bin_range;
    DAabb* bin_box = sc.bin_box; uint32_t* flags = sc.flags; uint32_t* bin_count = sc.bin_count;
    const int ni = (int)n;
#ifdef RT_EMU
    if (n > 1 && getenv("RT_EMU_SAH")) {
        // EXPERIMENT (emulation only): top-down binned-SAH binary tree instead of the LBVH, to measure how much
        // traversal work a higher-quality builder would save.  Not part of the product.
        emu_sah_binary_build(prim_boxes, ni, vals, bin_children, bin_parent, bin_count, bin_box);
    } else
#endif
    if (n > 1 && builder_kind() == 1) {
        if (int e = ploc_build(prim_boxes, ni, sc, stream)) return e;
    } else if (n > 1) {
        rt_memset(flags, 0, (size_t)(n - 1) * 4, stream);
        rt_launch(n - 1, stream, RT_LAMBDA(size_t i) { lbvh_build_node(keys, ni, (int)i, bin_children, bin_parent, bin_range, bin_count); });
        rt_launch(n, stream, RT_LAMBDA(size_t i) {
            const uint32_t leaf = (uint32_t)(ni - 1) + (uint32_t)i;
            bin_box[leaf] = prim_boxes[vals[i]];
            uint32_t cur = bin_parent[leaf];
            for (;;) {
                rt_threadfence();
                const uint32_t old = rt_atomic_add(&flags[cur], 1u);
                if (old == 0u) return;   // first arrival: the sibling subtree is not finished yet
                rt_threadfence();
                const int2 c = bin_children[cur];
                // volatile-style reads: the sibling's box was published before its atomic
                const volatile DAabb* lb = bin_box + c.x; const volatile DAabb* rb = bin_box + c.y;
                DAabb u;
                for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(lb->lo[k], rb->lo[k]); u.hi[k] = fmaxf(lb->hi[k], rb->hi[k]); }
                bin_box[cur] = u;
                if (cur == 0u) return;
                cur = bin_parent[cur];
            }
        });
    } else {
        rt_launch(1, stream, RT_LAMBDA(size_t) { bin_box[0] = prim_boxes[vals[0]]; });
    }
    // collapse, level by level.  The size of a level's work queue lives on the device ([8] / [9], ping-pong): the level
    // kernels are launched blind, sized by an upper bound (8x the previous one, at most max_nodes), and a thread beyond the real
    // queue size leaves at once — the host reads the counters back once per RT_COLLAPSE_BATCH levels instead of synchronising
    // after every one of the ~15 levels (an empty level costs a ~2 us launch).
    // counters: [0] wide nodes, [1] primitives emitted, [3] depth, [8] / [9] queue sizes
    uint32_t* counters = sc.counters;
    {
        uint32_t init[16] = {0}; init[0] = 1u; init[8] = 1u;
        rt_h2d(counters, init, sizeof init, stream);
    }
    const uint2 root_item = make_uint2(0u, 0u);
    rt_h2d(sc.q_a, &root_item, sizeof root_item, stream);
    uint2 *q_in = sc.q_a, *q_out = sc.q_b;
    uint32_t depth = 0, level = 0; size_t ub = 1;
    for (;;) {
        for (int k = 0; k < RT_COLLAPSE_BATCH; ++k, ++level) {
            const uint32_t* cin = counters + 8 + (level & 1u); uint32_t* cout = counters + 8 + ((level + 1u) & 1u);
            rt_memset(cout, 0, 4, stream);
            const uint2* qi = q_in; uint2* qo = q_out; const bool is_root = level == 0; const uint32_t lv = level;
            rt_launch(ub, stream, RT_LAMBDA(size_t i) {
                if (i >= (size_t)cin[0]) return;
                if (i == 0) counters[3] = lv + 1u;
                const uint2 it = qi[i];
                collapse_node(ni, bin_children, bin_box, bin_count, vals, it.x, it.y, is_root ? 0xFFFFFFFFu : 0u, out, counters, qo, cout);
            });
            uint2* t = q_in; q_in = q_out; q_out = t;
            ub = ub * 8 < (size_t)out.max_nodes ? ub * 8 : (size_t)out.max_nodes;
        }
        uint32_t host_counters[16];
        if (rt_d2h(host_counters, counters, sizeof host_counters, stream)) return 1;
        if (rt_stream_sync(stream)) return 1;
        if (host_counters[0] > out.max_nodes) return 2;
        info->n_nodes = host_counters[0]; info->n_prims = host_counters[1]; depth = host_counters[3];
        if (host_counters[8 + (level & 1u)] == 0u) break;      // nothing queued for the next level
        if (level > 512) return 3;
    }
    info->depth = depth;
    return 0;
}

// ---- refit -------------------------------------------------------------------------------------------
// Recomputes one wide node from its children: leaf children from `leaf_boxes` (one box per emitted primitive, in
// leaf order), inner children from node_box.  Keeps topology, slots and metas; re-quantises.
RT_D void refit_wide_node(float4* nodes, DAabb* node_box, const DAabb* leaf_boxes, uint32_t w) {
    float4* node = nodes + (size_t)w * RT_NODE_F4;
    const float4 n0 = node[0], n1 = node[1];
    const uint32_t imask = rt_float_as_uint(n0.w) >> 24;
    const uint32_t child_base = rt_float_as_uint(n1.x), prim_base = rt_float_as_uint(n1.y);
    uint32_t meta[8];
    for (int i = 0; i < 4; ++i) { meta[i] = (rt_float_as_uint(n1.z) >> (8 * i)) & 0xFFu; meta[4 + i] = (rt_float_as_uint(n1.w) >> (8 * i)) & 0xFFu; }
    DAabb cbox[8]; DAabb box; bool any = false;
    for (int k = 0; k < 3; ++k) { box.lo[k] = 0.0f; box.hi[k] = 0.0f; }
    for (int s = 0; s < 8; ++s) {
        if (!meta[s]) continue;
        DAabb b;
        if ((meta[s] & 0x18u) == 0x18u) {
            const uint32_t rel = (uint32_t)rt_popc(imask & ~(0xFFFFFFFFu << s));
            b = node_box[child_base + rel];
        } else {
            const uint32_t off = meta[s] & 0x1Fu, cnt = (uint32_t)rt_popc(meta[s] >> 5);
            b = leaf_boxes[prim_base + off];
            for (uint32_t k = 1; k < cnt; ++k) b = aabb_union(b, leaf_boxes[prim_base + off + k]);
        }
        cbox[s] = b;
        box = any ? aabb_union(box, b) : b; any = true;
    }
    node_box[w] = box;
    write_wide_node(node, box, cbox, meta, imask, child_base, prim_base);
}

#ifndef RT_EMU
// Warp-cooperative refit: 8 lanes per wide node (one per child slot), 4 nodes per warp.  Same arithmetic as
// refit_wide_node / write_wide_node (identical node bytes), but the per-child work — fetching the child's box, the
// six conservative bfloat16 roundings — runs in parallel; every lane stores its own three plane words.  The one-thread-per-node version spent 384 us on the 150 k nodes of config 4's character
// (a single lane per warp busy with ~2 k double-precision instructions per node).
// Bottom-up order: groups start at nodes without inner children; the group that completes a parent's last pending
// child (atomic counter) continues with the parent.
__global__ void __launch_bounds__(256) refit_nodes_kernel(float4* nodes, DAabb* node_box, const DAabb* leaf_boxes, const uint32_t* parent,
                                                         uint32_t* pending, uint32_t n_nodes) {
    const uint32_t lane = threadIdx.x & 31u, slot = lane & 7u, gbase = lane & 24u, gmask = 0xFFu << gbase;
    uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (w >= n_nodes) return;
    if ((__float_as_uint(__ldcg(&nodes[(size_t)w * RT_NODE_F4]).w) >> 24) != 0u) return;   // has inner children: reached from below
    for (;;) {
        float4* node = nodes + (size_t)w * RT_NODE_F4;
        const float4 n0 = __ldcg(node), n1 = __ldcg(node + 1);
        const uint32_t imask = __float_as_uint(n0.w) >> 24, child_base = __float_as_uint(n1.x), prim_base = __float_as_uint(n1.y);
        const uint32_t meta = (__float_as_uint(slot < 4u ? n1.z : n1.w) >> (8u * (slot & 3u))) & 0xFFu;
        const bool valid = meta != 0u;
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        if (valid) {
            if ((meta & 0x18u) == 0x18u) {
                const uint32_t rel = (uint32_t)__popc(imask & ~(0xFFFFFFFFu << slot));
                const float* b = reinterpret_cast<const float*>(node_box + child_base + rel);
                for (int a = 0; a < 3; ++a) { lo[a] = __ldcg(b + a); hi[a] = __ldcg(b + 3 + a); }   // written by another group: L2
            } else {
                const uint32_t off = meta & 0x1Fu, cnt = (uint32_t)__popc(meta >> 5);
                for (uint32_t k = 0; k < cnt; ++k) {
                    const DAabb b = leaf_boxes[prim_base + off + k];
                    for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b.lo[a]); hi[a] = fmaxf(hi[a], b.hi[a]); }
                }
            }
        }
        // node box = union of the valid children (min / max are exact: the order does not matter)
        float blo[3], bhi[3];
        for (int a = 0; a < 3; ++a) {
            float l = lo[a], h = hi[a];
            for (int d = 1; d < 8; d <<= 1) { l = fminf(l, __shfl_xor_sync(gmask, l, d)); h = fmaxf(h, __shfl_xor_sync(gmask, h, d)); }
            blo[a] = l; bhi[a] = h;
        }
        const bool any = __ballot_sync(gmask, valid) != 0u;
        if (!any) for (int a = 0; a < 3; ++a) { blo[a] = 0.0f; bhi[a] = 0.0f; }
        // this slot's three plane words + the node's plane bound (same arithmetic as write_wide_node: identical bytes)
        uint32_t hi_max = 0u;
        uint32_t* words = reinterpret_cast<uint32_t*>(node);
        for (int a = 0; a < 3; ++a) {
            uint32_t wv = RT_PLANE_EMPTY;
            if (valid) {
                const uint32_t l = bf16_lo_plane(lo[a], blo[a]), h = bf16_hi_plane(hi[a], blo[a]);
                wv = (h << 16) | l;
                if (h > hi_max) hi_max = h;
            }
            words[8 + 8 * a + slot] = wv;
        }
        for (int d = 1; d < 8; d <<= 1) { const uint32_t o = __shfl_xor_sync(gmask, hi_max, d); hi_max = o > hi_max ? o : hi_max; }
        if (slot == 0u) {
            DAabb box; for (int a = 0; a < 3; ++a) { box.lo[a] = blo[a]; box.hi[a] = bhi[a]; }
            node_box[w] = box;
            node[0] = make_float4(blo[0], blo[1], blo[2], __uint_as_float(plane_bound_exponent(hi_max) | (imask << 24)));
        }
        const uint32_t p = parent[w];
        if (p == 0xFFFFFFFFu) return;
        uint32_t old = 0u;
        if (slot == 0u) { __threadfence(); old = atomicAdd(&pending[p], 0xFFFFFFFFu); }   // publish, then decrement
        old = __shfl_sync(gmask, old, gbase);
        if (old != 1u) return;          // other children still pending
        __threadfence();
        w = p;
    }
}
#endif
