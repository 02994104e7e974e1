// mikktspace_gen.h — tangent generation for primitives that have a normal map but no TANGENT accessor.
//
// The reference calls `mikktspace::generate_tangents` (crate mikktspace 0.3.0, pinned in Cargo.lock; a Rust port of
// Morten S. Mikkelsen's mikktspace.c) from asset_loader/src/geometry.rs:192-212 through the `Geometry` impl at
// :296-350.  The crate is not under /root/reference, so this file restates the published algorithm for the case the
// reference exercises (indexed triangle lists, default 180 degree angular threshold), in float32 with the original's
// operation order:
//   1. weld vertices that are bit-identical in position, normal and texture coordinate
//   2. per triangle: first-order tangent / bitangent from the UV derivatives, orientation flag, magnitudes;
//      triangles with a degenerate UV mapping may join any group ("GROUP_WITH_ANY")
//   3. edge neighbours (first unassigned match per edge)
//   4. per (triangle, corner): grow a group over edge-adjacent triangles that share the welded vertex and the
//      orientation flag ("4-rule groups")
//   5. per group corner: sub-group of members within the angular threshold, angle-weighted sum of the members'
//      tangents projected into the vertex's tangent plane, normalised
//   6. degenerate (zero-area) triangles are left out of 2-5 and copy the tangent of a non-degenerate triangle that uses
//      the same vertex, if any
// The reference-side quirks are kept: results are written per (face, corner) into the *indexed* vertex, so the last
// face that touches a vertex wins (geometry.rs:321-323), and w = -1 when the bitangent preserves orientation, +1
// otherwise (geometry.rs:335-340).  Parity is unpinned (the reference has no test for this path).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace mikk {

struct V3 { float x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float length(V3 a) { return sqrtf(dot(a, a)); }
static inline bool not_zero(float f) { return fabsf(f) > FLT_MIN; }
static inline bool v_not_zero(V3 a) { return not_zero(a.x) || not_zero(a.y) || not_zero(a.z); }
static inline V3 normalize(V3 a) { return (1.0f / length(a)) * a; }
static inline bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

enum { MARK_DEGENERATE = 1, GROUP_WITH_ANY = 4, ORIENT_PRESERVING = 8 };

struct Group { int vertex; bool orient; std::vector<int> faces; };
struct TriInfo {
    int neighbor[3] = {-1, -1, -1};
    int group[3] = {-1, -1, -1};
    V3 os{0, 0, 0}, ot{0, 0, 0};
    float mag_s = 0, mag_t = 0;
    int flag = 0;
};
struct TSpace { V3 os{1, 0, 0}; float mag_s = 1; V3 ot{0, 1, 0}; float mag_t = 1; bool orient = false; };

struct Mesh {
    const float* pos; const float* nrm; const float* uv; size_t stride;   // stride in floats between vertices
    V3 p(uint32_t i) const { const float* q = pos + i * stride; return {q[0], q[1], q[2]}; }
    V3 n(uint32_t i) const { const float* q = nrm + i * stride; return {q[0], q[1], q[2]}; }
    float u(uint32_t i) const { return uv[i * stride]; }
    float v(uint32_t i) const { return uv[i * stride + 1]; }
};

// corner i of triangle f belongs to `g`; grow the group over the two edges around that corner (mikktspace AssignRecur)
static bool assign_recur(const std::vector<int>& tl, std::vector<TriInfo>& ti, std::vector<Group>& groups, int f, int g) {
    TriInfo& t = ti[f];
    const int vert = groups[g].vertex;
    const int* v = &tl[3 * f];
    int i = -1;
    if (v[0] == vert) i = 0; else if (v[1] == vert) i = 1; else if (v[2] == vert) i = 2;
    if (i < 0) return false;
    if (t.group[i] == g) return true;
    if (t.group[i] != -1) return false;
    if (t.flag & GROUP_WITH_ANY) {
        // first to claim this triangle decides its orientation
        if (t.group[0] == -1 && t.group[1] == -1 && t.group[2] == -1) {
            t.flag &= ~ORIENT_PRESERVING;
            t.flag |= groups[g].orient ? ORIENT_PRESERVING : 0;
        }
    }
    if (((t.flag & ORIENT_PRESERVING) != 0) != groups[g].orient) return false;
    groups[g].faces.push_back(f);
    t.group[i] = g;
    const int nl = t.neighbor[i], nr = t.neighbor[i > 0 ? i - 1 : 2];
    if (nl >= 0) assign_recur(tl, ti, groups, nl, g);
    if (nr >= 0) assign_recur(tl, ti, groups, nr, g);
    return true;
}

// angle-weighted tangent frame of corner `vert` over the member faces (mikktspace EvalTspace)
static TSpace eval_tspace(const std::vector<int>& faces, const std::vector<int>& tl, const std::vector<TriInfo>& ti, const Mesh& m, int vert) {
    TSpace r; r.os = {0, 0, 0}; r.ot = {0, 0, 0}; r.mag_s = 0; r.mag_t = 0;
    float angle_sum = 0;
    for (int f : faces) {
        if (ti[f].flag & GROUP_WITH_ANY) continue;
        const int* v = &tl[3 * f];
        int i = -1;
        if (v[0] == vert) i = 0; else if (v[1] == vert) i = 1; else if (v[2] == vert) i = 2;
        const V3 n = m.n((uint32_t)v[i]);
        V3 os = ti[f].os - dot(n, ti[f].os) * n, ot = ti[f].ot - dot(n, ti[f].ot) * n;
        if (v_not_zero(os)) os = normalize(os);
        if (v_not_zero(ot)) ot = normalize(ot);
        const int i2 = v[i < 2 ? i + 1 : 0], i1 = v[i], i0 = v[i > 0 ? i - 1 : 2];
        const V3 p0 = m.p((uint32_t)i0), p1 = m.p((uint32_t)i1), p2 = m.p((uint32_t)i2);
        V3 v1 = p0 - p1, v2 = p2 - p1;
        v1 = v1 - dot(n, v1) * n; if (v_not_zero(v1)) v1 = normalize(v1);
        v2 = v2 - dot(n, v2) * n; if (v_not_zero(v2)) v2 = normalize(v2);
        float c = dot(v1, v2); c = c > 1 ? 1 : (c < -1 ? -1 : c);
        const float angle = acosf(c);
        r.os = r.os + angle * os; r.ot = r.ot + angle * ot;
        r.mag_s += angle * ti[f].mag_s; r.mag_t += angle * ti[f].mag_t;
        angle_sum += angle;
    }
    if (v_not_zero(r.os)) r.os = normalize(r.os);
    if (v_not_zero(r.ot)) r.ot = normalize(r.ot);
    if (angle_sum > 0) { r.mag_s /= angle_sum; r.mag_t /= angle_sum; }
    return r;
}

// tangents: n_vertices x 4 floats, pre-filled by the caller (the reference pre-fills (1,0,0,0)); stride_t floats apart
static inline void generate(const Mesh& m, size_t n_vertices, const uint32_t* indices, size_t n_indices, float* tangents, size_t stride_t) {
    const int nf = (int)(n_indices / 3);
    if (nf <= 0) return;
    // 1. weld
    std::vector<int> tl(3 * (size_t)nf);
    {
        struct Key { uint32_t k[8]; bool operator<(const Key& o) const { return memcmp(k, o.k, sizeof k) < 0; } };
        std::map<Key, int> seen;
        for (size_t c = 0; c < tl.size(); ++c) {
            const uint32_t vi = indices[c];
            const V3 p = m.p(vi), n = m.n(vi); const float t[2] = {m.u(vi), m.v(vi)};
            Key key; memcpy(&key.k[0], &p, 12); memcpy(&key.k[3], &n, 12); memcpy(&key.k[6], t, 8);
            for (uint32_t& w : key.k) if (w == 0x80000000u) w = 0;    // -0 == +0 for the original's float compare
            auto it = seen.find(key);
            if (it == seen.end()) { seen.emplace(key, (int)vi); tl[c] = (int)vi; } else tl[c] = it->second;
        }
    }
    (void)n_vertices;
    std::vector<TriInfo> ti((size_t)nf);
    // degenerate triangles (two equal positions) are handled in the epilogue
    for (int f = 0; f < nf; ++f) {
        const V3 p0 = m.p(indices[3 * f]), p1 = m.p(indices[3 * f + 1]), p2 = m.p(indices[3 * f + 2]);
        if (veq(p0, p1) || veq(p0, p2) || veq(p1, p2)) ti[f].flag |= MARK_DEGENERATE;
    }
    // 2. per-triangle first-order derivatives (InitTriInfo)
    for (int f = 0; f < nf; ++f) {
        TriInfo& t = ti[f];
        if (t.flag & MARK_DEGENERATE) continue;
        t.flag |= GROUP_WITH_ANY;
        const uint32_t a = (uint32_t)tl[3 * f], b = (uint32_t)tl[3 * f + 1], c = (uint32_t)tl[3 * f + 2];
        const V3 v1 = m.p(a), v2 = m.p(b), v3 = m.p(c);
        const float t21x = m.u(b) - m.u(a), t21y = m.v(b) - m.v(a), t31x = m.u(c) - m.u(a), t31y = m.v(c) - m.v(a);
        const V3 d1 = v2 - v1, d2 = v3 - v1;
        const float area2 = t21x * t31y - t21y * t31x;
        V3 os = (t31y * d1) - (t21y * d2), ot = (-t31x * d1) + (t21x * d2);
        if (area2 > 0) t.flag |= ORIENT_PRESERVING;
        if (not_zero(area2)) {
            const float abs_area = fabsf(area2), len_os = length(os), len_ot = length(ot);
            const float s = (t.flag & ORIENT_PRESERVING) ? 1.0f : -1.0f;
            if (not_zero(len_os)) t.os = (s / len_os) * os;
            if (not_zero(len_ot)) t.ot = (s / len_ot) * ot;
            t.mag_s = len_os / abs_area; t.mag_t = len_ot / abs_area;
            if (not_zero(t.mag_s) && not_zero(t.mag_t)) t.flag &= ~GROUP_WITH_ANY;
        }
    }
    // 3. edge neighbours: edges keyed by the ordered welded index pair, paired first-come among unassigned ones
    {
        struct Edge { int i0, i1, f, e; };
        std::vector<Edge> edges; edges.reserve(3 * (size_t)nf);
        for (int f = 0; f < nf; ++f) {
            if (ti[f].flag & MARK_DEGENERATE) continue;
            for (int e = 0; e < 3; ++e) {
                const int a = tl[3 * f + e], b = tl[3 * f + (e < 2 ? e + 1 : 0)];
                edges.push_back({a < b ? a : b, a < b ? b : a, f, e});
            }
        }
        std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) {
            if (a.i0 != b.i0) return a.i0 < b.i0; if (a.i1 != b.i1) return a.i1 < b.i1; if (a.f != b.f) return a.f < b.f; return a.e < b.e; });
        for (size_t i = 0; i < edges.size(); ++i) {
            const Edge& ea = edges[i];
            if (ti[ea.f].neighbor[ea.e] != -1) continue;
            for (size_t j = i + 1; j < edges.size() && edges[j].i0 == ea.i0 && edges[j].i1 == ea.i1; ++j) {
                const Edge& eb = edges[j];
                if (eb.f == ea.f || ti[eb.f].neighbor[eb.e] != -1) continue;
                ti[ea.f].neighbor[ea.e] = eb.f; ti[eb.f].neighbor[eb.e] = ea.f;
                break;
            }
        }
    }
    // 4. groups
    std::vector<Group> groups;
    for (int f = 0; f < nf; ++f) {
        if (ti[f].flag & MARK_DEGENERATE) continue;
        for (int i = 0; i < 3; ++i) {
            if ((ti[f].flag & GROUP_WITH_ANY) || ti[f].group[i] != -1) continue;
            Group g; g.vertex = tl[3 * f + i]; g.orient = (ti[f].flag & ORIENT_PRESERVING) != 0;
            groups.push_back(g);
            const int gi = (int)groups.size() - 1;
            ti[f].group[i] = gi; groups[gi].faces.push_back(f);
            const int nl = ti[f].neighbor[i], nr = ti[f].neighbor[i > 0 ? i - 1 : 2];
            if (nl >= 0) assign_recur(tl, ti, groups, nl, gi);
            if (nr >= 0) assign_recur(tl, ti, groups, nr, gi);
        }
    }
    // 5. tangent spaces per (face, corner)
    std::vector<TSpace> ts(3 * (size_t)nf);
    const float thres_cos = cosf((180.0f * 3.14159265358979323846f) / 180.0f);
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const Group& g = groups[gi];
        std::vector<std::vector<int>> uni_members; std::vector<TSpace> uni_ts;
        for (int f : g.faces) {
            int index = -1;
            if (ti[f].group[0] == (int)gi) index = 0; else if (ti[f].group[1] == (int)gi) index = 1; else if (ti[f].group[2] == (int)gi) index = 2;
            if (index < 0) continue;
            const int vert = tl[3 * f + index];
            const V3 n = m.n((uint32_t)vert);
            V3 os = ti[f].os - dot(n, ti[f].os) * n, ot = ti[f].ot - dot(n, ti[f].ot) * n;
            if (v_not_zero(os)) os = normalize(os);
            if (v_not_zero(ot)) ot = normalize(ot);
            std::vector<int> members;
            for (int t : g.faces) {
                V3 os2 = ti[t].os - dot(n, ti[t].os) * n, ot2 = ti[t].ot - dot(n, ti[t].ot) * n;
                if (v_not_zero(os2)) os2 = normalize(os2);
                if (v_not_zero(ot2)) ot2 = normalize(ot2);
                const bool any = ((ti[f].flag | ti[t].flag) & GROUP_WITH_ANY) != 0;
                const bool same_face = f == t;
                const float cs = dot(os, os2), ct = dot(ot, ot2);
                if (any || same_face || (cs > thres_cos && ct > thres_cos)) members.push_back(t);
            }
            std::sort(members.begin(), members.end());
            size_t l = 0;
            for (; l < uni_members.size(); ++l) if (uni_members[l] == members) break;
            if (l == uni_members.size()) { uni_members.push_back(members); uni_ts.push_back(eval_tspace(members, tl, ti, m, g.vertex)); }
            TSpace out = uni_ts[l]; out.orient = g.orient;
            ts[3 * (size_t)f + index] = out;
        }
    }
    // 6. degenerate triangles copy the frame of a good triangle that uses the same (un-welded) vertex
    {
        std::map<int, size_t> first_good;    // welded vertex -> first (face, corner) among the non-degenerate triangles
        bool any_degenerate = false;
        for (int f = 0; f < nf; ++f) if (ti[f].flag & MARK_DEGENERATE) { any_degenerate = true; break; }
        if (any_degenerate) {
            for (int t = 0; t < nf; ++t) {
                if (ti[t].flag & MARK_DEGENERATE) continue;
                for (int j = 0; j < 3; ++j) first_good.emplace(tl[3 * t + j], 3 * (size_t)t + j);
            }
            for (int f = 0; f < nf; ++f) {
                if (!(ti[f].flag & MARK_DEGENERATE)) continue;
                for (int i = 0; i < 3; ++i) {
                    auto it = first_good.find(tl[3 * f + i]);
                    if (it != first_good.end()) ts[3 * (size_t)f + i] = ts[it->second];
                }
            }
        }
    }
    // output through the reference's Geometry impl: last face wins, inverted handedness convention
    for (int f = 0; f < nf; ++f)
        for (int i = 0; i < 3; ++i) {
            const TSpace& t = ts[3 * (size_t)f + i];
            float* o = tangents + (size_t)indices[3 * f + i] * stride_t;
            o[0] = t.os.x; o[1] = t.os.y; o[2] = t.os.z; o[3] = t.orient ? -1.0f : 1.0f;
        }
}

}  // namespace mikk
