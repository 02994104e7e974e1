// gltf_host.cpp — host layer above the C ABI: glTF import with the reference's rules, scene graph,
// animation sampling, skin matrices, camera and per-frame UBO fill.  See include/gltf_host.h for the
// reference functions each entry point mirrors.  CPU only; produces rt_scene_desc / rt_ubo.
#include "../include/gltf_host.h"
#include "mikktspace_gen.h"

#include <dirent.h>
#include <zlib.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }
struct Error { std::string msg; };
[[noreturn]] void die(const std::string& m) { throw Error{m}; }

// ---------------------------------------------------------------------------------------------------
// minimal JSON
// ---------------------------------------------------------------------------------------------------
struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;
    const Json* get(const char* k) const {
        if (type != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool has(const char* k) const { return get(k) != nullptr; }
    double number(const char* k, double def) const { auto j = get(k); return (j && j->type == Num) ? j->num : def; }
    int integer(const char* k, int def) const { auto j = get(k); return (j && j->type == Num) ? (int)j->num : def; }
    std::string string(const char* k, const std::string& def = "") const { auto j = get(k); return (j && j->type == Str) ? j->str : def; }
    size_t size() const { return type == Arr ? arr.size() : 0; }
    const Json& operator[](size_t i) const { if (type != Arr || i >= arr.size()) die("json: array index out of range"); return arr[i]; }
};

struct JsonParser {
    const char* p; const char* e;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
    Json parse() { ws(); Json j = value(); return j; }
    Json value() {
        ws(); if (p >= e) die("json: unexpected end");
        Json j;
        switch (*p) {
            case '{': {
                j.type = Json::Obj; ++p; ws();
                if (*p == '}') { ++p; return j; }
                for (;;) {
                    ws(); std::string k = str(); ws(); if (*p != ':') die("json: expected ':'"); ++p;
                    j.obj.emplace_back(std::move(k), value()); ws();
                    if (*p == ',') { ++p; continue; }
                    if (*p == '}') { ++p; break; }
                    die("json: expected ',' or '}'");
                }
                return j;
            }
            case '[': {
                j.type = Json::Arr; ++p; ws();
                if (*p == ']') { ++p; return j; }
                for (;;) {
                    j.arr.push_back(value()); ws();
                    if (*p == ',') { ++p; continue; }
                    if (*p == ']') { ++p; break; }
                    die("json: expected ',' or ']'");
                }
                return j;
            }
            case '"': j.type = Json::Str; j.str = str(); return j;
            case 't': if (e - p >= 4 && !strncmp(p, "true", 4)) { p += 4; j.type = Json::Bool; j.b = true; return j; } break;
            case 'f': if (e - p >= 5 && !strncmp(p, "false", 5)) { p += 5; j.type = Json::Bool; j.b = false; return j; } break;
            case 'n': if (e - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return j; } break;
            default: {
                char* end = nullptr; j.num = strtod(p, &end);
                if (end == p) die("json: bad token");
                p = end; j.type = Json::Num; return j;
            }
        }
        die("json: bad literal");
    }
    std::string str() {
        if (*p != '"') die("json: expected string"); ++p;
        std::string s;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                ++p; if (p >= e) break;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': {
                        unsigned cp = 0; for (int i = 1; i <= 4 && p + i < e; ++i) { char c = p[i]; cp = cp * 16 + (c <= '9' ? c - '0' : (c | 32) - 'a' + 10); }
                        p += 4;
                        if (cp < 0x80) s += (char)cp; else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
                        else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
                        break;
                    }
                    default: s += *p;
                }
                ++p;
            } else s += *p++;
        }
        if (p >= e) die("json: unterminated string"); ++p;
        return s;
    }
};

// ---------------------------------------------------------------------------------------------------
// small fp32 linear algebra (column-major 4x4, glam-compatible operation order)
// ---------------------------------------------------------------------------------------------------
struct Mat4 { float m[16]; };
Mat4 identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
// glam Mat4::mul_mat4: each column = ((a.c0*b.x + a.c1*b.y) + a.c2*b.z) + a.c3*b.w
Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int i = 0; i < 4; ++i)
            r.m[c * 4 + i] = ((a.m[i] * b.m[c * 4] + a.m[4 + i] * b.m[c * 4 + 1]) + a.m[8 + i] * b.m[c * 4 + 2]) + a.m[12 + i] * b.m[c * 4 + 3];
    return r;
}
void mul_vec4(const Mat4& a, const float v[4], float out[4]) {
    for (int i = 0; i < 4; ++i) out[i] = ((a.m[i] * v[0] + a.m[4 + i] * v[1]) + a.m[8 + i] * v[2]) + a.m[12 + i] * v[3];
}
Mat4 from_trs(const float t[3], const float q[4], const float s[3]) {
    // gltf crate Transform::matrix(): T * R * S, rotation from unit quaternion (x,y,z,w)
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2, wx = w * x2, wy = w * y2, wz = w * z2;
    Mat4 r{};
    r.m[0] = (1.0f - (yy + zz)) * s[0]; r.m[1] = (xy + wz) * s[0]; r.m[2] = (xz - wy) * s[0]; r.m[3] = 0;
    r.m[4] = (xy - wz) * s[1]; r.m[5] = (1.0f - (xx + zz)) * s[1]; r.m[6] = (yz + wx) * s[1]; r.m[7] = 0;
    r.m[8] = (xz + wy) * s[2]; r.m[9] = (yz - wx) * s[2]; r.m[10] = (1.0f - (xx + yy)) * s[2]; r.m[11] = 0;
    r.m[12] = t[0]; r.m[13] = t[1]; r.m[14] = t[2]; r.m[15] = 1.0f;
    return r;
}
bool inverse(const float* m, float* out) {
    // general 4x4 inverse via cofactors, evaluated in double and rounded once (nalgebra try_inverse surrogate)
    double a[16]; for (int i = 0; i < 16; ++i) a[i] = m[i];
    double inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0.0) return false;
    double id = 1.0 / det;
    for (int i = 0; i < 16; ++i) out[i] = (float)(inv[i] * id);
    return true;
}

// ---------------------------------------------------------------------------------------------------
// files, base64, PNG
// ---------------------------------------------------------------------------------------------------
std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) die("cannot open " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
std::vector<uint8_t> base64_decode(const std::string& s, size_t start) {
    std::vector<uint8_t> out; out.reserve((s.size() - start) * 3 / 4);
    uint32_t acc = 0; int bits = 0;
    for (size_t i = start; i < s.size(); ++i) {
        char c = s[i]; int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52; else if (c == '+' || c == '-') v = 62; else if (c == '/' || c == '_') v = 63; else continue;
        acc = (acc << 6) | v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return out;
}
std::string url_decode(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size()) { o += (char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16); i += 2; } else o += s[i];
    }
    return o;
}

struct DecodedImage { std::vector<uint8_t> rgba; uint32_t w = 0, h = 0; };

bool decode_png(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || memcmp(d, sig, 8)) { err = "not a PNG"; return false; }
    size_t p = 8; uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    auto be32 = [&](size_t o) { return (uint32_t)d[o] << 24 | (uint32_t)d[o + 1] << 16 | (uint32_t)d[o + 2] << 8 | d[o + 3]; };
    while (p + 12 <= n) {
        uint32_t len = be32(p); const char* ty = (const char*)d + p + 4; const uint8_t* body = d + p + 8;
        if (p + 12 + len > n) { err = "truncated PNG"; return false; }
        if (!memcmp(ty, "IHDR", 4)) { w = be32(p + 8); h = be32(p + 12); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (!memcmp(ty, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(ty, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(ty, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(ty, "IEND", 4)) break;
        p += 12 + len;
    }
    if (depth != 8) { err = "PNG bit depth " + std::to_string(depth) + " unsupported (reference: Error::Support 16 bytes images)"; return false; }
    if (interlace) { err = "interlaced PNG unsupported"; return false; }
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) { err = "bad PNG colour type"; return false; }
    size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rl = raw.size();
    if (uncompress(raw.data(), &rl, idat.data(), idat.size()) != Z_OK || rl != raw.size()) { err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* src = &raw[(stride + 1) * y]; uint8_t ft = src[0]; ++src;
        uint8_t* dst = &img[stride * y]; const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            int a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0, v = src[x];
            switch (ft) {
                case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break;
                case 4: { int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: err = "bad PNG filter"; return false;
            }
            dst[x] = (uint8_t)v;
        }
    }
    out.w = w; out.h = h; out.rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t* o = &out.rgba[i * 4]; const uint8_t* s = &img[i * ch];
        switch (ctype) {
            case 0: o[0] = o[1] = o[2] = s[0]; o[3] = 255; break;                         // R8 -> (l,l,l,255)   image.rs:154
            case 4: o[0] = o[1] = o[2] = s[0]; o[3] = s[1]; break;                        // R8G8 -> (l,l,l,a)   image.rs:156-161
            case 2: o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = 255; break;             // R8G8B8              image.rs:162-167
            case 6: o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; break;
            case 3: { size_t k = s[0]; o[0] = k * 3 + 2 < plte.size() ? plte[k * 3] : 0; o[1] = k * 3 + 2 < plte.size() ? plte[k * 3 + 1] : 0;
                      o[2] = k * 3 + 2 < plte.size() ? plte[k * 3 + 2] : 0; o[3] = k < trns.size() ? trns[k] : 255; break; }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// JPEG (ITU-T T.81: baseline and progressive Huffman, 8 bit, 1 or 3 components)
// The reference decodes through the `image` 0.24.6 crate -> `jpeg-decoder` 0.3.0 (Cargo.lock; not under
// /root/reference).  T.81 leaves IDCT precision, chroma upsampling and colour conversion to the decoder; this one
// uses the LL&M 13-bit integer IDCT, triangle-filter ("fancy") upsampling for 2x1 / 2x2 chroma and the JFIF
// fixed-point YCbCr->RGB conversion, i.e. libjpeg's defaults, so it can be checked bit-for-bit against
// libjpeg-turbo (tests/test_host.py); against jpeg-decoder a few texels may differ by 1-2 LSB.
// ---------------------------------------------------------------------------------------------------
namespace jpg {
static const uint8_t ZZ[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
struct Huff {
    bool present = false;
    uint8_t vals[256]; int maxcode[18], valptr[17], mincode[17];
    uint16_t fast[512];   // 9-bit prefix -> (length << 8 | symbol), 0 = longer code
    void build(const uint8_t* counts, const uint8_t* symbols) {
        present = true; memset(fast, 0, sizeof fast);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k; mincode[len] = code;
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                vals[k] = symbols[k];
                if (len <= 9) { const int lo = code << (9 - len); for (int f = 0; f < (1 << (9 - len)); ++f) fast[lo + f] = (uint16_t)(len << 8 | symbols[k]); }
            }
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
    }
};
struct Comp { int id = 0, h = 1, v = 1, tq = 0, bw = 0, bh = 0, pw = 0, ph = 0, dc_pred = 0, td = 0, ta = 0; std::vector<int16_t> coef; std::vector<uint8_t> plane; };
struct Bits {
    const uint8_t* d; size_t n, p; uint32_t acc = 0; int cnt = 0; bool hit_marker = false;
    void fill() {
        while (cnt <= 24) {
            uint32_t b = 0;
            if (!hit_marker && p < n) {
                b = d[p];
                if (b == 0xFF) {
                    const uint8_t nx = p + 1 < n ? d[p + 1] : 0xD9;
                    if (nx == 0) p += 2; else { hit_marker = true; b = 0; }
                } else ++p;
            }
            acc |= b << (24 - cnt); cnt += 8;
        }
    }
    int peek(int k) { if (cnt < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; cnt -= k; }
    int get(int k) { if (!k) return 0; const int v = peek(k); skip(k); return v; }
    void reset() { acc = 0; cnt = 0; hit_marker = false; }
};
static inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
static int decode_sym(Bits& b, const Huff& h) {
    const int look = b.peek(9); const uint16_t f = h.fast[look];
    if (f) { b.skip(f >> 8); return f & 0xFF; }
    int code = b.peek(16), len = 10;
    for (; len <= 16; ++len) { const int c = code >> (16 - len); if (c <= h.maxcode[len] && h.maxcode[len] >= 0) { b.skip(len); return h.vals[h.valptr[len] + c - h.mincode[len]]; } }
    b.skip(16); return 0;
}
// LL&M integer inverse DCT (13-bit constants, two passes), the "slow-but-accurate" variant
static void idct(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
    const int CB = 13, P1 = 2;
    const long F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
    long ws[64];
    for (int c = 0; c < 8; ++c) {
        const int16_t* ip = in + c; const uint16_t* qp = q + c; long* wp = ws + c;
        if (!ip[8] && !ip[16] && !ip[24] && !ip[32] && !ip[40] && !ip[48] && !ip[56]) {
            const long dc = (long)ip[0] * qp[0] * (1 << P1);
            for (int r = 0; r < 8; ++r) wp[8 * r] = dc;
            continue;
        }
        long z2 = (long)ip[16] * qp[16], z3 = (long)ip[48] * qp[48];
        long z1 = (z2 + z3) * F0541, tmp2 = z1 - z3 * F1847, tmp3 = z1 + z2 * F0765;
        z2 = (long)ip[0] * qp[0]; z3 = (long)ip[32] * qp[32];
        long tmp0 = (z2 + z3) * (1L << CB), tmp1 = (z2 - z3) * (1L << CB);
        const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
        tmp0 = (long)ip[56] * qp[56]; tmp1 = (long)ip[40] * qp[40]; tmp2 = (long)ip[24] * qp[24]; tmp3 = (long)ip[8] * qp[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3; const long z5 = (z3 + z4) * F1175;
        tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390; z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int sh = CB - P1; const long rnd = 1L << (sh - 1);
        wp[0] = (t10 + tmp3 + rnd) >> sh; wp[56] = (t10 - tmp3 + rnd) >> sh; wp[8] = (t11 + tmp2 + rnd) >> sh; wp[48] = (t11 - tmp2 + rnd) >> sh;
        wp[16] = (t12 + tmp1 + rnd) >> sh; wp[40] = (t12 - tmp1 + rnd) >> sh; wp[24] = (t13 + tmp0 + rnd) >> sh; wp[32] = (t13 - tmp0 + rnd) >> sh;
    }
    for (int r = 0; r < 8; ++r) {
        const long* wp = ws + 8 * r; uint8_t* op = out + (size_t)r * stride;
        long z2 = wp[2], z3 = wp[6];
        long z1 = (z2 + z3) * F0541, tmp2 = z1 - z3 * F1847, tmp3 = z1 + z2 * F0765;
        long tmp0 = (wp[0] + wp[4]) * (1L << CB), tmp1 = (wp[0] - wp[4]) * (1L << CB);
        const long t10 = tmp0 + tmp3, t13 = tmp0 - tmp3, t11 = tmp1 + tmp2, t12 = tmp1 - tmp2;
        tmp0 = wp[7]; tmp1 = wp[5]; tmp2 = wp[3]; tmp3 = wp[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; long z4 = tmp1 + tmp3; const long z5 = (z3 + z4) * F1175;
        tmp0 *= F0298; tmp1 *= F2053; tmp2 *= F3072; tmp3 *= F1501;
        z1 *= -F0899; z2 *= -F2562; z3 *= -F1961; z4 *= -F0390; z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        const int sh = CB + P1 + 3; const long rnd = 1L << (sh - 1);
        auto px = [&](long v) { v = ((v + rnd) >> sh) + 128; return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
        op[0] = px(t10 + tmp3); op[7] = px(t10 - tmp3); op[1] = px(t11 + tmp2); op[6] = px(t11 - tmp2);
        op[2] = px(t12 + tmp1); op[5] = px(t12 - tmp1); op[3] = px(t13 + tmp0); op[4] = px(t13 - tmp0);
    }
}
}  // namespace jpg

bool decode_jpeg(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    using namespace jpg;
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) { err = "not a JPEG"; return false; }
    uint16_t qt[4][64]; bool qt_ok[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    std::vector<Comp> comps; int W = 0, H = 0, hmax = 1, vmax = 1, restart = 0; bool progressive = false, have_sof = false;
    int adobe_transform = -1;
    size_t p = 2;
    auto be16 = [&](size_t o) { return (int)d[o] << 8 | d[o + 1]; };
    while (p + 4 <= n) {
        if (d[p] != 0xFF) { ++p; continue; }
        const uint8_t m = d[p + 1];
        if (m == 0xFF) { ++p; continue; }
        if (m == 0xD9) break;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
        const size_t len = (size_t)be16(p + 2);
        if (len < 2 || p + 2 + len > n) { err = "truncated JPEG"; return false; }
        const uint8_t* b = d + p + 4; const size_t bl = len - 2;
        if (m == 0xDB) {
            for (size_t o = 0; o < bl;) {
                const int pq = b[o] >> 4, tq = b[o] & 15; ++o;
                if (tq > 3 || o + (pq ? 128u : 64u) > bl) { err = "bad JPEG DQT"; return false; }
                for (int i = 0; i < 64; ++i) { qt[tq][ZZ[i]] = (uint16_t)(pq ? (b[o] << 8 | b[o + 1]) : b[o]); o += pq ? 2 : 1; }
                qt_ok[tq] = true;
            }
        } else if (m == 0xC4) {
            for (size_t o = 0; o + 17 <= bl;) {
                const int tc = b[o] >> 4, th = b[o] & 15; int total = 0;
                for (int i = 0; i < 16; ++i) total += b[o + 1 + i];
                if (th > 3 || tc > 1 || total > 256 || o + 17 + total > bl) { err = "bad JPEG DHT"; return false; }
                (tc ? hac[th] : hdc[th]).build(b + o + 1, b + o + 17);
                o += 17 + total;
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (bl < 6 || b[0] != 8) { err = "JPEG sample precision unsupported"; return false; }
            progressive = m == 0xC2; H = be16(p + 5); W = be16(p + 7);
            const int nc = b[5];
            if ((nc != 1 && nc != 3) || bl < (size_t)(6 + 3 * nc) || !W || !H) { err = "JPEG component count unsupported"; return false; }
            comps.resize(nc);
            for (int i = 0; i < nc; ++i) { comps[i].id = b[6 + 3 * i]; comps[i].h = b[7 + 3 * i] >> 4; comps[i].v = b[7 + 3 * i] & 15; comps[i].tq = b[8 + 3 * i] & 3;
                                           if (comps[i].h < 1 || comps[i].h > 4 || comps[i].v < 1 || comps[i].v > 4) { err = "bad JPEG sampling"; return false; }
                                           hmax = std::max(hmax, comps[i].h); vmax = std::max(vmax, comps[i].v); }
            const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
            for (auto& c : comps) {
                c.bw = ((W * c.h + hmax - 1) / hmax + 7) / 8; c.bh = ((H * c.v + vmax - 1) / vmax + 7) / 8;   // blocks covering the component
                c.pw = mcux * c.h; c.ph = mcuy * c.v;                                                          // padded to whole MCUs
                c.coef.assign((size_t)c.pw * c.ph * 64, 0);
            }
            have_sof = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            err = "JPEG coding process unsupported (lossless / arithmetic / hierarchical)"; return false;
        } else if (m == 0xDD) {
            restart = be16(p + 4);
        } else if (m == 0xEE && bl >= 12 && !memcmp(b, "Adobe", 5)) {
            adobe_transform = b[11];
        } else if (m == 0xDA) {
            if (!have_sof) { err = "JPEG SOS before SOF"; return false; }
            const int ns = b[0];
            if (ns < 1 || ns > (int)comps.size() || bl < (size_t)(4 + 2 * ns)) { err = "bad JPEG SOS"; return false; }
            Comp* sc[4];
            for (int i = 0; i < ns; ++i) {
                sc[i] = nullptr;
                for (auto& c : comps) if (c.id == b[1 + 2 * i]) sc[i] = &c;
                if (!sc[i]) { err = "bad JPEG SOS component"; return false; }
                sc[i]->td = b[2 + 2 * i] >> 4; sc[i]->ta = b[2 + 2 * i] & 15;
                if (sc[i]->td > 3 || sc[i]->ta > 3) { err = "bad JPEG SOS table"; return false; }
            }
            int Ss = b[1 + 2 * ns], Se = b[2 + 2 * ns]; const int Ah = b[3 + 2 * ns] >> 4, Al = b[3 + 2 * ns] & 15;
            if (!progressive) { Ss = 0; Se = 63; }
            if (Ss > Se || Se > 63 || (progressive && Ss == 0 && Se != 0) || (progressive && Ss > 0 && ns != 1)) { err = "bad JPEG spectral selection"; return false; }
            Bits br{d, n, p + 2 + len};
            int eobrun = 0, todo = restart;
            for (auto& c : comps) c.dc_pred = 0;
            auto block = [&](Comp& c, int bx, int by) -> bool {
                int16_t* blk = &c.coef[((size_t)by * c.pw + bx) * 64];
                if (!progressive) {
                    const Huff &hd = hdc[c.td], &ha = hac[c.ta];
                    if (!hd.present || !ha.present) return false;
                    const int s = decode_sym(br, hd);
                    c.dc_pred += s ? extend(br.get(s), s) : 0;
                    blk[0] = (int16_t)c.dc_pred;
                    for (int k = 1; k < 64;) {
                        const int rs = decode_sym(br, ha), r = rs >> 4, sz = rs & 15;
                        if (!sz) { if (r != 15) break; k += 16; continue; }
                        k += r; if (k > 63) break;
                        blk[ZZ[k++]] = (int16_t)extend(br.get(sz), sz);
                    }
                    return true;
                }
                if (Ss == 0) {                                   // DC scan
                    if (Ah == 0) {
                        const Huff& hd = hdc[c.td]; if (!hd.present) return false;
                        const int s = decode_sym(br, hd);
                        c.dc_pred += s ? extend(br.get(s), s) : 0;
                        blk[0] = (int16_t)(c.dc_pred * (1 << Al));
                    } else if (br.get(1)) blk[0] |= (int16_t)(1 << Al);
                    return true;
                }
                const Huff& ha = hac[c.ta]; if (!ha.present) return false;
                if (Ah == 0) {                                   // AC first pass
                    if (eobrun > 0) { --eobrun; return true; }
                    for (int k = Ss; k <= Se; ++k) {
                        const int rs = decode_sym(br, ha), r = rs >> 4, sz = rs & 15;
                        if (sz) { k += r; if (k > 63) break; blk[ZZ[k]] = (int16_t)(extend(br.get(sz), sz) * (1 << Al)); }
                        else if (r == 15) k += 15;
                        else { eobrun = (1 << r) + (r ? br.get(r) : 0) - 1; break; }
                    }
                    return true;
                }
                // AC refinement (T.81 G.1.2.3)
                const int p1 = 1 << Al, m1 = -(1 << Al);
                int k = Ss;
                auto refine = [&](int16_t& cf) { if (br.get(1) && !(cf & p1)) cf = (int16_t)(cf + (cf >= 0 ? p1 : m1)); };
                if (eobrun == 0) {
                    for (; k <= Se; ++k) {
                        const int rs = decode_sym(br, ha); int r = rs >> 4; int s = rs & 15;
                        if (s) s = br.get(1) ? p1 : m1;
                        else if (r != 15) { eobrun = (1 << r) + (r ? br.get(r) : 0); break; }
                        for (; k <= Se; ++k) {
                            int16_t& cf = blk[ZZ[k]];
                            if (cf) refine(cf); else if (--r < 0) break;
                        }
                        if (s && k <= Se) blk[ZZ[k]] = (int16_t)s;
                    }
                }
                if (eobrun > 0) {
                    for (; k <= Se; ++k) { int16_t& cf = blk[ZZ[k]]; if (cf) refine(cf); }
                    --eobrun;
                }
                return true;
            };
            auto restart_check = [&]() -> bool {
                if (!restart) return true;
                if (--todo > 0) return true;
                // byte-align, expect RSTn
                br.reset();
                while (br.p + 1 < n && !(d[br.p] == 0xFF && d[br.p + 1] >= 0xD0 && d[br.p + 1] <= 0xD7)) {
                    if (d[br.p] == 0xFF && d[br.p + 1] != 0 && d[br.p + 1] != 0xFF) return true;   // some other marker: let the outer parser see it
                    ++br.p;
                }
                if (br.p + 1 < n) br.p += 2;
                todo = restart; eobrun = 0;
                for (auto& c : comps) c.dc_pred = 0;
                return true;
            };
            bool ok = true;
            if (ns == 1) {
                Comp& c = *sc[0];
                for (int by = 0; by < c.bh && ok; ++by) for (int bx = 0; bx < c.bw && ok; ++bx) { ok = block(c, bx, by); restart_check(); }
            } else {
                const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
                for (int my = 0; my < mcuy && ok; ++my) for (int mx = 0; mx < mcux && ok; ++mx) {
                    for (int i = 0; i < ns && ok; ++i) for (int v = 0; v < sc[i]->v && ok; ++v) for (int h = 0; h < sc[i]->h && ok; ++h)
                        ok = block(*sc[i], mx * sc[i]->h + h, my * sc[i]->v + v);
                    restart_check();
                }
            }
            if (!ok) { err = "JPEG scan references a missing Huffman table"; return false; }
            // continue after the entropy-coded segment: next marker that is not RSTn / stuffing
            size_t q = br.p < p + 2 + len ? p + 2 + len : br.p;          // the bit reader never steps over a marker
            while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0 && d[q + 1] != 0xFF && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7))) ++q;
            p = q; continue;
        }
        p += 2 + len;
    }
    if (!have_sof) { err = "JPEG without frame header"; return false; }
    for (auto& c : comps) {
        if (!qt_ok[c.tq]) { err = "JPEG quantisation table missing"; return false; }
        const int pw = c.pw * 8;
        c.plane.assign((size_t)pw * c.ph * 8, 0);
        for (int by = 0; by < c.ph; ++by) for (int bx = 0; bx < c.pw; ++bx)
            idct(&c.coef[((size_t)by * c.pw + bx) * 64], qt[c.tq], &c.plane[(size_t)by * 8 * pw + bx * 8], pw);
        c.coef.clear(); c.coef.shrink_to_fit();
    }
    // chroma upsampling to full resolution
    std::vector<std::vector<uint8_t>> full(comps.size());
    for (size_t ci = 0; ci < comps.size(); ++ci) {
        Comp& c = comps[ci]; const int pw = c.pw * 8;
        const int cw = (W * c.h + hmax - 1) / hmax, chh = (H * c.v + vmax - 1) / vmax;     // real component size
        const int fx = hmax / c.h, fy = vmax / c.v;
        std::vector<uint8_t>& o = full[ci]; o.resize((size_t)W * H);
        auto at = [&](int x, int y) -> int { return c.plane[(size_t)std::min(std::max(y, 0), chh - 1) * pw + std::min(std::max(x, 0), cw - 1)]; };
        if (fx == 1 && fy == 1 && hmax % c.h == 0 && vmax % c.v == 0) {
            for (int y = 0; y < H; ++y) memcpy(&o[(size_t)y * W], &c.plane[(size_t)y * pw], W);
        } else if (fx == 2 && fy == 1 && hmax == 2 * c.h && vmax == c.v && cw > 2) {
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                const int i = x >> 1, cur = at(i, y);
                int v;
                if (x & 1) v = (i == cw - 1) ? cur : (3 * cur + at(i + 1, y) + 2) >> 2;
                else v = (i == 0) ? cur : (3 * cur + at(i - 1, y) + 1) >> 2;
                o[(size_t)y * W + x] = (uint8_t)v;
            }
        } else if (fx == 2 && fy == 2 && hmax == 2 * c.h && vmax == 2 * c.v && cw > 2) {
            for (int y = 0; y < H; ++y) {
                const int iy = y >> 1, oy = (y & 1) ? iy + 1 : iy - 1;     // nearer row iy (weight 3), farther row oy (weight 1)
                for (int x = 0; x < W; ++x) {
                    const int i = x >> 1;
                    const int cur = 3 * at(i, iy) + at(i, oy);
                    int v;
                    if (x & 1) v = (i == cw - 1) ? (cur * 4 + 7) >> 4 : (cur * 3 + 3 * at(i + 1, iy) + at(i + 1, oy) + 7) >> 4;
                    else v = (i == 0) ? (cur * 4 + 8) >> 4 : (cur * 3 + 3 * at(i - 1, iy) + at(i - 1, oy) + 8) >> 4;
                    o[(size_t)y * W + x] = (uint8_t)v;
                }
            }
        } else {
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) o[(size_t)y * W + x] = (uint8_t)at(x * c.h / hmax, y * c.v / vmax);
        }
        c.plane.clear(); c.plane.shrink_to_fit();
    }
    out.w = (uint32_t)W; out.h = (uint32_t)H; out.rgba.resize((size_t)W * H * 4);
    const bool ycc = comps.size() == 3 && adobe_transform != 0;
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        uint8_t* o = &out.rgba[i * 4];
        if (comps.size() == 1) { o[0] = o[1] = o[2] = full[0][i]; }
        else if (!ycc) { o[0] = full[0][i]; o[1] = full[1][i]; o[2] = full[2][i]; }
        else {
            const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
            auto cl = [](int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
            const int r = y + ((91881 * cr + 32768) >> 16);
            const int g = y + ((-22554 * cb - 46802 * cr + 32768) >> 16);
            const int bl = y + ((116130 * cb + 32768) >> 16);
            o[0] = cl(r); o[1] = cl(g); o[2] = cl(bl);
        }
        o[3] = 255;
    }
    return true;
}

bool decode_image(const uint8_t* d, size_t n, DecodedImage& out, std::string& err) {
    if (n >= 2 && d[0] == 0xFF && d[1] == 0xD8) return decode_jpeg(d, n, out, err);
    return decode_png(d, n, out, err);
}

// ---------------------------------------------------------------------------------------------------
// Doc
// ---------------------------------------------------------------------------------------------------
struct Transform { bool decomposed = false; Mat4 matrix = identity(); float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1}; };

struct Node {
    int skin = -1, light = -1, mesh = -1;
    std::vector<int> children;
    Transform local;
    Mat4 parent_cache = identity();
    Mat4 local_matrix() const { return local.decomposed ? from_trs(local.t, local.r, local.s) : local.matrix; }
    Mat4 world() const { return mul(parent_cache, local_matrix()); }
};
struct Aabb { float lo[3], hi[3]; };
struct Primitive { uint32_t geo_id; Aabb aabb; };
struct Mesh { std::vector<Primitive> prims; };
struct LightDef { float color[3]; int kind; float range, intensity; };
struct Channel { int target; int prop; /*0 T,1 R,2 S,3 morph*/ std::vector<float> input; std::vector<float> out; int comps; int interp; /*0 linear,1 step,2 cubic*/ };
struct Skin { std::vector<int> joints; std::vector<Mat4> ibm; };

}  // namespace

struct gv_doc {
    int current_scene = 0;
    std::vector<std::vector<int>> scenes;
    std::vector<Node> nodes;
    std::vector<Mesh> meshes;
    std::vector<rt_material> materials;
    std::vector<LightDef> lights;
    std::vector<Channel> channels;
    std::vector<Skin> skins;
    // flat arrays (GeoBuilder)
    std::vector<rt_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<rt_prim_info> prim_infos;
    std::vector<rt_geometry> geometries;
    // textures
    std::vector<DecodedImage> images; std::vector<uint32_t> image_srgb;
    std::vector<rt_image_desc> image_descs;
    std::vector<rt_sampler_desc> samplers;
    std::vector<rt_texture_desc> textures;
    // derived
    std::vector<rt_instance> instances;
    std::vector<rt_light> dlights, plights;
    std::vector<float> skin_mats;
    Mat4 aabb_trans = identity();
    bool animated = false;
    std::vector<uint8_t> sky[6]; uint32_t sky_w = 0, sky_h = 0, sky_srgb = 1;
};

namespace {

struct Loader {
    gv_doc& doc; Json root; std::string dir;
    std::vector<std::vector<uint8_t>> buffers;
    std::vector<uint8_t> glb_bin;

    explicit Loader(gv_doc& d) : doc(d) {}

    void load(const std::string& path) {
        std::vector<uint8_t> file = read_file(path);
        size_t slash = path.find_last_of('/'); dir = slash == std::string::npos ? "." : path.substr(0, slash);
        std::string json;
        if (file.size() >= 12 && !memcmp(file.data(), "glTF", 4)) {
            size_t p = 12;
            while (p + 8 <= file.size()) {
                uint32_t len, type; memcpy(&len, &file[p], 4); memcpy(&type, &file[p + 4], 4);
                if (p + 8 + len > file.size()) die("truncated GLB");
                if (type == 0x4E4F534A) json.assign((const char*)&file[p + 8], len);
                else if (type == 0x004E4942) glb_bin.assign(file.begin() + p + 8, file.begin() + p + 8 + len);
                p += 8 + len;
            }
        } else json.assign((const char*)file.data(), file.size());
        JsonParser jp{json.data(), json.data() + json.size()};
        root = jp.parse();
        if (root.type != Json::Obj) die("glTF root is not an object");
        load_buffers();
        build();
    }

    std::vector<uint8_t> load_uri(const std::string& uri) {
        if (uri.rfind("data:", 0) == 0) { size_t c = uri.find(','); if (c == std::string::npos) die("bad data uri"); return base64_decode(uri, c + 1); }
        return read_file(dir + "/" + url_decode(uri));
    }
    void load_buffers() {
        const Json* bs = root.get("buffers");
        for (size_t i = 0; bs && i < bs->size(); ++i) {
            const Json& b = (*bs)[i];
            if (b.has("uri")) buffers.push_back(load_uri(b.string("uri"))); else buffers.push_back(glb_bin);
            if (buffers.back().size() < (size_t)b.number("byteLength", 0)) die("buffer shorter than byteLength");
        }
    }

    // ---- accessors (gltf crate accessor::util semantics; sparse accessors are not supported) ----
    struct View { const uint8_t* base; size_t stride; size_t count; int ctype; int ncomp; bool normalized; };
    static int ncomp_of(const std::string& t) { return t == "SCALAR" ? 1 : t == "VEC2" ? 2 : t == "VEC3" ? 3 : t == "VEC4" ? 4 : t == "MAT4" ? 16 : t == "MAT3" ? 9 : t == "MAT2" ? 4 : 0; }
    static int csize(int ct) { return (ct == 5120 || ct == 5121) ? 1 : (ct == 5122 || ct == 5123) ? 2 : 4; }
    View view(int acc) {
        const Json* as = root.get("accessors");
        if (!as || acc < 0 || (size_t)acc >= as->size()) die("accessor index out of range");
        const Json& a = (*as)[acc];
        if (a.has("sparse")) die("sparse accessors unsupported");
        View v; v.ctype = a.integer("componentType", 5126); v.ncomp = ncomp_of(a.string("type")); v.count = (size_t)a.number("count", 0);
        v.normalized = a.get("normalized") && a.get("normalized")->b;
        int bv = a.integer("bufferView", -1); if (bv < 0) die("accessor without bufferView unsupported");
        const Json* bvs = root.get("bufferViews");
        if (!bvs) die("accessor refers to a bufferView but the file has none");
        const Json& b = (*bvs)[(size_t)bv];
        size_t off = (size_t)b.number("byteOffset", 0) + (size_t)a.number("byteOffset", 0);
        size_t elem = (size_t)csize(v.ctype) * v.ncomp;
        v.stride = (size_t)b.number("byteStride", 0); if (!v.stride) v.stride = elem;
        const int bi = b.integer("buffer", 0);
        if (bi < 0 || (size_t)bi >= buffers.size()) die("bufferView refers to a missing buffer");
        const std::vector<uint8_t>& buf = buffers[(size_t)bi];
        if (v.ncomp <= 0) die("accessor with an unknown type");
        if (v.count && (v.count > buf.size() || off > buf.size() || v.stride > buf.size() || off + v.stride * (v.count - 1) + elem > buf.size())) die("accessor out of buffer range");
        v.base = buf.data() + off; return v;
    }
    static float comp_f32(const View& v, size_t i, int c, bool normalize_ints) {
        if (c >= v.ncomp || i >= v.count) return 0.0f;     // view() only vouches for ncomp components of count elements
        const uint8_t* p = v.base + i * v.stride + (size_t)c * csize(v.ctype);
        switch (v.ctype) {
            case 5126: { float f; memcpy(&f, p, 4); return f; }
            case 5121: return normalize_ints ? p[0] / 255.0f : (float)p[0];
            case 5123: { uint16_t u; memcpy(&u, p, 2); return normalize_ints ? u / 65535.0f : (float)u; }
            case 5120: { int8_t s = (int8_t)p[0]; return normalize_ints ? std::max(s / 127.0f, -1.0f) : (float)s; }
            case 5122: { int16_t s; memcpy(&s, p, 2); return normalize_ints ? std::max(s / 32767.0f, -1.0f) : (float)s; }
            case 5125: { uint32_t u; memcpy(&u, p, 4); return (float)u; }
        }
        return 0;
    }
    static uint32_t comp_u32(const View& v, size_t i, int c) {
        if (c >= v.ncomp || i >= v.count) return 0u;
        const uint8_t* p = v.base + i * v.stride + (size_t)c * csize(v.ctype);
        switch (v.ctype) {
            case 5121: return p[0];
            case 5123: { uint16_t u; memcpy(&u, p, 2); return u; }
            case 5125: { uint32_t u; memcpy(&u, p, 4); return u; }
        }
        return 0;
    }

    static rt_texture_info tex_info(const Json* j) {   // material.rs:33-56: index + 1, -1 = none
        rt_texture_info t{-1, -1};
        if (j && j->type == Json::Obj && j->has("index")) { t.index = 1 + j->integer("index", 0); t.coord = j->integer("texCoord", 0); }
        return t;
    }
    static void arr_to(const Json* j, float* out, int n) { if (j && j->type == Json::Arr) for (int i = 0; i < n && (size_t)i < j->size(); ++i) out[i] = (float)(*j)[i].num; }

    rt_material make_material(const Json& m) {   // material.rs:162-190, 332-380
        rt_material r; memset(&r, 0, sizeof r);
        const Json* ext = m.get("extensions");
        std::string am = m.string("alphaMode", "OPAQUE");
        r.alpha_mode = am == "MASK" ? 2 : am == "BLEND" ? 3 : 1;
        r.alpha_cutoff = (float)m.number("alphaCutoff", 0.5);
        r.double_sided = m.get("doubleSided") && m.get("doubleSided")->b;
        const Json* pbr = m.get("pbrMetallicRoughness"); Json empty; empty.type = Json::Obj; if (!pbr) pbr = &empty;
        float bc[4] = {1, 1, 1, 1}; arr_to(pbr->get("baseColorFactor"), bc, 4); memcpy(r.base_color, bc, 16);
        r.base_color_texture = tex_info(pbr->get("baseColorTexture"));
        r.metallic_factor = (float)pbr->number("metallicFactor", 1.0); r.roughness_factor = (float)pbr->number("roughnessFactor", 1.0);
        r.metallic_roughness_texture = tex_info(pbr->get("metallicRoughnessTexture"));
        r.normal_texture = tex_info(m.get("normalTexture")); r.emissive_texture = tex_info(m.get("emissiveTexture"));
        float em[3] = {0, 0, 0}; arr_to(m.get("emissiveFactor"), em, 3); r.emissive_factor[0] = em[0]; r.emissive_factor[1] = em[1]; r.emissive_factor[2] = em[2]; r.emissive_factor[3] = 0;
        r.occlusion_texture = tex_info(m.get("occlusionTexture"));
        r.ior = 1.5f; r.unlit = 0;
        // defaults: TransmissionInfo::default (texture fields zero!), VolumeInfo::default, SpecularInfo::default, SpecularGlossiness::default
        r.transmission_texture = rt_texture_info{-1, -1};   // derive(Default) on TransmissionInfo uses TextureInfo::default = (-1,-1)
        r.attenuation_color[0] = r.attenuation_color[1] = r.attenuation_color[2] = 1.0f; r.thickness_texture = rt_texture_info{-1, -1};
        r.attenuation_distance = 3.402823466e+38f;
        r.specular_texture = rt_texture_info{-1, -1}; r.specular_color_texture = rt_texture_info{-1, -1};
        for (int i = 0; i < 4; ++i) r.specular_color_factor[i] = 1.0f; r.specular_factor = 1.0f;
        r.sg_diffuse_texture = rt_texture_info{-1, -1}; r.sg_specular_glossiness_texture = rt_texture_info{-1, -1};
        if (ext) {
            if (const Json* e = ext->get("KHR_materials_ior")) r.ior = (float)e->number("ior", 1.5);
            if (ext->has("KHR_materials_unlit")) r.unlit = 1;
            if (const Json* e = ext->get("KHR_materials_transmission")) {
                r.transmission_exist = 1; r.transmission_factor = (float)e->number("transmissionFactor", 0.0);
                r.transmission_texture = tex_info(e->get("transmissionTexture"));
            }
            if (const Json* e = ext->get("KHR_materials_volume")) {
                r.volume_exists = 1; r.thickness_factor = (float)e->number("thicknessFactor", 0.0);
                r.thickness_texture = tex_info(e->get("thicknessTexture"));
                r.attenuation_distance = e->has("attenuationDistance") ? (float)e->number("attenuationDistance", 0) : INFINITY;
                float ac[3] = {1, 1, 1}; arr_to(e->get("attenuationColor"), ac, 3); memcpy(r.attenuation_color, ac, 12);
            }
            if (const Json* e = ext->get("KHR_materials_specular")) {
                r.specular_exist = 1; r.specular_factor = (float)e->number("specularFactor", 1.0);
                float sc[3] = {1, 1, 1}; arr_to(e->get("specularColorFactor"), sc, 3); memcpy(r.specular_color_factor, sc, 12); r.specular_color_factor[3] = 0;
                r.specular_texture = tex_info(e->get("specularTexture")); r.specular_color_texture = tex_info(e->get("specularColorTexture"));
            }
            if (const Json* e = ext->get("KHR_materials_pbrSpecularGlossiness")) {
                r.workflow = 1;
                float df[4] = {1, 1, 1, 1}; arr_to(e->get("diffuseFactor"), df, 4); memcpy(r.sg_diffuse_factor, df, 16);
                float sf[3] = {1, 1, 1}; arr_to(e->get("specularFactor"), sf, 3); memcpy(r.sg_specular_glossiness_factor, sf, 12);
                r.sg_specular_glossiness_factor[3] = (float)e->number("glossinessFactor", 1.0);
                r.sg_diffuse_texture = tex_info(e->get("diffuseTexture")); r.sg_specular_glossiness_texture = tex_info(e->get("specularGlossinessTexture"));
            }
        }
        return r;
    }

    void build() {
        // scenes / nodes (scene_graph.rs:105-113)
        const Json* scenes = root.get("scenes");
        if (!scenes || !scenes->size()) die("No scene");
        doc.current_scene = root.integer("scene", 0);
        for (size_t i = 0; i < scenes->size(); ++i) {
            std::vector<int> r; const Json* ns = (*scenes)[i].get("nodes");
            for (size_t k = 0; ns && k < ns->size(); ++k) r.push_back((int)(*ns)[k].num);
            doc.scenes.push_back(r);
        }
        if ((size_t)doc.current_scene >= doc.scenes.size()) doc.current_scene = 0;
        const Json* nodes = root.get("nodes");
        for (size_t i = 0; nodes && i < nodes->size(); ++i) {
            const Json& n = (*nodes)[i]; Node nd;
            nd.skin = n.integer("skin", -1); nd.mesh = n.integer("mesh", -1);
            if (const Json* e = n.get("extensions")) if (const Json* l = e->get("KHR_lights_punctual")) nd.light = l->integer("light", -1);
            if (const Json* c = n.get("children")) for (size_t k = 0; k < c->size(); ++k) nd.children.push_back((int)(*c)[k].num);
            if (const Json* m = n.get("matrix")) { for (int k = 0; k < 16 && (size_t)k < m->size(); ++k) nd.local.matrix.m[k] = (float)(*m)[k].num; }
            else { nd.local.decomposed = true; arr_to(n.get("translation"), nd.local.t, 3); arr_to(n.get("rotation"), nd.local.r, 4); arr_to(n.get("scale"), nd.local.s, 3); }
            doc.nodes.push_back(nd);
        }
        // lights (light.rs:129-140)
        if (const Json* e = root.get("extensions")) if (const Json* l = e->get("KHR_lights_punctual")) if (const Json* ls = l->get("lights"))
            for (size_t i = 0; i < ls->size(); ++i) {
                const Json& L = (*ls)[i]; LightDef d; d.color[0] = d.color[1] = d.color[2] = 1.0f; arr_to(L.get("color"), d.color, 3);
                std::string ty = L.string("type", "point"); d.kind = ty == "directional" ? 0 : 1;   // spot treated as point (light.rs:51-57)
                d.range = L.has("range") ? (float)L.number("range", 0) : 3.402823466e+38f; d.intensity = (float)L.number("intensity", 1.0);
                doc.lights.push_back(d);
            }
        // materials
        const Json* mats = root.get("materials");
        for (size_t i = 0; mats && i < mats->size(); ++i) doc.materials.push_back(make_material((*mats)[i]));
        if (doc.materials.empty()) { Json e; e.type = Json::Obj; doc.materials.push_back(make_material(e)); }
        // animations are read before meshes in the reference, order is irrelevant here
        load_animations();
        // meshes -> flat geometry (geometry.rs:100-268)
        const Json* meshes = root.get("meshes");
        for (size_t i = 0; meshes && i < meshes->size(); ++i) {
            Mesh mesh; const Json* prims = (*meshes)[i].get("primitives");
            for (size_t p = 0; prims && p < prims->size(); ++p) {
                const Json& pr = (*prims)[p]; const Json* at = pr.get("attributes");
                if (!at || !at->has("POSITION") || pr.integer("mode", 4) != 4) continue;   // is_primitive_supported
                mesh.prims.push_back(make_primitive(pr, *at));
            }
            doc.meshes.push_back(mesh);
        }
        load_textures();
        // skins (skinning.rs:19-37)
        const Json* skins = root.get("skins");
        for (size_t i = 0; skins && i < skins->size(); ++i) {
            const Json& s = (*skins)[i]; Skin sk; const Json* js = s.get("joints");
            for (size_t k = 0; js && k < js->size(); ++k) sk.joints.push_back((int)(*js)[k].num);
            int ibm = s.integer("inverseBindMatrices", -1); if (ibm < 0) die("skin without inverseBindMatrices (reference unwraps)");
            View v = view(ibm);
            for (size_t k = 0; k < v.count; ++k) { Mat4 m; for (int c = 0; c < 16; ++c) m.m[c] = comp_f32(v, k, c, false); sk.ibm.push_back(m); }
            if (sk.ibm.size() != sk.joints.size()) die("ibm/joint count mismatch");
            doc.skins.push_back(sk);
        }
        validate_graph();
        if (!doc.skins.empty()) tag_skinned_vertices();
        load_scene();
    }

    // The gltf crate validates indices at import and the reference unwraps / panics on what is left; here every index the
    // scene graph follows is checked once so that a malformed file is an error, never an out-of-bounds access.
    void validate_graph() {
        const int nn = (int)doc.nodes.size(), nm = (int)doc.meshes.size(), ns = (int)doc.skins.size();
        for (const auto& sc : doc.scenes) for (int r : sc) if (r < 0 || r >= nn) die("scene refers to a missing node");
        for (const Node& n : doc.nodes) {
            if (n.mesh < -1 || n.mesh >= nm) die("node refers to a missing mesh");
            if (n.skin < -1 || n.skin >= ns) die("node refers to a missing skin");
            for (int c : n.children) if (c < 0 || c >= nn) die("node refers to a missing child");
        }
        for (const Skin& sk : doc.skins) for (int j : sk.joints) if (j < 0 || j >= nn) die("skin refers to a missing joint node");
        for (const Channel& ch : doc.channels) if (ch.target < 0 || ch.target >= nn) die("animation channel targets a missing node");
        // the hierarchy must be a forest: every node reachable at most once from the scene roots
        std::vector<char> seen((size_t)nn, 0);
        std::vector<int> stack;
        for (const auto& sc : doc.scenes) {
            std::fill(seen.begin(), seen.end(), 0);
            for (int r : sc) stack.push_back(r);
            while (!stack.empty()) {
                const int i = stack.back(); stack.pop_back();
                if (seen[(size_t)i]) die("node hierarchy is not a tree (a node is reachable twice)");
                seen[(size_t)i] = 1;
                for (int c : doc.nodes[(size_t)i].children) stack.push_back(c);
            }
        }
    }

    Primitive make_primitive(const Json& pr, const Json& at) {
        int material_index = pr.integer("material", 0);   // DEFAULT_MATERIAL_INDEX geometry.rs:130,143
        if (material_index < 0 || (size_t)material_index >= doc.materials.size()) material_index = 0;
        uint32_t geo_id = (uint32_t)doc.prim_infos.size();
        View pos = view(at.integer("POSITION", -1));
        size_t nv = pos.count;
        std::vector<rt_vertex> verts(nv); memset(verts.data(), 0, nv * sizeof(rt_vertex));
        for (size_t i = 0; i < nv; ++i) for (int c = 0; c < 3; ++c) verts[i].position[c] = comp_f32(pos, i, c, false);
        std::vector<uint32_t> idx;
        if (pr.has("indices")) { View iv = view(pr.integer("indices", -1)); idx.resize(iv.count); for (size_t i = 0; i < iv.count; ++i) idx[i] = comp_u32(iv, i, 0); }
        else { idx.resize(nv); for (size_t i = 0; i < nv; ++i) idx[i] = (uint32_t)i; }
        for (uint32_t v : idx) if (v >= nv) die("vertex index out of range");      // the normal / tangent generators below index the vertices with it
        if (at.has("NORMAL")) { View v = view(at.integer("NORMAL", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) for (int c = 0; c < 3; ++c) verts[i].normal[c] = comp_f32(v, i, c, true); }
        else {   // create_geo_normal geometry.rs:270-290
            for (size_t t = 0; t + 2 < idx.size(); t += 3) {
                uint32_t i0 = idx[t], i1 = idx[t + 1], i2 = idx[t + 2];
                float a[3], b[3];
                for (int c = 0; c < 3; ++c) { a[c] = verts[i1].position[c] - verts[i0].position[c]; b[c] = verts[i2].position[c] - verts[i0].position[c]; }
                auto norm = [](float* v) { float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); float r = 1.0f / l; v[0] *= r; v[1] *= r; v[2] *= r; };
                norm(a); norm(b);
                float n[3] = {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]}; norm(n);
                for (uint32_t k : {i0, i1, i2}) if (k < nv) { verts[k].normal[0] = n[0]; verts[k].normal[1] = n[1]; verts[k].normal[2] = n[2]; }
            }
        }
        bool has_uv0 = at.has("TEXCOORD_0");
        if (has_uv0) { View v = view(at.integer("TEXCOORD_0", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) { verts[i].uv0[0] = comp_f32(v, i, 0, true); verts[i].uv0[1] = comp_f32(v, i, 1, true); } }
        if (at.has("TEXCOORD_1")) { View v = view(at.integer("TEXCOORD_1", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) { verts[i].uv1[0] = comp_f32(v, i, 0, true); verts[i].uv1[1] = comp_f32(v, i, 1, true); } }
        if (at.has("TANGENT")) { View v = view(at.integer("TANGENT", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) for (int c = 0; c < 4; ++c) verts[i].tangent[c] = comp_f32(v, i, c, true); }
        else {
            // geometry.rs:192-212: default (1,0,0,0); MikkTSpace when the material has a normal map
            for (size_t i = 0; i < nv; ++i) { verts[i].tangent[0] = 1.0f; verts[i].tangent[1] = verts[i].tangent[2] = verts[i].tangent[3] = 0.0f; }
            if (doc.materials[material_index].normal_texture.index >= 0) generate_tangents(verts, idx);
        }
        for (size_t i = 0; i < nv; ++i) for (int c = 0; c < 4; ++c) verts[i].color[c] = 1.0f;
        if (at.has("COLOR_0")) { View v = view(at.integer("COLOR_0", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) { for (int c = 0; c < v.ncomp && c < 4; ++c) verts[i].color[c] = comp_f32(v, i, c, true); if (v.ncomp == 3) verts[i].color[3] = 1.0f; } }
        if (at.has("WEIGHTS_0")) { View v = view(at.integer("WEIGHTS_0", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) for (int c = 0; c < 4; ++c) verts[i].weights[c] = comp_f32(v, i, c, true); }
        if (at.has("JOINTS_0")) { View v = view(at.integer("JOINTS_0", -1)); for (size_t i = 0; i < nv && i < v.count; ++i) for (int c = 0; c < 4; ++c) verts[i].joints[c] = comp_u32(v, i, c); }
        for (size_t i = 0; i < nv; ++i) verts[i].skin_index = -1;

        rt_prim_info pi{(uint32_t)doc.vertices.size(), (uint32_t)doc.indices.size(), (uint32_t)material_index, 0};
        rt_geometry g{(uint32_t)nv, (uint32_t)idx.size(), doc.materials[material_index].alpha_mode == 1 ? 1u : 0u, 0};
        if (idx.size() % 3) die("index count not a multiple of 3");
        for (uint32_t k : idx) if (k >= nv) die("index out of vertex range");
        doc.vertices.insert(doc.vertices.end(), verts.begin(), verts.end());
        doc.indices.insert(doc.indices.end(), idx.begin(), idx.end());
        doc.prim_infos.push_back(pi); doc.geometries.push_back(g);
        Primitive p; p.geo_id = geo_id;
        // primitive.bounding_box(): POSITION accessor min/max
        const Json& pa = (*root.get("accessors"))[at.integer("POSITION", 0)];
        float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}; arr_to(pa.get("min"), lo, 3); arr_to(pa.get("max"), hi, 3);
        memcpy(p.aabb.lo, lo, 12); memcpy(p.aabb.hi, hi, 12);
        return p;
    }

    // geometry.rs:192-212: mikktspace::generate_tangents over the indexed triangles (restated in mikktspace_gen.h)
    static void generate_tangents(std::vector<rt_vertex>& v, const std::vector<uint32_t>& idx) {
        if (v.empty() || idx.size() < 3) return;
        const size_t stride = sizeof(rt_vertex) / sizeof(float);
        mikk::Mesh m{v[0].position, v[0].normal, v[0].uv0, stride};
        mikk::generate(m, v.size(), idx.data(), idx.size(), v[0].tangent, stride);
    }

    void load_textures() {
        // linear set: material.rs:102-128 (normal, metallicRoughness, transmission, specular textures' *image* index)
        std::vector<bool> linear;
        const Json* imgs = root.get("images"); const Json* texs = root.get("textures"); const Json* mats = root.get("materials");
        size_t nimg = imgs ? imgs->size() : 0; linear.assign(nimg, false);
        auto mark = [&](const Json* ti) {
            if (!ti || !ti->has("index") || !texs) return; size_t t = (size_t)ti->integer("index", 0); if (t >= texs->size()) return;
            int src = (*texs)[t].integer("source", -1); if (src >= 0 && (size_t)src < nimg) linear[src] = true;
        };
        for (size_t i = 0; mats && i < mats->size(); ++i) {
            const Json& m = (*mats)[i]; mark(m.get("normalTexture"));
            if (const Json* p = m.get("pbrMetallicRoughness")) mark(p->get("metallicRoughnessTexture"));
            if (const Json* e = m.get("extensions")) {
                if (const Json* t = e->get("KHR_materials_transmission")) mark(t->get("transmissionTexture"));
                if (const Json* s = e->get("KHR_materials_specular")) mark(s->get("specularTexture"));
            }
        }
        // image 0 = 1x1 dummy (image.rs:31-43)
        DecodedImage dummy; dummy.w = dummy.h = 1; dummy.rgba = {1, 1, 1, 1};
        doc.images.push_back(dummy); doc.image_srgb.push_back(1);
        for (size_t i = 0; i < nimg; ++i) {
            const Json& im = (*imgs)[i]; std::vector<uint8_t> bytes;
            if (im.has("uri")) bytes = load_uri(im.string("uri"));
            else {
                int bv = im.integer("bufferView", -1); if (bv < 0) die("image without uri/bufferView");
                const Json& b = (*root.get("bufferViews"))[bv]; const auto& buf = buffers.at(b.integer("buffer", 0));
                size_t off = (size_t)b.number("byteOffset", 0), len = (size_t)b.number("byteLength", 0);
                if (off + len > buf.size()) die("image bufferView out of range");
                bytes.assign(buf.begin() + off, buf.begin() + off + len);
            }
            DecodedImage di; std::string err;
            if (!decode_image(bytes.data(), bytes.size(), di, err)) die("image " + std::to_string(i) + ": " + err + " (the C++ host decodes 8-bit PNG and baseline/progressive JPEG; pass other formats pre-decoded)");
            doc.images.push_back(std::move(di)); doc.image_srgb.push_back(linear[i] ? 0 : 1);
        }
        // samplers: slot 0 default (texture.rs:31-41)
        doc.samplers.push_back(rt_sampler_desc{RT_FILTER_LINEAR, RT_FILTER_LINEAR, RT_WRAP_REPEAT, RT_WRAP_REPEAT});
        const Json* smp = root.get("samplers");
        auto wrap = [](int w) { return w == 33071 ? (uint32_t)RT_WRAP_CLAMP : w == 33648 ? (uint32_t)RT_WRAP_MIRROR : (uint32_t)RT_WRAP_REPEAT; };
        for (size_t i = 0; smp && i < smp->size(); ++i) {
            const Json& s = (*smp)[i]; rt_sampler_desc d;
            int mag = s.integer("magFilter", 9729), mn = s.integer("minFilter", 9729);
            d.mag_filter = mag == 9728 ? RT_FILTER_NEAREST : RT_FILTER_LINEAR;
            d.min_filter = (mn == 9728 || mn == 9984 || mn == 9986) ? RT_FILTER_NEAREST : RT_FILTER_LINEAR;   // globals.rs:368-375
            d.wrap_s = wrap(s.integer("wrapS", 10497)); d.wrap_t = wrap(s.integer("wrapT", 10497));
            doc.samplers.push_back(d);
        }
        // textures: slot 0 default (texture.rs:21-29, 66-78)
        doc.textures.push_back(rt_texture_desc{0, 0});
        for (size_t i = 0; texs && i < texs->size(); ++i) {
            const Json& t = (*texs)[i]; rt_texture_desc d;
            d.image_index = (uint32_t)(t.integer("source", 0) + 1); d.sampler_index = t.has("sampler") ? (uint32_t)(t.integer("sampler", 0) + 1) : 0u;
            if (d.image_index >= doc.images.size() || d.sampler_index >= doc.samplers.size()) die("texture refers to missing image/sampler");
            doc.textures.push_back(d);
        }
    }

    void load_animations() {   // animation.rs:49-75, 212-226
        const Json* anims = root.get("animations");
        for (size_t a = 0; anims && a < anims->size(); ++a) {
            const Json& an = (*anims)[a]; const Json* chs = an.get("channels"); const Json* sms = an.get("samplers");
            for (size_t c = 0; chs && c < chs->size(); ++c) {
                const Json& ch = (*chs)[c]; const Json* tg = ch.get("target"); if (!tg || !tg->has("node")) continue;
                if (!sms) die("animation channel without samplers");
                const int si = ch.integer("sampler", 0); if (si < 0) die("animation channel refers to a missing sampler");
                const Json& sm = (*sms)[(size_t)si];
                Channel out; out.target = tg->integer("node", 0);
                std::string path = tg->string("path"); out.prop = path == "translation" ? 0 : path == "rotation" ? 1 : path == "scale" ? 2 : 3;
                std::string ip = sm.string("interpolation", "LINEAR"); out.interp = ip == "STEP" ? 1 : ip == "CUBICSPLINE" ? 2 : 0;
                View in = view(sm.integer("input", -1)); View ov = view(sm.integer("output", -1));
                out.input.resize(in.count); for (size_t i = 0; i < in.count; ++i) out.input[i] = comp_f32(in, i, 0, false);
                out.comps = ov.ncomp; out.out.resize(ov.count * ov.ncomp);
                for (size_t i = 0; i < ov.count; ++i) for (int k = 0; k < ov.ncomp; ++k) out.out[i * ov.ncomp + k] = comp_f32(ov, i, k, true);
                if (out.prop != 3) {   // the sampler below reads input.size() keys (three per key for CUBICSPLINE) of 3 or 4 floats
                    const size_t keys = out.interp == 2 ? out.input.size() * 3 : out.input.size();
                    if (out.comps < (out.prop == 1 ? 4 : 3) || out.out.size() < keys * (size_t)out.comps) die("animation sampler output does not match its input");
                }
                doc.channels.push_back(std::move(out));
            }
        }
    }

    void traverse(const std::function<void(const Node&)>& f) {   // scene_graph.rs:83-103
        std::function<void(int)> rec = [&](int n) { const Node& nd = doc.nodes[n]; f(nd); for (int c : nd.children) rec(c); };
        for (int r : doc.scenes[doc.current_scene]) rec(r);
    }

    void tag_skinned_vertices() {   // duplicate_mesh_for_non_affine_transform scene_graph.rs:191-235
        std::vector<int> mesh_skin(doc.meshes.size(), -1);
        traverse([&](const Node& n) {
            if (n.mesh < 0 || n.skin < 0) return;
            if (mesh_skin[n.mesh] >= 0) return;
            mesh_skin[n.mesh] = n.skin;
            for (const Primitive& p : doc.meshes[n.mesh].prims) {
                const rt_prim_info& pi = doc.prim_infos[p.geo_id]; const rt_geometry& g = doc.geometries[p.geo_id];
                for (uint32_t v = 0; v < g.v_len; ++v) doc.vertices[pi.v_offset + v].skin_index = n.skin;
            }
        });
    }

    // aabb.rs / scene_graph.rs:339-353.  Only the min and max corners are transformed (aabb.rs:73-85).
    static Aabb xf(const Aabb& a, const Mat4& m) {
        Aabb r; float lo[4] = {a.lo[0], a.lo[1], a.lo[2], 1}, hi[4] = {a.hi[0], a.hi[1], a.hi[2], 1}, o[4];
        mul_vec4(m, lo, o); memcpy(r.lo, o, 12); mul_vec4(m, hi, o); memcpy(r.hi, o, 12); return r;
    }
    static bool uni(const std::vector<Aabb>& v, Aabb& out) {
        if (v.empty()) return false;
        out = v[0];
        if (v.size() == 1) return true;
        for (const Aabb& a : v) for (int c = 0; c < 3; ++c) { out.lo[c] = std::min(out.lo[c], a.lo[c]); out.hi[c] = std::max(out.hi[c], a.hi[c]); }
        return true;
    }
    bool node_aabb(int n, Aabb& out) {
        const Node& cur = doc.nodes[n]; std::vector<Aabb> ch;
        for (int c : cur.children) { Aabb a; if (node_aabb(c, a)) ch.push_back(a); }
        if (cur.mesh >= 0) { std::vector<Aabb> ps; for (auto& p : doc.meshes[cur.mesh].prims) ps.push_back(p.aabb); Aabb m; if (uni(ps, m)) ch.push_back(m); }
        Mat4 l = cur.local_matrix();
        for (Aabb& a : ch) a = xf(a, l);
        return uni(ch, out);
    }
    void set_parent(int n, const Mat4& parent) {   // update_parent_transform scene_graph.rs:298-306
        Node& nd = doc.nodes[n]; nd.parent_cache = parent;
        Mat4 next = mul(parent, nd.local_matrix());
        for (int c : nd.children) set_parent(c, next);
    }
    void load_scene() {   // scene_graph.rs:277-289
        std::vector<Aabb> as;
        for (int r : doc.scenes[doc.current_scene]) { Aabb a; if (node_aabb(r, a)) as.push_back(a); }
        Aabb all; if (!uni(as, all)) die("scene has no geometry (reference unwraps None)");
        float size[3] = {std::fabs(all.hi[0] - all.lo[0]), std::fabs(all.hi[1] - all.lo[1]), std::fabs(all.hi[2] - all.lo[2])};
        float larger = (size[0] > size[1] && size[0] > size[2]) ? size[0] : (size[1] > size[2] ? size[1] : size[2]);
        float center[3]; for (int c = 0; c < 3; ++c) center[c] = all.lo[c] + (all.hi[c] - all.lo[c]) / 2.0f;
        float sc = 10.0f / larger;
        Mat4 T = identity(); T.m[12] = -center[0]; T.m[13] = -center[1]; T.m[14] = -center[2];
        Mat4 S = identity(); S.m[0] = S.m[5] = S.m[10] = sc;
        doc.aabb_trans = mul(S, T);
        for (int r : doc.scenes[doc.current_scene]) set_parent(r, doc.aabb_trans);
    }
};

// deterministic stand-in for rand::thread_rng() in LightRaw::random_light (light.rs:84-103): intensity is 0, so
// the positions never influence an image until the host overrides the lights.
struct Pcg32 { uint64_t s = 0x853c49e6748fea9bULL; uint32_t next() { uint64_t o = s; s = o * 6364136223846793005ULL + 1442695040888963407ULL; uint32_t x = (uint32_t)(((o >> 18u) ^ o) >> 27u), r = (uint32_t)(o >> 59u); return (x >> r) | (x << ((32 - r) & 31)); } float f() { return (next() >> 8) * (1.0f / 16777216.0f); } };

void refresh_derived(gv_doc& d) {
    Loader helper(d);
    // instances: create_top_as acceleration_structures.rs:143-172
    d.instances.clear();
    helper.traverse([&](const Node& n) {
        if (n.mesh < 0) return;
        Mat4 w = n.skin < 0 ? n.world() : identity();
        rt_instance in; memset(&in, 0, sizeof in);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) in.transform[r * 4 + c] = w.m[c * 4 + r];   // transpose -> row-major 3x4
        in.mask = 0xFF; in.flags = 0x1;   // TRIANGLE_FACING_CULL_DISABLE
        for (const Primitive& p : d.meshes[n.mesh].prims) { in.geo_id = p.geo_id; d.instances.push_back(in); }
    });
    // lights: get_lights_raw scene_graph.rs:55-81
    d.dlights.clear(); d.plights.clear();
    helper.traverse([&](const Node& n) {
        if (n.light < 0 || (size_t)n.light >= d.lights.size()) return;
        const LightDef& L = d.lights[n.light]; rt_light r; memset(&r, 0, sizeof r);
        r.color[0] = L.color[0]; r.color[1] = L.color[1]; r.color[2] = L.color[2]; r.color[3] = 0;
        Mat4 w = n.world(); float v[4] = {0, 0, 0, 1}; if (L.kind == 0) { v[2] = -1.0f; v[3] = 0; }
        mul_vec4(w, v, r.transform); r.kind = (uint32_t)L.kind; r.range = L.range; r.intensity = L.intensity;
        (L.kind == 0 ? d.dlights : d.plights).push_back(r);
    });
    if (d.plights.empty()) {
        Pcg32 rng;
        for (int i = 0; i < 5; ++i) {
            rt_light r; memset(&r, 0, sizeof r); for (int c = 0; c < 4; ++c) { r.color[c] = 1.0f; r.transform[c] = (rng.f() - 0.5f) * 2.0f * 10.0f; }
            r.kind = 1; r.range = INFINITY; r.intensity = 0.0f; d.plights.push_back(r);
        }
    }
    if (d.dlights.empty()) {   // LightRaw::default light.rs:113-124
        rt_light r; memset(&r, 0, sizeof r); for (int c = 0; c < 4; ++c) { r.color[c] = 1.0f; r.transform[c] = 1.0f; }
        r.kind = 0; r.range = INFINITY; r.intensity = 0.0f; d.dlights.push_back(r);
    }
    // skins: get_skins scene_graph.rs:329-337, skinning.rs:39-50,78-81
    d.skin_mats.assign(d.skins.size() * 4096, 0.0f);
    for (size_t s = 0; s < d.skins.size(); ++s) {
        const Skin& sk = d.skins[s]; size_t len = std::min<size_t>(sk.joints.size(), RT_MAX_JOINTS);
        for (size_t j = 0; j < len; ++j) { Mat4 m = mul(d.nodes[sk.joints[j]].world(), sk.ibm[j]); memcpy(&d.skin_mats[s * 4096 + j * 16], m.m, 64); }
    }
}

// glam Quat::slerp (0.24): shortest path, lerp fallback when dot > 0.9995
void quat_slerp(const float* a, const float* b_in, float s, float* out) {
    float b[4] = {b_in[0], b_in[1], b_in[2], b_in[3]};
    float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    if (dot < 0.0f) { for (float& x : b) x = -x; dot = -dot; }
    if (dot > 0.9995f) {
        float r[4]; for (int i = 0; i < 4; ++i) r[i] = a[i] + (b[i] - a[i]) * s;
        float l = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]); for (int i = 0; i < 4; ++i) out[i] = r[i] / l;
    } else {
        float theta = std::acos(dot), s1 = std::sin(theta * (1.0f - s)), s2 = std::sin(theta * s), st = 1.0f / std::sin(theta);
        for (int i = 0; i < 4; ++i) out[i] = (a[i] * s1 + b[i] * s2) * st;
    }
}

// Transform::decomposed for a matrix-form node (gltf crate): T = col3, S = column lengths, R from normalised basis
void decompose(const Mat4& m, float* t, float* r, float* s) {
    t[0] = m.m[12]; t[1] = m.m[13]; t[2] = m.m[14];
    float c[3][3];
    for (int k = 0; k < 3; ++k) { s[k] = std::sqrt(m.m[k * 4] * m.m[k * 4] + m.m[k * 4 + 1] * m.m[k * 4 + 1] + m.m[k * 4 + 2] * m.m[k * 4 + 2]); for (int i = 0; i < 3; ++i) c[k][i] = m.m[k * 4 + i] / s[k]; }
    float m00 = c[0][0], m11 = c[1][1], m22 = c[2][2], tr = m00 + m11 + m22;
    if (tr > 0) { float S = std::sqrt(tr + 1.0f) * 2; r[3] = 0.25f * S; r[0] = (c[1][2] - c[2][1]) / S; r[1] = (c[2][0] - c[0][2]) / S; r[2] = (c[0][1] - c[1][0]) / S; }
    else if (m00 > m11 && m00 > m22) { float S = std::sqrt(1.0f + m00 - m11 - m22) * 2; r[3] = (c[1][2] - c[2][1]) / S; r[0] = 0.25f * S; r[1] = (c[1][0] + c[0][1]) / S; r[2] = (c[2][0] + c[0][2]) / S; }
    else if (m11 > m22) { float S = std::sqrt(1.0f + m11 - m00 - m22) * 2; r[3] = (c[2][0] - c[0][2]) / S; r[0] = (c[1][0] + c[0][1]) / S; r[1] = 0.25f * S; r[2] = (c[2][1] + c[1][2]) / S; }
    else { float S = std::sqrt(1.0f + m22 - m00 - m11) * 2; r[3] = (c[0][1] - c[1][0]) / S; r[0] = (c[2][0] + c[0][2]) / S; r[1] = (c[2][1] + c[1][2]) / S; r[2] = 0.25f * S; }
}

}  // namespace

extern "C" {

const char* gv_last_error(void) { return g_err.c_str(); }

int gv_load_file(const char* path, gv_doc** out) {
    if (!path || !out) return fail("null argument");
    std::unique_ptr<gv_doc> d(new gv_doc());
    try { Loader l(*d); l.load(path); refresh_derived(*d); }
    catch (const Error& e) { return fail(std::string("gv_load_file: ") + e.msg); }
    catch (const std::exception& e) { return fail(std::string("gv_load_file: ") + e.what()); }
    *out = d.release(); return 0;
}
void gv_doc_free(gv_doc* d) { delete d; }

int gv_doc_scene_desc(gv_doc* d, rt_scene_desc* o) {
    if (!d || !o) return fail("null argument");
    memset(o, 0, sizeof *o);
    o->vertices = d->vertices.data(); o->n_vertices = (uint32_t)d->vertices.size();
    o->indices = d->indices.data(); o->n_indices = (uint32_t)d->indices.size();
    o->prim_infos = d->prim_infos.data(); o->geometries = d->geometries.data(); o->n_geometries = (uint32_t)d->geometries.size();
    o->materials = d->materials.data(); o->n_materials = (uint32_t)d->materials.size();
    o->instances = d->instances.data(); o->n_instances = (uint32_t)d->instances.size();
    d->image_descs.clear();
    for (size_t i = 0; i < d->images.size(); ++i) d->image_descs.push_back(rt_image_desc{d->images[i].rgba.data(), d->images[i].w, d->images[i].h, d->image_srgb[i], 0});
    o->images = d->image_descs.data(); o->n_images = (uint32_t)d->image_descs.size();
    o->samplers = d->samplers.data(); o->n_samplers = (uint32_t)d->samplers.size();
    o->textures = d->textures.data(); o->n_textures = (uint32_t)d->textures.size();
    o->dlights = d->dlights.data(); o->n_dlights = (uint32_t)d->dlights.size();
    o->plights = d->plights.data(); o->n_plights = (uint32_t)d->plights.size();
    o->skins = d->skin_mats.empty() ? nullptr : d->skin_mats.data(); o->n_skins = (uint32_t)d->skins.size();
    if (d->sky_w) { for (int f = 0; f < 6; ++f) o->skybox_faces[f] = d->sky[f].data(); o->skybox_width = d->sky_w; o->skybox_height = d->sky_h; o->skybox_srgb = d->sky_srgb; }
    return 0;
}
int gv_doc_fully_opaque(const gv_doc* d) { for (auto& m : d->materials) if (m.alpha_mode != 1) return 0; return 1; }
int gv_doc_static_scene(const gv_doc* d) { return d->channels.empty(); }
int gv_doc_need_compute(const gv_doc* d) { return !d->skins.empty(); }
void gv_doc_aabb_trans(const gv_doc* d, float o[16]) { memcpy(o, d->aabb_trans.m, 64); }

int gv_doc_animate(gv_doc* d, float t_in) {   // scene_graph.rs:308-323, animation.rs:77-146, Node::animate :397-416
    try {
        Loader helper(*d);
        for (const Channel& c : d->channels) {
            if (c.input.empty() || c.prop == 3) continue;   // morph: ignored by the reference (animation.rs:141-144)
            size_t len = c.input.size(); float mn = c.input[0], mx = c.input[len - 1], interval = mx - mn;
            float t = t_in > mn ? std::fmod(t_in - mn, interval) + mn : t_in;
            size_t s = 0, e = 0;
            for (size_t i = 0; i + 1 < len; ++i) if (t >= c.input[i] && t <= c.input[i + 1]) { s = i; e = s + 1; }
            float prev = c.input[s], next = c.input[e], factor = (t - prev) / (next - prev);
            float res[4] = {0, 0, 0, 0}; int n = c.comps;
            auto val = [&](size_t k) { return &c.out[k * n]; };
            if (c.prop == 1) {
                if (c.interp == 1) memcpy(res, val(s), 16); else quat_slerp(val(s), val(e), factor, res);
            } else {
                if (c.interp == 0) { for (int k = 0; k < 3; ++k) res[k] = val(s)[k] + (val(e)[k] - val(s)[k]) * factor; }
                else if (c.interp == 1) memcpy(res, val(s), 12);
                else {   // cubic_spline animation.rs:156-176
                    size_t s3 = s * 3; const float *p0 = val(s3 + 1), *m0v = val(s3 + 2), *p1 = val(s3 + 4), *m1v = val(s3 + 3);
                    float dt = next - prev, tt = factor;
                    for (int k = 0; k < 3; ++k) {
                        float m0 = dt * m0v[k], m1 = dt * m1v[k];
                        res[k] = (2.0f * tt * tt * tt - 3.0f * tt * tt + 1.0f) * p0[k] + (tt * tt * tt - 2.0f * tt * tt + tt) * m0 + (-2.0f * tt * tt * tt + 3.0f * tt * tt) * p1[k] + (tt * tt * tt - tt * tt) * m1;
                    }
                }
            }
            Node& nd = d->nodes.at(c.target);
            if (!nd.local.decomposed) { decompose(nd.local.matrix, nd.local.t, nd.local.r, nd.local.s); nd.local.decomposed = true; }
            if (c.prop == 0) memcpy(nd.local.t, res, 12); else if (c.prop == 1) memcpy(nd.local.r, res, 16); else memcpy(nd.local.s, res, 12);
            helper.set_parent(c.target, nd.parent_cache);   // update_local_transform
        }
        refresh_derived(*d);
    } catch (const Error& e) { return fail(std::string("gv_doc_animate: ") + e.msg); }
    catch (const std::exception& e) { return fail(std::string("gv_doc_animate: ") + e.what()); }
    return 0;
}
int gv_doc_get_skins(gv_doc* d, const float** mats, uint32_t* n) { *mats = d->skin_mats.data(); *n = (uint32_t)d->skins.size(); return 0; }
int gv_doc_get_instances(gv_doc* d, const rt_instance** inst, uint32_t* n) { *inst = d->instances.data(); *n = (uint32_t)d->instances.size(); return 0; }
int gv_doc_set_skybox(gv_doc* d, const uint8_t* const faces[6], uint32_t w, uint32_t h, uint32_t srgb) {
    for (int f = 0; f < 6; ++f) { if (!faces[f]) return fail("null skybox face"); d->sky[f].assign(faces[f], faces[f] + (size_t)w * h * 4); }
    d->sky_w = w; d->sky_h = h; d->sky_srgb = srgb; return 0;
}

void gv_camera_default(gv_camera* c, uint32_t w, uint32_t h) {
    c->position[0] = 0; c->position[1] = 0; c->position[2] = 1; c->direction[0] = 0; c->direction[1] = 0; c->direction[2] = -1;
    c->fov = 60.0f; c->aspect_ratio = (float)w / (float)h; c->z_near = 0.1f; c->z_far = 10.0f;
}
void gv_camera_view_matrix(const gv_camera* c, float o[16]) {
    // nalgebra Matrix4::look_at_rh(eye, target = eye + dir, up = +Y)
    float f[3] = {c->direction[0], c->direction[1], c->direction[2]};
    float fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]); for (float& x : f) x /= fl;
    float up[3] = {0, 1, 0};
    float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
    float sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]); for (float& x : s) x /= sl;
    float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    const float* e = c->position;
    o[0] = s[0]; o[4] = s[1]; o[8] = s[2];   o[12] = -(s[0] * e[0] + s[1] * e[1] + s[2] * e[2]);
    o[1] = u[0]; o[5] = u[1]; o[9] = u[2];   o[13] = -(u[0] * e[0] + u[1] * e[1] + u[2] * e[2]);
    o[2] = -f[0]; o[6] = -f[1]; o[10] = -f[2]; o[14] = (f[0] * e[0] + f[1] * e[1] + f[2] * e[2]);
    o[3] = 0; o[7] = 0; o[11] = 0; o[15] = 1;
}
void gv_camera_projection_matrix(const gv_camera* c, float o[16]) {
    // OPENGL_TO_VULKAN_RT * Matrix4::new_perspective(aspect, fovy, near, far)   camera.rs:106-118
    float fovy = c->fov * 3.14159265358979323846f / 180.0f, n = c->z_near, f = c->z_far;
    float t = std::tan(fovy / 2.0f);
    Mat4 P{}; P.m[0] = 1.0f / (c->aspect_ratio * t); P.m[5] = 1.0f / t; P.m[10] = (f + n) / (n - f); P.m[14] = 2.0f * f * n / (n - f); P.m[11] = -1.0f;
    Mat4 C = identity(); C.m[5] = -1.0f; C.m[10] = 0.5f; C.m[14] = 0.5f;   // rows (1,0,0,0),(0,-1,0,0),(0,0,.5,.5),(0,0,0,1)
    Mat4 r = mul(C, P); memcpy(o, r.m, 64);
}
int gv_mat4_inverse(const float in[16], float out[16]) { return inverse(in, out) ? 0 : fail("matrix not invertible"); }

void gv_gui_default(gv_gui* g) {   // gui_state.rs:303-332
    memset(g, 0, sizeof *g);
    g->aperture = 0.0f; g->focus_distance = 10.0f; g->number_of_samples = 3; g->number_of_bounces = 5; g->max_number_of_samples = 5000;
    g->acc = 1; g->sky = 0; g->antialiasing = 1; g->debug = 0; g->mapping = 0; g->animation = 0;
    g->map_scale = 1.0f; g->scale = 1.0f; g->orthographic_fov_dis = 0.0f; g->exposure = 5.0f; g->selected_tone_map_mode = 0;
}

void gv_build_ubo(const gv_camera* cam, const gv_gui* gui, uint32_t* total, uint32_t frame_count, uint32_t fully_opaque, uint32_t random_seed, rt_ubo* u) {
    memset(u, 0, sizeof *u);
    float view[16], proj[16];
    gv_camera_view_matrix(cam, view);
    float sc = gui->scale > 0.0f ? gui->scale : 1.0f / (std::fabs(gui->scale) + 1.0f);   // main.rs:197-202
    for (float& x : view) x *= sc;
    gv_camera_projection_matrix(cam, proj);
    memcpy(u->model_view, view, 64); memcpy(u->projection, proj, 64);
    inverse(view, u->model_view_inverse); inverse(proj, u->projection_inverse);
    // gui_state.rs:270-287 (dynamic sampling is a wall-clock controller and is not modelled)
    uint32_t n = gui->max_number_of_samples <= *total ? 0u : std::min(gui->max_number_of_samples - *total, gui->number_of_samples);
    bool acc = gui->acc && gui->mapping == 0 && !gui->animation;   // Gui::acc :289-291
    if (!acc) *total = 0;
    *total += n;
    u->aperture = gui->aperture; u->focus_distance = gui->focus_distance; u->fov_angle = 1.0f; u->orthographic_fov_dis = gui->orthographic_fov_dis;
    u->heatmap_scale = gui->map_scale; u->total_number_of_samples = *total; u->number_of_samples = n;
    u->number_of_bounces = (gui->mapping != 0 && gui->mapping != 1) ? 1u : gui->number_of_bounces;   // Gui::get_bounce :293-299
    u->random_seed = random_seed; u->has_sky = gui->sky; u->antialiasing = gui->antialiasing; u->mapping = gui->mapping;
    u->frame_count = frame_count; u->debug = gui->debug; u->fully_opaque = fully_opaque; u->exposure = gui->exposure;
    u->tone_mapping_mode = gui->selected_tone_map_mode;
}

int gv_decode_png(const uint8_t* data, size_t size, uint8_t** rgba, uint32_t* w, uint32_t* h) {
    DecodedImage d; std::string err;
    if (!decode_png(data, size, d, err)) return fail(err);
    *rgba = (uint8_t*)malloc(d.rgba.size()); memcpy(*rgba, d.rgba.data(), d.rgba.size()); *w = d.w; *h = d.h; return 0;
}
// SkyBox::new (cubumap.rs:86-106) + resource_manager::load_cubemap: the six *.png / *.jpg files of a directory, ordered
// by Face::get_index (cubumap.rs:29-50): posx,negx,posy,negy,posz,negz or right,left,top,bottom,front,back.
int gv_load_skybox_dir(const char* dir, uint8_t* faces[6], uint32_t* w, uint32_t* h) {
    static const char* names[6][2] = {{"posx", "right"}, {"negx", "left"}, {"posy", "top"}, {"negy", "bottom"}, {"posz", "front"}, {"negz", "back"}};
    for (int f = 0; f < 6; ++f) faces[f] = nullptr;
    DIR* dp = opendir(dir);
    if (!dp) return fail(std::string("cannot open skybox directory ") + dir);
    std::string found[6]; int count = 0;
    while (dirent* e = readdir(dp)) {
        const std::string fn = e->d_name;
        if (fn.size() < 5) continue;
        const std::string ext = fn.substr(fn.size() - 4);
        if (ext != ".png" && ext != ".jpg") continue;
        ++count;
        const std::string stem = fn.substr(0, fn.find('.'));
        for (int f = 0; f < 6; ++f) if (stem == names[f][0] || stem == names[f][1]) found[f] = std::string(dir) + "/" + fn;
    }
    closedir(dp);
    if (count != 6) return fail("skybox directory must hold exactly 6 .png/.jpg files (resource_manager::load_cubemap)");
    uint32_t fw = 0, fh = 0;
    for (int f = 0; f < 6; ++f) {
        if (found[f].empty()) { for (int g = 0; g < f; ++g) { free(faces[g]); faces[g] = nullptr; } return fail(std::string("skybox face missing: ") + names[f][0]); }
        std::vector<uint8_t> bytes;
        try { bytes = read_file(found[f]); } catch (const std::exception& ex) { for (int g = 0; g < f; ++g) { free(faces[g]); faces[g] = nullptr; } return fail(ex.what()); }
        DecodedImage d; std::string err;
        if (!decode_image(bytes.data(), bytes.size(), d, err) || (f && (d.w != fw || d.h != fh))) {
            for (int g = 0; g < f; ++g) { free(faces[g]); faces[g] = nullptr; }
            return fail(found[f] + ": " + (err.empty() ? "face size differs" : err));
        }
        fw = d.w; fh = d.h;
        faces[f] = (uint8_t*)malloc(d.rgba.size()); memcpy(faces[f], d.rgba.data(), d.rgba.size());
    }
    *w = fw; *h = fh;
    return 0;
}

void gv_generate_tangents(rt_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices) {
    if (!vertices || !indices || !n_vertices || n_indices < 3) return;
    const size_t stride = sizeof(rt_vertex) / sizeof(float);
    mikk::Mesh m{vertices[0].position, vertices[0].normal, vertices[0].uv0, stride};
    mikk::generate(m, n_vertices, indices, n_indices, vertices[0].tangent, stride);
}

int gv_decode_image(const uint8_t* data, size_t size, uint8_t** rgba, uint32_t* w, uint32_t* h) {
    DecodedImage d; std::string err;
    if (!decode_image(data, size, d, err)) return fail(err);
    *rgba = (uint8_t*)malloc(d.rgba.size()); memcpy(*rgba, d.rgba.data(), d.rgba.size()); *w = d.w; *h = d.h; return 0;
}
void gv_free(void* p) { free(p); }

}  // extern "C"
