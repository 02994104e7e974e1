// rt_platform.h — thin platform layer for the CUDA core.
//
// Product build: nvcc, sm_100a, everything below maps 1:1 onto CUDA (kernels, streams, cub radix sort,
// device atomics).  RT_EMU build (tests/emu only, never shipped, never loaded by the rustracer_b200
// package): the same sources are compiled by g++ with launches turned into loops so that the BVH builder,
// the traversal state machine and the shading code can be checked against the oracle on the CPU-only
// build box before GPU minutes are spent.  It is a development aid for the tests, not a fallback: the
// product library contains none of it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

#ifdef RT_EMU
// ----------------------------------------------------------------------------------------------------
// host emulation
// ----------------------------------------------------------------------------------------------------
#include <algorithm>
#include <cstdlib>
#include <vector>
using std::min; using std::max;
#define RT_HD inline
#define RT_D inline
#define RT_D_COLD inline
#define RT_LAMBDA [=]
#define RT_RESTRICT
typedef void* rt_stream_t;
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }

static inline uint32_t rt_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline unsigned long long rt_atomic_add64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline uint32_t rt_atomic_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline uint32_t rt_atomic_max(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline void rt_threadfence() {}
static inline float rt_ldg(const float* p) { return *p; }
template <class T> static inline T rt_ld(const T* p) { return *p; }

static inline float rt_fadd(float a, float b) { return a + b; }   // compiled with -ffp-contract=off
static inline float rt_fsub(float a, float b) { return a - b; }
static inline float rt_fmul(float a, float b) { return a * b; }
static inline float rt_fdiv(float a, float b) { return a / b; }
static inline double rt_dmul(double a, double b) { return a * b; }
static inline double rt_dsub(double a, double b) { return a - b; }
static inline float rt_rsqrt(float x) { return 1.0f / sqrtf(x); }
static inline float rt_rcp(float x) { return 1.0f / x; }
static inline float rt_rcp_approx(float x) { return 1.0f / x; }
static inline float rt_fma(float a, float b, float c) { return fmaf(a, b, c); }
static inline uint32_t rt_float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float rt_uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rt_byte_to_biased_float(uint32_t v, int i) { return rt_uint_as_float(0x4B000000u | ((v >> (8 * i)) & 0xFFu)); }
static inline uint32_t rt_byte_of(uint32_t v, int i) { return (v >> (8 * i)) & 0xFFu; }
static inline void rt_opaque(uint32_t&) {}
static inline uint32_t rt_shl_wrap(uint32_t v, uint32_t amount) { return v << (amount & 31u); }
static inline int rt_clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int rt_clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int rt_popc(uint32_t x) { return __builtin_popcount(x); }
static inline int rt_bfind(uint32_t x) { return x ? 31 - __builtin_clz(x) : -1; }   // index of highest set bit

inline int rt_malloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return *p ? 0 : 1; }
inline void rt_free(void* p) { free(p); }
inline int rt_h2d(void* d, const void* h, size_t n, rt_stream_t) { memcpy(d, h, n); return 0; }
inline int rt_d2h(void* h, const void* d, size_t n, rt_stream_t) { memcpy(h, d, n); return 0; }
inline int rt_d2d(void* d, const void* s, size_t n, rt_stream_t) { memcpy(d, s, n); return 0; }
inline int rt_memset(void* d, int v, size_t n, rt_stream_t) { memset(d, v, n); return 0; }
inline int rt_stream_sync(rt_stream_t) { return 0; }
inline const char* rt_platform_error() { return "emu"; }

template <class F> inline void rt_launch(size_t n, rt_stream_t, F f) { for (size_t i = 0; i < n; ++i) f(i); }

inline int rt_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, size_t n, void**, size_t*, rt_stream_t) {
    std::vector<uint32_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    for (size_t i = 0; i < n; ++i) { keys_tmp[i] = keys[idx[i]]; vals_tmp[i] = vals[idx[i]]; }
    memcpy(keys, keys_tmp, n * 8); memcpy(vals, vals_tmp, n * 4);
    return 0;
}

inline int rt_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void**, size_t*, rt_stream_t) {
    uint32_t acc = 0;
    for (size_t i = 0; i < n; ++i) { const uint32_t v = in[i]; out[i] = acc; acc += v; }
    return 0;
}

struct rt_timer { void create() {} void destroy() {} void record(rt_stream_t) {} bool ready() { return true; } };
inline float rt_timer_ms(rt_timer&, rt_timer&) { return 0.0f; }
struct rt_event { void create() {} void destroy() {} void record(rt_stream_t) {} void wait(rt_stream_t) {} int sync() { return 0; } };
inline int rt_stream_create(rt_stream_t* s) { *s = nullptr; return 0; }
inline void rt_stream_destroy(rt_stream_t) {}

#else
// ----------------------------------------------------------------------------------------------------
// CUDA (product)
// ----------------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <atomic>
#define RT_HD __host__ __device__ __forceinline__
#define RT_D __device__ __forceinline__
#define RT_D_COLD __device__ __noinline__   // rarely executed helpers kept out of line: smaller hot path, fewer registers
#define RT_LAMBDA [=] __device__
#define RT_RESTRICT __restrict__
typedef cudaStream_t rt_stream_t;

RT_D uint32_t rt_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
RT_D unsigned long long rt_atomic_add64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
RT_D uint32_t rt_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
RT_D uint32_t rt_atomic_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
RT_D void rt_threadfence() { __threadfence(); }
template <class T> RT_D T rt_ld(const T* p) { return __ldg(p); }

// round-to-nearest arithmetic that ptxas may not contract into FMA: the watertight test and the instance
// transform must produce the same bits as the oracle (g++ -ffp-contract=off)
RT_D float rt_fadd(float a, float b) { return __fadd_rn(a, b); }
RT_D float rt_fsub(float a, float b) { return __fsub_rn(a, b); }
RT_D float rt_fmul(float a, float b) { return __fmul_rn(a, b); }
RT_D float rt_fdiv(float a, float b) { return __fdiv_rn(a, b); }
RT_D double rt_dmul(double a, double b) { return __dmul_rn(a, b); }
RT_D double rt_dsub(double a, double b) { return __dsub_rn(a, b); }
RT_D float rt_rsqrt(float x) { return rsqrtf(x); }
RT_D float rt_rcp(float x) { return __frcp_rn(x); }
RT_D float rt_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// byte i of v -> float(2^23 + byte) with one PRMT (no I2F)
// (the constant is PRMT's first operand on purpose: the selector then encodes as an immediate.  With the constant as
//  the second operand ptxas rematerialised the selector into a register before every one of the 48 PRMTs of a node)
RT_D float rt_byte_to_biased_float(uint32_t v, int i) { return __uint_as_float(__byte_perm(0x4B000000u, v, 0x3004u + (uint32_t)i)); }
RT_D uint32_t rt_byte_of(uint32_t v, int i) { return __byte_perm(v, 0u, 0x4440u + (uint32_t)i); }          // one PRMT
RT_D void rt_opaque(uint32_t& v) { asm("" : "+r"(v)); }   // hides a value's provenance from the optimiser (no instruction)
RT_D uint32_t rt_shl_wrap(uint32_t v, uint32_t amount) { return __funnelshift_l(0u, v, amount); }            // SHF.L.W: low 5 bits of amount
RT_D float rt_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RT_D uint32_t rt_float_as_uint(float f) { return __float_as_uint(f); }
RT_D float rt_uint_as_float(uint32_t u) { return __uint_as_float(u); }
RT_D int rt_clz32(uint32_t x) { return __clz((int)x); }
RT_D int rt_clz64(uint64_t x) { return __clzll((long long)x); }
RT_D int rt_popc(uint32_t x) { return __popc(x); }
RT_D int rt_bfind(uint32_t x) { return 31 - __clz((int)x); }

const char* rt_platform_error();
int rt_malloc(void** p, size_t bytes);
void rt_free(void* p);
int rt_h2d(void* d, const void* h, size_t n, rt_stream_t s);
int rt_d2h(void* h, const void* d, size_t n, rt_stream_t s);
int rt_d2d(void* d, const void* s_, size_t n, rt_stream_t s);
int rt_memset(void* d, int v, size_t n, rt_stream_t s);
int rt_stream_sync(rt_stream_t s);
// tmp / tmp_bytes: caller-owned temporary storage (grown on demand; lives in the scene's BuildScratch, i.e. on that scene's device)
int rt_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, size_t n, void** tmp, size_t* tmp_bytes, rt_stream_t s);
int rt_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void** tmp, size_t* tmp_bytes, rt_stream_t s);

// generic element-wise launch: grid-stride, 256 threads, grid capped at a multiple of the SM count
template <class F> __global__ void __launch_bounds__(256) rt_foreach_kernel(size_t n, F f) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f(i);
}
extern int g_rt_sm_count;
extern std::atomic<unsigned long long> g_rt_launch_count;   // kernels launched by this library (all contexts of the process)
template <class F> inline void rt_launch(size_t n, rt_stream_t s, F f) {
    if (n == 0) return;
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)(g_rt_sm_count > 0 ? g_rt_sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    rt_foreach_kernel<<<(unsigned)blocks, 256, 0, s>>>(n, f);
    ++g_rt_launch_count;
}

struct rt_timer {
    cudaEvent_t e = nullptr;
    void create() { if (!e) cudaEventCreate(&e); }
    void destroy() { if (e) cudaEventDestroy(e); e = nullptr; }
    void record(rt_stream_t s) { cudaEventRecord(e, s); }
    bool ready() { return cudaEventQuery(e) == cudaSuccess; }
};
inline float rt_timer_ms(rt_timer& a, rt_timer& b) { float ms = 0; cudaEventElapsedTime(&ms, a.e, b.e); return ms; }
// ordering-only event (no timing): cross-stream dependencies of the frames in flight
struct rt_event {
    cudaEvent_t e = nullptr;
    void create() { if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
    void destroy() { if (e) cudaEventDestroy(e); e = nullptr; }
    void record(rt_stream_t s) { cudaEventRecord(e, s); }
    void wait(rt_stream_t s) { cudaStreamWaitEvent(s, e, 0); }
    int sync() { return cudaEventSynchronize(e) == cudaSuccess ? 0 : 1; }
};
inline int rt_stream_create(rt_stream_t* s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) == cudaSuccess ? 0 : 1; }
inline void rt_stream_destroy(rt_stream_t s) { if (s) cudaStreamDestroy(s); }
#endif
