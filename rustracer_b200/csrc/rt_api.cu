// rt_api.cu — product translation unit: CUDA platform functions, the C ABI (rt_core.h) and the multi-GPU
// peer-memory reduce.  Compiled for sm_100a only (see Makefile); no CPU path exists in this library.
#include "rt_core.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

int g_rt_sm_count = 148;
std::atomic<unsigned long long> g_rt_launch_count{0};
static thread_local cudaError_t g_last_cuda = cudaSuccess;

static int chk(cudaError_t e) { if (e != cudaSuccess) { g_last_cuda = e; return 1; } return 0; }
const char* rt_platform_error() { cudaError_t e = g_last_cuda != cudaSuccess ? g_last_cuda : cudaGetLastError(); g_last_cuda = cudaSuccess; return cudaGetErrorString(e); }
int rt_malloc(void** p, size_t bytes) { return chk(cudaMalloc(p, bytes ? bytes : 1)); }
void rt_free(void* p) { if (p) cudaFree(p); }
int rt_h2d(void* d, const void* h, size_t n, rt_stream_t s) { return chk(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
int rt_d2h(void* h, const void* d, size_t n, rt_stream_t s) { return chk(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
int rt_d2d(void* d, const void* s_, size_t n, rt_stream_t s) { return chk(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
int rt_memset(void* d, int v, size_t n, rt_stream_t s) { return chk(cudaMemsetAsync(d, v, n, s)); }
int rt_stream_sync(rt_stream_t s) { return chk(cudaStreamSynchronize(s)); }

// cub temporary storage is owned by the caller (BuildScratch of the scene being built): it lives on that scene's device and
// is only touched by work queued on that scene's stream.
static int grow_tmp(void** tmp, size_t* tmp_bytes, size_t need, rt_stream_t s) {
    if (need <= *tmp_bytes) return 0;
    if (*tmp) { cudaStreamSynchronize(s); cudaFree(*tmp); *tmp = nullptr; *tmp_bytes = 0; }
    const size_t cap = need + need / 4 + 256;
    if (chk(cudaMalloc(tmp, cap))) { *tmp = nullptr; return 1; }
    *tmp_bytes = cap;
    return 0;
}
int rt_sort_pairs_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, size_t n, void** tmp, size_t* tmp_bytes, rt_stream_t s) {
    size_t need = 0;
    if (chk(cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_tmp, vals, vals_tmp, (int)n, 0, 63, s))) return 1;
    if (grow_tmp(tmp, tmp_bytes, need, s)) return 1;
    size_t bytes = *tmp_bytes;
    if (chk(cub::DeviceRadixSort::SortPairs(*tmp, bytes, keys, keys_tmp, vals, vals_tmp, (int)n, 0, 63, s))) return 1;
    g_rt_launch_count += 4;
    if (rt_d2d(keys, keys_tmp, n * 8, s)) return 1;
    return rt_d2d(vals, vals_tmp, n * 4, s);
}

int rt_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void** tmp, size_t* tmp_bytes, rt_stream_t s) {
    size_t need = 0;
    if (chk(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, s))) return 1;
    if (grow_tmp(tmp, tmp_bytes, need, s)) return 1;
    size_t bytes = *tmp_bytes;
    if (chk(cub::DeviceScan::ExclusiveSum(*tmp, bytes, in, out, (int)n, s))) return 1;
    g_rt_launch_count += 1;
    return 0;
}

// ---- multi-GPU: peer access to the accumulation image across processes ----------------
#define RT_MAX_PEERS 8
struct PeerPtrs { const float4* p[RT_MAX_PEERS]; };

// Sums the peers' accumulation rows into this GPU's image over NVLink peer loads and tonemaps the result in
// the same pass (reduce fused with RayTracing.rgen:132-166).
__global__ void __launch_bounds__(256) reduce_peers_kernel(float4* acc, uint32_t* out, PeerPtrs peers, uint32_t n_peers, rt_ubo ubo, size_t begin, size_t end) {
    for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = acc[i];
        for (uint32_t k = 0; k < n_peers; ++k) {
            const float4 b = __ldcv(peers.p[k] + i);   // volatile load: peer memory, never cached stale
            a.x += b.x; a.y += b.y; a.z += b.z;
        }
        acc[i] = make_float4(a.x, a.y, a.z, 0.0f);
        rt_ubo u = ubo; u.number_of_samples = 0;
        accumulate_pixel(u, acc, out, i, mk3(0.0f), 0.0f, 0u);
    }
}

extern "C" {

#ifdef RT_PROBE
int rt_debug_probe(unsigned long long* out, uint32_t n_warps) {
    cudaDeviceSynchronize();
    return chk(cudaMemcpyFromSymbol(out, g_probe, (size_t)n_warps * 6 * sizeof(unsigned long long)));
}
#endif
int rt_ipc_export(rt_context* c, void* handle64) {
    if (!c || !handle64) return fail("rt_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    if (chk(cudaIpcGetMemHandle(&h, c->acc))) return fail(std::string("rt_ipc_export: ") + rt_platform_error());
    memcpy(handle64, &h, 64);
    c->ipc_exported = true;
    return 0;
}
int rt_ipc_open(rt_context* c, const void* handle64, void** peer_acc) {
    if (!c || !handle64 || !peer_acc) return fail("rt_ipc_open: null argument");
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h; memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (chk(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess))) return fail(std::string("rt_ipc_open: ") + rt_platform_error());
    c->ipc_opened.push_back(p); *peer_acc = p;
    return 0;
}
int rt_ipc_close(rt_context* c, void* peer_acc) {
    if (!c || !peer_acc) return fail("rt_ipc_close: null argument");
    for (size_t i = 0; i < c->ipc_opened.size(); ++i) if (c->ipc_opened[i] == peer_acc) { c->ipc_opened.erase(c->ipc_opened.begin() + i); break; }
    if (chk(cudaIpcCloseMemHandle(peer_acc))) return fail(std::string("rt_ipc_close: ") + rt_platform_error());
    return 0;
}
int rt_reduce_peers(rt_context* c, void* const* peer_acc, uint32_t n_peers, const rt_ubo* ubo, uint32_t row0, uint32_t row1, void* stream) {
    if (!c || !ubo || (n_peers && !peer_acc)) return fail("rt_reduce_peers: null argument");
    if (n_peers > RT_MAX_PEERS) return fail("rt_reduce_peers: too many peers");
    if (row1 > c->height || row0 > row1) return fail("rt_reduce_peers: bad row range");
    if (ubo->total_number_of_samples == 0) return fail("rt_reduce_peers: total_number_of_samples must be > 0");
    cudaSetDevice(c->device);
    PeerPtrs pp; for (uint32_t k = 0; k < RT_MAX_PEERS; ++k) pp.p[k] = k < n_peers ? (const float4*)peer_acc[k] : nullptr;
    const size_t begin = (size_t)row0 * c->width, end = (size_t)row1 * c->width;
    if (end == begin) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    size_t blocks = (end - begin + 255) / 256; const size_t cap = (size_t)g_rt_sm_count * 8; if (blocks > cap) blocks = cap;
    join_frames(c, st);          // frames still in flight on this context accumulate first
    reduce_peers_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->acc, c->slot[c->cur].fb.out, pp, n_peers, *ubo, begin, end);
    ++g_rt_launch_count;
    c->last_stream = st; consumer_ran(c, st);
    if (cudaPeekAtLastError() != cudaSuccess) return fail(std::string("rt_reduce_peers: ") + rt_platform_error());
    return 0;
}

}  // extern "C"
