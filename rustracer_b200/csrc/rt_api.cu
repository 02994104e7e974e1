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

// ---- multi-GPU combine with device-side synchronisation (rt_combine, include/rt_b200.h) ---------------------------
// One kernel per rank sums the peers' accumulation snapshots over NVLink peer loads for its band of the image, tonemaps
// it (RayTracing.rgen:132-166) and stores the RGBA8 band into its own AND the peers' display images (all-gather by peer
// stores) — reduce-scatter + tonemap + all-gather fused, no staging buffers.  Ordering between the GPUs is carried by
// three words in each rank's accumulation block (system-scope release / acquire), polled on the device:
//   SNAP: epoch of the snapshot a rank has published        READ: peer reads of my snapshot completed (count)
//   RECV: peer bands that have landed in my display (count)
// so that the host never waits for a GPU or for another process inside a combine.
enum { RT_SYNC_SNAP = 0, RT_SYNC_READ = 1, RT_SYNC_RECV = 2, RT_SYNC_ERR = 8 };
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void add_release_sys(uint32_t* p, uint32_t v) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

struct FlagPtrs { uint32_t* p[RT_MAX_PEERS]; };
#ifndef RT_COMBINE_TIMEOUT_NS
#define RT_COMBINE_TIMEOUT_NS 5000000000ull      // a peer that never arrives must not hang the GPU: give up, raise RT_SYNC_ERR
#endif
// one lane per flag: waits until *flag >= target (wrap-safe), single tiny block so that it never starves the kernels it waits for
__global__ void __launch_bounds__(32) wait_flags_kernel(FlagPtrs f, uint32_t n, uint32_t target, uint32_t* err) {
    if (threadIdx.x < n) {
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(f.p[threadIdx.x]) - target) < 0) {
            if (global_ns() - t0 > RT_COMBINE_TIMEOUT_NS) { st_release_sys(err, 1u); break; }
            __nanosleep(100);
        }
    }
}
__global__ void __launch_bounds__(32) publish_flag_kernel(uint32_t* flag, uint32_t value) {
    if (threadIdx.x == 0) { __threadfence_system(); st_release_sys(flag, value); }
}
// after the combine kernel: tell every peer that (a) its snapshot has been read, (b) my band is in its display
__global__ void __launch_bounds__(32) notify_peers_kernel(FlagPtrs read_flags, FlagPtrs recv_flags, uint32_t n) {
    if (threadIdx.x < n) {
        __threadfence_system();
        if (read_flags.p[threadIdx.x]) add_release_sys(read_flags.p[threadIdx.x], 1u);
        if (recv_flags.p[threadIdx.x]) add_release_sys(recv_flags.p[threadIdx.x], 1u);
    }
}

struct CombineArgs {
    const float4* src;                       // own snapshot (sample passes) or the accumulation image itself (tiles)
    float4* sum_out;                         // where the reduced RGBA32F band goes: own snapshot, or the root's (gather), may alias src
    uint32_t* display;                       // own display image
    const float4* peer_snap[RT_MAX_PEERS];   // summed when n_sum > 0
    uint32_t* peer_display[RT_MAX_PEERS];    // non-null entries receive the RGBA8 band (all-gather / gather to the root)
    uint32_t n_sum, n_peers;
    TilePart tp;                             // tiles: this rank's strips; else n_parts <= 1 and [begin, end) is a row band
    size_t begin, end;
    rt_ubo ubo;
};
__global__ void __launch_bounds__(256) combine_kernel(const CombineArgs a) {
    for (size_t i = a.begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.end; i += (size_t)gridDim.x * blockDim.x) {
        const size_t px = a.tp.n_parts > 1 ? (size_t)local_to_pixel(a.tp, (uint32_t)i) : i;
        float4 s = a.src[px];
        for (uint32_t k = 0; k < a.n_sum; ++k) {
            const float4 b = __ldcv(a.peer_snap[k] + px);           // peer memory over NVLink: never served from a stale L1 line
            s.x += b.x; s.y += b.y; s.z += b.z;
        }
        s.w = 0.0f;
        a.sum_out[px] = s;
        const uint32_t rgba = tonemap_rgba8(a.ubo, mk3(s.x, s.y, s.z));
        a.display[px] = rgba;
        for (uint32_t k = 0; k < a.n_peers; ++k) if (a.peer_display[k]) a.peer_display[k][px] = rgba;
    }
}

extern "C" {

int rt_combine(rt_context* c, const rt_combine_desc* d, const rt_ubo* ubo, void* stream) {
    if (!c || !d || !ubo || (d->n_peers && !d->peer_blocks)) return fail("rt_combine: null argument");
    if (d->n_peers > RT_MAX_PEERS) return fail("rt_combine: too many peers");
    if (ubo->total_number_of_samples == 0) return fail("rt_combine: total_number_of_samples must be > 0");
    if (d->epoch != c->combine_epoch + 1) return fail("rt_combine: epochs must be 1, 2, 3, ... (the same on every rank)");
    const bool tiles = d->n_parts > 1;
    if (tiles && (!d->strip_rows || d->part >= d->n_parts)) return fail("rt_combine: bad tile partition");
    if (!tiles && (d->row1 > c->height || d->row0 > d->row1)) return fail("rt_combine: bad row band");
    if (d->gather_to >= (int32_t)d->n_peers) return fail("rt_combine: gather_to is not a peer index");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t n = (size_t)c->width * c->height; const uint32_t e = d->epoch, np = d->n_peers;
    FlagPtrs own; for (auto& p : own.p) p = nullptr;
    uint32_t* err = c->sync + RT_SYNC_ERR;
    // 0. combines of one context are ordered, whichever streams they were given (the previous one still works on snap / display)
    if (c->combine_pending && c->combine_stream != st) c->ev_combine.wait(st);
    // 1. my previous snapshot must have been read by every peer before it is overwritten
    if (!tiles && np && e > 1) { own.p[0] = c->sync + RT_SYNC_READ; wait_flags_kernel<<<1, 32, 0, st>>>(own, 1, np * (e - 1), err); ++g_rt_launch_count; }
    // 2. frames in flight accumulate first; then the snapshot the peers will read (rendering may go on while they do)
    join_frames(c, st);
    CombineArgs a; memset(&a, 0, sizeof a);
    if (!tiles) {
        if (chk(cudaMemcpyAsync(c->snap, c->acc, n * sizeof(float4), cudaMemcpyDeviceToDevice, st))) return fail(std::string("rt_combine: ") + rt_platform_error());
        consumer_ran(c, st);      // the snapshot is the combine's only access to acc: later frames wait for the copy, not for the peers
        publish_flag_kernel<<<1, 32, 0, st>>>(c->sync + RT_SYNC_SNAP, e); ++g_rt_launch_count;
        a.src = c->snap; a.n_sum = np;
        a.begin = (size_t)d->row0 * c->width; a.end = (size_t)d->row1 * c->width;
        a.tp.n_parts = 1; a.tp.strip_rows = 1; a.tp.part = 0;
    } else {
        a.src = c->acc; a.n_sum = 0;
        a.tp.strip_rows = d->strip_rows; a.tp.n_parts = d->n_parts; a.tp.part = d->part;
        a.begin = 0; a.end = (size_t)owned_rows(TilePart{d->strip_rows, d->n_parts, d->part, c->width, c->height}) * c->width;
    }
    a.tp.width = c->width; a.tp.height = c->height;
    a.display = c->display; a.n_peers = np; a.ubo = *ubo;
    a.sum_out = c->snap;
    FlagPtrs snap_flags, read_flags, recv_flags;
    for (uint32_t k = 0; k < RT_MAX_PEERS; ++k) {
        snap_flags.p[k] = read_flags.p[k] = recv_flags.p[k] = nullptr;
        if (k >= np) continue;
        char* base = (char*)d->peer_blocks[k];
        if (!base) return fail("rt_combine: null peer block");
        a.peer_snap[k] = (const float4*)(base + acc_block_snap(n));
        uint32_t* psync = (uint32_t*)(base + acc_block_sync(n));
        snap_flags.p[k] = psync + RT_SYNC_SNAP;
        if (!tiles) read_flags.p[k] = psync + RT_SYNC_READ;
        const bool gets_band = d->gather_to == RT_GATHER_ALL || d->gather_to == (int32_t)k;
        if (gets_band) { a.peer_display[k] = (uint32_t*)(base + acc_block_display(n)); recv_flags.p[k] = psync + RT_SYNC_RECV; }
        if (d->gather_acc && d->gather_to == (int32_t)k) a.sum_out = (float4*)(base + acc_block_snap(n));   // full RGBA32F sum assembled on the root
    }
    // 3. every peer's snapshot of this epoch must be published before it is read
    if (!tiles && np) { wait_flags_kernel<<<1, 32, 0, st>>>(snap_flags, np, e, err); ++g_rt_launch_count; }
    if (a.end > a.begin) {
        size_t blocks = (a.end - a.begin + 255) / 256; const size_t cap = (size_t)g_rt_sm_count * 8; if (blocks > cap) blocks = cap;
        combine_kernel<<<(unsigned)blocks, 256, 0, st>>>(a); ++g_rt_launch_count;
    }
    // 4. peers may overwrite their snapshots / read their displays once my reads / stores are done
    if (np) { notify_peers_kernel<<<1, 32, 0, st>>>(read_flags, recv_flags, np); ++g_rt_launch_count; }
    // 5. my display is complete when every rank that sends me its band has done so
    uint32_t senders = 0;
    if (d->n_senders) senders = d->n_senders;
    if (senders) { own.p[0] = c->sync + RT_SYNC_RECV; wait_flags_kernel<<<1, 32, 0, st>>>(own, 1, c->recv_expected + senders, err); ++g_rt_launch_count; c->recv_expected += senders; }
    c->combine_epoch = e;
    if (tiles) consumer_ran(c, st);         // (tiles: the kernel read acc itself)
    c->combine_stream = st; c->ev_combine.record(st); c->combine_pending = true;
    if (cudaPeekAtLastError() != cudaSuccess) return fail(std::string("rt_combine: ") + rt_platform_error());
    return 0;
}

int rt_readback_display(rt_context* c, uint8_t* out_rgba8, float* sum_rgba32f) {
    if (!c) return fail("rt_readback_display: null context");
    cudaSetDevice(c->device);
    if (sync_all(c)) return fail(std::string("rt_readback_display: ") + rt_platform_error());
    const size_t n = (size_t)c->width * c->height;
    uint32_t err = 0;
    if (chk(cudaMemcpy(&err, c->sync + RT_SYNC_ERR, 4, cudaMemcpyDeviceToHost))) return fail(std::string("rt_readback_display: ") + rt_platform_error());
    if (err) { cudaMemset(c->sync + RT_SYNC_ERR, 0, 4); return fail("rt_readback_display: a combine timed out waiting for a peer GPU (RT_SYNC_ERR)"); }
    if (out_rgba8 && chk(cudaMemcpy(out_rgba8, c->display, n * 4, cudaMemcpyDeviceToHost))) return fail(std::string("rt_readback_display: ") + rt_platform_error());
    if (sum_rgba32f && chk(cudaMemcpy(sum_rgba32f, c->snap, n * 16, cudaMemcpyDeviceToHost))) return fail(std::string("rt_readback_display: ") + rt_platform_error());
    return 0;
}

int rt_combine_ptrs(rt_context* c, void** block, void** display, uint64_t* block_bytes) {
    if (!c) return fail("rt_combine_ptrs: null context");
    if (block) *block = c->acc;
    if (display) *display = c->display;
    if (block_bytes) *block_bytes = acc_block_bytes((size_t)c->width * c->height);
    return 0;
}

// ---- one process, several GPUs -------------------------------------------------------------------------------------
}  // extern "C"
struct rt_multi {
    std::vector<rt_context*> ctx; std::vector<rt_scene*> scene; std::vector<int> device;
    uint32_t mode = RT_PARTITION_TILES, width = 0, height = 0;
    uint64_t frames = 0; bool dirty = false;     // frames submitted / frames submitted since the last combine
    static const uint32_t strip_rows = 8;
};
extern "C" {

void rt_multi_destroy(rt_multi* m) {
    if (!m) return;
    for (rt_context* c : m->ctx) if (c) rt_synchronize(c);
    for (size_t i = 0; i < m->scene.size(); ++i) if (m->scene[i]) rt_scene_destroy(m->scene[i]);
    for (rt_context* c : m->ctx) if (c) rt_context_destroy(c);
    delete m;
}

int rt_multi_create(const int* devices, uint32_t n, uint32_t width, uint32_t height, uint32_t mode, rt_multi** out) {
    if (!devices || !out || n == 0 || n > RT_MAX_PEERS + 1) return fail("rt_multi_create: 1..9 devices");
    if (mode != RT_PARTITION_TILES && mode != RT_PARTITION_SAMPLE_PASSES) return fail("rt_multi_create: unknown partition mode");
    rt_multi* m = new rt_multi(); m->mode = mode; m->width = width; m->height = height;
    for (uint32_t i = 0; i < n; ++i) {
        rt_context* c = nullptr;
        if (rt_context_create(devices[i], width, height, &c)) { rt_multi_destroy(m); return 1; }
        m->ctx.push_back(c); m->device.push_back(devices[i]);
    }
    // peer access in every direction between distinct devices (replicas on one device need none)
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            if (devices[i] == devices[j]) continue;
            int can = 0; cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (!can) { rt_multi_destroy(m); return fail("rt_multi_create: devices " + std::to_string(devices[i]) + " and " + std::to_string(devices[j]) + " have no peer access"); }
            cudaSetDevice(devices[i]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rt_multi_destroy(m); return fail(std::string("rt_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    *out = m;
    return 0;
}

int rt_multi_scene_create(rt_multi* m, const rt_scene_desc* desc) {
    if (!m || !desc) return fail("rt_multi_scene_create: null argument");
    for (rt_scene* s : m->scene) if (s) rt_scene_destroy(s);
    m->scene.assign(m->ctx.size(), nullptr);
    for (size_t i = 0; i < m->ctx.size(); ++i) if (rt_scene_create(m->ctx[i], desc, &m->scene[i])) return 1;
    return 0;
}

int rt_multi_render(rt_multi* m, const rt_ubo* ubo) {
    if (!m || !ubo) return fail("rt_multi_render: null argument");
    if (m->scene.size() != m->ctx.size() || !m->scene[0]) return fail("rt_multi_render: no scene");
    const uint32_t n = (uint32_t)m->ctx.size();
    if (m->mode == RT_PARTITION_TILES) {
        for (uint32_t i = 0; i < n; ++i) {
            rt_render_opts o; o.flags = 0; o.strip_rows = rt_multi::strip_rows; o.n_parts = n; o.part = i;
            if (rt_render(m->ctx[i], m->scene[i], ubo, n > 1 ? &o : nullptr, nullptr)) return 1;
        }
    } else {
        const uint32_t i = (uint32_t)(m->frames % n);
        if (rt_render(m->ctx[i], m->scene[i], ubo, nullptr, nullptr)) return 1;
    }
    ++m->frames; m->dirty = true;
    return 0;
}

int rt_multi_combine(rt_multi* m, const rt_ubo* ubo) {
    if (!m || !ubo) return fail("rt_multi_combine: null argument");
    const uint32_t n = (uint32_t)m->ctx.size();
    for (uint32_t i = 0; i < n; ++i) {
        void* peers[RT_MAX_PEERS]; uint32_t np = 0; int32_t root = RT_GATHER_NONE;
        for (uint32_t j = 0; j < n; ++j) { if (j == i) continue; if (j == 0) root = (int32_t)np; peers[np++] = m->ctx[j]->acc; }
        rt_combine_desc d; memset(&d, 0, sizeof d);
        d.peer_blocks = peers; d.n_peers = np; d.epoch = m->ctx[i]->combine_epoch + 1;
        if (m->mode == RT_PARTITION_TILES && n > 1) { d.strip_rows = rt_multi::strip_rows; d.n_parts = n; d.part = i; }
        else { const uint32_t per = (m->height + n - 1) / n; d.row0 = per * i < m->height ? per * i : m->height; d.row1 = per * (i + 1) < m->height ? per * (i + 1) : m->height; }
        d.gather_to = i == 0 ? RT_GATHER_NONE : root; d.n_senders = i == 0 ? np : 0; d.gather_acc = 1;
        if (rt_combine(m->ctx[i], &d, ubo, nullptr)) return 1;
    }
    m->dirty = false;
    return 0;
}

int rt_multi_synchronize(rt_multi* m) {
    if (!m) return fail("rt_multi_synchronize: null argument");
    for (rt_context* c : m->ctx) if (rt_synchronize(c)) return 1;
    return 0;
}

int rt_multi_readback(rt_multi* m, const rt_ubo* ubo, float* acc, uint8_t* out) {
    if (!m || !ubo) return fail("rt_multi_readback: null argument");
    if (m->dirty || m->ctx[0]->combine_epoch == 0) if (rt_multi_combine(m, ubo)) return 1;
    if (rt_multi_synchronize(m)) return 1;      // every replica's stores into device 0's block have completed
    return rt_readback_display(m->ctx[0], out, acc);
}

int rt_multi_replica(rt_multi* m, uint32_t i, rt_context** ctx, rt_scene** scene) {
    if (!m || i >= m->ctx.size()) return fail("rt_multi_replica: index out of range");
    if (ctx) *ctx = m->ctx[i];
    if (scene) *scene = i < m->scene.size() ? m->scene[i] : nullptr;
    return 0;
}

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

}  // extern "C"
