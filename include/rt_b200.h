/*
 * rt_b200.h — C ABI of the B200-native path-tracing core for rustracer.
 *
 * This is the drop-in boundary.  The reference (KaminariOS/rustracer) has no
 * FFI of its own: its `gltf_viewer` + `asset_loader` crates call the in-tree `vulkan` crate for
 * everything that touches the GPU.  Each entry point below replaces one group of those calls and
 * cites it.  All structs are the reference's own `#[repr(C)]` / std430 layouts, byte for byte, so a
 * Rust host can pass its existing vectors unchanged.
 *
 * Conventions: every function returns 0 on success, non-zero on failure; rt_last_error() returns a
 * thread-local message.  Inputs are copied (caller keeps ownership).  A context and the scenes
 * created from it are not thread-safe (the reference drives Vulkan from one thread).
 * There is no CPU fallback: if no CUDA device is usable rt_context_create fails.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Binary layouts
 * ---------------------------------------------------------------------------------------- */

/* crates/libs/asset_loader/src/geometry.rs:10-22  ==  shaders/lib/RayTracingCommons.glsl:54-63 */
typedef struct rt_vertex {
    float    position[4];   /* xyz, w = 0 */
    float    normal[4];     /* xyz, w = 0 */
    float    tangent[4];    /* xyz + handedness */
    float    color[4];
    float    weights[4];
    uint32_t joints[4];
    float    uv0[2];
    float    uv1[2];
    int32_t  skin_index;    /* -1 = not skinned */
    uint32_t _pad[3];
} rt_vertex;                /* 128 B */

/* geometry.rs:65-73 == RayTracingCommons.glsl:47-52 */
typedef struct rt_prim_info {
    uint32_t v_offset;
    uint32_t i_offset;
    uint32_t material_id;
    uint32_t _pad;
} rt_prim_info;             /* 16 B */

/* material.rs:11-17 == Material.glsl:2-5.  index -1 = none; real indices are glTF index + 1 */
typedef struct rt_texture_info {
    int32_t index;
    int32_t coord;
} rt_texture_info;

/* material.rs:130-160 == Material.glsl:51-73 */
typedef struct rt_material {
    uint32_t        alpha_mode;          /*   0: 1 OPAQUE, 2 MASK, 3 BLEND (RayTracing.rahit:39-41) */
    float           alpha_cutoff;        /*   4 */
    uint32_t        double_sided;        /*   8 (never read by a shader) */
    uint32_t        workflow;            /*  12: 0 metallic-roughness, 1 specular-glossiness */
    float           _pad0[2];            /*  16 */
    rt_texture_info base_color_texture;  /*  24 */
    float           base_color[4];       /*  32 */
    float           metallic_factor;     /*  48 */
    float           roughness_factor;    /*  52 */
    rt_texture_info metallic_roughness_texture; /* 56 */
    rt_texture_info normal_texture;      /*  64 */
    rt_texture_info emissive_texture;    /*  72 */
    float           emissive_factor[4];  /*  80 */
    rt_texture_info occlusion_texture;   /*  96 (unused) */
    float           ior;                 /* 104 */
    uint32_t        unlit;               /* 108 */
    rt_texture_info transmission_texture;/* 112 */
    float           transmission_factor; /* 120 */
    uint32_t        transmission_exist;  /* 124 */
    float           attenuation_color[3];/* 128 */
    float           thickness_factor;    /* 140 */
    rt_texture_info thickness_texture;   /* 144 */
    float           attenuation_distance;/* 152 */
    uint32_t        volume_exists;       /* 156 */
    rt_texture_info specular_texture;    /* 160 */
    rt_texture_info specular_color_texture; /* 168 */
    float           specular_color_factor[4]; /* 176 */
    float           specular_factor;     /* 192 */
    uint32_t        specular_exist;      /* 196 */
    float           _pad1[2];            /* 200 */
    float           sg_diffuse_factor[4];/* 208 */
    float           sg_specular_glossiness_factor[4]; /* 224: rgb specular, a glossiness */
    rt_texture_info sg_diffuse_texture;  /* 240 */
    rt_texture_info sg_specular_glossiness_texture; /* 248 */
} rt_material;              /* 256 B */

/* light.rs:62-71 == PunctualLight.glsl:8-15 */
typedef struct rt_light {
    float    color[4];
    float    transform[4];  /* position (point) or direction (directional) */
    uint32_t kind;          /* 0 directional, 1 point */
    float    range;
    float    intensity;
    uint32_t _pad;
} rt_light;                 /* 48 B */

/* gltf_viewer/src/ubo.rs:3-31 == UniformBufferObject.glsl:3-30.  Matrices are column-major. */
typedef struct rt_ubo {
    float    model_view[16];
    float    projection[16];
    float    model_view_inverse[16];
    float    projection_inverse[16];
    float    aperture;
    float    focus_distance;
    float    fov_angle;
    float    orthographic_fov_dis;
    float    heatmap_scale;
    uint32_t total_number_of_samples;
    uint32_t number_of_samples;
    uint32_t number_of_bounces;
    uint32_t random_seed;   /* deviation D1: seeds the clockARB() surrogate */
    uint32_t has_sky;
    uint32_t antialiasing;
    uint32_t mapping;
    uint32_t frame_count;
    uint32_t debug;
    uint32_t fully_opaque;
    float    exposure;
    uint32_t tone_mapping_mode;
} rt_ubo;                   /* 324 B */

/* mapping values: RayTracingCommons.glsl:139-150 */
enum {
    RT_MAP_RENDER = 0, RT_MAP_HEAT = 1, RT_MAP_INSTANCE = 2, RT_MAP_TRIANGLE = 3, RT_MAP_DISTANCE = 4,
    RT_MAP_ALBEDO = 5, RT_MAP_METALLIC = 6, RT_MAP_ROUGHNESS = 7, RT_MAP_NORMAL = 8, RT_MAP_TANGENT = 9,
    RT_MAP_TRANSMISSION = 10, RT_MAP_GEO_ID = 11
};

#define RT_MAX_JOINTS 256   /* skinning.rs:9 ; one skin = 256 column-major mat4 = 16384 B */

/* geometry.rs:38-43 (GeoBuilder.len / .opaque): one entry per glTF primitive == one BLAS */
typedef struct rt_geometry {
    uint32_t v_len;
    uint32_t i_len;
    uint32_t opaque;        /* VkGeometryFlags OPAQUE iff material alphaMode == OPAQUE */
    uint32_t _pad;
} rt_geometry;

/* VkAccelerationStructureInstanceKHR as filled by acceleration_structures.rs:144-171 */
typedef struct rt_instance {
    float    transform[12]; /* object->world, 3x4 row-major */
    uint32_t geo_id;        /* instanceCustomIndex == BLAS index */
    uint32_t mask;          /* 0xFF */
    uint32_t flags;
    uint32_t _pad;
} rt_instance;              /* 64 B */

typedef struct rt_image_desc {      /* asset_loader/src/image.rs:14-23 */
    const uint8_t* rgba8;           /* width*height*4, row 0 first */
    uint32_t width, height;
    uint32_t srgb;                  /* TexGamma::Srgb => 1 */
    uint32_t _pad;
} rt_image_desc;

enum { RT_FILTER_NEAREST = 0, RT_FILTER_LINEAR = 1 };
enum { RT_WRAP_CLAMP = 0, RT_WRAP_MIRROR = 1, RT_WRAP_REPEAT = 2 };

typedef struct rt_sampler_desc {    /* asset_loader/src/texture.rs:23-29, globals.rs:362-382 */
    uint32_t mag_filter, min_filter;
    uint32_t wrap_s, wrap_t;
} rt_sampler_desc;

typedef struct rt_texture_desc {    /* globals.rs:336-340: [index, image_index, sampler_index] */
    uint32_t image_index;
    uint32_t sampler_index;
} rt_texture_desc;

typedef struct rt_scene_desc {
    const rt_vertex*       vertices;     uint32_t n_vertices;
    const uint32_t*        indices;      uint32_t n_indices;
    const rt_prim_info*    prim_infos;   /* n_geometries */
    const rt_geometry*     geometries;   uint32_t n_geometries;
    const rt_material*     materials;    uint32_t n_materials;
    const rt_instance*     instances;    uint32_t n_instances;
    const rt_image_desc*   images;       uint32_t n_images;
    const rt_sampler_desc* samplers;     uint32_t n_samplers;
    const rt_texture_desc* textures;     uint32_t n_textures;
    const rt_light*        dlights;      uint32_t n_dlights;
    const rt_light*        plights;      uint32_t n_plights;
    const float*           skins;        uint32_t n_skins;     /* n_skins * 256 * 16 floats */
    const uint8_t*         skybox_faces[6];                    /* +x,-x,+y,-y,+z,-z ; may be NULL */
    uint32_t               skybox_width, skybox_height, skybox_srgb;
    uint32_t               _pad;
} rt_scene_desc;

/* ------------------------------------------------------------------------------------------
 * Parity-test records (new; SURVEY.md §8b last rows)
 * ---------------------------------------------------------------------------------------- */
typedef struct rt_ray {
    float origin[3]; float tmin;
    float direction[3]; float tmax;
} rt_ray;                   /* 32 B */

typedef struct rt_hit {
    float    t;             /* < 0 on miss */
    float    u, v;          /* barycentric weights of vertex 1 and vertex 2 (gl HitAttributes) */
    uint32_t instance_id;   /* gl_InstanceID ; 0xFFFFFFFF on miss */
    uint32_t primitive_id;  /* gl_PrimitiveID */
    uint32_t geo_id;        /* gl_InstanceCustomIndexEXT */
} rt_hit;                   /* 24 B */

enum {
    RT_TRACE_OPAQUE = 1,    /* gl_RayFlagsOpaqueEXT: never run the alpha test */
    RT_TRACE_SCALAR = 2     /* cross-check path: one thread per ray, plain while-while loop.  Default (flag clear): the
                               ray set runs through the persistent wavefront traversal of the frame kernels
                               (dynamic fetch + warp-cooperative triangle rounds), i.e. the code rt_render times */
};

typedef struct rt_stats {
    float    ms_total, ms_raygen, ms_extend, ms_shade, ms_shadow, ms_accum;
    uint32_t n_extend_launches, n_kernel_launches;
    uint64_t rays_extend;   /* path segments traced (every primary traceRayEXT equivalent) */
    uint64_t rays_shadow;   /* shadow rays traced */
    uint64_t shaded_hits;
    uint64_t pixel_samples;
    /* traversal counters, only filled when RT_RENDER_COUNTERS was requested */
    uint64_t nodes, tris, insts, anyhits, tex_taps, light_cands;
} rt_stats;

enum {
    RT_RENDER_COUNTERS   = 1,   /* count nodes/tris/instances visited (slower; for the bytes/ray figure) */
    RT_RENDER_NO_TONEMAP = 2    /* skip the RGBA8 pass (used by sample-pass sharding before the reduce) */
};

typedef struct rt_render_opts {
    uint32_t flags;
    /* tile partition: rows are cut into strips of strip_rows; this context renders
       strips with (strip % n_parts) == part.  n_parts = 0 or 1 => whole frame. */
    uint32_t strip_rows, n_parts, part;
} rt_render_opts;

typedef struct rt_context rt_context;
typedef struct rt_scene   rt_scene;

/* ------------------------------------------------------------------------------------------
 * Entry points
 * ---------------------------------------------------------------------------------------- */
const char* rt_last_error(void);
const char* rt_version(void);

/* replaces ContextBuilder::build (vulkan/src/context.rs:88-179) + BaseApp::new image creation
   (app/src/lib.rs:305-325).  device = CUDA ordinal. */
int rt_context_create(int device, uint32_t width, uint32_t height, rt_context** out);
void rt_context_destroy(rt_context* ctx);
/* replaces BaseApp::recreate_swapchain (app/src/lib.rs:355-385): drops the accumulation.  With unchanged dimensions the
   images are cleared in place (allocations and peers' IPC mappings stay valid); a size change after rt_ipc_export fails. */
int rt_frame_resize(rt_context* ctx, uint32_t width, uint32_t height);

/* replaces InFlightFrames (app/src/lib.rs:34 IN_FLIGHT_FRAMES = 2, :329, :400-401 per-frame fence) and the
   per-swapchain-image storage images beside the single accumulation image (app/src/lib.rs:305-325).
   n = 1 (default): rt_render runs in strict order on the stream it is given.
   n = 2..4: every rt_render call gets its own slot (private stream, path-state / hit / shadow queues,
   frame radiance and RGBA8 output image); the path tracing of up to n frames overlaps on the GPU (the thin
   tails of one frame's late bounces fill with the next frame's rays) while the accumulate+tonemap steps
   run in submission order on the one shared accumulation image, so the images are bit-identical to n = 1.
   A frame starts after everything queued on rt_render's `stream` at the time of the call.  The library's
   own consumers (rt_readback, rt_tonemap, rt_combine, rt_last_frame_stats, rt_synchronize, the
   rt_scene_update_* calls, rt_frame_resize) wait for the frames in flight; work of your own on another
   stream must call rt_join first.  Keeps the accumulation image; synchronises. */
int rt_context_set_frames_in_flight(rt_context* ctx, uint32_t n);
/* makes `stream` (NULL = the context's) wait, on the device, for every frame submitted so far: after it, the images
   of rt_device_ptrs may be consumed on that stream (tests/test_gpu_parity.py::test_join_orders_foreign_stream...) */
int rt_join(rt_context* ctx, void* stream);
/* presentation of the frame submitted last (the storage image -> swapchain copy of app/src/lib.rs:563-611
   recorded in the same command buffer as the frame): queues the RGBA8 image's device->host copy behind that
   frame on its slot and returns the frame's ticket.  out_rgba8 should be pinned and must stay untouched until
   rt_frame_wait(ticket) returned (the in-flight fence wait, app/src/lib.rs:401); frames in flight need
   distinct host buffers.  The host may run further ahead than the number of slots (slot reuse is ordered on the
   device): keep one host buffer per frame between a submission and its rt_frame_wait. */
int rt_readback_async(rt_context* ctx, uint8_t* out_rgba8, uint64_t* ticket);
int rt_frame_wait(rt_context* ctx, uint64_t ticket);

/* replaces create_global + Buffers::new (asset_loader/src/globals.rs:309,43), create_as
   (acceleration_structures.rs:79-129), create_pipeline + SBT (gltf_viewer/src/pipeline_res.rs:21),
   create_descriptor_sets (desc_sets.rs:21) and the initial ComputeUnit::dispatch (main.rs:85-91) */
int rt_scene_create(rt_context* ctx, const rt_scene_desc* desc, rt_scene** out);
void rt_scene_destroy(rt_scene* scene);

/* replaces create_top_as + update_tlas (acceleration_structures.rs:136 ; main.rs:397-409,424-436) */
int rt_scene_update_instances(rt_scene* scene, const rt_instance* instances, uint32_t n);
/* replaces skin.copy_data_to_buffer + ComputeUnit::dispatch + create_as (main.rs:384-395,
   compute_unit.rs:30-68): skinning kernel, BLAS refit + TLAS refit (asynchronous: queued on the scene's stream,
   later frames wait for it on the device) or, with rebuild != 0, full BLAS + TLAS rebuild (synchronous) */
int rt_scene_update_skins(rt_scene* scene, const float* skin_mats, uint32_t n_skins, int rebuild);
/* Skinned scenes keep n (1..4, default 2 when the scene has skins, else 1) copies of what a skin update rewrites
   (skinned vertices, packed triangles, BLAS / TLAS nodes).  A refit writes the next copy and only waits, on the
   device, for the frames that still read it, so the update of frame f+1 overlaps the frames in flight; the host
   never blocks (rebuild != 0 and rt_scene_update_instances stay synchronous).  Synchronises. */
int rt_scene_set_versions(rt_scene* scene, uint32_t n);
/* A scene belongs to the context it was created on: updates wait for THAT context's frames.  Other contexts of the
   same device may render it (read-only) only while no rt_scene_update_* call is in progress. */
/* replaces dlights_buffer / plights_buffer.copy_data_to_buffer (main.rs:353-374) */
int rt_scene_update_lights(rt_scene* scene, const rt_light* dlights, uint32_t n_dlights,
                           const rt_light* plights, uint32_t n_plights);
/* replaces SkyboxResource::new + update_desc (globals.rs:175-199 ; main.rs:346-350) */
int rt_scene_set_skybox(rt_scene* scene, const uint8_t* const faces[6], uint32_t width,
                        uint32_t height, uint32_t srgb);

/* replaces ubo_buffer.copy_data_to_buffer + bind_* + trace_rays (main.rs:239,254-267): one frame of
   ubo->number_of_samples spp at ubo->number_of_bounces depth, accumulated per RayTracing.rgen:132-166.
   opts may be NULL.  stream is a cudaStream_t or NULL (context's own stream).  Asynchronous.
   With frames in flight (rt_context_set_frames_in_flight) the frame runs on its slot's stream. */
int rt_render(rt_context* ctx, rt_scene* scene, const rt_ubo* ubo, const rt_render_opts* opts,
              void* stream);
/* accumulate+tonemap only (RayTracing.rgen:132-166 tail) over the current accumulation image;
   used after a cross-GPU reduce of the accumulation buffer. */
int rt_tonemap(rt_context* ctx, const rt_ubo* ubo, void* stream);
int rt_synchronize(rt_context* ctx);

/* replaces the storage-image -> swapchain copy (app/src/lib.rs:563-611).  Either may be NULL.
   Synchronises the context first. */
int rt_readback(rt_context* ctx, float* acc_rgba32f, uint8_t* out_rgba8);
int rt_upload_accumulation(rt_context* ctx, const float* acc_rgba32f);     /* resume */
/* zero-copy interop: device pointers of the RGBA32F accumulation and RGBA8 output images */
int rt_device_ptrs(rt_context* ctx, void** acc, void** out);
/* replaces the GPU timestamp queries (app/src/lib.rs:549-555,652-656).  Synchronises. */
int rt_last_frame_stats(rt_context* ctx, rt_stats* stats);

/* parity-test entry points (no reference equivalent: traversal lives in the Vulkan driver).
   rng4: optional n x uvec4 payload RNG state used by the BLEND alpha test (NULL = zeros). */
int rt_trace_closest(rt_scene* scene, const rt_ray* rays, uint32_t n, uint32_t flags,
                     const uint32_t* rng4, rt_hit* hits);
int rt_trace_any(rt_scene* scene, const rt_ray* rays, uint32_t n, uint32_t flags,
                 const uint32_t* rng4, uint8_t* occluded);
/* read back the (skinned) vertex buffer the BLASes were built from (AnimationCompute.comp output) */
int rt_scene_read_vertices(rt_scene* scene, rt_vertex* out, uint32_t n);
/* diagnostics / tests: the 128-byte wide nodes (8 x float4: header, child meta, bfloat16 child planes) of one BLAS (geo >= 0: that geometry's own BLAS, geo == -1: the
   merged world-space BLAS, geo == -2: the TLAS).  *n_nodes receives the node count; out may be NULL to query it. */
int rt_scene_read_nodes(rt_scene* scene, int geo, float* out, uint32_t max_nodes, uint32_t* n_nodes);
/* BVH statistics for DESIGN/bench: nodes, max depth, bytes */
typedef struct rt_bvh_info {
    uint64_t blas_nodes, blas_tris, tlas_nodes, bytes;
    uint32_t max_depth_blas, max_depth_tlas;
    float    build_ms, refit_ms, skin_ms, tlas_ms;
} rt_bvh_info;
int rt_scene_bvh_info(rt_scene* scene, rt_bvh_info* info);

/* multi-GPU (new): cross-process peer access to a context's accumulation block (see rt_combine below).
   rt_ipc_export writes a 64-byte cudaIpcMemHandle of the block; rt_ipc_open maps a peer's handle and returns the block's
   device pointer in this process (close it with rt_ipc_close before the peer resizes). */
int rt_ipc_export(rt_context* ctx, void* handle64);
int rt_ipc_open(rt_context* ctx, const void* handle64, void** peer_block);
int rt_ipc_close(rt_context* ctx, void* peer_block);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU combine, device-synchronised
 * ------------------------------------------------------------------------------------------
 * Every context's accumulation image heads ONE device allocation, the "accumulation block":
 *     [ acc RGBA32F | snap RGBA32F | display RGBA8 | sync words ]
 * rt_ipc_export / rt_ipc_open (other processes) or rt_combine_ptrs (same process, peer access enabled) hand the block's
 * base pointer to the other GPUs.  rt_combine then runs, on each rank and without any host-side wait:
 *   snapshot acc -> snap; publish the epoch; wait (on the device) for the peers' snapshots of the same epoch;
 *   ONE kernel: sum the peers' snap rows of this rank's band over NVLink peer loads, tonemap (RayTracing.rgen:132-166),
 *   store the RGBA8 band into the own display AND (peer stores) into the peers' displays = reduce-scatter + tonemap +
 *   all-gather fused; notify the peers; wait until the bands this rank receives have landed.
 * acc is only read: frames keep accumulating while a combine is in flight (periodic display refresh on a side stream,
 * §8d config 5 "reduce every 64 frames").  After the call has completed on `stream`, the display holds the complete
 * tonemapped image (on every rank for RT_GATHER_ALL, on the root for a rooted gather) and snap holds the reduced RGBA32F
 * band of this rank (the whole image on the root with gather_acc).  With n_parts > 1 (tile partition, §8e A) nothing is
 * summed: each rank contributes the strips it rendered (bit-identical to the single-GPU image).
 * A peer that never shows up makes the device-side waits give up after ~5 s; rt_readback_display then reports it. */
enum { RT_GATHER_ALL = -1, RT_GATHER_NONE = -2 };
typedef struct rt_combine_desc {
    void* const* peer_blocks;   /* accumulation-block base pointers of the other ranks (rt_ipc_open / rt_combine_ptrs) */
    uint32_t n_peers;           /* <= 8 */
    uint32_t epoch;             /* 1, 2, 3, ... : the n-th combine; the same value on every rank */
    uint32_t row0, row1;        /* sample passes: the row band this rank reduces and tonemaps */
    uint32_t strip_rows, n_parts, part;   /* tiles (n_parts > 1): the strips this rank rendered (rt_render_opts) */
    int32_t  gather_to;         /* who receives this rank's band: RT_GATHER_ALL, RT_GATHER_NONE or an index into peer_blocks */
    uint32_t n_senders;         /* how many ranks store their band into THIS rank's display (n_peers for RT_GATHER_ALL / on the root, else 0) */
    uint32_t gather_acc;        /* != 0 with a rooted gather: the reduced RGBA32F bands are assembled in the root's snap as well */
} rt_combine_desc;
int rt_combine(rt_context* ctx, const rt_combine_desc* desc, const rt_ubo* ubo, void* stream);
/* synchronises, then copies the display image (and, optionally, snap = the reduced RGBA32F image / band) to the host */
int rt_readback_display(rt_context* ctx, uint8_t* out_rgba8, float* sum_rgba32f);
/* same-process peers: base pointer and size of the accumulation block, device pointer of the display image */
int rt_combine_ptrs(rt_context* ctx, void** block, void** display, uint64_t* block_bytes);

/* ------------------------------------------------------------------------------------------
 * One process, several GPUs (SURVEY.md §8b last row: "rt_context_create with n > 1 devices => replicas; partition mode
 * enum {TILES, SAMPLE_PASSES}").  The reference is a single process (app/src/lib.rs:133-252): rt_multi gives it N device
 * replicas behind one handle — contexts with peer access enabled in every direction, one scene replica per device
 * (deterministic builder: identical trees), frames partitioned by mode, rt_combine for the exchange.
 * ---------------------------------------------------------------------------------------- */
enum { RT_PARTITION_TILES = 0, RT_PARTITION_SAMPLE_PASSES = 1 };
typedef struct rt_multi rt_multi;
int  rt_multi_create(const int* devices, uint32_t n_devices, uint32_t width, uint32_t height, uint32_t mode, rt_multi** out);
void rt_multi_destroy(rt_multi* m);
/* replaces the scene upload + AS build on every replica */
int  rt_multi_scene_create(rt_multi* m, const rt_scene_desc* desc);
/* TILES: every device renders its interleaved 8-row strips of THIS frame (latency mode; image bit-identical to one GPU).
   SAMPLE_PASSES: the frame goes, whole, to device (frames submitted so far) % n (throughput mode; private RGBA32F sums). */
int  rt_multi_render(rt_multi* m, const rt_ubo* ubo);
/* the exchange step: rt_combine on every device (rooted gather to device 0 incl. the RGBA32F sum), asynchronous */
int  rt_multi_combine(rt_multi* m, const rt_ubo* ubo);
/* rt_multi_combine if frames were submitted since the last one, then the complete image from device 0 (either may be NULL) */
int  rt_multi_readback(rt_multi* m, const rt_ubo* ubo, float* acc_rgba32f, uint8_t* out_rgba8);
int  rt_multi_synchronize(rt_multi* m);
/* replica i, for calls this interface does not wrap (rt_scene_update_*, rt_last_frame_stats, ...) */
int  rt_multi_replica(rt_multi* m, uint32_t i, rt_context** ctx, rt_scene** scene);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
