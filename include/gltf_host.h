/*
 * gltf_host.h — C API of the host layer that sits above rt_b200.h and plays the role of the reference's
 * Rust host for the render path (no Rust toolchain in this environment; DESIGN.md deviation D6).
 * It mirrors, with the same names and argument meaning:
 *   asset_loader::load_file / Doc            crates/libs/asset_loader/src/scene_graph.rs:27-463
 *   Doc::animate / Doc::get_skins            scene_graph.rs:308-337, skinning.rs:39-50, animation.rs:77-146
 *   create_top_as instance list              acceleration_structures.rs:143-172
 *   app::camera::Camera                      crates/libs/app/src/camera.rs:98-118
 *   GltfViewer::update UBO fill              crates/examples/gltf_viewer/src/main.rs:189-242
 *   Gui defaults / sample budgeting          crates/examples/gltf_viewer/src/gui_state.rs:261-332
 * Pure CPU code, no CUDA: it produces the rt_scene_desc / rt_ubo the core consumes.
 */
#ifndef GLTF_HOST_H
#define GLTF_HOST_H
#include "rt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gv_doc gv_doc;

const char* gv_last_error(void);

/* asset_loader::load_file (scene_graph.rs:438-463): import .gltf/.glb, build flat arrays, tag skinned
   vertices, normalise the scene AABB (longest side = 10, centred). */
int  gv_load_file(const char* path, gv_doc** out);
void gv_doc_free(gv_doc* doc);
/* Fill desc with pointers into doc-owned storage (valid until the doc is freed or animated). */
int  gv_doc_scene_desc(gv_doc* doc, rt_scene_desc* desc);
int  gv_doc_fully_opaque(const gv_doc* doc);          /* GeoBuilder::fully_opaque geometry.rs:60-62 */
int  gv_doc_static_scene(const gv_doc* doc);          /* Doc::static_scene scene_graph.rs:325-327 */
int  gv_doc_need_compute(const gv_doc* doc);          /* Doc::need_compute scene_graph.rs:91-94 */
void gv_doc_aabb_trans(const gv_doc* doc, float out16[16]);   /* column-major */
/* Doc::animate(t) then refresh skin matrices and the instance list */
int  gv_doc_animate(gv_doc* doc, float t);
int  gv_doc_get_skins(gv_doc* doc, const float** mats, uint32_t* n_skins);        /* n_skins*256*16 */
int  gv_doc_get_instances(gv_doc* doc, const rt_instance** inst, uint32_t* n);
/* set decoded skybox faces (RGBA8, +x,-x,+y,-y,+z,-z) to be referenced by the scene desc */
int  gv_doc_set_skybox(gv_doc* doc, const uint8_t* const faces[6], uint32_t w, uint32_t h, uint32_t srgb);

/* app::camera::Camera (camera.rs:15-23) */
typedef struct gv_camera {
    float position[3];
    float direction[3];
    float fov;          /* degrees */
    float aspect_ratio;
    float z_near, z_far;
} gv_camera;
void gv_camera_default(gv_camera* cam, uint32_t width, uint32_t height);   /* app/src/lib.rs:331-338 */
void gv_camera_view_matrix(const gv_camera* cam, float out16[16]);         /* look_at_rh, column-major */
void gv_camera_projection_matrix(const gv_camera* cam, float out16[16]);   /* OPENGL_TO_VULKAN_RT * perspective */
int  gv_mat4_inverse(const float in16[16], float out16[16]);

/* gui_state.rs:15-41 (fields that reach the UBO) with Gui::new defaults (:303-332) */
typedef struct gv_gui {
    float    aperture, focus_distance;
    uint32_t number_of_samples, number_of_bounces, max_number_of_samples;
    uint32_t acc, sky, antialiasing, debug, mapping, animation;
    float    map_scale, scale, orthographic_fov_dis, exposure;
    uint32_t selected_tone_map_mode;
} gv_gui;
void gv_gui_default(gv_gui* gui);

/* per-frame bookkeeping of GltfViewer::update (main.rs:207-237): computes number_of_samples,
   updates *total_number_of_samples, fills the UBO.  random_seed is the D1 surrogate seed (3 in the reference). */
void gv_build_ubo(const gv_camera* cam, const gv_gui* gui, uint32_t* total_number_of_samples,
                  uint32_t frame_count, uint32_t fully_opaque, uint32_t random_seed, rt_ubo* out);

/* PNG decoder used for embedded glTF images (8-bit, non-interlaced): returns RGBA8 via malloc */
int  gv_decode_png(const uint8_t* data, size_t size, uint8_t** rgba, uint32_t* w, uint32_t* h);
/* geometry.rs:192-212,296-350: mikktspace::generate_tangents over an indexed triangle list (written into
   vertices[i].tangent; last face touching a vertex wins, w = -1 when the bitangent preserves orientation) */
void gv_generate_tangents(rt_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices);
/* PNG (8 bit) or JPEG (baseline / progressive, 8 bit) -> RGBA8; replaces image::ImageReader::decode (image.rs:60-83) */
int  gv_decode_image(const uint8_t* data, size_t size, uint8_t** rgba, uint32_t* w, uint32_t* h);
/* SkyBox::new (asset_loader/src/cubumap.rs:86-106): six faces of a directory in +x,-x,+y,-y,+z,-z order, RGBA8 (sRGB).
   Each faces[f] is malloc'd: release with gv_free. */
int  gv_load_skybox_dir(const char* dir, uint8_t* faces[6], uint32_t* w, uint32_t* h);
void gv_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
