// gltf_viewer.cpp — headless counterpart of the reference's `gltf_viewer` example
// (crates/examples/gltf_viewer/src/main.rs, args.rs:4-10), written against the two C APIs only:
// include/gltf_host.h (import, animation, camera, GUI state, UBO) and include/rt_b200.h (the render path).
//
//   gltf_viewer -f scene.gltf [-o out.png] [--width W --height H] [--spp N] [--samples-per-frame K] [--bounces B]
//               [--skybox DIR] [--mapping M] [--tone-map T] [--animate SECONDS] [--camera X Y Z] [--device D]
//               [--frames-in-flight N] [--gpus N [--same-device]]
//
// --gpus N renders every frame on N GPUs of this one process through rt_multi (tile partition: interleaved 8-row strips per
// device, gathered on device 0 by the fused peer-memory combine) — the image is bit-identical to one GPU.  --same-device puts
// all replicas on --device (single-GPU boxes, tests).
//
// The window, swapchain and imgui panel are out of scope; GUI defaults (gui_state.rs:303-332)
// stand in for everything not given on the command line.  The draw loop keeps IN_FLIGHT_FRAMES = 2 frames in flight
// like app/src/lib.rs:34,400-401.  Needs a CUDA device: the render path has no CPU fallback.
#include "../include/gltf_host.h"

#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void put_be32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static void png_chunk(std::vector<uint8_t>& out, const char* tag, const std::vector<uint8_t>& data) {
    put_be32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), tag, tag + 4); out.insert(out.end(), data.begin(), data.end());
    put_be32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}
// 8-bit RGB PNG from the RGBA8 storage image (the alpha channel is always 255: RayTracing.rgen:165)
static bool write_png(const char* path, const uint8_t* rgba, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw; raw.reserve((size_t)h * (1 + 3 * (size_t)w));
    for (uint32_t y = 0; y < h; ++y) {
        raw.push_back(0);
        for (uint32_t x = 0; x < w; ++x) { const uint8_t* p = rgba + 4 * ((size_t)y * w + x); raw.push_back(p[0]); raw.push_back(p[1]); raw.push_back(p[2]); }
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr;
    put_be32(ihdr, w); put_be32(ihdr, h); ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr); png_chunk(out, "IDAT", comp); png_chunk(out, "IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

static int die_rt(const char* what) { fprintf(stderr, "gltf_viewer: %s: %s\n", what, rt_last_error()); return 1; }
static int die_gv(const char* what) { fprintf(stderr, "gltf_viewer: %s: %s\n", what, gv_last_error()); return 1; }

int main(int argc, char** argv) {
    std::string file, output = "render.png", skybox;
    uint32_t width = 1920, height = 1080, spp = 64, per_frame = 3, bounces = 5, mapping = 0, tone_map = 0, in_flight = 2;
    int device = 0; bool animate = false, have_cam = false, same_device = false; float anim_t = 0.0f, cam[3] = {0, 0, 0};
    uint32_t gpus = 1;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto need = [&](int n) { if (i + n >= argc) { fprintf(stderr, "gltf_viewer: %s needs %d value(s)\n", a.c_str(), n); exit(2); } };
        if (a == "-f" || a == "--file") { need(1); file = argv[++i]; }
        else if (a == "-o" || a == "--output") { need(1); output = argv[++i]; }
        else if (a == "--width") { need(1); width = (uint32_t)atoi(argv[++i]); }
        else if (a == "--height") { need(1); height = (uint32_t)atoi(argv[++i]); }
        else if (a == "--spp") { need(1); spp = (uint32_t)atoi(argv[++i]); }
        else if (a == "--samples-per-frame") { need(1); per_frame = (uint32_t)atoi(argv[++i]); }
        else if (a == "--bounces") { need(1); bounces = (uint32_t)atoi(argv[++i]); }
        else if (a == "--skybox") { need(1); skybox = argv[++i]; }
        else if (a == "--mapping") { need(1); mapping = (uint32_t)atoi(argv[++i]); }
        else if (a == "--tone-map") { need(1); tone_map = (uint32_t)atoi(argv[++i]); }
        else if (a == "--animate") { need(1); animate = true; anim_t = (float)atof(argv[++i]); }
        else if (a == "--camera") { need(3); have_cam = true; for (int k = 0; k < 3; ++k) cam[k] = (float)atof(argv[++i]); }
        else if (a == "--device") { need(1); device = atoi(argv[++i]); }
        else if (a == "--frames-in-flight") { need(1); in_flight = (uint32_t)atoi(argv[++i]); }
        else if (a == "--gpus") { need(1); gpus = (uint32_t)atoi(argv[++i]); }
        else if (a == "--same-device") { same_device = true; }
        else { fprintf(stderr, "gltf_viewer: unknown argument %s\n", a.c_str()); return 2; }
    }
    if (file.empty()) { fprintf(stderr, "usage: gltf_viewer -f <file> [-o out.png] ...\n"); return 2; }

    gv_doc* doc = nullptr;
    if (gv_load_file(file.c_str(), &doc)) return die_gv("load_file");
    uint8_t* faces[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool sky = false;
    if (!skybox.empty()) {   // SkyBox::new, cubumap.rs:86-106
        uint32_t sw = 0, sh = 0;
        if (gv_load_skybox_dir(skybox.c_str(), faces, &sw, &sh)) return die_gv("skybox");
        if (gv_doc_set_skybox(doc, faces, sw, sh, 1)) return die_gv("skybox");
        sky = true;
    }
    rt_scene_desc desc;
    if (gv_doc_scene_desc(doc, &desc)) return die_gv("scene_desc");

    if (gpus < 1 || gpus > 8) { fprintf(stderr, "gltf_viewer: --gpus must be 1..8\n"); return 2; }
    // one replica (context + scene) per GPU behind rt_multi; a single GPU is the n = 1 case of the same interface
    rt_multi* multi = nullptr;
    std::vector<int> devices;
    for (uint32_t g = 0; g < gpus; ++g) devices.push_back(same_device ? device : device + (int)g);
    if (rt_multi_create(devices.data(), gpus, width, height, RT_PARTITION_TILES, &multi)) return die_rt("rt_multi_create");
    if (rt_multi_scene_create(multi, &desc)) return die_rt("rt_multi_scene_create");
    if (in_flight < 1) in_flight = 1;
    if (in_flight > 4) in_flight = 4;
    if (animate && gv_doc_animate(doc, anim_t)) return die_gv("animate");     // GltfViewer::state_change animation branch (main.rs:377-410)
    for (uint32_t g = 0; g < gpus; ++g) {
        rt_context* ctx = nullptr; rt_scene* scene = nullptr;
        if (rt_multi_replica(multi, g, &ctx, &scene)) return die_rt("rt_multi_replica");
        if (animate) {
            if (gv_doc_need_compute(doc)) {
                const float* mats = nullptr; uint32_t n_skins = 0;
                if (gv_doc_get_skins(doc, &mats, &n_skins)) return die_gv("get_skins");
                if (rt_scene_update_skins(scene, mats, n_skins, 0)) return die_rt("rt_scene_update_skins");
            }
            const rt_instance* inst = nullptr; uint32_t n = 0;
            if (gv_doc_get_instances(doc, &inst, &n)) return die_gv("get_instances");
            if (rt_scene_update_instances(scene, inst, n)) return die_rt("rt_scene_update_instances");
        }
        if (rt_context_set_frames_in_flight(ctx, in_flight)) return die_rt("rt_context_set_frames_in_flight");
    }

    gv_camera camera; gv_camera_default(&camera, width, height);
    if (have_cam) memcpy(camera.position, cam, sizeof cam);
    gv_gui gui; gv_gui_default(&gui);
    gui.number_of_samples = per_frame; gui.number_of_bounces = bounces; gui.max_number_of_samples = spp; gui.sky = sky ? 1u : 0u;
    gui.mapping = mapping; gui.selected_tone_map_mode = tone_map;
    const uint32_t opaque = (uint32_t)gv_doc_fully_opaque(doc);
    uint32_t total = 0, frames = 0;
    rt_ubo ubo, last_ubo; memset(&last_ubo, 0, sizeof last_ubo);
    for (;;) {               // GltfViewer::update + record_raytracing_commands (main.rs:189-270)
        gv_build_ubo(&camera, &gui, &total, frames, opaque, 3u, &ubo);
        if (ubo.number_of_samples == 0) break;
        if (rt_multi_render(multi, &ubo)) return die_rt("rt_multi_render");
        last_ubo = ubo; ++frames;
        if (mapping != 0) break;
    }
    if (!frames) { fprintf(stderr, "gltf_viewer: nothing to render (--spp 0?)\n"); return 2; }
    std::vector<uint8_t> rgba((size_t)width * height * 4);
    rt_context* ctx0 = nullptr;
    if (rt_multi_replica(multi, 0, &ctx0, nullptr)) return die_rt("rt_multi_replica");
    rt_stats st; memset(&st, 0, sizeof st);
    if (gpus == 1) {         // the storage image of the last frame (debug mappings such as HEAT / DISTANCE live there)
        if (rt_readback(ctx0, nullptr, rgba.data())) return die_rt("rt_readback");
    } else {                 // strips of every device gathered on device 0 (rt_combine over peer memory)
        if (rt_multi_readback(multi, &last_ubo, nullptr, rgba.data())) return die_rt("rt_multi_readback");
    }
    if (rt_last_frame_stats(ctx0, &st)) return die_rt("rt_last_frame_stats");
    if (!write_png(output.c_str(), rgba.data(), width, height)) { fprintf(stderr, "gltf_viewer: cannot write %s\n", output.c_str()); return 1; }
    printf("%s: %ux%u, %u spp in %u frames on %u GPU(s), last frame %.2f ms\n", output.c_str(), width, height, total, frames, gpus, st.ms_total);
    rt_multi_destroy(multi); gv_doc_free(doc);
    for (uint8_t* f : faces) gv_free(f);
    return 0;
}
