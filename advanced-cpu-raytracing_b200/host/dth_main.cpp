// raytracer_gpu — the reference's command line (src/main.cpp:132-202: `raytracer scene.xml`) on the GPU path.
// Parse with the host mirror, hand the flat scene to the C ABI once, render every camera with one dt_render
// call each (replacing the thread spawn / join / tonemap block, main.cpp:164-192), write <ImageName>.png.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dorktracer_host.h"

int main(int argc, char* argv[]) {
    if (argc < 2) { fprintf(stderr, "usage: %s scene.xml [--device N] [--seed S]\n", argv[0]); return 2; }
    int device = 0; unsigned long long seed = 1234;
    for (int i = 2; i + 1 < argc; i++) {
        if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--seed")) seed = strtoull(argv[i + 1], nullptr, 10);
    }
    dth_scene* hs = nullptr;
    if (dt_gpu_init(device) < 0) { fprintf(stderr, "%s\n", dt_last_error()); return 1; }
    dth_set_bvh_builder(dt_bvh2_build, 32768);        // Mesh::ConstructBVH of the large meshes on the GPU
    if (dth_scene_load_xml(argv[1], &hs) != DT_OK) { fprintf(stderr, "%s / %s\n", dth_last_error(), dt_last_error()); return 1; }
    for (int i = 0; i < dth_scene_desc(hs)->n_images; i++)
        if (!dth_scene_image_loaded(hs, i)) { fprintf(stderr, "image '%s' could not be decoded (PNG/EXR only in the stand-alone driver)\n", dth_scene_image_path(hs, i)); return 1; }
    dt_scene* gs = nullptr;
    if (dt_scene_create(dth_scene_desc(hs), &gs) != DT_OK) { fprintf(stderr, "%s\n", dt_last_error()); return 1; }
    // main.cpp:186-195 encodes HDR + PNG on the main thread after every camera; here the files are encoded by a writer queue
    // (host/dth_output.cpp) while the next camera renders, and joined once at the end (still inside "Rendering took").
    dth_writer* writer = dth_writer_create(2, 0);
    auto start = std::chrono::steady_clock::now();
    for (int c = 0; c < dth_scene_num_cameras(hs); c++) {
        const dt_camera_desc* cam = dth_scene_camera(hs, c);
        std::vector<uint8_t> ldr((size_t)cam->width * cam->height * 3);
        std::vector<float> hdr;
        if (cam->has_tonemapper) hdr.resize((size_t)cam->width * cam->height * 3);
        dt_render_params p; memset(&p, 0, sizeof p); p.seed = seed; p.tile_world = 1;
        dt_stats st;
        printf("Resolution: %dx%d, Running on: GPU %d\n", cam->width, cam->height, device);
        if (dt_render(gs, cam, &p, ldr.data(), hdr.empty() ? nullptr : hdr.data(), &st) != DT_OK) { fprintf(stderr, "%s\n", dt_last_error()); return 1; }
        std::string name = dth_scene_camera_image_name(hs, c);
        if (cam->has_tonemapper) dth_writer_submit_hdr(writer, name.c_str(), cam->width, cam->height, hdr.data());
        size_t dot = name.find_last_of('.');
        std::string png = name.substr(0, dot) + ".png";
        if (dth_writer_submit_png(writer, png.c_str(), cam->width, cam->height, ldr.data()) != DT_OK) { fprintf(stderr, "cannot queue %s\n", png.c_str()); return 1; }
        printf("%s: %llu closest + %llu shadow rays, %.3f ms on device (%.1f Mrays/s), %u waves\n", png.c_str(),
               (unsigned long long)st.rays_closest, (unsigned long long)st.rays_shadow, st.ms_total,
               (st.rays_closest + st.rays_shadow) / (st.ms_total * 1e3), st.waves);
    }
    double encode_s = 0.0;
    if (dth_writer_wait(writer, &encode_s) != DT_OK) { fprintf(stderr, "%s\n", dth_last_error()); return 1; }
    auto end = std::chrono::steady_clock::now();
    printf("Image encoding: %gs on the writer threads (overlapped with rendering)\n", encode_s);
    printf("Rendering took: %gs\n", std::chrono::duration<double>(end - start).count());
    dth_writer_destroy(writer);
    dt_scene_destroy(gs);
    dth_scene_free(hs);
    return 0;
}
