// main.cpp — the reference's `main` (src/main.zig:295-406) with its render loop replaced by the CUDA
// library: build the scene (unchanged API), Camera.init (unchanged), flatten, rtw_cuda_upload_scene,
// rtw_cuda_render, write the image (PPM, as north_star asks; the reference writes out.png).
//
// The reference has no CLI (all parameters are source constants); flags here expose
// those constants: --scene N --width W --spp S --depth D --seed K --grid G --variant V --gpus N --asset PATH --out FILE(.ppm|.png)
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rtw.hpp"

using namespace rtw_host;

int main(int argc, char **argv) {
    int scene = 6, grid = 3, width = 0, spp = 0, depth = 0, variant = 0, device = 0, gpus = 1;
    uint64_t seed = 42;
    std::string asset = "assets/sekaichizu.png", out = "out.ppm", lib = "";
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const char *v = argv[i + 1];
        if (k == "--scene") scene = atoi(v);
        else if (k == "--grid") grid = atoi(v);
        else if (k == "--width") width = atoi(v);
        else if (k == "--spp") spp = atoi(v);
        else if (k == "--depth") depth = atoi(v);
        else if (k == "--variant") variant = atoi(v);
        else if (k == "--device") device = atoi(v);
        else if (k == "--gpus") gpus = atoi(v);
        else if (k == "--seed") seed = strtoull(v, nullptr, 10);
        else if (k == "--asset") asset = v;
        else if (k == "--out") out = v;
        else if (k == "--lib") lib = v;
        else { fprintf(stderr, "unknown flag %s\n", k.c_str()); return 2; }
    }
    // The CUDA library is loaded at run time so this binary builds on machines without CUDA; it
    // still refuses to run without it — there is no CPU path.
    if (lib.empty()) {
        Dl_info info;
        std::string dir = ".";
        if (dladdr((void *)&main, &info) && info.dli_fname) {
            dir = info.dli_fname;
            const size_t p = dir.find_last_of('/');
            dir = p == std::string::npos ? "." : dir.substr(0, p);
        }
        lib = dir + "/librtw_cuda.so";
    }
    void *h = dlopen(lib.c_str(), RTLD_NOW);
    if (!h) { fprintf(stderr, "cannot load %s: %s\n", lib.c_str(), dlerror()); return 1; }
#define SYM(name) auto name##_ = reinterpret_cast<decltype(&name)>(dlsym(h, #name)); if (!name##_) { fprintf(stderr, "missing symbol %s\n", #name); return 1; }
    SYM(rtw_cuda_create) SYM(rtw_cuda_destroy) SYM(rtw_cuda_last_error) SYM(rtw_cuda_upload_scene) SYM(rtw_cuda_render) SYM(rtw_cuda_render_multi) SYM(rtw_cuda_create_multi) SYM(rtw_cuda_stats)

    SceneSetup s;
    try { s = makeScene(scene, grid, seed, asset); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return 1; }
    if (width > 0) { s.image_width = (uint32_t)width; s.image_height = (uint32_t)((double)width / s.aspect_ratio); }
    if (spp > 0) s.samples_per_pixel = (uint32_t)spp;
    if (depth > 0) s.max_depth = (uint32_t)depth;
    const Camera cam = Camera::init(s.look_from, s.look_at, {0, 1, 0}, s.vfov, s.aspect_ratio, s.aperture, 10.0, 0, 1);  // main.zig:366-376
    const FlatScene flat = flatten(s.world, 0, 1);
    const rtw_scene_desc desc = flat.desc();

    // --gpus N: devices 0..N-1 with peer mappings between every pair (rtw_cuda_create_multi); else the one device asked for
    std::vector<rtw_ctx *> ctxs((size_t)(gpus > 1 ? gpus : 1), nullptr);
    if (gpus > 1 ? rtw_cuda_create_multi_((uint32_t)gpus, ctxs.data()) : rtw_cuda_create_(device, &ctxs[0])) {
        fprintf(stderr, "rtw_cuda_create: %s\n", rtw_cuda_last_error_(nullptr));
        return 1;
    }
    for (size_t g = 0; g < ctxs.size(); ++g)
        if (rtw_cuda_upload_scene_(ctxs[g], &desc)) { fprintf(stderr, "upload: %s\n", rtw_cuda_last_error_(ctxs[g])); return 1; }
    rtw_ctx *ctx = ctxs[0];
    rtw_render_params p{};
    p.width = s.image_width; p.height = s.image_height;
    p.spp_begin = 0; p.spp_end = s.samples_per_pixel; p.spp_total = s.samples_per_pixel;
    p.max_depth = s.max_depth; p.variant = (uint32_t)variant; p.seed = seed;
    p.background[0] = s.background.x; p.background[1] = s.background.y; p.background[2] = s.background.z;
    std::vector<uint8_t> image((size_t)p.width * p.height * 3);
    const int rc = ctxs.size() > 1 ? rtw_cuda_render_multi_(ctxs.data(), (uint32_t)ctxs.size(), &cam.c, &p, image.data())
                                   : rtw_cuda_render_(ctx, &cam.c, &p, image.data(), nullptr);
    if (rc) { fprintf(stderr, "render: %s\n", rtw_cuda_last_error_(ctx)); return 1; }
    rtw_stats st{};
    rtw_cuda_stats_(ctx, &st);
    const double paths = (double)p.width * p.height * s.samples_per_pixel;
    fprintf(stderr, "scene %d: %u prims, %ux%u, %u spp on %zu GPU(s): %.2f ms wall (%.1f Mpaths/s); device 0: trace %.2f ms, resolve %.3f ms\n",
            scene, desc.n_prims, p.width, p.height, s.samples_per_pixel, ctxs.size(), st.ms_wall, paths / (st.ms_wall * 1e3), st.ms_trace, st.ms_resolve);
    const bool png = out.size() > 4 && out.compare(out.size() - 4, 4, ".png") == 0;  // the reference writes out.png (main.zig:405)
    if (!(png ? writePng(out, image.data(), p.width, p.height) : writePpm(out, image.data(), p.width, p.height))) {
        fprintf(stderr, "cannot write %s\n", out.c_str());
        return 1;
    }
    for (rtw_ctx *c : ctxs) rtw_cuda_destroy_(c);
    return 0;
}
