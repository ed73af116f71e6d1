// host_capi.cpp — C entry points of the C++ host twin, for the Python harness (tests, bench.py).
#include <cstring>
#include <string>

#include "rtw.hpp"

using namespace rtw_host;

namespace {
struct HostScene {
    SceneSetup setup;
    FlatScene flat;
    std::string err;
};
thread_local std::string g_err;
}  // namespace

extern "C" {

const char *rtw_host_last_error(void) { return g_err.c_str(); }

// Build scene `scene_id` (see makeScene), flatten it.  Returns null on failure.
void *rtw_host_scene_create(int scene_id, int grid, uint64_t seed, const char *asset_path) {
    try {
        auto *h = new HostScene();
        h->setup = makeScene(scene_id, grid, seed, asset_path ? asset_path : "");
        h->flat = flatten(h->setup.world, 0.0, 1.0);  // shutter of Camera.init call, main.zig:374-375
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void rtw_host_scene_destroy(void *h) { delete (HostScene *)h; }

void rtw_host_scene_desc(void *h, rtw_scene_desc *out) { *out = ((HostScene *)h)->flat.desc(); }

// look_from[3], look_at[3], vfov, aperture, aspect, background[3] | width, height, spp, max_depth
void rtw_host_scene_config(void *h, double *f12, uint32_t *u4) {
    const SceneSetup &s = ((HostScene *)h)->setup;
    f12[0] = s.look_from.x; f12[1] = s.look_from.y; f12[2] = s.look_from.z;
    f12[3] = s.look_at.x; f12[4] = s.look_at.y; f12[5] = s.look_at.z;
    f12[6] = s.vfov; f12[7] = s.aperture; f12[8] = s.aspect_ratio;
    f12[9] = s.background.x; f12[10] = s.background.y; f12[11] = s.background.z;
    u4[0] = s.image_width; u4[1] = s.image_height; u4[2] = s.samples_per_pixel; u4[3] = s.max_depth;
}

void rtw_host_camera_init(const double *look_from, const double *look_at, const double *vup, double vfov, double aspect,
                          double aperture, double focus_dist, double time0, double time1, rtw_camera *out) {
    *out = Camera::init({look_from[0], look_from[1], look_from[2]}, {look_at[0], look_at[1], look_at[2]},
                        {vup[0], vup[1], vup[2]}, vfov, aspect, aperture, focus_dist, time0, time1).c;
}

int rtw_host_write_ppm(const char *path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    return writePpm(path, rgb8, width, height) ? 0 : 1;
}

int rtw_host_write_png(const char *path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    return writePng(path, rgb8, width, height) ? 0 : 1;
}

// Decode a PNG into caller memory (rgba may be null to query the size).  Returns 0 on success.
int rtw_host_decode_png(const char *path, uint32_t *width, uint32_t *height, uint8_t *rgba, uint64_t capacity) {
    Image im;
    std::string err;
    if (!decodePng(path, im, err)) { g_err = err; return 1; }
    *width = im.width; *height = im.height;
    if (rgba) {
        if (capacity < im.rgba.size()) { g_err = "buffer too small"; return 2; }
        std::memcpy(rgba, im.rgba.data(), im.rgba.size());
    }
    return 0;
}

void rtw_host_random_real01(uint64_t seed, int n, double *out) {
    Random r(seed);
    for (int i = 0; i < n; ++i) out[i] = r.real01();
}

}  // extern "C"
