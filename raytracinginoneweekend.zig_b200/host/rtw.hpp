// rtw.hpp — C++ twin of the reference's Zig host side (`src/rtw.zig` + the scene half of
// `src/main.zig`).  The reference's toolchain (zig >= 0.14) is not in this image, so the host above
// the C ABI is written in C++ with the reference's own interface: same type names, same
// constructors (`makeSphere`, `Hittable::makeBox/makeTranslate/makeRotateY`, `Texture::makeSolid/
// makeChecker/makeNoise/makeImage`, `Camera::init`), same ownership model (`Rc<T>`).  The render
// loop itself (src/main.zig:382-402) is NOT here: it is the CUDA library behind include/rtw_cuda.h.
// zig/ holds the same flatten pass written against the real Zig types (uncompiled here).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <variant>
#include <vector>

#include "rtw_cuda.h"

namespace rtw_host {

// src/rc.zig:3-47 — shared ownership; pointer identity is the material identity the flatten pass
// dedupes on.
template <class T>
using Rc = std::shared_ptr<T>;

struct Vec3 {  // src/rtw/vec.zig:8-11
    double x = 0, y = 0, z = 0;
};
using Point3 = Vec3;
using Color = Vec3;
inline Color rgb(double r, double g, double b) { return {r, g, b}; }

// std.Random.DefaultPrng (Xoshiro256++ / SplitMix64 seeding) + Random.float(f64) +
// Random.uintLessThan, Zig >= 0.14 std; call sites src/main.zig:300-301, src/rtw/rand.zig:5-20.
class Random {
  public:
    explicit Random(uint64_t seed);
    uint64_t next();
    double real01();                         // rand.zig:13-15
    double real(double lo, double hi);       // rand.zig:18-20
    uint64_t intLessThan(uint64_t bound);    // rand.zig:7-10 with min = 0
  private:
    uint64_t s_[4];
};

// ---- textures: src/rtw/texture.zig:10-44 ----------------------------------------------------------
struct Texture;
struct Perlin {  // src/rtw/perlin.zig:10-38
    double ranvec[256][3];
    uint32_t perm[3][256];
    explicit Perlin(Random &rng);
};
struct Image {  // decoded RGBA8, row 0 = top (what zigimg hands texture.zig:131-137)
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> rgba;
};
struct SolidTexture { Color color; };
struct CheckerTexture { Rc<Texture> odd, even; };
struct NoiseTexture { Rc<Perlin> perlin; double scale; };
struct ImageTexture { Rc<Image> image; };
struct Texture {
    std::variant<SolidTexture, CheckerTexture, NoiseTexture, ImageTexture> v;
    static Texture makeSolid(Color c);
    static Texture makeChecker(Color odd, Color even);
    static Texture makeNoise(double scale, Random &rng);
    static Texture makeImage(const std::string &path);  // .png (zlib inflate) or raw .rgba with a sidecar size
    static Texture makeImageFromMemory(const uint8_t *rgba, uint32_t w, uint32_t h);
};

// ---- materials: src/rtw/material.zig:16-20 -----------------------------------------------------------
struct DiffuseMaterial { Texture albedo; };
struct MetalMaterial { Color albedo; double fuzz; };
struct DielectricMaterial { double ir; };
struct DiffuseLightMaterial { Texture emit; };
struct Material {
    std::variant<DiffuseMaterial, MetalMaterial, DielectricMaterial, DiffuseLightMaterial> v;
};

// ---- hittables: src/rtw/hittable.zig:22-33 -----------------------------------------------------------
struct Hittable;
struct Sphere { Point3 center; double radius; Rc<Material> material; };
struct MovingSphere { Point3 center0, center1; double time0, time1, radius; Rc<Material> material; };
struct HittableList { std::vector<Hittable> objects; };
struct XyRect { double x0, x1, y0, y1, k; Rc<Material> material; };
struct XzRect { double x0, x1, z0, z1, k; Rc<Material> material; };
struct YzRect { double y0, y1, z0, z1, k; Rc<Material> material; };
struct Box { Point3 min, max; std::shared_ptr<HittableList> sides; };
struct Translate { Rc<Hittable> object; Vec3 offset; };
struct RotateY { Rc<Hittable> object; double sin_t, cos_t; };
struct Hittable {
    std::variant<Sphere, MovingSphere, HittableList, XyRect, XzRect, YzRect, Box, Translate, RotateY> v;
    static Hittable makeBox(Point3 p0, Point3 p1, Rc<Material> material);     // hittable.zig:35-37, 434-451
    static Hittable makeTranslate(Rc<Hittable> obj, Vec3 offset);             // hittable.zig:39-41
    static Hittable makeRotateY(Rc<Hittable> obj, double angle);              // hittable.zig:43-45, 513-515
};
Hittable makeSphere(Point3 center, double radius, Rc<Material> material);     // main.zig:26-34

// ---- camera: src/main.zig:40-89 ------------------------------------------------------------------------
struct Camera {
    rtw_camera c;
    static Camera init(Point3 look_from, Point3 look_at, Vec3 vup, double vfov, double aspect_ratio, double aperture,
                       double focus_dist, double time0, double time1);
};

// ---- flatten: Rc graph -> the POD arrays of include/rtw_cuda.h ------------------------------------------
struct FlatScene {
    std::vector<rtw_prim> prims;
    std::vector<rtw_xform> xforms;
    std::vector<rtw_material> materials;
    std::vector<rtw_texture> textures;
    std::vector<rtw_image> images;
    std::vector<rtw_perlin> perlins;
    std::vector<Rc<Image>> keep_images;
    std::vector<Rc<Perlin>> keep_perlins;
    std::vector<std::vector<double>> ranvec_store;
    double time0 = 0, time1 = 1;
    rtw_scene_desc desc() const;
};
FlatScene flatten(const Hittable &world, double time0, double time1);

// ---- scenes: src/main.zig:124-293 and the per-scene settings of main.zig:304-362 ---------------------------
struct SceneSetup {
    Hittable world;
    Point3 look_from, look_at;
    double vfov = 40, aperture = 0, aspect_ratio = 3.0 / 2.0;
    Color background;
    uint32_t image_width = 600, image_height = 400, samples_per_pixel = 50, max_depth = 50;
};
// scene 1..6 = the reference's `scene` constant; 7 = config C3 (earth + glass + metal + checker
// ground), 8 = config C4 (grid x grid random spheres on the curved ground).  `grid` = half-extent
// of the a,b loops of scene 1 (reference 3) / scene 8.  `asset` = earth image path for 4 and 7.
SceneSetup makeScene(int scene, int grid, uint64_t seed, const std::string &asset);

bool writePpm(const std::string &path, const uint8_t *rgb8, uint32_t width, uint32_t height);
// 8-bit RGB PNG (what the reference writes: image.writeToFilePath("out.png"), src/main.zig:405), on zlib
bool writePng(const std::string &path, const uint8_t *rgb8, uint32_t width, uint32_t height);
bool decodePng(const std::string &path, Image &out, std::string &err);

}  // namespace rtw_host
