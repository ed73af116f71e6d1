// oracle_capi.cpp — C entry points of the CPU ORACLE (test infrastructure only; see the header
// of rtw_oracle.hpp for who may call this and for the "parity unpinned" statement).
//
// Restates the application layer of the reference: the six scene builders
// (src/main.zig:124-293), the per-scene camera/image settings (src/main.zig:304-362), the render
// loop and resolve (src/main.zig:382-402).  Struct layouts for scenes/cameras are the ones of
// include/rtw_cuda.h so tests can hand the same bytes to the oracle and to the CUDA library.
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rtw_cuda.h"
#include "rtw_oracle.hpp"

using namespace orc;

namespace {

struct SceneConfig {  // src/main.zig:304-362
    double look_from[3], look_at[3], vfov, aperture, aspect, background[3];
    uint32_t width, height, spp, max_depth;
};

// ---- a BVH over the top-level objects, used only where a linear scan is infeasible (C4) ----
struct CpuBvh {
    struct Node { Aabb box; int left = -1, right = -1, first = 0, count = 0; };
    std::vector<Node> nodes;
    std::vector<int> order;  // object indices
    bool built = false;
};

struct OrcScene {
    Scene scene;
    Rng rng_after_build{42};  // stream position after scene generation (main.zig:300-301,321)
    SceneConfig cfg{};
    CpuBvh bvh;
    // export storage
    std::vector<rtw_prim> x_prims;
    std::vector<rtw_xform> x_xforms;
    std::vector<rtw_material> x_mats;
    std::vector<rtw_texture> x_texs;
    std::vector<rtw_image> x_imgs;
    std::vector<rtw_perlin> x_perlins;
    std::vector<std::vector<double>> x_ranvec;
};

int add_solid(Scene &s, V3d c) {
    Texture t; t.kind = Texture::SOLID; t.color = c; s.textures.push_back(t);
    return (int)s.textures.size() - 1;
}
int add_checker(Scene &s, V3d odd, V3d even) {  // Texture.makeChecker texture.zig:20-26
    const int o = add_solid(s, odd), e = add_solid(s, even);
    Texture t; t.kind = Texture::CHECKER; t.a = o; t.b = e; s.textures.push_back(t);
    return (int)s.textures.size() - 1;
}
int add_noise(Scene &s, double scale, Rng &g) {  // Texture.makeNoise texture.zig:28-30, :89-94
    Perlin p; p.init(g); s.perlins.push_back(p);
    Texture t; t.kind = Texture::NOISE; t.a = (int)s.perlins.size() - 1; t.scale = scale; s.textures.push_back(t);
    return (int)s.textures.size() - 1;
}
int add_mat(Scene &s, Material::Kind k, int tex, V3d albedo, double param) {
    Material m; m.kind = k; m.texture = tex; m.albedo = albedo; m.param = param; s.materials.push_back(m);
    return (int)s.materials.size() - 1;
}

void set_cfg(SceneConfig &c, V3d from, V3d at, double vfov, double aperture, V3d bg) {
    c.look_from[0] = from.x; c.look_from[1] = from.y; c.look_from[2] = from.z;
    c.look_at[0] = at.x; c.look_at[1] = at.y; c.look_at[2] = at.z;
    c.vfov = vfov; c.aperture = aperture;
    c.background[0] = bg.x; c.background[1] = bg.y; c.background[2] = bg.z;
}

// generateRandomScene — src/main.zig:157-221.  `grid` = half-extent of the a,b loops
// (3 in the reference, :177-180; 11 in the book).
void build_random_scene(OrcScene &os, Rng &g, int grid) {
    Scene &s = os.scene;
    s.world = make_list();
    const int checker = add_checker(s, {0.2, 0.3, 0.1}, {0.9, 0.9, 0.9});
    const int mat_ground = add_mat(s, Material::DIFFUSE, checker, {0, 0, 0}, 0);
    const int mat1 = add_mat(s, Material::DIELECTRIC, -1, {0, 0, 0}, 1.5);
    const int mat2 = add_mat(s, Material::DIFFUSE, add_solid(s, {0.4, 0.2, 0.1}), {0, 0, 0}, 0);
    const int mat3 = add_mat(s, Material::METAL, -1, {0.7, 0.6, 0.5}, 0.0);
    s.world.objects.push_back(make_sphere({0, -1000, 0}, 1000, mat_ground));
    s.world.objects.push_back(make_sphere({0, 1, 0}, 1.0, mat1));
    s.world.objects.push_back(make_sphere({-4, 1, 0}, 1.0, mat2));
    s.world.objects.push_back(make_sphere({4, 1, 0}, 1.0, mat3));
    for (int a = -grid; a < grid; ++a)
        for (int b = -grid; b < grid; ++b) {
            const double choose_mat = g.real01();  // draw order :181-186
            const double cx = (double)a + 0.9 * g.real01();
            const double cz = (double)b + 0.9 * g.real01();
            const V3d center{cx, 0.2, cz};
            if (center.sub({4, 0.2, 0}).norm() <= 0.9) continue;  // :188-190
            if (choose_mat < 0.8) {  // :193-205 moving diffuse
                const V3d r1{g.real01(), g.real01(), g.real01()};
                const V3d r2{g.real01(), g.real01(), g.real01()};
                const int m = add_mat(s, Material::DIFFUSE, add_solid(s, r1.mulv(r2)), {0, 0, 0}, 0);
                const V3d c1 = center.add({0, g.real(0, 0.5), 0});
                s.world.objects.push_back(make_moving_sphere(center, c1, 0, 1, 0.2, m));
            } else if (choose_mat < 0.95) {  // :206-211 metal
                const V3d albedo{g.real(0.5, 1), g.real(0.5, 1), g.real(0.5, 1)};
                const double fuzz = g.real(0, 0.5);
                s.world.objects.push_back(make_sphere(center, 0.2, add_mat(s, Material::METAL, -1, albedo, fuzz)));
            } else {  // :212-216 glass
                s.world.objects.push_back(make_sphere(center, 0.2, add_mat(s, Material::DIELECTRIC, -1, {0, 0, 0}, 1.5)));
            }
        }
    set_cfg(os.cfg, {13, 2, 3}, {0, 0, 0}, 20.0, 0.1, {0.70, 0.80, 1.00});  // :320-326
}

void build_two_spheres(OrcScene &os) {  // main.zig:124-139, :327-332
    Scene &s = os.scene;
    s.world = make_list();
    const int checker = add_checker(s, {0.2, 0.3, 0.1}, {0.9, 0.9, 0.9});
    const int m1 = add_mat(s, Material::DIFFUSE, checker, {0, 0, 0}, 0);
    const int m2 = add_mat(s, Material::DIFFUSE, checker, {0, 0, 0}, 0);
    s.world.objects.push_back(make_sphere({0, -10, 0}, 10, m1));
    s.world.objects.push_back(make_sphere({0, 10, 0}, 10, m2));
    set_cfg(os.cfg, {13, 2, 3}, {0, 0, 0}, 20.0, 0.0, {0.70, 0.80, 1.00});
}

void build_two_perlin_spheres(OrcScene &os, Rng &g) {  // main.zig:141-155, :333-338
    Scene &s = os.scene;
    s.world = make_list();
    const int perlin = add_noise(s, 4.0, g);
    const int m1 = add_mat(s, Material::DIFFUSE, perlin, {0, 0, 0}, 0);
    const int m2 = add_mat(s, Material::DIFFUSE, perlin, {0, 0, 0}, 0);
    s.world.objects.push_back(make_sphere({0, -1000, 0}, 1000, m1));
    s.world.objects.push_back(make_sphere({0, 2, 0}, 2, m2));
    set_cfg(os.cfg, {13, 2, 3}, {0, 0, 0}, 20.0, 0.0, {0.70, 0.80, 1.00});
}

void build_earth(OrcScene &os, const uint8_t *rgba, uint32_t w, uint32_t h) {  // main.zig:223-234, :339-344
    Scene &s = os.scene;
    s.world = make_list();
    Image im; im.w = w; im.h = h; im.rgba.assign(rgba, rgba + (size_t)w * h * 4);
    s.images.push_back(std::move(im));
    Texture t; t.kind = Texture::IMAGE; t.a = 0; s.textures.push_back(t);
    const int m = add_mat(s, Material::DIFFUSE, (int)s.textures.size() - 1, {0, 0, 0}, 0);
    s.world.objects.push_back(make_sphere({0, 0, 0}, 2, m));
    set_cfg(os.cfg, {13, 2, 3}, {0, 0, 0}, 20.0, 0.0, {0.70, 0.80, 1.00});
}

void build_simple_light(OrcScene &os, Rng &g) {  // main.zig:236-257, :345-351
    Scene &s = os.scene;
    s.world = make_list();
    const int perlin = add_noise(s, 4.0, g);
    const int m1 = add_mat(s, Material::DIFFUSE, perlin, {0, 0, 0}, 0);
    const int m2 = add_mat(s, Material::DIFFUSE, perlin, {0, 0, 0}, 0);
    s.world.objects.push_back(make_sphere({0, -1000, 0}, 1000, m1));
    s.world.objects.push_back(make_sphere({0, 2, 0}, 2, m2));
    const int m3 = add_mat(s, Material::DIFFUSE_LIGHT, add_solid(s, {4, 4, 4}), {0, 0, 0}, 0);
    s.world.objects.push_back(make_rect(Hittable::XY_RECT, 3.0, 5.0, 1.0, 3.0, -2.0, m3));
    set_cfg(os.cfg, {26, 3, 6}, {0, 2, 0}, 20.0, 0.0, {0, 0, 0});
    os.cfg.spp = 400;
}

void build_cornell(OrcScene &os) {  // main.zig:259-293, :352-361
    Scene &s = os.scene;
    s.world = make_list();
    const int red = add_mat(s, Material::DIFFUSE, add_solid(s, {0.65, 0.05, 0.05}), {0, 0, 0}, 0);
    const int white = add_mat(s, Material::DIFFUSE, add_solid(s, {0.73, 0.73, 0.73}), {0, 0, 0}, 0);
    const int green = add_mat(s, Material::DIFFUSE, add_solid(s, {0.12, 0.45, 0.15}), {0, 0, 0}, 0);
    const int light = add_mat(s, Material::DIFFUSE_LIGHT, add_solid(s, {15, 15, 15}), {0, 0, 0}, 0);
    auto &o = s.world.objects;
    o.push_back(make_rect(Hittable::YZ_RECT, 0, 555, 0, 555, 555, green));
    o.push_back(make_rect(Hittable::YZ_RECT, 0, 555, 0, 555, 0, red));
    o.push_back(make_rect(Hittable::XZ_RECT, 213, 343, 227, 332, 554, light));
    o.push_back(make_rect(Hittable::XZ_RECT, 0, 555, 0, 555, 0, white));
    o.push_back(make_rect(Hittable::XZ_RECT, 0, 555, 0, 555, 555, white));
    o.push_back(make_rect(Hittable::XY_RECT, 0, 555, 0, 555, 555, white));
    const double pi = 3.14159265358979323846;
    o.push_back(make_translate(make_rotate_y(make_box({0, 0, 0}, {165, 330, 165}, white), 15.0 * pi / 180.0),
                               {265, 0, 295}));
    o.push_back(make_translate(make_rotate_y(make_box({0, 0, 0}, {165, 165, 165}, white), -18.0 * pi / 180.0),
                               {130, 0, 65}));
    set_cfg(os.cfg, {278, 278, -800}, {278, 278, 0}, 40.0, 0.0, {0, 0, 0});
    os.cfg.aspect = 1.0; os.cfg.width = 600; os.cfg.height = 600; os.cfg.spp = 200;
}

// ---- flatten: nested graph -> the POD arrays of include/rtw_cuda.h ---------------------------
void flatten(OrcScene &os, const Hittable &h, int chain) {
    switch (h.kind) {
        case Hittable::LIST:
        case Hittable::BOX:
            for (const auto &o : h.objects) flatten(os, o, chain);
            return;
        case Hittable::TRANSLATE: {
            rtw_xform x{}; x.kind = RTW_XFORM_TRANSLATE; x.outer = chain;
            x.v[0] = h.offset.x; x.v[1] = h.offset.y; x.v[2] = h.offset.z;
            os.x_xforms.push_back(x);
            flatten(os, *h.object, (int)os.x_xforms.size() - 1);
            return;
        }
        case Hittable::ROTATE_Y: {
            rtw_xform x{}; x.kind = RTW_XFORM_ROTATE_Y; x.outer = chain;
            x.v[0] = h.sin_t; x.v[1] = h.cos_t;
            os.x_xforms.push_back(x);
            flatten(os, *h.object, (int)os.x_xforms.size() - 1);
            return;
        }
        default: break;
    }
    rtw_prim p{};
    p.material = (uint32_t)h.material;
    p.xform = chain;
    switch (h.kind) {
        case Hittable::SPHERE:
            p.kind = RTW_PRIM_SPHERE;
            p.v[0] = h.center.x; p.v[1] = h.center.y; p.v[2] = h.center.z; p.v[3] = h.radius;
            break;
        case Hittable::MOVING_SPHERE:
            p.kind = RTW_PRIM_MOVING_SPHERE;
            p.v[0] = h.center.x; p.v[1] = h.center.y; p.v[2] = h.center.z;
            p.v[3] = h.center1.x; p.v[4] = h.center1.y; p.v[5] = h.center1.z;
            p.v[6] = h.time0; p.v[7] = h.time1; p.v[8] = h.radius;
            break;
        default:
            p.kind = h.kind == Hittable::XY_RECT ? RTW_PRIM_XY_RECT
                     : h.kind == Hittable::XZ_RECT ? RTW_PRIM_XZ_RECT : RTW_PRIM_YZ_RECT;
            p.v[0] = h.a0; p.v[1] = h.a1; p.v[2] = h.b0; p.v[3] = h.b1; p.v[4] = h.k;
    }
    os.x_prims.push_back(p);
}

void export_tables(OrcScene &os) {
    os.x_prims.clear(); os.x_xforms.clear(); os.x_mats.clear(); os.x_texs.clear();
    os.x_imgs.clear(); os.x_perlins.clear(); os.x_ranvec.clear();
    flatten(os, os.scene.world, -1);
    for (const auto &m : os.scene.materials) {
        rtw_material x{}; x.kind = (uint32_t)m.kind; x.texture = m.texture;
        x.albedo[0] = m.albedo.x; x.albedo[1] = m.albedo.y; x.albedo[2] = m.albedo.z; x.param = m.param;
        os.x_mats.push_back(x);
    }
    for (const auto &t : os.scene.textures) {
        rtw_texture x{}; x.kind = (uint32_t)t.kind; x.a = t.a; x.b = t.b;
        x.color[0] = t.color.x; x.color[1] = t.color.y; x.color[2] = t.color.z; x.scale = t.scale;
        os.x_texs.push_back(x);
    }
    for (const auto &im : os.scene.images) os.x_imgs.push_back({im.w, im.h, im.rgba.data()});
    for (const auto &p : os.scene.perlins) {
        os.x_ranvec.emplace_back(&p.ranvec[0][0], &p.ranvec[0][0] + 768);
    }
    for (size_t i = 0; i < os.scene.perlins.size(); ++i) {
        const auto &p = os.scene.perlins[i];
        os.x_perlins.push_back({os.x_ranvec[i].data(), p.perm[0], p.perm[1], p.perm[2]});
    }
}

Hittable leaf_from_prim(const rtw_prim &p, uint32_t id) {
    Hittable h;
    switch (p.kind) {
        case RTW_PRIM_SPHERE: h = make_sphere({p.v[0], p.v[1], p.v[2]}, p.v[3], (int)p.material); break;
        case RTW_PRIM_MOVING_SPHERE:
            h = make_moving_sphere({p.v[0], p.v[1], p.v[2]}, {p.v[3], p.v[4], p.v[5]}, p.v[6], p.v[7], p.v[8],
                                   (int)p.material);
            break;
        case RTW_PRIM_XY_RECT: h = make_rect(Hittable::XY_RECT, p.v[0], p.v[1], p.v[2], p.v[3], p.v[4], (int)p.material); break;
        case RTW_PRIM_XZ_RECT: h = make_rect(Hittable::XZ_RECT, p.v[0], p.v[1], p.v[2], p.v[3], p.v[4], (int)p.material); break;
        default: h = make_rect(Hittable::YZ_RECT, p.v[0], p.v[1], p.v[2], p.v[3], p.v[4], (int)p.material); break;
    }
    h.prim_id = id;
    return h;
}

Camera camera_from(const rtw_camera *c) {
    Camera k;
    auto v = [](const double *p) { return V3d{p[0], p[1], p[2]}; };
    k.origin = v(c->origin); k.horizontal = v(c->horizontal); k.vertical = v(c->vertical);
    k.lower_left_corner = v(c->lower_left_corner); k.u = v(c->u); k.v = v(c->v); k.w = v(c->w);
    k.lens_radius = c->lens_radius; k.time0 = c->time0; k.time1 = c->time1;
    return k;
}
void camera_to(const Camera &k, rtw_camera *c) {
    auto w = [](double *p, V3d v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; };
    w(c->origin, k.origin); w(c->horizontal, k.horizontal); w(c->vertical, k.vertical);
    w(c->lower_left_corner, k.lower_left_corner); w(c->u, k.u); w(c->v, k.v); w(c->w, k.w);
    c->lens_radius = k.lens_radius; c->time0 = k.time0; c->time1 = k.time1;
}

// ---- CPU BVH (median split over top-level objects; conservative slab test) -------------------
void bvh_build(OrcScene &os) {
    CpuBvh &b = os.bvh;
    const auto &objs = os.scene.world.objects;
    const int n = (int)objs.size();
    std::vector<Aabb> boxes(n);
    for (int i = 0; i < n; ++i) bounding_box(objs[i], os.scene.time0, os.scene.time1, boxes[i]);
    b.order.resize(n);
    std::iota(b.order.begin(), b.order.end(), 0);
    b.nodes.clear();
    b.nodes.reserve(2 * n);
    struct Item { int node, lo, hi; };
    std::vector<Item> stack;
    b.nodes.emplace_back();
    stack.push_back({0, 0, n});
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        Aabb bb = boxes[b.order[it.lo]];
        for (int i = it.lo + 1; i < it.hi; ++i) bb = Aabb::surrounding(bb, boxes[b.order[i]]);
        b.nodes[it.node].box = bb;
        if (it.hi - it.lo <= 4) { b.nodes[it.node].first = it.lo; b.nodes[it.node].count = it.hi - it.lo; continue; }
        const double ex = bb.max.x - bb.min.x, ey = bb.max.y - bb.min.y, ez = bb.max.z - bb.min.z;
        const int ax = ex > ey ? (ex > ez ? 0 : 2) : (ey > ez ? 1 : 2);
        auto cen = [&](int i) {
            const Aabb &x = boxes[i];
            return ax == 0 ? x.min.x + x.max.x : ax == 1 ? x.min.y + x.max.y : x.min.z + x.max.z;
        };
        const int mid = (it.lo + it.hi) / 2;
        std::nth_element(b.order.begin() + it.lo, b.order.begin() + mid, b.order.begin() + it.hi,
                         [&](int x, int y) { return cen(x) < cen(y); });
        const int l = (int)b.nodes.size(); b.nodes.emplace_back();
        const int r = (int)b.nodes.size(); b.nodes.emplace_back();
        b.nodes[it.node].left = l; b.nodes[it.node].right = r;
        stack.push_back({l, it.lo, mid});
        stack.push_back({r, mid, it.hi});
    }
    b.built = true;
}

template <class R>
bool slab_conservative(const Aabb &bx, const Ray<R> &r, R t_min, R t_max) {
    // Same test as aabb.zig:8-45 but widened by a few ulps so rounding can never cull a box
    // that contains a hit the linear scan would accept.
    const R mn[3] = {R(bx.min.x), R(bx.min.y), R(bx.min.z)}, mx[3] = {R(bx.max.x), R(bx.max.y), R(bx.max.z)};
    const R o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    const R eps = std::numeric_limits<R>::epsilon() * R(8);
    R lo = t_min, hi = t_max;
    for (int a = 0; a < 3; ++a) {
        R s0 = (mn[a] - o[a]) / d[a], s1 = (mx[a] - o[a]) / d[a];
        if (s0 != s0 || s1 != s1) continue;  // 0/0: origin on the slab plane, parallel ray
        R t0 = std::fmin(s0, s1), t1 = std::fmax(s0, s1);
        t0 -= std::fabs(t0) * eps; t1 += std::fabs(t1) * eps;
        lo = std::fmax(t0, lo); hi = std::fmin(t1, hi);
        if (hi < lo) return false;
    }
    return true;
}

template <class R>
bool hit_bvh(const OrcScene &os, const Ray<R> &r, R t_min, R t_max, HitRecord<R> &rec) {
    const CpuBvh &b = os.bvh;
    const auto &objs = os.scene.world.objects;
    bool any = false;
    R closest = t_max;
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const CpuBvh::Node &n = b.nodes[stack[--sp]];
        if (!slab_conservative<R>(n.box, r, t_min, closest)) continue;
        if (n.count) {
            for (int i = 0; i < n.count; ++i) {
                HitRecord<R> tmp;
                // t_max inclusive (hittable.zig:110); equal t: the later list element wins
                if (hit<R>(objs[b.order[n.first + i]], r, t_min, closest, tmp) &&
                    (!any || tmp.t < closest || tmp.prim_id > rec.prim_id)) {
                    any = true; closest = tmp.t; rec = tmp;
                }
            }
        } else {
            stack[sp++] = n.left;
            stack[sp++] = n.right;
        }
    }
    return any;
}

template <class R>
void trace_rays_t(OrcScene &os, uint32_t n, const double *rays, int use_bvh, uint32_t *ids, double *t,
                  double *normal, double *uv) {
    if (use_bvh && !os.bvh.built) bvh_build(os);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const double *q = rays + 7 * i;
        Ray<R> r{{R(q[0]), R(q[1]), R(q[2])}, {R(q[3]), R(q[4]), R(q[5])}, R(q[6])};
        HitRecord<R> rec;
        const R inf = std::numeric_limits<R>::infinity();
        const bool h = use_bvh ? hit_bvh<R>(os, r, R(0.001), inf, rec) : hit<R>(os.scene.world, r, R(0.001), inf, rec);
        ids[i] = h ? rec.prim_id : 0xFFFFFFFFu;
        if (t) t[i] = h ? (double)rec.t : 0.0;
        if (normal) {
            normal[3 * i + 0] = h ? (double)rec.normal.x : 0.0;
            normal[3 * i + 1] = h ? (double)rec.normal.y : 0.0;
            normal[3 * i + 2] = h ? (double)rec.normal.z : 0.0;
        }
        if (uv) { uv[2 * i] = h ? (double)rec.u : 0.0; uv[2 * i + 1] = h ? (double)rec.v : 0.0; }
    }
}

template <class R>
void primary_hits_t(OrcScene &os, const Camera &cam, uint32_t W, uint32_t H, int use_bvh, uint32_t *ids,
                    double *t, double *normal) {
    if (use_bvh && !os.bvh.built) bvh_build(os);
    const R time = R(cam.time0) + R(0.5) * (R(cam.time1) - R(cam.time0));  // rand.zig:18-20 with xi=.5
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t j = 0; j < (int64_t)H; ++j)
        for (uint32_t i = 0; i < W; ++i) {
            const R u = (R(i) + R(0.5)) / (R(W) - R(1));  // main.zig:390-391 with xi=.5
            const R v = (R((uint32_t)j) + R(0.5)) / (R(H) - R(1));
            const Ray<R> r = cam.ray_from<R>(V3<R>{0, 0, 0}, u, v, time);
            HitRecord<R> rec;
            const R inf = std::numeric_limits<R>::infinity();
            const bool h = use_bvh ? hit_bvh<R>(os, r, R(0.001), inf, rec)
                                   : hit<R>(os.scene.world, r, R(0.001), inf, rec);
            const size_t k = (size_t)j * W + i;
            ids[k] = h ? rec.prim_id : 0xFFFFFFFFu;
            if (t) t[k] = h ? (double)rec.t : 0.0;
            if (normal) {
                normal[3 * k + 0] = h ? (double)rec.normal.x : 0.0;
                normal[3 * k + 1] = h ? (double)rec.normal.y : 0.0;
                normal[3 * k + 2] = h ? (double)rec.normal.z : 0.0;
            }
        }
}

// The render loop src/main.zig:382-394.  nthreads == 1: one sequential stream exactly as the
// reference (scan order j, i, s).  nthreads > 1: rows in parallel, one substream per row — NOT
// reference behaviour (the reference is single-threaded), offered as an all-core CPU baseline.
template <class R>
double render_t(OrcScene &os, const Camera &cam, const rtw_render_params *p, int nthreads, int continue_stream,
                double *accum, uint64_t *paths, uint64_t *rays) {
    const uint32_t W = p->width, H = p->height;
    const uint32_t spp = p->spp_end - p->spp_begin;
    const V3<R> bg{R(p->background[0]), R(p->background[1]), R(p->background[2])};
    uint64_t tot_paths = 0, tot_rays = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto do_row = [&](uint32_t j, Rng &g, Counters &cn) {
        for (uint32_t i = 0; i < W; ++i) {
            V3<R> px{0, 0, 0};
            for (uint32_t s = 0; s < spp; ++s) {
                const R u = (R(i) + R(g.real01())) / (R(W) - R(1));
                const R v = (R(j) + R(g.real01())) / (R(H) - R(1));
                const Ray<R> r = cam.get_ray<R>(g, u, v);
                ++cn.paths;
                px = px.add(os.scene.ray_color<R>(r, bg, g, p->max_depth, cn));
            }
            double *a = accum + ((size_t)j * W + i) * 3;
            a[0] = (double)px.x; a[1] = (double)px.y; a[2] = (double)px.z;
        }
    };
    if (nthreads <= 1) {
        Rng g = continue_stream ? os.rng_after_build : Rng(p->seed);
        Counters cn;
        for (uint32_t j = 0; j < H; ++j) do_row(j, g, cn);
        tot_paths = cn.paths; tot_rays = cn.rays;
    } else {
#ifdef _OPENMP
        omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_paths, tot_rays)
        for (int64_t j = 0; j < (int64_t)H; ++j) {
            Rng g(p->seed * 0x9E3779B97F4A7C15ull + (uint64_t)j * 0xD1B54A32D192ED03ull + 1);
            Counters cn;
            do_row((uint32_t)j, g, cn);
            tot_paths += cn.paths; tot_rays += cn.rays;
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (paths) *paths = tot_paths;
    if (rays) *rays = tot_rays;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

// ---- scene lifetime ---------------------------------------------------------------------------
// scene_id 1..6 = the reference's `scene` constant (src/main.zig:310,320-362).  `grid` is the
// half-extent of scene 1's loops (reference: 3).  image_* only for scene 4.
void *orc_scene_builtin(int scene_id, int grid, uint64_t seed, const uint8_t *image_rgba, uint32_t image_w,
                        uint32_t image_h) {
    auto *os = new OrcScene();
    os->cfg.aspect = 3.0 / 2.0;  // main.zig:304-308
    os->cfg.width = 600;
    os->cfg.spp = 50;
    os->cfg.max_depth = 50;
    os->cfg.height = 0;
    Rng g(seed);  // main.zig:300
    switch (scene_id) {
        case 1: build_random_scene(*os, g, grid); break;
        case 2: build_two_spheres(*os); break;
        case 3: build_two_perlin_spheres(*os, g); break;
        case 4:
            if (!image_rgba) { delete os; return nullptr; }
            build_earth(*os, image_rgba, image_w, image_h);
            break;
        case 5: build_simple_light(*os, g); break;
        case 6: build_cornell(*os); break;
        default: delete os; return nullptr;
    }
    if (os->cfg.height == 0)  // main.zig:306 @divTrunc(width, aspect)
        os->cfg.height = (uint32_t)std::trunc((double)os->cfg.width / os->cfg.aspect);
    os->rng_after_build = g;
    os->scene.time0 = 0; os->scene.time1 = 1;  // main.zig:374-375
    os->scene.assign_ids();
    export_tables(*os);
    return os;
}

void *orc_scene_from_desc(const rtw_scene_desc *d) {
    auto *os = new OrcScene();
    Scene &s = os->scene;
    s.world = make_list();
    for (uint32_t i = 0; i < d->n_prims; ++i) {
        Hittable h = leaf_from_prim(d->prims[i], i);
        int x = d->prims[i].xform;
        while (x >= 0) {
            const rtw_xform &xf = d->xforms[x];
            if (xf.kind == RTW_XFORM_TRANSLATE) h = make_translate(std::move(h), {xf.v[0], xf.v[1], xf.v[2]});
            else h = make_rotate_y_sc(std::move(h), xf.v[0], xf.v[1]);
            x = xf.outer;
        }
        s.world.objects.push_back(std::move(h));
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const auto &m = d->materials[i];
        add_mat(s, (Material::Kind)m.kind, m.texture, {m.albedo[0], m.albedo[1], m.albedo[2]}, m.param);
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const auto &t = d->textures[i];
        Texture x; x.kind = (Texture::Kind)t.kind; x.a = t.a; x.b = t.b;
        x.color = {t.color[0], t.color[1], t.color[2]}; x.scale = t.scale;
        s.textures.push_back(x);
    }
    for (uint32_t i = 0; i < d->n_images; ++i) {
        Image im; im.w = d->images[i].width; im.h = d->images[i].height;
        im.rgba.assign(d->images[i].rgba8, d->images[i].rgba8 + (size_t)im.w * im.h * 4);
        s.images.push_back(std::move(im));
    }
    for (uint32_t i = 0; i < d->n_perlins; ++i) {
        Perlin p;
        std::memcpy(p.ranvec, d->perlins[i].ranvec, sizeof(p.ranvec));
        std::memcpy(p.perm[0], d->perlins[i].perm_x, 1024);
        std::memcpy(p.perm[1], d->perlins[i].perm_y, 1024);
        std::memcpy(p.perm[2], d->perlins[i].perm_z, 1024);
        s.perlins.push_back(p);
    }
    s.time0 = d->time0; s.time1 = d->time1;
    s.n_leaves = d->n_prims;
    os->rng_after_build = Rng(42);
    export_tables(*os);
    return os;
}

void orc_scene_free(void *h) { delete (OrcScene *)h; }

// Pointers stay valid until orc_scene_free.
void orc_scene_export(void *h, rtw_scene_desc *out) {
    auto *os = (OrcScene *)h;
    out->n_prims = (uint32_t)os->x_prims.size(); out->prims = os->x_prims.data();
    out->n_xforms = (uint32_t)os->x_xforms.size(); out->xforms = os->x_xforms.data();
    out->n_materials = (uint32_t)os->x_mats.size(); out->materials = os->x_mats.data();
    out->n_textures = (uint32_t)os->x_texs.size(); out->textures = os->x_texs.data();
    out->n_images = (uint32_t)os->x_imgs.size(); out->images = os->x_imgs.data();
    out->n_perlins = (uint32_t)os->x_perlins.size(); out->perlins = os->x_perlins.data();
    out->time0 = os->scene.time0; out->time1 = os->scene.time1;
}

// look_from[3], look_at[3], vfov, aperture, aspect, background[3] | width,height,spp,max_depth
void orc_scene_config(void *h, double *f10, uint32_t *u4) {
    auto *os = (OrcScene *)h;
    const SceneConfig &c = os->cfg;
    for (int i = 0; i < 3; ++i) { f10[i] = c.look_from[i]; f10[3 + i] = c.look_at[i]; f10[9 + i] = c.background[i]; }
    f10[6] = c.vfov; f10[7] = c.aperture; f10[8] = c.aspect;
    u4[0] = c.width; u4[1] = c.height; u4[2] = c.spp; u4[3] = c.max_depth;
}

uint64_t orc_scene_draws_used(void *h) { return ((OrcScene *)h)->rng_after_build.draws; }

// Camera.init src/main.zig:52-89
void orc_camera_init(const double *look_from, const double *look_at, const double *vup, double vfov,
                     double aspect, double aperture, double focus_dist, double time0, double time1,
                     rtw_camera *out) {
    camera_to(Camera::init({look_from[0], look_from[1], look_from[2]}, {look_at[0], look_at[1], look_at[2]},
                           {vup[0], vup[1], vup[2]}, vfov, aspect, aperture, focus_dist, time0, time1),
              out);
}

// ---- queries ----------------------------------------------------------------------------------
int orc_trace_rays(void *h, uint32_t n, const double *rays, int precision, int use_bvh, uint32_t *ids,
                   double *t, double *normal, double *uv) {
    auto *os = (OrcScene *)h;
    if (precision == 64) trace_rays_t<double>(*os, n, rays, use_bvh, ids, t, normal, uv);
    else if (precision == 32) trace_rays_t<float>(*os, n, rays, use_bvh, ids, t, normal, uv);
    else return 1;
    return 0;
}

int orc_primary_hits(void *h, const rtw_camera *cam, uint32_t W, uint32_t H, int precision, int use_bvh,
                     uint32_t *ids, double *t, double *normal) {
    auto *os = (OrcScene *)h;
    const Camera c = camera_from(cam);
    if (precision == 64) primary_hits_t<double>(*os, c, W, H, use_bvh, ids, t, normal);
    else if (precision == 32) primary_hits_t<float>(*os, c, W, H, use_bvh, ids, t, normal);
    else return 1;
    return 0;
}

// accum: H*W*3 doubles (sums), row j = scanline j (bottom first).  rgb8 (optional): H*W*3 bytes,
// row 0 = top (main.zig:396).  Returns seconds spent in the loop nest.
double orc_render(void *h, const rtw_camera *cam, const rtw_render_params *p, int precision, int nthreads,
                  int continue_stream, double *accum, uint8_t *rgb8, uint64_t *paths, uint64_t *rays) {
    auto *os = (OrcScene *)h;
    const Camera c = camera_from(cam);
    const double secs = precision == 32 ? render_t<float>(*os, c, p, nthreads, continue_stream, accum, paths, rays)
                                        : render_t<double>(*os, c, p, nthreads, continue_stream, accum, paths, rays);
    if (rgb8) {
        const uint32_t W = p->width, H = p->height;
        const double scale = 1.0 / (double)(p->spp_total ? p->spp_total : p->spp_end - p->spp_begin);
        for (uint32_t j = 0; j < H; ++j)
            for (uint32_t i = 0; i < W; ++i)
                for (int ch = 0; ch < 3; ++ch)
                    rgb8[((size_t)(H - 1 - j) * W + i) * 3 + ch] =
                        resolve_channel(accum[((size_t)j * W + i) * 3 + ch], scale);
    }
    return secs;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- known-answer-test hooks --------------------------------------------
void orc_kat_xoshiro(uint64_t seed, int n, uint64_t *out, uint64_t *state4) {
    Xoshiro256pp g(seed);
    if (state4) std::memcpy(state4, g.s, 32);
    for (int i = 0; i < n; ++i) out[i] = g.next();
}
void orc_kat_real01(uint64_t seed, int n, double *out) {
    Rng g(seed);
    for (int i = 0; i < n; ++i) out[i] = g.real01();
}
void orc_kat_uint_less_than(uint64_t seed, uint64_t bound, int n, uint64_t *out) {
    Rng g(seed);
    for (int i = 0; i < n; ++i) out[i] = g.uint_less_than(bound);
}
void orc_kat_sphere_uv(const double *p, double *uv) { sphere_uv<double>({p[0], p[1], p[2]}, uv[0], uv[1]); }
double orc_kat_reflectance(double c, double idx) { return reflectance<double>(c, idx); }
void orc_kat_reflect(const double *v, const double *n, double *o) {
    const V3d r = reflect<double>({v[0], v[1], v[2]}, {n[0], n[1], n[2]});
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void orc_kat_refract(const double *v, const double *n, double eta, double *o) {
    const V3d r = refract<double>({v[0], v[1], v[2]}, {n[0], n[1], n[2]}, eta);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
uint8_t orc_kat_resolve(double sum, uint32_t spp) { return resolve_channel(sum, 1.0 / (double)spp); }
void orc_kat_texture(void *h, int tex, double u, double v, const double *p, double *rgb) {
    const V3d c = ((OrcScene *)h)->scene.texture_value<double>(tex, u, v, {p[0], p[1], p[2]});
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
double orc_kat_perlin_turb(void *h, int perlin, const double *p, int depth) {
    return ((OrcScene *)h)->scene.perlins[perlin].turb<double>({p[0], p[1], p[2]}, depth);
}
// bounding box of top-level object i (reference boudingBox rules); returns 0/1
int orc_kat_bounding_box(void *h, uint32_t i, double *min3, double *max3) {
    auto *os = (OrcScene *)h;
    Aabb b;
    if (i >= os->scene.world.objects.size()) return 0;
    if (!bounding_box(os->scene.world.objects[i], os->scene.time0, os->scene.time1, b)) return 0;
    min3[0] = b.min.x; min3[1] = b.min.y; min3[2] = b.min.z;
    max3[0] = b.max.x; max3[1] = b.max.y; max3[2] = b.max.z;
    return 1;
}
int orc_kat_aabb_hit(const double *min3, const double *max3, const double *ray7, double t_min, double t_max) {
    Aabb b{{min3[0], min3[1], min3[2]}, {max3[0], max3[1], max3[2]}};
    Ray<double> r{{ray7[0], ray7[1], ray7[2]}, {ray7[3], ray7[4], ray7[5]}, ray7[6]};
    return b.hit<double>(r, t_min, t_max) ? 1 : 0;
}
// full closest-hit record of one ray (f64): t, p[3], n[3], u, v, front, prim, material
int orc_kat_hit_record(void *h, const double *ray7, double *out12) {
    auto *os = (OrcScene *)h;
    Ray<double> r{{ray7[0], ray7[1], ray7[2]}, {ray7[3], ray7[4], ray7[5]}, ray7[6]};
    HitRecord<double> rec;
    if (!hit<double>(os->scene.world, r, 0.001, std::numeric_limits<double>::infinity(), rec)) return 0;
    out12[0] = rec.t; out12[1] = rec.p.x; out12[2] = rec.p.y; out12[3] = rec.p.z;
    out12[4] = rec.normal.x; out12[5] = rec.normal.y; out12[6] = rec.normal.z;
    out12[7] = rec.u; out12[8] = rec.v; out12[9] = rec.front_face ? 1.0 : 0.0;
    out12[10] = (double)rec.prim_id; out12[11] = (double)rec.material;
    return 1;
}

// ---- unit-level hooks (pinned against the transpiled reference in tests/test_ref_pin.py; the *_given forms are
//      what the device's unit probe is compared with) ---------------------------------------------------------
static Ray<double> ray_of(const double *q) { return {{q[0], q[1], q[2]}, {q[3], q[4], q[5]}, q[6]}; }
static void ray_to(const Ray<double> &r, double *q) {
    q[0] = r.o.x; q[1] = r.o.y; q[2] = r.o.z; q[3] = r.d.x; q[4] = r.d.y; q[5] = r.d.z; q[6] = r.time;
}
static void rec_to(const HitRecord<double> &rec, double *o) {
    o[0] = rec.t; o[1] = rec.p.x; o[2] = rec.p.y; o[3] = rec.p.z;
    o[4] = rec.normal.x; o[5] = rec.normal.y; o[6] = rec.normal.z;
    o[7] = rec.u; o[8] = rec.v; o[9] = rec.front_face ? 1.0 : 0.0;
    o[10] = (double)rec.prim_id; o[11] = (double)rec.material;
}
// n rays against world.hit (hittable.zig:47-59) with explicit [t_min, t_max]; out12 as orc_kat_hit_record
void orc_kat_hit_records(void *h, uint32_t n, const double *rays7, double t_min, double t_max, double *out12, uint8_t *hit_mask) {
    auto *os = (OrcScene *)h;
    for (uint32_t i = 0; i < n; ++i) {
        HitRecord<double> rec;
        const bool ok = hit<double>(os->scene.world, ray_of(rays7 + 7 * (size_t)i), t_min, t_max, rec);
        hit_mask[i] = ok ? 1 : 0;
        if (ok) rec_to(rec, out12 + 12 * (size_t)i);
    }
}
double orc_kat_perlin_noise(void *h, int perlin, const double *p) {
    return ((OrcScene *)h)->scene.perlins[perlin].noise<double>({p[0], p[1], p[2]});
}
void orc_kat_perlin_tables(void *h, int perlin, double *ranvec768, uint32_t *perm768) {
    const Perlin &pn = ((OrcScene *)h)->scene.perlins[perlin];
    std::memcpy(ranvec768, pn.ranvec, sizeof pn.ranvec);
    std::memcpy(perm768, pn.perm, sizeof pn.perm);
}
static HitRecord<double> rec_of(const double *rec10, int material) {
    HitRecord<double> rec;
    rec.p = {rec10[0], rec10[1], rec10[2]}; rec.normal = {rec10[3], rec10[4], rec10[5]};
    rec.u = rec10[6]; rec.v = rec10[7]; rec.front_face = rec10[8] != 0.0; rec.t = rec10[9];
    rec.material = material;
    return rec;
}
// Material.scatter (material.zig:22-29) + emitted (:31-38) with the reference's own sampling from Rng(seed);
// rec10 = p[3], normal[3], u, v, front_face, t.  out14 = attenuation[3], scattered ray[7], emitted[3], draws used
int orc_kat_scatter(void *h, int material, const double *ray7, const double *rec10, uint64_t seed, double *out14) {
    auto *os = (OrcScene *)h;
    const HitRecord<double> rec = rec_of(rec10, material);
    Rng g(seed);
    V3d att{0, 0, 0};
    Ray<double> sc{{0, 0, 0}, {0, 0, 0}, 0};
    const bool ok = os->scene.scatter<double>(ray_of(ray7), rec, att, sc, g);
    const V3d em = os->scene.emitted<double>(rec);
    out14[0] = att.x; out14[1] = att.y; out14[2] = att.z;
    ray_to(sc, out14 + 3);
    out14[10] = em.x; out14[11] = em.y; out14[12] = em.z;
    out14[13] = (double)g.draws;
    return ok ? 1 : 0;
}
// same with the random choices supplied: vec3 = the unit vector (diffuse) / ball point (metal), xi = the uniform (dielectric)
int orc_kat_scatter_given(void *h, int material, const double *ray7, const double *rec10, const double *vec3, double xi,
                          double *out13) {
    auto *os = (OrcScene *)h;
    const HitRecord<double> rec = rec_of(rec10, material);
    GivenSampler smp;
    smp.vec = {vec3[0], vec3[1], vec3[2]};
    smp.xi = xi;
    V3d att{0, 0, 0};
    Ray<double> sc{{0, 0, 0}, {0, 0, 0}, 0};
    const bool ok = os->scene.scatter_s<double>(ray_of(ray7), rec, att, sc, smp);
    const V3d em = os->scene.emitted<double>(rec);
    out13[0] = att.x; out13[1] = att.y; out13[2] = att.z;
    ray_to(sc, out13 + 3);
    out13[10] = em.x; out13[11] = em.y; out13[12] = em.z;
    return ok ? 1 : 0;
}
// Camera.getRay (main.zig:91-100) from Rng(seed); returns draws used
uint64_t orc_kat_get_ray(const rtw_camera *cam, uint64_t seed, double s, double t, double *ray7) {
    const Camera c = camera_from(cam);
    Rng g(seed);
    ray_to(c.get_ray<double>(g, s, t), ray7);
    return g.draws;
}
// same with the lens-disk point and the time uniform supplied
void orc_kat_get_ray_given(const rtw_camera *cam, const double *disk2, double time_xi, double s, double t, double *ray7) {
    const Camera c = camera_from(cam);
    GivenSampler smp;
    smp.vec = {disk2[0], disk2[1], 0.0};
    smp.xi = time_xi;
    ray_to(c.get_ray_s<double>(smp, s, t), ray7);
}
// rayColor (main.zig:103-122) of one ray from Rng(seed): out5 = colour[3], rays traced, draws used
void orc_kat_ray_color(void *h, const double *ray7, const double *bg3, uint32_t depth, uint64_t seed, double *out5) {
    auto *os = (OrcScene *)h;
    Rng g(seed);
    Counters cn;
    const V3d c = os->scene.ray_color<double>(ray_of(ray7), {bg3[0], bg3[1], bg3[2]}, g, depth, cn);
    out5[0] = c.x; out5[1] = c.y; out5[2] = c.z; out5[3] = (double)cn.rays; out5[4] = (double)g.draws;
}
// rand.zig:22-40 from Rng(seed): which = 0 randomPointInUnitSphere, 1 randomPointInUnitDisk, 2 randomUnitVector
void orc_kat_samplers(uint64_t seed, int which, uint32_t n, double *out3) {
    Rng g(seed);
    for (uint32_t i = 0; i < n; ++i) {
        const V3d v = which == 0 ? random_in_unit_sphere<double>(g) : which == 1 ? random_in_unit_disk<double>(g) : random_unit_vector<double>(g);
        out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
    }
}

}  // extern "C"
