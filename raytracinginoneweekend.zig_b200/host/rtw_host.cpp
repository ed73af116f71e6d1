// rtw_host.cpp — implementation of the C++ twin of the reference's host side (see rtw.hpp).
#include "rtw.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>

namespace rtw_host {

// ------------------------------------------------------------------------------------------------
// Random — Zig std DefaultPrng + float(f64) + uintLessThan
// ------------------------------------------------------------------------------------------------
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

Random::Random(uint64_t seed) {
    uint64_t z = seed;
    for (auto &w : s_) {
        z += 0x9e3779b97f4a7c15ull;
        uint64_t v = z;
        v = (v ^ (v >> 30)) * 0xbf58476d1ce4e5b9ull;
        v = (v ^ (v >> 27)) * 0x94d049bb133111ebull;
        w = v ^ (v >> 31);
    }
}
uint64_t Random::next() {
    const uint64_t result = rotl64(s_[0] + s_[3], 23) + s_[0];
    const uint64_t t = s_[1] << 17;
    s_[2] ^= s_[0]; s_[3] ^= s_[1]; s_[1] ^= s_[2]; s_[0] ^= s_[3];
    s_[2] ^= t;
    s_[3] = rotl64(s_[3], 45);
    return result;
}
double Random::real01() {
    const uint64_t r = next();
    unsigned lz = r ? (unsigned)__builtin_clzll(r) : 64u;
    if (lz >= 12) {
        lz = 12;
        while (true) {
            const uint64_t more = next();
            const unsigned a = more ? (unsigned)__builtin_clzll(more) : 64u;
            lz += a;
            if (a != 64) break;
            if (lz >= 1022) { lz = 1022; break; }
        }
    }
    const uint64_t bits = ((uint64_t)(1022 - lz) << 52) | (r & ((1ull << 52) - 1));
    double out;
    std::memcpy(&out, &bits, sizeof out);
    return out;
}
double Random::real(double lo, double hi) { return lo + real01() * (hi - lo); }
uint64_t Random::intLessThan(uint64_t bound) {
    unsigned __int128 m = (unsigned __int128)next() * bound;
    uint64_t low = (uint64_t)m;
    if (low < bound) {
        const uint64_t threshold = (0 - bound) % bound;
        while (low < threshold) {
            m = (unsigned __int128)next() * bound;
            low = (uint64_t)m;
        }
    }
    return (uint64_t)(m >> 64);
}

// ------------------------------------------------------------------------------------------------
// Perlin tables, textures
// ------------------------------------------------------------------------------------------------
Perlin::Perlin(Random &rng) {
    for (int i = 0; i < 256; ++i) {
        double v[3] = {rng.real(-1, 1), rng.real(-1, 1), rng.real(-1, 1)};
        const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int a = 0; a < 3; ++a) ranvec[i][a] = n == 0.0 ? v[a] : v[a] / n;
        for (auto &p : perm) p[i] = (uint32_t)i;
    }
    for (auto &p : perm)
        for (int i = 255; i > 0; --i) {  // exclusive upper bound: perlin.zig:96
            const uint64_t target = rng.intLessThan((uint64_t)i);
            std::swap(p[i], p[target]);
        }
}

Texture Texture::makeSolid(Color c) { return {SolidTexture{c}}; }
Texture Texture::makeChecker(Color odd, Color even) {
    return {CheckerTexture{std::make_shared<Texture>(makeSolid(odd)), std::make_shared<Texture>(makeSolid(even))}};
}
Texture Texture::makeNoise(double scale, Random &rng) { return {NoiseTexture{std::make_shared<Perlin>(rng), scale}}; }
Texture Texture::makeImageFromMemory(const uint8_t *rgba, uint32_t w, uint32_t h) {
    auto im = std::make_shared<Image>();
    im->width = w; im->height = h;
    im->rgba.assign(rgba, rgba + (size_t)w * h * 4);
    return {ImageTexture{im}};
}
Texture Texture::makeImage(const std::string &path) {
    auto im = std::make_shared<Image>();
    std::string err;
    if (!decodePng(path, *im, err)) throw std::runtime_error("makeImage(" + path + "): " + err);
    return {ImageTexture{im}};
}

// ------------------------------------------------------------------------------------------------
// PNG decode (8-bit RGB/RGBA/grey, non-interlaced) on top of zlib — the subset zigimg is asked for
// by texture.zig:111 on assets/sekaichizu.png.
// ------------------------------------------------------------------------------------------------
static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

bool decodePng(const std::string &path, Image &out, std::string &err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open file"; return false; }
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) { err = "not a PNG"; return false; }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    size_t pos = 8;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const char *tag = reinterpret_cast<const char *>(&buf[pos + 4]);
        if (pos + 12 + len > buf.size()) { err = "truncated chunk"; return false; }
        const uint8_t *data = &buf[pos + 8];
        if (!std::strncmp(tag, "IHDR", 4)) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::strncmp(tag, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::strncmp(tag, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!w || !h) { err = "no IHDR"; return false; }
    if (depth != 8 || interlace != 0 || (ctype != 6 && ctype != 2 && ctype != 0 && ctype != 4)) {
        err = "unsupported PNG flavour (need 8-bit, non-interlaced, grey/RGB/RGBA)";
        return false;
    }
    const int ch = ctype == 6 ? 4 : ctype == 2 ? 3 : ctype == 4 ? 2 : 1;
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) {
        err = "inflate failed";
        return false;
    }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *src = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)ch ? dst[x - ch] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) / 2; break;
                case 4: {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: err = "bad filter type"; return false;
            }
            dst[x] = (uint8_t)(src[x] + pred);
        }
    }
    out.width = w; out.height = h;
    out.rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const uint8_t *p = &img[i * ch];
        uint8_t *q = &out.rgba[i * 4];
        if (ch == 4) { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[3]; }
        else if (ch == 3) { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = 255; }
        else if (ch == 2) { q[0] = q[1] = q[2] = p[0]; q[3] = p[1]; }
        else { q[0] = q[1] = q[2] = p[0]; q[3] = 255; }
    }
    return true;
}

bool writePpm(const std::string &path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%u %u\n255\n", width, height);
    const size_t n = (size_t)width * height * 3;
    const bool ok = std::fwrite(rgb8, 1, n, f) == n;
    return std::fclose(f) == 0 && ok;
}

bool writePng(const std::string &path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    // filter type 0 on every row, one zlib stream, one IDAT chunk
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw[(stride + 1) * y] = 0;
        std::memcpy(&raw[(stride + 1) * y + 1], rgb8 + stride * y, stride);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto put32 = [](uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
    auto chunk = [&](const char *tag, const uint8_t *data, uint32_t len) {
        uint8_t hdr[8];
        put32(hdr, len);
        std::memcpy(hdr + 4, tag, 4);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t tail[4];
        put32(tail, (uint32_t)crc);
        return std::fwrite(hdr, 1, 8, f) == 8 && (len == 0 || std::fwrite(data, 1, len, f) == len) && std::fwrite(tail, 1, 4, f) == 4;
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t ihdr[13];
    put32(ihdr, width); put32(ihdr + 4, height);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;  // 8-bit, RGB, deflate, adaptive, no interlace
    bool ok = std::fwrite(sig, 1, 8, f) == 8 && chunk("IHDR", ihdr, 13) && chunk("IDAT", z.data(), (uint32_t)zlen) && chunk("IEND", nullptr, 0);
    return std::fclose(f) == 0 && ok;
}

// ------------------------------------------------------------------------------------------------
// Hittable constructors
// ------------------------------------------------------------------------------------------------
Hittable makeSphere(Point3 center, double radius, Rc<Material> material) { return {Sphere{center, radius, std::move(material)}}; }

Hittable Hittable::makeBox(Point3 p0, Point3 p1, Rc<Material> m) {
    auto sides = std::make_shared<HittableList>();
    auto &o = sides->objects;  // side order of Box.init, hittable.zig:437-442
    o.push_back({XyRect{p0.x, p1.x, p0.y, p1.y, p1.z, m}});
    o.push_back({XyRect{p0.x, p1.x, p0.y, p1.y, p0.z, m}});
    o.push_back({XzRect{p0.x, p1.x, p0.z, p1.z, p1.y, m}});
    o.push_back({XzRect{p0.x, p1.x, p0.z, p1.z, p0.y, m}});
    o.push_back({YzRect{p0.y, p1.y, p0.z, p1.z, p1.x, m}});
    o.push_back({YzRect{p0.y, p1.y, p0.z, p1.z, p0.x, m}});
    return {Box{p0, p1, sides}};
}
Hittable Hittable::makeTranslate(Rc<Hittable> obj, Vec3 offset) { return {Translate{std::move(obj), offset}}; }
Hittable Hittable::makeRotateY(Rc<Hittable> obj, double angle) {
    return {RotateY{std::move(obj), std::sin(angle), std::cos(angle)}};
}

// ------------------------------------------------------------------------------------------------
// Camera.init — src/main.zig:52-89
// ------------------------------------------------------------------------------------------------
namespace {
Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
Vec3 scale(Vec3 a, double t) { return {a.x * t, a.y * t, a.z * t}; }
Vec3 over(Vec3 a, double t) { return {a.x / t, a.y / t, a.z / t}; }
Vec3 cross(Vec3 u, Vec3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
double length(Vec3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
Vec3 unit(Vec3 a) { const double n = length(a); return n == 0.0 ? a : over(a, n); }
void put(double *dst, Vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
}  // namespace

Camera Camera::init(Point3 look_from, Point3 look_at, Vec3 vup, double vfov, double aspect_ratio, double aperture,
                    double focus_dist, double time0, double time1) {
    const double theta = vfov * 3.14159265358979323846 / 180.0;
    const double viewport_height = 2.0 * std::tan(theta / 2);
    const double viewport_width = aspect_ratio * viewport_height;
    const Vec3 w = unit(sub(look_from, look_at));
    const Vec3 u = unit(cross(vup, w));
    const Vec3 v = cross(w, u);
    const Vec3 horizontal = scale(u, viewport_width * focus_dist);
    const Vec3 vertical = scale(v, viewport_height * focus_dist);
    const Vec3 llc = sub(sub(sub(look_from, over(horizontal, 2.0)), over(vertical, 2.0)), scale(w, focus_dist));
    Camera cam{};
    put(cam.c.origin, look_from); put(cam.c.horizontal, horizontal); put(cam.c.vertical, vertical);
    put(cam.c.lower_left_corner, llc); put(cam.c.u, u); put(cam.c.v, v); put(cam.c.w, w);
    cam.c.lens_radius = aperture / 2.0;
    cam.c.time0 = time0;
    cam.c.time1 = time1;
    return cam;
}

// ------------------------------------------------------------------------------------------------
// flatten — walk the Rc graph depth-first; leaves get ids in append order; Translate/RotateY become
// instance nodes linked outward; materials are deduped on Rc pointer identity.
// ------------------------------------------------------------------------------------------------
namespace {
struct Flattener {
    FlatScene out;
    std::map<const Material *, uint32_t> mat_index;
    std::map<const Image *, int> image_index;
    std::map<const Perlin *, int> perlin_index;

    int texture(const Texture &t) {
        rtw_texture x{};
        if (auto *s = std::get_if<SolidTexture>(&t.v)) {
            x.kind = RTW_TEX_SOLID; x.a = x.b = -1;
            x.color[0] = s->color.x; x.color[1] = s->color.y; x.color[2] = s->color.z;
        } else if (auto *c = std::get_if<CheckerTexture>(&t.v)) {
            x.kind = RTW_TEX_CHECKER;
            x.a = texture(*c->odd);
            x.b = texture(*c->even);
        } else if (auto *n = std::get_if<NoiseTexture>(&t.v)) {
            x.kind = RTW_TEX_NOISE; x.b = -1; x.scale = n->scale;
            auto it = perlin_index.find(n->perlin.get());
            if (it == perlin_index.end()) {
                out.keep_perlins.push_back(n->perlin);
                out.ranvec_store.emplace_back(&n->perlin->ranvec[0][0], &n->perlin->ranvec[0][0] + 768);
                it = perlin_index.emplace(n->perlin.get(), (int)out.keep_perlins.size() - 1).first;
            }
            x.a = it->second;
        } else {
            const auto &im = std::get<ImageTexture>(t.v);
            x.kind = RTW_TEX_IMAGE; x.b = -1;
            auto it = image_index.find(im.image.get());
            if (it == image_index.end()) {
                out.keep_images.push_back(im.image);
                it = image_index.emplace(im.image.get(), (int)out.keep_images.size() - 1).first;
            }
            x.a = it->second;
        }
        out.textures.push_back(x);
        return (int)out.textures.size() - 1;
    }

    uint32_t material(const Rc<Material> &m) {
        auto it = mat_index.find(m.get());
        if (it != mat_index.end()) return it->second;
        rtw_material x{};
        x.texture = -1;
        if (auto *d = std::get_if<DiffuseMaterial>(&m->v)) { x.kind = RTW_MAT_DIFFUSE; x.texture = texture(d->albedo); }
        else if (auto *mt = std::get_if<MetalMaterial>(&m->v)) {
            x.kind = RTW_MAT_METAL; x.albedo[0] = mt->albedo.x; x.albedo[1] = mt->albedo.y; x.albedo[2] = mt->albedo.z; x.param = mt->fuzz;
        } else if (auto *g = std::get_if<DielectricMaterial>(&m->v)) { x.kind = RTW_MAT_DIELECTRIC; x.param = g->ir; }
        else { x.kind = RTW_MAT_DIFFUSE_LIGHT; x.texture = texture(std::get<DiffuseLightMaterial>(m->v).emit); }
        out.materials.push_back(x);
        const uint32_t idx = (uint32_t)out.materials.size() - 1;
        mat_index.emplace(m.get(), idx);
        return idx;
    }

    void rect(uint32_t kind, double a0, double a1, double b0, double b1, double k, const Rc<Material> &m, int chain) {
        rtw_prim p{};
        p.kind = kind; p.material = material(m); p.xform = chain;
        p.v[0] = a0; p.v[1] = a1; p.v[2] = b0; p.v[3] = b1; p.v[4] = k;
        out.prims.push_back(p);
    }

    void walk(const Hittable &h, int chain) {
        if (auto *s = std::get_if<Sphere>(&h.v)) {
            rtw_prim p{};
            p.kind = RTW_PRIM_SPHERE; p.material = material(s->material); p.xform = chain;
            p.v[0] = s->center.x; p.v[1] = s->center.y; p.v[2] = s->center.z; p.v[3] = s->radius;
            out.prims.push_back(p);
        } else if (auto *ms = std::get_if<MovingSphere>(&h.v)) {
            rtw_prim p{};
            p.kind = RTW_PRIM_MOVING_SPHERE; p.material = material(ms->material); p.xform = chain;
            p.v[0] = ms->center0.x; p.v[1] = ms->center0.y; p.v[2] = ms->center0.z;
            p.v[3] = ms->center1.x; p.v[4] = ms->center1.y; p.v[5] = ms->center1.z;
            p.v[6] = ms->time0; p.v[7] = ms->time1; p.v[8] = ms->radius;
            out.prims.push_back(p);
        } else if (auto *l = std::get_if<HittableList>(&h.v)) {
            for (const auto &o : l->objects) walk(o, chain);
        } else if (auto *r = std::get_if<XyRect>(&h.v)) {
            rect(RTW_PRIM_XY_RECT, r->x0, r->x1, r->y0, r->y1, r->k, r->material, chain);
        } else if (auto *r2 = std::get_if<XzRect>(&h.v)) {
            rect(RTW_PRIM_XZ_RECT, r2->x0, r2->x1, r2->z0, r2->z1, r2->k, r2->material, chain);
        } else if (auto *r3 = std::get_if<YzRect>(&h.v)) {
            rect(RTW_PRIM_YZ_RECT, r3->y0, r3->y1, r3->z0, r3->z1, r3->k, r3->material, chain);
        } else if (auto *b = std::get_if<Box>(&h.v)) {
            for (const auto &o : b->sides->objects) walk(o, chain);
        } else if (auto *t = std::get_if<Translate>(&h.v)) {
            rtw_xform x{};
            x.kind = RTW_XFORM_TRANSLATE; x.outer = chain;
            x.v[0] = t->offset.x; x.v[1] = t->offset.y; x.v[2] = t->offset.z;
            out.xforms.push_back(x);
            walk(*t->object, (int)out.xforms.size() - 1);
        } else {
            const auto &ry = std::get<RotateY>(h.v);
            rtw_xform x{};
            x.kind = RTW_XFORM_ROTATE_Y; x.outer = chain;
            x.v[0] = ry.sin_t; x.v[1] = ry.cos_t;
            out.xforms.push_back(x);
            walk(*ry.object, (int)out.xforms.size() - 1);
        }
    }
};
}  // namespace

FlatScene flatten(const Hittable &world, double time0, double time1) {
    Flattener f;
    f.walk(world, -1);
    f.out.time0 = time0;
    f.out.time1 = time1;
    for (const auto &im : f.out.keep_images) f.out.images.push_back({im->width, im->height, im->rgba.data()});
    for (size_t i = 0; i < f.out.keep_perlins.size(); ++i) {
        const auto &p = f.out.keep_perlins[i];
        f.out.perlins.push_back({f.out.ranvec_store[i].data(), p->perm[0], p->perm[1], p->perm[2]});
    }
    return std::move(f.out);
}

rtw_scene_desc FlatScene::desc() const {
    rtw_scene_desc d{};
    d.n_prims = (uint32_t)prims.size(); d.prims = prims.data();
    d.n_xforms = (uint32_t)xforms.size(); d.xforms = xforms.data();
    d.n_materials = (uint32_t)materials.size(); d.materials = materials.data();
    d.n_textures = (uint32_t)textures.size(); d.textures = textures.data();
    d.n_images = (uint32_t)images.size(); d.images = images.data();
    d.n_perlins = (uint32_t)perlins.size(); d.perlins = perlins.data();
    d.time0 = time0; d.time1 = time1;
    return d;
}

// ------------------------------------------------------------------------------------------------
// Scenes
// ------------------------------------------------------------------------------------------------
namespace {
Rc<Material> mat(Material m) { return std::make_shared<Material>(std::move(m)); }
Rc<Material> diffuse(Texture t) { return mat({DiffuseMaterial{std::move(t)}}); }
Rc<Material> metal(Color a, double fuzz) { return mat({MetalMaterial{a, fuzz}}); }
Rc<Material> glass(double ir) { return mat({DielectricMaterial{ir}}); }
Rc<Material> light(Texture t) { return mat({DiffuseLightMaterial{std::move(t)}}); }

// Body of generateRandomScene's a,b loops (main.zig:177-218).  `lift` puts the small spheres on the
// curved ground (config C4); the reference scene uses y = 0.2.
void scatterSmallSpheres(std::vector<Hittable> &objs, Random &rng, int grid, bool lift) {
    for (int a = -grid; a < grid; ++a)
        for (int b = -grid; b < grid; ++b) {
            const double choose_mat = rng.real01();
            Point3 center{(double)a + 0.9 * rng.real01(), 0.2, (double)b + 0.9 * rng.real01()};
            if (length(sub(center, {4, 0.2, 0})) <= 0.9) continue;
            if (lift) center.y = std::sqrt(1000.2 * 1000.2 - center.x * center.x - center.z * center.z) - 1000.0;
            if (choose_mat < 0.8) {
                const Color c1{rng.real01(), rng.real01(), rng.real01()};
                const Color c2{rng.real01(), rng.real01(), rng.real01()};
                auto m = diffuse(Texture::makeSolid({c1.x * c2.x, c1.y * c2.y, c1.z * c2.z}));
                const Point3 center1{center.x, center.y + rng.real(0, 0.5), center.z};
                objs.push_back({MovingSphere{center, center1, 0, 1, 0.2, m}});
            } else if (choose_mat < 0.95) {
                const Color albedo{rng.real(0.5, 1), rng.real(0.5, 1), rng.real(0.5, 1)};
                const double fuzz = rng.real(0, 0.5);
                objs.push_back(makeSphere(center, 0.2, metal(albedo, fuzz)));
            } else {
                objs.push_back(makeSphere(center, 0.2, glass(1.5)));
            }
        }
}

Hittable randomScene(Random &rng, int grid, bool lift) {  // main.zig:157-221
    HittableList l;
    l.objects.push_back(makeSphere({0, -1000, 0}, 1000, diffuse(Texture::makeChecker(rgb(0.2, 0.3, 0.1), rgb(0.9, 0.9, 0.9)))));
    l.objects.push_back(makeSphere({0, 1, 0}, 1.0, glass(1.5)));
    l.objects.push_back(makeSphere({-4, 1, 0}, 1.0, diffuse(Texture::makeSolid(rgb(0.4, 0.2, 0.1)))));
    l.objects.push_back(makeSphere({4, 1, 0}, 1.0, metal(rgb(0.7, 0.6, 0.5), 0.0)));
    scatterSmallSpheres(l.objects, rng, grid, lift);
    return {std::move(l)};
}

Hittable twoSpheres() {  // main.zig:124-139
    HittableList l;
    l.objects.push_back(makeSphere({0, -10, 0}, 10, diffuse(Texture::makeChecker(rgb(0.2, 0.3, 0.1), rgb(0.9, 0.9, 0.9)))));
    l.objects.push_back(makeSphere({0, 10, 0}, 10, diffuse(Texture::makeChecker(rgb(0.2, 0.3, 0.1), rgb(0.9, 0.9, 0.9)))));
    return {std::move(l)};
}

Hittable twoPerlinSpheres(Random &rng, bool with_light) {  // main.zig:141-155, 236-257
    HittableList l;
    const Texture perlin = Texture::makeNoise(4.0, rng);
    l.objects.push_back(makeSphere({0, -1000, 0}, 1000, diffuse(perlin)));
    l.objects.push_back(makeSphere({0, 2, 0}, 2, diffuse(perlin)));
    if (with_light) l.objects.push_back({XyRect{3.0, 5.0, 1.0, 3.0, -2.0, light(Texture::makeSolid(rgb(4, 4, 4)))}});
    return {std::move(l)};
}

Hittable earth(const std::string &asset) {  // main.zig:223-234
    HittableList l;
    l.objects.push_back(makeSphere({0, 0, 0}, 2, diffuse(Texture::makeImage(asset))));
    return {std::move(l)};
}

Hittable earthGlassMetal(const std::string &asset) {  // config C3: scene 4 + glass + metal + checker ground
    HittableList l;
    l.objects.push_back(makeSphere({0, 0, 0}, 2, diffuse(Texture::makeImage(asset))));
    l.objects.push_back(makeSphere({0.72, -1, -3.12}, 1.0, glass(1.5)));
    l.objects.push_back(makeSphere({-0.72, -1, 3.12}, 1.0, metal(rgb(0.7, 0.6, 0.5), 0.0)));
    l.objects.push_back(makeSphere({0, -1002, 0}, 1000, diffuse(Texture::makeChecker(rgb(0.2, 0.3, 0.1), rgb(0.9, 0.9, 0.9)))));
    return {std::move(l)};
}

Hittable cornellBox() {  // main.zig:259-293
    HittableList l;
    auto red = diffuse(Texture::makeSolid(rgb(0.65, 0.05, 0.05)));
    auto white = diffuse(Texture::makeSolid(rgb(0.73, 0.73, 0.73)));
    auto green = diffuse(Texture::makeSolid(rgb(0.12, 0.45, 0.15)));
    auto lamp = light(Texture::makeSolid(rgb(15, 15, 15)));
    l.objects.push_back({YzRect{0, 555, 0, 555, 555, green}});
    l.objects.push_back({YzRect{0, 555, 0, 555, 0, red}});
    l.objects.push_back({XzRect{213, 343, 227, 332, 554, lamp}});
    l.objects.push_back({XzRect{0, 555, 0, 555, 0, white}});
    l.objects.push_back({XzRect{0, 555, 0, 555, 555, white}});
    l.objects.push_back({XyRect{0, 555, 0, 555, 555, white}});
    const double deg = 3.14159265358979323846 / 180.0;
    auto box1 = std::make_shared<Hittable>(Hittable::makeBox({0, 0, 0}, {165, 330, 165}, white));
    auto box1r = std::make_shared<Hittable>(Hittable::makeRotateY(box1, 15.0 * deg));
    l.objects.push_back(Hittable::makeTranslate(box1r, {265, 0, 295}));
    auto box2 = std::make_shared<Hittable>(Hittable::makeBox({0, 0, 0}, {165, 165, 165}, white));
    auto box2r = std::make_shared<Hittable>(Hittable::makeRotateY(box2, -18.0 * deg));
    l.objects.push_back(Hittable::makeTranslate(box2r, {130, 0, 65}));
    return {std::move(l)};
}
}  // namespace

SceneSetup makeScene(int scene, int grid, uint64_t seed, const std::string &asset) {
    Random rng(seed);  // main.zig:300
    SceneSetup s;
    s.background = rgb(0.70, 0.80, 1.00);
    s.look_from = {13, 2, 3};
    s.look_at = {0, 0, 0};
    s.vfov = 20.0;
    switch (scene) {  // main.zig:320-362
        case 1: s.world = randomScene(rng, grid, false); s.aperture = 0.1; break;
        case 2: s.world = twoSpheres(); break;
        case 3: s.world = twoPerlinSpheres(rng, false); break;
        case 4: s.world = earth(asset); break;
        case 5:
            s.world = twoPerlinSpheres(rng, true);
            s.background = rgb(0, 0, 0); s.look_from = {26, 3, 6}; s.look_at = {0, 2, 0}; s.samples_per_pixel = 400;
            break;
        case 6:
            s.world = cornellBox();
            s.background = rgb(0, 0, 0); s.look_from = {278, 278, -800}; s.look_at = {278, 278, 0}; s.vfov = 40.0;
            s.aspect_ratio = 1.0; s.image_width = 600; s.samples_per_pixel = 200;
            break;
        case 7:
            s.world = earthGlassMetal(asset);
            s.aspect_ratio = 16.0 / 9.0; s.image_width = 1920; s.samples_per_pixel = 1000;
            break;
        case 8:
            s.world = randomScene(rng, grid, true);
            s.look_from = {0, 300, 600}; s.vfov = 40.0;
            s.aspect_ratio = 16.0 / 9.0; s.image_width = 1920; s.samples_per_pixel = 256;
            break;
        default: throw std::runtime_error("unknown scene id");
    }
    s.image_height = (uint32_t)std::trunc((double)s.image_width / s.aspect_ratio);  // main.zig:306
    return s;
}

}  // namespace rtw_host
