// rtw_oracle.hpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the per-pixel path-tracing loop of nsfisis/RayTracingInOneWeekend.zig,
// used to CHECK the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call anything in oracle/.  The
// product (raytracinginoneweekend.zig_b200/) never includes, links or executes this code.
//
// PINNING.  The reference ships no tests, golden vectors or fixtures and cannot be compiled here (no
// `zig` toolchain; zigimg is fetched from the network, build.zig.zon:69-72; only scene 6 type-checks, SURVEY.md §0
// D4), so there is no oracle/_ref binary.  What pins this restatement instead is the reference's SOURCE TEXT, executed:
// tests/ref_transpile.py transpiles /root/reference/src/*.zig to Python at test time (same f64 statements, same order)
// and tests/test_ref_pin.py demands bit-for-bit equality with this oracle for main() on all six scenes, every
// Hittable.hit / boudingBox body, Aabb.hit, the materials, textures, Perlin, Camera, rayColor and the rejection
// samplers (>= 10^4 random inputs per family); tests/golden/ref_golden.json carries the reference's outputs to machines
// without /root/reference (tests/test_ref_golden.py).  Still NOT pinned by anything in the reference's tree: Zig std's
// DefaultPrng / Random.float(f64) / uintLessThan and libm (restated from the published algorithms, SURVEY.md App. E;
// the Xoshiro256++ vector of Zig's own std test is checked in tests/test_oracle_kat.py) — they only decide WHICH
// scene seed 42 generates and which stream the CPU samples, never a statement of the path.
//
// Everything is templated on `Real`:
//   Real = double  — the reference's semantics (Vec3 is 3 x f64, src/rtw/vec.zig:8-11);
//   Real = float   — the same statements evaluated in fp32 on the (float)-rounded scene: what
//                    the device's fp32 arithmetic must reproduce bit-for-bit when both sides
//                    keep FMA contraction off (primary-hit id parity).
// Each function cites the reference lines it follows.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// Vec3 — src/rtw/vec.zig:8-109 (scalar struct, left-to-right evaluation order preserved)
// ------------------------------------------------------------------------------------------
template <class R>
struct V3 {
    R x, y, z;
    R norm2() const { return x * x + y * y + z * z; }                     // vec.zig:17-19
    R norm() const { return std::sqrt(norm2()); }                          // vec.zig:13-15
    static R dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; } // vec.zig:21-23
    static V3 cross(V3 u, V3 v) {                                          // vec.zig:25-31
        return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
    }
    V3 normalized() const {  // vec.zig:33-40: zero vector returned unchanged
        R n = norm();
        if (n == R(0)) return *this;
        return div(n);
    }
    V3 add(V3 v) const { return {x + v.x, y + v.y, z + v.z}; }
    V3 sub(V3 v) const { return {x - v.x, y - v.y, z - v.z}; }
    V3 mul(R t) const { return {x * t, y * t, z * t}; }
    V3 mulv(V3 v) const { return {x * v.x, y * v.y, z * v.z}; }
    V3 div(R t) const { return {x / t, y / t, z / t}; }
    bool near_zero() const {  // vec.zig:98-101
        const R e = R(1e-8);
        return std::fabs(x) < e && std::fabs(y) < e && std::fabs(z) < e;
    }
    template <class S>
    V3<S> as() const { return {S(x), S(y), S(z)}; }
};
using V3d = V3<double>;

// ------------------------------------------------------------------------------------------
// RNG — Zig std.Random.DefaultPrng (= Xoshiro256++ seeded through SplitMix64) and
// Random.float(f64) / uintLessThan, Zig >= 0.14 std (not under /root/reference; restated from
// the published algorithm).  Call sites: src/main.zig:300-301,
// src/rtw/rand.zig:5-20.
// ------------------------------------------------------------------------------------------
struct Xoshiro256pp {
    uint64_t s[4];
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    explicit Xoshiro256pp(uint64_t seed = 0) { reseed(seed); }
    void reseed(uint64_t seed) {
        uint64_t sm = seed;
        for (int i = 0; i < 4; ++i) {  // SplitMix64.next()
            sm += 0x9e3779b97f4a7c15ull;
            uint64_t z = sm;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
};

struct Rng {
    Xoshiro256pp g;
    uint64_t draws = 0;
    explicit Rng(uint64_t seed = 42) : g(seed) {}
    uint64_t u64() { ++draws; return g.next(); }
    static int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
    // Random.float(f64): 52 mantissa bits, geometrically distributed exponent.  rand.zig:13-15
    double real01() {
        const uint64_t r = u64();
        int lz = clz64(r);
        if (lz >= 12) {
            lz = 12;
            for (;;) {
                const int a = clz64(u64());
                lz += a;
                if (a != 64) break;
                if (lz >= 1022) { lz = 1022; break; }
            }
        }
        const uint64_t bits = (uint64_t(1022 - lz) << 52) | (r & 0xFFFFFFFFFFFFFull);
        double d;
        std::memcpy(&d, &bits, 8);
        return d;
    }
    double real(double lo, double hi) { return lo + real01() * (hi - lo); }  // rand.zig:18-20
    // Random.uintLessThan(u64, n): Lemire multiply-shift with rejection.  rand.zig:7-10
    uint64_t uint_less_than(uint64_t n) {
        uint64_t x = u64();
        unsigned __int128 m = (unsigned __int128)x * n;
        uint64_t l = (uint64_t)m;
        if (l < n) {
            const uint64_t t = (0 - n) % n;
            while (l < t) {
                x = u64();
                m = (unsigned __int128)x * n;
                l = (uint64_t)m;
            }
        }
        return (uint64_t)(m >> 64);
    }
};

template <class R>
inline V3<R> random_in_unit_sphere(Rng &g) {  // rand.zig:22-28 (rejection on norm() >= 1)
    for (;;) {
        V3<R> p{R(g.real(-1.0, 1.0)), R(g.real(-1.0, 1.0)), R(g.real(-1.0, 1.0))};
        if (p.norm() >= R(1)) continue;
        return p;
    }
}
template <class R>
inline V3<R> random_in_unit_disk(Rng &g) {  // rand.zig:30-36
    for (;;) {
        V3<R> p{R(g.real(-1.0, 1.0)), R(g.real(-1.0, 1.0)), R(0)};
        if (p.norm() >= R(1)) continue;
        return p;
    }
}
template <class R>
inline V3<R> random_unit_vector(Rng &g) { return random_in_unit_sphere<R>(g).normalized(); }  // rand.zig:38-40

// ------------------------------------------------------------------------------------------
// Ray / HitRecord — src/rtw/ray.zig:5-13, src/rtw/hit_record.zig:7-21
// ------------------------------------------------------------------------------------------
template <class R>
struct Ray {
    V3<R> o, d;
    R time;
    V3<R> at(R t) const { return o.add(d.mul(t)); }
};
template <class R>
struct HitRecord {
    V3<R> p, normal;
    int material = -1;
    R t = 0, u = 0, v = 0;  // MovingSphere leaves u,v undefined (hittable.zig:189-198): 0 here
    bool front_face = false;
    uint32_t prim_id = 0xFFFFFFFFu;  // added: depth-first leaf index
};

// ------------------------------------------------------------------------------------------
// Perlin — src/rtw/perlin.zig:10-125
// ------------------------------------------------------------------------------------------
struct Perlin {
    double ranvec[256][3];
    uint32_t perm[3][256];
    void init(Rng &g) {  // perlin.zig:18-38
        for (int i = 0; i < 256; ++i) {
            V3d v{g.real(-1, 1), g.real(-1, 1), g.real(-1, 1)};
            v = v.normalized();
            ranvec[i][0] = v.x; ranvec[i][1] = v.y; ranvec[i][2] = v.z;
            perm[0][i] = perm[1][i] = perm[2][i] = (uint32_t)i;
        }
        for (int a = 0; a < 3; ++a)  // permute: exclusive upper bound, perlin.zig:93-101
            for (int i = 255; i > 0; --i) {
                const uint64_t tgt = g.uint_less_than((uint64_t)i);
                std::swap(perm[a][i], perm[a][tgt]);
            }
    }
    template <class R>
    R noise(V3<R> p) const {  // perlin.zig:47-77 + perlinInterp :103-124
        const R fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
        const R u = p.x - fx, v = p.y - fy, w = p.z - fz;
        const R uu = u * u * (R(3) - R(2) * u);
        const R vv = v * v * (R(3) - R(2) * v);
        const R ww = w * w * (R(3) - R(2) * w);
        const int i = (int)fx, j = (int)fy, k = (int)fz;
        R accum = 0;
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj)
                for (int dk = 0; dk < 2; ++dk) {
                    const uint32_t idx = perm[0][(i + di) & 255] ^ perm[1][(j + dj) & 255] ^
                                         perm[2][(k + dk) & 255];
                    const V3<R> c{R(ranvec[idx][0]), R(ranvec[idx][1]), R(ranvec[idx][2])};
                    const R ti = R(di), tj = R(dj), tk = R(dk);
                    // perlin.zig:77 hands the Hermite-SMOOTHED u_, v_, w_ to perlinInterp, whose weight vector
                    // (perlin.zig:114) is therefore (u_ - i, v_ - j, w_ - k) — not the book's raw (u - i, ...)
                    const V3<R> wv{uu - ti, vv - tj, ww - tk};
                    accum += (ti * uu + (R(1) - ti) * (R(1) - uu)) *
                             (tj * vv + (R(1) - tj) * (R(1) - vv)) *
                             (tk * ww + (R(1) - tk) * (R(1) - ww)) * V3<R>::dot(c, wv);
                }
        return accum;
    }
    template <class R>
    R turb(V3<R> p, int depth) const {  // perlin.zig:79-91
        R accum = 0, weight = 1;
        for (int i = 0; i < depth; ++i) {
            accum += weight * noise<R>(p);
            weight *= R(0.5);
            p = p.mul(R(2));
        }
        return std::fabs(accum);
    }
};

// ------------------------------------------------------------------------------------------
// Textures — src/rtw/texture.zig:10-145
// ------------------------------------------------------------------------------------------
struct Image {
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> rgba;
};
struct Texture {
    enum Kind { SOLID = 0, CHECKER = 1, NOISE = 2, IMAGE = 3 } kind = SOLID;
    int a = -1, b = -1;  // checker: odd, even texture; noise: perlin index; image: image index
    V3d color{0, 0, 0};
    double scale = 1;
};

// ------------------------------------------------------------------------------------------
// Materials — src/rtw/material.zig:16-121
// ------------------------------------------------------------------------------------------
struct Material {
    enum Kind { DIFFUSE = 0, METAL = 1, DIELECTRIC = 2, DIFFUSE_LIGHT = 3 } kind = DIFFUSE;
    int texture = -1;
    V3d albedo{0, 0, 0};
    double param = 0;  // fuzz | ir
};

template <class R>
inline V3<R> reflect(V3<R> v, V3<R> n) { return v.sub(n.mul(R(2) * V3<R>::dot(v, n))); }  // material.zig:112-114
template <class R>
inline V3<R> refract(V3<R> uv, V3<R> n, R eta) {  // material.zig:116-121
    const R cos_theta = std::fmin(V3<R>::dot(uv.mul(R(-1)), n), R(1));
    const V3<R> perp = uv.add(n.mul(cos_theta)).mul(eta);
    const V3<R> par = n.mul(-std::sqrt(std::fabs(R(1) - perp.norm2())));
    return perp.add(par);
}
template <class R>
inline R reflectance(R cosv, R idx) {  // material.zig:87-91 (Schlick)
    const R r0 = (R(1) - idx) / (R(1) + idx);
    const R r1 = r0 * r0;
    return r1 + (R(1) - r1) * std::pow(R(1) - cosv, R(5));
}

// ------------------------------------------------------------------------------------------
// Hittable graph — src/rtw/hittable.zig:22-608.  Stored in f64 (the reference's type); every
// `hit` converts the fields it reads to Real, which is the identity for Real=double.
// ------------------------------------------------------------------------------------------
struct Hittable {
    enum Kind { SPHERE, MOVING_SPHERE, LIST, XY_RECT, XZ_RECT, YZ_RECT, BOX, TRANSLATE, ROTATE_Y } kind = LIST;
    // sphere / moving sphere
    V3d center{0, 0, 0}, center1{0, 0, 0};
    double radius = 0, time0 = 0, time1 = 1;
    // rects: a = first in-plane axis, b = second (xy: a=x b=y; xz: a=x b=z; yz: a=y b=z)
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;
    int material = -1;
    uint32_t prim_id = 0xFFFFFFFFu;  // leaves only; assigned by Scene::assign_ids()
    // list / box
    std::vector<Hittable> objects;
    V3d box_min{0, 0, 0}, box_max{0, 0, 0};
    // translate / rotateY
    std::shared_ptr<Hittable> object;
    V3d offset{0, 0, 0};
    double sin_t = 0, cos_t = 1;
};

struct Aabb {  // src/rtw/aabb.zig:4-61
    V3d min, max;
    static Aabb surrounding(const Aabb &a, const Aabb &b) {  // aabb.zig:47-60
        return {{std::fmin(a.min.x, b.min.x), std::fmin(a.min.y, b.min.y), std::fmin(a.min.z, b.min.z)},
                {std::fmax(a.max.x, b.max.x), std::fmax(a.max.y, b.max.y), std::fmax(a.max.z, b.max.z)}};
    }
    template <class R>
    bool hit(const Ray<R> &r, R t_min, R t_max) const {  // aabb.zig:8-45 (division form)
        const R mn[3] = {R(min.x), R(min.y), R(min.z)}, mx[3] = {R(max.x), R(max.y), R(max.z)};
        const R o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
        for (int a = 0; a < 3; ++a) {
            const R s0 = (mn[a] - o[a]) / d[a];
            const R s1 = (mx[a] - o[a]) / d[a];
            const R t0 = std::fmin(s0, s1), t1 = std::fmax(s0, s1);
            t_min = std::fmax(t0, t_min);
            t_max = std::fmin(t1, t_max);
            if (t_max <= t_min) return false;
        }
        return true;
    }
};

inline Hittable make_sphere(V3d c, double r, int mat) {
    Hittable h; h.kind = Hittable::SPHERE; h.center = c; h.radius = r; h.material = mat; return h;
}
inline Hittable make_moving_sphere(V3d c0, V3d c1, double t0, double t1, double r, int mat) {
    Hittable h; h.kind = Hittable::MOVING_SPHERE; h.center = c0; h.center1 = c1; h.time0 = t0; h.time1 = t1;
    h.radius = r; h.material = mat; return h;
}
inline Hittable make_rect(Hittable::Kind kind, double a0, double a1, double b0, double b1, double k, int mat) {
    Hittable h; h.kind = kind; h.a0 = a0; h.a1 = a1; h.b0 = b0; h.b1 = b1; h.k = k; h.material = mat; return h;
}
inline Hittable make_list() { Hittable h; h.kind = Hittable::LIST; return h; }
inline Hittable make_box(V3d p0, V3d p1, int mat) {  // Box.init hittable.zig:434-451 (side order!)
    Hittable h; h.kind = Hittable::BOX; h.box_min = p0; h.box_max = p1;
    h.objects.push_back(make_rect(Hittable::XY_RECT, p0.x, p1.x, p0.y, p1.y, p1.z, mat));
    h.objects.push_back(make_rect(Hittable::XY_RECT, p0.x, p1.x, p0.y, p1.y, p0.z, mat));
    h.objects.push_back(make_rect(Hittable::XZ_RECT, p0.x, p1.x, p0.z, p1.z, p1.y, mat));
    h.objects.push_back(make_rect(Hittable::XZ_RECT, p0.x, p1.x, p0.z, p1.z, p0.y, mat));
    h.objects.push_back(make_rect(Hittable::YZ_RECT, p0.y, p1.y, p0.z, p1.z, p1.x, mat));
    h.objects.push_back(make_rect(Hittable::YZ_RECT, p0.y, p1.y, p0.z, p1.z, p0.x, mat));
    return h;
}
inline Hittable make_translate(Hittable child, V3d offset) {  // hittable.zig:39-41
    Hittable h; h.kind = Hittable::TRANSLATE; h.object = std::make_shared<Hittable>(std::move(child));
    h.offset = offset; return h;
}
inline Hittable make_rotate_y_sc(Hittable child, double s, double c) {
    Hittable h; h.kind = Hittable::ROTATE_Y; h.object = std::make_shared<Hittable>(std::move(child));
    h.sin_t = s; h.cos_t = c; return h;
}
inline Hittable make_rotate_y(Hittable child, double angle) {  // RotateY.init hittable.zig:513-515
    return make_rotate_y_sc(std::move(child), std::sin(angle), std::cos(angle));
}

// boudingBox (sic) — the reference's leaf/instance box rules, hittable.zig:61-73 and the
// per-variant bodies cited below.  Dead code at render time in the reference; here
// the spec the new BVH's leaf boxes are checked against.
inline bool bounding_box(const Hittable &h, double time0, double time1, Aabb &out) {
    switch (h.kind) {
        case Hittable::SPHERE: {  // hittable.zig:133-143
            const V3d rr{h.radius, h.radius, h.radius};
            out = {h.center.sub(rr), h.center.add(rr)};
            return true;
        }
        case Hittable::MOVING_SPHERE: {  // hittable.zig:203-217, centre(t) :219-221
            auto ctr = [&](double t) {
                return h.center.add(h.center1.sub(h.center).mul((t - h.time0) / (h.time1 - h.time0)));
            };
            const V3d rr{h.radius, h.radius, h.radius};
            const V3d o0 = ctr(time0), o1 = ctr(time1);
            out = Aabb::surrounding({o0.sub(rr), o0.add(rr)}, {o1.sub(rr), o1.add(rr)});
            return true;
        }
        case Hittable::XY_RECT:  // hittable.zig:305-316
            out = {{h.a0, h.b0, h.k - 0.0001}, {h.a1, h.b1, h.k + 0.0001}};
            return true;
        case Hittable::XZ_RECT:  // hittable.zig:358-369
            out = {{h.a0, h.k - 0.0001, h.b0}, {h.a1, h.k + 0.0001, h.b1}};
            return true;
        case Hittable::YZ_RECT:  // hittable.zig:411-422
            out = {{h.k - 0.0001, h.a0, h.b0}, {h.k + 0.0001, h.a1, h.b1}};
            return true;
        case Hittable::BOX:  // hittable.zig:457-465
            out = {h.box_min, h.box_max};
            return true;
        case Hittable::LIST: {  // hittable.zig:246-260
            if (h.objects.empty()) return false;
            bool first = true;
            Aabb tmp;
            for (const auto &o : h.objects) {
                if (!bounding_box(o, time0, time1, tmp)) return false;
                out = first ? tmp : Aabb::surrounding(out, tmp);
                first = false;
            }
            return true;
        }
        case Hittable::TRANSLATE: {  // hittable.zig:491-498
            const bool r = bounding_box(*h.object, time0, time1, out);
            out = {out.min.add(h.offset), out.max.add(h.offset)};
            return r;
        }
        case Hittable::ROTATE_Y: {  // RotateY.init hittable.zig:516-556 (8 corners), :598-603
            Aabb bb;
            (void)bounding_box(*h.object, 0, 0, bb);
            const double inf = std::numeric_limits<double>::infinity();
            V3d mn{inf, inf, inf}, mx{-inf, -inf, -inf};
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int k = 0; k < 2; ++k) {
                        const double x = i * bb.max.x + (1.0 - i) * bb.min.x;
                        const double y = j * bb.max.y + (1.0 - j) * bb.min.y;
                        const double z = k * bb.max.z + (1.0 - k) * bb.min.z;
                        const double nx = h.cos_t * x + h.sin_t * z;
                        const double nz = -h.sin_t * x + h.cos_t * z;
                        mn = {std::fmin(mn.x, nx), std::fmin(mn.y, y), std::fmin(mn.z, nz)};
                        mx = {std::fmax(mx.x, nx), std::fmax(mx.y, y), std::fmax(mx.z, nz)};
                    }
            out = {mn, mx};
            return true;
        }
    }
    return false;
}

// ---- hit(): closest-hit with the reference's exact statements -------------------------------
template <class R>
bool hit(const Hittable &h, const Ray<R> &r, R t_min, R t_max, HitRecord<R> &rec);

template <class R>
inline void sphere_uv(V3<R> p, R &u, R &v) {  // getSphereUv hittable.zig:145-150
    const R pi = R(3.14159265358979323846);
    const R phi = std::atan2(-p.z, p.x) + pi;
    const R theta = std::acos(-p.y);
    u = phi / (R(2) * pi);
    v = theta / pi;
}

template <class R>
inline bool hit_sphere_at(const Hittable &h, V3<R> center, bool write_uv, const Ray<R> &r, R t_min, R t_max,
                          HitRecord<R> &rec) {
    // Sphere.hit hittable.zig:95-131 / MovingSphere.hit :165-201 (identical but for centre, uv)
    const R radius = R(h.radius);
    const V3<R> oc = r.o.sub(center);
    const R a = r.d.norm2();
    const R half_b = V3<R>::dot(oc, r.d);
    const R c = oc.norm2() - radius * radius;
    const R disc = half_b * half_b - a * c;
    if (disc < R(0)) return false;
    const R sqrtd = std::sqrt(disc);
    R root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    rec.t = root;
    rec.p = r.at(root);
    const V3<R> outward = rec.p.sub(center).div(radius);
    rec.front_face = V3<R>::dot(outward, r.d) < R(0);
    rec.normal = rec.front_face ? outward : outward.mul(R(-1));
    if (write_uv) sphere_uv<R>(outward, rec.u, rec.v);
    else { rec.u = 0; rec.v = 0; }
    rec.material = h.material;
    rec.prim_id = h.prim_id;
    return true;
}

template <class R>
inline bool hit_rect(const Hittable &h, const Ray<R> &r, R t_min, R t_max, HitRecord<R> &rec) {
    // XyRect/XzRect/YzRect.hit hittable.zig:278-303 / 331-356 / 384-409
    const R o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    int ax_a, ax_b, ax_k;
    if (h.kind == Hittable::XY_RECT) { ax_a = 0; ax_b = 1; ax_k = 2; }
    else if (h.kind == Hittable::XZ_RECT) { ax_a = 0; ax_b = 2; ax_k = 1; }
    else { ax_a = 1; ax_b = 2; ax_k = 0; }
    const R a0 = R(h.a0), a1 = R(h.a1), b0 = R(h.b0), b1 = R(h.b1), k = R(h.k);
    const R t = (k - o[ax_k]) / d[ax_k];
    if (t < t_min || t > t_max) return false;  // NaN t falls through like the reference
    const R pa = o[ax_a] + t * d[ax_a];
    const R pb = o[ax_b] + t * d[ax_b];
    if (pa < a0 || pa > a1 || pb < b0 || pb > b1) return false;
    rec.u = (pa - a0) / (a1 - a0);
    rec.v = (pb - b0) / (b1 - b0);
    rec.t = t;
    rec.material = h.material;
    rec.p = r.at(t);
    V3<R> outward{0, 0, 0};
    (ax_k == 0 ? outward.x : ax_k == 1 ? outward.y : outward.z) = R(1);
    rec.front_face = V3<R>::dot(outward, r.d) < R(0);
    rec.normal = rec.front_face ? outward : outward.mul(R(-1));
    rec.prim_id = h.prim_id;
    return true;
}

template <class R>
inline bool hit_list(const std::vector<Hittable> &objs, const Ray<R> &r, R t_min, R t_max, HitRecord<R> &rec) {
    // HittableList.hit hittable.zig:231-244: linear scan, shrinking t_max, later wins on ties
    bool any = false;
    R closest = t_max;
    for (const auto &o : objs) {
        HitRecord<R> tmp;
        if (hit<R>(o, r, t_min, closest, tmp)) {
            any = true;
            closest = tmp.t;
            rec = tmp;
        }
    }
    return any;
}

template <class R>
bool hit(const Hittable &h, const Ray<R> &r, R t_min, R t_max, HitRecord<R> &rec) {  // hittable.zig:47-59
    switch (h.kind) {
        case Hittable::SPHERE:
            return hit_sphere_at<R>(h, h.center.as<R>(), true, r, t_min, t_max, rec);
        case Hittable::MOVING_SPHERE: {  // centre(t) hittable.zig:219-221
            const V3<R> c0 = h.center.as<R>(), c1 = h.center1.as<R>();
            const V3<R> c = c0.add(c1.sub(c0).mul((r.time - R(h.time0)) / (R(h.time1) - R(h.time0))));
            return hit_sphere_at<R>(h, c, false, r, t_min, t_max, rec);
        }
        case Hittable::LIST:
            return hit_list<R>(h.objects, r, t_min, t_max, rec);
        case Hittable::XY_RECT:
        case Hittable::XZ_RECT:
        case Hittable::YZ_RECT:
            return hit_rect<R>(h, r, t_min, t_max, rec);
        case Hittable::BOX:  // hittable.zig:453-455
            return hit_list<R>(h.objects, r, t_min, t_max, rec);
        case Hittable::TRANSLATE: {  // hittable.zig:478-489
            const Ray<R> rr{r.o.sub(h.offset.as<R>()), r.d, r.time};
            if (!hit<R>(*h.object, rr, t_min, t_max, rec)) return false;
            rec.p = rec.p.add(h.offset.as<R>());
            return true;
        }
        case Hittable::ROTATE_Y: {  // hittable.zig:558-596
            const R c = R(h.cos_t), s = R(h.sin_t);
            V3<R> o = r.o, d = r.d;
            o.x = c * r.o.x - s * r.o.z;
            o.z = s * r.o.x + c * r.o.z;
            d.x = c * r.d.x - s * r.d.z;
            d.z = s * r.d.x + c * r.d.z;
            const Ray<R> rr{o, d, r.time};
            if (!hit<R>(*h.object, rr, t_min, t_max, rec)) return false;
            V3<R> p = rec.p, n = rec.normal;
            p.x = c * rec.p.x + s * rec.p.z;
            p.z = -s * rec.p.x + c * rec.p.z;
            n.x = c * rec.normal.x + s * rec.normal.z;
            n.z = -s * rec.normal.x + c * rec.normal.z;
            rec.p = p;
            rec.normal = n;
            return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// Sources of randomness for getRay / scatter.  RngSampler = the reference's behaviour (rand.zig:13-40 on the one
// sequential stream); GivenSampler = values supplied by a test.
// ------------------------------------------------------------------------------------------
struct RngSampler {
    Rng &g;
    double real01() { return g.real01(); }
    template <class R> V3<R> in_unit_sphere() { return random_in_unit_sphere<R>(g); }
    template <class R> V3<R> in_unit_disk() { return random_in_unit_disk<R>(g); }
    template <class R> V3<R> unit_vector() { return random_unit_vector<R>(g); }
};
struct GivenSampler {
    V3d vec{0, 0, 0};  // the unit vector / ball point / disk point the call will consume
    double xi = 0;     // the uniform it will consume
    double real01() { return xi; }
    template <class R> V3<R> in_unit_sphere() { return vec.as<R>(); }
    template <class R> V3<R> in_unit_disk() { return vec.as<R>(); }
    template <class R> V3<R> unit_vector() { return vec.as<R>(); }
};

// ------------------------------------------------------------------------------------------
// Camera — src/main.zig:40-101
// ------------------------------------------------------------------------------------------
struct Camera {
    V3d origin, horizontal, vertical, lower_left_corner, u, v, w;
    double lens_radius, time0, time1;
    static Camera init(V3d look_from, V3d look_at, V3d vup, double vfov, double aspect, double aperture,
                       double focus_dist, double time0, double time1) {  // main.zig:52-89
        const double pi = 3.14159265358979323846;
        const double theta = vfov * pi / 180.0;  // deg2rad main.zig:36-38
        const double h = std::tan(theta / 2);
        const double vh = 2.0 * h;
        const double vw = aspect * vh;
        Camera c;
        c.w = look_from.sub(look_at).normalized();
        c.u = V3d::cross(vup, c.w).normalized();
        c.v = V3d::cross(c.w, c.u);
        c.origin = look_from;
        c.horizontal = c.u.mul(vw * focus_dist);
        c.vertical = c.v.mul(vh * focus_dist);
        c.lower_left_corner = c.origin.sub(c.horizontal.div(2.0)).sub(c.vertical.div(2.0)).sub(c.w.mul(focus_dist));
        c.lens_radius = aperture / 2.0;
        c.time0 = time0;
        c.time1 = time1;
        return c;
    }
    template <class R>
    Ray<R> ray_from(V3<R> lens_offset, R s, R t, R time) const {  // body of getRay main.zig:93-99
        const V3<R> dir = lower_left_corner.as<R>()
                              .add(horizontal.as<R>().mul(s))
                              .add(vertical.as<R>().mul(t))
                              .sub(origin.as<R>())
                              .sub(lens_offset);
        return {origin.as<R>().add(lens_offset), dir, time};
    }
    // getRay main.zig:91-100 with the source of randomness abstracted: S provides in_unit_disk() and real01().
    // RngSampler draws them from the sequential stream exactly as the reference does; GivenSampler replays
    // caller-supplied values (unit-level comparison with the device, whose samplers are rejection-free).
    template <class R, class S>
    Ray<R> get_ray_s(S &smp, R s, R t) const {
        const V3<R> rd = smp.template in_unit_disk<R>().mul(R(lens_radius));
        const V3<R> offset = u.as<R>().mul(rd.x).add(v.as<R>().mul(rd.y));
        Ray<R> r = ray_from<R>(offset, s, t, R(0));
        r.time = R(time0 + smp.real01() * (time1 - time0));  // randomReal rand.zig:18-20
        return r;
    }
    template <class R>
    Ray<R> get_ray(Rng &g, R s, R t) const;
};
template <class R>
inline Ray<R> Camera::get_ray(Rng &g, R s, R t) const {
    RngSampler smp{g};
    return get_ray_s<R>(smp, s, t);
}

// ------------------------------------------------------------------------------------------
// Scene = world + tables
// ------------------------------------------------------------------------------------------
struct Counters {
    uint64_t paths = 0, rays = 0;
};

struct Scene {
    Hittable world;  // always a LIST at the top (main.zig builders return .list)
    std::vector<Material> materials;
    std::vector<Texture> textures;
    std::vector<Image> images;
    std::vector<Perlin> perlins;
    double time0 = 0, time1 = 1;
    uint32_t n_leaves = 0;

    static void assign(Hittable &h, uint32_t &next) {
        switch (h.kind) {
            case Hittable::LIST:
            case Hittable::BOX:
                for (auto &o : h.objects) assign(o, next);
                break;
            case Hittable::TRANSLATE:
            case Hittable::ROTATE_Y:
                assign(*h.object, next);
                break;
            default:
                h.prim_id = next++;
        }
    }
    void assign_ids() {  // depth-first append order
        uint32_t n = 0;
        assign(world, n);
        n_leaves = n;
    }

    template <class R>
    V3<R> texture_value(int ti, R u, R v, V3<R> p) const {  // Texture.value texture.zig:36-43
        const Texture &tx = textures[ti];
        switch (tx.kind) {
            case Texture::SOLID:  // texture.zig:46-55
                return tx.color.as<R>();
            case Texture::CHECKER: {  // texture.zig:79-82
                const R sines = std::sin(R(10) * p.x) * std::sin(R(10) * p.y) * std::sin(R(10) * p.z);
                return sines < R(0) ? texture_value<R>(tx.a, u, v, p) : texture_value<R>(tx.b, u, v, p);
            }
            case Texture::NOISE: {  // texture.zig:100-104
                const R s = R(0.5) * (R(1) + std::sin(R(tx.scale) * p.z + R(10) * perlins[tx.a].turb<R>(p, 7)));
                return V3<R>{R(1), R(1), R(1)}.mul(s);
            }
            case Texture::IMAGE: {  // texture.zig:121-144
                const Image &im = images[tx.a];
                const R uc = std::fmin(std::fmax(u, R(0)), R(1));
                const R vc = R(1) - std::fmin(std::fmax(v, R(0)), R(1));
                size_t i = (size_t)(uc * R(im.w));
                size_t j = (size_t)(vc * R(im.h));
                if (i > im.w - 1) i = im.w - 1;
                // texture.zig:130 clamps j with WIDTH-1 (out-of-bounds read when v==0 and
                // h < w).  Deliberately fixed here to height-1 (measure-zero).
                if (j > im.h - 1) j = im.h - 1;
                const R cs = R(1) / R(255);
                const size_t off = j * im.w * 4 + i * 4;
                const R rr = R(im.rgba[off + 0]), gg = R(im.rgba[off + 1]), bb = R(im.rgba[off + 2]);
                if (im.rgba[off + 3] == 0) return {R(0), R(0), R(1)};  // "Ocean"
                return {cs * rr, cs * gg, cs * bb};
            }
        }
        return {0, 0, 0};
    }

    template <class R>
    V3<R> emitted(const HitRecord<R> &rec) const {  // Material.emitted material.zig:31-38, :107-109
        const Material &m = materials[rec.material];
        if (m.kind != Material::DIFFUSE_LIGHT) return {0, 0, 0};
        return texture_value<R>(m.texture, rec.u, rec.v, rec.p);
    }

    template <class R>
    bool scatter(const Ray<R> &r_in, const HitRecord<R> &rec, V3<R> &att, Ray<R> &out, Rng &g) const {
        RngSampler smp{g};
        return scatter_s<R>(r_in, rec, att, out, smp);
    }
    template <class R, class S>
    bool scatter_s(const Ray<R> &r_in, const HitRecord<R> &rec, V3<R> &att, Ray<R> &out, S &smp) const {
        const Material &m = materials[rec.material];  // Material.scatter material.zig:22-29
        switch (m.kind) {
            case Material::DIFFUSE: {  // material.zig:44-52
                V3<R> dir = rec.normal.add(smp.template unit_vector<R>());
                if (dir.near_zero()) dir = rec.normal;
                out = {rec.p, dir, r_in.time};
                att = texture_value<R>(m.texture, rec.u, rec.v, rec.p);
                return true;
            }
            case Material::METAL: {  // material.zig:59-65 (tests the UN-fuzzed reflection, Q17)
                const V3<R> refl = reflect<R>(r_in.d.normalized(), rec.normal);
                out = {rec.p, refl.add(smp.template in_unit_sphere<R>().mul(R(m.param))), r_in.time};
                att = m.albedo.as<R>();
                return V3<R>::dot(refl, rec.normal) > R(0);
            }
            case Material::DIELECTRIC: {  // material.zig:72-85
                const R ir = R(m.param);
                const R ratio = rec.front_face ? R(1) / ir : ir;
                const V3<R> ud = r_in.d.normalized();
                const R cos_theta = std::fmin(V3<R>::dot(ud.mul(R(-1)), rec.normal), R(1));
                const R sin_theta = std::sqrt(R(1) - cos_theta * cos_theta);
                const bool can_refract = ratio * sin_theta <= R(1);
                // short-circuit: the draw happens only when refraction is possible (Q13)
                const V3<R> dir = (can_refract && reflectance<R>(cos_theta, ratio) < R(smp.real01()))
                                      ? refract<R>(ud, rec.normal, ratio)
                                      : reflect<R>(ud, rec.normal);
                out = {rec.p, dir, r_in.time};
                att = {R(1), R(1), R(1)};
                return true;
            }
            case Material::DIFFUSE_LIGHT:  // material.zig:97-105
                return false;
        }
        return false;
    }

    // rayColor — src/main.zig:103-122 (recursive, as the reference)
    template <class R>
    V3<R> ray_color(const Ray<R> &r, V3<R> background, Rng &g, uint32_t depth, Counters &cn) const {
        if (depth == 0) return {0, 0, 0};
        HitRecord<R> rec;
        ++cn.rays;
        if (!hit<R>(world, r, R(0.001), std::numeric_limits<R>::infinity(), rec)) return background;
        Ray<R> scattered;
        V3<R> att;
        const V3<R> em = emitted<R>(rec);
        if (scatter<R>(r, rec, att, scattered, g))
            return em.add(att.mulv(ray_color<R>(scattered, background, g, depth - 1, cn)));
        return em;
    }
};

// resolve — src/main.zig:395-400
inline uint8_t resolve_channel(double sum, double scale) {
    double c = std::sqrt(sum * scale);
    c = c < 0.0 ? 0.0 : (c > 0.999 ? 0.999 : c);  // std.math.clamp
    return (uint8_t)(256.0 * c);
}

}  // namespace orc
