// rtw_trace.cuh — production fp32 device code: RNG, closest-hit (flat scan and BVH), shading.
//
// What each piece replaces in the reference (all f64 there, fp32 here):
//   Philox4x32-10 + samplers  <- the single sequential Xoshiro stream and the rejection samplers
//                                 (src/main.zig:300-301, src/rtw/rand.zig:13-40)
//   closest_hit_flat          <- HittableList.hit linear scan (src/rtw/hittable.zig:231-244)
//   closest_hit_bvh           <- new (the reference has no BVH; Aabb.hit aabb.zig:8-45 is dead code)
//   sphere_test / rect_test   <- Sphere.hit / MovingSphere.hit / *Rect.hit (hittable.zig:95-131,
//                                 165-201, 278-303, 331-356, 384-409)
//   finalise_hit              <- the HitRecord fill of those functions + Translate/RotateY back-map
//   shade                     <- Material.emitted/scatter (material.zig:16-121), Texture.value
//                                 (texture.zig:36-145), one level of rayColor (main.zig:103-122)
#pragma once
#include "rtw_device.cuh"

namespace rtw {

// ---------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10.  key = (seed_lo, seed_hi), counter = (pixel, sample, block, 0).
// block 0 = camera ray, block b>=1 = bounce b.  Every draw of a (pixel, sample, bounce) is a pure
// function of its indices, so any partition of the samples over threads/GPUs gives the same paths.
// ---------------------------------------------------------------------------------------------
// The ten round keys (key + i * Weyl constants) are the same for every call of a launch: the host expands them once
// into DevRender::philox_keys, so a round is 2 IMAD.WIDE + 2 LOP3 with a constant-bank operand and no key arithmetic.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, const uint32_t (&keys)[20]) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ keys[2 * i], lo1, hi0 ^ ctr.w ^ keys[2 * i + 1], lo0);
    }
    return ctr;
}
// 24-bit uniform in [0,1): exactly the fp32 lattice, never 1.0
__device__ __forceinline__ float u01_24(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }

// 32 bytes per thread in ONE load instruction (LDG.E.256, new with sm_100): the BVH kernels fetch per-lane records from
// scattered addresses, where the L1 data pipe serves one wavefront per lane per load INSTRUCTION — ncu on the 10^6-sphere
// scene: l1tex__data_pipe_lsu_wavefronts at 77 % of peak with four LDG.128 per 64-byte sibling pair — so half as many
// instructions is half as many wavefronts.  `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void *p, float4 &a, float4 &b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT, ~1 ulp, no slow path
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, ~1 ulp, no slow path
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, time;
};

struct Hit {
    float t;        // distance along the UNIT direction (see normalise_ray), not the reference's parameter
    uint32_t slot;  // index into the prim array that was searched; kMiss = no hit
};

// Rays are TRACED AND SHADED with a unit direction u = d/|d| and distances s = t |d|.  The reference leaves d
// un-normalised (main.zig:94, material.zig:45-48) and its t_min = 0.001 is in units of THAT parameter (SURVEY App. B
// Q18), so the closest-hit searches normalise the ray in place when they start and scale t_min by |d|; with u.u = 1 the
// quadratic's `a` drops out of every sphere and bound test, and Vec3.normalized (vec.zig:33-40) out of metal and dielectric
// scattering.  Hit points are o + s u = o + t d.  Only the probes convert back (t = s * rl) for comparison with the oracle.
// Returns |d|; rl = 1/|d| (MUFU.RSQ, ~2 ulp: |u|^2 = 1 +- 3e-7, covered by the slack of every culling test).
__device__ __forceinline__ float normalise_ray(Ray &r, float &rl) {
    const float add = fmaf(r.dx, r.dx, fmaf(r.dy, r.dy, r.dz * r.dz));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rl) : "f"(add));
    r.dx *= rl; r.dy *= rl; r.dz *= rl;
    return add * rl;
}

template <bool STATS>
struct Counters {
    __device__ __forceinline__ void add(int, uint32_t = 1) {}
    __device__ __forceinline__ void flush(unsigned long long *) {}
};
template <>
struct Counters<true> {
    uint32_t c[ST_COUNT];
    __device__ Counters() {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) c[i] = 0;
    }
    __device__ __forceinline__ void add(int slot, uint32_t n = 1) { c[slot] += n; }
    __device__ void flush(unsigned long long *g) {
#pragma unroll
        for (int i = 0; i < ST_COUNT; ++i) {
            unsigned long long v = c[i];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(g + i, v);
            c[i] = 0;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Primitive tests on a unit-direction ray `u`; distances s in [s_min, s_max].
// Accept rule = the reference's: nearest root in the range, both ends inclusive (hittable.zig:106-116).
// ---------------------------------------------------------------------------------------------
// Sphere: numerically robust fp32 form.  The textbook b^2 - a*c of hittable.zig:96-101 cancels
// catastrophically in fp32 for small far spheres (measured: -0.5 % image bias from false
// self-intersections); instead the discriminant comes from the perpendicular offset l of the
// centre from the ray line (disc' = r^2 - |l|^2), and the roots from q = b' + sign(b') sqrt(disc'),
// s = {c/q, q}.  Same roots in exact arithmetic.
// discriminant of the robust form: returns disc', writes oc and b' = oc.u
__device__ __forceinline__ float sphere_disc(const Ray &u, float cx, float cy, float cz, float rr,
                                             float &ocx, float &ocy, float &ocz, float &bp) {
    ocx = u.ox - cx; ocy = u.oy - cy; ocz = u.oz - cz;
    bp = fmaf(ocx, u.dx, fmaf(ocy, u.dy, ocz * u.dz));
    const float lx = fmaf(-bp, u.dx, ocx), ly = fmaf(-bp, u.dy, ocy), lz = fmaf(-bp, u.dz, ocz);
    return fmaf(-lx, lx, fmaf(-ly, ly, fmaf(-lz, lz, rr)));
}
// roots from (c, b', disc'): nearest root in [s_min, s_max]
__device__ __forceinline__ bool sphere_root(float c, float bp, float discp, float s_min, float s_max, float &s_out) {
    const float sq = sqrt_approx(discp);
    const float bq = -bp;
    const float q = bq + copysignf(sq, bq);
    const float t0 = c * rcp_approx(q), t1 = q;
    float root = fminf(t0, t1);
    if (root < s_min || s_max < root) {
        root = fmaxf(t0, t1);
        if (root < s_min || s_max < root) return false;
    }
    if (!(root == root)) return false;  // q == 0 (ray through the centre of a zero-disc sphere), NaN disc'
    s_out = root;
    return true;
}
// |o - centre|^2 - r^2 of a big sphere about a reference point q on its surface: no 1e6 - 1e6 cancellation for the
// r = 1000 ground sphere of main.zig:172
__device__ __forceinline__ float big_sphere_c(const Ray &u, const DevBigSphere &g) {
    const float ax = u.ox - g.qx, ay = u.oy - g.qy, az = u.oz - g.qz;
    const float aa = fmaf(ax, ax, fmaf(ay, ay, az * az));
    const float am = fmaf(ax, g.mx, fmaf(ay, g.my, az * g.mz));
    return fmaf(2.0f, am, aa) + g.K;
}
// the BVH leaves' sphere test: the same operations in the same order as the flat scan's member tests, so both searches
// return bit-identical distances
template <bool STATS>
__device__ __forceinline__ bool sphere_test(const Ray &u, const DevPrim &p, const DevBigSphere *bigs, float s_min, float s_max,
                                            float &s_out, Counters<STATS> &cn) {
    cn.add(ST_SPHERE_TESTS);
    const float cx = fmaf(p.b.x, u.time, p.a.x), cy = fmaf(p.b.y, u.time, p.a.y), cz = fmaf(p.b.z, u.time, p.a.z);
    const float rr = p.a.w * p.a.w;
    float ocx, ocy, ocz, bp;
    const float discp = sphere_disc(u, cx, cy, cz, rr, ocx, ocy, ocz, bp);
    if (!(discp >= 0.0f)) return false;
    cn.add(ST_SPHERE_ROOTS);
    const uint32_t big = (__float_as_uint(p.b.w) >> 8) & 0xFFFu;
    const float c = big ? big_sphere_c(u, bigs[big - 1]) : fmaf(ocx, ocx, fmaf(ocy, ocy, fmaf(ocz, ocz, -rr)));
    return sphere_root(c, bp, discp, s_min, s_max, s_out);
}

// Axis-aligned rect in its object space.  kind: PK_XY (k on z), PK_XZ (k on y), PK_YZ (k on x).
__device__ __forceinline__ bool rect_test_os(uint32_t kind, float ox, float oy, float oz, float dx, float dy,
                                             float dz, const DevPrim &p, float t_min, float t_max, float &t_out) {
    float ok, dk, oa, da, ob, db;
    if (kind == PK_XY) { ok = oz; dk = dz; oa = ox; da = dx; ob = oy; db = dy; }
    else if (kind == PK_XZ) { ok = oy; dk = dy; oa = ox; da = dx; ob = oz; db = dz; }
    else { ok = ox; dk = dx; oa = oy; da = dy; ob = oz; db = dz; }
    const float t = __fdividef(p.b.x - ok, dk);  // hittable.zig:279 (2-ulp division: production arithmetic)
    if (t < t_min || t > t_max) return false;
    const float pa = fmaf(t, da, oa), pb = fmaf(t, db, ob);
    if (pa < p.a.x || pa > p.a.y || pb < p.a.z || pb > p.a.w) return false;
    t_out = t;
    return true;
}

template <bool STATS>
__device__ __forceinline__ bool rect_test(const Ray &r, const DevPrim &p, uint32_t kind, const DevXform *xforms,
                                          float t_min, float t_max, float &t_out, Counters<STATS> &cn) {
    cn.add(ST_RECT_TESTS);
    const int xi = __float_as_int(p.b.y);
    float ox = r.ox, oy = r.oy, oz = r.oz, dx = r.dx, dy = r.dy, dz = r.dz;
    if (xi >= 0) {  // world -> object (Translate.hit + RotateY.hit, hittable.zig:479-483, 560-573)
        cn.add(ST_XFORM_APPS);
        const DevXform x = xforms[xi];
        const float wx = ox, wz = oz, vx = dx, vz = dz;
        ox = fmaf(x.c, wx, -x.s * wz) + x.tx;
        oy = oy + x.ty;
        oz = fmaf(x.s, wx, x.c * wz) + x.tz;
        dx = fmaf(x.c, vx, -x.s * vz);
        dz = fmaf(x.s, vx, x.c * vz);
    }
    const bool h = rect_test_os(kind, ox, oy, oz, dx, dy, dz, p, t_min, t_max, t_out);
    if (h) cn.add(ST_RECT_ACCEPTS);
    return h;
}

template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ bool prim_test(const Ray &u, const DevPrim &p, const DevScene &sc, float s_min, float s_max,
                                          float &s_out, Counters<STATS> &cn) {
    const uint32_t kind = __float_as_uint(p.b.w) & 0xFFu;
    if (!(FEAT & FF_RECTS) || ((FEAT & FF_SPHERES) && kind == PK_SPHERE)) return sphere_test<STATS>(u, p, sc.bigs, s_min, s_max, s_out, cn);
    return rect_test<STATS>(u, p, kind, sc.xforms, s_min, s_max, s_out, cn);
}

// ---------------------------------------------------------------------------------------------
// Flat scan: the scene staged in shared memory as a FlatLayout blob.  Every lane reads the same
// record (smem broadcast), so the scan itself has no divergence; loop bodies are branch-free up to
// the (rare) accept path and unrolled for ILP.  Small spheres (static and moving together) sit in spatial
// groups of up to four behind a bounding sphere: each lane tests four bounds, ONE warp reduction (REDUX.OR)
// decides which of the four groups anybody needs, and the whole warp tests only those groups' members —
// culling without divergence.  The ground and the few spheres that dwarf the rest are tested first, individually.
// ALL 32 lanes must call this (inactive lanes pass active = false): it contains warp reductions.
// Grouping changes the visiting order, so the reference's tie rule ("later list element wins",
// hittable.zig:235-242) is applied explicitly: a candidate replaces the current hit if t < best, or
// t == best and its prim id is larger.
// ---------------------------------------------------------------------------------------------
struct FlatBest {
    float t;
    uint32_t id;
};

// Candidates reach this with t <= b.t.  Ties go to the larger prim id; compared as SIGNED integers so that "no hit yet"
// (kMiss = -1) loses against every id without a test of its own (prim ids stay below 2^28, the node-reference limit).
__device__ __forceinline__ bool flat_better(const FlatBest &b, float t, uint32_t id) {
    return t < b.t || (int32_t)id > (int32_t)b.id;
}
__device__ __forceinline__ void flat_consider(FlatBest &b, float t, uint32_t id) {
    if (flat_better(b, t, id)) { b.t = t; b.id = id; }
}

// does the ray (s >= 0) possibly touch the bounding sphere?  line test + "entirely behind" test
// `reach` = distance of the current closest hit: the bound is also skipped when even its nearest point lies beyond it
// (the big spheres — the ground — are tested first so that this culls)
// b = (centre, R): the radius itself is stored (R^2 is one multiply away, R would be a MUFU)
// This is a CULLING test, so it uses the cheap form |oc|^2 - (oc.u)^2 <= R^2 instead of the cancellation-free
// perpendicular offset the member tests need, and pays for the fp32 cancellation with slack: the left side carries an
// absolute error below 16 * 2^-24 * |oc|^2 (three-term dot products, one FMA rounding, |u|^2 = 1 +- 3e-7), so the bound
// is kept whenever lhs <= R^2 + 2e-6 |oc|^2.  False positives only cost time.
__device__ __forceinline__ bool bound_hit(const Ray &u, const float4 b, float reach) {
    const float ox = u.ox - b.x, oy = u.oy - b.y, oz = u.oz - b.z;
    const float rr = b.w * b.w;
    const float oo = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
    const float bp = fmaf(ox, u.dx, fmaf(oy, u.dy, oz * u.dz));
    const float lhs = fmaf(-bp, bp, oo);
    const float rhs = fmaf(oo, 2e-6f, rr);
    // origin outside the bound and either the bound is behind, or |oc| - R > reach, i.e. |oc|^2 > (reach + R)^2
    const float far = reach + b.w, lim = far * far;
    // bitwise, not short-circuit: the test stays one straight line of predicate logic (lim >= rr, so oo > lim implies
    // the origin is outside).  A NaN radius (padding entries of the bounds array) fails lhs <= rhs.
    return (lhs <= rhs) & !((oo > lim) | ((oo > rr) & (bp > 0.0f)));
}
// the same test, OR-ing `bit` into `mask` when it passes: four chained FSETP and one predicated LOP3 (left to itself
// the compiler materialises the predicates with three SEL per bound)
__device__ __forceinline__ void bound_hit_into(const Ray &u, const float4 b, float reach, uint32_t &mask, uint32_t bit) {
    const float ox = u.ox - b.x, oy = u.oy - b.y, oz = u.oz - b.z;
    const float rr = b.w * b.w;
    const float oo = fmaf(ox, ox, fmaf(oy, oy, oz * oz));
    const float bp = fmaf(ox, u.dx, fmaf(oy, u.dy, oz * u.dz));
    const float lhs = fmaf(-bp, bp, oo);
    const float rhs = fmaf(oo, 2e-6f, rr);
    const float far = reach + b.w, lim = far * far;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %1, 0f00000000;\n\t"
        "setp.gt.and.f32 p, %2, %3, p;\n\t"
        "setp.gt.or.f32 p, %2, %4, p;\n\t"
        "setp.le.and.f32 p, %5, %6, !p;\n\t"
        "@p or.b32 %0, %0, %7;\n\t}"
        : "+r"(mask) : "f"(bp), "f"(oo), "f"(rr), "f"(lim), "f"(lhs), "f"(rhs), "r"(bit));
}

template <bool STATS>
__device__ __forceinline__ void flat_static_pair(const Ray &u, const float4 a0, const float4 a1,
                                                 uint32_t id0, uint32_t id1, bool active, float s_min, FlatBest &best,
                                                 Counters<STATS> &cn) {
    float o0x, o0y, o0z, b0, o1x, o1y, o1z, b1;
    const float d0 = sphere_disc(u, a0.x, a0.y, a0.z, a0.w, o0x, o0y, o0z, b0);
    const float d1 = sphere_disc(u, a1.x, a1.y, a1.z, a1.w, o1x, o1y, o1z, b1);
    if (active && d0 >= 0.0f) {
        cn.add(ST_SPHERE_ROOTS);
        const float c = fmaf(o0x, o0x, fmaf(o0y, o0y, fmaf(o0z, o0z, -a0.w)));
        float t;
        if (sphere_root(c, b0, d0, s_min, best.t, t)) flat_consider(best, t, id0);
    }
    if (active && d1 >= 0.0f) {
        cn.add(ST_SPHERE_ROOTS);
        const float c = fmaf(o1x, o1x, fmaf(o1y, o1y, fmaf(o1z, o1z, -a1.w)));
        float t;
        if (sphere_root(c, b1, d1, s_min, best.t, t)) flat_consider(best, t, id1);
    }
}

// A run of rects sharing orientation and instance transform: t = (k - o_k)/d_k (hittable.zig:279), in-plane
// bounds inclusive (hittable.zig:283), range [t_min, best] inclusive, ties to the larger prim id.
template <bool STATS>
__device__ __forceinline__ void rect_run(float ok, float dk, float oa, float da, float ob, float db, const float4 *rp,
                                         uint32_t count, bool active, float t_min, FlatBest &best, Counters<STATS> &cn) {
    const float inv = rcp_approx(dk);  // dk == 0: t = +-inf or NaN
#pragma unroll 2
    for (uint32_t i = 0; i < count; ++i) {
        const float4 a = rp[2 * i], b = rp[2 * i + 1];
        const float t = (b.x - ok) * inv;
        const float pa = fmaf(t, da, oa), pb = fmaf(t, db, ob);
        // rejection written exactly as the reference does (hittable.zig:280,285): a NaN t (origin on the plane of a
        // parallel ray) fails every comparison and is therefore NOT rejected there either
        const bool hit = active && !(t < t_min || t > best.t || pa < a.x || pa > a.y || pb < a.z || pb > a.w);
        const uint32_t id = __float_as_uint(b.y);
        if (STATS) { if (hit) cn.add(ST_RECT_ACCEPTS); }
        const bool take = hit && flat_better(best, t, id);
        best.t = take ? t : best.t;
        best.id = take ? id : best.id;
    }
}

// `r` comes in with any direction and leaves with the unit direction (normalise_ray); the returned Hit::t is the distance
// along it.  `rl_out` (optional) receives 1/|d| for callers that report the reference's parameter t = s / |d|.
template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ Hit closest_hit_flat(Ray &r, bool active, const float4 *s, const FlatLayout &L,
                                                const DevScene &sc, float t_min, Counters<STATS> &cn, float *rl_out = nullptr) {
    float rl;
    t_min *= normalise_ray(r, rl);  // s_min = t_min |d|
    if (rl_out) *rl_out = rl;
    const Ray &u = r;
    // inactive lanes scan with reach 0: a bound then passes only if their (stale) origin lies inside it, so they
    // practically never keep a group alive in the votes below and the votes need no `active` term
    FlatBest best{active ? __int_as_float(0x7f800000) : 0.0f, kMiss};
    const uint32_t *ids = reinterpret_cast<const uint32_t *>(s + L.off_ids);

    // ---- big static spheres: c term about the reference point ----
    if constexpr ((FEAT & FF_SPHERES) != 0u) {
        const float4 *sp = s + L.off_big;
        for (uint32_t i = 0; i < L.n_big; ++i) {
            const float4 a0 = sp[i];
            float ox, oy, oz, bp;
            const float d0 = sphere_disc(u, a0.x, a0.y, a0.z, a0.w, ox, oy, oz, bp);
            if (active) cn.add(ST_SPHERE_TESTS);
            if (active && d0 >= 0.0f) {
                cn.add(ST_SPHERE_ROOTS);
                const float c = big_sphere_c(u, sc.bigs[i]);
                float t;
                if (sphere_root(c, bp, d0, t_min, best.t, t)) flat_consider(best, t, ids[i]);
            }
        }
    }
    // ---- sphere groups (static and moving spheres together, grouped by position), FOUR AT A TIME: the four bounds as straight-line code, each lane collecting
    //      its own pass bits; ONE warp reduction (REDUX.OR) turns them into the set of groups somebody needs; then only those
    //      groups' members are tested, by every lane.  Against a vote and a branch per group this is 8 fewer instructions
    //      per bound; the reach (distance of the closest hit so far) is refreshed between chunks.  Scenes with fewer than
    //      three groups skip the bounds (L.flags & kFlatNoBounds): the test costs what it saves there ----
    if constexpr ((FEAT & FF_SPHERES) != 0u) {
        const uint32_t ng = L.n_mov_groups;
        const float4 *bnd = s + L.off_bounds;  // padded to a multiple of four with NaN radii (never pass)
        const float4 *mov = s + L.off_mov;
        const uint32_t mov_ids = (L.n_big + 3u) & ~3u;  // ids: big (padded to x4) | four per group
#pragma unroll 1
        for (uint32_t base = 0; base < ng; base += 4u, bnd += 4) {
            const uint32_t nb = min(4u, ng - base);
            uint32_t m = (1u << nb) - 1u;
            if (!(L.flags & kFlatNoBounds)) {
                if (active) cn.add(ST_SPHERE_TESTS, nb);  // the bound is a sphere test too
                const float reach = best.t;
                m = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) bound_hit_into(u, bnd[k], reach, m, 1u << k);
                m = __reduce_or_sync(0xffffffffu, m);
            }
#pragma unroll 1
            while (m) {
                const uint32_t g = base + (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                if (active) cn.add(ST_SPHERE_TESTS, 4);
                // every member is a (centre, velocity) pair: centre(time) = cb + vel*time (hittable.zig:219-221), vel = 0 for
                // a static sphere (static and moving spheres share the groups: they are formed by position alone)
                const float4 *gp = mov + 8 * g;
                const uint4 id = *reinterpret_cast<const uint4 *>(ids + mov_ids + 4 * g);
                const float4 a0 = gp[0], v0 = gp[1], a1 = gp[2], v1 = gp[3], a2 = gp[4], v2 = gp[5], a3 = gp[6], v3 = gp[7];
                if (STATS) {
                    if (active) cn.add(ST_MOVING_TESTS, (v0.x != 0.f || v0.y != 0.f || v0.z != 0.f) + (v1.x != 0.f || v1.y != 0.f || v1.z != 0.f) +
                                                            (v2.x != 0.f || v2.y != 0.f || v2.z != 0.f) + (v3.x != 0.f || v3.y != 0.f || v3.z != 0.f));
                }
                const float4 c0 = make_float4(fmaf(v0.x, u.time, a0.x), fmaf(v0.y, u.time, a0.y), fmaf(v0.z, u.time, a0.z), a0.w);
                const float4 c1 = make_float4(fmaf(v1.x, u.time, a1.x), fmaf(v1.y, u.time, a1.y), fmaf(v1.z, u.time, a1.z), a1.w);
                const float4 c2 = make_float4(fmaf(v2.x, u.time, a2.x), fmaf(v2.y, u.time, a2.y), fmaf(v2.z, u.time, a2.z), a2.w);
                const float4 c3 = make_float4(fmaf(v3.x, u.time, a3.x), fmaf(v3.y, u.time, a3.y), fmaf(v3.z, u.time, a3.z), a3.w);
                flat_static_pair<STATS>(u, c0, c1, id.x, id.y, active, t_min, best, cn);
                flat_static_pair<STATS>(u, c2, c3, id.z, id.w, active, t_min, best, cn);
            }
        }
    }
    // ---- boxes: up to six rects that are the faces of one axis-aligned box of an instance's object space (the reference's
    //      Box, hittable.zig:429-470, and rooms): one transform, three slabs, six parameter comparisons ----
    if constexpr ((FEAT & FF_RECTS) != 0u) {
        const float4 *bx = s + L.off_boxes;
        for (uint32_t q = 0; q < L.n_boxes; ++q, bx += kBoxF4) {
            const float4 A = bx[0], B = bx[1];
            const uint4 I0 = *reinterpret_cast<const uint4 *>(bx + 2);
            const float4 X = bx[3], T = bx[4];  // (id4, id5, cos, sin), (tx, ty, tz, -)
            const uint2 I1 = make_uint2(__float_as_uint(X.x), __float_as_uint(X.y));
            const uint32_t xf = __float_as_uint(B.z), mask = __float_as_uint(B.w);  // mask: which faces exist (event counters only)
            float ox = u.ox, oy = u.oy, oz = u.oz, dx = u.dx, dy = u.dy, dz = u.dz;
            if (xf) {  // world -> object (Translate.hit + RotateY.hit, hittable.zig:479-483, 560-573); the composed chain
                       // sits in the record itself (same floats as DevScene::xforms[xf - 1], which finalise_hit reads)
                if (active) cn.add(ST_XFORM_APPS);
                ox = fmaf(X.z, u.ox, -X.w * u.oz) + T.x; oy = u.oy + T.y; oz = fmaf(X.w, u.ox, X.z * u.oz) + T.z;
                dx = fmaf(X.z, u.dx, -X.w * u.dz); dz = fmaf(X.w, u.dx, X.z * u.dz);
            }
            if (active) cn.add(ST_RECT_TESTS, __popc(mask));
            const float ix = rcp_approx(dx), iy = rcp_approx(dy), iz = rcp_approx(dz);
            const float tx0 = (A.x - ox) * ix, tx1 = (A.y - ox) * ix;  // t = (k - o) / d, hittable.zig:279
            const float ty0 = (A.z - oy) * iy, ty1 = (A.w - oy) * iy;
            const float tz0 = (B.x - oz) * iz, tz1 = (B.y - oz) * iz;
            // fminf/fmaxf drop a NaN (origin on a plane of a parallel ray): that slab then does not constrain
            const float lox = fminf(tx0, tx1), hix = fmaxf(tx0, tx1);
            const float loy = fminf(ty0, ty1), hiy = fmaxf(ty0, ty1);
            const float loz = fminf(tz0, tz1), hiz = fmaxf(tz0, tz1);
            // A face's in-plane bounds test (hittable.zig:283-287) in parameters: the NEAR face of axis a (at t = lo_a) is
            // hit iff lo_a >= the other two lows and <= the other two highs, i.e. (lo_a <= hi_a always) iff
            // lo_a == t_en = max(lows) and t_en <= t_ex = min(highs); the FAR face iff hi_a == t_ex and t_en <= t_ex.  So of
            // the six faces only those AT t_en or t_ex can be hit, several of them on an edge or corner, where the list scan
            // keeps the later element (hittable.zig:235-242) = the largest prim id; absent faces (a room's open side) carry
            // id -1 and lose every signed comparison.  The box's closest hit is the t_en candidate when it is in range and
            // present, else the t_ex one: ONE candidate per box meets the running best instead of six.
            const float t_en = fmaxf(fmaxf(lox, loy), loz), t_ex = fminf(fminf(hix, hiy), hiz);
            const bool px = tx0 <= tx1, py = ty0 <= ty1, pz = tz0 <= tz1;  // near face = the "0" face of the axis
            const int32_t nx = (int32_t)(px ? I1.y : I1.x), fx = (int32_t)(px ? I1.x : I1.y);
            const int32_t ny = (int32_t)(py ? I0.w : I0.z), fy = (int32_t)(py ? I0.z : I0.w);
            const int32_t nz = (int32_t)(pz ? I0.y : I0.x), fz = (int32_t)(pz ? I0.x : I0.y);
            const int32_t id_ex = max(max(hix == t_ex ? fx : -1, hiy == t_ex ? fy : -1), hiz == t_ex ? fz : -1);
            // a ray that only touches an edge or corner (t_en == t_ex) meets near and far faces at the same t
            const int32_t id_en = max(max(max(lox == t_en ? nx : -1, loy == t_en ? ny : -1), loz == t_en ? nz : -1), t_en == t_ex ? id_ex : -1);
            const bool cross = active & (t_en <= t_ex);
            const bool en_ok = cross & (t_en >= t_min) & (t_en <= best.t) & (id_en >= 0);
            const bool ex_ok = cross & (t_ex >= t_min) & (t_ex <= best.t) & (id_ex >= 0);
            const float tb = en_ok ? t_en : t_ex;
            const uint32_t ib = (uint32_t)(en_ok ? id_en : id_ex);
            if (STATS) { if (en_ok | ex_ok) cn.add(ST_RECT_ACCEPTS); }
            const bool take = (en_ok | ex_ok) && flat_better(best, tb, ib);
            best.t = take ? tb : best.t;
            best.id = take ? ib : best.id;
        }
    }
    // ---- rects: runs of equal (instance transform, orientation); consecutive runs of one instance (a box = three
    //      runs) share the object-space ray, which is therefore set up once per instance, not once per run ----
    if constexpr ((FEAT & FF_RECTS) != 0u) {
        const uint4 *runs = reinterpret_cast<const uint4 *>(s + L.off_runs);
        float o[3] = {u.ox, u.oy, u.oz}, d[3] = {u.dx, u.dy, u.dz};
        for (uint32_t q = 0; q < L.n_runs; ++q) {
            const uint4 run = runs[q];
            if (!(run.x & kRunSameXform)) {
                o[0] = u.ox; o[1] = u.oy; o[2] = u.oz; d[0] = u.dx; d[1] = u.dy; d[2] = u.dz;
                if (run.x) {  // world -> object (Translate.hit + RotateY.hit, hittable.zig:479-483, 560-573)
                    if (active) cn.add(ST_XFORM_APPS);
                    const DevXform x = sc.xforms[run.x - 1u];
                    o[0] = fmaf(x.c, u.ox, -x.s * u.oz) + x.tx; o[1] = u.oy + x.ty; o[2] = fmaf(x.s, u.ox, x.c * u.oz) + x.tz;
                    d[0] = fmaf(x.c, u.dx, -x.s * u.dz); d[2] = fmaf(x.s, u.dx, x.c * u.dz);
                }
            }
            const float4 *rp = s + L.off_rect + 2 * run.z;
            if (active) cn.add(ST_RECT_TESTS, run.w);
            if (run.y == PK_XY) rect_run<STATS>(o[2], d[2], o[0], d[0], o[1], d[1], rp, run.w, active, t_min, best, cn);
            else if (run.y == PK_XZ) rect_run<STATS>(o[1], d[1], o[0], d[0], o[2], d[2], rp, run.w, active, t_min, best, cn);
            else rect_run<STATS>(o[0], d[0], o[1], d[1], o[2], d[2], rp, run.w, active, t_min, best, cn);
        }
    }
    return Hit{best.t, best.id};
}

// ---------------------------------------------------------------------------------------------
// BVH traversal: binary BVH, 32-byte nodes, siblings adjacent (one 64-byte aligned fetch per
// step), ordered descent, per-thread stack.  Slab test = aabb.zig:8-45 with 1/d precomputed and
// the far bound widened by 2 ulps so rounding can never cull a box holding an acceptable hit.
// Ties: a candidate replaces the current hit if t < best, or t == best and its prim id is larger
// (= "later list element wins", hittable.zig:235-242).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool slab(const BvhNode &n, const Ray &r, float idx, float idy, float idz, float t_min,
                                     float t_max, float &t_near) {
    const float x0 = (n.mnx - r.ox) * idx, x1 = (n.mxx - r.ox) * idx;
    const float y0 = (n.mny - r.oy) * idy, y1 = (n.mxy - r.oy) * idy;
    const float z0 = (n.mnz - r.oz) * idz, z1 = (n.mxz - r.oz) * idz;
    // fminf/fmaxf drop NaNs (0*inf when the origin lies on a slab plane of a parallel ray)
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), t_min));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), t_max));
    t_near = tn;
    return tn <= tf * 1.0000004f;
}

constexpr int kBvhStack = 64;

// The traversal stack: node references in local memory (the entries of one depth form a 128-byte line per warp).
// Measured and removed (profiles/r02_m): entry distances kept on the stack so that stale entries are dropped by the pop (only
// 1.6-3 % of the node tests are stale with the ordered descent; the second word per entry cost 10 % on the 10^6-sphere scene),
// and the first 16 / 24 entries in shared memory, one bank per lane (+3 % on 485 spheres, +10 % on 10^6: the address
// arithmetic and the L1 capacity given to the carve-out cost more than the local-memory lines did).
struct BvhStack {
    uint32_t ref[kBvhStack];
    __device__ __forceinline__ void put(int at, uint32_t v) { ref[at] = v; }
    __device__ __forceinline__ uint32_t get(int at) const { return ref[at]; }
};

// stack entry / node reference: index in the low 28 bits, leaf primitive count in the top 4.  The builders store
// BvhNode::a in exactly this form (kNodeRefShift), so a child reference is the loaded word itself.

// Traversal state of one ray.  `step` performs at most one interior visit (both children of the
// current node) followed by at most one leaf visit, so a warp can interleave traversal steps of its
// lanes with shading/regeneration of the lanes that are done (k_megakernel_bvh).
struct BvhTraversal {
    float idx, idy, idz, s_min;
    Hit h;
    uint32_t best_id;
    uint32_t cur;  // reference of the node to visit next
    uint32_t pend; // speculative schedule only: a postponed leaf (0 = none; leaf references are never 0)
    int sp;
    // The stack itself is NOT a member: a dynamically indexed array inside the struct forces the whole struct into
    // local memory (ncu r01_m: four LDL and three STL per interior visit); as a separate array only the pushes and
    // pops touch local memory and the scalars above stay in registers.

    // normalises `r` in place (normalise_ray); returns true when the traversal is already finished (empty scene)
    __device__ __forceinline__ bool init(Ray &r, const DevScene &sc, float t_min, float *rl_out = nullptr) {
        float rl;
        s_min = t_min * normalise_ray(r, rl);
        if (rl_out) *rl_out = rl;
        idx = rcp_approx(r.dx); idy = rcp_approx(r.dy); idz = rcp_approx(r.dz);
        h = Hit{__int_as_float(0x7f800000), kMiss};
        best_id = 0;
        sp = 0;
        pend = 0u;
        if (sc.n_prims == 0) return true;
        cur = sc.nodes[0].a;
        return false;
    }

    __device__ __forceinline__ bool at_leaf() const { return (cur >> 28) != 0u; }

    // one interior visit: both children of `cur` (precondition: !at_leaf()).  Returns true when finished.
    template <bool STATS, class Stack>
    __device__ __forceinline__ bool interior_step(const Ray &r, const DevScene &sc, Stack &stack,
                                                  Counters<STATS> &cn) {
        const float4 *q = reinterpret_cast<const float4 *>(sc.nodes + cur);
        float4 l0, l1, r0, r1;
        ldg256(q, l0, l1);
        ldg256(q + 2, r0, r1);
        const BvhNode L{l0.x, l0.y, l0.z, __float_as_uint(l0.w), l1.x, l1.y, l1.z, __float_as_uint(l1.w)};
        const BvhNode R{r0.x, r0.y, r0.z, __float_as_uint(r0.w), r1.x, r1.y, r1.z, __float_as_uint(r1.w)};
        cn.add(ST_NODE_TESTS, 2);
        float tl, tr;
        const bool hl = slab(L, r, idx, idy, idz, s_min, h.t, tl);
        const bool hr = slab(R, r, idx, idy, idz, s_min, h.t, tr);
        const uint32_t el = L.a, er = R.a;
        if (hl && hr) {
            const bool left_first = tl <= tr;
            stack.put(sp++, left_first ? er : el);  // depth is bounded by the builder (<= kBvhStack)
            cur = left_first ? el : er;
        } else if (hl || hr) {
            cur = hl ? el : er;
        } else {
            return !pop(stack);
        }
        return false;
    }

    // next node to visit from the stack; false when there is none (the traversal is finished)
    template <class Stack>
    __device__ __forceinline__ bool pop(Stack &stack) {
        if (sp == 0) return false;
        cur = stack.get(--sp);
        return true;
    }

    // the primitives of leaf reference `leaf`: the only primitive-test site
    template <bool STATS, uint32_t FEAT>
    __device__ __forceinline__ void test_leaf(uint32_t leaf, const Ray &r, const DevScene &sc, Counters<STATS> &cn) {
        const uint32_t cnt = leaf >> 28, at = leaf & 0x0FFFFFFFu;
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t slot = at + i;
            DevPrim p;
            ldg256(sc.prims_bvh + slot, p.a, p.b);
            float t;
            if (prim_test<STATS, FEAT>(r, p, sc, s_min, h.t, t, cn)) {
                const uint32_t id = __ldg(sc.bvh_prim_id + slot);
                if (t < h.t || h.slot == kMiss || id > best_id) { h.t = t; h.slot = slot; best_id = id; }
            }
        }
    }

    // one leaf visit (precondition: at_leaf()).  Returns true when finished.
    template <bool STATS, uint32_t FEAT = FF_ALL, class Stack>
    __device__ __forceinline__ bool leaf_step(const Ray &r, const DevScene &sc, Stack &stack,
                                              Counters<STATS> &cn) {
        test_leaf<STATS, FEAT>(cur, r, sc, cn);
        return !pop(stack);
    }

    // ---- speculative schedule (Aila & Laine 2009, "speculative traversal"): a lane that reaches a leaf POSTPONES it (`pend`)
    //      and goes on with the next stack entry instead of idling until enough lanes have a leaf to test.  The closest hit does
    //      not depend on the order of the visits (the tie rule is symmetric); a postponed leaf only delays the shrinking of h.t,
    //      i.e. a few node tests more.  Invariant: `cur` is always a node still to visit; pend != 0 is a second one (a leaf).
    template <class Stack>
    __device__ __forceinline__ void postpone(Stack &stack) {
        if (at_leaf() && pend == 0u && sp > 0) {
            const uint32_t leaf = cur;
            if (pop(stack)) pend = leaf;  // nothing else worth a visit: stay parked at the leaf
            else cur = leaf;
        }
    }
    template <bool STATS, class Stack>
    __device__ __forceinline__ bool spec_interior_step(const Ray &r, const DevScene &sc, Stack &stack,
                                                       Counters<STATS> &cn) {
        if (interior_step<STATS>(r, sc, stack, cn)) {  // nothing left but the postponed leaf
            if (pend == 0u) return true;
            cur = pend; pend = 0u;
            return false;
        }
        postpone(stack);
        return false;
    }
    // precondition: pend != 0 || at_leaf()
    template <bool STATS, uint32_t FEAT = FF_ALL, class Stack>
    __device__ __forceinline__ bool spec_leaf_step(const Ray &r, const DevScene &sc, Stack &stack,
                                                   Counters<STATS> &cn) {
        const bool from_pend = pend != 0u;
        test_leaf<STATS, FEAT>(from_pend ? pend : cur, r, sc, cn);
        if (from_pend) pend = 0u;
        else if (!pop(stack)) return true;
        postpone(stack);
        return false;
    }

    // at most one interior visit followed by at most one leaf visit.  Returns true when finished.
    template <bool STATS, class Stack>
    __device__ __forceinline__ bool step(const Ray &r, const DevScene &sc, Stack &stack, Counters<STATS> &cn) {
        if (!at_leaf() && interior_step<STATS>(r, sc, stack, cn)) return true;
        if (at_leaf()) return leaf_step<STATS>(r, sc, stack, cn);
        return false;
    }
};

// `r` leaves with the unit direction, Hit::t is the distance along it (see closest_hit_flat)
template <bool STATS>
__device__ __forceinline__ Hit closest_hit_bvh(Ray &r, const DevScene &sc, float t_min, Counters<STATS> &cn, float *rl_out = nullptr) {
    BvhTraversal tv;
    BvhStack stack;
    if (tv.init(r, sc, t_min, rl_out)) return tv.h;
    while (!tv.template step<STATS>(r, sc, stack, cn)) {}
    return tv.h;
}

// ---------------------------------------------------------------------------------------------
// Hit record of the winning primitive (computed once per ray, not per candidate as the reference
// does at hittable.zig:118-128).
// ---------------------------------------------------------------------------------------------
// getSphereUv (hittable.zig:145-150) needs atan2 and acos of the outward normal: CUDA's atan2f + acosf were 9 % of all
// warp instructions of the image-texture config at 9.4 lanes (profiles/r02_c).  Polynomials fitted on [0, 1] (least squares on
// Chebyshev nodes, errors measured over 2*10^6 fp32 arguments): |atan error| <= 1.8e-7 rad, |acos error| <= 3.5e-7 rad, i.e.
// u, v to 6e-8 / 1.2e-7 — 2.4e-4 of a texel of the 2048-wide asset.
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.0f ? mn * rcp_approx(mx) : 0.0f;  // atan2(0, 0) = 0 like libm
    const float z = a * a;
    float p = -0.0048311310820281506f;
    p = fmaf(p, z, 0.024756666272878647f); p = fmaf(p, z, -0.060218989849090576f); p = fmaf(p, z, 0.09967915713787079f);
    p = fmaf(p, z, -0.14040136337280273f); p = fmaf(p, z, 0.1997368186712265f); p = fmaf(p, z, -0.33332303166389465f);
    p = fmaf(p, z, 0.9999999403953552f);
    float r = p * a;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.0f ? 3.14159265358979323846f - r : r;
    return copysignf(r, y);
}
__device__ __forceinline__ float fast_acos(float x) {  // x in [-1, 1]
    const float ax = fabsf(x);
    float p = 0.0022513726726174355f;
    p = fmaf(p, ax, -0.011012407019734383f); p = fmaf(p, ax, 0.026749366894364357f); p = fmaf(p, ax, -0.048724427819252014f);
    p = fmaf(p, ax, 0.08873733878135681f); p = fmaf(p, ax, -0.21458369493484497f); p = fmaf(p, ax, 1.5707961320877075f);
    const float r = sqrt_approx(1.0f - ax) * p;
    return x < 0.0f ? 3.14159265358979323846f - r : r;
}
// u, v of getSphereUv for the outward unit normal n (hittable.zig:145-150)
__device__ __forceinline__ void sphere_uv(float nx, float ny, float nz, float &u, float &v) {
    u = (fast_atan2(-nz, nx) + 3.14159265358979323846f) * 0.15915494309189533577f;
    v = fast_acos(fminf(fmaxf(-ny, -1.0f), 1.0f)) * 0.31830988618379067154f;
}

struct Surface {
    float px, py, pz;     // hit point, world space
    float nx, ny, nz;     // face-corrected normal (HitRecord.normal)
    float onx, ony, onz;  // outward normal (for sphere uv, hittable.zig:127)
    float u, v;           // sphere (instanced): final uv.  rect: numerators of uv, divided lazily by the image texture
    float ru, rv;         // rect: denominators of uv (1 otherwise)
    bool front_face;
    bool is_sphere;       // plain sphere: uv = getSphereUv(outward normal), computed lazily by the image texture
};

template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ Surface finalise_hit(const Ray &r, const DevPrim &p, float t, const DevScene &sc,
                                                Counters<STATS> &cn) {
    Surface s;
    const uint32_t kind = __float_as_uint(p.b.w) & 0xFFu;
    s.px = fmaf(t, r.dx, r.ox); s.py = fmaf(t, r.dy, r.oy); s.pz = fmaf(t, r.dz, r.oz);  // Ray.at ray.zig:10-12
    if (!(FEAT & FF_RECTS) || ((FEAT & FF_SPHERES) && kind == PK_SPHERE)) {
        cn.add(ST_SPHERE_FINAL);
        const float cx = fmaf(p.b.x, r.time, p.a.x), cy = fmaf(p.b.y, r.time, p.a.y), cz = fmaf(p.b.z, r.time, p.a.z);
        const float inv_r = rcp_approx(p.a.w);
        s.onx = (s.px - cx) * inv_r; s.ony = (s.py - cy) * inv_r; s.onz = (s.pz - cz) * inv_r;
        s.is_sphere = true;
        s.u = 0.0f; s.v = 0.0f; s.ru = 1.0f; s.rv = 1.0f;
        const uint32_t xf = __float_as_uint(p.b.w) >> 20;
        if ((FEAT & FF_TEX) && xf) {  // instanced sphere: getSphereUv sees the OBJECT-space normal (hittable.zig:127 inside Translate/RotateY)
            const DevXform x = sc.xforms[xf - 1];
            const float ux = fmaf(x.c, s.onx, -x.s * s.onz), uz = fmaf(x.s, s.onx, x.c * s.onz);
            sphere_uv(ux, s.ony, uz, s.u, s.v);
            s.is_sphere = false;  // uv already final
        }
    } else {
        // object-space rect: outward normal is the +axis (hittable.zig:295-301), uv = normalised
        // in-plane coordinates (hittable.zig:288-289)
        const int xi = __float_as_int(p.b.y);
        float ox = r.ox, oy = r.oy, oz = r.oz, dx = r.dx, dy = r.dy, dz = r.dz;
        DevXform x{1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0, 0, 0};
        if (xi >= 0) {
            x = sc.xforms[xi];
            const float wx = ox, wz = oz, vx = dx, vz = dz;
            ox = fmaf(x.c, wx, -x.s * wz) + x.tx; oy = oy + x.ty; oz = fmaf(x.s, wx, x.c * wz) + x.tz;
            dx = fmaf(x.c, vx, -x.s * vz); dz = fmaf(x.s, vx, x.c * vz);
        }
        const float qx = fmaf(t, dx, ox), qy = fmaf(t, dy, oy), qz = fmaf(t, dz, oz);
        float pa, pb, nxo = 0.0f, nyo = 0.0f, nzo = 0.0f;
        if (kind == PK_XY) { pa = qx; pb = qy; nzo = 1.0f; }
        else if (kind == PK_XZ) { pa = qx; pb = qz; nyo = 1.0f; }
        else { pa = qy; pb = qz; nxo = 1.0f; }
        // (x - x0) / (x1 - x0) hittable.zig:289-290, on the exact bounds: the record's are widened by `slack` (rtw_api.cpp
        // rect_slack).  Kept un-divided: only an image texture ever reads u, v (texture_value divides there).
        s.u = pa - (p.a.x + p.b.z); s.v = pb - (p.a.z + p.b.z);
        s.ru = (p.a.y - p.a.x) - 2.0f * p.b.z; s.rv = (p.a.w - p.a.z) - 2.0f * p.b.z;
        // object -> world for the normal (RotateY.hit hittable.zig:588-590): n = A^T n_obj
        s.onx = fmaf(x.c, nxo, x.s * nzo);
        s.ony = nyo;
        s.onz = fmaf(-x.s, nxo, x.c * nzo);
        s.is_sphere = false;
    }
    s.front_face = fmaf(s.onx, r.dx, fmaf(s.ony, r.dy, s.onz * r.dz)) < 0.0f;
    const float sg = s.front_face ? 1.0f : -1.0f;
    s.nx = sg * s.onx; s.ny = sg * s.ony; s.nz = sg * s.onz;
    return s;
}

// ---------------------------------------------------------------------------------------------
// Textures — Texture.value texture.zig:36-145
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float perlin_noise(const DevPerlin &pn, float x, float y, float z) {  // perlin.zig:47-77,103-124
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const float u = x - fx, v = y - fy, w = z - fz;
    const float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const uint32_t idx = pn.perm[0][(i + di) & 255] ^ pn.perm[1][(j + dj) & 255] ^ pn.perm[2][(k + dk) & 255];
                const float4 c = pn.ranvec[idx];
                const float wx = di ? uu : 1.0f - uu, wy = dj ? vv : 1.0f - vv, wz = dk ? ww : 1.0f - ww;
                // the weight vector uses the SMOOTHED coordinates: perlin.zig:77 passes u_, v_, w_ to perlinInterp (:114)
                accum += wx * wy * wz * (c.x * (uu - di) + c.y * (vv - dj) + c.z * (ww - dk));
            }
    return accum;
}
__device__ __forceinline__ float perlin_turb(const DevPerlin &pn, float x, float y, float z) {  // perlin.zig:79-91, depth 7
    float accum = 0.0f, weight = 1.0f;
#pragma unroll 1
    for (int i = 0; i < 7; ++i) {
        accum += weight * perlin_noise(pn, x, y, z);
        weight *= 0.5f;
        x *= 2.0f; y *= 2.0f; z *= 2.0f;
    }
    return fabsf(accum);
}

template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ float3 texture_value(const DevScene &sc, int ti, const Surface &s, Counters<STATS> &cn) {
    DevTexture tx = sc.textures[ti];
    // Checker (texture.zig:79-82): sign(sin(10x) sin(10y) sin(10z)) < 0 -> odd.  sin(10 x) < 0 iff
    // floor(10 x / pi) is odd, so the sign of the product is the parity of the three floors: the
    // same function except on the measure-zero cell boundaries, without three sin evaluations.
    // Children are textures themselves (texture.zig:59-60): loop handles nesting.
    for (int guard = 0; tx.kind == 1u && guard < 8; ++guard) {
        cn.add(ST_TEX_CHECKER);
        const float k = 3.18309886183790672f;  // 10/pi
        const int par = (int)floorf(s.px * k) + (int)floorf(s.py * k) + (int)floorf(s.pz * k);
        tx = sc.textures[(par & 1) ? tx.a : tx.b];
    }
    if (!(FEAT & FF_TEX) || tx.kind <= 1u) return make_float3(tx.r, tx.g, tx.bl);  // solid texture.zig:46-55 (a checker still here = nesting beyond the guard: refused at upload)
    if (tx.kind == 2u) {  // noise texture.zig:100-104
        cn.add(ST_TEX_NOISE);
        const float v = 0.5f * (1.0f + sinf(tx.scale * s.pz + 10.0f * perlin_turb(sc.perlins[tx.a], s.px, s.py, s.pz)));
        return make_float3(v, v, v);
    }
    // image texture.zig:121-144 — nearest texel by truncation, alpha==0 -> (0,0,1)
    cn.add(ST_TEX_IMAGE);
    float u = __fdividef(s.u, s.ru), v = __fdividef(s.v, s.rv);
    if (s.is_sphere) sphere_uv(s.onx, s.ony, s.onz, u, v);  // getSphereUv hittable.zig:145-150
    const DevImage im = sc.images[tx.a];
    const float uc = fminf(fmaxf(u, 0.0f), 1.0f);
    const float vc = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
    int i = (int)(uc * (float)im.w), j = (int)(vc * (float)im.h);
    i = min(i, (int)im.w - 1);
    j = min(j, (int)im.h - 1);  // reference clamps with width-1 (texture.zig:130, OOB bug): fixed
    const uchar4 px = tex2D<uchar4>(im.tex, (float)i + 0.5f, (float)j + 0.5f);
    if (px.w == 0) return make_float3(0.0f, 0.0f, 1.0f);
    const float cs = 1.0f / 255.0f;
    return make_float3(cs * px.x, cs * px.y, cs * px.z);
}

// ---------------------------------------------------------------------------------------------
// Samplers: same distributions as rand.zig:22-40, rejection-free (no divergent loops).
// ---------------------------------------------------------------------------------------------
// One Philox block per ray: block 0 of a (pixel, sample) feeds the camera ray, block b >= 1 the scatter at bounce b.
// Both kinds of ray turn the block into the SAME intermediate — an angle (sin, cos of 2 pi ub) and a radius
// (sqrt(ua) for the lens disk, sqrt(1 - z^2) with z = 1 - 2 ua for the unit vector) — so a warp whose lanes are partly
// starting paths and partly scattering runs this once for all 32 lanes (k_megakernel_flat).
//   ua, ub  lens-disk point (camera) | unit vector (diffuse, metal fuzz direction) | ua = the dielectric's uniform
//   uc      pixel jitter u (camera)  | fuzz radius^3 (metal)
//   ud      pixel jitter v (camera)
//   ue      shutter time (camera): the low bytes of three words — bits no other variate uses
struct Draw {
    float ua, uc, ud, ue;
    float x, y, z;  // camera: (x, y) = point in the unit disk; scatter: (x, y, z) = unit vector
};
// the maps themselves, from uniforms in [0,1):
//   camera  (x, y) = sqrt(ua) (cos, sin)(2 pi ub): uniform in the unit disk          = randomPointInUnitDisk rand.zig:30-36
//   else    z = 1 - 2 ua, (x, y) = sqrt(1 - z^2) (cos, sin)(2 pi ub): uniform on the sphere = randomUnitVector rand.zig:38-40
//           and cbrt(uc) (x, y, z): uniform in the unit ball                             = randomPointInUnitSphere rand.zig:22-28
__device__ __forceinline__ Draw draw_from(float ua, float ub, float uc, float ud, float ue, bool camera) {
    Draw d;
    d.ua = ua; d.uc = uc; d.ud = ud; d.ue = ue;
    d.z = 1.0f - 2.0f * d.ua;
    const float rad = sqrt_approx(camera ? d.ua : fmaxf(0.0f, fmaf(-d.z, d.z, 1.0f)));
    float sn, cs;
    __sincosf(6.28318530717958647692f * ub, &sn, &cs);
    d.x = rad * cs; d.y = rad * sn;
    return d;
}
__device__ __forceinline__ Draw make_draw(const uint4 rn, bool camera) {
    return draw_from(u01_24(rn.x), u01_24(rn.y), u01_24(rn.z), u01_24(rn.w),
                     u01_24((rn.x << 24) | ((rn.y & 0xFFu) << 16) | ((rn.z & 0xFFu) << 8)), camera);
}

// The five uniforms of one camera ray in the order (ju, jv, l1, l2, tm): pixel jitter, lens disk, shutter time.
__device__ __forceinline__ void camera_uniforms(const DevRender &rp, uint32_t pixel, uint32_t sample, float (&u)[5]) {
    const uint4 rn = philox4x32_10(make_uint4(pixel, sample, 0u, 0u), rp.philox_keys);
    u[0] = u01_24(rn.z); u[1] = u01_24(rn.w); u[2] = u01_24(rn.x); u[3] = u01_24(rn.y);
    u[4] = u01_24((rn.x << 24) | ((rn.y & 0xFFu) << 16) | ((rn.z & 0xFFu) << 8));
}

// Camera.getRay main.zig:91-100 + the (u,v) jitter of main.zig:390-391, from a drawn block.
__device__ __forceinline__ Ray camera_from_draw(const DevCamera &cam, const DevRender &rp, uint32_t i, uint32_t j, const Draw &dw) {
    const float s = ((float)i + dw.uc) * rp.inv_wm1;  // main.zig:390-391 (division by W-1 as a multiply)
    const float t = ((float)j + dw.ud) * rp.inv_hm1;
    const float rdx = dw.x * cam.lens_radius, rdy = dw.y * cam.lens_radius;
    const float offx = cam.ux * rdx + cam.wx * rdy, offy = cam.uy * rdx + cam.wy * rdy, offz = cam.uz * rdx + cam.wz * rdy;
    Ray r;
    r.ox = cam.ox + offx; r.oy = cam.oy + offy; r.oz = cam.oz + offz;
    r.dx = fmaf(cam.vx, t, fmaf(cam.hx, s, cam.lx)) - cam.ox - offx;
    r.dy = fmaf(cam.vy, t, fmaf(cam.hy, s, cam.ly)) - cam.oy - offy;
    r.dz = fmaf(cam.vz, t, fmaf(cam.hz, s, cam.lz)) - cam.oz - offz;
    r.time = fmaf(dw.ue, cam.time1 - cam.time0, cam.time0);
    return r;
}
__device__ __forceinline__ Ray camera_ray(const DevCamera &cam, const DevRender &rp, uint32_t pixel, uint32_t i,
                                          uint32_t j, uint32_t sample) {
    const uint4 rn = philox4x32_10(make_uint4(pixel, sample, 0u, 0u), rp.philox_keys);
    return camera_from_draw(cam, rp, i, j, make_draw(rn, true));
}

// ---------------------------------------------------------------------------------------------
// One level of rayColor (main.zig:103-122) made iterative: `beta` is the product of attenuations so far, `L` the
// radiance gathered.  Split around the random draw:
//   shade_prepare  everything Material.emitted / scatter decide WITHOUT randomness: hit record, emitted light, the
//                  attenuation (texture lookup), the metal absorb test, the dielectric's two candidate directions.
//                  Leaves the next ray half-built in `r` (origin = hit point, direction = the deterministic part) and
//                  the rest in `Pending`.  Returns false when the path ends here.
//   scatter_finish adds the drawn part: n + unit vector | reflected + fuzz * ball point | reflect-or-refract by xi.
// ---------------------------------------------------------------------------------------------
struct Pending {
    float bx, by, bz;  // dielectric: the refracted direction (r.d holds the reflected one)
    float s;           // metal: fuzz; dielectric: Schlick reflectance (1 when refraction is impossible: xi < 1 never exceeds it)
    uint32_t kind;     // RTW_MAT_* of the surface being left
};

template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ bool shade_prepare(const DevScene &sc, Ray &r, const DevPrim &prim, uint32_t prim_id, float t,
                                              float3 &beta, float3 &L, Pending &pd, Counters<STATS> &cn) {
    const Surface s = finalise_hit<STATS, FEAT>(r, prim, t, sc, cn);
    const DevMaterial m = sc.materials[sc.prim_material[prim_id]];
    pd.kind = m.kind;
    if (m.kind == 3u) {  // diffuse_light: emitted on both faces, never scatters (material.zig:97-109)
        cn.add(ST_EMIT);
        const float3 e = texture_value<STATS, FEAT>(sc, m.tex, s, cn);
        L.x = fmaf(beta.x, e.x, L.x); L.y = fmaf(beta.y, e.y, L.y); L.z = fmaf(beta.z, e.z, L.z);
        return false;
    }
    if (m.kind == 0u) {  // diffuse material.zig:44-52: direction = normal + unit vector
        cn.add(ST_SC_DIFFUSE);
        const float3 a = texture_value<STATS, FEAT>(sc, m.tex, s, cn);
        beta.x *= a.x; beta.y *= a.y; beta.z *= a.z;
        r.dx = s.nx; r.dy = s.ny; r.dz = s.nz;
    } else {
        const float ux = r.dx, uy = r.dy, uz = r.dz;  // Vec3.normalized vec.zig:33-40: the search left r.d a unit vector
        const float udn = fmaf(ux, s.nx, fmaf(uy, s.ny, uz * s.nz));
        const float rx = fmaf(-2.0f * udn, s.nx, ux), ry = fmaf(-2.0f * udn, s.ny, uy), rz = fmaf(-2.0f * udn, s.nz, uz);  // reflect material.zig:112-114
        r.dx = rx; r.dy = ry; r.dz = rz;
        if (m.kind == 1u) {  // metal material.zig:59-65
            cn.add(ST_SC_METAL);
            pd.s = m.param;
            beta.x *= m.r; beta.y *= m.g; beta.z *= m.b;
            // absorbed iff the UN-fuzzed reflection points into the surface (material.zig:64)
            if (!(fmaf(rx, s.nx, fmaf(ry, s.ny, rz * s.nz)) > 0.0f)) return false;
        } else {  // dielectric material.zig:72-85
            cn.add(ST_SC_DIELECTRIC);
            const float ratio = s.front_face ? rcp_approx(m.param) : m.param;
            const float cos_theta = fminf(-udn, 1.0f);
            const float sin_theta = sqrt_approx(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
            const bool can_refract = ratio * sin_theta <= 1.0f;
            float r0 = __fdividef(1.0f - ratio, 1.0f + ratio);
            r0 *= r0;
            const float om = 1.0f - cos_theta, om2 = om * om;
            const float refl = fmaf(1.0f - r0, om2 * om2 * om, r0);  // Schlick material.zig:87-91
            pd.s = can_refract ? refl : 1.0f;                         // refract iff can_refract && refl < xi (material.zig:81)
            const float px = ratio * fmaf(cos_theta, s.nx, ux), py = ratio * fmaf(cos_theta, s.ny, uy),
                        pz = ratio * fmaf(cos_theta, s.nz, uz);       // refract material.zig:116-121
            const float par = -sqrt_approx(fabsf(1.0f - fmaf(px, px, fmaf(py, py, pz * pz))));
            pd.bx = fmaf(par, s.nx, px); pd.by = fmaf(par, s.ny, py); pd.bz = fmaf(par, s.nz, pz);
        }
    }
    r.ox = s.px; r.oy = s.py; r.oz = s.pz;  // time is kept (material.zig:49)
    return true;
}

__device__ __forceinline__ void scatter_finish(Ray &r, const Pending &pd, const Draw &dw) {
    if (pd.kind == 0u) {  // diffuse: normal + randomUnitVector, the normal alone if that sum is near zero (material.zig:45-48)
        const float ndx = r.dx + dw.x, ndy = r.dy + dw.y, ndz = r.dz + dw.z;
        const bool tiny = fabsf(ndx) < 1e-8f && fabsf(ndy) < 1e-8f && fabsf(ndz) < 1e-8f;
        r.dx = tiny ? r.dx : ndx; r.dy = tiny ? r.dy : ndy; r.dz = tiny ? r.dz : ndz;
    } else if (pd.kind == 1u) {  // metal: reflected + fuzz * randomPointInUnitSphere (material.zig:62)
        if (pd.s > 0.0f) {
            const float k = pd.s * cbrtf(dw.uc);
            r.dx = fmaf(k, dw.x, r.dx); r.dy = fmaf(k, dw.y, r.dy); r.dz = fmaf(k, dw.z, r.dz);
        }
    } else if (pd.s < dw.ua) {  // dielectric: refract when Schlick's reflectance is below the uniform (material.zig:81)
        r.dx = pd.bx; r.dy = pd.by; r.dz = pd.bz;
    }
}

// prepare + draw + finish in one call (deterministic, BVH and wavefront kernels, unit probe).  Returns true when the
// path continues with `r` replaced by the scattered ray.
template <bool STATS, uint32_t FEAT = FF_ALL>
__device__ __forceinline__ bool shade(const DevScene &sc, const DevRender &rp, Ray &r, const DevPrim &prim,
                                      uint32_t prim_id, float t, uint32_t pixel, uint32_t sample, uint32_t bounce,
                                      float3 &beta, float3 &L, Counters<STATS> &cn) {
    Pending pd;
    if (!shade_prepare<STATS, FEAT>(sc, r, prim, prim_id, t, beta, L, pd, cn)) return false;
    const uint4 rn = philox4x32_10(make_uint4(pixel, sample, bounce, 0u), rp.philox_keys);
    scatter_finish(r, pd, make_draw(rn, false));
    return true;
}

}  // namespace rtw
