// rtw_probe.cu — reference-ORDER closest-hit probe (parity instrument, K3).  Compiled with
// -fmad=false: no FMA contraction, IEEE div/sqrt, so that for Real=float and Real=double every
// +,-,*,/,sqrt rounds exactly as the same statement does on a CPU without contraction.
//
// This is the reference's `world.hit(r, 0.001, inf, &rec)` (src/main.zig:109) written statement for
// statement in the reference's evaluation order (src/rtw/hittable.zig:95-131, 165-201, 219-221,
// 278-303, 331-356, 384-409, 478-489, 558-596; src/rtw/vec.zig), templated on Real.  It exists so
// that primary-hit ids can be compared BIT-EXACT with the CPU restatement; the renderer itself uses
// the robust fp32 forms in rtw_trace.cuh.
#include <cuda_runtime.h>

#include "../../include/rtw_cuda.h"
#include "rtw_kernels.h"

namespace rtw {

template <class R>
struct V3 {
    R x, y, z;
};
template <class R> __device__ __forceinline__ V3<R> add(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> __device__ __forceinline__ V3<R> sub(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> __device__ __forceinline__ V3<R> mul(V3<R> a, R t) { return {a.x * t, a.y * t, a.z * t}; }
template <class R> __device__ __forceinline__ V3<R> divs(V3<R> a, R t) { return {a.x / t, a.y / t, a.z / t}; }
template <class R> __device__ __forceinline__ R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> __device__ __forceinline__ R norm2(V3<R> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

template <class R> __device__ __forceinline__ R sqrt_rn_(R x);
template <> __device__ __forceinline__ float sqrt_rn_<float>(float x) { return __fsqrt_rn(x); }
template <> __device__ __forceinline__ double sqrt_rn_<double>(double x) { return __dsqrt_rn(x); }
template <class R> __device__ __forceinline__ R inf_();
template <> __device__ __forceinline__ float inf_<float>() { return __int_as_float(0x7f800000); }
template <> __device__ __forceinline__ double inf_<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

template <class R>
struct RRay {
    V3<R> o, d;
    R time;
};
template <class R>
struct RRec {
    V3<R> p, n;
    R t, u, v;
    bool front;
};

template <class R>
__device__ bool ref_sphere(V3<R> center, R radius, bool write_uv, const RRay<R> &r, R t_min, R t_max, RRec<R> &rec) {
    const V3<R> oc = sub(r.o, center);  // hittable.zig:96-101
    const R a = norm2(r.d);
    const R half_b = dot(oc, r.d);
    const R c = norm2(oc) - radius * radius;
    const R disc = half_b * half_b - a * c;
    if (disc < R(0)) return false;
    const R sqrtd = sqrt_rn_<R>(disc);
    R root = (-half_b - sqrtd) / a;  // hittable.zig:108-116
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    rec.t = root;
    rec.p = add(r.o, mul(r.d, root));
    const V3<R> outward = divs(sub(rec.p, center), radius);
    rec.front = dot(outward, r.d) < R(0);
    rec.n = rec.front ? outward : mul(outward, R(-1));
    if (write_uv) {  // getSphereUv hittable.zig:145-150 (libm-class functions: tolerance, not bit parity)
        const R pi = R(3.14159265358979323846);
        rec.u = (atan2(-outward.z, outward.x) + pi) / (R(2) * pi);
        rec.v = acos(-outward.y) / pi;
    } else {
        rec.u = R(0); rec.v = R(0);
    }
    return true;
}

template <class R>
__device__ bool ref_rect(uint32_t kind, const double *v, const RRay<R> &r, R t_min, R t_max, RRec<R> &rec) {
    const R o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    int aa, ab, ak;
    if (kind == RTW_PRIM_XY_RECT) { aa = 0; ab = 1; ak = 2; }
    else if (kind == RTW_PRIM_XZ_RECT) { aa = 0; ab = 2; ak = 1; }
    else { aa = 1; ab = 2; ak = 0; }
    const R a0 = R(v[0]), a1 = R(v[1]), b0 = R(v[2]), b1 = R(v[3]), k = R(v[4]);
    const R t = (k - o[ak]) / d[ak];  // hittable.zig:279-282
    if (t < t_min || t > t_max) return false;
    const R pa = o[aa] + t * d[aa];
    const R pb = o[ab] + t * d[ab];
    if (pa < a0 || pa > a1 || pb < b0 || pb > b1) return false;
    rec.u = (pa - a0) / (a1 - a0);
    rec.v = (pb - b0) / (b1 - b0);
    rec.t = t;
    rec.p = add(r.o, mul(r.d, t));
    V3<R> outward{R(0), R(0), R(0)};
    if (ak == 0) outward.x = R(1); else if (ak == 1) outward.y = R(1); else outward.z = R(1);
    rec.front = dot(outward, r.d) < R(0);
    rec.n = rec.front ? outward : mul(outward, R(-1));
    return true;
}

template <class R>
__device__ bool ref_prim(const RawScene &sc, uint32_t id, const RRay<R> &rw, R t_min, R t_max, RRec<R> &rec) {
    const RawPrim &p = sc.prims[id];
    RRay<R> r = rw;
    for (uint32_t k = 0; k < p.chain_len; ++k) {  // outermost first
        const RawXform &x = sc.chains[p.chain_begin + k];
        if (x.kind == RTW_XFORM_TRANSLATE) {  // Translate.hit hittable.zig:479-483
            r.o = sub(r.o, V3<R>{R(x.v[0]), R(x.v[1]), R(x.v[2])});
        } else {  // RotateY.hit hittable.zig:560-573
            const R s = R(x.v[0]), c = R(x.v[1]);
            const V3<R> o = r.o, d = r.d;
            r.o.x = c * o.x - s * o.z;
            r.o.z = s * o.x + c * o.z;
            r.d.x = c * d.x - s * d.z;
            r.d.z = s * d.x + c * d.z;
        }
    }
    bool h;
    if (p.kind == RTW_PRIM_SPHERE) {
        h = ref_sphere<R>({R(p.v[0]), R(p.v[1]), R(p.v[2])}, R(p.v[3]), true, r, t_min, t_max, rec);
    } else if (p.kind == RTW_PRIM_MOVING_SPHERE) {  // centre(t) hittable.zig:219-221
        const V3<R> c0{R(p.v[0]), R(p.v[1]), R(p.v[2])}, c1{R(p.v[3]), R(p.v[4]), R(p.v[5])};
        const V3<R> c = add(c0, mul(sub(c1, c0), (r.time - R(p.v[6])) / (R(p.v[7]) - R(p.v[6]))));
        h = ref_sphere<R>(c, R(p.v[8]), false, r, t_min, t_max, rec);
    } else {
        h = ref_rect<R>(p.kind, p.v, r, t_min, t_max, rec);
    }
    if (!h) return false;
    for (int k = (int)p.chain_len - 1; k >= 0; --k) {  // innermost first on the way out
        const RawXform &x = sc.chains[p.chain_begin + k];
        if (x.kind == RTW_XFORM_TRANSLATE) {  // hittable.zig:487
            rec.p = add(rec.p, V3<R>{R(x.v[0]), R(x.v[1]), R(x.v[2])});
        } else {  // hittable.zig:583-593
            const R s = R(x.v[0]), c = R(x.v[1]);
            const V3<R> q = rec.p, n = rec.n;
            rec.p.x = c * q.x + s * q.z;
            rec.p.z = -s * q.x + c * q.z;
            rec.n.x = c * n.x + s * n.z;
            rec.n.z = -s * n.x + c * n.z;
        }
    }
    return true;
}

template <class R>
__device__ bool ref_slab(const BvhNode &n, const RRay<R> &r, R t_min, R t_max) {
    // aabb.zig:8-45 (division form), widened by 8 ulps so a box is never culled by rounding
    const R eps = (sizeof(R) == 4 ? R(1.1920929e-7) : R(2.220446049250313e-16)) * R(8);
    const R mn[3] = {R(n.mnx), R(n.mny), R(n.mnz)}, mx[3] = {R(n.mxx), R(n.mxy), R(n.mxz)};
    const R o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    R lo = t_min, hi = t_max;
    for (int a = 0; a < 3; ++a) {
        const R s0 = (mn[a] - o[a]) / d[a], s1 = (mx[a] - o[a]) / d[a];
        if (s0 != s0 || s1 != s1) continue;
        R t0 = s0 < s1 ? s0 : s1, t1 = s0 < s1 ? s1 : s0;
        t0 -= fabs(t0) * eps; t1 += fabs(t1) * eps;
        lo = t0 > lo ? t0 : lo; hi = t1 < hi ? t1 : hi;
        if (hi < lo) return false;
    }
    return true;
}

template <class R, int VARIANT>
__device__ bool ref_closest(const RawScene &sc, const RRay<R> &r, RRec<R> &best, uint32_t &best_id) {
    const R t_min = R(0.001);
    R closest = inf_<R>();
    bool any = false;
    if (VARIANT == VAR_FLAT) {  // HittableList.hit hittable.zig:231-244
        for (uint32_t i = 0; i < sc.n_prims; ++i) {
            RRec<R> tmp;
            if (ref_prim<R>(sc, i, r, t_min, closest, tmp)) { any = true; closest = tmp.t; best = tmp; best_id = i; }
        }
        return any;
    }
    auto leaf = [&](uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            const uint32_t id = sc.bvh_prim_id[first + k];
            RRec<R> tmp;
            if (ref_prim<R>(sc, id, r, t_min, closest, tmp) && (!any || tmp.t < closest || id > best_id)) {
                any = true; closest = tmp.t; best = tmp; best_id = id;
            }
        }
    };
    if (sc.root_is_leaf) { leaf(sc.nodes[0].a & kNodeRefIndexMask, sc.nodes[0].b); return any; }
    uint32_t stack[64];
    int sp = 0;
    stack[sp++] = sc.nodes[0].a;
    while (sp) {
        const uint32_t pair = stack[--sp];
        for (int side = 0; side < 2; ++side) {
            const BvhNode n = sc.nodes[pair + side];
            if (!ref_slab<R>(n, r, t_min, closest)) continue;
            if (n.b) leaf(n.a & kNodeRefIndexMask, n.b);
            else if (sp < 64) stack[sp++] = n.a;
        }
    }
    return any;
}

template <class R, int VARIANT>
__global__ void __launch_bounds__(128) k_ref_probe(const RawScene sc, uint32_t n, const double *__restrict__ rays,
                                                   const RawCamera cam, uint32_t width, uint32_t height, bool use_cam,
                                                   uint32_t *prim_id, double *t_out, double *normal, double *uv) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    RRay<R> r;
    if (use_cam) {
        // parity-mode primary ray: main.zig:390-391 with xi = 0.5, Camera.getRay main.zig:93-99 with a
        // zero lens offset, time = time0 + 0.5 (time1 - time0) (rand.zig:18-20 with xi = 0.5)
        const uint32_t j = idx / width, i = idx - j * width;
        const R s = (R(i) + R(0.5)) / (R(width) - R(1));
        const R t = (R(j) + R(0.5)) / (R(height) - R(1));
        const V3<R> org{R(cam.origin[0]), R(cam.origin[1]), R(cam.origin[2])};
        const V3<R> hor{R(cam.horizontal[0]), R(cam.horizontal[1]), R(cam.horizontal[2])};
        const V3<R> ver{R(cam.vertical[0]), R(cam.vertical[1]), R(cam.vertical[2])};
        const V3<R> llc{R(cam.llc[0]), R(cam.llc[1]), R(cam.llc[2])};
        const V3<R> zero{R(0), R(0), R(0)};
        r.d = sub(sub(add(add(llc, mul(hor, s)), mul(ver, t)), org), zero);
        r.o = add(org, zero);
        r.time = R(cam.time0) + R(0.5) * (R(cam.time1) - R(cam.time0));
    } else {
        const double *q = rays + 7 * (size_t)idx;
        r.o = {R(q[0]), R(q[1]), R(q[2])};
        r.d = {R(q[3]), R(q[4]), R(q[5])};
        r.time = R(q[6]);
    }
    RRec<R> rec;
    uint32_t id = kMiss;
    const bool h = ref_closest<R, VARIANT>(sc, r, rec, id);
    prim_id[idx] = h ? id : kMiss;
    if (t_out) t_out[idx] = h ? (double)rec.t : 0.0;
    if (normal) {
        normal[3 * (size_t)idx + 0] = h ? (double)rec.n.x : 0.0;
        normal[3 * (size_t)idx + 1] = h ? (double)rec.n.y : 0.0;
        normal[3 * (size_t)idx + 2] = h ? (double)rec.n.z : 0.0;
    }
    if (uv) {
        uv[2 * (size_t)idx + 0] = h ? (double)rec.u : 0.0;
        uv[2 * (size_t)idx + 1] = h ? (double)rec.v : 0.0;
    }
}

cudaError_t launch_ref_probe(int precision, int variant, const RawScene &sc, uint32_t n, const double *rays,
                             const RawCamera *cam, uint32_t width, uint32_t height, uint32_t *prim_id, double *t,
                             double *normal, double *uv, cudaStream_t st) {
    const int grid = (int)((n + 127) / 128);
    RawCamera c{};
    const bool use_cam = cam != nullptr;
    if (use_cam) c = *cam;
#define RTW_LAUNCH(R, V) k_ref_probe<R, V><<<grid, 128, 0, st>>>(sc, n, rays, c, width, height, use_cam, prim_id, t, normal, uv)
    if (precision == 32) {
        if (variant == VAR_FLAT) RTW_LAUNCH(float, VAR_FLAT); else RTW_LAUNCH(float, VAR_BVH);
    } else {
        if (variant == VAR_FLAT) RTW_LAUNCH(double, VAR_FLAT); else RTW_LAUNCH(double, VAR_BVH);
    }
#undef RTW_LAUNCH
    return cudaGetLastError();
}

}  // namespace rtw
