// rtw_api.cpp — implementation of the C ABI of include/rtw_cuda.h.
//
// Host side of the seam cut around the reference's render loop (src/main.zig:382-402): validates and
// copies the flattened scene, lowers it to the fp32 device layout of rtw_device.cuh (this is where
// f64 -> fp32 happens), builds the BVH, owns the device buffers, launches the kernels.  No CPU
// fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtw_cuda.h"
#include "rtw_bvh.h"
#include "rtw_kernels.h"
#include "rtw_lbvh.h"

using namespace rtw;

namespace {

// Spheres at least this large get the reference-point form of |o-c|^2 - r^2 (rtw_trace.cuh).
constexpr double kBigSphereRadius = 64.0;
// Scenes up to this many primitives default to the warp-uniform flat scan (measured crossover).
constexpr uint32_t kFlatAutoMax = 256;  // measured crossover on sphere scenes (round 2 kernels, 1080p x 50 spp): flat 27 % ahead at 145 prims, 12 % at 198, level at 256, BVH 10 % ahead at 326
constexpr uint32_t kLbvhAutoMin = 1u << 16;  // scenes at least this large build their BVH on the device
constexpr uint32_t kFlatHardMax = 6000;  // the shared-memory image must stay under ~200 KB
constexpr uint32_t kMaxCheckerDepth = 8;  // = the guard of the device loop in texture_value (rtw_trace.cuh)
constexpr uint32_t kBatchSpp = 64;       // pooled kernel: a batch = one 8x4 tile x 64 samples = 2048 paths

thread_local std::string g_create_error;

float bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Device array that only ever grows: a re-upload of a scene of the same or a smaller size reuses the allocation
// (cudaMalloc / cudaFree of 100 MB-class buffers were measured to stall for hundreds of milliseconds now and then).
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;    // elements in use
    size_t cap = 0;  // elements allocated
    cudaError_t alloc(size_t count) {
        const size_t want = std::max<size_t>(1, count);  // keep a valid pointer so kernels never see null tables
        if (want > cap) {
            release();
            cudaError_t e = cudaMalloc(&p, want * sizeof(T));
            if (e != cudaSuccess) { p = nullptr; return e; }
            cap = want;
        }
        n = count;
        return cudaSuccess;
    }
    cudaError_t upload(const std::vector<T> &h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = cap = 0;
    }
};

// Splits [0, n) into fixed chunks (independent of the machine, so chunk-ordered reductions are reproducible) and runs
// fn(chunk, lo, hi) over them on a few threads when the scene is large; small scenes stay on the calling thread.
constexpr uint32_t kChunk = 1u << 15;
inline uint32_t n_chunks_of(uint32_t n) { return (n + kChunk - 1) / kChunk; }
template <class F>
void for_chunks(uint32_t n, unsigned max_threads, F &&fn) {
    const uint32_t nc = n_chunks_of(n);
    unsigned nt = 1;
    if (nc >= 4) {
        nt = max_threads ? max_threads : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        nt = std::min<unsigned>(nt, nc);
    }
    auto run = [&](unsigned t) {
        for (uint32_t c = t; c < nc; c += nt) fn(c, c * kChunk, std::min(n, (c + 1) * kChunk));
    };
    if (nt <= 1) { run(0); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(run, t);
    run(0);
    for (auto &t : pool) t.join();
}

// Tuning knobs.  Each has an environment variable of the same name, read ONCE in rtw_cuda_create (never on the render
// path); rtw_cuda_set_option changes one on a live context.
const char *const kOptionNames[] = {"RTW_SPP_CHUNK", "RTW_BATCH_SPP", "RTW_BVH_LEAF", "RTW_BVH_THRESH", "RTW_BVH_STEPS", "RTW_WF_SLOTS",
                                    "RTW_BVH_LEAF_MAX", "RTW_BVH_BUILDER", "RTW_BUILD_THREADS", "RTW_UPLOAD_TRACE",
                                    "RTW_FLAT_KERNEL", "RTW_BOX_PRIMS", "RTW_FLAT_SPECIALISE", "RTW_MID_SPHERES", "RTW_GROUP_ROUND", "RTW_BVH_KERNEL", "RTW_LBVH_POW"};
struct Options {
    std::map<std::string, std::string> v;
    const char *get(const char *name) const {
        auto it = v.find(name);
        return it == v.end() ? nullptr : it->second.c_str();
    }
    long num(const char *name, long dflt) const {
        const char *e = get(name);
        return e && *e ? atol(e) : dflt;
    }
    static bool known(const char *name) {
        for (const char *k : kOptionNames) if (std::strcmp(k, name) == 0) return true;
        return false;
    }
    void read_env() {
        for (const char *k : kOptionNames) if (const char *e = getenv(k)) v[k] = e;
    }
};

// option RTW_UPLOAD_TRACE=1: per-stage wall times of rtw_cuda_upload_scene on stderr
struct Laps {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool on = false;
    explicit Laps(bool enabled) : on(enabled) {}
    void lap(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[rtw upload] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

}  // namespace

struct rtw_ctx {
    int device = 0;
    int n_sms = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {};
    cudaEvent_t ev_launch = nullptr;   // end of the last accumulate launch: the next one (any stream) is ordered after it
    cudaEvent_t ev_trace = nullptr;    // render_multi: "this device's samples are in its buffer"
    bool launch_recorded = false;
    bool have_scene = false;
    Options opt;
    std::vector<int> peers_enabled;    // devices whose memory this context's device can already map
    unsigned build_threads() const { return (unsigned)std::max(0l, opt.num("RTW_BUILD_THREADS", 0)); }

    DevBuf<DevPrim> prims_flat, prims_bvh;
    DevBuf<float4> flat_blob;
    DevBuf<uint32_t> bvh_prim_id, prim_material;
    DevBuf<BvhNode> nodes;
    DevBuf<DevXform> xforms;
    DevBuf<DevBigSphere> bigs;
    DevBuf<DevMaterial> materials;
    DevBuf<DevTexture> textures;
    DevBuf<DevImage> images;
    DevBuf<DevPerlin> perlins;
    DevBuf<RawPrim> raw_prims;
    DevBuf<RawXform> raw_chains;
    DevBuf<float> lbvh_boxes;      // device BVH build: fp32 leaf boxes, primitive ids, work arena (kept between uploads)
    DevBuf<uint32_t> lbvh_ids;
    DevBuf<unsigned char> lbvh_arena;
    std::vector<Box3d> scratch_boxes;  // host scratch reused between uploads (first touch of 50-100 MB is not free)
    std::vector<cudaArray_t> image_arrays;
    std::vector<cudaTextureObject_t> image_tex;
    DevScene scene{};
    RawScene raw{};
    // host copy kept for the reference-order probe (object-space f64 fields + nested chains), uploaded on first use
    std::vector<rtw_prim> host_prims;  // instanced spheres already moved to world space ...
    std::vector<std::pair<uint32_t, std::array<double, 10>>> host_inst_orig;  // ... their object-space fields
    std::vector<rtw_xform> host_xforms;
    uint32_t n_prims = 0;
    uint32_t flat_feat = FF_ALL;  // FF_* features of the uploaded scene (selects the specialised flat megakernel)
    bool root_is_leaf = false;

    // wavefront variant: path-state slots
    DevBuf<float4> wf_ro, wf_rd, wf_beta, wf_rad;
    DevBuf<uint32_t> wf_hit, wf_free0, wf_free1, wf_cursor;
    DevBuf<WfCounters> wf_counters;
    WfCounters *wf_host_counters = nullptr;
    uint32_t wf_slots = 0;

    DevBuf<float4> accum;       // internal accumulation buffer of rtw_cuda_render
    DevBuf<uint8_t> rgb8;
    DevBuf<unsigned int> tile_counter;
    DevBuf<unsigned long long> stat_counters;
    rtw_stats stats{};
};

namespace {

int fail(rtw_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(ctx, 2, "%s failed: %s", #call, cudaGetErrorString(e_));   \
    } while (0)

void free_images(rtw_ctx *c) {
    for (auto t : c->image_tex) cudaDestroyTextureObject(t);
    for (auto a : c->image_arrays) cudaFreeArray(a);
    c->image_tex.clear();
    c->image_arrays.clear();
}

// ---- leaf boxes: the reference's boudingBox rules -------------------------------------------------
// sphere c +- r (hittable.zig:133-143); moving sphere = union of the boxes at time0 and time1
// (hittable.zig:203-217); rects padded +-1e-4 on the thin axis (hittable.zig:305-316, 358-369,
// 411-422); RotateY = box of the 8 rotated corners (hittable.zig:516-556); Translate shifts
// (hittable.zig:491-498).
// Production rect tests evaluate t = (k - o) * rcp(d) and p = o + t d in fp32 with a 1-ulp reciprocal, so a ray aimed
// EXACTLY at the shared edge of two abutting rects (Cornell's wall/floor seams seen from the symmetric camera: pixel
// centres on the image diagonals) can land a few 1e-5 outside both and leak through (found by the seam check of
// tests/test_gpu_parity.py: 63 of 360 000 parity-mode pixels).  The fp32 records therefore carry in-plane bounds widened
// by 2^-19 of the rect's own coordinate magnitude — below the rounding noise of the test itself, invisible in the image,
// and watertight.  The f64 tables of the reference-order probe keep the exact bounds.
double rect_slack(const rtw_prim &p) {
    double mag = 1.0;
    for (int a = 0; a < 5; ++a) mag = std::max(mag, std::fabs(p.v[a]));
    return mag * 1.9073486328125e-6;
}

Box3d leaf_box(const rtw_scene_desc *s, const rtw_prim &p) {
    Box3d b{};
    auto set = [&](double x0, double y0, double z0, double x1, double y1, double z1) {
        b.mn[0] = x0; b.mn[1] = y0; b.mn[2] = z0; b.mx[0] = x1; b.mx[1] = y1; b.mx[2] = z1;
    };
    switch (p.kind) {
        // A negative radius is legal in the reference (the hollow-glass idiom: Sphere.hit uses r*r and divides by r,
        // hittable.zig:99,120, so the normal flips).  Its boudingBox (hittable.zig:136-141) would come out inverted;
        // the box a BVH needs is the one of |r|.
        case RTW_PRIM_SPHERE: {
            const double r = std::fabs(p.v[3]);
            set(p.v[0] - r, p.v[1] - r, p.v[2] - r, p.v[0] + r, p.v[1] + r, p.v[2] + r);
            break;
        }
        case RTW_PRIM_MOVING_SPHERE: {
            const double r = std::fabs(p.v[8]);
            for (int a = 0; a < 3; ++a) {
                const double c0 = p.v[a], c1 = p.v[3 + a];
                const double o0 = c0 + (c1 - c0) * ((s->time0 - p.v[6]) / (p.v[7] - p.v[6]));
                const double o1 = c0 + (c1 - c0) * ((s->time1 - p.v[6]) / (p.v[7] - p.v[6]));
                b.mn[a] = std::fmin(o0 - r, o1 - r);
                b.mx[a] = std::fmax(o0 + r, o1 + r);
            }
            break;
        }
        // rects: thin axis padded by 1e-4 as the reference does; in-plane extent = the widened bounds of the fp32 record
        case RTW_PRIM_XY_RECT: { const double e = rect_slack(p); set(p.v[0] - e, p.v[2] - e, p.v[4] - 0.0001, p.v[1] + e, p.v[3] + e, p.v[4] + 0.0001); break; }
        case RTW_PRIM_XZ_RECT: { const double e = rect_slack(p); set(p.v[0] - e, p.v[4] - 0.0001, p.v[2] - e, p.v[1] + e, p.v[4] + 0.0001, p.v[3] + e); break; }
        default: { const double e = rect_slack(p); set(p.v[4] - 0.0001, p.v[0] - e, p.v[2] - e, p.v[4] + 0.0001, p.v[1] + e, p.v[3] + e); break; }
    }
    for (int x = p.xform; x >= 0; x = s->xforms[x].outer) {  // innermost -> outermost
        const rtw_xform &xf = s->xforms[x];
        if (xf.kind == RTW_XFORM_TRANSLATE) {
            for (int a = 0; a < 3; ++a) { b.mn[a] += xf.v[a]; b.mx[a] += xf.v[a]; }
        } else {
            const double sn = xf.v[0], cs = xf.v[1];
            Box3d o{};
            for (int a = 0; a < 3; ++a) { o.mn[a] = INFINITY; o.mx[a] = -INFINITY; }
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int k = 0; k < 2; ++k) {
                        const double x0 = i ? b.mx[0] : b.mn[0], y0 = j ? b.mx[1] : b.mn[1], z0 = k ? b.mx[2] : b.mn[2];
                        const double q[3] = {cs * x0 + sn * z0, y0, -sn * x0 + cs * z0};
                        for (int a = 0; a < 3; ++a) { o.mn[a] = std::fmin(o.mn[a], q[a]); o.mx[a] = std::fmax(o.mx[a], q[a]); }
                    }
            b = o;
        }
    }
    return b;
}

// Compose an instance chain (innermost index `x`) into object = A*world + t, A = Ry.
struct XformD {
    double c = 1.0, s = 0.0, t[3] = {0, 0, 0};
    // world = A^T (object - t)
    void to_world(const double o[3], double w[3]) const {
        const double x = o[0] - t[0], y = o[1] - t[1], z = o[2] - t[2];
        w[0] = c * x + s * z; w[1] = y; w[2] = -s * x + c * z;
    }
};
XformD compose_chain_d(const rtw_scene_desc *s, int x) {
    std::vector<int> chain;  // outermost first
    for (int k = x; k >= 0; k = s->xforms[k].outer) chain.push_back(k);
    std::reverse(chain.begin(), chain.end());
    XformD r;
    for (int k : chain) {
        const rtw_xform &xf = s->xforms[k];
        if (xf.kind == RTW_XFORM_TRANSLATE) {  // p <- p - offset
            for (int a = 0; a < 3; ++a) r.t[a] -= xf.v[a];
        } else {  // p <- R p with R: x' = c x - s z, z' = s x + c z   (hittable.zig:563-564)
            const double s2 = xf.v[0], c2 = xf.v[1];
            const double nc = c2 * r.c - s2 * r.s, ns = s2 * r.c + c2 * r.s;
            const double tx = c2 * r.t[0] - s2 * r.t[2], tz = s2 * r.t[0] + c2 * r.t[2];
            r.c = nc; r.s = ns; r.t[0] = tx; r.t[2] = tz;
        }
    }
    return r;
}
DevXform compose_chain(const rtw_scene_desc *s, int x) {
    const XformD r = compose_chain_d(s, x);
    DevXform d{};
    d.c = (float)r.c; d.s = (float)r.s; d.tx = (float)r.t[0]; d.ty = (float)r.t[1]; d.tz = (float)r.t[2];
    return d;
}

int validate(rtw_ctx *ctx, const rtw_scene_desc *s) {
    if (!s) return fail(ctx, 1, "scene is null");
    if (s->n_prims && !s->prims) return fail(ctx, 1, "prims is null");
    if (s->n_xforms && !s->xforms) return fail(ctx, 1, "xforms is null");
    if (s->n_materials && !s->materials) return fail(ctx, 1, "materials is null");
    if (s->n_textures && !s->textures) return fail(ctx, 1, "textures is null");
    if (s->n_images && !s->images) return fail(ctx, 1, "images is null");
    if (s->n_perlins && !s->perlins) return fail(ctx, 1, "perlins is null");
    if (!(s->time1 >= s->time0)) return fail(ctx, 1, "time1 < time0");
    for (uint32_t i = 0; i < s->n_xforms; ++i) {
        const rtw_xform &x = s->xforms[i];
        if (x.kind > RTW_XFORM_ROTATE_Y) return fail(ctx, 1, "xform %u: bad kind %u", i, x.kind);
        if (x.outer < -1 || x.outer >= (int)s->n_xforms || x.outer == (int)i)
            return fail(ctx, 1, "xform %u: bad outer index %d", i, x.outer);
    }
    for (uint32_t i = 0; i < s->n_xforms; ++i) {  // chains must terminate
        uint32_t steps = 0;
        for (int k = (int)i; k >= 0; k = s->xforms[k].outer)
            if (++steps > s->n_xforms) return fail(ctx, 1, "xform %u: cyclic chain", i);
    }
    for (uint32_t i = 0; i < s->n_prims; ++i) {
        const rtw_prim &p = s->prims[i];
        if (p.kind > RTW_PRIM_YZ_RECT) return fail(ctx, 1, "prim %u: bad kind %u", i, p.kind);
        if (p.material >= s->n_materials) return fail(ctx, 1, "prim %u: material %u out of range", i, p.material);
        if (p.xform < -1 || p.xform >= (int)s->n_xforms) return fail(ctx, 1, "prim %u: xform %d out of range", i, p.xform);
        if (p.kind == RTW_PRIM_MOVING_SPHERE && p.v[7] == p.v[6]) return fail(ctx, 1, "prim %u: time1 == time0", i);
    }
    for (uint32_t i = 0; i < s->n_materials; ++i) {
        const rtw_material &m = s->materials[i];
        if (m.kind > RTW_MAT_DIFFUSE_LIGHT) return fail(ctx, 1, "material %u: bad kind %u", i, m.kind);
        if ((m.kind == RTW_MAT_DIFFUSE || m.kind == RTW_MAT_DIFFUSE_LIGHT) && (m.texture < 0 || m.texture >= (int)s->n_textures))
            return fail(ctx, 1, "material %u: texture %d out of range", i, m.texture);
    }
    for (uint32_t i = 0; i < s->n_textures; ++i) {
        const rtw_texture &t = s->textures[i];
        if (t.kind > RTW_TEX_IMAGE) return fail(ctx, 1, "texture %u: bad kind %u", i, t.kind);
        if (t.kind == RTW_TEX_CHECKER && (t.a < 0 || t.b < 0 || t.a >= (int)s->n_textures || t.b >= (int)s->n_textures))
            return fail(ctx, 1, "texture %u: checker child out of range", i);
        if (t.kind == RTW_TEX_NOISE && (t.a < 0 || t.a >= (int)s->n_perlins)) return fail(ctx, 1, "texture %u: perlin out of range", i);
        if (t.kind == RTW_TEX_IMAGE && (t.a < 0 || t.a >= (int)s->n_images)) return fail(ctx, 1, "texture %u: image out of range", i);
    }
    // checker children are textures themselves (texture.zig:59-60): the graph must be acyclic and no deeper than the
    // device loop follows (kMaxCheckerDepth levels of checker above a non-checker texture)
    for (uint32_t i = 0; i < s->n_textures; ++i) {
        if (s->textures[i].kind != RTW_TEX_CHECKER) continue;
        std::vector<std::pair<int, uint32_t>> st{{(int)i, 1u}};
        uint64_t visited = 0;
        while (!st.empty()) {
            const auto [t, depth] = st.back();
            st.pop_back();
            if (s->textures[t].kind != RTW_TEX_CHECKER) continue;
            if (depth > kMaxCheckerDepth) return fail(ctx, 1, "texture %u: checker nesting deeper than %u (or cyclic)", i, kMaxCheckerDepth);
            if (++visited > 4096) return fail(ctx, 1, "texture %u: checker graph too large", i);
            st.push_back({s->textures[t].a, depth + 1});
            st.push_back({s->textures[t].b, depth + 1});
        }
    }
    for (uint32_t i = 0; i < s->n_images; ++i)
        if (!s->images[i].rgba8 || !s->images[i].width || !s->images[i].height) return fail(ctx, 1, "image %u: empty", i);
    for (uint32_t i = 0; i < s->n_perlins; ++i)
        if (!s->perlins[i].ranvec || !s->perlins[i].perm_x || !s->perlins[i].perm_y || !s->perlins[i].perm_z) return fail(ctx, 1, "perlin %u: null table", i);
    return 0;
}

DevCamera lower_camera(const rtw_camera *c) {
    DevCamera d{};
    d.ox = (float)c->origin[0]; d.oy = (float)c->origin[1]; d.oz = (float)c->origin[2];
    d.hx = (float)c->horizontal[0]; d.hy = (float)c->horizontal[1]; d.hz = (float)c->horizontal[2];
    d.vx = (float)c->vertical[0]; d.vy = (float)c->vertical[1]; d.vz = (float)c->vertical[2];
    d.lx = (float)c->lower_left_corner[0]; d.ly = (float)c->lower_left_corner[1]; d.lz = (float)c->lower_left_corner[2];
    d.ux = (float)c->u[0]; d.uy = (float)c->u[1]; d.uz = (float)c->u[2];
    d.wx = (float)c->v[0]; d.wy = (float)c->v[1]; d.wz = (float)c->v[2];
    d.lens_radius = (float)c->lens_radius; d.time0 = (float)c->time0; d.time1 = (float)c->time1;
    return d;
}

int pick_variant(rtw_ctx *ctx, uint32_t requested, int *out, bool uses_smem = true) {
    int v;
    switch (requested) {
        case RTW_VARIANT_AUTO: v = ctx->n_prims <= kFlatAutoMax ? VAR_FLAT : VAR_BVH; break;
        case RTW_VARIANT_MEGA_FLAT: v = VAR_FLAT; break;
        case RTW_VARIANT_MEGA_BVH: v = VAR_BVH; break;
        case RTW_VARIANT_WAVEFRONT: v = ctx->n_prims <= kFlatAutoMax ? VAR_FLAT : VAR_BVH; break;  // traversal of the wavefront
        default: return fail(ctx, 1, "unknown variant %u", requested);
    }
    if (v == VAR_FLAT && uses_smem && (ctx->n_prims > kFlatHardMax || (size_t)ctx->scene.flat.total_f4 * 16 > 200 * 1024))
        return fail(ctx, 1, "flat variant: the scene (%u primitives, %u B image) does not fit in shared memory", ctx->n_prims,
                    ctx->scene.flat.total_f4 * 16u);
    *out = v;
    return 0;
}

}  // namespace

extern "C" {

uint32_t rtw_cuda_abi_version(void) { return RTW_ABI_VERSION; }

const char *rtw_cuda_last_error(const rtw_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rtw_cuda_create(int device, rtw_ctx **out) {
    rtw_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, 1, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, 3, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, 1, "device %d out of range (have %d)", device, count);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(nullptr, 3, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    ctx = new rtw_ctx();
    ctx->device = device;
    ctx->n_sms = prop.multiProcessorCount;
    // A BLOCKING stream on purpose: cudaMemcpy from pageable memory returns once the data is staged, possibly
    // before the DMA lands; work in a blocking stream is ordered after it, work in a non-blocking stream is not.
    if (cudaStreamCreate(&ctx->stream) != cudaSuccess) { delete ctx; return fail(nullptr, 2, "stream create failed"); }
    for (auto &ev : ctx->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) { delete ctx; return fail(nullptr, 2, "event create failed"); }
    if (cudaEventCreateWithFlags(&ctx->ev_launch, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_trace, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, 2, "event create failed");
    }
    ctx->opt.read_env();  // the only place the environment is read
    if (ctx->tile_counter.alloc(1) != cudaSuccess || ctx->stat_counters.alloc(ST_COUNT) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, 2, "device allocation failed");
    }
    cudaMemset(ctx->stat_counters.p, 0, ST_COUNT * sizeof(unsigned long long));
    *out = ctx;
    return 0;
}

void rtw_cuda_destroy(rtw_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    free_images(ctx);
    ctx->flat_blob.release();
    ctx->wf_ro.release(); ctx->wf_rd.release(); ctx->wf_beta.release(); ctx->wf_rad.release(); ctx->wf_hit.release();
    ctx->wf_free0.release(); ctx->wf_free1.release(); ctx->wf_cursor.release(); ctx->wf_counters.release();
    if (ctx->wf_host_counters) cudaFreeHost(ctx->wf_host_counters);
    ctx->prims_flat.release(); ctx->prims_bvh.release(); ctx->bvh_prim_id.release(); ctx->prim_material.release();
    ctx->nodes.release(); ctx->xforms.release(); ctx->bigs.release(); ctx->materials.release(); ctx->textures.release();
    ctx->images.release(); ctx->perlins.release(); ctx->raw_prims.release(); ctx->raw_chains.release();
    ctx->lbvh_boxes.release(); ctx->lbvh_ids.release(); ctx->lbvh_arena.release();
    ctx->accum.release(); ctx->rgb8.release(); ctx->tile_counter.release(); ctx->stat_counters.release();
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_launch) cudaEventDestroy(ctx->ev_launch);
    if (ctx->ev_trace) cudaEventDestroy(ctx->ev_trace);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rtw_cuda_set_option(rtw_ctx *ctx, const char *name, const char *value) {
    if (!ctx || !name) return fail(ctx, 1, "null argument");
    if (!Options::known(name)) return fail(ctx, 1, "unknown option %s", name);
    if (value) ctx->opt.v[name] = value; else ctx->opt.v.erase(name);
    return 0;
}

// Map every other context's device into `ctx`'s device (NVLink peer access), once.
static int enable_peers(rtw_ctx *ctx, rtw_ctx *const *ctxs, uint32_t n_ctx) {
    CK(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < n_ctx; ++i) {
        const int dev = ctxs[i]->device;
        if (dev == ctx->device || std::find(ctx->peers_enabled.begin(), ctx->peers_enabled.end(), dev) != ctx->peers_enabled.end()) continue;
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, ctx->device, dev));
        if (!can) return fail(ctx, 3, "device %d cannot map device %d's memory (no peer access)", ctx->device, dev);
        const cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, 2, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
        ctx->peers_enabled.push_back(dev);
    }
    return 0;
}

int rtw_cuda_create_multi(uint32_t n_gpus, rtw_ctx **out) {
    if (!out) return fail(nullptr, 1, "out is null");
    if (n_gpus == 0 || n_gpus > kMaxResolveBufs) return fail(nullptr, 1, "n_gpus must be 1..%u", kMaxResolveBufs);
    for (uint32_t i = 0; i < n_gpus; ++i) out[i] = nullptr;
    for (uint32_t i = 0; i < n_gpus; ++i) {
        if (int rc = rtw_cuda_create((int)i, &out[i])) {
            for (uint32_t k = 0; k < i; ++k) { rtw_cuda_destroy(out[k]); out[k] = nullptr; }
            return rc;
        }
    }
    // all-pairs peer mappings now, not inside the first timed frame: the slab-parallel resolve of rtw_cuda_render_multi
    // has every GPU read every other GPU's buffer
    for (uint32_t i = 0; i < n_gpus && n_gpus > 1; ++i) {
        if (int rc = enable_peers(out[i], out, n_gpus)) {
            g_create_error = out[i]->err;
            for (uint32_t k = 0; k < n_gpus; ++k) { rtw_cuda_destroy(out[k]); out[k] = nullptr; }
            return rc;
        }
    }
    return 0;
}

// Device-side BVH build (rtw_lbvh.cu) into ctx->nodes / bvh_prim_id / prims_bvh; ctx->prims_flat must be uploaded.
// Primitives whose box dwarfs the spread of the rest would drag every Morton ancestor up to scene size, so they get
// their own small SAH subtree (host) and the two subtrees meet under a new root:
//   nodes: [0] root  [1] pad  [2] big subtree root  [3] Morton subtree root  [4..] big pairs, then Morton pairs
//   slots: [0, nb) big primitives in their leaf order, [nb, n) the rest in Morton order
// *built = false (and rc 0) when the result would not fit the traversal (stack depth) or the input is degenerate.
static int build_bvh_on_device(rtw_ctx *ctx, const std::vector<Box3d> &boxes, uint32_t leaf_max, uint32_t *n_nodes,
                               uint32_t *depth, bool *built) {
    *built = false;
    Laps laps(ctx->opt.get("RTW_UPLOAD_TRACE") != nullptr);
    const unsigned bt = ctx->build_threads();
    const uint32_t n = (uint32_t)boxes.size();
    std::vector<float> fb(6 * (size_t)n);
    struct Part {
        float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};  // centroids
        float bmn[3] = {INFINITY, INFINITY, INFINITY}, bmx[3] = {-INFINITY, -INFINITY, -INFINITY};  // boxes
        bool finite = true;
        std::vector<uint32_t> big, small;
        void merge(const Part &o) {
            for (int a = 0; a < 3; ++a) {
                cmn[a] = std::min(cmn[a], o.cmn[a]); cmx[a] = std::max(cmx[a], o.cmx[a]);
                bmn[a] = std::min(bmn[a], o.bmn[a]); bmx[a] = std::max(bmx[a], o.bmx[a]);
            }
            finite = finite && o.finite;
        }
    };
    std::vector<Part> parts(n_chunks_of(n));
    for_chunks(n, bt, [&](uint32_t chunk, uint32_t lo, uint32_t hi) {
        Part pt;
        for (uint32_t i = lo; i < hi; ++i) {
            float *b = &fb[6 * (size_t)i];
            box_to_f32(boxes[i], b);
            for (int a = 0; a < 3; ++a) {
                const float c = 0.5f * (b[a] + b[3 + a]);
                if (!(std::fabs(c) <= 3e38f)) pt.finite = false;
                pt.cmn[a] = std::min(pt.cmn[a], c); pt.cmx[a] = std::max(pt.cmx[a], c);
                pt.bmn[a] = std::min(pt.bmn[a], b[a]); pt.bmx[a] = std::max(pt.bmx[a], b[3 + a]);
            }
        }
        parts[chunk] = std::move(pt);
    });
    Part all;
    for (const Part &pt : parts) all.merge(pt);
    if (!all.finite) return 0;  // non-finite boxes: leave them to the host builder
    const float spread = std::max(all.cmx[0] - all.cmn[0], std::max(all.cmx[1] - all.cmn[1], all.cmx[2] - all.cmn[2]));
    // second pass: split off the big ones, Morton grid over the centroids of what is left
    for_chunks(n, bt, [&](uint32_t chunk, uint32_t lo, uint32_t hi) {
        Part &pt = parts[chunk];
        pt = Part{};
        for (uint32_t i = lo; i < hi; ++i) {
            const float *b = &fb[6 * (size_t)i];
            const float ext = std::max(b[3] - b[0], std::max(b[4] - b[1], b[5] - b[2]));
            if (ext > 0.125f * spread) { pt.big.push_back(i); continue; }
            pt.small.push_back(i);
            for (int a = 0; a < 3; ++a) {
                const float c = 0.5f * (b[a] + b[3 + a]);
                pt.cmn[a] = std::min(pt.cmn[a], c); pt.cmx[a] = std::max(pt.cmx[a], c);
            }
        }
    });
    constexpr uint32_t kMaxBig = 4096;
    std::vector<uint32_t> big_ids, small_ids;
    small_ids.reserve(n);
    Part rest;
    for (const Part &pt : parts) {
        rest.merge(pt);
        big_ids.insert(big_ids.end(), pt.big.begin(), pt.big.end());
        small_ids.insert(small_ids.end(), pt.small.begin(), pt.small.end());
    }
    if (big_ids.size() > kMaxBig) return 0;  // no clear size gap: a job for the SAH builder
    const uint32_t nb = (uint32_t)big_ids.size(), ns = (uint32_t)small_ids.size();
    if (ns < 2 || ns <= leaf_max) return 0;
    const float *all_mn = all.bmn, *all_mx = all.bmx, *smn = rest.cmn, *smx = rest.cmx;
    laps.lap("  lbvh host prep");
    const float sext[3] = {smx[0] - smn[0], smx[1] - smn[1], smx[2] - smn[2]};

    BvhResult big;
    std::vector<BvhNode> head;
    std::vector<uint32_t> big_order(nb);
    uint32_t root_slot = 0, pair_base = 2;
    if (nb) {
        std::vector<Box3d> bb(nb);
        for (uint32_t k = 0; k < nb; ++k) bb[k] = boxes[big_ids[k]];
        big = build_bvh(bb, leaf_max, bt);
        for (uint32_t k = 0; k < nb; ++k) big_order[k] = big_ids[big.order[k]];
        const uint32_t B = (uint32_t)big.nodes.size();  // [0] root, [1] pad, pairs from 2
        head.assign((size_t)B + 2, BvhNode{0, 0, 0, 0, 0, 0, 0, 0});
        head[0] = BvhNode{all_mn[0], all_mn[1], all_mn[2], 2u, all_mx[0], all_mx[1], all_mx[2], 0u};
        auto moved = [](BvhNode nd) { if (nd.b == 0) nd.a += 2; return nd; };  // interior: its pair moves up by two
        head[2] = moved(big.nodes[0]);
        for (uint32_t k = 2; k < B; ++k) head[k + 2] = moved(big.nodes[k]);
        root_slot = 3;
        pair_base = B + 2;
    } else {
        head.assign(2, BvhNode{0, 0, 0, 0, 0, 0, 0, 0});
    }
    CK(cudaSetDevice(ctx->device));
    DevBuf<float> &d_boxes = ctx->lbvh_boxes;
    DevBuf<uint32_t> &d_ids = ctx->lbvh_ids;
    CK(d_boxes.upload(fb));
    CK(d_ids.upload(small_ids));
    CK(ctx->lbvh_arena.alloc(lbvh_arena_bytes(ns)));
    CK(ctx->nodes.alloc((size_t)pair_base + 2 * (size_t)(ns - 1)));
    CK(ctx->bvh_prim_id.alloc(n));
    CK(ctx->prims_bvh.alloc(n));
    CK(cudaMemcpy(ctx->nodes.p, head.data(), head.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
    if (nb) CK(cudaMemcpy(ctx->bvh_prim_id.p, big_order.data(), (size_t)nb * 4, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    laps.lap("  lbvh alloc + H2D");
    LbvhInfo info;
    // Morton grid candidates: cells of axis a are (extent_a / widest extent)^-p times as long as the cubic grid's, p in
    // {0, 1, 2, 4}: a thin axis then contributes its splits later.  On a scene that is a curved SHEET of spheres (10^6 spheres on
    // the r = 1000 ground: extents 1000 x 293 x 1000) the cubic grid's height splits cut rings out of discs — 51.7 node tests
    // per ray against 47.0 (p = 2), 46.9 (p = 4) and 46.9 for the host SAH builder; for a scene that fills its bounds the cubic
    // grid is the right one.  build_lbvh fits every candidate and emits the one with the smallest surface-area cost.
    // Option RTW_LBVH_POW = p pins one exponent.
    float axis_w[4 * 3];
    int n_cand = 0;
    {
        const float widest = std::max(sext[0], std::max(sext[1], sext[2]));
        const float thinnest = std::min(sext[0], std::min(sext[1], sext[2]));
        const char *pw = ctx->opt.get("RTW_LBVH_POW");
        std::vector<float> exps = pw ? std::vector<float>{(float)atof(pw)} : std::vector<float>{0.f, 1.f, 2.f, 4.f};
        if (!pw && !(thinnest < 0.9f * widest)) exps = {0.f};  // nothing to choose between
        for (float p : exps) {
            for (int a = 0; a < 3; ++a)
                axis_w[3 * n_cand + a] = (p > 0.f && widest > 0.f) ? std::pow(std::max(sext[a] / widest, 1e-6f), p) : 1.f;
            ++n_cand;
        }
    }
    CK(build_lbvh(d_boxes.p, d_ids.p, ns, smn, sext, leaf_max, ctx->nodes.p, root_slot, pair_base, ctx->bvh_prim_id.p, nb,
                  ctx->lbvh_arena.p, ctx->lbvh_arena.cap, ctx->stream, &info, axis_w, n_cand));
    if (laps.on && n_cand > 1)
        fprintf(stderr, "[rtw upload]   lbvh grid candidates: cost %.6g %.6g %.6g %.6g -> p index %d\n", info.cost[0], info.cost[1], info.cost[2], info.cost[3], info.chosen);
    laps.lap("  lbvh device build");
    const uint32_t total_depth = nb ? 1 + std::max(big.depth, info.depth) : info.depth;
    if (total_depth > 64) return 0;
    CK(gather_prims(ctx->prims_flat.p, ctx->bvh_prim_id.p, ctx->prims_bvh.p, n, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    laps.lap("  lbvh gather");
    *n_nodes = info.n_nodes;
    *depth = total_depth;
    *built = true;
    return 0;
}


int rtw_cuda_upload_scene(rtw_ctx *ctx, const rtw_scene_desc *s) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    const auto t_begin = std::chrono::steady_clock::now();
    if (int rc = validate(ctx, s)) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->have_scene = false;
    const uint32_t n = s->n_prims;

    // Instanced spheres (Translate / RotateY around a sphere): a rigid transform keeps a sphere a sphere, so they
    // are lowered to world-space spheres; the chain is kept only for the texture coordinates (getSphereUv works on
    // the object-space normal, hittable.zig:127).  `wprims` = the prims with such centres moved to world space.
    Laps laps(ctx->opt.get("RTW_UPLOAD_TRACE") != nullptr);
    const unsigned bt = ctx->build_threads();
    std::vector<rtw_prim> wprims = std::move(ctx->host_prims);  // storage of the previous scene's copy, if any
    wprims.resize(n);
    std::vector<Box3d> &boxes = ctx->scratch_boxes;
    boxes.resize(n);
    // Spheres tested individually ahead of the groups ("big"): |r| >= kBigSphereRadius always (their c term needs the
    // reference-point form), and in scenes small enough for the flat scan also the few static spheres that dwarf the rest
    // (> 2.5 x the median radius, at most eight: the three r = 1 spheres of main.zig:202-218 among r = 0.2 ones) — inside a
    // group they inflate its bound (a vote passes a group roughly in proportion to R^2), alone they cost one
    // discriminant each and, being hit early, shorten the reach for every bound after them.
    double big_thr = kBigSphereRadius;
    if (n <= kFlatAutoMax && ctx->opt.num("RTW_MID_SPHERES", 1) != 0) {
        std::vector<double> radii;
        for (uint32_t i = 0; i < n; ++i) {
            const rtw_prim &p = s->prims[i];
            if (p.kind == RTW_PRIM_SPHERE) radii.push_back(std::fabs(p.v[3]));
            else if (p.kind == RTW_PRIM_MOVING_SPHERE) radii.push_back(std::fabs(p.v[8]));
        }
        if (radii.size() >= 8) {
            std::vector<double> sorted(radii);
            std::nth_element(sorted.begin(), sorted.begin() + sorted.size() / 2, sorted.end());
            const double cut = 2.5 * sorted[sorted.size() / 2];
            uint32_t n_mid = 0;
            double smallest = kBigSphereRadius;
            for (uint32_t i = 0; i < n; ++i) {
                const rtw_prim &p = s->prims[i];
                if (p.kind == RTW_PRIM_SPHERE && p.xform < 0 && std::fabs(p.v[3]) > cut) { ++n_mid; smallest = std::min(smallest, std::fabs(p.v[3])); }
            }
            if (n_mid > 0 && n_mid <= 8) big_thr = smallest;
        }
    }
    // reference point for big spheres: centroid of the centres of everything that is not big (summed per fixed
    // chunk, chunks in order: the same bits on every machine)
    std::vector<std::array<double, 4>> cen_part(n_chunks_of(n), std::array<double, 4>{0, 0, 0, 0});
    for_chunks(n, bt, [&](uint32_t chunk, uint32_t lo, uint32_t hi) {
        std::array<double, 4> acc{0, 0, 0, 0};
        for (uint32_t i = lo; i < hi; ++i) {
            rtw_prim &p = wprims[i];
            p = s->prims[i];
            if (p.xform >= 0 && (p.kind == RTW_PRIM_SPHERE || p.kind == RTW_PRIM_MOVING_SPHERE)) {
                const XformD x = compose_chain_d(s, p.xform);
                x.to_world(&s->prims[i].v[0], &p.v[0]);
                if (p.kind == RTW_PRIM_MOVING_SPHERE) x.to_world(&s->prims[i].v[3], &p.v[3]);
                rtw_prim q = p;
                q.xform = -1;  // already in world space: exact box instead of the box of a rotated box ...
                boxes[i] = leaf_box(s, q);
                // ... padded: the reference-order fp32 probe intersects in OBJECT space, so its hit points carry the
                // rounding of the rotated ray (~1e-7 |o|) and may fall a hair outside the exact world-space box
                double mag = 1.0;
                for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(boxes[i].mn[a]), std::fabs(boxes[i].mx[a])));
                for (int a = 0; a < 3; ++a) { boxes[i].mn[a] -= 4e-6 * mag; boxes[i].mx[a] += 4e-6 * mag; }
            } else {
                boxes[i] = leaf_box(s, s->prims[i]);
            }
            if (p.kind == RTW_PRIM_SPHERE && p.xform < 0 && std::fabs(p.v[3]) >= big_thr) continue;
            for (int a = 0; a < 3; ++a) acc[a] += 0.5 * (boxes[i].mn[a] + boxes[i].mx[a]);
            acc[3] += 1.0;
        }
        cen_part[chunk] = acc;
    });
    double cen[3] = {0, 0, 0}, ncen = 0;
    for (const auto &part : cen_part) { cen[0] += part[0]; cen[1] += part[1]; cen[2] += part[2]; ncen += part[3]; }
    if (ncen > 0) for (double &c : cen) c /= ncen;
    laps.lap("copy + leaf boxes");

    // ---- lower to fp32 records --------------------------------------------------------------------
    std::vector<DevPrim> flat(n);
    std::vector<uint32_t> prim_mat(n);
    std::vector<DevBigSphere> bigs;
    std::vector<DevXform> xforms;
    std::map<int, int> xform_slot;
    auto xform_slot_of = [&](int x) {
        auto it = xform_slot.find(x);
        if (it == xform_slot.end()) {
            xforms.push_back(compose_chain(s, x));
            it = xform_slot.emplace(x, (int)xforms.size() - 1).first;
        }
        return it->second;
    };
    // Plain primitives (no instance chain, not a big sphere) touch no shared table and are lowered in parallel; the
    // rest follow in index order, which fixes the numbering of `bigs` and of the transform slots.
    auto lower = [&](uint32_t i) -> bool {
        const rtw_prim &p = wprims[i];
        DevPrim d{};
        prim_mat[i] = p.material;
        uint32_t meta = 0;
        uint32_t sphere_xf = 0;  // xform slot + 1 of an instanced sphere (uv only)
        if (p.xform >= 0 && (p.kind == RTW_PRIM_SPHERE || p.kind == RTW_PRIM_MOVING_SPHERE)) {
            sphere_xf = (uint32_t)xform_slot_of(p.xform) + 1u;
            if (sphere_xf > 0xFFFu) return false;
        }
        if (p.kind == RTW_PRIM_SPHERE) {
            d.a = make_float4((float)p.v[0], (float)p.v[1], (float)p.v[2], (float)p.v[3]);
            d.b = make_float4(0.f, 0.f, 0.f, 0.f);
            meta = PK_SPHERE;
            if ((std::fabs(p.v[3]) >= kBigSphereRadius || (p.xform < 0 && std::fabs(p.v[3]) >= big_thr)) && bigs.size() < 0xFFEu) {
                // q = point of the sphere surface nearest the scene's centre of interest
                double dir[3] = {cen[0] - p.v[0], cen[1] - p.v[1], cen[2] - p.v[2]};
                double len = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
                if (!(len > 0)) { dir[0] = 0; dir[1] = 1; dir[2] = 0; len = 1; }
                // the device works from the fp32-rounded centre and q: keep the identities exact in f64
                const double cf[3] = {(double)(float)p.v[0], (double)(float)p.v[1], (double)(float)p.v[2]};
                const double rf = std::fabs((double)(float)p.v[3]);
                DevBigSphere g{};
                double q[3], m[3], mm = 0;
                for (int a = 0; a < 3; ++a) q[a] = (double)(float)(cf[a] + dir[a] / len * rf);
                for (int a = 0; a < 3; ++a) { m[a] = q[a] - cf[a]; }
                g.qx = (float)q[0]; g.qy = (float)q[1]; g.qz = (float)q[2];
                g.mx = (float)m[0]; g.my = (float)m[1]; g.mz = (float)m[2];
                // K uses the fp32-rounded m so that |a + m|^2 - r^2 is what the kernel expands
                for (int a = 0; a < 3; ++a) { const double mf = (double)(&g.mx)[a]; mm += mf * mf; }
                g.K = (float)(mm - rf * rf);
                bigs.push_back(g);
                meta |= (uint32_t)bigs.size() << 8;
            }
        } else if (p.kind == RTW_PRIM_MOVING_SPHERE) {
            // centre(time) = c0 + (c1-c0) (time-t0)/(t1-t0) = cb + vel*time   (hittable.zig:219-221)
            const double inv = 1.0 / (p.v[7] - p.v[6]);
            float vel[3], cb[3];
            for (int a = 0; a < 3; ++a) {
                const double v = (p.v[3 + a] - p.v[a]) * inv;
                vel[a] = (float)v;
                cb[a] = (float)(p.v[a] - v * p.v[6]);
            }
            d.a = make_float4(cb[0], cb[1], cb[2], (float)p.v[8]);
            d.b = make_float4(vel[0], vel[1], vel[2], 0.f);
            meta = PK_SPHERE;
        } else {
            const double e = rect_slack(p);  // watertight seams, see rect_slack
            d.a = make_float4((float)(p.v[0] - e), (float)(p.v[1] + e), (float)(p.v[2] - e), (float)(p.v[3] + e));
            const int slot = p.xform >= 0 ? xform_slot_of(p.xform) : -1;
            d.b = make_float4((float)p.v[4], bits_to_float((uint32_t)slot), (float)e, 0.f);
            meta = p.kind == RTW_PRIM_XY_RECT ? PK_XY : p.kind == RTW_PRIM_XZ_RECT ? PK_XZ : PK_YZ;
        }
        d.b.w = bits_to_float(meta | (sphere_xf << 20));
        flat[i] = d;

        return true;
    };
    std::vector<std::vector<uint32_t>> special(n_chunks_of(n));
    for_chunks(n, bt, [&](uint32_t chunk, uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            const rtw_prim &p = wprims[i];
            if (p.xform >= 0 || (p.kind == RTW_PRIM_SPHERE && std::fabs(p.v[3]) >= big_thr)) special[chunk].push_back(i);
            else lower(i);
        }
    });
    for (const auto &part : special)
        for (uint32_t i : part)
            if (!lower(i)) return fail(ctx, 1, "too many distinct instance chains on spheres (max 4095)");
    laps.lap("lower to fp32");

    // which primitive / texture kinds the scene has: selects the specialised megakernels (FF_* in rtw_device.cuh)
    {
        uint32_t feat = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const rtw_prim &p = s->prims[i];
            const bool sphere = p.kind == RTW_PRIM_SPHERE || p.kind == RTW_PRIM_MOVING_SPHERE;
            feat |= sphere ? FF_SPHERES : FF_RECTS;
            if (sphere && p.xform >= 0) feat |= FF_TEX;  // instanced sphere: uv in object space (finalise_hit)
        }
        for (uint32_t i = 0; i < s->n_textures; ++i)
            if (s->textures[i].kind != RTW_TEX_SOLID && s->textures[i].kind != RTW_TEX_CHECKER) feat |= FF_TEX;
        ctx->flat_feat = feat;
    }

    // ---- BVH ------------------------------------------------------------------------------------------
    // Small and medium scenes: binned SAH on the host (best trees, microseconds to milliseconds).  Large scenes:
    // Morton/Karras build on the device (rtw_lbvh.cu), with the few primitives that dwarf the rest (the ground
    // sphere) kept out of the Morton order and grafted next to the root as their own host-built subtree.
    // Leaves of up to four primitives where the surface-area heuristic prefers them — for scenes with rects and boxes (Cornell:
    // leaves of one primitive +2 % time); sphere-only scenes get one sphere per leaf (10^6 spheres -0.8 %, 485 spheres -1.6 %,
    // 9*10^4 spheres -1.9 %: a leaf's sphere test costs more than the node test a fold saves)
    uint32_t leaf_max = (ctx->flat_feat & FF_RECTS) ? 4 : 1;
    leaf_max = (uint32_t)std::max(1l, std::min(15l, ctx->opt.num("RTW_BVH_LEAF_MAX", leaf_max)));  // 4-bit count in node refs
    CK(ctx->prims_flat.upload(flat));
    const auto t_bvh = std::chrono::steady_clock::now();
    bool use_lbvh = n >= kLbvhAutoMin;
    if (const char *e = ctx->opt.get("RTW_BVH_BUILDER")) use_lbvh = std::strcmp(e, "lbvh") == 0 && n >= 64;
    uint32_t bvh_n_nodes = 0, bvh_depth = 0;
    bool bvh_root_is_leaf = false, bvh_on_device = false;
    if (use_lbvh) {
        // not suitable (too deep for the traversal stack / degenerate input): the host builder below takes over
        if (int rc = build_bvh_on_device(ctx, boxes, leaf_max, &bvh_n_nodes, &bvh_depth, &bvh_on_device)) return rc;
    }
    if (!bvh_on_device) {
        BvhResult bvh = build_bvh(boxes, leaf_max, bt);
        if (bvh.depth > (uint32_t)64) return fail(ctx, 2, "BVH depth %u exceeds the traversal stack", bvh.depth);
        std::vector<DevPrim> leaf_order(n);
        for (uint32_t k = 0; k < n; ++k) leaf_order[k] = flat[bvh.order[k]];
        CK(ctx->prims_bvh.upload(leaf_order));
        CK(ctx->bvh_prim_id.upload(bvh.order));
        CK(ctx->nodes.upload(bvh.nodes));
        bvh_n_nodes = (uint32_t)bvh.nodes.size(); bvh_depth = bvh.depth; bvh_root_is_leaf = bvh.root_is_leaf;
    }
    laps.lap("bvh");
    ctx->stats.ms_bvh_build = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_bvh).count();
    ctx->stats.bvh_builder = bvh_on_device ? RTW_BVH_BUILDER_LBVH : RTW_BVH_BUILDER_SAH;

    // ---- shared-memory image for the flat scan: segmented by kind, small spheres in groups of four --------
    FlatLayout fl{};
    std::vector<float4> blob;
    if (n <= kFlatHardMax) {  // big scenes never run the flat scan: skip the (costly) grouping
        std::vector<uint32_t> stat_ids, mov_ids, big_ids, rect_ids;
        for (uint32_t i = 0; i < n; ++i) {
            const rtw_prim &p = s->prims[i];
            if (p.kind == RTW_PRIM_SPHERE) {
                uint32_t meta;
                std::memcpy(&meta, &flat[i].b.w, 4);
                (((meta >> 8) & 0xFFFu) ? big_ids : stat_ids).push_back(i);  // big_ids in the same order as `bigs`
            } else if (p.kind == RTW_PRIM_MOVING_SPHERE) mov_ids.push_back(i);
            else rect_ids.push_back(i);
        }
        // spatial groups of exactly 4 (the last one may be short): recursive median splits on the widest
        // axis of the centroids, the left half always a multiple of four so no slot is wasted
        // Groups of at most four by recursive median splits on the widest axis of the centroids.  The scan tests bounds
        // four at a time, so the NUMBER of groups is rounded up to a multiple of four (option RTW_GROUP_ROUND=0: exactly
        // ceil(n / 4) groups of four): the bounds that would be padding enclose three spheres instead of nothing.
        const bool round_groups = ctx->opt.num("RTW_GROUP_ROUND", 1) != 0;
        auto make_groups = [&](const std::vector<uint32_t> &sub) {
            std::vector<std::vector<uint32_t>> groups;
            std::vector<uint32_t> work(sub);
            struct Range { size_t lo, hi, g; };  // [lo, hi) becomes g groups
            std::vector<Range> st;
            size_t G = (work.size() + 3) / 4;
            if (round_groups && G >= 3) G = std::min((G + 3) / 4 * 4, work.size());
            if (!work.empty()) st.push_back({0, work.size(), G});
            while (!st.empty()) {
                const Range rg = st.back();
                st.pop_back();
                const size_t cnt = rg.hi - rg.lo;
                if (rg.g <= 1) {
                    std::vector<uint32_t> m(work.begin() + rg.lo, work.begin() + rg.hi);
                    std::sort(m.begin(), m.end());
                    groups.push_back(m);
                    continue;
                }
                double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
                auto cen = [&](uint32_t i, int a) { return 0.5 * (boxes[i].mn[a] + boxes[i].mx[a]); };
                for (size_t k = rg.lo; k < rg.hi; ++k)
                    for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], cen(work[k], a)); mx[a] = std::max(mx[a], cen(work[k], a)); }
                int ax = 0;
                for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
                // g_left of the g groups go left with a proportional share of the spheres, at most four per group on
                // either side
                const size_t gl = rg.g / 2, gr = rg.g - gl;
                size_t left = (cnt * gl + rg.g / 2) / rg.g;
                left = std::min(left, 4 * gl);
                if (cnt - left > 4 * gr) left = cnt - 4 * gr;
                left = std::max<size_t>(left, gl);                    // no empty group
                if (cnt - left < gr) left = cnt - gr;
                std::nth_element(work.begin() + rg.lo, work.begin() + rg.lo + left, work.begin() + rg.hi,
                                 [&](uint32_t x, uint32_t y) { return cen(x, ax) < cen(y, ax); });
                st.push_back({rg.lo, rg.lo + left, gl});
                st.push_back({rg.lo + left, rg.hi, gr});
            }
            return groups;
        };
        // sphere centre of prim i at time t (f64)
        auto centre_at = [&](uint32_t i, double t, double c[3]) {
            const rtw_prim &p = wprims[i];
            if (p.kind == RTW_PRIM_SPHERE) { c[0] = p.v[0]; c[1] = p.v[1]; c[2] = p.v[2]; return std::fabs(p.v[3]); }
            const double u = (t - p.v[6]) / (p.v[7] - p.v[6]);
            for (int a = 0; a < 3; ++a) c[a] = p.v[a] + (p.v[3 + a] - p.v[a]) * u;
            return std::fabs(p.v[8]);
        };
        // Bound of a group = the smallest ball enclosing its members over the shutter interval (both end positions of a
        // moving sphere; the sweep between them is a straight segment, covered by convexity).  Badoiu-Clarkson iteration
        // from the centroid: step towards the farthest member by 1/(k+1).  A vote passes a group roughly in proportion to
        // the cross-section R^2 of its bound, so a tighter ball is fewer member tests.
        struct Ball { double c[3], r; };
        auto balls_of = [&](const std::vector<uint32_t> &m, Ball *balls) {  // <= 8 balls: 4 members x both shutter ends
            size_t nb = 0;
            for (uint32_t i : m)
                for (double t : {s->time0, s->time1}) {
                    Ball b;
                    b.r = centre_at(i, t, b.c);
                    if (nb && balls[nb - 1].r == b.r && balls[nb - 1].c[0] == b.c[0] && balls[nb - 1].c[1] == b.c[1] && balls[nb - 1].c[2] == b.c[2])
                        continue;  // a static sphere is one ball, not two
                    balls[nb++] = b;
                }
            return nb;
        };
        auto ball_of = [&](const std::vector<uint32_t> &m, double cen[3], int iters = 400) {
            Ball balls[8];
            const size_t nb = balls_of(m, balls);
            cen[0] = cen[1] = cen[2] = 0;
            for (size_t k = 0; k < nb; ++k) for (int a = 0; a < 3; ++a) cen[a] += balls[k].c[a] / (double)nb;
            auto radius_at = [&](const double c[3], size_t *far) {
                double R = 0;
                for (size_t k = 0; k < nb; ++k) {
                    const double d = std::sqrt((balls[k].c[0] - c[0]) * (balls[k].c[0] - c[0]) + (balls[k].c[1] - c[1]) * (balls[k].c[1] - c[1]) +
                                               (balls[k].c[2] - c[2]) * (balls[k].c[2] - c[2])) + balls[k].r;
                    if (d > R) { R = d; if (far) *far = k; }
                }
                return R;
            };
            double best[3] = {cen[0], cen[1], cen[2]}, best_R = radius_at(cen, nullptr), c[3] = {cen[0], cen[1], cen[2]};
            for (int k = 1; k <= iters && nb; ++k) {
                size_t far = 0;
                const double R = radius_at(c, &far);
                if (R < best_R) { best_R = R; best[0] = c[0]; best[1] = c[1]; best[2] = c[2]; }
                const Ball &f = balls[far];
                const double d = R - f.r;  // distance to the far ball's centre
                if (!(d > 0)) break;
                // the point of the far ball farthest from c lies on the ray c -> f.c at distance R
                const double step = 1.0 / (k + 1.0);
                for (int a = 0; a < 3; ++a) c[a] += step * (f.c[a] - c[a]) * (R / d);
            }
            cen[0] = best[0]; cen[1] = best[1]; cen[2] = best[2];
            return best_R;
        };
        // No enclosing ball is smaller than half the extent of any two members: a lower bound that costs 28 distances
        // and rejects almost every exchange candidate of refine_groups before the iteration above runs.
        auto ball_lower_bound = [&](const std::vector<uint32_t> &m) {
            Ball balls[8];
            const size_t nb = balls_of(m, balls);
            double lb = 0;
            for (size_t i = 0; i < nb; ++i) {
                lb = std::max(lb, balls[i].r);
                for (size_t j = i + 1; j < nb; ++j) {
                    const double d = std::sqrt((balls[i].c[0] - balls[j].c[0]) * (balls[i].c[0] - balls[j].c[0]) + (balls[i].c[1] - balls[j].c[1]) * (balls[i].c[1] - balls[j].c[1]) +
                                               (balls[i].c[2] - balls[j].c[2]) * (balls[i].c[2] - balls[j].c[2]));
                    lb = std::max(lb, 0.5 * (d + balls[i].r + balls[j].r));
                }
            }
            return lb;
        };
        auto bound_of = [&](const std::vector<uint32_t> &m) {
            double cen[3];
            double R = ball_of(m, cen);
            R = R * (1.0 + 1e-5) + 1e-6;  // fp32 evaluation slack
            const float cf[3] = {(float)cen[0], (float)cen[1], (float)cen[2]};
            for (int a = 0; a < 3; ++a) R += std::fabs((double)cf[a] - cen[a]);
            return make_float4(cf[0], cf[1], cf[2], (float)(R * (1.0 + 1e-6)));  // the radius, not its square
        };
        // Median splits give spatially compact groups; a few passes of pairwise member exchanges between groups then
        // lower the sum of R^2 (deterministic: fixed order, strict improvement only).
        auto refine_groups = [&](std::vector<std::vector<uint32_t>> &groups) {
            if (groups.size() < 2 || groups.size() > 64) return;
            std::vector<double> R2(groups.size());
            double cen[3];
            for (size_t g = 0; g < groups.size(); ++g) { const double R = ball_of(groups[g], cen); R2[g] = R * R; }
            for (int pass = 0; pass < 4; ++pass) {
                bool improved = false;
                for (size_t g = 0; g < groups.size(); ++g)
                    for (size_t h2 = g + 1; h2 < groups.size(); ++h2)
                        for (size_t a = 0; a < groups[g].size(); ++a)
                            for (size_t b2 = 0; b2 < groups[h2].size(); ++b2) {
                                std::swap(groups[g][a], groups[h2][b2]);
                                const double now = (R2[g] + R2[h2]) * (1.0 - 1e-9);
                                bool better = false;
                                const double Lg = ball_lower_bound(groups[g]), Lh = ball_lower_bound(groups[h2]);
                                if (Lg * Lg + Lh * Lh < now) {
                                    // every iterate's radius is an upper bound of the optimum: 40 steps screen, 400 decide
                                    const double Sg = ball_of(groups[g], cen, 40), Sh = ball_of(groups[h2], cen, 40);
                                    if (Sg * Sg + Sh * Sh < now * 1.1) {
                                        const double Rg = ball_of(groups[g], cen), Rh = ball_of(groups[h2], cen);
                                        if (Rg * Rg + Rh * Rh < now) { R2[g] = Rg * Rg; R2[h2] = Rh * Rh; improved = true; better = true; }
                                    }
                                }
                                if (!better) std::swap(groups[g][a], groups[h2][b2]);
                            }
                if (!improved) break;
            }
            for (auto &g : groups) std::sort(g.begin(), g.end());
        };
        // Static and moving spheres are grouped TOGETHER, by position alone (a static sphere is a moving one with zero
        // velocity: DevPrim::b.xyz = 0, and fma(0, time, c) = c exactly, so the BVH leaves' test stays bit-identical):
        // tighter bounds than per-kind groups over the same cells, one member code path in the scan, fewer padding slots.
        std::vector<uint32_t> small_ids(stat_ids);
        small_ids.insert(small_ids.end(), mov_ids.begin(), mov_ids.end());
        std::sort(small_ids.begin(), small_ids.end());
        std::vector<std::vector<uint32_t>> sgroups;  // (no separate static groups any more; the layout keeps the slot)
        auto mgroups = make_groups(small_ids);
        refine_groups(mgroups);
        std::vector<float4> sph, big, mov, rect, bounds;
        std::vector<uint32_t> ids;
        const float4 dummy = make_float4(0.f, 0.f, 0.f, -1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (const auto &g : sgroups) {
            bounds.push_back(bound_of(g));
            for (int k = 0; k < 4; ++k) {
                if (k < (int)g.size()) {
                    const DevPrim &d = flat[g[k]];
                    sph.push_back(make_float4(d.a.x, d.a.y, d.a.z, d.a.w * d.a.w));
                    ids.push_back(g[k]);
                } else { sph.push_back(dummy); ids.push_back(0); }
            }
        }
        for (uint32_t i : big_ids) {
            const DevPrim &d = flat[i];
            big.push_back(make_float4(d.a.x, d.a.y, d.a.z, d.a.w * d.a.w));
            ids.push_back(i);
        }
        while (ids.size() & 3) ids.push_back(0);
        for (const auto &g : mgroups) {
            bounds.push_back(bound_of(g));
            for (int k = 0; k < 4; ++k) {
                if (k < (int)g.size()) {
                    const DevPrim &d = flat[g[k]];
                    mov.push_back(make_float4(d.a.x, d.a.y, d.a.z, d.a.w * d.a.w));
                    mov.push_back(make_float4(d.b.x, d.b.y, d.b.z, d.a.w));
                    ids.push_back(g[k]);
                } else { mov.push_back(dummy); mov.push_back(zero4); ids.push_back(0); }
            }
        }
        // boxes: rects of one instance that are faces of ONE axis-aligned box (exact equality of the f64 fields) are
        // scanned as three slabs (box_face in rtw_trace.cuh).  Found from pairs of differently oriented rects, which fix
        // all six numbers; a box needs at least three faces to be worth a record.  Option RTW_BOX_PRIMS=0 turns it off.
        std::vector<float4> boxes4;
        std::vector<char> boxed(n, 0);
        if (ctx->opt.num("RTW_BOX_PRIMS", 1) != 0) {
            struct Cand { double x0, x1, y0, y1, z0, z1; };
            auto ranges = [&](uint32_t i, double (&lo)[3], double (&hi)[3], int &thin, double &k) {
                const rtw_prim &p = s->prims[i];  // in-plane axes (a, b) and the thin axis per orientation
                const int ax = p.kind == RTW_PRIM_XY_RECT ? 0 : p.kind == RTW_PRIM_XZ_RECT ? 0 : 1;
                const int bx = p.kind == RTW_PRIM_XY_RECT ? 1 : 2;
                thin = 3 - ax - bx;
                for (int a = 0; a < 3; ++a) { lo[a] = NAN; hi[a] = NAN; }
                lo[ax] = p.v[0]; hi[ax] = p.v[1]; lo[bx] = p.v[2]; hi[bx] = p.v[3];
                k = p.v[4];
            };
            auto face_of = [&](uint32_t i, const Cand &c) -> int {  // face slot 0..5 (z1, z0, y1, y0, x1, x0) or -1
                double lo[3], hi[3], k; int thin;
                ranges(i, lo, hi, thin, k);
                const double clo[3] = {c.x0, c.y0, c.z0}, chi[3] = {c.x1, c.y1, c.z1};
                for (int a = 0; a < 3; ++a)
                    if (a != thin && !(lo[a] == clo[a] && hi[a] == chi[a])) return -1;
                if (k == chi[thin]) return 2 * (2 - thin);
                if (k == clo[thin]) return 2 * (2 - thin) + 1;
                return -1;
            };
            for (size_t ia = 0; ia < rect_ids.size(); ++ia) {
                const uint32_t a = rect_ids[ia];
                if (boxed[a]) continue;
                for (size_t ib = ia + 1; ib < rect_ids.size(); ++ib) {
                    const uint32_t b = rect_ids[ib];
                    if (boxed[b] || s->prims[b].kind == s->prims[a].kind || s->prims[b].xform != s->prims[a].xform) continue;
                    double la[3], ha[3], lb[3], hb[3], ka, kb; int ta, tb;
                    ranges(a, la, ha, ta, ka); ranges(b, lb, hb, tb, kb);
                    Cand c;
                    double lo[3], hi[3];
                    bool ok = true;
                    for (int x = 0; x < 3; ++x) {  // every axis is in-plane for at least one of two differently oriented rects
                        const bool ina = x != ta, inb = x != tb;
                        if (ina && inb && !(la[x] == lb[x] && ha[x] == hb[x])) ok = false;
                        lo[x] = ina ? la[x] : lb[x]; hi[x] = ina ? ha[x] : hb[x];
                    }
                    if (!ok || !(lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2])) continue;
                    c = Cand{lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]};
                    if (face_of(a, c) < 0 || face_of(b, c) < 0) continue;
                    uint32_t ids6[6], mask = 0;
                    for (size_t ic = ia; ic < rect_ids.size(); ++ic) {
                        const uint32_t r3 = rect_ids[ic];
                        if (boxed[r3] || s->prims[r3].xform != s->prims[a].xform) continue;
                        const int f = face_of(r3, c);
                        if (f >= 0 && !(mask & (1u << f))) { mask |= 1u << f; ids6[f] = r3; }
                    }
                    if (__builtin_popcount(mask) < 3) continue;
                    for (int f = 0; f < 6; ++f) if (mask & (1u << f)) boxed[ids6[f]] = 1; else ids6[f] = kMiss;  // absent face: id -1 never wins
                    const int slot = s->prims[a].xform >= 0 ? xform_slot_of(s->prims[a].xform) : -1;
                    boxes4.push_back(make_float4((float)c.x0, (float)c.x1, (float)c.y0, (float)c.y1));
                    boxes4.push_back(make_float4((float)c.z0, (float)c.z1, bits_to_float((uint32_t)(slot + 1)), bits_to_float(mask)));
                    boxes4.push_back(make_float4(bits_to_float(ids6[0]), bits_to_float(ids6[1]), bits_to_float(ids6[2]), bits_to_float(ids6[3])));
                    const DevXform bxf = slot >= 0 ? xforms[slot] : DevXform{1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    boxes4.push_back(make_float4(bits_to_float(ids6[4]), bits_to_float(ids6[5]), bxf.c, bxf.s));
                    boxes4.push_back(make_float4(bxf.tx, bxf.ty, bxf.tz, 0.f));
                    break;
                }
            }
            rect_ids.erase(std::remove_if(rect_ids.begin(), rect_ids.end(), [&](uint32_t i) { return boxed[i] != 0; }), rect_ids.end());
        }
        // rect runs: equal (xform slot, orientation); transforms in order of first appearance, the runs of one
        // transform adjacent so the scan sets up the object-space ray once per instance
        std::vector<float4> runs;
        {
            std::vector<std::pair<int, uint32_t>> keys;  // (xform slot, kind)
            std::vector<std::vector<uint32_t>> members;
            for (uint32_t i : rect_ids) {
                int slot; uint32_t meta;
                std::memcpy(&slot, &flat[i].b.y, 4); std::memcpy(&meta, &flat[i].b.w, 4);
                const std::pair<int, uint32_t> key{slot, meta & 0xFFu};
                size_t k = 0;
                for (; k < keys.size(); ++k) if (keys[k] == key) break;
                if (k == keys.size()) { keys.push_back(key); members.emplace_back(); }
                members[k].push_back(i);
            }
            std::vector<size_t> order_k(keys.size());
            for (size_t k = 0; k < keys.size(); ++k) order_k[k] = k;
            auto first_with_slot = [&](int slot) { size_t k = 0; while (keys[k].first != slot) ++k; return k; };
            std::stable_sort(order_k.begin(), order_k.end(), [&](size_t a, size_t b) {
                return first_with_slot(keys[a].first) < first_with_slot(keys[b].first);
            });
            int prev_slot = 0;
            bool have_prev = false;
            for (size_t k : order_k) {
                const uint32_t first = (uint32_t)(rect.size() / 2);
                for (uint32_t i : members[k]) {
                    rect.push_back(flat[i].a);
                    rect.push_back(make_float4(flat[i].b.x, bits_to_float(i), 0.f, 0.f));
                }
                float4 rd;
                const bool same = have_prev && prev_slot == keys[k].first;
                const uint32_t w[4] = {(uint32_t)(keys[k].first + 1) | (same ? kRunSameXform : 0u), keys[k].second, first,
                                       (uint32_t)members[k].size()};
                std::memcpy(&rd, w, 16);
                runs.push_back(rd);
                prev_slot = keys[k].first;
                have_prev = true;
            }
        }
        fl.flags = sgroups.size() + mgroups.size() < 3 ? kFlatNoBounds : 0u;
        fl.n_sph_groups = (uint32_t)sgroups.size(); fl.n_big = (uint32_t)big_ids.size();
        fl.n_mov_groups = (uint32_t)mgroups.size(); fl.n_rect = (uint32_t)rect_ids.size();
        fl.off_sph = 0;
        fl.off_big = fl.off_sph + (uint32_t)sph.size();
        fl.off_mov = fl.off_big + (uint32_t)big.size();
        fl.off_rect = fl.off_mov + (uint32_t)mov.size();
        fl.off_runs = fl.off_rect + (uint32_t)rect.size();
        fl.n_runs = (uint32_t)runs.size();
        fl.off_boxes = fl.off_runs + (uint32_t)runs.size();
        fl.n_boxes = (uint32_t)(boxes4.size() / kBoxF4);
        while (bounds.size() & 3) bounds.push_back(make_float4(0.f, 0.f, 0.f, NAN));  // the scan tests bounds four at a time
        fl.off_bounds = fl.off_boxes + (uint32_t)boxes4.size();
        fl.off_ids = fl.off_bounds + (uint32_t)bounds.size();
        blob.insert(blob.end(), sph.begin(), sph.end());
        blob.insert(blob.end(), big.begin(), big.end());
        blob.insert(blob.end(), mov.begin(), mov.end());
        blob.insert(blob.end(), rect.begin(), rect.end());
        blob.insert(blob.end(), runs.begin(), runs.end());
        blob.insert(blob.end(), boxes4.begin(), boxes4.end());
        blob.insert(blob.end(), bounds.begin(), bounds.end());
        const size_t at = blob.size();
        blob.resize(at + ids.size() / 4);
        if (!ids.empty()) std::memcpy(&blob[at], ids.data(), ids.size() * 4);
        fl.total_f4 = (uint32_t)blob.size();
    }

    laps.lap("flat-scan image");
    // ---- materials / textures -------------------------------------------------------------------------
    std::vector<DevMaterial> mats(s->n_materials);
    for (uint32_t i = 0; i < s->n_materials; ++i) {
        const rtw_material &m = s->materials[i];
        mats[i] = DevMaterial{m.kind, m.texture, (float)m.albedo[0], (float)m.albedo[1], (float)m.albedo[2], (float)m.param, 0.f, 0.f};
    }
    std::vector<DevTexture> texs(s->n_textures);
    for (uint32_t i = 0; i < s->n_textures; ++i) {
        const rtw_texture &t = s->textures[i];
        texs[i] = DevTexture{t.kind, t.a, t.b, (float)t.scale, (float)t.color[0], (float)t.color[1], (float)t.color[2], 0u};
    }
    std::vector<DevPerlin> perl(s->n_perlins);
    for (uint32_t i = 0; i < s->n_perlins; ++i) {
        const rtw_perlin &p = s->perlins[i];
        if (!p.ranvec || !p.perm_x || !p.perm_y || !p.perm_z) return fail(ctx, 1, "perlin %u: null table", i);
        for (int k = 0; k < 256; ++k) {
            perl[i].ranvec[k] = make_float4((float)p.ranvec[3 * k], (float)p.ranvec[3 * k + 1], (float)p.ranvec[3 * k + 2], 0.f);
            perl[i].perm[0][k] = (uint8_t)p.perm_x[k]; perl[i].perm[1][k] = (uint8_t)p.perm_y[k]; perl[i].perm[2][k] = (uint8_t)p.perm_z[k];
        }
    }
    free_images(ctx);
    std::vector<DevImage> imgs(s->n_images);
    for (uint32_t i = 0; i < s->n_images; ++i) {
        const rtw_image &im = s->images[i];
        cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        CK(cudaMallocArray(&arr, &fd, im.width, im.height));
        ctx->image_arrays.push_back(arr);
        CK(cudaMemcpy2DToArray(arr, 0, 0, im.rgba8, (size_t)im.width * 4, (size_t)im.width * 4, im.height, cudaMemcpyHostToDevice));
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;  // nearest texel, as texture.zig:126-133; bilinear would smear the alpha test
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaTextureObject_t tex = 0;
        CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        ctx->image_tex.push_back(tex);
        imgs[i] = DevImage{tex, im.width, im.height, 0, 0};
    }

    CK(ctx->flat_blob.upload(blob));
    CK(ctx->prim_material.upload(prim_mat));
    CK(ctx->xforms.upload(xforms));
    CK(ctx->bigs.upload(bigs));
    CK(ctx->materials.upload(mats));
    CK(ctx->textures.upload(texs));
    CK(ctx->images.upload(imgs));
    CK(ctx->perlins.upload(perl));
    ctx->raw_prims.release(); ctx->raw_chains.release();  // reference-order probe tables: built on first use
    ctx->raw = RawScene{};
    ctx->host_inst_orig.clear();
    for (uint32_t i = 0; i < n; ++i)
        if (wprims[i].xform >= 0 && (wprims[i].kind == RTW_PRIM_SPHERE || wprims[i].kind == RTW_PRIM_MOVING_SPHERE)) {
            std::array<double, 10> v;
            std::memcpy(v.data(), s->prims[i].v, sizeof(double) * 10);
            ctx->host_inst_orig.emplace_back(i, v);
        }
    ctx->host_prims = std::move(wprims);
    ctx->host_xforms.assign(s->xforms, s->xforms + s->n_xforms);
    laps.lap("tables + upload");

    DevScene &d = ctx->scene;
    d.flat_blob = ctx->flat_blob.p; d.flat = fl;
    d.prims_flat = ctx->prims_flat.p; d.prims_bvh = ctx->prims_bvh.p; d.bvh_prim_id = ctx->bvh_prim_id.p;
    d.prim_material = ctx->prim_material.p; d.nodes = ctx->nodes.p; d.xforms = ctx->xforms.p; d.bigs = ctx->bigs.p;
    d.materials = ctx->materials.p; d.textures = ctx->textures.p; d.images = ctx->images.p; d.perlins = ctx->perlins.p;
    d.n_prims = n; d.n_nodes = bvh_n_nodes; d.n_xforms = (uint32_t)xforms.size();
    d.n_materials = s->n_materials; d.n_textures = s->n_textures;
    d.root_is_leaf = bvh_root_is_leaf ? 1u : 0u;
    ctx->n_prims = n;
    ctx->root_is_leaf = bvh_root_is_leaf;
    ctx->stats.bvh_nodes = bvh_n_nodes;
    ctx->stats.bvh_depth = bvh_depth;
    CK(cudaDeviceSynchronize());
    ctx->stats.ms_upload = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    ctx->have_scene = true;
    return 0;
}

// the launch-independent part of DevRender: image size, sample range, seed -> Philox round keys, background
static DevRender base_render(const rtw_render_params *p) {
    DevRender rp{};
    rp.width = p->width; rp.height = p->height;
    rp.spp_begin = p->spp_begin; rp.spp_end = p->spp_end;
    rp.max_depth = p->max_depth;
    rp.seed_lo = (uint32_t)p->seed; rp.seed_hi = (uint32_t)(p->seed >> 32);
    for (uint32_t i = 0; i < 10; ++i) {
        rp.philox_keys[2 * i] = rp.seed_lo + i * 0x9E3779B9u;
        rp.philox_keys[2 * i + 1] = rp.seed_hi + i * 0xBB67AE85u;
    }
    rp.bg_r = (float)p->background[0]; rp.bg_g = (float)p->background[1]; rp.bg_b = (float)p->background[2];
    rp.inv_wm1 = (float)(1.0 / ((double)p->width - 1.0)); rp.inv_hm1 = (float)(1.0 / ((double)p->height - 1.0));
    return rp;
}

static int accumulate_impl(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *p, float *d_accum,
                           cudaStream_t st, bool timed) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!ctx->have_scene) return fail(ctx, 1, "no scene uploaded");
    if (!cam || !p || !d_accum) return fail(ctx, 1, "null argument");
    if (p->width == 0 || p->height == 0) return fail(ctx, 1, "empty image");
    if (p->spp_end < p->spp_begin) return fail(ctx, 1, "spp_end < spp_begin");
    if ((uint64_t)p->width * p->height > 0x7FFFFFFFull || p->width > 0xFFFFu || p->height > 0xFFFFu) return fail(ctx, 1, "image too large");
    int variant = 0;
    if (int rc = pick_variant(ctx, p->variant, &variant)) return rc;
    CK(cudaSetDevice(ctx->device));
    const bool stats = (p->flags & RTW_FLAG_COUNT_EVENTS) != 0;

    if (p->variant == RTW_VARIANT_WAVEFRONT && (p->max_depth > 63 || p->spp_end > (1u << 26)))
        return fail(ctx, 1, "wavefront variant packs (sample, bounce) into 26 + 6 bits: max_depth <= 63 and spp_end <= 2^26");
    DevRender rp = base_render(p);
    rp.accum = reinterpret_cast<float4 *>(d_accum);
    rp.tile_counter = ctx->tile_counter.p;
    rp.stats = ctx->stat_counters.p;
    rp.tiles_x = (p->width + 7) / 8;
    rp.n_tiles = rp.tiles_x * ((p->height + 3) / 4);

    // RTW_FLAG_DETERMINISTIC (or RTW_SPP_CHUNK=0): lane-owns-pixel kernel, one chunk, fixed summation order.
    // Default: pooled kernel (warp-level path queue + one vector atomic per path).
    const char *env = ctx->opt.get("RTW_SPP_CHUNK");
    const bool env_set = env && *env;
    // option RTW_FLAT_KERNEL=1: the first schedule of the pooled flat kernel (kept for A/B measurements)
    // 1 first schedule, 2 second schedule at 9 CTAs/SM (56 registers, spills), 3 (default) second schedule at 8 CTAs/SM
    // (64 registers, no spills): measured 42.8 / 40.6 / 39.9 ms on scene 1 (1080p x 200 spp), 28.8 / 27.2 / 26.2 ms on C3
    const long fk = ctx->opt.num("RTW_FLAT_KERNEL", 3);
    int pooled = ((p->flags & RTW_FLAG_DETERMINISTIC) || env_set) ? 0 : (fk == 1 ? 1 : fk == 2 ? 2 : 3);
    // the default flat kernel exists in three specialisations on the scene's features (code for absent primitive and
    // texture kinds compiled out); RTW_FLAT_SPECIALISE=0 runs the generic one.  Event-counter builds are generic.
    if (pooled == 3 && !stats && ctx->opt.num("RTW_FLAT_SPECIALISE", 1) != 0) {  // (the BVH megakernel has the spheres-only specialisation)
        const uint32_t f = ctx->flat_feat;
        if ((f & ~(uint32_t)FF_SPHERES) == 0) pooled = 4;
        else if ((f & ~(uint32_t)(FF_SPHERES | FF_TEX)) == 0) pooled = 5;
        else if ((f & ~(uint32_t)FF_RECTS) == 0) pooled = 6;
    }
    // option RTW_BVH_KERNEL: 1 = per-lane state machine, 2 = per-warp ray queue in shared memory (k_megakernel_bvhq)
    // 3 = the state machine with speculative traversal (a lane postpones one leaf and goes on descending)
    if (variant == VAR_BVH && pooled >= 1) {
        const long bk = ctx->opt.num("RTW_BVH_KERNEL", 1);
        if (bk == 2) pooled = pooled == 4 ? 8 : 7;
        else if (bk == 3) pooled = pooled == 4 ? 10 : 9;
    }
    const int per_sm = megakernel_ctas_per_sm(variant, stats, pooled, ctx->scene);
    if (per_sm <= 0) return fail(ctx, 2, "megakernel does not fit on an SM (flat image %u B)", ctx->scene.flat.total_f4 * 16u);
    const int grid = per_sm * ctx->n_sms;

    const uint32_t spp = p->spp_end - p->spp_begin;
    uint32_t chunk = spp;
    if (env_set) {
        const long v = atol(env);
        chunk = v <= 0 ? spp : (uint32_t)v;
    }
    if (chunk == 0) chunk = 1;
    rp.spp_chunk = chunk;
    rp.n_chunks = spp ? (spp + chunk - 1) / chunk : 0;
    // Batch = one 8x4 tile x batch_spp samples.  Large frames use 64 samples per batch; small frames shrink the
    // batch so that every resident warp still gets >= ~24 batches (otherwise the launch is all tail: the
    // reference's own 600x400x50 frame would be 1.4 batches per warp).
    uint32_t batch_spp = kBatchSpp;
    if (spp > 0) {
        const uint64_t warps = (uint64_t)grid * 4, want = warps * 24;
        const uint64_t sblocks_wanted = (want + rp.n_tiles - 1) / rp.n_tiles;
        if (sblocks_wanted > 1) batch_spp = (uint32_t)std::max<uint64_t>(2, std::min<uint64_t>(kBatchSpp, spp / sblocks_wanted));
        if (ctx->opt.get("RTW_BATCH_SPP")) batch_spp = (uint32_t)std::max(1l, std::min(4096l, ctx->opt.num("RTW_BATCH_SPP", batch_spp)));
    }
    rp.batch_spp = batch_spp;
    rp.n_sblocks = (spp + batch_spp - 1) / batch_spp;
    if ((uint64_t)rp.n_sblocks * rp.n_tiles > 0xFFFFFFF0ull) return fail(ctx, 1, "too many path batches for one launch");
    rp.n_batches = rp.n_sblocks * rp.n_tiles;
    rp.service_threshold = 24; rp.steps_per_round = 3; rp.leaf_threshold = 4;  // swept on the 10^6- and the 485-sphere scene (tools/bvh_sweep.py): -3 % / -1 % vs (20, 2, 8)
    rp.leaf_threshold = (uint32_t)std::max(1l, std::min(32l, ctx->opt.num("RTW_BVH_LEAF", rp.leaf_threshold)));
    rp.service_threshold = (uint32_t)std::max(1l, std::min(32l, ctx->opt.num("RTW_BVH_THRESH", rp.service_threshold)));
    rp.steps_per_round = (uint32_t)std::max(1l, std::min(64l, ctx->opt.num("RTW_BVH_STEPS", rp.steps_per_round)));

    if (spp == 0) { ctx->stats.n_launches = 0; return 0; }
    // One work queue and one set of event counters per context: launches of one context are ordered ON THE DEVICE, whatever
    // streams the caller uses (a second rtw_cuda_accumulate on another stream waits for the first; it does not race).
    if (ctx->launch_recorded) CK(cudaStreamWaitEvent(st, ctx->ev_launch, 0));
    if (stats) CK(cudaMemsetAsync(ctx->stat_counters.p, 0, ST_COUNT * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(unsigned int), st));
    const DevCamera dc = lower_camera(cam);
    if (timed) CK(cudaEventRecord(ctx->ev[0], st));
    if (p->max_depth == 0) {
        // rayColor returns black before intersecting anything when depth == 0 (main.zig:105-108): every sample
        // contributes (0, 0, 0); only the sample count moves
        CK(launch_add_samples(rp.accum, p->width * p->height, (float)spp, st));
        if (timed) CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaEventRecord(ctx->ev_launch, st));
        ctx->launch_recorded = true;
        ctx->stats.n_launches = 1;
        ctx->stats.variant_used = p->variant;
        return 0;
    }
    if (p->variant == RTW_VARIANT_WAVEFRONT) {
        // K2: host loop of generate / extend / shade launches over the path-state slots (blocks the caller)
        uint32_t slots = 1u << 21;
        slots = (uint32_t)std::max(128l, ctx->opt.num("RTW_WF_SLOTS", slots));
        slots = (slots + 127u) & ~127u;
        if (ctx->wf_slots != slots) {
            CK(ctx->wf_ro.alloc(slots)); CK(ctx->wf_rd.alloc(slots)); CK(ctx->wf_beta.alloc(slots)); CK(ctx->wf_rad.alloc(slots));
            CK(ctx->wf_hit.alloc(slots)); CK(ctx->wf_free0.alloc(slots)); CK(ctx->wf_free1.alloc(slots));
            CK(ctx->wf_cursor.alloc(1)); CK(ctx->wf_counters.alloc(1));
            if (!ctx->wf_host_counters) CK(cudaMallocHost(&ctx->wf_host_counters, sizeof(WfCounters)));
            ctx->wf_slots = slots;
        }
        WfState ws{};
        ws.ro = ctx->wf_ro.p; ws.rd = ctx->wf_rd.p; ws.beta = ctx->wf_beta.p; ws.rad = ctx->wf_rad.p; ws.hit = ctx->wf_hit.p;
        ws.free_list[0] = ctx->wf_free0.p; ws.free_list[1] = ctx->wf_free1.p;
        ws.counters = ctx->wf_counters.p; ws.extend_cursor = ctx->wf_cursor.p; ws.n_slots = slots;
        uint32_t nl = 0;
        CK(wavefront_accumulate(ws, variant, stats, ctx->scene, dc, rp, ctx->n_sms, ctx->wf_host_counters, st, &nl));
        if (timed) CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaEventRecord(ctx->ev_launch, st));
        ctx->launch_recorded = true;
        ctx->stats.n_launches = nl;
        ctx->stats.variant_used = RTW_VARIANT_WAVEFRONT;
        return 0;
    }
    CK(launch_megakernel(variant, stats, pooled, ctx->scene, dc, rp, grid, st));
    if (timed) CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaEventRecord(ctx->ev_launch, st));
    ctx->launch_recorded = true;
    ctx->stats.n_launches = 1;
    ctx->stats.variant_used = variant == VAR_FLAT ? RTW_VARIANT_MEGA_FLAT : RTW_VARIANT_MEGA_BVH;
    return 0;
}

static int fetch_counters(rtw_ctx *ctx) {
    unsigned long long h[ST_COUNT];
    CK(cudaMemcpy(h, ctx->stat_counters.p, sizeof h, cudaMemcpyDeviceToHost));
    rtw_stats &s = ctx->stats;
    s.paths = h[ST_PATHS]; s.rays = h[ST_RAYS]; s.node_tests = h[ST_NODE_TESTS]; s.sphere_tests = h[ST_SPHERE_TESTS];
    s.sphere_roots = h[ST_SPHERE_ROOTS]; s.moving_tests = h[ST_MOVING_TESTS]; s.rect_tests = h[ST_RECT_TESTS];
    s.rect_accepts = h[ST_RECT_ACCEPTS]; s.xform_apps = h[ST_XFORM_APPS]; s.sphere_finalise = h[ST_SPHERE_FINAL];
    s.scatter_diffuse = h[ST_SC_DIFFUSE]; s.scatter_metal = h[ST_SC_METAL]; s.scatter_dielectric = h[ST_SC_DIELECTRIC];
    s.emit_hits = h[ST_EMIT]; s.tex_checker = h[ST_TEX_CHECKER]; s.tex_image = h[ST_TEX_IMAGE]; s.tex_noise = h[ST_TEX_NOISE];
    s.nan_pixels = h[ST_NAN_PIXELS];
    return 0;
}

int rtw_cuda_accumulate(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *p, float *d_accum, void *stream) {
    return accumulate_impl(ctx, cam, p, d_accum, (cudaStream_t)stream, false);
}

int rtw_cuda_resolve_multi(rtw_ctx *ctx, const float *const *d_accums, uint32_t n_bufs, uint32_t width, uint32_t height,
                           uint32_t spp_total, uint8_t *d_rgb8, void *stream) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!d_accums || !d_rgb8 || n_bufs == 0 || n_bufs > kMaxResolveBufs) return fail(ctx, 1, "bad buffer list");
    if (width == 0 || height == 0 || spp_total == 0) return fail(ctx, 1, "empty image or spp_total == 0");
    CK(cudaSetDevice(ctx->device));
    ResolveArgs a{};
    for (uint32_t i = 0; i < n_bufs; ++i) {
        if (!d_accums[i]) return fail(ctx, 1, "accumulation buffer %u is null", i);
        a.bufs[i] = reinterpret_cast<const float4 *>(d_accums[i]);
    }
    a.n_bufs = n_bufs; a.width = width; a.height = height;
    a.row_begin = 0; a.row_end = height;
    a.scale = 1.0f / (float)spp_total;
    a.rgb8 = d_rgb8;
    a.nan_counter = ctx->stat_counters.p + ST_NAN_PIXELS;
    CK(cudaMemsetAsync(a.nan_counter, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    CK(launch_resolve(a, (cudaStream_t)stream));
    return 0;
}

int rtw_cuda_resolve(rtw_ctx *ctx, const float *d_accum, uint32_t width, uint32_t height, uint32_t spp_total,
                     uint8_t *d_rgb8, void *stream) {
    const float *bufs[1] = {d_accum};
    return rtw_cuda_resolve_multi(ctx, bufs, 1, width, height, spp_total, d_rgb8, stream);
}

int rtw_cuda_render(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *p, uint8_t *rgb8_out, float *accum_out) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    const auto t_begin = std::chrono::steady_clock::now();
    if (!p || !rgb8_out) return fail(ctx, 1, "null argument");
    if (p->width == 0 || p->height == 0) return fail(ctx, 1, "empty image");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)p->width * p->height;
    if (ctx->accum.n != npx) { CK(ctx->accum.alloc(npx)); CK(ctx->rgb8.alloc(npx * 3)); }
    cudaStream_t st = ctx->stream;
    CK(cudaMemsetAsync(ctx->accum.p, 0, npx * sizeof(float4), st));
    if (int rc = accumulate_impl(ctx, cam, p, reinterpret_cast<float *>(ctx->accum.p), st, true)) return rc;
    const uint32_t spp = p->spp_end - p->spp_begin;
    const uint32_t spp_total = p->spp_total ? p->spp_total : spp;
    if (spp_total == 0) return fail(ctx, 1, "no samples to resolve");
    CK(cudaEventRecord(ctx->ev[2], st));
    if (int rc = rtw_cuda_resolve(ctx, reinterpret_cast<const float *>(ctx->accum.p), p->width, p->height, spp_total, ctx->rgb8.p, st)) return rc;
    CK(cudaEventRecord(ctx->ev[3], st));
    CK(cudaMemcpyAsync(rgb8_out, ctx->rgb8.p, npx * 3, cudaMemcpyDeviceToHost, st));
    if (accum_out) CK(cudaMemcpyAsync(accum_out, ctx->accum.p, npx * sizeof(float4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (spp > 0) { CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->stats.ms_trace = ms; } else ctx->stats.ms_trace = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    ctx->stats.ms_resolve = ms;
    ctx->stats.n_launches += 1;
    if (int rc = fetch_counters(ctx)) return rc;
    if (!(p->flags & RTW_FLAG_COUNT_EVENTS)) {
        const rtw_stats keep = ctx->stats;  // no event counters in this mode: keep timings and scene facts only
        ctx->stats = rtw_stats{};
        ctx->stats.nan_pixels = keep.nan_pixels; ctx->stats.ms_upload = keep.ms_upload; ctx->stats.ms_trace = keep.ms_trace;
        ctx->stats.ms_resolve = keep.ms_resolve; ctx->stats.n_launches = keep.n_launches; ctx->stats.variant_used = keep.variant_used;
        ctx->stats.bvh_nodes = keep.bvh_nodes; ctx->stats.bvh_depth = keep.bvh_depth;
        ctx->stats.ms_bvh_build = keep.ms_bvh_build; ctx->stats.bvh_builder = keep.bvh_builder;
        ctx->stats.paths = (uint64_t)npx * spp;
    }
    ctx->stats.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return 0;
}

// Resolve rows [row_begin, row_end) of the sum of `bufs` on `ctx`'s device into its rgb8 buffer (device pointers).
static int resolve_slab(rtw_ctx *ctx, const float *const *bufs, uint32_t n_bufs, uint32_t width, uint32_t height, uint32_t row_begin,
                        uint32_t row_end, uint32_t spp_total, uint8_t *d_rgb8, cudaStream_t st) {
    ResolveArgs a{};
    for (uint32_t i = 0; i < n_bufs; ++i) a.bufs[i] = reinterpret_cast<const float4 *>(bufs[i]);
    a.n_bufs = n_bufs; a.width = width; a.height = height;
    a.row_begin = row_begin; a.row_end = row_end;
    a.scale = 1.0f / (float)spp_total;
    a.rgb8 = d_rgb8;
    a.nan_counter = ctx->stat_counters.p + ST_NAN_PIXELS;
    CK(cudaMemsetAsync(a.nan_counter, 0, sizeof(unsigned long long), st));
    CK(launch_resolve(a, st));
    return 0;
}

// Multi-GPU frame in one process.  Sample split as SURVEY §8(e): context i traces its share of the sample indices of
// EVERY pixel into its own fp32 buffer.  The exchange is reduce-scatter shaped and fused with the resolve: GPU g owns
// the scanline slab g, its resolve kernel reads that slab from all N buffers — its own from HBM, N-1 over NVLink peer
// mappings — sums, quantises and writes its rows of the image, and copies them straight into the caller's host buffer.
// Every GPU therefore ingests (N-1)/N of ONE buffer (round 1 had GPU 0 ingest N-1 whole buffers), all links carry
// traffic at once, and the N device-to-host copies run in parallel.  Cross-device ordering is by events (each resolve
// waits for every trace), the host blocks once at the end.  stats: ms_wall = host wall clock of the whole call,
// ms_trace / ms_resolve = device 0's kernel times.
int rtw_cuda_render_multi(rtw_ctx *const *ctxs, uint32_t n_ctx, const rtw_camera *cam, const rtw_render_params *p,
                          uint8_t *rgb8_out) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return fail(nullptr, 1, "no contexts");
    rtw_ctx *ctx = ctxs[0];
    const auto t_begin = std::chrono::steady_clock::now();
    if (n_ctx > kMaxResolveBufs) return fail(ctx, 1, "at most %u contexts", kMaxResolveBufs);
    if (!cam || !p || !rgb8_out) return fail(ctx, 1, "null argument");
    if (p->variant == RTW_VARIANT_WAVEFRONT) return fail(ctx, 1, "render_multi runs the megakernel variants only");
    if (p->width == 0 || p->height == 0 || p->spp_end <= p->spp_begin) return fail(ctx, 1, "empty image or sample range");
    const size_t npx = (size_t)p->width * p->height;
    const uint32_t spp = p->spp_end - p->spp_begin;
    const uint32_t spp_total = p->spp_total ? p->spp_total : spp;
    const float *bufs[kMaxResolveBufs];
    for (uint32_t i = 0; i < n_ctx; ++i) {
        if (!ctxs[i]) return fail(ctx, 1, "context %u is null", i);
        for (uint32_t k = 0; k < i; ++k)
            if (ctxs[k]->device == ctxs[i]->device) return fail(ctx, 1, "contexts %u and %u share device %d", k, i, ctxs[i]->device);
    }
    // peer mappings: already there after rtw_cuda_create_multi; contexts created one by one get them here, once
    for (uint32_t i = 0; i < n_ctx && n_ctx > 1; ++i)
        if (int rc = enable_peers(ctxs[i], ctxs, n_ctx)) { if (ctxs[i] != ctx) ctx->err = ctxs[i]->err; return rc; }
    for (uint32_t i = 0; i < n_ctx; ++i) {
        rtw_ctx *c = ctxs[i];
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, 2, "cudaSetDevice(%d) failed", c->device);
        if (c->accum.n != npx) {
            if (c->accum.alloc(npx) != cudaSuccess || c->rgb8.alloc(npx * 3) != cudaSuccess) return fail(ctx, 2, "allocation failed on device %d", c->device);
        }
        if (cudaMemsetAsync(c->accum.p, 0, npx * sizeof(float4), c->stream) != cudaSuccess) return fail(ctx, 2, "memset failed");
        rtw_render_params sub = *p;
        const uint32_t base = spp / n_ctx, extra = spp % n_ctx;
        sub.spp_begin = p->spp_begin + i * base + std::min(i, extra);
        sub.spp_end = sub.spp_begin + base + (i < extra ? 1u : 0u);
        if (int rc = accumulate_impl(c, cam, &sub, reinterpret_cast<float *>(c->accum.p), c->stream, i == 0)) {
            if (c != ctx) ctx->err = c->err;
            return rc;
        }
        if (cudaEventRecord(c->ev_trace, c->stream) != cudaSuccess) return fail(ctx, 2, "event record failed on device %d", c->device);
        bufs[i] = reinterpret_cast<const float *>(c->accum.p);
    }
    // slab g = scanlines [g H / N, (g+1) H / N); image rows are flipped (main.zig:396), so slab g lands in the host
    // buffer at rows [H - j1, H - j0)
    for (uint32_t g = 0; g < n_ctx; ++g) {
        rtw_ctx *c = ctxs[g];
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, 2, "cudaSetDevice(%d) failed", c->device);
        for (uint32_t k = 0; k < n_ctx; ++k)
            if (k != g && cudaStreamWaitEvent(c->stream, ctxs[k]->ev_trace, 0) != cudaSuccess) return fail(ctx, 2, "cross-device event wait failed");
        const uint32_t j0 = (uint32_t)((uint64_t)p->height * g / n_ctx), j1 = (uint32_t)((uint64_t)p->height * (g + 1) / n_ctx);
        // own buffer first: the local HBM read overlaps the first peer reads
        const float *order[kMaxResolveBufs];
        for (uint32_t k = 0; k < n_ctx; ++k) order[k] = bufs[(g + k) % n_ctx];
        if (g == 0) CK(cudaEventRecord(c->ev[2], c->stream));
        if (int rc = resolve_slab(c, order, n_ctx, p->width, p->height, j0, j1, spp_total, c->rgb8.p, c->stream)) { if (c != ctx) ctx->err = c->err; return rc; }
        if (g == 0) CK(cudaEventRecord(c->ev[3], c->stream));
        const size_t off = (size_t)(p->height - j1) * p->width * 3, bytes = (size_t)(j1 - j0) * p->width * 3;
        if (bytes && cudaMemcpyAsync(rgb8_out + off, c->rgb8.p + off, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
            return fail(ctx, 2, "device-to-host copy failed on device %d", c->device);
    }
    unsigned long long nan_total = 0;
    for (uint32_t i = 0; i < n_ctx; ++i) {
        cudaSetDevice(ctxs[i]->device);
        if (cudaStreamSynchronize(ctxs[i]->stream) != cudaSuccess) return fail(ctx, 2, "device %d: kernel failed: %s", ctxs[i]->device, cudaGetErrorString(cudaGetLastError()));
        unsigned long long nn = 0;
        if (cudaMemcpy(&nn, ctxs[i]->stat_counters.p + ST_NAN_PIXELS, sizeof nn, cudaMemcpyDeviceToHost) == cudaSuccess) nan_total += nn;
    }
    CK(cudaSetDevice(ctx->device));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.ms_trace = ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
    ctx->stats.ms_resolve = ms;
    ctx->stats.n_launches = 2 * n_ctx;
    ctx->stats.nan_pixels = nan_total;
    ctx->stats.paths = (uint64_t)npx * spp;
    ctx->stats.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return 0;
}

// Tables of the reference-order probe (the reference's own object-space f64 fields and nested Translate/RotateY
// chains, hittable.zig:472-596): test infrastructure of the parity suite, so built and uploaded on first use only.
static int ensure_raw_scene(rtw_ctx *ctx) {
    if (ctx->raw.prims) return 0;
    const uint32_t n = ctx->n_prims;
    std::vector<RawPrim> raw(n);
    std::vector<RawXform> chains;
    for (uint32_t i = 0; i < n; ++i) {
        const rtw_prim &p = ctx->host_prims[i];
        RawPrim &r = raw[i];
        r.kind = p.kind; r.material = p.material;
        std::memcpy(r.v, p.v, sizeof r.v);
        r.chain_begin = (uint32_t)chains.size();
        r.chain_len = 0;
        if (p.xform < 0) continue;
        std::vector<RawXform> tmp;
        for (int x = p.xform; x >= 0; x = ctx->host_xforms[x].outer) {
            RawXform rx{};
            rx.kind = ctx->host_xforms[x].kind;
            rx.v[0] = ctx->host_xforms[x].v[0]; rx.v[1] = ctx->host_xforms[x].v[1]; rx.v[2] = ctx->host_xforms[x].v[2];
            tmp.push_back(rx);
        }
        std::reverse(tmp.begin(), tmp.end());  // outermost first
        r.chain_len = (uint32_t)tmp.size();
        chains.insert(chains.end(), tmp.begin(), tmp.end());
    }
    for (const auto &orig : ctx->host_inst_orig) std::memcpy(raw[orig.first].v, orig.second.data(), sizeof raw[orig.first].v);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->raw_prims.upload(raw));
    CK(ctx->raw_chains.upload(chains));
    ctx->raw = RawScene{ctx->raw_prims.p, ctx->raw_chains.p, ctx->nodes.p, ctx->bvh_prim_id.p, n, ctx->scene.root_is_leaf};
    return 0;
}

static int probe_impl(rtw_ctx *ctx, uint32_t n, const double *rays, const rtw_camera *cam, uint32_t width, uint32_t height,
                      uint32_t precision, uint32_t variant_req, uint32_t *prim_id, double *t, double *normal, double *uv) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!ctx->have_scene) return fail(ctx, 1, "no scene uploaded");
    if (!prim_id) return fail(ctx, 1, "prim_id is null");
    if (precision != 0 && precision != 32 && precision != 64) return fail(ctx, 1, "precision must be 0, 32 or 64");
    if (n == 0) return 0;
    int variant = 0;
    if (int rc = pick_variant(ctx, variant_req, &variant, precision == 0)) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<uint32_t> d_id;
    DevBuf<double> d_t, d_n, d_uv, d_rays;
    DevBuf<float> f_rays, f_t, f_n, f_uv;
    int rc = 0;
    auto cleanup = [&]() { d_id.release(); d_t.release(); d_n.release(); d_uv.release(); d_rays.release(); f_rays.release(); f_t.release(); f_n.release(); f_uv.release(); };
#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ctx, 2, "%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
    CKP(d_id.alloc(n));
    std::vector<double> host_rays;
    if (precision == 0 && !rays) {
        // production probe takes explicit rays: generate the parity-mode primary rays on the host in f64
        host_rays.resize((size_t)n * 7);
        for (uint32_t j = 0; j < height; ++j)
            for (uint32_t i = 0; i < width; ++i) {
                const double s = ((double)i + 0.5) / ((double)width - 1.0), tt = ((double)j + 0.5) / ((double)height - 1.0);
                double *q = &host_rays[((size_t)j * width + i) * 7];
                for (int a = 0; a < 3; ++a) {
                    q[a] = cam->origin[a];
                    q[3 + a] = cam->lower_left_corner[a] + cam->horizontal[a] * s + cam->vertical[a] * tt - cam->origin[a];
                }
                q[6] = cam->time0 + 0.5 * (cam->time1 - cam->time0);
            }
        rays = host_rays.data();
    }
    if (precision == 0) {
        std::vector<float> fr((size_t)n * 7);
        for (size_t k = 0; k < fr.size(); ++k) fr[k] = (float)rays[k];
        CKP(f_rays.upload(fr));
        CKP(f_t.alloc(n)); CKP(f_n.alloc((size_t)n * 3)); CKP(f_uv.alloc((size_t)n * 2));
        CKP(cudaDeviceSynchronize());  // the pageable H2D copies above must have landed
        CKP(launch_probe(variant, ctx->scene, n, f_rays.p, d_id.p, f_t.p, f_n.p, f_uv.p, st));
        CKP(cudaStreamSynchronize(st));
        std::vector<float> ht(n), hn((size_t)n * 3), hu((size_t)n * 2);
        CKP(cudaMemcpy(prim_id, d_id.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CKP(cudaMemcpy(ht.data(), f_t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
        CKP(cudaMemcpy(hn.data(), f_n.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        CKP(cudaMemcpy(hu.data(), f_uv.p, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost));
        if (t) for (uint32_t k = 0; k < n; ++k) t[k] = ht[k];
        if (normal) for (size_t k = 0; k < (size_t)n * 3; ++k) normal[k] = hn[k];
        if (uv) for (size_t k = 0; k < (size_t)n * 2; ++k) uv[k] = hu[k];
    } else {
        if ((rc = ensure_raw_scene(ctx))) { cleanup(); return rc; }
        RawCamera rc_cam{};
        if (!rays) {
            for (int a = 0; a < 3; ++a) {
                rc_cam.origin[a] = cam->origin[a]; rc_cam.horizontal[a] = cam->horizontal[a];
                rc_cam.vertical[a] = cam->vertical[a]; rc_cam.llc[a] = cam->lower_left_corner[a];
            }
            rc_cam.time0 = cam->time0; rc_cam.time1 = cam->time1;
        } else {
            std::vector<double> hr(rays, rays + (size_t)n * 7);
            CKP(d_rays.upload(hr));
        }
        CKP(d_t.alloc(n)); CKP(d_n.alloc((size_t)n * 3)); CKP(d_uv.alloc((size_t)n * 2));
        CKP(cudaDeviceSynchronize());  // the pageable H2D copies above must have landed
        CKP(launch_ref_probe((int)precision, variant, ctx->raw, n, rays ? d_rays.p : nullptr, rays ? nullptr : &rc_cam, width,
                             height, d_id.p, d_t.p, d_n.p, d_uv.p, st));
        CKP(cudaStreamSynchronize(st));
        CKP(cudaMemcpy(prim_id, d_id.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (t) CKP(cudaMemcpy(t, d_t.p, n * sizeof(double), cudaMemcpyDeviceToHost));
        if (normal) CKP(cudaMemcpy(normal, d_n.p, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        if (uv) CKP(cudaMemcpy(uv, d_uv.p, (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    }
#undef CKP
    cleanup();
    return 0;
}

int rtw_cuda_trace_rays(rtw_ctx *ctx, uint32_t n, const double *rays, uint32_t precision, uint32_t variant,
                        uint32_t *prim_id, double *t, double *normal, double *uv) {
    if (ctx && n && !rays) return fail(ctx, 1, "rays is null");
    return probe_impl(ctx, n, rays, nullptr, 0, 0, precision, variant, prim_id, t, normal, uv);
}

int rtw_cuda_primary_hits(rtw_ctx *ctx, const rtw_camera *cam, uint32_t width, uint32_t height, uint32_t precision,
                          uint32_t variant, uint32_t *prim_id, double *t, double *normal) {
    if (ctx && !cam) return fail(ctx, 1, "camera is null");
    if (ctx && ((uint64_t)width * height > 0x7FFFFFFFull)) return fail(ctx, 1, "image too large");
    return probe_impl(ctx, width * height, nullptr, cam, width, height, precision, variant, prim_id, t, normal, nullptr);
}

// ---- unit probes of the stochastic device code (rtw_unit.cu) -------------------------------------------------------
extern "C++" {
namespace {
template <class T>
int to_device(rtw_ctx *ctx, DevBuf<T> &d, const T *h, size_t n) {
    CK(d.alloc(n));
    CK(cudaMemcpy(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <class T>
int to_host(rtw_ctx *ctx, T *h, const DevBuf<T> &d, size_t n) {
    CK(cudaMemcpy(h, d.p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
}
}  // namespace
}  // extern "C++"

int rtw_cuda_unit_camera(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *p, uint32_t n, const uint32_t *ijs, float *out) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!cam || !p || (n && (!ijs || !out))) return fail(ctx, 1, "null argument");
    if (p->width == 0 || p->height == 0) return fail(ctx, 1, "empty image");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    DevBuf<uint32_t> d_in;
    DevBuf<float> d_out;
    int rc = to_device(ctx, d_in, ijs, (size_t)n * 3);
    if (!rc && d_out.alloc((size_t)n * 14) != cudaSuccess) rc = fail(ctx, 2, "allocation failed");
    if (!rc) {
        cudaDeviceSynchronize();
        const cudaError_t e = launch_unit_camera(lower_camera(cam), base_render(p), n, d_in.p, d_out.p, ctx->stream);
        if (e != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, 2, "unit camera kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) rc = to_host(ctx, out, d_out, (size_t)n * 14);
    d_in.release(); d_out.release();
    return rc;
}

int rtw_cuda_unit_samplers(rtw_ctx *ctx, uint32_t n, const float *u3, float *out) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (n && (!u3 || !out)) return fail(ctx, 1, "null argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    DevBuf<float> d_in, d_out;
    int rc = to_device(ctx, d_in, u3, (size_t)n * 3);
    if (!rc && d_out.alloc((size_t)n * 8) != cudaSuccess) rc = fail(ctx, 2, "allocation failed");
    if (!rc) {
        cudaDeviceSynchronize();
        const cudaError_t e = launch_unit_samplers(n, d_in.p, d_out.p, ctx->stream);
        if (e != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, 2, "unit samplers kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) rc = to_host(ctx, out, d_out, (size_t)n * 8);
    d_in.release(); d_out.release();
    return rc;
}

int rtw_cuda_unit_uniforms(rtw_ctx *ctx, const rtw_render_params *p, uint32_t n, const uint32_t *psb, float *out) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!p || (n && (!psb || !out))) return fail(ctx, 1, "null argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(ctx->device));
    DevBuf<uint32_t> d_in;
    DevBuf<float> d_out;
    int rc = to_device(ctx, d_in, psb, (size_t)n * 3);
    if (!rc && d_out.alloc((size_t)n * 8) != cudaSuccess) rc = fail(ctx, 2, "allocation failed");
    if (!rc) {
        cudaDeviceSynchronize();
        const cudaError_t e = launch_unit_uniforms(base_render(p), n, d_in.p, d_out.p, ctx->stream);
        if (e != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, 2, "unit uniforms kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) rc = to_host(ctx, out, d_out, (size_t)n * 8);
    d_in.release(); d_out.release();
    return rc;
}

int rtw_cuda_unit_shade(rtw_ctx *ctx, const rtw_render_params *p, uint32_t n, const double *rays, const uint32_t *psb,
                        uint32_t *prim_id, float *out) {
    if (!ctx) return fail(nullptr, 1, "ctx is null");
    if (!ctx->have_scene) return fail(ctx, 1, "no scene uploaded");
    if (!p || (n && (!rays || !psb || !prim_id || !out))) return fail(ctx, 1, "null argument");
    if (n == 0) return 0;
    int variant = 0;
    if (int rc = pick_variant(ctx, p->variant, &variant)) return rc;
    CK(cudaSetDevice(ctx->device));
    std::vector<float> fr((size_t)n * 7);
    for (size_t k = 0; k < fr.size(); ++k) fr[k] = (float)rays[k];
    DevBuf<float> d_rays, d_out;
    DevBuf<uint32_t> d_psb, d_id;
    int rc = to_device(ctx, d_rays, fr.data(), fr.size());
    if (!rc) rc = to_device(ctx, d_psb, psb, (size_t)n * 3);
    if (!rc && (d_out.alloc((size_t)n * 20) != cudaSuccess || d_id.alloc(n) != cudaSuccess)) rc = fail(ctx, 2, "allocation failed");
    if (!rc) {
        cudaDeviceSynchronize();
        const cudaError_t e = launch_unit_shade(variant, ctx->scene, base_render(p), n, d_rays.p, d_psb.p, d_id.p, d_out.p, ctx->stream);
        if (e != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, 2, "unit shade kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (!rc) rc = to_host(ctx, out, d_out, (size_t)n * 20);
    if (!rc) rc = to_host(ctx, prim_id, d_id, (size_t)n);
    d_rays.release(); d_out.release(); d_psb.release(); d_id.release();
    return rc;
}

int rtw_cuda_stats(rtw_ctx *ctx, rtw_stats *out) {
    if (!ctx || !out) return fail(ctx, 1, "null argument");
    *out = ctx->stats;
    return 0;
}

int rtw_cuda_measure_fp32_peak(rtw_ctx *ctx, double *tflops, double *sm_clock_mhz) {
    if (!ctx || !tflops) return fail(ctx, 1, "null argument");
    CK(cudaSetDevice(ctx->device));
    const int grid = ctx->n_sms * 8, iters = 4096;
    DevBuf<float> out;
    CK(out.alloc((size_t)grid * 256));
    cudaStream_t st = ctx->stream;
    CK(launch_ffma_peak(out.p, grid, 64, st));  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev[0], st));
        CK(launch_ffma_peak(out.p, grid, iters, st));
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        const double flops = (double)grid * 256 * iters * 16 * 8 * 2;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    out.release();
    *tflops = best;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *sm_clock_mhz = khz / 1000.0;
    }
    return 0;
}

}  // extern "C"
