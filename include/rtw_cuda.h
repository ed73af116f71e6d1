/*
 * rtw_cuda.h — C ABI of the B200-native path-tracing loop.
 *
 * This is the drop-in boundary for the per-pixel path-tracing loop of
 * nsfisis/RayTracingInOneWeekend.zig.  The reference has no FFI of its own: the loop is
 * inline in `main` (reference src/main.zig:382-402) and calls the private `Camera.getRay`
 * (src/main.zig:91-100) and `rayColor` (src/main.zig:103-122).  The seam is cut around that
 * loop: everything the loop reads (world, camera, background, image size, spp, depth)
 * crosses this header as plain-old-data; what the loop writes (the rgb24 pixel array,
 * src/main.zig:395-400) comes back.
 *
 * All scene/camera scalars are `double` because the reference's `Vec3` is 3 x f64
 * (src/rtw/vec.zig:8-11): the host-language flatten pass copies fields without rounding
 * and the library owns every precision decision (device arithmetic is fp32).
 *
 * Ownership (reference: host owns the Rc graph and frees it after render,
 * src/main.zig:363): the library COPIES everything passed to rtw_cuda_upload_scene; the
 * caller owns all output buffers.  Threading: one caller thread per context.
 * Errors: every entry point returns 0 on success, non-zero otherwise;
 * rtw_cuda_last_error() gives the text.  There is no CPU fallback: without a CUDA
 * device rtw_cuda_create fails.
 */
#ifndef RTW_CUDA_H
#define RTW_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTW_ABI_VERSION 3u

/* ---- primitives: the leaf variants of `Hittable` (src/rtw/hittable.zig:22-33) -------- */
enum {
    RTW_PRIM_SPHERE        = 0, /* v = {cx,cy,cz, r}                        hittable.zig:90-93   */
    RTW_PRIM_MOVING_SPHERE = 1, /* v = {c0x,c0y,c0z, c1x,c1y,c1z, t0,t1, r} hittable.zig:157-163 */
    RTW_PRIM_XY_RECT       = 2, /* v = {x0,x1, y0,y1, k}                    hittable.zig:270-276 */
    RTW_PRIM_XZ_RECT       = 3, /* v = {x0,x1, z0,z1, k}                    hittable.zig:323-329 */
    RTW_PRIM_YZ_RECT       = 4  /* v = {y0,y1, z0,z1, k}                    hittable.zig:376-382 */
};

/* One leaf of the hittable graph.  Array order = primitive id = depth-first append order
 * of the reference's nested lists (`HittableList`, `Box` sides hittable.zig:437-442), which
 * is what makes the reference's tie rule ("later list element wins", hittable.zig:235-242)
 * reproducible.  `Box` is flattened to its 6 rects; `list` nodes vanish. */
typedef struct rtw_prim {
    uint32_t kind;      /* RTW_PRIM_*                                                        */
    uint32_t material;  /* index into rtw_scene_desc.materials (identity of Rc(Material).ptr) */
    int32_t  xform;     /* innermost instance node wrapping this leaf, or -1                  */
    uint32_t reserved;
    double   v[10];
} rtw_prim;

/* ---- instance nodes: `Translate` / `RotateY` (hittable.zig:472-489, 505-596) --------- */
enum {
    RTW_XFORM_TRANSLATE = 0, /* v = {ox,oy,oz}       */
    RTW_XFORM_ROTATE_Y  = 1  /* v = {sin_t, cos_t}   as cached by RotateY.init hittable.zig:513-515 */
};
typedef struct rtw_xform {
    uint32_t kind;
    int32_t  outer;  /* the instance node that wraps this one, or -1 (chain goes outward) */
    double   v[4];
} rtw_xform;

/* ---- materials (src/rtw/material.zig:16-20) ------------------------------------------- */
enum {
    RTW_MAT_DIFFUSE       = 0, /* texture = albedo texture          material.zig:41-53  */
    RTW_MAT_METAL         = 1, /* albedo[3], param = fuzz           material.zig:55-66  */
    RTW_MAT_DIELECTRIC    = 2, /* param = index of refraction       material.zig:68-92  */
    RTW_MAT_DIFFUSE_LIGHT = 3  /* texture = emit texture            material.zig:94-110 */
};
typedef struct rtw_material {
    uint32_t kind;
    int32_t  texture;
    double   albedo[3];
    double   param;
} rtw_material;

/* ---- textures (src/rtw/texture.zig:10-14) ---------------------------------------------- */
enum {
    RTW_TEX_SOLID   = 0, /* color[3]                                   texture.zig:46-55   */
    RTW_TEX_CHECKER = 1, /* a = odd texture, b = even texture          texture.zig:57-83   */
    RTW_TEX_NOISE   = 2, /* a = index into perlins, scale              texture.zig:85-105  */
    RTW_TEX_IMAGE   = 3  /* a = index into images                      texture.zig:107-145 */
};
typedef struct rtw_texture {
    uint32_t kind;
    int32_t  a;
    int32_t  b;
    uint32_t reserved;
    double   color[3];
    double   scale;
} rtw_texture;

/* Decoded image, 4 bytes per texel RGBA, row 0 = top (texture.zig:131-137). */
typedef struct rtw_image {
    uint32_t       width;
    uint32_t       height;
    const uint8_t *rgba8;
} rtw_image;

/* Perlin tables (src/rtw/perlin.zig:10-16): 256 unit vectors + 3 permutations of 0..255. */
typedef struct rtw_perlin {
    const double   *ranvec; /* 256*3 */
    const uint32_t *perm_x; /* 256   */
    const uint32_t *perm_y;
    const uint32_t *perm_z;
} rtw_perlin;

typedef struct rtw_scene_desc {
    uint32_t            n_prims;
    const rtw_prim     *prims;
    uint32_t            n_xforms;
    const rtw_xform    *xforms;
    uint32_t            n_materials;
    const rtw_material *materials;
    uint32_t            n_textures;
    const rtw_texture  *textures;
    uint32_t            n_images;
    const rtw_image    *images;
    uint32_t            n_perlins;
    const rtw_perlin   *perlins;
    /* shutter interval the leaf boxes of moving spheres must cover
     * (MovingSphere.boudingBox(time0,time1), hittable.zig:203-217) */
    double              time0;
    double              time1;
} rtw_scene_desc;

/* The ten fields of the reference `Camera` (src/main.zig:40-51), produced by Camera.init
 * (src/main.zig:52-89), which stays on the host. */
typedef struct rtw_camera {
    double origin[3];
    double horizontal[3];
    double vertical[3];
    double lower_left_corner[3];
    double u[3];
    double v[3];
    double w[3];
    double lens_radius;
    double time0;
    double time1;
} rtw_camera;

/* Traversal / kernel selection. */
enum {
    RTW_VARIANT_AUTO      = 0, /* library picks per scene                                     */
    RTW_VARIANT_MEGA_FLAT = 1, /* persistent megakernel, warp-uniform scan of smem-staged prims */
    RTW_VARIANT_MEGA_BVH  = 2, /* persistent megakernel, stack-based BVH traversal            */
    RTW_VARIANT_WAVEFRONT = 3  /* per-bounce ray queues, compaction between bounces           */
};

/* Who built the BVH of the uploaded scene (rtw_stats.bvh_builder).  Scenes of >= 65536 primitives build on
 * the device; the environment variable RTW_BVH_BUILDER=sah|lbvh overrides the choice. */
enum {
    RTW_BVH_BUILDER_SAH  = 0, /* host: binned surface-area heuristic                          */
    RTW_BVH_BUILDER_LBVH = 1  /* device: Morton order + parallel radix tree                   */
};

typedef struct rtw_render_params {
    uint32_t width;
    uint32_t height;
    uint32_t spp_begin;   /* sample indices [spp_begin, spp_end) are rendered               */
    uint32_t spp_end;
    uint32_t spp_total;   /* divisor for the resolve (src/main.zig:395); 0 = spp_end-spp_begin */
    uint32_t max_depth;   /* src/main.zig:307 */
    uint32_t variant;     /* RTW_VARIANT_* */
    uint32_t flags;       /* RTW_FLAG_* */
    uint64_t seed;        /* Philox key; (seed, pixel, sample, bounce) fixes every draw      */
    double   background[3]; /* src/main.zig:318 */
} rtw_render_params;

enum {
    RTW_FLAG_COUNT_EVENTS = 1u, /* run the instrumented kernel build and fill rtw_stats counters */
    RTW_FLAG_DETERMINISTIC = 2u /* lane-owns-pixel kernel: fixed fp32 summation order (bit-reproducible
                                   frames) at some cost in lane utilisation; default is the pooled kernel
                                   whose per-path vector atomics commute only up to fp32 rounding */
};

/* Event counters of the last render / timing of its kernels.  Event classes are the rows
 * of SURVEY.md §8(d) "algorithmic flops per unit of work". */
typedef struct rtw_stats {
    uint64_t paths;            /* (pixel,sample) pairs = calls of rayColor from main.zig:393   */
    uint64_t rays;             /* calls of world.hit from main.zig:109 (all bounces)           */
    uint64_t node_tests;       /* BVH slab tests                                               */
    uint64_t sphere_tests;     /* sphere candidates tested                                     */
    uint64_t sphere_roots;     /* ... that passed the discriminant                             */
    uint64_t moving_tests;     /* moving-sphere centre evaluations                             */
    uint64_t rect_tests;
    uint64_t rect_accepts;
    uint64_t xform_apps;       /* instance transforms applied (in+out)                         */
    uint64_t sphere_finalise;  /* winning sphere hits shaded                                   */
    uint64_t scatter_diffuse;
    uint64_t scatter_metal;
    uint64_t scatter_dielectric;
    uint64_t emit_hits;
    uint64_t tex_checker;
    uint64_t tex_image;
    uint64_t tex_noise;
    uint64_t nan_pixels;       /* pixels whose sum was NaN at resolve      */
    double   ms_trace;         /* device time of the path-tracing kernel(s), CUDA events       */
    double   ms_resolve;
    double   ms_upload;        /* last rtw_cuda_upload_scene incl. BVH build                   */
    uint32_t n_launches;       /* kernels launched by the last render call                     */
    uint32_t variant_used;
    uint32_t bvh_nodes;
    uint32_t bvh_depth;
    double   ms_bvh_build;     /* BVH part of ms_upload                                         */
    uint32_t bvh_builder;      /* RTW_BVH_BUILDER_*                                             */
    uint32_t reserved0;
    double   ms_wall;          /* rtw_cuda_render / rtw_cuda_render_multi: host wall clock of the whole call     */
} rtw_stats;

typedef struct rtw_ctx rtw_ctx;

/* Create a context bound to CUDA device `device`.  Fails (non-zero) when no device. */
int rtw_cuda_create(int device, rtw_ctx **out);
void rtw_cuda_destroy(rtw_ctx *ctx);
const char *rtw_cuda_last_error(const rtw_ctx *ctx); /* ctx may be NULL: last create error */
uint32_t rtw_cuda_abi_version(void);

/* Contexts on devices 0..n_gpus-1 (out: array of n_gpus pointers) with NVLink peer mappings between every pair
 * enabled up front — the form rtw_cuda_render_multi wants.  (SURVEY §8b sketched `create(n_gpus)` returning ONE
 * context; one context per device is kept instead because `torchrun`-style hosts own exactly one device per process
 * and use rtw_cuda_create + rtw_cuda_accumulate + their own collective.)  All-or-nothing: on failure nothing is left. */
int rtw_cuda_create_multi(uint32_t n_gpus, rtw_ctx **out);

/* Tuning knobs (INTEGRATION.md lists them).  Each knob also has an environment variable of the same name, read ONCE
 * in rtw_cuda_create; this call overrides it on a live context, value NULL restores the built-in default.  Unknown
 * names are an error. */
int rtw_cuda_set_option(rtw_ctx *ctx, const char *name, const char *value);

/* Copy the scene, build leaf boxes (rules of the reference's `boudingBox` methods,
 * hittable.zig:133-143,203-217,305-316,358-369,411-422,491-498,598-603), build the BVH
 * (new: the reference scans a list, hittable.zig:231-244), upload SoA buffers. */
int rtw_cuda_upload_scene(rtw_ctx *ctx, const rtw_scene_desc *scene);

/* Replaces the whole loop nest src/main.zig:382-402: trace samples [spp_begin,spp_end) of
 * every pixel, average, sqrt, clamp, x256 -> u8, write row (H-1-j) (main.zig:396).
 * rgb8_out: width*height*3 bytes, HOST memory, row 0 = top.  accum_out (optional, HOST):
 * width*height*4 floats (r,g,b sums + sample count), row j = reference scanline j
 * (bottom row first, as the reference iterates). Blocking. */
int rtw_cuda_render(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *params,
                    uint8_t *rgb8_out, float *accum_out);

/* Device-pointer halves of rtw_cuda_render, for the multi-GPU spp split: each rank
 * accumulates its spp range into its own fp32 buffer, the buffers are summed (NCCL reduce),
 * rank 0 resolves.  d_accum: DEVICE, width*height*4 floats; it is ADDED to, not cleared.
 * `stream` is a cudaStream_t (NULL = default stream).  Asynchronous. */
int rtw_cuda_accumulate(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *params,
                        float *d_accum, void *stream);
/* Replaces src/main.zig:395-400.  d_rgb8: DEVICE, width*height*3 bytes. */
int rtw_cuda_resolve(rtw_ctx *ctx, const float *d_accum, uint32_t width, uint32_t height,
                     uint32_t spp_total, uint8_t *d_rgb8, void *stream);
/* Same, but sums n_bufs accumulation buffers (peer-mapped buffers of other GPUs allowed)
 * inside the resolve kernel: reduce + resolve fused over NVLink peer memory. */
int rtw_cuda_resolve_multi(rtw_ctx *ctx, const float *const *d_accums, uint32_t n_bufs,
                           uint32_t width, uint32_t height, uint32_t spp_total,
                           uint8_t *d_rgb8, void *stream);

/* Single-process multi-GPU form of rtw_cuda_render: `ctxs[i]` live on different devices and hold the SAME
 * uploaded scene.  Context i traces its share of [spp_begin, spp_end) (first spp mod n contexts take one extra
 * sample) into its own fp32 buffer; then EVERY context resolves one scanline slab of the image, its kernel reading
 * that slab from all n buffers through NVLink peer mappings (reduction fused into the resolve, reduce-scatter
 * shaped: each GPU ingests (n-1)/n of one buffer), and copies its rows into rgb8_out (HOST).  Needs peer access
 * between all pairs (rtw_cuda_create_multi sets it up; otherwise it is enabled on first use).  One thread calls
 * this for one set of contexts at a time.  Blocking; rtw_stats.ms_wall of ctxs[0] = wall clock of the call. */
int rtw_cuda_render_multi(rtw_ctx *const *ctxs, uint32_t n_ctx, const rtw_camera *cam,
                          const rtw_render_params *params, uint8_t *rgb8_out);

/* Parity probes: the closest-hit query `world.hit(r, 0.001, inf, &rec)` of
 * src/main.zig:109 for explicit rays.  precision = 32 (production arithmetic) or 64
 * (same device code instantiated in double: bit-comparable with the reference's f64).
 * Both run with FMA contraction off and IEEE div/sqrt.  rays: n x 7 doubles
 * {ox,oy,oz, dx,dy,dz, time}.  Outputs (HOST): prim_id[n] (0xFFFFFFFF = miss), t[n],
 * normal[3n] (face-corrected, as HitRecord.normal), uv[2n] (may be NULL). */
int rtw_cuda_trace_rays(rtw_ctx *ctx, uint32_t n, const double *rays, uint32_t precision,
                        uint32_t variant, uint32_t *prim_id, double *t, double *normal,
                        double *uv);
/* Primary rays in parity mode: pixel centre u=(i+.5)/(W-1), v=(j+.5)/(H-1), pinhole,
 * time=(time0+time1)/2.  Row j = reference scanline j (bottom first). */
int rtw_cuda_primary_hits(rtw_ctx *ctx, const rtw_camera *cam, uint32_t width, uint32_t height,
                          uint32_t precision, uint32_t variant, uint32_t *prim_id, double *t,
                          double *normal);

/* Unit-level probes of the production STOCHASTIC device code (parity instruments).  They run the same device
 * functions as the render kernels and return the random choices along with the result, so a host can replay the
 * reference's statements (f64) with exactly those choices.  All buffers HOST; `params` supplies width, height (the
 * (W-1),(H-1) denominators of main.zig:390-391) and the seed.
 *
 * rtw_cuda_unit_camera — Camera.getRay main.zig:91-100 + the jitter of :390-391 for n (i, j, sample) triples
 *   (ijs: 3n uint32).  out: 14n floats = ray {o, d, time} | uniforms {ju, jv, l1, l2, tm} | lens-disk point {x, y}.
 * rtw_cuda_unit_samplers — the rejection-free maps that replace rand.zig:22-40 for caller-supplied uniforms
 *   (u3: 3n floats in [0,1)).  out: 8n floats = unit vector | ball point | disk point.
 * rtw_cuda_unit_uniforms — the Philox4x32-10 draws of (pixel, sample, block) (psb: 3n uint32).  out: 8n floats =
 *   the four 24-bit uniforms | the four raw words (bit patterns).
 * rtw_cuda_unit_shade — one level of rayColor (main.zig:109-121) for explicit rays (7n doubles): production closest hit,
 *   then Material.emitted/scatter (material.zig:16-110) keyed by psb (pixel, sample, bounce).  prim_id[n] (RTW_MISS = no
 *   hit, rest of the row zero); out: 20n floats = t | scattered ray {o, d, time} | attenuation | emitted | continues
 *   (0/1) | sample vector consumed (diffuse: unit vector, metal: ball point) | uniform consumed (dielectric) | material kind. */
int rtw_cuda_unit_camera(rtw_ctx *ctx, const rtw_camera *cam, const rtw_render_params *params, uint32_t n,
                         const uint32_t *ijs, float *out);
int rtw_cuda_unit_samplers(rtw_ctx *ctx, uint32_t n, const float *u3, float *out);
int rtw_cuda_unit_uniforms(rtw_ctx *ctx, const rtw_render_params *params, uint32_t n, const uint32_t *psb, float *out);
int rtw_cuda_unit_shade(rtw_ctx *ctx, const rtw_render_params *params, uint32_t n, const double *rays,
                        const uint32_t *psb, uint32_t *prim_id, float *out);

int rtw_cuda_stats(rtw_ctx *ctx, rtw_stats *out);

/* Measured FP32 peak of this device: dependent-FFMA-chain microbenchmark, TFLOP/s
 * (2 flops per FFMA).  The roofline denominator bench.py reports against. */
int rtw_cuda_measure_fp32_peak(rtw_ctx *ctx, double *tflops, double *sm_clock_mhz);

#define RTW_MISS 0xFFFFFFFFu

#ifdef __cplusplus
}
#endif
#endif /* RTW_CUDA_H */
