// rtw_device.cuh — device-side data layout of the flattened scene (all fp32, SoA-of-records).
//
// The reference keeps a tagged-union graph of f64 structs (src/rtw/hittable.zig:22-33,
// material.zig:16-20, texture.zig:10-14).  On the device every table is an array of 32-byte
// records so one record = one or two 16-byte vector loads (coalesced / broadcast friendly):
//   DevPrim      2 x float4   primitives, either in reference order (flat scan) or BVH leaf order
//   BvhNode      2 x float4   {min, a | max, b}; siblings adjacent -> one aligned 64-byte fetch
//   DevXform     composed instance transform (Translate/RotateY chain, hittable.zig:472-596)
//   DevBigSphere reference-point data for spheres whose radius makes |oc|^2 - r^2 cancel in fp32
//   DevMaterial / DevTexture
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtw {

constexpr uint32_t kMiss = 0xFFFFFFFFu;

// Scene-feature mask the flat megakernel is specialised on (k_megakernel_flat<STATS, MINB, FEAT>): code for primitive kinds
// and texture kinds the uploaded scene does not contain is not compiled into the kernel the scene runs (a 47 KB kernel body
// against a ~6 KB L0 / 32 KB L1.5 instruction cache: ncu r02_a shows 1.6 warps per issue waiting for instructions).
enum : uint32_t { FF_SPHERES = 1u, FF_RECTS = 2u, FF_TEX = 4u, FF_ALL = 7u };  // FF_TEX = noise / image textures, instanced-sphere uv


// prim.b.w (as uint) = kind | (extra << 8)
enum : uint32_t { PK_SPHERE = 0, PK_XY = 2, PK_XZ = 3, PK_YZ = 4 };

struct __align__(16) DevPrim {
    // sphere: a = (cbx, cby, cbz, r)   centre(time) = cb + vel*time   (moving: hittable.zig:219-221)
    //         b = (vx, vy, vz, meta)   meta = PK_SPHERE | (big_index+1) << 8
    // rect:   a = (a0, a1, b0, b1)     in-plane bounds, widened by `slack` on every side (watertight seams)
    //         b = (k, bits(xform index or -1), slack, meta)
    float4 a, b;
};

struct __align__(16) BvhNode {
    float mnx, mny, mnz;
    uint32_t a;  // interior: index of left child (right = a+1); leaf: first slot | count << kNodeRefShift
    float mxx, mxy, mxz;
    uint32_t b;  // interior: 0; leaf: primitive count (1..15)
};
constexpr uint32_t kNodeRefShift = 28, kNodeRefIndexMask = (1u << kNodeRefShift) - 1u;


struct __align__(16) DevXform {
    // object = A*world + t with A = rotation about Y: x' = c*x - s*z, z' = s*x + c*z
    float c, s, tx, ty, tz, pad0, pad1, pad2;
};

struct __align__(16) DevBigSphere {
    // o - centre = (o - q) + m, |o-centre|^2 - r^2 = |o-q|^2 + 2 (o-q).m + K, K = |m|^2 - r^2
    float qx, qy, qz, K;
    float mx, my, mz, pad;
};

struct __align__(16) DevMaterial {
    uint32_t kind;  // RTW_MAT_*
    int32_t tex;
    float r, g, b;  // metal albedo
    float param;    // fuzz | ir
    float pad0, pad1;
};

struct __align__(16) DevTexture {
    uint32_t kind;  // RTW_TEX_*
    int32_t a, b;   // checker: odd, even; noise: perlin index; image: image index
    float scale;
    float r, g, bl;
    uint32_t pad;
};

struct DevImage {
    cudaTextureObject_t tex;  // uchar4, point filter, unnormalised coords, clamp
    uint32_t w, h;
    uint32_t pad0, pad1;
};

struct DevPerlin {
    float4 ranvec[256];
    uint8_t perm[3][256];
};

// Shared-memory image of a small scene for the flat scan, as one blob of float4 (offsets in float4
// units).  Primitives are SEGMENTED BY KIND and small spheres — static and moving together, by position — are packed into
// spatial GROUPS of four behind a bounding sphere, so that a whole warp can skip a group:
//   big        : 1 x float4  (cx, cy, cz, r^2) per individually tested static sphere (|r| >= 64, or one of the few that dwarf
//                the rest of a small scene); DevBigSphere i in scene.bigs
//   groups     : 8 x float4  = 4 members x {(cbx, cby, cbz, r^2), (vx, vy, vz, r)}: centre(time) = cb + vel * time, vel = 0
//                for a static sphere (FlatLayout::n_mov_groups / off_mov; the n_sph_groups / off_sph slots are unused: 0)
//   bounds     : 1 x float4 per group: (cx, cy, cz, R), the smallest ball enclosing the members over the shutter interval;
//                padded to a multiple of four with NaN radii (never pass): the scan tests four bounds per warp reduction
//   rect       : 2 x float4 per rect: (a0, a1, b0, b1), (k, bits(prim id), -, -), sorted into RUNS of equal
//                (instance transform, orientation) so the object-space ray is set up once per run and the
//                inner loop is specialised per orientation and branch-free
//   runs       : uint4 per run: (xform slot + 1 or 0 | kRunSameXform when the previous run has the same
//                transform, PK_XY/XZ/YZ, first rect, count); the runs of one transform are adjacent
//   boxes      : kBoxF4 x float4 per box: rects of one instance that are faces of ONE axis-aligned box of its object space —
//                the reference's `Box` (6 faces, hittable.zig:429-470) and rooms like the Cornell box's five walls (faces
//                missing is fine: id -1).  (x0, x1, y0, y1), (z0, z1, bits(xform slot + 1 or 0), bits(face mask)), four prim
//                ids, (two prim ids, cos, sin), (tx, ty, tz, -): ids in Box.init order z1, z0, y1, y0, x1, x0
//                (hittable.zig:437-442), the instance's composed transform inline.  Tested as three slabs (entry and exit
//                face), see closest_hit_flat in rtw_trace.cuh; rects that belong to a box are not in the runs.
//   ids        : uint32 prim id per slot, in the order big (padded to x4) | four per group
// Unused member slots have r^2 = -1 (never hit).
constexpr uint32_t kRunSameXform = 0x80000000u;
constexpr uint32_t kBoxF4 = 5;
constexpr uint32_t kFlatNoBounds = 1u;  // fewer than three sphere groups: the scan tests their members without the bounds
struct FlatLayout {
    uint32_t n_sph_groups, n_big, n_mov_groups, n_rect;
    uint32_t off_sph, off_big, off_mov, off_rect, off_ids;
    uint32_t total_f4;
    uint32_t flags;          // kFlatNoBounds
    uint32_t n_runs, off_runs, n_boxes, off_boxes, off_bounds;
};

struct DevScene {
    const float4 *flat_blob;      // FlatLayout image, staged into shared memory by the flat kernels
    FlatLayout flat;
    const DevPrim *prims_flat;    // reference order: slot == prim id
    const DevPrim *prims_bvh;     // BVH leaf order
    const uint32_t *bvh_prim_id;  // slot -> prim id
    const uint32_t *prim_material;  // indexed by prim id
    const BvhNode *nodes;
    const DevXform *xforms;
    const DevBigSphere *bigs;
    const DevMaterial *materials;
    const DevTexture *textures;
    const DevImage *images;
    const DevPerlin *perlins;
    uint32_t n_prims, n_nodes, n_xforms, n_materials, n_textures;
    uint32_t root_is_leaf;
};

struct DevCamera {  // src/main.zig:40-51, rounded to fp32
    float ox, oy, oz;
    float hx, hy, hz;
    float vx, vy, vz;
    float lx, ly, lz;
    float ux, uy, uz;
    float wx, wy, wz;  // camera v
    float lens_radius, time0, time1;
};

// Event counters.  Order must match rtw_stats in include/rtw_cuda.h.
enum StatSlot {
    ST_PATHS = 0, ST_RAYS, ST_NODE_TESTS, ST_SPHERE_TESTS, ST_SPHERE_ROOTS, ST_MOVING_TESTS, ST_RECT_TESTS,
    ST_RECT_ACCEPTS, ST_XFORM_APPS, ST_SPHERE_FINAL, ST_SC_DIFFUSE, ST_SC_METAL, ST_SC_DIELECTRIC, ST_EMIT,
    ST_TEX_CHECKER, ST_TEX_IMAGE, ST_TEX_NOISE, ST_NAN_PIXELS, ST_COUNT
};

struct DevRender {
    uint32_t width, height;
    uint32_t spp_begin, spp_end;
    uint32_t max_depth;
    uint32_t seed_lo, seed_hi;
    uint32_t philox_keys[20];      // Philox4x32-10 round keys: (seed_lo + i*0x9E3779B9, seed_hi + i*0xBB67AE85), i = 0..9
    float bg_r, bg_g, bg_b;
    float inv_wm1, inv_hm1;        // 1/(W-1), 1/(H-1)  (main.zig:390-391)
    float4 *accum;                 // width*height, row j = reference scanline j
    unsigned int *tile_counter;    // persistent-kernel work queue (tiles x chunks, or path batches)
    uint32_t n_batches, n_sblocks, batch_spp;  // pooled kernel: batch = one tile x batch_spp samples
    uint32_t service_threshold, steps_per_round, leaf_threshold;  // BVH state machine tuning
    unsigned long long *stats;     // ST_COUNT counters (instrumented build only)
    uint32_t n_tiles, tiles_x, spp_chunk, n_chunks;
};

}  // namespace rtw
