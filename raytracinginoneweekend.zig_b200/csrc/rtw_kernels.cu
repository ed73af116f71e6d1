// rtw_kernels.cu — production kernels (fp32, FMA contraction on), sm_100a.
//
//   k_megakernel_flat<STATS,MINB,FEAT> K1 for scenes of <= 256 primitives (default): flat scan out of shared memory, second
//                                    schedule (one Philox block + one evaluation of the sampling maps per iteration for the
//                                    whole warp), specialised on the scene's features (FF_* in rtw_device.cuh)
//   k_megakernel_bvh<STATS,FEAT,SPEC> K1 for larger scenes: BVH traversal as a per-lane state machine (SPEC: with one postponed
//                                    leaf per lane, RTW_BVH_KERNEL=3, measured slower)
//   k_megakernel_bvhq<STATS,FEAT,SLOTS,MINB> the same paths with a per-warp ray queue in shared memory (RTW_BVH_KERNEL=2; measured
//                                    slower than the state machine, profiles/r02_l: kept for A/B measurements)
//   k_megakernel_pooled<VARIANT,STATS> first flat schedule (RTW_FLAT_KERNEL=1, kept for A/B measurements)
//   k_megakernel<VARIANT,STATS>      K1, deterministic form (lane owns a pixel; RTW_FLAG_DETERMINISTIC)
//        all of these replace the loop nest src/main.zig:382-394 and everything below it
//   k_resolve / k_resolve4           K4: replaces src/main.zig:395-400 (average, sqrt, clamp, x256 -> u8, row flip),
//                                    optionally summing several (peer-mapped) accumulation buffers first
//   k_probe<VARIANT>                 production-arithmetic closest-hit probe (parity instrument)
//   k_ffma_peak                      FP32 roofline denominator, measured on the device
//
// Megakernel design.  The work is embarrassingly parallel and FP32/issue bound, with two sources of SIMT
// inefficiency: paths of very different length (1..50 rays) and per-ray divergence in traversal and material
// code.  All kernels are persistent: a grid of (#SMs x resident CTAs) CTAs pulls batches (an 8x4 pixel tile x a block
// of samples) from an atomic queue.  In the pooled kernels a warp owns the batch as a pool of path indices: whenever
// lanes finish their paths they are handed the next indices (ballot + popc prefix), so no lane waits for a longer
// neighbour, and a finished path is added to the frame with one vector reduction.  DESIGN.md §3 has the ncu history.
#include <cuda_runtime.h>

#include "rtw_kernels.h"
#include "rtw_trace.cuh"

namespace rtw {

constexpr int kBlock = 128;  // threads per CTA

// stage the FlatLayout blob into shared memory (once per CTA)
__device__ __forceinline__ void stage_flat(const DevScene &sc, float4 *s_flat) {
    for (uint32_t i = threadIdx.x; i < sc.flat.total_f4; i += blockDim.x) s_flat[i] = sc.flat_blob[i];
    __syncthreads();
}

// One ray of one path: closest hit, then miss / shade.  Returns false when the path ended (its
// radiance is then complete in L).  Called by ALL lanes of the warp (the flat scan votes);
// lanes without a path pass active = false and get false back.
template <int VARIANT, bool STATS>
__device__ __forceinline__ bool path_step(const DevScene &sc, const DevRender &rp, const float4 *s_flat, bool active,
                                          Ray &r, uint32_t pixel, uint32_t sample, uint32_t &bounce, float3 &beta,
                                          float3 &L, Counters<STATS> &cn) {
    Hit h{0.0f, kMiss};
    if (VARIANT == VAR_FLAT) h = closest_hit_flat<STATS>(r, active, s_flat, sc.flat, sc, 0.001f, cn);
    else if (active) h = closest_hit_bvh<STATS>(r, sc, 0.001f, cn);
    if (!active) return false;
    cn.add(ST_RAYS);
    if (h.slot == kMiss) {  // main.zig:109-112
        L.x = fmaf(beta.x, rp.bg_r, L.x); L.y = fmaf(beta.y, rp.bg_g, L.y); L.z = fmaf(beta.z, rp.bg_b, L.z);
        return false;
    }
    DevPrim prim;
    uint32_t prim_id;
    if (VARIANT == VAR_FLAT) { prim_id = h.slot; prim = sc.prims_flat[prim_id]; }
    else { prim = sc.prims_bvh[h.slot]; prim_id = sc.bvh_prim_id[h.slot]; }
    ++bounce;
    const bool go = shade<STATS>(sc, rp, r, prim, prim_id, h.t, pixel, sample, bounce, beta, L, cn);
    // depth exhausted: the next rayColor call returns black before intersecting (main.zig:105-108)
    return go && bounce < rp.max_depth;
}

// ---------------------------------------------------------------------------------------------
// Deterministic variant: lane = pixel, every pixel's samples summed in index order by one lane.
// ---------------------------------------------------------------------------------------------
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(kBlock) k_megakernel(const DevScene sc, const DevCamera cam, const DevRender rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    if (VARIANT == VAR_FLAT) stage_flat(sc, s_flat);
    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t total_items = rp.n_tiles * rp.n_chunks;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(rp.tile_counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total_items) break;
        // chunk-major order: all tiles of chunk 0 first, so concurrent warps rarely share a pixel
        const uint32_t chunk = item / rp.n_tiles, tile = item - chunk * rp.n_tiles;
        const uint32_t ty = tile / rp.tiles_x, tx = tile - ty * rp.tiles_x;
        const uint32_t i = tx * 8 + (lane & 7), j = ty * 4 + (lane >> 3);
        const bool inside = i < rp.width && j < rp.height;
        const uint32_t pixel = j * rp.width + i;
        uint32_t sample = rp.spp_begin + chunk * rp.spp_chunk;
        const uint32_t sample_end = inside ? min(sample + rp.spp_chunk, rp.spp_end) : sample;

        float3 sum = make_float3(0.0f, 0.0f, 0.0f);
        float3 beta = make_float3(1.0f, 1.0f, 1.0f), L = make_float3(0.0f, 0.0f, 0.0f);
        Ray r;
        uint32_t bounce = 0, cur_sample = 0;
        bool alive = false;
        for (;;) {
            if (!alive && sample < sample_end) {
                cur_sample = sample++;
                r = camera_ray(cam, rp, pixel, i, j, cur_sample);
                beta = make_float3(1.0f, 1.0f, 1.0f);
                L = make_float3(0.0f, 0.0f, 0.0f);
                bounce = 0;
                alive = true;
                cn.add(ST_PATHS);
            }
            if (!__any_sync(0xffffffffu, alive)) break;
            const bool was = alive;
            alive = path_step<VARIANT, STATS>(sc, rp, s_flat, was, r, pixel, cur_sample, bounce, beta, L, cn);
            if (was && !alive) { sum.x += L.x; sum.y += L.y; sum.z += L.z; }
        }
        if (inside) {
            float4 *dst = rp.accum + pixel;
            const float ns = (float)(sample_end - (rp.spp_begin + chunk * rp.spp_chunk));
            if (rp.n_chunks == 1) {  // this warp owns the pixel: plain read-modify-write, deterministic
                float4 a = *dst;
                a.x += sum.x; a.y += sum.y; a.z += sum.z; a.w += ns;
                *dst = a;
            } else {
                atomicAdd(&dst->x, sum.x); atomicAdd(&dst->y, sum.y); atomicAdd(&dst->z, sum.z); atomicAdd(&dst->w, ns);
            }
        }
    }
    if (STATS) cn.flush(rp.stats);
}

// ---------------------------------------------------------------------------------------------
// Pooled variant (default): the warp owns a pool of path indices (a batch = one 8x4 pixel tile x
// batch_spp samples, pulled from a global atomic queue).  Whenever lanes finish their paths they are
// handed the next indices of the pool — ballot + popc prefix, no lane ever waits for its neighbours —
// and a finished path is added to the frame with ONE vector reduction (red.global.add.v4.f32: r, g, b
// and the sample count).  There is no per-tile tail; only the last batches of the launch run under-full.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float4 *addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(kBlock, 9) k_megakernel_pooled(const DevScene sc, const DevCamera cam, const DevRender rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    if (VARIANT == VAR_FLAT) stage_flat(sc, s_flat);
    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // warp-uniform pool state
    uint32_t pool_next = 0, pool_end = 0, pool_tx = 0, pool_ty = 0, pool_s0 = 0;
    bool more = true;

    float3 beta = make_float3(1.0f, 1.0f, 1.0f), L = make_float3(0.0f, 0.0f, 0.0f);
    Ray r;
    uint32_t bounce = 0, cur_sample = 0, pixel = 0;
    bool alive = false;
    for (;;) {
        // ---- hand new paths to idle lanes ----
        for (;;) {
            const uint32_t need = __ballot_sync(0xffffffffu, !alive);
            if (!need) break;
            if (pool_next == pool_end) {
                if (!more) break;
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(rp.tile_counter, 1u);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= rp.n_batches) { more = false; break; }
                // sample-block-major: neighbouring warps work on neighbouring tiles
                const uint32_t sb = b / rp.n_tiles, tile = b - sb * rp.n_tiles;
                pool_ty = tile / rp.tiles_x; pool_tx = tile - pool_ty * rp.tiles_x;
                pool_s0 = rp.spp_begin + sb * rp.batch_spp;
                pool_next = 0;
                pool_end = 32u * min(rp.batch_spp, rp.spp_end - pool_s0);
            }
            const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
            const uint32_t rank = __popc(need & lt_mask);
            if (!alive && rank < take) {
                const uint32_t idx = pool_next + rank;
                const uint32_t pl = idx & 31u;
                const uint32_t i = pool_tx * 8 + (pl & 7), j = pool_ty * 4 + (pl >> 3);
                if (i < rp.width && j < rp.height) {  // ragged edge tiles: the slot is consumed, no path starts
                    pixel = j * rp.width + i;
                    cur_sample = pool_s0 + (idx >> 5);
                    r = camera_ray(cam, rp, pixel, i, j, cur_sample);
                    beta = make_float3(1.0f, 1.0f, 1.0f);
                    L = make_float3(0.0f, 0.0f, 0.0f);
                    bounce = 0;
                    alive = true;
                    cn.add(ST_PATHS);
                }
            }
            pool_next += take;
        }
        if (!__any_sync(0xffffffffu, alive)) {
            if (!more && pool_next == pool_end) break;
            continue;
        }
        const bool was = alive;
        alive = path_step<VARIANT, STATS>(sc, rp, s_flat, was, r, pixel, cur_sample, bounce, beta, L, cn);
        if (was && !alive) red_add_v4(rp.accum + pixel, L.x, L.y, L.z, 1.0f);
    }
    if (STATS) cn.flush(rp.stats);
}

// ---------------------------------------------------------------------------------------------
// Flat megakernel, second schedule (default for flat scenes): same pool, same scan, but the iteration is rotated so
// that the two divergent halves of the old one — "shade the lanes that hit" and "regenerate the lanes that ended",
// which between them cover the whole warp — share everything that is common to them:
//
//     scan                       all lanes                 closest hit of the 32 rays
//     A  hit lanes: hit record, emitted light, attenuation, metal absorb test, dielectric candidates (shade_prepare)
//        miss lanes: background;  ended lanes: one vector reduction into the frame
//     B  ended lanes take the next path indices of the pool (ballot + popc)
//     C  all lanes               ONE Philox block: counter (pixel, sample, bounce) — bounce 0 = a new camera ray —
//                                and ONE evaluation of the shared sampling maps (make_draw)
//     D  new lanes: camera ray;  scattering lanes: add the drawn part (scatter_finish)
//
// ncu on the first schedule (profiles/r01_l): regenerate 146 instructions at 14.6 lanes + shade/Philox ~270 at 6-19
// lanes per scan of ~800; here Philox and the sin/cos/sqrt of the sampling run once per iteration at 32 lanes.
// Samples are identical to every other kernel's (same counters, same maps): only the schedule differs.
// ---------------------------------------------------------------------------------------------
template <bool STATS, int MINB, uint32_t FEAT>
__global__ void __launch_bounds__(kBlock, MINB) k_megakernel_flat(const DevScene sc, const DevCamera cam, const DevRender rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    stage_flat(sc, s_flat);
    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t pool_next = 0, pool_end = 0, pool_tx = 0, pool_ty = 0, pool_s0 = 0;
    bool more = true;

    float3 beta = make_float3(1.0f, 1.0f, 1.0f), L = make_float3(0.0f, 0.0f, 0.0f);
    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    Pending pd{0.f, 0.f, 0.f, 0.f, 0u};
    uint32_t bounce = 0, cur_sample = 0, pixel = 0, pij = 0;  // pij = i | j << 16
    bool alive = false;
    for (;;) {
        // ---- B: hand new paths to lanes without one ----
        bool fresh = false;
        for (;;) {
            const uint32_t need = __ballot_sync(0xffffffffu, !alive && !fresh);
            if (!need) break;
            if (pool_next == pool_end) {
                if (!more) break;
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(rp.tile_counter, 1u);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= rp.n_batches) { more = false; break; }
                const uint32_t sb = b / rp.n_tiles, tile = b - sb * rp.n_tiles;
                pool_ty = tile / rp.tiles_x; pool_tx = tile - pool_ty * rp.tiles_x;
                pool_s0 = rp.spp_begin + sb * rp.batch_spp;
                pool_next = 0;
                pool_end = 32u * min(rp.batch_spp, rp.spp_end - pool_s0);
            }
            const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
            const uint32_t rank = __popc(need & lt_mask);
            if (!alive && !fresh && rank < take) {
                const uint32_t idx = pool_next + rank;
                const uint32_t pl = idx & 31u;
                const uint32_t i = pool_tx * 8 + (pl & 7), j = pool_ty * 4 + (pl >> 3);
                if (i < rp.width && j < rp.height) {  // ragged edge tiles: the slot is consumed, no path starts
                    pixel = j * rp.width + i;
                    pij = i | (j << 16);
                    cur_sample = pool_s0 + (idx >> 5);
                    bounce = 0;
                    fresh = true;
                    cn.add(ST_PATHS);
                }
            }
            pool_next += take;
        }
        if (!__any_sync(0xffffffffu, alive || fresh)) {
            if (!more && pool_next == pool_end) break;
            continue;
        }
        // ---- C: one Philox block and one evaluation of the sampling maps for the whole warp ----
        const uint4 rn = philox4x32_10(make_uint4(pixel, cur_sample, bounce, 0u), rp.philox_keys);
        const Draw dw = make_draw(rn, fresh);
        // ---- D: finish the rays ----
        if (fresh) {
            r = camera_from_draw(cam, rp, pij & 0xFFFFu, pij >> 16, dw);
            beta = make_float3(1.0f, 1.0f, 1.0f);
            L = make_float3(0.0f, 0.0f, 0.0f);
            alive = true;
        } else if (alive) {
            scatter_finish(r, pd, dw);
        }
        // ---- scan ----
        const Hit h = closest_hit_flat<STATS, FEAT>(r, alive, s_flat, sc.flat, sc, 0.001f, cn);
        // ---- A: what the hit decides without randomness ----
        if (alive) {
            cn.add(ST_RAYS);
            if (h.slot == kMiss) {  // main.zig:109-112
                L.x = fmaf(beta.x, rp.bg_r, L.x); L.y = fmaf(beta.y, rp.bg_g, L.y); L.z = fmaf(beta.z, rp.bg_b, L.z);
                alive = false;
            } else {
                const DevPrim prim = sc.prims_flat[h.slot];
                ++bounce;
                // depth exhausted: the next rayColor call returns black before intersecting (main.zig:105-108)
                alive = shade_prepare<STATS, FEAT>(sc, r, prim, h.slot, h.t, beta, L, pd, cn) && bounce < rp.max_depth;
            }
            if (!alive) red_add_v4(rp.accum + pixel, L.x, L.y, L.z, 1.0f);
        }
    }
    if (STATS) cn.flush(rp.stats);
}

// ---------------------------------------------------------------------------------------------
// BVH megakernel (pooled): a per-lane state machine.  BVH traversals have very different lengths
// (measured: 11.6 of 32 lanes active when every lane runs its traversal to completion before the warp
// shades), so here the warp interleaves: lanes that are traversing take one traversal step per
// iteration; lanes whose traversal is finished wait until at least rp.service_threshold of them have
// gathered (or nobody is traversing), then they are shaded, dead paths are replaced from the pool and
// new traversals start — all in one "service" phase that therefore runs reasonably full.
// ---------------------------------------------------------------------------------------------

template <bool STATS, uint32_t FEAT, bool SPEC = false>
__global__ void __launch_bounds__(kBlock) k_megakernel_bvh(const DevScene sc, const DevCamera cam, const DevRender rp) {
    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t pool_next = 0, pool_end = 0, pool_tx = 0, pool_ty = 0, pool_s0 = 0;
    bool more = true;

    float3 beta = make_float3(1.0f, 1.0f, 1.0f), L = make_float3(0.0f, 0.0f, 0.0f);
    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    uint32_t bounce = 0, cur_sample = 0, pixel = 0;
    bool alive = false, trav = false;
    BvhTraversal tv;
    BvhStack stack;
    tv.sp = 0; tv.cur = 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !trav);
        if (__popc(idle) >= (int)rp.service_threshold) {
            // ---- service phase: shade finished traversals ----
            if (!trav && alive) {
                cn.add(ST_RAYS);
                if (tv.h.slot == kMiss) {  // main.zig:109-112
                    L.x = fmaf(beta.x, rp.bg_r, L.x); L.y = fmaf(beta.y, rp.bg_g, L.y); L.z = fmaf(beta.z, rp.bg_b, L.z);
                    alive = false;
                } else {
                    const DevPrim prim = sc.prims_bvh[tv.h.slot];
                    const uint32_t prim_id = sc.bvh_prim_id[tv.h.slot];
                    ++bounce;
                    alive = shade<STATS, FEAT>(sc, rp, r, prim, prim_id, tv.h.t, pixel, cur_sample, bounce, beta, L, cn) &&
                            bounce < rp.max_depth;  // main.zig:105-108
                }
                if (!alive) red_add_v4(rp.accum + pixel, L.x, L.y, L.z, 1.0f);
            }
            // ---- hand new paths to lanes without one ----
            for (;;) {
                const uint32_t need = __ballot_sync(0xffffffffu, !alive);
                if (!need) break;
                if (pool_next == pool_end) {
                    if (!more) break;
                    uint32_t b = 0;
                    if (lane == 0) b = atomicAdd(rp.tile_counter, 1u);
                    b = __shfl_sync(0xffffffffu, b, 0);
                    if (b >= rp.n_batches) { more = false; break; }
                    const uint32_t sb = b / rp.n_tiles, tile = b - sb * rp.n_tiles;
                    pool_ty = tile / rp.tiles_x; pool_tx = tile - pool_ty * rp.tiles_x;
                    pool_s0 = rp.spp_begin + sb * rp.batch_spp;
                    pool_next = 0;
                    pool_end = 32u * min(rp.batch_spp, rp.spp_end - pool_s0);
                }
                const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
                const uint32_t rank = __popc(need & lt_mask);
                if (!alive && rank < take) {
                    const uint32_t idx = pool_next + rank;
                    const uint32_t pl = idx & 31u;
                    const uint32_t i = pool_tx * 8 + (pl & 7), j = pool_ty * 4 + (pl >> 3);
                    if (i < rp.width && j < rp.height) {
                        pixel = j * rp.width + i;
                        cur_sample = pool_s0 + (idx >> 5);
                        r = camera_ray(cam, rp, pixel, i, j, cur_sample);
                        beta = make_float3(1.0f, 1.0f, 1.0f);
                        L = make_float3(0.0f, 0.0f, 0.0f);
                        bounce = 0;
                        alive = true;
                        cn.add(ST_PATHS);
                    }
                }
                pool_next += take;
            }
            // ---- start the traversal of every lane that has a ray and is not traversing ----
            if (alive && !trav) trav = !tv.init(r, sc, 0.001f);
            if (!__any_sync(0xffffffffu, alive)) break;  // pool exhausted and every path finished
        }
        // ---- leaf phase: lanes parked at a leaf wait until enough of them have gathered (or nobody is
        //      left descending), then test their primitives together ----
        {
            const uint32_t m_leaf = __ballot_sync(0xffffffffu, trav && tv.at_leaf());
            const uint32_t m_int = __ballot_sync(0xffffffffu, trav && !tv.at_leaf());
            if (m_leaf && (__popc(m_leaf) >= (int)rp.leaf_threshold || m_int == 0u)) {
                if constexpr (SPEC) {
                    // SPEC: a lane parked at a leaf (m_leaf) has either nothing else to visit or a postponed leaf already; lanes
                    // that are still descending bring their postponed leaf along (the phase runs anyway)
                    if (trav && (tv.pend != 0u || tv.at_leaf())) trav = !tv.template spec_leaf_step<STATS, FEAT>(r, sc, stack, cn);
                } else {
                    if (trav && tv.at_leaf()) trav = !tv.template leaf_step<STATS, FEAT>(r, sc, stack, cn);
                }
            }
        }
        // ---- interior phase ----
#pragma unroll 1
        for (uint32_t k = 0; k < rp.steps_per_round; ++k) {
            if constexpr (SPEC) {
                if (trav && !tv.at_leaf()) trav = !tv.template spec_interior_step<STATS>(r, sc, stack, cn);
            } else {
                if (trav && !tv.at_leaf()) trav = !tv.template interior_step<STATS>(r, sc, stack, cn);
            }
        }
    }
    if (STATS) cn.flush(rp.stats);
}

// ---------------------------------------------------------------------------------------------
// BVH megakernel, second schedule: a per-WARP ray queue in shared memory.  ncu on the state machine above
// (profiles/r02_j, r02_k): the interior phase runs at 18.5-18.9 of 32 lanes — a lane whose traversal is finished sits idle
// until >= 24 lanes have gathered for the service phase.  Here a warp owns SLOTS (> 32) path slots in shared memory
// (ray, hit, path state: 18 words each, field-major so that lanes touching different slots hit different banks) and
// two rings of slot numbers:
//     ready   rays waiting to be traversed          done   finished traversals waiting to be shaded
// A lane whose traversal ends pushes its slot on `done` and takes the next slot from `ready` AT ONCE (ballot + popc
// ranks), so every lane is traversing as long as `ready` is not empty; when `done` holds 32 slots the service phase
// shades them at 32 of 32 lanes (whichever lanes produced them), replaces dead paths from the pool and pushes the new rays
// on `ready`.  With 64 slots and 32 lanes traversing, ready + done = 32: `ready` runs empty exactly when `done` is full.
// Only the traversal state lives in registers; the path state (beta, L, pixel, sample, bounce) never does outside the
// service phase.  Same samples, same paths as every other kernel (Philox counters): only the schedule differs.
// ---------------------------------------------------------------------------------------------
enum : uint32_t { QF_OX, QF_OY, QF_OZ, QF_DX, QF_DY, QF_DZ, QF_TIME, QF_T, QF_HIT, QF_PIXEL, QF_SAMPLE, QF_BOUNCE,
                  QF_BR, QF_BG, QF_BB, QF_LR, QF_LG, QF_LB, QF_COUNT };
constexpr uint32_t kQEmpty = 0xFFFFFFFFu;  // QF_PIXEL of a slot without a path

constexpr int kBvhqSlots = 64;

template <bool STATS, uint32_t FEAT, int SLOTS, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) k_megakernel_bvhq(const DevScene sc, const DevCamera cam, const DevRender rp) {
    static_assert(SLOTS > 32 && SLOTS <= 255, "slot numbers are bytes; 32 lanes traverse while the rest wait in the rings");
    __shared__ uint32_t q_words[kBlock / 32][QF_COUNT * SLOTS];
    __shared__ uint8_t q_rings[kBlock / 32][2][SLOTS];
    uint32_t *const S = q_words[threadIdx.x >> 5];
    uint8_t *const ready_q = q_rings[threadIdx.x >> 5][0], *const done_q = q_rings[threadIdx.x >> 5][1];
#define QW(f, s) S[(f) * SLOTS + (s)]
#define QF(f, s) __uint_as_float(S[(f) * SLOTS + (s)])
    auto wrap = [](uint32_t x) { return x >= (uint32_t)SLOTS ? x - (uint32_t)SLOTS : x; };

    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t pool_next = 0, pool_end = 0, pool_tx = 0, pool_ty = 0, pool_s0 = 0;
    bool more = true;
    // every slot starts on `done` without a path: the first service phases start the first paths
    for (uint32_t s = lane; s < (uint32_t)SLOTS; s += 32u) { QW(QF_PIXEL, s) = kQEmpty; done_q[s] = (uint8_t)s; }
    uint32_t ready_head = 0, ready_count = 0, done_head = 0, done_count = SLOTS, live = SLOTS;
    __syncwarp();

    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    BvhTraversal tv;
    BvhStack stack;
    tv.sp = 0; tv.cur = 0;
    uint32_t my = 0;   // slot of the ray this lane is traversing
    bool trav = false, fin = false;  // fin: traversal finished, slot not yet pushed on `done`
    for (;;) {
        // ---- service phase: shade up to 32 finished traversals, one per lane ----
        const uint32_t thr = min(rp.service_threshold, live > 32u ? live - 32u : 1u);
        if (done_count >= thr) {
            const uint32_t n = min(32u, done_count);
            const bool mine = lane < n;
            const uint32_t s = mine ? done_q[wrap(done_head + lane)] : 0u;
            done_head = wrap(done_head + n); done_count -= n;
            Ray q{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
            float3 beta = make_float3(1.0f, 1.0f, 1.0f), L = make_float3(0.0f, 0.0f, 0.0f);
            uint32_t pixel = mine ? QW(QF_PIXEL, s) : kQEmpty, cur_sample = 0, bounce = 0;
            bool alive = mine && pixel != kQEmpty;
            if (alive) {
                cn.add(ST_RAYS);
                q = Ray{QF(QF_OX, s), QF(QF_OY, s), QF(QF_OZ, s), QF(QF_DX, s), QF(QF_DY, s), QF(QF_DZ, s), QF(QF_TIME, s)};
                beta = make_float3(QF(QF_BR, s), QF(QF_BG, s), QF(QF_BB, s));
                L = make_float3(QF(QF_LR, s), QF(QF_LG, s), QF(QF_LB, s));
                cur_sample = QW(QF_SAMPLE, s); bounce = QW(QF_BOUNCE, s);
                const uint32_t hit = QW(QF_HIT, s);
                if (hit == kMiss) {  // main.zig:109-112
                    L.x = fmaf(beta.x, rp.bg_r, L.x); L.y = fmaf(beta.y, rp.bg_g, L.y); L.z = fmaf(beta.z, rp.bg_b, L.z);
                    alive = false;
                } else {
                    const DevPrim prim = sc.prims_bvh[hit];
                    const uint32_t prim_id = sc.bvh_prim_id[hit];
                    ++bounce;
                    alive = shade<STATS, FEAT>(sc, rp, q, prim, prim_id, QF(QF_T, s), pixel, cur_sample, bounce, beta, L, cn) &&
                            bounce < rp.max_depth;  // main.zig:105-108
                }
                if (!alive) red_add_v4(rp.accum + pixel, L.x, L.y, L.z, 1.0f);
            }
            // ---- hand new paths to the slots without one ----
            for (;;) {
                const uint32_t need = __ballot_sync(0xffffffffu, mine && !alive);
                if (!need) break;
                if (pool_next == pool_end) {
                    if (!more) break;
                    uint32_t b = 0;
                    if (lane == 0) b = atomicAdd(rp.tile_counter, 1u);
                    b = __shfl_sync(0xffffffffu, b, 0);
                    if (b >= rp.n_batches) { more = false; break; }
                    const uint32_t sb = b / rp.n_tiles, tile = b - sb * rp.n_tiles;
                    pool_ty = tile / rp.tiles_x; pool_tx = tile - pool_ty * rp.tiles_x;
                    pool_s0 = rp.spp_begin + sb * rp.batch_spp;
                    pool_next = 0;
                    pool_end = 32u * min(rp.batch_spp, rp.spp_end - pool_s0);
                }
                const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
                const uint32_t rank = __popc(need & lt_mask);
                if (mine && !alive && rank < take) {
                    const uint32_t idx = pool_next + rank;
                    const uint32_t pl = idx & 31u;
                    const uint32_t i = pool_tx * 8 + (pl & 7), j = pool_ty * 4 + (pl >> 3);
                    if (i < rp.width && j < rp.height) {  // ragged edge tiles: the index is consumed, no path starts
                        pixel = j * rp.width + i;
                        cur_sample = pool_s0 + (idx >> 5);
                        q = camera_ray(cam, rp, pixel, i, j, cur_sample);
                        beta = make_float3(1.0f, 1.0f, 1.0f);
                        L = make_float3(0.0f, 0.0f, 0.0f);
                        bounce = 0;
                        alive = true;
                        cn.add(ST_PATHS);
                    }
                }
                pool_next += take;
            }
            // ---- write the slots back; rays go on `ready`, slots that found no new path are retired ----
            const uint32_t m_go = __ballot_sync(0xffffffffu, alive);
            if (alive) {
                QW(QF_OX, s) = __float_as_uint(q.ox); QW(QF_OY, s) = __float_as_uint(q.oy); QW(QF_OZ, s) = __float_as_uint(q.oz);
                QW(QF_DX, s) = __float_as_uint(q.dx); QW(QF_DY, s) = __float_as_uint(q.dy); QW(QF_DZ, s) = __float_as_uint(q.dz);
                QW(QF_TIME, s) = __float_as_uint(q.time);
                QW(QF_PIXEL, s) = pixel; QW(QF_SAMPLE, s) = cur_sample; QW(QF_BOUNCE, s) = bounce;
                QW(QF_BR, s) = __float_as_uint(beta.x); QW(QF_BG, s) = __float_as_uint(beta.y); QW(QF_BB, s) = __float_as_uint(beta.z);
                QW(QF_LR, s) = __float_as_uint(L.x); QW(QF_LG, s) = __float_as_uint(L.y); QW(QF_LB, s) = __float_as_uint(L.z);
                ready_q[wrap(wrap(ready_head + ready_count) + (uint32_t)__popc(m_go & lt_mask))] = (uint8_t)s;
            }
            ready_count += (uint32_t)__popc(m_go);
            live -= n - (uint32_t)__popc(m_go);
            __syncwarp();
        }
        if (live == 0u) break;  // pool exhausted and every path finished
        // ---- fetch: lanes without a traversal take the next rays of `ready` ----
        {
            const uint32_t need = __ballot_sync(0xffffffffu, !trav);
            const uint32_t take = min((uint32_t)__popc(need), ready_count);
            if (take) {
                const uint32_t rank = __popc(need & lt_mask);
                if (!trav && rank < take) {
                    my = ready_q[wrap(ready_head + rank)];
                    r = Ray{QF(QF_OX, my), QF(QF_OY, my), QF(QF_OZ, my), QF(QF_DX, my), QF(QF_DY, my), QF(QF_DZ, my), QF(QF_TIME, my)};
                    fin = tv.init(r, sc, 0.001f);  // true: empty scene, the ray is a miss already
                    trav = !fin;
                    // shade continues from the unit direction (normalise_ray)
                    QW(QF_DX, my) = __float_as_uint(r.dx); QW(QF_DY, my) = __float_as_uint(r.dy); QW(QF_DZ, my) = __float_as_uint(r.dz);
                }
                ready_head = wrap(ready_head + take); ready_count -= take;
            }
        }
        // ---- leaf phase: lanes parked at a leaf wait until enough of them have gathered (or nobody is left descending) ----
        {
            const uint32_t m_leaf = __ballot_sync(0xffffffffu, trav && tv.at_leaf());
            const uint32_t m_int = __ballot_sync(0xffffffffu, trav && !tv.at_leaf());
            if (m_leaf && (__popc(m_leaf) >= (int)rp.leaf_threshold || m_int == 0u)) {
                if (trav && tv.at_leaf()) { fin = tv.template leaf_step<STATS, FEAT>(r, sc, stack, cn); trav = !fin; }
            }
        }
        // ---- interior phase ----
#pragma unroll 1
        for (uint32_t k = 0; k < rp.steps_per_round; ++k) {
            if (trav && !tv.at_leaf()) { fin = tv.template interior_step<STATS>(r, sc, stack, cn); trav = !fin; }
        }
        // ---- finished traversals go on `done` ----
        {
            const uint32_t m_fin = __ballot_sync(0xffffffffu, fin);
            if (m_fin) {
                if (fin) {
                    QW(QF_T, my) = __float_as_uint(tv.h.t); QW(QF_HIT, my) = tv.h.slot;
                    done_q[wrap(wrap(done_head + done_count) + (uint32_t)__popc(m_fin & lt_mask))] = (uint8_t)my;
                    fin = false;
                }
                done_count += (uint32_t)__popc(m_fin);
                __syncwarp();
            }
        }
    }
#undef QW
#undef QF
    if (STATS) cn.flush(rp.stats);
}

// ---------------------------------------------------------------------------------------------
// K4 resolve: out = 256 * clamp(sqrt(sum / spp), 0, 0.999) truncated to u8, written to row H-1-j.
// Sums up to kMaxResolveBufs accumulation buffers first; with peer access enabled those may live on
// other GPUs, i.e. the cross-GPU reduction and the resolve are one kernel over NVLink peer memory.
// Works on the scanline slab [row_begin, row_end): in the multi-GPU render every GPU resolves its own slab
// from all N buffers (reduce-scatter shaped: each GPU ingests (N-1)/N of ONE buffer over NVLink).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t quantise(float sum, float scale, unsigned int &nan_flag) {
    float c = sqrtf(sum * scale);
    if (!(c == c)) { c = 0.0f; nan_flag = 1u; }  // reference: UB on NaN; here 0
    c = fminf(fmaxf(c, 0.0f), 0.999f);
    return (uint32_t)(256.0f * c);
}

__global__ void __launch_bounds__(256) k_resolve(ResolveArgs a) {
    const uint32_t n = a.row_end * a.width;
    const uint32_t idx = a.row_begin * a.width + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int nan_flag = 0u;
    if (idx < n) {
        float4 s = __ldcs(a.bufs[0] + idx);
        for (uint32_t b = 1; b < a.n_bufs; ++b) {
            const float4 q = __ldcs(a.bufs[b] + idx);
            s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
        }
        const uint32_t j = idx / a.width, i = idx - j * a.width;
        uint8_t *o = a.rgb8 + ((size_t)(a.height - 1u - j) * a.width + i) * 3u;
        o[0] = (uint8_t)quantise(s.x, a.scale, nan_flag);
        o[1] = (uint8_t)quantise(s.y, a.scale, nan_flag);
        o[2] = (uint8_t)quantise(s.z, a.scale, nan_flag);
    }
    if (a.nan_counter) {
        const unsigned int m = __ballot_sync(0xffffffffu, nan_flag != 0u);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(a.nan_counter, (unsigned long long)__popc(m));
    }
}

// Same, four pixels per thread when the width is a multiple of four: 4 x 16-byte loads per buffer, one
// 12-byte (3 x u32) store — the byte-wise stores of k_resolve reach only ~1.2 TB/s.
__global__ void __launch_bounds__(256) k_resolve4(ResolveArgs a) {
    const uint32_t n4 = (a.row_end * a.width) >> 2;
    const uint32_t q = ((a.row_begin * a.width) >> 2) + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int nan_flag = 0u;
    if (q < n4) {
        const uint32_t idx = q << 2;
        float4 s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = __ldcs(a.bufs[0] + idx + k);
        for (uint32_t b = 1; b < a.n_bufs; ++b) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 v = __ldcs(a.bufs[b] + idx + k);
                s[k].x += v.x; s[k].y += v.y; s[k].z += v.z; s[k].w += v.w;
            }
        }
        uint32_t c[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned int f = 0u;
            c[3 * k + 0] = quantise(s[k].x, a.scale, f);
            c[3 * k + 1] = quantise(s[k].y, a.scale, f);
            c[3 * k + 2] = quantise(s[k].z, a.scale, f);
            nan_flag += f;  // NaN pixels among this thread's four
        }
        const uint32_t j = idx / a.width, i = idx - j * a.width;
        uint32_t *o = reinterpret_cast<uint32_t *>(a.rgb8 + ((size_t)(a.height - 1u - j) * a.width + i) * 3u);  // 12-byte aligned: i % 4 == 0
        o[0] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
        o[1] = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
        o[2] = c[8] | (c[9] << 8) | (c[10] << 16) | (c[11] << 24);
    }
    if (a.nan_counter) {
        const unsigned int tot = __reduce_add_sync(0xffffffffu, nan_flag);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(a.nan_counter, (unsigned long long)tot);
    }
}

// ---------------------------------------------------------------------------------------------
// Production-arithmetic probe: closest hit of explicit rays with exactly the code the megakernel
// runs (robust sphere form, FMA contraction on).
// ---------------------------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(kBlock) k_probe(const DevScene sc, uint32_t n, const float *__restrict__ rays,
                                                  uint32_t *prim_id, float *t_out, float *normal, float *uv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    if (VARIANT == VAR_FLAT) stage_flat(sc, s_flat);
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = idx < n;
    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    if (active) {
        const float *q = rays + 7 * (size_t)idx;
        r = Ray{q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
    }
    Counters<false> cn;
    Hit h{0.0f, kMiss};
    float rl = 1.0f;  // 1/|d|: the searches work with the unit direction and return distances
    if (VARIANT == VAR_FLAT) h = closest_hit_flat<false>(r, active, s_flat, sc.flat, sc, 0.001f, cn, &rl);
    else if (active) h = closest_hit_bvh<false>(r, sc, 0.001f, cn, &rl);
    if (!active) return;
    if (h.slot == kMiss) {
        prim_id[idx] = kMiss; t_out[idx] = 0.0f;
        normal[3 * idx] = normal[3 * idx + 1] = normal[3 * idx + 2] = 0.0f;
        if (uv) { uv[2 * idx] = 0.0f; uv[2 * idx + 1] = 0.0f; }
        return;
    }
    DevPrim prim;
    uint32_t id;
    if (VARIANT == VAR_FLAT) { id = h.slot; prim = sc.prims_flat[id]; }
    else { prim = sc.prims_bvh[h.slot]; id = sc.bvh_prim_id[h.slot]; }
    const Surface s = finalise_hit<false>(r, prim, h.t, sc, cn);
    prim_id[idx] = id; t_out[idx] = h.t * rl;  // the reference's parameter t = s / |d|
    normal[3 * idx] = s.nx; normal[3 * idx + 1] = s.ny; normal[3 * idx + 2] = s.nz;
    if (uv) {
        float u = __fdividef(s.u, s.ru), v = __fdividef(s.v, s.rv);  // rect: numerator / denominator (lazy division)
        if (s.is_sphere) sphere_uv(s.onx, s.ony, s.onz, u, v);
        uv[2 * idx] = u; uv[2 * idx + 1] = v;
    }
}

// max_depth == 0: rayColor returns black before intersecting (main.zig:105-108); only the sample count moves
__global__ void __launch_bounds__(256) k_add_samples(float4 *accum, uint32_t n, float count) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) accum[idx].w += count;
}
cudaError_t launch_add_samples(float4 *accum, uint32_t n_pixels, float count, cudaStream_t st) {
    k_add_samples<<<(n_pixels + 255) / 256, 256, 0, st>>>(accum, n_pixels, count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FP32 peak: 8 independent FFMA chains per thread, 2 flops per FFMA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ffma_peak(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from rtw_api.cpp through rtw_kernels.h)
// ---------------------------------------------------------------------------------------------
// POOLED: 0 = deterministic lane-owns-pixel kernel, 1 = pooled (flat: first schedule), 2 = flat second schedule at 9 CTAs
// per SM (56 registers, a few spills), 3 = the same at 8 CTAs per SM (64 registers), 4 / 5 / 6 = 3 specialised on the scene's
// features (FF_SPHERES, FF_SPHERES | FF_TEX, FF_RECTS; event-counter builds always run the generic kernel)
template <int VARIANT, bool STATS, int POOLED>
static auto mega_kernel_ptr() {
    if constexpr (POOLED == 10 && VARIANT == VAR_BVH) return k_megakernel_bvh<STATS, FF_SPHERES, true>;  // speculative schedule
    else if constexpr (POOLED == 9 && VARIANT == VAR_BVH) return k_megakernel_bvh<STATS, FF_ALL, true>;
    else if constexpr (POOLED == 8 && VARIANT == VAR_BVH) return k_megakernel_bvhq<STATS, FF_SPHERES, kBvhqSlots, 8>;  // ray-queue schedule
    else if constexpr (POOLED == 7 && VARIANT == VAR_BVH) return k_megakernel_bvhq<STATS, FF_ALL, kBvhqSlots, 7>;
    else if constexpr (POOLED == 4 && VARIANT == VAR_BVH) return k_megakernel_bvh<STATS, FF_SPHERES>;  // spheres, solid / checker textures
    else if constexpr (POOLED && VARIANT == VAR_BVH) return k_megakernel_bvh<STATS, FF_ALL>;
    else if constexpr (POOLED == 4) return k_megakernel_flat<STATS, 9, FF_SPHERES>;            // spheres, solid / checker textures
    else if constexpr (POOLED == 5) return k_megakernel_flat<STATS, 8, FF_SPHERES | FF_TEX>;   // spheres, every texture kind
    else if constexpr (POOLED == 6) return k_megakernel_flat<STATS, 8, FF_RECTS>;              // rects and boxes only, solid / checker
    else if constexpr (POOLED == 3) return k_megakernel_flat<STATS, 8, FF_ALL>;
    else if constexpr (POOLED == 2) return k_megakernel_flat<STATS, 9, FF_ALL>;
    else if constexpr (POOLED) return k_megakernel_pooled<VARIANT, STATS>;
    else return k_megakernel<VARIANT, STATS>;
}

template <int VARIANT, bool STATS, int POOLED>
static cudaError_t launch_mega_t(const DevScene &sc, const DevCamera &cam, const DevRender &rp, int grid,
                                 size_t smem, cudaStream_t st) {
    auto kern = mega_kernel_ptr<VARIANT, STATS, POOLED>();
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, kBlock, smem, st>>>(sc, cam, rp);
    return cudaGetLastError();
}

// BVH kernels: 1 = state machine, 4 = its spheres-only specialisation, 7 = ray-queue schedule, 8 = its spheres-only
// specialisation; event-counter builds run the generic kernel of the schedule
static int bvh_pooled(int pooled, bool stats) {
    if (pooled == 7 || pooled == 8) return (pooled == 8 && !stats) ? 8 : 7;
    if (pooled == 9 || pooled == 10) return (pooled == 10 && !stats) ? 10 : 9;
    return (pooled == 4 && !stats) ? 4 : 1;
}

static size_t mega_smem(int variant, const DevScene &sc) {
    return variant == VAR_FLAT ? (size_t)sc.flat.total_f4 * sizeof(float4) : 0;
}

cudaError_t launch_megakernel(int variant, bool stats, int pooled, const DevScene &sc, const DevCamera &cam,
                              const DevRender &rp, int grid, cudaStream_t st) {
    const size_t smem = mega_smem(variant, sc);
    if (variant == VAR_BVH && pooled >= 1) pooled = bvh_pooled(pooled, stats);
#define RTW_CASE(V, S, P) if (variant == V && stats == S && pooled == P) return launch_mega_t<V, S, P>(sc, cam, rp, grid, smem, st)
    RTW_CASE(VAR_FLAT, false, 0); RTW_CASE(VAR_FLAT, true, 0); RTW_CASE(VAR_FLAT, false, 1); RTW_CASE(VAR_FLAT, true, 1);
    RTW_CASE(VAR_FLAT, false, 2); RTW_CASE(VAR_FLAT, true, 2); RTW_CASE(VAR_FLAT, false, 3); RTW_CASE(VAR_FLAT, true, 3);
    RTW_CASE(VAR_FLAT, false, 4); RTW_CASE(VAR_FLAT, false, 5); RTW_CASE(VAR_FLAT, false, 6);
    RTW_CASE(VAR_BVH, false, 0); RTW_CASE(VAR_BVH, true, 0); RTW_CASE(VAR_BVH, false, 1); RTW_CASE(VAR_BVH, true, 1);
    RTW_CASE(VAR_BVH, false, 4); RTW_CASE(VAR_BVH, false, 7); RTW_CASE(VAR_BVH, true, 7); RTW_CASE(VAR_BVH, false, 8);
    RTW_CASE(VAR_BVH, false, 9); RTW_CASE(VAR_BVH, true, 9); RTW_CASE(VAR_BVH, false, 10);
#undef RTW_CASE
    return cudaErrorInvalidValue;
}

template <int VARIANT, bool STATS, int POOLED>
static int occ_t(size_t smem) {
    auto kern = mega_kernel_ptr<VARIANT, STATS, POOLED>();
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n = 0;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBlock, smem) == cudaSuccess ? n : 0;
}

int megakernel_ctas_per_sm(int variant, bool stats, int pooled, const DevScene &sc) {
    const size_t smem = mega_smem(variant, sc);
    if (variant == VAR_BVH && pooled >= 1) pooled = bvh_pooled(pooled, stats);
#define RTW_CASE(V, S, P) if (variant == V && stats == S && pooled == P) return occ_t<V, S, P>(smem)
    RTW_CASE(VAR_FLAT, false, 0); RTW_CASE(VAR_FLAT, true, 0); RTW_CASE(VAR_FLAT, false, 1); RTW_CASE(VAR_FLAT, true, 1);
    RTW_CASE(VAR_FLAT, false, 2); RTW_CASE(VAR_FLAT, true, 2); RTW_CASE(VAR_FLAT, false, 3); RTW_CASE(VAR_FLAT, true, 3);
    RTW_CASE(VAR_FLAT, false, 4); RTW_CASE(VAR_FLAT, false, 5); RTW_CASE(VAR_FLAT, false, 6);
    RTW_CASE(VAR_BVH, false, 0); RTW_CASE(VAR_BVH, true, 0); RTW_CASE(VAR_BVH, false, 1); RTW_CASE(VAR_BVH, true, 1);
    RTW_CASE(VAR_BVH, false, 4); RTW_CASE(VAR_BVH, false, 7); RTW_CASE(VAR_BVH, true, 7); RTW_CASE(VAR_BVH, false, 8);
    RTW_CASE(VAR_BVH, false, 9); RTW_CASE(VAR_BVH, true, 9); RTW_CASE(VAR_BVH, false, 10);
#undef RTW_CASE
    return 0;
}

cudaError_t launch_resolve(const ResolveArgs &a, cudaStream_t st) {
    const uint32_t n = a.width * (a.row_end - a.row_begin);  // rows [row_begin, row_end) of the accumulation buffers
    if (n == 0) return cudaSuccess;
    // 4-pixel path needs rows that are multiples of 4 pixels (12-byte groups stay 4-byte aligned) and an aligned base
    if ((a.width & 3u) == 0 && (reinterpret_cast<uintptr_t>(a.rgb8) & 3u) == 0) k_resolve4<<<((n >> 2) + 255) / 256, 256, 0, st>>>(a);
    else k_resolve<<<(n + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_probe(int variant, const DevScene &sc, uint32_t n, const float *rays, uint32_t *prim_id, float *t,
                         float *normal, float *uv, cudaStream_t st) {
    const int grid = (int)((n + kBlock - 1) / kBlock);
    if (variant == VAR_FLAT) {
        const size_t smem = mega_smem(variant, sc);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_probe<VAR_FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k_probe<VAR_FLAT><<<grid, kBlock, smem, st>>>(sc, n, rays, prim_id, t, normal, uv);
    } else {
        k_probe<VAR_BVH><<<grid, kBlock, 0, st>>>(sc, n, rays, prim_id, t, normal, uv);
    }
    return cudaGetLastError();
}

cudaError_t launch_ffma_peak(float *out, int grid, int iters, cudaStream_t st) {
    k_ffma_peak<<<grid, 256, 0, st>>>(out, iters, 1.0000001f, 1e-9f);
    return cudaGetLastError();
}

}  // namespace rtw
