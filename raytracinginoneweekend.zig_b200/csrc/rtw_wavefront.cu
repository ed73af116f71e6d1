// rtw_wavefront.cu — K2: the wavefront alternative to the megakernel (RTW_VARIANT_WAVEFRONT).
//
// Same path (src/main.zig:382-394 and below), same device functions (rtw_trace.cuh), different schedule.
// Path state lives in global memory (64 B per slot, SoA of float4) and every iteration of the host loop runs
// three kernels over the N slots:
//
//   k_wf_generate  refills the slots freed by the previous iteration with the next paths (camera rays);
//                  the free-slot list it consumes was COMPACTED by k_wf_shade with warp ballot + popc prefix
//                  sums and one atomicAdd per warp;
//   k_wf_extend    closest hit of every live ray.  Flat scenes: one thread per slot, the warp-voted group scan at
//                  32/32 lanes (no regeneration or shading inside the loop).  BVH scenes: persistent warps with
//                  dynamic ray fetch — a lane whose traversal finishes takes the next ray of the warp's pool, so
//                  every step runs with all lanes traversing;
//   k_wf_shade     miss / emit / scatter (material divergence is confined to this kernel), adds finished paths
//                  to the frame with one vector reduction, appends dead slots to the free list.
//
// Because freed slots are refilled at once the wave stays full: the number of iterations is total rays / N, not
// max_depth per wave.  Counters are double-buffered by iteration parity so no kernel zeroes what another reads.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "rtw_kernels.h"
#include "rtw_trace.cuh"

namespace rtw {

constexpr int kWfBlock = 128;
constexpr uint32_t kDead = 0xFFFFFFFFu;

__device__ __forceinline__ void wf_red_add_v4(float4 *addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// path index -> (pixel, sample): tile-major like the megakernel's batches, so that the 32 paths of a warp are the
// 32 pixels of one 8x4 tile (coherent primary rays)
__device__ __forceinline__ bool wf_path_coords(const DevRender &rp, unsigned long long path, uint32_t &pixel, uint32_t &i,
                                               uint32_t &j, uint32_t &sample) {
    // dense numbering: full sample blocks of batch_spp samples, then one shorter last block
    const uint32_t spp = rp.spp_end - rp.spp_begin;
    const unsigned long long per_full = 32ull * rp.batch_spp * rp.n_tiles;
    const uint32_t sb = min((uint32_t)(path / per_full), rp.n_sblocks - 1u);
    const unsigned long long in_sb = path - (unsigned long long)sb * per_full;
    const uint32_t s0 = sb * rp.batch_spp;
    const uint32_t ns = min(rp.batch_spp, spp - s0);  // samples in this block
    const uint32_t tile = (uint32_t)(in_sb / (32u * ns));
    const uint32_t within = (uint32_t)(in_sb - (unsigned long long)tile * 32u * ns);
    const uint32_t pl = within & 31u, s = within >> 5;
    const uint32_t ty = tile / rp.tiles_x, tx = tile - ty * rp.tiles_x;
    i = tx * 8 + (pl & 7);
    j = ty * 4 + (pl >> 3);
    sample = rp.spp_begin + s0 + s;
    pixel = j * rp.width + i;
    return i < rp.width && j < rp.height;  // false only in ragged edge tiles
}

template <bool STATS>
__global__ void __launch_bounds__(kWfBlock) k_wf_generate(WfState st, const DevCamera cam, const DevRender rp, uint32_t parity) {
    Counters<STATS> cn;
    const uint32_t n_free = st.counters->n_free[parity];
    const unsigned long long base = st.counters->next_path[parity];
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0)  // hand the path cursor to the next iteration (n_free[parity^1] was zeroed by the previous shade)
        st.counters->next_path[parity ^ 1u] = min(base + n_free, st.total_paths);
    bool back = false;  // slot goes straight back to the free list (no path for it)
    uint32_t slot = 0;
    if (idx < n_free) {
        slot = st.free_list[parity][idx];
        const unsigned long long path = base + idx;
        uint32_t pixel, i, j, sample;
        if (path < st.total_paths && wf_path_coords(rp, path, pixel, i, j, sample)) {
            const Ray r = camera_ray(cam, rp, pixel, i, j, sample);
            st.ro[slot] = make_float4(r.ox, r.oy, r.oz, r.time);
            st.rd[slot] = make_float4(r.dx, r.dy, r.dz, 0.0f);
            st.beta[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel));
            st.rad[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(sample << 6));  // bounce in the low 6 bits
            cn.add(ST_PATHS);
        } else {
            // ragged index outside the image, or the frame is exhausted
            st.beta[slot].w = __uint_as_float(kDead);
            back = true;
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, back);  // warp-aggregated append
    if (m) {
        const uint32_t lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        uint32_t at = 0;
        if (lane == leader) at = atomicAdd(&st.counters->n_free[parity ^ 1u], (uint32_t)__popc(m));
        at = __shfl_sync(0xffffffffu, at, leader);
        if (back) st.free_list[parity ^ 1u][at + __popc(m & ((1u << lane) - 1u))] = slot;
    }
    if (STATS) cn.flush(rp.stats);
}

// ---- extend, flat scenes: one thread per slot, fully converged warps ------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(kWfBlock) k_wf_extend_flat(WfState st, const DevScene sc, const DevRender rp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    for (uint32_t i = threadIdx.x; i < sc.flat.total_f4; i += blockDim.x) s_flat[i] = sc.flat_blob[i];
    __syncthreads();
    Counters<STATS> cn;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;  // grid covers exactly the slot count
    const float4 o = st.ro[slot], d = st.rd[slot];
    const bool active = __float_as_uint(st.beta[slot].w) != kDead;
    Ray r{o.x, o.y, o.z, d.x, d.y, d.z, o.w};
    const Hit h = closest_hit_flat<STATS>(r, active, s_flat, sc.flat, sc, 0.001f, cn);
    if (active) {
        st.rd[slot] = make_float4(r.dx, r.dy, r.dz, h.t);  // the unit direction the search left, and the distance along it
        st.hit[slot] = h.slot;  // flat: prim id
    }
    if (STATS) cn.flush(rp.stats);
}

// ---- extend, BVH scenes: persistent warps, dynamic ray fetch ------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(kWfBlock) k_wf_extend_bvh(WfState st, const DevScene sc, const DevRender rp) {
    Counters<STATS> cn;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t pool_next = 0, pool_end = 0;
    bool more = true;
    bool trav = false;
    uint32_t slot = 0;
    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    BvhTraversal tv;
    BvhStack stack;
    tv.sp = 0; tv.cur = 0;
    constexpr uint32_t kChunk = 32 * 8;
    for (;;) {
        // hand new rays to idle lanes
        for (;;) {
            const uint32_t need = __ballot_sync(0xffffffffu, !trav);
            if (!need) break;
            if (pool_next == pool_end) {
                if (!more) break;
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(st.extend_cursor, kChunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= st.n_slots) { more = false; break; }
                pool_next = b;
                pool_end = min(b + kChunk, st.n_slots);
            }
            const uint32_t take = min((uint32_t)__popc(need), pool_end - pool_next);
            const uint32_t rank = __popc(need & lt_mask);
            if (!trav && rank < take) {
                slot = pool_next + rank;
                if (__float_as_uint(st.beta[slot].w) != kDead) {
                    const float4 o = st.ro[slot], d = st.rd[slot];
                    r = Ray{o.x, o.y, o.z, d.x, d.y, d.z, o.w};
                    if (tv.init(r, sc, 0.001f)) { st.rd[slot] = make_float4(r.dx, r.dy, r.dz, tv.h.t); st.hit[slot] = kMiss; }
                    else trav = true;
                }
            }
            pool_next += take;
        }
        if (!__any_sync(0xffffffffu, trav)) {
            if (!more && pool_next == pool_end) break;
            continue;
        }
        // leaf phase (gathered: run when >= leaf_threshold lanes are parked at a leaf or nobody descends), then
        // one interior step
        bool done = false;
        const uint32_t m_leaf = __ballot_sync(0xffffffffu, trav && tv.at_leaf());
        const uint32_t m_int = __ballot_sync(0xffffffffu, trav && !tv.at_leaf());
        if (m_leaf && (__popc(m_leaf) >= (int)rp.leaf_threshold || m_int == 0u)) {
            if (trav && tv.at_leaf()) done = tv.template leaf_step<STATS>(r, sc, stack, cn);
        }
        if (trav && !done && !tv.at_leaf()) done = tv.template interior_step<STATS>(r, sc, stack, cn);
        if (done) {
            trav = false;
            st.rd[slot] = make_float4(r.dx, r.dy, r.dz, tv.h.t);  // unit direction + distance
            st.hit[slot] = tv.h.slot;  // BVH: leaf-order slot
        }
    }
    if (STATS) cn.flush(rp.stats);
}

// ---- shade: miss / emit / scatter, accumulate finished paths, compact the dead slots --------------------------------
template <int VARIANT, bool STATS>
__global__ void __launch_bounds__(kWfBlock) k_wf_shade(WfState st, const DevScene sc, const DevRender rp, uint32_t parity) {
    Counters<STATS> cn;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    if (slot == 0) st.counters->n_free[parity] = 0;  // the list this iteration's generate consumed
    const float4 bp = st.beta[slot];
    const uint32_t pixel = __float_as_uint(bp.w);
    bool freed = false;
    if (pixel != kDead) {
        const float4 o = st.ro[slot], d = st.rd[slot], lm = st.rad[slot];
        Ray r{o.x, o.y, o.z, d.x, d.y, d.z, o.w};
        float3 beta = make_float3(bp.x, bp.y, bp.z), L = make_float3(lm.x, lm.y, lm.z);
        const uint32_t meta = __float_as_uint(lm.w);
        const uint32_t sample = meta >> 6;
        uint32_t bounce = meta & 63u;
        const uint32_t hslot = st.hit[slot];
        cn.add(ST_RAYS);
        bool alive;
        if (hslot == kMiss) {  // main.zig:109-112
            L.x = fmaf(beta.x, rp.bg_r, L.x); L.y = fmaf(beta.y, rp.bg_g, L.y); L.z = fmaf(beta.z, rp.bg_b, L.z);
            alive = false;
        } else {
            DevPrim prim;
            uint32_t prim_id;
            if (VARIANT == VAR_FLAT) { prim_id = hslot; prim = sc.prims_flat[prim_id]; }
            else { prim = sc.prims_bvh[hslot]; prim_id = sc.bvh_prim_id[hslot]; }
            ++bounce;
            alive = shade<STATS>(sc, rp, r, prim, prim_id, d.w, pixel, sample, bounce, beta, L, cn) && bounce < rp.max_depth;
        }
        if (alive) {
            st.ro[slot] = make_float4(r.ox, r.oy, r.oz, r.time);
            st.rd[slot] = make_float4(r.dx, r.dy, r.dz, 0.0f);
            st.beta[slot] = make_float4(beta.x, beta.y, beta.z, bp.w);
            st.rad[slot] = make_float4(L.x, L.y, L.z, __uint_as_float((sample << 6) | bounce));
        } else {
            wf_red_add_v4(rp.accum + pixel, L.x, L.y, L.z, 1.0f);
            st.beta[slot].w = __uint_as_float(kDead);
            freed = true;
        }
    }
    // compaction of the free-slot list: ballot + popc prefix, one atomicAdd per warp
    const uint32_t m = __ballot_sync(0xffffffffu, freed);
    if (m) {
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&st.counters->n_free[parity ^ 1u], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (freed) st.free_list[parity ^ 1u][base + __popc(m & ((1u << lane) - 1u))] = slot;
    }
    if (STATS) cn.flush(rp.stats);
}

__global__ void k_wf_init(WfState st) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < st.n_slots) {
        st.free_list[0][idx] = idx;
        st.beta[idx].w = __uint_as_float(kDead);
    }
    if (idx == 0) {
        st.counters->n_free[0] = st.n_slots;
        st.counters->n_free[1] = 0;
        st.counters->next_path[0] = 0;
        st.counters->next_path[1] = 0;
    }
}

// ---- host driver ---------------------------------------------------------------------------------------------------
cudaError_t wavefront_accumulate(WfState &st, int variant, bool stats, const DevScene &sc, const DevCamera &cam,
                                 const DevRender &rp, int n_sms, WfCounters *h_counters, cudaStream_t stream,
                                 uint32_t *n_launches_out) {
    const uint32_t spp = rp.spp_end - rp.spp_begin;
    st.total_paths = 32ull * rp.n_tiles * spp;
    const int grid = (int)(st.n_slots / kWfBlock);
    uint32_t launches = 0;
    cudaError_t e;
    k_wf_init<<<grid, kWfBlock, 0, stream>>>(st);
    ++launches;
    const size_t smem = variant == VAR_FLAT ? (size_t)sc.flat.total_f4 * sizeof(float4) : 0;
    if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(k_wf_extend_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_wf_extend_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    }
    int bvh_occ = 0;
    if (variant == VAR_BVH) {
        if (stats) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bvh_occ, k_wf_extend_bvh<true>, kWfBlock, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bvh_occ, k_wf_extend_bvh<false>, kWfBlock, 0);
        if (bvh_occ <= 0) return cudaErrorLaunchOutOfResources;
    }
    const uint32_t check_every = 4;
    for (uint32_t it = 0;; ++it) {
        const uint32_t parity = it & 1u;
        if (stats) k_wf_generate<true><<<grid, kWfBlock, 0, stream>>>(st, cam, rp, parity);
        else k_wf_generate<false><<<grid, kWfBlock, 0, stream>>>(st, cam, rp, parity);
        if (variant == VAR_FLAT) {
            if (stats) k_wf_extend_flat<true><<<grid, kWfBlock, smem, stream>>>(st, sc, rp);
            else k_wf_extend_flat<false><<<grid, kWfBlock, smem, stream>>>(st, sc, rp);
        } else {
            if ((e = cudaMemsetAsync(st.extend_cursor, 0, sizeof(uint32_t), stream)) != cudaSuccess) return e;
            if (stats) k_wf_extend_bvh<true><<<bvh_occ * n_sms, kWfBlock, 0, stream>>>(st, sc, rp);
            else k_wf_extend_bvh<false><<<bvh_occ * n_sms, kWfBlock, 0, stream>>>(st, sc, rp);
        }
        if (variant == VAR_FLAT) {
            if (stats) k_wf_shade<VAR_FLAT, true><<<grid, kWfBlock, 0, stream>>>(st, sc, rp, parity);
            else k_wf_shade<VAR_FLAT, false><<<grid, kWfBlock, 0, stream>>>(st, sc, rp, parity);
        } else {
            if (stats) k_wf_shade<VAR_BVH, true><<<grid, kWfBlock, 0, stream>>>(st, sc, rp, parity);
            else k_wf_shade<VAR_BVH, false><<<grid, kWfBlock, 0, stream>>>(st, sc, rp, parity);
        }
        launches += 3;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((it + 1) % check_every == 0) {
            // finished when the frame's paths are all handed out and every slot is free again
            if ((e = cudaMemcpyAsync(h_counters, st.counters, sizeof(WfCounters), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
            const uint32_t np = (it + 1) & 1u;  // the counters the NEXT iteration will read
            if (h_counters->next_path[np] >= st.total_paths && h_counters->n_free[np] == st.n_slots) break;
        }
        if (it > (1u << 24)) return cudaErrorUnknown;  // cannot happen: every iteration advances each live path
    }
    if (n_launches_out) *n_launches_out = launches;
    return cudaSuccess;
}

}  // namespace rtw
