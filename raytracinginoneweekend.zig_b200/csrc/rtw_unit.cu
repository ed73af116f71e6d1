// rtw_unit.cu — unit-level probes of the production STOCHASTIC device code (parity instruments, not timed).
//
// The image gates of the parity suite see camera_ray / the sampling maps / shade only through converged means; these
// kernels call the very same __device__ functions the megakernels call (rtw_trace.cuh, same compile flags as
// rtw_kernels.cu) for caller-chosen indices and hand back both the result AND the random choices it was made from,
// so the host can replay Camera.getRay (src/main.zig:91-100) and Material.scatter (src/rtw/material.zig:22-110) in the
// f64 oracle with exactly those choices and compare to fp32 rounding.
#include <cuda_runtime.h>

#include "rtw_kernels.h"
#include "rtw_trace.cuh"

namespace rtw {

// out[n][14]: ray (o, d, time), the five uniforms (ju, jv, l1, l2, tm), the lens-disk point they map to
__global__ void __launch_bounds__(128) k_unit_camera(const DevCamera cam, const DevRender rp, uint32_t n, const uint32_t *__restrict__ ijs,
                                                     float *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t i = ijs[3 * k], j = ijs[3 * k + 1], sample = ijs[3 * k + 2];
    const uint32_t pixel = j * rp.width + i;
    const Ray r = camera_ray(cam, rp, pixel, i, j, sample);
    // the same slicing of the same Philox block as camera_ray (kept in one place: camera_uniforms)
    float u[5];
    camera_uniforms(rp, pixel, sample, u);
    const Draw dw = make_draw(philox4x32_10(make_uint4(pixel, sample, 0u, 0u), rp.philox_keys), true);
    const float2 dk = make_float2(dw.x, dw.y);  // the lens-disk point camera_ray used
    float *o = out + 14 * (size_t)k;
    o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz; o[6] = r.time;
    o[7] = u[0]; o[8] = u[1]; o[9] = u[2]; o[10] = u[3]; o[11] = u[4];
    o[12] = dk.x; o[13] = dk.y;
}

// out[n][8]: the maps of draw_from for caller-supplied uniforms: unit vector (u1,u2) | ball point (u1,u2,u3) | disk point (u1,u2)
__global__ void __launch_bounds__(128) k_unit_samplers(uint32_t n, const float *__restrict__ u3, float *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float a = u3[3 * k], b = u3[3 * k + 1], c = u3[3 * k + 2];
    const Draw sv = draw_from(a, b, c, 0.f, 0.f, false), sd = draw_from(a, b, c, 0.f, 0.f, true);
    const float kr = cbrtf(sv.uc);  // scatter_finish: reflected + fuzz * cbrt(uc) * unit vector
    const float3 v = make_float3(sv.x, sv.y, sv.z), bl = make_float3(kr * sv.x, kr * sv.y, kr * sv.z);
    const float2 d = make_float2(sd.x, sd.y);
    float *o = out + 8 * (size_t)k;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = bl.x; o[4] = bl.y; o[5] = bl.z; o[6] = d.x; o[7] = d.y;
}

// out[n][8]: the uniforms the production kernels draw for (pixel, sample, block): u01_24 of the four Philox words, and
// the raw words (as float bit patterns) — the distribution tests read these
__global__ void __launch_bounds__(128) k_unit_uniforms(const DevRender rp, uint32_t n, const uint32_t *__restrict__ psb, float *__restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint4 rn = philox4x32_10(make_uint4(psb[3 * k], psb[3 * k + 1], psb[3 * k + 2], 0u), rp.philox_keys);
    float *o = out + 8 * (size_t)k;
    o[0] = u01_24(rn.x); o[1] = u01_24(rn.y); o[2] = u01_24(rn.z); o[3] = u01_24(rn.w);
    o[4] = __uint_as_float(rn.x); o[5] = __uint_as_float(rn.y); o[6] = __uint_as_float(rn.z); o[7] = __uint_as_float(rn.w);
}

// One rayColor level (main.zig:109-121) for explicit rays with production arithmetic: closest hit, then shade() with
// beta = 1, L = 0.  out[n][20]: t | scattered ray (o, d, time) | attenuation | emitted | continues | the sample vector
// shade consumed (diffuse: unit vector, metal: ball point) | the uniform (dielectric) | material kind
template <int VARIANT>
__global__ void __launch_bounds__(128) k_unit_shade(const DevScene sc, const DevRender rp, uint32_t n, const float *__restrict__ rays,
                                                    const uint32_t *__restrict__ psb, uint32_t *__restrict__ prim_id, float *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_flat = reinterpret_cast<float4 *>(smem_raw);
    if (VARIANT == VAR_FLAT) {
        for (uint32_t i = threadIdx.x; i < sc.flat.total_f4; i += blockDim.x) s_flat[i] = sc.flat_blob[i];
        __syncthreads();
    }
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < n;
    Ray r{0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    if (active) {
        const float *q = rays + 7 * (size_t)k;
        r = Ray{q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
    }
    Counters<false> cn;
    Hit h{0.0f, kMiss};
    float rl = 1.0f;  // 1/|d|: the searches normalise r in place and return distances
    if (VARIANT == VAR_FLAT) h = closest_hit_flat<false>(r, active, s_flat, sc.flat, sc, 0.001f, cn, &rl);
    else if (active) h = closest_hit_bvh<false>(r, sc, 0.001f, cn, &rl);
    if (!active) return;
    float *o = out + 20 * (size_t)k;
    for (int q = 0; q < 20; ++q) o[q] = 0.0f;
    if (h.slot == kMiss) { prim_id[k] = kMiss; return; }
    DevPrim prim;
    uint32_t id;
    if (VARIANT == VAR_FLAT) { id = h.slot; prim = sc.prims_flat[id]; }
    else { prim = sc.prims_bvh[h.slot]; id = sc.bvh_prim_id[h.slot]; }
    prim_id[k] = id;
    const uint32_t pixel = psb[3 * k], sample = psb[3 * k + 1], bounce = psb[3 * k + 2];
    const DevMaterial m = sc.materials[sc.prim_material[id]];
    const uint4 rn = philox4x32_10(make_uint4(pixel, sample, bounce, 0u), rp.philox_keys);
    float3 vec = make_float3(0.f, 0.f, 0.f);
    const Draw dw = make_draw(rn, false);  // what scatter_finish consumes
    if (m.kind == 0u) vec = make_float3(dw.x, dw.y, dw.z);
    else if (m.kind == 1u && m.param > 0.0f) { const float k = cbrtf(dw.uc); vec = make_float3(k * dw.x, k * dw.y, k * dw.z); }
    float3 beta = make_float3(1.f, 1.f, 1.f), L = make_float3(0.f, 0.f, 0.f);
    const bool go = shade<false>(sc, rp, r, prim, id, h.t, pixel, sample, bounce, beta, L, cn);
    o[0] = h.t * rl;  // the reference's parameter
    o[1] = r.ox; o[2] = r.oy; o[3] = r.oz; o[4] = r.dx; o[5] = r.dy; o[6] = r.dz; o[7] = r.time;
    o[8] = beta.x; o[9] = beta.y; o[10] = beta.z;
    o[11] = L.x; o[12] = L.y; o[13] = L.z;
    o[14] = go ? 1.0f : 0.0f;
    o[15] = vec.x; o[16] = vec.y; o[17] = vec.z;
    o[18] = u01_24(rn.x);
    o[19] = (float)m.kind;
}

cudaError_t launch_unit_camera(const DevCamera &cam, const DevRender &rp, uint32_t n, const uint32_t *ijs, float *out, cudaStream_t st) {
    k_unit_camera<<<(n + 127) / 128, 128, 0, st>>>(cam, rp, n, ijs, out);
    return cudaGetLastError();
}
cudaError_t launch_unit_samplers(uint32_t n, const float *u3, float *out, cudaStream_t st) {
    k_unit_samplers<<<(n + 127) / 128, 128, 0, st>>>(n, u3, out);
    return cudaGetLastError();
}
cudaError_t launch_unit_uniforms(const DevRender &rp, uint32_t n, const uint32_t *psb, float *out, cudaStream_t st) {
    k_unit_uniforms<<<(n + 127) / 128, 128, 0, st>>>(rp, n, psb, out);
    return cudaGetLastError();
}
cudaError_t launch_unit_shade(int variant, const DevScene &sc, const DevRender &rp, uint32_t n, const float *rays, const uint32_t *psb,
                              uint32_t *prim_id, float *out, cudaStream_t st) {
    const int grid = (int)((n + 127) / 128);
    if (variant == VAR_FLAT) {
        const size_t smem = (size_t)sc.flat.total_f4 * sizeof(float4);
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_unit_shade<VAR_FLAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k_unit_shade<VAR_FLAT><<<grid, 128, smem, st>>>(sc, rp, n, rays, psb, prim_id, out);
    } else {
        k_unit_shade<VAR_BVH><<<grid, 128, 0, st>>>(sc, rp, n, rays, psb, prim_id, out);
    }
    return cudaGetLastError();
}

}  // namespace rtw
