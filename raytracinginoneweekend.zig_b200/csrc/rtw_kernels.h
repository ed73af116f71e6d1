// rtw_kernels.h — launcher interface between the C-ABI layer (rtw_api.cpp) and the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtw_device.cuh"

namespace rtw {

enum { VAR_FLAT = 1, VAR_BVH = 2 };
constexpr uint32_t kMaxResolveBufs = 8;

struct ResolveArgs {
    const float4 *bufs[kMaxResolveBufs];
    uint32_t n_bufs;
    uint32_t width, height;
    uint32_t row_begin, row_end;  // reference scanlines j (bottom row = 0) this launch resolves
    float scale;  // 1 / spp_total
    uint8_t *rgb8;
    unsigned long long *nan_counter;  // may be null
};

// rtw_kernels.cu (production arithmetic)
// pooled: 0 = deterministic lane-owns-pixel kernel, 1 = pooled path queue (BVH state machine / flat first schedule),
//         2 / 3 = flat second schedule (k_megakernel_flat) at 9 / 8 CTAs per SM, 4 / 5 / 6 = 3 specialised on the scene's features;
//         BVH variant: 7 / 8 = ray-queue schedule (generic / spheres only), 4 = spheres-only state machine, every other value >= 1 = state machine
cudaError_t launch_megakernel(int variant, bool stats, int pooled, const DevScene &sc, const DevCamera &cam,
                              const DevRender &rp, int grid, cudaStream_t st);
int megakernel_ctas_per_sm(int variant, bool stats, int pooled, const DevScene &sc);
cudaError_t launch_resolve(const ResolveArgs &a, cudaStream_t st);
cudaError_t launch_probe(int variant, const DevScene &sc, uint32_t n, const float *rays, uint32_t *prim_id, float *t,
                         float *normal, float *uv, cudaStream_t st);
cudaError_t launch_ffma_peak(float *out, int grid, int iters, cudaStream_t st);
cudaError_t launch_add_samples(float4 *accum, uint32_t n_pixels, float count, cudaStream_t st);

// rtw_unit.cu (unit-level probes of the production stochastic code: parity instruments)
cudaError_t launch_unit_camera(const DevCamera &cam, const DevRender &rp, uint32_t n, const uint32_t *ijs, float *out, cudaStream_t st);
cudaError_t launch_unit_samplers(uint32_t n, const float *u3, float *out, cudaStream_t st);
cudaError_t launch_unit_uniforms(const DevRender &rp, uint32_t n, const uint32_t *psb, float *out, cudaStream_t st);
cudaError_t launch_unit_shade(int variant, const DevScene &sc, const DevRender &rp, uint32_t n, const float *rays, const uint32_t *psb,
                              uint32_t *prim_id, float *out, cudaStream_t st);

// rtw_wavefront.cu (K2: wavefront schedule of the same path)
struct WfCounters {
    uint32_t n_free[2];
    uint32_t pad[2];
    unsigned long long next_path[2];
};
struct WfState {
    float4 *ro, *rd, *beta, *rad;  // (o.xyz, time) (d.xyz, t_hit) (beta.rgb, pixel) (L.rgb, sample<<6|bounce)
    uint32_t *hit;
    uint32_t *free_list[2];
    WfCounters *counters;
    uint32_t *extend_cursor;
    uint32_t n_slots;
    unsigned long long total_paths;
};
cudaError_t wavefront_accumulate(WfState &st, int variant, bool stats, const DevScene &sc, const DevCamera &cam,
                                 const DevRender &rp, int n_sms, WfCounters *h_counters, cudaStream_t stream,
                                 uint32_t *n_launches_out);

// rtw_probe.cu (reference-order arithmetic, compiled with -fmad=false; parity instrument only).
// The raw scene keeps the reference's f64 fields and its nested instance chains.
struct RawXform {
    uint32_t kind;  // RTW_XFORM_*
    uint32_t pad;
    double v[3];    // translate: offset; rotate_y: {sin, cos, -}
};
struct RawPrim {
    uint32_t kind;         // RTW_PRIM_*
    uint32_t material;
    uint32_t chain_begin;  // into the RawXform array, OUTERMOST first
    uint32_t chain_len;
    double v[10];
};
struct RawScene {
    const RawPrim *prims;        // reference order
    const RawXform *chains;
    const BvhNode *nodes;        // same BVH as production
    const uint32_t *bvh_prim_id; // slot -> prim id
    uint32_t n_prims;
    uint32_t root_is_leaf;
};
struct RawCamera {
    double origin[3], horizontal[3], vertical[3], llc[3];
    double time0, time1;
};
// precision 32|64; variant VAR_FLAT|VAR_BVH.  rays: n x 7 doubles, or null with cam/width/height set
// (parity-mode primary rays).  Outputs are doubles regardless of precision.
cudaError_t launch_ref_probe(int precision, int variant, const RawScene &sc, uint32_t n, const double *rays,
                             const RawCamera *cam, uint32_t width, uint32_t height, uint32_t *prim_id, double *t,
                             double *normal, double *uv, cudaStream_t st);

}  // namespace rtw
