// rtw_lbvh.h — device-side BVH construction for large scenes (new work: the
// reference scans a list, src/rtw/hittable.zig:231-244, and its `boudingBox` methods only define
// the leaf boxes).  Morton order + Karras' parallel hierarchy, emitted straight into the node
// layout the traversal reads (rtw_device.cuh: 32-byte nodes, siblings adjacent, <= max_leaf
// primitives per leaf), so a million-sphere scene is ready in milliseconds instead of the
// host builder's fraction of a second.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rtw_device.cuh"

namespace rtw {

struct LbvhInfo {
    uint32_t n_nodes = 0;  // one past the last node index written
    uint32_t depth = 0;    // levels below and including the subtree root (root = 1)
    static constexpr int kMaxCandidates = 8;
    int chosen = 0;                    // index of the Morton-grid candidate that was emitted
    float cost[kMaxCandidates] = {};   // surface-area cost of each candidate's tree (only filled when there are several)
};

// Builds a BVH over the `ns` primitives listed in d_ids (ns >= 2, ns > max_leaf).
//   d_boxes   6 floats per primitive id (min xyz, max xyz), already rounded outward
//   cmin/cext centroid bounds of those primitives (the Morton grid)
//   d_nodes   the subtree root goes to d_nodes[root_slot], child pairs from pair_base (even) on;
//             capacity needed: pair_base + 2*(ns-1)
//   d_order   leaf slot -> primitive id, written at [slot_base, slot_base + ns)
//   d_arena   work memory of at least lbvh_arena_bytes(ns) bytes, 256-byte aligned (the caller keeps it between builds)
size_t lbvh_arena_bytes(uint32_t ns);
cudaError_t build_lbvh(const float *d_boxes, const uint32_t *d_ids, uint32_t ns, const float cmin[3], const float cext[3],
                       uint32_t max_leaf, BvhNode *d_nodes, uint32_t root_slot, uint32_t pair_base, uint32_t *d_order,
                       uint32_t slot_base, void *d_arena, size_t arena_bytes, cudaStream_t st, LbvhInfo *info,
                       const float *axis_weights = nullptr, int n_candidates = 0);
//   axis_weights  n_candidates x 3 factors (<= 1) on the cubic Morton grid's resolution per axis; every candidate's tree is fitted
//                 (0.6 ms per 10^6 primitives) and the one with the smallest surface-area cost is emitted; none = the cubic grid

// dst[k] = src[order[k]] for k < n (primitives into leaf order)
cudaError_t gather_prims(const DevPrim *src, const uint32_t *d_order, DevPrim *dst, uint32_t n, cudaStream_t st);

}  // namespace rtw
