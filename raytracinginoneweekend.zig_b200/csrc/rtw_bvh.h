// rtw_bvh.h — host-side BVH construction (new work: the reference scans a list,
// src/rtw/hittable.zig:231-244; its `boudingBox` methods only define the leaf boxes).
#pragma once
#include <stdint.h>

#include <vector>

#include "rtw_device.cuh"

namespace rtw {

struct Box3d {
    double mn[3], mx[3];
};

struct BvhResult {
    std::vector<BvhNode> nodes;   // nodes[0] = root, nodes[1] = padding, children in adjacent pairs
    std::vector<uint32_t> order;  // slot -> primitive id (leaf order)
    uint32_t depth = 0;
    bool root_is_leaf = false;
};

// Binned-SAH binary BVH over `boxes` (one per primitive, reference order).  Leaves hold at most
// `max_leaf` primitives.  Node boxes are fp32, rounded outward from the f64 boxes.
// max_threads = 0: as many as the machine offers (capped at 16)
BvhResult build_bvh(const std::vector<Box3d> &boxes, uint32_t max_leaf = 4, unsigned max_threads = 0);

// The f64 box rounded outward to fp32: out = (min xyz, max xyz).
void box_to_f32(const Box3d &b, float out[6]);

}  // namespace rtw
