// rtw_bvh.cpp — binned-SAH BVH builder, linearised into 32-byte nodes with adjacent siblings.
#include "rtw_bvh.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

namespace rtw {
namespace {

struct Aabb {
    float mn[3], mx[3];
    void reset() {
        for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::infinity(); mx[a] = -mn[a]; }
    }
    void grow(const Aabb &b) {
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], b.mn[a]); mx[a] = std::max(mx[a], b.mx[a]); }
    }
    float half_area() const {
        const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

float down(double v) {  // largest float <= v
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
float up(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

void write_node(BvhNode &n, const Aabb &b, uint32_t a, uint32_t cnt) {
    n.mnx = b.mn[0]; n.mny = b.mn[1]; n.mnz = b.mn[2]; n.a = a;
    n.mxx = b.mx[0]; n.mxy = b.mx[1]; n.mxz = b.mx[2]; n.b = cnt;
}

}  // namespace

BvhResult build_bvh(const std::vector<Box3d> &boxes, uint32_t max_leaf) {
    BvhResult out;
    const uint32_t n = (uint32_t)boxes.size();
    out.order.resize(n);
    std::iota(out.order.begin(), out.order.end(), 0u);
    out.nodes.resize(2);
    write_node(out.nodes[1], Aabb{{0, 0, 0}, {0, 0, 0}}, 0, 0);
    if (n == 0) {
        Aabb e{{0, 0, 0}, {0, 0, 0}};
        write_node(out.nodes[0], e, 0, 0);
        out.root_is_leaf = true;
        return out;
    }
    std::vector<Aabb> pb(n);
    std::vector<float> cen(3 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            pb[i].mn[a] = down(boxes[i].mn[a]);
            pb[i].mx[a] = up(boxes[i].mx[a]);
            cen[3 * (size_t)i + a] = 0.5f * (pb[i].mn[a] + pb[i].mx[a]);
        }

    struct Task { uint32_t node, lo, hi, depth; };
    std::vector<Task> stack;
    stack.push_back({0, 0, n, 1});
    constexpr int kBins = 16;
    constexpr uint32_t kSahDepthLimit = 36;  // deeper than this: median splits (log2 n more levels at most)
    while (!stack.empty()) {
        const Task tk = stack.back();
        stack.pop_back();
        out.depth = std::max(out.depth, tk.depth);
        Aabb bounds, cb;
        bounds.reset(); cb.reset();
        for (uint32_t i = tk.lo; i < tk.hi; ++i) {
            const uint32_t id = out.order[i];
            bounds.grow(pb[id]);
            for (int a = 0; a < 3; ++a) {
                cb.mn[a] = std::min(cb.mn[a], cen[3 * (size_t)id + a]);
                cb.mx[a] = std::max(cb.mx[a], cen[3 * (size_t)id + a]);
            }
        }
        const uint32_t count = tk.hi - tk.lo;
        auto make_leaf = [&]() { write_node(out.nodes[tk.node], bounds, tk.lo, count); };
        if (count == 1) { make_leaf(); continue; }

        // binned SAH
        int best_axis = -1, best_split = -1;
        float best_cost = std::numeric_limits<float>::infinity();
        if (tk.depth < kSahDepthLimit) {
            for (int a = 0; a < 3; ++a) {
                const float ext = cb.mx[a] - cb.mn[a];
                if (!(ext > 0.0f)) continue;
                Aabb bb[kBins];
                uint32_t bc[kBins] = {};
                for (auto &b : bb) b.reset();
                const float k = (float)kBins / ext;
                for (uint32_t i = tk.lo; i < tk.hi; ++i) {
                    const uint32_t id = out.order[i];
                    const int bi = std::min(kBins - 1, (int)((cen[3 * (size_t)id + a] - cb.mn[a]) * k));
                    bb[bi].grow(pb[id]);
                    ++bc[bi];
                }
                float right_area[kBins];
                uint32_t right_cnt[kBins];
                Aabb acc;
                acc.reset();
                uint32_t c = 0;
                for (int b = kBins - 1; b > 0; --b) {
                    if (bc[b]) acc.grow(bb[b]);
                    c += bc[b];
                    right_area[b] = c ? acc.half_area() : 0.0f;
                    right_cnt[b] = c;
                }
                acc.reset();
                c = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    if (bc[b]) acc.grow(bb[b]);
                    c += bc[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    const float cost = acc.half_area() * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
                }
            }
        }
        const float leaf_cost = bounds.half_area() * (float)count;
        if (count <= max_leaf && (best_axis < 0 || best_cost + bounds.half_area() >= leaf_cost)) { make_leaf(); continue; }

        uint32_t mid;
        if (best_axis >= 0) {
            const float ext = cb.mx[best_axis] - cb.mn[best_axis];
            const float k = (float)kBins / ext;
            auto it = std::partition(out.order.begin() + tk.lo, out.order.begin() + tk.hi, [&](uint32_t id) {
                const int bi = std::min(kBins - 1, (int)((cen[3 * (size_t)id + best_axis] - cb.mn[best_axis]) * k));
                return bi <= best_split;
            });
            mid = (uint32_t)(it - out.order.begin());
        } else {
            mid = tk.lo;
        }
        if (mid == tk.lo || mid == tk.hi) {  // degenerate (coincident centroids / depth limit): median by widest axis
            int ax = 0;
            for (int a = 1; a < 3; ++a)
                if (cb.mx[a] - cb.mn[a] > cb.mx[ax] - cb.mn[ax]) ax = a;
            mid = tk.lo + count / 2;
            std::nth_element(out.order.begin() + tk.lo, out.order.begin() + mid, out.order.begin() + tk.hi,
                             [&](uint32_t x, uint32_t y) { return cen[3 * (size_t)x + ax] < cen[3 * (size_t)y + ax]; });
        }
        const uint32_t left = (uint32_t)out.nodes.size();
        out.nodes.emplace_back();
        out.nodes.emplace_back();
        write_node(out.nodes[tk.node], bounds, left, 0);
        stack.push_back({left, tk.lo, mid, tk.depth + 1});
        stack.push_back({left + 1, mid, tk.hi, tk.depth + 1});
    }
    out.root_is_leaf = out.nodes[0].b != 0;
    // keep reference order inside each leaf (ties then resolve like the linear scan without id lookups)
    for (const BvhNode &nd : out.nodes)
        if (nd.b > 1) std::sort(out.order.begin() + nd.a, out.order.begin() + nd.a + nd.b);
    return out;
}

}  // namespace rtw
