// rtw_bvh.cpp — binned-SAH BVH builder, linearised into 32-byte nodes with adjacent siblings.
#include "rtw_bvh.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <numeric>
#include <thread>

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
    n.mnx = b.mn[0]; n.mny = b.mn[1]; n.mnz = b.mn[2]; n.a = a | (cnt << kNodeRefShift);  // = the traversal's node reference
    n.mxx = b.mx[0]; n.mxy = b.mx[1]; n.mxz = b.mx[2]; n.b = cnt;
}

}  // namespace

void box_to_f32(const Box3d &b, float out[6]) {
    for (int a = 0; a < 3; ++a) { out[a] = down(b.mn[a]); out[3 + a] = up(b.mx[a]); }
}

BvhResult build_bvh(const std::vector<Box3d> &boxes, uint32_t max_leaf, unsigned max_threads) {
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

    // Nodes are preallocated (a binary tree over n leaves has < 2n nodes; +2 for the root slot and its pad) and pairs
    // are handed out from per-thread chunks of an atomic cursor, so subtrees built by one thread stay contiguous.
    unsigned n_threads = 1;
    constexpr uint32_t kSerialBelow = 8192;  // subtrees smaller than this are finished by the thread that reaches them
    if (n >= 4 * kSerialBelow) {
        n_threads = max_threads ? max_threads : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    }
    const uint32_t chunk_pairs = n_threads > 1 ? 512u : 1u;
    out.nodes.resize(2 * (size_t)n + 2 + (size_t)n_threads * 2 * chunk_pairs);
    std::atomic<uint32_t> node_cursor{2};
    std::atomic<uint32_t> max_depth{1};
    constexpr int kBins = 16;
    constexpr uint32_t kSahDepthLimit = 36;  // deeper than this: median splits (log2 n more levels at most)

    struct Task { uint32_t node, lo, hi, depth; };
    struct Alloc { uint32_t next = 0, end = 0; };
    auto alloc_pair = [&](Alloc &al) {
        if (al.next == al.end) { al.next = node_cursor.fetch_add(2 * chunk_pairs); al.end = al.next + 2 * chunk_pairs; }
        const uint32_t at = al.next;
        al.next += 2;
        return at;
    };
    // Splits one node; returns false for a leaf, else fills the two child tasks.
    auto split = [&](const Task &tk, Alloc &al, Task &lt, Task &rt) -> bool {
        uint32_t d = max_depth.load(std::memory_order_relaxed);
        while (tk.depth > d && !max_depth.compare_exchange_weak(d, tk.depth)) {}
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
        if (count == 1) { make_leaf(); return false; }

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
        if (count <= max_leaf && (best_axis < 0 || best_cost + bounds.half_area() >= leaf_cost)) { make_leaf(); return false; }

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
        const uint32_t left = alloc_pair(al);
        write_node(out.nodes[tk.node], bounds, left, 0);
        lt = Task{left, tk.lo, mid, tk.depth + 1};
        rt = Task{left + 1, mid, tk.hi, tk.depth + 1};
        return true;
    };
    auto build_serial = [&](const Task &root, Alloc &al) {  // depth-first, private stack
        std::vector<Task> st{root};
        while (!st.empty()) {
            const Task tk = st.back();
            st.pop_back();
            Task l, r;
            if (split(tk, al, l, r)) { st.push_back(l); st.push_back(r); }
        }
    };

    if (n_threads <= 1) {
        Alloc al;
        build_serial(Task{0, 0, n, 1}, al);
    } else {
        // shared queue of large subtrees; a worker finishes small subtrees itself
        std::mutex mu;
        std::condition_variable cv;
        std::vector<Task> queue{Task{0, 0, n, 1}};
        unsigned busy = 0;
        bool done = false;
        auto worker = [&]() {
            Alloc al;
            for (;;) {
                Task tk;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return done || !queue.empty(); });
                    if (queue.empty()) return;
                    tk = queue.back();
                    queue.pop_back();
                    ++busy;
                }
                if (tk.hi - tk.lo < kSerialBelow) {
                    build_serial(tk, al);
                } else {
                    Task l, r;
                    if (split(tk, al, l, r)) {
                        std::lock_guard<std::mutex> lk(mu);
                        queue.push_back(l);
                        queue.push_back(r);
                        cv.notify_all();
                    }
                }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    --busy;
                    if (busy == 0 && queue.empty()) { done = true; cv.notify_all(); }
                }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(worker);
        for (auto &t : pool) t.join();
    }
    out.depth = max_depth.load();
    // unused tail of the preallocation (chunks leave holes: harmless, never referenced) is trimmed to the cursor
    out.nodes.resize(std::min<size_t>(out.nodes.size(), node_cursor.load()));
    out.root_is_leaf = out.nodes[0].b != 0;
    // keep reference order inside each leaf (ties then resolve like the linear scan without id lookups)
    for (const BvhNode &nd : out.nodes)
        if (nd.b > 1) std::sort(out.order.begin() + (nd.a & kNodeRefIndexMask), out.order.begin() + (nd.a & kNodeRefIndexMask) + nd.b);
    return out;
}

}  // namespace rtw
