// rtw_lbvh.cu — BVH construction on the device: 63-bit Morton codes of the leaf-box centroids, one
// radix sort, Karras' parallel binary radix tree, a bottom-up box fit, and an emit pass that folds
// every subtree of <= max_leaf primitives into one leaf and writes the traversal's node layout
// (rtw_device.cuh) directly, in pre-order.  Leaf boxes are the ones the reference's `boudingBox` methods define
// (src/rtw/hittable.zig:133-143,203-217,305-316,358-369,411-422,491-498,598-603), computed by the caller.
#include "rtw_lbvh.h"

#include <algorithm>
#include <cstdint>

#include <cub/cub.cuh>

namespace rtw {
namespace {

__device__ __forceinline__ uint64_t spread21(uint32_t v) {  // bit i -> bit 3i
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x001f00000000ffffull;
    x = (x | x << 16) & 0x001f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

struct Grid {
    float mnx, mny, mnz, sx, sy, sz;  // q = (c - mn) * s, s = 2^21 / widest extent
};

__global__ void k_morton(const float *__restrict__ boxes, const uint32_t *__restrict__ ids, uint32_t ns, Grid g,
                         uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ns) return;
    const uint32_t id = ids[k];
    const float *b = boxes + 6 * (size_t)id;
    const float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]);
    const float lim = 2097151.0f;
    const uint32_t qx = (uint32_t)fminf(fmaxf((cx - g.mnx) * g.sx, 0.f), lim);
    const uint32_t qy = (uint32_t)fminf(fmaxf((cy - g.mny) * g.sy, 0.f), lim);
    const uint32_t qz = (uint32_t)fminf(fmaxf((cz - g.mnz) * g.sz, 0.f), lim);
    keys[k] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    vals[k] = id;
}

// length of the common prefix of keys i and j, ties broken by position; -1 outside the array
__device__ __forceinline__ int prefix(const uint64_t *__restrict__ keys, int n, int i, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t kj = keys[j];
    return ki == kj ? 64 + __clz(i ^ j) : __clzll((long long)(ki ^ kj));
}

// Karras 2012: internal node i of the radix tree over sorted keys covers [lo, hi] and splits after `gamma`.
__global__ void k_hierarchy(const uint64_t *__restrict__ keys, uint32_t ns, uint2 *__restrict__ range,
                            uint32_t *__restrict__ gamma, uint32_t *__restrict__ parent_int,
                            uint32_t *__restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)ns;
    if (i >= n - 1) return;
    const uint64_t ki = keys[i];
    const int d = prefix(keys, n, i, ki, i + 1) - prefix(keys, n, i, ki, i - 1) >= 0 ? 1 : -1;
    const int dmin = prefix(keys, n, i, ki, i - d);
    int lmax = 2;
    while (prefix(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (prefix(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = prefix(keys, n, i, ki, j);
    int s = 0;
    for (int t = l;;) {
        t = (t + 1) >> 1;
        if (prefix(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int g = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    range[i] = make_uint2((uint32_t)lo, (uint32_t)hi);
    gamma[i] = (uint32_t)g;
    if (lo == g) parent_leaf[g] = (uint32_t)i; else parent_int[g] = (uint32_t)i;
    if (hi == g + 1) parent_leaf[g + 1] = (uint32_t)i; else parent_int[g + 1] = (uint32_t)i;
    if (i == 0) parent_int[0] = 0xFFFFFFFFu;
}

struct Box {
    float mnx, mny, mnz, mxx, mxy, mxz;
};

__device__ __forceinline__ Box prim_box(const float *__restrict__ boxes, uint32_t id) {
    const float *b = boxes + 6 * (size_t)id;
    return Box{b[0], b[1], b[2], b[3], b[4], b[5]};
}

// ibox[2i] = (min, bits(levels)), ibox[2i+1] = (max, bits(pairs)); written once by the second thread to reach node i.
// levels = height of the emitted subtree (a leaf = 1), pairs = sibling pairs it emits (0 when folded into a leaf).
__device__ __forceinline__ Box int_box(const float4 *ibox, uint32_t i, uint32_t &levels, uint32_t &pairs) {
    const float4 a = __ldcg(ibox + 2 * (size_t)i), b = __ldcg(ibox + 2 * (size_t)i + 1);
    levels = __float_as_uint(a.w);
    pairs = __float_as_uint(b.w);
    return Box{a.x, a.y, a.z, b.x, b.y, b.z};
}

__device__ __forceinline__ float half_area(const Box &b) {
    const float dx = b.mxx - b.mnx, dy = b.mxy - b.mny, dz = b.mxz - b.mnz;
    return fmaf(dx, dy, fmaf(dy, dz, dz * dx));
}

// Bottom-up: boxes, and which subtrees fold into one leaf.  A subtree of <= max_leaf primitives folds when the surface
// area heuristic prefers it -- area x count against the node's own area plus the cost of its two (already optimally
// folded) children -- the criterion of the host builder (rtw_bvh.cpp; weighting the primitive test below 1 folds more
// and was slower: 0.75 / 0.5 / 0.35 -> 58.3 / 60.5 / 61.5 ms against 57.8).  survive[i] = 1 for nodes that stay interior.
__global__ void k_fit(const float *__restrict__ boxes, const uint32_t *__restrict__ vals, uint32_t ns, uint32_t max_leaf,
                      const uint2 *__restrict__ range, const uint32_t *__restrict__ gamma,
                      const uint32_t *__restrict__ parent_int, const uint32_t *__restrict__ parent_leaf,
                      unsigned int *__restrict__ arrivals, float4 *ibox, float *icost, uint32_t *__restrict__ survive) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ns) return;
    uint32_t cur = parent_leaf[p];
    while (cur != 0xFFFFFFFFu) {
        if (atomicAdd(arrivals + cur, 1u) == 0u) return;  // the sibling subtree is not finished: its thread continues
        __threadfence();
        const uint2 rg = range[cur];
        const uint32_t g = gamma[cur];
        uint32_t ll = 1, lr = 1, pl = 0, pr = 0;
        const Box L = rg.x == g ? prim_box(boxes, vals[g]) : int_box(ibox, g, ll, pl);
        const Box R = rg.y == g + 1 ? prim_box(boxes, vals[g + 1]) : int_box(ibox, g + 1, lr, pr);
        const Box B{fminf(L.mnx, R.mnx), fminf(L.mny, R.mny), fminf(L.mnz, R.mnz),
                    fmaxf(L.mxx, R.mxx), fmaxf(L.mxy, R.mxy), fmaxf(L.mxz, R.mxz)};
        const uint32_t count = rg.y - rg.x + 1;
        const float cl = rg.x == g ? half_area(L) : __ldcg(icost + g);
        const float cr = rg.y == g + 1 ? half_area(R) : __ldcg(icost + g + 1);
        const float area = half_area(B);
        const float split_cost = area + cl + cr, leaf_cost = area * (float)count;
        const bool folded = count <= max_leaf && !(split_cost < leaf_cost);
        const uint32_t levels = folded ? 1u : 1u + max(ll, lr), pairs = folded ? 0u : 1u + pl + pr;
        icost[cur] = folded ? leaf_cost : split_cost;
        survive[cur] = folded ? 0u : 1u;
        ibox[2 * (size_t)cur] = make_float4(B.mnx, B.mny, B.mnz, __uint_as_float(levels));
        ibox[2 * (size_t)cur + 1] = make_float4(B.mxx, B.mxy, B.mxz, __uint_as_float(pairs));
        __threadfence();
        cur = parent_int[cur];
    }
}

// Pre-order position (in pairs) of the pair each surviving node emits: the left subtree directly behind its parent, the
// right subtree behind the left one -- the layout the host builder produces, so a descent to the left stays in the
// same or the next cache line.  Each node sums its own path to the root (depth <= ~64 steps).
constexpr uint32_t kInsideLeaf = 0xFFFFFFFFu;
__global__ void k_offsets(uint32_t ns, const uint2 *__restrict__ range, const uint32_t *__restrict__ gamma,
                          const uint32_t *__restrict__ parent_int, const uint32_t *__restrict__ survive,
                          const float4 *__restrict__ ibox, uint32_t *__restrict__ pair_index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns - 1 || !survive[i]) return;
    uint32_t off = 0, c = i;
    bool inside_leaf = false;  // an ancestor folded: this node is part of that leaf and emits nothing
    while (c != 0u) {
        const uint32_t p = parent_int[c];
        const uint32_t g = gamma[p];
        inside_leaf |= survive[p] == 0u;
        off += 1u;
        if (c == g + 1u && range[p].x != g) off += __float_as_uint(ibox[2 * (size_t)g + 1].w);  // pairs of the left sibling
        c = p;
    }
    pair_index[i] = inside_leaf ? kInsideLeaf : off;
}

__device__ __forceinline__ void store_node(BvhNode *nodes, uint32_t at, const Box &b, uint32_t a, uint32_t cnt) {
    float4 *q = reinterpret_cast<float4 *>(nodes + at);
    q[0] = make_float4(b.mnx, b.mny, b.mnz, __uint_as_float(a | (cnt << kNodeRefShift)));  // = the traversal's node reference
    q[1] = make_float4(b.mxx, b.mxy, b.mxz, __uint_as_float(cnt));
}

// Every surviving internal node (more than max_leaf primitives) writes its two children side by side.
__global__ void k_emit(const float *__restrict__ boxes, const uint32_t *__restrict__ vals, uint32_t ns,
                       const uint2 *__restrict__ range, const uint32_t *__restrict__ gamma,
                       const uint32_t *__restrict__ survive, const uint32_t *__restrict__ pair_index,
                       const float4 *__restrict__ ibox, BvhNode *__restrict__ nodes, uint32_t root_slot, uint32_t pair_base,
                       uint32_t *__restrict__ order, uint32_t slot_base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns - 1 || !survive[i] || pair_index[i] == kInsideLeaf) return;
    const uint2 rg = range[i];
    const uint32_t g = gamma[i];
    const uint32_t at = pair_base + 2u * pair_index[i];
    uint32_t lv, np;
    if (i == 0) store_node(nodes, root_slot, int_box(ibox, 0, lv, np), at, 0u);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const uint32_t c = g + side;
        const bool single = side == 0 ? rg.x == c : rg.y == c;
        if (single) {
            const uint32_t id = vals[c];
            order[slot_base + c] = id;
            store_node(nodes, at + side, prim_box(boxes, id), slot_base + c, 1u);
            continue;
        }
        const Box b = int_box(ibox, c, lv, np);
        if (survive[c]) {
            store_node(nodes, at + side, b, pair_base + 2u * pair_index[c], 0u);
            continue;
        }
        // folded subtree: one leaf, members in ascending primitive id (the order the reference's scan meets them)
        const uint2 cr = range[c];
        const uint32_t cnt = cr.y - cr.x + 1;
        uint32_t m[16];
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t v = vals[cr.x + k];
            uint32_t q = k;
            while (q > 0 && m[q - 1] > v) { m[q] = m[q - 1]; --q; }
            m[q] = v;
        }
        for (uint32_t k = 0; k < cnt; ++k) order[slot_base + cr.x + k] = m[k];
        store_node(nodes, at + side, b, slot_base + cr.x, cnt);
    }
}

__global__ void k_gather(const DevPrim *__restrict__ src, const uint32_t *__restrict__ order, DevPrim *__restrict__ dst, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 *s = reinterpret_cast<const float4 *>(src + order[k]);
    float4 *d = reinterpret_cast<float4 *>(dst + k);
    d[0] = s[0];
    d[1] = s[1];
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

cudaError_t gather_prims(const DevPrim *src, const uint32_t *d_order, DevPrim *dst, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_gather<<<(n + 255) / 256, 256, 0, st>>>(src, d_order, dst, n);
    return cudaGetLastError();
}

namespace {
struct ArenaLayout {
    size_t keys0, keys1, vals0, vals1, range, gamma, pint, pleaf, surv, pair, arrivals, ibox, icost, tmp, total, sort_bytes;
};
cudaError_t arena_layout(uint32_t ns, ArenaLayout &a) {
    a.sort_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, a.sort_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)ns, 0, 63, 0);
    if (e != cudaSuccess) return e;
    const size_t n = ns;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes); return at; };
    a.keys0 = take(8 * n); a.keys1 = take(8 * n); a.vals0 = take(4 * n); a.vals1 = take(4 * n);
    a.range = take(8 * n); a.gamma = take(4 * n); a.pint = take(4 * n); a.pleaf = take(4 * n);
    a.surv = take(4 * n); a.pair = take(4 * n); a.arrivals = take(4 * n); a.ibox = take(32 * n); a.icost = take(4 * n);
    a.tmp = take(a.sort_bytes);
    a.total = off;
    return cudaSuccess;
}
}  // namespace

size_t lbvh_arena_bytes(uint32_t ns) {
    ArenaLayout a;
    return arena_layout(ns, a) == cudaSuccess ? a.total : 0;
}

cudaError_t build_lbvh(const float *d_boxes, const uint32_t *d_ids, uint32_t ns, const float cmin[3], const float cext[3],
                       uint32_t max_leaf, BvhNode *d_nodes, uint32_t root_slot, uint32_t pair_base, uint32_t *d_order,
                       uint32_t slot_base, void *d_arena, size_t arena_bytes, cudaStream_t st, LbvhInfo *info,
                       const float *axis_weights, int n_candidates) {
    if (ns < 2 || ns <= max_leaf || max_leaf < 1 || max_leaf > 15 || (pair_base & 1u)) return cudaErrorInvalidValue;
    cudaError_t e;
    ArenaLayout L;
    if ((e = arena_layout(ns, L)) != cudaSuccess) return e;
    if (!d_arena || arena_bytes < L.total || (reinterpret_cast<uintptr_t>(d_arena) & 255u)) return cudaErrorInvalidValue;
    const size_t n = ns;
    char *arena = static_cast<char *>(d_arena);
    auto *keys0 = (uint64_t *)(arena + L.keys0), *keys1 = (uint64_t *)(arena + L.keys1);
    auto *vals0 = (uint32_t *)(arena + L.vals0), *vals1 = (uint32_t *)(arena + L.vals1);
    auto *range = (uint2 *)(arena + L.range);
    auto *gamma = (uint32_t *)(arena + L.gamma), *pint = (uint32_t *)(arena + L.pint), *pleaf = (uint32_t *)(arena + L.pleaf);
    auto *surv = (uint32_t *)(arena + L.surv), *pair = (uint32_t *)(arena + L.pair);
    auto *arrivals = (unsigned int *)(arena + L.arrivals);
    auto *ibox = (float4 *)(arena + L.ibox);
    auto *icost = (float *)(arena + L.icost);
    void *tmp = arena + L.tmp;
    const size_t sort_bytes = L.sort_bytes;

    // one scale for the three axes (cubic cells): a thin axis then shares its leading bits across the whole scene and
    // the radix tree simply has no split there, instead of slicing a flat scene into overlapping slabs.  Candidate
    // `axis_weights` (<= 1 per axis) make the cells of an axis longer still; the candidate whose tree has the smallest
    // surface-area cost (k_fit sums it bottom-up: icost[0]) is the one that is emitted.
    const float widest = fmaxf(cext[0], fmaxf(cext[1], cext[2]));
    const float scale = widest > 0.f ? 2097152.0f / widest : 0.f;
    const uint32_t tpb = 256, blocks = (ns + tpb - 1) / tpb;
    uint32_t root_info[2] = {0, 0};  // levels, pairs
    auto done = [&](cudaError_t err) { return err; };
    static const float kCubic[3] = {1.f, 1.f, 1.f};
    if (n_candidates < 1 || !axis_weights) { axis_weights = kCubic; n_candidates = 1; }
    auto fit = [&](const float *w) -> cudaError_t {  // Morton codes -> sort -> radix tree -> boxes, folding, costs
        Grid g;
        g.mnx = cmin[0]; g.mny = cmin[1]; g.mnz = cmin[2];
        g.sx = scale * w[0]; g.sy = scale * w[1]; g.sz = scale * w[2];
        k_morton<<<blocks, tpb, 0, st>>>(d_boxes, d_ids, ns, g, keys0, vals0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        size_t tb = sort_bytes;
        if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys0, keys1, vals0, vals1, (int)ns, 0, 63, st)) != cudaSuccess) return e;
        k_hierarchy<<<blocks, tpb, 0, st>>>(keys1, ns, range, gamma, pint, pleaf);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(arrivals, 0, 4 * n, st)) != cudaSuccess) return e;
        k_fit<<<blocks, tpb, 0, st>>>(d_boxes, vals1, ns, max_leaf, range, gamma, pint, pleaf, arrivals, ibox, icost, surv);
        return cudaGetLastError();
    };
    int best = 0;
    float best_cost = 0.f;
    if (n_candidates > 1) {
        for (int c = 0; c < n_candidates; ++c) {
            float cost = 0.f;
            if ((e = fit(axis_weights + 3 * c)) != cudaSuccess) return done(e);
            if ((e = cudaMemcpyAsync(&cost, icost, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(e);
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
            if (info && c < LbvhInfo::kMaxCandidates) info->cost[c] = cost;
            if (c == 0 || cost < best_cost) { best = c; best_cost = cost; }
        }
    }
    if (n_candidates == 1 || best != n_candidates - 1)
        if ((e = fit(axis_weights + 3 * best)) != cudaSuccess) return done(e);
    k_offsets<<<blocks, tpb, 0, st>>>(ns, range, gamma, pint, surv, ibox, pair);
    if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    k_emit<<<blocks, tpb, 0, st>>>(d_boxes, vals1, ns, range, gamma, surv, pair, ibox, d_nodes, root_slot, pair_base, d_order, slot_base);
    if ((e = cudaGetLastError()) != cudaSuccess) return done(e);
    if ((e = cudaMemcpyAsync(&root_info[0], reinterpret_cast<const char *>(ibox) + 12, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(e);
    if ((e = cudaMemcpyAsync(&root_info[1], reinterpret_cast<const char *>(ibox) + 28, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return done(e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(e);
    if (info) {
        info->chosen = best;
        info->n_nodes = pair_base + 2u * root_info[1];
        info->depth = root_info[0];
    }
    return done(cudaSuccess);
}

}  // namespace rtw
