// rt_rebuild.hpp — SURVEY §8 row f4: a second, faster topology over the SAME elements (host-side builder).
//
// The reference tree (median split on a round-robin axis, src/lib/hittable.rs:382-427) stays the definition of the
// result.  This tree is only an accelerator: the fast traversal finds the element with the smallest accepted t over
// ALL elements, and the result is used only when the reference provably returns the same element (unique minimum and
// the hit point strictly inside the box of the element's parent in the reference tree, see ref_visible in
// rt_device.cuh); every other ray is traced again on the reference tree.
//
// Builder: top-down surface-area heuristic over the elements' own bounding boxes (Geometry::bounding_box), exact sweep
// for small ranges, 32 centroid bins above; one element per leaf, pre-order numbering (children have larger indices);
// depth is capped (the traversal stack lives in shared memory), falling back to a median split where the cap binds.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace rt_rebuild {

struct Box3 {
    float mn[3], mx[3];
};

struct Tree {
    std::vector<uint32_t> left, right;  // child: 0x80000000 | element, or node index
    std::vector<uint32_t> axis;         // split axis: the left child holds the lower centroids on it
    std::vector<Box3> box;              // each node's own box (exact union of the element boxes below it)
    uint32_t root = 0;
    uint32_t depth = 0;
};

namespace detail {

inline Box3 empty_box() {
    Box3 b;
    for (int k = 0; k < 3; ++k) { b.mn[k] = INFINITY; b.mx[k] = -INFINITY; }
    return b;
}
inline void grow(Box3& b, const Box3& o) {
    for (int k = 0; k < 3; ++k) { b.mn[k] = std::fmin(b.mn[k], o.mn[k]); b.mx[k] = std::fmax(b.mx[k], o.mx[k]); }
}
inline double half_area(const Box3& b) {
    const double x = std::max(0.0, (double)b.mx[0] - b.mn[0]), y = std::max(0.0, (double)b.mx[1] - b.mn[1]), z = std::max(0.0, (double)b.mx[2] - b.mn[2]);
    return x * y + y * z + z * x;
}

struct Builder {
    const std::vector<Box3>& leaf;
    Tree& out;
    std::vector<uint32_t> idx;
    std::vector<float> cen[3];
    uint32_t max_depth;
    static const int BINS = 32;
    static const uint32_t SWEEP_MAX = 96;

    Builder(const std::vector<Box3>& l, Tree& t, uint32_t md) : leaf(l), out(t), max_depth(md) {
        idx.resize(leaf.size());
        std::iota(idx.begin(), idx.end(), 0u);
        for (int k = 0; k < 3; ++k) {
            cen[k].resize(leaf.size());
            for (size_t i = 0; i < leaf.size(); ++i) cen[k][i] = 0.5f * leaf[i].mn[k] + 0.5f * leaf[i].mx[k];
        }
    }

    // returns the split position (lo < mid < hi) after reordering idx[lo, hi)
    uint32_t split(uint32_t lo, uint32_t hi, uint32_t levels_left, uint32_t* axis_out) {
        const uint32_t n = hi - lo;
        const uint64_t child_cap = levels_left >= 1 ? (levels_left - 1 >= 63 ? ~0ull : (1ull << (levels_left - 1))) : 1ull;
        float cmin[3], cmax[3];
        for (int k = 0; k < 3; ++k) { cmin[k] = INFINITY; cmax[k] = -INFINITY; }
        for (uint32_t i = lo; i < hi; ++i)
            for (int k = 0; k < 3; ++k) { cmin[k] = std::fmin(cmin[k], cen[k][idx[i]]); cmax[k] = std::fmax(cmax[k], cen[k][idx[i]]); }
        int best_axis = -1;
        uint32_t best_pos = 0;  // sweep: number of elements on the left; bins: last left bin
        double best_cost = INFINITY;
        if (n <= SWEEP_MAX) {
            std::vector<uint32_t> ord(n);
            std::vector<double> right_area(n);
            for (int k = 0; k < 3; ++k) {
                if (!(cmax[k] > cmin[k])) continue;
                std::copy(idx.begin() + lo, idx.begin() + hi, ord.begin());
                std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return cen[k][a] < cen[k][b]; });
                Box3 acc = empty_box();
                for (uint32_t i = n; i-- > 1;) { grow(acc, leaf[ord[i]]); right_area[i] = half_area(acc); }
                acc = empty_box();
                for (uint32_t i = 1; i < n; ++i) {
                    grow(acc, leaf[ord[i - 1]]);
                    const double c = half_area(acc) * i + right_area[i] * (n - i);
                    if (c < best_cost && i <= child_cap && n - i <= child_cap) { best_cost = c; best_axis = k; best_pos = i; }
                }
            }
            if (best_axis >= 0) {
                const int k = best_axis;
                std::stable_sort(idx.begin() + lo, idx.begin() + hi, [&](uint32_t a, uint32_t b) { return cen[k][a] < cen[k][b]; });
                *axis_out = (uint32_t)k;
                return lo + best_pos;
            }
        } else {
            for (int k = 0; k < 3; ++k) {
                if (!(cmax[k] > cmin[k]) || !std::isfinite(cmax[k] - cmin[k])) continue;
                const float scale = (float)BINS / (cmax[k] - cmin[k]);
                Box3 bb[BINS];
                uint32_t cnt[BINS];
                for (int b = 0; b < BINS; ++b) { bb[b] = empty_box(); cnt[b] = 0; }
                for (uint32_t i = lo; i < hi; ++i) {
                    int b = (int)((cen[k][idx[i]] - cmin[k]) * scale);
                    b = std::min(BINS - 1, std::max(0, b));
                    grow(bb[b], leaf[idx[i]]);
                    ++cnt[b];
                }
                double ra[BINS];
                uint32_t rc[BINS];
                Box3 acc = empty_box();
                uint32_t c = 0;
                for (int b = BINS - 1; b >= 1; --b) { grow(acc, bb[b]); c += cnt[b]; ra[b] = half_area(acc); rc[b] = c; }
                acc = empty_box();
                c = 0;
                for (int b = 0; b < BINS - 1; ++b) {
                    grow(acc, bb[b]);
                    c += cnt[b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    const double cost = half_area(acc) * c + ra[b + 1] * rc[b + 1];
                    if (cost < best_cost && c <= child_cap && rc[b + 1] <= child_cap) { best_cost = cost; best_axis = k; best_pos = (uint32_t)b; }
                }
            }
            if (best_axis >= 0) {
                const int k = best_axis;
                const float scale = (float)BINS / (cmax[k] - cmin[k]);
                const float c0 = cmin[k];
                auto it = std::partition(idx.begin() + lo, idx.begin() + hi, [&](uint32_t a) {
                    int b = (int)((cen[k][a] - c0) * scale);
                    b = std::min(BINS - 1, std::max(0, b));
                    return (uint32_t)b <= best_pos;
                });
                const uint32_t mid = (uint32_t)(it - idx.begin());
                if (mid > lo && mid < hi) { *axis_out = (uint32_t)k; return mid; }
            }
        }
        // no useful plane (coincident centroids, NaNs, or the depth cap binds): median split on the widest centroid axis
        int k = 0;
        for (int a = 1; a < 3; ++a)
            if (cmax[a] - cmin[a] > cmax[k] - cmin[k]) k = a;
        const uint32_t mid = lo + n / 2;
        std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](uint32_t a, uint32_t b) { return cen[k][a] < cen[k][b]; });
        *axis_out = (uint32_t)k;
        return mid;
    }

    uint32_t build(uint32_t lo, uint32_t hi, uint32_t depth) {
        if (hi - lo == 1) return 0x80000000u | idx[lo];
        const uint32_t me = (uint32_t)out.left.size();
        out.left.push_back(0);
        out.right.push_back(0);
        out.box.push_back(empty_box());
        out.axis.push_back(0);
        out.depth = std::max(out.depth, depth + 1);
        uint32_t ax = 0;
        const uint32_t mid = split(lo, hi, max_depth - depth, &ax);
        out.axis[me] = ax;
        const uint32_t l = build(lo, mid, depth + 1);
        const uint32_t r = build(mid, hi, depth + 1);
        out.left[me] = l;
        out.right[me] = r;
        Box3 b = empty_box();
        for (uint32_t c : {l, r}) grow(b, (c & 0x80000000u) ? leaf[c & 0x7FFFFFFFu] : out.box[c]);
        out.box[me] = b;
        return me;
    }
};

}  // namespace detail

// max_depth: number of internal levels allowed (>= ceil(log2 n)).
inline void build(const std::vector<Box3>& leaf, uint32_t max_depth, Tree& out) {
    out = Tree();
    uint32_t need = 0;
    while ((1ull << need) < leaf.size()) ++need;
    detail::Builder b(leaf, out, std::max(max_depth, need));
    out.left.reserve(leaf.size());
    out.right.reserve(leaf.size());
    out.box.reserve(leaf.size());
    out.root = b.build(0, (uint32_t)leaf.size(), 0);
}

}  // namespace rt_rebuild
