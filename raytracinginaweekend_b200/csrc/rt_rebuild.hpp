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
#include <atomic>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

namespace rt_rebuild {

// Static chunking over hardware threads; fn(begin, end) works on disjoint index ranges, so results do not depend on the thread count.
template <class F>
inline void parallel_for(size_t n, size_t min_chunk, F&& fn) {
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t parts = std::max<size_t>(1, std::min(hw, n / std::max<size_t>(1, min_chunk)));
    if (parts <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    for (size_t p = 1; p < parts; ++p) pool.emplace_back([&, p] { fn(n * p / parts, n * (p + 1) / parts); });
    fn((size_t)0, n / parts);
    for (auto& t : pool) t.join();
}

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
    for (int k = 0; k < 3; ++k) {  // plain comparisons (std::fmin is a libm call at -O2); NaN boxes are rejected upstream
        b.mn[k] = o.mn[k] < b.mn[k] ? o.mn[k] : b.mn[k];
        b.mx[k] = o.mx[k] > b.mx[k] ? o.mx[k] : b.mx[k];
    }
}
inline double half_area(const Box3& b) {
    const double x = std::max(0.0, (double)b.mx[0] - b.mn[0]), y = std::max(0.0, (double)b.mx[1] - b.mn[1]), z = std::max(0.0, (double)b.mx[2] - b.mn[2]);
    return x * y + y * z + z * x;
}

struct Builder {
    const std::vector<Box3>& leaf;
    Tree& out;
    std::vector<uint32_t> own_idx;
    std::vector<float> own_cen[3];
    std::vector<uint32_t>& idx;        // element order (partitioned in place); shared between the builders of one build()
    const std::vector<float>* cen;     // [3] centroids
    uint32_t max_depth;
    static const int BINS = 32;
    static const uint32_t SWEEP_MAX = 96;
    static const uint32_t PAR_SPLIT_MIN = 1u << 18;

    Builder(const std::vector<Box3>& l, Tree& t, uint32_t md) : leaf(l), out(t), idx(own_idx), cen(own_cen), max_depth(md) {
        own_idx.resize(leaf.size());
        std::iota(own_idx.begin(), own_idx.end(), 0u);
        for (int k = 0; k < 3; ++k) own_cen[k].resize(leaf.size());
        parallel_for(leaf.size(), 65536, [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i)
                for (int k = 0; k < 3; ++k) own_cen[k][i] = 0.5f * leaf[i].mn[k] + 0.5f * leaf[i].mx[k];
        });
    }
    Builder(const std::vector<Box3>& l, Tree& t, uint32_t md, std::vector<uint32_t>* shared_idx, const std::vector<float>* shared_cen)
        : leaf(l), out(t), idx(*shared_idx), cen(shared_cen), max_depth(md) {}

    // returns the split position (lo < mid < hi) after reordering idx[lo, hi)
    uint32_t split(uint32_t lo, uint32_t hi, uint32_t levels_left, uint32_t* axis_out) {
        const uint32_t n = hi - lo;
        const uint64_t child_cap = levels_left >= 1 ? (levels_left - 1 >= 63 ? ~0ull : (1ull << (levels_left - 1))) : 1ull;
        float cmin[3], cmax[3];
        for (int k = 0; k < 3; ++k) { cmin[k] = INFINITY; cmax[k] = -INFINITY; }
        const bool big = n >= PAR_SPLIT_MIN;  // the first few levels of a large build: scan and bin on all cores
        if (big) {
            std::mutex mu;
            parallel_for(n, 65536, [&](size_t a, size_t b) {
                float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (size_t i = lo + a; i < lo + b; ++i)
                    for (int k = 0; k < 3; ++k) {
                        const float c = cen[k][idx[i]];
                        mn[k] = c < mn[k] ? c : mn[k];
                        mx[k] = c > mx[k] ? c : mx[k];
                    }
                std::lock_guard<std::mutex> lk(mu);
                for (int k = 0; k < 3; ++k) { cmin[k] = std::min(cmin[k], mn[k]); cmax[k] = std::max(cmax[k], mx[k]); }
            });
        } else {
            for (uint32_t i = lo; i < hi; ++i)
                for (int k = 0; k < 3; ++k) {
                    const float c = cen[k][idx[i]];
                    cmin[k] = c < cmin[k] ? c : cmin[k];
                    cmax[k] = c > cmax[k] ? c : cmax[k];
                }
        }
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
                if (big) {  // per-thread bins, merged (box union and counts are order-independent: same result as the serial loop)
                    std::mutex mu;
                    parallel_for(n, 65536, [&](size_t a, size_t e) {
                        Box3 lb[BINS];
                        uint32_t lc[BINS];
                        for (int b = 0; b < BINS; ++b) { lb[b] = empty_box(); lc[b] = 0; }
                        for (size_t i = lo + a; i < lo + e; ++i) {
                            int b = (int)((cen[k][idx[i]] - cmin[k]) * scale);
                            b = std::min(BINS - 1, std::max(0, b));
                            grow(lb[b], leaf[idx[i]]);
                            ++lc[b];
                        }
                        std::lock_guard<std::mutex> lk(mu);
                        for (int b = 0; b < BINS; ++b) { if (lc[b]) grow(bb[b], lb[b]); cnt[b] += lc[b]; }
                    });
                } else
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

// ------------------------------------------------------------------------------------------------
// Reference pre-splitting.  The reference's OBJ loader fans every polygon around the mesh origin (Q4), so half of a mesh's
// triangles are long slivers whose boxes all overlap around one point; a tree over such boxes cannot separate them.
// Because this tree is only an accelerator, an element may appear in several leaves: a triangle's reference is cut at the
// middle of its box's longest axis (exact polygon clipping in double, boxes rounded outward to float) as long as the two
// clipped boxes together are < 0.9x the area of the unsplit one, largest boxes first, until `budget` references exist.
// Every point of the triangle stays inside the box of at least one of its references, which is all the traversal needs.
// ------------------------------------------------------------------------------------------------
struct Ref {
    uint32_t item;
    Box3 box;
};

namespace detail {
struct Poly {
    int n = 0;
    double v[12][3];
};
inline Poly clip_half(const Poly& in, int axis, double c, double sign) {  // keep sign * (x[axis] - c) >= 0
    Poly out;
    for (int i = 0; i < in.n; ++i) {
        const double* a = in.v[i];
        const double* b = in.v[(i + 1) % in.n];
        const double da = sign * (a[axis] - c), db = sign * (b[axis] - c);
        if (da >= 0.0 && out.n < 12) { for (int k = 0; k < 3; ++k) out.v[out.n][k] = a[k]; ++out.n; }
        if ((da >= 0.0) != (db >= 0.0) && out.n < 12) {
            const double t = da / (da - db);
            for (int k = 0; k < 3; ++k) out.v[out.n][k] = a[k] + t * (b[k] - a[k]);
            out.v[out.n][axis] = c;
            ++out.n;
        }
    }
    return out;
}
inline float round_down(double v) {
    float f = (float)v;
    return (double)f > v ? std::nextafter(f, -INFINITY) : f;
}
inline float round_up(double v) {
    float f = (float)v;
    return (double)f < v ? std::nextafter(f, INFINITY) : f;
}
inline Box3 poly_box(const Poly& p, const Box3& within) {
    Box3 b = empty_box();
    for (int k = 0; k < 3; ++k) {
        double lo = p.v[0][k], hi = p.v[0][k];
        for (int i = 1; i < p.n; ++i) { lo = p.v[i][k] < lo ? p.v[i][k] : lo; hi = p.v[i][k] > hi ? p.v[i][k] : hi; }
        const float flo = round_down(lo), fhi = round_up(hi);
        b.mn[k] = flo > within.mn[k] ? flo : within.mn[k];  // never beyond the box being split
        b.mx[k] = fhi < within.mx[k] ? fhi : within.mx[k];
    }
    return b;
}
}  // namespace detail

// tri: 9 floats (p0, p1, p2) per element, only read where is_tri[i] != 0.
// Largest boxes first, in ROUNDS: each round cuts every candidate whose box area is at least tau, in parallel, and feeds the
// pieces back as candidates; tau halves when nothing that large is left.  A piece is always smaller than the reference it came
// from (accepted cuts shrink the summed area to < 0.9x), so this visits the references in the order of a sequential "pop the
// largest" loop up to a factor 2 in area, and the budget ends on the largest candidates of the last round.  A reference carries no polygon: its piece of the triangle is re-derived on demand by
// clipping the triangle against the reference's REGION (the axis-aligned cell its cuts have carved out: 48 B instead of a
// 300 B polygon — the sequential polygon-copying version was memory-bound at ~0.9 s for the 10^6 triangles of C4).
// Deterministic for any thread count.
inline void presplit(const std::vector<Box3>& leaf, const std::vector<float>& tri, const std::vector<uint8_t>& is_tri, size_t budget,
                     std::vector<Ref>& refs) {
    using detail::Poly;
    refs.clear();
    refs.reserve(std::max(budget, leaf.size()));
    for (uint32_t i = 0; i < leaf.size(); ++i) refs.push_back(Ref{i, leaf[i]});
    if (budget <= leaf.size()) return;
    struct Region { double lo[3], hi[3]; };
    std::vector<Region> region;
    region.reserve(budget);
    region.resize(leaf.size(), Region{{-INFINITY, -INFINITY, -INFINITY}, {INFINITY, INFINITY, INFINITY}});
    struct Cand { double area; uint32_t ref; };
    std::vector<Cand> cand;
    for (uint32_t i = 0; i < leaf.size(); ++i)
        if (is_tri[i]) cand.push_back({detail::half_area(leaf[i]), i});
    struct Cut { Box3 bl, bh; double mid; int ax; bool ok; };
    std::vector<Cut> cuts;
    double tau = 0.0;  // area threshold of the current round: everything at least this large is cut now
    for (const Cand& c : cand) tau = std::max(tau, c.area);
    tau *= 0.5;
    while (!cand.empty() && refs.size() < budget && tau > 0.0) {
        const size_t room = budget - refs.size();
        // candidates of this round at the END of the array (they are consumed, their pieces appended: no copying of the rest)
        auto first_big = std::partition(cand.begin(), cand.end(), [&](const Cand& c) { return c.area < tau; });
        size_t k = (size_t)(cand.end() - first_big);
        if (k == 0) { tau *= 0.5; continue; }
        if (k > room) {  // the budget ends inside this round: the `room` largest of them (ties by reference index: deterministic)
            auto larger = [](const Cand& a, const Cand& b) { return a.area > b.area || (a.area == b.area && a.ref < b.ref); };
            std::nth_element(first_big, first_big + room, cand.end(), larger);
            cand.resize((size_t)(first_big - cand.begin()) + room);
            k = room;
        }
        const size_t base = cand.size() - k;
        cuts.resize(k);
        parallel_for(k, 2048, [&](size_t lo_i, size_t hi_i) {
            for (size_t c = lo_i; c < hi_i; ++c) {
                Cut& cu = cuts[c];
                cu.ok = false;
                const uint32_t r = cand[base + c].ref;
                const Box3 b = refs[r].box;
                int ax = 0;
                for (int q = 1; q < 3; ++q)
                    if (b.mx[q] - b.mn[q] > b.mx[ax] - b.mn[ax]) ax = q;
                if (!(b.mx[ax] > b.mn[ax])) continue;
                const double mid = 0.5 * ((double)b.mn[ax] + (double)b.mx[ax]);
                // this reference's piece of the triangle = triangle clipped against its region
                Poly src;
                src.n = 3;
                for (int v = 0; v < 3; ++v)
                    for (int q = 0; q < 3; ++q) src.v[v][q] = tri[9 * (size_t)refs[r].item + 3 * v + q];
                const Region& rg = region[r];
                for (int q = 0; q < 3 && src.n >= 3; ++q) {
                    if (rg.lo[q] > -INFINITY) src = detail::clip_half(src, q, rg.lo[q], 1.0);
                    if (rg.hi[q] < INFINITY && src.n >= 3) src = detail::clip_half(src, q, rg.hi[q], -1.0);
                }
                if (src.n < 3) continue;
                const Poly lo = detail::clip_half(src, ax, mid, -1.0), hi = detail::clip_half(src, ax, mid, 1.0);
                if (lo.n < 3 || hi.n < 3) continue;
                cu.bl = detail::poly_box(lo, b);
                cu.bh = detail::poly_box(hi, b);
                const double a0 = detail::half_area(b), al = detail::half_area(cu.bl), ah = detail::half_area(cu.bh);
                cu.ok = al + ah < 0.9 * a0;
                cu.mid = mid; cu.ax = ax;
            }
        });
        // apply in batch order (sequential, cheap): the low half replaces the reference, the high half is appended; both pieces
        // are smaller than the reference they came from and become candidates again
        std::vector<uint32_t> batch(k);
        for (size_t c = 0; c < k; ++c) batch[c] = cand[base + c].ref;
        cand.resize(base);
        for (size_t c = 0; c < k; ++c) {
            if (!cuts[c].ok || refs.size() >= budget) continue;
            const uint32_t r = batch[c];
            refs[r].box = cuts[c].bl;
            const uint32_t r2 = (uint32_t)refs.size();
            refs.push_back(Ref{refs[r].item, cuts[c].bh});
            Region hi_region = region[r];
            region[r].hi[cuts[c].ax] = cuts[c].mid;
            hi_region.lo[cuts[c].ax] = cuts[c].mid;
            region.push_back(hi_region);
            cand.push_back({detail::half_area(cuts[c].bl), r});
            cand.push_back({detail::half_area(cuts[c].bh), r2});
        }
    }
}

// max_depth: number of internal levels allowed (>= ceil(log2 n)).  Leaves are indices into `leaf` (boxes of references).
// Ranges larger than n / (4 x threads) are split on the calling thread; the subtrees below are built concurrently (they work
// on disjoint ranges of the index array) and appended, so every child still has a larger index than its parent.
inline void build(const std::vector<Box3>& leaf, uint32_t max_depth, Tree& out) {
    out = Tree();
    uint32_t need = 0;
    while ((1ull << need) < leaf.size()) ++need;
    max_depth = std::max(max_depth, need);
    const unsigned hw = std::min(std::max(1u, std::thread::hardware_concurrency()), leaf.size() < 100000 ? 8u : 32u);
    const uint32_t PAR_MIN = (uint32_t)std::max<size_t>(1024, leaf.size() / (4 * (size_t)hw));
    if (leaf.size() < 4096 || hw < 2) {
        detail::Builder b(leaf, out, max_depth);
        out.left.reserve(leaf.size());
        out.right.reserve(leaf.size());
        out.box.reserve(leaf.size());
        out.root = b.build(0, (uint32_t)leaf.size(), 0);
        return;
    }
    const uint32_t TASK_BIT = 0x40000000u;
    struct Task {
        uint32_t lo, hi, depth;
        Tree tree;
    };
    std::vector<Task> tasks;
    detail::Builder top(leaf, out, max_depth);
    // Top part, phase 1: only the SPLITS of the ranges above PAR_MIN, fork-join (the two halves of a split work on disjoint ranges of
    // the index array, so they run on two threads down to FORK_LEVELS; a split of >= 2^18 references also bins on all cores).
    // The result of a split depends on its range alone, so the tree is the same for any thread count.
    struct TopNode {
        uint32_t lo = 0, hi = 0, depth = 0, mid = 0, axis = 0;
        bool task = false;
        std::unique_ptr<TopNode> l, r;
    };
    const uint32_t FORK_LEVELS = hw >= 16 ? 4u : (hw >= 4 ? 3u : (hw >= 2 ? 1u : 0u));
    struct Splitter {
        detail::Builder& b;
        uint32_t par_min, fork_levels;
        void go(TopNode* nd) {
            if (nd->hi - nd->lo == 1) return;
            if (nd->hi - nd->lo <= par_min) { nd->task = true; return; }
            nd->mid = b.split(nd->lo, nd->hi, b.max_depth - nd->depth, &nd->axis);
            nd->l.reset(new TopNode); nd->r.reset(new TopNode);
            nd->l->lo = nd->lo; nd->l->hi = nd->mid; nd->r->lo = nd->mid; nd->r->hi = nd->hi;
            nd->l->depth = nd->r->depth = nd->depth + 1;
            if (nd->depth < fork_levels && nd->l->hi - nd->l->lo > par_min && nd->r->hi - nd->r->lo > par_min) {
                std::thread other([&] { go(nd->l.get()); });
                go(nd->r.get());
                other.join();
            } else {
                go(nd->l.get());
                go(nd->r.get());
            }
        }
    } splitter{top, PAR_MIN, FORK_LEVELS};
    TopNode top_root;
    top_root.lo = 0; top_root.hi = (uint32_t)leaf.size(); top_root.depth = 0;
    splitter.go(&top_root);
    // phase 2: number the top nodes in pre-order and collect the subtree tasks left to right (sequential, a few hundred nodes)
    struct Numberer {
        detail::Builder& b;
        std::vector<Task>& tasks;
        uint32_t task_bit;
        uint32_t go(const TopNode* nd) {
            if (nd->hi - nd->lo == 1) return 0x80000000u | b.idx[nd->lo];
            if (nd->task) {
                tasks.push_back(Task{nd->lo, nd->hi, nd->depth, Tree()});
                return task_bit | (uint32_t)(tasks.size() - 1);
            }
            const uint32_t me = (uint32_t)b.out.left.size();
            b.out.left.push_back(0); b.out.right.push_back(0); b.out.box.push_back(detail::empty_box()); b.out.axis.push_back(nd->axis);
            b.out.depth = std::max(b.out.depth, nd->depth + 1);
            const uint32_t l = go(nd->l.get()), r = go(nd->r.get());
            b.out.left[me] = l; b.out.right[me] = r;
            return me;
        }
    } numberer{top, tasks, TASK_BIT};
    const uint32_t root = numberer.go(&top_root);
    const uint32_t n_top = (uint32_t)out.left.size();
    // subtrees: each task gets its own Builder view (shared, already partitioned index array; disjoint ranges)
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t t = next.fetch_add(1);
            if (t >= tasks.size()) return;
            Task& tk = tasks[t];
            detail::Builder b(leaf, tk.tree, max_depth, &top.idx, top.cen);
            tk.tree.root = b.build(tk.lo, tk.hi, tk.depth);
        }
    };
    std::vector<std::thread> pool;
    const unsigned n_thr = (unsigned)std::min<size_t>(hw, tasks.size());
    for (unsigned i = 1; i < n_thr; ++i) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    // splice: offsets by prefix sum, then every task copies its nodes into place (parallel over tasks)
    std::vector<uint32_t> task_root(tasks.size()), task_off(tasks.size());
    {
        uint32_t off = n_top;
        for (size_t t = 0; t < tasks.size(); ++t) { task_off[t] = off; off += (uint32_t)tasks[t].tree.left.size(); }
        out.left.resize(off); out.right.resize(off); out.axis.resize(off); out.box.resize(off);
    }
    parallel_for(tasks.size(), 1, [&](size_t t_lo, size_t t_hi) {
        for (size_t t = t_lo; t < t_hi; ++t) {
            const Tree& tr = tasks[t].tree;
            const uint32_t off = task_off[t];
            auto fix = [&](uint32_t c) { return (c & 0x80000000u) ? c : c + off; };
            for (size_t i = 0; i < tr.left.size(); ++i) {
                out.left[off + i] = fix(tr.left[i]); out.right[off + i] = fix(tr.right[i]);
                out.axis[off + i] = tr.axis[i]; out.box[off + i] = tr.box[i];
            }
            task_root[t] = fix(tr.root);
        }
    });
    for (size_t t = 0; t < tasks.size(); ++t) out.depth = std::max(out.depth, tasks[t].tree.depth);
    for (uint32_t i = n_top; i-- > 0;) {  // patch task references, then the top boxes bottom-up
        for (uint32_t* c : {&out.left[i], &out.right[i]})
            if (!(*c & 0x80000000u) && (*c & TASK_BIT)) *c = task_root[*c & ~TASK_BIT];
        Box3 bx = detail::empty_box();
        for (uint32_t c : {out.left[i], out.right[i]}) detail::grow(bx, (c & 0x80000000u) ? leaf[c & 0x7FFFFFFFu] : out.box[c]);
        out.box[i] = bx;
    }
    out.root = (!(root & 0x80000000u) && (root & TASK_BIT)) ? task_root[root & ~TASK_BIT] : root;
}

}  // namespace rt_rebuild
