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
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

#include <math.h>
#ifdef __CUDACC__
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif

namespace rt_rebuild {

// Worker threads shared by every parallel_for of the process (created on first use).  Starting a thread costs 20-30 us; the scene
// hand-off of C5 issues ~100 parallel loops (24 pre-split rounds among them), which on 16 cores was a fifth of its run time.
// Calls may come from several threads at once and from inside a chunk: a caller runs its own first chunk, then helps with
// whatever is queued until its own chunks are done, so nobody ever waits on a thread that waits on it.
class Pool {
  public:
    struct Job {
        const std::function<void(size_t, size_t)>* fn;
        size_t a, b;
        std::atomic<size_t>* pending;
    };
    static constexpr unsigned MAX_THREADS = 32;  // these loops are memory-bound well before that; several processes per box (one per GPU) share the cores
    static Pool& get() {
        static Pool p;
        return p;
    }
    size_t threads() const { return workers_.size() + 1; }
    void run(size_t n, size_t parts, const std::function<void(size_t, size_t)>& fn) {
        std::atomic<size_t> pending{parts - 1};
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t p = 1; p < parts; ++p) q_.push_back(Job{&fn, n * p / parts, n * (p + 1) / parts, &pending});
        }
        // wake as many workers as there are chunks to take, not all of them: on a 128-core box a round of four chunks would otherwise
        // send 127 threads through the queue's mutex to find it empty
        if (parts - 1 >= workers_.size()) cv_.notify_all();
        else for (size_t p = 1; p < parts; ++p) cv_.notify_one();
        fn((size_t)0, n / parts);
        for (int idle = 0; pending.load(std::memory_order_acquire) != 0;) {
            Job j;
            if (pop(j)) { exec(j); idle = 0; }
            else if (++idle < 256) std::this_thread::yield();
            else std::this_thread::sleep_for(std::chrono::microseconds(50));  // a long chunk elsewhere: stop burning a core the workers could use
        }
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
    Pool() {
        const unsigned hw = std::min(MAX_THREADS, std::max(1u, std::thread::hardware_concurrency()));
        for (unsigned i = 1; i < hw; ++i)
            workers_.emplace_back([this] {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                        if (q_.empty()) return;
                        j = q_.front();
                        q_.pop_front();
                    }
                    exec(j);
                }
            });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    bool pop(Job& j) {
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty()) return false;
        j = q_.front();
        q_.pop_front();
        return true;
    }
    static void exec(const Job& j) {
        (*j.fn)(j.a, j.b);
        j.pending->fetch_sub(1, std::memory_order_release);
    }
};

// Static chunking over hardware threads; fn(begin, end) works on disjoint index ranges, so results do not depend on the thread count.
template <class F>
inline void parallel_for(size_t n, size_t min_chunk, F&& fn) {
    const size_t hw = std::min(Pool::MAX_THREADS, std::max(1u, std::thread::hardware_concurrency()));
    const size_t parts = std::max<size_t>(1, std::min(hw, n / std::max<size_t>(1, min_chunk)));
    if (parts <= 1) { fn((size_t)0, n); return; }
    const std::function<void(size_t, size_t)> f = [&fn](size_t a, size_t b) { fn(a, b); };
    Pool::get().run(n, parts, f);
}

struct Box3 {
    float mn[3], mx[3];
};

// Uninitialised storage for trivially copyable records: a std::vector would zero-fill tens of megabytes on the calling thread
// before the parallel loop that writes them (first touch, page faults included, then happens on the worker threads instead).
template <class T>
struct Buffer {
    std::unique_ptr<T[]> p;
    size_t n = 0, cap = 0;
    void reserve_discard(size_t c) { if (c > cap) { p.reset(new T[c]); cap = c; } }
    void resize_discard(size_t c) { reserve_discard(c); n = c; }  // contents undefined
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T* data() { return p.get(); }
    const T* data() const { return p.get(); }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* begin() { return p.get(); }
    T* end() { return p.get() + n; }
    const T* begin() const { return p.get(); }
    const T* end() const { return p.get() + n; }
};

struct Tree {
    // n - 1 nodes over n leaves in PRE-ORDER: the node over the leaf positions [lo, hi) has its left subtree at me + 1 and its right
    // subtree at me + (mid - lo), so every subtree owns a known index range before it is built (children have larger indices).
    Buffer<uint32_t> left, right;  // child: 0x80000000 | element, or node index
    Buffer<uint32_t> axis;         // split axis: the left child holds the lower centroids on it
    Buffer<Box3> box;              // each node's own box (exact union of the element boxes below it)
    Buffer<float> band_a, band_c;  // only with LeafSrc::item_a: maxima of the leaves' two guard-band coefficients below each node
    uint32_t root = 0;
    uint32_t depth = 0;
};


// The leaves handed to build(): boxes at a byte stride (a plain Box3 array, or the boxes inside pre-split references), and optionally
// two per-ELEMENT coefficients whose subtree maxima the builder collects on its way back up (the guard band of the tight cull).
struct LeafSrc {
    const char* box0 = nullptr;   // &leaf[0].box
    size_t stride = sizeof(Box3);
    size_t n = 0;
    const char* item0 = nullptr;  // &leaf[0].item (same stride), or NULL: the leaf is its own element
    const float* item_a = nullptr;
    const float* item_c = nullptr;
    const Box3& box(size_t i) const { return *reinterpret_cast<const Box3*>(box0 + i * stride); }
    uint32_t item(size_t i) const { return item0 ? *reinterpret_cast<const uint32_t*>(item0 + i * stride) : (uint32_t)i; }
};

namespace detail {

RT_HD inline Box3 empty_box() {
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
RT_HD inline double half_area(const Box3& b) {
    double x = (double)b.mx[0] - b.mn[0], y = (double)b.mx[1] - b.mn[1], z = (double)b.mx[2] - b.mn[2];
    x = 0.0 < x ? x : 0.0; y = 0.0 < y ? y : 0.0; z = 0.0 < z ? z : 0.0;  // std::max(0.0, .) without the host-only call (device code uses this too)
    return x * y + y * z + z * x;
}

// One record per reference: its box and its index, moved together when a range is partitioned, so that every pass over a
// range streams through memory (the first version kept an index array and gathered boxes and three centroid arrays through it:
// 4-6 cache lines per reference and pass, 2/3 of the build time for the 2 * 10^6 references of C4).
struct Rec {
    Box3 box;
    uint32_t id;
    uint32_t pad;
};
inline float centroid(const Rec& r, int k) { return 0.5f * r.box.mn[k] + 0.5f * r.box.mx[k]; }

struct Builder {
    const LeafSrc& src;
    Tree& out;
    Buffer<Rec> own_recs;
    Buffer<Rec>& recs;                 // reference order (partitioned in place); shared between the builders of one build()
    uint32_t max_depth;
    uint32_t depth_seen = 0;           // deepest internal level this builder made
    static const int BINS = 32;
    static const uint32_t SWEEP_MAX = 96;
    static const uint32_t PAR_SPLIT_MIN = 1u << 18;

    Builder(const LeafSrc& l, Tree& t, uint32_t md) : src(l), out(t), recs(own_recs), max_depth(md) {
        own_recs.resize_discard(l.n);
        parallel_for(l.n, 65536, [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) { own_recs[i].box = l.box(i); own_recs[i].id = (uint32_t)i; own_recs[i].pad = 0; }
        });
    }
    Builder(const LeafSrc& l, Tree& t, uint32_t md, Buffer<Rec>* shared) : src(l), out(t), recs(*shared), max_depth(md) {}
    void child_band(uint32_t c, uint32_t leaf_pos, float& a, float& cc) const {
        if (!src.item_a) return;
        if (c & 0x80000000u) {
            const uint32_t it = src.item(recs[leaf_pos].id);
            a = std::max(a, src.item_a[it]); cc = std::max(cc, src.item_c[it]);
        } else {
            a = std::max(a, out.band_a[c]); cc = std::max(cc, out.band_c[c]);
        }
    }

    // stable merge sort of (key, position) pairs, n <= SWEEP_MAX, no allocation: the unique stable order, i.e. what std::stable_sort gives
    struct KeyPos { float key; uint32_t pos; };
    static void stable_sort_small(KeyPos* a, uint32_t n, KeyPos* tmp) {
        for (uint32_t i = 1; i < n; ++i) {  // insertion sort of runs of 8
            if (i % 8 == 0) continue;
            KeyPos v = a[i];
            uint32_t j = i;
            while (j % 8 != 0 && v.key < a[j - 1].key) { a[j] = a[j - 1]; --j; }
            a[j] = v;
        }
        KeyPos *src = a, *dst = tmp;
        for (uint32_t w = 8; w < n; w *= 2) {
            for (uint32_t s = 0; s < n; s += 2 * w) {
                const uint32_t m = std::min(s + w, n), e = std::min(s + 2 * w, n);
                uint32_t i = s, j = m, o = s;
                while (i < m && j < e) dst[o++] = (src[j].key < src[i].key) ? src[j++] : src[i++];  // ties: left run first (stable)
                while (i < m) dst[o++] = src[i++];
                while (j < e) dst[o++] = src[j++];
            }
            std::swap(src, dst);
        }
        if (src != a) std::copy(src, src + n, a);
    }

    // returns the split position (lo < mid < hi) after reordering recs[lo, hi)
    uint32_t split(uint32_t lo, uint32_t hi, uint32_t levels_left, uint32_t* axis_out) {
        const uint32_t n = hi - lo;
        const uint64_t child_cap = levels_left >= 1 ? (levels_left - 1 >= 63 ? ~0ull : (1ull << (levels_left - 1))) : 1ull;
        float cmin[3], cmax[3];
        for (int k = 0; k < 3; ++k) { cmin[k] = INFINITY; cmax[k] = -INFINITY; }
        const bool big = n >= PAR_SPLIT_MIN;  // the first few levels of a large build: scan and bin on all cores
        Rec* const R = recs.data();
        auto bounds = [&](size_t a, size_t b, float* mn, float* mx) {
            for (size_t i = a; i < b; ++i)
                for (int k = 0; k < 3; ++k) {
                    const float c = centroid(R[i], k);
                    mn[k] = c < mn[k] ? c : mn[k];
                    mx[k] = c > mx[k] ? c : mx[k];
                }
        };
        if (big) {
            std::mutex mu;
            parallel_for(n, 65536, [&](size_t a, size_t b) {
                float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
                bounds(lo + a, lo + b, mn, mx);
                std::lock_guard<std::mutex> lk(mu);
                for (int k = 0; k < 3; ++k) { cmin[k] = std::min(cmin[k], mn[k]); cmax[k] = std::max(cmax[k], mx[k]); }
            });
        } else {
            bounds(lo, hi, cmin, cmax);
        }
        int best_axis = -1;
        uint32_t best_pos = 0;  // sweep: number of references on the left; bins: last left bin
        double best_cost = INFINITY;
        if (n <= SWEEP_MAX) {
            KeyPos ord[3][SWEEP_MAX], tmp[SWEEP_MAX];
            double right_area[SWEEP_MAX];
            for (int k = 0; k < 3; ++k) {
                if (!(cmax[k] > cmin[k])) continue;
                KeyPos* o = ord[k];
                for (uint32_t i = 0; i < n; ++i) o[i] = KeyPos{centroid(R[lo + i], k), i};
                stable_sort_small(o, n, tmp);
                Box3 acc = empty_box();
                for (uint32_t i = n; i-- > 1;) { grow(acc, R[lo + o[i].pos].box); right_area[i] = half_area(acc); }
                acc = empty_box();
                for (uint32_t i = 1; i < n; ++i) {
                    grow(acc, R[lo + o[i - 1].pos].box);
                    const double c = half_area(acc) * i + right_area[i] * (n - i);
                    if (c < best_cost && i <= child_cap && n - i <= child_cap) { best_cost = c; best_axis = k; best_pos = i; }
                }
            }
            if (best_axis >= 0) {
                Rec sorted[SWEEP_MAX];
                const KeyPos* o = ord[best_axis];
                for (uint32_t i = 0; i < n; ++i) sorted[i] = R[lo + o[i].pos];
                std::copy(sorted, sorted + n, R + lo);
                *axis_out = (uint32_t)best_axis;
                return lo + best_pos;
            }
        } else {
            // all three axes in one pass over the range
            struct Bins { Box3 bb[3][BINS]; uint32_t cnt[3][BINS]; };
            auto clear = [](Bins& B) { for (int k = 0; k < 3; ++k) for (int b = 0; b < BINS; ++b) { B.bb[k][b] = empty_box(); B.cnt[k][b] = 0; } };
            bool use[3];
            float scale[3];
            for (int k = 0; k < 3; ++k) {
                use[k] = (cmax[k] > cmin[k]) && std::isfinite(cmax[k] - cmin[k]);
                scale[k] = use[k] ? (float)BINS / (cmax[k] - cmin[k]) : 0.0f;
            }
            auto bin_range = [&](size_t a, size_t e, Bins& B) {
                for (size_t i = a; i < e; ++i)
                    for (int k = 0; k < 3; ++k) {
                        if (!use[k]) continue;
                        int b = (int)((centroid(R[i], k) - cmin[k]) * scale[k]);
                        b = std::min(BINS - 1, std::max(0, b));
                        grow(B.bb[k][b], R[i].box);
                        ++B.cnt[k][b];
                    }
            };
            Bins B;
            clear(B);
            if (big) {  // per-thread bins, merged (box union and counts are order-independent: same result as the serial loop)
                std::mutex mu;
                parallel_for(n, 65536, [&](size_t a, size_t e) {
                    Bins L;
                    clear(L);
                    bin_range(lo + a, lo + e, L);
                    std::lock_guard<std::mutex> lk(mu);
                    for (int k = 0; k < 3; ++k)
                        for (int b = 0; b < BINS; ++b) { if (L.cnt[k][b]) grow(B.bb[k][b], L.bb[k][b]); B.cnt[k][b] += L.cnt[k][b]; }
                });
            } else {
                bin_range(lo, hi, B);
            }
            for (int k = 0; k < 3; ++k) {
                if (!use[k]) continue;
                double ra[BINS];
                uint32_t rc[BINS];
                Box3 acc = empty_box();
                uint32_t c = 0;
                for (int b = BINS - 1; b >= 1; --b) { grow(acc, B.bb[k][b]); c += B.cnt[k][b]; ra[b] = half_area(acc); rc[b] = c; }
                acc = empty_box();
                c = 0;
                for (int b = 0; b < BINS - 1; ++b) {
                    grow(acc, B.bb[k][b]);
                    c += B.cnt[k][b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    const double cost = half_area(acc) * c + ra[b + 1] * rc[b + 1];
                    if (cost < best_cost && c <= child_cap && rc[b + 1] <= child_cap) { best_cost = cost; best_axis = k; best_pos = (uint32_t)b; }
                }
            }
            if (best_axis >= 0) {
                const int k = best_axis;
                const float sc = scale[k], c0 = cmin[k];
                auto goes_left = [&](const Rec& a) {
                    int b = (int)((centroid(a, k) - c0) * sc);
                    b = std::min(BINS - 1, std::max(0, b));
                    return (uint32_t)b <= best_pos;
                };
                uint32_t mid;
                if (big) {
                    // stable partition on all cores through a scratch copy: chunk counts, prefix, scatter, copy back (the chunking is
                    // fixed, and a stable partition has one result anyway: no dependence on the thread count)
                    const size_t CH = 65536, n_ch = (n + CH - 1) / CH;
                    std::vector<uint32_t> n_left(n_ch + 1, 0);
                    parallel_for(n_ch, 1, [&](size_t a, size_t e) {
                        for (size_t c = a; c < e; ++c) {
                            uint32_t cnt = 0;
                            for (size_t i = lo + c * CH, end = std::min<size_t>(hi, lo + (c + 1) * CH); i < end; ++i) cnt += goes_left(R[i]) ? 1u : 0u;
                            n_left[c + 1] = cnt;
                        }
                    });
                    for (size_t c = 0; c < n_ch; ++c) n_left[c + 1] += n_left[c];
                    const uint32_t total_left = n_left[n_ch];
                    Buffer<Rec> tmp;
                    tmp.resize_discard(n);
                    parallel_for(n_ch, 1, [&](size_t a, size_t e) {
                        for (size_t c = a; c < e; ++c) {
                            size_t ol = n_left[c], orr = total_left + (c * CH - n_left[c]);
                            for (size_t i = lo + c * CH, end = std::min<size_t>(hi, lo + (c + 1) * CH); i < end; ++i) {
                                if (goes_left(R[i])) tmp[ol++] = R[i]; else tmp[orr++] = R[i];
                            }
                        }
                    });
                    parallel_for(n, 65536, [&](size_t a, size_t e) { std::copy(tmp.data() + a, tmp.data() + e, R + lo + a); });
                    mid = lo + total_left;
                } else {
                    mid = (uint32_t)(std::partition(R + lo, R + hi, goes_left) - R);
                }
                if (mid > lo && mid < hi) { *axis_out = (uint32_t)k; return mid; }
            }
        }
        // no useful plane (coincident centroids, NaNs, or the depth cap binds): median split on the widest centroid axis
        int k = 0;
        for (int a = 1; a < 3; ++a)
            if (cmax[a] - cmin[a] > cmax[k] - cmin[k]) k = a;
        const uint32_t mid = lo + n / 2;
        std::nth_element(R + lo, R + mid, R + hi, [&](const Rec& a, const Rec& b) { return centroid(a, k) < centroid(b, k); });
        *axis_out = (uint32_t)k;
        return mid;
    }

    // node `me` over recs[lo, hi); returns the child reference its parent stores
    uint32_t build(uint32_t lo, uint32_t hi, uint32_t depth, uint32_t me) {
        if (hi - lo == 1) return 0x80000000u | recs[lo].id;
        depth_seen = std::max(depth_seen, depth + 1);
        uint32_t ax = 0;
        const uint32_t mid = split(lo, hi, max_depth - depth, &ax);
        out.axis[me] = ax;
        const uint32_t l = build(lo, mid, depth + 1, me + 1);
        const uint32_t r = build(mid, hi, depth + 1, me + (mid - lo));
        finish_node(me, l, r, lo, mid);
        return me;
    }
    void finish_node(uint32_t me, uint32_t l, uint32_t r, uint32_t lo, uint32_t mid) {
        out.left[me] = l;
        out.right[me] = r;
        // a leaf child's box is still at its position: recs[lo] / recs[mid] when that side has one reference
        Box3 b = empty_box();
        grow(b, (l & 0x80000000u) ? recs[lo].box : out.box[l]);
        grow(b, (r & 0x80000000u) ? recs[mid].box : out.box[r]);
        out.box[me] = b;
        if (src.item_a) {
            float a = 0.0f, c = 0.0f;
            child_band(l, lo, a, c);
            child_band(r, mid, a, c);
            out.band_a[me] = a; out.band_c[me] = c;
        }
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
RT_HD inline void clip_half(const Poly& in, int axis, double c, double sign, Poly& out) {  // keep sign * (x[axis] - c) >= 0
    out.n = 0;
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
}
RT_HD inline float round_down(double v) {
    float f = (float)v;
    return (double)f > v ? nextafterf(f, -INFINITY) : f;
}
RT_HD inline float round_up(double v) {
    float f = (float)v;
    return (double)f < v ? nextafterf(f, INFINITY) : f;
}
RT_HD inline Box3 poly_box(const Poly& p, const Box3& within) {
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

// The cell a reference's cuts have carved out of space (its piece of the triangle = the triangle clipped against it).
struct Region {
    double lo[3], hi[3];
};
struct Cut {
    Box3 bl, bh;  // boxes of the low / high piece
    double mid;
    int ax;
    bool ok;
};
// One attempted cut of a reference (box b, region rg) of the triangle tri9 at the middle of the box's longest axis.  The ONE
// implementation, compiled for the host loop below and for the device kernel (rt_presplit.cuh): double arithmetic without
// contraction on both sides (-ffp-contract=off / --fmad=false), so both produce the same boxes bit for bit.
RT_HD inline void compute_cut(const Box3& b, const float* tri9, const Region& rg, Cut& cu) {
    using detail::Poly;
    cu.ok = false;
    cu.mid = 0.0;
    cu.ax = 0;
    int ax = 0;
    for (int q = 1; q < 3; ++q)
        if (b.mx[q] - b.mn[q] > b.mx[ax] - b.mn[ax]) ax = q;
    if (!(b.mx[ax] > b.mn[ax])) return;
    const double mid = 0.5 * ((double)b.mn[ax] + (double)b.mx[ax]);
    Poly p0, p1;
    p0.n = 3;
    for (int v = 0; v < 3; ++v)
        for (int q = 0; q < 3; ++q) p0.v[v][q] = tri9[3 * v + q];
    Poly* src = &p0;
    Poly* dst = &p1;
    for (int q = 0; q < 3 && src->n >= 3; ++q) {
        if (rg.lo[q] > -INFINITY) { detail::clip_half(*src, q, rg.lo[q], 1.0, *dst); Poly* t = src; src = dst; dst = t; }
        if (rg.hi[q] < INFINITY && src->n >= 3) { detail::clip_half(*src, q, rg.hi[q], -1.0, *dst); Poly* t = src; src = dst; dst = t; }
    }
    if (src->n < 3) return;
    Poly lo;
    detail::clip_half(*src, ax, mid, -1.0, lo);
    detail::clip_half(*src, ax, mid, 1.0, *dst);
    if (lo.n < 3 || dst->n < 3) return;
    cu.bl = detail::poly_box(lo, b);
    cu.bh = detail::poly_box(*dst, b);
    const double a0 = detail::half_area(b), al = detail::half_area(cu.bl), ah = detail::half_area(cu.bh);
    cu.ok = al + ah < 0.9 * a0;
    cu.mid = mid;
    cu.ax = ax;
}

// tri: 9 floats (p0, p1, p2) per element, only read where is_tri[i] != 0.
// Largest boxes first, in ROUNDS: each round cuts every candidate whose box area is at least tau, in parallel, and feeds the
// pieces back as candidates; tau halves when nothing that large is left.  A piece is always smaller than the reference it came
// from (accepted cuts shrink the summed area to < 0.9x), so this visits the references in the order of a sequential "pop the
// largest" loop up to a factor 2 in area, and the budget ends on the largest candidates of the last round (ties by reference index).
// Candidates wait in one bucket per halving of tau (a round touches only its own candidates; the first version re-partitioned the
// whole candidate array every round: 2/3 of the pre-split time on C4), and a round's results are written in parallel at positions
// given by a prefix sum over the accepted cuts.  A reference carries no polygon: its piece of the triangle is re-derived on demand
// by clipping the triangle against the reference's REGION (the axis-aligned cell its cuts have carved out: 48 B instead of a
// 300 B polygon).  Deterministic for any thread count.
inline void presplit(const Box3* leaf, size_t n0, const float* tri, const uint8_t* is_tri, size_t budget, Buffer<Ref>& refs) {
    refs.resize_discard(std::max(budget, n0));  // sized once; `n_refs` of them are live
    Buffer<double> area;  // scratch: box areas computed in the parallel loops, consumed by the sequential bucket fills
    area.resize_discard(std::max(n0, budget > n0 ? 2 * (budget - n0) : (size_t)0));
    parallel_for(n0, 65536, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) { refs[i] = Ref{(uint32_t)i, leaf[i]}; area[i] = detail::half_area(leaf[i]); }
    });
    size_t n_refs = n0;
    if (budget <= n0) { refs.n = n0; return; }
    Buffer<Region> region;
    region.resize_discard(budget);
    parallel_for(n0, 65536, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) region[i] = Region{{-INFINITY, -INFINITY, -INFINITY}, {INFINITY, INFINITY, INFINITY}};
    });
    struct Cand { double area; uint32_t ref; };
    double tau0 = 0.0;
    for (uint32_t i = 0; i < n0; ++i)
        if (is_tri[i]) tau0 = std::max(tau0, area[i]);
    tau0 *= 0.5;
    if (!(tau0 > 0.0) || !std::isfinite(tau0)) { refs.n = n0; return; }
    // bucket j holds the candidates with tau0 / 2^j <= area < tau0 / 2^(j-1)  (bucket 0: everything >= tau0)
    std::vector<double> tau;  // tau[j] = tau0 / 2^j while that is a normal number
    for (double t = tau0; t >= 2.3e-308; t *= 0.5) tau.push_back(t);
    auto exponent = [](double v) -> int { uint64_t b; std::memcpy(&b, &v, 8); return (int)((b >> 52) & 0x7FF); };
    const int e0 = exponent(tau0);
    auto bucket_of = [&](double area) -> int {
        if (!(area >= tau.back())) return -1;  // zero, NaN or below every threshold: never reached
        if (area >= tau0) return 0;
        int j = std::max(e0 - exponent(area) - 1, 1);  // tau[j] has exponent e0 - j: area >= tau[j] needs j >= e0 - exponent(area) - 1 ... + 1
        while (area < tau[(size_t)j]) ++j;
        return j;
    };
    std::vector<std::vector<Cand>> buckets(tau.size());
    auto put = [&](const Cand& c) {
        const int j = bucket_of(c.area);
        if (j >= 0) buckets[(size_t)j].push_back(c);
    };
    for (uint32_t i = 0; i < n0; ++i)
        if (is_tri[i]) put(Cand{area[i], i});
    Buffer<Cut> cuts;
    Buffer<int> piece_bucket;  // of the two pieces of cut c: computed in the parallel loop, the sequential queueing only appends
    std::vector<Cand> cur;
    std::vector<uint32_t> slot;
    for (size_t j = 0; j < buckets.size() && n_refs < budget;) {
        if (buckets[j].empty()) { ++j; continue; }
        cur.clear();
        cur.swap(buckets[j]);
        const size_t room = budget - n_refs;
        size_t k = cur.size();
        if (k > room) {  // the budget ends inside this round: the `room` largest of them (ties by reference index: deterministic)
            auto larger = [](const Cand& a, const Cand& b) { return a.area > b.area || (a.area == b.area && a.ref < b.ref); };
            std::nth_element(cur.begin(), cur.begin() + room, cur.end(), larger);
            cur.resize(room);
            k = room;
        }
        cuts.resize_discard(k);
        parallel_for(k, 192, [&](size_t lo_i, size_t hi_i) {  // (a cut costs ~0.3 us: rounds of a few hundred candidates already pay for the pool)
            for (size_t c = lo_i; c < hi_i; ++c) {
                const uint32_t r = cur[c].ref;
                compute_cut(refs[r].box, tri + 9 * (size_t)refs[r].item, region[r], cuts[c]);
            }
        });
        // apply: the low half replaces the reference, the high half goes to a new reference at n_refs + (accepted cuts before it);
        // both pieces are smaller than the reference they came from and become candidates again
        slot.resize(k);
        piece_bucket.resize_discard(2 * k);
        uint32_t n_ok = 0;
        for (size_t c = 0; c < k; ++c) { slot[c] = n_ok; n_ok += cuts[c].ok ? 1u : 0u; }
        const size_t base = n_refs;
        parallel_for(k, 4096, [&](size_t lo_i, size_t hi_i) {
            for (size_t c = lo_i; c < hi_i; ++c) {
                if (!cuts[c].ok) continue;
                const uint32_t r = cur[c].ref, r2 = (uint32_t)(base + slot[c]);
                refs[r].box = cuts[c].bl;
                refs[r2] = Ref{refs[r].item, cuts[c].bh};
                Region hi_region = region[r];
                region[r].hi[cuts[c].ax] = cuts[c].mid;
                hi_region.lo[cuts[c].ax] = cuts[c].mid;
                region[r2] = hi_region;
                area[2 * c] = detail::half_area(cuts[c].bl);
                area[2 * c + 1] = detail::half_area(cuts[c].bh);
                piece_bucket[2 * c] = bucket_of(area[2 * c]);
                piece_bucket[2 * c + 1] = bucket_of(area[2 * c + 1]);
            }
        });
        n_refs += n_ok;
        if (n_refs >= budget) break;  // nothing will be cut any more: the pieces of this round need not be queued
        for (size_t c = 0; c < k; ++c) {
            if (!cuts[c].ok) continue;
            if (piece_bucket[2 * c] >= 0) buckets[(size_t)piece_bucket[2 * c]].push_back(Cand{area[2 * c], cur[c].ref});
            if (piece_bucket[2 * c + 1] >= 0) buckets[(size_t)piece_bucket[2 * c + 1]].push_back(Cand{area[2 * c + 1], (uint32_t)(base + slot[c])});
        }
    }
    refs.n = n_refs;
}

inline void presplit(const std::vector<Box3>& leaf, const std::vector<float>& tri, const std::vector<uint8_t>& is_tri, size_t budget,
                     Buffer<Ref>& refs) {
    presplit(leaf.data(), leaf.size(), tri.data(), is_tri.data(), budget, refs);
}

// max_depth: number of internal levels allowed (>= ceil(log2 n)).  Leaves are indices into `leaf` (boxes of references).
// Ranges larger than n / (4 x threads) are split on the calling thread; the subtrees below are built concurrently (they work
// on disjoint ranges of the index array) and appended, so every child still has a larger index than its parent.
inline void build(const LeafSrc& leaf, uint32_t max_depth, Tree& out) {
    out = Tree();
    if (leaf.n == 0) return;
    uint32_t need = 0;
    while ((1ull << need) < leaf.n) ++need;
    max_depth = std::max(max_depth, need);
    const size_t n_nodes = leaf.n - 1;
    out.left.resize_discard(n_nodes); out.right.resize_discard(n_nodes); out.axis.resize_discard(n_nodes); out.box.resize_discard(n_nodes);
    if (leaf.item_a) { out.band_a.resize_discard(n_nodes); out.band_c.resize_discard(n_nodes); }
    const unsigned hw = std::min(std::max(1u, std::thread::hardware_concurrency()), leaf.n < 100000 ? 8u : 32u);
    detail::Builder top(leaf, out, max_depth);
    if (leaf.n < 4096 || hw < 2) {
        out.root = top.build(0, (uint32_t)leaf.n, 0, 0);
        out.depth = top.depth_seen;
        return;
    }
    // Ranges above PAR_MIN are split here, fork-join (the two halves of a split work on disjoint ranges of the record array, so they
    // run on two threads down to FORK_LEVELS; a split of >= 2^18 references also scans, bins and partitions on all cores).  The
    // ranges below become tasks for a pool; a task writes its nodes straight into its own index range of the tree.  The result of a
    // split depends on its range alone, so the tree is the same for any thread count.
    const uint32_t PAR_MIN = (uint32_t)std::max<size_t>(1024, leaf.n / (16 * (size_t)hw));
    struct TopNode {
        uint32_t lo = 0, hi = 0, depth = 0, mid = 0, axis = 0, me = 0;
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
            nd->l->me = nd->me + 1; nd->r->me = nd->me + (nd->mid - nd->lo);
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
    top_root.lo = 0; top_root.hi = (uint32_t)leaf.n; top_root.depth = 0; top_root.me = 0;
    splitter.go(&top_root);
    // the top nodes in pre-order, the tasks left to right (sequential, a few thousand nodes)
    std::vector<TopNode*> order, tasks;
    {
        std::vector<TopNode*> st{&top_root};
        while (!st.empty()) {
            TopNode* nd = st.back();
            st.pop_back();
            if (nd->hi - nd->lo == 1) continue;
            if (nd->task) { tasks.push_back(nd); continue; }
            order.push_back(nd);
            st.push_back(nd->r.get());
            st.push_back(nd->l.get());
        }
    }
    std::atomic<size_t> next{0};
    std::atomic<uint32_t> depth_max{0};
    auto worker = [&]() {
        uint32_t deepest = 0;
        for (;;) {
            const size_t t = next.fetch_add(1);
            if (t >= tasks.size()) break;
            TopNode* tk = tasks[t];
            detail::Builder b(leaf, out, max_depth, &top.recs);
            b.build(tk->lo, tk->hi, tk->depth, tk->me);
            deepest = std::max(deepest, b.depth_seen);
        }
        uint32_t seen = depth_max.load();
        while (seen < deepest && !depth_max.compare_exchange_weak(seen, deepest)) {}
    };
    {
        const size_t n_thr = std::min<size_t>(hw, tasks.size());
        parallel_for(n_thr, 1, [&](size_t a, size_t b) { for (size_t i = a; i < b; ++i) worker(); });  // the shared pool's threads pull tasks
    }
    out.depth = depth_max.load();
    auto ref_of = [&](const TopNode* c) -> uint32_t { return c->hi - c->lo == 1 ? (0x80000000u | top.recs[c->lo].id) : c->me; };
    for (size_t i = order.size(); i-- > 0;) {  // the top nodes bottom-up: children, boxes, band maxima
        const TopNode* nd = order[i];
        out.axis[nd->me] = nd->axis;
        top.finish_node(nd->me, ref_of(nd->l.get()), ref_of(nd->r.get()), nd->lo, nd->mid);
        out.depth = std::max(out.depth, nd->depth + 1);
    }
    out.root = ref_of(&top_root);
}
inline void build(const std::vector<Box3>& leaf, uint32_t max_depth, Tree& out) {
    LeafSrc src;
    src.box0 = reinterpret_cast<const char*>(leaf.data());
    src.n = leaf.size();
    build(src, max_depth, out);
}
// over pre-split references, collecting the subtree maxima of item_a / item_c (per element) in out.band_a / band_c
inline void build(const Buffer<Ref>& refs, const float* item_a, const float* item_c, uint32_t max_depth, Tree& out) {
    LeafSrc src;
    src.box0 = reinterpret_cast<const char*>(&refs.data()->box);
    src.item0 = reinterpret_cast<const char*>(&refs.data()->item);
    src.stride = sizeof(Ref);
    src.n = refs.size();
    src.item_a = item_a;
    src.item_c = item_c;
    build(src, max_depth, out);
}

}  // namespace rt_rebuild
