// Host-only invariants of the accelerator-tree builder (raytracinginaweekend_b200/csrc/rt_rebuild.hpp), run by tests/test_rebuild_cpu.py.
//   rebuild_check <n_elements> <seed> <budget_factor> <max_depth_extra> <mode>     mode 0: random triangles + spheres, 1: slivers fanned
//   around one point (the reference's OBJ fan, Q4), 2: all elements identical (coincident centroids)
// Checks: every reference box lies inside its element's box; sampled points of every triangle are covered by one of its
// references; every node box is the exact union of its children; children have larger indices than their parent; every
// reference is a leaf exactly once; the depth cap holds.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../raytracinginaweekend_b200/csrc/rt_rebuild.hpp"

using namespace rt_rebuild;
static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails < 10) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } ++fails; } } while (0)

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const size_t n = std::strtoul(argv[1], nullptr, 10);
    std::mt19937 rng((unsigned)std::atoi(argv[2]));
    const double factor = std::atof(argv[3]);
    const int extra = std::atoi(argv[4]), mode = std::atoi(argv[5]);
    if (mode == 3) {
        // The worker pool behind every parallel loop (rt_rebuild::Pool): loops issued from several threads at once, each with a loop nested
        // inside its chunks (the scene hand-off does both: helper threads beside the builder, whose big splits loop again).  A caller runs
        // its first chunk and then helps with whatever is queued, so this must neither deadlock nor lose a chunk.
        std::atomic<uint64_t> total{0};
        std::vector<std::thread> callers;
        for (int t = 0; t < 3; ++t)
            callers.emplace_back([&, t] {
                for (int rep = 0; rep < 4; ++rep)
                    parallel_for(n, 64, [&](size_t a, size_t b) {
                        uint64_t local = 0;
                        parallel_for(b - a, 16, [&](size_t c, size_t d) {
                            uint64_t s = 0;
                            for (size_t i = a + c; i < a + d; ++i) s += i % 7 + (uint64_t)t;
                            total.fetch_add(s);
                        });
                        (void)local;
                    });
            });
        for (auto& th : callers) th.join();
        uint64_t expect = 0;
        for (int t = 0; t < 3; ++t)
            for (size_t i = 0; i < n; ++i) expect += 4 * (i % 7 + (uint64_t)t);
        CHECK(total.load() == expect, "pool lost or repeated work: %llu vs %llu", (unsigned long long)total.load(), (unsigned long long)expect);
        std::printf("%s pool n=%zu\n", fails ? "FAILED" : "OK", n);
        return fails ? 1 : 0;
    }
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    std::vector<Box3> leaf(n);
    std::vector<float> tri(9 * n, 0.0f);
    std::vector<uint8_t> is_tri(n, 0);
    for (size_t i = 0; i < n; ++i) {
        const bool t = mode != 0 || (i % 3) != 0;
        is_tri[i] = t;
        if (t) {
            float* p = &tri[9 * i];
            if (mode == 1) {  // sliver: one vertex at the hub, two nearby vertices far away
                const float a = U(rng) * 3.14159f, r = 2.0f + U(rng);
                p[0] = 0.5f; p[1] = 0.25f; p[2] = -0.5f;
                p[3] = 0.5f + r * std::cos(a); p[4] = 0.25f + U(rng); p[5] = -0.5f + r * std::sin(a);
                p[6] = p[3] + 0.05f * U(rng); p[7] = p[4] + 0.05f * U(rng); p[8] = p[5] + 0.05f * U(rng);
            } else if (mode == 2) {
                const float q[9] = {0, 0, 0, 1, 0, 0, 0, 1, 0};
                for (int k = 0; k < 9; ++k) p[k] = q[k];
            } else {
                for (int k = 0; k < 3; ++k) p[k] = 10.0f * U(rng);
                for (int k = 3; k < 9; ++k) p[k] = p[k % 3] + U(rng);
            }
            for (int k = 0; k < 3; ++k) {
                leaf[i].mn[k] = std::fmin(std::fmin(p[k], p[3 + k]), p[6 + k]);
                leaf[i].mx[k] = std::fmax(std::fmax(p[k], p[3 + k]), p[6 + k]);
            }
        } else {
            const float r = 0.05f + 0.5f * std::fabs(U(rng));
            for (int k = 0; k < 3; ++k) { const float c = 10.0f * U(rng); leaf[i].mn[k] = c - r; leaf[i].mx[k] = c + r; }
        }
    }
    Buffer<Ref> refs;
    presplit(leaf, tri, is_tri, (size_t)((double)n * factor), refs);
    CHECK(refs.size() >= n && refs.size() <= std::max(n, (size_t)((double)n * factor)), "reference count %zu", refs.size());
    std::vector<std::vector<uint32_t>> of(n);
    for (uint32_t r = 0; r < refs.size(); ++r) {
        CHECK(refs[r].item < n, "item out of range");
        of[refs[r].item].push_back(r);
        for (int k = 0; k < 3; ++k)
            CHECK(refs[r].box.mn[k] >= leaf[refs[r].item].mn[k] && refs[r].box.mx[k] <= leaf[refs[r].item].mx[k], "ref %u leaves its element's box", r);
    }
    std::uniform_real_distribution<float> B(0.0f, 1.0f);
    for (size_t i = 0; i < n; ++i) {
        CHECK(!of[i].empty(), "element %zu has no reference", i);
        if (!is_tri[i]) { CHECK(of[i].size() == 1, "non-triangle was split"); continue; }
        const float* p = &tri[9 * i];
        for (int s = 0; s < 24; ++s) {  // points of the triangle (incl. vertices and edges) must be inside one of its reference boxes
            float u = B(rng), v = B(rng);
            if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
            if (s < 3) { u = s == 1; v = s == 2; }
            if (s == 3) v = 0.0f;
            double q[3];
            for (int k = 0; k < 3; ++k) q[k] = (double)p[k] + (double)u * ((double)p[3 + k] - p[k]) + (double)v * ((double)p[6 + k] - p[k]);
            bool in = false;
            for (uint32_t r : of[i]) {
                bool ok = true;
                for (int k = 0; k < 3; ++k) ok = ok && q[k] >= (double)refs[r].box.mn[k] - 1e-6 * (1.0 + std::fabs(q[k])) && q[k] <= (double)refs[r].box.mx[k] + 1e-6 * (1.0 + std::fabs(q[k]));
                in = in || ok;
            }
            CHECK(in, "point of triangle %zu not covered by its %zu references", i, of[i].size());
        }
    }
    std::vector<Box3> rb(refs.size());
    for (size_t i = 0; i < refs.size(); ++i) rb[i] = refs[i].box;
    uint32_t need = 0;
    while ((1ull << need) < rb.size()) ++need;
    Tree T;
    build(rb, need + (uint32_t)extra, T);
    if (rb.size() == 1) {
        CHECK(T.left.empty() && (T.root & 0x80000000u), "single element tree");
    } else {
        CHECK(T.left.size() == rb.size() - 1, "node count %zu for %zu leaves", T.left.size(), rb.size());
        CHECK(T.depth <= need + (uint32_t)extra, "depth %u over the cap %u", T.depth, need + extra);
        std::vector<uint32_t> seen(rb.size(), 0);
        std::vector<uint32_t> depth(T.left.size(), 0);
        uint32_t maxd = 0;
        CHECK(T.root == 0, "root index");
        if (!T.left.empty()) depth[0] = 1;
        for (uint32_t i = 0; i < T.left.size(); ++i) {
            Box3 u = detail::empty_box();
            for (uint32_t c : {T.left[i], T.right[i]}) {
                if (c & 0x80000000u) { const uint32_t r = c & 0x7FFFFFFFu; CHECK(r < rb.size(), "leaf out of range"); if (r < rb.size()) { ++seen[r]; detail::grow(u, rb[r]); } }
                else { CHECK(c > i && c < T.left.size(), "child %u of node %u is not after its parent", c, i); if (c > i && c < T.left.size()) { depth[c] = depth[i] + 1; detail::grow(u, T.box[c]); } }
            }
            maxd = std::max(maxd, depth[i]);
            for (int k = 0; k < 3; ++k) CHECK(u.mn[k] == T.box[i].mn[k] && u.mx[k] == T.box[i].mx[k], "node %u box is not the union of its children", i);
            CHECK(T.axis[i] < 3, "axis");
        }
        for (size_t r = 0; r < rb.size(); ++r) CHECK(seen[r] == 1, "reference %zu appears %u times", r, seen[r]);
        CHECK(maxd == T.depth, "depth bookkeeping %u vs %u", maxd, T.depth);
    }
    std::printf("%s n=%zu refs=%zu nodes=%zu depth=%u\n", fails ? "FAILED" : "OK", n, refs.size(), T.left.size(), T.depth);
    return fails ? 1 : 0;
}
