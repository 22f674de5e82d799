// b200rt.cu — C ABI (include/b200rt.h) + host runtime of the B200 render hot path.
//
// Host side: validates the flattened reference scene, repacks it into the device layout of rt_device.cuh
// (children-in-parent 64 B nodes, 64 B primitive / shading records), owns device buffers and drives the
// wavefront loop.  There is NO CPU fallback: without a CUDA device every entry point returns RT_ECUDA.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo -O3 -Xcompiler -ffp-contract=off
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200rt.h"
#include "rt_kernels.cuh"
#include "rt_rebuild.hpp"

// Device memory of scenes and of the scene hand-off's scratch goes through a small per-device cache of freed blocks (defined below):
// cudaMalloc / cudaFree go to the kernel-mode driver and serialise with everything else that does; inside bench.py's later configs
// single calls were seen to take 0.1 - 1.3 s each (profiles/README.md, round 2 late).  A caller that hands scenes over again and
// again (a shim's render() does, every call) re-uses the blocks of the scene it has just destroyed.
namespace rt_devcache {
cudaError_t alloc(void** out, size_t bytes);
void release(void* p);
void trim(int device);  // -1: every device
}  // namespace rt_devcache

#include "rt_presplit.cuh"
#include "rt_sahbuild.cuh"

using namespace rt;

namespace {

thread_local std::string g_err;
// Knobs (rt_set_option).  Writers and readers go through g_opt_mu; rt_scene_create and every render take ONE snapshot at entry
// (Options), so a concurrent rt_set_option never changes a call half-way.
std::mutex g_opt_mu;
int g_opt_traversal = 3;      // 0 = reference node test only (exact), 1 = + tight cull (exact), 2 = + reciprocal fast path,
                              // 3 = rebuilt topology + per-ray proof, unproven rays on the reference tree with 2 (plain scenes; else 2)
int g_opt_presplit = 800;     // reference budget of the rebuilt tree in per cent of the element count (<= 100: no pre-splitting)
int g_opt_presplit_large = 200;  // the same budget for scenes above 150 000 elements (capped: the accelerator tree of C4 is already > L2)
int64_t g_opt_device_cache_mb = 2048;  // freed device blocks kept per GPU for the next scene (0: every release is a cudaFree)
int g_opt_build_device = 1;     // after a device pre-split, also build the accelerator tree there (rt_sahbuild.cuh): nothing comes back to the host
int g_opt_presplit_device = 1;  // pre-split on the device (rt_presplit.cuh) for scenes of >= 16 Ki elements; 0 = on the host cores
int g_opt_rebuild_min_items = 32;  // below this the reference tree is as good (a Cornell box has 8 elements) and nothing has to be proven
int g_opt_rebuild_wrapped = 1;  // also build it for scenes with wrapped elements / volumes (true world-space bounds)
int g_opt_rebuild = 1;        // build the second topology at rt_scene_create (0: traversal 3 behaves like 2)
int g_opt_rebuild_depth_extra = 4;  // depth cap of the rebuilt tree = reference depth + this (the stack lives in shared memory)
int g_opt_pipes = 0;         // independent wavefront pipelines (streams) per render: one pipeline's extend tail overlaps the others' work.
                              // 0 = auto: 2 for renders with >= 64 Mi path slots (throughput-bound), 4 below (launch-tail-bound)
int g_opt_leaf_weight = 24;   // extend phase vote: leaf phase when n_leaf * w > n_node * 16
int g_opt_refill = 28;        // extend refills idle lanes when fewer than this many of 32 are busy
int g_opt_tail_budget = 64;   // rounds an extend warp keeps going after it can no longer refill, before it parks its rays (0 = never)
int g_opt_carveout = 75;        // cudaFuncAttributePreferredSharedMemoryCarveout of the extend kernel in per cent (-1 = driver default): the stacks get
                                // 171 KB of the SM's 228 KB, the rest stays L1 for node / primitive records (C4 +4.8 %, C2 unchanged; profiles/README.md)
int g_opt_debug_skip_redo = 0;  // measurement only: keep unproven fast-path results (films are then NOT the reference's)
int g_opt_count = 0;        // 1 = COUNT kernel variant (node visits / primitive tests per ray)
int g_opt_profile = 0;        // 1 = time every extend / shade launch with CUDA events on the launching stream (no extra syncs)
int g_opt_check_every = 8;    // host polls the live-path count every N iterations
int g_opt_blocks_per_sm = 0;  // 0 = from occupancy query
int64_t g_opt_max_slots = 160ll << 20;  // path slots in flight per render (bounds the workspace: 120 B per slot)

struct Options {
    int traversal, presplit, rebuild_min_items, rebuild_wrapped, rebuild, rebuild_depth_extra, pipes, leaf_weight, refill, count, profile,
        check_every, blocks_per_sm, debug_skip_redo, tail_budget, carveout;
    int64_t max_slots;
    int presplit_device, presplit_large, build_device;
};
Options options_snapshot() {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    return Options{g_opt_traversal, g_opt_presplit, g_opt_rebuild_min_items, g_opt_rebuild_wrapped, g_opt_rebuild, g_opt_rebuild_depth_extra,
                   g_opt_pipes, g_opt_leaf_weight, g_opt_refill, g_opt_count, g_opt_profile, g_opt_check_every, g_opt_blocks_per_sm, g_opt_debug_skip_redo, g_opt_tail_budget, g_opt_carveout,
                   g_opt_max_slots, g_opt_presplit_device, g_opt_presplit_large, g_opt_build_device};
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            return fail(RT_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
    } while (0)

}  // namespace
namespace rt_devcache {
struct Block {
    void* p;
    size_t bytes;
    int device;
};
static std::mutex g_mu;
static std::map<void*, Block> g_live;          // blocks handed out
static std::vector<Block> g_free;              // blocks kept for re-use, oldest first
static size_t g_cached[64] = {0};              // bytes kept per device
cudaError_t alloc(void** out, size_t bytes) {
    *out = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    bytes = (bytes + 511) & ~(size_t)511;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        size_t best = g_free.size();
        for (size_t i = 0; i < g_free.size(); ++i)  // smallest kept block that fits without wasting more than a quarter
            if (g_free[i].device == dev && g_free[i].bytes >= bytes && g_free[i].bytes <= bytes + bytes / 4 + ((size_t)1 << 20) &&
                (best == g_free.size() || g_free[i].bytes < g_free[best].bytes))
                best = i;
        if (best != g_free.size()) {
            const Block b = g_free[best];
            g_free.erase(g_free.begin() + (long)best);
            if (dev >= 0 && dev < 64) g_cached[dev] -= b.bytes;
            g_live[b.p] = b;
            *out = b.p;
            return cudaSuccess;
        }
    }
    void* p = nullptr;
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {  // out of memory with blocks still kept: give them back and try once more
        cudaGetLastError();
        trim(dev);
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_live[p] = Block{p, bytes, dev};
    *out = p;
    return cudaSuccess;
}
void release(void* p) {
    if (!p) return;
    std::vector<void*> to_free;
    {
        int64_t cap_mb;
        { std::lock_guard<std::mutex> ol(g_opt_mu); cap_mb = g_opt_device_cache_mb; }
        const size_t cap = (size_t)cap_mb << 20;
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_live.find(p);
        if (it == g_live.end()) { to_free.push_back(p); }
        else {
            const Block b = it->second;
            g_live.erase(it);
            if (b.bytes > cap || b.device < 0 || b.device >= 64) to_free.push_back(p);  // e.g. the 19 GB path queues: never kept
            else {
                g_free.push_back(b);
                g_cached[b.device] += b.bytes;
                for (size_t i = 0; g_cached[b.device] > cap && i < g_free.size();) {  // over the cap: the oldest go first
                    if (g_free[i].device != b.device) { ++i; continue; }
                    to_free.push_back(g_free[i].p);
                    g_cached[b.device] -= g_free[i].bytes;
                    g_free.erase(g_free.begin() + (long)i);
                }
            }
        }
    }
    for (void* q : to_free) cudaFree(q);  // (a pointer is valid on any current device)
}
void trim(int device) {
    std::vector<void*> to_free;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (size_t i = 0; i < g_free.size();) {
            if (device >= 0 && g_free[i].device != device) { ++i; continue; }
            to_free.push_back(g_free[i].p);
            if (g_free[i].device >= 0 && g_free[i].device < 64) g_cached[g_free[i].device] -= g_free[i].bytes;
            g_free.erase(g_free.begin() + (long)i);
        }
    }
    for (void* q : to_free) cudaFree(q);
}
}  // namespace rt_devcache
namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) return cudaSuccess;
        const cudaError_t e = rt_devcache::alloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();  // clear the sticky-looking last error: the next call must not report this OOM again
            return e;
        }
        n = count;
        return cudaSuccess;
    }
    void release() {
        if (p) rt_devcache::release(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

}  // namespace

// Render workspace (path-state queues, accumulators, counters): one per device, shared by all scenes of the process and
// reused across renders, so creating / destroying a scene never pays for gigabytes of cudaMalloc.  Renders on the same
// device are serialised by `mu` (rt_render* stays blocking and thread-safe).
struct Workspace {
    std::mutex mu;
    DevBuf<float4> A[2], B[2], C[2], acc;
    DevBuf<int2> D[2];
    DevBuf<float> Tm[2];
    DevBuf<uint2> hits;
    DevBuf<Counters> ctr;
    size_t ws_slots2 = 0;            // slots the second queue set (A[1]..D[1]) is sized for
    DevBuf<float> film, sumsq;
    static const int MAX_PIPES = 8;
    Counters* h_ctr = nullptr;  // pinned, MAX_PIPES x 2 entries (double-buffered poll per pipeline)
    cudaEvent_t poll_event[MAX_PIPES][2] = {};
    cudaStream_t pipe_stream[MAX_PIPES] = {};
    cudaEvent_t fork_event = nullptr, join_event[MAX_PIPES] = {};
    cudaStream_t resume_stream[MAX_PIPES] = {};  // RESUME launches of k_extend (parked traversals), beside the queue launch of the same pipeline
    cudaEvent_t resume_fork[MAX_PIPES] = {}, resume_join[MAX_PIPES] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // render brackets (created once: no per-call create / destroy to leak on error paths)
    std::vector<cudaEvent_t> prof_events[MAX_PIPES];
    DevBuf<uint32_t> iter_log;  // [MAX_PIPES][ITER_LOG_CAP], only with RT_ITER_LOG set
    DevBuf<uint2> susp;         // suspend pools of all pipelines: [pipe][2][susp_cap][4 + stack_cap]
    DevBuf<float> long_rays;    // 4096 x 8 floats, only with RT_LONG_RAYS set (debug: the rays that set a launch's tail)
    static const int ITER_LOG_CAP = 16384;
    size_t ws_slots = 0;
    // film download into a caller's pageable buffer: two pinned staging chunks, DMA of chunk i+1 overlapped with the host copy of chunk i
    static const size_t STAGE_BYTES = (size_t)16 << 20;
    char* h_stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_event[2] = {};
    cudaStream_t copy_stream = nullptr;
};
static std::mutex g_ws_mu;
static std::map<int, Workspace*> g_ws;  // never freed: lives as long as the process (the CUDA context may already be gone at exit)
static Workspace& workspace_of(int device) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    Workspace*& w = g_ws[device];
    if (!w) w = new Workspace();
    return *w;
}

struct RtScene {
    int device = 0;
    int sm_count = 0;
    SceneDev dev{};
    DevBuf<NodeDev> nodes;
    DevBuf<PrimDev> prims;
    DevBuf<ShadeDev> shade;
    DevBuf<float4> own_box;
    DevBuf<WrapDev> wraps;
    DevBuf<NodeDev> nodes2;    // row f4: rebuilt topology (plain scenes only)
    DevBuf<float4> vis_box;
    DevBuf<uint32_t> visit_rank;
    uint32_t depth2 = 0;
    bool has_animation = false;
    bool boxes_are_bounds = true;  // every node box contains its children's boxes (false: MODE 0 only, see pack_scene)
    bool has_wrapped = false, has_volume = false;  // pick the k_extend instantiation (plain / wrappers / wrappers + volumes)
    DevBuf<MaterialDev> materials;
    DevBuf<TextureDev> textures;
    std::vector<DevBuf<uint8_t>*> images;
    uint64_t device_bytes = 0;
    uint32_t depth = 0;
    ~RtScene() {
        for (auto* b : images) delete b;
    }
};

namespace {

struct Box {
    float mn[3], mx[3];
};

// Geometry::bounding_box (hittable.rs:150-157) for leaf children; used only by the tight cull.
Box leaf_box(const RtItem& it) {
    Box b;
    const float* g = it.geom;
    if (it.kind == RT_GEO_SPHERE) {  // aabb.rs:20-25
        for (int i = 0; i < 3; ++i) { b.mn[i] = g[i] - g[3]; b.mx[i] = g[i] + g[3]; }
    } else if (it.kind == RT_GEO_RECT) {  // rect_geometry.rs:87-98, thickness 0.01
        int plane = (int)g[0];
        int p0 = plane == RT_RECT_YZ ? 1 : 0, p1 = plane == RT_RECT_XY ? 1 : 2, n = plane == RT_RECT_XY ? 2 : (plane == RT_RECT_XZ ? 1 : 0);
        b.mn[p0] = g[2]; b.mx[p0] = g[3];
        b.mn[p1] = g[4]; b.mx[p1] = g[5];
        b.mn[n] = g[1] - 0.01f; b.mx[n] = g[1] + 0.01f;
    } else if (it.kind == RT_GEO_BOX) {  // Geometry::AxisAlignedBox: the box itself (hittable.rs:154)
        for (int i = 0; i < 3; ++i) { b.mn[i] = g[i]; b.mx[i] = g[3 + i]; }
    } else {  // triangle_geometry.rs:47-49
        for (int i = 0; i < 3; ++i) {
            b.mn[i] = std::fmin(std::fmin(g[i], g[3 + i]), g[6 + i]);
            b.mx[i] = std::fmax(std::fmax(g[i], g[3 + i]), g[6 + i]);
        }
    }
    return b;
}

// True world-space bounds of a wrapped element (for the rebuilt tree only; the reference's own box for it is the Q2 box):
// the local box's corners through Transformation::apply_point (transformations.rs) and, for Animation, translated by
// velocity * t at both ends of the shutter interval (linear motion: the union of the two end boxes bounds the sweep).
Box world_box(const RtItem& it, const Box& lb, double t_lo, double t_hi) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int c = 0; c < 8; ++c) {
        double p[3] = {(c & 1) ? lb.mx[0] : lb.mn[0], (c & 2) ? lb.mx[1] : lb.mn[1], (c & 4) ? lb.mx[2] : lb.mn[2]};
        if (it.wrap_flags & RT_WRAP_TRANSFORMATION) {
            const double sn = it.transform[3], cs = it.transform[4];
            const double x = cs * p[0] + sn * p[2], z = -sn * p[0] + cs * p[2];
            p[0] = x + it.transform[0]; p[1] += it.transform[1]; p[2] = z + it.transform[2];
        }
        for (double t : {t_lo, t_hi}) {
            for (int k = 0; k < 3; ++k) {
                const double q = p[k] + ((it.wrap_flags & RT_WRAP_ANIMATION) ? (double)it.velocity[k] * t : 0.0);
                mn[k] = std::min(mn[k], q); mx[k] = std::max(mx[k], q);
            }
        }
    }
    Box b;
    for (int k = 0; k < 3; ++k) {
        b.mn[k] = (float)mn[k]; if ((double)b.mn[k] > mn[k]) b.mn[k] = std::nextafter(b.mn[k], -INFINITY);
        b.mx[k] = (float)mx[k]; if ((double)b.mx[k] < mx[k]) b.mx[k] = std::nextafter(b.mx[k], INFINITY);
    }
    return b;
}

struct H3 {
    float x, y, z;
};
inline H3 hsub(H3 a, H3 b) { return H3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float hdot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline H3 hcross(H3 a, H3 b) { return H3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Per-triangle constants of TriangleGeometry::hit (triangle_geometry.rs:14-29), same op order, IEEE single.
PrimDev pack_triangle(const float* g) {
    H3 p0{g[0], g[1], g[2]}, p1{g[3], g[4], g[5]}, p2{g[6], g[7], g[8]};
    H3 dir1 = hsub(p1, p0), dir2 = hsub(p2, p0);
    H3 c = hcross(dir1, dir2);
    float len = std::sqrt(hdot(c, c));
    float s = 1.0f / len;  // Dir3::unit = self * (1.0 / length)
    H3 n{c.x * s, c.y * s, c.z * s};
    H3 vt1 = hcross(n, dir2);
    float d1 = hdot(dir1, vt1);
    H3 vt2 = hcross(n, dir1);
    float d2 = hdot(dir2, vt2);
    PrimDev p;
    p.q0 = make_float4(p0.x, p0.y, p0.z, n.x);
    p.q1 = make_float4(n.y, n.z, vt1.x, vt1.y);
    p.q2 = make_float4(vt1.z, d1, vt2.x, vt2.y);
    p.q3 = make_float4(vt2.z, d2, 0.0f, 0.0f);
    return p;
}

inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// Pageable <-> pinned copies of a staging chunk on up to four threads: one core copies at 5-6 GB/s, less than a quarter of what the DMA
// engine moves, so a single-threaded copy was what a 100 MB film download or a 128 MB node upload waited for.
static void copy_mt(void* dst, const void* src, size_t len) {
    const size_t parts = std::min<size_t>(4, len >> 21);
    if (parts < 2) { std::memcpy(dst, src, len); return; }
    std::thread th[3];
    for (size_t p = 1; p < parts; ++p)
        th[p - 1] = std::thread([=] { std::memcpy(static_cast<char*>(dst) + len * p / parts, static_cast<const char*>(src) + len * p / parts, len * (p + 1) / parts - len * p / parts); });
    std::memcpy(dst, src, len / parts);
    for (size_t p = 1; p < parts; ++p) th[p - 1].join();
}
// Pinned staging chunks of a device's workspace (created on first use; the caller holds ws.mu).
int ensure_staging(Workspace& ws) {
    if (ws.h_stage[0]) return RT_OK;
    for (int b = 0; b < 2; ++b) {
        CU(cudaMallocHost(&ws.h_stage[b], Workspace::STAGE_BYTES));
        CU(cudaEventCreateWithFlags(&ws.stage_event[b], cudaEventDisableTiming));
    }
    CU(cudaStreamCreateWithFlags(&ws.copy_stream, cudaStreamNonBlocking));
    return RT_OK;
}
// Pageable host -> device through the pinned chunks: the host copy of chunk i+1 overlaps the DMA of chunk i.  A plain cudaMemcpy
// from pageable memory ran at 1.3 GB/s here (166 ms of C4's 1.1 s scene hand-off).  Synchronous; takes the workspace lock.
int upload_staged(int device, void* d_dst, const void* src, size_t bytes) {
    if (bytes < ((size_t)1 << 20)) { CU(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice)); return RT_OK; }
    Workspace& ws = workspace_of(device);
    std::lock_guard<std::mutex> lk(ws.mu);
    int rc = ensure_staging(ws);
    if (rc) return rc;
    const size_t C = Workspace::STAGE_BYTES;
    const size_t n_chunks = (bytes + C - 1) / C;
    for (size_t i = 0; i < n_chunks; ++i) {
        const size_t off = i * C, len = std::min(C, bytes - off);
        if (i >= 2) CU(cudaEventSynchronize(ws.stage_event[i & 1]));  // the DMA that last read this chunk has finished
        copy_mt(ws.h_stage[i & 1], static_cast<const char*>(src) + off, len);
        CU(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, ws.h_stage[i & 1], len, cudaMemcpyHostToDevice, ws.copy_stream));
        CU(cudaEventRecord(ws.stage_event[i & 1], ws.copy_stream));
    }
    CU(cudaStreamSynchronize(ws.copy_stream));
    return RT_OK;
}

// Device -> pageable host through the pinned chunks (the caller holds ws.mu): the DMA of chunk i+1 overlaps the host copy of chunk i.
static int download_staged(Workspace& ws, void* dst, const void* d_src, size_t bytes) {
    {
        int rc = ensure_staging(ws);
        if (rc) return rc;
    }
    const size_t C = Workspace::STAGE_BYTES;
    const size_t n_chunks = (bytes + C - 1) / C;
    auto issue = [&](size_t i) -> int {
        const size_t off = i * C, len = std::min(C, bytes - off);
        CU(cudaMemcpyAsync(ws.h_stage[i & 1], static_cast<const char*>(d_src) + off, len, cudaMemcpyDeviceToHost, ws.copy_stream));
        CU(cudaEventRecord(ws.stage_event[i & 1], ws.copy_stream));
        return RT_OK;
    };
    if (n_chunks) { int rc = issue(0); if (rc) return rc; }
    for (size_t i = 0; i < n_chunks; ++i) {
        if (i + 1 < n_chunks) { int rc = issue(i + 1); if (rc) { cudaStreamSynchronize(ws.copy_stream); return rc; } }
        CU(cudaEventSynchronize(ws.stage_event[i & 1]));
        const size_t off = i * C, len = std::min(C, bytes - off);
        copy_mt(static_cast<char*>(dst) + off, ws.h_stage[i & 1], len);
    }
    return RT_OK;
}

// A helper thread of pack_scene that is always joined when it goes out of scope (early returns, exceptions of the packing code).
struct ScopedThread {
    std::thread t;
    std::string error;  // set by the body (exceptions included)
    bool inline_mode = false;  // small scenes: the body runs on the calling thread (starting and waking threads costs more than the work)
    template <class F>
    void start(F&& body) {
        auto run = [this, body]() {
            try { body(error); } catch (const std::exception& e) { error = e.what(); } catch (...) { error = "unknown exception"; }
        };
        if (inline_mode) run(); else t = std::thread(run);
    }
    void join() { if (t.joinable()) t.join(); }
    ~ScopedThread() { join(); }
};
// Host -> device copies of the scene arrays on a thread of their own, in the order they are posted, so that the upload of what is
// already packed (primitives, shading records, the reference tree) runs beside the accelerator build.  The staging chunks of a
// device's workspace serve one copy at a time, so every large scene upload goes through this one thread.
struct Uploader {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<int()>> jobs;
    bool closing = false;
    int rc = RT_OK;
    std::string error;
    double busy_ms = 0.0;
    std::thread t;
    bool inline_mode = false;  // small scenes: every job runs inside post()
    void run_job(std::function<int()>& job) {
        if (rc != RT_OK) return;  // after a failure the remaining jobs are dropped
        const auto t0 = std::chrono::steady_clock::now();
        int r;
        try { r = job(); } catch (const std::exception& e) { r = fail(RT_ENOMEM, e.what()); } catch (...) { r = fail(RT_EINVAL, "unknown exception"); }
        busy_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r != RT_OK) { rc = r; error = g_err; }  // g_err is thread-local: carried to the caller by finish()
    }
    void start() {
        if (inline_mode) return;
        t = std::thread([this] {
            for (;;) {
                std::function<int()> job;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return closing || !jobs.empty(); });
                    if (jobs.empty()) return;
                    job = std::move(jobs.front());
                    jobs.pop_front();
                }
                run_job(job);
            }
        });
    }
    void post(std::function<int()> job) {
        if (inline_mode) { std::lock_guard<std::mutex> lk(mu); run_job(job); return; }
        { std::lock_guard<std::mutex> lk(mu); jobs.push_back(std::move(job)); }
        cv.notify_one();
    }
    void stop() {  // waits for everything posted
        { std::lock_guard<std::mutex> lk(mu); closing = true; }
        cv.notify_one();
        if (t.joinable()) t.join();
    }
    int finish() {  // ... and reports the first failure on the calling thread
        stop();
        return rc == RT_OK ? RT_OK : fail(rc, error);
    }
    ~Uploader() { stop(); }  // early return of the packing code: its own error stays the reported one
};

int pack_scene(const RtSceneDesc& d, RtScene& sc, const Options& O) {
    // RT_PACK_LOG=1: phase times of the scene hand-off on stderr (host packing, accelerator build, upload)
    const bool pack_log = std::getenv("RT_PACK_LOG") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!pack_log) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rt_scene_create] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const uint32_t N = d.n_items;
    if (N == 0) return fail(RT_EINVAL, "scene has no items (the reference BVH builder recurses forever on an empty list)");
    if (N > RT_REF_ITEM_MASK) return fail(RT_EINVAL, "too many items");
    if (!d.items) return fail(RT_EINVAL, "items is NULL");
    if (d.n_nodes != N - 1) return fail(RT_EINVAL, "n_nodes must be n_items - 1 (reference BVH topology)");
    if (d.n_nodes && !d.nodes) return fail(RT_EINVAL, "nodes is NULL");
    if (d.n_materials == 0 || !d.materials) return fail(RT_EINVAL, "no materials");
    for (uint32_t i = 0; i < d.n_materials; ++i) {
        const RtMaterial& m = d.materials[i];
        if (m.kind > RT_MAT_ISOTROPIC) return fail(RT_EINVAL, "unknown material kind");
        if (m.kind != RT_MAT_DIELECTRIC && m.texture >= d.n_textures) return fail(RT_EINVAL, "material texture index out of range");
    }
    for (uint32_t i = 0; i < d.n_textures; ++i) {
        const RtTexture& t = d.textures[i];
        if (t.kind > RT_TEX_MARBLE) return fail(RT_EINVAL, "unknown texture kind");
        if (t.kind == RT_TEX_MARBLE && (t.image >= d.n_perlins || !d.perlins[t.image].ranvec || !d.perlins[t.image].perm_x || !d.perlins[t.image].perm_y || !d.perlins[t.image].perm_z))
            return fail(RT_EINVAL, "marble texture refers to a missing Perlin table");
        if (t.kind == RT_TEX_CHECKER && (t.even >= d.n_textures || t.odd >= d.n_textures)) return fail(RT_EINVAL, "checker texture child out of range");
        if (t.kind == RT_TEX_IMAGE && (t.image >= d.n_images || !d.images[t.image].rgb8 || d.images[t.image].width == 0 || d.images[t.image].height == 0))
            return fail(RT_EINVAL, "image texture refers to a missing image");
    }
    // a checker must resolve to a non-checker texture within 8 levels whichever way it branches (the device walks at most 8)
    for (uint32_t i = 0; i < d.n_textures; ++i) {
        if (d.textures[i].kind != RT_TEX_CHECKER) continue;
        std::vector<uint32_t> level{i};
        bool ok = false;
        for (int depth = 0; depth < 8 && !ok; ++depth) {
            std::vector<uint32_t> next;
            for (uint32_t t : level)
                for (uint32_t c : {d.textures[t].even, d.textures[t].odd})
                    if (d.textures[c].kind == RT_TEX_CHECKER && std::find(next.begin(), next.end(), c) == next.end()) next.push_back(c);
            ok = next.empty();
            level.swap(next);
        }
        if (!ok) return fail(RT_EINVAL, "checker texture nesting deeper than 8 levels (or cyclic)");
    }
    auto ref_of_item = [&](uint32_t item) {
        return RT_REF_LEAF | (d.items[item].kind << RT_REF_KIND_SHIFT) | (d.items[item].wrap_flags ? RT_REF_WRAPPED : 0u) | item;
    };
    auto conv_ref = [&](uint32_t r, uint32_t* out) -> bool {
        if (r & RT_LEAF_BIT) {
            uint32_t item = r & ~RT_LEAF_BIT;
            if (item >= N) return false;
            *out = ref_of_item(item);
        } else {
            if (r >= d.n_nodes) return false;
            *out = r;
        }
        return true;
    };

    // uninitialised buffers: first touch (page faults included) happens in the parallel loops that fill them
    using rt_rebuild::Buffer;
    Buffer<PrimDev> prims;
    Buffer<ShadeDev> shade;
    Buffer<Box> lbox;
    prims.resize_discard(N); shade.resize_discard(N); lbox.resize_discard(N);
    // Guard band of the tight cull (DESIGN.md "Tight node test"): a primitive test can report a hit for a ray that
    // misses the primitive's true AABB by the rounding slack of that test.  m(t) = A + |t| (1e-5 + C |t|) with
    //   triangles / rects: A = 8 eps max|coordinate|            (position = o + t d is rounded at coordinate scale)
    //   spheres          : A = 8 eps (max|coordinate| + 3 r), C = 8 * 3 eps / r   (|disc error| <~ 11 eps |oc|^2)
    Buffer<float> leafA, leafC, fastA, fastC;
    Buffer<Box> fbox;
    leafA.resize_discard(N); leafC.resize_discard(N); fastA.resize_discard(N); fastC.resize_discard(N); fbox.resize_discard(N);
    // ray times the camera can produce: U[t0, t1) + dot(shutter_pace, point), point in [0, 1 + 1/(w-1)]^2 (camera.rs:185-190)
    const bool times_ok = std::isfinite(d.camera.time_start) && std::isfinite(d.camera.time_end) && std::isfinite(d.camera.shutter_pace[0]) &&
                          std::isfinite(d.camera.shutter_pace[1]);
    const double pace_lo = 2.0 * (std::min(0.0, (double)d.camera.shutter_pace[0]) + std::min(0.0, (double)d.camera.shutter_pace[1]));
    const double pace_hi = 2.0 * (std::max(0.0, (double)d.camera.shutter_pace[0]) + std::max(0.0, (double)d.camera.shutter_pace[1]));
    const double shutter_lo = times_ok ? std::min<double>(d.camera.time_start, d.camera.time_end) + pace_lo : 0.0;
    const double shutter_hi = times_ok ? std::max<double>(d.camera.time_start, d.camera.time_end) + pace_hi : 0.0;
    bool any_wrap = false;
    Buffer<WrapDev> wraps;
    wraps.resize_discard(N);
    const float eps = 5.9604645e-8f, safety = 8.0f;
    std::atomic<bool> flag_wrap{false}, flag_anim{false}, flag_vol{false};
    std::atomic<const char*> first_error{nullptr};
    rt_rebuild::parallel_for(N, 16384, [&](size_t i_lo, size_t i_hi) {
        const char* item_error = nullptr;
        auto one = [&](uint32_t i) {
        const RtItem& it = d.items[i];
        if (it.kind > RT_GEO_BOX) { item_error = "unknown geometry kind"; return; }
        if (it.wrap_flags & ~7u) { item_error = "unknown wrapper flag"; return; }
        if (it.wrap_flags != 0u) flag_wrap.store(true, std::memory_order_relaxed);
        if (it.wrap_flags & RT_WRAP_ANIMATION) flag_anim.store(true, std::memory_order_relaxed);
        
        if (it.wrap_flags & RT_WRAP_VOLUME) flag_vol.store(true, std::memory_order_relaxed);
        if (it.material >= d.n_materials) { item_error = "item material index out of range"; return; }
        const float* g = it.geom;
        PrimDev p;
        p.q0 = p.q1 = p.q2 = p.q3 = make_float4(0, 0, 0, 0);
        ShadeDev s;
        s.q0 = s.q1 = s.q2 = s.q3 = make_float4(0, 0, 0, 0);
        if (it.kind == RT_GEO_SPHERE) {
            if (!(g[3] >= 0.0f)) { item_error = "sphere radius must be >= 0 (sphere_geometry.rs:13)"; return; }
            p.q0 = make_float4(g[0], g[1], g[2], g[3]);
        } else if (it.kind == RT_GEO_RECT) {
            uint32_t plane = (uint32_t)g[0];
            if (plane > RT_RECT_YZ) { item_error = "rect plane out of range"; return; }
            p.q0 = make_float4(g[1], g[2], g[3], g[4]);
            p.q1 = make_float4(g[5], u2f(plane), 0.0f, 0.0f);
        } else if (it.kind == RT_GEO_BOX) {
            p.q0 = make_float4(g[0], g[1], g[2], 0.0f);
            p.q1 = make_float4(g[3], g[4], g[5], 0.0f);
        } else {
            p = pack_triangle(g);
            s.q0 = make_float4(g[9], g[10], g[11], g[12]);
            s.q1 = make_float4(g[13], g[14], g[15], g[16]);
            s.q2 = make_float4(g[17], g[18], g[19], g[20]);
            s.q3 = make_float4(g[21], g[22], g[23], 0.0f);
        }
        s.q3.w = u2f(it.material);
        prims[i] = p;
        shade[i] = s;
        lbox[i] = leaf_box(it);
        float maxabs = 0.0f;
        for (int k = 0; k < 3; ++k) maxabs = std::max(maxabs, std::max(std::fabs(lbox[i].mn[k]), std::fabs(lbox[i].mx[k])));
        if (it.kind == RT_GEO_SPHERE) {
            leafA[i] = safety * eps * (maxabs + 3.0f * g[3]);
            leafC[i] = g[3] > 0.0f ? safety * 3.0f * eps / g[3] : INFINITY;
        } else if (it.kind == RT_GEO_TRIANGLE) {
            // The reference's triangle test is ill-conditioned for slivers (and its OBJ fan, Q4, produces them): the unit normal
            // comes from a cross product of nearly parallel edges (tilt ~ eps L^2 / 2A, i.e. a plane that misses the far vertices by
            // ~ eps kappa L), and the barycentric w0 = 1 - w1 - w2 moves the far edge by ~ eps (3L + M) (|e1| + |e2|) / |e3|
            // <= 2 eps (3L + M) kappa, with kappa = L^2 / 2A (L = longest edge, 2A = |e1 x e2|, M = coordinate magnitude).
            // So hits are accepted up to that far OUTSIDE the true triangle; the band must cover it (tests/test_guard_band_cpu.py).
            const double e1[3] = {(double)g[3] - g[0], (double)g[4] - g[1], (double)g[5] - g[2]};
            const double e2[3] = {(double)g[6] - g[0], (double)g[7] - g[1], (double)g[8] - g[2]};
            const double e3[3] = {e2[0] - e1[0], e2[1] - e1[1], e2[2] - e1[2]};
            const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            const double two_a = std::sqrt(cx * cx + cy * cy + cz * cz);
            const double l2 = std::max({e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2],
                                        e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]});
            const double kappa = two_a > 0.0 ? l2 / two_a : INFINITY;
            // generic term + the derived sliver bound eps kappa (3M + 10L); the worst slack observed over 1.4 M adversarial grazing rays
            // on 473 slivers (kappa 30 .. 5e4) is 0.06 of that bound (tests/test_guard_band_cpu.py keeps checking it)
            leafA[i] = (float)(safety * eps * (double)maxabs + eps * kappa * (3.0 * (double)maxabs + 10.0 * std::sqrt(l2)));
            if (!(leafA[i] >= 0.0f)) leafA[i] = INFINITY;  // degenerate triangle (NaN): no tight cull around it
            leafC[i] = 0.0f;
        } else {
            leafA[i] = safety * eps * maxabs;
            leafC[i] = 0.0f;
        }
        WrapDev w{};
        w.a = make_float4(it.transform[0], it.transform[1], it.transform[2], it.transform[3]);
        w.b = make_float4(it.transform[4], it.velocity[0], it.velocity[1], it.velocity[2]);
        w.flags = it.wrap_flags;
        w.neg_inv_density = it.neg_inv_density;
        wraps[i] = w;
        // Q2: the reference does not transform / motion-extend the AABB of a wrapped element, so its box (and every ancestor's)
        // is NOT a true bound: no tight cull there, the reference's weak per-axis rule alone decides (guard band = +inf)
        fastA[i] = leafA[i]; fastC[i] = leafC[i]; fbox[i] = lbox[i];
        if (it.wrap_flags & 3u) {
            // the rebuilt tree (row f4) bounds the element where it really is; its band adds the rounding of the ray's reverse
            // transform (rotation + offsets at world / local coordinate scale)
            fbox[i] = world_box(it, lbox[i], shutter_lo, shutter_hi);
            float wabs = maxabs;
            for (int k = 0; k < 3; ++k) wabs = std::max(wabs, std::max(std::fabs(fbox[i].mn[k]), std::fabs(fbox[i].mx[k])));
            fastA[i] += 16.0f * eps * wabs;
            leafA[i] = INFINITY; leafC[i] = 0.0f;
        }
        };
        for (size_t i = i_lo; i < i_hi && !item_error; ++i) one((uint32_t)i);
        if (item_error) { const char* expect = nullptr; first_error.compare_exchange_strong(expect, item_error); }
    });
    if (first_error.load()) return fail(RT_EINVAL, first_error.load());
    any_wrap = flag_wrap.load();
    sc.has_animation = flag_anim.load();
    sc.has_wrapped = any_wrap;
    sc.has_volume = flag_vol.load();
    lap("validate + pack items");
    auto bf16_up = [](float v) -> uint32_t {  // smallest bf16 >= v (v >= 0)
        uint32_t b;
        std::memcpy(&b, &v, 4);
        if (b & 0xFFFFu) b = (b | 0xFFFFu) + 1u;
        return b >> 16;
    };
    auto pack_margin = [&](float a, float c) -> uint32_t { return (bf16_up(a) << 16) | bf16_up(c); };

    // every child reference of the caller's tree, range-checked once (the passes below, several of them on helper threads, index with them)
    {
        std::atomic<const char*> node_error{nullptr};
        rt_rebuild::parallel_for(d.n_nodes, 65536, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                uint32_t tmp;
                if (!conv_ref(d.nodes[i].left, &tmp) || !conv_ref(d.nodes[i].right, &tmp)) { node_error.store("BVH child reference out of range"); return; }
                if (d.nodes[i].axis_id > 2) { node_error.store("BVH axis_id out of range"); return; }
            }
        });
        if (node_error.load()) return fail(RT_EINVAL, node_error.load());
    }
    uint32_t root;
    if (!conv_ref(d.root, &root)) return fail(RT_EINVAL, "root reference out of range");

    // everything the helper threads write (declared before them: they are joined before these buffers go away)
    Buffer<NodeDev> nodes;
    Buffer<float4> own_box;
    Buffer<float> nodeA, nodeC;
    nodes.resize_discard(d.n_nodes); own_box.resize_discard(2 * (size_t)d.n_nodes); nodeA.resize_discard(d.n_nodes); nodeC.resize_discard(d.n_nodes);
    Buffer<NodeDev> nodes2;
    Buffer<float4> vis_box;
    Buffer<uint32_t> visit_rank;
    double ms_ref_nodes = 0.0, ms_vis = 0.0, ms_ranks = 0.0;
    auto ms_since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };

    // Uploads start now, beside the rest of the packing: primitives and shading records first.
#define UP(dst, src, bytes) do { int rc_up_ = upload_staged(sc.device, (dst), (src), (bytes)); if (rc_up_) return rc_up_; } while (0)
    const bool threaded = N >= 16384u;  // below: everything on this thread, in the same order
    Uploader up;
    up.inline_mode = !threaded;
    up.start();
    up.post([&]() -> int {
        CU(cudaSetDevice(sc.device));
        {   // one attribute, cached per device (cudaGetDeviceProperties walks every attribute and can take tens of milliseconds)
            static std::mutex mu;
            static std::map<int, int> sm_of;
            std::lock_guard<std::mutex> lk(mu);
            auto it = sm_of.find(sc.device);
            if (it == sm_of.end()) {
                int v = 0;
                CU(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, sc.device));
                it = sm_of.emplace(sc.device, v).first;
            }
            sc.sm_count = it->second;
        }
        return RT_OK;
    });
    auto post_item_uploads = [&]() {  // primitives and shading records (after the inputs of a device pre-split, which the build waits for)
        up.post([&]() -> int {
            CU(sc.prims.alloc(N));
            CU(sc.shade.alloc(N));
            if (any_wrap) {
                CU(sc.wraps.alloc(N));
                UP(sc.wraps.p, wraps.data(), N * sizeof(WrapDev));
                sc.device_bytes += (uint64_t)N * sizeof(WrapDev);
            }
            UP(sc.prims.p, prims.data(), N * sizeof(PrimDev));
            UP(sc.shade.p, shade.data(), N * sizeof(ShadeDev));
            sc.device_bytes += (uint64_t)N * (sizeof(PrimDev) + sizeof(ShadeDev));
            return RT_OK;
        });
    };

    auto child_box = [&](uint32_t r) -> Box {
        if (r & RT_LEAF_BIT) return lbox[r & ~RT_LEAF_BIT];
        Box b;
        for (int k = 0; k < 3; ++k) { b.mn[k] = d.nodes[r].aabb_min[k]; b.mx[k] = d.nodes[r].aabb_max[k]; }
        return b;
    };
    // Helper 1: the reference tree's device records.  Subtree maxima of the guard band bottom-up (children always have larger
    // indices than their parent in the pre-order numbering; verified, otherwise fall back to the global maximum), the 64-byte
    // node records, each node's own box and its parent link; then the upload of both arrays is posted.
    ScopedThread h_ref;
    h_ref.inline_mode = !threaded;
    h_ref.start([&](std::string&) {
        const auto t0 = std::chrono::steady_clock::now();
        bool preorder = true;
        for (uint32_t i = 0; i < d.n_nodes && preorder; ++i)
            for (uint32_t c : {d.nodes[i].left, d.nodes[i].right})
                if (!(c & RT_LEAF_BIT) && c <= i) preorder = false;
        if (preorder) {
            for (uint32_t ii = d.n_nodes; ii-- > 0;) {
                float a = 0.0f, c = 0.0f;
                for (uint32_t ch : {d.nodes[ii].left, d.nodes[ii].right}) {
                    if (ch & RT_LEAF_BIT) { const uint32_t item = ch & ~RT_LEAF_BIT; a = std::max(a, leafA[item]); c = std::max(c, leafC[item]); }
                    else { a = std::max(a, nodeA[ch]); c = std::max(c, nodeC[ch]); }
                }
                nodeA[ii] = a; nodeC[ii] = c;
            }
        } else {
            float a = 0.0f, c = 0.0f;
            for (uint32_t i = 0; i < N; ++i) { a = std::max(a, leafA[i]); c = std::max(c, leafC[i]); }
            std::fill(nodeA.begin(), nodeA.end(), a);
            std::fill(nodeC.begin(), nodeC.end(), c);
        }
        rt_rebuild::parallel_for(d.n_nodes, 32768, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const RtBvhNode& n = d.nodes[i];
                uint32_t l = 0, r = 0;
                conv_ref(n.left, &l); conv_ref(n.right, &r);  // range-checked above
                Box bl = child_box(n.left), br = child_box(n.right);
                NodeDev nd;
                nd.q0 = make_float4(bl.mn[0], bl.mn[1], bl.mn[2], bl.mx[0]);
                nd.q1 = make_float4(bl.mx[1], bl.mx[2], br.mn[0], br.mn[1]);
                nd.q2 = make_float4(br.mn[2], br.mx[0], br.mx[1], br.mx[2]);
                nd.q3 = make_uint4(l, r, n.axis_id, pack_margin(nodeA[i], nodeC[i]));
                nodes[i] = nd;
                // .w of the first half: the node's parent in the reference tree (ref_reaches_exact walks the chain); filled below
                own_box[2 * i] = make_float4(n.aabb_min[0], n.aabb_min[1], n.aabb_min[2], u2f(0xFFFFFFFFu));
                own_box[2 * i + 1] = make_float4(n.aabb_max[0], n.aabb_max[1], n.aabb_max[2], 0.0f);
            }
        });
        for (uint32_t i = 0; i < d.n_nodes; ++i)  // parent links
            for (uint32_t ch : {d.nodes[i].left, d.nodes[i].right})
                if (!(ch & RT_LEAF_BIT)) own_box[2 * (size_t)ch].w = u2f(i);
        ms_ref_nodes = ms_since(t0);
        up.post([&]() -> int {
            CU(sc.nodes.alloc(std::max<size_t>(nodes.size(), 1)));
            CU(sc.own_box.alloc(std::max<size_t>(own_box.size(), 2)));
            if (!nodes.empty()) UP(sc.nodes.p, nodes.data(), nodes.size() * sizeof(NodeDev));
            if (!nodes.empty()) UP(sc.own_box.p, own_box.data(), own_box.size() * sizeof(float4));
            sc.device_bytes += nodes.size() * (sizeof(NodeDev) + 2 * sizeof(float4));
            return RT_OK;
        });
    });

    // Helper 0: depth of the tree (iterative DFS; also rejects cycles by bounding the visit count).  Sequential, and only the
    // accelerator build needs the number: it runs beside the pre-split.  The visit-rank walks below wait for its verdict first
    // (they would not terminate on a cyclic graph).
    uint32_t depth = 0;
    std::promise<bool> depth_promise;
    std::shared_future<bool> tree_ok = depth_promise.get_future().share();
    ScopedThread h_depth;
    h_depth.inline_mode = !threaded;
    h_depth.start([&](std::string& err) {
        bool ok = false;
        try {
            ok = true;
            if (!(root & RT_REF_LEAF)) {
                std::vector<std::pair<uint32_t, uint32_t>> st;
                st.push_back({root, 1u});
                size_t visited = 0;
                while (!st.empty() && ok) {
                    auto [nd, dep] = st.back();
                    st.pop_back();
                    if (++visited > (size_t)d.n_nodes) { ok = false; break; }
                    depth = std::max(depth, dep);
                    for (uint32_t c : {d.nodes[nd].left, d.nodes[nd].right})
                        if (!(c & RT_LEAF_BIT)) st.push_back({c, dep + 1});
                }
            }
        } catch (const std::exception& e) { err = e.what(); ok = false; }
        depth_promise.set_value(ok);
    });

    // The tight cull (MODE >= 1) and the visibility proof of MODE 3 rely on what the reference builder guarantees: a node's box
    // contains the boxes of its children (it is their union).  A caller-made tree that breaks this is still traversed exactly
    // as given, but only with the reference's own node rule (MODE 0) and without the rebuilt topology.
    {
        std::atomic<bool> bounds_ok{true};
        rt_rebuild::parallel_for(d.n_nodes, 65536, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi && bounds_ok.load(std::memory_order_relaxed); ++i)
                for (uint32_t ch : {d.nodes[i].left, d.nodes[i].right}) {
                    if ((ch & RT_LEAF_BIT) && d.items[ch & ~RT_LEAF_BIT].wrap_flags) continue;  // Q2 boxes: no tight cull around them anyway
                    const Box cb = child_box(ch);
                    for (int k = 0; k < 3; ++k)
                        if (!(cb.mn[k] >= d.nodes[i].aabb_min[k] && cb.mx[k] <= d.nodes[i].aabb_max[k])) bounds_ok.store(false, std::memory_order_relaxed);
                }
        });
        sc.boxes_are_bounds = bounds_ok.load();
    }
    lap("bounds check");

    // Row f4: second topology (SAH) over the same elements.  Wrapped elements enter it with their true world-space bounds
    // (fbox; the reference's own boxes for them are not bounds, Q2); a volume's draw is keyed by the element, so it does not
    // depend on the tree that reaches it.
    uint32_t root2 = 0;
    const bool rebuild = N >= (uint32_t)std::max(2, O.rebuild_min_items) && O.rebuild && sc.boxes_are_bounds && (!any_wrap || (O.rebuild_wrapped && times_ok));
    const bool want_ranks = rebuild && !any_wrap && !(root & RT_REF_LEAF);
    // Helper 2: the reference parent's box of every element — the visibility proof of a fast-path hit (ref_visible).
    // Helper 3: equal-t ties on the fast path (plain scenes): the reference returns the element its ordered DFS visits first, and
    // that order depends only on the signs of the ray direction (hittable.rs:441-445) — one leaf order per octant, 8 x n_items ranks.
    ScopedThread h_vis, h_ranks;
    h_vis.inline_mode = h_ranks.inline_mode = !threaded;
    if (rebuild) {
        vis_box.resize_discard(2 * (size_t)N);
        h_vis.start([&](std::string&) {
            const auto t0 = std::chrono::steady_clock::now();
            rt_rebuild::parallel_for(vis_box.size(), 1 << 18, [&](size_t lo, size_t hi) { std::fill(vis_box.data() + lo, vis_box.data() + hi, make_float4(0, 0, 0, 0)); });
            rt_rebuild::parallel_for(d.n_nodes, 65536, [&](size_t lo, size_t hi) {  // every element has one parent: the writes do not collide
                for (size_t i = lo; i < hi; ++i)
                    for (uint32_t ch : {d.nodes[i].left, d.nodes[i].right})
                        if (ch & RT_LEAF_BIT) {
                            const uint32_t item = ch & ~RT_LEAF_BIT;
                            vis_box[2 * (size_t)item] = make_float4(d.nodes[i].aabb_min[0], d.nodes[i].aabb_min[1], d.nodes[i].aabb_min[2], u2f((uint32_t)i));  // .w: the parent node
                            vis_box[2 * (size_t)item + 1] = make_float4(d.nodes[i].aabb_max[0], d.nodes[i].aabb_max[1], d.nodes[i].aabb_max[2], 0.0f);
                        }
            });
            ms_vis = ms_since(t0);
            up.post([&]() -> int {
                CU(sc.vis_box.alloc(vis_box.size()));
                UP(sc.vis_box.p, vis_box.data(), vis_box.size() * sizeof(float4));
                sc.device_bytes += vis_box.size() * sizeof(float4);
                return RT_OK;
            });
        });
    }
    if (want_ranks) {
        visit_rank.resize_discard(8 * (size_t)N);
        h_ranks.start([&](std::string&) {
            if (!tree_ok.get()) return;
            const auto t0 = std::chrono::steady_clock::now();
            rt_rebuild::parallel_for(visit_rank.size(), 1 << 20, [&](size_t lo, size_t hi) { std::fill(visit_rank.data() + lo, visit_rank.data() + hi, 0u); });
            // octants 0..3 by an ordered DFS each; the complementary octant (7 - o) flips the child order at EVERY node, i.e. visits the
            // leaves in exactly the reverse order: rank = N - 1 - rank
            rt_rebuild::parallel_for(4, 1, [&](size_t o_lo, size_t o_hi) {
                std::vector<uint32_t> st;
                for (size_t oct = o_lo; oct < o_hi; ++oct) {
                    uint32_t next = 0;
                    st.clear();
                    st.push_back(d.root);
                    uint32_t* fwd = visit_rank.data() + oct * N;
                    uint32_t* rev = visit_rank.data() + (7 - oct) * N;
                    while (!st.empty()) {
                        const uint32_t r = st.back();
                        st.pop_back();
                        if (r & RT_LEAF_BIT) {
                            const uint32_t item = r & ~RT_LEAF_BIT;
                            fwd[item] = next; rev[item] = N - 1u - next;
                            ++next;
                            continue;
                        }
                        const RtBvhNode& n = d.nodes[r];
                        const bool left_first = ((oct >> n.axis_id) & 1u) != 0u;
                        st.push_back(left_first ? n.right : n.left);  // visited second
                        st.push_back(left_first ? n.left : n.right);
                    }
                }
            });
            ms_ranks = ms_since(t0);
            up.post([&]() -> int {
                CU(sc.visit_rank.alloc(visit_rank.size()));
                UP(sc.visit_rank.p, visit_rank.data(), visit_rank.size() * sizeof(uint32_t));
                sc.device_bytes += visit_rank.size() * sizeof(uint32_t);
                return RT_OK;
            });
        });
    }
    if (rebuild) {
        Buffer<rt_rebuild::Box3> lb;
        lb.resize_discard(N);
        Buffer<float> tri;
        Buffer<uint8_t> is_tri;
        std::atomic<bool> finite_boxes{true};  // the device builders order floats by their bits: anything else stays on the host path
        rt_rebuild::parallel_for(N, 65536, [&](size_t lo, size_t hi) {
            bool fin = true;
            for (size_t i = lo; i < hi; ++i)
                for (int k = 0; k < 3; ++k) {
                    lb[i].mn[k] = fbox[i].mn[k]; lb[i].mx[k] = fbox[i].mx[k];
                    fin = fin && std::isfinite(lb[i].mn[k]) && std::isfinite(lb[i].mx[k]);
                }
            if (!fin) finite_boxes.store(false, std::memory_order_relaxed);
        });
        // pre-split sliver triangles into several references (rt_rebuild.hpp); the budget shrinks for very large scenes
        const double factor = O.presplit <= 100 ? 1.0 : std::min((double)O.presplit / 100.0, N <= 150000u ? 16.0 : (N <= 2000000u ? (double)O.presplit_large / 100.0 : 1.0));
        if (factor > 1.0) {
            tri.resize_discard(9 * (size_t)N);
            is_tri.resize_discard(N);
            rt_rebuild::parallel_for(N, 65536, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) {
                    const bool t = d.items[i].kind == RT_GEO_TRIANGLE && !d.items[i].wrap_flags;  // wrapped triangles keep one reference
                    is_tri[i] = t ? 1 : 0;
                    for (int k = 0; k < 9; ++k) tri[9 * i + k] = t ? d.items[i].geom[k] : 0.0f;
                }
            });
        }
        Buffer<rt_rebuild::Ref> refs;
        const size_t budget = (size_t)((double)N * factor);
        const bool on_device = factor > 1.0 && budget > N && O.presplit_device && threaded && finite_boxes.load() && !std::getenv("RT_PACK_DRYRUN");
        const bool build_on_device = on_device && O.build_device;
        bool built = false;
        const uint32_t depth_extra = (uint32_t)std::max(0, O.rebuild_depth_extra);
        if (on_device) {
            // rt_presplit.cuh: the elements' boxes and triangles go up first; the references come back for the host builder, or stay
            // for the device one (rt_sahbuild.cuh), which also needs the per-element band coefficients and leaf references
            DevBuf<rt_rebuild::Box3> d_lb;
            DevBuf<float> d_tri, d_a, d_c;
            DevBuf<uint8_t> d_is;
            DevBuf<uint32_t> d_iref;
            DevBuf<rt_rebuild::Ref> d_refs;
            Buffer<uint32_t> iref;
            if (build_on_device) {
                iref.resize_discard(N);
                rt_rebuild::parallel_for(N, 65536, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; ++i) iref[i] = ref_of_item((uint32_t)i); });
            }
            struct Signal {  // set when the job ends, whichever way
                std::promise<int>& p;
                int rc = RT_ECUDA;
                ~Signal() { p.set_value(rc); }
            };
            std::promise<int> in_done;
            std::future<int> in_rc = in_done.get_future();
            up.post([&]() -> int {
                Signal sig{in_done};
                CU(d_lb.alloc(N)); CU(d_tri.alloc(9 * (size_t)N)); CU(d_is.alloc(N)); CU(d_refs.alloc(budget));
                UP(d_lb.p, lb.data(), (size_t)N * sizeof(rt_rebuild::Box3));
                UP(d_tri.p, tri.data(), 9 * (size_t)N * sizeof(float));
                UP(d_is.p, is_tri.data(), N);
                if (build_on_device) {
                    CU(d_a.alloc(N)); CU(d_c.alloc(N)); CU(d_iref.alloc(N));
                    UP(d_a.p, fastA.data(), (size_t)N * 4);
                    UP(d_c.p, fastC.data(), (size_t)N * 4);
                    UP(d_iref.p, iref.data(), (size_t)N * 4);
                }
                sig.rc = RT_OK;
                return RT_OK;
            });
            post_item_uploads();
            const bool in_ok = in_rc.get() == RT_OK;
            CU(cudaSetDevice(sc.device));  // (this thread frees the DevBufs above when the block ends)
            if (!in_ok) return up.finish();
            cudaStream_t st = nullptr;
            CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            size_t n_refs = 0;
            cudaError_t e = rt_presplit::run(d_lb.p, d_tri.p, d_is.p, N, budget, d_refs.p, &n_refs, st);
            if (e != cudaSuccess) { cudaStreamDestroy(st); return fail(RT_ECUDA, std::string("device pre-split: ") + cudaGetErrorString(e)); }
            lap("pre-split references (device)");
            if (build_on_device && n_refs > rt_sahbuild::SMALL) {
                if (!tree_ok.get()) { cudaStreamDestroy(st); return fail(RT_EINVAL, h_depth.error.empty() ? "BVH is not a tree" : "scene packing: " + h_depth.error); }
                uint32_t extra_levels = 0;
                while (((size_t)N << extra_levels) < n_refs) ++extra_levels;
                if (sc.nodes2.alloc(n_refs - 1) != cudaSuccess) { cudaStreamDestroy(st); return fail(RT_ENOMEM, "out of device memory (accelerator tree)"); }
                rt_sahbuild::Result res;
                e = rt_sahbuild::run(d_refs.p, (uint32_t)n_refs, d_a.p, d_c.p, d_iref.p, depth + extra_levels + depth_extra, sc.nodes2.p, &res, st);
                cudaStreamDestroy(st);
                if (e != cudaSuccess) return fail(RT_ECUDA, std::string("device accelerator build: ") + cudaGetErrorString(e));
                sc.device_bytes += (uint64_t)(n_refs - 1) * sizeof(NodeDev);
                root2 = res.root;
                sc.depth2 = res.depth;
                built = true;
                lap("SAH build (device)");
            } else {
                cudaStreamDestroy(st);
                refs.resize_discard(n_refs);
                std::promise<int> out_done;
                std::future<int> out_rc = out_done.get_future();
                up.post([&]() -> int {
                    Signal sig{out_done};
                    Workspace& ws = workspace_of(sc.device);
                    std::lock_guard<std::mutex> lk(ws.mu);
                    const int rc = download_staged(ws, refs.data(), d_refs.p, n_refs * sizeof(rt_rebuild::Ref));
                    sig.rc = rc;
                    return rc;
                });
                if (out_rc.get() != RT_OK) return up.finish();
                lap("references back to the host");
            }
        } else {
            post_item_uploads();
            rt_rebuild::presplit(lb.data(), N, tri.data(), is_tri.data(), budget, refs);
            lap("pre-split references");
        }
        if (!built) {
        if (!tree_ok.get()) return fail(RT_EINVAL, h_depth.error.empty() ? "BVH is not a tree" : "scene packing: " + h_depth.error);
        rt_rebuild::Tree F;
        uint32_t extra_levels = 0;
        while (((size_t)N << extra_levels) < refs.size()) ++extra_levels;
        // the builder also collects, bottom-up, each node's maxima of the two guard-band coefficients of the elements below it
        rt_rebuild::build(refs, fastA.data(), fastC.data(), depth + extra_levels + depth_extra, F);
        lap("SAH build");
        const uint32_t n2 = (uint32_t)F.left.size();
        auto item_of = [&](uint32_t leafref) -> uint32_t { return refs[leafref & ~RT_LEAF_BIT].item; };
        nodes2.resize_discard(n2);
        auto box2 = [&](uint32_t r) -> rt_rebuild::Box3 { return (r & RT_LEAF_BIT) ? refs[r & ~RT_LEAF_BIT].box : F.box[r]; };
        auto ref2 = [&](uint32_t r) -> uint32_t { return (r & RT_LEAF_BIT) ? ref_of_item(item_of(r)) : r; };
        // packed and uploaded in slices: the copy of one slice runs while the next is being packed
        if (n2) up.post([&]() -> int { CU(sc.nodes2.alloc(nodes2.size())); sc.device_bytes += nodes2.size() * sizeof(NodeDev); return RT_OK; });
        const size_t SLICE = (size_t)1 << 19;  // 32 MB of node records
        for (size_t s0 = 0; s0 < n2; s0 += SLICE) {
            const size_t s1 = std::min<size_t>(n2, s0 + SLICE);
            rt_rebuild::parallel_for(s1 - s0, 32768, [&](size_t lo, size_t hi) {
                for (size_t i = s0 + lo; i < s0 + hi; ++i) {
                    const rt_rebuild::Box3 bl = box2(F.left[i]), br = box2(F.right[i]);
                    NodeDev nd;
                    nd.q0 = make_float4(bl.mn[0], bl.mn[1], bl.mn[2], bl.mx[0]);
                    nd.q1 = make_float4(bl.mx[1], bl.mx[2], br.mn[0], br.mn[1]);
                    nd.q2 = make_float4(br.mn[2], br.mx[0], br.mx[1], br.mx[2]);
                    nd.q3 = make_uint4(ref2(F.left[i]), ref2(F.right[i]), F.axis[i], pack_margin(F.band_a[i], F.band_c[i]));
                    nodes2[i] = nd;
                }
            });
            up.post([&, s0, s1]() -> int { UP(sc.nodes2.p + s0, nodes2.data() + s0, (s1 - s0) * sizeof(NodeDev)); return RT_OK; });
        }
        root2 = ref2(F.root);
        sc.depth2 = F.depth;
        lap("pack accelerator nodes");
        }

    }
    if (!tree_ok.get()) return fail(RT_EINVAL, h_depth.error.empty() ? "BVH is not a tree" : "scene packing: " + h_depth.error);
    sc.depth = depth;
    if (!rebuild) post_item_uploads();
    h_ref.join();
    h_vis.join();
    h_ranks.join();
    for (const std::string* e : {&h_ref.error, &h_vis.error, &h_ranks.error})
        if (!e->empty()) return fail(RT_ENOMEM, "scene packing: " + *e);
    {
        const int rc_up = up.finish();
        lap("helpers + uploads (tail)");
        if (pack_log)
            std::fprintf(stderr, "[rt_scene_create]   beside the build: reference nodes %.1f ms, parent boxes %.1f ms, visit ranks %.1f ms, uploads %.1f ms\n",
                         ms_ref_nodes, ms_vis, ms_ranks, up.busy_ms);
        if (rc_up != RT_OK) return rc_up;
    }
    SceneDev& S = sc.dev;
    std::memset(&S, 0, sizeof(S));
    S.root = root;
    S.n_items = N;
    if (!(root & RT_REF_LEAF)) {
        for (int k = 0; k < 3; ++k) {
            S.root_min[k] = d.nodes[root].aabb_min[k];
            S.root_max[k] = d.nodes[root].aabb_max[k];
        }
        S.root_margin = pack_margin(nodeA[root], nodeC[root]);
    }
    S.stack_cap = std::max(depth, sc.depth2) + 2;
    S.root2 = root2;

    const RtCamera& c = d.camera;
    S.cam.position = make_float3(c.position[0], c.position[1], c.position[2]);
    S.cam.ulc = make_float3(c.upper_left_corner[0], c.upper_left_corner[1], c.upper_left_corner[2]);
    S.cam.unit_right = make_float3(c.unit_right[0], c.unit_right[1], c.unit_right[2]);
    S.cam.unit_up = make_float3(c.unit_up[0], c.unit_up[1], c.unit_up[2]);
    S.cam.scaled_right = make_float3(c.scaled_right[0], c.scaled_right[1], c.scaled_right[2]);
    S.cam.scaled_up = make_float3(c.scaled_up[0], c.scaled_up[1], c.scaled_up[2]);
    S.cam.lens_radius = c.lens_radius;
    S.cam.time_start = c.time_start;
    S.cam.time_end = c.time_end;
    S.cam.pace_x = c.shutter_pace[0];
    S.cam.pace_y = c.shutter_pace[1];
    S.bg_is_sky = d.background_is_sky ? 1u : 0u;
    for (int k = 0; k < 3; ++k) S.bg[k] = d.background_rgb[k];
    S.has_light = d.has_light ? 1u : 0u;
    if (d.has_light && d.light_plane > RT_RECT_YZ) return fail(RT_EINVAL, "light plane out of range");
    S.light_plane = d.light_plane;
    S.light_dist = d.light_dist;
    S.light_r0lo = d.light_r0[0]; S.light_r0hi = d.light_r0[1];
    S.light_r1lo = d.light_r1[0]; S.light_r1hi = d.light_r1[1];

    // the large arrays are on the device already (Uploader above); the small tables follow from this thread
    CU(cudaSetDevice(sc.device));

    std::vector<const uint8_t*> img_dev(d.n_images, nullptr);
    for (uint32_t i = 0; i < d.n_images; ++i) {
        size_t bytes = (size_t)d.images[i].width * d.images[i].height * 3;
        auto* b = new DevBuf<uint8_t>();
        sc.images.push_back(b);
        CU(b->alloc(bytes));
        CU(cudaMemcpy(b->p, d.images[i].rgb8, bytes, cudaMemcpyHostToDevice));
        img_dev[i] = b->p;
        sc.device_bytes += bytes;
    }
    // Perlin tables: 768 floats (gradients) + 768 u32 (perm_x | perm_y | perm_z) per table, one device buffer each
    std::vector<const float*> perlin_dev(d.n_perlins, nullptr);
    for (uint32_t i = 0; i < d.n_perlins; ++i) {
        std::vector<float> blk(1536);
        std::memcpy(blk.data(), d.perlins[i].ranvec, 768 * sizeof(float));
        std::memcpy(blk.data() + 768, d.perlins[i].perm_x, 256 * 4);
        std::memcpy(blk.data() + 1024, d.perlins[i].perm_y, 256 * 4);
        std::memcpy(blk.data() + 1280, d.perlins[i].perm_z, 256 * 4);
        auto* b = new DevBuf<uint8_t>();
        sc.images.push_back(b);
        CU(b->alloc(1536 * 4));
        CU(cudaMemcpy(b->p, blk.data(), 1536 * 4, cudaMemcpyHostToDevice));
        perlin_dev[i] = reinterpret_cast<const float*>(b->p);
        sc.device_bytes += 1536 * 4;
    }
    std::vector<MaterialDev> mats(d.n_materials);
    for (uint32_t i = 0; i < d.n_materials; ++i) mats[i] = MaterialDev{d.materials[i].kind, d.materials[i].kind == RT_MAT_DIELECTRIC ? 0u : d.materials[i].texture, d.materials[i].param, 0u};
    std::vector<TextureDev> texs(std::max<uint32_t>(d.n_textures, 1));
    std::memset(texs.data(), 0, texs.size() * sizeof(TextureDev));
    for (uint32_t i = 0; i < d.n_textures; ++i) {
        TextureDev t{};
        t.kind = d.textures[i].kind;
        t.r = d.textures[i].rgb[0]; t.g = d.textures[i].rgb[1]; t.b = d.textures[i].rgb[2];
        if (t.kind == RT_TEX_IMAGE) {
            t.width = d.images[d.textures[i].image].width;
            t.height = d.images[d.textures[i].image].height;
            t.texels = img_dev[d.textures[i].image];
        }
        if (t.kind == RT_TEX_CHECKER) {
            t.inv_frequency = d.textures[i].inv_frequency;
            t.even = d.textures[i].even;
            t.odd = d.textures[i].odd;
        }
        if (t.kind == RT_TEX_MARBLE) {
            t.inv_frequency = d.textures[i].inv_frequency;  // scale
            t.ranvec = perlin_dev[d.textures[i].image];
            t.perm = reinterpret_cast<const uint32_t*>(perlin_dev[d.textures[i].image] + 768);
        }
        texs[i] = t;
    }
    CU(sc.materials.alloc(mats.size()));
    CU(sc.textures.alloc(texs.size()));
    CU(cudaMemcpy(sc.materials.p, mats.data(), mats.size() * sizeof(MaterialDev), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc.textures.p, texs.data(), texs.size() * sizeof(TextureDev), cudaMemcpyHostToDevice));
    sc.device_bytes += mats.size() * sizeof(MaterialDev) + texs.size() * sizeof(TextureDev);
    S.nodes = sc.nodes.p;
    S.prims = sc.prims.p;
    S.shade = sc.shade.p;
    S.own_box = sc.own_box.p;
    S.nodes2 = sc.nodes2.p;
    S.fast_trace_ok = sc.has_animation ? 0u : 1u;
    S.visit_rank = sc.visit_rank.p;
    S.vis_box = sc.vis_box.p;
    S.wraps = sc.wraps.p;
    S.trace_time = 0.0f;
    S.materials = sc.materials.p;
    S.textures = sc.textures.p;
#undef UP
    lap("upload");
    return RT_OK;
}

size_t stack_smem_bytes(const RtScene& sc, int block) { return (size_t)sc.dev.stack_cap * block * 8; }

template <class K>
int persistent_grid(const RtScene& sc, K kernel, int block, size_t smem, int* out_grid, int blocks_per_sm_cap = 0) {
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) return fail(RT_ECUDA, "kernel does not fit on an SM (BVH too deep for the shared-memory stack)");
    if (blocks_per_sm_cap > 0) per_sm = std::min(per_sm, blocks_per_sm_cap);
    *out_grid = per_sm * sc.sm_count;
    return RT_OK;
}

// Grows the per-device workspace.  On any allocation failure every queue buffer is released and the sizes reset, so a later
// (smaller) render never runs on a half-allocated workspace.
void release_workspace_queues(Workspace& w) {
    for (int q = 0; q < 2; ++q) { w.A[q].release(); w.B[q].release(); w.C[q].release(); w.D[q].release(); w.Tm[q].release(); }
    w.hits.release();
    w.acc.release();
    w.ws_slots = 0;
    w.ws_slots2 = 0;
}
int ensure_workspace(Workspace& sc, size_t slots, size_t acc_elems, size_t npix, bool need_time, bool two_queues) {
    cudaError_t e = cudaSuccess;
    auto grow = [&](auto& buf, size_t count) {
        if (e == cudaSuccess && buf.n < count) e = buf.alloc(count);
    };
    if (sc.ws_slots < slots) {
        grow(sc.A[0], slots); grow(sc.B[0], slots); grow(sc.C[0], slots); grow(sc.D[0], slots);
        grow(sc.hits, slots);
        sc.Tm[0].release();
        if (e == cudaSuccess) sc.ws_slots = slots;
    }
    if (two_queues && sc.ws_slots2 < slots) {  // the wavefront loop ping-pongs between two queue sets
        grow(sc.A[1], slots); grow(sc.B[1], slots); grow(sc.C[1], slots); grow(sc.D[1], slots);
        sc.Tm[1].release();
        if (e == cudaSuccess) sc.ws_slots2 = slots;
    }
    grow(sc.acc, acc_elems);
    if (need_time) { grow(sc.Tm[0], sc.ws_slots); if (two_queues) grow(sc.Tm[1], sc.ws_slots2); }
    grow(sc.film, npix * 3);
    grow(sc.sumsq, npix);
    if (e != cudaSuccess) {
        release_workspace_queues(sc);
        return fail(e == cudaErrorMemoryAllocation ? RT_ENOMEM : RT_ECUDA, std::string("render workspace: ") + cudaGetErrorString(e));
    }
    if (!sc.ctr.p) CU(sc.ctr.alloc(Workspace::MAX_PIPES));
    if (!sc.h_ctr) CU(cudaMallocHost(&sc.h_ctr, Workspace::MAX_PIPES * 2 * sizeof(Counters)));
    if (!sc.fork_event) {
        CU(cudaEventCreateWithFlags(&sc.fork_event, cudaEventDisableTiming));
        for (int p = 0; p < Workspace::MAX_PIPES; ++p) {
            CU(cudaStreamCreateWithFlags(&sc.pipe_stream[p], cudaStreamNonBlocking));
            int prio_lo = 0, prio_hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CU(cudaStreamCreateWithPriority(&sc.resume_stream[p], cudaStreamNonBlocking, prio_hi));
            CU(cudaEventCreateWithFlags(&sc.resume_fork[p], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sc.resume_join[p], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sc.join_event[p], cudaEventDisableTiming));
            for (int b = 0; b < 2; ++b) CU(cudaEventCreate(&sc.poll_event[p][b]));
        }
    }
    return RT_OK;
}

// The wavefront loop.  Writes the (unscaled unless apply_scale) film into d_out.
using ExtendFn = void (*)(SceneDev, RenderDev, int);
template <int FEAT, bool RESUME>
static ExtendFn pick_extend_feat(int mode, bool count) {
    return count ? (mode == 3 ? k_extend<3, true, FEAT, RESUME> : (mode == 2 ? k_extend<2, true, FEAT, RESUME> : (mode == 1 ? k_extend<1, true, FEAT, RESUME> : k_extend<0, true, FEAT, RESUME>)))
                 : (mode == 3 ? k_extend<3, false, FEAT, RESUME> : (mode == 2 ? k_extend<2, false, FEAT, RESUME> : (mode == 1 ? k_extend<1, false, FEAT, RESUME> : k_extend<0, false, FEAT, RESUME>)));
}
template <bool RESUME>
static ExtendFn pick_extend(int mode, bool count, int feat) {
    return feat == 2 ? pick_extend_feat<2, RESUME>(mode, count) : (feat == 1 ? pick_extend_feat<1, RESUME>(mode, count) : pick_extend_feat<0, RESUME>(mode, count));
}

// What one rank renders: validated parameters + the lane / track split.
struct RenderPlan {
    RenderDev R{};       // rank-level description (k_finalize and the multi-GPU reduce use it as is)
    uint64_t n_tasks = 0;
    uint32_t n_local = 0;  // lanes this rank owns (0: it contributes nothing)
    size_t npix = 0, slots = 0, acc_elems = 0;  // slots: paths in flight on this rank (all pipelines)
};

int make_plan(const RtScene& sc, const Options& O, const RtRenderParams& P, uint32_t lane_first, uint32_t lane_stride, RenderPlan& pl) {
    if (P.width < 2 || P.height < 2) return fail(RT_EINVAL, "image must be at least 2x2 (reference divides by width-1 / height-1)");
    if (P.samples_per_pixel == 0) return fail(RT_EINVAL, "samples_per_pixel must be >= 1");
    if (P.thread_count == 0) return fail(RT_EINVAL, "thread_count must be >= 1");
    if (P.render_mode > RT_MODE_NORMALS) return fail(RT_EINVAL, "unknown render mode");
    if (lane_stride == 0) return fail(RT_EINVAL, "lane_stride must be >= 1");
    const uint64_t T = P.thread_count;
    pl.n_tasks = std::min<uint64_t>(T, P.samples_per_pixel);
    if (P.samples_per_pixel / T + 1 > 0x1FFFFFFFull) return fail(RT_EINVAL, "too many samples per lane (2^29 at most)");
    pl.npix = (size_t)P.width * P.height;
    if (pl.npix > 0x7FFFFFFFull) return fail(RT_EINVAL, "image too large");
    pl.n_local = lane_first >= pl.n_tasks ? 0u : (uint32_t)((pl.n_tasks - lane_first + lane_stride - 1) / lane_stride);
    // The library, not the caller's thread_count, decides how many paths are in flight: at most max_slots path slots, fed from
    // the rank's chains (one chain = all samples of one (lane, pixel), rendered in order by one slot; a slot whose chain has
    // ended takes the next one).  The path-state workspace is bounded while every lane still owns its plane (merge_planes
    // needs all T of them: 16 B x T x pixels).
    if (pl.n_local) {
        const uint64_t budget = (uint64_t)std::max<int64_t>(O.max_slots, 1);
        const uint64_t chains = (uint64_t)pl.n_local * pl.npix;
        pl.slots = (size_t)std::min<uint64_t>(chains, std::max<uint64_t>(budget, 4 * Workspace::MAX_PIPES));  // >= 1 slot for every pipeline
        if (pl.slots > 0x7FFFFFF0ull) return fail(RT_EINVAL, "more than 2^31 path slots (lower max_slots)");
        pl.acc_elems = (size_t)pl.n_local * pl.npix;
    }
    RenderDev& R = pl.R;
    R.npix = (uint32_t)pl.npix; R.width = (uint32_t)P.width; R.height = (uint32_t)P.height;
    R.max_depth = P.max_depth;
    R.n_tasks = (uint32_t)pl.n_tasks;
    R.lane_first = lane_first; R.lane_stride = lane_stride; R.n_local_lanes = pl.n_local;
    R.spp_whole = (uint32_t)(P.samples_per_pixel / T);
    R.spp_rem = (uint32_t)(P.samples_per_pixel % T);
    R.seed0 = (uint32_t)P.seed; R.seed1 = (uint32_t)(P.seed >> 32);
    R.sx = 1.0f / (float)(P.width - 1);
    R.sy = 1.0f / (float)(P.height - 1);
    R.mode = P.render_mode;
    R.ref_only = (O.traversal == 0 || !sc.boxes_are_bounds) ? 1u : 0u;
    R.debug_skip_redo = O.debug_skip_redo ? 1u : 0u;
    R.tail_budget = (uint32_t)std::abs(O.tail_budget);   // < 0: fixed budget of that many rounds (no adaptation)
    R.tail_fixed = O.tail_budget < 0 ? 1u : 0u;
    R.refill_threshold = (uint32_t)std::min(32, std::max(1, O.refill));
    R.leaf_weight = (uint32_t)std::max(1, O.leaf_weight);
    return RT_OK;
}

// The wavefront loop of one rank.  The caller holds ws.mu.  Leaves the per-lane sums in ws.acc (pl.R.acc); with `finalize` also
// writes the film (unscaled unless apply_scale) into d_out and, if d_sumsq, the per-pixel sum of squared samples.
// On any error every pipeline stream is drained before returning, so the shared workspace is never reused under running kernels.
int render_locked(RtScene& sc, Workspace& ws, const Options& O, RenderPlan& pl, float* d_out, bool apply_scale, float* d_sumsq,
                  bool finalize, cudaStream_t stream, RtStats* stats) {
    const size_t npix = pl.npix;
    CU(cudaSetDevice(sc.device));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (pl.n_local == 0) {  // this rank owns no lane: contributes zeros
        if (finalize) {
            CU(cudaMemsetAsync(d_out, 0, npix * 3 * sizeof(float), stream));
            if (d_sumsq) CU(cudaMemsetAsync(d_sumsq, 0, npix * sizeof(float), stream));
        }
        return RT_OK;
    }
    int rc = ensure_workspace(ws, pl.slots, pl.acc_elems, npix, sc.has_animation, true);
    if (rc) return rc;
    if (!ws.ev0) {
        CU(cudaEventCreate(&ws.ev0));
        CU(cudaEventCreate(&ws.ev1));
    }
    RenderDev& R = pl.R;
    R.acc = ws.acc.p;
    int mode = std::min(3, std::max(0, O.traversal));
    if (!sc.boxes_are_bounds) mode = 0;
    if (mode == 3 && !sc.nodes2.p) mode = 2;  // no rebuilt topology for this scene (wrapped elements, single element, rebuild off)
    const int feat = sc.has_volume ? 2 : (sc.has_wrapped ? 1 : 0);
    const int eblock = 128;
    const size_t esmem = stack_smem_bytes(sc, eblock);

    // Split the rank's lanes over G independent wavefront pipelines, each on its own stream with its own queues and
    // counters.  A persistent extend launch ends with a tail in which a few very long rays keep a few warps busy; with two
    // pipelines in flight the other pipeline's blocks backfill the SMs during that tail (and during shade / launch gaps).
    const char* iter_log_path = std::getenv("RT_ITER_LOG");  // debug: dump (pipe, iteration, live paths, extend ms, shade ms)
    if (iter_log_path && !ws.iter_log.p) CU(ws.iter_log.alloc((size_t)Workspace::MAX_PIPES * Workspace::ITER_LOG_CAP * 9));
    const char* long_rays_path = O.count ? std::getenv("RT_LONG_RAYS") : nullptr;  // debug: dump the rays that took >= 4096 traversal rounds
    if (long_rays_path) {
        if (!ws.long_rays.p) CU(ws.long_rays.alloc(4096 * 8));
        CU(cudaMemsetAsync(ws.long_rays.p, 0, 4096 * 8 * sizeof(float), stream));
    }
    std::vector<std::array<float, 2>> iter_ms[Workspace::MAX_PIPES];
    const int want_pipes = O.pipes > 0 ? O.pipes : (pl.slots >= ((size_t)64 << 20) ? 2 : 4);  // measured on C1..C5 (profiles/README.md)
    const int G = (int)std::min<uint32_t>((uint32_t)std::max(1, std::min(want_pipes, (int)Workspace::MAX_PIPES)), pl.n_local);
    struct Pipe {
        RenderDev R;
        cudaStream_t st;
        int q = 0, buf = 0, q_after[2] = {0, 0};
        bool done = false;
        uint64_t iters = 0;
        const Counters* final_ctr = nullptr;
    };
    std::vector<Pipe> pipes(G);
    // suspend pools (RenderDev::susp): two per pipeline, 1/64 of its slots (1 Ki .. 64 Ki entries) of 4 + stack_cap 8-byte words each
    const uint32_t susp_stride = 4u + sc.dev.stack_cap;
    const uint32_t susp_cap = R.tail_budget ? (uint32_t)std::min<size_t>(65536, std::max<size_t>(1024, pl.slots / (size_t)G / 64)) : 0u;
    if (susp_cap) {
        const size_t need = (size_t)G * 2 * susp_cap * susp_stride;
        if (ws.susp.n < need && ws.susp.alloc(need) != cudaSuccess) { cudaGetLastError(); return fail(RT_ENOMEM, "suspend pool"); }
    }
    {
        size_t off = 0;
        for (int p = 0; p < G; ++p) {
            Pipe& pp = pipes[p];
            // pipeline p renders the rank's local lanes p, p + G, ...: its chains, and its share of the slots
            const uint32_t lanes_p = (pl.n_local - (uint32_t)p + (uint32_t)G - 1u) / (uint32_t)G;
            pp.R = R;
            pp.R.pipe = (uint32_t)p; pp.R.n_pipes = (uint32_t)G;
            pp.R.n_chains = (unsigned long long)lanes_p * npix;
            pp.R.cap = (uint32_t)std::min<uint64_t>(pp.R.n_chains, std::max<uint64_t>(1, (uint64_t)pl.slots * lanes_p / pl.n_local));
            for (int q = 0; q < 2; ++q) {
                pp.R.A[q] = ws.A[q].p + off; pp.R.B[q] = ws.B[q].p + off; pp.R.C[q] = ws.C[q].p + off; pp.R.D[q] = ws.D[q].p + off;
                pp.R.Tm[q] = sc.has_animation ? ws.Tm[q].p + off : nullptr;
            }
            pp.R.hits = ws.hits.p + off;
            pp.R.susp_cap = susp_cap; pp.R.susp_stride = susp_stride;
            for (int b = 0; b < 2; ++b) pp.R.susp[b] = susp_cap ? ws.susp.p + ((size_t)p * 2 + b) * susp_cap * susp_stride : nullptr;
            pp.R.ctr = ws.ctr.p + p;
            pp.st = ws.pipe_stream[p];
            pp.R.iter_log = iter_log_path ? ws.iter_log.p + (size_t)p * Workspace::ITER_LOG_CAP * 9 : nullptr;
            pp.R.long_rays = long_rays_path ? ws.long_rays.p : nullptr;
            off += pp.R.cap;
        }
    }
    // error exits: drain the pipeline streams so the workspace is quiescent when ws.mu is released
    struct Drain {
        std::vector<Pipe>& pipes;
        bool armed = true;
        ~Drain() {
            if (!armed) return;
            for (Pipe& pp : pipes) cudaStreamSynchronize(pp.st);
            cudaDeviceSynchronize();  // RESUME launches run on their own streams
            cudaGetLastError();
        }
    } drain{pipes};

    int egrid = 0, sgrid = 0;
    ExtendFn ext = pick_extend<false>(mode, O.count != 0, feat);
    ExtendFn ext_resume = pick_extend<true>(mode, O.count != 0, feat);  // the same kernel over the parked traversals (RenderDev::susp)
    CU(cudaFuncSetAttribute(ext, cudaFuncAttributePreferredSharedMemoryCarveout, O.carveout >= 0 ? std::min(100, O.carveout) : (int)cudaSharedmemCarveoutDefault));
    rc = persistent_grid(sc, ext, eblock, esmem, &egrid, O.blocks_per_sm);
    if (rc) return rc;
    void (*shade)(SceneDev, RenderDev, int) = mode == 3 ? k_shade<true> : k_shade<false>;
    rc = persistent_grid(sc, shade, 128, 0, &sgrid, O.blocks_per_sm);
    if (rc) return rc;
    int rgrid = 0;  // RESUME launch: one warp per 32 pool entries
    if (susp_cap) {
        int unused = 0;
        rc = persistent_grid(sc, ext_resume, eblock, esmem, &unused, 0);  // sets the shared-memory attribute, checks that it fits
        if (rc) return rc;
        rgrid = (int)((susp_cap + 127u) / 128u);
    }
    for (Pipe& pp : pipes) pp.R.extend_warps = (uint32_t)egrid * (uint32_t)(eblock / 32);

    uint64_t launches = 0;
    double extend_ms = 0.0, shade_ms = 0.0;
    const int chunk = std::max(1, O.check_every);
    // per-launch timing without extra synchronisation: 3 events per iteration, read back after the chunk's poll
    if (O.profile) {
        for (int p = 0; p < G; ++p)
            while (ws.prof_events[p].size() < (size_t)(2 * 3 * chunk)) {
                cudaEvent_t e;
                CU(cudaEventCreate(&e));
                ws.prof_events[p].push_back(e);
            }
    }
    CU(cudaEventRecord(ws.ev0, stream));
    CU(cudaMemsetAsync(ws.acc.p, 0, pl.acc_elems * sizeof(float4), stream));  // every lane plane starts at zero
    CU(cudaEventRecord(ws.fork_event, stream));
    ++launches;
    for (int p = 0; p < G; ++p) {
        Pipe& pp = pipes[p];
        CU(cudaStreamWaitEvent(pp.st, ws.fork_event, 0));
        int ggrid = (int)std::min<size_t>(((size_t)pp.R.cap + 255) / 256, (size_t)sc.sm_count * 8);
        k_generate<<<ggrid, 256, 0, pp.st>>>(sc.dev, pp.R);
        ++launches;
    }
    // a chain takes at most (spp_whole + 1) samples x max_depth segments (x 2: k_shade may send a segment back once, to be traced again
    // on the reference tree); while chains are left every slot advances one every iteration
    uint64_t max_iters = 0;
    for (const Pipe& pp : pipes)
        max_iters = std::max<uint64_t>(max_iters, ((pp.R.n_chains + pp.R.cap - 1) / pp.R.cap + 1) * (uint64_t)(R.spp_whole + 1u) * 2ull * (uint64_t)std::max(R.max_depth, 1));
    max_iters = max_iters * 64 + 65536 + 2 * (uint64_t)chunk;  // parked traversals (tail_budget) stretch a segment over several iterations; a hang guard, not a tight bound
    // The host never knows the live-path count of the iteration it is launching: kernels read it from device memory
    // (grids are persistent, sized by the SM count).  Every `chunk` iterations the counters are copied to pinned memory;
    // the host keeps one chunk in flight per pipeline while it waits for the previous chunk's copy.
    auto enqueue_chunk = [&](int p, int buf) -> int {
        Pipe& pp = pipes[p];
        for (int it = 0; it < chunk; ++it) {
            cudaEvent_t* pe = O.profile ? &ws.prof_events[p][(size_t)(buf * chunk + it) * 3] : nullptr;
            if (pe) CU(cudaEventRecord(pe[0], pp.st));
            pp.R.iter = (uint32_t)std::min<uint64_t>(pp.iters, Workspace::ITER_LOG_CAP - 1);
            if (rgrid) {  // the parked traversals of the previous iteration resume beside the queue launch, on their own (high-priority) stream
                CU(cudaEventRecord(ws.resume_fork[p], pp.st));
                CU(cudaStreamWaitEvent(ws.resume_stream[p], ws.resume_fork[p], 0));
                ext_resume<<<rgrid, eblock, esmem, ws.resume_stream[p]>>>(sc.dev, pp.R, pp.q);
                CU(cudaEventRecord(ws.resume_join[p], ws.resume_stream[p]));
                ++launches;
            }
            ext<<<egrid, eblock, esmem, pp.st>>>(sc.dev, pp.R, pp.q);
            if (pe) CU(cudaEventRecord(pe[1], pp.st));
            if (rgrid) CU(cudaStreamWaitEvent(pp.st, ws.resume_join[p], 0));
            shade<<<sgrid, 128, 0, pp.st>>>(sc.dev, pp.R, pp.q);
            if (pe) CU(cudaEventRecord(pe[2], pp.st));
            launches += 2;
            ++pp.iters;
            pp.q ^= 1;
        }
        CU(cudaMemcpyAsync(&ws.h_ctr[p * 2 + buf], pp.R.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, pp.st));
        CU(cudaEventRecord(ws.poll_event[p][buf], pp.st));
        pp.q_after[buf] = pp.q;  // the queue the next iteration would read
        return RT_OK;
    };
    auto harvest = [&](int p, int buf) -> int {
        if (!O.profile) return RT_OK;
        for (int it = 0; it < chunk; ++it) {
            cudaEvent_t* pe = &ws.prof_events[p][(size_t)(buf * chunk + it) * 3];
            float a = 0.0f, b2 = 0.0f;
            CU(cudaEventElapsedTime(&a, pe[0], pe[1]));
            CU(cudaEventElapsedTime(&b2, pe[1], pe[2]));
            extend_ms += a;   // with G > 1 the pipelines overlap: these are per-kernel device times, not wall time
            shade_ms += b2;
            if (iter_log_path) iter_ms[p].push_back({a, b2});
        }
        return RT_OK;
    };
    for (int p = 0; p < G; ++p) {
        rc = enqueue_chunk(p, 0);
        if (rc) return rc;
    }
    int live = G;
    while (live > 0) {
        for (int p = 0; p < G; ++p) {
            Pipe& pp = pipes[p];
            if (pp.done) continue;
            rc = enqueue_chunk(p, pp.buf ^ 1);
            if (rc) return rc;
            CU(cudaEventSynchronize(ws.poll_event[p][pp.buf]));
            rc = harvest(p, pp.buf);
            if (rc) return rc;
            const Counters& hc = ws.h_ctr[p * 2 + pp.buf];
            const bool drained = hc.nb[pp.q_after[pp.buf]] + hc.nn[pp.q_after[pp.buf]] == 0;
            pp.buf ^= 1;
            if (drained) {  // the chunk already in flight only runs empty iterations
                CU(cudaEventSynchronize(ws.poll_event[p][pp.buf]));
                rc = harvest(p, pp.buf);
                if (rc) return rc;
                pp.final_ctr = &ws.h_ctr[p * 2 + pp.buf];
                pp.done = true;
                --live;
                CU(cudaEventRecord(ws.join_event[p], pp.st));
                CU(cudaStreamWaitEvent(stream, ws.join_event[p], 0));
            }
            if (pp.iters > max_iters) return fail(RT_ECUDA, "wavefront loop did not terminate");
        }
    }
    if (finalize) {
        int fgrid = (int)std::min<size_t>((npix + 255) / 256, (size_t)sc.sm_count * 8);
        k_finalize<<<fgrid, 256, 0, stream>>>(R, d_out, 1.0f / (float)pl.n_tasks, apply_scale ? 1 : 0, d_sumsq);
        ++launches;
    }
    CU(cudaEventRecord(ws.ev1, stream));
    CU(cudaEventSynchronize(ws.ev1));
    CU(cudaGetLastError());
    drain.armed = false;
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, ws.ev0, ws.ev1));
    if (long_rays_path) {
        std::vector<float> lr(4096 * 8);
        CU(cudaMemcpy(lr.data(), ws.long_rays.p, lr.size() * sizeof(float), cudaMemcpyDeviceToHost));
        uint32_t cnt;
        std::memcpy(&cnt, lr.data(), 4);
        if (FILE* f = std::fopen(long_rays_path, "w")) {
            std::fprintf(f, "ox,oy,oz,dx,dy,dz,rounds,on_reference_tree\n");
            for (uint32_t i = 0; i < std::min(cnt, 4095u); ++i) {
                const float* r = &lr[8 * (i + 1)];
                std::fprintf(f, "%.9g,%.9g,%.9g,%.9g,%.9g,%.9g,%.0f,%.0f\n", r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
            }
            std::fclose(f);
        }
    }
    if (iter_log_path) {
        std::vector<uint32_t> lg((size_t)Workspace::MAX_PIPES * Workspace::ITER_LOG_CAP * 9);
        CU(cudaMemcpy(lg.data(), ws.iter_log.p, lg.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (FILE* f = std::fopen(iter_log_path, "w")) {
            std::fprintf(f, "pipe,iter,rays,extend_ms,shade_ms,cum_nodes_div64,cum_prims_div64,cum_rounds_div16,cum_busy_div256,cum_trounds_div16,cum_tbusy_div256,parked,budget\n");
            for (int p = 0; p < G; ++p)
                for (size_t i = 0; i < pipes[p].iters && i < (size_t)Workspace::ITER_LOG_CAP; ++i) {
                    const uint32_t* L = &lg[(size_t)p * Workspace::ITER_LOG_CAP * 9];
                    std::fprintf(f, "%d,%zu,%u,%.4f,%.4f,%u,%u,%u,%u,%u,%u,%u,%u\n", p, i, L[i],
                                 i < iter_ms[p].size() ? iter_ms[p][i][0] : -1.0f, i < iter_ms[p].size() ? iter_ms[p][i][1] : -1.0f,
                                 L[Workspace::ITER_LOG_CAP + i], L[2 * Workspace::ITER_LOG_CAP + i], L[4 * Workspace::ITER_LOG_CAP + i],
                                 L[5 * Workspace::ITER_LOG_CAP + i], L[6 * Workspace::ITER_LOG_CAP + i], L[7 * Workspace::ITER_LOG_CAP + i],
                                 L[3 * Workspace::ITER_LOG_CAP + i], L[8 * Workspace::ITER_LOG_CAP + i]);
                }
            std::fclose(f);
        }
    }
    Counters total{};
    uint64_t iters = 0;
    for (const Pipe& pp : pipes) {
        total.rays += pp.final_ctr->rays; total.paths += pp.final_ctr->paths;
        total.node_visits += pp.final_ctr->node_visits; total.prim_tests += pp.final_ctr->prim_tests; total.exact_steps += pp.final_ctr->exact_steps;
        total.redo_total += pp.final_ctr->redo_total;
        total.ref_rounds += pp.final_ctr->ref_rounds; total.ref_lanes += pp.final_ctr->ref_lanes;
        total.tail_rounds += pp.final_ctr->tail_rounds; total.tail_busy_lanes += pp.final_ctr->tail_busy_lanes;
        total.rounds += pp.final_ctr->rounds; total.busy_lanes += pp.final_ctr->busy_lanes;
        total.max_sp = std::max(total.max_sp, pp.final_ctr->max_sp);
        for (int h = 0; h < 24; ++h) total.hist[h] += pp.final_ctr->hist[h];
        iters = std::max(iters, pp.iters);
    }
    if (O.count && std::getenv("RT_TAIL_LOG"))
        std::fprintf(stderr, "[rt_render] warp rounds %llu (lanes holding a ray %.2f), after the queue ran dry %llu (%.2f lanes), with >= 1 lane on the reference tree %llu (%.2f such lanes), redo rays %llu of %llu\n",
                     total.rounds, total.rounds ? (double)total.busy_lanes / total.rounds : 0.0, total.tail_rounds, total.tail_rounds ? (double)total.tail_busy_lanes / total.tail_rounds : 0.0,
                     total.ref_rounds, total.ref_rounds ? (double)total.ref_lanes / total.ref_rounds : 0.0, total.redo_total, total.rays);
    if (stats) {
        stats->ray_segments = total.rays;
        stats->paths = total.paths;
        stats->iterations = iters;
        stats->kernel_launches = launches;
        stats->render_ms = ms;
        stats->extend_ms = extend_ms;
        stats->shade_ms = shade_ms;
        stats->node_visits = total.node_visits;
        stats->prim_tests = total.prim_tests;
        stats->exact_steps = total.exact_steps;
        stats->redo_rays = total.redo_total;
        stats->max_stack = total.max_sp;
        stats->stack_capacity = sc.dev.stack_cap;
        stats->slots = (uint32_t)pl.slots;
        stats->pipes = (uint32_t)G;
        stats->warp_rounds = total.rounds; stats->busy_lane_rounds = total.busy_lanes;
        for (int h = 0; h < 24; ++h) stats->rounds_hist[h] = total.hist[h];
    }
    return RT_OK;
}

int render_core(RtScene& sc, const RtRenderParams& P, uint32_t lane_first, uint32_t lane_stride, float* d_out, bool apply_scale,
                float* d_sumsq, cudaStream_t stream, RtStats* stats) {
    const Options O = options_snapshot();
    RenderPlan pl;
    int rc = make_plan(sc, O, P, lane_first, lane_stride, pl);
    if (rc) return rc;
    Workspace& ws = workspace_of(sc.device);
    std::lock_guard<std::mutex> ws_lock(ws.mu);
    return render_locked(sc, ws, O, pl, d_out, apply_scale, d_sumsq, true, stream, stats);
}

}  // namespace

// The C ABI never lets a C++ exception (std::bad_alloc from the host-side packing, std::system_error from the builder threads)
// cross the boundary: the reference panics, this library returns an error code.
template <class F>
static int guarded(F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return fail(RT_ENOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(RT_ECUDA, std::string("internal error: ") + e.what());
    } catch (...) {
        return fail(RT_ECUDA, "internal error");
    }
}

// Device -> pageable host copy through the workspace's pinned staging chunks (the caller holds ws.mu and has synchronised the
// producer of `d_src`).  A plain cudaMemcpy into pageable memory ran at 1.5 GB/s here (16 ms for a 1920x1080 film, 64 ms at 4K).
extern "C" {

const char* rt_last_error(void) { return g_err.c_str(); }

int rt_device_count(int* out_count) {
    if (!out_count) return fail(RT_EINVAL, "out_count is NULL");
    CU(cudaGetDeviceCount(out_count));
    return RT_OK;
}

int rt_presplit_refs(const float* boxes6, const float* tri9, const uint8_t* is_tri, uint32_t n, uint64_t budget, int device,
                     uint32_t* out_items, float* out_boxes6, uint64_t* out_n) {
    if (!boxes6 || !tri9 || !is_tri || !out_items || !out_boxes6 || !out_n) return fail(RT_EINVAL, "NULL argument");
    return guarded([&]() -> int {
        static_assert(sizeof(rt_rebuild::Box3) == 24 && sizeof(rt_rebuild::Ref) == 28, "reference layout");
        const rt_rebuild::Box3* leaf = reinterpret_cast<const rt_rebuild::Box3*>(boxes6);
        rt_rebuild::Buffer<rt_rebuild::Ref> refs;
        if (device < 0) {
            rt_rebuild::presplit(leaf, n, tri9, is_tri, (size_t)budget, refs);
        } else {
            CU(cudaSetDevice(device));
            const size_t cap = std::max<size_t>((size_t)budget, n);
            DevBuf<rt_rebuild::Box3> d_lb;
            DevBuf<float> d_tri;
            DevBuf<uint8_t> d_is;
            DevBuf<rt_rebuild::Ref> d_refs;
            CU(d_lb.alloc(n)); CU(d_tri.alloc(9 * (size_t)n)); CU(d_is.alloc(n)); CU(d_refs.alloc(cap));
            CU(cudaMemcpy(d_lb.p, leaf, (size_t)n * 24, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d_tri.p, tri9, (size_t)n * 36, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d_is.p, is_tri, n, cudaMemcpyHostToDevice));
            cudaStream_t st = nullptr;
            CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            size_t n_refs = 0;
            const cudaError_t e = rt_presplit::run(d_lb.p, d_tri.p, d_is.p, n, (size_t)budget, d_refs.p, &n_refs, st);
            cudaStreamDestroy(st);
            if (e != cudaSuccess) return fail(RT_ECUDA, std::string("device pre-split: ") + cudaGetErrorString(e));
            refs.resize_discard(n_refs);
            CU(cudaMemcpy(refs.data(), d_refs.p, n_refs * sizeof(rt_rebuild::Ref), cudaMemcpyDeviceToHost));
        }
        for (size_t i = 0; i < refs.size(); ++i) {
            out_items[i] = refs[i].item;
            std::memcpy(out_boxes6 + 6 * i, &refs[i].box, 24);
        }
        *out_n = refs.size();
        return RT_OK;
    });
}

int rt_set_option(const char* key, int64_t value) {
    if (!key) return fail(RT_EINVAL, "key is NULL");
    std::string k(key);
    std::lock_guard<std::mutex> lk(g_opt_mu);
    if (k == "traversal") g_opt_traversal = (int)value;
    else if (k == "rebuild") g_opt_rebuild = (int)value;
    else if (k == "rebuild_wrapped") g_opt_rebuild_wrapped = (int)value;
    else if (k == "rebuild_min_items") g_opt_rebuild_min_items = (int)value;
    else if (k == "presplit") g_opt_presplit = (int)value;
    else if (k == "presplit_device") g_opt_presplit_device = (int)value;
    else if (k == "build_device") g_opt_build_device = (int)value;
    else if (k == "device_cache_mb") { g_opt_device_cache_mb = std::max<int64_t>(0, value); }
    else if (k == "presplit_large") g_opt_presplit_large = (int)std::max<int64_t>(100, std::min<int64_t>(value, 1600));
    else if (k == "rebuild_depth_extra") g_opt_rebuild_depth_extra = (int)value;
    else if (k == "profile") g_opt_profile = (int)value;
    else if (k == "count") g_opt_count = (int)value;
    else if (k == "debug_skip_redo") g_opt_debug_skip_redo = (int)value;
    else if (k == "carveout") g_opt_carveout = (int)value;
    else if (k == "tail_budget") g_opt_tail_budget = (int)std::max<int64_t>(-(1 << 24), std::min<int64_t>(value, 1 << 24));
    else if (k == "refill") g_opt_refill = (int)value;
    else if (k == "leaf_weight") g_opt_leaf_weight = (int)value;
    else if (k == "pipes") g_opt_pipes = (int)value;
    else if (k == "check_every") g_opt_check_every = std::max<int>(1, (int)value);
    else if (k == "blocks_per_sm") g_opt_blocks_per_sm = (int)value;
    else if (k == "max_slots") g_opt_max_slots = std::max<int64_t>(1, value);
    else return fail(RT_EINVAL, "unknown option " + k);
    return RT_OK;
}

int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out_scene) {
    if (!desc || !out_scene) return fail(RT_EINVAL, "NULL argument");
    *out_scene = nullptr;
    // RT_PACK_DRYRUN (profiling of the host phases on a box without a GPU, together with RT_PACK_LOG): the device check is skipped
    // and the call fails at the upload instead
    if (!std::getenv("RT_PACK_DRYRUN")) {
        int count = 0;
        CU(cudaGetDeviceCount(&count));
        if (device < 0 || device >= count) return fail(RT_ECUDA, "CUDA device ordinal out of range (no CPU fallback exists)");
    }
    RtScene* sc = new (std::nothrow) RtScene();
    if (!sc) return fail(RT_ENOMEM, "out of host memory");
    sc->device = device;
    const Options O = options_snapshot();
    const auto t_create = std::chrono::steady_clock::now();
    int rc = guarded([&] { return pack_scene(*desc, *sc, O); });
    if (std::getenv("RT_PACK_LOG"))
        std::fprintf(stderr, "[rt_scene_create] total (with the release of the host buffers) %8.1f ms\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create).count());
    if (rc != RT_OK) {
        delete sc;
        return rc;
    }
    *out_scene = sc;
    return RT_OK;
}

void rt_scene_destroy(RtScene* scene) {
    if (!scene) return;
    cudaSetDevice(scene->device);
    const bool log = std::getenv("RT_PACK_LOG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    // The scene's blocks go back to the block cache, not to cudaFree (which would wait for the device by itself): work a caller has
    // queued against this scene (rt_trace_device / rt_render*_device on its own stream) must be over before another scene gets them.
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    delete scene;
    if (log)
        std::fprintf(stderr, "[rt_scene_destroy] device busy for %8.1f ms, frees %8.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
}

int rt_scene_set_ray_time(RtScene* scene, float time) {
    if (!scene) return fail(RT_EINVAL, "scene is NULL");
    scene->dev.trace_time = time;
    return RT_OK;
}

int rt_scene_device_bytes(const RtScene* scene, uint64_t* out_bytes) {
    if (!scene || !out_bytes) return fail(RT_EINVAL, "NULL argument");
    *out_bytes = scene->device_bytes;
    return RT_OK;
}

int rt_scene_accelerator_nodes(const RtScene* scene, void* out_nodes64, uint64_t cap_nodes, uint64_t* out_n, uint32_t* out_root, uint32_t* out_depth) {
    if (!scene || !out_n || !out_root || !out_depth) return fail(RT_EINVAL, "NULL argument");
    *out_n = scene->nodes2.n;
    *out_root = scene->dev.root2;
    *out_depth = scene->depth2;
    if (!out_nodes64 || cap_nodes < scene->nodes2.n || scene->nodes2.n == 0) return RT_OK;
    CU(cudaSetDevice(scene->device));
    CU(cudaMemcpy(out_nodes64, scene->nodes2.p, scene->nodes2.n * sizeof(NodeDev), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_trace_device(RtScene* scene, const float* d_org, const float* d_dir, uint64_t n, float t_min, float t_max,
                    uint32_t* d_hit_id, float* d_t, float* d_normal, uint8_t* d_front_face, float* d_uv, void* cuda_stream) {
    if (!scene) return fail(RT_EINVAL, "scene is NULL");
    if (n == 0) return RT_OK;
    if (!d_org || !d_dir) return fail(RT_EINVAL, "ray buffers are NULL");
    CU(cudaSetDevice(scene->device));
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const int block = 128;
    const size_t smem = stack_smem_bytes(*scene, block);
    int grid = 0, rc;
    const int mode = scene->boxes_are_bounds ? std::min(3, std::max(0, options_snapshot().traversal)) : 0;
    auto trc = mode == 3 ? k_trace<3> : (mode == 2 ? k_trace<2> : (mode == 1 ? k_trace<1> : k_trace<0>));
    rc = persistent_grid(*scene, trc, block, smem, &grid);
    if (rc) return rc;
    grid = (int)std::min<uint64_t>((n + block - 1) / block, (uint64_t)grid * 4);
    trc<<<grid, block, smem, stream>>>(scene->dev, d_org, d_dir, n, t_min, t_max, d_hit_id, d_t, d_normal, d_front_face, d_uv);
    CU(cudaGetLastError());
    return RT_OK;
}

int rt_trace(RtScene* scene, const float* org, const float* dir, uint64_t n, float t_min, float t_max, uint32_t* hit_id,
             float* t, float* normal, uint8_t* front_face, float* uv) {
    if (!scene) return fail(RT_EINVAL, "scene is NULL");
    if (n == 0) return RT_OK;
    if (!org || !dir) return fail(RT_EINVAL, "ray buffers are NULL");
    CU(cudaSetDevice(scene->device));
    DevBuf<float> d_org, d_dir, d_t, d_n, d_uv;
    DevBuf<uint32_t> d_id;
    DevBuf<uint8_t> d_ff;
    CU(d_org.alloc(n * 3));
    CU(d_dir.alloc(n * 3));
    CU(d_id.alloc(n));
    CU(d_t.alloc(n));
    CU(d_n.alloc(n * 3));
    CU(d_ff.alloc(n));
    CU(d_uv.alloc(n * 2));
    CU(cudaMemcpy(d_org.p, org, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_dir.p, dir, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    int rc = rt_trace_device(scene, d_org.p, d_dir.p, n, t_min, t_max, d_id.p, d_t.p, d_n.p, d_ff.p, d_uv.p, nullptr);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    if (hit_id) CU(cudaMemcpy(hit_id, d_id.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (t) CU(cudaMemcpy(t, d_t.p, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (normal) CU(cudaMemcpy(normal, d_n.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (front_face) CU(cudaMemcpy(front_face, d_ff.p, n, cudaMemcpyDeviceToHost));
    if (uv) CU(cudaMemcpy(uv, d_uv.p, n * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_render_planes_device(RtScene* scene, const RtRenderParams* params, uint32_t lane_first, uint32_t lane_stride,
                            float* d_plane_sum, void* cuda_stream, RtStats* out_stats) {
    if (!scene || !params || !d_plane_sum) return fail(RT_EINVAL, "NULL argument");
    return guarded([&] { return render_core(*scene, *params, lane_first, lane_stride, d_plane_sum, false, nullptr, (cudaStream_t)cuda_stream, out_stats); });
}

int rt_render_device(RtScene* scene, const RtRenderParams* params, float* d_out_rgb, void* cuda_stream, RtStats* out_stats) {
    if (!scene || !params || !d_out_rgb) return fail(RT_EINVAL, "NULL argument");
    return guarded([&] { return render_core(*scene, *params, 0, 1, d_out_rgb, true, nullptr, (cudaStream_t)cuda_stream, out_stats); });
}

int rt_render_moments_device(RtScene* scene, const RtRenderParams* params, float* d_out_rgb, float* d_out_sumsq, void* cuda_stream,
                             RtStats* out_stats) {
    if (!scene || !params || !d_out_rgb || !d_out_sumsq) return fail(RT_EINVAL, "NULL argument");
    return guarded([&] { return render_core(*scene, *params, 0, 1, d_out_rgb, true, d_out_sumsq, (cudaStream_t)cuda_stream, out_stats); });
}

int rt_render(RtScene* scene, const RtRenderParams* params, float* out_rgb, RtStats* out_stats) {
    if (!scene || !params || !out_rgb) return fail(RT_EINVAL, "NULL argument");
    if (params->width < 2 || params->height < 2) return fail(RT_EINVAL, "image must be at least 2x2");
    CU(cudaSetDevice(scene->device));
    return guarded([&]() -> int {
        const size_t n = (size_t)params->width * params->height * 3;
        const Options O = options_snapshot();
        RenderPlan pl;
        int rc = make_plan(*scene, O, *params, 0, 1, pl);
        if (rc) return rc;
        Workspace& ws = workspace_of(scene->device);
        std::lock_guard<std::mutex> ws_lock(ws.mu);
        if (ws.film.n < n) {  // the film lives in the per-device workspace: no cudaMalloc / cudaFree per call
            if (ws.film.alloc(n) != cudaSuccess) return fail(RT_ENOMEM, "film buffer");
        }
        rc = render_locked(*scene, ws, O, pl, ws.film.p, true, nullptr, true, nullptr, out_stats);
        if (rc) return rc;
        return download_staged(ws, out_rgb, ws.film.p, n * sizeof(float));  // render_locked has synchronised the render
    });
}

// render() on several GPUs of one box from ONE call.  The reference splits the SAMPLES over its worker threads and
// averages their planes (rendering.rs:161-210, 222-252); here the T lanes are dealt round-robin to the GPUs (lane id % n == g), one
// host thread drives each GPU, and the film combine is k_merge_peers: a fixed-order reduce over NVLink peer memory fused with
// finalize, so the returned film has the same bits for every n (including n = 1 and rt_render).
static int enable_peers(const std::vector<int>& devs) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, bool> done;
    std::lock_guard<std::mutex> lk(mu);
    for (int a : devs)
        for (int b : devs) {
            if (a == b || done.count({a, b})) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) return fail(RT_EUNSUPPORTED, "rt_render_multi needs peer access between the GPUs (NVLink / NVSwitch)");
            CU(cudaSetDevice(a));
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) return fail(RT_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            done[{a, b}] = true;
        }
    return RT_OK;
}

static int render_multi_impl(RtScene* const* scenes, int n, const RtRenderParams& P, float* out_rgb, float* const* d_out_rgb, RtStats* out_stats) {
    if (n < 1 || n > RT_MAX_GPUS) return fail(RT_EINVAL, "n_scenes must be 1..16");
    std::vector<int> devs(n);
    for (int g = 0; g < n; ++g) {
        if (!scenes[g]) return fail(RT_EINVAL, "scene is NULL");
        devs[g] = scenes[g]->device;
        for (int h = 0; h < g; ++h)
            if (devs[h] == devs[g]) return fail(RT_EINVAL, "rt_render_multi: every scene must live on a different GPU");
    }
    const Options O = options_snapshot();
    std::vector<RenderPlan> plans(n);
    for (int g = 0; g < n; ++g) {
        int rc = make_plan(*scenes[g], O, P, (uint32_t)g, (uint32_t)n, plans[g]);
        if (rc) return rc;
    }
    int rc = enable_peers(devs);
    if (rc) return rc;
    // lock every involved workspace (ascending device order: no deadlock with a concurrent rt_render_multi) until the merge is done
    std::vector<int> order(n);
    for (int g = 0; g < n; ++g) order[g] = g;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return devs[a] < devs[b]; });
    std::vector<Workspace*> wss(n);
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int g : order) {
        wss[g] = &workspace_of(devs[g]);
        locks.emplace_back(wss[g]->mu);
    }
    const size_t npix = plans[0].npix;
    std::vector<int> rcs(n, RT_OK);
    std::vector<std::string> errs(n);
    std::vector<RtStats> stats(n);
    auto run_all = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int g = 1; g < n; ++g) th.emplace_back([&, g] { rcs[g] = guarded([&] { return fn(g); }); errs[g] = g_err; });
        rcs[0] = guarded([&] { return fn(0); });
        errs[0] = g_err;
        for (auto& t : th) t.join();
        for (int g = 0; g < n; ++g)
            if (rcs[g]) return fail(rcs[g], "GPU " + std::to_string(devs[g]) + ": " + errs[g]);
        return (int)RT_OK;
    };
    // phase 1: every GPU renders its lanes into its own workspace (per-lane sums stay in HBM)
    rc = run_all([&](int g) -> int {
        int r = render_locked(*scenes[g], *wss[g], O, plans[g], nullptr, false, nullptr, false, nullptr, &stats[g]);
        if (r) return r;
        if (wss[g]->film.n < npix * 3 && wss[g]->film.alloc(npix * 3) != cudaSuccess) return fail(RT_ENOMEM, "film buffer");
        return RT_OK;
    });
    if (rc) return rc;
    // phase 2: GPU g merges pixel slice g from all ranks' lane sums (peer loads) and stores it (host copy, or peer stores into every film)
    PeerPlanes PP{};
    PP.n_ranks = (uint32_t)n;
    for (int g = 0; g < n; ++g) PP.acc[g] = plans[g].n_local ? wss[g]->acc.p : nullptr;
    std::vector<float> merge_ms(n, 0.0f);
    rc = run_all([&](int g) -> int {
        CU(cudaSetDevice(devs[g]));
        Workspace& ws = *wss[g];
        const uint32_t p0 = (uint32_t)(npix * (size_t)g / n), p1 = (uint32_t)(npix * (size_t)(g + 1) / n);
        PeerPlanes pp = PP;
        if (d_out_rgb) { pp.n_out = (uint32_t)n; for (int h = 0; h < n; ++h) pp.out[h] = d_out_rgb[h]; }
        else { pp.n_out = 1; pp.out[0] = ws.film.p; }
        cudaStream_t st = ws.pipe_stream[0];
        CU(cudaEventRecord(ws.ev0, st));
        if (p1 > p0) {
            const int grid = (int)std::min<size_t>(((size_t)(p1 - p0) + 255) / 256, (size_t)scenes[g]->sm_count * 8);
            k_merge_peers<<<grid, 256, 0, st>>>(pp, plans[g].R, p0, p1, 1.0f / (float)plans[g].n_tasks);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(ws.ev1, st));
        CU(cudaEventSynchronize(ws.ev1));
        if (out_rgb && p1 > p0) {  // every GPU downloads its own slice through its own pinned staging chunks, in parallel
            int r = download_staged(ws, out_rgb + 3 * (size_t)p0, ws.film.p + 3 * (size_t)p0, (size_t)(p1 - p0) * 3 * sizeof(float));
            if (r) return r;
        }
        CU(cudaEventElapsedTime(&merge_ms[g], ws.ev0, ws.ev1));
        return RT_OK;
    });
    if (rc) return rc;
    if (out_stats)
        for (int g = 0; g < n; ++g) {
            out_stats[g] = stats[g];
            out_stats[g].merge_ms = merge_ms[g];
            out_stats[g].kernel_launches += 1;
        }
    return RT_OK;
}

int rt_render_multi(RtScene* const* scenes, int n_scenes, const RtRenderParams* params, float* out_rgb, RtStats* out_stats) {
    if (!scenes || !params || !out_rgb) return fail(RT_EINVAL, "NULL argument");
    return guarded([&] { return render_multi_impl(scenes, n_scenes, *params, out_rgb, nullptr, out_stats); });
}

int rt_render_multi_device(RtScene* const* scenes, int n_scenes, const RtRenderParams* params, float* const* d_out_rgb, RtStats* out_stats) {
    if (!scenes || !params || !d_out_rgb) return fail(RT_EINVAL, "NULL argument");
    for (int g = 0; g < n_scenes; ++g)
        if (!d_out_rgb[g]) return fail(RT_EINVAL, "NULL film pointer");
    return guarded([&] { return render_multi_impl(scenes, n_scenes, *params, nullptr, d_out_rgb, out_stats); });
}

// Releases every per-device render workspace (queues, accumulators, streams, events).  Optional: a process may simply exit; a
// long-lived caller (or a benchmark that wants a clean device before it returns) calls this when no render is in flight.
int rt_shutdown(void) {
    rt_devcache::trim(-1);
    std::lock_guard<std::mutex> lk(g_ws_mu);
    for (auto& kv : g_ws) {
        Workspace* w = kv.second;
        if (!w) continue;
        std::lock_guard<std::mutex> wl(w->mu);
        if (cudaSetDevice(kv.first) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaDeviceSynchronize();
        release_workspace_queues(*w);
        w->ctr.release(); w->film.release(); w->sumsq.release(); w->iter_log.release(); w->long_rays.release(); w->susp.release();
        if (w->h_ctr) { cudaFreeHost(w->h_ctr); w->h_ctr = nullptr; }
        for (int b = 0; b < 2; ++b) {
            if (w->h_stage[b]) { cudaFreeHost(w->h_stage[b]); w->h_stage[b] = nullptr; }
            if (w->stage_event[b]) { cudaEventDestroy(w->stage_event[b]); w->stage_event[b] = nullptr; }
        }
        if (w->copy_stream) { cudaStreamDestroy(w->copy_stream); w->copy_stream = nullptr; }
        for (int p = 0; p < Workspace::MAX_PIPES; ++p) {
            if (w->pipe_stream[p]) { cudaStreamDestroy(w->pipe_stream[p]); w->pipe_stream[p] = nullptr; }
            if (w->resume_stream[p]) { cudaStreamDestroy(w->resume_stream[p]); w->resume_stream[p] = nullptr; }
            if (w->resume_fork[p]) { cudaEventDestroy(w->resume_fork[p]); w->resume_fork[p] = nullptr; }
            if (w->resume_join[p]) { cudaEventDestroy(w->resume_join[p]); w->resume_join[p] = nullptr; }
            if (w->join_event[p]) { cudaEventDestroy(w->join_event[p]); w->join_event[p] = nullptr; }
            for (int b = 0; b < 2; ++b)
                if (w->poll_event[p][b]) { cudaEventDestroy(w->poll_event[p][b]); w->poll_event[p][b] = nullptr; }
            for (cudaEvent_t e : w->prof_events[p]) cudaEventDestroy(e);
            w->prof_events[p].clear();
        }
        if (w->fork_event) { cudaEventDestroy(w->fork_event); w->fork_event = nullptr; }
        if (w->ev0) { cudaEventDestroy(w->ev0); w->ev0 = nullptr; }
        if (w->ev1) { cudaEventDestroy(w->ev1); w->ev1 = nullptr; }
        cudaGetLastError();
    }
    return RT_OK;
}

int rt_film_scale_device(float* d_rgb, uint64_t n_floats, float multiplier, void* cuda_stream) {
    if (!d_rgb) return fail(RT_EINVAL, "NULL argument");
    if (n_floats == 0) return RT_OK;
    int grid = (int)std::min<uint64_t>((n_floats + 255) / 256, 148 * 16);
    k_scale<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_rgb, n_floats, multiplier);
    CU(cudaGetLastError());
    return RT_OK;
}

int rt_to_rgb8_gamma2_device(const float* d_rgb, uint64_t n_pixels, uint8_t* d_out, void* cuda_stream) {
    if (!d_rgb || !d_out) return fail(RT_EINVAL, "NULL argument");
    if (n_pixels == 0) return RT_OK;
    uint64_t n = n_pixels * 3;
    int grid = (int)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    k_gamma2<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(d_rgb, n, d_out);
    CU(cudaGetLastError());
    return RT_OK;
}

}  // extern "C"
