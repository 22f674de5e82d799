// rt_sahbuild.cuh — SURVEY §8 row f4: the accelerator tree (rt_rebuild.hpp, build) built on the device, straight into the 64-byte
// node records k_extend reads.  Same heuristic as the host builder: top-down surface-area heuristic over the references' boxes,
// 32 centroid bins per axis for ranges above 32 references, an exact search over every centroid plane below; one reference per
// leaf; pre-order numbering by position (the node over leaf positions [lo, hi) has its left subtree at me + 1 and its right
// subtree at me + (mid - lo)); the depth cap binds the same way (a split is only allowed if both sides still fit).
//
// * Ranges above 32 references, level by level (k_bounds, k_bin, k_eval, k_side + one prefix sum, k_scatter): every reference adds
//   itself to its node's centroid bounds and bins with atomics — through shared memory when its whole block sits in one node, which
//   is every block of the first levels —, one warp per node scans the bins (lane = bin), and a stable partition by a prefix sum over
//   "goes left" moves the reference indices.  The boxes of a node's two children are the bin unions of the chosen split: the
//   children-in-parent record is complete the moment the split is chosen, no bottom-up pass.
// * Ranges of <= 32 references: one warp builds the whole subtree (k_small), one reference per lane; every centroid of every axis
//   is tried as the plane (what the host's sorted sweep enumerates, up to equal centroids), side boxes by warp reductions.
// The tree is the same from run to run (nothing depends on the order in which threads finish).
#pragma once
#include <cub/device/device_scan.cuh>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "rt_device.cuh"
#include "rt_rebuild.hpp"

namespace rt_sahbuild {

using rt::NodeDev;
using rt_rebuild::Box3;
using rt_rebuild::Ref;

static const int BINS = 32;
static const uint32_t SMALL = 32;       // ranges up to this many references go to k_small
static const uint32_t NONE = 0xFFFFFFFFu;

struct Active {  // a range above SMALL references being split at the current level
    uint32_t lo, hi, me, depth;
};
struct Split {   // what k_eval decided for an active node (read by k_side / k_scatter / k_fallback_boxes / k_emit)
    float cmin, scale;      // bin of a centroid c on `axis`: clamp((c - cmin) * scale)
    uint32_t axis, bin;     // goes left: bin index <= bin;  axis == 3: median fallback (left = the first n/2 positions)
    uint32_t mid;
    uint32_t slot_l, slot_r;  // next level's active slots of the children, or NONE
    uint32_t pad;
    float lbox[6], rbox[6];
    uint32_t band_a, band_c;  // float bits (>= 0): maxima over the node's references
};
struct NodeBins {
    int cmin[3], cmax[3];          // centroid bounds, floats mapped to ordered ints
    uint32_t band_a, band_c;
    int bmin[3][BINS][3], bmax[3][BINS][3];  // per axis and bin: box union (ordered ints)
    uint32_t cnt[3][BINS];
};
struct Counts {
    uint32_t n_next;    // active nodes of the next level
    uint32_t n_small;   // small tasks
    uint32_t depth;     // deepest internal level made
    uint32_t n_done;    // finished big nodes (for k_fix_leaves)
};

__device__ __forceinline__ int f2o(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }  // order-preserving
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }
__device__ __forceinline__ float centroid(const Box3& b, int k) { return 0.5f * b.mn[k] + 0.5f * b.mx[k]; }
__device__ __forceinline__ uint32_t bf16_up(float v) {  // smallest bf16 >= v (v >= 0)
    uint32_t b = __float_as_uint(v);
    if (b & 0xFFFFu) b = (b | 0xFFFFu) + 1u;
    return b >> 16;
}
__device__ __forceinline__ double half_area6(const float* b) {  // (mn xyz, mx xyz); empty boxes give 0
    double x = (double)b[3] - b[0], y = (double)b[4] - b[1], z = (double)b[5] - b[2];
    x = 0.0 < x ? x : 0.0; y = 0.0 < y ? y : 0.0; z = 0.0 < z ? z : 0.0;
    return x * y + y * z + z * x;
}
__device__ __forceinline__ int bin_of(float c, float cmin, float scale) {
    int b = (int)((c - cmin) * scale);
    return min(BINS - 1, max(0, b));
}

__global__ void k_iota(uint32_t* idx, uint32_t* node_of, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = i; node_of[i] = 0u; }
}

__global__ void k_clear_bins(NodeBins* bins, uint32_t n_active) {
    const uint32_t total = n_active * (uint32_t)(sizeof(NodeBins) / 4);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t w = i % (uint32_t)(sizeof(NodeBins) / 4);
        int v;
        // layout in words: cmin[3] cmax[3] band_a band_c | bmin[288] | bmax[288] | cnt[96]
        if (w < 3) v = 0x7FFFFFFF; else if (w < 6) v = (int)0x80000000; else if (w < 8) v = 0;
        else if (w < 8 + 288) v = 0x7FFFFFFF; else if (w < 8 + 576) v = (int)0x80000000; else v = 0;
        reinterpret_cast<int*>(bins)[i] = v;
    }
}

// centroid bounds and band maxima of every active node
__global__ void k_bounds(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ node_of, uint32_t n, const Ref* __restrict__ refs,
                         const float* __restrict__ item_a, const float* __restrict__ item_c, NodeBins* bins) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t slot = p < n ? node_of[p] : NONE;
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    uint32_t ba = 0u, bc = 0u;
    if (slot != NONE) {
        const Ref r = refs[idx[p]];
        for (int k = 0; k < 3; ++k) { const int o = f2o(centroid(r.box, k)); lo[k] = o; hi[k] = o; }
        ba = __float_as_uint(item_a[r.item]); bc = __float_as_uint(item_c[r.item]);
    }
    // one node for the whole warp (the usual case above 32 references): reduce first
    const uint32_t s0 = __shfl_sync(0xFFFFFFFFu, slot, 0);
    if (__all_sync(0xFFFFFFFFu, slot == s0)) {
        if (s0 == NONE) return;
        for (int off = 16; off > 0; off >>= 1) {
            for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off)); hi[k] = max(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off)); }
            ba = max(ba, __shfl_xor_sync(0xFFFFFFFFu, ba, off)); bc = max(bc, __shfl_xor_sync(0xFFFFFFFFu, bc, off));
        }
        if ((threadIdx.x & 31u) == 0u) {
            NodeBins& B = bins[s0];
            for (int k = 0; k < 3; ++k) { atomicMin(&B.cmin[k], lo[k]); atomicMax(&B.cmax[k], hi[k]); }
            atomicMax(&B.band_a, ba); atomicMax(&B.band_c, bc);
        }
    } else if (slot != NONE) {
        NodeBins& B = bins[slot];
        for (int k = 0; k < 3; ++k) { atomicMin(&B.cmin[k], lo[k]); atomicMax(&B.cmax[k], hi[k]); }
        atomicMax(&B.band_a, ba); atomicMax(&B.band_c, bc);
    }
}

// every reference into its node's bins, all three axes.  A block whose first and last position belong to the same node (positions
// of a node are contiguous) bins in shared memory and flushes the bins it touched.
__global__ void __launch_bounds__(256) k_bin(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ node_of, uint32_t n, const Ref* __restrict__ refs,
                                             NodeBins* bins) {
    __shared__ int s_min[3][BINS][3], s_max[3][BINS][3];
    __shared__ uint32_t s_cnt[3][BINS];
    const uint32_t p0 = blockIdx.x * blockDim.x, p = p0 + threadIdx.x;
    const uint32_t first = node_of[min(p0, n - 1)], last = node_of[min(p0 + blockDim.x - 1, n - 1)];
    const bool uniform = first == last;
    if (uniform && first == NONE) return;
    if (uniform) {
        for (uint32_t i = threadIdx.x; i < 3 * BINS * 3; i += blockDim.x) { (&s_min[0][0][0])[i] = 0x7FFFFFFF; (&s_max[0][0][0])[i] = (int)0x80000000; }
        for (uint32_t i = threadIdx.x; i < 3 * BINS; i += blockDim.x) (&s_cnt[0][0])[i] = 0u;
        __syncthreads();
    }
    const uint32_t slot = p < n ? node_of[p] : NONE;
    if (slot != NONE) {
        const NodeBins& B = bins[slot];
        const Ref r = refs[idx[p]];
        int bo[6];
        for (int k = 0; k < 3; ++k) { bo[k] = f2o(r.box.mn[k]); bo[3 + k] = f2o(r.box.mx[k]); }
        for (int k = 0; k < 3; ++k) {
            const float cmin = o2f(B.cmin[k]), cmax = o2f(B.cmax[k]);
            if (!(cmax > cmin) || !isfinite(cmax - cmin)) continue;
            const int b = bin_of(centroid(r.box, k), cmin, (float)BINS / (cmax - cmin));
            if (uniform) {
                for (int a = 0; a < 3; ++a) { atomicMin(&s_min[k][b][a], bo[a]); atomicMax(&s_max[k][b][a], bo[3 + a]); }
                atomicAdd(&s_cnt[k][b], 1u);
            } else {
                NodeBins& W = bins[slot];
                for (int a = 0; a < 3; ++a) { atomicMin(&W.bmin[k][b][a], bo[a]); atomicMax(&W.bmax[k][b][a], bo[3 + a]); }
                atomicAdd(&W.cnt[k][b], 1u);
            }
        }
    }
    if (uniform) {
        __syncthreads();
        NodeBins& W = bins[first];
        for (uint32_t i = threadIdx.x; i < 3 * BINS; i += blockDim.x) {
            const uint32_t c = (&s_cnt[0][0])[i];
            if (!c) continue;
            const int k = i / BINS, b = i % BINS;
            atomicAdd(&W.cnt[k][b], c);
            for (int a = 0; a < 3; ++a) { atomicMin(&W.bmin[k][b][a], s_min[k][b][a]); atomicMax(&W.bmax[k][b][a], s_max[k][b][a]); }
        }
    }
}

// one warp per active node, lane = bin: the split, the children's boxes, the children's slots
__global__ void k_eval(const Active* __restrict__ act, uint32_t n_active, const NodeBins* __restrict__ bins, Split* split, Active* next, uint4* small,
                       Counts* cnt, uint32_t max_depth, Active* done) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (w >= n_active) return;
    const Active A = act[w];
    const NodeBins& B = bins[w];
    const uint32_t n = A.hi - A.lo;
    const uint32_t levels_left = max_depth - A.depth;
    const unsigned long long child_cap = levels_left >= 1 ? (levels_left - 1 >= 63 ? ~0ull : (1ull << (levels_left - 1))) : 1ull;
    double best_cost = INFINITY;
    uint32_t best_axis = 3u, best_bin = 0u, best_nl = 0u;
    float best_cmin = 0.0f, best_scale = 0.0f;
    float bl[6], br[6];
    for (int q = 0; q < 6; ++q) { bl[q] = 0.0f; br[q] = 0.0f; }
    for (int k = 0; k < 3; ++k) {
        const float cmin = o2f(B.cmin[k]), cmax = o2f(B.cmax[k]);
        if (!(cmax > cmin) || !isfinite(cmax - cmin)) continue;
        // inclusive prefix (bins 0..lane) and suffix (bins lane..31) unions
        float pre[6], suf[6];
        for (int a = 0; a < 3; ++a) { pre[a] = o2f(B.bmin[k][lane][a]); pre[3 + a] = o2f(B.bmax[k][lane][a]); }
        for (int q = 0; q < 6; ++q) suf[q] = pre[q];
        uint32_t cp = B.cnt[k][lane], cs = cp;
        for (int off = 1; off < 32; off <<= 1) {
            for (int q = 0; q < 6; ++q) {
                const float up = __shfl_up_sync(0xFFFFFFFFu, pre[q], off), dn = __shfl_down_sync(0xFFFFFFFFu, suf[q], off);
                if (lane >= (uint32_t)off) pre[q] = q < 3 ? fminf(pre[q], up) : fmaxf(pre[q], up);
                if (lane + off < 32u) suf[q] = q < 3 ? fminf(suf[q], dn) : fmaxf(suf[q], dn);
            }
            const uint32_t cu = __shfl_up_sync(0xFFFFFFFFu, cp, off), cd = __shfl_down_sync(0xFFFFFFFFu, cs, off);
            if (lane >= (uint32_t)off) cp += cu;
            if (lane + off < 32u) cs += cd;
        }
        // split after bin `lane`: left = pre, right = suffix of lane + 1
        float rs[6];
        for (int q = 0; q < 6; ++q) rs[q] = __shfl_down_sync(0xFFFFFFFFu, suf[q], 1);
        const uint32_t cr = __shfl_down_sync(0xFFFFFFFFu, cs, 1);
        double cost = INFINITY;
        if (lane < 31u && cp != 0u && cr != 0u && cp <= child_cap && cr <= child_cap) cost = half_area6(pre) * cp + half_area6(rs) * cr;
        // argmin over the lanes, the lower bin on ties (the host loop keeps the first strictly smaller cost)
        double c = cost;
        uint32_t arg = lane;
        for (int off = 16; off > 0; off >>= 1) {
            const double oc = __shfl_xor_sync(0xFFFFFFFFu, c, off);
            const uint32_t oa = __shfl_xor_sync(0xFFFFFFFFu, arg, off);
            if (oc < c || (oc == c && oa < arg)) { c = oc; arg = oa; }
        }
        if (c < best_cost) {
            best_cost = c; best_axis = (uint32_t)k; best_bin = arg; best_cmin = cmin; best_scale = (float)BINS / (cmax - cmin);
            best_nl = __shfl_sync(0xFFFFFFFFu, cp, arg);
            for (int q = 0; q < 6; ++q) { bl[q] = __shfl_sync(0xFFFFFFFFu, pre[q], arg); br[q] = __shfl_sync(0xFFFFFFFFu, rs[q], arg); }
        }
    }
    if (lane != 0u) return;
    Split S;
    S.cmin = best_cmin; S.scale = best_scale; S.axis = best_axis; S.bin = best_bin; S.pad = 0u;
    if (best_axis == 3u) best_nl = n / 2u;  // no useful plane (coincident centroids, or the depth cap binds): the first half of the positions;
                                            // k_fallback_boxes collects the two boxes
    const uint32_t mid = A.lo + best_nl;
    S.mid = mid;
    for (int q = 0; q < 6; ++q) { S.lbox[q] = bl[q]; S.rbox[q] = br[q]; }
    if (best_axis == 3u)
        for (int q = 0; q < 6; ++q) { S.lbox[q] = S.rbox[q] = q < 3 ? INFINITY : -INFINITY; }
    S.band_a = B.band_a; S.band_c = B.band_c;
    S.slot_l = S.slot_r = NONE;
    const uint32_t nl = best_nl, nr = n - best_nl;
    const uint32_t me_l = A.me + 1u, me_r = A.me + nl;
    if (nl > SMALL) { S.slot_l = atomicAdd(&cnt->n_next, 1u); next[S.slot_l] = Active{A.lo, mid, me_l, A.depth + 1u}; }
    else if (nl >= 2u) { small[atomicAdd(&cnt->n_small, 1u)] = make_uint4(A.lo, nl, me_l, A.depth + 1u); }
    if (nr > SMALL) { S.slot_r = atomicAdd(&cnt->n_next, 1u); next[S.slot_r] = Active{mid, A.hi, me_r, A.depth + 1u}; }
    else if (nr >= 2u) { small[atomicAdd(&cnt->n_small, 1u)] = make_uint4(mid, nr, me_r, A.depth + 1u); }
    split[w] = S;
    atomicMax(&cnt->depth, A.depth + 1u);
    done[atomicAdd(&cnt->n_done, 1u)] = Active{A.lo, A.hi, A.me, mid};  // (depth field reused: mid) for k_fix_leaves
}

__global__ void k_side(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ node_of, uint32_t n, const Ref* __restrict__ refs,
                       const Active* __restrict__ act, const Split* __restrict__ split, unsigned long long* flag) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t slot = node_of[p];
    unsigned long long f = 0ull;
    if (slot != NONE) {
        const Split& S = split[slot];
        if (S.axis == 3u) f = p < S.mid ? 1ull : 0ull;
        else f = (uint32_t)bin_of(centroid(refs[idx[p]].box, (int)S.axis), S.cmin, S.scale) <= S.bin ? 1ull : 0ull;
    }
    flag[p] = f;
}

// stable partition of every active node's range: lefts keep their order at [lo, mid), rights at [mid, hi)
__global__ void k_scatter(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ node_of, uint32_t n, const Active* __restrict__ act,
                          const Split* __restrict__ split, const unsigned long long* __restrict__ flag, const unsigned long long* __restrict__ pos,
                          uint32_t* idx_next, uint32_t* node_next) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t slot = node_of[p];
    if (slot == NONE) { idx_next[p] = idx[p]; node_next[p] = NONE; return; }
    const Active A = act[slot];
    const Split& S = split[slot];
    const uint32_t lefts_before = (uint32_t)(pos[p] - pos[A.lo]);
    const bool left = flag[p] != 0ull;
    const uint32_t q = left ? A.lo + lefts_before : S.mid + ((p - A.lo) - lefts_before);
    idx_next[q] = idx[p];
    node_next[q] = left ? S.slot_l : S.slot_r;
}

// median fallback only: the children's boxes (unions over the two halves)
__global__ void k_fallback_boxes(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ node_of, uint32_t n, const Ref* __restrict__ refs,
                                 Split* split, int* scratch /* 12 ordered ints per slot, cleared */) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t slot = node_of[p];
    if (slot == NONE || split[slot].axis != 3u) return;
    const Ref r = refs[idx[p]];
    int* s = scratch + 12 * (size_t)slot + (p < split[slot].mid ? 0 : 6);
    for (int a = 0; a < 3; ++a) { atomicMin(&s[a], f2o(r.box.mn[a])); atomicMax(&s[3 + a], f2o(r.box.mx[a])); }
}

// the node records of the level (children that are single references are patched by k_fix_leaves once positions are final)
__global__ void k_emit(const Active* __restrict__ act, uint32_t n_active, const Split* __restrict__ split, const int* __restrict__ scratch, NodeDev* nodes) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_active) return;
    const Active A = act[w];
    const Split& S = split[w];
    float l[6], r[6];
    for (int q = 0; q < 6; ++q) { l[q] = S.lbox[q]; r[q] = S.rbox[q]; }
    if (S.axis == 3u)
        for (int q = 0; q < 6; ++q) { l[q] = o2f(scratch[12 * (size_t)w + q]); r[q] = o2f(scratch[12 * (size_t)w + 6 + q]); }
    NodeDev nd;
    nd.q0 = make_float4(l[0], l[1], l[2], l[3]);
    nd.q1 = make_float4(l[4], l[5], r[0], r[1]);
    nd.q2 = make_float4(r[2], r[3], r[4], r[5]);
    const uint32_t margin = (bf16_up(__uint_as_float(S.band_a)) << 16) | bf16_up(__uint_as_float(S.band_c));
    nd.q3 = make_uint4(A.me + 1u, A.me + (S.mid - A.lo), S.axis == 3u ? 0u : S.axis, margin);
    nodes[A.me] = nd;
}
__global__ void k_clear_scratch(int* scratch, uint32_t n_active) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 12u * n_active) scratch[i] = (i % 6u) < 3u ? 0x7FFFFFFF : (int)0x80000000;
}

// children of one reference: the leaf reference of the element at that (final) position
__global__ void k_fix_leaves(const Active* __restrict__ done, uint32_t n_done, const uint32_t* __restrict__ idx, const Ref* __restrict__ refs,
                             const uint32_t* __restrict__ item_ref, NodeDev* nodes) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_done) return;
    const Active D = done[w];
    const uint32_t mid = D.depth;
    if (mid - D.lo == 1u) nodes[D.me].q3.x = item_ref[refs[idx[D.lo]].item];
    if (D.hi - mid == 1u) nodes[D.me].q3.y = item_ref[refs[idx[mid]].item];
}

// One warp per range of <= 32 references: the whole subtree.  Lane i holds the reference at position lo + i.
__global__ void __launch_bounds__(128) k_small(const uint4* __restrict__ tasks /* lo, n, me, depth */, uint32_t n_tasks, const uint32_t* __restrict__ idx,
                                               const Ref* __restrict__ refs, const float* __restrict__ item_a, const float* __restrict__ item_c,
                                               const uint32_t* __restrict__ item_ref, NodeDev* nodes, Counts* cnt, uint32_t max_depth) {
    __shared__ uint32_t s_stack[4][32][3];  // (start | count << 8, me, depth)
    __shared__ uint32_t s_perm[4][32];
    const uint32_t wib = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tasks) return;
    const uint4 info = tasks[t];
    const uint32_t lo = info.x, n0 = info.y;
    uint32_t ref_of_lane = 0u;
    float bx[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float cen[3] = {0.0f, 0.0f, 0.0f};
    float ba = 0.0f, bc = 0.0f;
    if (lane < n0) {
        const Ref r = refs[idx[lo + lane]];
        for (int a = 0; a < 3; ++a) { bx[a] = r.box.mn[a]; bx[3 + a] = r.box.mx[a]; cen[a] = centroid(r.box, a); }
        ba = item_a[r.item]; bc = item_c[r.item];
        ref_of_lane = item_ref[r.item];
    }
    uint32_t sp = 0u, deepest = 0u;
    if (lane == 0u) { s_stack[wib][0][0] = 0u | (n0 << 8); s_stack[wib][0][1] = info.z; s_stack[wib][0][2] = info.w; }
    sp = 1u;
    __syncwarp();
    while (sp > 0u) {
        --sp;
        const uint32_t e0 = s_stack[wib][sp][0], me = s_stack[wib][sp][1], depth = s_stack[wib][sp][2];
        __syncwarp();
        const uint32_t s = e0 & 0xFFu, m = e0 >> 8;
        const bool in = lane >= s && lane < s + m;
        const uint32_t levels_left = max_depth - depth;
        const uint32_t child_cap = levels_left >= 1u ? (levels_left - 1u >= 31u ? 0xFFFFFFFFu : (1u << (levels_left - 1u))) : 1u;
        deepest = max(deepest, depth + 1u);
        double best_cost = INFINITY;
        uint32_t best_mask = 0u;
        for (int k = 0; k < 3; ++k) {
            for (uint32_t j = 0; j < m; ++j) {
                const float c = __shfl_sync(0xFFFFFFFFu, cen[k], (int)(s + j));
                const bool left = in && cen[k] <= c;
                const uint32_t lm = __ballot_sync(0xFFFFFFFFu, left), rm = __ballot_sync(0xFFFFFFFFu, in && !left);
                const uint32_t nl = __popc(lm), nr = __popc(rm);
                if (nr == 0u || nl > child_cap || nr > child_cap) continue;  // (warp-uniform)
                float L[6], R[6];
                for (int q = 0; q < 6; ++q) {
                    const float e = q < 3 ? INFINITY : -INFINITY;
                    L[q] = left ? bx[q] : e;
                    R[q] = (in && !left) ? bx[q] : e;
                }
                for (int off = 16; off > 0; off >>= 1)
                    for (int q = 0; q < 6; ++q) {
                        const float ol = __shfl_xor_sync(0xFFFFFFFFu, L[q], off), orr = __shfl_xor_sync(0xFFFFFFFFu, R[q], off);
                        L[q] = q < 3 ? fminf(L[q], ol) : fmaxf(L[q], ol);
                        R[q] = q < 3 ? fminf(R[q], orr) : fmaxf(R[q], orr);
                    }
                const double cost = half_area6(L) * nl + half_area6(R) * nr;
                if (cost < best_cost) { best_cost = cost; best_mask = lm; }
            }
        }
        if (best_mask == 0u) {  // no plane separates them within the cap: the first half of the lanes
            const uint32_t h = m / 2u;
            best_mask = (h >= 32u ? 0xFFFFFFFFu : ((1u << h) - 1u)) << s;
        }
        const bool left = (best_mask >> lane) & 1u;
        const uint32_t right_mask = __ballot_sync(0xFFFFFFFFu, in && !left);
        const uint32_t nl = __popc(best_mask), nr = m - nl;
        // boxes and bands of the node
        float L[6], R[6];
        for (int q = 0; q < 6; ++q) {
            const float e = q < 3 ? INFINITY : -INFINITY;
            L[q] = (in && left) ? bx[q] : e;
            R[q] = (in && !left) ? bx[q] : e;
        }
        float na = in ? ba : 0.0f, nc = in ? bc : 0.0f;
        for (int off = 16; off > 0; off >>= 1) {
            for (int q = 0; q < 6; ++q) {
                const float ol = __shfl_xor_sync(0xFFFFFFFFu, L[q], off), orr = __shfl_xor_sync(0xFFFFFFFFu, R[q], off);
                L[q] = q < 3 ? fminf(L[q], ol) : fmaxf(L[q], ol);
                R[q] = q < 3 ? fminf(R[q], orr) : fmaxf(R[q], orr);
            }
            na = fmaxf(na, __shfl_xor_sync(0xFFFFFFFFu, na, off));
            nc = fmaxf(nc, __shfl_xor_sync(0xFFFFFFFFu, nc, off));
        }
        // stable partition of the lanes [s, s + m): lefts first
        uint32_t newpos = lane;
        if (in) newpos = left ? s + __popc(best_mask & ((1u << lane) - 1u)) : s + nl + __popc(right_mask & ((1u << lane) - 1u));
        s_perm[wib][newpos] = lane;
        __syncwarp();
        const int src = (int)s_perm[wib][lane];
        __syncwarp();
        for (int q = 0; q < 6; ++q) bx[q] = __shfl_sync(0xFFFFFFFFu, bx[q], src);
        for (int a = 0; a < 3; ++a) cen[a] = __shfl_sync(0xFFFFFFFFu, cen[a], src);
        ba = __shfl_sync(0xFFFFFFFFu, ba, src); bc = __shfl_sync(0xFFFFFFFFu, bc, src);
        ref_of_lane = __shfl_sync(0xFFFFFFFFu, ref_of_lane, src);
        // the record
        const uint32_t leaf_l = __shfl_sync(0xFFFFFFFFu, ref_of_lane, (int)s), leaf_r = __shfl_sync(0xFFFFFFFFu, ref_of_lane, (int)(s + nl));
        if (lane == 0u) {
            NodeDev nd;
            nd.q0 = make_float4(L[0], L[1], L[2], L[3]);
            nd.q1 = make_float4(L[4], L[5], R[0], R[1]);
            nd.q2 = make_float4(R[2], R[3], R[4], R[5]);
            nd.q3 = make_uint4(nl == 1u ? leaf_l : me + 1u, nr == 1u ? leaf_r : me + nl, 0u, (bf16_up(na) << 16) | bf16_up(nc));
            nodes[me] = nd;
            if (nr >= 2u) { s_stack[wib][sp][0] = (s + nl) | (nr << 8); s_stack[wib][sp][1] = me + nl; s_stack[wib][sp][2] = depth + 1u; }
            if (nl >= 2u) { const uint32_t o = sp + (nr >= 2u ? 1u : 0u); s_stack[wib][o][0] = s | (nl << 8); s_stack[wib][o][1] = me + 1u; s_stack[wib][o][2] = depth + 1u; }
        }
        sp += (nr >= 2u ? 1u : 0u) + (nl >= 2u ? 1u : 0u);
        __syncwarp();
    }
    if (lane == 0u) atomicMax(&cnt->depth, deepest);
}

struct Result {
    uint32_t n_nodes = 0, depth = 0, root = 0;
};

// d_refs: n > 32 references (device).  item_a / item_c: guard-band coefficients per ELEMENT, item_ref: leaf reference per element
// (device).  d_nodes: room for n - 1 records.  Synchronous (one small read-back per level sizes the next level's launches).
inline cudaError_t run(const Ref* d_refs, uint32_t n, const float* d_item_a, const float* d_item_c, const uint32_t* d_item_ref, uint32_t max_depth,
                       NodeDev* d_nodes, Result* out, cudaStream_t st) {
#define SB_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { if (base) rt_devcache::release(base); return e_; } } while (0)
    char* base = nullptr;
    if (n <= SMALL) return cudaErrorInvalidValue;
    uint32_t need = 0;
    while ((1ull << need) < n) ++need;
    max_depth = max_depth > need ? max_depth : need;
    const size_t cap = (size_t)n / (SMALL + 1) + 2;  // active ranges of one level: disjoint, each above SMALL references
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t scan_bytes = 0;
    SB_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n, st));
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align(bytes); return o; };
    const size_t o_idx0 = take((size_t)n * 4), o_idx1 = take((size_t)n * 4), o_nod0 = take((size_t)n * 4), o_nod1 = take((size_t)n * 4), o_flag = take((size_t)n * 8),
                 o_pos = take((size_t)n * 8), o_act0 = take(cap * sizeof(Active)), o_act1 = take(cap * sizeof(Active)), o_split = take(cap * sizeof(Split)),
                 o_bins = take(cap * sizeof(NodeBins)), o_scr = take(cap * 12 * 4), o_small = take(((size_t)n / 2 + 2) * sizeof(uint4)),
                 o_done = take((size_t)n * sizeof(Active)), o_cnt = take(sizeof(Counts)), o_scan = take(scan_bytes);
    SB_CU(rt_devcache::alloc(reinterpret_cast<void**>(&base), off));
    uint32_t* idx[2] = {reinterpret_cast<uint32_t*>(base + o_idx0), reinterpret_cast<uint32_t*>(base + o_idx1)};
    uint32_t* nod[2] = {reinterpret_cast<uint32_t*>(base + o_nod0), reinterpret_cast<uint32_t*>(base + o_nod1)};
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(base + o_flag);
    unsigned long long* pos = reinterpret_cast<unsigned long long*>(base + o_pos);
    Active* act[2] = {reinterpret_cast<Active*>(base + o_act0), reinterpret_cast<Active*>(base + o_act1)};
    Split* split = reinterpret_cast<Split*>(base + o_split);
    NodeBins* bins = reinterpret_cast<NodeBins*>(base + o_bins);
    int* scratch = reinterpret_cast<int*>(base + o_scr);
    uint4* small = reinterpret_cast<uint4*>(base + o_small);
    Active* done = reinterpret_cast<Active*>(base + o_done);
    Counts* cnt = reinterpret_cast<Counts*>(base + o_cnt);
    void* scan_tmp = base + o_scan;
    const int TB = 256;
    auto grid = [&](size_t items) { return (unsigned)((items + TB - 1) / TB); };

    SB_CU(cudaMemsetAsync(cnt, 0, sizeof(Counts), st));
    k_iota<<<grid(n), TB, 0, st>>>(idx[0], nod[0], n);
    const Active root{0u, n, 0u, 0u};
    SB_CU(cudaMemcpyAsync(act[0], &root, sizeof(Active), cudaMemcpyHostToDevice, st));
    uint32_t n_active = 1u;
    int cur = 0, levels = 0;
    Counts h{};
    while (n_active > 0u) {
        k_clear_bins<<<std::min<unsigned>(grid((size_t)n_active * (sizeof(NodeBins) / 4)), 148u * 8u), TB, 0, st>>>(bins, n_active);
        k_clear_scratch<<<grid(12 * (size_t)n_active), TB, 0, st>>>(scratch, n_active);
        k_bounds<<<grid(n), TB, 0, st>>>(idx[cur], nod[cur], n, d_refs, d_item_a, d_item_c, bins);
        k_bin<<<grid(n), 256, 0, st>>>(idx[cur], nod[cur], n, d_refs, bins);
        SB_CU(cudaMemsetAsync(&cnt->n_next, 0, 4, st));
        k_eval<<<grid((size_t)n_active * 32), TB, 0, st>>>(act[cur], n_active, bins, split, act[cur ^ 1], small, cnt, max_depth, done);
        k_side<<<grid(n), TB, 0, st>>>(idx[cur], nod[cur], n, d_refs, act[cur], split, flag);
        SB_CU(cudaGetLastError());
        SB_CU(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, pos, (int)n, st));
        k_fallback_boxes<<<grid(n), TB, 0, st>>>(idx[cur], nod[cur], n, d_refs, split, scratch);
        k_emit<<<grid(n_active), TB, 0, st>>>(act[cur], n_active, split, scratch, d_nodes);
        k_scatter<<<grid(n), TB, 0, st>>>(idx[cur], nod[cur], n, act[cur], split, flag, pos, idx[cur ^ 1], nod[cur ^ 1]);
        SB_CU(cudaGetLastError());
        SB_CU(cudaMemcpyAsync(&h, cnt, sizeof(Counts), cudaMemcpyDeviceToHost, st));
        SB_CU(cudaStreamSynchronize(st));
        n_active = h.n_next;
        cur ^= 1;
        if (++levels > 200 || n_active > cap) { rt_devcache::release(base); return cudaErrorUnknown; }
    }
    if (h.n_small) k_small<<<(h.n_small + 3u) / 4u, 128, 0, st>>>(small, h.n_small, idx[cur], d_refs, d_item_a, d_item_c, d_item_ref, d_nodes, cnt, max_depth);
    if (h.n_done) k_fix_leaves<<<grid(h.n_done), TB, 0, st>>>(done, h.n_done, idx[cur], d_refs, d_item_ref, d_nodes);
    SB_CU(cudaGetLastError());
    SB_CU(cudaMemcpyAsync(&h, cnt, sizeof(Counts), cudaMemcpyDeviceToHost, st));
    SB_CU(cudaStreamSynchronize(st));
    if (std::getenv("RT_PACK_LOG"))
        std::fprintf(stderr, "[rt_scene_create]   device SAH build: %d levels above %u references, %u warp-built subtrees, depth %u\n", levels, SMALL, h.n_small, h.depth);
    rt_devcache::release(base);
    out->n_nodes = n - 1u;
    out->depth = h.depth;
    out->root = 0u;
    return cudaSuccess;
#undef SB_CU
}

}  // namespace rt_sahbuild
