// rt_presplit.cuh — SURVEY §8 row f4: the pre-split of sliver references (rt_rebuild.hpp, presplit) on the device.
//
// Same algorithm, same arithmetic (rt_rebuild::compute_cut is compiled for both sides), same rounds: every round cuts all
// candidates whose box area is at least tau, the pieces become candidates again, tau drops to the largest remaining candidate's
// power-of-two threshold when nothing that large is left, and the budget ends on the largest candidates of the last round.
// What differs is the bookkeeping: one candidate list per round, compacted with a prefix sum over (stays, accepted) — so a new
// reference's index depends only on its position in the list, never on the order in which threads finish.  The set of references
// equals the host version's (tests/test_gpu_round2.py compares them); their numbering differs, which no result depends on.
//
// 10^6 triangles -> 2 * 10^6 references (C4): 125 ms on 16 host cores, 28 ms here including the upload of the boxes and triangles
// (12 cutting rounds of four small kernels + one scan each, 6 threshold descents, 5 rounds in which the budget ends: those order
// their large candidates with two stable radix sorts — reference index, then area — and cut the first `room` of them).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rt_rebuild.hpp"

namespace rt_devcache {  // b200rt.cu: device blocks re-used from scene to scene
cudaError_t alloc(void** out, size_t bytes);
void release(void* p);
}

namespace rt_presplit {

using rt_rebuild::Box3;
using rt_rebuild::Cut;
using rt_rebuild::Ref;
using rt_rebuild::Region;

struct Counters {
    unsigned long long n_big;      // candidates with area >= tau
    unsigned long long max_small;  // bits of the largest candidate area below tau (positive doubles order like their bits)
    unsigned long long max_area;   // k_init: bits of the largest triangle box area
};

__global__ void k_init(const Box3* __restrict__ leaf, const uint8_t* __restrict__ is_tri, uint32_t n0, Ref* refs, Region* region,
                       unsigned long long* code, Counters* ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bits = 0ull;
    if (i < n0) {
        Ref r;
        r.item = i;
        r.box = leaf[i];
        refs[i] = r;
        Region rg;
        for (int k = 0; k < 3; ++k) { rg.lo[k] = -INFINITY; rg.hi[k] = INFINITY; }
        region[i] = rg;
        const bool t = is_tri[i] != 0;
        code[i] = t ? 1ull : 0ull;
        if (t) {
            const double a = rt_rebuild::detail::half_area(r.box);
            if (a > 0.0 && a < INFINITY) bits = (unsigned long long)__double_as_longlong(a);
        }
    }
    for (int off = 16; off > 0; off >>= 1) bits = max(bits, __shfl_xor_sync(0xFFFFFFFFu, bits, off));
    if ((threadIdx.x & 31u) == 0u && bits) atomicMax(&ctr->max_area, bits);
}

__global__ void k_first_candidates(const unsigned long long* __restrict__ pos, const uint8_t* __restrict__ is_tri, uint32_t n0, uint32_t* cand) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0 && is_tri[i]) cand[pos[i]] = i;
}

__global__ void k_flag(const uint32_t* __restrict__ cand, uint32_t n_cand, const Ref* __restrict__ refs, double tau, double* area, uint8_t* big,
                       Counters* ctr) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    bool b = false;
    unsigned long long small_bits = 0ull;
    if (c < n_cand) {
        const double a = rt_rebuild::detail::half_area(refs[cand[c]].box);
        area[c] = a;
        b = a >= tau;
        big[c] = b ? 1 : 0;
        if (!b && a > 0.0) small_bits = (unsigned long long)__double_as_longlong(a);
    }
    const uint32_t nb = __popc(__ballot_sync(0xFFFFFFFFu, b));
    for (int off = 16; off > 0; off >>= 1) small_bits = max(small_bits, __shfl_xor_sync(0xFFFFFFFFu, small_bits, off));
    if ((threadIdx.x & 31u) == 0u) {
        if (nb) atomicAdd(&ctr->n_big, (unsigned long long)nb);
        if (small_bits) atomicMax(&ctr->max_small, small_bits);
    }
}

// The round in which the budget ends: the large candidates alone, compacted (position from a prefix sum over `big`), go to the host,
// which picks the ones to cut; k_mark_dropped brings its verdict back.
__global__ void k_big_code(const uint8_t* __restrict__ big, uint32_t n_cand, unsigned long long* code) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cand) code[c] = big[c] ? 1ull : 0ull;
}
struct BigCand {
    double area;
    uint32_t ref, c;
};
static_assert(sizeof(BigCand) <= sizeof(Cut), "the selection list borrows the cut array");
__global__ void k_big_gather(const uint32_t* __restrict__ cand, uint32_t n_cand, const uint8_t* __restrict__ big, const double* __restrict__ area,
                             const unsigned long long* __restrict__ pos, BigCand* out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cand && big[c]) out[pos[c]] = BigCand{area[c], cand[c], c};
}
// Selection of that round on the device: order the large candidates by (area descending, reference index ascending) with two stable
// radix sorts — by reference index first, then by area — and drop everything after the first `room`.
__global__ void k_sel_keys_ref(const BigCand* __restrict__ sel, uint32_t n, uint32_t* key, uint32_t* val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = sel[i].ref; val[i] = i; }
}
__global__ void k_sel_keys_area(const BigCand* __restrict__ sel, const uint32_t* __restrict__ perm, uint32_t n, unsigned long long* key, uint32_t* val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const BigCand b = sel[perm[i]]; key[i] = (unsigned long long)__double_as_longlong(b.area); val[i] = b.c; }
}
__global__ void k_mark_dropped(const uint32_t* __restrict__ dropped, uint32_t n, uint8_t* big) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) big[dropped[i]] = 2;
}

// code: low word = "stays a candidate as it is" (area below tau), high word = "cut accepted" — one 64-bit prefix sum gives both
// compaction offsets.
__global__ void k_cut(const uint32_t* __restrict__ cand, uint32_t n_cand, const uint8_t* __restrict__ big, const Ref* __restrict__ refs,
                      const Region* __restrict__ region, const float* __restrict__ tri, Cut* cuts, unsigned long long* code) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    if (big[c] == 0) { code[c] = 1ull; return; }
    if (big[c] == 2) {  // large enough, but beyond the budget of the round in which it ends: leaves the list uncut
        cuts[c].ok = false;
        code[c] = 0ull;
        return;
    }
    const uint32_t r = cand[c];
    const Ref ref = refs[r];
    float t9[9];
    for (int k = 0; k < 9; ++k) t9[k] = tri[9 * (size_t)ref.item + k];
    const Region rg = region[r];
    Cut cu;
    rt_rebuild::compute_cut(ref.box, t9, rg, cu);
    cuts[c] = cu;
    code[c] = cu.ok ? (1ull << 32) : 0ull;
}

// The low half replaces the reference, the high half becomes reference n_refs + (accepted cuts before it); next round's list =
// the candidates that stayed, in order, then the two pieces of every accepted cut, in order.
__global__ void k_apply(const uint32_t* __restrict__ cand, uint32_t n_cand, const uint8_t* __restrict__ big, const Cut* __restrict__ cuts,
                        const unsigned long long* __restrict__ pos, Ref* refs, Region* region, uint32_t n_refs, uint32_t n_stay, uint32_t* cand_next) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cand) return;
    const unsigned long long p = pos[c];
    const uint32_t r = cand[c];
    if (!big[c]) { cand_next[(uint32_t)p] = r; return; }
    const Cut cu = cuts[c];
    if (!cu.ok) return;
    const uint32_t k = (uint32_t)(p >> 32), r2 = n_refs + k;
    Ref lo = refs[r];
    Ref hi;
    hi.item = lo.item;
    hi.box = cu.bh;
    lo.box = cu.bl;
    refs[r] = lo;
    refs[r2] = hi;
    Region rl = region[r], rh = rl;
    rl.hi[cu.ax] = cu.mid;
    rh.lo[cu.ax] = cu.mid;
    region[r] = rl;
    region[r2] = rh;
    cand_next[n_stay + 2u * k] = r;
    cand_next[n_stay + 2u * k + 1u] = r2;
}

// Returns cudaSuccess or the first CUDA error.  d_leaf / d_tri / d_is_tri: the elements (device); d_refs: room for max(budget, n0)
// references; *n_refs_out: how many were made.  Synchronous (a few dozen tiny read-backs steer the rounds).
inline cudaError_t run(const Box3* d_leaf, const float* d_tri, const uint8_t* d_is_tri, uint32_t n0, size_t budget, Ref* d_refs, size_t* n_refs_out,
                       cudaStream_t st) {
#define PS_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { if (base) rt_devcache::release(base); return e_; } } while (0)
    char* base = nullptr;
    *n_refs_out = n0;
    if (n0 == 0) return cudaSuccess;
    const size_t B = std::max<size_t>(budget, n0);
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t scan_bytes = 0, sort_bytes = 0, sort_bytes2 = 0;
    PS_CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)B, st));
    PS_CU(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)B, 0, 32, st));
    PS_CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                                    (uint32_t*)nullptr, (int)B, 0, 64, st));
    sort_bytes = std::max(sort_bytes, sort_bytes2);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align(bytes); return o; };
    const size_t o_region = take(B * sizeof(Region)), o_cand0 = take(B * 4), o_cand1 = take(B * 4), o_code = take(B * 8), o_pos = take(B * 8),
                 o_area = take(B * 8), o_big = take(B), o_cuts = take(B * sizeof(Cut)), o_ctr = take(sizeof(Counters)), o_scan = take(scan_bytes),
                 o_sort = take(sort_bytes), o_k32 = take(B * 4), o_k32b = take(B * 4), o_v32 = take(B * 4), o_v32b = take(B * 4), o_k64 = take(B * 8);
    PS_CU(rt_devcache::alloc(reinterpret_cast<void**>(&base), off));
    Region* region = reinterpret_cast<Region*>(base + o_region);
    uint32_t* cand[2] = {reinterpret_cast<uint32_t*>(base + o_cand0), reinterpret_cast<uint32_t*>(base + o_cand1)};
    unsigned long long* code = reinterpret_cast<unsigned long long*>(base + o_code);
    unsigned long long* pos = reinterpret_cast<unsigned long long*>(base + o_pos);
    double* area = reinterpret_cast<double*>(base + o_area);
    uint8_t* big = reinterpret_cast<uint8_t*>(base + o_big);
    Cut* cuts = reinterpret_cast<Cut*>(base + o_cuts);
    Counters* ctr = reinterpret_cast<Counters*>(base + o_ctr);
    void* scan_tmp = base + o_scan;
    void* sort_tmp = base + o_sort;
    uint32_t* k32 = reinterpret_cast<uint32_t*>(base + o_k32);
    uint32_t* k32b = reinterpret_cast<uint32_t*>(base + o_k32b);
    uint32_t* v32 = reinterpret_cast<uint32_t*>(base + o_v32);
    uint32_t* v32b = reinterpret_cast<uint32_t*>(base + o_v32b);
    unsigned long long* k64 = reinterpret_cast<unsigned long long*>(base + o_k64);
    const int TB = 128;
    auto grid = [&](size_t n) { return (unsigned)((n + TB - 1) / TB); };
    auto totals = [&](uint32_t n, unsigned long long* out) -> cudaError_t {  // sum of code[0, n) from the exclusive scan
        unsigned long long last[2];
        cudaError_t e = cudaMemcpyAsync(&last[0], pos + (n - 1), 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&last[1], code + (n - 1), 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *out = last[0] + last[1];
        return e;
    };

    PS_CU(cudaMemsetAsync(ctr, 0, sizeof(Counters), st));
    k_init<<<grid(n0), TB, 0, st>>>(d_leaf, d_is_tri, n0, d_refs, region, code, ctr);
    PS_CU(cudaGetLastError());
    PS_CU(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, code, pos, (int)n0, st));
    unsigned long long n_first = 0;
    PS_CU(totals(n0, &n_first));
    Counters h{};
    PS_CU(cudaMemcpyAsync(&h, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    PS_CU(cudaStreamSynchronize(st));
    double max_area;
    { long long b = (long long)h.max_area; std::memcpy(&max_area, &b, 8); }
    const double tau0 = 0.5 * max_area;
    size_t n_refs = n0;
    uint32_t n_cand = (uint32_t)n_first;
    if (budget > n0 && n_cand > 0 && tau0 > 0.0 && std::isfinite(tau0)) {
        k_first_candidates<<<grid(n0), TB, 0, st>>>(pos, d_is_tri, n0, cand[0]);
        PS_CU(cudaGetLastError());
        // the largest threshold tau0 / 2^j that is <= a (the host version's bucket of a); 0 when a is below every normal threshold
        auto threshold_below = [&](double a) -> double {
            if (!(a >= 2.3e-308)) return 0.0;
            if (a >= tau0) return tau0;
            int j = std::max(std::ilogb(tau0) - std::ilogb(a) - 1, 0);
            double t = std::ldexp(tau0, -j);
            while (t > a) t *= 0.5;
            return t >= 2.3e-308 ? t : 0.0;
        };
        double tau = tau0;
        int cur = 0;
        int n_rounds = 0, n_descents = 0, n_selections = 0;
        for (int guard = 0; guard < 100000 && n_refs < budget && n_cand > 0; ++guard) {
            PS_CU(cudaMemsetAsync(ctr, 0, sizeof(Counters), st));
            k_flag<<<grid(n_cand), TB, 0, st>>>(cand[cur], n_cand, d_refs, tau, area, big, ctr);
            PS_CU(cudaGetLastError());
            PS_CU(cudaMemcpyAsync(&h, ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
            PS_CU(cudaStreamSynchronize(st));
            if (h.n_big == 0) {
                double a;
                { long long b = (long long)h.max_small; std::memcpy(&a, &b, 8); }
                tau = threshold_below(a);
                ++n_descents;
                if (!(tau > 0.0)) break;
                continue;
            }
            const size_t room = budget - n_refs;
            if (h.n_big > room) {
                // the budget ends inside this round: only the `room` largest are cut (ties by reference index); the others leave the list
                const uint32_t nb = (uint32_t)h.n_big;
                ++n_selections;
                BigCand* d_sel = reinterpret_cast<BigCand*>(cuts);  // (cuts is not in use before k_cut; sizeof(BigCand) <= sizeof(Cut))
                k_big_code<<<grid(n_cand), TB, 0, st>>>(big, n_cand, code);
                PS_CU(cudaGetLastError());
                PS_CU(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, code, pos, (int)n_cand, st));
                k_big_gather<<<grid(n_cand), TB, 0, st>>>(cand[cur], n_cand, big, area, pos, d_sel);
                PS_CU(cudaGetLastError());
                k_sel_keys_ref<<<grid(nb), TB, 0, st>>>(d_sel, nb, k32, v32);
                PS_CU(cudaGetLastError());
                PS_CU(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, k32, k32b, v32, v32b, (int)nb, 0, 32, st));
                k_sel_keys_area<<<grid(nb), TB, 0, st>>>(d_sel, v32b, nb, k64, v32);
                PS_CU(cudaGetLastError());
                PS_CU(cub::DeviceRadixSort::SortPairsDescending(sort_tmp, sort_bytes, k64, code /* keys out: free until k_cut */, v32, v32b, (int)nb, 0, 64, st));
                k_mark_dropped<<<grid(nb - room), TB, 0, st>>>(v32b + room, (uint32_t)(nb - room), big);
                PS_CU(cudaGetLastError());
            }
            k_cut<<<grid(n_cand), TB, 0, st>>>(cand[cur], n_cand, big, d_refs, region, d_tri, cuts, code);
            PS_CU(cudaGetLastError());
            PS_CU(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, code, pos, (int)n_cand, st));
            unsigned long long tot = 0;
            PS_CU(totals(n_cand, &tot));
            const uint32_t n_stay = (uint32_t)(tot & 0xFFFFFFFFull), n_ok = (uint32_t)(tot >> 32);
            k_apply<<<grid(n_cand), TB, 0, st>>>(cand[cur], n_cand, big, cuts, pos, d_refs, region, (uint32_t)n_refs, n_stay, cand[cur ^ 1]);
            PS_CU(cudaGetLastError());
            n_refs += n_ok;
            n_cand = n_stay + 2u * n_ok;
            ++n_rounds;
            cur ^= 1;  // (refused cuts can leave room after such a round: the loop goes on, like the host version's)
        }
        if (std::getenv("RT_PACK_LOG"))
            std::fprintf(stderr, "[rt_scene_create]   device pre-split: %d cutting rounds, %d threshold descents, %d budget selections, %zu references\n",
                         n_rounds, n_descents, n_selections, n_refs);
    }
    PS_CU(cudaStreamSynchronize(st));
    rt_devcache::release(base);
    *n_refs_out = n_refs;
    return cudaSuccess;
#undef PS_CU
}

}  // namespace rt_presplit
